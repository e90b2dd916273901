/*
 * ref_harness_fsk.c - flat, ctypes-friendly entry points around the UNMODIFIED reference FSK modem
 * (src/fsk.c: fsk_tx, fsk_rx) and the integer DDS (src/dds_int.c).  TEST INFRASTRUCTURE ONLY.
 *
 * Compiled INTO oracle/_ref/libspandsp_ref_{strict,fast}.so together with the reference's own
 * sources (taken in place from /root/reference/src; nothing is copied into this repository).
 */
#include "config.h"

#include <inttypes.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include <time.h>
#include <stdbool.h>

#include "spandsp/telephony.h"
#include "spandsp/alloc.h"
#include "spandsp/logging.h"
#include "spandsp/fast_convert.h"
#include "spandsp/saturated.h"
#include "spandsp/complex.h"
#include "spandsp/dds.h"
#include "spandsp/awgn.h"
#include "spandsp/async.h"
#include "spandsp/power_meter.h"
#include "spandsp/fsk.h"

#include "spandsp/private/logging.h"
#include "spandsp/private/power_meter.h"
#include "spandsp/private/fsk.h"
#include "spandsp/private/awgn.h"

#include "ref_harness.h"

#define EXPORT __attribute__((visibility("default")))

/* Bit source: PRBS x^23 + x^18 + 1, either raw (char_bits == 0) or as start-stop characters: `idle` mark bits,
   then start (0), char_bits data bits LSB first, an optional parity bit, stop (1). */
typedef struct
{
    uint32_t lfsr;
    int char_bits;
    int parity;
    int idle;
    int pos;            /* position inside the current character, -idle .. char_bits + 2 */
    int ones;
} fsk_src_t;

static int prbs_next(fsk_src_t *p)
{
    const int bit = ((p->lfsr >> 22) ^ (p->lfsr >> 17)) & 1;
    p->lfsr = ((p->lfsr << 1) | bit) & 0x7FFFFF;
    return bit;
}

static int fsk_src_get_bit(void *user)
{
    fsk_src_t *p = (fsk_src_t *) user;
    int bit;

    if (p->char_bits == 0)
        return prbs_next(p);
    if (p->pos < 0)
    {
        p->pos++;
        return 1;
    }
    if (p->pos == 0)
    {
        p->pos++;
        p->ones = 0;
        return 0;
    }
    if (p->pos <= p->char_bits)
    {
        bit = prbs_next(p);
        p->ones += bit;
        p->pos++;
        return bit;
    }
    if (p->pos == p->char_bits + 1  &&  p->parity != ASYNC_PARITY_NONE)
    {
        p->pos++;
        switch (p->parity)
        {
        case ASYNC_PARITY_EVEN:
            return p->ones & 1;
        case ASYNC_PARITY_ODD:
            return (p->ones & 1) ^ 1;
        case ASYNC_PARITY_MARK:
            return 1;
        default:
            return 0;
        }
    }
    /* stop bit, then `idle` mark bits before the next character */
    p->pos = -p->idle;
    return 1;
}

/* `lead` samples of silence, then an FSK burst `burst` samples long (to the end if < 0), AWGN over everything.
   level_dbm0 > 0 means "the spec's own tx level". */
EXPORT int ref_fsk_generate(int16_t *amp, int n, int spec, float level_dbm0, uint32_t lfsr_seed, int char_bits, int parity, int idle,
                            int lead, int burst, int noise_seed, float noise_dbm0)
{
    fsk_tx_state_t *tx;
    fsk_src_t src;
    awgn_state_t *noise;
    int pos;
    int len;
    int i;

    memset(amp, 0, sizeof(int16_t)*n);
    src.lfsr = (lfsr_seed & 0x7FFFFF)  ?  (lfsr_seed & 0x7FFFFF)  :  1;
    src.char_bits = char_bits;
    src.parity = parity;
    src.idle = idle;
    src.pos = -idle;
    src.ones = 0;
    tx = fsk_tx_init(NULL, &preset_fsk_specs[spec], fsk_src_get_bit, &src);
    if (tx == NULL)
        return -1;
    if (level_dbm0 <= 0.0f)
        fsk_tx_power(tx, level_dbm0);
    pos = (lead > n)  ?  n  :  lead;
    len = (burst < 0  ||  burst > n - pos)  ?  (n - pos)  :  burst;
    fsk_tx(tx, amp + pos, len);
    fsk_tx_free(tx);
    if (noise_dbm0 > -99.0f)
    {
        noise = awgn_init_dbm0(NULL, noise_seed, noise_dbm0);
        for (i = 0;  i < n;  i++)
            amp[i] = sat_add16(amp[i], awgn(noise));
        awgn_free(noise);
    }
    return pos + len;
}

typedef struct
{
    int16_t *out;
    int cap;
    int n;
} fsk_rec_t;

static void fsk_put_bit(void *user, int bit)
{
    fsk_rec_t *r = (fsk_rec_t *) user;
    if (r->n < r->cap)
        r->out[r->n] = (int16_t) bit;
    r->n++;
}

void ref_fsk_final(fsk_rx_state_t *rx, int32_t *final, int32_t *window)
{
    int j;
    int k;

    if (final)
    {
        final[0] = rx->baud_rate;
        final[1] = rx->framing_mode;
        final[2] = rx->data_bits;
        final[3] = rx->parity;
        final[4] = rx->stop_bits;
        final[5] = rx->total_data_bits;
        final[6] = rx->carrier_on_power;
        final[7] = rx->carrier_off_power;
        final[8] = rx->power.reading;
        final[9] = rx->last_sample;
        final[10] = rx->signal_present;
        final[11] = rx->phase_rate[0];
        final[12] = rx->phase_rate[1];
        final[13] = (int32_t) rx->phase_acc[0];
        final[14] = (int32_t) rx->phase_acc[1];
        final[15] = rx->correlation_span;
        final[16] = rx->dot[0].re;
        final[17] = rx->dot[0].im;
        final[18] = rx->dot[1].re;
        final[19] = rx->dot[1].im;
        final[20] = rx->buf_ptr;
        final[21] = rx->frame_pos;
        final[22] = rx->frame_in_progress;
        final[23] = rx->baud_phase;
        final[24] = rx->last_bit;
        final[25] = rx->scaling_shift;
        final[26] = rx->parity_errors;
        final[27] = rx->framing_errors;
    }
    if (window)
    {
        for (j = 0;  j < 2;  j++)
        {
            for (k = 0;  k < FSK_MAX_WINDOW_LEN;  k++)
            {
                window[(j*FSK_MAX_WINDOW_LEN + k)*2] = rx->window[j][k].re;
                window[(j*FSK_MAX_WINDOW_LEN + k)*2 + 1] = rx->window[j][k].im;
            }
        }
    }
}

/* One channel.  out[] receives what put_bit delivered, in order (bits / characters and negative SIG_STATUS_* codes).
   data_bits > 0: fsk_rx_set_frame_parameters(data_bits, parity, stop_bits) after init.
   restart_at >= 0: fsk_rx_restart(rx, &preset_fsk_specs[restart_spec], restart_mode) before the rx call that starts at
   the first chunk boundary >= restart_at.  fillin_at >= 0: fsk_rx_fillin(rx, fillin_len) INSTEAD of the rx call(s)
   covering [fillin_at, fillin_at + fillin_len) (chunk-aligned by the caller).
   final[28]: the receiver's integer state (order: sb_fsk_rx.cuh K_*); window[2*128*2]. */
EXPORT int ref_fsk_run(const int16_t *amp, int n, int chunk, int spec, int framing_mode, float cutoff,
                       int data_bits, int parity, int stop_bits,
                       int restart_at, int restart_spec, int restart_mode, int fillin_at, int fillin_len,
                       int16_t *out, int out_cap, int32_t *nout, int32_t *final, int32_t *window)
{
    fsk_rx_state_t *rx;
    fsk_rec_t rec;
    int pos;
    int len;

    rec.out = out;
    rec.cap = out_cap;
    rec.n = 0;
    rx = fsk_rx_init(NULL, &preset_fsk_specs[spec], framing_mode, fsk_put_bit, &rec);
    if (rx == NULL)
        return -1;
    if (cutoff > -99.0f)
        fsk_rx_set_signal_cutoff(rx, cutoff);
    if (data_bits > 0)
        fsk_rx_set_frame_parameters(rx, data_bits, parity, stop_bits);
    if (chunk <= 0)
        chunk = n;
    for (pos = 0;  pos < n;  pos += len)
    {
        if (restart_at >= 0  &&  pos >= restart_at)
        {
            fsk_rx_restart(rx, &preset_fsk_specs[restart_spec], restart_mode);
            restart_at = -1;
        }
        len = (n - pos < chunk)  ?  (n - pos)  :  chunk;
        if (fillin_at >= 0  &&  pos >= fillin_at  &&  pos < fillin_at + fillin_len)
            fsk_rx_fillin(rx, len);
        else
            fsk_rx(rx, amp + pos, len);
    }
    *nout = rec.n;
    ref_fsk_final(rx, final, window);
    fsk_rx_free(rx);
    return 0;
}

typedef struct
{
    const int16_t *amp;
    int64_t stride;
    int c0;
    int c1;
    int n;
    int chunk;
    int spec;
    int framing_mode;
} fsk_job_t;

static void *fsk_worker(void *arg)
{
    fsk_job_t *j = (fsk_job_t *) arg;
    int32_t nout;
    int16_t scratch[16];
    int c;

    for (c = j->c0;  c < j->c1;  c++)
        ref_fsk_run(j->amp + (int64_t) c*j->stride, j->n, j->chunk, j->spec, j->framing_mode, -100.0f, 0, 0, 0, -1, 0, 0, -1, 0, scratch, 0, &nout, NULL, NULL);
    return NULL;
}

/* Many channels on nthreads host threads; returns elapsed seconds (CPU baseline). */
EXPORT double ref_fsk_run_batch(const int16_t *amp, int64_t stride, int channels, int n, int chunk, int spec, int framing_mode, int nthreads)
{
    pthread_t *th;
    fsk_job_t *jobs;
    struct timespec t0;
    struct timespec t1;
    int i;

    if (nthreads < 1)
        nthreads = 1;
    if (nthreads > channels)
        nthreads = channels;
    th = (pthread_t *) malloc(sizeof(pthread_t)*nthreads);
    jobs = (fsk_job_t *) malloc(sizeof(fsk_job_t)*nthreads);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (i = 0;  i < nthreads;  i++)
    {
        jobs[i].amp = amp;
        jobs[i].stride = stride;
        jobs[i].c0 = (int) ((int64_t) channels*i/nthreads);
        jobs[i].c1 = (int) ((int64_t) channels*(i + 1)/nthreads);
        jobs[i].n = n;
        jobs[i].chunk = chunk;
        jobs[i].spec = spec;
        jobs[i].framing_mode = framing_mode;
        if (nthreads == 1)
            fsk_worker(&jobs[i]);
        else
            pthread_create(&th[i], NULL, fsk_worker, &jobs[i]);
    }
    if (nthreads > 1)
    {
        for (i = 0;  i < nthreads;  i++)
            pthread_join(th[i], NULL);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    free(jobs);
    return (double) (t1.tv_sec - t0.tv_sec) + 1.0e-9*(double) (t1.tv_nsec - t0.tv_nsec);
}

/* The integer DDS quarter wave as the reference's dds_lookup() sees it (257 entries), the presets' derived
   constants ({rate0, rate1, on_power, off_power} per preset), and the preset table itself
   ({freq_zero, freq_one, tx_level, min_level, baud_rate} per preset). */
EXPORT void ref_fsk_tables(int16_t *sine, int32_t *derived, int32_t *presets)
{
    fsk_rx_state_t *rx;
    int i;

    for (i = 0;  i <= 256;  i++)
        sine[i] = dds_lookup((uint32_t) i << 22);
    for (i = 0;  i <= FSK_V21CH1_110;  i++)
    {
        rx = fsk_rx_init(NULL, &preset_fsk_specs[i], FSK_FRAME_MODE_ASYNC, NULL, NULL);
        derived[4*i] = rx->phase_rate[0];
        derived[4*i + 1] = rx->phase_rate[1];
        derived[4*i + 2] = rx->carrier_on_power;
        derived[4*i + 3] = rx->carrier_off_power;
        fsk_rx_free(rx);
        presets[5*i] = preset_fsk_specs[i].freq_zero;
        presets[5*i + 1] = preset_fsk_specs[i].freq_one;
        presets[5*i + 2] = preset_fsk_specs[i].tx_level;
        presets[5*i + 3] = preset_fsk_specs[i].min_level;
        presets[5*i + 4] = preset_fsk_specs[i].baud_rate;
    }
}
