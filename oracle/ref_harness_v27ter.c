/*
 * ref_harness_v27ter.c - flat, ctypes-friendly entry points around the UNMODIFIED reference V.27ter
 * transmitter and receiver (src/v27ter_tx.c, src/v27ter_rx.c).  TEST INFRASTRUCTURE ONLY.
 *
 * Compiled INTO oracle/_ref/libspandsp_ref_{strict,fast}.so together with the reference's own
 * sources (taken in place from /root/reference/src; nothing is copied into this repository).
 * A separate translation unit from ref_harness.c because the generated V.17 and V.29 headers
 * define the same static names (rx_pulseshaper_re, godard_desc, ...).
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
#include "spandsp/godard.h"
#include "spandsp/v29rx.h"
#include "spandsp/v27ter_rx.h"
#include "spandsp/v27ter_tx.h"

#include "spandsp/private/logging.h"
#include "spandsp/private/power_meter.h"
#include "spandsp/private/godard.h"
#include "spandsp/private/v27ter_rx.h"
#include "spandsp/private/awgn.h"

#include "ref_harness.h"

/* The tables the V.27ter receiver is built on (static const in the generated headers). */
#include "v27ter_rx_4800_rrc.h"
#include "v27ter_rx_2400_rrc.h"

#define EXPORT __attribute__((visibility("default")))

typedef struct
{
    uint32_t lfsr;
} prbs27_t;

static int prbs27_get_bit(void *user)
{
    prbs27_t *p = (prbs27_t *) user;
    /* x^23 + x^18 + 1 */
    const int bit = ((p->lfsr >> 22) ^ (p->lfsr >> 17)) & 1;
    p->lfsr = ((p->lfsr << 1) | bit) & 0x7FFFFF;
    return bit;
}

static void add_awgn27(int16_t *amp, int n, int seed, float level_dbm0)
{
    awgn_state_t *noise;
    int i;

    if (level_dbm0 <= -99.0f)
        return;
    noise = awgn_init_dbm0(NULL, seed, level_dbm0);
    for (i = 0;  i < n;  i++)
        amp[i] = sat_add16(amp[i], awgn(noise));
    awgn_free(noise);
}

/* `lead` samples of silence, a V.27ter burst of PRBS data `burst1` samples long (cut, not shut down), then -
   if burst2 > 0 - `gap` samples of silence and a second burst of `burst2` samples; silence to the end of the
   buffer; AWGN over everything. */
EXPORT int ref_v27ter_generate(int16_t *amp, int n, int bit_rate, int tep, float power_dbm0, uint32_t lfsr_seed,
                               int lead, int burst1, int gap, int burst2, int noise_seed, float noise_dbm0)
{
    v27ter_tx_state_t *tx;
    prbs27_t prbs;
    int pos;
    int len;

    prbs.lfsr = (lfsr_seed & 0x7FFFFF)  ?  (lfsr_seed & 0x7FFFFF)  :  1;
    memset(amp, 0, sizeof(int16_t)*n);
    pos = (lead > n)  ?  n  :  lead;
    tx = v27ter_tx_init(NULL, bit_rate, tep != 0, prbs27_get_bit, &prbs);
    if (tx == NULL)
        return -1;
    v27ter_tx_power(tx, power_dbm0);
    len = (burst1 < 0  ||  burst1 > n - pos)  ?  (n - pos)  :  burst1;
    v27ter_tx(tx, amp + pos, len);
    pos += len;
    if (burst2 > 0)
    {
        pos += gap;
        if (pos < n)
        {
            v27ter_tx_restart(tx, bit_rate, tep != 0);
            len = (burst2 > n - pos)  ?  (n - pos)  :  burst2;
            v27ter_tx(tx, amp + pos, len);
            pos += len;
        }
    }
    v27ter_tx_free(tx);
    add_awgn27(amp, n, noise_seed, noise_dbm0);
    return pos;
}

typedef struct
{
    int8_t *bits;
    int cap;
    int n;
    ref_v29_sym_t *syms;
    int sym_cap;
    int nsyms;
} v27_rec_t;

static void v27_put_bit(void *user, int bit)
{
    v27_rec_t *r = (v27_rec_t *) user;
    if (r->n < r->cap)
        r->bits[r->n] = (int8_t) bit;
    r->n++;
}

static void v27_qam(void *user, const complexf_t *z, const complexf_t *target, int symbol)
{
    v27_rec_t *r = (v27_rec_t *) user;
    if (r->nsyms < r->sym_cap)
    {
        /* Gardner hops are reported with NULL pointers and the integrator value (src/v27ter_rx.c:517-518):
           recorded as a NaN symbol carrying that value */
        r->syms[r->nsyms].re = (z)  ?  z->re  :  NAN;
        r->syms[r->nsyms].im = (z)  ?  z->im  :  NAN;
        r->syms[r->nsyms].tre = (target)  ?  target->re  :  NAN;
        r->syms[r->nsyms].tim = (target)  ?  target->im  :  NAN;
        r->syms[r->nsyms].state = symbol;
    }
    r->nsyms++;
}

/* One channel.  bits[] receives what put_bit delivered, in order (0/1 and negative SIG_STATUS_* codes).
   If restart_at >= 0, v27ter_rx_restart(rx, bit_rate, restart_short != 0) is called before the rx call that
   starts at the first chunk boundary >= restart_at (the way an application re-arms the receiver for
   the next page).
   final[]: {training_stage, carrier_phase_rate, eq_put_step, signal_present, agc_scaling bits,
             total_baud_timing_correction, constellation_state, carrier_phase, gardner_integrate, gardner_step} */
EXPORT int ref_v27ter_run(const int16_t *amp, int n, int chunk, int bit_rate, float cutoff, int want_qam,
                       int restart_at, int restart_short,
                       int8_t *bits, int bits_cap, int32_t *nbits,
                       ref_v29_sym_t *syms, int sym_cap, int32_t *nsyms,
                       float *eq_coeff, int32_t *final)
{
    v27ter_rx_state_t *rx;
    v27_rec_t rec;
    int pos;
    int len;
    int i;

    rec.bits = bits;
    rec.cap = bits_cap;
    rec.n = 0;
    rec.syms = syms;
    rec.sym_cap = sym_cap;
    rec.nsyms = 0;
    rx = v27ter_rx_init(NULL, bit_rate, v27_put_bit, &rec);
    if (rx == NULL)
        return -1;
    if (cutoff > -99.0f)
        v27ter_rx_set_signal_cutoff(rx, cutoff);
    if (want_qam)
        v27ter_rx_set_qam_report_handler(rx, v27_qam, &rec);
    for (pos = 0;  pos < n;  pos += len)
    {
        if (restart_at >= 0  &&  pos >= restart_at)
        {
            v27ter_rx_restart(rx, bit_rate, restart_short != 0);
            restart_at = -1;
        }
        len = (n - pos < chunk)  ?  (n - pos)  :  chunk;
        v27ter_rx(rx, amp + pos, len);
    }
    *nbits = rec.n;
    *nsyms = rec.nsyms;
    if (eq_coeff)
    {
        for (i = 0;  i < V27TER_EQUALIZER_LEN;  i++)
        {
            eq_coeff[2*i] = rx->eq_coeff[i].re;
            eq_coeff[2*i + 1] = rx->eq_coeff[i].im;
        }
    }
    if (final)
    {
        final[0] = rx->training_stage;
        final[1] = rx->carrier_phase_rate;
        final[2] = rx->eq_put_step;
        final[3] = rx->signal_present;
        memcpy(&final[4], &rx->agc_scaling, 4);
        final[5] = rx->total_baud_timing_correction;
        final[6] = rx->constellation_state;
        final[7] = (int32_t) rx->carrier_phase;
        final[8] = rx->gardner_integrate;
        final[9] = rx->gardner_step;
    }
    v27ter_rx_free(rx);
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
    int bit_rate;
    float cutoff;
} v27_job_t;

static void *v27_worker(void *arg)
{
    v27_job_t *j = (v27_job_t *) arg;
    int32_t nsyms;
    int32_t nb;
    int8_t scratch[16];
    int c;

    for (c = j->c0;  c < j->c1;  c++)
        ref_v27ter_run(j->amp + (int64_t) c*j->stride, j->n, j->chunk, j->bit_rate, j->cutoff, 0, -1, 0, scratch, 0, &nb, NULL, 0, &nsyms, NULL, NULL);
    return NULL;
}

/* Many channels on nthreads host threads; returns elapsed seconds (CPU baseline). */
EXPORT double ref_v27ter_run_batch(const int16_t *amp, int64_t stride, int channels, int n, int chunk, int bit_rate, float cutoff, int nthreads)
{
    pthread_t *th;
    v27_job_t *jobs;
    struct timespec t0;
    struct timespec t1;
    int i;

    if (nthreads < 1)
        nthreads = 1;
    if (nthreads > channels)
        nthreads = channels;
    th = (pthread_t *) malloc(sizeof(pthread_t)*nthreads);
    jobs = (v27_job_t *) malloc(sizeof(v27_job_t)*nthreads);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (i = 0;  i < nthreads;  i++)
    {
        jobs[i].amp = amp;
        jobs[i].stride = stride;
        jobs[i].c0 = (int) ((int64_t) channels*i/nthreads);
        jobs[i].c1 = (int) ((int64_t) channels*(i + 1)/nthreads);
        jobs[i].n = n;
        jobs[i].chunk = chunk;
        jobs[i].bit_rate = bit_rate;
        jobs[i].cutoff = cutoff;
        if (nthreads == 1)
            v27_worker(&jobs[i]);
        else
            pthread_create(&th[i], NULL, v27_worker, &jobs[i]);
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

/* The constant tables the V.27ter receiver uses, as the reference build sees them.
   rrc4800_*: [8][27]; rrc2400_*: [12][27]; ints: 8. */
EXPORT void ref_v27ter_tables(float *rrc4800_re, float *rrc4800_im, float *rrc2400_re, float *rrc2400_im, int32_t *ints)
{
    int i;
    int j;

    for (i = 0;  i < RX_PULSESHAPER_4800_COEFF_SETS;  i++)
    {
        for (j = 0;  j < 27;  j++)
        {
            rrc4800_re[i*27 + j] = rx_pulseshaper_4800_re[i][j];
            rrc4800_im[i*27 + j] = rx_pulseshaper_4800_im[i][j];
        }
    }
    for (i = 0;  i < RX_PULSESHAPER_2400_COEFF_SETS;  i++)
    {
        for (j = 0;  j < 27;  j++)
        {
            rrc2400_re[i*27 + j] = rx_pulseshaper_2400_re[i][j];
            rrc2400_im[i*27 + j] = rx_pulseshaper_2400_im[i][j];
        }
    }
    ints[0] = RX_PULSESHAPER_4800_COEFF_SETS;
    ints[1] = RX_PULSESHAPER_2400_COEFF_SETS;
    ints[2] = DDS_PHASE_RATE(1800.0f);
    ints[3] = DDS_PHASE_RATE(1800.0f - 20.0f);
    ints[4] = DDS_PHASE_RATE(1800.0f + 20.0f);
    ints[5] = DDS_PHASE(45.0f);
    ints[6] = DDS_PHASE(-45.0f);
    ints[7] = DDS_PHASE(180.0f);
}
