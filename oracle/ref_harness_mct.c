/*
 * ref_harness_mct.c - flat, ctypes-friendly entry points around the UNMODIFIED reference modem connect tone
 * generator and detector (src/modem_connect_tones.c) and, for the FAX preamble, the V.21 channel 2 FSK
 * transmitter (src/fsk.c).  TEST INFRASTRUCTURE ONLY.
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
#include "spandsp/tone_detect.h"
#include "spandsp/tone_generate.h"
#include "spandsp/super_tone_rx.h"
#include "spandsp/modem_connect_tones.h"

#include "spandsp/private/logging.h"
#include "spandsp/private/power_meter.h"
#include "spandsp/private/fsk.h"
#include "spandsp/private/awgn.h"
#include "spandsp/private/modem_connect_tones.h"

#include "ref_harness.h"

#define EXPORT __attribute__((visibility("default")))

/* Bit source of the preamble generator: `flags` HDLC flag octets (01111110, the same in either bit order), then
   PRBS x^23 + x^18 + 1 data (what follows a real preamble is a stuffed frame; for the detector any non-flag
   pattern is "the body"). */
typedef struct
{
    uint32_t lfsr;
    int flags;
    int pos;
} mct_src_t;

static int mct_src_get_bit(void *user)
{
    mct_src_t *p = (mct_src_t *) user;
    int bit;

    if (p->pos < 8*p->flags)
    {
        bit = (0x7E >> (p->pos & 7)) & 1;
        p->pos++;
        return bit;
    }
    bit = ((p->lfsr >> 22) ^ (p->lfsr >> 17)) & 1;
    p->lfsr = ((p->lfsr << 1) | bit) & 0x7FFFFF;
    return bit;
}

/* ADDS a signal to amp[0..n): `lead` untouched samples, then `burst` samples (to the end if < 0) of
   - tone_type 1..5, 8, 9: modem_connect_tones_tx() output.  freq > 0 replaces the tone frequency, level_dbm0 <= 0
     the level, mod_freq > 0 the AM frequency (as tests/modem_connect_tones_tests.c does for its sweeps, by writing
     the transmitter's phase rate and level fields);
   - tone_type 6: V.21 channel 2 FSK carrying `flags` HDLC flags then PRBS data (level_dbm0 <= 0 sets the level);
   then AWGN (if noise_dbm0 > -99) over all n samples.  Calling it several times on one buffer composes scenarios. */
EXPORT int ref_mct_generate(int16_t *amp, int n, int tone_type, float freq, float level_dbm0, float mod_freq,
                            int lead, int burst, int flags, uint32_t lfsr_seed, int noise_seed, float noise_dbm0)
{
    modem_connect_tones_tx_state_t *tx;
    fsk_tx_state_t *ftx;
    awgn_state_t *noise;
    mct_src_t src;
    int16_t *tmp;
    int pos;
    int len;
    int got;
    int i;

    pos = (lead > n)  ?  n  :  lead;
    len = (burst < 0  ||  burst > n - pos)  ?  (n - pos)  :  burst;
    tmp = (int16_t *) calloc((len > 0)  ?  len  :  1, sizeof(int16_t));
    if (tone_type == MODEM_CONNECT_TONES_FAX_PREAMBLE)
    {
        src.lfsr = (lfsr_seed & 0x7FFFFF)  ?  (lfsr_seed & 0x7FFFFF)  :  1;
        src.flags = flags;
        src.pos = 0;
        if ((ftx = fsk_tx_init(NULL, &preset_fsk_specs[FSK_V21CH2], mct_src_get_bit, &src)) == NULL)
        {
            free(tmp);
            return -1;
        }
        if (level_dbm0 <= 0.0f)
            fsk_tx_power(ftx, level_dbm0);
        fsk_tx(ftx, tmp, len);
        fsk_tx_free(ftx);
    }
    else if (tone_type != MODEM_CONNECT_TONES_NONE)
    {
        if ((tx = modem_connect_tones_tx_init(NULL, tone_type)) == NULL)
        {
            free(tmp);
            return -1;
        }
        if (freq > 0.0f)
            tx->tone_phase_rate = dds_phase_rate(freq);
        if (level_dbm0 <= 0.0f)
        {
            tx->level = dds_scaling_dbm0(level_dbm0);
            if (tx->mod_level)
                tx->mod_level = tx->level/5;
        }
        if (mod_freq > 0.0f  &&  tx->mod_phase_rate)
            tx->mod_phase_rate = dds_phase_rate(mod_freq);
        for (i = 0;  i < len;  i += got)
        {
            if ((got = modem_connect_tones_tx(tx, tmp + i, len - i)) <= 0)
                break;
        }
        modem_connect_tones_tx_free(tx);
    }
    for (i = 0;  i < len;  i++)
        amp[pos + i] = sat_add16(amp[pos + i], tmp[i]);
    free(tmp);
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
    int32_t *ev;            /* {index of the rx call, tone, level} */
    int cap;
    int n;
    int call;
} mct_rec_t;

static void mct_report(void *user, int tone, int level, int delay)
{
    mct_rec_t *r = (mct_rec_t *) user;
    if (r->n < r->cap)
    {
        r->ev[3*r->n] = r->call;
        r->ev[3*r->n + 1] = tone;
        r->ev[3*r->n + 2] = level;
    }
    r->n++;
}

static int32_t fbits(float f)
{
    int32_t v;
    memcpy(&v, &f, 4);
    return v;
}

/* One channel through modem_connect_tones_rx() in `chunk`-sample calls.
   use_callback != 0: ev[] receives {call index, tone, level} per tone_callback.
   use_callback == 0: no callback; modem_connect_tones_rx_get() is polled after every call and ev[] receives
   {call index, hit, 0} for every non-zero hit.
   final[16]: tone_type, notch_level, channel_level, am_level, tone_present, tone_on, tone_cycle_duration, good_cycles,
   raw_bit_stream, num_bits, flags_seen, framing_ok_announced, znotch_1, znotch_2, z15hz_1, z15hz_2 (floats as bits)
   - the order of sb_mct_rx.cuh's M_* fields; fsk_final[28] as ref_fsk_run(). */
extern void ref_fsk_final(fsk_rx_state_t *rx, int32_t *final, int32_t *window);

static int mct_run_core(const int16_t *amp, int n, int chunk, const int32_t *lens, int ncalls, int tone_type, int use_callback,
                        int32_t *ev, int ev_cap, int32_t *nev, int32_t *final, int32_t *fsk_final)
{
    modem_connect_tones_rx_state_t *rx;
    mct_rec_t rec;
    int pos;
    int len;
    int hit;

    rec.ev = ev;
    rec.cap = ev_cap;
    rec.n = 0;
    rec.call = 0;
    rx = modem_connect_tones_rx_init(NULL, tone_type, (use_callback)  ?  mct_report  :  NULL, &rec);
    if (rx == NULL)
        return -1;
    if (chunk <= 0)
        chunk = n;
    for (pos = 0;  (lens)  ?  (rec.call < ncalls)  :  (pos < n);  pos += len)
    {
        if (lens)
            len = lens[rec.call];
        else
            len = (n - pos < chunk)  ?  (n - pos)  :  chunk;
        modem_connect_tones_rx(rx, amp + pos, len);
        if (!use_callback)
        {
            if ((hit = modem_connect_tones_rx_get(rx)) != MODEM_CONNECT_TONES_NONE)
                mct_report(&rec, hit, 0, 0);
        }
        rec.call++;
    }
    *nev = rec.n;
    if (final)
    {
        final[0] = rx->tone_type;
        final[1] = rx->notch_level;
        final[2] = rx->channel_level;
        final[3] = rx->am_level;
        final[4] = rx->tone_present;
        final[5] = rx->tone_on;
        final[6] = rx->tone_cycle_duration;
        final[7] = rx->good_cycles;
        final[8] = (int32_t) rx->raw_bit_stream;
        final[9] = rx->num_bits;
        final[10] = rx->flags_seen;
        final[11] = rx->framing_ok_announced;
        final[12] = fbits(rx->znotch_1);
        final[13] = fbits(rx->znotch_2);
        final[14] = fbits(rx->z15hz_1);
        final[15] = fbits(rx->z15hz_2);
    }
    if (fsk_final)
    {
        if (rx->tone_type == MODEM_CONNECT_TONES_FAX_PREAMBLE  ||  rx->tone_type == MODEM_CONNECT_TONES_FAX_CED_OR_PREAMBLE)
            ref_fsk_final(&rx->v21rx, fsk_final, NULL);
        else
            memset(fsk_final, 0, sizeof(int32_t)*28);
    }
    modem_connect_tones_rx_free(rx);
    return 0;
}

EXPORT int ref_mct_run(const int16_t *amp, int n, int chunk, int tone_type, int use_callback,
                       int32_t *ev, int ev_cap, int32_t *nev, int32_t *final, int32_t *fsk_final)
{
    return mct_run_core(amp, n, chunk, NULL, 0, tone_type, use_callback, ev, ev_cap, nev, final, fsk_final);
}

/* The same with an explicit list of call lengths (sum = the samples available), callback mode. */
EXPORT int ref_mct_run_calls(const int16_t *amp, const int32_t *lens, int ncalls, int tone_type,
                             int32_t *ev, int ev_cap, int32_t *nev, int32_t *final, int32_t *fsk_final)
{
    return mct_run_core(amp, 0, 0, lens, ncalls, tone_type, 1, ev, ev_cap, nev, final, fsk_final);
}

typedef struct
{
    const int16_t *amp;
    int64_t stride;
    int c0;
    int c1;
    int n;
    int chunk;
    int tone_type;
} mct_job_t;

static void *mct_worker(void *arg)
{
    mct_job_t *j = (mct_job_t *) arg;
    int32_t nev;
    int32_t scratch[3];
    int c;

    for (c = j->c0;  c < j->c1;  c++)
        ref_mct_run(j->amp + (int64_t) c*j->stride, j->n, j->chunk, j->tone_type, 1, scratch, 0, &nev, NULL, NULL);
    return NULL;
}

/* Many channels on nthreads host threads; returns elapsed seconds (CPU baseline). */
EXPORT double ref_mct_run_batch(const int16_t *amp, int64_t stride, int channels, int n, int chunk, int tone_type, int nthreads)
{
    pthread_t *th;
    mct_job_t *jobs;
    struct timespec t0;
    struct timespec t1;
    int i;

    if (nthreads < 1)
        nthreads = 1;
    if (nthreads > channels)
        nthreads = channels;
    th = (pthread_t *) malloc(sizeof(pthread_t)*nthreads);
    jobs = (mct_job_t *) malloc(sizeof(mct_job_t)*nthreads);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (i = 0;  i < nthreads;  i++)
    {
        jobs[i].amp = amp;
        jobs[i].stride = stride;
        jobs[i].c0 = (int) ((int64_t) channels*i/nthreads);
        jobs[i].c1 = (int) ((int64_t) channels*(i + 1)/nthreads);
        jobs[i].n = n;
        jobs[i].chunk = chunk;
        jobs[i].tone_type = tone_type;
        if (nthreads == 1)
            mct_worker(&jobs[i]);
        else
            pthread_create(&th[i], NULL, mct_worker, &jobs[i]);
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
