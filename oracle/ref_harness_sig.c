/*
 * ref_harness_sig.c - flat, ctypes-friendly entry points around the UNMODIFIED reference in-band signalling tone
 * generator and receiver (src/sig_tone.c).  TEST INFRASTRUCTURE ONLY.
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
#include <stdbool.h>

#include "spandsp/telephony.h"
#include "spandsp/alloc.h"
#include "spandsp/logging.h"
#include "spandsp/fast_convert.h"
#include "spandsp/saturated.h"
#include "spandsp/complex.h"
#include "spandsp/dds.h"
#include "spandsp/awgn.h"
#include "spandsp/power_meter.h"
#include "spandsp/tone_detect.h"
#include "spandsp/super_tone_rx.h"
#include "spandsp/sig_tone.h"

#include "spandsp/private/power_meter.h"
#include "spandsp/private/sig_tone.h"

#define EXPORT __attribute__((visibility("default")))

static void tx_update(void *user, int what, int level, int duration)
{
}

/* sig_tone_tx() driven by a script: steps[] = {mode bits, duration in samples} pairs; each step is
   sig_tone_tx_set_mode(mode, 0) followed by `duration` samples of sig_tone_tx() ADDED to amp (so that speech-like or
   noise backgrounds can be put under the tones).  tone_db > -99: the tone levels are replaced (both the high and the
   low one) as tests/sig_tone_tests.c does by writing tone_scaling[][].  freq_offset shifts both tones (Hz).  Then AWGN
   (noise_dbm0 > -99) over everything.  Returns the samples produced. */
EXPORT int ref_sig_generate(int16_t *amp, int n, int tone_type, const int32_t *steps, int nsteps, float tone_db, float freq_offset,
                            int noise_seed, float noise_dbm0)
{
    sig_tone_tx_state_t *tx;
    awgn_state_t *noise;
    int16_t *tmp;
    int pos;
    int len;
    int i;
    int k;

    if ((tx = sig_tone_tx_init(NULL, tone_type, tx_update, NULL)) == NULL)
        return -1;
    if (tone_db > -99.0f)
    {
        for (k = 0;  k < 2;  k++)
        {
            tx->tone_scaling[k][0] = dds_scaling_dbm0(tone_db);
            tx->tone_scaling[k][1] = dds_scaling_dbm0(tone_db);
        }
    }
    if (freq_offset != 0.0f)
    {
        for (k = 0;  k < 2;  k++)
        {
            if (tx->desc->tone_freq[k])
                tx->phase_rate[k] = dds_phase_rate((float) tx->desc->tone_freq[k] + freq_offset);
        }
    }
    tmp = (int16_t *) calloc((n > 0)  ?  n  :  1, sizeof(int16_t));
    pos = 0;
    for (i = 0;  i < nsteps  &&  pos < n;  i++)
    {
        len = steps[2*i + 1];
        if (len > n - pos)
            len = n - pos;
        sig_tone_tx_set_mode(tx, steps[2*i], 0);
        sig_tone_tx(tx, tmp + pos, len);
        pos += len;
    }
    sig_tone_tx_free(tx);
    for (i = 0;  i < pos;  i++)
        amp[i] = sat_add16(amp[i], tmp[i]);
    free(tmp);
    if (noise_dbm0 > -99.0f)
    {
        noise = awgn_init_dbm0(NULL, noise_seed, noise_dbm0);
        for (i = 0;  i < n;  i++)
            amp[i] = sat_add16(amp[i], awgn(noise));
        awgn_free(noise);
    }
    return pos;
}

typedef struct
{
    int32_t *ev;            /* {index of the rx call, signalling_state, duration} */
    int cap;
    int n;
    int call;
} sig_rec_t;

static void rx_update(void *user, int what, int level, int duration)
{
    sig_rec_t *r = (sig_rec_t *) user;
    if (r->n < r->cap)
    {
        r->ev[3*r->n] = r->call;
        r->ev[3*r->n + 1] = what;
        r->ev[3*r->n + 2] = duration;
    }
    r->n++;
}

static int32_t fbits(float f)
{
    int32_t v;
    memcpy(&v, &f, 4);
    return v;
}

/* One channel through sig_tone_rx() IN PLACE, in calls of the lengths in lens[] (or of `chunk` samples when lens is
   NULL).  modes[] = {call index, mode} pairs: sig_tone_rx_set_mode(mode) before that call (call 0 = right after init).
   final[31]: the receiver state in the order of sb_sig_rx.cuh's T_* fields. */
EXPORT int ref_sig_run(int16_t *amp, int n, int chunk, const int32_t *lens, int ncalls, int tone_type, const int32_t *modes, int nmodes,
                       int32_t *ev, int ev_cap, int32_t *nev, int32_t *final)
{
    sig_tone_rx_state_t *rx;
    sig_rec_t rec;
    int pos;
    int len;
    int j;
    int i;
    int m;

    rec.ev = ev;
    rec.cap = ev_cap;
    rec.n = 0;
    rec.call = 0;
    if ((rx = sig_tone_rx_init(NULL, tone_type, rx_update, &rec)) == NULL)
        return -1;
    if (chunk <= 0)
        chunk = n;
    for (pos = 0;  (lens)  ?  (rec.call < ncalls)  :  (pos < n);  pos += len)
    {
        for (m = 0;  m < nmodes;  m++)
        {
            if (modes[2*m] == rec.call)
                sig_tone_rx_set_mode(rx, modes[2*m + 1], 0);
        }
        if (lens)
            len = lens[rec.call];
        else
            len = (n - pos < chunk)  ?  (n - pos)  :  chunk;
        sig_tone_rx(rx, amp + pos, len);
        rec.call++;
    }
    *nev = rec.n;
    if (final)
    {
        final[0] = (int32_t) (rx->desc->tones == 2  ?  3  :  (rx->desc->tone_freq[0] == 2280  ?  1  :  2));
        final[1] = rx->current_rx_tone;
        final[2] = rx->current_notch_filter;
        for (j = 0;  j < 3;  j++)
        {
            for (i = 0;  i < 2;  i++)
            {
                final[3 + 2*j + i] = fbits(rx->tone[j].notch_z1[i]);
                final[9 + 2*j + i] = fbits(rx->tone[j].notch_z2[i]);
            }
            final[15 + j] = rx->tone[j].power.reading;
        }
        final[18] = fbits(rx->flat_z[0]);
        final[19] = fbits(rx->flat_z[1]);
        final[20] = rx->flat_power.reading;
        final[21] = rx->tone_persistence_timeout;
        final[22] = rx->last_sample_tone_present;
        final[23] = rx->flat_detection_threshold;
        final[24] = rx->sharp_detection_threshold;
        final[25] = rx->detection_ratio;
        final[26] = rx->flat_mode;
        final[27] = rx->flat_mode_timeout;
        final[28] = rx->notch_insertion_timeout;
        final[29] = rx->signalling_state;
        final[30] = rx->signalling_state_duration;
    }
    sig_tone_rx_free(rx);
    return 0;
}
