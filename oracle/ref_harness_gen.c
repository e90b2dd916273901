/*
 * ref_harness_gen.c - flat, ctypes-friendly entry points around the UNMODIFIED reference signal sources:
 * dtmf_tx (src/dtmf.c), tone_gen (src/tone_generate.c), awgn (src/awgn.c), the float DDS table (src/dds_float.c).
 * TEST INFRASTRUCTURE ONLY.
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
#include "spandsp/queue.h"
#include "spandsp/complex.h"
#include "spandsp/dds.h"
#include "spandsp/awgn.h"
#include "spandsp/tone_detect.h"
#include "spandsp/tone_generate.h"
#include "spandsp/super_tone_rx.h"
#include "spandsp/dtmf.h"

#include "spandsp/private/tone_generate.h"
#include "spandsp/private/awgn.h"

#define EXPORT __attribute__((visibility("default")))

/* One transmitter: optional dtmf_tx_set_level / dtmf_tx_set_timing, dtmf_tx_put(digits), then ncalls calls of
   dtmf_tx(); call k may write max_lens[k] samples at amp + (sum of the earlier max_lens) and its return value goes to
   out_lens[k].  digits2 (may be NULL) is put before call number put2_before_call.  put_results[0..1] receive what the
   two dtmf_tx_put() calls returned.  What dtmf_tx() does not write keeps the caller's contents. */
EXPORT int ref_dtmf_tx_calls(int16_t *amp, const int32_t *max_lens, int ncalls, const char *digits, const char *digits2, int put2_before_call,
                             int set_level, int level, int twist, int set_timing, int on_ms, int off_ms,
                             int32_t *out_lens, int32_t *put_results)
{
    dtmf_tx_state_t *tx;
    int k;
    int pos;

    if ((tx = dtmf_tx_init(NULL, NULL, NULL)) == NULL)
        return -1;
    if (set_level)
        dtmf_tx_set_level(tx, level, twist);
    if (set_timing)
        dtmf_tx_set_timing(tx, on_ms, off_ms);
    put_results[0] = dtmf_tx_put(tx, digits, -1);
    put_results[1] = 0;
    pos = 0;
    for (k = 0;  k < ncalls;  k++)
    {
        if (digits2  &&  k == put2_before_call)
            put_results[1] = dtmf_tx_put(tx, digits2, -1);
        out_lens[k] = dtmf_tx(tx, amp + pos, max_lens[k]);
        pos += max_lens[k];
    }
    dtmf_tx_free(tx);
    return 0;
}

/* n x awgn(): amp[i] = sat_add16(amp[i], awgn()) (add != 0) or amp[i] = awgn(); level in dBm0 or (dbov != 0) dBov;
   calls[] (ncalls entries summing to n; may be NULL) only documents that the state carries across calls - awgn() has
   no per-call behaviour. */
EXPORT int ref_awgn_run(int16_t *amp, int n, int seed, float level, int dbov, int add)
{
    awgn_state_t *s;
    int i;

    s = (dbov)  ?  awgn_init_dbov(NULL, seed, level)  :  awgn_init_dbm0(NULL, seed, level);
    if (s == NULL)
        return -1;
    for (i = 0;  i < n;  i++)
        amp[i] = (add)  ?  sat_add16(amp[i], awgn(s))  :  awgn(s);
    awgn_free(s);
    return 0;
}

/* The float DDS table as dds_lookupf() sees it (2048 entries), and the constants the DTMF transmitter derives:
   consts[0..7] = dds_phase_ratef() of the row and column frequencies (as int32 bits in floats' place: use the int view),
   gains[0] = dds_scaling_dbm0f(-10), gains[1] = dds_scaling_dbm0f(-13), gains[2] = dds_scaling_dbm0f(0). */
EXPORT void ref_gen_tables(float *sine, int32_t *rates, float *gains)
{
    static const int row[4] = {697, 770, 852, 941};
    static const int col[4] = {1209, 1336, 1477, 1633};
    int i;

    for (i = 0;  i < 2048;  i++)
        sine[i] = dds_lookupf((uint32_t) i << 21);
    for (i = 0;  i < 4;  i++)
    {
        rates[i] = dds_phase_ratef((float) row[i]);
        rates[4 + i] = dds_phase_ratef((float) col[i]);
    }
    gains[0] = dds_scaling_dbm0f(-10.0f);
    gains[1] = dds_scaling_dbm0f(-13.0f);
    gains[2] = dds_scaling_dbm0f(0.0f);
}

/* One burst as tests/dtmf_rx_tests.c:my_dtmf_gen_init() + my_dtmf_generate() make it: tone_gen_descriptor_init(f1, l1,
   f2, l2, on_ms, off_ms, 0, 0, false) - frequencies as int, like the test passes them - then one tone_gen() of at most
   max_samples samples (1000 in dtmf_rx_tests.c, 9999 in bell_mf_rx_tests.c).  Returns the samples written. */
EXPORT int ref_tone_burst(int16_t *amp, int max_samples, int f1, int l1, int f2, int l2, int on_ms, int off_ms)
{
    tone_gen_descriptor_t desc;
    tone_gen_state_t tone;

    tone_gen_descriptor_init(&desc, f1, l1, f2, l2, on_ms, off_ms, 0, 0, false);
    tone_gen_init(&tone, &desc);
    return tone_gen(&tone, amp, max_samples);
}

/* The stimulus of tests/super_tone_rx_tests.c:detection_range_tests(): 350 Hz + 440 Hz from the integer DDS with
   running phases, at every level from first_level to last_level dBm0 in 1 dB steps, `chunks` chunks of 160 samples per
   level.  Returns the samples written. */
EXPORT int ref_super_tone_range_stimulus(int16_t *amp, int max_samples, int first_level, int last_level, int chunks)
{
    uint32_t phase[2];
    int32_t phase_inc[2];
    int scale;
    int level;
    int i;
    int j;
    int n;

    phase[0] = 0;
    phase_inc[0] = dds_phase_rate(350.0f);
    phase[1] = 0;
    phase_inc[1] = dds_phase_rate(440.0f);
    n = 0;
    for (level = first_level;  level <= last_level;  level++)
    {
        scale = dds_scaling_dbm0(level);
        for (j = 0;  j < chunks;  j++)
        {
            for (i = 0;  i < 160;  i++)
            {
                if (n >= max_samples)
                    return n;
                amp[n] = (dds(&phase[0], phase_inc[0])*scale) >> 15;
                amp[n] += (dds(&phase[1], phase_inc[1])*scale) >> 15;
                n++;
            }
        }
    }
    return n;
}
