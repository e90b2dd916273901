/*
 * ref_harness_tx.c - flat, ctypes-friendly entry points around the UNMODIFIED reference transmit-side sources that the
 * device generator banks restate: tone_gen with a full descriptor (src/tone_generate.c) and v29_tx (src/v29tx.c).
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
#include "spandsp/complex.h"
#include "spandsp/vector_float.h"
#include "spandsp/complex_vector_float.h"
#include "spandsp/async.h"
#include "spandsp/dds.h"
#include "spandsp/power_meter.h"
#include "spandsp/tone_generate.h"
#include "spandsp/v29tx.h"

#include "spandsp/private/logging.h"
#include "spandsp/private/tone_generate.h"
#include "spandsp/private/v29tx.h"

#include "v29tx_rrc.h"

#define EXPORT __attribute__((visibility("default")))

/* tone_gen_descriptor_init(desc[0..8] = f1, l1, f2, l2, d1, d2, d3, d4, repeat) + tone_gen_init(), then ncalls calls of
   tone_gen(); call k may write max_lens[k] samples after those of the earlier calls and its return value goes to
   out_lens[k].  What tone_gen() does not write keeps the caller's contents. */
EXPORT int ref_tone_gen_calls(int16_t *amp, const int32_t *max_lens, int ncalls, const int32_t *desc, int32_t *out_lens)
{
    tone_gen_descriptor_t d;
    tone_gen_state_t tone;
    int k;
    int pos;

    tone_gen_descriptor_init(&d, desc[0], desc[1], desc[2], desc[3], desc[4], desc[5], desc[6], desc[7], desc[8] != 0);
    tone_gen_init(&tone, &d);
    pos = 0;
    for (k = 0;  k < ncalls;  k++)
    {
        out_lens[k] = tone_gen(&tone, amp + pos, max_lens[k]);
        pos += max_lens[k];
    }
    return 0;
}

typedef struct
{
    int mode;                   /* 0: x^23 + x^18 + 1 sequence, 1: caller bits (LSB first), then SIG_STATUS_END_OF_DATA */
    uint32_t lfsr;
    const uint8_t *bits;
    int nbits;
    int pos;
    int status;                 /* bit 0: SIG_STATUS_END_OF_DATA reported, bit 1: SIG_STATUS_SHUTDOWN_COMPLETE reported */
} tx_src_t;

static int tx_get_bit(void *user)
{
    tx_src_t *p = (tx_src_t *) user;
    int bit;

    if (p->mode == 0)
    {
        bit = ((p->lfsr >> 22) ^ (p->lfsr >> 17)) & 1;
        p->lfsr = ((p->lfsr << 1) | bit) & 0x7FFFFF;
        return bit;
    }
    if (p->pos >= p->nbits)
        return SIG_STATUS_END_OF_DATA;
    bit = (p->bits[p->pos >> 3] >> (p->pos & 7)) & 1;
    p->pos++;
    return bit;
}

static void tx_status(void *user, int status)
{
    tx_src_t *p = (tx_src_t *) user;

    if (status == SIG_STATUS_END_OF_DATA)
        p->status |= 1;
    else if (status == SIG_STATUS_SHUTDOWN_COMPLETE)
        p->status |= 2;
}

/* v29_tx_init(bit_rate, tep) + v29_tx_power(power_dbm0), then ncalls calls of v29_tx() laid out as in
   ref_tone_gen_calls.  If restart_before_call >= 0, v29_tx_restart(restart_rate, restart_tep) is called before that
   call.  *status receives the status bits above. */
EXPORT int ref_v29_tx_calls(int16_t *amp, const int32_t *max_lens, int ncalls, int bit_rate, int tep, float power_dbm0,
                            int src_mode, uint32_t lfsr_seed, const uint8_t *bits, int nbits,
                            int restart_before_call, int restart_rate, int restart_tep,
                            int32_t *out_lens, int32_t *status)
{
    v29_tx_state_t *tx;
    tx_src_t src;
    int k;
    int pos;

    memset(&src, 0, sizeof(src));
    src.mode = src_mode;
    src.lfsr = (lfsr_seed & 0x7FFFFF)  ?  (lfsr_seed & 0x7FFFFF)  :  1;
    src.bits = bits;
    src.nbits = nbits;
    if ((tx = v29_tx_init(NULL, bit_rate, tep != 0, tx_get_bit, &src)) == NULL)
        return -1;
    v29_tx_set_modem_status_handler(tx, tx_status, &src);
    v29_tx_power(tx, power_dbm0);
    pos = 0;
    for (k = 0;  k < ncalls;  k++)
    {
        if (k == restart_before_call)
            v29_tx_restart(tx, restart_rate, restart_tep != 0);
        out_lens[k] = v29_tx(tx, amp + pos, max_lens[k]);
        pos += max_lens[k];
    }
    *status = src.status;
    v29_tx_free(tx);
    return 0;
}

/* The transmit pulse shaper as the reference build sees it (generated header v29tx_rrc.h): [10][9] */
EXPORT void ref_v29_tx_tables(float *shaper)
{
    int i;
    int j;

    for (j = 0;  j < TX_PULSESHAPER_COEFF_SETS;  j++)
    {
        for (i = 0;  i < V29_TX_FILTER_STEPS;  i++)
            shaper[j*V29_TX_FILTER_STEPS + i] = tx_pulseshaper[j][i];
    }
}
