/*
 * ref_harness.c - flat, ctypes-friendly entry points around the UNMODIFIED reference
 * implementation of the hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is compiled INTO oracle/_ref/libspandsp_ref_{strict,fast}.so together with
 * the reference's own sources (taken in place from /root/reference/src; nothing is copied
 * into this repository).  It contains no DSP of its own: every sample goes through the
 * reference's dtmf_rx()/bell_mf_rx()/r2_mf_rx()/super_tone_rx()/v29_rx() and every callback
 * is recorded verbatim as a ref_event_t.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load the resulting library.
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
#include "spandsp/queue.h"
#include "spandsp/complex.h"
#include "spandsp/dds.h"
#include "spandsp/awgn.h"
#include "spandsp/tone_detect.h"
#include "spandsp/tone_generate.h"
#include "spandsp/super_tone_rx.h"
#include "spandsp/super_tone_tx.h"
#include "spandsp/dtmf.h"
#include "spandsp/bell_r2_mf.h"
#include "spandsp/async.h"
#include "spandsp/power_meter.h"
#include "spandsp/godard.h"
#include "spandsp/v29rx.h"
#include "spandsp/v29tx.h"
#include "spandsp/v17rx.h"
#include "spandsp/v17tx.h"

#include "spandsp/private/logging.h"
#include "spandsp/private/tone_detect.h"
#include "spandsp/private/power_meter.h"
#include "spandsp/private/godard.h"
#include "spandsp/private/v29rx.h"
#include "spandsp/private/v17rx.h"
#include "spandsp/private/super_tone_rx.h"
#include "spandsp/private/tone_generate.h"
#include "spandsp/private/super_tone_tx.h"

#include "spandsp/math_fixed.h"
#include "spandsp/bit_operations.h"
#include "spandsp/g711.h"
#include "ref_harness.h"

/* The generated tables the V.29 receiver is built on (static const in the generated headers). */
#include "v29rx_rrc.h"
#include "v29rx_godard.h"
#include "math_fixed_tables.h"

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------ */
/* event recording                                                                      */

typedef struct
{
    ref_event_t *ev;
    int64_t cap;
    int64_t n;
    int32_t chunk;      /* index of the rx call during which the callback fired */
} recorder_t;

static void rec_push(recorder_t *r, int kind, int a, int b, int c)
{
    if (r->n < r->cap)
    {
        r->ev[r->n].chunk = r->chunk;
        r->ev[r->n].kind = kind;
        r->ev[r->n].a = a;
        r->ev[r->n].b = b;
        r->ev[r->n].c = c;
    }
    r->n++;
}

static void digits_cb(void *user, const char *digits, int len)
{
    int i;

    for (i = 0;  i < len;  i++)
        rec_push((recorder_t *) user, REF_EV_DIGIT, (unsigned char) digits[i], len, i);
}

static void realtime_cb(void *user, int code, int level, int delay)
{
    rec_push((recorder_t *) user, REF_EV_TONE, code, level, delay);
}

static void segment_cb(void *user, int f1, int f2, int duration)
{
    rec_push((recorder_t *) user, REF_EV_SEGMENT, f1, f2, duration);
}

/* ------------------------------------------------------------------------------------ */
/* signal generation (reference dtmf_tx / bell_mf_tx / r2_mf_tx / awgn)                   */

EXPORT int ref_abi_version(void)
{
    return REF_HARNESS_ABI;
}

EXPORT int ref_sizeof(int what)
{
    switch (what)
    {
    case 0: return (int) sizeof(goertzel_state_t);
    case 4: return (int) sizeof(v29_rx_state_t);
    case 5: return (int) sizeof(v17_rx_state_t);
    }
    return -1;
}

static void add_awgn(int16_t *amp, int n, int seed, float level_dbm0)
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

/* dtmf_tx -> (+ awgn).  on/off in ms (<0: defaults), level in dBm0, twist in dB. */
EXPORT int ref_dtmf_generate(int16_t *amp, int max_samples, const char *digits,
                             int level, int twist, int on_ms, int off_ms,
                             int noise_seed, float noise_dbm0)
{
    dtmf_tx_state_t *tx;
    int len;
    int total;
    int n;
    const char *p;

    tx = dtmf_tx_init(NULL, NULL, NULL);
    dtmf_tx_set_level(tx, level, twist);
    dtmf_tx_set_timing(tx, on_ms, off_ms);
    memset(amp, 0, sizeof(int16_t)*max_samples);
    total = 0;
    p = digits;
    n = (int) strlen(digits);
    /* The tx queue holds 128 digits; feed it in pieces. */
    while (total < max_samples)
    {
        int piece = (n > 100)  ?  100  :  n;
        if (piece > 0)
        {
            dtmf_tx_put(tx, p, piece);
            p += piece;
            n -= piece;
        }
        len = dtmf_tx(tx, amp + total, max_samples - total);
        total += len;
        if (len == 0  &&  n == 0)
            break;
    }
    dtmf_tx_free(tx);
    add_awgn(amp, max_samples, noise_seed, noise_dbm0);
    return total;
}

EXPORT int ref_bell_mf_generate(int16_t *amp, int max_samples, const char *digits,
                                int noise_seed, float noise_dbm0)
{
    bell_mf_tx_state_t *tx;
    int total;

    tx = bell_mf_tx_init(NULL);
    memset(amp, 0, sizeof(int16_t)*max_samples);
    bell_mf_tx_put(tx, digits, -1);
    total = bell_mf_tx(tx, amp, max_samples);
    bell_mf_tx_free(tx);
    add_awgn(amp, max_samples, noise_seed, noise_dbm0);
    return total;
}

/* Each digit is keyed for on_samples and then silence for off_samples. */
EXPORT int ref_r2_mf_generate(int16_t *amp, int max_samples, const char *digits, int fwd,
                              int on_samples, int off_samples,
                              int noise_seed, float noise_dbm0)
{
    r2_mf_tx_state_t *tx;
    int total;
    int len;
    const char *p;

    tx = r2_mf_tx_init(NULL, fwd != 0);
    memset(amp, 0, sizeof(int16_t)*max_samples);
    total = 0;
    for (p = digits;  *p  &&  total < max_samples;  p++)
    {
        len = on_samples;
        if (len > max_samples - total)
            len = max_samples - total;
        r2_mf_tx_put(tx, *p);
        total += r2_mf_tx(tx, amp + total, len);
        len = off_samples;
        if (len > max_samples - total)
            len = max_samples - total;
        r2_mf_tx_put(tx, 0);
        r2_mf_tx(tx, amp + total, len);     /* generates nothing; buffer is already silent */
        total += len;
    }
    r2_mf_tx_free(tx);
    add_awgn(amp, max_samples, noise_seed, noise_dbm0);
    return total;
}

/* A cadenced dual tone: nsteps x {f1,f2 (Hz, 0 = none), level dBm0, length ms}, repeated. */
EXPORT int ref_cadence_generate(int16_t *amp, int max_samples, const int *steps, int nsteps,
                                int noise_seed, float noise_dbm0)
{
    super_tone_tx_step_t *tree;
    super_tone_tx_step_t *last;
    super_tone_tx_step_t *st;
    super_tone_tx_state_t *tx;
    int i;
    int total;

    tree = NULL;
    last = NULL;
    for (i = 0;  i < nsteps;  i++)
    {
        st = super_tone_tx_make_step(NULL,
                                     (float) steps[4*i + 0], (float) steps[4*i + 2],
                                     (float) steps[4*i + 1], (float) steps[4*i + 2],
                                     steps[4*i + 3], 1);
        if (last)
            last->next = st;
        else
            tree = st;
        last = st;
    }
    memset(amp, 0, sizeof(int16_t)*max_samples);
    total = 0;
    /* The tone tree is played once per init; loop it to fill the buffer. */
    while (total < max_samples)
    {
        int len;
        tx = super_tone_tx_init(NULL, tree);
        len = super_tone_tx(tx, amp + total, max_samples - total);
        super_tone_tx_free(tx);
        if (len <= 0)
            break;
        total += len;
    }
    super_tone_tx_free_tone(tree);
    add_awgn(amp, max_samples, noise_seed, noise_dbm0);
    return total;
}

EXPORT void ref_awgn_add(int16_t *amp, int n, int seed, float level_dbm0)
{
    add_awgn(amp, n, seed, level_dbm0);
}

/* ------------------------------------------------------------------------------------ */
/* one-channel runners                                                                   */

static int run_dtmf(const ref_params_t *p, const int16_t *amp, int n, recorder_t *rec, ref_final_t *fin)
{
    dtmf_rx_state_t *s;
    int pos;
    int len;
    char buf[256];

    s = dtmf_rx_init(NULL, (p->mode == REF_MODE_DIGITS_CB)  ?  digits_cb  :  NULL, rec);
    if (p->mode == REF_MODE_REALTIME)
        dtmf_rx_set_realtime_callback(s, realtime_cb, rec);
    if (p->dtmf_set_parms)
        dtmf_rx_parms(s, p->dtmf_filter_dialtone, p->dtmf_twist, p->dtmf_reverse_twist, p->dtmf_threshold);
    rec->chunk = 0;
    for (pos = 0;  pos < n;  pos += len)
    {
        len = (n - pos < p->chunk)  ?  (n - pos)  :  p->chunk;
        if (p->fillin_every > 0  &&  rec->chunk > 0  &&  (rec->chunk % p->fillin_every) == 0)
            dtmf_rx_fillin(s, len);
        else
            dtmf_rx(s, amp + pos, len);
        rec->chunk++;
    }
    if (fin)
    {
        fin->status = dtmf_rx_status(s);
        if (p->mode == REF_MODE_POLL)
        {
            /* Drain what accumulated in the digit buffer */
            fin->ndigits = (int) dtmf_rx_get(s, buf, 255);
            memcpy(fin->digits, buf, fin->ndigits + 1);
        }
    }
    dtmf_rx_free(s);
    return 0;
}

static int run_bell_mf(const ref_params_t *p, const int16_t *amp, int n, recorder_t *rec, ref_final_t *fin)
{
    bell_mf_rx_state_t *s;
    int pos;
    int len;
    char buf[256];

    s = bell_mf_rx_init(NULL, (p->mode == REF_MODE_DIGITS_CB)  ?  digits_cb  :  NULL, rec);
    rec->chunk = 0;
    for (pos = 0;  pos < n;  pos += len)
    {
        len = (n - pos < p->chunk)  ?  (n - pos)  :  p->chunk;
        bell_mf_rx(s, amp + pos, len);
        rec->chunk++;
    }
    if (fin  &&  p->mode == REF_MODE_POLL)
    {
        fin->ndigits = (int) bell_mf_rx_get(s, buf, 255);
        memcpy(fin->digits, buf, fin->ndigits + 1);
    }
    bell_mf_rx_free(s);
    return 0;
}

static int run_r2_mf(const ref_params_t *p, const int16_t *amp, int n, recorder_t *rec, ref_final_t *fin)
{
    r2_mf_rx_state_t *s;
    int pos;
    int len;

    s = r2_mf_rx_init(NULL, p->r2_fwd != 0, realtime_cb, rec);
    rec->chunk = 0;
    for (pos = 0;  pos < n;  pos += len)
    {
        len = (n - pos < p->chunk)  ?  (n - pos)  :  p->chunk;
        r2_mf_rx(s, amp + pos, len);
        rec->chunk++;
    }
    if (fin)
        fin->status = r2_mf_rx_get(s);
    r2_mf_rx_free(s);
    return 0;
}

/* Super-tone descriptor tables are passed flat: tone t has tone_segs[t] elements, each
   {f1, f2, min_ms, max_ms}, concatenated in `elements`. */
static super_tone_rx_descriptor_t *build_descriptor(const ref_params_t *p)
{
    super_tone_rx_descriptor_t *desc;
    int t;
    int e;
    int k;
    int tone;

    desc = super_tone_rx_make_descriptor(NULL);
    k = 0;
    for (t = 0;  t < p->st_tones;  t++)
    {
        tone = super_tone_rx_add_tone(desc);
        for (e = 0;  e < p->st_tone_segs[t];  e++, k++)
        {
            super_tone_rx_add_element(desc, tone,
                                      p->st_elements[4*k + 0], p->st_elements[4*k + 1],
                                      p->st_elements[4*k + 2], p->st_elements[4*k + 3]);
        }
    }
    return desc;
}

static int run_super_tone(const ref_params_t *p, const int16_t *amp, int n, recorder_t *rec, ref_final_t *fin)
{
    super_tone_rx_descriptor_t *desc;
    super_tone_rx_state_t *s;
    int pos;
    int len;

    desc = build_descriptor(p);
    s = super_tone_rx_init(NULL, desc, realtime_cb, rec);
    if (p->mode == REF_MODE_SEGMENTS)
        super_tone_rx_segment_callback(s, segment_cb);
    /* The reference leaves `rotation` uninitialised (super_tone_rx.c:507-547); it is always
       written before it is read, so zeroing it here changes nothing observable. */
    s->rotation = 0;
    rec->chunk = 0;
    for (pos = 0;  pos < n;  pos += len)
    {
        len = (n - pos < p->chunk)  ?  (n - pos)  :  p->chunk;
        super_tone_rx(s, amp + pos, len);
        rec->chunk++;
    }
    if (fin)
    {
        fin->status = s->detected_tone;
        fin->ndigits = desc->monitored_frequencies;
    }
    super_tone_rx_free(s);
    super_tone_rx_free_descriptor(desc);
    return 0;
}

/* Report the Goertzel coefficient table the reference builds for a super-tone descriptor. */
EXPORT int ref_super_tone_bins(const ref_params_t *p, float *fac, int max)
{
    super_tone_rx_descriptor_t *desc;
    int i;
    int n;

    desc = build_descriptor(p);
    n = desc->monitored_frequencies;
    for (i = 0;  i < n  &&  i < max;  i++)
        fac[i] = desc->desc[i].fac;
    super_tone_rx_free_descriptor(desc);
    return n;
}

static int run_one(const ref_params_t *p, const int16_t *amp, int n, recorder_t *rec, ref_final_t *fin)
{
    if (fin)
        memset(fin, 0, sizeof(*fin));
    switch (p->detector)
    {
    case REF_DET_DTMF:
        return run_dtmf(p, amp, n, rec, fin);
    case REF_DET_BELL_MF:
        return run_bell_mf(p, amp, n, rec, fin);
    case REF_DET_R2_MF:
        return run_r2_mf(p, amp, n, rec, fin);
    case REF_DET_SUPER_TONE:
        return run_super_tone(p, amp, n, rec, fin);
    }
    return -1;
}

/* ------------------------------------------------------------------------------------ */
/* multi-channel, multi-thread runner (also the CPU baseline timer)                      */

typedef struct
{
    const ref_params_t *p;
    const int16_t *amp;
    int64_t stride;
    int c0;
    int c1;
    int n;
    ref_event_t *ev;
    int64_t ev_cap;
    int64_t *ev_count;
    ref_final_t *fin;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *) arg;
    recorder_t rec;
    int c;

    for (c = j->c0;  c < j->c1;  c++)
    {
        rec.ev = (j->ev)  ?  (j->ev + (int64_t) c*j->ev_cap)  :  NULL;
        rec.cap = (j->ev)  ?  j->ev_cap  :  0;
        rec.n = 0;
        rec.chunk = 0;
        run_one(j->p, j->amp + (int64_t) c*j->stride, j->n, &rec, (j->fin)  ?  &j->fin[c]  :  NULL);
        if (j->ev_count)
            j->ev_count[c] = rec.n;
    }
    return NULL;
}

/* Run `channels` independent detectors, channel c reading amp[c*stride .. c*stride+n).
   Events of channel c go to ev[c*ev_cap ..]; ev_count[c] receives the number produced
   (which may exceed ev_cap: the excess is dropped).  Returns elapsed seconds. */
EXPORT double ref_run(const ref_params_t *p, const int16_t *amp, int64_t stride, int channels, int n,
                      int nthreads, ref_event_t *ev, int64_t ev_cap, int64_t *ev_count, ref_final_t *fin)
{
    pthread_t *th;
    job_t *jobs;
    struct timespec t0;
    struct timespec t1;
    int i;

    if (nthreads < 1)
        nthreads = 1;
    if (nthreads > channels)
        nthreads = channels;
    th = (pthread_t *) malloc(sizeof(pthread_t)*nthreads);
    jobs = (job_t *) malloc(sizeof(job_t)*nthreads);
    /* Make sure the reference's lazy global descriptor tables are built before threads race
       for them (dtmf.c:482-491, bell_r2_mf.c:698,895). */
    {
        recorder_t rec = {NULL, 0, 0, 0};
        int16_t z[1] = {0};
        run_one(p, z, 0, &rec, NULL);
    }
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (i = 0;  i < nthreads;  i++)
    {
        jobs[i].p = p;
        jobs[i].amp = amp;
        jobs[i].stride = stride;
        jobs[i].c0 = (int) ((int64_t) channels*i/nthreads);
        jobs[i].c1 = (int) ((int64_t) channels*(i + 1)/nthreads);
        jobs[i].n = n;
        jobs[i].ev = ev;
        jobs[i].ev_cap = ev_cap;
        jobs[i].ev_count = ev_count;
        jobs[i].fin = fin;
        if (nthreads == 1)
            worker(&jobs[i]);
        else
            pthread_create(&th[i], NULL, worker, &jobs[i]);
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

/* ------------------------------------------------------------------------------------ */
/* raw Goertzel (goertzel_update / goertzel_result on a caller-visible state)             */

EXPORT float ref_goertzel_fac(float freq, int samples)
{
    goertzel_descriptor_t d;

    make_goertzel_descriptor(&d, freq, samples);
    return d.fac;
}

/* Runs one Goertzel filter over amp[0..n) in blocks of `samples`; out[b] = goertzel_result. */
EXPORT int ref_goertzel_blocks(float freq, int samples, const int16_t *amp, int n, float *out)
{
    goertzel_descriptor_t d;
    goertzel_state_t s;
    int pos;
    int nb;

    make_goertzel_descriptor(&d, freq, samples);
    goertzel_init(&s, &d);
    nb = 0;
    for (pos = 0;  pos < n;  )
    {
        pos += goertzel_update(&s, amp + pos, n - pos);
        if (s.current_sample >= samples)
            out[nb++] = goertzel_result(&s);
    }
    return nb;
}

/* ------------------------------------------------------------------------------------ */
/* V.29 (v29_tx -> awgn -> v29_rx)                                                        */

typedef struct
{
    uint32_t lfsr;
} prbs_t;

static int prbs_get_bit(void *user)
{
    prbs_t *p = (prbs_t *) user;
    /* x^23 + x^18 + 1 */
    const int bit = ((p->lfsr >> 22) ^ (p->lfsr >> 17)) & 1;
    p->lfsr = ((p->lfsr << 1) | bit) & 0x7FFFFF;
    return bit;
}

/* `lead` samples of silence, then a V.29 transmission of PRBS data for the rest of the buffer. */
EXPORT int ref_v29_generate(int16_t *amp, int n, int bit_rate, int tep, float power_dbm0, uint32_t lfsr_seed,
                            int lead, int noise_seed, float noise_dbm0)
{
    v29_tx_state_t *tx;
    prbs_t prbs;
    int len;

    prbs.lfsr = (lfsr_seed & 0x7FFFFF)  ?  (lfsr_seed & 0x7FFFFF)  :  1;
    memset(amp, 0, sizeof(int16_t)*n);
    if (lead > n)
        lead = n;
    tx = v29_tx_init(NULL, bit_rate, tep != 0, prbs_get_bit, &prbs);
    v29_tx_power(tx, power_dbm0);
    len = v29_tx(tx, amp + lead, n - lead);
    v29_tx_free(tx);
    add_awgn(amp, n, noise_seed, noise_dbm0);
    return len;
}

typedef struct
{
    int8_t *bits;
    int cap;
    int n;
    ref_v29_sym_t *syms;
    int sym_cap;
    int nsyms;
} v29_rec_t;

static void v29_put_bit(void *user, int bit)
{
    v29_rec_t *r = (v29_rec_t *) user;
    if (r->n < r->cap)
        r->bits[r->n] = (int8_t) bit;
    r->n++;
}

static void v29_qam(void *user, const complexf_t *z, const complexf_t *target, int symbol)
{
    v29_rec_t *r = (v29_rec_t *) user;
    if (r->nsyms < r->sym_cap)
    {
        r->syms[r->nsyms].re = z->re;
        r->syms[r->nsyms].im = z->im;
        r->syms[r->nsyms].tre = target->re;
        r->syms[r->nsyms].tim = target->im;
        r->syms[r->nsyms].state = symbol;
    }
    r->nsyms++;
}

/* One channel.  bits[] receives what put_bit delivered, in order: 0/1 data bits and the negative
   SIG_STATUS_* codes (no separate status handler is installed, src/v29rx.c:171-178).
   final[]: {training_stage, carrier_phase_rate, eq_put_step, signal_present, agc_scaling bits,
             total_baud_timing_correction, constellation_state, carrier_phase} */
EXPORT int ref_v29_run_ex(const int16_t *amp, int n, int chunk, int bit_rate, float cutoff, int want_qam,
                          int restart_at, int restart_old_train,
                          int8_t *bits, int bits_cap, int32_t *nbits,
                          ref_v29_sym_t *syms, int sym_cap, int32_t *nsyms,
                          float *eq_coeff, int32_t *final);

EXPORT int ref_v29_run(const int16_t *amp, int n, int chunk, int bit_rate, float cutoff, int want_qam,
                       int8_t *bits, int bits_cap, int32_t *nbits,
                       ref_v29_sym_t *syms, int sym_cap, int32_t *nsyms,
                       float *eq_coeff, int32_t *final)
{
    return ref_v29_run_ex(amp, n, chunk, bit_rate, cutoff, want_qam, -1, 0, bits, bits_cap, nbits, syms, sym_cap, nsyms, eq_coeff, final);
}

/* As ref_v29_run; if restart_at >= 0, v29_rx_restart(rx, bit_rate, restart_old_train) is called before the
   rx call that starts at the first chunk boundary >= restart_at. */
EXPORT int ref_v29_run_ex(const int16_t *amp, int n, int chunk, int bit_rate, float cutoff, int want_qam,
                          int restart_at, int restart_old_train,
                          int8_t *bits, int bits_cap, int32_t *nbits,
                          ref_v29_sym_t *syms, int sym_cap, int32_t *nsyms,
                          float *eq_coeff, int32_t *final)
{
    v29_rx_state_t *rx;
    v29_rec_t rec;
    int pos;
    int len;
    int i;

    rec.bits = bits;
    rec.cap = bits_cap;
    rec.n = 0;
    rec.syms = syms;
    rec.sym_cap = sym_cap;
    rec.nsyms = 0;
    rx = v29_rx_init(NULL, bit_rate, v29_put_bit, &rec);
    if (rx == NULL)
        return -1;
    if (cutoff > -99.0f)
        v29_rx_set_signal_cutoff(rx, cutoff);
    if (want_qam)
        v29_rx_set_qam_report_handler(rx, v29_qam, &rec);
    for (pos = 0;  pos < n;  pos += len)
    {
        if (restart_at >= 0  &&  pos >= restart_at)
        {
            v29_rx_restart(rx, bit_rate, restart_old_train != 0);
            restart_at = -1;
        }
        len = (n - pos < chunk)  ?  (n - pos)  :  chunk;
        v29_rx(rx, amp + pos, len);
    }
    *nbits = rec.n;
    *nsyms = rec.nsyms;
    if (eq_coeff)
    {
        for (i = 0;  i < V29_EQUALIZER_LEN;  i++)
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
        final[5] = rx->godard.total_baud_timing_correction;
        final[6] = rx->constellation_state;
        final[7] = (int32_t) rx->carrier_phase;
    }
    v29_rx_free(rx);
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
    int8_t *bits;
    int64_t bits_cap;
    int32_t *nbits;
} v29_job_t;

static void *v29_worker(void *arg)
{
    v29_job_t *j = (v29_job_t *) arg;
    int32_t nsyms;
    int32_t nb;
    int c;
    int8_t *scratch = NULL;

    if (j->bits == NULL)
        scratch = (int8_t *) malloc(16);
    for (c = j->c0;  c < j->c1;  c++)
    {
        ref_v29_run(j->amp + (int64_t) c*j->stride, j->n, j->chunk, j->bit_rate, j->cutoff, 0,
                    (j->bits)  ?  (j->bits + (int64_t) c*j->bits_cap)  :  scratch, (j->bits)  ?  (int) j->bits_cap  :  0, &nb,
                    NULL, 0, &nsyms, NULL, NULL);
        if (j->nbits)
            j->nbits[c] = nb;
    }
    free(scratch);
    return NULL;
}

/* Many channels on nthreads host threads; returns elapsed seconds (CPU baseline for cfg4). */
EXPORT double ref_v29_run_batch(const int16_t *amp, int64_t stride, int channels, int n, int chunk, int bit_rate, float cutoff,
                                int nthreads, int8_t *bits, int64_t bits_cap, int32_t *nbits)
{
    pthread_t *th;
    v29_job_t *jobs;
    struct timespec t0;
    struct timespec t1;
    int i;

    if (nthreads < 1)
        nthreads = 1;
    if (nthreads > channels)
        nthreads = channels;
    th = (pthread_t *) malloc(sizeof(pthread_t)*nthreads);
    jobs = (v29_job_t *) malloc(sizeof(v29_job_t)*nthreads);
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
        jobs[i].bits = bits;
        jobs[i].bits_cap = bits_cap;
        jobs[i].nbits = nbits;
        if (nthreads == 1)
            v29_worker(&jobs[i]);
        else
            pthread_create(&th[i], NULL, v29_worker, &jobs[i]);
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

/* The constant tables the receiver uses, as the reference build sees them. */
EXPORT void ref_v29_tables(float *rrc_re, float *rrc_im, float *sine, uint16_t *sqrt_tab, float *godard, int32_t *ints)
{
    int i;
    int j;

    for (i = 0;  i < RX_PULSESHAPER_COEFF_SETS;  i++)
    {
        for (j = 0;  j < 27;  j++)
        {
            rrc_re[i*27 + j] = rx_pulseshaper_re[i][j];
            rrc_im[i*27 + j] = rx_pulseshaper_im[i][j];
        }
    }
    for (i = 0;  i < 2048;  i++)
        sine[i] = dds_lookupf((uint32_t) i << 21);
    for (i = 0;  i < 193;  i++)
        sqrt_tab[i] = fixed_sqrt_table[i];
    godard[0] = godard_desc.low_band_edge_coeff[0];
    godard[1] = godard_desc.low_band_edge_coeff[1];
    godard[2] = godard_desc.low_band_edge_coeff[2];
    godard[3] = godard_desc.high_band_edge_coeff[0];
    godard[4] = godard_desc.high_band_edge_coeff[1];
    godard[5] = godard_desc.high_band_edge_coeff[2];
    godard[6] = godard_desc.mixed_band_edges_coeff_3;
    godard[7] = godard_desc.coarse_trigger;
    godard[8] = godard_desc.fine_trigger;
    ints[0] = godard_desc.coarse_step;
    ints[1] = godard_desc.fine_step;
    ints[2] = DDS_PHASE_RATE(1700.0f);
    ints[3] = DDS_PHASE_RATE(1700.0f - 20.0f);
    ints[4] = DDS_PHASE_RATE(1700.0f + 20.0f);
    ints[5] = DDS_PHASE(45.0f);
    ints[6] = DDS_PHASE(-45.0f);
    ints[7] = power_meter_level_dbm0(-28.5f + 2.5f);
    ints[8] = power_meter_level_dbm0(-28.5f - 2.5f);
}

/* ------------------------------------------------------------------------------------ */
/* G.711 (src/spandsp/g711.h inline functions)                                            */

/* law: 0 = u-law, 1 = A-law */
EXPORT void ref_g711_encode(int law, const int16_t *in, uint8_t *out, int n)
{
    int i;

    for (i = 0;  i < n;  i++)
        out[i] = (law)  ?  linear_to_alaw(in[i])  :  linear_to_ulaw(in[i]);
}

EXPORT void ref_g711_expand(int law, const uint8_t *in, int16_t *out, int n)
{
    int i;

    for (i = 0;  i < n;  i++)
        out[i] = (law)  ?  alaw_to_linear(in[i])  :  ulaw_to_linear(in[i]);
}
