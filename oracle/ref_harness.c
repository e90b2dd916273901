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

#include "ref_harness.h"

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
