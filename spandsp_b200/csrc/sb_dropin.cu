// sb_dropin.cu - the spandsp-named per-channel API (include/spandsp_b200_dropin.h) on top of the
// bank engine.  Host logic only mirrors what the reference does on the host side of a callback
// (digit buffering, callback selection); every sample is processed by the CUDA kernels.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <vector>

#include "sb_engine.h"
#pragma GCC visibility push(default)
#include "../../include/spandsp_b200_dropin.h"
#include "../../include/spandsp_b200_v29.h"
#include "../../include/spandsp_b200_v17.h"
#include "../../include/spandsp_b200_v27ter.h"
#include "../../include/spandsp_b200_fsk.h"
#include "../../include/spandsp_b200_mct.h"
#include "../../include/spandsp_b200_sig.h"
#pragma GCC visibility pop

#define SB_MAGIC    0x5350414E42323030ULL       /* "SPANB200" */

// ------------------------------------------------------------------------------------------
// state objects (all start with the same header)
struct sb_member_t
{
    unsigned long long magic;
    span_b200_group_s *grp;
    int index;
    int det;
    int staged;                 // samples staged since the last flush
    int heap;                   // allocated by *_init(NULL, ...)
};

struct dtmf_rx_state_s
{
    sb_member_t m;
    digits_rx_callback_t digits_callback;
    void *digits_callback_data;
    span_tone_report_func_t realtime_callback;
    void *realtime_callback_data;
    int lost_digits;
    int current_digits;
    char digits[MAX_DTMF_DIGITS + 1];
};

struct bell_mf_rx_state_s
{
    sb_member_t m;
    digits_rx_callback_t digits_callback;
    void *digits_callback_data;
    int lost_digits;
    int current_digits;
    char digits[MAX_BELL_MF_DIGITS + 1];
};

struct r2_mf_rx_state_s
{
    sb_member_t m;
    span_tone_report_func_t callback;
    void *callback_data;
    int fwd;
};

struct super_tone_rx_state_s
{
    sb_member_t m;
    super_tone_rx_descriptor_t *desc;
    span_tone_report_func_t tone_callback;
    tone_segment_func_t segment_callback;
    void *callback_data;
};

static_assert(sizeof(dtmf_rx_state_s) <= 432, "must fit the reference's dtmf_rx_state_t (private/dtmf.h:54-116)");
static_assert(sizeof(bell_mf_rx_state_s) <= 288, "must fit the reference's bell_mf_rx_state_t (private/bell_r2_mf.h:48-67)");
static_assert(sizeof(r2_mf_rx_state_s) <= 152, "must fit the reference's r2_mf_rx_state_t (private/bell_r2_mf.h:85-100)");
static_assert(sizeof(super_tone_rx_state_s) <= 272, "must fit the reference's super_tone_rx_state_t (private/super_tone_rx.h:51-62)");

struct super_tone_rx_descriptor_s
{
    std::vector<std::vector<int> > tones;       // per tone: {f1, f2, min_ms, max_ms} x elements
    int heap;
};

struct span_b200_group_s
{
    span_b200_ctx_t *ctx;
    int det;
    int members;
    int max_samples;
    int single;                 // group of one made by *_init(): flushes on every rx call
    int arg;
    span_b200_bank_t *bank;
    int16_t *h_stage;           // pinned, [members][max_samples]
    std::vector<void *> owned;  // member objects allocated by group_create
    std::vector<sb_member_t *> member;
    std::vector<span_b200_event_t> events;
};

static std::recursive_mutex g_lock;
static span_b200_ctx_t *g_default_ctx = NULL;

extern "C" span_b200_ctx_t *span_b200_default_ctx(void)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (g_default_ctx == NULL)
    {
        const char *e = getenv("SPANDSP_B200_DEVICE");
        g_default_ctx = span_b200_ctx_create((e)  ?  atoi(e)  :  0);
    }
    return g_default_ctx;
}

static size_t state_size(int det)
{
    switch (det)
    {
    case SPAN_B200_DET_DTMF: return sizeof(dtmf_rx_state_s);
    case SPAN_B200_DET_BELL_MF: return sizeof(bell_mf_rx_state_s);
    case SPAN_B200_DET_R2_MF: return sizeof(r2_mf_rx_state_s);
    default: return sizeof(super_tone_rx_state_s);
    }
}

static int alloc_stage(span_b200_group_s *g, int max_samples)
{
    sb_device_guard sb_dg_((g)  ?  span_b200_ctx_device(g->ctx)  :  -1);
    int16_t *p = NULL;
    if (cudaMallocHost(&p, sizeof(int16_t)*(size_t) g->members*max_samples) != cudaSuccess)
    {
        sb_set_error("cannot allocate pinned staging for %d members x %d samples", g->members, max_samples);
        return -1;
    }
    if (g->h_stage)
    {
        // keep what is already staged
        for (int i = 0;  i < g->members;  i++)
        {
            const int n = (g->member[i])  ?  g->member[i]->staged  :  0;
            if (n > 0)
                memcpy(p + (size_t) i*max_samples, g->h_stage + (size_t) i*g->max_samples, sizeof(int16_t)*n);
        }
        cudaFreeHost(g->h_stage);
    }
    g->h_stage = p;
    g->max_samples = max_samples;
    return 0;
}

static span_b200_group_s *group_new(span_b200_ctx_t *ctx, int det, int members, int max_samples, int arg,
                                    super_tone_rx_descriptor_t *desc, int single)
{
    if (ctx == NULL)
        ctx = span_b200_default_ctx();
    if (ctx == NULL)
        return NULL;
    if (members < 1  ||  max_samples < 1)
    {
        sb_set_error("bad group size");
        return NULL;
    }
    span_b200_group_s *g = new span_b200_group_s();
    g->ctx = ctx;
    g->det = det;
    g->members = members;
    g->single = single;
    g->arg = arg;
    g->member.assign(members, (sb_member_t *) NULL);
    switch (det)
    {
    case SPAN_B200_DET_DTMF:
        g->bank = span_b200_dtmf_bank_create(ctx, members);
        break;
    case SPAN_B200_DET_BELL_MF:
        g->bank = span_b200_bell_mf_bank_create(ctx, members);
        break;
    case SPAN_B200_DET_R2_MF:
        g->bank = span_b200_r2_mf_bank_create(ctx, members, arg);
        break;
    case SPAN_B200_DET_SUPER_TONE:
        {
            if (desc == NULL)
            {
                sb_set_error("super-tone group needs a descriptor");
                delete g;
                return NULL;
            }
            std::vector<int32_t> segs;
            std::vector<int32_t> el;
            for (size_t t = 0;  t < desc->tones.size();  t++)
            {
                segs.push_back((int32_t) (desc->tones[t].size()/4));
                el.insert(el.end(), desc->tones[t].begin(), desc->tones[t].end());
            }
            if (el.empty())
                el.assign(4, 0);
            if (segs.empty())
                segs.push_back(0);
            span_b200_super_tone_desc_t d;
            d.tones = (int32_t) desc->tones.size();
            d.tone_segs = segs.data();
            d.elements = el.data();
            // Segment events are always produced; the host drops them when no handler is installed.
            g->bank = span_b200_super_tone_bank_create(ctx, members, &d, 1);
        }
        break;
    default:
        sb_set_error("unknown detector %d", det);
        break;
    }
    if (g->bank == NULL)
    {
        delete g;
        return NULL;
    }
    if (alloc_stage(g, max_samples) != 0)
    {
        span_b200_bank_destroy(g->bank);
        delete g;
        return NULL;
    }
    return g;
}

static void member_attach(span_b200_group_s *g, int index, sb_member_t *m, int heap)
{
    m->magic = SB_MAGIC;
    m->grp = g;
    m->index = index;
    m->det = g->det;
    m->staged = 0;
    m->heap = heap;
    g->member[index] = m;
}

extern "C" span_b200_group_t *span_b200_group_create(span_b200_ctx_t *ctx, int detector, int members, int max_samples,
                                                     int arg, super_tone_rx_descriptor_t *desc)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    span_b200_group_s *g = group_new(ctx, detector, members, max_samples, arg, desc, 0);
    if (g == NULL)
        return NULL;
    const size_t sz = state_size(detector);
    for (int i = 0;  i < members;  i++)
    {
        void *p = calloc(1, sz);
        g->owned.push_back(p);
        member_attach(g, i, (sb_member_t *) p, 0);
        if (detector == SPAN_B200_DET_R2_MF)
            ((r2_mf_rx_state_s *) p)->fwd = arg;
        if (detector == SPAN_B200_DET_SUPER_TONE)
            ((super_tone_rx_state_s *) p)->desc = desc;
    }
    return g;
}

extern "C" void *span_b200_group_member(span_b200_group_t *g, int index)
{
    if (g == NULL  ||  index < 0  ||  index >= g->members)
        return NULL;
    return g->member[index];
}

extern "C" void span_b200_group_destroy(span_b200_group_t *g)
{
    sb_device_guard sb_dg_((g)  ?  span_b200_ctx_device(g->ctx)  :  -1);
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (g == NULL)
        return;
    span_b200_bank_destroy(g->bank);
    if (g->h_stage)
        cudaFreeHost(g->h_stage);
    for (size_t i = 0;  i < g->owned.size();  i++)
    {
        ((sb_member_t *) g->owned[i])->magic = 0;
        free(g->owned[i]);
    }
    delete g;
}

// ------------------------------------------------------------------------------------------
// host side of the callbacks

// src/dtmf.c:322-338 / src/bell_r2_mf.c:640-655
template <class S>
static int deliver_digit(S *s, int hit)
{
    int fired = 0;
    if (s->current_digits < 128)
    {
        s->digits[s->current_digits++] = (char) hit;
        s->digits[s->current_digits] = '\0';
        if (s->digits_callback)
        {
            s->digits_callback(s->digits_callback_data, s->digits, s->current_digits);
            s->current_digits = 0;
            fired = 1;
        }
    }
    else
    {
        s->lost_digits++;
    }
    return fired;
}

// src/dtmf.c:352-358 / src/bell_r2_mf.c:667-672: digits buffered before a callback was installed
template <class S>
static int trailing_flush(S *s)
{
    if (s->current_digits  &&  s->digits_callback)
    {
        s->digits_callback(s->digits_callback_data, s->digits, s->current_digits);
        s->digits[0] = '\0';
        s->current_digits = 0;
        return 1;
    }
    return 0;
}

static int dispatch(span_b200_group_s *g, const span_b200_event_t &e)
{
    sb_member_t *m = g->member[e.channel];
    if (m == NULL)
        return 0;
    switch (g->det)
    {
    case SPAN_B200_DET_DTMF:
        {
            dtmf_rx_state_s *s = (dtmf_rx_state_s *) m;
            if (e.kind == SPAN_B200_EV_TONE)
            {
                if (s->realtime_callback)
                {
                    s->realtime_callback(s->realtime_callback_data, e.a, e.b, e.c);
                    return 1;
                }
                return 0;
            }
            return deliver_digit(s, e.a);
        }
    case SPAN_B200_DET_BELL_MF:
        return deliver_digit((bell_mf_rx_state_s *) m, e.a);
    case SPAN_B200_DET_R2_MF:
        {
            r2_mf_rx_state_s *s = (r2_mf_rx_state_s *) m;
            if (s->callback)
            {
                s->callback(s->callback_data, e.a, e.b, e.c);
                return 1;
            }
            return 0;
        }
    default:
        {
            super_tone_rx_state_s *s = (super_tone_rx_state_s *) m;
            if (e.kind == SPAN_B200_EV_SEGMENT)
            {
                if (s->segment_callback)
                {
                    s->segment_callback(s->callback_data, e.a, e.b, e.c);
                    return 1;
                }
                return 0;
            }
            if (s->tone_callback)
            {
                s->tone_callback(s->callback_data, e.a, e.b, e.c);
                return 1;
            }
            return 0;
        }
    }
}

extern "C" int span_b200_group_flush(span_b200_group_t *g)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (g == NULL)
        return -1;
    int n = 0;
    for (int i = 0;  i < g->members;  i++)
    {
        const int st = (g->member[i])  ?  g->member[i]->staged  :  0;
        if (st > n)
            n = st;
    }
    if (n == 0)
        return 0;
    for (int i = 0;  i < g->members;  i++)
    {
        const int st = (g->member[i])  ?  g->member[i]->staged  :  0;
        if (st != n)
        {
            sb_set_error("group flush: member %d staged %d samples, others %d (all members must be fed the same amount per flush)", i, st, n);
            return -1;
        }
    }
    if (span_b200_bank_rx_host(g->bank, g->h_stage, g->max_samples, n, NULL) != 0)
        return -1;
    int overflow = 0;
    const int64_t total = span_b200_bank_event_count(g->bank, &overflow);
    if (total < 0)
        return -1;
    if (overflow)
    {
        // Cannot happen with the default (worst case) event capacity; a caller-set capacity was too small.
        // The reports that fit are still delivered below, but the flush is reported as failed.
        sb_set_error("group flush: the bank's event buffer overflowed; reports were lost");
    }
    g->events.resize((size_t) total);
    if (total > 0  &&  span_b200_bank_events(g->bank, g->events.data(), total) < 0)
        return -1;
    // Fire callbacks channel by channel, each channel's events in time order.
    std::stable_sort(g->events.begin(), g->events.end(),
                     [](const span_b200_event_t &a, const span_b200_event_t &b) { return a.channel < b.channel; });
    int fired = 0;
    for (size_t i = 0;  i < g->events.size();  i++)
        fired += dispatch(g, g->events[i]);
    for (int i = 0;  i < g->members;  i++)
    {
        sb_member_t *m = g->member[i];
        if (m == NULL)
            continue;
        if (g->det == SPAN_B200_DET_DTMF  &&  ((dtmf_rx_state_s *) m)->realtime_callback == NULL)
            fired += trailing_flush((dtmf_rx_state_s *) m);
        else if (g->det == SPAN_B200_DET_BELL_MF)
            fired += trailing_flush((bell_mf_rx_state_s *) m);
        m->staged = 0;
    }
    return (overflow)  ?  -1  :  fired;
}

static int stage_samples(sb_member_t *m, const int16_t amp[], int samples)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    span_b200_group_s *g = m->grp;
    if (samples <= 0)
        return 0;
    if (m->staged + samples > g->max_samples)
    {
        // More audio than the staging area holds (a group that is flushed less often than its max_samples
        // suggests, or a long call on a plain state): grow it, keeping what every member has staged.  Audio is
        // never dropped; the only failure left is running out of pinned memory.
        if (alloc_stage(g, std::max(2*g->max_samples, m->staged + samples)) != 0)
            return -1;
    }
    memcpy(g->h_stage + (size_t) m->index*g->max_samples + m->staged, amp, sizeof(int16_t)*samples);
    m->staged += samples;
    if (g->single)
        return (span_b200_group_flush(g) < 0)  ?  -1  :  0;
    return 0;
}

// Make (or re-use) the backing of a state object.  `s` NULL: allocate.  `s` a live member: keep
// its slot, reset its channel.  Anything else is taken as caller storage: a new group of one.
static sb_member_t *state_open(void *s, int det, int arg, super_tone_rx_descriptor_t *desc)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    sb_member_t *m = (sb_member_t *) s;
    if (m != NULL  &&  m->magic == SB_MAGIC  &&  m->grp != NULL  &&  m->det == det
        &&  m->index >= 0  &&  m->index < m->grp->members  &&  m->grp->member[m->index] == m
        &&  !(det == SPAN_B200_DET_R2_MF  &&  m->grp->single  &&  m->grp->arg != arg)
        &&  !(det == SPAN_B200_DET_SUPER_TONE  &&  m->grp->single))
    {
        if (m->staged > 0  &&  !m->grp->single)
            m->staged = 0;
        span_b200_bank_reset(m->grp->bank, m->index, 1);
        return m;
    }
    if (m != NULL  &&  m->magic == SB_MAGIC  &&  m->grp != NULL  &&  m->grp->single  &&  m->det == det)
    {
        // same storage, different construction arguments: rebuild the backing
        span_b200_group_s *old = m->grp;
        m->magic = 0;
        span_b200_group_destroy(old);
    }
    int heap = 0;
    if (m == NULL)
    {
        m = (sb_member_t *) calloc(1, state_size(det));
        if (m == NULL)
            return NULL;
        heap = 1;
    }
    span_b200_group_s *g = group_new(NULL, det, 1, 1024, arg, desc, 1);
    if (g == NULL)
    {
        if (heap)
            free(m);
        return NULL;
    }
    memset(m, 0, state_size(det));
    member_attach(g, 0, m, heap);
    return m;
}

static int state_close(void *s, int do_free)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    sb_member_t *m = (sb_member_t *) s;
    if (m == NULL  ||  m->magic != SB_MAGIC)
        return 0;
    span_b200_group_s *g = m->grp;
    const int heap = m->heap;
    if (g  &&  g->single)
    {
        m->magic = 0;
        span_b200_group_destroy(g);
        if (do_free  &&  heap)
            free(m);
    }
    return 0;
}

static int flush_if_pending(sb_member_t *m)
{
    // Control calls act on the channel "now": in a batched group anything staged is processed first.
    if (m->grp  &&  m->staged > 0)
        return span_b200_group_flush(m->grp);
    return 0;
}

// ------------------------------------------------------------------------------------------
// DTMF  (src/dtmf.c:363-519)
extern "C" dtmf_rx_state_t *dtmf_rx_init(dtmf_rx_state_t *s, digits_rx_callback_t callback, void *user_data)
{
    sb_member_t *m = state_open(s, SPAN_B200_DET_DTMF, 0, NULL);
    if (m == NULL)
        return NULL;
    dtmf_rx_state_s *d = (dtmf_rx_state_s *) m;
    d->digits_callback = callback;
    d->digits_callback_data = user_data;
    d->realtime_callback = NULL;
    d->realtime_callback_data = NULL;
    d->lost_digits = 0;
    d->current_digits = 0;
    d->digits[0] = '\0';
    return d;
}

extern "C" int dtmf_rx_release(dtmf_rx_state_t *s)
{
    return state_close(s, 0);
}

extern "C" int dtmf_rx_free(dtmf_rx_state_t *s)
{
    return state_close(s, 1);
}

extern "C" void dtmf_rx_set_realtime_callback(dtmf_rx_state_t *s, span_tone_report_func_t callback, void *user_data)
{
    flush_if_pending(&s->m);
    s->realtime_callback = callback;
    s->realtime_callback_data = user_data;
    span_b200_dtmf_bank_realtime(s->m.grp->bank, s->m.index, 1, callback != NULL);
}

extern "C" void dtmf_rx_parms(dtmf_rx_state_t *s, int filter_dialtone, float twist, float reverse_twist, float threshold)
{
    flush_if_pending(&s->m);
    span_b200_dtmf_bank_parms(s->m.grp->bank, s->m.index, 1, filter_dialtone, twist, reverse_twist, threshold);
}

extern "C" int dtmf_rx(dtmf_rx_state_t *s, const int16_t amp[], int samples)
{
    // "The number of samples unprocessed" (src/spandsp/dtmf.h:176): 0 as in src/dtmf.c:360, or all of them when
    // the device path failed (span_b200_last_error() says why)
    return (stage_samples(&s->m, amp, samples) == 0)  ?  0  :  samples;
}

extern "C" int dtmf_rx_fillin(dtmf_rx_state_t *s, int samples)
{
    (void) samples;
    flush_if_pending(&s->m);
    span_b200_dtmf_bank_fillin(s->m.grp->bank, s->m.index, 1);
    return 0;
}

extern "C" int dtmf_rx_status(dtmf_rx_state_t *s)
{
    int32_t st = 0;
    flush_if_pending(&s->m);
    span_b200_bank_status(s->m.grp->bank, s->m.index, 1, &st);
    return st;
}

template <class S>
static size_t digits_get(S *s, char *buf, int max)
{
    // src/dtmf.c:394-408
    if (max > s->current_digits)
        max = s->current_digits;
    if (max > 0)
    {
        memcpy(buf, s->digits, max);
        memmove(s->digits, s->digits + max, s->current_digits - max);
        s->current_digits -= max;
    }
    buf[max] = '\0';
    return max;
}

extern "C" size_t dtmf_rx_get(dtmf_rx_state_t *s, char *buf, int max)
{
    flush_if_pending(&s->m);
    return digits_get(s, buf, max);
}

extern "C" logging_state_t *dtmf_rx_get_logging_state(dtmf_rx_state_t *s)
{
    (void) s;
    return NULL;        // the "Potentially 'x'" debug trace of src/dtmf.c:260-275 is not produced
}

// ------------------------------------------------------------------------------------------
// Bell MF  (src/bell_r2_mf.c:676-745)
extern "C" bell_mf_rx_state_t *bell_mf_rx_init(bell_mf_rx_state_t *s, digits_rx_callback_t callback, void *user_data)
{
    sb_member_t *m = state_open(s, SPAN_B200_DET_BELL_MF, 0, NULL);
    if (m == NULL)
        return NULL;
    bell_mf_rx_state_s *d = (bell_mf_rx_state_s *) m;
    d->digits_callback = callback;
    d->digits_callback_data = user_data;
    d->lost_digits = 0;
    d->current_digits = 0;
    d->digits[0] = '\0';
    return d;
}

extern "C" int bell_mf_rx_release(bell_mf_rx_state_t *s) { return state_close(s, 0); }
extern "C" int bell_mf_rx_free(bell_mf_rx_state_t *s) { return state_close(s, 1); }

extern "C" int bell_mf_rx(bell_mf_rx_state_t *s, const int16_t amp[], int samples)
{
    return (stage_samples(&s->m, amp, samples) == 0)  ?  0  :  samples;         // unprocessed samples
}

extern "C" size_t bell_mf_rx_get(bell_mf_rx_state_t *s, char *buf, int max)
{
    flush_if_pending(&s->m);
    return digits_get(s, buf, max);
}

// ------------------------------------------------------------------------------------------
// MFC/R2  (src/bell_r2_mf.c:883-951)
extern "C" r2_mf_rx_state_t *r2_mf_rx_init(r2_mf_rx_state_t *s, bool fwd, span_tone_report_func_t callback, void *user_data)
{
    sb_member_t *m = state_open(s, SPAN_B200_DET_R2_MF, (fwd)  ?  1  :  0, NULL);
    if (m == NULL)
        return NULL;
    r2_mf_rx_state_s *d = (r2_mf_rx_state_s *) m;
    d->callback = callback;
    d->callback_data = user_data;
    d->fwd = (fwd)  ?  1  :  0;
    return d;
}

extern "C" int r2_mf_rx_release(r2_mf_rx_state_t *s) { return state_close(s, 0); }
extern "C" int r2_mf_rx_free(r2_mf_rx_state_t *s) { return state_close(s, 1); }

extern "C" int r2_mf_rx(r2_mf_rx_state_t *s, const int16_t amp[], int samples)
{
    return (stage_samples(&s->m, amp, samples) == 0)  ?  0  :  samples;         // unprocessed samples
}

extern "C" int r2_mf_rx_get(r2_mf_rx_state_t *s)
{
    int32_t st = 0;
    flush_if_pending(&s->m);
    span_b200_bank_status(s->m.grp->bank, s->m.index, 1, &st);
    return st;
}

// ------------------------------------------------------------------------------------------
// supervisory tones  (src/super_tone_rx.c:125-161,231-287,454-566)
extern "C" super_tone_rx_descriptor_t *super_tone_rx_make_descriptor(super_tone_rx_descriptor_t *desc)
{
    if (desc == NULL)
    {
        desc = new (std::nothrow) super_tone_rx_descriptor_s();
        if (desc == NULL)
            return NULL;
        desc->heap = 1;
    }
    else
    {
        new (desc) super_tone_rx_descriptor_s();
        desc->heap = 0;
    }
    return desc;
}

extern "C" int super_tone_rx_free_descriptor(super_tone_rx_descriptor_t *desc)
{
    if (desc)
    {
        if (desc->heap)
            delete desc;
        else
            desc->~super_tone_rx_descriptor_s();
    }
    return 0;
}

extern "C" int super_tone_rx_add_tone(super_tone_rx_descriptor_t *desc)
{
    desc->tones.push_back(std::vector<int>());
    return (int) desc->tones.size() - 1;
}

extern "C" int super_tone_rx_add_element(super_tone_rx_descriptor_t *desc, int tone, int f1, int f2, int min, int max)
{
    if (tone < 0  ||  tone >= (int) desc->tones.size())
        return -1;
    std::vector<int> &t = desc->tones[tone];
    t.push_back(f1);
    t.push_back(f2);
    t.push_back(min);
    t.push_back(max);
    return (int) t.size()/4 - 1;
}

extern "C" super_tone_rx_state_t *super_tone_rx_init(super_tone_rx_state_t *s, super_tone_rx_descriptor_t *desc,
                                                     span_tone_report_func_t callback, void *user_data)
{
    if (desc == NULL  ||  callback == NULL)     // src/super_tone_rx.c:514-519
        return NULL;
    sb_member_t *m = state_open(s, SPAN_B200_DET_SUPER_TONE, 0, desc);
    if (m == NULL)
        return NULL;
    super_tone_rx_state_s *d = (super_tone_rx_state_s *) m;
    d->desc = desc;
    d->tone_callback = callback;
    d->segment_callback = NULL;
    d->callback_data = user_data;
    return d;
}

extern "C" int super_tone_rx_release(super_tone_rx_state_t *s) { return state_close(s, 0); }
extern "C" int super_tone_rx_free(super_tone_rx_state_t *s) { return state_close(s, 1); }

extern "C" void super_tone_rx_tone_callback(super_tone_rx_state_t *s, span_tone_report_func_t callback, void *user_data)
{
    s->tone_callback = callback;
    s->callback_data = user_data;
}

extern "C" void super_tone_rx_segment_callback(super_tone_rx_state_t *s, tone_segment_func_t callback)
{
    s->segment_callback = callback;
}

extern "C" int super_tone_rx(super_tone_rx_state_t *s, const int16_t amp[], int samples)
{
    // "The number of samples processed" (src/spandsp/super_tone_rx.h:154): all of them (src/super_tone_rx.c:489), or
    // none when the device path failed
    return (stage_samples(&s->m, amp, samples) == 0)  ?  samples  :  0;
}

extern "C" int super_tone_rx_fillin(super_tone_rx_state_t *s, int samples)
{
    (void) s;
    (void) samples;
    return 0;                                   // src/super_tone_rx.c:493-497
}

// ------------------------------------------------------------------------------------------
// Goertzel primitives on a caller-owned public structure (src/tone_detect.c:60-205).
// One device thread runs the recurrence so that there is a single arithmetic implementation.
__global__ void goertzel_update_one(float *st, const int16_t *amp, int n)
{
    if (threadIdx.x != 0  ||  blockIdx.x != 0)
        return;
    float v2 = st[0];
    float v3 = st[1];
    const float fac = st[2];
    for (int i = 0;  i < n;  i++)
    {
        const float v1 = v2;
        v2 = v3;
        v3 = __fadd_rn(__fsub_rn(__fmul_rn(fac, v2), v1), (float) amp[i]);      // src/tone_detect.c:141-151
    }
    st[0] = v2;
    st[1] = v3;
}

__global__ void goertzel_result_one(float *st)
{
    if (threadIdx.x != 0  ||  blockIdx.x != 0)
        return;
    float v1 = st[0];
    float v2 = st[1];
    const float fac = st[2];
    const float v3 = __fsub_rn(__fmul_rn(fac, v2), v1);                         // src/tone_detect.c:174-181
    v1 = __fsub_rn(__fadd_rn(__fmul_rn(v3, v3), __fmul_rn(v2, v2)), __fmul_rn(__fmul_rn(v2, v3), fac));
    st[3] = __fmul_rn(v1, 2.0f);                                                // src/tone_detect.c:200-201
}

extern "C" void make_goertzel_descriptor(goertzel_descriptor_t *t, float freq, int samples)
{
    t->fac = sb_goertzel_fac(freq);
    t->samples = samples;
}

extern "C" goertzel_state_t *goertzel_init(goertzel_state_t *s, goertzel_descriptor_t *t)
{
    if (s == NULL)
    {
        if ((s = (goertzel_state_t *) malloc(sizeof(*s))) == NULL)
            return NULL;
    }
    s->v2 = 0.0f;
    s->v3 = 0.0f;
    s->fac = t->fac;
    s->samples = t->samples;
    s->current_sample = 0;
    return s;
}

extern "C" int goertzel_release(goertzel_state_t *s)
{
    (void) s;
    return 0;
}

extern "C" int goertzel_free(goertzel_state_t *s)
{
    if (s)
        free(s);
    return 0;
}

extern "C" void goertzel_reset(goertzel_state_t *s)
{
    s->v2 = 0.0f;
    s->v3 = 0.0f;
    s->current_sample = 0;
}

static float *g_gz_state = NULL;
static int16_t *g_gz_amp = NULL;
static int g_gz_cap = 0;

static int gz_prepare(int n)
{
    span_b200_ctx_t *ctx = span_b200_default_ctx();
    if (ctx == NULL)
        return -1;
    sb_device_guard sb_dg_(span_b200_ctx_device(ctx));
    if (g_gz_state == NULL  &&  cudaMalloc(&g_gz_state, sizeof(float)*4) != cudaSuccess)
        return -1;
    if (n > g_gz_cap)
    {
        if (g_gz_amp)
            cudaFree(g_gz_amp);
        g_gz_cap = std::max(n, 1024);
        if (cudaMalloc(&g_gz_amp, sizeof(int16_t)*g_gz_cap) != cudaSuccess)
        {
            g_gz_cap = 0;
            g_gz_amp = NULL;
            return -1;
        }
    }
    return 0;
}

extern "C" int goertzel_update(goertzel_state_t *s, const int16_t amp[], int samples)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    // src/tone_detect.c:135-137: never run past the end of the block
    if (samples > s->samples - s->current_sample)
        samples = s->samples - s->current_sample;
    if (samples <= 0)
        return 0;
    if (gz_prepare(samples) != 0)
    {
        sb_set_error("goertzel_update: no CUDA device (there is no CPU fallback)");
        return 0;
    }
    sb_device_guard sb_dg_(span_b200_ctx_device(span_b200_default_ctx()));
    cudaStream_t st = (cudaStream_t) sb_ctx_stream(span_b200_default_ctx());
    const float h[3] = {s->v2, s->v3, s->fac};
    cudaMemcpyAsync(g_gz_state, h, sizeof(h), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(g_gz_amp, amp, sizeof(int16_t)*samples, cudaMemcpyHostToDevice, st);
    goertzel_update_one<<<1, 32, 0, st>>>(g_gz_state, g_gz_amp, samples);
    float o[2];
    cudaMemcpyAsync(o, g_gz_state, sizeof(o), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess)
    {
        sb_set_error("goertzel_update: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    s->v2 = o[0];
    s->v3 = o[1];
    s->current_sample += samples;
    return samples;
}

extern "C" float goertzel_result(goertzel_state_t *s)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (gz_prepare(0) != 0)
    {
        sb_set_error("goertzel_result: no CUDA device (there is no CPU fallback)");
        return 0.0f;
    }
    sb_device_guard sb_dg_(span_b200_ctx_device(span_b200_default_ctx()));
    cudaStream_t st = (cudaStream_t) sb_ctx_stream(span_b200_default_ctx());
    const float h[3] = {s->v2, s->v3, s->fac};
    cudaMemcpyAsync(g_gz_state, h, sizeof(h), cudaMemcpyHostToDevice, st);
    goertzel_result_one<<<1, 32, 0, st>>>(g_gz_state);
    float o = 0.0f;
    cudaMemcpyAsync(&o, g_gz_state + 3, sizeof(o), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    goertzel_reset(s);                          // src/tone_detect.c:203
    return o;
}

// ------------------------------------------------------------------------------------------
// V.29 receiver, one bank of one per state object (src/v29rx.c:145-195,867-1153)
struct v29_rx_state_s
{
    unsigned long long magic;
    span_b200_v29_bank_t *bank;
    int heap;
    int bit_rate;
    span_put_bit_func_t put_bit;
    void *put_bit_user_data;
    span_modem_status_func_t status_handler;
    void *status_user_data;
    qam_report_handler_t qam_report;
    void *qam_user_data;
    complexf_t eq_coeff[33];
    std::vector<int8_t> *bits;
    std::vector<span_b200_v29_symbol_t> *syms;
};

static_assert(sizeof(v29_rx_state_s) <= 1272, "must fit the reference's v29_rx_state_t (private/v29rx.h)");

extern "C" v29_rx_state_t *v29_rx_init(v29_rx_state_t *s, int bit_rate, span_put_bit_func_t put_bit, void *user_data)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (bit_rate != 9600  &&  bit_rate != 7200  &&  bit_rate != 4800)
        return NULL;                                        // src/v29rx.c:1102-1111
    span_b200_ctx_t *ctx = span_b200_default_ctx();
    if (ctx == NULL)
        return NULL;
    int heap = 0;
    if (s != NULL  &&  s->magic == SB_MAGIC  &&  s->bank != NULL)
    {
        span_b200_v29_bank_destroy(s->bank);
        delete s->bits;
        delete s->syms;
        heap = s->heap;
    }
    else if (s == NULL)
    {
        if ((s = (v29_rx_state_t *) calloc(1, sizeof(*s))) == NULL)
            return NULL;
        heap = 1;
    }
    memset(s, 0, sizeof(*s));
    s->bank = span_b200_v29_bank_create(ctx, 1, bit_rate, 1);
    if (s->bank == NULL)
    {
        if (heap)
            free(s);
        return NULL;
    }
    s->magic = SB_MAGIC;
    s->heap = heap;
    s->bit_rate = bit_rate;
    s->put_bit = put_bit;
    s->put_bit_user_data = user_data;
    s->bits = new std::vector<int8_t>();
    s->syms = new std::vector<span_b200_v29_symbol_t>();
    return s;
}

extern "C" int v29_rx_restart(v29_rx_state_t *s, int bit_rate, bool old_train)
{
    if (bit_rate != 9600  &&  bit_rate != 7200  &&  bit_rate != 4800)
        return -1;                                          // src/v29rx.c:1033
    s->bit_rate = bit_rate;
    return span_b200_v29_bank_restart_ex(s->bank, 0, 1, bit_rate, (old_train)  ?  1  :  0);
}

static int v29_close(v29_rx_state_t *s, int do_free)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (s == NULL  ||  s->magic != SB_MAGIC)
        return 0;
    span_b200_v29_bank_destroy(s->bank);
    delete s->bits;
    delete s->syms;
    const int heap = s->heap;
    s->magic = 0;
    s->bank = NULL;
    if (do_free  &&  heap)
        free(s);
    return 0;
}

extern "C" int v29_rx_release(v29_rx_state_t *s) { return v29_close(s, 0); }
extern "C" int v29_rx_free(v29_rx_state_t *s) { return v29_close(s, 1); }
extern "C" logging_state_t *v29_rx_get_logging_state(v29_rx_state_t *s) { (void) s; return NULL; }

extern "C" void v29_rx_set_put_bit(v29_rx_state_t *s, span_put_bit_func_t put_bit, void *user_data)
{
    s->put_bit = put_bit;
    s->put_bit_user_data = user_data;
}

extern "C" void v29_rx_set_modem_status_handler(v29_rx_state_t *s, span_modem_status_func_t handler, void *user_data)
{
    s->status_handler = handler;
    s->status_user_data = user_data;
}

extern "C" void v29_rx_set_qam_report_handler(v29_rx_state_t *s, qam_report_handler_t handler, void *user_data)
{
    s->qam_report = handler;
    s->qam_user_data = user_data;
}

// One entry of the put_bit stream: data bits go to put_bit, status codes to the status handler if one
// is installed, else to put_bit (src/v29rx.c:171-178).
static void v29_deliver(v29_rx_state_t *s, int v)
{
    if (v < 0  &&  s->status_handler)
        s->status_handler(s->status_user_data, v);
    else if (s->put_bit)
        s->put_bit(s->put_bit_user_data, v);
}

extern "C" int v29_rx(v29_rx_state_t *s, const int16_t amp[], int len)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (len <= 0)
        return 0;
    if (span_b200_v29_bank_rx_host(s->bank, amp, len, len, NULL) != 0)
        return 0;
    int32_t nb = 0;
    int32_t ns = 0;
    if (span_b200_v29_bank_counts(s->bank, &nb, &ns) != 0)
        return 0;
    s->bits->resize((size_t) std::max(nb, 1));
    s->syms->resize((size_t) std::max(ns, 1));
    if (nb > 0)
        span_b200_v29_bank_bits(s->bank, 0, s->bits->data(), nb);
    if (ns > 0)
        span_b200_v29_bank_symbols(s->bank, 0, s->syms->data(), ns);
    // Replay in the reference's order: the bits of a baud come before that baud's qam report.
    int done = 0;
    for (int k = 0;  k < ns;  k++)
    {
        const span_b200_v29_symbol_t &y = (*s->syms)[k];
        for (  ;  done < y.bit_pos  &&  done < nb;  done++)
            v29_deliver(s, (*s->bits)[done]);
        if (s->qam_report)
        {
            const complexf_t z = {y.re, y.im};
            const complexf_t t = {y.target_re, y.target_im};
            s->qam_report(s->qam_user_data, &z, &t, y.state);
        }
    }
    for (  ;  done < nb;  done++)
        v29_deliver(s, (*s->bits)[done]);
    return 0;                                               // src/v29rx.c:963
}

extern "C" int v29_rx_fillin(v29_rx_state_t *s, int len)
{
    return (span_b200_v29_bank_fillin(s->bank, 0, 1, len) == 0)  ?  0  :  0;
}

extern "C" int v29_rx_equalizer_state(v29_rx_state_t *s, complexf_t **coeffs)
{
    float eq[66];
    span_b200_v29_bank_channel_state(s->bank, 0, eq, NULL);
    for (int i = 0;  i < 33;  i++)
    {
        s->eq_coeff[i].re = eq[2*i];
        s->eq_coeff[i].im = eq[2*i + 1];
    }
    *coeffs = s->eq_coeff;
    return 33;                                              // V29_EQUALIZER_LEN
}

extern "C" float v29_rx_carrier_frequency(v29_rx_state_t *s)
{
    int32_t info[10];
    span_b200_v29_bank_channel_state(s->bank, 0, NULL, info);
    return (float) info[1]*(float) 8000/(65536.0f*65536.0f);           // dds_frequencyf, src/dds_float.c:2115-2118
}

extern "C" float v29_rx_symbol_timing_correction(v29_rx_state_t *s)
{
    int32_t info[10];
    span_b200_v29_bank_channel_state(s->bank, 0, NULL, info);
    return (float) info[5]/((float) 48*10.0f/3.0f);                    // src/v29rx.c:151-154
}

extern "C" float v29_rx_signal_power(v29_rx_state_t *s)
{
    int32_t info[10];
    span_b200_v29_bank_channel_state(s->bank, 0, NULL, info);
    // power_meter_current_dbm0() + 3.98f: src/power_meter.c:115-122, src/v29rx.c:157-160
    if (info[8] <= 0)
        return (-96.329f + (3.14f + 3.02f)) + 3.98f;
    return (10.0f*log10f((float) info[8]/(32767.0f*32767.0f) + 1.0e-10f) + (3.14f + 3.02f)) + 3.98f;
}

extern "C" void v29_rx_set_signal_cutoff(v29_rx_state_t *s, float cutoff)
{
    span_b200_v29_bank_set_signal_cutoff(s->bank, 0, 1, cutoff);
}

// ------------------------------------------------------------------------------------------
// V.17 receiver, one bank of one per state object (src/v17rx.c:157-204,1214-1541)
struct v17_rx_state_s
{
    unsigned long long magic;
    span_b200_v17_bank_t *bank;
    int heap;
    int bit_rate;
    span_put_bit_func_t put_bit;
    void *put_bit_user_data;
    span_modem_status_func_t status_handler;
    void *status_user_data;
    qam_report_handler_t qam_report;
    void *qam_user_data;
    complexf_t eq_coeff[33];
    std::vector<int8_t> *bits;
    std::vector<span_b200_v17_symbol_t> *syms;
};

static_assert(sizeof(v17_rx_state_s) <= 2352, "must fit the reference's v17_rx_state_t (private/v29rx.h)");

extern "C" v17_rx_state_t *v17_rx_init(v17_rx_state_t *s, int bit_rate, span_put_bit_func_t put_bit, void *user_data)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (bit_rate != 14400  &&  bit_rate != 12000  &&  bit_rate != 9600  &&  bit_rate != 7200  &&  bit_rate != 4800)
        return NULL;                                        // src/v17rx.c:1498-1510
    span_b200_ctx_t *ctx = span_b200_default_ctx();
    if (ctx == NULL)
        return NULL;
    int heap = 0;
    if (s != NULL  &&  s->magic == SB_MAGIC  &&  s->bank != NULL)
    {
        span_b200_v17_bank_destroy(s->bank);
        delete s->bits;
        delete s->syms;
        heap = s->heap;
    }
    else if (s == NULL)
    {
        if ((s = (v17_rx_state_t *) calloc(1, sizeof(*s))) == NULL)
            return NULL;
        heap = 1;
    }
    memset(s, 0, sizeof(*s));
    s->bank = span_b200_v17_bank_create(ctx, 1, bit_rate, 1);
    if (s->bank == NULL)
    {
        if (heap)
            free(s);
        return NULL;
    }
    s->magic = SB_MAGIC;
    s->heap = heap;
    s->bit_rate = bit_rate;
    s->put_bit = put_bit;
    s->put_bit_user_data = user_data;
    s->bits = new std::vector<int8_t>();
    s->syms = new std::vector<span_b200_v17_symbol_t>();
    return s;
}

extern "C" int v17_rx_restart(v17_rx_state_t *s, int bit_rate, int short_train)
{
    if (bit_rate != 14400  &&  bit_rate != 12000  &&  bit_rate != 9600  &&  bit_rate != 7200  &&  bit_rate != 4800)
        return -1;                                          // src/v17rx.c:1425
    s->bit_rate = bit_rate;
    return span_b200_v17_bank_restart(s->bank, 0, 1, bit_rate, short_train);
}

static int v17_close(v17_rx_state_t *s, int do_free)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (s == NULL  ||  s->magic != SB_MAGIC)
        return 0;
    span_b200_v17_bank_destroy(s->bank);
    delete s->bits;
    delete s->syms;
    const int heap = s->heap;
    s->magic = 0;
    s->bank = NULL;
    if (do_free  &&  heap)
        free(s);
    return 0;
}

extern "C" int v17_rx_release(v17_rx_state_t *s) { return v17_close(s, 0); }
extern "C" int v17_rx_free(v17_rx_state_t *s) { return v17_close(s, 1); }
extern "C" logging_state_t *v17_rx_get_logging_state(v17_rx_state_t *s) { (void) s; return NULL; }

extern "C" void v17_rx_set_put_bit(v17_rx_state_t *s, span_put_bit_func_t put_bit, void *user_data)
{
    s->put_bit = put_bit;
    s->put_bit_user_data = user_data;
}

extern "C" void v17_rx_set_modem_status_handler(v17_rx_state_t *s, span_modem_status_func_t handler, void *user_data)
{
    s->status_handler = handler;
    s->status_user_data = user_data;
}

extern "C" void v17_rx_set_qam_report_handler(v17_rx_state_t *s, qam_report_handler_t handler, void *user_data)
{
    s->qam_report = handler;
    s->qam_user_data = user_data;
}

// One entry of the put_bit stream: data bits go to put_bit, status codes to the status handler if one
// is installed, else to put_bit (src/v17rx.c:181-189).
static void v17_deliver(v17_rx_state_t *s, int v)
{
    if (v < 0  &&  s->status_handler)
        s->status_handler(s->status_user_data, v);
    else if (s->put_bit)
        s->put_bit(s->put_bit_user_data, v);
}

extern "C" int v17_rx(v17_rx_state_t *s, const int16_t amp[], int len)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (len <= 0)
        return 0;
    if (span_b200_v17_bank_rx_host(s->bank, amp, len, len, NULL) != 0)
        return 0;
    int32_t nb = 0;
    int32_t ns = 0;
    if (span_b200_v17_bank_counts(s->bank, &nb, &ns) != 0)
        return 0;
    s->bits->resize((size_t) std::max(nb, 1));
    s->syms->resize((size_t) std::max(ns, 1));
    if (nb > 0)
        span_b200_v17_bank_bits(s->bank, 0, s->bits->data(), nb);
    if (ns > 0)
        span_b200_v17_bank_symbols(s->bank, 0, s->syms->data(), ns);
    // Replay in the reference's order: the bits of a baud come before that baud's qam report.
    int done = 0;
    for (int k = 0;  k < ns;  k++)
    {
        const span_b200_v17_symbol_t &y = (*s->syms)[k];
        for (  ;  done < y.bit_pos  &&  done < nb;  done++)
            v17_deliver(s, (*s->bits)[done]);
        if (s->qam_report)
        {
            const complexf_t z = {y.re, y.im};
            const complexf_t t = {y.target_re, y.target_im};
            s->qam_report(s->qam_user_data, &z, &t, y.state);
        }
    }
    for (  ;  done < nb;  done++)
        v17_deliver(s, (*s->bits)[done]);
    return 0;                                               // src/v17rx.c:1310
}

extern "C" int v17_rx_fillin(v17_rx_state_t *s, int len)
{
    return (span_b200_v17_bank_fillin(s->bank, 0, 1, len) == 0)  ?  0  :  0;
}

extern "C" int v17_rx_equalizer_state(v17_rx_state_t *s, complexf_t **coeffs)
{
    float eq[66];
    span_b200_v17_bank_channel_state(s->bank, 0, eq, NULL);
    for (int i = 0;  i < 33;  i++)
    {
        s->eq_coeff[i].re = eq[2*i];
        s->eq_coeff[i].im = eq[2*i + 1];
    }
    *coeffs = s->eq_coeff;
    return 33;                                              // V17_EQUALIZER_LEN
}

extern "C" float v17_rx_carrier_frequency(v17_rx_state_t *s)
{
    int32_t info[12];
    span_b200_v17_bank_channel_state(s->bank, 0, NULL, info);
    return (float) info[1]*(float) 8000/(65536.0f*65536.0f);           // dds_frequencyf, src/dds_float.c:2115-2118
}

extern "C" float v17_rx_symbol_timing_correction(v17_rx_state_t *s)
{
    int32_t info[12];
    span_b200_v17_bank_channel_state(s->bank, 0, NULL, info);
    return (float) info[5]/((float) 192*10.0f/3.0f);                   // src/v17rx.c:163-166
}

extern "C" float v17_rx_signal_power(v17_rx_state_t *s)
{
    int32_t info[12];
    span_b200_v17_bank_channel_state(s->bank, 0, NULL, info);
    // power_meter_current_dbm0() + 3.98f: src/power_meter.c:115-122, src/v17rx.c:169-172
    if (info[8] <= 0)
        return (-96.329f + (3.14f + 3.02f)) + 3.98f;
    return (10.0f*log10f((float) info[8]/(32767.0f*32767.0f) + 1.0e-10f) + (3.14f + 3.02f)) + 3.98f;
}

extern "C" void v17_rx_set_signal_cutoff(v17_rx_state_t *s, float cutoff)
{
    span_b200_v17_bank_set_signal_cutoff(s->bank, 0, 1, cutoff);
}

// ------------------------------------------------------------------------------------------
// V.27ter receiver, one bank of one per state object (src/v27ter_rx.c:137-189,863-1210)
struct v27ter_rx_state_s
{
    unsigned long long magic;
    span_b200_v27ter_bank_t *bank;
    int heap;
    int bit_rate;
    span_put_bit_func_t put_bit;
    void *put_bit_user_data;
    span_modem_status_func_t status_handler;
    void *status_user_data;
    qam_report_handler_t qam_report;
    void *qam_user_data;
    complexf_t eq_coeff[32];
    std::vector<int8_t> *bits;
    std::vector<span_b200_v27ter_symbol_t> *syms;
};

static_assert(sizeof(v27ter_rx_state_s) <= 1184, "must fit the reference's v27ter_rx_state_t (private/v27ter_rx.h)");

extern "C" v27ter_rx_state_t *v27ter_rx_init(v27ter_rx_state_t *s, int bit_rate, span_put_bit_func_t put_bit, void *user_data)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (bit_rate != 4800  &&  bit_rate != 2400)
        return NULL;                                        // src/v27ter_rx.c:1163-1171
    span_b200_ctx_t *ctx = span_b200_default_ctx();
    if (ctx == NULL)
        return NULL;
    int heap = 0;
    if (s != NULL  &&  s->magic == SB_MAGIC  &&  s->bank != NULL)
    {
        span_b200_v27ter_bank_destroy(s->bank);
        delete s->bits;
        delete s->syms;
        heap = s->heap;
    }
    else if (s == NULL)
    {
        if ((s = (v27ter_rx_state_t *) calloc(1, sizeof(*s))) == NULL)
            return NULL;
        heap = 1;
    }
    memset(s, 0, sizeof(*s));
    s->bank = span_b200_v27ter_bank_create(ctx, 1, bit_rate, 1);
    if (s->bank == NULL)
    {
        if (heap)
            free(s);
        return NULL;
    }
    s->magic = SB_MAGIC;
    s->heap = heap;
    s->bit_rate = bit_rate;
    s->put_bit = put_bit;
    s->put_bit_user_data = user_data;
    s->bits = new std::vector<int8_t>();
    s->syms = new std::vector<span_b200_v27ter_symbol_t>();
    return s;
}

extern "C" int v27ter_rx_restart(v27ter_rx_state_t *s, int bit_rate, bool old_train)
{
    if (bit_rate != 4800  &&  bit_rate != 2400)
        return -1;                                          // src/v27ter_rx.c:1095-1103
    s->bit_rate = bit_rate;
    return span_b200_v27ter_bank_restart(s->bank, 0, 1, bit_rate, (old_train)  ?  1  :  0);
}

static int v27ter_close(v27ter_rx_state_t *s, int do_free)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (s == NULL  ||  s->magic != SB_MAGIC)
        return 0;
    span_b200_v27ter_bank_destroy(s->bank);
    delete s->bits;
    delete s->syms;
    const int heap = s->heap;
    s->magic = 0;
    s->bank = NULL;
    if (do_free  &&  heap)
        free(s);
    return 0;
}

extern "C" int v27ter_rx_release(v27ter_rx_state_t *s) { return v27ter_close(s, 0); }
extern "C" int v27ter_rx_free(v27ter_rx_state_t *s) { return v27ter_close(s, 1); }
extern "C" logging_state_t *v27ter_rx_get_logging_state(v27ter_rx_state_t *s) { (void) s; return NULL; }

extern "C" void v27ter_rx_set_put_bit(v27ter_rx_state_t *s, span_put_bit_func_t put_bit, void *user_data)
{
    s->put_bit = put_bit;
    s->put_bit_user_data = user_data;
}

extern "C" void v27ter_rx_set_modem_status_handler(v27ter_rx_state_t *s, span_modem_status_func_t handler, void *user_data)
{
    s->status_handler = handler;
    s->status_user_data = user_data;
}

extern "C" void v27ter_rx_set_qam_report_handler(v27ter_rx_state_t *s, qam_report_handler_t handler, void *user_data)
{
    s->qam_report = handler;
    s->qam_user_data = user_data;
}

// One entry of the put_bit stream: data bits go to put_bit, status codes to the status handler if one
// is installed, else to put_bit (src/v27ter_rx.c:166-174).
static void v27ter_deliver(v27ter_rx_state_t *s, int v)
{
    if (v < 0  &&  s->status_handler)
        s->status_handler(s->status_user_data, v);
    else if (s->put_bit)
        s->put_bit(s->put_bit_user_data, v);
}

extern "C" int v27ter_rx(v27ter_rx_state_t *s, const int16_t amp[], int len)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (len <= 0)
        return 0;
    if (span_b200_v27ter_bank_rx_host(s->bank, amp, len, len, NULL) != 0)
        return 0;
    int32_t nb = 0;
    int32_t ns = 0;
    if (span_b200_v27ter_bank_counts(s->bank, &nb, &ns) != 0)
        return 0;
    s->bits->resize((size_t) std::max(nb, 1));
    s->syms->resize((size_t) std::max(ns, 1));
    if (nb > 0)
        span_b200_v27ter_bank_bits(s->bank, 0, s->bits->data(), nb);
    if (ns > 0)
        span_b200_v27ter_bank_symbols(s->bank, 0, s->syms->data(), ns);
    // Replay in the reference's order: the bits of a baud come before that baud's qam report.
    int done = 0;
    for (int k = 0;  k < ns;  k++)
    {
        const span_b200_v27ter_symbol_t &y = (*s->syms)[k];
        for (  ;  done < y.bit_pos  &&  done < nb;  done++)
            v27ter_deliver(s, (*s->bits)[done]);
        if (s->qam_report)
        {
            if (y.re != y.re)
            {
                // A Gardner timing hop: the reference passes NULL pointers and the integrator value (src/v27ter_rx.c:517-518)
                s->qam_report(s->qam_user_data, NULL, NULL, y.state);
            }
            else
            {
                const complexf_t z = {y.re, y.im};
                const complexf_t t = {y.target_re, y.target_im};
                s->qam_report(s->qam_user_data, &z, &t, y.state);
            }
        }
    }
    for (  ;  done < nb;  done++)
        v27ter_deliver(s, (*s->bits)[done]);
    return 0;                                               // src/v27ter_rx.c:1026
}

extern "C" int v27ter_rx_fillin(v27ter_rx_state_t *s, int len)
{
    return (span_b200_v27ter_bank_fillin(s->bank, 0, 1, len) == 0)  ?  0  :  0;
}

extern "C" int v27ter_rx_equalizer_state(v27ter_rx_state_t *s, complexf_t **coeffs)
{
    float eq[64];
    span_b200_v27ter_bank_channel_state(s->bank, 0, eq, NULL);
    for (int i = 0;  i < 32;  i++)
    {
        s->eq_coeff[i].re = eq[2*i];
        s->eq_coeff[i].im = eq[2*i + 1];
    }
    *coeffs = s->eq_coeff;
    return 32;                                              // V27TER_EQUALIZER_LEN
}

extern "C" float v27ter_rx_carrier_frequency(v27ter_rx_state_t *s)
{
    int32_t info[12];
    span_b200_v27ter_bank_channel_state(s->bank, 0, NULL, info);
    return (float) info[1]*(float) 8000/(65536.0f*65536.0f);           // dds_frequencyf, src/dds_float.c:2115-2118
}

extern "C" float v27ter_rx_symbol_timing_correction(v27ter_rx_state_t *s)
{
    int32_t info[12];
    span_b200_v27ter_bank_channel_state(s->bank, 0, NULL, info);
    const int steps_per_symbol = (info[11] == 4800)  ?  8*5  :  12*20/3;
    return (float) info[5]/(float) steps_per_symbol;                    // src/v27ter_rx.c:143-149
}

extern "C" float v27ter_rx_signal_power(v27ter_rx_state_t *s)
{
    int32_t info[12];
    span_b200_v27ter_bank_channel_state(s->bank, 0, NULL, info);
    // power_meter_current_dbm0() + 3.98f: src/power_meter.c:115-122, src/v27ter_rx.c:152-155
    if (info[10] <= 0)
        return (-96.329f + (3.14f + 3.02f)) + 3.98f;
    return (10.0f*log10f((float) info[10]/(32767.0f*32767.0f) + 1.0e-10f) + (3.14f + 3.02f)) + 3.98f;
}

extern "C" void v27ter_rx_set_signal_cutoff(v27ter_rx_state_t *s, float cutoff)
{
    span_b200_v27ter_bank_set_signal_cutoff(s->bank, 0, 1, cutoff);
}

// ------------------------------------------------------------------------------------------
// FSK receiver, one bank of one per state object (src/fsk.c:271-354,396-760)
struct fsk_rx_state_s
{
    unsigned long long magic;
    span_b200_fsk_bank_t *bank;
    int heap;
    span_put_bit_func_t put_bit;
    void *put_bit_user_data;
    span_modem_status_func_t status_handler;
    void *status_user_data;
    std::vector<int16_t> *out;
};

static_assert(sizeof(fsk_rx_state_s) <= 2200, "must fit the reference's fsk_rx_state_t (private/fsk.h)");
static_assert(sizeof(fsk_spec_t) == sizeof(span_b200_fsk_spec_t), "fsk_spec_t layout");

// preset_fsk_specs[] (src/fsk.c:60-156), exported as data like the reference's (src/spandsp/fsk.h:131)
extern "C" __attribute__((visibility("default"))) const fsk_spec_t preset_fsk_specs[] =
{
    {"V21 ch 1", 1080 + 100, 1080 - 100, -14, -30, 300*100},
    {"V21 ch 2", 1750 + 100, 1750 - 100, -14, -30, 300*100},
    {"V23 ch 1", 1700 + 400, 1700 - 400, -14, -30, 1200*100},
    {"V23 ch 2", 420 + 30, 420 - 30, -14, -30, 75*100},
    {"Bell103 ch 1", 1170 - 100, 1170 + 100, -14, -30, 300*100},
    {"Bell103 ch 2", 2125 - 100, 2125 + 100, -14, -30, 300*100},
    {"Bell202", 1700 + 500, 1700 - 500, -14, -30, 1200*100},
    {"Weitbrecht 45.45", 1600 + 200, 1600 - 200, -14, -30, 4545},
    {"Weitbrecht 50", 1600 + 200, 1600 - 200, -14, -30, 50*100},
    {"Weitbrecht 47.6", 1600 + 200, 1600 - 200, -14, -30, 4760},
    {"V21 (110bps) ch 1", 1080 + 100, 1080 - 100, -14, -30, 110*100}
};

extern "C" fsk_rx_state_t *fsk_rx_init(fsk_rx_state_t *s, const fsk_spec_t *spec, int framing_mode, span_put_bit_func_t put_bit, void *user_data)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (spec == NULL)
        return NULL;
    span_b200_ctx_t *ctx = span_b200_default_ctx();
    if (ctx == NULL)
        return NULL;
    int heap = 0;
    if (s != NULL  &&  s->magic == SB_MAGIC  &&  s->bank != NULL)
    {
        span_b200_fsk_bank_destroy(s->bank);
        delete s->out;
        heap = s->heap;
    }
    else if (s == NULL)
    {
        if ((s = (fsk_rx_state_t *) calloc(1, sizeof(*s))) == NULL)
            return NULL;
        heap = 1;
    }
    memset(s, 0, sizeof(*s));
    s->bank = span_b200_fsk_bank_create(ctx, 1, (const span_b200_fsk_spec_t *) spec, framing_mode);
    if (s->bank == NULL)
    {
        if (heap)
            free(s);
        return NULL;
    }
    s->magic = SB_MAGIC;
    s->heap = heap;
    s->put_bit = put_bit;
    s->put_bit_user_data = user_data;
    s->out = new std::vector<int16_t>();
    return s;
}

extern "C" int fsk_rx_restart(fsk_rx_state_t *s, const fsk_spec_t *spec, int framing_mode)
{
    span_b200_fsk_bank_restart(s->bank, 0, 1, (const span_b200_fsk_spec_t *) spec, framing_mode);
    return 0;                                               // src/fsk.c:721
}

static int fsk_close(fsk_rx_state_t *s, int do_free)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (s == NULL  ||  s->magic != SB_MAGIC)
        return 0;
    span_b200_fsk_bank_destroy(s->bank);
    delete s->out;
    const int heap = s->heap;
    s->magic = 0;
    s->bank = NULL;
    if (do_free  &&  heap)
        free(s);
    return 0;
}

extern "C" int fsk_rx_release(fsk_rx_state_t *s) { return fsk_close(s, 0); }
extern "C" int fsk_rx_free(fsk_rx_state_t *s) { return fsk_close(s, 1); }

extern "C" void fsk_rx_set_put_bit(fsk_rx_state_t *s, span_put_bit_func_t put_bit, void *user_data)
{
    s->put_bit = put_bit;
    s->put_bit_user_data = user_data;
}

extern "C" void fsk_rx_set_modem_status_handler(fsk_rx_state_t *s, span_modem_status_func_t handler, void *user_data)
{
    s->status_handler = handler;
    s->status_user_data = user_data;
}

extern "C" void fsk_rx_set_signal_cutoff(fsk_rx_state_t *s, float cutoff)
{
    span_b200_fsk_bank_set_signal_cutoff(s->bank, 0, 1, cutoff);
}

extern "C" float fsk_rx_signal_power(fsk_rx_state_t *s)
{
    return span_b200_fsk_bank_signal_power(s->bank, 0);
}

extern "C" void fsk_rx_set_frame_parameters(fsk_rx_state_t *s, int data_bits, int parity, int stop_bits)
{
    span_b200_fsk_bank_set_frame_parameters(s->bank, 0, 1, data_bits, parity, stop_bits);
}

extern "C" int fsk_rx_get_parity_errors(fsk_rx_state_t *s, bool reset)
{
    int32_t e = 0;
    span_b200_fsk_bank_errors(s->bank, 0, &e, NULL, (reset)  ?  1  :  0);
    return e;
}

extern "C" int fsk_rx_get_framing_errors(fsk_rx_state_t *s, bool reset)
{
    int32_t e = 0;
    span_b200_fsk_bank_errors(s->bank, 0, NULL, &e, (reset)  ?  1  :  0);
    return e;
}

extern "C" int fsk_rx(fsk_rx_state_t *s, const int16_t *amp, int len)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (len <= 0)
        return 0;
    if (span_b200_fsk_bank_rx_host(s->bank, amp, len, len, NULL) != 0)
        return 0;
    int32_t n = 0;
    if (span_b200_fsk_bank_counts(s->bank, &n) != 0)
        return 0;
    s->out->resize((size_t) std::max(n, 1));
    if (n > 0)
        n = (int32_t) span_b200_fsk_bank_output(s->bank, 0, s->out->data(), n);
    for (int i = 0;  i < n;  i++)
    {
        const int v = (*s->out)[i];
        // status reports go to the status handler if one is installed, else through put_bit (src/fsk.c:347-354)
        if (v < 0  &&  s->status_handler)
            s->status_handler(s->status_user_data, v);
        else if (s->put_bit)
            s->put_bit(s->put_bit_user_data, v);
    }
    return 0;                                               // src/fsk.c:625
}

extern "C" int fsk_rx_fillin(fsk_rx_state_t *s, int len)
{
    span_b200_fsk_bank_fillin(s->bank, 0, 1, len);
    return 0;
}

// ------------------------------------------------------------------------------------------
// Modem connect tone detector, one bank of one per state object (src/modem_connect_tones.c:74-118,419-892)
struct modem_connect_tones_rx_state_s
{
    unsigned long long magic;
    span_b200_mct_bank_t *bank;
    int heap;
    int hit;
    span_tone_report_func_t tone_callback;
    void *callback_data;
    std::vector<span_b200_mct_event_t> *ev;
};

static_assert(sizeof(modem_connect_tones_rx_state_s) <= 2300, "must fit the reference's modem_connect_tones_rx_state_t");

extern "C" const char *modem_connect_tone_to_str(int tone)
{
    switch (tone)                                           // src/modem_connect_tones.c:74-103
    {
    case MODEM_CONNECT_TONES_NONE:
        return "No tone";
    case MODEM_CONNECT_TONES_FAX_CNG:
        return "FAX CNG";
    case MODEM_CONNECT_TONES_ANS:
        return "ANS or FAX CED";
    case MODEM_CONNECT_TONES_ANS_PR:
        return "ANS/";
    case MODEM_CONNECT_TONES_ANSAM:
        return "ANSam";
    case MODEM_CONNECT_TONES_ANSAM_PR:
        return "ANSam/";
    case MODEM_CONNECT_TONES_FAX_PREAMBLE:
        return "FAX preamble";
    case MODEM_CONNECT_TONES_FAX_CED_OR_PREAMBLE:
        return "FAX CED or preamble";
    case MODEM_CONNECT_TONES_BELL_ANS:
        return "Bell ANS";
    case MODEM_CONNECT_TONES_CALLING_TONE:
        return "Calling tone";
    }
    return "???";
}

extern "C" modem_connect_tones_rx_state_t *modem_connect_tones_rx_init(modem_connect_tones_rx_state_t *s, int tone_type,
                                                                       span_tone_report_func_t tone_callback, void *user_data)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    span_b200_ctx_t *ctx = span_b200_default_ctx();
    if (ctx == NULL)
        return NULL;
    int heap = 0;
    if (s != NULL  &&  s->magic == SB_MAGIC  &&  s->bank != NULL)
    {
        // a second init of a live state keeps its bank (the reference re-initialises in place)
        if (span_b200_mct_bank_init(s->bank, 0, 1, tone_type) != 0)
            return NULL;
        s->hit = MODEM_CONNECT_TONES_NONE;
        s->tone_callback = tone_callback;
        s->callback_data = user_data;
        return s;
    }
    if (s == NULL)
    {
        if ((s = (modem_connect_tones_rx_state_t *) calloc(1, sizeof(*s))) == NULL)
            return NULL;
        heap = 1;
    }
    memset(s, 0, sizeof(*s));
    s->bank = span_b200_mct_bank_create(ctx, 1, tone_type);
    if (s->bank == NULL)
    {
        if (heap)
            free(s);
        return NULL;
    }
    s->magic = SB_MAGIC;
    s->heap = heap;
    s->hit = MODEM_CONNECT_TONES_NONE;
    s->tone_callback = tone_callback;
    s->callback_data = user_data;
    s->ev = new std::vector<span_b200_mct_event_t>();
    return s;
}

static int mct_close(modem_connect_tones_rx_state_t *s, int do_free)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (s == NULL  ||  s->magic != SB_MAGIC)
        return 0;
    span_b200_mct_bank_destroy(s->bank);
    delete s->ev;
    const int heap = s->heap;
    s->magic = 0;
    s->bank = NULL;
    if (do_free  &&  heap)
        free(s);
    return 0;
}

extern "C" int modem_connect_tones_rx_release(modem_connect_tones_rx_state_t *s) { return mct_close(s, 0); }
extern "C" int modem_connect_tones_rx_free(modem_connect_tones_rx_state_t *s) { return mct_close(s, 1); }

extern "C" int modem_connect_tones_rx(modem_connect_tones_rx_state_t *s, const int16_t amp[], int len)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (len <= 0)
        return 0;
    if (span_b200_mct_bank_rx_host(s->bank, amp, len, len, NULL) != 0)
        return 0;
    int64_t n = span_b200_mct_bank_events(s->bank, NULL, 0);
    if (n <= 0)
        return 0;
    s->ev->resize((size_t) n);
    n = span_b200_mct_bank_events(s->bank, s->ev->data(), n);
    for (int64_t i = 0;  i < n;  i++)
    {
        const span_b200_mct_event_t &e = (*s->ev)[(size_t) i];
        // report_tone_state() (src/modem_connect_tones.c:419-438): the callback, or else the hit
        if (s->tone_callback)
            s->tone_callback(s->callback_data, e.tone, e.level, 0);
        else if (e.tone != MODEM_CONNECT_TONES_NONE)
            s->hit = e.tone;
    }
    return 0;                                               // src/modem_connect_tones.c:803
}

extern "C" int modem_connect_tones_rx_fillin(modem_connect_tones_rx_state_t *s, int len)
{
    return 0;                                               // src/modem_connect_tones.c:806-809: nothing is done
}

extern "C" int modem_connect_tones_rx_get(modem_connect_tones_rx_state_t *s)
{
    const int x = s->hit;                                   // src/modem_connect_tones.c:812-820
    s->hit = MODEM_CONNECT_TONES_NONE;
    return x;
}

// ------------------------------------------------------------------------------------------
// In-band signalling tone receiver, one bank of one per state object (src/sig_tone.c:402-734)
struct sig_tone_rx_state_s
{
    unsigned long long magic;
    span_b200_sig_bank_t *bank;
    int heap;
    span_tone_report_func_t sig_update;
    void *user_data;
    std::vector<span_b200_sig_event_t> *ev;
};

static_assert(sizeof(sig_tone_rx_state_s) <= 160, "must fit the reference's sig_tone_rx_state_t (private/sig_tone.h)");

extern "C" sig_tone_rx_state_t *sig_tone_rx_init(sig_tone_rx_state_t *s, int tone_type, span_tone_report_func_t sig_update, void *user_data)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (sig_update == NULL  ||  tone_type < 1  ||  tone_type > 3)
        return NULL;                                        // src/sig_tone.c:679-680
    span_b200_ctx_t *ctx = span_b200_default_ctx();
    if (ctx == NULL)
        return NULL;
    int heap = 0;
    if (s != NULL  &&  s->magic == SB_MAGIC  &&  s->bank != NULL)
    {
        if (span_b200_sig_bank_init(s->bank, 0, 1, tone_type) != 0)
            return NULL;
        s->sig_update = sig_update;
        s->user_data = user_data;
        return s;
    }
    if (s == NULL)
    {
        if ((s = (sig_tone_rx_state_t *) calloc(1, sizeof(*s))) == NULL)
            return NULL;
        heap = 1;
    }
    memset(s, 0, sizeof(*s));
    s->bank = span_b200_sig_bank_create(ctx, 1, tone_type);
    if (s->bank == NULL)
    {
        if (heap)
            free(s);
        return NULL;
    }
    s->magic = SB_MAGIC;
    s->heap = heap;
    s->sig_update = sig_update;
    s->user_data = user_data;
    s->ev = new std::vector<span_b200_sig_event_t>();
    return s;
}

static int sig_close(sig_tone_rx_state_t *s, int do_free)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (s == NULL  ||  s->magic != SB_MAGIC)
        return 0;
    span_b200_sig_bank_destroy(s->bank);
    delete s->ev;
    const int heap = s->heap;
    s->magic = 0;
    s->bank = NULL;
    if (do_free  &&  heap)
        free(s);
    return 0;
}

extern "C" int sig_tone_rx_release(sig_tone_rx_state_t *s) { return sig_close(s, 0); }
extern "C" int sig_tone_rx_free(sig_tone_rx_state_t *s) { return sig_close(s, 1); }

extern "C" void sig_tone_rx_set_mode(sig_tone_rx_state_t *s, int mode, int duration)
{
    span_b200_sig_bank_set_mode(s->bank, 0, 1, mode);       // the duration is not used on the receive side (src/sig_tone.c:666-669)
}

extern "C" int sig_tone_rx(sig_tone_rx_state_t *s, int16_t amp[], int len)
{
    std::lock_guard<std::recursive_mutex> lk(g_lock);
    if (len <= 0)
        return len;
    if (span_b200_sig_bank_rx_host(s->bank, amp, len, len, NULL) != 0)
        return len;
    int64_t n = span_b200_sig_bank_events(s->bank, NULL, 0);
    if (n > 0)
    {
        s->ev->resize((size_t) n);
        n = span_b200_sig_bank_events(s->bank, s->ev->data(), n);
        for (int64_t i = 0;  i < n;  i++)
        {
            const span_b200_sig_event_t &e = (*s->ev)[(size_t) i];
            if (s->sig_update)
                s->sig_update(s->user_data, e.signalling_state, 0, e.duration);
        }
    }
    return len;                                             // src/sig_tone.c:663
}

// ------------------------------------------------------------------------------------------
// FAX receive front end (src/fax_modems.c:177-333): host glue over the drop-in receivers above
struct span_b200_fax_rx_s
{
    int fast_modem;
    v29_rx_state_t *v29;
    v17_rx_state_t *v17;
    v27ter_rx_state_t *v27ter;
    fsk_rx_state_t *v21;
    span_put_bit_func_t fast_put_bit;
    void *fast_user_data;
    int current;
    int rx_frame_received;
};

// v29_rx_status_handler() / v17_rx_status_handler() / v27ter_rx_status_handler() (src/fax_modems.c:197-209,244-256,291-303)
static void fax_fast_status(void *user_data, int status)
{
    span_b200_fax_rx_t *s = (span_b200_fax_rx_t *) user_data;
    if (status == SIG_STATUS_TRAINING_SUCCEEDED)
    {
        // "Switching from V.xx + V.21 to V.xx": the fast modem keeps the line, and its status goes through put_bit again
        s->current = SPAN_B200_FAX_RX_FAST;
        switch (s->fast_modem)
        {
        case 29:
            v29_rx_set_modem_status_handler(s->v29, NULL, s);
            break;
        case 17:
            v17_rx_set_modem_status_handler(s->v17, NULL, s);
            break;
        default:
            v27ter_rx_set_modem_status_handler(s->v27ter, NULL, s);
            break;
        }
    }
    if (s->fast_put_bit)
        s->fast_put_bit(s->fast_user_data, status);
}

extern "C" int span_b200_fax_rx_free(span_b200_fax_rx_t *s)
{
    if (s == NULL)
        return 0;
    if (s->v29)
        v29_rx_free(s->v29);
    if (s->v17)
        v17_rx_free(s->v17);
    if (s->v27ter)
        v27ter_rx_free(s->v27ter);
    if (s->v21)
        fsk_rx_free(s->v21);
    delete s;
    return 0;
}

extern "C" span_b200_fax_rx_t *span_b200_fax_rx_init(int fast_modem, int bit_rate, int short_train,
                                                     span_put_bit_func_t fast_put_bit, void *fast_user_data,
                                                     span_put_bit_func_t v21_put_bit, void *v21_user_data)
{
    span_b200_fax_rx_t *s = new span_b200_fax_rx_s();
    memset(s, 0, sizeof(*s));
    s->fast_modem = fast_modem;
    s->fast_put_bit = fast_put_bit;
    s->fast_user_data = fast_user_data;
    s->current = SPAN_B200_FAX_RX_BOTH;
    bool ok = false;
    switch (fast_modem)
    {
    case 29:
        if ((s->v29 = v29_rx_init(NULL, bit_rate, fast_put_bit, fast_user_data)) != NULL)
        {
            v29_rx_set_modem_status_handler(s->v29, fax_fast_status, s);
            ok = true;
        }
        break;
    case 17:
        if ((s->v17 = v17_rx_init(NULL, bit_rate, fast_put_bit, fast_user_data)) != NULL)
        {
            if (short_train)
                v17_rx_restart(s->v17, bit_rate, short_train);
            v17_rx_set_modem_status_handler(s->v17, fax_fast_status, s);
            ok = true;
        }
        break;
    case 27:
        if ((s->v27ter = v27ter_rx_init(NULL, bit_rate, fast_put_bit, fast_user_data)) != NULL)
        {
            v27ter_rx_set_modem_status_handler(s->v27ter, fax_fast_status, s);
            ok = true;
        }
        break;
    default:
        sb_set_error("fast modem %d (17, 27 or 29)", fast_modem);
        break;
    }
    if (ok)
    {
        // fax_modems_start_slow_modem(FAX_MODEM_V21_RX), src/fax_modems.c:340-343
        s->v21 = fsk_rx_init(NULL, &preset_fsk_specs[FSK_V21CH2], FSK_FRAME_MODE_SYNC, v21_put_bit, v21_user_data);
        if (s->v21)
            fsk_rx_set_signal_cutoff(s->v21, -39.09f);
        else
            ok = false;
    }
    if (!ok)
    {
        span_b200_fax_rx_free(s);
        return NULL;
    }
    return s;
}

extern "C" void span_b200_fax_rx_frame_received(span_b200_fax_rx_t *s)
{
    s->rx_frame_received = 1;
}

extern "C" int span_b200_fax_rx_current(const span_b200_fax_rx_t *s)
{
    return s->current;
}

static void fax_fast_rx(span_b200_fax_rx_t *s, const int16_t amp[], int len)
{
    switch (s->fast_modem)
    {
    case 29:
        v29_rx(s->v29, amp, len);
        break;
    case 17:
        v17_rx(s->v17, amp, len);
        break;
    default:
        v27ter_rx(s->v27ter, amp, len);
        break;
    }
}

// fax_modems_v29_v21_rx() and its two siblings (src/fax_modems.c:213-228,260-275,307-322) - and, once the handler has
// been swapped, the plain receiver the reference's rx handler then points at
extern "C" int span_b200_fax_rx(span_b200_fax_rx_t *s, const int16_t amp[], int len)
{
    switch (s->current)
    {
    case SPAN_B200_FAX_RX_FAST:
        fax_fast_rx(s, amp, len);
        return 0;
    case SPAN_B200_FAX_RX_V21:
        fsk_rx(s->v21, amp, len);
        return 0;
    }
    // The status handler may swap to the fast modem in the middle of this call; the reference still runs fsk_rx() on
    // this block (the swap only changes who gets the NEXT block)
    fax_fast_rx(s, amp, len);
    fsk_rx(s->v21, amp, len);
    if (s->rx_frame_received)
    {
        // "We have received something, and the fast modem has not trained. We must be receiving valid V.21" - checked
        // after both receivers have run, so V.21 also wins a tie inside one block (:313-319)
        s->current = SPAN_B200_FAX_RX_V21;
    }
    return 0;
}

extern "C" int span_b200_fax_rx_fillin(span_b200_fax_rx_t *s, int len)
{
    if (s->current != SPAN_B200_FAX_RX_V21)
    {
        switch (s->fast_modem)
        {
        case 29:
            v29_rx_fillin(s->v29, len);
            break;
        case 17:
            v17_rx_fillin(s->v17, len);
            break;
        default:
            v27ter_rx_fillin(s->v27ter, len);
            break;
        }
    }
    if (s->current != SPAN_B200_FAX_RX_FAST)
        fsk_rx_fillin(s->v21, len);
    return 0;
}
