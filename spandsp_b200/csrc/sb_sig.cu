// sb_sig.cu - C ABI of the signalling tone receiver banks (include/spandsp_b200_sig.h).  The receiver is
// sb_sig_rx.cuh.  Reference: src/sig_tone.c.
#include <vector>

#include "sb_engine.h"
#include "sb_sig_rx.cuh"

#pragma GCC visibility push(default)
#include "../../include/spandsp_b200_sig.h"
#pragma GCC visibility pop

using namespace sbs;

#define CK(call) \
    do \
    { \
        cudaError_t e_ = (call); \
        if (e_ != cudaSuccess) \
        { \
            sb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return -1; \
        } \
    } \
    while (0)

struct span_b200_sig_bank_s
{
    span_b200_ctx_t *ctx;
    int channels;
    int *state;
    int2 *ev;
    long long ev_cap;
    int *nev;
    int16_t *d_io;
    size_t d_io_bytes;
    cudaStream_t last_stream;
    bool have_last;
    std::vector<int> *h_nev;
    std::vector<int2> *h_ev;
};

static SigArgs sig_args(span_b200_sig_bank_t *b, int16_t *d_amp, int64_t stride, int n)
{
    SigArgs a;
    a.amp = d_amp;
    a.stride = stride;
    a.n = n;
    a.channels = b->channels;
    a.state = b->state;
    a.ev = b->ev;
    a.ev_cap = b->ev_cap;
    a.nev = b->nev;
    return a;
}

static int sig_quiesce(span_b200_sig_bank_t *b)
{
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    return 0;
}

static int sig_ctl(span_b200_sig_bank_t *b, int first, int count, int mode, int ia, int ib, int ic, int id)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b == NULL  ||  first < 0  ||  count < 0  ||  first + count > b->channels)
    {
        sb_set_error("channel range out of bounds");
        return -1;
    }
    if (count == 0)
        return 0;
    if (sig_quiesce(b) != 0)
        return -1;
    cudaStream_t st = (cudaStream_t) sb_ctx_stream(b->ctx);
    SigArgs a = sig_args(b, NULL, 0, 0);
    sig_ctl_kernel<<<(count + 127)/128, 128, 0, st>>>(a, first, count, mode, ia, ib, ic, id);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int span_b200_sig_bank_init(span_b200_sig_bank_t *b, int first, int count, int tone_type)
{
    if (tone_type < 1  ||  tone_type > 3)
    {
        sb_set_error("bad signalling tone type %d", tone_type);     // src/sig_tone.c:679-680
        return -1;
    }
    int32_t thr[3];
    host_sig_thresholds(tone_type, thr);
    return sig_ctl(b, first, count, 0, tone_type, thr[0], thr[1], thr[2]);
}

extern "C" int span_b200_sig_bank_set_mode(span_b200_sig_bank_t *b, int first, int count, int mode)
{
    return sig_ctl(b, first, count, 1, mode, 0, 0, 0);
}

extern "C" void span_b200_sig_bank_destroy(span_b200_sig_bank_t *b)
{
    if (b == NULL)
        return;
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b->have_last)
        cudaStreamSynchronize(b->last_stream);
    cudaFree(b->state);
    cudaFree(b->ev);
    cudaFree(b->nev);
    cudaFree(b->d_io);
    delete b->h_nev;
    delete b->h_ev;
    delete b;
}

extern "C" span_b200_sig_bank_t *span_b200_sig_bank_create(span_b200_ctx_t *ctx, int channels, int tone_type)
{
    if (ctx == NULL  ||  channels <= 0  ||  tone_type < 1  ||  tone_type > 3)
    {
        sb_set_error("bad signalling tone bank arguments");
        return NULL;
    }
    SB_DEVICE_CKP(span_b200_ctx_device(ctx));
    span_b200_sig_bank_t *b = new span_b200_sig_bank_s();
    memset(b, 0, sizeof(*b));
    b->ctx = ctx;
    b->channels = channels;
    b->h_nev = new std::vector<int>();
    b->h_ev = new std::vector<int2>();
    const size_t C = channels;
    bool ok = cudaMalloc(&b->state, sizeof(int)*T_COUNT*C) == cudaSuccess
              &&  cudaMalloc(&b->nev, sizeof(int)*C) == cudaSuccess
              &&  cudaMemset(b->nev, 0, sizeof(int)*C) == cudaSuccess;
    if (!ok)
    {
        sb_set_error("signalling tone bank allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        span_b200_sig_bank_destroy(b);
        return NULL;
    }
    if (span_b200_sig_bank_init(b, 0, channels, tone_type) != 0)
    {
        span_b200_sig_bank_destroy(b);
        return NULL;
    }
    return b;
}

extern "C" int span_b200_sig_bank_channels(const span_b200_sig_bank_t *b)
{
    return b->channels;
}

static int sig_realloc(void **p, size_t bytes)
{
    if (*p)
        CK(cudaFree(*p));
    *p = NULL;
    CK(cudaMalloc(p, bytes));
    return 0;
}

extern "C" int span_b200_sig_bank_rx_device(span_b200_sig_bank_t *b, int16_t *d_amp, int64_t stride, int n, void *stream)
{
    if (b == NULL  ||  n < 0  ||  (n > 0  &&  d_amp == NULL))
    {
        sb_set_error("bad rx arguments");
        return -1;
    }
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    cudaStream_t st = (stream)  ?  (cudaStream_t) stream  :  (cudaStream_t) sb_ctx_stream(b->ctx);
    if (b->have_last  &&  b->last_stream != st)
        CK(cudaStreamSynchronize(b->last_stream));
    // A tone is confirmed after 24 consistent samples and dropped after 64: two reports per 88 samples at most in the
    // sharp detector; the flat detector can report a loss at once but hands back to the sharp one when it does
    const long long want = (long long) n/40 + 8;
    if (b->ev_cap < want)
    {
        if (b->have_last)
            CK(cudaStreamSynchronize(b->last_stream));
        if (sig_realloc((void **) &b->ev, sizeof(int2)*(size_t) want*b->channels) != 0)
            return -1;
        b->ev_cap = want;
    }
    SigArgs a = sig_args(b, d_amp, stride, n);
    sig_rx_kernel<<<(b->channels + 63)/64, 64, 0, st>>>(a);
    CK(cudaGetLastError());
    b->last_stream = st;
    b->have_last = true;
    return 0;
}

extern "C" int span_b200_sig_bank_rx_host(span_b200_sig_bank_t *b, int16_t *h_amp, int64_t stride, int n, void *stream)
{
    if (b == NULL  ||  n < 0  ||  (n > 0  &&  h_amp == NULL))
    {
        sb_set_error("bad rx arguments");
        return -1;
    }
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    cudaStream_t st = (stream)  ?  (cudaStream_t) stream  :  (cudaStream_t) sb_ctx_stream(b->ctx);
    if (b->have_last  &&  b->last_stream != st)
        CK(cudaStreamSynchronize(b->last_stream));
    const size_t row = ((size_t) n + 7) & ~(size_t) 7;
    const size_t want = sizeof(int16_t)*row*b->channels + 16;
    if (b->d_io_bytes < want)
    {
        if (b->have_last)
            CK(cudaStreamSynchronize(b->last_stream));
        if (sig_realloc((void **) &b->d_io, want) != 0)
            return -1;
        b->d_io_bytes = want;
    }
    if (n > 0)
        CK(cudaMemcpy2DAsync(b->d_io, sizeof(int16_t)*row, h_amp, sizeof(int16_t)*stride, sizeof(int16_t)*(size_t) n,
                             b->channels, cudaMemcpyHostToDevice, st));
    if (span_b200_sig_bank_rx_device(b, b->d_io, (int64_t) row, n, (void *) st) != 0)
        return -1;
    if (n > 0)
        CK(cudaMemcpy2DAsync(h_amp, sizeof(int16_t)*stride, b->d_io, sizeof(int16_t)*row, sizeof(int16_t)*(size_t) n,
                             b->channels, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int64_t span_b200_sig_bank_events(span_b200_sig_bank_t *b, span_b200_sig_event_t *events, int64_t max)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b == NULL)
        return -1;
    if (sig_quiesce(b) != 0)
        return -1;
    if (!b->have_last  ||  b->ev_cap == 0)
        return 0;
    const size_t C = b->channels;
    b->h_nev->resize(C);
    CK(cudaMemcpy(b->h_nev->data(), b->nev, sizeof(int)*C, cudaMemcpyDeviceToHost));
    int most = 0;
    for (size_t c = 0;  c < C;  c++)
    {
        if ((*b->h_nev)[c] > most)
            most = (*b->h_nev)[c];
    }
    if (most == 0)
        return 0;
    if (most > b->ev_cap)
    {
        sb_set_error("signalling tone report buffer overflow (%d > %lld)", most, b->ev_cap);
        return -1;
    }
    b->h_ev->resize(C*(size_t) most);
    CK(cudaMemcpy2D(b->h_ev->data(), sizeof(int2)*most, b->ev, sizeof(int2)*(size_t) b->ev_cap, sizeof(int2)*most, C, cudaMemcpyDeviceToHost));
    int64_t total = 0;
    for (size_t c = 0;  c < C;  c++)
    {
        for (int i = 0;  i < (*b->h_nev)[c];  i++)
        {
            if (total < max  &&  events)
            {
                const int2 e = (*b->h_ev)[c*(size_t) most + i];
                events[total].channel = (int32_t) c;
                events[total].signalling_state = e.x;
                events[total].duration = e.y;
            }
            total++;
        }
    }
    return total;
}

extern "C" int span_b200_sig_bank_channel_state(span_b200_sig_bank_t *b, int channel, int32_t *info)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b == NULL  ||  channel < 0  ||  channel >= b->channels  ||  info == NULL)
        return -1;
    if (sig_quiesce(b) != 0)
        return -1;
    const size_t C = b->channels;
    CK(cudaMemcpy2D(info, sizeof(int), b->state + channel, sizeof(int)*C, sizeof(int), T_COUNT, cudaMemcpyDeviceToHost));
    return 0;
}
