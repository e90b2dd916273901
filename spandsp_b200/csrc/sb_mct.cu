// sb_mct.cu - C ABI of the modem connect tone detector banks (include/spandsp_b200_mct.h).  The detector is
// sb_mct_rx.cuh.  Reference: src/modem_connect_tones.c.
#include <vector>

#include "sb_engine.h"
#define SBF_NO_FSK_KERNELS
#include "sb_mct_rx.cuh"

#pragma GCC visibility push(default)
#include "../../include/spandsp_b200_mct.h"
#pragma GCC visibility pop

using namespace sbf;

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

struct span_b200_mct_bank_s
{
    span_b200_ctx_t *ctx;
    int channels;
    int *state;                 // [M_COUNT][channels]
    int2 *window;               // the V.21 receivers' correlation windows
    short *sine;
    int2 *ev;                   // [channel][ev_cap]
    long long ev_cap;
    int *nev;
    int16_t *d_in;
    size_t d_in_bytes;
    cudaStream_t last_stream;
    bool have_last;
    bool configured;
    std::vector<int> *h_nev;
    std::vector<int2> *h_ev;
};

// V.21 channel 2 at 300 baud: 26 samples of correlation window (src/fsk.c:694-696)
static const int MCT_WSPAN = 8000*100/(300*100);

static MctArgs mct_args(span_b200_mct_bank_t *b, const int16_t *d_amp, int64_t stride, int n)
{
    MctArgs a;
    a.f.amp = d_amp;
    a.f.stride = stride;
    a.f.n = n;
    a.f.channels = b->channels;
    a.f.state = b->state;
    a.f.window = b->window;
    a.f.sine = b->sine;
    a.f.out = NULL;
    a.f.out_cap = 0;
    a.f.nout = NULL;
    a.f.wspan = MCT_WSPAN;
    a.ev = b->ev;
    a.ev_cap = b->ev_cap;
    a.nev = b->nev;
    return a;
}

static int mct_quiesce(span_b200_mct_bank_t *b)
{
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    return 0;
}

static int mct_range_ok(span_b200_mct_bank_t *b, int first, int count)
{
    if (b == NULL  ||  first < 0  ||  count < 0  ||  first + count > b->channels)
    {
        sb_set_error("channel range out of bounds");
        return 0;
    }
    return 1;
}

// What fsk_rx_init(V.21 ch 2, synchronous) + fsk_rx_set_signal_cutoff(-45.5) derive on the host
// (src/modem_connect_tones.c:845-846, src/fsk.c:271-277,676-690)
static FskSetup mct_v21_setup(void)
{
    FskSetup su;
    su.baud_rate = 300*100;
    su.framing_mode = FRAME_MODE_SYNC;
    su.rate0 = host_dds_int_phase_rate((float) (1750 + 100));
    su.rate1 = host_dds_int_phase_rate((float) (1750 - 100));
    su.on_power = host_level_dbm0(-45.5f + 2.5f - 5.3f);
    su.off_power = host_level_dbm0(-45.5f - 2.5f - 5.3f);
    return su;
}

extern "C" int span_b200_mct_bank_init(span_b200_mct_bank_t *b, int first, int count, int tone_type)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (!mct_range_ok(b, first, count))
        return -1;
    if (count == 0)
        return 0;
    if (mct_quiesce(b) != 0)
        return -1;
    cudaStream_t st = (cudaStream_t) sb_ctx_stream(b->ctx);
    MctArgs a = mct_args(b, NULL, 0, 0);
    mct_init_kernel<<<(count + 127)/128, 128, 0, st>>>(a, first, count, tone_type, mct_v21_setup());
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" void span_b200_mct_bank_destroy(span_b200_mct_bank_t *b)
{
    if (b == NULL)
        return;
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b->have_last)
        cudaStreamSynchronize(b->last_stream);
    cudaFree(b->state);
    cudaFree(b->window);
    cudaFree(b->sine);
    cudaFree(b->ev);
    cudaFree(b->nev);
    cudaFree(b->d_in);
    delete b->h_nev;
    delete b->h_ev;
    delete b;
}

extern "C" span_b200_mct_bank_t *span_b200_mct_bank_create(span_b200_ctx_t *ctx, int channels, int tone_type)
{
    if (ctx == NULL  ||  channels <= 0)
    {
        sb_set_error("bad modem connect tone bank arguments");
        return NULL;
    }
    SB_DEVICE_CKP(span_b200_ctx_device(ctx));
    span_b200_mct_bank_t *b = new span_b200_mct_bank_s();
    memset(b, 0, sizeof(*b));
    b->ctx = ctx;
    b->channels = channels;
    b->h_nev = new std::vector<int>();
    b->h_ev = new std::vector<int2>();
    std::vector<short> t;
    make_dds_int_table(t);
    const size_t C = channels;
    bool ok = cudaMalloc(&b->state, sizeof(int)*M_COUNT*C) == cudaSuccess
              &&  cudaMalloc(&b->window, sizeof(int2)*2*SBF_MAX_WINDOW*C) == cudaSuccess
              &&  cudaMalloc(&b->sine, sizeof(short)*SBF_SINE_PAD) == cudaSuccess
              &&  cudaMalloc(&b->nev, sizeof(int)*C) == cudaSuccess
              &&  cudaMemcpy(b->sine, t.data(), sizeof(short)*SBF_SINE_PAD, cudaMemcpyHostToDevice) == cudaSuccess
              &&  cudaMemset(b->nev, 0, sizeof(int)*C) == cudaSuccess;
    if (!ok)
    {
        sb_set_error("modem connect tone bank allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        span_b200_mct_bank_destroy(b);
        return NULL;
    }
    if (span_b200_mct_bank_init(b, 0, channels, tone_type) != 0)
    {
        span_b200_mct_bank_destroy(b);
        return NULL;
    }
    return b;
}

extern "C" int span_b200_mct_bank_channels(const span_b200_mct_bank_t *b)
{
    return b->channels;
}

static int mct_realloc(void **p, size_t bytes)
{
    if (*p)
        CK(cudaFree(*p));
    *p = NULL;
    CK(cudaMalloc(p, bytes));
    return 0;
}

extern "C" int span_b200_mct_bank_rx_device(span_b200_mct_bank_t *b, const int16_t *d_amp, int64_t stride, int n, void *stream)
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
    // A tone is declared after >= 415 ms of it (3320 samples) or, for the preamble, 5 flags (1066 samples), and may end
    // one sample later: two reports per ~1000 samples per detector pass is the worst case, and there are two passes
    const long long want = 4*((long long) n/1000 + 2);
    if (b->ev_cap < want)
    {
        if (b->have_last)
            CK(cudaStreamSynchronize(b->last_stream));
        if (mct_realloc((void **) &b->ev, sizeof(int2)*(size_t) want*b->channels) != 0)
            return -1;
        b->ev_cap = want;
    }
    if (!b->configured)
    {
        CK(cudaFuncSetAttribute(mct_rx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fsk_smem_bytes(MCT_WSPAN)));
        b->configured = true;
    }
    MctArgs a = mct_args(b, d_amp, stride, n);
    mct_rx_kernel<<<(b->channels + 31)/32, 32, fsk_smem_bytes(MCT_WSPAN), st>>>(a);
    CK(cudaGetLastError());
    b->last_stream = st;
    b->have_last = true;
    return 0;
}

extern "C" int span_b200_mct_bank_rx_host(span_b200_mct_bank_t *b, const int16_t *h_amp, int64_t stride, int n, void *stream)
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
    // rows padded to a multiple of 8 samples so that every row starts 16-byte aligned
    const size_t row = ((size_t) n + 7) & ~(size_t) 7;
    const size_t want = sizeof(int16_t)*row*b->channels + 16;
    if (b->d_in_bytes < want)
    {
        if (b->have_last)
            CK(cudaStreamSynchronize(b->last_stream));
        if (mct_realloc((void **) &b->d_in, want) != 0)
            return -1;
        b->d_in_bytes = want;
    }
    if (n > 0)
        CK(cudaMemcpy2DAsync(b->d_in, sizeof(int16_t)*row, h_amp, sizeof(int16_t)*stride, sizeof(int16_t)*(size_t) n,
                             b->channels, cudaMemcpyHostToDevice, st));
    return span_b200_mct_bank_rx_device(b, b->d_in, (int64_t) row, n, (void *) st);
}

extern "C" int64_t span_b200_mct_bank_events(span_b200_mct_bank_t *b, span_b200_mct_event_t *events, int64_t max)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b == NULL)
        return -1;
    if (mct_quiesce(b) != 0)
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
        sb_set_error("modem connect tone report buffer overflow (%d > %lld)", most, b->ev_cap);
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
                events[total].tone = e.x & 0xFFFF;
                events[total].level = host_mct_level(e.x >> 16, e.y);
            }
            total++;
        }
    }
    return total;
}

extern "C" int span_b200_mct_bank_get(span_b200_mct_bank_t *b, int first, int count, int32_t *hits)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (!mct_range_ok(b, first, count))
        return -1;
    if (count == 0)
        return 0;
    if (mct_quiesce(b) != 0)
        return -1;
    int *p = b->state + (size_t) M_HIT*b->channels + first;
    if (hits)
        CK(cudaMemcpy(hits, p, sizeof(int)*(size_t) count, cudaMemcpyDeviceToHost));
    CK(cudaMemset(p, 0, sizeof(int)*(size_t) count));
    return 0;
}

extern "C" int span_b200_mct_bank_channel_state(span_b200_mct_bank_t *b, int channel, int32_t *info, int32_t *fsk_info)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b == NULL  ||  channel < 0  ||  channel >= b->channels)
        return -1;
    if (mct_quiesce(b) != 0)
        return -1;
    const size_t C = b->channels;
    if (info)
        CK(cudaMemcpy2D(info, sizeof(int), b->state + (size_t) K_COUNT*C + channel, sizeof(int)*C, sizeof(int), M_COUNT - K_COUNT, cudaMemcpyDeviceToHost));
    if (fsk_info)
        CK(cudaMemcpy2D(fsk_info, sizeof(int), b->state + channel, sizeof(int)*C, sizeof(int), K_COUNT, cudaMemcpyDeviceToHost));
    return 0;
}
