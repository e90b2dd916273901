// sb_fsk.cu - C ABI of the FSK receiver banks (include/spandsp_b200_fsk.h).  The receiver is sb_fsk_rx.cuh.
// Reference: src/fsk.c.
#include <vector>

#include "sb_engine.h"
#include "sb_fsk_rx.cuh"

#pragma GCC visibility push(default)
#include "../../include/spandsp_b200_fsk.h"
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

struct span_b200_fsk_bank_s
{
    span_b200_ctx_t *ctx;
    int channels;
    int *state;
    int2 *window;
    short *sine;
    short *out;
    long long out_cap;
    int *nout;
    int16_t *d_in;
    size_t d_in_bytes;
    int wspan;                  // largest correlation span any channel has been configured with
    int max_baud;               // largest baud rate x 100 any channel has been configured with
    cudaStream_t last_stream;
    bool have_last;
    bool configured;
};

// preset_fsk_specs[] (src/fsk.c:60-156)
static const span_b200_fsk_spec_t fsk_presets[SPAN_B200_FSK_PRESETS] =
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

extern "C" const span_b200_fsk_spec_t *span_b200_fsk_preset(int which)
{
    if (which < 0  ||  which >= SPAN_B200_FSK_PRESETS)
        return NULL;
    return &fsk_presets[which];
}

extern "C" int span_b200_dds_int_table(int16_t *table)
{
    std::vector<short> t;
    make_dds_int_table(t);
    memcpy(table, t.data(), sizeof(int16_t)*SBF_SINE_WORDS);
    return 0;
}

static FskArgs fsk_args(span_b200_fsk_bank_t *b, const int16_t *d_amp, int64_t stride, int n)
{
    FskArgs a;
    a.amp = d_amp;
    a.stride = stride;
    a.n = n;
    a.channels = b->channels;
    a.state = b->state;
    a.window = b->window;
    a.sine = b->sine;
    a.out = b->out;
    a.out_cap = b->out_cap;
    a.nout = b->nout;
    a.wspan = b->wspan;
    return a;
}

static int fsk_quiesce(span_b200_fsk_bank_t *b)
{
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    return 0;
}

static int fsk_range_ok(span_b200_fsk_bank_t *b, int first, int count)
{
    if (b == NULL  ||  first < 0  ||  count < 0  ||  first + count > b->channels)
    {
        sb_set_error("channel range out of bounds");
        return 0;
    }
    return 1;
}

static int fsk_configure(span_b200_fsk_bank_t *b)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (!b->configured)
    {
        CK(cudaFuncSetAttribute(fsk_rx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fsk_smem_bytes(SBF_MAX_WINDOW)));
        CK(cudaFuncSetAttribute(fsk_ctl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fsk_smem_bytes(SBF_MAX_WINDOW)));
        b->configured = true;
    }
    return 0;
}

static int fsk_ctl(span_b200_fsk_bank_t *b, int first, int count, int mode, const FskSetup &su, int aux)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (count <= 0)
        return 0;
    if (fsk_quiesce(b) != 0  ||  fsk_configure(b) != 0)
        return -1;
    cudaStream_t st = (cudaStream_t) sb_ctx_stream(b->ctx);
    FskArgs a = fsk_args(b, NULL, 0, 0);
    fsk_ctl_kernel<<<(count + 31)/32, 32, fsk_smem_bytes(b->wspan), st>>>(a, first, count, mode, su, aux);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

static bool fsk_spec_ok(const span_b200_fsk_spec_t *spec, int framing_mode)
{
    return spec != NULL  &&  spec->baud_rate > 0  &&  framing_mode >= 0  &&  framing_mode <= 2;
}

// What fsk_rx_restart() derives from the spec on the host (src/fsk.c:676-690,271-277)
static FskSetup fsk_setup(span_b200_fsk_bank_t *b, const span_b200_fsk_spec_t *spec, int framing_mode)
{
    FskSetup su;
    su.baud_rate = spec->baud_rate;
    su.framing_mode = framing_mode;
    su.rate0 = host_dds_int_phase_rate((float) spec->freq_zero);
    su.rate1 = host_dds_int_phase_rate((float) spec->freq_one);
    const float cutoff = (float) spec->min_level;
    su.on_power = host_level_dbm0(cutoff + 2.5f - 5.3f);
    su.off_power = host_level_dbm0(cutoff - 2.5f - 5.3f);
    int span = 8000*100/spec->baud_rate;
    if (span > SBF_MAX_WINDOW)
        span = SBF_MAX_WINDOW;
    if (span > b->wspan)
        b->wspan = span;
    if (spec->baud_rate > b->max_baud)
        b->max_baud = spec->baud_rate;
    return su;
}

extern "C" void span_b200_fsk_bank_destroy(span_b200_fsk_bank_t *b)
{
    if (b == NULL)
        return;
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b->have_last)
        cudaStreamSynchronize(b->last_stream);
    cudaFree(b->state);
    cudaFree(b->window);
    cudaFree(b->sine);
    cudaFree(b->out);
    cudaFree(b->nout);
    cudaFree(b->d_in);
    delete b;
}

extern "C" span_b200_fsk_bank_t *span_b200_fsk_bank_create(span_b200_ctx_t *ctx, int channels, const span_b200_fsk_spec_t *spec, int framing_mode)
{
    if (ctx == NULL  ||  channels <= 0  ||  !fsk_spec_ok(spec, framing_mode))
    {
        sb_set_error("bad FSK bank arguments");
        return NULL;
    }
    SB_DEVICE_CKP(span_b200_ctx_device(ctx));
    span_b200_fsk_bank_t *b = new span_b200_fsk_bank_s();
    memset(b, 0, sizeof(*b));
    b->ctx = ctx;
    b->channels = channels;
    b->wspan = 1;
    std::vector<short> t;
    make_dds_int_table(t);
    const size_t C = channels;
    bool ok = cudaMalloc(&b->state, sizeof(int)*K_COUNT*C) == cudaSuccess
              &&  cudaMalloc(&b->window, sizeof(int2)*2*SBF_MAX_WINDOW*C) == cudaSuccess
              &&  cudaMalloc(&b->sine, sizeof(short)*SBF_SINE_PAD) == cudaSuccess
              &&  cudaMalloc(&b->nout, sizeof(int)*C) == cudaSuccess
              &&  cudaMemcpy(b->sine, t.data(), sizeof(short)*SBF_SINE_PAD, cudaMemcpyHostToDevice) == cudaSuccess
              &&  cudaMemset(b->nout, 0, sizeof(int)*C) == cudaSuccess;
    if (!ok)
    {
        sb_set_error("FSK bank allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        span_b200_fsk_bank_destroy(b);
        return NULL;
    }
    const FskSetup su = fsk_setup(b, spec, framing_mode);
    if (fsk_ctl(b, 0, channels, 0, su, 0) != 0)
    {
        span_b200_fsk_bank_destroy(b);
        return NULL;
    }
    return b;
}

extern "C" int span_b200_fsk_bank_channels(const span_b200_fsk_bank_t *b)
{
    return b->channels;
}

extern "C" int span_b200_fsk_bank_restart(span_b200_fsk_bank_t *b, int first, int count, const span_b200_fsk_spec_t *spec, int framing_mode)
{
    if (!fsk_range_ok(b, first, count)  ||  !fsk_spec_ok(spec, framing_mode))
    {
        sb_set_error("bad restart arguments");
        return -1;
    }
    const FskSetup su = fsk_setup(b, spec, framing_mode);
    return fsk_ctl(b, first, count, 1, su, 0);
}

extern "C" int span_b200_fsk_bank_set_signal_cutoff(span_b200_fsk_bank_t *b, int first, int count, float cutoff)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (!fsk_range_ok(b, first, count))
        return -1;
    if (fsk_quiesce(b) != 0)
        return -1;
    const size_t C = b->channels;
    std::vector<int> v(count, host_level_dbm0(cutoff + 2.5f - 5.3f));
    CK(cudaMemcpy(b->state + (size_t) K_ON_POWER*C + first, v.data(), sizeof(int)*count, cudaMemcpyHostToDevice));
    v.assign(count, host_level_dbm0(cutoff - 2.5f - 5.3f));
    CK(cudaMemcpy(b->state + (size_t) K_OFF_POWER*C + first, v.data(), sizeof(int)*count, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int span_b200_fsk_bank_set_frame_parameters(span_b200_fsk_bank_t *b, int first, int count, int data_bits, int parity, int stop_bits)
{
    if (!fsk_range_ok(b, first, count)  ||  data_bits < 1  ||  parity < 0  ||  parity > 4  ||  stop_bits < 0  ||  stop_bits > 255
        ||  data_bits + ((parity != 0)  ?  1  :  0) > 15)
    {
        sb_set_error("bad frame parameters");
        return -1;
    }
    FskSetup su;
    memset(&su, 0, sizeof(su));
    return fsk_ctl(b, first, count, 3, su, data_bits | (parity << 8) | (stop_bits << 16));
}

extern "C" int span_b200_fsk_bank_fillin(span_b200_fsk_bank_t *b, int first, int count, int samples)
{
    if (!fsk_range_ok(b, first, count)  ||  samples < 0)
    {
        sb_set_error("bad fillin arguments");
        return -1;
    }
    FskSetup su;
    memset(&su, 0, sizeof(su));
    return fsk_ctl(b, first, count, 2, su, samples);
}

static int fsk_realloc(void **p, size_t bytes)
{
    if (*p)
        CK(cudaFree(*p));
    *p = NULL;
    CK(cudaMalloc(p, bytes));
    return 0;
}

extern "C" int span_b200_fsk_bank_rx_device(span_b200_fsk_bank_t *b, const int16_t *d_amp, int64_t stride, int n, void *stream)
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
    // Worst case per sample: one status report (carrier up or down) plus the bits of the baud clock
    const long long want = (long long) n + (long long) n*b->max_baud/(8000*100) + 64;
    if (b->out_cap < want)
    {
        if (b->have_last)
            CK(cudaStreamSynchronize(b->last_stream));
        if (fsk_realloc((void **) &b->out, sizeof(short)*(size_t) want*b->channels) != 0)
            return -1;
        b->out_cap = want;
    }
    if (fsk_configure(b) != 0)
        return -1;
    FskArgs a = fsk_args(b, d_amp, stride, n);
    fsk_rx_kernel<<<(b->channels + 31)/32, 32, fsk_smem_bytes(b->wspan), st>>>(a);
    CK(cudaGetLastError());
    b->last_stream = st;
    b->have_last = true;
    return 0;
}

extern "C" int span_b200_fsk_bank_rx_host(span_b200_fsk_bank_t *b, const int16_t *h_amp, int64_t stride, int n, void *stream)
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
        if (fsk_realloc((void **) &b->d_in, want) != 0)
            return -1;
        b->d_in_bytes = want;
    }
    if (n > 0)
        CK(cudaMemcpy2DAsync(b->d_in, sizeof(int16_t)*row, h_amp, sizeof(int16_t)*stride, sizeof(int16_t)*(size_t) n,
                             b->channels, cudaMemcpyHostToDevice, st));
    return span_b200_fsk_bank_rx_device(b, b->d_in, (int64_t) row, n, (void *) st);
}

extern "C" int span_b200_fsk_bank_counts(span_b200_fsk_bank_t *b, int32_t *nout)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (fsk_quiesce(b) != 0)
        return -1;
    CK(cudaMemcpy(nout, b->nout, sizeof(int)*(size_t) b->channels, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int64_t span_b200_fsk_bank_output(span_b200_fsk_bank_t *b, int channel, int16_t *out, int64_t max)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (channel < 0  ||  channel >= b->channels)
        return -1;
    if (fsk_quiesce(b) != 0)
        return -1;
    int n = 0;
    CK(cudaMemcpy(&n, b->nout + channel, sizeof(int), cudaMemcpyDeviceToHost));
    long long k = n;
    if (k > b->out_cap)
        k = b->out_cap;
    if (k > max)
        k = max;
    if (k > 0)
        CK(cudaMemcpy(out, b->out + (size_t) channel*b->out_cap, sizeof(short)*(size_t) k, cudaMemcpyDeviceToHost));
    return k;
}

extern "C" int span_b200_fsk_bank_output_layout(span_b200_fsk_bank_t *b, const int16_t **d_out, int64_t *out_cap, const int32_t **d_nout)
{
    if (d_out)
        *d_out = b->out;
    if (out_cap)
        *out_cap = b->out_cap;
    if (d_nout)
        *d_nout = b->nout;
    return 0;
}

extern "C" int span_b200_fsk_bank_errors(span_b200_fsk_bank_t *b, int channel, int32_t *parity_errors, int32_t *framing_errors, int reset)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (channel < 0  ||  channel >= b->channels)
        return -1;
    if (fsk_quiesce(b) != 0)
        return -1;
    const size_t C = b->channels;
    const int zero = 0;
    if (parity_errors)
    {
        CK(cudaMemcpy(parity_errors, b->state + (size_t) K_PARITY_ERRORS*C + channel, sizeof(int), cudaMemcpyDeviceToHost));
        if (reset)
            CK(cudaMemcpy(b->state + (size_t) K_PARITY_ERRORS*C + channel, &zero, sizeof(int), cudaMemcpyHostToDevice));
    }
    if (framing_errors)
    {
        CK(cudaMemcpy(framing_errors, b->state + (size_t) K_FRAMING_ERRORS*C + channel, sizeof(int), cudaMemcpyDeviceToHost));
        if (reset)
            CK(cudaMemcpy(b->state + (size_t) K_FRAMING_ERRORS*C + channel, &zero, sizeof(int), cudaMemcpyHostToDevice));
    }
    return 0;
}

extern "C" float span_b200_fsk_bank_signal_power(span_b200_fsk_bank_t *b, int channel)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    int reading = 0;
    if (channel < 0  ||  channel >= b->channels  ||  fsk_quiesce(b) != 0)
        return -96.329f + (3.14f + 3.02f);
    if (cudaMemcpy(&reading, b->state + (size_t) K_READING*b->channels + channel, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
        return -96.329f + (3.14f + 3.02f);
    // power_meter_current_dbm0() (src/power_meter.c:115-122)
    if (reading <= 0)
        return -96.329f + (3.14f + 3.02f);
    return 10.0f*log10f((float) reading/(32767.0f*32767.0f) + 1.0e-10f) + (3.14f + 3.02f);
}

extern "C" int span_b200_fsk_bank_channel_state(span_b200_fsk_bank_t *b, int channel, int32_t *info, int32_t *window)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (channel < 0  ||  channel >= b->channels)
        return -1;
    if (fsk_quiesce(b) != 0)
        return -1;
    const size_t C = b->channels;
    if (info)
        CK(cudaMemcpy2D(info, sizeof(int), b->state + channel, sizeof(int)*C, sizeof(int), K_COUNT, cudaMemcpyDeviceToHost));
    if (window)
        CK(cudaMemcpy2D(window, sizeof(int2), b->window + channel, sizeof(int2)*C, sizeof(int2), 2*SBF_MAX_WINDOW, cudaMemcpyDeviceToHost));
    return 0;
}
