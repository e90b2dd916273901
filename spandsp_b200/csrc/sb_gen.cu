// sb_gen.cu - C ABI of the signal source banks (include/spandsp_b200_gen.h).  The generators are sb_gen.cuh.
// Reference: src/dtmf.c (transmit half), src/tone_generate.c, src/dds_float.c, src/awgn.c.
#include <vector>

#include "sb_engine.h"
#include "sb_gen.cuh"
#include "sb_modem.cuh"          // host generators of the modem filter tables (make_tx_rrc)

#pragma GCC visibility push(default)
#include "../../include/spandsp_b200_gen.h"
#pragma GCC visibility pop

using namespace sbg;

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

struct span_b200_dtmf_tx_bank_s
{
    span_b200_ctx_t *ctx;
    int channels;
    int *state;
    unsigned char *queue;
    float *sine;
    int *lens;
    int *result;                // per channel outcome of a put
    unsigned char *d_digits;    // staging for put
    size_t d_digits_bytes;
    int *d_lens;
    int16_t *d_out;             // staging for tx_host
    size_t d_out_bytes;
    cudaStream_t last_stream;
    bool have_last;
};

extern "C" int span_b200_dds_float_table(float *table)
{
    host_make_sine_table(table);
    return 0;
}

static DtmfTxArgs tx_args(span_b200_dtmf_tx_bank_t *b, int16_t *d_amp, int64_t stride, int max_samples, int zero_fill)
{
    // dtmf_row[] / dtmf_col[] (src/dtmf.c:114-121)
    static const float row[4] = {697.0f, 770.0f, 852.0f, 941.0f};
    static const float col[4] = {1209.0f, 1336.0f, 1477.0f, 1633.0f};
    DtmfTxArgs a;
    a.amp = d_amp;
    a.stride = stride;
    a.max_samples = max_samples;
    a.channels = b->channels;
    a.zero_fill = zero_fill;
    a.state = b->state;
    a.queue = b->queue;
    a.sine = b->sine;
    a.lens = b->lens;
    for (int i = 0;  i < 4;  i++)
    {
        // tone_gen_descriptor_init() takes the frequencies as int (src/dtmf.c:533-536, tone_generate.c:92,105)
        a.rates[i] = host_dds_phase_ratef((float) (int) row[i]);
        a.rates[4 + i] = host_dds_phase_ratef((float) (int) col[i]);
    }
    return a;
}

static int tx_quiesce(span_b200_dtmf_tx_bank_t *b)
{
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    return 0;
}

static int tx_range_ok(span_b200_dtmf_tx_bank_t *b, int first, int count)
{
    if (b == NULL  ||  first < 0  ||  count < 0  ||  first + count > b->channels)
    {
        sb_set_error("channel range out of bounds");
        return 0;
    }
    return 1;
}

static int tx_ctl(span_b200_dtmf_tx_bank_t *b, int first, int count, int mode, float ga, float gb, int ia, int ib)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (!tx_range_ok(b, first, count))
        return -1;
    if (count == 0)
        return 0;
    if (tx_quiesce(b) != 0)
        return -1;
    cudaStream_t st = (cudaStream_t) sb_ctx_stream(b->ctx);
    DtmfTxArgs a = tx_args(b, NULL, 0, 0, 0);
    dtmf_tx_ctl_kernel<<<(count + 127)/128, 128, 0, st>>>(a, first, count, mode, ga, gb, ia, ib);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int span_b200_dtmf_tx_bank_init(span_b200_dtmf_tx_bank_t *b, int first, int count)
{
    // DEFAULT_DTMF_TX_LEVEL = -10 dBm0 (src/dtmf.c:67)
    return tx_ctl(b, first, count, 0, host_dds_scaling_dbm0f(-10.0f), 0.0f, 0, 0);
}

extern "C" int span_b200_dtmf_tx_bank_set_level(span_b200_dtmf_tx_bank_t *b, int first, int count, int level, int twist)
{
    return tx_ctl(b, first, count, 1, host_dds_scaling_dbm0f((float) level), host_dds_scaling_dbm0f((float) (level + twist)), 0, 0);
}

extern "C" int span_b200_dtmf_tx_bank_set_timing(span_b200_dtmf_tx_bank_t *b, int first, int count, int on_time, int off_time)
{
    return tx_ctl(b, first, count, 2, 0.0f, 0.0f, ((on_time >= 0)  ?  on_time  :  50)*8000/1000, ((off_time >= 0)  ?  off_time  :  55)*8000/1000);
}

extern "C" void span_b200_dtmf_tx_bank_destroy(span_b200_dtmf_tx_bank_t *b)
{
    if (b == NULL)
        return;
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b->have_last)
        cudaStreamSynchronize(b->last_stream);
    cudaFree(b->state);
    cudaFree(b->queue);
    cudaFree(b->sine);
    cudaFree(b->lens);
    cudaFree(b->result);
    cudaFree(b->d_digits);
    cudaFree(b->d_lens);
    cudaFree(b->d_out);
    delete b;
}

extern "C" span_b200_dtmf_tx_bank_t *span_b200_dtmf_tx_bank_create(span_b200_ctx_t *ctx, int channels)
{
    if (ctx == NULL  ||  channels <= 0)
    {
        sb_set_error("bad DTMF transmitter bank arguments");
        return NULL;
    }
    SB_DEVICE_CKP(span_b200_ctx_device(ctx));
    span_b200_dtmf_tx_bank_t *b = new span_b200_dtmf_tx_bank_s();
    memset(b, 0, sizeof(*b));
    b->ctx = ctx;
    b->channels = channels;
    std::vector<float> t(SBG_SINE_WORDS);
    host_make_sine_table(t.data());
    const size_t C = channels;
    bool ok = cudaMalloc(&b->state, sizeof(int)*D_COUNT*C) == cudaSuccess
              &&  cudaMalloc(&b->queue, SBG_QUEUE*C) == cudaSuccess
              &&  cudaMalloc(&b->sine, sizeof(float)*SBG_SINE_WORDS) == cudaSuccess
              &&  cudaMalloc(&b->lens, sizeof(int)*C) == cudaSuccess
              &&  cudaMalloc(&b->result, sizeof(int)*C) == cudaSuccess
              &&  cudaMalloc(&b->d_lens, sizeof(int)*C) == cudaSuccess
              &&  cudaMemcpy(b->sine, t.data(), sizeof(float)*SBG_SINE_WORDS, cudaMemcpyHostToDevice) == cudaSuccess
              &&  cudaMemset(b->queue, 0, SBG_QUEUE*C) == cudaSuccess
              &&  cudaMemset(b->lens, 0, sizeof(int)*C) == cudaSuccess;
    if (!ok)
    {
        sb_set_error("DTMF transmitter bank allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        span_b200_dtmf_tx_bank_destroy(b);
        return NULL;
    }
    if (span_b200_dtmf_tx_bank_init(b, 0, channels) != 0)
    {
        span_b200_dtmf_tx_bank_destroy(b);
        return NULL;
    }
    return b;
}

extern "C" int span_b200_dtmf_tx_bank_channels(const span_b200_dtmf_tx_bank_t *b)
{
    return b->channels;
}

static int gen_realloc(void **p, size_t bytes)
{
    if (*p)
        CK(cudaFree(*p));
    *p = NULL;
    CK(cudaMalloc(p, bytes));
    return 0;
}

static int tx_put(span_b200_dtmf_tx_bank_t *b, int first, int count, const char *digits, int64_t stride, const int32_t *lens, int len_all)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (!tx_range_ok(b, first, count)  ||  digits == NULL)
    {
        sb_set_error("bad put arguments");
        return -1;
    }
    if (count == 0)
        return 0;
    if (tx_quiesce(b) != 0)
        return -1;
    // the digits the kernel will read: one string, or count strings packed at a pitch of SBG_QUEUE
    const size_t want = (lens)  ?  (size_t) count*SBG_QUEUE  :  (size_t) ((len_all > 0)  ?  len_all  :  1);
    if (b->d_digits_bytes < want)
    {
        if (gen_realloc((void **) &b->d_digits, want) != 0)
            return -1;
        b->d_digits_bytes = want;
    }
    std::vector<int> hl;
    if (lens)
    {
        std::vector<unsigned char> pack((size_t) count*SBG_QUEUE, 0);
        hl.assign(lens, lens + count);
        for (int i = 0;  i < count;  i++)
        {
            if (hl[i] < 0)
                hl[i] = (int) strlen(digits + (size_t) i*stride);
            // more than the queue holds can never fit: keep the length (the kernel reports the deficit), copy nothing extra
            memcpy(&pack[(size_t) i*SBG_QUEUE], digits + (size_t) i*stride, (size_t) ((hl[i] < SBG_QUEUE)  ?  hl[i]  :  SBG_QUEUE));
        }
        CK(cudaMemcpy(b->d_digits, pack.data(), pack.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(b->d_lens, hl.data(), sizeof(int)*(size_t) count, cudaMemcpyHostToDevice));
    }
    else if (len_all > 0)
    {
        CK(cudaMemcpy(b->d_digits, digits, (size_t) ((len_all < SBG_QUEUE)  ?  len_all  :  SBG_QUEUE), cudaMemcpyHostToDevice));
    }
    cudaStream_t st = (cudaStream_t) sb_ctx_stream(b->ctx);
    DtmfTxArgs a = tx_args(b, NULL, 0, 0, 0);
    dtmf_tx_put_kernel<<<(count + 127)/128, 128, 0, st>>>(a, first, count, b->d_digits, (lens)  ?  SBG_QUEUE  :  0, (lens)  ?  b->d_lens  :  NULL,
                                                        len_all, b->result);
    CK(cudaGetLastError());
    std::vector<int> res(count);
    CK(cudaMemcpyAsync(res.data(), b->result, sizeof(int)*(size_t) count, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    int worst = 0;
    for (int i = 0;  i < count;  i++)
    {
        if (res[i] > worst)
            worst = res[i];
    }
    return worst;
}

extern "C" int span_b200_dtmf_tx_bank_put(span_b200_dtmf_tx_bank_t *b, int first, int count, const char *digits, int len)
{
    if (digits == NULL)
    {
        sb_set_error("bad put arguments");
        return -1;
    }
    if (len < 0)
    {
        if ((len = (int) strlen(digits)) == 0)
            return 0;                                       // src/dtmf.c:607-611
    }
    return tx_put(b, first, count, digits, 0, NULL, len);
}

extern "C" int span_b200_dtmf_tx_bank_put_each(span_b200_dtmf_tx_bank_t *b, int first, int count, const char *digits, int64_t stride,
                                                const int32_t *lens)
{
    if (lens == NULL)
    {
        sb_set_error("bad put arguments");
        return -1;
    }
    return tx_put(b, first, count, digits, stride, lens, 0);
}

extern "C" int span_b200_dtmf_tx_bank_tx_device(span_b200_dtmf_tx_bank_t *b, int16_t *d_amp, int64_t stride, int max_samples, int zero_fill,
                                                 void *stream)
{
    if (b == NULL  ||  max_samples < 0  ||  (max_samples > 0  &&  d_amp == NULL))
    {
        sb_set_error("bad tx arguments");
        return -1;
    }
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    cudaStream_t st = (stream)  ?  (cudaStream_t) stream  :  (cudaStream_t) sb_ctx_stream(b->ctx);
    if (b->have_last  &&  b->last_stream != st)
        CK(cudaStreamSynchronize(b->last_stream));
    DtmfTxArgs a = tx_args(b, d_amp, stride, max_samples, zero_fill);
    dtmf_tx_kernel<<<(b->channels + 127)/128, 128, 0, st>>>(a);
    CK(cudaGetLastError());
    b->last_stream = st;
    b->have_last = true;
    return 0;
}

extern "C" int span_b200_dtmf_tx_bank_tx_host(span_b200_dtmf_tx_bank_t *b, int16_t *h_amp, int64_t stride, int max_samples, int zero_fill)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b == NULL  ||  max_samples < 0  ||  (max_samples > 0  &&  h_amp == NULL))
    {
        sb_set_error("bad tx arguments");
        return -1;
    }
    if (tx_quiesce(b) != 0)
        return -1;
    const size_t row = ((size_t) max_samples + 7) & ~(size_t) 7;
    const size_t want = sizeof(int16_t)*row*b->channels + 16;
    if (b->d_out_bytes < want)
    {
        if (gen_realloc((void **) &b->d_out, want) != 0)
            return -1;
        b->d_out_bytes = want;
    }
    cudaStream_t st = (cudaStream_t) sb_ctx_stream(b->ctx);
    if (!zero_fill  &&  max_samples > 0)
    {
        // what a channel does not generate must keep the caller's contents
        CK(cudaMemcpy2DAsync(b->d_out, sizeof(int16_t)*row, h_amp, sizeof(int16_t)*stride, sizeof(int16_t)*(size_t) max_samples,
                             b->channels, cudaMemcpyHostToDevice, st));
    }
    if (span_b200_dtmf_tx_bank_tx_device(b, b->d_out, (int64_t) row, max_samples, zero_fill, (void *) st) != 0)
        return -1;
    if (max_samples > 0)
        CK(cudaMemcpy2DAsync(h_amp, sizeof(int16_t)*stride, b->d_out, sizeof(int16_t)*row, sizeof(int16_t)*(size_t) max_samples,
                             b->channels, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int span_b200_dtmf_tx_bank_lens(span_b200_dtmf_tx_bank_t *b, int32_t *lens)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b == NULL  ||  lens == NULL)
        return -1;
    if (tx_quiesce(b) != 0)
        return -1;
    CK(cudaMemcpy(lens, b->lens, sizeof(int)*(size_t) b->channels, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int span_b200_dtmf_tx_bank_sync(span_b200_dtmf_tx_bank_t *b)
{
    if (b == NULL)
        return -1;
    return tx_quiesce(b);
}

// ------------------------------------------------------------------------------------------
struct span_b200_awgn_bank_s
{
    span_b200_ctx_t *ctx;
    int channels;
    int *istate;
    double *dstate;
    int *d_seeds;
    cudaStream_t last_stream;
    bool have_last;
};

static AwgnArgs awgn_args(span_b200_awgn_bank_t *b, int16_t *d_amp, int64_t stride, int n, int add)
{
    AwgnArgs a;
    a.amp = d_amp;
    a.stride = stride;
    a.n = n;
    a.channels = b->channels;
    a.add = add;
    a.istate = b->istate;
    a.dstate = b->dstate;
    return a;
}

static int awgn_init(span_b200_awgn_bank_t *b, int first, int count, const int32_t *seeds, int seed0, float level_dbov)
{
    if (b == NULL  ||  first < 0  ||  count < 0  ||  first + count > b->channels)
    {
        sb_set_error("channel range out of bounds");
        return -1;
    }
    if (count == 0)
        return 0;
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    if (seeds)
        CK(cudaMemcpy(b->d_seeds, seeds, sizeof(int)*(size_t) count, cudaMemcpyHostToDevice));
    // s->rms = pow(10.0, level/20.0)*32768.0 (src/awgn.c:145), level a float
    const double rms = pow(10.0, level_dbov/20.0)*32768.0;
    cudaStream_t st = (cudaStream_t) sb_ctx_stream(b->ctx);
    AwgnArgs a = awgn_args(b, NULL, 0, 0, 0);
    awgn_init_kernel<<<(count + 127)/128, 128, 0, st>>>(a, first, count, (seeds)  ?  b->d_seeds  :  NULL, seed0, rms);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int span_b200_awgn_bank_init_dbov(span_b200_awgn_bank_t *b, int first, int count, const int32_t *seeds, int seed0, float level)
{
    return awgn_init(b, first, count, seeds, seed0, level);
}

extern "C" int span_b200_awgn_bank_init_dbm0(span_b200_awgn_bank_t *b, int first, int count, const int32_t *seeds, int seed0, float level)
{
    return awgn_init(b, first, count, seeds, seed0, level - (3.14f + 3.02f));     // DBM0_MAX_POWER (src/awgn.c:154)
}

extern "C" void span_b200_awgn_bank_destroy(span_b200_awgn_bank_t *b)
{
    if (b == NULL)
        return;
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b->have_last)
        cudaStreamSynchronize(b->last_stream);
    cudaFree(b->istate);
    cudaFree(b->dstate);
    cudaFree(b->d_seeds);
    delete b;
}

extern "C" span_b200_awgn_bank_t *span_b200_awgn_bank_create(span_b200_ctx_t *ctx, int channels, const int32_t *seeds, int seed0, float level_dbm0)
{
    if (ctx == NULL  ||  channels <= 0)
    {
        sb_set_error("bad noise bank arguments");
        return NULL;
    }
    SB_DEVICE_CKP(span_b200_ctx_device(ctx));
    span_b200_awgn_bank_t *b = new span_b200_awgn_bank_s();
    memset(b, 0, sizeof(*b));
    b->ctx = ctx;
    b->channels = channels;
    const size_t C = channels;
    bool ok = cudaMalloc(&b->istate, sizeof(int)*4*C) == cudaSuccess
              &&  cudaMalloc(&b->dstate, sizeof(double)*(2 + SBG_RAN_TABLE)*C) == cudaSuccess
              &&  cudaMalloc(&b->d_seeds, sizeof(int)*C) == cudaSuccess;
    if (!ok)
    {
        sb_set_error("noise bank allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        span_b200_awgn_bank_destroy(b);
        return NULL;
    }
    if (span_b200_awgn_bank_init_dbm0(b, 0, channels, seeds, seed0, level_dbm0) != 0)
    {
        span_b200_awgn_bank_destroy(b);
        return NULL;
    }
    return b;
}

extern "C" int span_b200_awgn_bank_channels(const span_b200_awgn_bank_t *b)
{
    return b->channels;
}

static int awgn_run(span_b200_awgn_bank_t *b, int16_t *d_amp, int64_t stride, int n, int add, void *stream)
{
    if (b == NULL  ||  n < 0  ||  (n > 0  &&  d_amp == NULL))
    {
        sb_set_error("bad noise arguments");
        return -1;
    }
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    cudaStream_t st = (stream)  ?  (cudaStream_t) stream  :  (cudaStream_t) sb_ctx_stream(b->ctx);
    if (b->have_last  &&  b->last_stream != st)
        CK(cudaStreamSynchronize(b->last_stream));
    AwgnArgs a = awgn_args(b, d_amp, stride, n, add);
    awgn_kernel<<<(b->channels + 31)/32, 32, 0, st>>>(a);
    CK(cudaGetLastError());
    b->last_stream = st;
    b->have_last = true;
    return 0;
}

extern "C" int span_b200_awgn_bank_add_device(span_b200_awgn_bank_t *b, int16_t *d_amp, int64_t stride, int samples, void *stream)
{
    return awgn_run(b, d_amp, stride, samples, 1, stream);
}

extern "C" int span_b200_awgn_bank_fill_device(span_b200_awgn_bank_t *b, int16_t *d_amp, int64_t stride, int samples, void *stream)
{
    return awgn_run(b, d_amp, stride, samples, 0, stream);
}

extern "C" int span_b200_awgn_bank_sync(span_b200_awgn_bank_t *b)
{
    if (b == NULL)
        return -1;
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    return 0;
}

// ------------------------------------------------------------------------------------------
// generic bank plumbing shared by the tone_gen and V.29 transmitter banks
struct gen_bank_base
{
    span_b200_ctx_t *ctx;
    int channels;
    int *state;
    float *sine;
    int *lens;
    cudaStream_t last_stream;
    bool have_last;
};

static int gen_quiesce(gen_bank_base *b)
{
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    return 0;
}

static int gen_range_ok(const gen_bank_base *b, int first, int count)
{
    if (b == NULL  ||  first < 0  ||  count < 0  ||  first + count > b->channels)
    {
        sb_set_error("channel range out of bounds");
        return 0;
    }
    return 1;
}

static bool gen_base_alloc(gen_bank_base *b, span_b200_ctx_t *ctx, int channels, int fields)
{
    sb_device_guard sb_dg_(span_b200_ctx_device(ctx));
    b->ctx = ctx;
    b->channels = channels;
    std::vector<float> t(SBG_SINE_WORDS);
    host_make_sine_table(t.data());
    const size_t C = channels;
    return cudaMalloc(&b->state, sizeof(int)*(size_t) fields*C) == cudaSuccess
           &&  cudaMemset(b->state, 0, sizeof(int)*(size_t) fields*C) == cudaSuccess
           &&  cudaMalloc(&b->sine, sizeof(float)*SBG_SINE_WORDS) == cudaSuccess
           &&  cudaMalloc(&b->lens, sizeof(int)*C) == cudaSuccess
           &&  cudaMemcpy(b->sine, t.data(), sizeof(float)*SBG_SINE_WORDS, cudaMemcpyHostToDevice) == cudaSuccess
           &&  cudaMemset(b->lens, 0, sizeof(int)*C) == cudaSuccess;
}

static void gen_base_free(gen_bank_base *b)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b->have_last)
        cudaStreamSynchronize(b->last_stream);
    cudaFree(b->state);
    cudaFree(b->sine);
    cudaFree(b->lens);
}

static int gen_lens(gen_bank_base *b, int32_t *lens)
{
    if (b == NULL  ||  lens == NULL)
        return -1;
    if (gen_quiesce(b) != 0)
        return -1;
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    CK(cudaMemcpy(lens, b->lens, sizeof(int)*(size_t) b->channels, cudaMemcpyDeviceToHost));
    return 0;
}

// ------------------------------------------------------------------------------------------
// tone_gen banks
struct span_b200_tone_gen_bank_s : gen_bank_base
{
    ToneDesc *d_descs;
    size_t d_descs_n;
};

static ToneGenArgs tg_args(span_b200_tone_gen_bank_t *b, int16_t *d_amp, int64_t stride, int max_samples, int zero_fill)
{
    ToneGenArgs a;
    a.amp = d_amp;
    a.stride = stride;
    a.max_samples = max_samples;
    a.channels = b->channels;
    a.zero_fill = zero_fill;
    a.state = b->state;
    a.sine = b->sine;
    a.lens = b->lens;
    return a;
}

extern "C" void span_b200_tone_gen_bank_destroy(span_b200_tone_gen_bank_t *b)
{
    if (b == NULL)
        return;
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    gen_base_free(b);
    cudaFree(b->d_descs);
    delete b;
}

extern "C" span_b200_tone_gen_bank_t *span_b200_tone_gen_bank_create(span_b200_ctx_t *ctx, int channels)
{
    if (ctx == NULL  ||  channels <= 0)
    {
        sb_set_error("bad tone generator bank arguments");
        return NULL;
    }
    SB_DEVICE_CKP(span_b200_ctx_device(ctx));
    span_b200_tone_gen_bank_t *b = new span_b200_tone_gen_bank_s();
    memset(b, 0, sizeof(*b));
    if (!gen_base_alloc(b, ctx, channels, T_COUNT))
    {
        sb_set_error("tone generator bank allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        span_b200_tone_gen_bank_destroy(b);
        return NULL;
    }
    // idle: tone_gen() returns 0 while current_section < 0 (src/tone_generate.c:139-140)
    std::vector<int> idle((size_t) channels, -1);
    if (cudaMemcpy(b->state + (size_t) T_SECTION*channels, idle.data(), sizeof(int)*(size_t) channels, cudaMemcpyHostToDevice) != cudaSuccess)
    {
        sb_set_error("tone generator bank setup failed");
        span_b200_tone_gen_bank_destroy(b);
        return NULL;
    }
    return b;
}

extern "C" int span_b200_tone_gen_bank_channels(const span_b200_tone_gen_bank_t *b)
{
    return b->channels;
}

static int tg_init(span_b200_tone_gen_bank_t *b, int first, int count, const span_b200_tone_desc_t *descs, int same)
{
    if (!gen_range_ok(b, first, count)  ||  descs == NULL)
    {
        sb_set_error("bad tone generator init arguments");
        return -1;
    }
    if (count == 0)
        return 0;
    if (gen_quiesce(b) != 0)
        return -1;
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    const size_t n = (same)  ?  1  :  (size_t) count;
    std::vector<ToneDesc> h(n);
    for (size_t i = 0;  i < n;  i++)
        host_tone_descriptor(h[i], descs[i].f1, descs[i].l1, descs[i].f2, descs[i].l2, descs[i].d1, descs[i].d2, descs[i].d3, descs[i].d4, descs[i].repeat);
    if (b->d_descs_n < n)
    {
        if (gen_realloc((void **) &b->d_descs, sizeof(ToneDesc)*n) != 0)
            return -1;
        b->d_descs_n = n;
    }
    CK(cudaMemcpy(b->d_descs, h.data(), sizeof(ToneDesc)*n, cudaMemcpyHostToDevice));
    cudaStream_t st = (cudaStream_t) sb_ctx_stream(b->ctx);
    tone_gen_init_kernel<<<(count + 127)/128, 128, 0, st>>>(tg_args(b, NULL, 0, 0, 0), first, count, b->d_descs, same);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int span_b200_tone_gen_bank_init(span_b200_tone_gen_bank_t *b, int first, int count, const span_b200_tone_desc_t *desc)
{
    return tg_init(b, first, count, desc, 1);
}

extern "C" int span_b200_tone_gen_bank_init_each(span_b200_tone_gen_bank_t *b, int first, int count, const span_b200_tone_desc_t *descs)
{
    return tg_init(b, first, count, descs, 0);
}

extern "C" int span_b200_tone_gen_bank_tx_device(span_b200_tone_gen_bank_t *b, int16_t *d_amp, int64_t stride, int max_samples, int zero_fill,
                                                  void *stream)
{
    if (b == NULL  ||  max_samples < 0  ||  (max_samples > 0  &&  d_amp == NULL))
    {
        sb_set_error("bad tx arguments");
        return -1;
    }
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    cudaStream_t st = (stream)  ?  (cudaStream_t) stream  :  (cudaStream_t) sb_ctx_stream(b->ctx);
    if (b->have_last  &&  b->last_stream != st)
        CK(cudaStreamSynchronize(b->last_stream));
    tone_gen_kernel<<<(b->channels + 127)/128, 128, 0, st>>>(tg_args(b, d_amp, stride, max_samples, zero_fill));
    CK(cudaGetLastError());
    b->last_stream = st;
    b->have_last = true;
    return 0;
}

extern "C" int span_b200_tone_gen_bank_lens(span_b200_tone_gen_bank_t *b, int32_t *lens)
{
    return gen_lens(b, lens);
}

extern "C" int span_b200_tone_gen_bank_sync(span_b200_tone_gen_bank_t *b)
{
    return (b)  ?  gen_quiesce(b)  :  -1;
}

// ------------------------------------------------------------------------------------------
// V.29 transmitter banks
struct span_b200_v29_tx_bank_s : gen_bank_base
{
    float *shaper;
    unsigned char *bits;
    int64_t bits_stride;
    unsigned int *d_seeds;
    int *d_counts;
};

extern "C" int span_b200_v29_tx_tables(float *shaper)
{
    std::vector<float> t;
    sbm::make_tx_rrc(t, SBG_V29_TX_SETS, SBG_V29_TX_STEPS, 0.25);          // src/make_modem_filter.c:401-413
    memcpy(shaper, t.data(), sizeof(float)*t.size());
    return 0;
}

static V29TxArgs vt_args(span_b200_v29_tx_bank_t *b, int16_t *d_amp, int64_t stride, int max_samples, int zero_fill)
{
    V29TxArgs a;
    a.amp = d_amp;
    a.stride = stride;
    a.max_samples = max_samples;
    a.channels = b->channels;
    a.zero_fill = zero_fill;
    a.state = b->state;
    a.bits = b->bits;
    a.bits_stride = b->bits_stride;
    a.sine = b->sine;
    a.shaper = b->shaper;
    a.lens = b->lens;
    a.carrier_phase_rate = host_dds_phase_ratef(1700.0f);                   // CARRIER_NOMINAL_FREQ, src/v29tx.c:79
    return a;
}

// v29_tx_power() (src/v29tx.c:323-338), float build: TX_PULSESHAPER_GAIN is 1
static float v29_tx_base_gain(float power)
{
    return powf(10.0f, (power - 3.14f)/20.0f)*32768.0f/1.000000f;
}

static int vt_ctl(span_b200_v29_tx_bank_t *b, int first, int count, int mode, float ga, int ia, int ib, const unsigned int *seeds, const int *counts)
{
    if (!gen_range_ok(b, first, count))
        return -1;
    if (count == 0)
        return 0;
    if (gen_quiesce(b) != 0)
        return -1;
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    cudaStream_t st = (cudaStream_t) sb_ctx_stream(b->ctx);
    v29_tx_ctl_kernel<<<(count + 127)/128, 128, 0, st>>>(vt_args(b, NULL, 0, 0, 0), first, count, mode, ga, ia, ib, seeds, counts);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

static bool v29_rate_ok(int bit_rate)
{
    return bit_rate == 9600  ||  bit_rate == 7200  ||  bit_rate == 4800;
}

extern "C" void span_b200_v29_tx_bank_destroy(span_b200_v29_tx_bank_t *b)
{
    if (b == NULL)
        return;
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    gen_base_free(b);
    cudaFree(b->shaper);
    cudaFree(b->bits);
    cudaFree(b->d_seeds);
    cudaFree(b->d_counts);
    delete b;
}

extern "C" span_b200_v29_tx_bank_t *span_b200_v29_tx_bank_create(span_b200_ctx_t *ctx, int channels, int bit_rate, int tep)
{
    if (ctx == NULL  ||  channels <= 0  ||  !v29_rate_ok(bit_rate))
    {
        sb_set_error("bad V.29 transmitter bank arguments");        // src/v29tx.c:403-412: unknown rates are refused
        return NULL;
    }
    SB_DEVICE_CKP(span_b200_ctx_device(ctx));
    span_b200_v29_tx_bank_t *b = new span_b200_v29_tx_bank_s();
    memset(b, 0, sizeof(*b));
    float shaper[SBG_V29_TX_SETS*SBG_V29_TX_STEPS];
    span_b200_v29_tx_tables(shaper);
    const size_t C = channels;
    bool ok = gen_base_alloc(b, ctx, channels, X_COUNT)
              &&  cudaMalloc(&b->shaper, sizeof(shaper)) == cudaSuccess
              &&  cudaMemcpy(b->shaper, shaper, sizeof(shaper), cudaMemcpyHostToDevice) == cudaSuccess
              &&  cudaMalloc(&b->d_seeds, sizeof(unsigned int)*C) == cudaSuccess
              &&  cudaMalloc(&b->d_counts, sizeof(int)*C) == cudaSuccess;
    if (!ok)
    {
        sb_set_error("V.29 transmitter bank allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        span_b200_v29_tx_bank_destroy(b);
        return NULL;
    }
    if (vt_ctl(b, 0, channels, 0, v29_tx_base_gain(-14.0f), bit_rate, tep != 0, NULL, NULL) != 0
        ||  vt_ctl(b, 0, channels, 3, 0.0f, 1, 0, NULL, NULL) != 0)
    {
        span_b200_v29_tx_bank_destroy(b);
        return NULL;
    }
    return b;
}

extern "C" int span_b200_v29_tx_bank_channels(const span_b200_v29_tx_bank_t *b)
{
    return b->channels;
}

extern "C" int span_b200_v29_tx_bank_restart(span_b200_v29_tx_bank_t *b, int first, int count, int bit_rate, int tep)
{
    if (!v29_rate_ok(bit_rate))
    {
        sb_set_error("V.29 bit rate %d", bit_rate);
        return -1;                                              // src/v29tx.c:377-378
    }
    return vt_ctl(b, first, count, 1, 0.0f, bit_rate, tep != 0, NULL, NULL);
}

extern "C" int span_b200_v29_tx_bank_power(span_b200_v29_tx_bank_t *b, int first, int count, float power)
{
    return vt_ctl(b, first, count, 2, v29_tx_base_gain(power), 0, 0, NULL, NULL);
}

extern "C" int span_b200_v29_tx_bank_set_prbs(span_b200_v29_tx_bank_t *b, int first, int count, const uint32_t *seeds, uint32_t seed0)
{
    if (!gen_range_ok(b, first, count))
        return -1;
    if (count == 0)
        return 0;
    if (seeds)
    {
        if (gen_quiesce(b) != 0)
            return -1;
        SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
        CK(cudaMemcpy(b->d_seeds, seeds, sizeof(unsigned int)*(size_t) count, cudaMemcpyHostToDevice));
    }
    return vt_ctl(b, first, count, 3, 0.0f, (int) seed0, 0, (seeds)  ?  b->d_seeds  :  NULL, NULL);
}

extern "C" int span_b200_v29_tx_bank_set_bits(span_b200_v29_tx_bank_t *b, int first, int count, const uint8_t *bits, int64_t stride_bytes,
                                               const int32_t *nbits)
{
    if (!gen_range_ok(b, first, count)  ||  bits == NULL  ||  nbits == NULL  ||  stride_bytes < 0)
    {
        sb_set_error("bad bit source arguments");
        return -1;
    }
    if (count == 0)
        return 0;
    if (gen_quiesce(b) != 0)
        return -1;
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    int64_t need = 1;
    for (int i = 0;  i < count;  i++)
    {
        if (nbits[i] < 0  ||  ((int64_t) nbits[i] + 7)/8 > stride_bytes)
        {
            sb_set_error("channel %d: %d bits do not fit a stride of %lld bytes", first + i, nbits[i], (long long) stride_bytes);
            return -1;
        }
        need = std::max(need, ((int64_t) nbits[i] + 7)/8);
    }
    if (b->bits == NULL  ||  b->bits_stride < need)
    {
        // one pitch for the whole bank; a wider request re-creates the buffer (sources of other channels are lost:
        // set the widest first)
        if (gen_realloc((void **) &b->bits, (size_t) need*b->channels) != 0)
            return -1;
        CK(cudaMemset(b->bits, 0, (size_t) need*b->channels));
        b->bits_stride = need;
    }
    CK(cudaMemcpy2D(b->bits + (size_t) first*b->bits_stride, (size_t) b->bits_stride, bits, (size_t) stride_bytes, (size_t) need, (size_t) count,
                    cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b->d_counts, nbits, sizeof(int)*(size_t) count, cudaMemcpyHostToDevice));
    return vt_ctl(b, first, count, 4, 0.0f, 0, 0, NULL, b->d_counts);
}

extern "C" int span_b200_v29_tx_bank_tx_device(span_b200_v29_tx_bank_t *b, int16_t *d_amp, int64_t stride, int max_samples, int zero_fill,
                                                void *stream)
{
    if (b == NULL  ||  max_samples < 0  ||  (max_samples > 0  &&  d_amp == NULL))
    {
        sb_set_error("bad tx arguments");
        return -1;
    }
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    cudaStream_t st = (stream)  ?  (cudaStream_t) stream  :  (cudaStream_t) sb_ctx_stream(b->ctx);
    if (b->have_last  &&  b->last_stream != st)
        CK(cudaStreamSynchronize(b->last_stream));
    v29_tx_kernel<<<(b->channels + 127)/128, 128, 0, st>>>(vt_args(b, d_amp, stride, max_samples, zero_fill));
    CK(cudaGetLastError());
    b->last_stream = st;
    b->have_last = true;
    return 0;
}

extern "C" int span_b200_v29_tx_bank_lens(span_b200_v29_tx_bank_t *b, int32_t *lens)
{
    return gen_lens(b, lens);
}

extern "C" int span_b200_v29_tx_bank_status(span_b200_v29_tx_bank_t *b, int32_t *status)
{
    if (b == NULL  ||  status == NULL)
        return -1;
    if (gen_quiesce(b) != 0)
        return -1;
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    CK(cudaMemcpy(status, b->state + (size_t) X_STATUS*b->channels, sizeof(int)*(size_t) b->channels, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int span_b200_v29_tx_bank_sync(span_b200_v29_tx_bank_t *b)
{
    return (b)  ?  gen_quiesce(b)  :  -1;
}
