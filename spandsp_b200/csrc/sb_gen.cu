// sb_gen.cu - C ABI of the signal source banks (include/spandsp_b200_gen.h).  The generators are sb_gen.cuh.
// Reference: src/dtmf.c (transmit half), src/tone_generate.c, src/dds_float.c, src/awgn.c.
#include <vector>

#include "sb_engine.h"
#include "sb_gen.cuh"

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
    sb_device_guard sb_dg_(span_b200_ctx_device(b->ctx));
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
    sb_device_guard sb_dg_(span_b200_ctx_device(b->ctx));
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
