// sb_modem_bank.cuh - host side of the modem receiver banks (V.29, V.17, V.27ter): device tables, per-channel
// state arrays, launches, result read-back.  The C ABI functions of include/spandsp_b200_v29.h and
// include/spandsp_b200_v17.h are thin wrappers over these templates.
#pragma once

#include <vector>

#include "sb_engine.h"
#include "sb_modem.cuh"

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

#define CKP(call) \
    do \
    { \
        cudaError_t e_ = (call); \
        if (e_ != cudaSuccess) \
        { \
            sb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return NULL; \
        } \
    } \
    while (0)

namespace sbm {

// The kernel a bank's rx call runs: by default the receiver itself, one thread per receiver, one warp per CTA
template <class RX>
struct RxKernel
{
    typedef RX type;
    static const int WARPS = 1;
};

#define SBM_HOST_PIECES 4

template <class RX>
struct ModemBank
{
    span_b200_ctx_t *ctx;
    int channels;
    int bit_rate;
    float *fstate;
    int *istate;
    std::vector<void *> owned;          // device allocations holding the constant tables
    typename RX::Consts k;
    unsigned int *words;                // packed data bits, [channel][words_cap]
    long long words_cap;
    int *status;                        // status reports, [channel][status_cap][2]
    long long status_cap;
    int *nbits;
    int *nstatus;
    span_b200_v29_symbol_t *syms;
    long long sym_cap;
    int *nsyms;
    int want_symbols;
    int16_t *d_in;
    size_t d_in_bytes;
    cudaStream_t copy_stream;           // rx_host: the pieces of a call cross PCIe here while the kernel of the piece before runs
    cudaEvent_t copied[SBM_HOST_PIECES];
    cudaEvent_t in_free;                // the kernels that read d_in have finished
    cudaStream_t last_stream;
    bool have_last;
    int on_power;
    int off_power;
    int bits_per_sample_x2;             // output capacity: put_bit calls per input sample, times two
};

template <class T>
static int modem_upload(std::vector<void *> &owned, const T **dst, const void *src, size_t bytes)
{
    void *p = NULL;
    CK(cudaMalloc(&p, bytes));
    owned.push_back(p);
    CK(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
    *dst = (const T *) p;
    return 0;
}

// Tables every receiver needs: RRC sets, sine table, sqrt table, Godard descriptor, carrier constants.
template <class RX>
static int modem_core_tables(ModemBank<RX> *b, double carrier_hz, double godard_fine_trigger, int godard_coarse_step, float agc_target,
                             float agc_reference = 735.0f, const std::vector<float> *rrc_re_rows = NULL, const std::vector<float> *rrc_im_rows = NULL)
{
    std::vector<float> re;
    std::vector<float> im;
    std::vector<float> st;
    std::vector<unsigned short> sq;
    godard_desc_t g;
    if (rrc_re_rows)
    {
        re = *rrc_re_rows;          // [RX::SETS][27], prepared by the caller
        im = *rrc_im_rows;
    }
    else
    {
        make_rx_rrc(re, im, RX::SETS, carrier_hz);
    }
    make_sine_table(st);
    make_sqrt_table(sq);
    make_godard(g, carrier_hz, godard_fine_trigger, godard_coarse_step);
    if (modem_upload(b->owned, &b->k.rrc_re, re.data(), sizeof(float)*re.size()) != 0
        ||
        modem_upload(b->owned, &b->k.rrc_im, im.data(), sizeof(float)*im.size()) != 0
        ||
        modem_upload(b->owned, &b->k.sine, st.data(), sizeof(float)*st.size()) != 0
        ||
        modem_upload(b->owned, &b->k.sqrt_tab, sq.data(), sizeof(unsigned short)*sq.size()) != 0)
    {
        return -1;
    }
    for (int i = 0;  i < 3;  i++)
    {
        b->k.g_low[i] = g.low[i];
        b->k.g_high[i] = g.high[i];
    }
    b->k.g_mixed3 = g.mixed3;
    b->k.g_coarse_trigger = g.coarse_trigger;
    b->k.g_fine_trigger = g.fine_trigger;
    b->k.g_coarse_step = g.coarse_step;
    b->k.g_fine_step = g.fine_step;
    const float carrier = (float) carrier_hz;
    b->k.rate_nominal = host_dds_phase_rate(carrier);
    b->k.rate_low = host_dds_phase_rate(carrier - 20.0f);
    b->k.rate_high = host_dds_phase_rate(carrier + 20.0f);
    b->k.agc_target = agc_target/1.000000f;                 // RX_PULSESHAPER_GAIN is 1.0 in the float build
    b->k.agc_initial = (agc_target/1.000000f)/agc_reference;    // src/v29rx.c:1078, src/v17rx.c:1474, src/v27ter_rx.c:1141
    return 0;
}

// KRX: the receiver type of the kernel the arguments are for (RX itself, or its RxKernel<RX>::type)
template <class KRX, class RX>
static KernelArgs<KRX> modem_args_for(ModemBank<RX> *b, const int16_t *d_amp, int64_t stride, int n)
{
    KernelArgs<KRX> ka;
    ka.a.amp = d_amp;
    ka.a.stride = stride;
    ka.a.n = n;
    ka.a.channels = b->channels;
    ka.a.fstate = b->fstate;
    ka.a.istate = b->istate;
    ka.a.words = b->words;
    ka.a.words_cap = b->words_cap;
    ka.a.status = b->status;
    ka.a.status_cap = b->status_cap;
    ka.a.nbits = b->nbits;
    ka.a.nstatus = b->nstatus;
    ka.a.syms = (b->want_symbols)  ?  b->syms  :  NULL;
    ka.a.sym_cap = b->sym_cap;
    ka.a.nsyms = b->nsyms;
    ka.a.append = 0;
    ka.k = b->k;
    return ka;
}

template <class RX>
static KernelArgs<RX> modem_args(ModemBank<RX> *b, const int16_t *d_amp, int64_t stride, int n)
{
    return modem_args_for<RX, RX>(b, d_amp, stride, n);
}

template <class RX>
static int modem_configure()
{
    // function attributes are per device: one flag per device ordinal, not one per process
    static bool configured[64] = {false};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0  ||  dev >= 64  ||  !configured[dev])
    {
        typedef typename RxKernel<RX>::type KRX;
        const int smem = (int) sizeof(float)*modem_smem_words<RX>();
        const int ksmem = (int) sizeof(float)*modem_smem_words<KRX>(RxKernel<RX>::WARPS);
        CK(cudaFuncSetAttribute(modem_rx_kernel<KRX, RxKernel<RX>::WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, ksmem));
        CK(cudaFuncSetAttribute(modem_init_kernel<RX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (dev >= 0  &&  dev < 64)
            configured[dev] = true;
    }
    return 0;
}

// mode < 0: xxx_rx_init(); mode >= 0: xxx_rx_restart(s, bit_rate, mode)
template <class RX>
static int modem_init_channels(ModemBank<RX> *b, int first, int count, int bit_rate, int mode)
{
    if (count <= 0)
        return 0;
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    if (modem_configure<RX>() != 0)
        return -1;
    const int smem = (int) sizeof(float)*modem_smem_words<RX>();
    cudaStream_t st = (cudaStream_t) sb_ctx_stream(b->ctx);
    KernelArgs<RX> ka = modem_args(b, NULL, 0, 0);
    modem_init_kernel<RX><<<(count + 31)/32, 32, smem, st>>>(ka, first, count, bit_rate, mode, b->on_power, b->off_power);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    return 0;
}

template <class RX>
static int modem_alloc_state(ModemBank<RX> *b)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    const size_t C = b->channels;
    CK(cudaMalloc(&b->fstate, sizeof(float)*RX::F_COUNT*C));
    CK(cudaMalloc(&b->istate, sizeof(int)*RX::I_COUNT*C));
    CK(cudaMemset(b->fstate, 0, sizeof(float)*RX::F_COUNT*C));
    CK(cudaMemset(b->istate, 0, sizeof(int)*RX::I_COUNT*C));
    CK(cudaMalloc(&b->nbits, sizeof(int)*C));
    CK(cudaMalloc(&b->nstatus, sizeof(int)*C));
    CK(cudaMalloc(&b->nsyms, sizeof(int)*C));
    CK(cudaMemset(b->nbits, 0, sizeof(int)*C));
    CK(cudaMemset(b->nstatus, 0, sizeof(int)*C));
    CK(cudaMemset(b->nsyms, 0, sizeof(int)*C));
    return 0;
}

template <class RX>
static void modem_destroy(ModemBank<RX> *b)
{
    if (b == NULL)
        return;
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b->have_last)
        cudaStreamSynchronize(b->last_stream);
    cudaFree(b->fstate);
    cudaFree(b->istate);
    for (size_t i = 0;  i < b->owned.size();  i++)
        cudaFree(b->owned[i]);
    cudaFree(b->words);
    cudaFree(b->status);
    cudaFree(b->nbits);
    cudaFree(b->nstatus);
    cudaFree(b->syms);
    cudaFree(b->nsyms);
    cudaFree(b->d_in);
    if (b->copy_stream)
    {
        cudaStreamSynchronize(b->copy_stream);
        cudaStreamDestroy(b->copy_stream);
    }
    for (int i = 0;  i < SBM_HOST_PIECES;  i++)
    {
        if (b->copied[i])
            cudaEventDestroy(b->copied[i]);
    }
    if (b->in_free)
        cudaEventDestroy(b->in_free);
    delete b;
}

template <class RX>
static int modem_quiesce(ModemBank<RX> *b)
{
    SB_DEVICE_CK(span_b200_ctx_device(b->ctx));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    return 0;
}

template <class RX>
static int modem_range_ok(ModemBank<RX> *b, int first, int count)
{
    if (b == NULL  ||  first < 0  ||  count < 0  ||  first + count > b->channels)
    {
        sb_set_error("channel range out of bounds");
        return 0;
    }
    return 1;
}

// xxx_rx_set_signal_cutoff() (src/v29rx.c:163-168, src/v17rx.c:173-178)
template <class RX>
static int modem_set_signal_cutoff(ModemBank<RX> *b, int first, int count, float cutoff)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (!modem_range_ok(b, first, count))
        return -1;
    if (modem_quiesce(b) != 0)
        return -1;
    const int on = (int32_t) (host_power_meter_level_dbm0(cutoff + 2.5f)*0.4f);
    const int off = (int32_t) (host_power_meter_level_dbm0(cutoff - 2.5f)*0.4f);
    if (first == 0  &&  count == b->channels)
    {
        b->on_power = on;
        b->off_power = off;
    }
    std::vector<int> v(count, on);
    CK(cudaMemcpy(b->istate + (size_t) I_ON_POWER*b->channels + first, v.data(), sizeof(int)*count, cudaMemcpyHostToDevice));
    v.assign(count, off);
    CK(cudaMemcpy(b->istate + (size_t) I_OFF_POWER*b->channels + first, v.data(), sizeof(int)*count, cudaMemcpyHostToDevice));
    return 0;
}

// xxx_rx_fillin(): integer bookkeeping only (src/v29rx.c:967-996, src/v17rx.c:1313-1343, src/v27ter_rx.c:1030-1068); done on the
// host copy of four fields.
template <class RX>
static int modem_fillin(ModemBank<RX> *b, int first, int count, int samples)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (!modem_range_ok(b, first, count)  ||  samples < 0)
    {
        sb_set_error("bad fillin arguments");
        return -1;
    }
    if (modem_quiesce(b) != 0)
        return -1;
    const size_t C = b->channels;
    std::vector<int> present(count), stage(count), phase(count), rate(count), put(count), bps(count);
    CK(cudaMemcpy(bps.data(), b->istate + I_BIT_RATE*C + first, sizeof(int)*count, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(present.data(), b->istate + I_SIGNAL_PRESENT*C + first, sizeof(int)*count, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(stage.data(), b->istate + I_STAGE*C + first, sizeof(int)*count, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(phase.data(), b->istate + I_CARRIER_PHASE*C + first, sizeof(int)*count, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rate.data(), b->istate + I_PHASE_RATE*C + first, sizeof(int)*count, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(put.data(), b->istate + I_EQ_PUT_STEP*C + first, sizeof(int)*count, cudaMemcpyDeviceToHost));
    for (int c = 0;  c < count;  c++)
    {
        if (present[c] <= 0  ||  stage[c] == RX::STAGE_PARKED)
            continue;
        unsigned int ph = (unsigned int) phase[c];
        for (int i = 0;  i < samples;  i++)
        {
            ph += (unsigned int) rate[c];
            put[c] -= RX::fillin_sets(bps[c]);
            if (put[c] <= 0)
                put[c] += RX::fillin_half_baud(bps[c]);
        }
        phase[c] = (int) ph;
    }
    CK(cudaMemcpy(b->istate + I_CARRIER_PHASE*C + first, phase.data(), sizeof(int)*count, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b->istate + I_EQ_PUT_STEP*C + first, put.data(), sizeof(int)*count, cudaMemcpyHostToDevice));
    return 0;
}

static inline int modem_realloc(void **p, size_t bytes)
{
    if (*p)
        CK(cudaFree(*p));
    *p = NULL;
    CK(cudaMalloc(p, bytes));
    return 0;
}

// append: this call continues the output of the previous one (modem_rx_host feeds a long call in pieces); n_total: the
// samples of the whole call, which the output buffers are sized for
template <class RX>
static int modem_rx_device(ModemBank<RX> *b, const int16_t *d_amp, int64_t stride, int n, void *stream, bool append = false, int n_total = -1)
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
    // Worst case: bits per baud at 2400 baud/8000 Hz plus timing drift; status reports: a carrier cycle (up, training,
    // result, down) takes hundreds of samples.
    if (n_total < n)
        n_total = n;
    const long long want_words = ((long long) n_total*b->bits_per_sample_x2/2 + 64 + 31)/32 + 1;
    if (b->words_cap < want_words)
    {
        if (b->have_last)
            CK(cudaStreamSynchronize(b->last_stream));
        if (modem_realloc((void **) &b->words, sizeof(unsigned int)*(size_t) want_words*b->channels) != 0)
            return -1;
        b->words_cap = want_words;
    }
    const long long want_status = (long long) n_total/64 + 16;
    if (b->status_cap < want_status)
    {
        if (b->have_last)
            CK(cudaStreamSynchronize(b->last_stream));
        if (modem_realloc((void **) &b->status, sizeof(int)*2*(size_t) want_status*b->channels) != 0)
            return -1;
        b->status_cap = want_status;
    }
    const long long want_syms = (long long) n_total*2/5 + 16;
    if (b->want_symbols  &&  b->sym_cap < want_syms)
    {
        if (b->have_last)
            CK(cudaStreamSynchronize(b->last_stream));
        if (modem_realloc((void **) &b->syms, (size_t) want_syms*b->channels*sizeof(span_b200_v29_symbol_t)) != 0)
            return -1;
        b->sym_cap = want_syms;
    }
    if (modem_configure<RX>() != 0)
        return -1;
    typedef typename RxKernel<RX>::type KRX;
    const int warps = RxKernel<RX>::WARPS;
    KernelArgs<KRX> ka = modem_args_for<KRX>(b, d_amp, stride, n);
    ka.a.append = (append)  ?  1  :  0;
    const int smem = (int) sizeof(float)*modem_smem_words<KRX>(warps);
    const int per_cta = warps*KRX::LS;                  // receivers per CTA
    modem_rx_kernel<KRX, RxKernel<RX>::WARPS><<<(b->channels + per_cta - 1)/per_cta, warps*32, smem, st>>>(ka);
    CK(cudaGetLastError());
    b->last_stream = st;
    b->have_last = true;
    return 0;
}

template <class RX>
static int modem_rx_host(ModemBank<RX> *b, const int16_t *h_amp, int64_t stride, int n, void *stream)
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
    const size_t want = sizeof(int16_t)*(size_t) n*b->channels + 16;
    if (b->d_in_bytes < want)
    {
        if (b->have_last)
            CK(cudaStreamSynchronize(b->last_stream));
        if (modem_realloc((void **) &b->d_in, want) != 0)
            return -1;
        b->d_in_bytes = want;
    }
    // A long call is fed in SBM_HOST_PIECES pieces of time: piece k + 1 crosses PCIe on the copy stream while the kernel
    // of piece k runs (a receiver kernel lasts as long as its sample count whatever the channel count, so the split is
    // over samples, not channels); the kernels of pieces 1.. continue the output of the one before.
    const int piece = ((n + SBM_HOST_PIECES - 1)/SBM_HOST_PIECES + 7) & ~7;
    if (n < 8000)
    {
        if (n > 0)
            CK(cudaMemcpy2DAsync(b->d_in, sizeof(int16_t)*(size_t) n, h_amp, sizeof(int16_t)*stride, sizeof(int16_t)*(size_t) n,
                                 b->channels, cudaMemcpyHostToDevice, st));
        return modem_rx_device(b, b->d_in, n, n, (void *) st);
    }
    if (b->copy_stream == NULL)
    {
        CK(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
        for (int i = 0;  i < SBM_HOST_PIECES;  i++)
            CK(cudaEventCreateWithFlags(&b->copied[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&b->in_free, cudaEventDisableTiming));
    }
    // the copies must not overtake what the caller's stream still does with d_in (the previous call's kernels), nor
    // start before the work already queued on the caller's stream (its ordering is the caller's contract)
    CK(cudaEventRecord(b->in_free, st));
    CK(cudaStreamWaitEvent(b->copy_stream, b->in_free, 0));
    int pieces = 0;
    for (int s0 = 0;  s0 < n;  s0 += piece, pieces++)
    {
        const int len = (n - s0 < piece)  ?  (n - s0)  :  piece;
        CK(cudaMemcpy2DAsync(b->d_in + s0, sizeof(int16_t)*(size_t) n, h_amp + s0, sizeof(int16_t)*stride, sizeof(int16_t)*(size_t) len,
                             b->channels, cudaMemcpyHostToDevice, b->copy_stream));
        CK(cudaEventRecord(b->copied[pieces], b->copy_stream));
    }
    pieces = 0;
    for (int s0 = 0;  s0 < n;  s0 += piece, pieces++)
    {
        const int len = (n - s0 < piece)  ?  (n - s0)  :  piece;
        CK(cudaStreamWaitEvent(st, b->copied[pieces], 0));
        if (modem_rx_device(b, b->d_in + s0, n, len, (void *) st, pieces > 0, n) != 0)
            return -1;
    }
    return 0;
}

template <class RX>
static int modem_counts(ModemBank<RX> *b, int32_t *nbits, int32_t *nsyms)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (modem_quiesce(b) != 0)
        return -1;
    if (nbits)
        CK(cudaMemcpy(nbits, b->nbits, sizeof(int)*(size_t) b->channels, cudaMemcpyDeviceToHost));
    if (nsyms)
        CK(cudaMemcpy(nsyms, b->nsyms, sizeof(int)*(size_t) b->channels, cudaMemcpyDeviceToHost));
    return 0;
}

// The put_bit sequence of one channel as the reference's callback would have seen it - 0 / 1 and the negative status
// values in their places - rebuilt from the packed data bits and the status list.  Returns the number of entries.
static inline long long modem_unpack(const unsigned int *words, long long words_cap, const int *status, long long status_cap,
                                     int nbits, int nstatus, int8_t *out, long long max)
{
    long long ns = nstatus;
    if (ns > status_cap)
        ns = status_cap;
    long long k = 0;            // position in the put_bit sequence
    long long d = 0;            // data bits consumed
    long long si = 0;
    const long long total = (nbits < max)  ?  nbits  :  max;
    while (k < total)
    {
        if (si < ns  &&  status[2*si] == k)
        {
            out[k++] = (int8_t) status[2*si + 1];
            si++;
            continue;
        }
        if ((d >> 5) >= words_cap)
            break;
        out[k++] = (int8_t) ((words[d >> 5] >> (d & 31)) & 1u);
        d++;
    }
    return k;
}

template <class RX>
static int64_t modem_bits(ModemBank<RX> *b, int channel, int8_t *out, int64_t max)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (channel < 0  ||  channel >= b->channels)
        return -1;
    if (modem_quiesce(b) != 0)
        return -1;
    int n = 0;
    int ns = 0;
    CK(cudaMemcpy(&n, b->nbits + channel, sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&ns, b->nstatus + channel, sizeof(int), cudaMemcpyDeviceToHost));
    if (n <= 0  ||  b->words == NULL)
        return 0;
    long long nw = ((long long) n + 31)/32;
    if (nw > b->words_cap)
        nw = b->words_cap;
    long long nst = (ns < b->status_cap)  ?  ns  :  b->status_cap;
    std::vector<unsigned int> w((size_t) nw + 1);
    std::vector<int> s2((size_t) 2*nst + 2);
    CK(cudaMemcpy(w.data(), b->words + (size_t) channel*b->words_cap, sizeof(unsigned int)*(size_t) nw, cudaMemcpyDeviceToHost));
    if (nst > 0)
        CK(cudaMemcpy(s2.data(), b->status + (size_t) channel*b->status_cap*2, sizeof(int)*2*(size_t) nst, cudaMemcpyDeviceToHost));
    return modem_unpack(w.data(), nw, s2.data(), nst, n, ns, out, max);
}

// Every channel's output in the packed form the kernel writes, one transfer per array (the bulk read-back):
//   words  [channels][words_stride]   the data bits, 32 to a word, first bit = bit 0 (row c holds channel c's)
//   status [channels][status_stride][2]  {position in the put_bit sequence, negative SIG_STATUS_* value}
//   nbits / nstatus [channels]        put_bit calls (bits + reports) / reports of each channel
// Rows are cut to the strides given.  Returns the largest nbits, or -1.
template <class RX>
static int64_t modem_output_packed(ModemBank<RX> *b, uint32_t *words, int64_t words_stride, int32_t *nbits,
                                   int32_t *status, int64_t status_stride, int32_t *nstatus)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b == NULL  ||  nbits == NULL  ||  nstatus == NULL  ||  words_stride < 0  ||  status_stride < 0)
    {
        sb_set_error("bad arguments");
        return -1;
    }
    if (modem_quiesce(b) != 0)
        return -1;
    const size_t C = (size_t) b->channels;
    CK(cudaMemcpy(nbits, b->nbits, sizeof(int)*C, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(nstatus, b->nstatus, sizeof(int)*C, cudaMemcpyDeviceToHost));
    long long mx = 0;
    long long ms = 0;
    for (size_t c = 0;  c < C;  c++)
    {
        if (nbits[c] > mx)
            mx = nbits[c];
        if (nstatus[c] > ms)
            ms = nstatus[c];
    }
    long long w = (mx + 31)/32;
    if (w > b->words_cap)
        w = b->words_cap;
    if (w > words_stride)
        w = words_stride;
    if (words  &&  w > 0)
        CK(cudaMemcpy2D(words, sizeof(uint32_t)*(size_t) words_stride, b->words, sizeof(unsigned int)*(size_t) b->words_cap,
                        sizeof(uint32_t)*(size_t) w, C, cudaMemcpyDeviceToHost));
    if (ms > b->status_cap)
        ms = b->status_cap;
    if (ms > status_stride)
        ms = status_stride;
    if (status  &&  ms > 0)
        CK(cudaMemcpy2D(status, sizeof(int32_t)*2*(size_t) status_stride, b->status, sizeof(int)*2*(size_t) b->status_cap,
                        sizeof(int32_t)*2*(size_t) ms, C, cudaMemcpyDeviceToHost));
    return mx;
}

// The put_bit sequences of all channels as bytes: out[c*out_stride ..] receives the first min(count, out_stride) entries
// of channel c; nbits (may be NULL) the per-channel counts.  A convenience over modem_output_packed (the unpacking runs
// on the host).  Returns the largest count.
template <class RX>
static int64_t modem_bits_all(ModemBank<RX> *b, int8_t *out, int64_t out_stride, int32_t *nbits)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (b == NULL  ||  out == NULL  ||  out_stride < 0)
    {
        sb_set_error("bad arguments");
        return -1;
    }
    const size_t C = (size_t) b->channels;
    std::vector<int32_t> n(C);
    std::vector<int32_t> ns(C);
    const long long wcap = (b->words_cap > 0)  ?  b->words_cap  :  1;
    const long long scap = (b->status_cap > 0)  ?  b->status_cap  :  1;
    std::vector<uint32_t> w(C*(size_t) wcap);
    std::vector<int32_t> st(C*(size_t) scap*2);
    const int64_t mx = modem_output_packed(b, w.data(), wcap, n.data(), st.data(), scap, ns.data());
    if (mx < 0)
        return -1;
    for (size_t c = 0;  c < C;  c++)
    {
        if (nbits)
            nbits[c] = n[c];
        modem_unpack(&w[c*(size_t) wcap], wcap, &st[c*(size_t) scap*2], scap, n[c], ns[c], out + c*(size_t) out_stride, out_stride);
    }
    return mx;
}

template <class RX>
static int64_t modem_symbols(ModemBank<RX> *b, int channel, span_b200_v29_symbol_t *out, int64_t max)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (channel < 0  ||  channel >= b->channels  ||  !b->want_symbols)
        return -1;
    if (modem_quiesce(b) != 0)
        return -1;
    int n = 0;
    CK(cudaMemcpy(&n, b->nsyms + channel, sizeof(int), cudaMemcpyDeviceToHost));
    long long k = n;
    if (k > b->sym_cap)
        k = b->sym_cap;
    if (k > max)
        k = max;
    if (k > 0)
        CK(cudaMemcpy(out, b->syms + (size_t) channel*b->sym_cap, sizeof(span_b200_v29_symbol_t)*(size_t) k, cudaMemcpyDeviceToHost));
    return k;
}

template <class RX>
static int modem_output_layout(ModemBank<RX> *b, const uint32_t **d_words, int64_t *words_cap, const int32_t **d_nbits,
                               const int32_t **d_status, int64_t *status_cap, const int32_t **d_nstatus,
                               const span_b200_v29_symbol_t **d_syms, int64_t *sym_cap, const int32_t **d_nsyms)
{
    if (d_words)
        *d_words = (const uint32_t *) b->words;
    if (words_cap)
        *words_cap = b->words_cap;
    if (d_nbits)
        *d_nbits = b->nbits;
    if (d_status)
        *d_status = b->status;
    if (status_cap)
        *status_cap = b->status_cap;
    if (d_nstatus)
        *d_nstatus = b->nstatus;
    if (d_syms)
        *d_syms = b->syms;
    if (sym_cap)
        *sym_cap = b->sym_cap;
    if (d_nsyms)
        *d_nsyms = b->nsyms;
    return 0;
}

// eq_coeff: eq_len complex taps; info[i] = istate field fields[i] (or, for fields[i] < 0, the float field
// -1 - fields[i] as its bit pattern).
template <class RX>
static int modem_channel_state(ModemBank<RX> *b, int channel, float *eq_coeff, int32_t *info, const int *fields, int nfields,
                               int eq_len = SBM_EQ_LEN)
{
    if (channel < 0  ||  channel >= b->channels)
        return -1;
    if (modem_quiesce(b) != 0)
        return -1;
    const size_t C = b->channels;
    if (eq_coeff)
    {
        for (int i = 0;  i < 2*eq_len;  i++)
            CK(cudaMemcpy(&eq_coeff[i], b->fstate + (size_t) (F_EQ_COEFF + i)*C + channel, sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (info)
    {
        for (int i = 0;  i < nfields;  i++)
        {
            if (fields[i] >= 0)
                CK(cudaMemcpy(&info[i], b->istate + (size_t) fields[i]*C + channel, sizeof(int), cudaMemcpyDeviceToHost));
            else
                CK(cudaMemcpy(&info[i], b->fstate + (size_t) (-1 - fields[i])*C + channel, sizeof(float), cudaMemcpyDeviceToHost));
        }
    }
    return 0;
}

}  // namespace sbm
