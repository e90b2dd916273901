// sb_engine.cu - host side of the B200 tone-bank engine: contexts, banks, launches, events.
// C ABI declared in include/spandsp_b200.h.  There is no CPU fallback anywhere in this file:
// without a usable sm_100 device every entry point fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <dlfcn.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <ctype.h>
#include <nccl.h>             // types and prototypes only: libnccl.so.2 is resolved at run time (dlopen), never linked

#include <algorithm>
#include <vector>
#include <mutex>
#include <map>

#include "sb_engine.h"
#include "sb_detectors.cuh"

using namespace sb;

// ------------------------------------------------------------------------------------------
// errors
static thread_local char g_err[512] = "";

void sb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

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

// Inside a *_create function, once the object exists: a failure releases what was already acquired
// (the destroy functions tolerate NULL members)
#define CKB(call) \
    do \
    { \
        cudaError_t e_ = (call); \
        if (e_ != cudaSuccess) \
        { \
            sb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            span_b200_bank_destroy(b); \
            return NULL; \
        } \
    } \
    while (0)

#define CKC(call) \
    do \
    { \
        cudaError_t e_ = (call); \
        if (e_ != cudaSuccess) \
        { \
            sb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            span_b200_ctx_destroy(ctx); \
            return NULL; \
        } \
    } \
    while (0)

extern "C" const char *span_b200_last_error(void)
{
    return g_err;
}

extern "C" int span_b200_abi_version(void)
{
    return SPAN_B200_ABI_VERSION;
}

// ------------------------------------------------------------------------------------------
// coefficient and level tables (host libm, same expressions as the reference)

// src/tone_detect.c:60-68: 2.0f*cosf(2.0f*M_PI*(freq/8000.0f)); the cosf argument is formed in
// double because M_PI is a double constant, then narrowed by the call.
float sb_goertzel_fac(float freq)
{
    return 2.0f*cosf((float) (2.0f*M_PI*(freq/8000.0f)));
}

// G.711 expansion (src/spandsp/g711.h:165-172 u-law with bias 0x84, :239-252 A-law with AMI mask 0x55)
int sb_ulaw_to_linear(unsigned char ulaw)
{
    ulaw = (unsigned char) ~ulaw;
    const int t = (((ulaw & 0x0F) << 3) + 0x84) << (((int) ulaw & 0x70) >> 4);
    return (short) ((ulaw & 0x80)  ?  (0x84 - t)  :  (t - 0x84));
}

int sb_alaw_to_linear(unsigned char alaw)
{
    alaw ^= 0x55;
    int i = ((alaw & 0x0F) << 4);
    const int seg = (((int) alaw & 0x70) >> 4);
    if (seg)
        i = (i + 0x108) << (seg - 1);
    else
        i += 8;
    return (short) ((alaw & 0x80)  ?  i  :  -i);
}

// src/dtmf.c:110,314 with lfastrintf = C truncation on x86-64 (src/spandsp/fast_convert.h:194-197)
static int host_dtmf_level(float energy)
{
    return (int) (long int) (10.0f*log10f(energy) - 107.255f);
}

static float bits_to_float(uint32_t u)
{
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// Tabulate the steps of host_dtmf_level over all positive finite floats (monotone in the bit
// pattern).  thr[k] = smallest float whose level is >= level_min + k + 1.
static void build_level_table(std::vector<float> &thr, int &level_min)
{
    const uint32_t lo_bits = 1;                 // smallest denormal
    const uint32_t hi_bits = 0x7F7FFFFFu;       // FLT_MAX
    level_min = host_dtmf_level(bits_to_float(lo_bits));
    const int level_max = host_dtmf_level(bits_to_float(hi_bits));
    thr.clear();
    uint32_t start = lo_bits;
    for (int target = level_min + 1;  target <= level_max;  target++)
    {
        uint32_t a = start;                     // level(a) < target
        uint32_t b = hi_bits;                   // level(b) >= target
        while (b - a > 1)
        {
            const uint32_t m = a + (b - a)/2;
            if (host_dtmf_level(bits_to_float(m)) >= target)
                b = m;
            else
                a = m;
        }
        thr.push_back(bits_to_float(b));
        start = a;
    }
}

// ------------------------------------------------------------------------------------------
// context
struct span_b200_ctx_s
{
    int device;
    int sm_count;
    int smem_optin;
    cudaStream_t stream;
    float *d_level_thr;
    int level_n;
    int level_min;
    float *d_g711_lut;              // [2][256]: u-law, A-law expansion as float
    int numa_node;                  // NUMA node the GPU hangs off (-1: unknown / single node)
    std::mutex host_lock;
    std::map<void *, size_t> host_blocks;   // span_b200_host_alloc() blocks
};

static int upload_constants(int device)
{
    float f[8];
    static const float row[4] = {697.0f, 770.0f, 852.0f, 941.0f};           // src/dtmf.c:114-117
    static const float col[4] = {1209.0f, 1336.0f, 1477.0f, 1633.0f};       // src/dtmf.c:118-121
    static const int bell[6] = {700, 900, 1100, 1300, 1500, 1700};          // src/bell_r2_mf.c:251-254
    static const int r2f[6] = {1380, 1500, 1620, 1740, 1860, 1980};         // src/bell_r2_mf.c:264-267
    static const int r2b[6] = {1140, 1020, 900, 780, 660, 540};             // src/bell_r2_mf.c:269-272
    (void) device;
    for (int i = 0;  i < 4;  i++)
    {
        f[2*i] = sb_goertzel_fac(row[i]);
        f[2*i + 1] = sb_goertzel_fac(col[i]);
    }
    CK(cudaMemcpyToSymbol(c_dtmf_fac, f, sizeof(float)*8));
    CK(cudaMemcpyToSymbol(c_dtmf_positions, "123A456B789C*0#D", 17));
    for (int i = 0;  i < 6;  i++)
        f[i] = sb_goertzel_fac((float) bell[i]);
    CK(cudaMemcpyToSymbol(c_bell_mf_fac, f, sizeof(float)*6));
    for (int i = 0;  i < 6;  i++)
        f[i] = sb_goertzel_fac((float) r2f[i]);
    CK(cudaMemcpyToSymbol(c_r2_fwd_fac, f, sizeof(float)*6));
    for (int i = 0;  i < 6;  i++)
        f[i] = sb_goertzel_fac((float) r2b[i]);
    CK(cudaMemcpyToSymbol(c_r2_back_fac, f, sizeof(float)*6));
    CK(cudaMemcpyToSymbol(c_bell_mf_positions, "1247C-358A--69*---0B----#", 26));      // src/bell_r2_mf.c:262
    CK(cudaMemcpyToSymbol(c_r2_mf_positions, "1247B-358C--69D---0E----F", 26));        // src/bell_r2_mf.c:276
    return 0;
}

static int gpu_numa_node(int device);

extern "C" span_b200_ctx_t *span_b200_ctx_create(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess  ||  count == 0)
    {
        sb_set_error("no CUDA device available (%s); spandsp_b200 has no CPU fallback",
                     (e != cudaSuccess)  ?  cudaGetErrorString(e)  :  "device count is 0");
        return NULL;
    }
    if (device < 0)
        CKP(cudaGetDevice(&device));
    if (device >= count)
    {
        sb_set_error("device %d out of range (%d devices)", device, count);
        return NULL;
    }
    cudaDeviceProp prop;
    CKP(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
    {
        sb_set_error("device %d is sm_%d%d; spandsp_b200 kernels are built for sm_100a only", device, prop.major, prop.minor);
        return NULL;
    }
    SB_DEVICE_CKP(device);
    span_b200_ctx_t *ctx = new span_b200_ctx_s();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int) prop.sharedMemPerBlockOptin;
    ctx->numa_node = gpu_numa_node(device);
    CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    if (upload_constants(device) != 0)
    {
        span_b200_ctx_destroy(ctx);
        return NULL;
    }
    std::vector<float> thr;
    build_level_table(thr, ctx->level_min);
    ctx->level_n = (int) thr.size();
    CKC(cudaMalloc(&ctx->d_level_thr, sizeof(float)*thr.size()));
    CKC(cudaMemcpy(ctx->d_level_thr, thr.data(), sizeof(float)*thr.size(), cudaMemcpyHostToDevice));
    float lut[512];
    for (int i = 0;  i < 256;  i++)
    {
        lut[i] = (float) sb_ulaw_to_linear((unsigned char) i);
        lut[256 + i] = (float) sb_alaw_to_linear((unsigned char) i);
    }
    CKC(cudaMalloc(&ctx->d_g711_lut, sizeof(lut)));
    CKC(cudaMemcpy(ctx->d_g711_lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
    return ctx;
}

extern "C" void span_b200_ctx_destroy(span_b200_ctx_t *ctx)
{
    if (ctx == NULL)
        return;
    sb_device_guard sb_dg_(ctx->device);
    if (ctx->stream)
        cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_level_thr);
    cudaFree(ctx->d_g711_lut);
    if (ctx->stream)
        cudaStreamDestroy(ctx->stream);
    for (std::map<void *, size_t>::iterator it = ctx->host_blocks.begin();  it != ctx->host_blocks.end();  ++it)
    {
        cudaHostUnregister(it->first);
        munmap(it->first, it->second);
    }
    delete ctx;
}

extern "C" int span_b200_ctx_device(const span_b200_ctx_t *ctx)
{
    return (ctx)  ?  ctx->device  :  -1;
}

// ---- NUMA-local pinned host memory --------------------------------------------------------------
// The end-to-end rate of the host interface is the PCIe rate, and with several GPUs per box what limits that is where
// the staging memory lives: a buffer on the other socket crosses the inter-socket link on every DMA.  A block from
// span_b200_host_alloc() is placed on the NUMA node of the context's GPU (mbind, preferred policy: it degrades to
// any node where the container's cpuset does not allow that one) and pinned (cudaHostRegister).
static int gpu_numa_node(int device)
{
    char id[64];
    if (cudaDeviceGetPCIBusId(id, sizeof(id), device) != cudaSuccess)
        return -1;
    for (char *p = id;  *p;  p++)
        *p = (char) tolower(*p);
    char path[160];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", id);
    FILE *f = fopen(path, "r");
    if (f == NULL)
        return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1)
        node = -1;
    fclose(f);
    return node;
}

extern "C" int span_b200_ctx_numa_node(const span_b200_ctx_t *ctx)
{
    return ctx->numa_node;
}

extern "C" void *span_b200_host_alloc(span_b200_ctx_t *ctx, size_t bytes)
{
    if (ctx == NULL  ||  bytes == 0)
    {
        sb_set_error("bad host allocation arguments");
        return NULL;
    }
    SB_DEVICE_CKP(ctx->device);
    const size_t page = 2u << 20;
    const size_t len = (bytes + page - 1)/page*page;
    void *p = mmap(NULL, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED)
    {
        sb_set_error("mmap of %zu bytes failed", len);
        return NULL;
    }
    if (ctx->numa_node >= 0  &&  ctx->numa_node < 1024)
    {
        unsigned long mask[16];
        memset(mask, 0, sizeof(mask));
        mask[ctx->numa_node/(8*sizeof(unsigned long))] |= 1ul << (ctx->numa_node%(8*sizeof(unsigned long)));
        // MPOL_PREFERRED = 1: pages come from this node while it has room and the cpuset allows it
        (void) syscall(SYS_mbind, p, len, 1, mask, (unsigned long) (8*sizeof(mask)), 0u);
    }
    cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterDefault);
    if (e != cudaSuccess)
    {
        sb_set_error("cudaHostRegister of %zu bytes failed: %s", len, cudaGetErrorString(e));
        munmap(p, len);
        return NULL;
    }
    std::lock_guard<std::mutex> lk(ctx->host_lock);
    ctx->host_blocks[p] = len;
    return p;
}

extern "C" void span_b200_host_free(span_b200_ctx_t *ctx, void *p)
{
    if (ctx == NULL  ||  p == NULL)
        return;
    sb_device_guard sb_dg_(ctx->device);
    size_t len = 0;
    {
        std::lock_guard<std::mutex> lk(ctx->host_lock);
        std::map<void *, size_t>::iterator it = ctx->host_blocks.find(p);
        if (it == ctx->host_blocks.end())
            return;
        len = it->second;
        ctx->host_blocks.erase(it);
    }
    cudaHostUnregister(p);
    munmap(p, len);
}

extern "C" int span_b200_ctx_sm_count(const span_b200_ctx_t *ctx)
{
    return ctx->sm_count;
}

void *sb_ctx_stream(span_b200_ctx_t *ctx)
{
    return (void *) ctx->stream;
}

// ------------------------------------------------------------------------------------------
// communicator (multi-GPU gather of the event records); functions further down
#define SB_PULL_STREAMS 4

struct span_b200_comm_s
{
    span_b200_ctx_t *ctx;
    int nranks;
    int rank;
    ncclComm_t nccl;
    cudaStream_t stream;
    // root: the gathered records of the two buffers in flight
    span_b200_wire_event_t *gather[2];
    long long gather_cap[2];
    unsigned long long *d_counts[2];        // [nranks][SB_META_WORDS] all-gathered {record count, buffer generation, IPC handle}
    unsigned long long *h_counts[2];        // pinned
    unsigned long long *d_sync;             // [1 + nranks] completion collective of the peer-copy transport
    int transport;                          // SPAN_B200_GATHER_*
    void **peer_ptr[2];                     // [nranks] the ranks' record buffers as mapped here (root, peer-copy transport)
    unsigned long long *peer_gen[2];        // [nranks] generation of the mapping
    cudaEvent_t counts_ready[2];
    cudaEvent_t done[2];                    // the transfer that read / filled buffer i has finished
    // root, peer-copy transport: the ranks' records are pulled on several streams at once (one copy engine pulling
    // over NVLink is latency-limited well below the link rate), forked from / joined to `stream` by events
    cudaStream_t pull[SB_PULL_STREAMS];
    cudaEvent_t fork;
    cudaEvent_t joined[SB_PULL_STREAMS];
    bool done_valid[2];
    int begun_slot;                         // slot of the last _gather_begin without _gather_end, or -1
    int ended_slot;                         // slot of the last _gather_end, or -1
    long long total[2];
};

static int comm_gather_reserve(span_b200_comm_t *cm, int slot, long long records, cudaStream_t st, long long keep);

// What a rank contributes to the counts exchange, per record buffer: [0] = records of the call (written by the scan
// kernel), [1] = generation of the buffer (it changes when the buffer is re-allocated), [2..9] = its cudaIpcMemHandle_t
#define SB_META_WORDS   10

// ------------------------------------------------------------------------------------------
// bank
struct span_b200_bank_s
{
    span_b200_ctx_t *ctx;
    int det;
    int channels;
    int block;
    int bins;
    int npairs;
    int fwd;
    int want_segments;

    // carried bank state
    float *v2;
    float *v3;
    float *energy;
    int *cs;
    // DTMF
    float *thr;
    float *ntw;
    float *rtw;
    unsigned char *flags;
    float *z;
    unsigned char *last_hit;
    unsigned char *in_digit;
    int *duration;
    std::vector<float> h_thr;
    std::vector<float> h_ntw;
    std::vector<float> h_rtw;
    std::vector<unsigned char> h_flags;
    int n_filter;
    // MF
    unsigned char *hits;
    // super-tone
    StParams stp;
    int *segments;
    int *detected;
    int *rotation;
    unsigned char *pending;
    void *st_log;                       // [groups][SB_ST_LOG] records the super-tone count pass keeps for the emit pass
    int *d_tone_segs;
    int *d_tone_first;
    int4 *d_elements;
    int tones;
    int total_elements;
    std::vector<float> h_fac;

    // per-call scratch
    void *code;
    size_t code_bytes;
    float *eout;
    size_t eout_bytes;
    unsigned int *counts;
    unsigned int *offsets;
    unsigned long long *d_total;            // [2][SB_META_WORDS]: per record buffer (wire mode alternates; else slot 0), see SB_META_WORDS
    unsigned long long *h_total;            // [2][SB_META_WORDS], pinned: [0] comes back from the device, [1..] go up
    unsigned long long wire_gen;
    span_b200_event_t *events;
    long long ev_cap;
    long long ev_cap_user;
    int16_t *d_in;
    size_t d_in_bytes;
    cudaStream_t copy_stream;       // rx_host: the H2D copies run beside the filter-bank kernels
    cudaEvent_t copied[8];
    cudaEvent_t in_free;

    // wire mode (12-byte records, two buffers used alternately) and the multi-GPU gather
    int wire_on;
    unsigned int channel_base;
    span_b200_wire_event_t *wire[2];
    long long wire_cap[2];
    int slot;                       // buffer of the last rx call
    cudaEvent_t emitted[2];         // the emit pass that filled buffer i has finished
    span_b200_comm_t *comm;
    int root;

    int uniform_cs;                 // common block phase of all channels, -1 if they differ
    int last_nb;
    int last_launches;
    const char *last_path;
    cudaStream_t last_stream;
    bool have_last;

    int tune_timing;
    std::vector<cudaEvent_t> ev_pool;
    int ev_used;

    int tune_slice;
    int tune_variant;
    int tune_direct;
    int tune_packed;
};

static int ensure(void **p, size_t *have, size_t want)
{
    if (*have >= want)
        return 0;
    if (*p)
        CK(cudaFree(*p));
    *p = NULL;
    *have = 0;
    size_t sz = want + want/8 + 256;
    CK(cudaMalloc(p, sz));
    *have = sz;
    return 0;
}

// The super-tone bank kernel is instantiated for a fixed list of pair counts; a descriptor runs on the next larger
// one (the surplus bins have a zero coefficient and are never read by the decision).  The carried state arrays
// must have the rows of the instantiation that runs, not of the descriptor: load_carry()/store_carry() touch
// all 2*NPAIRS rows.
static int st_template_pairs(int npairs)
{
    static const int sizes[] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 24, 32};
    for (size_t i = 0;  i < sizeof(sizes)/sizeof(sizes[0]);  i++)
    {
        if (npairs <= sizes[i])
            return sizes[i];
    }
    return -1;
}

static span_b200_bank_t *bank_alloc(span_b200_ctx_t *ctx, int det, int channels, int block, int bins)
{
    if (ctx == NULL  ||  channels <= 0)
    {
        sb_set_error("bad bank arguments");
        return NULL;
    }
    SB_DEVICE_CKP(ctx->device);
    span_b200_bank_t *b = new span_b200_bank_s();
    b->ctx = ctx;
    b->det = det;
    b->channels = channels;
    b->block = block;
    b->bins = bins;
    b->npairs = (bins + 1)/2;
    if (det == SPAN_B200_DET_SUPER_TONE)
        b->npairs = st_template_pairs(b->npairs);
    b->uniform_cs = 0;
    b->tune_packed = 5;
    b->last_path = "";
    const size_t C = channels;
    CKB(cudaMalloc(&b->v2, sizeof(float)*2*b->npairs*C));
    CKB(cudaMalloc(&b->v3, sizeof(float)*2*b->npairs*C));
    CKB(cudaMalloc(&b->energy, sizeof(float)*C));
    CKB(cudaMalloc(&b->cs, sizeof(int)*C));
    // one count / offset per group of channels that share a warp in the sequencer (32; the super-tone sequencer: SB_ST_CPW)
    CKB(cudaMalloc(&b->counts, sizeof(unsigned int)*((C + SB_ST_CPW - 1)/SB_ST_CPW + 1)));
    CKB(cudaMalloc(&b->offsets, sizeof(unsigned int)*((C + SB_ST_CPW - 1)/SB_ST_CPW + 1)));
    CKB(cudaMalloc(&b->d_total, 2*SB_META_WORDS*sizeof(unsigned long long)));
    CKB(cudaMemset(b->d_total, 0, 2*SB_META_WORDS*sizeof(unsigned long long)));
    CKB(cudaMallocHost(&b->h_total, 2*SB_META_WORDS*sizeof(unsigned long long)));
    memset(b->h_total, 0, 2*SB_META_WORDS*sizeof(unsigned long long));
    return b;
}

extern "C" int span_b200_bank_reset(span_b200_bank_t *b, int first, int count)
{
    if (first < 0  ||  count < 0  ||  first + count > b->channels)
    {
        sb_set_error("channel range out of bounds");
        return -1;
    }
    if (count == 0)
        return 0;
    SB_DEVICE_CK(b->ctx->device);
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    const size_t C = b->channels;
    for (int i = 0;  i < 2*b->npairs;  i++)
    {
        CK(cudaMemset(b->v2 + i*C + first, 0, sizeof(float)*count));
        CK(cudaMemset(b->v3 + i*C + first, 0, sizeof(float)*count));
    }
    CK(cudaMemset(b->energy + first, 0, sizeof(float)*count));
    CK(cudaMemset(b->cs + first, 0, sizeof(int)*count));
    switch (b->det)
    {
    case SPAN_B200_DET_DTMF:
        for (int i = first;  i < first + count;  i++)
        {
            b->h_thr[i] = 171029200.0f;         // src/dtmf.c:104
            b->h_ntw[i] = 6.309f;               // src/dtmf.c:105
            b->h_rtw[i] = 2.512f;               // src/dtmf.c:106
            if (b->h_flags[i] & SB_DTMF_FLAG_FILTER)
                b->n_filter--;
            b->h_flags[i] = 0;
        }
        CK(cudaMemcpy(b->thr + first, &b->h_thr[first], sizeof(float)*count, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(b->ntw + first, &b->h_ntw[first], sizeof(float)*count, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(b->rtw + first, &b->h_rtw[first], sizeof(float)*count, cudaMemcpyHostToDevice));
        CK(cudaMemset(b->flags + first, 0, count));
        for (int i = 0;  i < 4;  i++)
            CK(cudaMemset(b->z + i*C + first, 0, sizeof(float)*count));
        CK(cudaMemset(b->last_hit + first, 0, count));
        CK(cudaMemset(b->in_digit + first, 0, count));
        CK(cudaMemset(b->duration + first, 0, sizeof(int)*count));
        break;
    case SPAN_B200_DET_BELL_MF:
        for (int i = 0;  i < 5;  i++)
            CK(cudaMemset(b->hits + i*C + first, 0, count));
        break;
    case SPAN_B200_DET_R2_MF:
        CK(cudaMemset(b->hits + first, 0, count));
        break;
    case SPAN_B200_DET_SUPER_TONE:
        for (int i = 0;  i < 11;  i++)
        {
            // src/super_tone_rx.c:527-533: f1 = f2 = -1, min_duration = 0
            CK(cudaMemset(b->segments + (3*i)*C + first, 0xFF, sizeof(int)*count));
            CK(cudaMemset(b->segments + (3*i + 1)*C + first, 0xFF, sizeof(int)*count));
            CK(cudaMemset(b->segments + (3*i + 2)*C + first, 0, sizeof(int)*count));
        }
        CK(cudaMemset(b->detected + first, 0xFF, sizeof(int)*count));
        CK(cudaMemset(b->rotation + first, 0, sizeof(int)*count));
        CK(cudaMemset(b->pending + first, 0, count));
        break;
    }
    // Block phases: all zero again only if the whole bank was reset or already uniform at zero.
    if (!(count == b->channels  ||  b->uniform_cs == 0))
        b->uniform_cs = -1;
    else
        b->uniform_cs = 0;
    return 0;
}

extern "C" span_b200_bank_t *span_b200_dtmf_bank_create(span_b200_ctx_t *ctx, int channels)
{
    sb_device_guard sb_dg_(span_b200_ctx_device(ctx));
    span_b200_bank_t *b = bank_alloc(ctx, SPAN_B200_DET_DTMF, channels, 102, 8);
    if (b == NULL)
        return NULL;
    const size_t C = channels;
    CKB(cudaMalloc(&b->thr, sizeof(float)*C));
    CKB(cudaMalloc(&b->ntw, sizeof(float)*C));
    CKB(cudaMalloc(&b->rtw, sizeof(float)*C));
    CKB(cudaMalloc(&b->flags, C));
    CKB(cudaMalloc(&b->z, sizeof(float)*4*C));
    CKB(cudaMalloc(&b->last_hit, C));
    CKB(cudaMalloc(&b->in_digit, C));
    CKB(cudaMalloc(&b->duration, sizeof(int)*C));
    b->h_thr.assign(C, 0.0f);
    b->h_ntw.assign(C, 0.0f);
    b->h_rtw.assign(C, 0.0f);
    b->h_flags.assign(C, 0);
    b->n_filter = 0;
    if (span_b200_bank_reset(b, 0, channels) != 0)
    {
        span_b200_bank_destroy(b);
        return NULL;
    }
    return b;
}

extern "C" span_b200_bank_t *span_b200_bell_mf_bank_create(span_b200_ctx_t *ctx, int channels)
{
    sb_device_guard sb_dg_(span_b200_ctx_device(ctx));
    span_b200_bank_t *b = bank_alloc(ctx, SPAN_B200_DET_BELL_MF, channels, 120, 6);
    if (b == NULL)
        return NULL;
    CKB(cudaMalloc(&b->hits, (size_t) 5*channels));
    if (span_b200_bank_reset(b, 0, channels) != 0)
    {
        span_b200_bank_destroy(b);
        return NULL;
    }
    return b;
}

extern "C" span_b200_bank_t *span_b200_r2_mf_bank_create(span_b200_ctx_t *ctx, int channels, int fwd)
{
    sb_device_guard sb_dg_(span_b200_ctx_device(ctx));
    span_b200_bank_t *b = bank_alloc(ctx, SPAN_B200_DET_R2_MF, channels, 133, 6);
    if (b == NULL)
        return NULL;
    b->fwd = (fwd != 0);
    CKB(cudaMalloc(&b->hits, (size_t) channels));
    if (span_b200_bank_reset(b, 0, channels) != 0)
    {
        span_b200_bank_destroy(b);
        return NULL;
    }
    return b;
}

// Host-side descriptor construction: src/super_tone_rx.c:81-161.
struct st_build_t
{
    int used;
    int monitored;
    int pitches[64][2];
    float fac[64];
};

static int st_add_freq(st_build_t *d, int freq)
{
    int i;

    if (freq == 0)
        return -1;
    for (i = 0;  i < d->used;  i++)
    {
        if (d->pitches[i][0] == freq)
            return d->pitches[i][1];
    }
    for (i = 0;  i < d->used;  i++)
    {
        if ((d->pitches[i][0] - 10) <= freq  &&  freq <= (d->pitches[i][0] + 10))
        {
            // Close to a tone we already monitor: share its detector, retuned to the mean.
            if (d->used >= 64)
                return -2;
            d->pitches[d->used][0] = freq;
            d->pitches[d->used][1] = i;
            d->fac[d->pitches[i][1]] = sb_goertzel_fac((float) (freq + d->pitches[i][0])/2);
            d->used++;
            return d->pitches[i][1];
        }
    }
    if (d->used >= 64)
        return -2;
    d->pitches[i][0] = freq;
    d->pitches[i][1] = d->monitored;
    d->fac[d->monitored++] = sb_goertzel_fac((float) freq);
    d->used++;
    return d->pitches[i][1];
}

extern "C" span_b200_bank_t *span_b200_super_tone_bank_create(span_b200_ctx_t *ctx, int channels,
                                                              const span_b200_super_tone_desc_t *desc,
                                                              int want_segments)
{
    if (desc == NULL  ||  desc->tones < 0)
    {
        sb_set_error("super-tone descriptor missing");
        return NULL;
    }
    st_build_t bd;
    memset(&bd, 0, sizeof(bd));
    std::vector<int> tone_first;
    std::vector<int4> elements;
    int k = 0;
    for (int t = 0;  t < desc->tones;  t++)
    {
        tone_first.push_back(k);
        for (int e = 0;  e < desc->tone_segs[t];  e++, k++)
        {
            int4 el;
            el.x = st_add_freq(&bd, desc->elements[4*k + 0]);
            el.y = st_add_freq(&bd, desc->elements[4*k + 1]);
            if (el.x == -2  ||  el.y == -2)
            {
                sb_set_error("super-tone descriptor uses more than 64 distinct frequencies");
                return NULL;
            }
            el.z = desc->elements[4*k + 2]*8;                                           // src/super_tone_rx.c:157
            el.w = (desc->elements[4*k + 3] == 0)  ?  0x7FFFFFFF  :  desc->elements[4*k + 3]*8;   // :158
            elements.push_back(el);
        }
    }
    if (bd.monitored < 1)
    {
        sb_set_error("super-tone descriptor monitors no frequency");
        return NULL;
    }
    if (bd.monitored > 2*SB_ST_MAX_PAIRS)
    {
        sb_set_error("super-tone descriptor monitors %d frequencies; the limit is %d (src/spandsp/private/super_tone_rx.h:44)",
                     bd.monitored, 2*SB_ST_MAX_PAIRS);
        return NULL;
    }
    span_b200_bank_t *b = bank_alloc(ctx, SPAN_B200_DET_SUPER_TONE, channels, 128, bd.monitored);
    if (b == NULL)
        return NULL;
    b->want_segments = (want_segments != 0);
    memset(&b->stp, 0, sizeof(b->stp));
    b->stp.bins = bd.monitored;
    for (int i = 0;  i < bd.monitored;  i++)
        b->stp.fac[i] = bd.fac[i];
    b->h_fac.assign(bd.fac, bd.fac + bd.monitored);
    b->tones = desc->tones;
    b->total_elements = (int) elements.size();
    const size_t C = channels;
    CKB(cudaMalloc(&b->segments, sizeof(int)*33*C));
    CKB(cudaMalloc(&b->detected, sizeof(int)*C));
    CKB(cudaMalloc(&b->rotation, sizeof(int)*C));
    CKB(cudaMalloc(&b->pending, C));
    CKB(cudaMalloc(&b->st_log, (size_t) ((C + SB_ST_CPW - 1)/SB_ST_CPW)*SB_ST_LOG*sizeof(span_b200_event_t)));
    CKB(cudaMalloc(&b->d_tone_segs, sizeof(int)*(desc->tones + 1)));
    CKB(cudaMalloc(&b->d_tone_first, sizeof(int)*(desc->tones + 1)));
    CKB(cudaMalloc(&b->d_elements, sizeof(int4)*(elements.size() + 1)));
    if (desc->tones)
    {
        CKB(cudaMemcpy(b->d_tone_segs, desc->tone_segs, sizeof(int)*desc->tones, cudaMemcpyHostToDevice));
        CKB(cudaMemcpy(b->d_tone_first, tone_first.data(), sizeof(int)*desc->tones, cudaMemcpyHostToDevice));
    }
    if (!elements.empty())
        CKB(cudaMemcpy(b->d_elements, elements.data(), sizeof(int4)*elements.size(), cudaMemcpyHostToDevice));
    if (span_b200_bank_reset(b, 0, channels) != 0)
    {
        span_b200_bank_destroy(b);
        return NULL;
    }
    return b;
}

extern "C" void span_b200_bank_destroy(span_b200_bank_t *b)
{
    if (b == NULL)
        return;
    sb_device_guard sb_dg_(b->ctx->device);
    if (b->have_last)
        cudaStreamSynchronize(b->last_stream);
    cudaFree(b->v2);
    cudaFree(b->v3);
    cudaFree(b->energy);
    cudaFree(b->cs);
    cudaFree(b->thr);
    cudaFree(b->ntw);
    cudaFree(b->rtw);
    cudaFree(b->flags);
    cudaFree(b->z);
    cudaFree(b->last_hit);
    cudaFree(b->in_digit);
    cudaFree(b->duration);
    cudaFree(b->hits);
    cudaFree(b->segments);
    cudaFree(b->detected);
    cudaFree(b->rotation);
    cudaFree(b->pending);
    cudaFree(b->st_log);
    cudaFree(b->d_tone_segs);
    cudaFree(b->d_tone_first);
    cudaFree(b->d_elements);
    cudaFree(b->code);
    cudaFree(b->eout);
    cudaFree(b->counts);
    cudaFree(b->offsets);
    cudaFree(b->d_total);
    cudaFreeHost(b->h_total);
    if (b->copy_stream)
    {
        cudaStreamSynchronize(b->copy_stream);
        cudaStreamDestroy(b->copy_stream);
        for (int i = 0;  i < 8;  i++)
        {
            if (b->copied[i])
                cudaEventDestroy(b->copied[i]);
        }
        if (b->in_free)
            cudaEventDestroy(b->in_free);
    }
    for (int i = 0;  i < 2;  i++)
    {
        cudaFree(b->wire[i]);
        if (b->emitted[i])
            cudaEventDestroy(b->emitted[i]);
    }
    cudaFree(b->events);
    cudaFree(b->d_in);
    for (size_t i = 0;  i < b->ev_pool.size();  i++)
        cudaEventDestroy(b->ev_pool[i]);
    delete b;
}

extern "C" int span_b200_bank_channels(const span_b200_bank_t *b) { return b->channels; }
extern "C" int span_b200_bank_detector(const span_b200_bank_t *b) { return b->det; }
extern "C" int span_b200_bank_block_len(const span_b200_bank_t *b) { return b->block; }
extern "C" int span_b200_bank_bins(const span_b200_bank_t *b) { return b->bins; }
extern "C" const char *span_b200_bank_last_path(const span_b200_bank_t *b) { return b->last_path; }
extern "C" int span_b200_bank_last_launches(const span_b200_bank_t *b) { return b->last_launches; }
extern "C" int span_b200_bank_last_blocks(span_b200_bank_t *b) { return b->last_nb; }

extern "C" int span_b200_bank_coefficients(const span_b200_bank_t *b, float *fac, int max)
{
    static const float row[4] = {697.0f, 770.0f, 852.0f, 941.0f};
    static const float col[4] = {1209.0f, 1336.0f, 1477.0f, 1633.0f};
    static const int bell[6] = {700, 900, 1100, 1300, 1500, 1700};
    static const int r2f[6] = {1380, 1500, 1620, 1740, 1860, 1980};
    static const int r2b[6] = {1140, 1020, 900, 780, 660, 540};
    int n = 0;
    for (int i = 0;  i < b->bins  &&  i < max;  i++, n++)
    {
        switch (b->det)
        {
        case SPAN_B200_DET_DTMF:
            fac[i] = sb_goertzel_fac((i & 1)  ?  col[i >> 1]  :  row[i >> 1]);
            break;
        case SPAN_B200_DET_BELL_MF:
            fac[i] = sb_goertzel_fac((float) bell[i]);
            break;
        case SPAN_B200_DET_R2_MF:
            fac[i] = sb_goertzel_fac((float) ((b->fwd)  ?  r2f[i]  :  r2b[i]));
            break;
        default:
            fac[i] = b->h_fac[i];
            break;
        }
    }
    return n;
}

extern "C" int span_b200_bank_tune(span_b200_bank_t *b, int what, int value)
{
    switch (what)
    {
    case 0: b->tune_slice = value; return 0;
    case 1: b->tune_variant = value; return 0;
    case 2: b->tune_direct = value; return 0;
    case 3: b->tune_packed = value; return 0;
    case 4: b->tune_timing = value; return 0;
    }
    sb_set_error("unknown tuning knob %d", what);
    return -1;
}

extern "C" int span_b200_bank_set_event_capacity(span_b200_bank_t *b, int64_t events)
{
    b->ev_cap_user = events;
    return 0;
}

// ---- DTMF control plane ---------------------------------------------------------------------
static int range_ok(span_b200_bank_t *b, int det, int first, int count)
{
    if (b->det != det)
    {
        sb_set_error("wrong detector type for this call");
        return 0;
    }
    if (first < 0  ||  count < 0  ||  first + count > b->channels)
    {
        sb_set_error("channel range out of bounds");
        return 0;
    }
    return 1;
}

extern "C" int span_b200_dtmf_bank_parms(span_b200_bank_t *b, int first, int count,
                                         int filter_dialtone, float twist, float reverse_twist, float threshold)
{
    if (!range_ok(b, SPAN_B200_DET_DTMF, first, count))
        return -1;
    if (count == 0)
        return 0;
    SB_DEVICE_CK(b->ctx->device);
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    const size_t C = b->channels;
    // src/dtmf.c:421-445
    for (int i = first;  i < first + count;  i++)
    {
        if (filter_dialtone >= 0)
        {
            if (b->h_flags[i] & SB_DTMF_FLAG_FILTER)
                b->n_filter--;
            b->h_flags[i] &= ~SB_DTMF_FLAG_FILTER;
            if (filter_dialtone)
            {
                b->h_flags[i] |= SB_DTMF_FLAG_FILTER;
                b->n_filter++;
            }
        }
        if (twist >= 0.0f)
            b->h_ntw[i] = powf(10.0f, twist/10.0f);                     // db_to_power_ratio, telephony.h:141
        if (reverse_twist >= 0.0f)
            b->h_rtw[i] = powf(10.0f, reverse_twist/10.0f);
        if (threshold > -99.0f)                                          // goertzel_threshold_dbm0, tone_detect.h:66
            b->h_thr[i] = (float) ((102*102*32768.0f*32768.0f/2.0f)*powf(10.0f, (threshold - 3.14f)/10.0f));
    }
    if (filter_dialtone >= 0)
    {
        for (int i = 0;  i < 4;  i++)
            CK(cudaMemset(b->z + i*C + first, 0, sizeof(float)*count));
    }
    CK(cudaMemcpy(b->flags + first, &b->h_flags[first], count, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b->thr + first, &b->h_thr[first], sizeof(float)*count, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b->ntw + first, &b->h_ntw[first], sizeof(float)*count, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b->rtw + first, &b->h_rtw[first], sizeof(float)*count, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int span_b200_dtmf_bank_realtime(span_b200_bank_t *b, int first, int count, int on)
{
    if (!range_ok(b, SPAN_B200_DET_DTMF, first, count))
        return -1;
    if (count == 0)
        return 0;
    SB_DEVICE_CK(b->ctx->device);
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    for (int i = first;  i < first + count;  i++)
    {
        b->h_flags[i] &= ~SB_DTMF_FLAG_REALTIME;
        if (on)
            b->h_flags[i] |= SB_DTMF_FLAG_REALTIME;
    }
    CK(cudaMemcpy(b->flags + first, &b->h_flags[first], count, cudaMemcpyHostToDevice));
    CK(cudaMemset(b->duration + first, 0, sizeof(int)*count));          // src/dtmf.c:417
    return 0;
}

extern "C" int span_b200_dtmf_bank_fillin(span_b200_bank_t *b, int first, int count)
{
    if (!range_ok(b, SPAN_B200_DET_DTMF, first, count))
        return -1;
    if (count == 0)
        return 0;
    SB_DEVICE_CK(b->ctx->device);
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    const size_t C = b->channels;
    // src/dtmf.c:363-379: restart the Goertzels and the energy sum; hit history is kept.
    for (int i = 0;  i < 8;  i++)
    {
        CK(cudaMemset(b->v2 + i*C + first, 0, sizeof(float)*count));
        CK(cudaMemset(b->v3 + i*C + first, 0, sizeof(float)*count));
    }
    CK(cudaMemset(b->energy + first, 0, sizeof(float)*count));
    CK(cudaMemset(b->cs + first, 0, sizeof(int)*count));
    if (!(count == b->channels  ||  b->uniform_cs == 0))
        b->uniform_cs = -1;
    else
        b->uniform_cs = 0;
    return 0;
}

extern "C" int span_b200_bank_status(span_b200_bank_t *b, int first, int count, int32_t *status)
{
    if (first < 0  ||  count < 0  ||  first + count > b->channels)
    {
        sb_set_error("channel range out of bounds");
        return -1;
    }
    if (count == 0)
        return 0;
    SB_DEVICE_CK(b->ctx->device);
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    std::vector<unsigned char> a(count);
    std::vector<unsigned char> l(count);
    switch (b->det)
    {
    case SPAN_B200_DET_DTMF:
        CK(cudaMemcpy(a.data(), b->in_digit + first, count, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(l.data(), b->last_hit + first, count, cudaMemcpyDeviceToHost));
        for (int i = 0;  i < count;  i++)
            status[i] = (a[i])  ?  a[i]  :  ((l[i])  ?  'x'  :  0);    // src/dtmf.c:382-391
        break;
    case SPAN_B200_DET_R2_MF:
        CK(cudaMemcpy(a.data(), b->hits + first, count, cudaMemcpyDeviceToHost));
        for (int i = 0;  i < count;  i++)
            status[i] = a[i];
        break;
    case SPAN_B200_DET_SUPER_TONE:
        CK(cudaMemcpy(status, b->detected + first, sizeof(int)*count, cudaMemcpyDeviceToHost));
        break;
    default:
        for (int i = 0;  i < count;  i++)
            status[i] = 0;
        break;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// launches

struct Geometry
{
    int cs0;
    int nb;             // complete blocks (uniform) or the per-channel maximum
    int slice_blocks;
    int nslices;
    bool staged;
};

template <class DET, int SEG_VEC, int NSTAGE, int WARPS, int MINB, int NPACK, bool IN8, bool FILTK>
static int launch_staged_k(const BankArgs<DET> &a, cudaStream_t st)
{
    typedef StageCfg<SEG_VEC, NSTAGE> cfg;
    const int smem = cfg::WARP_BYTES*WARPS + ((IN8)  ?  1024  :  0);
    auto kern = bank_kernel_staged<DET, SEG_VEC, NSTAGE, WARPS, MINB, NPACK, IN8, FILTK>;
    // function attributes are per device: one flag per device ordinal, not one per process
    static bool configured[64] = {false};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0  ||  dev >= 64  ||  !configured[dev])
    {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (dev >= 0  &&  dev < 64)
            configured[dev] = true;
    }
    const long long ngroups = (a.channels + 31)/32;
    const long long items = ngroups*a.nslices;
    const long long grid = (items + WARPS - 1)/WARPS;
    kern<<<(unsigned int) grid, WARPS*32, smem, st>>>(a);
    CK(cudaGetLastError());
    return 0;
}

// `filter`: some channel of the bank has the DTMF dial-tone notch on (only DtmfDet has an instantiation for it)
template <class DET, int SEG_VEC, int NSTAGE, int WARPS, int MINB, int NPACK, bool IN8 = false>
static int launch_staged(const BankArgs<DET> &a, bool filter, cudaStream_t st)
{
    if constexpr (DET::FILTER)
    {
        if (filter)
            return launch_staged_k<DET, SEG_VEC, NSTAGE, WARPS, MINB, NPACK, IN8, true>(a, st);
    }
    return launch_staged_k<DET, SEG_VEC, NSTAGE, WARPS, MINB, NPACK, IN8, false>(a, st);
}

template <class DET, int NPACK>
static int launch_variant(const BankArgs<DET> &a, bool filter, cudaStream_t st, int variant = 0)
{
    // Occupancy experiments (DTMF only, tuning knob 1): more resident warps per SM at a lower register cap
    if constexpr (std::is_same<DET, DtmfDet>::value  &&  (NPACK == DET::NPAIRS  ||  NPACK == DET::NPAIRS + 1))
    {
        if (variant == 1)
            return launch_staged<DET, 8, 2, 4, 6, NPACK>(a, filter, st);    // 24 warps/SM, <= 80 registers
        if (variant == 2)
            return launch_staged<DET, 8, 2, 8, 3, NPACK>(a, filter, st);    // 24 warps/SM in 3 CTAs of 8 warps
        if (variant == 3)
            return launch_staged<DET, 8, 2, 2, 10, NPACK>(a, filter, st);   // 20 warps/SM in 10 CTAs of 2 warps
        if constexpr (NPACK == DET::NPAIRS + 1)
        {
            if (variant == 4)
                return launch_staged<DET, 8, 2, 4, 5, NPACK>(a, filter, st);    // 20 warps/SM, <= 96 registers
            if (variant == 5)
                return launch_staged<DET, 8, 3, 4, 4, NPACK>(a, filter, st);    // three stages in flight
        }
    }
    // (row bytes per stage = SEG_VEC*16, stages, warps per CTA, min CTAs per SM).  The round-1 sweep
    // (profiles/r01_sweep_dtmf*.json) covered nine shapes; the fastest is kept: 128-byte row segments,
    // 2 stages, 16-20 resident warps per SM.
    // More than 32 bins (super-tone descriptors with 33..64 monitored frequencies): the resonators alone need
    // 4*NPAIRS registers, so two CTAs per SM instead of four
    if constexpr (DET::NPAIRS > 16)
        return launch_staged<DET, 8, 2, 4, 2, NPACK>(a, filter, st);
    else
        return launch_staged<DET, 8, 2, 4, 4, NPACK>(a, filter, st);    // 272 B/row
}

template <class DET>
static int launch_bank(span_b200_bank_t *b, BankArgs<DET> &a, const Geometry &g, cudaStream_t st, bool all_variants)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    const bool filter = (b->det == SPAN_B200_DET_DTMF  &&  b->n_filter > 0);
    if (a.lut)
    {
        // 8-bit companded input: one staged shape, or the direct kernel
        a.cs0 = g.cs0;
        a.slice_blocks = g.slice_blocks;
        a.nslices = g.nslices;
        a.nblocks = g.nb;
        if (g.staged)
        {
            b->last_path = "staged";
            if constexpr (DET::NPAIRS > 16)
                return launch_staged<DET, 8, 2, 4, 2, DET::NPAIRS + 1, true>(a, filter, st);
            else
                return launch_staged<DET, 8, 2, 4, 4, DET::NPAIRS + 1, true>(a, filter, st);
        }
        b->last_path = "direct";
        bank_kernel_direct<DET, DET::NPAIRS + 1, true><<<(a.channels + 127)/128, 128, 0, st>>>(a);
        CK(cudaGetLastError());
        return 0;
    }
    a.cs0 = g.cs0;
    a.slice_blocks = g.slice_blocks;
    a.nslices = g.nslices;
    a.nblocks = g.nb;
    if (g.staged)
    {
        b->last_path = "staged";
        // Default: 2-wide multiply, subtract and add (NPACK = NPAIRS + 1).  DTMF keeps the earlier forms behind
        // knob 3 for the comparison in DESIGN.md: 0 = all scalar, 4 = scalar FMUL + FADD2 (also what the
        // occupancy variants of knob 1 are built with).
        if constexpr (std::is_same<DET, DtmfDet>::value)
        {
            if (all_variants  &&  b->tune_packed == 0)
                return launch_variant<DET, 0>(a, filter, st);
            if (all_variants  &&  b->tune_packed == 4)
                return launch_variant<DET, DET::NPAIRS>(a, filter, st, b->tune_variant);
            if (all_variants  &&  b->tune_variant != 0)
                return launch_variant<DET, DET::NPAIRS + 1>(a, filter, st, b->tune_variant);
        }
        return launch_variant<DET, DET::NPAIRS + 1>(a, filter, st);
    }
    b->last_path = "direct";
    const int grid = (a.channels + 127)/128;
    if (b->tune_packed  ||  !all_variants)
        bank_kernel_direct<DET, DET::NPAIRS + 1, false><<<grid, 128, 0, st>>>(a);
    else
        bank_kernel_direct<DET, 0, false><<<grid, 128, 0, st>>>(a);
    CK(cudaGetLastError());
    return 0;
}

template <int NP>
static int launch_st(span_b200_bank_t *b, BankArgs<SuperToneDet<NP> > &a, const Geometry &g, cudaStream_t st)
{
    return launch_bank<SuperToneDet<NP> >(b, a, g, st, false);
}

// One launch covers the channels [c0, c0 + count) of the bank: every per-channel pointer is advanced to c0, the
// arrays keep the whole bank's row pitch.  d_amp is the row of channel c0 (bytes per sample: 2, or 1 for G.711).
template <class DET>
static void fill_common(span_b200_bank_t *b, BankArgs<DET> &a, const int16_t *d_amp, int64_t stride, int n, int c0, int count)
{
    memset(&a, 0, sizeof(a));
    a.amp = d_amp;
    a.stride = stride;
    a.n = n;
    a.channels = count;
    a.cstride = b->channels;
    a.v2 = b->v2 + c0;
    a.v3 = b->v3 + c0;
    a.energy = b->energy + c0;
    a.cs = b->cs + c0;
    a.code = (typename DET::code_t *) b->code + c0;
    a.eout = (b->eout)  ?  (b->eout + c0)  :  NULL;
    a.block_rt = b->block;
}

template <int NP>
static int run_st(span_b200_bank_t *b, const int16_t *d_amp, int64_t stride, int n, const Geometry &g, cudaStream_t st, const float *lut,
                  int c0, int count)
{
    BankArgs<SuperToneDet<NP> > a;
    fill_common(b, a, d_amp, stride, n, c0, count);
    a.lut = lut;
    a.det = b->stp;
    return launch_st<NP>(b, a, g, st);
}

static long long worst_case_events(const span_b200_bank_t *b, int nb)
{
    const long long C = b->channels;
    switch (b->det)
    {
    case SPAN_B200_DET_DTMF:
        // An event needs two consecutive blocks that differ from in_digit, and it leaves last_hit == in_digit
        // (src/dtmf.c:304-346: last_hit takes the debounced hit), so the block after an event cannot fire:
        // at most one event per two blocks in either mode.
        return C*(nb/2 + 1);
    case SPAN_B200_DET_BELL_MF:
        return C*(nb/2 + 1);
    case SPAN_B200_DET_R2_MF:
        return C*(long long) nb;
    default:
        return C*(3LL*nb + 2);          // per block: tone lost + segment + tone found (+ owed re-chunk)
    }
}

// One rx call = prepare (geometry, scratch, record buffers) + the filter-bank kernel over all channels - in one launch,
// or one launch per channel range when the samples arrive in pieces (rx_host) - + finish (the sequencers).
struct RxCall
{
    Geometry g;
    int n;
    int law;                        // -1 = int16 linear samples; 0 = u-law bytes; 1 = A-law bytes (stride in samples either way)
    const float *lut;
    cudaStream_t st;
    int slot;
    span_b200_wire_event_t *wire_out;
    long long out_cap;
    cudaEvent_t t1;
    cudaEvent_t before_emit;        // the gather transfer that still reads the record buffer this call's emit pass overwrites, or NULL
};

static int rx_prepare(span_b200_bank_t *b, const int16_t *d_amp, int64_t stride, int n, void *stream, int law, RxCall &rc)
{
    if (b == NULL  ||  n < 0  ||  (n > 0  &&  d_amp == NULL))
    {
        sb_set_error("bad rx arguments");
        return -1;
    }
    cudaStream_t st = (stream)  ?  (cudaStream_t) stream  :  b->ctx->stream;
    if (b->have_last  &&  b->last_stream != st)
        CK(cudaStreamSynchronize(b->last_stream));
    const int B = b->block;
    Geometry &g = rc.g;
    rc.n = n;
    rc.law = law;
    rc.st = st;
    rc.t1 = NULL;
    rc.before_emit = NULL;
    g.cs0 = b->uniform_cs;
    const bool aligned = ((((uintptr_t) d_amp) & 15) == 0)  &&  ((stride & ((law >= 0)  ?  15  :  7)) == 0);
    g.staged = (g.cs0 >= 0)  &&  aligned  &&  !b->tune_direct  &&  n > 0;
    rc.lut = (law >= 0)  ?  (b->ctx->d_g711_lut + 256*law)  :  NULL;
    g.nb = (g.cs0 >= 0)  ?  (g.cs0 + n)/B  :  (B - 1 + n)/B;
    // ---- time slicing ----
    g.nslices = 1;
    g.slice_blocks = g.nb + 1;
    if (g.staged  &&  !(b->det == SPAN_B200_DET_DTMF  &&  b->n_filter > 0))
    {
        int L = b->tune_slice;
        if (L <= 0)
        {
            // Enough work items for ~16 waves of resident warps, but slices no shorter than 16 blocks.
            const long long ngroups = (b->channels + 31)/32;
            const long long want_items = (long long) b->ctx->sm_count*16*16;
            long long slices = (want_items + ngroups - 1)/ngroups;
            if (slices < 1)
                slices = 1;
            L = (int) ((g.nb + slices - 1)/slices);
            if (L < 16)
                L = 16;
        }
        if (g.nb > L)
        {
            g.slice_blocks = L;
            g.nslices = (g.nb + L - 1)/L;
        }
    }
    // ---- scratch ----
    const size_t C = b->channels;
    const size_t code_sz = (b->det == SPAN_B200_DET_SUPER_TONE)  ?  2  :  1;
    if (ensure(&b->code, &b->code_bytes, code_sz*C*(size_t) std::max(g.nb, 1)) != 0)
        return -1;
    if (b->det == SPAN_B200_DET_DTMF)
    {
        if (ensure((void **) &b->eout, &b->eout_bytes, sizeof(float)*C*(size_t) std::max(g.nb, 1)) != 0)
            return -1;
    }
    long long want_ev = (b->ev_cap_user > 0)  ?  b->ev_cap_user  :  worst_case_events(b, g.nb);
    if (want_ev < 1)
        want_ev = 1;
    span_b200_wire_event_t *wire_out = NULL;
    long long out_cap = 0;
    int slot = 0;
    if (b->wire_on)
    {
        if (g.nb > SPAN_B200_WIRE_MAX_BLOCKS)
        {
            sb_set_error("wire records hold block numbers below %d; this call has %d blocks per channel", SPAN_B200_WIRE_MAX_BLOCKS, g.nb);
            return -1;
        }
        slot = b->slot ^ 1;
        const bool to_gather = (b->comm != NULL  &&  b->comm->rank == b->root);
        // the transfer that read this buffer two calls ago must be over before the emit pass overwrites it: the wait
        // goes in front of the emit pass, not of the filter bank, so the transfer has this call's filter kernel to hide
        // behind as well
        rc.before_emit = (b->comm  &&  b->comm->done_valid[slot])  ?  b->comm->done[slot]  :  NULL;
        if (to_gather)
        {
            // root: emit straight into the gather buffer; room for every rank's worst case (grown on demand later)
            if (comm_gather_reserve(b->comm, slot, want_ev*b->comm->nranks, st, 0) != 0)
                return -1;
            wire_out = b->comm->gather[slot];
            out_cap = want_ev;
        }
        else
        {
            if (b->wire_cap[slot] < want_ev  ||  (b->ev_cap_user > 0  &&  b->wire_cap[slot] != want_ev))
            {
                if (b->wire[slot])
                {
                    // (another rank may still be reading it)
                    if (b->comm  &&  b->comm->done_valid[slot])
                        CK(cudaEventSynchronize(b->comm->done[slot]));
                    CK(cudaFree(b->wire[slot]));
                }
                b->wire[slot] = NULL;
                b->wire_cap[slot] = 0;
                CK(cudaMalloc(&b->wire[slot], sizeof(span_b200_wire_event_t)*(size_t) want_ev));
                b->wire_cap[slot] = want_ev;
                // what another rank needs to read this buffer over NVLink (peer-copy transport of the gather)
                cudaIpcMemHandle_t h;
                static_assert(sizeof(h) == 8*sizeof(unsigned long long), "cudaIpcMemHandle_t is 64 bytes");
                unsigned long long *hm = b->h_total + slot*SB_META_WORDS;
                if (cudaIpcGetMemHandle(&h, b->wire[slot]) == cudaSuccess)
                {
                    hm[1] = ++b->wire_gen;
                    memcpy(&hm[2], &h, sizeof(h));
                }
                else
                {
                    cudaGetLastError();
                    hm[1] = 0;                  // no handle: only the NCCL transport can move these records
                    memset(&hm[2], 0, sizeof(h));
                }
                CK(cudaMemcpyAsync(b->d_total + slot*SB_META_WORDS + 1, hm + 1, (SB_META_WORDS - 1)*sizeof(unsigned long long),
                                   cudaMemcpyHostToDevice, st));
            }
            wire_out = b->wire[slot];
            out_cap = b->wire_cap[slot];
        }
        if (b->emitted[slot] == NULL)
            CK(cudaEventCreateWithFlags(&b->emitted[slot], cudaEventDisableTiming));
    }
    else
    {
        if (b->ev_cap < want_ev  ||  (b->ev_cap_user > 0  &&  b->ev_cap != want_ev))
        {
            if (b->events)
                CK(cudaFree(b->events));
            b->events = NULL;
            b->ev_cap = 0;
            CK(cudaMalloc(&b->events, sizeof(span_b200_event_t)*(size_t) want_ev));
            b->ev_cap = want_ev;
        }
        out_cap = b->ev_cap;
    }
    rc.slot = slot;
    rc.wire_out = wire_out;
    rc.out_cap = out_cap;
    b->last_launches = 0;
    b->last_path = "empty";
    if (b->tune_timing  &&  n > 0)
    {
        while ((int) b->ev_pool.size() < b->ev_used + 2)
        {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            b->ev_pool.push_back(e);
        }
        cudaEvent_t t0 = b->ev_pool[b->ev_used++];
        rc.t1 = b->ev_pool[b->ev_used++];
        CK(cudaEventRecord(t0, st));
    }
    return 0;
}

// The filter-bank kernel for channels [c0, c0 + count); d_amp = the row of channel c0
static int rx_bank_launch(span_b200_bank_t *b, const RxCall &rc, const int16_t *d_amp, int64_t stride, int c0, int count)
{
    if (rc.n <= 0  ||  count <= 0)
        return 0;
    const Geometry &g = rc.g;
    cudaStream_t st = rc.st;
    const float *lut = rc.lut;
    const int n = rc.n;
    int rcode = 0;
    switch (b->det)
    {
    case SPAN_B200_DET_DTMF:
        {
            BankArgs<DtmfDet> a;
            fill_common(b, a, d_amp, stride, n, c0, count);
            a.lut = lut;
            a.det.threshold = b->thr + c0;
            a.det.normal_twist = b->ntw + c0;
            a.det.reverse_twist = b->rtw + c0;
            a.det.flags = b->flags + c0;
            a.det.z = b->z + c0;
            rcode = launch_bank<DtmfDet>(b, a, g, st, true);
        }
        break;
    case SPAN_B200_DET_BELL_MF:
        {
            BankArgs<BellMfDet> a;
            fill_common(b, a, d_amp, stride, n, c0, count);
            a.lut = lut;
            rcode = launch_bank<BellMfDet>(b, a, g, st, false);
        }
        break;
    case SPAN_B200_DET_R2_MF:
        {
            BankArgs<R2MfDet> a;
            fill_common(b, a, d_amp, stride, n, c0, count);
            a.lut = lut;
            a.det.fwd = b->fwd;
            rcode = launch_bank<R2MfDet>(b, a, g, st, false);
        }
        break;
    case SPAN_B200_DET_SUPER_TONE:
        if (b->npairs <= 1)
            rcode = run_st<1>(b, d_amp, stride, n, g, st, lut, c0, count);
        else if (b->npairs <= 2)
            rcode = run_st<2>(b, d_amp, stride, n, g, st, lut, c0, count);
        else if (b->npairs <= 3)
            rcode = run_st<3>(b, d_amp, stride, n, g, st, lut, c0, count);
        else if (b->npairs <= 4)
            rcode = run_st<4>(b, d_amp, stride, n, g, st, lut, c0, count);
        else if (b->npairs <= 5)
            rcode = run_st<5>(b, d_amp, stride, n, g, st, lut, c0, count);
        else if (b->npairs <= 6)
            rcode = run_st<6>(b, d_amp, stride, n, g, st, lut, c0, count);
        else if (b->npairs <= 8)
            rcode = run_st<8>(b, d_amp, stride, n, g, st, lut, c0, count);
        else if (b->npairs <= 10)
            rcode = run_st<10>(b, d_amp, stride, n, g, st, lut, c0, count);
        else if (b->npairs <= 12)
            rcode = run_st<12>(b, d_amp, stride, n, g, st, lut, c0, count);
        else if (b->npairs <= 16)
            rcode = run_st<16>(b, d_amp, stride, n, g, st, lut, c0, count);
        else if (b->npairs <= 20)
            rcode = run_st<20>(b, d_amp, stride, n, g, st, lut, c0, count);
        else if (b->npairs <= 24)
            rcode = run_st<24>(b, d_amp, stride, n, g, st, lut, c0, count);
        else
            rcode = run_st<32>(b, d_amp, stride, n, g, st, lut, c0, count);
        break;
    }
    if (rcode != 0)
        return -1;
    b->last_launches++;
    return 0;
}

// The sequencers: count, scan, emit
static int rx_finish(span_b200_bank_t *b, const RxCall &rc)
{
    const Geometry &g = rc.g;
    cudaStream_t st = rc.st;
    const int n = rc.n;
    const int B = b->block;
    const int slot = rc.slot;
    if (rc.t1)
        CK(cudaEventRecord(rc.t1, st));
    SeqCommon q;
    q.channels = b->channels;
    q.n = n;
    q.cs0 = g.cs0;
    q.cs = b->cs;
    q.offsets = b->offsets;
    q.counts = b->counts;
    q.events = (b->wire_on)  ?  NULL  :  b->events;
    q.wire = rc.wire_out;
    q.channel_base = b->channel_base;
    q.capacity = rc.out_cap;
    const int sgrid = (b->channels + 127)/128;
    const int cpw = (b->det == SPAN_B200_DET_SUPER_TONE)  ?  SB_ST_CPW  :  32;      // channels per sequencer warp
    for (int pass = 0;  pass < 2;  pass++)
    {
        if (pass == 1  &&  rc.before_emit)
            CK(cudaStreamWaitEvent(st, rc.before_emit, 0));
        switch (b->det)
        {
        case SPAN_B200_DET_DTMF:
            {
                DtmfSeqArgs s;
                s.q = q;
                s.code = (const unsigned char *) b->code;
                s.eout = b->eout;
                s.flags = b->flags;
                s.last_hit = b->last_hit;
                s.in_digit = b->in_digit;
                s.duration = b->duration;
                s.level_thr = b->ctx->d_level_thr;
                s.level_n = b->ctx->level_n;
                s.level_min = b->ctx->level_min;
                if (pass == 0)
                    dtmf_sequencer<false><<<sgrid, 128, 0, st>>>(s);
                else
                    dtmf_sequencer<true><<<sgrid, 128, 0, st>>>(s);
            }
            break;
        case SPAN_B200_DET_BELL_MF:
        case SPAN_B200_DET_R2_MF:
            {
                MfSeqArgs s;
                s.q = q;
                s.code = (const unsigned char *) b->code;
                s.hits = b->hits;
                if (b->det == SPAN_B200_DET_BELL_MF)
                {
                    if (pass == 0)
                        bell_mf_sequencer<false><<<sgrid, 128, 0, st>>>(s);
                    else
                        bell_mf_sequencer<true><<<sgrid, 128, 0, st>>>(s);
                }
                else
                {
                    if (pass == 0)
                        r2_mf_sequencer<false><<<sgrid, 128, 0, st>>>(s);
                    else
                        r2_mf_sequencer<true><<<sgrid, 128, 0, st>>>(s);
                }
            }
            break;
        case SPAN_B200_DET_SUPER_TONE:
            {
                StSeqArgs s;
                s.q = q;
                s.code = (const unsigned short *) b->code;
                s.t.tones = b->tones;
                s.t.tone_segs = b->d_tone_segs;
                s.t.tone_first = b->d_tone_first;
                s.t.elements = b->d_elements;
                s.t.total_elements = b->total_elements;
                s.segments = b->segments;
                s.detected_tone = b->detected;
                s.rotation = b->rotation;
                s.pending = b->pending;
                s.want_segments = b->want_segments;
                s.log = b->st_log;
                const int stgrid = (b->channels + SB_ST_CPC - 1)/SB_ST_CPC;
                const bool small = (b->tones <= SB_ST_SMEM_TONES  &&  b->total_elements <= SB_ST_SMEM_ELEMENTS);
                if (pass == 0)
                {
                    if (small)
                        super_tone_sequencer<false, true><<<stgrid, 128, 0, st>>>(s);
                    else
                        super_tone_sequencer<false, false><<<stgrid, 128, 0, st>>>(s);
                }
                else
                {
                    if (small)
                        super_tone_sequencer<true, true><<<stgrid, 128, 0, st>>>(s);
                    else
                        super_tone_sequencer<true, false><<<stgrid, 128, 0, st>>>(s);
                }
            }
            break;
        }
        CK(cudaGetLastError());
        b->last_launches++;
        if (pass == 0)
        {
            scan_counts<<<1, 1024, 0, st>>>(b->counts, b->offsets, (b->channels + cpw - 1)/cpw, b->d_total + slot*SB_META_WORDS);
            CK(cudaGetLastError());
            b->last_launches++;
            CK(cudaMemcpyAsync(b->h_total + slot*SB_META_WORDS, b->d_total + slot*SB_META_WORDS, sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost, st));
        }
    }
    if (b->wire_on)
    {
        CK(cudaEventRecord(b->emitted[slot], st));
        b->slot = slot;
    }
    if (g.cs0 >= 0)
        b->uniform_cs = (g.cs0 + n) % B;
    b->last_nb = g.nb;
    b->last_stream = st;
    b->have_last = true;
    return 0;
}

static int rx_core(span_b200_bank_t *b, const int16_t *d_amp, int64_t stride, int n, void *stream, int law)
{
    if (b == NULL)
    {
        sb_set_error("bad rx arguments");
        return -1;
    }
    SB_DEVICE_CK(b->ctx->device);
    RxCall rc;
    if (rx_prepare(b, d_amp, stride, n, stream, law, rc) != 0)
        return -1;
    if (rx_bank_launch(b, rc, d_amp, stride, 0, b->channels) != 0)
        return -1;
    return rx_finish(b, rc);
}

extern "C" int span_b200_bank_rx_device(span_b200_bank_t *b, const int16_t *d_amp, int64_t stride,
                                        int n, void *stream)
{
    return rx_core(b, d_amp, stride, n, stream, -1);
}

extern "C" int span_b200_bank_rx_device_g711(span_b200_bank_t *b, const uint8_t *d_data, int64_t stride,
                                             int n, int alaw, void *stream)
{
    return rx_core(b, (const int16_t *) d_data, stride, n, stream, (alaw)  ?  1  :  0);
}

// Host samples.  The copy and the filter bank are pipelined over channel ranges: the rows of range k + 1 cross PCIe
// (on the bank's copy stream) while the filter-bank kernel of range k runs; the sequencers follow once, over all
// channels.  Pinned host memory makes the copies asynchronous; pageable memory works, serialised by the driver.
#define SB_RX_HOST_PIECES   8

static int rx_host_core(span_b200_bank_t *b, const void *h_data, int64_t stride, int n, void *stream, int law)
{
    if (b == NULL  ||  n < 0  ||  (n > 0  &&  h_data == NULL))
    {
        sb_set_error("bad rx arguments");
        return -1;
    }
    SB_DEVICE_CK(b->ctx->device);
    cudaStream_t st = (stream)  ?  (cudaStream_t) stream  :  b->ctx->stream;
    if (b->have_last  &&  b->last_stream != st)
        CK(cudaStreamSynchronize(b->last_stream));
    const int bps = (law >= 0)  ?  1  :  2;
    // Device rows are padded to a multiple of 16 bytes so that the staged kernel applies.
    const int64_t dstride = (law >= 0)  ?  ((n + 15) & ~15LL)  :  ((n + 7) & ~7LL);
    if (ensure((void **) &b->d_in, &b->d_in_bytes, (size_t) bps*(size_t) dstride*b->channels + 16) != 0)
        return -1;
    RxCall rc;
    if (rx_prepare(b, b->d_in, dstride, n, (void *) st, law, rc) != 0)
        return -1;
    if (n > 0)
    {
        if (b->copy_stream == NULL)
        {
            CK(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
            for (int i = 0;  i < SB_RX_HOST_PIECES;  i++)
                CK(cudaEventCreateWithFlags(&b->copied[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&b->in_free, cudaEventDisableTiming));
        }
        // the previous call's kernels (on st) may still read d_in
        CK(cudaEventRecord(b->in_free, st));
        CK(cudaStreamWaitEvent(b->copy_stream, b->in_free, 0));
        // pieces of whole 32-channel groups; small banks go in one piece
        int pieces = SB_RX_HOST_PIECES;
        while (pieces > 1  &&  (long long) b->channels*n < (long long) pieces*(4 << 20))
            pieces >>= 1;
        const int groups = (b->channels + 31)/32;
        for (int k = 0;  k < pieces;  k++)
        {
            const int c0 = (int) ((long long) groups*k/pieces)*32;
            int c1 = (int) ((long long) groups*(k + 1)/pieces)*32;
            if (c1 > b->channels)
                c1 = b->channels;
            if (c1 <= c0)
                continue;
            const char *src = (const char *) h_data + (size_t) c0*(size_t) stride*bps;
            char *dst = (char *) b->d_in + (size_t) c0*(size_t) dstride*bps;
            if (stride == n  &&  dstride == n)
            {
                // contiguous on both sides: one flat copy (measured 55 GB/s vs ~50 GB/s for the pitched form)
                CK(cudaMemcpyAsync(dst, src, (size_t) bps*(size_t) n*(size_t) (c1 - c0), cudaMemcpyHostToDevice, b->copy_stream));
            }
            else
            {
                CK(cudaMemcpy2DAsync(dst, (size_t) bps*dstride, src, (size_t) bps*stride, (size_t) bps*n, (size_t) (c1 - c0),
                                     cudaMemcpyHostToDevice, b->copy_stream));
            }
            CK(cudaEventRecord(b->copied[k], b->copy_stream));
            CK(cudaStreamWaitEvent(st, b->copied[k], 0));
            if (rx_bank_launch(b, rc, (const int16_t *) dst, dstride, c0, c1 - c0) != 0)
                return -1;
        }
    }
    return rx_finish(b, rc);
}

extern "C" int span_b200_bank_rx_host_g711(span_b200_bank_t *b, const uint8_t *h_data, int64_t stride,
                                           int n, int alaw, void *stream)
{
    return rx_host_core(b, h_data, stride, n, stream, (alaw)  ?  1  :  0);
}

extern "C" int span_b200_bank_rx_host(span_b200_bank_t *b, const int16_t *h_amp, int64_t stride,
                                      int n, void *stream)
{
    return rx_host_core(b, h_amp, stride, n, stream, -1);
}

extern "C" int64_t span_b200_bank_event_count(span_b200_bank_t *b, int *overflow)
{
    if (overflow)
        *overflow = 0;
    if (!b->have_last)
        return 0;
    SB_DEVICE_CK(b->ctx->device);
    CK(cudaStreamSynchronize(b->last_stream));
    const int slot = (b->wire_on)  ?  b->slot  :  0;
    long long total = (long long) b->h_total[slot*SB_META_WORDS];
    long long cap = b->ev_cap;
    if (b->wire_on)
        cap = (b->comm  &&  b->comm->rank == b->root)  ?  (b->comm->gather_cap[slot]/b->comm->nranks)  :  b->wire_cap[slot];
    if (total > cap)
    {
        if (overflow)
            *overflow = 1;
        total = cap;
    }
    return total;
}

static int not_in_wire_mode(const span_b200_bank_t *b)
{
    if (b->wire_on)
    {
        sb_set_error("the bank is in wire mode: use span_b200_bank_events_wire() / the gather calls");
        return 0;
    }
    return 1;
}

extern "C" int64_t span_b200_bank_events(span_b200_bank_t *b, span_b200_event_t *out, int64_t max)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (!not_in_wire_mode(b))
        return -1;
    int64_t total = span_b200_bank_event_count(b, NULL);
    if (total < 0)
        return -1;
    if (total > max)
        total = max;
    if (total > 0)
        CK(cudaMemcpy(out, b->events, sizeof(span_b200_event_t)*(size_t) total, cudaMemcpyDeviceToHost));
    return total;
}

extern "C" int64_t span_b200_bank_events_to_device(span_b200_bank_t *b, span_b200_event_t *d_out, int64_t max, void *stream)
{
    sb_device_guard sb_dg_((b)  ?  span_b200_ctx_device(b->ctx)  :  -1);
    if (!not_in_wire_mode(b))
        return -1;
    int64_t total = span_b200_bank_event_count(b, NULL);
    if (total < 0)
        return -1;
    if (total > max)
        total = max;
    cudaStream_t st = (stream)  ?  (cudaStream_t) stream  :  b->ctx->stream;
    if (total > 0)
        CK(cudaMemcpyAsync(d_out, b->events, sizeof(span_b200_event_t)*(size_t) total, cudaMemcpyDeviceToDevice, st));
    return total;
}

extern "C" double span_b200_bank_kernel_ms(span_b200_bank_t *b, int *launches)
{
    double ms = 0.0;
    if (launches)
        *launches = 0;
    sb_device_guard sb_dg_(b->ctx->device);
    if (!sb_dg_.ok())
        return -1.0;
    if (b->have_last  &&  cudaStreamSynchronize(b->last_stream) != cudaSuccess)
        return -1.0;
    for (int i = 0;  i + 1 < b->ev_used;  i += 2)
    {
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, b->ev_pool[i], b->ev_pool[i + 1]) == cudaSuccess)
            ms += t;
        if (launches)
            (*launches)++;
    }
    b->ev_used = 0;
    return ms;
}

extern "C" const span_b200_event_t *span_b200_bank_events_device(span_b200_bank_t *b)
{
    return (b->wire_on)  ?  NULL  :  b->events;
}

extern "C" int span_b200_bank_block_codes(span_b200_bank_t *b, uint16_t *codes, int64_t max)
{
    if (!b->have_last)
        return 0;
    SB_DEVICE_CK(b->ctx->device);
    CK(cudaStreamSynchronize(b->last_stream));
    const int64_t total = (int64_t) b->last_nb*b->channels;
    const int64_t n = (total < max)  ?  total  :  max;
    if (n <= 0)
        return 0;
    if (b->det == SPAN_B200_DET_SUPER_TONE)
    {
        CK(cudaMemcpy(codes, b->code, sizeof(uint16_t)*(size_t) n, cudaMemcpyDeviceToHost));
    }
    else
    {
        std::vector<unsigned char> tmp((size_t) n);
        CK(cudaMemcpy(tmp.data(), b->code, (size_t) n, cudaMemcpyDeviceToHost));
        for (int64_t i = 0;  i < n;  i++)
            codes[i] = tmp[(size_t) i];
    }
    return (int) n;
}

// ------------------------------------------------------------------------------------------
// raw Goertzel banks
template <int NP, bool E>
static int run_raw(span_b200_ctx_t *ctx, const float *fac, int bins, int block_len, const int16_t *d_amp,
                   int64_t stride, int channels, int n, float *d_out, int64_t cap, float *d_energy, cudaStream_t st)
{
    BankArgs<RawDet<NP, E> > a;
    memset(&a, 0, sizeof(a));
    a.amp = d_amp;
    a.stride = stride;
    a.n = n;
    a.channels = channels;
    a.block_rt = block_len;
    a.raw = d_out;
    a.raw_capacity = cap;
    a.eout = d_energy;
    a.cstride = channels;
    a.det.bins = bins;
    for (int i = 0;  i < bins;  i++)
        a.det.fac[i] = fac[i];
    const int nb = n/block_len;
    a.cs0 = 0;
    a.nblocks = nb;
    // Carried state is not supported for raw banks: the tail (n % block_len samples) is dropped,
    // exactly what a caller sees if it only looks at goertzel_result() of complete blocks.
    a.n = nb*block_len;
    if (a.n == 0)
        return 0;
    const bool aligned = ((((uintptr_t) d_amp) & 15) == 0)  &&  ((stride & 7) == 0);
    if (aligned)
    {
        const long long ngroups = (channels + 31)/32;
        const long long want_items = (long long) ctx->sm_count*8*16;
        long long slices = (want_items + ngroups - 1)/ngroups;
        int L = (int) ((nb + slices - 1)/slices);
        if (L < 16)
            L = 16;
        a.slice_blocks = (nb > L)  ?  L  :  (nb + 1);
        a.nslices = (nb > L)  ?  ((nb + L - 1)/L)  :  1;
        return launch_staged<RawDet<NP, E>, 8, 2, 4, 4, NP>(a, false, st);
    }
    bank_kernel_direct<RawDet<NP, E>, RawDet<NP, E>::NPAIRS + 1, false><<<(channels + 127)/128, 128, 0, st>>>(a);
    CK(cudaGetLastError());
    return 0;
}

template <bool E>
static int goertzel_blocks(span_b200_ctx_t *ctx, const float *fac, int bins, int block_len, const int16_t *d_amp, int64_t stride, int channels,
                           int n, float *d_out, int64_t out_capacity, float *d_energy, void *stream)
{
    if (ctx == NULL  ||  fac == NULL  ||  bins < 1  ||  bins > SB_RAW_MAX_BINS  ||  block_len < 1  ||  channels < 1  ||  n < 0)
    {
        sb_set_error("bad goertzel bank arguments (1..%d bins)", SB_RAW_MAX_BINS);
        return -1;
    }
    SB_DEVICE_CK(ctx->device);
    cudaStream_t st = (stream)  ?  (cudaStream_t) stream  :  ctx->stream;
    const int np = (bins + 1)/2;
    int rc;
    if (np <= 1)
        rc = run_raw<1, E>(ctx, fac, bins, block_len, d_amp, stride, channels, n, d_out, out_capacity, d_energy, st);
    else if (np <= 2)
        rc = run_raw<2, E>(ctx, fac, bins, block_len, d_amp, stride, channels, n, d_out, out_capacity, d_energy, st);
    else if (np <= 4)
        rc = run_raw<4, E>(ctx, fac, bins, block_len, d_amp, stride, channels, n, d_out, out_capacity, d_energy, st);
    else if (np <= 8)
        rc = run_raw<8, E>(ctx, fac, bins, block_len, d_amp, stride, channels, n, d_out, out_capacity, d_energy, st);
    else
        rc = run_raw<16, E>(ctx, fac, bins, block_len, d_amp, stride, channels, n, d_out, out_capacity, d_energy, st);
    if (rc != 0)
        return -1;
    return n/block_len;
}

extern "C" int span_b200_goertzel_blocks_device(span_b200_ctx_t *ctx, const float *fac, int bins, int block_len,
                                                const int16_t *d_amp, int64_t stride, int channels, int n,
                                                float *d_out, int64_t out_capacity, void *stream)
{
    return goertzel_blocks<false>(ctx, fac, bins, block_len, d_amp, stride, channels, n, d_out, out_capacity, NULL, stream);
}

extern "C" int span_b200_goertzel_blocks_energy_device(span_b200_ctx_t *ctx, const float *fac, int bins, int block_len,
                                                       const int16_t *d_amp, int64_t stride, int channels, int n,
                                                       float *d_out, int64_t out_capacity, float *d_energy, void *stream)
{
    if (d_energy == NULL)
    {
        sb_set_error("no energy buffer");
        return -1;
    }
    return goertzel_blocks<true>(ctx, fac, bins, block_len, d_amp, stride, channels, n, d_out, out_capacity, d_energy, stream);
}

// The two remaining users of the Goertzel primitives in the reference, as tone sets for the raw bank
extern "C" int span_b200_goertzel_tone_set(int which, float *freqs, int max, int *block_len)
{
    static const float ademco[2] = {1400.0f, 2300.0f};                      // src/ademco_contactid.c:1179-1180, block 55 (:446)
    static const float v18[9] = {390.0f, 980.0f, 1180.0f, 1270.0f, 1300.0f, 1400.0f, 1650.0f, 1800.0f, 2225.0f};    // src/v18.c:200-211, block 102 (:177)
    const float *f;
    int n;
    int bl;
    switch (which)
    {
    case SPAN_B200_TONE_SET_ADEMCO_CONTACTID:
        f = ademco;
        n = 2;
        bl = 55;
        break;
    case SPAN_B200_TONE_SET_V18:
        f = v18;
        n = 9;
        bl = 102;
        break;
    default:
        sb_set_error("unknown tone set %d", which);
        return -1;
    }
    if (block_len)
        *block_len = bl;
    for (int i = 0;  i < n  &&  i < max;  i++)
        freqs[i] = f[i];
    return n;
}

extern "C" float span_b200_goertzel_coefficient(float freq)
{
    return sb_goertzel_fac(freq);
}

// ------------------------------------------------------------------------------------------
// RFC 4733 telephone-event payloads for the DTMF reports (the wire form of a detected digit)
extern "C" int span_b200_rfc4733_event_code(int digit)
{
    if (digit >= '0'  &&  digit <= '9')
        return digit - '0';
    switch (digit)
    {
    case '*': return 10;
    case '#': return 11;
    case 'A': return 12;
    case 'B': return 13;
    case 'C': return 14;
    case 'D': return 15;
    }
    return -1;
}

extern "C" void span_b200_rfc4733_pack(uint8_t out[4], int event, int end, int volume, int duration)
{
    if (volume < 0)
        volume = 0;
    if (volume > 63)
        volume = 63;
    if (duration < 0)
        duration = 0;
    if (duration > 0xFFFF)
        duration = 0xFFFF;
    out[0] = (uint8_t) event;
    out[1] = (uint8_t) (((end)  ?  0x80  :  0) | volume);
    out[2] = (uint8_t) (duration >> 8);
    out[3] = (uint8_t) (duration & 0xFF);
}

extern "C" int span_b200_rfc4733_dtmf(span_b200_rfc4733_state_t *st, int code, int level, int duration, uint8_t out[8])
{
    int n = 0;
    if (st->event >= 0)
    {
        // whatever the report says, the event in progress ends here; `duration` is how long it lasted
        span_b200_rfc4733_pack(out, st->event, 1, st->volume, duration);
        n++;
        st->event = -1;
    }
    if (code != 0)
    {
        const int ev = span_b200_rfc4733_event_code(code);
        if (ev >= 0)
        {
            st->event = ev;
            st->volume = (level < 0)  ?  -level  :  0;          // "power level of the tone, expressed in dBm0 after dropping the sign"
            if (st->volume > 63)
                st->volume = 63;
            span_b200_rfc4733_pack(out + 4*n, ev, 0, st->volume, 0);
            n++;
        }
    }
    return n;
}


// ------------------------------------------------------------------------------------------
// wire records and the multi-GPU gather: channels shard over ranks with no exchange on the filter
// path; the only traffic is the detected-digit / tone records on their way to the rank that replays the callbacks.

extern "C" int span_b200_bank_set_wire(span_b200_bank_t *b, int on, uint32_t channel_base)
{
    if (b == NULL)
        return -1;
    SB_DEVICE_CK(b->ctx->device);
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    if (!on  &&  b->comm)
    {
        sb_set_error("the bank is attached to a communicator; it stays in wire mode");
        return -1;
    }
    b->wire_on = (on != 0);
    b->channel_base = channel_base;
    b->h_total[0] = b->h_total[SB_META_WORDS] = 0;
    b->slot = 0;
    return 0;
}

extern "C" int64_t span_b200_bank_events_wire(span_b200_bank_t *b, span_b200_wire_event_t *out, int64_t max)
{
    if (b == NULL  ||  !b->wire_on)
    {
        sb_set_error("the bank is not in wire mode");
        return -1;
    }
    int64_t total = span_b200_bank_event_count(b, NULL);
    if (total < 0)
        return -1;
    if (total > max)
        total = max;
    SB_DEVICE_CK(b->ctx->device);
    const span_b200_wire_event_t *src = (b->comm  &&  b->comm->rank == b->root)  ?  b->comm->gather[b->slot]  :  b->wire[b->slot];
    if (total > 0)
        CK(cudaMemcpy(out, src, sizeof(span_b200_wire_event_t)*(size_t) total, cudaMemcpyDeviceToHost));
    return total;
}

extern "C" void span_b200_wire_expand(const span_b200_wire_event_t *in, span_b200_event_t *out, int64_t n, uint32_t channel_base)
{
    for (int64_t i = 0;  i < n;  i++)
    {
        out[i].channel = (int32_t) (in[i].channel - channel_base);
        out[i].block = SPAN_B200_WIRE_BLOCK(in[i]);
        out[i].kind = SPAN_B200_WIRE_KIND(in[i]);
        out[i].a = in[i].a;
        out[i].b = in[i].b;
        out[i].c = in[i].c;
    }
}

// ---- NCCL, resolved at run time ----------------------------------------------------------------
struct nccl_api_t
{
    void *handle;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRankConfig)(ncclComm_t *, int, ncclUniqueId, int, ncclConfig_t *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    ncclResult_t (*GetVersion)(int *);
};

static nccl_api_t g_nccl;
static std::mutex g_nccl_lock;

static const nccl_api_t *nccl_api()
{
    std::lock_guard<std::mutex> lk(g_nccl_lock);
    if (g_nccl.handle)
        return &g_nccl;
    // The copy the process already has (a host application that brought its own NCCL - PyTorch does), else the system's
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (h == NULL)
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == NULL)
    {
        sb_set_error("libnccl.so.2 not found (%s); the multi-GPU gather needs NCCL", dlerror());
        return NULL;
    }
    nccl_api_t a;
    memset(&a, 0, sizeof(a));
    a.GetUniqueId = (decltype(a.GetUniqueId)) dlsym(h, "ncclGetUniqueId");
    a.CommInitRankConfig = (decltype(a.CommInitRankConfig)) dlsym(h, "ncclCommInitRankConfig");
    a.CommInitRank = (decltype(a.CommInitRank)) dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy)) dlsym(h, "ncclCommDestroy");
    a.GetErrorString = (decltype(a.GetErrorString)) dlsym(h, "ncclGetErrorString");
    a.AllGather = (decltype(a.AllGather)) dlsym(h, "ncclAllGather");
    a.Send = (decltype(a.Send)) dlsym(h, "ncclSend");
    a.Recv = (decltype(a.Recv)) dlsym(h, "ncclRecv");
    a.GroupStart = (decltype(a.GroupStart)) dlsym(h, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd)) dlsym(h, "ncclGroupEnd");
    a.GetVersion = (decltype(a.GetVersion)) dlsym(h, "ncclGetVersion");
    if (!a.GetUniqueId  ||  !a.CommInitRank  ||  !a.CommDestroy  ||  !a.GetErrorString  ||  !a.AllGather  ||  !a.Send  ||  !a.Recv
        ||  !a.GroupStart  ||  !a.GroupEnd)
    {
        sb_set_error("libnccl.so.2 lacks a required entry point");
        return NULL;
    }
    a.handle = h;
    g_nccl = a;
    return &g_nccl;
}

#define NK(call) \
    do \
    { \
        ncclResult_t r_ = (call); \
        if (r_ != ncclSuccess) \
        { \
            sb_set_error("%s failed: %s (%s:%d)", #call, nc->GetErrorString(r_), __FILE__, __LINE__); \
            return -1; \
        } \
    } \
    while (0)

extern "C" int span_b200_comm_unique_id(unsigned char id[SPAN_B200_COMM_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) == SPAN_B200_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    const nccl_api_t *nc = nccl_api();
    if (nc == NULL)
        return -1;
    ncclUniqueId u;
    NK(nc->GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return 0;
}

extern "C" void span_b200_comm_destroy(span_b200_comm_t *cm)
{
    if (cm == NULL)
        return;
    sb_device_guard sb_dg_(cm->ctx->device);
    if (cm->stream)
        cudaStreamSynchronize(cm->stream);
    if (cm->nccl  &&  g_nccl.handle)
        g_nccl.CommDestroy(cm->nccl);
    cudaFree(cm->d_sync);
    for (int i = 0;  i < 2;  i++)
    {
        if (cm->peer_ptr[i])
        {
            for (int r = 0;  r < cm->nranks;  r++)
            {
                if (cm->peer_ptr[i][r])
                    cudaIpcCloseMemHandle(cm->peer_ptr[i][r]);
            }
            delete[] cm->peer_ptr[i];
            delete[] cm->peer_gen[i];
        }
        cudaFree(cm->gather[i]);
        cudaFree(cm->d_counts[i]);
        if (cm->h_counts[i])
            cudaFreeHost(cm->h_counts[i]);
        if (cm->counts_ready[i])
            cudaEventDestroy(cm->counts_ready[i]);
        if (cm->done[i])
            cudaEventDestroy(cm->done[i]);
    }
    for (int i = 0;  i < SB_PULL_STREAMS;  i++)
    {
        if (cm->pull[i])
        {
            cudaStreamSynchronize(cm->pull[i]);
            cudaStreamDestroy(cm->pull[i]);
        }
        if (cm->joined[i])
            cudaEventDestroy(cm->joined[i]);
    }
    if (cm->fork)
        cudaEventDestroy(cm->fork);
    if (cm->stream)
        cudaStreamDestroy(cm->stream);
    delete cm;
}

extern "C" span_b200_comm_t *span_b200_comm_create(span_b200_ctx_t *ctx, const unsigned char id[SPAN_B200_COMM_ID_BYTES], int nranks, int rank,
                                                   int max_ctas)
{
    if (ctx == NULL  ||  id == NULL  ||  nranks < 1  ||  rank < 0  ||  rank >= nranks)
    {
        sb_set_error("bad communicator arguments");
        return NULL;
    }
    const nccl_api_t *nc = nccl_api();
    if (nc == NULL)
        return NULL;
    SB_DEVICE_CKP(ctx->device);
    span_b200_comm_t *cm = new span_b200_comm_s();
    cm->ctx = ctx;
    cm->nranks = nranks;
    cm->rank = rank;
    cm->begun_slot = -1;
    cm->ended_slot = -1;
    // How the records travel: by default the root pulls them out of the ranks' buffers with its copy engines over
    // NVLink (no SM of any GPU is used, the filter kernels are compute-bound); SPANDSP_B200_GATHER=nccl selects
    // exact-count ncclSend / ncclRecv instead.  The counts and the completion are NCCL collectives either way.
    const char *tr = getenv("SPANDSP_B200_GATHER");
    cm->transport = (tr  &&  strcmp(tr, "nccl") == 0)  ?  SPAN_B200_GATHER_NCCL  :  SPAN_B200_GATHER_PEER_COPY;
    bool ok = cudaStreamCreateWithFlags(&cm->stream, cudaStreamNonBlocking) == cudaSuccess
              &&  cudaMalloc(&cm->d_sync, sizeof(unsigned long long)*(1 + nranks)) == cudaSuccess
              &&  cudaMemset(cm->d_sync, 0, sizeof(unsigned long long)*(1 + nranks)) == cudaSuccess;
    ok = ok  &&  cudaEventCreateWithFlags(&cm->fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0;  ok  &&  i < SB_PULL_STREAMS;  i++)
    {
        ok = cudaStreamCreateWithFlags(&cm->pull[i], cudaStreamNonBlocking) == cudaSuccess
             &&  cudaEventCreateWithFlags(&cm->joined[i], cudaEventDisableTiming) == cudaSuccess;
    }
    for (int i = 0;  ok  &&  i < 2;  i++)
    {
        cm->peer_ptr[i] = new void *[nranks]();
        cm->peer_gen[i] = new unsigned long long[nranks]();
        ok = cudaMalloc(&cm->d_counts[i], sizeof(unsigned long long)*SB_META_WORDS*nranks) == cudaSuccess
             &&  cudaMallocHost(&cm->h_counts[i], sizeof(unsigned long long)*SB_META_WORDS*nranks) == cudaSuccess
             &&  cudaEventCreateWithFlags(&cm->counts_ready[i], cudaEventDisableTiming) == cudaSuccess
             &&  cudaEventCreateWithFlags(&cm->done[i], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok)
    {
        sb_set_error("communicator allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        span_b200_comm_destroy(cm);
        return NULL;
    }
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclResult_t r;
    if (nc->CommInitRankConfig)
    {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.blocking = 1;
        if (max_ctas > 0)
        {
            cfg.minCTAs = 1;
            cfg.maxCTAs = max_ctas;
        }
        r = nc->CommInitRankConfig(&cm->nccl, nranks, u, rank, &cfg);
    }
    else
    {
        r = nc->CommInitRank(&cm->nccl, nranks, u, rank);
    }
    if (r != ncclSuccess)
    {
        sb_set_error("ncclCommInitRank failed: %s", nc->GetErrorString(r));
        cm->nccl = NULL;
        span_b200_comm_destroy(cm);
        return NULL;
    }
    return cm;
}

extern "C" int span_b200_comm_set_transport(span_b200_comm_t *cm, int transport)
{
    if (cm == NULL  ||  (transport != SPAN_B200_GATHER_NCCL  &&  transport != SPAN_B200_GATHER_PEER_COPY))
    {
        sb_set_error("bad transport");
        return -1;
    }
    cm->transport = transport;
    return 0;
}

extern "C" int span_b200_comm_transport(const span_b200_comm_t *cm) { return cm->transport; }

extern "C" int span_b200_comm_rank(const span_b200_comm_t *cm) { return cm->rank; }
extern "C" int span_b200_comm_nranks(const span_b200_comm_t *cm) { return cm->nranks; }

extern "C" int span_b200_comm_sync(span_b200_comm_t *cm)
{
    if (cm == NULL)
        return -1;
    SB_DEVICE_CK(cm->ctx->device);
    CK(cudaStreamSynchronize(cm->stream));
    return 0;
}

// Root: make gather[slot] hold `records` records.  keep > 0: the first `keep` records (the root's own, already
// written by the emit pass on stream `st`) survive a re-allocation.
static int comm_gather_reserve(span_b200_comm_t *cm, int slot, long long records, cudaStream_t st, long long keep)
{
    if (cm->gather_cap[slot] >= records)
        return 0;
    span_b200_wire_event_t *p = NULL;
    CK(cudaMalloc(&p, sizeof(span_b200_wire_event_t)*(size_t) records));
    if (cm->gather[slot])
    {
        if (keep > 0)
        {
            CK(cudaMemcpyAsync(p, cm->gather[slot], sizeof(span_b200_wire_event_t)*(size_t) keep, cudaMemcpyDeviceToDevice, st));
            CK(cudaStreamSynchronize(st));
        }
        CK(cudaStreamSynchronize(cm->stream));
        CK(cudaFree(cm->gather[slot]));
    }
    cm->gather[slot] = p;
    cm->gather_cap[slot] = records;
    return 0;
}

extern "C" int span_b200_bank_attach_comm(span_b200_bank_t *b, span_b200_comm_t *cm, int root)
{
    if (b == NULL  ||  cm == NULL  ||  root < 0  ||  root >= cm->nranks  ||  cm->ctx != b->ctx)
    {
        sb_set_error("bad attach arguments (bank and communicator must share a context)");
        return -1;
    }
    if (!b->wire_on)
    {
        sb_set_error("switch the bank to wire records first (span_b200_bank_set_wire)");
        return -1;
    }
    SB_DEVICE_CK(b->ctx->device);
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    b->comm = cm;
    b->root = root;
    return 0;
}

extern "C" int span_b200_bank_gather_begin(span_b200_bank_t *b)
{
    if (b == NULL  ||  b->comm == NULL  ||  !b->have_last)
    {
        sb_set_error("gather: no communicator attached, or no rx call yet");
        return -1;
    }
    span_b200_comm_t *cm = b->comm;
    const nccl_api_t *nc = nccl_api();
    if (nc == NULL)
        return -1;
    if (cm->begun_slot >= 0)
    {
        sb_set_error("gather: the previous _gather_begin has not been completed by _gather_end");
        return -1;
    }
    SB_DEVICE_CK(b->ctx->device);
    const int slot = b->slot;
    CK(cudaStreamWaitEvent(cm->stream, b->emitted[slot], 0));
    NK(nc->AllGather(b->d_total + slot*SB_META_WORDS, cm->d_counts[slot], SB_META_WORDS, ncclUint64, cm->nccl, cm->stream));
    CK(cudaMemcpyAsync(cm->h_counts[slot], cm->d_counts[slot], sizeof(unsigned long long)*SB_META_WORDS*cm->nranks, cudaMemcpyDeviceToHost,
                       cm->stream));
    CK(cudaEventRecord(cm->counts_ready[slot], cm->stream));
    cm->begun_slot = slot;
    return 0;
}

extern "C" int64_t span_b200_bank_gather_end(span_b200_bank_t *b, int64_t *counts)
{
    if (b == NULL  ||  b->comm == NULL  ||  b->comm->begun_slot < 0)
    {
        sb_set_error("gather: _gather_end without _gather_begin");
        return -1;
    }
    span_b200_comm_t *cm = b->comm;
    const nccl_api_t *nc = nccl_api();
    if (nc == NULL)
        return -1;
    SB_DEVICE_CK(b->ctx->device);
    const int slot = cm->begun_slot;
    CK(cudaEventSynchronize(cm->counts_ready[slot]));
    long long total = 0;
    const unsigned long long *meta = cm->h_counts[slot];
    const long long own_cap = (cm->rank == b->root)  ?  (cm->gather_cap[slot]/cm->nranks)  :  b->wire_cap[slot];
    for (int r = 0;  r < cm->nranks;  r++)
    {
        if (counts)
            counts[r] = (int64_t) meta[r*SB_META_WORDS];
        total += (long long) meta[r*SB_META_WORDS];
    }
    long long own = (long long) meta[cm->rank*SB_META_WORDS];
    if (own > own_cap)
    {
        sb_set_error("gather: this rank's record buffer overflowed (%lld records, room for %lld)", own, own_cap);
        return -1;
    }
    if (cm->transport == SPAN_B200_GATHER_PEER_COPY)
    {
        if (cm->rank == b->root)
        {
            if (comm_gather_reserve(cm, slot, total, b->last_stream, own) != 0)
                return -1;
            long long off = own;
            int pulls = 0;
            CK(cudaEventRecord(cm->fork, cm->stream));
            for (int r = 0;  r < cm->nranks;  r++)
            {
                const long long n = (long long) meta[r*SB_META_WORDS];
                if (r == cm->rank  ||  n == 0)
                    continue;
                const unsigned long long gen = meta[r*SB_META_WORDS + 1];
                if (gen == 0)
                {
                    sb_set_error("gather: rank %d exported no memory handle; use SPANDSP_B200_GATHER=nccl", r);
                    return -1;
                }
                if (cm->peer_gen[slot][r] != gen)
                {
                    // that rank (re-)allocated its buffer: map the new one
                    if (cm->peer_ptr[slot][r])
                        CK(cudaIpcCloseMemHandle(cm->peer_ptr[slot][r]));
                    cm->peer_ptr[slot][r] = NULL;
                    cudaIpcMemHandle_t h;
                    memcpy(&h, &meta[r*SB_META_WORDS + 2], sizeof(h));
                    CK(cudaIpcOpenMemHandle(&cm->peer_ptr[slot][r], h, cudaIpcMemLazyEnablePeerAccess));
                    cm->peer_gen[slot][r] = gen;
                }
                // the root's copy engine reads the rank's records over NVLink: exactly n records, behind the previous rank's
                cudaStream_t ps = cm->pull[pulls % SB_PULL_STREAMS];
                if (pulls < SB_PULL_STREAMS)
                    CK(cudaStreamWaitEvent(ps, cm->fork, 0));
                CK(cudaMemcpyAsync(cm->gather[slot] + off, cm->peer_ptr[slot][r], (size_t) n*sizeof(span_b200_wire_event_t),
                                   cudaMemcpyDefault, ps));
                pulls++;
                off += n;
            }
            for (int i = 0;  i < pulls  &&  i < SB_PULL_STREAMS;  i++)
            {
                CK(cudaEventRecord(cm->joined[i], cm->pull[i]));
                CK(cudaStreamWaitEvent(cm->stream, cm->joined[i], 0));
            }
        }
        // completion: a tiny collective behind the copies on the root's stream - a rank's buffer is free again when
        // this has completed on that rank, because the root takes part only after its copies have finished
        NK(nc->AllGather(cm->d_sync, cm->d_sync + 1, 1, ncclUint64, cm->nccl, cm->stream));
    }
    else if (cm->rank == b->root)
    {
        if (comm_gather_reserve(cm, slot, total, b->last_stream, own) != 0)
            return -1;
        long long off = own;
        NK(nc->GroupStart());
        for (int r = 0;  r < cm->nranks;  r++)
        {
            const long long n = (long long) meta[r*SB_META_WORDS];
            if (r == cm->rank  ||  n == 0)
                continue;
            NK(nc->Recv(cm->gather[slot] + off, (size_t) n*sizeof(span_b200_wire_event_t), ncclUint8, r, cm->nccl, cm->stream));
            off += n;
        }
        NK(nc->GroupEnd());
    }
    else if (own > 0)
    {
        NK(nc->GroupStart());
        NK(nc->Send(b->wire[slot], (size_t) own*sizeof(span_b200_wire_event_t), ncclUint8, b->root, cm->nccl, cm->stream));
        NK(nc->GroupEnd());
    }
    CK(cudaEventRecord(cm->done[slot], cm->stream));
    cm->done_valid[slot] = true;
    cm->total[slot] = total;
    cm->ended_slot = slot;
    cm->begun_slot = -1;
    return total;
}

extern "C" int64_t span_b200_bank_gathered(span_b200_bank_t *b, const span_b200_wire_event_t **d_records)
{
    if (b == NULL  ||  b->comm == NULL  ||  b->comm->ended_slot < 0)
    {
        sb_set_error("gather: nothing gathered yet");
        return -1;
    }
    span_b200_comm_t *cm = b->comm;
    SB_DEVICE_CK(b->ctx->device);
    const int slot = cm->ended_slot;
    CK(cudaEventSynchronize(cm->done[slot]));
    if (d_records)
        *d_records = (cm->rank == b->root)  ?  cm->gather[slot]  :  NULL;
    return cm->total[slot];
}

extern "C" int64_t span_b200_bank_gathered_host(span_b200_bank_t *b, span_b200_wire_event_t *out, int64_t max)
{
    const span_b200_wire_event_t *d = NULL;
    int64_t total = span_b200_bank_gathered(b, &d);
    if (total < 0)
        return -1;
    if (d == NULL)
        return 0;
    if (total > max)
        total = max;
    SB_DEVICE_CK(b->ctx->device);
    if (total > 0)
        CK(cudaMemcpy(out, d, sizeof(span_b200_wire_event_t)*(size_t) total, cudaMemcpyDeviceToHost));
    return total;
}
