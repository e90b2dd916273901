// sb_bank.cuh - the Goertzel filter-bank kernels.
//
// One thread owns one channel for a time slice.  Each thread keeps the 2*NPAIRS resonators of
// its channel in registers and streams the channel's int16 samples through them:
//
//      v1 = v2; v2 = v3; v3 = fac*v2 - v1 + x          (reference: src/spandsp/tone_detect.h:184-190)
//
// and at every block boundary pushes one zero sample, forms 2*(v3*v3 + v2*v2 - v2*v3*fac)
// (src/tone_detect.c:174-203), runs the detector's block decision and writes one code per
// (block, channel).  The sequential per-channel state machines (debounce, cadence, digit
// history) run afterwards in the sequencer kernels (sb_detectors.cuh) over those codes.  That
// split is what lets the time axis be cut into independent slices: a slice always starts on a
// block boundary, where the resonators are zero by definition.
//
// Memory path (bank_kernel_staged): input is channel-major [channel][sample] int16.  A warp
// (32 channels) cooperatively copies SEG_VEC*16 contiguous bytes of each of its 32 rows per
// stage with 16-byte cp.async (LDGSTS) - each group of SEG_VEC lanes reads one contiguous,
// 16-byte-aligned run of a row, so every DRAM sector that is fetched is fully used - into a
// per-warp shared-memory ring of NSTAGE stages.  Row r of the ring starts at r*ROW_BYTES with
// ROW_BYTES = 16 (mod 128), so when the 32 lanes then read "their" row with one 128-bit
// ld.shared per 8 samples, every quarter-warp touches 32 distinct banks (conflict-free).
// Warps are fully independent: no __syncthreads, only __syncwarp + cp.async groups.
//
// The generic kernel (bank_kernel_direct) handles what the staged path does not: per-channel
// block phases that differ inside a bank, and rows that are not 16-byte aligned.
#pragma once

#include <type_traits>

#include "sb_common.cuh"

namespace sb {

template <class DET>
struct BankArgs
{
    const int16_t *amp;             // [channel][sample], row stride in samples
    long long stride;
    int n;                          // samples per channel in this call
    int channels;                   // channels this launch covers (all pointers already point at its first channel)
    int cstride;                    // channels of the whole bank: the row pitch of the [row][channel] state / code arrays
    int blk0;                       // index of the first block decision this launch writes (0 unless a call is run in pieces)
    int cs0;                        // uniform block phase at entry (samples already in the open block)
    int slice_blocks;               // blocks per time slice (staged kernel)
    int nslices;
    int nblocks;                    // complete blocks produced per channel (uniform phase)
    float *v2;                      // carried resonator state, [2*NPAIRS][channels]
    float *v3;
    float *energy;                  // carried block energy, [channels]
    const int *cs;                  // block phase at entry, [channels] (direct kernel; the sequencer advances it)
    typename DET::code_t *code;     // [block][channel] decisions of this call
    float *eout;                    // [block][channel] block energy (only where DET::ENERGY_OUT)
    float *raw;                     // [block][bin][channel] bin energies (only where DET::RAW)
    long long raw_capacity;         // floats
    int block_rt;                   // block length where DET::BLOCK == 0 (raw Goertzel banks)
    const float *lut;               // 8-bit (G.711) input: 256-entry expansion table, as float
    typename DET::Params det;
};

template <class DET>
__device__ __forceinline__ int block_len(const BankArgs<DET> &a)
{
    return (DET::BLOCK > 0)  ?  DET::BLOCK  :  a.block_rt;
}

// ------------------------------------------------------------------------------------------
// Per-thread channel runner
template <class DET, int NPACK>
struct Runner
{
    static constexpr int NP = DET::NPAIRS;

    pair_t v2[NP];
    pair_t v3[NP];
    pair_t fac[NP];
    float energy;
    float z[4];                     // DTMF dial-tone notch state: z350[0], z350[1], z440[0], z440[1]
    int cs;                         // samples in the open block
    int blk;                        // index of the next block decision to write
    int c;                          // channel
    int channels;                   // row pitch of the [row][channel] arrays (BankArgs::cstride)
    bool active;
    bool filt;
    typename DET::Local loc;

    __device__ __forceinline__ void zero_state()
    {
#pragma unroll
        for (int p = 0;  p < NP;  p++)
        {
            v2[p].x = v2[p].y = 0.0f;
            v3[p].x = v3[p].y = 0.0f;
        }
        energy = 0.0f;
    }

    __device__ __forceinline__ void load_carry(const BankArgs<DET> &a)
    {
#pragma unroll
        for (int p = 0;  p < NP;  p++)
        {
            v2[p].x = a.v2[(size_t) (2*p)*channels + c];
            v2[p].y = a.v2[(size_t) (2*p + 1)*channels + c];
            v3[p].x = a.v3[(size_t) (2*p)*channels + c];
            v3[p].y = a.v3[(size_t) (2*p + 1)*channels + c];
        }
        energy = (DET::ENERGY)  ?  a.energy[c]  :  0.0f;
    }

    __device__ __forceinline__ void store_carry(const BankArgs<DET> &a)
    {
        if (!active)
            return;
#pragma unroll
        for (int p = 0;  p < NP;  p++)
        {
            a.v2[(size_t) (2*p)*channels + c] = v2[p].x;
            a.v2[(size_t) (2*p + 1)*channels + c] = v2[p].y;
            a.v3[(size_t) (2*p)*channels + c] = v3[p].x;
            a.v3[(size_t) (2*p + 1)*channels + c] = v3[p].y;
        }
        if (DET::ENERGY)
            a.energy[c] = energy;
        // The block phase cs[] is advanced by the sequencer kernel, which still needs the
        // entry value to reproduce the reference's duration arithmetic.
    }

    // src/dtmf.c:167-183.  Both biquads, strict left-to-right evaluation.
    __device__ __forceinline__ float notch(float famp)
    {
        float v1;

        v1 = fsub(fadd(fmul(0.98356f, famp), fmul(1.8954426f, z[0])), fmul(0.9691396f, z[1]));
        famp = fadd(fsub(v1, fmul(1.9251480f, z[0])), z[1]);
        z[1] = z[0];
        z[0] = v1;
        v1 = fsub(fadd(fmul(0.98456f, famp), fmul(1.8529543f, z[2])), fmul(0.9691396f, z[3]));
        famp = fadd(fsub(v1, fmul(1.8819938f, z[2])), z[3]);
        z[3] = z[2];
        z[2] = v1;
        return famp;
    }

    // The resonators of all bins, one sample
    __device__ __forceinline__ void bins(float x)
    {
        // The first NPACK pairs use the 2-wide add/sub (FADD2), the rest scalar FADDs: the mix is a
        // tuning knob (FADD2 saves issue slots but is not spread over the FP32 sub-pipes like FADD).
        // NPACK > NP: the multiply is 2-wide as well (FFMA2 with a zero addend, sb_common.cuh): 3 issue slots
        // per pair and sample instead of 4.
#pragma unroll
        for (int p = 0;  p < NP;  p++)
        {
            const pair_t v1 = v2[p];
            v2[p] = v3[p];
            if (NPACK > NP)
                v3[p] = padd_scalar<true>(psub<true>(pmul_packed(fac[p], v2[p]), v1), x);
            else if (p < NPACK)
                v3[p] = padd_scalar<true>(psub<true>(pmul(fac[p], v2[p]), v1), x);
            else
                v3[p] = padd_scalar<false>(psub<false>(pmul(fac[p], v2[p]), v1), x);
        }
    }

    template <bool FILT>
    __device__ __forceinline__ void step(float x)
    {
        if (DET::FILTER  &&  FILT)
        {
            const float f = notch(x);
            x = (filt)  ?  f  :  x;
        }
        if (DET::ENERGY)
            energy = fadd(energy, fmul(x, x));              // src/dtmf.c:189, super_tone_rx.c:477
        bins(x);
    }

    // End of a detection block: finish the bins, decide, emit, reset.
    __device__ __forceinline__ void block_end(const BankArgs<DET> &a)
    {
        float e[2*NP];

#pragma unroll
        for (int p = 0;  p < NP;  p++)
        {
            // src/tone_detect.c:174-203
            pair_t v1 = v2[p];
            pair_t w2 = v3[p];
            pair_t w3 = psub<false>(pmul(fac[p], w2), v1);
            e[2*p] = fmul(fsub(fadd(fmul(w3.x, w3.x), fmul(w2.x, w2.x)), fmul(fmul(w2.x, w3.x), fac[p].x)), 2.0f);
            e[2*p + 1] = fmul(fsub(fadd(fmul(w3.y, w3.y), fmul(w2.y, w2.y)), fmul(fmul(w2.y, w3.y), fac[p].y)), 2.0f);
        }
        if constexpr (DET::RAW)
        {
            if (active)
            {
#pragma unroll
                for (int i = 0;  i < 2*NP;  i++)
                {
                    const long long idx = ((long long) blk*loc.bins + i)*channels + c;
                    if (i < loc.bins  &&  idx < a.raw_capacity)
                        a.raw[idx] = e[i];
                }
                if (DET::ENERGY  &&  a.eout)
                    a.eout[(size_t) blk*channels + c] = energy;
            }
        }
        else
        {
            const int code = DET::decide(e, energy, loc);
            if (active)
            {
                a.code[(size_t) blk*channels + c] = (typename DET::code_t) code;
                if (DET::ENERGY_OUT  &&  code != 0)
                    a.eout[(size_t) blk*channels + c] = energy;
            }
        }
        blk++;
        cs = 0;
    }

    // Eight consecutive samples, no block boundary inside.
    template <bool FILT>
    __device__ __forceinline__ void fast8f(const float (&x)[8])
    {
        if constexpr (DET::ENERGY  &&  (NPACK > NP)  &&  !(DET::FILTER  &&  FILT))
        {
            // The squares for the block energy two at a time (a square is never -0, so pmul_packed is exact);
            // the sum itself stays sequential, in the reference's order
            float sq[8];
#pragma unroll
            for (int i = 0;  i < 8;  i += 2)
            {
                pair_t xx;
                xx.x = x[i];
                xx.y = x[i + 1];
                const pair_t q = pmul_packed(xx, xx);
                sq[i] = q.x;
                sq[i + 1] = q.y;
            }
#pragma unroll
            for (int i = 0;  i < 8;  i++)
            {
                energy = fadd(energy, sq[i]);
                bins(x[i]);
            }
        }
        else
        {
#pragma unroll
            for (int i = 0;  i < 8;  i++)
                step<FILT>(x[i]);
        }
    }

    template <bool FILT>
    __device__ __forceinline__ void fast8(const uint4 &v)
    {
        float x[8];
        x[0] = sample_of<0>(v);
        x[1] = sample_of<1>(v);
        x[2] = sample_of<2>(v);
        x[3] = sample_of<3>(v);
        x[4] = sample_of<4>(v);
        x[5] = sample_of<5>(v);
        x[6] = sample_of<6>(v);
        x[7] = sample_of<7>(v);
        fast8f<FILT>(x);
    }

    // Eight companded samples (two 32-bit words) through the expansion table in shared memory.
    template <bool FILT>
    __device__ __forceinline__ void fast8_g711(unsigned int w0, unsigned int w1, const float *lut)
    {
        float x[8];
#pragma unroll
        for (int i = 0;  i < 4;  i++)
        {
            x[i] = lut[(w0 >> (8*i)) & 0xFFu];
            x[4 + i] = lut[(w1 >> (8*i)) & 0xFFu];
        }
        fast8f<FILT>(x);
    }

    // What "reset" means at a block end: goertzel_result()/goertzel_reset() clear the bins
    // (src/tone_detect.c:110-120,203) and the detectors clear their energy sum.
    __device__ __forceinline__ void zero_after_block()
    {
#pragma unroll
        for (int p = 0;  p < NP;  p++)
        {
            v2[p].x = v2[p].y = 0.0f;
            v3[p].x = v3[p].y = 0.0f;
        }
        energy = 0.0f;
    }

    // Samples [lo, hi) of a vector, one at a time, with a block-boundary check after each.  Used
    // for the first/last vector of a slice and for the samples around a block boundary.
    // IN8: the vector holds 16 companded bytes instead of 8 int16.
    template <bool FILT, bool IN8>
    __device__ __forceinline__ void partial(uint4 v, int lo, int hi, const BankArgs<DET> &a, const float *lut)
    {
        const int B = block_len(a);
        if (IN8)
        {
            for (int i = 0;  i < lo;  i++)
                shift_vec8(v);
        }
        else
        {
            shift_vec_n(v, lo);
        }
#pragma unroll 1
        for (int e = lo;  e < hi;  e++)
        {
            float x;
            if (IN8)
            {
                x = lut[v.x & 0xFFu];
                shift_vec8(v);
            }
            else
            {
                x = (float) (short) (v.x & 0xFFFFu);
                shift_vec(v);
            }
            step<FILT>(x);
            if (++cs == B)
            {
                block_end(a);
                zero_after_block();
            }
        }
    }

    // Eight samples with a block boundary inside or right after them: k (1..8) samples finish the open block,
    // the other 8 - k start the next one.  k is warp-uniform in the staged kernel.  The tail of the old block is a
    // straight-line chain of steps left early after k of them; the head of the new block is the same chain
    // entered late (fall-through switch) - no per-sample loop, and the resonator registers rotate by renaming
    // except for one fix-up at the exit.  Needs B >= 8.
    template <bool FILT, int I>
    __device__ __forceinline__ void tail_chain(const float (&x)[8], int k)
    {
        step<FILT>(x[I]);
        if constexpr (I < 7)
        {
            if (k > I + 1)
                tail_chain<FILT, I + 1>(x, k);
        }
    }

    template <bool FILT>
    __device__ __forceinline__ void straddle8(const float (&x)[8], int k, const BankArgs<DET> &a)
    {
        tail_chain<FILT, 0>(x, k);
        block_end(a);
        zero_after_block();
        switch (k)
        {
        case 1:
            step<FILT>(x[1]);
            [[fallthrough]];
        case 2:
            step<FILT>(x[2]);
            [[fallthrough]];
        case 3:
            step<FILT>(x[3]);
            [[fallthrough]];
        case 4:
            step<FILT>(x[4]);
            [[fallthrough]];
        case 5:
            step<FILT>(x[5]);
            [[fallthrough]];
        case 6:
            step<FILT>(x[6]);
            [[fallthrough]];
        case 7:
            step<FILT>(x[7]);
            break;
        default:
            break;
        }
        cs = 8 - k;
    }

    template <bool FILT>
    __device__ __forceinline__ void straddle8_i16(const uint4 &v, int k, const BankArgs<DET> &a)
    {
        float x[8];
        x[0] = sample_of<0>(v);
        x[1] = sample_of<1>(v);
        x[2] = sample_of<2>(v);
        x[3] = sample_of<3>(v);
        x[4] = sample_of<4>(v);
        x[5] = sample_of<5>(v);
        x[6] = sample_of<6>(v);
        x[7] = sample_of<7>(v);
        straddle8<FILT>(x, k, a);
    }

    template <bool FILT>
    __device__ __forceinline__ void straddle8_g711(unsigned int w0, unsigned int w1, const float *lut, int k, const BankArgs<DET> &a)
    {
        float x[8];
#pragma unroll
        for (int i = 0;  i < 4;  i++)
        {
            x[i] = lut[(w0 >> (8*i)) & 0xFFu];
            x[4 + i] = lut[(w1 >> (8*i)) & 0xFFu];
        }
        straddle8<FILT>(x, k, a);
    }

    // A whole vector.  The common case - no block boundary inside a group of eight samples - takes the
    // unrolled path.  All conditions are warp-uniform in the staged kernel.
    template <bool FILT, bool IN8>
    __device__ __forceinline__ void full(const uint4 &v, const BankArgs<DET> &a, const float *lut)
    {
        const int B = block_len(a);
        if (!IN8)
        {
            if (cs + 8 < B)
            {
                fast8<FILT>(v);
                cs += 8;
            }
            else if (DET::BLOCK > 0  ||  B >= 8)
            {
                straddle8_i16<FILT>(v, B - cs, a);
            }
            else
            {
                partial<FILT, false>(v, 0, 8, a, lut);
            }
        }
        else
        {
            // two groups of eight companded samples
#pragma unroll 1
            for (int h = 0;  h < 2;  h++)
            {
                const unsigned int w0 = (h)  ?  v.z  :  v.x;
                const unsigned int w1 = (h)  ?  v.w  :  v.y;
                if (cs + 8 < B)
                {
                    fast8_g711<FILT>(w0, w1, lut);
                    cs += 8;
                }
                else if (DET::BLOCK > 0  ||  B >= 8)
                {
                    straddle8_g711<FILT>(w0, w1, lut, B - cs, a);
                }
                else
                {
                    partial<FILT, true>(v, 8*h, 8*h + 8, a, lut);
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// Staged kernel: uniform block phase across the bank, 16-byte aligned rows.
template <int SEG_VEC, int NSTAGE>
struct StageCfg
{
    static constexpr int ROW_BYTES = NSTAGE*SEG_VEC*16 + 16;
    static constexpr int WARP_BYTES = 32*ROW_BYTES;
    static_assert((NSTAGE*SEG_VEC) % 8 == 0, "ring must be a multiple of 128 bytes so that ROW_BYTES = 16 mod 128");
    static_assert(SEG_VEC == 8  ||  SEG_VEC == 16  ||  SEG_VEC == 32, "SEG_VEC lanes copy one row segment");
};

// FILTK: this instantiation carries the DTMF dial-tone notch (used when any channel of the bank has it on).
template <class DET, int SEG_VEC, int NSTAGE, int WARPS, int MINB, int NPACK, bool IN8, bool FILTK>
__global__ void __launch_bounds__(WARPS*32, MINB) bank_kernel_staged(const BankArgs<DET> a)
{
    constexpr int VSH = (IN8)  ?  4  :  3;          // log2(samples per 16-byte vector)
    constexpr int VMASK = (1 << VSH) - 1;
    constexpr int BPS = (IN8)  ?  1  :  2;          // bytes per sample
    typedef StageCfg<SEG_VEC, NSTAGE> cfg;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    // The broadcast tells the compiler that the warp index - and with it the slice geometry and the block
    // phase - is warp-uniform: uniform registers and plain branches instead of divergence bookkeeping.
    const int warp = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0);
    // 8-bit input: the 256-entry expansion table sits in front of the rings (1 KB per CTA)
    const float *lut = (const float *) smem_raw;
    if (IN8)
    {
        float *w = (float *) smem_raw;
        for (int i = threadIdx.x;  i < 256;  i += WARPS*32)
            w[i] = a.lut[i];
        __syncthreads();
    }
    const int ngroups = (a.channels + 31) >> 5;
    const long long item = (long long) blockIdx.x*WARPS + warp;
    const int group = (int) (item % ngroups);
    const int slice = (int) (item / ngroups);
    if (slice >= a.nslices)
        return;

    // ---- sample range of this slice (block geometry: see DESIGN.md "time slicing") ----
    const int B = block_len(a);
    const bool last = (slice == a.nslices - 1);
    long long s0 = (long long) slice*a.slice_blocks*B - a.cs0;
    if (s0 < 0)
        s0 = 0;
    long long s1 = (last)  ?  (long long) a.n  :  ((long long) (slice + 1)*a.slice_blocks*B - a.cs0);
    if (s1 > a.n)
        s1 = a.n;
    const int start = (int) s0;
    const int end = (int) s1;

    Runner<DET, NPACK> r;
    r.channels = a.cstride;
    r.c = group*32 + lane;
    r.active = (r.c < a.channels);
    if (!r.active)
        r.c = a.channels - 1;           // shadow the last channel; never writes
    DET::load_local(a.det, r.c, r.loc);
#pragma unroll
    for (int p = 0;  p < DET::NPAIRS;  p++)
    {
        r.fac[p] = DET::fac(a.det, p);
        // Keep the coefficients in registers: without this the compiler re-materialises them from
        // the constant bank (LDCU + MOV) in every unrolled vector, 10 extra issue slots per 8 samples.
        asm volatile("" : "+f"(r.fac[p].x), "+f"(r.fac[p].y));
    }
    r.blk = a.blk0 + slice*a.slice_blocks;
    r.filt = false;
    if (slice == 0  &&  a.cs0 > 0)
    {
        r.load_carry(a);
        r.cs = a.cs0;
    }
    else
    {
        r.zero_state();
        r.cs = 0;
    }
    constexpr bool any_filter = DET::FILTER  &&  FILTK;
    if (any_filter)
    {
        r.filt = DET::filter_on(a.det, r.c);
        DET::load_filter(a.det, r.c, a.cstride, r.z);
    }

    // ---- staging geometry ----
    const uint32_t warp_smem = (uint32_t) __cvta_generic_to_shared(smem_raw) + ((IN8)  ?  1024  :  0) + warp*cfg::WARP_BYTES;
    const uint32_t my_row = warp_smem + lane*cfg::ROW_BYTES;
    const int v_lo = start >> VSH;
    const int v_hi = (end + VMASK) >> VSH;
    const int g_lo = v_lo/SEG_VEC;
    const int g_hi = (v_hi + SEG_VEC - 1)/SEG_VEC;
    const long long row_bytes = (long long) a.n*BPS;

    // copy role of this lane: SEG_VEC lanes cover one row segment; 32/SEG_VEC rows per instruction
    constexpr int RPI = 32/SEG_VEC;
    const int cp_vec = lane % SEG_VEC;
    const int cp_row0 = lane/SEG_VEC;

    // Row pointer of this lane's first copy and the byte step between its successive copies.
    const bool full_group = (group*32 + 32 <= a.channels);
    const char *cp_src0 = (const char *) a.amp + (long long) (group*32 + cp_row0)*a.stride*BPS + (long long) cp_vec*16;
    const long long cp_step = (long long) RPI*a.stride*BPS;
    const uint32_t cp_dst0 = warp_smem + cp_row0*cfg::ROW_BYTES + cp_vec*16;

    auto issue = [&](int g)
    {
        if (g < g_hi)
        {
            const long long off = (long long) g*(SEG_VEC*16);
            uint32_t dst = cp_dst0 + (g % NSTAGE)*(SEG_VEC*16);
            // Interior segment of a full group of rows: every 16-byte piece lies inside the slice and the row
            const bool interior = full_group  &&  g*SEG_VEC >= v_lo  &&  (g + 1)*SEG_VEC <= v_hi  &&  off + SEG_VEC*16 <= row_bytes;
            if (interior)
            {
                const char *src = cp_src0 + off;
#pragma unroll
                for (int i = 0;  i < SEG_VEC;  i++)
                {
                    cp_async_16_full(dst, src);
                    src += cp_step;
                    dst += RPI*cfg::ROW_BYTES;
                }
            }
            else
            {
                const int V = g*SEG_VEC + cp_vec;
                long long nbytes = row_bytes - (off + cp_vec*16);
                int sb = (nbytes >= 16)  ?  16  :  (nbytes > 0)  ?  (int) nbytes  :  0;
                if (V < v_lo  ||  V >= v_hi)
                    sb = 0;
#pragma unroll 1
                for (int i = 0;  i < SEG_VEC;  i++)
                {
                    int ch = group*32 + i*RPI + cp_row0;
                    if (ch >= a.channels)
                        ch = a.channels - 1;
                    const char *src = (const char *) a.amp + (long long) ch*a.stride*BPS + off + cp_vec*16;
                    cp_async_16(dst, src, sb);
                    dst += RPI*cfg::ROW_BYTES;
                }
            }
        }
        cp_async_commit();
    };

    for (int s = 0;  s < NSTAGE - 1;  s++)
        issue(g_lo + s);

    // Vector roles inside the slice.  A slice starts on a block boundary with all-zero resonators, for which
    // zero samples are exact no-ops: the leading samples of its first vector (they belong to the previous
    // slice) are masked to zero and the block phase starts negative, so the first vector runs through full()
    // like any other.  A slice that is not the last one ends on a block boundary: what full() computes past
    // it (the head of a block that belongs to the next slice) is never emitted or stored.  Only the last
    // vector of the last slice must stop exactly at `end`: that one takes partial().  (8-bit input keeps
    // partial() for the head too: A-law has no code for zero.)
    const int lead = start & VMASK;
    const bool head_partial = (IN8  &&  lead != 0);
    const bool tail_exact = (last  &&  (end & VMASK) != 0);
    const int vf_lo = (head_partial)  ?  (v_lo + 1)  :  v_lo;          // vectors [vf_lo, vf_hi) go through full()
    const int vf_hi = (tail_exact)  ?  (v_hi - 1)  :  v_hi;
    constexpr bool FILT = any_filter;
    for (int g = g_lo;  g < g_hi;  g++)
    {
        cp_async_wait<NSTAGE - 2>();
        __syncwarp();
        issue(g + NSTAGE - 1);
        const uint32_t stage = my_row + (g % NSTAGE)*SEG_VEC*16;
        const int Vbase = g*SEG_VEC;
        if (head_partial  &&  v_lo >= Vbase  &&  v_lo < Vbase + SEG_VEC)
        {
            const int hi_h = (v_hi - 1 == v_lo)  ?  (((end - 1) & VMASK) + 1)  :  (VMASK + 1);
            r.template partial<FILT, IN8>(lds128(stage + (v_lo - Vbase)*16), lead, hi_h, a, lut);
        }
        const int ja = (vf_lo > Vbase)  ?  (vf_lo - Vbase)  :  0;
        const int jb = (vf_hi < Vbase + SEG_VEC)  ?  (vf_hi - Vbase)  :  SEG_VEC;
        if (ja < jb)
        {
            uint4 cur = lds128(stage + ja*16);
            if (!IN8  &&  lead != 0  &&  Vbase + ja == v_lo)
            {
                // mask the `lead` samples that precede the slice
                cur.x = (lead >= 2)  ?  0u  :  (cur.x & 0xFFFF0000u);
                cur.y = (lead >= 4)  ?  0u  :  (lead == 3)  ?  (cur.y & 0xFFFF0000u)  :  cur.y;
                cur.z = (lead >= 6)  ?  0u  :  (lead == 5)  ?  (cur.z & 0xFFFF0000u)  :  cur.z;
                cur.w = (lead == 7)  ?  (cur.w & 0xFFFF0000u)  :  cur.w;
                r.cs = -lead;
            }
            // Two vectors per trip, each fetched while the previous one is processed (the fetch after the last
            // vector of a stage lands in the 16 bytes of padding that follow the ring, or in the next stage;
            // its value is not used).  Ping-pong between the two register sets: no copies.
            int j = ja;
#pragma unroll 1
            for (;;)
            {
                const uint4 nxt = lds128(stage + j*16 + 16);
                r.template full<FILT, IN8>(cur, a, lut);
                if (++j >= jb)
                    break;
                cur = lds128(stage + j*16 + 16);
                r.template full<FILT, IN8>(nxt, a, lut);
                if (++j >= jb)
                    break;
            }
        }
        if (tail_exact  &&  v_hi - 1 >= Vbase  &&  v_hi - 1 < Vbase + SEG_VEC  &&  !(head_partial  &&  v_hi - 1 == v_lo))
        {
            const int lo_t = (v_hi - 1 == v_lo)  ?  lead  :  0;
            r.template partial<FILT, IN8>(lds128(stage + (v_hi - 1 - Vbase)*16), lo_t, ((end - 1) & VMASK) + 1, a, lut);
        }
    }
    cp_async_wait<0>();

    if (last  &&  !DET::RAW)
    {
        r.store_carry(a);
        if (DET::FILTER  &&  any_filter  &&  r.active)
            DET::store_filter(a.det, r.c, a.cstride, r.z);
    }
}

// ------------------------------------------------------------------------------------------
// Direct kernel: any block phase per channel, any alignment.  One thread per channel over the
// whole call; samples are read with 16-bit loads through L1.
template <class DET, int NPACK, bool IN8>
__global__ void __launch_bounds__(128) bank_kernel_direct(const BankArgs<DET> a)
{
    const int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= a.channels)
        return;

    Runner<DET, NPACK> r;
    r.channels = a.cstride;
    r.c = c;
    r.active = true;
    DET::load_local(a.det, c, r.loc);
#pragma unroll
    for (int p = 0;  p < DET::NPAIRS;  p++)
        r.fac[p] = DET::fac(a.det, p);
    r.blk = a.blk0;
    r.cs = (DET::RAW)  ?  0  :  a.cs[c];
    if (r.cs > 0)
        r.load_carry(a);
    else
        r.zero_state();
    r.filt = false;
    if (DET::FILTER)
    {
        r.filt = DET::filter_on(a.det, c);
        DET::load_filter(a.det, c, a.cstride, r.z);
    }
    const int16_t *row = a.amp + (long long) c*a.stride;
    const unsigned char *row8 = (const unsigned char *) a.amp + (long long) c*a.stride;
    if (DET::FILTER  &&  r.filt)
    {
#pragma unroll 1
        for (int i = 0;  i < a.n;  i++)
        {
            r.template step<true>((IN8)  ?  __ldg(a.lut + __ldg(row8 + i))  :  (float) __ldg(row + i));
            if (++r.cs == block_len(a))
            {
                r.block_end(a);
                r.zero_after_block();
            }
        }
        DET::store_filter(a.det, c, a.cstride, r.z);
    }
    else
    {
#pragma unroll 1
        for (int i = 0;  i < a.n;  i++)
        {
            r.template step<false>((IN8)  ?  __ldg(a.lut + __ldg(row8 + i))  :  (float) __ldg(row + i));
            if (++r.cs == block_len(a))
            {
                r.block_end(a);
                r.zero_after_block();
            }
        }
    }
    if (!DET::RAW)
        r.store_carry(a);
}

}  // namespace sb
