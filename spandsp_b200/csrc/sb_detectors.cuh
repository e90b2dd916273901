// sb_detectors.cuh - per-detector block decisions (device) and the per-channel sequencers.
//
// Block decisions run inside the bank kernel, once per (channel, block); they are pure
// functions of the block's bin energies.  Sequencers are the reference's small per-channel
// state machines (debounce, hit history, cadence tracking); they run as a second, tiny pass
// over the [block][channel] decision codes and turn them into callback-equivalent event
// records.  Events are produced with a count -> exclusive-scan -> emit scheme, so the event
// buffer is compact, ordered by (channel, time) and written without atomics.
#pragma once

#include <limits.h>

#include "sb_bank.cuh"
#include "../../include/spandsp_b200.h"

namespace sb {


struct SeqCommon
{
    int channels;
    int n;                      // samples per channel in this call
    int cs0;                    // uniform entry phase, or -1: per channel from cs[]
    int *cs;                    // [channels] block phase, advanced by the emit pass
    const unsigned int *offsets;    // exclusive scan of the per-warp counts (emit pass)
    unsigned int *counts;       // per-warp (32 channels) event counts (count pass)
    span_b200_event_t *events;  // 24-byte records, or
    span_b200_wire_event_t *wire;   // 12-byte wire records (multi-GPU gather, compact host read-back); one of the two is NULL
    unsigned int channel_base;  // wire records carry the global channel number
    long long capacity;
};

__device__ __forceinline__ void put_event(const SeqCommon &q, unsigned int pos, int c, int blk, int kind, int a, int b, int cc)
{
    if ((long long) pos < q.capacity)
    {
        if (q.wire)
        {
            // kind 1 / 2 / 5 -> 1 / 2 / 3 in the top two bits of block_kind (include/spandsp_b200.h)
            span_b200_wire_event_t w;
            w.channel = q.channel_base + (unsigned int) c;
            w.c = cc;
            w.block_kind = (unsigned short) ((blk & 0x3FFF) | (((kind == SPAN_B200_EV_SEGMENT)  ?  3  :  kind) << 14));
            w.a = (signed char) a;
            w.b = (signed char) b;
            q.wire[pos] = w;
        }
        else
        {
            span_b200_event_t e;
            e.channel = c;
            e.block = blk;
            e.kind = kind;
            e.a = a;
            e.b = b;
            e.c = cc;
            q.events[pos] = e;
        }
    }
}

// Fill in field b of a record written earlier (the deferred DTMF level)
__device__ __forceinline__ void set_event_b(const SeqCommon &q, unsigned int pos, int b)
{
    if ((long long) pos < q.capacity)
    {
        if (q.wire)
            q.wire[pos].b = (signed char) b;
        else
            q.events[pos].b = b;
    }
}

// Warp-aggregated event output.  The 32 channels of a warp walk their blocks in lock step; at
// every emission point the lanes that have an event take consecutive slots of the warp's region
// of the event buffer (ballot + popc), so the records of one step land in one contiguous run
// instead of 32 scattered ones.  Order inside the buffer: (group of 32 channels, time, channel);
// the order of any single channel's events is preserved.  All lanes must call push() together.
template <bool EMIT>
struct EventSink
{
    const SeqCommon &q;
    unsigned int base;
    unsigned int lt;

    int cpw;                        // channels per warp (32, or fewer where a sequencer spreads a bank over more warps)

    __device__ __forceinline__ EventSink(const SeqCommon &qq, int warp_global, int channels_per_warp = 32) : q(qq)
    {
        cpw = channels_per_warp;
        // (a CTA's trailing warps may lie wholly beyond the last channel: they have no slot in offsets[] / counts[])
        base = (EMIT  &&  warp_global*cpw < qq.channels)  ?  qq.offsets[warp_global]  :  0;
        lt = (1u << (threadIdx.x & 31)) - 1u;
    }

    // Returns the position the event was written to (emit pass, has == true), else 0
    __device__ __forceinline__ unsigned int push(bool has, int c, int blk, int kind, int a, int b, int cc)
    {
        unsigned int pos = 0;
        if (EMIT)
        {
            const unsigned int m = __ballot_sync(0xFFFFFFFFu, has);
            pos = base + __popc(m & lt);
            if (has)
                put_event(q, pos, c, blk, kind, a, b, cc);
            base += __popc(m);
        }
        else
        {
            base += (has)  ?  1u  :  0u;        // count pass: per lane, reduced once at the end
        }
        return pos;
    }

    __device__ __forceinline__ void finish(int warp_global)
    {
        if (!EMIT)
        {
            const unsigned int total = __reduce_add_sync(0xFFFFFFFFu, base);
            if ((threadIdx.x & 31) == 0  &&  warp_global*cpw < q.channels)
                q.counts[warp_global] = total;
        }
    }
};

// ==========================================================================================
// DTMF  (reference: src/dtmf.c)
__constant__ float c_dtmf_fac[8];       // row0, col0, row1, col1, ... (src/dtmf.c:114-121)
__constant__ char c_dtmf_positions[17]; // "123A456B789C*0#D" (src/dtmf.c:123)

struct DtmfParams
{
    const float *threshold;             // per channel (dtmf_rx_parms, src/dtmf.c:421-445)
    const float *normal_twist;
    const float *reverse_twist;
    const unsigned char *flags;         // SB_DTMF_FLAG_*
    float *z;                           // [4][channels] dial-tone notch state
};

#define SB_DTMF_FLAG_FILTER     1
#define SB_DTMF_FLAG_REALTIME   2

struct DtmfLocal
{
    float threshold;
    float normal_twist;
    float reverse_twist;
};

struct DtmfDet
{
    static constexpr int NPAIRS = 4;
    static constexpr int BLOCK = 102;               // src/dtmf.c:71
    static constexpr bool ENERGY = true;
    static constexpr bool ENERGY_OUT = true;
    static constexpr bool FILTER = true;
    static constexpr bool RAW = false;
    typedef unsigned char code_t;
    typedef DtmfParams Params;
    typedef DtmfLocal Local;

    __device__ static __forceinline__ pair_t fac(const Params &, int p)
    {
        pair_t f;
        f.x = c_dtmf_fac[2*p];
        f.y = c_dtmf_fac[2*p + 1];
        return f;
    }

    __device__ static __forceinline__ void load_local(const Params &d, int c, Local &l)
    {
        l.threshold = d.threshold[c];
        l.normal_twist = d.normal_twist[c];
        l.reverse_twist = d.reverse_twist[c];
    }

    __device__ static __forceinline__ bool filter_on(const Params &d, int c)
    {
        return (d.flags[c] & SB_DTMF_FLAG_FILTER) != 0;
    }

    __device__ static __forceinline__ void load_filter(const Params &d, int c, int channels, float (&z)[4])
    {
#pragma unroll
        for (int i = 0;  i < 4;  i++)
            z[i] = d.z[(size_t) i*channels + c];
    }

    __device__ static __forceinline__ void store_filter(const Params &d, int c, int channels, const float (&z)[4])
    {
#pragma unroll
        for (int i = 0;  i < 4;  i++)
            d.z[(size_t) i*channels + c] = z[i];
    }

    // src/dtmf.c:211-258.  e[2i] = row i, e[2i+1] = col i.
    __device__ static __forceinline__ int decide(const float (&e)[8], float energy, const Local &l)
    {
        float rbest = e[0];
        float cbest = e[1];
        int best_row = 0;
        int best_col = 0;
#pragma unroll
        for (int i = 1;  i < 4;  i++)
        {
            if (e[2*i] > rbest)
            {
                rbest = e[2*i];
                best_row = i;
            }
            if (e[2*i + 1] > cbest)
            {
                cbest = e[2*i + 1];
                best_col = i;
            }
        }
        bool ok = (rbest >= l.threshold)  &&  (cbest >= l.threshold);
        ok = ok  &&  (cbest < fmul(rbest, l.reverse_twist))  &&  (fmul(cbest, l.normal_twist) > rbest);
#pragma unroll
        for (int i = 0;  i < 4;  i++)
        {
            if ((i != best_col  &&  fmul(e[2*i + 1], 6.309f) > cbest)
                ||
                (i != best_row  &&  fmul(e[2*i], 6.309f) > rbest))
                ok = false;
        }
        ok = ok  &&  (fadd(rbest, cbest) > fmul(83.868f, energy));
        const int ch = (int) c_dtmf_positions[(best_row << 2) + best_col];        // src/dtmf.c:123,256
        return (ok)  ?  ch  :  0;
    }
};

struct DtmfSeqArgs
{
    SeqCommon q;
    const unsigned char *code;          // [block][channel]
    const float *eout;                  // [block][channel]
    const unsigned char *flags;
    unsigned char *last_hit;
    unsigned char *in_digit;
    int *duration;
    // level(energy) = (int) (10*log10f(energy) - 107.255f) (src/dtmf.c:110,314) is a monotone
    // step function of the float energy.  The host tabulates its steps once with its own libm
    // (the same one the CPU oracle uses) so that the device reproduces it exactly:
    // level = level_min + #{k : level_thr[k] <= energy}.
    const float *level_thr;
    int level_n;
    int level_min;
};

__device__ __forceinline__ int dtmf_level(const DtmfSeqArgs &s, float energy)
{
    int lo = 0;
    int hi = s.level_n;
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (s.level_thr[mid] <= energy)
            lo = mid + 1;
        else
            hi = mid;
    }
    return s.level_min + lo;
}

// One block decision through the debounce / duration logic of src/dtmf.c:201-207,304-347
// DEFER: the level of a "tone on" report is not computed here (it needs the block energy from global memory and a
// table search, neither of which the state machine depends on); need_level tells the caller to fill it in later.
template <bool EMIT, bool DEFER>
__device__ __forceinline__ void dtmf_seq_block(const DtmfSeqArgs &s, int c, int b, int len, int hit, bool realtime,
                                               int &in_digit, int &last_hit, int &dur,
                                               bool &ev, int &ev_kind, int &ev_a, int &ev_b, int &ev_c, bool &need_level)
{
    if (dur < INT_MAX - len)
        dur += len;
    if (hit != in_digit  &&  last_hit != in_digit)
    {
        hit = (hit  &&  hit == last_hit)  ?  hit  :  0;
        if (realtime)
        {
            if (in_digit  ||  hit)
            {
                ev = true;
                ev_kind = SPAN_B200_EV_TONE;
                ev_a = hit;
                ev_c = dur;
                if (in_digit  &&  !hit)
                    ev_b = -99;
                else if (EMIT  &&  DEFER)
                    need_level = true;
                else if (EMIT)
                    ev_b = dtmf_level(s, s.eout[(size_t) b*s.q.channels + c]);
                dur = 0;
            }
        }
        else if (hit)
        {
            ev = true;
            ev_kind = SPAN_B200_EV_DIGIT;
            ev_a = hit;
        }
        in_digit = hit;
    }
    last_hit = hit;
}

#define SB_SEQ_TILE     112             // block rows staged per tile (cfg2: 784 blocks = 7 tiles)
#define SB_SEQ_LEVELS   1024            // level table entries kept in shared memory (the table has ~830)
#define SB_SEQ_QUEUE    16              // deferred levels per thread and tile

// dtmf_level() over a copy of the table in shared memory
__device__ __forceinline__ int dtmf_level_tab(const float *tab, int n, int level_min, float energy)
{
    int lo = 0;
    int hi = n;
    while (lo < hi)
    {
        const int mid = (lo + hi) >> 1;
        if (tab[mid] <= energy)
            lo = mid + 1;
        else
            hi = mid;
    }
    return level_min + lo;
}

// The per-channel sequencer is a short serial state machine over the block decisions; left to itself it is bound
// by the latency of its one-byte loads (the round-1 version took 0.25 ms per pass for 51 MB of codes).  Where the
// bank has one block phase and a channel count that keeps the rows 16-byte aligned, the CTA stages the decisions
// of its 128 channels through shared memory in tiles of SB_SEQ_TILE block rows (16-byte cp.async, double
// buffered): a handful of memory round trips per CTA instead of one or more per block.
template <bool EMIT>
__global__ void __launch_bounds__(128) dtmf_sequencer(const DtmfSeqArgs s)
{
    __shared__ __align__(16) unsigned char tile[2][SB_SEQ_TILE*128];
    // Emit pass: the level table, and per thread a short queue of "tone on" reports whose level is still owed
    __shared__ float lvl_tab[(EMIT)  ?  SB_SEQ_LEVELS  :  1];
    __shared__ unsigned int q_pos[(EMIT)  ?  SB_SEQ_QUEUE  :  1][128];
    __shared__ unsigned short q_row[(EMIT)  ?  SB_SEQ_QUEUE  :  1][128];
    const int gc = blockIdx.x*blockDim.x + threadIdx.x;
    const int wg = gc >> 5;
    const bool live = (gc < s.q.channels);
    const int c = (live)  ?  gc  :  (s.q.channels - 1);
    const int B = DtmfDet::BLOCK;
    const int cs_old = (s.q.cs0 >= 0)  ?  s.q.cs0  :  s.q.cs[c];
    const int nb = (live)  ?  ((cs_old + s.q.n)/B)  :  0;
    const bool realtime = (s.flags[c] & SB_DTMF_FLAG_REALTIME) != 0;
    int in_digit = s.in_digit[c];
    int last_hit = s.last_hit[c];
    int dur = s.duration[c];
    EventSink<EMIT> sink(s.q, wg);

    if (s.q.cs0 >= 0  &&  (s.q.channels & 15) == 0  &&  s.level_n <= SB_SEQ_LEVELS)
    {
        if (EMIT)
        {
            for (int i = threadIdx.x;  i < s.level_n;  i += 128)
                lvl_tab[i] = s.level_thr[i];                // visible after the first __syncthreads() below
        }
        int qn = 0;
        const int nbu = (s.q.cs0 + s.q.n)/B;                // the same for every channel
        const int ntiles = (nbu + SB_SEQ_TILE - 1)/SB_SEQ_TILE;
        const int c0 = blockIdx.x*128;
        auto issue = [&](int t)
        {
            if (t < ntiles)
            {
                const uint32_t dst0 = (uint32_t) __cvta_generic_to_shared(tile[t & 1]);
                for (int k = threadIdx.x;  k < SB_SEQ_TILE*8;  k += 128)
                {
                    const int row = k >> 3;
                    const int piece = k & 7;
                    const int b = t*SB_SEQ_TILE + row;
                    const bool ok = (b < nbu  &&  c0 + 16*piece < s.q.channels);
                    const unsigned char *src = (ok)  ?  (s.code + (size_t) b*s.q.channels + c0 + 16*piece)  :  s.code;
                    cp_async_16(dst0 + row*128 + piece*16, src, (ok)  ?  16  :  0);
                }
            }
            cp_async_commit();
        };
        issue(0);
        for (int t = 0;  t < ntiles;  t++)
        {
            issue(t + 1);
            cp_async_wait<1>();
            __syncthreads();
            const unsigned char *col = tile[t & 1] + threadIdx.x;
            const int rows = (nbu - t*SB_SEQ_TILE < SB_SEQ_TILE)  ?  (nbu - t*SB_SEQ_TILE)  :  SB_SEQ_TILE;
#pragma unroll 4
            for (int row = 0;  row < rows;  row++)
            {
                const int b = t*SB_SEQ_TILE + row;
                bool ev = false;
                int ev_kind = 0;
                int ev_a = 0;
                int ev_b = 0;
                int ev_c = 0;
                bool need_level = false;
                if (live)
                    dtmf_seq_block<EMIT, true>(s, c, b, (b == 0)  ?  (B - cs_old)  :  B, col[row*128], realtime, in_digit, last_hit, dur,
                                               ev, ev_kind, ev_a, ev_b, ev_c, need_level);
                const unsigned int pos = sink.push(ev, c, b, ev_kind, ev_a, ev_b, ev_c);
                if (EMIT  &&  need_level)
                {
                    if (qn < SB_SEQ_QUEUE)
                    {
                        q_pos[qn][threadIdx.x] = pos;
                        q_row[qn][threadIdx.x] = (unsigned short) row;
                        qn++;
                    }
                    else
                    {
                        set_event_b(s.q, pos, dtmf_level_tab(lvl_tab, s.level_n, s.level_min, s.eout[(size_t) b*s.q.channels + c]));
                    }
                }
            }
            if (EMIT)
            {
                // The levels owed for this tile: all energy loads first (independent, one memory round trip), then
                // the table searches and the 4-byte stores into the records written above
                float en[SB_SEQ_QUEUE];
#pragma unroll
                for (int i = 0;  i < SB_SEQ_QUEUE;  i++)
                    en[i] = (i < qn)  ?  s.eout[(size_t) (t*SB_SEQ_TILE + q_row[i][threadIdx.x])*s.q.channels + c]  :  0.0f;
#pragma unroll
                for (int i = 0;  i < SB_SEQ_QUEUE;  i++)
                {
                    if (i < qn)
                    {
                        set_event_b(s.q, q_pos[i][threadIdx.x], dtmf_level_tab(lvl_tab, s.level_n, s.level_min, en[i]));
                    }
                }
                qn = 0;
            }
            __syncthreads();
        }
        cp_async_wait<0>();
    }
    else
    {
        const int nb_max = __reduce_max_sync(0xFFFFFFFFu, nb);
        // The decision codes of a group of blocks are fetched together: the loads do not depend on the
        // state machine, and issuing them back to back hides the memory latency a one-by-one walk exposes.
        constexpr int G = 16;
        for (int b0 = 0;  b0 < nb_max;  b0 += G)
        {
            unsigned char codes[G];
#pragma unroll
            for (int i = 0;  i < G;  i++)
                codes[i] = (b0 + i < nb)  ?  s.code[(size_t) (b0 + i)*s.q.channels + c]  :  (unsigned char) 0;
#pragma unroll
            for (int i = 0;  i < G;  i++)
            {
                const int b = b0 + i;
                bool ev = false;
                int ev_kind = 0;
                int ev_a = 0;
                int ev_b = 0;
                int ev_c = 0;
                bool need_level = false;
                if (b < nb)
                    dtmf_seq_block<EMIT, false>(s, c, b, (b == 0)  ?  (B - cs_old)  :  B, codes[i], realtime, in_digit, last_hit, dur,
                                                ev, ev_kind, ev_a, ev_b, ev_c, need_level);
                if (b < nb_max)
                    sink.push(ev, c, b, ev_kind, ev_a, ev_b, ev_c);
            }
        }
    }
    sink.finish(wg);
    if (EMIT  &&  live)
    {
        const int tail = (nb == 0)  ?  s.q.n  :  (cs_old + s.q.n - nb*B);
        if (dur < INT_MAX - tail)
            dur += tail;
        s.in_digit[c] = (unsigned char) in_digit;
        s.last_hit[c] = (unsigned char) last_hit;
        s.duration[c] = dur;
        s.q.cs[c] = cs_old + s.q.n - nb*B;
    }
}

// The staged walk of dtmf_sequencer as a helper for the detectors without a per-tile epilogue: the CTA's 128 channels
// step through all nbu block rows of 8-bit decision codes, tile by tile (16-byte cp.async, double buffered); f(block,
// code) is called by all threads together for every row.  Needs a uniform block phase and channels % 16 == 0.
template <class F>
__device__ __forceinline__ void walk_codes8(const unsigned char *code, int channels, int nbu, unsigned char (*tile)[SB_SEQ_TILE*128], F f)
{
    const int ntiles = (nbu + SB_SEQ_TILE - 1)/SB_SEQ_TILE;
    const int c0 = blockIdx.x*128;
    auto issue = [&](int t)
    {
        if (t < ntiles)
        {
            const uint32_t dst0 = (uint32_t) __cvta_generic_to_shared(tile[t & 1]);
            for (int k = threadIdx.x;  k < SB_SEQ_TILE*8;  k += 128)
            {
                const int row = k >> 3;
                const int piece = k & 7;
                const int b = t*SB_SEQ_TILE + row;
                const bool ok = (b < nbu  &&  c0 + 16*piece < channels);
                const unsigned char *src = (ok)  ?  (code + (size_t) b*channels + c0 + 16*piece)  :  code;
                cp_async_16(dst0 + row*128 + piece*16, src, (ok)  ?  16  :  0);
            }
        }
        cp_async_commit();
    };
    issue(0);
    for (int t = 0;  t < ntiles;  t++)
    {
        issue(t + 1);
        cp_async_wait<1>();
        __syncthreads();
        const unsigned char *col = tile[t & 1] + threadIdx.x;
        const int rows = (nbu - t*SB_SEQ_TILE < SB_SEQ_TILE)  ?  (nbu - t*SB_SEQ_TILE)  :  SB_SEQ_TILE;
#pragma unroll 4
        for (int row = 0;  row < rows;  row++)
            f(t*SB_SEQ_TILE + row, (int) col[row*128]);
        __syncthreads();
    }
    cp_async_wait<0>();
}

// ==========================================================================================
// Bell MF and MFC/R2  (reference: src/bell_r2_mf.c)
__constant__ float c_bell_mf_fac[6];    // 700 ... 1700 Hz   (src/bell_r2_mf.c:251-254)
__constant__ float c_r2_fwd_fac[6];     // 1380 ... 1980 Hz  (src/bell_r2_mf.c:264-267)
__constant__ float c_r2_back_fac[6];    // 1140 ... 540 Hz   (src/bell_r2_mf.c:269-272)
__constant__ char c_bell_mf_positions[26];      // "1247C-358A--69*---0B----#"
__constant__ char c_r2_mf_positions[26];        // "1247B-358C--69D---0E----F"

struct MfParams
{
    int fwd;
};

struct MfLocal
{
};

// src/bell_r2_mf.c:554-628 / 793-863: returns index into the 25-character position table, or -1.
__device__ __forceinline__ int mf_pick(const float (&e)[6], float threshold, float twist, float relative_peak)
{
    int best;
    int second;
    float ebest;
    float esecond;

    if (e[0] > e[1])
    {
        best = 0;
        second = 1;
        ebest = e[0];
        esecond = e[1];
    }
    else
    {
        best = 1;
        second = 0;
        ebest = e[1];
        esecond = e[0];
    }
#pragma unroll
    for (int i = 2;  i < 6;  i++)
    {
        if (e[i] >= ebest)
        {
            second = best;
            esecond = ebest;
            best = i;
            ebest = e[i];
        }
        else if (e[i] >= esecond)
        {
            second = i;
            esecond = e[i];
        }
    }
    bool ok = (ebest >= threshold)  &&  (esecond >= threshold)
              &&  (ebest < fmul(esecond, twist))  &&  (fmul(ebest, twist) > esecond);
#pragma unroll
    for (int i = 0;  i < 6;  i++)
    {
        if (i != best  &&  i != second  &&  fmul(e[i], relative_peak) >= esecond)
            ok = false;
    }
    if (second < best)
    {
        const int t = best;
        best = second;
        second = t;
    }
    return (ok)  ?  (best*5 + second - 1)  :  -1;
}

struct BellMfDet
{
    static constexpr int NPAIRS = 3;
    static constexpr int BLOCK = 120;               // src/bell_r2_mf.c:204
    static constexpr bool ENERGY = false;
    static constexpr bool ENERGY_OUT = false;
    static constexpr bool FILTER = false;
    static constexpr bool RAW = false;
    typedef unsigned char code_t;
    typedef MfParams Params;
    typedef MfLocal Local;

    __device__ static __forceinline__ pair_t fac(const Params &, int p)
    {
        pair_t f;
        f.x = c_bell_mf_fac[2*p];
        f.y = c_bell_mf_fac[2*p + 1];
        return f;
    }
    __device__ static __forceinline__ void load_local(const Params &, int, Local &) {}
    __device__ static __forceinline__ bool filter_on(const Params &, int) { return false; }
    __device__ static __forceinline__ void load_filter(const Params &, int, int, float (&)[4]) {}
    __device__ static __forceinline__ void store_filter(const Params &, int, int, const float (&)[4]) {}

    __device__ static __forceinline__ int decide(const float (&e)[6], float, const Local &)
    {
        const int k = mf_pick(e, 3343803100.0f, 3.981f, 12.589f);      // src/bell_r2_mf.c:236-238
        return (k >= 0)  ?  (int) c_bell_mf_positions[k]  :  0;
    }
};

struct R2MfDet
{
    static constexpr int NPAIRS = 3;
    static constexpr int BLOCK = 133;               // src/bell_r2_mf.c:206
    static constexpr bool ENERGY = false;
    static constexpr bool ENERGY_OUT = false;
    static constexpr bool FILTER = false;
    static constexpr bool RAW = false;
    typedef unsigned char code_t;
    typedef MfParams Params;
    typedef MfLocal Local;

    __device__ static __forceinline__ pair_t fac(const Params &d, int p)
    {
        pair_t f;
        f.x = (d.fwd)  ?  c_r2_fwd_fac[2*p]  :  c_r2_back_fac[2*p];
        f.y = (d.fwd)  ?  c_r2_fwd_fac[2*p + 1]  :  c_r2_back_fac[2*p + 1];
        return f;
    }
    __device__ static __forceinline__ void load_local(const Params &, int, Local &) {}
    __device__ static __forceinline__ bool filter_on(const Params &, int) { return false; }
    __device__ static __forceinline__ void load_filter(const Params &, int, int, float (&)[4]) {}
    __device__ static __forceinline__ void store_filter(const Params &, int, int, const float (&)[4]) {}

    __device__ static __forceinline__ int decide(const float (&e)[6], float, const Local &)
    {
        const int k = mf_pick(e, 1031766650.0f, 5.012f, 12.589f);      // src/bell_r2_mf.c:240-242
        return (k >= 0)  ?  (int) c_r2_mf_positions[k]  :  0;
    }
};

struct MfSeqArgs
{
    SeqCommon q;
    const unsigned char *code;
    unsigned char *hits;                // Bell: [5][channels]; R2: [1][channels] = current_digit
};

// src/bell_r2_mf.c:629-661: a digit is reported when the last two blocks agree with this one and
// the two before differ (KP '*': the last three agree and the two before differ).
template <bool EMIT>
__global__ void __launch_bounds__(128) bell_mf_sequencer(const MfSeqArgs s)
{
    const int gc = blockIdx.x*blockDim.x + threadIdx.x;
    const int wg = gc >> 5;
    const bool live = (gc < s.q.channels);
    const int c = (live)  ?  gc  :  (s.q.channels - 1);
    const int B = BellMfDet::BLOCK;
    const int cs_old = (s.q.cs0 >= 0)  ?  s.q.cs0  :  s.q.cs[c];
    const int nb = (live)  ?  ((cs_old + s.q.n)/B)  :  0;
    const int nb_max = __reduce_max_sync(0xFFFFFFFFu, nb);
    const size_t C = s.q.channels;
    int h0 = s.hits[c];
    int h1 = s.hits[C + c];
    int h2 = s.hits[2*C + c];
    int h3 = s.hits[3*C + c];
    int h4 = s.hits[4*C + c];
    EventSink<EMIT> sink(s.q, wg);

    auto one_block = [&](int b, bool run, int hit)
    {
        bool ev = false;
        if (run)
        {
            ev = hit
                 &&  hit == h4  &&  hit == h3
                 &&  ((hit != '*'  &&  hit != h2  &&  hit != h1)
                      ||
                      (hit == '*'  &&  hit == h2  &&  hit != h1  &&  hit != h0));
            h0 = h1;
            h1 = h2;
            h2 = h3;
            h3 = h4;
            h4 = hit;
        }
        sink.push(ev, c, b, SPAN_B200_EV_DIGIT, hit, 0, 0);
    };

    if (s.q.cs0 >= 0  &&  (s.q.channels & 15) == 0)
    {
        __shared__ __align__(16) unsigned char tile[2][SB_SEQ_TILE*128];
        walk_codes8(s.code, s.q.channels, (s.q.cs0 + s.q.n)/B, tile, [&](int b, int hit) { one_block(b, live, hit); });
    }
    else
    {
        constexpr int G = 16;
        for (int b0 = 0;  b0 < nb_max;  b0 += G)
        {
            unsigned char codes[G];
#pragma unroll
            for (int i = 0;  i < G;  i++)
                codes[i] = (b0 + i < nb)  ?  s.code[(size_t) (b0 + i)*C + c]  :  (unsigned char) 0;
#pragma unroll
            for (int i = 0;  i < G;  i++)
            {
                if (b0 + i < nb_max)
                    one_block(b0 + i, b0 + i < nb, codes[i]);
            }
        }
    }
    sink.finish(wg);
    if (EMIT  &&  live)
    {
        s.hits[c] = (unsigned char) h0;
        s.hits[C + c] = (unsigned char) h1;
        s.hits[2*C + c] = (unsigned char) h2;
        s.hits[3*C + c] = (unsigned char) h3;
        s.hits[4*C + c] = (unsigned char) h4;
        s.q.cs[c] = cs_old + s.q.n - nb*B;
    }
}

// src/bell_r2_mf.c:864-876: report every change of the block decision.
template <bool EMIT>
__global__ void __launch_bounds__(128) r2_mf_sequencer(const MfSeqArgs s)
{
    const int gc = blockIdx.x*blockDim.x + threadIdx.x;
    const int wg = gc >> 5;
    const bool live = (gc < s.q.channels);
    const int c = (live)  ?  gc  :  (s.q.channels - 1);
    const int B = R2MfDet::BLOCK;
    const int cs_old = (s.q.cs0 >= 0)  ?  s.q.cs0  :  s.q.cs[c];
    const int nb = (live)  ?  ((cs_old + s.q.n)/B)  :  0;
    const int nb_max = __reduce_max_sync(0xFFFFFFFFu, nb);
    const size_t C = s.q.channels;
    int current = s.hits[c];
    EventSink<EMIT> sink(s.q, wg);

    auto one_block = [&](int b, bool run, int hit)
    {
        bool ev = false;
        if (run)
        {
            ev = (hit != current);
            current = hit;
        }
        sink.push(ev, c, b, SPAN_B200_EV_TONE, hit, (hit)  ?  -10  :  -99, 0);
    };

    if (s.q.cs0 >= 0  &&  (s.q.channels & 15) == 0)
    {
        __shared__ __align__(16) unsigned char tile[2][SB_SEQ_TILE*128];
        walk_codes8(s.code, s.q.channels, (s.q.cs0 + s.q.n)/B, tile, [&](int b, int hit) { one_block(b, live, hit); });
    }
    else
    {
        constexpr int G = 16;
        for (int b0 = 0;  b0 < nb_max;  b0 += G)
        {
            unsigned char codes[G];
#pragma unroll
            for (int i = 0;  i < G;  i++)
                codes[i] = (b0 + i < nb)  ?  s.code[(size_t) (b0 + i)*C + c]  :  (unsigned char) 0;
#pragma unroll
            for (int i = 0;  i < G;  i++)
            {
                if (b0 + i < nb_max)
                    one_block(b0 + i, b0 + i < nb, codes[i]);
            }
        }
    }
    sink.finish(wg);
    if (EMIT  &&  live)
    {
        s.hits[c] = (unsigned char) current;
        s.q.cs[c] = cs_old + s.q.n - nb*B;
    }
}

// ==========================================================================================
// Supervisory tones  (reference: src/super_tone_rx.c)
#define SB_ST_MAX_PAIRS     32          // 64 monitored bins: the reference's limit (src/spandsp/private/super_tone_rx.h:29,44)
#define SB_RAW_MAX_BINS     32          // raw Goertzel banks

struct StParams
{
    float fac[2*SB_ST_MAX_PAIRS];       // padded with zeros
    int bins;                           // monitored_frequencies
};

struct StLocal
{
    int bins;
};

// code layout: bits 0-6 = k1 + 1, bits 7-13 = k2 + 1, bit 15 = "one-bin quirk" (see below)
#define SB_ST_CODE(k1, k2)      ((unsigned short) ((((k2) + 1) << 7) | ((k1) + 1)))
#define SB_ST_QUIRK             0x8000

template <int NP>
struct SuperToneDet
{
    static constexpr int NPAIRS = NP;
    static constexpr int BLOCK = 128;               // src/spandsp/private/super_tone_rx.h:29
    static constexpr bool ENERGY = true;
    static constexpr bool ENERGY_OUT = false;
    static constexpr bool FILTER = false;
    static constexpr bool RAW = false;
    typedef unsigned short code_t;
    typedef StParams Params;
    typedef StLocal Local;

    __device__ static __forceinline__ pair_t fac(const Params &d, int p)
    {
        pair_t f;
        f.x = d.fac[2*p];
        f.y = d.fac[2*p + 1];
        return f;
    }
    __device__ static __forceinline__ void load_local(const Params &d, int, Local &l) { l.bins = d.bins; }
    __device__ static __forceinline__ bool filter_on(const Params &, int) { return false; }
    __device__ static __forceinline__ void load_filter(const Params &, int, int, float (&)[4]) {}
    __device__ static __forceinline__ void store_filter(const Params &, int, int, const float (&)[4]) {}

    // src/super_tone_rx.c:300-365
    __device__ static __forceinline__ int decide(const float (&res)[2*NP], float energy, const Local &l)
    {
        if (energy < 2104205.6f)                    // src/super_tone_rx.c:75,300
            return SB_ST_CODE(-1, -1);
        if (l.bins < 2)
        {
            // src/super_tone_rx.c:312-316: k1 = k2 = 0 and the bins are left un-read; the reference
            // then re-enters super_tone_chunk with zero energy.  The sequencer replays that.
            return SB_ST_CODE(0, 0) | SB_ST_QUIRK;
        }
        int k1;
        int k2;
        float r1;
        float r2;
        if (res[0] > res[1])
        {
            k1 = 0;
            k2 = 1;
            r1 = res[0];
            r2 = res[1];
        }
        else
        {
            k1 = 1;
            k2 = 0;
            r1 = res[1];
            r2 = res[0];
        }
#pragma unroll
        for (int j = 2;  j < 2*NP;  j++)
        {
            if (j < l.bins)
            {
                if (res[j] >= r1)
                {
                    k2 = k1;
                    r2 = r1;
                    k1 = j;
                    r1 = res[j];
                }
                else if (res[j] >= r2)
                {
                    k2 = j;
                    r2 = res[j];
                }
            }
        }
        if (fadd(r1, r2) < fmul(1.995f, energy))
        {
            k1 = -1;
            k2 = -1;
        }
        else if (r1 > fmul(3.981f, r2))
        {
            k2 = -1;
        }
        else if (k2 < k1)
        {
            const int t = k1;
            k1 = k2;
            k2 = t;
        }
        return SB_ST_CODE(k1, k2);
    }
};

// Raw Goertzel bank: arbitrary coefficients and block length, energies out
// (goertzel_update()/goertzel_result(), src/tone_detect.c:123-205).
// WITH_ENERGY: also the block's total energy (sequential float sum of x*x, as every caller of these primitives in the
// reference keeps it beside its Goertzels: src/ademco_contactid.c:901-905, src/v18.c:1560-1566) -> eout[block][channel]
template <int NP, bool WITH_ENERGY = false>
struct RawDet
{
    static constexpr int NPAIRS = NP;
    static constexpr int BLOCK = 0;                 // run-time block length
    static constexpr bool ENERGY = WITH_ENERGY;
    static constexpr bool ENERGY_OUT = false;
    static constexpr bool FILTER = false;
    static constexpr bool RAW = true;
    typedef unsigned char code_t;
    typedef StParams Params;
    typedef StLocal Local;

    __device__ static __forceinline__ pair_t fac(const Params &d, int p)
    {
        pair_t f;
        f.x = d.fac[2*p];
        f.y = d.fac[2*p + 1];
        return f;
    }
    __device__ static __forceinline__ void load_local(const Params &d, int, Local &l) { l.bins = d.bins; }
    __device__ static __forceinline__ bool filter_on(const Params &, int) { return false; }
    __device__ static __forceinline__ void load_filter(const Params &, int, int, float (&)[4]) {}
    __device__ static __forceinline__ void store_filter(const Params &, int, int, const float (&)[4]) {}
    __device__ static __forceinline__ int decide(const float (&)[2*NP], float, const Local &) { return 0; }
};

// Device copy of a super-tone descriptor's cadence templates (src/spandsp/private/super_tone_rx.h:40-49)
struct StTemplates
{
    int tones;
    const int *tone_segs;               // [tones]
    const int *tone_first;              // [tones] index of the tone's first element in `elements`
    const int4 *elements;               // {f1, f2, min_duration, max_duration} in samples
    int total_elements;
};

struct StSeqArgs
{
    SeqCommon q;
    const unsigned short *code;
    StTemplates t;
    int *segments;                      // [33][channels]: segments[i].{f1,f2,min_duration} at 3*i + {0,1,2}
    int *detected_tone;
    int *rotation;
    unsigned char *pending;             // one-bin quirk: a zero-energy re-chunk is owed at the next call
    int want_segments;
    void *log;                          // [groups][SB_ST_LOG] records kept by the count pass (wire or 24-byte, as the call asks), or NULL
};

#define SB_ST_TILE          48          // block rows of decisions staged per tile
#define SB_ST_CPW           8           // channels per warp
#define SB_ST_CPC           (4*SB_ST_CPW)   // channels per CTA (four warps)

// The segment history of one channel (src/spandsp/private/super_tone_rx.h:57: segments[11]); slot 10 holds the block
// pair seen last, slot 9 the segment in progress - those two live in registers in the sequencer - and slots 0..8 the
// finished segments.  Slots 0..8 are kept in shared memory, [field][channel], as a ring of 16: the cadence tests index
// them with run-time positions, which registers cannot do, and the reference's "shift everything down by one" becomes
// a step of the ring's head.
struct StSegs
{
    int *base;                          // this channel's column of int [48][SB_ST_CPC]
    int head;

    __device__ __forceinline__ int &f1(int i) const { return base[(3*((head + i) & 15))*SB_ST_CPC]; }
    __device__ __forceinline__ int &f2(int i) const { return base[(3*((head + i) & 15) + 1)*SB_ST_CPC]; }
    __device__ __forceinline__ int &dur(int i) const { return base[(3*((head + i) & 15) + 2)*SB_ST_CPC]; }
};

#define SB_ST_SMEM_ELEMENTS 160         // cadence template elements kept in shared memory (larger descriptors read them from global memory)
#define SB_ST_SMEM_TONES    64
#define SB_ST_LOG           (32*SB_ST_CPW)  // records per group of SB_ST_CPW channels the count pass keeps for the emit pass

// src/super_tone_rx.c:366-448.  The cadence logic is a chain of dependent, data-dependent branches per block: what it
// costs is latency, and a warp pays for every path any of its channels takes.  So
// - a warp carries only SB_ST_CPW channels (lanes 0..7; the bank then spreads over four times the warps, which is what
//   hides the latency: there are only tens of thousands of channels);
// - the per-channel history, the cadence templates and - where the bank has one block phase and the rows are 16-byte
//   aligned - tiles of the decision codes live in shared memory;
// - the cadence tests, which the reference runs every block, are run only at the blocks where their outcome can change
//   (see keep_until / scan_at below);
// - the count pass keeps the records it counts (up to SB_ST_LOG per group, in s.log) and, when they all fitted, commits
//   the channel state itself; the emit pass then only moves those records to their final places.  A group with more
//   records than that is walked again by the emit pass, from the untouched state, as in the other sequencers.
// Event order: (group of SB_ST_CPW channels, block, channel).
// SMALL: the descriptor's templates fit the shared-memory copies (the usual case; a compile-time fact so that they are
// read with shared-memory loads, not through generic pointers)
template <bool EMIT, bool SMALL>
__global__ void __launch_bounds__(128) super_tone_sequencer(const StSeqArgs s)
{
    __shared__ int seg[48*SB_ST_CPC];                       // (lanes beyond SB_ST_CPW never run a block: they share their warp's first column, untouched)
    __shared__ __align__(16) unsigned short tile[2][(SB_ST_TILE + 1)*SB_ST_CPC];
    __shared__ int4 s_elements[SB_ST_SMEM_ELEMENTS];
    __shared__ int s_tone_segs[SB_ST_SMEM_TONES];
    __shared__ int s_tone_first[SB_ST_SMEM_TONES];
    __shared__ int s_last_key[SB_ST_SMEM_TONES];                // the frequency pair of each tone's last step, as a decision code (-1: no steps)
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wg = blockIdx.x*4 + warp;                         // group of SB_ST_CPW channels
    const int gc = wg*SB_ST_CPW + lane;
    const bool group_there = (wg*SB_ST_CPW < s.q.channels);
    const size_t record = (s.q.wire)  ?  sizeof(span_b200_wire_event_t)  :  sizeof(span_b200_event_t);
    bool walk = true;
    if (EMIT  &&  s.log)
    {
        // Groups whose records the count pass kept: move them.  The CTA walks only if one of its groups overflowed
        // (its barriers need all four warps), and then only that group's lanes are live.
        const unsigned int count = (group_there)  ?  s.q.counts[wg]  :  0;
        walk = (count > SB_ST_LOG);
        if (!walk  &&  count)
        {
            const unsigned int base = s.q.offsets[wg];
            const unsigned int *src = (const unsigned int *) ((const char *) s.log + (size_t) wg*SB_ST_LOG*record);
            unsigned int *dst = (s.q.wire)  ?  (unsigned int *) (s.q.wire + base)  :  (unsigned int *) (s.q.events + base);
            const unsigned int words_per = (unsigned int) (record/4);
            const long long room = s.q.capacity - (long long) base;
            const unsigned int fit = (room <= 0)  ?  0  :  ((room < (long long) count)  ?  (unsigned int) room  :  count);
            for (unsigned int k = lane;  k < fit*words_per;  k += 32)
                dst[k] = src[k];
        }
        if (!__syncthreads_or(walk))
            return;
    }
    const bool live = (walk  &&  lane < SB_ST_CPW  &&  gc < s.q.channels);
    const int c = (lane < SB_ST_CPW  &&  gc < s.q.channels)  ?  gc  :  (s.q.channels - 1);
    const int col_in_cta = warp*SB_ST_CPW + ((lane < SB_ST_CPW)  ?  lane  :  0);
    const int B = 128;
    const size_t C = s.q.channels;
    const int cs_old = (s.q.cs0 >= 0)  ?  s.q.cs0  :  s.q.cs[c];
    const int nb = (live)  ?  ((cs_old + s.q.n)/B)  :  0;
    StSegs t;
    t.base = seg + col_in_cta;
    t.head = 0;
    // slots 9 (segment in progress) and 10 (pair seen last) are touched by every block: registers
    int f1_9 = 0;
    int f2_9 = 0;
    int dur_9 = 0;
    int key10 = 0;                      // slot 10, in the form of the decision codes
    if (lane < SB_ST_CPW)
    {
        for (int i = 0;  i < 27;  i++)
            seg[i*SB_ST_CPC + col_in_cta] = s.segments[(size_t) i*C + c];
        f1_9 = s.segments[(size_t) 27*C + c];
        f2_9 = s.segments[(size_t) 28*C + c];
        dur_9 = s.segments[(size_t) 29*C + c];
        key10 = SB_ST_CODE(s.segments[(size_t) 30*C + c], s.segments[(size_t) 31*C + c]);
    }
    // templates: shared memory copies where they fit
    constexpr bool small = SMALL;
    if (small)
    {
        for (int i = threadIdx.x;  i < s.t.total_elements;  i += 128)
            s_elements[i] = s.t.elements[i];
        for (int i = threadIdx.x;  i < s.t.tones;  i += 128)
        {
            const int steps = s.t.tone_segs[i];
            const int first = s.t.tone_first[i];
            s_tone_segs[i] = steps;
            s_tone_first[i] = first;
            s_last_key[i] = -1;
            if (steps > 0)
            {
                const int4 p = s.t.elements[first + steps - 1];
                // (a pair no decision code can carry never matches: any other non-negative value will do)
                s_last_key[i] = (p.x >= -1  &&  p.x < 127  &&  p.y >= -1  &&  p.y < 127)  ?  (int) SB_ST_CODE(p.x, p.y)  :  0x4000;
            }
        }
    }
    __syncthreads();
    const int4 *elements = s.t.elements;
    const int *tone_segs = s.t.tone_segs;
    const int *tone_first = s.t.tone_first;
    if constexpr (SMALL)
    {
        elements = s_elements;
        tone_segs = s_tone_segs;
        tone_first = s_tone_first;
    }
    const int ntones = s.t.tones;
    int detected = s.detected_tone[c];
    int rotation = s.rotation[c];
    int pending = s.pending[c];

    // Record output.  Emit pass: the usual warp-aggregated sink.  Count pass: the same aggregation, into the group's
    // piece of the log, so that the records already lie in their final order.
    EventSink<EMIT> sink(s.q, wg, SB_ST_CPW);
    SeqCommon lq = s.q;
    if (!EMIT)
    {
        lq.capacity = (s.log)  ?  SB_ST_LOG  :  0;
        if (s.q.wire)
            lq.wire = (span_b200_wire_event_t *) s.log + (size_t) wg*SB_ST_LOG;
        else
            lq.events = (span_b200_event_t *) s.log + (size_t) wg*SB_ST_LOG;
    }
    unsigned int logged = 0;
    const unsigned int lt = (1u << lane) - 1u;
    auto push = [&](bool has, int blk, int kind, int a, int b, int cc)
    {
        if (EMIT)
        {
            sink.push(has, c, blk, kind, a, b, cc);
        }
        else
        {
            const unsigned int m = __ballot_sync(0xFFFFFFFFu, has);
            if (has)
                put_event(lq, logged + __popc(m & lt), c, blk, kind, a, b, cc);
            logged += __popc(m);
        }
    };

    // While a block only lengthens the segment in progress, the outcome of the cadence tests is a function of that
    // one duration.  keep_until: the largest duration for which the detected tone's test still passes (-1: not
    // worked out yet); scan_at: the smallest duration at which the search over all tones can succeed at all.  Both
    // are recomputed whenever the history shifts, a tone is found or lost, and at the start of every call, so the
    // tests the reference runs every block are run here only at the blocks where their result can change.
    // repeat_until / differ_until fold the two into "the largest duration from which a repeat of the segment's pair /
    // a pair other than the one seen last is a bare increment"; key9 is slot 9 in the form of the decision codes.
    int keep_until = -1;
    int scan_at = 0;
    int repeat_until = -1;
    int differ_until = -1;
    int key9 = -1;
    // rotation modulo the detected tone's step count, stepped with it (-1: not worked out yet): the reference's
    // (rotation + steps - k) % steps without a division per test
    int rmod = -1;

    // One super_tone_chunk() step.  It can raise up to three callbacks, in this order: tone lost,
    // segment report, tone found.  They are recorded here and pushed by all lanes together.
    bool e_lost;
    bool e_seg;
    bool e_found;
    int seg_f1;
    int seg_f2;
    int seg_ms;
    int found_id;

    // st_test_cadence(pattern, steps > 0, rotation >= 0) (src/super_tone_rx.c:187-196) as a bound on the duration
    auto work_out_keep_until = [&]()
    {
        const int steps = tone_segs[detected];
        int j = 0;
        if (steps > 0)
        {
            if (rmod < 0)
                rmod = rotation%steps;
            const int y = rmod + steps - 1;
            j = (y >= steps)  ?  (y - steps)  :  y;
        }
        const int4 p = elements[tone_first[detected] + j];
        keep_until = (p.x != f1_9  ||  p.y != f2_9  ||  p.w < 0)  ?  0  :  (p.w >> 7);
    };

    auto chunk = [&](bool run, int k1, int k2)
    {
        e_lost = false;
        e_seg = false;
        e_found = false;
        seg_f1 = seg_f2 = seg_ms = found_id = 0;
        if (!run)
            return;
        if ((int) SB_ST_CODE(k1, k2) != key10)
        {
            key10 = SB_ST_CODE(k1, k2);
            dur_9++;
        }
        else if (k1 != f1_9  ||  k2 != f2_9)
        {
            if (detected >= 0)
            {
                // st_test_cadence(pattern, -steps, rotation++) (src/super_tone_rx.c:172-196): the finished segment
                // (slot 9, about to become slot 8... here still in registers) against the step it should have been,
                // with both duration bounds; the one before it (slot 8) likewise
                const int4 *pattern = elements + tone_first[detected];
                const int steps = tone_segs[detected];
                bool ok = true;
                int j = 0;
                if (steps > 0)
                {
                    if (rmod < 0)
                        rmod = rotation%steps;
                    const int x = rmod + steps - 2;
                    j = (x < 0)  ?  0  :  ((x >= steps)  ?  (x - steps)  :  x);
                    const int4 p = pattern[j];
                    const int d8 = t.dur(8)*128;
                    if (p.x != t.f1(8)  ||  p.y != t.f2(8)  ||  p.z > d8  ||  p.w < d8)
                        ok = false;
                    const int y = rmod + steps - 1;
                    j = (y >= steps)  ?  (y - steps)  :  y;
                    rmod = (rmod + 1 >= steps)  ?  0  :  (rmod + 1);
                }
                if (ok)
                {
                    const int4 p = pattern[j];
                    if (p.x != f1_9  ||  p.y != f2_9  ||  p.w < dur_9*128)
                        ok = false;
                }
                rotation++;
                if (!ok)
                {
                    detected = -1;
                    e_lost = true;
                }
            }
            if (s.want_segments)
            {
                e_seg = true;
                seg_f1 = f1_9;
                seg_f2 = f2_9;
                seg_ms = dur_9*128/8;
            }
            // segments[i] = segments[i + 1] for i = 0..8 (src/super_tone_rx.c:409-410): the ring steps, the finished
            // segment becomes slot 8
            t.head = (t.head + 1) & 15;
            t.f1(8) = f1_9;
            t.f2(8) = f2_9;
            t.dur(8) = dur_9;
            f1_9 = k1;
            f2_9 = k2;
            dur_9 = 1;
            keep_until = -1;
            scan_at = 0;
        }
        else
        {
            if (detected >= 0)
            {
                if (keep_until < 0)
                    work_out_keep_until();
                if (dur_9 > keep_until)
                {
                    detected = -1;
                    e_lost = true;
                    scan_at = 0;
                }
            }
            dur_9++;
        }
        if (detected < 0  &&  dur_9 >= scan_at)
        {
            // The search of src/super_tone_rx.c:425-437 (st_test_cadence with rotation < 0, src/super_tone_rx.c:
            // 198-212), tones in order, first match wins.  The newest segment is compared first (it is the one that
            // rules most tones out - by its frequency pair alone, kept per tone in s_last_key; the test is a
            // conjunction, so the order does not matter).  A tone whose only failing check is "segment in progress
            // still too short" names the duration at which to look again.
            const int mykey = SB_ST_CODE(f1_9, f2_9);
            const int d9 = dur_9*128;
            int next = 0x7FFFFFFF;
            for (int j = 0;  j < ntones;  j++)
            {
                if (small)
                {
                    const int key = s_last_key[j];
                    if (key != mykey  &&  key >= 0)
                        continue;
                }
                const int4 *pattern = elements + tone_first[j];
                const int steps = tone_segs[j];
                bool ok = true;
                bool short_yet = false;
                int need = 0;
                if (steps > 0)
                {
                    const int4 p = pattern[steps - 1];
                    if (p.x != f1_9  ||  p.y != f2_9  ||  p.w < d9)
                    {
                        ok = false;
                    }
                    else if (p.z > d9)
                    {
                        short_yet = true;
                        need = (p.z + 127) >> 7;
                    }
                }
                for (int i = steps - 2;  i >= 0  &&  ok;  i--)
                {
                    const int k = i + 10 - steps;
                    const int4 p = pattern[i];
                    const int d = t.dur(k)*128;
                    if (p.x != t.f1(k)  ||  p.y != t.f2(k)  ||  p.w < d  ||  p.z > d)
                        ok = false;
                }
                if (ok  &&  !short_yet)
                {
                    detected = j;
                    rotation = 0;
                    rmod = 0;
                    e_found = true;
                    found_id = j;
                    keep_until = -1;
                    break;
                }
                if (ok  &&  need < next)
                    next = need;
            }
            scan_at = next;
        }
        // (worked out here, not at the next block, so that the next block can already take the fast path)
        if (detected >= 0  &&  keep_until < 0)
            work_out_keep_until();
        key9 = SB_ST_CODE(f1_9, f2_9);
        const int no_scan_until = (scan_at >= 2)  ?  (scan_at - 2)  :  -1;
        repeat_until = (detected >= 0)  ?  keep_until  :  no_scan_until;
        differ_until = (detected >= 0)  ?  0x7FFFFFFF  :  no_scan_until;      // (that branch tests nothing while a tone is held)
    };

    auto flush = [&](int blk)
    {
        // (most blocks raise nothing in any channel of the warp: one vote instead of three)
        if (__any_sync(0xFFFFFFFFu, e_lost  ||  e_seg  ||  e_found))
        {
            push(e_lost, blk, SPAN_B200_EV_TONE, -1, -10, 0);
            push(e_seg, blk, SPAN_B200_EV_SEGMENT, seg_f1, seg_f2, seg_ms);
            push(e_found, blk, SPAN_B200_EV_TONE, found_id, -10, 0);
        }
    };

    const long long consumed_at_end = (long long) cs_old + s.q.n;
    // one decision code through the reference's block logic, including the one-bin quirk
    auto one_block = [&](int b, bool run, int code)
    {
        // Nearly always, in every channel of the warp: either the pair of the segment in progress again, or a pair
        // that differs from the one seen last (two tones too close for a 16 ms block to separate beat, and the
        // decision alternates: the reference then just notes the pair and lengthens the segment) - and no test due.
        const bool same = (code == key10);
        const bool fast = !run  ||  ((same)  ?  (key10 == key9  &&  dur_9 <= repeat_until)
                                             :  (!(code & SB_ST_QUIRK)  &&  dur_9 <= differ_until));
        if (__all_sync(0xFFFFFFFFu, fast))
        {
            if (run)
            {
                key10 = code;
                dur_9++;
            }
            return;
        }
        chunk(run, (code & 0x7F) - 1, ((code >> 7) & 0x7F) - 1);
        flush(b);
        const bool quirk = run  &&  (code & SB_ST_QUIRK);
        if (__any_sync(0xFFFFFFFFu, quirk))
        {
            // The reference loops straight back into super_tone_chunk with zero energy unless the
            // block ended exactly at the end of the caller's buffer (src/super_tone_rx.c:466-486).
            const bool again = quirk  &&  ((long long) (b + 1)*B < consumed_at_end);
            chunk(again, -1, -1);
            flush(b);
            if (quirk  &&  !again)
                pending = 1;
        }
    };

    // A zero-energy re-chunk owed from the previous call (one-bin descriptors only).
    {
        const bool owe = live  &&  pending  &&  s.q.n > 0;
        if (__any_sync(0xFFFFFFFFu, owe))
        {
            chunk(owe, -1, -1);
            flush(0);
        }
        if (owe)
            pending = 0;
    }
    if (s.q.cs0 >= 0  &&  (s.q.channels & 7) == 0)
    {
        const int nbu = (s.q.cs0 + s.q.n)/B;                // the same for every channel
        const int ntiles = (nbu + SB_ST_TILE - 1)/SB_ST_TILE;
        const int c0 = blockIdx.x*SB_ST_CPC;
        constexpr int PIECES = SB_ST_CPC/8;                 // 16-byte pieces per row
        auto issue = [&](int tl)
        {
            if (tl < ntiles)
            {
                const uint32_t dst0 = (uint32_t) __cvta_generic_to_shared(tile[tl & 1]);
                for (int k = threadIdx.x;  k < SB_ST_TILE*PIECES;  k += 128)
                {
                    const int row = k/PIECES;
                    const int piece = k - row*PIECES;
                    const int b = tl*SB_ST_TILE + row;
                    const bool ok = (b < nbu  &&  c0 + 8*piece < s.q.channels);
                    const unsigned short *src = (ok)  ?  (s.code + (size_t) b*C + c0 + 8*piece)  :  s.code;
                    cp_async_16(dst0 + row*(SB_ST_CPC*2) + piece*16, src, (ok)  ?  16  :  0);
                }
            }
            cp_async_commit();
        };
        issue(0);
        for (int tl = 0;  tl < ntiles;  tl++)
        {
            issue(tl + 1);
            cp_async_wait<1>();
            __syncthreads();
            const unsigned short *col = tile[tl & 1] + col_in_cta;
            const int rows = (nbu - tl*SB_ST_TILE < SB_ST_TILE)  ?  (nbu - tl*SB_ST_TILE)  :  SB_ST_TILE;
            // (the next row's code is fetched before this row's chain of branches, not after it; the tile has a spare
            // row so that the fetch behind the last one stays inside it)
            int code = (int) col[0];
            int blk = tl*SB_ST_TILE;
            const int blk_end = blk + rows;
#pragma unroll 1
            for (  ;  blk < blk_end;  blk++)
            {
                col += SB_ST_CPC;
                const int next = (int) col[0];
                one_block(blk, live, code);
                code = next;
            }
            __syncthreads();
        }
        cp_async_wait<0>();
    }
    else
    {
        const int nb_max = __reduce_max_sync(0xFFFFFFFFu, nb);
#pragma unroll 1
        for (int b = 0;  b < nb_max;  b++)
        {
            const bool run = (b < nb);
            const int code = (run)  ?  (int) s.code[(size_t) b*C + c]  :  0;
            one_block(b, run, code);
        }
    }
    bool commit = EMIT;
    if (!EMIT)
    {
        if (lane == 0  &&  group_there)
            s.q.counts[wg] = logged;
        commit = (s.log != NULL  &&  logged <= SB_ST_LOG);
    }
    if (commit  &&  live)
    {
        for (int i = 0;  i < 9;  i++)
        {
            s.segments[(size_t) (3*i)*C + c] = t.f1(i);
            s.segments[(size_t) (3*i + 1)*C + c] = t.f2(i);
            s.segments[(size_t) (3*i + 2)*C + c] = t.dur(i);
        }
        s.segments[(size_t) 27*C + c] = f1_9;
        s.segments[(size_t) 28*C + c] = f2_9;
        s.segments[(size_t) 29*C + c] = dur_9;
        s.segments[(size_t) 30*C + c] = (key10 & 0x7F) - 1;
        s.segments[(size_t) 31*C + c] = ((key10 >> 7) & 0x7F) - 1;
        s.detected_tone[c] = detected;
        s.rotation[c] = rotation;
        s.pending[c] = (unsigned char) pending;
        s.q.cs[c] = cs_old + s.q.n - nb*B;
    }
}

// ==========================================================================================
// Exclusive scan of per-channel event counts (single CTA; channels <= a few million).
__global__ void __launch_bounds__(1024) scan_counts(const unsigned int *counts, unsigned int *offsets, int n, unsigned long long *total)
{
    __shared__ unsigned long long part[1024];
    const int tid = threadIdx.x;
    const int per = (n + 1023)/1024;
    const int i0 = tid*per;
    const int i1 = (i0 + per < n)  ?  (i0 + per)  :  n;
    unsigned long long sum = 0;

    for (int i = i0;  i < i1;  i++)
        sum += counts[i];
    part[tid] = sum;
    __syncthreads();
    for (int d = 1;  d < 1024;  d <<= 1)
    {
        unsigned long long v = (tid >= d)  ?  part[tid - d]  :  0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    unsigned long long run = part[tid] - sum;
    for (int i = i0;  i < i1;  i++)
    {
        offsets[i] = (unsigned int) run;
        run += counts[i];
    }
    if (tid == 1023)
        total[0] = part[1023];
}

}  // namespace sb
