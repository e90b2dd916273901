// sb_mct_rx.cuh - the modem connect tone detector of src/modem_connect_tones.c:419-804: FAX CNG (1100 Hz), ANS / CED
// (2100 Hz) with its phase-reversal and 15 Hz AM variants (ANS/, ANSam, ANSam/), Bell ANS (2225 Hz), calling tone
// (1300 Hz), and the V.21 FAX preamble (a V.21 channel 2 fsk_rx followed by an HDLC flag counter).  Per sample: one
// biquad notch (plus a 15 Hz band-pass on the rectified signal for ANS), two integer level trackers and a small
// state machine.  Written __host__ __device__ so that tests/hostsim runs the same code on the CPU.
#pragma once

#include "sb_fsk_rx.cuh"

namespace sbf {

enum
{
    MCT_NONE = 0, MCT_FAX_CNG = 1, MCT_ANS = 2, MCT_ANS_PR = 3, MCT_ANSAM = 4, MCT_ANSAM_PR = 5, MCT_FAX_PREAMBLE = 6,
    MCT_FAX_CED_OR_PREAMBLE = 7, MCT_BELL_ANS = 8, MCT_CALLING_TONE = 9            // src/spandsp/modem_connect_tones.h:57-84
};

// How the level argument of a report is derived (the host finishes it with its own libm, like the reference would)
enum
{
    MCT_LEVEL_MINUS_99 = 0,         // tone lost: -99
    MCT_LEVEL_CHANNEL = 1,          // from channel_level (src/modem_connect_tones.c:563 etc.)
    MCT_LEVEL_FSK_POWER = 2         // from the V.21 receiver's power meter (src/modem_connect_tones.c:493)
};

// Per-channel state after the K_COUNT fields of the embedded FSK receiver, [field][channel]; floats as their bits
enum
{
    M_TONE_TYPE = K_COUNT, M_NOTCH_LEVEL, M_CHANNEL_LEVEL, M_AM_LEVEL, M_TONE_PRESENT, M_TONE_ON, M_CYCLE_DURATION,
    M_GOOD_CYCLES, M_RAW_BIT_STREAM, M_NUM_BITS, M_FLAGS_SEEN, M_FRAMING_OK, M_ZNOTCH_1, M_ZNOTCH_2, M_Z15HZ_1, M_Z15HZ_2,
    M_HIT, M_COUNT
};

#if !defined(SBF_FLOAT_OPS)
#define SBF_FLOAT_OPS
#if defined(__CUDA_ARCH__)
SB_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
SB_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
SB_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
// lfastrintf() is a plain (long int) cast on x86-64 gcc builds (src/spandsp/fast_convert.h:185-197): truncation
SB_HD int f_rint(float a) { return __float2int_rz(a); }
#else
// The host build of tests/hostsim is compiled without contraction; plain operators are the strict ones
SB_HD float f_mul(float a, float b) { volatile float r = a*b; return r; }
SB_HD float f_add(float a, float b) { volatile float r = a + b; return r; }
SB_HD float f_sub(float a, float b) { volatile float r = a - b; return r; }
SB_HD int f_rint(float a) { return (int) (long int) a; }
#endif
#endif

struct MctRx
{
    int tone_type, notch_level, channel_level, am_level, tone_present, tone_on, tone_cycle_duration, good_cycles;
    int raw_bit_stream, num_bits, flags_seen, framing_ok_announced;
    int hit;                    // what modem_connect_tones_rx_get() hands out: the last tone declared since the last get
    float znotch_1, znotch_2, z15hz_1, z15hz_2;
    FskRx fsk;
    int2 *ev;                   // reports: x = tone | (level kind << 16), y = the raw level source
    int ev_cap;
    int nev;

    SB_HD static int fbits(float f)
    {
#if defined(__CUDA_ARCH__)
        return __float_as_int(f);
#else
        int v;
        memcpy(&v, &f, 4);
        return v;
#endif
    }

    SB_HD static float bitsf(int v)
    {
#if defined(__CUDA_ARCH__)
        return __int_as_float(v);
#else
        float f;
        memcpy(&f, &v, 4);
        return f;
#endif
    }

    template <class V> SB_HD void visit_own(V &v, bool load)
    {
        v(M_TONE_TYPE, tone_type);
        v(M_NOTCH_LEVEL, notch_level);
        v(M_CHANNEL_LEVEL, channel_level);
        v(M_AM_LEVEL, am_level);
        v(M_TONE_PRESENT, tone_present);
        v(M_TONE_ON, tone_on);
        v(M_CYCLE_DURATION, tone_cycle_duration);
        v(M_GOOD_CYCLES, good_cycles);
        v(M_RAW_BIT_STREAM, raw_bit_stream);
        v(M_NUM_BITS, num_bits);
        v(M_FLAGS_SEEN, flags_seen);
        v(M_FRAMING_OK, framing_ok_announced);
        v(M_HIT, hit);
        int a = fbits(znotch_1);
        int b = fbits(znotch_2);
        int c = fbits(z15hz_1);
        int d = fbits(z15hz_2);
        v(M_ZNOTCH_1, a);
        v(M_ZNOTCH_2, b);
        v(M_Z15HZ_1, c);
        v(M_Z15HZ_2, d);
        if (load)
        {
            znotch_1 = bitsf(a);
            znotch_2 = bitsf(b);
            z15hz_1 = bitsf(c);
            z15hz_2 = bitsf(d);
        }
    }

    // report_tone_state() (src/modem_connect_tones.c:419-438).  Both of the reference's outlets are kept: the report
    // record (what a tone_callback would be given) and `hit` (what a state without a callback accumulates); which one
    // an application sees is the host's business.
    SB_HD void report(int tone, int kind, int raw)
    {
        if (tone != tone_present)
        {
            if (nev < ev_cap)
                ev[nev] = make_int2(tone | (kind << 16), raw);
            nev++;
            if (tone != MCT_NONE)
                hit = tone;
            tone_present = tone;
        }
    }

    // v21_put_bit() (src/modem_connect_tones.c:441-518): five back-to-back HDLC flags announce the FAX preamble
    SB_HD void v21_put_bit(int bit)
    {
        if (bit < 0)
        {
            if (bit == -1)                                  // SIG_STATUS_CARRIER_DOWN
            {
                if (tone_present == MCT_FAX_PREAMBLE)
                    report(MCT_NONE, MCT_LEVEL_MINUS_99, 0);
            }
            if (bit == -1  ||  bit == -2)                   // ... falls through to SIG_STATUS_CARRIER_UP
            {
                raw_bit_stream = 0;
                num_bits = 0;
                flags_seen = 0;
                framing_ok_announced = 0;
            }
            return;
        }
        raw_bit_stream = (int) (((unsigned int) raw_bit_stream << 1) | (((unsigned int) bit << 8) & 0x100u));
        num_bits++;
        if ((raw_bit_stream & 0x7F00) == 0x7E00)
        {
            if ((raw_bit_stream & 0x8000))
            {
                flags_seen = 0;                             // HDLC abort
            }
            else
            {
                if (flags_seen < 5)                         // HDLC_FRAMING_OK_THRESHOLD
                {
                    if (num_bits != 8)
                        flags_seen = 0;
                    if (++flags_seen >= 5  &&  !framing_ok_announced)
                    {
                        report(MCT_FAX_PREAMBLE, MCT_LEVEL_FSK_POWER, fsk.reading);
                        framing_ok_announced = 1;
                    }
                }
            }
            num_bits = 0;
        }
        else
        {
            if (flags_seen >= 5)
            {
                if (num_bits == 8)
                {
                    framing_ok_announced = 0;
                    flags_seen = 0;
                }
            }
        }
    }

    // One sample through the V.21 receiver; what it delivers goes to v21_put_bit() (src/modem_connect_tones.c:573-580)
    SB_HD void preamble_sample(int amp)
    {
        short tmp[4];
        fsk.out = tmp;
        fsk.out_cap = 4;
        fsk.nout = 0;
        fsk.sample(amp);
        for (int i = 0;  i < fsk.nout  &&  i < 4;  i++)
            v21_put_bit(tmp[i]);
    }

    // The single-notch detectors: FAX CNG 1100 Hz (:532-571), Bell ANS 2225 Hz (:696-735), calling tone 1300 Hz (:737-776)
    SB_HD void notch_sample(int amp, int tone, float b0, float a1, float a2, float b1)
    {
        float famp = (float) amp;
        const float v1 = f_sub(f_add(f_mul(b0, famp), f_mul(a1, znotch_1)), f_mul(a2, znotch_2));
        famp = f_add(f_add(v1, f_mul(b1, znotch_1)), znotch_2);
        znotch_2 = znotch_1;
        znotch_1 = v1;
        const int notched = (int) (short) f_rint(famp);
        channel_level += ((abs(amp) - channel_level) >> 5);
        notch_level += ((abs(notched) - notch_level) >> 5);
        if (channel_level > 70  &&  notch_level*6 < channel_level)
        {
            if (tone_present != tone)
            {
                if (++tone_cycle_duration >= 415*8)
                    report(tone, MCT_LEVEL_CHANNEL, channel_level);
            }
        }
        else
        {
            if (tone_present == tone)
                report(MCT_NONE, MCT_LEVEL_MINUS_99, 0);
            tone_cycle_duration = 0;
        }
    }

    // The 2100 Hz detector with phase-reversal timing and 15 Hz AM detection (src/modem_connect_tones.c:581-694)
    SB_HD void ans_sample(int amp)
    {
        float famp = (float) amp;
        float v1 = f_sub(f_add(fabsf(famp), f_mul(1.996667f, z15hz_1)), f_mul(0.9968004f, z15hz_2));
        const float filtered = f_mul(0.001599787f, f_sub(v1, z15hz_2));
        z15hz_2 = z15hz_1;
        z15hz_1 = v1;
        am_level += abs(f_rint(filtered)) - (am_level >> 8);
        v1 = f_sub(f_sub(f_mul(0.7552f, famp), f_mul(0.1183852f, znotch_1)), f_mul(0.5104039f, znotch_2));
        famp = f_add(f_add(v1, f_mul(0.1567596f, znotch_1)), znotch_2);
        znotch_2 = znotch_1;
        znotch_1 = v1;
        const int notched = (int) (short) f_rint(famp);
        channel_level += ((abs(amp) - channel_level) >> 5);
        notch_level += ((abs(notched) - notch_level) >> 4);
        if (channel_level <= 70)
        {
            if (tone_present != MCT_NONE)
                report(MCT_NONE, MCT_LEVEL_MINUS_99, 0);
            tone_cycle_duration = 0;
            good_cycles = 0;
            tone_on = 0;
            return;
        }
        tone_cycle_duration++;
        if (notch_level*6 < channel_level)
        {
            if (!tone_on)
            {
                if (tone_cycle_duration >= (450 - 25)*8)
                {
                    if (++good_cycles == 3)
                        report((am_level*15/256 > channel_level)  ?  MCT_ANSAM_PR  :  MCT_ANS_PR, MCT_LEVEL_CHANNEL, channel_level);
                }
                else
                {
                    good_cycles = 0;
                }
                tone_cycle_duration = 0;
            }
            else
            {
                if (tone_cycle_duration >= (450 + 100)*8)
                {
                    if (tone_present == MCT_NONE)
                        report((am_level*15/256 > channel_level)  ?  MCT_ANSAM  :  MCT_ANS, MCT_LEVEL_CHANNEL, channel_level);
                    good_cycles = 0;
                    tone_cycle_duration = (450 + 100)*8;
                }
            }
            tone_on = 1;
        }
        else if (notch_level*5 > channel_level)
        {
            if (tone_present == MCT_ANS)
            {
                report(MCT_NONE, MCT_LEVEL_MINUS_99, 0);
                good_cycles = 0;
            }
            else
            {
                if (tone_cycle_duration >= (450 + 25)*8)
                {
                    if (tone_present == MCT_ANS_PR  ||  tone_present == MCT_ANSAM_PR)
                        report(MCT_NONE, MCT_LEVEL_MINUS_99, 0);
                    good_cycles = 0;
                }
            }
            tone_on = 0;
        }
    }

    SB_HD bool uses_v21() const { return tone_type == MCT_FAX_PREAMBLE  ||  tone_type == MCT_FAX_CED_OR_PREAMBLE; }
    SB_HD bool uses_ans() const { return tone_type == MCT_FAX_CED_OR_PREAMBLE  ||  tone_type == MCT_ANS; }

    // The second pass of modem_connect_tones_rx() over one sample (the first, for the preamble types, is the V.21 pass)
    SB_HD void tone_sample(int amp)
    {
        switch (tone_type)
        {
        case MCT_FAX_CNG:
            notch_sample(amp, MCT_FAX_CNG, 0.792928f, 1.0018744927985f, 0.54196833412465f, -1.2994747954630f);
            break;
        case MCT_FAX_CED_OR_PREAMBLE:
        case MCT_ANS:
            ans_sample(amp);
            break;
        case MCT_BELL_ANS:
            notch_sample(amp, MCT_BELL_ANS, 0.739651f, -0.257384f, 0.510404f, 0.351437f);
            break;
        case MCT_CALLING_TONE:
            notch_sample(amp, MCT_CALLING_TONE, 0.755582f, 0.820887174515f, 0.541968324778f, -1.0456667108f);
            break;
        default:
            break;
        }
    }

    // modem_connect_tones_rx_init() (src/modem_connect_tones.c:823-875); the V.21 receiver is set up by the caller
    SB_HD void init(int type)
    {
        tone_type = type & 0xFFF;
        if (tone_type == MCT_ANS_PR  ||  tone_type == MCT_ANSAM  ||  tone_type == MCT_ANSAM_PR)
            tone_type = MCT_ANS;
        channel_level = 0;
        notch_level = 0;
        am_level = 0;
        tone_present = MCT_NONE;
        tone_cycle_duration = 0;
        good_cycles = 0;
        tone_on = 0;
        znotch_1 = znotch_2 = z15hz_1 = z15hz_2 = 0.0f;
        num_bits = 0;
        flags_seen = 0;
        framing_ok_announced = 0;
        raw_bit_stream = 0;
        hit = MCT_NONE;
    }
};

// The level argument of a report, finished with the host's libm as the reference computes it
static inline int host_mct_level(int kind, int raw)
{
    const float dbm0_max_power = 3.14f + 3.02f;             // DBM0_MAX_POWER, src/spandsp/telephony.h
    switch (kind)
    {
    case MCT_LEVEL_CHANNEL:
        // src/modem_connect_tones.c:563: level of a sine from its mean rectified amplitude
        return (int) (long int) (((raw == 0)  ?  (-96.329f + dbm0_max_power)  :  (20.0f*log10f((float) raw/32768.0f))) + dbm0_max_power + 0.8f);
    case MCT_LEVEL_FSK_POWER:
        // lfastrintf(fsk_rx_signal_power()) = power_meter_current_dbm0() (src/power_meter.c:115-122)
        if (raw <= 0)
            return (int) (long int) (-96.329f + dbm0_max_power);
        return (int) (long int) (10.0f*log10f((float) raw/(32767.0f*32767.0f) + 1.0e-10f) + dbm0_max_power);
    default:
        return -99;
    }
}

struct MctArgs
{
    FskArgs f;                      // samples, channel count, state [M_COUNT][channels], window, sine
    int2 *ev;                       // [channel][ev_cap]
    long long ev_cap;
    int *nev;                       // [channels]
};

#if defined(__CUDACC__)

// One channel's samples in order, 16 bytes per load where the row is aligned
template <class F> __device__ __forceinline__ void mct_for_each_sample(const int16_t *row, int n, F f)
{
    int pos = 0;
    if ((((size_t) row) & 15) == 0)
    {
#pragma unroll 1
        for (  ;  pos + 8 <= n;  pos += 8)
        {
            const uint4 v = __ldg((const uint4 *) (row + pos));
            const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0;  k < 4;  k++)
            {
                f((int) (short) (w[k] & 0xFFFFu));
                f((int) (short) (w[k] >> 16));
            }
        }
    }
#pragma unroll 1
    for (  ;  pos < n;  pos++)
        f((int) __ldg(row + pos));
}

// modem_connect_tones_rx() for 32 channels per CTA, one thread per channel.  As in the reference, a call first runs
// the V.21 receiver over all its samples (preamble types) and then the tone detector over the same samples: the two
// share tone_present, so the order is part of the result.
__global__ void __launch_bounds__(32) mct_rx_kernel(const MctArgs a)
{
    extern __shared__ int fsk_smem[];
    const int lane = threadIdx.x;
    const int c = blockIdx.x*32 + lane;
    const bool active = (c < a.f.channels);
    MctRx r;
    int live;
    const int cc = (active)  ?  c  :  (a.f.channels - 1);
    fsk_bind(r.fsk, a.f, fsk_smem, lane, cc, live);
    if (!active)
        return;
    FskLoader ld = {a.f.state, (size_t) a.f.channels, (size_t) c};
    r.visit_own(ld, true);
    r.ev = a.ev + (size_t) c*a.ev_cap;
    r.ev_cap = (int) a.ev_cap;
    r.nev = 0;
    const int16_t *row = a.f.amp + (long long) c*a.f.stride;
    if (r.uses_v21())
        mct_for_each_sample(row, a.f.n, [&r](int amp) { r.preamble_sample(amp); });
    mct_for_each_sample(row, a.f.n, [&r](int amp) { r.tone_sample(amp); });
    r.fsk.out = NULL;
    FskStorer st = {a.f.state, (size_t) a.f.channels, (size_t) c};
    r.visit_own(st, false);
    fsk_unbind(r.fsk, a.f, c, live);
    a.nev[c] = r.nev;
}

// modem_connect_tones_rx_init() for channels [first, first + count): all state and the V.21 window zeroed, the V.21
// receiver initialised as fsk_rx_init(V.21 ch 2, synchronous) + fsk_rx_set_signal_cutoff(-45.5) where it is used
__global__ void mct_init_kernel(const MctArgs a, int first, int count, int tone_type, FskSetup su)
{
    const int idx = blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= count)
        return;
    const int c = first + idx;
    for (int f = 0;  f < M_COUNT;  f++)
        a.f.state[(size_t) f*a.f.channels + c] = 0;
    for (int k = 0;  k < 2*SBF_MAX_WINDOW;  k++)
        a.f.window[(size_t) k*a.f.channels + c] = make_int2(0, 0);
    MctRx r;
    FskLoader ld = {a.f.state, (size_t) a.f.channels, (size_t) c};
    r.fsk.visit(ld);
    r.visit_own(ld, true);
    r.init(tone_type);
    if (r.uses_v21())
        r.fsk.restart(su.baud_rate, su.framing_mode, su.rate0, su.rate1, su.on_power, su.off_power);
    FskStorer st = {a.f.state, (size_t) a.f.channels, (size_t) c};
    r.fsk.visit(st);
    r.visit_own(st, false);
}

#endif  // __CUDACC__

}  // namespace sbf
