// sb_sig_rx.cuh - the in-band signalling tone receiver of src/sig_tone.c:402-664 (2280 Hz AC15-style, 2600 Hz, and the
// 2400 + 2600 Hz pair of SS5): per tone a notch made of two cascaded biquads that doubles as the guard filter, leaky
// power meters on the notched and the flat signal, a "sharp" detector (power ratio through the notch, with on / off
// persistence timers) that hands over to a "flat" one (band-pass + threshold) while a tone lasts, and the notch
// insertion logic that removes the tone from the audio - the receiver rewrites its input buffer.
// Float build arithmetic in the reference's expression order.  Written __host__ __device__ so that tests/hostsim runs
// the same code on the CPU.
#pragma once

#include <stdint.h>
#include <limits.h>
#include <math.h>
#include <string.h>

#include <cuda_runtime.h>

#if !defined(SB_HD)
#define SB_HD __host__ __device__ __forceinline__
#endif

namespace sbs {

enum
{
    SIG_TONE_1_PRESENT = 0x001, SIG_TONE_1_CHANGE = 0x002, SIG_TONE_2_PRESENT = 0x004, SIG_TONE_2_CHANGE = 0x008,
    SIG_TONE_RX_PASSTHROUGH = 0x040, SIG_TONE_RX_FILTER_TONE = 0x080           // src/spandsp/sig_tone.h:67-88
};

// Per-channel state, one int per field, stored [field][channel]; floats as their bits
enum
{
    T_TYPE = 0, T_RX_TONE, T_NOTCH_FILTER,
    T_Z1 = 3,                   // notch_z1[tone][2], 6 fields
    T_Z2 = 9,                   // notch_z2[tone][2], 6 fields
    T_POWER = 15,               // tone[].power.reading, 3 fields
    T_FLAT_Z = 18,              // 2 fields
    T_FLAT_POWER = 20, T_PERSISTENCE, T_LAST_PRESENT, T_FLAT_THR, T_SHARP_THR, T_RATIO, T_FLAT_MODE, T_FLAT_TIMEOUT,
    T_NOTCH_TIMEOUT, T_STATE, T_DURATION, T_COUNT
};

#if defined(__CUDA_ARCH__)
SB_HD float s_fmul(float a, float b) { return __fmul_rn(a, b); }
SB_HD float s_fadd(float a, float b) { return __fadd_rn(a, b); }
// float -> int16_t as gcc does it on x86-64 (cvttss2si, low 16 bits kept): the argument of power_meter_update()
SB_HD int s_trunc16(float a) { return (int) (short) __float2int_rz(a); }
SB_HD int s_lrintf(float a) { return __float2int_rn(a); }
SB_HD int s_fbits(float f) { return __float_as_int(f); }
SB_HD float s_bitsf(int v) { return __int_as_float(v); }
#else
SB_HD float s_fmul(float a, float b) { volatile float r = a*b; return r; }
SB_HD float s_fadd(float a, float b) { volatile float r = a + b; return r; }
SB_HD int s_trunc16(float a) { return (int) (short) (int) a; }
SB_HD int s_lrintf(float a) { return (int) lrintf(a); }
SB_HD int s_fbits(float f) { int v; memcpy(&v, &f, 4); return v; }
SB_HD float s_bitsf(int v) { float f; memcpy(&f, &v, 4); return f; }
#endif

// sig_tone_notch_coeffs_t: a1, b1, a2, b2 (src/sig_tone.c:79-126); index 0 = 2280 Hz, 1 = 2400 Hz, 2 = 2600 Hz
struct NotchCoeffs
{
    float a1[3];
    float b1[3];
    float a2[3];
    float b2[3];
};

SB_HD NotchCoeffs notch_coeffs(int set)
{
    NotchCoeffs k;
    if (set == 0)
    {
        k.a1[0] = 0.878906f;  k.a1[1] = 0.439362f;  k.a1[2] = 1.0f;
        k.b1[0] = 0.0f;  k.b1[1] = -0.287627f;  k.b1[2] = -0.883605f;
        k.a2[0] = 0.0f;  k.a2[1] = 0.433228f;  k.a2[2] = 1.0f;
        k.b2[0] = 0.0f;  k.b2[1] = -0.530792f;  k.b2[2] = -0.883605f;
    }
    else if (set == 1)
    {
        k.a1[0] = 0.862000f;  k.a1[1] = 0.612055f;  k.a1[2] = 1.0f;
        k.b1[0] = 0.0f;  k.b1[1] = -0.456264f;  k.b1[2] = -0.864899f;
        k.a2[0] = 0.0f;  k.a2[1] = 0.621021f;  k.a2[2] = 1.0f;
        k.b2[0] = 0.0f;  k.b2[1] = -0.690738f;  k.b2[2] = -0.864899f;
    }
    else
    {
        k.a1[0] = 0.862000f;  k.a1[1] = 0.902374f;  k.a1[2] = 1.0f;
        k.b1[0] = 0.0f;  k.b1[1] = -0.732727f;  k.b1[2] = -0.864899f;
        k.a2[0] = 0.0f;  k.a2[1] = 0.910766f;  k.a2[2] = 1.0f;
        k.b2[0] = 0.0f;  k.b2[1] = -0.952393f;  k.b2[2] = -0.864899f;
    }
    return k;
}

// sig_tone_descriptor_t, the receive-side fields (src/sig_tone.c:143-219); tone_type 1 = 2280 Hz, 2 = 2600 Hz,
// 3 = 2400 Hz + 2600 Hz
struct Desc
{
    int tones;
    int notch_set[2];
    int has_flat;
    int sharp_flat_timeout;
    int notch_lag_time;
    int tone_on_check_time;
    int tone_off_check_time;
};

SB_HD Desc descriptor(int tone_type)
{
    Desc d;
    d.notch_lag_time = 225*8;
    d.tone_on_check_time = 3*8;
    d.tone_off_check_time = 8*8;
    if (tone_type == 1)
    {
        d.tones = 1;
        d.notch_set[0] = 0;
        d.notch_set[1] = -1;
        d.has_flat = 1;
        d.sharp_flat_timeout = 225*8;
    }
    else if (tone_type == 2)
    {
        d.tones = 1;
        d.notch_set[0] = 2;
        d.notch_set[1] = -1;
        d.has_flat = 0;
        d.sharp_flat_timeout = 0;
    }
    else
    {
        d.tones = 2;
        d.notch_set[0] = 1;
        d.notch_set[1] = 2;
        d.has_flat = 0;
        d.sharp_flat_timeout = 0;
    }
    return d;
}

// Host side of sig_tone_rx_init() (src/sig_tone.c:714-716): {flat threshold, sharp threshold, detection ratio}
static inline void host_sig_thresholds(int tone_type, int32_t out[3])
{
    const float ratio_db = (tone_type == 1)  ?  13.0f  :  15.6f;
    // power_meter_level_dbm0(-30.0f) (src/power_meter.c:82-93)
    float level = -30.0f - (3.14f + 3.02f);
    if (level > 0.0)
        level = 0.0;
    const int32_t thr = (int32_t) (powf(10.0f, level/10.0f)*(32767.0f*32767.0f));
    out[0] = thr;
    out[1] = thr;
    out[2] = (int32_t) (powf(10.0f, ratio_db/10.0f) + 1.0f);
}

struct SigLoader
{
    const int *state;
    size_t channels;
    size_t c;
    SB_HD int operator()(int field) const { return state[(size_t) field*channels + c]; }
};

struct SigStorer
{
    int *state;
    size_t channels;
    size_t c;
    SB_HD void operator()(int field, int v) const { state[(size_t) field*channels + c] = v; }
};

struct SigRx
{
    int tone_type, current_rx_tone, current_notch_filter;
    float z1[3][2];
    float z2[3][2];
    int power[3];
    float flat_z[2];
    int flat_power, persistence, last_present, flat_thr, sharp_thr, ratio, flat_mode, flat_timeout, notch_timeout;
    int state, duration;
    Desc d;
    NotchCoeffs nk[2];
    int2 *ev;                   // sig_update() calls: x = signalling_state, y = signalling_state_duration
    int ev_cap;
    int nev;

    SB_HD void load(const SigLoader &ld)
    {
        tone_type = ld(T_TYPE);
        current_rx_tone = ld(T_RX_TONE);
        current_notch_filter = ld(T_NOTCH_FILTER);
        for (int j = 0;  j < 3;  j++)
        {
            for (int i = 0;  i < 2;  i++)
            {
                z1[j][i] = s_bitsf(ld(T_Z1 + 2*j + i));
                z2[j][i] = s_bitsf(ld(T_Z2 + 2*j + i));
            }
            power[j] = ld(T_POWER + j);
        }
        flat_z[0] = s_bitsf(ld(T_FLAT_Z));
        flat_z[1] = s_bitsf(ld(T_FLAT_Z + 1));
        flat_power = ld(T_FLAT_POWER);
        persistence = ld(T_PERSISTENCE);
        last_present = ld(T_LAST_PRESENT);
        flat_thr = ld(T_FLAT_THR);
        sharp_thr = ld(T_SHARP_THR);
        ratio = ld(T_RATIO);
        flat_mode = ld(T_FLAT_MODE);
        flat_timeout = ld(T_FLAT_TIMEOUT);
        notch_timeout = ld(T_NOTCH_TIMEOUT);
        state = ld(T_STATE);
        duration = ld(T_DURATION);
        bind();
    }

    SB_HD void store(const SigStorer &st) const
    {
        st(T_TYPE, tone_type);
        st(T_RX_TONE, current_rx_tone);
        st(T_NOTCH_FILTER, current_notch_filter);
        for (int j = 0;  j < 3;  j++)
        {
            for (int i = 0;  i < 2;  i++)
            {
                st(T_Z1 + 2*j + i, s_fbits(z1[j][i]));
                st(T_Z2 + 2*j + i, s_fbits(z2[j][i]));
            }
            st(T_POWER + j, power[j]);
        }
        st(T_FLAT_Z, s_fbits(flat_z[0]));
        st(T_FLAT_Z + 1, s_fbits(flat_z[1]));
        st(T_FLAT_POWER, flat_power);
        st(T_PERSISTENCE, persistence);
        st(T_LAST_PRESENT, last_present);
        st(T_FLAT_THR, flat_thr);
        st(T_SHARP_THR, sharp_thr);
        st(T_RATIO, ratio);
        st(T_FLAT_MODE, flat_mode);
        st(T_FLAT_TIMEOUT, flat_timeout);
        st(T_NOTCH_TIMEOUT, notch_timeout);
        st(T_STATE, state);
        st(T_DURATION, duration);
    }

    SB_HD void bind()
    {
        d = descriptor(tone_type);
        nk[0] = notch_coeffs(d.notch_set[0]);
        nk[1] = notch_coeffs((d.notch_set[1] >= 0)  ?  d.notch_set[1]  :  d.notch_set[0]);
    }

    // sig_tone_rx_init() (src/sig_tone.c:672-719): everything zero except these
    SB_HD void init(int type, int flat_threshold, int sharp_threshold, int detection_ratio)
    {
        tone_type = type;
        current_rx_tone = 0;
        current_notch_filter = 0;
        for (int j = 0;  j < 3;  j++)
        {
            z1[j][0] = z1[j][1] = z2[j][0] = z2[j][1] = 0.0f;
            power[j] = 0;
        }
        flat_z[0] = flat_z[1] = 0.0f;
        flat_power = 0;
        persistence = 0;
        last_present = -1;
        flat_thr = flat_threshold;
        sharp_thr = sharp_threshold;
        ratio = detection_ratio;
        flat_mode = 0;
        flat_timeout = 0;
        notch_timeout = 0;
        state = 0;
        duration = 0;
        bind();
    }

    // power_meter_update() with shift 5 (src/power_meter.c:65-69, sig_tone.c:709-712)
    SB_HD static int meter(int &reading, int amp16)
    {
        reading += ((amp16*amp16 - reading) >> 5);
        return reading;
    }

    // One sample of sig_tone_rx()'s loop; returns what the reference leaves in amp[i]
    SB_HD int sample(int amp)
    {
        const int l = (d.tones == 2)  ?  3  :  1;
        float notched[3] = {0.0f, 0.0f, 0.0f};
        int notch_power[3];
        notch_power[0] = 0;
        notch_power[1] = INT_MAX;
        notch_power[2] = INT_MAX;
        if (duration < INT_MAX)
            duration++;
        float signal = (float) amp;
        for (int j = 0;  j < l;  j++)
        {
            const NotchCoeffs &k = nk[(j == 1)  ?  1  :  0];            // coeff_sets[] = {0, 1, 0} (src/sig_tone.c:235-240)
            float v = s_fadd(s_fadd(s_fmul(signal, k.a1[0]), s_fmul(z1[j][0], k.b1[1])), s_fmul(z1[j][1], k.b1[2]));
            float x = v;
            v = s_fadd(v, s_fadd(s_fmul(z1[j][0], k.a1[1]), s_fmul(z1[j][1], k.a1[2])));
            z1[j][1] = z1[j][0];
            z1[j][0] = x;
            v = s_fadd(v, s_fadd(s_fmul(z2[j][0], k.b2[1]), s_fmul(z2[j][1], k.b2[2])));
            x = v;
            v = s_fadd(v, s_fadd(s_fmul(z2[j][0], k.a2[1]), s_fmul(z2[j][1], k.a2[2])));
            z2[j][1] = z2[j][0];
            z2[j][0] = x;
            notched[j] = v;
            notch_power[j] = meter(power[j], s_trunc16(v));
            if (j == 1)
                signal = v;
        }
        if ((state & (SIG_TONE_1_PRESENT | SIG_TONE_2_PRESENT)))
        {
            if (flat_timeout  &&  --flat_timeout == 0)
                flat_mode = 1;
        }
        else
        {
            flat_timeout = d.sharp_flat_timeout;
            flat_mode = 0;
        }
        int immediate = -1;
        if (flat_mode)
        {
            float bandpass = (float) amp;
            if (d.has_flat)
            {
                // flat_coeffs[0] (src/sig_tone.c:128-141)
                float v = s_fadd(s_fadd(s_fmul((float) amp, 0.393676f), s_fmul(flat_z[0], -0.261778f)), s_fmul(flat_z[1], -0.359985f));
                const float x = v;
                v = s_fadd(v, s_fadd(s_fmul(flat_z[0], -0.5f), s_fmul(flat_z[1], -0.5f)));
                flat_z[1] = flat_z[0];
                flat_z[0] = x;
                bandpass = v;
            }
            const int fp = meter(flat_power, s_trunc16(bandpass));
            if ((state & (SIG_TONE_1_PRESENT | SIG_TONE_2_PRESENT)))
            {
                if (fp < flat_thr)
                {
                    state &= ~SIG_TONE_1_PRESENT;
                    state |= SIG_TONE_1_CHANGE;
                }
            }
            else
            {
                if (fp > flat_thr)
                    state |= (SIG_TONE_1_PRESENT | SIG_TONE_1_CHANGE);
            }
            if ((state & (SIG_TONE_1_PRESENT | SIG_TONE_2_PRESENT)))
            {
                notch_timeout = d.notch_lag_time;
            }
            else
            {
                if (notch_timeout)
                    notch_timeout--;
            }
        }
        else
        {
            const int fp = meter(flat_power, amp);
            if (fp >= sharp_thr)
            {
                const int m = (notch_power[0] < notch_power[1])  ?  0  :  1;
                if ((notch_power[m] >> 6)*ratio < (fp >> 6))
                    immediate = m;
                else if ((notch_power[2] >> 6)*ratio < (fp >> 7))
                    immediate = 2;
            }
            if ((state & (SIG_TONE_1_PRESENT | SIG_TONE_2_PRESENT)))
            {
                if (immediate != current_notch_filter)
                {
                    if (--persistence == 0)
                    {
                        persistence = d.tone_on_check_time;
                        state |= ((state & (SIG_TONE_1_PRESENT | SIG_TONE_2_PRESENT)) << 1);
                        state &= ~(SIG_TONE_1_PRESENT | SIG_TONE_2_PRESENT);
                    }
                }
                else
                {
                    persistence = d.tone_off_check_time;
                }
            }
            else
            {
                if (notch_timeout)
                    notch_timeout--;
                if (immediate >= 0  &&  immediate == last_present)
                {
                    if (--persistence == 0)
                    {
                        persistence = d.tone_off_check_time;
                        notch_timeout = d.notch_lag_time;
                        const int bits = (immediate == 0)  ?  (SIG_TONE_1_PRESENT | SIG_TONE_1_CHANGE)
                                       : (immediate == 1)  ?  (SIG_TONE_2_PRESENT | SIG_TONE_2_CHANGE)
                                       : (SIG_TONE_1_PRESENT | SIG_TONE_2_PRESENT | SIG_TONE_1_CHANGE | SIG_TONE_2_CHANGE);
                        state |= bits;
                        current_notch_filter = immediate;
                    }
                }
                else
                {
                    persistence = d.tone_on_check_time;
                }
            }
        }
        if ((state & (SIG_TONE_1_CHANGE | SIG_TONE_2_CHANGE)))
        {
            if (nev < ev_cap)
                ev[nev] = make_int2(state, duration);
            nev++;
            state &= ~(SIG_TONE_1_CHANGE | SIG_TONE_2_CHANGE);
            duration = 0;
        }
        int out = amp;
        if ((current_rx_tone & SIG_TONE_RX_PASSTHROUGH))
        {
            if ((current_rx_tone & SIG_TONE_RX_FILTER_TONE)  ||  notch_timeout)
            {
                // fsaturatef() (src/spandsp/saturated.h:142-149)
                const float f = notched[current_notch_filter];
                out = (f > 32767.0f)  ?  32767  :  (f < -32768.0f)  ?  -32768  :  (int) (short) s_lrintf(f);
            }
        }
        else
        {
            out = 0;
        }
        last_present = immediate;
        return out;
    }
};

struct SigArgs
{
    int16_t *amp;                   // [channel][sample], rewritten in place
    long long stride;
    int n;
    int channels;
    int *state;                     // [T_COUNT][channels]
    int2 *ev;                       // [channel][ev_cap]
    long long ev_cap;
    int *nev;                       // [channels]
};

#if defined(__CUDACC__)

// sig_tone_rx() for every channel: thread per channel, samples read and written back 16 bytes at a time
__global__ void __launch_bounds__(64) sig_rx_kernel(const SigArgs a)
{
    const int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= a.channels)
        return;
    SigRx r;
    SigLoader ld = {a.state, (size_t) a.channels, (size_t) c};
    r.load(ld);
    r.ev = a.ev + (size_t) c*a.ev_cap;
    r.ev_cap = (int) a.ev_cap;
    r.nev = 0;
    int16_t *row = a.amp + (long long) c*a.stride;
    int pos = 0;
    if ((((size_t) row) & 15) == 0)
    {
#pragma unroll 1
        for (  ;  pos + 8 <= a.n;  pos += 8)
        {
            const uint4 v = *((const uint4 *) (row + pos));
            unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0;  k < 4;  k++)
            {
                const int lo = r.sample((int) (short) (w[k] & 0xFFFFu));
                const int hi = r.sample((int) (short) (w[k] >> 16));
                w[k] = ((unsigned int) lo & 0xFFFFu) | ((unsigned int) hi << 16);
            }
            *((uint4 *) (row + pos)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
#pragma unroll 1
    for (  ;  pos < a.n;  pos++)
        row[pos] = (int16_t) r.sample((int) row[pos]);
    SigStorer st = {a.state, (size_t) a.channels, (size_t) c};
    r.store(st);
    a.nev[c] = r.nev;
}

// mode 0: sig_tone_rx_init(tone_type = ia; thresholds ib, ic, id); 1: sig_tone_rx_set_mode(mode = ia)
__global__ void sig_ctl_kernel(const SigArgs a, int first, int count, int mode, int ia, int ib, int ic, int id)
{
    const int idx = blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= count)
        return;
    const int c = first + idx;
    SigRx r;
    SigLoader ld = {a.state, (size_t) a.channels, (size_t) c};
    SigStorer st = {a.state, (size_t) a.channels, (size_t) c};
    if (mode == 0)
    {
        r.init(ia, ib, ic, id);
    }
    else
    {
        r.load(ld);
        r.current_rx_tone = ia;
    }
    r.store(st);
}

#endif  // __CUDACC__

}  // namespace sbs
