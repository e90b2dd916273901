// sb_gen.cuh - the signal sources of the reference's own tests and of BASELINE's workloads, on the device:
//   * tone_gen()   cadenced multi-tone generator       (src/tone_generate.c:125-230, float DDS src/dds_float.c:2102-2198)
//   * dtmf_tx()    digit queue -> tone_gen per digit    (src/dtmf.c:521-676)
//   * awgn()       ran1-style uniform source + polar Gaussian (src/awgn.c:82-196)
// with these the multi-channel inputs (cfg2 / cfg5: dtmf_tx + awgn per channel) are produced where
// they are consumed, instead of on the host and over PCIe.  Written __host__ __device__ so that tests/hostsim runs
// the same code on the CPU.  The tone path is exact by construction (integer phase accumulators, the 2048-entry
// float sine table, one float multiply per tone, float adds in order, truncating conversion).  The noise path is
// exact up to the one transcendental in it: log() of the device library against the host libm's (both below 1 ulp;
// a differing last bit moves an output sample only if it straddles a rounding boundary, ~1e-14 per sample).
#pragma once

#include <stdint.h>
#include <math.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime.h>

#if !defined(SB_HD)
#define SB_HD __host__ __device__ __forceinline__
#endif

namespace sbg {

#define SBG_SINE_WORDS      2048        // src/dds_float.c:48-51 (SLENK = 11)
#define SBG_QUEUE           128         // MAX_DTMF_DIGITS, src/spandsp/dtmf.h:74
#define SBG_RAN_TABLE       97          // src/spandsp/private/awgn.h

#if defined(__CUDA_ARCH__)
SB_HD float g_fmul(float a, float b) { return __fmul_rn(a, b); }
SB_HD float g_fadd(float a, float b) { return __fadd_rn(a, b); }
SB_HD double g_dmul(double a, double b) { return __dmul_rn(a, b); }
SB_HD double g_dadd(double a, double b) { return __dadd_rn(a, b); }
SB_HD double g_ddiv(double a, double b) { return __ddiv_rn(a, b); }
SB_HD double g_dsqrt(double a) { return __dsqrt_rn(a); }
SB_HD int g_lrint(double a) { return __double2int_rn(a); }
#else
SB_HD float g_fmul(float a, float b) { volatile float r = a*b; return r; }
SB_HD float g_fadd(float a, float b) { volatile float r = a + b; return r; }
SB_HD double g_dmul(double a, double b) { volatile double r = a*b; return r; }
SB_HD double g_dadd(double a, double b) { volatile double r = a + b; return r; }
SB_HD double g_ddiv(double a, double b) { volatile double r = a/b; return r; }
SB_HD double g_dsqrt(double a) { return sqrt(a); }
SB_HD int g_lrint(double a) { return (int) lrint(a); }
#endif

SB_HD int g_fbits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int v;
    memcpy(&v, &f, 4);
    return v;
#endif
}

SB_HD float g_bitsf(int v)
{
#if defined(__CUDA_ARCH__)
    return __int_as_float(v);
#else
    float f;
    memcpy(&f, &v, 4);
    return f;
#endif
}

// dds_phase_ratef() (src/dds_float.c:2108-2111) and dds_scaling_dbm0f() (:2120-2123), host side
static inline int32_t host_dds_phase_ratef(float frequency)
{
    return (int32_t) (frequency*65536.0f*65536.0f/8000);
}

static inline float host_dds_scaling_dbm0f(float level)
{
    return powf(10.0f, (level - 3.14f)/20.0f)*32767.0f;     // DBM0_MAX_SINE_POWER = 3.14
}

// src/dds_float.c:51-2101: sin(2*pi*i/2048) as 8-decimal literals which the compiler parses as float
static inline void host_make_sine_table(float *t)
{
    char buf[64];
    for (int i = 0;  i < SBG_SINE_WORDS;  i++)
    {
        snprintf(buf, sizeof(buf), "%.8f", sin(2.0*3.14159265358979323846*(double) i/2048.0));
        t[i] = strtof(buf, NULL);
    }
}

// Sequential int16 writer for one channel row: collects eight samples and stores them as 16 bytes where the row is
// 16-byte aligned (the device rows are), scalar stores otherwise and for the tail.
struct RowOut
{
    int16_t *row;
    int pos;
    unsigned int w[4];
    bool vec;

    SB_HD void begin(int16_t *r)
    {
        row = r;
        pos = 0;
        vec = ((((size_t) r) & 15) == 0);
        w[0] = w[1] = w[2] = w[3] = 0;
    }

    SB_HD void put(int v)
    {
        if (!vec)
        {
            row[pos++] = (int16_t) v;
            return;
        }
        const int k = pos & 7;
        const unsigned int h = (unsigned int) v & 0xFFFFu;
        if ((k & 1) == 0)
            w[k >> 1] = h;
        else
            w[k >> 1] |= h << 16;
        pos++;
        if (k == 7)
        {
#if defined(__CUDA_ARCH__)
            *((uint4 *) (row + pos - 8)) = make_uint4(w[0], w[1], w[2], w[3]);
#else
            memcpy(row + pos - 8, w, 16);
#endif
        }
    }

    // The samples of an unfinished group of eight
    SB_HD void flush()
    {
        if (!vec)
            return;
        const int k = pos & 7;
        for (int i = 0;  i < k;  i++)
            row[pos - k + i] = (int16_t) ((w[i >> 1] >> ((i & 1)*16)) & 0xFFFFu);
    }
};

// tone_gen_state_t (src/spandsp/private/tone_generate.h) and tone_gen() (src/tone_generate.c:125-230), float build
struct ToneGen
{
    int rate[4];
    float gain[4];
    unsigned int phase[4];
    int duration[4];
    int repeat;
    int section;
    int position;
    const float *sine;

    // dds_modf() (src/dds_float.c:2169-2176)
    SB_HD float dds_modf(int i, int r, float g)
    {
        const float amp = g_fmul(sine[phase[i] >> 21], g);
        phase[i] += (unsigned int) r;
        return amp;
    }

    // lfastrintf() is a truncating cast on x86-64 gcc builds (src/spandsp/fast_convert.h:185-197)
    SB_HD static int to_amp(float x)
    {
#if defined(__CUDA_ARCH__)
        return (int) (short) __float2int_rz(x);
#else
        return (int) (short) (long int) x;
#endif
    }

    template <class OUT> SB_HD int run(OUT &out, int max_samples)
    {
        if (section < 0)
            return 0;
        int samples = 0;
        while (samples < max_samples)
        {
            int limit = samples + duration[section] - position;
            if (limit > max_samples)
                limit = max_samples;
            position += (limit - samples);
            if ((section & 1))
            {
                for (  ;  samples < limit;  samples++)
                    out.put(0);
            }
            else if (rate[0] < 0)
            {
                // modulated tone: exactly two tones
                for (  ;  samples < limit;  samples++)
                {
                    const float a = dds_modf(0, -rate[0], gain[0]);
                    const float b = dds_modf(1, rate[1], gain[1]);
                    out.put(to_amp(g_fmul(a, g_fadd(1.0f, b))));
                }
            }
            else
            {
                for (  ;  samples < limit;  samples++)
                {
                    float xamp = 0.0f;
                    for (int i = 0;  i < 4;  i++)
                    {
                        if (rate[i] == 0)
                            break;
                        xamp = g_fadd(xamp, dds_modf(i, rate[i], gain[i]));
                    }
                    out.put(to_amp(xamp));
                }
            }
            if (position >= duration[section])
            {
                position = 0;
                if (++section > 3  ||  duration[section] == 0)
                {
                    if (!repeat)
                    {
                        section = -1;
                        break;
                    }
                    section = 0;
                }
            }
        }
        return samples;
    }
};

// Per-channel state of a DTMF transmitter, one int per field, stored [field][channel]; floats as their bits
enum
{
    D_SECTION = 0, D_POSITION, D_PHASE0, D_PHASE1, D_RATE0, D_RATE1, D_GAIN0, D_GAIN1, D_DUR0, D_DUR1,
    D_LOW, D_HIGH, D_ON, D_OFF, D_QRD, D_QCNT, D_COUNT
};

struct GenLoader
{
    const int *state;
    size_t channels;
    size_t c;
    SB_HD int operator()(int field) const { return state[(size_t) field*channels + c]; }
};

struct GenStorer
{
    int *state;
    size_t channels;
    size_t c;
    SB_HD void operator()(int field, int v) const { state[(size_t) field*channels + c] = v; }
};

// dtmf_tx_state_t (src/spandsp/private/dtmf.h:31-52) and dtmf_tx() (src/dtmf.c:550-597)
struct DtmfTx
{
    ToneGen tones;
    float low_level;
    float high_level;
    int on_time;
    int off_time;
    int qrd;                        // read position in the channel's digit ring
    int qcnt;                       // digits waiting
    const unsigned char *queue;     // ring element k at queue[k*qstride]
    size_t qstride;
    const int *row_rate;            // dds_phase_ratef() of the four row and four column frequencies
    const int *col_rate;

    SB_HD void load(const GenLoader &ld)
    {
        tones.section = ld(D_SECTION);
        tones.position = ld(D_POSITION);
        tones.phase[0] = (unsigned int) ld(D_PHASE0);
        tones.phase[1] = (unsigned int) ld(D_PHASE1);
        tones.phase[2] = tones.phase[3] = 0;
        tones.rate[0] = ld(D_RATE0);
        tones.rate[1] = ld(D_RATE1);
        tones.rate[2] = tones.rate[3] = 0;
        tones.gain[0] = g_bitsf(ld(D_GAIN0));
        tones.gain[1] = g_bitsf(ld(D_GAIN1));
        tones.gain[2] = tones.gain[3] = 0.0f;
        tones.duration[0] = ld(D_DUR0);
        tones.duration[1] = ld(D_DUR1);
        tones.duration[2] = tones.duration[3] = 0;
        tones.repeat = 0;
        low_level = g_bitsf(ld(D_LOW));
        high_level = g_bitsf(ld(D_HIGH));
        on_time = ld(D_ON);
        off_time = ld(D_OFF);
        qrd = ld(D_QRD);
        qcnt = ld(D_QCNT);
    }

    SB_HD void store(const GenStorer &st) const
    {
        st(D_SECTION, tones.section);
        st(D_POSITION, tones.position);
        st(D_PHASE0, (int) tones.phase[0]);
        st(D_PHASE1, (int) tones.phase[1]);
        st(D_RATE0, tones.rate[0]);
        st(D_RATE1, tones.rate[1]);
        st(D_GAIN0, g_fbits(tones.gain[0]));
        st(D_GAIN1, g_fbits(tones.gain[1]));
        st(D_DUR0, tones.duration[0]);
        st(D_DUR1, tones.duration[1]);
        st(D_LOW, g_fbits(low_level));
        st(D_HIGH, g_fbits(high_level));
        st(D_ON, on_time);
        st(D_OFF, off_time);
        st(D_QRD, qrd);
        st(D_QCNT, qcnt);
    }

    // dtmf_tx_init() (src/dtmf.c:637-660): default level and timing, nothing queued, generator idle
    SB_HD void init(float default_gain)
    {
        for (int i = 0;  i < 4;  i++)
        {
            tones.rate[i] = 0;
            tones.gain[i] = 0.0f;
            tones.phase[i] = 0;
            tones.duration[i] = 0;
        }
        tones.rate[0] = row_rate[0];                // tone_gen_init(&s->tones, &dtmf_digit_tones[0])
        tones.rate[1] = col_rate[0];
        tones.gain[0] = default_gain;
        tones.gain[1] = default_gain;
        tones.duration[0] = 50*8;
        tones.duration[1] = 55*8;
        tones.repeat = 0;
        tones.section = -1;
        tones.position = 0;
        low_level = default_gain;
        high_level = default_gain;
        on_time = 50*8;
        off_time = 55*8;
        qrd = 0;
        qcnt = 0;
    }

    // "123A456B789C*0#D" (src/dtmf.c:123); returns -1 for a character that is not a DTMF digit
    SB_HD static int digit_index(int ch)
    {
        switch (ch)
        {
        case '1': return 0;
        case '2': return 1;
        case '3': return 2;
        case 'A': return 3;
        case '4': return 4;
        case '5': return 5;
        case '6': return 6;
        case 'B': return 7;
        case '7': return 8;
        case '8': return 9;
        case '9': return 10;
        case 'C': return 11;
        case '*': return 12;
        case '0': return 13;
        case '#': return 14;
        case 'D': return 15;
        }
        return -1;
    }

    template <class OUT> SB_HD int tx(OUT &out, int max_samples)
    {
        int len = 0;
        if (tones.section >= 0)
            len = tones.run(out, max_samples);
        while (len < max_samples)
        {
            if (qcnt <= 0)
                break;                                      // no callback to ask for more (src/dtmf.c:567-569)
            const int digit = queue[(size_t) qrd*qstride];
            qrd = (qrd + 1) & (SBG_QUEUE - 1);
            qcnt--;
            if (digit == 0)
                continue;
            const int idx = digit_index(digit);
            if (idx < 0)
                continue;
            // tone_gen_init() from dtmf_digit_tones[idx], then this transmitter's levels and timing
            tones.rate[0] = row_rate[idx >> 2];
            tones.rate[1] = col_rate[idx & 3];
            tones.phase[0] = 0;
            tones.phase[1] = 0;
            tones.gain[0] = low_level;
            tones.gain[1] = high_level;
            tones.duration[0] = on_time;
            tones.duration[1] = off_time;
            tones.section = 0;
            tones.position = 0;
            len += tones.run(out, max_samples - len);
        }
        return len;
    }
};

// awgn_state_t (src/spandsp/private/awgn.h) and awgn() (src/awgn.c:82-196)
struct Awgn
{
    int ix1;
    int ix2;
    int ix3;
    int odd;
    double amp2;
    double rms;
    double *r;                      // table element j at r[j*rs]
    size_t rs;

    // ran_init() (src/awgn.c:82-103)
    SB_HD void ran_init(int idum)
    {
        const double rm1 = 1.0/259200.0;
        const double rm2 = 1.0/134456.0;
        if (idum < 0)
            idum = -idum;
        ix1 = (54773 + idum)%259200;
        ix1 = (7141*ix1 + 54773)%259200;
        ix2 = ix1%134456;
        ix1 = (7141*ix1 + 54773)%259200;
        ix3 = ix1%243000;
        for (int j = 0;  j < SBG_RAN_TABLE;  j++)
        {
            ix1 = (7141*ix1 + 54773)%259200;
            ix2 = (8121*ix2 + 28411)%134456;
            r[(size_t) j*rs] = g_dmul(g_dadd((double) ix1, g_dmul((double) ix2, rm2)), rm1);
        }
    }

    // ran() (src/awgn.c:106-129)
    SB_HD double ran()
    {
        const double rm1 = 1.0/259200.0;
        const double rm2 = 1.0/134456.0;
        ix1 = (7141*ix1 + 54773)%259200;
        ix2 = (8121*ix2 + 28411)%134456;
        ix3 = (4561*ix3 + 51349)%243000;
        const int j = (97*ix3)/243000;
        if (j > 96  ||  j < 0)
            return -1.0;
        const double temp = r[(size_t) j*rs];
        r[(size_t) j*rs] = g_dmul(g_dadd((double) ix1, g_dmul((double) ix2, rm2)), rm1);
        return temp;
    }

    // awgn() (src/awgn.c:169-196) with fsaturate() (src/spandsp/saturated.h:152-159)
    SB_HD int sample()
    {
        double amp;
        odd = !odd;
        if (odd)
        {
            amp = amp2;
        }
        else
        {
            double v1;
            double v2;
            double rr;
            do
            {
                v1 = g_dadd(g_dmul(2.0, ran()), -1.0);
                v2 = g_dadd(g_dmul(2.0, ran()), -1.0);
                rr = g_dadd(g_dmul(v1, v1), g_dmul(v2, v2));
            }
            while (rr >= 1.0);
            rr = g_dsqrt(g_ddiv(g_dmul(-2.0, log(rr)), rr));
            amp2 = g_dmul(v1, rr);
            amp = g_dmul(v2, rr);
        }
        amp = g_dmul(amp, rms);
        if (amp > 32767.0)
            return 32767;
        if (amp < -32768.0)
            return -32768;
        return (int) (short) g_lrint(amp);
    }
};

// sat_add16() (src/spandsp/saturated.h)
SB_HD int sat_add16(int a, int b)
{
    const int s = a + b;
    return (s > 32767)  ?  32767  :  (s < -32768)  ?  -32768  :  s;
}

struct DtmfTxArgs
{
    int16_t *amp;                   // [channel][sample], row stride in samples
    long long stride;
    int max_samples;
    int channels;
    int zero_fill;                  // write zeros after the last generated sample (the reference leaves them alone)
    int *state;                     // [D_COUNT][channels]
    unsigned char *queue;           // [SBG_QUEUE][channels]
    const float *sine;              // SBG_SINE_WORDS
    int *lens;                      // [channels]: what dtmf_tx() returned
    int rates[8];                   // rows, columns
};

struct AwgnArgs
{
    int16_t *amp;
    long long stride;
    int n;
    int channels;
    int add;                        // 1: amp[i] = sat_add16(amp[i], awgn());  0: amp[i] = awgn()
    int *istate;                    // [4][channels]: ix1, ix2, ix3, odd
    double *dstate;                 // [2 + SBG_RAN_TABLE][channels]: amp2, rms, r[]
};

// ------------------------------------------------------------------------------------------
// Cadenced tone generator banks: tone_gen_descriptor_init() + tone_gen_init() + tone_gen() per channel
// (src/tone_generate.c:60-230).  Per-channel state, one int per field, [field][channel]; floats as their bits.
enum
{
    T_SECTION = 0, T_POSITION, T_REPEAT,
    T_PHASE0, T_RATE0 = T_PHASE0 + 4, T_GAIN0 = T_RATE0 + 4, T_DUR0 = T_GAIN0 + 4, T_COUNT = T_DUR0 + 4
};

SB_HD void tone_gen_load(ToneGen &t, const GenLoader &ld)
{
    t.section = ld(T_SECTION);
    t.position = ld(T_POSITION);
    t.repeat = ld(T_REPEAT);
    for (int i = 0;  i < 4;  i++)
    {
        t.phase[i] = (unsigned int) ld(T_PHASE0 + i);
        t.rate[i] = ld(T_RATE0 + i);
        t.gain[i] = g_bitsf(ld(T_GAIN0 + i));
        t.duration[i] = ld(T_DUR0 + i);
    }
}

SB_HD void tone_gen_store(const ToneGen &t, const GenStorer &st)
{
    st(T_SECTION, t.section);
    st(T_POSITION, t.position);
    st(T_REPEAT, t.repeat);
    for (int i = 0;  i < 4;  i++)
    {
        st(T_PHASE0 + i, (int) t.phase[i]);
        st(T_RATE0 + i, t.rate[i]);
        st(T_GAIN0 + i, g_fbits(t.gain[i]));
        st(T_DUR0 + i, t.duration[i]);
    }
}

// What tone_gen_descriptor_init() computes on the host (libm powf), per channel
struct ToneDesc
{
    int rate[2];
    float gain[2];
    int duration[4];
    int repeat;
};

static inline void host_tone_descriptor(ToneDesc &d, int f1, int l1, int f2, int l2, int d1, int d2, int d3, int d4, int repeat)
{
    memset(&d, 0, sizeof(d));
    if (f1)
    {
        d.rate[0] = host_dds_phase_ratef((float) f1);
        if (f2 < 0)
            d.rate[0] = -d.rate[0];
        d.gain[0] = host_dds_scaling_dbm0f((float) l1);
    }
    if (f2)
    {
        d.rate[1] = host_dds_phase_ratef((float) abs(f2));
        d.gain[1] = (f2 < 0)  ?  ((float) l2/100.0f)  :  host_dds_scaling_dbm0f((float) l2);
    }
    d.duration[0] = d1*8000/1000;
    d.duration[1] = d2*8000/1000;
    d.duration[2] = d3*8000/1000;
    d.duration[3] = d4*8000/1000;
    d.repeat = repeat;
}

// tone_gen_init() (src/tone_generate.c:232-270): copy the descriptor, zero the phases, start at section 0
SB_HD void tone_gen_start(ToneGen &t, const ToneDesc &d)
{
    for (int i = 0;  i < 4;  i++)
    {
        t.rate[i] = (i < 2)  ?  d.rate[i]  :  0;
        t.gain[i] = (i < 2)  ?  d.gain[i]  :  0.0f;
        t.phase[i] = 0;
        t.duration[i] = d.duration[i];
    }
    t.repeat = d.repeat;
    t.section = 0;
    t.position = 0;
}

struct ToneGenArgs
{
    int16_t *amp;
    long long stride;
    int max_samples;
    int channels;
    int zero_fill;
    int *state;                     // [T_COUNT][channels]
    const float *sine;
    int *lens;
};

// ------------------------------------------------------------------------------------------
// v29_tx() (src/v29tx.c:100-322), float build.  The data bits come from a per-channel source on the device: a
// maximal-length sequence (x^23 + x^18 + 1, the generator the oracle harness feeds the reference with) or a buffer
// of caller bits; when the buffer runs out the transmitter sees SIG_STATUS_END_OF_DATA and shuts down as the
// reference does (32 bauds of scrambled ones, then silence).
#define SBG_V29_TX_STEPS        9       // V29_TX_FILTER_STEPS, src/spandsp/private/v29tx.h:30
#define SBG_V29_TX_SETS         10      // TX_PULSESHAPER_COEFF_SETS

enum
{
    X_BIT_RATE = 0, X_RRC_STEP, X_SCRAMBLE, X_TRAIN_SCRAMBLE, X_IN_TRAINING, X_TRAINING_STEP, X_TRAINING_OFFSET,
    X_CARRIER_PHASE, X_BAUD_PHASE, X_CONSTELLATION, X_REAL_BITS, X_SRC_MODE, X_LFSR, X_BIT_POS, X_BIT_COUNT,
    X_STATUS, X_GAIN, X_BASE_GAIN, X_RRC_RE, X_RRC_IM = X_RRC_RE + SBG_V29_TX_STEPS, X_COUNT = X_RRC_IM + SBG_V29_TX_STEPS
};

struct V29Tx
{
    int bit_rate;
    int rrc_step;
    unsigned int scramble_reg;
    unsigned int training_scramble_reg;
    int in_training;
    int training_step;
    int training_offset;
    unsigned int carrier_phase;
    int carrier_phase_rate;
    int baud_phase;
    int constellation_state;
    int real_bits;                  // current_get_bit == get_bit (else fake_get_bit)
    int src_mode;                   // 0: PRBS, 1: caller bits
    unsigned int lfsr;
    int bit_pos;
    int bit_count;
    int status;                     // bit 0: SIG_STATUS_END_OF_DATA reported, bit 1: SIG_STATUS_SHUTDOWN_COMPLETE reported
    float gain;
    float base_gain;
    float rrc_re[SBG_V29_TX_STEPS];
    float rrc_im[SBG_V29_TX_STEPS];
    const unsigned char *bits;      // caller bits of this channel, LSB first
    const float *sine;              // [2048]
    const float *shaper;            // [10][9]

    SB_HD void load(const GenLoader &ld)
    {
        bit_rate = ld(X_BIT_RATE);
        rrc_step = ld(X_RRC_STEP);
        scramble_reg = (unsigned int) ld(X_SCRAMBLE);
        training_scramble_reg = (unsigned int) ld(X_TRAIN_SCRAMBLE);
        in_training = ld(X_IN_TRAINING);
        training_step = ld(X_TRAINING_STEP);
        training_offset = ld(X_TRAINING_OFFSET);
        carrier_phase = (unsigned int) ld(X_CARRIER_PHASE);
        baud_phase = ld(X_BAUD_PHASE);
        constellation_state = ld(X_CONSTELLATION);
        real_bits = ld(X_REAL_BITS);
        src_mode = ld(X_SRC_MODE);
        lfsr = (unsigned int) ld(X_LFSR);
        bit_pos = ld(X_BIT_POS);
        bit_count = ld(X_BIT_COUNT);
        status = ld(X_STATUS);
        gain = g_bitsf(ld(X_GAIN));
        base_gain = g_bitsf(ld(X_BASE_GAIN));
        for (int i = 0;  i < SBG_V29_TX_STEPS;  i++)
        {
            rrc_re[i] = g_bitsf(ld(X_RRC_RE + i));
            rrc_im[i] = g_bitsf(ld(X_RRC_IM + i));
        }
    }

    SB_HD void store(const GenStorer &st) const
    {
        st(X_BIT_RATE, bit_rate);
        st(X_RRC_STEP, rrc_step);
        st(X_SCRAMBLE, (int) scramble_reg);
        st(X_TRAIN_SCRAMBLE, (int) training_scramble_reg);
        st(X_IN_TRAINING, in_training);
        st(X_TRAINING_STEP, training_step);
        st(X_TRAINING_OFFSET, training_offset);
        st(X_CARRIER_PHASE, (int) carrier_phase);
        st(X_BAUD_PHASE, baud_phase);
        st(X_CONSTELLATION, constellation_state);
        st(X_REAL_BITS, real_bits);
        st(X_SRC_MODE, src_mode);
        st(X_LFSR, (int) lfsr);
        st(X_BIT_POS, bit_pos);
        st(X_BIT_COUNT, bit_count);
        st(X_STATUS, status);
        st(X_GAIN, g_fbits(gain));
        st(X_BASE_GAIN, g_fbits(base_gain));
        for (int i = 0;  i < SBG_V29_TX_STEPS;  i++)
        {
            st(X_RRC_RE + i, g_fbits(rrc_re[i]));
            st(X_RRC_IM + i, g_fbits(rrc_im[i]));
        }
    }

    // set_working_gain() (src/v29tx.c:299-319)
    SB_HD void set_working_gain()
    {
        if (bit_rate == 9600)
            gain = g_fmul(0.387f, base_gain);
        else if (bit_rate == 7200)
            gain = g_fmul(0.605f, base_gain);
        else if (bit_rate == 4800)
            gain = g_fmul(0.470f, base_gain);
    }

    // v29_tx_restart() (src/v29tx.c:362-398)
    SB_HD int restart(int rate, int tep)
    {
        bit_rate = rate;
        set_working_gain();
        if (rate == 9600)
            training_offset = 0;
        else if (rate == 7200)
            training_offset = 2;
        else if (rate == 4800)
            training_offset = 4;
        else
            return -1;
        for (int i = 0;  i < SBG_V29_TX_STEPS;  i++)
            rrc_re[i] = rrc_im[i] = 0.0f;
        rrc_step = 0;
        scramble_reg = 0;
        training_scramble_reg = 0x2A;
        in_training = 1;
        training_step = (tep)  ?  0  :  480;        // V29_TRAINING_SEG_TEP / V29_TRAINING_SEG_1
        carrier_phase = 0;
        baud_phase = 0;
        constellation_state = 0;
        real_bits = 0;
        return 0;
    }

    // The caller's get_bit: 0/1, or -1 for SIG_STATUS_END_OF_DATA
    SB_HD int source_bit()
    {
        if (src_mode == 0)
        {
            const int bit = (int) ((lfsr >> 22) ^ (lfsr >> 17)) & 1;
            lfsr = ((lfsr << 1) | (unsigned int) bit) & 0x7FFFFFu;
            return bit;
        }
        if (bit_pos >= bit_count)
            return -1;
        const int bit = (bits[bit_pos >> 3] >> (bit_pos & 7)) & 1;
        bit_pos++;
        return bit;
    }

    // get_scrambled_bit() (src/v29tx.c:104-125)
    SB_HD int get_scrambled_bit()
    {
        int bit = 1;                                    // fake_get_bit()
        if (real_bits)
        {
            bit = source_bit();
            if (bit < 0)
            {
                status |= 1;
                real_bits = 0;
                in_training = 1;
                bit = 1;
            }
        }
        const int out_bit = (bit ^ (int) (scramble_reg >> (18 - 1)) ^ (int) (scramble_reg >> (23 - 1))) & 1;
        scramble_reg = (scramble_reg << 1) | (unsigned int) out_bit;
        return out_bit;
    }

    // getbaud() (src/v29tx.c:128-222); constellations: src/v29tx_constellation_maps.h:28-79
    SB_HD void getbaud(float &vre, float &vim)
    {
        const float c16[16][2] =
        {
            { 3.0f,  0.0f}, { 1.0f,  1.0f}, { 0.0f,  3.0f}, {-1.0f,  1.0f}, {-3.0f,  0.0f}, {-1.0f, -1.0f}, { 0.0f, -3.0f}, { 1.0f, -1.0f},
            { 5.0f,  0.0f}, { 3.0f,  3.0f}, { 0.0f,  5.0f}, {-3.0f,  3.0f}, {-5.0f,  0.0f}, {-3.0f, -3.0f}, { 0.0f, -5.0f}, { 3.0f, -3.0f}
        };
        const float abab[6][2] = {{3.0f, -3.0f}, {-3.0f, 0.0f}, {1.0f, -1.0f}, {-3.0f, 0.0f}, {0.0f, -3.0f}, {-3.0f, 0.0f}};
        const float cdcd[6][2] = {{3.0f, 0.0f}, {-3.0f, 3.0f}, {3.0f, 0.0f}, {-1.0f, 1.0f}, {3.0f, 0.0f}, {0.0f, 3.0f}};
        const int steps_9600[8] = {1, 0, 2, 3, 6, 7, 5, 4};
        const int steps_4800[4] = {0, 2, 6, 4};

        if (in_training)
        {
            if (++training_step <= 480 + 48 + 128 + 384)            // V29_TRAINING_SEG_4
            {
                if (training_step <= 480 + 48 + 128)                // V29_TRAINING_SEG_3
                {
                    if (training_step <= 480)                       // talker echo protection: unmodulated carrier
                    {
                        vre = c16[0][0];
                        vim = c16[0][1];
                        return;
                    }
                    if (training_step <= 480 + 48)                  // segment 1: silence
                    {
                        vre = vim = 0.0f;
                        return;
                    }
                    const int k = (training_step & 1) + training_offset;       // segment 2: ABAB
                    vre = abab[k][0];
                    vim = abab[k][1];
                    return;
                }
                // segment 3: CDCD through the 1 + x^-6 + x^-7 training scrambler (the register is a uint8_t)
                const int bit = (int) (training_scramble_reg & 1);
                training_scramble_reg >>= 1;
                training_scramble_reg |= ((((unsigned int) bit ^ training_scramble_reg) & 1) << 6);
                training_scramble_reg &= 0xFFu;
                vre = cdcd[bit + training_offset][0];
                vim = cdcd[bit + training_offset][1];
                return;
            }
            if (training_step == 480 + 48 + 128 + 384 + 48 + 1)     // V29_TRAINING_END + 1: over to the real bits
            {
                real_bits = 1;
                in_training = 0;
            }
            if (training_step == 480 + 48 + 128 + 384 + 48 + 32)    // V29_TRAINING_SHUTDOWN_END
                status |= 2;
        }
        int amp = 0;
        if (bit_rate == 9600  &&  get_scrambled_bit())
            amp = 8;
        int b = get_scrambled_bit();
        b = (b << 1) | get_scrambled_bit();
        if (bit_rate == 4800)
        {
            b = steps_4800[b];
        }
        else
        {
            b = (b << 1) | get_scrambled_bit();
            b = steps_9600[b];
        }
        constellation_state = (constellation_state + b) & 7;
        vre = c16[amp | constellation_state][0];
        vim = c16[amp | constellation_state][1];
    }

    // vec_circular_dot_prodf() with the scalar vec_dot_prodf() (src/vector_float.c:890-939)
    SB_HD float shape(const float *x, const float *y) const
    {
        float za = 0.0f;
        float zb = 0.0f;
        const int first = SBG_V29_TX_STEPS - rrc_step;
        for (int i = 0;  i < SBG_V29_TX_STEPS;  i++)
        {
            if (i < first)
                za = g_fadd(za, g_fmul(x[rrc_step + i], y[i]));
            else
                zb = g_fadd(zb, g_fmul(x[i - first], y[i]));
        }
        return g_fadd(za, zb);
    }

    // v29_tx() (src/v29tx.c:226-296)
    template <class OUT> SB_HD int tx(OUT &out, int len)
    {
        if (training_step >= 480 + 48 + 128 + 384 + 48 + 32)
            return 0;
        for (int sample = 0;  sample < len;  sample++)
        {
            if ((baud_phase += 3) >= 10)
            {
                baud_phase -= 10;
                float vre;
                float vim;
                getbaud(vre, vim);
                rrc_re[rrc_step] = vre;
                rrc_im[rrc_step] = vim;
                if (++rrc_step >= SBG_V29_TX_STEPS)
                    rrc_step = 0;
            }
            const float *row = shaper + (SBG_V29_TX_SETS - 1 - baud_phase)*SBG_V29_TX_STEPS;
            const float xre = shape(rrc_re, row);
            const float xim = shape(rrc_im, row);
            const float zre = sine[(carrier_phase + (1u << 30)) >> 21];        // dds_complexf(), src/dds_float.c:2183-2190
            const float zim = sine[carrier_phase >> 21];
            carrier_phase += (unsigned int) carrier_phase_rate;
            const float famp = g_fadd(g_fmul(xre, zre), -g_fmul(xim, zim));
            out.put(ToneGen::to_amp(g_fmul(famp, gain)));
        }
        return len;
    }
};

struct V29TxArgs
{
    int16_t *amp;
    long long stride;
    int max_samples;
    int channels;
    int zero_fill;
    int *state;                     // [X_COUNT][channels]
    const unsigned char *bits;      // [channels][bits_stride] caller bits, or NULL
    long long bits_stride;
    const float *sine;
    const float *shaper;
    int *lens;
    int carrier_phase_rate;         // dds_phase_ratef(1700.0f)
};

#if defined(__CUDACC__)

// dtmf_tx() for every channel: thread per channel, the sine table in shared memory
__global__ void __launch_bounds__(128) dtmf_tx_kernel(const DtmfTxArgs a)
{
    __shared__ float s_sine[SBG_SINE_WORDS];
    for (int i = threadIdx.x;  i < SBG_SINE_WORDS;  i += blockDim.x)
        s_sine[i] = a.sine[i];
    __syncthreads();
    const int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= a.channels)
        return;
    DtmfTx t;
    GenLoader ld = {a.state, (size_t) a.channels, (size_t) c};
    t.load(ld);
    t.tones.sine = s_sine;
    t.queue = a.queue + c;
    t.qstride = (size_t) a.channels;
    t.row_rate = a.rates;
    t.col_rate = a.rates + 4;
    RowOut out;
    out.begin(a.amp + (long long) c*a.stride);
    const int len = t.tx(out, a.max_samples);
    if (a.zero_fill)
    {
        for (int i = len;  i < a.max_samples;  i++)
            out.put(0);
    }
    out.flush();
    GenStorer st = {a.state, (size_t) a.channels, (size_t) c};
    t.store(st);
    a.lens[c] = len;
}

// mode 0: dtmf_tx_init(); 1: dtmf_tx_set_level (ga = low gain, gb = high gain); 2: dtmf_tx_set_timing (ia = on, ib = off samples)
__global__ void dtmf_tx_ctl_kernel(const DtmfTxArgs a, int first, int count, int mode, float ga, float gb, int ia, int ib)
{
    const int idx = blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= count)
        return;
    const int c = first + idx;
    DtmfTx t;
    GenLoader ld = {a.state, (size_t) a.channels, (size_t) c};
    GenStorer st = {a.state, (size_t) a.channels, (size_t) c};
    t.row_rate = a.rates;
    t.col_rate = a.rates + 4;
    if (mode == 0)
    {
        t.init(ga);
    }
    else
    {
        t.load(ld);
        if (mode == 1)
        {
            t.low_level = ga;
            t.high_level = gb;
        }
        else
        {
            t.on_time = ia;
            t.off_time = ib;
        }
    }
    t.store(st);
}

// dtmf_tx_put() (src/dtmf.c:600-621) per channel: digits[c*dstride .. + lens[c]) (dstride 0: one string for all);
// all or nothing per channel; result[c] = 0 or the number of digits that did not fit
__global__ void dtmf_tx_put_kernel(const DtmfTxArgs a, int first, int count, const unsigned char *digits, long long dstride,
                                   const int *lens, int len_all, int *result)
{
    const int idx = blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= count)
        return;
    const int c = first + idx;
    const size_t C = (size_t) a.channels;
    const int len = (lens)  ?  lens[idx]  :  len_all;
    const int cnt = a.state[(size_t) D_QCNT*C + c];
    const int space = SBG_QUEUE - cnt;
    if (space < len)
    {
        result[idx] = len - space;
        return;
    }
    const int rd = a.state[(size_t) D_QRD*C + c];
    const unsigned char *src = digits + (size_t) idx*dstride;
    for (int i = 0;  i < len;  i++)
        a.queue[(size_t) ((rd + cnt + i) & (SBG_QUEUE - 1))*C + c] = src[i];
    a.state[(size_t) D_QCNT*C + c] = cnt + len;
    result[idx] = 0;
}

// awgn() for every channel, one warp per CTA: the 97-entry shuffle tables of the 32 channels sit in shared memory
// lane-interleaved.  Samples in groups of eight per lane (16-byte loads / stores) where the rows are aligned.
__global__ void __launch_bounds__(32) awgn_kernel(const AwgnArgs a)
{
    __shared__ double s_r[SBG_RAN_TABLE*32];
    const int lane = threadIdx.x;
    const int c = blockIdx.x*32 + lane;
    if (c >= a.channels)
        return;
    const size_t C = (size_t) a.channels;
    Awgn g;
    g.ix1 = a.istate[c];
    g.ix2 = a.istate[C + c];
    g.ix3 = a.istate[2*C + c];
    g.odd = a.istate[3*C + c];
    g.amp2 = a.dstate[c];
    g.rms = a.dstate[C + c];
    g.r = s_r + lane;
    g.rs = 32;
    for (int j = 0;  j < SBG_RAN_TABLE;  j++)
        s_r[j*32 + lane] = a.dstate[(size_t) (2 + j)*C + c];
    int16_t *row = a.amp + (long long) c*a.stride;
    int pos = 0;
    if ((((size_t) row) & 15) == 0)
    {
#pragma unroll 1
        for (  ;  pos + 8 <= a.n;  pos += 8)
        {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (a.add)
                v = *((const uint4 *) (row + pos));
            unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0;  k < 4;  k++)
            {
                const int lo = sat_add16((int) (short) (w[k] & 0xFFFFu), g.sample());
                const int hi = sat_add16((int) (short) (w[k] >> 16), g.sample());
                w[k] = ((unsigned int) lo & 0xFFFFu) | ((unsigned int) hi << 16);
            }
            *((uint4 *) (row + pos)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
#pragma unroll 1
    for (  ;  pos < a.n;  pos++)
        row[pos] = (int16_t) sat_add16((a.add)  ?  (int) row[pos]  :  0, g.sample());
    a.istate[c] = g.ix1;
    a.istate[C + c] = g.ix2;
    a.istate[2*C + c] = g.ix3;
    a.istate[3*C + c] = g.odd;
    a.dstate[c] = g.amp2;
    for (int j = 0;  j < SBG_RAN_TABLE;  j++)
        a.dstate[(size_t) (2 + j)*C + c] = s_r[j*32 + lane];
}

// awgn_init_dbov() (src/awgn.c:131-150) for channels [first, first + count): seeds[idx] (or seed0 + idx), rms from the host
__global__ void awgn_init_kernel(const AwgnArgs a, int first, int count, const int *seeds, int seed0, double rms)
{
    const int idx = blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= count)
        return;
    const int c = first + idx;
    const size_t C = (size_t) a.channels;
    Awgn g;
    g.r = a.dstate + 2*C + c;
    g.rs = C;
    g.ran_init((seeds)  ?  seeds[idx]  :  (seed0 + idx));
    a.istate[c] = g.ix1;
    a.istate[C + c] = g.ix2;
    a.istate[2*C + c] = g.ix3;
    a.istate[3*C + c] = 1;          // odd = true
    a.dstate[c] = 0.0;              // amp2
    a.dstate[C + c] = rms;
}

// tone_gen() for every channel
__global__ void __launch_bounds__(128) tone_gen_kernel(const ToneGenArgs a)
{
    __shared__ float s_sine[SBG_SINE_WORDS];
    for (int i = threadIdx.x;  i < SBG_SINE_WORDS;  i += blockDim.x)
        s_sine[i] = a.sine[i];
    __syncthreads();
    const int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= a.channels)
        return;
    ToneGen t;
    GenLoader ld = {a.state, (size_t) a.channels, (size_t) c};
    tone_gen_load(t, ld);
    t.sine = s_sine;
    RowOut out;
    out.begin(a.amp + (long long) c*a.stride);
    const int len = t.run(out, a.max_samples);
    if (a.zero_fill)
    {
        for (int i = len;  i < a.max_samples;  i++)
            out.put(0);
    }
    out.flush();
    GenStorer st = {a.state, (size_t) a.channels, (size_t) c};
    tone_gen_store(t, st);
    a.lens[c] = len;
}

// tone_gen_init() with descs[idx] (or descs[0] for every channel when `same`) for channels [first, first + count)
__global__ void tone_gen_init_kernel(const ToneGenArgs a, int first, int count, const ToneDesc *descs, int same)
{
    const int idx = blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= count)
        return;
    ToneGen t;
    tone_gen_start(t, descs[(same)  ?  0  :  idx]);
    GenStorer st = {a.state, (size_t) a.channels, (size_t) (first + idx)};
    tone_gen_store(t, st);
}

// v29_tx() for every channel
__global__ void __launch_bounds__(128) v29_tx_kernel(const V29TxArgs a)
{
    __shared__ float s_sine[SBG_SINE_WORDS];
    __shared__ float s_shaper[SBG_V29_TX_SETS*SBG_V29_TX_STEPS];
    for (int i = threadIdx.x;  i < SBG_SINE_WORDS;  i += blockDim.x)
        s_sine[i] = a.sine[i];
    for (int i = threadIdx.x;  i < SBG_V29_TX_SETS*SBG_V29_TX_STEPS;  i += blockDim.x)
        s_shaper[i] = a.shaper[i];
    __syncthreads();
    const int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= a.channels)
        return;
    V29Tx t;
    GenLoader ld = {a.state, (size_t) a.channels, (size_t) c};
    t.load(ld);
    t.sine = s_sine;
    t.shaper = s_shaper;
    t.carrier_phase_rate = a.carrier_phase_rate;
    t.bits = (a.bits)  ?  (a.bits + (long long) c*a.bits_stride)  :  NULL;
    RowOut out;
    out.begin(a.amp + (long long) c*a.stride);
    const int len = t.tx(out, a.max_samples);
    if (a.zero_fill)
    {
        for (int i = len;  i < a.max_samples;  i++)
            out.put(0);
    }
    out.flush();
    GenStorer st = {a.state, (size_t) a.channels, (size_t) c};
    t.store(st);
    a.lens[c] = len;
}

// mode 0: v29_tx_init() (all state zeroed, power = base gain `ga`, restart); 1: v29_tx_restart(ia = bit rate, ib = tep);
// 2: v29_tx_power (ga = base gain); 3: PRBS source, seed seeds[idx] or ia + idx; 4: caller bits, counts[idx] of them
__global__ void v29_tx_ctl_kernel(const V29TxArgs a, int first, int count, int mode, float ga, int ia, int ib,
                                  const unsigned int *seeds, const int *counts)
{
    const int idx = blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= count)
        return;
    const int c = first + idx;
    V29Tx t;
    GenLoader ld = {a.state, (size_t) a.channels, (size_t) c};
    GenStorer st = {a.state, (size_t) a.channels, (size_t) c};
    if (mode == 0)
    {
        t.scramble_reg = 0;
        t.status = 0;
        t.src_mode = 0;
        t.lfsr = 1;
        t.bit_pos = 0;
        t.bit_count = 0;
        t.base_gain = ga;
        t.gain = 0.0f;
        t.bit_rate = ia;
        t.restart(ia, ib);
    }
    else
    {
        t.load(ld);
        if (mode == 1)
        {
            t.restart(ia, ib);
        }
        else if (mode == 2)
        {
            t.base_gain = ga;
            t.set_working_gain();
        }
        else if (mode == 3)
        {
            const unsigned int seed = ((seeds)  ?  seeds[idx]  :  (unsigned int) (ia + idx)) & 0x7FFFFFu;
            t.src_mode = 0;
            t.lfsr = (seed)  ?  seed  :  1u;
        }
        else
        {
            t.src_mode = 1;
            t.bit_pos = 0;
            t.bit_count = counts[idx];
            t.status &= ~1;
        }
    }
    t.store(st);
}

#endif  // __CUDACC__

}  // namespace sbg
