// sb_modem.cuh - what the V.29 and V.17 receiver banks share: strict arithmetic wrappers, the host-side
// generators of the constant tables, the receiver core (signal detect, RRC band-pass FIR, Godard timing,
// AGC, T/2 equalizer buffer, complex equalizer / LMS, carrier loop), the kernels and the bank bookkeeping.
//
// Reference: src/v29rx.c, src/v17rx.c (the two receivers share this structure line for line),
// src/godard.c:144-220, src/power_meter.c:65-69, src/math_fixed.c:158-169, src/dds_float.c:2135-2180,
// src/spandsp/arctan2.h:47-80, src/vector_float.c:890-939 (scalar dot products),
// src/complex_vector_float.c:137-219.
//
// Arithmetic contract: every float operation is an explicitly rounded single operation in the reference's
// operand order (the pinned oracle is the strict build with sequential dot products); integers follow C
// semantics of the reference on x86-64 (arithmetic right shifts, int16 wrap-around, cvttss2si for
// float->int32).  The one libm call in the sample loop, cosf/sinf at the equalizer "spin", is reproduced
// exactly (host_sincosf below).
//
// Everything a receiver does per sample is written as __host__ __device__ code.  The product only ever
// runs it on the GPU (modem_rx_kernel); tests/hostsim compiles the same receiver for the host so that
// the training state machines can be debugged against the oracle in a container without a GPU.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#pragma GCC visibility push(default)
#include "../../include/spandsp_b200_v29.h"
#pragma GCC visibility pop

#define SB_HD __host__ __device__ __forceinline__

#define SBM_FILTER_STEPS    27      // V29_RX_FILTER_STEPS / V17_RX_FILTER_STEPS
#define SBM_EQ_LEN          33      // V29_EQUALIZER_LEN / V17_EQUALIZER_LEN
#define SBM_EQ_PRE_LEN      16
#define SBM_SINE_WORDS      2048    // dds sine table
#define SBM_SQRT_WORDS      100     // 193 uint16 of the fixed_sqrt table, padded to keep 16-byte alignment
#define SBM_IN_RING         32      // samples per lane in the input staging ring (power of two)
#define SBM_RRC_ROW         28      // coefficient rows in shared memory, padded from 27 so that they load as float4
#define SBM_RRC_SKEW        16      // words between the real and the imaginary table: row r of the two then sits 16 banks apart

#define SIG_STATUS_CARRIER_DOWN             (-1)    // src/spandsp/async.h:66-103
#define SIG_STATUS_CARRIER_UP               (-2)
#define SIG_STATUS_TRAINING_IN_PROGRESS     (-3)
#define SIG_STATUS_TRAINING_SUCCEEDED       (-4)
#define SIG_STATUS_TRAINING_FAILED          (-5)

namespace sbm {

// ------------------------------------------------------------------------------------------
// strict arithmetic

#if defined(__CUDA_ARCH__)
SB_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
SB_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
SB_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
SB_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
SB_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
SB_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
SB_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
SB_HD int clz32(unsigned int x) { return __clz((int) x); }
SB_HD unsigned int float_bits(float f) { return __float_as_uint(f); }
SB_HD int float_to_int_rz(float f) { return __float2int_rz(f); }
SB_HD int double_to_int_rz(double f) { return __double2int_rz(f); }
template <class T> SB_HD T ldg(const T *p) { return __ldg(p); }
#else
// Host build (tests/hostsim only): plain IEEE single operations; compiled without FMA contraction.
SB_HD float fmul(float a, float b) { volatile float r = a*b; return r; }
SB_HD float fadd(float a, float b) { volatile float r = a + b; return r; }
SB_HD float fsub(float a, float b) { volatile float r = a - b; return r; }
SB_HD float fdiv(float a, float b) { volatile float r = a/b; return r; }
SB_HD double dmul(double a, double b) { volatile double r = a*b; return r; }
SB_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
SB_HD double dsub(double a, double b) { volatile double r = a - b; return r; }
SB_HD int clz32(unsigned int x) { return __builtin_clz(x); }
SB_HD unsigned int float_bits(float f) { unsigned int u; memcpy(&u, &f, 4); return u; }
SB_HD int float_to_int_rz(float f) { return (int) f; }
SB_HD int double_to_int_rz(double f) { return (int) f; }
template <class T> SB_HD T ldg(const T *p) { return *p; }
#endif

// (int32_t) of a float the way x86-64's cvttss2si does it: out-of-range and NaN give INT_MIN.
SB_HD int f2i(float f)
{
    if (!(f > -2147483904.0f  &&  f < 2147483648.0f))
        return (int) 0x80000000;
    return float_to_int_rz(f);
}

// cosf()/sinf() as the host C library computes them.  The reference calls libm's cosf/sinf once per
// training (src/v29rx.c:618-623, src/v17rx.c:707-711,784-788); the values feed the adaptive loops, whose
// discrete timing decisions amplify a 1-ulp difference into visible (1e-3) excursions of the soft
// symbols, so they have to be reproduced exactly.  Third-party arithmetic: GNU libc 2.39 (the image's
// libm.so.6), sysdeps/ieee754/flt-32/s_cosf.c, s_sinf.c, s_sincosf.h - the "sincosf" of ARM's optimized
// routines: reduce by pi/2 in double with a 2^24-prescaled 2/pi, then an odd/even minimax polynomial in
// double, rounded once to float.  Constants are the published ones (they can be read back from
// __sincosf_table in libm.so.6).  Arguments here are phases in [0, 2*pi), so only the two fast paths
// (|x| < pi/4, |x| < 120) are needed.  tools/check_host_sincosf.py checks this restatement against the
// live libm on 1e6 arguments.
SB_HD double sincosf_poly(double x, double x2, bool cos_table_negated, int n)
{
    const double c0 = (cos_table_negated)  ?  -0x1p0  :  0x1p0;
    const double c1 = (cos_table_negated)  ?  0x1.ffffffd0c621cp-2  :  -0x1.ffffffd0c621cp-2;
    const double c2 = (cos_table_negated)  ?  -0x1.55553e1068f19p-5  :  0x1.55553e1068f19p-5;
    const double c3 = (cos_table_negated)  ?  0x1.6c087e89a359dp-10  :  -0x1.6c087e89a359dp-10;
    const double c4 = (cos_table_negated)  ?  -0x1.99343027bf8c3p-16  :  0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3;
    const double s2 = 0x1.1107605230bc4p-7;
    const double s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0)
    {
        const double x3 = dmul(x, x2);
        const double t1 = dadd(s2, dmul(x2, s3));
        const double x7 = dmul(x3, x2);
        const double sv = dadd(x, dmul(x3, s1));
        return dadd(sv, dmul(x7, t1));
    }
    const double x4 = dmul(x2, x2);
    const double t2 = dadd(c3, dmul(x2, c4));
    const double t1 = dadd(c0, dmul(x2, c1));
    const double x6 = dmul(x4, x2);
    const double cv = dadd(t1, dmul(x4, c2));
    return dadd(cv, dmul(x6, t2));
}

SB_HD unsigned int abstop12(float f)
{
    return (float_bits(f) >> 20) & 0x7FFu;
}

// is_cos: 1 for cosf, 0 for sinf
__host__ __device__ inline float host_sincosf(float y, int is_cos)
{
    double x = (double) y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f))
    {
        if (abstop12(y) < abstop12(0x1p-12f))
            return (is_cos)  ?  1.0f  :  y;
        return (float) sincosf_poly(x, dmul(x, x), false, is_cos);
    }
    const double r = dmul(x, 0x1.45F306DC9C883p+23);
    const int n = (double_to_int_rz(r) + 0x800000) >> 24;
    x = dsub(x, dmul((double) n, 0x1.921FB54442D18p0));
    const double sgn = ((n & 3) == 1  ||  (n & 3) == 2)  ?  -1.0  :  1.0;
    return (float) sincosf_poly(dmul(x, sgn), dmul(x, x), (n & 2) != 0, (is_cos)  ?  (n ^ 1)  :  n);
}

SB_HD float host_cosf(float y) { return host_sincosf(y, 1); }
SB_HD float host_sinf(float y) { return host_sincosf(y, 0); }

// src/spandsp/arctan2.h:47-80
SB_HD int arctan2(float y, float x)
{
    if (y == 0.0f)
        return (x < 0.0f)  ?  (int) 0x80000000  :  0;
    if (x == 0.0f)
        return (y < 0.0f)  ?  (int) 0xC0000000  :  0x40000000;
    const float abs_y = fabsf(y);
    float angle;
    if (x < 0.0f)
        angle = fsub(3.0f, fdiv(fadd(x, abs_y), fsub(abs_y, x)));
    else
        angle = fsub(1.0f, fdiv(fsub(x, abs_y), fadd(abs_y, x)));
    angle = fmul(angle, 536870912.0f);
    if (y < 0.0f)
        angle = -angle;
    return f2i(angle);
}

// dds_phase_to_radians (src/dds_float.c:2103-2106)
SB_HD float phase_to_radians(unsigned int phase)
{
    return fdiv(fmul(fmul((float) phase, 2.0f), 3.1415926f), fmul(65536.0f, 65536.0f));
}

// ------------------------------------------------------------------------------------------
// constant tables: host generators.  The reference builds these with generator programs at build time;
// the library redoes them (including the print-to-decimal / parse-as-float step) so that the float tables
// are bit-identical (tests/test_v29_tables.py, tests/test_v17_tables.py).

static inline float decimal_roundtrip(double v, int decimals)
{
    // The generators print "%.<d>f" into a C header and the compiler parses the literal as float.
    char buf[64];
    snprintf(buf, sizeof(buf), "%.*f", decimals, v);
    return strtof(buf, NULL);
}

// Radix-2 decimation-in-time transform with e^{+j} twiddles from a table built with the
// generator's own constant for pi (src/filter_tools.c:73-124).
struct cplx
{
    double re;
    double im;
};

static inline void dit_transform(cplx *data, cplx *temp, int n, const std::vector<cplx> &circle, int full)
{
    if (n <= 1)
        return;
    const int h = n/2;
    for (int i = 0;  i < h;  i++)
    {
        temp[i] = data[2*i];
        temp[h + i] = data[2*i + 1];
    }
    dit_transform(&temp[0], &data[0], h, circle, full);
    dit_transform(&temp[h], &data[h], h, circle, full);
    int p = 0;
    const int t = full/n;
    for (int i = 0;  i < h;  i++)
    {
        const cplx &w = circle[p];
        const cplx &o = temp[h + i];
        cplx wkt;
        wkt.re = w.re*o.re - w.im*o.im;
        wkt.im = w.re*o.im + w.im*o.re;
        data[i].re = temp[i].re + wkt.re;
        data[i].im = temp[i].im + wkt.im;
        data[h + i].re = temp[i].re - wkt.re;
        data[h + i].im = temp[i].im - wkt.im;
        p += t;
    }
}

// Root raised cosine prototype by frequency sampling (src/filter_tools.c:126-190), then the polyphase
// band-pass sets (src/make_modem_filter.c:155-271).  V.29: 48 sets x 27 taps at 1700 Hz (:401-413);
// V.17: 192 sets x 27 taps at 1800 Hz (:319-332); both 2400 baud, excess bandwidth 0.5.  V.27ter: 12 sets at
// 1200 baud and 8 sets at 1600 baud, 1800 Hz, excess bandwidth 0.5 (:375-400).
static inline void rrc_prototype(std::vector<double> &coeffs, int coeff_sets, int total, double alpha, double beta)
{
    const int SEQ_LEN = 8192;
    const double GEN_PI = 3.1415926535;
    const double f1 = (1.0 - beta)*alpha;
    const double f2 = (1.0 + beta)*alpha;
    const double tau = 0.5/alpha;

    std::vector<cplx> vec(SEQ_LEN);
    std::vector<cplx> temp(SEQ_LEN);
    for (int i = 0;  i < SEQ_LEN;  i++)
        vec[i].re = vec[i].im = 0.0;
    for (int i = 0;  i <= SEQ_LEN/2;  i++)
    {
        const double f = (double) i/(double) SEQ_LEN;
        double v;
        if (f <= f1)
            v = 1.0;
        else if (f <= f2)
            v = 0.5*(1.0 + cos((GEN_PI*tau/beta)*(f - f1)));
        else
            v = 0.0;
        vec[i].re = v;
        vec[i].im = 0.0;
    }
    for (int i = 0;  i <= SEQ_LEN/2;  i++)
        vec[i].re = sqrt(vec[i].re);
    for (int i = 0;  i <= SEQ_LEN/2;  i++)
        vec[i].re *= tau;
    for (int i = 1;  i < SEQ_LEN/2;  i++)
        vec[SEQ_LEN - i] = vec[i];
    std::vector<cplx> circle(SEQ_LEN/2);
    for (int i = 0;  i < SEQ_LEN/2;  i++)
    {
        const double x = (2.0*GEN_PI*i)/(double) SEQ_LEN;
        circle[i].re = cos(x);
        circle[i].im = sin(x);
    }
    dit_transform(vec.data(), temp.data(), SEQ_LEN, circle, SEQ_LEN);
    coeffs.resize(total);
    const int h = (total - 1)/2;
    for (int i = 0;  i < total;  i++)
        coeffs[i] = vec[(SEQ_LEN - h + i) % SEQ_LEN].re/(double) SEQ_LEN;
    // unity DC gain (src/make_modem_filter.c:77-86,175-184)
    double gain = 0.0;
    for (int i = coeff_sets/2;  i < total;  i += coeff_sets)
        gain += coeffs[i];
    for (int i = 0;  i < total;  i++)
        coeffs[i] /= gain;
}

static inline void make_rx_rrc(std::vector<float> &re, std::vector<float> &im, int coeff_sets, double carrier_hz, double baud_rate = 2400.0)
{
    const double GEN_PI = 3.1415926535;
    const int per_filter = SBM_FILTER_STEPS;
    const int total = coeff_sets*per_filter + 1;
    std::vector<double> coeffs;
    rrc_prototype(coeffs, coeff_sets, total, baud_rate/(2.0*(double) (coeff_sets*8000)), 0.5);
    double carrier = carrier_hz;
    carrier *= 2.0*GEN_PI/8000;
    re.assign(coeff_sets*per_filter, 0.0f);
    im.assign(coeff_sets*per_filter, 0.0f);
    for (int j = 0;  j < coeff_sets;  j++)
    {
        for (int i = 0;  i < per_filter;  i++)
        {
            const int m = i - (per_filter >> 1);
            const int x = i*coeff_sets + j;
            re[j*per_filter + i] = decimal_roundtrip(coeffs[x]*cos(carrier*m), 10);
            im[j*per_filter + i] = decimal_roundtrip(coeffs[x]*sin(carrier*m), 10);
        }
    }
}

// The transmit pulse shaper (src/make_modem_filter.c:48-151): baseband root raised cosine, `coeff_sets` polyphase
// sets of `per_filter` taps, alpha = 1/(2*coeff_sets).  V.29: 10 sets x 9 taps, excess bandwidth 0.25 (:401-413).
static inline void make_tx_rrc(std::vector<float> &taps, int coeff_sets, int per_filter, double excess_bandwidth)
{
    const int total = coeff_sets*per_filter + 1;
    std::vector<double> coeffs;
    rrc_prototype(coeffs, coeff_sets, total, 1.0/(2.0*(double) coeff_sets), excess_bandwidth);
    taps.assign(coeff_sets*per_filter, 0.0f);
    for (int j = 0;  j < coeff_sets;  j++)
    {
        for (int i = 0;  i < per_filter;  i++)
            taps[j*per_filter + i] = decimal_roundtrip(coeffs[i*coeff_sets + j], 10);
    }
}

// src/dds_float.c:51-2101: sin(2*pi*i/2048) as 8-decimal literals.
static inline void make_sine_table(std::vector<float> &t)
{
    t.resize(2048);
    for (int i = 0;  i < 2048;  i++)
        t[i] = decimal_roundtrip(sin(2.0*M_PI*(double) i/2048.0), 8);
}

// src/make_math_fixed_tables.c:59-72
static inline void make_sqrt_table(std::vector<unsigned short> &t)
{
    t.resize(193);
    for (int i = 64;  i <= 256;  i++)
    {
        int v = (int) (sqrt(i/256.0)*65536.0 + 0.5);
        if (v > 65535)
            v = 65535;
        t[i - 64] = (unsigned short) v;
    }
}

// src/make_modem_godard_descriptor.c:60-80; arguments from src/Makefile.am (V.29: 1700.0 2400.0 0.99
// 1000.0 30.0 5 1; V.17: 1800.0 2400.0 0.99 1000.0 100.0 15 1); floats are printed with 6 decimals.
struct godard_desc_t
{
    float low[3];
    float high[3];
    float mixed3;
    float coarse_trigger;
    float fine_trigger;
    int coarse_step;
    int fine_step;
};

static inline void make_godard(godard_desc_t &g, double carrier, double fine_trigger, int coarse_step)
{
    const double alpha = 0.99;
    const double low_edge = 2.0*M_PI*(carrier - 2400.0/2.0)/8000.0;
    const double high_edge = 2.0*M_PI*(carrier + 2400.0/2.0)/8000.0;
    g.low[0] = decimal_roundtrip(2.0*alpha*cos(low_edge), 6);
    g.high[0] = decimal_roundtrip(2.0*alpha*cos(high_edge), 6);
    g.low[1] = g.high[1] = decimal_roundtrip(-alpha*alpha, 6);
    g.low[2] = decimal_roundtrip(-alpha*sin(low_edge), 6);
    g.high[2] = decimal_roundtrip(-alpha*sin(high_edge), 6);
    g.mixed3 = decimal_roundtrip(-alpha*alpha*(sin(high_edge)*cos(low_edge) - sin(low_edge)*cos(high_edge)), 6);
    g.coarse_trigger = decimal_roundtrip(1000.0, 6);
    g.fine_trigger = decimal_roundtrip(fine_trigger, 6);
    g.coarse_step = coarse_step;
    g.fine_step = 1;
}

// src/power_meter.c:86-96
static inline int host_power_meter_level_dbm0(float level)
{
    float l;

    level -= (3.14f + 3.02f);
    if (level > 0.0)
        level = 0.0;
    l = powf(10.0f, level/10.0f)*(32767.0f*32767.0f);
    return (int) l;
}

// DDS_PHASE_RATE / DDS_PHASE (src/spandsp/dds.h:31-32)
static inline int32_t host_dds_phase_rate(float hz) { return (int32_t) (hz*65536.0f*65536.0f/8000); }
static inline int32_t host_dds_phase(float angle)
{
    return (int32_t) ((uint32_t) (((angle < 0.0f)  ?  (360.0f + angle)  :  angle)*65536.0f*65536.0f/360.0f));
}

// ------------------------------------------------------------------------------------------
// device side: what every receiver needs

struct CoreConsts
{
    const float *rrc_re;                // [COEFF_SETS][27]
    const float *rrc_im;
    const float *sine;                  // [2048]
    const unsigned short *sqrt_tab;     // [193]
    float g_low[3];
    float g_high[3];
    float g_mixed3;
    float g_coarse_trigger;
    float g_fine_trigger;
    int g_coarse_step;
    int g_fine_step;
    int rate_nominal;                   // DDS_PHASE_RATE(carrier)
    int rate_low;                       // DDS_PHASE_RATE(carrier - 20)
    int rate_high;                      // DDS_PHASE_RATE(carrier + 20)
    float agc_initial;                  // (target/RX_PULSESHAPER_GAIN)/735.0f
    float agc_target;                   // target/RX_PULSESHAPER_GAIN
};

// Per-channel state in global memory, structure of arrays: field f of channel c at [f*C + c].
enum
{
    F_AGC = 0, F_AGC_SAVE, F_EQ_DELTA, F_TRAINING_ERROR, F_TRACK_P, F_TRACK_I,
    F_LBE0, F_LBE1, F_HBE0, F_HBE1, F_DC0, F_DC1, F_BAUD_PHASE,
    F_EQ_COEFF,                                     // 66
    F_EQ_COEFF_SAVE = F_EQ_COEFF + 2*SBM_EQ_LEN,    // 66
    F_EQ_BUF = F_EQ_COEFF_SAVE + 2*SBM_EQ_LEN,      // 66
    F_RRC = F_EQ_BUF + 2*SBM_EQ_LEN,                // 27
    F_CORE_COUNT = F_RRC + SBM_FILTER_STEPS
};

enum
{
    I_BIT_RATE = 0, I_RRC_STEP, I_SCRAMBLE, I_STAGE, I_TRAIN_COUNT,
    I_LAST_SAMPLE, I_SIGNAL_PRESENT, I_DROP_PENDING, I_LOW_SAMPLES, I_HIGH_SAMPLE, I_CARRIER_PHASE, I_PHASE_RATE,
    I_PHASE_RATE_SAVE, I_POWER, I_ON_POWER, I_OFF_POWER, I_EQ_STEP, I_EQ_PUT_STEP, I_EQ_SKIP, I_BAUD_HALF,
    I_LAST_ANGLE0, I_LAST_ANGLE1, I_TOTAL_TIMING,
    I_DIFF_ANGLES,                      // 16
    I_CORE_COUNT = I_DIFF_ANGLES + 16
};

struct ModemArgs
{
    const int16_t *amp;
    long long stride;
    int n;
    int channels;
    float *fstate;
    int *istate;
    // The put_bit stream of a call, per channel: the data bits packed 32 to a word (first bit = bit 0), the status
    // reports (negative SIG_STATUS_* values through put_bit, src/v29rx.c:171-178) as {position in the put_bit
    // sequence, value} pairs beside them.  nbits = put_bit calls (bits + reports), nstatus = reports.
    unsigned int *words;                // [channel][words_cap]
    long long words_cap;
    int *status;                        // [channel][status_cap][2]
    long long status_cap;
    int *nbits;                         // [channel]
    int *nstatus;                       // [channel]
    span_b200_v29_symbol_t *syms;       // [channel][sym_cap], or NULL
    long long sym_cap;
    int *nsyms;
    int append;                         // the call continues the previous one's output (a long call fed in pieces): counts and
                                        // the unfinished word are taken up where that call left them
};

// State visitors: one list of fields (visit()) serves loading and storing.  Every field is visited by
// reference, whether it lives in a register or in this lane's column of shared memory.
struct StateLoader
{
    const float *F;
    const int *I;
    size_t C;
    size_t c;
    SB_HD void f(int slot, float &v) { v = F[(size_t) slot*C + c]; }
    SB_HD void i(int slot, int &v) { v = I[(size_t) slot*C + c]; }
    SB_HD void u(int slot, unsigned int &v) { v = (unsigned int) I[(size_t) slot*C + c]; }
};

struct StateStorer
{
    float *F;
    int *I;
    size_t C;
    size_t c;
    SB_HD void f(int slot, float &v) { F[(size_t) slot*C + c] = v; }
    SB_HD void i(int slot, int &v) { I[(size_t) slot*C + c] = v; }
    SB_HD void u(int slot, unsigned int &v) { I[(size_t) slot*C + c] = (int) v; }
};

// One receiver.  Scalars live in registers; the per-channel arrays live in shared memory,
// lane-interleaved (element e of lane l at [e*32 + l]) so that any per-lane index is conflict-free.
// D is the concrete receiver (CRTP): it supplies restart_after_carrier_down() and process_baud().
// LPC = lanes per channel.  1: one thread runs one receiver (every receiver; also the host build).  4: four lanes
// share one receiver - the scalar part of the receiver runs replicated on all four (same inputs, same arithmetic,
// same results), the three loops that carry the flops are split between them:
//   * the RRC FIR pair: lanes {real, imaginary} x {first, second segment of the reference's circular dot product};
//   * the complex equalizer dot product: the same four roles (real / imaginary accumulator x segment);
//   * the LMS update: taps i = lane, lane + 4, ...
// The segment split is exact: vec_circular_dot_prodf() sums the taps before and after the ring's physical wrap in two
// separate chains and adds the two sums, so the chains are independent.  To keep all lanes on one instruction
// stream whatever the ring position, the rings are stored between two zero areas, [zeros | ring | zeros]: a lane
// that owns the first segment reads the ring from the current position on and runs into the trailing zeros, a lane
// that owns the second segment starts in the leading zeros and runs into the ring - every lane executes all N taps,
// and a product with a zero sample is +-0, which leaves a sum that started at +0 unchanged (x + (+-0) = x, and
// (+0) + (-0) = +0), so each chain has exactly the value the reference's shorter loop gives.
template <class D, int COEFF_SETS, int EQ_LEN = SBM_EQ_LEN, int LPC = 1>
struct RxCore
{
    static const int SETS = COEFF_SETS;
    static const int EQ_TAPS = EQ_LEN;      // equalizer length (33: V.29, V.17; 32: V.27ter); arrays are sized for 33
    static const int LANES = LPC;
    static const int LS = 32/LPC;           // receivers per warp = element pitch of the per-receiver arrays in shared memory
    // Symbol clock of xxx_rx_fillin(): coefficient-set steps per sample and per T/2 (a receiver with several
    // baud rates overrides these)
    static int fillin_sets(int) { return COEFF_SETS; }
    static int fillin_half_baud(int) { return COEFF_SETS*10/(3*2); }
    // float state
    float agc_scaling, agc_scaling_save, eq_delta, training_error, track_p, track_i;
    float lbe0, lbe1, hbe0, hbe1, dc0, dc1, baud_phase;
    // int state
    int bit_rate, rrc_step, training_stage, training_count;
    unsigned int scramble_reg, carrier_phase;
    int last_sample, signal_present, drop_pending, low_samples, high_sample;
    int phase_rate, phase_rate_save, power, on_power, off_power;
    int eq_step, eq_put_step, eq_skip, baud_half, total_timing;
    int last_angle0, last_angle1;
    // shared-memory arrays of this lane (dynamic indexing would force the whole receiver out of registers
    // if they were member arrays).  Lane-interleaved: element e of lane l at [e*32 + l] (complex: float2
    // elements), so any per-lane index is bank-conflict free.  The two rings (RRC history, equalizer
    // buffer) are kept twice, the copy of slot k at k + N, so that the N taps that start at any ring
    // position are contiguous and the inner loops need no wrap-around arithmetic.
    const float *sine;              // [2048] dds sine table (shared memory copy)
    const unsigned short *sqrt_tab; // [193] fixed_sqrt table (shared memory copy)
    int *diff_angles;       // [16]
    float2 *eq_coeff;       // [33]
    float2 *eq_buf;         // LPC 1: [66], ring of 33 kept twice.  LPC 4: the ring, with 33 zeros before and after it
    float *rrc;             // LPC 1: [54], ring of 27 kept twice.  LPC 4: the ring, with 27 zeros before and after it
    float2 *eq_coef_re;     // LPC 4: (yr, -yi) per tap: what the lanes of the real accumulator multiply with
    float2 *eq_coef_im;     // LPC 4: (yi, yr) per tap: the imaginary accumulator's
    int lms_pending;        // LPC 4: an LMS update / a save of the coefficients is owed for this baud (done by run())
    int save_pending;
    float lms_ere;
    float lms_eim;
    float h_zre;            // LPC 4: the equalizer output of this baud, computed by run() with the warp converged
    float h_zim;
    int sub;                // LPC 4: this lane's role, 0 = real / first segment, 1 = real / second, 2 = imag / first, 3 = imag / second
    unsigned int gmask;     // LPC 4: the four lanes of this receiver
    // outputs.  Device kernels pack the data bits (words / status); the host build writes one byte per put_bit call.
    signed char *bits;
    int nbits;
    int bits_cap;
    unsigned int *words;
    int words_cap;
    int nwords;
    unsigned int bit_acc;
    int bit_fill;
    int *status;
    int status_cap;
    int nstatus;
    span_b200_v29_symbol_t *syms;
    int nsyms;
    int sym_cap;
    // channel
    int c;
    int channels;
    float *fstate;

    SB_HD D &self() { return *static_cast<D *>(this); }

    // words per receiver of the lane-interleaved block
    static const int CORE_LANE_WORDS = (LPC == 1)  ?  (2*SBM_EQ_LEN + 4*SBM_EQ_LEN + 2*SBM_FILTER_STEPS + 16)
                                                   :  (6*SBM_EQ_LEN + 6*SBM_EQ_LEN + 3*SBM_FILTER_STEPS + 16 + 1);
    static const int IN_RING_WORDS = SBM_IN_RING/2;     // per receiver, contiguous (not interleaved): cp.async needs 16 bytes in a row

    // block: the warp's lane-interleaved area (element e of receiver r at [e*LS + r])
    SB_HD void bind_core(float *block, int lane)
    {
        if (LPC == 1)
        {
            eq_coeff = ((float2 *) block) + lane;
            eq_buf = ((float2 *) (block + (2*SBM_EQ_LEN)*32)) + lane;
            rrc = block + (6*SBM_EQ_LEN)*32 + lane;
            diff_angles = (int *) (block + (6*SBM_EQ_LEN + 2*SBM_FILTER_STEPS)*32) + lane;
            eq_coef_re = eq_coef_im = NULL;
            sub = 0;
            gmask = 0xFFFFFFFFu;
        }
        else
        {
            const int r = lane/LPC;
            float2 *p2 = (float2 *) block;
            eq_coeff = p2 + r;
            eq_coef_re = p2 + SBM_EQ_LEN*LS + r;
            eq_coef_im = p2 + 2*SBM_EQ_LEN*LS + r;
            eq_buf = p2 + 3*SBM_EQ_LEN*LS + EQ_LEN*LS + r;              // the ring; EQ_LEN zeros on either side
            float *pf = block + 12*SBM_EQ_LEN*LS;
            rrc = pf + SBM_FILTER_STEPS*LS + r;                          // the ring; 27 zeros on either side
            diff_angles = (int *) (pf + 3*SBM_FILTER_STEPS*LS) + r;
            sub = lane%LPC;
            gmask = ((1u << LPC) - 1u) << (lane - sub);
        }
    }

    // LPC 4: the zero areas around the two rings (written once per launch)
    SB_HD void zero_pads()
    {
        if (LPC > 1)
        {
            for (int k = 0;  k < EQ_LEN;  k++)
            {
                eq_buf[(k - EQ_LEN)*LS] = make_float2(0.0f, 0.0f);
                eq_buf[(k + EQ_LEN)*LS] = make_float2(0.0f, 0.0f);
            }
            for (int k = 0;  k < SBM_FILTER_STEPS;  k++)
            {
                rrc[(k - SBM_FILTER_STEPS)*LS] = 0.0f;
                rrc[(k + SBM_FILTER_STEPS)*LS] = 0.0f;
            }
        }
    }

    // One equalizer coefficient, in every form that is kept of it
    SB_HD void set_coef(int i, float yr, float yi)
    {
        eq_coeff[i*LS] = make_float2(yr, yi);
        if (LPC > 1)
        {
            eq_coef_re[i*LS] = make_float2(yr, -yi);
            eq_coef_im[i*LS] = make_float2(yi, yr);
        }
    }

    SB_HD void group_sync()
    {
#if defined(__CUDA_ARCH__)
        if (LPC > 1)
            __syncwarp(gmask);
#endif
    }

    template <class V> SB_HD void visit_core(V &v)
    {
        v.f(F_AGC, agc_scaling);
        v.f(F_AGC_SAVE, agc_scaling_save);
        v.f(F_EQ_DELTA, eq_delta);
        v.f(F_TRAINING_ERROR, training_error);
        v.f(F_TRACK_P, track_p);
        v.f(F_TRACK_I, track_i);
        v.f(F_LBE0, lbe0);
        v.f(F_LBE1, lbe1);
        v.f(F_HBE0, hbe0);
        v.f(F_HBE1, hbe1);
        v.f(F_DC0, dc0);
        v.f(F_DC1, dc1);
        v.f(F_BAUD_PHASE, baud_phase);
        for (int k = 0;  k < EQ_LEN;  k++)
        {
            v.f(F_EQ_COEFF + 2*k, eq_coeff[k*LS].x);
            v.f(F_EQ_COEFF + 2*k + 1, eq_coeff[k*LS].y);
            v.f(F_EQ_BUF + 2*k, eq_buf[k*LS].x);
            v.f(F_EQ_BUF + 2*k + 1, eq_buf[k*LS].y);
        }
        for (int k = 0;  k < SBM_FILTER_STEPS;  k++)
            v.f(F_RRC + k, rrc[k*LS]);
        v.i(I_BIT_RATE, bit_rate);
        v.i(I_RRC_STEP, rrc_step);
        v.u(I_SCRAMBLE, scramble_reg);
        v.i(I_STAGE, training_stage);
        v.i(I_TRAIN_COUNT, training_count);
        v.i(I_LAST_SAMPLE, last_sample);
        v.i(I_SIGNAL_PRESENT, signal_present);
        v.i(I_DROP_PENDING, drop_pending);
        v.i(I_LOW_SAMPLES, low_samples);
        v.i(I_HIGH_SAMPLE, high_sample);
        v.u(I_CARRIER_PHASE, carrier_phase);
        v.i(I_PHASE_RATE, phase_rate);
        v.i(I_PHASE_RATE_SAVE, phase_rate_save);
        v.i(I_POWER, power);
        v.i(I_ON_POWER, on_power);
        v.i(I_OFF_POWER, off_power);
        v.i(I_EQ_STEP, eq_step);
        v.i(I_EQ_PUT_STEP, eq_put_step);
        v.i(I_EQ_SKIP, eq_skip);
        v.i(I_BAUD_HALF, baud_half);
        v.i(I_LAST_ANGLE0, last_angle0);
        v.i(I_LAST_ANGLE1, last_angle1);
        v.i(I_TOTAL_TIMING, total_timing);
        for (int k = 0;  k < 16;  k++)
            v.i(I_DIFF_ANGLES + k, diff_angles[k*LS]);
    }

    // After loading: LPC 1 fills the second copy of the two rings; LPC 4 derives the two extra forms of the coefficients
    SB_HD void mirror_rings()
    {
        if (LPC == 1)
        {
            for (int k = 0;  k < EQ_LEN;  k++)
                eq_buf[(k + EQ_LEN)*LS] = eq_buf[k*LS];
            for (int k = 0;  k < SBM_FILTER_STEPS;  k++)
                rrc[(k + SBM_FILTER_STEPS)*LS] = rrc[k*LS];
        }
        else
        {
            for (int k = 0;  k < EQ_LEN;  k++)
            {
                const float2 y = eq_coeff[k*LS];
                set_coef(k, y.x, y.y);
            }
        }
    }

    SB_HD void rrc_clear()
    {
        for (int i = 0;  i < ((LPC == 1)  ?  2  :  1)*SBM_FILTER_STEPS;  i++)
            rrc[i*LS] = 0.0f;
        rrc_step = 0;
    }

    // One put_bit() call of the reference: a data bit, or (negative) a status report
    SB_HD void out_bit(int v)
    {
        if (words == NULL)
        {
            if (nbits < bits_cap)
                bits[nbits] = (signed char) v;
        }
        else if (v < 0)
        {
            if (sub == 0  &&  nstatus < status_cap)
            {
                status[2*nstatus] = nbits;
                status[2*nstatus + 1] = v;
            }
            nstatus++;
        }
        else
        {
            bit_acc |= (unsigned int) (v & 1) << bit_fill;
            if (++bit_fill == 32)
            {
                if (sub == 0  &&  nwords < words_cap)
                    words[nwords] = bit_acc;
                nwords++;
                bit_acc = 0;
                bit_fill = 0;
            }
        }
        nbits++;
    }

    // n (1..4) data bits at once, first bit in bit 0 of v
    SB_HD void out_bits(unsigned int v, int n)
    {
        if (words == NULL)
        {
            for (int i = 0;  i < n;  i++)
                out_bit((int) ((v >> i) & 1u));
            return;
        }
        const unsigned long long acc = (unsigned long long) bit_acc | ((unsigned long long) v << bit_fill);
        bit_fill += n;
        bit_acc = (unsigned int) acc;
        if (bit_fill >= 32)
        {
            if (sub == 0  &&  nwords < words_cap)
                words[nwords] = bit_acc;
            nwords++;
            bit_acc = (unsigned int) (acc >> 32);
            bit_fill -= 32;
        }
        nbits += n;
    }

    // End of a call: the bits of an unfinished word
    SB_HD void out_flush()
    {
        if (words != NULL  &&  bit_fill > 0)
        {
            if (sub == 0  &&  nwords < words_cap)
                words[nwords] = bit_acc;
            nwords++;
            bit_acc = 0;
            bit_fill = 0;
        }
    }

    // src/v29rx.c:171-178, src/v17rx.c:181-189: without a status handler the status goes through put_bit
    SB_HD void report_status(int status)
    {
        out_bit(status);
    }

    SB_HD void report_symbol(float zre, float zim, float tre, float tim, int state)
    {
        if (syms)
        {
            if (sub == 0  &&  nsyms < sym_cap)
            {
                span_b200_v29_symbol_t s;
                s.re = zre;
                s.im = zim;
                s.target_re = tre;
                s.target_im = tim;
                s.state = state;
                s.bit_pos = nbits;
                syms[nsyms] = s;
            }
            nsyms++;
        }
    }

    // src/v29rx.c:214-258, src/v17rx.c:219-264
    // (with LPC 4 all lanes of a receiver write the same values to the same places: harmless)
    SB_HD void equalizer_reset()
    {
        for (int i = 0;  i < EQ_LEN;  i++)
            set_coef(i, 0.0f, 0.0f);
        for (int i = 0;  i < ((LPC == 1)  ?  2  :  1)*EQ_LEN;  i++)
            eq_buf[i*LS] = make_float2(0.0f, 0.0f);
        set_coef(SBM_EQ_PRE_LEN, 3.0f, 0.0f);
        eq_put_step = COEFF_SETS*10/(3*2) - 1;
        eq_step = 0;
    }

    SB_HD void equalizer_restore()
    {
        for (int i = 0;  i < EQ_LEN;  i++)
        {
            set_coef(i, fstate[(size_t) (F_EQ_COEFF_SAVE + 2*i)*channels + c],
                        fstate[(size_t) (F_EQ_COEFF_SAVE + 2*i + 1)*channels + c]);
        }
        for (int i = 0;  i < ((LPC == 1)  ?  2  :  1)*EQ_LEN;  i++)
            eq_buf[i*LS] = make_float2(0.0f, 0.0f);
        eq_put_step = COEFF_SETS*10/(3*2) - 1;
        eq_step = 0;
    }

    SB_HD void equalizer_save()
    {
        if (LPC > 1)
        {
            save_pending = 1;               // after the deferred LMS update of this baud (run())
            return;
        }
        equalizer_save_now();
    }

    SB_HD void equalizer_save_now()
    {
        if (sub == 0)
        {
            for (int i = 0;  i < EQ_LEN;  i++)
            {
                const float2 y = eq_coeff[i*LS];
                fstate[(size_t) (F_EQ_COEFF_SAVE + 2*i)*channels + c] = y.x;
                fstate[(size_t) (F_EQ_COEFF_SAVE + 2*i + 1)*channels + c] = y.y;
            }
        }
    }

    // godard_ted_init (src/godard.c:222-240)
    SB_HD void godard_init()
    {
        lbe0 = lbe1 = hbe0 = hbe1 = dc0 = dc1 = baud_phase = 0.0f;
        total_timing = 0;
    }

    // The real and the imaginary RRC band-pass FIR of one sample, together (src/v29rx.c:914,946,
    // src/v17rx.c:1259,1290).  Each is vec_circular_dot_prodf (src/vector_float.c:932-939) with the scalar
    // vec_dot_prodf (:890-900): two segments - the taps before and after the ring's physical wrap - each
    // summed in order from 0.0f, then added.  rrc_step is the position after the insert, i.e. the oldest
    // sample; coefficient rows are padded to 28 floats so that they load as float4.
    SB_HD void rrc_dot2(const float *row_re, const float *row_im, float &v_re, float &v_im)
    {
#if defined(__CUDA_ARCH__)
        if (LPC > 1)
        {
            // Four lanes, four chains: {real, imaginary} x {first, second segment}.  A lane of the first segment reads
            // the ring from rrc_step on and runs into the zeros behind it; a lane of the second segment starts in the
            // zeros before the ring (see the note at the top of RxCore).
            const float *x = rrc + (rrc_step + ((sub & 1)  ?  -SBM_FILTER_STEPS  :  0))*LS;
            const float4 *c4 = (const float4 *) ((sub & 2)  ?  row_im  :  row_re);
            float z = 0.0f;
#pragma unroll
            for (int q = 0;  q < 7;  q++)
            {
                const float4 cc = c4[q];
                const float cs[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
                for (int e = 0;  e < 4;  e++)
                {
                    const int i = 4*q + e;
                    if (i < SBM_FILTER_STEPS)
                        z = fadd(z, fmul(x[i*LS], cs[e]));
                }
            }
            // first + second segment (the partner lane holds the other one; addition commutes), then the other
            // component from the lane two further on.  Every lane of the warp is here (run() keeps the warp converged
            // around this call), so the shuffles take the full mask.
            // (three independent shuffles instead of two dependent butterfly steps: one shuffle latency, not two)
            const float p1 = __shfl_xor_sync(0xFFFFFFFFu, z, 1);
            const float p2 = __shfl_xor_sync(0xFFFFFFFFu, z, 2);
            const float p3 = __shfl_xor_sync(0xFFFFFFFFu, z, 3);
            z = fadd(z, p1);
            const float o = fadd(p2, p3);
            v_re = (sub & 2)  ?  o  :  z;
            v_im = (sub & 2)  ?  z  :  o;
            return;
        }
#endif
        float zar = 0.0f, zbr = 0.0f, zai = 0.0f, zbi = 0.0f;
        const int first = SBM_FILTER_STEPS - rrc_step;      // taps in the first segment
        const float *x = rrc + rrc_step*LS;
        const float4 *cr4 = (const float4 *) row_re;
        const float4 *ci4 = (const float4 *) row_im;
#pragma unroll
        for (int q = 0;  q < 7;  q++)
        {
            const float4 cr = cr4[q];
            const float4 ci = ci4[q];
            const float crs[4] = {cr.x, cr.y, cr.z, cr.w};
            const float cis[4] = {ci.x, ci.y, ci.z, ci.w};
#pragma unroll
            for (int e = 0;  e < 4;  e++)
            {
                const int i = 4*q + e;
                if (i < SBM_FILTER_STEPS)
                {
                    const float xv = x[i*LS];
                    const float pr = fmul(xv, crs[e]);
                    const float pi = fmul(xv, cis[e]);
                    if (i < first)
                    {
                        zar = fadd(zar, pr);
                        zai = fadd(zai, pi);
                    }
                    else
                    {
                        zbr = fadd(zbr, pr);
                        zbi = fadd(zbi, pi);
                    }
                }
            }
        }
        v_re = fadd(zar, zbr);
        v_im = fadd(zai, zbi);
    }

    // src/complex_vector_float.c:137-150,187-196
    SB_HD void equalizer_get(float &zre, float &zim)
    {
#if defined(__CUDA_ARCH__)
        if (LPC > 1)
        {
            // The same four roles.  The real lanes multiply with (yr, -yi): xr*yr + xi*(-yi) is xr*yr - xi*yi with the
            // same two roundings (negation is exact); the imaginary lanes with (yi, yr): xr*yi + xi*yr.
            const float2 *xb = eq_buf + (eq_step + ((sub & 1)  ?  -EQ_LEN  :  0))*LS;
            const float2 *yb = (sub & 2)  ?  eq_coef_im  :  eq_coef_re;
            float z = 0.0f;
#pragma unroll
            for (int i = 0;  i < EQ_LEN;  i++)
            {
                const float2 x = xb[i*LS];
                const float2 y = yb[i*LS];
                z = fadd(z, fadd(fmul(x.x, y.x), fmul(x.y, y.y)));
            }
            const float p1 = __shfl_xor_sync(0xFFFFFFFFu, z, 1);
            const float p2 = __shfl_xor_sync(0xFFFFFFFFu, z, 2);
            const float p3 = __shfl_xor_sync(0xFFFFFFFFu, z, 3);
            z = fadd(z, p1);
            const float o = fadd(p2, p3);
            zre = (sub & 2)  ?  o  :  z;
            zim = (sub & 2)  ?  z  :  o;
            return;
        }
#endif
        float are = 0.0f, aim = 0.0f, bre = 0.0f, bim = 0.0f;
        const int first = EQ_LEN - eq_step;
        const float2 *xb = eq_buf + eq_step*LS;
#pragma unroll
        for (int i = 0;  i < EQ_LEN;  i++)
        {
            const float2 x = xb[i*LS];
            const float2 y = eq_coeff[i*LS];
            const float pr = fsub(fmul(x.x, y.x), fmul(x.y, y.y));
            const float pi = fadd(fmul(x.x, y.y), fmul(x.y, y.x));
            if (i < first)
            {
                are = fadd(are, pr);
                aim = fadd(aim, pi);
            }
            else
            {
                bre = fadd(bre, pr);
                bim = fadd(bim, pi);
            }
        }
        zre = fadd(are, bre);
        zim = fadd(aim, bim);
    }

    // src/v29rx.c:281-290, src/v17rx.c:296-307 + src/complex_vector_float.c:199-219
    SB_HD void tune_equalizer(float zre, float zim, float tre, float tim)
    {
        const float ere = fmul(fsub(tre, zre), eq_delta);
        const float eim = fmul(fsub(tim, zim), eq_delta);
        if (LPC > 1)
        {
            // Deferred: run() does the update with the whole warp once the per-baud work of all receivers is through
            // (each receiver adapts every tenth baud in data mode, at its own bauds: a warp would otherwise walk the
            // 33 taps with four lanes almost every baud).
            lms_pending = 1;
            lms_ere = ere;
            lms_eim = eim;
            return;
        }
        const float2 *xb = eq_buf + eq_step*LS;
#pragma unroll
        for (int i = 0;  i < EQ_LEN;  i++)
        {
            const float2 x = xb[i*LS];
            float2 y = eq_coeff[i*LS];
            y.x = fadd(fmul(y.x, 0.9999f), fadd(fmul(x.y, eim), fmul(x.x, ere)));
            y.y = fadd(fmul(y.y, 0.9999f), fsub(fmul(x.x, eim), fmul(x.y, ere)));
            eq_coeff[i*LS] = y;
        }
    }

    // The equalizer "spin" (src/v29rx.c:618-625, src/v17rx.c:707-713,784-790); LPC 1: both ring copies
    SB_HD void spin_equalizer_buffer(unsigned int phase_step)
    {
        const float p = phase_to_radians(phase_step);
        const float cr = host_cosf(p);
        const float ci = -host_sinf(p);
        group_sync();                       // a read-modify-write of shared data: one lane does it
        if (sub == 0)
        {
            for (int q = 0;  q < ((LPC == 1)  ?  2  :  1)*EQ_LEN;  q++)
            {
                const float2 x = eq_buf[q*LS];
                eq_buf[q*LS] = make_float2(fsub(fmul(x.x, cr), fmul(x.y, ci)), fadd(fmul(x.x, ci), fmul(x.y, cr)));
            }
        }
        group_sync();
    }

    // src/v29rx.c:297-331, src/v17rx.c:313-338
    SB_HD void track_carrier(float zre, float zim, float tre, float tim)
    {
        const float error = fsub(fmul(zim, tre), fmul(zre, tim));
        phase_rate += f2i(fmul(track_i, error));
        carrier_phase += (unsigned int) f2i(fmul(track_p, error));
    }

    // src/godard.c:165-220
    SB_HD int godard_per_baud(const CoreConsts &k)
    {
        float v = fadd(fsub(fmul(fmul(lbe1, hbe0), k.g_low[2]), fmul(fmul(lbe0, hbe1), k.g_high[2])),
                       fmul(fmul(lbe1, hbe1), k.g_mixed3));
        const float p = fsub(v, dc1);
        dc1 = dc0;
        dc0 = v;
        baud_phase = fsub(baud_phase, p);
        v = fabsf(baud_phase);
        int corr = 0;
        if (v > k.g_fine_trigger)
        {
            int i = (v > k.g_coarse_trigger)  ?  k.g_coarse_step  :  k.g_fine_step;
            if (baud_phase < 0.0f)
                i = -i;
            corr = i;
            total_timing += i;
        }
        return corr;
    }

    // src/v29rx.c:788-864, src/v17rx.c:1136-1208 (IAXMODEM_STUFF is defined in both files)
    template <class K> SB_HD int signal_detect(const K &k, short amp)
    {
        const short x = (short) (amp >> 1);
        short diff = (short) (x - (short) last_sample);
        last_sample = x;
        power += (((int) diff*(int) diff - power) >> 4);                // power_meter_update, shift 4
        const int pw = power;
        diff = (short) abs((int) diff);
        if (10*(int) diff < high_sample)
        {
            if (++low_samples > 120)
            {
                power = 0;
                high_sample = 0;
                low_samples = 0;
            }
        }
        else
        {
            low_samples = 0;
            if ((int) diff > high_sample)
                high_sample = diff;
        }
        if (signal_present > 0)
        {
            if (drop_pending  ||  pw < off_power)
            {
                if (--signal_present <= 0)
                {
                    self().restart_after_carrier_down(k);
                    report_status(SIG_STATUS_CARRIER_DOWN);
                    return 0;
                }
                drop_pending = 1;
            }
        }
        else
        {
            if (pw < on_power)
                return 0;
            signal_present = 1;
            drop_pending = 0;
            report_status(SIG_STATUS_CARRIER_UP);
        }
        return pw;
    }

    // src/math_fixed.c:158-169
    SB_HD int fixed_sqrt32(const CoreConsts &k, unsigned int x)
    {
        if (x == 0)
            return 0;
        const int shift = 30 - ((31 - clz32(x)) & ~1);
        x <<= shift;
        return (int) sqrt_tab[((x >> 24) & 0xFF) - 64] >> (shift >> 1);
    }

    // xxx_rx()'s per-sample body (src/v29rx.c:885-960, src/v17rx.c:1231-1308) is split in three so that
    // the 32 channels of a warp can be kept in step on *symbol* time rather than sample time (see
    // modem_rx_kernel):
    //   front(): everything up to and including the real FIR and the Godard filters; tells whether this
    //            sample is a T/2 instant (eq_put_step <= 0);
    //   half():  the T/2 work: AGC, imaginary FIR, down-mix, equalizer buffer insert; tells whether a
    //            whole baud is now complete;
    //   baud():  timing correction, equalizer, training state machine / slicer, qam report.
    // The carrier NCO advance that ends the reference's loop body is done by whichever part ends the sample.
    int h_pw;
    float h_sre;
    float h_vim;

    template <class K> SB_HD bool front(const K &k, const float *s_rrc_re, const float *s_rrc_im, short amp)
    {
        const float xv = (float) amp;
        rrc[rrc_step*LS] = xv;
        if (LPC == 1)
            rrc[(rrc_step + SBM_FILTER_STEPS)*LS] = xv;
        if (++rrc_step >= SBM_FILTER_STEPS)
            rrc_step = 0;
        const int pw = signal_detect(k, amp);
        if (pw == 0)
            return false;
        if (training_stage == D::STAGE_PARKED)
            return false;
        eq_put_step -= COEFF_SETS;
        int step = -eq_put_step;
        if (step < 0)
            step += COEFF_SETS;
        if (step < 0)
            step = 0;
        else if (step > COEFF_SETS - 1)
            step = COEFF_SETS - 1;
        // Both FIRs at once: the imaginary one is only needed at T/2 instants, but computed together the
        // four accumulation chains overlap and the samples are fetched once.
        float v;
        rrc_dot2(s_rrc_re + step*SBM_RRC_ROW, s_rrc_im + step*SBM_RRC_ROW, v, h_vim);
        const float sre = fmul(v, agc_scaling);
        // godard_ted_rx, src/godard.c:144-161
        {
            float t = fadd(fadd(fmul(lbe0, k.g_low[0]), fmul(lbe1, k.g_low[1])), sre);
            lbe1 = lbe0;
            lbe0 = t;
            t = fadd(fadd(fmul(hbe0, k.g_high[0]), fmul(hbe1, k.g_high[1])), sre);
            hbe1 = hbe0;
            hbe0 = t;
        }
        if (eq_put_step <= 0)
        {
            h_pw = pw;
            h_sre = sre;
            return true;
        }
        carrier_phase += (unsigned int) phase_rate;
        return false;
    }

    template <class K> SB_HD bool half(const K &k)
    {
        if (agc_scaling_save == 0.0f)
        {
            int root_power = fixed_sqrt32(k, (unsigned int) h_pw);
            if (root_power == 0)
                root_power = 1;
            agc_scaling = fdiv(k.agc_target, (float) root_power);
        }
        const float sim = fmul(h_vim, agc_scaling);
        const float zr = sine[(carrier_phase + (1u << 30)) >> 21];       // dds_lookup_complexf, src/dds_float.c:2177
        const float zi = sine[carrier_phase >> 21];
        const float zzre = fsub(fmul(h_sre, zr), fmul(sim, zi));
        const float zzim = fsub(fmul(-h_sre, zi), fmul(sim, zr));
        eq_put_step += COEFF_SETS*10/(3*2);
        // process_half_baud, first part (src/v29rx.c:516-525, src/v17rx.c:638-647)
        eq_buf[eq_step*LS] = make_float2(zzre, zzim);
        if (LPC == 1)
            eq_buf[(eq_step + EQ_LEN)*LS] = make_float2(zzre, zzim);
        if (++eq_step >= EQ_LEN)
            eq_step = 0;
        if ((baud_half ^= 1))
        {
            carrier_phase += (unsigned int) phase_rate;
            return false;
        }
        return true;
    }

    template <class K> SB_HD void baud(const K &k)
    {
        self().process_baud(k);
        carrier_phase += (unsigned int) phase_rate;
    }

    // All samples of one call, lanes kept in step on symbol time: one trip of the outer loop takes every
    // receiving channel through one whole baud - two T/2 instants, each reached after one or two input
    // samples - so the expensive parts (imaginary FIR, equalizer, training/slicer) run with all lanes
    // converged even though the channels' symbol clocks sit at different sample phases.  Channels without
    // carrier (or parked) simply consume up to four samples per trip.  Each channel still sees its own
    // samples in order, which is all the reference's per-channel semantics require.
    // Input staging.  Each lane consumes its own row at its own pace (3 or 4 samples per baud), and a
    // warp has one scoreboard, so a per-lane register prefetch would make every lane wait for the most
    // recent load of any lane.  Instead each lane owns a ring of SBM_IN_RING samples in shared memory that
    // it tops up with 16-byte cp.async copies well ahead of use; the copies never pass through registers
    // and the only wait is cp.async.wait_group for groups issued several bauds earlier.
    short *in_ring;

#if defined(__CUDA_ARCH__)
    static __device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, int src_bytes)
    {
        const unsigned int dst = (unsigned int) __cvta_generic_to_shared(smem_dst);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" :: "r"(dst), "l"(gmem_src), "r"(src_bytes) : "memory");
    }
    static __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
    template <int N> static __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory"); }
#endif

    bool feed_staged;
    int feed_fill;
    const int16_t *feed_row;
    int feed_n;

    SB_HD void feed_open(const int16_t *row, int n)
    {
        feed_row = row;
        feed_n = n;
        feed_fill = 0;
        feed_staged = false;
#if defined(__CUDA_ARCH__)
        feed_staged = (((size_t) row) & 15) == 0;
        if (feed_staged)
        {
            for (int g = 0;  g < (SBM_IN_RING - 8)/8;  g++)
            {
                if (feed_fill < n)
                {
                    cp_async16(in_ring + (feed_fill & (SBM_IN_RING - 1)), row + feed_fill, (n - feed_fill >= 8)  ?  16  :  2*(n - feed_fill));
                    feed_fill += 8;
                }
            }
            cp_async_commit();
            cp_async_wait<0>();
        }
#endif
    }

    // The four samples at pos .. pos + 3, packed (zeros beyond the end of the row); tops the ring up by eight
    SB_HD unsigned long long feed_peek4(int pos)
    {
        unsigned long long cur;
#if defined(__CUDA_ARCH__)
        if (feed_staged)
        {
            if (feed_fill < feed_n  &&  feed_fill - pos <= SBM_IN_RING - 8)
            {
                cp_async16(in_ring + (feed_fill & (SBM_IN_RING - 1)), feed_row + feed_fill, (feed_n - feed_fill >= 8)  ?  16  :  2*(feed_n - feed_fill));
                feed_fill += 8;
            }
            cp_async_commit();
            cp_async_wait<2>();
            const unsigned int a0 = (unsigned short) in_ring[pos & (SBM_IN_RING - 1)];
            const unsigned int a1 = (unsigned short) in_ring[(pos + 1) & (SBM_IN_RING - 1)];
            const unsigned int a2 = (unsigned short) in_ring[(pos + 2) & (SBM_IN_RING - 1)];
            const unsigned int a3 = (unsigned short) in_ring[(pos + 3) & (SBM_IN_RING - 1)];
            cur = (unsigned long long) (a0 | (a1 << 16)) | ((unsigned long long) (a2 | (a3 << 16)) << 32);
        }
        else
#endif
        {
            cur = 0;
            for (int i = 0;  i < 4;  i++)
            {
                if (pos + i < feed_n)
                    cur |= (unsigned long long) (unsigned short) ldg(feed_row + pos + i) << (16*i);
            }
        }
        return cur;
    }

    // The first part of front(): the sample into the ring
    SB_HD void front_insert(short amp)
    {
        rrc[rrc_step*LS] = (float) amp;
        if (++rrc_step >= SBM_FILTER_STEPS)
            rrc_step = 0;
    }

    // The coefficient set front() will use for the next sample (eq_put_step after its decrement)
    SB_HD int front_step() const
    {
        int step = -(eq_put_step - COEFF_SETS);
        if (step < 0)
            step += COEFF_SETS;
        if (step < 0)
            step = 0;
        else if (step > COEFF_SETS - 1)
            step = COEFF_SETS - 1;
        return step;
    }

    // The rest of front(), given the FIR pair of the sample just inserted
    template <class K> SB_HD bool front_rest(const K &k, short amp, float v, float vim)
    {
        const int pw = signal_detect(k, amp);
        if (pw == 0)
            return false;
        if (training_stage == D::STAGE_PARKED)
            return false;
        eq_put_step -= COEFF_SETS;
        h_vim = vim;
        const float sre = fmul(v, agc_scaling);
        {
            float t = fadd(fadd(fmul(lbe0, k.g_low[0]), fmul(lbe1, k.g_low[1])), sre);
            lbe1 = lbe0;
            lbe0 = t;
            t = fadd(fadd(fmul(hbe0, k.g_high[0]), fmul(hbe1, k.g_high[1])), sre);
            hbe1 = hbe0;
            hbe0 = t;
        }
        if (eq_put_step <= 0)
        {
            h_pw = pw;
            h_sre = sre;
            return true;
        }
        carrier_phase += (unsigned int) phase_rate;
        return false;
    }

#if defined(__CUDA_ARCH__)
    // The deferred LMS updates of this baud, the whole warp on one receiver at a time: lane i takes tap i (lane 0 also
    // tap 32).  Same arithmetic per tap as the one-lane form; the taps are independent of each other.
    SB_HD void warp_lms()
    {
        const int lane = threadIdx.x & 31;
        unsigned int need = __ballot_sync(0xFFFFFFFFu, lms_pending != 0  &&  sub == 0);
        while (need)
        {
            const int src = __ffs((int) need) - 1;
            need &= need - 1;
            const float ere = __shfl_sync(0xFFFFFFFFu, lms_ere, src);
            const float eim = __shfl_sync(0xFFFFFFFFu, lms_eim, src);
            const int es = __shfl_sync(0xFFFFFFFFu, eq_step, src);
            const int shift = src/LPC - lane/LPC;               // that receiver's column relative to mine
#pragma unroll
            for (int rep = 0;  rep < 2;  rep++)
            {
                const int i = lane + 32*rep;
                if (i < EQ_LEN)
                {
                    int p = es + i;
                    if (p >= EQ_LEN)
                        p -= EQ_LEN;
                    const float2 x = eq_buf[p*LS + shift];
                    const float2 y = eq_coeff[i*LS + shift];
                    const float yr = fadd(fmul(y.x, 0.9999f), fadd(fmul(x.y, eim), fmul(x.x, ere)));
                    const float yi = fadd(fmul(y.y, 0.9999f), fsub(fmul(x.x, eim), fmul(x.y, ere)));
                    eq_coeff[i*LS + shift] = make_float2(yr, yi);
                    eq_coef_re[i*LS + shift] = make_float2(yr, -yi);
                    eq_coef_im[i*LS + shift] = make_float2(yi, yr);
                }
            }
        }
        lms_pending = 0;
        __syncwarp();
        if (__any_sync(0xFFFFFFFFu, save_pending != 0))
        {
            if (save_pending)
                equalizer_save_now();
            save_pending = 0;
        }
    }
#endif

    // Four lanes per receiver: the same walk, but the warp stays converged around the three cooperative pieces - the
    // FIR pair is evaluated by every lane in every round (for a receiver that takes no sample in a round it is
    // evaluated on whatever the ring holds and discarded: the instruction stream costs the same), the equalizer
    // once per trip, the LMS updates after it.
    template <class K> SB_HD void run_group(const K &k, const float *s_rrc_re, const float *s_rrc_im, const int16_t *row, int n)
    {
#if defined(__CUDA_ARCH__)
        int pos = 0;
        feed_open(row, n);
        lms_pending = 0;
        save_pending = 0;
#pragma unroll 1
        while (__any_sync(0xFFFFFFFFu, pos < n))
        {
            unsigned long long cur = feed_peek4(pos);
            bool whole = false;
            // (rolled: four copies of the sample path do not fit the instruction cache - measured, profiles/r02_ncu_v29_*)
#pragma unroll 1
            for (int h = 0;  h < 2;  h++)
            {
                const bool on = !(h == 0  &&  baud_half);
                bool due = false;
#pragma unroll 1
                for (int q = 0;  q < 2;  q++)
                {
                    const bool take = on  &&  pos < n  &&  !due;
                    const short amp = (short) (cur & 0xFFFFu);
                    if (take)
                    {
                        cur >>= 16;
                        front_insert(amp);
                        pos++;
                    }
                    const int step = front_step();
                    float v;
                    float vim;
                    rrc_dot2(s_rrc_re + step*SBM_RRC_ROW, s_rrc_im + step*SBM_RRC_ROW, v, vim);
                    if (take)
                        due = front_rest(k, amp, v, vim);
                }
                if (due)
                    whole = half(k);
            }
            // the equalizer on the buffer as it stands (for a receiver that completed a baud: with both T/2 samples in)
            equalizer_get(h_zre, h_zim);
            if (whole)
                baud(k);
            warp_lms();
        }
#endif
    }

    template <class K> SB_HD void run(const K &k, const float *s_rrc_re, const float *s_rrc_im, const int16_t *row, int n)
    {
        if (LPC > 1)
        {
            run_group(k, s_rrc_re, s_rrc_im, row, n);
            return;
        }
        int pos = 0;
        feed_open(row, n);
#pragma unroll 1
        while (pos < n)
        {
            // The (up to) four samples of this trip, packed
            unsigned long long cur = feed_peek4(pos);
#pragma unroll 1
            for (int h = 0;  h < 2;  h++)
            {
                // A channel that enters the trip half-way through a baud sits out the first slot, so that
                // every channel completes its baud in the second slot (and is baud-aligned from then on).
                if (h == 0  &&  baud_half)
                    continue;
                bool due = false;
#pragma unroll 1
                for (int q = 0;  q < 2;  q++)
                {
                    if (pos < n  &&  !due)
                    {
                        const short amp = (short) (cur & 0xFFFFu);
                        cur >>= 16;
                        due = front(k, s_rrc_re, s_rrc_im, amp);
                        pos++;
                    }
                }
                if (due)
                {
                    if (half(k))
                        baud(k);
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// kernels

#if defined(__CUDACC__)

template <class RX>
struct KernelArgs
{
    ModemArgs a;
    typename RX::Consts k;
};

// Shared memory: [rrc_re | rrc_im | sine | sqrt | RX tables | per warp: lane-interleaved per-receiver arrays, input rings]
template <class RX> __host__ __device__ constexpr int modem_table_words()
{
    return 2*RX::SETS*SBM_RRC_ROW + SBM_RRC_SKEW + SBM_SINE_WORDS + SBM_SQRT_WORDS + RX::TABLE_WORDS;
}

template <class RX> __host__ __device__ constexpr int modem_warp_words()
{
    return (RX::LANE_WORDS + RX::IN_RING_WORDS)*RX::LS;
}

template <class RX> __host__ __device__ constexpr int modem_smem_words(int warps = 1)
{
    return modem_table_words<RX>() + warps*modem_warp_words<RX>();
}

// The CTA's tables (all threads), then this thread's receiver bound to its warp's area
template <class RX>
__device__ __forceinline__ void modem_bind(RX &r, const KernelArgs<RX> &ka, float *smem, int lane, int c,
                                           const float *&s_rrc_re, const float *&s_rrc_im)
{
    float *w_rrc_re = smem;
    float *w_rrc_im = smem + RX::SETS*SBM_RRC_ROW + SBM_RRC_SKEW;
    float *w_sine = smem + 2*RX::SETS*SBM_RRC_ROW + SBM_RRC_SKEW;
    unsigned int *w_sqrt = (unsigned int *) (w_sine + SBM_SINE_WORDS);
    float *tables = w_sine + SBM_SINE_WORDS + SBM_SQRT_WORDS;
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    for (int i = tid;  i < SBM_SINE_WORDS;  i += nthr)
        w_sine[i] = ka.k.sine[i];
    for (int i = tid;  i < 97;  i += nthr)
        w_sqrt[i] = (unsigned int) ka.k.sqrt_tab[2*i] | ((2*i + 1 < 193)  ?  ((unsigned int) ka.k.sqrt_tab[2*i + 1] << 16)  :  0u);
    for (int i = tid;  i < RX::SETS*SBM_RRC_ROW;  i += nthr)
    {
        const int row = i/SBM_RRC_ROW;
        const int tap = i - row*SBM_RRC_ROW;
        w_rrc_re[i] = (tap < SBM_FILTER_STEPS)  ?  ka.k.rrc_re[row*SBM_FILTER_STEPS + tap]  :  0.0f;
        w_rrc_im[i] = (tap < SBM_FILTER_STEPS)  ?  ka.k.rrc_im[row*SBM_FILTER_STEPS + tap]  :  0.0f;
    }
    RX::fill_tables(tables, ka.k, tid, nthr);
    __syncthreads();
    float *lane_block = smem + modem_table_words<RX>() + (tid >> 5)*modem_warp_words<RX>();
    s_rrc_re = w_rrc_re;
    s_rrc_im = w_rrc_im;
    r.c = c;
    r.channels = ka.a.channels;
    r.fstate = ka.a.fstate;
    r.sine = w_sine;
    r.sqrt_tab = (const unsigned short *) w_sqrt;
    r.in_ring = (short *) (lane_block + RX::LANE_WORDS*RX::LS) + (lane/RX::LANES)*SBM_IN_RING;
    r.bind(tables, lane_block, lane);
}

// A warp runs 32/RX::LANES receivers; WARPS warps per CTA share one copy of the tables.  Few channels exist
// (thousands): with one lane per receiver a CTA is one warp so that they spread over all SMs; with four lanes per
// receiver there are four times the warps and two of them share a CTA's tables.
template <class RX, int WARPS>
__global__ void __launch_bounds__(WARPS*32) modem_rx_kernel(const KernelArgs<RX> ka)
{
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x*WARPS + (threadIdx.x >> 5))*RX::LS + lane/RX::LANES;
    RX r;
    const float *s_rrc_re;
    const float *s_rrc_im;
    // Lanes beyond the last channel: with one lane per receiver they leave; with several lanes per receiver the warp
    // has to stay whole for its shuffles, so they shadow the last channel (same input, same state, same arithmetic)
    // and write nothing.
    const bool live = (c < ka.a.channels);
    const int cc = (live)  ?  c  :  (ka.a.channels - 1);
    modem_bind(r, ka, smem, lane, cc, s_rrc_re, s_rrc_im);
    if (!live  &&  RX::LANES == 1)
        return;
    r.zero_pads();
    StateLoader ld = {ka.a.fstate, ka.a.istate, (size_t) ka.a.channels, (size_t) cc};
    r.visit(ld);
    r.mirror_rings();
    r.bits = NULL;
    r.bits_cap = 0;
    r.nbits = 0;
    r.words = ka.a.words + (size_t) cc*ka.a.words_cap;
    r.words_cap = (live)  ?  (int) ka.a.words_cap  :  0;
    r.nwords = 0;
    r.bit_acc = 0;
    r.bit_fill = 0;
    r.status = ka.a.status + (size_t) cc*ka.a.status_cap*2;
    r.status_cap = (live)  ?  (int) ka.a.status_cap  :  0;
    r.nstatus = 0;
    r.syms = (ka.a.syms  &&  live)  ?  (ka.a.syms + (size_t) cc*ka.a.sym_cap)  :  NULL;
    r.sym_cap = (int) ka.a.sym_cap;
    r.nsyms = 0;
    if (ka.a.append)
    {
        r.nbits = ka.a.nbits[cc];
        r.nstatus = ka.a.nstatus[cc];
        const int data_bits = r.nbits - r.nstatus;
        r.nwords = data_bits >> 5;
        r.bit_fill = data_bits & 31;
        r.bit_acc = (r.bit_fill != 0  &&  r.nwords < (int) ka.a.words_cap)  ?  r.words[r.nwords]  :  0u;
        if (ka.a.nsyms)
            r.nsyms = ka.a.nsyms[cc];
    }
    r.group_sync();
    r.run(ka.k, s_rrc_re, s_rrc_im, ka.a.amp + (long long) cc*ka.a.stride, ka.a.n);
    r.out_flush();
    r.group_sync();
    if (r.sub == 0  &&  live)
    {
        StateStorer st = {ka.a.fstate, ka.a.istate, (size_t) ka.a.channels, (size_t) c};
        r.visit(st);
        ka.a.nbits[c] = r.nbits;
        ka.a.nstatus[c] = r.nstatus;
        if (ka.a.nsyms)
            ka.a.nsyms[c] = r.nsyms;
    }
}

// xxx_rx_init() (mode < 0: all state zeroed first, then init with `mode` = -1 - restart argument) or
// xxx_rx_restart() (mode >= 0) for channels [first, first + count).  `mode` is the receiver's own restart
// argument (V.29: old_train; V.17: short_train).
template <class RX>
__global__ void __launch_bounds__(32) modem_init_kernel(const KernelArgs<RX> ka, int first, int count, int bit_rate, int mode,
                                                        int on_power, int off_power)
{
    extern __shared__ float smem[];
    const int lane = threadIdx.x;
    const int idx = blockIdx.x*32 + lane;
    const int c = first + idx;
    RX r;
    const float *s_rrc_re;
    const float *s_rrc_im;
    modem_bind(r, ka, smem, lane, (idx < count)  ?  c  :  first, s_rrc_re, s_rrc_im);
    if (idx >= count)
        return;
    r.zero_pads();
    r.bits = NULL;
    r.bits_cap = 0;
    r.nbits = 0;
    r.words = NULL;
    r.words_cap = 0;
    r.nwords = 0;
    r.bit_acc = 0;
    r.bit_fill = 0;
    r.status = NULL;
    r.status_cap = 0;
    r.nstatus = 0;
    r.syms = NULL;
    r.sym_cap = 0;
    r.nsyms = 0;
    if (mode < 0)
    {
        r.init(ka.k, bit_rate, on_power, off_power);
    }
    else
    {
        StateLoader ld = {ka.a.fstate, ka.a.istate, (size_t) ka.a.channels, (size_t) c};
        r.visit(ld);
        r.mirror_rings();
        r.restart(ka.k, bit_rate, mode);
    }
    StateStorer st = {ka.a.fstate, ka.a.istate, (size_t) ka.a.channels, (size_t) c};
    r.visit(st);
}

#endif  // __CUDACC__

}  // namespace sbm
