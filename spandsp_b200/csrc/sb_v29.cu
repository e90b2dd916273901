// sb_v29.cu - V.29 receiver banks: RRC band-pass FIR, Godard timing, T/2 complex adaptive equalizer,
// slicer, carrier loop, descrambler.  One CUDA thread runs one channel; a bank holds N channels.
//
// Reference: src/v29rx.c (whole file), src/godard.c:144-220, src/power_meter.c:65-69,
// src/math_fixed.c:158-169, src/dds_float.c:2135-2180, src/spandsp/arctan2.h:47-80,
// src/vector_float.c:890-939 (scalar dot products), src/complex_vector_float.c:137-219.
//
// Arithmetic contract: every float operation is an explicitly rounded __f*_rn in the reference's
// operand order (the pinned oracle is the strict build with sequential dot products), integers
// follow C semantics of the reference on x86-64 (arithmetic right shifts, int16 wrap-around,
// cvttss2si for float->int32).  The only libm calls in the loop are cosf/sinf at the one-off
// equalizer "spin" (src/v29rx.c:618-623); the device versions may differ in the last ulp there.
//
// The constant tables (48x27 RRC coefficient sets x2, 2048-entry sine table, sqrt table, Godard
// descriptor) are computed at context creation by our own generators below, which redo the
// reference's build-time generator programs (src/make_modem_filter.c:155-271, src/filter_tools.c:60-190,
// src/make_modem_godard_descriptor.c:60-80, src/make_math_fixed_tables.c:59-72) including their
// print-to-decimal / parse-as-float step, so that the float tables are bit-identical
// (tests/test_v29_tables.py checks them against tables dumped from the compiled reference).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "sb_engine.h"
#pragma GCC visibility push(default)
#include "../../include/spandsp_b200_v29.h"
#pragma GCC visibility pop

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

// ------------------------------------------------------------------------------------------
// constant tables: host generators

#define V29_COEFF_SETS      48
#define V29_FILTER_STEPS    27
#define V29_EQ_LEN          33
#define V29_EQ_PRE_LEN      16

static float decimal_roundtrip(double v, int decimals)
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

static void dit_transform(cplx *data, cplx *temp, int n, const std::vector<cplx> &circle, int full)
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

// Root raised cosine prototype by frequency sampling (src/filter_tools.c:126-190), then the
// polyphase band-pass sets (src/make_modem_filter.c:155-271; V.29: 48 sets x 27 taps, 1700 Hz,
// 2400 baud, excess bandwidth 0.5, :401-413).
static void make_v29_rrc(std::vector<float> &re, std::vector<float> &im)
{
    const int SEQ_LEN = 8192;
    const double GEN_PI = 3.1415926535;
    const int coeff_sets = V29_COEFF_SETS;
    const int per_filter = V29_FILTER_STEPS;
    const int total = coeff_sets*per_filter + 1;
    const double alpha = 2400.0/(2.0*(double) (coeff_sets*8000));
    const double beta = 0.5;
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
    std::vector<double> coeffs(total);
    const int h = (total - 1)/2;
    for (int i = 0;  i < total;  i++)
        coeffs[i] = vec[(SEQ_LEN - h + i) % SEQ_LEN].re/(double) SEQ_LEN;
    double gain = 0.0;
    for (int i = coeff_sets/2;  i < total;  i += coeff_sets)
        gain += coeffs[i];
    for (int i = 0;  i < total;  i++)
        coeffs[i] /= gain;
    double carrier = 1700.0;
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

// src/dds_float.c:51-2101: sin(2*pi*i/2048) as 8-decimal literals.
static void make_sine_table(std::vector<float> &t)
{
    t.resize(2048);
    for (int i = 0;  i < 2048;  i++)
        t[i] = decimal_roundtrip(sin(2.0*M_PI*(double) i/2048.0), 8);
}

// src/make_math_fixed_tables.c:59-72
static void make_sqrt_table(std::vector<unsigned short> &t)
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

// src/make_modem_godard_descriptor.c:60-80 with the V.29 arguments of src/Makefile.am:556-560:
// 1700.0 2400.0 0.99 1000.0 30.0 5 1; floats are printed with 6 decimals.
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

static void make_godard(godard_desc_t &g)
{
    const double alpha = 0.99;
    const double low_edge = 2.0*M_PI*(1700.0 - 2400.0/2.0)/8000.0;
    const double high_edge = 2.0*M_PI*(1700.0 + 2400.0/2.0)/8000.0;
    g.low[0] = decimal_roundtrip(2.0*alpha*cos(low_edge), 6);
    g.high[0] = decimal_roundtrip(2.0*alpha*cos(high_edge), 6);
    g.low[1] = g.high[1] = decimal_roundtrip(-alpha*alpha, 6);
    g.low[2] = decimal_roundtrip(-alpha*sin(low_edge), 6);
    g.high[2] = decimal_roundtrip(-alpha*sin(high_edge), 6);
    g.mixed3 = decimal_roundtrip(-alpha*alpha*(sin(high_edge)*cos(low_edge) - sin(low_edge)*cos(high_edge)), 6);
    g.coarse_trigger = decimal_roundtrip(1000.0, 6);
    g.fine_trigger = decimal_roundtrip(30.0, 6);
    g.coarse_step = 5;
    g.fine_step = 1;
}

// src/power_meter.c:86-96
static int host_power_meter_level_dbm0(float level)
{
    float l;

    level -= (3.14f + 3.02f);
    if (level > 0.0)
        level = 0.0;
    l = powf(10.0f, level/10.0f)*(32767.0f*32767.0f);
    return (int) l;
}

extern "C" int span_b200_v29_tables(float *rrc_re, float *rrc_im, float *sine, uint16_t *sqrt_tab, float *godard, int32_t *ints)
{
    std::vector<float> re;
    std::vector<float> im;
    std::vector<float> st;
    std::vector<unsigned short> sq;
    godard_desc_t g;
    make_v29_rrc(re, im);
    make_sine_table(st);
    make_sqrt_table(sq);
    make_godard(g);
    memcpy(rrc_re, re.data(), sizeof(float)*re.size());
    memcpy(rrc_im, im.data(), sizeof(float)*im.size());
    memcpy(sine, st.data(), sizeof(float)*st.size());
    memcpy(sqrt_tab, sq.data(), sizeof(unsigned short)*sq.size());
    godard[0] = g.low[0];
    godard[1] = g.low[1];
    godard[2] = g.low[2];
    godard[3] = g.high[0];
    godard[4] = g.high[1];
    godard[5] = g.high[2];
    godard[6] = g.mixed3;
    godard[7] = g.coarse_trigger;
    godard[8] = g.fine_trigger;
    ints[0] = g.coarse_step;
    ints[1] = g.fine_step;
    ints[2] = (int32_t) (1700.0f*65536.0f*65536.0f/8000);                   // DDS_PHASE_RATE, src/spandsp/dds.h:31
    ints[3] = (int32_t) ((1700.0f - 20.0f)*65536.0f*65536.0f/8000);
    ints[4] = (int32_t) ((1700.0f + 20.0f)*65536.0f*65536.0f/8000);
    ints[5] = (int32_t) ((uint32_t) (45.0f*65536.0f*65536.0f/360.0f));      // DDS_PHASE, src/spandsp/dds.h:32
    ints[6] = (int32_t) ((uint32_t) ((360.0f + -45.0f)*65536.0f*65536.0f/360.0f));
    ints[7] = host_power_meter_level_dbm0(-28.5f + 2.5f);
    ints[8] = host_power_meter_level_dbm0(-28.5f - 2.5f);
    return 0;
}

// ------------------------------------------------------------------------------------------
// device side

namespace v29 {

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// (int32_t) of a float the way x86-64's cvttss2si does it: out-of-range and NaN give INT_MIN.
__device__ __forceinline__ int f2i(float f)
{
    if (!(f > -2147483904.0f  &&  f < 2147483648.0f))
        return (int) 0x80000000;
    return __float2int_rz(f);
}

// cosf()/sinf() as the host C library computes them.  The reference calls libm's cosf/sinf once per
// training (src/v29rx.c:618-623); the values feed the adaptive loops, whose discrete timing decisions
// amplify a 1-ulp difference into visible (1e-3) excursions of the soft symbols, so they have to be
// reproduced exactly.  Third-party arithmetic: GNU libc 2.39 (the image's libm.so.6), sysdeps/ieee754/
// flt-32/s_cosf.c, s_sinf.c, s_sincosf.h - the "sincosf" of ARM's optimized routines: reduce by pi/2 in
// double with a 2^24-prescaled 2/pi, then an odd/even minimax polynomial in double, rounded once to
// float.  Constants are the published ones (they can be read back from __sincosf_table in libm.so.6).
// Arguments here are phases in [0, 2*pi), so only the two fast paths (|x| < pi/4, |x| < 120) are needed.
// tools/check_host_sincosf.py checks this restatement against the live libm on 1e6 arguments.
__device__ __forceinline__ double sincosf_poly(double x, double x2, bool cos_table_negated, int n)
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
        const double x3 = __dmul_rn(x, x2);
        const double t1 = __dadd_rn(s2, __dmul_rn(x2, s3));
        const double x7 = __dmul_rn(x3, x2);
        const double sv = __dadd_rn(x, __dmul_rn(x3, s1));
        return __dadd_rn(sv, __dmul_rn(x7, t1));
    }
    const double x4 = __dmul_rn(x2, x2);
    const double t2 = __dadd_rn(c3, __dmul_rn(x2, c4));
    const double t1 = __dadd_rn(c0, __dmul_rn(x2, c1));
    const double x6 = __dmul_rn(x4, x2);
    const double cv = __dadd_rn(t1, __dmul_rn(x4, c2));
    return __dadd_rn(cv, __dmul_rn(x6, t2));
}

__device__ __forceinline__ unsigned int abstop12(float f)
{
    return (__float_as_uint(f) >> 20) & 0x7FFu;
}

// is_cos: 1 for cosf, 0 for sinf
__device__ float host_sincosf(float y, int is_cos)
{
    double x = (double) y;
    if (abstop12(y) < abstop12(0x1.921FB6p-1f))
    {
        if (abstop12(y) < abstop12(0x1p-12f))
            return (is_cos)  ?  1.0f  :  y;
        return (float) sincosf_poly(x, __dmul_rn(x, x), false, is_cos);
    }
    const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = __dsub_rn(x, __dmul_rn((double) n, 0x1.921FB54442D18p0));
    const double sgn = ((n & 3) == 1  ||  (n & 3) == 2)  ?  -1.0  :  1.0;
    return (float) sincosf_poly(__dmul_rn(x, sgn), __dmul_rn(x, x), (n & 2) != 0, (is_cos)  ?  (n ^ 1)  :  n);
}

__device__ __forceinline__ float host_cosf(float y) { return host_sincosf(y, 1); }
__device__ __forceinline__ float host_sinf(float y) { return host_sincosf(y, 0); }

struct Consts
{
    const float *rrc_re;                // [48][27]
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
    int rate_nominal;                   // DDS_PHASE_RATE(1700)
    int rate_low;                       // DDS_PHASE_RATE(1680)
    int rate_high;                      // DDS_PHASE_RATE(1720)
    int phase_p45;
    int phase_m45;
    float agc_initial;                  // (1.25f/1.0f)/735.0f
    float eq_delta;                     // 0.21f/33
};

__constant__ unsigned char c_space_map_9600[20][20];    // src/v29rx.c:119-143
__constant__ float c_constellation[16][2];              // src/v29tx_constellation_maps.h:57-79
__constant__ unsigned char c_phase_steps_9600[8];       // src/v29rx.c:404-407
__constant__ unsigned char c_phase_steps_4800[4];       // src/v29rx.c:408-411
__constant__ int c_cdcd_pos[6];                         // src/v29rx.c:488-493

// Per-channel state in global memory, structure of arrays: field f of channel c at [f*C + c].
enum
{
    F_AGC = 0, F_AGC_SAVE, F_TRAINING_ERROR, F_TRACK_P, F_TRACK_I,
    F_LBE0, F_LBE1, F_HBE0, F_HBE1, F_DC0, F_DC1, F_BAUD_PHASE,
    F_EQ_COEFF,                                     // 66
    F_EQ_COEFF_SAVE = F_EQ_COEFF + 2*V29_EQ_LEN,    // 66
    F_EQ_BUF = F_EQ_COEFF_SAVE + 2*V29_EQ_LEN,      // 66
    F_RRC = F_EQ_BUF + 2*V29_EQ_LEN,                // 27
    F_COUNT = F_RRC + V29_FILTER_STEPS
};

enum
{
    I_BIT_RATE = 0, I_TRAINING_CD, I_OLD_TRAIN, I_RRC_STEP, I_SCRAMBLE, I_TRAIN_SCRAMBLE, I_STAGE, I_TRAIN_COUNT,
    I_LAST_SAMPLE, I_SIGNAL_PRESENT, I_DROP_PENDING, I_LOW_SAMPLES, I_HIGH_SAMPLE, I_CARRIER_PHASE, I_PHASE_RATE,
    I_PHASE_RATE_SAVE, I_POWER, I_ON_POWER, I_OFF_POWER, I_EQ_STEP, I_EQ_PUT_STEP, I_EQ_SKIP, I_BAUD_HALF,
    I_LAST_ANGLE0, I_LAST_ANGLE1, I_DIFF_ANGLES,    // 16
    I_CONSTELLATION = I_DIFF_ANGLES + 16, I_TOTAL_TIMING, I_COUNT
};

enum
{
    STAGE_NORMAL = 0, STAGE_SYMBOL_ACQUISITION, STAGE_LOG_PHASE, STAGE_WAIT_FOR_CDCD, STAGE_TRAIN_ON_CDCD,
    STAGE_TRAIN_ON_CDCD_AND_TEST, STAGE_TEST_ONES, STAGE_PARKED
};

#define SIG_STATUS_CARRIER_DOWN             (-1)
#define SIG_STATUS_CARRIER_UP               (-2)
#define SIG_STATUS_TRAINING_IN_PROGRESS     (-3)
#define SIG_STATUS_TRAINING_SUCCEEDED       (-4)
#define SIG_STATUS_TRAINING_FAILED          (-5)

struct Args
{
    const int16_t *amp;
    long long stride;
    int n;
    int channels;
    float *fstate;
    int *istate;
    signed char *bits;                  // [channel][bits_cap]: 0/1 data bits and negative status codes
    long long bits_cap;
    int *nbits;                         // [channel]
    span_b200_v29_symbol_t *syms;       // [channel][sym_cap], or NULL
    long long sym_cap;
    int *nsyms;
    Consts k;
};

// One receiver.  Scalars live in registers; the three per-channel arrays live in shared memory,
// lane-interleaved (element e of lane l at [e*32 + l]) so that any per-lane index is conflict-free.
struct Rx
{
    // float state
    float agc_scaling, agc_scaling_save, training_error, track_p, track_i;
    float lbe0, lbe1, hbe0, hbe1, dc0, dc1, baud_phase;
    // int state
    int bit_rate, training_cd, rrc_step, training_stage, training_count;
    unsigned int scramble_reg, training_scramble_reg, carrier_phase;
    int last_sample, signal_present, drop_pending, low_samples, high_sample;
    int phase_rate, phase_rate_save, power, on_power, off_power;
    int eq_step, eq_put_step, eq_skip, baud_half, constellation_state, total_timing;
    int last_angle0, last_angle1;
    int *diff_angles;       // [16], shared memory, lane-interleaved (dynamic indexing would force the
                            // whole receiver out of registers if it were a member array)
    // shared-memory arrays of this lane
    float *eq_coeff;        // [66]
    float *eq_buf;          // [66]
    float *rrc;             // [27]
    // outputs
    signed char *bits;
    int nbits;
    int bits_cap;
    span_b200_v29_symbol_t *syms;
    int nsyms;
    int sym_cap;
    // channel
    int c;
    int channels;
    float *fstate;

    __device__ __forceinline__ void out_bit(int v)
    {
        if (nbits < bits_cap)
            bits[nbits] = (signed char) v;
        nbits++;
    }

    // src/v29rx.c:171-178: without a status handler the status goes through put_bit
    __device__ __forceinline__ void report_status(int status)
    {
        out_bit(status);
    }

    // src/v29rx.c:214-258
    __device__ __forceinline__ void equalizer_reset(const Consts &k)
    {
        for (int i = 0;  i < 2*V29_EQ_LEN;  i++)
        {
            eq_coeff[i*32] = 0.0f;
            eq_buf[i*32] = 0.0f;
        }
        eq_coeff[(2*V29_EQ_PRE_LEN)*32] = 3.0f;
        eq_put_step = V29_COEFF_SETS*10/(3*2) - 1;
        eq_step = 0;
    }

    __device__ __forceinline__ void equalizer_restore(const Consts &k)
    {
        for (int i = 0;  i < 2*V29_EQ_LEN;  i++)
        {
            eq_coeff[i*32] = fstate[(size_t) (F_EQ_COEFF_SAVE + i)*channels + c];
            eq_buf[i*32] = 0.0f;
        }
        eq_put_step = V29_COEFF_SETS*10/(3*2) - 1;
        eq_step = 0;
    }

    __device__ __forceinline__ void equalizer_save()
    {
        for (int i = 0;  i < 2*V29_EQ_LEN;  i++)
            fstate[(size_t) (F_EQ_COEFF_SAVE + i)*channels + c] = eq_coeff[i*32];
    }

    // src/v29rx.c:1019-1097
    __device__ __forceinline__ void restart(const Consts &k, int rate, bool old_train)
    {
        training_cd = (rate == 9600)  ?  0  :  (rate == 7200)  ?  2  :  4;
        bit_rate = rate;
        for (int i = 0;  i < V29_FILTER_STEPS;  i++)
            rrc[i*32] = 0.0f;
        rrc_step = 0;
        scramble_reg = 0;
        training_scramble_reg = 0x2A;
        training_stage = STAGE_SYMBOL_ACQUISITION;
        training_count = 0;
        signal_present = 0;
        high_sample = 0;
        low_samples = 0;
        drop_pending = 0;
        for (int i = 0;  i < 16;  i++)
            diff_angles[i*32] = 0;
        carrier_phase = 0;
        power = 0;                                  // power_meter_init(&s->power, 4)
        constellation_state = 0;
        if (old_train)
        {
            phase_rate = phase_rate_save;
            equalizer_restore(k);
            agc_scaling = agc_scaling_save;
        }
        else
        {
            phase_rate = k.rate_nominal;
            equalizer_reset(k);
            agc_scaling_save = 0.0f;
            agc_scaling = k.agc_initial;
        }
        track_i = 8000.0f;
        track_p = 8000000.0f;
        last_sample = 0;
        eq_skip = 0;
        lbe0 = lbe1 = hbe0 = hbe1 = dc0 = dc1 = baud_phase = 0.0f;    // godard_ted_init
        total_timing = 0;
        baud_half = 0;
    }

    // src/spandsp/arctan2.h:47-80
    __device__ __forceinline__ int arctan2(float y, float x)
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

    // src/vector_float.c:932-939 with the scalar vec_dot_prodf (:890-900): two segments, each summed
    // in order from 0.0f, then added.
    __device__ __forceinline__ float rrc_dot(const float *coef)
    {
        float za = 0.0f;
        float zb = 0.0f;
        const int first = V29_FILTER_STEPS - rrc_step;      // taps in the first segment
        int j = rrc_step;
#pragma unroll 9
        for (int i = 0;  i < V29_FILTER_STEPS;  i++)
        {
            const float p = fmul(rrc[j*32], coef[i]);
            if (i < first)
                za = fadd(za, p);
            else
                zb = fadd(zb, p);
            if (++j >= V29_FILTER_STEPS)
                j = 0;
        }
        return fadd(za, zb);
    }

    // src/complex_vector_float.c:137-150,187-196
    __device__ __forceinline__ void equalizer_get(float &zre, float &zim)
    {
        float are = 0.0f, aim = 0.0f, bre = 0.0f, bim = 0.0f;
        const int first = V29_EQ_LEN - eq_step;
        int j = eq_step;
#pragma unroll 3
        for (int i = 0;  i < V29_EQ_LEN;  i++)
        {
            const float xr = eq_buf[(2*j)*32];
            const float xi = eq_buf[(2*j + 1)*32];
            const float yr = eq_coeff[(2*i)*32];
            const float yi = eq_coeff[(2*i + 1)*32];
            const float pr = fsub(fmul(xr, yr), fmul(xi, yi));
            const float pi = fadd(fmul(xr, yi), fmul(xi, yr));
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
            if (++j >= V29_EQ_LEN)
                j = 0;
        }
        zre = fadd(are, bre);
        zim = fadd(aim, bim);
    }

    // src/v29rx.c:281-290 + src/complex_vector_float.c:199-219
    __device__ __forceinline__ void tune_equalizer(const Consts &k, float zre, float zim, float tre, float tim)
    {
        const float ere = fmul(fsub(tre, zre), k.eq_delta);
        const float eim = fmul(fsub(tim, zim), k.eq_delta);
        int j = eq_step;
#pragma unroll 3
        for (int i = 0;  i < V29_EQ_LEN;  i++)
        {
            const float xr = eq_buf[(2*j)*32];
            const float xi = eq_buf[(2*j + 1)*32];
            const float yr = eq_coeff[(2*i)*32];
            const float yi = eq_coeff[(2*i + 1)*32];
            eq_coeff[(2*i)*32] = fadd(fmul(yr, 0.9999f), fadd(fmul(xi, eim), fmul(xr, ere)));
            eq_coeff[(2*i + 1)*32] = fadd(fmul(yi, 0.9999f), fsub(fmul(xr, eim), fmul(xi, ere)));
            if (++j >= V29_EQ_LEN)
                j = 0;
        }
    }

    // src/v29rx.c:297-331
    __device__ __forceinline__ void track_carrier(float zre, float zim, float tre, float tim)
    {
        const float error = fsub(fmul(zim, tre), fmul(zre, tim));
        phase_rate += f2i(fmul(track_i, error));
        carrier_phase += (unsigned int) f2i(fmul(track_p, error));
    }

    // src/v29rx.c:350-361
    __device__ __forceinline__ int scrambled_training_bit()
    {
        const int bit = training_scramble_reg & 1;
        training_scramble_reg >>= 1;
        if (bit ^ (int) (training_scramble_reg & 1))
            training_scramble_reg |= 0x40;
        return bit;
    }

    // src/v29rx.c:365-396
    __device__ __forceinline__ void put_bit(int bit)
    {
        bit &= 1;
        const int out = (bit ^ (int) (scramble_reg >> (18 - 1)) ^ (int) (scramble_reg >> (23 - 1))) & 1;
        scramble_reg = (scramble_reg << 1) | (unsigned int) bit;
        if (training_stage == STAGE_NORMAL)
            out_bit(out);
    }

    // src/v29rx.c:402-480
    __device__ __forceinline__ void decode_baud(const Consts &k, float zre, float zim)
    {
        int nearest;
        int raw_bits;

        if (bit_rate == 4800)
        {
            const int b1 = (zim > zre);
            const int b2 = (zim < -zre);
            nearest = ((b2 << 1) | (b1 ^ b2)) << 1;
            raw_bits = c_phase_steps_4800[((nearest - constellation_state) >> 1) & 3];
            put_bit(raw_bits);
            put_bit(raw_bits >> 1);
        }
        else
        {
            int re = f2i(fmul(fadd(zre, 5.0f), 2.0f));
            int im = f2i(fmul(fadd(zim, 5.0f), 2.0f));
            re = (re > 19)  ?  19  :  (re < 0)  ?  0  :  re;
            im = (im > 19)  ?  19  :  (im < 0)  ?  0  :  im;
            nearest = c_space_map_9600[re][im];
            if (bit_rate == 9600)
                put_bit(nearest >> 3);
            else
                nearest &= 7;
            raw_bits = c_phase_steps_9600[(nearest - constellation_state) & 7];
            for (int i = 0;  i < 3;  i++)
            {
                put_bit(raw_bits);
                raw_bits >>= 1;
            }
        }
        const float tre = c_constellation[nearest][0];
        const float tim = c_constellation[nearest][1];
        track_carrier(zre, zim, tre, tim);
        if (--eq_skip <= 0)
        {
            eq_skip = 10;
            tune_equalizer(k, zre, zim, tre, tim);
        }
        constellation_state = nearest;
    }

    // src/godard.c:165-220
    __device__ __forceinline__ int godard_per_baud(const Consts &k)
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

    __device__ __forceinline__ void park()
    {
        agc_scaling_save = 0.0f;
        training_stage = STAGE_PARKED;
        report_status(SIG_STATUS_TRAINING_FAILED);
    }

    // src/v29rx.c:526-785: the once-per-baud part of process_half_baud()
    __device__ __forceinline__ void process_baud(const Consts &k)
    {
        eq_put_step += godard_per_baud(k);
        float zre;
        float zim;
        equalizer_get(zre, zim);
        float tre = 0.0f;
        float tim = 0.0f;

        switch (training_stage)
        {
        case STAGE_NORMAL:
            decode_baud(k, zre, zim);
            tre = c_constellation[constellation_state][0];
            tim = c_constellation[constellation_state][1];
            break;
        case STAGE_SYMBOL_ACQUISITION:
            if (++training_count >= 60)
            {
                training_stage = STAGE_LOG_PHASE;
                for (int i = 0;  i < 16;  i++)
                    diff_angles[i*32] = 0;
                last_angle0 = arctan2(zim, zre);
                if (agc_scaling_save == 0.0f)
                    agc_scaling_save = agc_scaling;
            }
            break;
        case STAGE_LOG_PHASE:
            last_angle1 = arctan2(zim, zre);
            training_count = 1;
            training_stage = STAGE_WAIT_FOR_CDCD;
            break;
        case STAGE_WAIT_FOR_CDCD:
            {
                const int angle = arctan2(zim, zre);
                int i = training_count + 1;
                int ang = angle - ((i & 1)  ?  last_angle1  :  last_angle0);
                if (i & 1)
                    last_angle1 = angle;
                else
                    last_angle0 = angle;
                diff_angles[(i & 0xF)*32] = diff_angles[((i - 2) & 0xF)*32] + (ang >> 4);
                if ((ang > k.phase_p45  ||  ang < k.phase_m45)  &&  training_count >= 13)
                {
                    i = (training_count - 8) & ~1;
                    if (i > 1)
                    {
                        const int j = i & 0xF;
                        ang = (diff_angles[j*32] + diff_angles[(j | 0x1)*32])/(i - 1);
                        phase_rate += 3*16*(ang/20);
                    }
                    if (phase_rate < k.rate_low  ||  phase_rate > k.rate_high)
                    {
                        park();
                        break;
                    }
                    // dds_phase_to_radians (src/dds_float.c:2103-2106) of the angle as uint32
                    const float p = fdiv(fmul(fmul((float) (unsigned int) angle, 2.0f), 3.1415926f), fmul(65536.0f, 65536.0f));
                    const float cr = host_cosf(p);
                    const float ci = -host_sinf(p);
                    for (int q = 0;  q < V29_EQ_LEN;  q++)
                    {
                        const float xr = eq_buf[(2*q)*32];
                        const float xi = eq_buf[(2*q + 1)*32];
                        eq_buf[(2*q)*32] = fsub(fmul(xr, cr), fmul(xi, ci));
                        eq_buf[(2*q + 1)*32] = fadd(fmul(xr, ci), fmul(xi, cr));
                    }
                    carrier_phase += (unsigned int) angle;
                    const int bit = scrambled_training_bit();
                    constellation_state = c_cdcd_pos[training_cd + bit];
                    tre = c_constellation[constellation_state][0];
                    tim = c_constellation[constellation_state][1];
                    training_count = 1;
                    training_stage = STAGE_TRAIN_ON_CDCD;
                    report_status(SIG_STATUS_TRAINING_IN_PROGRESS);
                    break;
                }
                if (++training_count > 128)
                    park();
            }
            break;
        case STAGE_TRAIN_ON_CDCD:
            {
                const int bit = scrambled_training_bit();
                constellation_state = c_cdcd_pos[training_cd + bit];
                tre = c_constellation[constellation_state][0];
                tim = c_constellation[constellation_state][1];
                track_carrier(zre, zim, tre, tim);
                tune_equalizer(k, zre, zim, tre, tim);
                if (++training_count >= 384 - 48)
                {
                    training_stage = STAGE_TRAIN_ON_CDCD_AND_TEST;
                    training_error = 0.0f;
                    track_i = 200.0f;
                    track_p = 1000000.0f;
                }
            }
            break;
        case STAGE_TRAIN_ON_CDCD_AND_TEST:
            {
                const int bit = scrambled_training_bit();
                constellation_state = c_cdcd_pos[training_cd + bit];
                tre = c_constellation[constellation_state][0];
                tim = c_constellation[constellation_state][1];
                track_carrier(zre, zim, tre, tim);
                tune_equalizer(k, zre, zim, tre, tim);
                const float dr = fsub(zre, tre);
                const float di = fsub(zim, tim);
                training_error = fadd(training_error, fadd(fmul(dr, dr), fmul(di, di)));
                if (++training_count >= 384)
                {
                    if (training_error < fmul(48.0f, 2.0f))
                    {
                        training_error = 0.0f;
                        training_count = 0;
                        constellation_state = 0;
                        training_stage = STAGE_TEST_ONES;
                    }
                    else
                    {
                        park();
                    }
                }
            }
            break;
        case STAGE_TEST_ONES:
            {
                decode_baud(k, zre, zim);
                tre = c_constellation[constellation_state][0];
                tim = c_constellation[constellation_state][1];
                const float dr = fsub(zre, tre);
                const float di = fsub(zim, tim);
                training_error = fadd(training_error, fadd(fmul(dr, dr), fmul(di, di)));
                if (++training_count >= 48)
                {
                    if (training_error < fmul(48.0f, 1.0f))
                    {
                        report_status(SIG_STATUS_TRAINING_SUCCEEDED);
                        signal_present = 60;
                        training_stage = STAGE_NORMAL;
                        equalizer_save();
                        phase_rate_save = phase_rate;
                        agc_scaling_save = agc_scaling;
                    }
                    else
                    {
                        park();
                    }
                }
            }
            break;
        default:
            break;
        }
        if (syms)
        {
            if (nsyms < sym_cap)
            {
                span_b200_v29_symbol_t s;
                s.re = zre;
                s.im = zim;
                s.target_re = tre;
                s.target_im = tim;
                s.state = constellation_state;
                s.bit_pos = nbits;
                syms[nsyms] = s;
            }
            nsyms++;
        }
    }

    // src/v29rx.c:788-864 (IAXMODEM_STUFF is defined at src/v29rx.c:1)
    __device__ __forceinline__ int signal_detect(const Consts &k, short amp)
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
                    restart(k, bit_rate, false);
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
    __device__ __forceinline__ int fixed_sqrt32(const Consts &k, unsigned int x)
    {
        if (x == 0)
            return 0;
        const int shift = 30 - ((31 - __clz(x)) & ~1);
        x <<= shift;
        return (int) k.sqrt_tab[((x >> 24) & 0xFF) - 64] >> (shift >> 1);
    }

    // v29_rx()'s per-sample body (src/v29rx.c:885-960) is split in three so that the 32 channels of a
    // warp can be kept in step on *symbol* time rather than sample time (see v29_rx_kernel):
    //   front(): everything up to and including the real FIR and the Godard filters; tells whether this
    //            sample is a T/2 instant (eq_put_step <= 0);
    //   half():  the T/2 work: AGC, imaginary FIR, down-mix, equalizer buffer insert; tells whether a
    //            whole baud is now complete;
    //   baud():  timing correction, equalizer, training state machine / slicer, qam report.
    // The carrier NCO advance that ends the reference's loop body is done by whichever part ends the sample.
    int h_step;
    int h_pw;
    float h_sre;

    __device__ __forceinline__ bool front(const Consts &k, const float *s_rrc_re, short amp)
    {
        rrc[rrc_step*32] = (float) amp;
        if (++rrc_step >= V29_FILTER_STEPS)
            rrc_step = 0;
        const int pw = signal_detect(k, amp);
        if (pw == 0)
            return false;
        if (training_stage == STAGE_PARKED)
            return false;
        eq_put_step -= V29_COEFF_SETS;
        int step = -eq_put_step;
        if (step < 0)
            step += V29_COEFF_SETS;
        if (step < 0)
            step = 0;
        else if (step > V29_COEFF_SETS - 1)
            step = V29_COEFF_SETS - 1;
        const float v = rrc_dot(s_rrc_re + step*V29_FILTER_STEPS);
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
            h_step = step;
            h_pw = pw;
            h_sre = sre;
            return true;
        }
        carrier_phase += (unsigned int) phase_rate;
        return false;
    }

    __device__ __forceinline__ bool half(const Consts &k, const float *s_rrc_im)
    {
        if (agc_scaling_save == 0.0f)
        {
            int root_power = fixed_sqrt32(k, (unsigned int) h_pw);
            if (root_power == 0)
                root_power = 1;
            agc_scaling = fdiv(fdiv(1.25f, 1.0f), (float) root_power);
        }
        const float v = rrc_dot(s_rrc_im + h_step*V29_FILTER_STEPS);
        const float sim = fmul(v, agc_scaling);
        const float zr = k.sine[(carrier_phase + (1u << 30)) >> 21];     // dds_lookup_complexf, src/dds_float.c:2177
        const float zi = k.sine[carrier_phase >> 21];
        const float zzre = fsub(fmul(h_sre, zr), fmul(sim, zi));
        const float zzim = fsub(fmul(-h_sre, zi), fmul(sim, zr));
        eq_put_step += V29_COEFF_SETS*10/(3*2);
        // process_half_baud, first part (src/v29rx.c:516-525)
        eq_buf[(2*eq_step)*32] = zzre;
        eq_buf[(2*eq_step + 1)*32] = zzim;
        if (++eq_step >= V29_EQ_LEN)
            eq_step = 0;
        if ((baud_half ^= 1))
        {
            carrier_phase += (unsigned int) phase_rate;
            return false;
        }
        return true;
    }

    __device__ __forceinline__ void baud(const Consts &k)
    {
        process_baud(k);
        carrier_phase += (unsigned int) phase_rate;
    }
};

#define V29_SMEM_FLOATS_PER_WARP    ((2*V29_EQ_LEN + 2*V29_EQ_LEN + V29_FILTER_STEPS + 16)*32)

// 32 channels per CTA (one warp): few channels exist (thousands), so spread them over all SMs.
__global__ void __launch_bounds__(32) v29_rx_kernel(const Args a)
{
    extern __shared__ float smem[];
    float *s_rrc_re = smem;
    float *s_rrc_im = smem + V29_COEFF_SETS*V29_FILTER_STEPS;
    float *lane_base = smem + 2*V29_COEFF_SETS*V29_FILTER_STEPS;
    const int lane = threadIdx.x;
    for (int i = lane;  i < V29_COEFF_SETS*V29_FILTER_STEPS;  i += 32)
    {
        s_rrc_re[i] = a.k.rrc_re[i];
        s_rrc_im[i] = a.k.rrc_im[i];
    }
    __syncwarp();
    const int c = blockIdx.x*32 + lane;
    if (c >= a.channels)
        return;
    const size_t C = a.channels;

    Rx r;
    r.c = c;
    r.channels = a.channels;
    r.fstate = a.fstate;
    r.eq_coeff = lane_base + lane;
    r.eq_buf = lane_base + (2*V29_EQ_LEN)*32 + lane;
    r.rrc = lane_base + (4*V29_EQ_LEN)*32 + lane;
    r.diff_angles = (int *) (lane_base + (4*V29_EQ_LEN + V29_FILTER_STEPS)*32 + lane);
    const float *F = a.fstate;
    const int *I = a.istate;
#define LF(f) F[(size_t) (f)*C + c]
#define LI(f) I[(size_t) (f)*C + c]
    r.agc_scaling = LF(F_AGC);
    r.agc_scaling_save = LF(F_AGC_SAVE);
    r.training_error = LF(F_TRAINING_ERROR);
    r.track_p = LF(F_TRACK_P);
    r.track_i = LF(F_TRACK_I);
    r.lbe0 = LF(F_LBE0);
    r.lbe1 = LF(F_LBE1);
    r.hbe0 = LF(F_HBE0);
    r.hbe1 = LF(F_HBE1);
    r.dc0 = LF(F_DC0);
    r.dc1 = LF(F_DC1);
    r.baud_phase = LF(F_BAUD_PHASE);
    for (int i = 0;  i < 2*V29_EQ_LEN;  i++)
    {
        r.eq_coeff[i*32] = LF(F_EQ_COEFF + i);
        r.eq_buf[i*32] = LF(F_EQ_BUF + i);
    }
    for (int i = 0;  i < V29_FILTER_STEPS;  i++)
        r.rrc[i*32] = LF(F_RRC + i);
    r.bit_rate = LI(I_BIT_RATE);
    r.training_cd = LI(I_TRAINING_CD);
    r.rrc_step = LI(I_RRC_STEP);
    r.scramble_reg = (unsigned int) LI(I_SCRAMBLE);
    r.training_scramble_reg = (unsigned int) LI(I_TRAIN_SCRAMBLE);
    r.training_stage = LI(I_STAGE);
    r.training_count = LI(I_TRAIN_COUNT);
    r.last_sample = LI(I_LAST_SAMPLE);
    r.signal_present = LI(I_SIGNAL_PRESENT);
    r.drop_pending = LI(I_DROP_PENDING);
    r.low_samples = LI(I_LOW_SAMPLES);
    r.high_sample = LI(I_HIGH_SAMPLE);
    r.carrier_phase = (unsigned int) LI(I_CARRIER_PHASE);
    r.phase_rate = LI(I_PHASE_RATE);
    r.phase_rate_save = LI(I_PHASE_RATE_SAVE);
    r.power = LI(I_POWER);
    r.on_power = LI(I_ON_POWER);
    r.off_power = LI(I_OFF_POWER);
    r.eq_step = LI(I_EQ_STEP);
    r.eq_put_step = LI(I_EQ_PUT_STEP);
    r.eq_skip = LI(I_EQ_SKIP);
    r.baud_half = LI(I_BAUD_HALF);
    r.last_angle0 = LI(I_LAST_ANGLE0);
    r.last_angle1 = LI(I_LAST_ANGLE1);
    for (int i = 0;  i < 16;  i++)
        r.diff_angles[i*32] = LI(I_DIFF_ANGLES + i);
    r.constellation_state = LI(I_CONSTELLATION);
    r.total_timing = LI(I_TOTAL_TIMING);
    r.bits = a.bits + (size_t) c*a.bits_cap;
    r.bits_cap = (int) a.bits_cap;
    r.nbits = 0;
    r.syms = (a.syms)  ?  (a.syms + (size_t) c*a.sym_cap)  :  NULL;
    r.sym_cap = (int) a.sym_cap;
    r.nsyms = 0;

    // Lanes are kept in step on symbol time: one trip of this loop takes every receiving channel
    // through one whole baud - two T/2 instants, each reached after one or two input samples - so the
    // expensive parts (imaginary FIR, equalizer, training/slicer) run with all lanes converged even
    // though the channels' symbol clocks sit at different sample phases.  Channels without carrier (or
    // parked) simply consume up to four samples per trip.  Each channel still sees its own samples in
    // order, which is all the reference's per-channel semantics require.
    const int16_t *row = a.amp + (long long) c*a.stride;
    int pos = 0;
#pragma unroll 1
    while (pos < a.n)
    {
#pragma unroll 1
        for (int h = 0;  h < 2;  h++)
        {
            // A channel that enters the trip half-way through a baud sits out the first slot, so that
            // every channel completes its baud in the second slot (and is baud-aligned from then on).
            if (h == 0  &&  r.baud_half)
                continue;
            bool due = false;
#pragma unroll 1
            for (int q = 0;  q < 2;  q++)
            {
                if (pos < a.n  &&  !due)
                {
                    due = r.front(a.k, s_rrc_re, __ldg(row + pos));
                    pos++;
                }
            }
            if (due)
            {
                if (r.half(a.k, s_rrc_im))
                    r.baud(a.k);
            }
        }
    }

    float *FW = a.fstate;
    int *IW = a.istate;
#define SF(f, v) FW[(size_t) (f)*C + c] = (v)
#define SI(f, v) IW[(size_t) (f)*C + c] = (int) (v)
    SF(F_AGC, r.agc_scaling);
    SF(F_AGC_SAVE, r.agc_scaling_save);
    SF(F_TRAINING_ERROR, r.training_error);
    SF(F_TRACK_P, r.track_p);
    SF(F_TRACK_I, r.track_i);
    SF(F_LBE0, r.lbe0);
    SF(F_LBE1, r.lbe1);
    SF(F_HBE0, r.hbe0);
    SF(F_HBE1, r.hbe1);
    SF(F_DC0, r.dc0);
    SF(F_DC1, r.dc1);
    SF(F_BAUD_PHASE, r.baud_phase);
    for (int i = 0;  i < 2*V29_EQ_LEN;  i++)
    {
        SF(F_EQ_COEFF + i, r.eq_coeff[i*32]);
        SF(F_EQ_BUF + i, r.eq_buf[i*32]);
    }
    for (int i = 0;  i < V29_FILTER_STEPS;  i++)
        SF(F_RRC + i, r.rrc[i*32]);
    SI(I_BIT_RATE, r.bit_rate);
    SI(I_TRAINING_CD, r.training_cd);
    SI(I_RRC_STEP, r.rrc_step);
    SI(I_SCRAMBLE, r.scramble_reg);
    SI(I_TRAIN_SCRAMBLE, r.training_scramble_reg);
    SI(I_STAGE, r.training_stage);
    SI(I_TRAIN_COUNT, r.training_count);
    SI(I_LAST_SAMPLE, r.last_sample);
    SI(I_SIGNAL_PRESENT, r.signal_present);
    SI(I_DROP_PENDING, r.drop_pending);
    SI(I_LOW_SAMPLES, r.low_samples);
    SI(I_HIGH_SAMPLE, r.high_sample);
    SI(I_CARRIER_PHASE, r.carrier_phase);
    SI(I_PHASE_RATE, r.phase_rate);
    SI(I_PHASE_RATE_SAVE, r.phase_rate_save);
    SI(I_POWER, r.power);
    SI(I_EQ_STEP, r.eq_step);
    SI(I_EQ_PUT_STEP, r.eq_put_step);
    SI(I_EQ_SKIP, r.eq_skip);
    SI(I_BAUD_HALF, r.baud_half);
    SI(I_LAST_ANGLE0, r.last_angle0);
    SI(I_LAST_ANGLE1, r.last_angle1);
    for (int i = 0;  i < 16;  i++)
        SI(I_DIFF_ANGLES + i, r.diff_angles[i*32]);
    SI(I_CONSTELLATION, r.constellation_state);
    SI(I_TOTAL_TIMING, r.total_timing);
    a.nbits[c] = r.nbits;
    if (a.nsyms)
        a.nsyms[c] = r.nsyms;
}

}  // namespace v29

// ------------------------------------------------------------------------------------------
// host side: banks

struct span_b200_v29_bank_s
{
    span_b200_ctx_t *ctx;
    int channels;
    int bit_rate;
    float *fstate;
    int *istate;
    float *d_rrc_re;
    float *d_rrc_im;
    float *d_sine;
    unsigned short *d_sqrt;
    v29::Consts k;
    signed char *bits;
    long long bits_cap;
    int *nbits;
    span_b200_v29_symbol_t *syms;
    long long sym_cap;
    int *nsyms;
    int want_symbols;
    int16_t *d_in;
    size_t d_in_bytes;
    cudaStream_t last_stream;
    bool have_last;
    int on_power;
    int off_power;
};

static int v29_upload_const()
{
    static const float constellation[16][2] =
    {
        { 3.0f,  0.0f}, { 1.0f,  1.0f}, { 0.0f,  3.0f}, {-1.0f,  1.0f}, {-3.0f,  0.0f}, {-1.0f, -1.0f}, { 0.0f, -3.0f}, { 1.0f, -1.0f},
        { 5.0f,  0.0f}, { 3.0f,  3.0f}, { 0.0f,  5.0f}, {-3.0f,  3.0f}, {-5.0f,  0.0f}, {-3.0f, -3.0f}, { 0.0f, -5.0f}, { 3.0f, -3.0f}
    };
    // The slicer map: nearest constellation point for each 0.5 x 0.5 cell of [-5, 5) x [-5, 5), evaluated
    // at the cell centre, first index winning ties (the rule of src/make_v29_constellation_map.c:63-84).
    // The table actually compiled into the reference (src/v29rx.c:119-143) deviates from that rule in
    // eight cells on the diagonals, which are listed explicitly.
    unsigned char space_map[20][20];
    for (int ire = 0;  ire < 20;  ire++)
    {
        const double re = (ire - 10)/2.0 + 0.25;
        for (int iim = 0;  iim < 20;  iim++)
        {
            const double im = (iim - 10)/2.0 + 0.25;
            int best = 0;
            double best_distance = 1000000.0;
            for (int l = 0;  l < 16;  l++)
            {
                const double d = (re - constellation[l][0])*(re - constellation[l][0])
                               + (im - constellation[l][1])*(im - constellation[l][1]);
                if (d < best_distance)
                {
                    best = l;
                    best_distance = d;
                }
            }
            space_map[ire][iim] = (unsigned char) best;
        }
    }
    static const unsigned char exceptions[8][3] =
    {
        {5, 6, 13}, {5, 13, 11}, {6, 5, 13}, {6, 14, 11}, {13, 5, 15}, {13, 14, 9}, {14, 6, 15}, {14, 13, 9}
    };
    for (int i = 0;  i < 8;  i++)
        space_map[exceptions[i][0]][exceptions[i][1]] = exceptions[i][2];
    static const unsigned char ps9600[8] = {4, 0, 2, 6, 7, 3, 1, 5};
    static const unsigned char ps4800[4] = {0, 2, 3, 1};
    static const int cdcd[6] = {0, 11, 0, 3, 0, 2};
    CK(cudaMemcpyToSymbol(v29::c_space_map_9600, space_map, sizeof(space_map)));
    CK(cudaMemcpyToSymbol(v29::c_constellation, constellation, sizeof(constellation)));
    CK(cudaMemcpyToSymbol(v29::c_phase_steps_9600, ps9600, sizeof(ps9600)));
    CK(cudaMemcpyToSymbol(v29::c_phase_steps_4800, ps4800, sizeof(ps4800)));
    CK(cudaMemcpyToSymbol(v29::c_cdcd_pos, cdcd, sizeof(cdcd)));
    return 0;
}

// v29_rx_restart(s, bit_rate, false) after v29_rx_init's memset: src/v29rx.c:1019-1145
static int v29_reset_channels(span_b200_v29_bank_t *b, int first, int count, int bit_rate)
{
    const size_t C = b->channels;
    std::vector<float> f((size_t) v29::F_COUNT*count, 0.0f);
    std::vector<int> in((size_t) v29::I_COUNT*count, 0);
#define HF(field, i) f[(size_t) (field)*count + (i)]
#define HI(field, i) in[(size_t) (field)*count + (i)]
    for (int i = 0;  i < count;  i++)
    {
        HF(v29::F_AGC, i) = b->k.agc_initial;
        HF(v29::F_TRACK_P, i) = 8000000.0f;
        HF(v29::F_TRACK_I, i) = 8000.0f;
        HF(v29::F_EQ_COEFF + 2*V29_EQ_PRE_LEN, i) = 3.0f;
        HI(v29::I_BIT_RATE, i) = bit_rate;
        HI(v29::I_TRAINING_CD, i) = (bit_rate == 9600)  ?  0  :  (bit_rate == 7200)  ?  2  :  4;
        HI(v29::I_TRAIN_SCRAMBLE, i) = 0x2A;
        HI(v29::I_STAGE, i) = v29::STAGE_SYMBOL_ACQUISITION;
        HI(v29::I_PHASE_RATE, i) = b->k.rate_nominal;
        HI(v29::I_ON_POWER, i) = b->on_power;
        HI(v29::I_OFF_POWER, i) = b->off_power;
        HI(v29::I_EQ_PUT_STEP, i) = V29_COEFF_SETS*10/(3*2) - 1;
    }
    for (int fld = 0;  fld < v29::F_COUNT;  fld++)
        CK(cudaMemcpy(b->fstate + fld*C + first, &f[(size_t) fld*count], sizeof(float)*count, cudaMemcpyHostToDevice));
    for (int fld = 0;  fld < v29::I_COUNT;  fld++)
        CK(cudaMemcpy(b->istate + fld*C + first, &in[(size_t) fld*count], sizeof(int)*count, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" span_b200_v29_bank_t *span_b200_v29_bank_create(span_b200_ctx_t *ctx, int channels, int bit_rate, int want_symbols)
{
    if (ctx == NULL  ||  channels <= 0  ||  (bit_rate != 9600  &&  bit_rate != 7200  &&  bit_rate != 4800))
    {
        sb_set_error("bad V.29 bank arguments (bit rate must be 9600, 7200 or 4800)");      // src/v29rx.c:1102-1111
        return NULL;
    }
    CKP(cudaSetDevice(span_b200_ctx_device(ctx)));
    if (v29_upload_const() != 0)
        return NULL;
    span_b200_v29_bank_t *b = new span_b200_v29_bank_s();
    b->ctx = ctx;
    b->channels = channels;
    b->bit_rate = bit_rate;
    b->want_symbols = (want_symbols != 0);
    std::vector<float> re;
    std::vector<float> im;
    std::vector<float> st;
    std::vector<unsigned short> sq;
    godard_desc_t g;
    make_v29_rrc(re, im);
    make_sine_table(st);
    make_sqrt_table(sq);
    make_godard(g);
    CKP(cudaMalloc(&b->d_rrc_re, sizeof(float)*re.size()));
    CKP(cudaMalloc(&b->d_rrc_im, sizeof(float)*im.size()));
    CKP(cudaMalloc(&b->d_sine, sizeof(float)*st.size()));
    CKP(cudaMalloc(&b->d_sqrt, sizeof(unsigned short)*sq.size()));
    CKP(cudaMemcpy(b->d_rrc_re, re.data(), sizeof(float)*re.size(), cudaMemcpyHostToDevice));
    CKP(cudaMemcpy(b->d_rrc_im, im.data(), sizeof(float)*im.size(), cudaMemcpyHostToDevice));
    CKP(cudaMemcpy(b->d_sine, st.data(), sizeof(float)*st.size(), cudaMemcpyHostToDevice));
    CKP(cudaMemcpy(b->d_sqrt, sq.data(), sizeof(unsigned short)*sq.size(), cudaMemcpyHostToDevice));
    b->k.rrc_re = b->d_rrc_re;
    b->k.rrc_im = b->d_rrc_im;
    b->k.sine = b->d_sine;
    b->k.sqrt_tab = b->d_sqrt;
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
    b->k.rate_nominal = (int32_t) (1700.0f*65536.0f*65536.0f/8000);
    b->k.rate_low = (int32_t) ((1700.0f - 20.0f)*65536.0f*65536.0f/8000);
    b->k.rate_high = (int32_t) ((1700.0f + 20.0f)*65536.0f*65536.0f/8000);
    b->k.phase_p45 = (int32_t) ((uint32_t) (45.0f*65536.0f*65536.0f/360.0f));
    b->k.phase_m45 = (int32_t) ((uint32_t) ((360.0f + -45.0f)*65536.0f*65536.0f/360.0f));
    b->k.agc_initial = (1.25f/1.000000f)/735.0f;                        // src/v29rx.c:1078
    b->k.eq_delta = 0.21f/V29_EQ_LEN;                                    // src/v29rx.c:97,240
    // v29_rx_set_signal_cutoff(s, -28.5f): src/v29rx.c:163-168,1129
    b->on_power = (int32_t) (host_power_meter_level_dbm0(-28.5f + 2.5f)*0.4f);
    b->off_power = (int32_t) (host_power_meter_level_dbm0(-28.5f - 2.5f)*0.4f);
    const size_t C = channels;
    CKP(cudaMalloc(&b->fstate, sizeof(float)*v29::F_COUNT*C));
    CKP(cudaMalloc(&b->istate, sizeof(int)*v29::I_COUNT*C));
    CKP(cudaMalloc(&b->nbits, sizeof(int)*C));
    CKP(cudaMalloc(&b->nsyms, sizeof(int)*C));
    CKP(cudaMemset(b->nbits, 0, sizeof(int)*C));
    CKP(cudaMemset(b->nsyms, 0, sizeof(int)*C));
    if (v29_reset_channels(b, 0, channels, bit_rate) != 0)
        return NULL;
    return b;
}

extern "C" void span_b200_v29_bank_destroy(span_b200_v29_bank_t *b)
{
    if (b == NULL)
        return;
    cudaSetDevice(span_b200_ctx_device(b->ctx));
    if (b->have_last)
        cudaStreamSynchronize(b->last_stream);
    cudaFree(b->fstate);
    cudaFree(b->istate);
    cudaFree(b->d_rrc_re);
    cudaFree(b->d_rrc_im);
    cudaFree(b->d_sine);
    cudaFree(b->d_sqrt);
    cudaFree(b->bits);
    cudaFree(b->nbits);
    cudaFree(b->syms);
    cudaFree(b->nsyms);
    cudaFree(b->d_in);
    delete b;
}

extern "C" int span_b200_v29_bank_channels(const span_b200_v29_bank_t *b)
{
    return b->channels;
}

extern "C" int span_b200_v29_bank_restart(span_b200_v29_bank_t *b, int first, int count, int bit_rate)
{
    if (first < 0  ||  count < 0  ||  first + count > b->channels  ||  (bit_rate != 9600  &&  bit_rate != 7200  &&  bit_rate != 4800))
    {
        sb_set_error("bad restart arguments");
        return -1;                                  // src/v29rx.c:1033
    }
    CK(cudaSetDevice(span_b200_ctx_device(b->ctx)));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    return v29_reset_channels(b, first, count, bit_rate);
}

extern "C" int span_b200_v29_bank_set_signal_cutoff(span_b200_v29_bank_t *b, int first, int count, float cutoff)
{
    if (first < 0  ||  count < 0  ||  first + count > b->channels)
    {
        sb_set_error("channel range out of bounds");
        return -1;
    }
    CK(cudaSetDevice(span_b200_ctx_device(b->ctx)));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    // src/v29rx.c:163-168
    const int on = (int32_t) (host_power_meter_level_dbm0(cutoff + 2.5f)*0.4f);
    const int off = (int32_t) (host_power_meter_level_dbm0(cutoff - 2.5f)*0.4f);
    if (first == 0  &&  count == b->channels)
    {
        b->on_power = on;
        b->off_power = off;
    }
    std::vector<int> v(count, on);
    CK(cudaMemcpy(b->istate + (size_t) v29::I_ON_POWER*b->channels + first, v.data(), sizeof(int)*count, cudaMemcpyHostToDevice));
    v.assign(count, off);
    CK(cudaMemcpy(b->istate + (size_t) v29::I_OFF_POWER*b->channels + first, v.data(), sizeof(int)*count, cudaMemcpyHostToDevice));
    return 0;
}

// v29_rx_fillin(): integer bookkeeping only (src/v29rx.c:967-996); done on the host copy of four fields.
extern "C" int span_b200_v29_bank_fillin(span_b200_v29_bank_t *b, int first, int count, int samples)
{
    if (first < 0  ||  count < 0  ||  first + count > b->channels  ||  samples < 0)
    {
        sb_set_error("bad fillin arguments");
        return -1;
    }
    CK(cudaSetDevice(span_b200_ctx_device(b->ctx)));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    const size_t C = b->channels;
    std::vector<int> present(count), stage(count), phase(count), rate(count), put(count);
    CK(cudaMemcpy(present.data(), b->istate + v29::I_SIGNAL_PRESENT*C + first, sizeof(int)*count, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(stage.data(), b->istate + v29::I_STAGE*C + first, sizeof(int)*count, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(phase.data(), b->istate + v29::I_CARRIER_PHASE*C + first, sizeof(int)*count, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rate.data(), b->istate + v29::I_PHASE_RATE*C + first, sizeof(int)*count, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(put.data(), b->istate + v29::I_EQ_PUT_STEP*C + first, sizeof(int)*count, cudaMemcpyDeviceToHost));
    for (int c = 0;  c < count;  c++)
    {
        if (present[c] <= 0  ||  stage[c] == v29::STAGE_PARKED)
            continue;
        unsigned int ph = (unsigned int) phase[c];
        for (int i = 0;  i < samples;  i++)
        {
            ph += (unsigned int) rate[c];
            put[c] -= V29_COEFF_SETS;
            if (put[c] <= 0)
                put[c] += V29_COEFF_SETS*10/(3*2);
        }
        phase[c] = (int) ph;
    }
    CK(cudaMemcpy(b->istate + v29::I_CARRIER_PHASE*C + first, phase.data(), sizeof(int)*count, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b->istate + v29::I_EQ_PUT_STEP*C + first, put.data(), sizeof(int)*count, cudaMemcpyHostToDevice));
    return 0;
}

static int v29_ensure(void **p, size_t have_elems, size_t want_elems, size_t elem)
{
    (void) have_elems;
    if (*p)
        CK(cudaFree(*p));
    *p = NULL;
    CK(cudaMalloc(p, want_elems*elem));
    return 0;
}

extern "C" int span_b200_v29_bank_rx_device(span_b200_v29_bank_t *b, const int16_t *d_amp, int64_t stride, int n, void *stream)
{
    if (b == NULL  ||  n < 0  ||  (n > 0  &&  d_amp == NULL))
    {
        sb_set_error("bad rx arguments");
        return -1;
    }
    CK(cudaSetDevice(span_b200_ctx_device(b->ctx)));
    cudaStream_t st = (stream)  ?  (cudaStream_t) stream  :  (cudaStream_t) sb_ctx_stream(b->ctx);
    if (b->have_last  &&  b->last_stream != st)
        CK(cudaStreamSynchronize(b->last_stream));
    // Worst case: 4 bits per baud at 2400 baud/8000 Hz plus timing drift, plus status reports.
    const long long want_bits = (long long) n*3/2 + 64;
    if (b->bits_cap < want_bits)
    {
        if (b->have_last)
            CK(cudaStreamSynchronize(b->last_stream));
        if (v29_ensure((void **) &b->bits, 0, (size_t) want_bits*b->channels, 1) != 0)
            return -1;
        b->bits_cap = want_bits;
    }
    const long long want_syms = (long long) n*2/5 + 16;
    if (b->want_symbols  &&  b->sym_cap < want_syms)
    {
        if (b->have_last)
            CK(cudaStreamSynchronize(b->last_stream));
        if (v29_ensure((void **) &b->syms, 0, (size_t) want_syms*b->channels, sizeof(span_b200_v29_symbol_t)) != 0)
            return -1;
        b->sym_cap = want_syms;
    }
    v29::Args a;
    a.amp = d_amp;
    a.stride = stride;
    a.n = n;
    a.channels = b->channels;
    a.fstate = b->fstate;
    a.istate = b->istate;
    a.bits = b->bits;
    a.bits_cap = b->bits_cap;
    a.nbits = b->nbits;
    a.syms = (b->want_symbols)  ?  b->syms  :  NULL;
    a.sym_cap = b->sym_cap;
    a.nsyms = b->nsyms;
    a.k = b->k;
    const int smem = (int) sizeof(float)*(2*V29_COEFF_SETS*V29_FILTER_STEPS + V29_SMEM_FLOATS_PER_WARP);
    static bool configured = false;
    if (!configured)
    {
        CK(cudaFuncSetAttribute(v29::v29_rx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    v29::v29_rx_kernel<<<(b->channels + 31)/32, 32, smem, st>>>(a);
    CK(cudaGetLastError());
    b->last_stream = st;
    b->have_last = true;
    return 0;
}

extern "C" int span_b200_v29_bank_rx_host(span_b200_v29_bank_t *b, const int16_t *h_amp, int64_t stride, int n, void *stream)
{
    if (b == NULL  ||  n < 0  ||  (n > 0  &&  h_amp == NULL))
    {
        sb_set_error("bad rx arguments");
        return -1;
    }
    CK(cudaSetDevice(span_b200_ctx_device(b->ctx)));
    cudaStream_t st = (stream)  ?  (cudaStream_t) stream  :  (cudaStream_t) sb_ctx_stream(b->ctx);
    if (b->have_last  &&  b->last_stream != st)
        CK(cudaStreamSynchronize(b->last_stream));
    const size_t want = sizeof(int16_t)*(size_t) n*b->channels + 16;
    if (b->d_in_bytes < want)
    {
        if (b->have_last)
            CK(cudaStreamSynchronize(b->last_stream));
        if (b->d_in)
            CK(cudaFree(b->d_in));
        b->d_in = NULL;
        CK(cudaMalloc(&b->d_in, want));
        b->d_in_bytes = want;
    }
    if (n > 0)
        CK(cudaMemcpy2DAsync(b->d_in, sizeof(int16_t)*(size_t) n, h_amp, sizeof(int16_t)*stride, sizeof(int16_t)*(size_t) n,
                             b->channels, cudaMemcpyHostToDevice, st));
    return span_b200_v29_bank_rx_device(b, b->d_in, n, n, (void *) st);
}

extern "C" int span_b200_v29_bank_counts(span_b200_v29_bank_t *b, int32_t *nbits, int32_t *nsyms)
{
    CK(cudaSetDevice(span_b200_ctx_device(b->ctx)));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    if (nbits)
        CK(cudaMemcpy(nbits, b->nbits, sizeof(int)*(size_t) b->channels, cudaMemcpyDeviceToHost));
    if (nsyms)
        CK(cudaMemcpy(nsyms, b->nsyms, sizeof(int)*(size_t) b->channels, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int64_t span_b200_v29_bank_bits(span_b200_v29_bank_t *b, int channel, int8_t *out, int64_t max)
{
    if (channel < 0  ||  channel >= b->channels)
        return -1;
    CK(cudaSetDevice(span_b200_ctx_device(b->ctx)));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    int n = 0;
    CK(cudaMemcpy(&n, b->nbits + channel, sizeof(int), cudaMemcpyDeviceToHost));
    long long k = n;
    if (k > b->bits_cap)
        k = b->bits_cap;
    if (k > max)
        k = max;
    if (k > 0)
        CK(cudaMemcpy(out, b->bits + (size_t) channel*b->bits_cap, (size_t) k, cudaMemcpyDeviceToHost));
    return k;
}

extern "C" int64_t span_b200_v29_bank_symbols(span_b200_v29_bank_t *b, int channel, span_b200_v29_symbol_t *out, int64_t max)
{
    if (channel < 0  ||  channel >= b->channels  ||  !b->want_symbols)
        return -1;
    CK(cudaSetDevice(span_b200_ctx_device(b->ctx)));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
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

extern "C" int span_b200_v29_bank_output_layout(span_b200_v29_bank_t *b, const int8_t **d_bits, int64_t *bits_cap,
                                                const int32_t **d_nbits, const span_b200_v29_symbol_t **d_syms,
                                                int64_t *sym_cap, const int32_t **d_nsyms)
{
    if (d_bits)
        *d_bits = (const int8_t *) b->bits;
    if (bits_cap)
        *bits_cap = b->bits_cap;
    if (d_nbits)
        *d_nbits = b->nbits;
    if (d_syms)
        *d_syms = b->syms;
    if (sym_cap)
        *sym_cap = b->sym_cap;
    if (d_nsyms)
        *d_nsyms = b->nsyms;
    return 0;
}

// Receiver status of one channel: v29_rx_equalizer_state / carrier_frequency / symbol_timing_correction /
// signal_power (src/v29rx.c:145-161,180-195) material.
extern "C" int span_b200_v29_bank_channel_state(span_b200_v29_bank_t *b, int channel, float *eq_coeff, int32_t *info)
{
    if (channel < 0  ||  channel >= b->channels)
        return -1;
    CK(cudaSetDevice(span_b200_ctx_device(b->ctx)));
    if (b->have_last)
        CK(cudaStreamSynchronize(b->last_stream));
    const size_t C = b->channels;
    if (eq_coeff)
    {
        for (int i = 0;  i < 2*V29_EQ_LEN;  i++)
            CK(cudaMemcpy(&eq_coeff[i], b->fstate + (size_t) (v29::F_EQ_COEFF + i)*C + channel, sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (info)
    {
        static const int fields[10] = {v29::I_STAGE, v29::I_PHASE_RATE, v29::I_EQ_PUT_STEP, v29::I_SIGNAL_PRESENT, -1,
                                       v29::I_TOTAL_TIMING, v29::I_CONSTELLATION, v29::I_CARRIER_PHASE, v29::I_POWER, v29::I_BIT_RATE};
        for (int i = 0;  i < 10;  i++)
        {
            if (fields[i] >= 0)
                CK(cudaMemcpy(&info[i], b->istate + (size_t) fields[i]*C + channel, sizeof(int), cudaMemcpyDeviceToHost));
        }
        CK(cudaMemcpy(&info[4], b->fstate + (size_t) v29::F_AGC*C + channel, sizeof(float), cudaMemcpyDeviceToHost));
    }
    return 0;
}
