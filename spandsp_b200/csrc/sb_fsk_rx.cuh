// sb_fsk_rx.cuh - the FSK receiver (V.21, V.23, Bell 103/202, Weitbrecht) of src/fsk.c:396-626: non-coherent
// demodulation by quadrature correlation with the two tones over a sliding one-baud window, integer DDS, power
// meter carrier detect, and the three bit-clock modes (synchronous, asynchronous, framed).  All integer
// arithmetic; the results are bit-exact by construction.  Written __host__ __device__ like the modem receivers,
// so that tests/hostsim can run the very same code on the CPU.
#pragma once

#include <stdint.h>
#include <math.h>
#include <string.h>
#include <vector>

#include <cuda_runtime.h>

#if !defined(SB_HD)
#define SB_HD __host__ __device__ __forceinline__
#endif

namespace sbf {

#define SBF_MAX_WINDOW      128         // FSK_MAX_WINDOW_LEN, src/spandsp/fsk.h:140
#define SBF_SINE_WORDS      257         // one quadrant + 1, src/dds_int.c:48-55
#define SBF_SINE_PAD        260

enum
{
    FRAME_MODE_ASYNC = 0,               // src/spandsp/fsk.h:126-128
    FRAME_MODE_SYNC = 1,
    FRAME_MODE_FRAMED = 2
};

enum
{
    PARITY_NONE = 0,                    // src/spandsp/async.h:151-157
    PARITY_EVEN,
    PARITY_ODD,
    PARITY_MARK,
    PARITY_SPACE
};

// Per-channel state, one int per field, stored [field][channel]
enum
{
    K_BAUD_RATE = 0, K_FRAMING_MODE, K_DATA_BITS, K_PARITY, K_STOP_BITS, K_TOTAL_DATA_BITS, K_ON_POWER, K_OFF_POWER,
    K_READING, K_LAST_SAMPLE, K_SIGNAL_PRESENT, K_RATE0, K_RATE1, K_ACC0, K_ACC1, K_SPAN,
    K_DOT0_RE, K_DOT0_IM, K_DOT1_RE, K_DOT1_IM, K_BUF_PTR, K_FRAME_POS, K_FRAME_IN_PROGRESS, K_BAUD_PHASE, K_LAST_BIT,
    K_SHIFT, K_PARITY_ERRORS, K_FRAMING_ERRORS, K_COUNT
};

// The quarter-wave table of src/dds_int.c:55-315: round(32767*sin(pi/2*i/256)), i = 0..256
static inline void make_dds_int_table(std::vector<short> &t)
{
    t.resize(SBF_SINE_PAD);
    for (int i = 0;  i < SBF_SINE_PAD;  i++)
        t[i] = 0;
    for (int i = 0;  i < SBF_SINE_WORDS;  i++)
        t[i] = (short) floor(32767.0*sin(1.5707963267948966*(double) i/256.0) + 0.5);
}

// dds_phase_rate() (src/dds_int.c:316-319)
static inline int32_t host_dds_int_phase_rate(float frequency)
{
    return (int32_t) (frequency*65536.0f*65536.0f/8000);
}

// power_meter_level_dbm0() (src/power_meter.c:82-93)
static inline int32_t host_level_dbm0(float level)
{
    float l;

    level -= (3.14f + 3.02f);
    if (level > 0.0)
        level = 0.0;
    l = powf(10.0f, level/10.0f)*(32767.0f*32767.0f);
    return (int32_t) l;
}

struct FskLoader
{
    const int *state;
    size_t channels;
    size_t c;
    SB_HD void operator()(int field, int &v) const { v = state[(size_t) field*channels + c]; }
};

struct FskStorer
{
    int *state;
    size_t channels;
    size_t c;
    SB_HD void operator()(int field, int &v) const { state[(size_t) field*channels + c] = v; }
};

struct FskRx
{
    int baud_rate, framing_mode, data_bits, parity, stop_bits, total_data_bits, on_power, off_power;
    int reading, last_sample, signal_present, rate0, rate1, acc0, acc1, span;
    int dot0_re, dot0_im, dot1_re, dot1_im, buf_ptr, frame_pos, frame_in_progress, baud_phase, last_bit;
    int shift, parity_errors, framing_errors;

    int2 *win;                  // window element (tone j, slot k) at win[(j*wspan + k)*ls]
    int wspan;
    int ls;
    const short *sine;
    short *out;                 // the put_bit stream: bits / characters >= 0, SIG_STATUS_* < 0
    int out_cap;
    int nout;

    template <class V> SB_HD void visit(V &v)
    {
        v(K_BAUD_RATE, baud_rate);
        v(K_FRAMING_MODE, framing_mode);
        v(K_DATA_BITS, data_bits);
        v(K_PARITY, parity);
        v(K_STOP_BITS, stop_bits);
        v(K_TOTAL_DATA_BITS, total_data_bits);
        v(K_ON_POWER, on_power);
        v(K_OFF_POWER, off_power);
        v(K_READING, reading);
        v(K_LAST_SAMPLE, last_sample);
        v(K_SIGNAL_PRESENT, signal_present);
        v(K_RATE0, rate0);
        v(K_RATE1, rate1);
        v(K_ACC0, acc0);
        v(K_ACC1, acc1);
        v(K_SPAN, span);
        v(K_DOT0_RE, dot0_re);
        v(K_DOT0_IM, dot0_im);
        v(K_DOT1_RE, dot1_re);
        v(K_DOT1_IM, dot1_im);
        v(K_BUF_PTR, buf_ptr);
        v(K_FRAME_POS, frame_pos);
        v(K_FRAME_IN_PROGRESS, frame_in_progress);
        v(K_BAUD_PHASE, baud_phase);
        v(K_LAST_BIT, last_bit);
        v(K_SHIFT, shift);
        v(K_PARITY_ERRORS, parity_errors);
        v(K_FRAMING_ERRORS, framing_errors);
    }

    // put_bit / report_status_change without a status handler (src/fsk.c:347-354)
    SB_HD void put(int v)
    {
        if (nout < out_cap)
            out[nout] = (short) v;
        nout++;
    }

    // dds_lookup() (src/dds_int.c:340-356)
    SB_HD int lookup(unsigned int phase) const
    {
        phase >>= 22;
        unsigned int step = phase & 255u;
        if ((phase & 256u))
            step = 256u - step;
        const int amp = sine[step];
        return ((phase & 512u))  ?  -amp  :  amp;
    }

    // src/fsk.c:357-393
    SB_HD void put_frame(unsigned int frame)
    {
        frame &= 0xFFFFu;
        if (parity != PARITY_NONE)
        {
            const unsigned int parity_bit_a = (frame >> 15) & 1u;
            unsigned int parity_bit_b;
            frame &= 0x7FFFu;
            frame >>= (16 - total_data_bits);
            unsigned int x = frame & 0xFFu;                 // parity8() takes a uint8_t (spandsp/bit_operations.h:284-288)
            x = (x ^ (x >> 4)) & 0x0Fu;
            const unsigned int p8 = (0x6996u >> x) & 1u;
            switch (parity)
            {
            case PARITY_ODD:
                parity_bit_b = p8 ^ 1u;
                break;
            case PARITY_EVEN:
                parity_bit_b = p8;
                break;
            case PARITY_MARK:
                parity_bit_b = 1;
                break;
            default:
                parity_bit_b = 0;
                break;
            }
            if (parity_bit_a == parity_bit_b)
                put((int) frame);
            else
                parity_errors++;
        }
        else
        {
            frame >>= (16 - total_data_bits);
            put((int) frame);
        }
    }

    // One tone of the sliding correlation (src/fsk.c:414-431); returns the squared magnitude
    SB_HD int correlate(int j, int amp, int &acc, int rate, int &dre, int &dim)
    {
        int2 *slot = win + (size_t) (j*wspan + buf_ptr)*ls;
        const int2 old = *slot;
        const int re = lookup((unsigned int) acc + (1u << 30));
        const int im = lookup((unsigned int) acc);
        acc = (int) ((unsigned int) acc + (unsigned int) rate);
        int2 nw;
        nw.x = (re*amp) >> shift;
        nw.y = (im*amp) >> shift;
        *slot = nw;
        dre = (int) ((unsigned int) dre - (unsigned int) old.x + (unsigned int) nw.x);
        dim = (int) ((unsigned int) dim - (unsigned int) old.y + (unsigned int) nw.y);
        const int a = dre >> 15;
        const int b = dim >> 15;
        return (int) ((unsigned int) (a*a) + (unsigned int) (b*b));
    }

    // One pass of fsk_rx()'s sample loop (src/fsk.c:408-621).  The reference leaves the loop body early
    // (`continue`) while no carrier is present - WITHOUT advancing the window pointer; kept.
    SB_HD void sample(int amp)
    {
        const int sum0 = correlate(0, amp, acc0, rate0, dot0_re, dot0_im);
        const int sum1 = correlate(1, amp, acc1, rate1, dot1_re, dot1_im);
        const int x = amp >> 1;
        const int d = (int) (short) (x - last_sample);
        reading += ((d*d - reading) >> 4);                  // power_meter_update(), shift 4 (src/fsk.c:715)
        last_sample = x;
        const int power = reading;
        if (signal_present)
        {
            if (power < off_power)
            {
                if (--signal_present <= 0)
                {
                    put(-1);                                // SIG_STATUS_CARRIER_DOWN
                    baud_phase = 0;
                    return;
                }
            }
        }
        else
        {
            if (power < on_power)
            {
                baud_phase = 0;
                return;
            }
            if (baud_phase < (span >> 1) - 30)
            {
                baud_phase++;
                return;
            }
            signal_present = 1;
            baud_phase = 0;
            frame_pos = -2;
            frame_in_progress = 0;
            last_bit = 0;
            put(-2);                                        // SIG_STATUS_CARRIER_UP
        }
        const int baudstate = (sum0 < sum1);
        if (framing_mode == FRAME_MODE_SYNC)
        {
            if (last_bit != baudstate)
            {
                last_bit = baudstate;
                if (baud_phase < 8000*50)
                    baud_phase += (baud_rate >> 3);
                else
                    baud_phase -= (baud_rate >> 3);
            }
            if ((baud_phase += baud_rate) >= 8000*100)
            {
                baud_phase -= 8000*100;
                put(baudstate);
            }
        }
        else if (framing_mode == FRAME_MODE_ASYNC)
        {
            if (last_bit != baudstate)
            {
                last_bit = baudstate;
                baud_phase = 8000*50;
            }
            if ((baud_phase += baud_rate) >= 8000*100)
            {
                baud_phase -= 8000*100;
                put(baudstate);
            }
        }
        else
        {
            if (frame_pos == -2)
            {
                if (baudstate == 0)
                {
                    baud_phase = 8000*(100 - 40)/2;
                    frame_pos = -1;
                    frame_in_progress = 0;
                    last_bit = -1;
                }
            }
            else if (frame_pos == -1)
            {
                if (baudstate != 0)
                {
                    frame_pos = -2;
                }
                else
                {
                    baud_phase += baud_rate;
                    if (baud_phase >= 8000*100)
                    {
                        frame_pos = 0;
                        last_bit = baudstate;
                    }
                }
            }
            else
            {
                baud_phase += baud_rate;
                if (baud_phase >= 8000*(100 - 40))
                {
                    if (last_bit < 0)
                        last_bit = baudstate;
                    if (last_bit != baudstate)
                    {
                        frame_pos = -2;
                        framing_errors++;
                    }
                    else if (baud_phase >= 8000*100)
                    {
                        if (frame_pos++ > total_data_bits)
                        {
                            if (baudstate == 1)
                                put_frame((unsigned int) frame_in_progress);
                            else
                                framing_errors++;
                            frame_pos = -2;
                        }
                        else
                        {
                            frame_in_progress = (int) ((((unsigned int) frame_in_progress & 0xFFFFu) >> 1) | ((unsigned int) baudstate << 15));
                        }
                        baud_phase -= 8000*100;
                        last_bit = -1;
                    }
                }
            }
        }
        if (++buf_ptr >= span)
            buf_ptr = 0;
    }

    // One pass of fsk_rx_fillin()'s loop (src/fsk.c:637-666): a zero into the current window slot, tone phases advance.
    // The reference's loop does not step the window pointer (every pass clears the same slot); kept.
    SB_HD void fillin_sample()
    {
        for (int j = 0;  j < 2;  j++)
        {
            int2 *slot = win + (size_t) (j*wspan + buf_ptr)*ls;
            const int2 old = *slot;
            *slot = make_int2(0, 0);
            if (j == 0)
            {
                dot0_re = (int) ((unsigned int) dot0_re - (unsigned int) old.x);
                dot0_im = (int) ((unsigned int) dot0_im - (unsigned int) old.y);
                acc0 = (int) ((unsigned int) acc0 + (unsigned int) rate0);
            }
            else
            {
                dot1_re = (int) ((unsigned int) dot1_re - (unsigned int) old.x);
                dot1_im = (int) ((unsigned int) dot1_im - (unsigned int) old.y);
                acc1 = (int) ((unsigned int) acc1 + (unsigned int) rate1);
            }
        }
    }

    // fsk_rx_set_frame_parameters() (src/fsk.c:300-316)
    SB_HD void set_frame_parameters(int data, int par, int stop)
    {
        if (framing_mode == FRAME_MODE_FRAMED)
        {
            data_bits = data;
            parity = par;
            stop_bits = stop;
            total_data_bits = data_bits;
            if (parity != PARITY_NONE)
                total_data_bits++;
        }
    }

    // fsk_rx_restart() (src/fsk.c:670-722).  The window and the correlation sums are NOT cleared by the
    // reference's restart (only by fsk_rx_init's memset); kept.  rate0/rate1/on/off are computed by the host.
    SB_HD void restart(int baud, int mode, int r0, int r1, int on_pw, int off_pw)
    {
        baud_rate = baud;
        framing_mode = mode;
        if (framing_mode == FRAME_MODE_FRAMED)
            set_frame_parameters(8, PARITY_NONE, 1);
        on_power = on_pw;
        off_power = off_pw;
        rate0 = r0;
        rate1 = r1;
        acc0 = 0;
        acc1 = 0;
        last_sample = 0;
        span = 8000*100/baud_rate;
        if (span > SBF_MAX_WINDOW)
            span = SBF_MAX_WINDOW;
        shift = 0;
        for (int chop = span;  chop != 0;  chop >>= 1)
            shift++;
        baud_phase = 0;
        frame_pos = -2;
        frame_in_progress = 0;
        last_bit = 0;
        reading = 0;                                        // power_meter_init(&s->power, 4)
        signal_present = 0;
    }
};

// What the host passes for a restart / init of a channel range
struct FskSetup
{
    int baud_rate;
    int framing_mode;
    int rate0;
    int rate1;
    int on_power;
    int off_power;
};

struct FskArgs
{
    const int16_t *amp;             // [channel][sample], row stride in samples
    long long stride;
    int n;
    int channels;
    int *state;                     // [K_COUNT][channels]
    int2 *window;                   // [2][SBF_MAX_WINDOW][channels]
    const short *sine;              // SBF_SINE_PAD entries
    short *out;                     // [channel][out_cap]
    long long out_cap;
    int *nout;                      // [channels]
    int wspan;                      // window slots kept in shared memory (>= every channel's span)
};

#if defined(__CUDACC__)

__device__ __forceinline__ void fsk_bind(FskRx &r, const FskArgs &a, int *smem, int lane, int c, int &live)
{
    short *s_sine = (short *) smem;
    for (int i = lane;  i < SBF_SINE_PAD;  i += 32)
        s_sine[i] = a.sine[i];
    __syncwarp();
    r.sine = s_sine;
    r.win = (int2 *) (smem + SBF_SINE_PAD/2 + 2) + lane;
    r.wspan = a.wspan;
    r.ls = 32;
    FskLoader ld = {a.state, (size_t) a.channels, (size_t) c};
    r.visit(ld);
    // The live part of the window: global [tone][slot][channel] -> shared [tone][slot][lane].  After a restart to
    // a shorter span the window pointer may still sit beyond the new span (the reference does not reset it,
    // src/fsk.c:670-722) and the next sample lands there: those slots are live too.
    live = (r.buf_ptr + 1 > r.span)  ?  (r.buf_ptr + 1)  :  r.span;
    for (int j = 0;  j < 2;  j++)
    {
        for (int k = 0;  k < live;  k++)
            r.win[(size_t) (j*a.wspan + k)*32] = a.window[((size_t) (j*SBF_MAX_WINDOW + k))*a.channels + c];
    }
}

__device__ __forceinline__ void fsk_unbind(FskRx &r, const FskArgs &a, int c, int live)
{
    FskStorer st = {a.state, (size_t) a.channels, (size_t) c};
    r.visit(st);
    for (int j = 0;  j < 2;  j++)
    {
        for (int k = 0;  k < live;  k++)
            a.window[((size_t) (j*SBF_MAX_WINDOW + k))*a.channels + c] = r.win[(size_t) (j*a.wspan + k)*32];
    }
}

constexpr int fsk_smem_bytes(int wspan)
{
    return (SBF_SINE_PAD/2 + 2)*4 + wspan*2*32*8;
}

#if !defined(SBF_NO_FSK_KERNELS)       // sb_mct.cu embeds the receiver in its own kernel and does not want these

// fsk_rx() for 32 channels per CTA (one warp), one thread per channel: the receiver is a short sequential
// integer state machine per sample; the per-channel correlation windows sit in shared memory lane-interleaved
// (conflict-free for any per-lane window position), samples arrive as 16-byte loads per lane.
__global__ void __launch_bounds__(32) fsk_rx_kernel(const FskArgs a)
{
    extern __shared__ int fsk_smem[];
    const int lane = threadIdx.x;
    const int c = blockIdx.x*32 + lane;
    const bool active = (c < a.channels);
    FskRx r;
    int live;
    fsk_bind(r, a, fsk_smem, lane, (active)  ?  c  :  (a.channels - 1), live);
    if (!active)
        return;
    r.out = a.out + (size_t) c*a.out_cap;
    r.out_cap = (int) a.out_cap;
    r.nout = 0;
    const int16_t *row = a.amp + (long long) c*a.stride;
    int pos = 0;
    if ((((size_t) row) & 15) == 0)
    {
#pragma unroll 1
        for (  ;  pos + 8 <= a.n;  pos += 8)
        {
            const uint4 v = __ldg((const uint4 *) (row + pos));
            r.sample((short) (v.x & 0xFFFFu));
            r.sample((short) (v.x >> 16));
            r.sample((short) (v.y & 0xFFFFu));
            r.sample((short) (v.y >> 16));
            r.sample((short) (v.z & 0xFFFFu));
            r.sample((short) (v.z >> 16));
            r.sample((short) (v.w & 0xFFFFu));
            r.sample((short) (v.w >> 16));
        }
    }
#pragma unroll 1
    for (  ;  pos < a.n;  pos++)
        r.sample(__ldg(row + pos));
    fsk_unbind(r, a, c, live);
    a.nout[c] = r.nout;
}

// mode 0: fsk_rx_init() (state and window zeroed, then restart); 1: fsk_rx_restart() - neither touches the
// window; 2: fsk_rx_fillin(len = aux); 3: fsk_rx_set_frame_parameters(data_bits = aux & 255,
// parity = (aux >> 8) & 255, stop_bits = aux >> 16)
__global__ void __launch_bounds__(32) fsk_ctl_kernel(const FskArgs a, int first, int count, int mode, FskSetup su, int aux)
{
    extern __shared__ int fsk_smem[];
    const int lane = threadIdx.x;
    const int idx = blockIdx.x*32 + lane;
    const int c = first + idx;
    const bool active = (idx < count);
    FskRx r;
    if (mode == 2)
    {
        int live;
        fsk_bind(r, a, fsk_smem, lane, (active)  ?  c  :  first, live);
        if (!active)
            return;
        for (int i = 0;  i < aux;  i++)
            r.fillin_sample();
        fsk_unbind(r, a, c, live);
        return;
    }
    if (!active)
        return;
    if (mode == 0)
    {
        for (int f = 0;  f < K_COUNT;  f++)
            a.state[(size_t) f*a.channels + c] = 0;
        for (int k = 0;  k < 2*SBF_MAX_WINDOW;  k++)
            a.window[(size_t) k*a.channels + c] = make_int2(0, 0);
    }
    FskLoader ld = {a.state, (size_t) a.channels, (size_t) c};
    r.visit(ld);
    if (mode == 0  ||  mode == 1)
        r.restart(su.baud_rate, su.framing_mode, su.rate0, su.rate1, su.on_power, su.off_power);
    else
        r.set_frame_parameters(aux & 255, (aux >> 8) & 255, aux >> 16);
    FskStorer st = {a.state, (size_t) a.channels, (size_t) c};
    r.visit(st);
}

#endif  // SBF_NO_FSK_KERNELS

#endif  // __CUDACC__

}  // namespace sbf
