// sb_v27ter_rx.cuh - the V.27ter receiver (4800 bit/s at 1600 baud, 2400 bit/s at 1200 baud) on top of the
// shared modem core: RRC band-pass FIR pair evaluated at the T/2 instants only, Gardner symbol timing,
// 32-tap T/2 complex equalizer, DPSK slicer (quadrant / octant), descrambler with the V.27ter repeating
// pattern guard, training state machine.  Reference: src/v27ter_rx.c:137-1210.
#pragma once

#include "sb_modem.cuh"

namespace sbm {

#define V27TER_EQ_LEN               32
#define V27TER_EQ_PRE_LEN           16
#define V27TER_SETS_4800            8
#define V27TER_SETS_2400            12
#define V27TER_TRAINING_SEG_3_LEN   50
#define V27TER_TRAINING_SEG_5_LEN   1074
#define V27TER_TRAINING_SEG_6_LEN   8

struct V27terConsts : CoreConsts
{
    int phase_p45;                      // DDS_PHASE(45.0f)
    int phase_m45;                      // DDS_PHASE(-45.0f)
    int phase_180;                      // DDS_PHASE(180.0f)
    float eq_delta;                     // 0.25f/32
};

// [20][27]: rows 0..7 = the 4800 bit/s sets (1600 baud), rows 8..19 = the 2400 bit/s sets (1200 baud)
// (src/make_modem_filter.c:375-400)
static inline void make_v27ter_rrc(std::vector<float> &re, std::vector<float> &im)
{
    std::vector<float> re24;
    std::vector<float> im24;
    make_rx_rrc(re, im, V27TER_SETS_4800, 1800.0, 1600.0);
    make_rx_rrc(re24, im24, V27TER_SETS_2400, 1800.0, 1200.0);
    re.insert(re.end(), re24.begin(), re24.end());
    im.insert(im.end(), im24.begin(), im24.end());
}

// The coefficient rows of both rates live in one table: rows 0..7 = 4800 bit/s, rows 8..19 = 2400 bit/s.
struct RxV27ter : RxCore<RxV27ter, V27TER_SETS_4800 + V27TER_SETS_2400, V27TER_EQ_LEN>
{
    typedef V27terConsts Consts;
    typedef RxCore<RxV27ter, V27TER_SETS_4800 + V27TER_SETS_2400, V27TER_EQ_LEN> Core;

    enum
    {
        STAGE_NORMAL = 0, STAGE_SYMBOL_ACQUISITION, STAGE_LOG_PHASE, STAGE_WAIT_FOR_HOP, STAGE_TRAIN_ON_ABAB,
        STAGE_TEST_ONES, STAGE_PARKED
    };
    enum
    {
        I_PATTERN_COUNT = I_CORE_COUNT, I_TRAINING_BC, I_CONSTELLATION, I_GARDNER_INTEGRATE, I_GARDNER_STEP, I_COUNT
    };
    enum
    {
        F_COUNT = F_CORE_COUNT
    };
    static const int TABLE_WORDS = 0;
    static const int LANE_WORDS = Core::CORE_LANE_WORDS;

    int scrambler_pattern_count;
    int training_bc;
    int constellation_state;
    int gardner_integrate;
    int gardner_step;

    static SB_HD void fill_tables(float *, const Consts &, int, int)
    {
    }

    SB_HD void bind(const float *, float *lane_block, int lane)
    {
        bind_core(lane_block, lane);
    }

    template <class V> SB_HD void visit(V &v)
    {
        visit_core(v);
        v.i(I_PATTERN_COUNT, scrambler_pattern_count);
        v.i(I_TRAINING_BC, training_bc);
        v.i(I_CONSTELLATION, constellation_state);
        v.i(I_GARDNER_INTEGRATE, gardner_integrate);
        v.i(I_GARDNER_STEP, gardner_step);
    }

    static int fillin_sets(int rate) { return (rate == 4800)  ?  V27TER_SETS_4800  :  V27TER_SETS_2400; }
    static int fillin_half_baud(int rate) { return (rate == 4800)  ?  V27TER_SETS_4800*5/2  :  V27TER_SETS_2400*20/(3*2); }
    SB_HD int sets() const { return (bit_rate == 4800)  ?  V27TER_SETS_4800  :  V27TER_SETS_2400; }
    // T/2 in coefficient-set steps: RX_PULSESHAPER_4800_COEFF_SETS*5/2, RX_PULSESHAPER_2400_COEFF_SETS*20/(3*2)
    SB_HD int half_baud_steps() const { return (bit_rate == 4800)  ?  V27TER_SETS_4800*5/2  :  V27TER_SETS_2400*20/(3*2); }

    // v27ter_constellation[8] (src/v27ter_rx.c:121-135)
    static SB_HD float con_re(int i)
    {
        return (i == 0)  ?  1.414f  :  (i == 1  ||  i == 7)  ?  1.0f  :  (i == 2  ||  i == 6)  ?  0.0f  :  (i == 4)  ?  -1.414f  :  -1.0f;
    }
    static SB_HD float con_im(int i)
    {
        return (i == 2)  ?  1.414f  :  (i == 1  ||  i == 3)  ?  1.0f  :  (i == 0  ||  i == 4)  ?  0.0f  :  (i == 6)  ?  -1.414f  :  -1.0f;
    }

    // src/v27ter_rx.c:219-241
    SB_HD void equalizer_reset27(const Consts &k)
    {
        for (int i = 0;  i < V27TER_EQ_LEN;  i++)
            eq_coeff[i*32] = make_float2(0.0f, 0.0f);
        for (int i = 0;  i < 2*V27TER_EQ_LEN;  i++)
            eq_buf[i*32] = make_float2(0.0f, 0.0f);
        eq_coeff[(V27TER_EQ_PRE_LEN + 1)*32] = make_float2(1.414f, 0.0f);
        eq_delta = k.eq_delta;
        eq_put_step = half_baud_steps();
        eq_step = 0;
    }

    // v27ter_rx_restart (src/v27ter_rx.c:1091-1158).  The reference never stores its old_train argument (the
    // test at :1132 reads a field that stays false after v27ter_rx_init's memset), so every restart is a full
    // retrain; the argument is accepted and ignored here in the same way.
    SB_HD int restart(const Consts &k, int rate, int)
    {
        if (rate != 4800  &&  rate != 2400)
            return -1;
        bit_rate = rate;
        rrc_clear();
        training_error = 0.0f;
        scramble_reg = 0x3C;
        scrambler_pattern_count = 0;
        training_stage = STAGE_SYMBOL_ACQUISITION;
        training_bc = 0;
        training_count = 0;
        signal_present = 0;
        high_sample = 0;
        low_samples = 0;
        drop_pending = 0;
        for (int i = 0;  i < 16;  i++)
            diff_angles[i*32] = 0;
        carrier_phase = 0;
        track_i = 200000.0f;
        track_p = 10000000.0f;
        power = 0;                                  // power_meter_init(&s->power, 4)
        constellation_state = 0;
        phase_rate = k.rate_nominal;
        agc_scaling = k.agc_initial;
        equalizer_reset27(k);
        eq_skip = 0;
        last_sample = 0;
        gardner_integrate = 0;
        total_timing = 0;
        gardner_step = 512;
        baud_half = 0;
        return 0;
    }

    // v27ter_rx_init (src/v27ter_rx.c:1161-1188): memset, signal cutoff, restart
    SB_HD void init(const Consts &k, int rate, int on_pw, int off_pw)
    {
        agc_scaling_save = 0.0f;
        last_angle0 = last_angle1 = 0;
        phase_rate_save = 0;
        godard_init();
        on_power = on_pw;
        off_power = off_pw;
        restart(k, rate, 0);
    }

    SB_HD void restart_after_carrier_down(const Consts &k)
    {
        restart(k, bit_rate, 0);                // src/v27ter_rx.c:840
    }

    // src/v27ter_rx.c:377-412
    SB_HD int descramble(int in_bit)
    {
        in_bit &= 1;
        int out_bit = (in_bit ^ (int) (scramble_reg >> 5) ^ (int) (scramble_reg >> 6)) & 1;
        const bool training = (training_stage > STAGE_NORMAL  &&  training_stage < STAGE_TEST_ONES);
        if (scrambler_pattern_count >= 33)
        {
            out_bit ^= 1;
            scrambler_pattern_count = 0;
        }
        else if (training)
        {
            scrambler_pattern_count = 0;
        }
        else
        {
            if ((((int) (scramble_reg >> 7) ^ in_bit) & ((int) (scramble_reg >> 8) ^ in_bit) & ((int) (scramble_reg >> 11) ^ in_bit) & 1))
                scrambler_pattern_count = 0;
            else
                scrambler_pattern_count++;
        }
        scramble_reg <<= 1;
        scramble_reg |= (unsigned int) ((training)  ?  out_bit  :  in_bit);
        return out_bit;
    }

    // src/v27ter_rx.c:415-435
    SB_HD void put_bit(int bit)
    {
        const int out = descramble(bit);
        if (training_stage == STAGE_NORMAL)
            out_bit(out);
    }

    // src/v27ter_rx.c:438-483 with find_quadrant (:318-331) and find_octant (:334-374)
    SB_HD void decode_baud(float zre, float zim)
    {
        int nearest;
        if (bit_rate == 2400)
        {
            const int b1 = (zim > zre);
            const int b2 = (zim < -zre);
            nearest = (b2 << 1) | (b1 ^ b2);
            // phase_steps_2400[4] = {0, 2, 3, 1}
            const int raw_bits = (0x78 >> (2*((nearest - constellation_state) & 3))) & 3;
            put_bit(raw_bits);
            put_bit(raw_bits >> 1);
            constellation_state = nearest;
            nearest <<= 1;
        }
        else
        {
            const float abs_re = fabsf(zre);
            const float abs_im = fabsf(zim);
            if (fmul(abs_im, 1.0f) > fmul(abs_re, 0.4142136f)  &&  fmul(abs_im, 1.0f) < fmul(abs_re, 2.4142136f))
            {
                const int b1 = (zre < 0.0f);
                const int b2 = (zim < 0.0f);
                nearest = (b2 << 2) | ((b1 ^ b2) << 1) | 1;
            }
            else
            {
                const int b1 = (zim > zre);
                const int b2 = (zim < -zre);
                nearest = (b2 << 2) | ((b1 ^ b2) << 1);
            }
            // phase_steps_4800[8] = {4, 0, 2, 6, 7, 3, 1, 5}
            const int raw_bits = (int) ((0x51376204u >> (4*((nearest - constellation_state) & 7))) & 7u);
            put_bit(raw_bits);
            put_bit(raw_bits >> 1);
            put_bit(raw_bits >> 2);
            constellation_state = nearest;
        }
        const float tre = con_re(nearest);
        const float tim = con_im(nearest);
        track_carrier(zre, zim, tre, tim);
        if (--eq_skip <= 0)
        {
            eq_skip = 100;
            tune_equalizer(zre, zim, tre, tim);
        }
    }

    // The Gardner test for baud alignment (src/v27ter_rx.c:486-525).  A hop is reported through qam_report
    // with NULL pointers and the integrator value; the record carries NaN coordinates for that.
    SB_HD void symbol_sync()
    {
        const float2 a = eq_buf[((eq_step - 3) & (V27TER_EQ_LEN - 1))*32];
        const float2 b = eq_buf[((eq_step - 1) & (V27TER_EQ_LEN - 1))*32];
        const float2 m = eq_buf[((eq_step - 2) & (V27TER_EQ_LEN - 1))*32];
        const float p = fmul(fsub(a.x, b.x), m.x);
        const float q = fmul(fsub(a.y, b.y), m.y);
        gardner_integrate += (fadd(p, q) > 0.0f)  ?  gardner_step  :  -gardner_step;
        if (abs(gardner_integrate) >= 128)
        {
            eq_put_step += gardner_integrate/128;
            total_timing += gardner_integrate/128;
            const float nan = __int_as_float_hd(0x7FC00000);
            report_symbol(nan, nan, nan, nan, gardner_integrate);
            gardner_integrate = 0;
        }
    }

    static SB_HD float __int_as_float_hd(int v)
    {
#if defined(__CUDA_ARCH__)
        return __int_as_float(v);
#else
        float f;
        memcpy(&f, &v, 4);
        return f;
#endif
    }

    SB_HD void park()
    {
        training_stage = STAGE_PARKED;
        report_status(SIG_STATUS_TRAINING_FAILED);
    }

    SB_HD void next_abab(float &tre, float &tim)
    {
        training_bc ^= descramble(1);
        descramble(1);
        descramble(1);
        constellation_state = (training_bc)  ?  4  :  0;         // abab_pos[2] = {0, 4}
        tre = con_re(constellation_state);
        tim = con_im(constellation_state);
    }

    // src/v27ter_rx.c:560-781: the once-per-baud part of process_half_baud()
    SB_HD void process_baud(const Consts &k)
    {
        symbol_sync();
        float zre;
        float zim;
        equalizer_get(zre, zim);
        float tre = 0.0f;
        float tim = 0.0f;

        switch (training_stage)
        {
        case STAGE_NORMAL:
            {
                decode_baud(zre, zim);
                const int cs = (bit_rate == 4800)  ?  constellation_state  :  (constellation_state << 1);
                tre = con_re(cs);
                tim = con_im(cs);
            }
            break;
        case STAGE_SYMBOL_ACQUISITION:
            if (++training_count >= 30)
            {
                gardner_step = 32;
                training_stage = STAGE_LOG_PHASE;
                for (int i = 0;  i < 16;  i++)
                    diff_angles[i*32] = 0;
                last_angle0 = arctan2(zim, zre);
            }
            break;
        case STAGE_LOG_PHASE:
            last_angle1 = arctan2(zim, zre);
            training_count = 1;
            training_stage = STAGE_WAIT_FOR_HOP;
            break;
        case STAGE_WAIT_FOR_HOP:
            {
                int angle = arctan2(zim, zre);
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
                        if (bit_rate == 4800)
                            phase_rate += 16*(ang/10);
                        else
                            phase_rate += 3*16*(ang/40);
                    }
                    if (phase_rate < k.rate_low  ||  phase_rate > k.rate_high)
                    {
                        park();
                        break;
                    }
                    angle = (int) ((unsigned int) angle + (unsigned int) k.phase_180);
                    spin_equalizer_buffer((unsigned int) angle);
                    carrier_phase += (unsigned int) angle;
                    gardner_step = 2;
                    // The first element of the scrambled sequence has just been seen, so skip it
                    training_bc = 1;
                    next_abab(tre, tim);
                    training_count = 1;
                    training_stage = STAGE_TRAIN_ON_ABAB;
                    report_status(SIG_STATUS_TRAINING_IN_PROGRESS);
                }
                else if (++training_count > V27TER_TRAINING_SEG_3_LEN)
                {
                    park();
                }
            }
            break;
        case STAGE_TRAIN_ON_ABAB:
            {
                next_abab(tre, tim);
                track_carrier(zre, zim, tre, tim);
                tune_equalizer(zre, zim, tre, tim);
                const float left = (float) (V27TER_TRAINING_SEG_5_LEN - training_count);
                track_i = fadd(400.0f, fdiv(fmul(fsub(200000.0f, 400.0f), left), (float) V27TER_TRAINING_SEG_5_LEN));
                track_p = fadd(1000000.0f, fdiv(fmul(fsub(10000000.0f, 1000000.0f), left), (float) V27TER_TRAINING_SEG_5_LEN));
                if (++training_count >= V27TER_TRAINING_SEG_5_LEN)
                {
                    constellation_state = (bit_rate == 4800)  ?  4  :  2;
                    training_count = 0;
                    training_stage = STAGE_TEST_ONES;
                }
            }
            break;
        case STAGE_TEST_ONES:
            {
                decode_baud(zre, zim);
                const int cs = (bit_rate == 4800)  ?  constellation_state  :  (constellation_state << 1);
                tre = con_re(cs);
                tim = con_im(cs);
                const float dr = fsub(zre, tre);
                const float di = fsub(zim, tim);
                training_error = fadd(training_error, fadd(fmul(dr, dr), fmul(di, di)));
                if (++training_count >= V27TER_TRAINING_SEG_6_LEN)
                {
                    if ((bit_rate == 4800  &&  training_error < fmul((float) V27TER_TRAINING_SEG_6_LEN, 0.25f))
                        ||
                        (bit_rate == 2400  &&  training_error < fmul((float) V27TER_TRAINING_SEG_6_LEN, 0.5f)))
                    {
                        report_status(SIG_STATUS_TRAINING_SUCCEEDED);
                        signal_present = (bit_rate == 4800)  ?  90  :  120;
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
        report_symbol(zre, zim, tre, tim, constellation_state);
    }

    // v27ter_rx()'s per-sample body (src/v27ter_rx.c:882-944,950-1012), split like the other receivers:
    //   front27(): ring insert, carrier detect, symbol clock; tells whether this sample is a T/2 instant;
    //   half27():  AGC, the FIR pair, down-mix, equalizer insert; tells whether a whole baud is complete;
    //   baud():    Gardner timing, equalizer, training state machine / slicer, qam report.
    template <class K> SB_HD bool front27(const K &k, short amp)
    {
        const float xv = (float) amp;
        rrc[rrc_step*32] = xv;
        rrc[(rrc_step + SBM_FILTER_STEPS)*32] = xv;
        if (++rrc_step >= SBM_FILTER_STEPS)
            rrc_step = 0;
        const int pw = signal_detect(k, amp);
        if (pw == 0)
            return false;
        if (training_stage == STAGE_PARKED)
            return false;
        if ((eq_put_step -= sets()) <= 0)
        {
            h_pw = pw;
            return true;
        }
        carrier_phase += (unsigned int) phase_rate;
        return false;
    }

    template <class K> SB_HD bool half27(const K &k, const float *s_rrc_re, const float *s_rrc_im)
    {
        if (training_stage == STAGE_SYMBOL_ACQUISITION)
        {
            int root_power = fixed_sqrt32(k, (unsigned int) h_pw);
            if (root_power == 0)
                root_power = 1;
            agc_scaling = fdiv(k.agc_target, (float) root_power);
        }
        int step = -eq_put_step;
        if (step > sets() - 1)
            step = sets() - 1;
        const int row = ((bit_rate == 4800)  ?  0  :  V27TER_SETS_4800) + step;
        float v_re;
        float v_im;
        rrc_dot2(s_rrc_re + row*SBM_RRC_ROW, s_rrc_im + row*SBM_RRC_ROW, v_re, v_im);
        const float sre = fmul(v_re, agc_scaling);
        const float sim = fmul(v_im, agc_scaling);
        const float zr = sine[(carrier_phase + (1u << 30)) >> 21];       // dds_lookup_complexf, src/dds_float.c:2177
        const float zi = sine[carrier_phase >> 21];
        const float zzre = fsub(fmul(sre, zr), fmul(sim, zi));
        const float zzim = fsub(fmul(-sre, zi), fmul(sim, zr));
        eq_put_step += half_baud_steps();
        // process_half_baud, first part (src/v27ter_rx.c:549-558)
        eq_buf[eq_step*32] = make_float2(zzre, zzim);
        eq_buf[(eq_step + V27TER_EQ_LEN)*32] = make_float2(zzre, zzim);
        if (++eq_step >= V27TER_EQ_LEN)
            eq_step = 0;
        if ((baud_half ^= 1))
        {
            carrier_phase += (unsigned int) phase_rate;
            return false;
        }
        return true;
    }

    // One trip = one baud: two T/2 instants, each reached after two to four input samples (2.5 per T/2 at 1600
    // baud, 3.33 at 1200 baud); lanes stay converged in the FIR pair and the per-baud work.
    template <class K> SB_HD void run(const K &k, const float *s_rrc_re, const float *s_rrc_im, const int16_t *row, int n)
    {
        int pos = 0;
        feed_open(row, n);
#pragma unroll 1
        while (pos < n)
        {
#pragma unroll 1
            for (int h = 0;  h < 2;  h++)
            {
                if (h == 0  &&  baud_half)
                    continue;
                unsigned long long cur = feed_peek4(pos);
                bool due = false;
#pragma unroll 1
                for (int q = 0;  q < 4;  q++)
                {
                    if (pos < n  &&  !due)
                    {
                        const short amp = (short) (cur & 0xFFFFu);
                        cur >>= 16;
                        due = front27(k, amp);
                        pos++;
                    }
                }
                if (due)
                {
                    if (half27(k, s_rrc_re, s_rrc_im))
                        baud(k);
                }
            }
        }
    }
};

}  // namespace sbm
