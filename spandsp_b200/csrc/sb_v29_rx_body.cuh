// sb_v29_rx_body.cuh - the V.29 receiver struct, included once per variant by sb_v29_rx.cuh with
//   SBM_V29_NAME   the struct's name
//   SBM_V29_LPC    lanes per channel (1: one thread per receiver, also the host build; 4: see RxCore)
// Training state machine, slicer, differential decoder, descrambler.  Reference: src/v29rx.c:350-785,1019-1097.
struct SBM_V29_NAME : RxCore<SBM_V29_NAME, V29_COEFF_SETS, SBM_EQ_LEN, SBM_V29_LPC>
{
    typedef V29Consts Consts;
    typedef RxCore<SBM_V29_NAME, V29_COEFF_SETS, SBM_EQ_LEN, SBM_V29_LPC> Core;

    enum
    {
        STAGE_NORMAL = 0, STAGE_SYMBOL_ACQUISITION, STAGE_LOG_PHASE, STAGE_WAIT_FOR_CDCD, STAGE_TRAIN_ON_CDCD,
        STAGE_TRAIN_ON_CDCD_AND_TEST, STAGE_TEST_ONES, STAGE_PARKED
    };
    enum
    {
        I_TRAINING_CD = I_CORE_COUNT, I_OLD_TRAIN, I_TRAIN_SCRAMBLE, I_CONSTELLATION, I_COUNT
    };
    enum
    {
        F_COUNT = F_CORE_COUNT
    };
    static const int TABLE_WORDS = ((sizeof(V29Tables) + 15)/16)*4;     // keeps what follows 16-byte aligned
    static const int LANE_WORDS = Core::CORE_LANE_WORDS;

    int training_cd;
    int old_train;
    unsigned int training_scramble_reg;
    int constellation_state;
    const V29Tables *t;

    static SB_HD void fill_tables(float *dst, const Consts &k, int lane, int nlanes)
    {
        const unsigned int *src = (const unsigned int *) k.tables;
        unsigned int *d = (unsigned int *) dst;
        for (int i = lane;  i < (int) ((sizeof(*k.tables) + 3)/4);  i += nlanes)
            d[i] = src[i];
    }

    SB_HD void bind(const float *tables, float *lane_block, int lane)
    {
        t = (const V29Tables *) tables;
        bind_core(lane_block, lane);
    }

    template <class V> SB_HD void visit(V &v)
    {
        visit_core(v);
        v.i(I_TRAINING_CD, training_cd);
        v.i(I_OLD_TRAIN, old_train);
        v.u(I_TRAIN_SCRAMBLE, training_scramble_reg);
        v.i(I_CONSTELLATION, constellation_state);
    }

    // src/v29rx.c:1019-1097
    SB_HD int restart(const Consts &k, int rate, int use_old_train)
    {
        training_cd = (rate == 9600)  ?  0  :  (rate == 7200)  ?  2  :  4;
        bit_rate = rate;
        rrc_clear();
        scramble_reg = 0;
        training_scramble_reg = 0x2A;
        training_stage = STAGE_SYMBOL_ACQUISITION;
        training_count = 0;
        signal_present = 0;
        high_sample = 0;
        low_samples = 0;
        drop_pending = 0;
        old_train = use_old_train;
        for (int i = 0;  i < 16;  i++)
            diff_angles[i*LS] = 0;
        carrier_phase = 0;
        power = 0;                                  // power_meter_init(&s->power, 4)
        constellation_state = 0;
        eq_delta = k.eq_delta;
        if (use_old_train)
        {
            phase_rate = phase_rate_save;
            equalizer_restore();
            agc_scaling = agc_scaling_save;
        }
        else
        {
            phase_rate = k.rate_nominal;
            equalizer_reset();
            agc_scaling_save = 0.0f;
            agc_scaling = k.agc_initial;
        }
        track_i = 8000.0f;
        track_p = 8000000.0f;
        last_sample = 0;
        eq_skip = 0;
        godard_init();
        baud_half = 0;
        return 0;
    }

    // v29_rx_init (src/v29rx.c:1100-1134): memset, signal cutoff, restart
    SB_HD void init(const Consts &k, int rate, int on_pw, int off_pw)
    {
        training_error = 0.0f;
        last_angle0 = last_angle1 = 0;
        eq_step = 0;
        eq_put_step = 0;
        on_power = on_pw;
        off_power = off_pw;
        phase_rate_save = 0;
        agc_scaling_save = 0.0f;
        restart(k, rate, 0);
    }

    SB_HD void restart_after_carrier_down(const Consts &k)
    {
        restart(k, bit_rate, 0);                // src/v29rx.c:836
    }

    // src/v29rx.c:350-361
    SB_HD int scrambled_training_bit()
    {
        const int bit = training_scramble_reg & 1;
        training_scramble_reg >>= 1;
        if (bit ^ (int) (training_scramble_reg & 1))
            training_scramble_reg |= 0x40;
        return bit;
    }

    // src/v29rx.c:365-396
    SB_HD void put_bit(int bit)
    {
        bit &= 1;
        const int out = (bit ^ (int) (scramble_reg >> (18 - 1)) ^ (int) (scramble_reg >> (23 - 1))) & 1;
        scramble_reg = (scramble_reg << 1) | (unsigned int) bit;
        if (training_stage == STAGE_NORMAL)
            out_bit(out);
    }

    // n (2..4) consecutive put_bit() calls at once, first bit in bit 0.  The descrambler's taps are 18 and 23 bits
    // back, so the bits of one baud do not feed each other: bit i is b[i] ^ reg[17 - i] ^ reg[22 - i] of the register
    // as it stood before the baud.
    SB_HD static unsigned int rev4(unsigned int x)
    {
#if defined(__CUDA_ARCH__)
        return __brev(x) >> 28;
#else
        return ((x & 1u) << 3) | ((x & 2u) << 1) | ((x & 4u) >> 1) | ((x & 8u) >> 3);
#endif
    }

    SB_HD void put_bits(unsigned int b, int n)
    {
        const unsigned int mask = (1u << n) - 1u;
        b &= mask;
        const unsigned int taps = ((scramble_reg >> 14) ^ (scramble_reg >> 19)) & 0xFu;
        const unsigned int out = (b ^ rev4(taps)) & mask;
        scramble_reg = (scramble_reg << n) | (rev4(b) >> (4 - n));
        if (training_stage == STAGE_NORMAL)
            out_bits(out, n);
    }

    // src/v29rx.c:402-480
    SB_HD void decode_baud(float zre, float zim)
    {
        int nearest;
        int raw_bits;

        if (bit_rate == 4800)
        {
            const int b1 = (zim > zre);
            const int b2 = (zim < -zre);
            nearest = ((b2 << 1) | (b1 ^ b2)) << 1;
            raw_bits = t->phase_steps_4800[((nearest - constellation_state) >> 1) & 3];
            put_bits((unsigned int) raw_bits, 2);
        }
        else
        {
            int re = f2i(fmul(fadd(zre, 5.0f), 2.0f));
            int im = f2i(fmul(fadd(zim, 5.0f), 2.0f));
            re = (re > 19)  ?  19  :  (re < 0)  ?  0  :  re;
            im = (im > 19)  ?  19  :  (im < 0)  ?  0  :  im;
            nearest = t->space_map[re][im];
            const unsigned int amp_bit = (unsigned int) (nearest >> 3) & 1u;
            if (bit_rate != 9600)
                nearest &= 7;
            raw_bits = t->phase_steps_9600[(nearest - constellation_state) & 7];
            if (bit_rate == 9600)
                put_bits(amp_bit | ((unsigned int) (raw_bits & 7) << 1), 4);
            else
                put_bits((unsigned int) raw_bits, 3);
        }
        const float tre = t->constellation[nearest][0];
        const float tim = t->constellation[nearest][1];
        track_carrier(zre, zim, tre, tim);
        if (--eq_skip <= 0)
        {
            eq_skip = 10;
            tune_equalizer(zre, zim, tre, tim);
        }
        constellation_state = nearest;
    }

    SB_HD void park()
    {
        agc_scaling_save = 0.0f;
        training_stage = STAGE_PARKED;
        report_status(SIG_STATUS_TRAINING_FAILED);
    }

    // src/v29rx.c:526-785: the once-per-baud part of process_half_baud()
    SB_HD void process_baud(const Consts &k)
    {
        eq_put_step += godard_per_baud(k);
        float zre;
        float zim;
        if (Core::LANES > 1)
        {
            zre = h_zre;                    // computed by run_group() with the warp converged
            zim = h_zim;
        }
        else
        {
            equalizer_get(zre, zim);
        }
        float tre = 0.0f;
        float tim = 0.0f;

        switch (training_stage)
        {
        case STAGE_NORMAL:
            decode_baud(zre, zim);
            tre = t->constellation[constellation_state][0];
            tim = t->constellation[constellation_state][1];
            break;
        case STAGE_SYMBOL_ACQUISITION:
            if (++training_count >= 60)
            {
                training_stage = STAGE_LOG_PHASE;
                for (int i = 0;  i < 16;  i++)
                    diff_angles[i*LS] = 0;
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
                diff_angles[(i & 0xF)*LS] = diff_angles[((i - 2) & 0xF)*LS] + (ang >> 4);
                if ((ang > k.phase_p45  ||  ang < k.phase_m45)  &&  training_count >= 13)
                {
                    i = (training_count - 8) & ~1;
                    if (i > 1)
                    {
                        const int j = i & 0xF;
                        ang = (diff_angles[j*LS] + diff_angles[(j | 0x1)*LS])/(i - 1);
                        phase_rate += 3*16*(ang/20);
                    }
                    if (phase_rate < k.rate_low  ||  phase_rate > k.rate_high)
                    {
                        park();
                        break;
                    }
                    spin_equalizer_buffer((unsigned int) angle);
                    carrier_phase += (unsigned int) angle;
                    const int bit = scrambled_training_bit();
                    constellation_state = t->cdcd_pos[training_cd + bit];
                    tre = t->constellation[constellation_state][0];
                    tim = t->constellation[constellation_state][1];
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
                constellation_state = t->cdcd_pos[training_cd + bit];
                tre = t->constellation[constellation_state][0];
                tim = t->constellation[constellation_state][1];
                track_carrier(zre, zim, tre, tim);
                tune_equalizer(zre, zim, tre, tim);
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
                constellation_state = t->cdcd_pos[training_cd + bit];
                tre = t->constellation[constellation_state][0];
                tim = t->constellation[constellation_state][1];
                track_carrier(zre, zim, tre, tim);
                tune_equalizer(zre, zim, tre, tim);
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
                decode_baud(zre, zim);
                tre = t->constellation[constellation_state][0];
                tim = t->constellation[constellation_state][1];
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
        report_symbol(zre, zim, tre, tim, constellation_state);
    }
};
