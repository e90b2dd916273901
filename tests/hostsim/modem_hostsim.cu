// modem_hostsim.cu - TEST INFRASTRUCTURE ONLY.  Compiles the receivers of spandsp_b200/csrc/sb_v29_rx.cuh and
// sb_v17_rx.cuh (the very code the CUDA kernels run, written __host__ __device__) for the HOST, one channel
// at a time, so that the training state machines, the trellis decoder and the state save/restore can be
// compared with the oracle in a container without a GPU (tests/test_hostsim_modem.py).  Nothing in the
// product library links or loads this; the product path exists on the GPU only.
#include "../../spandsp_b200/csrc/sb_modem.cuh"
#include "../../spandsp_b200/csrc/sb_v29_rx.cuh"
#include "../../spandsp_b200/csrc/sb_v17_rx.cuh"
#include "../../spandsp_b200/csrc/sb_v27ter_rx.cuh"
#include "../../spandsp_b200/csrc/sb_fsk_rx.cuh"
#include "../../spandsp_b200/csrc/sb_mct_rx.cuh"
#include "../../spandsp_b200/csrc/sb_gen.cuh"
#include "../../spandsp_b200/csrc/sb_sig_rx.cuh"

using namespace sbm;

#define EXPORT extern "C" __attribute__((visibility("default")))

struct HostTables
{
    std::vector<float> re, im, sine;
    std::vector<unsigned short> sq;
    V29Tables t29;
    V17Tables t17;
    std::vector<unsigned char> maps, map4800;
};

static void core_consts(CoreConsts &k, HostTables &h, int sets, double carrier, double fine, int coarse, float agc_target)
{
    godard_desc_t g;
    make_rx_rrc(h.re, h.im, sets, carrier);
    make_sine_table(h.sine);
    make_sqrt_table(h.sq);
    make_godard(g, carrier, fine, coarse);
    k.rrc_re = h.re.data();
    k.rrc_im = h.im.data();
    k.sine = h.sine.data();
    k.sqrt_tab = h.sq.data();
    for (int i = 0;  i < 3;  i++)
    {
        k.g_low[i] = g.low[i];
        k.g_high[i] = g.high[i];
    }
    k.g_mixed3 = g.mixed3;
    k.g_coarse_trigger = g.coarse_trigger;
    k.g_fine_trigger = g.fine_trigger;
    k.g_coarse_step = g.coarse_step;
    k.g_fine_step = g.fine_step;
    k.rate_nominal = host_dds_phase_rate((float) carrier);
    k.rate_low = host_dds_phase_rate((float) carrier - 20.0f);
    k.rate_high = host_dds_phase_rate((float) carrier + 20.0f);
    k.agc_target = agc_target/1.000000f;
    k.agc_initial = (agc_target/1.000000f)/735.0f;
}

template <class RX>
struct HostChannel
{
    typename RX::Consts k;
    std::vector<float> smem;
    std::vector<float> fstate;
    std::vector<int> istate;
    const float *s_re;
    const float *s_im;
    float *tables;
    float *lane;

    void setup()
    {
        smem.assign(2*RX::SETS*SBM_RRC_ROW + RX::TABLE_WORDS + RX::LANE_WORDS*32, 0.0f);
        fstate.assign(RX::F_COUNT, 0.0f);
        istate.assign(RX::I_COUNT, 0);
        for (int row = 0;  row < RX::SETS;  row++)
        {
            for (int tap = 0;  tap < SBM_FILTER_STEPS;  tap++)
            {
                smem[row*SBM_RRC_ROW + tap] = k.rrc_re[row*SBM_FILTER_STEPS + tap];
                smem[RX::SETS*SBM_RRC_ROW + row*SBM_RRC_ROW + tap] = k.rrc_im[row*SBM_FILTER_STEPS + tap];
            }
        }
        s_re = smem.data();
        s_im = smem.data() + RX::SETS*SBM_RRC_ROW;
        tables = smem.data() + 2*RX::SETS*SBM_RRC_ROW;
        lane = tables + RX::TABLE_WORDS;
        RX::fill_tables(tables, k, 0, 1);
    }

    void attach(RX &r)
    {
        r.c = 0;
        r.channels = 1;
        r.fstate = fstate.data();
        r.in_ring = NULL;
        r.sine = k.sine;
        r.sqrt_tab = k.sqrt_tab;
        r.bind(tables, lane, 0);
        r.bits = NULL;
        r.bits_cap = 0;
        r.nbits = 0;
        r.syms = NULL;
        r.sym_cap = 0;
        r.nsyms = 0;
    }

    void init(int bit_rate, int on_power, int off_power)
    {
        RX r;
        memset(&r, 0, sizeof(r));
        attach(r);
        r.init(k, bit_rate, on_power, off_power);
        StateStorer st = {fstate.data(), istate.data(), 1, 0};
        r.visit(st);
    }

    int restart(int bit_rate, int mode)
    {
        RX r;
        memset(&r, 0, sizeof(r));
        attach(r);
        StateLoader ld = {fstate.data(), istate.data(), 1, 0};
        r.visit(ld);
        r.mirror_rings();
        const int rc = r.restart(k, bit_rate, mode);
        StateStorer st = {fstate.data(), istate.data(), 1, 0};
        r.visit(st);
        return rc;
    }

    // one rx call: load state, run, store state - exactly what modem_rx_kernel does per launch
    void rx(const int16_t *amp, int n, int8_t *bits, int bits_cap, int *nbits, span_b200_v29_symbol_t *syms, int sym_cap, int *nsyms)
    {
        RX r;
        memset(&r, 0, sizeof(r));
        attach(r);
        StateLoader ld = {fstate.data(), istate.data(), 1, 0};
        r.visit(ld);
        r.mirror_rings();
        r.bits = (signed char *) bits;
        r.bits_cap = bits_cap;
        r.syms = syms;
        r.sym_cap = sym_cap;
        r.run(k, s_re, s_im, amp, n);
        StateStorer st = {fstate.data(), istate.data(), 1, 0};
        r.visit(st);
        *nbits = r.nbits;
        *nsyms = r.nsyms;
    }
};

template <class RX>
static int run_channel(HostChannel<RX> &ch, const int16_t *amp, int n, int chunk, int bit_rate, int restart_at, int restart_mode,
                       int8_t *bits, int bits_cap, int32_t *nbits, span_b200_v29_symbol_t *syms, int sym_cap, int32_t *nsyms,
                       float *eq_coeff)
{
    int tb = 0;
    int ts = 0;
    int len;
    if (chunk <= 0)
        chunk = n;
    for (int pos = 0;  pos < n;  pos += len)
    {
        if (restart_at >= 0  &&  pos >= restart_at)
        {
            ch.restart(bit_rate, restart_mode);
            restart_at = -1;
        }
        len = (n - pos < chunk)  ?  (n - pos)  :  chunk;
        int nb = 0;
        int ns = 0;
        // bit positions in the symbol records count from the start of each rx call; make them global
        span_b200_v29_symbol_t *sp = (syms  &&  ts < sym_cap)  ?  (syms + ts)  :  NULL;
        ch.rx(amp + pos, len, (tb < bits_cap)  ?  (bits + tb)  :  bits, (tb < bits_cap)  ?  (bits_cap - tb)  :  0, &nb,
              sp, (sp)  ?  (sym_cap - ts)  :  0, &ns);
        if (sp)
        {
            for (int i = 0;  i < ns  &&  ts + i < sym_cap;  i++)
                sp[i].bit_pos += tb;
        }
        tb += nb;
        ts += ns;
    }
    *nbits = tb;
    *nsyms = ts;
    if (eq_coeff)
    {
        for (int i = 0;  i < 2*SBM_EQ_LEN;  i++)
            eq_coeff[i] = ch.fstate[F_EQ_COEFF + i];
    }
    return 0;
}

// final[]: as oracle ref_v17_run: {training_stage, carrier_phase_rate, eq_put_step, signal_present, agc_scaling bits,
//           total_baud_timing_correction, diff, carrier_phase, short_train, trellis_ptr}
EXPORT int hostsim_v17_run(const int16_t *amp, int n, int chunk, int bit_rate, float cutoff, int restart_at, int restart_short,
                           int8_t *bits, int bits_cap, int32_t *nbits, span_b200_v29_symbol_t *syms, int sym_cap, int32_t *nsyms,
                           float *eq_coeff, int32_t *final)
{
    static HostTables h;
    HostChannel<RxV17> ch;
    core_consts(ch.k, h, V17_COEFF_SETS, 1800.0, 100.0, 15, 2.17f);
    make_v17_tables(h.t17);
    make_v17_maps(h.t17, h.maps, h.map4800);
    ch.k.tables = &h.t17;
    ch.k.maps = h.maps.data();
    ch.k.map4800 = h.map4800.data();
    ch.k.phase_p90 = host_dds_phase(90.0f);
    ch.k.phase_m90 = host_dds_phase(-90.0f);
    ch.k.phase_180 = host_dds_phase(180.0f);
    ch.k.phase_a = host_dds_phase(270.0f + 18.433f);
    ch.k.phase_b = host_dds_phase(180.0f + 18.433f);
    ch.k.phase_c = host_dds_phase(18.433f);
    const float fast = 0.21f/SBM_EQ_LEN;
    ch.k.eq_delta_fast = fast;
    ch.k.eq_delta_slow = 0.1f*fast;
    ch.setup();
    if (cutoff <= -99.0f)
        cutoff = -45.5f;
    ch.init(bit_rate, (int32_t) (host_power_meter_level_dbm0(cutoff + 2.5f)*0.4f), (int32_t) (host_power_meter_level_dbm0(cutoff - 2.5f)*0.4f));
    run_channel(ch, amp, n, chunk, bit_rate, restart_at, restart_short, bits, bits_cap, nbits, syms, sym_cap, nsyms, eq_coeff);
    if (final)
    {
        final[0] = ch.istate[I_STAGE];
        final[1] = ch.istate[I_PHASE_RATE];
        final[2] = ch.istate[I_EQ_PUT_STEP];
        final[3] = ch.istate[I_SIGNAL_PRESENT];
        memcpy(&final[4], &ch.fstate[F_AGC], 4);
        final[5] = ch.istate[I_TOTAL_TIMING];
        final[6] = ch.istate[RxV17::I_DIFF];
        final[7] = ch.istate[I_CARRIER_PHASE];
        final[8] = ch.istate[RxV17::I_SHORT_TRAIN];
        final[9] = ch.istate[RxV17::I_TRELLIS_PTR];
    }
    return 0;
}

// final[]: as oracle ref_v29_run: {training_stage, carrier_phase_rate, eq_put_step, signal_present, agc_scaling bits,
//           total_baud_timing_correction, constellation_state, carrier_phase}
EXPORT int hostsim_v29_run(const int16_t *amp, int n, int chunk, int bit_rate, float cutoff, int restart_at, int restart_old_train,
                           int8_t *bits, int bits_cap, int32_t *nbits, span_b200_v29_symbol_t *syms, int sym_cap, int32_t *nsyms,
                           float *eq_coeff, int32_t *final)
{
    static HostTables h;
    HostChannel<RxV29> ch;
    core_consts(ch.k, h, V29_COEFF_SETS, 1700.0, 30.0, 5, 1.25f);
    make_v29_tables(h.t29);
    ch.k.tables = &h.t29;
    ch.k.phase_p45 = host_dds_phase(45.0f);
    ch.k.phase_m45 = host_dds_phase(-45.0f);
    ch.k.eq_delta = 0.21f/SBM_EQ_LEN;
    ch.setup();
    if (cutoff <= -99.0f)
        cutoff = -28.5f;
    ch.init(bit_rate, (int32_t) (host_power_meter_level_dbm0(cutoff + 2.5f)*0.4f), (int32_t) (host_power_meter_level_dbm0(cutoff - 2.5f)*0.4f));
    run_channel(ch, amp, n, chunk, bit_rate, restart_at, restart_old_train, bits, bits_cap, nbits, syms, sym_cap, nsyms, eq_coeff);
    if (final)
    {
        final[0] = ch.istate[I_STAGE];
        final[1] = ch.istate[I_PHASE_RATE];
        final[2] = ch.istate[I_EQ_PUT_STEP];
        final[3] = ch.istate[I_SIGNAL_PRESENT];
        memcpy(&final[4], &ch.fstate[F_AGC], 4);
        final[5] = ch.istate[I_TOTAL_TIMING];
        final[6] = ch.istate[RxV29::I_CONSTELLATION];
        final[7] = ch.istate[I_CARRIER_PHASE];
    }
    return 0;
}

// final[]: as oracle ref_v27ter_run: {training_stage, carrier_phase_rate, eq_put_step, signal_present, agc_scaling bits,
//           total_baud_timing_correction, constellation_state, carrier_phase, gardner_integrate, gardner_step}
EXPORT int hostsim_v27ter_run(const int16_t *amp, int n, int chunk, int bit_rate, float cutoff, int restart_at, int restart_old_train,
                              int8_t *bits, int bits_cap, int32_t *nbits, span_b200_v29_symbol_t *syms, int sym_cap, int32_t *nsyms,
                              float *eq_coeff, int32_t *final)
{
    static HostTables h;
    HostChannel<RxV27ter> ch;
    core_consts(ch.k, h, V27TER_SETS_4800, 1800.0, 30.0, 5, 1.414f);
    make_v27ter_rrc(h.re, h.im);
    ch.k.rrc_re = h.re.data();
    ch.k.rrc_im = h.im.data();
    ch.k.agc_initial = (1.414f/1.000000f)/283.0f;
    ch.k.phase_p45 = host_dds_phase(45.0f);
    ch.k.phase_m45 = host_dds_phase(-45.0f);
    ch.k.phase_180 = host_dds_phase(180.0f);
    ch.k.eq_delta = 0.25f/V27TER_EQ_LEN;
    ch.setup();
    if (cutoff <= -99.0f)
        cutoff = -45.5f;
    ch.init(bit_rate, (int32_t) (host_power_meter_level_dbm0(cutoff + 2.5f)*0.4f), (int32_t) (host_power_meter_level_dbm0(cutoff - 2.5f)*0.4f));
    run_channel(ch, amp, n, chunk, bit_rate, restart_at, restart_old_train, bits, bits_cap, nbits, syms, sym_cap, nsyms, eq_coeff);
    if (final)
    {
        final[0] = ch.istate[I_STAGE];
        final[1] = ch.istate[I_PHASE_RATE];
        final[2] = ch.istate[I_EQ_PUT_STEP];
        final[3] = ch.istate[I_SIGNAL_PRESENT];
        memcpy(&final[4], &ch.fstate[F_AGC], 4);
        final[5] = ch.istate[I_TOTAL_TIMING];
        final[6] = ch.istate[RxV27ter::I_CONSTELLATION];
        final[7] = ch.istate[I_CARRIER_PHASE];
        final[8] = ch.istate[RxV27ter::I_GARDNER_INTEGRATE];
        final[9] = ch.istate[RxV27ter::I_GARDNER_STEP];
    }
    return 0;
}

// ---- FSK receiver (sb_fsk_rx.cuh) on the host: same call sequence as oracle ref_fsk_run ----------------------
// spec5 = {freq_zero, freq_one, tx_level, min_level, baud_rate}; rspec5: the spec of the restart (if restart_at >= 0)
static void fsk_host_restart(sbf::FskRx &r, const int32_t *spec5, int mode)
{
    const float cutoff = (float) spec5[3];
    r.restart(spec5[4], mode, sbf::host_dds_int_phase_rate((float) spec5[0]), sbf::host_dds_int_phase_rate((float) spec5[1]),
              sbf::host_level_dbm0(cutoff + 2.5f - 5.3f), sbf::host_level_dbm0(cutoff - 2.5f - 5.3f));
}

EXPORT int hostsim_fsk_run(const int16_t *amp, int n, int chunk, const int32_t *spec5, int framing_mode, float cutoff,
                           int data_bits, int parity, int stop_bits,
                           int restart_at, const int32_t *rspec5, int restart_mode, int fillin_at, int fillin_len,
                           int16_t *out, int out_cap, int32_t *nout, int32_t *final, int32_t *window)
{
    static std::vector<short> sine;
    if (sine.empty())
        sbf::make_dds_int_table(sine);
    std::vector<int2> win(2*SBF_MAX_WINDOW, make_int2(0, 0));
    std::vector<int> state(sbf::K_COUNT, 0);
    sbf::FskRx r;
    sbf::FskLoader ld = {state.data(), 1, 0};
    r.visit(ld);
    r.win = win.data();
    r.wspan = SBF_MAX_WINDOW;
    r.ls = 1;
    r.sine = sine.data();
    r.out = out;
    r.out_cap = out_cap;
    r.nout = 0;
    fsk_host_restart(r, spec5, framing_mode);
    if (cutoff > -99.0f)
    {
        r.on_power = sbf::host_level_dbm0(cutoff + 2.5f - 5.3f);
        r.off_power = sbf::host_level_dbm0(cutoff - 2.5f - 5.3f);
    }
    if (data_bits > 0)
        r.set_frame_parameters(data_bits, parity, stop_bits);
    if (chunk <= 0)
        chunk = n;
    int len;
    for (int pos = 0;  pos < n;  pos += len)
    {
        if (restart_at >= 0  &&  pos >= restart_at)
        {
            fsk_host_restart(r, rspec5, restart_mode);
            restart_at = -1;
        }
        len = (n - pos < chunk)  ?  (n - pos)  :  chunk;
        // every chunk goes through the state arrays, as every kernel launch does
        sbf::FskStorer st = {state.data(), 1, 0};
        r.visit(st);
        r.visit(ld);
        if (fillin_at >= 0  &&  pos >= fillin_at  &&  pos < fillin_at + fillin_len)
        {
            for (int i = 0;  i < len;  i++)
                r.fillin_sample();
        }
        else
        {
            for (int i = 0;  i < len;  i++)
                r.sample(amp[pos + i]);
        }
    }
    *nout = r.nout;
    sbf::FskStorer st = {state.data(), 1, 0};
    r.visit(st);
    if (final)
        memcpy(final, state.data(), sizeof(int)*sbf::K_COUNT);
    if (window)
        memcpy(window, win.data(), sizeof(int2)*2*SBF_MAX_WINDOW);
    return 0;
}

// modem_connect_tones_rx() in `chunk`-sample calls on the host: ev[] = {call index, tone, level} per report;
// final[17] = the M_* fields, fsk_final[28] = the K_* fields of the embedded V.21 receiver
EXPORT int hostsim_mct_run(const int16_t *amp, int n, int chunk, int tone_type, int32_t *ev, int ev_cap, int32_t *nev,
                           int32_t *final, int32_t *fsk_final)
{
    static std::vector<short> sine;
    if (sine.empty())
        sbf::make_dds_int_table(sine);
    std::vector<int2> win(2*SBF_MAX_WINDOW, make_int2(0, 0));
    std::vector<int> state(sbf::M_COUNT, 0);
    std::vector<int2> rep(64);
    sbf::MctRx r;
    sbf::FskLoader ld = {state.data(), 1, 0};
    sbf::FskStorer st = {state.data(), 1, 0};
    r.fsk.visit(ld);
    r.visit_own(ld, true);
    r.fsk.win = win.data();
    r.fsk.wspan = SBF_MAX_WINDOW;
    r.fsk.ls = 1;
    r.fsk.sine = sine.data();
    r.init(tone_type);
    if (r.uses_v21())
    {
        r.fsk.restart(300*100, sbf::FRAME_MODE_SYNC, sbf::host_dds_int_phase_rate(1850.0f), sbf::host_dds_int_phase_rate(1650.0f),
                      sbf::host_level_dbm0(-45.5f + 2.5f - 5.3f), sbf::host_level_dbm0(-45.5f - 2.5f - 5.3f));
    }
    if (chunk <= 0)
        chunk = n;
    int len;
    int total = 0;
    int call = 0;
    for (int pos = 0;  pos < n;  pos += len, call++)
    {
        len = (n - pos < chunk)  ?  (n - pos)  :  chunk;
        // every chunk goes through the state arrays, as every kernel launch does
        r.fsk.visit(st);
        r.visit_own(st, false);
        r.fsk.visit(ld);
        r.visit_own(ld, true);
        r.ev = rep.data();
        r.ev_cap = (int) rep.size();
        r.nev = 0;
        if (r.uses_v21())
        {
            for (int i = 0;  i < len;  i++)
                r.preamble_sample(amp[pos + i]);
        }
        for (int i = 0;  i < len;  i++)
            r.tone_sample(amp[pos + i]);
        if (r.nev > r.ev_cap)
            return -1;
        for (int i = 0;  i < r.nev;  i++, total++)
        {
            if (total < ev_cap)
            {
                ev[3*total] = call;
                ev[3*total + 1] = rep[i].x & 0xFFFF;
                ev[3*total + 2] = sbf::host_mct_level(rep[i].x >> 16, rep[i].y);
            }
        }
    }
    *nev = total;
    r.fsk.visit(st);
    r.visit_own(st, false);
    if (final)
        memcpy(final, state.data() + sbf::K_COUNT, sizeof(int)*(sbf::M_COUNT - sbf::K_COUNT));
    if (fsk_final)
        memcpy(fsk_final, state.data(), sizeof(int)*sbf::K_COUNT);
    return 0;
}

EXPORT void hostsim_fsk_tables(int16_t *sine)
{
    std::vector<short> t;
    sbf::make_dds_int_table(t);
    memcpy(sine, t.data(), sizeof(short)*257);
}

// ------------------------------------------------------------------------------------------
// The signal sources of sb_gen.cuh on the host (same layout of arguments as oracle/ref_harness_gen.c)

EXPORT int hostsim_dtmf_tx_calls(int16_t *amp, const int32_t *max_lens, int ncalls, const char *digits, const char *digits2, int put2_before_call,
                                 int set_level, int level, int twist, int set_timing, int on_ms, int off_ms,
                                 int32_t *out_lens, int32_t *put_results)
{
    static std::vector<float> sine;
    if (sine.empty())
    {
        sine.resize(SBG_SINE_WORDS);
        sbg::host_make_sine_table(sine.data());
    }
    static const int row[4] = {697, 770, 852, 941};
    static const int col[4] = {1209, 1336, 1477, 1633};
    int rates[8];
    for (int i = 0;  i < 4;  i++)
    {
        rates[i] = sbg::host_dds_phase_ratef((float) row[i]);
        rates[4 + i] = sbg::host_dds_phase_ratef((float) col[i]);
    }
    std::vector<unsigned char> queue(SBG_QUEUE, 0);
    std::vector<int> state(sbg::D_COUNT, 0);
    sbg::GenLoader ld = {state.data(), 1, 0};
    sbg::GenStorer st = {state.data(), 1, 0};
    sbg::DtmfTx t;
    t.tones.sine = sine.data();
    t.queue = queue.data();
    t.qstride = 1;
    t.row_rate = rates;
    t.col_rate = rates + 4;
    t.init(sbg::host_dds_scaling_dbm0f(-10.0f));
    if (set_level)
    {
        t.low_level = sbg::host_dds_scaling_dbm0f((float) level);
        t.high_level = sbg::host_dds_scaling_dbm0f((float) (level + twist));
    }
    if (set_timing)
    {
        t.on_time = ((on_ms >= 0)  ?  on_ms  :  50)*8;
        t.off_time = ((off_ms >= 0)  ?  off_ms  :  55)*8;
    }
    auto put = [&](const char *d) -> int
    {
        const int len = (int) strlen(d);
        if (len == 0)
            return 0;
        const int space = SBG_QUEUE - t.qcnt;
        if (space < len)
            return len - space;
        for (int i = 0;  i < len;  i++)
            queue[(t.qrd + t.qcnt + i) & (SBG_QUEUE - 1)] = (unsigned char) d[i];
        t.qcnt += len;
        return 0;
    };
    put_results[0] = put(digits);
    put_results[1] = 0;
    int pos = 0;
    for (int k = 0;  k < ncalls;  k++)
    {
        if (digits2  &&  k == put2_before_call)
            put_results[1] = put(digits2);
        // every call goes through the state arrays, as every kernel launch does
        t.store(st);
        t.load(ld);
        sbg::RowOut out;
        out.begin(amp + pos);
        out_lens[k] = t.tx(out, max_lens[k]);
        out.flush();
        pos += max_lens[k];
    }
    return 0;
}

EXPORT int hostsim_awgn_run(int16_t *amp, int n, int seed, float level, int dbov, int add)
{
    std::vector<double> r(SBG_RAN_TABLE);
    sbg::Awgn g;
    g.r = r.data();
    g.rs = 1;
    g.ran_init(seed);
    if (!dbov)
        level = level - (3.14f + 3.02f);
    g.rms = pow(10.0, level/20.0)*32768.0;
    g.amp2 = 0.0;
    g.odd = 1;
    for (int i = 0;  i < n;  i++)
        amp[i] = (int16_t) ((add)  ?  sbg::sat_add16(amp[i], g.sample())  :  g.sample());
    return 0;
}

// tone_gen() with a full descriptor (same arguments as oracle/ref_harness_tx.c: ref_tone_gen_calls)
EXPORT int hostsim_tone_gen_calls(int16_t *amp, const int32_t *max_lens, int ncalls, const int32_t *desc, int32_t *out_lens)
{
    static std::vector<float> sine;
    if (sine.empty())
    {
        sine.resize(SBG_SINE_WORDS);
        sbg::host_make_sine_table(sine.data());
    }
    sbg::ToneDesc d;
    sbg::host_tone_descriptor(d, desc[0], desc[1], desc[2], desc[3], desc[4], desc[5], desc[6], desc[7], desc[8] != 0);
    std::vector<int> state(sbg::T_COUNT, 0);
    sbg::GenLoader ld = {state.data(), 1, 0};
    sbg::GenStorer st = {state.data(), 1, 0};
    sbg::ToneGen t;
    t.sine = sine.data();
    sbg::tone_gen_start(t, d);
    int pos = 0;
    for (int k = 0;  k < ncalls;  k++)
    {
        sbg::tone_gen_store(t, st);
        sbg::tone_gen_load(t, ld);
        sbg::RowOut out;
        out.begin(amp + pos);
        out_lens[k] = t.run(out, max_lens[k]);
        out.flush();
        pos += max_lens[k];
    }
    return 0;
}

// v29_tx() (same arguments as oracle/ref_harness_tx.c: ref_v29_tx_calls)
EXPORT int hostsim_v29_tx_calls(int16_t *amp, const int32_t *max_lens, int ncalls, int bit_rate, int tep, float power_dbm0,
                                int src_mode, uint32_t lfsr_seed, const uint8_t *bits, int nbits,
                                int restart_before_call, int restart_rate, int restart_tep,
                                int32_t *out_lens, int32_t *status)
{
    static std::vector<float> sine;
    static std::vector<float> shaper;
    if (sine.empty())
    {
        sine.resize(SBG_SINE_WORDS);
        sbg::host_make_sine_table(sine.data());
        sbm::make_tx_rrc(shaper, SBG_V29_TX_SETS, SBG_V29_TX_STEPS, 0.25);
    }
    if (bit_rate != 9600  &&  bit_rate != 7200  &&  bit_rate != 4800)
        return -1;
    std::vector<int> state(sbg::X_COUNT, 0);
    sbg::GenLoader ld = {state.data(), 1, 0};
    sbg::GenStorer st = {state.data(), 1, 0};
    sbg::V29Tx t;
    t.load(ld);
    t.sine = sine.data();
    t.shaper = shaper.data();
    t.bits = bits;
    t.carrier_phase_rate = sbg::host_dds_phase_ratef(1700.0f);
    t.base_gain = powf(10.0f, (-14.0f - 3.14f)/20.0f)*32768.0f/1.000000f;
    t.restart(bit_rate, tep);
    t.base_gain = powf(10.0f, (power_dbm0 - 3.14f)/20.0f)*32768.0f/1.000000f;
    t.set_working_gain();
    t.src_mode = src_mode;
    t.lfsr = (lfsr_seed & 0x7FFFFF)  ?  (lfsr_seed & 0x7FFFFF)  :  1;
    t.bit_pos = 0;
    t.bit_count = nbits;
    int pos = 0;
    for (int k = 0;  k < ncalls;  k++)
    {
        if (k == restart_before_call)
            t.restart(restart_rate, restart_tep);
        t.store(st);
        t.load(ld);
        sbg::RowOut out;
        out.begin(amp + pos);
        out_lens[k] = t.tx(out, max_lens[k]);
        out.flush();
        pos += max_lens[k];
    }
    *status = t.status;
    return 0;
}

EXPORT void hostsim_v29_tx_tables(float *shaper)
{
    std::vector<float> t;
    sbm::make_tx_rrc(t, SBG_V29_TX_SETS, SBG_V29_TX_STEPS, 0.25);
    memcpy(shaper, t.data(), sizeof(float)*t.size());
}

EXPORT void hostsim_gen_tables(float *sine)
{
    sbg::host_make_sine_table(sine);
}

// ------------------------------------------------------------------------------------------
// The signalling tone receiver of sb_sig_rx.cuh on the host (same arguments as oracle/ref_harness_sig.c: ref_sig_run)
EXPORT int hostsim_sig_run(int16_t *amp, int n, int chunk, const int32_t *lens, int ncalls, int tone_type, const int32_t *modes, int nmodes,
                           int32_t *ev, int ev_cap, int32_t *nev, int32_t *final)
{
    if (tone_type < 1  ||  tone_type > 3)
        return -1;
    std::vector<int> state(sbs::T_COUNT, 0);
    std::vector<int2> rep(16384);
    sbs::SigLoader ld = {state.data(), 1, 0};
    sbs::SigStorer st = {state.data(), 1, 0};
    sbs::SigRx r;
    int32_t thr[3];
    sbs::host_sig_thresholds(tone_type, thr);
    r.init(tone_type, thr[0], thr[1], thr[2]);
    if (chunk <= 0)
        chunk = n;
    int len;
    int call = 0;
    int total = 0;
    for (int pos = 0;  (lens)  ?  (call < ncalls)  :  (pos < n);  pos += len, call++)
    {
        for (int m = 0;  m < nmodes;  m++)
        {
            if (modes[2*m] == call)
                r.current_rx_tone = modes[2*m + 1];
        }
        len = (lens)  ?  lens[call]  :  ((n - pos < chunk)  ?  (n - pos)  :  chunk);
        // every call goes through the state arrays, as every kernel launch does
        r.store(st);
        r.load(ld);
        r.ev = rep.data();
        r.ev_cap = (int) rep.size();
        r.nev = 0;
        for (int i = 0;  i < len;  i++)
            amp[pos + i] = (int16_t) r.sample(amp[pos + i]);
        if (r.nev > r.ev_cap)
            return -1;
        for (int i = 0;  i < r.nev;  i++, total++)
        {
            if (total < ev_cap)
            {
                ev[3*total] = call;
                ev[3*total + 1] = rep[i].x;
                ev[3*total + 2] = rep[i].y;
            }
        }
    }
    *nev = total;
    r.store(st);
    if (final)
        memcpy(final, state.data(), sizeof(int)*sbs::T_COUNT);
    return 0;
}
