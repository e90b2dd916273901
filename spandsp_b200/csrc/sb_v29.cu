// sb_v29.cu - C ABI of the V.29 receiver banks (include/spandsp_b200_v29.h).  The receiver itself is
// sb_v29_rx.cuh on top of the shared modem core (sb_modem.cuh); the host bookkeeping is sb_modem_bank.cuh.
// Reference: src/v29rx.c.
#include "sb_modem_bank.cuh"
#include "sb_v29_rx.cuh"

namespace sbm {

// V.29 banks run the four-lanes-per-receiver form of the receiver, two warps per CTA (same state arrays, same results)
template <>
struct RxKernel<RxV29>
{
    typedef RxV29x4 type;
    static const int WARPS = 2;
};

}  // namespace sbm

using namespace sbm;

struct span_b200_v29_bank_s : ModemBank<RxV29>
{
};

static bool v29_rate_ok(int bit_rate)
{
    return bit_rate == 9600  ||  bit_rate == 7200  ||  bit_rate == 4800;
}

extern "C" int span_b200_v29_tables(float *rrc_re, float *rrc_im, float *sine, uint16_t *sqrt_tab, float *godard, int32_t *ints)
{
    std::vector<float> re;
    std::vector<float> im;
    std::vector<float> st;
    std::vector<unsigned short> sq;
    godard_desc_t g;
    make_rx_rrc(re, im, V29_COEFF_SETS, 1700.0);
    make_sine_table(st);
    make_sqrt_table(sq);
    make_godard(g, 1700.0, 30.0, 5);
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
    ints[2] = host_dds_phase_rate(1700.0f);
    ints[3] = host_dds_phase_rate(1700.0f - 20.0f);
    ints[4] = host_dds_phase_rate(1700.0f + 20.0f);
    ints[5] = host_dds_phase(45.0f);
    ints[6] = host_dds_phase(-45.0f);
    ints[7] = host_power_meter_level_dbm0(-28.5f + 2.5f);
    ints[8] = host_power_meter_level_dbm0(-28.5f - 2.5f);
    return 0;
}

extern "C" span_b200_v29_bank_t *span_b200_v29_bank_create(span_b200_ctx_t *ctx, int channels, int bit_rate, int want_symbols)
{
    if (ctx == NULL  ||  channels <= 0  ||  !v29_rate_ok(bit_rate))
    {
        sb_set_error("bad V.29 bank arguments (bit rate must be 9600, 7200 or 4800)");      // src/v29rx.c:1102-1111
        return NULL;
    }
    SB_DEVICE_CKP(span_b200_ctx_device(ctx));
    span_b200_v29_bank_t *b = new span_b200_v29_bank_s();
    b->ctx = ctx;
    b->channels = channels;
    b->bit_rate = bit_rate;
    b->want_symbols = (want_symbols != 0);
    b->bits_per_sample_x2 = 3;          // 4 bits per baud, 0.3 baud per sample, plus margin
    V29Tables t;
    make_v29_tables(t);
    // v29_rx_set_signal_cutoff(s, -28.5f): src/v29rx.c:163-168,1129
    b->on_power = (int32_t) (host_power_meter_level_dbm0(-28.5f + 2.5f)*0.4f);
    b->off_power = (int32_t) (host_power_meter_level_dbm0(-28.5f - 2.5f)*0.4f);
    b->k.phase_p45 = host_dds_phase(45.0f);
    b->k.phase_m45 = host_dds_phase(-45.0f);
    b->k.eq_delta = 0.21f/SBM_EQ_LEN;                                   // src/v29rx.c:97,240
    if (modem_core_tables(b, 1700.0, 30.0, 5, 1.25f) != 0               // src/v29rx.c:91,1078; src/Makefile.am:556-560
        ||
        modem_upload(b->owned, &b->k.tables, &t, sizeof(t)) != 0
        ||
        modem_alloc_state(b) != 0
        ||
        modem_init_channels(b, 0, channels, bit_rate, -1) != 0)
    {
        modem_destroy(b);
        return NULL;
    }
    return b;
}

extern "C" void span_b200_v29_bank_destroy(span_b200_v29_bank_t *b)
{
    modem_destroy(b);
}

extern "C" int span_b200_v29_bank_channels(const span_b200_v29_bank_t *b)
{
    return b->channels;
}

extern "C" int span_b200_v29_bank_restart_ex(span_b200_v29_bank_t *b, int first, int count, int bit_rate, int old_train)
{
    if (!modem_range_ok(b, first, count)  ||  !v29_rate_ok(bit_rate))
    {
        sb_set_error("bad restart arguments");
        return -1;                                  // src/v29rx.c:1033
    }
    if (modem_quiesce(b) != 0)
        return -1;
    return modem_init_channels(b, first, count, bit_rate, (old_train)  ?  1  :  0);
}

extern "C" int span_b200_v29_bank_restart(span_b200_v29_bank_t *b, int first, int count, int bit_rate)
{
    return span_b200_v29_bank_restart_ex(b, first, count, bit_rate, 0);
}

extern "C" int span_b200_v29_bank_set_signal_cutoff(span_b200_v29_bank_t *b, int first, int count, float cutoff)
{
    return modem_set_signal_cutoff(b, first, count, cutoff);
}

extern "C" int span_b200_v29_bank_fillin(span_b200_v29_bank_t *b, int first, int count, int samples)
{
    return modem_fillin(b, first, count, samples);
}

extern "C" int span_b200_v29_bank_rx_device(span_b200_v29_bank_t *b, const int16_t *d_amp, int64_t stride, int n, void *stream)
{
    return modem_rx_device(b, d_amp, stride, n, stream);
}

extern "C" int span_b200_v29_bank_rx_host(span_b200_v29_bank_t *b, const int16_t *h_amp, int64_t stride, int n, void *stream)
{
    return modem_rx_host(b, h_amp, stride, n, stream);
}

extern "C" int span_b200_v29_bank_counts(span_b200_v29_bank_t *b, int32_t *nbits, int32_t *nsyms)
{
    return modem_counts(b, nbits, nsyms);
}

extern "C" int64_t span_b200_v29_bank_bits(span_b200_v29_bank_t *b, int channel, int8_t *out, int64_t max)
{
    return modem_bits(b, channel, out, max);
}

extern "C" int64_t span_b200_v29_bank_bits_all(span_b200_v29_bank_t *b, int8_t *out, int64_t out_stride, int32_t *nbits)
{
    return modem_bits_all(b, out, out_stride, nbits);
}

extern "C" int64_t span_b200_v29_bank_symbols(span_b200_v29_bank_t *b, int channel, span_b200_v29_symbol_t *out, int64_t max)
{
    return modem_symbols(b, channel, out, max);
}

extern "C" int span_b200_v29_bank_output_layout(span_b200_v29_bank_t *b, const uint32_t **d_words, int64_t *words_cap, const int32_t **d_nbits,
                                                const int32_t **d_status, int64_t *status_cap, const int32_t **d_nstatus,
                                                const span_b200_v29_symbol_t **d_syms, int64_t *sym_cap, const int32_t **d_nsyms)
{
    return modem_output_layout(b, d_words, words_cap, d_nbits, d_status, status_cap, d_nstatus, d_syms, sym_cap, d_nsyms);
}

extern "C" int64_t span_b200_v29_bank_output_packed(span_b200_v29_bank_t *b, uint32_t *words, int64_t words_stride, int32_t *nbits,
                                                   int32_t *status, int64_t status_stride, int32_t *nstatus)
{
    return modem_output_packed(b, words, words_stride, nbits, status, status_stride, nstatus);
}

// Receiver status of one channel: v29_rx_equalizer_state / carrier_frequency / symbol_timing_correction /
// signal_power (src/v29rx.c:145-161,180-195) material.
extern "C" int span_b200_v29_bank_channel_state(span_b200_v29_bank_t *b, int channel, float *eq_coeff, int32_t *info)
{
    static const int fields[10] = {I_STAGE, I_PHASE_RATE, I_EQ_PUT_STEP, I_SIGNAL_PRESENT, -1 - F_AGC,
                                   I_TOTAL_TIMING, RxV29::I_CONSTELLATION, I_CARRIER_PHASE, I_POWER, I_BIT_RATE};
    return modem_channel_state(b, channel, eq_coeff, info, fields, 10);
}
