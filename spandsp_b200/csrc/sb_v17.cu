// sb_v17.cu - C ABI of the V.17 receiver banks (include/spandsp_b200_v17.h).  The receiver itself is
// sb_v17_rx.cuh on top of the shared modem core (sb_modem.cuh); the host bookkeeping is sb_modem_bank.cuh.
// Reference: src/v17rx.c.
#include "sb_modem_bank.cuh"
#include "sb_v17_rx.cuh"

#pragma GCC visibility push(default)
#include "../../include/spandsp_b200_v17.h"
#pragma GCC visibility pop

using namespace sbm;

struct span_b200_v17_bank_s : ModemBank<RxV17>
{
};

static bool v17_rate_ok(int bit_rate)
{
    return bit_rate == 14400  ||  bit_rate == 12000  ||  bit_rate == 9600  ||  bit_rate == 7200  ||  bit_rate == 4800;
}

extern "C" int span_b200_v17_tables(float *rrc_re, float *rrc_im, float *godard, int32_t *ints, float *constellations,
                                    uint8_t *maps, uint8_t *map4800)
{
    std::vector<float> re;
    std::vector<float> im;
    godard_desc_t g;
    make_rx_rrc(re, im, V17_COEFF_SETS, 1800.0);
    make_godard(g, 1800.0, 100.0, 15);
    memcpy(rrc_re, re.data(), sizeof(float)*re.size());
    memcpy(rrc_im, im.data(), sizeof(float)*im.size());
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
    ints[2] = host_dds_phase_rate(1800.0f);
    ints[3] = host_dds_phase_rate(1800.0f - 20.0f);
    ints[4] = host_dds_phase_rate(1800.0f + 20.0f);
    ints[5] = host_dds_phase(90.0f);
    ints[6] = host_dds_phase(-90.0f);
    ints[7] = host_dds_phase(180.0f);
    ints[8] = host_dds_phase(270.0f + 18.433f);
    ints[9] = host_dds_phase(180.0f + 18.433f);
    ints[10] = host_dds_phase(18.433f);
    ints[11] = V17_COEFF_SETS;
    V17Tables t;
    make_v17_tables(t);
    memcpy(constellations, t.constellation, sizeof(t.constellation));
    std::vector<unsigned char> m;
    std::vector<unsigned char> m48;
    make_v17_maps(t, m, m48);
    memcpy(maps, m.data(), m.size());
    memcpy(map4800, m48.data(), m48.size());
    return 0;
}

extern "C" span_b200_v17_bank_t *span_b200_v17_bank_create(span_b200_ctx_t *ctx, int channels, int bit_rate, int want_symbols)
{
    if (ctx == NULL  ||  channels <= 0  ||  !v17_rate_ok(bit_rate))
    {
        sb_set_error("bad V.17 bank arguments (bit rate must be 14400, 12000, 9600, 7200 or 4800)");    // src/v17rx.c:1498-1510
        return NULL;
    }
    SB_DEVICE_CKP(span_b200_ctx_device(ctx));
    span_b200_v17_bank_t *b = new span_b200_v17_bank_s();
    b->ctx = ctx;
    b->channels = channels;
    b->bit_rate = bit_rate;
    b->want_symbols = (want_symbols != 0);
    b->bits_per_sample_x2 = 4;          // 6 bits per baud, 0.3 baud per sample, plus margin
    V17Tables t;
    make_v17_tables(t);
    std::vector<unsigned char> m;
    std::vector<unsigned char> m48;
    make_v17_maps(t, m, m48);
    // v17_rx_set_signal_cutoff(s, -45.5f): src/v17rx.c:173-178,1524
    b->on_power = (int32_t) (host_power_meter_level_dbm0(-45.5f + 2.5f)*0.4f);
    b->off_power = (int32_t) (host_power_meter_level_dbm0(-45.5f - 2.5f)*0.4f);
    b->k.phase_p90 = host_dds_phase(90.0f);
    b->k.phase_m90 = host_dds_phase(-90.0f);
    b->k.phase_180 = host_dds_phase(180.0f);
    b->k.phase_a = host_dds_phase(270.0f + 18.433f);
    b->k.phase_b = host_dds_phase(180.0f + 18.433f);
    b->k.phase_c = host_dds_phase(18.433f);
    const float fast = 0.21f/SBM_EQ_LEN;                                // EQUALIZER_FAST_ADAPTION_DELTA, src/v17rx.c:112
    b->k.eq_delta_fast = fast;
    b->k.eq_delta_slow = 0.1f*fast;                                     // EQUALIZER_SLOW_ADAPTION_DELTA, src/v17rx.c:114
    if (modem_core_tables(b, 1800.0, 100.0, 15, 2.17f) != 0             // src/v17rx.c:106,1474; src/Makefile.am:487-491
        ||
        modem_upload(b->owned, &b->k.tables, &t, sizeof(t)) != 0
        ||
        modem_upload(b->owned, &b->k.maps, m.data(), m.size()) != 0
        ||
        modem_upload(b->owned, &b->k.map4800, m48.data(), m48.size()) != 0
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

extern "C" void span_b200_v17_bank_destroy(span_b200_v17_bank_t *b)
{
    modem_destroy(b);
}

extern "C" int span_b200_v17_bank_channels(const span_b200_v17_bank_t *b)
{
    return b->channels;
}

extern "C" int span_b200_v17_bank_restart(span_b200_v17_bank_t *b, int first, int count, int bit_rate, int short_train)
{
    if (!modem_range_ok(b, first, count)  ||  !v17_rate_ok(bit_rate)  ||  short_train < 0  ||  short_train > 2)
    {
        sb_set_error("bad restart arguments");
        return -1;                                  // src/v17rx.c:1425
    }
    if (modem_quiesce(b) != 0)
        return -1;
    return modem_init_channels(b, first, count, bit_rate, short_train);
}

extern "C" int span_b200_v17_bank_set_signal_cutoff(span_b200_v17_bank_t *b, int first, int count, float cutoff)
{
    return modem_set_signal_cutoff(b, first, count, cutoff);
}

extern "C" int span_b200_v17_bank_fillin(span_b200_v17_bank_t *b, int first, int count, int samples)
{
    return modem_fillin(b, first, count, samples);
}

extern "C" int span_b200_v17_bank_rx_device(span_b200_v17_bank_t *b, const int16_t *d_amp, int64_t stride, int n, void *stream)
{
    return modem_rx_device(b, d_amp, stride, n, stream);
}

extern "C" int span_b200_v17_bank_rx_host(span_b200_v17_bank_t *b, const int16_t *h_amp, int64_t stride, int n, void *stream)
{
    return modem_rx_host(b, h_amp, stride, n, stream);
}

extern "C" int span_b200_v17_bank_counts(span_b200_v17_bank_t *b, int32_t *nbits, int32_t *nsyms)
{
    return modem_counts(b, nbits, nsyms);
}

extern "C" int64_t span_b200_v17_bank_bits(span_b200_v17_bank_t *b, int channel, int8_t *out, int64_t max)
{
    return modem_bits(b, channel, out, max);
}

extern "C" int64_t span_b200_v17_bank_bits_all(span_b200_v17_bank_t *b, int8_t *out, int64_t out_stride, int32_t *nbits)
{
    return modem_bits_all(b, out, out_stride, nbits);
}

extern "C" int64_t span_b200_v17_bank_symbols(span_b200_v17_bank_t *b, int channel, span_b200_v17_symbol_t *out, int64_t max)
{
    return modem_symbols(b, channel, out, max);
}

extern "C" int span_b200_v17_bank_output_layout(span_b200_v17_bank_t *b, const uint32_t **d_words, int64_t *words_cap, const int32_t **d_nbits,
                                                const int32_t **d_status, int64_t *status_cap, const int32_t **d_nstatus,
                                                const span_b200_v29_symbol_t **d_syms, int64_t *sym_cap, const int32_t **d_nsyms)
{
    return modem_output_layout(b, d_words, words_cap, d_nbits, d_status, status_cap, d_nstatus, d_syms, sym_cap, d_nsyms);
}

extern "C" int64_t span_b200_v17_bank_output_packed(span_b200_v17_bank_t *b, uint32_t *words, int64_t words_stride, int32_t *nbits,
                                                   int32_t *status, int64_t status_stride, int32_t *nstatus)
{
    return modem_output_packed(b, words, words_stride, nbits, status, status_stride, nstatus);
}

extern "C" int span_b200_v17_bank_channel_state(span_b200_v17_bank_t *b, int channel, float *eq_coeff, int32_t *info)
{
    static const int fields[12] = {I_STAGE, I_PHASE_RATE, I_EQ_PUT_STEP, I_SIGNAL_PRESENT, -1 - F_AGC,
                                   I_TOTAL_TIMING, RxV17::I_DIFF, I_CARRIER_PHASE, I_POWER, I_BIT_RATE,
                                   RxV17::I_SHORT_TRAIN, RxV17::I_TRELLIS_PTR};
    return modem_channel_state(b, channel, eq_coeff, info, fields, 12);
}
