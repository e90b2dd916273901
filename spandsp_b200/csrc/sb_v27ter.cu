// sb_v27ter.cu - C ABI of the V.27ter receiver banks (include/spandsp_b200_v27ter.h).  The receiver itself is
// sb_v27ter_rx.cuh on top of the shared modem core (sb_modem.cuh); the host bookkeeping is sb_modem_bank.cuh.
// Reference: src/v27ter_rx.c.
#include "sb_modem_bank.cuh"
#include "sb_v27ter_rx.cuh"

#pragma GCC visibility push(default)
#include "../../include/spandsp_b200_v27ter.h"
#pragma GCC visibility pop

using namespace sbm;

struct span_b200_v27ter_bank_s : ModemBank<RxV27ter>
{
};

static bool v27ter_rate_ok(int bit_rate)
{
    return bit_rate == 4800  ||  bit_rate == 2400;
}

extern "C" int span_b200_v27ter_tables(float *rrc4800_re, float *rrc4800_im, float *rrc2400_re, float *rrc2400_im, int32_t *ints)
{
    std::vector<float> re;
    std::vector<float> im;
    make_v27ter_rrc(re, im);
    const size_t n48 = (size_t) V27TER_SETS_4800*SBM_FILTER_STEPS;
    const size_t n24 = (size_t) V27TER_SETS_2400*SBM_FILTER_STEPS;
    memcpy(rrc4800_re, re.data(), sizeof(float)*n48);
    memcpy(rrc4800_im, im.data(), sizeof(float)*n48);
    memcpy(rrc2400_re, re.data() + n48, sizeof(float)*n24);
    memcpy(rrc2400_im, im.data() + n48, sizeof(float)*n24);
    ints[0] = V27TER_SETS_4800;
    ints[1] = V27TER_SETS_2400;
    ints[2] = host_dds_phase_rate(1800.0f);
    ints[3] = host_dds_phase_rate(1800.0f - 20.0f);
    ints[4] = host_dds_phase_rate(1800.0f + 20.0f);
    ints[5] = host_dds_phase(45.0f);
    ints[6] = host_dds_phase(-45.0f);
    ints[7] = host_dds_phase(180.0f);
    return 0;
}

extern "C" span_b200_v27ter_bank_t *span_b200_v27ter_bank_create(span_b200_ctx_t *ctx, int channels, int bit_rate, int want_symbols)
{
    if (ctx == NULL  ||  channels <= 0  ||  !v27ter_rate_ok(bit_rate))
    {
        sb_set_error("bad V.27ter bank arguments (bit rate must be 4800 or 2400)");         // src/v27ter_rx.c:1163-1171
        return NULL;
    }
    SB_DEVICE_CKP(span_b200_ctx_device(ctx));
    span_b200_v27ter_bank_t *b = new span_b200_v27ter_bank_s();
    b->ctx = ctx;
    b->channels = channels;
    b->bit_rate = bit_rate;
    b->want_symbols = (want_symbols != 0);
    b->bits_per_sample_x2 = 2;          // 3 bits per baud at 0.2 baud per sample, plus margin
    // v27ter_rx_set_signal_cutoff(s, -45.5f): src/v27ter_rx.c:158-163,1183
    b->on_power = (int32_t) (host_power_meter_level_dbm0(-45.5f + 2.5f)*0.4f);
    b->off_power = (int32_t) (host_power_meter_level_dbm0(-45.5f - 2.5f)*0.4f);
    b->k.phase_p45 = host_dds_phase(45.0f);
    b->k.phase_m45 = host_dds_phase(-45.0f);
    b->k.phase_180 = host_dds_phase(180.0f);
    b->k.eq_delta = 0.25f/V27TER_EQ_LEN;                                // EQUALIZER_DELTA, src/v27ter_rx.c:94,237
    std::vector<float> re;
    std::vector<float> im;
    make_v27ter_rrc(re, im);
    // The Godard descriptor is not used by this receiver (Gardner timing); the arguments only fill the fields.
    if (modem_core_tables(b, 1800.0, 30.0, 5, 1.414f, 283.0f, &re, &im) != 0    // src/v27ter_rx.c:87,1141
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

extern "C" void span_b200_v27ter_bank_destroy(span_b200_v27ter_bank_t *b)
{
    modem_destroy(b);
}

extern "C" int span_b200_v27ter_bank_channels(const span_b200_v27ter_bank_t *b)
{
    return b->channels;
}

extern "C" int span_b200_v27ter_bank_restart(span_b200_v27ter_bank_t *b, int first, int count, int bit_rate, int old_train)
{
    if (!modem_range_ok(b, first, count)  ||  !v27ter_rate_ok(bit_rate))
    {
        sb_set_error("bad restart arguments");
        return -1;                                  // src/v27ter_rx.c:1095-1103
    }
    if (modem_quiesce(b) != 0)
        return -1;
    return modem_init_channels(b, first, count, bit_rate, (old_train)  ?  1  :  0);
}

extern "C" int span_b200_v27ter_bank_set_signal_cutoff(span_b200_v27ter_bank_t *b, int first, int count, float cutoff)
{
    return modem_set_signal_cutoff(b, first, count, cutoff);
}

extern "C" int span_b200_v27ter_bank_fillin(span_b200_v27ter_bank_t *b, int first, int count, int samples)
{
    return modem_fillin(b, first, count, samples);
}

extern "C" int span_b200_v27ter_bank_rx_device(span_b200_v27ter_bank_t *b, const int16_t *d_amp, int64_t stride, int n, void *stream)
{
    return modem_rx_device(b, d_amp, stride, n, stream);
}

extern "C" int span_b200_v27ter_bank_rx_host(span_b200_v27ter_bank_t *b, const int16_t *h_amp, int64_t stride, int n, void *stream)
{
    return modem_rx_host(b, h_amp, stride, n, stream);
}

extern "C" int span_b200_v27ter_bank_counts(span_b200_v27ter_bank_t *b, int32_t *nbits, int32_t *nsyms)
{
    return modem_counts(b, nbits, nsyms);
}

extern "C" int64_t span_b200_v27ter_bank_bits(span_b200_v27ter_bank_t *b, int channel, int8_t *out, int64_t max)
{
    return modem_bits(b, channel, out, max);
}

extern "C" int64_t span_b200_v27ter_bank_bits_all(span_b200_v27ter_bank_t *b, int8_t *out, int64_t out_stride, int32_t *nbits)
{
    return modem_bits_all(b, out, out_stride, nbits);
}

extern "C" int64_t span_b200_v27ter_bank_symbols(span_b200_v27ter_bank_t *b, int channel, span_b200_v27ter_symbol_t *out, int64_t max)
{
    return modem_symbols(b, channel, out, max);
}

extern "C" int span_b200_v27ter_bank_output_layout(span_b200_v27ter_bank_t *b, const uint32_t **d_words, int64_t *words_cap, const int32_t **d_nbits,
                                                const int32_t **d_status, int64_t *status_cap, const int32_t **d_nstatus,
                                                const span_b200_v29_symbol_t **d_syms, int64_t *sym_cap, const int32_t **d_nsyms)
{
    return modem_output_layout(b, d_words, words_cap, d_nbits, d_status, status_cap, d_nstatus, d_syms, sym_cap, d_nsyms);
}

extern "C" int64_t span_b200_v27ter_bank_output_packed(span_b200_v27ter_bank_t *b, uint32_t *words, int64_t words_stride, int32_t *nbits,
                                                   int32_t *status, int64_t status_stride, int32_t *nstatus)
{
    return modem_output_packed(b, words, words_stride, nbits, status, status_stride, nstatus);
}

extern "C" int span_b200_v27ter_bank_channel_state(span_b200_v27ter_bank_t *b, int channel, float *eq_coeff, int32_t *info)
{
    static const int fields[12] = {I_STAGE, I_PHASE_RATE, I_EQ_PUT_STEP, I_SIGNAL_PRESENT, -1 - F_AGC,
                                   I_TOTAL_TIMING, RxV27ter::I_CONSTELLATION, I_CARRIER_PHASE, RxV27ter::I_GARDNER_INTEGRATE,
                                   RxV27ter::I_GARDNER_STEP, I_POWER, I_BIT_RATE};
    return modem_channel_state(b, channel, eq_coeff, info, fields, 12, V27TER_EQ_LEN);
}
