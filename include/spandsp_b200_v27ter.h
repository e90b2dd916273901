/*
 * spandsp_b200_v27ter.h - C ABI of the V.27ter receiver banks (bulk interface).
 *
 * A bank = N independent V.27ter receivers (src/v27ter_rx.c) processed by one call; channel c reads
 * d_amp[c*stride .. c*stride + samples).  What the reference delivers through callbacks is returned
 * as per-channel streams, exactly as for V.29 (spandsp_b200_v29.h):
 *   - the put_bit stream (span_put_bit_func_t, src/spandsp/async.h:123): one int8 per call, 0/1 for
 *     data bits and the negative SIG_STATUS_* codes (async.h:66-103) exactly where the reference would
 *     have delivered them (no separate status handler: src/v27ter_rx.c:166-174);
 *   - optionally the qam_report_handler_t stream (src/spandsp/v29rx.h:130): the equalized symbol once per
 *     baud in every training stage (src/v27ter_rx.c:765-780), and the Gardner timing hops, which the
 *     reference reports with NULL pointers and the integrator value (src/v27ter_rx.c:517-518): those
 *     records have NaN coordinates and state = the integrator value.
 */
#if !defined(_SPANDSP_B200_V27TER_H_)
#define _SPANDSP_B200_V27TER_H_

#include <stdint.h>

#include "spandsp_b200.h"
#include "spandsp_b200_v29.h"

#if defined(__cplusplus)
extern "C"
{
#endif

typedef struct span_b200_v27ter_bank_s span_b200_v27ter_bank_t;

typedef span_b200_v29_symbol_t span_b200_v27ter_symbol_t;

/* v27ter_rx_init(NULL, bit_rate, ...) x channels (src/v27ter_rx.c:1161-1188).  bit_rate: 4800 or 2400;
   anything else fails as in the reference.  want_symbols != 0 records the qam_report stream. */
span_b200_v27ter_bank_t *span_b200_v27ter_bank_create(span_b200_ctx_t *ctx, int channels, int bit_rate, int want_symbols);
void span_b200_v27ter_bank_destroy(span_b200_v27ter_bank_t *bank);
int span_b200_v27ter_bank_channels(const span_b200_v27ter_bank_t *bank);
/* v27ter_rx_restart(s, bit_rate, old_train) (src/v27ter_rx.c:1091-1158) for channels [first, first+count).
   Channels of one bank may run at different rates.  The reference never records old_train (the test at
   :1132 reads a field that is always false), so every restart is a full retrain; same here. */
int span_b200_v27ter_bank_restart(span_b200_v27ter_bank_t *bank, int first, int count, int bit_rate, int old_train);
/* v27ter_rx_set_signal_cutoff() (src/v27ter_rx.c:158-163) */
int span_b200_v27ter_bank_set_signal_cutoff(span_b200_v27ter_bank_t *bank, int first, int count, float cutoff);
/* v27ter_rx_fillin() (src/v27ter_rx.c:1030-1068) */
int span_b200_v27ter_bank_fillin(span_b200_v27ter_bank_t *bank, int first, int count, int samples);

/* v27ter_rx() (src/v27ter_rx.c:863) for every channel; device / host sample memory as in spandsp_b200.h. */
int span_b200_v27ter_bank_rx_device(span_b200_v27ter_bank_t *bank, const int16_t *d_amp, int64_t stride, int samples, void *stream);
int span_b200_v27ter_bank_rx_host(span_b200_v27ter_bank_t *bank, const int16_t *h_amp, int64_t stride, int samples, void *stream);

/* Results of the last rx call.  counts: per channel number of put_bit calls / qam reports. */
int span_b200_v27ter_bank_counts(span_b200_v27ter_bank_t *bank, int32_t *nbits, int32_t *nsyms);
int64_t span_b200_v27ter_bank_bits(span_b200_v27ter_bank_t *bank, int channel, int8_t *out, int64_t max);
/* The put_bit sequences of every channel as bytes: out + c*out_stride receives the first min(count, out_stride)
   entries of channel c, nbits (may be NULL) the per-channel counts.  Returns the largest count. */
int64_t span_b200_v27ter_bank_bits_all(span_b200_v27ter_bank_t *bank, int8_t *out, int64_t out_stride, int32_t *nbits);
int64_t span_b200_v27ter_bank_symbols(span_b200_v27ter_bank_t *bank, int channel, span_b200_v27ter_symbol_t *out, int64_t max);
/* The output of the last rx call in the form the kernel writes it, for every channel, one transfer per array:
     words   [channels][words_stride]      the data bits, 32 to a word, first bit = bit 0
     status  [channels][status_stride][2]  the status reports: {position in the put_bit sequence, SIG_STATUS_* value}
     nbits / nstatus [channels]            put_bit calls (bits + reports) / reports
   Rows are cut to the strides given; words / status may be NULL.  Returns the largest nbits, or -1.  This is the bulk
   read-back (1 bit per data bit over PCIe); *_bank_bits() and *_bank_bits_all() rebuild the byte-per-call sequence
   from it on the host. */
int64_t span_b200_v27ter_bank_output_packed(span_b200_v27ter_bank_t *bank, uint32_t *words, int64_t words_stride, int32_t *nbits,
                                        int32_t *status, int64_t status_stride, int32_t *nstatus);
/* Device-side layout of those buffers ([channel][capacity]) for callers that consume them on the GPU. */
int span_b200_v27ter_bank_output_layout(span_b200_v27ter_bank_t *bank, const uint32_t **d_words, int64_t *words_cap, const int32_t **d_nbits,
                                     const int32_t **d_status, int64_t *status_cap, const int32_t **d_nstatus,
                                     const span_b200_v29_symbol_t **d_syms, int64_t *sym_cap, const int32_t **d_nsyms);
/* eq_coeff: 32 complex taps (v27ter_rx_equalizer_state, src/v27ter_rx.c:177-189); info[12] =
   {training_stage, carrier_phase_rate, eq_put_step, signal_present, agc_scaling (float bits),
    total_baud_timing_correction, constellation_state, carrier_phase, gardner_integrate, gardner_step,
    power meter reading, bit_rate}. */
int span_b200_v27ter_bank_channel_state(span_b200_v27ter_bank_t *bank, int channel, float *eq_coeff, int32_t *info);

/* The constant tables the receiver is built on, as computed by this library's own generators (for
   verification against the reference's generated headers).  rrc4800_*: [8][27]; rrc2400_*: [12][27];
   ints: 8 = {sets at 4800, sets at 2400, carrier phase rate nominal / low / high, DDS phases 45, -45, 180 degrees}. */
int span_b200_v27ter_tables(float *rrc4800_re, float *rrc4800_im, float *rrc2400_re, float *rrc2400_im, int32_t *ints);

#if defined(__cplusplus)
}
#endif

#endif
