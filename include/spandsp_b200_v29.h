/*
 * spandsp_b200_v29.h - C ABI of the V.29 receiver banks (bulk interface).
 *
 * A bank = N independent V.29 receivers (src/v29rx.c) processed by one call; channel c reads
 * d_amp[c*stride .. c*stride + samples).  What the reference delivers through callbacks is returned
 * as per-channel streams:
 *   - the put_bit stream (span_put_bit_func_t, src/spandsp/async.h:123): one int8 per call, 0/1 for
 *     data bits and the negative SIG_STATUS_* codes (async.h:66-103) exactly where the reference would
 *     have delivered them (no separate status handler: src/v29rx.c:171-178);
 *   - optionally the equalized symbols of qam_report_handler_t (src/spandsp/v29rx.h:130).
 */
#if !defined(_SPANDSP_B200_V29_H_)
#define _SPANDSP_B200_V29_H_

#include <stdint.h>

#include "spandsp_b200.h"

#if defined(__cplusplus)
extern "C"
{
#endif

typedef struct span_b200_v29_bank_s span_b200_v29_bank_t;

/* One qam_report: equalizer output z, the constellation target, and constellation_state. */
typedef struct
{
    float re;
    float im;
    float target_re;
    float target_im;
    int32_t state;
    int32_t bit_pos;        /* number of put_bit calls of this rx call that preceded the report */
} span_b200_v29_symbol_t;

/* v29_rx_init(NULL, bit_rate, ...) x channels (src/v29rx.c:1100-1134).  bit_rate: 9600, 7200 or 4800;
   anything else fails as in the reference.  want_symbols != 0 records the qam_report stream. */
span_b200_v29_bank_t *span_b200_v29_bank_create(span_b200_ctx_t *ctx, int channels, int bit_rate, int want_symbols);
void span_b200_v29_bank_destroy(span_b200_v29_bank_t *bank);
int span_b200_v29_bank_channels(const span_b200_v29_bank_t *bank);
/* v29_rx_restart(s, bit_rate, false) (src/v29rx.c:1019) for channels [first, first+count). */
int span_b200_v29_bank_restart(span_b200_v29_bank_t *bank, int first, int count, int bit_rate);
/* v29_rx_restart(s, bit_rate, old_train): old_train != 0 reuses the saved equalizer, carrier and gain
   (src/v29rx.c:1064-1069). */
int span_b200_v29_bank_restart_ex(span_b200_v29_bank_t *bank, int first, int count, int bit_rate, int old_train);
/* v29_rx_set_signal_cutoff() (src/v29rx.c:163-168) */
int span_b200_v29_bank_set_signal_cutoff(span_b200_v29_bank_t *bank, int first, int count, float cutoff);
/* v29_rx_fillin() (src/v29rx.c:967-996): sustain carrier phase and symbol timing over `samples` lost samples. */
int span_b200_v29_bank_fillin(span_b200_v29_bank_t *bank, int first, int count, int samples);

/* v29_rx() (src/v29rx.c:867) for every channel; device / host sample memory as in spandsp_b200.h. */
int span_b200_v29_bank_rx_device(span_b200_v29_bank_t *bank, const int16_t *d_amp, int64_t stride, int samples, void *stream);
int span_b200_v29_bank_rx_host(span_b200_v29_bank_t *bank, const int16_t *h_amp, int64_t stride, int samples, void *stream);

/* Results of the last rx call.  counts: per channel number of put_bit calls / qam reports. */
int span_b200_v29_bank_counts(span_b200_v29_bank_t *bank, int32_t *nbits, int32_t *nsyms);
int64_t span_b200_v29_bank_bits(span_b200_v29_bank_t *bank, int channel, int8_t *out, int64_t max);
/* The put_bit streams of every channel in one transfer: out + c*out_stride receives the first
   min(count, out_stride) entries of channel c, nbits (may be NULL) the per-channel counts.  Returns the largest count. */
int64_t span_b200_v29_bank_bits_all(span_b200_v29_bank_t *bank, int8_t *out, int64_t out_stride, int32_t *nbits);
int64_t span_b200_v29_bank_symbols(span_b200_v29_bank_t *bank, int channel, span_b200_v29_symbol_t *out, int64_t max);
/* The output of the last rx call in the form the kernel writes it, for every channel, one transfer per array:
     words   [channels][words_stride]      the data bits, 32 to a word, first bit = bit 0
     status  [channels][status_stride][2]  the status reports: {position in the put_bit sequence, SIG_STATUS_* value}
     nbits / nstatus [channels]            put_bit calls (bits + reports) / reports
   Rows are cut to the strides given; words / status may be NULL.  Returns the largest nbits, or -1.  This is the bulk
   read-back (1 bit per data bit over PCIe); *_bank_bits() and *_bank_bits_all() rebuild the byte-per-call sequence
   from it on the host. */
int64_t span_b200_v29_bank_output_packed(span_b200_v29_bank_t *bank, uint32_t *words, int64_t words_stride, int32_t *nbits,
                                        int32_t *status, int64_t status_stride, int32_t *nstatus);
/* Device-side layout of those buffers ([channel][capacity]) for callers that consume them on the GPU. */
int span_b200_v29_bank_output_layout(span_b200_v29_bank_t *bank, const uint32_t **d_words, int64_t *words_cap, const int32_t **d_nbits,
                                     const int32_t **d_status, int64_t *status_cap, const int32_t **d_nstatus,
                                     const span_b200_v29_symbol_t **d_syms, int64_t *sym_cap, const int32_t **d_nsyms);
/* eq_coeff: 33 complex taps (v29_rx_equalizer_state, src/v29rx.c:180-195); info[10] =
   {training_stage, carrier_phase_rate, eq_put_step, signal_present, agc_scaling (float bits),
    total_baud_timing_correction, constellation_state, carrier_phase, power meter reading, bit_rate}. */
int span_b200_v29_bank_channel_state(span_b200_v29_bank_t *bank, int channel, float *eq_coeff, int32_t *info);

/* The constant tables the receiver is built on, as computed by this library's own generators
   (for verification against the reference's generated headers). rrc_*: [48][27]; sine: [2048];
   sqrt_tab: [193]; godard: 9 floats; ints: 9 (see sb_v29.cu). */
int span_b200_v29_tables(float *rrc_re, float *rrc_im, float *sine, uint16_t *sqrt_tab, float *godard, int32_t *ints);

#if defined(__cplusplus)
}
#endif

#endif
