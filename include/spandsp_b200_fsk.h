/*
 * spandsp_b200_fsk.h - C ABI of the FSK receiver banks (bulk interface): V.21, V.23, Bell 103, Bell 202,
 * Weitbrecht - the V.21 channel 2 receiver is the one a FAX front end runs beside its fast modem
 * (src/fax_modems.c:207-330).
 *
 * A bank = N independent fsk_rx receivers (src/fsk.c:396-626) processed by one call; channel c reads
 * d_amp[c*stride .. c*stride + samples).  What the reference delivers through put_bit (and, absent a status
 * handler, its status reports - src/fsk.c:347-354) is returned as one int16 stream per channel: 0/1 for the
 * bits of the synchronous and asynchronous modes, the character value in framed mode, and the negative
 * SIG_STATUS_* codes (src/spandsp/async.h:66-103) exactly where the reference would have delivered them.
 * Everything in this receiver is integer arithmetic: results are identical to the reference's, bit for bit.
 */
#if !defined(_SPANDSP_B200_FSK_H_)
#define _SPANDSP_B200_FSK_H_

#include <stdint.h>

#include "spandsp_b200.h"

#if defined(__cplusplus)
extern "C"
{
#endif

typedef struct span_b200_fsk_bank_s span_b200_fsk_bank_t;

/* fsk_spec_t (src/spandsp/fsk.h:92-106) */
typedef struct
{
    const char *name;
    int freq_zero;
    int freq_one;
    int tx_level;
    int min_level;
    int baud_rate;          /* baud rate x 100 */
} span_b200_fsk_spec_t;

/* Indices into span_b200_fsk_presets[] = the reference's preset_fsk_specs[] (src/fsk.c:60-156, fsk.h:108-121) */
enum
{
    SPAN_B200_FSK_V21CH1 = 0,
    SPAN_B200_FSK_V21CH2,
    SPAN_B200_FSK_V23CH1,
    SPAN_B200_FSK_V23CH2,
    SPAN_B200_FSK_BELL103CH1,
    SPAN_B200_FSK_BELL103CH2,
    SPAN_B200_FSK_BELL202,
    SPAN_B200_FSK_WEITBRECHT_4545,
    SPAN_B200_FSK_WEITBRECHT_50,
    SPAN_B200_FSK_WEITBRECHT_476,
    SPAN_B200_FSK_V21CH1_110,
    SPAN_B200_FSK_PRESETS
};
const span_b200_fsk_spec_t *span_b200_fsk_preset(int which);

/* Framing modes (src/spandsp/fsk.h:124-129) and parity (src/spandsp/async.h:149-158) */
enum
{
    SPAN_B200_FSK_FRAME_MODE_ASYNC = 0,
    SPAN_B200_FSK_FRAME_MODE_SYNC = 1,
    SPAN_B200_FSK_FRAME_MODE_FRAMED = 2
};

/* fsk_rx_init(NULL, spec, framing_mode, ...) x channels (src/fsk.c:725-744). */
span_b200_fsk_bank_t *span_b200_fsk_bank_create(span_b200_ctx_t *ctx, int channels, const span_b200_fsk_spec_t *spec, int framing_mode);
void span_b200_fsk_bank_destroy(span_b200_fsk_bank_t *bank);
int span_b200_fsk_bank_channels(const span_b200_fsk_bank_t *bank);
/* fsk_rx_restart(s, spec, framing_mode) (src/fsk.c:670-722) for channels [first, first+count); channels of one
   bank may run different specs.  As in the reference, a restart does not clear the correlation window. */
int span_b200_fsk_bank_restart(span_b200_fsk_bank_t *bank, int first, int count, const span_b200_fsk_spec_t *spec, int framing_mode);
/* fsk_rx_set_signal_cutoff() (src/fsk.c:271-277) */
int span_b200_fsk_bank_set_signal_cutoff(span_b200_fsk_bank_t *bank, int first, int count, float cutoff);
/* fsk_rx_set_frame_parameters() (src/fsk.c:300-316); framed mode only, total bits per character <= 15 */
int span_b200_fsk_bank_set_frame_parameters(span_b200_fsk_bank_t *bank, int first, int count, int data_bits, int parity, int stop_bits);
/* fsk_rx_fillin() (src/fsk.c:628-667) */
int span_b200_fsk_bank_fillin(span_b200_fsk_bank_t *bank, int first, int count, int samples);

/* fsk_rx() (src/fsk.c:396) for every channel; device / host sample memory as in spandsp_b200.h. */
int span_b200_fsk_bank_rx_device(span_b200_fsk_bank_t *bank, const int16_t *d_amp, int64_t stride, int samples, void *stream);
int span_b200_fsk_bank_rx_host(span_b200_fsk_bank_t *bank, const int16_t *h_amp, int64_t stride, int samples, void *stream);

/* Results of the last rx call: per channel number of put_bit calls, and one channel's stream. */
int span_b200_fsk_bank_counts(span_b200_fsk_bank_t *bank, int32_t *nout);
int64_t span_b200_fsk_bank_output(span_b200_fsk_bank_t *bank, int channel, int16_t *out, int64_t max);
/* Device-side layout of the result buffer ([channel][capacity]) for callers that consume it on the GPU. */
int span_b200_fsk_bank_output_layout(span_b200_fsk_bank_t *bank, const int16_t **d_out, int64_t *out_cap, const int32_t **d_nout);
/* fsk_rx_get_parity_errors() / fsk_rx_get_framing_errors() (src/fsk.c:319-344) */
int span_b200_fsk_bank_errors(span_b200_fsk_bank_t *bank, int channel, int32_t *parity_errors, int32_t *framing_errors, int reset);
/* fsk_rx_signal_power() (src/fsk.c:280-283) */
float span_b200_fsk_bank_signal_power(span_b200_fsk_bank_t *bank, int channel);
/* info[28]: the receiver's integer state in the order of src/spandsp/private/fsk.h:55-113 (see sb_fsk_rx.cuh K_*);
   window (may be NULL): 2 x 128 complex int32 correlation window entries, [tone][slot][re, im]. */
int span_b200_fsk_bank_channel_state(span_b200_fsk_bank_t *bank, int channel, int32_t *info, int32_t *window);

/* The integer DDS quarter-wave table (257 entries) as computed by this library (src/dds_int.c:55-315 holds the
   same numbers as literals), for verification. */
int span_b200_dds_int_table(int16_t *table);

#if defined(__cplusplus)
}
#endif

#endif
