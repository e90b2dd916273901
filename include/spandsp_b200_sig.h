/*
 * spandsp_b200_sig.h - C ABI of the in-band signalling tone receiver banks (bulk interface): 2280 Hz (AC15 and other
 * European protocols), 2600 Hz, and 2400 Hz + 2600 Hz (SS5) - SURVEY 8(f) rank 4 (src/sig_tone.c:402-734).
 *
 * A bank = N independent sig_tone_rx receivers processed by one call; channel c reads AND REWRITES
 * d_amp[c*stride .. c*stride + samples): like the reference the receiver notches the signalling tone out of the
 * audio, mutes it, or passes it, as sig_tone_rx_set_mode() asks.  Every sig_update() callback the reference would
 * have made comes back as one record.  Float filters in the reference's expression order: audio, reports and the
 * complete receiver state are identical to the strict reference build's, bit for bit.
 */
#if !defined(_SPANDSP_B200_SIG_H_)
#define _SPANDSP_B200_SIG_H_

#include <stdint.h>

#include "spandsp_b200.h"

#if defined(__cplusplus)
extern "C"
{
#endif

typedef struct span_b200_sig_bank_s span_b200_sig_bank_t;

/* Tone types (src/spandsp/sig_tone.h:56-64) and the state / mode bits (:67-88) */
enum
{
    SPAN_B200_SIG_TONE_2280HZ = 1,
    SPAN_B200_SIG_TONE_2600HZ = 2,
    SPAN_B200_SIG_TONE_2400HZ_2600HZ = 3
};

enum
{
    SPAN_B200_SIG_TONE_1_PRESENT = 0x001,
    SPAN_B200_SIG_TONE_1_CHANGE = 0x002,
    SPAN_B200_SIG_TONE_2_PRESENT = 0x004,
    SPAN_B200_SIG_TONE_2_CHANGE = 0x008,
    SPAN_B200_SIG_TONE_RX_PASSTHROUGH = 0x040,
    SPAN_B200_SIG_TONE_RX_FILTER_TONE = 0x080
};

/* One sig_update(user_data, signalling_state, 0, duration) of the reference (src/sig_tone.c:633-640) */
typedef struct
{
    int32_t channel;
    int32_t signalling_state;       /* the *_PRESENT and *_CHANGE bits at the moment of the report */
    int32_t duration;               /* samples the previous state lasted */
} span_b200_sig_event_t;

/* sig_tone_rx_init(NULL, tone_type, ...) x channels (src/sig_tone.c:672-722).  NULL for a tone type outside 1..3. */
span_b200_sig_bank_t *span_b200_sig_bank_create(span_b200_ctx_t *ctx, int channels, int tone_type);
void span_b200_sig_bank_destroy(span_b200_sig_bank_t *bank);
int span_b200_sig_bank_channels(const span_b200_sig_bank_t *bank);
/* sig_tone_rx_init() again for channels [first, first+count); channels of one bank may use different tone types */
int span_b200_sig_bank_init(span_b200_sig_bank_t *bank, int first, int count, int tone_type);
/* sig_tone_rx_set_mode() (src/sig_tone.c:666-669): SPAN_B200_SIG_TONE_RX_PASSTHROUGH / _FILTER_TONE, or 0 = mute */
int span_b200_sig_bank_set_mode(span_b200_sig_bank_t *bank, int first, int count, int mode);

/* sig_tone_rx() (src/sig_tone.c:402-664) for every channel, in place; device / host sample memory as in
   spandsp_b200.h (the host form copies the processed audio back). */
int span_b200_sig_bank_rx_device(span_b200_sig_bank_t *bank, int16_t *d_amp, int64_t stride, int samples, void *stream);
int span_b200_sig_bank_rx_host(span_b200_sig_bank_t *bank, int16_t *h_amp, int64_t stride, int samples, void *stream);

/* The reports of the last rx call, ordered by channel and, within a channel, in time.  Returns their number (which
   may exceed max; only max are written), or -1. */
int64_t span_b200_sig_bank_events(span_b200_sig_bank_t *bank, span_b200_sig_event_t *events, int64_t max);
/* info[31]: the receiver's state in the order of sb_sig_rx.cuh's T_* fields (src/spandsp/private/sig_tone.h:150-205):
   tone_type, current_rx_tone, current_notch_filter, notch_z1[3][2], notch_z2[3][2] (floats as bit patterns),
   tone[].power[3], flat_z[2], flat_power, tone_persistence_timeout, last_sample_tone_present, flat / sharp
   detection thresholds, detection_ratio, flat_mode, flat_mode_timeout, notch_insertion_timeout, signalling_state,
   signalling_state_duration. */
int span_b200_sig_bank_channel_state(span_b200_sig_bank_t *bank, int channel, int32_t *info);

#if defined(__cplusplus)
}
#endif

#endif
