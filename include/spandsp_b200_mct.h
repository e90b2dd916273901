/*
 * spandsp_b200_mct.h - C ABI of the modem connect tone detector banks (bulk interface): FAX CNG, ANS / CED with its
 * phase-reversal and AM variants, Bell ANS, calling tone, and the V.21 FAX preamble (the
 * detector a FAX / V.8 front end runs before it starts its modems, src/modem_connect_tones.c:521).
 *
 * A bank = N independent modem_connect_tones_rx receivers (src/modem_connect_tones.c:419-804) processed by one
 * call; channel c reads d_amp[c*stride .. c*stride + samples).  One call of span_b200_mct_bank_rx_*() is one call of
 * modem_connect_tones_rx() per channel - including the reference's order inside a call for the CED-or-preamble
 * type (the V.21 receiver sees the whole buffer before the 2100 Hz detector does, :573-580).  Every report the
 * reference would have made through its tone_callback comes back as one record; the `hit` a state without a
 * callback accumulates (:428-431) is kept too and fetched with span_b200_mct_bank_get().
 * Filter arithmetic is IEEE single precision in the reference's operation order; results are identical to the
 * strict reference build's, bit for bit (reports, levels and the complete detector state).
 */
#if !defined(_SPANDSP_B200_MCT_H_)
#define _SPANDSP_B200_MCT_H_

#include <stdint.h>

#include "spandsp_b200.h"

#if defined(__cplusplus)
extern "C"
{
#endif

typedef struct span_b200_mct_bank_s span_b200_mct_bank_t;

/* The reference's tone codes (src/spandsp/modem_connect_tones.h:57-90) */
enum
{
    SPAN_B200_MCT_NONE = 0,
    SPAN_B200_MCT_FAX_CNG = 1,
    SPAN_B200_MCT_ANS = 2,
    SPAN_B200_MCT_ANS_PR = 3,
    SPAN_B200_MCT_ANSAM = 4,
    SPAN_B200_MCT_ANSAM_PR = 5,
    SPAN_B200_MCT_FAX_PREAMBLE = 6,
    SPAN_B200_MCT_FAX_CED_OR_PREAMBLE = 7,
    SPAN_B200_MCT_BELL_ANS = 8,
    SPAN_B200_MCT_CALLING_TONE = 9,
    SPAN_B200_MCT_REAL_TIME_REPORTS = 0x1000
};

/* One tone_callback(user_data, tone, level, 0) of the reference (src/modem_connect_tones.c:423-425) */
typedef struct
{
    int32_t channel;
    int32_t tone;
    int32_t level;              /* dBm0, or -99 when the tone ends */
} span_b200_mct_event_t;

/* modem_connect_tones_rx_init(NULL, tone_type, ...) x channels (src/modem_connect_tones.c:823-875).  As in the
   reference a tone type the receiver does not know is accepted and detects nothing. */
span_b200_mct_bank_t *span_b200_mct_bank_create(span_b200_ctx_t *ctx, int channels, int tone_type);
void span_b200_mct_bank_destroy(span_b200_mct_bank_t *bank);
int span_b200_mct_bank_channels(const span_b200_mct_bank_t *bank);
/* modem_connect_tones_rx_init(s, tone_type, ...) again for channels [first, first+count); channels of one bank
   may look for different tones. */
int span_b200_mct_bank_init(span_b200_mct_bank_t *bank, int first, int count, int tone_type);

/* modem_connect_tones_rx() (src/modem_connect_tones.c:521-804) for every channel; device / host sample memory as
   in spandsp_b200.h. */
int span_b200_mct_bank_rx_device(span_b200_mct_bank_t *bank, const int16_t *d_amp, int64_t stride, int samples, void *stream);
int span_b200_mct_bank_rx_host(span_b200_mct_bank_t *bank, const int16_t *h_amp, int64_t stride, int samples, void *stream);

/* The reports of the last rx call, ordered by channel and, within a channel, in the order the reference would have
   made them.  Returns their number (which may exceed max; only max are written), or -1. */
int64_t span_b200_mct_bank_events(span_b200_mct_bank_t *bank, span_b200_mct_event_t *events, int64_t max);
/* modem_connect_tones_rx_get() (src/modem_connect_tones.c:812-820) for channels [first, first+count): the last tone
   declared since the previous get, then cleared. */
int span_b200_mct_bank_get(span_b200_mct_bank_t *bank, int first, int count, int32_t *hits);
/* info[17]: tone_type, notch_level, channel_level, am_level, tone_present, tone_on, tone_cycle_duration, good_cycles,
   raw_bit_stream, num_bits, flags_seen, framing_ok_announced, znotch_1, znotch_2, z15hz_1, z15hz_2 (floats as their
   bit patterns), hit (src/spandsp/private/modem_connect_tones.h:57-103); fsk_info (may be NULL): the 28 integers
   of the embedded V.21 receiver as span_b200_fsk_bank_channel_state() gives them. */
int span_b200_mct_bank_channel_state(span_b200_mct_bank_t *bank, int channel, int32_t *info, int32_t *fsk_info);

#if defined(__cplusplus)
}
#endif

#endif
