/*
 * spandsp_b200_dropin.h - the spandsp-named per-channel API of the tone-detect path, served by
 * the B200 engine.
 *
 * Same function names, argument meaning, callback types, ownership and return conventions as the
 * reference headers cited at each group, so a caller of that path links unchanged.  State objects
 * are opaque here (the reference keeps their bodies in spandsp/private/ too); `*_init(NULL, ...)`
 * allocates, `*_init(s, ...)` re-initialises in place, and sizeof(the reference's struct) bytes of
 * caller storage are always enough (ours are smaller).
 *
 * Execution model.  Every state object is a slot of a GROUP.  A state made by `*_init()` alone is
 * a group of one and behaves synchronously: `dtmf_rx()` copies the samples to the GPU, runs the
 * kernels and fires the callbacks before it returns - exactly the reference's observable
 * behaviour, at batch-of-one cost.  For real channel counts create a group
 * (span_b200_group_create), hand its members to your per-channel code, and call
 * span_b200_group_flush() once per tick: `*_rx()` on a member then only stages the samples, and the
 * flush processes all members in one launch and fires every callback, in channel order, on the
 * flushing thread.  There is no CPU implementation behind any of these calls.
 */
#if !defined(_SPANDSP_B200_DROPIN_H_)
#define _SPANDSP_B200_DROPIN_H_

#include <stdint.h>
#include <stddef.h>
#include <stdbool.h>

#include "spandsp_b200.h"

#if defined(__cplusplus)
extern "C"
{
#endif

/* ---- callback types (src/spandsp/dtmf.h:76, src/spandsp/super_tone_rx.h:56-58 and tone_detect users) */
typedef void (*digits_rx_callback_t)(void *user_data, const char *digits, int len);
typedef void (*span_tone_report_func_t)(void *user_data, int code, int level, int delay);
typedef void (*tone_segment_func_t)(void *user_data, int f1, int f2, int duration);

typedef struct dtmf_rx_state_s dtmf_rx_state_t;
typedef struct bell_mf_rx_state_s bell_mf_rx_state_t;
typedef struct r2_mf_rx_state_s r2_mf_rx_state_t;
typedef struct super_tone_rx_state_s super_tone_rx_state_t;
typedef struct super_tone_rx_descriptor_s super_tone_rx_descriptor_t;
typedef struct logging_state_s logging_state_t;

#define MAX_DTMF_DIGITS     128     /* src/spandsp/dtmf.h:74 */
#define MAX_BELL_MF_DIGITS  128     /* src/spandsp/bell_r2_mf.h */

/* ---- groups (new; see "Execution model" above) ------------------------------------------------ */
typedef struct span_b200_group_s span_b200_group_t;

/* detector: SPAN_B200_DET_*; arg: R2 MF forward flag, or for super-tone a super_tone_rx_descriptor_t
   passed through `desc`.  max_samples sizes the pinned staging area (samples one member stages between flushes); it grows when a
   member is fed more than that, audio is never dropped. */
span_b200_group_t *span_b200_group_create(span_b200_ctx_t *ctx, int detector, int members, int max_samples,
                                          int arg, super_tone_rx_descriptor_t *desc);
/* Member i as the detector's state type (cast to dtmf_rx_state_t * etc.).  Members are created in
   the *_rx_init() state with no callbacks; install callbacks with *_rx_init(member, cb, user). */
void *span_b200_group_member(span_b200_group_t *group, int index);
/* Process what the members staged since the last flush (all members that staged anything must have
   staged the same number of samples; members that staged nothing are fed nothing) and fire the
   callbacks.  Returns the number of callbacks fired, or -1. */
int span_b200_group_flush(span_b200_group_t *group);
void span_b200_group_destroy(span_b200_group_t *group);
/* The default context used by states created without a group (device from SPANDSP_B200_DEVICE, else 0). */
span_b200_ctx_t *span_b200_default_ctx(void);

/* ---- DTMF receiver: src/spandsp/dtmf.h:153-228, src/dtmf.c:132-519 ---------------------------- */
dtmf_rx_state_t *dtmf_rx_init(dtmf_rx_state_t *s, digits_rx_callback_t callback, void *user_data);
int dtmf_rx_release(dtmf_rx_state_t *s);
int dtmf_rx_free(dtmf_rx_state_t *s);
void dtmf_rx_set_realtime_callback(dtmf_rx_state_t *s, span_tone_report_func_t callback, void *user_data);
void dtmf_rx_parms(dtmf_rx_state_t *s, int filter_dialtone, float twist, float reverse_twist, float threshold);
int dtmf_rx(dtmf_rx_state_t *s, const int16_t amp[], int samples);
int dtmf_rx_fillin(dtmf_rx_state_t *s, int samples);
int dtmf_rx_status(dtmf_rx_state_t *s);
size_t dtmf_rx_get(dtmf_rx_state_t *s, char *digits, int max);
logging_state_t *dtmf_rx_get_logging_state(dtmf_rx_state_t *s);

/* ---- Bell MF receiver: src/spandsp/bell_r2_mf.h:199-228, src/bell_r2_mf.c:507-745 -------------- */
bell_mf_rx_state_t *bell_mf_rx_init(bell_mf_rx_state_t *s, digits_rx_callback_t callback, void *user_data);
int bell_mf_rx_release(bell_mf_rx_state_t *s);
int bell_mf_rx_free(bell_mf_rx_state_t *s);
int bell_mf_rx(bell_mf_rx_state_t *s, const int16_t amp[], int samples);
size_t bell_mf_rx_get(bell_mf_rx_state_t *s, char *buf, int max);

/* ---- MFC/R2 receiver: src/spandsp/bell_r2_mf.h:236-266, src/bell_r2_mf.c:750-951 --------------- */
r2_mf_rx_state_t *r2_mf_rx_init(r2_mf_rx_state_t *s, bool fwd, span_tone_report_func_t callback, void *user_data);
int r2_mf_rx_release(r2_mf_rx_state_t *s);
int r2_mf_rx_free(r2_mf_rx_state_t *s);
int r2_mf_rx(r2_mf_rx_state_t *s, const int16_t amp[], int samples);
int r2_mf_rx_get(r2_mf_rx_state_t *s);

/* ---- supervisory tone receiver: src/spandsp/super_tone_rx.h:76-164, src/super_tone_rx.c ---------- */
super_tone_rx_descriptor_t *super_tone_rx_make_descriptor(super_tone_rx_descriptor_t *desc);
int super_tone_rx_free_descriptor(super_tone_rx_descriptor_t *desc);
int super_tone_rx_add_tone(super_tone_rx_descriptor_t *desc);
int super_tone_rx_add_element(super_tone_rx_descriptor_t *desc, int tone, int f1, int f2, int min, int max);
super_tone_rx_state_t *super_tone_rx_init(super_tone_rx_state_t *s, super_tone_rx_descriptor_t *desc,
                                          span_tone_report_func_t callback, void *user_data);
int super_tone_rx_release(super_tone_rx_state_t *s);
int super_tone_rx_free(super_tone_rx_state_t *s);
void super_tone_rx_tone_callback(super_tone_rx_state_t *s, span_tone_report_func_t callback, void *user_data);
void super_tone_rx_segment_callback(super_tone_rx_state_t *s, tone_segment_func_t callback);
int super_tone_rx(super_tone_rx_state_t *s, const int16_t amp[], int samples);
int super_tone_rx_fillin(super_tone_rx_state_t *s, int samples);

/* ---- Goertzel primitives: src/spandsp/tone_detect.h:32-58,86-127, src/tone_detect.c:60-205 ------ */
/* These two structures are public in the reference (callers embed them and the header-inline
   goertzel_sample()/goertzel_samplex() touch the fields), so their layout is kept. */
typedef struct
{
    float fac;
    int samples;
} goertzel_descriptor_t;

typedef struct
{
    float v2;
    float v3;
    float fac;
    int samples;
    int current_sample;
} goertzel_state_t;

void make_goertzel_descriptor(goertzel_descriptor_t *t, float freq, int samples);
goertzel_state_t *goertzel_init(goertzel_state_t *s, goertzel_descriptor_t *t);
int goertzel_release(goertzel_state_t *s);
int goertzel_free(goertzel_state_t *s);
void goertzel_reset(goertzel_state_t *s);
/* Both run on the device as a batch of one (state and samples go up, state comes back). */
int goertzel_update(goertzel_state_t *s, const int16_t amp[], int samples);
float goertzel_result(goertzel_state_t *s);

/* ---- what the modem receivers report through put_bit / the status handler: src/spandsp/async.h:62-103 ---- */
enum
{
    SIG_STATUS_CARRIER_DOWN = -1,
    SIG_STATUS_CARRIER_UP = -2,
    SIG_STATUS_TRAINING_IN_PROGRESS = -3,
    SIG_STATUS_TRAINING_SUCCEEDED = -4,
    SIG_STATUS_TRAINING_FAILED = -5,
    SIG_STATUS_FRAMING_OK = -6,
    SIG_STATUS_END_OF_DATA = -7,
    SIG_STATUS_ABORT = -8,
    SIG_STATUS_BREAK = -9,
    SIG_STATUS_SHUTDOWN_COMPLETE = -10,
    SIG_STATUS_OCTET_REPORT = -11,
    SIG_STATUS_POOR_SIGNAL_QUALITY = -12,
    SIG_STATUS_MODEM_RETRAIN_OCCURRED = -13,
    SIG_STATUS_LINK_CONNECTED = -14,
    SIG_STATUS_LINK_DISCONNECTED = -15,
    SIG_STATUS_LINK_ERROR = -16,
    SIG_STATUS_LINK_IDLE = -17
};

#if !defined(SAMPLE_RATE)
#define SAMPLE_RATE         8000    /* src/spandsp/telephony.h:45 */
#endif

/* ---- V.29 receiver: src/spandsp/v29rx.h:130-244, src/v29rx.c:145-195,867-1153 -------------------- */
typedef struct
{
    float re;
    float im;
} complexf_t;                                                           /* src/spandsp/complex.h:42-48 */
typedef void (*span_put_bit_func_t)(void *user_data, int bit);          /* src/spandsp/async.h:123 */
typedef void (*span_modem_status_func_t)(void *user_data, int status);  /* src/spandsp/async.h:131 */
typedef void (*qam_report_handler_t)(void *user_data, const complexf_t *constel, const complexf_t *target, int symbol);
typedef struct v29_rx_state_s v29_rx_state_t;

/* Synchronous, one receiver per state (a bank of one).  Banks of many receivers: spandsp_b200_v29.h. */
v29_rx_state_t *v29_rx_init(v29_rx_state_t *s, int bit_rate, span_put_bit_func_t put_bit, void *user_data);
int v29_rx_restart(v29_rx_state_t *s, int bit_rate, bool old_train);
int v29_rx_release(v29_rx_state_t *s);
int v29_rx_free(v29_rx_state_t *s);
logging_state_t *v29_rx_get_logging_state(v29_rx_state_t *s);
void v29_rx_set_put_bit(v29_rx_state_t *s, span_put_bit_func_t put_bit, void *user_data);
void v29_rx_set_modem_status_handler(v29_rx_state_t *s, span_modem_status_func_t handler, void *user_data);
int v29_rx(v29_rx_state_t *s, const int16_t amp[], int len);
int v29_rx_fillin(v29_rx_state_t *s, int len);
int v29_rx_equalizer_state(v29_rx_state_t *s, complexf_t **coeffs);
float v29_rx_carrier_frequency(v29_rx_state_t *s);
float v29_rx_symbol_timing_correction(v29_rx_state_t *s);
float v29_rx_signal_power(v29_rx_state_t *s);
void v29_rx_set_signal_cutoff(v29_rx_state_t *s, float cutoff);
void v29_rx_set_qam_report_handler(v29_rx_state_t *s, qam_report_handler_t handler, void *user_data);

/* ---- V.17 receiver: src/spandsp/v17rx.h:236-333, src/v17rx.c:157-204,1214-1541 -------------------- */
typedef struct v17_rx_state_s v17_rx_state_t;

/* Synchronous, one receiver per state (a bank of one).  Banks of many receivers: spandsp_b200_v17.h. */
v17_rx_state_t *v17_rx_init(v17_rx_state_t *s, int bit_rate, span_put_bit_func_t put_bit, void *user_data);
int v17_rx_restart(v17_rx_state_t *s, int bit_rate, int short_train);
int v17_rx_release(v17_rx_state_t *s);
int v17_rx_free(v17_rx_state_t *s);
logging_state_t *v17_rx_get_logging_state(v17_rx_state_t *s);
void v17_rx_set_put_bit(v17_rx_state_t *s, span_put_bit_func_t put_bit, void *user_data);
void v17_rx_set_modem_status_handler(v17_rx_state_t *s, span_modem_status_func_t handler, void *user_data);
int v17_rx(v17_rx_state_t *s, const int16_t amp[], int len);
int v17_rx_fillin(v17_rx_state_t *s, int len);
int v17_rx_equalizer_state(v17_rx_state_t *s, complexf_t **coeffs);
float v17_rx_carrier_frequency(v17_rx_state_t *s);
float v17_rx_symbol_timing_correction(v17_rx_state_t *s);
float v17_rx_signal_power(v17_rx_state_t *s);
void v17_rx_set_signal_cutoff(v17_rx_state_t *s, float cutoff);
void v17_rx_set_qam_report_handler(v17_rx_state_t *s, qam_report_handler_t handler, void *user_data);

/* ---- V.27ter receiver: src/spandsp/v27ter_rx.h:71-165, src/v27ter_rx.c:137-189,863-1210 ---------- */
typedef struct v27ter_rx_state_s v27ter_rx_state_t;

/* Synchronous, one receiver per state (a bank of one).  Banks of many receivers: spandsp_b200_v27ter.h.
   The qam report handler also receives the Gardner timing hops (NULL, NULL, integrator value), as in the reference. */
v27ter_rx_state_t *v27ter_rx_init(v27ter_rx_state_t *s, int bit_rate, span_put_bit_func_t put_bit, void *user_data);
int v27ter_rx_restart(v27ter_rx_state_t *s, int bit_rate, bool old_train);
int v27ter_rx_release(v27ter_rx_state_t *s);
int v27ter_rx_free(v27ter_rx_state_t *s);
logging_state_t *v27ter_rx_get_logging_state(v27ter_rx_state_t *s);
void v27ter_rx_set_put_bit(v27ter_rx_state_t *s, span_put_bit_func_t put_bit, void *user_data);
void v27ter_rx_set_modem_status_handler(v27ter_rx_state_t *s, span_modem_status_func_t handler, void *user_data);
int v27ter_rx(v27ter_rx_state_t *s, const int16_t amp[], int len);
int v27ter_rx_fillin(v27ter_rx_state_t *s, int len);
int v27ter_rx_equalizer_state(v27ter_rx_state_t *s, complexf_t **coeffs);
float v27ter_rx_carrier_frequency(v27ter_rx_state_t *s);
float v27ter_rx_symbol_timing_correction(v27ter_rx_state_t *s);
float v27ter_rx_signal_power(v27ter_rx_state_t *s);
void v27ter_rx_set_signal_cutoff(v27ter_rx_state_t *s, float cutoff);
void v27ter_rx_set_qam_report_handler(v27ter_rx_state_t *s, qam_report_handler_t handler, void *user_data);

/* ---- FSK receiver (V.21, V.23, Bell 103/202, Weitbrecht): src/spandsp/fsk.h:92-269, src/fsk.c:60-156,271-760 ---- */
typedef struct
{
    const char *name;
    int freq_zero;
    int freq_one;
    int tx_level;
    int min_level;
    int baud_rate;
} fsk_spec_t;                                                           /* src/spandsp/fsk.h:92-106 */

enum
{
    FSK_V21CH1 = 0, FSK_V21CH2, FSK_V23CH1, FSK_V23CH2, FSK_BELL103CH1, FSK_BELL103CH2, FSK_BELL202,
    FSK_WEITBRECHT_4545, FSK_WEITBRECHT_50, FSK_WEITBRECHT_476, FSK_V21CH1_110
};                                                                      /* src/spandsp/fsk.h:108-121 */

enum
{
    FSK_FRAME_MODE_ASYNC = 0,
    FSK_FRAME_MODE_SYNC = 1,
    FSK_FRAME_MODE_FRAMED = 2
};                                                                      /* src/spandsp/fsk.h:124-129 */

extern const fsk_spec_t preset_fsk_specs[];                             /* src/spandsp/fsk.h:131 */

typedef struct fsk_rx_state_s fsk_rx_state_t;

/* Synchronous, one receiver per state (a bank of one).  Banks of many receivers: spandsp_b200_fsk.h. */
fsk_rx_state_t *fsk_rx_init(fsk_rx_state_t *s, const fsk_spec_t *spec, int framing_mode, span_put_bit_func_t put_bit, void *user_data);
int fsk_rx_restart(fsk_rx_state_t *s, const fsk_spec_t *spec, int framing_mode);
int fsk_rx_release(fsk_rx_state_t *s);
int fsk_rx_free(fsk_rx_state_t *s);
int fsk_rx(fsk_rx_state_t *s, const int16_t *amp, int len);
int fsk_rx_fillin(fsk_rx_state_t *s, int len);
void fsk_rx_set_put_bit(fsk_rx_state_t *s, span_put_bit_func_t put_bit, void *user_data);
void fsk_rx_set_modem_status_handler(fsk_rx_state_t *s, span_modem_status_func_t handler, void *user_data);
void fsk_rx_set_signal_cutoff(fsk_rx_state_t *s, float cutoff);
float fsk_rx_signal_power(fsk_rx_state_t *s);
void fsk_rx_set_frame_parameters(fsk_rx_state_t *s, int data_bits, int parity, int stop_bits);
int fsk_rx_get_parity_errors(fsk_rx_state_t *s, bool reset);
int fsk_rx_get_framing_errors(fsk_rx_state_t *s, bool reset);

/* ---- Modem connect tone detector (FAX CNG, CED / ANS and its variants, Bell ANS, calling tone, V.21 preamble):
        src/spandsp/modem_connect_tones.h:57-188, src/modem_connect_tones.c:74-118,419-892 ---- */
enum
{
    MODEM_CONNECT_TONES_NONE = 0,
    MODEM_CONNECT_TONES_FAX_CNG = 1,
    MODEM_CONNECT_TONES_ANS = 2,
    MODEM_CONNECT_TONES_ANS_PR = 3,
    MODEM_CONNECT_TONES_ANSAM = 4,
    MODEM_CONNECT_TONES_ANSAM_PR = 5,
    MODEM_CONNECT_TONES_FAX_PREAMBLE = 6,
    MODEM_CONNECT_TONES_FAX_CED_OR_PREAMBLE = 7,
    MODEM_CONNECT_TONES_BELL_ANS = 8,
    MODEM_CONNECT_TONES_CALLING_TONE = 9,
    MODEM_CONNECT_TONES_REAL_TIME_REPORTS = 0x1000
};                                                                      /* src/spandsp/modem_connect_tones.h:57-90 */
#define MODEM_CONNECT_TONES_FAX_CED MODEM_CONNECT_TONES_ANS             /* src/spandsp/modem_connect_tones.h:93 */

typedef struct modem_connect_tones_rx_state_s modem_connect_tones_rx_state_t;

/* Synchronous, one detector per state (a bank of one).  Banks of many detectors: spandsp_b200_mct.h.
   With a tone_callback every change of the detected tone is reported through it, (user_data, tone, level, 0);
   without one the last tone declared is kept for modem_connect_tones_rx_get(). */
modem_connect_tones_rx_state_t *modem_connect_tones_rx_init(modem_connect_tones_rx_state_t *s, int tone_type,
                                                            span_tone_report_func_t tone_callback, void *user_data);
int modem_connect_tones_rx_release(modem_connect_tones_rx_state_t *s);
int modem_connect_tones_rx_free(modem_connect_tones_rx_state_t *s);
int modem_connect_tones_rx(modem_connect_tones_rx_state_t *s, const int16_t amp[], int len);
int modem_connect_tones_rx_fillin(modem_connect_tones_rx_state_t *s, int len);
int modem_connect_tones_rx_get(modem_connect_tones_rx_state_t *s);
const char *modem_connect_tone_to_str(int tone);

/* ---- In-band signalling tone receiver (2280 Hz, 2600 Hz, 2400 + 2600 Hz): src/spandsp/sig_tone.h:56-134,
        src/sig_tone.c:402-734 ---- */
enum
{
    SIG_TONE_2280HZ = 1,
    SIG_TONE_2600HZ,
    SIG_TONE_2400HZ_2600HZ
};                                                                      /* src/spandsp/sig_tone.h:56-64 */

enum
{
    SIG_TONE_1_PRESENT = 0x001,
    SIG_TONE_1_CHANGE = 0x002,
    SIG_TONE_2_PRESENT = 0x004,
    SIG_TONE_2_CHANGE = 0x008,
    SIG_TONE_TX_PASSTHROUGH = 0x010,
    SIG_TONE_RX_PASSTHROUGH = 0x040,
    SIG_TONE_RX_FILTER_TONE = 0x080,
    SIG_TONE_TX_UPDATE_REQUEST = 0x100,
    SIG_TONE_RX_UPDATE_REQUEST = 0x200
};                                                                      /* src/spandsp/sig_tone.h:67-88 */

typedef struct sig_tone_rx_state_s sig_tone_rx_state_t;

/* Synchronous, one receiver per state (a bank of one); amp[] is rewritten in place as in the reference, and
   sig_update(user_data, signalling_state, 0, duration) is called for every change.  Banks: spandsp_b200_sig.h. */
sig_tone_rx_state_t *sig_tone_rx_init(sig_tone_rx_state_t *s, int tone_type, span_tone_report_func_t sig_update, void *user_data);
int sig_tone_rx_release(sig_tone_rx_state_t *s);
int sig_tone_rx_free(sig_tone_rx_state_t *s);
int sig_tone_rx(sig_tone_rx_state_t *s, int16_t amp[], int len);
void sig_tone_rx_set_mode(sig_tone_rx_state_t *s, int mode, int duration);

/* ---- FAX receive front end: the fast modem beside the V.21 receiver until one of them has the signal ---------
   src/fax_modems.c:177-333: fax_modems_v17_v21_rx() / fax_modems_v27ter_v21_rx() / fax_modems_v29_v21_rx() hand every
   block of audio to the fast modem and to the V.21 channel 2 receiver; the first to produce something keeps the line:
   the fast modem when it reports SIG_STATUS_TRAINING_SUCCEEDED (its status handler swaps the rx handler,
   :197-209,244-256,291-303), V.21 when the HDLC layer above it has accepted a frame (rx_frame_received, :219-225,
   :266-272,:313-319).  The HDLC framer itself stays with the caller (it is not part of this library): it tells the
   front end through span_b200_fax_rx_frame_received(). */
typedef struct span_b200_fax_rx_s span_b200_fax_rx_t;

enum
{
    SPAN_B200_FAX_RX_BOTH = 0,      /* the race is on: both receivers get the audio */
    SPAN_B200_FAX_RX_FAST = 1,      /* the fast modem trained: only it runs from here on */
    SPAN_B200_FAX_RX_V21 = 2        /* a V.21 frame was accepted: only the V.21 receiver runs from here on */
};

/* fast_modem: 17, 27 or 29 (V.17, V.27ter, V.29) at bit_rate; fast_put_bit receives the fast modem's bits and status
   reports exactly as fax_modems.c passes them on (the status also when the status handler consumed it, :304), v21_put_bit
   those of the V.21 receiver (fsk_rx_init(&preset_fsk_specs[FSK_V21CH2], FSK_FRAME_MODE_SYNC, ...) with the -39.09 dBm0
   cutoff of fax_modems_start_slow_modem(), :341-343).  NULL on a bad modem / rate or without a device. */
span_b200_fax_rx_t *span_b200_fax_rx_init(int fast_modem, int bit_rate, int short_train,
                                          span_put_bit_func_t fast_put_bit, void *fast_user_data,
                                          span_put_bit_func_t v21_put_bit, void *v21_user_data);
int span_b200_fax_rx_free(span_b200_fax_rx_t *s);
/* fax_modems_vXX_v21_rx(): returns 0 */
int span_b200_fax_rx(span_b200_fax_rx_t *s, const int16_t amp[], int len);
/* fax_modems_vXX_v21_rx_fillin() */
int span_b200_fax_rx_fillin(span_b200_fax_rx_t *s, int len);
/* The caller's HDLC layer accepted a frame from the V.21 bit stream (s->rx_frame_received = true in the reference) */
void span_b200_fax_rx_frame_received(span_b200_fax_rx_t *s);
/* SPAN_B200_FAX_RX_* */
int span_b200_fax_rx_current(const span_b200_fax_rx_t *s);

#if defined(__cplusplus)
}
#endif

#endif
