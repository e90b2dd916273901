/*
 * ref_harness.h - shared record layouts for the CPU oracles (oracle/_ref harness and the
 * plain-C restatement in tonebank_oracle.c).  TEST INFRASTRUCTURE ONLY.
 */
#if !defined(_REF_HARNESS_H_)
#define _REF_HARNESS_H_

#include <stdint.h>

#define REF_HARNESS_ABI         4

enum
{
    REF_DET_DTMF = 0,
    REF_DET_BELL_MF = 1,
    REF_DET_R2_MF = 2,
    REF_DET_SUPER_TONE = 3
};

enum
{
    REF_MODE_DIGITS_CB = 0,     /* digits delivered through digits_rx_callback_t */
    REF_MODE_REALTIME = 1,      /* dtmf_rx_set_realtime_callback() style */
    REF_MODE_POLL = 2,          /* no callback; digits fetched with *_rx_get() at the end */
    REF_MODE_SEGMENTS = 3       /* super-tone: tone callback + segment callback */
};

enum
{
    REF_EV_DIGIT = 1,           /* a = digit char, b = len of the callback's string, c = index in it */
    REF_EV_TONE = 2,            /* a = code, b = level, c = delay/duration (span_tone_report_func_t) */
    REF_EV_SEGMENT = 5          /* a = f1, b = f2, c = duration ms (tone_segment_func_t) */
};

typedef struct
{
    int32_t chunk;              /* index of the *_rx() call that fired the callback */
    int32_t kind;
    int32_t a;
    int32_t b;
    int32_t c;
} ref_event_t;

#define REF_MAX_ST_TONES        32
#define REF_MAX_ST_ELEMENTS     128

typedef struct
{
    int32_t detector;
    int32_t mode;
    int32_t chunk;              /* samples per *_rx() call */
    int32_t fillin_every;       /* DTMF: >0 -> every k-th call is dtmf_rx_fillin() instead of dtmf_rx() */
    /* dtmf_rx_parms() */
    int32_t dtmf_set_parms;
    int32_t dtmf_filter_dialtone;
    float dtmf_twist;
    float dtmf_reverse_twist;
    float dtmf_threshold;
    /* r2_mf_rx_init() */
    int32_t r2_fwd;
    /* super-tone descriptor, flattened */
    int32_t st_tones;
    int32_t st_tone_segs[REF_MAX_ST_TONES];
    int32_t st_elements[4*REF_MAX_ST_ELEMENTS];   /* {f1, f2, min_ms, max_ms} */
} ref_params_t;

typedef struct
{
    int32_t status;             /* dtmf_rx_status() / r2_mf_rx_get() / super-tone detected_tone */
    int32_t ndigits;            /* POLL mode: digits drained at the end (super-tone: monitored bins) */
    char digits[256];
} ref_final_t;

/* One equalized symbol as delivered to qam_report_handler_t (src/spandsp/v29rx.h:130) */
typedef struct
{
    float re;
    float im;
    float tre;
    float tim;
    int32_t state;
} ref_v29_sym_t;

#endif
