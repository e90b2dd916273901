/*
 * spandsp_b200.h - C ABI of the B200 tone-bank engine (batch / device-resident interface).
 *
 * This is the "bulk" boundary described in SURVEY.md 8(b): many channels of one detector type
 * are processed by one call, with the int16 samples laid out channel-major
 * ([channel][sample], row stride given in samples).  The spandsp-named per-channel drop-in API
 * (dtmf_rx(), bell_mf_rx(), r2_mf_rx(), super_tone_rx(), goertzel_*()) in spandsp_b200_dropin.h
 * is a thin host layer over these entry points.
 *
 * Plain C types only: pointers, sizes, ints.  No CUDA or torch types appear in any signature; a
 * CUDA stream is passed as an opaque `void *` (a cudaStream_t; NULL = the context's own stream).
 *
 * Each entry point names the reference function whose arithmetic it reproduces
 * (paths relative to the spandsp source tree).
 */
#if !defined(_SPANDSP_B200_H_)
#define _SPANDSP_B200_H_

#include <stdint.h>
#include <stddef.h>

#if defined(__cplusplus)
extern "C"
{
#endif

#define SPAN_B200_ABI_VERSION       1

typedef struct span_b200_ctx_s span_b200_ctx_t;
typedef struct span_b200_bank_s span_b200_bank_t;

/* Detector types of a bank */
enum
{
    SPAN_B200_DET_DTMF = 0,         /* src/dtmf.c:132 dtmf_rx() */
    SPAN_B200_DET_BELL_MF = 1,      /* src/bell_r2_mf.c:507 bell_mf_rx() */
    SPAN_B200_DET_R2_MF = 2,        /* src/bell_r2_mf.c:750 r2_mf_rx() */
    SPAN_B200_DET_SUPER_TONE = 3    /* src/super_tone_rx.c:454 super_tone_rx() */
};

/* Event kinds.  One event corresponds to one callback invocation of the reference. */
enum
{
    /* digits_rx_callback_t material: a = digit character (src/dtmf.c:322-333, bell_r2_mf.c:640-651) */
    SPAN_B200_EV_DIGIT = 1,
    /* span_tone_report_func_t(user, a = code, b = level, c = delay):
       DTMF realtime callback (src/dtmf.c:309-318), R2 MF (bell_r2_mf.c:869-875),
       super-tone tone callback (super_tone_rx.c:391,419,438) */
    SPAN_B200_EV_TONE = 2,
    /* tone_segment_func_t(user, a = f1, b = f2, c = duration ms) (super_tone_rx.c:399-405) */
    SPAN_B200_EV_SEGMENT = 5
};

/* 24-byte event record.  Within one rx call the events of any single channel appear in time
   order; across channels the buffer is ordered by (group of 32 channels, block, channel). */
typedef struct
{
    int32_t channel;
    int32_t block;                  /* index, within the rx call, of the detection block that fired it */
    int32_t kind;
    int32_t a;
    int32_t b;
    int32_t c;
} span_b200_event_t;

/* Flattened supervisory tone descriptor: what super_tone_rx_add_tone()/add_element()
   (src/super_tone_rx.c:125-161) were called with. */
typedef struct
{
    int32_t tones;
    const int32_t *tone_segs;       /* [tones] number of elements of each tone */
    const int32_t *elements;        /* all elements, tone after tone: {f1 Hz, f2 Hz, min ms, max ms} */
} span_b200_super_tone_desc_t;

/* ---- context ---------------------------------------------------------------------------- */

/* Bind to a CUDA device (-1: the calling thread's current device).  Fails (NULL) when there
   is no usable sm_100 device; there is no CPU fallback. */
span_b200_ctx_t *span_b200_ctx_create(int device);
void span_b200_ctx_destroy(span_b200_ctx_t *ctx);
int span_b200_abi_version(void);
/* Last error text of the calling thread ("" if none). */
const char *span_b200_last_error(void);
int span_b200_ctx_device(const span_b200_ctx_t *ctx);
int span_b200_ctx_sm_count(const span_b200_ctx_t *ctx);

/* ---- banks ------------------------------------------------------------------------------ */

/* Every channel starts in the state the reference's *_rx_init() leaves it in
   (src/dtmf.c:454-504, bell_r2_mf.c:693-733,889-933, super_tone_rx.c:507-547). */
span_b200_bank_t *span_b200_dtmf_bank_create(span_b200_ctx_t *ctx, int channels);
span_b200_bank_t *span_b200_bell_mf_bank_create(span_b200_ctx_t *ctx, int channels);
span_b200_bank_t *span_b200_r2_mf_bank_create(span_b200_ctx_t *ctx, int channels, int fwd);
span_b200_bank_t *span_b200_super_tone_bank_create(span_b200_ctx_t *ctx, int channels,
                                                   const span_b200_super_tone_desc_t *desc,
                                                   int want_segments);
void span_b200_bank_destroy(span_b200_bank_t *bank);
int span_b200_bank_channels(const span_b200_bank_t *bank);
int span_b200_bank_detector(const span_b200_bank_t *bank);
int span_b200_bank_block_len(const span_b200_bank_t *bank);
/* Number of Goertzel bins per channel (super-tone: the descriptor's monitored frequencies). */
int span_b200_bank_bins(const span_b200_bank_t *bank);
/* Copy the bank's Goertzel coefficients 2*cos(2*pi*f/8000) (src/tone_detect.c:60-68). */
int span_b200_bank_coefficients(const span_b200_bank_t *bank, float *fac, int max);

/* Re-initialise channels [first, first+count) (the *_rx_init(s, ...) in-place form). */
int span_b200_bank_reset(span_b200_bank_t *bank, int first, int count);

/* dtmf_rx_parms() (src/dtmf.c:421-445) for channels [first, first+count).  Same argument
   meaning: filter_dialtone < 0, twist < 0, reverse_twist < 0, threshold <= -99 leave the
   respective setting unchanged. */
int span_b200_dtmf_bank_parms(span_b200_bank_t *bank, int first, int count,
                              int filter_dialtone, float twist, float reverse_twist, float threshold);
/* dtmf_rx_set_realtime_callback() (src/dtmf.c:411-418): on != 0 selects TONE events with level
   and duration (and zeroes the duration counter), on == 0 selects DIGIT events. */
int span_b200_dtmf_bank_realtime(span_b200_bank_t *bank, int first, int count, int on);
/* dtmf_rx_fillin() (src/dtmf.c:363-379) */
int span_b200_dtmf_bank_fillin(span_b200_bank_t *bank, int first, int count);
/* dtmf_rx_status() (src/dtmf.c:382-391) / r2_mf_rx_get() (bell_r2_mf.c:883) / super-tone detected
   tone, for channels [first, first+count).  Synchronises with the bank's pending work. */
int span_b200_bank_status(span_b200_bank_t *bank, int first, int count, int32_t *status);

/* ---- processing ------------------------------------------------------------------------- */

/* Run `samples` samples of every channel through the detector: channel c reads
   d_amp[c*stride .. c*stride + samples).  d_amp is DEVICE memory.  The work is enqueued on
   `stream` (cudaStream_t, NULL = the context stream) and the call returns without waiting.
   Arbitrary `samples` are accepted; partial detection blocks are carried to the next call exactly
   as the reference carries them in its state structure.  Returns 0, or -1 on error.
   Replaces N calls of dtmf_rx()/bell_mf_rx()/r2_mf_rx()/super_tone_rx(). */
int span_b200_bank_rx_device(span_b200_bank_t *bank, const int16_t *d_amp, int64_t stride,
                             int samples, void *stream);

/* Same, with HOST samples: copies them to the device first (pinned memory makes the copy
   asynchronous; pageable memory works too).  The copy is part of the call. */
int span_b200_bank_rx_host(span_b200_bank_t *bank, const int16_t *h_amp, int64_t stride,
                           int samples, void *stream);

/* Same two calls for 8-bit companded samples (G.711), one byte per sample, channel-major
   [channel][sample] with the row stride in samples (= bytes): the expansion of
   ulaw_to_linear()/alaw_to_linear() (src/spandsp/g711.h:165-172,239-252) is fused into the kernel's load,
   so the detectors see exactly the int16 stream the reference would after expanding, at half the
   bytes per sample.  alaw: 0 = u-law, 1 = A-law. */
int span_b200_bank_rx_device_g711(span_b200_bank_t *bank, const uint8_t *d_data, int64_t stride,
                                  int samples, int alaw, void *stream);
int span_b200_bank_rx_host_g711(span_b200_bank_t *bank, const uint8_t *h_data, int64_t stride,
                                int samples, int alaw, void *stream);

/* Wait for the last rx call of this bank and return how many events it produced
   (-1 on error).  *overflow is set when the event buffer capacity was exceeded (the surplus
   is dropped). */
int64_t span_b200_bank_event_count(span_b200_bank_t *bank, int *overflow);
/* Copy up to max events of the last rx call to host memory.  Returns the number copied. */
int64_t span_b200_bank_events(span_b200_bank_t *bank, span_b200_event_t *out, int64_t max);
/* Device pointer of the (already resolved) event records of the last rx call, for callers that
   forward them device-to-device (e.g. an NCCL gather).  Valid until the next rx call. */
const span_b200_event_t *span_b200_bank_events_device(span_b200_bank_t *bank);
/* Copy up to max event records of the last rx call into caller-provided DEVICE memory (enqueued on
   `stream`, after waiting for the rx call).  Returns the number copied. */
int64_t span_b200_bank_events_to_device(span_b200_bank_t *bank, span_b200_event_t *d_out, int64_t max, void *stream);
/* Override the event buffer capacity (events per rx call).  0 = size for the worst case. */
int span_b200_bank_set_event_capacity(span_b200_bank_t *bank, int64_t events);

/* Per-block diagnostics of the last rx call (tests, tuning): the [block][channel] decision
   codes and, for DTMF, the block energies of blocks whose decision was a hit. */
int span_b200_bank_last_blocks(span_b200_bank_t *bank);
int span_b200_bank_block_codes(span_b200_bank_t *bank, uint16_t *codes, int64_t max);

/* Tuning knobs (benchmarks only).  what: 0 = blocks per time slice (0 = auto),
   1 = staging variant (0 = auto), 2 = force the direct (unstaged) kernel, 3 = packed f32x2 adds,
   4 = record CUDA events around the filter-bank kernel of every rx call (see *_kernel_ms). */
int span_b200_bank_tune(span_b200_bank_t *bank, int what, int value);
/* Name of the kernel path the last rx call took ("staged", "direct"). */
const char *span_b200_bank_last_path(const span_b200_bank_t *bank);
/* With tuning knob 4 on: waits for the bank's stream and returns the summed device time (ms) of
   the filter-bank kernel over the rx calls since the previous query; *launches = how many. */
double span_b200_bank_kernel_ms(span_b200_bank_t *bank, int *launches);
/* Number of kernels the last rx call launched. */
int span_b200_bank_last_launches(const span_b200_bank_t *bank);

/* ---- raw Goertzel banks (goertzel_update()/goertzel_result(), src/tone_detect.c:123-205) --- */

/* channels x bins independent Goertzel filters with arbitrary coefficients and block length;
   every completed block of every channel yields `bins` energies.  d_out receives
   [block][bin][channel] floats (device memory, capacity in floats).  Returns the number of
   complete blocks, or -1. */
int span_b200_goertzel_blocks_device(span_b200_ctx_t *ctx, const float *fac, int bins, int block_len,
                                     const int16_t *d_amp, int64_t stride, int channels, int samples,
                                     float *d_out, int64_t out_capacity, void *stream);

#if defined(__cplusplus)
}
#endif

#endif
