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
   order; across channels the buffer is ordered by (group of channels, block, channel), the groups being runs of
   32 consecutive channels (super-tone banks: 8). */
typedef struct
{
    int32_t channel;
    int32_t block;                  /* index, within the rx call, of the detection block that fired it */
    int32_t kind;
    int32_t a;
    int32_t b;
    int32_t c;
} span_b200_event_t;

/* Compact 12-byte form of the same record, used where records travel: the multi-GPU gather (NCCL) and the host
   read-back of large banks.  channel is the GLOBAL channel number (the bank's channel_base + the channel inside
   the bank); block_kind = block index in bits 0-13 and the kind in bits 14-15 (1 = DIGIT, 2 = TONE, 3 = SEGMENT);
   a and b are the 24-byte record's a and b as signed bytes (digit / code character, tone id, level in dB, segment
   frequency indices: all within -128..127), c its c.  One rx call may hold at most 16384 blocks in this form. */
typedef struct
{
    uint32_t channel;
    int32_t c;
    uint16_t block_kind;
    int8_t a;
    int8_t b;
} span_b200_wire_event_t;

#define SPAN_B200_WIRE_BLOCK(r)     ((int) ((r).block_kind & 0x3FFF))
#define SPAN_B200_WIRE_KIND(r)      ((((r).block_kind >> 14) == 3)  ?  SPAN_B200_EV_SEGMENT  :  (int) ((r).block_kind >> 14))
#define SPAN_B200_WIRE_MAX_BLOCKS   16384

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
/* NUMA node the context's GPU is attached to (-1: unknown).  span_b200_host_alloc() returns pinned host memory placed
   on that node where the OS allows it - the staging memory for *_rx_host(): with several GPUs in a box the host
   interface runs at PCIe speed only if each GPU's buffers live on its own socket.  Free with span_b200_host_free()
   (or implicitly with the context). */
int span_b200_ctx_numa_node(const span_b200_ctx_t *ctx);
void *span_b200_host_alloc(span_b200_ctx_t *ctx, size_t bytes);
void span_b200_host_free(span_b200_ctx_t *ctx, void *p);

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

/* ---- compact records and the multi-GPU gather -------------------------------------- */

/* Switch a bank to 12-byte wire records (on != 0): from the next rx call on, the emit pass writes
   span_b200_wire_event_t records - into one of two buffers used alternately, so that the records of call k can
   still be read (or travel) while call k + 1 runs - and span_b200_bank_events() / _events_device() /
   _events_to_device() are refused; span_b200_bank_event_count() keeps working.  channel_base is added to every
   channel number (the first global channel of this bank's shard). */
int span_b200_bank_set_wire(span_b200_bank_t *bank, int on, uint32_t channel_base);
/* Copy up to max wire records of the last rx call to host memory.  Returns the number copied. */
int64_t span_b200_bank_events_wire(span_b200_bank_t *bank, span_b200_wire_event_t *out, int64_t max);
/* Host helper: wire records -> 24-byte records (channel numbers are reduced by channel_base). */
void span_b200_wire_expand(const span_b200_wire_event_t *in, span_b200_event_t *out, int64_t n, uint32_t channel_base);

/* One communicator per process / GPU; all ranks of a job create it with the same id.  The id is made by
   span_b200_comm_unique_id() on one rank and handed to the others by whatever out-of-band channel the application
   has (the benchmark broadcasts it with torch.distributed).  It is an NCCL communicator (ncclCommInitRankConfig);
   libnccl.so.2 is looked up at run time, the library does not link against it.  max_ctas > 0 caps the thread
   blocks NCCL may use per operation (the filter kernels are compute-bound: NCCL should stay small). */
typedef struct span_b200_comm_s span_b200_comm_t;
#define SPAN_B200_COMM_ID_BYTES     128
int span_b200_comm_unique_id(unsigned char id[SPAN_B200_COMM_ID_BYTES]);
span_b200_comm_t *span_b200_comm_create(span_b200_ctx_t *ctx, const unsigned char id[SPAN_B200_COMM_ID_BYTES], int nranks, int rank,
                                        int max_ctas);
void span_b200_comm_destroy(span_b200_comm_t *comm);
/* How the records travel to the root once the counts are known (all ranks must use the same):
     SPAN_B200_GATHER_PEER_COPY (default)  the root reads exactly each rank's records out of that rank's buffer with its
         copy engines over NVLink / NVSwitch (the buffers are shared through CUDA IPC handles that ride in the counts
         exchange): no SM of any GPU is used, which matters because the filter kernels are compute-bound; the counts
         and the completion are NCCL collectives;
     SPAN_B200_GATHER_NCCL  exact-count ncclSend (ranks) / ncclRecv (root), one group.
   The environment variable SPANDSP_B200_GATHER=nccl selects the second at creation. */
enum
{
    SPAN_B200_GATHER_NCCL = 0,
    SPAN_B200_GATHER_PEER_COPY = 1
};
int span_b200_comm_set_transport(span_b200_comm_t *comm, int transport);
int span_b200_comm_transport(const span_b200_comm_t *comm);
int span_b200_comm_rank(const span_b200_comm_t *comm);
int span_b200_comm_nranks(const span_b200_comm_t *comm);

/* Attach a bank (in wire mode) to a communicator: its records will be gathered to rank `root`.  On the root the
   emit pass writes the root's own records straight into the gather buffer (no copy). */
int span_b200_bank_attach_comm(span_b200_bank_t *bank, span_b200_comm_t *comm, int root);
/* The gather of the records of the bank's last rx call, in two halves so that it can overlap the next call:
     _begin: (collective, asynchronous) all ranks exchange their record counts (ncclAllGather on the communicator's
             own stream, ordered after the rx call by an event);
     _end:   (collective) waits on the HOST for those counts, then enqueues the transfer of exactly each rank's records
             to the root, which lays them out behind its own in rank order (peer copies or ncclSend / ncclRecv, see
             span_b200_comm_set_transport).  Returns the total number of records (all ranks), or -1; counts (may be
             NULL) receives the nranks per-rank counts.
   Pipelined use (what the benchmark does): rx(k); _end(k-1); _begin(k); ... which lets the records of call k-1
   travel while the kernels of call k run.  Two record buffers are in flight; a third rx call waits (on the device)
   for the transfer that still reads the buffer it is about to overwrite. */
int span_b200_bank_gather_begin(span_b200_bank_t *bank);
int64_t span_b200_bank_gather_end(span_b200_bank_t *bank, int64_t *counts);
/* Root only: wait for the transfer started by the last _gather_end and return its records (device memory, valid until
   the rx call after next) / copy them to the host.  Order: the root's records, then the other ranks' in rank order;
   inside a rank as in the bank's own buffer. */
int64_t span_b200_bank_gathered(span_b200_bank_t *bank, const span_b200_wire_event_t **d_records);
int64_t span_b200_bank_gathered_host(span_b200_bank_t *bank, span_b200_wire_event_t *out, int64_t max);
/* Wait for everything the communicator's stream holds */
int span_b200_comm_sync(span_b200_comm_t *comm);

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

/* The same with the block's total energy beside the bins: d_energy [block][channel] receives the sequential float sum of
   x*x over each block - what the reference's remaining users of these primitives keep for their "fraction of total energy"
   tests (src/ademco_contactid.c:901-931, src/v18.c:1560-1600). */
int span_b200_goertzel_blocks_energy_device(span_b200_ctx_t *ctx, const float *fac, int bins, int block_len,
                                            const int16_t *d_amp, int64_t stride, int channels, int samples,
                                            float *d_out, int64_t out_capacity, float *d_energy, void *stream);
/* Those users' tone sets: frequencies (Hz) and block length.  Returns the number of frequencies. */
enum
{
    SPAN_B200_TONE_SET_ADEMCO_CONTACTID = 0,    /* 1400 / 2300 Hz handshake, 55-sample blocks (src/ademco_contactid.c:446,1179-1180) */
    SPAN_B200_TONE_SET_V18 = 1                  /* the nine V.18 probing tones, 102-sample blocks (src/v18.c:177,200-211) */
};
int span_b200_goertzel_tone_set(int which, float *freqs, int max, int *block_len);
/* make_goertzel_descriptor()'s coefficient 2*cos(2*pi*f/8000) as the library computes it (src/tone_detect.c:60-68) */
float span_b200_goertzel_coefficient(float freq);

/* ---- RFC 4733 telephone-event payloads (the on-the-wire form of the digit reports) ------------ */

/* DTMF digit character -> event code 0-15 ('0'-'9', '*', '#', 'A'-'D'); -1 for anything else */
int span_b200_rfc4733_event_code(int digit);
/* One 4-byte payload: event, E bit, volume (0..63 = -dBm0), duration in timestamp units (samples at 8 kHz), network order */
void span_b200_rfc4733_pack(uint8_t out[4], int event, int end, int volume, int duration);
/* A channel's formatter state: the event in progress (-1: none) and its volume.  Initialise event to -1. */
typedef struct
{
    int32_t event;
    int32_t volume;
} span_b200_rfc4733_state_t;
/* One realtime DTMF report (what dtmf_rx's realtime callback / a SPAN_B200_EV_TONE record carries: code = digit or 0,
   level in dBm0, duration in samples since the previous report) -> the payloads it implies, 4 bytes each in out: the end
   packet (E = 1, duration) of the digit in progress, then the first packet (E = 0, duration 0) of a new digit.
   Returns how many (0, 1 or 2). */
int span_b200_rfc4733_dtmf(span_b200_rfc4733_state_t *st, int code, int level, int duration, uint8_t out[8]);

#if defined(__cplusplus)
}
#endif

#endif
