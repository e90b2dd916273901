/*
 * spandsp_b200_gen.h - C ABI of the signal source banks (bulk interface): the generators the reference's own tests
 * and BASELINE's workloads are built from - dtmf_tx() (src/dtmf.c:521-676) on tone_gen() (src/tone_generate.c:125-230)
 * and awgn() (src/awgn.c:82-196) - run for N channels on the device, so that a multi-channel input is produced where
 * it is consumed instead of on the host and over PCIe.
 *
 * Tones are exact: integer phase accumulators, the reference's 2048-entry sine table, one float multiply per tone,
 * float adds in the reference's order, its truncating conversion to int16 - the samples equal dtmf_tx()'s bit for
 * bit.  The noise source reproduces the reference's generator (three linear congruential sequences, the 97-entry
 * shuffle table, polar method in double precision, rounding and saturation) with IEEE operations in the reference's
 * order; the one library call in it, log(), is the device library's (< 1 ulp, like the host's): a last-bit
 * difference there reaches an output sample only if it straddles a rounding boundary (about 1e-14 per sample).
 */
#if !defined(_SPANDSP_B200_GEN_H_)
#define _SPANDSP_B200_GEN_H_

#include <stdint.h>

#include "spandsp_b200.h"

#if defined(__cplusplus)
extern "C"
{
#endif

typedef struct span_b200_dtmf_tx_bank_s span_b200_dtmf_tx_bank_t;
typedef struct span_b200_awgn_bank_s span_b200_awgn_bank_t;

/* dtmf_tx_init(NULL, NULL, NULL) x channels (src/dtmf.c:637-660): -10 dBm0 per tone, 50 ms on / 55 ms off,
   nothing queued.  There is no "more digits" callback: a transmitter stops when its queue is empty. */
span_b200_dtmf_tx_bank_t *span_b200_dtmf_tx_bank_create(span_b200_ctx_t *ctx, int channels);
void span_b200_dtmf_tx_bank_destroy(span_b200_dtmf_tx_bank_t *bank);
int span_b200_dtmf_tx_bank_channels(const span_b200_dtmf_tx_bank_t *bank);
/* dtmf_tx_init() again for channels [first, first+count) */
int span_b200_dtmf_tx_bank_init(span_b200_dtmf_tx_bank_t *bank, int first, int count);
/* dtmf_tx_set_level() (src/dtmf.c:623-627), dtmf_tx_set_timing() (:630-634; a negative time selects the default) */
int span_b200_dtmf_tx_bank_set_level(span_b200_dtmf_tx_bank_t *bank, int first, int count, int level, int twist);
int span_b200_dtmf_tx_bank_set_timing(span_b200_dtmf_tx_bank_t *bank, int first, int count, int on_time, int off_time);
/* dtmf_tx_put() (src/dtmf.c:600-621): the same digits for every channel of the range (len < 0: strlen), or one
   string per channel (digits + i*stride, lens[i]).  As in the reference a channel takes all of its digits or none
   (the queue holds 128).  Returns 0 if every channel took them, else the largest number of digits that did not
   fit in some channel, or -1 on error. */
int span_b200_dtmf_tx_bank_put(span_b200_dtmf_tx_bank_t *bank, int first, int count, const char *digits, int len);
int span_b200_dtmf_tx_bank_put_each(span_b200_dtmf_tx_bank_t *bank, int first, int count, const char *digits, int64_t stride,
                                    const int32_t *lens);
/* dtmf_tx(s, amp, max_samples) (src/dtmf.c:550-597) for every channel: channel c writes d_amp[c*stride ..).  Like the
   reference a channel stops early when its digits run out and leaves the rest of its row alone - unless zero_fill
   is set, which writes zeros there.  span_b200_dtmf_tx_bank_lens() returns what each dtmf_tx() returned. */
int span_b200_dtmf_tx_bank_tx_device(span_b200_dtmf_tx_bank_t *bank, int16_t *d_amp, int64_t stride, int max_samples, int zero_fill,
                                     void *stream);
int span_b200_dtmf_tx_bank_tx_host(span_b200_dtmf_tx_bank_t *bank, int16_t *h_amp, int64_t stride, int max_samples, int zero_fill);
int span_b200_dtmf_tx_bank_lens(span_b200_dtmf_tx_bank_t *bank, int32_t *lens);
/* The *_device calls are asynchronous on the stream given (NULL: the context's own non-blocking stream, which is
   not ordered with the legacy default stream).  Wait for the bank's last call to finish: */
int span_b200_dtmf_tx_bank_sync(span_b200_dtmf_tx_bank_t *bank);

/* awgn_init_dbm0(NULL, seed, level) x channels (src/awgn.c:152-155); channel i is seeded with seeds[i], or with
   seed0 + i when seeds is NULL. */
span_b200_awgn_bank_t *span_b200_awgn_bank_create(span_b200_ctx_t *ctx, int channels, const int32_t *seeds, int seed0, float level_dbm0);
void span_b200_awgn_bank_destroy(span_b200_awgn_bank_t *bank);
int span_b200_awgn_bank_channels(const span_b200_awgn_bank_t *bank);
/* awgn_init_dbm0() / awgn_init_dbov() again for channels [first, first+count) */
int span_b200_awgn_bank_init_dbm0(span_b200_awgn_bank_t *bank, int first, int count, const int32_t *seeds, int seed0, float level);
int span_b200_awgn_bank_init_dbov(span_b200_awgn_bank_t *bank, int first, int count, const int32_t *seeds, int seed0, float level);
/* samples x awgn() per channel (src/awgn.c:169-196): amp[i] = sat_add16(amp[i], awgn()) - how the reference's tests
   put noise on a line - or amp[i] = awgn(). */
int span_b200_awgn_bank_add_device(span_b200_awgn_bank_t *bank, int16_t *d_amp, int64_t stride, int samples, void *stream);
int span_b200_awgn_bank_fill_device(span_b200_awgn_bank_t *bank, int16_t *d_amp, int64_t stride, int samples, void *stream);
int span_b200_awgn_bank_sync(span_b200_awgn_bank_t *bank);

/* The 2048-entry float sine table of the DDS (src/dds_float.c:51-2101) as computed by this library, for verification */
int span_b200_dds_float_table(float *table);

#if defined(__cplusplus)
}
#endif

#endif
