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

/* ---- cadenced tone generators: tone_gen() (src/tone_generate.c:60-230) ------------------------------------------ */
typedef struct span_b200_tone_gen_bank_s span_b200_tone_gen_bank_t;

/* What tone_gen_descriptor_init() takes (src/spandsp/tone_generate.h): frequencies in Hz (f2 < 0: f1 amplitude
   modulated by |f2|, l2 = depth in percent), levels in dBm0, four durations in ms (on, off, on, off), repeat. */
typedef struct
{
    int32_t f1;
    int32_t l1;
    int32_t f2;
    int32_t l2;
    int32_t d1;
    int32_t d2;
    int32_t d3;
    int32_t d4;
    int32_t repeat;
} span_b200_tone_desc_t;

/* N idle generators (they produce nothing until initialised). */
span_b200_tone_gen_bank_t *span_b200_tone_gen_bank_create(span_b200_ctx_t *ctx, int channels);
void span_b200_tone_gen_bank_destroy(span_b200_tone_gen_bank_t *bank);
int span_b200_tone_gen_bank_channels(const span_b200_tone_gen_bank_t *bank);
/* tone_gen_descriptor_init() + tone_gen_init() for channels [first, first+count): the same descriptor for all of them,
   or descs[i] for channel first + i. */
int span_b200_tone_gen_bank_init(span_b200_tone_gen_bank_t *bank, int first, int count, const span_b200_tone_desc_t *desc);
int span_b200_tone_gen_bank_init_each(span_b200_tone_gen_bank_t *bank, int first, int count, const span_b200_tone_desc_t *descs);
/* tone_gen(s, amp, max_samples) for every channel (src/tone_generate.c:125-230); a generator without repeat stops at
   the end of its cadence and leaves the rest of its row alone unless zero_fill is set. */
int span_b200_tone_gen_bank_tx_device(span_b200_tone_gen_bank_t *bank, int16_t *d_amp, int64_t stride, int max_samples, int zero_fill,
                                      void *stream);
int span_b200_tone_gen_bank_lens(span_b200_tone_gen_bank_t *bank, int32_t *lens);
int span_b200_tone_gen_bank_sync(span_b200_tone_gen_bank_t *bank);

/* ---- V.29 transmitters: v29_tx() (src/v29tx.c:100-470) ------------------------------------------------------------- */
typedef struct span_b200_v29_tx_bank_s span_b200_v29_tx_bank_t;

/* v29_tx_init(NULL, bit_rate, tep, get_bit, ...) x channels (src/v29tx.c:401-429): -14 dBm0, training from the start.
   The data bits a get_bit callback would supply come from a per-channel source on the device: by default a
   maximal-length sequence x^23 + x^18 + 1 seeded with channel + 1 (see *_set_prbs / *_set_bits). */
span_b200_v29_tx_bank_t *span_b200_v29_tx_bank_create(span_b200_ctx_t *ctx, int channels, int bit_rate, int tep);
void span_b200_v29_tx_bank_destroy(span_b200_v29_tx_bank_t *bank);
int span_b200_v29_tx_bank_channels(const span_b200_v29_tx_bank_t *bank);
/* v29_tx_restart() (src/v29tx.c:362-398), v29_tx_power() (:323-338) for channels [first, first+count) */
int span_b200_v29_tx_bank_restart(span_b200_v29_tx_bank_t *bank, int first, int count, int bit_rate, int tep);
int span_b200_v29_tx_bank_power(span_b200_v29_tx_bank_t *bank, int first, int count, float power_dbm0);
/* Data source: the 23-bit sequence, channel first + i seeded with seeds[i] (or seed0 + i when seeds is NULL; 0 -> 1) */
int span_b200_v29_tx_bank_set_prbs(span_b200_v29_tx_bank_t *bank, int first, int count, const uint32_t *seeds, uint32_t seed0);
/* Data source: caller bits, LSB first, channel first + i takes nbits[i] bits from bits + i*stride_bytes (host memory,
   copied).  When they run out the transmitter gets SIG_STATUS_END_OF_DATA (src/v29tx.c:109-119) and shuts down. */
int span_b200_v29_tx_bank_set_bits(span_b200_v29_tx_bank_t *bank, int first, int count, const uint8_t *bits, int64_t stride_bytes,
                                   const int32_t *nbits);
/* v29_tx(s, amp, max_samples) for every channel (src/v29tx.c:226-296); lens = what each call returned */
int span_b200_v29_tx_bank_tx_device(span_b200_v29_tx_bank_t *bank, int16_t *d_amp, int64_t stride, int max_samples, int zero_fill,
                                    void *stream);
int span_b200_v29_tx_bank_lens(span_b200_v29_tx_bank_t *bank, int32_t *lens);
/* Per channel: bit 0 = SIG_STATUS_END_OF_DATA seen, bit 1 = SIG_STATUS_SHUTDOWN_COMPLETE reported */
int span_b200_v29_tx_bank_status(span_b200_v29_tx_bank_t *bank, int32_t *status);
int span_b200_v29_tx_bank_sync(span_b200_v29_tx_bank_t *bank);
/* The transmit pulse shaper [10][9] (src/v29tx_rrc.h, generated by src/make_modem_filter.c) as computed by this library */
int span_b200_v29_tx_tables(float *shaper);

/* The 2048-entry float sine table of the DDS (src/dds_float.c:51-2101) as computed by this library, for verification */
int span_b200_dds_float_table(float *table);

#if defined(__cplusplus)
}
#endif

#endif
