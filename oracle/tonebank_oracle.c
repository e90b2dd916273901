/*
 * tonebank_oracle.c - plain-C CPU restatement of the reference's tone-detect hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under spandsp_b200/ may link, load or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement event-for-event against
 * the reference's own sources compiled as oracle/_ref/libspandsp_ref_strict.so (where
 * /root/reference is available) and against the golden fixtures in tests/golden/ that were
 * generated from that library (tests/golden/make_golden.py).
 *
 * The restatement is organised the way the GPU engine is (bank -> block decision -> per-channel
 * sequencer) rather than the way the reference is written, but every arithmetic expression keeps
 * the reference's operand order under strict IEEE-754 binary32 evaluation (no FMA contraction,
 * no re-association): compile with -fno-fast-math -ffp-contract=off.
 *
 * Reference locations (all under /root/reference/src):
 *   Goertzel update   tone_detect.h:172-192, tone_detect.c:123-156
 *   Goertzel result   tone_detect.c:160-205
 *   coefficient       tone_detect.c:60-68
 *   DTMF              dtmf.c:71,104-123 (constants), 132-361 (dtmf_rx), 363-379 (fillin), 421-445 (parms)
 *   Bell MF           bell_r2_mf.c:204,236-262 (constants), 507-673
 *   R2 MF             bell_r2_mf.c:206,240-276 (constants), 750-880
 *   Super-tone        super_tone_rx.c:75-77 (constants), 81-161 (descriptor), 164-228 (cadence),
 *                     289-451 (chunk), 454-490 (rx)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <math.h>
#include <time.h>
#include <pthread.h>

#include "ref_harness.h"
#include "tonebank_oracle.h"

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------ */
/* Goertzel bank                                                                         */

/* tone_detect.c:60-68: the argument of cosf is formed in double (M_PI is a double constant)
   and then narrowed; the product with 2.0f is a float multiply. */
EXPORT float tbo_goertzel_fac(float freq)
{
    return 2.0f*cosf((float) (2.0f*M_PI*(freq/8000.0f)));
}

typedef struct
{
    float v2;
    float v3;
    float fac;
} bin_t;

/* tone_detect.h:184-190:  v1 = v2; v2 = v3; v3 = fac*v2 - v1 + amp  (evaluated left to right) */
static inline void bin_step(bin_t *b, float x)
{
    float v1;

    v1 = b->v2;
    b->v2 = b->v3;
    b->v3 = b->fac*b->v2 - v1 + x;
}

/* tone_detect.c:174-203: push one zero sample, then 2*(v3*v3 + v2*v2 - v2*v3*fac), reset. */
static inline float bin_finish(bin_t *b)
{
    float v1;

    v1 = b->v2;
    b->v2 = b->v3;
    b->v3 = b->fac*b->v2 - v1;
    v1 = b->v3*b->v3 + b->v2*b->v2 - b->v2*b->v3*b->fac;
    v1 *= 2.0f;
    b->v2 = 0.0f;
    b->v3 = 0.0f;
    return v1;
}

/* ------------------------------------------------------------------------------------ */
/* event recording (same record layout as the reference harness)                          */

typedef struct
{
    ref_event_t *ev;
    int64_t cap;
    int64_t n;
    int32_t chunk;
} recorder_t;

static void rec_push(recorder_t *r, int kind, int a, int b, int c)
{
    if (r->n < r->cap)
    {
        r->ev[r->n].chunk = r->chunk;
        r->ev[r->n].kind = kind;
        r->ev[r->n].a = a;
        r->ev[r->n].b = b;
        r->ev[r->n].c = c;
    }
    r->n++;
}

/* A digit string buffer with the reference's overflow rule (dtmf.c:322-338, bell_r2_mf.c:640-655):
   at most 128 buffered digits; with a callback installed every digit is delivered at once. */
typedef struct
{
    char digits[129];
    int current;
    int lost;
} digit_buf_t;

static void digit_buf_put(digit_buf_t *d, int hit, int has_cb, recorder_t *rec)
{
    if (d->current < 128)
    {
        d->digits[d->current++] = (char) hit;
        d->digits[d->current] = '\0';
        if (has_cb)
        {
            int i;
            for (i = 0;  i < d->current;  i++)
                rec_push(rec, REF_EV_DIGIT, (unsigned char) d->digits[i], d->current, i);
            d->current = 0;
        }
    }
    else
    {
        d->lost++;
    }
}

/* ------------------------------------------------------------------------------------ */
/* DTMF                                                                                  */

static const float dtmf_row_hz[4] = {697.0f, 770.0f, 852.0f, 941.0f};      /* dtmf.c:114-117 */
static const float dtmf_col_hz[4] = {1209.0f, 1336.0f, 1477.0f, 1633.0f};  /* dtmf.c:118-121 */
static const char dtmf_positions[] = "123A" "456B" "789C" "*0#D";          /* dtmf.c:123 */

#define DTMF_BLOCK              102             /* dtmf.c:71 */
#define DTMF_THRESHOLD          171029200.0f    /* dtmf.c:104 */
#define DTMF_NORMAL_TWIST       6.309f          /* dtmf.c:105 */
#define DTMF_REVERSE_TWIST      2.512f          /* dtmf.c:106 */
#define DTMF_RELATIVE_PEAK      6.309f          /* dtmf.c:107-108 */
#define DTMF_TO_TOTAL_ENERGY    83.868f         /* dtmf.c:109 */
#define DTMF_POWER_OFFSET       107.255f        /* dtmf.c:110 */

typedef struct
{
    bin_t row[4];
    bin_t col[4];
    float energy;
    float z350[2];
    float z440[2];
    int filter_dialtone;
    float normal_twist;
    float reverse_twist;
    float threshold;
    int current_sample;
    int duration;
    uint8_t last_hit;
    uint8_t in_digit;
    digit_buf_t buf;
} dtmf_chan_t;

static void dtmf_chan_init(dtmf_chan_t *s)
{
    int i;

    memset(s, 0, sizeof(*s));
    for (i = 0;  i < 4;  i++)
    {
        s->row[i].fac = tbo_goertzel_fac(dtmf_row_hz[i]);
        s->col[i].fac = tbo_goertzel_fac(dtmf_col_hz[i]);
    }
    s->normal_twist = DTMF_NORMAL_TWIST;
    s->reverse_twist = DTMF_REVERSE_TWIST;
    s->threshold = DTMF_THRESHOLD;
}

/* dtmf.c:421-445; goertzel_threshold_dbm0 is tone_detect.h:66; db_to_power_ratio telephony.h:141 */
static void dtmf_chan_parms(dtmf_chan_t *s, int filter_dialtone, float twist, float reverse_twist, float threshold)
{
    if (filter_dialtone >= 0)
    {
        s->z350[0] = s->z350[1] = s->z440[0] = s->z440[1] = 0.0f;
        s->filter_dialtone = filter_dialtone;
    }
    if (twist >= 0.0f)
        s->normal_twist = powf(10.0f, twist/10.0f);
    if (reverse_twist >= 0.0f)
        s->reverse_twist = powf(10.0f, reverse_twist/10.0f);
    if (threshold > -99.0f)
        s->threshold = (float) ((DTMF_BLOCK*DTMF_BLOCK*32768.0f*32768.0f/2.0f)*powf(10.0f, (threshold - 3.14f)/10.0f));
}

/* dtmf.c:211-258: block decision from the eight bin energies and the block's sample energy. */
EXPORT int tbo_dtmf_decide(const float row_energy[4], const float col_energy[4], float energy,
                           float threshold, float normal_twist, float reverse_twist)
{
    int best_row;
    int best_col;
    int i;

    best_row = 0;
    best_col = 0;
    for (i = 1;  i < 4;  i++)
    {
        if (row_energy[i] > row_energy[best_row])
            best_row = i;
        if (col_energy[i] > col_energy[best_col])
            best_col = i;
    }
    if (row_energy[best_row] < threshold  ||  col_energy[best_col] < threshold)
        return 0;
    if (!(col_energy[best_col] < row_energy[best_row]*reverse_twist
          &&
          col_energy[best_col]*normal_twist > row_energy[best_row]))
        return 0;
    for (i = 0;  i < 4;  i++)
    {
        if ((i != best_col  &&  col_energy[i]*DTMF_RELATIVE_PEAK > col_energy[best_col])
            ||
            (i != best_row  &&  row_energy[i]*DTMF_RELATIVE_PEAK > row_energy[best_row]))
            return 0;
    }
    if (!((row_energy[best_row] + col_energy[best_col]) > DTMF_TO_TOTAL_ENERGY*energy))
        return 0;
    return dtmf_positions[(best_row << 2) + best_col];
}

/* dtmf.c:304-347: two-block debounce.  Returns nothing; reports through rec. */
static void dtmf_sequence(dtmf_chan_t *s, int hit, int mode, recorder_t *rec)
{
    if (hit != s->in_digit  &&  s->last_hit != s->in_digit)
    {
        hit = (hit  &&  hit == s->last_hit)  ?  hit  :  0;
        if (mode == REF_MODE_REALTIME)
        {
            if (s->in_digit  ||  hit)
            {
                /* dtmf.c:314: lfastrintf is C truncation on x86-64 (fast_convert.h:194-197) */
                int level = (s->in_digit  &&  !hit)  ?  -99  :  (int) (long int) (10.0f*log10f(s->energy) - DTMF_POWER_OFFSET);
                rec_push(rec, REF_EV_TONE, hit, level, s->duration);
                s->duration = 0;
            }
        }
        else if (hit)
        {
            digit_buf_put(&s->buf, hit, mode == REF_MODE_DIGITS_CB, rec);
        }
        s->in_digit = (uint8_t) hit;
    }
    s->last_hit = (uint8_t) hit;
}

static void dtmf_chan_rx(dtmf_chan_t *s, const int16_t amp[], int samples, int mode, recorder_t *rec)
{
    int sample;
    int limit;
    int j;
    int i;
    float xamp;
    float famp;
    float v1;
    float row_energy[4];
    float col_energy[4];

    for (sample = 0;  sample < samples;  sample = limit)
    {
        /* dtmf.c:154-161 */
        if ((samples - sample) >= (DTMF_BLOCK - s->current_sample))
            limit = sample + (DTMF_BLOCK - s->current_sample);
        else
            limit = samples;
        for (j = sample;  j < limit;  j++)
        {
            xamp = amp[j];
            if (s->filter_dialtone)
            {
                /* dtmf.c:167-183 */
                famp = xamp;
                v1 = 0.98356f*famp + 1.8954426f*s->z350[0] - 0.9691396f*s->z350[1];
                famp = v1 - 1.9251480f*s->z350[0] + s->z350[1];
                s->z350[1] = s->z350[0];
                s->z350[0] = v1;
                v1 = 0.98456f*famp + 1.8529543f*s->z440[0] - 0.9691396f*s->z440[1];
                famp = v1 - 1.8819938f*s->z440[0] + s->z440[1];
                s->z440[1] = s->z440[0];
                s->z440[0] = v1;
                xamp = famp;
            }
            s->energy += xamp*xamp;                 /* dtmf.c:189 */
            for (i = 0;  i < 4;  i++)
            {
                bin_step(&s->row[i], xamp);         /* dtmf.c:191-198 */
                bin_step(&s->col[i], xamp);
            }
        }
        if (s->duration < INT_MAX - (limit - sample))   /* dtmf.c:201-203 */
            s->duration += (limit - sample);
        s->current_sample += (limit - sample);
        if (s->current_sample < DTMF_BLOCK)
            continue;
        for (i = 0;  i < 4;  i++)
        {
            row_energy[i] = bin_finish(&s->row[i]);
            col_energy[i] = bin_finish(&s->col[i]);
        }
        dtmf_sequence(s,
                      tbo_dtmf_decide(row_energy, col_energy, s->energy, s->threshold, s->normal_twist, s->reverse_twist),
                      mode, rec);
        s->energy = 0.0f;
        s->current_sample = 0;
    }
    /* dtmf.c:352-358 never fires in our three modes: with a callback the buffer is always empty here. */
}

/* dtmf.c:363-379 */
static void dtmf_chan_fillin(dtmf_chan_t *s)
{
    int i;

    for (i = 0;  i < 4;  i++)
    {
        s->row[i].v2 = s->row[i].v3 = 0.0f;
        s->col[i].v2 = s->col[i].v3 = 0.0f;
    }
    s->energy = 0.0f;
    s->current_sample = 0;
}

static void run_dtmf(const ref_params_t *p, const int16_t *amp, int n, recorder_t *rec, ref_final_t *fin)
{
    dtmf_chan_t s;
    int pos;
    int len;

    dtmf_chan_init(&s);
    if (p->dtmf_set_parms)
        dtmf_chan_parms(&s, p->dtmf_filter_dialtone, p->dtmf_twist, p->dtmf_reverse_twist, p->dtmf_threshold);
    rec->chunk = 0;
    for (pos = 0;  pos < n;  pos += len)
    {
        len = (n - pos < p->chunk)  ?  (n - pos)  :  p->chunk;
        if (p->fillin_every > 0  &&  rec->chunk > 0  &&  (rec->chunk % p->fillin_every) == 0)
            dtmf_chan_fillin(&s);
        else
            dtmf_chan_rx(&s, amp + pos, len, p->mode, rec);
        rec->chunk++;
    }
    if (fin)
    {
        /* dtmf.c:382-391 */
        fin->status = (s.in_digit)  ?  s.in_digit  :  ((s.last_hit)  ?  'x'  :  0);
        if (p->mode == REF_MODE_POLL)
        {
            fin->ndigits = (s.buf.current > 255)  ?  255  :  s.buf.current;
            memcpy(fin->digits, s.buf.digits, fin->ndigits);
            fin->digits[fin->ndigits] = '\0';
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Bell MF and R2 MF: six bins, pick the best two                                         */

static const int bell_mf_hz[6] = {700, 900, 1100, 1300, 1500, 1700};        /* bell_r2_mf.c:251-254 */
static const int r2_fwd_hz[6] = {1380, 1500, 1620, 1740, 1860, 1980};       /* bell_r2_mf.c:264-267 */
static const int r2_back_hz[6] = {1140, 1020, 900, 780, 660, 540};          /* bell_r2_mf.c:269-272 */
static const char bell_mf_positions[] = "1247C-358A--69*---0B----#";        /* bell_r2_mf.c:262 */
static const char r2_mf_positions[] = "1247B-358C--69D---0E----F";          /* bell_r2_mf.c:276 */

#define BELL_MF_BLOCK           120             /* bell_r2_mf.c:204 */
#define BELL_MF_THRESHOLD       3343803100.0f   /* bell_r2_mf.c:236 */
#define BELL_MF_TWIST           3.981f
#define BELL_MF_RELATIVE_PEAK   12.589f
#define R2_MF_BLOCK             133             /* bell_r2_mf.c:206 */
#define R2_MF_THRESHOLD         1031766650.0f   /* bell_r2_mf.c:240 */
#define R2_MF_TWIST             5.012f
#define R2_MF_RELATIVE_PEAK     12.589f

/* bell_r2_mf.c:554-628 / 793-863: returns the index into the 25-character position table, or -1. */
EXPORT int tbo_mf_decide(const float energy[6], float threshold, float twist, float relative_peak)
{
    int best;
    int second_best;
    int i;

    if (energy[0] > energy[1])
    {
        best = 0;
        second_best = 1;
    }
    else
    {
        best = 1;
        second_best = 0;
    }
    for (i = 2;  i < 6;  i++)
    {
        if (energy[i] >= energy[best])
        {
            second_best = best;
            best = i;
        }
        else if (energy[i] >= energy[second_best])
        {
            second_best = i;
        }
    }
    if (!(energy[best] >= threshold
          &&  energy[second_best] >= threshold
          &&  energy[best] < energy[second_best]*twist
          &&  energy[best]*twist > energy[second_best]))
        return -1;
    for (i = 0;  i < 6;  i++)
    {
        if (i != best  &&  i != second_best  &&  energy[i]*relative_peak >= energy[second_best])
            return -1;
    }
    if (second_best < best)
    {
        i = best;
        best = second_best;
        second_best = i;
    }
    return best*5 + second_best - 1;
}

typedef struct
{
    bin_t out[6];
    int current_sample;
    uint8_t hits[5];
    digit_buf_t buf;
    int current_digit;
} mf_chan_t;

static void mf_chan_init(mf_chan_t *s, const int hz[6])
{
    int i;

    memset(s, 0, sizeof(*s));
    for (i = 0;  i < 6;  i++)
        s->out[i].fac = tbo_goertzel_fac((float) hz[i]);
}

static void run_bell_mf(const ref_params_t *p, const int16_t *amp, int n, recorder_t *rec, ref_final_t *fin)
{
    mf_chan_t s;
    int pos;
    int len;
    int j;
    int i;
    int k;
    int hit;
    float energy[6];

    mf_chan_init(&s, bell_mf_hz);
    rec->chunk = 0;
    for (pos = 0;  pos < n;  pos += len)
    {
        len = (n - pos < p->chunk)  ?  (n - pos)  :  p->chunk;
        for (j = 0;  j < len;  j++)
        {
            float x = (float) amp[pos + j];
            for (i = 0;  i < 6;  i++)
                bin_step(&s.out[i], x);
            if (++s.current_sample < BELL_MF_BLOCK)
                continue;
            for (i = 0;  i < 6;  i++)
                energy[i] = bin_finish(&s.out[i]);
            k = tbo_mf_decide(energy, BELL_MF_THRESHOLD, BELL_MF_TWIST, BELL_MF_RELATIVE_PEAK);
            hit = (k >= 0)  ?  bell_mf_positions[k]  :  0;
            /* bell_r2_mf.c:629-661: two (KP: four) identical clean blocks preceded by two different ones */
            if (hit
                &&  hit == s.hits[4]  &&  hit == s.hits[3]
                &&  ((hit != '*'  &&  hit != s.hits[2]  &&  hit != s.hits[1])
                     ||
                     (hit == '*'  &&  hit == s.hits[2]  &&  hit != s.hits[1]  &&  hit != s.hits[0])))
            {
                digit_buf_put(&s.buf, hit, p->mode == REF_MODE_DIGITS_CB, rec);
            }
            s.hits[0] = s.hits[1];
            s.hits[1] = s.hits[2];
            s.hits[2] = s.hits[3];
            s.hits[3] = s.hits[4];
            s.hits[4] = (uint8_t) hit;
            s.current_sample = 0;
        }
        rec->chunk++;
    }
    if (fin  &&  p->mode == REF_MODE_POLL)
    {
        fin->ndigits = (s.buf.current > 255)  ?  255  :  s.buf.current;
        memcpy(fin->digits, s.buf.digits, fin->ndigits);
        fin->digits[fin->ndigits] = '\0';
    }
}

static void run_r2_mf(const ref_params_t *p, const int16_t *amp, int n, recorder_t *rec, ref_final_t *fin)
{
    mf_chan_t s;
    int pos;
    int len;
    int j;
    int i;
    int k;
    int hit_digit;
    float energy[6];

    mf_chan_init(&s, (p->r2_fwd)  ?  r2_fwd_hz  :  r2_back_hz);
    rec->chunk = 0;
    for (pos = 0;  pos < n;  pos += len)
    {
        len = (n - pos < p->chunk)  ?  (n - pos)  :  p->chunk;
        for (j = 0;  j < len;  j++)
        {
            float x = (float) amp[pos + j];
            for (i = 0;  i < 6;  i++)
                bin_step(&s.out[i], x);
            if (++s.current_sample < R2_MF_BLOCK)
                continue;
            for (i = 0;  i < 6;  i++)
                energy[i] = bin_finish(&s.out[i]);
            k = tbo_mf_decide(energy, R2_MF_THRESHOLD, R2_MF_TWIST, R2_MF_RELATIVE_PEAK);
            hit_digit = (k >= 0)  ?  r2_mf_positions[k]  :  0;
            if (s.current_digit != hit_digit)       /* bell_r2_mf.c:869-875 */
                rec_push(rec, REF_EV_TONE, hit_digit, (hit_digit)  ?  -10  :  -99, 0);
            s.current_digit = hit_digit;
            s.current_sample = 0;
        }
        rec->chunk++;
    }
    if (fin)
        fin->status = s.current_digit;
}

/* ------------------------------------------------------------------------------------ */
/* Supervisory tones                                                                     */

#define ST_BLOCK                128             /* private/super_tone_rx.h:29 */
#define ST_THRESHOLD            2104205.6f      /* super_tone_rx.c:75 */
#define ST_TWIST                3.981f          /* super_tone_rx.c:76 */
#define ST_TO_TOTAL_ENERGY      1.995f          /* super_tone_rx.c:77 */
#define ST_MAX_BINS             64

EXPORT void tbo_st_descriptor_init(tbo_st_descriptor_t *d)
{
    memset(d, 0, sizeof(*d));
}

/* super_tone_rx.c:81-122.  Frequencies within +-10 Hz of a monitored one share its bin, which is
   retuned to the mean of the two. */
static int st_add_freq(tbo_st_descriptor_t *d, int freq)
{
    int i;

    if (freq == 0)
        return -1;
    for (i = 0;  i < d->used_frequencies;  i++)
    {
        if (d->pitches[i][0] == freq)
            return d->pitches[i][1];
    }
    for (i = 0;  i < d->used_frequencies;  i++)
    {
        if ((d->pitches[i][0] - 10) <= freq  &&  freq <= (d->pitches[i][0] + 10))
        {
            d->pitches[d->used_frequencies][0] = freq;
            d->pitches[d->used_frequencies][1] = i;
            d->fac[d->pitches[i][1]] = tbo_goertzel_fac((float) (freq + d->pitches[i][0])/2);
            d->used_frequencies++;
            return d->pitches[i][1];
        }
    }
    d->pitches[i][0] = freq;
    d->pitches[i][1] = d->monitored_frequencies;
    d->fac[d->monitored_frequencies++] = tbo_goertzel_fac((float) freq);
    d->used_frequencies++;
    return d->pitches[i][1];
}

EXPORT int tbo_st_add_tone(tbo_st_descriptor_t *d)
{
    d->tone_segs[d->tones] = 0;
    return d->tones++;
}

/* super_tone_rx.c:140-161 */
EXPORT int tbo_st_add_element(tbo_st_descriptor_t *d, int tone, int f1, int f2, int min, int max)
{
    int step;
    tbo_st_segment_t *seg;

    step = d->tone_segs[tone];
    seg = &d->tone_list[tone][step];
    seg->f1 = st_add_freq(d, f1);
    seg->f2 = st_add_freq(d, f2);
    seg->min_duration = min*8;
    seg->max_duration = (max == 0)  ?  0x7FFFFFFF  :  max*8;
    d->tone_segs[tone]++;
    return step;
}

/* super_tone_rx.c:164-228 */
static int st_test_cadence(const tbo_st_segment_t *pattern, int steps, const tbo_st_segment_t *test, int rotation)
{
    int i;
    int j;

    if (rotation >= 0)
    {
        j = 0;
        if (steps < 0)
        {
            steps = -steps;
            j = (rotation + steps - 2)%steps;
            if (pattern[j].f1 != test[8].f1  ||  pattern[j].f2 != test[8].f2)
                return 0;
            if (pattern[j].min_duration > test[8].min_duration*ST_BLOCK
                ||  pattern[j].max_duration < test[8].min_duration*ST_BLOCK)
                return 0;
        }
        if (steps)
            j = (rotation + steps - 1)%steps;
        if (pattern[j].f1 != test[9].f1  ||  pattern[j].f2 != test[9].f2)
            return 0;
        if (pattern[j].max_duration < test[9].min_duration*ST_BLOCK)
            return 0;
    }
    else
    {
        for (i = 0;  i < steps;  i++)
        {
            j = i + 10 - steps;
            if (pattern[i].f1 != test[j].f1  ||  pattern[i].f2 != test[j].f2)
                return 0;
            if (pattern[i].min_duration > test[j].min_duration*ST_BLOCK
                ||  pattern[i].max_duration < test[j].min_duration*ST_BLOCK)
                return 0;
        }
    }
    return 1;
}

/* super_tone_rx.c:300-365: which (k1,k2) pair of bins, if any, carries this block. */
EXPORT void tbo_st_decide(const float res[], int bins, float energy, int *pk1, int *pk2)
{
    int k1;
    int k2;
    int j;

    if (res[0] > res[1])
    {
        k1 = 0;
        k2 = 1;
    }
    else
    {
        k1 = 1;
        k2 = 0;
    }
    for (j = 2;  j < bins;  j++)
    {
        if (res[j] >= res[k1])
        {
            k2 = k1;
            k1 = j;
        }
        else if (res[j] >= res[k2])
        {
            k2 = j;
        }
    }
    if ((res[k1] + res[k2]) < ST_TO_TOTAL_ENERGY*energy)
    {
        k1 = -1;
        k2 = -1;
    }
    else if (res[k1] > ST_TWIST*res[k2])
    {
        k2 = -1;
    }
    else if (k2 < k1)
    {
        j = k1;
        k1 = k2;
        k2 = j;
    }
    *pk1 = k1;
    *pk2 = k2;
}

typedef struct
{
    const tbo_st_descriptor_t *desc;
    bin_t state[ST_MAX_BINS];
    int current_sample;     /* the reference keeps one per bin; they always move together */
    float energy;
    int detected_tone;
    int rotation;
    tbo_st_segment_t segments[11];
} st_chan_t;

/* super_tone_rx.c:366-448: the cadence sequencer fed with one block decision. */
static void st_sequence(st_chan_t *s, int k1, int k2, int mode, recorder_t *rec)
{
    const tbo_st_descriptor_t *d = s->desc;
    int j;

    if (k1 != s->segments[10].f1  ||  k2 != s->segments[10].f2)
    {
        s->segments[10].f1 = k1;
        s->segments[10].f2 = k2;
        s->segments[9].min_duration++;
    }
    else if (k1 != s->segments[9].f1  ||  k2 != s->segments[9].f2)
    {
        if (s->detected_tone >= 0)
        {
            if (!st_test_cadence(d->tone_list[s->detected_tone], -d->tone_segs[s->detected_tone], s->segments, s->rotation++))
            {
                s->detected_tone = -1;
                rec_push(rec, REF_EV_TONE, -1, -10, 0);
            }
        }
        if (mode == REF_MODE_SEGMENTS)
            rec_push(rec, REF_EV_SEGMENT, s->segments[9].f1, s->segments[9].f2, s->segments[9].min_duration*ST_BLOCK/8);
        memmove(&s->segments[0], &s->segments[1], 9*sizeof(s->segments[0]));
        s->segments[9].f1 = k1;
        s->segments[9].f2 = k2;
        s->segments[9].min_duration = 1;
    }
    else
    {
        if (s->detected_tone >= 0)
        {
            if (!st_test_cadence(d->tone_list[s->detected_tone], d->tone_segs[s->detected_tone], s->segments, s->rotation))
            {
                s->detected_tone = -1;
                rec_push(rec, REF_EV_TONE, -1, -10, 0);
            }
        }
        s->segments[9].min_duration++;
    }
    if (s->detected_tone < 0)
    {
        for (j = 0;  j < d->tones;  j++)
        {
            if (st_test_cadence(d->tone_list[j], d->tone_segs[j], s->segments, -1))
            {
                s->detected_tone = j;
                s->rotation = 0;
                rec_push(rec, REF_EV_TONE, j, -10, 0);
                break;
            }
        }
    }
}

static void st_block_end(st_chan_t *s, int mode, recorder_t *rec)
{
    const tbo_st_descriptor_t *d = s->desc;
    float res[ST_MAX_BINS];
    int k1;
    int k2;
    int i;
    int bins = d->monitored_frequencies;

    if (s->energy < ST_THRESHOLD)
    {
        k1 = k2 = -1;
        for (i = 0;  i < bins;  i++)
            s->state[i].v2 = s->state[i].v3 = 0.0f;
        s->current_sample = 0;
    }
    else if (bins < 2)
    {
        /* super_tone_rx.c:312-316: the bins are neither read nor reset.  The caller's next
           goertzel_update consumes 0 samples and the chunk logic runs again with energy 0,
           which takes the branch above. */
        k1 = k2 = 0;
    }
    else
    {
        for (i = 0;  i < bins;  i++)
            res[i] = bin_finish(&s->state[i]);
        s->current_sample = 0;
        tbo_st_decide(res, bins, s->energy, &k1, &k2);
    }
    st_sequence(s, k1, k2, mode, rec);
    s->energy = 0.0f;
}

static void fill_descriptor(tbo_st_descriptor_t *d, const ref_params_t *p)
{
    int t;
    int e;
    int k;
    int tone;

    tbo_st_descriptor_init(d);
    k = 0;
    for (t = 0;  t < p->st_tones;  t++)
    {
        tone = tbo_st_add_tone(d);
        for (e = 0;  e < p->st_tone_segs[t];  e++, k++)
            tbo_st_add_element(d, tone, p->st_elements[4*k + 0], p->st_elements[4*k + 1],
                               p->st_elements[4*k + 2], p->st_elements[4*k + 3]);
    }
}

EXPORT int ref_super_tone_bins(const ref_params_t *p, float *fac, int max)
{
    tbo_st_descriptor_t *d;
    int i;
    int n;

    d = (tbo_st_descriptor_t *) malloc(sizeof(*d));
    fill_descriptor(d, p);
    n = d->monitored_frequencies;
    for (i = 0;  i < n  &&  i < max;  i++)
        fac[i] = d->fac[i];
    free(d);
    return n;
}

static void run_super_tone(const ref_params_t *p, const int16_t *amp, int n, recorder_t *rec, ref_final_t *fin)
{
    tbo_st_descriptor_t *d;
    st_chan_t *s;
    int pos;
    int len;
    int i;
    int j;
    int bins;

    d = (tbo_st_descriptor_t *) malloc(sizeof(*d));
    s = (st_chan_t *) calloc(1, sizeof(*s));
    fill_descriptor(d, p);
    bins = d->monitored_frequencies;
    s->desc = d;
    for (i = 0;  i < 11;  i++)
    {
        s->segments[i].f1 = -1;
        s->segments[i].f2 = -1;
        s->segments[i].min_duration = 0;
    }
    s->detected_tone = -1;
    for (i = 0;  i < bins;  i++)
        s->state[i].fac = d->fac[i];
    rec->chunk = 0;
    for (pos = 0;  pos < n;  pos += len)
    {
        len = (n - pos < p->chunk)  ?  (n - pos)  :  p->chunk;
        j = 0;
        while (j < len)
        {
            /* super_tone_rx.c:466-486.  With zero monitored bins the reference's loop never
               advances (x stays 0); descriptors always have at least one bin in practice. */
            int x = len - j;
            if (x > ST_BLOCK - s->current_sample)
                x = ST_BLOCK - s->current_sample;
            for (i = 0;  i < bins;  i++)
            {
                int k;
                for (k = 0;  k < x;  k++)
                    bin_step(&s->state[i], (float) amp[pos + j + k]);
            }
            {
                int k;
                for (k = 0;  k < x;  k++)
                {
                    float xamp = (float) amp[pos + j + k];
                    s->energy += xamp*xamp;
                }
            }
            s->current_sample += x;
            j += x;
            if (s->current_sample >= ST_BLOCK)
                st_block_end(s, p->mode, rec);
        }
        rec->chunk++;
    }
    if (fin)
    {
        fin->status = s->detected_tone;
        fin->ndigits = bins;
    }
    free(s);
    free(d);
}

/* ------------------------------------------------------------------------------------ */
/* multi-channel runner, same signature as the reference harness                          */

static int run_one(const ref_params_t *p, const int16_t *amp, int n, recorder_t *rec, ref_final_t *fin)
{
    if (fin)
        memset(fin, 0, sizeof(*fin));
    switch (p->detector)
    {
    case REF_DET_DTMF:
        run_dtmf(p, amp, n, rec, fin);
        return 0;
    case REF_DET_BELL_MF:
        run_bell_mf(p, amp, n, rec, fin);
        return 0;
    case REF_DET_R2_MF:
        run_r2_mf(p, amp, n, rec, fin);
        return 0;
    case REF_DET_SUPER_TONE:
        run_super_tone(p, amp, n, rec, fin);
        return 0;
    }
    return -1;
}

typedef struct
{
    const ref_params_t *p;
    const int16_t *amp;
    int64_t stride;
    int c0;
    int c1;
    int n;
    ref_event_t *ev;
    int64_t ev_cap;
    int64_t *ev_count;
    ref_final_t *fin;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *) arg;
    recorder_t rec;
    int c;

    for (c = j->c0;  c < j->c1;  c++)
    {
        rec.ev = (j->ev)  ?  (j->ev + (int64_t) c*j->ev_cap)  :  NULL;
        rec.cap = (j->ev)  ?  j->ev_cap  :  0;
        rec.n = 0;
        rec.chunk = 0;
        run_one(j->p, j->amp + (int64_t) c*j->stride, j->n, &rec, (j->fin)  ?  &j->fin[c]  :  NULL);
        if (j->ev_count)
            j->ev_count[c] = rec.n;
    }
    return NULL;
}

EXPORT int ref_abi_version(void)
{
    return REF_HARNESS_ABI;
}

EXPORT double ref_run(const ref_params_t *p, const int16_t *amp, int64_t stride, int channels, int n,
                      int nthreads, ref_event_t *ev, int64_t ev_cap, int64_t *ev_count, ref_final_t *fin)
{
    pthread_t *th;
    job_t *jobs;
    struct timespec t0;
    struct timespec t1;
    int i;

    if (nthreads < 1)
        nthreads = 1;
    if (nthreads > channels)
        nthreads = channels;
    th = (pthread_t *) malloc(sizeof(pthread_t)*nthreads);
    jobs = (job_t *) malloc(sizeof(job_t)*nthreads);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (i = 0;  i < nthreads;  i++)
    {
        jobs[i].p = p;
        jobs[i].amp = amp;
        jobs[i].stride = stride;
        jobs[i].c0 = (int) ((int64_t) channels*i/nthreads);
        jobs[i].c1 = (int) ((int64_t) channels*(i + 1)/nthreads);
        jobs[i].n = n;
        jobs[i].ev = ev;
        jobs[i].ev_cap = ev_cap;
        jobs[i].ev_count = ev_count;
        jobs[i].fin = fin;
        if (nthreads == 1)
            worker(&jobs[i]);
        else
            pthread_create(&th[i], NULL, worker, &jobs[i]);
    }
    if (nthreads > 1)
    {
        for (i = 0;  i < nthreads;  i++)
            pthread_join(th[i], NULL);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    free(jobs);
    return (double) (t1.tv_sec - t0.tv_sec) + 1.0e-9*(double) (t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------------------ */
/* raw Goertzel                                                                          */

EXPORT float ref_goertzel_fac(float freq, int samples)
{
    (void) samples;
    return tbo_goertzel_fac(freq);
}

EXPORT int ref_goertzel_blocks(float freq, int samples, const int16_t *amp, int n, float *out)
{
    bin_t b;
    int i;
    int cs;
    int nb;

    b.v2 = b.v3 = 0.0f;
    b.fac = tbo_goertzel_fac(freq);
    cs = 0;
    nb = 0;
    for (i = 0;  i < n;  i++)
    {
        bin_step(&b, (float) amp[i]);
        if (++cs >= samples)
        {
            out[nb++] = bin_finish(&b);
            cs = 0;
        }
    }
    return nb;
}

/* Per-block bank energies for a whole row (used by the GPU "energies" parity tests):
   out[b*bins + i] for block b and bin i, plus energy[b] = sum x*x over the block. */
EXPORT int tbo_bank_blocks(const float *fac, int bins, int block, const int16_t *amp, int n, float *out, float *energy)
{
    bin_t b[ST_MAX_BINS];
    float e;
    int i;
    int k;
    int cs;
    int nb;

    for (k = 0;  k < bins;  k++)
    {
        b[k].v2 = b[k].v3 = 0.0f;
        b[k].fac = fac[k];
    }
    e = 0.0f;
    cs = 0;
    nb = 0;
    for (i = 0;  i < n;  i++)
    {
        float x = (float) amp[i];
        e += x*x;
        for (k = 0;  k < bins;  k++)
            bin_step(&b[k], x);
        if (++cs >= block)
        {
            for (k = 0;  k < bins;  k++)
                out[nb*bins + k] = bin_finish(&b[k]);
            if (energy)
                energy[nb] = e;
            e = 0.0f;
            cs = 0;
            nb++;
        }
    }
    return nb;
}
