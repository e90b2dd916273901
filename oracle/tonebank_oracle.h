/*
 * tonebank_oracle.h - declarations for the plain-C CPU restatement (tonebank_oracle.c).
 * TEST INFRASTRUCTURE ONLY.
 */
#if !defined(_TONEBANK_ORACLE_H_)
#define _TONEBANK_ORACLE_H_

#include <stdint.h>

#define TBO_ST_MAX_TONES        32
#define TBO_ST_MAX_STEPS        32

typedef struct
{
    int f1;
    int f2;
    int recognition_duration;
    int min_duration;
    int max_duration;
} tbo_st_segment_t;

/* The reference grows these with realloc (super_tone_rx.c:111-115,127-131,149-152); fixed
   capacity is enough for a checker. */
typedef struct
{
    int used_frequencies;
    int monitored_frequencies;
    int pitches[64][2];
    int tones;
    tbo_st_segment_t tone_list[TBO_ST_MAX_TONES][TBO_ST_MAX_STEPS];
    int tone_segs[TBO_ST_MAX_TONES];
    float fac[64];
} tbo_st_descriptor_t;

float tbo_goertzel_fac(float freq);
int tbo_dtmf_decide(const float row_energy[4], const float col_energy[4], float energy,
                    float threshold, float normal_twist, float reverse_twist);
int tbo_mf_decide(const float energy[6], float threshold, float twist, float relative_peak);
void tbo_st_decide(const float res[], int bins, float energy, int *pk1, int *pk2);
void tbo_st_descriptor_init(tbo_st_descriptor_t *d);
int tbo_st_add_tone(tbo_st_descriptor_t *d);
int tbo_st_add_element(tbo_st_descriptor_t *d, int tone, int f1, int f2, int min, int max);
int tbo_bank_blocks(const float *fac, int bins, int block, const int16_t *amp, int n, float *out, float *energy);

#endif
