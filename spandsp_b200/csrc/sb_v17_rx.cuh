// sb_v17_rx.cuh - the V.17 receiver on top of the shared core: long and short training state machines,
// 8-candidate soft slicer, 8-state trellis (Viterbi) decoder with 16-step traceback, differential decoder,
// descrambler; also the 4800 bit/s V.32bis mode without trellis that the reference carries.
// Reference: src/v17rx.c:340-1134,1386-1494, src/spandsp/private/v17rx.h.
#pragma once

#include "sb_modem.cuh"

namespace sbm {

#define V17_COEFF_SETS              192
#define V17_TRELLIS_STORAGE_DEPTH   16
#define V17_TRELLIS_LOOKBACK_DEPTH  16

#define V17_TRAINING_SEG_1_LEN          256
#define V17_TRAINING_SEG_2_LEN          2976
#define V17_TRAINING_SHORT_SEG_2_LEN    38
#define V17_TRAINING_SEG_3_LEN          64
#define V17_TRAINING_SEG_4A_LEN         15
#define V17_TRAINING_SEG_4_LEN          48
#define V17_BRIDGE_WORD                 0x8880

// Offsets of the five constellations inside V17Tables::constellation
#define V17_CON_14400   0
#define V17_CON_12000   128
#define V17_CON_9600    192
#define V17_CON_7200    224
#define V17_CON_4800    240
#define V17_CON_POINTS  244

// Small lookup tables, staged in shared memory (lanes index them divergently).
struct V17Tables
{
    float constellation[V17_CON_POINTS][2];     // src/v17_v32bis_tx_constellation_maps.h
};

struct V17Consts : CoreConsts
{
    const V17Tables *tables;            // global memory
    const unsigned char *maps;          // [4][36][36][8]  (src/v17_v32bis_rx_constellation_maps.h)
    const unsigned char *map4800;       // [36][36]
    int phase_p90;                      // DDS_PHASE(90.0f)
    int phase_m90;                      // DDS_PHASE(-90.0f)
    int phase_180;                      // DDS_PHASE(180.0f)
    int phase_a;                        // DDS_PHASE(270.0f + 18.433f)
    int phase_b;                        // DDS_PHASE(180.0f + 18.433f)
    int phase_c;                        // DDS_PHASE(18.433f)
    float eq_delta_fast;                // 0.21f/33
    float eq_delta_slow;                // 0.1f*(0.21f/33)
};

// The V.17 signal space (ITU-T V.17 figures 2 to 5, as tabulated in
// src/v17_v32bis_tx_constellation_maps.h).  Every constellation is invariant under 90 degree rotation and
// the low three index bits walk that symmetry: with A = point 8g and B = point 8g + 1 of group g, and
// rot(x, y) = (y, -x), the group is {A, B, rot B, rot A, -A, -B, -rot B, -rot A}.  Only the A/B pairs are
// tabulated here; tests/test_v17_tables.py checks the expansion against the reference's tables.
static inline void make_v17_tables(V17Tables &t)
{
    static const signed char base_14400[16][4] =
    {
        {-8, -3, 9, 2}, {-8, 1, 9, -2}, {-4, -3, 5, 2}, {-4, 1, 5, -2}, {4, -3, -3, 2}, {4, 1, -3, -2}, {0, -3, 1, 2}, {0, 1, 1, -2},
        {8, -3, -7, 2}, {8, 1, -7, -2}, {-4, -7, 5, 6}, {-4, 5, 5, -6}, {4, -7, -3, 6}, {4, 5, -3, -6}, {0, -7, 1, 6}, {0, 5, 1, -6}
    };
    static const signed char base_12000[8][4] =
    {
        {7, 1, -5, -1}, {3, -3, -1, 3}, {7, -7, -5, 7}, {-1, -7, 3, 7}, {3, 5, -1, -5}, {-1, 1, 3, -1}, {-5, 5, 7, -5}, {-5, -3, 7, 3}
    };
    static const signed char base_9600[4][4] =
    {
        {-8, 2, -6, -4}, {0, 2, -6, 4}, {0, -6, 2, -4}, {8, 2, 2, 4}
    };
    static const signed char base_7200[2][4] =
    {
        {6, -6, -2, 6}, {-2, 2, 6, -2}
    };
    static const signed char con_4800[4][2] =
    {
        {-6, -2}, {-2, 6}, {2, -6}, {6, 2}
    };
    struct
    {
        const signed char (*base)[4];
        int groups;
        int offset;
    } sets[4] =
    {
        {base_14400, 16, V17_CON_14400}, {base_12000, 8, V17_CON_12000}, {base_9600, 4, V17_CON_9600}, {base_7200, 2, V17_CON_7200}
    };
    for (int s = 0;  s < 4;  s++)
    {
        for (int g = 0;  g < sets[s].groups;  g++)
        {
            const int ax = sets[s].base[g][0];
            const int ay = sets[s].base[g][1];
            const int bx = sets[s].base[g][2];
            const int by = sets[s].base[g][3];
            const int pts[8][2] = {{ax, ay}, {bx, by}, {by, -bx}, {ay, -ax}, {-ax, -ay}, {-bx, -by}, {-by, bx}, {-ay, ax}};
            for (int i = 0;  i < 8;  i++)
            {
                t.constellation[sets[s].offset + 8*g + i][0] = (float) pts[i][0];
                t.constellation[sets[s].offset + 8*g + i][1] = (float) pts[i][1];
            }
        }
    }
    for (int i = 0;  i < 4;  i++)
    {
        t.constellation[V17_CON_4800 + i][0] = (float) con_4800[i][0];
        t.constellation[V17_CON_4800 + i][1] = (float) con_4800[i][1];
    }
}

// The soft-decision maps: for every 0.5 x 0.5 cell of [-9, 9) x [-9, 9), evaluated at the cell centre,
// the nearest constellation point in each of the 8 trellis subsets (index mod 8)
// (src/make_v17_v32_constellation_map.c:62-301).  maps: [4][36][36][8]; map4800: [36][36].
// Where two points of a subset are exactly equally near (1184 cells, all far outside the constellation)
// the generator's rule lets the last one win, but the table checked in as
// src/v17_v32bis_rx_constellation_maps.h - the one the receiver is compiled with - holds the first one in
// 396 of them, without a pattern an evaluation rule would reproduce.  v17_tie_first[] records, one bit per
// tie in scan order (map, re cell, im cell, subset), where the first point is held
// (tools/make_v17_tie_bitmap.py derives it; tests/test_v17_tables.py checks the result byte for byte).
static inline void make_v17_maps(const V17Tables &t, std::vector<unsigned char> &maps, std::vector<unsigned char> &map4800)
{
    static const int offset[4] = {V17_CON_14400, V17_CON_12000, V17_CON_9600, V17_CON_7200};
    static const int points[4] = {128, 64, 32, 16};
    static const unsigned int v17_tie_first[37] =
    {
        0x03C0F0F0u, 0x0380E03Cu, 0x00000802u, 0x3F0F1C75u, 0x0300C03Cu, 0xF0F0C03Cu, 0x00A28A28u, 0x00540054u,
        0x004A004Au, 0x28002328u, 0x08EB0023u, 0x0008EB00u, 0xBE0082BEu, 0xF18CB282u, 0xE4F18CB2u, 0x23E4E423u,
        0xD808D8E4u, 0x70D808D8u, 0x0270B202u, 0x4C828CB2u, 0x224C828Cu, 0x2822B328u, 0x4CEB09B3u, 0x854CEB09u,
        0xC5178517u, 0x828BC58Bu, 0x110028A0u, 0x55000000u, 0xF5333333u, 0x28A28A28u, 0x5050B2CAu, 0x48585848u,
        0x22222323u, 0x090B0B09u, 0x05050505u, 0x20828585u, 0x28A08208u
    };
    int tie = 0;
    maps.assign(4*36*36*8, 0);
    map4800.assign(36*36, 0);
    for (int m = 0;  m < 4;  m++)
    {
        for (int ire = 0;  ire <= 35;  ire++)
        {
            const double re = (ire - 18)/2.0 + 0.25;
            for (int iim = 0;  iim <= 35;  iim++)
            {
                const double im = (iim - 18)/2.0 + 0.25;
                for (int i = 0;  i < 8;  i++)
                {
                    int best = 0;
                    int first_best = 0;
                    double best_distance = 1000000.0;
                    for (int l = i;  l < points[m];  l += 8)
                    {
                        const double cr = t.constellation[offset[m] + l][0];
                        const double ci = t.constellation[offset[m] + l][1];
                        const double distance = (re - cr)*(re - cr) + (im - ci)*(im - ci);
                        if (distance < best_distance)
                            first_best = l;
                        if (distance <= best_distance)
                        {
                            best = l;
                            best_distance = distance;
                        }
                    }
                    if (best != first_best)
                    {
                        if (tie < 37*32  &&  ((v17_tie_first[tie >> 5] >> (tie & 31)) & 1u))
                            best = first_best;
                        tie++;
                    }
                    maps[((m*36 + ire)*36 + iim)*8 + i] = (unsigned char) best;
                }
            }
        }
    }
    for (int ire = 0;  ire <= 35;  ire++)
    {
        const double re = (ire - 18)/2.0 + 0.25;
        for (int iim = 0;  iim <= 35;  iim++)
        {
            const double im = (iim - 18)/2.0 + 0.25;
            int best = 0;
            double best_distance = 1000000.0;
            for (int l = 0;  l < 4;  l++)
            {
                const double cr = t.constellation[V17_CON_4800 + l][0];
                const double ci = t.constellation[V17_CON_4800 + l][1];
                const double distance = (re - cr)*(re - cr) + (im - ci)*(im - ci);
                if (distance <= best_distance)
                {
                    best = l;
                    best_distance = distance;
                }
            }
            map4800[ire*36 + iim] = (unsigned char) best;
        }
    }
}

struct RxV17 : RxCore<RxV17, V17_COEFF_SETS>
{
    typedef V17Consts Consts;
    typedef RxCore<RxV17, V17_COEFF_SETS> Core;

    enum
    {
        STAGE_NORMAL = 0, STAGE_SYMBOL_ACQUISITION, STAGE_LOG_PHASE, STAGE_SHORT_WAIT_FOR_CDBA, STAGE_WAIT_FOR_CDBA,
        STAGE_COARSE_TRAIN_ON_CDBA, STAGE_FINE_TRAIN_ON_CDBA, STAGE_SHORT_TRAIN_ON_CDBA_AND_TEST,
        STAGE_TRAIN_ON_CDBA_AND_TEST, STAGE_BRIDGE, STAGE_TCM_WINDUP, STAGE_TEST_ONES, STAGE_PARKED
    };
    // One trellis step = 8 states x {past state (3 bits), constellation point (7 bits)} packed as 8 x 16 bits.
    static const int TRELLIS_WORDS = V17_TRELLIS_STORAGE_DEPTH*4;
    enum
    {
        I_DIFF = I_CORE_COUNT, I_SHORT_TRAIN, I_SPACE_MAP, I_BITS_PER_SYMBOL, I_TRELLIS_PTR, I_CON_OFFSET,
        I_TRELLIS,                                  // 64 words
        I_COUNT = I_TRELLIS + TRELLIS_WORDS
    };
    enum
    {
        F_DISTANCES = F_CORE_COUNT,                 // 8
        F_COUNT = F_DISTANCES + 8
    };
    static const int TABLE_WORDS = ((sizeof(V17Tables) + 15)/16)*4;     // keeps what follows 16-byte aligned
    static const int LANE_WORDS = Core::CORE_LANE_WORDS + TRELLIS_WORDS;

    int diff;
    int short_train;
    int space_map;
    int bits_per_symbol;
    int trellis_ptr;
    int con_offset;                     // which constellation of V17Tables (s->constellation)
    float dist0, dist1, dist2, dist3, dist4, dist5, dist6, dist7;      // s->distances[8]
    int *trellis;                       // [64] words, lane-interleaved
    const V17Tables *t;

    static SB_HD void fill_tables(float *dst, const Consts &k, int lane, int nlanes)
    {
        const unsigned int *src = (const unsigned int *) k.tables;
        unsigned int *d = (unsigned int *) dst;
        for (int i = lane;  i < (int) ((sizeof(*k.tables) + 3)/4);  i += nlanes)
            d[i] = src[i];
    }

    SB_HD void bind(const float *tables, float *lane_block, int lane)
    {
        t = (const V17Tables *) tables;
        bind_core(lane_block, lane);
        trellis = (int *) (lane_block + Core::CORE_LANE_WORDS*32) + lane;
    }

    template <class V> SB_HD void visit(V &v)
    {
        visit_core(v);
        v.i(I_DIFF, diff);
        v.i(I_SHORT_TRAIN, short_train);
        v.i(I_SPACE_MAP, space_map);
        v.i(I_BITS_PER_SYMBOL, bits_per_symbol);
        v.i(I_TRELLIS_PTR, trellis_ptr);
        v.i(I_CON_OFFSET, con_offset);
        for (int w = 0;  w < TRELLIS_WORDS;  w++)
            v.i(I_TRELLIS + w, trellis[w*32]);
        v.f(F_DISTANCES + 0, dist0);
        v.f(F_DISTANCES + 1, dist1);
        v.f(F_DISTANCES + 2, dist2);
        v.f(F_DISTANCES + 3, dist3);
        v.f(F_DISTANCES + 4, dist4);
        v.f(F_DISTANCES + 5, dist5);
        v.f(F_DISTANCES + 6, dist6);
        v.f(F_DISTANCES + 7, dist7);
    }

    SB_HD void equalizer_restore17(const Consts &k)
    {
        equalizer_restore();
        eq_delta = k.eq_delta_slow;
        eq_skip = 0;
    }

    SB_HD void equalizer_reset17(const Consts &k)
    {
        equalizer_reset();
        eq_delta = k.eq_delta_fast;
        eq_skip = 0;
    }

    // v17_rx_restart (src/v17rx.c:1386-1494).  Returns -1 for a bad bit rate.
    SB_HD int restart(const Consts &k, int rate, int short_train_arg)
    {
        switch (rate)
        {
        case 14400:
            con_offset = V17_CON_14400;
            space_map = 0;
            bits_per_symbol = 6;
            break;
        case 12000:
            con_offset = V17_CON_12000;
            space_map = 1;
            bits_per_symbol = 5;
            break;
        case 9600:
            con_offset = V17_CON_9600;
            space_map = 2;
            bits_per_symbol = 4;
            break;
        case 7200:
            con_offset = V17_CON_7200;
            space_map = 3;
            bits_per_symbol = 3;
            break;
        case 4800:
            con_offset = V17_CON_4800;
            space_map = 0;
            bits_per_symbol = 2;
            break;
        default:
            return -1;
        }
        bit_rate = rate;
        rrc_clear();
        training_error = 0.0f;
        diff = 1;
        scramble_reg = 0x2ECDD5;
        training_stage = STAGE_SYMBOL_ACQUISITION;
        training_count = 0;
        signal_present = 0;
        high_sample = 0;
        low_samples = 0;
        drop_pending = 0;
        if (short_train_arg != 2)
            short_train = short_train_arg;
        last_angle0 = last_angle1 = 0;
        for (int i = 0;  i < 16;  i++)
            diff_angles[i*32] = 0;
        dist0 = 0.0f;
        dist1 = dist2 = dist3 = dist4 = dist5 = dist6 = dist7 = fmul(99.0f, 1.0f);
        for (int i = 0;  i < TRELLIS_WORDS;  i++)
            trellis[i*32] = 0;
        trellis_ptr = 14;
        carrier_phase = 0;
        power = 0;                                  // power_meter_init(&s->power, 4)
        if (short_train)
        {
            phase_rate = phase_rate_save;
            equalizer_restore17(k);
            agc_scaling = agc_scaling_save;
            track_i = 0.0f;
            track_p = 40000.0f;
        }
        else
        {
            phase_rate = k.rate_nominal;
            equalizer_reset17(k);
            agc_scaling_save = 0.0f;
            agc_scaling = k.agc_initial;
            track_i = 5000.0f;
            track_p = 40000.0f;
        }
        last_sample = 0;
        godard_init();
        baud_half = 0;
        return 0;
    }

    // v17_rx_init (src/v17rx.c:1496-1530): memset, scrambler tap, signal cutoff, saved carrier, restart
    SB_HD void init(const Consts &k, int rate, int on_pw, int off_pw)
    {
        eq_step = 0;
        eq_put_step = 0;
        eq_skip = 0;
        eq_delta = 0.0f;
        on_power = on_pw;
        off_power = off_pw;
        agc_scaling_save = 0.0f;
        short_train = 0;
        phase_rate_save = k.rate_nominal;
        restart(k, rate, 0);
    }

    SB_HD void restart_after_carrier_down(const Consts &k)
    {
        restart(k, bit_rate, short_train);      // src/v17rx.c:1186
    }

    // src/v17rx.c:340-354 (scrambler_tap = 18 - 1, :1523)
    SB_HD int descramble(int in_bit)
    {
        in_bit &= 1;
        const int out_bit = (in_bit ^ (int) (scramble_reg >> (18 - 1)) ^ (int) (scramble_reg >> (23 - 1))) & 1;
        scramble_reg <<= 1;
        if (training_stage > STAGE_NORMAL  &&  training_stage < STAGE_TCM_WINDUP)
            scramble_reg |= (unsigned int) out_bit;
        else
            scramble_reg |= (unsigned int) in_bit;
        return out_bit;
    }

    // src/v17rx.c:357-376
    SB_HD void put_bit(int bit)
    {
        const int out = descramble(bit);
        if (training_stage == STAGE_NORMAL)
            out_bit(out);
    }

    SB_HD float con_re(int idx) const { return t->constellation[con_offset + idx][0]; }
    SB_HD float con_im(int idx) const { return t->constellation[con_offset + idx][1]; }

    // src/v17rx.c:394-589
    SB_HD int decode_baud(const Consts &k, float zre, float zim)
    {
        int re = f2i(fmul(fadd(zre, 9.0f), 2.0f));
        int im = f2i(fmul(fadd(zim, 9.0f), 2.0f));
        re = (re > 35)  ?  35  :  (re < 0)  ?  0  :  re;
        im = (im > 35)  ?  35  :  (im < 0)  ?  0  :  im;
        if (bits_per_symbol == 2)
        {
            // 4800 bit/s V.32bis mode, without trellis coding.  v32bis_4800_differential_decoder[4][4]
            // (src/v17rx.c:401-407) packed two bits per entry, row = previous state.
            const int constellation_state = ldg(k.map4800 + re*36 + im);
            const unsigned int dd = (2u | (3u << 2) | (0u << 4) | (1u << 6))
                                  | ((0u | (2u << 2) | (1u << 4) | (3u << 6)) << 8)
                                  | ((3u | (1u << 2) | (2u << 4) | (0u << 6)) << 16)
                                  | ((1u | (0u << 2) | (3u << 4) | (2u << 6)) << 24);
            const int raw = (int) ((dd >> (8*diff + 2*constellation_state)) & 3u);
            diff = constellation_state;
            put_bit(raw);
            put_bit(raw >> 1);
            return constellation_state;
        }
        // The 8 candidate points, one per trellis subset, and their squared distances
        const uint2 cand8 = ldg((const uint2 *) (k.maps + (size_t) ((space_map*36 + re)*36 + im)*8));
        int cand[8];
        float distances[8];
        float min = 9999999.0f;
        int min_index = 0;
#pragma unroll
        for (int i = 0;  i < 8;  i++)
        {
            cand[i] = (int) ((((i < 4)  ?  cand8.x  :  cand8.y) >> (8*(i & 3))) & 0xFFu);
            const float dr = fsub(con_re(cand[i]), zre);
            const float di = fsub(con_im(cand[i]), zim);
            distances[i] = fadd(fmul(dr, dr), fmul(di, di));
            if (min > distances[i])
            {
                min = distances[i];
                min_index = i;
            }
        }
        int constellation_state = cand[0];
#pragma unroll
        for (int i = 1;  i < 8;  i++)
            constellation_state = (min_index == i)  ?  cand[i]  :  constellation_state;
        track_carrier(zre, zim, con_re(constellation_state), con_im(constellation_state));

        // Trellis: update the accumulated distance to each of the 8 states (src/v17rx.c:512-541)
        if (++trellis_ptr >= V17_TRELLIS_STORAGE_DEPTH)
            trellis_ptr = 0;
        const float old[8] = {dist0, dist1, dist2, dist3, dist4, dist5, dist6, dist7};
        float nd[8];
        unsigned int packed[8];
        // tcm_paths[8][4], src/v17rx.c:415-425
        const int tcm_paths[8][4] =
        {
            {0, 6, 2, 4}, {6, 0, 4, 2}, {2, 4, 0, 6}, {4, 2, 6, 0}, {1, 3, 7, 5}, {5, 7, 3, 1}, {7, 5, 1, 3}, {3, 1, 5, 7}
        };
#pragma unroll
        for (int i = 0;  i < 8;  i++)
        {
            const int set = i >> 2;
            float m = fadd(distances[tcm_paths[i][0]], old[set]);
            float best_old = old[set];
            float best_d = distances[tcm_paths[i][0]];
            int best_cand = cand[tcm_paths[i][0]];
            int best_k = set;
#pragma unroll
            for (int j = 1;  j < 4;  j++)
            {
                const int kk = (j << 1) + set;
                const float v = fadd(distances[tcm_paths[i][j]], old[kk]);
                if (m > v)
                {
                    m = v;
                    best_old = old[kk];
                    best_d = distances[tcm_paths[i][j]];
                    best_cand = cand[tcm_paths[i][j]];
                    best_k = kk;
                }
            }
            // An elementary IIR filter tracks the distance to date
            nd[i] = fadd(fmul(best_old, 0.9f), fmul(best_d, 0.1f));
            packed[i] = (unsigned int) best_k | ((unsigned int) best_cand << 3);
        }
        dist0 = nd[0];
        dist1 = nd[1];
        dist2 = nd[2];
        dist3 = nd[3];
        dist4 = nd[4];
        dist5 = nd[5];
        dist6 = nd[6];
        dist7 = nd[7];
#pragma unroll
        for (int w = 0;  w < 4;  w++)
            trellis[(trellis_ptr*4 + w)*32] = (int) (packed[2*w] | (packed[2*w + 1] << 16));

        // The state with the minimum distance to date starts the path back (src/v17rx.c:543-556)
        float mn = nd[0];
        int kst = 0;
#pragma unroll
        for (int i = 1;  i < 8;  i++)
        {
            if (mn > nd[i])
            {
                mn = nd[i];
                kst = i;
            }
        }
        // Trace back (src/v17rx.c:557-571)
        int j = trellis_ptr;
#pragma unroll 1
        for (int i = 0;  i < V17_TRELLIS_LOOKBACK_DEPTH - 1;  i++)
        {
            const unsigned int w = (unsigned int) trellis[(j*4 + (kst >> 1))*32];
            kst = (int) ((w >> (16*(kst & 1))) & 7u);
            if (--j < 0)
                j = V17_TRELLIS_STORAGE_DEPTH - 1;
        }
        const unsigned int w = (unsigned int) trellis[(j*4 + (kst >> 1))*32];
        const int nearest = (int) (((w >> (16*(kst & 1))) >> 3) & 0x7Fu) >> 1;

        // Differentially decode: v17_differential_decoder[d][x] = (x - d) & 3 (src/v17rx.c:408-414)
        int raw = (nearest & 0x3C) | (((nearest & 0x03) - diff) & 3);
        diff = nearest & 0x03;
        for (int i = 0;  i < bits_per_symbol;  i++)
        {
            put_bit(raw);
            raw >>= 1;
        }
        return constellation_state;
    }

    SB_HD void cdba(int bit, float &tre, float &tim) const
    {
        // src/v17rx.c:602-608
        tre = (bit == 0)  ?  6.0f  :  (bit == 1)  ?  -2.0f  :  (bit == 2)  ?  2.0f  :  -6.0f;
        tim = (bit == 0)  ?  2.0f  :  (bit == 1)  ?  6.0f  :  (bit == 2)  ?  -6.0f  :  -2.0f;
    }

    SB_HD float spacing() const
    {
        // constellation_spacing[4], src/v17rx.c:145-155
        return (space_map == 0)  ?  1.414f  :  (space_map == 1)  ?  2.0f  :  (space_map == 2)  ?  2.828f  :  4.0f;
    }

    SB_HD void park(bool clear_agc)
    {
        if (clear_agc)
            agc_scaling_save = 0.0f;
        training_stage = STAGE_PARKED;
        report_status(SIG_STATUS_TRAINING_FAILED);
    }

    SB_HD int next_cdba_bits()
    {
        int bit = descramble(1);
        bit = (bit << 1) | descramble(1);
        return bit;
    }

    SB_HD void add_training_error(float zre, float zim, float tre, float tim, bool accumulate)
    {
        const float dr = fsub(zre, tre);
        const float di = fsub(zim, tim);
        const float p = fadd(fmul(dr, dr), fmul(di, di));
        training_error = (accumulate)  ?  fadd(training_error, p)  :  p;
    }

    // src/v17rx.c:649-1133: the once-per-baud part of process_half_baud()
    SB_HD void process_baud(const Consts &k)
    {
        eq_put_step += godard_per_baud(k);
        float zre;
        float zim;
        equalizer_get(zre, zim);
        float tre = 0.0f;
        float tim = 0.0f;
        int constellation_state = 0;

        switch (training_stage)
        {
        case STAGE_NORMAL:
            constellation_state = decode_baud(k, zre, zim);
            tre = con_re(constellation_state);
            tim = con_im(constellation_state);
            break;
        case STAGE_SYMBOL_ACQUISITION:
            if (++training_count >= 100)
            {
                training_stage = STAGE_LOG_PHASE;
                for (int i = 0;  i < 16;  i++)
                    diff_angles[i*32] = 0;
                last_angle0 = arctan2(zim, zre);
                if (agc_scaling_save == 0.0f)
                    agc_scaling_save = agc_scaling;
            }
            break;
        case STAGE_LOG_PHASE:
            {
                int angle = arctan2(zim, zre);
                training_count = 1;
                if (short_train)
                {
                    // We should already know the accurate carrier frequency; only the phase is needed.
                    if ((unsigned int) (angle - last_angle0) < (unsigned int) k.phase_180)
                    {
                        angle = last_angle0;
                        last_angle0 = k.phase_a;
                        last_angle1 = k.phase_b;
                    }
                    else
                    {
                        last_angle0 = k.phase_b;
                        last_angle1 = k.phase_a;
                    }
                    const unsigned int phase_step = (unsigned int) angle - (unsigned int) k.phase_b;
                    spin_equalizer_buffer(phase_step);
                    track_p = 500000.0f;
                    carrier_phase += phase_step;
                    training_stage = STAGE_SHORT_WAIT_FOR_CDBA;
                }
                else
                {
                    last_angle1 = angle;
                    training_stage = STAGE_WAIT_FOR_CDBA;
                }
            }
            break;
        case STAGE_WAIT_FOR_CDBA:
            {
                const int angle = arctan2(zim, zre);
                int i = training_count + 1;
                int ang = angle - ((i & 1)  ?  last_angle1  :  last_angle0);
                if (i & 1)
                    last_angle1 = angle;
                else
                    last_angle0 = angle;
                diff_angles[(i & 0xF)*32] = diff_angles[((i - 2) & 0xF)*32] + (ang >> 4);
                if ((ang > k.phase_p90  ||  ang < k.phase_m90)  &&  training_count >= 13)
                {
                    // A phase reversal: slam the carrier frequency into line (src/v17rx.c:743-768)
                    i = (training_count - 8) & ~1;
                    if (i > 1)
                    {
                        const int j = i & 0xF;
                        ang = (diff_angles[j*32] + diff_angles[(j | 0x1)*32])/(i - 1);
                        phase_rate += 3*16*(ang/20);
                    }
                    if (phase_rate < k.rate_low  ||  phase_rate > k.rate_high)
                    {
                        park(true);
                        break;
                    }
                    const unsigned int phase_step = (unsigned int) angle - (unsigned int) k.phase_c;
                    spin_equalizer_buffer(phase_step);
                    carrier_phase += phase_step;
                    // The first symbol of the scrambled sequence has just been seen, so skip it
                    cdba(next_cdba_bits(), tre, tim);
                    training_count = 1;
                    training_stage = STAGE_COARSE_TRAIN_ON_CDBA;
                    report_status(SIG_STATUS_TRAINING_IN_PROGRESS);
                    break;
                }
                if (++training_count > V17_TRAINING_SEG_1_LEN)
                    park(true);
            }
            break;
        case STAGE_COARSE_TRAIN_ON_CDBA:
            cdba(next_cdba_bits(), tre, tim);
            track_carrier(zre, zim, tre, tim);
            tune_equalizer(zre, zim, tre, tim);
            add_training_error(zre, zim, tre, tim, false);
            if (++training_count == V17_TRAINING_SEG_2_LEN - 2000  ||  training_error < 1.0f  ||  training_error > 200.0f)
            {
                eq_delta = k.eq_delta_slow;
                track_i = 1000.0f;
                training_stage = STAGE_FINE_TRAIN_ON_CDBA;
            }
            break;
        case STAGE_FINE_TRAIN_ON_CDBA:
            cdba(next_cdba_bits(), tre, tim);
            track_carrier(zre, zim, tre, tim);
            tune_equalizer(zre, zim, tre, tim);
            if (++training_count >= V17_TRAINING_SEG_2_LEN - 48)
            {
                training_error = 0.0f;
                track_i = 100.0f;
                track_p = 500000.0f;
                training_stage = STAGE_TRAIN_ON_CDBA_AND_TEST;
            }
            break;
        case STAGE_TRAIN_ON_CDBA_AND_TEST:
            cdba(next_cdba_bits(), tre, tim);
            if (++training_count < V17_TRAINING_SEG_2_LEN - 20)
            {
                track_carrier(zre, zim, tre, tim);
                tune_equalizer(zre, zim, tre, tim);
                add_training_error(zre, zim, tre, tim, true);
            }
            else if (training_count >= V17_TRAINING_SEG_2_LEN)
            {
                if (training_error < fmul(fmul(20.0f, 1.414f), spacing()))
                {
                    training_error = 0.0f;
                    training_count = 0;
                    training_stage = STAGE_BRIDGE;
                }
                else
                {
                    park(true);
                }
            }
            break;
        case STAGE_BRIDGE:
            descramble(V17_BRIDGE_WORD >> ((training_count & 0x7) << 1));
            descramble(V17_BRIDGE_WORD >> (((training_count & 0x7) << 1) + 1));
            tre = zre;
            tim = zim;
            if (++training_count >= V17_TRAINING_SEG_3_LEN)
            {
                training_error = 0.0f;
                training_count = 0;
                if (bits_per_symbol == 2)
                {
                    diff = (short_train)  ?  0  :  1;
                    training_stage = STAGE_TEST_ONES;
                }
                else
                {
                    training_stage = STAGE_TCM_WINDUP;
                }
            }
            break;
        case STAGE_SHORT_WAIT_FOR_CDBA:
            {
                const int angle = arctan2(zim, zre);
                const int ang = angle - ((training_count & 1)  ?  last_angle1  :  last_angle0);
                if (ang > k.phase_p90  ||  ang < k.phase_m90)
                {
                    cdba(next_cdba_bits(), tre, tim);
                    training_error = 0.0f;
                    training_count = 1;
                    training_stage = STAGE_SHORT_TRAIN_ON_CDBA_AND_TEST;
                }
                else
                {
                    cdba((training_count & 1) + 2, tre, tim);
                    track_carrier(zre, zim, tre, tim);
                    if (++training_count > V17_TRAINING_SEG_1_LEN)
                        park(false);
                }
            }
            break;
        case STAGE_SHORT_TRAIN_ON_CDBA_AND_TEST:
            cdba(next_cdba_bits(), tre, tim);
            track_carrier(zre, zim, tre, tim);
            if (training_count > 8)
                add_training_error(zre, zim, tre, tim, true);
            if (++training_count >= V17_TRAINING_SHORT_SEG_2_LEN)
            {
                track_i = 100.0f;
                track_p = 500000.0f;
                if (training_error < fmul(fmul(fmul((float) (V17_TRAINING_SHORT_SEG_2_LEN - 8), 4.0f), 1.0f), spacing()))
                {
                    training_count = 0;
                    if (bits_per_symbol == 2)
                    {
                        diff = (short_train)  ?  0  :  1;
                        training_error = 0.0f;
                        training_stage = STAGE_TEST_ONES;
                    }
                    else
                    {
                        training_stage = STAGE_TCM_WINDUP;
                    }
                    report_status(SIG_STATUS_TRAINING_IN_PROGRESS);
                }
                else
                {
                    park(false);
                }
            }
            break;
        case STAGE_TCM_WINDUP:
            constellation_state = decode_baud(k, zre, zim);
            tre = con_re(constellation_state);
            tim = con_im(constellation_state);
            add_training_error(zre, zim, tre, tim, true);
            if (++training_count >= V17_TRAINING_SEG_4A_LEN)
            {
                training_error = 0.0f;
                training_count = 0;
                diff = (short_train)  ?  0  :  1;
                training_stage = STAGE_TEST_ONES;
            }
            break;
        case STAGE_TEST_ONES:
            constellation_state = decode_baud(k, zre, zim);
            tre = con_re(constellation_state);
            tim = con_im(constellation_state);
            add_training_error(zre, zim, tre, tim, true);
            if (++training_count >= V17_TRAINING_SEG_4_LEN)
            {
                if (training_error < fmul(fmul(fmul((float) V17_TRAINING_SEG_4_LEN, 1.0f), 1.0f), spacing()))
                {
                    report_status(SIG_STATUS_TRAINING_SUCCEEDED);
                    signal_present = 60;
                    equalizer_save();
                    phase_rate_save = phase_rate;
                    short_train = 1;
                    training_stage = STAGE_NORMAL;
                }
                else
                {
                    park(!short_train);
                }
            }
            break;
        default:
            break;
        }
        report_symbol(zre, zim, tre, tim, constellation_state);
    }
};

}  // namespace sbm
