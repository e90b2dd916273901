// sb_v29_rx.cuh - the V.29 receiver on top of the shared core: training state machine, slicer,
// differential decoder, descrambler.  Reference: src/v29rx.c:350-785,1019-1097.
#pragma once

#include "sb_modem.cuh"

namespace sbm {

#define V29_COEFF_SETS      48

// Small lookup tables, staged in shared memory (lanes index them divergently).
struct V29Tables
{
    float constellation[16][2];         // src/v29tx_constellation_maps.h:57-79
    int cdcd_pos[6];                    // src/v29rx.c:488-493
    unsigned char space_map[20][20];    // src/v29rx.c:119-143
    unsigned char phase_steps_9600[8];  // src/v29rx.c:404-407
    unsigned char phase_steps_4800[4];  // src/v29rx.c:408-411
};

struct V29Consts : CoreConsts
{
    const V29Tables *tables;            // global memory
    int phase_p45;                      // DDS_PHASE(45.0f)
    int phase_m45;                      // DDS_PHASE(-45.0f)
    float eq_delta;                     // 0.21f/33
};

// The slicer map: nearest constellation point for each 0.5 x 0.5 cell of [-5, 5) x [-5, 5), evaluated
// at the cell centre, first index winning ties (the rule of src/make_v29_constellation_map.c:63-84).
// The table actually compiled into the reference (src/v29rx.c:119-143) deviates from that rule in
// eight cells on the diagonals, which are listed explicitly.
static inline void make_v29_tables(V29Tables &t)
{
    static const float constellation[16][2] =
    {
        { 3.0f,  0.0f}, { 1.0f,  1.0f}, { 0.0f,  3.0f}, {-1.0f,  1.0f}, {-3.0f,  0.0f}, {-1.0f, -1.0f}, { 0.0f, -3.0f}, { 1.0f, -1.0f},
        { 5.0f,  0.0f}, { 3.0f,  3.0f}, { 0.0f,  5.0f}, {-3.0f,  3.0f}, {-5.0f,  0.0f}, {-3.0f, -3.0f}, { 0.0f, -5.0f}, { 3.0f, -3.0f}
    };
    memset(&t, 0, sizeof(t));
    memcpy(t.constellation, constellation, sizeof(constellation));
    for (int ire = 0;  ire < 20;  ire++)
    {
        const double re = (ire - 10)/2.0 + 0.25;
        for (int iim = 0;  iim < 20;  iim++)
        {
            const double im = (iim - 10)/2.0 + 0.25;
            int best = 0;
            double best_distance = 1000000.0;
            for (int l = 0;  l < 16;  l++)
            {
                const double d = (re - constellation[l][0])*(re - constellation[l][0])
                               + (im - constellation[l][1])*(im - constellation[l][1]);
                if (d < best_distance)
                {
                    best = l;
                    best_distance = d;
                }
            }
            t.space_map[ire][iim] = (unsigned char) best;
        }
    }
    static const unsigned char exceptions[8][3] =
    {
        {5, 6, 13}, {5, 13, 11}, {6, 5, 13}, {6, 14, 11}, {13, 5, 15}, {13, 14, 9}, {14, 6, 15}, {14, 13, 9}
    };
    for (int i = 0;  i < 8;  i++)
        t.space_map[exceptions[i][0]][exceptions[i][1]] = exceptions[i][2];
    static const unsigned char ps9600[8] = {4, 0, 2, 6, 7, 3, 1, 5};
    static const unsigned char ps4800[4] = {0, 2, 3, 1};
    static const int cdcd[6] = {0, 11, 0, 3, 0, 2};
    memcpy(t.phase_steps_9600, ps9600, sizeof(ps9600));
    memcpy(t.phase_steps_4800, ps4800, sizeof(ps4800));
    memcpy(t.cdcd_pos, cdcd, sizeof(cdcd));
}

#define SBM_V29_NAME    RxV29
#define SBM_V29_LPC     1
#include "sb_v29_rx_body.cuh"
#undef SBM_V29_NAME
#undef SBM_V29_LPC

#if defined(__CUDACC__)
// Four lanes per receiver: the kernel the banks run (the state arrays in global memory are the same)
#define SBM_V29_NAME    RxV29x4
#define SBM_V29_LPC     4
#include "sb_v29_rx_body.cuh"
#undef SBM_V29_NAME
#undef SBM_V29_LPC
#define SBM_HAVE_V29X4
#endif

}  // namespace sbm
