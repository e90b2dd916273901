/*
 * spandsp/bell_r2_mf.h - so that a caller written against the reference compiles unchanged with -I<this repo>/include:
 * what src/spandsp/bell_r2_mf.h: bell_mf_rx_*, r2_mf_rx_* declares is declared, for the paths this library
 * replaces, by spandsp_b200_dropin.h.
 */
#if !defined(_SPANDSP_B200_FWD_BELL_R2_MF_H_)
#define _SPANDSP_B200_FWD_BELL_R2_MF_H_

#include "../spandsp_b200_dropin.h"

#endif
