/*
 * spandsp/complex.h - so that a caller written against the reference compiles unchanged with -I<this repo>/include:
 * what src/spandsp/complex.h: complexf_t declares is declared, for the paths this library
 * replaces, by spandsp_b200_dropin.h.
 */
#if !defined(_SPANDSP_B200_FWD_COMPLEX_H_)
#define _SPANDSP_B200_FWD_COMPLEX_H_

#include "../spandsp_b200_dropin.h"

#endif
