/*
 * spandsp/tone_detect.h - so that a caller written against the reference compiles unchanged with -I<this repo>/include:
 * what src/spandsp/tone_detect.h: make_goertzel_descriptor, goertzel_* declares is declared, for the paths this library
 * replaces, by spandsp_b200_dropin.h.
 */
#if !defined(_SPANDSP_B200_FWD_TONE_DETECT_H_)
#define _SPANDSP_B200_FWD_TONE_DETECT_H_

#include "../spandsp_b200_dropin.h"

#endif
