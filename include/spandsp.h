/*
 * spandsp.h - umbrella header under the reference's name (src/spandsp.h.in): a caller that includes <spandsp.h> and uses
 * the receive paths this library replaces compiles unchanged with -I<this repo>/include and links against
 * libspandsp_b200.so.
 */
#if !defined(_SPANDSP_B200_FWD_SPANDSP_H_)
#define _SPANDSP_B200_FWD_SPANDSP_H_

#include "spandsp_b200_dropin.h"

#endif
