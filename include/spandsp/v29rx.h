/*
 * spandsp/v29rx.h - so that a caller written against the reference compiles unchanged with -I<this repo>/include:
 * what src/spandsp/v29rx.h: v29_rx_* declares is declared, for the paths this library
 * replaces, by spandsp_b200_dropin.h.
 */
#if !defined(_SPANDSP_B200_FWD_V29RX_H_)
#define _SPANDSP_B200_FWD_V29RX_H_

#include "../spandsp_b200_dropin.h"

#endif
