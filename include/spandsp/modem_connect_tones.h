/*
 * spandsp/modem_connect_tones.h - so that a caller written against the reference compiles unchanged with -I<this repo>/include:
 * what src/spandsp/modem_connect_tones.h: modem_connect_tones_rx_* declares is declared, for the paths this library
 * replaces, by spandsp_b200_dropin.h.
 */
#if !defined(_SPANDSP_B200_FWD_MODEM_CONNECT_TONES_H_)
#define _SPANDSP_B200_FWD_MODEM_CONNECT_TONES_H_

#include "../spandsp_b200_dropin.h"

#endif
