/*
 * spandsp/async.h - so that a caller written against the reference compiles unchanged with -I<this repo>/include:
 * what src/spandsp/async.h: SIG_STATUS_*, span_put_bit_func_t, span_modem_status_func_t declares is declared, for the paths this library
 * replaces, by spandsp_b200_dropin.h.
 */
#if !defined(_SPANDSP_B200_FWD_ASYNC_H_)
#define _SPANDSP_B200_FWD_ASYNC_H_

#include "../spandsp_b200_dropin.h"

#endif
