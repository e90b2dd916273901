// sb_engine.h - internal helpers shared between the engine and the drop-in layer.
#pragma once

// The library is built with -fvisibility=hidden; only what the public headers declare is exported.
#pragma GCC visibility push(default)
#include "../../include/spandsp_b200.h"
#pragma GCC visibility pop

void sb_set_error(const char *fmt, ...);
float sb_goertzel_fac(float freq);
void *sb_ctx_stream(span_b200_ctx_t *ctx);
int sb_ulaw_to_linear(unsigned char ulaw);
int sb_alaw_to_linear(unsigned char alaw);
