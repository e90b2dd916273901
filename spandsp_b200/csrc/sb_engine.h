// sb_engine.h - internal helpers shared between the engine and the drop-in layer.
#pragma once

// The library is built with -fvisibility=hidden; only what the public headers declare is exported.
#pragma GCC visibility push(default)
#include "../../include/spandsp_b200.h"
#pragma GCC visibility pop

#include <cuda_runtime.h>

void sb_set_error(const char *fmt, ...);

// Every ABI entry point works on its context's device and leaves the caller's current device as it found it
// (a host application's later CUDA or torch work must not be redirected by a call into this library).
struct sb_device_guard
{
    int prev;
    int want;
    bool good;
    explicit sb_device_guard(int device) : prev(-1), want(device), good(true)
    {
        if (want < 0)
            return;                     // no object to act on: the caller's own argument check reports it
        if (cudaGetDevice(&prev) != cudaSuccess)
            prev = -1;
        if (prev != want)
            good = (cudaSetDevice(want) == cudaSuccess);
    }
    ~sb_device_guard()
    {
        if (want >= 0  &&  prev >= 0  &&  prev != want)
            cudaSetDevice(prev);
    }
    bool ok() const { return good; }
};

#define SB_DEVICE_CK(dev) \
    sb_device_guard sb_dg_(dev); \
    do \
    { \
        if (!sb_dg_.ok()) \
        { \
            sb_set_error("cudaSetDevice(%d) failed (%s:%d)", (int) (dev), __FILE__, __LINE__); \
            return -1; \
        } \
    } \
    while (0)

#define SB_DEVICE_CKP(dev) \
    sb_device_guard sb_dg_(dev); \
    do \
    { \
        if (!sb_dg_.ok()) \
        { \
            sb_set_error("cudaSetDevice(%d) failed (%s:%d)", (int) (dev), __FILE__, __LINE__); \
            return NULL; \
        } \
    } \
    while (0)

float sb_goertzel_fac(float freq);
void *sb_ctx_stream(span_b200_ctx_t *ctx);
int sb_ulaw_to_linear(unsigned char ulaw);
int sb_alaw_to_linear(unsigned char alaw);
