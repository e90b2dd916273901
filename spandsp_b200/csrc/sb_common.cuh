// sb_common.cuh - device helpers shared by the tone-bank kernels (sm_100a only).
//
// Arithmetic contract: every value that feeds a detection decision is computed with
// explicitly rounded, never-contracted IEEE-754 binary32 operations (__fmul_rn / __fadd_rn /
// __fsub_rn, or their f32x2 forms), in the operand order of the reference source, so block
// energies are bit-identical to the strict CPU oracle.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "spandsp_b200 kernels are written for sm_100a only"
#endif

namespace sb {

typedef unsigned long long u64;

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

// ---- two-bin packed state --------------------------------------------------------------
// A pair holds two independent Goertzel bins (for DTMF: row_i in .x, col_i in .y).  On sm_100a
// add/sub have a 2-wide form (FADD2) that halves the issue slots of the recurrence.  The 2-wide
// multiply cannot be written as mul.rn.f32x2: ptxas 12.9 contracts mul.rn.f32x2 + sub.rn.f32x2 into a
// single FFMA2 even with explicit .rn and --fmad=false, which would change the rounding.  pmul_packed()
// below gets the 2-wide multiply another way.
struct pair_t
{
    float x;
    float y;
};

__device__ __forceinline__ u64 pack2(pair_t a)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}

__device__ __forceinline__ pair_t unpack2(u64 v)
{
    pair_t r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

template <bool PACKED>
__device__ __forceinline__ pair_t psub(pair_t a, pair_t b)
{
    if (PACKED)
    {
        u64 r;
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a)), "l"(pack2(b)));
        return unpack2(r);
    }
    pair_t r;
    r.x = fsub(a.x, b.x);
    r.y = fsub(a.y, b.y);
    return r;
}

template <bool PACKED>
__device__ __forceinline__ pair_t padd_scalar(pair_t a, float s)
{
    if (PACKED)
    {
        pair_t b;
        b.x = s;
        b.y = s;
        u64 r;
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(a)), "l"(pack2(b)));
        return unpack2(r);
    }
    pair_t r;
    r.x = fadd(a.x, s);
    r.y = fadd(a.y, s);
    return r;
}

__device__ __forceinline__ pair_t pmul(pair_t a, pair_t b)
{
    pair_t r;
    r.x = fmul(a.x, b.x);
    r.y = fmul(a.y, b.y);
    return r;
}

// The 2-wide multiply as an explicit FMA with a +0 addend: fma(a, b, +0) rounds the exact product once, i.e. it IS
// the IEEE product, except that a product of -0 comes out as +0.  ptxas keeps it as FFMA2 Rd, Ra, Rb, RZ (it has
// nothing to contract it with).  The one caller is the Goertzel recurrence v3 = (fac*v2 - v1) + x, where the sign
// of a zero product cannot reach v3: (+-0 - v1) differs only for v1 = +0 (giving -0 vs +0), and adding x - a
// sample, never -0 - maps both to the same value.
__device__ __forceinline__ pair_t pmul_packed(pair_t a, pair_t b)
{
    u64 r;
    u64 z;
    asm("mov.b64 %0, 0;" : "=l"(z));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pack2(a)), "l"(pack2(b)), "l"(z));
    return unpack2(r);
}

// ---- cp.async (LDGSTS) 16-byte copies with zero fill ------------------------------------
__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void *gptr, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(smem_addr), "l"(gptr), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ void cp_async_16_full(uint32_t smem_addr, const void *gptr)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(smem_addr), "l"(gptr) : "memory");
}

__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" :: "n"(N) : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t smem_addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr));
    return v;
}

// int16 sample e (0..7) of a 16-byte vector, as float.  Compiles to I2F.S16 with a .H0/.H1
// operand select, which runs on the conversion pipe, not the FP32 pipe.
template <int E>
__device__ __forceinline__ float sample_of(const uint4 &v)
{
    const uint32_t w = (E < 2)  ?  v.x  :  (E < 4)  ?  v.y  :  (E < 6)  ?  v.z  :  v.w;
    const short s = (short) ((E & 1)  ?  (w >> 16)  :  (w & 0xFFFFu));
    return (float) s;
}

// Shift a 16-byte vector right by one int16 lane (slow-path iteration helper).
__device__ __forceinline__ void shift_vec(uint4 &v)
{
    v.x = __funnelshift_r(v.x, v.y, 16);
    v.y = __funnelshift_r(v.y, v.z, 16);
    v.z = __funnelshift_r(v.z, v.w, 16);
    v.w = v.w >> 16;
}

// Shift a 16-byte vector right by one byte (companded input).
__device__ __forceinline__ void shift_vec8(uint4 &v)
{
    v.x = __funnelshift_r(v.x, v.y, 8);
    v.y = __funnelshift_r(v.y, v.z, 8);
    v.z = __funnelshift_r(v.z, v.w, 8);
    v.w = v.w >> 8;
}

__device__ __forceinline__ void shift_vec_n(uint4 &v, int lanes)
{
    for (int i = 0;  i < lanes;  i++)
        shift_vec(v);
}

}  // namespace sb
