// pipe_microbench.cu - measures the issue rate of the instructions the Goertzel bank is made of
// (FMUL, FADD, FADD2, FFMA, FFMA2, I2F.S16) on the device it runs on.  Output: warp-instructions
// per clock per SM for 4/8/16 warps per SM.  Used to decide scalar vs f32x2 adds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_microbench pipe_microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

typedef unsigned long long u64;

#define ITERS 4096

template <int OP>
__global__ void bench(float *out, u64 *cycles, float seed)
{
    float a[8];
    u64 p[8];
    int w[8];
#pragma unroll
    for (int i = 0;  i < 8;  i++)
    {
        a[i] = seed + i + threadIdx.x;
        p[i] = ((u64) __float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.0f);
        w[i] = (int) (seed*1000) + i*77 + threadIdx.x;
    }
    const float m = 0.999f + seed*1e-9f;
    const u64 m2 = ((u64) __float_as_uint(m) << 32) | __float_as_uint(m);
    __syncthreads();
    const u64 t0 = clock64();
#pragma unroll 1
    for (int it = 0;  it < ITERS;  it++)
    {
#pragma unroll
        for (int i = 0;  i < 8;  i++)
        {
            if (OP == 0) a[i] = __fmul_rn(a[i], m);
            if (OP == 1) a[i] = __fadd_rn(a[i], m);
            if (OP == 2) a[i] = __fmaf_rn(a[i], m, m);
            if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(m2));
            if (OP == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(m2));
            if (OP == 5) { short s = (short) (w[i] & 0xFFFF); float f = (float) s; w[i] = __float_as_int(f) ^ it; }
            if (OP == 6)
            {   // the bank's mix: 2 FMUL + 1 FADD2(sub) + 1 FADD2(add) per pair
                float lo = __uint_as_float((unsigned) (p[i] & 0xFFFFFFFFu));
                float hi = __uint_as_float((unsigned) (p[i] >> 32));
                lo = __fmul_rn(lo, m);
                hi = __fmul_rn(hi, m);
                u64 q = ((u64) __float_as_uint(hi) << 32) | __float_as_uint(lo);
                asm volatile("sub.rn.f32x2 %0, %0, %1;" : "+l"(q) : "l"(m2));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q) : "l"(m2));
                p[i] = q;
            }
            if (OP == 7)
            {   // scalar version of the same: 2 FMUL + 4 FADD
                float lo = __uint_as_float((unsigned) (p[i] & 0xFFFFFFFFu));
                float hi = __uint_as_float((unsigned) (p[i] >> 32));
                lo = __fadd_rn(__fsub_rn(__fmul_rn(lo, m), m), m);
                hi = __fadd_rn(__fsub_rn(__fmul_rn(hi, m), m), m);
                p[i] = ((u64) __float_as_uint(hi) << 32) | __float_as_uint(lo);
            }
        }
    }
    const u64 t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0;  i < 8;  i++)
        s += a[i] + (float) (p[i] & 0xFFFF) + w[i];
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0)
        cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(const char *name, int instr_per_slot, int sms)
{
    float *out;
    u64 *cyc;
    cudaMalloc(&out, sizeof(float)*sms*1024);
    cudaMalloc(&cyc, sizeof(u64)*sms);
    for (int threads = 128;  threads <= 1024;  threads *= 2)
    {
        bench<OP><<<sms, threads>>>(out, cyc, 1.0f);
        bench<OP><<<sms, threads>>>(out, cyc, 1.0f);
        cudaDeviceSynchronize();
        u64 h[256];
        cudaMemcpy(h, cyc, sizeof(u64)*sms, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0;  i < sms;  i++)
            avg += (double) h[i];
        avg /= sms;
        const double winst = (double) ITERS*8*instr_per_slot*(threads/32);
        printf("%-28s warps/SM=%2d  cycles=%9.0f  warp-instr/clk/SM=%6.3f\n", name, threads/32, avg, winst/avg);
    }
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device %s, %d SMs, sm_%d%d, clock %d kHz\n", p.name, p.multiProcessorCount, p.major, p.minor, p.clockRate);
    const int sms = p.multiProcessorCount;
    run<0>("FMUL", 1, sms);
    run<1>("FADD", 1, sms);
    run<2>("FFMA", 1, sms);
    run<3>("FADD2 (add.rn.f32x2)", 1, sms);
    run<4>("FFMA2 (fma.rn.f32x2)", 1, sms);
    run<5>("I2F.S16 (+LOP)", 2, sms);
    run<6>("pair: 2 FMUL + 2 FADD2", 4, sms);
    run<7>("pair: 2 FMUL + 4 FADD", 6, sms);
    return 0;
}
