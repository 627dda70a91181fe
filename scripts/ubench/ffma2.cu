// Microbenchmark: issue cost of FFMA2 (fma.rn.f32x2) against scalar FFMA / FMUL+FADD on sm_100a.
// Each thread runs NCHAIN independent dependency chains of ITER operations; 8 warps per SM sub-partition (1024 threads per SM x 1 CTA,
// 148 CTAs).  Prints warp-instructions per cycle per SMSP for each variant (clock64 on the device).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
constexpr int ITER = 4096, NCHAIN = 8;
__device__ __forceinline__ u64 pfma(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float sfma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float smul(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float sadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint32_t iadd(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int MODE> __global__ void __launch_bounds__(1024) k(u64 *out, long long *cyc, u64 kb, u64 kc, float fb, float fc)
{
    u64 p[NCHAIN]; float f[NCHAIN]; uint32_t q[NCHAIN];
    for (int i = 0; i < NCHAIN; ++i) { p[i] = kb + threadIdx.x + i; f[i] = fb + threadIdx.x + i; q[i] = threadIdx.x + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) {
            if (MODE == 0) p[i] = pfma(p[i], kb, kc);                         // FFMA2, 1 per step
            if (MODE == 1) f[i] = sfma(f[i], fb, fc);                         // FFMA, 1 per step
            if (MODE == 2) { f[i] = smul(f[i], fb); f[i] = sadd(f[i], fc); }  // FMUL + FADD, 2 per step
            if (MODE == 3) { p[i] = pfma(p[i], kb, kc); q[i] = iadd(q[i], 3u); }   // FFMA2 + IADD: does the ALU pipe fill FFMA2's second slot?
            if (MODE == 4) { f[i] = sfma(f[i], fb, fc); q[i] = iadd(q[i], 3u); }   // FFMA + IADD
            if (MODE == 5) q[i] = iadd(q[i], 3u);                             // IADD alone
        }
    }
    const long long t1 = clock64();
    u64 acc = 0;
    for (int i = 0; i < NCHAIN; ++i) acc += p[i] + (u64)f[i] + q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name, int per_step, u64 *out, long long *cyc)
{
    k<MODE><<<148, 1024>>>(out, cyc, 0x3f8000013f800001ull, 0x3a0000003a000000ull, 1.0000001f, 0.0005f);
    cudaDeviceSynchronize();
    k<MODE><<<148, 1024>>>(out, cyc, 0x3f8000013f800001ull, 0x3a0000003a000000ull, 1.0000001f, 0.0005f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += (double)h[i]; c /= 148;
    const double winst = (double)ITER * NCHAIN * per_step * 32.0 / 4.0;       // warp instructions per SMSP (32 warps per SM = 8 per SMSP)
    printf("%-14s cycles %.0f  warp-inst/cycle/SMSP %.3f  (cycles per listed instruction %.3f)\n", name, c, winst / c, c / winst);
}
int main()
{
    u64 *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    run<0>("FFMA2", 1, out, cyc);
    run<1>("FFMA", 1, out, cyc);
    run<2>("FMUL+FADD", 2, out, cyc);
    run<3>("FFMA2+IADD", 2, out, cyc);
    run<4>("FFMA+IADD", 2, out, cyc);
    run<5>("IADD", 1, out, cyc);
    return 0;
}
