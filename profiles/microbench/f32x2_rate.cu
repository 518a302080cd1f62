// f32x2_rate.cu -- issue-rate microbenchmark on sm_100a: scalar FFMA/FADD vs packed FFMA2/FADD2 (fma.rn.f32x2 / add.rn.f32x2),
// alone and interleaved with shared-memory loads.  Prints cycles per warp-instruction per SM sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o f32x2_rate f32x2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fadd1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

template <int MODE>
__global__ void k(float *out, long long *cyc, int iters) {
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-3f;
    __syncthreads();
    float a[8];
    u64 p[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 0.5f); }
    const float b = 1.0001f, c = 0.0001f;
    const u64 pb = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b), pc = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {          // 16 scalar FFMA
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = ffma1(a[i], b, c);
        } else if (MODE == 1) {   // 8 packed FFMA2 (= 16 FMAs)
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], pb, pc);
        } else if (MODE == 2) {   // 16 scalar FADD
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fadd1(a[i], c);
        } else if (MODE == 3) {   // 8 packed FADD2
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fadd2(p[i], pc);
        } else if (MODE == 4) {   // 16 scalar FFMA + 8 LDS.64
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = ffma1(a[i], b, c);
                float2 v; { unsigned ad = (unsigned)__cvta_generic_to_shared(&sm[((threadIdx.x + it + i * 64) & 2047) * 2]); asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(ad)); }
                acc += v.x;
                a[i] = ffma1(a[i], b, v.y);
            }
        } else if (MODE == 5) {   // 8 packed FFMA2 + 8 LDS.64  (same flops as MODE 4 minus the acc adds)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float2 v; { unsigned ad = (unsigned)__cvta_generic_to_shared(&sm[((threadIdx.x + it + i * 64) & 2047) * 2]); asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(ad)); }
                u64 pv = ((u64)__float_as_uint(v.y) << 32) | __float_as_uint(v.x);
                p[i] = ffma2(p[i], pb, pv);
            }
        }
    }
    long long t1 = clock64();
    float s = acc;
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int fp_per_iter, int threads) {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 20000;
    k<MODE><<<148, threads>>>(out, cyc, 10);
    k<MODE><<<148, threads>>>(out, cyc, iters);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double warps_per_smsp = threads / 32 / 4.0;
    printf("%-34s threads/SM %4d: %.3f cycles per FP warp-instr per SMSP (%d FP instr/iter)\n", name, threads,
           c / ((double)iters * fp_per_iter * warps_per_smsp), fp_per_iter);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int th : {128, 512, 1024}) {
        run<0>("FFMA  x16", 16, th);
        run<1>("FFMA2 x8 (16 FMAs)", 8, th);
        run<2>("FADD  x16", 16, th);
        run<3>("FADD2 x8", 8, th);
        run<4>("FFMA x16 + LDS.64 x8 + FADD x8", 16, th);
        run<5>("FFMA2 x8 + LDS.64 x8", 8, th);
    }
    return 0;
}
