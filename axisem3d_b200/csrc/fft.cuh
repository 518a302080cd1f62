// fft.cuh -- in-shared-memory mixed-radix complex FFT passes for the azimuthal transforms.
//
// Replaces SolverFFTW_{1,3,N3,N6,N9}::computeC2R/computeR2C (S/core/fftw/SolverFFTW_N6.cpp:45-53),
// i.e. FFTW's fftwf_plan_many_dft_c2r / r2c on lucky-number lengths N = 2^a 3^b 5^c 7^d (11|13)^{<=1}
// (S/preloop/utilities/PreloopFFTW.cpp:59-99).
//
// Design (DESIGN.md "FFT"):
//  * two real columns are transformed as ONE complex column z = x + i y ("two-for-one"), which works
//    for odd and even N alike; the Hermitian fill-in / split is fused into the producers/consumers;
//  * c2r runs as an in-place decimation-in-frequency transform (natural -> digit-reversed order),
//    r2c as the in-place decimation-in-time transpose (digit-reversed -> natural).  The pointwise
//    physics in between does not care about the sample order, so no reordering pass exists: all
//    phi-dependent material arrays are stored digit-reversed at upload time;
//  * one thread owns one radix-R butterfly in registers; passes are separated by __syncthreads().
#pragma once
#include <cuda_runtime.h>

#define AX_MAX_STAGES 10

struct FftPlan {
    int N;
    int nstages;
    int radix[AX_MAX_STAGES];
    int tw_off;    // offset (float2) of W_N[k] = exp(+2 pi i k / N), k < N, in the twiddle pool
    int perm_off;  // offset (int) of perm[pos] = sample index n stored at position pos after the DIF c2r
    // per-stage twiddle tables for the fused element kernel (fused.cuh): stage s with block length L_s and radix R_s
    // owns T_s[j * R_s + p] = exp(+2 pi i j p / L_s), j < L_s / R_s, p < R_s, at stwpool[stw_off[s]] (-1: no twiddles,
    // i.e. L_s == R_s).  All tables of one plan are contiguous: [stw_base, stw_base + stw_len).
    int stw_off[AX_MAX_STAGES];
    int stw_base, stw_len;
    // the same tables once more in p-major order, T2_s[p * (L_s / R_s) + j], for the warp-per-point kernel (fused_wp.cuh: lanes
    // run over j, so the p-major rows are read conflict-free): table of stage s at stwpool[stw_off[s] + stw2_delta]
    int stw2_delta;
};

// Radix sequence of the DIF transform of length N (shared by the host planner and the compile-time specialised
// kernels, which must agree because phi-dependent arrays are uploaded in the digit-reversed order of this plan):
// powers of two first (16s, then the 8/4/2 remainder), then odd primes 13, 11, 7, 5, 3 -- so that the LAST DIF
// stage (stride-1 butterflies) has an odd radix whenever N has an odd factor.  n == 0 marks an unsupported N.
// maxr2 = 16: powers of two as 16s + remainder (one thread per butterfly, many columns per CTA: fused.cuh, kernels.cuh);
// maxr2 = 8: as 8s and 4s (2^4 -> 4 4, 2^5 -> 8 4, 2^6 -> 8 8, ...) for the warp-per-point kernel, whose 72-register
// budget and 3 columns per warp favour small butterflies (fused_wp.cuh).
struct RadixList {
    int n;
    int r[AX_MAX_STAGES];
};
__host__ __device__ constexpr RadixList choose_radices_ct(int N, int maxr2 = 16) {
    RadixList out{0, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};
    int n = N, e = 0;
    while (n % 2 == 0 && n > 0) { n /= 2; ++e; }
    if (maxr2 >= 16) {
        while (e >= 4) { if (out.n < AX_MAX_STAGES) out.r[out.n] = 16; ++out.n; e -= 4; }
    } else {
        while (e > 4 || e == 3) { if (out.n < AX_MAX_STAGES) out.r[out.n] = 8; ++out.n; e -= 3; }
        while (e >= 2) { if (out.n < AX_MAX_STAGES) out.r[out.n] = 4; ++out.n; e -= 2; }
    }
    if (e == 3) { if (out.n < AX_MAX_STAGES) out.r[out.n] = 8; ++out.n; }
    else if (e == 2) { if (out.n < AX_MAX_STAGES) out.r[out.n] = 4; ++out.n; }
    else if (e == 1) { if (out.n < AX_MAX_STAGES) out.r[out.n] = 2; ++out.n; }
    const int ps[5] = {13, 11, 7, 5, 3};
    for (int i = 0; i < 5; ++i)
        while (n % ps[i] == 0) { if (out.n < AX_MAX_STAGES) out.r[out.n] = ps[i]; ++out.n; n /= ps[i]; }
    if (n != 1 || out.n > AX_MAX_STAGES) out.n = -1;
    return out;
}

// cos/sin(2 pi k / R) for the in-register butterflies, R <= 16
__constant__ float c_cos[17][16];
__constant__ float c_sin[17][16];

// Complex helpers.  sm_100a has packed fp32 arithmetic (PTX add/sub/mul/fma.rn.f32x2 -> SASS FADD2/FMUL2/FFMA2): one issue slot
// per complex add or real x complex FMA instead of two.  The fused element kernel is issue-bound, not FP32-pipe-bound
// (profiles/microbench/f32x2_rate.cu: FFMA2 = 2 cycles per warp-instruction per sub-partition, i.e. the same flops in half the
// issue slots), so every complex helper goes through the packed forms.  Rounding is identical to the scalar forms (rn).
#ifndef AX_F32X2
#define AX_F32X2 1
#endif
typedef unsigned long long ax_u64;
__device__ __forceinline__ ax_u64 f2_pack(float2 a) { ax_u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
__device__ __forceinline__ ax_u64 f2_splat(float s) { ax_u64 r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(s)); return r; }
__device__ __forceinline__ float2 f2_unpack(ax_u64 r) { float2 a; asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r)); return a; }
#if AX_F32X2
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    ax_u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b))); return f2_unpack(r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    ax_u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b))); return f2_unpack(r);
}
__device__ __forceinline__ float2 cscale(float2 a, float s) {
    ax_u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_splat(s))); return f2_unpack(r);
}
// acc + r * a (r real)
__device__ __forceinline__ float2 cfma(float r, float2 a, float2 acc) {
    ax_u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_splat(r)), "l"(f2_pack(a)), "l"(f2_pack(acc))); return f2_unpack(d);
}
// acc + a .* b (component-wise)
__device__ __forceinline__ float2 cfma2(float2 a, float2 b, float2 acc) {
    ax_u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(acc))); return f2_unpack(d);
}
__device__ __forceinline__ float2 cmul2(float2 a, float2 b) {
    ax_u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b))); return f2_unpack(d);
}
// a * b = a.x * (b.x, b.y) + a.y * (-b.y, b.x)
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return cfma(a.x, b, cscale(make_float2(-b.y, b.x), a.y));
}
// a * conj(b) = a.x * (b.x, -b.y) + a.y * (b.y, b.x)
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {
    return cfma(a.x, make_float2(b.x, -b.y), cscale(make_float2(b.y, b.x), a.y));
}
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float2 cfma(float r, float2 a, float2 acc) {
    return make_float2(fmaf(r, a.x, acc.x), fmaf(r, a.y, acc.y));
}
__device__ __forceinline__ float2 cfma2(float2 a, float2 b, float2 acc) {
    return make_float2(fmaf(a.x, b.x, acc.x), fmaf(a.y, b.y, acc.y));
}
__device__ __forceinline__ float2 cmul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
#endif
// multiply by SIGN * i
template <int SIGN>
__device__ __forceinline__ float2 cmul_i(float2 a) {
    return SIGN > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// ---------------------------------------------------------------- butterflies: y_p = sum_q a_q w^{pq},
// w = exp(SIGN * 2 pi i / R)
template <int R, int SIGN>
struct Dft;

template <int SIGN>
struct Dft<2, SIGN> {
    static __device__ __forceinline__ void run(float2 (&a)[2]) {
        float2 t = a[0];
        a[0] = cadd(t, a[1]);
        a[1] = csub(t, a[1]);
    }
};

template <int SIGN>
struct Dft<4, SIGN> {
    static __device__ __forceinline__ void run(float2 (&a)[4]) {
        float2 s02 = cadd(a[0], a[2]), d02 = csub(a[0], a[2]);
        float2 s13 = cadd(a[1], a[3]), d13 = cmul_i<SIGN>(csub(a[1], a[3]));
        a[0] = cadd(s02, s13);
        a[2] = csub(s02, s13);
        a[1] = cadd(d02, d13);
        a[3] = csub(d02, d13);
    }
};

template <int SIGN>
struct Dft<8, SIGN> {
    static __device__ __forceinline__ void run(float2 (&a)[8]) {
        float2 e[4] = {a[0], a[2], a[4], a[6]};
        float2 o[4] = {a[1], a[3], a[5], a[7]};
        Dft<4, SIGN>::run(e);
        Dft<4, SIGN>::run(o);
        const float h = 0.70710678118654752440f;
        // w8^p = exp(SIGN i pi p / 4)
        float2 w1 = make_float2(h, SIGN * h), w3 = make_float2(-h, SIGN * h);
        o[1] = cmul(o[1], w1);
        o[2] = cmul_i<SIGN>(o[2]);
        o[3] = cmul(o[3], w3);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            a[p] = cadd(e[p], o[p]);
            a[p + 4] = csub(e[p], o[p]);
        }
    }
};

template <int SIGN>
struct Dft<16, SIGN> {
    static __device__ __forceinline__ void run(float2 (&a)[16]) {
        float2 e[8], o[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            e[q] = a[2 * q];
            o[q] = a[2 * q + 1];
        }
        Dft<8, SIGN>::run(e);
        Dft<8, SIGN>::run(o);
#pragma unroll
        for (int p = 1; p < 8; ++p) {
            if (p == 4) {
                o[p] = cmul_i<SIGN>(o[p]);
            } else {
                o[p] = cmul(o[p], make_float2(c_cos[16][p], SIGN * c_sin[16][p]));
            }
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            a[p] = cadd(e[p], o[p]);
            a[p + 8] = csub(e[p], o[p]);
        }
    }
};

// odd prime radix via the symmetric/antisymmetric split (halves the multiplies)
template <int R, int SIGN>
struct DftOdd {
    static __device__ __forceinline__ void run(float2 (&a)[R]) {
        constexpr int H = (R - 1) / 2;
        float2 s[H], d[H];
#pragma unroll
        for (int q = 1; q <= H; ++q) {
            s[q - 1] = cadd(a[q], a[R - q]);
            d[q - 1] = csub(a[q], a[R - q]);
        }
        float2 a0 = a[0];
        float2 y0 = a0;
#pragma unroll
        for (int q = 0; q < H; ++q) y0 = cadd(y0, s[q]);
        a[0] = y0;
#pragma unroll
        for (int p = 1; p <= H; ++p) {
            float2 A = a0, B = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 1; q <= H; ++q) {
                const int k = (p * q) % R;
                const float c = c_cos[R][k], sn = c_sin[R][k];
                A = cfma(c, s[q - 1], A);
                B = cfma(sn, d[q - 1], B);
            }
            // y_p = A + SIGN i B ; y_{R-p} = A - SIGN i B
            float2 iB = cmul_i<SIGN>(B);
            a[p] = cadd(A, iB);
            a[R - p] = csub(A, iB);
        }
    }
};
template <int SIGN> struct Dft<3, SIGN> : DftOdd<3, SIGN> {};
template <int SIGN> struct Dft<5, SIGN> : DftOdd<5, SIGN> {};
template <int SIGN> struct Dft<7, SIGN> : DftOdd<7, SIGN> {};
template <int SIGN> struct Dft<11, SIGN> : DftOdd<11, SIGN> {};
template <int SIGN> struct Dft<13, SIGN> : DftOdd<13, SIGN> {};

// ---------------------------------------------------------------- one pass over `ncols` columns
// z: column c starts at z + c * ldz; tw = W_N table (shared or global); L = current block length.
// DIF (c2r, SIGN=+1): butterfly then twiddle w_L^{j p}.  DIT (r2c, SIGN=-1): twiddle w_L^{j q} then butterfly.
template <int R, int SIGN, bool DIF>
__device__ __forceinline__ void fft_pass(float2 *z, int ldz, int ncols, int N, int L,
                                         const float2 *__restrict__ tw, int tid, int nthreads) {
    const int Ls = L / R;
    const int nb = N / R;
    const int tstride = N / L;
    const int total = ncols * nb;
    for (int idx = tid; idx < total; idx += nthreads) {
        const int col = idx / nb;
        const int b = idx - col * nb;
        const int blk = b / Ls;
        const int j = b - blk * Ls;
        float2 *x = z + (size_t)col * ldz + blk * L + j;
        float2 a[R];
#pragma unroll
        for (int q = 0; q < R; ++q) a[q] = x[q * Ls];
        if (!DIF && j != 0) {
            const int step = j * tstride;
#pragma unroll
            for (int q = 1; q < R; ++q) {
                float2 w = tw[q * step];
                w.y = SIGN * w.y;
                a[q] = cmul(a[q], w);
            }
        }
        Dft<R, SIGN>::run(a);
        if (DIF && j != 0) {
            const int step = j * tstride;
#pragma unroll
            for (int p = 1; p < R; ++p) {
                float2 w = tw[p * step];
                w.y = SIGN * w.y;
                a[p] = cmul(a[p], w);
            }
        }
#pragma unroll
        for (int q = 0; q < R; ++q) x[q * Ls] = a[q];
    }
}

template <int SIGN, bool DIF>
__device__ __forceinline__ void fft_pass_dispatch(int R, float2 *z, int ldz, int ncols, int N, int L,
                                                  const float2 *__restrict__ tw, int tid, int nthreads) {
    switch (R) {
        case 2: fft_pass<2, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
        case 3: fft_pass<3, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
        case 4: fft_pass<4, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
        case 5: fft_pass<5, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
        case 7: fft_pass<7, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
        case 8: fft_pass<8, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
        case 11: fft_pass<11, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
        case 13: fft_pass<13, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
        case 16: fft_pass<16, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
        default: break;
    }
}

// c2r side: natural-order spectrum Z[k] -> samples z[n] stored at digit-reversed positions (unnormalised,
// sign +, like FFTW's backward transform).  Ends with a __syncthreads().
__device__ __forceinline__ void fft_inverse_dif(const FftPlan &pl, float2 *z, int ldz, int ncols,
                                                const float2 *__restrict__ tw, int tid, int nthreads) {
    int L = pl.N;
    for (int s = 0; s < pl.nstages; ++s) {
        const int R = pl.radix[s];
        fft_pass_dispatch<+1, true>(R, z, ldz, ncols, pl.N, L, tw, tid, nthreads);
        L /= R;
        __syncthreads();
    }
}

// r2c side: samples at digit-reversed positions -> natural-order spectrum (unnormalised, sign -).
__device__ __forceinline__ void fft_forward_dit(const FftPlan &pl, float2 *z, int ldz, int ncols,
                                                const float2 *__restrict__ tw, int tid, int nthreads) {
    int L = 1;
    for (int s = pl.nstages - 1; s >= 0; --s) {
        const int R = pl.radix[s];
        L *= R;
        fft_pass_dispatch<-1, false>(R, z, ldz, ncols, pl.N, L, tw, tid, nthreads);
        __syncthreads();
    }
}
