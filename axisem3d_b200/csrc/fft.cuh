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
#include <algorithm>
#include <functional>
#include <vector>

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
};

// Radix sequence of the DIF transform of length N (host planner; every phi-dependent array is uploaded in the digit-reversed
// order of this plan, so all kernels of a process share it).  Radices: the primes up to 13, 4, 8, 16 and the composites
// 6, 9, 10, 12 (DftCT; 14 and 15 would save another 1.5 % of the stages of cfg4 for the two most register-hungry butterflies) -- chosen to minimise the number of stages: each stage is one trip of the whole spectrum
// through shared memory, one barrier and (all but the last) one twiddle multiplication per sample.  Among the factorisations
// with the fewest stages the most even one wins (smallest sum of squares); the stages are ordered largest radix first, except
// that an odd radix, if there is one, goes last (stride-1 butterflies: an odd radix keeps lanes that walk over butterflies of
// one column on different banks).  n < 0 marks an unsupported N (a prime factor above 13).
struct RadixList {
    int n;
    int r[AX_MAX_STAGES];
};
inline RadixList choose_radices_plan(int N) {
    static const int RAD[13] = {16, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2};
    RadixList out{0, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};
    if (N <= 1) return out;
    // best[d] for every divisor d of N, by increasing d: (stages, sum of squares, first radix)
    struct Best { int stages; long long sq; int first; };
    std::vector<int> divs;
    for (int d = 1; d <= N; ++d)
        if (N % d == 0) divs.push_back(d);
    std::vector<Best> best(divs.size(), Best{1 << 20, 0, 0});
    auto at = [&](int d) { return (size_t)(std::lower_bound(divs.begin(), divs.end(), d) - divs.begin()); };
    best[0] = Best{0, 0, 0};
    for (size_t i = 1; i < divs.size(); ++i) {
        const int d = divs[i];
        for (int r : RAD) {
            if (d % r) continue;
            const Best &sub = best[at(d / r)];
            if (sub.stages >= (1 << 20)) continue;
            const Best cand{sub.stages + 1, sub.sq + (long long)r * r, r};
            if (cand.stages < best[i].stages || (cand.stages == best[i].stages && cand.sq < best[i].sq)) best[i] = cand;
        }
    }
    if (best.back().stages >= (1 << 20) || best.back().stages > AX_MAX_STAGES) { out.n = -1; return out; }
    std::vector<int> rs;
    for (int d = N; d > 1; d /= best[at(d)].first) rs.push_back(best[at(d)].first);
    std::sort(rs.begin(), rs.end(), std::greater<int>());
    int odd = -1;
    for (int k = (int)rs.size() - 1; k >= 0; --k)
        if (rs[k] % 2) { odd = k; break; }          // the smallest odd radix
    if (odd >= 0) { const int r = rs[odd]; rs.erase(rs.begin() + odd); rs.push_back(r); }
    out.n = (int)rs.size();
    for (int k = 0; k < out.n; ++k) out.r[k] = rs[k];
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
// composite radix R = R1 * R2 in registers (Cooley-Tukey, every index a compile-time constant): R2 transforms of length R1 over
// a[R2 n1 + n2], the twiddles w_R^(n2 k1) from the constant tables, then R1 transforms of length R2 -> a[k1 + R1 k2].
// A stage of radix 12 or 10 replaces two stages (4 x 3, 2 x 5): one trip through shared memory, one barrier and one twiddle
// table less per transform (the planner below minimises the number of stages).
template <int R1, int R2, int SIGN>
struct DftCT {
    static __device__ __forceinline__ void run(float2 (&a)[R1 * R2]) {
        constexpr int R = R1 * R2;
        float2 t[R2][R1];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) {
            float2 c[R1];
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1) c[n1] = a[R2 * n1 + n2];
            Dft<R1, SIGN>::run(c);
#pragma unroll
            for (int k1 = 0; k1 < R1; ++k1) {
                const int m = (n2 * k1) % R;
                if (m == 0) t[n2][k1] = c[k1];
                else if (4 * m == R) t[n2][k1] = cmul_i<SIGN>(c[k1]);
                else if (2 * m == R) t[n2][k1] = make_float2(-c[k1].x, -c[k1].y);
                else if (4 * m == 3 * R) t[n2][k1] = cmul_i<-SIGN>(c[k1]);
                else t[n2][k1] = cmul(c[k1], make_float2(c_cos[R][m], SIGN * c_sin[R][m]));
            }
        }
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            float2 c[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) c[n2] = t[n2][k1];
            Dft<R2, SIGN>::run(c);
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) a[k1 + R1 * k2] = c[k2];
        }
    }
};
template <int SIGN> struct Dft<3, SIGN> : DftOdd<3, SIGN> {};
template <int SIGN> struct Dft<5, SIGN> : DftOdd<5, SIGN> {};
template <int SIGN> struct Dft<7, SIGN> : DftOdd<7, SIGN> {};
template <int SIGN> struct Dft<11, SIGN> : DftOdd<11, SIGN> {};
template <int SIGN> struct Dft<13, SIGN> : DftOdd<13, SIGN> {};
template <int SIGN> struct Dft<6, SIGN> : DftCT<2, 3, SIGN> {};
template <int SIGN> struct Dft<9, SIGN> : DftCT<3, 3, SIGN> {};
template <int SIGN> struct Dft<10, SIGN> : DftCT<2, 5, SIGN> {};
template <int SIGN> struct Dft<12, SIGN> : DftCT<4, 3, SIGN> {};

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
        case 6: fft_pass<6, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
        case 7: fft_pass<7, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
        case 8: fft_pass<8, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
        case 9: fft_pass<9, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
        case 10: fft_pass<10, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
        case 12: fft_pass<12, SIGN, DIF>(z, ldz, ncols, N, L, tw, tid, nthreads); break;
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
