// fused.cuh -- one-CTA-per-element kernel for elements with 3D (phi-dependent) material whose whole strain
// spectrum fits in shared memory: SolidElement::computeStiff / FluidElement::computeStiff
// (SolidElement.cpp:43-65, 404-443; FluidElement.cpp:43-65, 333-355) with no HBM/L2 round trip between
// gather -> grad -> [rotate] -> c2r -> stress(+SLS) -> r2c -> [rotate^-1] -> quad -> scatter.
//
// Shared memory of one CTA (float2 units):
//   Z  [NPAIR * 25][ldz]   "Z-form" columns: two real strain/stress components of one GLL point as one complex
//                           column of length N = Nr (ldz = (N + 1) | 1 is odd, so lanes that run over columns are
//                           bank-conflict free, and slot N of every column is a spare);
//   U  [NC * 25][Mt]       gathered displacement of Mt Fourier modes (Mt = Nu + 1 when it fits); reused by the
//                           quad phase for the pointwise term r;
//   TW [stw_len]           per-stage twiddle tables of the plan.
// Phases (8 barriers for a two-stage plan with Mt = Nu + 1):
//   gather | grad (thread = (mode, point), writes Z-form) | DIF stages (thread = (column, butterfly), column fastest)
//   | stress (thread = (point, phi)) | DIT stages | quad-pre (in place: slot beta <- X, slot N - beta <- Y, U <- r)
//   | quad-post + scatter (RED.ADD.F32x2).
#pragma once
#include "kernels.cuh"

// ---------------------------------------------------------------- one FFT stage over all columns
// DIF (c2r, SIGN = +1): butterfly, then twiddle T[j][p].  DIT (r2c, SIGN = -1): conj twiddle, then butterfly.
template <int R, int SIGN, bool DIF>
__device__ __forceinline__ void fused_stage(float2 *__restrict__ z, int ldz, int ncols, int N, int L,
                                            const float2 *__restrict__ T, int tid, int nt) {
    const int Ls = L / R;
    const int nb = N / R;
    const int total = ncols * nb;
    int idx = tid;
    int b = idx / ncols, col = idx - b * ncols;
    const int db = nt / ncols, dc = nt - db * ncols;
    for (; idx < total; idx += nt) {
        int blk, j;
        if (Ls == 1) { blk = b; j = 0; }
        else if (L == N) { blk = 0; j = b; }
        else { blk = b / Ls; j = b - blk * Ls; }
        float2 *x = z + col * ldz + blk * L + j;
        float2 a[R];
#pragma unroll
        for (int q = 0; q < R; ++q) a[q] = x[q * Ls];
        if (!DIF && j != 0) {
            const float2 *t = T + j * R;
#pragma unroll
            for (int q = 1; q < R; ++q) {
                const float2 w = t[q];
                a[q] = make_float2(a[q].x * w.x + a[q].y * w.y, a[q].y * w.x - a[q].x * w.y);   // * conj(w)
            }
        }
        Dft<R, SIGN>::run(a);
        if (DIF && j != 0) {
            const float2 *t = T + j * R;
#pragma unroll
            for (int p = 1; p < R; ++p) a[p] = cmul(a[p], t[p]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) x[q * Ls] = a[q];
        col += dc;
        b += db;
        if (col >= ncols) { col -= ncols; ++b; }
    }
}

template <int SIGN, bool DIF>
__device__ __forceinline__ void fused_stage_dispatch(int R, float2 *z, int ldz, int ncols, int N, int L,
                                                     const float2 *T, int tid, int nt) {
    switch (R) {
        case 2: fused_stage<2, SIGN, DIF>(z, ldz, ncols, N, L, T, tid, nt); break;
        case 3: fused_stage<3, SIGN, DIF>(z, ldz, ncols, N, L, T, tid, nt); break;
        case 4: fused_stage<4, SIGN, DIF>(z, ldz, ncols, N, L, T, tid, nt); break;
        case 5: fused_stage<5, SIGN, DIF>(z, ldz, ncols, N, L, T, tid, nt); break;
        case 7: fused_stage<7, SIGN, DIF>(z, ldz, ncols, N, L, T, tid, nt); break;
        case 8: fused_stage<8, SIGN, DIF>(z, ldz, ncols, N, L, T, tid, nt); break;
        case 11: fused_stage<11, SIGN, DIF>(z, ldz, ncols, N, L, T, tid, nt); break;
        case 13: fused_stage<13, SIGN, DIF>(z, ldz, ncols, N, L, T, tid, nt); break;
        case 16: fused_stage<16, SIGN, DIF>(z, ldz, ncols, N, L, T, tid, nt); break;
        default: break;
    }
}

__host__ __device__ __forceinline__ int fused_ldz(int N) { return (N + 1) | 1; }

// ---------------------------------------------------------------- the kernel
// grid: one CTA per element elems[blockIdx.x]; block NT threads; 512 / NT CTAs per SM.
template <bool FLUID, int NT>
__global__ void __launch_bounds__(NT, 512 / NT)
    k_elem3d_fused(const ElemDesc *__restrict__ elems, const FftPlan *__restrict__ plans, const float2 *__restrict__ stwpool,
                   const float *__restrict__ geom, const float *__restrict__ coef, const float *__restrict__ attpar,
                   float *__restrict__ attstate, const float2 *__restrict__ displ, float2 *__restrict__ stiff) {
    constexpr int NC = FLUID ? 1 : 3, NPAIR = FLUID ? 2 : 3;
    constexpr int NHW = NT / 16;
    extern __shared__ float2 smem[];
    __shared__ ElemDesc sE;
    __shared__ FftPlan sP;
    const int tid = threadIdx.x;
    {
        const int *src = reinterpret_cast<const int *>(elems + blockIdx.x);
        int *dst = reinterpret_cast<int *>(&sE);
        for (int k = tid; k < (int)(sizeof(ElemDesc) / sizeof(int)); k += NT) dst[k] = src[k];
    }
    __syncthreads();
    {
        const int *src = reinterpret_cast<const int *>(plans + sE.plan_id);
        int *dst = reinterpret_cast<int *>(&sP);
        for (int k = tid; k < (int)(sizeof(FftPlan) / sizeof(int)); k += NT) dst[k] = src[k];
    }
    const ElemDesc &E = sE;
    const int N = E.nr, nu = E.nu, M = nu + 1, Mt = E.mt;
    const int ldz = fused_ldz(N);
    const bool nyq = E.nyq != 0;
    float2 *Z = smem;
    float2 *U = Z + NPAIR * AX_NPE * ldz;
    float2 *TW = U + NC * AX_NPE * Mt;
    const int hw = tid >> 4, t = tid & 15;
    __syncthreads();
    for (int k = tid; k < sP.stw_len; k += NT) TW[k] = stwpool[sP.stw_base + k];

    // ------------------------------------------------------------ gather + grad, Mt modes at a time
    for (int a0 = 0; a0 < M; a0 += Mt) {
        const int mt = min(Mt, M - a0);
        if (a0) __syncthreads();
        // Point::scatterDisplToElement (SolidPoint.cpp:175-195): half-warp per (component, point) row
        for (int row = hw; row < NC * AX_NPE; row += NHW) {
            const int c = row / AX_NPE, p = row - c * AX_NPE;
            const float2 *src = displ + (size_t)E.pt_off[p] + (size_t)c * E.pt_stride[p];
            const int nlive = E.pt_nlive[p];
            float2 *dst = U + row * Mt;
            for (int a = t; a < mt; a += 16) {
                const int al = a0 + a;
                float2 u = al < nlive ? __ldg(src + al) : czero();
                if (al == 0) u.y = 0.f;
                dst[a] = u;
            }
        }
        __syncthreads();
        for (int p = hw; p < AX_NPE; p += NHW) {
            const int i = p / 5, j = p - 5 * i;
            GCoef gc;
            load_gcoef(gc, E.axial, i, j);
            const PointGeom g = load_geom(geom, E.geom_off, p);
            const bool ax0 = E.axial && i == 0;
            float tr[4] = {0.f, 1.f, 0.f, 1.f};
            if (!FLUID && E.tiso) {
#pragma unroll
                for (int k = 0; k < 4; ++k) tr[k] = geom[E.trig_off + k * AX_NPE + p];
            }
            float2 *zp = Z + p * ldz;
            for (int a = t; a < mt; a += 16) {
                const int alpha = a0 + a;
                const bool dead = nyq && alpha == nu;
                if constexpr (!FLUID) {
                    float2 e[6];
                    grad6_point(U, Mt, a, i, j, gc, g, (float)alpha, ax0, e);
                    if (dead) {
#pragma unroll
                        for (int c = 0; c < 6; ++c) e[c] = czero();
                    }
                    if (E.tiso) rot_spz_to_rtz(e, tr[0], tr[1], tr[2], tr[3]);
#pragma unroll
                    for (int pr = 0; pr < 3; ++pr) zform_store(zp + pr * AX_NPE * ldz, N, alpha, e[2 * pr], e[2 * pr + 1]);
                } else {
                    float2 e[3];
                    grad_fluid_point(U, Mt, a, i, j, gc, g, (float)alpha, ax0, e);
                    if (dead) e[0] = e[1] = e[2] = czero();
                    zform_store(zp, N, alpha, e[0], e[1]);
                    zform_store(zp + AX_NPE * ldz, N, alpha, e[2], czero());
                }
            }
        }
    }
    __syncthreads();

    // ------------------------------------------------------------ c2r (SolverFFTW_N6::computeC2R, unnormalised, sign +)
    {
        int L = N;
        for (int s = 0; s < sP.nstages; ++s) {
            const int R = sP.radix[s];
            fused_stage_dispatch<+1, true>(R, Z, ldz, NPAIR * AX_NPE, N, L, TW + (sP.stw_off[s] - sP.stw_base), tid, NT);
            L /= R;
            __syncthreads();
        }
    }

    // ------------------------------------------------------------ physical space: stress (+ SLS attenuation)
    {
        const int total = AX_NPE * N;
        int idx = tid;
        int p = idx / N, pos = idx - p * N;
        const int dp = NT / N, dpos = NT - dp * N;
        const int cs = AX_NPE * ldz;
        for (; idx < total; idx += NT) {
            float2 *zc = Z + p * ldz + pos;
            if constexpr (!FLUID) {
                const float2 z0 = zc[0], z1 = zc[cs], z2 = zc[2 * cs];
                float e[6] = {z0.x, z0.y, z1.x, z1.y, z2.x, z2.y}, s[6];
                const float *cf = coef + E.coef_off + idx;   // [k][point][pos] with point * N + pos == idx
                stress_law<float>(E.law, e, s, [&](int k) { return __ldcs(cf + (size_t)k * total); });
                if (E.att_kind != ATT_NONE) {
                    const int P = E.att_kind == ATT_CG4 ? 4 : AX_NPE;
                    const int q = E.att_kind == ATT_CG4 ? cg4_index(p) : p;
                    if (q >= 0) {
                        const float *ap = attpar + E.att_par_off;
                        const float *mod = ap + 3 * E.nsls;
                        float *stt = attstate + E.att_state_off;
                        const size_t cell = (size_t)q * N + pos;
                        const size_t PN = (size_t)P * N, sl = 6 * PN;
                        attenuation_cell<float>(
                            E.nsls, ap, mod[cell], mod[PN + cell], mod[2 * PN + cell], E.do_kappa != 0, e, s,
                            [&](int k, int c) -> float & { return stt[k * sl + c * PN + cell]; },
                            [&](int c) -> float & { return stt[E.nsls * sl + c * PN + cell]; });
                    }
                }
                zc[0] = make_float2(s[0], s[1]);
                zc[cs] = make_float2(s[2], s[3]);
                zc[2 * cs] = make_float2(s[4], s[5]);
            } else {
                const float K = __ldcs(coef + E.coef_off + idx);   // Acoustic3D.cpp:9-16
                const float2 a = zc[0], b = zc[cs];
                zc[0] = cscale(a, K);
                zc[cs] = make_float2(b.x * K, 0.f);
            }
            p += dp;
            pos += dpos;
            if (pos >= N) { pos -= N; ++p; }
        }
    }
    __syncthreads();

    // ------------------------------------------------------------ r2c (computeR2C; the 1/Nr is applied at load below)
    {
        int L = 1;
        for (int s = sP.nstages - 1; s >= 0; --s) {
            const int R = sP.radix[s];
            L *= R;
            fused_stage_dispatch<-1, false>(R, Z, ldz, NPAIR * AX_NPE, N, L, TW + (sP.stw_off[s] - sP.stw_base), tid, NT);
            __syncthreads();
        }
    }

    // ------------------------------------------------------------ quad + scatter
    const float sc = 1.f / (float)N;   // SolverFFTW_N6::computeR2C scaling (SolverFFTW_N6.cpp:47-48)
    for (int a0 = 0; a0 < M; a0 += Mt) {
        const int mt = min(Mt, M - a0);
        if (a0) __syncthreads();
        // pointwise half, in place: slot beta <- X, slot N - beta <- Y (beta = 0: the spare slot N), U <- r
        for (int p = hw; p < AX_NPE; p += NHW) {
            const int i = p / 5;
            const PointGeom g = load_geom(geom, E.geom_off, p);
            const bool ax0 = E.axial && i == 0;
            float tr[4] = {0.f, 1.f, 0.f, 1.f};
            if (!FLUID && E.tiso) {
#pragma unroll
                for (int k = 0; k < 4; ++k) tr[k] = geom[E.trig_off + k * AX_NPE + p];
            }
            float2 *zp = Z + p * ldz;
            for (int a = t; a < mt; a += 16) {
                const int beta = a0 + a;
                if (nyq && beta == nu) continue;
                if constexpr (!FLUID) {
                    float2 s[6], X[3], Y[3], r[3];
#pragma unroll
                    for (int pr = 0; pr < 3; ++pr) zform_load(zp + pr * AX_NPE * ldz, N, beta, sc, s[2 * pr], s[2 * pr + 1]);
                    if (E.tiso) rot_rtz_to_spz(s, tr[0], tr[1], tr[2], tr[3]);
                    quad6_pre(s, g, (float)beta, ax0, X, Y, r);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        zp[c * AX_NPE * ldz + beta] = X[c];
                        zp[c * AX_NPE * ldz + N - beta] = Y[c];
                        U[(c * AX_NPE + p) * Mt + a] = r[c];
                    }
                } else {
                    float2 s[3], X, Y, r, dummy;
                    zform_load(zp, N, beta, sc, s[0], s[1]);
                    zform_load(zp + AX_NPE * ldz, N, beta, sc, s[2], dummy);
                    quad_fluid_pre(s, g, (float)beta, ax0, X, Y, r);
                    zp[beta] = X;
                    zp[N - beta] = Y;
                    U[p * Mt + a] = r;
                }
            }
        }
        __syncthreads();
        // tensor-product half + Point::gatherStiffFromElement (SolidPoint.cpp:197-209)
        for (int p = hw; p < AX_NPE; p += NHW) {
            const int i = p / 5, j = p - 5 * i;
            GCoef gc;
            load_gcoef(gc, E.axial, i, j);
            const int nlive = E.pt_nlive[p];
            const size_t base = (size_t)E.pt_off[p];
            const int st = E.pt_stride[p];
            for (int a = t; a < mt; a += 16) {
                const int beta = a0 + a;
                if ((nyq && beta == nu) || beta >= nlive) continue;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    float2 f = U[(c * AX_NPE + p) * Mt + a];
                    const float2 *zx = Z + (c * AX_NPE + j) * ldz + beta;           // X(k, j), k = 0..4
                    const float2 *zy = Z + (c * AX_NPE + i * 5) * ldz + N - beta;   // Y(i, k)
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        f = cfma(gc.gxi_row[k], zx[k * 5 * ldz], f);
                        f = cfma(gc.geta_row[k], zy[k * ldz], f);
                    }
                    if (beta == 0) f.y = 0.f;
                    scatter_sub(stiff, base + (size_t)c * st + beta, f);
                }
            }
        }
    }
}
