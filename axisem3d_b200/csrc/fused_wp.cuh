// fused_wp.cuh -- warp-per-GLL-point body of the persistent element kernel (k_elem3d_fused<..., WP = true>, fused.cuh).
//
// Same path as fused_element (SolidElement::computeStiff / FluidElement::computeStiff, SolidElement.cpp:43-65, 404-443;
// FluidElement.cpp:43-65, 333-355) and the same shared-memory tile (U mode-major, Z-form columns), but a different work
// decomposition.  The only couplings between the 25 GLL points of an element are the two 5x5 tensor-product contractions:
// grad reads the gathered displacement of all points, the second half of quad reads X/Y of all points.  Everything in
// between -- strain of one point for all modes, [rotation], c2r, stress (+SLS), r2c, [rotation^-1], pointwise half of quad --
// touches the 3 (fluid: 2) Z-form columns of ONE point only.  So warp p owns point p for the whole element:
//
//   CTA barrier A  (gather of this element has landed in U)
//     warp p: grad (lane = mode) -> Z columns of p | DIF stages (lane = butterfly) | stress (lane = phi) | DIT stages
//             | quad-pre (lane = mode; X, Y in place, r in registers)        ... only __syncwarp() between these phases
//   CTA barrier B  (X, Y of all points complete; U dead: the next element's gather starts here)
//     warp p: quad-post + scatter (lane = mode, RED.ADD.F32x2, 256 contiguous bytes per warp instruction)
//
// Two CTA barriers per element instead of one per FFT stage and phase, and the 25 warps drift apart between them, so the
// LDS-bound contractions, the FMA-bound butterflies and the L2-latency-bound stress of different points overlap on the SM
// instead of running in lock step.  All point indices are warp-uniform: gradient coefficients and geometry are uniform
// loads and the integer index arithmetic of the thread-per-(mode, point) mapping disappears.  25 compute warps (+ the
// Newmark warps) = 864 threads, 72 registers per thread: butterflies are radix 2..8 and odd primes (plans built with
// choose_radices_ct(N, 8)), twiddle tables are read p-major so that lanes (= j) read consecutive words.
#pragma once
#include "kernels.cuh"

#define AX_WP_NT (32 * AX_NPE)   // compute threads: one warp per GLL point
#ifndef AX_WP_QR
#define AX_WP_QR 4               // 32-mode rounds per quad super-chunk (the pointwise term r stays in registers across barrier B)
#endif

// One FFT stage over the NPAIR columns of one point, executed by one warp.  The columns of a point are contiguous
// (zp[pr * N + n]), so a stage is simply NPAIR * N / L blocks of length L: task idx -> block idx / Ls, offset idx % Ls.
// T2: p-major twiddles of this stage, T2[p * Ls + j] = exp(+2 pi i j p / L).
// NCT > 0: N = NCT and L = LCT are compile-time (strides and trip counts become immediates).
template <int R, int SIGN, bool DIF, int NPAIR, int NCT = 0, int LCT = 0>
__device__ __forceinline__ void wp_stage(float2 *__restrict__ zp, int N_rt, int L_rt, const float2 *__restrict__ T2, int lane) {
    const int N = NCT ? NCT : N_rt;
    const int L = NCT ? LCT : L_rt;
    const int Ls = L / R;
    const int total = NPAIR * (N / R);
    const float inv_ls = 1.0f / (float)Ls;
#pragma unroll(NCT ? 8 : 1)
    for (int idx = lane; idx < total; idx += 32) {
        int blk = idx, j = 0;
        if (Ls > 1) {
            if (NCT) blk = idx / Ls;                                  // constant divisor
            else blk = __float2int_rz(((float)idx + 0.5f) * inv_ls);   // exact: idx, Ls < 2^11
            j = idx - blk * Ls;
        }
        float2 *x = zp + blk * L + j;
        float2 a[R];
#pragma unroll
        for (int q = 0; q < R; ++q) a[q] = x[q * Ls];
        if (!DIF && Ls > 1) {
#pragma unroll
            for (int q = 1; q < R; ++q) a[q] = cmul_conj(a[q], T2[q * Ls + j]);
        }
        Dft<R, SIGN>::run(a);
        if (DIF && Ls > 1) {
#pragma unroll
            for (int p = 1; p < R; ++p) a[p] = cmul(a[p], T2[p * Ls + j]);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) x[q * Ls] = a[q];
    }
}

template <int SIGN, bool DIF, int NPAIR>
__device__ __forceinline__ void wp_stage_dispatch(int R, float2 *zp, int N, int L, const float2 *T2, int lane) {
    switch (R) {
        case 2: wp_stage<2, SIGN, DIF, NPAIR>(zp, N, L, T2, lane); break;
        case 3: wp_stage<3, SIGN, DIF, NPAIR>(zp, N, L, T2, lane); break;
        case 4: wp_stage<4, SIGN, DIF, NPAIR>(zp, N, L, T2, lane); break;
        case 5: wp_stage<5, SIGN, DIF, NPAIR>(zp, N, L, T2, lane); break;
        case 7: wp_stage<7, SIGN, DIF, NPAIR>(zp, N, L, T2, lane); break;
        case 8: wp_stage<8, SIGN, DIF, NPAIR>(zp, N, L, T2, lane); break;
        case 11: wp_stage<11, SIGN, DIF, NPAIR>(zp, N, L, T2, lane); break;
        case 13: wp_stage<13, SIGN, DIF, NPAIR>(zp, N, L, T2, lane); break;
        default: break;   // radix 16 is never planned for this kernel (choose_radices_ct(N, 8))
    }
}

// compile-time plan of the length-NCT transform (radices of choose_radices_ct(NCT, 8), like the host planner):
// stage S has block length L, its p-major twiddle table starts at T2 + TWOFF
template <int NPAIR, int NCT, int S, int L, int TWOFF>
struct CtWp {
    static __device__ __forceinline__ void inverse(float2 *zp, const float2 *T2, int lane) {
        if constexpr (S < choose_radices_ct(NCT, 8).n) {
            constexpr int R = choose_radices_ct(NCT, 8).r[S];
            constexpr int Ls = L / R;
            wp_stage<R, +1, true, NPAIR, NCT, L>(zp, NCT, L, T2 + TWOFF, lane);
            __syncwarp();
            CtWp<NPAIR, NCT, S + 1, Ls, TWOFF + (Ls > 1 ? L : 0)>::inverse(zp, T2, lane);
        }
    }
    static __device__ __forceinline__ void forward(float2 *zp, const float2 *T2, int lane) {
        if constexpr (S < choose_radices_ct(NCT, 8).n) {
            constexpr int R = choose_radices_ct(NCT, 8).r[S];
            constexpr int Ls = L / R;
            CtWp<NPAIR, NCT, S + 1, Ls, TWOFF + (Ls > 1 ? L : 0)>::forward(zp, T2, lane);
            wp_stage<R, -1, false, NPAIR, NCT, L>(zp, NCT, L, T2 + TWOFF, lane);
            __syncwarp();
        }
    }
};

// geometry of point p = tid / 32 of element E, one value per lane: lanes 0..4 dsdxii, dsdeta, dzdxii, dzdeta, inv_s; lanes 5..8
// sin t, cos t, sin 2t, cos 2t (TI / anisotropic elements).  Requested one element ahead (behind barrier B of the previous
// element) and broadcast by shuffles where needed: one live register per thread, no exposed load latency.
__device__ __forceinline__ float wp_load_geom(const ElemDesc &E, const float *__restrict__ geom, int tid, bool fluid) {
    const int p = tid >> 5, lane = tid & 31;
    float gv = (lane == 6 || lane == 8) ? 1.f : 0.f;
    if (lane < 5) gv = geom[E.geom_off + lane * AX_NPE + p];
    else if (lane < 9 && !fluid && E.tiso != 0) gv = geom[E.trig_off + (lane - 5) * AX_NPE + p];
    return gv;
}

// stress of one point for all phi, no attenuation: the moduli of NB rounds are requested before the first use
// (Isotropic3D.cpp:10-27, TransverselyIsotropic3D.cpp:10-28, Anisotropic3D.cpp:10-54).  cf: first modulus of this point.
template <int NCOEF, int NB>
__device__ __forceinline__ void wp_stress(int law, float2 *__restrict__ zp, const float *__restrict__ cf, int N, int cf_stride, int lane) {
    for (int pos0 = lane; pos0 < N; pos0 += 32 * NB) {
        float c[NB][NCOEF];
#pragma unroll
        for (int u = 0; u < NB; ++u)
            if (pos0 + 32 * u < N) {
#pragma unroll
                for (int k = 0; k < NCOEF; ++k) c[u][k] = __ldcs(cf + (size_t)k * cf_stride + pos0 + 32 * u);
            }
#pragma unroll
        for (int u = 0; u < NB; ++u)
            if (pos0 + 32 * u < N) {
                float2 *zc = zp + pos0 + 32 * u;
                const float2 z0 = zc[0], z1 = zc[N], z2 = zc[2 * N];
                float ee[6] = {z0.x, z0.y, z1.x, z1.y, z2.x, z2.y}, s[6];
                stress_law<float>(law, ee, s, [&](int k) { return c[u][k]; });
                zc[0] = make_float2(s[0], s[1]);
                zc[N] = make_float2(s[2], s[3]);
                zc[2 * N] = make_float2(s[4], s[5]);
            }
    }
}

// E, P: descriptor and plan of this element (shared memory).  On entry the first gather tile of this element is in
// flight (cp.async); `after_first_sync` runs behind barrier A (every thread has left the previous element),
// `after_u_dead` behind the first barrier B (it starts the next element's gather into U).
//
// Z tile of this body: column (point p, pair pr) at Z[(p * NPAIR + pr) * N], no padding.  After quad-pre, slot beta of
// column (p, c) holds X_c(beta) and slot N - beta holds Y_c(beta); mode 0 (X, Y real) is packed as (X_c(0), Y_c(0)) in slot 0.
// gv: wp_load_geom of this element.  NCT > 0: body specialised for Nr == NCT.
template <bool FLUID, int NT, int NWW, int NCT, typename GatherFn, typename AfterSyncFn, typename AfterUFn>
__device__ __forceinline__ void wp_element(const FusedCtx<FLUID> &cx, const ElemDesc &E, const FftPlan &P, int tid, const float gv,
                                           GatherFn gather, AfterSyncFn after_first_sync, AfterUFn after_u_dead) {
    static_assert(NT == AX_WP_NT, "one warp per GLL point");
    constexpr int NC = FLUID ? 1 : 3, NPAIR = FLUID ? 2 : 3;
    constexpr int US = NC * AX_NPE;
    constexpr int NHW = NT / 16;
    constexpr int QR = AX_WP_QR;
    float2 *const U = cx.U, *const TW = cx.TW, *const Z = cx.Z;
    const int p = tid >> 5, lane = tid & 31;       // warp = GLL point
    const int i = p / 5, j = p - 5 * i;
    const int N = NCT ? NCT : E.nr, nu = N / 2, M = nu + 1, Mt = E.mt;
    const bool nyq = (N & 1) == 0;
    const bool axial = E.axial != 0, tiso = !FLUID && E.tiso != 0;
    const bool ax0 = axial && i == 0;
    float2 *const zp = Z + p * NPAIR * N;

    auto geom_of = [&]() {
        PointGeom g;
        g.dsdxii = __shfl_sync(0xffffffffu, gv, 0);
        g.dsdeta = __shfl_sync(0xffffffffu, gv, 1);
        g.dzdxii = __shfl_sync(0xffffffffu, gv, 2);
        g.dzdeta = __shfl_sync(0xffffffffu, gv, 3);
        g.inv_s = __shfl_sync(0xffffffffu, gv, 4);
        return g;
    };

    // ------------------------------------------------------------ gather (prefetched) + grad, Mt modes at a time   @phase wp gather wait + grad
    for (int a0 = 0; a0 < M; a0 += Mt) {
        const int mt = min(Mt, M - a0);
        if (a0) {
            cta_sync<NT, NWW>();
            gather(E, a0, mt, -1);
        }
        cp_async_wait_all();
        if (a0 == 0 && (tid & 15) == 0)   // Im(u) of mode 0 is not used (Gradient.cpp:209-224): this thread copied these entries
            for (int row = tid >> 4; row < US; row += NHW) U[row].y = 0.f;
        cta_sync<NT, NWW>();              // barrier A
        if (a0 == 0) after_first_sync();
        GCoef gc;
        load_gcoef(gc, axial, i, j);
        const PointGeom g = geom_of();
        float tr[4] = {0.f, 1.f, 0.f, 1.f};
        if (tiso) {
#pragma unroll
            for (int k = 0; k < 4; ++k) tr[k] = __shfl_sync(0xffffffffu, gv, 5 + k);
        }
        for (int a = lane; a < mt; a += 32) {
            const int alpha = a0 + a;
            const bool dead = nyq && alpha == nu;
            if constexpr (!FLUID) {
                float2 ee[6];
                grad6_mm(U + a * US, i, j, gc, g, (float)alpha, ax0, ee);
                if (dead) {
#pragma unroll
                    for (int c = 0; c < 6; ++c) ee[c] = czero();
                }
                if (tiso) rot_spz_to_rtz(ee, tr[0], tr[1], tr[2], tr[3]);
#pragma unroll
                for (int pr = 0; pr < 3; ++pr) zform_store(zp + pr * N, N, alpha, ee[2 * pr], ee[2 * pr + 1]);
            } else {
                float2 ee[3];
                grad_fluid_mm(U + a * US, i, j, gc, g, (float)alpha, ax0, ee);
                if (dead) ee[0] = ee[1] = ee[2] = czero();
                zform_store(zp, N, alpha, ee[0], ee[1]);
                zform_store(zp + N, N, alpha, ee[2], czero());
            }
        }
    }
    __syncwarp();

    // ------------------------------------------------------------ c2r (SolverFFTW_N6::computeC2R, unnormalised, sign +)   @phase wp c2r
    const float2 *const T2 = TW - P.stw_base;   // TW holds the p-major tables [stw_base + stw2_delta ...) of this plan
    if constexpr (NCT != 0) {
        CtWp<NPAIR, NCT, 0, NCT, 0>::inverse(zp, TW, lane);
    } else {
        int L = N;
        for (int s = 0; s < P.nstages; ++s) {
            const int R = P.radix[s];
            wp_stage_dispatch<+1, true, NPAIR>(R, zp, N, L, T2 + P.stw_off[s], lane);
            L /= R;
            __syncwarp();
        }
    }

    // ------------------------------------------------------------ physical space: stress (+ SLS attenuation)   @phase wp stress
    {
        bool done = false;
        if constexpr (!FLUID) {
            if (E.att_kind == ATT_NONE) {
                const float *cf = cx.coef + E.coef_off + (size_t)p * N;
                const int law = E.law;
                if (law == LAW_ISO) wp_stress<2, 4>(law, zp, cf, N, AX_NPE * N, lane);
                else if (law == LAW_TI) wp_stress<5, 2>(law, zp, cf, N, AX_NPE * N, lane);
                else wp_stress<21, 1>(law, zp, cf, N, AX_NPE * N, lane);
                done = true;
            }
        }
        if (!done) physical_space<FLUID, 32>(E, cx.coef, cx.attpar, cx.attstate, zp, N, N, p, 1, lane, N);
    }
    __syncwarp();

    // ------------------------------------------------------------ r2c (computeR2C; the 1/Nr is applied at load below)   @phase wp r2c
    if constexpr (NCT != 0) {
        CtWp<NPAIR, NCT, 0, NCT, 0>::forward(zp, TW, lane);
    } else {
        int L = 1;
        for (int s = P.nstages - 1; s >= 0; --s) {
            const int R = P.radix[s];
            L *= R;
            wp_stage_dispatch<-1, false, NPAIR>(R, zp, N, L, T2 + P.stw_off[s], lane);
            __syncwarp();
        }
    }

    // ------------------------------------------------------------ quad + scatter, 32 * QR modes at a time   @phase wp quad-pre
    const float sc = 1.f / (float)N;   // SolverFFTW_N6::computeR2C scaling (SolverFFTW_N6.cpp:47-48)
    for (int a0 = 0; a0 < M; a0 += 32 * QR) {
        float2 r[QR][NC];
        {
            // pointwise half, in place: slot beta <- X, slot N - beta <- Y (beta = 0: both real, packed in slot 0); r stays in registers
            const PointGeom g = geom_of();
            float tr[4] = {0.f, 1.f, 0.f, 1.f};
            if (tiso) {
#pragma unroll
                for (int k = 0; k < 4; ++k) tr[k] = __shfl_sync(0xffffffffu, gv, 5 + k);
            }
#pragma unroll
            for (int q = 0; q < QR; ++q) {
                const int beta = a0 + 32 * q + lane;
                if (beta < M && !(nyq && beta == nu)) {
                    if constexpr (!FLUID) {
                        float2 s[6], X[3], Y[3];
#pragma unroll
                        for (int pr = 0; pr < 3; ++pr) zform_load(zp + pr * N, N, beta, sc, s[2 * pr], s[2 * pr + 1]);
                        if (tiso) rot_rtz_to_spz(s, tr[0], tr[1], tr[2], tr[3]);
                        quad6_pre(s, g, (float)beta, ax0, X, Y, r[q]);
                        if (beta == 0) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) zp[c * N] = make_float2(X[c].x, Y[c].x);
                        } else {
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                zp[c * N + beta] = X[c];
                                zp[c * N + N - beta] = Y[c];
                            }
                        }
                    } else {
                        float2 s[3], X, Y, dummy;
                        zform_load(zp, N, beta, sc, s[0], s[1]);
                        zform_load(zp + N, N, beta, sc, s[2], dummy);
                        quad_fluid_pre(s, g, (float)beta, ax0, X, Y, r[q][0]);
                        if (beta == 0) zp[0] = make_float2(X.x, Y.x);
                        else { zp[beta] = X; zp[N - beta] = Y; }
                    }
                }
            }
        }
        cta_sync<NT, NWW>();   // barrier B
        if (a0 == 0) after_u_dead();   // @phase wp next-element gather issue
        // tensor-product half + Point::gatherStiffFromElement (SolidPoint.cpp:197-209)   @phase wp quad-post + scatter
        {
            GCoef gc;
            load_gcoef(gc, axial, i, j);
            const int nlive = E.pt_nlive[p];
            float2 *const dst = cx.stiff + (size_t)E.pt_off[p];
            const int st = E.pt_stride[p];
            const int cN = NPAIR * N;                       // distance between the columns of two points
            const float2 *const zx0 = Z + j * cN;           // X_c(k, j) at zx0 + 5 k cN + c N + beta
            const float2 *const zy0 = Z + i * 5 * cN;       // Y_c(i, k) at zy0 + k cN + c N + N - beta
#pragma unroll
            for (int q = 0; q < QR; ++q) {
                const int beta = a0 + 32 * q + lane;
                if (beta != 0 && beta < M && !(nyq && beta == nu) && beta < nlive) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        float2 f = r[q][c];
                        const float2 *zx = zx0 + c * N + beta;
                        const float2 *zy = zy0 + c * N + N - beta;
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            f = cfma(gc.gxi_row[k], zx[k * 5 * cN], f);
                            f = cfma(gc.geta_row[k], zy[k * cN], f);
                        }
                        atomicAdd(dst + (size_t)c * st + beta, make_float2(-f.x, -f.y));   // stiff -= f (RED.ADD.F32x2)
                    }
                }
            }
            if (a0 == 0 && lane == 0 && nlive > 0) {   // mode 0 (one lane per element and point): real, X and Y packed in slot 0
#pragma unroll 1
                for (int c = 0; c < NC; ++c) {
                    float f = r[0][c].x;
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        f = fmaf(gc.gxi_row[k], zx0[k * 5 * cN + c * N].x, f);
                        f = fmaf(gc.geta_row[k], zy0[k * cN + c * N].y, f);
                    }
                    atomicAdd(dst + (size_t)c * st, make_float2(-f, 0.f));
                }
            }
        }
        // no barrier before the next super-chunk: its quad-pre rewrites slots [a0', a0' + 32 QR) and their mirrors of the
        // warp's own columns only, which no quad-post of this super-chunk reads
    }
    // barrier A of the next element separates these reads of Z from its grad
}
