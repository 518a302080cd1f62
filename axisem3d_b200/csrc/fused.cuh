// fused.cuh -- persistent one-CTA-per-element kernel for elements with 3D (phi-dependent) material:
// SolidElement::computeStiff / FluidElement::computeStiff (SolidElement.cpp:43-65, 404-443; FluidElement.cpp:43-65, 333-355)
// with no HBM/L2 round trip between gather -> grad -> [rotate] -> c2r -> stress(+SLS) -> r2c -> [rotate^-1] -> quad -> scatter.
//
// An element is processed in `ng` passes over groups of GLL rows (xi index): ng = 1 (all 25 points at once) while the whole
// strain spectrum fits in shared memory, 2 (rows 0-2 | 3-4), 3 (0-1 | 2-3 | 4) or 5 (one row per pass) for longer azimuthal
// expansions -- a pass holds the Z-form columns of its np = 5 * rows points only, so that one GLL row of Nr <= ~1300 fits.
// What couples the rows is cheap to repeat: the xi-derivative of the gradient needs the displacement of all 25 points, so
// every pass gathers it again (L2 hits), and the xi-part of the quadrature of a pass is a partial sum for all 25 points,
// scattered with one more RED per point (f = G_xi X + Y G_eta^T + r is linear in the rows of X).
//
// Shared memory of one CTA (float2 units), tile region [0, R) laid out per element (host: plan_fused_element):
//   U  [0, twoff)          gathered displacement of a tile of `mt` modes, mode-major: U[a * NC*25 + c * 25 + point] (all
//                           contraction operands of one thread sit at immediate offsets; the odd row stride keeps lanes =
//                           modes conflict free);
//   TW [twoff, zoff)       per-stage twiddle tables of the plan;
//   Z  [zoff, R)           "Z-form" columns of the pass, [pair][local point][ldz]: two real strain/stress components of one
//                           GLL point as one complex column of length N = Nr.  ldz = (N + 1) | 1 is odd, so lanes that run
//                           over columns are bank-conflict free, and slot N of every column is a spare.
// Behind the tile region: the stage buffers of the in-kernel Newmark warps.
// One CTA takes elements off a device-side work counter (elements are sorted by Nr, largest first, so the queue is
// LPT-like).  While a pass is in its FFT/stress/quad phases the first displacement tile of the NEXT pass streams into U
// with cp.async (U is dead after grad: the quad phase keeps its pointwise term in registers), the next element's
// descriptor and plan are prefetched, and the moduli of an element are prefetched into L2 at its top.  Phases of one pass:
//   grad (thread = (mode, point of the group), writes Z-form) | DIF stages (thread = (column, butterfly), column fastest)
//   | stress (thread = (point, phi)) | DIT stages | quad-pre (in place: slot beta <- X, slot N - beta <- Y)
//   | quad-post + scatter (RED.ADD.F32x2; the points outside the group receive the xi-partial sums only).
#pragma once
#include "kernels.cuh"

__host__ __device__ constexpr int fused_ldz(int N) { return (N + 1) | 1; }

__device__ __forceinline__ void cp_async8(float2 *dst_smem, const float2 *src, bool pred) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int sz = pred ? 8 : 0;   // src-size 0: nothing is read, the 8 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }


// ---------------------------------------------------------------- compute-warp barrier
// With in-kernel Newmark warps (NWW > 0) the CTA has NT compute threads + 32 * NWW Newmark threads; the compute warps
// synchronise among themselves on named barrier 1 and the Newmark warps never join a CTA-wide barrier.
template <int NT, int NWW>
__device__ __forceinline__ void cta_sync() {
    if constexpr (NWW > 0) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
    else __syncthreads();
}

// ---------------------------------------------------------------- one FFT stage over all columns
// DIF (c2r, SIGN = +1): butterfly, then twiddle T[j][p].  DIT (r2c, SIGN = -1): conj twiddle, then butterfly.
#ifndef AX_STAGE_INLINE
#define AX_STAGE_INLINE 0   // 0: every radix stage is a real function with its own register allocation (B200: cfg4 0.816 vs 0.823 ms per step inlined)
#endif
#if AX_STAGE_INLINE
#define AX_STAGE_ATTR __forceinline__
#else
#define AX_STAGE_ATTR __noinline__
#endif
template <int R, int SIGN, bool DIF, int NT>
__device__ AX_STAGE_ATTR void fused_stage(float2 *__restrict__ z, int N, int L, int ncols, const float2 *__restrict__ T, int tid) {
    const int ldz = fused_ldz(N);
    const int Ls = L / R;
    const int nb = N / R;
    const int total = ncols * nb;
    int idx = tid;
    int b = idx / ncols, col = idx - b * ncols;
    const int db = NT / ncols, dc = NT - db * ncols;
    for (; idx < total; idx += NT) {
        int blk, j;
        if (Ls == 1) { blk = b; j = 0; }
        else if (L == N) { blk = 0; j = b; }
        else { blk = b / Ls; j = b - blk * Ls; }
        float2 *x = z + col * ldz + blk * L + j;
        float2 a[R];
#pragma unroll
        for (int q = 0; q < R; ++q) a[q] = x[q * Ls];
        if (!DIF && Ls > 1 && j != 0) {
            const float2 *t = T + j * R;
#pragma unroll
            for (int q = 1; q < R; ++q) {
                const float2 w = t[q];
                a[q] = cmul_conj(a[q], w);
            }
        }
        Dft<R, SIGN>::run(a);
        if (DIF && Ls > 1 && j != 0) {
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

template <int SIGN, bool DIF, int NT>
__device__ __forceinline__ void fused_stage_dispatch(int R, float2 *z, int N, int L, int ncols, const float2 *T, int tid) {
    switch (R) {
        case 2: fused_stage<2, SIGN, DIF, NT>(z, N, L, ncols, T, tid); break;
        case 3: fused_stage<3, SIGN, DIF, NT>(z, N, L, ncols, T, tid); break;
        case 4: fused_stage<4, SIGN, DIF, NT>(z, N, L, ncols, T, tid); break;
        case 5: fused_stage<5, SIGN, DIF, NT>(z, N, L, ncols, T, tid); break;
        case 6: fused_stage<6, SIGN, DIF, NT>(z, N, L, ncols, T, tid); break;
        case 7: fused_stage<7, SIGN, DIF, NT>(z, N, L, ncols, T, tid); break;
        case 8: fused_stage<8, SIGN, DIF, NT>(z, N, L, ncols, T, tid); break;
        case 9: fused_stage<9, SIGN, DIF, NT>(z, N, L, ncols, T, tid); break;
        case 10: fused_stage<10, SIGN, DIF, NT>(z, N, L, ncols, T, tid); break;
        case 12: fused_stage<12, SIGN, DIF, NT>(z, N, L, ncols, T, tid); break;
        case 11: fused_stage<11, SIGN, DIF, NT>(z, N, L, ncols, T, tid); break;
        case 13: fused_stage<13, SIGN, DIF, NT>(z, N, L, ncols, T, tid); break;
        case 16: fused_stage<16, SIGN, DIF, NT>(z, N, L, ncols, T, tid); break;
        default: break;
    }
}

// ---------------------------------------------------------------- grad / quad on the mode-major tile
// um = U + a * NC*25: the displacement of one mode, [c * 25 + point].
__device__ __forceinline__ void grad6_mm(const float2 *__restrict__ um, int i, int j, const GCoef &gc, const PointGeom &g,
                                         float alpha, bool axial_row0, float2 (&e)[6]) {
    const float2 *col = um + j, *row = um + i * 5;
    float2 GU[3], UG[3], u[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float2 a = czero(), b = czero();
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            a = cfma(gc.gxi_col[k], col[c * AX_NPE + k * 5], a);
            b = cfma(gc.geta_col[k], row[c * AX_NPE + k], b);
        }
        GU[c] = a;
        UG[c] = b;
        u[c] = row[c * AX_NPE + j];
    }
    float2 ds[3], dz[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        ds[c] = cfma(g.dzdeta, GU[c], cscale(UG[c], g.dzdxii));
        dz[c] = cfma(g.dsdeta, GU[c], cscale(UG[c], g.dsdxii));
    }
    const float2 v0 = cadd(u[0], mul_ialpha(u[1], alpha));
    const float2 v1 = csub(mul_ialpha(u[0], alpha), u[1]);
    const float2 v2 = mul_ialpha(u[2], alpha);
    e[0] = ds[0];
    e[1] = cscale(v0, g.inv_s);
    e[2] = dz[2];
    e[3] = cfma(g.inv_s, v2, dz[1]);
    e[4] = cadd(dz[0], ds[2]);
    e[5] = cfma(g.inv_s, v1, ds[1]);
    if (axial_row0) {   // L'Hopital rows on the axis (Gradient.cpp:221-224, 245-254)
        const float2 gv0 = cadd(GU[0], mul_ialpha(GU[1], alpha));
        const float2 gv1 = csub(mul_ialpha(GU[0], alpha), GU[1]);
        const float2 gv2 = mul_ialpha(GU[2], alpha);
        e[1] = cfma(g.dzdeta, gv0, e[1]);
        e[5] = cfma(g.dzdeta, gv1, e[5]);
        e[3] = cfma(g.dzdeta, gv2, e[3]);
        if (alpha == 1.f) {
            const float2 uv0 = cadd(UG[0], mul_ialpha(UG[1], alpha));
            const float2 uv1 = csub(mul_ialpha(UG[0], alpha), UG[1]);
            e[1] = cfma(g.dzdxii, uv0, e[1]);
            e[5] = cfma(g.dzdxii, uv1, e[5]);
        }
    }
}
__device__ __forceinline__ void grad_fluid_mm(const float2 *__restrict__ um, int i, int j, const GCoef &gc, const PointGeom &g,
                                              float alpha, bool axial_row0, float2 (&e)[3]) {
    const float2 *col = um + j, *row = um + i * 5;
    float2 GU = czero(), UG = czero();
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        GU = cfma(gc.gxi_col[k], col[k * 5], GU);
        UG = cfma(gc.geta_col[k], row[k], UG);
    }
    const float2 v = mul_ialpha(row[j], alpha);
    e[0] = cfma(g.dzdeta, GU, cscale(UG, g.dzdxii));
    e[1] = cscale(v, g.inv_s);
    e[2] = cfma(g.dsdeta, GU, cscale(UG, g.dsdxii));
    if (axial_row0) e[1] = cfma(g.dzdeta, mul_ialpha(GU, alpha), e[1]);
}

// ---------------------------------------------------------------- one element   @phase prologue
template <bool FLUID>
struct FusedCtx {
    const float *__restrict__ geom;
    const float *__restrict__ coef;
    const float *__restrict__ attpar;
    float *__restrict__ attstate;
    const float2 *__restrict__ displ;
    float2 *__restrict__ stiff;
    float2 *U, *TW, *Z;
    const float *sgeom;   // shared-memory copy of this element's geometry [5][25] + trig [4][25] (staged with the gather)
};

// stress of NB (point, phi) cells at once: all moduli are requested before the first use (Isotropic3D.cpp:10-27,
// TransverselyIsotropic3D.cpp:10-28, Anisotropic3D.cpp:10-54; no attenuation on this path)
// Z holds the columns of `np` consecutive points of one element: pair k of local point pl at Z + (k * np + pl) * ldz.
// cf points at the first of those points in the element's moduli ([k][25][N]: cf_stride = 25 * N between moduli).
#ifndef AX_STRESS_PIPE
#define AX_STRESS_PIPE 0   // measured on B200 (profiles/microbench/ab.sh): no gain alone, +2 % step time together with AX_SGEOM (registers)
#endif
#ifndef AX_SGEOM
#define AX_SGEOM 1     // element geometry staged in shared memory with the gather (0: read from global memory where used)
#endif
template <int NCOEF, int NB, int NT>
__device__ __forceinline__ void stress_batch(int law, float2 *__restrict__ Z, const float *__restrict__ cf, int total, int cf_stride, int cs,
                                             int ldz, int N, int tid) {
    constexpr bool PIPE = AX_STRESS_PIPE && NCOEF * NB <= 10;   // moduli of the next batch are requested before this batch is computed
    const int dp = NT / N, dpos = NT - dp * N;
    int idx = tid;
    int p = idx / N, pos = idx - p * N;
    float c[NB][NCOEF], cn[PIPE ? NB : 1][PIPE ? NCOEF : 1];
    auto load = [&](float (*dst)[PIPE ? NCOEF : 1], int at) {
#pragma unroll
        for (int u = 0; u < NB; ++u)
            if (at + u * NT < total) {
#pragma unroll
                for (int k = 0; k < NCOEF; ++k) dst[u][k] = __ldcs(cf + (size_t)k * cf_stride + at + u * NT);
            }
    };
    if constexpr (PIPE) load(c, idx);
    for (; idx < total; idx += NB * NT) {
        if constexpr (PIPE) {
            load(cn, idx + NB * NT);
        } else {
#pragma unroll
            for (int u = 0; u < NB; ++u)
                if (idx + u * NT < total) {
#pragma unroll
                    for (int k = 0; k < NCOEF; ++k) c[u][k] = __ldcs(cf + (size_t)k * cf_stride + idx + u * NT);
                }
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            if (idx + u * NT < total) {
                float2 *zc = Z + p * ldz + pos;
                const float2 z0 = zc[0], z1 = zc[cs], z2 = zc[2 * cs];
                float ee[6] = {z0.x, z0.y, z1.x, z1.y, z2.x, z2.y}, s[6];
                stress_law<float>(law, ee, s, [&](int k) { return c[u][k]; });
                zc[0] = make_float2(s[0], s[1]);
                zc[cs] = make_float2(s[2], s[3]);
                zc[2 * cs] = make_float2(s[4], s[5]);
            }
            p += dp;
            pos += dpos;
            if (pos >= N) { pos -= N; ++p; }
        }
        if constexpr (PIPE) {
#pragma unroll
            for (int u = 0; u < NB; ++u)
#pragma unroll
                for (int k = 0; k < NCOEF; ++k) c[u][k] = cn[u][k];
        }
    }
}

// Physical space of `np` points [p0, p0 + np) of element E: stress (+ SLS attenuation) on the c2r output, in place
// (Isotropic3D.cpp:10-27, TransverselyIsotropic3D.cpp:10-28, Anisotropic3D.cpp:10-54, Attenuation3D_{Full,CG4}.cpp,
// Acoustic3D.cpp:9-16).  Phi positions are in the digit-reversed order of the plan, like the uploaded moduli.
template <bool FLUID, int NT, bool PRT = false>
__device__ __forceinline__ void physical_space(const ElemDesc &E, const float *__restrict__ coef, const float *__restrict__ attpar,
                                               float *__restrict__ attstate, float2 *__restrict__ Z, int N, int ldz, int p0, int np,
                                               int tid, int cs_in = 0) {
    // cs: distance between the columns of two pairs of one point (cs_in != 0: Z is a window into a full 25-point tile)
    const int total = np * N, cf_stride = AX_NPE * N, cs = cs_in ? cs_in : np * ldz;
    const float *cf0 = coef + E.coef_off + (size_t)p0 * N;
    const int law = E.law;
    if constexpr (!FLUID) {
        if (E.att_kind == ATT_NONE && !(PRT && E.prt != 0)) {
            if (law == LAW_ISO) stress_batch<2, 4, NT>(law, Z, cf0, total, cf_stride, cs, ldz, N, tid);
            else if (law == LAW_TI) stress_batch<5, 2, NT>(law, Z, cf0, total, cf_stride, cs, ldz, N, tid);
            else stress_batch<21, 1, NT>(law, Z, cf0, total, cf_stride, cs, ldz, N, tid);
            return;
        }
    }
    int idx = tid;
    int pl = idx / N, pos = idx - pl * N;
    const int dp = NT / N, dpos = NT - dp * N;
    const bool prt = PRT && E.prt != 0;
    const float *px0 = coef + E.prt_off + (size_t)p0 * N;   // X [4][25][N], digit-reversed phi like the moduli
    for (; idx < total; idx += NT) {
        float2 *zc = Z + pl * ldz + pos;
        if constexpr (!FLUID) {
            float ee[6], s[6], px[4] = {1.f, 0.f, 0.f, 1.f};
            if (prt) {   // PRT_3D::sphericalToUndulated (PRT_3D.cpp:41-62): 5 Z-form pairs = 9 components -> 6 strains
#pragma unroll
                for (int k = 0; k < 4; ++k) px[k] = __ldcs(px0 + (size_t)k * cf_stride + idx);
                const float2 q0 = zc[0], q1 = zc[cs], q2 = zc[2 * cs], q3 = zc[3 * cs], q4 = zc[4 * cs];
                const float e9[9] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y, q3.x, q3.y, q4.x};
                prt_s2u_solid(e9, px, ee);
            } else {
                const float2 z0 = zc[0], z1 = zc[cs], z2 = zc[2 * cs];
                ee[0] = z0.x; ee[1] = z0.y; ee[2] = z1.x; ee[3] = z1.y; ee[4] = z2.x; ee[5] = z2.y;
            }
            const float *cf = cf0 + idx;   // [k][point][pos] with local point * N + pos == idx
            stress_law<float>(law, ee, s, [&](int k) { return __ldcs(cf + (size_t)k * cf_stride); });
            const int p = p0 + pl;
            const int Pn = E.att_kind == ATT_CG4 ? 4 : AX_NPE;
            const int q = E.att_kind == ATT_NONE ? -1 : E.att_kind == ATT_CG4 ? cg4_index(p) : p;   // ATT_NONE reaches here only with PRT
            if (q >= 0) {
                const float *ap = attpar + E.att_par_off;
                const float *mod = ap + 3 * E.nsls;
                float *stt = attstate + E.att_state_off;
                const size_t cell = (size_t)q * N + pos;
                const size_t PN = (size_t)Pn * N, sl = 6 * PN;
                const int nsls = E.nsls;
                attenuation_cell<float>(
                    nsls, ap, mod[cell], mod[PN + cell], mod[2 * PN + cell], E.do_kappa != 0, ee, s,
                    [&](int k, int c) -> float & { return stt[k * sl + c * PN + cell]; },
                    [&](int c) -> float & { return stt[nsls * sl + c * PN + cell]; });
            }
            if (prt) {   // PRT_3D::undulatedToSpherical (PRT_3D.cpp:64-86)
                float s9[9];
                prt_u2s_solid(s, px, s9);
                zc[0] = make_float2(s9[0], s9[1]);
                zc[cs] = make_float2(s9[2], s9[3]);
                zc[2 * cs] = make_float2(s9[4], s9[5]);
                zc[3 * cs] = make_float2(s9[6], s9[7]);
                zc[4 * cs] = make_float2(s9[8], 0.f);
            } else {
                zc[0] = make_float2(s[0], s[1]);
                zc[cs] = make_float2(s[2], s[3]);
                zc[2 * cs] = make_float2(s[4], s[5]);
            }
        } else {
            const float K = __ldcs(cf0 + idx);
            const float2 a = zc[0], b = zc[cs];
            if (prt) {   // PRT_3D on the fluid's 3 components (PRT_3D.cpp:21-39) around Acoustic3D::strainToStress
                float px[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) px[k] = __ldcs(px0 + (size_t)k * cf_stride + idx);
                float e3[3] = {a.x, a.y, b.x};
                prt_s2u_fluid(e3, px);
#pragma unroll
                for (int c = 0; c < 3; ++c) e3[c] *= K;
                prt_u2s_fluid(e3, px);
                zc[0] = make_float2(e3[0], e3[1]);
                zc[cs] = make_float2(e3[2], 0.f);
            } else {
                zc[0] = cscale(a, K);
                zc[cs] = make_float2(b.x * K, 0.f);
            }
        }
        pl += dp;
        pos += dpos;
        if (pos >= N) { pos -= N; ++pl; }
    }
}

// Row groups of an element processed in ng passes: rows [row_begin(ng, g), row_begin(ng, g + 1)); ng in {1, 2, 3, 5}.
__host__ __device__ __forceinline__ constexpr int fused_rows_per_group(int ng) { return (5 + ng - 1) / ng; }
__host__ __device__ __forceinline__ constexpr int fused_row_begin(int ng, int g) {
    return g * fused_rows_per_group(ng) < 5 ? g * fused_rows_per_group(ng) : 5;
}
// ng = 10: a pass is part of ONE GLL row -- columns 0-2 or 3-4 -- for the longest expansions (a whole row of Nr = 2016 does not
// fit).  Points of pass g: [p0, p0 + np); most points of any pass of an element with ng passes: fused_np_max(ng).
#define AX_ITEM_ALL 15        // work-item code (element << 4 | g): g = 15 = every pass of the element in turn
__host__ __device__ __forceinline__ constexpr int fused_np_max(int ng) { return ng == 10 ? 3 : 5 * fused_rows_per_group(ng); }
__host__ __device__ __forceinline__ void fused_group(int ng, int g, int &p0, int &np) {
    if (ng == 10) {
        p0 = 5 * (g >> 1) + ((g & 1) ? 3 : 0);
        np = (g & 1) ? 2 : 3;
    } else {
        p0 = 5 * fused_row_begin(ng, g);
        np = 5 * (fused_row_begin(ng, g + 1) - fused_row_begin(ng, g));
    }
}

// One pass = one row group of one element.  E, P: descriptor and plan of this element (shared memory).  On entry the
// first gather tile of this pass (`first_mt` modes) is in flight (cp.async); `after_first_sync` runs behind the first
// barrier of the pass, `after_grad` once U is dead (it starts the next pass's gather).
template <bool FLUID, int NT, int NWW, typename GatherFn, typename AfterSyncFn, typename AfterGradFn>
__device__ __forceinline__ void fused_pass(const FusedCtx<FLUID> &cx, const ElemDesc &E, const FftPlan &P, int g, int first_mt, int tid,
                                           GatherFn gather, AfterSyncFn after_first_sync, AfterGradFn after_grad) {
    constexpr int NC = FLUID ? 1 : 3, NPAIR = FLUID ? 2 : 3;
    constexpr int US = NC * AX_NPE;                   // row stride of U (odd)
    constexpr int NHW = NT / 16;
    constexpr int QIT = 4;                            // in-group (point, 16-mode chunk) items per half-warp and quad tile: their pointwise term r lives in registers
    float2 *const U = cx.U, *const TW = cx.TW, *const Z = cx.Z;
    const int hw = tid >> 4, t = tid & 15;
    const int N = E.nr, nu = N / 2, M = nu + 1, Mt = E.mt;
    const int ldz = fused_ldz(N);
    const bool nyq = (N & 1) == 0;
    const bool axial = E.axial != 0, tiso = !FLUID && E.tiso != 0;
    const int ng = E.ng;
    int p0, np;                                       // points of this pass: [p0, p0 + np)
    fused_group(ng, g, p0, np);
    const bool partial = ng == 10;                    // part of one GLL row: columns [j0, j1) of row r0
    const int r0 = p0 / 5, r1 = partial ? r0 + 1 : r0 + np / 5;
    const int j0 = p0 - 5 * r0, j1 = partial ? j0 + np : 5;
    const int ncols = NPAIR * np, cs = np * ldz;      // column of (pair pr, local point pl) = Z + (pr * np + pl) * ldz

    // Static thread mapping of the pass: kpp = NHW / np half-warps share one point of the group and interleave its 16-mode
    // chunks (np = 25: one half-warp per point; 15: two; 10: three; 5: six, at NHW = 32), lanes run over the 16 modes of a chunk.
    // Everything that depends on the point only (G coefficients, geometry, rotation) is loaded once per pass.
    const int kpp = max(1, NHW / np);
    const int pl = hw / kpp, sub = hw - pl * kpp;
    const bool act = pl < np;
    const int p = p0 + (act ? pl : 0), i = p / 5, j = p - 5 * i;
    const bool ax0 = axial && i == 0;
    GCoef gc;
    load_gcoef(gc, axial, i, j);
    PointGeom gm = {0.f, 0.f, 0.f, 0.f, 0.f};   // geometry / rotation of the point: read behind the first barrier of the pass (the
    float tr[4] = {0.f, 1.f, 0.f, 1.f};          // geometry tile of a new element arrives with its first gather)

    // ------------------------------------------------------------ gather (prefetched) + grad, one tile of modes at a time   @phase gather wait + grad
    for (int a0 = 0, mt = first_mt; a0 < M; a0 += mt, mt = min(Mt, M - a0)) {
        if (a0) {
            cta_sync<NT, NWW>();
            gather(E, a0, mt, -1);
        }
        cp_async_wait_all();
        if (a0 == 0 && t == 0)   // Im(u) of mode 0 is not used (Gradient.cpp:209-224): this thread copied these entries
            for (int row = hw; row < US; row += NHW) U[row].y = 0.f;
        cta_sync<NT, NWW>();
        if (a0 == 0) {
            after_first_sync();
            gm = load_geom(cx.sgeom, 0, p);
            if (tiso) {
#pragma unroll
                for (int k = 0; k < 4; ++k) tr[k] = cx.sgeom[(5 + k) * AX_NPE + p];
            }
        }
        if (act) {
            float2 *zp = Z + pl * ldz;
            for (int a = sub * 16 + t; a < mt; a += 16 * kpp) {
                const int alpha = a0 + a;
                const bool dead = nyq && alpha == nu;
                if constexpr (!FLUID) {
                    float2 ee[6];
                    grad6_mm(U + a * US, i, j, gc, gm, (float)alpha, ax0, ee);
                    if (dead) {
#pragma unroll
                        for (int c = 0; c < 6; ++c) ee[c] = czero();
                    }
                    if (tiso) rot_spz_to_rtz(ee, tr[0], tr[1], tr[2], tr[3]);
#pragma unroll
                    for (int pr = 0; pr < 3; ++pr) zform_store(zp + pr * cs, N, alpha, ee[2 * pr], ee[2 * pr + 1]);
                } else {
                    float2 ee[3];
                    grad_fluid_mm(U + a * US, i, j, gc, gm, (float)alpha, ax0, ee);
                    if (dead) ee[0] = ee[1] = ee[2] = czero();
                    zform_store(zp, N, alpha, ee[0], ee[1]);
                    zform_store(zp + cs, N, alpha, ee[2], czero());
                }
            }
        }
    }
    cta_sync<NT, NWW>();   // Z complete, U dead
    after_grad();      // @phase next-pass gather issue

    // ------------------------------------------------------------ c2r (SolverFFTW_N6::computeC2R, unnormalised, sign +)   @phase c2r
    {
        int L = N;
        for (int s = 0; s < P.nstages; ++s) {
            const int R = P.radix[s];
            fused_stage_dispatch<+1, true, NT>(R, Z, N, L, ncols, TW + (P.stw_off[s] - P.stw_base), tid);
            L /= R;
            cta_sync<NT, NWW>();
        }
    }

    // ------------------------------------------------------------ physical space: stress (+ SLS attenuation)   @phase stress
    physical_space<FLUID, NT>(E, cx.coef, cx.attpar, cx.attstate, Z, N, ldz, p0, np, tid);
    cta_sync<NT, NWW>();

    // ------------------------------------------------------------ r2c (computeR2C; the 1/Nr is applied at load below)   @phase r2c
    {
        int L = 1;
        for (int s = P.nstages - 1; s >= 0; --s) {
            const int R = P.radix[s];
            L *= R;
            fused_stage_dispatch<-1, false, NT>(R, Z, N, L, ncols, TW + (P.stw_off[s] - P.stw_base), tid);
            cta_sync<NT, NWW>();
        }
    }

    // ------------------------------------------------------------ quad + scatter, QIT * kpp 16-mode chunks at a time   @phase quad-pre
    // Same mapping as grad: the half-warp keeps its point; chunk q of its tile is c0 + q * kpp + sub, pre and post, so that the
    // pointwise term r never leaves its registers.
    const float sc = 1.f / (float)N;   // SolverFFTW_N6::computeR2C scaling (SolverFFTW_N6.cpp:47-48)
    const int nchq = (M + 15) >> 4;
    const int nout = AX_NPE - np;      // points outside the group (full rows): xi-partial sums only
    const int nlive = E.pt_nlive[p], st = E.pt_stride[p];
    float2 *const dst = cx.stiff + (size_t)E.pt_off[p];
    float2 *const zp = Z + pl * ldz;
    for (int c0 = 0; c0 < nchq; c0 += QIT * kpp) {
        float2 r[QIT][NC];
        // pointwise half, in place: slot beta <- X, slot N - beta <- Y (beta = 0: the spare slot N)
        if (act) {
#pragma unroll
            for (int q = 0; q < QIT; ++q) {
                const int beta = (c0 + q * kpp + sub) * 16 + t;
                if (beta < M && !(nyq && beta == nu)) {
                    if constexpr (!FLUID) {
                        float2 sg[6], X[3], Y[3];
#pragma unroll
                        for (int pr = 0; pr < 3; ++pr) zform_load(zp + pr * cs, N, beta, sc, sg[2 * pr], sg[2 * pr + 1]);
                        if (tiso) rot_rtz_to_spz(sg, tr[0], tr[1], tr[2], tr[3]);
                        quad6_pre(sg, gm, (float)beta, ax0, X, Y, r[q]);
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            zp[c * cs + beta] = X[c];
                            zp[c * cs + N - beta] = Y[c];
                        }
                    } else {
                        float2 sg[3], X, Y, dummy;
                        zform_load(zp, N, beta, sc, sg[0], sg[1]);
                        zform_load(zp + cs, N, beta, sc, sg[2], dummy);
                        quad_fluid_pre(sg, gm, (float)beta, ax0, X, Y, r[q][0]);
                        zp[beta] = X;
                        zp[N - beta] = Y;
                    }
                }
            }
        }
        cta_sync<NT, NWW>();
        // tensor-product half + Point::gatherStiffFromElement (SolidPoint.cpp:197-209)   @phase quad-post + scatter
        if (act) {
#pragma unroll
            for (int q = 0; q < QIT; ++q) {
                const int beta = (c0 + q * kpp + sub) * 16 + t;
                if (beta < M && !(nyq && beta == nu) && beta < nlive) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        float2 f = r[q][c];
                        const float2 *zx = Z + c * cs + j * ldz + beta;                         // X(k, j), k = r0 .. r1 - 1
                        const float2 *zy = Z + c * cs + (i - r0) * 5 * ldz + N - beta;          // Y(i, k), k = 0 .. 4
                        if (partial) {            // one row, columns [j0, j1): X(r0, j) and Y(r0, k), k in [j0, j1), at local index k - j0
                            const float2 *zr = Z + c * cs - j0 * ldz;
#pragma unroll
                            for (int k = 0; k < 5; ++k) {
                                if (k == r0) f = cfma(gc.gxi_row[k], zr[j * ldz + beta], f);
                                if (k >= j0 && k < j1) f = cfma(gc.geta_row[k], zr[k * ldz + N - beta], f);
                            }
                        } else if (ng == 1) {
#pragma unroll
                            for (int k = 0; k < 5; ++k) {
                                f = cfma(gc.gxi_row[k], zx[k * 5 * ldz], f);
                                f = cfma(gc.geta_row[k], zy[k * ldz], f);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 5; ++k) {   // static register indices: the row range is a (warp-uniform) predicate
                                if (k >= r0 && k < r1) f = cfma(gc.gxi_row[k], zx[(k - r0) * 5 * ldz], f);
                                f = cfma(gc.geta_row[k], zy[k * ldz], f);
                            }
                        }
                        if (beta == 0) f.y = 0.f;
                        atomicAdd(dst + (size_t)c * st + beta, make_float2(-f.x, -f.y));   // stiff -= f (RED.ADD.F32x2)
                    }
                }
            }
        }
        // points outside the pass: f(i', j) += sum_{k in group} G_xi(i', k) X(k, j); a partial row also owes the rest of its row
        // the eta-sums f(r0, j') += sum_{k in [j0, j1)} G_eta(j', k) Y(r0, k)
        const int nout_x = partial ? 4 * np : nout, nout_all = partial ? 4 * np + (5 - np) : nout;
        if (nout_all) {
            const int nchs = min(QIT * kpp, nchq - c0);
            for (int it = hw; it < nout_all * nchs; it += NHW) {
                const int po = it / nchs, ch = it - po * nchs;
                int p2;
                const bool xi_target = po < nout_x;
                if (!partial) p2 = po < p0 ? po : po + np;          // skip [p0, p0 + np)
                else if (xi_target) { const int ii = po / np; p2 = 5 * (ii < r0 ? ii : ii + 1) + j0 + (po - ii * np); }
                else { const int q2 = po - nout_x; p2 = 5 * r0 + (q2 < j0 ? q2 : q2 + np); }
                const int i2 = p2 / 5, j2 = p2 - 5 * i2;
                const int beta = (c0 + ch) * 16 + t;
                if (beta < M && !(nyq && beta == nu) && beta < E.pt_nlive[p2]) {
                    const float *Gxi = c_G[axial ? 1 : 0];
                    const float *Geta = c_G[0];
                    float2 *const dst2 = cx.stiff + (size_t)E.pt_off[p2];
                    const int st2 = E.pt_stride[p2];
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        float2 f = czero();
                        if (!partial) {
                            const float2 *zx = Z + c * cs + j2 * ldz + beta;
                            for (int k = r0; k < r1; ++k) f = cfma(Gxi[i2 * 5 + k], zx[(k - r0) * 5 * ldz], f);
                        } else if (xi_target) {
                            f = cscale(Z[c * cs + (j2 - j0) * ldz + beta], Gxi[i2 * 5 + r0]);
                        } else {
                            const float2 *zy = Z + c * cs + N - beta;
                            for (int k = j0; k < j1; ++k) f = cfma(Geta[j2 * 5 + k], zy[(k - j0) * ldz], f);
                        }
                        if (beta == 0) f.y = 0.f;
                        atomicAdd(dst2 + (size_t)c * st2 + beta, make_float2(-f.x, -f.y));
                    }
                }
            }
        }
    }
    // the barrier after cp_async_wait_all at the top of the next pass separates these reads of Z from its grad
}

// ---------------------------------------------------------------- split pipeline, middle kernel
// Elements whose spectrum does not fit one SM's shared memory go through k_grad3d -> k_fft3d_v2 -> k_quad3d with the
// Z-form spectrum staged in an L2-resident scratch ring ([pair][25][Nr] per element).  This kernel is the FFT / physical
// space part for NP consecutive GLL points of one element (the constitutive law couples the 6 components of one point,
// never two points): load NPAIR * NP columns, c2r, stress (+SLS), r2c, store -- the same stage and stress code as the
// fused kernel, on a tile that is 1/5 (NP = 5) or 1/25 (NP = 1) of an element, so that even Nr = 2016 fits.
#ifndef AX_FFT_MIN_CTAS
#define AX_FFT_MIN_CTAS 4   // resident 256-thread CTAs per SM the register allocation of k_fft3d_v2 must allow (64 registers, 76 B
                            // of spills).  B200, cfg3 (profiles/microbench/occ_ab*.sh): 2 -> 1.271, 3 -> 1.108, 4 -> 1.099 ms per step
#endif
template <bool FLUID, int NP, int NT, int NPAIR_ = 0, bool PRT = false>
__global__ void __launch_bounds__(NT, NT <= 256 ? AX_FFT_MIN_CTAS : NT <= 512 ? AX_FFT_MIN_CTAS / 2 : 1) k_fft3d_v2(const ElemDesc *__restrict__ elems, const FftItem *__restrict__ items,
                                                 const FftPlan *__restrict__ plans, const float2 *__restrict__ stwpool,
                                                 const float *__restrict__ coef, const float *__restrict__ attpar,
                                                 float *__restrict__ attstate, float2 *__restrict__ scratch) {
    constexpr int NPAIR = NPAIR_ ? NPAIR_ : (FLUID ? 2 : 3), NCOLS = NPAIR * NP;   // NPAIR_ = 5: solid elements with PRT (9 components)
    constexpr int PLAN_W = (int)(sizeof(FftPlan) / sizeof(int));
    static_assert(PLAN_W <= NT, "plan loader");
    extern __shared__ __align__(16) float2 smem[];
    __shared__ FftPlan sP;
    const int tid = threadIdx.x;
    const FftItem it = items[blockIdx.x];
    const ElemDesc &E = elems[it.elem];
    if (tid < PLAN_W) reinterpret_cast<int *>(&sP)[tid] = reinterpret_cast<const int *>(plans + E.plan_id)[tid];
    const int N = E.nr, ldz = fused_ldz(N), p0 = it.p0;
    __syncthreads();
    const FftPlan &P = sP;
    float2 *const TW = smem, *const Z = smem + ((P.stw_len + 1) & ~1);
    for (int k = tid; k < P.stw_len; k += NT) TW[k] = stwpool[P.stw_base + k];
    float2 *const gz = scratch + E.scratch_off;
#pragma unroll
    for (int c = 0; c < NCOLS; ++c) {
        const int pr = c / NP, pl = c - pr * NP;
        const float2 *src = gz + ((size_t)pr * AX_NPE + p0 + pl) * N;
        float2 *dst = Z + c * ldz;
        for (int k = tid; k < N; k += NT) dst[k] = __ldcs(src + k);
    }
    __syncthreads();
    {
        int L = N;
        for (int s = 0; s < P.nstages; ++s) {
            const int R = P.radix[s];
            fused_stage_dispatch<+1, true, NT>(R, Z, N, L, NCOLS, TW + (P.stw_off[s] - P.stw_base), tid);
            L /= R;
            __syncthreads();
        }
    }
    physical_space<FLUID, NT, PRT>(E, coef, attpar, attstate, Z, N, ldz, p0, NP, tid);
    __syncthreads();
    {
        int L = 1;
        for (int s = P.nstages - 1; s >= 0; --s) {
            const int R = P.radix[s];
            L *= R;
            fused_stage_dispatch<-1, false, NT>(R, Z, N, L, NCOLS, TW + (P.stw_off[s] - P.stw_base), tid);
            __syncthreads();
        }
    }
#pragma unroll
    for (int c = 0; c < NCOLS; ++c) {
        const int pr = c / NP, pl = c - pr * NP;
        float2 *dst = gz + ((size_t)pr * AX_NPE + p0 + pl) * N;
        const float2 *src = Z + c * ldz;
        for (int k = tid; k < N; k += NT) dst[k] = src[k];
    }
}

// ---------------------------------------------------------------- in-kernel Newmark   @phase newmark warps
// The point update (SolidPoint::updateNewmark, SolidPoint.cpp:23-38) is HBM-bound, the element pipeline above is
// latency-bound and leaves HBM ~93 % idle.  So the update of step n+1 runs *under* the element kernel of step n: a "plain"
// solid point (pure solid, Mass1D, not axial, not on a halo, not in a source element, touched by fused elements only) can
// be advanced as soon as the last element that touches it has scattered its force.  Compute warps count arrivals per point
// (nw.cnt) and push completed points into a ready queue; NWW specialised warps per CTA (and, once the element queue is
// empty, the compute warps too) pop points, fetch stiff/accel/veloc/displ with TMA bulk loads into a double-buffered
// shared-memory stage (no registers held across the HBM latency), update, and store with plain coalesced stores.
// All other points are advanced by k_newmark_* at the start of the next step, as before.
struct NwArgs {
    int on;                    // 0: this launch does no in-kernel Newmark (single step, last step of a run, verb-wise calls)
    int n_plain;               // queue length = number of plain points
    int *cnt;                  // [ns] arrivals per point (consumer resets to 0)
    int *queue;                // [n_plain] ready queue, -1 = empty (consumer resets)
    unsigned *ctl;             // [0] tail (producers), [1] head (consumers), [2] finished warps; last warp re-arms
    const unsigned *off;       // per solid point: block offset (float2 units), Nu, Nr, inverse mass
    const int *nu;
    const int *nr;
    const float *invmass;
    float2 *displ, *veloc, *accel, *stiff;
    float half_dt, dt, half_dt_dt;
    unsigned long long *dbg;   // optional [grid][4] globaltimer stamps: start, elements done, consumers done (AX3D_NW_DEBUG)
};

#ifndef AX_FUSED_NT
#define AX_FUSED_NT 512        // compute threads of the fused element kernel (one CTA per SM: 512 -> 128 registers per thread)
#endif
#ifndef AX_NWW
#define AX_NWW 2               // Newmark warps per CTA of the solid launch
#endif
#ifndef AX_NW_NT
#define AX_NW_NT (AX_FUSED_NT - 32 * AX_NWW)   // compute threads of that launch
#endif
#ifndef NW_CHR
#define NW_CHR 96              // rows (complex entries) per chunk (160: 31 KB of stages per CTA, 96: 19 KB)
#endif
#define NW_CHS (NW_CHR + 2)    // stage row capacity: the 16-byte aligned superset of a chunk
#ifndef NW_NSTAGE
#define NW_NSTAGE 3
#endif
#define NW_WARP_SMEM (NW_NSTAGE * 4 * NW_CHS * 8)   // bytes of stage buffers per consumer warp

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(a), "r"(parity) : "memory");
}
// 1D TMA bulk load global -> shared, completion on an mbarrier (bytes and both addresses are multiples of 16)
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst_smem)),
                 "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// One consumer warp.  stage: NW_WARP_SMEM bytes of shared memory owned by this warp; bar: NW_NSTAGE mbarriers owned by it.
// service(): called by the whole warp whenever it would otherwise wait (Newmark warp 0 uses it to post the arrivals of
// the elements its CTA has finished; a no-op elsewhere).
template <typename ServiceFn>
__device__ __forceinline__ void nw_consumer(const NwArgs &nw, float2 *stage, unsigned long long *bar, int lane, ServiceFn service) {
    int cur_p = -1, cur_r = 0, cur_total = 0;          // point being chunked (warp-uniform)
    unsigned cur_off = 0;
    int q_p[NW_NSTAGE], q_r0[NW_NSTAGE], q_n[NW_NSTAGE], q_skew[NW_NSTAGE];
    unsigned q_g0[NW_NSTAGE], q_par[NW_NSTAGE];
    bool q_ok[NW_NSTAGE];
#pragma unroll
    for (int s = 0; s < NW_NSTAGE; ++s) { q_ok[s] = false; q_par[s] = 0u; }

    // next chunk -> slot s, loads issued by lane 0; false when the queue is exhausted
    auto fetch = [&](int s) -> bool {
        if (cur_p < 0 || cur_r >= cur_total) {
            int p = -1;
            unsigned slot = 0;
            if (lane == 0) slot = atomicAdd(&nw.ctl[1], 1u);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot >= (unsigned)nw.n_plain) return false;
            volatile int *qs = nw.queue + slot;
            for (unsigned spins = 0;; ++spins) {          // warp-wide wait for the producer of this slot
                if (lane == 0) p = *qs;
                p = __shfl_sync(0xffffffffu, p, 0);
                if (p >= 0) break;
                service();
                __nanosleep(200);
                if (spins > (1u << 24)) {                 // ~4 s of waiting for a producer: a bookkeeping error.  Never hang the GPU,
                    if (lane == 0) nw.ctl[3] = 1u;        // never carry on with a point that was not advanced: flag it and abort the
                    __threadfence_system();               // launch (the next API call reports the failed launch)
                    __trap();
                }
            }
            if (lane == 0) {
                __threadfence();                           // acquire: the forces of every element touching p are visible
                asm volatile("fence.proxy.async;\n" ::: "memory");   // ... also to the TMA loads below
                *qs = -1;                                  // re-arm the slot and the arrival counter for the next launch
                nw.cnt[p] = 0;
            }
            __syncwarp();
            cur_p = p;
            cur_r = 0;
            cur_total = 3 * (nw.nu[p] + 1);
            cur_off = nw.off[p];
        }
        const int n = min(NW_CHR, cur_total - cur_r);
        const unsigned g0 = cur_off + (unsigned)cur_r;     // first entry of the chunk (float2 index)
        const unsigned a0 = g0 & ~1u;                      // 16-byte aligned superset [a0, a1)
        const unsigned a1 = (g0 + (unsigned)n + 1u) & ~1u;
        q_p[s] = cur_p; q_r0[s] = cur_r; q_n[s] = n; q_g0[s] = g0; q_skew[s] = (int)(g0 - a0); q_ok[s] = true;
        cur_r += n;
        if (lane == 0) {
            const unsigned bytes = (a1 - a0) * 8u;
            float2 *st = stage + s * 4 * NW_CHS;
            mbar_expect_tx(bar + s, 4u * bytes);
            bulk_load(st + 0 * NW_CHS, nw.stiff + a0, bytes, bar + s);
            bulk_load(st + 1 * NW_CHS, nw.accel + a0, bytes, bar + s);
            bulk_load(st + 2 * NW_CHS, nw.veloc + a0, bytes, bar + s);
            bulk_load(st + 3 * NW_CHS, nw.displ + a0, bytes, bar + s);
        }
        return true;
    };

    bool more = true;
#pragma unroll
    for (int s = 0; s < NW_NSTAGE; ++s)
        if (more) more = fetch(s);
    for (;;) {
        bool any = false;
#pragma unroll
        for (int s = 0; s < NW_NSTAGE; ++s) {   // unrolled: the q_* arrays stay in registers
            if (!q_ok[s]) continue;
            any = true;
            service();
            mbar_wait(bar + s, q_par[s]);
            q_par[s] ^= 1u;
            const int p = q_p[s];
            const int nu = nw.nu[p], stp = nu + 1;
            const bool nyq = (nw.nr[p] & 1) == 0;
            const float im = nw.invmass[p];
            const float2 *st = stage + s * 4 * NW_CHS + q_skew[s];
            for (int i = lane; i < q_n[s]; i += 32) {
                const int row = q_r0[s] + i;
                const int c = row >= 2 * stp ? 2 : (row >= stp ? 1 : 0);
                const int alpha = row - c * stp;
                float2 f = st[i];
                const float2 a_old = st[NW_CHS + i];
                float2 v = st[2 * NW_CHS + i], u = st[3 * NW_CHS + i];
                // SolidPoint::maskField for a non-axial point (SolidPoint.cpp:216-238), M^-1, mask again
                if (alpha == 0) f.y = 0.f;
                if (nyq && alpha == nu) f = czero();
                f = make_float2(f.x * im, f.y * im);
                if (alpha == 0) f.y = 0.f;
                if (nyq && alpha == nu) f = czero();
                newmark_entry(f, a_old, v, u, nw.half_dt, nw.dt, nw.half_dt_dt);
                const size_t g = (size_t)q_g0[s] + i;
                nw.veloc[g] = v;
                nw.accel[g] = f;
                nw.displ[g] = u;
                nw.stiff[g] = czero();
            }
            __syncwarp();            // every lane has read the stage before the next bulk load overwrites it
            q_ok[s] = false;
            if (more) more = fetch(s);
        }
        if (!any) break;
    }
}

// ---------------------------------------------------------------- in-kernel halo put   @phase halo put
// Domain::assembleStiff, send half (Domain.cpp:111-131), overlapped with the interior elements: boundary elements (those that
// touch a point shared with another rank) are first in the work queue of the solid launch; every CTA that finishes one bumps
// `bcnt`, and the CTA that brings it to `nb` -- every boundary force of the rank is then in memory: the other element
// kernels of the step were launched before this one -- couples the solid-fluid points on the halo
// (SolidFluidPoint::coupleSolidFluid comes before the exchange, Newmark.cpp:57-59) and stores every neighbour's segment
// into that neighbour's window over NVLink, then goes back to the element queue.  k_halo_wait_add runs behind the kernel.
struct HaloArgs {
    const HaloTab *tab;   // nullptr: no in-kernel put in this launch
    unsigned *bcnt;       // boundary elements finished in this launch (re-armed by the CTA that sends)
    int nb;               // boundary elements of the launch (0: CTA 0 sends before its first element)
    unsigned *cost;       // nullptr, or [nelem]: SM clock cycles this launch spent on every element (ax3d_measure_costs: the
                          // measured element costs that weight the partition, Mesh.cpp:412-588)
};

template <int NT, int NWW>
__device__ __forceinline__ void halo_put_cta(const HaloTab &ht, const float2 *__restrict__ s_displ, float2 *__restrict__ s_stiff, int tid) {
    __threadfence();   // acquire: the forces every other CTA fenced before its bump of bcnt
    for (int r = tid; r < ht.sf_halo.nrows; r += NT) sf_couple_row(ht.sf_halo, r, s_displ, s_stiff, ht.f_stiff);
    __threadfence();
    cta_sync<NT, NWW>();
    const unsigned s = __ldcg(ht.step);
    for (int n = 0; n < ht.nneigh; ++n) {
        const HaloPeer &P = ht.peer[n];
        float2 *dst = P.win + (size_t)(s & 1u) * P.stride;
        for (int i0 = tid; i0 < P.n; i0 += 4 * NT) {   // four independent index -> force -> remote store chains per thread
            float2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * NT;
                if (i < P.n) {
                    const unsigned k = P.idx[i];
                    v[u] = (k >> 31) ? __ldcg(ht.f_stiff + (k & 0x7fffffffu)) : __ldcg(s_stiff + k);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u * NT < P.n) __stcg(dst + i0 + u * NT, v[u]);
        }
    }
    __threadfence_system();
    cta_sync<NT, NWW>();
    if (tid < ht.nneigh) atomicAdd_system(ht.peer[tid].count, (unsigned)ht.peer[tid].nblocks);
}

// ---------------------------------------------------------------- the kernel   @phase kernel loop
// Work items: items[w] = (element index << 4) | g -- one row-group pass g of one element (g = 15: every pass of the element in
// turn).  Passes of one element are independent (each gathers the displacement itself and scatters with RED), so they are
// scheduled separately: a third of a large element is a finer LPT quantum than the element (ranks whose part of the mesh
// is a few hundred large elements lose 7 % to the quantisation otherwise, profiles/r2_scaling.md).
// grid: persistent, one CTA per SM; block NT compute threads (+ 32 * NWW Newmark threads).  work[0] = next item index
// (starts at gridDim.x), work[1] = number of warps that have finished; the last one re-arms the counters for the next
// launch (graph replay).  tile_cap: float2 capacity R of the tile region (the per-element offsets twoff / zoff of the
// descriptors refer to it); the Newmark warps' stage buffers start right behind it.
#define NW_RING 8   // arrival-code ring (elements finished by the compute warps, not yet posted by Newmark warp 0)
template <bool FLUID, int NT, int NWW>
__global__ void __launch_bounds__(NT + 32 * NWW, 1)
    k_elem3d_fused(const ElemDesc *__restrict__ elems, const int *__restrict__ items, int nitems, const FftPlan *__restrict__ plans,
                   const float2 *__restrict__ stwpool, const float *__restrict__ geom, const float *__restrict__ coef,
                   const float *__restrict__ attpar, float *__restrict__ attstate, const float2 *__restrict__ displ,
                   float2 *__restrict__ stiff, int tile_cap, unsigned *__restrict__ work, const NwArgs nw, const HaloArgs halo) {
    constexpr int NC = FLUID ? 1 : 3;
    constexpr int US = NC * AX_NPE;
    constexpr int NHW = NT / 16;
    constexpr int NWARP = NT / 32 + NWW;
    constexpr int DESC_W = (int)(sizeof(ElemDesc) / sizeof(int)), PLAN_W = (int)(sizeof(FftPlan) / sizeof(int));
    static_assert(DESC_W <= NT && DESC_W + PLAN_W <= NT, "descriptor loaders need NT >= descriptor words");
    extern __shared__ __align__(16) float2 smem[];
    __shared__ ElemDesc sE[2];
    __shared__ FftPlan sP[2];
    __shared__ float sGeom[2][9 * AX_NPE];   // geometry (+ trig) of the current / next element, staged with its first gather
    __shared__ int sIdx[3];   // ring of work-item indices: current, next, next-next
    __shared__ int sCode[3];  // ... and their codes (element << 4 | pass)
    __shared__ unsigned long long sBar[NWARP][NW_NSTAGE];
    __shared__ int sArrCode[NW_RING][AX_NPE];   // ring: pt_nw codes of the elements whose scatter is complete ...
    __shared__ volatile int sArrHead;           // ... up to this count (written by compute thread 0 behind a barrier)
    __shared__ volatile int sArrSeen;           // ... of which Newmark warp 0 has posted this many (back-pressure for the ring)
    __shared__ volatile int sCtaDone;           // the compute warps have handed over their last element
    __shared__ int sPut;                        // this CTA finished the rank's last boundary element: it sends the halo
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const bool nw_on = NWW > 0 && nw.on != 0;

    if (NWW > 0 && nw_on && lane == 0) {
#pragma unroll
        for (int s = 0; s < NW_NSTAGE; ++s) mbar_init(&sBar[warp][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid == 0) { sArrHead = 0; sArrSeen = 0; sCtaDone = 0; }
    auto stamp = [&](int k) {
        if (nw.dbg) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            nw.dbg[blockIdx.x * 4 + k] = t;
        }
    };
    if (tid == 0) stamp(0);
    __syncthreads();   // the only CTA-wide barrier: roles split below

    if (tid < NT) {
        const int hw = tid >> 4, t = tid & 15;
        FusedCtx<FLUID> cx{geom, coef, attpar, attstate, displ, stiff, smem, smem, smem, sGeom[0]};
        float2 *const U = smem;

        // descriptor + plan of element el -> slot s (plain loads; visible after the next barrier)
        auto load_desc = [&](int s, int el) {
            if (tid < DESC_W) reinterpret_cast<int *>(&sE[s])[tid] = reinterpret_cast<const int *>(elems + el)[tid];
            if (tid >= DESC_W && tid < DESC_W + PLAN_W) {
                const int pid = elems[el].plan_id;
                reinterpret_cast<int *>(&sP[s])[tid - DESC_W] = reinterpret_cast<const int *>(plans + pid)[tid - DESC_W];
            }
        };
        // Point::scatterDisplToElement (SolidPoint.cpp:175-195) for modes [a0, a0 + mt): half-warp per (component, point)
        // slot >= 0: first tile of an element -- its geometry goes to sGeom[slot] with the same cp.async group
        auto gather = [&](const ElemDesc &E, int a0, int mt, int slot) {
            if (slot >= 0) {
                const bool ti = !FLUID && E.tiso != 0;
                if (tid < 5 * AX_NPE || (ti && tid >= 128 && tid < 128 + 4 * AX_NPE)) {
                    const float *src = tid < 128 ? geom + E.geom_off + tid : geom + E.trig_off + (tid - 128);
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(&sGeom[slot][tid < 128 ? tid : 5 * AX_NPE + tid - 128]);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src) : "memory");
                }
            }
            const unsigned u0 = (unsigned)__cvta_generic_to_shared(U + t * US);
            for (int row = hw; row < US; row += NHW) {
                const int c = row / AX_NPE, p = row - c * AX_NPE;
                const float2 *src = displ + (size_t)E.pt_off[p] + (size_t)c * E.pt_stride[p] + a0 + t;
                const int nlive = E.pt_nlive[p] - a0 - t;   // entries [0, nlive) of this lane's stride-16 sequence are live
                unsigned dst = u0 + row * 8u;
                for (int a = 0; a < mt - t; a += 16) {
                    const int sz = a < nlive ? 8 : 0;        // src-size 0: nothing is read, the 8 bytes are zero-filled
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(sz ? src + a : displ), "r"(sz) : "memory");
                    dst += 16u * US * 8u;
                }
            }
        };
        int e = blockIdx.x;     // work-item index
        const int nelem = nitems;
        int n_done = 0;     // elements of this CTA whose scatter has been issued (their codes are in the ring)
        if (!FLUID && halo.tab != nullptr && halo.nb == 0 && blockIdx.x == 0) halo_put_cta<NT, NWW>(*halo.tab, displ, stiff, tid);
        if (e < nelem) {
            const int code0 = items[e];
            load_desc(0, code0 >> 4);
            if (tid == 0) {
                sCode[0] = code0;
                const int nx = (int)atomicAdd(&work[0], 1u);
                sIdx[1] = nx;
                sCode[1] = nx < nelem ? items[nx] : 0;
            }
            cta_sync<NT, NWW>();
            int first_mt = min(sE[0].mt, sE[0].nu + 1);   // modes of the gather tile in flight (first tile of the coming pass)
            gather(sE[0], 0, first_mt, 0);
            int tw_plan = -1, tw_at = -1;
            for (int it = 0, k = 0; e < nelem; it ^= 1, k = (k == 2 ? 0 : k + 1)) {
                const ElemDesc &E = sE[it];
                const FftPlan &P = sP[it];
                cx.sgeom = sGeom[it];
                cx.TW = smem + E.twoff;
                cx.Z = smem + E.zoff;
                const int kn = k == 2 ? 0 : k + 1, knn = kn == 2 ? 0 : kn + 1;   // ring slots of the next two elements
                const int ng = E.ng;
                const int gsel = sCode[k] & 15, el = sCode[k] >> 4;
                const int g0 = gsel == AX_ITEM_ALL ? 0 : gsel, g1 = gsel == AX_ITEM_ALL ? ng : gsel + 1;
                // moduli of the points of this item ([k][25][Nr]: rows [5 r0, 5 r1) of every modulus) -> L2 while gather/grad/c2r run
                {
                    const int ncoef = FLUID ? 1 : (E.law == LAW_ISO ? 2 : E.law == LAW_TI ? 5 : 21);
                    int pa, pn, pb, pm;
                    fused_group(ng, g0, pa, pn);
                    fused_group(ng, g1 - 1, pb, pm);
                    pb += pm;
                    const int per = ((pb - pa) * E.nr + 31) / 32;          // 128-byte lines per modulus
                    const float *cb = coef + E.coef_off + (size_t)pa * E.nr;
                    for (int q = tid; q < ncoef * per; q += NT) {
                        const int kc = q / per, ln = q - kc * per;
                        prefetch_l2(cb + (size_t)kc * AX_NPE * E.nr + (size_t)ln * 32);
                    }
                }
                long long t_el = 0;
                if (halo.cost != nullptr && tid == 0) t_el = clock64();
                for (int g = g0; g < g1; ++g) {
                    const bool last = g + 1 == g1;
                    int next_mt = 0;
                    // next element: descriptor behind the first barrier of the element, displacement (cp.async into the dead U) behind grad
                    auto after_first_sync = [&]() {
                        if (g != g0) return;
                        if (nw_on && tid == 0) {   // the previous element's scatter is complete (barrier): hand it to Newmark warp 0
                            __threadfence_block();
                            sArrHead = n_done;
                        }
                        const int en = sIdx[kn];   // fetched during the previous item
                        if (en < nelem) load_desc(it ^ 1, sCode[kn] >> 4);
                        // twiddle tables of this element's plan.  Its TW region may overlap the previous element's Z: written only
                        // now, behind the barrier every thread passes after the previous element's last read of Z; the
                        // barrier behind grad publishes it before the first FFT stage
                        if (tw_plan != E.plan_id || tw_at != E.twoff) {
                            const int tb = P.stw_base;
                            for (int q = tid; q < P.stw_len; q += NT) cx.TW[q] = stwpool[tb + q];
                            tw_plan = E.plan_id;
                            tw_at = E.twoff;
                        }
                    };
                    auto after_grad = [&]() {
                        if (!last) {               // next row group of this element: same displacement again
                            next_mt = min(E.mt, E.nu + 1);
                            gather(E, 0, next_mt, -1);
                            return;
                        }
                        if (tid == 0) {            // needed one item from now: latency hidden
                            const int nx = (int)atomicAdd(&work[0], 1u);
                            sIdx[knn] = nx;
                            sCode[knn] = nx < nelem ? items[nx] : 0;
                        }
                        if (sIdx[kn] < nelem) {
                            const ElemDesc &En = sE[it ^ 1];
                            // the first tile of the next element lands in U while this element's TW / Z are live: stay below both
                            next_mt = min(min(En.mt, En.nu + 1), (E.twoff / US) & ~15);
                            gather(En, 0, next_mt, it ^ 1);
                        }
                    };
                    fused_pass<FLUID, NT, NWW>(cx, E, P, g, first_mt, tid, gather, after_first_sync, after_grad);
                    first_mt = next_mt;
                }
                if (halo.cost != nullptr && tid == 0) atomicAdd(&halo.cost[el], (unsigned)(clock64() - t_el));
                if (!FLUID && halo.tab != nullptr && E.bnd) {
                    __threadfence();          // this thread's scatter of the element before the count
                    cta_sync<NT, NWW>();
                    if (tid == 0) {
                        const unsigned old = atomicAdd(halo.bcnt, 1u);
                        sPut = (old + 1u == (unsigned)halo.nb) ? 1 : 0;
                        if (sPut) *halo.bcnt = 0u;   // re-arm for the next launch (graph replay)
                    }
                    cta_sync<NT, NWW>();
                    if (sPut) halo_put_cta<NT, NWW>(*halo.tab, displ, stiff, tid);
                }
                if (nw_on) {
                    if (tid < AX_NPE) {
                        while (n_done - sArrSeen >= NW_RING) __nanosleep(100);   // ring full: Newmark warp 0 is behind
                        sArrCode[n_done & (NW_RING - 1)][tid] = E.pt_nw[tid];
                    }
                    ++n_done;
                }
                e = sIdx[kn];
            }
            if (nw_on) {
                cta_sync<NT, NWW>();   // last element's scatter complete; shared memory is free from here on
                if (tid == 0) {
                    __threadfence_block();
                    sArrHead = n_done;
                }
            }
        }
        if (nw_on && tid == 0) {
            __threadfence_block();
            sCtaDone = 1;
        }
        if (tid == 0) stamp(1);
        // element queue empty: the compute warps join the Newmark consumers (their stages live in the now idle tile memory)
        if (nw_on && (warp + 1) * NW_WARP_SMEM <= tile_cap * (int)sizeof(float2)) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy tile traffic before the TMA writes
            nw_consumer(nw, smem + warp * (NW_WARP_SMEM / 8), sBar[warp], lane, [] {});
        }
    } else if (NWW > 0) {
        if (nw_on) {
            float2 *stage = smem + ((tile_cap + 1) & ~1) + (warp - NT / 32) * (NW_WARP_SMEM / 8);
            if (warp == NT / 32) {
                // Newmark warp 0 posts the arrivals: a point whose last element has scattered goes to the ready queue.
                // (Kept off the compute warps: the release fence would sit on their critical path once per element.)
                int seen = 0;
                auto service = [&]() {
                    const int h = sArrHead;
                    if (seen < h) {
                        __threadfence_block();
                        __threadfence();   // release: the CTA's forces (ordered before sArrHead by its barrier) before the counts
                        for (; seen < h; ++seen) {
                            const int code = lane < AX_NPE ? sArrCode[seen & (NW_RING - 1)][lane] : -1;
                            if (code >= 0) {
                                const int p = code & 0xffffff, need = code >> 24;
                                if (atomicAdd(&nw.cnt[p], 1) + 1 == need) {
                                    const unsigned slot = atomicAdd(&nw.ctl[0], 1u);
                                    __threadfence();
                                    atomicExch(&nw.queue[slot], p);
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) sArrSeen = seen;
                    }
                };
                nw_consumer(nw, stage, sBar[warp], lane, service);
                // the ready queue may be fully claimed long before this CTA's last element has scattered
                while (!sCtaDone) {
                    service();
                    __nanosleep(500);
                }
                service();
            } else {
                nw_consumer(nw, stage, sBar[warp], lane, [] {});
            }
        }
    }
    // re-arm the counters once every warp of every CTA is done
    __syncwarp();
    if (tid == NT) stamp(2);
    if (tid == 0) stamp(3);
    if (lane == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&work[1], 1u);
        if (done == gridDim.x * NWARP - 1) {
            work[0] = gridDim.x;
            work[1] = 0u;
            if (NWW > 0 && nw_on) { nw.ctl[0] = 0u; nw.ctl[1] = 0u; }
            __threadfence();
        }
    }
}
