// fused.cuh -- persistent one-CTA-per-element kernel for elements with 3D (phi-dependent) material whose whole
// strain spectrum fits in shared memory: SolidElement::computeStiff / FluidElement::computeStiff
// (SolidElement.cpp:43-65, 404-443; FluidElement.cpp:43-65, 333-355) with no HBM/L2 round trip between
// gather -> grad -> [rotate] -> c2r -> stress(+SLS) -> r2c -> [rotate^-1] -> quad -> scatter.
//
// Shared memory of one CTA (float2 units), fixed offsets for every element of a launch:
//   U  [u_cap]             gathered displacement, mode-major: U[a * NC*25 + c * 25 + point] (all contraction operands
//                           of one thread sit at immediate offsets; the odd row stride keeps lanes = modes conflict free);
//   TW [tw_cap]            per-stage twiddle tables of the plan;
//   Z  [NPAIR * 25][ldz]   "Z-form" columns: two real strain/stress components of one GLL point as one complex
//                           column of length N = Nr.  ldz = (N + 1) | 1 is odd, so lanes that run over columns are
//                           bank-conflict free, and slot N of every column is a spare.
// One CTA takes elements off a device-side work counter (elements are sorted by cost, largest first, so the queue is
// LPT-like).  While element e is in its FFT/stress/quad phases the displacement of the NEXT element streams into U
// with cp.async (U is dead after grad: the quad phase keeps its pointwise term in registers), its descriptor and plan
// are prefetched, and the moduli of e are prefetched into L2 at the top of e.  Phases of one element:
//   grad (thread = (mode, point), writes Z-form) | DIF stages (thread = (column, butterfly), column fastest)
//   | stress (thread = (point, phi)) | DIT stages | quad-pre (in place: slot beta <- X, slot N - beta <- Y)
//   | quad-post + scatter (RED.ADD.F32x2).
// The element body exists twice in a kernel instance: generic (driven by the FftPlan at run time) and, when NCT1 > 0,
// specialised for the one compile-time Nr = NCT1 (strides, radices and twiddle offsets become immediates) -- the host
// picks the instance whose NCT1 is the most frequent Nr of the domain.
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

// ---------------------------------------------------------------- one FFT stage over all columns
// DIF (c2r, SIGN = +1): butterfly, then twiddle T[j][p].  DIT (r2c, SIGN = -1): conj twiddle, then butterfly.
// NCT > 0: N = NCT and L = LCT are compile-time.
template <int R, int SIGN, bool DIF, int NT, int NCOLS, int NCT, int LCT>
__device__ __forceinline__ void fused_stage(float2 *__restrict__ z, int N_rt, int L_rt, const float2 *__restrict__ T, int tid) {
    const int N = NCT ? NCT : N_rt;
    const int L = NCT ? LCT : L_rt;
    const int ldz = fused_ldz(N);
    const int Ls = L / R;
    const int nb = N / R;
    const int total = NCOLS * nb;
    int idx = tid;
    int b = idx / NCOLS, col = idx - b * NCOLS;
    constexpr int db = NT / NCOLS, dc = NT - db * NCOLS;
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
        if (col >= NCOLS) { col -= NCOLS; ++b; }
    }
}

template <int SIGN, bool DIF, int NT, int NCOLS>
__device__ __forceinline__ void fused_stage_dispatch(int R, float2 *z, int N, int L, const float2 *T, int tid) {
    switch (R) {
        case 2: fused_stage<2, SIGN, DIF, NT, NCOLS, 0, 0>(z, N, L, T, tid); break;
        case 3: fused_stage<3, SIGN, DIF, NT, NCOLS, 0, 0>(z, N, L, T, tid); break;
        case 4: fused_stage<4, SIGN, DIF, NT, NCOLS, 0, 0>(z, N, L, T, tid); break;
        case 5: fused_stage<5, SIGN, DIF, NT, NCOLS, 0, 0>(z, N, L, T, tid); break;
        case 7: fused_stage<7, SIGN, DIF, NT, NCOLS, 0, 0>(z, N, L, T, tid); break;
        case 8: fused_stage<8, SIGN, DIF, NT, NCOLS, 0, 0>(z, N, L, T, tid); break;
        case 11: fused_stage<11, SIGN, DIF, NT, NCOLS, 0, 0>(z, N, L, T, tid); break;
        case 13: fused_stage<13, SIGN, DIF, NT, NCOLS, 0, 0>(z, N, L, T, tid); break;
        case 16: fused_stage<16, SIGN, DIF, NT, NCOLS, 0, 0>(z, N, L, T, tid); break;
        default: break;
    }
}

// compile-time plan: stage S of length-NCT transform has block length L; TWOFF = offset of its twiddle table
template <int NT, int NCOLS, int NCT, int S, int L, int TWOFF>
struct CtFft {
    static __device__ __forceinline__ void inverse(float2 *Z, const float2 *TW, int tid) {
        if constexpr (S < choose_radices_ct(NCT).n) {
            constexpr int R = choose_radices_ct(NCT).r[S];
            constexpr int Ls = L / R;
            fused_stage<R, +1, true, NT, NCOLS, NCT, L>(Z, NCT, L, TW + TWOFF, tid);
            __syncthreads();
            CtFft<NT, NCOLS, NCT, S + 1, Ls, TWOFF + (Ls > 1 ? L : 0)>::inverse(Z, TW, tid);
        }
    }
    static __device__ __forceinline__ void forward(float2 *Z, const float2 *TW, int tid) {
        if constexpr (S < choose_radices_ct(NCT).n) {
            constexpr int R = choose_radices_ct(NCT).r[S];
            constexpr int Ls = L / R;
            CtFft<NT, NCOLS, NCT, S + 1, Ls, TWOFF + (Ls > 1 ? L : 0)>::forward(Z, TW, tid);
            fused_stage<R, -1, false, NT, NCOLS, NCT, L>(Z, NCT, L, TW + TWOFF, tid);
            __syncthreads();
        }
    }
};

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
};

// stress of NB (point, phi) cells at once: all moduli are requested before the first use (Isotropic3D.cpp:10-27,
// TransverselyIsotropic3D.cpp:10-28, Anisotropic3D.cpp:10-54; no attenuation on this path)
template <int NCOEF, int NB, int NT>
__device__ __forceinline__ void stress_batch(int law, float2 *__restrict__ Z, const float *__restrict__ cf, int total, int ldz, int N,
                                             int tid) {
    const int cs = AX_NPE * ldz;
    const int dp = NT / N, dpos = NT - dp * N;
    int idx = tid;
    int p = idx / N, pos = idx - p * N;
    for (; idx < total; idx += NB * NT) {
        float c[NB][NCOEF];
#pragma unroll
        for (int u = 0; u < NB; ++u)
            if (idx + u * NT < total) {
#pragma unroll
                for (int k = 0; k < NCOEF; ++k) c[u][k] = __ldcs(cf + (size_t)k * total + idx + u * NT);
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
    }
}

// E, P: descriptor and plan of this element (shared memory).  On entry the first gather tile of this element is in
// flight (cp.async); `after_first_sync` runs behind the first barrier (every thread has left the previous element: its
// descriptor slot is free), `after_grad` once U is dead (it starts the next element's gather).
template <bool FLUID, int NT, int NCT, typename GatherFn, typename AfterSyncFn, typename AfterGradFn>
__device__ __forceinline__ void fused_element(const FusedCtx<FLUID> &cx, const ElemDesc &E, const FftPlan &P, int tid,
                                              GatherFn gather, AfterSyncFn after_first_sync, AfterGradFn after_grad) {
    constexpr int NC = FLUID ? 1 : 3, NPAIR = FLUID ? 2 : 3;
    constexpr int US = NC * AX_NPE;                   // row stride of U (odd)
    constexpr int NCOLS = NPAIR * AX_NPE;
    constexpr int NHW = NT / 16;
    constexpr int PP = (AX_NPE + NHW - 1) / NHW;      // point passes per thread
    constexpr int QIT = NT >= 256 ? 4 : 2;   // 16-mode chunks per quad tile (the pointwise term r lives in registers)
    float2 *const U = cx.U, *const TW = cx.TW, *const Z = cx.Z;
    const int hw = tid >> 4, t = tid & 15;
    const int N = NCT ? NCT : E.nr, nu = N / 2, M = nu + 1, Mt = E.mt;
    const int ldz = fused_ldz(N);
    const bool nyq = (N & 1) == 0;
    const bool axial = E.axial != 0, tiso = !FLUID && E.tiso != 0;
    const int law = E.law;
    const long long geom_off = E.geom_off, trig_off = E.trig_off, coef_off = E.coef_off;

    // ------------------------------------------------------------ gather (prefetched) + grad, Mt modes at a time   @phase gather wait + grad
    for (int a0 = 0; a0 < M; a0 += Mt) {
        const int mt = min(Mt, M - a0);
        if (a0) {
            __syncthreads();
            gather(E, a0, mt);
        }
        cp_async_wait_all();
        if (a0 == 0 && t == 0)   // Im(u) of mode 0 is not used (Gradient.cpp:209-224): this thread copied these entries
            for (int row = hw; row < US; row += NHW) U[row].y = 0.f;
        __syncthreads();
        if (a0 == 0) after_first_sync();
#pragma unroll
        for (int pp = 0; pp < PP; ++pp) {
            const int p = pp * NHW + hw;
            if (p < AX_NPE) {
                const int i = p / 5, j = p - 5 * i;
                GCoef gc;
                load_gcoef(gc, axial, i, j);
                const PointGeom g = load_geom(cx.geom, geom_off, p);
                const bool ax0 = axial && i == 0;
                float tr[4] = {0.f, 1.f, 0.f, 1.f};
                if (tiso) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) tr[k] = cx.geom[trig_off + k * AX_NPE + p];
                }
                float2 *zp = Z + p * ldz;
                for (int a = t; a < mt; a += 16) {
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
                        for (int pr = 0; pr < 3; ++pr) zform_store(zp + pr * AX_NPE * ldz, N, alpha, ee[2 * pr], ee[2 * pr + 1]);
                    } else {
                        float2 ee[3];
                        grad_fluid_mm(U + a * US, i, j, gc, g, (float)alpha, ax0, ee);
                        if (dead) ee[0] = ee[1] = ee[2] = czero();
                        zform_store(zp, N, alpha, ee[0], ee[1]);
                        zform_store(zp + AX_NPE * ldz, N, alpha, ee[2], czero());
                    }
                }
            }
        }
    }
    __syncthreads();   // Z complete, U dead
    after_grad();      // @phase next-element gather issue

    // ------------------------------------------------------------ c2r (SolverFFTW_N6::computeC2R, unnormalised, sign +)   @phase c2r
    if constexpr (NCT != 0) {
        CtFft<NT, NCOLS, NCT, 0, NCT, 0>::inverse(Z, TW, tid);
    } else {
        int L = N;
        for (int s = 0; s < P.nstages; ++s) {
            const int R = P.radix[s];
            fused_stage_dispatch<+1, true, NT, NCOLS>(R, Z, N, L, TW + (P.stw_off[s] - P.stw_base), tid);
            L /= R;
            __syncthreads();
        }
    }

    // ------------------------------------------------------------ physical space: stress (+ SLS attenuation)   @phase stress
    {
        const int total = AX_NPE * N;
        const float *cf0 = cx.coef + coef_off;
        bool done = false;
        if constexpr (!FLUID) {
            if (E.att_kind == ATT_NONE) {
                if (law == LAW_ISO) stress_batch<2, 4, NT>(law, Z, cf0, total, ldz, N, tid);
                else if (law == LAW_TI) stress_batch<5, 2, NT>(law, Z, cf0, total, ldz, N, tid);
                else stress_batch<21, 1, NT>(law, Z, cf0, total, ldz, N, tid);
                done = true;
            }
        }
        if (!done) {
            int idx = tid;
            int p = idx / N, pos = idx - p * N;
            const int dp = NT / N, dpos = NT - dp * N;
            const int cs = AX_NPE * ldz;
            for (; idx < total; idx += NT) {
                float2 *zc = Z + p * ldz + pos;
                if constexpr (!FLUID) {
                    const float2 z0 = zc[0], z1 = zc[cs], z2 = zc[2 * cs];
                    float ee[6] = {z0.x, z0.y, z1.x, z1.y, z2.x, z2.y}, s[6];
                    const float *cf = cf0 + idx;   // [k][point][pos] with point * N + pos == idx
                    stress_law<float>(law, ee, s, [&](int k) { return __ldcs(cf + (size_t)k * total); });
                    const int Pn = E.att_kind == ATT_CG4 ? 4 : AX_NPE;
                    const int q = E.att_kind == ATT_CG4 ? cg4_index(p) : p;
                    if (q >= 0) {
                        const float *ap = cx.attpar + E.att_par_off;
                        const float *mod = ap + 3 * E.nsls;
                        float *stt = cx.attstate + E.att_state_off;
                        const size_t cell = (size_t)q * N + pos;
                        const size_t PN = (size_t)Pn * N, sl = 6 * PN;
                        const int nsls = E.nsls;
                        attenuation_cell<float>(
                            nsls, ap, mod[cell], mod[PN + cell], mod[2 * PN + cell], E.do_kappa != 0, ee, s,
                            [&](int k, int c) -> float & { return stt[k * sl + c * PN + cell]; },
                            [&](int c) -> float & { return stt[nsls * sl + c * PN + cell]; });
                    }
                    zc[0] = make_float2(s[0], s[1]);
                    zc[cs] = make_float2(s[2], s[3]);
                    zc[2 * cs] = make_float2(s[4], s[5]);
                } else {
                    const float K = __ldcs(cf0 + idx);   // Acoustic3D.cpp:9-16
                    const float2 a = zc[0], b = zc[cs];
                    zc[0] = cscale(a, K);
                    zc[cs] = make_float2(b.x * K, 0.f);
                }
                p += dp;
                pos += dpos;
                if (pos >= N) { pos -= N; ++p; }
            }
        }
    }
    __syncthreads();

    // ------------------------------------------------------------ r2c (computeR2C; the 1/Nr is applied at load below)   @phase r2c
    if constexpr (NCT != 0) {
        CtFft<NT, NCOLS, NCT, 0, NCT, 0>::forward(Z, TW, tid);
    } else {
        int L = 1;
        for (int s = P.nstages - 1; s >= 0; --s) {
            const int R = P.radix[s];
            L *= R;
            fused_stage_dispatch<-1, false, NT, NCOLS>(R, Z, N, L, TW + (P.stw_off[s] - P.stw_base), tid);
            __syncthreads();
        }
    }

    // ------------------------------------------------------------ quad + scatter, 16 * QIT modes at a time   @phase quad-pre
    const float sc = 1.f / (float)N;   // SolverFFTW_N6::computeR2C scaling (SolverFFTW_N6.cpp:47-48)
    for (int a0 = 0; a0 < M; a0 += 16 * QIT) {
        float2 r[PP][QIT][NC];
        // pointwise half, in place: slot beta <- X, slot N - beta <- Y (beta = 0: the spare slot N); r stays in registers
#pragma unroll
        for (int pp = 0; pp < PP; ++pp) {
            const int p = pp * NHW + hw;
            if (p < AX_NPE) {
                const int i = p / 5;
                const PointGeom g = load_geom(cx.geom, geom_off, p);
                const bool ax0 = axial && i == 0;
                float tr[4] = {0.f, 1.f, 0.f, 1.f};
                if (tiso) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) tr[k] = cx.geom[trig_off + k * AX_NPE + p];
                }
                float2 *zp = Z + p * ldz;
#pragma unroll
                for (int q = 0; q < QIT; ++q) {
                    const int beta = a0 + 16 * q + t;
                    if (beta < M && !(nyq && beta == nu)) {
                        if constexpr (!FLUID) {
                            float2 s[6], X[3], Y[3];
#pragma unroll
                            for (int pr = 0; pr < 3; ++pr)
                                zform_load(zp + pr * AX_NPE * ldz, N, beta, sc, s[2 * pr], s[2 * pr + 1]);
                            if (tiso) rot_rtz_to_spz(s, tr[0], tr[1], tr[2], tr[3]);
                            quad6_pre(s, g, (float)beta, ax0, X, Y, r[pp][q]);
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                zp[c * AX_NPE * ldz + beta] = X[c];
                                zp[c * AX_NPE * ldz + N - beta] = Y[c];
                            }
                        } else {
                            float2 s[3], X, Y, dummy;
                            zform_load(zp, N, beta, sc, s[0], s[1]);
                            zform_load(zp + AX_NPE * ldz, N, beta, sc, s[2], dummy);
                            quad_fluid_pre(s, g, (float)beta, ax0, X, Y, r[pp][q][0]);
                            zp[beta] = X;
                            zp[N - beta] = Y;
                        }
                    }
                }
            }
        }
        __syncthreads();
        // tensor-product half + Point::gatherStiffFromElement (SolidPoint.cpp:197-209)   @phase quad-post + scatter
#pragma unroll
        for (int pp = 0; pp < PP; ++pp) {
            const int p = pp * NHW + hw;
            if (p < AX_NPE) {
                const int i = p / 5, j = p - 5 * i;
                GCoef gc;
                load_gcoef(gc, axial, i, j);
                const int nlive = E.pt_nlive[p];
                float2 *const dst = cx.stiff + (size_t)E.pt_off[p];
                const int st = E.pt_stride[p];
#pragma unroll
                for (int q = 0; q < QIT; ++q) {
                    const int beta = a0 + 16 * q + t;
                    if (beta < M && !(nyq && beta == nu) && beta < nlive) {
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            float2 f = r[pp][q][c];
                            const float2 *zx = Z + (c * AX_NPE + j) * ldz + beta;           // X(k, j), k = 0..4
                            const float2 *zy = Z + (c * AX_NPE + i * 5) * ldz + N - beta;   // Y(i, k)
#pragma unroll
                            for (int k = 0; k < 5; ++k) {
                                f = cfma(gc.gxi_row[k], zx[k * 5 * ldz], f);
                                f = cfma(gc.geta_row[k], zy[k * ldz], f);
                            }
                            if (beta == 0) f.y = 0.f;
                            atomicAdd(dst + (size_t)c * st + beta, make_float2(-f.x, -f.y));   // stiff -= f (RED.ADD.F32x2)
                        }
                    }
                }
            }
        }
    }
    // the barrier after cp_async_wait_all at the top of the next element separates these reads of Z from its grad
}

// ---------------------------------------------------------------- the kernel   @phase kernel loop
// grid: persistent, one CTA per SM; block NT threads.  work[0] = next element index (starts at gridDim.x),
// work[1] = number of CTAs that have finished; the last one re-arms both for the next launch (graph replay).
template <bool FLUID, int NT, int NCT1>
__global__ void __launch_bounds__(NT, (512 / NT) > 0 ? (512 / NT) : 1)
    k_elem3d_fused(const ElemDesc *__restrict__ elems, int nelem, const FftPlan *__restrict__ plans,
                   const float2 *__restrict__ stwpool, const float *__restrict__ geom, const float *__restrict__ coef,
                   const float *__restrict__ attpar, float *__restrict__ attstate, const float2 *__restrict__ displ,
                   float2 *__restrict__ stiff, int u_cap, int tw_cap, unsigned *__restrict__ work) {
    constexpr int NC = FLUID ? 1 : 3;
    constexpr int US = NC * AX_NPE;
    constexpr int NHW = NT / 16;
    constexpr int DESC_W = (int)(sizeof(ElemDesc) / sizeof(int)), PLAN_W = (int)(sizeof(FftPlan) / sizeof(int));
    static_assert(DESC_W <= NT && 64 + PLAN_W <= NT, "descriptor loaders need NT >= descriptor words");
    extern __shared__ float2 smem[];
    __shared__ ElemDesc sE[2];
    __shared__ FftPlan sP[2];
    __shared__ int sIdx[3];   // ring of element indices: current, next, next-next
    const int tid = threadIdx.x;
    const int hw = tid >> 4, t = tid & 15;
    FusedCtx<FLUID> cx{geom, coef, attpar, attstate, displ, stiff, smem, smem + u_cap, smem + u_cap + tw_cap};
    float2 *const U = cx.U;

    // descriptor + plan of element el -> slot s (plain loads; visible after the next barrier)
    auto load_desc = [&](int s, int el) {
        if (tid < DESC_W) reinterpret_cast<int *>(&sE[s])[tid] = reinterpret_cast<const int *>(elems + el)[tid];
        if (tid >= 64 && tid < 64 + PLAN_W) {
            const int pid = elems[el].plan_id;
            reinterpret_cast<int *>(&sP[s])[tid - 64] = reinterpret_cast<const int *>(plans + pid)[tid - 64];
        }
    };
    // Point::scatterDisplToElement (SolidPoint.cpp:175-195) for modes [a0, a0 + mt): half-warp per (component, point)
    auto gather = [&](const ElemDesc &E, int a0, int mt) {
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

    int e = blockIdx.x;
    if (e < nelem) {
        load_desc(0, e);
        if (tid == 0) sIdx[1] = (int)atomicAdd(&work[0], 1u);
        __syncthreads();
        gather(sE[0], 0, min(sE[0].mt, sE[0].nu + 1));
        int tw_plan = -1;
        for (int it = 0, k = 0; e < nelem; it ^= 1, k = (k == 2 ? 0 : k + 1)) {
            const ElemDesc &E = sE[it];
            const FftPlan &P = sP[it];
            const int kn = k == 2 ? 0 : k + 1, knn = kn == 2 ? 0 : kn + 1;   // ring slots of the next two elements
            if (tw_plan != E.plan_id) {   // TW is idle here: the FFT stages of the previous element are barrier-separated
                for (int k = tid; k < P.stw_len; k += NT) cx.TW[k] = stwpool[P.stw_base + k];
                tw_plan = E.plan_id;
            }
            // moduli of this element -> L2 while gather/grad/c2r run
            {
                const int ncoef = FLUID ? 1 : (E.law == LAW_ISO ? 2 : E.law == LAW_TI ? 5 : 21);
                const float *cb = coef + E.coef_off;
                const int nline = (ncoef * AX_NPE * E.nr + 31) / 32;
                for (int k = tid; k < nline; k += NT) prefetch_l2(cb + (size_t)k * 32);
            }
            // next element: descriptor behind the first barrier, displacement (cp.async into the dead U) behind grad
            auto after_first_sync = [&]() {
                const int en = sIdx[kn];   // fetched during the previous element
                if (en < nelem) load_desc(it ^ 1, en);
            };
            auto after_grad = [&]() {
                if (tid == 0) sIdx[knn] = (int)atomicAdd(&work[0], 1u);   // needed one element from now: latency hidden
                if (sIdx[kn] < nelem) {
                    const ElemDesc &En = sE[it ^ 1];
                    gather(En, 0, min(En.mt, En.nu + 1));
                }
            };
            if (NCT1 != 0 && E.nr == NCT1) fused_element<FLUID, NT, NCT1>(cx, E, P, tid, gather, after_first_sync, after_grad);
            else fused_element<FLUID, NT, 0>(cx, E, P, tid, gather, after_first_sync, after_grad);
            e = sIdx[kn];
        }
    }
    // re-arm the work counter once every CTA is done
    if (tid == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&work[1], 1u);
        if (done == gridDim.x - 1) {
            work[0] = gridDim.x;
            work[1] = 0u;
            __threadfence();
        }
    }
}
