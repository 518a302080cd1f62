// cluster.cuh -- one thread-block CLUSTER of 5 CTAs per element: SolidElement::computeStiff / FluidElement::computeStiff
// (SolidElement.cpp:43-65, 404-443; FluidElement.cpp:43-65, 333-355) for elements with 3D material, with the element's
// spectrum spread over the shared memory of the 5 CTAs (distributed shared memory, sm_90+/sm_100a).
//
// Why: the single-CTA kernel (fused.cuh) needs the whole spectrum of an element (75 Z-form columns x Nr) in ONE SM's
// shared memory, which caps it at Nr ~ 340 and at one CTA (16 warps) per SM; larger elements used to take the split
// pipeline (k_grad3d -> k_fft3d_v2 -> k_quad3d) with the spectrum staged in an L2 scratch ring.  Here CTA r of the cluster
// owns GLL row i = r (points p = 5 r .. 5 r + 4): 1/5 of the gathered displacement and 15 of the 75 columns.
//   * c2r, stress (+SLS) and r2c couple only the components of one point: entirely CTA-local;
//   * the 5x5 tensor-product derivative couples points: the eta-derivative (along j) stays inside the row, the
//     xi-derivative (along i) reads the same (component, j, mode) entry of the other 4 CTAs through DSMEM
//     (cluster.map_shared_rank) -- in grad from the gathered displacement U, in quad from the X columns.
// Per-CTA shared memory: TW (stage twiddles) + U [NC*5][Mp] + Z [NPAIR*5][ldz]; Nr = 208 -> 40 KB (5 CTAs of different
// elements per SM, phases decorrelated), Nr = 1008 -> 190 KB.  Four cluster barriers per element.
#pragma once
#include <cooperative_groups.h>

#include "fused.cuh"

#define AX_CL 5   // CTAs per cluster = GLL rows per element

__host__ __device__ constexpr int cl_mp(int M) { return M | 1; }   // column stride of U (float2)

// shared-memory plan of one CTA (float2 units), identical in the 5 CTAs of a cluster so that DSMEM offsets line up
struct ClLayout {
    int tw, u, z, total;
};
__host__ __device__ inline ClLayout cl_layout(bool fluid, int N, int stw_len) {
    const int NC = fluid ? 1 : 3, NPAIR = fluid ? 2 : 3, M = N / 2 + 1;
    ClLayout l;
    l.tw = 0;
    l.u = (stw_len + 1) & ~1;
    l.z = l.u + ((NC * 5 * cl_mp(M) + 1) & ~1);
    l.total = l.z + NPAIR * 5 * fused_ldz(N);
    return l;
}

template <bool FLUID, int NT>
__global__ void __launch_bounds__(NT, NT >= 512 ? 1 : NT >= 256 ? 2 : 4) k_elem3d_cluster(const ElemDesc *__restrict__ elems, const int *__restrict__ list,
                                                      const FftPlan *__restrict__ plans, const float2 *__restrict__ stwpool,
                                                      const float *__restrict__ geom, const float *__restrict__ coef,
                                                      const float *__restrict__ attpar, float *__restrict__ attstate,
                                                      const float2 *__restrict__ displ, float2 *__restrict__ stiff) {
    namespace cg = cooperative_groups;
    constexpr int NC = FLUID ? 1 : 3, NPAIR = FLUID ? 2 : 3, NCOLS = NPAIR * 5;
    constexpr int PLAN_W = (int)(sizeof(FftPlan) / sizeof(int));
    static_assert(PLAN_W <= NT, "plan loader");
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) float2 smem[];
    __shared__ FftPlan sP;
    __shared__ float sgeom[9 * 5];   // [5 geometry + 4 trig][j] of this row
    const int tid = threadIdx.x;
    const int row = (int)cluster.block_rank();                 // GLL row i of this CTA
    const ElemDesc &E = elems[list[blockIdx.x / AX_CL]];
    if (tid < PLAN_W) reinterpret_cast<int *>(&sP)[tid] = reinterpret_cast<const int *>(plans + E.plan_id)[tid];
    const int N = E.nr, nu = N / 2, M = nu + 1, Mp = cl_mp(M), ldz = fused_ldz(N);
    const bool nyq = (N & 1) == 0, axial = E.axial != 0, tiso = !FLUID && E.tiso != 0;
    const bool ax0 = axial && row == 0;
    if (tid < 45) {
        const int k = tid / 5, j = tid - 5 * k;
        float v = (k == 6 || k == 8) ? 1.f : 0.f;             // identity rotation when the element has none
        if (k < 5) v = geom[E.geom_off + k * AX_NPE + row * 5 + j];
        else if (tiso) v = geom[E.trig_off + (k - 5) * AX_NPE + row * 5 + j];
        sgeom[k * 5 + j] = v;
    }
    __syncthreads();
    const FftPlan &P = sP;
    const ClLayout lay = cl_layout(FLUID, N, P.stw_len);
    float2 *const TW = smem + lay.tw, *const U = smem + lay.u, *const Z = smem + lay.z;
    for (int k = tid; k < P.stw_len; k += NT) TW[k] = stwpool[P.stw_base + k];
    {   // moduli (and SLS state) of this row are first touched two phases from now: pull their lines towards L2 meanwhile
        const int ncoef = FLUID ? 1 : (E.law == LAW_ISO ? 2 : E.law == LAW_TI ? 5 : 21);
        const int lines = (5 * N * 4 + 127) / 128;
        for (int q = tid; q < ncoef * lines; q += NT) {
            const int kc = q / lines, l = q - kc * lines;
            prefetch_l2(coef + E.coef_off + ((size_t)kc * AX_NPE + row * 5) * N + (size_t)l * 32);
        }
    }

    // ------------------------------------------------------------ gather this row (Point::scatterDisplToElement, SolidPoint.cpp:175-195)
#pragma unroll
    for (int cj = 0; cj < NC * 5; ++cj) {
        const int c = cj / 5, j = cj - 5 * c, p = row * 5 + j;
        const float2 *src = displ + (size_t)E.pt_off[p] + (size_t)c * E.pt_stride[p];
        const int nlive = E.pt_nlive[p];
        float2 *dst = U + cj * Mp;
        for (int a = tid; a < M; a += NT) {
            float2 v = a < nlive ? src[a] : czero();
            if (a == 0) v.y = 0.f;                             // Im(u) of mode 0 is not used (Gradient.cpp:209-224)
            dst[a] = v;
        }
    }
    cluster.sync();   // every row of U is in place

    // ------------------------------------------------------------ grad (Gradient::computeGrad6 / computeGrad, Gradient.cpp:26-57, 206-265)
    const float2 *Uk[AX_CL];
#pragma unroll
    for (int k = 0; k < AX_CL; ++k) Uk[k] = k == row ? U : cluster.map_shared_rank(U, k);
    for (int j = 0; j < 5; ++j) {
        GCoef gc;
        load_gcoef(gc, axial, row, j);
        PointGeom g;
        g.dsdxii = sgeom[0 * 5 + j]; g.dsdeta = sgeom[1 * 5 + j]; g.dzdxii = sgeom[2 * 5 + j]; g.dzdeta = sgeom[3 * 5 + j]; g.inv_s = sgeom[4 * 5 + j];
        const float tr0 = sgeom[5 * 5 + j], tr1 = sgeom[6 * 5 + j], tr2 = sgeom[7 * 5 + j], tr3 = sgeom[8 * 5 + j];
        for (int a = tid; a < M; a += NT) {
            const float alpha = (float)a;
            const bool dead = nyq && a == nu;
            float2 GU[NC], UG[NC], u[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                float2 s = czero(), t = czero();
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    s = cfma(gc.gxi_col[k], Uk[k][(c * 5 + j) * Mp + a], s);   // (Gxi^T u)(i, j): rows k live in CTA k
                    t = cfma(gc.geta_col[k], U[(c * 5 + k) * Mp + a], t);      // (u Geta)(i, j): this row
                }
                GU[c] = s;
                UG[c] = t;
                u[c] = U[(c * 5 + j) * Mp + a];
            }
            if constexpr (!FLUID) {
                float2 ds[3], dz[3], e[6];
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
                if (ax0) {   // L'Hopital rows on the axis (Gradient.cpp:221-224, 245-254)
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
                if (dead) {
#pragma unroll
                    for (int c = 0; c < 6; ++c) e[c] = czero();
                }
                if (tiso) rot_spz_to_rtz(e, tr0, tr1, tr2, tr3);
#pragma unroll
                for (int pr = 0; pr < 3; ++pr) zform_store(Z + (pr * 5 + j) * ldz, N, a, e[2 * pr], e[2 * pr + 1]);
            } else {
                float2 e[3];
                const float2 v = mul_ialpha(u[0], alpha);
                e[0] = cfma(g.dzdeta, GU[0], cscale(UG[0], g.dzdxii));
                e[1] = cscale(v, g.inv_s);
                e[2] = cfma(g.dsdeta, GU[0], cscale(UG[0], g.dsdxii));
                if (ax0) e[1] = cfma(g.dzdeta, mul_ialpha(GU[0], alpha), e[1]);
                if (dead) e[0] = e[1] = e[2] = czero();
                zform_store(Z + j * ldz, N, a, e[0], e[1]);
                zform_store(Z + (5 + j) * ldz, N, a, e[2], czero());
            }
        }
    }
    cluster.sync();   // Z of this row complete; nobody reads anybody's U any more (it is reused for r below)

    // ------------------------------------------------------------ c2r -> stress (+ SLS) -> r2c on this row's 15 columns
    {
        int L = N;
        for (int s = 0; s < P.nstages; ++s) {
            const int R = P.radix[s];
            fused_stage_dispatch<+1, true, NT, NCOLS>(R, Z, N, L, TW + (P.stw_off[s] - P.stw_base), tid);
            L /= R;
            __syncthreads();
        }
    }
    physical_space<FLUID, NT>(E, coef, attpar, attstate, Z, N, ldz, row * 5, 5, tid);
    __syncthreads();
    {
        int L = 1;
        for (int s = P.nstages - 1; s >= 0; --s) {
            const int R = P.radix[s];
            L *= R;
            fused_stage_dispatch<-1, false, NT, NCOLS>(R, Z, N, L, TW + (P.stw_off[s] - P.stw_base), tid);
            __syncthreads();
        }
    }

    // ------------------------------------------------------------ quad, pointwise half (Gradient::computeQuad6 / computeQuad), in place:
    // column c of point j: slot beta <- X_c, slot N - beta <- Y_c (beta = 0: the spare slot N); r_c -> U
    const float sc = 1.f / (float)N;   // SolverFFTW_N6::computeR2C scaling (SolverFFTW_N6.cpp:47-48)
    for (int j = 0; j < 5; ++j) {
        PointGeom g;
        g.dsdxii = sgeom[0 * 5 + j]; g.dsdeta = sgeom[1 * 5 + j]; g.dzdxii = sgeom[2 * 5 + j]; g.dzdeta = sgeom[3 * 5 + j]; g.inv_s = sgeom[4 * 5 + j];
        const float tr0 = sgeom[5 * 5 + j], tr1 = sgeom[6 * 5 + j], tr2 = sgeom[7 * 5 + j], tr3 = sgeom[8 * 5 + j];
        float2 *zp = Z + j * ldz;
        for (int beta = tid; beta < M; beta += NT) {
            if (nyq && beta == nu) continue;
            if constexpr (!FLUID) {
                float2 s[6], X[3], Y[3], r[3];
#pragma unroll
                for (int pr = 0; pr < 3; ++pr) zform_load(zp + pr * 5 * ldz, N, beta, sc, s[2 * pr], s[2 * pr + 1]);
                if (tiso) rot_rtz_to_spz(s, tr0, tr1, tr2, tr3);
                quad6_pre(s, g, (float)beta, ax0, X, Y, r);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    zp[c * 5 * ldz + beta] = X[c];
                    zp[c * 5 * ldz + N - beta] = Y[c];
                    U[(c * 5 + j) * Mp + beta] = r[c];
                }
            } else {
                float2 s[3], X, Y, r, dummy;
                zform_load(zp, N, beta, sc, s[0], s[1]);
                zform_load(zp + 5 * ldz, N, beta, sc, s[2], dummy);
                quad_fluid_pre(s, g, (float)beta, ax0, X, Y, r);
                zp[beta] = X;
                zp[N - beta] = Y;
                U[j * Mp + beta] = r;
            }
        }
    }
    cluster.sync();   // X of every row is in place

    // ------------------------------------------------------------ quad, tensor-product half + Point::gatherStiffFromElement (SolidPoint.cpp:197-209)
    const float2 *Zk[AX_CL];
#pragma unroll
    for (int k = 0; k < AX_CL; ++k) Zk[k] = k == row ? Z : cluster.map_shared_rank(Z, k);
    for (int j = 0; j < 5; ++j) {
        GCoef gc;
        load_gcoef(gc, axial, row, j);
        const int p = row * 5 + j;
        const int nlive = min(E.pt_nlive[p], M);
        float2 *const dst = stiff + (size_t)E.pt_off[p];
        const int st = E.pt_stride[p];
        for (int beta = tid; beta < nlive; beta += NT) {
            if (nyq && beta == nu) continue;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                float2 f = U[(c * 5 + j) * Mp + beta];
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    f = cfma(gc.gxi_row[k], Zk[k][(c * 5 + j) * ldz + beta], f);        // X(k, j): row k lives in CTA k
                    f = cfma(gc.geta_row[k], Z[(c * 5 + k) * ldz + N - beta], f);       // Y(i, k): this row
                }
                if (beta == 0) f.y = 0.f;
                atomicAdd(dst + (size_t)c * st + beta, make_float2(-f.x, -f.y));        // stiff -= f (RED.ADD.F32x2)
            }
        }
    }
    cluster.sync();   // no CTA may exit while a peer still reads its X columns
}
