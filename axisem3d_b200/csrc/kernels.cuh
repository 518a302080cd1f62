// kernels.cuh -- sm_100a kernels of the per-timestep hot path (Newmark.cpp:47-93).
//
//   k_newmark_{solid,fluid}   Domain::updateNewmark      (SolidPoint.cpp:23-38, FluidPoint.cpp:23-44)
//   k_mass3d                  Mass3D::computeAccel        (Mass3D.cpp:13-57)
//   k_source                  Domain::applySource         (SourceTerm.cpp:29-35)
//   k_elem1d<FLUID>           Element::computeStiff, 1D material (no FFT): gather -> grad -> [rotate] ->
//                             stress(+SLS) -> [rotate^-1] -> quad -> scatter, one CTA per (element, mode tile)
//   k_grad3d / k_fft3d / k_quad3d  the same path for 3D material, split at the two points where the data
//                             dependency changes axis (points <-> modes); the spectrum between them lives in an
//                             L2-resident scratch ring ("Z-form": two real columns per complex column)
//   k_sf_couple               Domain::coupleSolidFluid    (SFCoupling1D.cpp:9-19)
//   k_pack / k_unpack_add     Domain::assembleStiff       (SolidPoint.cpp:163-173)
//   k_check_finite            Domain::checkStability      (Domain.cpp:237-275)
#pragma once
#include "elem.cuh"

struct ElemDesc {
    int nr, nu, nyq, axial;
    int tiso, law, att_kind, nsls;
    int do_kappa, plan_id, ppb, is3d;
    int mt, bucket;            // fused kernel (fused.cuh): modes per gather tile; launch bucket (-1: split pipeline)
    unsigned pt_off[AX_NPE];   // offset of the point's block in the solid / fluid field array (float2 units)
    int pt_stride[AX_NPE];     // Nu_p + 1 (component stride inside the block)
    int pt_nlive[AX_NPE];      // rows [0, nlive) are gathered / scattered (Nu_p - nyq_p + 1)
    int pt_nw[AX_NPE];         // in-kernel Newmark (fused.cuh): -1, or solid point index | (number of fused elements touching it << 24)
    long long geom_off;        // float  [5][25]: dsdxii, dsdeta, dzdxii, dzdeta, inv_s
    long long trig_off;        // float  [4][25]: sin t, cos t, sin 2t, cos 2t
    long long coef_off;        // float  1D: [ncoef][25]      3D: [ncoef][25][Nr] (digit-reversed phi)
    long long att_par_off;     // float  abg[3][nsls] then {3 dkappa, dmu, 2 dmu}: 1D [3][P], 3D [3][P][Nr]
    long long att_state_off;   // 1D: float2 [nsls+1][6][P][M]   3D: float [nsls+1][6][P][Nr]; slot nsls = stressR
    long long scratch_off;     // float2 [NPAIR][25][Nr] inside the scratch ring
    long long prt_off;         // float  particle relabelling X: 1D [4][25], 3D [4][25][Nr] (digit-reversed phi), in the moduli pool
    int prt;                   // 1: the element carries a PRT (9-component path; trig_off is valid for fluid elements too)
    int ng;                    // fused kernel: number of row-group passes (1, 2, 3 or 5)
    int zoff, twoff;           // fused kernel: float2 offsets of the Z-form columns and of the twiddle tables in the tile region
    int bnd, pad_;             // 1: the element touches a point shared with another rank (boundary element: first in the work queue)
};

struct PointTab {              // one per field family (solid: ncomp = 3, fluid: ncomp = 1)
    const unsigned *off;       // [npoint] block offset (float2 units)
    const int *nu;             // [npoint]
    const int *nr;             // [npoint]
    const unsigned char *flags;// bit0 axial, bit1 fluidSurf, bit2 mass3D
    const float *invmass;      // [npoint] (1.0 for Mass3D points: k_mass3d already applied it)
    const int *row_point;      // [nrows] point index of every (point, mode) row
    const int *row_start;      // [npoint] first row of the point
    int nrows;
};

// ------------------------------------------------------------------------------------ masks
// SolidPoint::maskField (SolidPoint.cpp:216-238) on the 3 components of one mode
__device__ __forceinline__ void mask_solid(float2 (&f)[3], int alpha, int nu, bool axial, bool nyq) {
    if (alpha == 0) { f[0].y = 0.f; f[1].y = 0.f; f[2].y = 0.f; }
    if (axial) {
        if (alpha == 0) { f[0] = czero(); f[1] = czero(); }
        else if (alpha == 1) {
            float2 s0 = f[0], s1 = f[1];
            f[0] = make_float2(0.5f * (s0.x + s1.y), 0.5f * (s0.y - s1.x));   // half (s0 - i s1)
            f[1] = make_float2(0.5f * (s1.x - s0.y), 0.5f * (s1.y + s0.x));   // half (s1 + i s0)
            f[2] = czero();
        } else { f[0] = czero(); f[1] = czero(); f[2] = czero(); }
    }
    if (nyq && alpha == nu) { f[0] = czero(); f[1] = czero(); f[2] = czero(); }
}
// FluidPoint::maskField (FluidPoint.cpp:197-207)
__device__ __forceinline__ void mask_fluid(float2 &f, int alpha, int nu, bool axial, bool nyq) {
    if (alpha == 0) f.y = 0.f;
    if (axial && alpha > 0) f = czero();
    if (nyq && alpha == nu) f = czero();
}

// ------------------------------------------------------------------------------------ Newmark
// SolidPoint::updateNewmark (SolidPoint.cpp:31-36) on one complex entry, f = masked acceleration.  One definition for
// the stand-alone kernels and the in-kernel Newmark warps of fused.cuh, so that both round identically.
__device__ __forceinline__ void newmark_entry(float2 f, float2 a_old, float2 &v, float2 &u, float half_dt, float dt, float half_dt_dt) {
    v.x = fmaf(half_dt, a_old.x + f.x, v.x);
    v.y = fmaf(half_dt, a_old.y + f.y, v.y);
    u.x += fmaf(dt, v.x, half_dt_dt * f.x);
    u.y += fmaf(dt, v.y, half_dt_dt * f.y);
}

// pt.row_point / row_start describe the rows this launch covers (all points, or the "special" ones the in-kernel Newmark
// warps do not own); everything else is indexed by point.
__global__ void __launch_bounds__(256) k_newmark_solid(PointTab pt, float2 *__restrict__ displ, float2 *__restrict__ veloc,
                                                       float2 *__restrict__ accel, float2 *__restrict__ stiff,
                                                       float half_dt, float dt, float half_dt_dt) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= pt.nrows) return;
    const int p = pt.row_point[r];
    const int alpha = r - pt.row_start[p];
    const int nu = pt.nu[p];
    const int st = nu + 1;
    const unsigned char fl = pt.flags[p];
    const bool axial = fl & 1, nyq = (pt.nr[p] & 1) == 0;
    const size_t base = (size_t)pt.off[p] + alpha;
    float2 f[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) f[c] = stiff[base + (size_t)c * st];
    mask_solid(f, alpha, nu, axial, nyq);
    const float im = pt.invmass[p];
#pragma unroll
    for (int c = 0; c < 3; ++c) f[c] = make_float2(f[c].x * im, f[c].y * im);
    mask_solid(f, alpha, nu, axial, nyq);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const size_t i = base + (size_t)c * st;
        float2 a_old = accel[i], v = veloc[i], u = displ[i];
        newmark_entry(f[c], a_old, v, u, half_dt, dt, half_dt_dt);
        veloc[i] = v;
        accel[i] = f[c];
        displ[i] = u;
        stiff[i] = czero();
    }
}

__global__ void __launch_bounds__(256) k_newmark_fluid(PointTab pt, float2 *__restrict__ displ, float2 *__restrict__ veloc,
                                                       float2 *__restrict__ accel, float2 *__restrict__ stiff,
                                                       float half_dt, float dt, float half_dt_dt) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= pt.nrows) return;
    const int p = pt.row_point[r];
    const int alpha = r - pt.row_start[p];
    const int nu = pt.nu[p];
    const unsigned char fl = pt.flags[p];
    const size_t i = (size_t)pt.off[p] + alpha;
    if (fl & 2) {   // surface fluid point: pinned to zero (FluidPoint.cpp:25-28)
        displ[i] = veloc[i] = accel[i] = stiff[i] = czero();
        return;
    }
    const bool axial = fl & 1, nyq = (pt.nr[p] & 1) == 0;
    float2 f = stiff[i];
    mask_fluid(f, alpha, nu, axial, nyq);
    const float im = pt.invmass[p];
    f = make_float2(f.x * im, f.y * im);
    mask_fluid(f, alpha, nu, axial, nyq);
    float2 a_old = accel[i], v = veloc[i], u = displ[i];
    newmark_entry(f, a_old, v, u, half_dt, dt, half_dt_dt);
    veloc[i] = v;
    accel[i] = f;
    displ[i] = u;
    stiff[i] = czero();
}

// ------------------------------------------------------------------------------------ point FFT helpers
// spectrum X[k], k <= Nu (Hermitian) for up to two real columns -> complex column z[n], n < N:
// Z[k] = A[k] + i B[k], Z[N-k] = conj(A[k]) + i conj(B[k])
__device__ __forceinline__ void zform_store(float2 *z, int N, int k, float2 a, float2 b) {
    z[k] = make_float2(a.x - b.y, a.y + b.x);
    if (k >= 1 && 2 * k < N) z[N - k] = make_float2(a.x + b.y, b.x - a.y);
}
__device__ __forceinline__ void zform_load(const float2 *z, int N, int k, float scale, float2 &a, float2 &b) {
    float2 zk = z[k];
    if (k >= 1 && 2 * k < N) {
        float2 w = z[N - k];
        a = make_float2(0.5f * scale * (zk.x + w.x), 0.5f * scale * (zk.y - w.y));
        b = make_float2(0.5f * scale * (zk.y + w.y), 0.5f * scale * (w.x - zk.x));
    } else {
        a = make_float2(scale * zk.x, 0.f);
        b = make_float2(scale * zk.y, 0.f);
    }
}

// Mass3D::computeAccel (Mass3D.cpp:13-57) for the points listed in `list`: mask -> c2r -> * invMass(phi) -> r2c/Nr.
// One CTA per point; NCOMP = 3 (solid: columns s,phi | z,0) or 1 (fluid).  Result overwrites stiff.
struct Mass3DItem {
    int point;         // index into the point table
    int plan_id;
    long long im_off;  // float [Nr] inverse mass, digit-reversed phi
    long long ns_off;  // -1, or MassOcean3D: float [3][Nr] unit normal * sqrt(m_ocean / (m (m + m_ocean))), digit-reversed phi
};
template <int NCOMP>
__global__ void __launch_bounds__(128) k_mass3d(PointTab pt, const Mass3DItem *__restrict__ items,
                                                const FftPlan *__restrict__ plans, const float2 *__restrict__ twpool,
                                                const float *__restrict__ impool, float2 *__restrict__ stiff) {
    extern __shared__ float2 smem[];
    const Mass3DItem it = items[blockIdx.x];
    const FftPlan pl = plans[it.plan_id];
    const int N = pl.N, nu = pt.nu[it.point], st = nu + 1;
    const bool axial = pt.flags[it.point] & 1, nyq = (N & 1) == 0;
    constexpr int NCOL = NCOMP == 3 ? 2 : 1;
    float2 *z = smem;               // [NCOL][N]
    float2 *tw = smem + NCOL * N;   // [N]
    const size_t base = pt.off[it.point];
    for (int k = threadIdx.x; k < N; k += blockDim.x) tw[k] = twpool[pl.tw_off + k];
    for (int k = threadIdx.x; k <= nu; k += blockDim.x) {
        if (NCOMP == 3) {
            float2 f[3] = {stiff[base + k], stiff[base + st + k], stiff[base + 2 * st + k]};
            mask_solid(f, k, nu, axial, nyq);
            zform_store(z, N, k, f[0], f[1]);
            zform_store(z + N, N, k, f[2], czero());
        } else {
            float2 f = stiff[base + k];
            mask_fluid(f, k, nu, axial, nyq);
            zform_store(z, N, k, f, czero());
        }
    }
    __syncthreads();
    fft_inverse_dif(pl, z, N, NCOL, tw, threadIdx.x, blockDim.x);
    const float *im = impool + it.im_off;
    if (NCOMP == 3 && it.ns_off >= 0) {   // MassOcean3D::computeAccel (MassOcean3D.cpp:18-49): a = f / m - (f . n') n'
        const float *ns = impool + it.ns_off;
        for (int pos = threadIdx.x; pos < N; pos += blockDim.x) {
            const float2 z0 = z[pos], z1 = z[N + pos];
            const float n0 = ns[pos], n1 = ns[N + pos], n2 = ns[2 * N + pos], m = im[pos];
            const float fn = z0.x * n0 + z0.y * n1 + z1.x * n2;
            z[pos] = make_float2(z0.x * m - fn * n0, z0.y * m - fn * n1);
            z[N + pos] = make_float2(z1.x * m - fn * n2, 0.f);
        }
    } else {
        for (int idx = threadIdx.x; idx < NCOL * N; idx += blockDim.x) {
            const int pos = idx % N;
            z[idx] = cscale(z[idx], im[pos]);
        }
    }
    __syncthreads();
    fft_forward_dit(pl, z, N, NCOL, tw, threadIdx.x, blockDim.x);
    const float sc = 1.f / (float)N;
    for (int k = threadIdx.x; k <= nu; k += blockDim.x) {
        float2 a, b;
        zform_load(z, N, k, sc, a, b);
        stiff[base + k] = a;
        if (NCOMP == 3) {
            stiff[base + st + k] = b;
            float2 c, d;
            zform_load(z + N, N, k, sc, c, d);
            stiff[base + 2 * st + k] = c;
        }
    }
}

// MassOcean1D::computeAccel (MassOcean1D.cpp:15-28) for the solid points with an axisymmetric ocean load: mask, rotate (s, z)
// to (normal, tangent), scale by 1 / (m + m_ocean) and 1 / m, rotate back; phi component / m.  Overwrites stiff; the Newmark
// kernel then sees inverse mass 1 (like the Mass3D points).  Thread = (item, mode).
struct Ocean1DItem {
    int point;
    float imZ, imR, sint, cost;
};
__global__ void k_mass_ocean1d(PointTab pt, int nitems, const Ocean1DItem *__restrict__ items, int max_m, float2 *__restrict__ stiff) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = r / max_m, alpha = r - k * max_m;
    if (k >= nitems) return;
    const Ocean1DItem it = items[k];
    const int nu = pt.nu[it.point], st = nu + 1;
    if (alpha > nu) return;
    const bool axial = pt.flags[it.point] & 1, nyq = (pt.nr[it.point] & 1) == 0;
    const size_t base = (size_t)pt.off[it.point] + alpha;
    float2 f[3] = {stiff[base], stiff[base + st], stiff[base + 2 * st]};
    mask_solid(f, alpha, nu, axial, nyq);
    float2 Z = cadd(cscale(f[0], it.sint), cscale(f[2], it.cost));
    float2 R = csub(cscale(f[0], it.cost), cscale(f[2], it.sint));
    Z = cscale(Z, it.imZ);
    R = cscale(R, it.imR);
    stiff[base] = cadd(cscale(Z, it.sint), cscale(R, it.cost));
    stiff[base + 2 * st] = csub(cscale(Z, it.cost), cscale(R, it.sint));
    stiff[base + st] = cscale(f[1], it.imR);
}

// ------------------------------------------------------------------------------------ source
// stf is read from device memory (written by a 4-byte H2D copy per step) so that the launch can be replayed
// from a CUDA graph with a different source factor every step.
__global__ void k_source(int n, const unsigned *__restrict__ off, const float2 *__restrict__ val,
                         const float *__restrict__ stf_ptr, float2 *__restrict__ stiff) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float stf = *stf_ptr;
    float2 v = val[i];
    atomicAdd(&stiff[off[i]], make_float2(v.x * stf, v.y * stf));
}

// ------------------------------------------------------------------------------------ element helpers
__device__ __forceinline__ PointGeom load_geom(const float *__restrict__ geom, long long off, int p) {
    PointGeom g;
    g.dsdxii = geom[off + 0 * AX_NPE + p];
    g.dsdeta = geom[off + 1 * AX_NPE + p];
    g.dzdxii = geom[off + 2 * AX_NPE + p];
    g.dzdeta = geom[off + 3 * AX_NPE + p];
    g.inv_s = geom[off + 4 * AX_NPE + p];
    return g;
}

// gather of SolidPoint::scatterDisplToElement (SolidPoint.cpp:175-195) for one (alpha, point)
template <int NC>
__device__ __forceinline__ void gather_tile(const ElemDesc &E, const float2 *__restrict__ displ, float2 *sU, int t, int p,
                                            int alpha) {
    const bool live = alpha < E.pt_nlive[p];
    const size_t base = (size_t)E.pt_off[p] + alpha;
    const int st = E.pt_stride[p];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        float2 u = live ? displ[base + (size_t)c * st] : czero();
        if (alpha == 0) u.y = 0.f;
        sU[(c * AX_NPE + p) * AX_TILE + t] = u;
    }
}

// scatter of SolidPoint::gatherStiffFromElement (SolidPoint.cpp:197-209): stiff -= f on live rows
__device__ __forceinline__ void scatter_sub(float2 *__restrict__ stiff, size_t i, float2 f) {
    atomicAdd(&stiff[i], make_float2(-f.x, -f.y));   // sm_90+ vector atomic (RED.E.ADD.F32x2)
}

static __device__ __forceinline__ int cg4_index(int p) {   // (1,1),(1,3),(3,1),(3,3) -> 0..3, else -1
    return p == 6 ? 0 : p == 8 ? 1 : p == 16 ? 2 : p == 18 ? 3 : -1;
}

// ------------------------------------------------------------------------------------ 1D-material element kernel
// grid: one CTA per work item (element, first mode of the tile); block: AX_TILE x 25 threads, lane = mode.
#ifndef AX_ELEM1D_MIN_CTAS
#define AX_ELEM1D_MIN_CTAS 3   // resident 400-thread CTAs per SM the register allocation must allow for the solid instance; B200, cfg1
                               // (profiles/microbench/elem1d_ab.sh): 1 -> 0.0472, 2 -> 0.0403, 3 -> 0.0371, 4 -> 0.0437 ms per step
#endif
// PRT: instance that also handles elements with particle relabelling (only launched when the class has any)
template <bool FLUID, bool PRT = false>
__global__ void __launch_bounds__(AX_TILE *AX_NPE, FLUID ? 4 : AX_ELEM1D_MIN_CTAS) k_elem1d(const ElemDesc *__restrict__ elems, const int *__restrict__ w_elem,
                                                            const int *__restrict__ w_a0, const float *__restrict__ geom,
                                                            const float *__restrict__ coef, const float *__restrict__ attpar,
                                                            float2 *__restrict__ attstate, const float2 *__restrict__ displ,
                                                            float2 *__restrict__ stiff) {
    constexpr int NC = FLUID ? 1 : 3;
    __shared__ float2 sU[NC * AX_NPE * AX_TILE];
    __shared__ float2 sX[NC * AX_NPE * AX_TILE];
    __shared__ float2 sY[NC * AX_NPE * AX_TILE];
    const int t = threadIdx.x % AX_TILE, p = threadIdx.x / AX_TILE;
    const int i = p / 5, j = p % 5;
    // work item: (element, first mode of a 16-mode tile), or a pack of -a0 consecutive small elements of equal M sharing
    // the 16 lanes (lane = sub-element * Mpad + mode); all tiles below are indexed by lane, so a pack needs no other change
    const int e0 = w_elem[blockIdx.x], a0 = w_a0[blockIdx.x];
    int sub = 0, alpha = a0 + t;
    bool lane_on = true;
    if (a0 < 0) {
        const int Mpad = elems[e0].nu + 1 <= 4 ? 4 : 8;
        sub = t / Mpad;
        alpha = t - sub * Mpad;
        lane_on = sub < -a0;
        if (!lane_on) sub = 0;
    }
    const ElemDesc &E = elems[e0 + sub];
    const int M = E.nu + 1;
    const bool active = lane_on && alpha < M;
    gather_tile<NC>(E, displ, sU, t, p, active ? alpha : (1 << 30));
    GCoef gc;
    load_gcoef(gc, E.axial, i, j);
    const PointGeom g = load_geom(geom, E.geom_off, p);
    const bool ax0 = E.axial && i == 0;
    const bool dead = !active || (E.nyq && alpha == E.nu);
    __syncthreads();
    float2 r[NC];
    if constexpr (!FLUID) {
        float2 e[6], s[6], X[3], Y[3];
        float tr[4] = {0.f, 1.f, 0.f, 1.f};
        if (E.tiso) {
#pragma unroll
            for (int k = 0; k < 4; ++k) tr[k] = geom[E.trig_off + k * AX_NPE + p];
        }
        float px[4] = {1.f, 0.f, 0.f, 1.f};
        if (PRT && E.prt) {   // SolidElement.cpp:405-413: grad9 -> rotate -> sphericalToUndulated (PRT_1D, Fourier space)
#pragma unroll
            for (int k = 0; k < 4; ++k) px[k] = coef[E.prt_off + k * AX_NPE + p];
            float2 e9[9];
            grad9_point(sU, AX_TILE, t, i, j, gc, g, (float)alpha, ax0, e9);
            if (dead) {
#pragma unroll
                for (int c = 0; c < 9; ++c) e9[c] = czero();
            }
            rot9(e9, tr[0], tr[1], tr[2], tr[3], false);
            prt_s2u_solid(e9, px, e);
        } else {
            grad6_point(sU, AX_TILE, t, i, j, gc, g, (float)alpha, ax0, e);
            if (dead) {
#pragma unroll
                for (int c = 0; c < 6; ++c) e[c] = czero();
            }
            if (E.tiso) rot_spz_to_rtz(e, tr[0], tr[1], tr[2], tr[3]);
        }
        const float *cf = coef + E.coef_off + p;
        stress_law<float2>(E.law, e, s, [&](int k) { return cf[k * AX_NPE]; });
        if (E.att_kind != ATT_NONE && active) {
            const int P = E.att_kind == ATT_CG4 ? 4 : AX_NPE;
            const int q = E.att_kind == ATT_CG4 ? cg4_index(p) : p;
            if (q >= 0) {
                const float *ap = attpar + E.att_par_off;
                const float *mod = ap + 3 * E.nsls;
                float2 *stt = attstate + E.att_state_off;
                const size_t cell = (size_t)q * M + alpha;
                const size_t sl = (size_t)6 * P * M;
                attenuation_cell<float2>(
                    E.nsls, ap, mod[q], mod[P + q], mod[2 * P + q], E.do_kappa != 0, e, s,
                    [&](int k, int c) -> float2 & { return stt[k * sl + (size_t)c * P * M + cell]; },
                    [&](int c) -> float2 & { return stt[E.nsls * sl + (size_t)c * P * M + cell]; });
            }
        }
        if (PRT && E.prt) {   // SolidElement.cpp:424-432: undulatedToSpherical -> rotate back -> quad9
            float2 s9[9];
            prt_u2s_solid(s, px, s9);
            rot9(s9, tr[0], tr[1], tr[2], tr[3], true);
            quad9_pre(s9, g, (float)alpha, ax0, X, Y, r);
        } else {
            if (E.tiso) rot_rtz_to_spz(s, tr[0], tr[1], tr[2], tr[3]);
            quad6_pre(s, g, (float)alpha, ax0, X, Y, r);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            sX[(c * AX_NPE + p) * AX_TILE + t] = X[c];
            sY[(c * AX_NPE + p) * AX_TILE + t] = Y[c];
        }
    } else {
        float2 e[3], s[3], X, Y;
        grad_fluid_point(sU, AX_TILE, t, i, j, gc, g, (float)alpha, ax0, e);
        const float K = dead ? 0.f : coef[E.coef_off + p];
        float px[4] = {1.f, 0.f, 0.f, 1.f}, s1 = 0.f, c1 = 1.f;
        if (PRT && E.prt) {   // FluidElement.cpp:335-343: rotate -> sphericalToUndulated
#pragma unroll
            for (int k = 0; k < 4; ++k) px[k] = coef[E.prt_off + k * AX_NPE + p];
            s1 = geom[E.trig_off + p];
            c1 = geom[E.trig_off + AX_NPE + p];
            rot3_fluid(e, s1, c1, false);
            prt_s2u_fluid(e, px);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) s[c] = cscale(e[c], K);   // Acoustic1D.cpp:8-16
        if (PRT && E.prt) {   // FluidElement.cpp:345-352
            prt_u2s_fluid(s, px);
            rot3_fluid(s, s1, c1, true);
        }
        quad_fluid_pre(s, g, (float)alpha, ax0, X, Y, r[0]);
        sX[p * AX_TILE + t] = X;
        sY[p * AX_TILE + t] = Y;
    }
    __syncthreads();
    if (dead || alpha >= E.pt_nlive[p]) return;
    const size_t base = (size_t)E.pt_off[p] + alpha;
    const int st = E.pt_stride[p];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        float2 f = quad_post(sX, sY, AX_TILE, c, t, i, j, gc, r[c]);
        if (alpha == 0) f.y = 0.f;
        scatter_sub(stiff, base + (size_t)c * st, f);
    }
}

#ifndef AX_GQ3D_MIN_CTAS
#define AX_GQ3D_MIN_CTAS 3   // resident 400-thread CTAs per SM the register allocation of k_grad3d / k_quad3d must allow
#endif
// ------------------------------------------------------------------------------------ 3D material: stage A (grad)
template <bool FLUID, bool PRT = false>
__global__ void __launch_bounds__(AX_TILE *AX_NPE, AX_GQ3D_MIN_CTAS) k_grad3d(const ElemDesc *__restrict__ elems, const int *__restrict__ w_elem,
                                                            const int *__restrict__ w_a0, const float *__restrict__ geom,
                                                            const float2 *__restrict__ displ, float2 *__restrict__ scratch) {
    constexpr int NC = FLUID ? 1 : 3;
    __shared__ float2 sU[NC * AX_NPE * AX_TILE];
    const ElemDesc &E = elems[w_elem[blockIdx.x]];
    const int t = threadIdx.x % AX_TILE, p = threadIdx.x / AX_TILE;
    const int i = p / 5, j = p % 5;
    const int alpha = w_a0[blockIdx.x] + t;
    const int N = E.nr;
    const bool active = alpha <= E.nu;
    gather_tile<NC>(E, displ, sU, t, p, active ? alpha : (1 << 30));
    GCoef gc;
    load_gcoef(gc, E.axial, i, j);
    const PointGeom g = load_geom(geom, E.geom_off, p);
    const bool ax0 = E.axial && i == 0;
    __syncthreads();
    if (!active) return;
    const bool dead = E.nyq && alpha == E.nu;
    float2 *z = scratch + E.scratch_off + (size_t)p * N;
    if constexpr (!FLUID) {
        if (PRT && E.prt) {   // 9 components -> 5 Z-form pairs (the last one half empty)
            float2 e9[9];
            grad9_point(sU, AX_TILE, t, i, j, gc, g, (float)alpha, ax0, e9);
            if (dead) {
#pragma unroll
                for (int c = 0; c < 9; ++c) e9[c] = czero();
            }
            float tr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) tr[k] = geom[E.trig_off + k * AX_NPE + p];
            rot9(e9, tr[0], tr[1], tr[2], tr[3], false);
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) zform_store(z + (size_t)pr * AX_NPE * N, N, alpha, e9[2 * pr], e9[2 * pr + 1]);
            zform_store(z + (size_t)4 * AX_NPE * N, N, alpha, e9[8], czero());
            return;
        }
        float2 e[6];
        grad6_point(sU, AX_TILE, t, i, j, gc, g, (float)alpha, ax0, e);
        if (dead) {
#pragma unroll
            for (int c = 0; c < 6; ++c) e[c] = czero();
        }
        if (E.tiso) {
            float tr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) tr[k] = geom[E.trig_off + k * AX_NPE + p];
            rot_spz_to_rtz(e, tr[0], tr[1], tr[2], tr[3]);
        }
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) zform_store(z + (size_t)pr * AX_NPE * N, N, alpha, e[2 * pr], e[2 * pr + 1]);
    } else {
        float2 e[3];
        grad_fluid_point(sU, AX_TILE, t, i, j, gc, g, (float)alpha, ax0, e);
        if (PRT && E.prt) rot3_fluid(e, geom[E.trig_off + p], geom[E.trig_off + AX_NPE + p], false);
        if (dead) e[0] = e[1] = e[2] = czero();
        zform_store(z, N, alpha, e[0], e[1]);
        zform_store(z + (size_t)AX_NPE * N, N, alpha, e[2], czero());
    }
}

// ------------------------------------------------------------------------------------ 3D material: stage B (c2r, stress, r2c)
// grid: one CTA per (element, group of `ppb` points); block 256.  smem: z[NPAIR*ppb][N] + tw[N].
struct FftItem {
    int elem;
    int p0;
};
// ------------------------------------------------------------------------------------ 3D material: stage C (quad)
template <bool FLUID, bool PRT = false>
__global__ void __launch_bounds__(AX_TILE *AX_NPE, AX_GQ3D_MIN_CTAS) k_quad3d(const ElemDesc *__restrict__ elems, const int *__restrict__ w_elem,
                                                            const int *__restrict__ w_a0, const float *__restrict__ geom,
                                                            const float2 *__restrict__ scratch, float2 *__restrict__ stiff) {
    constexpr int NC = FLUID ? 1 : 3;
    __shared__ float2 sX[NC * AX_NPE * AX_TILE];
    __shared__ float2 sY[NC * AX_NPE * AX_TILE];
    const ElemDesc &E = elems[w_elem[blockIdx.x]];
    const int t = threadIdx.x % AX_TILE, p = threadIdx.x / AX_TILE;
    const int i = p / 5, j = p % 5;
    const int beta = w_a0[blockIdx.x] + t;
    const int N = E.nr;
    const bool active = beta <= E.nu;
    const bool dead = !active || (E.nyq && beta == E.nu);
    GCoef gc;
    load_gcoef(gc, E.axial, i, j);
    const PointGeom g = load_geom(geom, E.geom_off, p);
    const bool ax0 = E.axial && i == 0;
    const float2 *z = scratch + E.scratch_off + (size_t)p * N;
    const float sc = 1.f / (float)N;   // SolverFFTW_N6::computeR2C scaling (SolverFFTW_N6.cpp:47-48)
    float2 r[NC];
    if constexpr (!FLUID) {
        float2 s[6], X[3], Y[3];
        if (PRT && E.prt) {
            float2 s9[9], dummy;
            if (!dead) {
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) zform_load(z + (size_t)pr * AX_NPE * N, N, beta, sc, s9[2 * pr], s9[2 * pr + 1]);
                zform_load(z + (size_t)4 * AX_NPE * N, N, beta, sc, s9[8], dummy);
                float tr[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) tr[k] = geom[E.trig_off + k * AX_NPE + p];
                rot9(s9, tr[0], tr[1], tr[2], tr[3], true);
            } else {
#pragma unroll
                for (int c = 0; c < 9; ++c) s9[c] = czero();
            }
            quad9_pre(s9, g, (float)beta, ax0, X, Y, r);
        } else {
        if (!dead) {
#pragma unroll
            for (int pr = 0; pr < 3; ++pr) zform_load(z + (size_t)pr * AX_NPE * N, N, beta, sc, s[2 * pr], s[2 * pr + 1]);
            if (E.tiso) {
                float tr[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) tr[k] = geom[E.trig_off + k * AX_NPE + p];
                rot_rtz_to_spz(s, tr[0], tr[1], tr[2], tr[3]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 6; ++c) s[c] = czero();
        }
        quad6_pre(s, g, (float)beta, ax0, X, Y, r);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            sX[(c * AX_NPE + p) * AX_TILE + t] = X[c];
            sY[(c * AX_NPE + p) * AX_TILE + t] = Y[c];
        }
    } else {
        float2 s[3], X, Y, dummy;
        if (!dead) {
            zform_load(z, N, beta, sc, s[0], s[1]);
            zform_load(z + (size_t)AX_NPE * N, N, beta, sc, s[2], dummy);
            if (PRT && E.prt) rot3_fluid(s, geom[E.trig_off + p], geom[E.trig_off + AX_NPE + p], true);
        } else {
            s[0] = s[1] = s[2] = czero();
        }
        quad_fluid_pre(s, g, (float)beta, ax0, X, Y, r[0]);
        sX[p * AX_TILE + t] = X;
        sY[p * AX_TILE + t] = Y;
    }
    __syncthreads();
    if (dead || beta >= E.pt_nlive[p]) return;
    const size_t base = (size_t)E.pt_off[p] + beta;
    const int st = E.pt_stride[p];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        float2 f = quad_post(sX, sY, AX_TILE, c, t, i, j, gc, r[c]);
        if (beta == 0) f.y = 0.f;
        scatter_sub(stiff, base + (size_t)c * st, f);
    }
}

// ------------------------------------------------------------------------------------ solid-fluid coupling
// SolidFluidPoint::coupleSolidFluid (SolidFluidPoint.cpp:106-110) with SFCoupling1D (SFCoupling1D.cpp:9-19);
// thread = (sf point, mode).  cpl: [4] = ns, nz, ns_invmf, nz_invmf.
struct SFTab {
    const unsigned *s_off, *f_off;
    const int *nu;
    const float *cpl;
    const int *row_point, *row_start;
    int nrows;
};
// one (sf point, mode) row.  L2-only accesses (ld.cg / st.cg): the in-kernel caller (fused.cuh: halo_put_cta) reads forces that
// other SMs have just RED-added
__device__ __forceinline__ void sf_couple_row(const SFTab &sf, int r, const float2 *__restrict__ s_displ, float2 *__restrict__ s_stiff,
                                              float2 *__restrict__ f_stiff) {
    const int q = sf.row_point[r];
    const int alpha = r - sf.row_start[q];
    const int st = sf.nu[q] + 1;
    const float *c = sf.cpl + 4 * q;
    const size_t so = (size_t)sf.s_off[q] + alpha, fo = (size_t)sf.f_off[q] + alpha;
    float2 us = __ldcg(s_displ + so), uz = __ldcg(s_displ + so + 2 * (size_t)st);
    float2 ff = __ldcg(f_stiff + fo);
    ff.x += c[0] * us.x + c[1] * uz.x;
    ff.y += c[0] * us.y + c[1] * uz.y;
    __stcg(f_stiff + fo, ff);
    float2 a = __ldcg(s_stiff + so), b = __ldcg(s_stiff + so + 2 * (size_t)st);
    a.x -= c[2] * ff.x; a.y -= c[2] * ff.y;
    b.x -= c[3] * ff.x; b.y -= c[3] * ff.y;
    __stcg(s_stiff + so, a);
    __stcg(s_stiff + so + 2 * (size_t)st, b);
}
__global__ void k_sf_couple(SFTab sf, const float2 *__restrict__ s_displ, float2 *__restrict__ s_stiff,
                            float2 *__restrict__ f_stiff) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= sf.nrows) return;
    sf_couple_row(sf, r, s_displ, s_stiff, f_stiff);
}

// SFCoupling3D (SFCoupling3D.cpp:9-54): one CTA per 3D solid-fluid point.
struct SF3DItem {
    unsigned s_off, f_off;
    int nu, plan_id;
    long long n_off;   // float [6][Nr]: n_un(s,phi,z), n_as_invmf(s,phi,z), digit-reversed phi
};
__global__ void __launch_bounds__(128) k_sf_couple3d(const SF3DItem *__restrict__ items, const FftPlan *__restrict__ plans,
                                                     const float2 *__restrict__ twpool, const float *__restrict__ npool,
                                                     const float2 *__restrict__ s_displ, float2 *__restrict__ s_stiff,
                                                     float2 *__restrict__ f_stiff) {
    extern __shared__ float2 smem[];
    const SF3DItem it = items[blockIdx.x];
    const FftPlan pl = plans[it.plan_id];
    const int N = pl.N, nu = it.nu, st = nu + 1;
    float2 *z = smem, *tw = smem + 2 * N;
    const float *nn = npool + it.n_off;
    const float sc = 1.f / (float)N;
    for (int k = threadIdx.x; k < N; k += blockDim.x) tw[k] = twpool[pl.tw_off + k];
    // solid displacement -> physical space (columns: s + i phi | z)
    for (int k = threadIdx.x; k <= nu; k += blockDim.x) {
        zform_store(z, N, k, s_displ[it.s_off + k], s_displ[it.s_off + st + k]);
        zform_store(z + N, N, k, s_displ[it.s_off + 2 * st + k], czero());
    }
    __syncthreads();
    fft_inverse_dif(pl, z, N, 2, tw, threadIdx.x, blockDim.x);
    for (int pos = threadIdx.x; pos < N; pos += blockDim.x) {
        float2 a = z[pos], b = z[N + pos];
        z[pos] = make_float2(nn[pos] * a.x + nn[N + pos] * a.y + nn[2 * N + pos] * b.x, 0.f);
    }
    __syncthreads();
    fft_forward_dit(pl, z, N, 1, tw, threadIdx.x, blockDim.x);
    // fluid stiff += ...; then fluid stiff -> physical space
    for (int k = threadIdx.x; k <= nu; k += blockDim.x) {
        float2 a, b;
        zform_load(z, N, k, sc, a, b);
        float2 ff = f_stiff[it.f_off + k];
        ff.x += a.x; ff.y += a.y;
        f_stiff[it.f_off + k] = ff;
        z[N + k] = ff;   // stash (second column is free)
    }
    __syncthreads();
    for (int k = threadIdx.x; k <= nu; k += blockDim.x) {
        float2 ff = z[N + k];
        zform_store(z, N, k, ff, czero());
    }
    __syncthreads();
    fft_inverse_dif(pl, z, N, 1, tw, threadIdx.x, blockDim.x);
    for (int pos = threadIdx.x; pos < N; pos += blockDim.x) {
        const float fr = z[pos].x;
        z[pos] = make_float2(nn[3 * N + pos] * fr, nn[4 * N + pos] * fr);
        z[N + pos] = make_float2(nn[5 * N + pos] * fr, 0.f);
    }
    __syncthreads();
    fft_forward_dit(pl, z, N, 2, tw, threadIdx.x, blockDim.x);
    for (int k = threadIdx.x; k <= nu; k += blockDim.x) {
        float2 a, b, c, d;
        zform_load(z, N, k, sc, a, b);
        zform_load(z + N, N, k, sc, c, d);
        float2 *s0 = &s_stiff[it.s_off + k], *s1 = &s_stiff[it.s_off + st + k], *s2 = &s_stiff[it.s_off + 2 * st + k];
        s0->x -= a.x; s0->y -= a.y;
        s1->x -= b.x; s1->y -= b.y;
        s2->x -= c.x; s2->y -= c.y;
    }
}

// ------------------------------------------------------------------------------------ halo
// idx[i] : bit31 = fluid array, low bits = offset in the stiff array
__global__ void k_pack(int n, const unsigned *__restrict__ idx, const float2 *__restrict__ s_stiff,
                       const float2 *__restrict__ f_stiff, float2 *__restrict__ buf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = idx[i];
    buf[i] = (k >> 31) ? f_stiff[k & 0x7fffffffu] : s_stiff[k];
}
__global__ void k_unpack_add(int n, const unsigned *__restrict__ idx, const float2 *__restrict__ buf,
                             float2 *__restrict__ s_stiff, float2 *__restrict__ f_stiff) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = idx[i];
    float2 *d = (k >> 31) ? &f_stiff[k & 0x7fffffffu] : &s_stiff[k];
    float2 v = buf[i];
    d->x += v.x;
    d->y += v.y;
}

// Peer-memory halo (Domain::assembleStiff over NVLink without a library call).  Every rank owns a receive window
// win[2][total] (float2; parity = exchange number & 1) and one arrival counter per neighbour; the windows of the neighbours
// are mapped into this process (cudaIpcOpenMemHandle).  k_halo_put packs the boundary stiffness of one neighbour straight
// into that neighbour's window (SolidPoint::feedBuffer order, SolidPoint.cpp:163-167) and then bumps the neighbour's
// counter once per block; k_halo_wait_add spins until all blocks of the matching put have arrived and does the
// extractBuffer add (SolidPoint.cpp:169-173).  `step` (device memory) numbers the exchanges: it is read by both kernels
// and advanced by k_halo_advance behind them, so the whole exchange is plain stream-ordered kernels and replays from a
// CUDA graph.  The parity double buffer is enough: put(s + 2) into a window is ordered behind the owner's wait_add(s) by
// the owner's own put(s + 1), which the writer waits for in its wait_add(s + 1).
__global__ void k_halo_put(int n, const unsigned *__restrict__ idx, const float2 *__restrict__ s_stiff,
                           const float2 *__restrict__ f_stiff, float2 *__restrict__ peer_win, size_t peer_parity_stride,
                           unsigned *__restrict__ peer_count, const unsigned *__restrict__ step) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned s = *step;
    if (i < n) {
        const unsigned k = idx[i];
        const float2 v = (k >> 31) ? f_stiff[k & 0x7fffffffu] : s_stiff[k];
        __stcg(peer_win + (size_t)(s & 1u) * peer_parity_stride + i, v);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd_system(peer_count, 1u);
}
__global__ void k_halo_wait_add(int n, const unsigned *__restrict__ idx, const float2 *__restrict__ win, size_t parity_stride,
                                const unsigned *__restrict__ count, const unsigned *__restrict__ step, float2 *__restrict__ s_stiff,
                                float2 *__restrict__ f_stiff, int *__restrict__ timeout_flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned s = *step;
    if (threadIdx.x == 0) {
        const unsigned want = (s + 1u) * gridDim.x;   // the matching put announces the same number of blocks
        unsigned long long spins = 0;
        for (;;) {
            unsigned c;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(c) : "l"(count) : "memory");
            if ((int)(c - want) >= 0) break;
            __nanosleep(100);
            if (++spins > (1ull << 27)) {   // ~20 s without the neighbour's boundary forces: never add stale data, never hang the
                *timeout_flag = 1;          // GPU -- flag it and abort the launch (the next API call reports the failure)
                __threadfence_system();
                __trap();
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (i < n) {
        const unsigned k = idx[i];
        float2 *d = (k >> 31) ? &f_stiff[k & 0x7fffffffu] : &s_stiff[k];
        const float2 v = __ldcv(win + (size_t)(s & 1u) * parity_stride + i);   // written by the peer: never from L1
        d->x += v.x;
        d->y += v.y;
    }
}
// In-kernel form of k_halo_put (fused.cuh: halo_put_cta): one CTA of the solid element kernel sends the boundary forces of every
// neighbour as soon as the last boundary element of the rank has scattered, while the interior elements are still being computed.
#define AX_MAX_NEIGH 16
struct HaloPeer {
    const unsigned *idx;          // pack list of this neighbour (k_halo_put order)
    float2 *win;                  // my segment in the neighbour's window (parity 0)
    unsigned long long stride;    // its parity stride
    unsigned *count;              // its arrival counter for me
    int n, nblocks;               // entries; blocks the matching k_halo_wait_add expects per exchange (= its grid)
};
struct HaloTab {
    int nneigh;
    HaloPeer peer[AX_MAX_NEIGH];
    const unsigned *step;         // exchange number (k_halo_advance)
    float2 *f_stiff;              // fluid stiffness (the kernel's own arrays are the solid ones)
    SFTab sf_halo;                // (sf point, mode) rows of the solid-fluid points on the halo: coupled before they are sent
};
__global__ void k_halo_advance(unsigned *step) { *step += 1u; }

// ------------------------------------------------------------------------------------ stability
__global__ void k_check_finite(size_t n, const float *__restrict__ a, int *__restrict__ bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    int b = 0;
    for (; i < n; i += stride) b |= !isfinite(a[i]);
    if (b) atomicOr(bad, 1);
}

// ------------------------------------------------------------------------------------ wisdom learning
// Point::learnWisdom (SolidPoint.cpp:240-266, FluidPoint.cpp:209-229): one warp per (point, component).  When the Hilbert
// norm h2 = |u|^2 - |u_0|^2 / 2 exceeds its running maximum, the smallest order newNu < Nu with
// |u|^2 - |u_{0..newNu}|^2 <= cutoff^2 h2 is stored (Nu if none).  wis_max starts at -1, wis_nu at Nu (SolidPoint.cpp:16).
template <int NCOMP>
__global__ void k_learn_wisdom(PointTab pt, int npoint, const float2 *__restrict__ displ, float cutoff, float *__restrict__ wis_max,
                               int *__restrict__ wis_nu) {
    const int w = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= npoint * NCOMP) return;
    const int p = w / NCOMP, c = w - p * NCOMP;
    const int nu = pt.nu[p];
    const float2 *u = displ + (size_t)pt.off[p] + (size_t)c * (nu + 1);
    float l2 = 0.f;
    for (int a = lane; a <= nu; a += 32) l2 += u[a].x * u[a].x + u[a].y * u[a].y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l2 += __shfl_xor_sync(0xffffffffu, l2, o);
    const float2 u0 = u[0];
    const float h2 = l2 - 0.5f * (u0.x * u0.x + u0.y * u0.y);
    if (h2 <= wis_max[w]) return;          // warp-uniform
    const float tol = h2 * cutoff * cutoff;
    int found = nu;
    float carry = 0.f;
    for (int base = 0; base < nu; base += 32) {
        const int a = base + lane;
        float v = a <= nu ? u[a].x * u[a].x + u[a].y * u[a].y : 0.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {   // inclusive scan
            const float t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        const float pre = carry + v;         // |u_{0..a}|^2
        const unsigned m = __ballot_sync(0xffffffffu, a < nu && l2 - pre <= tol);
        if (m) { found = base + __ffs(m) - 1; break; }
        carry = __shfl_sync(0xffffffffu, pre, 31);
    }
    if (lane == 0) {
        wis_max[w] = h2;
        wis_nu[w] = found;
    }
}

// ------------------------------------------------------------------------------------ receivers
// SolidElement::computeGroundMotion (SolidElement.cpp:189-216); one CTA per receiver, 128 threads.
struct RecvItem {
    int elem;
    float phi;
};
// Element::computeGroundMotion (SolidElement.cpp:189-216): u(phi) = sum_p w_p (u_0 + 2 Re sum_{alpha>=1} u_alpha e^{i alpha phi}).
// One CTA of AX_REC_NT threads per receiver; thread = (point, mode) pairs, warp-shuffle + shared reduction.
#define AX_REC_NT 512
__global__ void __launch_bounds__(AX_REC_NT) k_ground_motion(const ElemDesc *__restrict__ elems, const RecvItem *__restrict__ rec,
                                                             const float *__restrict__ weights, const float2 *__restrict__ displ,
                                                             float *__restrict__ out, const int *__restrict__ slot, int slot_stride) {
    if (slot) out += (size_t)(*slot) * slot_stride;   // device-side record ring (ax3d_run_steps_record): row of this step
    __shared__ float s_w[AX_NPE];
    __shared__ unsigned s_off[AX_NPE];
    __shared__ int s_stride[AX_NPE], s_nlive[AX_NPE];
    __shared__ float red[3][AX_REC_NT / 32];
    const RecvItem R = rec[blockIdx.x];
    const ElemDesc &E = elems[R.elem];
    if (threadIdx.x < AX_NPE) {
        const int p = threadIdx.x;
        s_w[p] = weights[(size_t)blockIdx.x * AX_NPE + p];
        s_off[p] = E.pt_off[p];
        s_stride[p] = E.pt_stride[p];
        s_nlive[p] = E.pt_nlive[p];
    }
    const int top = E.nu - E.nyq;   // last mode used
    __syncthreads();
    float acc[3] = {0.f, 0.f, 0.f};
    for (int idx = threadIdx.x; idx < AX_NPE * (top + 1); idx += AX_REC_NT) {
        const int p = idx / (top + 1), alpha = idx - p * (top + 1);
        const float wp = s_w[p];
        if (fabsf(wp) < 1e-10f || alpha >= s_nlive[p]) continue;
        float sn, cs;
        sincosf((float)alpha * R.phi, &sn, &cs);
        const float fac = alpha == 0 ? 1.f : 2.f;
        const size_t base = (size_t)s_off[p] + alpha;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float2 u = displ[base + (size_t)c * s_stride[p]];
            const float re = alpha == 0 ? u.x : (cs * u.x - sn * u.y);
            acc[c] += wp * fac * re;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        if ((threadIdx.x & 31) == 0) red[c][threadIdx.x >> 5] = acc[c];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < AX_REC_NT / 32; ++k) t += red[threadIdx.x][k];
        out[blockIdx.x * 3 + threadIdx.x] = t;
    }
}

// SolidElement::computeStrain (SolidElement.cpp:219-279) and computeCurl (281-345) without PRT: gather -> computeGrad6 /
// computeGrad9 -> transformSPZ_RTZ (the recorder has called forceTIso, SolidElement.cpp:347-352, so every receiver element
// rotates) -> evaluation at azimuth phi with the 25 interpolation weights.  One CTA of 16 modes x 25 points per receiver,
// mode tiles staged like k_grad3d.  out: 6 Voigt strains (RTZ) or 3 curl components per receiver.
template <bool CURL>
__global__ void __launch_bounds__(AX_TILE *AX_NPE) k_strain_curl(const ElemDesc *__restrict__ elems, const RecvItem *__restrict__ rec,
                                                                 const float *__restrict__ weights, const float *__restrict__ geom,
                                                                 const float2 *__restrict__ displ, float *__restrict__ out) {
    constexpr int NOUT = CURL ? 3 : 6;
    __shared__ float2 sU[3 * AX_NPE * AX_TILE];
    __shared__ float red[NOUT][AX_TILE * AX_NPE / 32 + 1];
    const RecvItem R = rec[blockIdx.x];
    const ElemDesc &E = elems[R.elem];
    const int t = threadIdx.x % AX_TILE, p = threadIdx.x / AX_TILE;
    const int i = p / 5, j = p % 5;
    const int top = E.nu - E.nyq;
    GCoef gc;
    load_gcoef(gc, E.axial, i, j);
    const PointGeom g = load_geom(geom, E.geom_off, p);
    const bool ax0 = E.axial && i == 0;
    float tr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) tr[k] = geom[E.trig_off + k * AX_NPE + p];
    const float wp = weights[(size_t)blockIdx.x * AX_NPE + p];
    float acc[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) acc[k] = 0.f;
    for (int a0 = 0; a0 <= top; a0 += AX_TILE) {
        const int alpha = a0 + t;
        __syncthreads();
        gather_tile<3>(E, displ, sU, t, p, alpha <= E.nu ? alpha : (1 << 30));
        __syncthreads();
        if (alpha > top || fabsf(wp) < 1e-10f) continue;
        float sn, cs;
        sincosf((float)alpha * R.phi, &sn, &cs);
        const float f = (alpha == 0 ? 1.f : 2.f) * wp;
        auto ev = [&](float2 v) { return f * (alpha == 0 ? v.x : cs * v.x - sn * v.y); };
        if constexpr (CURL) {
            float2 e9[9];
            grad9_point(sU, AX_TILE, t, i, j, gc, g, (float)alpha, ax0, e9);
            rot9(e9, tr[0], tr[1], tr[2], tr[3], false);
            acc[0] += ev(csub(e9[7], e9[5]));      // dUZdT - dUTdZ
            acc[1] += ev(csub(e9[2], e9[6]));      // dURdZ - dUZdR
            acc[2] += ev(csub(e9[3], e9[1]));      // dUTdR - dURdT
        } else {
            float2 e[6];
            grad6_point(sU, AX_TILE, t, i, j, gc, g, (float)alpha, ax0, e);
            rot_spz_to_rtz(e, tr[0], tr[1], tr[2], tr[3]);
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[k] += ev(e[k]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < NOUT) {
        float s = 0.f;
        for (int k = 0; k < (AX_TILE * AX_NPE + 31) / 32; ++k) s += red[threadIdx.x][k];
        out[blockIdx.x * NOUT + threadIdx.x] = s;
    }
}

// FluidElement::computeStrain (FluidElement.cpp:219-306) for Acoustic1D elements without particle relabelling: the fluid
// displacement u = K grad(chi) of a mode tile is formed for all 25 points in shared memory (with a 1D K the SPZ <-> RTZ
// rotations around Acoustic1D::strainToStress cancel), then Gradient::computeGrad6 of it, rotated to RTZ (forceTIso) and
// evaluated at azimuth phi with the 25 interpolation weights -- the second half is k_strain_curl's.
__global__ void __launch_bounds__(AX_TILE *AX_NPE) k_strain_fluid1d(const ElemDesc *__restrict__ elems, const RecvItem *__restrict__ rec,
                                                                    const float *__restrict__ weights, const float *__restrict__ geom,
                                                                    const float *__restrict__ coef, const float2 *__restrict__ displ,
                                                                    float *__restrict__ out) {
    __shared__ float2 sC[AX_NPE * AX_TILE];
    __shared__ float2 sU[3 * AX_NPE * AX_TILE];
    __shared__ float red[6][AX_TILE * AX_NPE / 32 + 1];
    const RecvItem R = rec[blockIdx.x];
    const ElemDesc &E = elems[R.elem];
    const int t = threadIdx.x % AX_TILE, p = threadIdx.x / AX_TILE;
    const int i = p / 5, j = p % 5;
    const int top = E.nu - E.nyq;
    GCoef gc;
    load_gcoef(gc, E.axial, i, j);
    const PointGeom g = load_geom(geom, E.geom_off, p);
    const bool ax0 = E.axial && i == 0;
    float tr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) tr[k] = geom[E.trig_off + k * AX_NPE + p];
    const float K = coef[E.coef_off + p];
    const float wp = weights[(size_t)blockIdx.x * AX_NPE + p];
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int a0 = 0; a0 <= top; a0 += AX_TILE) {
        const int alpha = a0 + t;
        __syncthreads();
        gather_tile<1>(E, displ, sC, t, p, alpha <= E.nu ? alpha : (1 << 30));
        __syncthreads();
        {
            float2 e3[3];
            grad_fluid_point(sC, AX_TILE, t, i, j, gc, g, (float)alpha, ax0, e3);
            if (alpha > top) e3[0] = e3[1] = e3[2] = czero();
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float2 u = cscale(e3[c], K);
                if (alpha == 0) u.y = 0.f;
                sU[(c * AX_NPE + p) * AX_TILE + t] = u;
            }
        }
        __syncthreads();
        if (alpha > top || fabsf(wp) < 1e-10f) continue;
        float sn, cs;
        sincosf((float)alpha * R.phi, &sn, &cs);
        const float f = (alpha == 0 ? 1.f : 2.f) * wp;
        float2 e[6];
        grad6_point(sU, AX_TILE, t, i, j, gc, g, (float)alpha, ax0, e);
        rot_spz_to_rtz(e, tr[0], tr[1], tr[2], tr[3]);
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[k] += f * (alpha == 0 ? e[k].x : cs * e[k].x - sn * e[k].y);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float sm = 0.f;
        for (int k = 0; k < (AX_TILE * AX_NPE + 31) / 32; ++k) sm += red[threadIdx.x][k];
        out[blockIdx.x * 6 + threadIdx.x] = sm;
    }
}

// FluidElement::computeGroundMotion (FluidElement.cpp:163-215): the fluid "displacement" is the acoustic stress of the
// potential, u = K grad(chi) -- gather -> Gradient::computeGrad -> [c2r -> K(phi) -> r2c for 3D material] -- evaluated at
// azimuth phi and interpolated with the receiver's 25 weights.  One CTA per receiver, one GLL row (5 points) at a time so
// that the 3D case needs only 10 Z-form columns of shared memory whatever Nr; the potential is read straight from the
// point arrays (a receiver kernel: <= a few hundred CTAs per recorded step, off the hot path).
#define AX_RECF_NT 256
__device__ __forceinline__ float2 recf_u(const ElemDesc &E, const float2 *__restrict__ displ, int p, int a) {
    float2 v = a < E.pt_nlive[p] ? displ[(size_t)E.pt_off[p] + a] : czero();
    if (a == 0) v.y = 0.f;
    return v;
}
__device__ __forceinline__ void recf_grad(const ElemDesc &E, const float2 *__restrict__ displ, const float *__restrict__ geom, int i, int j,
                                          int a, float2 (&e)[3]) {
    GCoef gc;
    load_gcoef(gc, E.axial, i, j);
    const PointGeom g = load_geom(geom, E.geom_off, i * 5 + j);
    float2 GU = czero(), UG = czero();
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        GU = cfma(gc.gxi_col[k], recf_u(E, displ, k * 5 + j, a), GU);
        UG = cfma(gc.geta_col[k], recf_u(E, displ, i * 5 + k, a), UG);
    }
    const float alpha = (float)a;
    const float2 v = mul_ialpha(recf_u(E, displ, i * 5 + j, a), alpha);
    e[0] = cfma(g.dzdeta, GU, cscale(UG, g.dzdxii));
    e[1] = cscale(v, g.inv_s);
    e[2] = cfma(g.dsdeta, GU, cscale(UG, g.dsdxii));
    if (E.axial && i == 0) e[1] = cfma(g.dzdeta, mul_ialpha(GU, alpha), e[1]);
    if (E.nyq && a == E.nu) e[0] = e[1] = e[2] = czero();
    if (E.prt) rot3_fluid(e, geom[E.trig_off + i * 5 + j], geom[E.trig_off + AX_NPE + i * 5 + j], false);   // FluidElement.cpp:176-178
}
__global__ void __launch_bounds__(AX_RECF_NT) k_ground_motion_fluid(const ElemDesc *__restrict__ elems, const RecvItem *__restrict__ rec,
                                                                    const float *__restrict__ weights, const FftPlan *__restrict__ plans,
                                                                    const float2 *__restrict__ twpool, const float *__restrict__ geom,
                                                                    const float *__restrict__ coef, const float2 *__restrict__ displ,
                                                                    float *__restrict__ out, const int *__restrict__ slot, int slot_stride) {
    if (slot) out += (size_t)(*slot) * slot_stride;
    extern __shared__ float2 zsm[];            // 3D material only: [2 pairs][5 points][ldz]
    __shared__ float red[3][AX_RECF_NT / 32];
    __shared__ FftPlan sP;
    const int tid = threadIdx.x;
    const RecvItem R = rec[blockIdx.x];
    const ElemDesc &E = elems[R.elem];
    const int N = E.nr, top = E.nu - E.nyq, nm = top + 1;
    const bool is3d = E.is3d != 0;
    const int ldz = (N + 1) | 1;
    if (is3d && tid < (int)(sizeof(FftPlan) / sizeof(int))) reinterpret_cast<int *>(&sP)[tid] = reinterpret_cast<const int *>(plans + E.plan_id)[tid];
    __syncthreads();
    const float *w = weights + (size_t)blockIdx.x * AX_NPE;
    float acc[3] = {0.f, 0.f, 0.f};
    auto add = [&](int p, int a, const float2 (&s)[3]) {
        float sn, cs;
        sincosf((float)a * R.phi, &sn, &cs);
        const float f = (a == 0 ? 1.f : 2.f) * w[p];
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += f * (a == 0 ? s[c].x : cs * s[c].x - sn * s[c].y);
    };
    for (int i = 0; i < 5; ++i) {
        if (!is3d) {   // Acoustic1D::strainToStress (Acoustic1D.cpp:8-14): K per point, in Fourier space
            for (int idx = tid; idx < 5 * nm; idx += AX_RECF_NT) {
                const int j = idx / nm, a = idx - j * nm, p = i * 5 + j;
                float2 e[3];
                recf_grad(E, displ, geom, i, j, a, e);
                const float K = coef[E.coef_off + p];
                float px[4] = {1.f, 0.f, 0.f, 1.f};
                if (E.prt) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) px[k] = coef[E.prt_off + k * AX_NPE + p];
                    prt_s2u_fluid(e, px);
                }
                float2 s[3] = {cscale(e[0], K), cscale(e[1], K), cscale(e[2], K)};
                if (E.prt) {
                    prt_u2s_fluid(s, px);
                    rot3_fluid(s, geom[E.trig_off + p], geom[E.trig_off + AX_NPE + p], true);
                }
                add(p, a, s);
            }
            continue;
        }
        const int M = E.nu + 1;
        for (int idx = tid; idx < 5 * M; idx += AX_RECF_NT) {
            const int j = idx / M, a = idx - j * M;
            float2 e[3];
            recf_grad(E, displ, geom, i, j, a, e);
            zform_store(zsm + j * ldz, N, a, e[0], e[1]);
            zform_store(zsm + (5 + j) * ldz, N, a, e[2], czero());
        }
        __syncthreads();
        fft_inverse_dif(sP, zsm, ldz, 10, twpool + sP.tw_off, tid, AX_RECF_NT);
        for (int idx = tid; idx < 5 * N; idx += AX_RECF_NT) {   // Acoustic3D::strainToStress (Acoustic3D.cpp:9-16), digit-reversed phi
            const int j = idx / N, pos = idx - j * N;
            const float K = coef[E.coef_off + (size_t)(i * 5 + j) * N + pos];
            const float2 z0 = zsm[j * ldz + pos], z1 = zsm[(5 + j) * ldz + pos];
            float ee[3] = {z0.x, z0.y, z1.x};
            float px[4] = {1.f, 0.f, 0.f, 1.f};
            if (E.prt) {
#pragma unroll
                for (int k = 0; k < 4; ++k) px[k] = coef[E.prt_off + ((size_t)k * AX_NPE + i * 5 + j) * N + pos];
                prt_s2u_fluid(ee, px);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) ee[c] *= K;
            if (E.prt) prt_u2s_fluid(ee, px);
            zsm[j * ldz + pos] = make_float2(ee[0], ee[1]);
            zsm[(5 + j) * ldz + pos] = make_float2(ee[2], 0.f);
        }
        __syncthreads();
        fft_forward_dit(sP, zsm, ldz, 10, twpool + sP.tw_off, tid, AX_RECF_NT);
        const float sc = 1.f / (float)N;
        for (int idx = tid; idx < 5 * nm; idx += AX_RECF_NT) {
            const int j = idx / nm, a = idx - j * nm;
            float2 s[3], dummy;
            zform_load(zsm + j * ldz, N, a, sc, s[0], s[1]);
            zform_load(zsm + (5 + j) * ldz, N, a, sc, s[2], dummy);
            if (E.prt) rot3_fluid(s, geom[E.trig_off + i * 5 + j], geom[E.trig_off + AX_NPE + i * 5 + j], true);
            add(i * 5 + j, a, s);
        }
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        if ((tid & 31) == 0) red[c][tid >> 5] = acc[c];
    }
    __syncthreads();
    if (tid < 3) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < AX_RECF_NT / 32; ++k) t += red[tid][k];
        out[blockIdx.x * 3 + tid] = t;
    }
}
