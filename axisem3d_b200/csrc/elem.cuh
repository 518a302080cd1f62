// elem.cuh -- per-(mode, GLL point) device functions of the element stiffness path.
//
// One thread owns one (Fourier mode alpha, GLL point p = ipol*5 + jpol) of one element.  The only
// cross-point coupling (the 5x5 tensor-product derivative) goes through a shared-memory tile
// laid out [component][point][mode-in-tile] so that lanes (= consecutive modes) read consecutive
// float2 words.  Everything else is pointwise in registers.
#pragma once
#include "fft.cuh"

#define AX_NPE 25
#define AX_TILE 16   // Fourier modes per CTA tile in the grad/quad kernels (400 threads)

enum { LAW_ISO = 0, LAW_TI = 1, LAW_ANISO = 2 };
enum { ATT_NONE = 0, ATT_FULL = 1, ATT_CG4 = 2 };

// i * alpha * a   and   -i * beta * a
__device__ __forceinline__ float2 mul_ialpha(float2 a, float al) { return make_float2(-al * a.y, al * a.x); }
__device__ __forceinline__ float2 mul_mibeta(float2 a, float be) { return make_float2(be * a.y, -be * a.x); }
__device__ __forceinline__ float2 czero() { return make_float2(0.f, 0.f); }

struct PointGeom {
    float dsdxii, dsdeta, dzdxii, dzdeta, inv_s;
};

// G matrices in constant memory: c_G[0] = G_GLL, c_G[1] = G_GLJ; G[i][j] = l_i'(x_j), row-major
// (Gradient::setGMat, Gradient.cpp:324-329).
__constant__ float c_G[2][25];

struct GCoef {
    float gxi_col[5];   // Gxi[k][i]  : (Gxi^T u)(i, j) = sum_k Gxi[k][i] u[k][j]
    float geta_col[5];  // Geta[k][j] : (u Geta)(i, j)  = sum_k u[i][k] Geta[k][j]
    float gxi_row[5];   // Gxi[i][k]  : (Gxi X)(i, j)   = sum_k Gxi[i][k] X[k][j]
    float geta_row[5];  // Geta[j][k] : (Y Geta^T)(i, j) = sum_k Y[i][k] Geta[j][k]
};

__device__ __forceinline__ void load_gcoef(GCoef &g, int axial, int i, int j) {
    const float *Gxi = c_G[axial ? 1 : 0];
    const float *Geta = c_G[0];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        g.gxi_col[k] = Gxi[k * 5 + i];
        g.geta_col[k] = Geta[k * 5 + j];
        g.gxi_row[k] = Gxi[i * 5 + k];
        g.geta_row[k] = Geta[j * 5 + k];
    }
}

// ------------------------------------------------------------------------------------------
// Gradient::computeGrad6 (Gradient.cpp:206-265) at one (alpha, point).  sU: [3][25][T] tile.
// sU layout: [component][point][T] (T = modes per tile, runtime).
// Writes e[6] in Voigt order [ss, pp, zz, pz, sz, sp] (engineering shear).
__device__ __forceinline__ void grad6_point(const float2 *sU, int T, int t, int i, int j, const GCoef &gc,
                                            const PointGeom &g, float alpha, bool axial_row0,
                                            float2 (&e)[6]) {
    float2 GU[3], UG[3], u[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float2 a = czero(), b = czero();
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            a = cfma(gc.gxi_col[k], sU[(c * AX_NPE + k * 5 + j) * T + t], a);
            b = cfma(gc.geta_col[k], sU[(c * AX_NPE + i * 5 + k) * T + t], b);
        }
        GU[c] = a;
        UG[c] = b;
        u[c] = sU[(c * AX_NPE + i * 5 + j) * T + t];
    }
    float2 ds[3], dz[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        ds[c] = cfma(g.dzdeta, GU[c], cscale(UG[c], g.dzdxii));
        dz[c] = cfma(g.dsdeta, GU[c], cscale(UG[c], g.dsdxii));
    }
    float2 v0 = cadd(u[0], mul_ialpha(u[1], alpha));
    float2 v1 = csub(mul_ialpha(u[0], alpha), u[1]);
    float2 v2 = mul_ialpha(u[2], alpha);
    e[0] = ds[0];
    e[1] = cscale(v0, g.inv_s);
    e[2] = dz[2];
    e[3] = cfma(g.inv_s, v2, dz[1]);
    e[4] = cadd(dz[0], ds[2]);
    e[5] = cfma(g.inv_s, v1, ds[1]);
    if (axial_row0) {   // L'Hopital rows on the axis (Gradient.cpp:221-224, 245-254)
        float2 gv0 = cadd(GU[0], mul_ialpha(GU[1], alpha));
        float2 gv1 = csub(mul_ialpha(GU[0], alpha), GU[1]);
        float2 gv2 = mul_ialpha(GU[2], alpha);
        e[1] = cfma(g.dzdeta, gv0, e[1]);
        e[5] = cfma(g.dzdeta, gv1, e[5]);
        e[3] = cfma(g.dzdeta, gv2, e[3]);
        if (alpha == 1.f) {
            float2 uv0 = cadd(UG[0], mul_ialpha(UG[1], alpha));
            float2 uv1 = csub(mul_ialpha(UG[0], alpha), UG[1]);
            e[1] = cfma(g.dzdxii, uv0, e[1]);
            e[5] = cfma(g.dzdxii, uv1, e[5]);
        }
    }
}

// Gradient::computeGrad (fluid, Gradient.cpp:26-57).  sU: [1][25][T]; e[3].
__device__ __forceinline__ void grad_fluid_point(const float2 *sU, int T, int t, int i, int j, const GCoef &gc,
                                                 const PointGeom &g, float alpha, bool axial_row0,
                                                 float2 (&e)[3]) {
    float2 GU = czero(), UG = czero();
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        GU = cfma(gc.gxi_col[k], sU[(k * 5 + j) * T + t], GU);
        UG = cfma(gc.geta_col[k], sU[(i * 5 + k) * T + t], UG);
    }
    float2 u = sU[(i * 5 + j) * T + t];
    float2 v = mul_ialpha(u, alpha);
    e[0] = cfma(g.dzdeta, GU, cscale(UG, g.dzdxii));
    e[1] = cscale(v, g.inv_s);
    e[2] = cfma(g.dsdeta, GU, cscale(UG, g.dsdxii));
    if (axial_row0) e[1] = cfma(g.dzdeta, mul_ialpha(GU, alpha), e[1]);
}

// ------------------------------------------------------------------------------------------
// Gradient::computeQuad6 (Gradient.cpp:267-322), pointwise half: from s[6] build X_c, Y_c (with the
// axial terms folded in) and r_c = inv_s * g_c.
__device__ __forceinline__ void quad6_pre(const float2 (&s)[6], const PointGeom &g, float beta, bool axial_row0,
                                          float2 (&X)[3], float2 (&Y)[3], float2 (&r)[3]) {
    float2 g0 = cadd(s[1], mul_mibeta(s[5], beta));
    float2 g1 = csub(mul_mibeta(s[1], beta), s[5]);
    float2 g2 = mul_mibeta(s[3], beta);
    const int pa[3] = {0, 5, 4}, pb[3] = {4, 3, 2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        X[c] = cfma(g.dzdeta, s[pa[c]], cscale(s[pb[c]], g.dsdeta));
        Y[c] = cfma(g.dzdxii, s[pa[c]], cscale(s[pb[c]], g.dsdxii));
    }
    float2 gg[3] = {g0, g1, g2};
#pragma unroll
    for (int c = 0; c < 3; ++c) r[c] = cscale(gg[c], g.inv_s);
    if (axial_row0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) X[c] = cfma(g.dzdeta, gg[c], X[c]);
        if (beta == 1.f) {
            Y[0] = cfma(g.dzdxii, gg[0], Y[0]);
            Y[1] = cfma(g.dzdxii, gg[1], Y[1]);
        }
    }
}

// tensor-product half: f = Gxi X + Y Geta^T + r.  sX, sY: [NC][25][T].
__device__ __forceinline__ float2 quad_post(const float2 *sX, const float2 *sY, int T, int c, int t, int i, int j,
                                            const GCoef &gc, float2 r) {
    float2 f = r;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        f = cfma(gc.gxi_row[k], sX[(c * AX_NPE + k * 5 + j) * T + t], f);
        f = cfma(gc.geta_row[k], sY[(c * AX_NPE + i * 5 + k) * T + t], f);
    }
    return f;
}

// Gradient::computeQuad (fluid, Gradient.cpp:59-82), pointwise half.
__device__ __forceinline__ void quad_fluid_pre(const float2 (&s)[3], const PointGeom &g, float beta, bool axial_row0,
                                               float2 &X, float2 &Y, float2 &r) {
    float2 gg = mul_mibeta(s[1], beta);
    X = cfma(g.dzdeta, s[0], cscale(s[2], g.dsdeta));
    Y = cfma(g.dzdxii, s[0], cscale(s[2], g.dsdxii));
    r = cscale(gg, g.inv_s);
    if (axial_row0) X = cfma(g.dzdeta, gg, X);
}

// ------------------------------------------------------------------------------------------
// CrdTransTIsoSolid (S/core/element/crd/CrdTransTIsoSolid.cpp:14-42); trig = {sin t, cos t, sin 2t, cos 2t}
template <typename V>
__device__ __forceinline__ V vadd(V a, V b);
template <> __device__ __forceinline__ float vadd(float a, float b) { return a + b; }
template <> __device__ __forceinline__ float2 vadd(float2 a, float2 b) { return cadd(a, b); }
template <typename V>
__device__ __forceinline__ V vsub(V a, V b);
template <> __device__ __forceinline__ float vsub(float a, float b) { return a - b; }
template <> __device__ __forceinline__ float2 vsub(float2 a, float2 b) { return csub(a, b); }
template <typename V>
__device__ __forceinline__ V vscale(V a, float s);
template <> __device__ __forceinline__ float vscale(float a, float s) { return a * s; }
template <> __device__ __forceinline__ float2 vscale(float2 a, float s) { return cscale(a, s); }

template <typename V>
__device__ __forceinline__ void rot_spz_to_rtz(V (&u)[6], float s1, float c1, float s2, float c2) {
    V sum02 = vadd(u[0], u[2]), dif02 = vsub(u[0], u[2]), u3 = u[3];
    u[0] = vscale(vsub(vadd(sum02, vscale(dif02, c2)), vscale(u[4], s2)), 0.5f);
    u[2] = vsub(sum02, u[0]);
    u[4] = vadd(vscale(u[4], c2), vscale(dif02, s2));
    u[3] = vadd(vscale(u3, c1), vscale(u[5], s1));
    u[5] = vsub(vscale(u[5], c1), vscale(u3, s1));
}
template <typename V>
__device__ __forceinline__ void rot_rtz_to_spz(V (&u)[6], float s1, float c1, float s2, float c2) {
    V sum02 = vadd(u[0], u[2]), dif02 = vscale(vsub(u[0], u[2]), 0.5f), u3 = u[3];
    u[0] = vadd(vadd(vscale(sum02, 0.5f), vscale(dif02, c2)), vscale(u[4], s2));
    u[2] = vsub(sum02, u[0]);
    u[4] = vsub(vscale(u[4], c2), vscale(dif02, s2));
    u[3] = vsub(vscale(u3, c1), vscale(u[5], s1));
    u[5] = vadd(vscale(u[5], c1), vscale(u3, s1));
}

// ------------------------------------------------------------------------------------------
// Particle relabelling (undulated interfaces): the 9-component path of SolidElement::displToStiff (SolidElement.cpp:405-432)
// and the PRT steps of FluidElement::displToStiff (FluidElement.cpp:333-355).
//
// Gradient::computeGrad9 (Gradient.cpp:84-141) at one (alpha, point): e[3 c + d] = d-th derivative of component c,
// d = (s, phi incl. curvature terms, z).  sU as in grad6_point.
__device__ __forceinline__ void grad9_point(const float2 *sU, int T, int t, int i, int j, const GCoef &gc, const PointGeom &g,
                                            float alpha, bool axial_row0, float2 (&e)[9]) {
    float2 GU[3], UG[3], u[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float2 a = czero(), b = czero();
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            a = cfma(gc.gxi_col[k], sU[(c * AX_NPE + k * 5 + j) * T + t], a);
            b = cfma(gc.geta_col[k], sU[(c * AX_NPE + i * 5 + k) * T + t], b);
        }
        GU[c] = a;
        UG[c] = b;
        u[c] = sU[(c * AX_NPE + i * 5 + j) * T + t];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        e[3 * c + 0] = cfma(g.dzdeta, GU[c], cscale(UG[c], g.dzdxii));
        e[3 * c + 2] = cfma(g.dsdeta, GU[c], cscale(UG[c], g.dsdxii));
    }
    const float2 v0 = cadd(u[0], mul_ialpha(u[1], alpha));
    const float2 v1 = csub(mul_ialpha(u[0], alpha), u[1]);
    const float2 v2 = mul_ialpha(u[2], alpha);
    e[1] = cscale(v1, g.inv_s);
    e[4] = cscale(v0, g.inv_s);
    e[7] = cscale(v2, g.inv_s);
    if (axial_row0) {   // Gradient.cpp:101-104, 127-136
        const float2 gv0 = cadd(GU[0], mul_ialpha(GU[1], alpha));
        const float2 gv1 = csub(mul_ialpha(GU[0], alpha), GU[1]);
        const float2 gv2 = mul_ialpha(GU[2], alpha);
        e[4] = cfma(g.dzdeta, gv0, e[4]);
        e[1] = cfma(g.dzdeta, gv1, e[1]);
        e[7] = cfma(g.dzdeta, gv2, e[7]);
        if (alpha == 1.f) {
            const float2 uv0 = cadd(UG[0], mul_ialpha(UG[1], alpha));
            const float2 uv1 = csub(mul_ialpha(UG[0], alpha), UG[1]);
            e[4] = cfma(g.dzdxii, uv0, e[4]);
            e[1] = cfma(g.dzdxii, uv1, e[1]);
        }
    }
}

// Gradient::computeQuad9 (Gradient.cpp:143-204), pointwise half: same X, Y, r as quad6_pre, from 9 stresses.
__device__ __forceinline__ void quad9_pre(const float2 (&s)[9], const PointGeom &g, float beta, bool axial_row0,
                                          float2 (&X)[3], float2 (&Y)[3], float2 (&r)[3]) {
    const float2 gg[3] = {cadd(s[4], mul_mibeta(s[1], beta)), csub(mul_mibeta(s[4], beta), s[1]), mul_mibeta(s[7], beta)};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        X[c] = cfma(g.dzdeta, s[3 * c], cscale(s[3 * c + 2], g.dsdeta));
        Y[c] = cfma(g.dzdxii, s[3 * c], cscale(s[3 * c + 2], g.dsdxii));
        r[c] = cscale(gg[c], g.inv_s);
    }
    if (axial_row0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) X[c] = cfma(g.dzdeta, gg[c], X[c]);
        if (beta == 1.f) {
            Y[0] = cfma(g.dzdxii, gg[0], Y[0]);
            Y[1] = cfma(g.dzdxii, gg[1], Y[1]);
        }
    }
}

// CrdTransTIsoSolid::transformSPZ_RTZ / RTZ_SPZ on 9 components (CrdTransTIsoSolid.cpp:44-83); back: s1, s2 -> -s1, -s2
template <typename V>
__device__ __forceinline__ void rot9(V (&u)[9], float s1, float c1, float s2, float c2, bool back) {
    if (back) { s1 = -s1; s2 = -s2; }
    const V sum08 = vadd(u[0], u[8]), dif08 = vsub(u[0], u[8]);
    const V sum26 = vadd(u[2], u[6]), dif26 = vsub(u[2], u[6]);
    const V u1 = u[1], u3 = u[3];
    u[0] = vscale(vsub(vadd(sum08, vscale(dif08, c2)), vscale(sum26, s2)), 0.5f);
    u[2] = vscale(vadd(vadd(dif26, vscale(sum26, c2)), vscale(dif08, s2)), 0.5f);
    u[6] = vsub(u[2], dif26);
    u[8] = vsub(sum08, u[0]);
    u[1] = vsub(vscale(u1, c1), vscale(u[7], s1));
    u[7] = vadd(vscale(u[7], c1), vscale(u1, s1));
    u[3] = vsub(vscale(u3, c1), vscale(u[5], s1));
    u[5] = vadd(vscale(u[5], c1), vscale(u3, s1));
}
// CrdTransTIsoFluid::transformSPZ_RTZ / RTZ_SPZ on 3 components (CrdTransTIsoFluid.cpp:16-34)
template <typename V>
__device__ __forceinline__ void rot3_fluid(V (&u)[3], float s1, float c1, bool back) {
    if (back) s1 = -s1;
    const V u0 = u[0];
    u[0] = vsub(vscale(u0, c1), vscale(u[2], s1));
    u[2] = vadd(vscale(u[2], c1), vscale(u0, s1));
}

// PRT_1D/3D::sphericalToUndulated(SolidResponse) (PRT_1D.cpp:32-52, PRT_3D.cpp:41-62): 9 -> 6; X[4] at this point (and phi)
template <typename V>
__device__ __forceinline__ void prt_s2u_solid(const V (&p)[9], const float (&X)[4], V (&u)[6]) {
    u[0] = vadd(vscale(p[0], X[0]), vscale(p[2], X[1]));
    u[1] = vadd(vscale(p[4], X[0]), vscale(p[5], X[2]));
    u[2] = vscale(p[8], X[3]);
    u[3] = vadd(vadd(vscale(p[7], X[0]), vscale(p[8], X[2])), vscale(p[5], X[3]));
    u[4] = vadd(vadd(vscale(p[6], X[0]), vscale(p[8], X[1])), vscale(p[2], X[3]));
    u[5] = vadd(vadd(vscale(vadd(p[3], p[1]), X[0]), vscale(p[5], X[1])), vscale(p[2], X[2]));
}
// PRT_1D/3D::undulatedToSpherical(SolidResponse) (PRT_1D.cpp:54-76, PRT_3D.cpp:64-86): 6 -> 9
template <typename V>
__device__ __forceinline__ void prt_u2s_solid(const V (&u)[6], const float (&X)[4], V (&p)[9]) {
    p[0] = vscale(u[0], X[0]);
    p[1] = vscale(u[5], X[0]);
    p[2] = vadd(vadd(vscale(u[0], X[1]), vscale(u[4], X[3])), vscale(u[5], X[2]));
    p[3] = p[1];
    p[4] = vscale(u[1], X[0]);
    p[5] = vadd(vadd(vscale(u[1], X[2]), vscale(u[3], X[3])), vscale(u[5], X[1]));
    p[6] = vscale(u[4], X[0]);
    p[7] = vscale(u[3], X[0]);
    p[8] = vadd(vadd(vscale(u[2], X[3]), vscale(u[3], X[2])), vscale(u[4], X[1]));
}
// PRT_1D/3D::sphericalToUndulated / undulatedToSpherical(FluidResponse) (PRT_1D.cpp:9-30, PRT_3D.cpp:21-39), in place
template <typename V>
__device__ __forceinline__ void prt_s2u_fluid(V (&e)[3], const float (&X)[4]) {
    e[0] = vadd(vscale(e[0], X[0]), vscale(e[2], X[1]));
    e[1] = vadd(vscale(e[1], X[0]), vscale(e[2], X[2]));
    e[2] = vscale(e[2], X[3]);
}
template <typename V>
__device__ __forceinline__ void prt_u2s_fluid(V (&s)[3], const float (&X)[4]) {
    s[2] = vadd(vadd(vscale(s[0], X[1]), vscale(s[1], X[2])), vscale(s[2], X[3]));
    s[0] = vscale(s[0], X[0]);
    s[1] = vscale(s[1], X[0]);
}

// ------------------------------------------------------------------------------------------
// constitutive laws on V = float (physical space, 3D classes) or float2 (Fourier space, 1D classes).
// coef(k) returns the k-th modulus at this (point[, phi]).
template <typename V, typename CoefFn>
__device__ __forceinline__ void stress_law(int law, const V (&e)[6], V (&s)[6], CoefFn coef) {
    if (law == LAW_ISO) {          // Isotropic1D.cpp:9-26 / Isotropic3D.cpp:10-27
        const float lam = coef(0), mu = coef(1), mu2 = mu + mu;
        V sii = vscale(vadd(vadd(e[0], e[1]), e[2]), lam);
        s[0] = vadd(sii, vscale(e[0], mu2));
        s[1] = vadd(sii, vscale(e[1], mu2));
        s[2] = vadd(sii, vscale(e[2], mu2));
        s[3] = vscale(e[3], mu);
        s[4] = vscale(e[4], mu);
        s[5] = vscale(e[5], mu);
    } else if (law == LAW_TI) {    // TransverselyIsotropic1D.cpp:9-26 / 3D:10-28
        const float A = coef(0), C = coef(1), F = coef(2), L = coef(3), N = coef(4), N2 = N + N;
        V e01 = vadd(e[0], e[1]);
        V t = vadd(vscale(e01, A), vscale(e[2], F));
        s[0] = vsub(t, vscale(e[1], N2));
        s[1] = vsub(t, vscale(e[0], N2));
        s[2] = vadd(vscale(e[2], C), vscale(e01, F));
        s[3] = vscale(e[3], L);
        s[4] = vscale(e[4], L);
        s[5] = vscale(e[5], N);
    } else {                        // Anisotropic1D.cpp:9-54 / Anisotropic3D.cpp:10-54
        float C[21];
#pragma unroll
        for (int k = 0; k < 21; ++k) C[k] = coef(k);
        // upper triangle, row by row: (0,0..5) -> 0..5, (1,1..5) -> 6..10, (2,2..5) -> 11..14, ...
        const int idx[6][6] = {{0, 1, 2, 3, 4, 5},    {1, 6, 7, 8, 9, 10},   {2, 7, 11, 12, 13, 14},
                               {3, 8, 12, 15, 16, 17}, {4, 9, 13, 16, 18, 19}, {5, 10, 14, 17, 19, 20}};
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            V acc = vscale(e[0], C[idx[i][0]]);
#pragma unroll
            for (int j = 1; j < 6; ++j) acc = vadd(acc, vscale(e[j], C[idx[i][j]]));
            s[i] = acc;
        }
    }
}

// SLS attenuation at one state cell (Attenuation{1D,3D}_{Full,CG4}.cpp): stress -= sum memvar;
// memvar = alpha memvar + beta R_old; R_new from strain; memvar += gamma R_new.
// mem(s, c) / Rst(c) are references into the persistent state.
template <typename V, typename MemFn, typename RFn>
__device__ __forceinline__ void attenuation_cell(int nsls, const float *__restrict__ abg /* [3][nsls] */,
                                                 float dk3, float dmu, float dmu2, bool do_kappa,
                                                 const V (&e)[6], V (&s)[6], MemFn mem, RFn Rst) {
    const float third = (float)(1.0 / 3.0);
    V Rn[6];
    V e3 = vscale(vadd(vadd(e[0], e[1]), e[2]), third);
    if (do_kappa) {
        V s3 = vscale(e3, dk3);
        Rn[0] = vadd(s3, vscale(vsub(e[0], e3), dmu2));
        Rn[1] = vadd(s3, vscale(vsub(e[1], e3), dmu2));
        Rn[2] = vadd(s3, vscale(vsub(e[2], e3), dmu2));
    } else {
        Rn[0] = vscale(vsub(e[0], e3), dmu2);
        Rn[1] = vscale(vsub(e[1], e3), dmu2);
        Rn[2] = vscale(vadd(Rn[0], Rn[1]), -1.f);
    }
    Rn[3] = vscale(e[3], dmu);
    Rn[4] = vscale(e[4], dmu);
    Rn[5] = vscale(e[5], dmu);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        V &Rc = Rst(c);
        V Rold = Rc;
        for (int k = 0; k < nsls; ++k) {
            V &m = mem(k, c);
            V mv = m;
            s[c] = vsub(s[c], mv);
            mv = vadd(vscale(mv, abg[k]), vscale(Rold, abg[nsls + k]));
            mv = vadd(mv, vscale(Rn[c], abg[2 * nsls + k]));
            m = mv;
        }
        Rc = Rn[c];
    }
}
