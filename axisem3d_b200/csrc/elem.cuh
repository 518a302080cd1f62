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
