"""CPU ORACLE (test infrastructure, not product code) -- numpy restatement of the reference's
per-timestep stiffness + Newmark hot path.

PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCES (with a stated caveat).  The reference ships no golden vectors,
tests or fixtures for this path and its real dependencies (Eigen 3.3, FFTW 3.3, MPI, METIS) are
not in this container.  Its hot-path sources do compile unmodified, where they lie, against the plain-loop stand-ins
of oracle/shim/ (Eigen subset, FFTW's plan_many r2c/c2r as direct DFT sums, single-process mpi.h); oracle/Makefile.ref
builds them into oracle/_ref/ref_driver and ref_connectivity, oracle/make_golden*.py run those on seeded synthetic
domains and commit what the reference's classes produced under tests/golden/.  tests/test_golden_reference.py holds
this oracle (fp32 and fp64), oracle/oracle.c and the CUDA path to those vectors (displacements, stiffness forces,
receiver ground motion; rel. L2 <= 2e-6 for the oracles), tests/test_connectivity.py the index maps and halos (bit
exact).  Caveat: third-party arithmetic is the stand-ins', so agreement with a true Eigen/FFTW build is to fp32
rounding (summation order), not bit for bit.  The reference's self-checks restated in tests/test_oracle_invariants.py
(self-adjoint / positive operators, 1D-vs-3D equivalence, round trips, null space) stay as independent evidence.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.

Everything is vectorised over *groups* of elements with the same signature; every function
cites the reference file:line it follows (S/ = /root/reference/SOLVER/src/).
dtype=np.float32 follows the reference's single-precision build (Real=float, S/global.h:12-18),
dtype=np.float64 is the "truth" twin.
"""
from __future__ import annotations

import numpy as np

nPE = 25
CG4_IPNT = (1 * 5 + 1, 1 * 5 + 3, 3 * 5 + 1, 3 * 5 + 3)   # Attenuation1D_CG4.cpp:24-27
ANISO_IJ = [(i, j) for i in range(6) for j in range(i, 6)]


def _cdtype(dtype):
    return np.complex64 if np.dtype(dtype) == np.float32 else np.complex128


# ============================================================================ gradient
class GradOps:
    """Gradient::computeGrad6/Quad6/Grad/Quad (S/core/element/grad/Gradient.cpp) for a batch
    of E elements sharing the axial flag.  Fields are [E, M, ncomp, 5, 5] complex."""

    def __init__(self, G_GLL, G_GLJ, dsdxii, dsdeta, dzdxii, dzdeta, inv_s, axial, dtype):
        rd = np.dtype(dtype)
        self.rd, self.cd = rd, _cdtype(dtype)
        G_GLL = np.asarray(G_GLL, dtype=np.float64).reshape(5, 5).astype(rd)
        G_GLJ = np.asarray(G_GLJ, dtype=np.float64).reshape(5, 5).astype(rd)
        self.axial = bool(axial)
        self.Gxi = G_GLJ if axial else G_GLL          # Gradient.cpp:15-23
        self.Geta = G_GLL
        sh = lambda a: np.asarray(a, dtype=np.float64).reshape(-1, 1, 5, 5).astype(rd)   # [E,1,5,5]
        self.dsdxii, self.dsdeta = sh(dsdxii), sh(dsdeta)
        self.dzdxii, self.dzdeta = sh(dzdxii), sh(dzdeta)
        self.inv_s = sh(inv_s)

    # tensor-product derivatives: GU = Gxi^T u, UG = u Geta
    def _GU(self, u):
        return np.einsum("ki,...kj->...ij", self.Gxi, u).astype(self.cd, copy=False)

    def _UG(self, u):
        return np.einsum("...ik,kj->...ij", u, self.Geta).astype(self.cd, copy=False)

    def _ialpha(self, M):
        return (1j * np.arange(M, dtype=self.rd)).astype(self.cd).reshape(1, M, 1, 1)

    def grad6(self, u, nyquist):
        """Gradient.cpp:206-265.  u [E,M,3,5,5] -> strain [E,M,6,5,5], Voigt order
        [ss, pp, zz, pz, sz, sp] with engineering shear."""
        E, M = u.shape[:2]
        u = u.copy()
        u[:, 0] = u[:, 0].real                          # alpha = 0 uses .real() only (209-224)
        ia = self._ialpha(M)
        u0, u1, u2 = u[:, :, 0], u[:, :, 1], u[:, :, 2]
        GU = [self._GU(c) for c in (u0, u1, u2)]
        UG = [self._UG(c) for c in (u0, u1, u2)]
        ds = [self.dzdeta * GU[c] + self.dzdxii * UG[c] for c in range(3)]
        dz = [self.dsdeta * GU[c] + self.dsdxii * UG[c] for c in range(3)]
        v0 = u0 + ia * u1
        v1 = ia * u0 - u1
        v2 = ia * u2
        e = np.zeros((E, M, 6, 5, 5), dtype=self.cd)
        e[:, :, 0] = ds[0]
        e[:, :, 1] = self.inv_s * v0
        e[:, :, 2] = dz[2]
        e[:, :, 3] = dz[1] + self.inv_s * v2
        e[:, :, 4] = dz[0] + ds[2]
        e[:, :, 5] = ds[1] + self.inv_s * v1
        if self.axial:                                  # 221-224, 245-254
            g0 = self.Gxi[:, 0]                         # (Gxi^T).row(0)
            row = lambda v: np.einsum("k,...kj->...j", g0, v)
            e[:, :, 1, 0, :] += self.dzdeta[:, :, 0, :] * row(v0)
            e[:, :, 5, 0, :] += self.dzdeta[:, :, 0, :] * row(v1)
            e[:, :, 3, 0, :] += self.dzdeta[:, :, 0, :] * row(v2)
            if M > 1:
                e[:, 1, 1, 0, :] += self.dzdxii[:, 0, 0, :] * (v0[:, 1, 0, :] @ self.Geta)
                e[:, 1, 5, 0, :] += self.dzdxii[:, 0, 0, :] * (v1[:, 1, 0, :] @ self.Geta)
        e[:, 0] = e[:, 0].real
        if nyquist:
            e[:, M - 1] = 0                             # 256-264
        return e.astype(self.cd, copy=False)

    def quad6(self, s, nyquist):
        """Gradient.cpp:267-322.  stress [E,M,6,5,5] -> force [E,M,3,5,5]."""
        E, M = s.shape[:2]
        s = s.copy()
        s[:, 0] = s[:, 0].real
        ib = -self._ialpha(M)                           # iibeta = -beta*ii (287)
        g = [s[:, :, 1] + ib * s[:, :, 5],
             ib * s[:, :, 1] - s[:, :, 5],
             ib * s[:, :, 3]]
        pairs = [(0, 4), (5, 3), (4, 2)]
        f = np.zeros((E, M, 3, 5, 5), dtype=self.cd)
        for c, (a, b) in enumerate(pairs):
            X = self.dzdeta * s[:, :, a] + self.dsdeta * s[:, :, b]
            Y = self.dzdxii * s[:, :, a] + self.dsdxii * s[:, :, b]
            f[:, :, c] = (np.einsum("ik,...kj->...ij", self.Gxi, X)
                          + np.einsum("...ik,jk->...ij", Y, self.Geta)
                          + self.inv_s * g[c])
            if self.axial:                              # 282-285, 304-312
                gr = self.dzdeta[:, :, 0, :] * g[c][:, :, 0, :]            # [E,M,5]
                f[:, :, c] += self.Gxi[:, 0].reshape(1, 1, 5, 1) * gr[:, :, None, :]
                if c < 2 and M > 1:
                    f[:, 1, c, 0, :] += (self.dzdxii[:, 0, 0, :] * g[c][:, 1, 0, :]) @ self.Geta.T
        f[:, 0] = f[:, 0].real
        if nyquist:
            f[:, M - 1] = 0                             # 316-321
        return f.astype(self.cd, copy=False)

    def grad9(self, u, nyquist):
        """Gradient::computeGrad9 (Gradient.cpp:84-141): u [E,M,3,5,5] -> u_{i,j} [E,M,9,5,5], component 3 i + j with
        j = (d/ds, (1/s) d/dphi incl. the curvature terms, d/dz)."""
        E, M = u.shape[:2]
        u = u.copy()
        u[:, 0] = u[:, 0].real
        ia = self._ialpha(M)
        uc = [u[:, :, 0], u[:, :, 1], u[:, :, 2]]
        GU = [self._GU(c) for c in uc]
        UG = [self._UG(c) for c in uc]
        v = [uc[0] + ia * uc[1], ia * uc[0] - uc[1], ia * uc[2]]       # v0, v1, v2 (109-111)
        e = np.zeros((E, M, 9, 5, 5), dtype=self.cd)
        for c in range(3):
            e[:, :, 3 * c + 0] = self.dzdeta * GU[c] + self.dzdxii * UG[c]
            e[:, :, 3 * c + 2] = self.dsdeta * GU[c] + self.dsdxii * UG[c]
        e[:, :, 1] = self.inv_s * v[1]
        e[:, :, 4] = self.inv_s * v[0]
        e[:, :, 7] = self.inv_s * v[2]
        if self.axial:                                  # 101-104, 127-136
            g0 = self.Gxi[:, 0]
            row = lambda w: np.einsum("k,...kj->...j", g0, w)
            for comp, w in ((4, v[0]), (1, v[1]), (7, v[2])):
                e[:, :, comp, 0, :] += self.dzdeta[:, :, 0, :] * row(w)
            if M > 1:
                e[:, 1, 4, 0, :] += self.dzdxii[:, 0, 0, :] * (v[0][:, 1, 0, :] @ self.Geta)
                e[:, 1, 1, 0, :] += self.dzdxii[:, 0, 0, :] * (v[1][:, 1, 0, :] @ self.Geta)
        e[:, 0] = e[:, 0].real
        if nyquist:
            e[:, M - 1] = 0
        return e.astype(self.cd, copy=False)

    def quad9(self, s, nyquist):
        """Gradient::computeQuad9 (Gradient.cpp:143-204): [E,M,9,5,5] -> force [E,M,3,5,5]."""
        E, M = s.shape[:2]
        s = s.copy()
        s[:, 0] = s[:, 0].real
        ib = -self._ialpha(M)
        g = [s[:, :, 4] + ib * s[:, :, 1], ib * s[:, :, 4] - s[:, :, 1], ib * s[:, :, 7]]
        f = np.zeros((E, M, 3, 5, 5), dtype=self.cd)
        for c in range(3):
            X = self.dzdeta * s[:, :, 3 * c] + self.dsdeta * s[:, :, 3 * c + 2]
            Y = self.dzdxii * s[:, :, 3 * c] + self.dsdxii * s[:, :, 3 * c + 2]
            f[:, :, c] = (np.einsum("ik,...kj->...ij", self.Gxi, X)
                          + np.einsum("...ik,jk->...ij", Y, self.Geta)
                          + self.inv_s * g[c])
            if self.axial:                              # 158-161, 186-195
                gr = self.dzdeta[:, :, 0, :] * g[c][:, :, 0, :]
                f[:, :, c] += self.Gxi[:, 0].reshape(1, 1, 5, 1) * gr[:, :, None, :]
                if c < 2 and M > 1:
                    f[:, 1, c, 0, :] += (self.dzdxii[:, 0, 0, :] * g[c][:, 1, 0, :]) @ self.Geta.T
        f[:, 0] = f[:, 0].real
        if nyquist:
            f[:, M - 1] = 0
        return f.astype(self.cd, copy=False)

    def grad_fluid(self, u, nyquist):
        """Gradient::computeGrad (26-57).  u [E,M,5,5] -> [E,M,3,5,5]."""
        E, M = u.shape[:2]
        u = u.copy()
        u[:, 0] = u[:, 0].real
        ia = self._ialpha(M)
        GU, UG = self._GU(u), self._UG(u)
        v = ia * u
        e = np.zeros((E, M, 3, 5, 5), dtype=self.cd)
        e[:, :, 0] = self.dzdeta * GU + self.dzdxii * UG
        e[:, :, 1] = self.inv_s * v
        e[:, :, 2] = self.dsdeta * GU + self.dsdxii * UG
        if self.axial:
            e[:, :, 1, 0, :] += self.dzdeta[:, :, 0, :] * np.einsum("k,...kj->...j", self.Gxi[:, 0], v)
        e[:, 0] = e[:, 0].real
        e[:, 0, 1] = 0                                  # u_i[0][1].real().setZero() (33)
        if nyquist:
            e[:, M - 1] = 0
        return e.astype(self.cd, copy=False)

    def quad_fluid(self, s, nyquist):
        """Gradient::computeQuad (59-82).  [E,M,3,5,5] -> [E,M,5,5]."""
        E, M = s.shape[:2]
        s = s.copy()
        s[:, 0] = s[:, 0].real
        ib = -self._ialpha(M)
        g = ib * s[:, :, 1]
        X = self.dzdeta * s[:, :, 0] + self.dsdeta * s[:, :, 2]
        Y = self.dzdxii * s[:, :, 0] + self.dsdxii * s[:, :, 2]
        f = (np.einsum("ik,...kj->...ij", self.Gxi, X) + np.einsum("...ik,jk->...ij", Y, self.Geta)
             + self.inv_s * g)
        if self.axial:
            gr = self.dzdeta[:, :, 0, :] * g[:, :, 0, :]
            f = f + self.Gxi[:, 0].reshape(1, 1, 5, 1) * gr[:, :, None, :]
        f[:, 0] = f[:, 0].real
        if nyquist:
            f[:, M - 1] = 0
        return f.astype(self.cd, copy=False)


# ============================================================================ rotation
def tiso_spz_to_rtz(u, theta, rd):
    """CrdTransTIsoSolid::transformSPZ_RTZ (S/core/element/crd/CrdTransTIsoSolid.cpp:14-27).
    u [E,M,6,5,5] in place; theta [E,5,5] (double, trig evaluated in double then cast)."""
    th = np.asarray(theta, dtype=np.float64).reshape(-1, 1, 5, 5)
    s1, c1 = np.sin(th).astype(rd), np.cos(th).astype(rd)
    s2, c2 = np.sin(2 * th).astype(rd), np.cos(2 * th).astype(rd)
    half = rd.type(0.5)
    sum02 = u[:, :, 0] + u[:, :, 2]
    dif02 = u[:, :, 0] - u[:, :, 2]
    u3 = u[:, :, 3].copy()
    u[:, :, 0] = half * (sum02 + c2 * dif02 - s2 * u[:, :, 4])
    u[:, :, 2] = sum02 - u[:, :, 0]
    u[:, :, 4] = c2 * u[:, :, 4] + s2 * dif02
    u[:, :, 3] = c1 * u3 + s1 * u[:, :, 5]
    u[:, :, 5] = c1 * u[:, :, 5] - s1 * u3
    return u


def tiso_rtz_to_spz(u, theta, rd):
    """CrdTransTIsoSolid::transformRTZ_SPZ (CrdTransTIsoSolid.cpp:29-42)."""
    th = np.asarray(theta, dtype=np.float64).reshape(-1, 1, 5, 5)
    s1, c1 = np.sin(th).astype(rd), np.cos(th).astype(rd)
    s2, c2 = np.sin(2 * th).astype(rd), np.cos(2 * th).astype(rd)
    half = rd.type(0.5)
    sum02 = u[:, :, 0] + u[:, :, 2]
    dif02 = (u[:, :, 0] - u[:, :, 2]) * half
    u3 = u[:, :, 3].copy()
    u[:, :, 0] = half * sum02 + c2 * dif02 + s2 * u[:, :, 4]
    u[:, :, 2] = sum02 - u[:, :, 0]
    u[:, :, 4] = c2 * u[:, :, 4] - s2 * dif02
    u[:, :, 3] = c1 * u3 - s1 * u[:, :, 5]
    u[:, :, 5] = c1 * u[:, :, 5] + s1 * u3
    return u


def _trig(theta, rd):
    th = np.asarray(theta, dtype=np.float64).reshape(-1, 1, 5, 5)
    return np.sin(th).astype(rd), np.cos(th).astype(rd), np.sin(2 * th).astype(rd), np.cos(2 * th).astype(rd)


def tiso9_rotate(u, theta, rd, back):
    """CrdTransTIsoSolid::transformSPZ_RTZ / RTZ_SPZ on 9 components (CrdTransTIsoSolid.cpp:44-83).  u [E,M,9,5,5]."""
    s1, c1, s2, c2 = _trig(theta, rd)
    if back:
        s1, s2 = -s1, -s2
    half = rd.type(0.5)
    sum08, dif08 = u[:, :, 0] + u[:, :, 8], u[:, :, 0] - u[:, :, 8]
    sum26, dif26 = u[:, :, 2] + u[:, :, 6], u[:, :, 2] - u[:, :, 6]
    u1, u3 = u[:, :, 1].copy(), u[:, :, 3].copy()
    u[:, :, 0] = half * (sum08 + c2 * dif08 - s2 * sum26)
    u[:, :, 2] = half * (dif26 + c2 * sum26 + s2 * dif08)
    u[:, :, 6] = u[:, :, 2] - dif26
    u[:, :, 8] = sum08 - u[:, :, 0]
    u[:, :, 1] = c1 * u1 - s1 * u[:, :, 7]
    u[:, :, 7] = c1 * u[:, :, 7] + s1 * u1
    u[:, :, 3] = c1 * u3 - s1 * u[:, :, 5]
    u[:, :, 5] = c1 * u[:, :, 5] + s1 * u3
    return u


def tiso_fluid_rotate(u, theta, rd, back):
    """CrdTransTIsoFluid::transformSPZ_RTZ / RTZ_SPZ on 3 components (CrdTransTIsoFluid.cpp:16-34).  u [E,M,3,5,5]."""
    s1, c1, _, _ = _trig(theta, rd)
    if back:
        s1 = -s1
    u0 = u[:, :, 0].copy()
    u[:, :, 0] = c1 * u0 - s1 * u[:, :, 2]
    u[:, :, 2] = c1 * u[:, :, 2] + s1 * u0
    return u


# ================================================================================= particle relabelling
# X [4, E, rows, 25] (rows = 1: PRT_1D on Fourier coefficients, rows = Nr: PRT_3D on phi samples); fields [E, R, ncomp, 25]
def prt_s2u_solid(sph, X):
    """PRT_1D/3D::sphericalToUndulated(SolidResponse) (PRT_1D.cpp:32-52, PRT_3D.cpp:41-62): 9 -> 6."""
    X0, X1, X2, X3 = X
    c = lambda k: sph[:, :, k]
    und = np.zeros(sph.shape[:2] + (6,) + sph.shape[3:], dtype=sph.dtype)
    und[:, :, 0] = X0 * c(0) + X1 * c(2)
    und[:, :, 1] = X0 * c(4) + X2 * c(5)
    und[:, :, 2] = X3 * c(8)
    und[:, :, 3] = X0 * c(7) + X2 * c(8) + X3 * c(5)
    und[:, :, 4] = X0 * c(6) + X1 * c(8) + X3 * c(2)
    und[:, :, 5] = X0 * (c(3) + c(1)) + X1 * c(5) + X2 * c(2)
    return und


def prt_u2s_solid(und, X):
    """PRT_1D/3D::undulatedToSpherical(SolidResponse) (PRT_1D.cpp:54-76, PRT_3D.cpp:64-86): 6 -> 9."""
    X0, X1, X2, X3 = X
    c = lambda k: und[:, :, k]
    sph = np.zeros(und.shape[:2] + (9,) + und.shape[3:], dtype=und.dtype)
    sph[:, :, 0] = X0 * c(0)
    sph[:, :, 1] = X0 * c(5)
    sph[:, :, 2] = X1 * c(0) + X3 * c(4) + X2 * c(5)
    sph[:, :, 3] = sph[:, :, 1]
    sph[:, :, 4] = X0 * c(1)
    sph[:, :, 5] = X2 * c(1) + X3 * c(3) + X1 * c(5)
    sph[:, :, 6] = X0 * c(4)
    sph[:, :, 7] = X0 * c(3)
    sph[:, :, 8] = X3 * c(2) + X2 * c(3) + X1 * c(4)
    return sph


def prt_s2u_fluid(e, X):
    """PRT_1D/3D::sphericalToUndulated(FluidResponse) (PRT_1D.cpp:9-18, PRT_3D.cpp:21-30): in place on 3 components."""
    X0, X1, X2, X3 = X
    out = np.empty_like(e)
    out[:, :, 0] = X0 * e[:, :, 0] + X1 * e[:, :, 2]
    out[:, :, 1] = X0 * e[:, :, 1] + X2 * e[:, :, 2]
    out[:, :, 2] = X3 * e[:, :, 2]
    return out


def prt_u2s_fluid(s, X):
    """PRT_1D/3D::undulatedToSpherical(FluidResponse) (PRT_1D.cpp:20-30, PRT_3D.cpp:32-39)."""
    X0, X1, X2, X3 = X
    out = np.empty_like(s)
    out[:, :, 2] = X1 * s[:, :, 0] + X2 * s[:, :, 1] + X3 * s[:, :, 2]
    out[:, :, 0] = X0 * s[:, :, 0]
    out[:, :, 1] = X0 * s[:, :, 1]
    return out


# ================================================================================= FFT
def c2r(x, Nr, rd):
    """FieldFFT::transformF2P + SolverFFTW_N*::computeC2R (FieldFFT.cpp:11-33,
    SolverFFTW_N6.cpp:51-53): backward, unnormalised.  x [E, M, ...] -> [E, Nr, ...].
    numpy's pocketfft c2r ignores Im(DC) and Im(Nyquist) exactly as FFTW's c2r does."""
    y = np.fft.irfft(x, n=Nr, axis=1) * Nr
    return y.astype(rd, copy=False)


def r2c(x, Nr, cd):
    """FieldFFT::transformP2F + computeR2C (FieldFFT.cpp:35-57, SolverFFTW_N6.cpp:45-49):
    forward, then scaled by 1/Nr.  x [E, Nr, ...] -> [E, Nr/2+1, ...]."""
    y = np.fft.rfft(x, axis=1)
    y = y * x.dtype.type(1.0 / Nr) if x.dtype == np.float32 else y / Nr
    return y.astype(cd, copy=False)


# ====================================================================== constitutive laws
def stress_iso(e, lam, mu):
    """Isotropic1D.cpp:9-26 / Isotropic3D.cpp:10-27.  e [...,6,P] (component axis = -2),
    lam/mu broadcastable to [...,P]."""
    mu2 = mu + mu                                       # mMu2(two * mu), Isotropic1D.h:15
    sii = lam * (e[..., 0, :] + e[..., 1, :] + e[..., 2, :])
    s = np.empty_like(e)
    s[..., 0, :] = sii + mu2 * e[..., 0, :]
    s[..., 1, :] = sii + mu2 * e[..., 1, :]
    s[..., 2, :] = sii + mu2 * e[..., 2, :]
    s[..., 3, :] = mu * e[..., 3, :]
    s[..., 4, :] = mu * e[..., 4, :]
    s[..., 5, :] = mu * e[..., 5, :]
    return s


def stress_ti(e, A, C, F, L, N):
    """TransverselyIsotropic1D.cpp:9-26 / TransverselyIsotropic3D.cpp:10-28 (strain in RTZ)."""
    N2 = N + N
    e01 = e[..., 0, :] + e[..., 1, :]
    t = A * e01 + F * e[..., 2, :]
    s = np.empty_like(e)
    s[..., 0, :] = t - N2 * e[..., 1, :]
    s[..., 1, :] = t - N2 * e[..., 0, :]
    s[..., 2, :] = C * e[..., 2, :] + F * e01
    s[..., 3, :] = L * e[..., 3, :]
    s[..., 4, :] = L * e[..., 4, :]
    s[..., 5, :] = N * e[..., 5, :]
    return s


def stress_aniso(e, C21):
    """Anisotropic1D.cpp:9-54 / Anisotropic3D.cpp:10-54: sigma = C eps, symmetric 6x6,
    C21[k] in ANISO_IJ order."""
    s = np.zeros_like(e)
    for k, (i, j) in enumerate(ANISO_IJ):
        s[..., i, :] += C21[k] * e[..., j, :]
        if i != j:
            s[..., j, :] += C21[k] * e[..., i, :]
    return s


class AttState:
    """SLS memory variables of one element group (Attenuation{1D,3D}_{Full,CG4}.cpp).
    State arrays: memvar [nsls, E, R, 6, P], stressR [E, R, 6, P]; R = M (1D, complex) or
    Nr (3D, real); P = 25 (Full) or 4 (CG4)."""

    def __init__(self, atts, R, is3D, rd):
        a0 = atts[0]
        self.nsls, self.cg4, self.doKappa = a0.nsls, a0.cg4, a0.doKappa
        for a in atts:
            if (a.nsls, a.cg4, a.doKappa) != (self.nsls, self.cg4, self.doKappa):
                raise ValueError("inhomogeneous attenuation group")
        E = len(atts)
        P = 4 if self.cg4 else nPE
        self.P, self.rd = P, rd
        dt = rd if is3D else _cdtype(rd)
        rows = R if is3D else 1
        self.alpha = np.stack([a.alpha for a in atts]).astype(rd)          # [E, nsls]
        self.beta = np.stack([a.beta for a in atts]).astype(rd)
        self.gamma = np.stack([a.gamma for a in atts]).astype(rd)
        dk = np.stack([np.asarray(a.dkappa).reshape(rows, P) for a in atts]).astype(rd)
        dm = np.stack([np.asarray(a.dmu).reshape(rows, P) for a in atts]).astype(rd)
        three, two = rd.type(3.0), rd.type(2.0)
        self.dk3, self.dmu, self.dmu2 = three * dk, dm, two * dm          # ctor: three*dkappa, two*dmu
        self.memvar = np.zeros((self.nsls, E, R, 6, P), dtype=dt)
        self.stressR = np.zeros((E, R, 6, P), dtype=dt)
        self.sel = list(CG4_IPNT) if self.cg4 else slice(None)

    def reset(self):
        self.memvar[:] = 0
        self.stressR[:] = 0

    def apply_and_update(self, strain, stress):
        """strain/stress [E, R, 6, 25]; stress is modified in place.
        applyToStress then updateMemoryVariables (e.g. Attenuation3D_Full.cpp:16-50)."""
        rd = self.rd
        for s in range(self.nsls):                       # sigma -= sum memvar
            stress[..., self.sel] -= self.memvar[s]
        a = self.alpha.T.reshape(self.nsls, -1, 1, 1, 1)
        b = self.beta.T.reshape(self.nsls, -1, 1, 1, 1)
        g = self.gamma.T.reshape(self.nsls, -1, 1, 1, 1)
        self.memvar = a * self.memvar + b * self.stressR[None]
        e = strain[..., self.sel]
        third = rd.type(1.0 / 3.0)
        e3 = (e[..., 0, :] + e[..., 1, :] + e[..., 2, :]) * third
        R = self.stressR
        if self.doKappa:
            s3 = self.dk3 * e3
            R[..., 0, :] = s3 + self.dmu2 * (e[..., 0, :] - e3)
            R[..., 1, :] = s3 + self.dmu2 * (e[..., 1, :] - e3)
            R[..., 2, :] = s3 + self.dmu2 * (e[..., 2, :] - e3)
        else:
            R[..., 0, :] = self.dmu2 * (e[..., 0, :] - e3)
            R[..., 1, :] = self.dmu2 * (e[..., 1, :] - e3)
            R[..., 2, :] = -(R[..., 0, :] + R[..., 1, :])
        R[..., 3, :] = self.dmu * e[..., 3, :]
        R[..., 4, :] = self.dmu * e[..., 4, :]
        R[..., 5, :] = self.dmu * e[..., 5, :]
        self.memvar = (self.memvar + g * R[None]).astype(self.memvar.dtype, copy=False)


# ======================================================================== element groups
class _Group:
    pass


class OracleDomain:
    """Restates Domain (S/core/domain/Domain.cpp) + the Point/Element classes it owns.
    Build with the descriptor objects of axisem3d_b200.model (duck-typed; not imported)."""

    def __init__(self, dtype=np.float32):
        self.rd = np.dtype(dtype)
        self.cd = _cdtype(dtype)
        self.points, self.elements, self.sources = [], [], []
        self.G_GLL = self.G_GLJ = None
        self.msg = None
        self.exchange = None
        self._final = False

    # ---- construction (Domain.cpp:46-56)
    def setGMat(self, G_GLL, G_GLJ):
        self.G_GLL, self.G_GLJ = np.asarray(G_GLL, float).reshape(5, 5), np.asarray(G_GLJ, float).reshape(5, 5)

    def addPoint(self, p):
        p.domain_tag = len(self.points)
        self.points.append(p)
        return p.domain_tag

    def addElement(self, e):
        e.domain_tag = len(self.elements)
        self.elements.append(e)
        return e.domain_tag

    def addSourceTerm(self, st):
        self.sources.append(st)

    def setMessaging(self, info, exchange):
        """exchange(list_of_send_buffers) -> list_of_recv_buffers (one per neighbour)."""
        self.msg, self.exchange = info, exchange

    # ---- finalize: flatten points, group elements
    def finalize(self):
        rd, cd = self.rd, self.cd
        P = self.points
        nP = len(P)
        self.p_nr = np.array([p.nr for p in P], dtype=np.int64)
        self.p_nu = self.p_nr // 2
        self.p_nyq = (self.p_nr % 2 == 0).astype(np.int64)
        self.p_axial = np.array([p.axial for p in P], dtype=bool)
        self.Mmax = int(self.p_nu.max()) + 1 if nP else 1
        self.s_idx = -np.ones(nP, dtype=np.int64)
        self.f_idx = -np.ones(nP, dtype=np.int64)
        s_pts, f_pts = [], []
        for t, p in enumerate(P):
            if p.kind in ("solid", "solidfluid"):
                self.s_idx[t] = len(s_pts)
                s_pts.append((t, p if p.kind == "solid" else p.solid))
            if p.kind in ("fluid", "solidfluid"):
                self.f_idx[t] = len(f_pts)
                f_pts.append((t, p if p.kind == "fluid" else p.fluid))
        self.s_tag = np.array([t for t, _ in s_pts], dtype=np.int64)
        self.f_tag = np.array([t for t, _ in f_pts], dtype=np.int64)
        self.s_mass = [q.mass for _, q in s_pts]
        self.f_mass = [q.mass for _, q in f_pts]
        self.f_surf = np.array([q.fluidSurf for _, q in f_pts], dtype=bool)
        nS, nF, M = len(s_pts), len(f_pts), self.Mmax
        self.S = {k: np.zeros((nS, 3, M), dtype=cd) for k in ("displ", "veloc", "accel", "stiff")}
        self.F = {k: np.zeros((nF, M), dtype=cd) for k in ("displ", "veloc", "accel", "stiff")}
        al = np.arange(M)
        # rows alpha <= Nu_p - nyq_p are "live" (SolidPoint.cpp:175-209)
        self.s_live = al[None, :] <= (self.p_nu - self.p_nyq)[self.s_tag][:, None]
        self.f_live = al[None, :] <= (self.p_nu - self.p_nyq)[self.f_tag][:, None]
        self.s_rows = al[None, :] <= self.p_nu[self.s_tag][:, None]
        self.f_rows = al[None, :] <= self.p_nu[self.f_tag][:, None]
        self.sf_tags = [t for t, p in enumerate(P) if p.kind == "solidfluid"]
        self._build_groups()
        self._final = True

    def _sig(self, e):
        if e.kind == "solid":
            el = e.elastic
            att = el.att
            asig = None if att is None else (att.nsls, att.cg4, att.doKappa)
            return ("solid", e.maxNr, e.axial(), el.law, el.is3D, asig, e.prt is not None)
        return ("fluid", e.maxNr, e.axial(), e.acoustic.is3D, e.prt is not None)

    def _build_groups(self):
        rd = self.rd
        by = {}
        for e in self.elements:
            by.setdefault(self._sig(e), []).append(e)
        self.groups = []
        for sig, els in by.items():
            g = _Group()
            g.sig, g.kind, g.Nr, g.axial = sig, sig[0], sig[1], sig[2]
            g.Nu = g.Nr // 2
            g.M = g.Nu + 1
            g.nyq = int(g.Nr % 2 == 0)
            g.tags = np.array([e.domain_tag for e in els], dtype=np.int64)
            idx = self.s_idx if g.kind == "solid" else self.f_idx
            g.pidx = np.array([[idx[p.domain_tag] for p in e.points] for e in els], dtype=np.int64)
            if (g.pidx < 0).any():
                raise RuntimeError("Point::scatterDisplToElement || Incompatible point type.")
            gr = [e.grad for e in els]
            g.grad = GradOps(self.G_GLL, self.G_GLJ,
                             np.stack([x.dsdxii for x in gr]), np.stack([x.dsdeta for x in gr]),
                             np.stack([x.dzdxii for x in gr]), np.stack([x.dzdeta for x in gr]),
                             np.stack([x.inv_s for x in gr]), g.axial, rd)
            g.elem3D = els[0].elem3D
            rows = g.Nr if g.elem3D else 1
            g.hasPRT = els[0].prt is not None
            g.X = np.stack([e.prt.X for e in els], axis=1).astype(rd) if g.hasPRT else None      # [4, E, rows, 25]
            if g.kind == "solid":
                g.law = sig[3]
                g.inTIso = els[0].inTIso
                g.theta = np.stack([e.formThetaMat() for e in els]) if g.inTIso else None
                # coef [ncoef, E, rows, 25]
                g.coef = np.stack([e.elastic.coef for e in els], axis=1).astype(rd)
                atts = [e.elastic.att for e in els]
                g.att = None if atts[0] is None else AttState(atts, g.Nr if g.elem3D else g.M, g.elem3D, rd)
            else:
                g.K = np.stack([e.acoustic.K.reshape(rows, nPE) for e in els]).astype(rd)
                g.inTIso = g.hasPRT
                g.theta = np.stack([e.formThetaMat() for e in els]) if g.hasPRT else None
            self.groups.append(g)
        # keep reference order of first appearance irrelevant: scatter is a sum

    # ---- element level ------------------------------------------------------------
    def _gather_solid(self, g):
        """SolidPoint::scatterDisplToElement (SolidPoint.cpp:175-195) for all 25 points."""
        M = g.M
        d = self.S["displ"][g.pidx]                       # [E,25,3,Mmax]
        live = self.s_live[g.pidx]                        # [E,25,Mmax]
        d = np.where(live[:, :, None, :], d, 0)[..., :M]
        if M > d.shape[-1]:
            raise AssertionError
        u = np.transpose(d, (0, 3, 2, 1)).reshape(len(g.tags), M, 3, 5, 5)
        return np.ascontiguousarray(u)

    def _gather_fluid(self, g):
        M = g.M
        d = self.F["displ"][g.pidx]                       # [E,25,Mmax]
        d = np.where(self.f_live[g.pidx], d, 0)[..., :M]
        return np.ascontiguousarray(np.transpose(d, (0, 2, 1)).reshape(len(g.tags), M, 5, 5))

    def solid_displ_to_stiff(self, g, u):
        """SolidElement::displToStiff without PRT (SolidElement.cpp:404-443)."""
        rd, cd = self.rd, self.cd
        E, M = u.shape[:2]
        if g.hasPRT:                                                       # SolidElement.cpp:405-413, 424-432
            e9 = tiso9_rotate(g.grad.grad9(u, g.nyq), g.theta, rd, False).reshape(E, M, 9, nPE)
            if g.elem3D:
                und = prt_s2u_solid(c2r(e9, g.Nr, rd), g.X).astype(rd, copy=False)
                sph = prt_u2s_solid(self._stress(g, und), g.X).astype(rd, copy=False)
                s9 = r2c(sph, g.Nr, cd)
            else:
                und = prt_s2u_solid(e9, g.X).astype(cd, copy=False)
                s9 = prt_u2s_solid(self._stress(g, und), g.X).astype(cd, copy=False)
            s9 = tiso9_rotate(np.ascontiguousarray(s9).reshape(E, M, 9, 5, 5), g.theta, rd, True)
            return g.grad.quad9(s9, g.nyq)
        e = g.grad.grad6(u, g.nyq)
        if g.inTIso:
            e = tiso_spz_to_rtz(e, g.theta, rd)
        if g.elem3D:
            eR = c2r(e.reshape(E, M, 6, nPE), g.Nr, rd)                  # [E,Nr,6,25]
            sR = self._stress(g, eR)
            s = r2c(sR, g.Nr, cd).reshape(E, M, 6, 5, 5)
        else:
            ef = e.reshape(E, M, 6, nPE)
            s = self._stress(g, ef).reshape(E, M, 6, 5, 5)
        if g.inTIso:
            s = tiso_rtz_to_spz(s, g.theta, rd)
        return g.grad.quad6(s, g.nyq)

    def _stress(self, g, e):
        c = g.coef                                         # [ncoef,E,rows,25]
        if g.law == "iso":
            s = stress_iso(e, c[0], c[1])
        elif g.law == "ti":
            s = stress_ti(e, c[0], c[1], c[2], c[3], c[4])
        else:
            s = stress_aniso(e, c)
        s = s.astype(e.dtype, copy=False)
        if g.att is not None:
            g.att.apply_and_update(e, s)
        return s

    def fluid_displ_to_stiff(self, g, u):
        """FluidElement::displToStiff (FluidElement.cpp:333-355)."""
        return g.grad.quad_fluid(self._fluid_stress(g, u), g.nyq)

    def _fluid_stress(self, g, u):
        """gather -> computeGrad -> [rotate] -> [c2r] -> [PRT] -> K -> [PRT] -> [r2c] -> [rotate back]: the part shared by
        FluidElement::displToStiff and FluidElement::computeGroundMotion (FluidElement.cpp:163-215).  -> [E,M,3,5,5]"""
        rd, cd = self.rd, self.cd
        E, M = u.shape[:2]
        e = g.grad.grad_fluid(u, g.nyq)
        K = g.K[:, :, None, :]                             # [E,rows,1,25]
        if g.hasPRT:
            e = tiso_fluid_rotate(e, g.theta, rd, False)   # FluidElement.cpp:335-337
        if g.elem3D:
            eR = c2r(e.reshape(E, M, 3, nPE), g.Nr, rd)
            if g.hasPRT:
                eR = prt_s2u_fluid(eR, g.X).astype(rd, copy=False)
            sR = (K * eR).astype(rd, copy=False)           # Acoustic3D.cpp:9-16
            if g.hasPRT:
                sR = prt_u2s_fluid(sR, g.X).astype(rd, copy=False)
            s = r2c(sR, g.Nr, cd).reshape(E, M, 3, 5, 5)
        else:
            ef = e.reshape(E, M, 3, nPE)
            if g.hasPRT:
                ef = prt_s2u_fluid(ef, g.X).astype(cd, copy=False)
            sf = (K * ef).astype(cd, copy=False)
            if g.hasPRT:
                sf = prt_u2s_fluid(sf, g.X).astype(cd, copy=False)
            s = np.ascontiguousarray(sf).reshape(E, M, 3, 5, 5)
        if g.hasPRT:
            s = tiso_fluid_rotate(s, g.theta, rd, True)    # FluidElement.cpp:350-352
        return s

    def computeStiff(self):
        """Domain::computeStiff (Domain.cpp:82-94) -> Element::computeStiff
        (SolidElement.cpp:43-65; FluidElement.cpp:43-65): gather, displToStiff, stiff -= f
        truncated to every point's own live rows (SolidPoint.cpp:197-209)."""
        for g in self.groups:
            E, M = len(g.tags), g.M
            if g.kind == "solid":
                f = self.solid_displ_to_stiff(g, self._gather_solid(g))           # [E,M,3,5,5]
                f = np.transpose(f.reshape(E, M, 3, nPE), (0, 3, 2, 1))           # [E,25,3,M]
                live = self.s_live[g.pidx][..., :M]                               # [E,25,M]
                f = np.where(live[:, :, None, :], f, 0)
                stiff = self.S["stiff"]
                np.subtract.at(stiff, (g.pidx.reshape(-1), slice(None), slice(0, M)),
                               f.reshape(E * nPE, 3, M).astype(self.cd))
                pts = np.unique(g.pidx)
            else:
                f = self.fluid_displ_to_stiff(g, self._gather_fluid(g))           # [E,M,5,5]
                f = np.transpose(f.reshape(E, M, nPE), (0, 2, 1))                 # [E,25,M]
                f = np.where(self.f_live[g.pidx][..., :M], f, 0)
                np.subtract.at(self.F["stiff"], (g.pidx.reshape(-1), slice(0, M)),
                               f.reshape(E * nPE, M).astype(self.cd))
                pts = np.unique(g.pidx)
            # "mask Nyquist": the point's Nyquist row of stiff is set to zero (206-208)
            self._zero_point_nyquist(g.kind, pts)

    def _zero_point_nyquist(self, kind, pts):
        if kind == "solid":
            tags = self.s_tag[pts]
            sel = self.p_nyq[tags] == 1
            self.S["stiff"][pts[sel], :, self.p_nu[tags][sel]] = 0
        else:
            tags = self.f_tag[pts]
            sel = self.p_nyq[tags] == 1
            self.F["stiff"][pts[sel], self.p_nu[tags][sel]] = 0

    # ---- point level --------------------------------------------------------------
    def _mask_solid(self, f):
        """SolidPoint::maskField (SolidPoint.cpp:216-238) on [nS,3,M]."""
        cd = self.cd
        half = self.rd.type(0.5)
        f[:, :, 0] = f[:, :, 0].real
        ax = self.p_axial[self.s_tag]
        if ax.any():
            a = np.nonzero(ax)[0]
            f[a, 0, 0] = 0
            f[a, 1, 0] = 0
            if f.shape[2] > 1:
                has1 = a[self.p_nu[self.s_tag][a] >= 1]
                s0 = f[has1, 0, 1].copy()
                s1 = f[has1, 1, 1].copy()
                f[has1, 0, 1] = half * (s0 - cd(1j) * s1)
                f[has1, 1, 1] = half * (s1 + cd(1j) * s0)
                f[has1, 2, 1] = 0
                f[has1, :, 2:] = 0
        tags = self.s_tag
        sel = np.nonzero(self.p_nyq[tags] == 1)[0]
        f[sel, :, self.p_nu[tags][sel]] = 0
        return f

    def _mask_fluid(self, f):
        """FluidPoint::maskField (FluidPoint.cpp:197-207) on [nF,M]."""
        f[:, 0] = f[:, 0].real
        ax = np.nonzero(self.p_axial[self.f_tag])[0]
        f[ax, 1:] = 0
        tags = self.f_tag
        sel = np.nonzero(self.p_nyq[tags] == 1)[0]
        f[sel, self.p_nu[tags][sel]] = 0
        return f

    def _accel_full(self, stiff, masses, tags, comps):
        """Mass1D::computeAccel (Mass1D.cpp:12-18) / Mass3D::computeAccel (Mass3D.cpp:13-57)."""
        rd, cd = self.rd, self.cd
        out = stiff
        ocean = np.array([bool(getattr(m, "ocean", False)) for m in masses], dtype=bool)
        scal = np.array([(not m.is3D) for m in masses], dtype=bool) & ~ocean
        inv = np.array([m.invMass if (not m.is3D and not getattr(m, "ocean", False)) else 1.0 for m in masses], dtype=np.float64).astype(rd)
        for i in np.nonzero(ocean)[0]:
            m = masses[i]
            if not m.is3D:        # MassOcean1D::computeAccel (MassOcean1D.cpp:15-28)
                imZ, imR = rd.type(1.0 / (m.mass + m.massOcean)), rd.type(1.0 / m.mass)
                st, ct = rd.type(np.sin(m.theta)), rd.type(np.cos(m.theta))
                Z = (out[i, 0] * st + out[i, 2] * ct) * imZ
                R = (out[i, 0] * ct - out[i, 2] * st) * imR
                out[i, 0] = Z * st + R * ct
                out[i, 2] = Z * ct - R * st
                out[i, 1] = out[i, 1] * imR
            else:                 # MassOcean3D::computeAccel (MassOcean3D.cpp:18-49)
                Nr = int(self.p_nr[tags[i]])
                Nc = Nr // 2 + 1
                im = (1.0 / m.mass).astype(rd)
                sc = np.sqrt(m.massOcean / (m.mass * (m.mass + m.massOcean))).astype(rd)
                ns = (m.normal.astype(rd) * sc[:, None]).astype(rd).T               # [3, Nr]
                xr = (np.fft.irfft(out[i, :, :Nc], n=Nr, axis=-1) * Nr).astype(rd)    # [3, Nr]
                fn = (xr * ns).sum(axis=0).astype(rd)
                yr = (xr * im[None, :] - fn[None, :] * ns).astype(rd)
                y = np.fft.rfft(yr, axis=-1)
                out[i, :, :Nc] = (y * rd.type(1.0 / Nr)).astype(cd) if rd == np.float32 else (y / Nr)
        if comps:
            out[scal] = out[scal] * inv[scal][:, None, None]
        else:
            out[scal] = out[scal] * inv[scal][:, None]
        for i in np.nonzero(~scal & ~ocean)[0]:
            m = masses[i]
            Nr = int(self.p_nr[tags[i]])
            Nc = Nr // 2 + 1
            im = m.invMass.astype(rd)
            x = out[i, ..., :Nc]
            xr = (np.fft.irfft(x, n=Nr, axis=-1) * Nr).astype(rd)
            yr = (xr * im).astype(rd)
            y = np.fft.rfft(yr, axis=-1)
            y = (y * rd.type(1.0 / Nr)).astype(cd) if rd == np.float32 else (y / Nr)
            out[i, ..., :Nc] = y
        return out

    def updateNewmark(self, dt):
        """Domain::updateNewmark (Domain.cpp:165-177) -> SolidPoint::updateNewmark
        (SolidPoint.cpp:23-38), FluidPoint::updateNewmark (FluidPoint.cpp:23-44)."""
        rd = self.rd
        half_dt = 0.5 * dt
        half_dt_dt = half_dt * dt
        c_hdt, c_dt, c_hdd = rd.type(half_dt), rd.type(dt), rd.type(half_dt_dt)
        for fld, mask, masses, tags, comps, rows in (
                (self.S, self._mask_solid, self.s_mass, self.s_tag, True, self.s_rows),
                (self.F, self._mask_fluid, self.f_mass, self.f_tag, False, self.f_rows)):
            if len(tags) == 0:
                continue
            st = mask(fld["stiff"])
            st = self._accel_full(st, masses, tags, comps)
            st = mask(st)
            r = rows[:, None, :] if comps else rows
            st = np.where(r, st, 0).astype(self.cd, copy=False)
            fld["veloc"] += c_hdt * (fld["accel"] + st)
            fld["accel"][:] = st
            fld["displ"] += c_dt * fld["veloc"] + c_hdd * fld["accel"]
            fld["stiff"][:] = 0
        if len(self.f_tag) and self.f_surf.any():               # FluidPoint.cpp:25-28
            for k in ("displ", "veloc", "accel", "stiff"):
                self.F[k][self.f_surf] = 0

    def applySource(self, stf):
        """Domain::applySource (Domain.cpp:96-109) -> SourceTerm::apply (SourceTerm.cpp:29-35)
        -> SolidPoint::addToStiff (SolidPoint.cpp:211-214)."""
        stf = self.rd.type(stf)
        for st in self.sources:
            for i, p in enumerate(st.element.points):
                si = self.s_idx[p.domain_tag]
                if si < 0:
                    raise RuntimeError("Point::addToStiff || Incompatible point type.")
                f = st.force[i].astype(self.cd)
                self.S["stiff"][si, :, :f.shape[0]] += (f * stf).T

    def coupleSolidFluid(self):
        """Domain::coupleSolidFluid (Domain.cpp:179-191) -> SolidFluidPoint::coupleSolidFluid
        (SolidFluidPoint.cpp:106-110): fluid first, then solid (order matters)."""
        rd, cd = self.rd, self.cd
        for t in self.sf_tags:
            p = self.points[t]
            si, fi = self.s_idx[t], self.f_idx[t]
            c = p.couple
            n = p.nu + 1
            us = self.S["displ"][si, :, :n]
            fs = self.S["stiff"][si, :, :n]
            ff = self.F["stiff"][fi, :n]
            if not c.is3D:                                   # SFCoupling1D.cpp:9-19
                ff += rd.type(c.ns) * us[0] + rd.type(c.nz) * us[2]
                fs[0] -= rd.type(c.ns_invmf) * ff
                fs[2] -= rd.type(c.nz_invmf) * ff
            else:                                            # SFCoupling3D.cpp:9-54
                Nr = p.nr
                ur = (np.fft.irfft(us, n=Nr, axis=-1) * Nr).astype(rd)          # [3,Nr]
                n_un = c.n_un.astype(rd).T
                fr = (n_un[0] * ur[0] + n_un[1] * ur[1] + n_un[2] * ur[2]).astype(rd)
                add = np.fft.rfft(fr)
                add = (add * rd.type(1.0 / Nr)).astype(cd) if rd == np.float32 else add / Nr
                ff += add
                fr2 = (np.fft.irfft(ff, n=Nr) * Nr).astype(rd)
                n_as = c.n_as.astype(rd).T
                sr = (n_as * fr2[None, :]).astype(rd)
                sub = np.fft.rfft(sr, axis=-1)
                sub = (sub * rd.type(1.0 / Nr)).astype(cd) if rd == np.float32 else sub / Nr
                fs -= sub

    # ---- halo (Domain::assembleStiff, Domain.cpp:111-163) --------------------------
    def _pack(self, tags):
        out = []
        for t in tags:
            p = self.points[t]
            n = p.nu + 1
            if self.s_idx[t] >= 0:
                out.append(self.S["stiff"][self.s_idx[t], :, :n].reshape(-1))   # col-major CMatX3
            if self.f_idx[t] >= 0:
                out.append(self.F["stiff"][self.f_idx[t], :n])
        return np.concatenate(out) if out else np.zeros(0, dtype=self.cd)

    def _unpack_add(self, tags, buf):
        row = 0
        for t in tags:
            p = self.points[t]
            n = p.nu + 1
            if self.s_idx[t] >= 0:
                self.S["stiff"][self.s_idx[t], :, :n] += buf[row:row + 3 * n].reshape(3, n)
                row += 3 * n
            if self.f_idx[t] >= 0:
                self.F["stiff"][self.f_idx[t], :n] += buf[row:row + n]
                row += n

    def assembleStiff(self, phase=0):
        if self.msg is None or self.msg.mNProcComm == 0:
            return
        if phase <= 0:
            self._send = [self._pack(l) for l in self.msg.mILocalPoints]
            self._recv = self.exchange(self._send)
        if phase >= 0:
            for l, b in zip(self.msg.mILocalPoints, self._recv):
                self._unpack_add(l, np.asarray(b, dtype=self.cd))

    def checkStability(self):
        return bool(np.isfinite(self.S["displ"]).all() and np.isfinite(self.F["displ"]).all())

    def resetZero(self):
        for k in self.S:
            self.S[k][:] = 0
            self.F[k][:] = 0
        for g in self.groups:
            if g.kind == "solid" and g.att is not None:
                g.att.reset()

    # ---- step in the reference order (Newmark.cpp:47-93) ---------------------------
    def step(self, dt, stf):
        self.updateNewmark(dt)
        self.applySource(stf)
        self.computeStiff()
        self.coupleSolidFluid()
        self.assembleStiff(-1)
        self.assembleStiff(1)

    # ---- field access (by domain point tag) ----------------------------------------
    def get_solid(self, tag, which):
        n = self.points[tag].nu + 1
        return self.S[which][self.s_idx[tag], :, :n].T.copy()      # (Nu+1, 3) like CMatX3

    def get_fluid(self, tag, which):
        n = self.points[tag].nu + 1
        return self.F[which][self.f_idx[tag], :n].copy()

    def set_solid(self, tag, which, val):
        n = self.points[tag].nu + 1
        self.S[which][self.s_idx[tag], :, :n] = np.asarray(val).reshape(n, 3).T

    def set_fluid(self, tag, which, val):
        n = self.points[tag].nu + 1
        self.F[which][self.f_idx[tag], :n] = np.asarray(val).reshape(n)

    def maskDispl(self):
        """mask as randomDispl does (SolidPoint.cpp:47-58)."""
        self._mask_solid(self.S["displ"])
        self._mask_fluid(self.F["displ"])

    def ground_motion(self, elem_tag, phi, weights):
        """SolidElement::computeGroundMotion (SolidElement.cpp:189-216)."""
        e = self.elements[elem_tag]
        w = np.asarray(weights, float).reshape(nPE)
        if e.kind == "fluid":
            return self._ground_motion_fluid(e, phi, w)
        out = np.zeros(3)
        Nu, Nr = e.maxNu, e.maxNr
        top = Nu - int(Nr % 2 == 0)
        for i, p in enumerate(e.points):
            if abs(w[i]) < 1e-10:
                continue
            si = self.s_idx[p.domain_tag]
            d = np.where(self.s_live[si][None, :], self.S["displ"][si], 0)[:, :top + 1]
            al = np.arange(1, top + 1)
            ex = 2.0 * np.exp(1j * al * phi)
            up = d[:, 0].real + (ex[None, :] * d[:, 1:top + 1]).real.sum(axis=1)
            out += w[i] * up
        return out

    # ---- wisdom learning (Domain::learnWisdom, Domain.cpp:384-402) ----------------------
    def learnWisdom(self, cutoff):
        """Point::learnWisdom on every point (SolidPoint.cpp:240-266, FluidPoint.cpp:209-229): per component, when the
        Hilbert norm of the displacement exceeds its running maximum, the smallest order whose truncation error stays
        below cutoff^2 of it is remembered."""
        if not hasattr(self, "_wis"):
            self._wis = {}
        rd = np.dtype(self.rd).type
        for t, p in enumerate(self.points):
            n = p.nu + 1
            cols = []
            if self.s_idx[t] >= 0:
                cols += [("s", c, self.S["displ"][self.s_idx[t], c, :n]) for c in range(3)]
            if self.f_idx[t] >= 0:
                cols.append(("f", 0, self.F["displ"][self.f_idx[t], :n]))
            for fam, c, u in cols:
                e = (u.real.astype(rd) ** 2 + u.imag.astype(rd) ** 2).astype(rd)
                l2 = rd(e.sum(dtype=rd))
                h2 = rd(l2 - rd(0.5) * e[0])
                st = self._wis.setdefault((t, fam, c), [rd(-1.0), p.nu])
                if h2 <= st[0]:
                    continue
                st[0] = h2
                tol = rd(h2 * rd(cutoff) * rd(cutoff))
                diff = l2 - np.cumsum(e[:-1], dtype=rd)            # newNu = 0 .. nu-1
                hit = np.nonzero(diff <= tol)[0]
                st[1] = int(hit[0]) if hit.size else p.nu

    def getNuWisdom(self):
        """Point::getNuWisdom per point (SolidPoint.cpp:268-272: max over the 3 components; SolidFluidPoint.cpp:129-131)."""
        out = np.array([p.nu for p in self.points], dtype=np.int32)
        if hasattr(self, "_wis"):
            best = {}
            for (t, fam, c), st in self._wis.items():
                best[t] = max(best.get(t, 0), st[1])
            for t, v in best.items():
                out[t] = v
        return out

    def _solid_group_of(self, e):
        return next((g, int(np.nonzero(g.tags == e.domain_tag)[0][0])) for g in self.groups
                    if g.kind == "solid" and (g.tags == e.domain_tag).any())

    def _eval_phi(self, s, g, phi, w):
        """sum_p w_p (s_0 + 2 Re sum_{alpha >= 1} s_alpha e^{i alpha phi}) for s [M, ncomp, 25]"""
        top = g.Nu - g.nyq
        ex = 2.0 * np.exp(1j * np.arange(1, top + 1) * phi)
        up = s[0].real + (ex[:, None, None] * s[1:top + 1]).real.sum(axis=0)
        return (up * np.where(np.abs(w) < 1e-10, 0.0, w)[None, :]).sum(axis=1)

    def strain(self, elem_tag, phi, weights):
        """SolidElement::computeStrain (SolidElement.cpp:219-279) after forceTIso, no PRT: grad6 -> SPZ_RTZ -> evaluation.
        Fluid elements (FluidElement::computeStrain, FluidElement.cpp:219-306, no PRT): the fluid displacement (acoustic stress of
        the potential, as in computeGroundMotion) goes through the same computeGrad6 -> transformSPZ_RTZ -> evaluation."""
        e = self.elements[elem_tag]
        if e.prt is not None:
            raise NotImplementedError("strain receivers: elements without PRT")
        if e.kind == "fluid":
            g, k = next((g, int(np.nonzero(g.tags == e.domain_tag)[0][0])) for g in self.groups
                        if g.kind == "fluid" and (g.tags == e.domain_tag).any())
            u = self._fluid_stress(g, self._gather_fluid(g)).reshape(len(g.tags), g.M, 3, 5, 5)[k:k + 1]
            th = e.formThetaMat()[None]
            x = e.grad
            grad = GradOps(self.G_GLL, self.G_GLJ, x.dsdxii[None], x.dsdeta[None], x.dzdxii[None], x.dzdeta[None], x.inv_s[None],
                           g.axial, self.rd)
            s = tiso_spz_to_rtz(grad.grad6(u.astype(self.cd), g.nyq), th, self.rd)
            return self._eval_phi(s.reshape(1, g.M, 6, nPE)[0], g, phi, np.asarray(weights, float).reshape(nPE))
        g, k = self._solid_group_of(e)
        u = self._gather_solid(g)
        th = np.stack([self.elements[t].formThetaMat() for t in g.tags])
        s = tiso_spz_to_rtz(g.grad.grad6(u, g.nyq), th, self.rd)
        return self._eval_phi(s.reshape(u.shape[0], g.M, 6, nPE)[k], g, phi, np.asarray(weights, float).reshape(nPE))

    def curl(self, elem_tag, phi, weights):
        """SolidElement::computeCurl (SolidElement.cpp:281-345) after forceTIso, no PRT: grad9 -> SPZ_RTZ -> curl.
        FluidElement::computeCurl is identically zero (FluidElement.cpp:308-311)."""
        e = self.elements[elem_tag]
        if e.kind == "fluid":
            return np.zeros(3)
        if e.prt is not None:
            raise NotImplementedError("curl receivers: solid elements without PRT")
        g, k = self._solid_group_of(e)
        u = self._gather_solid(g)
        th = np.stack([self.elements[t].formThetaMat() for t in g.tags])
        s = tiso9_rotate(g.grad.grad9(u, g.nyq), th, self.rd, False).reshape(u.shape[0], g.M, 9, nPE)[k]
        v = self._eval_phi(s, g, phi, np.asarray(weights, float).reshape(nPE))
        return np.array([v[7] - v[5], v[2] - v[6], v[3] - v[1]])

    def _ground_motion_fluid(self, e, phi, w):
        """FluidElement::computeGroundMotion (FluidElement.cpp:163-215): gather, Gradient::computeGrad, [c2r, K, r2c | K],
        then the same azimuthal evaluation on the acoustic stress (= the fluid displacement)."""
        rd, cd = self.rd, self.cd
        g, k = next((g, int(np.nonzero(g.tags == e.domain_tag)[0][0])) for g in self.groups
                    if g.kind == "fluid" and (g.tags == e.domain_tag).any())
        M = g.M
        u = self._gather_fluid(g)                                   # [E,M,5,5]
        s = self._fluid_stress(g, u).reshape(u.shape[0], M, 3, nPE)[k]           # [M,3,25]
        top = g.Nu - g.nyq
        al = np.arange(1, top + 1)
        ex = 2.0 * np.exp(1j * al * phi)
        up = s[0].real + (ex[:, None, None] * s[1:top + 1]).real.sum(axis=0)     # [3,25]
        return (up * w[None, :]).sum(axis=1)
