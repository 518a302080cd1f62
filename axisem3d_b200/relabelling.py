"""3-D geometric models (undulated interfaces) by particle relabelling: what makes the preloop hand PRT_1D / PRT_3D operators,
Jacobian-scaled moduli and masses and tilted solid-fluid normals to the hot path.

    Geometric3D::buildInparam, Ellipticity::getDeltaR   S/3d_model/3d_geometric/Geometric3D.cpp:20-81, ellipticity/Ellipticity.cpp:9-18
    Relabelling (whole class)                            S/preloop/physics/relabelling/Relabelling.cpp:15-291
    Quad::getHminSlices, computeGradientScalar           S/preloop/mesh/Quad.cpp:630-659, 488-508
    PreloopGradient::gradScalar                          S/preloop/utilities/PreloopGradient.cpp:17-50
    XMath::gaussianSmoothing / trigonResampling / linearResampling / findClosestDist   S/preloop/utilities/XMath.cpp:18-30, 77-186
    Mesh::computeRadiusRef / computeRPhysical            S/preloop/mesh/Mesh.cpp:210-267

Only models without data files are restated: `Ellipticity` (MODEL_3D_ELLIPTICITY_MODE full).  A model is any object with
`delta_r(r, theta, phi)` on arrays (geocentric coordinates of the globe).
"""
from __future__ import annotations

import math

import numpy as np

from . import spectral as SP
from .volumetric import geocentric_global, linear_resampling, equal_rows

TINY = 1e-10


class Ellipticity:
    """Ellipticity::getDeltaR: the radial shift that turns the sphere of radius r into the ellipsoid of equal volume with the
    flattening of that depth."""

    def __init__(self, geodesy):
        self.g = geodesy

    def delta_r(self, r, theta, phi):
        r, theta = np.asarray(r, dtype=np.float64), np.asarray(theta, dtype=np.float64)
        f = self.g.flattening_array(r)
        rs = np.where(r < TINY, 1.0, r)
        b = (1.0 - f) ** (2.0 / 3.0) * rs
        a = b / (1.0 - f)
        tmp = (a * np.cos(theta)) ** 2 + (b * np.sin(theta)) ** 2
        return np.where(r < TINY, 0.0, a * b / np.sqrt(tmp) - rs)


def from_parameters(par, geodesy):
    """Geometric3D::buildInparam for the models that need no data files."""
    if par.get("MODEL_3D_GEOMETRIC_NUM", int) != 0:
        raise NotImplementedError("Geometric3D::buildInparam || MODEL_3D_GEOMETRIC_LIST models (crust1, EMC) need the reference's data files")
    return [Ellipticity(geodesy)] if par.get("MODEL_3D_ELLIPTICITY_MODE").lower() == "full" else []


def r_physical(models, r, theta, phi):
    return r + sum(float(m.delta_r(np.array([r]), np.array([theta]), np.array([phi]))[0]) for m in models)


def radius_ref(models, geodesy, r_outer, dist_tol, depth, lat, lon, src):
    """Mesh::computeRadiusRef: the radius in the undeformed (reference) mesh of a point `depth` below the physical surface.

    NB Mesh.cpp:217 tests `mPhi2D < -DBL_MAX * .9`, which is true exactly when the 2-D mode is OFF (Mesh.cpp:49 sets mPhi2D to
    -DBL_MAX then; Quad.cpp:477 reads the same test the other way round).  So in a normal 3-D run the point is rotated to the
    source-centred frame, its azimuth is replaced by -DBL_MAX, and it is rotated back before the undulation is evaluated.
    The reference's seismograms are computed with receivers placed this way, so the restatement does the same."""
    theta, phi = geodesy.lat2theta(lat, depth), geodesy.lon2phi(lon)
    rtp_s = geodesy.rotate_glob2src(np.array([1.0, theta, phi]), src.lat, src.lon, src.depth)
    rtp_s[2] = -1.7976931348623157e308
    rtp_g = geodesy.rotate_src2glob(rtp_s, src.lat, src.lon, src.depth)
    theta, phi = float(rtp_g[1]), float(rtp_g[2])
    if depth < TINY:
        return r_outer
    target = r_physical(models, r_outer, theta, phi) - depth
    tol = min(1e-5, dist_tol * 1e-5)
    cur, upper, lower = r_outer - depth, r_outer, 0.0
    for _ in range(10001):
        diff = r_physical(models, cur, theta, phi) - target
        if abs(diff) < tol:
            return cur
        if diff > 0.0:
            upper = cur
        else:
            lower = cur
        cur = 0.5 * (lower + upper)
    raise RuntimeError("Mesh::computeRadiusRef || Failed to find reference radius.")


# ------------------------------------------------------------------------------------------------------------------ XMath
def _r2c(x):
    """PreloopFFTW::computeR2C followed by the 1 / Nr the class applies (PreloopFFTW.cpp): Fourier coefficients of a real ring"""
    n = len(x)
    return np.fft.rfft(np.asarray(x, dtype=np.float64)) / n


def _c2r(c, n):
    """PreloopFFTW::computeC2R: unnormalised backward transform of Nu + 1 coefficients to n samples"""
    c = np.array(c, dtype=np.complex128)
    return np.fft.irfft(c, n) * n


def trigon_resampling(new_size, original):
    original = np.asarray(original, dtype=np.float64)
    n = len(original)
    if new_size == n:
        return original.copy()
    if all(abs(original[i] - original[0]) <= TINY * abs(original[0]) for i in range(1, n)):
        return np.full(new_size, original[0])
    four = _r2c(original)
    phi = 2.0 * math.pi / new_size * np.arange(new_size)
    val = np.full(new_size, four[0].real)
    for a in range(1, len(four)):
        fac = 1.0 if (n % 2 == 0 and a == len(four) - 1) else 2.0
        val += fac * (four[a] * np.exp(1j * a * phi)).real
    return val


def gaussian_smoothing(data, order, dev, period):
    data = np.asarray(data, dtype=np.float64)
    n = len(data)
    if n == 0:
        return data
    order = min(order, (n + 1) // 2 - 1)
    if order == 0:
        return data.copy()
    dev *= order
    k = np.arange(-order, order + 1)
    g = np.exp(-0.5 * k * k / (dev * dev))
    g /= g.sum()
    out = np.zeros(n)
    for i in range(n):
        for j in range(-order, order + 1):
            kk = i + j
            kk = kk % n if period else min(max(kk, 0), n - 1)
            out[i] += g[j + order] * data[kk]
    return out


def closest_dist(s, z):
    p = np.stack([np.asarray(s).ravel(), np.asarray(z).ravel()], 1)
    d = np.linalg.norm(p[:, None, :] - p[None, :, :], axis=2) + np.eye(len(p)) * 1e300
    return float(d.min())


# ------------------------------------------------------------------------------------------------------------ Relabelling
class Relabelling:
    """The particle relabelling of one quad: dZ = the radial shift of its 25 points on the quad's Nr azimuthal samples, its
    gradient (dZ/dR, dZ/dT, dZ/dZ in the local R-T-Z frame) and their resampling to every point's own Nr."""

    def __init__(self, mesh, iq, models, geodesy, src):
        self.mesh, self.iq = mesh, iq
        g = mesh.geo[iq]
        self.s, self.z = g["s"].reshape(25), g["z"].reshape(25)
        self.nr = int(mesh.e_nr[iq])
        self.pnr = mesh.e_pnr[iq].reshape(25).astype(int)
        self.Z = np.hypot(self.s, self.z)
        self.theta = np.array([0.0 if r < TINY else math.acos(zz / r) for r, zz in zip(self.Z, self.z)])     # Geodesy::theta
        dz = np.zeros((self.nr, 25))
        for ipnt in range(25):
            rg, tg, pg = geocentric_global(geodesy, src, self.Z[ipnt], self.theta[ipnt], self.nr)
            for m in models:
                dz[:, ipnt] += m.delta_r(rg, tg, pg)
        self.dZ = dz
        self.zero = bool(np.abs(dz).max() < TINY)
        if not self.zero:
            self._check_hmin()
            self._form_gradient()
            self._form_mass()

    # Quad::getHminSlices with this relabelling
    def hmin_slices(self):
        h = np.empty(self.nr)
        for k in range(self.nr):
            h[k] = closest_dist(self.s + self.dZ[k] * np.sin(self.theta), self.z + self.dZ[k] * np.cos(self.theta))
        return h

    def _check_hmin(self):
        max_order = (self.nr + 1) // 2 - 1
        original = self.dZ.copy()
        for order in range(max_order + 1):
            for ipnt in range(25):
                self.dZ[:, ipnt] = gaussian_smoothing(original[:, ipnt], order, 100, True)
            h = self.hmin_slices()
            if trigon_resampling(5 * self.nr, h).min() >= h.min() * 0.8:
                return
        raise RuntimeError("Relabelling::checkHmin || Program should not have reached here.")

    def _form_gradient(self):
        nr, nu = self.nr, self.nr // 2
        nyq = 1 if nr % 2 == 0 else 0
        if nyq:
            for ipnt in range(25):
                self.dZ[:, ipnt] = trigon_resampling(nr, trigon_resampling(nr - 1, self.dZ[:, ipnt]))
        mesh, iq = self.mesh, self.iq
        g = mesh.geo[iq]
        axial = bool(mesh.axial[iq])
        det = g["det"]
        dsdxii, dsdeta, dzdxii, dzdeta = g["J00"] / det, -g["J01"] / det, -g["J10"] / det, g["J11"] / det
        inv_s = np.where(g["s"] > 0, 1.0 / np.where(g["s"] > 0, g["s"], 1.0), 0.0)
        if axial:
            inv_s[0, :] = 0.0
        G_GLL, G_GLJ = np.asarray(SP.G_GLL).reshape(5, 5), np.asarray(SP.G_GLJ).reshape(5, 5)
        GT = (G_GLJ if axial else G_GLL).T
        u = np.stack([_r2c(self.dZ[:, ipnt]) for ipnt in range(25)], 1).reshape(nu + 1, 5, 5)        # [alpha][ipol][jpol]
        ui = np.zeros((nu + 1, 3, 5, 5), dtype=np.complex128)
        for a in range(nu - nyq + 1):                                  # PreloopGradient::gradScalar
            GU, UG = GT @ u[a], u[a] @ G_GLL
            ui[a, 0] = dzdeta * GU + dzdxii * UG
            ui[a, 1] = inv_s * (1j * a * u[a])
            ui[a, 2] = dsdeta * GU + dsdxii * UG
        if axial:
            ui[0, 0, 0, :] = 0.0
            ui[0, 1, 0, :] = 0.0
            if nu >= 1:
                ui[1, 1, 0, :] = 1j * ui[1, 0, 0, :]
                ui[1, 2, 0, :] = 0.0
            ui[2:, :, 0, :] = 0.0
        if nyq:
            ui[nu] = 0.0
        ui = ui.reshape(nu + 1, 3, 25)
        self.dZdT, self.dZdZ, self.dZdR = (np.zeros((nr, 25)) for _ in range(3))
        for ipnt in range(25):
            self.dZdT[:, ipnt] = _c2r(ui[:, 1, ipnt], nr)
            drds, drdz = _c2r(ui[:, 0, ipnt], nr), _c2r(ui[:, 2, ipnt], nr)
            ct, st = math.cos(self.theta[ipnt]), math.sin(self.theta[ipnt])
            self.dZdZ[:, ipnt] = drdz * ct + drds * st
            self.dZdR[:, ipnt] = -drdz * st + drds * ct

    def _form_mass(self):
        self.m_dZ, self.m_dZdR, self.m_dZdT, self.m_dZdZ = [], [], [], []
        for ipnt in range(25):
            n = int(self.pnr[ipnt])
            self.m_dZ.append(linear_resampling(n, self.dZ[:, ipnt]))
            self.m_dZdR.append(linear_resampling(n, self.dZdR[:, ipnt]))
            self.m_dZdT.append(linear_resampling(n, self.dZdT[:, ipnt]))
            self.m_dZdZ.append(linear_resampling(n, self.dZdZ[:, ipnt]))

    def is_par1d(self):
        return equal_rows(self.dZ)

    def stiff_jacobian(self):
        J = np.empty((self.nr, 25))
        for ipnt in range(25):
            j22 = 1.0 + self.dZdZ[:, ipnt]
            if self.Z[ipnt] < TINY:
                J[:, ipnt] = j22 * j22 * j22
            else:
                j00 = 1.0 + self.dZ[:, ipnt] / self.Z[ipnt]
                J[:, ipnt] = j00 * j00 * j22
        if J.min() <= 0.0:
            raise RuntimeError("Relabelling::getStiffJacobian || Negative Jacobian.")
        return J

    def stiff_x(self):
        """getStiffX as [4][Nr][25]"""
        X = np.empty((4, self.nr, 25))
        for ipnt in range(25):
            j0 = 1.0 + self.dZdZ[:, ipnt] if self.Z[ipnt] < TINY else 1.0 + self.dZ[:, ipnt] / self.Z[ipnt]
            j1, j2, j3 = self.dZdR[:, ipnt], self.dZdT[:, ipnt], 1.0 + self.dZdZ[:, ipnt]
            X[0, :, ipnt] = 1.0 / j0
            X[1, :, ipnt] = -j1 * (1.0 / (j0 * j3))
            X[2, :, ipnt] = -j2 * (1.0 / (j0 * j3))
            X[3, :, ipnt] = 1.0 / j3
        return X

    def mass_jacobian(self, ipnt):
        j22 = 1.0 + self.m_dZdZ[ipnt]
        if self.Z[ipnt] < TINY:
            J = j22 * j22 * j22
        else:
            j00 = 1.0 + self.m_dZ[ipnt] / self.Z[ipnt]
            J = j00 * j00 * j22
        if J.min() <= 0.0:
            raise RuntimeError("Relabelling::getMassJacobian || Negative Jacobian.")
        return J

    def sf_normal_rtz(self, ipnt):
        j0 = 1.0 + self.m_dZdZ[ipnt] if self.Z[ipnt] < TINY else 1.0 + self.m_dZ[ipnt] / self.Z[ipnt]
        return np.stack([-self.m_dZdR[ipnt] * j0, -self.m_dZdT[ipnt] * j0, j0 * j0], 1)            # [nr_p][3]
