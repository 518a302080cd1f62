"""The reference's input files -> the source, source-time function and receivers of the hot path.

The callers either side of the time loop that SURVEY.md section 8(f) ranks next (items 1 and 2): what `axisem_main`
(S/axisem.cpp:12-176) builds from `input/` besides the mesh, restated so that the unchanged `inparam.*`, `CMTSOLUTION` and
`STATIONS` files drive the CUDA path.  Every piece is pinned against the reference's own classes run on the template inputs
(oracle/_ref/axisem3d_dump, tests/test_preloop_reference.py).

    Parameters                      S/preloop/utilities/Parameters.cpp:15-137, Parameters.h:23-39
    Geodesy                         S/preloop/utilities/Geodesy.cpp:13-200
    NrField::buildInparam           S/preloop/nrfield/NrField.cpp:15-64, EmpNrField.cpp:22-53, ConstNrField.cpp:16-19
    Source::buildInparam / locate   S/preloop/source/Source.cpp:62-234
    Earthquake / PointForce         S/preloop/source/Earthquake.cpp:20-91, PointForce.cpp:18-55
    STF::buildInparam, Erf/Gauss/Ricker   S/preloop/source/stf/STF.cpp:21-62, ErfSTF.cpp:10-21, GaussSTF.cpp:10-21, RickerSTF.cpp:10-22
    ReceiverCollection / Receiver   S/preloop/receiver/ReceiverCollection.cpp:26-133, 265-389, Receiver.cpp:19-118
    PointwiseRecorder::record       S/core/output/pointwise/PointwiseRecorder.cpp:62-96 (SPZ -> RTZ / ENZ)
    Mesh::computeRadiusRef          S/preloop/mesh/Mesh.cpp:210-259 (no geometric 3-D models: R_ref = R_outer - depth)
    Mapping::invMapping             S/preloop/spectral/mapping/Mapping.cpp:15-26
"""
from __future__ import annotations

import bisect
import math
import os

import numpy as np

from . import model as M
from . import spectral as SP

DEGREE = math.pi / 180.0
TINY_DOUBLE, TINY_SINGLE = 1e-10, 1e-5          # S/global.h:8-9


# ---------------------------------------------------------------------------------------------------------------- Parameters
class Parameters:
    """inparam.model / inparam.nu / inparam.time_src_recv / inparam.advanced: `KEY value value ...` lines, `#` comments
    (Parameters.cpp:102-112); a key may appear several times, its values accumulate."""

    FILES = ("inparam.model", "inparam.nu", "inparam.time_src_recv", "inparam.advanced")

    def __init__(self, input_dir):
        self.input_dir = input_dir
        self.kv = {}
        for name in self.FILES:
            path = os.path.join(input_dir, name)
            if not os.path.exists(path):
                raise RuntimeError("Parameters::readParFile || Error opening parameter file: ||" + path)
            with open(path) as f:
                for line in f:
                    strs = [s for s in line.strip("\t \r\n").replace("\t", " ").split(" ") if s != ""]
                    if not strs or strs[0].startswith("#"):
                        continue
                    self.kv.setdefault(strs[0], []).extend(strs[1:])

    def size(self, key):
        return len(self.kv[key])

    def get(self, key, typ=str, index=0):
        try:
            val = self.kv[key][index]
            if typ is bool:
                u = val.upper()
                if u in ("TRUE", "YES", "ON", "1"):
                    return True
                if u in ("FALSE", "NO", "OFF", "0"):
                    return False
                raise ValueError(val)
            return typ(val)
        except (KeyError, IndexError, ValueError):
            raise RuntimeError("Parameters::getValue || Invalid parameter, keyword = " + key + ".")


# ------------------------------------------------------------------------------------------------------------------ Geodesy
class Geodesy:
    """Geographic <-> geocentric <-> source-centred coordinates (Geodesy.cpp).  `flattening` = 0 switches ellipticity off
    (MODEL_3D_ELLIPTICITY_MODE off); otherwise 1 / MODEL_3D_ELLIPTICITY_INVF scaled by the radial profile of the mesh file's
    `ellipticity` variable (ExodusModel.cpp:557-572)."""

    def __init__(self, r_outer=6371e3, flattening=0.0, knots=(), coeffs=()):
        self.r_outer, self.f0 = float(r_outer), float(flattening)
        self.knots, self.coeffs = np.asarray(knots, dtype=np.float64), np.asarray(coeffs, dtype=np.float64)
        self._knots_list = [float(k) for k in self.knots]
        self._q_cache = {}

    @classmethod
    def from_mesh(cls, mesh, par=None):
        mode = par.get("MODEL_3D_ELLIPTICITY_MODE") if par is not None else "off"
        if mode.lower() == "off" or mesh.ellip_knots is None:
            return cls(mesh.r_outer)
        inv_f = par.get("MODEL_3D_ELLIPTICITY_INVF", float)
        if inv_f <= 0.0:
            raise RuntimeError("ExodusModel::buildInparam || Invalid flattening.")
        return cls(mesh.r_outer, 1.0 / inv_f, mesh.ellip_knots, mesh.ellip_coeffs)

    def flattening(self, r):
        """getFlattening: linear interpolation in the first knot interval (i >= 1) whose upper end is >= r / R"""
        x = r / self.r_outer
        n = len(self.knots)
        i = max(bisect.bisect_left(self._knots_list, x), 1) if n else 1
        if i >= n:
            return self.f0
        f = (self.coeffs[i] - self.coeffs[i - 1]) / (self.knots[i] - self.knots[i - 1]) * (x - self.knots[i - 1]) + self.coeffs[i - 1]
        return float(f) * self.f0

    def flattening_array(self, r):
        """getFlattening on an array: the first knot interval (i >= 1) whose upper end is >= r / R; the full flattening beyond"""
        x = np.asarray(r, dtype=np.float64) / self.r_outer
        if len(self.knots) < 2:
            return np.full(x.shape, self.f0)
        i = np.maximum(np.searchsorted(self.knots, x, side="left"), 1)
        inside = i < len(self.knots)
        i = np.minimum(i, len(self.knots) - 1)
        k0, k1, c0, c1 = self.knots[i - 1], self.knots[i], self.coeffs[i - 1], self.coeffs[i]
        f = (c1 - c0) / (k1 - k0) * (x - k0) + c0
        return np.where(inside, f, 1.0) * self.f0

    @staticmethod
    def atan4(y, x):
        if math.sqrt(x * x + y * y) < TINY_DOUBLE:
            return 0.0, False
        t = math.atan2(y, x)
        if t < 0.0:
            t += math.pi
        if y < 0.0:
            t += math.pi
        return t, True

    def lat2theta(self, lat, depth):
        omf2 = (1.0 - self.flattening(self.r_outer - depth)) ** 2
        lim = 90.0 - TINY_DOUBLE
        if abs(lat) > lim:
            if abs(lat) > 90.1:
                raise RuntimeError("Geodesy::lat2Theta || Latitude out of range [-90, 90]")
            lat = math.copysign(lim, lat)
        return math.pi / 2.0 - math.atan(omf2 * math.tan(lat * DEGREE))

    def theta2lat(self, theta, depth):
        inv = 1.0 / (1.0 - self.flattening(self.r_outer - depth)) ** 2
        lat = math.pi / 2.0 - theta
        lim = (90.0 - TINY_DOUBLE) * DEGREE
        if abs(lat) > lim:
            if abs(lat) > 90.1 * DEGREE:
                raise RuntimeError("Geodesy::theta2Lat || Theta out of range [0, pi]")
            lat = math.copysign(lim, lat)
        return math.atan(inv * math.tan(lat)) / DEGREE

    @staticmethod
    def lon2phi(lon):
        return (lon + 360.0 if lon < 0.0 else lon) * DEGREE

    @staticmethod
    def phi2lon(phi):
        return (phi - 2.0 * math.pi if phi > math.pi else phi) / DEGREE

    @staticmethod
    def to_cartesian(rtp):
        r, t, p = rtp
        return np.array([r * math.sin(t) * math.cos(p), r * math.sin(t) * math.sin(p), r * math.cos(t)])

    @classmethod
    def to_spherical(cls, xyz):
        r = float(np.sqrt(xyz[0] * xyz[0] + xyz[1] * xyz[1] + xyz[2] * xyz[2]))
        t = 0.0 if r < TINY_DOUBLE else math.acos(xyz[2] / r)
        p, defined = cls.atan4(xyz[1], xyz[0])
        return np.array([r, t, p]), defined

    @staticmethod
    def rotation_matrix(theta, phi):
        ct, st, cp, sp = math.cos(theta), math.sin(theta), math.cos(phi), math.sin(phi)
        return np.array([[ct * cp, -sp, st * cp], [ct * sp, cp, st * sp], [-st, 0.0, ct]])

    def _q(self, srclat, srclon, srcdep):
        key = (srclat, srclon, srcdep)
        if key not in self._q_cache:
            self._q_cache[key] = self.rotation_matrix(self.lat2theta(srclat, srcdep), self.lon2phi(srclon))
        return self._q_cache[key]

    def rotate_glob2src(self, rtp_g, srclat, srclon, srcdep):
        q = self._q(srclat, srclon, srcdep)
        x = self.to_cartesian(rtp_g)
        xs = np.array([q[0, k] * x[0] + q[1, k] * x[1] + q[2, k] * x[2] for k in range(3)])       # Q^T x, summed in index order
        rtp, defined = self.to_spherical(xs)
        if not defined:
            rtp[2] = rtp_g[2]
        return rtp

    def rotate_src2glob(self, rtp_s, srclat, srclon, srcdep):
        q = self._q(srclat, srclon, srcdep)
        x = self.to_cartesian(rtp_s)
        xg = np.array([q[k, 0] * x[0] + q[k, 1] * x[1] + q[k, 2] * x[2] for k in range(3)])
        rtp, defined = self.to_spherical(xg)
        if not defined:
            rtp[2] = rtp_s[2]
        return rtp

    def back_azimuth(self, srclat, srclon, srcdep, reclat, reclon, recdep):
        st, sp = self.lat2theta(srclat, srcdep), self.lon2phi(srclon)
        d, e, f, c = math.sin(sp), -math.cos(sp), -math.sin(st), math.cos(st)
        a, b = f * e, -f * d
        rt, rp = self.lat2theta(reclat, recdep), self.lon2phi(reclon)
        d1, e1, f1, c1 = math.sin(rp), -math.cos(rp), -math.sin(rt), math.cos(rt)
        g1, h1 = -c1 * e1, c1 * d1
        ss = (a - d1) ** 2 + (b - e1) ** 2 + c * c - 2.0
        sc = (a - g1) ** 2 + (b - h1) ** 2 + (c - f1) ** 2 - 2.0
        return self.atan4(ss, sc)[0]


# ------------------------------------------------------------------------------------------------------------------ NrField
def nu_field(par, r_outer=6371e3):
    """NrField::buildInparam: (nu, nu_fn, lucky) for ExodusMesh / SynthMesh -- constant or the empirical law."""
    lucky = par.get("FFTW_LUCKY_NUMBER", bool)
    typ = par.get("NU_TYPE").lower()
    if typ == "constant":
        nu = par.get("NU_CONST", int)
        if nu < 0:
            raise RuntimeError("ConstNrField::ConstNrField || Negative Nu.")
        return nu, None, lucky
    if typ == "wisdom":
        # WisdomNrField (WisdomNrField.cpp:10-27): Nu from the 4 nearest points of a learning run's wisdom file, times a factor
        wis = NuWisdom.read(os.path.join(par.input_dir, par.get("NU_WISDOM_REUSE_INPUT")))
        factor = par.get("NU_WISDOM_REUSE_FACTOR", float)
        factor = factor if factor > TINY_DOUBLE else 1.0
        return None, (lambda s, z: int(round(wis.get_nu(s, z, 4) * factor))), lucky
    if typ != "empirical":
        raise NotImplementedError("NrField::build || NU_TYPE " + typ + " (user-defined fields are compiled into the reference)")
    nu_ref, nu_min = par.get("NU_EMP_REF", int), par.get("NU_EMP_MIN", int)
    sc_s, sc_t, sc_d = (par.get(k, bool) for k in ("NU_EMP_SCALE_AXIS", "NU_EMP_SCALE_THETA", "NU_EMP_SCALE_DEPTH"))
    pow_s, fact_pi = par.get("NU_EMP_POW_AXIS", float), par.get("NU_EMP_FACTOR_PI", float)
    start_t, pow_t = par.get("NU_EMP_THETA_START", float) * DEGREE, par.get("NU_EMP_POW_THETA", float)
    fact_d0 = par.get("NU_EMP_FACTOR_SURF", float)
    start_d, end_d = par.get("NU_EMP_DEPTH_START", float) * 1e3, par.get("NU_EMP_DEPTH_END", float) * 1e3

    def nu_fn(s, z):
        r = math.hypot(s, z)
        theta = 0.0 if r < TINY_DOUBLE else math.acos(z / r)
        d = r_outer - r
        nu = float(nu_ref)
        if sc_s:
            nu *= (s / r_outer) ** pow_s
        if sc_t and theta > start_t:
            nu *= 1.0 + (fact_pi - 1.0) * ((theta - start_t) / (math.pi - start_t)) ** pow_t
        if sc_d and d <= end_d:
            nu *= fact_d0 if d <= start_d else 1.0 + (fact_d0 - 1.0) / (start_d - end_d) * (d - end_d)
        return max(nu_min, int(math.ceil(nu)))

    return None, nu_fn, lucky


# ------------------------------------------------------------------------------------------------------------------ wisdom
class NuWisdom:
    """The Nu(s, z) wisdom a learning run leaves behind (NuWisdom.cpp): rows (s, z, learnt Nu, original Nu)."""

    def __init__(self, data):
        data = np.asarray(data, dtype=np.float64).reshape(len(data), -1)
        if data.shape[1] == 3:
            data = np.concatenate([data, data[:, 2:3]], axis=1)
        if data.shape[1] != 4:
            raise RuntimeError("NuWisdom::readFromFile || Inconsistent dimensions for wisdom data, must be a matrix of 3 or 4 columns")
        self.sz = data[:, :2].copy()
        self.nu_learn = np.round(data[:, 2]).astype(np.int64)
        self.nu_orign = np.round(data[:, 3]).astype(np.int64)

    @classmethod
    def read(cls, path):
        """variable `axisem3d_wisdom` of a NetCDF file: NetCDF-4 / HDF5 (what the reference writes) through h5lite, the
        classic format (what write() below produces) through scipy"""
        with open(path, "rb") as f:
            magic = f.read(8)
        if magic[:3] == b"CDF":
            from scipy.io import netcdf_file
            with netcdf_file(path, "r", mmap=False) as nc:
                return cls(np.array(nc.variables["axisem3d_wisdom"][:], dtype=np.float64))
        from . import h5lite
        return cls(h5lite.File(path)["axisem3d_wisdom"].read())

    def write(self, path):
        """NuWisdom::writeToFile; classic NetCDF (readable by the reference's NetCDF_Reader through libnetcdf)"""
        from scipy.io import netcdf_file
        data = np.concatenate([self.sz, self.nu_learn[:, None].astype(np.float64), self.nu_orign[:, None].astype(np.float64)], axis=1)
        with netcdf_file(path, "w") as nc:
            nc.createDimension("ncdim_%d" % len(data), len(data))
            nc.createDimension("ncdim_4", 4)
            v = nc.createVariable("axisem3d_wisdom", "d", ("ncdim_%d" % len(data), "ncdim_4"))
            v[:] = data

    def get_nu(self, s, z, num_samples=4):
        """NuWisdom::getNu: inverse-distance mean of the learnt Nu of the nearest points (an exact hit wins)"""
        if len(self.sz) == 0:
            raise RuntimeError("NuWisdom::getNu || Wisdom is empty.")
        d = np.hypot(self.sz[:, 0] - s, self.sz[:, 1] - z)
        k = min(num_samples, len(d))
        idx = np.argpartition(d, k - 1)[:k]
        idx = idx[np.argsort(d[idx], kind="stable")]
        tot = num = 0.0
        for i in idx:
            if d[i] < TINY_DOUBLE:
                return int(self.nu_learn[i])
            tot += 1.0 / d[i]
            num += self.nu_learn[i] / d[i]
        return int(round(num / tot))

    def compression_ratio(self):
        return float(self.nu_learn.sum()) / float(self.nu_orign.sum())


def learn_parameters(par):
    """LearnParameters (NuWisdom.cpp:134-149) -> (invoked, cutoff, interval, output file name)"""
    cutoff = min(max(par.get("NU_WISDOM_LEARN_EPSILON", float), 1e-5), 0.1)
    interval = par.get("NU_WISDOM_LEARN_INTERVAL", int)
    return par.get("NU_WISDOM_LEARN", bool), cutoff, interval if interval > 0 else 5, par.get("NU_WISDOM_LEARN_OUTPUT")


# ------------------------------------------------------------------------------------------------ mapping helpers on a mesh
def inv_mapping(mesh, iq, s, z):
    """Mapping::invMapping: Newton iteration from the element centre, 10 iterations, |ds| < 1e-7 -> (xi, eta) or None."""
    xi = eta = 0.0
    for _ in range(10):
        sz = mesh._map(iq, xi, eta)
        ds, dz = s - float(sz[0]), z - float(sz[1])
        if math.hypot(ds, dz) < 1e-7:
            return xi, eta
        J = mesh._map(iq, xi, eta, jac=True)
        j00, j01, j10, j11 = float(J[0, 0]), float(J[0, 1]), float(J[1, 0]), float(J[1, 1])
        det = j00 * j11 - j01 * j10
        xi += (j11 * ds - j01 * dz) / det
        eta += (-j10 * ds + j00 * dz) / det
    return None


def near_me(mesh, iq, s, z):
    n = mesh.nodes[iq]
    return not (s > n[0].max() + TINY_SINGLE or s < n[0].min() - TINY_SINGLE or z > n[1].max() + TINY_SINGLE or z < n[1].min() - TINY_SINGLE)


def interp_lagrange(target, bases):
    """XMath::interpLagrange."""
    bases = np.asarray(bases, dtype=np.float64)
    res = np.empty(len(bases))
    for k, x0 in enumerate(bases):
        p1 = p2 = 1.0
        for i, x in enumerate(bases):
            if i != k:
                p1 *= target - x
                p2 *= x0 - x
        res[k] = p1 / p2
    return res


def radius_ref(mesh, geodesy, depth, lat, lon):
    """Mesh::computeRadiusRef: without geometric 3-D models the physical and the reference radius coincide; with them the
    reference radius is found by bisection on the (monotonic) physical radius."""
    if depth < TINY_DOUBLE:
        return mesh.r_outer
    models = getattr(mesh, "geometric", None)
    if not models:
        return mesh.r_outer - depth
    from . import relabelling as REL
    return REL.radius_ref(models, geodesy, mesh.r_outer, mesh.dist_tol, depth, lat, lon, mesh.vol_src)


# ------------------------------------------------------------------------------------------------------------------- Source
class Source:
    """An axial point source: Earthquake (moment tensor, CMTSOLUTION) or PointForce."""

    def __init__(self, kind, depth, lat, lon, **comp):
        self.kind, self.depth, self.lat, self.lon = kind, float(depth), float(lat), float(lon)
        if abs(self.lat - 90.0) < TINY_DOUBLE:
            self.lat, self.lon = 90.0, 0.0
        if abs(self.lat + 90.0) < TINY_DOUBLE:
            self.lat, self.lon = -90.0, 0.0
        self.c = comp

    @staticmethod
    def _parse(path, keys):
        """Source::parseLine: the FIRST word of a line contains the key (case-insensitive) -> the next word is the value."""
        val = {}
        if not os.path.exists(path):
            raise RuntimeError("Source::buildInparam || Error opening CMT data file: ||" + path)
        with open(path) as f:
            for line in f:
                w = line.split()
                if len(w) < 2:
                    continue
                for k in keys:
                    if k.lower() in w[0].lower():
                        try:
                            val[k] = float(w[1])
                        except ValueError:
                            pass
        for k in keys:
            if k not in val:
                raise RuntimeError("Source::checkValue || Error initializing source parameter: " + k)
        return val

    @classmethod
    def from_parameters(cls, par):
        if par.get("DEVELOP_NON_SOURCE_MODE", bool):
            return None
        typ = par.get("SOURCE_TYPE").lower()
        path = os.path.join(par.input_dir, par.get("SOURCE_FILE"))
        if typ == "earthquake":
            v = cls._parse(path, ("latitude", "longitude", "depth", "Mrr", "Mtt", "Mpp", "Mrt", "Mrp", "Mtp"))
            m = {k: v[k] * 1e-7 for k in ("Mrr", "Mtt", "Mpp", "Mrt", "Mrp", "Mtp")}                # dyn cm -> N m
            return cls("earthquake", v["depth"] * 1e3, v["latitude"], v["longitude"], **m)
        if typ == "point_force":
            v = cls._parse(path, ("latitude", "longitude", "depth", "Ft", "Fp", "Fr"))
            return cls("point_force", v["depth"] * 1e3, v["latitude"], v["longitude"], px=v["Ft"], py=v["Fp"], pz=v["Fr"])
        raise RuntimeError("Source::buildInparam || Unknown source type: " + typ)

    def locate(self, mesh, geodesy):
        """Source::locate -> (quad, interpFactZ[5]): the axial solid element holding (0, R_ref)."""
        s, z = 0.0, radius_ref(mesh, geodesy, self.depth, self.lat, self.lon)
        for iq in range(mesh.nelem):
            if not mesh.axial[iq] or mesh.is_fluid[iq] or not near_me(mesh, iq, s, z):
                continue
            xe = inv_mapping(mesh, iq, s, z)
            if xe is not None and abs(xe[1]) <= 1.000001:
                if abs(xe[0] + 1.0) > TINY_SINGLE:
                    raise RuntimeError("Source::locate || Bad source location.")
                return iq, interp_lagrange(xe[1], SP.P_GLL)
        raise RuntimeError("Source::release || Error locating source.")

    def fouriers(self, mesh, iq, interp_z):
        """Earthquake::computeSourceFourier / PointForce::computeSourceFourier (no particle relabelling): 25 blocks [nrow][3]."""
        nrow = 3 if self.kind == "earthquake" else 2
        out = [np.zeros((nrow, 3), dtype=np.complex128) for _ in range(25)]
        G_GLL, G_GLJ = np.asarray(SP.G_GLL).reshape(5, 5), np.asarray(SP.G_GLJ).reshape(5, 5)
        axJ = []
        for j in range(5):
            J = np.asarray(mesh._map(iq, SP.P_GLJ[0], SP.P_GLL[j], jac=True), dtype=np.float64).reshape(2, 2)
            axJ.append(J / (J[0, 0] * J[1, 1] - J[0, 1] * J[1, 0]))
        c = self.c
        two_pi, four_pi = 2.0 * math.pi, 4.0 * math.pi
        relab = getattr(mesh, "relab", None)
        relab = relab[iq] if relab else None
        if relab is not None:              # VX0..VX3 / J on the axis (ipol = 0), first azimuthal sample (Earthquake.cpp:35-43)
            X = relab.stiff_x()
            VX = [X[k][0, 0:5] for k in range(4)]
            Jp = relab.stiff_jacobian()[0, 0:5]
        for ip in range(5):
            for jp in range(5):
                f = out[ip * 5 + jp]
                if self.kind == "point_force":
                    if ip == 0:
                        fact = interp_z[jp] * (Jp[jp] if relab is not None else 1.0)
                        f[0, 2] += fact * c["pz"] / two_pi
                        f[1, 0] += fact * (c["px"] - 1j * c["py"]) / four_pi
                        f[1, 1] += fact * (c["py"] + 1j * c["px"]) / four_pi
                    continue
                mxx, myy, mzz, mxy, mxz, myz = c["Mtt"], c["Mpp"], c["Mrr"], c["Mtp"], c["Mrt"], c["Mrp"]
                w = np.zeros((5, 5))
                w[ip, jp] = 1.0
                GU = G_GLJ.T @ w
                UG = w @ G_GLL
                for js in range(5):
                    fact, J = interp_z[js], axJ[js]
                    dwds = J[1, 1] * GU[0, js] - J[1, 0] * UG[0, js]
                    dwdz = J[0, 0] * UG[0, js] - J[0, 1] * GU[0, js]
                    if relab is not None:
                        x0, x1, x2, x3 = VX[0][js], VX[1][js], VX[2][js], VX[3][js]
                        f[0, 0] += fact * dwds * (mxx + myy) * x0 / two_pi
                        f[0, 2] += fact * dwdz * (mxz * x1 + myz * x2 + mzz * x3) / two_pi
                        dip = fact * dwdz * ((mxx - 1j * mxy) * x1 + (mxy - 1j * myy) * x2 + (mxz - 1j * myz) * x3) / four_pi
                        f[1, 0] += dip
                        f[1, 1] += dip * 1j
                        f[1, 2] += fact * dwds * (mxz - 1j * myz) * x0 / two_pi
                        f[2, 0] += fact * dwds * ((mxx - myy) / 2.0 - 1j * mxy) * x0 / two_pi
                        f[2, 1] += fact * dwds * ((mxx - myy) / 2.0 - 1j * mxy) * x0 / two_pi * 1j
                        continue
                    f[0, 0] += fact * dwds * (mxx + myy) / two_pi
                    f[0, 2] += fact * dwdz * mzz / two_pi
                    f[1, 0] += fact * dwdz * (mxz - 1j * myz) / four_pi
                    f[1, 1] += fact * dwdz * (mxz - 1j * myz) / four_pi * 1j
                    f[1, 2] += fact * dwds * (mxz - 1j * myz) / two_pi
                    f[2, 0] += fact * dwds * ((mxx - myy) / 2.0 - 1j * mxy) / two_pi
                    f[2, 1] += fact * dwds * ((mxx - myy) / 2.0 - 1j * mxy) / two_pi * 1j
        return out

    def release(self, mesh, geodesy, elements, dec=None):
        """Source::release -> SourceTerm on the rank that holds the source element (None elsewhere)."""
        iq, interp_z = self.locate(mesh, geodesy)
        if dec is not None:
            loc = np.nonzero(np.asarray(dec.local_elems) == iq)[0]
            if len(loc) == 0:
                return None
            el = elements[int(loc[0])]
        else:
            el = elements[iq]
        return M.SourceTerm(el, self.fouriers(mesh, iq, interp_z))


# ---------------------------------------------------------------------------------------------------------------------- STF
def make_stf(kind, dt, duration, hdur, decay=1.628, max_steps=0):
    """ErfSTF / GaussSTF / RickerSTF + the DEVELOP_MAX_TIME_STEPS cut of STF::buildInparam -> (series float32, shift)."""
    n_before = int(math.ceil(2.5 * hdur / dt))
    n_after = int(math.ceil(duration / dt))
    shift = n_before * dt
    t = -shift + np.arange(n_before + n_after + 1, dtype=np.float64) * dt
    a = decay / hdur
    kind = kind.lower()
    if kind == "erf":
        s = np.array([math.erf(a * x) * 0.5 + 0.5 for x in t])
    elif kind == "gauss":
        s = np.exp(-(a * t) ** 2) * decay / (hdur * math.sqrt(math.pi))
    elif kind == "ricker":
        s = 2.0 * a ** 2 * np.exp(-(a * t) ** 2) * (2.0 * t ** 2 * a ** 2 - 1.0)
    else:
        raise RuntimeError("STF::buildInparam || Unknown stf type: " + kind)
    if max_steps > 0 and len(s) > max_steps:
        s = s[:max_steps]
    return s.astype(np.float32), shift


def stf_from_parameters(par, dt):
    hdur = max(par.get("SOURCE_STF_HALF_DURATION", float), 5.0 * dt)
    return make_stf(par.get("SOURCE_TIME_FUNCTION"), dt, par.get("TIME_RECORD_LENGTH", float), hdur, 1.628,
                    par.get("DEVELOP_MAX_TIME_STEPS", int))


def delta_t(par, mesh):
    """axisem.cpp:84-93: TIME_DELTA_T (0 = the mesh's own estimate) times TIME_DELTA_T_FACTOR."""
    dt = par.get("TIME_DELTA_T", float)
    if dt < TINY_DOUBLE:
        dt = mesh.estimate_dt()
    fact = par.get("TIME_DELTA_T_FACTOR", float)
    return dt * (fact if fact >= TINY_DOUBLE else 1.0)


# ---------------------------------------------------------------------------------------------------------------- receivers
class Receivers:
    """ReceiverCollection: the STATIONS file (name, network, lat | theta, lon | phi, elevation (ignored), depth) located in
    the mesh; element, azimuth, interpolation weights, and the rotation of the recorded SPZ motion to RTZ or ENZ."""

    def __init__(self, names, networks, c1, c2, depths, geographic, src, geodesy, components="RTZ", duplicated=1):
        self.geographic, self.components = bool(geographic), components.upper()
        if self.components not in ("RTZ", "ENZ", "SPZ"):
            raise RuntimeError("ReceiverCollection::buildInparam || Invalid parameter, keyword = OUT_STATIONS_COMPONENTS.")
        self.keys = ["%s.%s" % (nw, nm) for nm, nw in zip(names, networks)]
        # NB ReceiverCollection.cpp:102-124 declares its list of seen keys INSIDE the loop, so "rename" / "error" never see a
        # duplicate; stations with equal keys are all kept under the same name.  Kept as is.
        n = len(names)
        self.depth = np.asarray(depths, dtype=np.float64)
        self.theta, self.phi, self.lat, self.lon, self.baz = (np.zeros(n) for _ in range(5))
        g = geodesy
        for i in range(n):
            if geographic:
                rtp_g = np.array([1.0, g.lat2theta(c1[i], self.depth[i]), g.lon2phi(c2[i])])
                rtp_s = g.rotate_glob2src(rtp_g, src.lat, src.lon, src.depth)
            else:
                rtp_s = np.array([1.0, c1[i] * DEGREE, c2[i] * DEGREE])
                rtp_g = g.rotate_src2glob(rtp_s, src.lat, src.lon, src.depth)
            self.theta[i], self.phi[i] = rtp_s[1], rtp_s[2]
            self.lat[i], self.lon[i] = g.theta2lat(rtp_g[1], self.depth[i]), g.phi2lon(rtp_g[2])
            self.baz[i] = g.back_azimuth(src.lat, src.lon, src.depth, self.lat[i], self.lon[i], self.depth[i])
        self.quad = self.weights = None
        self.geodesy = geodesy

    @staticmethod
    def read_stations(path):
        names, nets, c1, c2, dep = [], [], [], [], []
        if not os.path.exists(path):
            raise RuntimeError("ReceiverCollection::ReceiverCollection || Error opening station data file " + path + ".")
        with open(path) as f:
            for line in f:
                # Parameters::splitString: trim blanks and tabs, split with compression; a '\r' stays in the last field
                strs = [s for s in line.rstrip("\n").strip("\t ").replace("\t", " ").split(" ") if s != ""]
                if len(strs) < 6:
                    continue
                try:
                    a, b, d = float(strs[2]), float(strs[3]), float(strs[5])
                except ValueError:
                    continue
                if any(x.lower() in ("dump_strain", "dump_curl") for x in strs[6:]):
                    raise NotImplementedError("ReceiverCollection: dump_strain / dump_curl stations are read back through "
                                              "ax3d_record_strain / ax3d_record_curl by the host, not by this runner (" + strs[0] + ")")
                names.append(strs[0]); nets.append(strs[1]); c1.append(a); c2.append(b); dep.append(d)
        return names, nets, c1, c2, dep

    @classmethod
    def from_parameters(cls, par, src, geodesy):
        fname = par.get("OUT_STATIONS_FILE")
        sys_ = par.get("OUT_STATIONS_SYSTEM").lower()
        if sys_ not in ("source-centered", "geographic"):
            raise RuntimeError("ReceiverCollection::buildInparam || Invalid parameter, keyword = OUT_STATIONS_SYSTEM.")
        cols = ([], [], [], [], []) if fname.lower() == "none" else cls.read_stations(os.path.join(par.input_dir, fname))
        r = cls(*cols, sys_ == "geographic", src, geodesy, par.get("OUT_STATIONS_COMPONENTS"))
        r.record_interval = max(par.get("OUT_STATIONS_RECORD_INTERVAL", int), 1)
        return r

    def locate(self, mesh, depth_in_ref=False):
        """Receiver::locate + computeInterpFact for every station -> self.quad [n], self.weights [n][25] (fluid elements:
        divided by the integral factor, Receiver.cpp:104-113)."""
        n = len(self.keys)
        self.quad = np.full(n, -1, dtype=np.int64)
        self.weights = np.zeros((n, 25))
        smax, smin = mesh.nodes[:, 0].max(), mesh.nodes[:, 0].min()
        zmax, zmin = mesh.nodes[:, 1].max(), mesh.nodes[:, 1].min()
        lo = mesh.nodes.min(axis=2)
        hi = mesh.nodes.max(axis=2)
        for i in range(n):
            r = mesh.r_outer - self.depth[i] if depth_in_ref else radius_ref(mesh, self.geodesy, self.depth[i], self.lat[i], self.lon[i])
            s, z = r * math.sin(self.theta[i]), r * math.cos(self.theta[i])
            if not (s > smax + TINY_SINGLE or s < smin - TINY_SINGLE or z > zmax + TINY_SINGLE or z < zmin - TINY_SINGLE):
                cand = np.nonzero((s <= hi[:, 0] + TINY_SINGLE) & (s >= lo[:, 0] - TINY_SINGLE) &
                                  (z <= hi[:, 1] + TINY_SINGLE) & (z >= lo[:, 1] - TINY_SINGLE))[0]
                for iq in cand:
                    xe = inv_mapping(mesh, int(iq), s, z)
                    if xe is not None and abs(xe[0]) <= 1.000001 and abs(xe[1]) <= 1.000001:
                        self.quad[i] = iq
                        wx = interp_lagrange(xe[0], SP.P_GLJ if mesh.axial[iq] else SP.P_GLL)
                        we = interp_lagrange(xe[1], SP.P_GLL)
                        w = np.outer(wx, we)
                        if mesh.is_fluid[iq]:
                            w = w / np.asarray(mesh.ifact[iq]).reshape(5, 5)
                        self.weights[i] = w.reshape(-1)
                        break
            if self.quad[i] < 0:
                raise RuntimeError("ReceiverCollection::release || Error locating receiver || Name = " + self.keys[i])
        return self

    def release(self, domain, elements, dec=None):
        """Registers the stations of this rank's elements with the domain's recorder; returns their indices."""
        if dec is None:
            loc = {iq: iq for iq in set(int(q) for q in self.quad)}
        else:
            loc = {int(e): k for k, e in enumerate(dec.local_elems)}
        mine = [i for i in range(len(self.keys)) if int(self.quad[i]) in loc]
        if mine:
            domain.setReceivers([elements[loc[int(self.quad[i])]].domain_tag for i in mine], self.phi[mine], self.weights[mine])
        self.mine = np.array(mine, dtype=np.int64)
        return self.mine

    def rotate(self, spz, idx=None):
        """PointwiseRecorder::record: [..., nrec, 3] SPZ samples -> the OUT_STATIONS_COMPONENTS frame, in fp32 like the reference."""
        gm = np.array(spz, dtype=np.float32, copy=True)
        if self.components == "SPZ":
            return gm
        idx = np.arange(len(self.keys)) if idx is None else np.asarray(idx)
        cost, sint = np.cos(self.theta[idx]).astype(np.float32), np.sin(self.theta[idx]).astype(np.float32)
        ur = gm[..., 0] * sint + gm[..., 2] * cost
        ut = gm[..., 0] * cost - gm[..., 2] * sint
        if self.components == "ENZ":
            cb, sb = np.cos(self.baz[idx]).astype(np.float32), np.sin(self.baz[idx]).astype(np.float32)
            up = gm[..., 1].copy()
            gm[..., 0] = -ut * sb + up * cb
            gm[..., 1] = -ut * cb - up * sb
            gm[..., 2] = ur
        else:
            gm[..., 0] = ut
            gm[..., 2] = ur
        return gm
