"""3-D volumetric models of inparam.model (MODEL_3D_VOLUMETRIC_LIST) that need no data files, and their application to the
material of a quad -- what turns a 1-D element into one of the 3-D element kinds of the hot path.

    Volumetric3D::buildInparam      S/3d_model/3d_volumetric/Volumetric3D.cpp:23-88
    Volumetric3D_bubble / _cylinder S/3d_model/3d_volumetric/simple_shapes/Volumetric3D_bubble.cpp:11-134, Volumetric3D_cylinder.cpp:11-173
    Material::addVolumetric3D       S/preloop/physics/material/Material.cpp:96-231 (prepare3D: 902-932)
    Quad::computeGeocentricGlobal   S/preloop/mesh/Quad.cpp:468-481
    XMath::linearResampling         S/preloop/utilities/XMath.cpp:161-186

s20rts / s40rts / crust1 / EMC read data files that the checkout does not ship; full anisotropy (C_ij properties) is the
reference's.  A model is any object with `prop`, `ref`, `fluid` and `get(r, theta, phi) -> (in_range, value)` on arrays.
"""
from __future__ import annotations

import math

import numpy as np

from .preloop import DEGREE, Geodesy, Parameters

PROPS = ("VPV", "VPH", "VSV", "VSH", "RHO", "ANIS_ETA", "QKAPPA", "QMU", "VP", "VS")
ABS_SI = (1e3, 1e3, 1e3, 1e3, 1e3, 1.0, 1.0, 1.0, 1e3, 1e3)
REF_TYPES = {"absolute": 0, "abs": 0, "reference1d": 1, "ref1d": 1, "reference3d": 2, "ref3d": 2, "referenceperturb": 3, "refptb": 3}
KEYS = ("vpv", "vph", "vsv", "vsh", "rho", "eta", "qkp", "qmu")          # the eight property slots of Material


def _bool(s):
    u = s.upper()
    if u in ("TRUE", "YES", "ON", "1"):
        return True
    if u in ("FALSE", "NO", "OFF", "0"):
        return False
    raise RuntimeError("Parameters::castValue || Invalid argument encountered in Volumetric3D_bubble::initialize, arg = " + s + ".")


class Bubble:
    """bubble$<property>$<ref type>$<value inside>$<radius km>$<depth km>$<lat>$<lon>[$<source-centred>$<fluid>$<HWHM km>]"""

    def __init__(self, params, src, geodesy: Geodesy):
        if len(params) < 7:
            raise RuntimeError("Volumetric3D_bubble::initialize || Not enough parameters for a bubble-shaped heterogeneity. Need 7 at least.")
        names = [p.upper() for p in PROPS]
        if params[0].upper() not in names:
            raise RuntimeError("Volumetric3D_bubble::initialize || Unknown material property, name = " + params[0])
        self.prop = names.index(params[0].upper())
        if params[1].lower() not in REF_TYPES:
            raise RuntimeError("Volumetric3D_bubble::initialize || Unknown material reference type, type = " + params[1])
        self.ref = REF_TYPES[params[1].lower()]
        self.value = float(params[2])
        self.radius = float(params[3]) * 1e3
        depth, lat, lon = float(params[4]) * 1e3, float(params[5]), float(params[6])
        src_centred = _bool(params[7]) if len(params) > 7 else False
        self.fluid = _bool(params[8]) if len(params) > 8 else False
        self.hwhm = float(params[9]) * 1e3 if len(params) > 9 else -1.0
        if src_centred:
            rtp = geodesy.rotate_src2glob(np.array([geodesy.r_outer - depth, lat * DEGREE, lon * DEGREE]), src.lat, src.lon, src.depth)
        else:
            rtp = np.array([geodesy.r_outer - depth, geodesy.lat2theta(lat, depth), geodesy.lon2phi(lon)])
        self.xyz = geodesy.to_cartesian(rtp)
        if self.hwhm < 0.0:
            self.hwhm = self.radius * 0.2
        if self.ref == 0:
            self.hwhm = 0.0
            self.value *= ABS_SI[self.prop]

    def get(self, r, theta, phi):
        x = r * np.sin(theta) * np.cos(phi) - self.xyz[0]
        y = r * np.sin(theta) * np.sin(phi) - self.xyz[1]
        z = r * np.cos(theta) - self.xyz[2]
        dist = np.maximum(np.sqrt(x * x + y * y + z * z) - self.radius, 0.0)
        inside = ~(dist > 4.0 * self.hwhm)
        if self.hwhm > 0.0:
            std = self.hwhm / math.sqrt(2.0 * math.log(2.0))
            val = self.value * np.exp(-dist * dist / (std * std * 2.0))
        else:
            val = np.full(np.shape(dist), self.value)              # exp(-0 / 0) is never evaluated for a point in range
        return inside, val


class Cylinder:
    """cylinder$<property>$<ref type>$<value>$<radius km>$<depth1 km>$<lat1>$<lon1>$<depth2 km>$<lat2>$<lon2>
    [$<source-centred>$<fluid>$<lateral HWHM km>$<top-bottom HWHM km>]  (Volumetric3D_cylinder.cpp:11-173)"""

    def __init__(self, params, src, geodesy: Geodesy):
        if len(params) < 10:
            raise RuntimeError("Volumetric3D_cylinder::initialize || Not enough parameters for a cylinder-shaped heterogeneity. Need 10 at least.")
        names = [p.upper() for p in PROPS]
        if params[0].upper() not in names:
            raise RuntimeError("Volumetric3D_cylinder::initialize || Unknown material property, name = " + params[0])
        self.prop = names.index(params[0].upper())
        if params[1].lower() not in REF_TYPES:
            raise RuntimeError("Volumetric3D_cylinder::initialize || Unknown material reference type, type = " + params[1])
        self.ref = REF_TYPES[params[1].lower()]
        self.value = float(params[2])
        self.radius = float(params[3]) * 1e3
        d1, lat1, lon1 = float(params[4]) * 1e3, float(params[5]), float(params[6])
        d2, lat2, lon2 = float(params[7]) * 1e3, float(params[8]), float(params[9])
        src_centred = _bool(params[10]) if len(params) > 10 else False
        self.fluid = _bool(params[11]) if len(params) > 11 else False
        self.hwhm_lat = float(params[12]) * 1e3 if len(params) > 12 else -1.0
        self.hwhm_tb = float(params[13]) * 1e3 if len(params) > 13 else -1.0

        def end(d, lat, lon):
            if src_centred:
                rtp = geodesy.rotate_src2glob(np.array([geodesy.r_outer - d, lat * DEGREE, lon * DEGREE]), src.lat, src.lon, src.depth)
            else:
                rtp = np.array([geodesy.r_outer - d, geodesy.lat2theta(lat, d), geodesy.lon2phi(lon)])
            return geodesy.to_cartesian(rtp)
        self.p1, self.p2 = end(d1, lat1, lon1), end(d2, lat2, lon2)
        self.length = float(np.linalg.norm(self.p1 - self.p2))
        if self.hwhm_lat < 0.0:
            self.hwhm_lat = self.radius * 0.2
        if self.hwhm_tb < 0.0:
            self.hwhm_tb = self.length * 0.1
        if self.ref == 0:
            self.hwhm_lat = self.hwhm_tb = 0.0
            self.value *= ABS_SI[self.prop]

    def get(self, r, theta, phi):
        x = np.stack([r * np.sin(theta) * np.cos(phi), r * np.sin(theta) * np.sin(phi), r * np.cos(theta)], -1)
        a, b = x - self.p1, x - self.p2
        dline = np.linalg.norm(np.cross(a, b), axis=-1) / self.length
        inside = ~(dline > self.radius + 4.0 * self.hwhm_lat)
        dsurf = np.maximum(dline - self.radius, 0.0)
        if self.hwhm_lat > 0.0:
            std = self.hwhm_lat / math.sqrt(2.0 * math.log(2.0))
            val = self.value * np.exp(-dsurf * dsurf / (std * std * 2.0))
        else:
            val = np.full(np.shape(dline), self.value)
        dmax = np.maximum(np.linalg.norm(a, axis=-1), np.linalg.norm(b, axis=-1))
        dtb = np.sqrt(np.maximum(dmax * dmax - dline * dline, 0.0)) - self.length
        inside &= ~(dtb > 4.0 * self.hwhm_tb)
        if self.hwhm_tb > 0.0:
            std = self.hwhm_tb / math.sqrt(2.0 * math.log(2.0))
            val = np.where(dtb > 0.0, val * np.exp(-dtb * dtb / (std * std * 2.0)), val)
        return inside, val


def from_parameters(par: Parameters, src, geodesy):
    """Volumetric3D::buildInparam for the models that need no data files."""
    n = par.get("MODEL_3D_VOLUMETRIC_NUM", int)
    if n > par.size("MODEL_3D_VOLUMETRIC_LIST"):
        raise RuntimeError("Volumetric3D::buildInparam || Not enough model names provided in MODEL_3D_VOLUMETRIC_LIST")
    models = []
    for i in range(n):
        strs = [s for s in par.get("MODEL_3D_VOLUMETRIC_LIST", str, i).split("$") if s != ""]
        if strs[0].lower() == "bubble":
            models.append(Bubble(strs[1:], src, geodesy))
        elif strs[0].lower() == "cylinder":
            models.append(Cylinder(strs[1:], src, geodesy))
        else:
            raise NotImplementedError("Volumetric3D::buildInparam || model " + strs[0] + " (needs the reference's data files / is not restated)")
    return models


def geocentric_global(geodesy, src, r, theta, npnt):
    """Quad::computeGeocentricGlobal: the npnt azimuthal samples of the ring through (r, theta) in the source-centred frame,
    as geocentric (r, theta, phi) of the globe."""
    q = geodesy._q(src.lat, src.lon, src.depth)
    phi = 2.0 * math.pi / npnt * np.arange(npnt)
    xs = np.stack([r * math.sin(theta) * np.cos(phi), r * math.sin(theta) * np.sin(phi), np.full(npnt, r * math.cos(theta))])
    xg = np.stack([q[k, 0] * xs[0] + q[k, 1] * xs[1] + q[k, 2] * xs[2] for k in range(3)])
    rg = np.sqrt(xg[0] * xg[0] + xg[1] * xg[1] + xg[2] * xg[2])
    tg = np.where(rg < 1e-10, 0.0, np.arccos(np.clip(xg[2] / np.where(rg < 1e-10, 1.0, rg), -1.0, 1.0)))
    pg = np.arctan2(xg[1], xg[0])
    pg = np.where(pg < 0.0, pg + math.pi, pg)
    pg = np.where(xg[1] < 0.0, pg + math.pi, pg)                   # Geodesy::atan4
    undefined = np.sqrt(xg[0] * xg[0] + xg[1] * xg[1]) < 1e-10
    pg = np.where(undefined, phi, pg)
    return rg, tg, pg


def apply(models, geodesy, src, ref1d, s, z, nr, is_fluid):
    """Material::addVolumetric3D for one quad.  ref1d: {key: [25] 1-D values on the GLL points}; s, z: [25] point coordinates;
    nr: the quad's Nr.  Returns None when no sample of the quad is in range of a model (the material stays 1-D), else
    {key: [nr][25]}."""
    out = None
    slot = {0: (0,), 1: (1,), 2: (2,), 3: (3,), 4: (4,), 5: (5,), 6: (6,), 7: (7,), 8: (0, 1), 9: (2, 3)}       # VP -> VPV, VPH; VS -> VSV, VSH
    for ipnt in range(25):
        r = math.hypot(s[ipnt], z[ipnt])
        theta = 0.0 if r < 1e-10 else math.acos(z[ipnt] / r)
        rg, tg, pg = geocentric_global(geodesy, src, r, theta, nr)
        for m in models:
            if is_fluid and not m.fluid:
                continue
            inside, val = m.get(rg, tg, pg)
            if not inside.any():
                continue
            if out is None:
                out = {k: np.repeat(np.asarray(ref1d[k], dtype=np.float64).reshape(1, 25), nr, axis=0) for k in KEYS}
            for k in slot[m.prop]:
                col = out[KEYS[k]][:, ipnt]
                ref = float(ref1d[KEYS[k]][ipnt])
                if m.ref == 0:
                    new = val
                elif m.ref == 1:
                    new = ref * (1.0 + val)
                elif m.ref == 2:
                    new = col * (1.0 + val)
                else:
                    new = (col - ref) * (1.0 + val) + ref
                out[KEYS[k]][:, ipnt] = np.where(inside, new, col)
    return out


def linear_resampling(new_size, original):
    """XMath::linearResampling (equalRows with the reference's tolerance 1e-10)."""
    original = np.asarray(original, dtype=np.float64)
    n = len(original)
    if new_size == n:
        return original.copy()
    if all(abs(original[i] - original[0]) <= 1e-10 * abs(original[0]) for i in range(1, n)):
        return np.full(new_size, original[0])
    dphi, dphi0 = 2.0 * math.pi / new_size, 2.0 * math.pi / n
    out = np.empty(new_size)
    for i in range(new_size):
        phi = i * dphi
        l0 = int(phi / dphi0)
        l1 = 0 if l0 + 1 == n else l0 + 1
        out[i] = (original[l1] - original[l0]) / dphi0 * (phi - l0 * dphi0) + original[l0]
    return out


def equal_rows(a, tol=1e-10):
    """XMath::equalRows."""
    a = np.asarray(a, dtype=np.float64)
    n0 = np.linalg.norm(a[0])
    return all(np.linalg.norm(a[i] - a[0]) <= tol * n0 for i in range(1, a.shape[0]))
