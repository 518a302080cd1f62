"""ctypes front end of oracle/oracle.c (CPU ORACLE -- test infrastructure only).

`COracle(d)` takes a finalized float32 `OracleDomain` and runs its heavy verbs (computeStiff,
updateNewmark) in C/OpenMP on the very same numpy arrays, so the two oracles cross-check each other
and the C one serves as the timed CPU baseline of bench.py.  Supported subset: every element kind of
the numpy oracle; points with Mass1D (Mass3D points are delegated back to numpy)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "liboracle.so"])
    return LIB


class Group(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("E", "M", "Nr", "axial", "nyq", "fluid", "law", "is3d", "tiso", "ncoef",
                                       "att_kind", "nsls", "do_kappa", "P")] + \
               [("pidx", C.c_void_p)] + [(n, C.c_void_p) for n in ("dsdxii", "dsdeta", "dzdxii", "dzdeta", "inv_s")] + \
               [("theta", C.c_void_p), ("coef", C.c_void_p), ("alpha", C.c_void_p), ("beta", C.c_void_p),
                ("gamma", C.c_void_p), ("dk3", C.c_void_p), ("dmu", C.c_void_p), ("dmu2", C.c_void_p),
                ("memvar", C.c_void_p), ("stressR", C.c_void_p)]


_LAW = {"iso": 0, "ti": 1, "aniso": 2}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class COracle:
    def __init__(self, d):
        if d.rd != np.float32:
            raise ValueError("the C oracle is single precision")
        self.d = d
        self.lib = C.CDLL(build())
        self.lib.orc_num_threads.restype = C.c_int
        self.G_GLL = np.ascontiguousarray(d.G_GLL.reshape(25), dtype=np.float32)
        self.G_GLJ = np.ascontiguousarray(d.G_GLJ.reshape(25), dtype=np.float32)
        self.keep = []
        self.groups = []
        live = (d.p_nu - d.p_nyq + 1).astype(np.int32)
        self.s_nlive = np.ascontiguousarray(live[d.s_tag])
        self.f_nlive = np.ascontiguousarray(live[d.f_tag])
        for g in d.groups:
            if getattr(g, "hasPRT", False):
                raise NotImplementedError("oracle.c has no particle-relabelling path: PRT elements are checked with the numpy oracle only")
            G = Group()
            G.E, G.M, G.Nr, G.axial, G.nyq = len(g.tags), g.M, g.Nr, int(g.axial), g.nyq
            G.fluid = int(g.kind == "fluid")
            G.is3d = int(g.elem3D)
            c = lambda a, t=np.float32: np.ascontiguousarray(a, dtype=t)
            arrs = dict(pidx=c(g.pidx, np.int32))
            for k in ("dsdxii", "dsdeta", "dzdxii", "dzdeta", "inv_s"):
                arrs[k] = c(getattr(g.grad, k).reshape(G.E, 25))
            att = None
            if g.kind == "solid":
                G.law, G.tiso, G.ncoef = _LAW[g.law], int(g.inTIso), g.coef.shape[0]
                arrs["coef"] = c(g.coef)
                if g.inTIso:
                    arrs["theta"] = c(g.theta.reshape(G.E, 25), np.float64)
                att = g.att
            else:
                G.law, G.tiso, G.ncoef = 0, 0, 1
                arrs["coef"] = c(g.K)
            if att is not None:
                G.att_kind, G.nsls, G.do_kappa, G.P = (2 if att.cg4 else 1), att.nsls, int(att.doKappa), att.P
                for k in ("alpha", "beta", "gamma", "dk3", "dmu", "dmu2"):
                    arrs[k] = c(getattr(att, k))
            for k, a in arrs.items():
                setattr(G, k, _ptr(a))
            self.keep.append(arrs)
            self.groups.append((g, G, att))

    def threads(self):
        return self.lib.orc_num_threads()

    def set_threads(self, n):
        """explicit OpenMP thread count (launchers such as torchrun export OMP_NUM_THREADS=1)"""
        self.lib.orc_set_num_threads(C.c_int(int(n)))
        return self.threads()

    def computeStiff(self):
        d = self.d
        for g, G, att in self.groups:
            if att is not None:
                att.memvar = np.ascontiguousarray(att.memvar)
                att.stressR = np.ascontiguousarray(att.stressR)
                G.memvar, G.stressR = _ptr(att.memvar), _ptr(att.stressR)
            fld, nlive = (d.F, self.f_nlive) if g.kind == "fluid" else (d.S, self.s_nlive)
            self.lib.orc_compute_stiff(C.byref(G), _ptr(self.G_GLL), _ptr(self.G_GLJ), _ptr(fld["displ"]),
                                       _ptr(fld["stiff"]), _ptr(nlive), C.c_int(d.Mmax))

    def updateNewmark(self, dt):
        d = self.d
        if any(m.is3D or getattr(m, "ocean", False) for m in d.s_mass) or any(m.is3D for m in d.f_mass):
            return d.updateNewmark(dt)
        if not hasattr(self, "_pt"):
            mk = lambda tags, masses: dict(
                nu=np.ascontiguousarray(d.p_nu[tags], dtype=np.int32), nr=np.ascontiguousarray(d.p_nr[tags], dtype=np.int32),
                ax=np.ascontiguousarray(d.p_axial[tags], dtype=np.uint8),
                im=np.array([m.invMass for m in masses], dtype=np.float32))
            self._pt = (mk(d.s_tag, d.s_mass), mk(d.f_tag, d.f_mass), np.ascontiguousarray(d.f_surf, dtype=np.uint8))
        ps, pf, surf = self._pt
        for fld, p, ncomp, sf in ((d.S, ps, 3, None), (d.F, pf, 1, surf)):
            n = len(p["nu"])
            if n == 0:
                continue
            self.lib.orc_update_newmark(C.c_int(n), C.c_int(ncomp), C.c_int(d.Mmax), _ptr(p["nu"]), _ptr(p["nr"]), _ptr(p["ax"]),
                                        _ptr(sf) if sf is not None else None, _ptr(p["im"]), _ptr(fld["displ"]),
                                        _ptr(fld["veloc"]), _ptr(fld["accel"]), _ptr(fld["stiff"]), C.c_double(dt))

    def step(self, dt, stf):
        d = self.d
        self.updateNewmark(dt)
        d.applySource(stf)
        self.computeStiff()
        d.coupleSolidFluid()
