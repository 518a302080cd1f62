"""Serialises a domain (the sequence of constructor calls of Mesh::release) for tests/cpp/host_driver.cpp.

`DumpDomain` is duck-typed like axisem3d_b200.domain.Domain / the oracle domain: mesh.release(d, dt) and
d.addSourceTerm(...) fill it; write(path, dt, stf) produces the binary the C++ driver replays through the facade
classes of axisem3d_b200/host/ax3d_host.hpp.  Layouts are those of include/axisem3d_b200.h."""
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER_SRC = os.path.join(ROOT, "tests", "cpp", "host_driver.cpp")
DRIVER = os.path.join(ROOT, "tests", "cpp", "host_driver")
_LAW = {"iso": 0, "ti": 1, "aniso": 2}


def build_driver(force=False):
    """g++ the facade test driver against the in-tree library (no nvcc needed: the facade is plain C++)."""
    lib_dir = os.path.join(ROOT, "axisem3d_b200")
    deps = [DRIVER_SRC, os.path.join(ROOT, "tests", "cpp", "release_domain.inc"), os.path.join(lib_dir, "host", "ax3d_host.hpp"),
            os.path.join(ROOT, "include", "axisem3d_b200.h")]
    if not force and os.path.exists(DRIVER) and all(os.path.getmtime(DRIVER) >= os.path.getmtime(d) for d in deps):
        return DRIVER
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-o", DRIVER, DRIVER_SRC, "-L" + lib_dir, "-laxisem3d_b200",
                           "-Wl,-rpath," + lib_dir])
    return DRIVER


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _colmajor(a):
    return _f32(np.asarray(a).T).reshape(-1)


def _mass(m):
    return _f32(m.invMass).reshape(-1) if m.is3D else np.array([m.invMass], dtype=np.float32)


class DumpDomain:
    def __init__(self):
        self.points, self.elements, self.sources = [], [], []
        self.G = None

    def setGMat(self, G_GLL, G_GLJ):
        self.G = (np.asarray(G_GLL, dtype=np.float64).reshape(25), np.asarray(G_GLJ, dtype=np.float64).reshape(25))

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

    def finalize(self):
        pass

    def write(self, path, dt, stf):
        out = [b"AX3D", self.G[0].tobytes(), self.G[1].tobytes()]
        i32 = lambda *v: out.append(struct.pack("<%di" % len(v), *[int(x) for x in v]))
        arr = lambda a: out.append(np.ascontiguousarray(a).tobytes())

        def mass(m):
            if getattr(m, "ocean", False):     # marker -1000000 - rows, then doubles: mass, massOcean, theta | normal (nr x 3, column-major)
                rows = m.mass.size if m.is3D else 1
                i32(-1000000 - rows)
                if m.is3D:
                    arr(np.asarray(m.mass, np.float64)); arr(np.asarray(m.massOcean, np.float64))
                    arr(np.ascontiguousarray(m.normal.T.reshape(-1), dtype=np.float64))
                else:
                    arr(np.array([m.mass, m.massOcean, m.theta], dtype=np.float64))
                return
            v = _mass(m)
            i32(v.size)
            arr(v)
        i32(len(self.points))
        for p in self.points:
            kind = {"solid": 0, "fluid": 1}.get(p.kind, 2)
            i32(kind, p.nr, p.axial)
            arr(np.asarray(p.crds, dtype=np.float64))
            if kind == 0:
                mass(p.mass)
            elif kind == 1:
                mass(p.mass)
                i32(p.fluidSurf)
            else:
                mass(p.solid.mass)
                mass(p.fluid.mass)
                i32(p.fluid.fluidSurf)
                c = p.couple
                if c.is3D:
                    i32(p.nr)
                    arr(_colmajor(c.n_un))
                    arr(_colmajor(c.n_as))
                else:
                    i32(1)
                    arr(np.array([c.ns, 0.0, c.nz], dtype=np.float32))
                    arr(np.array([c.ns_invmf, 0.0, c.nz_invmf], dtype=np.float32))
        i32(len(self.elements))
        for e in self.elements:
            g = e.grad
            i32(0 if e.kind == "solid" else 1, g.axial)
            arr(np.array([p.domain_tag for p in e.points], dtype=np.int32))
            arr(np.stack([g.dsdxii, g.dsdeta, g.dzdxii, g.dzdeta, g.inv_s]).reshape(-1).astype(np.float64))
            prt = getattr(e, "prt", None)                 # rows of X: 0 = none, 1 = PRT_1D, Nr = PRT_3D; then [4][25][rows]
            if prt is None:
                i32(0)
            else:
                i32(prt.X.shape[1])
                arr(_f32(np.transpose(prt.X, (0, 2, 1))).reshape(-1))
            if e.kind == "solid":
                el = e.elastic
                rows = el.coef.shape[1]
                i32(_LAW[el.law], rows)
                arr(_f32(np.transpose(el.coef, (0, 2, 1))).reshape(-1))
                a = el.att
                if a is None:
                    i32(0)
                else:
                    P = 4 if a.cg4 else 25
                    i32(2 if a.cg4 else 1, a.nsls, a.doKappa)
                    arr(_f32(a.alpha)); arr(_f32(a.beta)); arr(_f32(a.gamma))
                    arr(_colmajor(np.asarray(a.dkappa).reshape(rows, P)))
                    arr(_colmajor(np.asarray(a.dmu).reshape(rows, P)))
            else:
                K = e.acoustic.K
                i32(K.shape[0])
                arr(_colmajor(K))
        i32(len(self.sources))
        for st in self.sources:
            i32(st.element.domain_tag)
            arr(np.array([f.shape[0] for f in st.force], dtype=np.int32))
            for f in st.force:
                c = np.asarray(f, dtype=np.complex64).T.reshape(-1)       # column-major (nrow x 3)
                arr(np.stack([c.real, c.imag], 1).reshape(-1).astype(np.float32))
        stf = _f32(stf)
        i32(stf.size)
        out.append(struct.pack("<d", float(dt)))
        arr(stf)
        with open(path, "wb") as f:
            f.write(b"".join(out))


def read_displacement(path, points):
    """host_driver output -> {tag: (solid (Nu+1, 3) or None, fluid (Nu+1,) or None)}"""
    raw = np.fromfile(path, dtype=np.complex64)
    pos, res = 0, {}
    for p in points:
        n = p.nu + 1
        s = f = None
        if p.kind in ("solid", "solidfluid"):
            s = raw[pos:pos + 3 * n].reshape(3, n).T
            pos += 3 * n
        if p.kind in ("fluid", "solidfluid"):
            f = raw[pos:pos + n]
            pos += n
        res[p.domain_tag] = (s, f)
    assert pos == raw.size, (pos, raw.size)
    return res
