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


def parse_dump(raw):
    """The AX3D serialisation back as plain arrays: {"G", "points": [...], "elements": [...], "sources": [...], "dt", "stf",
    "receivers": {...} or None}.  Reads what DumpDomain.write produces and what oracle/ref_main_dump.cpp writes from the
    reference's own Domain (which appends a receiver section)."""
    raw = bytes(raw)
    pos = [0]

    def take(dtype, n=1):
        a = np.frombuffer(raw, dtype=dtype, count=n, offset=pos[0])
        pos[0] += a.nbytes
        return a

    def i32():
        return int(take("<i4")[0])

    assert raw[:4] == b"AX3D"
    pos[0] = 4
    res = {"G": (take("<f8", 25).copy(), take("<f8", 25).copy())}

    def mass():
        n = i32()
        if n <= -1000000:
            rows = -1000000 - n
            return {"ocean": True, "data": take("<f8", 3 if rows == 1 else 5 * rows).copy()}
        return take("<f4", n).copy()

    pts = []
    for _ in range(i32()):
        kind, nr, axial = i32(), i32(), i32()
        p = {"kind": kind, "nr": nr, "axial": axial, "crds": take("<f8", 2).copy()}
        if kind == 0:
            p["mass"] = mass()
        elif kind == 1:
            p["mass"] = mass()
            p["fluidSurf"] = i32()
        else:
            p["mass"] = mass()
            p["mass_fluid"] = mass()
            p["fluidSurf"] = i32()
            rows = i32()
            p["n_un"] = take("<f4", 3 * rows).reshape(3, rows).copy()
            p["n_as"] = take("<f4", 3 * rows).reshape(3, rows).copy()
        pts.append(p)
    res["points"] = pts
    els = []
    ncoef = {0: 2, 1: 5, 2: 21}
    for _ in range(i32()):
        e = {"fluid": i32(), "axial": i32(), "tags": take("<i4", 25).copy(), "grad": take("<f8", 125).reshape(5, 25).copy()}
        prt_rows = i32()
        if prt_rows:
            e["prt"] = take("<f4", 4 * 25 * prt_rows).reshape(4, 25, prt_rows).copy()
        if not e["fluid"]:
            law, rows = i32(), i32()
            e["law"], e["rows"] = law, rows
            e["coef"] = take("<f4", ncoef[law] * 25 * rows).reshape(ncoef[law], 25, rows).copy()
            att = i32()
            e["att"] = att
            if att:
                nsls, do_kappa = i32(), i32()
                P = 4 if att == 2 else 25
                e.update(nsls=nsls, doKappa=do_kappa, alpha=take("<f4", nsls).copy(), beta=take("<f4", nsls).copy(),
                         gamma=take("<f4", nsls).copy(), dkappa=take("<f4", rows * P).reshape(P, rows).copy(),
                         dmu=take("<f4", rows * P).reshape(P, rows).copy())
        else:
            rows = i32()
            e["rows"] = rows
            e["K"] = take("<f4", 25 * rows).reshape(25, rows).copy()
        els.append(e)
    res["elements"] = els
    srcs = []
    for _ in range(i32()):
        tag = i32()
        nrow = take("<i4", 25).copy()
        force = []
        for k in range(25):
            c = take("<f4", 2 * 3 * int(nrow[k])).reshape(3, int(nrow[k]), 2)
            force.append((c[..., 0] + 1j * c[..., 1]).T.copy())          # (nrow, 3)
        srcs.append({"element": tag, "force": force})
    res["sources"] = srcs
    n = i32()
    res["dt"] = float(take("<f8")[0])
    res["stf"] = take("<f4", n).copy()
    res["receivers"] = None
    if raw[pos[0]:pos[0] + 4] == b"RECV":
        pos[0] += 4
        rec = {"shift": float(take("<f8")[0])}
        nrec = i32()
        rec["components"] = raw[pos[0]:pos[0] + 3].decode() if nrec else ""
        pos[0] += 3 if nrec else 0
        keys, tags, par, w = [], [], [], []
        for _ in range(nrec):
            ln = i32()
            keys.append(raw[pos[0]:pos[0] + ln].decode())
            pos[0] += ln
            tags.append(i32())
            par.append(take("<f8", 6).copy())          # phi, theta, baz, lat, lon, dep
            w.append(take("<f8", 25).copy())
        rec.update(keys=keys, element=np.array(tags, dtype=np.int64), par=np.array(par).reshape(-1, 6),
                   weights=np.array(w).reshape(-1, 25))
        res["receivers"] = rec
    assert pos[0] == len(raw), (pos[0], len(raw))
    return res
