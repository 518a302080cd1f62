"""`Domain` with the reference's verbs (S/core/domain/Domain.h:27-89) over the CUDA C-ABI.

Mirrors the call sites of Newmark::solve (Newmark.cpp:47-93) and Mesh::release (Mesh.cpp:177-208):
addPoint / addElement / addSourceTerm / setMessaging, then updateNewmark, applySource, computeStiff,
coupleSolidFluid, assembleStiff, checkStability.  All arithmetic happens in libaxisem3d_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from . import model as M

_FIELD = {"displ": 0, "veloc": 1, "accel": 2, "stiff": 3}
_LAW = {"iso": 0, "ti": 1, "aniso": 2}


def _pf(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _pi(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _mass(m):
    if m.is3D:
        return _f32(m.invMass)
    return np.array([m.invMass], dtype=np.float32)


def _colmajor(a):
    """(rows, ncol) array -> Eigen column-major flat float32."""
    return _f32(np.asarray(a).T).reshape(-1)


def nccl_unique_id():
    """128-byte ncclUniqueId created on this rank (rank 0 creates it, the host broadcasts it)."""
    lib = capi.load()
    buf = (C.c_ubyte * 128)()
    capi.check(lib.ax3d_nccl_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf)


class Domain:
    def __init__(self, device=0):
        self.lib = capi.load()
        h = C.c_void_p()
        capi.check(self.lib.ax3d_create(int(device), C.byref(h)))
        self.h = h
        self.points, self.elements = [], []
        self._final = False

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.ax3d_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ setup
    def setGMat(self, G_GLL, G_GLJ):
        a = np.ascontiguousarray(np.asarray(G_GLL, dtype=np.float64).reshape(25))
        b = np.ascontiguousarray(np.asarray(G_GLJ, dtype=np.float64).reshape(25))
        capi.check(self.lib.ax3d_set_gmat(self.h, _pd(a), _pd(b)))

    def addPoint(self, p):
        tag = C.c_int(-1)
        crds = np.ascontiguousarray(p.crds, dtype=np.float64)
        if p.kind == "solid" and getattr(p.mass, "ocean", False):
            m = p.mass
            if m.is3D:
                a, b, c = np.ascontiguousarray(m.mass), np.ascontiguousarray(m.massOcean), np.ascontiguousarray(m.normal.T.reshape(-1))
            else:
                a, b, c = (np.array([v], dtype=np.float64) for v in (m.mass, m.massOcean, m.theta))
            capi.check(self.lib.ax3d_add_solid_point_ocean(self.h, p.nr, int(p.axial), _pd(crds), a.size, _pd(a), _pd(b), _pd(c),
                                                           C.byref(tag)))
        elif p.kind == "solid":
            im = _mass(p.mass)
            capi.check(self.lib.ax3d_add_solid_point(self.h, p.nr, int(p.axial), _pd(crds), im.size, _pf(im), C.byref(tag)))
        elif p.kind == "fluid":
            im = _mass(p.mass)
            capi.check(self.lib.ax3d_add_fluid_point(self.h, p.nr, int(p.axial), _pd(crds), im.size, _pf(im),
                                                     int(p.fluidSurf), C.byref(tag)))
        else:
            ims, imf = _mass(p.solid.mass), _mass(p.fluid.mass)
            c = p.couple
            if c.is3D:
                nun, nas, nsf = _colmajor(c.n_un), _colmajor(c.n_as), p.nr
            else:
                nun = np.array([c.ns, 0.0, c.nz], dtype=np.float32)
                nas = np.array([c.ns_invmf, 0.0, c.nz_invmf], dtype=np.float32)
                nsf = 1
            capi.check(self.lib.ax3d_add_solid_fluid_point(
                self.h, p.nr, int(p.axial), _pd(crds), ims.size, _pf(ims), imf.size, _pf(imf),
                int(p.fluid.fluidSurf), nsf, _pf(nun), _pf(nas), C.byref(tag)))
        p.domain_tag = tag.value
        self.points.append(p)
        return tag.value

    def addElement(self, e):
        tag = C.c_int(-1)
        tags = np.array([p.domain_tag for p in e.points], dtype=np.int32)
        g = e.grad
        geom = np.ascontiguousarray(np.stack([g.dsdxii, g.dsdeta, g.dzdxii, g.dzdeta, g.inv_s]).reshape(-1), dtype=np.float64)
        if e.kind == "solid":
            el = e.elastic
            rows = el.coef.shape[1]
            # (ncoef, rows, 25) -> per array column-major rows x 25
            coef = _f32(np.transpose(el.coef, (0, 2, 1))).reshape(-1)
            theta = np.ascontiguousarray(e.formThetaMat().reshape(-1), dtype=np.float64)
            att_ref = None
            keep = []
            if el.att is not None:
                a = el.att
                P = 4 if a.cg4 else 25
                al, be, ga = _f32(a.alpha), _f32(a.beta), _f32(a.gamma)
                dk = _colmajor(np.asarray(a.dkappa).reshape(rows, P))
                dm = _colmajor(np.asarray(a.dmu).reshape(rows, P))
                keep = [al, be, ga, dk, dm]
                att_ref = capi.Attenuation(2 if a.cg4 else 1, a.nsls, _pf(al), _pf(be), _pf(ga), _pf(dk), _pf(dm),
                                           int(a.doKappa))
            capi.check(self.lib.ax3d_add_solid_element(
                self.h, _pi(tags), _pd(geom), int(g.axial), _pd(theta), _LAW[el.law], rows, _pf(coef),
                C.byref(att_ref) if att_ref is not None else None, C.byref(tag)))
            del keep
        else:
            K = e.acoustic.K
            rows = K.shape[0]
            Kf = _colmajor(K)
            capi.check(self.lib.ax3d_add_fluid_element(self.h, _pi(tags), _pd(geom), int(g.axial), rows, _pf(Kf), C.byref(tag)))
        if getattr(e, "prt", None) is not None:          # the PRT* constructor argument (Quad.cpp:386-420)
            X = _f32(np.transpose(e.prt.X, (0, 2, 1))).reshape(-1)                 # [4][25][rows]
            theta = np.ascontiguousarray(e.formThetaMat().reshape(-1), dtype=np.float64)
            capi.check(self.lib.ax3d_set_element_prt(self.h, tag.value, int(e.prt.X.shape[1]), _pf(X), _pd(theta)))
        e.domain_tag = tag.value
        self.elements.append(e)
        return tag.value

    def addSourceTerm(self, st):
        nrow = np.array([f.shape[0] for f in st.force], dtype=np.int32)
        flat = []
        for f in st.force:
            c = np.asarray(f, dtype=np.complex64).T.reshape(-1)     # column-major (nrow x 3)
            flat.append(np.stack([c.real, c.imag], 1).reshape(-1))
        force = _f32(np.concatenate(flat))
        capi.check(self.lib.ax3d_add_source_term(self.h, st.element.domain_tag, _pi(nrow), _pf(force)))

    def setMessaging(self, info, rank=0, nproc=1, nccl_unique_id=None):
        neigh = np.array(info.mIProcComm, dtype=np.int32)
        npts = np.array(info.mNLocalPoints, dtype=np.int32)
        tags = np.array([t for l in info.mILocalPoints for t in l], dtype=np.int32)
        uid = None
        if nccl_unique_id is not None:
            uid = (C.c_ubyte * 128).from_buffer_copy(bytes(nccl_unique_id)[:128].ljust(128, b"\0"))
        capi.check(self.lib.ax3d_set_messaging(self.h, rank, nproc, C.cast(uid, C.c_void_p) if uid is not None else None,
                                               len(neigh), _pi(neigh), _pi(npts), _pi(tags)))

    def finalize(self):
        capi.check(self.lib.ax3d_finalize_setup(self.h))
        self._final = True

    # ------------------------------------------------------------------ peer-memory halo (include/axisem3d_b200.h)
    def haloExport(self, neigh_ranks):
        """This rank's receive window: {"rank", "handle" (64-byte cudaIpcMemHandle), "ptr" (device address in this process),
        "begin" {neighbour rank: start of its segment}, "total", "slot" {neighbour rank: index in my neighbour list}}."""
        nn = len(neigh_ranks)
        handle = (C.c_ubyte * 64)()
        ptr = C.c_void_p()
        begin = (C.c_longlong * (nn + 1))()
        total = C.c_longlong()
        capi.check(self.lib.ax3d_halo_export(self.h, C.cast(handle, C.c_void_p), C.byref(ptr), begin, C.byref(total)))
        return {"handle": bytes(handle), "ptr": int(ptr.value or 0), "total": int(total.value),
                "begin": {int(r): int(begin[k]) for k, r in enumerate(neigh_ranks)},
                "slot": {int(r): k for k, r in enumerate(neigh_ranks)}}

    def haloConnect(self, my_rank, neigh_ranks, windows, same_process=False):
        """windows: {neighbour rank: that rank's haloExport()}.  same_process: the neighbours' domains live in this
        process (raw device pointers instead of IPC handles)."""
        nn = len(neigh_ranks)
        pb = (C.c_longlong * nn)(*[windows[int(r)]["begin"][int(my_rank)] for r in neigh_ranks])
        pt = (C.c_longlong * nn)(*[windows[int(r)]["total"] for r in neigh_ranks])
        ps = np.array([windows[int(r)]["slot"][int(my_rank)] for r in neigh_ranks], dtype=np.int32)
        if same_process:
            ptrs = (C.c_void_p * nn)(*[windows[int(r)]["ptr"] for r in neigh_ranks])
            capi.check(self.lib.ax3d_halo_connect(self.h, nn, None, ptrs, pb, pt, _pi(ps)))
        else:
            blob = b"".join(windows[int(r)]["handle"] for r in neigh_ranks)
            buf = (C.c_ubyte * (64 * nn)).from_buffer_copy(blob)
            capi.check(self.lib.ax3d_halo_connect(self.h, nn, C.cast(buf, C.c_void_p), None, pb, pt, _pi(ps)))

    def connectHalo(self, info, rank, dist):
        """Collective over the torch.distributed group: exchange the windows and switch Domain::assembleStiff to the
        peer-memory kernels (ranks without neighbours only take part in the all_gather)."""
        neigh = [int(r) for r in info.mIProcComm]
        mine = self.haloExport(neigh) if neigh else None
        everyone = [None] * dist.get_world_size()
        dist.all_gather_object(everyone, mine)
        if neigh:
            self.haloConnect(rank, neigh, {r: everyone[r] for r in neigh})

    # ------------------------------------------------------------------ step verbs
    def updateNewmark(self, dt):
        capi.check(self.lib.ax3d_update_newmark(self.h, float(dt)))

    def applySource(self, stf):
        capi.check(self.lib.ax3d_apply_source(self.h, float(stf)))

    def computeStiff(self):
        capi.check(self.lib.ax3d_compute_stiff(self.h))

    def coupleSolidFluid(self):
        capi.check(self.lib.ax3d_couple_solid_fluid(self.h))

    def assembleStiff(self, phase=0):
        capi.check(self.lib.ax3d_assemble_stiff(self.h, int(phase)))

    def checkStability(self):
        ok = C.c_int(0)
        capi.check(self.lib.ax3d_check_stability(self.h, C.byref(ok)))
        return bool(ok.value)

    def setLearnParameters(self, invoked, cutoff, interval=1):
        """Domain::setLearnParameters (Domain.h:39)."""
        capi.check(self.lib.ax3d_set_learn_parameters(self.h, int(bool(invoked)), float(cutoff), int(interval)))

    def learnWisdom(self, tstep):
        """Domain::learnWisdom(tstep) (Domain.cpp:384-402) for verb-wise stepping; runSteps does it itself."""
        capi.check(self.lib.ax3d_learn_wisdom(self.h, int(tstep)))

    def getNuWisdom(self):
        """Point::getNuWisdom() per point tag (what Domain::dumpWisdom writes, Domain.cpp:404-440)."""
        out = np.zeros(len(self.points), dtype=np.int32)
        capi.check(self.lib.ax3d_get_nu_wisdom(self.h, _pi(out), len(out)))
        return out

    def resetZero(self):
        capi.check(self.lib.ax3d_reset_zero(self.h))

    def step(self, dt, stf):
        """One iteration of Newmark::solve's loop body (Newmark.cpp:49-91)."""
        self.updateNewmark(dt)
        self.applySource(stf)
        self.computeStiff()
        self.coupleSolidFluid()
        self.assembleStiff(-1)
        self.assembleStiff(1)

    def runSteps(self, dt, stf):
        stf = _f32(stf)
        capi.check(self.lib.ax3d_run_steps(self.h, stf.size, float(dt), _pf(stf)))

    def runStepsTimed(self, dt, stf):
        """runSteps timed with CUDA events on the library's stream; returns milliseconds."""
        stf = _f32(stf)
        ms = C.c_float(0)
        capi.check(self.lib.ax3d_run_steps_timed(self.h, stf.size, float(dt), _pf(stf), C.byref(ms)))
        return ms.value

    def runStepsRecord(self, dt, stf):
        """Newmark::solve with Domain::record after every update; returns the seismograms [nsteps][nrec][3] (s, phi, z)."""
        stf = _f32(stf)
        out = np.empty((stf.size, self._nrec, 3), dtype=np.float32)
        capi.check(self.lib.ax3d_run_steps_record(self.h, stf.size, float(dt), _pf(stf), _pf(out)))
        return out

    def setReceivers(self, elem_tags, phi, weights):
        et = np.ascontiguousarray(elem_tags, dtype=np.int32)
        ph = _f32(phi)
        w = _f32(np.asarray(weights).reshape(len(et), 25))
        self._nrec = len(et)
        capi.check(self.lib.ax3d_set_receivers(self.h, len(et), _pi(et), _pf(ph), _pf(w)))

    def record(self):
        out = np.zeros((self._nrec, 3), dtype=np.float32)
        capi.check(self.lib.ax3d_record(self.h, _pf(out)))
        return out

    def kernel_stats(self, reset=True):
        """{kernel name: (summed ms, launches, summed algorithmic bytes)} gathered while the timers were on."""
        n = C.c_int(0)
        capi.check(self.lib.ax3d_kernel_stats(self.h, -1, None, 0, None, None, None, C.byref(n), 0))
        out = {}
        for k in range(n.value):
            name = C.create_string_buffer(160)
            ms, nl, by = C.c_double(0), C.c_longlong(0), C.c_double(0)
            capi.check(self.lib.ax3d_kernel_stats(self.h, k, name, 160, C.byref(ms), C.byref(nl), C.byref(by), None,
                                                  1 if (reset and k == n.value - 1) else 0))
            out[name.value.decode()] = (ms.value, nl.value, by.value)
        return out

    def measure_costs(self, repeats=3):
        """Mesh::measure (Mesh.cpp:412-588): microseconds of one SM per element (domain tag order), measured on the device."""
        out = np.zeros(len(self.elements), dtype=np.float64)
        capi.check(self.lib.ax3d_measure_costs(self.h, int(repeats), _pd(out), len(out)))
        return out

    def synchronize(self):
        capi.check(self.lib.ax3d_synchronize(self.h))

    # ------------------------------------------------------------------ read-back / test hooks
    def get_solid(self, tag, which):
        n = self.points[tag].nu + 1
        out = np.zeros(3 * n * 2, dtype=np.float32)
        capi.check(self.lib.ax3d_get_point_field(self.h, tag, _FIELD[which], 0, _pf(out), 3 * n))
        return out.view(np.complex64).reshape(3, n).T.copy()

    def get_fluid(self, tag, which):
        n = self.points[tag].nu + 1
        out = np.zeros(n * 2, dtype=np.float32)
        capi.check(self.lib.ax3d_get_point_field(self.h, tag, _FIELD[which], 1, _pf(out), n))
        return out.view(np.complex64).copy()

    def set_solid(self, tag, which, val):
        n = self.points[tag].nu + 1
        v = np.ascontiguousarray(np.asarray(val, dtype=np.complex64).reshape(n, 3).T).view(np.float32).reshape(-1)
        capi.check(self.lib.ax3d_set_point_field(self.h, tag, _FIELD[which], 0, _pf(v), 3 * n))

    def set_fluid(self, tag, which, val):
        n = self.points[tag].nu + 1
        v = np.ascontiguousarray(np.asarray(val, dtype=np.complex64).reshape(n)).view(np.float32).reshape(-1)
        capi.check(self.lib.ax3d_set_point_field(self.h, tag, _FIELD[which], 1, _pf(v), n))

    def field_size(self, fluid):
        n = C.c_size_t(0)
        capi.check(self.lib.ax3d_field_size(self.h, int(fluid), C.byref(n)))
        return n.value

    def get_bulk(self, which, fluid):
        """All solid (fluid=False) or fluid blocks concatenated in point-tag order, complex64."""
        n = self.field_size(fluid)
        out = np.zeros(2 * n, dtype=np.float32)
        capi.check(self.lib.ax3d_get_field_bulk(self.h, _FIELD[which], int(fluid), _pf(out), n))
        return out.view(np.complex64)

    def set_bulk(self, which, fluid, val):
        v = np.ascontiguousarray(np.asarray(val, dtype=np.complex64)).view(np.float32)
        capi.check(self.lib.ax3d_set_field_bulk(self.h, _FIELD[which], int(fluid), _pf(v), v.size // 2))

    def ground_motion(self, elem_tags, phi, weights):
        et = np.ascontiguousarray(elem_tags, dtype=np.int32)
        ph = _f32(phi)
        w = _f32(np.asarray(weights).reshape(len(et), 25))
        out = np.zeros((len(et), 3), dtype=np.float32)
        capi.check(self.lib.ax3d_record_ground_motion(self.h, len(et), _pi(et), _pf(ph), _pf(w), _pf(out)))
        return out

    def strain(self, elem_tags, phi, weights):
        """Element::computeStrain after forceTIso (SolidElement.cpp:219-279): [nrec][6] Voigt strain in RTZ."""
        return self._strain_curl(self.lib.ax3d_record_strain, 6, elem_tags, phi, weights)

    def curl(self, elem_tags, phi, weights):
        """Element::computeCurl after forceTIso (SolidElement.cpp:281-345): [nrec][3]."""
        return self._strain_curl(self.lib.ax3d_record_curl, 3, elem_tags, phi, weights)

    def _strain_curl(self, fn, nout, elem_tags, phi, weights):
        et = np.ascontiguousarray(elem_tags, dtype=np.int32)
        ph = _f32(phi)
        w = _f32(np.asarray(weights).reshape(len(et), 25))
        out = np.zeros((len(et), nout), dtype=np.float32)
        capi.check(fn(self.h, len(et), _pi(et), _pf(ph), _pf(w), _pf(out)))
        return out

    # ------------------------------------------------------------------ measurement
    def launch_count(self):
        n = C.c_longlong(0)
        capi.check(self.lib.ax3d_launch_count(self.h, C.byref(n)))
        return n.value

    def work_per_step(self):
        n = C.c_longlong(0)
        capi.check(self.lib.ax3d_work_per_step(self.h, C.byref(n)))
        return n.value

    def algorithmic_bytes(self):
        out = np.zeros(3, dtype=np.float64)
        capi.check(self.lib.ax3d_algorithmic_bytes(self.h, _pd(out)))
        return out

    def enable_timers(self, on=True):
        capi.check(self.lib.ax3d_enable_timers(self.h, int(on)))

    def get_timers(self, reset=True):
        out = np.zeros(4, dtype=np.float64)
        capi.check(self.lib.ax3d_get_timers(self.h, _pd(out), int(reset)))
        return out
