"""The shipped Exodus mesh -> solver descriptors: restatement of the reference's preloop for 1-D (axisymmetric) background
models, so that BASELINE.json's configs run on `template/input/AxiSEM_prem_ani_one_crust_50.e` itself.

    ExodusModel::readRawData / formStructured / formAuxiliary   S/preloop/exodus/ExodusModel.cpp:53-247, 283-420, 580-635
    Mapping (spherical, linear, semi-spherical)                  S/preloop/spectral/mapping/*.cpp
    Quad: geometry, integral factor, Nr, normals, dt             S/preloop/mesh/Quad.cpp:39-240, 292-376, 509-715
    Material: nodal values -> GLL points, mass, moduli           S/preloop/physics/material/Material.cpp:34-95, 232-365, 521-537
    AttAxiSEM / AttBuilder                                       S/preloop/physics/attenuation/AttAxiSEM.cpp:14-57, AttBuilder.cpp:17-139
    GLLPoint::release                                            S/preloop/mesh/GLLPoint.cpp:48-128
    Mesh::getDeltaT                                              S/preloop/mesh/Mesh.cpp:111-152

The file is read with `h5lite` (no HDF5 / NetCDF library in this image).  The class offers the interface of
`mesh_synth.SynthMesh` (conn, e2g, e_nr, is_fluid, axial, release(), make_source(), estimate_dt(), work_per_step()), so the
tests and bench.py run unchanged on it.  `nu` is the constant Fourier order of inparam.nu (NU_TYPE constant) or a function
nu_fn(s, z); `model3d` adds the same synthetic azimuthal perturbation SynthMesh uses (volumetric 3-D models -- s40rts,
crust1, EMC files -- need their data files, which the checkout does not ship) so that the 3-D element kinds run on the real
geometry too.  Ellipticity, ocean load and particle relabelling are off, like the template's inparam.model.
"""
from __future__ import annotations

import numpy as np

from . import connectivity as CN
from . import h5lite
from . import model as M
from . import spectral as SP

Q2 = [np.array([[-1.0, 0.0], [0.0, -1.0]]), np.array([[0.0, -1.0], [1.0, 0.0]]),
      np.array([[1.0, 0.0], [0.0, 1.0]]), np.array([[0.0, 1.0], [-1.0, 0.0]])]     # Mapping::sOrthogQ2


def _p4(p):
    return p % 4


def _make_close(a, b):
    if a - b > np.pi:
        b += 2 * np.pi
    if b - a > np.pi:
        a += 2 * np.pi
    return a, b


def _strings(a):
    return [b"".join(r).split(b"\0")[0].decode().strip() for r in a]


class ExodusMesh:
    def __init__(self, path, nu=2, nu_fn=None, lucky=True, attenuation="cg4", model3d=False, perturb=0.02, fluid3d=False,
                 dtype_coef=np.float64, do_kappa=True, volumetric=None, src=None, geodesy=None, ocean_depth=0.0):
        self.path = path
        self.nu, self.nu_fn, self.lucky = nu, nu_fn, bool(lucky)
        self.att_kind = attenuation
        self.do_kappa = bool(do_kappa)          # ATTENUATION_QKAPPA (AttAxiSEM.cpp:43-50)
        self.ocean_depth = float(ocean_depth)   # MODEL_3D_OCEAN_LOAD constant$<km> in metres (OceanLoad3D_const; 0 = none)
        # 3-D volumetric models of inparam.model (volumetric.from_parameters) with the source they are rotated about
        # (`volumetric` may be a callable(mesh) -> (models, src, geodesy), evaluated once the file's globals are read)
        self._vol_factory = volumetric if callable(volumetric) else None
        self.volumetric, self.vol_src, self.vol_geodesy = ([] if callable(volumetric) else list(volumetric or [])), src, geodesy
        self.model3d, self.perturb, self.fluid3d = bool(model3d), float(perturb), bool(fluid3d)
        self.perturb_rho = False
        self.dtype_coef = dtype_coef
        self._read()
        self.geometric = []                     # 3-D geometric models (relabelling.from_parameters), 4th item of the factory's result
        if self._vol_factory is not None:
            res = self._vol_factory(self)
            self.volumetric, self.vol_src, self.vol_geodesy = res[:3]
            self.geometric = list(res[3]) if len(res) > 3 else []
        self._auxiliary()
        self._build_quads()
        self._build_points()

    # ------------------------------------------------------------------ ExodusModel::readRawData
    def _read(self):
        f = h5lite.File(self.path)
        gnames = _strings(f["name_glo_var"].read())
        gvals = f["vals_glo_var"].read().reshape(-1)
        self.glob = dict(zip(gnames, [float(v) for v in gvals]))
        self.records = {}
        for rec in _strings(f["info_records"].read()):
            if "=" in rec:
                k, v = rec.split("=", 1)
                self.records[k.strip()] = v.strip()
        if self.records.get("crdsys", "spherical").lower() == "cartesian":
            raise NotImplementedError("ExodusMesh: Cartesian meshes")
        self.conn = f["connect1"].read().astype(np.int64) - 1
        self.nodal_s = f["coordx"].read().astype(np.float64)
        self.nodal_z = f["coordy"].read().astype(np.float64)
        self.nelem = self.conn.shape[0]
        c = self.conn
        s, z = self.nodal_s, self.nodal_z
        d = [np.hypot(s[c[:, i]] - s[c[:, (i + 1) % 4]], z[c[:, i]] - z[c[:, (i + 1) % 4]]) for i in range(4)]
        self.dist_tol = float(np.min(d)) / 1000.0
        # side sets: value = side index of the element on the set, -1 elsewhere
        self.side_sets = {}
        for k, name in enumerate(_strings(f["ss_names"].read())):
            el = f["elem_ss%d" % (k + 1)].read().astype(np.int64) - 1
            sd = f["side_ss%d" % (k + 1)].read().astype(np.int64) - 1
            v = np.full(self.nelem, -1, dtype=np.int64)
            v[el] = sd
            self.side_sets[name] = v
        self.ss_axis = "t1" if "t1" in self.side_sets else "t0"
        self.ss_surface = "r1"
        names = _strings(f["name_elem_var"].read())
        self.var_names = names

        def var(name):
            return f["vals_elem_var%deb1" % (names.index(name) + 1)].read().reshape(-1).astype(np.float64)

        # depth-dependent variables are kept on the axis only and looked up by radius (ExodusModel.cpp:146-228)
        axial = self.side_sets[self.ss_axis]
        coords, quad_nodes = [], []
        import bisect
        for iq in range(self.nelem):
            side = int(axial[iq])
            if side < 0:
                continue
            other = 0 if side == 3 else side + 1
            z1, z2 = z[c[iq, side]], z[c[iq, other]]
            if min(z1, z2) < -self.dist_tol:
                continue
            z1 += (z2 - z1) / abs(z2 - z1) * self.dist_tol
            z2 -= (z2 - z1) / abs(z2 - z1) * self.dist_tol
            k = bisect.bisect_right(coords, z1)
            coords.insert(k, z1)
            quad_nodes.insert(k, (iq, side))
            k = bisect.bisect_right(coords, z2)
            coords.insert(k, z2)
            quad_nodes.insert(k, (iq, other))
        self.axis_coords = np.array(coords)
        qn = np.array(quad_nodes, dtype=np.int64)
        self.axis_vars = {}
        for name in names:
            if len(name) > 2 and name[-2] == "_" and name[-1] in "0123":
                base = name[:-2]
                if name[-1] != "0":
                    continue
                bufs = [var("%s_%d" % (base, i)) for i in range(4)]
                self.axis_vars[base] = np.array([bufs[n][q] for q, n in qn])
            else:
                buf = var(name)
                self.axis_vars[name] = buf[qn[:, 0]]
        self.elem_type = var("element_type")
        self.elem_dt = var("dt")
        self.isotropic = "VP_0" in names
        self.has_att = "QMU_0" in names and "nr_lin_solids" in self.glob
        self.r_outer = self.glob.get("radius", 6371e3)
        # radial ellipticity profile (ExodusModel.cpp:122-126): row 0 = knots r / R, row 1 = coefficients; used by preloop.Geodesy
        if "ellipticity" in f:
            ell = f["ellipticity"].read().astype(np.float64)
            self.ellip_knots, self.ellip_coeffs = ell[0].copy(), ell[1].copy()
        else:
            self.ellip_knots = self.ellip_coeffs = None
        if self.has_att:
            n = int(self.glob["nr_lin_solids"])
            self.sls_w = np.array([self.glob["w_%d" % i] for i in range(n)])
            self.sls_y = np.array([self.glob["y_%d" % i] for i in range(n)])
            self.f_min, self.f_max, self.f_ref = self.glob["f_min"], self.glob["f_max"], self.glob["f_ref"]

    def elem_var(self, name, iq):
        """ExodusModel::getElementalVariables (ExodusModel.cpp:580-635): nearest axis sample to the node's radius, moved
        towards the element centre by the distance tolerance."""
        c = self.conn[iq]
        r = np.hypot(self.nodal_s[c], self.nodal_z[c])
        cen = r.mean()
        if len(name) > 2 and name[-2] == "_" and name[-1] in "0123":
            coord = r[int(name[-1])]
            coord += self.dist_tol if coord < cen else -self.dist_tol
            return self.axis_vars[name[:-2]][int(np.argmin(np.abs(self.axis_coords - coord)))]
        return self.axis_vars[name][int(np.argmin(np.abs(self.axis_coords - cen)))]

    # ------------------------------------------------------------------ ExodusModel::formAuxiliary
    def _auxiliary(self):
        c, s, z = self.conn, self.nodal_s, self.nodal_z
        nnode = len(s)
        # average GLL spacing per node
        per = sum(np.hypot(s[c[:, i]] - s[c[:, (i + 1) % 4]], z[c[:, i]] - z[c[:, (i + 1) % 4]]) for i in range(4)) / 4.0 / CN.nPol
        cnt = np.zeros(nnode)
        acc = np.zeros(nnode)
        for i in range(4):
            np.add.at(cnt, c[:, i], 1.0)
            np.add.at(acc, c[:, i], per)
        self.ave_gll_spacing = acc / cnt
        # axial elements: rotate the nodes so that side 3 lies on the axis (ExodusModel.cpp:378-403)
        ax = self.side_sets[self.ss_axis]
        for iq in range(self.nelem):
            side = int(ax[iq])
            if side in (3, -1):
                continue
            con = c[iq].copy()
            for j in range(4):
                c[iq, j] = con[_p4(j + side - 3)]
            for v in self.side_sets.values():
                if v[iq] != -1:
                    v[iq] = _p4(int(v[iq]) - side + 3)
        self.axial = self.side_sets[self.ss_axis] >= 0
        # elements that touch the axis region without being axial (for the odd-Nr rule)
        near = np.zeros(nnode, dtype=bool)
        near[c[self.axial].reshape(-1)] = True
        self.vicinal = np.full((self.nelem, 4), -1, dtype=np.int64)
        for iq in range(self.nelem):
            if self.axial[iq]:
                continue
            for j in range(4):
                if near[c[iq, j]]:
                    self.vicinal[iq, j] = j

    # ------------------------------------------------------------------ mappings
    def _map(self, iq, xi, eta, jac=False):
        """Quad::mapping / jacobian at (xi, eta) arrays -> (s, z) or J[2][2] arrays."""
        nodes = self.nodes[iq]
        kind, co = self.map_kind[iq], self.curved_outer[iq]
        xi, eta = np.asarray(xi, dtype=np.float64), np.asarray(eta, dtype=np.float64)
        if kind == 1:      # LinearMapping
            xm, xp, em, ep = 1 - xi, 1 + xi, 1 - eta, 1 + eta
            if not jac:
                shp = np.stack([xm * em, xp * em, xp * ep, xm * ep]) / 4.0
                return np.tensordot(nodes, shp, axes=(1, 0))
            d0 = np.stack([-em, em, ep, -ep]) / 4.0
            d1 = np.stack([-xm, -xp, xp, xm]) / 4.0
            return np.stack([np.tensordot(nodes, d0, axes=(1, 0)), np.tensordot(nodes, d1, axes=(1, 0))], axis=1)   # J[:, 0] = d/dxi
        Q = Q2[co]
        n2 = Q @ nodes
        x2 = Q[0, 0] * xi + Q[0, 1] * eta
        e2 = Q[1, 0] * xi + Q[1, 1] * eta
        r = np.hypot(n2[0], n2[1])
        t = np.arctan2(n2[0], n2[1])
        i0, i1, i2, i3 = _p4(co - 2), _p4(co - 1), _p4(co), _p4(co + 1)
        t2, t3 = _make_close(t[i2], t[i3])
        r3 = r[i3]
        ang = ((1 - x2) * t3 + (1 + x2) * t2) / 2.0
        if kind == 0:      # SphericalMapping
            r0 = r[i0]
            t0, t1 = _make_close(t[i0], t[i1])
            ang0 = ((1 - x2) * t0 + (1 + x2) * t1) / 2.0
            if not jac:
                sz2 = np.stack([(1 + e2) * r3 / 2 * np.sin(ang) + (1 - e2) * r0 / 2 * np.sin(ang0),
                                (1 + e2) * r3 / 2 * np.cos(ang) + (1 - e2) * r0 / 2 * np.cos(ang0)])
                return np.tensordot(Q.T, sz2, axes=(1, 0))
            J00 = (1 + e2) * r3 * (t2 - t3) / 4 * np.cos(ang) + (1 - e2) * r0 * (t1 - t0) / 4 * np.cos(ang0)
            J01 = 0.5 * (r3 * np.sin(ang) - r0 * np.sin(ang0))
            J10 = -(1 + e2) * r3 * (t2 - t3) / 4 * np.sin(ang) - (1 - e2) * r0 * (t1 - t0) / 4 * np.sin(ang0)
            J11 = 0.5 * (r3 * np.cos(ang) - r0 * np.cos(ang0))
        else:              # SemiSphericalMapping
            s0, z0, s1, z1 = n2[0, i0], n2[1, i0], n2[0, i1], n2[1, i1]
            if not jac:
                sz2 = np.stack([(1 + e2) * r3 / 2 * np.sin(ang) + (1 - e2) / 2 * (((1 - x2) * s0 + (1 + x2) * s1) / 2),
                                (1 + e2) * r3 / 2 * np.cos(ang) + (1 - e2) / 2 * (((1 - x2) * z0 + (1 + x2) * z1) / 2)])
                return np.tensordot(Q.T, sz2, axes=(1, 0))
            J00 = (1 - e2) * (s1 - s0) / 4 + (1 + e2) * r3 * (t2 - t3) * np.cos(ang) / 4
            J01 = -((1 - x2) * s0 + (1 + x2) * s1) / 4 + r3 / 2 * np.sin(ang)
            J10 = (1 - e2) * (z1 - z0) / 4 - (1 + e2) * r3 * (t2 - t3) * np.sin(ang) / 4
            J11 = -((1 - x2) * z0 + (1 + x2) * z1) / 4 + r3 / 2 * np.cos(ang)
        J2 = np.array([[J00, J01], [J10, J11]])
        return np.einsum("ab,bc...,cd->ad...", Q.T, J2, Q)

    # ------------------------------------------------------------------ Quad::Quad
    def _build_quads(self):
        c, s, z = self.conn, self.nodal_s, self.nodal_z
        ne = self.nelem
        self.nodes = np.stack([s[c], z[c]], axis=1)                 # [ne][2][4]
        self.map_kind = np.zeros(ne, dtype=np.int64)
        self.curved_outer = np.full(ne, -1, dtype=np.int64)
        tol = self.dist_tol
        for iq in range(ne):
            r = np.hypot(self.nodes[iq, 0], self.nodes[iq, 1])
            et = self.elem_type[iq]
            if et < 0.5:
                self.map_kind[iq] = 0
                if abs(r[0] - r[1]) < tol and abs(r[2] - r[3]) < tol:
                    self.curved_outer[iq] = 2 if r[2] > r[0] else 0
                elif abs(r[1] - r[2]) < tol and abs(r[3] - r[0]) < tol:
                    self.curved_outer[iq] = 3 if r[3] > r[1] else 1
                else:
                    raise RuntimeError("Quad::Quad || Invalid spherical element shape.")
            elif et < 1.5:
                self.map_kind[iq] = 1
            else:
                self.map_kind[iq] = 2
                for k in range(4):
                    if abs(r[k] - r[(k + 1) % 4]) < tol and r[k] > r[(k + 2) % 4]:
                        self.curved_outer[iq] = k
                        break
                else:
                    raise RuntimeError("Quad::Quad || Invalid semi-spherical element shape.")
        fl = np.array([self.elem_var("fluid", iq) for iq in range(ne)])
        self.is_fluid = fl > 0.5
        self.sf_side = self.side_sets.get("solid_fluid_boundary", np.full(ne, -1, dtype=np.int64))
        self.surf_side = self.side_sets.get(self.ss_surface, np.full(ne, -1, dtype=np.int64))
        if np.any(self.axial & (self.side_sets[self.ss_axis] != 3)):
            raise RuntimeError("Quad::Quad || Axial side must be 3.")
        self.neighbours = CN.form_neighbourhood(self.conn)
        self.ngll, self.e2g = CN.form_elem_to_gll(self.conn, self.neighbours)
        # geometry on the 5 x 5 points
        self.geo = []
        for iq in range(ne):
            xi1 = SP.P_GLJ if self.axial[iq] else SP.P_GLL
            xi = xi1[:, None] * np.ones((1, 5))
            eta = np.ones((5, 1)) * SP.P_GLL[None, :]
            sz = self._map(iq, xi, eta)
            J = self._map(iq, xi, eta, jac=True)
            det = J[0, 0] * J[1, 1] - J[0, 1] * J[1, 0]
            sg = sz[0].copy()
            if self.axial[iq]:
                sg[0, :] = 0.0
            self.geo.append(dict(s=sg, z=sz[1], J00=J[0, 0], J01=J[0, 1], J10=J[1, 0], J11=J[1, 1], det=det, xi=xi1, xi2=xi, eta2=eta))

    # ------------------------------------------------------------------ Quad::formNrField, integral factor, masses, normals
    def _nr_at(self, iq, ip, jp):
        g = self.geo[iq]
        s, z = g["s"][ip, jp], g["z"][ip, jp]
        nu = self.nu_fn(s, z) if self.nu_fn is not None else self.nu
        nr = 2 * int(nu) + 1                                       # ConstNrField: 2 Nu + 1
        xi, eta = g["xi2"][ip, jp], g["eta2"][ip, jp]
        shp = np.array([(1 - xi) * (1 - eta), (1 + xi) * (1 - eta), (1 + xi) * (1 + eta), (1 - xi) * (1 + eta)]) / 4.0
        spacing = float(self.ave_gll_spacing[self.conn[iq]] @ shp)
        upper = max(int(2 * np.pi * s / spacing), 3)
        nr = min(nr, upper)
        force_odd = False
        if nr % 2 == 0:
            if self.axial[iq]:
                nr += 1
                force_odd = True
            elif self.vicinal[iq].max() >= 0:
                v = self.vicinal[iq]
                for i in range(4):
                    n0, n1 = v[i], v[(i + 1) % 4]
                    if n0 >= 0 and n1 >= 0:
                        on = (n0 == 0 and jp == 0) or (n0 == 1 and ip == 4) or (n0 == 2 and jp == 4) or (n0 == 3 and ip == 0)
                    elif n0 >= 0 and n1 < 0:
                        on = (n0 == 0 and ip == 0 and jp == 0) or (n0 == 1 and ip == 4 and jp == 0) or \
                             (n0 == 2 and ip == 4 and jp == 4) or (n0 == 3 and ip == 0 and jp == 4)
                    else:
                        on = False
                    if on:
                        nr += 1
                        force_odd = True
                        break
        if self.lucky:
            nr = SP.next_lucky_number(nr, force_odd)
        return nr

    def _nodal(self, base, iq):
        return np.array([self.elem_var("%s_%d" % (base, i), iq) for i in range(4)])

    def _interp(self, nodal, iq):
        g = self.geo[iq]
        xi, eta = g["xi2"], g["eta2"]
        return (nodal[0] * (1 - xi) * (1 - eta) + nodal[1] * (1 + xi) * (1 - eta) + nodal[2] * (1 + xi) * (1 + eta) +
                nodal[3] * (1 - xi) * (1 + eta)) / 4.0

    def _build_points(self):
        ne, ng = self.nelem, self.ngll
        # per-point Nr: the reference gives a shared point the value of the last quad that sets it up (GLLPoint::setup)
        self.p_nr = np.zeros(ng, dtype=np.int64)
        self.p_s, self.p_z = np.zeros(ng), np.zeros(ng)
        self.p_axis = np.zeros(ng, dtype=bool)
        self.e_pnr = np.zeros((ne, 5, 5), dtype=np.int64)
        for iq in range(ne):
            g, tags = self.geo[iq], self.e2g[iq]
            for ip in range(5):
                for jp in range(5):
                    self.e_pnr[iq, ip, jp] = self._nr_at(iq, ip, jp)
            self.p_s[tags], self.p_z[tags] = g["s"], g["z"]
            if self.axial[iq]:
                self.p_axis[tags[0, :]] = True
        for iq in range(ne):
            np.maximum.at(self.p_nr, self.e2g[iq].reshape(-1), self.e_pnr[iq].reshape(-1))
        self.e_nr = np.array([self.p_nr[self.e2g[iq]].max() for iq in range(ne)], dtype=np.int64)
        # material on the GLL points (Material::Material)
        self.mat = []
        for iq in range(ne):
            m = {}
            if self.isotropic:
                m["vpv"] = m["vph"] = self._interp(self._nodal("VP", iq), iq)
                m["vsv"] = m["vsh"] = self._interp(self._nodal("VS", iq), iq)
                m["eta"] = np.ones((5, 5))
                m["vmax_ref"] = self._nodal("VP", iq).max()
            else:
                for k, nm in (("vpv", "VPV"), ("vph", "VPH"), ("vsv", "VSV"), ("vsh", "VSH"), ("eta", "ETA")):
                    m[k] = self._interp(self._nodal(nm, iq), iq)
                m["vmax_ref"] = max(self._nodal("VPV", iq).max(), self._nodal("VPH", iq).max())
            m["rho"] = self._interp(self._nodal("RHO", iq), iq)
            if self.has_att:
                m["qkp"] = self._interp(self._nodal("QKAPPA", iq), iq)
                m["qmu"] = self._interp(self._nodal("QMU", iq), iq)
            self.mat.append(m)
        # 3-D volumetric models (Material::addVolumetric3D): None where no sample of the quad is in range of a model
        self.vol = [None] * ne
        if self.volumetric:
            from . import volumetric as VOL
            from .preloop import Geodesy
            geo = self.vol_geodesy or Geodesy(self.r_outer)
            for iq in range(ne):
                m, g = self.mat[iq], self.geo[iq]
                ref1d = {k: m[k].reshape(25) if k in m else np.zeros(25) for k in VOL.KEYS}
                self.vol[iq] = VOL.apply(self.volumetric, geo, self.vol_src, ref1d, g["s"].reshape(25), g["z"].reshape(25),
                                         int(self.e_nr[iq]), bool(self.is_fluid[iq]))
        # particle relabelling (Quad::addGeometric3D -> Relabelling::addUndulation): None where the quad is not displaced
        self.relab = [None] * ne
        if self.geometric:
            from .relabelling import Relabelling
            from .preloop import Geodesy
            geo = self.vol_geodesy or Geodesy(self.r_outer)
            for iq in range(ne):
                R = Relabelling(self, iq, self.geometric, geo, self.vol_src)
                if not R.zero:
                    self.relab[iq] = R
        # integral factor (Quad::formIntegralFactor)
        self.ifact = []
        for iq in range(ne):
            g = self.geo[iq]
            wxi = SP.W_GLJ if self.axial[iq] else SP.W_GLL
            w = wxi[:, None] * SP.W_GLL[None, :]
            if self.axial[iq]:
                f = w * g["s"] / (1.0 + g["xi"])[:, None].clip(1e-300) * g["det"]
                f[0, :] = (w * g["J00"] * g["det"])[0, :]
            else:
                f = w * g["s"] * g["det"]
            self.ifact.append(f)
        # masses and solid-fluid normals per GLL point (Quad::setupGLLPoints)
        self.mass_s = [np.zeros(n) for n in self.p_nr]
        self.mass_f = [np.zeros(n) for n in self.p_nr]
        self.sf_n = [None] * ng
        self.sf_contrib = {}
        self.p_surface = np.zeros(ng, dtype=bool)
        self.surf_n = {}
        for iq in range(ne):
            tags, m, f = self.e2g[iq], self.mat[iq], self.ifact[iq]
            for ip in range(5):
                for jp in range(5):
                    t = tags[ip, jp]
                    rho, vp = self._rho_vp(iq, ip, jp, int(self.p_nr[t]))
                    jm = 1.0 if self.relab[iq] is None else self.relab[iq].mass_jacobian(ip * 5 + jp)      # Material::computeElementalMass
                    if self.is_fluid[iq]:
                        self.mass_f[t] += f[ip, jp] / (rho * vp ** 2) * jm
                    else:
                        self.mass_s[t] += f[ip, jp] * rho * jm
            side = int(self.sf_side[iq])
            if side >= 0:
                for (ip, jp) in CN.EDGE_IJ[side]:
                    t = tags[ip, jp]
                    n = self._normal(iq, side, ip, jp)
                    if not self.is_fluid[iq]:
                        n = -n
                    if self.sf_n[t] is None:
                        self.sf_n[t] = np.zeros((self.p_nr[t], 3))
                    n = n[None, :] if n.ndim == 1 else n           # [Nr_p][3] under particle relabelling
                    self.sf_n[t] += 0.5 * n
                    self.sf_contrib.setdefault(int(t), []).append((iq, 0.5 * n))
            side = int(self.surf_side[iq])
            if side >= 0:
                for (ip, jp) in CN.EDGE_IJ[side]:
                    self.p_surface[tags[ip, jp]] = True
                    if self.ocean_depth > 0.0:         # GLLPoint::addSurfNormal: the surface area the ocean column stands on
                        t = tags[ip, jp]
                        self.surf_n[t] = self.surf_n.get(t, 0.0) + self._normal(iq, side, ip, jp)
        # a synthetic (theta index, radial index) pair per element, for the callers that stride over the mesh
        rc = np.hypot(self.nodes[:, 0].mean(axis=1), self.nodes[:, 1].mean(axis=1))
        tc = np.arctan2(self.nodes[:, 0].mean(axis=1), self.nodes[:, 1].mean(axis=1))
        self.nth, self.nr_ = 72, 28
        self.ab = np.stack([np.minimum((tc / np.pi * self.nth).astype(np.int64), self.nth - 1),
                            np.minimum((rc / self.r_outer * self.nr_).astype(np.int64), self.nr_ - 1)], 1)

    def _phi_pert(self, s, z, nr):
        if not self.model3d or self.perturb == 0.0:
            return np.zeros(nr)
        r = np.hypot(s, z)
        phi = 2 * np.pi * np.arange(nr) / nr
        amp = self.perturb * (s / self.r_outer) * np.sin(np.pi * r / self.r_outer)
        return amp * (np.cos(2 * phi + 0.3) + 0.5 * np.sin(3 * phi - z / self.r_outer) + 0.25 * np.cos(5 * phi))

    def _rho_vp(self, iq, ip, jp, nr):
        m, g = self.mat[iq], self.geo[iq]
        if self.vol[iq] is not None:           # mRhoMass3D / mVpFluid3D: the quad's samples resampled to the point's Nr
            from . import volumetric as VOL
            v = self.vol[iq]
            return VOL.linear_resampling(nr, v["rho"][:, ip * 5 + jp]), VOL.linear_resampling(nr, v["vpv"][:, ip * 5 + jp])
        p = self._phi_pert(g["s"][ip, jp], g["z"][ip, jp], nr)
        if self.is_fluid[iq] and not self.fluid3d:
            p = np.zeros(nr)
        return m["rho"][ip, jp] * (1.0 + 0.0 * p), m["vpv"][ip, jp] * (1.0 + 0.5 * p)

    def _normal(self, iq, side, ip, jp):
        """Quad::computeNormal (Quad.cpp:661-715), no relabelling."""
        nodes, g = self.nodes[iq], self.geo[iq]
        a, b = nodes[:, side], nodes[:, _p4(side + 1)]
        r0, r1 = np.hypot(*a), np.hypot(*b)
        th = lambda v: 0.0 if np.hypot(*v) < 1e-300 else np.arccos(np.clip(v[1] / np.hypot(*v), -1, 1))
        rsf = 0.5 * (r0 + r1)
        dth = abs(th(b) - th(a))
        half_r_dth, half_r2_dth = 0.5 * rsf * dth, 0.5 * rsf * rsf * dth
        sz = self._map(iq, g["xi2"][ip, jp], g["eta2"][ip, jp])
        rr = np.hypot(sz[0], sz[1])
        sint, cost = sz[0] / rr, sz[1] / rr
        n = np.array([sint, 0.0, cost])
        if self.relab[iq] is not None:
            # nRTZ . Q^T with the tilted normal of the undulated boundary (Relabelling::getSFNormalRTZ), one row per azimuth
            nrtz = self.relab[iq].sf_normal_rtz(ip * 5 + jp)
            n = np.stack([nrtz[:, 0] * cost + nrtz[:, 2] * sint, nrtz[:, 1], -nrtz[:, 0] * sint + nrtz[:, 2] * cost], 1)
        wxi = SP.W_GLJ if self.axial[iq] else SP.W_GLL
        wsf = wxi[ip] if side in (0, 2) else SP.W_GLL[jp]
        if self.axial[iq]:
            if ip == 0:
                n = n * wsf * g["J00"][ip, jp] * half_r_dth
            else:
                n = n * wsf / (1.0 + g["xi"][ip]) * sint * half_r2_dth
        else:
            n = n * wsf * sint * half_r2_dth
        if side != self.curved_outer[iq]:
            n = -n
        return n

    # ------------------------------------------------------------------ Mesh::getDeltaT
    def estimate_dt(self, factor=1.0):
        dt = np.inf
        for iq in range(self.nelem):
            nodes, g, m = self.nodes[iq], self.geo[iq], self.mat[iq]
            hmin_ref = min(np.hypot(*(nodes[:, i] - nodes[:, (i + 1) % 4])) for i in range(4))
            courant = self.elem_dt[iq] * m["vmax_ref"] / hmin_ref              # Quad::getCourant
            pts = np.stack([g["s"].ravel(), g["z"].ravel()], 1)
            d = np.linalg.norm(pts[:, None, :] - pts[None, :, :], axis=2) + np.eye(25) * 1e300
            vmax = max(m["vpv"].max(), m["vph"].max()) * (1.0 + (0.5 * self.perturb * 1.75 if self.model3d else 0.0))
            hmin = d.min()
            if self.vol[iq] is not None:       # Material::getVMax on the 3-D samples
                vmax = max(self.vol[iq]["vpv"].max(), self.vol[iq]["vph"].max())
            if self.relab[iq] is not None:     # Quad::getDeltaT: min over the azimuthal slices of courant * hmin / vmax
                hs = self.relab[iq].hmin_slices()
                if self.vol[iq] is not None:
                    vs = np.maximum(self.vol[iq]["vpv"].max(axis=1), self.vol[iq]["vph"].max(axis=1))
                    dt = min(dt, float((courant * hs / vs).min()))
                    continue
                hmin = hs.min()
            dt = min(dt, courant * hmin / vmax)
        return factor * dt

    # ------------------------------------------------------------------ release
    _att_factors = None

    def _att(self, dt, Qkp, Qmu):
        w, y = self.sls_w, self.sls_y
        ysum = y.sum()
        yd = y / ysum
        w0 = self.f_ref * 2 * np.pi
        w1 = np.sqrt(self.f_min * self.f_max) * 2 * np.pi
        fact = np.sum(yd * w * w / (w1 * w1 + w * w))
        alpha = np.exp(-w * dt)
        beta = ((1 - alpha) / (w * dt) - alpha) * yd
        gamma = ((alpha - 1) / (w * dt) + 1) * yd
        if self.do_kappa:
            kpNo = 1 + 2 * np.log(w1 / w0) / np.pi / Qkp
            dKp = kpNo / (Qkp / ysum + (1 - fact))
            kpAtt = kpNo + dKp * fact
        else:
            dKp, kpAtt, kpNo = np.zeros_like(Qkp), np.ones_like(Qkp), np.ones_like(Qkp)
        muNo = 1 + 2 * np.log(w1 / w0) / np.pi / Qmu
        dMu = muNo / (Qmu / ysum + (1 - fact))
        muAtt = muNo + dMu * fact
        return alpha, beta, gamma, dKp, kpAtt, kpNo, dMu, muAtt, muNo

    @staticmethod
    def _cg4_weights(f):
        w = np.zeros(4)
        w[0] = (f[0, 0] + f[0, 1] + f[1, 0] + f[1, 1] + 0.5 * (f[0, 2] + f[1, 2] + f[2, 0] + f[2, 1]) + 0.25 * f[2, 2]) / f[1, 1]
        w[1] = (f[0, 3] + f[0, 4] + f[1, 3] + f[1, 4] + 0.5 * (f[0, 2] + f[1, 2] + f[2, 3] + f[2, 4]) + 0.25 * f[2, 2]) / f[1, 3]
        w[2] = (f[3, 0] + f[3, 1] + f[4, 0] + f[4, 1] + 0.5 * (f[2, 0] + f[2, 1] + f[3, 2] + f[4, 2]) + 0.25 * f[2, 2]) / f[3, 1]
        w[3] = (f[3, 3] + f[3, 4] + f[4, 3] + f[4, 4] + 0.5 * (f[2, 3] + f[2, 4] + f[3, 2] + f[4, 2]) + 0.25 * f[2, 2]) / f[3, 3]
        return w

    def _make_point(self, t, local_mask=None):
        """GLLPoint::release (GLLPoint.cpp:48-128)."""
        nr = int(self.p_nr[t])
        crds = np.array([self.p_s[t], self.p_z[t]])
        axial = bool(self.p_axis[t])
        is_s, is_f = self.mass_s[t].any(), self.mass_f[t].any()

        def equal_rows(a):                         # XMath::equalRows: every row within 1e-10 (relative, 2-norm) of the first
            a = np.asarray(a, dtype=np.float64).reshape(len(a), -1)
            return bool((np.linalg.norm(a - a[0], axis=1) <= 1e-10 * np.linalg.norm(a[0])).all())

        def mk_mass(m):
            if equal_rows(m):
                return M.Mass1D(np.float32(1.0 / m[0]))
            return M.Mass3D((1.0 / m).astype(np.float32))
        sp = None
        if is_s and self.ocean_depth > 0.0 and int(t) in self.surf_n:
            # GLLPoint::release, ocean branch (GLLPoint.cpp:57-72): MassOcean1D where mass, depth and surface normal are the same
            # on every azimuthal sample, MassOcean3D (per-sample ocean mass and unit normal) otherwise
            ms = self.mass_s[t]
            sn = np.asarray(self.surf_n[int(t)], dtype=np.float64)
            if equal_rows(ms) and (sn.ndim == 1 or equal_rows(sn)):
                r = np.hypot(crds[0], crds[1])
                theta = 0.0 if r < 1e-10 else float(np.arccos(crds[1] / r))
                area = float(np.linalg.norm(sn if sn.ndim == 1 else sn[0]))
                sp = M.SolidPoint(nr, axial, crds, M.MassOcean1D(ms[0], 1027.0 * self.ocean_depth * area, theta))
            else:
                sn = sn if sn.ndim == 2 else np.repeat(sn[None, :], nr, axis=0)
                area = np.linalg.norm(sn, axis=1)
                sp = M.SolidPoint(nr, axial, crds, M.MassOcean3D(ms, 1027.0 * self.ocean_depth * area, sn / area[:, None]))
        elif is_s:
            sp = M.SolidPoint(nr, axial, crds, mk_mass(self.mass_s[t]))
        fp = M.FluidPoint(nr, axial, crds, mk_mass(self.mass_f[t]), bool(self.p_surface[t])) if is_f else None
        if sp is not None and fp is not None:
            n = self.sf_n[t]
            n_un = n
            if local_mask is not None:
                n_un = np.zeros_like(n)
                for e, c in self.sf_contrib[int(t)]:
                    if local_mask[e]:
                        n_un = n_un + c
            mf = self.mass_f[t]
            if equal_rows(mf) and equal_rows(n):       # GLLPoint::release: equalRows(mSFNormal_assmble) && equalRows(mMassFluid)
                c = M.SFCoupling1D(np.float32(n_un[0, 0]), np.float32(n_un[0, 2]), np.float32(n[0, 0] / mf[0]), np.float32(n[0, 2] / mf[0]))
            else:
                c = M.SFCoupling3D(n_un.astype(np.float32), (n / mf[:, None]).astype(np.float32))
            return M.SolidFluidPoint(sp, fp, c)
        return sp if sp is not None else fp

    def _make_element(self, iq, points, dt):
        g, m, f = self.geo[iq], self.mat[iq], self.ifact[iq]
        det = g["det"]
        inv_s = np.where(g["s"] > 0, 1.0 / np.where(g["s"] > 0, g["s"], 1.0), 0.0)
        if self.axial[iq]:
            inv_s[0, :] = 0.0
        grad = M.Gradient(g["J00"] / det, -g["J01"] / det, -g["J10"] / det, g["J11"] / det, inv_s, bool(self.axial[iq]))
        nr = int(self.e_nr[iq])
        ff = f.reshape(1, 25)
        vol = self.vol[iq]
        relab = self.relab[iq]
        relab3d = relab is not None and not relab.is_par1d()
        if vol is not None:
            # 3-D samples from the volumetric models; Quad::releaseSolid / releaseFluid: 1-D element if every property has equal rows
            from . import volumetric as VOL
            rows = nr
            with_att = self.att_kind is not None and self.has_att
            if self.is_fluid[iq]:
                is3d = not (VOL.equal_rows(vol["vpv"]) and VOL.equal_rows(vol["rho"]))                       # Material::isFluidPar1D
            else:
                is3d = not (all(VOL.equal_rows(vol[k]) for k in ("vpv", "vph", "vsv", "vsh", "rho", "eta")) and
                            (not with_att or (VOL.equal_rows(vol["qkp"]) and VOL.equal_rows(vol["qmu"]))))   # Material::isSolidPar1D
            cast = lambda x: np.ascontiguousarray(x if is3d else x[0:1]).astype(self.dtype_coef)
            rho, vpv, vph, vsv, vsh, eta = (vol[k].copy() for k in ("rho", "vpv", "vph", "vsv", "vsh", "eta"))
            qkp3, qmu3 = vol["qkp"], vol["qmu"]
            nrm = np.linalg.norm
            iso = nrm(vpv - vph) < 1e-10 * nrm(vpv) and nrm(vsv - vsh) < 1e-10 * nrm(vsv) and nrm(eta - 1.0) < 1e-10        # Material::isIsotropic
        else:
            rows = nr if ((self.model3d and (not self.is_fluid[iq] or self.fluid3d)) or relab3d) else 1
            pert = np.zeros((rows, 25))
            if rows > 1:
                for ip in range(5):
                    for jp in range(5):
                        pert[:, ip * 5 + jp] = self._phi_pert(g["s"][ip, jp], g["z"][ip, jp], nr)
            flat = lambda a: a.reshape(1, 25) * np.ones((rows, 1))
            is3d = bool(rows > 1 and (np.ptp(pert, axis=0).any() or relab3d))
            cast = lambda x: np.ascontiguousarray(x if is3d else x[0:1]).astype(self.dtype_coef)
            rho = flat(m["rho"])
            if not self.is_fluid[iq]:
                vpv, vph = flat(m["vpv"]) * (1 + 0.5 * pert), flat(m["vph"]) * (1 + 0.5 * pert)
                vsv, vsh = flat(m["vsv"]) * (1 + pert), flat(m["vsh"]) * (1 + pert)
                eta = flat(m["eta"])
                if self.has_att:
                    qkp3, qmu3 = flat(m["qkp"]), flat(m["qmu"])
            iso = np.allclose(m["vpv"], m["vph"]) and np.allclose(m["vsv"], m["vsh"]) and np.allclose(m["eta"], 1.0)   # Material::isIsotropic
        # Quad::releaseSolid / releaseFluid: elem1D = material 1-D and relabelling 1-D; PRT from Relabelling::createPRT
        prt, Jst = None, None
        if relab is not None:
            if vol is not None and relab3d:
                is3d = True
            Jst = relab.stiff_jacobian()                       # [Nr][25]
            X = relab.stiff_x()                                # [4][Nr][25]
            if is3d:
                prt = M.PRT_3D(np.concatenate([X[k] for k in range(4)], axis=1))
            else:
                prt = M.PRT_1D([X[k][0].reshape(5, 5) for k in range(4)])
            if rows == 1:
                Jst = Jst[0:1]
            cast = lambda x: np.ascontiguousarray(x if is3d else x[0:1]).astype(self.dtype_coef)
        if self.is_fluid[iq]:
            K = ff / rho
            if Jst is not None:
                K = K * Jst
            ac = M.Acoustic3D(cast(K * np.ones((rows, 1)))) if is3d else M.Acoustic1D(K[0].reshape(5, 5))
            return M.FluidElement(grad, prt, points, ac)
        A, C, L, N = rho * vph ** 2 * ff, rho * vpv ** 2 * ff, rho * vsv ** 2 * ff, rho * vsh ** 2 * ff
        F = eta * (A - 2 * L)
        if Jst is not None:                                    # must do relabelling before attenuation (Material.cpp:322-330)
            A, C, F, L, N = A * Jst, C * Jst, F * Jst, L * Jst, N * Jst
        att = None
        if self.att_kind is not None and self.has_att:
            kp = (4 * A + C + 4 * F - 4 * N) / 9.0                # Voigt average (Material.cpp:330-333)
            mu = (A + C - 2 * F + 6 * L + 5 * N) / 15.0
            A, C, F, L, N = A - (kp + 4 / 3 * mu), C - (kp + 4 / 3 * mu), F - (kp - 2 / 3 * mu), L - mu, N - mu
            al, be, ga, dKp, kpAtt, kpNo, dMu, muAtt, muNo = self._att(dt, qkp3, qmu3)
            nsls = len(al)
            if self.att_kind == "cg4":
                wc = self._cg4_weights(f)
                sel = [6, 8, 16, 18]
                dkp = np.stack([wc[i] * dKp[:, k] * kp[:, k] for i, k in enumerate(sel)], 1)
                dmu = np.stack([wc[i] * dMu[:, k] * mu[:, k] for i, k in enumerate(sel)], 1)
                kp, mu = kp * kpNo, mu * muNo
                for i, k in enumerate(sel):
                    kp[:, k] *= 1 + wc[i] * (kpAtt[:, k] / kpNo[:, k] - 1)
                    mu[:, k] *= 1 + wc[i] * (muAtt[:, k] / muNo[:, k] - 1)
                att = M.Attenuation3D_CG4(nsls, al, be, ga, dkp, dmu, self.do_kappa) if is3d else \
                    M.Attenuation1D_CG4(nsls, al, be, ga, nr // 2, dkp[0], dmu[0], self.do_kappa)
            else:
                dkp, dmu = dKp * kp, dMu * mu
                kp, mu = kp * kpAtt, mu * muAtt
                att = M.Attenuation3D_Full(nsls, al, be, ga, dkp, dmu, self.do_kappa) if is3d else \
                    M.Attenuation1D_Full(nsls, al, be, ga, nr // 2, dkp[0].reshape(5, 5), dmu[0].reshape(5, 5), self.do_kappa)
            A, C, F, L, N = A + (kp + 4 / 3 * mu), C + (kp + 4 / 3 * mu), F + (kp - 2 / 3 * mu), L + mu, N + mu
        if iso:
            el = (M.Isotropic3D if is3d else M.Isotropic1D)(cast(F), cast(L), att)
        else:
            el = (M.TransverselyIsotropic3D if is3d else M.TransverselyIsotropic1D)(cast(A), cast(C), cast(F), cast(L), cast(N), att)
        return M.SolidElement(grad, prt, points, el)

    def release(self, domain, dt, rank=0, elem_to_proc=None):
        """Mesh::release (Mesh.cpp:177-208)."""
        domain.setGMat(SP.G_GLL, SP.G_GLJ)
        if elem_to_proc is None:
            elem_to_proc = np.zeros(self.nelem, dtype=np.int64)
        dec = CN.decompose(self.conn, elem_to_proc, rank, self.e2g, self.neighbours)
        l2g = dec.local_to_global_gll
        local_mask = np.asarray(elem_to_proc) == rank
        pts = [self._make_point(int(t), None if local_mask.all() else local_mask) for t in l2g]
        for p in pts:
            domain.addPoint(p)
        elems = []
        for il, e in enumerate(dec.local_elems):
            tags = dec.elemToGllLocal[il].reshape(-1)
            el = self._make_element(int(e), [pts[t] for t in tags], dt)
            domain.addElement(el)
            elems.append(el)
        info = M.MessagingInfo(dec.iProcComm, dec.iLocalPoints)
        return dict(points=pts, elements=elems, msg=info, dec=dec)

    def source_element(self):
        """an axial solid element just below the surface on the northern axis (CMTSOLUTION depth ~ 12 km)"""
        cand = [iq for iq in range(self.nelem) if self.axial[iq] and not self.is_fluid[iq] and self.nodes[iq, 1].mean() > 0]
        return max(cand, key=lambda iq: self.nodes[iq, 1].mean() - 1e9 * (self.surf_side[iq] >= 0))

    def make_source(self, elements, dec=None, amp=1e20):
        e_glob = self.source_element()
        if dec is not None:
            loc = np.nonzero(dec.local_elems == e_glob)[0]
            if len(loc) == 0:
                return None
            el = elements[int(loc[0])]
        else:
            el = elements[e_glob]
        rs = np.random.default_rng(7)
        force = []
        for i in range(25):
            fc = (rs.standard_normal((3, 3)) + 1j * rs.standard_normal((3, 3))) * amp
            fc[0] = fc[0].real
            force.append(fc)
        return M.SourceTerm(el, force)

    def work_per_step(self):
        return int(np.sum(self.p_nr // 2 + 1))
