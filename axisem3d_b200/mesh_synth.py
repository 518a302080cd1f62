"""Synthetic meridional (s, z) spectral-element mesh + PREM-like model -> solver descriptors.

BASELINE.json's configs run on the shipped Exodus mesh, whose reader (NetCDF-4/HDF5) and the
whole preloop are outside the hot-path scope.  This generator produces the
same *kind* of data the reference's `Mesh::release` hands to `Domain` (Mesh.cpp:177-208):
  - a structured polar grid (n_theta x n_r quads, hollow centre) with spherical mapping,
    axial elements on both poles oriented so that side 3 (ipol = 0) lies on the axis
    (ExodusModel.cpp:378-403), GLJ nodes in xi for axial elements (Gradient.cpp:15-23);
  - geometry matrices as in Quad::createGraident (Quad.cpp:509-525), integral factor as in
    Quad::formIntegralFactor (Quad.cpp:527-547);
  - per-point Nr: NrField value capped by the circumference, odd on / next to the axis,
    lucky-rounded (Quad.cpp:549-616; PreloopFFTW.cpp:59-109);
  - masses / solid-fluid normals accumulated per GLL point as in Quad::setupGLLPoints
    (Quad.cpp:325-372), Quad::computeNormal (Quad.cpp:661-715), GLLPoint::release
    (GLLPoint.cpp:48-128);
  - moduli scaled by the integral factor (Material.cpp:257-330), SLS attenuation factors in the
    spirit of AttAxiSEM.cpp:14-57 / AttBuilder.cpp:17-139 (simplified, documented below).
It is an input generator ("data": "synthetic"), not part of the hot path.
"""
from __future__ import annotations

import numpy as np

from . import model as M
from . import spectral as SP
from . import connectivity as CN

R_EARTH = 6371e3


def prem_like(r):
    """Smooth PREM-like profiles (rho [kg/m3], vp, vs [m/s]) of radius r [m]; fluid between
    ICB (1221.5 km) and CMB (3480 km)."""
    x = r / R_EARTH
    rho = 13.0e3 - 8.5e3 * x ** 2 - 1.2e3 * x
    vp = 11.2e3 - 4.0e3 * x ** 2 + 1.5e3 * x * (1 - x)
    vs = 3.6e3 + 3.2e3 * (1 - x) * x * 2.0
    return rho, vp, vs


class SynthMesh:
    def __init__(self, n_theta=12, n_r=8, r_in=600e3, r_out=R_EARTH, fluid_layers=None,
                 nu=2, lucky=True, law="iso", model3d=False, perturb=0.02, perturb_rho=False,
                 fluid3d=False, attenuation=None, seed=20260101, nu_fn=None, dtype_coef=np.float64, prt=False, ocean=False):
        """law: 'iso' | 'ti' | 'aniso'.  model3d: phi-dependent material in solid elements.
        attenuation: None | 'cg4' | 'full'.  nu: constant Fourier order, or nu_fn(s, z) -> nu.
        fluid_layers: (b0, b1) radial element layers [b0, b1) that are fluid; default = the layers
        whose centre lies in the outer-core range; pass () for an all-solid mesh.
        prt: every element carries a particle-relabelling transform (PRT_1D with 1D material, PRT_3D with 3D material) with
        seeded X matrices close to the identity (X0, X3 ~ 1, X1, X2 ~ 0) -- arithmetic coverage, not a physical undulation."""
        self.prt = bool(prt)
        # ocean: solid points on the outer surface carry an ocean-load mass (MassOcean1D where the point's mass is
        # axisymmetric, MassOcean3D otherwise; GLLPoint.cpp:57-72) with a water column of ~3 km (+-20 % in phi for 3D)
        self.ocean = bool(ocean)
        self.nth, self.nr_ = int(n_theta), int(n_r)
        self.r_in, self.r_out = float(r_in), float(r_out)
        self.law, self.model3d, self.perturb = law, bool(model3d), float(perturb)
        self.perturb_rho, self.fluid3d = bool(perturb_rho), bool(fluid3d)
        self.att_kind = attenuation
        self.nu, self.nu_fn, self.lucky = nu, nu_fn, bool(lucky)
        self.rng = np.random.default_rng(seed)
        self.dtype_coef = dtype_coef
        self.th_edges = np.linspace(0.0, np.pi, self.nth + 1)
        self.r_edges = np.linspace(self.r_in, self.r_out, self.nr_ + 1)
        self.dth = np.pi / self.nth
        self.dr = (self.r_out - self.r_in) / self.nr_
        rc = 0.5 * (self.r_edges[:-1] + self.r_edges[1:])
        if fluid_layers is None:
            fl = (rc > 1221.5e3) & (rc < 3480e3)
        else:
            fl = np.zeros(self.nr_, dtype=bool)
            if len(fluid_layers) == 2:
                fl[fluid_layers[0]:fluid_layers[1]] = True
        self.layer_fluid = fl
        # SLS parameters (mesh globals of the shipped 50 s mesh)
        self.sls_w = np.array([0.0319, 0.5441, 5.308])
        self.sls_y = np.array([1.676, 1.504, 2.302])
        self.f_min, self.f_max, self.f_ref = 0.001, 1.0, 1.0
        self._build_topology()
        self._build_points()

    # ------------------------------------------------------------------ topology
    def _build_topology(self):
        nth, nr = self.nth, self.nr_
        nid = lambda a, b: a * (nr + 1) + b
        conn, flip, ab = [], [], []
        for a in range(nth):
            for b in range(nr):
                f = (a == nth - 1) and nth > 1
                if not f:
                    conn.append([nid(a, b), nid(a + 1, b), nid(a + 1, b + 1), nid(a, b + 1)])
                else:
                    conn.append([nid(a + 1, b + 1), nid(a, b + 1), nid(a, b), nid(a + 1, b)])
                flip.append(f)
                ab.append((a, b))
        self.conn = np.array(conn, dtype=np.int64)
        self.flip = np.array(flip, dtype=bool)
        self.ab = np.array(ab, dtype=np.int64)
        self.nelem = len(conn)
        self.axial = (self.ab[:, 0] == 0) | (self.ab[:, 0] == nth - 1)
        self.is_fluid = self.layer_fluid[self.ab[:, 1]]
        self.neighbours = CN.form_neighbourhood(self.conn)
        self.ngll, self.e2g = CN.form_elem_to_gll(self.conn, self.neighbours)

    def _elem_frame(self, e):
        a, b = self.ab[e]
        if not self.flip[e]:
            return self.th_edges[a], self.th_edges[a + 1], self.r_edges[b], self.r_edges[b + 1]
        return self.th_edges[a + 1], self.th_edges[a], self.r_edges[b + 1], self.r_edges[b]

    def _elem_geometry(self, e):
        """theta, r, s, z, J (2x2) and detJ on the 5x5 points of element e."""
        t0, t1, r0, r1 = self._elem_frame(e)
        xi = SP.P_GLJ if self.axial[e] else SP.P_GLL
        eta = SP.P_GLL
        th = (t0 + 0.5 * (xi + 1.0) * (t1 - t0))[:, None] * np.ones((1, 5))
        r = np.ones((5, 1)) * (r0 + 0.5 * (eta + 1.0) * (r1 - r0))[None, :]
        dth, dr = t1 - t0, r1 - r0
        s, z = r * np.sin(th), r * np.cos(th)
        if self.axial[e]:
            s[0, :] = 0.0
        J00, J01 = r * np.cos(th) * dth / 2, np.sin(th) * dr / 2
        J10, J11 = -r * np.sin(th) * dth / 2, np.cos(th) * dr / 2
        det = J00 * J11 - J01 * J10
        return dict(th=th, r=r, s=s, z=z, J00=J00, J01=J01, J10=J10, J11=J11, det=det,
                    xi=xi, eta=eta, dth=dth, dr=dr)

    # ---------------------------------------------------------------- GLL points
    def _nr_at(self, s, z, odd):
        nu = self.nu_fn(s, z) if self.nu_fn is not None else self.nu
        nr = 2 * int(nu) + 1
        r = np.hypot(s, z)
        spacing = 0.5 * (r * self.dth + self.dr) / 4.0
        upper = max(int(2 * np.pi * s / spacing), 3)
        nr = min(nr, upper)
        force_odd = False
        if nr % 2 == 0 and odd:
            nr += 1
            force_odd = True
        if self.lucky:
            nr = SP.next_lucky_number(nr, force_odd)
        return nr

    def _build_points(self):
        ng = self.ngll
        self.geo = [self._elem_geometry(e) for e in range(self.nelem)]
        self.p_s = np.zeros(ng)
        self.p_z = np.zeros(ng)
        self.p_axis = np.zeros(ng, dtype=bool)
        in_axial_elem = np.zeros(ng, dtype=bool)
        for e in range(self.nelem):
            g = self.geo[e]
            tags = self.e2g[e]
            self.p_s[tags] = g["s"]
            self.p_z[tags] = g["z"]
            if self.axial[e]:
                in_axial_elem[tags] = True
                self.p_axis[tags[0, :]] = True
        self.p_nr = np.array([self._nr_at(self.p_s[t], self.p_z[t], in_axial_elem[t]) for t in range(ng)],
                             dtype=np.int64)
        self.e_nr = np.array([self.p_nr[self.e2g[e]].max() for e in range(self.nelem)], dtype=np.int64)
        # integral factor (Quad.cpp:527-547)
        self.ifact = []
        for e in range(self.nelem):
            g = self.geo[e]
            wxi = SP.W_GLJ if self.axial[e] else SP.W_GLL
            w = wxi[:, None] * SP.W_GLL[None, :]
            if self.axial[e]:
                f = w * g["s"] / (1.0 + g["xi"])[:, None].clip(1e-300) * g["det"]
                f[0, :] = (w * g["J00"] * g["det"])[0, :]
            else:
                f = w * g["s"] * g["det"]
            self.ifact.append(f)
        # accumulate masses and solid-fluid normals per global GLL point
        self.mass_s = [np.zeros(n) for n in self.p_nr]
        self.mass_f = [np.zeros(n) for n in self.p_nr]
        self.sf_n = [None] * ng
        self.sf_contrib = {}          # GLL tag -> [(element, contribution)] for the rank-local (unassembled) normal
        for e in range(self.nelem):
            g = self.geo[e]
            tags = self.e2g[e]
            for ip in range(5):
                for jp in range(5):
                    t = tags[ip, jp]
                    rho, vp, vs = self._props(g["s"][ip, jp], g["z"][ip, jp], self.p_nr[t], self.is_fluid[e])
                    if self.is_fluid[e]:
                        self.mass_f[t] += self.ifact[e][ip, jp] / (rho * vp ** 2)
                    else:
                        self.mass_s[t] += self.ifact[e][ip, jp] * rho
            for side in self._sf_sides(e):
                for (ip, jp) in CN.EDGE_IJ[side]:
                    t = tags[ip, jp]
                    n = self._normal(e, side, ip, jp)
                    if not self.is_fluid[e]:
                        n = -n
                    if self.sf_n[t] is None:
                        self.sf_n[t] = np.zeros((self.p_nr[t], 3))
                    self.sf_n[t] += 0.5 * n[None, :]
                    self.sf_contrib.setdefault(int(t), []).append((e, 0.5 * n))

    def _sf_sides(self, e):
        a, b = self.ab[e]
        out = []
        outer, inner = (0, 2) if self.flip[e] else (2, 0)
        if b + 1 < self.nr_ and self.layer_fluid[b + 1] != self.layer_fluid[b]:
            out.append(outer)
        if b - 1 >= 0 and self.layer_fluid[b - 1] != self.layer_fluid[b]:
            out.append(inner)
        return out

    def _normal(self, e, side, ip, jp):
        """Quad::computeNormal (Quad.cpp:661-715) without relabelling: unit radial direction times
        the area element of the spherical side; sign + on the curved-outer side."""
        g = self.geo[e]
        t0, t1, r0, r1 = self._elem_frame(e)
        outer = 0 if self.flip[e] else 2
        rsf = max(r0, r1) if side == outer else min(r0, r1)
        half_r_dth = 0.5 * rsf * abs(t1 - t0)
        half_r2_dth = 0.5 * rsf * rsf * abs(t1 - t0)
        s, z = g["s"][ip, jp], g["z"][ip, jp]
        rr = np.hypot(s, z)
        sint, cost = s / rr, z / rr
        n = np.array([sint, 0.0, cost])
        wsf = (SP.W_GLJ if self.axial[e] else SP.W_GLL)[ip]
        if self.axial[e]:
            if ip == 0:
                n = n * wsf * g["J00"][ip, jp] * half_r_dth
            else:
                n = n * wsf / (1.0 + g["xi"][ip]) * sint * half_r2_dth
        else:
            n = n * wsf * sint * half_r2_dth
        if side != outer:
            n = -n
        return n

    # ------------------------------------------------------------------ material
    def _phi_pert(self, s, z, nr):
        """Relative perturbation on nr azimuthal samples phi_j = 2 pi j / nr (Quad.cpp:475-477);
        vanishes on the axis and outside the solid mantle-like region."""
        if not self.model3d or self.perturb == 0.0:
            return np.zeros(nr)
        r = np.hypot(s, z)
        phi = 2 * np.pi * np.arange(nr) / nr
        amp = self.perturb * (s / self.r_out) * np.sin(np.pi * (r - self.r_in) / (self.r_out - self.r_in))
        return amp * (np.cos(2 * phi + 0.3) + 0.5 * np.sin(3 * phi - z / self.r_out) + 0.25 * np.cos(5 * phi))

    def _props(self, s, z, nr, fluid):
        r = np.hypot(s, z)
        rho, vp, vs = prem_like(r)
        p = self._phi_pert(s, z, nr)
        if fluid and not self.fluid3d:
            p = np.zeros(nr)
        rho_a = rho * (1.0 + (0.4 * p if self.perturb_rho else 0.0 * p))
        vp_a = vp * (1.0 + 0.5 * p)
        vs_a = (0.0 if fluid else vs) * (1.0 + p)
        return rho_a, vp_a, vs_a

    def _att_factors(self, dt, Qkp, Qmu):
        """AttAxiSEM-style SLS factors (AttAxiSEM.cpp:24-57)."""
        w, y = self.sls_w, self.sls_y
        ysum = y.sum()
        yd = y / ysum
        w0 = self.f_ref * 2 * np.pi
        w1 = np.sqrt(self.f_min * self.f_max) * 2 * np.pi
        fact = np.sum(yd * w * w / (w1 * w1 + w * w))
        alpha = np.exp(-w * dt)
        beta = ((1 - alpha) / (w * dt) - alpha) * yd
        gamma = ((alpha - 1) / (w * dt) + 1) * yd
        kpNo = 1 + 2 * np.log(w1 / w0) / np.pi / Qkp
        dKp = kpNo / (Qkp / ysum + (1 - fact))
        kpAtt = kpNo + dKp * fact
        muNo = 1 + 2 * np.log(w1 / w0) / np.pi / Qmu
        dMu = muNo / (Qmu / ysum + (1 - fact))
        muAtt = muNo + dMu * fact
        return alpha, beta, gamma, dKp, kpAtt, kpNo, dMu, muAtt, muNo

    def _cg4_weights(self, f):
        """Quad::computeWeightsCG4 (Quad.cpp:438-466) on the 5x5 integral factor."""
        w = np.zeros(4)
        w[0] = (f[0, 0] + f[0, 1] + f[1, 0] + f[1, 1] + 0.5 * (f[0, 2] + f[1, 2] + f[2, 0] + f[2, 1]) + 0.25 * f[2, 2]) / f[1, 1]
        w[1] = (f[0, 3] + f[0, 4] + f[1, 3] + f[1, 4] + 0.5 * (f[0, 2] + f[1, 2] + f[2, 3] + f[2, 4]) + 0.25 * f[2, 2]) / f[1, 3]
        w[2] = (f[3, 0] + f[3, 1] + f[4, 0] + f[4, 1] + 0.5 * (f[2, 0] + f[2, 1] + f[3, 2] + f[4, 2]) + 0.25 * f[2, 2]) / f[3, 1]
        w[3] = (f[3, 3] + f[3, 4] + f[4, 3] + f[4, 4] + 0.5 * (f[2, 3] + f[2, 4] + f[3, 2] + f[4, 2]) + 0.25 * f[2, 2]) / f[3, 3]
        return w

    def estimate_dt(self, courant=0.4):
        hmin = np.inf
        for e in range(self.nelem):
            g = self.geo[e]
            pts = np.stack([g["s"].ravel(), g["z"].ravel()], 1)
            d = np.linalg.norm(pts[:, None, :] - pts[None, :, :], axis=2) + np.eye(25) * 1e30
            _, vp, _ = prem_like(np.hypot(g["s"], g["z"]))
            hmin = min(hmin, (d.min(axis=1) / (vp.ravel() * (1 + self.perturb))).min())
        # the azimuthal spacing 2 pi s / Nr is bounded below by the cap in _nr_at
        return courant * hmin * 0.5

    # ---------------------------------------------------------------- descriptors
    def _make_point(self, t, local_mask=None):
        """GLLPoint::release (GLLPoint.cpp:48-128).  Masses and the assembled SF normal are global sums (the
        reference assembles them over MPI at setup, Mesh.cpp:339-389); the *unassembled* normal handed to
        SFCoupling is the sum over this rank's elements only (GLLPoint.cpp:99-117), so that the halo sum of the
        fluid stiffness adds up to the assembled coupling term."""
        nr = int(self.p_nr[t])
        crds = np.array([self.p_s[t], self.p_z[t]])
        axial = bool(self.p_axis[t])
        is_s = self.mass_s[t].any()
        is_f = self.mass_f[t].any()

        def mk_mass(m):
            if np.ptp(m) <= 1e-12 * np.abs(m).max():     # XMath::equalRows
                return M.Mass1D(np.float32(1.0 / m[0]))
            return M.Mass3D((1.0 / m).astype(np.float32))
        ms = mk_mass(self.mass_s[t]) if is_s else None
        if is_s and self.ocean and not is_f and np.hypot(crds[0], crds[1]) > self.r_out * (1.0 - 1e-9):
            m = self.mass_s[t]
            theta = float(np.arccos(np.clip(crds[1] / np.hypot(crds[0], crds[1]), -1.0, 1.0)))
            if not ms.is3D and not self.perturb_rho:      # a phi-dependent water column makes the load 3D (GLLPoint.cpp:63-71)
                ms = M.MassOcean1D(m[0], 0.35 * m[0], theta)
            else:
                phi = 2.0 * np.pi * np.arange(nr) / nr
                nrm = np.stack([np.sin(theta) * np.ones(nr), 0.05 * np.sin(phi), np.cos(theta) * np.ones(nr)], 1)
                nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
                ms = M.MassOcean3D(m, 0.35 * m * (1.0 + 0.2 * np.cos(phi + 0.3)), nrm)
        sp = M.SolidPoint(nr, axial, crds, ms) if is_s else None
        fp = M.FluidPoint(nr, axial, crds, mk_mass(self.mass_f[t]), False) if is_f else None
        if sp is not None and fp is not None:
            n = self.sf_n[t]
            n_un = n
            if local_mask is not None:
                n_un = np.zeros_like(n)
                for e, c in self.sf_contrib[int(t)]:
                    if local_mask[e]:
                        n_un = n_un + c[None, :]
            mf = self.mass_f[t]
            if np.ptp(mf) <= 1e-12 * np.abs(mf).max() and np.abs(n - n[0]).max() <= 1e-12 * np.abs(n).max():
                c = M.SFCoupling1D(np.float32(n_un[0, 0]), np.float32(n_un[0, 2]),
                                   np.float32(n[0, 0] / mf[0]), np.float32(n[0, 2] / mf[0]))
            else:
                c = M.SFCoupling3D(n_un.astype(np.float32), (n / mf[:, None]).astype(np.float32))
            return M.SolidFluidPoint(sp, fp, c)
        return sp if sp is not None else fp

    def _make_element(self, e, points, dt):
        g = self.geo[e]
        det = g["det"]
        inv_s = np.where(g["s"] > 0, 1.0 / np.where(g["s"] > 0, g["s"], 1.0), 0.0)
        if self.axial[e]:
            inv_s[0, :] = 0.0
        grad = M.Gradient(g["J00"] / det, -g["J01"] / det, -g["J10"] / det, g["J11"] / det,
                          inv_s, bool(self.axial[e]))
        nr = int(self.e_nr[e])
        f = self.ifact[e]
        rho = np.zeros((nr, 25)); vp = np.zeros((nr, 25)); vs = np.zeros((nr, 25))
        for ip in range(5):
            for jp in range(5):
                a, b, c = self._props(g["s"][ip, jp], g["z"][ip, jp], nr, self.is_fluid[e])
                k = ip * 5 + jp
                rho[:, k], vp[:, k], vs[:, k] = a, b, c
        ff = f.reshape(1, 25)
        is3d = bool(np.ptp(rho, axis=0).any() or np.ptp(vp, axis=0).any() or np.ptp(vs, axis=0).any())
        cast = lambda x: np.ascontiguousarray(x if is3d else x[0:1]).astype(self.dtype_coef)
        if self.is_fluid[e]:
            K = ff / rho
            ac = M.Acoustic3D(cast(K)) if is3d else M.Acoustic1D(K[0].reshape(5, 5))
            return M.FluidElement(grad, self._make_prt(e, nr, is3d), points, ac)
        mu = rho * vs ** 2 * ff
        kp = rho * vp ** 2 * ff - 4.0 / 3.0 * mu
        att = None
        if self.att_kind is not None:
            Qmu = np.full_like(mu, 300.0)
            Qkp = np.full_like(mu, 57823.0)
            al, be, ga, dKp, kpAtt, kpNo, dMu, muAtt, muNo = self._att_factors(dt, Qkp, Qmu)
            nsls = len(al)
            if self.att_kind == "cg4":
                wc = self._cg4_weights(f)
                sel = [6, 8, 16, 18]
                dkp = np.stack([wc[i] * dKp[:, k] * kp[:, k] for i, k in enumerate(sel)], 1)
                dmu = np.stack([wc[i] * dMu[:, k] * mu[:, k] for i, k in enumerate(sel)], 1)
                kp = kp * kpNo
                mu = mu * muNo
                for i, k in enumerate(sel):
                    kp[:, k] *= 1 + wc[i] * (kpAtt[:, k] / kpNo[:, k] - 1)
                    mu[:, k] *= 1 + wc[i] * (muAtt[:, k] / muNo[:, k] - 1)
                if is3d:
                    att = M.Attenuation3D_CG4(nsls, al, be, ga, dkp, dmu, True)
                else:
                    att = M.Attenuation1D_CG4(nsls, al, be, ga, nr // 2, dkp[0], dmu[0], True)
            else:
                dkp, dmu = dKp * kp, dMu * mu
                kp, mu = kp * kpAtt, mu * muAtt
                if is3d:
                    att = M.Attenuation3D_Full(nsls, al, be, ga, dkp, dmu, True)
                else:
                    att = M.Attenuation1D_Full(nsls, al, be, ga, nr // 2, dkp[0].reshape(5, 5), dmu[0].reshape(5, 5), True)
        lam = kp - 2.0 / 3.0 * mu
        if self.law == "iso":
            el = (M.Isotropic3D if is3d else M.Isotropic1D)(cast(lam), cast(mu), att)
        elif self.law == "ti":
            # 2 % radial anisotropy: N = 1.04 L, A = 1.03 C, eta = 0.95
            L_, C_ = mu, lam + 2 * mu
            N_, A_ = 1.04 * L_, 1.03 * C_
            F_ = 0.95 * (A_ - 2 * L_)
            el = (M.TransverselyIsotropic3D if is3d else M.TransverselyIsotropic1D)(
                cast(A_), cast(C_), cast(F_), cast(L_), cast(N_), att)
        else:
            C6 = np.zeros((6, 6) + lam.shape)
            for i in range(3):
                for j in range(3):
                    C6[i, j] = lam
                C6[i, i] = lam + 2 * mu
                C6[i + 3, i + 3] = mu
            # deterministic symmetric perturbation (a few % of mu) so that all 21 moduli are live
            rs = np.random.default_rng(1000 + e)
            P = rs.uniform(-0.03, 0.03, size=(6, 6))
            P = 0.5 * (P + P.T)
            for i in range(6):
                for j in range(6):
                    C6[i, j] = C6[i, j] + P[i, j] * mu
            el = (M.Anisotropic3D if is3d else M.Anisotropic1D)([cast(C6[i, j]) for (i, j) in M.ANISO_IJ], att)
        return M.SolidElement(grad, self._make_prt(e, nr, is3d), points, el)

    def _make_prt(self, e, nr, is3d):
        """PRT_1D(array<RMatPP, 4>) / PRT_3D(RMatXN4) in the space of the element's material (SolidElement.cpp:27-32)."""
        if not self.prt:
            return None
        rs = np.random.default_rng(5000 + e)
        base = rs.uniform(-0.05, 0.05, size=(4, 25))
        base[0] += 1.0
        base[3] += 1.0
        if not is3d:
            return M.PRT_1D(base.reshape(4, 5, 5))
        phi = 2.0 * np.pi * np.arange(nr) / nr
        X = np.zeros((nr, 100))
        for k in range(4):
            amp, ph = rs.uniform(0.0, 0.02, size=25), rs.uniform(0, 2 * np.pi, size=25)
            X[:, 25 * k:25 * (k + 1)] = base[k][None, :] + amp[None, :] * np.cos((k + 1) * phi[:, None] + ph[None, :])
        return M.PRT_3D(X)

    def release(self, domain, dt, rank=0, elem_to_proc=None):
        """Mesh::release (Mesh.cpp:177-208): points in local GLL order, then elements in global-id
        order, then messaging.  Returns dict with local tags for convenience."""
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

    def make_source(self, elements, dec=None, amp=1e20):
        """Moment-tensor-like force on one axial element (orders 0..2 like Earthquake.cpp:20-88);
        returns SourceTerm or None when the element is not local."""
        e_glob = (self.nr_ * 0) + (self.nr_ - 2)          # axial column, second layer from the top
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
            f = (rs.standard_normal((3, 3)) + 1j * rs.standard_normal((3, 3))) * amp
            f[0] = f[0].real
            force.append(f)
        return M.SourceTerm(el, force)

    def work_per_step(self):
        """Metric numerator: sum over GLL points of (Nu_p + 1)."""
        return int(np.sum(self.p_nr // 2 + 1))
