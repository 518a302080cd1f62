"""Host-side mirror of the reference's solver-object constructors (descriptors only).

The reference hands data to its time loop by constructing C++ objects
(`Mesh::release`, /root/reference/SOLVER/src/preloop/mesh/Mesh.cpp:177-208;
`GLLPoint::release`, GLLPoint.cpp:48-128; `Quad::releaseSolid/Fluid`, Quad.cpp:386-420).
The classes below keep the same names and constructor arguments (Eigen matrices
become numpy arrays) and hold data only -- no arithmetic.  Two consumers walk them:
`axisem3d_b200.domain.Domain` (uploads through the C-ABI, CUDA path) and
`oracle/axisem_oracle.py` (CPU checker, used by tests only).

Array conventions (reference S/core/eigenc.h):
  RMatPP  -> (5, 5) array, index (ipol, jpol), row-major; flat point ipnt = ipol*5 + jpol
  RMatXN  -> (Nr, 25)   RMatX4 -> (Nr, 4)   RColX -> (Nr,)   RMatX3 -> (Nr, 3)
"""
from __future__ import annotations

import numpy as np

nPol = 4
nPntEdge = nPol + 1
nPntElem = nPntEdge * nPntEdge
nPE = nPntElem

# Voigt order of the 21 independent moduli in the Anisotropic{1D,3D} constructors
# (Anisotropic1D.h:14-20): C11 C12 C13 C14 C15 C16 C22 C23 ... C66
ANISO_IJ = [(i, j) for i in range(6) for j in range(i, 6)]


def _f(a, shape=None):
    a = np.asarray(a)
    a = np.ascontiguousarray(a if a.dtype == np.float32 else a.astype(np.float64))
    if shape is not None and a.shape != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (shape, a.shape))
    return a


# ----------------------------------------------------------------------------- mass
class Mass1D:
    """Mass1D(Real invMass) -- S/core/point/mass/Mass1D.h"""
    is3D = False

    def __init__(self, invMass):
        self.invMass = float(invMass)

    def checkCompatibility(self, nr):
        pass


class Mass3D:
    """Mass3D(const RColX &invMass) -- S/core/point/mass/Mass3D.cpp:9; rows must equal
    the point's Nr (Mass3D.cpp:59-64)."""
    is3D = True

    def __init__(self, invMass):
        self.invMass = _f(invMass).reshape(-1)

    def checkCompatibility(self, nr):
        if self.invMass.shape[0] != nr:
            raise RuntimeError("Mass3D::checkCompatibility || Incompatible size.")


class MassOcean1D:
    """MassOcean1D(double mass, double massOcean, double theta) -- MassOcean1D.cpp:8-13 (solid surface points under an
    axisymmetric ocean load, GLLPoint.cpp:57-62)."""
    is3D = False
    ocean = True

    def __init__(self, mass, massOcean, theta):
        self.mass, self.massOcean, self.theta = float(mass), float(massOcean), float(theta)
        self.invMass = 1.0 / self.mass            # mInvMassR; mInvMassZ = 1 / (mass + massOcean)

    def checkCompatibility(self, nr):
        pass


class MassOcean3D:
    """MassOcean3D(const RDColX &mass, const RDColX &massOcean, const RDMatX3 &normal) -- MassOcean3D.cpp:9-16."""
    is3D = True
    ocean = True

    def __init__(self, mass, massOcean, normal):
        self.mass = np.asarray(mass, dtype=np.float64).reshape(-1)
        self.massOcean = np.asarray(massOcean, dtype=np.float64).reshape(-1)
        self.normal = np.asarray(normal, dtype=np.float64).reshape(self.mass.size, 3)
        self.invMass = _f(1.0 / self.mass)

    def checkCompatibility(self, nr):
        if self.mass.shape[0] != nr:
            raise RuntimeError("MassOcean3D::checkCompatibility || Incompatible size.")


# --------------------------------------------------------------------------- points
class Point:
    def __init__(self, nr, axial, crds):
        self.nr = int(nr)
        self.nu = self.nr // 2          # Point.cpp:8
        self.axial = bool(axial)
        self.crds = _f(crds, (2,))
        self.domain_tag = -1


class SolidPoint(Point):
    """SolidPoint(int nr, bool axial, const RDCol2 &crds, Mass *mass) -- SolidPoint.cpp:9"""
    kind = "solid"

    def __init__(self, nr, axial, crds, mass):
        super().__init__(nr, axial, crds)
        mass.checkCompatibility(self.nr)
        self.mass = mass


class FluidPoint(Point):
    """FluidPoint(int nr, bool axial, crds, Mass *mass, bool fluidSurf) -- FluidPoint.cpp:9"""
    kind = "fluid"

    def __init__(self, nr, axial, crds, mass, fluidSurf):
        super().__init__(nr, axial, crds)
        mass.checkCompatibility(self.nr)
        self.mass = mass
        self.fluidSurf = bool(fluidSurf)


class SFCoupling1D:
    """SFCoupling1D(ns, nz, ns_invmf, nz_invmf) -- SFCoupling1D.h:11"""
    is3D = False

    def __init__(self, ns, nz, ns_invmf, nz_invmf):
        self.ns, self.nz = float(ns), float(nz)
        self.ns_invmf, self.nz_invmf = float(ns_invmf), float(nz_invmf)

    def checkCompatibility(self, nr):
        pass


class SFCoupling3D:
    """SFCoupling3D(RMatX3 n_unassembled, RMatX3 n_assembled_invMassFluid) -- SFCoupling3D.h"""
    is3D = True

    def __init__(self, normal_unassembled, normal_assembled_invMassFluid):
        self.n_un = _f(normal_unassembled)
        self.n_as = _f(normal_assembled_invMassFluid)
        if self.n_un.ndim != 2 or self.n_un.shape[1] != 3 or self.n_as.shape != self.n_un.shape:
            raise ValueError("SFCoupling3D expects two (Nr, 3) arrays")

    def checkCompatibility(self, nr):
        if self.n_un.shape[0] != nr or self.n_as.shape[0] != nr:
            raise RuntimeError("SFCoupling3D::checkCompatibility || Incompatible size.")


class SolidFluidPoint(Point):
    """SolidFluidPoint(SolidPoint*, FluidPoint*, SFCoupling*) -- SolidFluidPoint.cpp:12-19"""
    kind = "solidfluid"

    def __init__(self, sp, fp, couple):
        super().__init__(sp.nr, sp.axial, sp.crds)
        couple.checkCompatibility(self.nr)
        if sp.nr != fp.nr:
            raise RuntimeError("SolidFluidPoint::SolidFluidPoint || Incompatible size.")
        self.solid, self.fluid, self.couple = sp, fp, couple


# ------------------------------------------------------------------------- gradient
class Gradient:
    """Gradient(dsdxii, dsdeta, dzdxii, dzdeta, inv_s : RDMatPP, bool axial) -- Gradient.cpp:9"""

    def __init__(self, dsdxii, dsdeta, dzdxii, dzdeta, inv_s, axial):
        self.dsdxii = _f(dsdxii, (5, 5))
        self.dsdeta = _f(dsdeta, (5, 5))
        self.dzdxii = _f(dzdxii, (5, 5))
        self.dzdeta = _f(dzdeta, (5, 5))
        self.inv_s = _f(inv_s, (5, 5))
        self.axial = bool(axial)


# ---------------------------------------------------------------------- attenuation
class _Att:
    def __init__(self, nsls, alpha, beta, gamma, dkappa, dmu, doKappa):
        self.nsls = int(nsls)
        self.alpha = _f(alpha).reshape(-1)
        self.beta = _f(beta).reshape(-1)
        self.gamma = _f(gamma).reshape(-1)
        if not (self.alpha.size == self.beta.size == self.gamma.size == self.nsls):
            raise ValueError("alpha/beta/gamma must have nsls entries")
        self.dkappa = _f(dkappa)
        self.dmu = _f(dmu)
        self.doKappa = bool(doKappa)


class Attenuation1D_Full(_Att):
    """Attenuation1D_Full(nsls, alpha, beta, gamma, Nu, RMatPP dkappa, RMatPP dmu, doKappa)"""
    is3D, cg4 = False, False

    def __init__(self, nsls, alpha, beta, gamma, Nu, dkappa, dmu, doKappa):
        super().__init__(nsls, alpha, beta, gamma, _f(dkappa, (5, 5)), _f(dmu, (5, 5)), doKappa)
        self.Nu = int(Nu)

    def checkCompatibility(self, Nr):
        if Nr // 2 != self.Nu:
            raise RuntimeError("Attenuation1D_Full::checkCompatibility || Incompatible size.")


class Attenuation1D_CG4(_Att):
    """Attenuation1D_CG4(nsls, alpha, beta, gamma, Nu, RRow4 dkappa, RRow4 dmu, doKappa);
    the 4 points are (1,1),(1,3),(3,1),(3,3) (Attenuation1D_CG4.cpp:24-27)."""
    is3D, cg4 = False, True

    def __init__(self, nsls, alpha, beta, gamma, Nu, dkappa, dmu, doKappa):
        super().__init__(nsls, alpha, beta, gamma, _f(dkappa, (4,)), _f(dmu, (4,)), doKappa)
        self.Nu = int(Nu)

    def checkCompatibility(self, Nr):
        if Nr // 2 != self.Nu:
            raise RuntimeError("Attenuation1D_CG4::checkCompatibility || Incompatible size.")


class Attenuation3D_Full(_Att):
    """Attenuation3D_Full(nsls, alpha, beta, gamma, RMatXN dkappa, RMatXN dmu, doKappa)"""
    is3D, cg4 = True, False

    def __init__(self, nsls, alpha, beta, gamma, dkappa, dmu, doKappa):
        super().__init__(nsls, alpha, beta, gamma, dkappa, dmu, doKappa)
        if self.dkappa.ndim != 2 or self.dkappa.shape[1] != nPE or self.dmu.shape != self.dkappa.shape:
            raise ValueError("Attenuation3D_Full expects (Nr, 25) dkappa/dmu")

    def checkCompatibility(self, Nr):
        if self.dmu.shape[0] != Nr:
            raise RuntimeError("Attenuation3D_Full::checkCompatibility || Incompatible size.")


class Attenuation3D_CG4(_Att):
    """Attenuation3D_CG4(nsls, alpha, beta, gamma, RMatX4 dkappa, RMatX4 dmu, doKappa)"""
    is3D, cg4 = True, True

    def __init__(self, nsls, alpha, beta, gamma, dkappa, dmu, doKappa):
        super().__init__(nsls, alpha, beta, gamma, dkappa, dmu, doKappa)
        if self.dkappa.ndim != 2 or self.dkappa.shape[1] != 4 or self.dmu.shape != self.dkappa.shape:
            raise ValueError("Attenuation3D_CG4 expects (Nr, 4) dkappa/dmu")

    def checkCompatibility(self, Nr):
        if self.dmu.shape[0] != Nr:
            raise RuntimeError("Attenuation3D_CG4::checkCompatibility || Incompatible size.")


# -------------------------------------------------------------------------- elastic
class _Elastic:
    """coef: (ncoef, rows, 25) with rows = 1 (1D, Fourier space) or Nr (3D, physical space)."""

    def __init__(self, coefs, is3D, att):
        cs = []
        for c in coefs:
            c = _f(c)
            c = c.reshape(1, nPE) if not is3D else c
            if c.ndim != 2 or c.shape[1] != nPE:
                raise ValueError("%s: bad coefficient shape %s" % (type(self).__name__, c.shape))
            cs.append(c)
        self.coef = np.ascontiguousarray(np.stack(cs, 0))
        self.is3D = bool(is3D)
        if att is not None and att.is3D != self.is3D:
            raise ValueError("attenuation and elasticity live in different spaces")
        self.att = att

    def is1D(self):
        return not self.is3D

    def checkCompatibility(self, Nr):
        if self.att is not None:
            self.att.checkCompatibility(Nr)
        if self.is3D and self.coef.shape[1] != Nr:
            raise RuntimeError("%s::checkCompatibility || Incompatible size." % type(self).__name__)


class Isotropic1D(_Elastic):
    """Isotropic1D(RMatPP lambda, RMatPP mu, Attenuation1D*) -- Isotropic1D.h"""
    law, needTIso = "iso", False

    def __init__(self, lam, mu, att=None):
        super().__init__([lam, mu], False, att)


class Isotropic3D(_Elastic):
    """Isotropic3D(RMatXN lambda, RMatXN mu, Attenuation3D*) -- Isotropic3D.h"""
    law, needTIso = "iso", False

    def __init__(self, lam, mu, att=None):
        super().__init__([lam, mu], True, att)


class TransverselyIsotropic1D(_Elastic):
    """TransverselyIsotropic1D(A, C, F, L, N, att) -- TransverselyIsotropic1D.h"""
    law, needTIso = "ti", True

    def __init__(self, A, C, F, L, N, att=None):
        super().__init__([A, C, F, L, N], False, att)


class TransverselyIsotropic3D(_Elastic):
    law, needTIso = "ti", True

    def __init__(self, A, C, F, L, N, att=None):
        super().__init__([A, C, F, L, N], True, att)


class Anisotropic1D(_Elastic):
    """Anisotropic1D(C11, C12, ... C66 (21 RMatPP, order ANISO_IJ), att) -- Anisotropic1D.h:14-20"""
    law, needTIso = "aniso", True

    def __init__(self, C, att=None):
        if len(C) != 21:
            raise ValueError("Anisotropic1D expects 21 moduli")
        super().__init__(list(C), False, att)


class Anisotropic3D(_Elastic):
    law, needTIso = "aniso", True

    def __init__(self, C, att=None):
        if len(C) != 21:
            raise ValueError("Anisotropic3D expects 21 moduli")
        super().__init__(list(C), True, att)


class Acoustic1D:
    """Acoustic1D(RMatPP K) -- Acoustic1D.h"""
    is3D = False

    def __init__(self, K):
        self.K = _f(K).reshape(1, nPE)

    def is1D(self):
        return True

    def checkCompatibility(self, Nr):
        pass


class Acoustic3D:
    """Acoustic3D(RMatXN K) -- Acoustic3D.h; rows must equal the element Nr (Acoustic3D.cpp:18)."""
    is3D = True

    def __init__(self, K):
        self.K = _f(K)
        if self.K.ndim != 2 or self.K.shape[1] != nPE:
            raise ValueError("Acoustic3D expects (Nr, 25)")

    def is1D(self):
        return False

    def checkCompatibility(self, Nr):
        if self.K.shape[0] != Nr:
            raise RuntimeError("Acoustic3D::checkCompatibility || Incompatible size.")


# ------------------------------------------------------------------------- particle relabelling
class PRT_1D:
    """PRT_1D(const std::array<RMatPP, 4> &X) -- S/core/element/prt/PRT_1D.h:12.  X[k] is 5x5 (ipol, jpol)."""

    def __init__(self, X):
        self.X = _f(np.asarray(X, dtype=np.float64).reshape(4, 1, nPntElem))       # [4][rows = 1][25]

    def is1D(self):
        return True

    def checkCompatibility(self, nr):
        pass


class PRT_3D:
    """PRT_3D(const RMatXN4 &X) -- PRT_3D.cpp:13-19: X is Nr x 100 = [X0 | X1 | X2 | X3], each Nr x 25."""

    def __init__(self, X):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim != 2 or X.shape[1] != 4 * nPntElem:
            raise ValueError("PRT_3D expects an Nr x 100 matrix")
        self.X = _f(np.transpose(X.reshape(X.shape[0], 4, nPntElem), (1, 0, 2)))    # [4][Nr][25]

    def is1D(self):
        return False

    def checkCompatibility(self, nr):
        if self.X.shape[1] != nr:
            raise RuntimeError("PRT_3D::checkCompatibility || Incompatible size.")


# ------------------------------------------------------------------------- elements
class Element:
    def __init__(self, grad, prt, points):
        if len(points) != nPntElem:
            raise ValueError("an element has 25 points")
        self.grad = grad
        self.prt = prt
        self.points = list(points)
        self.maxNr = max(p.nr for p in points)      # Element.cpp:13-18
        self.maxNu = max(p.nu for p in points)
        if prt is not None:
            prt.checkCompatibility(self.maxNr)      # Element.cpp:19-21
        self.domain_tag = -1

    def axial(self):
        return self.points[0].axial                 # Element.cpp:37-39

    def formThetaMat(self):
        """Element.cpp:48-58 with Geodesy::theta: polar angle of every GLL point."""
        th = np.zeros(nPntElem)
        for i, p in enumerate(self.points):
            s, z = p.crds
            r = np.hypot(s, z)
            th[i] = 0.0 if r < 1e-10 else np.arccos(np.clip(z / r, -1.0, 1.0))
        return th.reshape(5, 5)


class SolidElement(Element):
    """SolidElement(Gradient*, PRT*, array<Point*,25>, Elastic*) -- SolidElement.cpp:15-34"""
    kind = "solid"

    def __init__(self, grad, prt, points, elastic):
        super().__init__(grad, prt, points)
        for p in points:
            if p.kind == "fluid":
                raise RuntimeError("Point::scatterDisplToElement || Incompatible point type.")
        elastic.checkCompatibility(self.maxNr)
        self.elastic = elastic
        self.inTIso = bool(prt is not None or elastic.needTIso)      # SolidElement.cpp:21
        if prt is not None and elastic.is1D() != prt.is1D():         # SolidElement.cpp:27-32
            raise RuntimeError("SolidElement::SolidElement || Particle Relabelling and Elasticity are generated in different spaces.")
        self.elem3D = not elastic.is1D()


class FluidElement(Element):
    """FluidElement(Gradient*, PRT*, array<Point*,25>, Acoustic*) -- FluidElement.cpp:15-34"""
    kind = "fluid"

    def __init__(self, grad, prt, points, acoustic):
        super().__init__(grad, prt, points)
        for p in points:
            if p.kind == "solid":
                raise RuntimeError("Point::scatterDisplToElement || Incompatible point type.")
        acoustic.checkCompatibility(self.maxNr)
        self.acoustic = acoustic
        self.inTIso = prt is not None               # only with PRT (FluidElement.cpp:21)
        if prt is not None and acoustic.is1D() != prt.is1D():        # FluidElement.cpp:27-32
            raise RuntimeError("FluidElement::FluidElement || Particle Relabelling and Elasticity are generated in different spaces.")
        self.elem3D = not acoustic.is1D()


class SourceTerm:
    """SourceTerm(Element*, const arPP_CMatX3 &force) -- SourceTerm.cpp:15-27: every point's
    force block is truncated to that point's Nu+1 rows."""

    def __init__(self, element, force):
        self.element = element
        self.force = []
        for i, f in enumerate(force):
            f = np.asarray(f, dtype=np.complex128).reshape(-1, 3)
            n = element.points[i].nu + 1
            self.force.append(f[:n].copy())


class MessagingInfo:
    """MessagingInfo (S/preloop/utilities/XMPI.h:330-343): neighbour ranks and, per neighbour,
    the local point tags whose stiffness is exchanged, in global-GLL-tag order."""

    def __init__(self, iProcComm, iLocalPoints):
        self.mIProcComm = [int(r) for r in iProcComm]
        self.mILocalPoints = [list(map(int, l)) for l in iLocalPoints]
        self.mNProcComm = len(self.mIProcComm)
        self.mNLocalPoints = [len(l) for l in self.mILocalPoints]
