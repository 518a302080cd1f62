// ax3d_host.hpp -- C++ host side of the drop-in boundary: the solver-object classes that AxiSEM3D's `Mesh::release`
// constructs and `Newmark::solve` drives, re-implemented as thin descriptors over the C-ABI of
// include/axisem3d_b200.h.  Class names, constructor argument order, method names and error text follow the reference
// (S/ = SOLVER/src of kuangdai/AxiSEM3D; file:line cited per class) so that the reference's driver sequence
//
//     Domain *domain = new Domain();  mesh->release(*domain);  source->release(...);  stf->release(...);
//     Newmark *nm = new Newmark(domain, reportInterval, checkStabInterval, randomDispl);  nm->solve(verbose);
//
// compiles against this header unchanged apart from the matrix typedefs: Eigen is not available in this build, so the
// fixed/dynamic Eigen types of S/core/eigenc.h are stood in for by the plain containers below (same storage order;
// with Eigen present `m.data()`, `m.rows()` of the Eigen objects are passed instead -- INTEGRATION.md).
//
// All arithmetic happens on the GPU inside libaxisem3d_b200.so; nothing here computes.  There is no CPU fallback: the
// Domain constructor throws if no CUDA device is usable.
#pragma once
#include <array>
#include <cmath>
#include <complex>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/axisem3d_b200.h"

namespace ax3d {

// ------------------------------------------------------------------------------------------ S/global.h, S/core/eigenc.h
typedef float Real;
typedef std::complex<Real> Complex;
static const int nPol = 4, nPntEdge = 5, nPntElem = 25, nPE = 25;
typedef std::array<double, 2> RDCol2;
typedef std::array<double, 25> RDMatPP;   // row-major (ipol, jpol)
typedef std::array<Real, 25> RMatPP;      // row-major (ipol, jpol)
typedef std::array<Real, 3> RRow3;
typedef std::array<Real, 6> RRow6;
typedef std::array<Real, 4> RRow4;        // eigen_cg4.h: the four CG4 points (1,1) (1,3) (3,1) (3,3)
typedef std::vector<Real> RColX;
struct RMatXN {                           // Nr x 25, column-major: (j, ipnt) at [ipnt * rows + j]
    int rows = 0;
    std::vector<Real> v;
    RMatXN() {}
    RMatXN(int r) : rows(r), v((size_t)r * 25, Real(0)) {}
    Real &operator()(int j, int ipnt) { return v[(size_t)ipnt * rows + j]; }
    const Real *data() const { return v.data(); }
};
struct RMatX3 {                           // Nr x 3, column-major
    int rows = 0;
    std::vector<Real> v;
    RMatX3() {}
    RMatX3(int r) : rows(r), v((size_t)r * 3, Real(0)) {}
    Real &operator()(int j, int c) { return v[(size_t)c * rows + j]; }
    const Real *data() const { return v.data(); }
};
struct RMatX4 {                           // Nr x 4, column-major (CG4 points)
    int rows = 0;
    std::vector<Real> v;
    RMatX4() {}
    RMatX4(int r) : rows(r), v((size_t)r * 4, Real(0)) {}
    const Real *data() const { return v.data(); }
};
struct CMatX3 {                           // (Nu + 1) x 3 complex, column-major
    int rows = 0;
    std::vector<Complex> v;
    CMatX3() {}
    CMatX3(int r) : rows(r), v((size_t)r * 3, Complex(0)) {}
    Complex &operator()(int a, int c) { return v[(size_t)c * rows + a]; }
    const Complex &operator()(int a, int c) const { return v[(size_t)c * rows + a]; }
};
typedef std::vector<Complex> CColX;
typedef std::array<CMatX3, 25> arPP_CMatX3;

inline void check(int rc) {
    if (rc != 0) throw std::runtime_error(ax3d_last_error());
}

// ------------------------------------------------------------------------------------------ point/mass  (S/core/point/mass)
class Mass {
public:
    virtual ~Mass() {}
    virtual int size() const = 0;          // 1 (Mass1D) or Nr (Mass3D)
    virtual const Real *data() const = 0;
    virtual void checkCompatibility(int nr) const {}
    // ocean-load masses hand double-precision descriptors to the library instead (0: not an ocean mass)
    virtual int oceanRows() const { return 0; }
    virtual const double *oceanMass() const { return 0; }
    virtual const double *oceanLoad() const { return 0; }
    virtual const double *oceanNormalOrTheta() const { return 0; }
};
class Mass1D : public Mass {               // Mass1D.cpp:8
public:
    explicit Mass1D(Real invMass) : mInvMass(invMass) {}
    int size() const { return 1; }
    const Real *data() const { return &mInvMass; }
private:
    Real mInvMass;
};
class Mass3D : public Mass {               // Mass3D.cpp:9; rows must equal the point's Nr (Mass3D.cpp:59-64)
public:
    explicit Mass3D(const RColX &invMass) : mInvMass(invMass) {}
    int size() const { return (int)mInvMass.size(); }
    const Real *data() const { return mInvMass.data(); }
    void checkCompatibility(int nr) const {
        if ((int)mInvMass.size() != nr) throw std::runtime_error("Mass3D::checkCompatibility || Incompatible size.");
    }
private:
    RColX mInvMass;
};

typedef std::vector<double> RDColX;
typedef std::vector<double> RDMatX3;         // Nr x 3, column-major
class MassOcean1D : public Mass {            // MassOcean1D.cpp:8-13: (mass, massOcean, theta)
public:
    MassOcean1D(double mass, double massOcean, double theta) : mMass(mass), mOcean(massOcean), mTheta(theta), mInvMass((Real)(1. / mass)) {}
    int size() const { return 1; }
    const Real *data() const { return &mInvMass; }
    int oceanRows() const { return 1; }
    const double *oceanMass() const { return &mMass; }
    const double *oceanLoad() const { return &mOcean; }
    const double *oceanNormalOrTheta() const { return &mTheta; }
private:
    double mMass, mOcean, mTheta;
    Real mInvMass;
};
class MassOcean3D : public Mass {            // MassOcean3D.cpp:9-16: (mass[Nr], massOcean[Nr], unit normal Nr x 3)
public:
    MassOcean3D(const RDColX &mass, const RDColX &massOcean, const RDMatX3 &normal) : mMass(mass), mOcean(massOcean), mNormal(normal) {
        for (double m : mass) mInvMass.push_back((Real)(1. / m));
    }
    int size() const { return (int)mMass.size(); }
    const Real *data() const { return mInvMass.data(); }
    void checkCompatibility(int nr) const {
        if ((int)mMass.size() != nr) throw std::runtime_error("MassOcean3D::checkCompatibility || Incompatible size.");
    }
    int oceanRows() const { return (int)mMass.size(); }
    const double *oceanMass() const { return mMass.data(); }
    const double *oceanLoad() const { return mOcean.data(); }
    const double *oceanNormalOrTheta() const { return mNormal.data(); }
private:
    RDColX mMass, mOcean;
    RDMatX3 mNormal;
    RColX mInvMass;
};

// ------------------------------------------------------------------------------------------ points  (S/core/point)
struct LearnParameters {                    // S/preloop/nrfield/wisdom/NuWisdom.h:37-43 (filled from inparam.nu there)
    bool mInvoked = false;
    double mCutoff = 0;
    int mInterval = 1;
    std::string mFileName;
};

class Domain;
class Point {                               // Point.h:13
public:
    Point(int nr, bool axial, const RDCol2 &crds) : mNr(nr), mNu(nr / 2), mAxial(axial), mCoords(crds) {}
    virtual ~Point() {}
    int getNr() const { return mNr; }
    int getNu() const { return mNu; }
    bool axial() const { return mAxial; }
    const RDCol2 &getCoords() const { return mCoords; }
    void setDomainTag(int tag) { mDomainTag = tag; }
    int getDomainTag() const { return mDomainTag; }
    virtual int sizeComm() const = 0;
    // read-back (Point.h:74-75): Fourier coefficients of the displacement, copied from the device
    CMatX3 getDispFourierSolid() const;
    CColX getDispFourierFluid() const;
protected:
    friend class Domain;
    virtual int release(ax3d_domain *dom) = 0;   // Domain::addPoint hands the descriptor to the library
    int mNr, mNu;
    bool mAxial;
    RDCol2 mCoords;
    int mDomainTag = -1;
    ax3d_domain *mDom = nullptr;
};

class SolidPoint : public Point {           // SolidPoint.cpp:9
public:
    SolidPoint(int nr, bool axial, const RDCol2 &crds, Mass *mass) : Point(nr, axial, crds), mMass(mass) {
        mMass->checkCompatibility(nr);
    }
    int sizeComm() const { return 3 * (mNu + 1); }   // SolidPoint.h:40
    const Mass *mass() const { return mMass.get(); }
protected:
    int release(ax3d_domain *dom) {
        int tag = -1;
        if (mMass->oceanRows())
            check(ax3d_add_solid_point_ocean(dom, mNr, mAxial, mCoords.data(), mMass->oceanRows(), mMass->oceanMass(), mMass->oceanLoad(),
                                             mMass->oceanNormalOrTheta(), &tag));
        else
            check(ax3d_add_solid_point(dom, mNr, mAxial, mCoords.data(), mMass->size(), mMass->data(), &tag));
        return tag;
    }
    std::unique_ptr<Mass> mMass;
};

class FluidPoint : public Point {           // FluidPoint.cpp:9
public:
    FluidPoint(int nr, bool axial, const RDCol2 &crds, Mass *mass, bool fluidSurf)
        : Point(nr, axial, crds), mMass(mass), mFluidSurf(fluidSurf) {
        mMass->checkCompatibility(nr);
    }
    int sizeComm() const { return mNu + 1; }         // FluidPoint.h:40
    const Mass *mass() const { return mMass.get(); }
    bool fluidSurf() const { return mFluidSurf; }
protected:
    int release(ax3d_domain *dom) {
        int tag = -1;
        check(ax3d_add_fluid_point(dom, mNr, mAxial, mCoords.data(), mMass->size(), mMass->data(), mFluidSurf, &tag));
        return tag;
    }
    std::unique_ptr<Mass> mMass;
    bool mFluidSurf;
};

class SFCoupling {                          // S/core/point/solid_fluid
public:
    virtual ~SFCoupling() {}
    virtual int size() const = 0;
    virtual const Real *normalUnassembled() const = 0;          // size() x 3 column-major
    virtual const Real *normalAssembledInvMassFluid() const = 0;
    virtual void checkCompatibility(int nr) const {}
};
class SFCoupling1D : public SFCoupling {    // SFCoupling1D.h:11
public:
    SFCoupling1D(Real ns, Real nz, Real ns_invmf, Real nz_invmf) : mUn{ns, 0, nz}, mAs{ns_invmf, 0, nz_invmf} {}
    int size() const { return 1; }
    const Real *normalUnassembled() const { return mUn; }
    const Real *normalAssembledInvMassFluid() const { return mAs; }
private:
    Real mUn[3], mAs[3];
};
class SFCoupling3D : public SFCoupling {    // SFCoupling3D.cpp:9-54
public:
    SFCoupling3D(const RMatX3 &n_unassembled, const RMatX3 &n_assembled_invMassFluid) : mUn(n_unassembled), mAs(n_assembled_invMassFluid) {}
    int size() const { return mUn.rows; }
    const Real *normalUnassembled() const { return mUn.data(); }
    const Real *normalAssembledInvMassFluid() const { return mAs.data(); }
    void checkCompatibility(int nr) const {
        if (mUn.rows != nr || mAs.rows != nr) throw std::runtime_error("SFCoupling3D::checkCompatibility || Incompatible size.");
    }
private:
    RMatX3 mUn, mAs;
};

class SolidFluidPoint : public Point {      // SolidFluidPoint.cpp:12; owns both sub-points and the coupling (:21-25)
public:
    SolidFluidPoint(SolidPoint *sp, FluidPoint *fp, SFCoupling *couple)
        : Point(sp->getNr(), sp->axial(), sp->getCoords()), mSolid(sp), mFluid(fp), mCouple(couple) {
        mCouple->checkCompatibility(mNr);
    }
    int sizeComm() const { return mSolid->sizeComm() + mFluid->sizeComm(); }
protected:
    int release(ax3d_domain *dom) {
        int tag = -1;
        check(ax3d_add_solid_fluid_point(dom, mNr, mAxial, mCoords.data(), mSolid->mass()->size(), mSolid->mass()->data(),
                                         mFluid->mass()->size(), mFluid->mass()->data(), mFluid->fluidSurf(), mCouple->size(),
                                         mCouple->normalUnassembled(), mCouple->normalAssembledInvMassFluid(), &tag));
        return tag;
    }
    std::unique_ptr<SolidPoint> mSolid;
    std::unique_ptr<FluidPoint> mFluid;
    std::unique_ptr<SFCoupling> mCouple;
};

// ------------------------------------------------------------------------------------------ element parts
class Gradient {                            // Gradient.cpp:9; static G matrices Gradient.cpp:324-329
public:
    Gradient(const RDMatPP &dsdxii, const RDMatPP &dsdeta, const RDMatPP &dzdxii, const RDMatPP &dzdeta, const RDMatPP &inv_s, bool axial)
        : mAxial(axial) {
        const RDMatPP *m[5] = {&dsdxii, &dsdeta, &dzdxii, &dzdeta, &inv_s};
        for (int k = 0; k < 5; ++k)
            for (int i = 0; i < 25; ++i) mGeom[k * 25 + i] = (*m[k])[i];
    }
    static void setGMat(const RDMatPP &G_GLL, const RDMatPP &G_GLJ) {
        sG_GLL() = G_GLL;
        sG_GLJ() = G_GLJ;
        sHaveG() = true;
    }
    bool axial() const { return mAxial; }
    const double *geom() const { return mGeom; }
    static RDMatPP &sG_GLL() { static RDMatPP g; return g; }
    static RDMatPP &sG_GLJ() { static RDMatPP g; return g; }
    static bool &sHaveG() { static bool h = false; return h; }
private:
    double mGeom[125];
    bool mAxial;
};

// ------------------------------------------------------------------------------------------ particle relabelling  (S/core/element/prt)
struct RMatXN4 {                          // Nr x 100 = [X0 | X1 | X2 | X3], column-major (eigenc.h: RMatXN4)
    int rows = 0;
    std::vector<Real> v;
    RMatXN4() {}
    RMatXN4(int r) : rows(r), v((size_t)r * 100, Real(0)) {}
    Real &operator()(int j, int col) { return v[(size_t)col * rows + j]; }
    const Real *data() const { return v.data(); }
};
class PRT {                                 // PRT.h:13-33
public:
    virtual ~PRT() {}
    virtual bool is1D() const = 0;
    virtual void checkCompatibility(int /*Nr*/) const {}
    int rows() const { return mRows; }
    const Real *X() const { return mX.data(); }     // [4][25][rows]
protected:
    int mRows = 1;
    std::vector<Real> mX;
};
class PRT_1D : public PRT {                 // PRT_1D.h:12
public:
    PRT_1D(const std::array<RMatPP, 4> &X) {
        mRows = 1;
        for (const RMatPP &x : X) mX.insert(mX.end(), x.begin(), x.end());
    }
    bool is1D() const { return true; }
};
class PRT_3D : public PRT {                 // PRT_3D.cpp:13-19
public:
    PRT_3D(const RMatXN4 &X) {
        mRows = X.rows;
        mX = X.v;                             // column-major Nr x 100 is exactly [4][25][Nr]
    }
    bool is1D() const { return false; }
    void checkCompatibility(int Nr) const {
        if (mRows != Nr) throw std::runtime_error("PRT_3D::checkCompatibility || Incompatible size.");
    }
};

class Attenuation {                         // S/core/element/material/attenuation
public:
    virtual ~Attenuation() {}
    ax3d_attenuation descriptor() const {
        ax3d_attenuation a;
        a.kind = mKind; a.nsls = (int)mAlpha.size();
        a.alpha = mAlpha.data(); a.beta = mBeta.data(); a.gamma = mGamma.data();
        a.dkappa = mDKappa.data(); a.dmu = mDMu.data(); a.do_kappa = mDoKappa;
        return a;
    }
    int rows() const { return mRows; }
protected:
    Attenuation(int kind, int nsls, const RColX &alpha, const RColX &beta, const RColX &gamma, int rows, int P, const Real *dkappa,
                const Real *dmu, bool doKappa)
        : mKind(kind), mRows(rows), mAlpha(alpha), mBeta(beta), mGamma(gamma), mDKappa(dkappa, dkappa + (size_t)rows * P),
          mDMu(dmu, dmu + (size_t)rows * P), mDoKappa(doKappa) {
        if ((int)alpha.size() != nsls || (int)beta.size() != nsls || (int)gamma.size() != nsls)
            throw std::runtime_error("Attenuation::Attenuation || Incompatible number of standard linear solids.");
    }
    int mKind, mRows;
    RColX mAlpha, mBeta, mGamma, mDKappa, mDMu;
    bool mDoKappa;
};
// the two intermediate bases of the reference (Attenuation1D.h:10, Attenuation3D.h:10): what the 1D / 3D elastic classes take
class Attenuation1D : public Attenuation {
protected:
    using Attenuation::Attenuation;
};
class Attenuation3D : public Attenuation {
protected:
    using Attenuation::Attenuation;
};
// Attenuation1D_Full.h / _CG4.h: (nsls, alpha, beta, gamma, Nu, dkappa, dmu, doKappa) with RMatPP / RRow4 moduli
class Attenuation1D_Full : public Attenuation1D {
public:
    Attenuation1D_Full(int nsls, const RColX &alpha, const RColX &beta, const RColX &gamma, int /*Nu*/, const RMatPP &dkappa,
                       const RMatPP &dmu, bool doKappa)
        : Attenuation1D(AX3D_ATT_FULL, nsls, alpha, beta, gamma, 1, 25, dkappa.data(), dmu.data(), doKappa) {}
};
class Attenuation1D_CG4 : public Attenuation1D {
public:
    Attenuation1D_CG4(int nsls, const RColX &alpha, const RColX &beta, const RColX &gamma, int /*Nu*/, const RRow4 &dkappa,
                      const RRow4 &dmu, bool doKappa)
        : Attenuation1D(AX3D_ATT_CG4, nsls, alpha, beta, gamma, 1, 4, dkappa.data(), dmu.data(), doKappa) {}
};
// Attenuation3D_Full.h / _CG4.h: (nsls, alpha, beta, gamma, dkappa, dmu, doKappa) with RMatXN / RMatX4 moduli
class Attenuation3D_Full : public Attenuation3D {
public:
    Attenuation3D_Full(int nsls, const RColX &alpha, const RColX &beta, const RColX &gamma, const RMatXN &dkappa, const RMatXN &dmu,
                       bool doKappa)
        : Attenuation3D(AX3D_ATT_FULL, nsls, alpha, beta, gamma, dkappa.rows, 25, dkappa.data(), dmu.data(), doKappa) {}
};
class Attenuation3D_CG4 : public Attenuation3D {
public:
    Attenuation3D_CG4(int nsls, const RColX &alpha, const RColX &beta, const RColX &gamma, const RMatX4 &dkappa, const RMatX4 &dmu,
                      bool doKappa)
        : Attenuation3D(AX3D_ATT_CG4, nsls, alpha, beta, gamma, dkappa.rows, 4, dkappa.data(), dmu.data(), doKappa) {}
};

class Elastic {                             // S/core/element/material/elastic; owns its Attenuation (Elastic1D.cpp:13-17)
public:
    virtual ~Elastic() {}
    int law() const { return mLaw; }
    int rows() const { return mRows; }
    const Real *coef() const { return mCoef.data(); }
    const Attenuation *attenuation() const { return mAtt.get(); }
    bool is1D() const { return mRows == 1; }
    bool needTIso() const { return mLaw != AX3D_ISO; }
    void checkCompatibility(int Nr) const {
        if (mRows != 1 && mRows != Nr) throw std::runtime_error("Elastic3D::checkCompatibility || Incompatible size.");
        if (mAtt && mAtt->rows() != mRows) throw std::runtime_error("Attenuation3D::checkCompatibility || Incompatible size.");
    }
protected:
    Elastic(int law, int rows, Attenuation *att) : mLaw(law), mRows(rows), mAtt(att) {}
    void push(const RMatPP &m) { mCoef.insert(mCoef.end(), m.begin(), m.end()); }   // one row per point: identical in both orders
    void push(const RMatXN &m) {
        if (m.rows != mRows) throw std::runtime_error("Elastic3D::Elastic3D || Incompatible size.");
        mCoef.insert(mCoef.end(), m.v.begin(), m.v.end());
    }
    int mLaw, mRows;
    std::vector<Real> mCoef;
    std::unique_ptr<Attenuation> mAtt;
};
class Isotropic1D : public Elastic {        // Isotropic1D.cpp:9-26
public:
    Isotropic1D(const RMatPP &lambda, const RMatPP &mu, Attenuation *att = 0) : Elastic(AX3D_ISO, 1, att) { push(lambda); push(mu); }
};
class Isotropic3D : public Elastic {        // Isotropic3D.cpp:10-27
public:
    Isotropic3D(const RMatXN &lambda, const RMatXN &mu, Attenuation *att = 0) : Elastic(AX3D_ISO, lambda.rows, att) { push(lambda); push(mu); }
};
class TransverselyIsotropic1D : public Elastic {   // TransverselyIsotropic1D.cpp:9-26
public:
    TransverselyIsotropic1D(const RMatPP &A, const RMatPP &C, const RMatPP &F, const RMatPP &L, const RMatPP &N, Attenuation *att = 0)
        : Elastic(AX3D_TI, 1, att) { push(A); push(C); push(F); push(L); push(N); }
};
class TransverselyIsotropic3D : public Elastic {   // TransverselyIsotropic3D.cpp:10-28
public:
    TransverselyIsotropic3D(const RMatXN &A, const RMatXN &C, const RMatXN &F, const RMatXN &L, const RMatXN &N, Attenuation *att = 0)
        : Elastic(AX3D_TI, A.rows, att) { push(A); push(C); push(F); push(L); push(N); }
};
class Anisotropic1D : public Elastic {      // Anisotropic1D.cpp:9-54: C11 C12 ... C16 C22 ... C66 (upper triangle, row by row)
public:
    Anisotropic1D(const std::array<RMatPP, 21> &Cij, Attenuation *att = 0) : Elastic(AX3D_ANISO, 1, att) { for (const RMatPP &c : Cij) push(c); }
    // the reference's own 22-argument form (Anisotropic1D.h:13-27)
    Anisotropic1D(const RMatPP &C11, const RMatPP &C12, const RMatPP &C13, const RMatPP &C14, const RMatPP &C15, const RMatPP &C16,
                  const RMatPP &C22, const RMatPP &C23, const RMatPP &C24, const RMatPP &C25, const RMatPP &C26,
                  const RMatPP &C33, const RMatPP &C34, const RMatPP &C35, const RMatPP &C36,
                  const RMatPP &C44, const RMatPP &C45, const RMatPP &C46, const RMatPP &C55, const RMatPP &C56, const RMatPP &C66,
                  Attenuation *att = 0)
        : Elastic(AX3D_ANISO, 1, att) {
        const RMatPP *all[21] = {&C11, &C12, &C13, &C14, &C15, &C16, &C22, &C23, &C24, &C25, &C26, &C33, &C34, &C35, &C36, &C44, &C45, &C46, &C55, &C56, &C66};
        for (const RMatPP *c : all) push(*c);
    }
};
class Anisotropic3D : public Elastic {      // Anisotropic3D.cpp:10-54
public:
    Anisotropic3D(const std::array<RMatXN, 21> &Cij, Attenuation *att = 0) : Elastic(AX3D_ANISO, Cij[0].rows, att) { for (const RMatXN &c : Cij) push(c); }
    // the reference's own 22-argument form (Anisotropic3D.h:13-27)
    Anisotropic3D(const RMatXN &C11, const RMatXN &C12, const RMatXN &C13, const RMatXN &C14, const RMatXN &C15, const RMatXN &C16,
                  const RMatXN &C22, const RMatXN &C23, const RMatXN &C24, const RMatXN &C25, const RMatXN &C26,
                  const RMatXN &C33, const RMatXN &C34, const RMatXN &C35, const RMatXN &C36,
                  const RMatXN &C44, const RMatXN &C45, const RMatXN &C46, const RMatXN &C55, const RMatXN &C56, const RMatXN &C66,
                  Attenuation *att = 0)
        : Elastic(AX3D_ANISO, C11.rows, att) {
        const RMatXN *all[21] = {&C11, &C12, &C13, &C14, &C15, &C16, &C22, &C23, &C24, &C25, &C26, &C33, &C34, &C35, &C36, &C44, &C45, &C46, &C55, &C56, &C66};
        for (const RMatXN *c : all) push(*c);
    }
};

class Acoustic {                            // S/core/element/material/acoustic
public:
    virtual ~Acoustic() {}
    int rows() const { return mRows; }
    const Real *K() const { return mK.data(); }
    void checkCompatibility(int Nr) const {
        if (mRows != 1 && mRows != Nr) throw std::runtime_error("Acoustic3D::checkCompatibility || Incompatible size.");
    }
protected:
    Acoustic(int rows, const Real *K) : mRows(rows), mK(K, K + (size_t)rows * 25) {}
    int mRows;
    std::vector<Real> mK;
};
class Acoustic1D : public Acoustic {        // Acoustic1D.cpp:8
public:
    explicit Acoustic1D(const RMatPP &K) : Acoustic(1, K.data()) {}
};
class Acoustic3D : public Acoustic {        // Acoustic3D.cpp:9-16
public:
    explicit Acoustic3D(const RMatXN &K) : Acoustic(K.rows, K.data()) {}
};

// ------------------------------------------------------------------------------------------ elements  (S/core/element)
class Element {                             // Element.cpp:13-29: owns Gradient (and PRT); points are shared, not owned
public:
    Element(Gradient *grad, PRT *prt, const std::array<Point *, 25> &points) : mGradient(grad), mPRT(prt), mPoints(points) {
        mMaxNr = -1;
        for (Point *p : points) mMaxNr = p->getNr() > mMaxNr ? p->getNr() : mMaxNr;   // Element.cpp:13-18
        mMaxNu = mMaxNr / 2;
        if (mPRT) mPRT->checkCompatibility(mMaxNr);                                   // Element.cpp:19-21
    }
    virtual ~Element() {}
    const Point *getPoint(int index) const { return mPoints[index]; }
    int getMaxNr() const { return mMaxNr; }
    int getMaxNu() const { return mMaxNu; }
    bool axial() const { return mGradient->axial(); }
    void setDomainTag(int tag) { mDomainTag = tag; }
    int getDomainTag() const { return mDomainTag; }
    RDMatPP formThetaMat() const {          // Element.cpp:48-58: polar angle of every GLL point
        RDMatPP th;
        for (int i = 0; i < 25; ++i) {
            const RDCol2 &c = mPoints[i]->getCoords();
            const double r = std::sqrt(c[0] * c[0] + c[1] * c[1]);
            th[i] = r < 1e-12 ? 0. : std::acos(c[1] / r);
        }
        return th;
    }
    // Element::computeGroundMotion (SolidElement.cpp:189-216): evaluated on the device from the current displacement
    void computeGroundMotion(Real phi, const RMatPP &weights, RRow3 &u_spz) const {
        check(ax3d_record_ground_motion(mDom, 1, &mDomainTag, &phi, weights.data(), u_spz.data()));
    }
    // Element.h:26-32 (SolidElement.cpp:219-352): the library applies forceTIso itself for strain / curl receivers
    void computeStrain(Real phi, const RMatPP &weights, RRow6 &strain) const {
        check(ax3d_record_strain(mDom, 1, &mDomainTag, &phi, weights.data(), strain.data()));
    }
    void computeCurl(Real phi, const RMatPP &weights, RRow3 &curl) const {
        check(ax3d_record_curl(mDom, 1, &mDomainTag, &phi, weights.data(), curl.data()));
    }
    void forceTIso() {}
    // the same three with the caller's own matrix types (anything with (i, j) and (i) access, e.g. Eigen's RMatPP / RRow3 / RRow6:
    // PointwiseRecorder.cpp:73-75, 103-105, 116-118 compile against these unchanged)
    template <class W, class O, class = decltype(std::declval<const W &>()(0, 0))>
    void computeGroundMotion(Real phi, const W &weights, O &u_spz) const {
        RMatPP w;
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) w[i * 5 + j] = (Real)weights(i, j);
        RRow3 u;
        computeGroundMotion(phi, w, u);
        for (int c = 0; c < 3; ++c) u_spz(c) = u[c];
    }
    template <class W, class O, class = decltype(std::declval<const W &>()(0, 0))>
    void computeStrain(Real phi, const W &weights, O &strain) const {
        RMatPP w;
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) w[i * 5 + j] = (Real)weights(i, j);
        RRow6 e;
        computeStrain(phi, w, e);
        for (int c = 0; c < 6; ++c) strain(c) = e[c];
    }
    template <class W, class O, class = decltype(std::declval<const W &>()(0, 0))>
    void computeCurl(Real phi, const W &weights, O &curl) const {
        RMatPP w;
        for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) w[i * 5 + j] = (Real)weights(i, j);
        RRow3 c3;
        computeCurl(phi, w, c3);
        for (int c = 0; c < 3; ++c) curl(c) = c3[c];
    }
protected:
    friend class Domain;
    virtual int release(ax3d_domain *dom) = 0;
    void tags(int out[25]) const {
        for (int i = 0; i < 25; ++i) {
            out[i] = mPoints[i]->getDomainTag();
            if (out[i] < 0) throw std::runtime_error("Domain::addElement || a point of this element has not been added to the domain.");
        }
    }
    // hands the PRT* constructor argument to the library (after ax3d_add_*_element)
    void releasePRT(ax3d_domain *dom, int tag) const {
        if (!mPRT) return;
        const RDMatPP theta = formThetaMat();
        check(ax3d_set_element_prt(dom, tag, mPRT->rows(), mPRT->X(), theta.data()));
    }
    std::unique_ptr<Gradient> mGradient;
    std::unique_ptr<PRT> mPRT;
    std::array<Point *, 25> mPoints;
    int mMaxNr, mMaxNu, mDomainTag = -1;
    ax3d_domain *mDom = nullptr;
};

class SolidElement : public Element {       // SolidElement.cpp:15-41: owns Elastic
public:
    SolidElement(Gradient *grad, PRT *prt, const std::array<Point *, 25> &points, Elastic *elas) : Element(grad, prt, points), mElastic(elas) {
        mElastic->checkCompatibility(mMaxNr);
    }
protected:
    int release(ax3d_domain *dom) {
        int t[25], tag = -1;
        tags(t);
        const RDMatPP theta = formThetaMat();
        ax3d_attenuation att;
        const ax3d_attenuation *pa = 0;
        if (mElastic->attenuation()) { att = mElastic->attenuation()->descriptor(); pa = &att; }
        check(ax3d_add_solid_element(dom, t, mGradient->geom(), mGradient->axial(), theta.data(), mElastic->law(), mElastic->rows(),
                                     mElastic->coef(), pa, &tag));
        releasePRT(dom, tag);
        return tag;
    }
    std::unique_ptr<Elastic> mElastic;
};

class FluidElement : public Element {       // FluidElement.cpp:15-41: owns Acoustic
public:
    FluidElement(Gradient *grad, PRT *prt, const std::array<Point *, 25> &points, Acoustic *acous) : Element(grad, prt, points), mAcoustic(acous) {
        mAcoustic->checkCompatibility(mMaxNr);
    }
protected:
    int release(ax3d_domain *dom) {
        int t[25], tag = -1;
        tags(t);
        check(ax3d_add_fluid_element(dom, t, mGradient->geom(), mGradient->axial(), mAcoustic->rows(), mAcoustic->K(), &tag));
        releasePRT(dom, tag);
        return tag;
    }
    std::unique_ptr<Acoustic> mAcoustic;
};

// ------------------------------------------------------------------------------------------ source  (S/core/source)
class SourceTerm {                          // SourceTerm.cpp:15-35
public:
    SourceTerm(Element *element, const arPP_CMatX3 &force) : mElement(element), mForce(force) {}
    Element *element() const { return mElement; }
    const arPP_CMatX3 &force() const { return mForce; }
private:
    Element *mElement;
    arPP_CMatX3 mForce;
};

class SourceTimeFunction {                  // SourceTimeFunction.h:10-24
public:
    SourceTimeFunction(const std::vector<Real> &stf, double dt, double shift) : mDeltaT(dt), mShift(shift), mSTF(stf) {}   // the reference's order
    int getSize() const { return (int)mSTF.size(); }
    double getDeltaT() const { return mDeltaT; }
    double getShift() const { return mShift; }
    Real getFactor(int tstep) const { return mSTF[tstep]; }
    const std::vector<Real> &samples() const { return mSTF; }
private:
    double mDeltaT, mShift;
    std::vector<Real> mSTF;
};

// ------------------------------------------------------------------------------------------ messaging  (XMPI.h:330-348)
struct MessagingInfo {
    int mNProcComm = 0;
    std::vector<int> mIProcComm;                    // neighbour ranks
    std::vector<int> mNLocalPoints;
    std::vector<std::vector<int>> mILocalPoints;    // per neighbour: local point tags in global-GLL-tag order
    // the halo sum runs over NCCL instead of MPI_Isend/Irecv: rank, size and the broadcast ncclUniqueId replace the requests
    int mRank = 0, mNProc = 1;
    std::array<unsigned char, 128> mNcclUniqueId{};
};
struct MessagingBuffer {};                          // the packed buffers live on the device

// ------------------------------------------------------------------------------------------ Domain  (S/core/domain/Domain.h)
class Domain {
public:
    explicit Domain(int device = 0) { check(ax3d_create(device, &mDom)); }
    ~Domain() {                                    // Domain.cpp:27-44: owns points, elements, sources, STF, messaging
        for (Point *p : mPoints) delete p;
        for (Element *e : mElements) delete e;
        for (SourceTerm *s : mSourceTerms) delete s;
        delete mSTF;
        delete mMsgInfo;
        delete mMsgBuffer;
        if (mDom) ax3d_destroy(mDom);
    }
    Domain(const Domain &) = delete;
    Domain &operator=(const Domain &) = delete;

    // ---- methods before the time loop
    int addPoint(Point *point) {                   // Domain.cpp:46-50: tag = insertion index
        sendGMat();
        const int tag = point->release(mDom);
        point->setDomainTag(tag);
        point->mDom = mDom;
        mPoints.push_back(point);
        return tag;
    }
    int addElement(Element *elem) {                // Domain.cpp:52-56
        sendGMat();
        const int tag = elem->release(mDom);
        elem->setDomainTag(tag);
        elem->mDom = mDom;
        mElements.push_back(elem);
        return tag;
    }
    void addSFPoint(SolidFluidPoint *) {}          // the library keeps its own list of solid-fluid points
    void addSourceTerm(SourceTerm *source) {
        int nrow[25];
        std::vector<float> flat;
        for (int i = 0; i < 25; ++i) {
            const CMatX3 &f = source->force()[i];
            nrow[i] = f.rows;
            for (const Complex &z : f.v) { flat.push_back(z.real()); flat.push_back(z.imag()); }
        }
        check(ax3d_add_source_term(mDom, source->element()->getDomainTag(), nrow, flat.data()));
        mSourceTerms.push_back(source);
    }
    void setSTF(SourceTimeFunction *stf) { mSTF = stf; }
    void setMessaging(MessagingInfo *msgInfo, MessagingBuffer *msgBuffer) {
        mMsgInfo = msgInfo;
        mMsgBuffer = msgBuffer;
        std::vector<int> npts, tags;
        for (const std::vector<int> &l : msgInfo->mILocalPoints) {
            npts.push_back((int)l.size());
            tags.insert(tags.end(), l.begin(), l.end());
        }
        check(ax3d_set_messaging(mDom, msgInfo->mRank, msgInfo->mNProc, msgInfo->mNcclUniqueId.data(), (int)msgInfo->mIProcComm.size(),
                                 msgInfo->mIProcComm.data(), npts.data(), tags.data()));
    }
    const SourceTimeFunction &getSTF() const { return *mSTF; }
    int getNumPoints() const { return (int)mPoints.size(); }
    int getNumElements() const { return (int)mElements.size(); }
    Point *getPoint(int index) const { return mPoints[index]; }
    Element *getElement(int index) const { return mElements[index]; }

    void resetZero() const { finalize(); check(ax3d_reset_zero(mDom)); }   // Domain.cpp:67-74

    // ---- methods during the time loop (Newmark.cpp:47-93)
    void computeStiff() const { finalize(); check(ax3d_compute_stiff(mDom)); }
    void applySource(int tstep) const { finalize(); check(ax3d_apply_source(mDom, mSTF ? mSTF->getFactor(tstep) : Real(0))); }
    void assembleStiff(int phase = 0) const { finalize(); check(ax3d_assemble_stiff(mDom, phase)); }
    void updateNewmark(double dt) const { finalize(); check(ax3d_update_newmark(mDom, dt)); }
    void coupleSolidFluid() const { finalize(); check(ax3d_couple_solid_fluid(mDom)); }
    void record(int /*tstep*/, double /*t*/) const {}   // PointwiseRecorder stays on the host side
    // wisdom learning (Domain.h:39; Domain.cpp:384-402, 404-440)
    void setLearnParameters(LearnParameters *lpar) {
        finalize();
        mLearnPar = lpar;
        check(ax3d_set_learn_parameters(mDom, lpar->mInvoked ? 1 : 0, (float)lpar->mCutoff, lpar->mInterval));
    }
    void learnWisdom(int tstep) const { if (mLearnPar && mLearnPar->mInvoked) check(ax3d_learn_wisdom(mDom, tstep)); }
    std::vector<int> getNuWisdom() const {   // Point::getNuWisdom() of every point, as Domain::dumpWisdom collects them
        std::vector<int> nw(mPoints.size());
        check(ax3d_get_nu_wisdom(mDom, nw.data(), (int)nw.size()));
        return nw;
    }
    // Mesh::measure (Mesh.cpp:530-588): measured cost of every element (microseconds of one SM), the METIS vertex weights of
    // the second partition pass; element order = domain tags
    std::vector<double> measureCosts(int repeats = 3) const {
        finalize();
        std::vector<double> cost(mElements.size());
        check(ax3d_measure_costs(mDom, repeats, cost.data(), (int)cost.size()));
        return cost;
    }
    void checkStability(double dt, int tstep, double t) const {            // Domain.cpp:237-275
        int ok = 1;
        check(ax3d_check_stability(mDom, &ok));
        if (!ok) {
            char buf[256];
            std::snprintf(buf, sizeof(buf), "Domain::checkStability || Simulation has blown up. || dt = %g || step = %d || t = %g", dt, tstep, t);
            throw std::runtime_error(buf);
        }
    }
    // whole time loop on the device, one CUDA-graph replay per step (what Newmark::solve uses when no recorder needs the host)
    void runSteps(int nsteps, double dt, const Real *stf) const { finalize(); check(ax3d_run_steps(mDom, nsteps, dt, stf)); }
    // Newmark::solve with Domain::record after every update, recorder buffer on the device (dump interval = nsteps)
    void runStepsRecord(int nsteps, double dt, const Real *stf, Real *out) const {
        finalize();
        check(ax3d_run_steps_record(mDom, nsteps, dt, stf, out));
    }
    void synchronize() const { check(ax3d_synchronize(mDom)); }
    ax3d_domain *handle() const { return mDom; }

private:
    void sendGMat() {
        if (!mSentG) {
            if (!Gradient::sHaveG()) throw std::runtime_error("Gradient::setGMat || G matrices have not been set.");
            check(ax3d_set_gmat(mDom, Gradient::sG_GLL().data(), Gradient::sG_GLJ().data()));
            mSentG = true;
        }
    }
    void finalize() const {                        // end of Mesh::release: first verb after the last add*
        if (!mFinal) {
            check(ax3d_finalize_setup(mDom));
            mFinal = true;
        }
    }
    ax3d_domain *mDom = nullptr;
    std::vector<Point *> mPoints;
    std::vector<Element *> mElements;
    std::vector<SourceTerm *> mSourceTerms;
    SourceTimeFunction *mSTF = nullptr;
    LearnParameters *mLearnPar = nullptr;
    MessagingInfo *mMsgInfo = nullptr;
    MessagingBuffer *mMsgBuffer = nullptr;
    bool mSentG = false;
    mutable bool mFinal = false;
};

inline CMatX3 Point::getDispFourierSolid() const {
    CMatX3 out(mNu + 1);
    check(ax3d_get_point_field(mDom, mDomainTag, AX3D_DISPL, 0, reinterpret_cast<float *>(out.v.data()), 3 * (mNu + 1)));
    return out;
}
inline CColX Point::getDispFourierFluid() const {
    CColX out(mNu + 1);
    check(ax3d_get_point_field(mDom, mDomainTag, AX3D_DISPL, 1, reinterpret_cast<float *>(out.data()), mNu + 1));
    return out;
}

// ------------------------------------------------------------------------------------------ Newmark  (S/core/newmark)
class Newmark {
public:
    Newmark(Domain *&domain, int reportInterval, int checkStabInterval, bool randomDispl)
        : mDomain(domain), mReportInterval(reportInterval <= 0 ? 100 : reportInterval),
          mCheckStabInterval(checkStabInterval <= 0 ? mReportInterval : checkStabInterval), mRandomDispl(randomDispl) {}
    // Newmark::solve (Newmark.cpp:19-93): same verb order per step; the recorder hook stays between the two assemble phases
    void solve(int verbose) const {
        double t = 0. - mDomain->getSTF().getShift();
        const double dt = mDomain->getSTF().getDeltaT();
        const int maxStep = mDomain->getSTF().getSize();
        mDomain->resetZero();
        if (mRandomDispl) throw std::runtime_error("Newmark::solve || DEVELOP_RANDOMIZE_DISP0 is not supported by the B200 path.");
        for (int tstep = 1; tstep <= maxStep; tstep++) {
            mDomain->updateNewmark(dt);
            mDomain->applySource(tstep - 1);
            mDomain->computeStiff();
            mDomain->coupleSolidFluid();
            mDomain->assembleStiff(-1);
            mDomain->record(tstep - 1, t);
            t += dt;
            if (tstep % mCheckStabInterval == 0) mDomain->checkStability(dt, tstep, t);
            if (verbose && tstep % mReportInterval == 0) std::printf("  step %d / %d   t = %g\n", tstep, maxStep, t);
            mDomain->learnWisdom(tstep - 1);
            mDomain->assembleStiff(1);
        }
        mDomain->synchronize();
    }
private:
    Domain *mDomain;
    int mReportInterval, mCheckStabInterval;
    bool mRandomDispl;
};

}   // namespace ax3d
