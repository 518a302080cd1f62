// oracle/ref_main_dump.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Runs the reference's own preloop (the call sequence of axisem_main, S/axisem.cpp:12-176, on the reference's classes compiled
// unmodified by oracle/Makefile.main) on an input/ directory and writes what Mesh::release / Source::release / STF::release /
// ReceiverCollection::release put into the reference's Domain, in the "AX3D" serialisation of tests/dump_domain.py (the
// sequence of constructor arguments at the boundary of include/axisem3d_b200.h) followed by a receiver section.  The
// repo's preloop restatement (axisem3d_b200/exodus_mesh.py, source.py, receiver.py) is pinned against this dump array by
// array (tests/test_preloop_reference.py), and the dump can be replayed through the CUDA path like any other dumped domain.
//
// The reference keeps these arrays in private members with no accessors; this one translation unit reads them by compiling
// the reference's HEADERS with `private` / `protected` spelled `public` (the object files it links against are the normal
// ones of Makefile.main; access specifiers change neither layout nor symbol names under this compiler).
//
//     ./axisem3d_dump <out.bin> [solve] [thin N]     (input/ and output/ next to the executable, like the reference)
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>
#include <Eigen/Dense>
#include <boost/algorithm/string.hpp>
#include <boost/lexical_cast.hpp>
#include <boost/geometry.hpp>
#include <netcdf.h>

#define private public
#define protected public
#include "axisem.h"
#include "MultilevelTimer.h"
#include "eigenc.h"
#include "eigenp.h"
#include "Point.h"
#include "SolidPoint.h"
#include "FluidPoint.h"
#include "SolidFluidPoint.h"
#include "Mass1D.h"
#include "Mass3D.h"
#include "MassOcean1D.h"
#include "MassOcean3D.h"
#include "SFCoupling1D.h"
#include "SFCoupling3D.h"
#include "Element.h"
#include "SolidElement.h"
#include "FluidElement.h"
#include "Gradient.h"
#include "PRT.h"
#include "PRT_1D.h"
#include "PRT_3D.h"
#include "Isotropic1D.h"
#include "TransverselyIsotropic1D.h"
#include "Anisotropic1D.h"
#include "Isotropic3D.h"
#include "TransverselyIsotropic3D.h"
#include "Anisotropic3D.h"
#include "Attenuation1D_CG4.h"
#include "Attenuation1D_Full.h"
#include "Attenuation3D_CG4.h"
#include "Attenuation3D_Full.h"
#include "Acoustic1D.h"
#include "Acoustic3D.h"
#include "SourceTerm.h"
#include "SourceTimeFunction.h"
#include "PointwiseRecorder.h"
#undef private
#undef protected

static std::vector<char> out;
static void put(const void *p, size_t n) { out.insert(out.end(), (const char *)p, (const char *)p + n); }
static void i32(int v) { put(&v, 4); }
static void f32(float v) { put(&v, 4); }
// "thin" dumps (large 3-D cases): the material / attenuation / PRT arrays of the elements whose index is not a multiple of
// g_thin are written as NaN (the container stays parseable and compresses well); everything else is complete
static int g_thin = 1;
static bool g_skip = false;
static void f64(double v) { put(&v, 8); }
[[noreturn]] static void die(const std::string &what) { throw std::runtime_error("ref_main_dump || " + what); }

// rows x cols matrix -> float32, column-major (the layout DumpDomain._colmajor writes)
template <class M> static void colmajor_f32(const M &m) {
    for (int j = 0; j < m.cols(); ++j)
        for (int i = 0; i < m.rows(); ++i) f32(g_skip ? std::nanf("") : (float)m(i, j));
}
// 5x5 structured matrix -> 25 floats in point order ipol * 5 + jpol
template <class M> static void struct_f32(const M &m) {
    for (int i = 0; i < nPntEdge; ++i)
        for (int j = 0; j < nPntEdge; ++j) f32(g_skip ? std::nanf("") : (float)m(i, j));
}

static void dump_mass(const Mass *m) {
    if (const Mass1D *m1 = dynamic_cast<const Mass1D *>(m)) {
        i32(1);
        f32((float)m1->mInvMass);
    } else if (const Mass3D *m3 = dynamic_cast<const Mass3D *>(m)) {
        i32((int)m3->mInvMass.rows());
        for (int i = 0; i < m3->mInvMass.rows(); ++i) f32((float)m3->mInvMass(i));
    } else if (const MassOcean1D *mo = dynamic_cast<const MassOcean1D *>(m)) {
        // DumpDomain's ocean marker, then mass, massOcean, theta as doubles -- recovered from the four fp32 members the class keeps
        // (MassOcean1D.cpp:8-13: 1 / (mass + massOcean), 1 / mass, sin theta, cos theta)
        i32(-1000001);
        const double mass = 1. / (double)mo->mInvMassR;
        f64(mass);
        f64(1. / (double)mo->mInvMassZ - mass);
        f64(std::atan2((double)mo->mSint, (double)mo->mCost));
    } else if (const MassOcean3D *m3o = dynamic_cast<const MassOcean3D *>(m)) {
        // MassOcean3D.cpp:9-16 keeps 1 / mass and normal * sqrt(massOcean / (mass (mass + massOcean))) in fp32; the unit normal,
        // mass and massOcean the constructor was given are recovered from them: marker, then mass[rows], massOcean[rows], normal
        // (rows x 3, column-major) as doubles
        const int rows = (int)m3o->mInvMass.rows();
        i32(-1000000 - rows);
        std::vector<double> mass(rows), scal(rows);
        for (int i = 0; i < rows; ++i) {
            mass[i] = 1. / (double)m3o->mInvMass(i);
            scal[i] = std::sqrt((double)m3o->mNormal_scal(i, 0) * m3o->mNormal_scal(i, 0) + (double)m3o->mNormal_scal(i, 1) * m3o->mNormal_scal(i, 1) +
                                (double)m3o->mNormal_scal(i, 2) * m3o->mNormal_scal(i, 2));
        }
        for (int i = 0; i < rows; ++i) f64(mass[i]);
        for (int i = 0; i < rows; ++i) f64(scal[i] * scal[i] * mass[i] * mass[i] / (1. - scal[i] * scal[i] * mass[i]));
        for (int c = 0; c < 3; ++c) for (int i = 0; i < rows; ++i) f64((double)m3o->mNormal_scal(i, c) / scal[i]);
    } else die("unknown Mass class");
}

static void dump_attenuation(const Attenuation *a, int rows) {
    if (!a) { i32(0); return; }
    const Attenuation1D_CG4 *c1 = dynamic_cast<const Attenuation1D_CG4 *>(a);
    const Attenuation1D_Full *f1 = dynamic_cast<const Attenuation1D_Full *>(a);
    const Attenuation3D_CG4 *c3 = dynamic_cast<const Attenuation3D_CG4 *>(a);
    const Attenuation3D_Full *f3 = dynamic_cast<const Attenuation3D_Full *>(a);
    const bool cg4 = c1 || c3;
    const bool doKappa = c1 ? c1->mDoKappa : f1 ? f1->mDoKappa : c3 ? c3->mDoKappa : f3->mDoKappa;
    i32(cg4 ? 2 : 1);
    i32(a->mNSLS);
    i32(doKappa ? 1 : 0);
    for (int i = 0; i < a->mNSLS; ++i) f32((float)a->mAlpha(i));
    for (int i = 0; i < a->mNSLS; ++i) f32((float)a->mBeta(i));
    for (int i = 0; i < a->mNSLS; ++i) f32((float)a->mGamma(i));
    // the classes keep 3 * dkappa and dmu; [rows][P] written column-major, P = 4 (CG4) or 25 (Full)
    if (c1) {
        for (int p = 0; p < 4; ++p) f32((float)((double)c1->mDKappa3(p) / 3.));
        for (int p = 0; p < 4; ++p) f32((float)c1->mDMu(p));
    } else if (f1) {
        for (int i = 0; i < nPntEdge; ++i) for (int j = 0; j < nPntEdge; ++j) f32((float)((double)f1->mDKappa3(i, j) / 3.));
        struct_f32(f1->mDMu);
    } else if (c3) {
        if (c3->mDKappa3.rows() != rows) die("Attenuation3D_CG4 rows");
        for (int p = 0; p < 4; ++p) for (int i = 0; i < rows; ++i) f32((float)((double)c3->mDKappa3(i, p) / 3.));
        colmajor_f32(c3->mDMu);
    } else {
        if (f3->mDKappa3.rows() != rows) die("Attenuation3D_Full rows");
        for (int p = 0; p < nPntElem; ++p) for (int i = 0; i < rows; ++i) f32((float)((double)f3->mDKappa3(i, p) / 3.));
        colmajor_f32(f3->mDMu);
    }
}

template <class M> static void coef1d(std::initializer_list<const M *> ms) {
    for (const M *m : ms) struct_f32(*m);
}
template <class M> static void coef3d(std::initializer_list<const M *> ms) {
    for (const M *m : ms) colmajor_f32(*m);           // [Nr][25] column-major = [25][Nr] = DumpDomain's transpose(coef, (0, 2, 1))
}

static void dump_domain(const Domain &d, double dt) {
    put("AX3D", 4);
    for (int i = 0; i < nPntEdge; ++i) for (int j = 0; j < nPntEdge; ++j) f64(SpectralConstants::getG_GLL()(i, j));
    for (int i = 0; i < nPntEdge; ++i) for (int j = 0; j < nPntEdge; ++j) f64(SpectralConstants::getG_GLJ()(i, j));
    // ---- points
    i32((int)d.mPoints.size());
    for (const Point *p : d.mPoints) {
        const SolidFluidPoint *sf = dynamic_cast<const SolidFluidPoint *>(p);
        const SolidPoint *sp = dynamic_cast<const SolidPoint *>(p);
        const FluidPoint *fp = dynamic_cast<const FluidPoint *>(p);
        i32(sf ? 2 : sp ? 0 : 1);
        i32(p->mNr);
        i32(p->mAxial ? 1 : 0);
        f64(p->mCoords(0));
        f64(p->mCoords(1));
        if (sf) {
            dump_mass(sf->mSolidPoint->mMass);
            dump_mass(sf->mFluidPoint->mMass);
            i32(sf->mFluidPoint->mFluidSurf ? 1 : 0);
            if (const SFCoupling1D *c1 = dynamic_cast<const SFCoupling1D *>(sf->mSFCoupling)) {
                i32(1);
                f32((float)c1->mNormalS_unassembled); f32(0.f); f32((float)c1->mNormalZ_unassembled);
                f32((float)c1->mNormalS_assembled_invMassFluid); f32(0.f); f32((float)c1->mNormalZ_assembled_invMassFluid);
            } else {
                const SFCoupling3D *c3 = dynamic_cast<const SFCoupling3D *>(sf->mSFCoupling);
                i32(p->mNr);
                colmajor_f32(c3->mNormal_unassembled);
                colmajor_f32(c3->mNormal_assembled_invMassFluid);
            }
        } else if (sp) {
            dump_mass(sp->mMass);
        } else {
            dump_mass(fp->mMass);
            i32(fp->mFluidSurf ? 1 : 0);
        }
    }
    // ---- elements
    i32((int)d.mElements.size());
    int ielem = 0;
    for (const Element *e : d.mElements) {
        g_skip = (ielem++ % g_thin) != 0;
        const SolidElement *se = dynamic_cast<const SolidElement *>(e);
        const FluidElement *fe = dynamic_cast<const FluidElement *>(e);
        const Gradient *g = e->mGradient;
        i32(se ? 0 : 1);
        i32(g->mAxial ? 1 : 0);
        for (int k = 0; k < nPntElem; ++k) i32(e->mPoints[k]->getDomainTag());
        for (const RMatPP *m : {&g->mDsDxii, &g->mDsDeta, &g->mDzDxii, &g->mDzDeta, &g->mInv_s})
            for (int i = 0; i < nPntEdge; ++i) for (int j = 0; j < nPntEdge; ++j) f64((double)(*m)(i, j));
        // particle relabelling: rows of X (0 = none, 1 = PRT_1D, Nr = PRT_3D), then X as [4][25][rows]
        if (!e->mHasPRT) {
            i32(0);
        } else if (const PRT_1D *p1 = dynamic_cast<const PRT_1D *>(e->mPRT)) {
            i32(1);
            for (int k = 0; k < 4; ++k) struct_f32(p1->mXStruct[k]);
        } else {
            const PRT_3D *p3 = dynamic_cast<const PRT_3D *>(e->mPRT);
            i32((int)p3->mXFlat0.rows());
            coef3d<RMatXN>({&p3->mXFlat0, &p3->mXFlat1, &p3->mXFlat2, &p3->mXFlat3});
        }
        if (se) {
            const Elastic *el = se->mElastic;
            const Attenuation *att = 0;
            int rows = 1;
            if (const Isotropic1D *m = dynamic_cast<const Isotropic1D *>(el)) {
                i32(0); i32(1); coef1d<RMatPP>({&m->mLambda, &m->mMu}); att = m->mAttenuation;
            } else if (const TransverselyIsotropic1D *m = dynamic_cast<const TransverselyIsotropic1D *>(el)) {
                i32(1); i32(1); coef1d<RMatPP>({&m->mA, &m->mC, &m->mF, &m->mL, &m->mN}); att = m->mAttenuation;
            } else if (const Anisotropic1D *m = dynamic_cast<const Anisotropic1D *>(el)) {
                i32(2); i32(1);
                coef1d<RMatPP>({&m->mC11, &m->mC12, &m->mC13, &m->mC14, &m->mC15, &m->mC16, &m->mC22, &m->mC23, &m->mC24, &m->mC25,
                                &m->mC26, &m->mC33, &m->mC34, &m->mC35, &m->mC36, &m->mC44, &m->mC45, &m->mC46, &m->mC55, &m->mC56,
                                &m->mC66});
                att = m->mAttenuation;
            } else if (const Isotropic3D *m = dynamic_cast<const Isotropic3D *>(el)) {
                rows = (int)m->mLambda.rows();
                i32(0); i32(rows); coef3d<RMatXN>({&m->mLambda, &m->mMu}); att = m->mAttenuation;
            } else if (const TransverselyIsotropic3D *m = dynamic_cast<const TransverselyIsotropic3D *>(el)) {
                rows = (int)m->mA.rows();
                i32(1); i32(rows); coef3d<RMatXN>({&m->mA, &m->mC, &m->mF, &m->mL, &m->mN}); att = m->mAttenuation;
            } else if (const Anisotropic3D *m = dynamic_cast<const Anisotropic3D *>(el)) {
                rows = (int)m->mC11.rows();
                i32(2); i32(rows);
                coef3d<RMatXN>({&m->mC11, &m->mC12, &m->mC13, &m->mC14, &m->mC15, &m->mC16, &m->mC22, &m->mC23, &m->mC24, &m->mC25,
                                &m->mC26, &m->mC33, &m->mC34, &m->mC35, &m->mC36, &m->mC44, &m->mC45, &m->mC46, &m->mC55, &m->mC56,
                                &m->mC66});
                att = m->mAttenuation;
            } else die("unknown Elastic class");
            dump_attenuation(att, rows);
        } else {
            if (const Acoustic1D *a = dynamic_cast<const Acoustic1D *>(fe->mAcoustic)) {
                i32(1);
                struct_f32(a->mKStruct);
            } else {
                const Acoustic3D *a3 = dynamic_cast<const Acoustic3D *>(fe->mAcoustic);
                i32((int)a3->mKFlat.rows());
                colmajor_f32(a3->mKFlat);
            }
        }
    }
    g_skip = false;
    // ---- source terms
    i32((int)d.mSourceTerms.size());
    for (const SourceTerm *st : d.mSourceTerms) {
        i32(st->mElement->getDomainTag());
        for (int k = 0; k < nPntElem; ++k) i32((int)st->mForce[k].rows());
        for (int k = 0; k < nPntElem; ++k)
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < st->mForce[k].rows(); ++r) {
                    f32((float)st->mForce[k](r, c).real());
                    f32((float)st->mForce[k](r, c).imag());
                }
    }
    // ---- source time function
    const SourceTimeFunction *stf = d.mSTF;
    i32((int)stf->mSTF.size());
    f64(dt);
    for (Real v : stf->mSTF) f32((float)v);
    // ---- receivers (after the part tests/cpp/host_driver.cpp reads)
    put("RECV", 4);
    f64(stf->mShift);
    const PointwiseRecorder *rec = d.mPointwiseRecorder;
    i32(rec ? (int)rec->mPointwiseInfo.size() : 0);
    if (rec) {
        put(rec->mComponents.c_str(), 3);
        for (const PointwiseInfo &r : rec->mPointwiseInfo) {
            const std::string key = r.mNetwork + "." + r.mName;
            i32((int)key.size());
            put(key.data(), key.size());
            i32(r.mElement->getDomainTag());
            f64(r.mPhi); f64(r.mTheta); f64(r.mBAz); f64(r.mLat); f64(r.mLon); f64(r.mDep);
            for (int i = 0; i < nPntEdge; ++i) for (int j = 0; j < nPntEdge; ++j) f64((double)r.mWeights(i, j));
        }
    }
}

extern "C" void set_ftz();          // S/ftz.c, called first thing by the reference's main (main.cpp:12)

int main(int argc, char *argv[]) {
    set_ftz();
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <out.bin> [solve] [thin N]\n", argv[0]);
        return 2;
    }
    bool solve = false;
    for (int a = 2; a < argc; ++a) {
        if (std::string(argv[a]) == "solve") solve = true;
        if (std::string(argv[a]) == "thin" && a + 1 < argc) g_thin = std::max(1, std::atoi(argv[++a]));
    }
    try {
        // the call sequence of axisem_main (axisem.cpp:12-176)
        PreloopVariables pl;
        SolverVariables sv;
        XMPI::initialize(argc, argv);
        SpectralConstants::initialize(nPol);
        int verbose;
        Parameters::buildInparam(pl.mParameters, verbose);
        MultilevelTimer::initialize(Parameters::sOutputDirectory + "/develop/preloop_timer.txt", 4);
        ExodusModel::buildInparam(pl.mExodusModel, *(pl.mParameters), pl.mAttParameters, verbose);
        NrField::buildInparam(pl.mNrField, *(pl.mParameters), verbose);
        Source::buildInparam(pl.mSource, *(pl.mParameters), verbose);
        const double srcLat = pl.mSource->getLatitude(), srcLon = pl.mSource->getLongitude(), srcDep = pl.mSource->getDepth();
        Volumetric3D::buildInparam(pl.mVolumetric3D, *(pl.mParameters), pl.mExodusModel, srcLat, srcLon, srcDep, verbose);
        Geometric3D::buildInparam(pl.mGeometric3D, *(pl.mParameters), verbose);
        OceanLoad3D::buildInparam(pl.mOceanLoad3D, *(pl.mParameters), verbose);
        pl.mMesh = new Mesh(pl.mExodusModel, pl.mNrField, srcLat, srcLon, srcDep, *(pl.mParameters), verbose);
        pl.mMesh->setVolumetric3D(pl.mVolumetric3D);
        pl.mMesh->setGeometric3D(pl.mGeometric3D);
        pl.mMesh->setOceanLoad3D(pl.mOceanLoad3D);
        pl.mMesh->buildUnweighted();
        initializeSolverStatic(pl.mMesh->getMaxNr(), pl.mParameters->getValue<bool>("FFTW_DISABLE_WISDOM"));
        double dt = pl.mParameters->getValue<double>("TIME_DELTA_T");
        if (dt < tinyDouble) dt = pl.mMesh->getDeltaT();
        double dt_fact = pl.mParameters->getValue<double>("TIME_DELTA_T_FACTOR");
        if (dt_fact < tinyDouble) dt_fact = 1.0;
        dt *= dt_fact;
        AttBuilder::buildInparam(pl.mAttBuilder, *(pl.mParameters), pl.mAttParameters, dt, verbose);
        pl.mMesh->setAttBuilder(pl.mAttBuilder);
        pl.mMesh->buildWeighted();
        STF::buildInparam(pl.mSTF, *(pl.mParameters), dt, verbose);
        ReceiverCollection::buildInparam(pl.mReceivers, *(pl.mParameters), srcLat, srcLon, srcDep, pl.mSTF->getSize(), verbose);
        sv.mDomain = new Domain();
        pl.mMesh->release(*(sv.mDomain));
        pl.mSource->release(*(sv.mDomain), *(pl.mMesh));
        pl.mSTF->release(*(sv.mDomain));
        pl.mReceivers->release(*(sv.mDomain), *(pl.mMesh), pl.mParameters->getValue<bool>("OUT_STATIONS_DEPTH_REF"));
        sv.mDomain->initializeRecorders();
        MultilevelTimer::finalize();

        dump_domain(*sv.mDomain, dt);
        std::ofstream f(argv[1], std::ios::binary);
        f.write(out.data(), (std::streamsize)out.size());
        f.close();
        std::printf("ref_main_dump: %zu points, %zu elements, %zu bytes, dt %.17g\n", sv.mDomain->mPoints.size(),
                    sv.mDomain->mElements.size(), out.size(), dt);

        if (solve) {
            const int infoInt = pl.mParameters->getValue<int>("OPTION_LOOP_INFO_INTERVAL");
            const int stabInt = pl.mParameters->getValue<int>("OPTION_STABILITY_INTERVAL");
            sv.mNewmark = new Newmark(sv.mDomain, infoInt, stabInt, pl.mParameters->getValue<bool>("DEVELOP_RANDOMIZE_DISP0"));
            pl.finalize();
            sv.mNewmark->solve(0);
            sv.mDomain->finalizeRecorders();
        }
        finalizeSolverStatic();
        XMPI::finalize();
    } catch (const std::exception &e) {
        XMPI::cout.setp(XMPI::rank());
        XMPI::printException(e);
        return 1;
    }
    return 0;
}
