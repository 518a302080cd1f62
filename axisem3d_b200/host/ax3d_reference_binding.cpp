// ax3d_reference_binding.cpp -- the members of the binding's Domain / Newmark that need the reference's own classes complete:
// PointwiseRecorder (compiled from S/core/output/pointwise, unmodified), LearnParameters / NuWisdom, MessagingInfo.
// Built only by oracle/Makefile.dropin, next to the reference's sources (see ax3d_reference_binding.hpp).
#include "ax3d_reference_binding.hpp"

#include <cmath>
#include <cstdio>

#include "NuWisdom.h"
#include "PointwiseRecorder.h"
#include "SurfaceRecorder.h"
#include "XMPI.h"

static int device_from_env() {
    const char *e = std::getenv("AX3D_DEVICE");
    return e ? std::atoi(e) : 0;
}

Domain::Domain() : mImpl(device_from_env()) {}

Domain::~Domain() {                                     // Domain.cpp:27-44; the ax3d points / elements / sources go with mImpl
    for (Point *p : mPoints) delete p;                  // shells
    for (Element *e : mElements) delete e;
    delete mPointwiseRecorder;
    delete mSurfaceRecorder;
    if (mSTF) { mSTF->take(); delete mSTF; }            // the ax3d::SourceTimeFunction is deleted by ax3d::Domain
    delete mMsgInfo;
    delete mMsgBuffer;
    delete mLearn;
}

void Domain::setMessaging(MessagingInfo *msgInfo, MessagingBuffer *msgBuffer) {
    mMsgInfo = msgInfo;
    mMsgBuffer = msgBuffer;
    if (msgInfo->mNProcComm != 0)
        throw std::runtime_error("Domain::setMessaging || this binding is the serial build of the reference (one rank); "
                                 "the multi-rank halo goes through ax3d_set_messaging / ax3d_halo_connect (INTEGRATION.md section 1.4).");
}

void Domain::setLearnParameters(LearnParameters *lpar) {
    mLearn = lpar;
    mLearnBound.mInvoked = lpar->mInvoked;
    mLearnBound.mCutoff = lpar->mCutoff;
    mLearnBound.mInterval = lpar->mInterval;
    mLearnBound.mFileName = lpar->mFileName;
    // Mesh::release calls this before Source::release adds the source term (Mesh.cpp:207, axisem.cpp:133-140), and the library's
    // ax3d_set_learn_parameters finalizes the set-up: the parameters are handed on at the first verb of Newmark::solve
}

void Domain::resetZero() const {                        // Newmark.cpp:34
    if (mLearn && mLearn->mInvoked) mImpl.setLearnParameters(const_cast<ax3d::LearnParameters *>(&mLearnBound));
    mImpl.resetZero();
}

void Domain::initializeRecorders() const {              // Domain.cpp:193-205
    if (mPointwiseRecorder) mPointwiseRecorder->initialize();
    if (mSurfaceRecorder) throw std::runtime_error("Domain::initializeRecorders || OUT_STATIONS_WHOLE_SURFACE is not bound to the B200 path.");
}
void Domain::finalizeRecorders() const {
    if (mPointwiseRecorder) mPointwiseRecorder->finalize();
}
void Domain::record(int tstep, double t) const {        // Domain.cpp:207-219: the reference's own recorder, reading back through
    if (mPointwiseRecorder) mPointwiseRecorder->record(tstep, t);   // Element::computeGroundMotion -> ax3d_record_ground_motion
}
void Domain::dumpLeft() const {
    if (mPointwiseRecorder) mPointwiseRecorder->dumpToFile();
}

void Domain::dumpWisdom() const {                       // Domain.cpp:404-440
    if (!mLearn || !mLearn->mInvoked) return;
    const std::vector<int> nw = mImpl.getNuWisdom();
    NuWisdom wis;
    for (int i = 0; i < getNumPoints(); ++i) {
        const Point *p = getPoint(i);
        wis.insert(p->getCoords()(0), p->getCoords()(1), nw[i], p->getNu());
    }
    wis.writeToFile(mLearn->mFileName);
}

std::string Domain::verbose() const {                   // Domain.cpp:277-346, class names abridged to the kinds the boundary knows
    std::stringstream ss;
    ss << "\n=================== Computational Domain ===================" << std::endl;
    ss << "  Elements (B200 path)   =   " << getNumElements() << std::endl;
    ss << "  GLL Points             =   " << getNumPoints() << std::endl;
    ss << "=================== Computational Domain ===================\n" << std::endl;
    return ss.str();
}

Newmark::Newmark(Domain *&domain, int reportInterval, int checkStabInterval, bool randomDispl)
    : mDomain(domain), mReportInterval(reportInterval <= 0 ? 100 : reportInterval),
      mCheckStabInterval(checkStabInterval <= 0 ? mReportInterval : checkStabInterval), mRandomDispl(randomDispl) {}

void Newmark::solve(int verbose) const {
    double t = 0. - mDomain->getSTF().getShift();
    const double dt = mDomain->getSTF().getDeltaT();
    const int maxStep = mDomain->getSTF().getSize();
    mDomain->resetZero();
    if (mRandomDispl) mDomain->initDisplTinyRandom();
    for (int tstep = 1; tstep <= maxStep; tstep++) {
        mDomain->updateNewmark(dt);
        mDomain->applySource(tstep - 1);
        mDomain->computeStiff();
        mDomain->coupleSolidFluid();
        mDomain->assembleStiff(-1);
        mDomain->record(tstep - 1, t);
        t += dt;
        if (tstep % mCheckStabInterval == 0) mDomain->checkStability(dt, tstep, t);
        if (verbose && tstep % mReportInterval == 0) {
            XMPI::cout << "  SIMULATION TIME / sec     =   " << t << XMPI::endl;
            XMPI::cout << "  TIME STEP / TOTAL STEPS   =   " << tstep << " / " << maxStep << XMPI::endl << XMPI::endl;
        }
        mDomain->learnWisdom(tstep - 1);
        mDomain->assembleStiff(1);
    }
    mDomain->synchronize();
    mDomain->dumpLeft();
    mDomain->dumpWisdom();
}
