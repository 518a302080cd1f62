// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Runs the REFERENCE's own hot-path classes on a domain serialised by tests/dump_domain.py.  Everything that does
// arithmetic here is the unmodified reference source, compiled where it lies under /root/reference/SOLVER/src by
// oracle/Makefile.ref: SolidPoint/FluidPoint/SolidFluidPoint, Mass1D/3D, SFCoupling1D/3D, Gradient, FieldFFT,
// SolverFFTW_{1,3,N3,N6,N9}, CrdTransTIso*, every Elastic/Acoustic/Attenuation class, SolidElement/FluidElement,
// SourceTerm.  What is NOT the reference: Eigen and FFTW (absent from the image; oracle/shim/ provides plain-loop
// stand-ins with the documented semantics) and this file, which plays the part of Mesh::release (construction,
// Mesh.cpp:177-208: the text of tests/cpp/release_domain.inc, shared with the facade's test driver), axisem.cpp:219-232 (static initialisation) and the serial Newmark::solve / Domain verbs
// (Newmark.cpp:47-93, Domain.cpp:82-109,165-191) -- Domain.cpp itself drags in the recorders, NetCDF and Boost.
//
//   usage: ref_driver <dump.bin> <out.bin> [kick.bin|-] [recv.in recv.out] [wisdom.out|- cutoff] [series.out]
// series.out (optional, needs recv.in): Domain::record after the update of every step (Newmark.cpp:64-70) -- float
// [nsteps][nrec][3] seismograms from Element::computeGroundMotion, the quantity BASELINE.json's 2000-step check is about.
// wisdom.out (optional): Point::learnWisdom(cutoff) is called on every point after every step (Domain::learnWisdom with
// interval 1, Domain.cpp:384-402); the file receives int32 getNuWisdom() per point (Domain::dumpWisdom, Domain.cpp:404-440).
// recv.in (optional): int32 nrec, then per receiver int32 element tag, float phi, float weights[25] (ipol-major);
// recv.out: float[nrec][3] = Element::computeGroundMotion(phi, weights) of the final state (PointwiseRecorder.cpp:62-144).
// kick.bin (optional): one complex64 buffer per point in Point::feedBuffer order; it is added to the stiffness with
// Point::extractBuffer before the first step, so the first updateNewmark turns it into a broadband displacement
// (u = dt^2 M^-1 f) through the reference's own code -- the reference has no public displacement setter.
// out.bin: all displacements (tests/dump_domain.py:read_displacement order) followed by all stiffness buffers
// (Point::feedBuffer order) as they stand after the last step.
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "Acoustic1D.h"
#include "Acoustic3D.h"
#include "Anisotropic1D.h"
#include "Anisotropic3D.h"
#include "Attenuation1D_CG4.h"
#include "Attenuation1D_Full.h"
#include "Attenuation3D_CG4.h"
#include "Attenuation3D_Full.h"
#include "FluidElement.h"
#include "FluidPoint.h"
#include "Gradient.h"
#include "Isotropic1D.h"
#include "Isotropic3D.h"
#include "Mass1D.h"
#include "Mass3D.h"
#include "MassOcean1D.h"
#include "MassOcean3D.h"
#include "PRT_1D.h"
#include "PRT_3D.h"
#include "SFCoupling1D.h"
#include "SFCoupling3D.h"
#include "SolidElement.h"
#include "SolidFluidPoint.h"
#include "SolidPoint.h"
#include "SolverFFTW_1.h"
#include "SolverFFTW_3.h"
#include "SolverFFTW_N3.h"
#include "SolverFFTW_N6.h"
#include "SolverFFTW_N9.h"
#include "SourceTerm.h"
#include "TransverselyIsotropic1D.h"
#include "TransverselyIsotropic3D.h"

struct Reader {
    std::ifstream f;
    explicit Reader(const char *path) : f(path, std::ios::binary) {
        if (!f) throw std::runtime_error(std::string("ref_driver || cannot open ") + path);
    }
    template <typename T> T get() { T v; f.read(reinterpret_cast<char *>(&v), sizeof(T)); return v; }
    template <typename T> std::vector<T> vec(size_t n) {
        std::vector<T> v(n);
        f.read(reinterpret_cast<char *>(v.data()), n * sizeof(T));
        return v;
    }
};

// matrix adapters of this side: raw floats / doubles -> the Eigen typedefs of S/core/eigenc.h and S/preloop/eigenp.h
static RMatXN mk_xn(const std::vector<float> &src, size_t k, int rows) {
    RMatXN m(rows, nPntElem);
    std::memcpy(m.data(), src.data() + k * (size_t)rows * 25, (size_t)rows * 25 * sizeof(float));
    return m;
}
static RMatPP mk_pp(const std::vector<float> &src, size_t k) {
    RMatPP m;
    std::memcpy(m.data(), src.data() + k * 25, 25 * sizeof(float));
    return m;
}
static RMatPP take_pp(const std::vector<float> &src, size_t k) { return mk_pp(src, k); }
static RDMatPP mk_dpp(const double *p) { RDMatPP m; std::memcpy(m.data(), p, 25 * 8); return m; }
static RDCol2 mk_crds(double s, double z) { RDCol2 c; c(0) = s; c(1) = z; return c; }
static RColX mk_col(const std::vector<float> &v) {
    RColX c((int)v.size());
    std::memcpy(c.data(), v.data(), v.size() * sizeof(float));
    return c;
}
static RDColX mk_dcol(const std::vector<double> &v) {
    RDColX c((int)v.size());
    std::memcpy(c.data(), v.data(), v.size() * 8);
    return c;
}
static RDMatX3 mk_dx3(const std::vector<double> &v, int rows) {
    RDMatX3 m(rows, 3);
    std::memcpy(m.data(), v.data(), v.size() * 8);
    return m;
}
static RMatX3 mk_x3(const std::vector<float> &v, int rows) {
    RMatX3 m(rows, 3);
    std::memcpy(m.data(), v.data(), v.size() * sizeof(float));
    return m;
}
static RMatX4 mk_x4(const std::vector<float> &v, int rows) {
    RMatX4 m(rows, 4);
    std::memcpy(m.data(), v.data(), v.size() * sizeof(float));
    return m;
}
static RMatXN4 mk_xn4(const std::vector<float> &v, int rows) {
    RMatXN4 m(rows, 4 * nPntElem);
    std::memcpy(m.data(), v.data(), v.size() * sizeof(float));   // [k][point][row] = column-major Nr x 100
    return m;
}
static RRow4 mk_row4(const float *p) {
    RRow4 r;
    for (int i = 0; i < 4; ++i) r(i) = p[i];
    return r;
}
static CMatX3 mk_cx3(const std::vector<float> &v, int nrow) {
    CMatX3 m(nrow, 3);
    std::memcpy(static_cast<void *>(m.data()), v.data(), v.size() * sizeof(float));
    return m;
}

// the part of Domain that Mesh::release talks to (Domain.h:30-39, Domain.cpp:46-56): containers only
struct RefDomain {
    std::vector<Point *> points;
    std::vector<SolidFluidPoint *> sfpoints;
    std::vector<Element *> elements;
    std::vector<SourceTerm *> sources;
    int addPoint(Point *p) { points.push_back(p); return (int)points.size() - 1; }
    void addSFPoint(SolidFluidPoint *p) { sfpoints.push_back(p); }
    int addElement(Element *e) { elements.push_back(e); return (int)elements.size() - 1; }
    void addSourceTerm(SourceTerm *s) { sources.push_back(s); }
    Point *getPoint(int i) const { return points[i]; }
    Element *getElement(int i) const { return elements[i]; }
};

#include "../tests/cpp/release_domain.inc"

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: ref_driver dump.bin out.bin [kick.bin|-] [recv.in recv.out] [wisdom.out|- cutoff] [series.out]\n"); return 2; }
    try {
        Reader r(argv[1]);
        RefDomain dom;
        ReleaseInfo info;
        release_domain(r, &dom, info);                          // the construction text shared with tests/cpp/host_driver.cpp
        std::vector<Point *> &points = dom.points;
        std::vector<SolidFluidPoint *> &sfpoints = dom.sfpoints;
        std::vector<Element *> &elements = dom.elements;
        std::vector<SourceTerm *> &sources = dom.sources;
        const int npoints = info.npoints, nelems = info.nelems, nsteps = info.nsteps, maxNr = info.maxNr;
        const double dt = info.dt;
        const std::vector<float> &stf = info.stf;
        // ---- static solver state (axisem.cpp:219-232); wisdom import/export is file IO only and skipped
        SolverFFTW_1::initialize(maxNr);
        SolverFFTW_3::initialize(maxNr);
        SolverFFTW_N3::initialize(maxNr);
        SolverFFTW_N6::initialize(maxNr);
        SolverFFTW_N9::initialize(maxNr);
        SolidElement::initWorkspace(maxNr / 2);
        FluidElement::initWorkspace(maxNr / 2);

        // ---- Newmark::solve (Newmark.cpp:19-93), serial: Domain::resetZero, then the step verbs in order
        for (Element *e : elements) e->resetZero();
        for (Point *p : points) p->resetZero();
        if (argc > 3 && std::string(argv[3]) != "-") {
            std::ifstream kf(argv[3], std::ios::binary);
            if (!kf) throw std::runtime_error("ref_driver || cannot open kick file");
            for (Point *p : points) {
                CColX buf(p->sizeComm());
                kf.read(reinterpret_cast<char *>(buf.data()), (size_t)p->sizeComm() * sizeof(Complex));
                int row = 0;
                p->extractBuffer(buf, row);
            }
        }
        struct Recv { int etag; float phi; RMatPP w; };
        std::vector<Recv> recvs;
        std::ofstream series;
        if (argc > 8) {
            Reader rr(argv[4]);
            const int nrec = rr.get<int32_t>();
            for (int ir = 0; ir < nrec; ++ir) {
                Recv q;
                q.etag = rr.get<int32_t>();
                q.phi = rr.get<float>();
                q.w = take_pp(rr.vec<float>(25), 0);
                recvs.push_back(q);
            }
            series.open(argv[8], std::ios::binary);
        }
        const bool learn = argc > 7 && std::string(argv[6]) != "-";
        for (int tstep = 1; tstep <= nsteps; ++tstep) {
            for (Point *p : points) p->updateNewmark(dt);                     // Domain.cpp:165-176
            for (const Recv &q : recvs) {                                     // Domain::record (Domain.cpp:207-220)
                RRow3 u;
                elements[q.etag]->computeGroundMotion(q.phi, q.w, u);
                const float o[3] = {u(0), u(1), u(2)};
                series.write(reinterpret_cast<const char *>(o), sizeof(o));
            }
            for (SourceTerm *s : sources) s->apply(stf[tstep - 1]);          // Domain.cpp:96-109
            for (Element *e : elements) e->computeStiff();                    // Domain.cpp:82-94
            for (SolidFluidPoint *sf : sfpoints) sf->coupleSolidFluid();      // Domain.cpp:178-191
            if (learn)
                for (Point *p : points) p->learnWisdom((Real)std::atof(argv[7]));   // Domain.cpp:384-402
        }
        if (learn) {
            std::ofstream wo(argv[6], std::ios::binary);
            for (Point *p : points) {
                const int32_t nw = p->getNuWisdom();
                wo.write(reinterpret_cast<const char *>(&nw), 4);
            }
        }

        std::ofstream out(argv[2], std::ios::binary);
        for (size_t ip = 0; ip < points.size(); ++ip) {
            Point *p = points[ip];
            if (info.point_kind[ip] != 1) {
                const CMatX3 &u = p->getDispFourierSolid();
                out.write(reinterpret_cast<const char *>(u.data()), (size_t)u.size() * sizeof(Complex));
            }
            if (info.point_kind[ip] != 0) {
                const CColX &u = p->getDispFourierFluid();
                out.write(reinterpret_cast<const char *>(u.data()), (size_t)u.size() * sizeof(Complex));
            }
        }
        for (Point *p : points) {
            CColX buf(p->sizeComm());
            int row = 0;
            p->feedBuffer(buf, row);
            out.write(reinterpret_cast<const char *>(buf.data()), (size_t)buf.size() * sizeof(Complex));
        }
        if (argc > 5) {
            const bool has_prt = info.has_prt;
            Reader rr(argv[4]);
            std::ofstream ro(argv[5], std::ios::binary);
            const int nrec = rr.get<int32_t>();
            for (int ir = 0; ir < nrec; ++ir) {
                const int etag = rr.get<int32_t>();
                const float phi = rr.get<float>();
                std::vector<float> wv = rr.vec<float>(25);
                RMatPP w = take_pp(wv, 0);
                RRow3 u;
                elements[etag]->computeGroundMotion(phi, w, u);
                const float o[3] = {u(0), u(1), u(2)};
                ro.write(reinterpret_cast<const char *>(o), sizeof(o));
            }
            // ... then strain and curl at the same receivers (PointwiseRecorder.cpp:96-135), solid elements without PRT only;
            // forceTIso first, as ReceiverCollection::release does for stations that dump them
            rr.f.seekg(4);
            for (int ir = 0; ir < nrec; ++ir) {
                const int etag = rr.get<int32_t>();
                const float phi = rr.get<float>();
                std::vector<float> wv = rr.vec<float>(25);
                float o[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                if (dynamic_cast<SolidElement *>(elements[etag]) && !has_prt) {
                    RMatPP w = take_pp(wv, 0);
                    elements[etag]->forceTIso();
                    RRow6 st;
                    RRow3 cu;
                    elements[etag]->computeStrain(phi, w, st);
                    elements[etag]->computeCurl(phi, w, cu);
                    for (int k = 0; k < 6; ++k) o[k] = st(k);
                    for (int k = 0; k < 3; ++k) o[6 + k] = cu(k);
                }
                ro.write(reinterpret_cast<const char *>(o), sizeof(o));
            }
        }
        std::printf("ref_driver ok: %d points, %d elements, %d steps, maxNr %d\n", npoints, nelems, nsteps, maxNr);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
}
