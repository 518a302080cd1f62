// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Runs the REFERENCE's own hot-path classes on a domain serialised by tests/dump_domain.py.  Everything that does
// arithmetic here is the unmodified reference source, compiled where it lies under /root/reference/SOLVER/src by
// oracle/Makefile.ref: SolidPoint/FluidPoint/SolidFluidPoint, Mass1D/3D, SFCoupling1D/3D, Gradient, FieldFFT,
// SolverFFTW_{1,3,N3,N6,N9}, CrdTransTIso*, every Elastic/Acoustic/Attenuation class, SolidElement/FluidElement,
// SourceTerm.  What is NOT the reference: Eigen and FFTW (absent from the image; oracle/shim/ provides plain-loop
// stand-ins with the documented semantics) and this file, which plays the part of Mesh::release (construction,
// Mesh.cpp:177-208), axisem.cpp:219-232 (static initialisation) and the serial Newmark::solve / Domain verbs
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

static RMatXN take_xn(const std::vector<float> &src, size_t k, int rows) {
    RMatXN m(rows, nPntElem);
    std::memcpy(m.data(), src.data() + k * (size_t)rows * 25, (size_t)rows * 25 * sizeof(float));
    return m;
}
static RMatPP take_pp(const std::vector<float> &src, size_t k) {
    RMatPP m;
    std::memcpy(m.data(), src.data() + k * 25, 25 * sizeof(float));
    return m;
}
static RColX take_col(const std::vector<float> &v) {
    RColX c((int)v.size());
    std::memcpy(c.data(), v.data(), v.size() * sizeof(float));
    return c;
}
static Mass *read_mass(Reader &r) {
    const int n = r.get<int32_t>();
    if (n <= -1000000) {                              // ocean load (GLLPoint.cpp:57-72)
        const int rows = -n - 1000000;
        if (rows == 1) {
            std::vector<double> v = r.vec<double>(3);
            return new MassOcean1D(v[0], v[1], v[2]);
        }
        std::vector<double> m = r.vec<double>(rows), mo = r.vec<double>(rows), nv = r.vec<double>((size_t)3 * rows);
        RDColX mass(rows), massOcean(rows);
        RDMatX3 normal(rows, 3);
        std::memcpy(mass.data(), m.data(), rows * 8);
        std::memcpy(massOcean.data(), mo.data(), rows * 8);
        std::memcpy(normal.data(), nv.data(), (size_t)3 * rows * 8);
        return new MassOcean3D(mass, massOcean, normal);
    }
    std::vector<float> v = r.vec<float>(n);
    if (n == 1) return new Mass1D(v[0]);
    return new Mass3D(take_col(v));
}

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: ref_driver dump.bin out.bin [kick.bin]\n"); return 2; }
    try {
        Reader r(argv[1]);
        char magic[4];
        r.f.read(magic, 4);
        if (std::memcmp(magic, "AX3D", 4) != 0) throw std::runtime_error("ref_driver || bad magic");
        RDMatPP G_GLL, G_GLJ;
        r.f.read(reinterpret_cast<char *>(G_GLL.data()), 25 * 8);
        r.f.read(reinterpret_cast<char *>(G_GLJ.data()), 25 * 8);
        Gradient::setGMat(G_GLL, G_GLJ);

        // ---- points (GLLPoint::release, GLLPoint.cpp:48-128)
        std::vector<Point *> points;
        std::vector<SolidFluidPoint *> sfpoints;
        const int npoints = r.get<int32_t>();
        int maxNr = 1;
        struct PendingPoint { int kind, nr, axial, surf; RDCol2 crds; Mass *m0, *m1; SFCoupling *c; };
        std::vector<PendingPoint> pend;
        for (int ip = 0; ip < npoints; ++ip) {
            PendingPoint q;
            q.kind = r.get<int32_t>(); q.nr = r.get<int32_t>(); q.axial = r.get<int32_t>();
            q.surf = 0; q.m0 = q.m1 = 0; q.c = 0;
            r.f.read(reinterpret_cast<char *>(q.crds.data()), 16);
            maxNr = std::max(maxNr, q.nr);
            pend.push_back(q);
            PendingPoint &p = pend.back();
            if (p.kind == 0) {
                p.m0 = read_mass(r);
            } else if (p.kind == 1) {
                p.m0 = read_mass(r);
                p.surf = r.get<int32_t>();
            } else {
                p.m0 = read_mass(r);
                p.m1 = read_mass(r);
                p.surf = r.get<int32_t>();
                const int nsf = r.get<int32_t>();
                std::vector<float> un = r.vec<float>(3 * (size_t)nsf), as = r.vec<float>(3 * (size_t)nsf);
                if (nsf == 1) {
                    p.c = new SFCoupling1D(un[0], un[2], as[0], as[2]);
                } else {
                    RMatX3 a(nsf, 3), b(nsf, 3);
                    std::memcpy(a.data(), un.data(), un.size() * sizeof(float));
                    std::memcpy(b.data(), as.data(), as.size() * sizeof(float));
                    p.c = new SFCoupling3D(a, b);
                }
            }
        }
        // ---- static solver state (axisem.cpp:219-232); wisdom import/export is file IO only and skipped
        SolverFFTW_1::initialize(maxNr);
        SolverFFTW_3::initialize(maxNr);
        SolverFFTW_N3::initialize(maxNr);
        SolverFFTW_N6::initialize(maxNr);
        SolverFFTW_N9::initialize(maxNr);
        SolidElement::initWorkspace(maxNr / 2);
        FluidElement::initWorkspace(maxNr / 2);
        for (PendingPoint &p : pend) {
            if (p.kind == 0) points.push_back(new SolidPoint(p.nr, p.axial != 0, p.crds, p.m0));
            else if (p.kind == 1) points.push_back(new FluidPoint(p.nr, p.axial != 0, p.crds, p.m0, p.surf != 0));
            else {
                SolidFluidPoint *sf = new SolidFluidPoint(new SolidPoint(p.nr, p.axial != 0, p.crds, p.m0),
                                                          new FluidPoint(p.nr, p.axial != 0, p.crds, p.m1, p.surf != 0), p.c);
                points.push_back(sf);
                sfpoints.push_back(sf);
            }
        }
        // ---- elements (Quad::release, Quad.cpp:378-420)
        std::vector<Element *> elements;
        const int nelems = r.get<int32_t>();
        for (int ie = 0; ie < nelems; ++ie) {
            const int kind = r.get<int32_t>(), axial = r.get<int32_t>();
            std::vector<int32_t> tags = r.vec<int32_t>(25);
            std::vector<double> geom = r.vec<double>(125);
            RDMatPP g[5];
            for (int k = 0; k < 5; ++k) std::memcpy(g[k].data(), geom.data() + 25 * k, 25 * 8);
            Gradient *grad = new Gradient(g[0], g[1], g[2], g[3], g[4], axial != 0);
            PRT *prt = 0;                                                     // Quad::createRelabelling (Quad.cpp:527-547)
            const int prt_rows = r.get<int32_t>();
            if (prt_rows > 0) {
                std::vector<float> X = r.vec<float>((size_t)4 * 25 * prt_rows);
                if (prt_rows == 1) {
                    std::array<RMatPP, 4> Xs;
                    for (int k = 0; k < 4; ++k) Xs[k] = take_pp(X, k);
                    prt = new PRT_1D(Xs);
                } else {
                    RMatXN4 Xf(prt_rows, 4 * nPntElem);
                    std::memcpy(Xf.data(), X.data(), X.size() * sizeof(float));   // [k][point][row] = column-major Nr x 100
                    prt = new PRT_3D(Xf);
                }
            }
            std::array<Point *, nPntElem> pts;
            for (int i = 0; i < 25; ++i) pts[i] = points[tags[i]];
            if (kind == 0) {
                const int law = r.get<int32_t>(), rows = r.get<int32_t>();
                const int ncoef = law == 0 ? 2 : law == 1 ? 5 : 21;
                std::vector<float> coef = r.vec<float>((size_t)ncoef * rows * 25);
                const int att_kind = r.get<int32_t>();
                Attenuation1D *att1 = 0;
                Attenuation3D *att3 = 0;
                if (att_kind != 0) {
                    const int nsls = r.get<int32_t>(), dok = r.get<int32_t>();
                    RColX al = take_col(r.vec<float>(nsls)), be = take_col(r.vec<float>(nsls)), ga = take_col(r.vec<float>(nsls));
                    const int P = att_kind == 2 ? 4 : 25;
                    std::vector<float> dk = r.vec<float>((size_t)rows * P), dm = r.vec<float>((size_t)rows * P);
                    int maxNu = 0;
                    for (int i = 0; i < 25; ++i) maxNu = std::max(maxNu, pts[i]->getNu());
                    if (rows == 1 && P == 25) att1 = new Attenuation1D_Full(nsls, al, be, ga, maxNu, take_pp(dk, 0), take_pp(dm, 0), dok != 0);
                    else if (rows == 1) {
                        RRow4 a, b;
                        for (int i = 0; i < 4; ++i) { a(i) = dk[i]; b(i) = dm[i]; }
                        att1 = new Attenuation1D_CG4(nsls, al, be, ga, maxNu, a, b, dok != 0);
                    } else if (P == 25) att3 = new Attenuation3D_Full(nsls, al, be, ga, take_xn(dk, 0, rows), take_xn(dm, 0, rows), dok != 0);
                    else {
                        RMatX4 a(rows, 4), b(rows, 4);
                        std::memcpy(a.data(), dk.data(), dk.size() * sizeof(float));
                        std::memcpy(b.data(), dm.data(), dm.size() * sizeof(float));
                        att3 = new Attenuation3D_CG4(nsls, al, be, ga, a, b, dok != 0);
                    }
                }
                Elastic *el;
                if (rows == 1) {
                    RMatPP C[21];
                    for (int k = 0; k < ncoef; ++k) C[k] = take_pp(coef, k);
                    if (law == 0) el = new Isotropic1D(C[0], C[1], att1);
                    else if (law == 1) el = new TransverselyIsotropic1D(C[0], C[1], C[2], C[3], C[4], att1);
                    else el = new Anisotropic1D(C[0], C[1], C[2], C[3], C[4], C[5], C[6], C[7], C[8], C[9], C[10], C[11], C[12], C[13],
                                                C[14], C[15], C[16], C[17], C[18], C[19], C[20], att1);
                } else {
                    std::vector<RMatXN> C;
                    for (int k = 0; k < ncoef; ++k) C.push_back(take_xn(coef, k, rows));
                    if (law == 0) el = new Isotropic3D(C[0], C[1], att3);
                    else if (law == 1) el = new TransverselyIsotropic3D(C[0], C[1], C[2], C[3], C[4], att3);
                    else el = new Anisotropic3D(C[0], C[1], C[2], C[3], C[4], C[5], C[6], C[7], C[8], C[9], C[10], C[11], C[12], C[13],
                                                C[14], C[15], C[16], C[17], C[18], C[19], C[20], att3);
                }
                elements.push_back(new SolidElement(grad, prt, pts, el));
            } else {
                const int rows = r.get<int32_t>();
                std::vector<float> K = r.vec<float>((size_t)rows * 25);
                Acoustic *ac = rows == 1 ? (Acoustic *)new Acoustic1D(take_pp(K, 0)) : (Acoustic *)new Acoustic3D(take_xn(K, 0, rows));
                elements.push_back(new FluidElement(grad, prt, pts, ac));
            }
        }
        // ---- sources (Source::release, Source.cpp:30-59)
        std::vector<SourceTerm *> sources;
        const int nsrc = r.get<int32_t>();
        for (int is = 0; is < nsrc; ++is) {
            const int etag = r.get<int32_t>();
            std::vector<int32_t> nrow = r.vec<int32_t>(25);
            arPP_CMatX3 force;
            for (int i = 0; i < 25; ++i) {
                force[i] = CMatX3(nrow[i], 3);
                std::vector<float> v = r.vec<float>((size_t)6 * nrow[i]);
                std::memcpy(static_cast<void *>(force[i].data()), v.data(), v.size() * sizeof(float));
            }
            sources.push_back(new SourceTerm(elements[etag], force));
        }
        const int nsteps = r.get<int32_t>();
        const double dt = r.get<double>();
        std::vector<float> stf = r.vec<float>(nsteps);

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
            if (pend[ip].kind != 1) {
                const CMatX3 &u = p->getDispFourierSolid();
                out.write(reinterpret_cast<const char *>(u.data()), (size_t)u.size() * sizeof(Complex));
            }
            if (pend[ip].kind != 0) {
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
        }
        std::printf("ref_driver ok: %d points, %d elements, %d steps, maxNr %d\n", npoints, nelems, nsteps, maxNr);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
}
