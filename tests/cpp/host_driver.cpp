// host_driver.cpp -- test driver of the C++ facade (axisem3d_b200/host/ax3d_host.hpp): plays the role of the reference's
// axisem_main (S/axisem.cpp:127-181) for a domain serialised by tests/dump_domain.py -- constructs the solver objects
// exactly as Mesh::release / Source::release / STF::release would, runs Newmark::solve, and writes every point's
// displacement so that the Python test can compare it with the oracle.
//   usage: host_driver <dump.bin> <out.bin>      exit code 0 = ran; 3 = no CUDA device (message on stderr)
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>

#include "../../axisem3d_b200/host/ax3d_host.hpp"

using namespace ax3d;

struct Reader {
    std::ifstream f;
    explicit Reader(const char *path) : f(path, std::ios::binary) {
        if (!f) throw std::runtime_error(std::string("host_driver || cannot open ") + path);
    }
    template <typename T> T get() { T v; f.read(reinterpret_cast<char *>(&v), sizeof(T)); return v; }
    template <typename T> std::vector<T> vec(size_t n) {
        std::vector<T> v(n);
        f.read(reinterpret_cast<char *>(v.data()), n * sizeof(T));
        return v;
    }
};

static RMatXN take_xn(const std::vector<float> &src, size_t k, int rows) {
    RMatXN m(rows);
    std::memcpy(m.v.data(), src.data() + k * (size_t)rows * 25, (size_t)rows * 25 * sizeof(float));
    return m;
}
static RMatPP take_pp(const std::vector<float> &src, size_t k) {
    RMatPP m;
    std::memcpy(m.data(), src.data() + k * 25, 25 * sizeof(float));
    return m;
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
        return new MassOcean3D(m, mo, nv);
    }
    std::vector<float> v = r.vec<float>(n);
    if (n == 1) return new Mass1D(v[0]);
    return new Mass3D(v);
}

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: host_driver dump.bin out.bin\n"); return 2; }
    try {
        Reader r(argv[1]);
        char magic[4];
        r.f.read(magic, 4);
        if (std::memcmp(magic, "AX3D", 4) != 0) throw std::runtime_error("host_driver || bad magic");
        RDMatPP G_GLL, G_GLJ;
        r.f.read(reinterpret_cast<char *>(G_GLL.data()), 25 * 8);
        r.f.read(reinterpret_cast<char *>(G_GLJ.data()), 25 * 8);
        Gradient::setGMat(G_GLL, G_GLJ);                       // SpectralConstants::initialize -> Gradient::setGMat

        Domain *domain = new Domain(0);                         // axisem.cpp:127
        // ---- GLLPoint::release (GLLPoint.cpp:48-128)
        const int npoints = r.get<int32_t>();
        for (int ip = 0; ip < npoints; ++ip) {
            const int kind = r.get<int32_t>(), nr = r.get<int32_t>(), axial = r.get<int32_t>();
            RDCol2 crds;
            r.f.read(reinterpret_cast<char *>(crds.data()), 16);
            if (kind == 0) {
                domain->addPoint(new SolidPoint(nr, axial != 0, crds, read_mass(r)));
            } else if (kind == 1) {
                Mass *m = read_mass(r);
                const int surf = r.get<int32_t>();
                domain->addPoint(new FluidPoint(nr, axial != 0, crds, m, surf != 0));
            } else {
                Mass *ms = read_mass(r);
                Mass *mf = read_mass(r);
                const int surf = r.get<int32_t>();
                const int nsf = r.get<int32_t>();
                std::vector<float> un = r.vec<float>(3 * (size_t)nsf), as = r.vec<float>(3 * (size_t)nsf);
                SFCoupling *c;
                if (nsf == 1) {
                    c = new SFCoupling1D(un[0], un[2], as[0], as[2]);
                } else {
                    RMatX3 a(nsf), b(nsf);
                    a.v = un;
                    b.v = as;
                    c = new SFCoupling3D(a, b);
                }
                SolidFluidPoint *sfp = new SolidFluidPoint(new SolidPoint(nr, axial != 0, crds, ms), new FluidPoint(nr, axial != 0, crds, mf, surf != 0), c);
                domain->addPoint(sfp);
                domain->addSFPoint(sfp);
            }
        }
        // ---- Quad::release (Quad.cpp:378-420)
        const int nelems = r.get<int32_t>();
        for (int ie = 0; ie < nelems; ++ie) {
            const int kind = r.get<int32_t>(), axial = r.get<int32_t>();
            std::vector<int32_t> tags = r.vec<int32_t>(25);
            std::vector<double> geom = r.vec<double>(125);
            RDMatPP g[5];
            for (int k = 0; k < 5; ++k) std::memcpy(g[k].data(), geom.data() + 25 * k, 25 * 8);
            Gradient *grad = new Gradient(g[0], g[1], g[2], g[3], g[4], axial != 0);
            PRT *prt = 0;                                             // Quad::createRelabelling (Quad.cpp:527-547)
            const int prt_rows = r.get<int32_t>();
            if (prt_rows > 0) {
                std::vector<float> X = r.vec<float>((size_t)4 * 25 * prt_rows);
                if (prt_rows == 1) {
                    std::array<RMatPP, 4> Xs;
                    for (int k = 0; k < 4; ++k) Xs[k] = take_pp(X, k);
                    prt = new PRT_1D(Xs);
                } else {
                    RMatXN4 Xf(prt_rows);
                    Xf.v = X;
                    prt = new PRT_3D(Xf);
                }
            }
            std::array<Point *, 25> pts;
            for (int i = 0; i < 25; ++i) pts[i] = domain->getPoint(tags[i]);
            if (kind == 0) {
                const int law = r.get<int32_t>(), rows = r.get<int32_t>();
                const int ncoef = law == AX3D_ISO ? 2 : law == AX3D_TI ? 5 : 21;
                std::vector<float> coef = r.vec<float>((size_t)ncoef * rows * 25);
                const int att_kind = r.get<int32_t>();
                Attenuation *att = 0;
                if (att_kind != AX3D_ATT_NONE) {
                    const int nsls = r.get<int32_t>(), dok = r.get<int32_t>();
                    RColX al = r.vec<float>(nsls), be = r.vec<float>(nsls), ga = r.vec<float>(nsls);
                    const int P = att_kind == AX3D_ATT_CG4 ? 4 : 25;
                    std::vector<float> dk = r.vec<float>((size_t)rows * P), dm = r.vec<float>((size_t)rows * P);
                    if (rows == 1 && P == 25) att = new Attenuation1D_Full(nsls, al, be, ga, pts[0]->getNu(), take_pp(dk, 0), take_pp(dm, 0), dok != 0);
                    else if (rows == 1) {
                        std::array<Real, 4> a{dk[0], dk[1], dk[2], dk[3]}, b{dm[0], dm[1], dm[2], dm[3]};
                        att = new Attenuation1D_CG4(nsls, al, be, ga, pts[0]->getNu(), a, b, dok != 0);
                    } else if (P == 25) att = new Attenuation3D_Full(nsls, al, be, ga, take_xn(dk, 0, rows), take_xn(dm, 0, rows), dok != 0);
                    else {
                        RMatX4 a(rows), b(rows);
                        a.v = dk;
                        b.v = dm;
                        att = new Attenuation3D_CG4(nsls, al, be, ga, a, b, dok != 0);
                    }
                }
                Elastic *el;
                if (rows == 1) {
                    if (law == AX3D_ISO) el = new Isotropic1D(take_pp(coef, 0), take_pp(coef, 1), att);
                    else if (law == AX3D_TI) el = new TransverselyIsotropic1D(take_pp(coef, 0), take_pp(coef, 1), take_pp(coef, 2), take_pp(coef, 3), take_pp(coef, 4), att);
                    else {
                        std::array<RMatPP, 21> C;
                        for (int k = 0; k < 21; ++k) C[k] = take_pp(coef, k);
                        el = new Anisotropic1D(C, att);
                    }
                } else {
                    if (law == AX3D_ISO) el = new Isotropic3D(take_xn(coef, 0, rows), take_xn(coef, 1, rows), att);
                    else if (law == AX3D_TI)
                        el = new TransverselyIsotropic3D(take_xn(coef, 0, rows), take_xn(coef, 1, rows), take_xn(coef, 2, rows), take_xn(coef, 3, rows), take_xn(coef, 4, rows), att);
                    else {
                        std::vector<RMatXN> C;
                        for (int k = 0; k < 21; ++k) C.push_back(take_xn(coef, k, rows));
                        el = new Anisotropic3D(C[0], C[1], C[2], C[3], C[4], C[5], C[6], C[7], C[8], C[9], C[10], C[11], C[12], C[13], C[14],
                                               C[15], C[16], C[17], C[18], C[19], C[20], att);   // the reference's signature
                    }
                }
                domain->addElement(new SolidElement(grad, prt, pts, el));
            } else {
                const int rows = r.get<int32_t>();
                std::vector<float> K = r.vec<float>((size_t)rows * 25);
                Acoustic *ac = rows == 1 ? (Acoustic *)new Acoustic1D(take_pp(K, 0)) : (Acoustic *)new Acoustic3D(take_xn(K, 0, rows));
                domain->addElement(new FluidElement(grad, prt, pts, ac));
            }
        }
        // ---- Source::release (Source.cpp:30-59)
        const int nsrc = r.get<int32_t>();
        for (int is = 0; is < nsrc; ++is) {
            const int etag = r.get<int32_t>();
            std::vector<int32_t> nrow = r.vec<int32_t>(25);
            arPP_CMatX3 force;
            for (int i = 0; i < 25; ++i) {
                force[i] = CMatX3(nrow[i]);
                std::vector<float> v = r.vec<float>((size_t)6 * nrow[i]);
                std::memcpy(static_cast<void *>(force[i].v.data()), v.data(), v.size() * sizeof(float));
            }
            domain->addSourceTerm(new SourceTerm(domain->getElement(etag), force));
        }
        // ---- STF::release (STF.cpp:9-12)
        const int nsteps = r.get<int32_t>();
        const double dt = r.get<double>();
        std::vector<float> stf = r.vec<float>(nsteps);
        domain->setSTF(new SourceTimeFunction(stf, dt, 0.));

        Newmark *newmark = new Newmark(domain, 1000000, 20, false);   // axisem.cpp:169
        newmark->solve(0);                                            // axisem.cpp:181

        // ---- read-back (what PointwiseRecorder / wisdom learning use: Point::getDispFourierSolid/Fluid)
        std::ofstream out(argv[2], std::ios::binary);
        for (int ip = 0; ip < domain->getNumPoints(); ++ip) {
            Point *p = domain->getPoint(ip);
            if (dynamic_cast<SolidPoint *>(p) || dynamic_cast<SolidFluidPoint *>(p)) {
                CMatX3 u = p->getDispFourierSolid();
                out.write(reinterpret_cast<const char *>(u.v.data()), u.v.size() * sizeof(Complex));
            }
            if (dynamic_cast<FluidPoint *>(p) || dynamic_cast<SolidFluidPoint *>(p)) {
                CColX u = p->getDispFourierFluid();
                out.write(reinterpret_cast<const char *>(u.data()), u.size() * sizeof(Complex));
            }
        }
        // one receiver through Element::computeGroundMotion on the last solid element
        delete newmark;
        delete domain;
        std::printf("host_driver ok: %d points, %d elements, %d steps\n", npoints, nelems, nsteps);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return std::strstr(e.what(), "no CUDA device") ? 3 : 1;
    }
}
