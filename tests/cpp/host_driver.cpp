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

// matrix adapters of this side: raw floats / doubles -> the plain containers ax3d_host.hpp stands in for the Eigen typedefs
static RMatXN mk_xn(const std::vector<float> &src, size_t k, int rows) {
    RMatXN m(rows);
    std::memcpy(m.v.data(), src.data() + k * (size_t)rows * 25, (size_t)rows * 25 * sizeof(float));
    return m;
}
static RMatPP mk_pp(const std::vector<float> &src, size_t k) {
    RMatPP m;
    std::memcpy(m.data(), src.data() + k * 25, 25 * sizeof(float));
    return m;
}
static RDMatPP mk_dpp(const double *p) { RDMatPP m; std::memcpy(m.data(), p, 25 * 8); return m; }
static RDCol2 mk_crds(double s, double z) { return RDCol2{s, z}; }
static RColX mk_col(const std::vector<float> &v) { return v; }
static RDColX mk_dcol(const std::vector<double> &v) { return v; }
static RDMatX3 mk_dx3(const std::vector<double> &v, int) { return v; }
static RMatX3 mk_x3(const std::vector<float> &v, int rows) { RMatX3 m(rows); m.v = v; return m; }
static RMatX4 mk_x4(const std::vector<float> &v, int rows) { RMatX4 m(rows); m.v = v; return m; }
static RMatXN4 mk_xn4(const std::vector<float> &v, int rows) { RMatXN4 m(rows); m.v = v; return m; }
static RRow4 mk_row4(const float *p) { return RRow4{p[0], p[1], p[2], p[3]}; }
static CMatX3 mk_cx3(const std::vector<float> &v, int nrow) {
    CMatX3 m(nrow);
    std::memcpy(static_cast<void *>(m.v.data()), v.data(), v.size() * sizeof(float));
    return m;
}

#include "release_domain.inc"

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: host_driver dump.bin out.bin\n"); return 2; }
    try {
        Reader r(argv[1]);
        Domain *domain = new Domain(0);                         // axisem.cpp:127
        ReleaseInfo info;
        release_domain(r, domain, info);                        // Mesh::release, Source::release (axisem.cpp:133-145)
        const int npoints = info.npoints, nelems = info.nelems, nsteps = info.nsteps;
        domain->setSTF(new SourceTimeFunction(info.stf, info.dt, 0.));   // STF::release (STF.cpp:9-12)

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
