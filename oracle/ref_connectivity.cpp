// oracle/ref_connectivity.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Runs the REFERENCE's Connectivity::decompose (S/preloop/graph/Connectivity.cpp:96-221, compiled unmodified where it
// lies) for every rank of a given partition and writes what Mesh::buildLocal would receive: the local element -> GLL
// numbering (the gather/scatter index maps) and the per-neighbour shared-point lists (the halos).
// Not the reference: Eigen and MPI (oracle/shim), and DualGraph below.  The reference's DualGraph.cpp is two METIS calls
// (METIS 5.1.0 is not in the image): METIS_MeshToDual -- the elements sharing >= ncommon nodes, restated here by direct
// incidence (Connectivity.cpp does not depend on the adjacency order: a shared point gets the same tag from any
// lower-numbered neighbour, and the halo maps are keyed std::maps) -- and METIS_PartGraphKway, whose output elemToProc is
// this harness's INPUT (the reference's own partition depends on wall-clock cost measurements, Mesh.cpp:412-588).
//
//   usage: ref_connectivity <in.bin> <out.bin>
//   in : int32 nelem, nproc; int32 conn[nelem][4]; int32 elemToProc[nelem]
//   out: per rank: int32 nElemLocal, nGllLocal; int32 procMask[nelem]; int32 elemToGll[nElemLocal][5][5] (ipol-major);
//        int32 nProcComm; per neighbour: int32 rank, n; int32 localPoints[n]
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <set>
#include <stdexcept>
#include <vector>

#include "Connectivity.h"
#include "DualGraph.h"
#include "XMPI.h"

int ax_mpi_rank = 0, ax_mpi_size = 1;
static std::vector<int32_t> g_elemToProc;

void DualGraph::formNeighbourhood(const IMatX4 &connectivity, int ncommon, std::vector<IColX> &neighbours) {
    const int nelem = connectivity.rows();
    int nnode = 0;
    for (int e = 0; e < nelem; ++e)
        for (int k = 0; k < 4; ++k) nnode = std::max(nnode, connectivity(e, k) + 1);
    std::vector<std::vector<int>> node2el(nnode);
    for (int e = 0; e < nelem; ++e)
        for (int k = 0; k < 4; ++k) node2el[connectivity(e, k)].push_back(e);
    neighbours.clear();
    for (int e = 0; e < nelem; ++e) {
        std::set<int> cand;
        for (int k = 0; k < 4; ++k)
            for (int o : node2el[connectivity(e, k)])
                if (o != e) cand.insert(o);
        std::vector<int> keep;
        for (int o : cand) {
            int shared = 0;
            for (int a = 0; a < 4; ++a)
                for (int b = 0; b < 4; ++b) shared += connectivity(e, a) == connectivity(o, b);
            if (shared >= ncommon) keep.push_back(o);
        }
        IColX col((int)keep.size());
        for (size_t i = 0; i < keep.size(); ++i) col((int)i) = keep[i];
        neighbours.push_back(col);
    }
}

void DualGraph::decompose(const IMatX4 &connectivity, const DecomposeOption &, IColX &elemToProc) {
    elemToProc = IColX(connectivity.rows());
    for (int e = 0; e < connectivity.rows(); ++e) elemToProc(e) = g_elemToProc[e];
}

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: ref_connectivity in.bin out.bin\n"); return 2; }
    try {
        std::ifstream f(argv[1], std::ios::binary);
        if (!f) throw std::runtime_error("ref_connectivity || cannot open input");
        int32_t nelem, nproc;
        f.read(reinterpret_cast<char *>(&nelem), 4);
        f.read(reinterpret_cast<char *>(&nproc), 4);
        std::vector<int32_t> conn((size_t)nelem * 4);
        f.read(reinterpret_cast<char *>(conn.data()), conn.size() * 4);
        g_elemToProc.resize(nelem);
        f.read(reinterpret_cast<char *>(g_elemToProc.data()), (size_t)nelem * 4);
        IMatX4 excon(nelem, 4);
        for (int e = 0; e < nelem; ++e)
            for (int k = 0; k < 4; ++k) excon(e, k) = conn[(size_t)e * 4 + k];
        Connectivity con(excon);
        std::ofstream out(argv[2], std::ios::binary);
        auto put = [&out](int32_t v) { out.write(reinterpret_cast<const char *>(&v), 4); };
        ax_mpi_size = nproc;
        for (int r = 0; r < nproc; ++r) {
            ax_mpi_rank = r;
            int nGllLocal = 0;
            std::vector<IMatPP> e2g;
            MessagingInfo msg;
            IColX procMask;
            con.decompose(DecomposeOption(), nGllLocal, e2g, msg, procMask);
            put((int32_t)e2g.size());
            put(nGllLocal);
            for (int e = 0; e < nelem; ++e) put(procMask(e));
            for (const IMatPP &m : e2g)
                for (int i = 0; i <= nPol; ++i)
                    for (int j = 0; j <= nPol; ++j) put(m(i, j));
            put(msg.mNProcComm);
            for (int i = 0; i < msg.mNProcComm; ++i) {
                put(msg.mIProcComm[i]);
                put(msg.mNLocalPoints[i]);
                for (int t : msg.mILocalPoints[i]) put(t);
            }
        }
        std::printf("ref_connectivity ok: %d elements, %d ranks\n", nelem, nproc);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
}
