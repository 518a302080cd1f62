// dual_graph.cpp -- host side of the multi-GPU path: the element partition.
//
// B200-side counterpart of DualGraph::decompose / formNeighbourhood (SOLVER/src/preloop/graph/DualGraph.cpp:12-94):
// the quadrilateral mesh becomes its dual graph (two elements are adjacent when they share `ncommon` nodes: 2 = an edge for
// the partition, 1 = a corner for the neighbour discovery of Connectivity::decompose), and METIS's multilevel k-way
// partitioner cuts it into `nproc` contiguous parts of equal total weight (vertex weight = element cost, Mesh.cpp:88-101,
// 578) minimising the edge cut -- one part per GPU.  METIS itself is the 64-bit-index build that ships inside the CUDA
// toolkit (libmetis_static.a, a cuSOLVER dependency); no header comes with it, so the five entry points used are declared
// here with that build's types (idx_t = int64, real_t = float; checked at run time by ax3d_metis_selftest).
//
// C-ABI (plain pointers and sizes), bound from Python by axisem3d_b200/partition.py and from C++ by ax3d_host.hpp.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <queue>
#include <string>
#include <vector>

typedef int64_t idx_t;
typedef float real_t;
extern "C" {
int METIS_SetDefaultOptions(idx_t *options);
int METIS_MeshToDual(idx_t *ne, idx_t *nn, idx_t *eptr, idx_t *eind, idx_t *ncommon, idx_t *numflag, idx_t **r_xadj, idx_t **r_adjncy);
int METIS_PartGraphKway(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *vsize, idx_t *adjwgt, idx_t *nparts,
                        real_t *tpwgts, real_t *ubvec, idx_t *options, idx_t *edgecut, idx_t *part);
int METIS_Free(void *ptr);
}
#define AX_METIS_OK 1
#define AX_METIS_NOPTIONS 40
// option slots of METIS 5 (moptions_et).  5.1 inserted NO2HOP at 9, which moved MINCONN / CONTIG from 9 / 10 to 10 / 11: both
// candidate slots of CONTIG are set -- the other one is MINCONN (5.1) or COMPRESS (5.0, ordering only), harmless for k-way.
enum { OPT_OBJTYPE = 1, OPT_NCUTS = 7, OPT_SEED = 8, OPT_CONTIG_50 = 10, OPT_CONTIG_51 = 11 };

static thread_local std::string g_err;
extern "C" const char *ax3d_partition_last_error(void) { return g_err.c_str(); }

static bool mesh_to_dual(int64_t nelem, const int64_t *conn, int ncommon, idx_t **xadj, idx_t **adjncy) {
    // unique node list -> compact node ids (DualGraph.cpp:96-117)
    std::vector<int64_t> nodes(conn, conn + 4 * nelem);
    std::sort(nodes.begin(), nodes.end());
    nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
    idx_t ne = nelem, nn = (idx_t)nodes.size(), nc = ncommon, numflag = 0;
    std::vector<idx_t> eptr(nelem + 1), eind(4 * nelem);
    for (int64_t e = 0; e <= nelem; ++e) eptr[e] = 4 * e;
    for (int64_t k = 0; k < 4 * nelem; ++k) eind[k] = std::lower_bound(nodes.begin(), nodes.end(), conn[k]) - nodes.begin();
    if (METIS_MeshToDual(&ne, &nn, eptr.data(), eind.data(), &nc, &numflag, xadj, adjncy) != AX_METIS_OK) {
        g_err = "DualGraph::metisError || Error in metis function: METIS_MeshToDual";
        return false;
    }
    return true;
}

extern "C" {

// DualGraph::formNeighbourhood: CSR adjacency of the dual graph.  Call with adjncy == NULL to size it (xadj[nelem] entries).
int ax3d_dual_graph(int64_t nelem, const int64_t *conn, int ncommon, int64_t *xadj_out, int64_t *adjncy_out, int64_t adjncy_cap) {
    idx_t *xadj = nullptr, *adjncy = nullptr;
    if (!mesh_to_dual(nelem, conn, ncommon, &xadj, &adjncy)) return 1;
    for (int64_t e = 0; e <= nelem; ++e) xadj_out[e] = xadj[e];
    int rc = 0;
    if (adjncy_out) {
        if (adjncy_cap < xadj[nelem]) { g_err = "DualGraph::formNeighbourhood || adjacency buffer too small"; rc = 1; }
        else for (int64_t k = 0; k < xadj[nelem]; ++k) adjncy_out[k] = adjncy[k];
    }
    METIS_Free(xadj);
    METIS_Free(adjncy);
    return rc;
}

// DualGraph::decompose (DualGraph.cpp:35-94): weights may be NULL (unit weights); `ntrials` partitions are made with seeds
// seed0 .. seed0 + ntrials - 1 and the one with the smallest edge cut wins (the reference runs one trial per MPI rank and
// broadcasts the best).  Out: elem_to_proc[nelem], *edgecut, *imbalance_out = max part weight / mean part weight,
// *contiguous = 1 when every part is connected in the ncommon = 2 dual graph.
int ax3d_partition_kway(int64_t nelem, const int64_t *conn, const double *weights, int nproc, double imbalance, int ncuts, int seed0,
                        int ntrials, int64_t *elem_to_proc, int64_t *edgecut, double *imbalance_out, int *contiguous) {
    if (nelem <= 0 || nproc <= 0) { g_err = "DualGraph::decompose || empty mesh or no processors"; return 1; }
    for (int64_t e = 0; e < nelem; ++e) elem_to_proc[e] = 0;
    if (edgecut) *edgecut = 0;
    if (imbalance_out) *imbalance_out = 1.0;
    if (contiguous) *contiguous = 1;
    if (nproc == 1) return 0;
    idx_t *xadj = nullptr, *adjncy = nullptr;
    if (!mesh_to_dual(nelem, conn, 2, &xadj, &adjncy)) return 1;
    // integer vertex weights scaled to 0.9 * INT32_MAX in total, as the reference does (DualGraph.cpp:70-79)
    std::vector<idx_t> vwgt;
    if (weights) {
        double sum = 0;
        for (int64_t e = 0; e < nelem; ++e) sum += weights[e];
        if (!(sum > 0)) { g_err = "DualGraph::decompose || element weights must be positive"; METIS_Free(xadj); METIS_Free(adjncy); return 1; }
        const double imax = 0.9 * (double)std::numeric_limits<int32_t>::max();
        vwgt.resize(nelem);
        for (int64_t e = 0; e < nelem; ++e) vwgt[e] = std::max<idx_t>(1, (idx_t)std::llround(weights[e] / sum * imax));
    }
    std::vector<idx_t> part(nelem), best(nelem);
    idx_t best_cut = std::numeric_limits<idx_t>::max();
    int rc = 0;
    for (int t = 0; t < std::max(1, ntrials); ++t) {
        idx_t opt[AX_METIS_NOPTIONS];
        METIS_SetDefaultOptions(opt);
        opt[OPT_OBJTYPE] = 0;          // METIS_OBJTYPE_CUT
        opt[OPT_CONTIG_50] = 1;
        opt[OPT_CONTIG_51] = 1;
        opt[OPT_NCUTS] = std::max(1, ncuts);
        opt[OPT_SEED] = seed0 + t;
        idx_t nv = nelem, ncon = 1, np = nproc, cut = 0;
        real_t ub = (real_t)(1.0 + imbalance);
        if (METIS_PartGraphKway(&nv, &ncon, xadj, adjncy, weights ? vwgt.data() : nullptr, nullptr, nullptr, &np, nullptr, &ub, opt, &cut,
                                part.data()) != AX_METIS_OK) {
            g_err = "DualGraph::metisError || Error in metis function: METIS_PartGraphKway";
            rc = 1;
            break;
        }
        if (cut < best_cut) { best_cut = cut; best = part; }
    }
    if (!rc) {
        for (int64_t e = 0; e < nelem; ++e) elem_to_proc[e] = best[e];
        if (edgecut) *edgecut = best_cut;
        std::vector<double> pw(nproc, 0.0);
        double tot = 0;
        for (int64_t e = 0; e < nelem; ++e) { const double w = weights ? weights[e] : 1.0; pw[best[e]] += w; tot += w; }
        if (imbalance_out) *imbalance_out = *std::max_element(pw.begin(), pw.end()) / (tot / nproc);
        if (contiguous) {   // every part connected?
            std::vector<char> seen(nelem, 0);
            std::vector<int> comps(nproc, 0);
            for (int64_t s = 0; s < nelem; ++s) {
                if (seen[s]) continue;
                comps[best[s]]++;
                std::queue<int64_t> q;
                q.push(s);
                seen[s] = 1;
                while (!q.empty()) {
                    const int64_t u = q.front();
                    q.pop();
                    for (idx_t k = xadj[u]; k < xadj[u + 1]; ++k) {
                        const idx_t v = adjncy[k];
                        if (!seen[v] && best[v] == best[u]) { seen[v] = 1; q.push(v); }
                    }
                }
            }
            *contiguous = 1;
            for (int p = 0; p < nproc; ++p) if (comps[p] != 1) *contiguous = 0;
        }
    }
    METIS_Free(xadj);
    METIS_Free(adjncy);
    return rc;
}

// the METIS build really has 64-bit indices: partition a 6 x 6 grid in two and look at the answer
int ax3d_metis_selftest(void) {
    const int n = 6;
    std::vector<int64_t> conn;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const int64_t a = i * (n + 1) + j;
            conn.insert(conn.end(), {a, a + n + 1, a + n + 2, a + 1});
        }
    std::vector<int64_t> part(n * n);
    int64_t cut = -1;
    double imb = 0;
    int contig = 0;
    if (ax3d_partition_kway(n * n, conn.data(), nullptr, 2, 0.01, 1, 0, 1, part.data(), &cut, &imb, &contig)) return 1;
    int64_t c0 = 0;
    for (int64_t p : part) {
        if (p != 0 && p != 1) { g_err = "DualGraph::check_idx_t || Incompatible METIS build (index width)"; return 1; }
        c0 += p == 0;
    }
    if (c0 != 18 || cut < n || cut > 2 * n || !contig) { g_err = "DualGraph::check_idx_t || METIS self test failed"; return 1; }
    return 0;
}

}   // extern "C"
