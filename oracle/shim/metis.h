// oracle/shim/metis.h -- TEST INFRASTRUCTURE, not product code.
//
// The METIS 5 interface DualGraph.cpp:35-144 is written against: 32-bit idx_t / real_t (DualGraph.cpp:138-144 refuses anything
// else).  The only METIS in this image is the one inside the CUDA toolkit (libmetis_static.a, no header, 64-bit idx_t, 32-bit
// real_t -- probed by axisem3d_b200/host/dual_graph.cpp's self-test), so the calls below are C++ overloads on int* that widen
// the arrays, call the library's C symbols and narrow the result.  Arrays handed back (xadj / adjncy of METIS_MeshToDual) are
// malloc'ed here and released by the METIS_Free overload.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <vector>
#define IDXTYPEWIDTH 32
#define REALTYPEWIDTH 32
typedef int32_t idx_t;
typedef float real_t;
#define METIS_NOPTIONS 40
#define METIS_OK 1
enum { METIS_OBJTYPE_CUT, METIS_OBJTYPE_VOL, METIS_OBJTYPE_NODE };
enum { METIS_OPTION_PTYPE, METIS_OPTION_OBJTYPE, METIS_OPTION_CTYPE, METIS_OPTION_IPTYPE, METIS_OPTION_RTYPE, METIS_OPTION_DBGLVL,
       METIS_OPTION_NITER, METIS_OPTION_NCUTS, METIS_OPTION_SEED, METIS_OPTION_NO2HOP, METIS_OPTION_MINCONN, METIS_OPTION_CONTIG,
       METIS_OPTION_COMPRESS, METIS_OPTION_CCORDER, METIS_OPTION_PFACTOR, METIS_OPTION_NSEPS, METIS_OPTION_UFACTOR,
       METIS_OPTION_NUMBERING };
namespace ax_metis64 {
extern "C" {
int METIS_SetDefaultOptions(int64_t *options);
int METIS_MeshToDual(int64_t *ne, int64_t *nn, int64_t *eptr, int64_t *eind, int64_t *ncommon, int64_t *numflag, int64_t **r_xadj,
                     int64_t **r_adjncy);
int METIS_PartGraphKway(int64_t *nvtxs, int64_t *ncon, int64_t *xadj, int64_t *adjncy, int64_t *vwgt, int64_t *vsize,
                        int64_t *adjwgt, int64_t *nparts, float *tpwgts, float *ubvec, int64_t *options, int64_t *edgecut,
                        int64_t *part);
int METIS_Free(void *ptr);
}
}  // namespace ax_metis64

inline int METIS_SetDefaultOptions(idx_t *options) {
    int64_t o[METIS_NOPTIONS];
    const int rc = ax_metis64::METIS_SetDefaultOptions(o);
    for (int i = 0; i < METIS_NOPTIONS; ++i) options[i] = (idx_t)o[i];
    return rc;
}
inline int METIS_MeshToDual(idx_t *ne, idx_t *nn, idx_t *eptr, idx_t *eind, idx_t *ncommon, idx_t *numflag, idx_t **r_xadj,
                            idx_t **r_adjncy) {
    int64_t ne64 = *ne, nn64 = *nn, nc64 = *ncommon, nf64 = *numflag, *xadj = nullptr, *adjncy = nullptr;
    std::vector<int64_t> p(eptr, eptr + *ne + 1), i(eind, eind + eptr[*ne]);
    const int rc = ax_metis64::METIS_MeshToDual(&ne64, &nn64, p.data(), i.data(), &nc64, &nf64, &xadj, &adjncy);
    if (rc != METIS_OK) return rc;
    *r_xadj = (idx_t *)std::malloc(sizeof(idx_t) * (*ne + 1));
    for (int k = 0; k <= *ne; ++k) (*r_xadj)[k] = (idx_t)xadj[k];
    *r_adjncy = (idx_t *)std::malloc(sizeof(idx_t) * (xadj[*ne] > 0 ? xadj[*ne] : 1));
    for (int64_t k = 0; k < xadj[*ne]; ++k) (*r_adjncy)[k] = (idx_t)adjncy[k];
    ax_metis64::METIS_Free(xadj);
    ax_metis64::METIS_Free(adjncy);
    return rc;
}
inline int METIS_PartGraphKway(idx_t *nvtxs, idx_t *ncon, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *vsize, idx_t *adjwgt,
                               idx_t *nparts, real_t *tpwgts, real_t *ubvec, idx_t *options, idx_t *edgecut, idx_t *part) {
    const int n = *nvtxs;
    int64_t n64 = n, ncon64 = *ncon, np64 = *nparts, cut64 = 0;
    std::vector<int64_t> xa(xadj, xadj + n + 1), ad(adjncy, adjncy + xadj[n]), vw, vs, aw, op, pt(n, 0);
    if (vwgt) vw.assign(vwgt, vwgt + (size_t)n * *ncon);
    if (vsize) vs.assign(vsize, vsize + n);
    if (adjwgt) aw.assign(adjwgt, adjwgt + xadj[n]);
    if (options) op.assign(options, options + METIS_NOPTIONS);
    const int rc = ax_metis64::METIS_PartGraphKway(&n64, &ncon64, xa.data(), ad.data(), vwgt ? vw.data() : nullptr,
                                                   vsize ? vs.data() : nullptr, adjwgt ? aw.data() : nullptr, &np64, tpwgts, ubvec,
                                                   options ? op.data() : nullptr, &cut64, pt.data());
    *edgecut = (idx_t)cut64;
    for (int k = 0; k < n; ++k) part[k] = (idx_t)pt[k];
    return rc;
}
inline int METIS_Free(idx_t *ptr) {
    std::free(ptr);
    return METIS_OK;
}
