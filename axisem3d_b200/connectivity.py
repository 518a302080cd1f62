"""Element -> GLL-point numbering and partition halos (host-side integer work).

Restates S/preloop/graph/Connectivity.cpp: `formElemToGLL` (42-94), `decompose` (96-221),
`get_shared_DOF_quad` / `common_nodes` / `formNodeEdge` (223-316).  The reference obtains the
element neighbourhood from METIS_MeshToDual with ncommon=1 (DualGraph.cpp:96-136); here it is
built directly from the node->element incidence (same set; order is irrelevant because a
shared point receives the same tag from any lower-numbered neighbour).

The partition vector `elemToProc` itself is an input: the reference derives it from METIS
with wall-clock-measured weights (Mesh.cpp:412-588), which is not reproducible;
`partition_contiguous` below is a deterministic stand-in (weighted contiguous cuts of the
element order).  Everything downstream of `elemToProc` (local numbering, per-neighbour
shared-point lists in global-tag order) must be bit-exact and is tested as such.
"""
from __future__ import annotations

import numpy as np

nPol = 4
nPntEdge = 5

# Connectivity::formNodeEdge (Connectivity.cpp:265-309)
NODE_IJ = [(0, 0), (nPol, 0), (nPol, nPol), (0, nPol)]
EDGE_IJ = [
    [(i, 0) for i in range(nPntEdge)],
    [(nPol, i) for i in range(nPntEdge)],
    [(nPol - i, nPol) for i in range(nPntEdge)],
    [(0, nPol - i) for i in range(nPntEdge)],
]


def on_edge(ipol, jpol):
    return not (0 < ipol < nPol and 0 < jpol < nPol)


def form_neighbourhood(conn):
    """Elements sharing >= 1 node (DualGraph::formNeighbourhood with ncommon = 1)."""
    conn = np.asarray(conn, dtype=np.int64)
    nelem = conn.shape[0]
    node2el = {}
    for e in range(nelem):
        for n in conn[e]:
            node2el.setdefault(int(n), []).append(e)
    nb = []
    for e in range(nelem):
        s = set()
        for n in conn[e]:
            s.update(node2el[int(n)])
        s.discard(e)
        nb.append(sorted(s))
    return nb


def common_nodes(a, b):
    """Connectivity::common_nodes (236-263) -> (ncommon, aindex, bindex)."""
    ac, bc = [], []
    for i in range(4):
        for j in range(4):
            if a[i] == b[j]:
                ac.append(i)
                bc.append(j)
                break
    n = len(ac)
    if n == 1:
        return 1, ac[0], bc[0]
    if n == 2:
        ai, bi = min(ac), min(bc)
        if ai == 0 and max(ac) == 3:
            ai = 3
        if bi == 0 and max(bc) == 3:
            bi = 3
        return 2, ai, bi
    raise RuntimeError("Connectivity::common_nodes || Error topology in mesh.")


def shared_dof_quad(c1, c2):
    """Connectivity::get_shared_DOF_quad (223-234)."""
    n, i1, i2 = common_nodes(c1, c2)
    if n == 1:
        return [NODE_IJ[i1]], [NODE_IJ[i2]]
    return list(EDGE_IJ[i1]), list(reversed(EDGE_IJ[i2]))


def form_elem_to_gll(conn, neighbours=None):
    """Connectivity::formElemToGLL (42-94).  Returns (ngll, elemToGLL[nelem,5,5])."""
    conn = np.asarray(conn, dtype=np.int64)
    nelem = conn.shape[0]
    if neighbours is None:
        neighbours = form_neighbourhood(conn)
    e2g = -np.ones((nelem, nPntEdge, nPntEdge), dtype=np.int64)
    ngll = 0
    for ie in range(nelem):
        mask = np.ones((nPntEdge, nPntEdge), dtype=bool)
        for inb in neighbours[ie]:
            if inb >= ie:
                continue
            m1, m2 = shared_dof_quad(conn[inb], conn[ie])
            for (i1, j1), (i2, j2) in zip(m1, m2):
                mask[i2, j2] = False
                e2g[ie, i2, j2] = e2g[inb, i1, j1]
        for ip in range(nPntEdge):
            for jp in range(nPntEdge):
                if mask[ip, jp]:
                    e2g[ie, ip, jp] = ngll
                    ngll += 1
    return ngll, e2g


class Decomposition:
    pass


def decompose(conn, elem_to_proc, rank, global_e2g=None, neighbours=None):
    """Connectivity::decompose (96-221) for one rank, given elemToProc.

    Returns an object with: local_elems (global ids, ascending), nGllLocal, elemToGllLocal
    [nloc,5,5], iProcComm (ascending neighbour ranks), iLocalPoints (per neighbour: local GLL
    tags in ascending *global* GLL tag order), local_to_global_gll."""
    conn = np.asarray(conn, dtype=np.int64)
    elem_to_proc = np.asarray(elem_to_proc, dtype=np.int64)
    nelem = conn.shape[0]
    if neighbours is None:
        neighbours = form_neighbourhood(conn)
    if global_e2g is None:
        _, global_e2g = form_elem_to_gll(conn, neighbours)
    comm = {}                                            # rankOther -> {globalGll: (ielem, ip, jp)}
    edge_ij = [(i, j) for i in range(nPntEdge) for j in range(nPntEdge) if on_edge(i, j)]
    for ie in range(nelem):
        if elem_to_proc[ie] != rank:
            continue
        for inb in neighbours[ie]:
            other = int(elem_to_proc[inb])
            if other == rank:
                continue
            d = comm.setdefault(other, {})
            gll_other = set(int(global_e2g[inb, i, j]) for i, j in edge_ij)
            nfound = 0
            for i, j in edge_ij:
                t = int(global_e2g[ie, i, j])
                if t in gll_other:
                    d.setdefault(t, (ie, i, j))          # std::map::insert keeps the first
                    nfound += 1
            if nfound != nPntEdge and nfound != 1:
                raise RuntimeError("Connectivity::decompose || Domain decomposition failed.")
    local_elems = np.nonzero(elem_to_proc == rank)[0]
    glb2loc = -np.ones(nelem, dtype=np.int64)
    glb2loc[local_elems] = np.arange(len(local_elems))
    ngl, e2g_loc = form_elem_to_gll(conn[local_elems])
    out = Decomposition()
    out.local_elems = local_elems
    out.nGllLocal = ngl
    out.elemToGllLocal = e2g_loc
    out.iProcComm, out.iLocalPoints, out.iGlobalPoints = [], [], []
    for other in sorted(comm):
        out.iProcComm.append(other)
        loc, glb = [], []
        for t in sorted(comm[other]):
            ie, i, j = comm[other][t]
            loc.append(int(e2g_loc[glb2loc[ie], i, j]))
            glb.append(t)
        out.iLocalPoints.append(loc)
        out.iGlobalPoints.append(glb)
    l2g = -np.ones(ngl, dtype=np.int64)
    l2g[e2g_loc.reshape(-1)] = global_e2g[local_elems].reshape(-1)
    out.local_to_global_gll = l2g
    return out


def partition_contiguous(weights, nproc):
    """Deterministic stand-in for METIS k-way (DualGraph.cpp:35-94): cut the element order into
    `nproc` contiguous chunks of (nearly) equal total weight."""
    w = np.asarray(weights, dtype=np.float64)
    c = np.cumsum(w)
    tot = c[-1]
    proc = np.minimum((c - 0.5 * w) * nproc / tot, nproc - 1).astype(np.int64)
    return proc
