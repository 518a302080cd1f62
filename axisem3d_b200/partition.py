"""Element partition across GPUs: METIS k-way on the element dual graph, as the reference does it.

Front end of `axisem3d_b200/host/dual_graph.cpp` (DualGraph::decompose / formNeighbourhood,
S/preloop/graph/DualGraph.cpp:12-94) over its C-ABI.  The METIS underneath is the 64-bit-index build
inside the CUDA toolkit (`libmetis_static.a`); the library is linked in-tree (`libax3d_partition.so`).
`elemToProc` feeds `connectivity.decompose` (local numbering + halo lists, bit-exact vs Connectivity.cpp).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libax3d_partition.so")
SRC = os.path.join(HERE, "host", "dual_graph.cpp")
_lib = None


def _find_metis():
    for pat in ("/usr/local/cuda/targets/*/lib/libmetis_static.a", "/usr/local/cuda-*/targets/*/lib/libmetis_static.a",
                "/usr/lib/x86_64-linux-gnu/libmetis.a"):
        hits = sorted(glob.glob(pat))
        if hits:
            return hits[0]
    raise RuntimeError("axisem3d_b200.partition: libmetis_static.a not found in the CUDA toolkit")


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", SRC, _find_metis(), "-lm", "-o", LIB])
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.ax3d_partition_last_error.restype = C.c_char_p
        p64, pd = C.POINTER(C.c_int64), C.POINTER(C.c_double)
        lib.ax3d_partition_kway.argtypes = [C.c_int64, p64, pd, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, p64, p64, pd,
                                            C.POINTER(C.c_int)]
        lib.ax3d_dual_graph.argtypes = [C.c_int64, p64, C.c_int, p64, p64, C.c_int64]
        if lib.ax3d_metis_selftest() != 0:
            raise RuntimeError(lib.ax3d_partition_last_error().decode())
        _lib = lib
    return _lib


def _p64(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def dual_graph(conn, ncommon):
    """DualGraph::formNeighbourhood (DualGraph.cpp:12-33): list of neighbour arrays per element."""
    lib = load()
    conn = np.ascontiguousarray(conn, dtype=np.int64)
    ne = conn.shape[0]
    xadj = np.zeros(ne + 1, dtype=np.int64)
    if lib.ax3d_dual_graph(ne, _p64(conn), int(ncommon), _p64(xadj), None, 0):
        raise RuntimeError(lib.ax3d_partition_last_error().decode())
    adj = np.zeros(int(xadj[-1]), dtype=np.int64)
    if lib.ax3d_dual_graph(ne, _p64(conn), int(ncommon), _p64(xadj), _p64(adj), adj.size):
        raise RuntimeError(lib.ax3d_partition_last_error().decode())
    return [adj[xadj[e]:xadj[e + 1]] for e in range(ne)]


def partition_kway(conn, weights, nproc, imbalance=0.01, ncuts=1, seed=0, ntrials=4):
    """DualGraph::decompose (DualGraph.cpp:35-94).  Returns (elem_to_proc int64[nelem], info) with info =
    {"edgecut", "imbalance" (max part weight / mean), "contiguous", "method"}."""
    lib = load()
    conn = np.ascontiguousarray(conn, dtype=np.int64)
    ne = conn.shape[0]
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    out = np.zeros(ne, dtype=np.int64)
    cut, imb, contig = C.c_int64(0), C.c_double(1.0), C.c_int(1)
    rc = lib.ax3d_partition_kway(ne, _p64(conn), None if w is None else w.ctypes.data_as(C.POINTER(C.c_double)), int(nproc),
                                 float(imbalance), int(ncuts), int(seed), int(ntrials), _p64(out), C.byref(cut), C.byref(imb),
                                 C.byref(contig))
    if rc:
        raise RuntimeError(lib.ax3d_partition_last_error().decode())
    return out, {"edgecut": int(cut.value), "imbalance": float(imb.value), "contiguous": bool(contig.value),
                 "method": "METIS_PartGraphKway on the ncommon=2 dual graph (contiguous parts, ufactor %.2f, best of %d seeds)" % (imbalance, ntrials)}


def halo_stats(conn, elem_to_proc):
    """cut edges and neighbour count per rank of a partition (ncommon = 1 contacts, i.e. including corners)."""
    nb = dual_graph(conn, 1)
    e2p = np.asarray(elem_to_proc)
    nproc = int(e2p.max()) + 1
    neigh = [set() for _ in range(nproc)]
    for e, lst in enumerate(nb):
        for f in lst:
            if e2p[f] != e2p[e]:
                neigh[e2p[e]].add(int(e2p[f]))
    return {"neighbours_per_rank": [len(s) for s in neigh]}
