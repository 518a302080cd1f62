"""oracle/make_golden_connectivity.py -- TEST INFRASTRUCTURE.  tests/golden/conn_*.npz from the REFERENCE's own
Connectivity::decompose (S/preloop/graph/Connectivity.cpp, compiled unmodified into oracle/_ref/ref_connectivity by
oracle/Makefile.ref; METIS's two calls are stood in for as described in oracle/ref_connectivity.cpp).

Each fixture stores the inputs (quad connectivity, elemToProc) and, for every rank, what the reference returned: procMask,
nGllLocal, the local element -> GLL map [nloc][5][5], neighbour ranks and per-neighbour local point lists.
Runs in the build container only (needs /root/reference):  make -C oracle -f Makefile.ref && python oracle/make_golden_connectivity.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from axisem3d_b200 import connectivity as CN  # noqa: E402
from axisem3d_b200.mesh_synth import SynthMesh  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_connectivity")


def cases():
    """name -> (conn [nelem,4], elemToProc [nelem], nproc)"""
    out = {}
    m = SynthMesh(n_theta=8, n_r=6, nu=3)
    for n in (1, 2, 3, 4, 8):
        out["synth8x6_contig%d" % n] = (m.conn, CN.partition_contiguous(m.e_nr.astype(np.float64), n), n)
    # 2-D blocks: ranks that touch in a single corner point (the nfound == 1 branch, Connectivity.cpp:168-171)
    a, b = m.ab[:, 0], m.ab[:, 1]
    out["synth8x6_blocks4"] = (m.conn, (a >= 4).astype(np.int64) * 2 + (b >= 3).astype(np.int64), 4)
    # scrambled element order and rotated node orderings (every common_nodes / edge-reversal case), seeded
    rng = np.random.default_rng(20260102)
    m2 = SynthMesh(n_theta=7, n_r=5, nu=3)
    perm = rng.permutation(m2.nelem)
    conn = np.array([np.roll(m2.conn[e], int(rng.integers(0, 4))) for e in perm])
    ab = m2.ab[perm]
    out["scrambled7x5_blocks6"] = (conn, (ab[:, 0] * 3 // 7) * 2 + (ab[:, 1] >= 2).astype(np.int64), 6)
    out["scrambled7x5_random3"] = (conn, _connected_random(conn, 3, rng), 3)
    return out


def _connected_random(conn, nproc, rng):
    """region-growing partition from random seeds: irregular, but every cross-rank contact is an edge or a corner."""
    nb = CN.form_neighbourhood(conn)
    nelem = len(conn)
    part = -np.ones(nelem, dtype=np.int64)
    seeds = rng.choice(nelem, nproc, replace=False)
    front = [[int(s)] for s in seeds]
    for r, s in enumerate(seeds):
        part[s] = r
    while (part < 0).any():
        for r in range(nproc):
            nxt = []
            for e in front[r]:
                for o in nb[e]:
                    if part[o] < 0:
                        part[o] = r
                        nxt.append(o)
            front[r] = nxt or front[r]
    return part


def run_reference(conn, e2p, nproc, tmp):
    inp, outp = os.path.join(tmp, "c.in"), os.path.join(tmp, "c.out")
    with open(inp, "wb") as f:
        f.write(np.array([len(conn), nproc], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(conn, dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(e2p, dtype=np.int32).tobytes())
    r = subprocess.run([EXE, inp, outp], capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    raw = np.fromfile(outp, dtype=np.int32)
    pos, res = 0, {}
    nelem = len(conn)
    for rank in range(nproc):
        nloc, ngll = int(raw[pos]), int(raw[pos + 1]); pos += 2
        res["r%d_mask" % rank] = raw[pos:pos + nelem].copy(); pos += nelem
        res["r%d_e2g" % rank] = raw[pos:pos + nloc * 25].reshape(nloc, 5, 5).copy(); pos += nloc * 25
        res["r%d_ngll" % rank] = np.int32(ngll)
        ncomm = int(raw[pos]); pos += 1
        ranks, lists = [], []
        for _ in range(ncomm):
            other, n = int(raw[pos]), int(raw[pos + 1]); pos += 2
            ranks.append(other)
            lists.append(raw[pos:pos + n].copy()); pos += n
        res["r%d_comm" % rank] = np.array(ranks, dtype=np.int32)
        for other, l in zip(ranks, lists):
            res["r%d_to%d" % (rank, other)] = l
    assert pos == raw.size
    return res


def main():
    if not os.path.exists(EXE):
        raise SystemExit("build it first: make -C oracle -f Makefile.ref")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        for name, (conn, e2p, nproc) in cases().items():
            res = run_reference(conn, e2p, nproc, tmp)
            np.savez_compressed(os.path.join(GOLDEN_DIR, "conn_%s.npz" % name), conn=np.asarray(conn, np.int32),
                                elem_to_proc=np.asarray(e2p, np.int32), nproc=np.int32(nproc), **res)
            print("%-28s %4d elements %d ranks, halo sizes %s" % (
                name, len(conn), nproc, [int(sum(len(res["r%d_to%d" % (r, o)]) for o in res["r%d_comm" % r])) for r in range(nproc)]))


if __name__ == "__main__":
    main()
