"""oracle/make_golden_seismograms.py -- TEST INFRASTRUCTURE.  tests/golden/seis_*.npz: 2000-step station seismograms from
the REFERENCE's own sources (oracle/_ref/ref_driver, see oracle/make_golden.py), the quantity of BASELINE.json's third
correctness check ("station seismograms must agree to relative L2 misfit <= 1e-4 after 2000 steps").

Each case: seeded synthetic domain, moment-tensor-like source with a Gaussian source time function, receivers in solid
surface elements and in fluid elements; Domain::record after the update of every step; every STRIDE-th sample is stored.
Runs in the build container only:  make -C oracle -f Makefile.ref && python oracle/make_golden_seismograms.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import make_golden as mg  # noqa: E402
from axisem3d_b200.mesh_synth import SynthMesh  # noqa: E402
from dump_domain import DumpDomain  # noqa: E402

NSTEP, STRIDE, AMP = 2000, 10, 1e18
CASES = {
    # cfg1: 1D TI PREM-like + CG4 attenuation, Nu = 2 (the reference's own runnable configuration)
    "cfg1_ti1d_cg4": dict(n_theta=6, n_r=6, nu=2, law="ti", model3d=False, attenuation="cg4"),
    # 3D isotropic + 3D fluid (cfg2 shape, small)
    "cfg2_iso3d": dict(n_theta=4, n_r=6, nu=6, law="iso", model3d=True, attenuation=None, fluid3d=True),
}


def stf_samples(n=NSTEP):
    return np.exp(-((np.arange(n) - 30.0) / 8.0) ** 2).astype(np.float32)


def receivers(elements):
    """4 receivers in the outermost solid elements and 2 in fluid elements, seeded azimuths and weights"""
    rng = np.random.default_rng(mg.RECV_SEED + 1)
    sol = [e.domain_tag for e in elements if e.kind == "solid"][-4:]
    flu = [e.domain_tag for e in elements if e.kind == "fluid"][:2]
    tags = np.array(sol + flu, dtype=np.int32)
    phi = rng.uniform(0, 2 * np.pi, len(tags)).astype(np.float32)
    w = rng.uniform(0, 1, (len(tags), 25))
    return tags, phi, (w / w.sum(axis=1, keepdims=True)).astype(np.float32)


def build(name, domain):
    m = SynthMesh(**CASES[name])
    dt = m.estimate_dt()
    rel = m.release(domain, dt)
    domain.addSourceTerm(m.make_source(rel["elements"], rel["dec"], amp=AMP))
    return m, dt, rel


def main():
    os.makedirs(mg.GOLDEN_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        for name in CASES:
            d = DumpDomain()
            m, dt, rel = build(name, d)
            path, out, rin, rout, ser = (os.path.join(tmp, name + e) for e in (".bin", ".out", ".rin", ".rout", ".ser"))
            d.write(path, dt, stf_samples())
            tags, phi, w = receivers(d.elements)
            mg.write_receivers(rin, tags, phi, w)
            r = subprocess.run([mg.REF_DRIVER, path, out, "-", rin, rout, "-", "0", ser], capture_output=True, text=True, timeout=7200)
            if r.returncode != 0:
                raise SystemExit("%s: ref_driver failed: %s" % (name, r.stderr))
            series = np.fromfile(ser, dtype=np.float32).reshape(NSTEP, len(tags), 3)
            meta = dict(case=name, params=CASES[name], nstep=NSTEP, stride=STRIDE, dt=dt, amp=AMP,
                        source="oracle/_ref/ref_driver (reference sources + oracle/shim)")
            np.savez_compressed(os.path.join(mg.GOLDEN_DIR, "seis_%s.npz" % name), series=series[::STRIDE].copy(),
                                meta=np.array(json.dumps(meta)))
            print("%-16s %d receivers, max |u| solid %.3e fluid %.3e  %s" % (
                name, len(tags), np.abs(series[:, :4]).max(), np.abs(series[:, 4:]).max(), r.stdout.strip()))


if __name__ == "__main__":
    main()
