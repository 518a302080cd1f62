"""oracle/make_golden.py -- TEST INFRASTRUCTURE.  Generates tests/golden/ref_*.npz from the REFERENCE's own sources.

For every case below: build the seeded synthetic domain (axisem3d_b200.mesh_synth), serialise it (tests/dump_domain.py),
run oracle/_ref/ref_driver on it -- the reference's SolidElement/FluidElement/Gradient/FieldFFT/SolverFFTW_*/Elastic*/
Attenuation*/Point*/Mass*/SFCoupling*/SourceTerm classes compiled unmodified from /root/reference by oracle/Makefile.ref
against the Eigen/FFTW stand-ins of oracle/shim/ -- and store what it produced: every point's displacement and stiffness
after NSTEP Newmark steps that start from a broadband random stiffness "kick" (see ref_driver.cpp).  The fixtures carry
the reference's outputs only; inputs are re-created from the case parameters and seeds recorded in each file.

/root/reference does not exist on the GPU box, so this script runs in the build container only:
    make -C oracle -f Makefile.ref && python oracle/make_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

from axisem3d_b200.mesh_synth import SynthMesh  # noqa: E402
from dump_domain import DumpDomain  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
NSTEP = 3
KICK_SEED = 20260101
SOURCE_AMP = 1e9          # source-driven and kick-driven displacements both of order 1e-6


def ragged_nu(s, z):
    return int(2 + 30 * s / 6371e3)


CASES = {
    # cfg1: 1D TI + CG4 attenuation, Nu = 2 (Nr = 5), fluid core, SF coupling
    "cfg1_ti1d_cg4": dict(n_theta=4, n_r=8, nu=2, law="ti", model3d=False, attenuation="cg4"),
    "iso1d_full": dict(n_theta=3, n_r=8, nu=5, law="iso", model3d=False, attenuation="full"),
    "aniso1d": dict(n_theta=3, n_r=8, nu=4, law="aniso", model3d=False, attenuation=None),
    # cfg2: 3D isotropic, no attenuation
    "cfg2_iso3d": dict(n_theta=3, n_r=8, nu=52, law="iso", model3d=True, attenuation=None),
    "ti3d_cg4": dict(n_theta=3, n_r=8, nu=9, law="ti", model3d=True, attenuation="cg4"),
    # cfg3: 3D anisotropic + SLS, 3D fluid, (3D mass and 3D solid-fluid coupling with perturb_rho)
    "cfg3_aniso3d_cg4": dict(n_theta=3, n_r=8, nu=20, law="aniso", model3d=True, attenuation="cg4", fluid3d=True),
    "aniso3d_full_mass3d": dict(n_theta=3, n_r=8, nu=7, law="aniso", model3d=True, attenuation="full", fluid3d=True,
                                perturb_rho=True),
    # particle relabelling (9-component path): PRT_1D with 1D material, PRT_3D with 3D material, solid and fluid elements
    "prt1d_ti1d_cg4": dict(n_theta=3, n_r=8, nu=4, law="ti", model3d=False, attenuation="cg4", prt=True),
    "prt3d_aniso3d_full": dict(n_theta=3, n_r=8, nu=7, law="aniso", model3d=True, attenuation="full", fluid3d=True, prt=True),
    "prt3d_iso3d": dict(n_theta=3, n_r=8, nu=12, law="iso", model3d=True, attenuation=None, fluid3d=True, prt=True),
    # ocean-load masses on the surface points: MassOcean1D (1D model) and MassOcean3D (phi-dependent mass)
    "ocean1d_iso1d": dict(n_theta=3, n_r=8, nu=5, law="iso", model3d=False, attenuation=None, ocean=True),
    "ocean3d_iso3d_rho": dict(n_theta=3, n_r=8, nu=9, law="iso", model3d=True, attenuation=None, perturb_rho=True, ocean=True),
    # cfg4: ragged per-point Nu
    "cfg4_ragged": dict(n_theta=5, n_r=8, nu_fn=ragged_nu, law="iso", model3d=True, attenuation=None),
}


def case_mesh(name):
    return SynthMesh(**CASES[name])


def stf_samples(nstep=NSTEP):
    return np.exp(-((np.arange(nstep) - 1.0) / 1.5) ** 2).astype(np.float32)


def make_kick(points, dt, seed=KICK_SEED):
    """One complex64 buffer per point in Point::feedBuffer order (solid [3][Nu+1] comp-major, then fluid [Nu+1]), scaled
    so that the first updateNewmark (u = dt^2 M^-1 f) yields displacements of order 1e-6."""
    rng = np.random.default_rng(seed)
    out = []
    for p in points:
        n = p.nu + 1
        subs = [(p, 3)] if p.kind == "solid" else [(p, 1)] if p.kind == "fluid" else [(p.solid, 3), (p.fluid, 1)]
        for q, ncomp in subs:
            inv = float(np.mean(np.asarray(q.mass.invMass, dtype=np.float64)))
            scale = 1e-6 / (dt * dt * inv)
            out.append(((rng.standard_normal(ncomp * n) + 1j * rng.standard_normal(ncomp * n)) * scale).astype(np.complex64))
    return out


def dump_case(name, workdir):
    m = case_mesh(name)
    dt = m.estimate_dt()
    d = DumpDomain()
    rel = m.release(d, dt)
    d.addSourceTerm(m.make_source(rel["elements"], rel["dec"], amp=SOURCE_AMP))
    stf = stf_samples()
    path = os.path.join(workdir, name + ".bin")
    d.write(path, dt, stf)
    kick = make_kick(d.points, dt)
    kpath = os.path.join(workdir, name + ".kick")
    np.concatenate(kick).tofile(kpath)
    return m, dt, stf, d, path, kpath


RECV_SEED = 20260103
WISDOM_CUTOFF = 0.5       # Point::learnWisdom(cutoff): large, so that the white kick spectrum yields non-trivial orders


def make_receivers(elements, nper=4, seed=RECV_SEED):
    """nper receivers in solid and nper in fluid elements: (element tags, phi, weights[25]) -- arbitrary interpolation weights."""
    rng = np.random.default_rng(seed)
    sol = [e.domain_tag for e in elements if e.kind == "solid"]
    flu = [e.domain_tag for e in elements if e.kind == "fluid"]
    tags = list(rng.choice(sol, min(nper, len(sol)), replace=False)) + list(rng.choice(flu, min(nper, len(flu)), replace=False))
    phi = rng.uniform(0, 2 * np.pi, len(tags)).astype(np.float32)
    w = rng.uniform(0, 1, (len(tags), 25))
    w = (w / w.sum(axis=1, keepdims=True)).astype(np.float32)
    return np.array(tags, dtype=np.int32), phi, w


def write_receivers(path, tags, phi, w):
    import struct
    with open(path, "wb") as f:
        f.write(struct.pack("<i", len(tags)))
        for t, p, ww in zip(tags, phi, w):
            f.write(struct.pack("<if", int(t), float(p)))
            f.write(np.ascontiguousarray(ww, dtype=np.float32).tobytes())


def split_output(raw, points):
    """ref_driver out.bin -> (displ, stiff), each a flat complex64 array in point order."""
    n = sum((p.nu + 1) * {"solid": 3, "fluid": 1}.get(p.kind, 4) for p in points)
    assert raw.size == 2 * n, (raw.size, n)
    return raw[:n], raw[n:]


def main():
    if not os.path.exists(REF_DRIVER):
        raise SystemExit("build it first: make -C oracle -f Makefile.ref")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        for name in CASES:
            m, dt, stf, d, path, kpath = dump_case(name, tmp)
            out = os.path.join(tmp, name + ".out")
            rin, rout = os.path.join(tmp, name + ".rin"), os.path.join(tmp, name + ".rout")
            write_receivers(rin, *make_receivers(d.elements))
            wout = os.path.join(tmp, name + ".wout")
            r = subprocess.run([REF_DRIVER, path, out, kpath, rin, rout, wout, repr(WISDOM_CUTOFF)], capture_output=True, text=True,
                               timeout=3600)
            if r.returncode != 0:
                raise SystemExit("%s: ref_driver failed: %s" % (name, r.stderr))
            displ, stiff = split_output(np.fromfile(out, dtype=np.complex64), d.points)
            meta = dict(case=name, params={k: (v.__name__ if callable(v) else v) for k, v in CASES[name].items()}, nstep=NSTEP,
                        kick_seed=KICK_SEED, dt=dt, source="oracle/_ref/ref_driver (reference sources + oracle/shim)")
            raw_r = np.fromfile(rout, dtype=np.float32)
            nrec = raw_r.size // 12
            ground = raw_r[:3 * nrec].reshape(nrec, 3)
            strain_curl = raw_r[3 * nrec:].reshape(nrec, 9)      # 6 strains (RTZ, Voigt) + 3 curl; zeros where not applicable
            meta["recv_seed"] = RECV_SEED
            meta["wisdom_cutoff"] = WISDOM_CUTOFF
            nu_wisdom = np.fromfile(wout, dtype=np.int32)
            assert nu_wisdom.size == len(d.points)
            np.savez_compressed(os.path.join(GOLDEN_DIR, "ref_%s.npz" % name), displ=displ, stiff=stiff, ground=ground, strain_curl=strain_curl, nu_wisdom=nu_wisdom,
                                meta=np.array(json.dumps(meta)))
            print("%-24s %5d points %5d elements  |u| %.3e  |f| %.3e  %s" % (
                name, len(d.points), len(d.elements), np.abs(displ).max(), np.abs(stiff).max(), r.stdout.strip()))


if __name__ == "__main__":
    main()
