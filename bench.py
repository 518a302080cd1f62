#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: GLL-point x Fourier-mode Newmark steps / second, % of the HBM roofline.

A "step" is one iteration of Newmark::solve's loop body (Newmark.cpp:47-93: updateNewmark, applySource,
computeStiff, coupleSolidFluid, assembleStiff) over the whole mesh.  The default workload is the configuration the
north star's target is quoted on, configs[3] = "cfg4": the 50 s-period mesh size (2016 quads, ~32.7 k GLL points), 3D
model, per-point Nu 20 ... 500 (ragged FFT sizes), fluid outer core with solid-fluid coupling -- on the synthetic
structured mesh of axisem3d_b200/mesh_synth.py.  cfg1 / cfg2 / cfg3 / cfg5 are selected with --config.

Weak scaling (default): at N > 1 the mesh grows with N (72 N x 28 quads) and is cut into N parts by METIS k-way on the
element dual graph (axisem3d_b200/partition.py; DualGraph.cpp:35-94), boundary-point stiffness summed over NVLink every
step.  --scaling strong keeps the N = 8 mesh (576 x 28 = 16128 quads) for every N.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one JSON line on rank 0)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: oracle/oracle.c on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GLL-point x Fourier-mode Newmark steps per second"
UNIT = "point-modes/s"
N_THETA, N_R, NU = 72, 28, 100
STRONG_N = 8            # --scaling strong: the mesh every N runs is the N = 8 weak-scaling mesh


def cfg4_nu(s, z):
    """configs[3]: per-point Nu 20 ... 500, growing with the distance from the axis (wisdom-style ragged expansion)."""
    return int(20 + 480 * min(1.0, s / 6371e3))


def cfg5_nu(s, z):
    """configs[4]: empirical Nu law in the spirit of EmpNrField (preloop/nrfield/EmpNrField.cpp:22-53): Nu grows with the
    distance from the axis (the wavefield needs ~ 2 pi s / lambda_min azimuthal samples), capped at 1000."""
    return int(min(1000, 24 + 1100 * (s / 6371e3) ** 0.85))


CONFIGS = {
    # name: (mesh keyword arguments, description)
    "cfg1": (dict(nu=2, law="ti", model3d=False, attenuation="cg4"),
             "cfg1: 1D TI PREM-like + CG4 attenuation, Nu=2 (Nr=5), no FFT"),
    "cfg2": (dict(nu=NU, law="iso", model3d=True, attenuation=None),
             "cfg2: 3D isotropic, Nu=%d (Nr=208), no attenuation, fluid core + SF coupling" % NU),
    "cfg3": (dict(nu=200, law="aniso", model3d=True, attenuation="cg4", fluid3d=False),
             "cfg3: 3D anisotropic (21 C_ij) + CG4 SLS attenuation, Nu=200 (Nr=416), SF coupling through the outer core"),
    "cfg4": (dict(nu_fn=cfg4_nu, law="iso", model3d=True, attenuation=None),
             "cfg4: 3D isotropic, per-point Nu 20..500 (Nr up to 672 on this mesh), ragged FFT sizes, fluid core + SF coupling"),
    # configs[4]: the large anisotropic + SLS mesh with Nu up to 1000, partitioned over 8 GPUs (bench.py --config cfg5 --gpus 8).
    # Element count used: 240 x 96 = 23 040 quads (11.4 x the 50 s mesh, ~370 k GLL points, 166 M point-modes per step,
    # 52 GB of fields + moduli + memory variables in total).  The ~200 k-quad mesh BASELINE.md sketches (720 x 280) would fit
    # the 8 x 180 GB as well (~470 GB), but the per-element Python set-up of the synthetic generator would take ~10 min per
    # rank on the GPU box; the per-element work, the Nr range and the partition shape are the same.
    "cfg5": (dict(nu_fn=cfg5_nu, law="aniso", model3d=True, attenuation="cg4", fluid3d=False),
             "cfg5: synthetic mesh of 23 040 quads, 3D anisotropic (21 C_ij) + CG4 SLS attenuation, empirical Nu <= 1000 (Nr <= 2016)"),
}
CFG = "cfg4"
CFG5_THETA, CFG5_R = 240, 96


def mesh_shape(world, scaling):
    if CFG == "cfg5":
        return CFG5_THETA, CFG5_R
    return N_THETA * (STRONG_N if scaling == "strong" else world), N_R


def _iso_work_nu(base_nu, base_fn):
    """Weak scaling keeps the work per GPU fixed: the N-GPU mesh refines the N = 1 mesh in theta, and the circumference cap of
    the Fourier order (Quad.cpp:562-566: Nr <= 2 pi s / GLL spacing) would loosen with the finer spacing and give the near-axis
    points of the larger meshes more modes.  The order field of the weak-scaling meshes is therefore the N = 1 field: the
    configured order capped with the N = 1 mesh's GLL spacing."""
    from axisem3d_b200.mesh_synth import R_EARTH
    dth1, dr = np.pi / N_THETA, (R_EARTH - 600e3) / N_R

    def fn(s, z):
        nu = base_fn(s, z) if base_fn is not None else base_nu
        spacing = 0.5 * (np.hypot(s, z) * dth1 + dr) / 4.0
        upper = max(int(2 * np.pi * s / spacing), 3)
        return min(int(nu), (upper - 1) // 2)
    return fn


MESH = "synth"          # --mesh exodus: the shipped 50 s Exodus mesh itself (tests/golden/AxiSEM_prem_ani_one_crust_50.e), N = 1
EXODUS_FILE = os.path.join(ROOT, "tests", "golden", "AxiSEM_prem_ani_one_crust_50.e")


def make_mesh(n_theta, n_r=None, **kw):
    if MESH == "exodus":
        from axisem3d_b200.exodus_mesh import ExodusMesh
        c = CONFIGS[CFG][0]
        # the elastic law comes from the file (PREM: transversely isotropic in the upper mantle, isotropic elsewhere)
        return ExodusMesh(EXODUS_FILE, nu=c.get("nu", 2), nu_fn=c.get("nu_fn"), attenuation=c.get("attenuation"),
                          model3d=c.get("model3d", False), fluid3d=c.get("fluid3d", False), dtype_coef=np.float32)
    from axisem3d_b200.mesh_synth import SynthMesh
    args = dict(n_theta=n_theta, n_r=n_r or N_R, dtype_coef=np.float32)
    args.update(CONFIGS[CFG][0])
    args.update(kw)
    if CFG != "cfg5" and n_theta != N_THETA and (n_r or N_R) == N_R:
        args["nu_fn"] = _iso_work_nu(args.get("nu", 2), args.get("nu_fn"))
    return SynthMesh(**args)


def stf_series(n):
    t = np.arange(n, dtype=np.float64)
    return np.exp(-((t - 40.0) / 12.0) ** 2).astype(np.float32)


def workload_name(n_theta, n_r=N_R):
    if MESH == "exodus":
        return "%s; on the shipped mesh template/input/AxiSEM_prem_ani_one_crust_50.e (2016 quads; elastic law from the file)" % CONFIGS[CFG][1]
    return "%s; synthetic meridional mesh %d x %d = %d quads" % (CONFIGS[CFG][1], n_theta, n_r, n_theta * n_r)


def element_weights(mesh):
    """Element cost for the partitioner.  The reference measures every element's computeStiff and repartitions with the
    times as METIS vertex weights (Mesh.cpp:412-588); here the weights come from the cost model calibrated on costs
    measured on the device (ax3d_measure_costs; axisem3d_b200/cost_model.json, profiles/scripts/calibrate_costs.py):
    a Nr log2 Nr + b Nr + c per element kind, plus the Newmark update of the element's share of points.  Without the
    model file: Nr log2 Nr for 3D solid elements, a tenth of it for the cheap kinds."""
    nr = mesh.e_nr.astype(np.float64)
    nlog = nr * np.log2(np.maximum(nr, 2.0))
    fluid = np.asarray(mesh.is_fluid, dtype=bool)
    skey = "solid|%s|%s|%s" % ("3d" if mesh.model3d else "1d", mesh.law, mesh.att_kind or "none")
    fkey = "fluid|%s" % ("3d" if mesh.fluid3d else "1d")
    try:
        model = json.load(open(os.path.join(ROOT, "axisem3d_b200", "cost_model.json")))
        w = np.zeros_like(nr)
        for key, mask in ((skey, ~fluid), (fkey, fluid)):
            c = model[key]
            fit = np.maximum(c["a_nr_log2nr"] * nlog[mask] + c["b_nr"] * nr[mask] + c["c"], 0.05)
            if "table_nr" in c:      # inside the measured range: the measured mean cost per Nr, interpolated
                tn, tu = np.asarray(c["table_nr"], dtype=np.float64), np.asarray(c["table_us"], dtype=np.float64)
                inside = (nr[mask] >= tn[0]) & (nr[mask] <= tn[-1])
                fit = np.where(inside, np.interp(nr[mask], tn, tu), fit)
            w[mask] = fit
        # Newmark: ~16 unique points per element, (Nu + 1) modes each, 3 components for solid points
        w += model["_point_us_per_mode"] * 16.0 * (nr / 2 + 1) * np.where(fluid, 1.0 / 3.0, 1.0)
        return w
    except (OSError, KeyError):
        w = nlog + 16.0
        cheap = fluid if not mesh.fluid3d else np.zeros_like(fluid)
        if not mesh.model3d:
            cheap = np.ones_like(fluid)
        return np.where(cheap, 0.1 * w, w)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_median": float(np.median(pw)) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_arm(steps, warmup, stride=9, min_seconds=0.0, max_seconds=25.0, n_theta=N_THETA):
    """Times oracle/oracle.c (C + OpenMP restatement on all host cores; the reference itself cannot be built here) on a
    bounded sample of the workload: every `stride`-th theta-column of the n_theta x 28 mesh (same Nr distribution;
    stride 1 = the whole mesh).  Runs `steps` steps, keeps going until `min_seconds` of CPU work, stops at `max_seconds`.
    The OpenMP thread count is set explicitly to the host's core count (torchrun exports OMP_NUM_THREADS=1)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from axisem_oracle import OracleDomain
    from c_oracle import COracle
    mesh = make_mesh(n_theta)
    dt = mesh.estimate_dt()
    e2p = np.where(mesh.ab[:, 0] % stride == stride // 2, 0, 1)
    d = OracleDomain(np.float32)
    rel = mesh.release(d, dt, rank=0, elem_to_proc=e2p)
    d.finalize()
    co = COracle(d)
    try:
        ncore = len(os.sched_getaffinity(0))
    except AttributeError:
        ncore = os.cpu_count() or 1
    threads = co.set_threads(ncore)
    rng = np.random.default_rng(1)
    for fld in (d.S, d.F):
        a = fld["displ"]
        a[:] = ((rng.standard_normal(a.shape) + 1j * rng.standard_normal(a.shape)) * 1e-6).astype(a.dtype)
    d.maskDispl()
    work = int(np.sum(d.p_nu + 1))
    stf = stf_series(4096)
    for i in range(warmup):
        co.step(dt, float(stf[i % 4096]))
    t0 = time.perf_counter()
    done = 0
    while True:
        co.step(dt, float(stf[(warmup + done) % 4096]))
        done += 1
        el = time.perf_counter() - t0
        if el > max_seconds or (done >= steps and el >= min_seconds):
            break
    return dict(value=work * done / el, unit=UNIT, cores=threads, kind="port",
                sample="%d of %d quads (%s of the %s mesh, %d x %d), %d point-modes per step, %d steps, %.1f s of "
                       "C/OpenMP oracle on %d threads (of %d host cores)" % (
                           len(rel["elements"]), mesh.nelem, "every theta-column" if stride == 1 else "every %d-th theta-column" % stride,
                           CFG, n_theta, N_R, work, done, el, threads, ncore),
                ms_per_step=1e3 * el / done, steps=done, elements=len(rel["elements"]), mesh_elements=int(mesh.nelem))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if CFG == "cfg5":
        print(json.dumps({"impl": "reference", "unavailable": "cfg5 (23040 quads, Nu <= 1000: ~20 s of CPU work per step) does not fit the CPU arm's time budget; use cfg1-cfg4"}))
        return
    n_theta, n_r = mesh_shape(args.gpus, args.scaling)
    # the same N x mesh the GPU arm steps at --gpus N, whole mesh per step, all host threads
    r = cpu_arm(args.steps, args.warmup, stride=1, max_seconds=150.0, n_theta=n_theta)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(n_theta), "elements": r["mesh_elements"],
                       "note": "whole mesh per step on the host cores (C/OpenMP restatement of the reference path; the reference binary cannot be built in this image)"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def global_receivers(mesh, nrec=128, seed=5):
    """128 receivers (the template STATIONS file has 128) in the outermost solid layer of the GLOBAL mesh: element id,
    azimuth, interpolation weights -- the same on every rank, so that an N-rank run and a 1-rank run record the same stations."""
    if hasattr(mesh, "surf_side"):
        surf = np.nonzero((mesh.surf_side >= 0) & ~mesh.is_fluid)[0]
    else:
        surf = np.nonzero((mesh.ab[:, 1] == mesh.nr_ - 1) & ~mesh.is_fluid)[0]
    rng = np.random.default_rng(seed)
    eg = surf[rng.integers(0, len(surf), nrec)]
    phi = rng.uniform(0, 2 * np.pi, nrec)
    w = rng.uniform(0, 1, (nrec, 25))
    w /= w.sum(axis=1, keepdims=True)
    return eg, phi, w


def apply_kick(dom, rel, amp=1e12):
    """A broadband force on every GLL point and Fourier mode, a function of the GLOBAL point tag only, so that a partitioned
    run and a single-domain run start from the same state and the whole wavefield -- across every partition boundary -- is
    in motion from the first step of the parity check (the first Newmark update masks and mass-scales it)."""
    l2g = rel["dec"].local_to_global_gll
    ks, kf = [], []
    for t, p in enumerate(rel["points"]):
        g = float(l2g[t])
        a = np.arange(p.nu + 1, dtype=np.float64)
        if p.kind != "fluid":
            for c in range(3):
                ks.append(amp * (np.sin(0.37 * g + 1.3 * c + 0.11 * a) + 1j * np.cos(0.23 * g + 0.7 * c + 0.05 * a)) / (1.0 + a))
        if p.kind != "solid":
            kf.append(amp * (np.cos(0.41 * g + 0.13 * a) + 1j * np.sin(0.29 * g + 0.07 * a)) / (1.0 + a))
    if ks:
        dom.set_bulk("stiff", False, np.concatenate(ks).astype(np.complex64))
    if kf:
        dom.set_bulk("stiff", True, np.concatenate(kf).astype(np.complex64))


def register_receivers(dom, rel, eg, phi, w):
    """registers the receivers that lie in this rank's elements; returns their indices in the global list"""
    loc = {int(g): il for il, g in enumerate(rel["dec"].local_elems)}
    mine = [k for k, g in enumerate(eg) if int(g) in loc]
    if mine:
        dom.setReceivers([rel["elements"][loc[int(eg[k])]].domain_tag for k in mine], phi[mine], w[mine])
    return mine


def run_ours(args):
    import torch
    from axisem3d_b200 import connectivity as CN
    from axisem3d_b200 import partition as PT
    from axisem3d_b200.domain import Domain, nccl_unique_id

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    # stdout carries exactly one JSON line: library banners (NCCL prints its version to stdout) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n_theta, n_r = mesh_shape(world, args.scaling)
    mesh = make_mesh(n_theta, n_r)
    dt = mesh.estimate_dt()
    dom = Domain(local)
    e2p, part_info = None, None
    if world > 1:
        if args.partition == "metis":
            e2p, part_info = PT.partition_kway(mesh.conn, element_weights(mesh), world, imbalance=0.01, ntrials=4)
        else:
            e2p = CN.partition_contiguous(element_weights(mesh), world)
            part_info = {"method": "contiguous weighted cuts of the element order"}
        part_info.update(PT.halo_stats(mesh.conn, e2p))
    rel = mesh.release(dom, dt, rank=rank, elem_to_proc=e2p)
    src = mesh.make_source(rel["elements"], rel["dec"], amp=1e18)
    if src is not None:
        dom.addSourceTerm(src)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        dom.setMessaging(rel["msg"], rank, world, bytes(uid.cpu().numpy().tolist()))
    dom.finalize()
    halo = "none"
    if world > 1:
        halo = "nccl send/recv (eager steps)"
        if os.environ.get("AX3D_HALO", "peer") == "peer":
            dom.connectHalo(rel["msg"], rank, dist)     # NVLink peer-memory windows; steps replay as CUDA graphs
            halo = "peer-memory windows over NVLink (%s + k_halo_wait_add inside the step graph)" % (
                "in-kernel put from the solid element kernel" if os.environ.get("AX3D_INKERNEL_PUT", "0") not in ("", "0") else "k_halo_put")

    eg, phi, w = global_receivers(mesh)
    mine = register_receivers(dom, rel, eg, phi, w)

    K, W = args.steps, max(args.warmup, 3)
    NPAR = 10            # steps of the N-rank vs 1-rank parity check
    stf = stf_series(NPAR + W + 2 * K + 64)
    work_local = dom.work_per_step()
    alg = dom.algorithmic_bytes()

    def barrier():
        if dist is not None:
            dist.barrier()
        dom.synchronize()
        torch.cuda.synchronize()

    # ---- parity of the partitioned run (N > 1): the first NPAR steps from rest, seismograms at the 128 global receivers,
    #      compared below with a single-domain run of the same global mesh on rank 0's GPU
    seis_par = None
    if CFG == "cfg5":
        args.no_parity = True            # a single domain holding the whole cfg5 mesh is a set-up of its own
    if world > 1 and not args.no_parity:
        apply_kick(dom, rel)
        if mine:
            seis_par = dom.runStepsRecord(dt, stf[:NPAR])
        else:
            dom.runSteps(dt, stf[:NPAR])
            seis_par = np.zeros((NPAR, 0, 3), np.float32)
        barrier()

    # ---- device-timed region: inputs resident in HBM.  K steps per region (CUDA events on the launching stream, max over
    #      ranks); the region is repeated until >= args.min_seconds of GPU time so that the clock sampler sees the load,
    #      the median region is reported
    dom.runSteps(dt, stf[NPAR:NPAR + W])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = dom.launch_count()
    regions = []
    t_start = time.perf_counter()
    while True:
        barrier()
        regions.append(dom.runStepsTimed(dt, stf[NPAR + W:NPAR + W + K]))
        barrier()
        go = (time.perf_counter() - t_start) < args.min_seconds and len(regions) < 400
        if dist is not None:
            flag = torch.tensor([1.0 if go else 0.0], device="cuda")
            dist.broadcast(flag, 0)
            go = bool(flag.item() > 0)
        if not go:
            break
    launches = (dom.launch_count() - l0) // len(regions)
    clocks = sampler.stop() if rank == 0 else None
    stable = dom.checkStability()
    regions = np.array(regions, dtype=np.float64)

    # ---- end-to-end through the C-ABI with host buffers: ax3d_run_steps_record = Newmark::solve with the pointwise
    #      recorder (Newmark.cpp:49-70 order: update, record, source, stiff, couple, assemble).  Per step the host source
    #      factor goes H2D through a pinned slot (8 bytes) and the receiver samples come back D2H (one copy per batch of
    #      BATCH steps = the recorder's dump interval; the call returns only when the samples are in host memory).
    BATCH = 25
    e2e_regions, seis_bytes_per_step = [], 0
    def steps_rec(series):
        """K steps through the recording entry point; a rank without receivers steps through ax3d_run_steps and synchronises
        (every rank must take the same number of steps: the halo exchange pairs them)"""
        if mine:
            return dom.runStepsRecord(dt, series)
        dom.runSteps(dt, series)
        dom.synchronize()
        return None

    steps_rec(stf[:BATCH])      # graph variants of the recording path, record ring sized for a batch
    steps_rec(stf[:3])
    barrier()
    t_e2e = time.perf_counter()
    while True:
        barrier()
        t0 = time.perf_counter()
        done = 0
        while done < K:
            nb = min(BATCH, K - done)
            s_ = stf[NPAR + W + K + done:NPAR + W + K + done + nb]
            seis = steps_rec(s_)                    # returns with the samples in host memory
            if seis is not None:
                seis_bytes_per_step = int(seis[0].nbytes)
            done += nb
        e2e_regions.append(time.perf_counter() - t0)         # per rank; the max over ranks is taken below
        go = (time.perf_counter() - t_e2e) < 0.5 * args.min_seconds and len(e2e_regions) < 100
        if dist is not None:
            flag = torch.tensor([1.0 if go else 0.0], device="cuda")
            dist.broadcast(flag, 0)
            go = bool(flag.item() > 0)
        if not go:
            break
    barrier()
    e2e_regions = np.array(e2e_regions, dtype=np.float64)

    # ---- per-family and per-kernel device times: CUDA events on the launching stream around each family / each hot kernel,
    #      steps issued through ax3d_run_steps (eager, not the graph), so the in-kernel Newmark is active exactly as above
    dom.enable_timers(True)
    dom.get_timers(reset=True)
    dom.kernel_stats(reset=True)
    nt = max(min(K, 12), 4)
    dom.runSteps(dt, stf[:nt])
    fam = dom.get_timers(reset=True) / nt          # ms: newmark, elements, sf+source, halo
    kst = dom.kernel_stats(reset=True)
    dom.enable_timers(False)

    if dist is not None:
        t = torch.tensor(np.concatenate([regions, e2e_regions, [float(work_local), float(launches), float(seis_bytes_per_step)]]),
                         dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        regions = tmax[:len(regions)].cpu().numpy()
        e2e_regions = tmax[len(regions):len(regions) + len(e2e_regions)].cpu().numpy()
        work = mesh.work_per_step()          # global GLL points (shared points counted once)
        launches = int(tsum[-2])
        seis_bytes_per_step = int(tsum[-1])
        alg_all = torch.tensor(alg, dtype=torch.float64, device="cuda")
        dist.all_reduce(alg_all, op=dist.ReduceOp.SUM)
        alg_job = alg_all.cpu().numpy()
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, seis_par))
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"rank": rank, "elements": len(rel["elements"]), "point_modes": int(work_local),
                                          "family_ms": [round(float(x), 4) for x in fam],
                                          "kernels_ms": {k: round(v[0] / nt, 4) for k, v in kst.items()},
                                          "neighbours": len(rel["msg"].mIProcComm)})
    else:
        work = work_local
        alg_job = alg
        gathered = None
        per_rank = None
    ms = float(np.median(regions))
    e2e_s = float(np.median(e2e_regions))

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- parity: the same global mesh as ONE domain on this GPU, same source and receivers, first NPAR steps from rest
    parity = None
    if gathered is not None and not args.no_parity:
        try:
            one = Domain(local)
            rel1 = mesh.release(one, dt)
            s1 = mesh.make_source(rel1["elements"], rel1["dec"], amp=1e18)
            if s1 is not None:
                one.addSourceTerm(s1)
            one.finalize()
            register_receivers(one, rel1, eg, phi, w)
            apply_kick(one, rel1)
            ref = one.runStepsRecord(dt, stf[:NPAR]).astype(np.float64)
            got = np.zeros_like(ref)
            seen = np.zeros(len(eg), dtype=bool)
            for idx, sp in gathered:
                if idx:
                    got[:, idx, :] = sp
                    seen[idx] = True
            den = float(np.linalg.norm(ref))
            parity = {"what": "seismograms at the %d receivers, first %d steps after a broadband kick on every point and mode (+ the source): %d-rank run vs one domain holding the whole mesh" % (len(eg), NPAR, world),
                      "rel_l2": float(np.linalg.norm(got - ref) / den) if den > 0 else None, "tolerance": 1e-4,
                      "receivers_recorded": int(seen.sum())}
            parity["ok"] = bool(parity["rel_l2"] is not None and parity["rel_l2"] <= 1e-4 and seen.all())
            del one
        except Exception as ex:
            parity = {"ok": False, "error": repr(ex)}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # dominant kernel = the entry with the largest summed device time over the timed eager steps (rank 0's GPU)
    el_ms = float(fam[1])
    kernels = {k: {"ms_per_step": v[0] / nt, "launches_per_step": v[1] / nt, "algorithmic_bytes_per_step": v[2] / nt,
                   "GBps": (v[2] / (v[0] * 1e-3) / 1e9) if v[0] > 0 else None} for k, v in kst.items()}
    dk_name = max(kst, key=lambda k: kst[k][0]) if kst else "none"
    dk_ms_total, dk_n, dk_bytes_total = kst.get(dk_name, (0.0, 0, 0.0))
    dk_ms = dk_ms_total / dk_n if dk_n else 0.0
    dk_bytes = dk_bytes_total / dk_n if dk_n else 0.0
    achieved = dk_bytes / (dk_ms * 1e-3) / 1e9 if dk_ms > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get("%s|%s" % (CFG, dk_name.split(" ")[0])) if world == 1 else None
        if ent:
            traffic = ent.get("dram_bytes_per_launch")
    except Exception:
        pass
    step_bytes = float(alg_job.sum())
    ws_mb = (4 * 8 * (dom.field_size(False) + dom.field_size(True)) + alg[1] * 0.2) / 1e6
    line = {
        "metric": METRIC, "value": work * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic" if MESH == "synth" else "template/input Exodus mesh (1D PREM from the file%s), synthetic source and receivers" % (
            " + synthetic azimuthal perturbation" if CONFIGS[CFG][0].get("model3d") else ""),
        "config": {"workload": workload_name(n_theta, n_r), "elements": int(mesh.nelem), "gll_points": int(mesh.ngll),
                   "point_modes_per_step": int(work), "parallelism": "dd%d" % world, "halo": halo, "partition": part_info,
                   "timing": "median of %d regions of %d steps (CUDA events, max over ranks); min %.4f max %.4f ms/step" % (
                       len(regions), K, regions.min() / K, regions.max() / K),
                   "l2": "working set %.0f MB of point fields + moduli per GPU %s the 126 MB L2; no explicit flush" % (
                       ws_mb, "exceeds" if ws_mb > 126 else "DOES NOT exceed"),
                   "stable": bool(stable)},
        "clocks": clocks,
        "e2e": {"value": work * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * world, "d2h_bytes_per_step": seis_bytes_per_step,
                "batch_steps": BATCH, "ms_per_step": 1e3 * e2e_s / K, "regions": int(len(e2e_regions))},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dk_name,
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel_ms_per_launch": dk_ms, "kernel_algorithmic_bytes_per_launch": dk_bytes,
                     "kernel_share_of_step": (dk_ms_total / nt) / float(fam.sum()) if fam.sum() > 0 else None,
                     "kernels": kernels,
                     "algorithmic_bytes_per_step": {"points": float(alg_job[0]), "elements": float(alg_job[1]), "halo": float(alg_job[2])},
                     "family_ms": {"newmark": float(fam[0]), "elements": el_ms, "sf_source": float(fam[2]), "halo": float(fam[3])},
                     "whole_step": {"achieved": step_bytes / (ms / K * 1e-3) / 1e9,
                                    "frac": step_bytes / (ms / K * 1e-3) / 1e9 / (peak * world)}},
    }
    if parity is not None:
        line["parity"] = parity
    if per_rank is not None:
        line["per_rank"] = per_rank           # eager-step (timers on) family times of every rank: the load balance of the partition
    if world == 1 and not args.no_cpu and CFG != "cfg5":
        try:
            c = cpu_arm(6, 1, stride=3, min_seconds=12.0, max_seconds=25.0)
            line["cpu_baseline"] = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:       # the CPU leg must never sink the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the N-rank vs 1-rank seismogram check (N > 1)")
    ap.add_argument("--config", default="cfg4", choices=sorted(CONFIGS),
                    help="BASELINE.json config (cfg4 = configs[3], the variable-Nu config the target is quoted on, is the headline)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 72 N x 28 quads on N GPUs; strong: the N = 8 mesh (576 x 28) on every N")
    ap.add_argument("--partition", default="metis", choices=["metis", "contiguous"])
    ap.add_argument("--min-seconds", type=float, default=2.0, help="the K-step region is repeated until this much time has passed")
    ap.add_argument("--mesh", default="synth", choices=["synth", "exodus"],
                    help="exodus: run the configuration on the shipped 50 s Exodus mesh (read with axisem3d_b200/h5lite.py; N = 1)")
    args = ap.parse_args()
    global CFG, MESH
    CFG = args.config
    MESH = args.mesh
    if MESH == "exodus" and (args.gpus > 1 or CFG == "cfg5"):
        ap.error("--mesh exodus runs cfg1-cfg4 on one GPU")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
