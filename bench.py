#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on the cfg2 workload: GLL-point x Fourier-mode Newmark steps / second.

A "step" is one iteration of Newmark::solve's loop body (Newmark.cpp:47-93: updateNewmark, applySource,
computeStiff, coupleSolidFluid[, assembleStiff]) over the whole mesh.  Workload at N = 1 = configs[1]:
the 50 s-period mesh size (2016 quads, ~32.7 k GLL points), 3D isotropic elastic model, constant Nu = 100
(Nr = 208, capped by the circumference near the axis), no attenuation, fluid outer core with solid-fluid
coupling -- on the synthetic structured mesh of axisem3d_b200/mesh_synth.py (the Exodus reader is out of the
hot-path scope).  At N > 1 the mesh grows with N (72 N x 28 quads, weak scaling) and is cut into N
contiguous parts with an NCCL halo sum per step.

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


def cfg4_nu(s, z):
    """configs[3]: per-point Nu 20 ... 500, growing with the distance from the axis (wisdom-style ragged expansion)."""
    return int(20 + 480 * min(1.0, s / 6371e3))


CONFIGS = {
    # name: (mesh keyword arguments, description)
    "cfg1": (dict(nu=2, law="ti", model3d=False, attenuation="cg4"),
             "cfg1: 1D TI PREM-like + CG4 attenuation, Nu=2 (Nr=5), no FFT"),
    "cfg2": (dict(nu=NU, law="iso", model3d=True, attenuation=None),
             "cfg2: 3D isotropic, Nu=%d (Nr=208), no attenuation, fluid core + SF coupling" % NU),
    "cfg3": (dict(nu=200, law="aniso", model3d=True, attenuation="cg4", fluid3d=False),
             "cfg3: 3D anisotropic (21 C_ij) + CG4 SLS attenuation, Nu=200 (Nr=416), SF coupling through the outer core"),
    "cfg4": (dict(nu_fn=cfg4_nu, law="iso", model3d=True, attenuation=None),
             "cfg4: 3D isotropic, per-point Nu 20..500 (Nr 42..1008), ragged FFT sizes"),
}
CFG = "cfg2"


def make_mesh(n_theta, **kw):
    from axisem3d_b200.mesh_synth import SynthMesh
    args = dict(n_theta=n_theta, n_r=N_R, dtype_coef=np.float32)
    args.update(CONFIGS[CFG][0])
    args.update(kw)
    return SynthMesh(**args)


def stf_series(n):
    t = np.arange(n, dtype=np.float64)
    return np.exp(-((t - 40.0) / 12.0) ** 2).astype(np.float32)


def workload_name(n_theta):
    return "%s; 50 s-mesh-size synthetic meridional mesh (%d x %d = %d quads)" % (CONFIGS[CFG][1], n_theta, N_R, n_theta * N_R)


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
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_arm(steps, warmup, stride=9, min_seconds=0.0, max_seconds=25.0):
    """Times oracle/oracle.c (C + OpenMP restatement on all host cores; the reference itself cannot be built here) on a
    bounded sample of the cfg2 workload: every `stride`-th theta-column of the 72 x 28 mesh (same Nr distribution;
    stride 1 = the whole mesh).  Runs `steps` steps, keeps going until `min_seconds` of CPU work, stops at `max_seconds`."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from axisem_oracle import OracleDomain
    from c_oracle import COracle
    mesh = make_mesh(N_THETA)
    dt = mesh.estimate_dt()
    e2p = np.where(mesh.ab[:, 0] % stride == stride // 2, 0, 1)
    d = OracleDomain(np.float32)
    rel = mesh.release(d, dt, rank=0, elem_to_proc=e2p)
    d.finalize()
    co = COracle(d)
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
    return dict(value=work * done / el, unit=UNIT, cores=co.threads(), kind="port",
                sample="%d of %d quads (%s of the %s mesh), %d point-modes per step, %d steps, %.1f s of "
                       "C/OpenMP oracle on %d threads" % (len(rel["elements"]), mesh.nelem,
                                                          "every theta-column" if stride == 1 else "every %d-th theta-column" % stride,
                                                          CFG, work, done, el, co.threads()),
                ms_per_step=1e3 * el / done, steps=done)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_arm(args.steps, args.warmup, stride=1, max_seconds=150.0)   # whole cfg2 mesh per step
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"],
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(N_THETA), "note": "whole cfg2 mesh per step on the host cores (C/OpenMP restatement of the reference path; the reference binary cannot be built in this image)"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    from axisem3d_b200 import connectivity as CN
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

    n_theta = N_THETA * world
    mesh = make_mesh(n_theta)
    dt = mesh.estimate_dt()
    dom = Domain(local)
    e2p = None
    if world > 1:
        e2p = CN.partition_contiguous(mesh.e_nr.astype(np.float64), world)
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
            halo = "peer-memory windows over NVLink (k_halo_put / k_halo_wait_add), steps replayed as CUDA graphs"

    # 128 receivers (the template STATIONS file has 128) in the outermost solid layer
    surf = [e.domain_tag for e in rel["elements"] if e.kind == "solid"]
    rng = np.random.default_rng(5)
    nrec = 128
    etags = [surf[i] for i in rng.integers(0, len(surf), nrec)]
    w = rng.uniform(0, 1, (nrec, 25))
    w /= w.sum(axis=1, keepdims=True)
    dom.setReceivers(etags, rng.uniform(0, 2 * np.pi, nrec), w)

    K, W = args.steps, max(args.warmup, 3)
    stf = stf_series(W + 2 * K + 64)
    work_local = dom.work_per_step()
    alg = dom.algorithmic_bytes()

    def barrier():
        if dist is not None:
            dist.barrier()
        dom.synchronize()
        torch.cuda.synchronize()

    # ---- device-timed region: inputs resident in HBM
    dom.runSteps(dt, stf[:W])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = dom.launch_count()
    barrier()
    ms = dom.runStepsTimed(dt, stf[W:W + K])
    barrier()
    launches = dom.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    stable = dom.checkStability()

    # ---- end-to-end through the C-ABI with host buffers: ax3d_run_steps_record = Newmark::solve with the pointwise
    #      recorder (Newmark.cpp:49-70 order: update, record, source, stiff, couple, assemble).  Per step the host source
    #      factor goes H2D through a pinned slot (8 bytes) and the 128 receiver samples come back D2H (one copy per batch of
    #      BATCH steps = the recorder's dump interval; the call returns only when the samples are in host memory).
    BATCH = 25
    dom.runStepsRecord(dt, stf[:BATCH])      # graph variants of the recording path, record ring sized for a batch
    dom.runStepsRecord(dt, stf[:3])
    barrier()
    t0 = time.perf_counter()
    done = 0
    while done < K:
        nb = min(BATCH, K - done)
        seis = dom.runStepsRecord(dt, stf[W + K + done:W + K + done + nb])   # returns with the samples in host memory
        done += nb
    e2e_s = time.perf_counter() - t0         # per rank; the max over ranks is taken below
    barrier()
    seis_bytes_per_step = int(seis[0].nbytes)

    # ---- per-family device times and the dominant kernel's launch time: CUDA events on the launching stream around each
    #      family / around the solid k_elem3d_fused launch, steps issued through ax3d_run_steps (eager, not the graph), so
    #      the in-kernel Newmark is active exactly as in the timed region above
    dom.enable_timers(True)
    dom.get_timers(reset=True)
    dom.dominant_kernel(reset=True)
    nt = max(min(K, 12), 4)
    dom.runSteps(dt, stf[:nt])
    fam = dom.get_timers(reset=True) / nt          # ms: newmark, elements, sf+source, halo
    dk_ms, dk_bytes = dom.dominant_kernel(reset=True)
    dom.enable_timers(False)

    if dist is not None:
        t = torch.tensor([ms, e2e_s, float(work_local), float(launches)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, e2e_s = float(tmax[0]), float(tmax[1])
        work = mesh.work_per_step()          # global GLL points (shared points counted once)
        launches = int(tsum[3])
    else:
        work = work_local

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    traffic = None
    try:
        if CFG == "cfg2" and world == 1:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_elem3d_fused_dram_bytes_per_launch_cfg2")
    except Exception:
        pass
    el_ms = float(fam[1])
    if dk_ms > 0:
        dk_name = "k_elem3d_fused<solid> (gather+grad+c2r+stress+r2c+quad+scatter of the 3D solid elements" + \
                  (" + in-kernel Newmark of the plain solid points)" if dk_bytes[1] > 0 else ")")
        dk_total = float(dk_bytes.sum())
        achieved = dk_total / (dk_ms * 1e-3) / 1e9
    else:   # no fused launch in this configuration (e.g. cfg1: all elements 1D): fall back to the element family
        dk_name = "element stiffness family (k_elem1d)"
        dk_total = float(alg[1])
        achieved = alg[1] / (el_ms * 1e-3) / 1e9 if el_ms > 0 else 0.0
        dk_ms = el_ms
    step_bytes = float(alg.sum())
    line = {
        "metric": METRIC, "value": work * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(n_theta), "elements": int(mesh.nelem), "gll_points": int(mesh.ngll),
                   "point_modes_per_step": int(work), "parallelism": "dd%d" % world, "halo": halo,
                   "l2": "working set %.0f MB of point fields + moduli per GPU exceeds the 126 MB L2; no explicit flush" %
                         ((4 * 8 * (dom.field_size(False) + dom.field_size(True)) + alg[1] * 0.2) / 1e6),
                   "stable": bool(stable)},
        "clocks": clocks,
        "e2e": {"value": work * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": seis_bytes_per_step, "batch_steps": BATCH,
                "ms_per_step": 1e3 * e2e_s / K},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dk_name,
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel_ms_per_launch": dk_ms, "kernel_algorithmic_bytes_per_launch": dk_total,
                     "kernel_bytes_split": {"elements": float(dk_bytes[0]), "points_in_kernel": float(dk_bytes[1])},
                     "algorithmic_bytes_per_step": {"points": alg[0], "elements": alg[1], "halo": alg[2]},
                     "family_ms": {"newmark": float(fam[0]), "elements": el_ms, "sf_source": float(fam[2]), "halo": float(fam[3])},
                     "whole_step": {"achieved": step_bytes / (ms / K * 1e-3) / 1e9, "frac": step_bytes / (ms / K * 1e-3) / 1e9 / peak}},
    }
    if world == 1 and not args.no_cpu:
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
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS),
                    help="BASELINE.json config (cfg2 = configs[1] is the headline; the others are extra measurements)")
    args = ap.parse_args()
    global CFG
    CFG = args.config
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
