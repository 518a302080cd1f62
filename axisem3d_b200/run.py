"""`python -m axisem3d_b200.run <run_dir>`: what `./axisem3d` does in a run directory (S/axisem.cpp:12-176), on the CUDA path.

Reads `<run_dir>/input/` unchanged -- inparam.model, inparam.nu, inparam.time_src_recv, inparam.advanced, the Exodus mesh they
name, CMTSOLUTION (or the point-force file) and STATIONS -- builds the domain through the preloop restatement
(exodus_mesh.py, preloop.py, volumetric.py), runs the Newmark loop on the GPU (ax3d_run_steps_record: one CUDA graph per step,
device-side recorder) and writes `<run_dir>/output/stations/<network>.<name>.<RTZ|ENZ|SPZ>.ascii` in the layout of the
reference's PointwiseIOAscii (S/core/output/pointwise/PointwiseIOAscii.cpp: time and three components per line) and / or
`axisem3d_synthetics.nc` with the variables of PointwiseIONetCDF, as OUT_STATIONS_FORMAT asks.

Covered: 1-D background models, the volumetric models of volumetric.py, constant / empirical / wisdom Nu, wisdom learning, CG4 / full attenuation,
earthquake / point-force sources, a constant ocean load, erf / gauss / ricker source-time functions, geographic / source-centred stations,
ellipticity mode off / geographic / full (particle relabelling).  Anything else in the input files fails loudly (NotImplementedError) rather than being
ignored.  There is no CPU fallback: the CUDA library must load and a device must be present.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

from . import preloop as PL
from . import relabelling as REL
from . import volumetric as VOL
from .exodus_mesh import ExodusMesh


class Simulation:
    """Everything axisem_main builds before the time loop, from the input directory alone; `release(domain)` plays
    Mesh::release, Source::release and ReceiverCollection::release into any object with the Domain verbs."""

    def __init__(self, input_dir):
        par = self.par = PL.Parameters(input_dir)
        if par.get("MODEL_PLOT_SLICES_NUM") != "0":
            raise NotImplementedError("axisem3d_b200.run: MODEL_PLOT_SLICES_NUM != 0 (slice plots are the reference's)")
        if par.get("MODEL_2D_MODE").lower() != "off":
            raise NotImplementedError("axisem3d_b200.run: MODEL_2D_MODE " + par.get("MODEL_2D_MODE"))
        ocean = [x for x in par.get("MODEL_3D_OCEAN_LOAD").split("$") if x != ""]         # OceanLoad3D::buildInparam
        if ocean[0].lower() == "none":
            ocean_depth = 0.0
        elif ocean[0].lower() == "constant":
            ocean_depth = (float(ocean[1]) if len(ocean) > 1 else 3.0) * 1e3              # OceanLoad3D_const.h: default 3 km
        else:
            raise NotImplementedError("axisem3d_b200.run: MODEL_3D_OCEAN_LOAD " + ocean[0] + " (needs the reference's data files)")
        if par.get("ATTENUATION_SPECFEM_LEGACY", bool):
            raise NotImplementedError("axisem3d_b200.run: ATTENUATION_SPECFEM_LEGACY (AttSimplex is Fortran in the reference)")
        if par.get("OUT_STATIONS_WHOLE_SURFACE", bool):
            raise NotImplementedError("axisem3d_b200.run: whole-surface output is driven through the C-ABI verbs, not this runner")
        self.learn = PL.learn_parameters(par)           # (invoked, cutoff, interval, file): Domain::setLearnParameters
        nu, nu_fn, lucky = PL.nu_field(par)
        att = None if not par.get("ATTENUATION", bool) else ("cg4" if par.get("ATTENUATION_CG4", bool) else "full")
        self.source = PL.Source.from_parameters(par)

        def models(mesh):                       # Volumetric3D / Geometric3D::buildInparam, once the mesh file's globals are known
            g = PL.Geodesy.from_mesh(mesh, par)
            return VOL.from_parameters(par, self.source, g), self.source, g, REL.from_parameters(par, g)
        self.mesh = ExodusMesh(os.path.join(input_dir, par.get("MODEL_1D_EXODUS_MESH_FILE")), nu=nu if nu is not None else 0,
                               nu_fn=nu_fn, lucky=lucky, attenuation=att, do_kappa=par.get("ATTENUATION_QKAPPA", bool),
                               volumetric=models, ocean_depth=ocean_depth)
        self.geodesy = self.mesh.vol_geodesy
        self.dt = PL.delta_t(par, self.mesh)
        self.stf, self.shift = PL.stf_from_parameters(par, self.dt)
        self.receivers = PL.Receivers.from_parameters(par, self.source, self.geodesy).locate(
            self.mesh, par.get("OUT_STATIONS_DEPTH_REF", bool))

    def release(self, domain, rank=0, elem_to_proc=None):
        """Mesh::release + Source::release for rank `rank` of the partition `elem_to_proc` (None = one rank holds everything)"""
        rel = self.mesh.release(domain, self.dt, rank=rank, elem_to_proc=elem_to_proc)
        if self.source is not None:
            st = self.source.release(self.mesh, self.geodesy, rel["elements"], None if elem_to_proc is None else rel["dec"])
            if st is not None:
                domain.addSourceTerm(st)
        return rel

    def partition(self, nproc):
        """The reference's first decomposition pass: METIS k-way on the element dual graph, vertex weight = the element's Nr
        (Mesh.cpp:88-101, DualGraph.cpp:35-94); deterministic, so every rank computes the same vector."""
        from . import partition as PT
        e2p, info = PT.partition_kway(self.mesh.conn, self.mesh.e_nr.astype(np.float64), nproc, imbalance=0.01, ntrials=4)
        return e2p, info

    def times(self):
        """the time stamp of every recorded sample (Newmark.cpp:27, 64-65: t starts at -shift, recorded before it advances)"""
        return -self.shift + self.dt * np.arange(len(self.stf))


def write_ascii(out_dir, sim, series):
    """PointwiseIOAscii: one file per station, `time c1 c2 c3` per recorded step."""
    rc = sim.receivers
    os.makedirs(out_dir, exist_ok=True)
    t = sim.times()[::rc.record_interval]
    for i, key in enumerate(rc.keys):
        with open(os.path.join(out_dir, "%s.%s.ascii" % (key, rc.components)), "w") as f:
            for k in range(len(t)):
                f.write("%.6g %.6g %.6g %.6g\n" % (t[k], series[k, i, 0], series[k, i, 1], series[k, i, 2]))


def write_netcdf(path, sim, series):
    """PointwiseIONetCDF (S/core/output/pointwise/PointwiseIONetCDF.cpp:83-260): `time_points` [nstep] (double) and one variable
    `<network>.<name>.<components>` [nstep][3] (float) per station with its latitude / longitude / depth attributes, plus the
    source location as global attributes.  Written in the classic NetCDF format (scipy); the reference writes NetCDF-4."""
    from scipy.io import netcdf_file
    rc = sim.receivers
    t = sim.times()[::rc.record_interval]
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with netcdf_file(path, "w") as nc:
        nc.createDimension("ncdim_%d" % len(t), len(t))
        nc.createDimension("ncdim_3", 3)
        v = nc.createVariable("time_points", "d", ("ncdim_%d" % len(t),))
        v[:] = t
        if sim.source is not None:
            nc.source_latitude, nc.source_longitude, nc.source_depth = sim.source.lat, sim.source.lon, sim.source.depth
        for i, key in enumerate(rc.keys):
            v = nc.createVariable("%s.%s" % (key, rc.components), "f", ("ncdim_%d" % len(t), "ncdim_3"))
            v[:] = series[:, i, :]
            v.latitude, v.longitude, v.depth = float(rc.lat[i]), float(rc.lon[i]), float(rc.depth[i])


def write_wisdom(path, dom, rel, e2p, rank, dist):
    """Domain::dumpWisdom (Domain.cpp:404-440): (s, z, learnt Nu, Nu) of every point that no lower rank also holds, gathered on
    rank 0 and written as the `axisem3d_wisdom` variable a later run reads with NU_TYPE wisdom."""
    nu = dom.getNuWisdom()
    owned = np.ones(len(rel["points"]), dtype=bool)
    msg = rel["msg"]
    for r, pts in zip(msg.mIProcComm, msg.mILocalPoints):
        if r < rank:
            owned[np.asarray(pts, dtype=np.int64)] = False                      # Domain::pointInPreviousRank
    rows = np.array([[p.crds[0], p.crds[1], nu[t], p.nu] for t, p in enumerate(rel["points"]) if owned[t]], dtype=np.float64).reshape(-1, 4)
    if dist is not None:
        everyone = [None] * dist.get_world_size() if rank == 0 else None
        dist.gather_object(rows, everyone, dst=0)
        if rank == 0:
            rows = np.concatenate(everyone, axis=0)
    if rank == 0:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        PL.NuWisdom(rows).write(path)


def main(argv=None):
    """One process per GPU: `python -m axisem3d_b200.run <run_dir>` or, for N GPUs of one node,
    `python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 -m axisem3d_b200.run <run_dir>` (RANK, WORLD_SIZE,
    LOCAL_RANK, MASTER_* from the environment).  The halo sum runs over peer-memory windows inside the step graph; the
    torch.distributed group (gloo) only carries the set-up handles and the station traces gathered on rank 0."""
    argv = sys.argv[1:] if argv is None else argv
    run_dir = argv[0] if argv else "."
    from .domain import Domain
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    dist = None
    device = local
    if world > 1:
        import torch
        import torch.distributed as dist
        device = local % max(torch.cuda.device_count(), 1)      # ranks may share a device (tests on a one-GPU box)
        torch.cuda.set_device(device)
        if not dist.is_initialized():
            dist.init_process_group("gloo", rank=rank, world_size=world)
    t0 = time.time()
    sim = Simulation(os.path.join(run_dir, "input"))
    dom = Domain(device)
    e2p = None
    if world > 1:
        e2p, _ = sim.partition(world)
        box = [e2p if rank == 0 else None]
        dist.broadcast_object_list(box, 0)                      # one partition vector for everybody
        e2p = box[0]
    rel = sim.release(dom, rank, e2p)
    if world > 1:
        dom.setMessaging(rel["msg"], rank, world, None)
    dom.finalize()
    if world > 1:
        dom.connectHalo(rel["msg"], rank, dist)
    rc = sim.receivers
    mine = rc.release(dom, rel["elements"], None if e2p is None else rel["dec"])
    if sim.learn[0]:
        dom.setLearnParameters(True, sim.learn[1], sim.learn[2])            # Mesh::release (Mesh.cpp:207)
    t1 = time.time()
    # ax3d_run_steps_record takes at most 4096 steps per call; 1000 = the reference's OUT_STATIONS_DUMP_INTERVAL default
    chunk = min(max(sim.par.get("OUT_STATIONS_DUMP_INTERVAL", int), 1), 4096)
    parts = []
    for k in range(0, len(sim.stf), chunk):
        if len(mine):
            parts.append(np.asarray(dom.runStepsRecord(sim.dt, sim.stf[k:k + chunk])))
        else:                                                   # a rank without stations steps along (same number of exchanges)
            dom.runSteps(sim.dt, sim.stf[k:k + chunk])
    if not dom.checkStability():
        raise RuntimeError("Domain::checkStability || Simulation has blown up")
    dom.synchronize()
    t2 = time.time()
    series = np.concatenate(parts, axis=0) if parts else np.zeros((len(sim.stf), 0, 3), np.float32)    # [step][my receivers][3], SPZ
    series = rc.rotate(series, mine)[::rc.record_interval]
    if world > 1:
        everyone = [None] * world if rank == 0 else None
        dist.gather_object((mine, series), everyone, dst=0)
        if rank == 0:
            full = np.zeros((series.shape[0], len(rc.keys), 3), np.float32)
            for idx, ser in everyone:
                if len(idx):
                    full[:, idx] = ser
            series = full
        dist.barrier()
    if sim.learn[0]:
        write_wisdom(os.path.join(run_dir, "output", sim.learn[3]), dom, rel, e2p, rank, dist)
    if rank == 0:
        fmts = [sim.par.get("OUT_STATIONS_FORMAT", str, k).lower() for k in range(sim.par.size("OUT_STATIONS_FORMAT"))]
        if any(f not in ("ascii", "netcdf", "netcdf_no_assemble") for f in fmts):
            raise RuntimeError("ReceiverCollection::buildInparam || Invalid parameter, keyword = OUT_STATIONS_FORMAT.")
        if "ascii" in fmts or not fmts:
            write_ascii(os.path.join(run_dir, "output", "stations"), sim, series)
        if "netcdf" in fmts or "netcdf_no_assemble" in fmts:
            write_netcdf(os.path.join(run_dir, "output", "stations", "axisem3d_synthetics.nc"), sim, series)
        print("axisem3d_b200: %d rank(s), %d elements and %d points on rank 0, dt = %.6g s, %d steps, %d stations; preloop %.1f s, "
              "time loop %.2f s (%.3f ms / step)" % (world, len(rel["elements"]), len(rel["points"]), sim.dt, len(sim.stf), len(rc.keys),
                                                     t1 - t0, t2 - t1, 1e3 * (t2 - t1) / max(len(sim.stf), 1)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
