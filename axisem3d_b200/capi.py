"""ctypes binding of the C-ABI declared in include/axisem3d_b200.h (1:1, no logic)."""
from __future__ import annotations

import ctypes as C
import os

from . import _build

_lib = None

SYMBOLS = [
    "ax3d_create", "ax3d_destroy", "ax3d_last_error", "ax3d_version", "ax3d_set_gmat",
    "ax3d_add_solid_point", "ax3d_add_fluid_point", "ax3d_add_solid_fluid_point",
    "ax3d_add_solid_element", "ax3d_add_fluid_element", "ax3d_add_source_term", "ax3d_set_messaging",
    "ax3d_finalize_setup", "ax3d_update_newmark", "ax3d_apply_source", "ax3d_compute_stiff",
    "ax3d_couple_solid_fluid", "ax3d_assemble_stiff", "ax3d_check_stability", "ax3d_reset_zero",
    "ax3d_run_steps", "ax3d_synchronize", "ax3d_get_point_field", "ax3d_set_point_field",
    "ax3d_get_field_bulk", "ax3d_set_field_bulk", "ax3d_field_size", "ax3d_record_ground_motion",
    "ax3d_launch_count", "ax3d_work_per_step", "ax3d_algorithmic_bytes", "ax3d_enable_timers",
    "ax3d_get_timers", "ax3d_run_steps_timed", "ax3d_run_steps_record", "ax3d_kernel_stats", "ax3d_measure_costs", "ax3d_fft_plan", "ax3d_set_receivers", "ax3d_record", "ax3d_nccl_unique_id",
    "ax3d_halo_export", "ax3d_halo_connect", "ax3d_set_learn_parameters", "ax3d_learn_wisdom", "ax3d_get_nu_wisdom", "ax3d_set_element_prt", "ax3d_add_solid_point_ocean", "ax3d_record_strain", "ax3d_record_curl",
]


class Attenuation(C.Structure):
    _fields_ = [("kind", C.c_int), ("nsls", C.c_int), ("alpha", C.POINTER(C.c_float)),
                ("beta", C.POINTER(C.c_float)), ("gamma", C.POINTER(C.c_float)),
                ("dkappa", C.POINTER(C.c_float)), ("dmu", C.POINTER(C.c_float)), ("do_kappa", C.c_int)]


def lib_path():
    # AX3D_LIB: developer override to A/B a differently compiled build of the same source (profiles/microbench/variants)
    return os.environ.get("AX3D_LIB") or _build.LIB


def load(build_if_missing=True):
    """Load libaxisem3d_b200.so; fails loudly when it is missing (no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise RuntimeError("axisem3d_b200: CUDA library %s is missing; run __graft_entry__.build()" % path)
        _build.build()
    # The library links libnccl.so.2.  torch bundles a newer NCCL under the same SONAME than the system one; whichever is
    # mapped first wins for the whole process, and libtorch_cuda does not load against the older system copy.  Import
    # torch first so that one NCCL (torch's) serves both.
    # torch is optional for this binding: without it the system libnccl serves the library alone.
    try:
        import torch  # noqa: F401
    except ImportError:
        pass
    lib = C.CDLL(path)
    lib.ax3d_last_error.restype = C.c_char_p
    for s in SYMBOLS:
        getattr(lib, s)     # raises AttributeError if the .so does not export a declared symbol
    vp, i, d, f = C.c_void_p, C.c_int, C.c_double, C.c_float
    pi_, pd, pf = C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_float)
    lib.ax3d_create.argtypes = [i, C.POINTER(vp)]
    lib.ax3d_destroy.argtypes = [vp]
    lib.ax3d_set_gmat.argtypes = [vp, pd, pd]
    lib.ax3d_add_solid_point.argtypes = [vp, i, i, pd, i, pf, pi_]
    lib.ax3d_add_fluid_point.argtypes = [vp, i, i, pd, i, pf, i, pi_]
    lib.ax3d_add_solid_point_ocean.argtypes = [vp, i, i, pd, i, pd, pd, pd, pi_]
    lib.ax3d_add_solid_fluid_point.argtypes = [vp, i, i, pd, i, pf, i, pf, i, i, pf, pf, pi_]
    lib.ax3d_add_solid_element.argtypes = [vp, pi_, pd, i, pd, i, i, pf, C.POINTER(Attenuation), pi_]
    lib.ax3d_add_fluid_element.argtypes = [vp, pi_, pd, i, i, pf, pi_]
    lib.ax3d_add_source_term.argtypes = [vp, i, pi_, pf]
    lib.ax3d_set_element_prt.argtypes = [vp, i, i, pf, pd]
    lib.ax3d_set_messaging.argtypes = [vp, i, i, vp, i, pi_, pi_, pi_]
    lib.ax3d_finalize_setup.argtypes = [vp]
    lib.ax3d_update_newmark.argtypes = [vp, d]
    lib.ax3d_apply_source.argtypes = [vp, f]
    lib.ax3d_compute_stiff.argtypes = [vp]
    lib.ax3d_couple_solid_fluid.argtypes = [vp]
    lib.ax3d_assemble_stiff.argtypes = [vp, i]
    lib.ax3d_check_stability.argtypes = [vp, pi_]
    lib.ax3d_reset_zero.argtypes = [vp]
    lib.ax3d_set_learn_parameters.argtypes = [vp, i, f, i]
    lib.ax3d_learn_wisdom.argtypes = [vp, i]
    lib.ax3d_get_nu_wisdom.argtypes = [vp, pi_, i]
    lib.ax3d_run_steps.argtypes = [vp, i, d, pf]
    lib.ax3d_synchronize.argtypes = [vp]
    lib.ax3d_run_steps_timed.argtypes = [vp, i, d, pf, pf]
    lib.ax3d_run_steps_record.argtypes = [vp, i, d, pf, pf]
    lib.ax3d_measure_costs.argtypes = [vp, i, pd, i]
    lib.ax3d_fft_plan.argtypes = [i, pi_, i, pi_]
    lib.ax3d_kernel_stats.argtypes = [vp, i, C.c_char_p, i, pd, C.POINTER(C.c_longlong), pd, pi_, i]
    lib.ax3d_set_receivers.argtypes = [vp, i, pi_, pf, pf]
    lib.ax3d_record.argtypes = [vp, pf]
    lib.ax3d_nccl_unique_id.argtypes = [vp]
    lib.ax3d_halo_export.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    lib.ax3d_halo_connect.argtypes = [vp, i, vp, C.POINTER(vp), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), pi_]
    lib.ax3d_get_point_field.argtypes = [vp, i, i, i, pf, i]
    lib.ax3d_set_point_field.argtypes = [vp, i, i, i, pf, i]
    lib.ax3d_get_field_bulk.argtypes = [vp, i, i, pf, C.c_size_t]
    lib.ax3d_set_field_bulk.argtypes = [vp, i, i, pf, C.c_size_t]
    lib.ax3d_field_size.argtypes = [vp, i, C.POINTER(C.c_size_t)]
    lib.ax3d_record_ground_motion.argtypes = [vp, i, pi_, pf, pf, pf]
    lib.ax3d_record_strain.argtypes = [vp, i, pi_, pf, pf, pf]
    lib.ax3d_record_curl.argtypes = [vp, i, pi_, pf, pf, pf]
    lib.ax3d_launch_count.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.ax3d_work_per_step.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.ax3d_algorithmic_bytes.argtypes = [vp, pd]
    lib.ax3d_enable_timers.argtypes = [vp, i]
    lib.ax3d_get_timers.argtypes = [vp, pd, i]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError(load().ax3d_last_error().decode())
