"""axisem3d_b200 -- B200-native stiffness + Newmark hot path of AxiSEM3D behind the reference's
Domain / Element / Point interface.  The CUDA library (csrc/, C-ABI in include/axisem3d_b200.h)
is loaded lazily by `axisem3d_b200.capi`; there is no CPU fallback."""
__version__ = "0.1.0"
