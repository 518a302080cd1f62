// oracle/main_stubs.cpp -- TEST INFRASTRUCTURE, not product code.
//
// The reference links three Fortran files (SOLVER/CMakeLists.txt: simplex.f90, s20rts.f90, s40rts.f90); there is no Fortran
// compiler in this image.  Their entry points are only reached with ATTENUATION_SPECFEM_LEGACY (AttSimplex.cpp) or with the
// s20rts / s40rts volumetric models (whose data files are not in the checkout either), never in the runs oracle/ makes.  The
// symbols exist so that oracle/_ref/axisem3d_ref links; every one of them aborts loudly if it is ever called.
#include <cstdio>
#include <cstdlib>

static void ax_no_fortran(const char *what) {
    std::fprintf(stderr, "oracle/main_stubs.cpp: %s is Fortran code of the reference that could not be built here\n", what);
    std::abort();
}
extern "C" {
void __s20rts_MOD_initialize_s20rts() { ax_no_fortran("s20rts::initialize_s20rts"); }
void __s20rts_MOD_finalize_s20rts() { ax_no_fortran("s20rts::finalize_s20rts"); }
void __s20rts_MOD_perturb_s20rts() { ax_no_fortran("s20rts::perturb_s20rts"); }
void __s40rts_MOD_initialize_s40rts() { ax_no_fortran("s40rts::initialize_s40rts"); }
void __s40rts_MOD_finalize_s40rts() { ax_no_fortran("s40rts::finalize_s40rts"); }
void __s40rts_MOD_perturb_s40rts() { ax_no_fortran("s40rts::perturb_s40rts"); }
void __simplex_MOD_simplex_fminsearch() { ax_no_fortran("simplex::simplex_fminsearch"); }
}
