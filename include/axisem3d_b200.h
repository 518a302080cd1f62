/* axisem3d_b200.h -- C-ABI of the B200-native stiffness + Newmark hot path of AxiSEM3D.
 *
 * The reference has no FFI layer: its boundary is the C++ class API between Newmark/Mesh and
 * Domain/Element/Point.  Every entry point below names the reference interface
 * it replaces (S/ = SOLVER/src/ of kuangdai/AxiSEM3D).  INTEGRATION.md shows the binding a
 * maintainer adds on the reference side (a thin `Domain` facade, axisem3d_b200/host/).
 *
 * Conventions
 *  - plain pointers + sizes; all arrays are caller-owned HOST memory and are copied by the callee;
 *  - Real = float; complex = interleaved (re, im) floats; "RMatPP" = 25 values row-major (ipol, jpol);
 *  - per-element phi-dependent arrays ("RMatXN") are column-major Nr x 25 exactly as Eigen stores
 *    them in the reference: value(j, ipnt) at [ipnt * Nr + j], phi_j = 2 pi j / Nr;
 *  - every function returns 0 on success, non-zero on error; ax3d_last_error() returns the message
 *    in the reference's "Class::method || message" form (XMPI.cpp:52-90);
 *  - one host thread per domain (the reference is single-threaded per rank);
 *  - there is NO CPU fallback: every call fails if no CUDA device is usable.
 */
#ifndef AXISEM3D_B200_H
#define AXISEM3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ax3d_domain ax3d_domain;

/* ------------------------------------------------------------------ lifetime / errors */
/* Domain::Domain (S/core/domain/Domain.cpp:17-25) on CUDA device `device`. */
int ax3d_create(int device, ax3d_domain **out);
/* Domain::~Domain (Domain.cpp:27-44). */
int ax3d_destroy(ax3d_domain *dom);
/* XMPI::printException text of the last failure on this thread. */
const char *ax3d_last_error(void);
/* library/ABI version, device name; for load checks. */
int ax3d_version(void);

/* ------------------------------------------------------------------ setup (Mesh::release) */
/* Gradient::setGMat(G_GLL, G_GLJ) (S/core/element/grad/Gradient.cpp:324-329); G(i,j) = l_i'(x_j). */
int ax3d_set_gmat(ax3d_domain *dom, const double G_GLL[25], const double G_GLJ[25]);

/* mass descriptor: n_invmass == 1 -> Mass1D(invMass) (Mass1D.cpp:8), n_invmass == nr -> Mass3D(invMass[nr])
 * (Mass3D.cpp:9).  Domain::addPoint(new SolidPoint(nr, axial, crds, mass)) (SolidPoint.cpp:9; Domain.cpp:46).
 * Returns the domain tag (= insertion index) in *tag. */
int ax3d_add_solid_point(ax3d_domain *dom, int nr, int axial, const double crds[2],
                         int n_invmass, const float *invmass, int *tag);
/* Domain::addPoint(new SolidPoint(nr, axial, crds, new MassOcean1D / MassOcean3D(...))) (GLLPoint.cpp:57-72): a solid surface
 * point under an ocean load.  rows = 1: MassOcean1D(double mass, double massOcean, double theta) (MassOcean1D.cpp:8-28),
 * theta = normal_or_theta[0]; rows = nr: MassOcean3D(RDColX mass, RDColX massOcean, RDMatX3 unit_normal) (MassOcean3D.cpp:9-49),
 * normal column-major nr x 3. */
int ax3d_add_solid_point_ocean(ax3d_domain *dom, int nr, int axial, const double crds[2], int rows, const double *mass,
                               const double *mass_ocean, const double *normal_or_theta, int *tag);
/* Domain::addPoint(new FluidPoint(nr, axial, crds, mass, fluidSurf)) (FluidPoint.cpp:9). */
int ax3d_add_fluid_point(ax3d_domain *dom, int nr, int axial, const double crds[2],
                         int n_invmass, const float *invmass, int fluid_surf, int *tag);
/* Domain::addPoint + addSFPoint(new SolidFluidPoint(solid, fluid, couple)) (SolidFluidPoint.cpp:12;
 * GLLPoint.cpp:99-118).  n_sf == 1 -> SFCoupling1D(ns, nz, ns_invmf, nz_invmf) with
 * normal_un = {ns, 0, nz}, normal_as_invmf = {ns_invmf, 0, nz_invmf}; n_sf == nr -> SFCoupling3D with
 * two RMatX3 (column-major nr x 3). */
int ax3d_add_solid_fluid_point(ax3d_domain *dom, int nr, int axial, const double crds[2],
                               int n_invmass_solid, const float *invmass_solid,
                               int n_invmass_fluid, const float *invmass_fluid, int fluid_surf,
                               int n_sf, const float *normal_un, const float *normal_as_invmf, int *tag);

/* elastic law ids */
enum { AX3D_ISO = 0, AX3D_TI = 1, AX3D_ANISO = 2 };
/* attenuation ids */
enum { AX3D_ATT_NONE = 0, AX3D_ATT_FULL = 1, AX3D_ATT_CG4 = 2 };

/* Attenuation{1D,3D}_{Full,CG4}(nsls, alpha, beta, gamma, [Nu], dkappa, dmu, doKappa)
 * (S/core/element/material/attenuation).  rows = 1 (1D) or Nr (3D) must match the elastic rows;
 * dkappa/dmu: rows x P column-major with P = 25 (Full) or 4 (CG4; points (1,1),(1,3),(3,1),(3,3)). */
typedef struct ax3d_attenuation {
    int kind;            /* AX3D_ATT_* */
    int nsls;
    const float *alpha;  /* [nsls] */
    const float *beta;   /* [nsls] */
    const float *gamma;  /* [nsls] */
    const float *dkappa; /* [rows * P] */
    const float *dmu;    /* [rows * P] */
    int do_kappa;
} ax3d_attenuation;

/* Domain::addElement(new SolidElement(new Gradient(dsdxii, dsdeta, dzdxii, dzdeta, inv_s, axial), 0,
 * points, elastic)) (SolidElement.cpp:15; Gradient.cpp:9; Quad.cpp:386-404).
 *  geom: 5 x 25 doubles in the order dsdxii, dsdeta, dzdxii, dzdeta, inv_s;
 *  theta: 25 doubles (Element::formThetaMat, Element.cpp:48-58), used when law != AX3D_ISO;
 *  law/rows/coef: Isotropic{1D,3D}(lambda, mu) (2 arrays), TransverselyIsotropic{1D,3D}(A, C, F, L, N)
 *  (5), Anisotropic{1D,3D}(C11, C12, ..., C66) (21, upper triangle row by row); each array rows x 25
 *  column-major, rows = 1 (1D classes) or the element's Nr (3D classes);
 *  att: NULL or attenuation living in the same space.  Particle relabelling: ax3d_set_element_prt below. */
int ax3d_add_solid_element(ax3d_domain *dom, const int point_tags[25], const double *geom, int axial,
                           const double *theta, int law, int rows, const float *coef,
                           const ax3d_attenuation *att, int *tag);
/* Domain::addElement(new FluidElement(gradient, 0, points, new Acoustic{1D,3D}(K))) (FluidElement.cpp:15). */
int ax3d_add_fluid_element(ax3d_domain *dom, const int point_tags[25], const double *geom, int axial,
                           int rows, const float *K, int *tag);
/* The PRT* constructor argument of SolidElement / FluidElement (SolidElement.cpp:15-34, FluidElement.cpp:15-34; built by
 * Quad::createRelabelling, Quad.cpp:527-547): particle relabelling for undulated interfaces.  rows = 1: PRT_1D(const
 * std::array<RMatPP, 4> &X) (PRT_1D.h:12); rows = Nr: PRT_3D(const RMatXN4 &X) (PRT_3D.cpp:13-19).  X = [4][25][rows].
 * theta[25] = Element::formThetaMat() (Element.cpp:48-58).  Call after adding the element, before ax3d_finalize_setup.
 * The element then takes the 9-component path (computeGrad9 / computeQuad9, Gradient.cpp:84-204). */
int ax3d_set_element_prt(ax3d_domain *dom, int elem_tag, int rows, const float *X, const double theta[25]);

/* Domain::addSourceTerm(new SourceTerm(element, force)) (SourceTerm.cpp:15-27): force of point i is
 * nrow[i] x 3 complex column-major, concatenated over the 25 points; rows beyond the point's Nu+1 are dropped. */
int ax3d_add_source_term(ax3d_domain *dom, int elem_tag, const int nrow[25], const float *force);

/* Domain::setMessaging(MessagingInfo*, MessagingBuffer*) (Domain.h:36; XMPI.h:330-348; Mesh.cpp:189-204):
 * neighbour ranks and, per neighbour, the local point tags in global-GLL-tag order.  The halo sum runs over
 * NCCL: nccl_unique_id is the 128-byte ncclUniqueId every rank received from rank 0 (the C++ driver
 * broadcasts it with MPI_Bcast, the Python host with torch.distributed).  nproc == 1 or nneigh == 0 -> no-op. */
int ax3d_set_messaging(ax3d_domain *dom, int rank, int nproc, const void *nccl_unique_id,
                       int nneigh, const int *neigh_rank, const int *npoints, const int *point_tags);

/* XMPI::initialize analogue for the halo communicator: rank 0 creates the 128-byte ncclUniqueId and broadcasts it. */
int ax3d_nccl_unique_id(void *out128);

/* Peer-memory halo (replaces the MPI_Isend/Irecv pairs of Domain::assembleStiff, Domain.cpp:111-163, and
 * MessagingBuffer, S/core/domain/Domain.h:88-96, by direct NVLink stores into the neighbour's receive window).
 * Call after ax3d_finalize_setup on every rank that has neighbours:
 *   1. ax3d_halo_export: out_handle64 <- cudaIpcMemHandle_t of this rank's window (may be NULL), *out_ptr <- its device
 *      address (for neighbours living in the same process), neigh_begin[nneigh + 1] <- start of every neighbour's segment
 *      in the window (float2 units), *total <- window length per parity.
 *   2. the host exchanges (handle, neigh_begin, total, neighbour list) between neighbours (MPI_Allgather /
 *      torch.distributed.all_gather_object);
 *   3. ax3d_halo_connect: per neighbour n (ax3d_set_messaging order) its window as a 64-byte IPC handle (handles, other
 *      process of the same node) or a device pointer of this process (ptrs), peer_begin[n] = start of THIS rank's segment in
 *      it, peer_total[n] = its total, peer_slot[n] = index of this rank in ITS neighbour list.
 * Afterwards ax3d_assemble_stiff / ax3d_run_steps use the peer-memory kernels and multi-rank steps replay as CUDA graphs;
 * without these calls the NCCL send/recv path is used. */
int ax3d_halo_export(ax3d_domain *dom, void *out_handle64, void **out_ptr, long long *neigh_begin, long long *total);
int ax3d_halo_connect(ax3d_domain *dom, int nneigh, const void *handles, void *const *ptrs, const long long *peer_begin,
                      const long long *peer_total, const int *peer_slot);

/* End of Mesh::release: builds buckets, index maps, FFT plans, uploads everything. */
int ax3d_finalize_setup(ax3d_domain *dom);

/* ------------------------------------------------------------------ per step (Newmark::solve) */
/* Domain::updateNewmark(dt) (Domain.cpp:165-177). */
int ax3d_update_newmark(ax3d_domain *dom, double dt);
/* Domain::applySource(tstep) with stf = STF factor of that step (Domain.cpp:96-109). */
int ax3d_apply_source(ax3d_domain *dom, float stf);
/* Domain::computeStiff() (Domain.cpp:82-94). */
int ax3d_compute_stiff(ax3d_domain *dom);
/* Domain::coupleSolidFluid() (Domain.cpp:179-191). */
int ax3d_couple_solid_fluid(ax3d_domain *dom);
/* Domain::assembleStiff(phase): phase <= 0 pack + send/recv, phase >= 0 wait + unpack-add (Domain.cpp:111-163). */
int ax3d_assemble_stiff(ax3d_domain *dom, int phase);
/* Domain::checkStability (Domain.cpp:237-275): *stable = 0 if any displacement is non-finite. */
int ax3d_check_stability(ax3d_domain *dom, int *stable);
/* Domain::resetZero (Domain.cpp:67-74): all point fields and memory variables to zero. */
int ax3d_reset_zero(ax3d_domain *dom);
/* nsteps iterations of the Newmark::solve loop body (Newmark.cpp:47-93: update, source, stiff, couple,
 * assemble) with stf[i] as the source factor of step i; launched as one CUDA graph replay per step. */
int ax3d_run_steps(ax3d_domain *dom, int nsteps, double dt, const float *stf);
/* the same loop with Domain::record (Domain.cpp:207-220) after the update of every step, for the receivers registered
 * with ax3d_set_receivers: samples are buffered on the device (PointwiseRecorder's buffer, PointwiseRecorder.cpp:62-144,
 * with dump interval = nsteps <= 4096) and copied once: out[(i * nrec + r) * 3 + c], i = step, r = receiver. */
int ax3d_run_steps_record(ax3d_domain *dom, int nsteps, double dt, const float *stf, float *out);
/* the same loop, timed on the device with CUDA events on the launching stream; *ms = elapsed milliseconds. */
int ax3d_run_steps_timed(ax3d_domain *dom, int nsteps, double dt, const float *stf, float *ms);
/* blocks until the device finished all queued work of this domain. */
int ax3d_synchronize(ax3d_domain *dom);

/* ------------------------------------------------------------------ read-back / test hooks */
enum { AX3D_DISPL = 0, AX3D_VELOC = 1, AX3D_ACCEL = 2, AX3D_STIFF = 3 };
/* Point::getDispFourierSolid/Fluid and friends (Point.h:88-89): (Nu+1) x 3 complex column-major (solid part)
 * or (Nu+1) complex (fluid part) of point `tag`.  cap = capacity of out in complex numbers. */
int ax3d_get_point_field(ax3d_domain *dom, int tag, int field, int fluid_part, float *out, int cap);
int ax3d_set_point_field(ax3d_domain *dom, int tag, int field, int fluid_part, const float *in, int n);
/* bulk variants: all solid (fluid_part = 0) or fluid (= 1) blocks concatenated in point-tag order. */
int ax3d_get_field_bulk(ax3d_domain *dom, int field, int fluid_part, float *out, size_t cap_complex);
int ax3d_set_field_bulk(ax3d_domain *dom, int field, int fluid_part, const float *in, size_t n_complex);
/* number of complex entries of the solid / fluid field arrays. */
int ax3d_field_size(ax3d_domain *dom, int fluid_part, size_t *n_complex);
/* Element::computeGroundMotion(phi, weights, u_spz) for nrec receivers, out[3 * i + c]: SolidElement.cpp:189-216
 * (interpolated displacement) or FluidElement.cpp:163-215 (acoustic stress of the potential: gather, Gradient::computeGrad,
 * Acoustic1D/3D::strainToStress incl. the c2r/r2c pair for 3D material).  Evaluated on the device from the current
 * displacement, copied to host. */
int ax3d_record_ground_motion(ax3d_domain *dom, int nrec, const int *elem_tags, const float *phi,
                              const float *weights /* nrec x 25 */, float *out /* nrec x 3 */);

/* Element::computeStrain(phi, weights, RRow6&) and Element::computeCurl(phi, weights, RRow3&) after Element::forceTIso, as
 * PointwiseRecorder::record uses them for stations that dump strain / curl (PointwiseRecorder.cpp:96-135; SolidElement.cpp:
 * 219-345, 347-352): 6 Voigt strains resp. 3 curl components in the (R, T, Z) frame per receiver.  Solid elements without
 * particle relabelling. */
int ax3d_record_strain(ax3d_domain *dom, int nrec, const int *elem_tags, const float *phi, const float *weights, float *out /* nrec x 6 */);
int ax3d_record_curl(ax3d_domain *dom, int nrec, const int *elem_tags, const float *phi, const float *weights, float *out /* nrec x 3 */);

/* Domain::setPointwiseRecorder (Domain.h:34; ReceiverCollection.cpp:135-223): registers nrec receivers
 * (element, azimuth, 25 interpolation weights each) once ... */
int ax3d_set_receivers(ax3d_domain *dom, int nrec, const int *elem_tags, const float *phi, const float *weights);
/* ... and Domain::record -> PointwiseRecorder::record (Domain.cpp:207-220; PointwiseRecorder.cpp:62-144):
 * one displacement sample (s, phi, z) per registered receiver, nrec x 3 floats, device -> host. */
int ax3d_record(ax3d_domain *dom, float *out);

/* Wisdom learning (NU_WISDOM_LEARN in inparam.nu): Domain::setLearnParameters (Domain.h:39; LearnParameters = invoked, cutoff,
 * interval), Domain::learnWisdom(tstep) (Domain.cpp:384-402) -> Point::learnWisdom(cutoff) (SolidPoint.cpp:240-266,
 * FluidPoint.cpp:209-229), and the per-point Point::getNuWisdom() that Domain::dumpWisdom (Domain.cpp:404-440) writes.
 * ax3d_run_steps* call the learning step themselves once it is invoked (Newmark.cpp:89). */
int ax3d_set_learn_parameters(ax3d_domain *dom, int invoked, float cutoff, int interval);
int ax3d_learn_wisdom(ax3d_domain *dom, int tstep);
int ax3d_get_nu_wisdom(ax3d_domain *dom, int *nu_wisdom /* npoints, by point tag */, int npoints);

/* ------------------------------------------------------------------ measurement hooks */
/* number of CUDA kernel launches issued by this domain since creation (bench.py "gpu_launches"). */
int ax3d_launch_count(ax3d_domain *dom, long long *n);
/* sum over local GLL points of (Nu_p + 1): the work unit of BASELINE.json's metric. */
int ax3d_work_per_step(ax3d_domain *dom, long long *w);
/* algorithmic HBM bytes of one step (SURVEY.md §8d accounting) split per kernel family:
 * out[0] points, out[1] solid+fluid elements, out[2] halo. */
int ax3d_algorithmic_bytes(ax3d_domain *dom, double out[3]);
/* device time [ms] spent in each kernel family since the last call with reset != 0, measured with CUDA
 * events on the launching stream when profiling is enabled: out[0] newmark, [1] elements(stiff),
 * [2] solid-fluid + source, [3] halo. */
int ax3d_enable_timers(ax3d_domain *dom, int on);
/* Per-kernel statistics gathered while the timers are on (steps issued through ax3d_run_steps run eagerly then): one
 * CUDA-event pair on the launching stream around every hot kernel (the three launches of a split-pipeline chunk count as
 * one entry).  Entry `index` (name order) -> name, summed device time [ms], launches, summed ALGORITHMIC bytes of those
 * launches (elements of the launch; + 192 B per mode + mass for points advanced in-kernel).  *count = number
 * of entries; index < 0 reads only the count.  reset != 0 clears the table once the last entry (or the count) is read.
 * The caller picks the dominant kernel = the entry with the largest summed time (bench.py). */
int ax3d_kernel_stats(ax3d_domain *dom, int index, char *name, int name_cap, double *ms_total, long long *launches,
                      double *bytes_total, int *count, int reset);
int ax3d_get_timers(ax3d_domain *dom, double out_ms[4], int reset);
/* The plan of the azimuthal c2r / r2c of length nr that replaces SolverFFTW_N*::initialize's fftwf_plan_many_dft_*
 * (SolverFFTW_N6.cpp:16-43): radix sequence (2 ... 16, fewest stages) in DIF order.  Host only; no device needed. */
int ax3d_fft_plan(int nr, int *radices, int cap, int *nstages);
/* Measured element costs = the reference's cost-measure pass before its second partition (Mesh::measure,
 * Mesh.cpp:412-588: every element's computeStiff is timed, the times become the METIS vertex weights).  Runs
 * Domain::computeStiff `repeats` times with the device clocks on and returns cost_us[element tag] = microseconds of one
 * SM spent on the element (fused launches: clock64 around the element inside the kernel; k_elem1d / split pipeline:
 * event-timed launch duration shared by work units). */
int ax3d_measure_costs(ax3d_domain *dom, int repeats, double *cost_us, int nelem);

#ifdef __cplusplus
}
#endif
#endif /* AXISEM3D_B200_H */
