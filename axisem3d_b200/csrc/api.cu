// api.cu -- host side of the C-ABI (include/axisem3d_b200.h): collects the solver objects the reference's
// Mesh::release would construct, flattens them into SoA device arrays at ax3d_finalize_setup, and launches the
// kernels of kernels.cuh for the Domain step verbs.  No CPU fallback: every entry point needs a CUDA device.
#include "../../include/axisem3d_b200.h"
#include "kernels.cuh"
#include "fused.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef AX3D_WITH_NCCL
#include <nccl.h>
#endif

static thread_local std::string g_last_error;

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            throw std::runtime_error(std::string("ax3d::cuda || ") + cudaGetErrorString(e_) + " || " #call);  \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    void alloc(size_t count) {
        release();
        n = count;
        if (count) CK(cudaMalloc(&p, count * sizeof(T)));
    }
    // The domain's streams are non-blocking: they do not synchronise with the legacy stream these set-up copies use, a
    // pageable H2D cudaMemcpy may return before its DMA has landed and cudaMemset is asynchronous.  Every set-up copy
    // therefore drains the legacy stream before returning, so that work enqueued afterwards on any stream sees the data.
    void upload(const std::vector<T> &h) {
        alloc(h.size());
        if (n) {
            CK(cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice));
            CK(cudaStreamSynchronize(cudaStreamLegacy));
        }
    }
    void zero() {
        if (n) {
            CK(cudaMemset(p, 0, n * sizeof(T)));
            CK(cudaStreamSynchronize(cudaStreamLegacy));
        }
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

// ------------------------------------------------------------------------------------------ host descriptors
struct HPoint {
    int kind;   // 0 solid, 1 fluid, 2 solid-fluid
    int nr, nu;
    bool axial, fluid_surf;
    std::vector<float> im_s, im_f;       // inverse mass (1 or nr entries)
    double crds[2] = {0, 0};             // (s, z): Point::getCoords, used by Element::formThetaMat (Element.cpp:48-58)
    int ocean = 0;                       // 0 none, 1 MassOcean1D, 3 MassOcean3D (solid points only)
    float oc_imZ = 0, oc_sint = 0, oc_cost = 0;   // MassOcean1D (im_s[0] = 1 / m)
    std::vector<float> oc_ns;            // MassOcean3D: normal * scal, column-major nr x 3
    int n_sf = 0;
    std::vector<float> n_un, n_as;       // SF coupling (3 or 3*nr entries, column-major nr x 3)
    int s_idx = -1, f_idx = -1;
    size_t s_off = 0, f_off = 0;
};
struct HAtt {
    int kind = ATT_NONE, nsls = 0, do_kappa = 0;
    std::vector<float> alpha, beta, gamma, dkappa, dmu;
};
struct HElem {
    bool fluid;
    int pt[AX_NPE];
    double geom[5 * AX_NPE];
    double theta[AX_NPE];
    bool axial;
    int law, rows, nr, nu, ncoef;
    std::vector<float> coef;
    HAtt att;
    int prt_rows = 0;         // particle relabelling: 0 none, 1 PRT_1D, Nr PRT_3D
    std::vector<float> prtX;  // [4][25][prt_rows]
    int cls = -1, idx = -1;   // class and index inside the class after finalize
    bool bnd = false;         // touches a point shared with another rank
    double alg_b = 0;         // algorithmic bytes of one stiffness evaluation of this element
};
struct HSource {
    int elem;
    int nrow[AX_NPE];
    std::vector<float> force;
};

// 227 KB per CTA minus the fused kernel's static shared memory (descriptors, plans, mbarriers, arrival ring, geometry: < 4.5 KB)
#define AX_FUSED_DYN_MAX (232448 - 4608)
#ifndef AX_DUAL_DEFAULT
#define AX_DUAL_DEFAULT 2   // fluid chain of a step on a second stream (step_body): 0 off, 1 fork at the top of the step, 2 fork before the solid elements (B200, cfg2: 0.318 / 0.313 / 0.311 ms per step); AX3D_DUAL overrides
#endif
// k_fft3d_v2 instance by shared-memory residency: every instance is compiled for 1024 resident threads per SM (64 registers;
// the kernel is latency-bound and prefers warps to registers, profiles/microbench/occ_ab*.sh): four 256-thread CTAs when four
// tiles fit, two 512-thread CTAs when two or three fit, one 1024-thread CTA otherwise
static int fft_threads(size_t smem_bytes) {
    const size_t per_sm = (size_t)(228 * 1024) / (smem_bytes + 1024);
    return per_sm >= 4 ? 256 : per_sm >= 2 ? 512 : 1024;
}
enum { CLS_S1D = 0, CLS_F1D = 1, CLS_S3D = 2, CLS_F3D = 3, NCLS = 4 };

struct Chunk {   // a run of 3D elements of one class whose spectra fit the scratch ring together
    int cls;
    int w_begin, w_count;       // grad/quad work items
    int f_begin, f_count;       // fft work items
    size_t fft_smem;
    int fft_np, fft_nt;         // k_fft3d_v2 instance: points per CTA (5 or 1), threads (256: two CTAs per SM, or 512)
    int prt;                    // 1: solid elements with particle relabelling (5 Z-form pairs per point instead of 3)
    double alg_b;               // algorithmic bytes of the elements of the chunk
};

struct FusedLaunch {   // all fused 3D elements of one class: one persistent k_elem3d_fused<FLUID, ...> launch
    int cls, first, count;
    int nww;                     // Newmark warps per CTA (0: none; the solid launch uses AX_NWW when the domain has plain points)
    int tile_cap;                // float2 capacity R of the tile region (U | TW | Z, laid out per element by plan_fused_element)
    int nitems, nb_items;        // work items (element, row-group pass) of the launch; boundary items among them (first in the list)
    size_t smem;
    int grid;
    double alg_b;                // algorithmic bytes of the elements of the launch
};

struct ax3d_domain {
    int device = 0;
    bool finalized = false;
    cudaStream_t stream = nullptr;
    double G_GLL[25], G_GLJ[25];
    bool have_g = false;
    std::vector<HPoint> points;
    std::vector<HElem> elems;
    std::vector<HSource> sources;
    long long launches = 0;
    long long work = 0;
    int num_sm = 148;
    double alg_bytes[3] = {0, 0, 0};

    // ---- device: points
    size_t ns = 0, nf = 0;            // solid / fluid point counts
    size_t s_len = 0, f_len = 0;      // field lengths (float2)
    DevBuf<float2> s_field[4], f_field[4];
    DevBuf<unsigned> s_off, f_off;
    DevBuf<int> s_nu, f_nu, s_nr, f_nr, s_row_point, f_row_point, s_row_start, f_row_start;
    DevBuf<unsigned char> s_flags, f_flags;
    DevBuf<float> s_invmass, f_invmass;
    PointTab s_tab{}, f_tab{};
    std::vector<Mass3DItem> h_m3d_s, h_m3d_f;
    std::vector<Ocean1DItem> h_oc1d;
    DevBuf<Ocean1DItem> oc1d;
    int oc1d_max_m = 1;
    DevBuf<Mass3DItem> m3d_s, m3d_f;
    size_t m3d_smem_s = 0, m3d_smem_f = 0;
    DevBuf<float> impool;
    // ---- plans
    std::vector<FftPlan> h_plans;
    std::vector<std::vector<int>> h_perm;
    std::map<int, int> plan_of_n;
    DevBuf<FftPlan> plans;
    DevBuf<float2> twpool;
    std::vector<float2> h_stw;      // per-stage twiddle tables of all plans (fused kernel)
    DevBuf<float2> stwpool;
    std::vector<FusedLaunch> fused;
    DevBuf<unsigned> fused_work;    // per launch: {next work-item index, finished warps}
    DevBuf<int> fused_items[NCLS];  // per class: (element index in launch << 3) | pass, boundary items first, then by cost
    // ---- in-kernel Newmark (fused.cuh): "plain" solid points are advanced by the element kernel of the previous step
    bool nw_allowed = true;         // AX3D_NO_NW=1 switches the mechanism off
    int n_plain = 0;
    DevBuf<int> nw_cnt, nw_queue;
    DevBuf<unsigned> nw_ctl;        // tail, head, -, error flag
    DevBuf<int> s_row_point_sp, s_row_start_sp;
    PointTab s_tab_sp{};            // rows of the solid points the stand-alone Newmark kernel still owns
    DevBuf<unsigned long long> nw_dbg;
    // device-side record ring of ax3d_run_steps_record: [step][receiver][3]
    DevBuf<float> rec_ring;
    float *rec_ring_host = nullptr;
    size_t rec_ring_steps = 0;
    bool plain_advanced = false;    // the plain points already hold the state of the step about to start
    // ---- elements
    std::vector<ElemDesc> h_desc[NCLS];
    std::vector<double> h_alg_b[NCLS];   // algorithmic bytes of every element, in class order
    DevBuf<ElemDesc> desc[NCLS];
    DevBuf<float> geom, coef, attpar, attstate3d;
    DevBuf<float2> attstate1d;
    DevBuf<int> w_elem[NCLS], w_a0[NCLS];
    int n_work[NCLS] = {0, 0, 0, 0};
    bool cls_prt[NCLS] = {false, false, false, false};   // the class has elements with particle relabelling (1D classes: kernel instance)
    DevBuf<FftItem> fft_items[NCLS];
    std::vector<Chunk> chunks;
    DevBuf<float2> scratch;
    // ---- source, solid-fluid
    DevBuf<unsigned> src_off;
    DevBuf<float2> src_val;
    int n_src = 0;
    DevBuf<unsigned> sf_s_off, sf_f_off;
    DevBuf<int> sf_nu, sf_row_point, sf_row_start;
    DevBuf<float> sf_cpl;
    SFTab sf_tab{};
    std::vector<SF3DItem> h_sf3d;
    DevBuf<SF3DItem> sf3d;
    DevBuf<float> sf3d_pool;
    size_t sf3d_smem = 0;
    // ---- halo
    int rank = 0, nproc = 1;
    std::vector<int> neigh_rank;
    std::vector<std::vector<int>> neigh_pts;
    std::vector<size_t> neigh_begin;   // offsets into the packed buffers (float2)
    DevBuf<unsigned> halo_idx;
    DevBuf<float2> halo_send, halo_recv;
    // peer-memory halo (k_halo_put / k_halo_wait_add): own window = [2][total] float2 followed by one arrival counter
    // per neighbour; the neighbours' windows are mapped with cudaIpcOpenMemHandle (or passed as raw pointers in-process)
    DevBuf<float2> halo_win;
    DevBuf<unsigned> halo_step;
    std::vector<float2 *> peer_win;          // per neighbour: where my segment starts in its window (parity 0)
    std::vector<size_t> peer_stride;         // per neighbour: its parity stride (= its total)
    std::vector<unsigned *> peer_count;      // per neighbour: its arrival counter for me
    std::vector<void *> peer_mapped;         // bases returned by cudaIpcOpenMemHandle (closed at destroy)
    bool peer_halo = false;
    // in-kernel put (fused.cuh: halo_put_cta): the solid element kernel sends the boundary forces itself
    std::vector<char> is_halo_point;        // per point tag
    SFTab sf_tab_halo{}, sf_tab_rest{};     // 1D-coupled solid-fluid points on / off the halo
    DevBuf<unsigned> sfh_s_off, sfh_f_off, sfr_s_off, sfr_f_off;
    DevBuf<int> sfh_nu, sfh_row_point, sfh_row_start, sfr_nu, sfr_row_point, sfr_row_start;
    DevBuf<float> sfh_cpl, sfr_cpl;
    bool halo_sf3d = false;                 // a 3D-coupled solid-fluid point lies on the halo: no in-kernel put
    DevBuf<HaloTab> halo_tab;
    DevBuf<unsigned> halo_bcnt;
    int n_bnd_fused = 0;                    // boundary elements in the solid fused launch
    bool inkernel_put = false;
    // ax3d_measure_costs: per-element SM cycles recorded by the fused launches ([class offset + index in launch])
    DevBuf<unsigned> cost_buf;
    size_t cost_off[NCLS] = {0, 0, 0, 0};
    bool have_uid = false;
    unsigned char uid[128];
#ifdef AX3D_WITH_NCCL
    ncclComm_t comm = nullptr;
#endif
    // ---- misc
    DevBuf<int> bad_flag;
    bool timers = false;
    double timer_ms[4] = {0, 0, 0, 0};
    // per-kernel statistics while the timers are on (eager steps): name -> summed device time, launches, summed algorithmic
    // bytes of those launches.  bench.py takes the entry with the largest summed time as the dominant kernel.
    struct KStat { double ms = 0; long long n = 0; double bytes = 0; };
    std::map<std::string, KStat> kstats;
    double dom_bytes[2] = {0, 0};       // solid fused launch: algorithmic bytes {elements, in-kernel Newmark points}
    double pts_bytes[2] = {0, 0};       // algorithmic bytes of all solid / all fluid points
    double cls_bytes[4] = {0, 0, 0, 0}; // algorithmic bytes of the elements of each class that go through k_elem1d / the fused launch
    cudaEvent_t ev2 = nullptr, ev3 = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // fluid chain of a step on its own stream (step_body)
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool dual_chain = false;
    int dual_mode = 1;
    // receivers (scratch for ax3d_record_ground_motion)
    DevBuf<RecvItem> rec_items;
    DevBuf<float> rec_w, rec_out;
    float *rec_host = nullptr;
    size_t rec_cap = 0;
    // source factor: pinned host ring -> device scalar (4-byte H2D per step)
    DevBuf<float> stf_dev;
    float *stf_pinned = nullptr;
    int stf_slot = 0;
    // CUDA graph of one step (single-GPU path)
    cudaGraph_t graph[8] = {};          // [record * 4 + special_only * 2 + nw_on]
    cudaGraphExec_t graph_exec[8] = {};
    double graph_dt = 0;
    bool use_graph = true;
    // receivers registered with ax3d_set_receivers
    // wisdom learning (Domain::setLearnParameters / learnWisdom / dumpWisdom)
    bool learn_invoked = false;
    float learn_cutoff = 0.f;
    int learn_interval = 1;
    long long tstep = 0;                        // steps taken by ax3d_run_steps* since set-up (Newmark.cpp:47: tstep - 1)
    DevBuf<float> s_wis_max, f_wis_max;
    DevBuf<int> s_wis_nu, f_wis_nu;
    int nrec_c[NCLS] = {0, 0, 0, 0};            // per element class, rows of the device output in class order
    std::vector<int> rec_where_c[NCLS];
    DevBuf<RecvItem> rec_items_c[NCLS];
    DevBuf<float> rec_w_c[NCLS];
    size_t rec_fluid_smem = 0;                  // k_ground_motion_fluid: largest FFT tile of the registered fluid receivers
    int nrec_total() const { return nrec_c[0] + nrec_c[1] + nrec_c[2] + nrec_c[3]; }
};
#define STF_RING 4096

static void fail(const std::string &m) { throw std::runtime_error(m); }

// ------------------------------------------------------------------------------------------ FFT plans
static std::vector<int> choose_radices(int N) {
    const RadixList rl = choose_radices_plan(N);   // fft.cuh
    if (rl.n < 0) fail("ax3d::plan || Nr = " + std::to_string(N) + " is not a lucky number (prime factor > 13 or too many stages; PreloopFFTW.cpp:59-99)");
    std::vector<int> r(rl.r, rl.r + rl.n);
    if (r.empty()) r.push_back(1);
    return r;
}

static int get_plan(ax3d_domain *d, int N) {
    auto it = d->plan_of_n.find(N);
    if (it != d->plan_of_n.end()) return it->second;
    FftPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.N = N;
    std::vector<int> rad = choose_radices(N);
    if (rad.size() == 1 && rad[0] == 1) {
        pl.nstages = 0;
    } else {
        pl.nstages = (int)rad.size();
        for (int s = 0; s < pl.nstages; ++s) pl.radix[s] = rad[s];
    }
    // perm[pos] = n : sample index stored at position pos after the DIF transform
    std::vector<int> perm(N);
    for (int k = 0; k < N; ++k) {
        int kk = k, pos = 0, stride = N;
        for (int s = 0; s < pl.nstages; ++s) {
            stride /= pl.radix[s];
            pos += (kk % pl.radix[s]) * stride;
            kk /= pl.radix[s];
        }
        perm[pos] = k;
    }
    // per-stage twiddle tables T_s[j * R + p] = exp(+2 pi i j p / L_s) (fused.cuh)
    pl.stw_base = (int)d->h_stw.size();
    {
        int L = N;
        for (int s = 0; s < pl.nstages; ++s) {
            const int R = pl.radix[s], Ls = L / R;
            pl.stw_off[s] = -1;
            if (Ls > 1) {
                pl.stw_off[s] = (int)d->h_stw.size();
                for (int j = 0; j < Ls; ++j)
                    for (int p = 0; p < R; ++p) {
                        const double a = 2.0 * M_PI * (double)j * (double)p / (double)L;
                        d->h_stw.push_back(make_float2((float)cos(a), (float)sin(a)));
                    }
            }
            L = Ls;
        }
    }
    pl.stw_len = (int)d->h_stw.size() - pl.stw_base;
    int id = (int)d->h_plans.size();
    d->h_plans.push_back(pl);
    d->h_perm.push_back(perm);
    d->plan_of_n[N] = id;
    return id;
}

// ------------------------------------------------------------------------------------------ setup helpers
static void check_open(ax3d_domain *d) {
    if (!d) fail("ax3d || null domain");
    if (d->finalized) fail("Domain::add || setup already finalized");
}
static void check_final(ax3d_domain *d) {
    if (!d) fail("ax3d || null domain");
    if (!d->finalized) fail("Domain || ax3d_finalize_setup has not been called");
}

static void set_mass(std::vector<float> &dst, int nr, int n, const float *im, const char *who) {
    if (n != 1 && n != nr) fail(std::string(who) + " || Mass3D::checkCompatibility || Incompatible size.");
    dst.assign(im, im + n);
}

static void set_crds(ax3d_domain *d, int tag, const double crds[2]) {
    if (crds) { d->points[tag].crds[0] = crds[0]; d->points[tag].crds[1] = crds[1]; }
}
static int add_point(ax3d_domain *d, int kind, int nr, int axial, int n_s, const float *im_s, int n_f, const float *im_f,
                     int fluid_surf, int n_sf, const float *n_un, const float *n_as) {
    check_open(d);
    if (nr < 1) fail("Point::Point || nr must be positive");
    HPoint p;
    p.kind = kind;
    p.nr = nr;
    p.nu = nr / 2;
    p.axial = axial != 0;
    p.fluid_surf = fluid_surf != 0;
    if (kind == 0 || kind == 2) set_mass(p.im_s, nr, n_s, im_s, "SolidPoint::SolidPoint");
    if (kind == 1 || kind == 2) set_mass(p.im_f, nr, n_f, im_f, "FluidPoint::FluidPoint");
    if (kind == 2) {
        if (n_sf != 1 && n_sf != nr) fail("SFCoupling3D::checkCompatibility || Incompatible size.");
        p.n_sf = n_sf;
        p.n_un.assign(n_un, n_un + 3 * n_sf);
        p.n_as.assign(n_as, n_as + 3 * n_sf);
    }
    d->points.push_back(std::move(p));
    return (int)d->points.size() - 1;
}

static void common_elem(ax3d_domain *d, HElem &e, const int tags[25], const double *geom, int axial, bool fluid) {
    check_open(d);
    e.fluid = fluid;
    e.axial = axial != 0;
    int nr = -1;
    for (int i = 0; i < AX_NPE; ++i) {
        int t = tags[i];
        if (t < 0 || t >= (int)d->points.size()) fail("Element::Element || invalid point tag");
        const HPoint &p = d->points[t];
        if (fluid && p.kind == 0) fail("Point::scatterDisplToElement || Incompatible point type.");
        if (!fluid && p.kind == 1) fail("Point::scatterDisplToElement || Incompatible point type.");
        e.pt[i] = t;
        nr = std::max(nr, p.nr);
    }
    e.nr = nr;
    e.nu = nr / 2;
    memcpy(e.geom, geom, sizeof(e.geom));
}

// ------------------------------------------------------------------------------------------ finalize
static void set_fused_smem(int device, const FusedLaunch &f);
typedef void (*fft_kernel_t)(const ElemDesc *, const FftItem *, const FftPlan *, const float2 *, const float *, const float *, float *,
                             float2 *);
static fft_kernel_t fft_kernel(const Chunk &ch) {
    const bool fluid = ch.cls == CLS_F3D;
    if (ch.prt && !fluid) {   // solid elements with PRT: 5 Z-form pairs per point
        if (ch.fft_np == 1) return k_fft3d_v2<false, 1, 256, 5, true>;
        if (ch.fft_nt == 256) return k_fft3d_v2<false, 5, 256, 5, true>;
        if (ch.fft_nt == 512) return k_fft3d_v2<false, 5, 512, 5, true>;
        return k_fft3d_v2<false, 5, 1024, 5, true>;
    }
    if (ch.prt && fluid) {    // fluid elements with PRT: same 2 pairs, PRT applied around Acoustic3D
        if (ch.fft_np == 1) return k_fft3d_v2<true, 1, 256, 0, true>;
        if (ch.fft_nt == 256) return k_fft3d_v2<true, 5, 256, 0, true>;
        if (ch.fft_nt == 512) return k_fft3d_v2<true, 5, 512, 0, true>;
        return k_fft3d_v2<true, 5, 1024, 0, true>;
    }
    if (ch.fft_np == 1) return fluid ? k_fft3d_v2<true, 1, 256> : k_fft3d_v2<false, 1, 256>;
    if (ch.fft_nt == 256) return fluid ? k_fft3d_v2<true, 5, 256> : k_fft3d_v2<false, 5, 256>;
    if (ch.fft_nt == 512) return fluid ? k_fft3d_v2<true, 5, 512> : k_fft3d_v2<false, 5, 512>;
    return fluid ? k_fft3d_v2<true, 5, 1024> : k_fft3d_v2<false, 5, 1024>;
}

static void finalize(ax3d_domain *d) {
    check_open(d);
    if (!d->have_g) fail("Gradient::setGMat || ax3d_set_gmat has not been called");
    CK(cudaSetDevice(d->device));
    CK(cudaDeviceGetAttribute(&d->num_sm, cudaDevAttrMultiProcessorCount, d->device));
    // ---------------- constants
    {
        float G[2][25];
        for (int i = 0; i < 25; ++i) {
            G[0][i] = (float)d->G_GLL[i];
            G[1][i] = (float)d->G_GLJ[i];
        }
        CK(cudaMemcpyToSymbol(c_G, G, sizeof(G)));
        float hc[17][16], hs[17][16];
        memset(hc, 0, sizeof(hc));
        memset(hs, 0, sizeof(hs));
        for (int R = 1; R <= 16; ++R)
            for (int k = 0; k < R && k < 16; ++k) {
                hc[R][k] = (float)cos(2.0 * M_PI * k / R);
                hs[R][k] = (float)sin(2.0 * M_PI * k / R);
            }
        CK(cudaMemcpyToSymbol(c_cos, hc, sizeof(hc)));
        CK(cudaMemcpyToSymbol(c_sin, hs, sizeof(hs)));
    }
    // ---------------- points
    std::vector<unsigned> so, fo;
    std::vector<int> snu, fnu, snr, fnr, srp, frp, srs, frs;
    std::vector<unsigned char> sfl, ffl;
    std::vector<float> sim, fim, impool;
    size_t s_len = 0, f_len = 0;
    double bytes_pts = 0;
    for (size_t t = 0; t < d->points.size(); ++t) {
        HPoint &p = d->points[t];
        const int M = p.nu + 1;
        d->work += M;
        if (p.kind == 0 || p.kind == 2) {
            p.s_idx = (int)so.size();
            p.s_off = s_len;
            so.push_back((unsigned)s_len);
            snu.push_back(p.nu);
            snr.push_back(p.nr);
            srs.push_back((int)srp.size());
            for (int a = 0; a < M; ++a) srp.push_back(p.s_idx);
            const bool m3 = p.im_s.size() > 1;
            sfl.push_back((unsigned char)((p.axial ? 1 : 0) | (m3 ? 4 : 0)));
            sim.push_back((m3 || p.ocean) ? 1.f : p.im_s[0]);   // Mass3D / MassOcean*: the mass kernels have already applied it
            if (p.ocean == 1) d->h_oc1d.push_back(Ocean1DItem{p.s_idx, p.oc_imZ, p.im_s[0], p.oc_sint, p.oc_cost});
            if (m3) {
                Mass3DItem it;
                it.point = p.s_idx;
                it.plan_id = get_plan(d, p.nr);
                it.im_off = (long long)impool.size();
                it.ns_off = -1;
                const std::vector<int> &perm = d->h_perm[it.plan_id];
                for (int pos = 0; pos < p.nr; ++pos) impool.push_back(p.im_s[perm[pos]]);
                if (p.ocean == 3) {
                    it.ns_off = (long long)impool.size();
                    for (int c = 0; c < 3; ++c)
                        for (int pos = 0; pos < p.nr; ++pos) impool.push_back(p.oc_ns[(size_t)c * p.nr + perm[pos]]);
                }
                d->h_m3d_s.push_back(it);
                d->m3d_smem_s = std::max(d->m3d_smem_s, (size_t)3 * p.nr * sizeof(float2));
            }
            s_len += (size_t)3 * M;
            bytes_pts += 192.0 * M + (m3 ? 4.0 * p.nr : 4.0);
            d->pts_bytes[0] += 192.0 * M + (m3 ? 4.0 * p.nr : 4.0);
        }
        if (p.kind == 1 || p.kind == 2) {
            p.f_idx = (int)fo.size();
            p.f_off = f_len;
            fo.push_back((unsigned)f_len);
            fnu.push_back(p.nu);
            fnr.push_back(p.nr);
            frs.push_back((int)frp.size());
            for (int a = 0; a < M; ++a) frp.push_back(p.f_idx);
            const bool m3 = p.im_f.size() > 1;
            ffl.push_back((unsigned char)((p.axial ? 1 : 0) | (p.fluid_surf ? 2 : 0) | (m3 ? 4 : 0)));
            fim.push_back(m3 ? 1.f : p.im_f[0]);
            if (m3) {
                Mass3DItem it;
                it.point = p.f_idx;
                it.plan_id = get_plan(d, p.nr);
                it.im_off = (long long)impool.size();
                const std::vector<int> &perm = d->h_perm[it.plan_id];
                for (int pos = 0; pos < p.nr; ++pos) impool.push_back(p.im_f[perm[pos]]);
                d->h_m3d_f.push_back(it);
                d->m3d_smem_f = std::max(d->m3d_smem_f, (size_t)2 * p.nr * sizeof(float2));
            }
            f_len += (size_t)M;
            bytes_pts += 64.0 * M + (m3 ? 4.0 * p.nr : 4.0);
            d->pts_bytes[1] += 64.0 * M + (m3 ? 4.0 * p.nr : 4.0);
        }
    }
    if (s_len >= (1ull << 31) || f_len >= (1ull << 31)) fail("ax3d::finalize || field arrays exceed 2^31 complex entries per GPU");
    d->ns = so.size();
    d->nf = fo.size();
    d->s_len = s_len;
    d->f_len = f_len;
    for (int k = 0; k < 4; ++k) {
        d->s_field[k].alloc(s_len + 2);   // + 2: the in-kernel Newmark loads 16-byte aligned supersets of a point's block
        d->s_field[k].zero();
        d->f_field[k].alloc(f_len);
        d->f_field[k].zero();
    }
    d->s_off.upload(so); d->f_off.upload(fo);
    d->s_nu.upload(snu); d->f_nu.upload(fnu);
    d->s_nr.upload(snr); d->f_nr.upload(fnr);
    d->s_row_point.upload(srp); d->f_row_point.upload(frp);
    d->s_row_start.upload(srs); d->f_row_start.upload(frs);
    d->s_flags.upload(sfl); d->f_flags.upload(ffl);
    d->s_invmass.upload(sim); d->f_invmass.upload(fim);
    d->s_tab = PointTab{d->s_off.p, d->s_nu.p, d->s_nr.p, d->s_flags.p, d->s_invmass.p, d->s_row_point.p, d->s_row_start.p, (int)srp.size()};
    d->f_tab = PointTab{d->f_off.p, d->f_nu.p, d->f_nr.p, d->f_flags.p, d->f_invmass.p, d->f_row_point.p, d->f_row_start.p, (int)frp.size()};
    d->alg_bytes[0] = bytes_pts;

    // ---------------- elements: classify, sort 3D classes by Nr (descending) for load balance / chunking
    d->is_halo_point.assign(d->points.size(), 0);
    for (const auto &nb : d->neigh_pts)
        for (int t : nb) {
            if (t < 0 || t >= (int)d->points.size()) fail("Domain::setMessaging || invalid point tag");
            d->is_halo_point[t] = 1;
        }
    std::vector<int> order[NCLS];
    for (size_t e = 0; e < d->elems.size(); ++e) {
        HElem &E = d->elems[e];
        for (int i = 0; i < AX_NPE; ++i) E.bnd = E.bnd || d->is_halo_point[E.pt[i]];
        const bool is3d = E.rows > 1;
        E.cls = E.fluid ? (is3d ? CLS_F3D : CLS_F1D) : (is3d ? CLS_S3D : CLS_S1D);
        order[E.cls].push_back((int)e);
    }
    std::vector<float> geom, coef, attpar;
    size_t att1d_len = 0, att3d_len = 0;
    double bytes_el = 0;
    const char *env_sc = getenv("AX3D_SCRATCH_MB");
    const size_t scratch_cap = (size_t)((env_sc ? atof(env_sc) : 256.0) * 1024.0 * 1024.0 / sizeof(float2));
    size_t scratch_need = 0;
    const char *env_pk = getenv("AX3D_PACK_SMALL");
    const bool pack_small = env_pk ? atoi(env_pk) != 0 : true;
    const char *env_nf = getenv("AX3D_NO_FUSED");
    const bool use_fused = !(env_nf && atoi(env_nf) != 0);
    {
        // In-kernel Newmark (fused.cuh) is opt-in (AX3D_NW=1).  Measured on B200 (profiles/r2_nw_ab.md): the element kernel is
        // issue-bound, so the Newmark warps' polling and update instructions slow it down by more than the stand-alone,
        // HBM-bound k_newmark_solid costs (cfg4: 0.934 ms per step with, 0.820 ms without).
        const char *env_nw = getenv("AX3D_NW");
        d->nw_allowed = (env_nw && atoi(env_nw) != 0) && so.size() < (1u << 24);
    }

    for (int c = 0; c < NCLS; ++c) {
        const bool fluid = (c == CLS_F1D || c == CLS_F3D), is3d = (c == CLS_S3D || c == CLS_F3D);
        const int npair = fluid ? 2 : 3;
        std::vector<int> w_elem, w_a0;
        std::vector<FftItem> fitems;
        Chunk ch{c, 0, 0, 0, 0, 0, 5, 256, 0, 0.0};
        int cls_np = 0;   // points per k_fft3d_v2 CTA for this class: fixed by its largest split element
        size_t ch_scratch = 0;
        auto close_chunk = [&]() {
            if (ch.w_count > 0) d->chunks.push_back(ch);
            ch = Chunk{c, (int)w_elem.size(), 0, (int)fitems.size(), 0, 0, 5, 256, 0, 0.0};
            ch_scratch = 0;
        };
        FusedLaunch fl{};
        fl.cls = c;
        // tile region of the one resident CTA per SM (the solid launch keeps room for its Newmark warps' stages behind it)
        const size_t fused_lim = (size_t)AX_FUSED_DYN_MAX - ((c == CLS_S3D && d->nw_allowed) ? (size_t)AX_NWW * NW_WARP_SMEM + 8 : 0);
        fl.tile_cap = (int)((fused_lim / sizeof(float2)) & ~(size_t)1);
        // per-element shared-memory plan (fused.cuh): the smallest number of row groups ng in {1, 2, 3, 5} whose Z-form columns
        // fit next to the twiddle tables and a gather tile of >= 16 modes; Z sits at the end of the tile region, the twiddle
        // tables below it, the gather tile U = [0, twoff) in front
        auto plan_fused_element = [&](int N, int M, int stw_len, ElemDesc &D) -> bool {
            const int nc = fluid ? 1 : 3, us = nc * AX_NPE;
            // ng = 10 (partial-row passes, Nr up to ~2700) is opt-in (AX3D_PARTIAL_ROWS=1): on cfg5 (8 B200s) the split pipeline
            // with one point per CTA is still ~7 % faster for the Nr > ~1700 elements (summed element time 51.2 vs 54.6 ms)
            static const bool partial_rows = getenv("AX3D_PARTIAL_ROWS") && atoi(getenv("AX3D_PARTIAL_ROWS")) != 0;
            static const int ngs[5] = {1, 2, 3, 5, 10};
            for (int ng : ngs) {
                if (ng == 10 && !partial_rows) continue;
                const long long zsz = (long long)npair * fused_np_max(ng) * fused_ldz(N);
                const long long zoff = ((long long)fl.tile_cap - zsz) & ~1ll;
                const long long twoff = (zoff - ((stw_len + 1) & ~1)) & ~1ll;
                if (twoff < (long long)us * 16) continue;
                const int cap = (int)((twoff / us) & ~15ll);     // modes per gather tile
                D.ng = ng;
                D.zoff = (int)zoff;
                D.twoff = (int)twoff;
                D.mt = std::min(cap, M);
                return true;
            }
            return false;
        };
        // 3D classes: the fused launch takes the tail of the class order.  In front: what cannot be fused (particle relabelling,
        // one GLL row larger than shared memory; split pipeline), Nr descending; then the fused elements -- boundary elements
        // first (their forces go to the neighbours while the interior ones are computed), each part by Nr descending (LPT)
        int first_fused = (int)order[c].size();
        if (is3d) {
            std::vector<char> fit(d->elems.size(), 0);
            for (int e : order[c]) {
                const HElem &E = d->elems[e];
                ElemDesc tmp;
                const int pid = get_plan(d, E.nr);
                fit[e] = use_fused && E.prt_rows == 0 && plan_fused_element(E.nr, E.nu + 1, d->h_plans[pid].stw_len, tmp);
            }
            std::stable_sort(order[c].begin(), order[c].end(), [&](int a, int b) {
                const HElem &A = d->elems[a], &B = d->elems[b];
                if (fit[a] != fit[b]) return fit[a] < fit[b];
                if (!fit[a] && (A.prt_rows > 0) != (B.prt_rows > 0)) return A.prt_rows > 0;
                if (fit[a] && A.bnd != B.bnd) return A.bnd;
                return A.nr > B.nr;
            });
            first_fused = 0;
            while (first_fused < (int)order[c].size() && !fit[order[c][first_fused]]) ++first_fused;
        }
        auto close_fused = [&]() {
            if (fl.count > 0) {
                // work items: one per (element, row-group pass) -- passes are independent, and a pass is a finer scheduling
                // quantum than a large element -- or one per element (code 7 = all passes) when the in-kernel Newmark counts
                // arrivals per element.  Boundary items first, then by estimated cost (Nr log Nr x rows of the pass), largest first.
                struct It { int code; bool bnd; double cost; };
                std::vector<It> its;
                // Pass-level items cost a few per cent on their own (consecutive passes of one element no longer share a CTA's
                // warm descriptor / twiddles / L2 lines; B200 cfg4: 0.848 -> 0.875 ms per step), so they are used only where the
                // element-level queue would lose more to the quantisation: LPT makespan of the element costs over num_sm
                // bins more than 2.5 % above the ideal.  AX3D_PASS_ITEMS=0 / 1 forces the choice.
                auto lpt_excess = [&]() {
                    std::vector<double> cost;
                    double tot = 0;
                    for (int k = fl.first; k < fl.first + fl.count; ++k) {
                        const ElemDesc &D = d->h_desc[c][k];
                        // measured shape of the element cost (cost_model.json): ~ Nr log2 Nr, steeper once the element needs several passes
                        const double x = (double)D.nr * std::log2((double)std::max(D.nr, 2)) * (1.0 + 0.15 * (D.ng - 1)) + 150.0;
                        cost.push_back(x);
                        tot += x;
                    }
                    std::sort(cost.begin(), cost.end(), std::greater<double>());
                    std::vector<double> bin((size_t)d->num_sm, 0.0);
                    for (double x : cost) *std::min_element(bin.begin(), bin.end()) += x;
                    return *std::max_element(bin.begin(), bin.end()) / (tot / d->num_sm) - 1.0;
                };
                const char *env_pi = getenv("AX3D_PASS_ITEMS");
                const bool pass_items = env_pi ? atoi(env_pi) != 0 : lpt_excess() > 0.025;
                const bool per_element = (c == CLS_S3D && d->nw_allowed) || !pass_items;
                for (int k = fl.first; k < fl.first + fl.count; ++k) {
                    const ElemDesc &D = d->h_desc[c][k];
                    const double full = (double)D.nr * std::log2((double)std::max(D.nr, 2)) + 64.0;
                    if (per_element) { its.push_back(It{((k - fl.first) << 4) | AX_ITEM_ALL, D.bnd != 0, full}); continue; }
                    for (int g = 0; g < D.ng; ++g) {
                        int gp0, gnp;
                        fused_group(D.ng, g, gp0, gnp);
                        its.push_back(It{((k - fl.first) << 4) | g, D.bnd != 0, full * gnp / 25.0 + 16.0});
                    }
                }
                std::stable_sort(its.begin(), its.end(), [](const It &a, const It &b) {
                    if (a.bnd != b.bnd) return a.bnd;
                    return a.cost > b.cost;
                });
                std::vector<int> codes;
                fl.nb_items = 0;
                for (const It &x : its) { codes.push_back(x.code); fl.nb_items += x.bnd ? 1 : 0; }
                fl.nitems = (int)codes.size();
                d->fused_items[c].upload(codes);
                fl.smem = (size_t)AX_FUSED_DYN_MAX;   // tile region + (solid launch) the Newmark warps' stage buffers
                fl.grid = std::min(fl.nitems, d->num_sm);
                d->fused.push_back(fl);
            }
        };
        for (size_t k = 0; k < order[c].size(); ++k) {
            HElem &E = d->elems[order[c][k]];
            E.idx = (int)k;
            ElemDesc D;
            memset(&D, 0, sizeof(D));
            D.nr = E.nr;
            D.nu = E.nu;
            D.nyq = (E.nr % 2 == 0) ? 1 : 0;
            D.axial = E.axial ? 1 : 0;
            D.law = E.law;
            D.prt = E.prt_rows > 0 ? 1 : 0;
            D.bnd = E.bnd ? 1 : 0;
            if (D.prt) d->cls_prt[c] = true;
            D.tiso = (!fluid && (E.law != AX3D_ISO || D.prt)) ? 1 : 0;   // SolidElement.cpp:21: mInTIso = mHasPRT || needTIso
            D.is3d = is3d ? 1 : 0;
            D.att_kind = E.att.kind;
            D.nsls = E.att.nsls;
            D.do_kappa = E.att.do_kappa;
            const int M = E.nu + 1, N = E.nr;
            for (int i = 0; i < AX_NPE; ++i) {
                const HPoint &p = d->points[E.pt[i]];
                D.pt_off[i] = (unsigned)(fluid ? p.f_off : p.s_off);
                D.pt_stride[i] = p.nu + 1;
                D.pt_nlive[i] = p.nu - ((p.nr % 2 == 0) ? 1 : 0) + 1;
            }
            D.geom_off = (long long)geom.size();
            for (int i = 0; i < 5 * AX_NPE; ++i) geom.push_back((float)E.geom[i]);
            {                                  // fluid elements rotate only with PRT (FluidElement.cpp:21); every element
                                               // keeps its trig for the strain / curl read-back (forceTIso, SolidElement.cpp:347-352)
                D.trig_off = (long long)geom.size();
                for (int i = 0; i < AX_NPE; ++i) geom.push_back((float)sin(E.theta[i]));
                for (int i = 0; i < AX_NPE; ++i) geom.push_back((float)cos(E.theta[i]));
                for (int i = 0; i < AX_NPE; ++i) geom.push_back((float)sin(2.0 * E.theta[i]));
                for (int i = 0; i < AX_NPE; ++i) geom.push_back((float)cos(2.0 * E.theta[i]));
            }
            D.coef_off = (long long)coef.size();
            const std::vector<int> *perm = nullptr;
            if (is3d) {
                D.plan_id = get_plan(d, N);
                perm = &d->h_perm[D.plan_id];
                for (int kc = 0; kc < E.ncoef; ++kc)
                    for (int p = 0; p < AX_NPE; ++p)
                        for (int pos = 0; pos < N; ++pos) coef.push_back(E.coef[((size_t)kc * AX_NPE + p) * N + (*perm)[pos]]);
            } else {
                coef.insert(coef.end(), E.coef.begin(), E.coef.end());
            }
            if (D.prt) {   // PRT_1D: [4][25]; PRT_3D: [4][25][Nr] in the plan's digit-reversed phi order, next to the moduli
                if ((E.prt_rows == 1) == is3d) fail(std::string(fluid ? "FluidElement::FluidElement" : "SolidElement::SolidElement") +
                                                    " || Particle Relabelling and Elasticity are generated in different spaces.");
                D.prt_off = (long long)coef.size();
                if (is3d) {
                    for (int kc = 0; kc < 4; ++kc)
                        for (int p = 0; p < AX_NPE; ++p)
                            for (int pos = 0; pos < N; ++pos) coef.push_back(E.prtX[((size_t)kc * AX_NPE + p) * N + (*perm)[pos]]);
                } else {
                    coef.insert(coef.end(), E.prtX.begin(), E.prtX.end());
                }
            }
            if (E.att.kind != ATT_NONE) {
                const int P = E.att.kind == ATT_CG4 ? 4 : AX_NPE;
                D.att_par_off = (long long)attpar.size();
                attpar.insert(attpar.end(), E.att.alpha.begin(), E.att.alpha.end());
                attpar.insert(attpar.end(), E.att.beta.begin(), E.att.beta.end());
                attpar.insert(attpar.end(), E.att.gamma.begin(), E.att.gamma.end());
                const int rows = is3d ? N : 1;
                // three * dkappa, dmu, two * dmu in Real (Attenuation3D_Full.cpp:11-12)
                for (int which = 0; which < 3; ++which)
                    for (int q = 0; q < P; ++q)
                        for (int pos = 0; pos < rows; ++pos) {
                            const int j = is3d ? (*perm)[pos] : 0;
                            const float dk = E.att.dkappa[(size_t)q * rows + j], dm = E.att.dmu[(size_t)q * rows + j];
                            attpar.push_back(which == 0 ? 3.f * dk : which == 1 ? dm : 2.f * dm);
                        }
                const size_t cells = (size_t)(E.att.nsls + 1) * 6 * P * (is3d ? N : M);
                if (is3d) { D.att_state_off = (long long)att3d_len; att3d_len += cells; }
                else { D.att_state_off = (long long)att1d_len; att1d_len += cells; }
            }
            // algorithmic bytes
            {
                const double nin = fluid ? 1 : 3;
                double b = 2.0 * 200.0 * nin * M + 500.0 + (D.tiso ? 400.0 : 0.0);
                b += (is3d ? 100.0 * N : 100.0) * (E.ncoef + (D.prt ? 4 : 0));
                if (E.att.kind != ATT_NONE) {
                    const int P = E.att.kind == ATT_CG4 ? 4 : AX_NPE;
                    const double R = is3d ? 4.0 * N : 8.0 * M;
                    b += 2.0 * (E.att.nsls + 1) * 6 * P * R + 2.0 * P * (is3d ? 4.0 * N : 4.0);
                }
                bytes_el += b;
                E.alg_b = b;
            }
            D.bucket = -1;
            D.mt = M;
            if (is3d) {
                // fused one-CTA-per-element kernel (fused.cuh) when one row group of the element fits in shared memory
                D.plan_id = get_plan(d, N);
                const int stw_len = d->h_plans[D.plan_id].stw_len;
                const bool can_fuse = (int)k >= first_fused;
                if (can_fuse) {
                    if (!plan_fused_element(N, M, stw_len, D)) fail("ax3d::fused || internal: element does not fit its launch");
                    D.bucket = 0;
                    if (fl.count == 0) fl.first = (int)k;
                    fl.count++;
                    fl.alg_b += E.alg_b;
                    if (c == CLS_S3D && E.bnd) d->n_bnd_fused++;
                } else {
                    D.mt = M;
                    const int enp = (D.prt && !fluid) ? 5 : npair;   // Z-form pairs per point of this element
                    // one k_fft3d_v2 instance per chunk: same pair count and same points per CTA (5, or 1 when five points of
                    // this Nr do not fit in shared memory; elements come in descending Nr)
                    const int my_np = ((size_t)enp * 5 * fused_ldz(N) + 2 * (size_t)N) * sizeof(float2) <= (size_t)220 * 1024 ? 5 : 1;
                    if (ch.w_count > 0 && ((D.prt != 0) != (ch.prt != 0) || my_np != ch.fft_np)) close_chunk();
                    {   // elements come in descending Nr: start a new chunk where one more k_fft3d_v2 CTA per SM becomes resident
                        // (its 256-thread instance is compiled for AX_FFT_MIN_CTAS CTAs per SM), instead of running the whole class
                        // at the shared-memory footprint of its largest element
                        const size_t mine = ((size_t)enp * my_np * fused_ldz(N) + (size_t)((stw_len + 1) & ~1)) * sizeof(float2);
                        if (ch.w_count > 0 && fft_threads(mine) != fft_threads(ch.fft_smem)) close_chunk();
                    }
                    ch.prt = D.prt ? 1 : 0;
                    (void)cls_np;
                    D.ppb = my_np;
                    const size_t need_sc = (size_t)enp * AX_NPE * N;
                    if (need_sc > scratch_cap && ch.w_count > 0) close_chunk();
                    if (ch_scratch + need_sc > scratch_cap && ch.w_count > 0) close_chunk();
                    D.scratch_off = (long long)ch_scratch;
                    ch_scratch += need_sc;
                    scratch_need = std::max(scratch_need, ch_scratch);
                    for (int a0 = 0; a0 < M; a0 += AX_TILE) { w_elem.push_back((int)k); w_a0.push_back(a0); ch.w_count++; }
                    for (int p0 = 0; p0 < AX_NPE; p0 += D.ppb) { fitems.push_back(FftItem{(int)k, p0}); ch.f_count++; }
                    ch.fft_smem = std::max(ch.fft_smem, ((size_t)enp * D.ppb * fused_ldz(N) + (size_t)((stw_len + 1) & ~1)) * sizeof(float2));
                    ch.fft_np = D.ppb;
                    ch.fft_nt = fft_threads(ch.fft_smem);
                    ch.alg_b += E.alg_b;
                }
            } else {
                // 1D classes, small expansions (cfg1: M = 3): the 16 mode lanes of a CTA take up to 16 / Mpad consecutive
                // elements of the same M (Mpad = 4 or 8); such an item carries a0 = -(number of elements)
                const int Mpad = M <= 4 ? 4 : M <= 8 ? 8 : 0;
                if (pack_small && Mpad && !w_elem.empty() && w_a0.back() < 0 && -w_a0.back() < AX_TILE / Mpad &&
                    d->h_desc[c][w_elem.back()].nu + 1 == M && w_elem.back() - w_a0.back() == (int)k) {
                    w_a0.back() -= 1;                         // extend the open pack by this element
                } else if (pack_small && Mpad) {
                    w_elem.push_back((int)k);
                    w_a0.push_back(-1);
                } else {
                    for (int a0 = 0; a0 < M; a0 += AX_TILE) { w_elem.push_back((int)k); w_a0.push_back(a0); }
                }
            }
            d->h_desc[c].push_back(D);
            d->h_alg_b[c].push_back(E.alg_b);
            if (!is3d) d->cls_bytes[c] += E.alg_b;
        }
        if (is3d) { close_chunk(); close_fused(); }
        d->desc[c].upload(d->h_desc[c]);
        d->w_elem[c].upload(w_elem);
        d->w_a0[c].upload(w_a0);
        d->n_work[c] = (int)w_elem.size();
        d->fft_items[c].upload(fitems);
    }
    // ---------------- in-kernel Newmark: plain points, arrival counts, special-row table
    {
        const size_t ns = d->ns;
        std::vector<int> need(ns, 0);
        std::vector<char> ok(ns, 0);
        FusedLaunch *fs = nullptr;
        for (FusedLaunch &f : d->fused)
            if (f.cls == CLS_S3D) fs = &f;
        if (d->nw_allowed && fs) {
            for (const HPoint &p : d->points)
                if (p.kind == 0 && !p.axial && p.im_s.size() == 1 && !p.ocean) ok[p.s_idx] = 1;
            for (const auto &nb : d->neigh_pts)
                for (int t : nb)
                    if (t >= 0 && t < (int)d->points.size() && d->points[t].s_idx >= 0) ok[d->points[t].s_idx] = 0;
            for (const HSource &sc : d->sources)
                for (int i = 0; i < AX_NPE; ++i) {
                    const HPoint &p = d->points[d->elems[sc.elem].pt[i]];
                    if (p.s_idx >= 0) ok[p.s_idx] = 0;
                }
            for (const HElem &E : d->elems) {
                if (E.fluid) continue;
                const bool in_launch = E.cls == CLS_S3D && d->h_desc[CLS_S3D][E.idx].bucket == 0;
                for (int i = 0; i < AX_NPE; ++i) {
                    const int sp = d->points[E.pt[i]].s_idx;
                    if (sp < 0) continue;
                    if (in_launch) need[sp]++;
                    else ok[sp] = 0;
                }
            }
        }
        std::vector<int> srp_sp, srs_sp(ns, 0);
        int n_plain = 0;
        for (size_t sp = 0; sp < ns; ++sp) {
            if (ok[sp] && need[sp] > 0 && need[sp] < 128) { n_plain++; continue; }
            ok[sp] = 0;
            srs_sp[sp] = (int)srp_sp.size();
            for (int a = 0; a <= snu[sp]; ++a) srp_sp.push_back((int)sp);
        }
        if (n_plain * 8 < (int)ns) {   // not worth a specialised launch: everything stays with k_newmark_solid
            n_plain = 0;
            std::fill(ok.begin(), ok.end(), 0);
        }
        d->n_plain = n_plain;
        for (const HElem &E : d->elems)
            if (!E.fluid && E.cls == CLS_S3D && d->h_desc[CLS_S3D][E.idx].bucket == 0) d->dom_bytes[0] += E.alg_b;
        for (size_t sp = 0; sp < ns; ++sp)
            if (ok[sp]) d->dom_bytes[1] += 192.0 * (snu[sp] + 1) + 4.0;
        for (ElemDesc &D : d->h_desc[CLS_S3D]) {
            for (int i = 0; i < AX_NPE; ++i) D.pt_nw[i] = -1;
        }
        if (n_plain > 0) {
            for (const HElem &E : d->elems) {
                if (E.fluid || E.cls != CLS_S3D) continue;
                ElemDesc &D = d->h_desc[CLS_S3D][E.idx];
                if (D.bucket != 0) continue;
                for (int i = 0; i < AX_NPE; ++i) {
                    const int sp = d->points[E.pt[i]].s_idx;
                    // a point listed twice in one element would arrive twice: count per listing (need counted listings)
                    if (sp >= 0 && ok[sp]) D.pt_nw[i] = sp | (need[sp] << 24);
                }
            }
            d->desc[CLS_S3D].upload(d->h_desc[CLS_S3D]);
            fs->nww = AX_NWW;                                // their stages sit behind the tile region (plan_fused_element left the room)
            d->nw_cnt.alloc(ns);
            d->nw_cnt.zero();
            std::vector<int> q((size_t)n_plain, -1);
            d->nw_queue.upload(q);
            d->nw_ctl.alloc(4);
            d->nw_ctl.zero();
            if (getenv("AX3D_NW_DEBUG")) { d->nw_dbg.alloc(4 * 256); d->nw_dbg.zero(); }
            d->s_row_point_sp.upload(srp_sp);
            d->s_row_start_sp.upload(srs_sp);
            d->s_tab_sp = d->s_tab;
            d->s_tab_sp.row_point = d->s_row_point_sp.p;
            d->s_tab_sp.row_start = d->s_row_start_sp.p;
            d->s_tab_sp.nrows = (int)srp_sp.size();
        }
    }
    d->geom.upload(geom);
    d->coef.upload(coef);
    d->attpar.upload(attpar);
    d->attstate1d.alloc(att1d_len); d->attstate1d.zero();
    d->attstate3d.alloc(att3d_len); d->attstate3d.zero();
    d->scratch.alloc(scratch_need);
    d->alg_bytes[1] = bytes_el;

    // ---------------- sources: flatten to (offset, value)
    {
        std::vector<unsigned> off;
        std::vector<float2> val;
        for (const HSource &s : d->sources) {
            const HElem &E = d->elems[s.elem];
            size_t pos = 0;
            for (int i = 0; i < AX_NPE; ++i) {
                const HPoint &p = d->points[E.pt[i]];
                if (p.s_idx < 0) fail("Point::addToStiff || Incompatible point type.");
                const int nrow = s.nrow[i], keep = std::min(nrow, p.nu + 1);
                for (int c = 0; c < 3; ++c)
                    for (int a = 0; a < keep; ++a) {
                        off.push_back((unsigned)(p.s_off + (size_t)c * (p.nu + 1) + a));
                        val.push_back(make_float2(s.force[2 * (pos + (size_t)c * nrow + a)], s.force[2 * (pos + (size_t)c * nrow + a) + 1]));
                    }
                pos += (size_t)3 * nrow;
            }
        }
        d->n_src = (int)off.size();
        d->src_off.upload(off);
        d->src_val.upload(val);
    }
    // ---------------- solid-fluid points: the whole 1D-coupled table, and the same split into points on / off the halo
    {
        struct Tab { std::vector<unsigned> a, b; std::vector<int> nu, rp, rs; std::vector<float> cpl; };
        Tab all, onh, off;
        std::vector<float> pool;
        auto push = [](Tab &t, const HPoint &p) {
            const int q = (int)t.a.size();
            t.a.push_back((unsigned)p.s_off);
            t.b.push_back((unsigned)p.f_off);
            t.nu.push_back(p.nu);
            t.rs.push_back((int)t.rp.size());
            for (int k = 0; k <= p.nu; ++k) t.rp.push_back(q);
            t.cpl.push_back(p.n_un[0]); t.cpl.push_back(p.n_un[2]);
            t.cpl.push_back(p.n_as[0]); t.cpl.push_back(p.n_as[2]);
        };
        for (size_t tg = 0; tg < d->points.size(); ++tg) {
            const HPoint &p = d->points[tg];
            if (p.kind != 2) continue;
            if (p.n_sf == 1) {
                push(all, p);
                push(d->is_halo_point[tg] ? onh : off, p);
            } else {
                if (d->is_halo_point[tg]) d->halo_sf3d = true;
                SF3DItem it;
                it.s_off = (unsigned)p.s_off;
                it.f_off = (unsigned)p.f_off;
                it.nu = p.nu;
                it.plan_id = get_plan(d, p.nr);
                it.n_off = (long long)pool.size();
                const std::vector<int> &perm = d->h_perm[it.plan_id];
                for (int c = 0; c < 3; ++c)
                    for (int pos = 0; pos < p.nr; ++pos) pool.push_back(p.n_un[(size_t)c * p.nr + perm[pos]]);
                for (int c = 0; c < 3; ++c)
                    for (int pos = 0; pos < p.nr; ++pos) pool.push_back(p.n_as[(size_t)c * p.nr + perm[pos]]);
                d->h_sf3d.push_back(it);
                d->sf3d_smem = std::max(d->sf3d_smem, (size_t)3 * p.nr * sizeof(float2));
            }
        }
        d->sf_s_off.upload(all.a); d->sf_f_off.upload(all.b); d->sf_nu.upload(all.nu);
        d->sf_row_point.upload(all.rp); d->sf_row_start.upload(all.rs); d->sf_cpl.upload(all.cpl);
        d->sf_tab = SFTab{d->sf_s_off.p, d->sf_f_off.p, d->sf_nu.p, d->sf_cpl.p, d->sf_row_point.p, d->sf_row_start.p, (int)all.rp.size()};
        d->sfh_s_off.upload(onh.a); d->sfh_f_off.upload(onh.b); d->sfh_nu.upload(onh.nu);
        d->sfh_row_point.upload(onh.rp); d->sfh_row_start.upload(onh.rs); d->sfh_cpl.upload(onh.cpl);
        d->sf_tab_halo = SFTab{d->sfh_s_off.p, d->sfh_f_off.p, d->sfh_nu.p, d->sfh_cpl.p, d->sfh_row_point.p, d->sfh_row_start.p, (int)onh.rp.size()};
        d->sfr_s_off.upload(off.a); d->sfr_f_off.upload(off.b); d->sfr_nu.upload(off.nu);
        d->sfr_row_point.upload(off.rp); d->sfr_row_start.upload(off.rs); d->sfr_cpl.upload(off.cpl);
        d->sf_tab_rest = SFTab{d->sfr_s_off.p, d->sfr_f_off.p, d->sfr_nu.p, d->sfr_cpl.p, d->sfr_row_point.p, d->sfr_row_start.p, (int)off.rp.size()};
        d->sf3d.upload(d->h_sf3d);
        d->sf3d_pool.upload(pool);
    }
    // ---------------- Mass3D + plans
    d->m3d_s.upload(d->h_m3d_s);
    d->m3d_f.upload(d->h_m3d_f);
    d->oc1d.upload(d->h_oc1d);
    for (const HPoint &p : d->points)
        if (p.ocean == 1) d->oc1d_max_m = std::max(d->oc1d_max_m, p.nu + 1);
    d->impool.upload(impool);
    {
        std::vector<float2> tw;
        for (FftPlan &pl : d->h_plans) {
            pl.tw_off = (int)tw.size();
            for (int k = 0; k < pl.N; ++k) {
                const double a = 2.0 * M_PI * k / pl.N;
                tw.push_back(make_float2((float)cos(a), (float)sin(a)));
            }
        }
        d->twpool.upload(tw);
        d->stwpool.upload(d->h_stw);
        d->plans.upload(d->h_plans);
    }
    // ---------------- halo index lists
    {
        std::vector<unsigned> idx;
        d->neigh_begin.clear();
        double hb = 0;
        for (size_t n = 0; n < d->neigh_pts.size(); ++n) {
            d->neigh_begin.push_back(idx.size());
            for (int t : d->neigh_pts[n]) {
                if (t < 0 || t >= (int)d->points.size()) fail("Domain::setMessaging || invalid point tag");
                const HPoint &p = d->points[t];
                // SolidFluidPoint::feedBuffer: solid block then fluid block (SolidFluidPoint.cpp:76-79)
                if (p.s_idx >= 0)
                    for (int k = 0; k < 3 * (p.nu + 1); ++k) idx.push_back((unsigned)(p.s_off + k));
                if (p.f_idx >= 0)
                    for (int k = 0; k < p.nu + 1; ++k) idx.push_back((unsigned)(p.f_off + k) | 0x80000000u);
            }
        }
        d->neigh_begin.push_back(idx.size());
        hb = 16.0 * idx.size();
        d->halo_idx.upload(idx);
        d->halo_send.alloc(idx.size());
        d->halo_recv.alloc(idx.size());
        d->halo_win.alloc(2 * idx.size() + (d->neigh_rank.size() + 1) / 2 + 1);   // window + counters: one IPC handle maps both
        d->halo_win.zero();
        d->halo_step.alloc(1);
        d->halo_step.zero();
        d->alg_bytes[2] = hb;
    }
    d->bad_flag.alloc(2);   // [0] non-finite displacement (checkStability), [1] peer-halo wait timed out
    d->bad_flag.zero();
    d->stf_dev.alloc(2);   // {source factor (float), record-ring slot (int)} of the step being enqueued
    d->stf_dev.zero();
    CK(cudaMallocHost(&d->stf_pinned, 2 * STF_RING * sizeof(float)));
    {
        const char *g = getenv("AX3D_NO_GRAPH");
        d->use_graph = !(g && atoi(g) != 0);
    }
    {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo = least priority (numerically greatest)
        CK(cudaStreamCreateWithPriority(&d->stream, cudaStreamNonBlocking, hi));
        CK(cudaStreamCreateWithPriority(&d->stream2, cudaStreamNonBlocking, lo));
        CK(cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming));
        const char *e = getenv("AX3D_DUAL");
        d->dual_mode = e ? atoi(e) : AX_DUAL_DEFAULT;
        d->dual_chain = d->dual_mode != 0 && d->nf > 0 && d->ns > 0;
    }
    CK(cudaEventCreate(&d->ev0));
    CK(cudaEventCreate(&d->ev1));
    CK(cudaEventCreate(&d->ev2));
    CK(cudaEventCreate(&d->ev3));
    // opt in to large dynamic shared memory
    for (const Chunk &ch : d->chunks) {
        if (ch.fft_smem > (size_t)224 * 1024) fail("ax3d::finalize || Nr too large for the FFT stage (needs > 224 KB shared memory per point)");
        CK(cudaFuncSetAttribute((const void *)fft_kernel(ch), cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    }
    {
        std::vector<unsigned> w;
        for (const FusedLaunch &f : d->fused) {
            set_fused_smem(d->device, f);
            w.push_back((unsigned)f.grid);
            w.push_back(0u);
        }
        d->fused_work.upload(w);
    }
    if (d->m3d_smem_s > 48 * 1024) CK(cudaFuncSetAttribute(k_mass3d<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->m3d_smem_s));
    if (d->m3d_smem_f > 48 * 1024) CK(cudaFuncSetAttribute(k_mass3d<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->m3d_smem_f));
    if (d->sf3d_smem > 48 * 1024) CK(cudaFuncSetAttribute(k_sf_couple3d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->sf3d_smem));
#ifdef AX3D_WITH_NCCL
    if (d->nproc > 1 && d->have_uid) {
        ncclUniqueId id;
        memcpy(&id, d->uid, sizeof(id) < 128 ? sizeof(id) : 128);
        ncclResult_t r = ncclCommInitRank(&d->comm, d->nproc, id, d->rank);
        if (r != ncclSuccess) fail(std::string("XMPI::initialize || ncclCommInitRank: ") + ncclGetErrorString(r));
    }
#endif
    CK(cudaDeviceSynchronize());
    d->finalized = true;
}

// ------------------------------------------------------------------------------------------ step verbs
struct TimerScope {
    ax3d_domain *d;
    int slot;
    TimerScope(ax3d_domain *d_, int s) : d(d_), slot(s) {
        if (d->timers) cudaEventRecord(d->ev0, d->stream);
    }
    ~TimerScope() {
        if (d->timers) {
            cudaEventRecord(d->ev1, d->stream);
            cudaEventSynchronize(d->ev1);
            float ms = 0;
            cudaEventElapsedTime(&ms, d->ev0, d->ev1);
            d->timer_ms[slot] += ms;
        }
    }
};

// one kernel (or one back-to-back group of kernels) under its own event pair, nested inside the family's TimerScope
struct KTimer {
    ax3d_domain *d;
    const char *name;
    double bytes;
    KTimer(ax3d_domain *d_, const char *name_, double bytes_) : d(d_), name(name_), bytes(bytes_) {
        if (d->timers) cudaEventRecord(d->ev2, d->stream);
    }
    ~KTimer() {
        if (d->timers) {
            cudaEventRecord(d->ev3, d->stream);
            cudaEventSynchronize(d->ev3);
            float ms = 0;
            cudaEventElapsedTime(&ms, d->ev2, d->ev3);
            ax3d_domain::KStat &k = d->kstats[name];
            k.ms += ms;
            k.n += 1;
            k.bytes += bytes;
        }
    }
};

static inline int nblk(size_t n, int b) { return (int)((n + b - 1) / b); }

// special_only: the plain solid points were already advanced by the previous step's element kernel (fused.cuh)
// which: 1 = solid points, 2 = fluid points, 3 = both
static void update_newmark(ax3d_domain *d, double dt, bool special_only = false, int which = 3) {
    TimerScope ts(d, 0);
    const double half_dt = 0.5 * dt, half_dt_dt = half_dt * dt;   // SolidPoint.cpp:31-32 (double, then cast to Real)
    if ((which & 1) && !d->h_m3d_s.empty()) {
        k_mass3d<3><<<(int)d->h_m3d_s.size(), 128, d->m3d_smem_s, d->stream>>>(d->s_tab, d->m3d_s.p, d->plans.p, d->twpool.p, d->impool.p,
                                                                              d->s_field[AX3D_STIFF].p);
        d->launches++;
    }
    if ((which & 2) && !d->h_m3d_f.empty()) {
        k_mass3d<1><<<(int)d->h_m3d_f.size(), 128, d->m3d_smem_f, d->stream>>>(d->f_tab, d->m3d_f.p, d->plans.p, d->twpool.p, d->impool.p,
                                                                              d->f_field[AX3D_STIFF].p);
        d->launches++;
    }
    if ((which & 1) && !d->h_oc1d.empty()) {
        const size_t n = d->h_oc1d.size() * (size_t)d->oc1d_max_m;
        k_mass_ocean1d<<<nblk(n, 128), 128, 0, d->stream>>>(d->s_tab, (int)d->h_oc1d.size(), d->oc1d.p, d->oc1d_max_m, d->s_field[AX3D_STIFF].p);
        d->launches++;
    }
    const PointTab &stab = special_only ? d->s_tab_sp : d->s_tab;
    if ((which & 1) && stab.nrows) {
        KTimer kt(d, "k_newmark_solid", d->pts_bytes[0] - (special_only ? d->dom_bytes[1] : 0.0));
        k_newmark_solid<<<nblk(stab.nrows, 256), 256, 0, d->stream>>>(stab, d->s_field[0].p, d->s_field[1].p, d->s_field[2].p,
                                                                     d->s_field[3].p, (float)half_dt, (float)dt, (float)half_dt_dt);
        d->launches++;
    }
    if ((which & 2) && d->f_tab.nrows) {
        KTimer kt(d, "k_newmark_fluid", d->pts_bytes[1]);
        k_newmark_fluid<<<nblk(d->f_tab.nrows, 256), 256, 0, d->stream>>>(d->f_tab, d->f_field[0].p, d->f_field[1].p, d->f_field[2].p,
                                                                         d->f_field[3].p, (float)half_dt, (float)dt, (float)half_dt_dt);
        d->launches++;
    }
    CK(cudaGetLastError());
}

static void push_stf(ax3d_domain *d, float stf, int rec_slot = 0) {
    // a slot is reused only after STF_RING further steps; synchronise before wrapping around
    if (d->stf_slot == STF_RING) {
        CK(cudaStreamSynchronize(d->stream));
        d->stf_slot = 0;
    }
    float *h = d->stf_pinned + 2 * d->stf_slot;
    h[0] = stf;
    memcpy(h + 1, &rec_slot, sizeof(int));
    CK(cudaMemcpyAsync(d->stf_dev.p, h, 2 * sizeof(float), cudaMemcpyHostToDevice, d->stream));
    d->stf_slot++;
}

static size_t recf_smem(int N) { return (size_t)2 * 5 * fused_ldz(N) * sizeof(float2); }   // one GLL row: 2 Z-form pairs x 5 points

// device rows are grouped by element class; the caller's receiver order is restored on the host
static void unscramble_record(const ax3d_domain *d, const float *row, float *out) {
    int base = 0;
    for (int c = 0; c < NCLS; ++c) {
        for (int k = 0; k < d->nrec_c[c]; ++k)
            for (int cc = 0; cc < 3; ++cc) out[d->rec_where_c[c][k] * 3 + cc] = row[(base + k) * 3 + cc];
        base += d->nrec_c[c];
    }
}

// Domain::record -> PointwiseRecorder::record (Domain.cpp:207-220): one sample per registered receiver into row
// `*slot` of the device ring (slot == nullptr: row 0 of `out`).
// which: 1 = receivers in solid elements, 2 = in fluid elements (they read different displacement fields, which the dual-stream
// step updates on different streams), 3 = both
static void launch_record(ax3d_domain *d, float *out, const int *slot, int stride, int which = 3) {
    int row = 0;
    for (int c = 0; c < NCLS; ++c) {
        const int n = d->nrec_c[c];
        if (!n) continue;
        if (!(which & ((c == CLS_S1D || c == CLS_S3D) ? 1 : 2))) { row += n; continue; }
        if (c == CLS_S1D || c == CLS_S3D)
            k_ground_motion<<<n, AX_REC_NT, 0, d->stream>>>(d->desc[c].p, d->rec_items_c[c].p, d->rec_w_c[c].p, d->s_field[AX3D_DISPL].p,
                                                           out + (size_t)3 * row, slot, stride);
        else
            k_ground_motion_fluid<<<n, AX_RECF_NT, d->rec_fluid_smem, d->stream>>>(d->desc[c].p, d->rec_items_c[c].p, d->rec_w_c[c].p, d->plans.p,
                                                                                   d->twpool.p, d->geom.p, d->coef.p, d->f_field[AX3D_DISPL].p,
                                                                                   out + (size_t)3 * row, slot, stride);
        d->launches++;
        row += n;
    }
    CK(cudaGetLastError());
}

static void launch_source(ax3d_domain *d) {
    if (d->n_src) {
        k_source<<<nblk(d->n_src, 128), 128, 0, d->stream>>>(d->n_src, d->src_off.p, d->src_val.p, d->stf_dev.p, d->s_field[AX3D_STIFF].p);
        d->launches++;
        CK(cudaGetLastError());
    }
}

static void apply_source(ax3d_domain *d, float stf) {
    TimerScope ts(d, 2);
    if (d->n_src) push_stf(d, stf);
    launch_source(d);
}

// 512 threads = 128 registers/thread.  Measured on B200 (cfg2, elements family): 416 -> 0.270 ms, 448 -> 0.266, 512 -> 0.259,
// 640 (96 regs) -> 0.289, 768 (80 regs) -> 0.310: the kernel is not occupancy-limited.  With Newmark warps the CTA is
// 448 compute + 64 Newmark threads.
static int fused_nt(const FusedLaunch &f) {
    return f.nww ? AX_NW_NT + 32 * AX_NWW : AX_FUSED_NT;
}

typedef void (*fused_kernel_t)(const ElemDesc *, const int *, int, const FftPlan *, const float2 *, const float *, const float *, const float *,
                               float *, const float2 *, float2 *, int, unsigned *, const NwArgs, const HaloArgs);

static fused_kernel_t fused_kernel(const FusedLaunch &f) {
    const bool fluid = f.cls == CLS_F3D;
    if (fluid) return k_elem3d_fused<true, AX_FUSED_NT, 0>;
    return f.nww ? k_elem3d_fused<false, AX_NW_NT, AX_NWW> : k_elem3d_fused<false, AX_FUSED_NT, 0>;
}

// nw_on: this launch also advances the plain solid points to the next step (its dt is the step being integrated)
static void launch_fused(ax3d_domain *d, const FusedLaunch &f, int which, bool nw_on, double dt, bool put) {
    const int c = f.cls;
    const bool fluid = c == CLS_F3D;
    NwArgs nw;
    memset(&nw, 0, sizeof(nw));
    if (f.nww > 0 && nw_on) {
        const double half_dt = 0.5 * dt, half_dt_dt = half_dt * dt;
        nw.on = 1;
        nw.n_plain = d->n_plain;
        nw.cnt = d->nw_cnt.p;
        nw.queue = d->nw_queue.p;
        nw.ctl = d->nw_ctl.p;
        nw.off = d->s_off.p;
        nw.nu = d->s_nu.p;
        nw.nr = d->s_nr.p;
        nw.invmass = d->s_invmass.p;
        nw.displ = d->s_field[AX3D_DISPL].p;
        nw.veloc = d->s_field[AX3D_VELOC].p;
        nw.accel = d->s_field[AX3D_ACCEL].p;
        nw.stiff = d->s_field[AX3D_STIFF].p;
        nw.half_dt = (float)half_dt;
        nw.dt = (float)dt;
        nw.half_dt_dt = (float)half_dt_dt;
    }
    if (!fluid && d->nw_dbg.p && nw.on) nw.dbg = d->nw_dbg.p;
    const double kbytes = f.alg_b + (nw.on ? d->dom_bytes[1] : 0.0);
    HaloArgs halo{nullptr, nullptr, 0, nullptr};
    if (put && !fluid) halo = HaloArgs{d->halo_tab.p, d->halo_bcnt.p, f.nb_items, nullptr};
    if (d->cost_buf.p) halo.cost = d->cost_buf.p + d->cost_off[c];
    KTimer kt(d, fluid ? "k_elem3d_fused<fluid>" : (nw.on ? "k_elem3d_fused<solid> + in-kernel Newmark" : "k_elem3d_fused<solid>"), kbytes);
    fused_kernel(f)<<<f.grid, fused_nt(f), f.smem, d->stream>>>(
        d->desc[c].p + f.first, d->fused_items[c].p, f.nitems, d->plans.p, d->stwpool.p, d->geom.p, d->coef.p, d->attpar.p, d->attstate3d.p,
        fluid ? d->f_field[AX3D_DISPL].p : d->s_field[AX3D_DISPL].p, fluid ? d->f_field[AX3D_STIFF].p : d->s_field[AX3D_STIFF].p,
        f.tile_cap, d->fused_work.p + 2 * which, nw, halo);
}

static void set_fused_smem(int device, const FusedLaunch &f) {
    (void)device;
    if (f.smem > AX_FUSED_DYN_MAX) fail("ax3d::fused || shared-memory plan exceeds 227 KB");
    // several launches (and domains) may share one kernel instance: opt in to the maximum once
    CK(cudaFuncSetAttribute((const void *)fused_kernel(f), cudaFuncAttributeMaxDynamicSharedMemorySize, AX_FUSED_DYN_MAX));
}

// which: 1 = solid elements, 2 = fluid elements, 3 = both
// put: the solid fused launch sends the halo itself (in-kernel put; run_steps only)
static void compute_stiff(ax3d_domain *d, bool nw_on = false, double dt = 0.0, int which = 3, bool put = false) {
    TimerScope ts(d, 1);
    const int TB = AX_TILE * AX_NPE;
    if ((which & 1) && d->n_work[CLS_S1D]) {
        KTimer kt(d, "k_elem1d<solid>", d->cls_bytes[CLS_S1D]);
        (d->cls_prt[CLS_S1D] ? k_elem1d<false, true> : k_elem1d<false, false>)<<<d->n_work[CLS_S1D], TB, 0, d->stream>>>(d->desc[CLS_S1D].p, d->w_elem[CLS_S1D].p, d->w_a0[CLS_S1D].p, d->geom.p,
                                                                  d->coef.p, d->attpar.p, d->attstate1d.p, d->s_field[AX3D_DISPL].p,
                                                                  d->s_field[AX3D_STIFF].p);
        d->launches++;
    }
    if ((which & 2) && d->n_work[CLS_F1D]) {
        KTimer kt(d, "k_elem1d<fluid>", d->cls_bytes[CLS_F1D]);
        (d->cls_prt[CLS_F1D] ? k_elem1d<true, true> : k_elem1d<true, false>)<<<d->n_work[CLS_F1D], TB, 0, d->stream>>>(d->desc[CLS_F1D].p, d->w_elem[CLS_F1D].p, d->w_a0[CLS_F1D].p, d->geom.p,
                                                                 d->coef.p, d->attpar.p, d->attstate1d.p, d->f_field[AX3D_DISPL].p,
                                                                 d->f_field[AX3D_STIFF].p);
        d->launches++;
    }
    for (const Chunk &ch : d->chunks) {
        const int c = ch.cls;
        if (!(which & (c == CLS_S3D ? 1 : 2))) continue;
        KTimer kt(d, c == CLS_S3D ? "split pipeline <solid>: k_grad3d + k_fft3d_v2 + k_quad3d" : "split pipeline <fluid>: k_grad3d + k_fft3d_v2 + k_quad3d", ch.alg_b);
        if (c == CLS_S3D) {
            (ch.prt ? k_grad3d<false, true> : k_grad3d<false, false>)<<<ch.w_count, TB, 0, d->stream>>>(d->desc[c].p, d->w_elem[c].p + ch.w_begin, d->w_a0[c].p + ch.w_begin, d->geom.p,
                                                              d->s_field[AX3D_DISPL].p, d->scratch.p);
            fft_kernel(ch)<<<ch.f_count, ch.fft_np == 1 ? 256 : ch.fft_nt, ch.fft_smem, d->stream>>>(
                d->desc[c].p, d->fft_items[c].p + ch.f_begin, d->plans.p, d->stwpool.p, d->coef.p, d->attpar.p, d->attstate3d.p, d->scratch.p);
            (ch.prt ? k_quad3d<false, true> : k_quad3d<false, false>)<<<ch.w_count, TB, 0, d->stream>>>(d->desc[c].p, d->w_elem[c].p + ch.w_begin, d->w_a0[c].p + ch.w_begin, d->geom.p,
                                                              d->scratch.p, d->s_field[AX3D_STIFF].p);
        } else {
            (ch.prt ? k_grad3d<true, true> : k_grad3d<true, false>)<<<ch.w_count, TB, 0, d->stream>>>(d->desc[c].p, d->w_elem[c].p + ch.w_begin, d->w_a0[c].p + ch.w_begin, d->geom.p,
                                                             d->f_field[AX3D_DISPL].p, d->scratch.p);
            fft_kernel(ch)<<<ch.f_count, ch.fft_np == 1 ? 256 : ch.fft_nt, ch.fft_smem, d->stream>>>(
                d->desc[c].p, d->fft_items[c].p + ch.f_begin, d->plans.p, d->stwpool.p, d->coef.p, d->attpar.p, d->attstate3d.p, d->scratch.p);
            (ch.prt ? k_quad3d<true, true> : k_quad3d<true, false>)<<<ch.w_count, TB, 0, d->stream>>>(d->desc[c].p, d->w_elem[c].p + ch.w_begin, d->w_a0[c].p + ch.w_begin, d->geom.p,
                                                             d->scratch.p, d->f_field[AX3D_STIFF].p);
        }
        d->launches += 3;
    }
    for (int pass = 0; pass < 2; ++pass)   // the solid launch last: it may send the halo once every boundary force is in memory
        for (size_t k = 0; k < d->fused.size(); ++k) {
            const bool solid = d->fused[k].cls == CLS_S3D;
            if (solid != (pass == 1) || !(which & (solid ? 1 : 2))) continue;
            launch_fused(d, d->fused[k], (int)k, nw_on, dt, put);
            d->launches++;
        }
    CK(cudaGetLastError());
}

// rest_only: the solid-fluid points on the halo have been coupled by the in-kernel put already
static void couple_solid_fluid(ax3d_domain *d, bool rest_only = false) {
    TimerScope ts(d, 2);
    const SFTab &tab = rest_only ? d->sf_tab_rest : d->sf_tab;
    if (tab.nrows) {
        k_sf_couple<<<nblk(tab.nrows, 128), 128, 0, d->stream>>>(tab, d->s_field[AX3D_DISPL].p, d->s_field[AX3D_STIFF].p,
                                                                d->f_field[AX3D_STIFF].p);
        d->launches++;
    }
    if (!d->h_sf3d.empty()) {
        k_sf_couple3d<<<(int)d->h_sf3d.size(), 128, d->sf3d_smem, d->stream>>>(d->sf3d.p, d->plans.p, d->twpool.p, d->sf3d_pool.p,
                                                                              d->s_field[AX3D_DISPL].p, d->s_field[AX3D_STIFF].p,
                                                                              d->f_field[AX3D_STIFF].p);
        d->launches++;
    }
    CK(cudaGetLastError());
}

static void assemble_stiff(ax3d_domain *d, int phase) {
    if (d->nproc <= 1 || d->neigh_rank.empty()) return;
    TimerScope ts(d, 3);
    if (d->peer_halo) {
        const size_t total = d->neigh_begin.back();
        if (phase <= 0) {
            for (size_t n = 0; n < d->neigh_rank.size(); ++n) {
                const size_t b = d->neigh_begin[n], cnt = d->neigh_begin[n + 1] - b;
                if (!cnt) continue;
                k_halo_put<<<nblk(cnt, 256), 256, 0, d->stream>>>((int)cnt, d->halo_idx.p + b, d->s_field[AX3D_STIFF].p, d->f_field[AX3D_STIFF].p,
                                                                  d->peer_win[n], d->peer_stride[n], d->peer_count[n], d->halo_step.p);
                d->launches++;
            }
        }
        if (phase >= 0) {
            // neighbour order = reference order (Domain.cpp:143-149); one launch per neighbour keeps the sum order fixed
            for (size_t n = 0; n < d->neigh_rank.size(); ++n) {
                const size_t b = d->neigh_begin[n], cnt = d->neigh_begin[n + 1] - b;
                if (!cnt) continue;
                k_halo_wait_add<<<nblk(cnt, 256), 256, 0, d->stream>>>((int)cnt, d->halo_idx.p + b, d->halo_win.p + b, total,
                                                                       reinterpret_cast<unsigned *>(d->halo_win.p + 2 * total) + n,
                                                                       d->halo_step.p, d->s_field[AX3D_STIFF].p, d->f_field[AX3D_STIFF].p,
                                                                       d->bad_flag.p + 1);
                d->launches++;
            }
            k_halo_advance<<<1, 1, 0, d->stream>>>(d->halo_step.p);
            d->launches++;
        }
        CK(cudaGetLastError());
        return;
    }
#ifdef AX3D_WITH_NCCL
    if (!d->comm) fail("Domain::assembleStiff || no NCCL communicator (ax3d_set_messaging needs the unique id)");
    const size_t total = d->neigh_begin.back();
    if (phase <= 0) {
        k_pack<<<nblk(total, 256), 256, 0, d->stream>>>((int)total, d->halo_idx.p, d->s_field[AX3D_STIFF].p, d->f_field[AX3D_STIFF].p,
                                                        d->halo_send.p);
        d->launches++;
        ncclGroupStart();
        for (size_t n = 0; n < d->neigh_rank.size(); ++n) {
            const size_t b = d->neigh_begin[n], cnt = d->neigh_begin[n + 1] - b;
            ncclSend(d->halo_send.p + b, cnt * 2, ncclFloat, d->neigh_rank[n], d->comm, d->stream);
            ncclRecv(d->halo_recv.p + b, cnt * 2, ncclFloat, d->neigh_rank[n], d->comm, d->stream);
        }
        ncclGroupEnd();
    }
    if (phase >= 0) {
        // neighbour order = reference order (Domain.cpp:143-149); one launch per neighbour keeps the sum order fixed
        for (size_t n = 0; n < d->neigh_rank.size(); ++n) {
            const size_t b = d->neigh_begin[n], cnt = d->neigh_begin[n + 1] - b;
            if (!cnt) continue;
            k_unpack_add<<<nblk(cnt, 256), 256, 0, d->stream>>>((int)cnt, d->halo_idx.p + b, d->halo_recv.p + b, d->s_field[AX3D_STIFF].p,
                                                                d->f_field[AX3D_STIFF].p);
            d->launches++;
        }
    }
    CK(cudaGetLastError());
#else
    (void)phase;
    fail("Domain::assembleStiff || library built without NCCL");
#endif
}

// ------------------------------------------------------------------------------------------ extern "C"
#define API_BEGIN try {
#define API_END                          \
    return 0;                            \
    }                                    \
    catch (const std::exception &e) {    \
        g_last_error = e.what();         \
        return 1;                        \
    }                                    \
    catch (...) {                        \
        g_last_error = "ax3d || unknown exception"; \
        return 1;                        \
    }

extern "C" {

const char *ax3d_last_error(void) { return g_last_error.c_str(); }
int ax3d_version(void) { return 100; }

int ax3d_create(int device, ax3d_domain **out) {
    API_BEGIN
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) fail("Domain::Domain || no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= n) fail("Domain::Domain || invalid device ordinal");
    CK(cudaSetDevice(device));
    ax3d_domain *d = new ax3d_domain();
    d->device = device;
    *out = d;
    API_END
}

int ax3d_destroy(ax3d_domain *d) {
    API_BEGIN
    if (!d) return 0;
    cudaSetDevice(d->device);
    cudaDeviceSynchronize();
#ifdef AX3D_WITH_NCCL
    if (d->comm) ncclCommDestroy(d->comm);
#endif
    for (void *m : d->peer_mapped) cudaIpcCloseMemHandle(m);
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->ev0) cudaEventDestroy(d->ev0);
    if (d->ev1) cudaEventDestroy(d->ev1);
    if (d->ev2) cudaEventDestroy(d->ev2);
    if (d->ev3) cudaEventDestroy(d->ev3);
    if (d->rec_host) cudaFreeHost(d->rec_host);
    if (d->stf_pinned) cudaFreeHost(d->stf_pinned);
    if (d->rec_ring_host) cudaFreeHost(d->rec_ring_host);
    for (int v = 0; v < 8; ++v) {
        if (d->graph_exec[v]) cudaGraphExecDestroy(d->graph_exec[v]);
        if (d->graph[v]) cudaGraphDestroy(d->graph[v]);
    }
    delete d;
    API_END
}

int ax3d_set_gmat(ax3d_domain *d, const double G_GLL[25], const double G_GLJ[25]) {
    API_BEGIN
    check_open(d);
    memcpy(d->G_GLL, G_GLL, sizeof(d->G_GLL));
    memcpy(d->G_GLJ, G_GLJ, sizeof(d->G_GLJ));
    d->have_g = true;
    API_END
}

int ax3d_add_solid_point(ax3d_domain *d, int nr, int axial, const double crds[2], int n_invmass, const float *invmass, int *tag) {
    API_BEGIN
    *tag = add_point(d, 0, nr, axial, n_invmass, invmass, 0, nullptr, 0, 0, nullptr, nullptr);
    set_crds(d, *tag, crds);
    API_END
}
/* SolidPoint with an ocean load on top (GLLPoint.cpp:57-72): rows = 1 -> MassOcean1D(mass, massOcean, theta)
 * (MassOcean1D.cpp:8-13), normal_or_theta[0] = theta; rows = nr -> MassOcean3D(mass[nr], massOcean[nr], unit normal [nr x 3,
 * column-major]) (MassOcean3D.cpp:9-16).  The conversions to Real are done here, in double, as the reference does. */
int ax3d_add_solid_point_ocean(ax3d_domain *d, int nr, int axial, const double crds[2], int rows, const double *mass, const double *mass_ocean,
                               const double *normal_or_theta, int *tag) {
    API_BEGIN
    if (rows != 1 && rows != nr) fail("MassOcean3D::checkCompatibility || Incompatible size.");
    std::vector<float> im(rows);
    for (int i = 0; i < rows; ++i) im[i] = (float)(1.0 / mass[i]);
    *tag = add_point(d, 0, nr, axial, rows, im.data(), 0, nullptr, 0, 0, nullptr, nullptr);
    set_crds(d, *tag, crds);
    HPoint &p = d->points[*tag];
    if (rows == 1) {
        p.ocean = 1;
        p.oc_imZ = (float)(1.0 / (mass[0] + mass_ocean[0]));
        p.oc_sint = (float)sin(normal_or_theta[0]);
        p.oc_cost = (float)cos(normal_or_theta[0]);
    } else {
        p.ocean = 3;
        p.oc_ns.resize((size_t)3 * nr);
        for (int i = 0; i < nr; ++i) {
            const float scal = (float)sqrt(mass_ocean[i] / (mass[i] * (mass[i] + mass_ocean[i])));
            for (int c = 0; c < 3; ++c) p.oc_ns[(size_t)c * nr + i] = (float)normal_or_theta[(size_t)c * nr + i] * scal;
        }
    }
    API_END
}
int ax3d_add_fluid_point(ax3d_domain *d, int nr, int axial, const double crds[2], int n_invmass, const float *invmass, int fluid_surf,
                         int *tag) {
    API_BEGIN
    *tag = add_point(d, 1, nr, axial, 0, nullptr, n_invmass, invmass, fluid_surf, 0, nullptr, nullptr);
    set_crds(d, *tag, crds);
    API_END
}
int ax3d_add_solid_fluid_point(ax3d_domain *d, int nr, int axial, const double crds[2], int n_s, const float *im_s, int n_f,
                               const float *im_f, int fluid_surf, int n_sf, const float *n_un, const float *n_as, int *tag) {
    API_BEGIN
    *tag = add_point(d, 2, nr, axial, n_s, im_s, n_f, im_f, fluid_surf, n_sf, n_un, n_as);
    set_crds(d, *tag, crds);
    API_END
}

int ax3d_add_solid_element(ax3d_domain *d, const int tags[25], const double *geom, int axial, const double *theta, int law, int rows,
                           const float *coef, const ax3d_attenuation *att, int *tag) {
    API_BEGIN
    HElem e;
    common_elem(d, e, tags, geom, axial, false);
    if (law < 0 || law > 2) fail("SolidElement::SolidElement || unknown elastic law");
    if (rows != 1 && rows != e.nr) fail("Elastic3D::checkCompatibility || Incompatible size.");
    e.law = law;
    e.rows = rows;
    e.ncoef = law == AX3D_ISO ? 2 : law == AX3D_TI ? 5 : 21;
    e.coef.assign(coef, coef + (size_t)e.ncoef * rows * AX_NPE);
    if (theta) {
        memcpy(e.theta, theta, sizeof(e.theta));
    } else {   // Element::formThetaMat (Element.cpp:48-58) with Geodesy::theta of the points' (s, z)
        for (int i = 0; i < AX_NPE; ++i) {
            const double *c = d->points[e.pt[i]].crds;
            const double r = sqrt(c[0] * c[0] + c[1] * c[1]);
            e.theta[i] = r < 1e-10 ? 0.0 : acos(std::max(-1.0, std::min(1.0, c[1] / r)));
        }
    }
    if (att && att->kind != AX3D_ATT_NONE) {
        if (att->kind != AX3D_ATT_FULL && att->kind != AX3D_ATT_CG4) fail("Attenuation || unknown kind");
        const int P = att->kind == AX3D_ATT_CG4 ? 4 : AX_NPE;
        e.att.kind = att->kind;
        e.att.nsls = att->nsls;
        e.att.do_kappa = att->do_kappa;
        e.att.alpha.assign(att->alpha, att->alpha + att->nsls);
        e.att.beta.assign(att->beta, att->beta + att->nsls);
        e.att.gamma.assign(att->gamma, att->gamma + att->nsls);
        e.att.dkappa.assign(att->dkappa, att->dkappa + (size_t)rows * P);
        e.att.dmu.assign(att->dmu, att->dmu + (size_t)rows * P);
    }
    d->elems.push_back(std::move(e));
    *tag = (int)d->elems.size() - 1;
    API_END
}

int ax3d_add_fluid_element(ax3d_domain *d, const int tags[25], const double *geom, int axial, int rows, const float *K, int *tag) {
    API_BEGIN
    HElem e;
    common_elem(d, e, tags, geom, axial, true);
    if (rows != 1 && rows != e.nr) fail("Acoustic3D::checkCompatibility || Incompatible size.");
    e.law = AX3D_ISO;
    e.rows = rows;
    e.ncoef = 1;
    e.coef.assign(K, K + (size_t)rows * AX_NPE);
    for (int i = 0; i < AX_NPE; ++i) {   // Element::formThetaMat (Element.cpp:48-58): the strain read-back rotates to RTZ (forceTIso)
        const double *c = d->points[e.pt[i]].crds;
        const double r = sqrt(c[0] * c[0] + c[1] * c[1]);
        e.theta[i] = r < 1e-10 ? 0.0 : acos(std::max(-1.0, std::min(1.0, c[1] / r)));
    }
    d->elems.push_back(std::move(e));
    *tag = (int)d->elems.size() - 1;
    API_END
}

/* the PRT* argument of SolidElement / FluidElement (Quad.cpp:386-420, 527-547): PRT_1D(array<RMatPP, 4>) with rows = 1,
 * PRT_3D(RMatXN4) with rows = Nr; X = [4][25][rows].  theta = Element::formThetaMat() (Element.cpp:48-58), needed by
 * the fluid elements, which rotate only when they carry a PRT (FluidElement.cpp:21). */
int ax3d_set_element_prt(ax3d_domain *d, int elem_tag, int rows, const float *X, const double theta[25]) {
    API_BEGIN
    check_open(d);
    if (elem_tag < 0 || elem_tag >= (int)d->elems.size()) fail("Element::Element || invalid element tag");
    HElem &e = d->elems[elem_tag];
    if (rows != 1 && rows != e.nr) fail("PRT_3D::checkCompatibility || Incompatible size.");
    if ((rows == 1) != (e.rows == 1))
        fail(std::string(e.fluid ? "FluidElement::FluidElement" : "SolidElement::SolidElement") +
             " || Particle Relabelling and Elasticity are generated in different spaces.");
    e.prt_rows = rows;
    e.prtX.assign(X, X + (size_t)4 * AX_NPE * rows);
    memcpy(e.theta, theta, sizeof(e.theta));
    API_END
}

int ax3d_add_source_term(ax3d_domain *d, int elem_tag, const int nrow[25], const float *force) {
    API_BEGIN
    check_open(d);
    if (elem_tag < 0 || elem_tag >= (int)d->elems.size()) fail("SourceTerm::SourceTerm || invalid element tag");
    HSource s;
    s.elem = elem_tag;
    size_t tot = 0;
    for (int i = 0; i < AX_NPE; ++i) {
        s.nrow[i] = nrow[i];
        tot += (size_t)3 * nrow[i];
    }
    s.force.assign(force, force + 2 * tot);
    d->sources.push_back(std::move(s));
    API_END
}

int ax3d_set_messaging(ax3d_domain *d, int rank, int nproc, const void *uid, int nneigh, const int *neigh_rank, const int *npoints,
                       const int *point_tags) {
    API_BEGIN
    check_open(d);
    d->rank = rank;
    d->nproc = nproc;
    d->neigh_rank.assign(neigh_rank, neigh_rank + nneigh);
    d->neigh_pts.clear();
    size_t pos = 0;
    for (int n = 0; n < nneigh; ++n) {
        d->neigh_pts.emplace_back(point_tags + pos, point_tags + pos + npoints[n]);
        pos += npoints[n];
    }
    if (uid) {
        memcpy(d->uid, uid, 128);
        d->have_uid = true;
    }
    API_END
}

/* Peer-memory halo, step 1 (after ax3d_finalize_setup): describe this rank's receive window.  out_handle64 receives the
 * cudaIpcMemHandle_t of the window, out_ptr its address in this process (for neighbours that live in the same process),
 * neigh_begin[nneigh + 1] the start of every neighbour's segment (float2 units), counters follow the window:
 * the counter for neighbour n is at ((unsigned *)(win + 2 * total)) -- i.e. byte offset 16 * total + 4 * n -- see below. */
int ax3d_halo_export(ax3d_domain *d, void *out_handle64, void **out_ptr, long long *neigh_begin, long long *total) {
    API_BEGIN
    check_final(d);
    if (d->nproc <= 1 || d->neigh_rank.empty()) fail("Domain::setMessaging || this rank has no neighbours");
    // one allocation = window + counters, so that one handle maps both
    const size_t tot = d->neigh_begin.back(), nn = d->neigh_rank.size();
    if (out_handle64) {
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, d->halo_win.p));
        static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t");
        memcpy(out_handle64, &h, 64);
    }
    if (out_ptr) *out_ptr = d->halo_win.p;
    for (size_t n = 0; n <= nn; ++n) neigh_begin[n] = (long long)d->neigh_begin[n];
    *total = (long long)tot;
    API_END
}

/* Peer-memory halo, step 2: for every neighbour (in ax3d_set_messaging order) the neighbour's window -- as a 64-byte IPC
 * handle (handles != NULL: other process on the same node) or as a device pointer of this process (ptrs) --, the start
 * of MY segment in it (peer_begin, float2 units), its total (= parity stride) and my index in its neighbour list
 * (peer_slot: which of its counters is mine).  From here on Domain::assembleStiff uses k_halo_put / k_halo_wait_add
 * and multi-rank steps replay from the CUDA graph. */
int ax3d_halo_connect(ax3d_domain *d, int nneigh, const void *handles, void *const *ptrs, const long long *peer_begin,
                      const long long *peer_total, const int *peer_slot) {
    API_BEGIN
    check_final(d);
    if (nneigh != (int)d->neigh_rank.size()) fail("Domain::setMessaging || ax3d_halo_connect: neighbour count mismatch");
    d->peer_win.assign(nneigh, nullptr);
    d->peer_stride.assign(nneigh, 0);
    d->peer_count.assign(nneigh, nullptr);
    for (int n = 0; n < nneigh; ++n) {
        void *base = nullptr;
        if (handles) {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char *)handles + 64 * n, 64);
            CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            d->peer_mapped.push_back(base);
        } else {
            base = ptrs[n];
        }
        if (!base) fail("Domain::setMessaging || ax3d_halo_connect: null window");
        d->peer_win[n] = (float2 *)base + peer_begin[n];
        d->peer_stride[n] = (size_t)peer_total[n];
        d->peer_count[n] = (unsigned *)((float2 *)base + 2 * peer_total[n]) + peer_slot[n];
    }
    d->peer_halo = true;
    // in-kernel put (fused.cuh: halo_put_cta): needs a solid fused launch to carry it, at most AX_MAX_NEIGH neighbours and no
    // 3D-coupled solid-fluid point on the halo (its coupling is an FFT per point: k_sf_couple3d); otherwise k_halo_put kernels
    {
        bool solid_fused = false;
        for (const FusedLaunch &f : d->fused) solid_fused = solid_fused || f.cls == CLS_S3D;
        // Opt-in (AX3D_INKERNEL_PUT=1).  Measured on 2 and 4 B200s (cfg4 weak scaling, profiles/r2_scaling.md): the exchange
        // itself disappears from the step (halo family 0.135 -> 0.036 ms) but the step does not get shorter (0.951 vs 0.941 ms
        // at N = 4): what a rank waits for at the end of its step is its slower neighbour's compute, not the transfer, and the
        // fences and boundary-first order cost the element kernel ~1 %.  Default: k_halo_put kernels behind the coupling.
        const char *env = getenv("AX3D_INKERNEL_PUT");
        d->inkernel_put = solid_fused && nneigh <= AX_MAX_NEIGH && !d->halo_sf3d && (env && atoi(env) != 0);
        if (d->inkernel_put) {
            HaloTab ht;
            memset(&ht, 0, sizeof(ht));
            ht.nneigh = nneigh;
            for (int n = 0; n < nneigh; ++n) {
                const size_t b = d->neigh_begin[n], cnt = d->neigh_begin[n + 1] - b;
                ht.peer[n] = HaloPeer{d->halo_idx.p + b, d->peer_win[n], (unsigned long long)d->peer_stride[n], d->peer_count[n], (int)cnt,
                                      nblk(cnt, 256)};
            }
            ht.step = d->halo_step.p;
            ht.f_stiff = d->f_field[AX3D_STIFF].p;
            ht.sf_halo = d->sf_tab_halo;
            std::vector<HaloTab> v(1, ht);
            d->halo_tab.upload(v);
            d->halo_bcnt.alloc(1);
            d->halo_bcnt.zero();
            d->dual_chain = false;   // every other element kernel of the step must be complete when the solid launch sends
        }
    }
    for (int v = 0; v < 8; ++v) {   // step graphs captured before the connection do not contain the exchange
        if (d->graph_exec[v]) { cudaGraphExecDestroy(d->graph_exec[v]); d->graph_exec[v] = nullptr; }
        if (d->graph[v]) { cudaGraphDestroy(d->graph[v]); d->graph[v] = nullptr; }
    }
    API_END
}

/* ncclGetUniqueId on the calling rank (rank 0 calls it and broadcasts the 128 bytes; XMPI::initialize analogue). */
int ax3d_nccl_unique_id(void *out128) {
    API_BEGIN
#ifdef AX3D_WITH_NCCL
    ncclUniqueId id;
    ncclResult_t r = ncclGetUniqueId(&id);
    if (r != ncclSuccess) fail(std::string("XMPI::initialize || ncclGetUniqueId: ") + ncclGetErrorString(r));
    memset(out128, 0, 128);
    memcpy(out128, &id, sizeof(id) < 128 ? sizeof(id) : 128);
#else
    (void)out128;
    fail("XMPI::initialize || library built without NCCL");
#endif
    API_END
}

int ax3d_finalize_setup(ax3d_domain *d) {
    API_BEGIN
    finalize(d);
    API_END
}

int ax3d_update_newmark(ax3d_domain *d, double dt) {
    API_BEGIN
    check_final(d);
    update_newmark(d, dt);
    API_END
}
int ax3d_apply_source(ax3d_domain *d, float stf) {
    API_BEGIN
    check_final(d);
    apply_source(d, stf);
    API_END
}
int ax3d_compute_stiff(ax3d_domain *d) {
    API_BEGIN
    check_final(d);
    compute_stiff(d);
    API_END
}
int ax3d_couple_solid_fluid(ax3d_domain *d) {
    API_BEGIN
    check_final(d);
    couple_solid_fluid(d);
    API_END
}
int ax3d_assemble_stiff(ax3d_domain *d, int phase) {
    API_BEGIN
    check_final(d);
    assemble_stiff(d, phase);
    API_END
}

static void learn_wisdom(ax3d_domain *d) {
    if (d->ns) {
        k_learn_wisdom<3><<<nblk(d->ns * 3 * 32, 256), 256, 0, d->stream>>>(d->s_tab, (int)d->ns, d->s_field[AX3D_DISPL].p, d->learn_cutoff,
                                                                           d->s_wis_max.p, d->s_wis_nu.p);
        d->launches++;
    }
    if (d->nf) {
        k_learn_wisdom<1><<<nblk(d->nf * 32, 256), 256, 0, d->stream>>>(d->f_tab, (int)d->nf, d->f_field[AX3D_DISPL].p, d->learn_cutoff,
                                                                       d->f_wis_max.p, d->f_wis_nu.p);
        d->launches++;
    }
    CK(cudaGetLastError());
}

/* Domain::setLearnParameters(LearnParameters{invoked, cutoff, interval}) (Domain.h:39; axisem.cpp:158-166) */
int ax3d_set_learn_parameters(ax3d_domain *d, int invoked, float cutoff, int interval) {
    API_BEGIN
    check_final(d);
    if (interval <= 0) fail("Domain::setLearnParameters || interval must be positive");
    d->learn_invoked = invoked != 0;
    d->learn_cutoff = cutoff;
    d->learn_interval = interval;
    if (d->learn_invoked && !d->s_wis_max.p && !d->f_wis_max.p) {
        std::vector<float> m3(d->ns * 3, -1.f), m1(d->nf, -1.f);
        std::vector<int> n3(d->ns * 3), n1(d->nf);
        for (const HPoint &p : d->points) {
            if (p.s_idx >= 0) for (int c = 0; c < 3; ++c) n3[(size_t)p.s_idx * 3 + c] = p.nu;
            if (p.f_idx >= 0) n1[p.f_idx] = p.nu;
        }
        d->s_wis_max.upload(m3); d->s_wis_nu.upload(n3);
        d->f_wis_max.upload(m1); d->f_wis_nu.upload(n1);
    }
    API_END
}

/* Domain::learnWisdom(tstep) (Domain.cpp:384-402): Point::learnWisdom(cutoff) on every point when tstep % interval == 0 */
int ax3d_learn_wisdom(ax3d_domain *d, int tstep) {
    API_BEGIN
    check_final(d);
    if (!d->learn_invoked) return 0;
    if (tstep % d->learn_interval == 0) learn_wisdom(d);
    API_END
}

/* what Domain::dumpWisdom (Domain.cpp:404-440) collects: Point::getNuWisdom() per point tag (solid: max over the three
 * components, SolidPoint.cpp:268-272; solid-fluid: max of both parts, SolidFluidPoint.cpp:129-131) */
int ax3d_get_nu_wisdom(ax3d_domain *d, int *nu_wisdom, int npoints) {
    API_BEGIN
    check_final(d);
    if (npoints != (int)d->points.size()) fail("Domain::dumpWisdom || size mismatch");
    if (!d->learn_invoked) fail("Domain::dumpWisdom || wisdom learning was not invoked (ax3d_set_learn_parameters)");
    std::vector<int> n3(d->ns * 3), n1(d->nf);
    CK(cudaStreamSynchronize(d->stream));
    if (d->ns) CK(cudaMemcpy(n3.data(), d->s_wis_nu.p, n3.size() * sizeof(int), cudaMemcpyDeviceToHost));
    if (d->nf) CK(cudaMemcpy(n1.data(), d->f_wis_nu.p, n1.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (size_t t = 0; t < d->points.size(); ++t) {
        const HPoint &p = d->points[t];
        int v = 0;
        if (p.s_idx >= 0) for (int c = 0; c < 3; ++c) v = std::max(v, n3[(size_t)p.s_idx * 3 + c]);
        if (p.f_idx >= 0) v = std::max(v, n1[p.f_idx]);
        nu_wisdom[t] = v;
    }
    API_END
}

int ax3d_check_stability(ax3d_domain *d, int *stable) {
    API_BEGIN
    check_final(d);
    CK(cudaMemsetAsync(d->bad_flag.p, 0, sizeof(int), d->stream));
    if (d->s_len) k_check_finite<<<296, 256, 0, d->stream>>>(d->s_len * 2, (const float *)d->s_field[0].p, d->bad_flag.p);
    if (d->f_len) k_check_finite<<<296, 256, 0, d->stream>>>(d->f_len * 2, (const float *)d->f_field[0].p, d->bad_flag.p);
    d->launches += 2;
    int bad2[2] = {0, 0};
    CK(cudaMemcpyAsync(bad2, d->bad_flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream));
    if (bad2[1]) fail("Domain::assembleStiff || peer-memory halo: a neighbour's boundary stiffness never arrived (timeout)");
    const int bad = bad2[0];
    if (d->n_plain > 0) {
        unsigned ctl[4] = {0, 0, 0, 0};
        CK(cudaMemcpy(ctl, d->nw_ctl.p, sizeof(ctl), cudaMemcpyDeviceToHost));
        if (ctl[3]) fail("Domain::updateNewmark || in-kernel Newmark queue timed out (internal bookkeeping error)");
    }
    if (d->nw_dbg.p) {
        std::vector<unsigned long long> h(4 * 256);
        CK(cudaMemcpy(h.data(), d->nw_dbg.p, h.size() * 8, cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull, e_max = 0, e_min = ~0ull, n_max = 0, c_max = 0;
        for (int b = 0; b < d->num_sm && b < 256; ++b) {
            if (!h[4 * b]) continue;
            t0 = std::min(t0, h[4 * b]);
            e_max = std::max(e_max, h[4 * b + 1]); e_min = std::min(e_min, h[4 * b + 1]);
            n_max = std::max(n_max, h[4 * b + 2]); c_max = std::max(c_max, h[4 * b + 3]);
        }
        fprintf(stderr, "[nw-debug] elements done: first CTA %.1f us, last CTA %.1f us; newmark warps done %.1f us; compute warps done %.1f us (n_plain %d of %zu)\n",
                (e_min - t0) * 1e-3, (e_max - t0) * 1e-3, (n_max - t0) * 1e-3, (c_max - t0) * 1e-3, d->n_plain, d->ns);
    }
    *stable = bad ? 0 : 1;
    API_END
}

int ax3d_reset_zero(ax3d_domain *d) {
    API_BEGIN
    check_final(d);
    CK(cudaStreamSynchronize(d->stream));
    for (int k = 0; k < 4; ++k) {
        d->s_field[k].zero();
        d->f_field[k].zero();
    }
    d->attstate1d.zero();
    d->attstate3d.zero();
    d->tstep = 0;              // Newmark::solve restarts its step counter after Domain::resetZero (Newmark.cpp:30-47)
    d->plain_advanced = false;
    API_END
}

// One iteration of Newmark::solve (Newmark.cpp:47-93).  special_only: the plain solid points already hold this step's
// state (advanced under the previous step's element kernel); nw_on: this step's element kernel advances them to the next.
static void step_body(ax3d_domain *d, double dt, bool special_only = false, bool nw_on = false, bool record = false) {
    if (d->dual_chain && !d->timers && d->chunks.empty()) {
        // Solid and fluid points / elements only meet in coupleSolidFluid: the fluid chain (point update, fluid elements)
        // runs on a second, lower-priority stream next to the solid chain and fills the SMs the persistent solid element
        // kernel leaves idle at its start and in its tail (fork / join by events: capturable into the step graph).
        cudaStream_t s = d->stream;
        struct StreamSwap {   // the launch helpers enqueue on d->stream: restored on every exit path, exceptions included
            ax3d_domain *d;
            cudaStream_t keep;
            StreamSwap(ax3d_domain *d_, cudaStream_t to) : d(d_), keep(d_->stream) { d->stream = to; }
            ~StreamSwap() { d->stream = keep; }
        };
        auto fluid_chain = [&]() {
            CK(cudaEventRecord(d->ev_fork, s));
            CK(cudaStreamWaitEvent(d->stream2, d->ev_fork, 0));
            StreamSwap on_stream2(d, d->stream2);
            update_newmark(d, dt, special_only, 2);
            if (record) launch_record(d, d->rec_ring.p, reinterpret_cast<const int *>(d->stf_dev.p + 1), 3 * d->nrec_total(), 2);
            compute_stiff(d, nw_on, dt, 2);
            CK(cudaEventRecord(d->ev_join, d->stream2));
        };
        if (d->dual_mode == 1) fluid_chain();          // fork at the top of the step
        update_newmark(d, dt, special_only, 1);
        if (record) launch_record(d, d->rec_ring.p, reinterpret_cast<const int *>(d->stf_dev.p + 1), 3 * d->nrec_total(), 1);
        launch_source(d);
        if (d->dual_mode == 2) fluid_chain();          // fork just before the solid elements: the fluid chain fills their tail
        compute_stiff(d, nw_on, dt, 1);
        if (d->dual_mode == 3) fluid_chain();          // enqueued behind the solid element launch
        CK(cudaStreamWaitEvent(s, d->ev_join, 0));
        couple_solid_fluid(d);
        assemble_stiff(d, -1);
        assemble_stiff(d, 1);
        return;
    }
    update_newmark(d, dt, special_only);
    // the reference records after the update of the step (Newmark.cpp:64-70); it must also precede this step's element
    // kernel, which already advances the plain points to the next step
    if (record) launch_record(d, d->rec_ring.p, reinterpret_cast<const int *>(d->stf_dev.p + 1), 3 * d->nrec_total());
    launch_source(d);
    const bool put = d->inkernel_put && d->nproc > 1 && !d->neigh_rank.empty();
    compute_stiff(d, nw_on, dt, 3, put);
    couple_solid_fluid(d, put);
    if (!put) assemble_stiff(d, -1);   // no-ops on a single rank
    assemble_stiff(d, 1);
}

static long long count_step_launches(ax3d_domain *d, bool special_only, bool record) {
    long long n = 0;
    if (record) for (int c = 0; c < NCLS; ++c) n += d->nrec_c[c] > 0;
    n += !d->h_m3d_s.empty();
    n += !d->h_oc1d.empty();
    n += !d->h_m3d_f.empty();
    n += (special_only ? d->s_tab_sp.nrows : d->s_tab.nrows) > 0;
    n += d->f_tab.nrows > 0;
    n += d->n_src > 0;
    n += d->n_work[CLS_S1D] > 0;
    n += d->n_work[CLS_F1D] > 0;
    n += 3 * (long long)d->chunks.size();
    n += (long long)d->fused.size();
    n += d->sf_tab.nrows > 0;
    n += !d->h_sf3d.empty();
    if (d->nproc > 1 && !d->neigh_rank.empty()) {
        n += 1;   // k_pack (NCCL path) / k_halo_advance (peer path)
        for (size_t k = 0; k < d->neigh_rank.size(); ++k)
            n += (d->peer_halo && !d->inkernel_put ? 2 : 1) * (d->neigh_begin[k + 1] > d->neigh_begin[k]);   // [k_halo_put +] k_unpack_add / k_halo_wait_add
    }
    return n;
}

static void run_steps(ax3d_domain *d, int nsteps, double dt, const float *stf, bool record = false) {
    // Multi-rank steps run eagerly: capturing the NCCL send/recv group into the step graph hung on 2 x B200 (NCCL 2.28.9,
    // thread-local capture), so it stays an opt-in experiment (AX3D_HALO_GRAPH=1).
    static const bool halo_graph = getenv("AX3D_HALO_GRAPH") && atoi(getenv("AX3D_HALO_GRAPH")) != 0;
    const bool graph_ok = d->use_graph && !d->timers && (d->nproc <= 1 || d->neigh_rank.empty() || halo_graph || d->peer_halo);
    if (graph_ok && d->graph_dt != dt) {
        for (int v = 0; v < 8; ++v) {
            if (d->graph_exec[v]) { cudaGraphExecDestroy(d->graph_exec[v]); d->graph_exec[v] = nullptr; }
            if (d->graph[v]) { cudaGraphDestroy(d->graph[v]); d->graph[v] = nullptr; }
        }
        d->graph_dt = dt;
    }
    const bool can_nw = d->n_plain > 0;
    for (int i = 0; i < nsteps; ++i) {
        // every step but the last of this call advances the plain points to the next step under its element kernel,
        // so that the state after the call is exactly the reference's (all points at step i, stiff = this step's force)
        const bool special_only = d->plain_advanced;
        // Domain::learnWisdom(tstep - 1) (Newmark.cpp:89): needs every point at this step, so a learning step does not
        // advance the plain points under its element kernel
        const bool learn_now = d->learn_invoked && d->tstep % d->learn_interval == 0;
        const bool nw_on = can_nw && i + 1 < nsteps && !learn_now;
        if (d->n_src || record) push_stf(d, stf ? stf[i] : 0.f, i);
        if (graph_ok) {
            const int v = (record ? 4 : 0) + (special_only ? 2 : 0) + (nw_on ? 1 : 0);
            if (!d->graph_exec[v]) {
                const long long before = d->launches;
                CK(cudaStreamBeginCapture(d->stream, cudaStreamCaptureModeThreadLocal));
                try {
                    step_body(d, dt, special_only, nw_on, record);
                } catch (...) {   // never leave the stream capturing: end (and discard) the capture, then report
                    cudaGraph_t broken = nullptr;
                    cudaStreamEndCapture(d->stream, &broken);
                    if (broken) cudaGraphDestroy(broken);
                    cudaGetLastError();
                    d->launches = before;
                    throw;
                }
                CK(cudaStreamEndCapture(d->stream, &d->graph[v]));
                CK(cudaGraphInstantiate(&d->graph_exec[v], d->graph[v], 0));
                d->launches = before;   // capture enqueues nothing
            }
            CK(cudaGraphLaunch(d->graph_exec[v], d->stream));
            d->launches += count_step_launches(d, special_only, record);
        } else {
            step_body(d, dt, special_only, nw_on, record);
        }
        d->plain_advanced = nw_on;
        if (learn_now) learn_wisdom(d);
        d->tstep++;
    }
}

int ax3d_run_steps(ax3d_domain *d, int nsteps, double dt, const float *stf) {
    API_BEGIN
    check_final(d);
    run_steps(d, nsteps, dt, stf);
    API_END
}

/* Newmark::solve with the pointwise recorder buffering on the device (PointwiseRecorder's dump interval = nsteps):
 * out[(i * nrec + r) * 3 + c] = sample of receiver r at step i, copied to the host once at the end. */
int ax3d_run_steps_record(ax3d_domain *d, int nsteps, double dt, const float *stf, float *out) {
    API_BEGIN
    check_final(d);
    const int n = d->nrec_total();
    if (!n) fail("PointwiseRecorder::record || no receivers registered (ax3d_set_receivers)");
    if (nsteps <= 0) return 0;
    if (nsteps > STF_RING) fail("Newmark::solve || ax3d_run_steps_record takes at most 4096 steps per call");
    if ((size_t)nsteps > d->rec_ring_steps) {
        CK(cudaStreamSynchronize(d->stream));
        for (int v = 4; v < 8; ++v) {   // the ring address is baked into the recording graphs
            if (d->graph_exec[v]) { cudaGraphExecDestroy(d->graph_exec[v]); d->graph_exec[v] = nullptr; }
            if (d->graph[v]) { cudaGraphDestroy(d->graph[v]); d->graph[v] = nullptr; }
        }
        if (d->rec_ring_host) cudaFreeHost(d->rec_ring_host);
        d->rec_ring.alloc((size_t)nsteps * n * 3);
        CK(cudaMallocHost(&d->rec_ring_host, (size_t)nsteps * n * 3 * sizeof(float)));
        d->rec_ring_steps = nsteps;
    }
    run_steps(d, nsteps, dt, stf, true);
    CK(cudaMemcpyAsync(d->rec_ring_host, d->rec_ring.p, (size_t)nsteps * n * 3 * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream));
    for (int i = 0; i < nsteps; ++i) {
        const float *row = d->rec_ring_host + (size_t)i * n * 3;
        float *o = out + (size_t)i * n * 3;
        unscramble_record(d, row, o);
    }
    API_END
}

/* nsteps of the loop timed with CUDA events on the launching stream; *ms = device time. */
int ax3d_run_steps_timed(ax3d_domain *d, int nsteps, double dt, const float *stf, float *ms) {
    API_BEGIN
    check_final(d);
    CK(cudaStreamSynchronize(d->stream));
    CK(cudaEventRecord(d->ev0, d->stream));
    run_steps(d, nsteps, dt, stf);
    CK(cudaEventRecord(d->ev1, d->stream));
    CK(cudaEventSynchronize(d->ev1));
    CK(cudaEventElapsedTime(ms, d->ev0, d->ev1));
    API_END
}

/* PointwiseRecorder set-up (ReceiverCollection::release): registers nrec receivers once. */
int ax3d_set_receivers(ax3d_domain *d, int nrec, const int *elem_tags, const float *phi, const float *weights) {
    API_BEGIN
    check_final(d);
    // A second registration replaces the first: the recording graphs bake in the old item / weight pointers, grid sizes and
    // ring row stride, and the ring was sized for the old receiver count -- drain the stream, drop them and start over.
    CK(cudaStreamSynchronize(d->stream));
    for (int v = 4; v < 8; ++v) {
        if (d->graph_exec[v]) { cudaGraphExecDestroy(d->graph_exec[v]); d->graph_exec[v] = nullptr; }
        if (d->graph[v]) { cudaGraphDestroy(d->graph[v]); d->graph[v] = nullptr; }
    }
    d->rec_ring_steps = 0;
    std::vector<RecvItem> it[NCLS];
    std::vector<float> w[NCLS];
    for (int c = 0; c < NCLS; ++c) d->rec_where_c[c].clear();
    d->rec_fluid_smem = 0;
    for (int i = 0; i < nrec; ++i) {
        if (elem_tags[i] < 0 || elem_tags[i] >= (int)d->elems.size()) fail("PointwiseRecorder::record || invalid element tag");
        const HElem &E = d->elems[elem_tags[i]];
        RecvItem r{E.idx, phi[i]};
        it[E.cls].push_back(r);
        w[E.cls].insert(w[E.cls].end(), weights + (size_t)i * AX_NPE, weights + (size_t)(i + 1) * AX_NPE);
        d->rec_where_c[E.cls].push_back(i);
        if (E.cls == CLS_F3D) d->rec_fluid_smem = std::max(d->rec_fluid_smem, recf_smem(E.nr));
    }
    for (int c = 0; c < NCLS; ++c) {
        d->nrec_c[c] = (int)it[c].size();
        d->rec_items_c[c].upload(it[c]);
        d->rec_w_c[c].upload(w[c]);
    }
    if (d->rec_fluid_smem > (size_t)224 * 1024) fail("FluidElement::computeGroundMotion || Nr too large for the receiver kernel");
    if (d->rec_fluid_smem > (size_t)48 * 1024)
        CK(cudaFuncSetAttribute(k_ground_motion_fluid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d->rec_fluid_smem));
    if ((size_t)nrec > d->rec_cap) {
        if (d->rec_host) cudaFreeHost(d->rec_host);
        CK(cudaMallocHost(&d->rec_host, (size_t)nrec * 3 * sizeof(float)));
        d->rec_out.alloc((size_t)nrec * 3);
        d->rec_cap = nrec;
    }
    API_END
}

/* PointwiseRecorder::record (PointwiseRecorder.cpp:62-144) for the registered receivers: evaluates
 * Element::computeGroundMotion on the device and copies nrec x 3 floats to the host (pinned staging). */
int ax3d_record(ax3d_domain *d, float *out) {
    API_BEGIN
    check_final(d);
    const int n = d->nrec_total();
    if (!n) return 0;
    launch_record(d, d->rec_out.p, nullptr, 0);
    CK(cudaMemcpyAsync(d->rec_host, d->rec_out.p, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
    CK(cudaStreamSynchronize(d->stream));
    unscramble_record(d, d->rec_host, out);
    API_END
}

int ax3d_synchronize(ax3d_domain *d) {
    API_BEGIN
    check_final(d);
    CK(cudaStreamSynchronize(d->stream));
    API_END
}

static void locate(ax3d_domain *d, int tag, int fluid_part, float2 **base, int field, size_t *n) {
    if (tag < 0 || tag >= (int)d->points.size()) fail("Domain::getPoint || invalid point tag");
    if (field < 0 || field > 3) fail("ax3d || invalid field id");
    const HPoint &p = d->points[tag];
    if (!fluid_part) {
        if (p.s_idx < 0) fail("Point::getDispFourier || Incompatible point type.");
        *base = d->s_field[field].p + p.s_off;
        *n = (size_t)3 * (p.nu + 1);
    } else {
        if (p.f_idx < 0) fail("Point::getDispFourier || Incompatible point type.");
        *base = d->f_field[field].p + p.f_off;
        *n = (size_t)(p.nu + 1);
    }
}

int ax3d_get_point_field(ax3d_domain *d, int tag, int field, int fluid_part, float *out, int cap) {
    API_BEGIN
    check_final(d);
    float2 *b;
    size_t n;
    locate(d, tag, fluid_part, &b, field, &n);
    if ((size_t)cap < n) fail("ax3d_get_point_field || output buffer too small");
    CK(cudaStreamSynchronize(d->stream));
    CK(cudaMemcpy(out, b, n * sizeof(float2), cudaMemcpyDeviceToHost));
    API_END
}
int ax3d_set_point_field(ax3d_domain *d, int tag, int field, int fluid_part, const float *in, int n_in) {
    API_BEGIN
    check_final(d);
    float2 *b;
    size_t n;
    locate(d, tag, fluid_part, &b, field, &n);
    if ((size_t)n_in != n) fail("ax3d_set_point_field || size mismatch");
    CK(cudaStreamSynchronize(d->stream));
    CK(cudaMemcpy(b, in, n * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaStreamSynchronize(cudaStreamLegacy));   // pageable H2D: landed before anything later on the (non-blocking) compute streams
    API_END
}
int ax3d_field_size(ax3d_domain *d, int fluid_part, size_t *n) {
    API_BEGIN
    check_final(d);
    *n = fluid_part ? d->f_len : d->s_len;
    API_END
}
int ax3d_get_field_bulk(ax3d_domain *d, int field, int fluid_part, float *out, size_t cap) {
    API_BEGIN
    check_final(d);
    if (field < 0 || field > 3) fail("ax3d || invalid field id");
    const size_t n = fluid_part ? d->f_len : d->s_len;
    if (cap < n) fail("ax3d_get_field_bulk || output buffer too small");
    CK(cudaStreamSynchronize(d->stream));
    if (n) CK(cudaMemcpy(out, fluid_part ? d->f_field[field].p : d->s_field[field].p, n * sizeof(float2), cudaMemcpyDeviceToHost));
    API_END
}
int ax3d_set_field_bulk(ax3d_domain *d, int field, int fluid_part, const float *in, size_t n_in) {
    API_BEGIN
    check_final(d);
    if (field < 0 || field > 3) fail("ax3d || invalid field id");
    const size_t n = fluid_part ? d->f_len : d->s_len;
    if (n_in != n) fail("ax3d_set_field_bulk || size mismatch");
    CK(cudaStreamSynchronize(d->stream));
    if (n) CK(cudaMemcpy(fluid_part ? d->f_field[field].p : d->s_field[field].p, in, n * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaStreamSynchronize(cudaStreamLegacy));
    API_END
}

int ax3d_record_ground_motion(ax3d_domain *d, int nrec, const int *elem_tags, const float *phi, const float *weights, float *out) {
    API_BEGIN
    check_final(d);
    if (nrec <= 0) return 0;
    std::vector<RecvItem> items(nrec);
    for (int i = 0; i < nrec; ++i) {
        if (elem_tags[i] < 0 || elem_tags[i] >= (int)d->elems.size()) fail("PointwiseRecorder::record || invalid element tag");
        const HElem &E = d->elems[elem_tags[i]];
        items[i].elem = E.idx | (E.cls << 28);
        items[i].phi = phi[i];
        if (E.cls == CLS_F3D && recf_smem(E.nr) > (size_t)224 * 1024) fail("FluidElement::computeGroundMotion || Nr too large for the receiver kernel");
    }
    if ((size_t)nrec > d->rec_cap) {
        if (d->rec_host) cudaFreeHost(d->rec_host);
        CK(cudaMallocHost(&d->rec_host, (size_t)nrec * 3 * sizeof(float)));
        d->rec_items.alloc(nrec);
        d->rec_w.alloc((size_t)nrec * AX_NPE);
        d->rec_out.alloc((size_t)nrec * 3);
        d->rec_cap = nrec;
    }
    // receivers may sit in 1D or 3D elements: two launches over the respective descriptor arrays
    for (int c = 0; c < NCLS; ++c) {
        std::vector<RecvItem> sub;
        std::vector<int> where;
        size_t fsmem = 0;
        for (int i = 0; i < nrec; ++i) {
            if ((items[i].elem >> 28) == c) {
                RecvItem r = items[i];
                r.elem &= 0x0fffffff;
                sub.push_back(r);
                where.push_back(i);
                if (c == CLS_F3D) fsmem = std::max(fsmem, recf_smem(d->h_desc[c][r.elem].nr));
            }
        }
        if (sub.empty()) continue;
        // contiguous sub-launch: copy items + matching weights
        std::vector<float> w(sub.size() * AX_NPE);
        for (size_t k = 0; k < sub.size(); ++k) memcpy(&w[k * AX_NPE], weights + (size_t)where[k] * AX_NPE, AX_NPE * sizeof(float));
        CK(cudaMemcpyAsync(d->rec_items.p, sub.data(), sub.size() * sizeof(RecvItem), cudaMemcpyHostToDevice, d->stream));
        CK(cudaMemcpyAsync(d->rec_w.p, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice, d->stream));
        if (c == CLS_S1D || c == CLS_S3D) {
            k_ground_motion<<<(int)sub.size(), AX_REC_NT, 0, d->stream>>>(d->desc[c].p, d->rec_items.p, d->rec_w.p, d->s_field[AX3D_DISPL].p, d->rec_out.p, nullptr, 0);
        } else {
            if (fsmem > (size_t)48 * 1024) CK(cudaFuncSetAttribute(k_ground_motion_fluid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
            k_ground_motion_fluid<<<(int)sub.size(), AX_RECF_NT, fsmem, d->stream>>>(d->desc[c].p, d->rec_items.p, d->rec_w.p, d->plans.p, d->twpool.p,
                                                                                     d->geom.p, d->coef.p, d->f_field[AX3D_DISPL].p, d->rec_out.p, nullptr, 0);
        }
        d->launches++;
        CK(cudaMemcpyAsync(d->rec_host, d->rec_out.p, sub.size() * 3 * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
        CK(cudaStreamSynchronize(d->stream));
        for (size_t k = 0; k < sub.size(); ++k)
            for (int cc = 0; cc < 3; ++cc) out[where[k] * 3 + cc] = d->rec_host[k * 3 + cc];
    }
    API_END
}

/* Element::computeStrain (which = 0: out[nrec][6], Voigt strain in RTZ) / computeCurl (which = 1: out[nrec][3]) after
 * Element::forceTIso, as PointwiseRecorder::record uses them (PointwiseRecorder.cpp:96-135); SolidElement.cpp:219-345. */
static void record_strain_curl(ax3d_domain *d, int which, int nrec, const int *elem_tags, const float *phi, const float *weights, float *out) {
    check_final(d);
    if (nrec <= 0) return;
    const int nout = which ? 3 : 6;
    const char *fn = which ? "computeCurl" : "computeStrain";
    std::vector<RecvItem> items(nrec);
    for (int i = 0; i < nrec; ++i) {
        if (elem_tags[i] < 0 || elem_tags[i] >= (int)d->elems.size()) fail("PointwiseRecorder::record || invalid element tag");
        const HElem &E = d->elems[elem_tags[i]];
        // FluidElement::computeCurl is identically zero (FluidElement.cpp:308-311); computeStrain is built for Acoustic1D elements
        if (E.fluid && !which && E.rows > 1) fail("FluidElement::computeStrain || strain receivers in fluid elements with 3D material are not supported by the B200 path.");
        if (E.prt_rows) fail(std::string(E.fluid ? "FluidElement::" : "SolidElement::") + fn + " || strain / curl receivers in elements with particle relabelling are not supported by the B200 path.");
        items[i].elem = E.idx | (E.cls << 28);
        items[i].phi = phi[i];
    }
    DevBuf<RecvItem> di;
    DevBuf<float> dw, dout;
    for (int i = 0; i < nrec; ++i)   // fluid receivers: zero curl; the strain rows are filled by the fluid launch below
        if (d->elems[elem_tags[i]].fluid && which)
            for (int cc = 0; cc < nout; ++cc) out[i * nout + cc] = 0.f;
    for (int c : {CLS_S1D, CLS_S3D, CLS_F1D}) {
        if (c == CLS_F1D && which) continue;
        std::vector<RecvItem> sub;
        std::vector<int> where;
        std::vector<float> w;
        for (int i = 0; i < nrec; ++i)
            if ((items[i].elem >> 28) == c) {
                RecvItem r = items[i];
                r.elem &= 0x0fffffff;
                sub.push_back(r);
                where.push_back(i);
                w.insert(w.end(), weights + (size_t)i * AX_NPE, weights + (size_t)(i + 1) * AX_NPE);
            }
        if (sub.empty()) continue;
        di.upload(sub);
        dw.upload(w);
        dout.alloc(sub.size() * nout);
        if (c == CLS_F1D) k_strain_fluid1d<<<(int)sub.size(), AX_TILE * AX_NPE, 0, d->stream>>>(d->desc[c].p, di.p, dw.p, d->geom.p, d->coef.p, d->f_field[AX3D_DISPL].p, dout.p);
        else if (which) k_strain_curl<true><<<(int)sub.size(), AX_TILE * AX_NPE, 0, d->stream>>>(d->desc[c].p, di.p, dw.p, d->geom.p, d->s_field[AX3D_DISPL].p, dout.p);
        else k_strain_curl<false><<<(int)sub.size(), AX_TILE * AX_NPE, 0, d->stream>>>(d->desc[c].p, di.p, dw.p, d->geom.p, d->s_field[AX3D_DISPL].p, dout.p);
        d->launches++;
        CK(cudaGetLastError());
        std::vector<float> h(sub.size() * nout);
        CK(cudaStreamSynchronize(d->stream));
        CK(cudaMemcpy(h.data(), dout.p, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < sub.size(); ++k)
            for (int cc = 0; cc < nout; ++cc) out[where[k] * nout + cc] = h[k * nout + cc];
    }
}
int ax3d_record_strain(ax3d_domain *d, int nrec, const int *elem_tags, const float *phi, const float *weights, float *out) {
    API_BEGIN
    record_strain_curl(d, 0, nrec, elem_tags, phi, weights, out);
    API_END
}
int ax3d_record_curl(ax3d_domain *d, int nrec, const int *elem_tags, const float *phi, const float *weights, float *out) {
    API_BEGIN
    record_strain_curl(d, 1, nrec, elem_tags, phi, weights, out);
    API_END
}

/* Measured element costs (the reference's cost-measure pass, Mesh.cpp:412-588: each element's computeStiff is timed and the
 * times weight the second METIS partition).  One Domain::computeStiff is run `repeats` times with the clocks on:
 * elements of the fused launches report the SM cycles their CTA spent on them (clock64 around the element, gather prefetch
 * of the next one included); elements that go through k_elem1d or the split pipeline share the event-timed duration of
 * their launches in proportion to their number of 16-mode tiles (x Nr for 3D).  cost_us[elem tag] = microseconds of ONE SM. */
int ax3d_measure_costs(ax3d_domain *d, int repeats, double *cost_us, int nelem) {
    API_BEGIN
    check_final(d);
    if (nelem != (int)d->elems.size()) fail("Mesh::measure || size mismatch");
    if (repeats < 1) repeats = 1;
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, d->device));
    size_t tot = 0;
    for (int c = 0; c < NCLS; ++c) { d->cost_off[c] = tot; tot += d->h_desc[c].size(); }
    d->cost_buf.alloc(tot);
    std::vector<double> acc(d->elems.size(), 0.0);
    std::vector<unsigned> h(tot);
    const bool timers_were = d->timers;
    for (int r = 0; r <= repeats; ++r) {   // r = 0: warm-up
        d->cost_buf.zero();
        d->kstats.clear();
        d->timers = true;
        compute_stiff(d);
        d->timers = timers_were;
        CK(cudaStreamSynchronize(d->stream));
        if (r == 0) continue;
        CK(cudaMemcpy(h.data(), d->cost_buf.p, tot * sizeof(unsigned), cudaMemcpyDeviceToHost));
        // launches without in-kernel clocks: event time x resident CTAs share, split by work units
        double ms1d[NCLS] = {0, 0, 0, 0};
        for (const auto &kv : d->kstats) {
            if (kv.first == "k_elem1d<solid>") ms1d[CLS_S1D] += kv.second.ms;
            if (kv.first == "k_elem1d<fluid>") ms1d[CLS_F1D] += kv.second.ms;
            if (kv.first.rfind("split pipeline <solid>", 0) == 0) ms1d[CLS_S3D] += kv.second.ms;
            if (kv.first.rfind("split pipeline <fluid>", 0) == 0) ms1d[CLS_F3D] += kv.second.ms;
        }
        double units[NCLS] = {0, 0, 0, 0};
        auto unit_of = [&](const HElem &E) {
            const double tiles = (double)((E.nu + 1 + AX_TILE - 1) / AX_TILE);
            return E.rows > 1 ? tiles * E.nr : tiles;
        };
        for (const HElem &E : d->elems) {
            const bool in_fused = E.rows > 1 && d->h_desc[E.cls][E.idx].bucket == 0;
            if (!in_fused) units[E.cls] += unit_of(E);
        }
        for (size_t e = 0; e < d->elems.size(); ++e) {
            const HElem &E = d->elems[e];
            const bool in_fused = E.rows > 1 && d->h_desc[E.cls][E.idx].bucket == 0;
            if (in_fused) {
                // the fused launch indexes its elements from its first one
                int first = 0;
                for (const FusedLaunch &f : d->fused) if (f.cls == E.cls) first = f.first;
                acc[e] += (double)h[d->cost_off[E.cls] + (size_t)(E.idx - first)] / (double)khz * 1e3;
            } else if (units[E.cls] > 0) {
                acc[e] += ms1d[E.cls] * 1e3 * (double)d->num_sm * unit_of(E) / units[E.cls];
            }
        }
    }
    for (size_t e = 0; e < d->elems.size(); ++e) cost_us[e] = acc[e] / repeats;
    d->cost_buf.release();
    d->kstats.clear();
    API_END
}

/* The radix sequence the library plans for an azimuthal transform of length nr (host only, no device needed): radices[0 .. *nstages)
 * in DIF order, their product is nr.  Lucky numbers only (PreloopFFTW.cpp:59-99); fails for a prime factor above 13. */
int ax3d_fft_plan(int nr, int *radices, int cap, int *nstages) {
    API_BEGIN
    if (nr < 1) fail("ax3d::plan || nr must be positive");
    const RadixList rl = choose_radices_plan(nr);
    if (rl.n < 0) fail("ax3d::plan || Nr = " + std::to_string(nr) + " is not a lucky number (prime factor > 13 or too many stages; PreloopFFTW.cpp:59-99)");
    if (rl.n > cap) fail("ax3d::plan || output buffer too small");
    for (int k = 0; k < rl.n; ++k) radices[k] = rl.r[k];
    *nstages = rl.n;
    API_END
}

int ax3d_launch_count(ax3d_domain *d, long long *n) {
    API_BEGIN
    *n = d->launches;
    API_END
}
int ax3d_work_per_step(ax3d_domain *d, long long *w) {
    API_BEGIN
    check_final(d);
    *w = d->work;
    API_END
}
int ax3d_algorithmic_bytes(ax3d_domain *d, double out[3]) {
    API_BEGIN
    check_final(d);
    for (int i = 0; i < 3; ++i) out[i] = d->alg_bytes[i];
    API_END
}
int ax3d_enable_timers(ax3d_domain *d, int on) {
    API_BEGIN
    d->timers = on != 0;
    API_END
}
/* Per-kernel statistics gathered while the timers are on (ax3d_enable_timers): entry `index` in name order -> name, summed
 * device time (ms), launches, summed algorithmic bytes of those launches.  *count = number of entries (index < 0: only that).
 * reset != 0 clears the table after the read of the LAST entry (index == count - 1) or when index < 0. */
int ax3d_kernel_stats(ax3d_domain *d, int index, char *name, int name_cap, double *ms_total, long long *launches, double *bytes_total,
                      int *count, int reset) {
    API_BEGIN
    check_final(d);
    const int n = (int)d->kstats.size();
    if (count) *count = n;
    if (index >= 0) {
        if (index >= n) fail("ax3d_kernel_stats || index out of range");
        auto it = d->kstats.begin();
        std::advance(it, index);
        if (name && name_cap > 0) { strncpy(name, it->first.c_str(), (size_t)name_cap - 1); name[name_cap - 1] = 0; }
        if (ms_total) *ms_total = it->second.ms;
        if (launches) *launches = it->second.n;
        if (bytes_total) *bytes_total = it->second.bytes;
    }
    if (reset && (index < 0 || index == n - 1)) d->kstats.clear();
    API_END
}
int ax3d_get_timers(ax3d_domain *d, double out_ms[4], int reset) {
    API_BEGIN
    for (int i = 0; i < 4; ++i) {
        out_ms[i] = d->timer_ms[i];
        if (reset) d->timer_ms[i] = 0;
    }
    API_END
}

}   // extern "C"
