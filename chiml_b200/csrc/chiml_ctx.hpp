// Internal context of the B200 FDTD engine (not part of the C ABI).
//
// Device data layout (DESIGN.md "Data layout in HBM"):
//   * every field component is one SoA buffer, logical index (x, z, y) with x fastest, rows padded to a
//     multiple of 16 doubles (128 B) -> physical index  x + px * (z + lz * y)
//   * per component a uint16 "cell info" plane painted from the reference's update lists:
//       low byte  = material class (index into a small table of prefactors / pole constants)
//       high byte = flags: which of the reference's per-run operations apply to the cell
//   * CPML psi arrays exist only inside the slabs (the axis normal to the slab is compressed)
//   * polarisation (pole) state exists only on the x-span of dispersive cells of each grid row
#pragma once

#include <cuda_runtime.h>

#include <array>
#include <cstdint>
#include <string>
#include <map>
#include <string>
#include <vector>

#include "../../include/chiml_gpu.h"

namespace chiml {

constexpr int MAX_CHI = 4;        // chiral poles per material class
constexpr int MAX_POLES = 12;     // poles per material class (largest built-in metal has 6)
constexpr int MAX_SOURCES = 32;
constexpr int MAX_CLASSES = 255;

// cell-info flags (high byte)
constexpr uint16_t F_CURL  = 0x0100;  // interior curl run covers the cell (upE_/upH_/upD_)
constexpr uint16_t F_ISD   = 0x0200;  // ... and accumulates into D (upD_)
constexpr uint16_t F_PG0   = 0x0400;  // CPML part 0 grid term   (updateListGrid_k_)
constexpr uint16_t F_PS0   = 0x0800;  // CPML part 0 psi update  (updateListPsi_j_)
constexpr uint16_t F_PG1   = 0x1000;  // CPML part 1 grid term   (updateListGrid_j_)
constexpr uint16_t F_PS1   = 0x2000;  // CPML part 1 psi update  (updateListPsi_k_)
constexpr uint16_t F_D2E   = 0x4000;  // isotropic pole update + D->E (upLorD_)
constexpr uint16_t F_ORD2E = 0x8000;  // oriented-dipole D->E (upOrDipD_)
constexpr uint16_t CLS_MASK = 0x00FF;

struct ClassEntry
{
    double pf1, pf2;            // prefactors[1], prefactors[2] of the run
    double inv_eps;             // 1.0/eps           (DtoU dscal factor)
    double neg_inv_eps;         // -1.0/eps          (DtoU daxpy factor)
    double neg_half_inv_eps;    // -0.5/eps          (orDipDtoU daxpy factor)
    int npoles;
    int pad;
    double alpha[MAX_POLES], xi[MAX_POLES], gamma[MAX_POLES];
    double dip[MAX_POLES][3];
    // chiral poles of the class (UpdateChiral, UTIL/FDTD_up_eq.cpp:64-111; chiDtoU :920-925)
    int nchi, pad2;
    double chi_alpha[MAX_CHI], chi_xi[MAX_CHI], chi_g8[MAX_CHI], chi_gp8[MAX_CHI];   // chiAlpha, chiXi, chiGamma / 8.0, chiGammaPrev / 8.0
    double chi_fac;             // -1.0 / (-1.0 * eps) for E (D2E passes -eps), -1.0 / mu for H
};

// compact storage of per-row x-spans (pole state)
struct SpanTable
{
    int32_t* d_xmin = nullptr;   // per logical row (z + lz*y): first x of the span, or -1
    int32_t* d_xmax = nullptr;
    int64_t* d_base = nullptr;   // offset of the span in the pool
    int64_t total = 0;
    std::vector<int32_t> h_xmin, h_xmax;
    std::vector<int64_t> h_base;
    int32_t* d_rows = nullptr;   // compact list of rows with a span
    int nrows_used = 0, max_width = 0;
};

struct PmlPartDev
{
    int present = 0;
    int has_psi = 0;
    int axis = -1;               // derivative axis: 0 x, 1 y, 2 z
    int vfield = -1;             // ChimlField driving this part
    long off_logical = 0;        // indOff - ind (logical)
    double Db = 0.0;
    int nact = 0;                // number of coordinates along `axis` with an active psi
    long psi_pitch = 0, psi_count = 0;
    double* d_F = nullptr;       // per coordinate along axis
    double* d_b = nullptr;
    double* d_c = nullptr;
    int32_t* d_cmap = nullptr;   // coordinate -> compact coordinate (or -1)
    double* d_psi = nullptr;
    std::vector<int32_t> h_cmap;
};

struct SourceDev { int field; int32_t loc[3], sz[3]; };
struct DetectorDev
{
    int field; int32_t loc[3], sz[3]; int every;
    size_t sample_len = 0;
    // ring of samples: sample s (counted from the t = 0 sample) lives in slot s % cap; samples [base, count) are retained
    double* d_ring = nullptr; size_t cap = 0; size_t count = 0; size_t base = 0;
};

// kernels of the step loop, for launch / time / algorithmic-byte accounting
enum KernelKind { K_E_FAST = 0, K_E_UNIFORM, K_E_GENERAL, K_H_FAST, K_H_UNIFORM, K_H_GENERAL, K_ORDIP_POLES, K_SOURCE, K_DETECTOR, K_EMIT_ADDP, K_EMIT_DENSITY, K_EMIT_POP, K_HALO_PUSH, K_HALO_WAIT, K_DFT, K_STEPS_2D, K_WRAP, K_TFSF, K_WRAP_BLOCH, K_PREV_COPY, K_NKINDS };
struct KernelStat { int64_t launches = 0; double ms_total = 0.0; double alg_bytes = 0.0; int64_t timed = 0; };

// one parallelQE object on the device (chiml_emitters.cuh)
struct EmitterDev
{
    ChimlEmitterDesc d{};            // scalars; the pointer members are not used after add_emitters
    int n2 = 0, pz = 0;
    size_t pbox = 0;
    std::vector<double> h_h0, h_weight, h_mu, h_gam_val, h_eps;
    std::vector<int32_t> h_gam_ptr, h_gam_col, h_loc, h_pop_level;
    double *d_h0 = nullptr, *d_mu = nullptr, *d_gam_val = nullptr, *d_eps = nullptr;
    int32_t *d_gam_ptr = nullptr, *d_gam_col = nullptr, *d_loc = nullptr;
    double* d_P[3] = {};
    double* d_rho = nullptr;
    double* d_f[4] = {};             // ring of derivative histories
    int fbase = 0;                   // d_f[(fbase + k) % 4] holds d rho/dt at step n-k
    int mu_present[3] = {};
    double* d_pop_partial = nullptr; int nblocks = 0;
    int gam_maxrow = 0;              // longest row of gam_
    int group = 0;                   // lanes per emitter of k_emit_density_g (AoS state), or 0: k_emit_density, one thread per emitter (SoA state)
    double* d_pop = nullptr; size_t pop_cap = 0, pop_n = 0, pop_base = 0;   // ring like DetectorDev's: [npop][pop_cap] complex
    long tstep = 0;
};

// ---- y-slab halo (chiml_halo.cuh): flags the neighbours write into this context's memory, one 32-bit step counter each
// (HF_EY_FROM_UPPER / HF_EY_ACK: emitters on a ring -- the two seam rows of Ey travel at the END of a step, each side releasing its row first)
// (periodic ring of slabs: HF_H_FROM_UPPER = slab 0's Hx, Hz row 1 has arrived in the last slab's wrap row; HF_SEAM_ACK = the last slab has released that row for the step: read by the E half step before, its own discarded update through)
enum HaloFlag { HF_H_FROM_LOWER = 0, HF_OP_FROM_UPPER, HF_E_FROM_UPPER, HF_EY_FROM_LOWER, HF_QP_FROM_UPPER, HF_ERROR, HF_H_FROM_UPPER, HF_SEAM_ACK, HF_EY_FROM_UPPER, HF_EY_ACK, HF_NFLAGS = 16 };
constexpr size_t IPC_GRANULE = 2u << 20;     // exported buffers are whole 2 MiB allocations (never sub-allocated by the driver)

struct HaloPeer                   // one neighbouring slab, as mapped into this process
{
    bool present = false;
    int ly = 0;                   // its ghost-inclusive slab height
    double* field[6] = {};        // its E / H arrays (logical origin, like ChimlCtx::d_field)
    double* oPy_ghost[MAX_POLES] = {};   // its dense ghost row of node P_y (only the slab below needs ours: we write theirs)
    std::vector<double*> emitPy;  // its emitter P_y boxes, per emitter set (nullptr when that set does not exchange)
    std::vector<int> emit_bn1;    // box_n[1] of those sets
    int* flags = nullptr;         // its flag array
    std::vector<void*> opened;    // bases returned by cudaIpcOpenMemHandle
};

// one stored field of a flux / frequency detector (running DFT)
struct DftDev
{
    int field = 0, group = 0, every = 1, nfreq = 0, npts = 0, stride = 1;
    size_t nlines = 0, acc_len = 0;
    std::vector<ChimlDftLine> h_lines;
    ChimlDftLine* d_lines = nullptr;
    double *d_re = nullptr, *d_im = nullptr;
};

// one TFSF surface (chiml_gpu_add_tfsf_surface): host copies until commit, device copies after
struct TfsfDev
{
    ChimlTfsfSurface s{};
    std::vector<int32_t> h_pairs_D, h_pairs_U;
    std::vector<double> h_ep_mu;
    int32_t *d_pairs_D = nullptr, *d_pairs_U = nullptr;
    double* d_ep_mu = nullptr;
};

struct HostList { std::vector<ChimlRun> runs; };
struct HostPml { int present = 0, has_psi = 0; std::vector<ChimlPsiParams> psi; std::vector<ChimlGridParams> grid; };
struct HostObj { int npoles = 0, use_or_dip = 0; std::vector<double> alpha, xi, gamma, dip; int nmag = 0; std::vector<double> malpha, mxi, mgamma; int nchi = 0; std::vector<double> calpha, cxi, cgamma, cgprev; };

} // namespace chiml

struct ChimlCtx
{
    ChimlGridDesc g{};
    int device = 0;
    int lx = 0, ly = 0, lz = 0;
    long px = 0;                 // padded row pitch (doubles)
    long plane = 0;              // px * lz
    size_t nphys = 0;            // px * lz * ly
    size_t nlogical = 0;
    bool committed = false;
    std::string err;
    int march_fast = 0, march_uniform = 0;   // chiml_gpu_set_march: planes per column of k_fast / k_uniform (0 = automatic)

    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // state
    double* d_field[CHIML_NFIELDS] = {};       // points `guard` doubles into d_field_base
    double* d_field_base[CHIML_NFIELDS] = {};
    size_t guard = 0;
    bool have_off[6] = {};
    uint16_t* d_info[6] = {};
    chiml::ClassEntry* d_cls[6] = {};
    double2* d_pf[6] = {};       // {pf1, pf2} per class (interior fast path)
    // compact tile lists per family ([0] E, [1] H) and kind ([0] fast, [1] uniform, [2] general); element type chiml::TileRec
    void* d_tiles[2][3] = {};
    unsigned ntiles[2][3] = {};
    std::vector<chiml::ClassEntry> h_cls[6];
    int ncls[6] = {};
    chiml::PmlPartDev pml[6][2];

    // isotropic poles: per component pools [pole][cur/prev]; [0..2] electric (P of Ex..Ez), [3..5] magnetic (M of Hx..Hz, has_B)
    chiml::SpanTable span[6];
    int npoles_comp[6] = {};
    double* d_P[6][chiml::MAX_POLES][2] = {};
    int has_B = 0, pml_on_B = 0;       // chiml_gpu_set_magnetic
    // chiral media: chiral pole pools over the same spans, previous-field copies, the rows copied after each half step
    double* d_chi[6][chiml::MAX_CHI][2] = {};
    double* d_prev[6] = {};            // prevE_[0..2], prevH_[0..2] (logical origin, guarded like d_field)
    double* d_prev_base[6] = {};
    int nchi_comp[6] = {};
    long chi_off[6][3] = {};           // logical ind_i - ind, ind_j - ind, ind_k - ind of the CHID list of the component
    std::vector<int32_t> h_prev_rows;  // copy2PrevFields_: 4 per row {length, x, y, z}
    int4* d_prev_rows = nullptr; size_t n_prev_rows = 0;
    bool has_chi = false;
    int pcur = 0;                // which of the two buffers currently holds P (the other holds prevP)

    // oriented-dipole poles at nodes
    uint16_t* d_info_node = nullptr;
    chiml::ClassEntry* d_cls_node = nullptr;
    int ncls_node = 0;
    chiml::SpanTable span_node;
    int nordip = 0;              // pole grids of the whole grid: max(this slab's node list, chiml_gpu_set_ordip_pole_count)
    int nordip_global = 0;
    double* d_oP[3][chiml::MAX_POLES][2] = {};
    const double* h_dipg[3][chiml::MAX_POLES] = {};    // chiml_gpu_set_dip_grid: the caller's dipP_[c][p] on the logical grid, read at commit (not copied:
                                                       // six such grids of a C5-sized run are 25 GB)
    double* d_dipg[3][chiml::MAX_POLES] = {};          // ... gathered over the node spans (same index as d_oP), or nullptr
    bool has_dipg = false;
    long node_off[3] = {};       // logical offsets ind_i-ind, ind_j-ind, ind_k-ind of the node list
    long ordip_off[3] = {};      // logical offset ind_i-ind of upOrDipD_[c]

    // curl offsets per component (logical): ind_j - ind, ind_k - ind
    long off_j[6] = {}, off_k[6] = {};

    std::vector<chiml::SourceDev> sources;
    double* d_src_amp = nullptr; size_t src_amp_cap = 0;
    std::vector<chiml::DetectorDev> detectors;
    std::vector<chiml::EmitterDev> emitters;
    std::vector<chiml::DftDev> dfts;
    std::vector<int> dft_group_nfreq;            // nfreq of every group, in group order
    double* d_tw = nullptr; size_t tw_cap = 0;   // twiddles of the current step_n_dft call

    // host-side copies of the setup until commit
    chiml::HostList lists[6][6];
    chiml::HostPml hpml[6][2];
    std::vector<chiml::HostObj> objs;

    // complex fields: this context holds the real parts, `imag` the imaginary parts (chiml_gpu_bind_imag); the pair is stepped through this one
    ChimlCtx* imag = nullptr;
    ChimlCtx* real_part = nullptr;               // set in the imaginary part
    bool is_imag_part = false;
    double k_point[3] = {0.0, 0.0, 0.0};
    // TFSF surfaces and the incident-line table of the current chiml_gpu_step_n_tfsf call
    std::vector<chiml::TfsfDev> tfsf;
    double* d_tfsf_incd = nullptr; size_t tfsf_incd_cap = 0; size_t tfsf_per_step = 0;
    bool tfsf_table_ready = false;
    // periodic boundaries (chiml_gpu_set_periodic): wrap copies per component after its half step
    ChimlWrap wrap[6] = {};
    bool has_wrap[6] = {};
    bool periodic = false;

    void* d_tmaps = nullptr;                     // TMA descriptors of the field and psi arrays (chiml_kernels.cuh TMAP_*), 3-D grids
    // persistent multi-step kernel of 2-D grids (chiml_persist.cuh)
    void* d_persist_sa = nullptr;                // 3 StepArgs
    int persist_blocks = -1;                     // resident grid size, 0 = not available, -1 = not asked yet
    int persist_mode = -1;                       // -1 automatic (by grid size), 0 / 1 chiml_gpu_set_persistent

    long long step_count = 0;
    int64_t launches = 0;
    size_t dev_bytes = 0;

    // y-slab halo
    cudaStream_t hstream = nullptr;
    cudaEvent_t ev_main = nullptr, ev_push = nullptr;
    bool push_pending = false;
    int* d_flags = nullptr;                      // HF_NFLAGS ints, exported
    bool ring = false;                           // periodic run on several slabs: the slabs form a ring (slab 0 <-> slab nranks - 1)
    std::map<std::string, void*> ipc_cache;      // IPC handles already opened (a ring of two slabs names the same peer twice)
    unsigned* d_push_counter = nullptr;          // block counters of the push kernels
    int push_slot = 0;
    double* d_oPy_ghost[chiml::MAX_POLES] = {};  // dense (lx * lz) ghost row ny+1 of node P_y per pole, written by the slab above
    chiml::HaloPeer lower, upper;
    bool halo_bound = false;
    unsigned nbound[2][3] = {};                  // tiles of each list that touch a slab-boundary row (sorted to the front)

    // per-kernel device timing (chiml_gpu_set_kernel_timing / chiml_gpu_kernel_stat)
    bool timing = false;
    long long stat_step0 = 0;                                 // step_count at the last chiml_gpu_reset_kernel_stats
    chiml::KernelStat kstat[chiml::K_NKINDS];
    std::vector<cudaEvent_t> ev_pool;                         // recycled events
    std::vector<std::array<cudaEvent_t, 2>> ev_pending[chiml::K_NKINDS];
};
