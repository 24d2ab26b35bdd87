/* chiml_gpu.h -- C ABI of the B200 (sm_100a) time-stepping engine for chiML's FDTD hot path.
 *
 * This is the drop-in boundary: a maintainer of the reference replaces the body of
 * parallelFDTDFieldBase<double>::step() (reference src/FDTD_MANAGER/parallelFDTDField.hpp:1228-1303)
 * by calls into this library, handing over the data structures the reference constructor has
 * already built -- its run-length "update lists", CPML parameter lists, per-object pole constants,
 * source boxes and detector boxes -- unchanged.  Plain C types only; no exceptions cross the
 * boundary: every function returns 0 on success or a non-zero ChimlStatus, and
 * chiml_gpu_last_error() returns the message of the last failure on that context.
 *
 * Index space.  Every `ind` below is the reference's own local linear index
 *     ind = x + ln[0] * ( z + ln[2] * y )           (GRID/parallelGrid.hpp:363,563)
 * over the ghost-inclusive local extents ln = (nx+2, ny_loc+2, nz+2) (2-D: ln[2] = 1).  The device
 * stores rows padded to 128 bytes; that is invisible here.
 *
 * Threading: one host thread per context; contexts are independent.  All device work of a context
 * runs on streams it owns; calls are asynchronous unless stated otherwise.
 */
#ifndef CHIML_GPU_H
#define CHIML_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ChimlCtx ChimlCtx;

typedef enum ChimlStatus
{
    CHIML_OK = 0,
    CHIML_ERR_ARG = 1,         /* bad argument / inconsistent list */
    CHIML_ERR_CUDA = 2,        /* CUDA runtime error (message in chiml_gpu_last_error) */
    CHIML_ERR_UNSUPPORTED = 3, /* list describes something outside the supported hot path */
    CHIML_ERR_STATE = 4,       /* call order (e.g. step before commit) */
    CHIML_ERR_NO_DEVICE = 5    /* no CUDA device: there is deliberately no CPU fallback */
} ChimlStatus;

/* field components, in the reference's order E_[0..2], H_[0..2], D_[0..2] (parallelFDTDField.hpp:173-176) */
typedef enum ChimlField
{
    CHIML_EX = 0, CHIML_EY = 1, CHIML_EZ = 2,
    CHIML_HX = 3, CHIML_HY = 4, CHIML_HZ = 5,
    CHIML_DX = 6, CHIML_DY = 7, CHIML_DZ = 8,
    CHIML_BX = 9, CHIML_BY = 10, CHIML_BZ = 11,   /* B_[0..2]: exist after chiml_gpu_set_magnetic(has_B = 1) */
    CHIML_NFIELDS = 12
} ChimlField;

/* CompCell.pol / size_z selection of parallelFDTDField.hpp:391,418 */
typedef enum ChimlMode { CHIML_MODE_TE = 0 /* Ex,Ey,Hz */, CHIML_MODE_TM = 1 /* Ez,Hx,Hy */, CHIML_MODE_3D = 2 } ChimlMode;

typedef struct ChimlGridDesc
{
    int32_t mode;       /* ChimlMode */
    int32_t ln[3];      /* ghost-inclusive local extents (parallelGrid::ln_vec_) */
    double  d[3];       /* grid spacing d_ */
    double  dt;         /* time step dt_ */
    int32_t has_D;      /* D_ grids exist ("disp", parallelFDTDField.hpp:375-402) */
    int32_t pml_on_D;   /* dielectricMatInPML_: the E-side CPML acts on D (parallelFDTDField.cpp:68-77) */
    int32_t n_objects;  /* objArr_.size() (object 0 is the vacuum background) */
    int32_t rank;       /* y-slab index of this context (mpiInterface::rank) */
    int32_t nranks;     /* number of y-slabs */
} ChimlGridDesc;

/* One x-contiguous run: layout-identical to the reference's
 * std::pair<std::array<int,6>, std::array<double,4>> (UTIL/typedefs.hpp:14,
 * filled by populateUpLists, parallelFDTDField.hpp:628-649):
 *   {n, ind, ind_i, ind_j, ind_k, obj}  {1.0, -dt/(eps d_j), -dt/(eps d_k), eps}          */
typedef struct ChimlRun
{
    int32_t n, ind, ind_i, ind_j, ind_k, obj;
    double  pf[4];
} ChimlRun;

/* layout-identical to updatePsiParams / updateGridParams (PML/parallelPML.hpp:18-38) */
typedef struct ChimlPsiParams  { int32_t transSz, stride, ind, indOff; double b, c; } ChimlPsiParams;
typedef struct ChimlGridParams { int32_t nAx, stride, ind, indOff; double Db, DbField; } ChimlGridParams;

typedef enum ChimlListKind
{
    CHIML_LIST_U      = 0, /* upE_[c] / upH_[c]: direct curl update                 (updateE/updateH :1308-1323) */
    CHIML_LIST_D      = 1, /* upD_[c]: curl accumulated into D                       (updateD :1338-1343)        */
    CHIML_LIST_LORD   = 2, /* upLorD_[c]: isotropic pole update + D->E               (updatePolE :1355-1361, D2E :1456-1459) */
    CHIML_LIST_ORDIPD = 3, /* upOrDipD_[c]: oriented-dipole D->E (node->edge average) (D2E :1465-1466)            */
    CHIML_LIST_ORDIPP = 4, /* upOrDipP_: oriented-dipole pole update at nodes; comp ignored (updatePolE :1350-1354) */
    CHIML_LIST_CHID   = 5  /* upChiD_[c] (comp 0..2) / upChiB_[c] (comp 3..5): the cells of chiral objects -- achiral poles, chiral poles driven by the
                              8-point average of the other family's same component and of its previous value, D->E / B->H with both (updateChiE /
                              updateChiH :1392-1447, D2E :1460-1464, B2H :1484-1488); 3-D grids */
} ChimlListKind;

/* ---- life cycle ------------------------------------------------------------------------------ */
int  chiml_gpu_device_count(void);
int  chiml_gpu_create(const ChimlGridDesc* desc, int device, ChimlCtx** out);
void chiml_gpu_destroy(ChimlCtx* ctx);
const char* chiml_gpu_last_error(const ChimlCtx* ctx); /* ctx may be NULL: last create() failure */

/* ---- setup (before commit) -------------------------------------------------------------------- */
/* comp: 0..2 = Ex,Ey,Ez   3..5 = Hx,Hy,Hz */
int chiml_gpu_set_update_list(ChimlCtx* ctx, int kind, int comp, const ChimlRun* runs, size_t n);

/* Pole constants of object `obj`: Obj::alpha()/xi()/gamma() after setUpConsts(dt) (OBJECTS/Obj.cpp:299-371).
 * dip: 3*npoles doubles (x,y,z per pole) for oriented-dipole objects with a position-independent
 * dipole direction (ISOTROPIC -> 1,1,1; UNIDIRECTIONAL -> dipE; parallelFDTDField.hpp:998-1020), or NULL. */
int chiml_gpu_set_object(ChimlCtx* ctx, int obj, int npoles, const double* alpha, const double* xi, const double* gamma,
                         int use_or_dip, const double* dip);

/* One half of parallelCPML<T>::updateGrid() (PML/parallelPML.hpp:693-697) for component comp:
 *   part 0 = addGrid_j_(updateListGrid_k_, updateListPsi_j_, grid_i, psi_j, grid_k)
 *   part 1 = addGrid_k_(updateListGrid_j_, updateListPsi_k_, grid_i, psi_k, grid_j)
 * has_psi = 0 selects pmlUpdateFxnReal::addGridOnly (PML/parallelPML.cpp:23-30). */
int chiml_gpu_set_cpml(ChimlCtx* ctx, int comp, int part, int has_psi,
                       const ChimlPsiParams* psi, size_t npsi, const ChimlGridParams* grid, size_t ngrid);

/* Soft source box in local ghost-inclusive coordinates (SalveSource of SOURCE/parallelSource.hpp,
 * built by genDatStruct, parallelSourceNormal.hpp:100): field[box] += amp each step, where the host
 * passes amp = dt * Re(sum_p pulse_p(t)) (parallelSourceNormal.cpp:15-37).  Returns the source slot. */
int chiml_gpu_add_source(ChimlCtx* ctx, int field, const int32_t loc[3], const int32_t sz[3], int* slot);

/* Time-domain detector sampling (DTC/parallelStorageDTC.cpp:17-44): every `every` steps after the
 * step, and once at commit time (t = 0, parallelFDTDField.cpp:832-833), the raw values of `field`
 * in the box loc..loc+sz (local ghost-inclusive coordinates) are appended to a device ring that
 * chiml_gpu_read_detector[_range] reads and chiml_gpu_consume_detector drains.  The Yee-offset averaging and SI factors stay on
 * the host. */
int chiml_gpu_add_detector(ChimlCtx* ctx, int field, const int32_t loc[3], const int32_t sz[3], int every, int* slot);

/* Quantum-emitter cells of one parallelQE object (ML/parallelQE.hpp): every listed grid node carries, per level system
 * (= per Hamiltonian of levelSys_), an N x N density matrix propagated by PCABAM4 (:751-770) under H = H0 - mu.E (ML/Hamiltonian.cpp:59-69)
 * with the relaxation super-operator gam_ (:727-744), and feeds P = na Re<rho|mu> back into E (addQE :682-718).  Emitters are owned by the
 * slab that owns their node (the reference spreads them over all ranks, :394-420, and ships E / P boxes around; results are identical).
 * All complex arrays are (re, im) pairs, matrices row-major as the reference stores them. */
typedef struct ChimlEmitterDesc
{
    int32_t nlevel;            /* N = nlevel_ */
    int32_t nsys;              /* levelSys_.size() */
    int32_t nemit;             /* emitters (nodes) of this object inside this slab */
    int32_t box_lo[3];         /* local ghost-inclusive coordinates of the box corner = emitter-box minimum minus one node
                                  (SendEFieldRecvPField::loc_ after the procLoc shift, parallelQE.hpp:518-570) */
    int32_t box_n[3];          /* emitter bounding box n_vec = max - min + 1 (2-D: box_n[2] = 1) */
    double  dt;                /* dt_ */
    double  inv_hbar;          /* imag(one_over_hbar_) = 1/hbar_ (:201) */
    double  na;                /* na_ : molecular density * a^3 */
    const double* h0;          /* nsys * N*N complex : Hamiltonian::h0_ of every level system */
    const double* weight;      /* nsys : initial rho_00 (energyWeights_[q].second, density.hpp:57-60) */
    const double* mu;          /* 3 * N*N complex : x_, y_, z_expectation_ (dipole matrices times couplings, Hamiltonian.cpp:28-41) */
    const int32_t* gam_ptr;    /* N*N + 1 : CSR row pointers of gam_ ...                                                  */
    const int32_t* gam_col;    /* ... columns and values in the ITERATION ORDER of the reference's unordered_map rows (:740-742) */
    const double*  gam_val;
    const int32_t* loc;        /* 3 * nemit : emitter nodes relative to the box minimum (Density::x(), y(), z() after :477-484), in the
                                  reference's order */
    const double*  eps;        /* (box_n[0]+2)*(box_n[1]+2)*pz : eps_ (epsRelOrDip_) over the P box, index x + (n0+2)*(z + pz*y),
                                  pz = box_n[2]+2 in 3-D, 2 in 2-D (the shape of P_, :470) */
    int32_t npop;              /* population detectors (QEPopDtc) */
    const int32_t* pop_level;  /* npop : flat index into rho (QEPopDtc::level_, QEPopDtc.hpp:67) */
    int32_t pop_every;         /* timeInt_ in steps */
    int32_t npoints;           /* QEPopDtc::npoints_ = emitters of the whole object (all slabs) */
    int32_t object;            /* index of the object in qeArr_: with several slabs it pairs the sets of ONE object across a slab boundary
                                  (two emitter species filling the same region have equal boxes and differ only in this) */
} ChimlEmitterDesc;
int chiml_gpu_add_emitters(ChimlCtx* ctx, const ChimlEmitterDesc* desc, int* slot);

/* Running discrete Fourier transform of one stored field of a flux / frequency detector (parallelStorageFreqDTCReal::fieldIn,
 * DTC/parallelStorageFreqDTC.cpp:21-30; driven by parallelFluxDTC::fieldIn, DTC/parallelFlux.hpp:296-312): every `every` steps, after
 * the step, for every line l and point i < npts and frequency f < nfreq
 *     acc_re[lines[l].out + f + nfreq*i] += tw_re[f] * field[lines[l].ind + i*stride]      (the two dger_ rank-1 updates)
 *     acc_im[lines[l].out + f + nfreq*i] += tw_im[f] * field[lines[l].ind + i*stride]
 * with tw = exp(-i freq t) computed by the HOST for the time after the step (so that its rounding is the host's) and passed to
 * chiml_gpu_step_n_dft.  `group` = index of the flux object: all stored fields of one object share its twiddles. */
typedef struct ChimlDftLine { int32_t ind, out; } ChimlDftLine;   /* the (grid index, accumulator index) pairs of fInGridInds_ */
int chiml_gpu_add_dft(ChimlCtx* ctx, int field, int group, int every, int nfreq, int npts, int stride,
                      const ChimlDftLine* lines, size_t nlines, size_t acc_len, int* slot);

/* Periodic boundaries (CompCell.PBC with real fields, i.e. k-point = 0): after its half step every E / H component gets the wrap
 * copies of applyBC1Proc (UTIL/FDTD_up_eq.cpp:1058-1116) -- every ghost cell of the box [0, xmax] x [0, ymax] x [zmin-1, zmax]
 * receives the value of its periodic image inside (x = 0 <- xmax-1, x = xmax <- 1, likewise y and z; on 2-D grids, zmin = 0, rows
 * first, then columns over rows 1 .. ymax).  The seven numbers are the arguments the reference passes to applBCH_[c] / applBCE_[c]
 * (FDTD_MANAGER/parallelFDTDField.hpp:1267-1269,1285-1287; yHPBC_/yEPBC_/zMinPBC_/zMaxPBC_ from parallelFDTDField.cpp:153-171,
 * 322-336, parallelFDTDField.hpp:444-445): for a component trimmed by one point along an axis (fieldEnd) the last, never updated
 * point of that axis is the upper image.  comp 0..5 = Ex..Hz.  Oriented-dipole media are refused together with it.
 * A slab of several (desc.nranks > 1) passes ymax = ny = -1: the engine then wraps the x / z ghost cells of its owned rows only (applyBCProcMid,
 * UTIL/FDTD_up_eq.cpp:1036-1061, which the reference takes on every rank), and the y direction is the ring of ghost-row pushes: chiml_gpu_halo_bind
 * takes the blob of slab nranks - 1 as slab 0's lower neighbour and vice versa; the last slab sends its top owned row of Hx, Hz (ly - 3: these
 * components are one row short in y) upward across the seam and receives slab 0's row 1 of Hx, Hz in its wrap row ly - 2 (chiml_b200/slab.py).
 * Emitters: the two seam rows of Ey travel at the end of a step (the reference updates its emitters before it wraps E), polarisation boxes do not
 * cross the seam.  Interior results equal the single-slab run's bit for bit.  Covered there: real fields, no magnetic / chiral media, no TFSF. */
typedef struct ChimlWrap { int32_t nx, ny, nz, xmax, ymax, zmin, zmax; } ChimlWrap;
int chiml_gpu_set_periodic(ChimlCtx* ctx, int comp, const ChimlWrap* wrap);

/* Magnetic-dispersive media: the B / H / M mirror of the D / E / P path (updateMagH, updateB, B2H of step(), FDTD_MANAGER/parallelFDTDField.hpp:1230,
 * 1234, 1264; :1328-1333, :1370-1387, :1478-1500).  has_B: B_ grids exist (:403-406, :429-432); pml_on_B: magMatInPML_, the H-side CPML acts on B
 * (parallelFDTDField.cpp:229-246).  With has_B the H components accept the list kinds CHIML_LIST_D (upB_[c]: curl accumulated into B) and
 * CHIML_LIST_LORD (upLorB_[c]: magnetic pole update with H^n at the start of the step, then H = (B - sum M) / mu_inf after the H-side CPML and
 * the sources), comp 3..5.  Magnetic pole constants of an object: Obj::magAlpha() / magXi() / magGamma().  A soft source into H on a cell of
 * upLorB_ is overwritten by B2H as in the reference.  Runs on several slabs too (M and B are cell-local; the H row pushed upward is the one B2H
 * has written).  Magnetic oriented dipoles are refused. */
int chiml_gpu_set_magnetic(ChimlCtx* ctx, int has_B, int pml_on_B);
int chiml_gpu_set_object_magnetic(ChimlCtx* ctx, int obj, int npoles, const double* alpha, const double* xi, const double* gamma);
/* magnetic pole state lorM_[c][p] / prevLorM_[c][p] (c = 0..2 for Hx..Hz) expanded to the logical full grid */
int chiml_gpu_download_mag_pole(ChimlCtx* ctx, int comp, int pole, int prev, double* host);

/* Chiral media (UpdateChiral, UTIL/FDTD_up_eq.cpp:64-111; chiDtoU :920-925).  Per chiral pole p of an object, on the cells of CHIML_LIST_CHID:
 *   E side:  chiP_p = chiAlpha_p chiP_p + chiXi_p chiP_p,prev + (chiGamma_p / 8) sum_8 H_i + (chiGammaPrev_p / 8) sum_8 H_i,prev ;  E_i += (+1/eps) chiP_p
 *   H side:  chiM_p likewise from E_i and its previous value, before the H / B update of the step;                                H_i += (-1/mu) chiM_p
 * where the eight points are r, ind_j, ind_k, ind_j+ind_k-r, ind_i, ind_i+ind_j-r, ind_i+ind_k-r, ind_i+ind_j+ind_k-2r of the list entry, and "previous"
 * is the copy of the other family's fields taken right after the update that used them (copy2PrevFields_, parallelFDTDField.cpp:391-410: rows
 * {length, x, y, z} of a box around every chiral object).  Constants: Obj::chiAlpha() / chiXi() / chiGamma() / chiGammaPrev().  Needs D and B grids
 * (has_D, chiml_gpu_set_magnetic).  3-D grids, single slab; chiral oriented dipoles are refused. */
int chiml_gpu_set_object_chiral(ChimlCtx* ctx, int obj, int npoles, const double* alpha, const double* xi, const double* gamma, const double* gamma_prev);
int chiml_gpu_set_prev_copy(ChimlCtx* ctx, const int32_t* rows /* 4 per row: length, x, y, z */, size_t nrows);
/* chiral pole state lorChiHP_[c][p] (comp 0..2) / lorChiEM_[c][p] (comp 3..5) and their previous values; the previous-field copies prevE_ / prevH_ */
int chiml_gpu_download_chi_pole(ChimlCtx* ctx, int comp, int pole, int prev, double* host);
int chiml_gpu_download_prev_field(ChimlCtx* ctx, int comp, double* host);

/* Complex fields (Bloch-periodic runs: a k-point switches the reference to parallelFDTDFieldCplx, INPUTS/parallelInputs.cpp:108-112).  Every operator
 * of the step has real coefficients -- the reference's complex BLAS chains multiply by real factors -- so the real and the imaginary parts of all
 * arrays evolve as two real propagators over the SAME lists, coupled only by the phase factors exp(i k.L) of the periodic wrap copies
 * (applyBC1Proc, complex fields, UTIL/FDTD_up_eq.cpp:1248-1324) and driven by the real / imaginary parts of the pulse.  Two contexts set up
 * identically (lists, objects, CPML, sources, detectors, chiml_gpu_set_periodic on both) and committed are bound into a pair; from then on
 * chiml_gpu_step_n_cplx on the REAL part steps both (src_amp_re / src_amp_im: dt * Re / Im(sum pulse(t)) per step and source) and applies the Bloch
 * wrap copies; state is read from either context as usual.  Single slab, no emitters, running-DFT sets or TFSF surfaces. */
int chiml_gpu_bind_imag(ChimlCtx* re, ChimlCtx* im, const double* k_point /* 3 */);
int chiml_gpu_step_n_cplx(ChimlCtx* re, int n, const double* src_amp_re, const double* src_amp_im);

/* Total-field / scattered-field plane-wave source (SOURCE/parallelTFSF.hpp).  The 1-D auxiliary incident line (parallelTFSFBase::step,
 * :1148-1177: six complex 1-D fields with their own dispersion and CPML) does not depend on the main grid; it stays on the host --
 * the reference's own object steps it -- and the device applies the surface corrections of updateFields() (:1058-1073).  One surface
 * (paramStoreTFSF, :100-111, built by genSurface, :823-998) of target component comp (0..5 = Ex..Hz): for every pair
 * (incd index, main index) of indsD_ / indsU_ and i < n = szTrans_[0]
 *     target[ind_main + i * stride_main] += prefactor * Re(incd[ind_incd + i * stride_incd])                  (addIncdFields, parallelTFSF.cpp:77-83)
 *     ... += prefactor * (Re(incd[..]) / ep_mu[ind_incd + i * stride_incd])   for the U pairs when ep_mu != NULL  (addIncdFieldsEPChange, :93-105)
 * target = D_[comp] for the D pairs and E_/H_[comp] for the U pairs.  incd_offset selects the incident line inside the per-step table
 * handed to chiml_gpu_step_n_tfsf.  E / D surfaces are applied before the E half step and before the soft sources, H surfaces after the
 * H half step (step() item 5 sits between updateH and updateHPML_: a surface cell inside the CPML is refused).  Single slab only. */
typedef struct ChimlTfsfSurface
{
    int32_t comp;             /* target component 0..5 */
    int32_t incd_offset;      /* start of the surface's incident line inside one step's table */
    int32_t incd_len;         /* length of that line (bounds check) */
    int32_t n;                /* szTrans_[0] */
    int32_t stride_incd;      /* strideIncd_ (may be negative) */
    int32_t stride_main;      /* strideMain_ */
    int32_t npairs_D, npairs_U;
    double  prefactor;
    const int32_t* pairs_D;   /* 2 * npairs_D: indsD_ */
    const int32_t* pairs_U;   /* 2 * npairs_U: indsU_ */
    const double*  ep_mu;     /* incd_len doubles (eps_[c] / mu_[c] along the line) or NULL */
} ChimlTfsfSurface;
int chiml_gpu_add_tfsf_surface(ChimlCtx* ctx, const ChimlTfsfSurface* s);

/* Number of oriented-dipole pole grids of the WHOLE grid, orDipLorP_[c].size() = the largest pole count of any oriented-dipole
 * object (parallelFDTDField.hpp:452-478): every rank of the reference allocates and exchanges that many, whether or not its own slab
 * holds such an object.  Needed with several slabs only -- a slab that holds no oriented-dipole cell, or objects with fewer poles
 * than its neighbour's, still exchanges that many node P_y ghost rows.  0 (default) = this slab's own list decides. */
int chiml_gpu_set_ordip_pole_count(ChimlCtx* ctx, int n_poles_global);

/* Position-dependent dipole orientations (MAT_DIP_ORIENTAITON::REL_TO_NORM: "normal", "tangent", polar / azimuthal angles relative to the surface
 * normal, tanIso).  setupDipMoments (parallelFDTDField.hpp:960-1048) evaluates the object's surface gradient at every node; the result is the static
 * grid dipP_[comp][pole] that UpdateLorPolOrDip* (UTIL/FDTD_up_eq.cpp:450-631) multiplies with, node by node.  grid = that array, the whole local
 * ghost-inclusive grid (x fastest, then z, then y -- &dipP_[comp][pole]->point(0)); the engine keeps the values at the cells of CHIML_LIST_ORDIPP
 * only.  Where a grid is given it replaces the per-object direction of chiml_gpu_set_object for this (comp, pole) on every node; give it for every
 * component and pole the reference holds as soon as one pole of one object is oriented this way.  Before commit; the grid is read AT commit and
 * not copied (the reference's grids live as long as its propagator; six of them on a C5-sized slab are 25 GB): keep it valid until then. */
int chiml_gpu_set_dip_grid(ChimlCtx* ctx, int comp, int pole, const double* grid);

/* Column length of the y-marching kernels: how many stacked y planes of equal content one thread block walks, carrying the y-coupled
 * neighbour planes in registers (k_fast / k_uniform).  0 = automatic (from the grid size, up to 64 / 32).  Results do not depend on
 * it; it exists for tuning and so that the parity tests can force long columns on small grids (the environment variable
 * CHIML_B200_MARCH_NY="fast[,uniform]" does the same for contexts that never call this). */
int chiml_gpu_set_march(ChimlCtx* ctx, int fast_planes, int uniform_planes);

/* 2-D grids (single slab, no emitters) run all n steps of a chiml_gpu_step_n call in ONE cooperative launch with grid-wide barriers
 * between the phases of a step (csrc/chiml_persist.cuh) instead of 6-8 launches per step; results are identical.  Without this call the
 * engine chooses by grid size (one launch up to 1.5 M grid points: beyond that the launches cost less than the parallelism the resident
 * grid gives up); on = 1 forces the one-launch kernel, on = 0 the launch-per-phase path (also: environment variable
 * CHIML_B200_NO_PERSIST).  May be called at any time. */
int chiml_gpu_set_persistent(ChimlCtx* ctx, int on);

/* Freeze the setup: paints the per-cell update maps from the lists, builds the CPML coefficient
 * tables and compact psi / polarisation pools, zeroes all state. */
int chiml_gpu_commit(ChimlCtx* ctx);

/* ---- y-slabs (one context per GPU, one process per context) ---------------------------------------------------------------
 * Replaces parallelGrid::transferDat (GRID/parallelGrid.hpp:738-770) and the E / P box transfers of the emitter code
 * (ML/parallelQE.hpp:618-645,694-715): after commit every slab exports an opaque blob (CUDA IPC handles of the rows its neighbours
 * write, plus layout); the host transports the blobs (MPI, torch.distributed, a file ...) and hands each slab the blobs of the slab
 * below (rank-1) and above (rank+1), NULL where there is none.  From then on chiml_gpu_step_n exchanges ghost rows by peer-to-peer
 * stores over NVLink, boundary rows first, overlapped with the interior update.  All slabs must step in lock-step (same n). */
int chiml_gpu_halo_export(ChimlCtx* ctx, void* blob, size_t cap, size_t* size);   /* blob == NULL: only reports the size */
int chiml_gpu_halo_bind(ChimlCtx* ctx, const void* lower_blob, size_t lower_size, const void* upper_blob, size_t upper_size);

/* ---- stepping ---------------------------------------------------------------------------------- */
/* n leap-frog steps in the reference's order (step(), :1228-1303).  src_amp: n * n_sources doubles,
 * step-major (may be NULL when there are no sources). */
int chiml_gpu_step_n(ChimlCtx* ctx, int n, const double* src_amp);
/* same with running-DFT sets: twiddles = for every step k < n, for every group g in order, nfreq_g complex (re, im) numbers
 * exp(-i freq t_k), t_k = time after step k (only read on the steps the group samples).  chiml_gpu_step_n fails when DFT sets exist. */
int chiml_gpu_step_n_dft(ChimlCtx* ctx, int n, const double* src_amp, const double* twiddles);
/* same with TFSF surfaces: incd = for every step k < n a table of incd_per_step doubles holding the real parts of the incident lines the
 * surfaces of that step read -- for an H surface the incident E line BEFORE the line's step k, for an E / D surface the incident H line
 * AFTER it (updateFields: H surfaces, step(), E surfaces).  twiddles may be NULL without running-DFT sets. */
int chiml_gpu_step_n_tfsf(ChimlCtx* ctx, int n, const double* src_amp, const double* twiddles, const double* incd, size_t incd_per_step);
int chiml_gpu_sync(ChimlCtx* ctx);
/* same as step_n / step_n_dft (twiddles may be NULL when no running-DFT set is registered) but bracketed by CUDA events on the
 * context's stream; returns device milliseconds */
int chiml_gpu_step_n_timed(ChimlCtx* ctx, int n, const double* src_amp, const double* twiddles, float* ms);
/* Detector and population samples go to device rings.  chiml_gpu_step_n makes room for the samples of its n steps BEFORE it launches
 * anything (a ring whose retained samples leave no room is re-allocated there, once, never inside the step loop); a host that drains
 * the rings -- read, then consume -- between calls keeps them at their initial size for runs of any length.  reserve_steps does the
 * same sizing ahead of time for the next n_steps steps (e.g. for the whole run, when nothing is read before the end). */
int chiml_gpu_reserve_steps(ChimlCtx* ctx, long long n_steps);
/* number of kernels this context has launched since creation */
int64_t chiml_gpu_launch_count(const ChimlCtx* ctx);

/* ---- accounting --------------------------------------------------------------------------------- */
/* Per-kernel statistics of the step loop.  With kernel timing on, every launch made by chiml_gpu_step_n is bracketed by
 * CUDA events on the context's stream; chiml_gpu_kernel_stat synchronises and returns, for kernel `kind`
 * (0 <= kind < chiml_gpu_n_kernel_kinds()), the number of launches, the summed device time of the timed ones and the
 * ALGORITHMIC bytes one launch moves (time-varying state only, counted from the painted cells). */
typedef struct ChimlKernelStat
{
    char    name[32];
    int64_t launches;
    int64_t timed_launches;
    double  ms_total;
    double  alg_bytes_per_launch;  /* alg_bytes_per_step / launches per step (a slab with neighbours launches every tile list twice: the
                                      slab-boundary rows first, then the interior) */
    double  alg_bytes_per_step;    /* bytes all launches of this kernel move in ONE time step */
} ChimlKernelStat;
int chiml_gpu_set_kernel_timing(ChimlCtx* ctx, int on);
int chiml_gpu_n_kernel_kinds(void);
int chiml_gpu_kernel_stat(ChimlCtx* ctx, int kind, ChimlKernelStat* out);
int chiml_gpu_reset_kernel_stats(ChimlCtx* ctx);

/* ---- state access (synchronous) ---------------------------------------------------------------- */
/* host buffers use the reference's logical layout, ln[0]*ln[1]*ln[2] doubles, ghosts included */
int chiml_gpu_upload_field(ChimlCtx* ctx, int field, const double* host);
int chiml_gpu_download_field(ChimlCtx* ctx, int field, double* host);
/* isotropic pole state lorP_[c][p] / prevLorP_[c][p] expanded to the logical full grid (zeros elsewhere) */
int chiml_gpu_download_pole(ChimlCtx* ctx, int comp, int pole, int prev, double* host);
int chiml_gpu_upload_pole(ChimlCtx* ctx, int comp, int pole, int prev, const double* host);
/* oriented-dipole pole state orDipLorP_[c][p] / prevOrDipLorP_[c][p] (node-centred), expanded likewise */
int chiml_gpu_download_ordip_pole(ChimlCtx* ctx, int comp, int pole, int prev, double* host);
/* CPML psi of component comp, part 0/1, expanded to the logical full grid */
int chiml_gpu_download_psi(ChimlCtx* ctx, int comp, int part, double* host);
/* copies up to cap samples (each sz[0]*sz[1]*sz[2] doubles, x fastest then z then y), oldest retained sample first, and reports how
 * many are retained (= all samples since commit when nothing was consumed; sample 0 is the one taken at t = 0) */
int chiml_gpu_read_detector(ChimlCtx* ctx, int slot, double* out, size_t cap_samples, size_t* n_samples);
/* samples [first, first+n) by absolute number (what a host loop that drains the detector after every step reads); *n_read = how
 * many of them are retained */
int chiml_gpu_read_detector_range(ChimlCtx* ctx, int slot, size_t first, size_t n, double* out, size_t* n_read);
/* releases every sample with absolute number < upto: its ring slots are reused by later samples */
int chiml_gpu_consume_detector(ChimlCtx* ctx, int slot, size_t upto);

/* emitter state of level system `sys`: which = 0 rho, 1..4 = d rho/dt at n, n-1, n-2, n-3; out = nemit * N*N complex, emitter-major */
int chiml_gpu_download_emitter_state(ChimlCtx* ctx, int slot, int sys, int which, double* out);
/* the emitter polarisation box P_[comp] (shape as ChimlEmitterDesc::eps) */
int chiml_gpu_download_emitter_pol(ChimlCtx* ctx, int slot, int comp, double* out);
/* population detector `det` of emitter set `slot`: complex samples sum_emitters rho[level] / npoints of THIS slab's emitters
 * (QEPopDtc::accumPop; the host adds the slabs, QEPopDtc::toFile).  out = cap_samples complex. */
int chiml_gpu_read_population(ChimlCtx* ctx, int slot, int det, double* out, size_t cap_samples, size_t* n_samples);
/* releases the samples with absolute number < upto of EVERY population detector of emitter set `slot` (they share one ring) */
int chiml_gpu_consume_population(ChimlCtx* ctx, int slot, size_t upto);

/* accumulators of DFT set `slot`: acc_len doubles each (fInReal_, fInCplx_) */
int chiml_gpu_download_dft(ChimlCtx* ctx, int slot, double* re, double* im);

/* bytes of device memory held by the context */
size_t chiml_gpu_device_bytes(const ChimlCtx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* CHIML_GPU_H */
