/* chiml_plan.h -- on-disk form of everything the C ABI in chiml_gpu.h consumes ("plan file").
 *
 * A plan is the flattened output of the propagator constructor for ONE y-slab (rank): grid
 * description, update lists, object pole constants, CPML lists, source boxes with their per-step
 * amplitudes, detector boxes.  It is written by the host side of this repository
 * (chiml_b200/host, `chiml --dump-plan`) and -- for parity tests -- by oracle/ref_driver.cpp from the
 * internals of the unmodified reference, so the two can be compared list by list.
 *
 * File = sequence of records, little endian:  char tag[8] | uint64 nbytes | payload[nbytes]
 * The first record is CHIMLPLN (payload: int32 version).  Unknown tags are skipped by readers.
 */
#ifndef CHIML_PLAN_H
#define CHIML_PLAN_H

#include "chiml_gpu.h"

#define CHIML_PLAN_VERSION 1

#pragma pack(push, 1)
typedef struct ChimlPlanGrid          /* tag "GRID    " */
{
    ChimlGridDesc desc;
    int32_t y_start;      /* parallelGrid::procLoc(1): global row of local row 1 */
    int32_t n_global[3];  /* n_vec_ (points per direction, no ghosts) */
    int32_t n_steps;      /* ceil(tMax/dt) (main.cpp:54) */
    int32_t n_lor_poles;  /* max over components of lorP_[c].size() */
    int32_t n_ordip_poles;/* max over components of orDipLorP_[c].size() */
    int32_t pad;
    double  t_max;
} ChimlPlanGrid;

typedef struct ChimlPlanListHdr       /* tag "UPLIST  ": header + n * ChimlRun */
{
    int32_t kind, comp;
    uint64_t n;
} ChimlPlanListHdr;

typedef struct ChimlPlanObjectHdr     /* tag "OBJECT  ": header + alpha[np] xi[np] gamma[np] dip[3 np] */
{
    int32_t obj, npoles, use_or_dip, ml;
    double  eps_inf, mu_inf;
} ChimlPlanObjectHdr;

typedef struct ChimlPlanCpmlHdr       /* tag "CPML    ": header + npsi * ChimlPsiParams + ngrid * ChimlGridParams */
{
    int32_t comp, part, has_psi, pad;
    uint64_t npsi, ngrid;
} ChimlPlanCpmlHdr;

typedef struct ChimlPlanSourceHdr     /* tag "SOURCE  ": header + n_steps amplitudes (dt * Re sum pulse(t_k)) */
{
    int32_t field;
    int32_t loc[3];
    int32_t sz[3];
    int32_t n_steps;
} ChimlPlanSourceHdr;

typedef struct ChimlPlanDetector      /* tag "DETECTOR": one stored field box of a time-domain detector */
{
    int32_t detector;     /* index into dtcArr_ */
    int32_t field;
    int32_t loc[3];       /* GLOBAL grid coordinates (no ghosts) of the stored box */
    int32_t sz[3];        /* already grown by the Yee offset (parallelStorageDTC.hpp:62-63) */
    int32_t offset[3];
    int32_t every;        /* timeInterval_ in steps */
    int32_t type;         /* DTCTYPE as int */
    int32_t pad;
    double  conv;         /* convFactor_ */
    double  t_conv;       /* tConv_ */
} ChimlPlanDetector;
typedef struct ChimlPlanEmitterHdr    /* tag "EMITTER ": header, then in this order: h0[nsys*N*N*2] weight[nsys] mu[3*N*N*2]
                                         gam_ptr[N*N+1] gam_col[nnz] gam_val[nnz] loc[3*nemit] eps[(n0+2)(n1+2)pz] pop_level[npop] */
{
    int32_t object;       /* index into qeArr_ */
    int32_t nlevel, nsys, nemit;
    int32_t box_lo[3];
    int32_t box_n[3];
    int32_t nnz, npop, pop_every, npoints;
    int32_t pz, pad;
    double  dt, inv_hbar, na;
} ChimlPlanEmitterHdr;
typedef struct ChimlPlanDftHdr        /* tag "DFT     ": header, freq[nfreq] doubles (group's freqList_), then nlines ChimlDftLine */
{
    int32_t field, group, every, nfreq, npts, stride;
    uint64_t nlines, acc_len;
} ChimlPlanDftHdr;
typedef struct ChimlPlanPeriodic      /* tag "PERIODIC": the wrap copies of one component (chiml_gpu_set_periodic) */
{
    int32_t comp;
    ChimlWrap wrap;
} ChimlPlanPeriodic;
typedef struct ChimlPlanTfsfSurfaceHdr /* tag "TFSFSURF": header, pairs_D[2 npairs_D] pairs_U[2 npairs_U] int32, then ep_mu[incd_len] doubles if has_ep_mu */
{
    int32_t comp, incd_offset, incd_len, n, stride_incd, stride_main, npairs_D, npairs_U, has_ep_mu, pad;
    double  prefactor;
} ChimlPlanTfsfSurfaceHdr;
typedef struct ChimlPlanTfsfLinesHdr   /* tag "TFSFLINE": header, then n_steps * per_step doubles (chiml_gpu_step_n_tfsf's table); written AFTER the
                                          run by the reference driver (the line is stepped by the reference's own object) */
{
    int32_t n_steps, per_step;
} ChimlPlanTfsfLinesHdr;
typedef struct ChimlPlanMagnetic       /* tag "MAGNETIC": B_ grids exist / the H-side CPML acts on B (chiml_gpu_set_magnetic) */
{
    int32_t has_B, pml_on_B, n_mag_poles /* max over components of lorM_[c].size() */, pad;
} ChimlPlanMagnetic;
typedef struct ChimlPlanObjMagHdr      /* tag "OBJMAG  ": header + magAlpha[np] magXi[np] magGamma[np] of object obj */
{
    int32_t obj, npoles;
} ChimlPlanObjMagHdr;
typedef struct ChimlPlanObjChiHdr      /* tag "OBJCHI  ": header + chiAlpha[np] chiXi[np] chiGamma[np] chiGammaPrev[np] of object obj */
{
    int32_t obj, npoles;
} ChimlPlanObjChiHdr;
/* tag "DIPGRID ": int32 comp, int32 pole, then ln[0]*ln[1]*ln[2] doubles: dipP_[comp][pole] of setupDipMoments (written only when a pole is oriented
 * relative to the surface normal, REL_TO_NORM; then for every grid the reference holds) */
/* tag "PREVCOPY": uint64 nrows, then nrows x {int32 length, x, y, z}: copy2PrevFields_ (local ghost-inclusive coordinates) */
typedef struct ChimlPlanComplex        /* tag "COMPLEX ": the propagator holds complex fields (Bloch-periodic run, parallelFDTDFieldCplx): every field / psi /
                                          pole array has a real and an imaginary part, coupled only by the phase factors of the periodic wrap copies */
{
    int32_t cplx, pad;
    double  k_point[3];                /* k_point_ */
} ChimlPlanComplex;
/* tag "SRCIMAG ": int32 n_steps, then n_steps doubles dt * Im(sum pulse(t_k)) of the SOURCE record before it (complex fields only) */
#pragma pack(pop)

#endif /* CHIML_PLAN_H */
