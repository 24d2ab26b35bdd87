/* TEST INFRASTRUCTURE ONLY -- see chiml_oracle.h.  CPU restatement of the reference hot path.
 * Build: gcc -std=c99 -O2 -ffp-contract=off (oracle/Makefile).  All `file:line` citations are
 * relative to the reference tree (/root/reference/src). */
#define _POSIX_C_SOURCE 200809L
#include "chiml_oracle.h"

#include <complex.h>
#undef I
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define MAX_POLES 16
#define MAX_SRC 64

typedef struct { ChimlRun* r; size_t n; } RunList;
typedef struct { int has_psi, present; ChimlPsiParams* psi; size_t npsi; ChimlGridParams* grid; size_t ngrid; double* psi_grid; } CpmlPart;
typedef struct { int npoles, use_or_dip; double alpha[MAX_POLES], xi[MAX_POLES], gamma[MAX_POLES], dip[MAX_POLES][3]; } ObjConst;
typedef struct { int field; int32_t loc[3], sz[3]; } SrcBox;

/* one parallelQE object (ML/parallelQE.hpp): owned copies of the ChimlEmitterDesc arrays + state */
#define MAX_QE 8
#define MAX_N2 64
typedef struct
{
    ChimlEmitterDesc d;
    int n2, pz;
    size_t pbox;
    double *h0, *weight, *mu, *gam_val, *eps;
    int32_t *gam_ptr, *gam_col, *loc, *pop_level;
    double* P[3];          /* P_[c] boxes */
    double* st[5];         /* [w][sys][emitter][n2][re,im]: density_, density_deriv_n_, _n_minus_1_, _2_, _3_ (ML/density.hpp:24-28) */
    double* pop;           /* [det][sample][re,im] */
    size_t pop_cap, pop_n;
    double* curpop;        /* [det][re,im] */
    long tstep;
} QESet;

/* one stored field of a flux / frequency detector (DTC/parallelStorageFreqDTC.cpp:21-30) */
#define MAX_DFT 512
typedef struct { int field, group, every, nfreq, npts, stride; size_t nlines, acc_len; ChimlDftLine* lines; double *re, *im; } DftSet;

/* one TFSF surface (SOURCE/parallelTFSF.hpp:100-111) */
#define MAX_TFSF 256
typedef struct { ChimlTfsfSurface s; int32_t *pairs_D, *pairs_U; double* ep_mu; } TfsfSur;

struct OracleSim
{
    ChimlGridDesc g;
    size_t ncell;
    double* f[CHIML_NFIELDS];
    RunList up[6][6];              /* [kind][comp] */
    CpmlPart pml[6][2];
    ObjConst* obj;
    int npoles;                    /* number of allocated isotropic pole grids = max over objects (lorP_[c].size()) */
    int nordip;                    /* number of allocated oriented-dipole pole grids */
    double* P[3][MAX_POLES];       /* lorP_[c][p] */
    double* Pprev[3][MAX_POLES];   /* prevLorP_[c][p] */
    double* oP[3][MAX_POLES];      /* orDipLorP_[c][p] */
    double* oPprev[3][MAX_POLES];  /* prevOrDipLorP_[c][p] */
    double* dipgrid[3][MAX_POLES]; /* dipP_[c][p] (static) */
    double* dipgiven[3][MAX_POLES];/* dipP_[c][p] as handed over by oracle_set_dip_grid (position-dependent orientations), or NULL */
    int has_B, pml_on_B, nmag;     /* magnetic-dispersive media: B grids exist, the H-side CPML acts on B, number of lorM_ grids */
    double* M[3][MAX_POLES];       /* lorM_[c][p] */
    double* Mprev[3][MAX_POLES];   /* prevLorM_[c][p] */
    ObjConst* mobj;                /* magnetic pole constants per object (magAlpha, magXi, magGamma) */
    /* chiral media */
    ObjConst* cobj;                /* chiral constants per object: alpha = chiAlpha, xi = chiXi, gamma = chiGamma, dip[p][0] = chiGammaPrev */
    int nchi;                      /* number of chiral pole grids = max over objects */
    double* chi[6][MAX_POLES];     /* lorChiHP_[c][p] (0..2), lorChiEM_[c][p] (3..5) */
    double* chiprev[6][MAX_POLES];
    double* prevf[6];              /* prevE_[0..2], prevH_[0..2] */
    int32_t* prevcopy; size_t nprevcopy;   /* copy2PrevFields_ rows {length, x, y, z} */
    SrcBox src[MAX_SRC];
    int nsrc;
    int committed;
    /* threading */
    int nthreads;
    pthread_barrier_t bar;
    const double* src_amp;
    int nsteps;
    QESet qe[MAX_QE];
    int nqe;
    DftSet dft[MAX_DFT];
    int ndft;
    int dft_group_nfreq[256];
    const double* twiddles;        /* of the current oracle_step_n_dft call */
    long step_count;
    int phase_mask;                /* bit 0: H half step + sources, 1: node poles, 2: E half step + emitter addP, 3: emitter density */
    TfsfSur tfsf[MAX_TFSF];        /* TFSF surfaces (paramStoreTFSF) */
    int ntfsf;
    const double* tfsf_incd;       /* of the current oracle_step_n_tfsf call: [step][per_step] */
    size_t tfsf_per_step;
    ChimlWrap wrap[6];             /* periodic wrap copies per component (applBCE_ / applBCH_ arguments) */
    int has_wrap[6];
};

static int comp_exists(const OracleSim* s, int field)
{
    int c = field % 3, isH = (field >= 3 && field < 6);
    if(field >= CHIML_BX) return 0;                 /* B grids are made by oracle_set_magnetic */
    if(field >= 6 && !s->g.has_D) return 0;
    if(s->g.mode == CHIML_MODE_3D) return 1;
    if(s->g.mode == CHIML_MODE_TE) return isH ? (c == 2) : (c != 2);   /* Ex,Ey,Hz (parallelFDTDField.hpp:391-396) */
    return isH ? (c != 2) : (c == 2);                                    /* Ez,Hx,Hy (:418-423) */
}

OracleSim* oracle_create(const ChimlGridDesc* desc)
{
    OracleSim* s = (OracleSim*)calloc(1, sizeof(OracleSim));
    if(!s) return NULL;
    s->g = *desc;
    s->ncell = (size_t)desc->ln[0] * (size_t)desc->ln[1] * (size_t)desc->ln[2];
    s->obj = (ObjConst*)calloc((size_t)(desc->n_objects > 0 ? desc->n_objects : 1), sizeof(ObjConst));
    for(int fld = 0; fld < CHIML_NFIELDS; ++fld)
        if(comp_exists(s, fld)) s->f[fld] = (double*)calloc(s->ncell, sizeof(double));
    return s;
}

void oracle_destroy(OracleSim* s)
{
    if(!s) return;
    for(int i = 0; i < CHIML_NFIELDS; ++i) free(s->f[i]);
    for(int k = 0; k < 6; ++k) for(int c = 0; c < 6; ++c) free(s->up[k][c].r);
    for(int c = 0; c < 6; ++c) for(int p = 0; p < 2; ++p) { free(s->pml[c][p].psi); free(s->pml[c][p].grid); free(s->pml[c][p].psi_grid); }
    for(int c = 0; c < 3; ++c) for(int p = 0; p < MAX_POLES; ++p) { free(s->P[c][p]); free(s->Pprev[c][p]); free(s->oP[c][p]); free(s->oPprev[c][p]); free(s->dipgrid[c][p]); free(s->dipgiven[c][p]); }
    free(s->obj);
    free(s);
}

int oracle_set_update_list(OracleSim* s, int kind, int comp, const ChimlRun* runs, size_t n)
{
    if(kind < 0 || kind > 5 || comp < 0 || comp > 5) return CHIML_ERR_ARG;
    RunList* l = &s->up[kind][comp];
    free(l->r);
    l->r = (ChimlRun*)malloc((n ? n : 1) * sizeof(ChimlRun));
    if(n) memcpy(l->r, runs, n * sizeof(ChimlRun));
    l->n = n;
    return 0;
}

int oracle_set_object(OracleSim* s, int obj, int npoles, const double* alpha, const double* xi, const double* gamma, int use_or_dip, const double* dip)
{
    if(obj < 0 || obj >= s->g.n_objects || npoles < 0 || npoles > MAX_POLES) return CHIML_ERR_ARG;
    ObjConst* o = &s->obj[obj];
    o->npoles = npoles;
    o->use_or_dip = use_or_dip;
    for(int p = 0; p < npoles; ++p)
    {
        o->alpha[p] = alpha[p]; o->xi[p] = xi[p]; o->gamma[p] = gamma[p];
        for(int k = 0; k < 3; ++k) o->dip[p][k] = dip ? dip[3 * p + k] : 0.0;
    }
    return 0;
}

/* position-dependent dipole grids (include/chiml_gpu.h chiml_gpu_set_dip_grid): dipP_[comp][pole] of setupDipMoments
 * (parallelFDTDField.hpp:960-1048) for REL_TO_NORM orientations, the whole ghost-inclusive grid as the reference holds it */
int oracle_set_dip_grid(OracleSim* s, int comp, int pole, const double* grid)
{
    if(!s || !grid || comp < 0 || comp > 2 || pole < 0 || pole >= MAX_POLES) return CHIML_ERR_ARG;
    if(s->committed) return CHIML_ERR_STATE;
    free(s->dipgiven[comp][pole]);
    s->dipgiven[comp][pole] = (double*)malloc(s->ncell * sizeof(double));
    memcpy(s->dipgiven[comp][pole], grid, s->ncell * sizeof(double));
    return 0;
}

/* chiral media (include/chiml_gpu.h chiml_gpu_set_object_chiral / chiml_gpu_set_prev_copy) */
int oracle_set_object_chiral(OracleSim* s, int obj, int npoles, const double* alpha, const double* xi, const double* gamma, const double* gamma_prev)
{
    if(!s || obj < 0 || obj >= s->g.n_objects || npoles < 0 || npoles > MAX_POLES) return CHIML_ERR_ARG;
    if(!s->cobj) s->cobj = (ObjConst*)calloc((size_t)(s->g.n_objects > 0 ? s->g.n_objects : 1), sizeof(ObjConst));
    ObjConst* o = &s->cobj[obj];
    o->npoles = npoles;
    for(int p = 0; p < npoles; ++p) { o->alpha[p] = alpha[p]; o->xi[p] = xi[p]; o->gamma[p] = gamma[p]; o->dip[p][0] = gamma_prev[p]; }
    return 0;
}
int oracle_set_prev_copy(OracleSim* s, const int32_t* rows, size_t nrows)
{
    if(!s || (nrows && !rows)) return CHIML_ERR_ARG;
    free(s->prevcopy);
    s->prevcopy = (int32_t*)malloc((nrows ? nrows : 1) * 4 * sizeof(int32_t));
    if(nrows) memcpy(s->prevcopy, rows, nrows * 4 * sizeof(int32_t));
    s->nprevcopy = nrows;
    return 0;
}
double* oracle_chi_pole(OracleSim* s, int comp, int pole, int prev) { return (comp < 0 || comp > 5 || pole < 0 || pole >= MAX_POLES) ? NULL : (prev ? s->chiprev[comp][pole] : s->chi[comp][pole]); }
double* oracle_prev_field(OracleSim* s, int comp) { return (comp < 0 || comp > 5) ? NULL : s->prevf[comp]; }

/* magnetic-dispersive media (include/chiml_gpu.h chiml_gpu_set_magnetic / chiml_gpu_set_object_magnetic) */
int oracle_set_magnetic(OracleSim* s, int has_B, int pml_on_B)
{
    if(!s || s->committed) return CHIML_ERR_STATE;
    s->has_B = has_B; s->pml_on_B = pml_on_B;
    if(has_B)
        for(int c = 0; c < 3; ++c)
            if(s->f[CHIML_HX + c] && !s->f[CHIML_BX + c]) s->f[CHIML_BX + c] = (double*)calloc(s->ncell, sizeof(double));
    if(!s->mobj) s->mobj = (ObjConst*)calloc((size_t)(s->g.n_objects > 0 ? s->g.n_objects : 1), sizeof(ObjConst));
    return 0;
}
int oracle_set_object_magnetic(OracleSim* s, int obj, int npoles, const double* alpha, const double* xi, const double* gamma)
{
    if(!s || obj < 0 || obj >= s->g.n_objects || npoles < 0 || npoles > MAX_POLES) return CHIML_ERR_ARG;
    if(!s->mobj) s->mobj = (ObjConst*)calloc((size_t)(s->g.n_objects > 0 ? s->g.n_objects : 1), sizeof(ObjConst));
    ObjConst* o = &s->mobj[obj];
    o->npoles = npoles;
    for(int p = 0; p < npoles; ++p) { o->alpha[p] = alpha[p]; o->xi[p] = xi[p]; o->gamma[p] = gamma[p]; }
    return 0;
}
double* oracle_mag_pole(OracleSim* s, int comp, int pole, int prev) { return (comp < 0 || comp > 2 || pole < 0 || pole >= MAX_POLES) ? NULL : (prev ? s->Mprev[comp][pole] : s->M[comp][pole]); }
int oracle_n_mag_poles(OracleSim* s) { return s->nmag; }

int oracle_set_cpml(OracleSim* s, int comp, int part, int has_psi, const ChimlPsiParams* psi, size_t npsi, const ChimlGridParams* grid, size_t ngrid)
{
    if(comp < 0 || comp > 5 || part < 0 || part > 1) return CHIML_ERR_ARG;
    CpmlPart* p = &s->pml[comp][part];
    free(p->psi); free(p->grid);
    p->psi = (ChimlPsiParams*)malloc((npsi ? npsi : 1) * sizeof(ChimlPsiParams));
    p->grid = (ChimlGridParams*)malloc((ngrid ? ngrid : 1) * sizeof(ChimlGridParams));
    if(npsi) memcpy(p->psi, psi, npsi * sizeof(ChimlPsiParams));
    if(ngrid) memcpy(p->grid, grid, ngrid * sizeof(ChimlGridParams));
    p->npsi = npsi; p->ngrid = ngrid; p->has_psi = has_psi; p->present = 1;
    return 0;
}

int oracle_add_source(OracleSim* s, int field, const int32_t loc[3], const int32_t sz[3])
{
    if(s->nsrc >= MAX_SRC || field < 0 || field >= CHIML_NFIELDS || !s->f[field]) return CHIML_ERR_ARG;
    SrcBox* b = &s->src[s->nsrc++];
    b->field = field;
    for(int k = 0; k < 3; ++k) { b->loc[k] = loc[k]; b->sz[k] = sz[k]; }
    return 0;
}

/* Grid allocation mirrors the constructor: one P / prevP grid per pole index up to the largest pole
 * count of any object (parallelFDTDField.hpp:448-479), oriented-dipole grids likewise (:498-545), and
 * full-size psi grids (PML/parallelPML.hpp:143-156).  dipP_ grids: setupDipMoments (:960-1023) for
 * ISOTROPIC / UNIDIRECTIONAL orientations, evaluated from the node-centred object map which is
 * recovered here from the oriented-dipole node list (cells outside it are never read). */
int oracle_commit(OracleSim* s)
{
    s->npoles = 0; s->nordip = 0;
    for(int o = 0; o < s->g.n_objects; ++o)
    {
        if(s->obj[o].npoles > s->npoles) s->npoles = s->obj[o].npoles;
        if(s->obj[o].use_or_dip && s->obj[o].npoles > s->nordip) s->nordip = s->obj[o].npoles;
    }
    for(int c = 0; c < 3; ++c)
    {
        if(!s->f[CHIML_DX + c]) continue;
        for(int p = 0; p < s->npoles; ++p)
        {
            s->P[c][p] = (double*)calloc(s->ncell, sizeof(double));
            s->Pprev[c][p] = (double*)calloc(s->ncell, sizeof(double));
        }
        for(int p = 0; p < s->nordip; ++p)
        {
            s->oP[c][p] = (double*)calloc(s->ncell, sizeof(double));
            s->oPprev[c][p] = (double*)calloc(s->ncell, sizeof(double));
            s->dipgrid[c][p] = (double*)calloc(s->ncell, sizeof(double));
        }
    }
    /* lorM_ / prevLorM_: as many grids per H component as the largest magnetic pole count of any object */
    s->nmag = 0;
    if(s->has_B && s->mobj)
        for(int o = 0; o < s->g.n_objects; ++o) if(s->mobj[o].npoles > s->nmag) s->nmag = s->mobj[o].npoles;
    for(int c = 0; c < 3 && s->has_B; ++c)
    {
        if(!s->f[CHIML_BX + c]) continue;
        for(int p = 0; p < s->nmag; ++p)
        {
            s->M[c][p] = (double*)calloc(s->ncell, sizeof(double));
            s->Mprev[c][p] = (double*)calloc(s->ncell, sizeof(double));
        }
    }
    /* chiral media: lorChiHP_ / lorChiEM_ and their previous values, prevE_ / prevH_ (3-D: every component) */
    s->nchi = 0;
    if(s->cobj)
        for(int o = 0; o < s->g.n_objects; ++o) if(s->cobj[o].npoles > s->nchi) s->nchi = s->cobj[o].npoles;
    if(s->nchi > 0)
    {
        if(s->g.mode != CHIML_MODE_3D || !s->g.has_D || !s->has_B) return CHIML_ERR_UNSUPPORTED;
        for(int c = 0; c < 6; ++c)
        {
            s->prevf[c] = (double*)calloc(s->ncell, sizeof(double));
            for(int p = 0; p < s->nchi; ++p)
            {
                s->chi[c][p] = (double*)calloc(s->ncell, sizeof(double));
                s->chiprev[c][p] = (double*)calloc(s->ncell, sizeof(double));
            }
        }
    }
    const RunList* nl = &s->up[CHIML_LIST_ORDIPP][0];
    for(size_t e = 0; e < nl->n; ++e)
    {
        const ChimlRun* r = &nl->r[e];
        const ObjConst* o = &s->obj[r->obj];
        for(int c = 0; c < 3; ++c)
            for(int p = 0; p < o->npoles && p < s->nordip; ++p)
                if(s->dipgrid[c][p])
                    for(int i = 0; i < r->n; ++i) s->dipgrid[c][p][r->ind + i] = o->dip[p][c];
    }
    for(int c = 0; c < 3; ++c)
        for(int p = 0; p < s->nordip; ++p)
            if(s->dipgiven[c][p] && s->dipgrid[c][p]) memcpy(s->dipgrid[c][p], s->dipgiven[c][p], s->ncell * sizeof(double));
    for(int c = 0; c < 6; ++c)
        for(int p = 0; p < 2; ++p)
            if(s->pml[c][p].present && s->pml[c][p].has_psi)
                s->pml[c][p].psi_grid = (double*)calloc(s->ncell, sizeof(double));
    s->committed = 1;
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * BLAS level-1 restated (reference semantics: UTIL/utilities_MKL.hpp wrappers over daxpy_/dscal_/dcopy_)
 * ------------------------------------------------------------------------------------------- */
static inline void axpy(int n, double a, const double* x, int incx, double* y, int incy)
{ for(int i = 0; i < n; ++i) y[(size_t)i * incy] = y[(size_t)i * incy] + a * x[(size_t)i * incx]; }
static inline void scal(int n, double a, double* x, int incx)
{ for(int i = 0; i < n; ++i) x[(size_t)i * incx] = a * x[(size_t)i * incx]; }

/* UTIL/FDTD_up_eq.cpp:10-35 (OneCompCurlJ, OneCompCurlK, TwoCompCurl): the variant is selected by which
 * neighbour grids exist, as the constructor does (FDTD_MANAGER/parallelFDTDField.cpp:94-105,263-274). */
static void curl_run(const ChimlRun* r, double* Ui, const double* Vj, const double* Vk)
{
    if(Vj)
    {
        axpy(r->n,        r->pf[2], Vj + r->ind,   1, Ui + r->ind, 1);
        axpy(r->n, -1.0 * r->pf[2], Vj + r->ind_k, 1, Ui + r->ind, 1);
    }
    if(Vk)
    {
        axpy(r->n, -1.0 * r->pf[1], Vk + r->ind,   1, Ui + r->ind, 1);
        axpy(r->n,        r->pf[1], Vk + r->ind_j, 1, Ui + r->ind, 1);
    }
}

/* UTIL/FDTD_up_eq.cpp:425-448 UpdateLorPol */
static void lor_pol_run(const ChimlRun* r, const double* Ei, double** P, double** Pprev, const ObjConst* o, double* jstore)
{
    for(int pp = 0; pp < o->npoles; ++pp)
    {
        memcpy(jstore, P[pp] + r->ind, (size_t)r->n * sizeof(double));
        scal(r->n, o->alpha[pp], P[pp] + r->ind, 1);
        axpy(r->n, o->xi[pp], Pprev[pp] + r->ind, 1, P[pp] + r->ind, 1);
        axpy(r->n, o->gamma[pp], Ei + r->ind, 1, P[pp] + r->ind, 1);
        memcpy(Pprev[pp] + r->ind, jstore, (size_t)r->n * sizeof(double));
    }
}

/* UTIL/FDTD_up_eq.cpp:450-631 UpdateLorPolOrDip / ...XY / ...Z; multAvg = x*y/2.0 (UTIL/utilityFxns.hpp:38).
 * have[c] tells which E components exist (3-D: all; TE: x,y; TM: z). */
static void lor_pol_ordip_run(const OracleSim* s, const ChimlRun* r, const ObjConst* o, double* scratch)
{
    const int n = r->n;
    double* dotU = scratch;
    double* dotF = scratch + n;
    const int has_x = s->f[CHIML_EX] != NULL, has_z = s->f[CHIML_EZ] != NULL;
    const int offs[3] = { r->ind_i, r->ind_j, r->ind_k };
    for(int pp = 0; pp < o->npoles; ++pp)
    {
        double* jst[3] = { scratch + 2 * n, scratch + 3 * n, scratch + 4 * n };
        for(int c = 0; c < 3; ++c)
            if(s->f[c]) memcpy(jst[c], s->oP[c][pp] + r->ind, (size_t)n * sizeof(double));
        for(int c = 0; c < 3; ++c)
        {
            if(!s->f[c]) continue;
            scal(n, o->alpha[pp], s->oP[c][pp] + r->ind, 1);
            axpy(n, o->xi[pp], s->oPprev[c][pp] + r->ind, 1, s->oP[c][pp] + r->ind, 1);
        }
        if(has_x)
        {
            /* 3-D (UpdateLorPolOrDip) or TE (UpdateLorPolOrDipXY): x and y (and z) contributions, in this order */
            int first = 1;
            for(int c = 0; c < 3; ++c)
            {
                if(!s->f[c]) continue;
                const double* dip = s->dipgrid[c][pp] + r->ind;
                const double* E0 = s->f[c] + r->ind;
                const double* E1 = s->f[c] + offs[c];
                if(first)
                {
                    for(int i = 0; i < n; ++i) dotU[i] = dip[i] * E0[i] / 2.0;
                    for(int i = 0; i < n; ++i) dotF[i] = dip[i] * E1[i] / 2.0;
                    axpy(n, 1.0, dotF, 1, dotU, 1);
                    first = 0;
                }
                else
                {
                    for(int i = 0; i < n; ++i) dotF[i] = dip[i] * E0[i] / 2.0;
                    axpy(n, 1.0, dotF, 1, dotU, 1);
                    for(int i = 0; i < n; ++i) dotF[i] = dip[i] * E1[i] / 2.0;
                    axpy(n, 1.0, dotF, 1, dotU, 1);
                }
            }
            for(int c = 0; c < 3; ++c)
            {
                if(!s->f[c]) continue;
                const double* dip = s->dipgrid[c][pp] + r->ind;
                for(int i = 0; i < n; ++i) dotF[i] = dip[i] * dotU[i];
                axpy(n, o->gamma[pp], dotF, 1, s->oP[c][pp] + r->ind, 1);
            }
        }
        else if(has_z)
        {
            /* TM (UpdateLorPolOrDipZ :606-631): dotU = dip_z * Ez, P_z += gamma * dotU */
            const double* dip = s->dipgrid[2][pp] + r->ind;
            const double* E0 = s->f[CHIML_EZ] + r->ind;
            for(int i = 0; i < n; ++i) dotU[i] = dip[i] * E0[i];
            axpy(n, o->gamma[pp], dotU, 1, s->oP[2][pp] + r->ind, 1);
        }
        for(int c = 0; c < 3; ++c)
            if(s->f[c]) memcpy(s->oPprev[c][pp] + r->ind, jst[c], (size_t)n * sizeof(double));
    }
}

/* UTIL/FDTD_up_eq.cpp:838-848 DtoU: sums over ALL allocated pole grids */
static void dtou_run(const ChimlRun* r, const double* Di, double* Ui, double** P, int nP)
{
    const double eps = r->pf[3];
    memcpy(Ui + r->ind, Di + r->ind, (size_t)r->n * sizeof(double));
    scal(r->n, 1.0 / eps, Ui + r->ind, 1);
    for(int pp = 0; pp < nP; ++pp)
        axpy(r->n, -1.0 / eps, P[pp] + r->ind, 1, Ui + r->ind, 1);
}

/* UpdateChiral (UTIL/FDTD_up_eq.cpp:64-111): chiral pole of component i driven by the other family's component i at the eight corners spanned
 * by the entry's three neighbour indices, current and previous values */
static void chiral_run(const ChimlRun* r, const double* O, const double* Oprev, double** C, double** Cprev, const ObjConst* o, double* jstore)
{
    const long i1 = r->ind, i2 = r->ind_i, i3 = r->ind_j, i4 = r->ind_k;
    const long pts[8] = {i1, i3, i4, i3 + i4 - i1, i2, i2 + i3 - i1, i2 + i4 - i1, i2 + i3 + i4 - 2 * i1};
    for(int pp = 0; pp < o->npoles; ++pp)
    {
        memcpy(jstore, C[pp] + i1, (size_t)r->n * sizeof(double));
        scal(r->n, o->alpha[pp], C[pp] + i1, 1);
        axpy(r->n, o->xi[pp], Cprev[pp] + i1, 1, C[pp] + i1, 1);
        for(int k = 0; k < 8; ++k) axpy(r->n, o->gamma[pp] / 8.0, O + pts[k], 1, C[pp] + i1, 1);
        for(int k = 0; k < 8; ++k) axpy(r->n, o->dip[pp][0] / 8.0, Oprev + pts[k], 1, C[pp] + i1, 1);
        memcpy(Cprev[pp] + i1, jstore, (size_t)r->n * sizeof(double));
    }
}
/* chiDtoU (:920-925) with epMuInfty as D2E / B2H pass it: -eps for E, +mu for H */
static void chi_dtou_run(const ChimlRun* r, double epMuInfty, double* Ui, double** C, int nC)
{
    for(int pp = 0; pp < nC; ++pp) axpy(r->n, -1.0 / epMuInfty, C[pp] + r->ind, 1, Ui + r->ind, 1);
}
/* copy2PrevFields_ (FDTD_MANAGER/parallelFDTDField.hpp:1411-1416, 1441-1446): the three components of one family into their prev grids */
static void prev_copy(OracleSim* s, int isE)
{
    const size_t lx = (size_t)s->g.ln[0], lz = (size_t)s->g.ln[2];
    for(size_t q = 0; q < s->nprevcopy; ++q)
    {
        const int32_t* b = s->prevcopy + 4 * q;
        const size_t off = (size_t)b[1] + lx * ((size_t)b[3] + lz * (size_t)b[2]);
        for(int i = 0; i < 3; ++i)
        {
            const int f = (isE ? CHIML_EX : CHIML_HX) + i, pf = (isE ? 0 : 3) + i;
            if(s->f[f] && s->prevf[pf]) memcpy(s->prevf[pf] + off, s->f[f] + off, (size_t)b[0] * sizeof(double));
        }
    }
}

/* UTIL/FDTD_up_eq.cpp:862-889 orDipDtoU (node->edge average) and orDipDtoUZ (2-D TM Ez) */
static void ordip_dtou_run(const ChimlRun* r, const double* Di, double* Ui, double** P, int nP, int zvariant)
{
    const double eps = r->pf[3];
    memcpy(Ui + r->ind, Di + r->ind, (size_t)r->n * sizeof(double));
    scal(r->n, 1.0 / eps, Ui + r->ind, 1);
    for(int pp = 0; pp < nP; ++pp)
    {
        if(zvariant)
            axpy(r->n, -1.0 / eps, P[pp] + r->ind, 1, Ui + r->ind, 1);
        else
        {
            axpy(r->n, -0.5 / eps, P[pp] + r->ind,   1, Ui + r->ind, 1);
            axpy(r->n, -0.5 / eps, P[pp] + r->ind_i, 1, Ui + r->ind, 1);
        }
    }
}

/* PML/parallelPML.cpp:32-40 updatePsiField */
static void psi_entry(const ChimlPsiParams* p, double* psi, const double* V)
{
    scal(p->transSz,        p->b, psi + p->ind, p->stride);
    axpy(p->transSz,        p->c, V + p->ind,    p->stride, psi + p->ind, p->stride);
    axpy(p->transSz, -1.0 * p->c, V + p->indOff, p->stride, psi + p->ind, p->stride);
}
/* PML/parallelPML.cpp:12-30 addPsi / addGridOnly (the grid part) */
static void pml_grid_entry(const ChimlGridParams* p, double* Ui, const double* psi, const double* V)
{
    axpy(p->nAx,        p->DbField, V + p->ind,    p->stride, Ui + p->ind, p->stride);
    axpy(p->nAx, -1.0 * p->DbField, V + p->indOff, p->stride, Ui + p->ind, p->stride);
    if(psi)
        axpy(p->nAx, p->Db, psi + p->ind, p->stride, Ui + p->ind, p->stride);
}

/* ---------------------------------------------------------------------------------------------
 * step(): FDTD_MANAGER/parallelFDTDField.hpp:1228-1303, restricted to the lists covered here
 * (no magnetic / chiral media, no TFSF).  Work inside each phase is split over threads by list
 * entry; entries of one list write disjoint cells, and phases are separated by barriers.
 * ------------------------------------------------------------------------------------------- */
typedef struct { OracleSim* s; int tid; } Worker;


/* ------------------------------------------------------------------------------------------------
 * Maxwell-Liouville emitters (ML/parallelQE.hpp, ML/Hamiltonian.cpp, ML/density.hpp, UTIL/FDTD_up_eq.cpp:1367-1473).
 * Complex arithmetic follows the BLAS the reference is linked against in oracle/_ref (oracle/ref_shim/blas_shim.cpp):
 * the conventional 4-multiply complex product, netlib loop order in zgemm.
 * ------------------------------------------------------------------------------------------------ */
typedef struct { double re, im; } cx;
static cx cmul(cx a, cx b) { cx r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r; }
static cx cadd(cx a, cx b) { cx r; r.re = a.re + b.re; r.im = a.im + b.im; return r; }

/* zaxpy_(n, cplx(a,0), x, 1, y, 1) */
static void zaxpy_real(int n, double a, const cx* x, cx* y)
{
    cx ca; ca.re = a; ca.im = 0.0;
    for(int k = 0; k < n; ++k) y[k] = cadd(y[k], cmul(ca, x[k]));
}

int oracle_add_emitters(OracleSim* s, const ChimlEmitterDesc* d)
{
    if(s->nqe >= MAX_QE || d->nlevel < 1 || d->nlevel * d->nlevel > MAX_N2) return CHIML_ERR_ARG;
    QESet* q = &s->qe[s->nqe++];
    memset(q, 0, sizeof(*q));
    q->d = *d;
    const int n2 = d->nlevel * d->nlevel;
    q->n2 = n2;
    q->pz = s->g.ln[2] > 1 ? d->box_n[2] + 2 : 2;                       /* P_ shape, ML/parallelQE.hpp:470 */
    q->pbox = (size_t)(d->box_n[0] + 2) * (size_t)(d->box_n[1] + 2) * (size_t)q->pz;
#define DUP(dst, src, n, T) do { const size_t cnt_ = (size_t)(n); dst = (T*)malloc((cnt_ > 0 ? cnt_ : 1) * sizeof(T)); if(cnt_ > 0) memcpy(dst, src, cnt_ * sizeof(T)); } while(0)
    DUP(q->h0, d->h0, (size_t)d->nsys * n2 * 2, double);
    DUP(q->weight, d->weight, (size_t)d->nsys, double);
    DUP(q->mu, d->mu, (size_t)3 * n2 * 2, double);
    DUP(q->gam_ptr, d->gam_ptr, (size_t)n2 + 1, int32_t);
    DUP(q->gam_col, d->gam_col, (size_t)d->gam_ptr[n2], int32_t);
    DUP(q->gam_val, d->gam_val, (size_t)d->gam_ptr[n2], double);
    DUP(q->loc, d->loc, (size_t)3 * d->nemit, int32_t);
    DUP(q->eps, d->eps, q->pbox, double);
    DUP(q->pop_level, d->pop_level, (size_t)d->npop, int32_t);
#undef DUP
    for(int c = 0; c < 3; ++c) q->P[c] = (double*)calloc(q->pbox, sizeof(double));
    const size_t per = (size_t)d->nsys * (size_t)d->nemit * (size_t)n2 * 2;
    for(int w = 0; w < 5; ++w) q->st[w] = (double*)calloc(per ? per : 1, sizeof(double));
    /* Density::initializeDensity(weight): rho_00 = weight of the level system (ML/density.hpp:57-60, parallelQE.hpp:428-429) */
    for(int sy = 0; sy < d->nsys; ++sy)
        for(int e = 0; e < d->nemit; ++e)
            q->st[0][(((size_t)sy * d->nemit + e) * n2) * 2] = d->weight[sy];
    q->curpop = (double*)calloc((size_t)(d->npop ? d->npop : 1) * 2, sizeof(double));
    q->pop_cap = 1024;
    q->pop = (double*)calloc((size_t)(d->npop ? d->npop : 1) * q->pop_cap * 2, sizeof(double));
    return 0;
}

/* Hamiltonian::getHam (ML/Hamiltonian.cpp:59-69): H = h0 + Ex (-mu_x) + Ey (-mu_y) + Ez (-mu_z); a direction whose dipole matrix
 * is identically zero is skipped (:42-55) */
static void qe_get_ham(const QESet* q, int sys, const double e[3], cx* H)
{
    const int n2 = q->n2;
    const cx* h0 = (const cx*)q->h0 + (size_t)sys * n2;
    for(int k = 0; k < n2; ++k) H[k] = h0[k];
    if(e[0] == 0.0 && e[1] == 0.0 && e[2] == 0.0) return;
    for(int c = 0; c < 3; ++c)
    {
        const cx* mu = (const cx*)q->mu + (size_t)c * n2;
        int allzero = 1;
        for(int k = 0; k < n2; ++k) if(mu[k].re != 0.0 || mu[k].im != 0.0) allzero = 0;
        if(allzero) continue;
        cx a; a.re = e[c]; a.im = 0.0;
        for(int k = 0; k < n2; ++k)
        {
            cx neg; neg.re = -mu[k].re; neg.im = -mu[k].im;      /* neg_?_expectation_ = expectation * (-1.0 * coupling) */
            H[k] = cadd(H[k], cmul(a, neg));
        }
    }
}

/* parallelQEBase::denDeriv, MKL branch (ML/parallelQE.hpp:727-744): T = zgemm(i/hbar, H, rho) (column-major call on row-major data),
 * out = T + T^H (mkl_zomatadd 'R','N','C'), then the sparse relaxation rows */
static void qe_den_deriv(const QESet* q, const cx* H, const cx* den, cx* out)
{
    const int N = q->d.nlevel, n2 = q->n2;
    cx T[MAX_N2];
    cx alpha; alpha.re = 0.0; alpha.im = q->d.inv_hbar;
    for(int j = 0; j < N; ++j)
    {
        for(int i = 0; i < N; ++i) { T[i + j * N].re = 0.0; T[i + j * N].im = 0.0; }
        for(int l = 0; l < N; ++l)
        {
            const cx temp = cmul(alpha, den[l + j * N]);
            for(int i = 0; i < N; ++i) T[i + j * N] = cadd(T[i + j * N], cmul(temp, H[i + l * N]));
        }
    }
    cx one; one.re = 1.0; one.im = 0.0;
    for(int i = 0; i < N; ++i)
        for(int j = 0; j < N; ++j)
        {
            cx b; b.re = T[j * N + i].re; b.im = -T[j * N + i].im;
            out[i * N + j] = cadd(cmul(one, T[i * N + j]), cmul(one, b));
        }
    for(int ii = 0; ii < n2; ++ii)
        for(int k = q->gam_ptr[ii]; k < q->gam_ptr[ii + 1]; ++k)
        {
            const cx v = den[q->gam_col[k]];
            out[ii].re = out[ii].re + v.re * q->gam_val[k];
            out[ii].im = out[ii].im + v.im * q->gam_val[k];
        }
}

/* parallelQEBase::addQE (:682-718) + updateDensity (:614-678) on the slab that owns the nodes */
/* part: 1 = addP only, 2 = density update only, 3 = both (addQE) */
static void qe_add(OracleSim* s, QESet* q, int part)
{
    const ChimlEmitterDesc* d = &q->d;
    const int lnx = s->g.ln[0], lnz = s->g.ln[2];
    const int threeD = lnz > 1;
    const int zOff = threeD ? 1 : 0;
    const int bx = d->box_n[0] + 2, bz = q->pz;
    const int n2 = q->n2;
#define PB(i, j, k) ((size_t)(i) + (size_t)bx * ((size_t)(k) + (size_t)bz * (size_t)(j)))
#define GI(x, y, z) ((size_t)(x) + (size_t)lnx * ((size_t)(z) + (size_t)lnz * (size_t)(y)))
    /* addP (UTIL/FDTD_up_eq.cpp:1367-1380): E += -0.5 P[n]/eps[n], then E += -0.5 P[n+off]/eps[n+off] */
    for(int c = 0; c < 3 && (part & 1); ++c)
    {
        double* E = s->f[CHIML_EX + c];
        if(!E) continue;
        const int off[3] = { c == 0 ? 1 : 0, c == 1 ? 1 : 0, c == 2 ? zOff : 0 };
        const int nx = d->box_n[0] + 1, ny = d->box_n[1] + 1, nz = threeD ? d->box_n[2] + 1 : 1;
        for(int jj = 0; jj < nz; ++jj)
            for(int ii = 0; ii < ny; ++ii)
                for(int i = 0; i < nx; ++i)
                {
                    const size_t g = GI(d->box_lo[0] + i, d->box_lo[1] + ii, threeD ? d->box_lo[2] + jj : 0);
                    const double t1 = -0.5 * q->P[c][PB(i, ii, jj)] / q->eps[PB(i, ii, jj)];
                    E[g] = E[g] + 1.0 * t1;
                    const double t2 = -0.5 * q->P[c][PB(i + off[0], ii + off[1], jj + off[2])] / q->eps[PB(i + off[0], ii + off[1], jj + off[2])];
                    E[g] = E[g] + 1.0 * t2;
                }
    }
    if(!(part & 2)) return;
    /* zeroP_ (only the entries of this slab's emitters can be non-zero; rim rows owned by a neighbouring slab are kept) */
    for(int c = 0; c < 3; ++c)
        if(s->f[CHIML_EX + c])
            for(int e = 0; e < d->nemit; ++e)
            {
                const int* l = &q->loc[3 * e];
                q->P[c][PB(l[0] + 1, l[1] + 1, l[2] - 1 + zOff + 1)] = 0.0;
            }
    const int sample = (q->tstep % d->pop_every) == 0;
    cx H[MAX_N2], pred[MAX_N2], fpred[MAX_N2];
    const double dt = d->dt;
    for(int sy = 0; sy < d->nsys; ++sy)
        for(int e = 0; e < d->nemit; ++e)
        {
            const int* l = &q->loc[3 * e];
            /* getE_TE / getE_TM (UTIL/FDTD_up_eq.cpp:1398-1430): the node field */
            double ev[3] = {0.0, 0.0, 0.0};
            const int gx = d->box_lo[0] + 1 + l[0], gy = d->box_lo[1] + 1 + l[1], gz = threeD ? d->box_lo[2] + 1 + l[2] : 0;
            for(int c = 0; c < 3; ++c)
            {
                const double* E = s->f[CHIML_EX + c];
                if(!E) continue;
                if(c == 2 && !s->f[CHIML_EX]) ev[c] = E[GI(gx, gy, gz)];                         /* TM: copied */
                else ev[c] = 0.5 * E[GI(gx, gy, gz)] + 0.5 * E[GI(gx - (c == 0), gy - (c == 1), gz - (c == 2 ? zOff : 0))];
            }
            const size_t base = (((size_t)sy * d->nemit + e) * n2);
            cx* rho = (cx*)q->st[0] + base;
            cx* f0 = (cx*)q->st[1] + base; cx* f1 = (cx*)q->st[2] + base; cx* f2 = (cx*)q->st[3] + base; cx* f3 = (cx*)q->st[4] + base;
            /* PCABAM4 (:751-770) */
            for(int k = 0; k < n2; ++k) pred[k] = rho[k];
            zaxpy_real(n2,  55.0 * dt / 24.0, f0, pred);
            zaxpy_real(n2, -59.0 * dt / 24.0, f1, pred);
            zaxpy_real(n2,  37.0 * dt / 24.0, f2, pred);
            zaxpy_real(n2,  -9.0 * dt / 24.0, f3, pred);
            qe_get_ham(q, sy, ev, H);
            qe_den_deriv(q, H, pred, fpred);
            zaxpy_real(n2,  9.0 * dt / 24.0, fpred, rho);
            zaxpy_real(n2, 19.0 * dt / 24.0, f0, rho);
            zaxpy_real(n2, -5.0 * dt / 24.0, f1, rho);
            zaxpy_real(n2,        dt / 24.0, f2, rho);
            for(int k = 0; k < n2; ++k) { f3[k] = f2[k]; f2[k] = f1[k]; f1[k] = f0[k]; }   /* Density::moveDensity */
            qe_den_deriv(q, H, rho, f0);
            /* QEPopDtc::inPop (ML/QEPopDtc.hpp:67): the FLAT index level_ */
            if(sample)
                for(int p = 0; p < d->npop; ++p)
                {
                    q->curpop[2 * p] = q->curpop[2 * p] + rho[q->pop_level[p]].re;
                    q->curpop[2 * p + 1] = q->curpop[2 * p + 1] + rho[q->pop_level[p]].im;
                }
            /* updateQEPol (UTIL/FDTD_up_eq.hpp:620): P += na * Re(zdotc(rho, mu_c)) */
            for(int c = 0; c < 3; ++c)
            {
                if(!s->f[CHIML_EX + c]) continue;
                const cx* mu = (const cx*)q->mu + (size_t)c * n2;
                cx acc; acc.re = 0.0; acc.im = 0.0;
                for(int k = 0; k < n2; ++k)
                {
                    cx cj; cj.re = rho[k].re; cj.im = -rho[k].im;
                    acc = cadd(acc, cmul(cj, mu[k]));
                }
                const size_t pi = PB(l[0] + 1, l[1] + 1, l[2] - 1 + zOff + 1);
                q->P[c][pi] = q->P[c][pi] + d->na * acc.re;
            }
        }
    /* QEPopDtc::accumPop (ML/QEPopDtc.cpp:28-35) */
    if(sample && d->npop > 0)
    {
        if(q->pop_n >= q->pop_cap)
        {
            double* np_ = (double*)calloc((size_t)d->npop * q->pop_cap * 2 * 2, sizeof(double));
            for(int p = 0; p < d->npop; ++p) memcpy(np_ + (size_t)p * q->pop_cap * 2 * 2, q->pop + (size_t)p * q->pop_cap * 2, q->pop_n * 2 * sizeof(double));
            free(q->pop); q->pop = np_; q->pop_cap *= 2;
        }
        for(int p = 0; p < d->npop; ++p)
        {
            /* complex / double */
            q->pop[((size_t)p * q->pop_cap + q->pop_n) * 2] = q->curpop[2 * p] / (double)d->npoints;
            q->pop[((size_t)p * q->pop_cap + q->pop_n) * 2 + 1] = q->curpop[2 * p + 1] / (double)d->npoints;
        }
        ++q->pop_n;
    }
    for(int p = 0; p < d->npop; ++p) { q->curpop[2 * p] = 0.0; q->curpop[2 * p + 1] = 0.0; }
    ++q->tstep;
#undef PB
#undef GI
}

double* oracle_emitter_state(OracleSim* s, int slot, int which) { return (slot < 0 || slot >= s->nqe || which < 0 || which > 4) ? NULL : s->qe[slot].st[which]; }
double* oracle_emitter_P(OracleSim* s, int slot, int comp) { return (slot < 0 || slot >= s->nqe || comp < 0 || comp > 2) ? NULL : s->qe[slot].P[comp]; }
size_t  oracle_population(OracleSim* s, int slot, int det, double* out, size_t cap)
{
    if(slot < 0 || slot >= s->nqe || det < 0 || det >= s->qe[slot].d.npop) return 0;
    QESet* q = &s->qe[slot];
    size_t n = q->pop_n < cap ? q->pop_n : cap;
    if(out && n) memcpy(out, q->pop + (size_t)det * q->pop_cap * 2, n * 2 * sizeof(double));
    return q->pop_n;
}

#define SPLIT(n, lo, hi) size_t lo = (size_t)(n) * (size_t)tid / (size_t)nt, hi = (size_t)(n) * (size_t)(tid + 1) / (size_t)nt
#define BARRIER() do { if(nt > 1) pthread_barrier_wait(&s->bar); } while(0)

static void pml_component(OracleSim* s, int comp, double* target, int tid, int nt)
{
    /* parallelCPML<T>::updateGrid (PML/parallelPML.hpp:693-697): part 0 then part 1 */
    const int isE = comp < 3;
    const int i = comp % 3;
    const int base = isE ? CHIML_HX : CHIML_EX;           /* the PML of an E component is driven by H and vice versa */
    for(int part = 0; part < 2; ++part)
    {
        CpmlPart* p = &s->pml[comp][part];
        if(!p->present) { continue; }
        /* part 0: psi_j driven by grid_k; part 1: psi_k driven by grid_j (:695-696) */
        const double* V = s->f[base + (part == 0 ? (i + 2) % 3 : (i + 1) % 3)];
        if(!V) continue;
        if(p->has_psi)
        {
            SPLIT(p->npsi, lo, hi);
            for(size_t e = lo; e < hi; ++e) psi_entry(&p->psi[e], p->psi_grid, V);
        }
        BARRIER();
        {
            SPLIT(p->ngrid, lo, hi);
            for(size_t e = lo; e < hi; ++e) pml_grid_entry(&p->grid[e], target, p->has_psi ? p->psi_grid : NULL, V);
        }
        BARRIER();
    }
}

/* TFSF surfaces: tfsfUpdateFxnReal::addIncdFields / addIncdFieldsEPChange (SOURCE/parallelTFSF.cpp:77-105) on the records genSurface built
 * (SOURCE/parallelTFSF.hpp:823-998).  The incident lines come from the caller (the reference's own 1-D line, stepped on the host). */
int oracle_add_tfsf_surface(OracleSim* s, const ChimlTfsfSurface* t)
{
    if(!s || !t || s->ntfsf >= MAX_TFSF || t->comp < 0 || t->comp > 5 || !s->f[t->comp] || s->g.nranks != 1) return CHIML_ERR_ARG;
    if(t->npairs_D > 0 && (t->comp > 2 || !s->f[CHIML_DX + t->comp])) return CHIML_ERR_ARG;
    TfsfSur* q = &s->tfsf[s->ntfsf++];
    q->s = *t;
    q->pairs_D = (int32_t*)malloc((size_t)(2 * t->npairs_D + 1) * sizeof(int32_t));
    q->pairs_U = (int32_t*)malloc((size_t)(2 * t->npairs_U + 1) * sizeof(int32_t));
    if(t->npairs_D) memcpy(q->pairs_D, t->pairs_D, (size_t)(2 * t->npairs_D) * sizeof(int32_t));
    if(t->npairs_U) memcpy(q->pairs_U, t->pairs_U, (size_t)(2 * t->npairs_U) * sizeof(int32_t));
    q->ep_mu = NULL;
    if(t->ep_mu) { q->ep_mu = (double*)malloc((size_t)t->incd_len * sizeof(double)); memcpy(q->ep_mu, t->ep_mu, (size_t)t->incd_len * sizeof(double)); }
    return 0;
}
static void tfsf_add(OracleSim* s, const TfsfSur* q, const double* table)
{
    const double* incd = table + q->s.incd_offset;
    const int n = q->s.n;
    const long si = q->s.stride_incd, sm = q->s.stride_main;
    /* daxpy_(n, prefactor, Re incd, 2 * strideIncd, grid, strideMain): a negative increment starts at the far end (BLAS convention),
     * i.e. element i of x is x[(i - (n - 1)) * inc] ... walked together with element i of y */
    double* D = q->s.comp < 3 ? s->f[CHIML_DX + q->s.comp] : NULL;
    double* U = s->f[q->s.comp];
    for(int l = 0; l < q->s.npairs_D; ++l)
    {
        const long i0 = q->pairs_D[2 * l], m0 = q->pairs_D[2 * l + 1];
        const long ix0 = si < 0 ? (long)(1 - n) * si : 0;
        for(int i = 0; i < n; ++i) D[m0 + i * sm] = D[m0 + i * sm] + q->s.prefactor * incd[i0 + ix0 + i * si];
    }
    for(int l = 0; l < q->s.npairs_U; ++l)
    {
        const long i0 = q->pairs_U[2 * l], m0 = q->pairs_U[2 * l + 1];
        const long ix0 = si < 0 ? (long)(1 - n) * si : 0;
        for(int i = 0; i < n; ++i)
        {
            double v = incd[i0 + ix0 + i * si];
            if(q->ep_mu) v = v / q->ep_mu[i0 + ix0 + i * si];       /* dcopy (same negative-increment convention), then std::divides */
            U[m0 + i * sm] = U[m0 + i * sm] + q->s.prefactor * v;
        }
    }
}

/* applyBC1Proc, real fields (UTIL/FDTD_up_eq.cpp:1058-1116): the periodic wrap copies of one component on a single process, in the
 * reference's order of dcopy_ calls.  PT(x, y, z) = parallelGrid::point(x, y, z) (GRID/parallelGrid.hpp:363). */
int oracle_set_periodic(OracleSim* s, int comp, const ChimlWrap* w)
{
    if(!s || comp < 0 || comp > 5 || !w) return CHIML_ERR_ARG;
    if((s->g.nranks != 1) != (w->ymax < 0)) return CHIML_ERR_ARG;      /* a slab of several takes the x / z wraps only (ymax = -1) */
    s->wrap[comp] = *w; s->has_wrap[comp] = 1;
    return 0;
}
static void copy_strided(int n, const double* x, size_t incx, double* y, size_t incy) { for(int i = 0; i < n; ++i) y[(size_t)i * incy] = x[(size_t)i * incx]; }
static void apply_bc_1proc(const OracleSim* s, double* F, const ChimlWrap* w)
{
    const size_t lx = (size_t)s->g.ln[0], lz = (size_t)s->g.ln[2];
#define PT(x, y, z) (F + ((size_t)(x) + lx * ((size_t)(z) + lz * (size_t)(y))))
    const int nx = w->nx, ny = w->ny, nz = w->nz, xmax = w->xmax, ymax = w->ymax, zmin = w->zmin, zmax = w->zmax;
    if(ymax < 0)
    {
        /* one y-slab of several (applyBCProcMid, UTIL/FDTD_up_eq.cpp:1036-1061): the y direction is the ghost-row ring between the slabs
         * (transferDat); here only the x / z ghost cells of the owned rows take their periodic images */
        const int ly = s->g.ln[1];
        if(zmin != 0)
        {
            for(int jj = 1; jj < ly - 1; ++jj)
            {
                copy_strided(nz, PT(xmax - 1, jj, 1), lx, PT(0, jj, 1), lx);
                copy_strided(nz, PT(1, jj, 1), lx, PT(xmax, jj, 1), lx);
                copy_strided(nx, PT(1, jj, zmax - 1), 1, PT(1, jj, zmin - 1), 1);
                copy_strided(nx, PT(1, jj, zmin), 1, PT(1, jj, zmax), 1);
            }
            copy_strided(ly - 2, PT(1, 1, zmin), lx * lz, PT(xmax, 1, zmax), lx * lz);
            copy_strided(ly - 2, PT(xmax - 1, 1, zmax - 1), lx * lz, PT(0, 1, zmin - 1), lx * lz);
            copy_strided(ly - 2, PT(xmax - 1, 1, zmin), lx * lz, PT(0, 1, zmax), lx * lz);
            copy_strided(ly - 2, PT(1, 1, zmax - 1), lx * lz, PT(xmax, 1, zmin - 1), lx * lz);
        }
        else
        {
            copy_strided(ly - 2, PT(xmax - 1, 1, 0), lx, PT(0, 1, 0), lx);
            copy_strided(ly - 2, PT(1, 1, 0), lx, PT(xmax, 1, 0), lx);
        }
        return;
    }
    if(zmin != 0)
    {
        for(int kk = zmin; kk <= nz; ++kk)
        {
            copy_strided(nx, PT(1, ymax - 1, kk), 1, PT(1, 0, kk), 1);
            copy_strided(nx, PT(1, 1, kk), 1, PT(1, ymax, kk), 1);
        }
        for(int jj = 1; jj < ny; ++jj)
        {
            copy_strided(nz, PT(xmax - 1, jj, 1), lx, PT(0, jj, 1), lx);
            copy_strided(nz, PT(1, jj, 1), lx, PT(xmax, jj, 1), lx);
            copy_strided(nx, PT(1, jj, zmax - 1), 1, PT(1, jj, zmin - 1), 1);
            copy_strided(nx, PT(1, jj, zmin), 1, PT(1, jj, zmax), 1);
        }
        /* X edges */
        copy_strided(nx, PT(1, 1, zmin), 1, PT(1, ymax, zmax), 1);
        copy_strided(nx, PT(1, ymax - 1, zmin), 1, PT(1, 0, zmax), 1);
        copy_strided(nx, PT(1, 1, zmax - 1), 1, PT(1, ymax, zmin - 1), 1);
        copy_strided(nx, PT(1, ymax - 1, zmax - 1), 1, PT(1, 0, zmin - 1), 1);
        /* Y edges */
        copy_strided(ny - 1, PT(1, 1, zmin), lx * lz, PT(xmax, 1, zmax), lx * lz);
        copy_strided(ny - 1, PT(xmax - 1, 1, zmin), lx * lz, PT(0, 1, zmax), lx * lz);
        copy_strided(ny - 1, PT(1, 1, zmax - 1), lx * lz, PT(xmax, 1, zmin - 1), lx * lz);
        copy_strided(ny - 1, PT(xmax - 1, 1, zmax - 1), lx * lz, PT(0, 1, zmin - 1), lx * lz);
        /* Z edges */
        copy_strided(nz, PT(1, 1, 1), lx, PT(xmax, ymax, 1), lx);
        copy_strided(nz, PT(xmax - 1, 1, 1), lx, PT(0, ymax, 1), lx);
        copy_strided(nz, PT(1, ymax - 1, 1), lx, PT(xmax, 0, 1), lx);
        copy_strided(nz, PT(xmax - 1, ymax - 1, 1), lx, PT(0, 0, 1), lx);
        /* corners */
        *PT(xmax, ymax, zmax) = *PT(1, 1, zmin);
        *PT(0, ymax, zmax) = *PT(xmax - 1, 1, zmin);
        *PT(xmax, 0, zmax) = *PT(1, ymax - 1, zmin);
        *PT(0, 0, zmax) = *PT(xmax - 1, ymax - 1, zmin);
        *PT(xmax, ymax, zmin - 1) = *PT(1, 1, zmax - 1);
        *PT(0, ymax, zmin - 1) = *PT(xmax - 1, 1, zmax - 1);
        *PT(xmax, 0, zmin - 1) = *PT(1, ymax - 1, zmax - 1);
        *PT(0, 0, zmin - 1) = *PT(xmax - 1, ymax - 1, zmax - 1);
    }
    else
    {
        copy_strided(nx, PT(1, ymax - 1, 0), 1, PT(1, 0, 0), 1);
        copy_strided(nx, PT(1, 1, 0), 1, PT(1, ymax, 0), 1);
        copy_strided(ny, PT(xmax - 1, 1, 0), lx, PT(0, 1, 0), lx);
        copy_strided(ny, PT(1, 1, 0), lx, PT(xmax, 1, 0), lx);
    }
#undef PT
}

/* ---- complex fields (Bloch-periodic runs): the real and imaginary parts are two simulations over the same lists -- every operator of the
 * step has real coefficients, so the reference's complex BLAS chains (zaxpy with a real factor etc.) act on the two parts separately --
 * coupled only by the phase factors of the wrap copies.  applyBC1Proc, complex fields (UTIL/FDTD_up_eq.cpp:1248-1324), call by call:
 * zcopy_ then zscal_ with exp(i k.L) (dest = phase * dest, netlib product), the corners as written there -- they read the already
 * wrapped row ymax. */
typedef struct { double re, im; } cxd;
static cxd phase_of(double arg) { const double _Complex w = cexp(__builtin_complex(0.0, arg)); cxd r; r.re = creal(w); r.im = cimag(w); return r; }
static void zcopy_scal(int n, const double* sr, const double* si, size_t incs, double* dr, double* di, size_t incd, cxd a)
{
    for(int i = 0; i < n; ++i)
    {
        const double xr = sr[(size_t)i * incs], xi = si[(size_t)i * incs];
        dr[(size_t)i * incd] = a.re * xr - a.im * xi;
        di[(size_t)i * incd] = a.re * xi + a.im * xr;
    }
}
static void apply_bc_1proc_cplx(const OracleSim* s, double* R, double* Iq, const ChimlWrap* w, const double k[3])
{
    const size_t lx = (size_t)s->g.ln[0], lz = (size_t)s->g.ln[2];
    const double dx = s->g.d[0], dy = s->g.d[1], dz = s->g.d[2];
#define OF(x, y, z) ((size_t)(x) + lx * ((size_t)(z) + lz * (size_t)(y)))
#define CS(n, sx, sy, sz, inc, tx, ty, tz, arg) zcopy_scal(n, R + OF(sx, sy, sz), Iq + OF(sx, sy, sz), inc, R + OF(tx, ty, tz), Iq + OF(tx, ty, tz), inc, phase_of(arg))
    const int nx = w->nx, ny = w->ny, nz = w->nz, xmax = w->xmax, ymax = w->ymax, zmin = w->zmin, zmax = w->zmax;
    if(zmin != 0)
    {
        for(int kk = zmin; kk <= nz; ++kk)
        {
            CS(nx, 1, ymax - 1, kk, 1, 1, 0, kk, -1.0 * k[1] * dy * ymax);
            CS(nx, 1, 1, kk, 1, 1, ymax, kk, k[1] * dy * ymax);
        }
        for(int jj = 1; jj < ny; ++jj)
        {
            CS(nz, xmax - 1, jj, 1, lx, 0, jj, 1, -1.0 * k[0] * dx * xmax);
            CS(nz, 1, jj, 1, lx, xmax, jj, 1, k[0] * dx * xmax);
            CS(nx, 1, jj, zmax - 1, 1, 1, jj, zmin - 1, -1.0 * k[2] * dz * zmax);
            CS(nx, 1, jj, zmin, 1, 1, jj, zmax, k[2] * dz * zmax);
        }
        /* Y edges */
        CS(ny - 1, 1, 1, zmin, lx * lz, xmax, 1, zmax, k[0] * dx * xmax + k[2] * dz * zmax);
        CS(ny - 1, xmax - 1, 1, zmin, lx * lz, 0, 1, zmax, -1.0 * k[0] * dx * xmax + k[2] * dz * zmax);
        CS(ny - 1, 1, 1, zmax - 1, lx * lz, xmax, 1, zmin - 1, k[0] * dx * xmax - k[2] * dz * zmax);
        CS(ny - 1, xmax - 1, 1, zmax - 1, lx * lz, 0, 1, zmin - 1, -1.0 * k[0] * dx * xmax - k[2] * dz * zmax);
        /* X edges */
        CS(nx, 1, 1, zmin, 1, 1, ymax, zmax, k[1] * dy * ymax + k[2] * dz * zmax);
        CS(nx, 1, 1, zmax - 1, 1, 1, ymax, zmin - 1, k[1] * dy * ymax - k[2] * dz * zmax);
        CS(nx, 1, ymax - 1, zmin, 1, 1, 0, zmax, -1.0 * k[1] * dy * ymax + k[2] * dz * zmax);
        CS(nx, 1, ymax - 1, zmax - 1, 1, 1, 0, zmin - 1, -1.0 * k[1] * dy * ymax - k[2] * dz * zmax);
        /* Z edges */
        CS(nz, 1, 1, 1, lx, xmax, ymax, 1, k[0] * dx * xmax + k[1] * dy * ymax);
        CS(nz, xmax - 1, 1, 1, lx, 0, ymax, 1, -1.0 * k[0] * dx * xmax + k[1] * dy * ymax);
        CS(nz, 1, ymax - 1, 1, lx, xmax, 0, 1, k[0] * dx * xmax - k[1] * dy * ymax);
        CS(nz, xmax - 1, ymax - 1, 1, lx, 0, 0, 1, -1.0 * k[0] * dx * xmax - k[1] * dy * ymax);
        /* corners: every one of them reads row ymax, which the first loop has already filled with the phased image of row 1 */
        CS(1, 1, ymax, zmin, 1, xmax, ymax, zmax, k[0] * dx * xmax + k[1] * dy * ymax + k[2] * dz * zmax);
        CS(1, xmax - 1, ymax, zmin, 1, 0, ymax, zmax, -1.0 * k[0] * dx * xmax + k[1] * dy * ymax + k[2] * dz * zmax);
        CS(1, 1, ymax, zmax - 1, 1, xmax, ymax, zmin - 1, k[0] * dx * xmax + k[1] * dy * ymax - k[2] * dz * zmax);
        CS(1, xmax - 1, ymax, zmax - 1, 1, 0, ymax, zmin - 1, -1.0 * k[0] * dx * xmax + k[1] * dy * ymax - k[2] * dz * zmax);
        CS(1, 1, ymax, zmin, 1, xmax, 0, zmax, k[0] * dx * xmax - k[1] * dy * ymax + k[2] * dz * zmax);
        CS(1, xmax - 1, ymax, zmin, 1, 0, 0, zmax, -1.0 * k[0] * dx * xmax - k[1] * dy * ymax + k[2] * dz * zmax);
        CS(1, 1, ymax, zmax - 1, 1, xmax, 0, zmin - 1, k[0] * dx * xmax - k[1] * dy * ymax - k[2] * dz * zmax);
        CS(1, xmax - 1, ymax, zmax - 1, 1, 0, 0, zmin - 1, -1.0 * k[0] * dx * xmax - k[1] * dy * ymax - k[2] * dz * zmax);
    }
    else
    {
        CS(nx, 1, ymax - 1, 0, 1, 1, 0, 0, -1.0 * k[1] * dy * ymax);
        CS(nx, 1, 1, 0, 1, 1, ymax, 0, k[1] * dy * ymax);
        CS(ny, xmax - 1, 1, 0, lx, 0, 1, 0, -1.0 * k[0] * dx * xmax);
        CS(ny, 1, 1, 0, lx, xmax, 1, 0, k[0] * dx * xmax);
    }
#undef CS
#undef OF
}
int oracle_step_phase(OracleSim* s, int phase, const double* src_amp);
/* n steps of a complex-field run: re / im = the two parts (no periodic wraps of their own, no running-DFT sets, no emitters), amp_re / amp_im =
 * dt * Re / Im(sum pulse) per step and source, wrap / has_wrap = the applBCE_ / applBCH_ arguments per component */
int oracle_pair_step_n(OracleSim* re, OracleSim* im, int n, const double* amp_re, const double* amp_im, const ChimlWrap* wrap, const int* has_wrap,
                       const double* k_point)
{
    if(!re || !im || !wrap || !has_wrap || !k_point || re->ncell != im->ncell || re->ndft || im->ndft || re->nqe || im->nqe || re->ntfsf || im->ntfsf)
        return CHIML_ERR_ARG;
    for(int c = 0; c < 6; ++c) if(re->has_wrap[c] || im->has_wrap[c]) return CHIML_ERR_ARG;
    const size_t ns = (size_t)re->nsrc;
    for(int k = 0; k < n; ++k)
    {
        int rc;
        if((rc = oracle_step_phase(re, 0, amp_re + (size_t)k * ns)) || (rc = oracle_step_phase(im, 0, amp_im + (size_t)k * ns))) return rc;
        for(int i = 0; i < 3; ++i)          /* applBCH_ (FDTD_MANAGER/parallelFDTDField.hpp:1267-1269) */
            if(re->f[CHIML_HX + i] && has_wrap[3 + i]) apply_bc_1proc_cplx(re, re->f[CHIML_HX + i], im->f[CHIML_HX + i], &wrap[3 + i], k_point);
        for(int phase = 1; phase <= 3; ++phase)
            if((rc = oracle_step_phase(re, phase, amp_re + (size_t)k * ns)) || (rc = oracle_step_phase(im, phase, amp_im + (size_t)k * ns))) return rc;
        for(int i = 0; i < 3; ++i)          /* applBCE_ (:1285-1287) */
            if(re->f[CHIML_EX + i] && has_wrap[i]) apply_bc_1proc_cplx(re, re->f[CHIML_EX + i], im->f[CHIML_EX + i], &wrap[i], k_point);
    }
    return 0;
}

static void step_worker(OracleSim* s, int tid, int nt)
{
    const int lnx = s->g.ln[0], lnz = s->g.ln[2];
    double* scratch = (double*)malloc((size_t)(6 * lnx + 8) * sizeof(double));
    for(int step = 0; step < s->nsteps; ++step)
    {
        /* updateMagH (:1230, :1370-1387): magnetic poles from H^n, before any H / B update of the step */
        if(s->has_B)
        {
            for(int i = 0; i < 3 && (s->phase_mask & 1); ++i)
            {
                if(!s->f[CHIML_HX + i] || !s->f[CHIML_BX + i]) continue;
                RunList* l = &s->up[CHIML_LIST_LORD][3 + i];
                SPLIT(l->n, lo, hi);
                for(size_t e = lo; e < hi; ++e) lor_pol_run(&l->r[e], s->f[CHIML_HX + i], s->M[i], s->Mprev[i], &s->mobj[l->r[e].obj], scratch);
            }
            BARRIER();
        }
        /* updateChiH (:1231, :1422-1447): achiral magnetic poles of the chiral cells, chiral magnetisation from E and prevE, then E -> prevE */
        if(s->nchi > 0 && (s->phase_mask & 1))
        {
            for(int i = 0; i < 3; ++i)
            {
                RunList* l = &s->up[CHIML_LIST_CHID][3 + i];
                SPLIT(l->n, lo, hi);
                for(size_t e = lo; e < hi; ++e)
                {
                    lor_pol_run(&l->r[e], s->f[CHIML_HX + i], s->M[i], s->Mprev[i], &s->mobj[l->r[e].obj], scratch);
                    chiral_run(&l->r[e], s->f[CHIML_EX + i], s->prevf[i], s->chi[3 + i], s->chiprev[3 + i], &s->cobj[l->r[e].obj], scratch);
                }
            }
            BARRIER();
            if(tid == 0) prev_copy(s, 1);
            BARRIER();
        }
        /* updateB (:1234, :1328-1333) and updateH (:1308-1313) */
        for(int i = 0; i < 3 && (s->phase_mask & 1); ++i)
        {
            double* H = s->f[CHIML_HX + i];
            if(!H) continue;
            if(s->has_B && s->f[CHIML_BX + i])
            {
                RunList* lb = &s->up[CHIML_LIST_D][3 + i];
                SPLIT(lb->n, lo, hi);
                for(size_t e = lo; e < hi; ++e) curl_run(&lb->r[e], s->f[CHIML_BX + i], s->f[CHIML_EX + (i + 1) % 3], s->f[CHIML_EX + (i + 2) % 3]);
            }
            RunList* l = &s->up[CHIML_LIST_U][3 + i];
            SPLIT(l->n, lo, hi);
            for(size_t e = lo; e < hi; ++e) curl_run(&l->r[e], H, s->f[CHIML_EX + (i + 1) % 3], s->f[CHIML_EX + (i + 2) % 3]);
        }
        BARRIER();
        /* tfsf->updateFields() (:1238-1255, SOURCE/parallelTFSF.hpp:1058-1073): H surfaces, the line's own step (host), E surfaces */
        if(tid == 0 && (s->phase_mask & 1) && s->ntfsf > 0)
        {
            const double* table = s->tfsf_incd + (size_t)step * s->tfsf_per_step;
            for(int q = 0; q < s->ntfsf; ++q) if(s->tfsf[q].s.comp >= 3) tfsf_add(s, &s->tfsf[q], table);
            for(int q = 0; q < s->ntfsf; ++q) if(s->tfsf[q].s.comp < 3) tfsf_add(s, &s->tfsf[q], table);
        }
        BARRIER();
        /* updateHPML_ (:1258-1259) */
        for(int i = 0; i < 3 && (s->phase_mask & 1); ++i)
            if(s->f[CHIML_HX + i]) pml_component(s, 3 + i, (s->has_B && s->pml_on_B) ? s->f[CHIML_BX + i] : s->f[CHIML_HX + i], tid, nt);   /* on B when magMatInPML_ (parallelFDTDField.cpp:229-246) */
        /* src->addPul (:1261-1262, SOURCE/parallelSourceNormal.cpp:15-37): grid[box] += dt*Re(pulse), amp precomputed */
        if(tid == 0 && (s->phase_mask & 1))
        {
            for(int q = 0; q < s->nsrc; ++q)
            {
                const SrcBox* b = &s->src[q];
                const double amp = s->src_amp[(size_t)step * (size_t)s->nsrc + q];
                double* G = s->f[b->field];
                for(int y = 0; y < b->sz[1]; ++y)
                    for(int z = 0; z < b->sz[2]; ++z)
                        for(int x = 0; x < b->sz[0]; ++x)
                        {
                            size_t ind = (size_t)(b->loc[0] + x) + (size_t)lnx * ((size_t)(b->loc[2] + z) + (size_t)lnz * (size_t)(b->loc[1] + y));
                            G[ind] = G[ind] + amp;
                        }
            }
        }
        /* B2H (:1264, :1478-1500): H = B / mu_inf - sum M / mu_inf on the cells of upLorB_, after the sources */
        if(s->has_B && (s->phase_mask & 1))
        {
            BARRIER();
            for(int i = 0; i < 3; ++i)
            {
                if(!s->f[CHIML_HX + i] || !s->f[CHIML_BX + i]) continue;
                RunList* l = &s->up[CHIML_LIST_LORD][3 + i];
                SPLIT(l->n, lo, hi);
                for(size_t e = lo; e < hi; ++e) dtou_run(&l->r[e], s->f[CHIML_BX + i], s->f[CHIML_HX + i], s->M[i], s->nmag);
                RunList* lc = &s->up[CHIML_LIST_CHID][3 + i];
                SPLIT(lc->n, lo2, hi2);
                for(size_t e = lo2; e < hi2; ++e)
                {
                    dtou_run(&lc->r[e], s->f[CHIML_BX + i], s->f[CHIML_HX + i], s->M[i], s->nmag);
                    chi_dtou_run(&lc->r[e], lc->r[e].pf[3], s->f[CHIML_HX + i], s->chi[3 + i], s->nchi);
                }
            }
            BARRIER();
        }
        /* applBCH_ (:1267-1269): periodic wrap copies of the H components */
        if(tid == 0 && (s->phase_mask & 1))
            for(int i = 0; i < 3; ++i)
                if(s->f[CHIML_HX + i] && s->has_wrap[3 + i]) apply_bc_1proc(s, s->f[CHIML_HX + i], &s->wrap[3 + i]);
        BARRIER();
        /* updatePolE (:1348-1365): oriented-dipole poles at nodes, then isotropic poles per component */
        if(s->phase_mask & 2)
        {
            RunList* l = &s->up[CHIML_LIST_ORDIPP][0];
            SPLIT(l->n, lo, hi);
            for(size_t e = lo; e < hi; ++e) lor_pol_ordip_run(s, &l->r[e], &s->obj[l->r[e].obj], scratch);
        }
        for(int i = 0; i < 3 && (s->phase_mask & 4); ++i)
        {
            if(!s->f[CHIML_EX + i] || !s->f[CHIML_DX + i]) continue;
            RunList* l = &s->up[CHIML_LIST_LORD][i];
            SPLIT(l->n, lo, hi);
            for(size_t e = lo; e < hi; ++e) lor_pol_run(&l->r[e], s->f[CHIML_EX + i], s->P[i], s->Pprev[i], &s->obj[l->r[e].obj], scratch);
        }
        BARRIER();
        /* updateChiE (:1273, :1392-1417): achiral poles of the chiral cells, chiral polarisation from H and prevH, then H -> prevH */
        if(s->nchi > 0 && (s->phase_mask & 4))
        {
            for(int i = 0; i < 3; ++i)
            {
                RunList* l = &s->up[CHIML_LIST_CHID][i];
                SPLIT(l->n, lo, hi);
                for(size_t e = lo; e < hi; ++e)
                {
                    lor_pol_run(&l->r[e], s->f[CHIML_EX + i], s->P[i], s->Pprev[i], &s->obj[l->r[e].obj], scratch);
                    chiral_run(&l->r[e], s->f[CHIML_HX + i], s->prevf[3 + i], s->chi[i], s->chiprev[i], &s->cobj[l->r[e].obj], scratch);
                }
            }
            BARRIER();
            if(tid == 0) prev_copy(s, 0);
            BARRIER();
        }
        /* updateD (:1338-1343) and updateE (:1318-1323) */
        for(int i = 0; i < 3 && (s->phase_mask & 4); ++i)
        {
            if(!s->f[CHIML_EX + i]) continue;
            const double* Hj = s->f[CHIML_HX + (i + 1) % 3];
            const double* Hk = s->f[CHIML_HX + (i + 2) % 3];
            if(s->f[CHIML_DX + i])
            {
                RunList* l = &s->up[CHIML_LIST_D][i];
                SPLIT(l->n, lo, hi);
                for(size_t e = lo; e < hi; ++e) curl_run(&l->r[e], s->f[CHIML_DX + i], Hj, Hk);
            }
            RunList* l = &s->up[CHIML_LIST_U][i];
            SPLIT(l->n, lo, hi);
            for(size_t e = lo; e < hi; ++e) curl_run(&l->r[e], s->f[CHIML_EX + i], Hj, Hk);
        }
        BARRIER();
        /* updateEPML_ (:1279-1280): acts on D when material reaches the PML (parallelFDTDField.cpp:68-77,239-246) */
        for(int i = 0; i < 3 && (s->phase_mask & 4); ++i)
            if(s->f[CHIML_EX + i]) pml_component(s, i, s->g.pml_on_D ? s->f[CHIML_DX + i] : s->f[CHIML_EX + i], tid, nt);
        /* D2E (:1452-1473) */
        for(int i = 0; i < 3 && (s->phase_mask & 4); ++i)
        {
            if(!s->f[CHIML_EX + i] || !s->f[CHIML_DX + i]) continue;
            RunList* l = &s->up[CHIML_LIST_LORD][i];
            {
                SPLIT(l->n, lo, hi);
                for(size_t e = lo; e < hi; ++e) dtou_run(&l->r[e], s->f[CHIML_DX + i], s->f[CHIML_EX + i], s->P[i], s->npoles);
            }
            if(s->nchi > 0)
            {
                RunList* lc = &s->up[CHIML_LIST_CHID][i];
                SPLIT(lc->n, lo, hi);
                for(size_t e = lo; e < hi; ++e)
                {
                    dtou_run(&lc->r[e], s->f[CHIML_DX + i], s->f[CHIML_EX + i], s->P[i], s->npoles);
                    chi_dtou_run(&lc->r[e], -1.0 * lc->r[e].pf[3], s->f[CHIML_EX + i], s->chi[i], s->nchi);
                }
            }
            l = &s->up[CHIML_LIST_ORDIPD][i];
            {
                /* orDipDtoUZ only for Ez without Hz (FDTD_MANAGER/parallelFDTDField.cpp:293-296) */
                const int zvariant = (i == 2 && !s->f[CHIML_HZ]);
                SPLIT(l->n, lo, hi);
                for(size_t e = lo; e < hi; ++e) ordip_dtou_run(&l->r[e], s->f[CHIML_DX + i], s->f[CHIML_EX + i], s->oP[i], s->nordip, zvariant);
            }
        }
        BARRIER();
        /* qe->addQE() for every emitter object (:1282-1283); serial, as each reference rank runs it */
        if(tid == 0 && (s->phase_mask & 4))
            for(int q = 0; q < s->nqe; ++q) qe_add(s, &s->qe[q], 1);
        if(tid == 0 && (s->phase_mask & 8))
            for(int q = 0; q < s->nqe; ++q) qe_add(s, &s->qe[q], 2);
        /* applBCE_ (:1285-1287) */
        if(tid == 0 && (s->phase_mask & 8))
            for(int i = 0; i < 3; ++i)
                if(s->f[CHIML_EX + i] && s->has_wrap[i]) apply_bc_1proc(s, s->f[CHIML_EX + i], &s->wrap[i]);
        /* flux->fieldIn(tcur_) (:1300-1302): two dger_ rank-1 updates per line, F(f, i) += 1.0 * tw[f] * u[i] */
        if(tid == 0 && (s->phase_mask & 8))
        {
            ++s->step_count;
            size_t per_step = 0, goff[256];
            for(int g = 0; g < 256; ++g) { goff[g] = per_step; per_step += 2 * (size_t)s->dft_group_nfreq[g]; }
            for(int q = 0; q < s->ndft; ++q)
            {
                DftSet* d = &s->dft[q];
                if(s->step_count % d->every != 0) continue;
                const double* tw = s->twiddles + (size_t)step * per_step + goff[d->group];
                const double* G = s->f[d->field];
                for(size_t l = 0; l < d->nlines; ++l)
                    for(int i = 0; i < d->npts; ++i)
                    {
                        const double u = G[(size_t)d->lines[l].ind + (size_t)i * (size_t)d->stride];
                        const double t = 1.0 * u;
                        for(int f = 0; f < d->nfreq; ++f)
                        {
                            const size_t o = (size_t)d->lines[l].out + (size_t)f + (size_t)d->nfreq * (size_t)i;
                            d->re[o] = d->re[o] + tw[2 * f] * t;
                            d->im[o] = d->im[o] + tw[2 * f + 1] * t;
                        }
                    }
            }
        }
        BARRIER();
    }
    free(scratch);
}

static void* thread_main(void* arg)
{
    Worker* w = (Worker*)arg;
    step_worker(w->s, w->tid, w->s->nthreads);
    return NULL;
}

int oracle_step_n(OracleSim* s, int n, const double* src_amp, int nthreads);
int oracle_add_dft(OracleSim* s, int field, int group, int every, int nfreq, int npts, int stride, const ChimlDftLine* lines, size_t nlines, size_t acc_len)
{
    if(s->ndft >= MAX_DFT || group < 0 || group > 255 || field < 0 || field >= CHIML_NFIELDS || !s->f[field]) return CHIML_ERR_ARG;
    DftSet* d = &s->dft[s->ndft++];
    d->field = field; d->group = group; d->every = every; d->nfreq = nfreq; d->npts = npts; d->stride = stride; d->nlines = nlines; d->acc_len = acc_len;
    d->lines = (ChimlDftLine*)malloc((nlines ? nlines : 1) * sizeof(ChimlDftLine));
    if(nlines) memcpy(d->lines, lines, nlines * sizeof(ChimlDftLine));
    d->re = (double*)calloc(acc_len ? acc_len : 1, sizeof(double));
    d->im = (double*)calloc(acc_len ? acc_len : 1, sizeof(double));
    s->dft_group_nfreq[group] = nfreq;
    return 0;
}
double* oracle_dft(OracleSim* s, int slot, int imag) { return (slot < 0 || slot >= s->ndft) ? NULL : (imag ? s->dft[slot].im : s->dft[slot].re); }

int oracle_step_n_tfsf(OracleSim* s, int n, const double* src_amp, const double* twiddles, const double* incd, size_t incd_per_step, int nthreads)
{
    if(s->ntfsf > 0 && !incd) return CHIML_ERR_ARG;
    for(int q = 0; q < s->ntfsf; ++q) if((size_t)(s->tfsf[q].s.incd_offset + s->tfsf[q].s.incd_len) > incd_per_step) return CHIML_ERR_ARG;
    s->twiddles = twiddles; s->tfsf_incd = incd; s->tfsf_per_step = incd_per_step;
    const int rc = oracle_step_n(s, n, src_amp, nthreads);
    s->tfsf_incd = NULL;
    return rc;
}

int oracle_step_n_dft(OracleSim* s, int n, const double* src_amp, const double* twiddles, int nthreads)
{
    s->twiddles = twiddles;
    return oracle_step_n(s, n, src_amp, nthreads);
}

int oracle_step_n(OracleSim* s, int n, const double* src_amp, int nthreads)
{
    if(!s->committed) return CHIML_ERR_STATE;
    if(s->ndft > 0 && !s->twiddles) return CHIML_ERR_ARG;
    if(s->ntfsf > 0 && !s->tfsf_incd) return CHIML_ERR_ARG;
    if(nthreads < 1) nthreads = 1;
    s->nthreads = nthreads;
    s->src_amp = src_amp;
    s->nsteps = n;
    s->phase_mask = 15;
    if(nthreads == 1)
    {
        step_worker(s, 0, 1);
        return 0;
    }
    pthread_barrier_init(&s->bar, NULL, (unsigned)nthreads);
    pthread_t* th = (pthread_t*)malloc((size_t)nthreads * sizeof(pthread_t));
    Worker* w = (Worker*)malloc((size_t)nthreads * sizeof(Worker));
    for(int t = 0; t < nthreads; ++t) { w[t].s = s; w[t].tid = t; pthread_create(&th[t], NULL, thread_main, &w[t]); }
    for(int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    pthread_barrier_destroy(&s->bar);
    free(th); free(w);
    return 0;
}

/* one phase of ONE step (y-slab runs exchange ghost rows between the phases): 0 = H half step + sources, 1 = oriented-dipole
 * node poles, 2 = E half step + emitter addP, 3 = emitter density update.  src_amp = the n_sources amplitudes of this step. */
int oracle_step_phase(OracleSim* s, int phase, const double* src_amp)
{
    if(!s->committed || phase < 0 || phase > 3) return CHIML_ERR_STATE;
    if(phase == 3 && s->ndft > 0 && !s->twiddles) return CHIML_ERR_ARG;
    s->nthreads = 1;
    s->src_amp = src_amp;
    s->nsteps = 1;
    s->phase_mask = 1 << phase;
    step_worker(s, 0, 1);
    return 0;
}

/* same for runs with running-DFT sets: twiddles = the step's exp(-i freq t) of every group (read in phase 3) */
int oracle_step_phase_dft(OracleSim* s, int phase, const double* src_amp, const double* twiddles)
{
    s->twiddles = twiddles;
    return oracle_step_phase(s, phase, src_amp);
}

double* oracle_field(OracleSim* s, int field) { return (field >= 0 && field < CHIML_NFIELDS) ? s->f[field] : NULL; }
double* oracle_pole(OracleSim* s, int comp, int pole, int prev) { return (comp < 0 || comp > 2 || pole < 0 || pole >= MAX_POLES) ? NULL : (prev ? s->Pprev[comp][pole] : s->P[comp][pole]); }
double* oracle_ordip_pole(OracleSim* s, int comp, int pole, int prev) { return (comp < 0 || comp > 2 || pole < 0 || pole >= MAX_POLES) ? NULL : (prev ? s->oPprev[comp][pole] : s->oP[comp][pole]); }
double* oracle_psi(OracleSim* s, int comp, int part) { return (comp < 0 || comp > 5 || part < 0 || part > 1) ? NULL : s->pml[comp][part].psi_grid; }
int oracle_n_poles(OracleSim* s) { return s->npoles; }
