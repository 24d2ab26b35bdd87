/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of chiML's time-stepping hot path.
 *
 * Nothing under oracle/ is product code: it is imported only by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, and only as the
 * checker.  The product (chiml_b200/) never links, imports or executes it.
 *
 * What it restates: parallelFDTDFieldBase<double>::step() (reference
 * src/FDTD_MANAGER/parallelFDTDField.hpp:1228-1303) and the arithmetic it calls, as plain loops that
 * perform the same rounded operations in the same order as the reference's BLAS-level-1 call
 * chains (compile with -ffp-contract=off).  Each function cites the reference lines it follows.
 * It consumes the same flattened inputs as the C ABI (include/chiml_gpu.h): the reference's own
 * update lists, CPML lists and pole constants.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference ITSELF: oracle/_ref/chiml_ref is the unmodified
 * reference compiled in place (oracle/Makefile), and tests/test_oracle_vs_ref.py requires
 * bit-identical fields between the two on the committed cases.
 */
#ifndef CHIML_ORACLE_H
#define CHIML_ORACLE_H

#include "../include/chiml_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OracleSim OracleSim;

OracleSim* oracle_create(const ChimlGridDesc* desc);
void oracle_destroy(OracleSim* s);
int  oracle_set_update_list(OracleSim* s, int kind, int comp, const ChimlRun* runs, size_t n);
int  oracle_set_object(OracleSim* s, int obj, int npoles, const double* alpha, const double* xi, const double* gamma,
                       int use_or_dip, const double* dip);
int  oracle_set_cpml(OracleSim* s, int comp, int part, int has_psi, const ChimlPsiParams* psi, size_t npsi,
                     const ChimlGridParams* grid, size_t ngrid);
int  oracle_add_source(OracleSim* s, int field, const int32_t loc[3], const int32_t sz[3]);
int  oracle_add_emitters(OracleSim* s, const ChimlEmitterDesc* d);
int  oracle_add_dft(OracleSim* s, int field, int group, int every, int nfreq, int npts, int stride, const ChimlDftLine* lines, size_t nlines, size_t acc_len);
int  oracle_set_periodic(OracleSim* s, int comp, const ChimlWrap* w);
int  oracle_set_object_chiral(OracleSim* s, int obj, int npoles, const double* alpha, const double* xi, const double* gamma, const double* gamma_prev);
int  oracle_set_dip_grid(OracleSim* s, int comp, int pole, const double* grid);
int  oracle_set_prev_copy(OracleSim* s, const int32_t* rows, size_t nrows);
double* oracle_chi_pole(OracleSim* s, int comp, int pole, int prev);
double* oracle_prev_field(OracleSim* s, int comp);
int  oracle_set_magnetic(OracleSim* s, int has_B, int pml_on_B);
int  oracle_set_object_magnetic(OracleSim* s, int obj, int npoles, const double* alpha, const double* xi, const double* gamma);
double* oracle_mag_pole(OracleSim* s, int comp, int pole, int prev);
int  oracle_n_mag_poles(OracleSim* s);
int  oracle_add_tfsf_surface(OracleSim* s, const ChimlTfsfSurface* t);
int  oracle_commit(OracleSim* s);
/* nthreads > 1: rows of every list are split over POSIX threads (same arithmetic per cell) */
int  oracle_step_n(OracleSim* s, int n, const double* src_amp, int nthreads);

int  oracle_step_n_dft(OracleSim* s, int n, const double* src_amp, const double* twiddles, int nthreads);
int  oracle_step_n_tfsf(OracleSim* s, int n, const double* src_amp, const double* twiddles, const double* incd, size_t incd_per_step, int nthreads);
/* complex fields (Bloch-periodic runs) as two real simulations coupled by the phase factors of the wrap copies (see chiml_oracle.c) */
int  oracle_pair_step_n(OracleSim* re, OracleSim* im, int n, const double* amp_re, const double* amp_im, const ChimlWrap* wrap, const int* has_wrap,
                        const double* k_point);
double* oracle_dft(OracleSim* s, int slot, int imag);
/* one phase of one step, for y-slab runs that exchange ghost rows between phases (see chiml_oracle.c) */
int  oracle_step_phase(OracleSim* s, int phase, const double* src_amp);
int  oracle_step_phase_dft(OracleSim* s, int phase, const double* src_amp, const double* twiddles);

/* direct pointers to the full-size logical arrays (ln[0]*ln[1]*ln[2] doubles), NULL if absent */
double* oracle_field(OracleSim* s, int field);
double* oracle_pole(OracleSim* s, int comp, int pole, int prev);
double* oracle_ordip_pole(OracleSim* s, int comp, int pole, int prev);
double* oracle_psi(OracleSim* s, int comp, int part);
int     oracle_n_poles(OracleSim* s);
/* emitter state [sys][emitter][N*N][re,im] (which: 0 rho, 1..4 derivative histories), P boxes, population series */
double* oracle_emitter_state(OracleSim* s, int slot, int which);
double* oracle_emitter_P(OracleSim* s, int slot, int comp);
size_t  oracle_population(OracleSim* s, int slot, int det, double* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
