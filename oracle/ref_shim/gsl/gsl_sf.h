/* TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <gsl/gsl_sf.h>; only the Wigner 3j symbol the
 * reference calls (ML/BasisSet.cpp:50) is declared.  Defined in oracle/ref_shim/gsl_3j.cpp with the
 * closed-form Racah formula (GSL's gsl_sf_coupling_3j takes twice the angular momenta). */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
double gsl_sf_coupling_3j(int two_ja, int two_jb, int two_jc, int two_ma, int two_mb, int two_mc);
#ifdef __cplusplus
}
#endif
