// TEST INFRASTRUCTURE ONLY (oracle/): Wigner 3j symbol by the Racah formula, standing in for GSL's
// gsl_sf_coupling_3j (arguments are twice the angular momenta / projections).
#include <cmath>
#include <algorithm>
#include <gsl/gsl_sf.h>

static double lfact(int n) { return std::lgamma(double(n) + 1.0); }

extern "C" double gsl_sf_coupling_3j(int two_ja, int two_jb, int two_jc, int two_ma, int two_mb, int two_mc)
{
    if(two_ja < 0 || two_jb < 0 || two_jc < 0) return 0.0;
    if(two_ma + two_mb + two_mc != 0) return 0.0;
    if(std::abs(two_ma) > two_ja || std::abs(two_mb) > two_jb || std::abs(two_mc) > two_jc) return 0.0;
    if((two_ja + two_ma) % 2 || (two_jb + two_mb) % 2 || (two_jc + two_mc) % 2) return 0.0;
    if(two_jc > two_ja + two_jb || two_jc < std::abs(two_ja - two_jb)) return 0.0;
    if((two_ja + two_jb + two_jc) % 2) return 0.0;

    const int jca  = (-two_ja + two_jb + two_jc) / 2;
    const int jcb  = ( two_ja - two_jb + two_jc) / 2;
    const int jcc  = ( two_ja + two_jb - two_jc) / 2;
    const int jmma = ( two_ja - two_ma) / 2;
    const int jmmb = ( two_jb - two_mb) / 2;
    const int jmmc = ( two_jc - two_mc) / 2;
    const int jpma = ( two_ja + two_ma) / 2;
    const int jpmb = ( two_jb + two_mb) / 2;
    const int jpmc = ( two_jc + two_mc) / 2;
    const int jsum = ( two_ja + two_jb + two_jc) / 2;
    const int kmin = std::max(std::max(0, jpmb - jmmc), jmma - jpmc);
    const int kmax = std::min(std::min(jcc, jmma), jpmb);
    const double lnorm = 0.5 * (lfact(jca) + lfact(jcb) + lfact(jcc) - lfact(jsum + 1)
                               + lfact(jmma) + lfact(jmmb) + lfact(jmmc) + lfact(jpma) + lfact(jpmb) + lfact(jpmc));
    double sum = 0.0;
    for(int k = kmin; k <= kmax; ++k)
    {
        double lt = lfact(k) + lfact(jcc - k) + lfact(jmma - k) + lfact(jpmb - k) + lfact(jmmc - jpmb + k) + lfact(jpmc - jmma + k);
        double term = std::exp(lnorm - lt);
        sum += (k % 2 ? -term : term);
    }
    int phase = (two_ja - two_jb - two_mc) / 2;
    return (phase % 2 ? -sum : sum);
}
