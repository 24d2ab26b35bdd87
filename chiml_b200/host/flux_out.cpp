// Flux spectra of a single-rank run from the running-DFT accumulators of the flux regions: the surface-averaged frequency-domain
// fields, the Poynting integrand E x conj(H) on every surface, its Simpson integral, and the <flux name>.dat file.  Restates
// parallelFluxDTC::combineField / getFlux / simps / simps2D (DTC/parallelFlux.hpp:325-602) for one process, with the same complex
// expressions in the same order, so that the files equal the reference's character for character when the accumulators do.
// This is post-processing after the time loop: the hot path only produces the accumulators (chiml_gpu_add_dft).
#include <array>
#include <cmath>
#include <complex>
#include <fstream>
#include <iomanip>
#include <stdexcept>
#include <string>
#include <vector>

#include "setup.hpp"

namespace chiml_host {

namespace {

// netlib zaxpy with the product written out (what the reference links against here computes the same two expressions)
inline cplx cmul(const cplx& a, const cplx& b) { return cplx(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real()); }

// parallelFluxDTC::simps (DTC/parallelFlux.hpp:548-570)
cplx simps(const cplx* integrand, int n, double d)
{
    if(n == 1) return integrand[0] * d;
    cplx result(0.0, 0.0);
    if(n % 2 == 0)
    {
        for(int ii = 0; ii < (n - 2) / 2; ii++) result += d / 6.0 * (integrand[ii * 2] + 4.0 * integrand[ii * 2 + 1] + integrand[(ii + 1) * 2]);
        for(int ii = 1; ii < (n) / 2; ii++) result += d / 6.0 * (integrand[ii * 2 - 1] + 4.0 * integrand[ii * 2] + integrand[ii * 2 + 1]);
        result += d / 4.0 * (integrand[0] + integrand[1] + integrand[n - 1] + integrand[n - 2]);
    }
    else
    {
        for(int ii = 0; ii < (n - 1) / 2; ii++) result += d / 3.0 * (integrand[ii * 2] + 4.0 * integrand[ii * 2 + 1] + integrand[(ii + 1) * 2]);
    }
    return result;
}

// a complex grid {nx, ny} stored x fastest; simps2D (:578-591)
cplx simps2D(const std::vector<cplx>& g, int nx, int ny, double dx, double dy)
{
    if(ny > 1 && nx > 1)
    {
        std::vector<cplx> result(ny, 0.0);
        for(int jj = 0; jj < ny; ++jj) result[jj] = simps(&g[(size_t)jj * nx], nx, dx);
        return simps(result.data(), (int)result.size(), dy);
    }
    else if(nx > 1) return simps(g.data(), nx, dx);
    else if(ny > 1) return simps(g.data(), ny, dy);
    return cplx(0.0, 0.0);      // a one-point surface: the reference returns nothing here
}

struct Storage { const PlanDft* d; const std::vector<double>* re; const std::vector<double>* im; };
struct Surface
{
    int dir = 0; bool plus = true;
    std::vector<Storage> role[4];       // Ej, Ek, Hj, Hk
};

} // namespace

void write_flux_files(const Inputs& IP, const SlabPlan& P, const std::vector<std::vector<double>>& re, const std::vector<std::vector<double>>& im, long nSteps,
                      const std::vector<std::vector<cplx>>* incd)
{
    if(P.grid.desc.nranks != 1) throw std::logic_error("flux files are written by single-rank runs (several slabs write their accumulators)");
    const bool twoD = P.grid.desc.ln[2] == 1;
    const bool threeD = P.grid.desc.mode == CHIML_MODE_3D;
    for(size_t ff = 0; ff < IP.fluxes_.size(); ++ff)
    {
        const FluxInput& fx = IP.fluxes_[ff];
        const int nfreq = (int)fx.freqs.size();
        std::vector<Surface> surf;
        for(size_t q = 0; q < P.dfts.size(); ++q)
        {
            const PlanDft& d = P.dfts[q];
            if(d.group != (int)ff) continue;
            if((int)surf.size() <= d.surface) surf.resize(d.surface + 1);
            surf[d.surface].dir = d.dir; surf[d.surface].plus = d.plus;
            surf[d.surface].role[d.role].push_back({&d, &re[q], &im[q]});
        }
        if(surf.empty()) continue;
        // t_step_: how often fieldIn ran -- once in the propagator's constructor on the zero fields (parallelFDTDField.cpp:836-837),
        // then after every timeInt-th step
        const long nt = nSteps / fx.timeInt + 1;
        // unit conversions of the constructor (:146-173)
        double fluxConv = fx.weight, freqConv = 1.0;
        if(fx.SI)
        {
            freqConv = SPEED_OF_LIGHT / IP.a_;
            fluxConv = IP.I0_ / IP.a_ * IP.I0_ / (EPS0() * IP.a_ * SPEED_OF_LIGHT);
        }
        // the incident flux is scaled by the area of the region's faces unless cross sections are asked for (:156-172)
        double incConvX = 1.0, incConvY = 1.0, incConvZ = 1.0;
        if(fx.SI) incConvX = incConvY = incConvZ = IP.I0_ / IP.a_ * IP.I0_ / (EPS0() * IP.a_ * SPEED_OF_LIGHT);
        if(!fx.crossSec)
        {
            auto norm1 = [&](int n, double d) { std::vector<cplx> ones(n, cplx(1.0, 0.0)); return std::real(simps(ones.data(), n, d)); };
            auto norm2 = [&](int nx, int ny, double dx, double dy) {                       // simps2DNorm (:603-611)
                std::vector<cplx> result(ny, 0.0);
                for(int jj = 0; jj < ny; ++jj) { std::vector<cplx> ones(nx, cplx(1.0, 0.0)); result[jj] = simps(ones.data(), nx, dx); }
                return std::real(simps(result.data(), (int)result.size(), dy));
            };
            const double* d = P.grid.desc.d;
            if(threeD)
            {
                incConvX *= (fx.sz[2] > 1 && fx.sz[1] > 1) ? norm2(fx.sz[1], fx.sz[2], d[1], d[2]) : 0.0;
                incConvY *= (fx.sz[0] > 1 && fx.sz[2] > 1) ? norm2(fx.sz[0], fx.sz[2], d[0], d[2]) : 0.0;
                incConvZ *= (fx.sz[0] > 1 && fx.sz[1] > 1) ? norm2(fx.sz[0], fx.sz[1], d[0], d[1]) : 0.0;
            }
            else
            {
                incConvY *= norm1(fx.sz[0], d[0]);
                incConvX *= norm1(fx.sz[1], d[1]);
                incConvZ = 0.0;
            }
        }
        freqConv /= (M_PI * 2.0);

        // ---- loadFields(-1.0) (DTC/parallelFlux.hpp:664-722, called by the constructor, parallelFlux.cpp:70): the surface fields an
        // earlier run saved (normally the same cell without the scatterer) are read into Ej_freq_ .. Hk_freq_ and scaled by -1, so that
        // combineField below ADDS this run's fields to minus the incident ones.  Same order as saveFields: per role, per surface.
        std::vector<std::vector<cplx>> loaded[4];
        std::vector<std::array<int32_t, 3>> loadedN[4];
        if(fx.load)
        {
            std::ifstream lf(fx.incdFile.c_str(), std::ios::binary | std::ios::in);
            if(!lf) throw std::logic_error("flux load: cannot open " + fx.incdFile);
            int32_t hdr[4], cnt[4];
            lf.read(reinterpret_cast<char*>(hdr), sizeof(hdr));
            lf.read(reinterpret_cast<char*>(cnt), sizeof(cnt));
            if(nfreq != hdr[0] || fx.sz[0] != hdr[1] || fx.sz[1] != hdr[2] || fx.sz[2] != hdr[3])
                throw std::logic_error("Given incident fields do not match size and frequency numbers for the current calculations");
            for(int r = 0; r < 4; ++r)
                if(cnt[r] != (int32_t)surf.size()) throw std::logic_error("The size of the field vectors of the saved fields are not the same as those in the current calculation");
            for(int r = 0; r < 4; ++r)
                for(size_t vv = 0; vv < surf.size(); ++vv)
                {
                    std::array<int32_t, 3> nv;
                    lf.read(reinterpret_cast<char*>(nv.data()), sizeof(int32_t) * 3);
                    if(!lf || nv[0] < 0 || nv[1] < 0 || nv[2] < 0) throw std::logic_error("flux load: " + fx.incdFile + " is truncated");
                    std::vector<cplx> v((size_t)nv[0] * nv[1] * nv[2]);
                    lf.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * sizeof(cplx)));
                    if(!lf) throw std::logic_error("flux load: " + fx.incdFile + " is truncated");
                    for(cplx& c : v) c = cmul(cplx(-1.0, 0.0), c);             // zscal_(size, weight = -1.0)
                    loaded[r].push_back(std::move(v)); loadedN[r].push_back(nv);
                }
        }

        // ---- combineField: the stored fields of a surface averaged onto the face centres, per role a grid {nfreq, sz[tc1], sz[cor]}
        struct Face { int nx = 0, ny = 0; double dx = 0, dy = 0; std::vector<cplx> g[4]; bool has[4] = {false, false, false, false}; double weight = 1.0; };
        std::vector<Face> faces(surf.size());
        for(size_t vv = 0; vv < surf.size(); ++vv)
        {
            const Surface& s = surf[vv];
            int cor, tc1;
            std::array<int, 3> sz = fx.sz;
            if(s.dir == 0) { if(twoD) { cor = 1; tc1 = 2; } else { cor = 2; tc1 = 1; } sz[0] = 1; }
            else if(s.dir == 1) { cor = 0; tc1 = 2; sz[1] = 1; }
            else { cor = 0; tc1 = 1; sz[2] = 1; }
            const int corJ = (s.dir + 1) % 3, corK = (s.dir + 2) % 3;
            Face& F = faces[vv];
            F.nx = sz[cor]; F.ny = sz[tc1]; F.dx = P.grid.desc.d[cor]; F.dy = P.grid.desc.d[tc1];
            F.weight = s.plus ? 1.0 : -1.0;
            const int addIndex[2] = {0, 0};                      // one process: the surface starts in this process
            for(int r = 0; r < 4; ++r)
            {
                const std::vector<Storage>& arr = s.role[r];
                if(arr.empty()) continue;
                F.has[r] = true;
                F.g[r].assign((size_t)nfreq * F.nx * F.ny, cplx(0.0, 0.0));
                if(fx.load)
                {
                    static const char* roleName[4] = {"Ej_freq_", "Ek_freq_", "Hj_freq_", "Hk_freq_"};
                    const std::array<int32_t, 3>& nv = loadedN[r][vv];
                    if(nv[0] != nfreq || nv[1] != F.ny || nv[2] != F.nx)
                        throw std::logic_error(std::string("When loading in a ") + roleName[r] + " field in a flux region, one of the size elements did not agree with the file.");
                    F.g[r] = loaded[r][vv];
                }
                const int corJK = (r == 0 || r == 3) ? corJ : corK;                 // Ej, Hk: corJ; Ek, Hj: corK
                for(const Storage& st : arr)
                {
                    const int outY = st.d->nlines, outZ = st.d->npts;
                    // constructSzProcOffsetLists (:176-242): two shifted windows in 3-D (both in-plane fields exist), one otherwise
                    std::vector<std::array<int, 9>> ent(threeD ? 2 : 1, std::array<int, 9>{{0, outY, outZ, 0, 0, 0, 0, 0, 0}});
                    if(threeD)
                    {
                        const int end = st.d->gloc[corJK] + st.d->gsz[corJK] - 1;
                        const bool inside = (0 <= end) && (end < P.grid.n_global[corJK]);      // procLoc = 0, ln_vec - 2 = points of the grid
                        if(cor == corJK)
                        {
                            if(addIndex[1] == 0) { ent[1][6] = 1; ent[1][4] = 1; } else ent[0][8] = 1;
                            if(inside) ent[0][4] = 1;
                        }
                        else if(tc1 == corJK)
                        {
                            if(addIndex[0] == 0) { ent[1][5] = 1; ent[1][3] = 1; } else ent[0][7] = 1;
                            if(inside) ent[0][3] = 1;
                        }
                    }
                    const cplx w(1.0 / (static_cast<double>(arr.size() * ent.size())), 0.0);
                    for(const auto& e : ent)
                        for(int jj = 0; jj < e[2] - e[4]; ++jj)
                            for(int ii = 0; ii < e[1] - e[3]; ++ii)
                            {
                                const size_t src = (size_t)nfreq * ((size_t)(jj + e[6]) + (size_t)(ii + e[5]) * outZ);
                                const size_t dst = (size_t)nfreq * ((size_t)(jj + addIndex[1] + e[8]) + (size_t)(ii + addIndex[0] + e[7]) * F.nx);
                                if(dst + nfreq > F.g[r].size() || src + nfreq > st.re->size()) throw std::logic_error("flux output: a stored field does not fit its surface");
                                for(int f = 0; f < nfreq; ++f)
                                    F.g[r][dst + f] = F.g[r][dst + f] + cmul(w, cplx((*st.re)[src + f], (*st.im)[src + f]));
                            }
                }
            }
        }

        // ---- getFlux (:406-540)
        std::ofstream f((fx.name + ".dat").c_str());
        f << "#" << std::setw(16) << "freq\tabs(incd)\treal(incd)\timag(incd)";
        if(faces.size() > 1)
        {
            f << std::setw(16) << "\tabs(right)\treal(right)\timag(right)\tabs(left)\treal(left)\timag(left)\tabs(top)\treal(top)\timag(top)\tabs(bot)\treal(bot)\timag(bot)";
            if(fx.sz[0] > 1 && fx.sz[1] > 1 && fx.sz[2] > 1) f << std::setw(16) << "\tabs(front)\treal(front)\timag(front)\tabs(back)\treal(back)\timag(back)";
        }
        f << std::setw(16) << "freq\tabs(total)\treal(total)\timag(total)\n";
        for(int kf = 0; kf < nfreq; ++kf)
        {
            cplx flux(0.0, 0.0);
            // Fourier transform of the incident-field series of a TFSF source (:455-479): even entries (the field at the origin), then odd
            // ones (its Yee-offset partner), each weighted 1/2; real fields keep the real part (getIncdField_).  Without a TFSF source the
            // series are all zero.
            cplx inc[6];
            for(int s6 = 0; s6 < 6; ++s6)
            {
                inc[s6] = cplx(0.0, 0.0);
                if(!incd || incd->size() != 6) continue;
                const std::vector<cplx>& v = (*incd)[s6];
                for(size_t tt = 0; tt < v.size(); tt += 2)
                    inc[s6] += cplx(std::real(0.5 * v[tt])) * std::exp(cplx(0.0, fx.freqs[kf] * static_cast<double>(tt / 2) * (IP.dt_ / static_cast<double>(fx.timeInt))));
                for(size_t tt = 1; tt < v.size(); tt += 2)
                    inc[s6] += cplx(std::real(0.5 * v[tt])) * std::exp(cplx(0.0, fx.freqs[kf] * static_cast<double>(tt / 2) * (IP.dt_ / static_cast<double>(fx.timeInt))));
            }
            const cplx Ex_inc = inc[0], Ey_inc = inc[1], Ez_inc = inc[2], Hx_inc = inc[3], Hy_inc = inc[4], Hz_inc = inc[5];
            cplx flux_incd = std::pow(incConvX * (Ey_inc * std::conj(Hz_inc) - Ez_inc * std::conj(Hy_inc)) / (std::pow(static_cast<double>(nt * fx.timeInt), 2.0)), 2.0);
            flux_incd += std::pow(incConvY * (Ez_inc * std::conj(Hx_inc) - Ex_inc * std::conj(Hz_inc)) / (std::pow(static_cast<double>(nt * fx.timeInt), 2.0)), 2.0);
            flux_incd += std::pow(incConvZ * (Ex_inc * std::conj(Hy_inc) - Ey_inc * std::conj(Hx_inc)) / (std::pow(static_cast<double>(nt * fx.timeInt), 2.0)), 2.0);
            flux_incd = std::sqrt(flux_incd);
            std::vector<std::vector<cplx>> ijk(faces.size());
            for(size_t vv = 0; vv < faces.size(); ++vv)
            {
                const Face& F = faces[vv];
                const size_t n = (size_t)F.nx * F.ny;
                ijk[vv].assign(n, cplx(0.0, 0.0));
                std::vector<cplx> ikj(n, cplx(0.0, 0.0));
                if(F.has[0])          // Ej conj(Hk)
                    for(size_t e = 0; e < n; ++e)
                        ijk[vv][e] = F.g[0][kf + (size_t)nfreq * e] * std::conj(F.g[3][kf + (size_t)nfreq * e]) / std::pow(static_cast<double>(nt), 2.0);
                if(F.has[1])          // minus Ek conj(Hj)
                    for(size_t e = 0; e < n; ++e)
                    {
                        ikj[e] = F.g[1][kf + (size_t)nfreq * e] * std::conj(F.g[2][kf + (size_t)nfreq * e]) / std::pow(static_cast<double>(nt), 2.0);
                        ijk[vv][e] = ijk[vv][e] + cmul(cplx(-1.0, 0.0), ikj[e]);
                    }
            }
            f << std::setw(18) << std::setprecision(15) << freqConv * fx.freqs[kf] << "\t" << std::setw(16) << std::setprecision(15) << std::abs(flux_incd) << "\t"
              << std::setw(16) << std::setprecision(15) << std::abs(std::real(flux_incd)) << "\t" << std::setw(16) << std::setprecision(15) << std::imag(flux_incd) << "\t";
            for(size_t vv = 0; vv < faces.size(); ++vv)
            {
                const Face& F = faces[vv];
                const cplx tempFlux = fluxConv * F.weight * simps2D(ijk[vv], F.nx, F.ny, F.dx, F.dy);
                flux += tempFlux;
                f << std::setw(16) << std::setprecision(15) << std::abs(tempFlux) << "\t" << std::setw(16) << std::setprecision(15) << std::real(tempFlux) << "\t"
                  << std::setw(16) << std::setprecision(15) << std::imag(tempFlux) << "\t";
            }
            if(faces.size() > 1)
                f << std::setw(16) << std::setprecision(15) << std::abs(flux) << "\t" << std::setw(16) << std::setprecision(15) << std::real(flux) << "\t"
                  << std::setw(16) << std::setprecision(15) << std::imag(flux) << std::endl;
            else
                f << std::endl;
        }
        // ---- saveFields (DTC/parallelFlux.hpp:616-659): nfreq, the region's size, the number of surfaces per role, then per role and
        // surface the grid extents {nfreq, sz[transCor1], sz[cor]} and its complex values.  The reference walks every entry of
        // Ej_freq_ .. Hk_freq_, also the null ones of 2-D grids (a TE / TM surface stores one E and one H field): it can only save 3-D
        // regions, and so does this.
        if(fx.save)
        {
            for(const Face& F : faces)
                for(int r = 0; r < 4; ++r)
                    if(!F.has[r]) throw std::logic_error("flux save: the reference dereferences the missing in-plane field of a 2-D surface here; 3-D flux regions only");
            std::ofstream sf((fx.name + "_fields.dat").c_str(), std::ios::binary | std::ios::out);
            const int32_t hdr[4] = {nfreq, fx.sz[0], fx.sz[1], fx.sz[2]};
            sf.write(reinterpret_cast<const char*>(hdr), sizeof(hdr));
            const int32_t cnt[4] = {(int32_t)faces.size(), (int32_t)faces.size(), (int32_t)faces.size(), (int32_t)faces.size()};
            sf.write(reinterpret_cast<const char*>(cnt), sizeof(cnt));
            for(int r = 0; r < 4; ++r)
                for(const Face& F : faces)
                {
                    const int32_t nv[3] = {nfreq, F.ny, F.nx};
                    sf.write(reinterpret_cast<const char*>(nv), sizeof(nv));
                    sf.write(reinterpret_cast<const char*>(F.g[r].data()), (std::streamsize)(F.g[r].size() * sizeof(cplx)));
                }
        }
    }
}


// ---------------------------------------------------------------------------------------------------
// Frequency detectors (dtc_class "freq"): parallelDetectorFREQ_Base::collectFreqFields / fieldTranspose / toFile(incd, dt) / toMap(incd, dt)
// (DTC/parallelDTC_FREQ.hpp:275-343, 384-447, 497-583) for one process, as main.cpp:74-107 calls them (the variants with incident fields,
// whose series are all zero without a TFSF source: E_incd_ / H_incd_ hold 2 (nSteps + 1) zeros, FDTD_MANAGER/parallelFDTDField.hpp:250-255).
// ---------------------------------------------------------------------------------------------------
namespace {
// netlib zdotc with x = (1, 0) everywhere, as fieldOutFreqFunction / pwrOutFreqFunction call it (DTC/parallelDTCOutputFxn.cpp:9-18)
cplx sum_ones(int n, const cplx* y)
{
    cplx sacc(0.0, 0.0);
    const cplx one(1.0, 0.0);
    for(int i = 0; i < n; ++i) sacc = sacc + cmul(std::conj(one), y[i]);
    return sacc;
}
cplx field_out(int szFreq, const cplx* in, int, cplx*, int nt, double conv) { return conv * sum_ones(szFreq, in) / std::pow(static_cast<double>(nt), 1.0); }
cplx power_out(int szFreq, const cplx* in, int szPwr, cplx* pwr, int nt, double conv)
{
    for(int i = 0; i < szPwr; ++i) pwr[i] = in[i] * std::conj(in[i]);
    return conv * sum_ones(szFreq, pwr) / std::pow(static_cast<double>(nt), 2.0);
}
} // namespace

void write_freq_detector_files(const Inputs& IP, const SlabPlan& P, const std::vector<std::vector<double>>& re, const std::vector<std::vector<double>>& im, long nSteps)
{
    if(IP.freqDtcs_.empty()) return;
    if(P.grid.desc.nranks != 1) throw std::logic_error("frequency-detector files are written by single-rank runs");
    for(size_t k = 0; k < IP.freqDtcs_.size(); ++k)
    {
        const FreqDtcInput& q = IP.freqDtcs_[k];
        const int nfreq = (int)q.freqs.size();
        const int szFreq = q.sz[0] * q.sz[1] * q.sz[2];
        // the constructor's unit factors (:96-124)
        double freqConv = 1.0, convFactor = 1.0;
        if(q.SI)
        {
            convFactor = IP.I0_ / IP.a_;
            if(q.type == DTCTYPE::EX || q.type == DTCTYPE::EY || q.type == DTCTYPE::EZ) convFactor /= EPS0() * SPEED_OF_LIGHT;
            freqConv = SPEED_OF_LIGHT / IP.a_;
        }
        const bool pow = q.type == DTCTYPE::EPOW || q.type == DTCTYPE::HPOW;
        if(pow) convFactor = std::pow(convFactor, 2.0);
        freqConv /= (M_PI * 2.0);
        auto toOut = pow ? power_out : field_out;
        const int tStep = (int)(nSteps / q.timeInt + 1);        // output(): once in the propagator's constructor, then every timeInt-th step
        // collectFreqFields + fieldTranspose: per stored field a {szFreq, nfreq} grid, point j of frequency f at j + szFreq * f.  The box
        // point j = x + sz_x (z + sz_z y) is accumulator (line z + sz_z y, point x) of the volume storage; a 2-D grid stores no line at all
        // (getLocalSzEl of the absent z axis is 0) and its fields stay zero
        std::vector<std::vector<cplx>> trans;
        int nIncd = 0;
        for(size_t e = 0; e < P.dfts.size(); ++e)
        {
            const PlanDft& d = P.dfts[e];
            if(d.freq_dtc != (int)k) continue;
            std::vector<cplx> t((size_t)szFreq * nfreq, cplx(0.0, 0.0));
            const size_t pts = d.acc_len / (size_t)std::max(nfreq, 1);
            if(pts != 0 && pts != (size_t)szFreq) throw std::logic_error("frequency detector: the stored box does not match the detector's size");
            for(size_t j = 0; j < pts; ++j)
                for(int f = 0; f < nfreq; ++f) t[j + (size_t)szFreq * f] = cplx(re[e][f + (size_t)nfreq * j], im[e][f + (size_t)nfreq * j]);
            trans.push_back(std::move(t));
        }
        // which incident series main.cpp hands over: three for the power types, one otherwise
        nIncd = pow ? 3 : 1;
        const size_t incdLen = 2 * ((size_t)P.grid.n_steps + 1);
        auto incd_point = [&](int ii) {
            cplx pt(0.0, 0.0);
            const cplx zero(0.0, 0.0);
            for(size_t tt = 0; tt < incdLen; tt += 2)
                pt += std::real(0.5 * zero) * std::exp(cplx(0.0, -1.0 * q.freqs[ii] * static_cast<double>(tt / 2) * (IP.dt_ / static_cast<double>(q.timeInt))));
            for(size_t tt = 1; tt < incdLen; tt += 2)
                pt += std::real(0.5 * zero) * std::exp(cplx(0.0, -1.0 * q.freqs[ii] * static_cast<double>(tt / 2) * (IP.dt_ / static_cast<double>(q.timeInt))));
            pt /= static_cast<double>(tStep * q.timeInt);
            if(pow) pt *= std::conj(pt);
            pt *= convFactor / 2.0;
            return pt;
        };
        if(!q.outputMaps)
        {
            std::ofstream f(q.name.c_str());
            f << "#" << std::setw(16) << "freq\tabs(incd)\treal(incd)\timag(incd)";
            for(size_t nn = 1; nn < trans.size(); ++nn) f << "\tabs(field " << nn << ")\treal(field " << nn << ")\timag(field " << nn << ")";
            if(trans.size() > 1) f << "\tabs(total)\treal(total)\timag(total)";
            f << std::endl;
            std::vector<cplx> pwr(szFreq, 0.0);
            for(int ii = 0; ii < nfreq; ii++)
            {
                cplx incd_field(0.0, 0.0);
                f << std::setw(16) << std::setprecision(12) << freqConv * q.freqs[ii];
                for(int v = 0; v < nIncd; ++v)
                {
                    const cplx pt = incd_point(ii);
                    f << "\t" << std::setw(16) << std::setprecision(12) << std::real(pt) << "\t" << std::setw(16) << std::setprecision(12) << std::imag(pt) << "\t"
                      << std::setw(16) << std::setprecision(12) << std::abs(pt);
                    incd_field += pt;
                }
                if(nIncd > 2)
                    f << "\t" << std::setw(16) << std::setprecision(12) << std::real(incd_field) << "\t" << std::setw(16) << std::setprecision(12) << std::imag(incd_field) << "\t"
                      << std::setw(16) << std::setprecision(12) << std::abs(incd_field);
                cplx freq = 0;
                for(auto& grid : trans)
                {
                    const cplx pt = toOut(szFreq, &grid[(size_t)szFreq * ii], (int)pwr.size(), pwr.data(), tStep, convFactor / 2.0);
                    f << "\t" << std::setw(16) << std::setprecision(12) << std::real(pt) << "\t" << std::setw(16) << std::setprecision(12) << std::imag(pt) << "\t"
                      << std::setw(16) << std::setprecision(12) << std::abs(pt);
                    freq += pt;
                }
                if(trans.size() > 1)
                    f << '\t' << std::setw(16) << std::setprecision(12) << std::real(freq) << "\t" << std::setw(16) << std::setprecision(12) << std::imag(freq) << "\t"
                      << std::setw(16) << std::setprecision(12) << std::abs(freq) << std::endl;
                else
                    f << "\n";
            }
        }
        else
        {
            for(int ii = 0; ii < nfreq; ii++)
            {
                std::vector<cplx> pwr(1, 0.0);
                const std::string fileName = q.name + "." + std::to_string(freqConv * q.freqs[ii]);
                std::ofstream f(fileName.c_str());
                f << "#" << std::setw(16) << "x\ty\tz\tabs(incd)\treal(incd)\timag(incd)";
                for(size_t nn = 1; nn < trans.size(); ++nn) f << "\tabs(field " << nn << ")\treal(field " << nn << ")\timag(field " << nn << ")";
                if(trans.size() > 1) f << "\tabs(total)\treal(total)\timag(total)";
                f << std::endl;
                cplx incd_field(0.0, 0.0);
                std::vector<double> outIncdToFile;
                for(int v = 0; v < nIncd; ++v)
                {
                    const cplx pt = incd_point(ii);
                    outIncdToFile.push_back(std::real(pt)); outIncdToFile.push_back(std::imag(pt)); outIncdToFile.push_back(std::abs(pt));
                    incd_field += pt;
                }
                if(nIncd > 2) { outIncdToFile.push_back(std::real(incd_field)); outIncdToFile.push_back(std::imag(incd_field)); outIncdToFile.push_back(std::abs(incd_field)); }
                for(int yy = 0; yy < q.sz[1]; ++yy)
                    for(int zz = 0; zz < q.sz[2]; ++zz)
                        for(int xx = 0; xx < q.sz[0]; ++xx)
                        {
                            f << std::setw(16) << xx << '\t' << yy << '\t' << zz;
                            for(double val : outIncdToFile) f << "\t" << std::setw(16) << std::setprecision(12) << val;
                            const int jj = xx + zz * q.sz[0] + yy * q.sz[0] * q.sz[2];
                            cplx freq = 0;
                            for(auto& grid : trans)
                            {
                                const cplx pt = toOut(1, &grid[(size_t)jj + (size_t)szFreq * ii], (int)pwr.size(), pwr.data(), tStep, convFactor / 2.0);
                                f << "\t" << std::setw(16) << std::setprecision(12) << std::abs(pt) << "\t" << std::real(pt) << "\t" << std::imag(pt);
                                freq += pt;
                            }
                            if(trans.size() > 1)
                                f << '\t' << std::setw(16) << std::setprecision(12) << std::real(freq) << "\t" << std::setw(16) << std::setprecision(12) << std::imag(freq) << "\t"
                                  << std::setw(16) << std::setprecision(12) << std::abs(freq) << std::endl;
                            else
                                f << "\n";
                        }
            }
        }
    }
}

} // namespace chiml_host
