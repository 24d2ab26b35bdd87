// See setup.hpp.  Reference lines are cited at each piece of logic; indices are the reference's
// (local, ghost-inclusive, x + lnx*(z + lnz*y)).
#include "setup.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <thread>
#include <unordered_map>

namespace chiml_host {

namespace {

// ---------------------------------------------------------------------------------------------------
// slab geometry
// ---------------------------------------------------------------------------------------------------
struct Geom
{
    int rank, nranks;
    bool last;             // rank == size-1
    bool twoD;
    int n[3];              // n_vec_: grid points per direction (2-D: n[2] = 1)
    int ln[3];             // local ghost-inclusive extents
    int yStart;            // procLoc(1)
    double d[3], dt;
    int thick[3];          // pmlThickness_
    int mn[3], pl[3];      // local CPML thicknesses (parallelCPML::findLnVecs, PML/parallelPML.hpp:280-308)
    int ind(int x, int y, int z) const { return x + ln[0] * (z + y * ln[2]); }
    int procLoc(int dir) const { return dir == 1 ? yStart : 0; }
};

// the eight staggered maps of the constructor (parallelFDTDField.hpp:264-272): offset of the sample point
// inside the cell and which directions are one point short
struct GridSpec { double off[3]; int endOff[3]; bool E; };
const GridSpec SPEC_NODE_P = {{0.0, 0.0, 0.0}, {0, 0, 0}, true};
const GridSpec SPEC_COMP[6] = {
    {{0.5, 0.0, 0.0}, {1, 0, 0}, true},  {{0.0, 0.5, 0.0}, {0, 1, 0}, true},  {{0.0, 0.0, 0.5}, {0, 0, 1}, true},
    {{0.0, 0.5, 0.5}, {0, 1, 1}, false}, {{0.5, 0.0, 0.5}, {1, 0, 1}, false}, {{0.5, 0.5, 0.0}, {1, 1, 0}, false}};
// derivOff (parallelFDTDField.cpp:80-82,248-250)
const int DERIV_OFF[6][3] = {{0, -1, 0}, {0, 0, -1}, {-1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 0}};

bool comp_exists(int mode, int comp)
{
    const int c = comp % 3;
    const bool isH = comp >= 3;
    if(mode == CHIML_MODE_3D) return true;
    if(mode == CHIML_MODE_TE) return isH ? c == 2 : c != 2;
    return isH ? c != 2 : c == 2;
}

// ---------------------------------------------------------------------------------------------------
// object maps, one row at a time (replaces setupPhysFields, parallelFDTDField.hpp:867-947)
// ---------------------------------------------------------------------------------------------------
class Rasteriser
{
public:
    Rasteriser(const Inputs& IP, const Geom& g) : IP_(IP), g_(g)
    {
        for(int d = 0; d < 3; ++d) cen_[d] = (g.n[d] - g.n[d] % 2) / 2.0;
        for(size_t oo = 1; oo < IP.objArr_.size(); ++oo)
        {
            const Obj& o = *IP.objArr_[oo];
            if(!o.relevant()) continue;
            Cull c;
            c.obj = (int)oo;
            c.sameGeo = o.geoParam_ == o.geoParamML_;
            const std::array<double, 3> h = o.halfExtent(o.geoParamML_);
            for(int d = 0; d < 3; ++d)
            {
                // global index g with |(g + off - cen)*d - loc| <= h + margin, off in [0, 0.5]
                const double marg = 2.0 * g.d[d];
                c.lo[d] = std::isfinite(h[d]) ? (long)std::floor((o.location_[d] - h[d] - marg) / g.d[d] + cen_[d]) - 1 : std::numeric_limits<long>::min() / 2;
                c.hi[d] = std::isfinite(h[d]) ? (long)std::ceil((o.location_[d] + h[d] + marg) / g.d[d] + cen_[d]) + 1 : std::numeric_limits<long>::max() / 2;
            }
            cull_.push_back(c);
        }
    }

    // object id of the last relevant object containing the sample point at GLOBAL indices (gx, gy, gz); 0 = background
    int idAt(const GridSpec& s, long gx, long gy, long gz) const
    {
        int id = 0;
        const std::array<double, 3> pt = {{(gx + s.off[0] - cen_[0]) * g_.d[0], (gy + s.off[1] - cen_[1]) * g_.d[1], (gz + s.off[2] - cen_[2]) * g_.d[2]}};
        for(const Cull& c : cull_)
        {
            if(gx < c.lo[0] || gx > c.hi[0] || gy < c.lo[1] || gy > c.hi[1] || gz < c.lo[2] || gz > c.hi[2]) continue;
            const Obj& o = *IP_.objArr_[c.obj];
            if(o.isObj(pt, g_.d[0], o.geoParamML_)) id = c.obj;
        }
        return id;
    }

    // eps (or mu) map value at GLOBAL indices, without the border marks of row(): the value of the last relevant object whose
    // (non-enlarged) geometry contains the sample point, 1 elsewhere (parallelFDTDField.hpp:891-894)
    double epsAt(const GridSpec& s, long gx, long gy, long gz) const
    {
        double v = 1.0;
        const std::array<double, 3> pt = {{(gx + s.off[0] - cen_[0]) * g_.d[0], (gy + s.off[1] - cen_[1]) * g_.d[1], (gz + s.off[2] - cen_[2]) * g_.d[2]}};
        for(const Cull& c : cull_)
        {
            if(gx < c.lo[0] || gx > c.hi[0] || gy < c.lo[1] || gy > c.hi[1] || gz < c.lo[2] || gz > c.hi[2]) continue;
            const Obj& o = *IP_.objArr_[c.obj];
            if(o.isObj(pt, g_.d[0], o.geoParam_)) v = s.E ? o.eps_infty_ : o.mu_infty_;
        }
        return v;
    }

    // one local row (jj, kk) of the object-id and eps/mu maps, including the -1 / 0.0 border marks
    void row(const GridSpec& s, int jj, int kk, int* obj, double* eps) const
    {
        const int lx = g_.ln[0], ly = g_.ln[1], lz = g_.ln[2];
        std::fill(obj, obj + lx, 0);
        std::fill(eps, eps + lx, 1.0);
        const int zmin = g_.twoD ? 0 : 1, zmax = g_.twoD ? 1 : lz - 1;
        if(jj >= 1 && jj < ly - 1 && kk >= zmin && kk < zmax)
        {
            const long gy = (long)(jj - 1) + g_.yStart, gz = (long)(kk - 1);
            const double py = (gy + s.off[1] - cen_[1]) * g_.d[1], pz = (gz + s.off[2] - cen_[2]) * g_.d[2];
            for(const Cull& c : cull_)
            {
                if(gy < c.lo[1] || gy > c.hi[1] || gz < c.lo[2] || gz > c.hi[2]) continue;
                const Obj& o = *IP_.objArr_[c.obj];
                const int x0 = (int)std::max<long>(1, c.lo[0] + 1), x1 = (int)std::min<long>(lx - 2, c.hi[0] + 1);
                const double val = s.E ? o.eps_infty_ : o.mu_infty_;
                for(int ii = x0; ii <= x1; ++ii)
                {
                    const std::array<double, 3> pt = {{((ii - 1) + s.off[0] - cen_[0]) * g_.d[0], py, pz}};
                    const bool inML = o.isObj(pt, g_.d[0], o.geoParamML_);
                    if(inML) obj[ii] = c.obj;
                    if(c.sameGeo ? inML : o.isObj(pt, g_.d[0], o.geoParam_)) eps[ii] = val;
                }
            }
        }
        // borders between processes / domain edge are marked -1; the short last line additionally gets eps = 0 (:900-945)
        if(!g_.twoD)
        {
            if(kk == 0 || kk == lz - 1) std::fill(obj, obj + lx, -1);
            if(s.endOff[2] == 1 && kk == lz - 2) { std::fill(obj, obj + lx, -1); std::fill(eps, eps + lx, 0.0); }
        }
        if(jj == 0 || jj == ly - 1) std::fill(obj, obj + lx, -1);
        if(s.endOff[1] == 1 && g_.last && g_.n[1] + 2 * g_.nranks > 3 && jj == ly - 2) { std::fill(obj, obj + lx, -1); std::fill(eps, eps + lx, 0.0); }
        obj[0] = -1; obj[lx - 1] = -1;
        if(s.endOff[0] == 1 && g_.n[0] + 2 > 3) { obj[lx - 2] = -1; eps[lx - 2] = 0.0; }
    }

private:
    struct Cull { int obj; bool sameGeo; long lo[3], hi[3]; };
    const Inputs& IP_;
    const Geom& g_;
    double cen_[3];
    std::vector<Cull> cull_;
};

// ---------------------------------------------------------------------------------------------------
// does material reach into the CPML?  (parallelFDTDField.hpp:275-370, including its loop-bound quirk:
// min/max of the PML-normal index are overwritten inside the (jj,kk) loops, so the lower slab is
// scanned only on the first transverse line and the upper slab on every line)
// ---------------------------------------------------------------------------------------------------
void materials_in_pml(const Inputs& IP, const int n_vec[3], const double d[3], bool& dielectricMatInPML, bool& magMatInPML)
{
    dielectricMatInPML = false; magMatInPML = false;
    // a flag nobody can set is settled from the start: the scan stops once both are
    bool anyDie = false, anyMag = false;
    for(const auto& obj : IP.objArr_)
    {
        if(obj->eps_infty_ == 1.0 && obj->mu_infty_ == 1.0 && obj->gamma_.size() < 1 && obj->magGamma_.size() < 1 && obj->chiGamma_.size() < 1 && !obj->ML_) continue;
        anyDie = anyDie || !(obj->eps_infty_ == 1.0 && obj->gamma_.size() < 1 && obj->chiGamma_.size() < 1);
        anyMag = anyMag || !(obj->mu_infty_ == 1.0 && obj->magGamma_.size() < 1 && obj->chiGamma_.size() < 1);
    }
    auto settled = [&]() { return (dielectricMatInPML || !anyDie) && (magMatInPML || !anyMag); };
    for(int pp = 0; pp < 3 && !settled(); ++pp)
    {
        const int cor_ii = pp, cor_jj = (pp + 1) % 3, cor_kk = (pp + 2) % 3;
        int mn[3], mx[3];
        for(int k = 0; k < 3; ++k) { mn[k] = (int)(-1 * (double)n_vec[k] / 2); mx[k] = (int)((double)n_vec[k] / 2); }
        mx[cor_ii] = (int)(IP.pmlThickness_[cor_ii] - n_vec[cor_ii] / 2.0);
        for(const auto& obj : IP.objArr_)
        {
            if(obj->eps_infty_ == 1.0 && obj->mu_infty_ == 1.0 && obj->gamma_.size() < 1 && obj->magGamma_.size() < 1 && obj->chiGamma_.size() < 1 && !obj->ML_) continue;
            const bool dielcMat = !(obj->eps_infty_ == 1.0 && obj->gamma_.size() < 1 && obj->chiGamma_.size() < 1);
            const bool magMat = !(obj->mu_infty_ == 1.0 && obj->magGamma_.size() < 1 && obj->chiGamma_.size() < 1);
            // bounding box of the object in the centred integer coordinates of this scan, for culling
            const std::array<double, 3> h = obj->halfExtent(obj->geoParam_);
            long lo[3], hi[3];
            for(int k = 0; k < 3; ++k)
            {
                lo[k] = std::isfinite(h[k]) ? (long)std::floor((obj->location_[k] - h[k]) / d[k]) - 2 : std::numeric_limits<long>::min() / 2;
                hi[k] = std::isfinite(h[k]) ? (long)std::ceil((obj->location_[k] + h[k]) / d[k]) + 2 : std::numeric_limits<long>::max() / 2;
            }
            auto scan = [&](int jj, int kk, int i0, int i1) {
                const bool wantDie = dielcMat && !dielectricMatInPML, wantMag = magMat && !magMatInPML;
                if(!wantDie && !wantMag) return;
                if(jj < lo[cor_jj] || jj > hi[cor_jj] || kk < lo[cor_kk] || kk > hi[cor_kk]) return;
                for(int ii = (int)std::max<long>(i0, lo[cor_ii]); ii < i1 && ii <= hi[cor_ii]; ++ii)
                {
                    if(!(dielcMat && !dielectricMatInPML) && !(magMat && !magMatInPML)) break;
                    std::array<double, 3> pt = {{0, 0, 0}};
                    pt[cor_ii] = static_cast<double>(ii) * d[cor_ii];
                    pt[cor_jj] = static_cast<double>(jj) * d[cor_jj];
                    pt[cor_kk] = static_cast<double>(kk) * d[cor_kk];
                    // the six Yee points of the cell, reached by the reference's chain of offsets (including its d_[0] / d_[2] slips)
                    pt[0] += d[0] / 2.0;                                        // Ex point
                    if(!dielectricMatInPML && dielcMat && obj->isObj(pt, d[0], obj->geoParam_)) dielectricMatInPML = true;
                    pt[1] += d[1] / 2.0;                                        // Hz point
                    if(!magMatInPML && magMat && obj->isObj(pt, d[0], obj->geoParam_)) magMatInPML = true;
                    pt[0] -= d[0] / 2.0;                                        // Ey point
                    if(!dielectricMatInPML && dielcMat && obj->isObj(pt, d[0], obj->geoParam_)) dielectricMatInPML = true;
                    pt[2] += d[2] / 2.0;                                        // Hx point
                    if(!magMatInPML && magMat && obj->isObj(pt, d[0], obj->geoParam_)) magMatInPML = true;
                    pt[1] -= d[0] / 2.0;                                        // Ez point (the reference subtracts d_[0] here)
                    if(!dielectricMatInPML && dielcMat && obj->isObj(pt, d[0], obj->geoParam_)) dielectricMatInPML = true;
                    pt[0] += d[2] / 2.0;                                        // Hy point (the reference adds d_[2] here)
                    if(!magMatInPML && magMat && obj->isObj(pt, d[0], obj->geoParam_)) magMatInPML = true;
                }
            };
            for(int jj = mn[cor_jj]; jj < mx[cor_jj]; ++jj)
                for(int kk = mn[cor_kk]; kk < mx[cor_kk]; ++kk)
                {
                    scan(jj, kk, mn[cor_ii], mx[cor_ii]);
                    mn[cor_ii] = n_vec[cor_ii] / 2 - IP.pmlThickness_[cor_ii];
                    mx[cor_ii] = n_vec[cor_ii] / 2;
                    scan(jj, kk, mn[cor_ii], mx[cor_ii]);
                }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// update lists of one map (initializeList / getBlasLists / fillBlasLists / populateUpLists,
// parallelFDTDField.hpp:628-682,751-856)
// ---------------------------------------------------------------------------------------------------
struct Box { int mn[3], mx[3]; bool includeU; bool curl; };
struct RawRun { int x, y, z, n, obj; double eps; };

struct ListSet { std::vector<ChimlRun> U, D, LorD, OrDipD, ChiD; std::vector<std::array<int, 5>> chiLocs; /* {n, x, y, z, obj} of the ChiD runs */ };

// dieInPML / magInPML: dielectricMatInPML_ / magMatInPML_
void build_lists(const Inputs& IP, const Geom& g, const Rasteriser& ras, const GridSpec& spec, bool E, const int derivOff[3], const int fieldEnd[3],
                 double dj, double dk, bool dieInPML, bool magInPML, bool orDipField, int nthreads, ListSet& out)
{
    const int lx = g.ln[0];
    int mn[3] = {1, 1, g.twoD ? 0 : 1};
    int mx[3] = {g.ln[0] - 1 - fieldEnd[0], g.ln[1] - 1, g.twoD ? 1 : g.ln[2] - 1 - fieldEnd[2]};
    if(g.last) mx[1] -= fieldEnd[1];
    // getBlasLists (:802-856)
    const bool matInPML = dieInPML || magInPML;
    const bool inc = (E && dieInPML) || (!E && magInPML);
    int PML_x_left = inc ? g.mn[0] : 0, PML_x_right = inc ? g.pl[0] : 0;
    if(PML_x_right != 0) PML_x_right -= fieldEnd[0];
    int PML_y_bot = inc ? g.mn[1] : 0, PML_y_top = inc ? g.pl[1] : 0;
    if(PML_y_top != 0 && g.last) PML_y_top -= fieldEnd[1];
    int PML_z_back = (!g.twoD && inc) ? g.mn[2] : 0, PML_z_front = (!g.twoD && inc) ? g.pl[2] : 0;
    if(PML_z_front != 0) PML_z_front -= fieldEnd[2];
    const bool incU = !matInPML;
    std::vector<Box> boxes = {
        {{mn[0], mn[1], mn[2]}, {mx[0], mn[1] + PML_y_bot, mx[2]}, incU, false},
        {{mn[0], mx[1] - PML_y_top, mn[2]}, {mx[0], mx[1], mx[2]}, incU, false},
        {{mn[0], mn[1] + PML_y_bot, mn[2]}, {mx[0], mx[1] - PML_y_top, mn[2] + PML_z_back}, incU, false},
        {{mn[0], mn[1] + PML_y_bot, mx[2] - PML_z_front}, {mx[0], mx[1] - PML_y_top, mx[2]}, incU, false},
        {{mn[0], mn[1] + PML_y_bot, mn[2] + PML_z_back}, {mn[0] + PML_x_left, mx[1] - PML_y_top, mx[2] - PML_z_front}, incU, false},
        {{mx[0] - PML_x_right, mn[1] + PML_y_bot, mn[2] + PML_z_back}, {mx[0], mx[1] - PML_y_top, mx[2] - PML_z_front}, incU, false},
        {{mn[0] + PML_x_left, mn[1] + PML_y_bot, mn[2] + PML_z_back}, {mx[0] - PML_x_right, mx[1] - PML_y_top, mx[2] - PML_z_front}, true, false}};
    {
        Box c; c.includeU = true; c.curl = true;
        for(int k = 0; k < 3; ++k) { c.mn[k] = mn[k]; c.mx[k] = mx[k]; }
        c.mn[0] += g.mn[0]; c.mx[0] -= g.pl[0]; if(g.pl[0] != 0) c.mx[0] += fieldEnd[0];
        c.mn[1] += g.mn[1]; c.mx[1] -= g.pl[1]; if(g.pl[1] != 0 && g.last) c.mx[1] += fieldEnd[1];
        c.mn[2] += g.twoD ? 0 : g.mn[2]; c.mx[2] -= g.twoD ? 0 : g.pl[2]; if(!g.twoD && g.pl[2] != 0) c.mx[2] += fieldEnd[2];
        boxes.push_back(c);
    }
    const int nb = (int)boxes.size();
    if(nthreads < 1) nthreads = 1;
    const int ny = std::max(0, mx[1] - mn[1]);
    nthreads = std::max(1, std::min(nthreads, ny));
    // per thread, per box, per kind: raw runs in (y, z, x) order
    enum { K_U = 0, K_D = 1, K_ORD = 2, K_CHI = 3, NK = 4 };
    std::vector<std::vector<std::vector<RawRun>>> raw(nthreads, std::vector<std::vector<RawRun>>(nb * NK));
    auto work = [&](int t) {
        std::vector<int> obj(lx);
        std::vector<double> eps(lx);
        const int y0 = mn[1] + (int)((long)ny * t / nthreads), y1 = mn[1] + (int)((long)ny * (t + 1) / nthreads);
        for(int jj = y0; jj < y1; ++jj)
            for(int kk = mn[2]; kk < mx[2]; ++kk)
            {
                bool have = false;
                for(int b = 0; b < nb; ++b)
                {
                    const Box& bx = boxes[b];
                    if(jj < bx.mn[1] || jj >= bx.mx[1] || kk < bx.mn[2] || kk >= bx.mx[2] || bx.mx[0] <= bx.mn[0]) continue;
                    if(!have) { ras.row(spec, jj, kk, obj.data(), eps.data()); have = true; }
                    int ii = bx.mn[0];
                    while(ii < bx.mx[0])
                    {
                        const int iistore = ii;
                        while((ii < bx.mx[0] - 1) && (obj[ii] == obj[ii + 1]) && (eps[ii] == eps[ii + 1])) ++ii;
                        const int id = obj[iistore];
                        const Obj& o = *IP.objArr_[id < 0 ? 0 : id];
                        int kind;
                        // fillBlasLists :767-779 (oriented magnetic / chiral dipoles are refused before this point)
                        if(o.useOrientedDipols_ && E && o.gamma_.size() > 0) kind = K_ORD;
                        else if(o.chiGamma_.size() > 0) kind = K_CHI;
                        else if(bx.includeU && ((!E && o.magGamma_.size() < 1) || (E && o.gamma_.size() < 1 && !o.ML_))) kind = K_U;
                        else kind = K_D;
                        raw[t][b * NK + kind].push_back({iistore, jj, kk, ii - iistore + 1, id, eps[iistore]});
                        ++ii;
                    }
                }
            }
    };
    if(nthreads == 1) work(0);
    else
    {
        std::vector<std::thread> th;
        for(int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
        for(auto& x : th) x.join();
    }
    // populateUpLists (:628-649)
    auto emit = [&](std::vector<ChimlRun>& dst, const RawRun& r, bool upUList) {
        ChimlRun u;
        const double ep_mu = upUList ? r.eps : 1.0;
        const int x = r.x, y = r.y, z = r.z;
        u.n = r.n; u.ind = g.ind(x, y, z); u.obj = r.obj;
        if(!orDipField)
        {
            if(g.ln[2] == 1)
            {
                u.ind_i = g.ind(x - derivOff[1], y - derivOff[2], z);
                u.ind_j = g.ind(x + derivOff[0], y + derivOff[1], z);
                u.ind_k = g.ind(x + derivOff[2], y + derivOff[0], z);
            }
            else
            {
                u.ind_i = g.ind(x - derivOff[1], y - derivOff[2], z - derivOff[0]);
                u.ind_j = g.ind(x + derivOff[0], y + derivOff[1], z + derivOff[2]);
                u.ind_k = g.ind(x + derivOff[2], y + derivOff[0], z + derivOff[1]);
            }
        }
        else
        {
            u.ind_i = g.ind(x + derivOff[0], y, z);
            u.ind_j = g.ind(x, y + derivOff[1], z);
            u.ind_k = g.ind(x, y, z + derivOff[2]);
        }
        u.pf[0] = 1.0; u.pf[1] = -1.0 * g.dt / (ep_mu * dj); u.pf[2] = -1.0 * g.dt / (ep_mu * dk); u.pf[3] = r.eps;
        dst.push_back(u);
    };
    for(int b = 0; b < nb; ++b)
        for(int t = 0; t < nthreads; ++t)
        {
            if(boxes[b].curl)
            {
                for(const RawRun& r : raw[t][b * NK + K_U]) emit(out.U, r, true);
                // the curl pass files every non-U run under upD (getBlasLists :855 passes upDLists four times)
            }
            else
            {
                for(const RawRun& r : raw[t][b * NK + K_D]) emit(out.LorD, r, false);
                for(const RawRun& r : raw[t][b * NK + K_ORD]) emit(out.OrDipD, r, false);
                for(const RawRun& r : raw[t][b * NK + K_CHI])
                {
                    emit(out.ChiD, r, false);
                    out.chiLocs.push_back({{r.n, r.x, r.y, r.z, r.obj}});       // populateUpLists storeLocs (:645-646)
                }
            }
        }
    // upD keeps the reference's order inside the curl box: runs of all non-U kinds interleaved in (y, z, x) order
    {
        const int b = nb - 1;
        for(int t = 0; t < nthreads; ++t)
        {
            const std::vector<RawRun>* rr[3] = {&raw[t][b * NK + K_D], &raw[t][b * NK + K_ORD], &raw[t][b * NK + K_CHI]};
            size_t at[3] = {0, 0, 0};
            auto before = [](const RawRun& p, const RawRun& q) { return p.y != q.y ? p.y < q.y : (p.z != q.z ? p.z < q.z : p.x < q.x); };
            for(;;)
            {
                int best = -1;
                for(int k = 0; k < 3; ++k)
                    if(at[k] < rr[k]->size() && (best < 0 || before((*rr[k])[at[k]], (*rr[best])[at[best]]))) best = k;
                if(best < 0) break;
                emit(out.D, (*rr[best])[at[best]++], false);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// CPML (parallelCPML constructor and list builders, PML/parallelPML.hpp:102-182,192-275,321-451,462-656,708-765)
// ---------------------------------------------------------------------------------------------------
struct CpmlBuilder
{
    const Inputs& IP;
    const Geom& g;
    const Rasteriser& ras;
    int comp;
    bool E;
    int i_, j_, k_;
    int sh[3];                 // one point short in direction d (fieldEnd)
    bool hasGrid[2];           // part 0 needs grid_k, part 1 needs grid_j
    std::vector<double> eta[3][2];   // [direction][0 = minus side, 1 = plus side]
    std::vector<ChimlPsiParams> psi[2];
    std::vector<ChimlGridParams> grid[2];

    double kappa(double ii, double iiMax) const { return (0.0 <= ii && ii <= iiMax) ? 1.0 + (IP.pmlKappaMax_ - 1.0) * std::pow((iiMax - ii) / iiMax, IP.pmlM_) : 1.0; }
    double sigma(double ii, double iiMax, double eta_eff, double sigmaMax) const { return (0.0 <= ii && ii <= iiMax) ? sigmaMax / eta_eff * pow((iiMax - ii) / iiMax, IP.pmlM_) : 0.0; }
    double aVal(double ii, double iiMax) const { return (0.0 <= ii && ii <= iiMax) ? IP.pmlAMax_ * std::pow(ii / iiMax, IP.pmlMa_) : 0.0; }
    double b(double sig, double a, double kap) const { return std::exp(-1.0 * (sig / kap + a) * g.dt); }
    double c(double sig, double a, double kap) const { return (sig == 0 && a == 0) ? 0 : sig * (b(sig, a, kap) - 1.0) / (kap * (sig + kap * a)); }

    // genEtaEff / genEtaEffVal (:192-275): product of the plane means of eps_inf and mu_inf of the objects on the plane
    void genEta(int dir, bool pl)
    {
        const GridSpec& spec = SPEC_COMP[comp];
        int planeSzTemp[3] = {g.n[0] - sh[0], g.n[1] - sh[1], g.n[2] - sh[2]};
        if(g.twoD) planeSzTemp[2] = 1;
        const int lo = pl ? g.n[dir] - g.thick[dir] : 0, hi = pl ? g.n[dir] : g.thick[dir];
        std::vector<double>& out = eta[dir][pl ? 1 : 0];
        for(int ii = lo; ii < hi; ++ii)
        {
            double eps_sum = 0.0, mu_sum = 0.0;
            double N;
            auto add = [&](int id) {
                // ids on the sampled planes are never -1: the short last line is excluded by planeSzTemp
                eps_sum += IP.objArr_[id]->eps_infty_ / N;
                mu_sum += IP.objArr_[id]->mu_infty_ / N;
            };
            // a 2-D map has one z layer, kk = 0, whose sample point the reference puts at z = (kk - 1) d = -d (parallelFDTDField.hpp:890)
            const long zoff = g.twoD ? -1 : 0;
            if(dir == 0)
            {
                N = static_cast<double>(planeSzTemp[2] * planeSzTemp[1]);
                for(int yy = 0; yy < planeSzTemp[1]; ++yy)
                    for(int zz = 0; zz < planeSzTemp[2]; ++zz) add(ras.idAt(spec, ii, yy, zz + zoff));
            }
            else if(dir == 1)
            {
                N = static_cast<double>(planeSzTemp[0] * planeSzTemp[2]);
                for(int zz = 0; zz < planeSzTemp[2]; ++zz)
                    for(int xx = 0; xx < planeSzTemp[0]; ++xx) add(ras.idAt(spec, xx, ii, zz + zoff));
            }
            else
            {
                N = static_cast<double>(planeSzTemp[0] * planeSzTemp[1]);
                for(int yy = 0; yy < planeSzTemp[1]; ++yy)
                    for(int xx = 0; xx < planeSzTemp[0]; ++xx) add(ras.idAt(spec, xx, yy, ii));
            }
            const double etaSum = eps_sum * mu_sum;
            const double etaVal = (etaSum != 0.00) ? etaSum : 0;
            if((int)out.size() < g.thick[dir] && etaVal != 0.0) out.push_back(etaVal);
        }
    }

    // getPsiUpList (:321-451)
    void psiList(int dir, bool pl, int startPt, int nDir, int dirMax, std::vector<ChimlPsiParams>& list)
    {
        ChimlPsiParams param;
        int transSz2 = 1, cor_norm = dir, cor_trans1, cor_trans2, trans1FieldOff = 0, trans2FieldOff = 0, ccStart = 0;
        const double sigmaMax = IP.pmlSigOptRat_ * 0.8 * (IP.pmlM_ + 1) / g.d[dir];
        if(dir == 0)
        {
            cor_trans1 = g.twoD ? 1 : 2; cor_trans2 = g.twoD ? 2 : 1;
            param.stride = g.ln[0];
            if(sh[1] && g.last) (g.twoD ? trans1FieldOff : trans2FieldOff) = 1;
            if(sh[2]) (g.twoD ? trans2FieldOff : trans1FieldOff) = 1;
            if(sh[0]) { if(pl) ccStart = 1; else dirMax -= 1; }
        }
        else if(dir == 1)
        {
            cor_trans1 = 0; cor_trans2 = 2; param.stride = 1;
            if(sh[0]) trans1FieldOff = 1;
            if(sh[2]) trans2FieldOff = 1;
            if(pl && g.last && sh[1]) ccStart = 1;
            if(!pl && startPt + dirMax == nDir && sh[1]) dirMax -= 1;
        }
        else
        {
            cor_trans1 = 0; cor_trans2 = 1; param.stride = 1;
            if(sh[0]) trans1FieldOff = 1;
            if(sh[1] && g.last) trans2FieldOff = 1;
            if(sh[2]) { if(pl) ccStart = 1; else dirMax -= 1; }
        }
        const std::vector<double>& eta_eff = eta[dir][pl ? 1 : 0];
        param.transSz = g.ln[cor_trans1] - 2 - trans1FieldOff;
        transSz2 = g.ln[cor_trans2] - 2 - trans2FieldOff;
        if(g.twoD && cor_trans2 == 2) transSz2 = 1;
        const double distOff = E ? 0.0 : 0.5;
        int loc[3] = {-1, -1, -1};
        loc[cor_trans1] = 1;
        for(int cc = ccStart; cc < dirMax; ++cc)
        {
            loc[cor_norm] = pl ? g.ln[cor_norm] - 2 - cc : cc + 1;
            double dist = g.procLoc(cor_norm) + (loc[cor_norm] - 1) + distOff;
            if(pl) dist = g.n[cor_norm] - 1 - dist;
            const double sig = sigma(dist, static_cast<double>(nDir - 1), eta_eff.at((size_t)dist), sigmaMax);
            const double kap = kappa(dist, static_cast<double>(nDir - 1));
            const double a = aVal(dist, static_cast<double>(nDir - 1));
            param.b = b(sig, a, kap);
            param.c = c(sig, a, kap) / g.d[cor_norm];
            if(!E) param.c *= -1.0;
            for(int jj = 0; jj < transSz2; ++jj)
            {
                loc[cor_trans2] = 1 + jj;
                if(g.twoD) loc[2] = 0;
                int locOff[3] = {loc[0], loc[1], loc[2]};
                locOff[cor_norm] += E ? -1 : 1;
                param.ind = g.ind(loc[0], loc[1], loc[2]);
                param.indOff = g.ind(locOff[0], locOff[1], locOff[2]);
                list.push_back(param);
            }
        }
    }

    // getGridUpList + getAxLists (:462-656)
    void gridList(int dir, int derivDir, bool pl, std::vector<ChimlGridParams>& list)
    {
        ChimlGridParams param;
        int mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
        const int cor_ii = dir, cor_jj = (dir + 1) % 3, cor_kk = (dir + 2) % 3;
        int off_ii = 0;
        param.Db = g.dt;
        double DbFieldBase = g.dt;
        int offset[3] = {0, 0, 0};
        param.stride = dir == 0 ? g.ln[0] : 1;
        if(dir == 1)
        {
            if(pl && g.last && sh[1]) off_ii = -1;
            if(!pl) off_ii = 1;
        }
        else
        {
            if(pl && sh[dir]) off_ii = -1;
            else if(!pl) off_ii = 1;
        }
        // signs from Taflove ch. 7 (:555-581), by component: 0 Ex 1 Ey 2 Ez 3 Hx 4 Hy 5 Hz
        if(derivDir == 0)      { if(comp == 5 || comp == 1) param.Db *= -1.0; if(comp == 4 || comp == 1) DbFieldBase *= -1.0; }
        else if(derivDir == 1) { if(comp == 3 || comp == 2) param.Db *= -1.0; if(comp == 5 || comp == 2) DbFieldBase *= -1.0; }
        else                   { if(comp == 4 || comp == 0) param.Db *= -1.0; if(comp == 3 || comp == 0) DbFieldBase *= -1.0; }
        DbFieldBase /= g.d[derivDir];
        offset[derivDir] = E ? -1 : 1;
        mn[cor_ii] += pl ? g.ln[cor_ii] - g.pl[cor_ii] - 1 : 1;
        mx[cor_ii] += off_ii + (pl ? g.ln[cor_ii] - 1 : g.mn[cor_ii]);
        mn[cor_jj] = 1; mx[cor_jj] = g.ln[cor_jj] - 1;
        mn[cor_kk] = 1; mx[cor_kk] = g.ln[cor_kk] - 1;
        auto trim = [&](int ax) {
            mn[ax] += g.mn[ax];
            mx[ax] -= g.pl[ax] - ((E || (ax == 1 && !g.last)) ? 0 : 1);
        };
        if(dir == i_) { trim(cor_jj); trim(cor_kk); }
        else if(derivDir != cor_ii) trim(derivDir);
        if(cor_ii != 0 && sh[0]) mx[0] -= 1;
        if(cor_ii != 1 && g.last && sh[1]) mx[1] -= 1;
        if(cor_ii != 2 && !g.twoD && sh[2]) mx[2] -= 1;
        if(g.twoD) { mn[2] = 0; mx[2] = 1; }
        const double distOff = E ? 0.0 : 0.5;
        auto push = [&](int x, int y, int z, int n) {
            const int ax[3] = {x, y, z};
            double dist = g.procLoc(derivDir) + ax[derivDir] - 1 + distOff;
            if(pl) dist = g.n[derivDir] - 1 - dist;
            const double kap = kappa(dist, static_cast<double>(g.thick[derivDir] - 1));
            param.DbField = DbFieldBase / kap;
            param.ind = g.ind(x, y, z);
            param.indOff = g.ind(x + offset[0], y + offset[1], z + offset[2]);
            param.nAx = n;
            list.push_back(param);
        };
        if(param.stride == 1)
        {
            for(int kk = mn[2]; kk < mx[2]; ++kk)
                for(int jj = mn[1]; jj < mx[1]; ++jj) push(mn[0], jj, kk, mx[0] - mn[0]);
        }
        else if(g.twoD)
        {
            for(int kk = mn[2]; kk < mx[2]; ++kk)
                for(int ii = mn[0]; ii < mx[0]; ++ii) push(ii, mn[1], kk, mx[1] - mn[1]);
        }
        else
        {
            for(int jj = mn[1]; jj < mx[1]; ++jj)
                for(int ii = mn[0]; ii < mx[0]; ++ii) push(ii, jj, mn[2], mx[2] - mn[2]);
        }
    }

    // initalizeLists (:669-689) for the six faces in the constructor's order (:176-181)
    void build()
    {
        for(int dir = 0; dir < 3; ++dir)
        {
            if(dir == 2 && g.twoD) continue;
            genEta(dir, false);
            genEta(dir, true);
        }
        for(int dir = 0; dir < 3; ++dir)
            for(int side = 0; side < 2; ++side)
            {
                const bool pl = side == 1;
                const int ln_pml = pl ? g.pl[dir] : g.mn[dir];
                if(ln_pml == 0) continue;
                const int startPt = pl ? g.procLoc(dir) + g.ln[dir] - 2 : g.procLoc(dir);
                if(hasGrid[0])
                {
                    if(dir == j_) psiList(dir, pl, startPt, g.thick[dir], ln_pml, psi[0]);
                    gridList(dir, j_, pl, grid[0]);
                }
                if(hasGrid[1])
                {
                    if(dir == k_) psiList(dir, pl, startPt, g.thick[dir], ln_pml, psi[1]);
                    gridList(dir, k_, pl, grid[1]);
                }
            }
    }
};

// DTCTYPE -> stored field (parallelFDTDField.cpp:689-826).  Power detectors whose components sit at different Yee positions (E power
// in 3-D / TE, H power in 3-D / TM) are refused: the reference loops over the box of its FIRST stored field and indexes the boxes of
// the others -- which are grown along other axes -- with it (DTC/parallelDTC_TXT.cpp:38-50), reading past their ends.
int detector_field(DTCTYPE t, int mode)
{
    if(t == DTCTYPE::HPOW && mode == CHIML_MODE_TE) return CHIML_HZ;      // H power of a TE grid: Hz alone, offset 0 (:812-815)
    if(t == DTCTYPE::EPOW && mode == CHIML_MODE_TM) return CHIML_EZ;      // E power of a TM grid: Ez alone, offset 0 (:816-819)
    if(t == DTCTYPE::EPOW || t == DTCTYPE::HPOW)
        throw std::logic_error("power detectors over several field components: the reference indexes the stored boxes out of bounds; not reproduced");
    switch(t)
    {
        case DTCTYPE::EX: return CHIML_EX; case DTCTYPE::EY: return CHIML_EY; case DTCTYPE::EZ: return CHIML_EZ;
        case DTCTYPE::HX: return CHIML_HX; case DTCTYPE::HY: return CHIML_HY; case DTCTYPE::HZ: return CHIML_HZ;
        case DTCTYPE::DX: return CHIML_DX; case DTCTYPE::DY: return CHIML_DY; case DTCTYPE::DZ: return CHIML_DZ;
        default: throw std::logic_error("power / polarisation detectors are outside the covered hot path");
    }
}

#include "emitters.inc"
#include "flux.inc"

} // namespace

// First rows of the y-slabs as the reference cuts them: rows are weighted by an operation count per grid point (base cost, poles of
// the objects at the three E positions, CPML layers, flux surfaces) and every slab gets about the mean weight
// (parallelFDTDField.hpp:1097-1210 setupWeightsGrid, MPI/mpiInterface.cpp:21-53 getLocxLocyLocz(weights)).  Used with
// `chiml_plan --split reference`; the engine itself is indifferent to where the cuts are.
std::vector<int> reference_split(const Inputs& IP, const Geom& g, int nranks)
{
    const int n0 = g.n[0], n1 = g.n[1], n2 = g.n[2];
    std::vector<double> w(n1, 3.0 * 276.0 * n0 * n2);          // `IP.size_[2] != 1 ? 276.0 : 184.0` is 276 for every input
    for(const auto& obj : IP.objArr_)
    {
        const double dbAdd = obj->useOrientedDipols_ ? 189.0 : 169.0;
        const double add = dbAdd * obj->gamma_.size() + (obj->pols_.size() > 0 ? 30.0 : 0.0);
        if(add == 0.0) continue;
        for(int jj = 0; jj < n1; ++jj)
            for(int ii = 0; ii < n0; ++ii)
                for(int kk = 0; kk < n2; ++kk)
                {
                    std::array<double, 3> pt = {{(ii - (n0 - 1) / 2.0 + 0.5) * g.d[0], (jj - (n1 - 1) / 2.0) * g.d[1], (kk - (n2 - 1) / 2.0) * g.d[2]}};
                    if(obj->isObj(pt, g.d[0], obj->geoParam_)) w[jj] += add;
                    pt[1] += 0.5 * g.d[1]; pt[0] -= 0.5 * g.d[0];
                    if(obj->isObj(pt, g.d[0], obj->geoParam_)) w[jj] += add;
                    pt[1] -= 0.5 * g.d[1]; pt[2] += 0.5 * g.d[2];
                    if(obj->isObj(pt, g.d[0], obj->geoParam_)) w[jj] += add;
                }
    }
    // CPML layers: x layers on the Ey, Ez maps (every row alike), y layers on Ex, Ez, z layers on Ex, Ey
    for(int jj = 0; jj < n1; ++jj) w[jj] += 2.0 * 2.0 * g.thick[0] * 79.0 * n2;
    for(int yy = 0; yy < g.thick[1]; ++yy) { w[yy] += 2.0 * 79.0 * n0 * n2; w[n1 - 1 - yy] += 2.0 * 79.0 * n0 * n2; }
    if(!g.twoD) for(int jj = 0; jj < n1; ++jj) w[jj] += 2.0 * 2.0 * g.thick[2] * 79.0 * n0;
    if(!g.twoD)
        for(const FluxInput& f : IP.fluxes_)
        {
            const double fw = f.freqs.size() * 3 + 20.0;
            for(int yy = 0; yy < f.sz[1]; ++yy) w[f.loc[1] + yy] += 3.0 * fw * (2.0 * f.sz[0] + 2.0 * f.sz[2]);
            for(int zz = 0; zz < f.sz[2]; ++zz) { w[f.loc[1]] += 3.0 * fw * f.sz[0]; w[f.loc[1] + f.sz[1] - 1] += 3.0 * fw * f.sz[0]; }
        }
    double sum = 0.0;
    for(double v : w) sum += v;
    const double avg = sum / static_cast<double>(nranks);
    std::vector<int> start(nranks + 1, 0);
    start[nranks] = n1;
    double val = 0.0;
    int curY = 0;
    for(int cc = 0; cc < nranks - 1; ++cc)
    {
        curY = start[cc];
        while(val + w[curY] / 2.0 < avg || curY == start[cc]) { val += w[curY]; curY++; }
        val -= avg;
        start[cc + 1] = curY;
    }
    return start;
}

// ---------------------------------------------------------------------------------------------------
SlabPlan build_plan(Inputs& IP, int rank, int nranks, int nthreads, bool referenceSplit)
{
    if(nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    SlabPlan P;
    Geom g;
    g.rank = rank; g.nranks = nranks; g.last = rank == nranks - 1;
    g.twoD = IP.size_[2] == 0;
    for(int k = 0; k < 3; ++k)
    {
        g.n[k] = static_cast<int>(std::floor(IP.size_[k] / IP.d_[k] + 0.5)) + 1;   // toN_vec, parallelFDTDField.hpp:1714
        g.d[k] = IP.d_[k];
        g.thick[k] = IP.pmlThickness_[k];
    }
    g.dt = IP.dt_;
    // equal-height y-slabs (mpiInterface::getLocxLocyLocz(int,int,int), MPI/mpiInterface.cpp:55-58)
    const int base = g.n[1] / nranks, rem = g.n[1] % nranks;
    int nyloc = base + (rank < rem ? 1 : 0);
    g.yStart = rank * base + std::min(rank, rem);
    if(referenceSplit && nranks > 1)
    {
        // the reference's cost-weighted cuts instead (pole constants must exist: the weights count poles)
        for(auto& obj : IP.objArr_)
            if(!obj->constsSet_) obj->setUpConsts(IP.dt_);
        g.twoD = IP.size_[2] == 0;
        const std::vector<int> start = reference_split(IP, g, nranks);
        g.yStart = start[rank]; nyloc = start[rank + 1] - start[rank];
    }
    if(nyloc < 1) throw std::logic_error("a y-slab is empty: fewer grid rows than ranks");
    g.ln[0] = g.n[0] + 2; g.ln[1] = nyloc + 2; g.ln[2] = g.twoD ? 1 : g.n[2] + 2;
    // findLnVecs (PML/parallelPML.hpp:280-308)
    for(int k = 0; k < 3; ++k) { g.mn[k] = 0; g.pl[k] = 0; }
    for(int k = 0; k < 3; ++k)
    {
        if(k == 2 && g.twoD) continue;
        const int loc = g.procLoc(k), lsz = g.ln[k] - 2;
        if(loc < g.thick[k]) g.mn[k] = loc + lsz < g.thick[k] ? lsz : g.thick[k] - loc;
        if(loc + lsz > g.n[k] - g.thick[k]) g.pl[k] = loc > g.n[k] - g.thick[k] ? lsz : loc + lsz - (g.n[k] - g.thick[k]);
    }
    for(auto& obj : IP.objArr_)
        if(!obj->constsSet_) obj->setUpConsts(IP.dt_);

    const int mode = !g.twoD ? CHIML_MODE_3D
                             : ((IP.pol_ == POLARIZATION::HZ || IP.pol_ == POLARIZATION::EX || IP.pol_ == POLARIZATION::EY) ? CHIML_MODE_TE : CHIML_MODE_TM);
    const int nvec[3] = {g.n[0], g.n[1], g.n[2]};
    materials_in_pml(IP, nvec, g.d, P.dielectricMatInPML, P.magMatInPML);
    // which grids exist (parallelFDTDField.hpp:372-389): D for dispersive / chiral media, B for magnetic / chiral ones
    bool disp = P.dielectricMatInPML, magnetic = P.magMatInPML, chiral = false;
    int nLor = 0, nOrDip = 0, nMag = 0;
    for(const auto& obj : IP.objArr_)
    {
        if(obj->gamma_.size() > 0 || obj->ML_ || obj->eps_infty_ > 1.0) disp = true;
        if(obj->magGamma_.size() > 0 || obj->mu_infty_ > 1.0) magnetic = true;
        if(obj->chiGamma_.size() > 0) chiral = true;
        if(obj->useOrientedDipols_ && (obj->magGamma_.size() > 0 || obj->chiGamma_.size() > 0))
            throw std::logic_error("oriented-dipole objects with magnetic / chiral poles are outside the covered hot path");
        nLor = std::max(nLor, (int)obj->gamma_.size());
        nMag = std::max(nMag, (int)obj->magGamma_.size());
        if(obj->useOrientedDipols_) nOrDip = std::max(nOrDip, (int)obj->gamma_.size());
    }
    if(chiral && g.twoD) throw std::logic_error("chiral media on a 2-D grid are outside the covered hot path");
    if(chiral && nranks > 1) throw std::logic_error("chiral media are covered for single-slab runs");
    if(chiral) disp = true;
    P.has_B = magnetic || chiral;
    P.n_mag_poles = P.has_B ? nMag : 0;

    std::memset(&P.grid, 0, sizeof(P.grid));
    P.grid.desc.mode = mode;
    for(int k = 0; k < 3; ++k) { P.grid.desc.ln[k] = g.ln[k]; P.grid.desc.d[k] = g.d[k]; P.grid.n_global[k] = g.n[k]; }
    P.grid.desc.dt = g.dt;
    P.grid.desc.has_D = disp ? 1 : 0;
    P.grid.desc.pml_on_D = P.dielectricMatInPML ? 1 : 0;
    P.grid.desc.n_objects = (int)IP.objArr_.size();
    P.grid.desc.rank = rank; P.grid.desc.nranks = nranks;
    P.grid.y_start = g.yStart;
    P.grid.n_steps = int(std::ceil(IP.tMax_ / IP.dt_));
    P.grid.n_lor_poles = disp ? nLor : 0;
    P.grid.n_ordip_poles = disp ? nOrDip : 0;
    P.grid.t_max = IP.tMax_;

    // ---- periodic boundaries: the arguments step() hands to applBCE_ / applBCH_ (parallelFDTDField.hpp:1267-1269,1285-1287) with
    // yEPBC_ / yHPBC_ of the single-process branch (parallelFDTDField.cpp:163-171,329-336) and zMinPBC_ / zMaxPBC_ (hpp:444-445) ----
    if(IP.periodic_)
    {
        if(nOrDip > 0) throw std::logic_error("oriented-dipole media under periodic boundaries are outside the covered hot path");
        if(nranks > 1 && (IP.cplxFields_ || magnetic || chiral))
            throw std::logic_error("periodic runs on several slabs are covered for real fields without magnetic / chiral media");
        const int l0 = g.ln[0] - 2, l1 = g.ln[1] - 2;
        const int zMin = g.twoD ? 0 : 1, zMax = g.twoD ? 1 : g.ln[2] - 2;
        // y-limited components (fieldEnd[1] = 1: Ey, Hx, Hz) wrap at row ln_vec_[1], the others at ln_vec_[1] + 1
        const int yE[3] = {l1 + 1, l1, l1 + 1}, yH[3] = {l1, l1 + 1, l1};
        const ChimlWrap w[6] = {{l0 - 1, yE[0], zMax, l0, yE[0], zMin, zMax + 1},     {l0, yE[1], zMax, l0 + 1, yE[1], zMin, zMax + 1},
                                {l0, yE[2], zMax - 1, l0 + 1, yE[2], zMin, zMax},     {l0, yH[0], zMax - 1, l0 + 1, yH[0], zMin, zMax},
                                {l0 - 1, yH[1], zMax - 1, l0, yH[1], zMin, zMax},     {l0 - 1, yH[2], zMax, l0, yH[2], zMin, zMax + 1}};
        for(int comp = 0; comp < 6; ++comp)
            if(comp_exists(mode, comp))
            {
                ChimlPlanPeriodic pp; pp.comp = comp; pp.wrap = w[comp];
                // a slab of several: the x / z wraps of the owned rows only (applyBCProcMid on every rank); the y direction is the ghost-row ring
                // between the slabs, slab 0 <-> slab nranks - 1 included (include/chiml_gpu.h chiml_gpu_set_periodic)
                if(nranks > 1) { pp.wrap.ny = -1; pp.wrap.ymax = -1; }
                P.periodic.push_back(pp);
            }
    }
    if(IP.cplxFields_)
    {
        // parallelFDTDFieldCplx: everything above and below is the same set-up; emitters, flux regions and frequency detectors of a
        // complex run are not reproduced
        if(!IP.qes_.empty() || !IP.fluxes_.empty() || !IP.freqDtcs_.empty())
            throw std::logic_error("complex-field runs with emitters / flux regions / frequency detectors are outside the covered hot path");
        P.cplx = true;
        for(int k = 0; k < 3; ++k) P.k_point[k] = IP.k_point_[k];
    }

    Rasteriser ras(IP, g);
    std::vector<std::array<int, 5>> chiLocs;
    // ---- update lists (parallelFDTDField.cpp:80-92,248-261) ----
    for(int comp = 0; comp < 6; ++comp)
    {
        if(!comp_exists(mode, comp)) continue;
        const int i = comp % 3;
        const double dj = g.d[(i + 1) % 3], dk = g.d[(i + 2) % 3];
        ListSet ls;
        build_lists(IP, g, ras, SPEC_COMP[comp], comp < 3, DERIV_OFF[comp], SPEC_COMP[comp].endOff, dj, dk, P.dielectricMatInPML, P.magMatInPML, false, nthreads, ls);
        P.lists[CHIML_LIST_U][comp] = std::move(ls.U);
        if(comp < 3)
        {
            P.lists[CHIML_LIST_D][comp] = std::move(ls.D);
            P.lists[CHIML_LIST_LORD][comp] = std::move(ls.LorD);
            P.lists[CHIML_LIST_ORDIPD][comp] = std::move(ls.OrDipD);
        }
        else if(P.has_B)
        {
            // upB_ / upLorB_ (with B grids; without them an H component has nothing but upH_)
            P.lists[CHIML_LIST_D][comp] = std::move(ls.D);
            P.lists[CHIML_LIST_LORD][comp] = std::move(ls.LorD);
        }
        else if(!ls.D.empty() || !ls.LorD.empty()) throw std::logic_error("magnetic update lists without B grids");
        P.lists[CHIML_LIST_CHID][comp] = std::move(ls.ChiD);
        chiLocs.insert(chiLocs.end(), ls.chiLocs.begin(), ls.chiLocs.end());
    }
    // copy2PrevFields_ (parallelFDTDField.cpp:391-410): per chiral object the box of its ChiD / ChiB runs, one cell wider on every side
    if(chiral)
    {
        for(int oo = 0; oo < (int)IP.objArr_.size(); ++oo)
        {
            if(IP.objArr_[oo]->chiGamma_.empty()) continue;
            int mnC[3] = {g.n[0], g.n[1], g.n[2]}, mxC[3] = {0, 0, 0};
            for(const auto& ax : chiLocs)
            {
                if(ax[4] != oo) continue;
                mxC[0] = std::max(mxC[0], ax[1] + ax[0] - 1); mxC[1] = std::max(mxC[1], ax[2]); mxC[2] = std::max(mxC[2], ax[3]);
                mnC[0] = std::min(mnC[0], ax[1]); mnC[1] = std::min(mnC[1], ax[2]); mnC[2] = std::min(mnC[2], ax[3]);
            }
            const int sz = mxC[0] - mnC[0] + 3;
            for(int yy = mnC[1] - 1; yy <= mxC[1] + 1; ++yy)
                for(int zz = mnC[2] - 1; zz <= mxC[2] + 1; ++zz) P.prev_copy.push_back({{sz, mnC[0] - 1, yy, zz}});
        }
    }
    if(disp && nOrDip > 0)
    {
        // node-centred oriented-dipole list (parallelFDTDField.cpp:83-87,252-256): dipP_ exists for the in-plane components
        // whenever an oriented-dipole object carries electric poles
        const int dOff[3] = {-1, -1, -1}, fEnd[3] = {0, 0, 0};
        ListSet ls;
        build_lists(IP, g, ras, SPEC_NODE_P, true, dOff, fEnd, g.d[0], g.d[0], P.dielectricMatInPML, P.magMatInPML, true, nthreads, ls);
        P.lists[CHIML_LIST_ORDIPP][0] = std::move(ls.OrDipD);
    }
    // ---- dipole grids of orientations relative to the surface normal (setupDipMoments, parallelFDTDField.hpp:960-1048; getTangentDip :1058-1093):
    // evaluated at the cells of the node list -- the only ones the update reads -- for every component and every pole index of the grid
    {
        bool relToNorm = false;
        for(const auto& obj : IP.objArr_)
            if(obj->useOrientedDipols_)
                for(DIPOR o : obj->dipOr_) relToNorm = relToNorm || (o != DIPOR::ISOTROPIC && o != DIPOR::UNIDIRECTIONAL);
        if(relToNorm && disp && nOrDip > 0)
        {
            if(g.twoD) throw std::logic_error("oriented-dipole poles on a 2-D grid are outside the covered hot path");
            const size_t ncell = (size_t)g.ln[0] * g.ln[1] * g.ln[2];
            for(int c = 0; c < 3; ++c)
                for(int p = 0; p < nOrDip; ++p) P.dip_grids.push_back({c, p, std::vector<double>(ncell, 0.0)});
            auto tangentDip = [](const std::array<double, 3>& normVec, double latFact, double longFact) {
                double magSq = 0.0;
                for(double x : normVec) magSq = magSq + x * x;
                const double mag = std::sqrt(magSq);
                std::array<double, 3> sph = {{mag, std::acos(normVec[2] / mag), std::atan(normVec[1] / normVec[0])}};
                if(normVec[0] == 0.0) sph[2] = normVec[1] >= 0.0 ? M_PI / 2.0 : -1.0 * M_PI / 2.0;
                if(normVec[0] < 0) sph[2] += M_PI;
                if(mag < 1e-20) sph = {{0.0, 0.0, 0.0}};
                const std::array<double, 3> tLongSph = {{mag, sph[1] + M_PI / 2.0, sph[2]}};
                const std::array<double, 3> tLong = {{mag * std::sin(tLongSph[1]) * std::cos(tLongSph[2]), mag * std::sin(tLongSph[1]) * std::sin(tLongSph[2]), mag * std::cos(tLongSph[1])}};
                std::array<double, 3> tLat, out;
                for(int ii = 0; ii < 3; ++ii) tLat[ii] = normVec[(ii + 1) % 3] * tLong[(ii + 2) % 3] - normVec[(ii + 2) % 3] * tLong[(ii + 1) % 3];
                for(int ii = 0; ii < 3; ++ii) out[ii] = longFact * tLong[ii] + latFact * tLat[ii];
                return out;
            };
            for(const ChimlRun& r : P.lists[CHIML_LIST_ORDIPP][0])
            {
                const Obj& o = *IP.objArr_[r.obj];
                const int x0 = r.ind % g.ln[0], row = r.ind / g.ln[0], kk = row % g.ln[2], jj = row / g.ln[2];
                for(int p = 0; p < (int)o.gamma_.size() && p < nOrDip; ++p)
                {
                    const DIPOR how = o.dipOr_[p];
                    if(how == DIPOR::REL_TO_NORM && !o.identityAxes())
                        throw std::logic_error("surface-normal-relative dipoles of a rotated object are outside the covered hot path");
                    for(int i = 0; i < r.n; ++i)
                    {
                        const int ii = x0 + i;
                        std::array<double, 3> val = {{0.0, 0.0, 0.0}};
                        if(how == DIPOR::ISOTROPIC) val = {{1.0, 1.0, 1.0}};
                        else if(how == DIPOR::UNIDIRECTIONAL) val = o.dipE_[p];
                        else if(how == DIPOR::REL_TO_NORM)
                        {
                            const std::array<double, 3> pt = {{((ii - 1) + 0.0 + g.procLoc(0) - (g.n[0] - g.n[0] % 2) / 2.0) * g.d[0],
                                                               ((jj - 1) + 0.0 + g.procLoc(1) - (g.n[1] - g.n[1] % 2) / 2.0) * g.d[1],
                                                               ((kk - 1) + 0.0 + g.procLoc(2) - (g.n[2] - g.n[2] % 2) / 2.0) * g.d[2]}};
                            const std::array<double, 3> grad = o.findGradient(pt);
                            const std::array<double, 3> tan = tangentDip(grad, o.dipTanLatCompE_[p], o.dipTanLongCompE_[p]);
                            for(int c = 0; c < 3; ++c) val[c] = o.dipNormCompE_[p] * grad[c] + tan[c];
                        }
                        for(int c = 0; c < 3; ++c) P.dip_grids[(size_t)c * nOrDip + p].grid[(size_t)r.ind + i] = val[c];
                    }
                }
            }
        }
    }
    // ---- objects ----
    for(const auto& obj : IP.objArr_)
    {
        PlanObject o;
        o.npoles = (int)obj->gamma_.size(); o.use_or_dip = obj->useOrientedDipols_ ? 1 : 0; o.ml = obj->ML_ ? 1 : 0;
        if(g.twoD && o.use_or_dip && o.npoles > 0 && !o.ml)
            throw std::logic_error("oriented-dipole (unidirectional) poles on a 2-D grid are outside the covered hot path: the reference's update "
                                   "lists for them index z neighbours a 2-D grid does not have (its own parallelGrid::getInd assertion fails)");
        o.eps_inf = obj->eps_infty_; o.mu_inf = obj->mu_infty_;
        o.alpha = obj->alpha_; o.xi = obj->xi_; o.gamma = obj->gamma_;
        o.magAlpha = obj->magAlpha_; o.magXi = obj->magXi_; o.magGamma = obj->magGamma_;
        o.chiAlpha = obj->chiAlpha_; o.chiXi = obj->chiXi_; o.chiGamma = obj->chiGamma_; o.chiGammaPrev = obj->chiGammaPrev_;
        o.dip.assign(3 * (size_t)o.npoles, 0.0);
        if(o.use_or_dip)
            for(int p = 0; p < o.npoles; ++p)
                for(int k = 0; k < 3; ++k)      // setupDipMoments :998-1007; position-dependent orientations live in the dipole grids (record DIPGRID)
                    o.dip[3 * p + k] = obj->dipOr_[p] == DIPOR::ISOTROPIC ? 1.0 : (obj->dipOr_[p] == DIPOR::UNIDIRECTIONAL ? obj->dipE_[p][k] : 0.0);
        P.objects.push_back(o);
    }
    // ---- CPML (parallelFDTDField.cpp:60-77,229-246) ----
    for(int comp = 0; comp < 6; ++comp)
    {
        if(!comp_exists(mode, comp)) continue;
        const int i = comp % 3;
        CpmlBuilder cb{IP, g, ras, comp, comp < 3, i, (i + 1) % 3, (i + 2) % 3, {SPEC_COMP[comp].endOff[0], SPEC_COMP[comp].endOff[1], SPEC_COMP[comp].endOff[2]}, {false, false}, {}, {}, {}};
        const int other = comp < 3 ? 3 : 0;
        cb.hasGrid[0] = comp_exists(mode, other + (i + 2) % 3);   // grid_k
        cb.hasGrid[1] = comp_exists(mode, other + (i + 1) % 3);   // grid_j
        cb.build();
        for(int part = 0; part < 2; ++part)
        {
            if(!cb.hasGrid[part]) continue;
            // psi_j exists unless 2-D and j == Z; psi_k likewise (PML/parallelPML.hpp:143-156)
            const int axis = part == 0 ? cb.j_ : cb.k_;
            PlanCpml pc;
            pc.comp = comp; pc.part = part; pc.has_psi = (!g.twoD || axis != 2) ? 1 : 0;
            pc.psi = std::move(cb.psi[part]); pc.grid = std::move(cb.grid[part]);
            P.cpml.push_back(std::move(pc));
        }
    }
    // ---- sources (parallelFDTDField.cpp:457-560; SOURCE/parallelSourceNormal.hpp:69-140) ----
    for(const SourceInput& s : IP.sources_)
    {
        const int field = (int)s.pol;   // EX..HZ share the ChimlField numbering
        if(!comp_exists(mode, field)) throw std::logic_error("a source acts on a field component that does not exist in this mode");
        PlanSource ps;
        ps.field = field;
        bool inside = true;
        for(int k = 0; k < 3; ++k)
        {
            const int pl = g.procLoc(k), lsz = g.ln[k] - 2;
            int l;
            if(s.loc[k] >= pl && s.loc[k] < pl + lsz) l = s.loc[k] - pl + 1;
            else if(s.loc[k] < pl && s.loc[k] + s.sz[k] > pl) l = 1;
            else l = -1;
            if(k == 2 && g.twoD) l = 0;
            ps.loc[k] = l;
            if(l == -1) inside = false;
        }
        if(!inside) continue;
        for(int k = 0; k < 3; ++k)
        {
            const int pl = g.procLoc(k);
            if(s.sz[k] + s.loc[k] > pl + g.ln[k] - 2) ps.sz[k] = g.ln[k] - ps.loc[k] - 1;
            else ps.sz[k] = s.loc[k] + s.sz[k] - (pl + ps.loc[k] - 1);
            if(k == 2 && g.twoD) ps.sz[k] = 1;
        }
        ps.amp.resize(P.grid.n_steps);
        if(IP.cplxFields_) ps.amp_im.resize(P.grid.n_steps);
        double t = 0.0;
        for(int k = 0; k < P.grid.n_steps; ++k)
        {
            cplx pulVal = 0.0;
            for(size_t p = 0; p < s.shapes.size(); ++p) pulVal += pulseValue(s.shapes[p], t, s.fxn[p]);
            ps.amp[k] = g.dt * std::real(pulVal);
            if(IP.cplxFields_) ps.amp_im[k] = g.dt * std::imag(pulVal);      // parallelSourceNormalCplx::addPul: zaxpy_(n, dt_, pulVec_, ...)
            t += g.dt;
        }
        P.sources.push_back(std::move(ps));
    }
    // ---- detectors (DTC/parallelDTC.hpp:44-93, parallelStorageDTC.hpp:51-80) ----
    int dd = 0;
    for(const DetectorInput& d : IP.detectors_)
    {
        PlanDetector pd;
        pd.detector = dd++;
        pd.field = detector_field(d.type, mode);
        if(!comp_exists(mode, pd.field % 3 + (pd.field >= 3 && pd.field < 6 ? 3 : 0))) throw std::logic_error("a detector samples a field component that does not exist in this mode");
        for(int k = 0; k < 3; ++k) { pd.loc[k] = d.loc[k]; pd.sz[k] = d.sz[k]; pd.offset[k] = 0; }
        pd.every = static_cast<int>(std::floor(d.timeInt / IP.dt_));
        if(pd.every == 0) throw std::logic_error("The time step of a detector is less than the main grid or set to 0.");
        pd.type = (int)d.type;
        pd.conv = 1.0; pd.t_conv = 1.0;
        if(d.SI)
        {
            pd.t_conv *= IP.a_ / SPEED_OF_LIGHT;
            pd.conv = (IP.I0_ / IP.a_);
            if(d.type == DTCTYPE::EX || d.type == DTCTYPE::EY || d.type == DTCTYPE::EZ) pd.conv /= EPS0() * SPEED_OF_LIGHT;
        }
        if(d.type == DTCTYPE::EPOW || d.type == DTCTYPE::HPOW) pd.conv *= pd.conv;     // power: the square of the field factor (DTC/parallelDTC.hpp:87-91)
        P.detectors.push_back(pd);
    }
    // ---- emitters (parallelFDTDField.cpp:412-454) ----
    build_emitters(IP, g, ras, mode, P);
    // ---- flux regions: running-DFT sets (parallelFDTDField.cpp:650-682) ----
    build_fluxes(IP, g, mode, P);
    build_freq_detectors(IP, g, mode, P);
    return P;
}

// ---------------------------------------------------------------------------------------------------
namespace {
void put_rec(std::ofstream& out, const char* tag, const std::string& payload)
{
    char t[8];
    std::memset(t, ' ', 8);
    std::memcpy(t, tag, std::min<size_t>(8, std::strlen(tag)));
    const uint64_t n = payload.size();
    out.write(t, 8);
    out.write(reinterpret_cast<const char*>(&n), 8);
    out.write(payload.data(), (std::streamsize)n);
}
template <typename T> void app(std::string& s, const T& v) { s.append(reinterpret_cast<const char*>(&v), sizeof(T)); }
template <typename T> void app_vec(std::string& s, const std::vector<T>& v) { if(!v.empty()) s.append(reinterpret_cast<const char*>(v.data()), v.size() * sizeof(T)); }
}

void SlabPlan::write(const std::string& path) const
{
    std::ofstream out(path.c_str(), std::ios::binary);
    if(!out) throw std::runtime_error("cannot write " + path);
    { std::string p; int32_t v = CHIML_PLAN_VERSION; app(p, v); put_rec(out, "CHIMLPLN", p); }
    { std::string p; app(p, grid); put_rec(out, "GRID", p); }
    if(cplx)
    {
        ChimlPlanComplex pc; std::memset(&pc, 0, sizeof(pc));
        pc.cplx = 1; for(int k = 0; k < 3; ++k) pc.k_point[k] = k_point[k];
        std::string p; app(p, pc); put_rec(out, "COMPLEX", p);
    }
    for(const ChimlPlanPeriodic& pp : periodic) { std::string p; app(p, pp); put_rec(out, "PERIODIC", p); }
    // same record order as oracle/ref_driver.cpp
    for(int c = 0; c < 3; ++c)
    {
        const int kinds[5] = {CHIML_LIST_U, CHIML_LIST_U, CHIML_LIST_D, CHIML_LIST_LORD, CHIML_LIST_ORDIPD};
        const int comps[5] = {c, 3 + c, c, c, c};
        for(int k = 0; k < 5; ++k)
        {
            std::string p;
            ChimlPlanListHdr h; h.kind = kinds[k]; h.comp = comps[k]; h.n = lists[kinds[k]][comps[k]].size();
            app(p, h); app_vec(p, lists[kinds[k]][comps[k]]);
            put_rec(out, "UPLIST", p);
        }
    }
    for(int comp = 0; comp < 6; ++comp)
    {
        // upB_ / upLorB_ (B grids) and the chiral lists of either family
        const int kinds[3] = {CHIML_LIST_D, CHIML_LIST_LORD, CHIML_LIST_CHID};
        for(int k = 0; k < 3; ++k)
        {
            if(comp < 3 && k < 2) continue;
            if(k < 2 ? !has_B : lists[CHIML_LIST_CHID][comp].empty()) continue;
            std::string p;
            ChimlPlanListHdr h; h.kind = kinds[k]; h.comp = comp; h.n = lists[kinds[k]][comp].size();
            app(p, h); app_vec(p, lists[kinds[k]][comp]);
            put_rec(out, "UPLIST", p);
        }
    }
    if(has_B)
    {
        ChimlPlanMagnetic pm; std::memset(&pm, 0, sizeof(pm));
        pm.has_B = 1; pm.pml_on_B = magMatInPML ? 1 : 0; pm.n_mag_poles = n_mag_poles;
        std::string p; app(p, pm); put_rec(out, "MAGNETIC", p);
        for(size_t oo = 0; oo < objects.size(); ++oo)
        {
            const PlanObject& o = objects[oo];
            ChimlPlanObjMagHdr h; h.obj = (int)oo; h.npoles = (int)o.magGamma.size();
            std::string q; app(q, h); app_vec(q, o.magAlpha); app_vec(q, o.magXi); app_vec(q, o.magGamma);
            put_rec(out, "OBJMAG", q);
        }
    }
    {
        bool anyChi = false;
        for(const PlanObject& o : objects) anyChi = anyChi || !o.chiGamma.empty();
        if(anyChi)
        {
            for(size_t oo = 0; oo < objects.size(); ++oo)
            {
                const PlanObject& o = objects[oo];
                ChimlPlanObjChiHdr h; h.obj = (int)oo; h.npoles = (int)o.chiGamma.size();
                std::string q; app(q, h); app_vec(q, o.chiAlpha); app_vec(q, o.chiXi); app_vec(q, o.chiGamma); app_vec(q, o.chiGammaPrev);
                put_rec(out, "OBJCHI", q);
            }
            std::string q; const uint64_t nr = prev_copy.size(); app(q, nr); app_vec(q, prev_copy);
            put_rec(out, "PREVCOPY", q);
        }
    }
    for(const DipGrid& dg : dip_grids)
    {
        std::string q; const int32_t hd[2] = {dg.comp, dg.pole}; app(q, hd); app_vec(q, dg.grid);
        put_rec(out, "DIPGRID", q);
    }
    { std::string p; ChimlPlanListHdr h; h.kind = CHIML_LIST_ORDIPP; h.comp = 0; h.n = lists[CHIML_LIST_ORDIPP][0].size(); app(p, h); app_vec(p, lists[CHIML_LIST_ORDIPP][0]); put_rec(out, "UPLIST", p); }
    for(size_t oo = 0; oo < objects.size(); ++oo)
    {
        const PlanObject& o = objects[oo];
        ChimlPlanObjectHdr h; h.obj = (int)oo; h.npoles = o.npoles; h.use_or_dip = o.use_or_dip; h.ml = o.ml; h.eps_inf = o.eps_inf; h.mu_inf = o.mu_inf;
        std::string p; app(p, h); app_vec(p, o.alpha); app_vec(p, o.xi); app_vec(p, o.gamma); app_vec(p, o.dip);
        put_rec(out, "OBJECT", p);
    }
    for(const PlanCpml& c : cpml)
    {
        ChimlPlanCpmlHdr h; h.comp = c.comp; h.part = c.part; h.has_psi = c.has_psi; h.pad = 0; h.npsi = c.psi.size(); h.ngrid = c.grid.size();
        std::string p; app(p, h); app_vec(p, c.psi); app_vec(p, c.grid);
        put_rec(out, "CPML", p);
    }
    for(const PlanSource& s : sources)
    {
        ChimlPlanSourceHdr h; h.field = s.field;
        for(int k = 0; k < 3; ++k) { h.loc[k] = s.loc[k]; h.sz[k] = s.sz[k]; }
        h.n_steps = (int)s.amp.size();
        std::string p; app(p, h); app_vec(p, s.amp);
        put_rec(out, "SOURCE", p);
        if(cplx) { std::string q; const int32_t ns = (int32_t)s.amp_im.size(); app(q, ns); app_vec(q, s.amp_im); put_rec(out, "SRCIMAG", q); }
    }
    for(const PlanDetector& d : detectors)
    {
        ChimlPlanDetector r; std::memset(&r, 0, sizeof(r));
        r.detector = d.detector; r.field = d.field;
        for(int k = 0; k < 3; ++k) { r.loc[k] = d.loc[k]; r.sz[k] = d.sz[k]; r.offset[k] = d.offset[k]; }
        r.every = d.every; r.type = d.type; r.conv = d.conv; r.t_conv = d.t_conv;
        std::string p; app(p, r);
        put_rec(out, "DETECTOR", p);
    }
    for(const PlanEmitter& e : emitters)
    {
        ChimlPlanEmitterHdr h; std::memset(&h, 0, sizeof(h));
        h.object = e.object; h.nlevel = e.nlevel; h.nsys = e.nsys; h.nemit = e.nemit;
        for(int k = 0; k < 3; ++k) { h.box_lo[k] = e.box_lo[k]; h.box_n[k] = e.box_n[k]; }
        h.nnz = (int)e.gam_col.size(); h.npop = (int)e.pop_level.size(); h.pop_every = e.pop_every; h.npoints = e.npoints;
        h.pz = e.pz; h.dt = e.dt; h.inv_hbar = e.inv_hbar; h.na = e.na;
        std::string p; app(p, h);
        app_vec(p, e.h0); app_vec(p, e.weight); app_vec(p, e.mu); app_vec(p, e.gam_ptr); app_vec(p, e.gam_col); app_vec(p, e.gam_val);
        app_vec(p, e.loc); app_vec(p, e.eps); app_vec(p, e.pop_level);
        put_rec(out, "EMITTER", p);
    }
    for(const PlanDft& d : dfts)
    {
        ChimlPlanDftHdr h; std::memset(&h, 0, sizeof(h));
        h.field = d.field; h.group = d.group; h.every = d.every; h.nfreq = d.nfreq; h.npts = d.npts; h.stride = d.stride;
        h.nlines = d.lines.size(); h.acc_len = d.acc_len;
        std::string p; app(p, h); app_vec(p, d.freq); app_vec(p, d.lines);
        put_rec(out, "DFT", p);
    }
}

} // namespace chiml_host
