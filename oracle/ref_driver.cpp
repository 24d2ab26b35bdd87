// TEST INFRASTRUCTURE ONLY (oracle/): driver around the UNMODIFIED reference propagator
// (parallelFDTDFieldReal, compiled in place from /root/reference by oracle/Makefile against the shim
// headers in oracle/ref_shim).  It does what the reference's src/main.cpp:11-128 does for real
// fields -- strip comments, parse the JSON, construct the propagator, call step() nSteps times, write
// detector / flux / population outputs -- and additionally (a) runs R in-process "MPI ranks" as
// threads (the shim's communicator), (b) can stop after a given number of steps, (c) dumps every
// public field grid of every rank to a binary file so tests can compare full state, and (d) prints
// wall-clock seconds of the step loop as JSON.  Used to pin the restated oracle and as the
// `--impl reference` CPU arm of bench.py.
//
// usage: chiml_ref <input.json> [--ranks R] [--steps N] [--warmup W (untimed steps before the N timed ones)] [--dump FILE] [--plan PREFIX] [--no-output] [--quiet]
//   --plan PREFIX writes PREFIX.rank<r>.plan (include/chiml_plan.h) from the constructed propagator, before stepping
//
// dump file layout (little endian): magic "CHIMLDMP" | int32 nranks | then per rank, per grid:
//   int32 rank | char name[16] | int32 lnx, lny, lnz | int32 yStart(global row of local row 1) | float64 data[lnx*lny*lnz]
// with the reference's own index order x + lnx*(z + lnz*y), ghost cells included.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>
#include <boost/mpi.hpp>
#include <unordered_map>
#include <functional>
#include <iomanip>
// The plan dump reads the propagator's protected update lists; the reference offers no accessor for
// them, so this TEST driver opens the classes up (the reference sources themselves stay untouched).
#define protected public
#define private public
#include <FDTD_MANAGER/parallelFDTDField.hpp>
#undef protected
#undef private
#include "../include/chiml_plan.h"

namespace mpi = boost::mpi;

struct Options
{
    std::string input;
    int ranks = 1;
    int steps = -1;
    int warmup = 0;
    std::string dump;
    std::string plan;
    bool output = true;
    bool quiet = false;
};

struct GridDump
{
    int rank;
    std::string name;
    int ln[3];
    int yStart;
    std::vector<double> data;
};

static std::mutex g_dumpMtx;
static std::vector<GridDump> g_dumps;
static double g_stepSeconds = 0.0;
static int g_nStepsRun = 0;
static long g_cells = 0;

static void grabGrid(int rank, const std::string& name, std::shared_ptr<parallelGrid<double>> g)
{
    if(!g) return;
    GridDump d;
    d.rank = rank;
    d.name = name;
    d.ln[0] = g->local_x(); d.ln[1] = g->local_y(); d.ln[2] = g->local_z();
    d.yStart = g->procLoc(1);
    d.data.assign(&g->point(0), &g->point(0) + g->size());
    std::lock_guard<std::mutex> lk(g_dumpMtx);
    g_dumps.push_back(std::move(d));
}


// ---------------------------------------------------------------------------------------------
// plan dump (include/chiml_plan.h) from the reference's own data structures
// ---------------------------------------------------------------------------------------------
static void putRec(std::ofstream& out, const char* tag, const std::string& payload)
{
    char t[8];
    std::memset(t, ' ', 8);
    std::memcpy(t, tag, std::min<size_t>(8, std::strlen(tag)));
    uint64_t n = payload.size();
    out.write(t, 8);
    out.write(reinterpret_cast<const char*>(&n), 8);
    out.write(payload.data(), std::streamsize(n));
}
template <typename T> static void app(std::string& s, const T& v) { s.append(reinterpret_cast<const char*>(&v), sizeof(T)); }
template <typename T> static void appVec(std::string& s, const std::vector<T>& v) { if(!v.empty()) s.append(reinterpret_cast<const char*>(v.data()), v.size() * sizeof(T)); }

static void putList(std::ofstream& out, int kind, int comp, const upLists& l)
{
    static_assert(sizeof(upLists::value_type) == sizeof(ChimlRun), "upLists entry must be layout-identical to ChimlRun");
    std::string p;
    ChimlPlanListHdr h; h.kind = kind; h.comp = comp; h.n = l.size();
    app(p, h);
    if(!l.empty()) p.append(reinterpret_cast<const char*>(l.data()), l.size() * sizeof(ChimlRun));
    putRec(out, "UPLIST", p);
}

static void putCpml(std::ofstream& out, int comp, std::shared_ptr<parallelCPML<double>> pml)
{
    static_assert(sizeof(updatePsiParams) == sizeof(ChimlPsiParams), "updatePsiParams layout");
    static_assert(sizeof(updateGridParams) == sizeof(ChimlGridParams), "updateGridParams layout");
    if(!pml) return;
    for(int part = 0; part < 2; ++part)
    {
        const std::vector<updatePsiParams>&  psi  = part == 0 ? pml->updateListPsi_j_  : pml->updateListPsi_k_;
        const std::vector<updateGridParams>& grid = part == 0 ? pml->updateListGrid_k_ : pml->updateListGrid_j_;
        bool hasPsi  = part == 0 ? bool(pml->psi_j_)  : bool(pml->psi_k_);
        bool hasGrid = part == 0 ? bool(pml->grid_k_) : bool(pml->grid_j_);
        if(!hasGrid) continue;
        std::string p;
        ChimlPlanCpmlHdr h; h.comp = comp; h.part = part; h.has_psi = hasPsi ? 1 : 0; h.pad = 0; h.npsi = psi.size(); h.ngrid = grid.size();
        app(p, h);
        if(!psi.empty())  p.append(reinterpret_cast<const char*>(psi.data()),  psi.size()  * sizeof(ChimlPsiParams));
        if(!grid.empty()) p.append(reinterpret_cast<const char*>(grid.data()), grid.size() * sizeof(ChimlGridParams));
        putRec(out, "CPML", p);
    }
}

static int fieldId(parallelFDTDFieldReal& FF, const std::shared_ptr<parallelGrid<double>>& g)
{
    for(int i = 0; i < 3; ++i)
    {
        if(g && g == FF.E_[i]) return CHIML_EX + i;
        if(g && g == FF.H_[i]) return CHIML_HX + i;
        if(g && g == FF.D_[i]) return CHIML_DX + i;
    }
    return -1;
}

static void putEmitters(std::ofstream& out, parallelFDTDFieldReal& FF);
static void writePlan(const std::string& fname, parallelFDTDFieldReal& FF, const parallelProgramInputs& IP, int nSteps)
{
    std::ofstream out(fname.c_str(), std::ios::binary);
    { std::string p; int32_t v = CHIML_PLAN_VERSION; app(p, v); putRec(out, "CHIMLPLN", p); }
    std::shared_ptr<parallelGrid<double>> g0 = FF.E_[0] ? FF.E_[0] : FF.E_[2];
    ChimlPlanGrid pg;
    std::memset(&pg, 0, sizeof(pg));
    pg.desc.mode = (FF.E_[0] && FF.E_[2]) ? CHIML_MODE_3D : (FF.E_[0] ? CHIML_MODE_TE : CHIML_MODE_TM);
    for(int i = 0; i < 3; ++i) { pg.desc.ln[i] = g0->ln_vec(i); pg.desc.d[i] = FF.d_[i]; pg.n_global[i] = FF.n_vec_[i]; }
    pg.desc.dt = FF.dt_;
    pg.desc.has_D = (FF.D_[0] || FF.D_[2]) ? 1 : 0;
    pg.desc.pml_on_D = FF.dielectricMatInPML_ ? 1 : 0;
    pg.desc.n_objects = int(FF.objArr_.size());
    pg.desc.rank = FF.gridComm_->rank();
    pg.desc.nranks = FF.gridComm_->size();
    pg.y_start = g0->procLoc(1);
    pg.n_steps = nSteps;
    pg.n_lor_poles = int(std::max(FF.lorP_[0].size(), FF.lorP_[2].size()));
    pg.n_ordip_poles = int(std::max(FF.orDipLorP_[0].size(), FF.orDipLorP_[2].size()));
    pg.t_max = IP.tMax_;
    { std::string p; app(p, pg); putRec(out, "GRID", p); }

    if(FF.magMatInPML_) throw std::runtime_error("plan dump: magnetic material in the PML is outside the covered hot path");
    for(int c = 0; c < 3; ++c)
    {
        if(!FF.upB_[c].empty() || !FF.upLorB_[c].empty() || !FF.upChiD_[c].empty() || !FF.upChiB_[c].empty() || !FF.upOrDipB_[c].empty()
           || !FF.upOrDipChiD_[c].empty() || !FF.upOrDipChiB_[c].empty())
            throw std::runtime_error("plan dump: magnetic / chiral update lists are outside the covered hot path");
        putList(out, CHIML_LIST_U, c, FF.upE_[c]);
        putList(out, CHIML_LIST_U, 3 + c, FF.upH_[c]);
        putList(out, CHIML_LIST_D, c, FF.upD_[c]);
        putList(out, CHIML_LIST_LORD, c, FF.upLorD_[c]);
        putList(out, CHIML_LIST_ORDIPD, c, FF.upOrDipD_[c]);
    }
    if(!FF.upOrDipM_.empty() || !FF.upChiOrDipP_.empty() || !FF.upChiOrDipM_.empty())
        throw std::runtime_error("plan dump: magnetic / chiral oriented-dipole lists are outside the covered hot path");
    putList(out, CHIML_LIST_ORDIPP, 0, FF.upOrDipP_);

    for(size_t oo = 0; oo < FF.objArr_.size(); ++oo)
    {
        auto& obj = FF.objArr_[oo];
        ChimlPlanObjectHdr h;
        h.obj = int(oo); h.npoles = int(obj->gamma().size()); h.use_or_dip = obj->useOrdDip() ? 1 : 0; h.ml = obj->ML() ? 1 : 0;
        h.eps_inf = obj->epsInfty(); h.mu_inf = obj->muInfty();
        std::string p; app(p, h);
        appVec(p, obj->alpha()); appVec(p, obj->xi()); appVec(p, obj->gamma());
        std::vector<double> dip(3 * size_t(h.npoles), 0.0);
        if(h.use_or_dip)
        {
            for(int pp = 0; pp < h.npoles; ++pp)
            {
                MAT_DIP_ORIENTAITON ori = obj->dipOr(pp);
                if(ori == MAT_DIP_ORIENTAITON::ISOTROPIC) { dip[3*pp] = dip[3*pp+1] = dip[3*pp+2] = 1.0; }
                else if(ori == MAT_DIP_ORIENTAITON::UNIDIRECTIONAL) { for(int k = 0; k < 3; ++k) dip[3*pp+k] = obj->dipE(pp)[k]; }
                else throw std::runtime_error("plan dump: position-dependent dipole orientation (REL_TO_NORM) is outside the covered hot path");
            }
        }
        appVec(p, dip);
        putRec(out, "OBJECT", p);
    }
    for(int c = 0; c < 3; ++c)
    {
        putCpml(out, c, FF.EPML_[c]);
        putCpml(out, 3 + c, FF.HPML_[c]);
    }
    // sources: the box this rank adds to, and the per-step amplitude dt*Re(sum pulse(t_k)) with t_k accumulated as step() does
    for(auto& srcBase : FF.srcArr_)
    {
        auto src = std::dynamic_pointer_cast<parallelSourceNormalReal>(srcBase);
        if(!src) throw std::runtime_error("plan dump: only normal (axis-aligned) soft sources are on the covered hot path");
        if(!src->slave_) continue;
        ChimlPlanSourceHdr h;
        h.field = fieldId(FF, src->grid_);
        // undo the (length, trans1, trans2) permutation of genDatStruct so loc/sz are plain x,y,z
        std::array<int,3> sz = {{1, 1, 1}};
        const SalveSource& sl = *src->slave_;
        int ax1 = sl.addVec1_[0] ? 0 : (sl.addVec1_[1] ? 1 : 2);
        int ax2 = sl.addVec2_[0] ? 0 : (sl.addVec2_[1] ? 1 : 2);
        int longAxis = 3 - ax1 - ax2;
        sz[longAxis] = sl.sz_[0];
        sz[ax1] = sl.sz_[1];
        sz[ax2] = sl.sz_[2];
        for(int k = 0; k < 3; ++k) { h.loc[k] = sl.loc_[k]; h.sz[k] = sz[k]; }
        h.n_steps = nSteps;
        std::vector<double> amp(nSteps);
        double t = 0.0;
        for(int k = 0; k < nSteps; ++k)
        {
            cplx pulVal = 0.0;
            for(auto& pul : src->pulse_) pulVal += pul->pulse(t);
            amp[k] = FF.dt_ * std::real(pulVal);
            t += FF.dt_;
        }
        std::string p; app(p, h); appVec(p, amp);
        putRec(out, "SOURCE", p);
    }
    int dd = 0;
    for(auto& dtc : FF.dtcArr_)
    {
        for(auto& f : dtc->fields_)
        {
            ChimlPlanDetector d;
            std::memset(&d, 0, sizeof(d));
            d.detector = dd;
            d.field = fieldId(FF, f->grid_);
            for(int k = 0; k < 3; ++k) { d.loc[k] = f->loc_[k]; d.sz[k] = f->sz_[k]; d.offset[k] = f->offSet_[k]; }
            d.every = dtc->timeInterval_;
            d.type = int(dtc->type_);
            d.conv = dtc->convFactor_;
            d.t_conv = dtc->tConv_;
            std::string p; app(p, d);
            putRec(out, "DETECTOR", p);
        }
        ++dd;
    }
    if(!FF.qeArr_.empty() && FF.gridComm_->size() == 1) putEmitters(out, FF);
    // DFT records: every stored field of every flux object, in the order parallelFluxDTC::fieldIn walks them (DTC/parallelFlux.hpp:296-312)
    int group = 0;
    for(auto& flux : FF.fluxArr_)
    {
        for(auto& fp : flux->fInParam_)
            for(auto* vec : {&fp.Ej_dtc_, &fp.Ek_dtc_, &fp.Hj_dtc_, &fp.Hk_dtc_})
                for(auto& dtc : *vec)
                {
                    auto real = std::dynamic_pointer_cast<parallelStorageFreqDTCReal>(dtc);
                    if(!real || !real->fieldInFreq_) continue;
                    ChimlPlanDftHdr h; std::memset(&h, 0, sizeof(h));
                    h.field = fieldId(FF, real->grid_); h.group = group; h.every = flux->timeInt_; h.nfreq = real->nfreq_;
                    h.npts = real->fieldInFreq_->sz_[0]; h.stride = real->fieldInFreq_->stride_;
                    h.nlines = real->fieldInFreq_->fInGridInds_.size() / 2; h.acc_len = real->fInReal_.size();
                    std::string p; app(p, h); appVec(p, flux->freqList_);
                    for(size_t ii = 0; ii + 1 < real->fieldInFreq_->fInGridInds_.size(); ii += 2)
                    { ChimlDftLine l; l.ind = real->fieldInFreq_->fInGridInds_[ii]; l.out = real->fieldInFreq_->fInGridInds_[ii + 1]; app(p, l); }
                    putRec(out, "DFT", p);
                }
        ++group;
    }
}

// EMITTER records: one per parallelQE object, from the object's own members (single-rank runs only: with more ranks the
// reference spreads emitters over ranks irrespective of where their node lives, ML/parallelQE.hpp:394-420)
static void putEmitters(std::ofstream& out, parallelFDTDFieldReal& FF)
{
    int qq = 0;
    for(auto& qe : FF.qeArr_)
    {
        if(!qe->sameProcCalc_) throw std::runtime_error("plan dump: emitter sets need a single-rank run");
        ChimlPlanEmitterHdr h;
        std::memset(&h, 0, sizeof(h));
        h.object = qq;
        h.nlevel = qe->nlevel_;
        h.nsys = int(qe->levelSys_.size());
        h.nemit = int(qe->levelSys_[0].den_.size());
        auto eg = qe->e_[0] ? qe->e_[0] : qe->e_[2];
        auto Pg = qe->P_[0] ? qe->P_[0] : qe->P_[2];
        h.box_n[0] = eg->x(); h.box_n[1] = eg->y(); h.box_n[2] = eg->z();
        for(int k = 0; k < 3; ++k) h.box_lo[k] = qe->sameProcCalc_->loc_[k];
        h.pz = Pg->z();
        h.dt = qe->dt_; h.inv_hbar = std::imag(qe->one_over_hbar_); h.na = qe->na_;
        std::vector<int32_t> gptr(1, 0), gcol; std::vector<double> gval;
        for(auto& row : qe->gam_)
        {
            for(auto it = row.begin(); it != row.end(); ++it) { gcol.push_back(it->first); gval.push_back(it->second); }
            gptr.push_back(int32_t(gcol.size()));
        }
        while(int(gptr.size()) < h.nlevel * h.nlevel + 1) gptr.push_back(gptr.back());
        h.nnz = int(gcol.size());
        h.npop = int(qe->dtcPopArr_.size());
        h.pop_every = h.npop ? qe->dtcPopArr_[0]->timeInt_ : 1;
        h.npoints = h.npop ? qe->dtcPopArr_[0]->npoints_ : h.nemit;
        std::string p; app(p, h);
        const int n2 = h.nlevel * h.nlevel;
        for(auto& ls : qe->levelSys_) p.append(reinterpret_cast<const char*>(ls.ham_->h0_.data()), n2 * sizeof(cplx));
        for(auto& ew : qe->energyWeights_) app(p, ew.second);
        auto& ham = *qe->levelSys_[0].ham_;
        p.append(reinterpret_cast<const char*>(ham.x_expectation_.data()), n2 * sizeof(cplx));
        p.append(reinterpret_cast<const char*>(ham.y_expectation_.data()), n2 * sizeof(cplx));
        p.append(reinterpret_cast<const char*>(ham.z_expectation_.data()), n2 * sizeof(cplx));
        appVec(p, gptr); appVec(p, gcol); appVec(p, gval);
        for(auto& den : qe->levelSys_[0].den_) { int32_t l[3] = {den.x(), den.y(), den.z()}; p.append(reinterpret_cast<const char*>(l), sizeof(l)); }
        for(int y = 0; y < Pg->y(); ++y)
            for(int z = 0; z < Pg->z(); ++z)
                for(int x = 0; x < Pg->x(); ++x)
                {
                    double e = qe->eps_->z() == 1 ? qe->eps_->point(h.box_lo[0] + x, h.box_lo[1] + y, 0)
                                                  : qe->eps_->point(h.box_lo[0] + x, h.box_lo[1] + y, h.box_lo[2] + z);
                    app(p, e);
                }
        for(auto& d : qe->dtcPopArr_) { int32_t lv = d->level_; app(p, lv); }
        putRec(out, "EMITTER", p);
        ++qq;
    }
}

static void rankMain(int rank, const Options& opt)
{
    mpi::shim::myRank() = rank;
    std::shared_ptr<mpiInterface> gridComm = std::make_shared<mpiInterface>();
    std::string filename = opt.input;
    if(gridComm->rank() == 0)
        stripComments(filename);
    else
    {
        // mirror of what stripComments does to the name on rank 0
        std::string::size_type slash = filename.find_last_of('/');
        if(slash == std::string::npos) filename = "stripped_" + filename;
        else filename = filename.substr(0, slash + 1) + "stripped_" + filename.substr(slash + 1);
    }
    gridComm->barrier();
    boost::property_tree::ptree propTree;
    boost::property_tree::json_parser::read_json(filename, propTree);
    parallelProgramInputs IP(propTree, filename);
    gridComm->barrier();
    if(gridComm->rank() == 0)
        boost::filesystem::remove(filename);

    if(IP.cplxFields_)
        throw std::runtime_error("chiml_ref: complex-field runs are outside the hot path covered here");

    parallelFDTDFieldReal FF(IP, gridComm);
    int nSteps = int(std::ceil(IP.tMax_ / IP.dt_));
    if(opt.steps >= 0) nSteps = opt.steps;

    if(!opt.plan.empty())
        writePlan(opt.plan + ".rank" + std::to_string(rank) + ".plan", FF, IP, nSteps);

    for(int tt = 0; tt < opt.warmup; ++tt)
        FF.step();
    gridComm->barrier();
    auto t0 = std::chrono::steady_clock::now();
    for(int tt = 0; tt < nSteps; ++tt)
        FF.step();
    gridComm->barrier();
    auto t1 = std::chrono::steady_clock::now();
    if(rank == 0)
    {
        g_stepSeconds = std::chrono::duration<double>(t1 - t0).count();
        g_nStepsRun = nSteps;
        g_cells = long(FF.n_vec_[0]) * long(FF.n_vec_[1]) * long(FF.n_vec_[2] > 1 ? FF.n_vec_[2] : 1);   // grid points, PML included
    }

    if(!opt.dump.empty())
    {
        const char* c = "xyz";
        for(int i = 0; i < 3; ++i)
        {
            grabGrid(rank, std::string("E") + c[i], FF.E_[i]);
            grabGrid(rank, std::string("H") + c[i], FF.H_[i]);
            grabGrid(rank, std::string("D") + c[i], FF.D_[i]);
            grabGrid(rank, std::string("B") + c[i], FF.B_[i]);
            for(size_t p = 0; p < FF.lorP_[i].size(); ++p)
            {
                grabGrid(rank, std::string("P") + c[i] + std::to_string(p), FF.lorP_[i][p]);
                grabGrid(rank, std::string("pP") + c[i] + std::to_string(p), FF.prevLorP_[i][p]);
            }
            for(size_t p = 0; p < FF.orDipLorP_[i].size(); ++p)
            {
                grabGrid(rank, std::string("oP") + c[i] + std::to_string(p), FF.orDipLorP_[i][p]);
                grabGrid(rank, std::string("poP") + c[i] + std::to_string(p), FF.prevOrDipLorP_[i][p]);
            }
        }
    }

    if(!opt.dump.empty() && FF.gridComm_->size() == 1)
    {
        // emitter state: rho and the four derivative histories per level system, the P boxes, the population series
        int qq = 0;
        for(auto& qe : FF.qeArr_)
        {
            const int n2 = qe->nlevel_ * qe->nlevel_;
            for(size_t ss = 0; ss < qe->levelSys_.size(); ++ss)
                for(int w = 0; w < 5; ++w)
                {
                    GridDump d; d.rank = rank; d.name = "q" + std::to_string(qq) + "s" + std::to_string(ss) + "w" + std::to_string(w);
                    d.ln[0] = 2 * n2; d.ln[1] = int(qe->levelSys_[ss].den_.size()); d.ln[2] = 1; d.yStart = 0;
                    for(auto& den : qe->levelSys_[ss].den_)
                    {
                        std::vector<cplx>& v = w == 0 ? den.density_ : w == 1 ? den.density_deriv_n_ : w == 2 ? den.density_deriv_n_minus_1_
                                             : w == 3 ? den.density_deriv_n_minus_2_ : den.density_deriv_n_minus_3_;
                        for(auto& c : v) { d.data.push_back(c.real()); d.data.push_back(c.imag()); }
                    }
                    std::lock_guard<std::mutex> lk(g_dumpMtx); g_dumps.push_back(std::move(d));
                }
            for(int c = 0; c < 3; ++c)
            {
                if(!qe->P_[c] || !qe->E_[c]) continue;
                GridDump d; d.rank = rank; d.name = "q" + std::to_string(qq) + "P" + std::string(1, "xyz"[c]);
                d.ln[0] = qe->P_[c]->x(); d.ln[1] = qe->P_[c]->y(); d.ln[2] = qe->P_[c]->z(); d.yStart = 0;
                d.data.assign(&qe->P_[c]->point(0), &qe->P_[c]->point(0) + qe->P_[c]->size());
                std::lock_guard<std::mutex> lk(g_dumpMtx); g_dumps.push_back(std::move(d));
            }
            int dd = 0;
            for(auto& dtc : qe->dtcPopArr_)
            {
                GridDump d; d.rank = rank; d.name = "q" + std::to_string(qq) + "pop" + std::to_string(dd++);
                d.ln[0] = 2; d.ln[1] = int(dtc->allPop_.size()); d.ln[2] = 1; d.yStart = 0;
                for(auto& c : dtc->allPop_) { d.data.push_back(c.real()); d.data.push_back(c.imag()); }
                std::lock_guard<std::mutex> lk(g_dumpMtx); g_dumps.push_back(std::move(d));
            }
            ++qq;
        }
    }

    if(!opt.dump.empty())
    {
        int slot = 0;
        for(auto& flux : FF.fluxArr_)
            for(auto& fp : flux->fInParam_)
                for(auto* vec : {&fp.Ej_dtc_, &fp.Ek_dtc_, &fp.Hj_dtc_, &fp.Hk_dtc_})
                    for(auto& dtc : *vec)
                    {
                        auto real = std::dynamic_pointer_cast<parallelStorageFreqDTCReal>(dtc);
                        if(!real || !real->fieldInFreq_) continue;
                        for(int im = 0; im < 2; ++im)
                        {
                            GridDump d; d.rank = rank; d.name = "dft" + std::to_string(slot) + (im ? "i" : "r");
                            const std::vector<double>& v = im ? real->fInCplx_ : real->fInReal_;
                            d.ln[0] = int(v.size()); d.ln[1] = 1; d.ln[2] = 1; d.yStart = 0;
                            d.data = v;
                            std::lock_guard<std::mutex> lk(g_dumpMtx); g_dumps.push_back(std::move(d));
                        }
                        ++slot;
                    }
    }

    if(opt.output)
    {
        for(auto& flux : FF.fluxArr())
            flux->getFlux(FF.ExIncd(), FF.EyIncd(), FF.EzIncd(), FF.HxIncd(), FF.HyIncd(), FF.HzIncd(), true);
        for(auto& dtc : FF.dtcFreqArr())
        {
            if(dtc->outputMaps()) dtc->toMap();
            else dtc->toFile();
        }
        for(auto& dtc : FF.dtcArr())
            dtc->toFile();
        for(auto& qe : FF.qeArr())
        {
            if(qe->pAccuulate())
                qe->outputPol();
            for(auto& dtcPop : qe->dtcPopArr())
                dtcPop->toFile();
        }
    }
    gridComm->barrier();
}

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void segvHandler(int sig)
{
    void* frames[64];
    int n = backtrace(frames, 64);
    const char msg[] = "chiml_ref: fatal signal, backtrace:\n";
    (void)!write(2, msg, sizeof(msg) - 1);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

int main(int argc, char** argv)
{
    signal(SIGSEGV, segvHandler);
    signal(SIGABRT, segvHandler);
    Options opt;
    for(int a = 1; a < argc; ++a)
    {
        std::string s = argv[a];
        if(s == "--ranks" && a + 1 < argc) opt.ranks = std::atoi(argv[++a]);
        else if(s == "--steps" && a + 1 < argc) opt.steps = std::atoi(argv[++a]);
        else if(s == "--warmup" && a + 1 < argc) opt.warmup = std::atoi(argv[++a]);
        else if(s == "--dump" && a + 1 < argc) opt.dump = argv[++a];
        else if(s == "--plan" && a + 1 < argc) opt.plan = argv[++a];
        else if(s == "--no-output") opt.output = false;
        else if(s == "--quiet") opt.quiet = true;
        else if(opt.input.empty()) opt.input = s;
        else { std::fprintf(stderr, "chiml_ref: unknown argument %s\n", s.c_str()); return 2; }
    }
    if(opt.input.empty() || opt.ranks < 1)
    {
        std::fprintf(stderr, "usage: chiml_ref <input.json> [--ranks R] [--steps N] [--dump FILE] [--plan PREFIX] [--no-output] [--quiet]\n");
        return 2;
    }
    // --quiet: the reference prints from every rank; ranks are threads here, so the sink must be stateless (a shared
    // std::ostringstream is a data race that crashes with many ranks)
    struct NullBuf : std::streambuf
    {
        int overflow(int c) override { return traits_type::not_eof(c); }
        std::streamsize xsputn(const char*, std::streamsize n) override { return n; }
    };
    static NullBuf sink;
    std::streambuf* oldCout = nullptr;
    if(opt.quiet) oldCout = std::cout.rdbuf(&sink);

    mpi::shim::world().nranks = opt.ranks;
    std::vector<std::thread> threads;
    std::vector<std::string> errors(opt.ranks);
    for(int r = 0; r < opt.ranks; ++r)
        threads.emplace_back([r, &opt, &errors]() {
            try { rankMain(r, opt); }
            catch(std::exception& e) { errors[r] = e.what(); std::fprintf(stderr, "chiml_ref rank %d: %s\n", r, e.what()); std::_Exit(3); }
        });
    for(auto& t : threads) t.join();
    if(oldCout) std::cout.rdbuf(oldCout);

    if(!opt.dump.empty())
    {
        std::ofstream out(opt.dump.c_str(), std::ios::binary);
        out.write("CHIMLDMP", 8);
        int32_t nr = opt.ranks;
        out.write(reinterpret_cast<const char*>(&nr), 4);
        for(auto& d : g_dumps)
        {
            int32_t rk = d.rank;
            char name[16];
            std::memset(name, 0, sizeof(name));
            std::strncpy(name, d.name.c_str(), 15);
            int32_t ln[3] = { d.ln[0], d.ln[1], d.ln[2] };
            int32_t ys = d.yStart;
            out.write(reinterpret_cast<const char*>(&rk), 4);
            out.write(name, 16);
            out.write(reinterpret_cast<const char*>(ln), 12);
            out.write(reinterpret_cast<const char*>(&ys), 4);
            out.write(reinterpret_cast<const char*>(d.data.data()), std::streamsize(d.data.size() * sizeof(double)));
        }
    }
    std::printf("{\"impl\": \"chiml_ref\", \"ranks\": %d, \"steps\": %d, \"cells\": %ld, \"step_seconds\": %.9g}\n", opt.ranks, g_nStepsRun, g_cells, g_stepSeconds);
    return 0;
}
