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
// usage: chiml_ref <input.json> [--ranks R] [--steps N] [--warmup W (untimed steps before the N timed ones)] [--dump FILE] [--plan PREFIX] [--no-output] [--quiet] [--gpu]
//   --plan PREFIX writes PREFIX.rank<r>.plan (include/chiml_plan.h) from the constructed propagator, before stepping
//   --gpu         THE DROP-IN, COMPILED: the unmodified reference constructs everything, its lists go straight to the C ABI of
//                 chiml_b200/libchiml_b200.so (loaded with dlopen: this binary stays runnable without CUDA), the time loop runs on the
//                 GPU, and the reference's own writers (dtc->output / toFile, flux->getFlux, dtcPop->toFile) produce the files.  This
//                 is the bindGpu() / step() stub of INTEGRATION.md as code (reference FDTD_MANAGER/parallelFDTDField.hpp:1228-1303,
//                 main.cpp:54-118).  Single rank: the ranks of this driver are threads of one process, and CUDA IPC -- what the slabs'
//                 halo binds with -- needs one process per slab.
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
#include <dlfcn.h>
#include <unistd.h>

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
    bool gpu = false;
    std::string incdDump;      // --incd-dump FILE: the incident-field series E_incd_[0..2], H_incd_[0..2] after the run ("CHIMLINC", int32 n, 6 x n complex)
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

// complex grids: the real part under the grid's name, the imaginary part under <name>_im
static void grabGrid(int rank, const std::string& name, std::shared_ptr<parallelGrid<cplx>> g)
{
    if(!g) return;
    for(int im = 0; im < 2; ++im)
    {
        GridDump d;
        d.rank = rank;
        d.name = name + (im ? "_im" : "");
        d.ln[0] = g->local_x(); d.ln[1] = g->local_y(); d.ln[2] = g->local_z();
        d.yStart = g->procLoc(1);
        d.data.resize(size_t(g->size()));
        for(int i = 0; i < g->size(); ++i) d.data[size_t(i)] = im ? std::imag(g->point(i)) : std::real(g->point(i));
        std::lock_guard<std::mutex> lk(g_dumpMtx);
        g_dumps.push_back(std::move(d));
    }
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

template <class T> static void putCpml(std::ofstream& out, int comp, std::shared_ptr<parallelCPML<T>> pml)
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

template <class FFT, class T> static int fieldId(FFT& FF, const std::shared_ptr<parallelGrid<T>>& g)
{
    for(int i = 0; i < 3; ++i)
    {
        if(g && g == FF.E_[i]) return CHIML_EX + i;
        if(g && g == FF.H_[i]) return CHIML_HX + i;
        if(g && g == FF.D_[i]) return CHIML_DX + i;
    }
    return -1;
}

// the arguments step() passes to applBCH_[c] / applBCE_[c] (FDTD_MANAGER/parallelFDTDField.hpp:1267-1269,1285-1287), comp 0..5 = Ex..Hz
template <class FFT> static ChimlWrap wrapArgs(FFT& FF, int comp)
{
    const int l0 = FF.ln_vec_[0], zMin = FF.zMinPBC_, zMax = FF.zMaxPBC_;
    switch(comp)
    {
        case 0: return ChimlWrap{l0 - 1, FF.yEPBC_[0], zMax,     l0,     FF.yEPBC_[0], zMin, zMax + 1};
        case 1: return ChimlWrap{l0,     FF.yEPBC_[1], zMax,     l0 + 1, FF.yEPBC_[1], zMin, zMax + 1};
        case 2: return ChimlWrap{l0,     FF.yEPBC_[2], zMax - 1, l0 + 1, FF.yEPBC_[2], zMin, zMax};
        case 3: return ChimlWrap{l0,     FF.yHPBC_[0], zMax - 1, l0 + 1, FF.yHPBC_[0], zMin, zMax};
        case 4: return ChimlWrap{l0 - 1, FF.yHPBC_[1], zMax - 1, l0,     FF.yHPBC_[1], zMin, zMax};
        default: return ChimlWrap{l0 - 1, FF.yHPBC_[2], zMax,    l0,     FF.yHPBC_[2], zMin, zMax + 1};
    }
}

// ---------------------------------------------------------------------------------------------
// TFSF sources: the surfaces of every parallelTFSF object as ChimlTfsfSurface records and the layout of one step's incident-line
// table (include/chiml_gpu.h chiml_gpu_step_n_tfsf): per TFSF object E_incd_[0..2] then H_incd_[0..2], real parts
// ---------------------------------------------------------------------------------------------
struct TfsfLayout
{
    std::vector<int> off, len;      // [6 * t + k], k = 0..2 E lines, 3..5 H lines
    int perStep = 0;
};
static TfsfLayout tfsfLayout(parallelFDTDFieldReal& FF)
{
    TfsfLayout L;
    for(auto& tfsf : FF.tfsfArr_)
        for(int k = 0; k < 6; ++k)
        {
            auto& g = k < 3 ? tfsf->E_incd_[k] : tfsf->H_incd_[k - 3];
            L.off.push_back(L.perStep);
            L.len.push_back(g ? int(g->size()) : 0);
            L.perStep += L.len.back();
        }
    return L;
}
// the E lines (what the H surfaces of the coming step read) or the H lines (what its E surfaces read, after the line's own step)
static void tfsfGrab(parallelFDTDFieldReal& FF, const TfsfLayout& L, bool E, double* row)
{
    for(size_t t = 0; t < FF.tfsfArr_.size(); ++t)
        for(int k = E ? 0 : 3; k < (E ? 3 : 6); ++k)
        {
            auto& g = k < 3 ? FF.tfsfArr_[t]->E_incd_[k] : FF.tfsfArr_[t]->H_incd_[k - 3];
            for(int i = 0; i < L.len[6 * t + k]; ++i) row[L.off[6 * t + k] + i] = std::real(g->point(i));
        }
}
struct TfsfSurfaceRec { ChimlTfsfSurface s; std::vector<double> epMu; };
static std::vector<TfsfSurfaceRec> tfsfSurfaces(parallelFDTDFieldReal& FF, const TfsfLayout& L)
{
    std::vector<TfsfSurfaceRec> out;
    for(size_t t = 0; t < FF.tfsfArr_.size(); ++t)
    {
        auto& tfsf = FF.tfsfArr_[t];
        for(int side = 0; side < 2; ++side)          // updateFields(): H surfaces, then (after the line's step) E surfaces
            for(int ii = 0; ii < 3; ++ii)
            {
                const bool E = side == 1;
                auto& epMu = E ? tfsf->eps_[ii] : tfsf->mu_[ii];
                // the constructor's choice between addIncdFields and addIncdFieldsEPChange (SOURCE/parallelTFSF.cpp:125-200)
                const bool epChange = (E ? bool(tfsf->E_[ii]) : bool(tfsf->H_[ii])) && epMu &&
                                      std::any_of(epMu->data(), epMu->data() + epMu->size(), [&](double a) { return a != epMu->point(0); });
                for(auto& sur : (E ? tfsf->eSurfaces_[ii] : tfsf->hSurfaces_[ii]))
                {
                    TfsfSurfaceRec r;
                    std::memset(&r.s, 0, sizeof(r.s));
                    int k = -1;
                    for(int q = 0; q < 3; ++q)
                    {
                        if(sur->incdField_ == tfsf->E_incd_[q]) k = q;
                        if(sur->incdField_ == tfsf->H_incd_[q]) k = 3 + q;
                    }
                    if(k < 0) throw std::runtime_error("TFSF surface: unknown incident field");
                    if(E != (k >= 3)) throw std::runtime_error("TFSF surface: an E surface reads an incident E line (or H / H)");
                    if(!E && !sur->indsD_.empty()) throw std::runtime_error("TFSF surface inside a magnetic-dispersive medium (B target) is outside the covered hot path");
                    r.s.comp = (E ? 0 : 3) + ii;
                    r.s.incd_offset = L.off[6 * t + k]; r.s.incd_len = L.len[6 * t + k];
                    r.s.n = sur->szTrans_[0]; r.s.stride_incd = sur->strideIncd_; r.s.stride_main = sur->strideMain_;
                    r.s.npairs_D = int(sur->indsD_.size() / 2); r.s.npairs_U = int(sur->indsU_.size() / 2);
                    r.s.prefactor = sur->prefactor_;
                    r.s.pairs_D = sur->indsD_.data(); r.s.pairs_U = sur->indsU_.data();       // owned by the reference's surface object
                    if(epChange) { r.epMu.assign(epMu->data(), epMu->data() + epMu->size()); r.epMu.resize(size_t(r.s.incd_len), 1.0); }
                    out.push_back(std::move(r));
                }
            }
    }
    for(auto& r : out) r.s.ep_mu = r.epMu.empty() ? nullptr : r.epMu.data();
    return out;
}
static std::vector<double> g_tfsfTable;          // rank 0's table rows of the steps taken (appended to the plan after the run)
static int g_tfsfPerStep = 0;

// every running-DFT storage of the propagator in the order step() feeds them (FDTD_MANAGER/parallelFDTDField.hpp:1297-1302): the stored fields
// of the frequency detectors (dtcFreqArr_, DTC/parallelDTC_FREQ.hpp:258-264) after those of the flux objects (DTC/parallelFlux.hpp:296-312);
// group = which twiddle list a storage uses: flux objects first, then frequency detectors
struct DftStorageRef { std::shared_ptr<parallelStorageFreqDTCReal> st; int group; int every; const std::vector<double>* freq; };
static std::vector<DftStorageRef> allDftStorages(parallelFDTDFieldReal& FF)
{
    std::vector<DftStorageRef> out;
    int group = 0;
    for(auto& flux : FF.fluxArr_)
    {
        for(auto& fp : flux->fInParam_)
            for(auto* vec : {&fp.Ej_dtc_, &fp.Ek_dtc_, &fp.Hj_dtc_, &fp.Hk_dtc_})
                for(auto& dtc : *vec)
                {
                    auto real = std::dynamic_pointer_cast<parallelStorageFreqDTCReal>(dtc);
                    if(real && real->fieldInFreq_) out.push_back({real, group, flux->timeInt_, &flux->freqList_});
                }
        ++group;
    }
    for(auto& dtc : FF.dtcFreqArr_)
    {
        for(auto& g : dtc->gridsIn_)
        {
            auto real = std::dynamic_pointer_cast<parallelStorageFreqDTCReal>(g);
            if(real && real->fieldInFreq_) out.push_back({real, group, dtc->timeInt_, &dtc->freqList_});
        }
        ++group;
    }
    return out;
}
// the frequency lists of the groups that hold a storage on this rank, in group order (the twiddles of one step follow this order)
static std::vector<const std::vector<double>*> dftGroupFreqs(parallelFDTDFieldReal& FF)
{
    std::vector<const std::vector<double>*> out;
    int last = -1;
    for(const DftStorageRef& r : allDftStorages(FF)) if(r.group != last) { out.push_back(r.freq); last = r.group; }
    return out;
}

static void putEmitters(std::ofstream& out, parallelFDTDFieldReal& FF);
// real / complex propagator: the value type of its grids and the class of its soft sources
template <class FFT> struct FieldTraits;
template <> struct FieldTraits<parallelFDTDFieldReal> { typedef double value; typedef parallelSourceNormalReal source; static const bool cplx = false; };
template <> struct FieldTraits<parallelFDTDFieldCplx> { typedef cplx value;   typedef parallelSourceNormalCplx source; static const bool cplx = true; };
static void putTfsfRecords(std::ofstream& out, parallelFDTDFieldReal& FF);
static void putTfsfRecords(std::ofstream&, parallelFDTDFieldCplx& FF)
{ if(!FF.tfsfArr_.empty()) throw std::runtime_error("plan dump: TFSF sources with complex fields are outside the covered hot path"); }
static void putEmittersAndDfts(std::ofstream& out, parallelFDTDFieldReal& FF);
static void putEmittersAndDfts(std::ofstream&, parallelFDTDFieldCplx& FF)
{
    if(!FF.qeArr_.empty() || !FF.fluxArr_.empty() || !FF.dtcFreqArr_.empty())
        throw std::runtime_error("plan dump: emitters / flux regions / frequency detectors with complex fields are outside the covered hot path");
}
template <class FFT> static void writePlan(const std::string& fname, FFT& FF, const parallelProgramInputs& IP, int nSteps)
{
    typedef FieldTraits<FFT> TR;
    std::ofstream out(fname.c_str(), std::ios::binary);
    { std::string p; int32_t v = CHIML_PLAN_VERSION; app(p, v); putRec(out, "CHIMLPLN", p); }
    std::shared_ptr<parallelGrid<typename TR::value>> g0 = FF.E_[0] ? FF.E_[0] : FF.E_[2];
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
    if(TR::cplx)
    {
        // complex fields (Bloch-periodic runs): every field array has a real and an imaginary part; the wrap copies carry the phase
        // factors of the k-point (UTIL/FDTD_up_eq.cpp:1118-1324)
        if(!IP.periodic_) throw std::runtime_error("plan dump: complex fields without periodic boundaries are outside the covered hot path");
        ChimlPlanComplex pc; std::memset(&pc, 0, sizeof(pc));
        pc.cplx = 1; for(int k = 0; k < 3; ++k) pc.k_point[k] = FF.k_point_[k];
        std::string p; app(p, pc); putRec(out, "COMPLEX", p);
    }
    if(IP.periodic_)
    {
        if(FF.gridComm_->size() > 1) throw std::runtime_error("plan dump: periodic boundaries on several ranks are outside the covered hot path");
        for(int comp = 0; comp < 6; ++comp)
        {
            if(!(comp < 3 ? FF.E_[comp] : FF.H_[comp - 3])) continue;
            ChimlPlanPeriodic pp; pp.comp = comp; pp.wrap = wrapArgs(FF, comp);
            std::string p; app(p, pp); putRec(out, "PERIODIC", p);
        }
    }

    putTfsfRecords(out, FF);
    // magnetic-dispersive media: B grids, the H-side CPML on B, magnetic pole constants per object
    const bool hasB = bool(FF.B_[0]) || bool(FF.B_[2]);
    if(hasB)
    {
        ChimlPlanMagnetic pm; std::memset(&pm, 0, sizeof(pm));
        pm.has_B = 1; pm.pml_on_B = FF.magMatInPML_ ? 1 : 0; pm.n_mag_poles = int(std::max(FF.lorM_[0].size(), FF.lorM_[2].size()));
        std::string p; app(p, pm); putRec(out, "MAGNETIC", p);
        for(size_t oo = 0; oo < FF.objArr_.size(); ++oo)
        {
            auto& obj = FF.objArr_[oo];
            ChimlPlanObjMagHdr h; h.obj = int(oo); h.npoles = int(obj->magGamma().size());
            std::string q; app(q, h); appVec(q, obj->magAlpha()); appVec(q, obj->magXi()); appVec(q, obj->magGamma());
            putRec(out, "OBJMAG", q);
        }
    }
    else if(FF.magMatInPML_) throw std::runtime_error("plan dump: magnetic material in the PML without B grids");
    // chiral media: constants per object, the rows copied into prevE_ / prevH_
    {
        bool anyChi = false;
        for(auto& obj : FF.objArr_) anyChi = anyChi || !obj->chiGamma().empty();
        if(anyChi)
        {
            for(size_t oo = 0; oo < FF.objArr_.size(); ++oo)
            {
                auto& obj = FF.objArr_[oo];
                ChimlPlanObjChiHdr h; h.obj = int(oo); h.npoles = int(obj->chiGamma().size());
                std::string q; app(q, h); appVec(q, obj->chiAlpha()); appVec(q, obj->chiXi()); appVec(q, obj->chiGamma()); appVec(q, obj->chiGammaPrev());
                putRec(out, "OBJCHI", q);
            }
            std::string q; uint64_t nr = FF.copy2PrevFields_.size(); app(q, nr);
            for(auto& row : FF.copy2PrevFields_) for(int k = 0; k < 4; ++k) { int32_t v = row[k]; app(q, v); }
            putRec(out, "PREVCOPY", q);
        }
    }
    for(int c = 0; c < 3; ++c)
    {
        if(!FF.upOrDipB_[c].empty() || !FF.upOrDipChiD_[c].empty() || !FF.upOrDipChiB_[c].empty())
            throw std::runtime_error("plan dump: magnetic / chiral oriented-dipole update lists are outside the covered hot path");
        if(!FF.upChiD_[c].empty() || !FF.upChiB_[c].empty())
        {
            if(!hasB || !(FF.D_[0] && FF.D_[2])) throw std::runtime_error("plan dump: chiral lists need D and B grids on a 3-D grid");
            putList(out, CHIML_LIST_CHID, c, FF.upChiD_[c]);
            putList(out, CHIML_LIST_CHID, 3 + c, FF.upChiB_[c]);
        }
        if(hasB)
        {
            putList(out, CHIML_LIST_D, 3 + c, FF.upB_[c]);
            putList(out, CHIML_LIST_LORD, 3 + c, FF.upLorB_[c]);
        }
        else if(!FF.upB_[c].empty() || !FF.upLorB_[c].empty()) throw std::runtime_error("plan dump: magnetic update lists without B grids");
        putList(out, CHIML_LIST_U, c, FF.upE_[c]);
        putList(out, CHIML_LIST_U, 3 + c, FF.upH_[c]);
        putList(out, CHIML_LIST_D, c, FF.upD_[c]);
        putList(out, CHIML_LIST_LORD, c, FF.upLorD_[c]);
        putList(out, CHIML_LIST_ORDIPD, c, FF.upOrDipD_[c]);
    }
    if(!FF.upOrDipM_.empty() || !FF.upChiOrDipP_.empty() || !FF.upChiOrDipM_.empty())
        throw std::runtime_error("plan dump: magnetic / chiral oriented-dipole lists are outside the covered hot path");
    putList(out, CHIML_LIST_ORDIPP, 0, FF.upOrDipP_);

    bool relToNorm = false;
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
                else relToNorm = true;      // position-dependent: the grids of setupDipMoments follow as DIPGRID records
            }
        }
        appVec(p, dip);
        putRec(out, "OBJECT", p);
    }
    // dipP_[c][p] (setupDipMoments, parallelFDTDField.hpp:960-1048) when any pole is oriented relative to the surface normal
    if(relToNorm)
        for(int c = 0; c < 3; ++c)
            for(size_t pp = 0; pp < FF.dipP_[c].size(); ++pp)
            {
                int32_t hd[2] = {c, int32_t(pp)};
                std::string q; app(q, hd);
                const double* g = &FF.dipP_[c][pp]->point(0);
                q.append(reinterpret_cast<const char*>(g), sizeof(double) * size_t(pg.desc.ln[0]) * size_t(pg.desc.ln[1]) * size_t(pg.desc.ln[2]));
                putRec(out, "DIPGRID", q);
            }
    for(int c = 0; c < 3; ++c)
    {
        putCpml(out, c, FF.EPML_[c]);
        putCpml(out, 3 + c, FF.HPML_[c]);
    }
    // sources: the box this rank adds to, and the per-step amplitude dt*Re(sum pulse(t_k)) with t_k accumulated as step() does
    for(auto& srcBase : FF.srcArr_)
    {
        auto src = std::dynamic_pointer_cast<typename TR::source>(srcBase);
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
        std::vector<double> amp(nSteps), ampIm(nSteps);
        double t = 0.0;
        for(int k = 0; k < nSteps; ++k)
        {
            cplx pulVal = 0.0;
            for(auto& pul : src->pulse_) pulVal += pul->pulse(t);
            amp[k] = FF.dt_ * std::real(pulVal);
            ampIm[k] = FF.dt_ * std::imag(pulVal);           // complex fields: zaxpy_(n, dt_, pulVec_, ...) adds dt * pulse to both parts
            t += FF.dt_;
        }
        std::string p; app(p, h); appVec(p, amp);
        putRec(out, "SOURCE", p);
        if(TR::cplx) { std::string q; int32_t ns = nSteps; app(q, ns); appVec(q, ampIm); putRec(out, "SRCIMAG", q); }   // belongs to the SOURCE record before it
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
    putEmittersAndDfts(out, FF);
}

static void putTfsfRecords(std::ofstream& out, parallelFDTDFieldReal& FF)
{
    if(!FF.tfsfArr_.empty())
    {
        if(FF.gridComm_->size() > 1) throw std::runtime_error("plan dump: TFSF sources on several ranks are outside the covered hot path");
        const TfsfLayout L = tfsfLayout(FF);
        for(const TfsfSurfaceRec& r : tfsfSurfaces(FF, L))
        {
            ChimlPlanTfsfSurfaceHdr h; std::memset(&h, 0, sizeof(h));
            h.comp = r.s.comp; h.incd_offset = r.s.incd_offset; h.incd_len = r.s.incd_len; h.n = r.s.n; h.stride_incd = r.s.stride_incd;
            h.stride_main = r.s.stride_main; h.npairs_D = r.s.npairs_D; h.npairs_U = r.s.npairs_U; h.has_ep_mu = r.s.ep_mu ? 1 : 0; h.prefactor = r.s.prefactor;
            std::string p; app(p, h);
            p.append(reinterpret_cast<const char*>(r.s.pairs_D), size_t(r.s.npairs_D) * 8);
            p.append(reinterpret_cast<const char*>(r.s.pairs_U), size_t(r.s.npairs_U) * 8);
            if(r.s.ep_mu) p.append(reinterpret_cast<const char*>(r.s.ep_mu), size_t(r.s.incd_len) * 8);
            putRec(out, "TFSFSURF", p);
        }
    }
}

static void putEmittersAndDfts(std::ofstream& out, parallelFDTDFieldReal& FF)
{
    if(!FF.qeArr_.empty() && FF.gridComm_->size() == 1) putEmitters(out, FF);
    // DFT records: every stored field of every flux object and frequency detector
    for(const DftStorageRef& r : allDftStorages(FF))
    {
        auto& real = r.st;
        ChimlPlanDftHdr h; std::memset(&h, 0, sizeof(h));
        h.field = fieldId(FF, real->grid_); h.group = r.group; h.every = r.every; h.nfreq = real->nfreq_;
        h.npts = real->fieldInFreq_->sz_[0]; h.stride = real->fieldInFreq_->stride_;
        h.nlines = real->fieldInFreq_->fInGridInds_.size() / 2; h.acc_len = real->fInReal_.size();
        std::string p; app(p, h); appVec(p, *r.freq);
        for(size_t ii = 0; ii + 1 < real->fieldInFreq_->fInGridInds_.size(); ii += 2)
        { ChimlDftLine l; l.ind = real->fieldInFreq_->fInGridInds_[ii]; l.out = real->fieldInFreq_->fInGridInds_[ii + 1]; app(p, l); }
        putRec(out, "DFT", p);
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

// ---------------------------------------------------------------------------------------------
// --gpu: the reference's propagator bound to the CUDA engine through the C ABI
// ---------------------------------------------------------------------------------------------
struct GpuApi
{
    void* lib = nullptr;
#define CHIML_API(name) decltype(&::name) name = nullptr;
    CHIML_API(chiml_gpu_create) CHIML_API(chiml_gpu_destroy) CHIML_API(chiml_gpu_last_error) CHIML_API(chiml_gpu_set_update_list)
    CHIML_API(chiml_gpu_set_object) CHIML_API(chiml_gpu_set_cpml) CHIML_API(chiml_gpu_add_source) CHIML_API(chiml_gpu_add_detector)
    CHIML_API(chiml_gpu_add_emitters) CHIML_API(chiml_gpu_add_dft) CHIML_API(chiml_gpu_commit) CHIML_API(chiml_gpu_step_n)
    CHIML_API(chiml_gpu_step_n_dft) CHIML_API(chiml_gpu_sync) CHIML_API(chiml_gpu_read_detector_range) CHIML_API(chiml_gpu_consume_detector)
    CHIML_API(chiml_gpu_read_population) CHIML_API(chiml_gpu_download_dft) CHIML_API(chiml_gpu_download_field) CHIML_API(chiml_gpu_download_pole)
    CHIML_API(chiml_gpu_download_ordip_pole) CHIML_API(chiml_gpu_launch_count) CHIML_API(chiml_gpu_download_emitter_state)
    CHIML_API(chiml_gpu_download_emitter_pol) CHIML_API(chiml_gpu_set_periodic) CHIML_API(chiml_gpu_add_tfsf_surface) CHIML_API(chiml_gpu_step_n_tfsf)
    CHIML_API(chiml_gpu_set_magnetic) CHIML_API(chiml_gpu_set_object_magnetic) CHIML_API(chiml_gpu_download_mag_pole)
    CHIML_API(chiml_gpu_set_dip_grid) CHIML_API(chiml_gpu_set_object_chiral) CHIML_API(chiml_gpu_set_prev_copy) CHIML_API(chiml_gpu_download_chi_pole) CHIML_API(chiml_gpu_download_prev_field)
#undef CHIML_API
    void load()
    {
        const char* env = std::getenv("CHIML_B200_LIB");
        std::string path = env ? env : "";
        if(path.empty())
        {
            char exe[4096]; ssize_t n = readlink("/proc/self/exe", exe, sizeof(exe) - 1);
            std::string dir = n > 0 ? std::string(exe, size_t(n)) : std::string(".");
            dir = dir.substr(0, dir.find_last_of('/'));                 // .../oracle/_ref
            path = dir + "/../../chiml_b200/libchiml_b200.so";
        }
        lib = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if(!lib) throw std::runtime_error(std::string("--gpu: cannot load ") + path + ": " + dlerror());
#define CHIML_API(name) name = reinterpret_cast<decltype(&::name)>(dlsym(lib, #name)); if(!name) throw std::runtime_error("--gpu: " #name " missing in the library");
        CHIML_API(chiml_gpu_create) CHIML_API(chiml_gpu_destroy) CHIML_API(chiml_gpu_last_error) CHIML_API(chiml_gpu_set_update_list)
        CHIML_API(chiml_gpu_set_object) CHIML_API(chiml_gpu_set_cpml) CHIML_API(chiml_gpu_add_source) CHIML_API(chiml_gpu_add_detector)
        CHIML_API(chiml_gpu_add_emitters) CHIML_API(chiml_gpu_add_dft) CHIML_API(chiml_gpu_commit) CHIML_API(chiml_gpu_step_n)
        CHIML_API(chiml_gpu_step_n_dft) CHIML_API(chiml_gpu_sync) CHIML_API(chiml_gpu_read_detector_range) CHIML_API(chiml_gpu_consume_detector)
        CHIML_API(chiml_gpu_read_population) CHIML_API(chiml_gpu_download_dft) CHIML_API(chiml_gpu_download_field) CHIML_API(chiml_gpu_download_pole)
        CHIML_API(chiml_gpu_download_ordip_pole) CHIML_API(chiml_gpu_launch_count) CHIML_API(chiml_gpu_download_emitter_state)
        CHIML_API(chiml_gpu_download_emitter_pol) CHIML_API(chiml_gpu_set_periodic) CHIML_API(chiml_gpu_add_tfsf_surface) CHIML_API(chiml_gpu_step_n_tfsf)
        CHIML_API(chiml_gpu_set_magnetic) CHIML_API(chiml_gpu_set_object_magnetic) CHIML_API(chiml_gpu_download_mag_pole)
        CHIML_API(chiml_gpu_set_dip_grid) CHIML_API(chiml_gpu_set_object_chiral) CHIML_API(chiml_gpu_set_prev_copy) CHIML_API(chiml_gpu_download_chi_pole) CHIML_API(chiml_gpu_download_prev_field)
#undef CHIML_API
    }
};

struct GpuBinding
{
    GpuApi api;
    ChimlCtx* ctx = nullptr;
    std::vector<std::vector<int>> dtcSlots;                // per detector, per stored field: device slot
    std::vector<std::vector<std::array<int, 6>>> dtcBox;   // ... and its local box (loc, sz)
    std::vector<size_t> dtcSamples;                        // samples read so far per detector
    struct DftRef { std::shared_ptr<parallelStorageFreqDTCReal> st; int slot; };
    std::vector<DftRef> dfts;
    std::vector<const std::vector<double>*> dftFreqs;      // frequency lists of the groups with a storage here, in group order
    TfsfLayout tfsfL;                                      // TFSF: the incident-line table of one step and the surface records
    std::vector<TfsfSurfaceRec> tfsfSurf;
    void check(int rc, const char* what) { if(rc != CHIML_OK) throw std::runtime_error(std::string(what) + ": " + api.chiml_gpu_last_error(ctx)); }
};

// what a maintainer adds at the end of parallelFDTDFieldReal's constructor (INTEGRATION.md bindGpu)
static void bindGpu(parallelFDTDFieldReal& FF, GpuBinding& B)
{
    B.api.load();
    GpuApi& A = B.api;
    ChimlGridDesc g;
    std::memset(&g, 0, sizeof(g));
    g.mode = (FF.E_[0] && FF.E_[2]) ? CHIML_MODE_3D : (FF.E_[0] ? CHIML_MODE_TE : CHIML_MODE_TM);
    auto ref = FF.E_[0] ? FF.E_[0] : FF.E_[2];
    for(int k = 0; k < 3; ++k) { g.ln[k] = ref->ln_vec(k); g.d[k] = FF.d_[k]; }
    g.dt = FF.dt_; g.has_D = (FF.D_[0] || FF.D_[2]) ? 1 : 0; g.pml_on_D = FF.dielectricMatInPML_ ? 1 : 0; g.n_objects = int(FF.objArr_.size());
    g.rank = 0; g.nranks = 1;
    if(A.chiml_gpu_create(&g, 0, &B.ctx) != CHIML_OK) throw std::runtime_error(std::string("chiml_gpu_create: ") + A.chiml_gpu_last_error(nullptr));
    // magnetic-dispersive media: B grids, CPML on B, the lists upB_ / upLorB_ below, magnetic pole constants per object
    const bool hasB = bool(FF.B_[0]) || bool(FF.B_[2]);
    if(hasB) B.check(A.chiml_gpu_set_magnetic(B.ctx, 1, FF.magMatInPML_ ? 1 : 0), "set_magnetic");
    // TFSF sources: the surface records go to the device, the 1-D incident line stays with the reference's object (gpuStep)
    B.tfsfL = tfsfLayout(FF);
    B.tfsfSurf = tfsfSurfaces(FF, B.tfsfL);
    for(const TfsfSurfaceRec& r : B.tfsfSurf) B.check(A.chiml_gpu_add_tfsf_surface(B.ctx, &r.s), "add_tfsf_surface");

    static_assert(sizeof(upLists::value_type) == sizeof(ChimlRun), "upLists entries are handed over as ChimlRun");
    auto put = [&](int kind, int comp, const upLists& l) {
        B.check(A.chiml_gpu_set_update_list(B.ctx, kind, comp, reinterpret_cast<const ChimlRun*>(l.data()), l.size()), "set_update_list"); };
    for(int c = 0; c < 3; ++c)
    {
        if(!FF.upOrDipB_[c].empty() || !FF.upOrDipChiD_[c].empty() || !FF.upOrDipChiB_[c].empty())
            throw std::runtime_error("--gpu: magnetic / chiral oriented-dipole update lists are outside the covered hot path");
        if(hasB) { put(CHIML_LIST_D, 3 + c, FF.upB_[c]); put(CHIML_LIST_LORD, 3 + c, FF.upLorB_[c]); }
        if(!FF.upChiD_[c].empty() || !FF.upChiB_[c].empty()) { put(CHIML_LIST_CHID, c, FF.upChiD_[c]); put(CHIML_LIST_CHID, 3 + c, FF.upChiB_[c]); }   // chiral media
        put(CHIML_LIST_U, c, FF.upE_[c]);   put(CHIML_LIST_U, 3 + c, FF.upH_[c]);
        put(CHIML_LIST_D, c, FF.upD_[c]);   put(CHIML_LIST_LORD, c, FF.upLorD_[c]);
        put(CHIML_LIST_ORDIPD, c, FF.upOrDipD_[c]);
    }
    put(CHIML_LIST_ORDIPP, 0, FF.upOrDipP_);
    bool relToNorm = false;
    for(size_t oo = 0; oo < FF.objArr_.size(); ++oo)
    {
        auto& obj = FF.objArr_[oo];
        const int np = int(obj->gamma().size());
        std::vector<double> dip(3 * size_t(np), 0.0);
        if(obj->useOrdDip())
            for(int pp = 0; pp < np; ++pp)
            {
                MAT_DIP_ORIENTAITON ori = obj->dipOr(pp);
                if(ori == MAT_DIP_ORIENTAITON::ISOTROPIC) dip[3 * pp] = dip[3 * pp + 1] = dip[3 * pp + 2] = 1.0;
                else if(ori == MAT_DIP_ORIENTAITON::UNIDIRECTIONAL) for(int k = 0; k < 3; ++k) dip[3 * pp + k] = obj->dipE(pp)[k];
                else relToNorm = true;
            }
        B.check(A.chiml_gpu_set_object(B.ctx, int(oo), np, obj->alpha().data(), obj->xi().data(), obj->gamma().data(), obj->useOrdDip() ? 1 : 0, dip.data()), "set_object");
        if(hasB) B.check(A.chiml_gpu_set_object_magnetic(B.ctx, int(oo), int(obj->magGamma().size()), obj->magAlpha().data(), obj->magXi().data(), obj->magGamma().data()), "set_object_magnetic");
        if(!obj->chiGamma().empty())
            B.check(A.chiml_gpu_set_object_chiral(B.ctx, int(oo), int(obj->chiGamma().size()), obj->chiAlpha().data(), obj->chiXi().data(), obj->chiGamma().data(), obj->chiGammaPrev().data()), "set_object_chiral");
    }
    if(relToNorm)       // orientations relative to the surface normal: the engine reads the reference's own dipP_ grids
        for(int c = 0; c < 3; ++c)
            for(size_t pp = 0; pp < FF.dipP_[c].size(); ++pp)
                B.check(A.chiml_gpu_set_dip_grid(B.ctx, c, int(pp), &FF.dipP_[c][pp]->point(0)), "set_dip_grid");
    // periodic boundaries: the arguments of applBCE_ / applBCH_ (single rank: applyBC1Proc)
    if(FF.E_[0] ? FF.E_[0]->PBC() : FF.E_[2]->PBC())
        for(int comp = 0; comp < 6; ++comp)
            if(comp < 3 ? FF.E_[comp] : FF.H_[comp - 3])
            {
                const ChimlWrap w = wrapArgs(FF, comp);
                B.check(A.chiml_gpu_set_periodic(B.ctx, comp, &w), "set_periodic");
            }
    if(!FF.copy2PrevFields_.empty())                 // chiral media: the rows copied into prevE_ / prevH_ (std::array<int,4> rows are 4 ints each)
        B.check(A.chiml_gpu_set_prev_copy(B.ctx, reinterpret_cast<const int32_t*>(FF.copy2PrevFields_.data()), FF.copy2PrevFields_.size()), "set_prev_copy");
    static_assert(sizeof(updatePsiParams) == sizeof(ChimlPsiParams) && sizeof(updateGridParams) == sizeof(ChimlGridParams), "CPML list layouts");
    for(int c = 0; c < 3; ++c)
        for(int side = 0; side < 2; ++side)
        {
            auto pml = side == 0 ? FF.EPML_[c] : FF.HPML_[c];
            if(!pml) continue;
            const int comp = side == 0 ? c : 3 + c;
            if(pml->grid_k_)
                B.check(A.chiml_gpu_set_cpml(B.ctx, comp, 0, pml->psi_j_ ? 1 : 0, reinterpret_cast<const ChimlPsiParams*>(pml->updateListPsi_j_.data()), pml->updateListPsi_j_.size(),
                                             reinterpret_cast<const ChimlGridParams*>(pml->updateListGrid_k_.data()), pml->updateListGrid_k_.size()), "set_cpml");
            if(pml->grid_j_)
                B.check(A.chiml_gpu_set_cpml(B.ctx, comp, 1, pml->psi_k_ ? 1 : 0, reinterpret_cast<const ChimlPsiParams*>(pml->updateListPsi_k_.data()), pml->updateListPsi_k_.size(),
                                             reinterpret_cast<const ChimlGridParams*>(pml->updateListGrid_j_.data()), pml->updateListGrid_j_.size()), "set_cpml");
        }
    for(auto& srcBase : FF.srcArr_)
    {
        auto src = std::dynamic_pointer_cast<parallelSourceNormalReal>(srcBase);
        if(!src) throw std::runtime_error("--gpu: only normal (axis-aligned) soft sources are on the covered hot path");
        if(!src->slave_) throw std::runtime_error("--gpu: a source without a local box on a single rank");
        const SalveSource& sl = *src->slave_;
        const int ax1 = sl.addVec1_[0] ? 0 : (sl.addVec1_[1] ? 1 : 2), ax2 = sl.addVec2_[0] ? 0 : (sl.addVec2_[1] ? 1 : 2);
        int32_t loc[3], sz[3] = {1, 1, 1};
        sz[3 - ax1 - ax2] = sl.sz_[0]; sz[ax1] = sl.sz_[1]; sz[ax2] = sl.sz_[2];
        for(int k = 0; k < 3; ++k) loc[k] = sl.loc_[k];
        B.check(A.chiml_gpu_add_source(B.ctx, fieldId(FF, src->grid_), loc, sz, nullptr), "add_source");
    }
    const bool threeD = g.mode == CHIML_MODE_3D;
    for(auto& dtc : FF.dtcArr_)
    {
        std::vector<int> slots; std::vector<std::array<int, 6>> boxes;
        for(auto& f : dtc->fields_)
        {
            // stored box in global no-ghost coordinates, already grown by the Yee offset (DTC/parallelStorageDTC.hpp:51-80) -> local ghost-inclusive
            int32_t loc[3] = {f->loc_[0] + 1, f->loc_[1] - ref->procLoc(1) + 1, threeD ? f->loc_[2] + 1 : 0};
            int32_t sz[3] = {f->sz_[0], f->sz_[1], threeD ? f->sz_[2] : 1};
            int slot = -1;
            B.check(A.chiml_gpu_add_detector(B.ctx, fieldId(FF, f->grid_), loc, sz, dtc->timeInterval_, &slot), "add_detector");
            slots.push_back(slot); boxes.push_back({{loc[0], loc[1], loc[2], sz[0], sz[1], sz[2]}});
        }
        B.dtcSlots.push_back(slots); B.dtcBox.push_back(boxes); B.dtcSamples.push_back(1);    // sample 0 = t = 0, written by the constructor already
    }
    // emitters: the EMITTER record of the plan dump is exactly ChimlEmitterDesc + arrays; reuse that extraction through a memory stream
    // (kept in one place: putEmitters) -- done below by the caller, which owns the buffers
    static_assert(sizeof(ChimlDftLine) == 2 * sizeof(int), "fInGridInds_ is a list of (grid index, accumulator index) pairs");
    for(const DftStorageRef& r : allDftStorages(FF))
    {
        auto& st = r.st;
        int slot = -1;
        B.check(A.chiml_gpu_add_dft(B.ctx, fieldId(FF, st->grid_), r.group, r.every, st->nfreq_, st->fieldInFreq_->sz_[0], st->fieldInFreq_->stride_,
                                    reinterpret_cast<const ChimlDftLine*>(st->fieldInFreq_->fInGridInds_.data()), st->fieldInFreq_->fInGridInds_.size() / 2,
                                    st->fInReal_.size(), &slot), "add_dft");
        B.dfts.push_back({st, slot});
    }
    B.dftFreqs = dftGroupFreqs(FF);
}

// emitter objects: ChimlEmitterDesc from the members of parallelQEBase (what putEmitters writes into a plan file)
struct EmitterBuffers
{
    std::vector<double> h0, weight, mu, gval, eps; std::vector<int32_t> gptr, gcol, loc, pop;
};
static void bindEmitters(parallelFDTDFieldReal& FF, GpuBinding& B, std::vector<EmitterBuffers>& keep)
{
    int qq = 0;
    for(auto& qe : FF.qeArr_)
    {
        if(!qe->sameProcCalc_) throw std::runtime_error("--gpu: emitter sets need a single-rank run");
        keep.emplace_back();
        EmitterBuffers& b = keep.back();
        ChimlEmitterDesc d;
        std::memset(&d, 0, sizeof(d));
        d.object = qq++;
        d.nlevel = qe->nlevel_; d.nsys = int(qe->levelSys_.size()); d.nemit = int(qe->levelSys_[0].den_.size());
        auto eg = qe->e_[0] ? qe->e_[0] : qe->e_[2];
        auto Pg = qe->P_[0] ? qe->P_[0] : qe->P_[2];
        d.box_n[0] = eg->x(); d.box_n[1] = eg->y(); d.box_n[2] = eg->z();
        for(int k = 0; k < 3; ++k) d.box_lo[k] = qe->sameProcCalc_->loc_[k];
        d.dt = qe->dt_; d.inv_hbar = std::imag(qe->one_over_hbar_); d.na = qe->na_;
        const int n2 = d.nlevel * d.nlevel;
        for(auto& ls : qe->levelSys_) for(int k = 0; k < n2; ++k) { b.h0.push_back(ls.ham_->h0_[k].real()); b.h0.push_back(ls.ham_->h0_[k].imag()); }
        for(auto& ew : qe->energyWeights_) b.weight.push_back(ew.second);
        auto& ham = *qe->levelSys_[0].ham_;
        for(auto* v : {&ham.x_expectation_, &ham.y_expectation_, &ham.z_expectation_}) for(int k = 0; k < n2; ++k) { b.mu.push_back((*v)[k].real()); b.mu.push_back((*v)[k].imag()); }
        b.gptr.push_back(0);
        for(auto& row : qe->gam_) { for(auto it = row.begin(); it != row.end(); ++it) { b.gcol.push_back(it->first); b.gval.push_back(it->second); } b.gptr.push_back(int32_t(b.gcol.size())); }
        while(int(b.gptr.size()) < n2 + 1) b.gptr.push_back(b.gptr.back());
        for(auto& den : qe->levelSys_[0].den_) { b.loc.push_back(den.x()); b.loc.push_back(den.y()); b.loc.push_back(den.z()); }
        for(int y = 0; y < Pg->y(); ++y)
            for(int z = 0; z < Pg->z(); ++z)
                for(int x = 0; x < Pg->x(); ++x)
                    b.eps.push_back(qe->eps_->z() == 1 ? qe->eps_->point(d.box_lo[0] + x, d.box_lo[1] + y, 0) : qe->eps_->point(d.box_lo[0] + x, d.box_lo[1] + y, d.box_lo[2] + z));
        for(auto& p : qe->dtcPopArr_) b.pop.push_back(p->level_);
        d.h0 = b.h0.data(); d.weight = b.weight.data(); d.mu = b.mu.data(); d.gam_ptr = b.gptr.data(); d.gam_col = b.gcol.data(); d.gam_val = b.gval.data();
        d.loc = b.loc.data(); d.eps = b.eps.data(); d.npop = int(b.pop.size()); d.pop_level = b.pop.data();
        d.pop_every = d.npop ? qe->dtcPopArr_[0]->timeInt_ : 1;
        d.npoints = d.npop ? qe->dtcPopArr_[0]->npoints_ : d.nemit;
        B.check(B.api.chiml_gpu_add_emitters(B.ctx, &d, nullptr), "add_emitters");
    }
}

// the body of parallelFDTDFieldBase<double>::step() with the engine behind it (INTEGRATION.md step())
static void gpuStep(parallelFDTDFieldReal& FF, GpuBinding& B)
{
    GpuApi& A = B.api;
    std::vector<double> amp;                                  // SOURCE/parallelSourceNormal.cpp:15-37: dt * Re(sum pulse(t))
    for(auto& srcBase : FF.srcArr_)
    {
        auto src = std::dynamic_pointer_cast<parallelSourceNormalReal>(srcBase);
        cplx p = 0.0;
        for(auto& pul : src->pulse_) p += pul->pulse(FF.tcur_);
        amp.push_back(FF.dt_ * std::real(p));
    }
    std::vector<double> tfsfRow;
    if(!FF.tfsfArr_.empty())
    {
        // step() items around tfsf->updateFields() (FDTD_MANAGER/parallelFDTDField.hpp:1238-1255): the H surfaces read the incident E
        // lines as they are now, the line takes its own step on the host, the E surfaces read the incident H lines after it; the
        // incident-field series for the flux normalisation are recorded as the reference records them
        tfsfRow.assign(size_t(B.tfsfL.perStep), 0.0);
        tfsfGrab(FF, B.tfsfL, true, tfsfRow.data());
        for(auto& tfsf : FF.tfsfArr_)
        {
            FF.H_incd_[0][2 * FF.t_step_ + 0] = tfsf->get_incd_Hx(); FF.H_incd_[0][2 * FF.t_step_ + 1] = tfsf->get_incd_Hx_off();
            FF.H_incd_[1][2 * FF.t_step_ + 0] = tfsf->get_incd_Hy(); FF.H_incd_[1][2 * FF.t_step_ + 1] = tfsf->get_incd_Hy_off();
            FF.H_incd_[2][2 * FF.t_step_ + 0] = tfsf->get_incd_Hz(); FF.H_incd_[2][2 * FF.t_step_ + 1] = tfsf->get_incd_Hz_off();
            tfsf->step();
            FF.E_incd_[0][2 * FF.t_step_ + 0] = tfsf->get_incd_Ex(); FF.E_incd_[0][2 * FF.t_step_ + 1] = tfsf->get_incd_Ex_off();
            FF.E_incd_[1][2 * FF.t_step_ + 0] = tfsf->get_incd_Ey(); FF.E_incd_[1][2 * FF.t_step_ + 1] = tfsf->get_incd_Ey_off();
            FF.E_incd_[2][2 * FF.t_step_ + 0] = tfsf->get_incd_Ez(); FF.E_incd_[2][2 * FF.t_step_ + 1] = tfsf->get_incd_Ez_off();
        }
        tfsfGrab(FF, B.tfsfL, false, tfsfRow.data());
    }
    std::vector<double> tw;                                   // parallelFluxDTC::fieldIn: fftFact_ = exp(i * (-t * freq)), t = time after the step
    if(!B.dfts.empty())
    {
        const double t = FF.tcur_ + FF.dt_;
        for(const std::vector<double>* fl : B.dftFreqs)
            for(double f : *fl) { const cplx w = std::exp(cplx(0.0, -1.0 * t * f)); tw.push_back(w.real()); tw.push_back(w.imag()); }
    }
    if(!FF.tfsfArr_.empty())
        B.check(A.chiml_gpu_step_n_tfsf(B.ctx, 1, amp.empty() ? nullptr : amp.data(), tw.empty() ? nullptr : tw.data(), tfsfRow.data(), tfsfRow.size()), "step_n_tfsf");
    else if(B.dfts.empty()) B.check(A.chiml_gpu_step_n(B.ctx, 1, amp.empty() ? nullptr : amp.data()), "step_n");
    else B.check(A.chiml_gpu_step_n_dft(B.ctx, 1, amp.empty() ? nullptr : amp.data(), tw.data()), "step_n_dft");
    FF.tcur_ += FF.dt_;
    ++FF.t_step_;
    // detectors: the sampled boxes come back from the device ring into the reference's own grids, and its own writer formats them
    for(size_t d = 0; d < FF.dtcArr_.size(); ++d)
    {
        auto& dtc = FF.dtcArr_[d];
        if(FF.t_step_ % dtc->timeInt() != 0) continue;
        for(size_t f = 0; f < dtc->fields_.size(); ++f)
        {
            const auto& bx = B.dtcBox[d][f];
            std::vector<double> buf(size_t(bx[3]) * bx[4] * bx[5]);
            size_t got = 0;
            B.check(A.chiml_gpu_read_detector_range(B.ctx, B.dtcSlots[d][f], B.dtcSamples[d], 1, buf.data(), &got), "read_detector_range");
            if(got != 1) throw std::runtime_error("--gpu: detector sample missing");
            B.check(A.chiml_gpu_consume_detector(B.ctx, B.dtcSlots[d][f], B.dtcSamples[d] + 1), "consume_detector");
            auto& grid = dtc->fields_[f]->grid_;
            size_t i = 0;                                     // sample layout: x fastest, then z, then y
            for(int y = 0; y < bx[4]; ++y)
                for(int z = 0; z < bx[5]; ++z)
                    for(int x = 0; x < bx[3]; ++x) grid->point(bx[0] + x, bx[1] + y, bx[2] + z) = buf[i++];
        }
        ++B.dtcSamples[d];
        dtc->output(FF.tcur_);
    }
    // flux->fieldIn(tcur_) ran on the device (the running-DFT sets); what is left of it on the host is its sample counter, which
    // getFlux normalises with (DTC/parallelFlux.hpp:313,418)
    for(auto& flux : FF.fluxArr_)
        if(FF.t_step_ % flux->timeInt() == 0) ++flux->t_step_;
    for(auto& dtc : FF.dtcFreqArr_)                            // likewise parallelDetectorFREQ_Base::output (DTC/parallelDTC_FREQ.hpp:258-264)
        if(FF.t_step_ % dtc->timeInt() == 0) ++dtc->t_step_;
}

// after the loop (main.cpp:67-118): accumulators, populations and -- for the state dump of the tests -- every grid come back
static void gpuFinish(parallelFDTDFieldReal& FF, GpuBinding& B)
{
    GpuApi& A = B.api;
    B.check(A.chiml_gpu_sync(B.ctx), "sync");
    for(auto& d : B.dfts) B.check(A.chiml_gpu_download_dft(B.ctx, d.slot, d.st->fInReal_.data(), d.st->fInCplx_.data()), "download_dft");
    int qq = 0;
    for(auto& qe : FF.qeArr_)
    {
        int dd = 0;
        for(auto& pop : qe->dtcPopArr_)
        {
            size_t ns = 0;
            B.check(A.chiml_gpu_read_population(B.ctx, qq, dd, nullptr, 0, &ns), "read_population");
            std::vector<double> v(2 * ns);
            if(ns) B.check(A.chiml_gpu_read_population(B.ctx, qq, dd, v.data(), ns, &ns), "read_population");
            pop->allPop_.clear();
            for(size_t k = 0; k < ns; ++k) pop->allPop_.push_back(cplx(v[2 * k], v[2 * k + 1]));
            ++dd;
        }
        // density matrices, their derivative histories and the polarisation boxes, for callers that look at them (the state dump)
        const int n2 = qe->nlevel_ * qe->nlevel_;
        for(size_t ss = 0; ss < qe->levelSys_.size(); ++ss)
            for(int w = 0; w < 5; ++w)
            {
                std::vector<double> st(2 * size_t(n2) * qe->levelSys_[ss].den_.size());
                if(st.empty()) continue;
                B.check(A.chiml_gpu_download_emitter_state(B.ctx, qq, int(ss), w, st.data()), "download_emitter_state");
                size_t e = 0;
                for(auto& den : qe->levelSys_[ss].den_)
                {
                    std::vector<cplx>& v = w == 0 ? den.density_ : w == 1 ? den.density_deriv_n_ : w == 2 ? den.density_deriv_n_minus_1_
                                         : w == 3 ? den.density_deriv_n_minus_2_ : den.density_deriv_n_minus_3_;
                    for(int k = 0; k < n2; ++k) v[k] = cplx(st[2 * (e * n2 + k)], st[2 * (e * n2 + k) + 1]);
                    ++e;
                }
            }
        for(int c = 0; c < 3; ++c)
            if(qe->P_[c] && qe->E_[c]) B.check(A.chiml_gpu_download_emitter_pol(B.ctx, qq, c, &qe->P_[c]->point(0)), "download_emitter_pol");
        ++qq;
    }
    for(int c = 0; c < 3; ++c)
    {
        if(FF.E_[c]) B.check(A.chiml_gpu_download_field(B.ctx, CHIML_EX + c, &FF.E_[c]->point(0)), "download_field");
        if(FF.H_[c]) B.check(A.chiml_gpu_download_field(B.ctx, CHIML_HX + c, &FF.H_[c]->point(0)), "download_field");
        if(FF.D_[c]) B.check(A.chiml_gpu_download_field(B.ctx, CHIML_DX + c, &FF.D_[c]->point(0)), "download_field");
        if(FF.B_[c]) B.check(A.chiml_gpu_download_field(B.ctx, CHIML_BX + c, &FF.B_[c]->point(0)), "download_field");
        if(FF.prevE_[c] && !FF.copy2PrevFields_.empty()) B.check(A.chiml_gpu_download_prev_field(B.ctx, c, &FF.prevE_[c]->point(0)), "download_prev_field");
        if(FF.prevH_[c] && !FF.copy2PrevFields_.empty()) B.check(A.chiml_gpu_download_prev_field(B.ctx, 3 + c, &FF.prevH_[c]->point(0)), "download_prev_field");
        for(size_t p = 0; p < FF.lorChiHP_[c].size(); ++p)
        {
            B.check(A.chiml_gpu_download_chi_pole(B.ctx, c, int(p), 0, &FF.lorChiHP_[c][p]->point(0)), "download_chi_pole");
            B.check(A.chiml_gpu_download_chi_pole(B.ctx, c, int(p), 1, &FF.prevLorChiHP_[c][p]->point(0)), "download_chi_pole");
        }
        for(size_t p = 0; p < FF.lorChiEM_[c].size(); ++p)
        {
            B.check(A.chiml_gpu_download_chi_pole(B.ctx, 3 + c, int(p), 0, &FF.lorChiEM_[c][p]->point(0)), "download_chi_pole");
            B.check(A.chiml_gpu_download_chi_pole(B.ctx, 3 + c, int(p), 1, &FF.prevLorChiEM_[c][p]->point(0)), "download_chi_pole");
        }
        for(size_t p = 0; p < FF.lorM_[c].size(); ++p)
        {
            B.check(A.chiml_gpu_download_mag_pole(B.ctx, c, int(p), 0, &FF.lorM_[c][p]->point(0)), "download_mag_pole");
            B.check(A.chiml_gpu_download_mag_pole(B.ctx, c, int(p), 1, &FF.prevLorM_[c][p]->point(0)), "download_mag_pole");
        }
        for(size_t p = 0; p < FF.lorP_[c].size(); ++p)
        {
            B.check(A.chiml_gpu_download_pole(B.ctx, c, int(p), 0, &FF.lorP_[c][p]->point(0)), "download_pole");
            B.check(A.chiml_gpu_download_pole(B.ctx, c, int(p), 1, &FF.prevLorP_[c][p]->point(0)), "download_pole");
        }
        for(size_t p = 0; p < FF.orDipLorP_[c].size(); ++p)
        {
            B.check(A.chiml_gpu_download_ordip_pole(B.ctx, c, int(p), 0, &FF.orDipLorP_[c][p]->point(0)), "download_ordip_pole");
            B.check(A.chiml_gpu_download_ordip_pole(B.ctx, c, int(p), 1, &FF.prevOrDipLorP_[c][p]->point(0)), "download_ordip_pole");
        }
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
    {
        // Bloch-periodic run (k-point != 0 switches the reference to complex fields, INPUTS/parallelInputs.cpp:108-112): main.cpp:129-199 with
        // parallelFDTDFieldCplx; plan and state dumps as for real fields, every grid as a real and an imaginary array
        if(opt.gpu) throw std::runtime_error("chiml_ref --gpu: complex fields are driven by the host driver of this repository (two real field sets), not by this binding");
        parallelFDTDFieldCplx FC(IP, gridComm);
        int nStepsC = int(std::ceil(IP.tMax_ / IP.dt_));
        if(opt.steps >= 0) nStepsC = opt.steps;
        if(!opt.plan.empty()) writePlan(opt.plan + ".rank" + std::to_string(rank) + ".plan", FC, IP, nStepsC);
        for(int tt = 0; tt < opt.warmup; ++tt) FC.step();
        gridComm->barrier();
        auto c0 = std::chrono::steady_clock::now();
        for(int tt = 0; tt < nStepsC; ++tt) FC.step();
        gridComm->barrier();
        auto c1 = std::chrono::steady_clock::now();
        if(rank == 0)
        {
            g_stepSeconds = std::chrono::duration<double>(c1 - c0).count();
            g_nStepsRun = nStepsC;
            g_cells = long(FC.n_vec_[0]) * long(FC.n_vec_[1]) * long(FC.n_vec_[2] > 1 ? FC.n_vec_[2] : 1);
        }
        if(!opt.dump.empty())
        {
            const char* c = "xyz";
            for(int i = 0; i < 3; ++i)
            {
                grabGrid(rank, std::string("E") + c[i], FC.E_[i]);
                grabGrid(rank, std::string("H") + c[i], FC.H_[i]);
                grabGrid(rank, std::string("D") + c[i], FC.D_[i]);
                for(size_t p = 0; p < FC.lorP_[i].size(); ++p)
                {
                    grabGrid(rank, std::string("P") + c[i] + std::to_string(p), FC.lorP_[i][p]);
                    grabGrid(rank, std::string("pP") + c[i] + std::to_string(p), FC.prevLorP_[i][p]);
                }
            }
        }
        if(opt.output)
            for(auto& dtc : FC.dtcArr()) dtc->toFile();
        return;
    }

    parallelFDTDFieldReal FF(IP, gridComm);
    int nSteps = int(std::ceil(IP.tMax_ / IP.dt_));
    if(opt.steps >= 0) nSteps = opt.steps;

    if(!opt.plan.empty())
        writePlan(opt.plan + ".rank" + std::to_string(rank) + ".plan", FF, IP, nSteps);

    GpuBinding gpu;
    std::vector<EmitterBuffers> emitterBuffers;
    if(opt.gpu)
    {
        if(gridComm->size() != 1) throw std::runtime_error("--gpu runs a single rank (one process per slab is needed for the CUDA IPC halo)");
        bindGpu(FF, gpu);
        bindEmitters(FF, gpu, emitterBuffers);
        gpu.check(gpu.api.chiml_gpu_commit(gpu.ctx), "commit");
    }
    // a CPU run that writes a plan also records what the TFSF surfaces read from the reference's 1-D incident line, step by step
    const bool tfsfRecord = !FF.tfsfArr_.empty() && !opt.plan.empty() && !opt.gpu;
    const TfsfLayout tfsfL = tfsfLayout(FF);
    auto stepOnce = [&]() {
        if(opt.gpu) { gpuStep(FF, gpu); return; }
        if(!tfsfRecord) { FF.step(); return; }
        std::vector<double> row(size_t(tfsfL.perStep), 0.0);
        tfsfGrab(FF, tfsfL, true, row.data());
        FF.step();
        tfsfGrab(FF, tfsfL, false, row.data());
        g_tfsfTable.insert(g_tfsfTable.end(), row.begin(), row.end());
        g_tfsfPerStep = tfsfL.perStep;
    };
    for(int tt = 0; tt < opt.warmup; ++tt) stepOnce();
    gridComm->barrier();
    auto t0 = std::chrono::steady_clock::now();
    for(int tt = 0; tt < nSteps; ++tt) stepOnce();
    if(opt.gpu) gpuFinish(FF, gpu);
    if(tfsfRecord && g_tfsfPerStep > 0)
    {
        std::ofstream out((opt.plan + ".rank" + std::to_string(rank) + ".plan").c_str(), std::ios::binary | std::ios::app);
        ChimlPlanTfsfLinesHdr h; h.n_steps = int(g_tfsfTable.size() / size_t(g_tfsfPerStep)); h.per_step = g_tfsfPerStep;
        std::string p; app(p, h); appVec(p, g_tfsfTable);
        putRec(out, "TFSFLINE", p);
    }
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
            grabGrid(rank, std::string("vE") + c[i], FF.prevE_[i]);
            grabGrid(rank, std::string("vH") + c[i], FF.prevH_[i]);
            for(size_t p = 0; p < FF.lorChiHP_[i].size(); ++p)
            {
                grabGrid(rank, std::string("cP") + c[i] + std::to_string(p), FF.lorChiHP_[i][p]);
                grabGrid(rank, std::string("cvP") + c[i] + std::to_string(p), FF.prevLorChiHP_[i][p]);
            }
            for(size_t p = 0; p < FF.lorChiEM_[i].size(); ++p)
            {
                grabGrid(rank, std::string("cM") + c[i] + std::to_string(p), FF.lorChiEM_[i][p]);
                grabGrid(rank, std::string("cvM") + c[i] + std::to_string(p), FF.prevLorChiEM_[i][p]);
            }
            for(size_t p = 0; p < FF.lorM_[i].size(); ++p)
            {
                grabGrid(rank, std::string("M") + c[i] + std::to_string(p), FF.lorM_[i][p]);
                grabGrid(rank, std::string("pM") + c[i] + std::to_string(p), FF.prevLorM_[i][p]);
            }
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
        for(const DftStorageRef& r : allDftStorages(FF))
        {
            for(int im = 0; im < 2; ++im)
            {
                GridDump d; d.rank = rank; d.name = "dft" + std::to_string(slot) + (im ? "i" : "r");
                const std::vector<double>& v = im ? r.st->fInCplx_ : r.st->fInReal_;
                d.ln[0] = int(v.size()); d.ln[1] = 1; d.ln[2] = 1; d.yStart = 0;
                d.data = v;
                std::lock_guard<std::mutex> lk(g_dumpMtx); g_dumps.push_back(std::move(d));
            }
            ++slot;
        }
    }

    if(!opt.incdDump.empty() && rank == 0)
    {
        std::ofstream out(opt.incdDump.c_str(), std::ios::binary);
        const int32_t n = int32_t(FF.E_incd_[0].size());
        out.write("CHIMLINC", 8); out.write(reinterpret_cast<const char*>(&n), 4);
        for(int k = 0; k < 6; ++k)
        {
            const std::vector<cplx>& v = k < 3 ? FF.E_incd_[k] : FF.H_incd_[k - 3];
            out.write(reinterpret_cast<const char*>(v.data()), std::streamsize(v.size() * sizeof(cplx)));
        }
    }
    if(opt.output)
    {
        for(auto& flux : FF.fluxArr())
            flux->getFlux(FF.ExIncd(), FF.EyIncd(), FF.EzIncd(), FF.HxIncd(), FF.HyIncd(), FF.HzIncd(), true);
        for(auto& dtc : FF.dtcFreqArr())                      // main.cpp:74-107: the variant with incident fields first
        {
            try
            {
                std::vector<std::vector<cplx>> incdFields;
                if(dtc->type() == DTCTYPE::HPOW) incdFields = { FF.HxIncd(), FF.HyIncd(), FF.HzIncd() };
                else if(dtc->type() == DTCTYPE::EPOW) incdFields = { FF.ExIncd(), FF.EyIncd(), FF.EzIncd() };
                else if(dtc->type() == DTCTYPE::EX || dtc->type() == DTCTYPE::PX) incdFields = { FF.ExIncd() };
                else if(dtc->type() == DTCTYPE::EY || dtc->type() == DTCTYPE::PY) incdFields = { FF.EyIncd() };
                else if(dtc->type() == DTCTYPE::EZ || dtc->type() == DTCTYPE::PZ) incdFields = { FF.EzIncd() };
                else if(dtc->type() == DTCTYPE::HX || dtc->type() == DTCTYPE::MX) incdFields = { FF.HxIncd() };
                else if(dtc->type() == DTCTYPE::HY || dtc->type() == DTCTYPE::MY) incdFields = { FF.HyIncd() };
                else if(dtc->type() == DTCTYPE::HZ || dtc->type() == DTCTYPE::MZ) incdFields = { FF.HzIncd() };
                if(dtc->outputMaps()) dtc->toMap(incdFields, FF.dt());
                else dtc->toFile(incdFields, FF.dt());
            }
            catch(std::exception& e)
            {
                if(dtc->outputMaps()) dtc->toMap();
                else dtc->toFile();
            }
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
    if(opt.gpu)
    {
        if(rank == 0 && !opt.quiet) std::fprintf(stderr, "chiml_ref --gpu: %lld kernel launches\n", (long long)gpu.api.chiml_gpu_launch_count(gpu.ctx));
        gpu.api.chiml_gpu_destroy(gpu.ctx);
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
        else if(s == "--gpu") opt.gpu = true;
        else if(s == "--incd-dump" && a + 1 < argc) opt.incdDump = argv[++a];
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
