// Host-side construction of everything chiml_gpu.h consumes, for one y-slab: object maps -> update
// lists, CPML lists, pole constants, source boxes + amplitudes, detector boxes.  This is the work of
// the reference's propagator constructor (FDTD_MANAGER/parallelFDTDField.hpp:214-617,
// parallelFDTDField.cpp:13-852; PML/parallelPML.hpp:102-656) with a different architecture: the
// reference materialises eight full-size object-id and eps grids per rank; here rows of those maps
// are generated on the fly (one x-row at a time, objects culled by bounding box) and turned straight
// into run-length lists, so a 2048 x 256 x 1024 slab needs megabytes of host memory, not tens of
// gigabytes, and the row loop runs on all host threads.
#pragma once

#include <array>
#include <complex>
#include <memory>
#include <string>
#include <vector>

#include "../../include/chiml_plan.h"
#include "inputs.hpp"

namespace chiml_host {

struct PlanCpml { int comp, part, has_psi; std::vector<ChimlPsiParams> psi; std::vector<ChimlGridParams> grid; };
struct PlanObject
{
    int npoles, use_or_dip, ml; double eps_inf, mu_inf; std::vector<double> alpha, xi, gamma, dip;
    std::vector<double> magAlpha, magXi, magGamma;                  // magnetic poles (chiml_gpu_set_object_magnetic)
    std::vector<double> chiAlpha, chiXi, chiGamma, chiGammaPrev;    // chiral poles (chiml_gpu_set_object_chiral)
};
struct PlanSource { int field; int32_t loc[3], sz[3]; std::vector<double> amp, amp_im; };     // amp_im: dt * Im(sum pulse), complex fields only
struct PlanDetector { int detector, field; int32_t loc[3], sz[3], offset[3]; int every, type; double conv, t_conv; };

// one parallelQE object restricted to the slab (include/chiml_gpu.h ChimlEmitterDesc)
struct PlanEmitter
{
    int object = 0, nlevel = 0, nsys = 0, nemit = 0;
    int32_t box_lo[3] = {0, 0, 0}, box_n[3] = {0, 0, 0};
    int pz = 0, pop_every = 1, npoints = 0;
    double dt = 0, inv_hbar = 0, na = 0;
    std::vector<double> h0, weight, mu, gam_val, eps;
    std::vector<int32_t> gam_ptr, gam_col, loc, pop_level;
};

// one stored field of a flux region: a running-DFT set (include/chiml_gpu.h chiml_gpu_add_dft)
struct PlanDft
{
    int field = 0, group = 0, every = 1, nfreq = 0, npts = 0, stride = 1;
    uint64_t acc_len = 0;
    std::vector<double> freq;
    std::vector<ChimlDftLine> lines;
    // where the set sits in its flux region (not part of the plan file; used by the flux output, flux_out.cpp)
    int freq_dtc = -1;          // >= 0: a stored field of frequency detector freq_dtc (Inputs::freqDtcs_), not of a flux region
    int surface = 0;            // index into the region's surface list (parallelFluxDTC::fInParam_)
    int role = 0;               // 0 Ej, 1 Ek, 2 Hj, 3 Hk of that surface
    int dir = 0;                // normal of the surface: 0 x, 1 y, 2 z
    bool plus = true;           // the face at loc + sz - 1 (weight +1) or at loc (weight -1)
    int nlines = 0;             // fInParam::sz_[1]
    std::array<int, 3> gloc = {0, 0, 0}, gsz = {0, 0, 0};   // the storage's box in global grid points (parallelStorageFreqDTC::loc_, sz_)
};

// The flattened propagator of one rank ("plan", include/chiml_plan.h)
struct SlabPlan
{
    ChimlPlanGrid grid;
    std::vector<ChimlRun> lists[6][6];      // [ChimlListKind][component]
    std::vector<PlanObject> objects;
    std::vector<PlanCpml> cpml;
    std::vector<PlanSource> sources;
    std::vector<PlanDetector> detectors;
    std::vector<PlanEmitter> emitters;
    std::vector<PlanDft> dfts;
    std::vector<ChimlPlanPeriodic> periodic;   // wrap copies per component (CompCell.PBC)
    bool cplx = false;                         // complex fields (k-point != 0)
    double k_point[3] = {0.0, 0.0, 0.0};
    bool dielectricMatInPML = false;
    bool magMatInPML = false;                  // magnetic material reaches the CPML: the H-side CPML acts on B
    bool has_B = false;                        // B grids exist (magnetic or chiral media)
    int n_mag_poles = 0;
    std::vector<std::array<int32_t, 4>> prev_copy;   // copy2PrevFields_ rows {length, x, y, z} (chiral media)
    struct DipGrid { int comp, pole; std::vector<double> grid; };
    std::vector<DipGrid> dip_grids;                  // dipP_[comp][pole] at the oriented-dipole node cells (REL_TO_NORM orientations), 0 elsewhere

    void write(const std::string& path) const;
};

// Builds the plan of y-slab `rank` of `nranks` (equal-height slabs, mpiInterface::getLocxLocyLocz(int,int,int)).
// referenceSplit: cut the slabs where the reference's cost-weighted decomposition cuts them instead of at equal heights
SlabPlan build_plan(Inputs& IP, int rank, int nranks, int nthreads = 0, bool referenceSplit = false);

// Flux spectra files of a single-rank run (parallelFluxDTC::getFlux, DTC/parallelFlux.hpp:406-540): re[k] / im[k] = accumulators of
// P.dfts[k] (fInReal_ / fInCplx_), nSteps = time steps taken.  Writes <flux name>.dat for every flux region.
// Files of the frequency detectors of a single-rank run (parallelDetectorFREQ_Base::toFile / toMap with incident fields, DTC/parallelDTC_FREQ.hpp:384-583,
// as main.cpp:74-107 calls them; no TFSF source on this path: the incident columns are zero)
void write_freq_detector_files(const Inputs& IP, const SlabPlan& P, const std::vector<std::vector<double>>& re, const std::vector<std::vector<double>>& im, long nSteps);
// incd: the incident-field series of a TFSF run, E_incd_[0..2] then H_incd_[0..2] of the propagator (FDTD_MANAGER/parallelFDTDField.hpp:143-144,
// 1240-1254: two values per step), or nullptr / empty vectors when there is none
void write_flux_files(const Inputs& IP, const SlabPlan& P, const std::vector<std::vector<double>>& re, const std::vector<std::vector<double>>& im, long nSteps,
                      const std::vector<std::vector<std::complex<double>>>* incd = nullptr);

} // namespace chiml_host
