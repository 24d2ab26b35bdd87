// chiml: the host driver of the B200 engine -- the counterpart of the reference's src/main.cpp:11-128.  Reads the same JSON input,
// builds the propagator's lists on the host (build_plan), hands them to the CUDA library through the C ABI (include/chiml_gpu.h),
// runs the time loop on the GPU, and writes the detector / population files the reference writes (TXT detectors,
// DTC/parallelDTC_TXT.cpp:25-55; level populations, ML/QEPopDtc.cpp:37-61).
//
//   chiml <input.json> [--device D] [--steps N] [--rank R --nranks N --rendezvous DIR [--run-id ID]]
//
// One process drives one GPU.  With --nranks > 1 each process takes one y-slab; the halo blobs are exchanged through files in the
// rendezvous directory (any shared directory), after which the slabs talk over NVLink only.  The blob files carry the run id in
// their names (give every launch its own: a blob left behind by a killed run holds dead IPC handles) and are removed at the end.
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>

#include "../../include/chiml_gpu.h"
#include "setup.hpp"

using namespace chiml_host;

static void check(ChimlCtx* ctx, int rc, const char* what)
{
    if(rc != CHIML_OK) throw std::runtime_error(std::string(what) + ": " + chiml_gpu_last_error(ctx));
}

static void make_dirs(const std::string& file)
{
    for(size_t i = 1; i < file.size(); ++i)
        if(file[i] == '/') mkdir(file.substr(0, i).c_str(), 0777);
}

static std::vector<char> read_file_when_ready(const std::string& path)
{
    for(int tries = 0; tries < 6000; ++tries)
    {
        std::ifstream in(path.c_str(), std::ios::binary);
        if(in) { std::vector<char> v((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>()); if(!v.empty()) return v; }
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
    throw std::runtime_error("rendezvous: " + path + " did not appear");
}

// local ghost-inclusive box of a detector's stored field inside this slab (parallelStorageDTC::genDatStruct); false if outside
static bool local_box(const SlabPlan& P, const PlanDetector& d, int32_t loc[3], int32_t sz[3])
{
    const int ly = P.grid.desc.ln[1], lz = P.grid.desc.ln[2];
    const int y0 = d.loc[1] - P.grid.y_start + 1, y1 = y0 + d.sz[1];
    const int lo = std::max(y0, 1), hi = std::min(y1, ly - 1);
    if(hi <= lo) return false;
    loc[0] = d.loc[0] + 1; loc[1] = lo; loc[2] = lz > 1 ? d.loc[2] + 1 : 0;
    sz[0] = d.sz[0]; sz[1] = hi - lo; sz[2] = lz > 1 ? d.sz[2] : 1;
    return true;
}

int main(int argc, char** argv)
{
    std::string input, rendezvous, runId = "0";
    int device = 0, rank = 0, nranks = 1, steps = -1;
    for(int a = 1; a < argc; ++a)
    {
        const std::string s = argv[a];
        auto next = [&]() -> const char* { if(a + 1 >= argc) throw std::runtime_error("missing value after " + s); return argv[++a]; };
        try
        {
            if(s == "--device") device = std::atoi(next());
            else if(s == "--rank") rank = std::atoi(next());
            else if(s == "--nranks") nranks = std::atoi(next());
            else if(s == "--rendezvous") rendezvous = next();
            else if(s == "--run-id") runId = next();
            else if(s == "--steps") steps = std::atoi(next());
            else if(input.empty()) input = s;
            else throw std::runtime_error("unknown argument " + s);
        }
        catch(std::exception& e) { std::fprintf(stderr, "chiml: %s\n", e.what()); return 2; }
    }
    if(input.empty()) { std::fprintf(stderr, "usage: chiml <input.json> [--device D] [--steps N] [--rank R --nranks N --rendezvous DIR [--run-id ID]]\n"); return 2; }
    ChimlCtx* ctx = nullptr;
    ChimlCtx* ctxIm = nullptr;            // complex fields: the imaginary parts (a second propagator over the same lists)
    std::string myBlob;
    try
    {
        Json root = read_input_file(input);
        Inputs IP(root);
        SlabPlan P = build_plan(IP, rank, nranks);
        if(rank == 0) std::cout << "I TOOK ALL THE INPUT PARAMETERS" << std::endl;

        if(chiml_gpu_create(&P.grid.desc, device, &ctx) != CHIML_OK) throw std::runtime_error(std::string("chiml_gpu_create: ") + chiml_gpu_last_error(nullptr));
        // update lists and material constants of every kind (B grids first: the H components take D-type lists only with them)
        auto handOver = [&P](ChimlCtx* c) {
            if(P.has_B) check(c, chiml_gpu_set_magnetic(c, 1, P.magMatInPML ? 1 : 0), "set_magnetic");
            for(int kind = 0; kind < 6; ++kind)
                for(int comp = 0; comp < 6; ++comp)
                    if(!P.lists[kind][comp].empty())
                        check(c, chiml_gpu_set_update_list(c, kind, comp, P.lists[kind][comp].data(), P.lists[kind][comp].size()), "set_update_list");
            bool anyChi = false;
            for(size_t o = 0; o < P.objects.size(); ++o)
            {
                const PlanObject& ob = P.objects[o];
                check(c, chiml_gpu_set_object(c, (int)o, ob.npoles, ob.alpha.data(), ob.xi.data(), ob.gamma.data(), ob.use_or_dip, ob.dip.data()), "set_object");
                if(P.has_B) check(c, chiml_gpu_set_object_magnetic(c, (int)o, (int)ob.magGamma.size(), ob.magAlpha.data(), ob.magXi.data(), ob.magGamma.data()), "set_object_magnetic");
                if(!ob.chiGamma.empty())
                {
                    anyChi = true;
                    check(c, chiml_gpu_set_object_chiral(c, (int)o, (int)ob.chiGamma.size(), ob.chiAlpha.data(), ob.chiXi.data(), ob.chiGamma.data(), ob.chiGammaPrev.data()), "set_object_chiral");
                }
            }
            for(const SlabPlan::DipGrid& dg : P.dip_grids) check(c, chiml_gpu_set_dip_grid(c, dg.comp, dg.pole, dg.grid.data()), "set_dip_grid");
            if(anyChi) check(c, chiml_gpu_set_prev_copy(c, reinterpret_cast<const int32_t*>(P.prev_copy.data()), P.prev_copy.size()), "set_prev_copy");
        };
        handOver(ctx);
        for(const PlanCpml& c : P.cpml)
            check(ctx, chiml_gpu_set_cpml(ctx, c.comp, c.part, c.has_psi, c.psi.data(), c.psi.size(), c.grid.data(), c.grid.size()), "set_cpml");
        for(const PlanSource& s : P.sources) check(ctx, chiml_gpu_add_source(ctx, s.field, s.loc, s.sz, nullptr), "add_source");
        std::vector<int> detSlot(P.detectors.size(), -1);
        for(size_t d = 0; d < P.detectors.size(); ++d)
        {
            int32_t loc[3], sz[3];
            if(local_box(P, P.detectors[d], loc, sz)) check(ctx, chiml_gpu_add_detector(ctx, P.detectors[d].field, loc, sz, P.detectors[d].every, &detSlot[d]), "add_detector");
        }
        if(nranks > 1) check(ctx, chiml_gpu_set_ordip_pole_count(ctx, P.grid.n_ordip_poles), "set_ordip_pole_count");
        for(const ChimlPlanPeriodic& pp : P.periodic) check(ctx, chiml_gpu_set_periodic(ctx, pp.comp, &pp.wrap), "set_periodic");
        for(const PlanEmitter& e : P.emitters)
        {
            ChimlEmitterDesc d;
            std::memset(&d, 0, sizeof(d));
            d.object = e.object;
            d.nlevel = e.nlevel; d.nsys = e.nsys; d.nemit = e.nemit;
            for(int k = 0; k < 3; ++k) { d.box_lo[k] = e.box_lo[k]; d.box_n[k] = e.box_n[k]; }
            d.dt = e.dt; d.inv_hbar = e.inv_hbar; d.na = e.na;
            d.h0 = e.h0.data(); d.weight = e.weight.data(); d.mu = e.mu.data();
            d.gam_ptr = e.gam_ptr.data(); d.gam_col = e.gam_col.data(); d.gam_val = e.gam_val.data();
            d.loc = e.loc.data(); d.eps = e.eps.data();
            d.npop = (int)e.pop_level.size(); d.pop_level = e.pop_level.data(); d.pop_every = e.pop_every; d.npoints = e.npoints;
            check(ctx, chiml_gpu_add_emitters(ctx, &d, nullptr), "add_emitters");
        }
        std::vector<int> dftSlot(P.dfts.size(), -1);
        for(size_t q = 0; q < P.dfts.size(); ++q)
        {
            const PlanDft& d = P.dfts[q];
            check(ctx, chiml_gpu_add_dft(ctx, d.field, d.group, d.every, d.nfreq, d.npts, d.stride, d.lines.data(), d.lines.size(), d.acc_len, &dftSlot[q]), "add_dft");
        }
        check(ctx, chiml_gpu_commit(ctx), "commit");
        if(P.cplx)
        {
            // Bloch-periodic run: the imaginary parts of every array are a second context set up from the same lists (no detectors of its own:
            // the reference's TXT / BIN writers print the real part of the collected field, DTC/parallelDTC_TXT.cpp:76-93), bound to the first
            for(const PlanDetector& d : P.detectors)
                if(d.type == (int)DTCTYPE::EPOW || d.type == (int)DTCTYPE::HPOW) throw std::runtime_error("power detectors of a complex-field run are outside the covered hot path");
            if(chiml_gpu_create(&P.grid.desc, device, &ctxIm) != CHIML_OK) throw std::runtime_error(std::string("chiml_gpu_create: ") + chiml_gpu_last_error(nullptr));
            handOver(ctxIm);
            for(const PlanCpml& c : P.cpml)
                check(ctxIm, chiml_gpu_set_cpml(ctxIm, c.comp, c.part, c.has_psi, c.psi.data(), c.psi.size(), c.grid.data(), c.grid.size()), "set_cpml");
            for(const PlanSource& s : P.sources) check(ctxIm, chiml_gpu_add_source(ctxIm, s.field, s.loc, s.sz, nullptr), "add_source");
            for(const ChimlPlanPeriodic& pp : P.periodic) check(ctxIm, chiml_gpu_set_periodic(ctxIm, pp.comp, &pp.wrap), "set_periodic");
            check(ctxIm, chiml_gpu_commit(ctxIm), "commit");
            check(ctx, chiml_gpu_bind_imag(ctx, ctxIm, P.k_point), "bind_imag");
        }

        if(nranks > 1)
        {
            if(rendezvous.empty()) throw std::runtime_error("--nranks > 1 needs --rendezvous DIR");
            size_t n = 0;
            check(ctx, chiml_gpu_halo_export(ctx, nullptr, 0, &n), "halo_export");
            std::vector<char> mine(n);
            check(ctx, chiml_gpu_halo_export(ctx, mine.data(), n, &n), "halo_export");
            const std::string stem = rendezvous + "/halo." + runId + ".";
            const std::string my = stem + std::to_string(rank);
            myBlob = my;
            std::remove(my.c_str());
            { std::ofstream out((my + ".tmp").c_str(), std::ios::binary); out.write(mine.data(), (std::streamsize)n); }
            std::rename((my + ".tmp").c_str(), my.c_str());
            std::vector<char> lower, upper;
            const bool ring = !P.periodic.empty();      // a periodic run closes the slabs into a ring: slab nranks - 1 below slab 0
            if(rank > 0 || ring) lower = read_file_when_ready(stem + std::to_string((rank + nranks - 1) % nranks));
            if(rank + 1 < nranks || ring) upper = read_file_when_ready(stem + std::to_string((rank + 1) % nranks));
            check(ctx, chiml_gpu_halo_bind(ctx, lower.empty() ? nullptr : lower.data(), lower.size(), upper.empty() ? nullptr : upper.data(), upper.size()), "halo_bind");
        }
        if(rank == 0) std::cout << "made FF" << std::endl;

        const int nSteps = steps >= 0 ? steps : P.grid.n_steps;
        const int nsrc = (int)P.sources.size();
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<double> amp, twiddles;
        // complex twiddles per step: the flux regions that have a stored field on this slab, region order (chiml_gpu_step_n_dft)
        // (the groups of the frequency detectors follow those of the flux regions)
        const size_t nGroups = IP.fluxes_.size() + IP.freqDtcs_.size();
        auto groupFreqs = [&](size_t g) -> const std::vector<double>& { return g < IP.fluxes_.size() ? IP.fluxes_[g].freqs : IP.freqDtcs_[g - IP.fluxes_.size()].freqs; };
        std::vector<char> fluxHere(nGroups, 0);
        for(const PlanDft& d : P.dfts) fluxHere[d.group] = 1;
        size_t ntw = 0;
        for(size_t ff = 0; ff < nGroups; ++ff) if(fluxHere[ff]) ntw += groupFreqs(ff).size();
        double tFlux = 0.0;                               // the reference's tcur_ (tcur_ += dt_, parallelFDTDField.hpp:1290)
        // detector and population samples are drained from the device rings after every chunk of steps (read, then consume), so the
        // rings keep their initial size however long the run is
        std::vector<std::vector<double>> detData(P.detectors.size());
        std::vector<std::vector<std::vector<double>>> popData(P.emitters.size());
        for(size_t q = 0; q < P.emitters.size(); ++q) popData[q].resize(P.emitters[q].pop_level.size());
        auto drain = [&]() {
            for(size_t d = 0; d < P.detectors.size(); ++d)
            {
                if(detSlot[d] < 0) continue;
                int32_t loc[3], sz[3];
                local_box(P, P.detectors[d], loc, sz);
                const size_t len = (size_t)sz[0] * sz[1] * sz[2];
                size_t ns = 0;
                check(ctx, chiml_gpu_read_detector(ctx, detSlot[d], nullptr, 0, &ns), "read_detector");
                if(ns == 0) continue;
                const size_t have = detData[d].size();
                detData[d].resize(have + ns * len);
                check(ctx, chiml_gpu_read_detector(ctx, detSlot[d], detData[d].data() + have, ns, &ns), "read_detector");
                check(ctx, chiml_gpu_consume_detector(ctx, detSlot[d], have / len + ns), "consume_detector");
            }
            for(size_t q = 0; q < P.emitters.size(); ++q)
            {
                size_t total = 0;
                for(size_t dd = 0; dd < P.emitters[q].pop_level.size(); ++dd)
                {
                    size_t ns = 0;
                    check(ctx, chiml_gpu_read_population(ctx, (int)q, (int)dd, nullptr, 0, &ns), "read_population");
                    const size_t have = popData[q][dd].size();
                    popData[q][dd].resize(have + 2 * ns);
                    if(ns) check(ctx, chiml_gpu_read_population(ctx, (int)q, (int)dd, popData[q][dd].data() + have, ns, &ns), "read_population");
                    total = have / 2 + ns;
                }
                if(!P.emitters[q].pop_level.empty()) check(ctx, chiml_gpu_consume_population(ctx, (int)q, total), "consume_population");
            }
        };
        for(int done = 0; done < nSteps;)
        {
            const int n = std::min(256, nSteps - done);
            amp.assign((size_t)n * std::max(nsrc, 1), 0.0);
            for(int k = 0; k < n; ++k)
                for(int q = 0; q < nsrc; ++q)
                    if((size_t)(done + k) < P.sources[q].amp.size()) amp[(size_t)k * nsrc + q] = P.sources[q].amp[done + k];
            if(P.cplx)
            {
                std::vector<double> ampIm((size_t)n * std::max(nsrc, 1), 0.0);
                for(int k = 0; k < n; ++k)
                    for(int q = 0; q < nsrc; ++q)
                        if((size_t)(done + k) < P.sources[q].amp_im.size()) ampIm[(size_t)k * nsrc + q] = P.sources[q].amp_im[done + k];
                check(ctx, chiml_gpu_step_n_cplx(ctx, n, nsrc ? amp.data() : nullptr, nsrc ? ampIm.data() : nullptr), "step_n_cplx");
            }
            else if(P.dfts.empty()) check(ctx, chiml_gpu_step_n(ctx, n, nsrc ? amp.data() : nullptr), "step_n");
            else
            {
                // fftFact_ = exp(i * (-t * freq)) with the time after the step (parallelFluxDTC::fieldIn, DTC/parallelFlux.hpp:298)
                twiddles.resize((size_t)n * ntw * 2);
                for(int k = 0; k < n; ++k)
                {
                    tFlux += P.grid.desc.dt;
                    size_t j = (size_t)k * ntw;
                    for(size_t ff = 0; ff < nGroups; ++ff)
                        for(double freq : groupFreqs(ff))
                        {
                            if(!fluxHere[ff]) break;
                            const std::complex<double> w = std::exp(std::complex<double>(0.0, -1.0 * tFlux * freq));
                            twiddles[2 * j] = w.real(); twiddles[2 * j + 1] = w.imag(); ++j;
                        }
                }
                check(ctx, chiml_gpu_step_n_dft(ctx, n, nsrc ? amp.data() : nullptr, twiddles.data()), "step_n_dft");
            }
            done += n;
            drain();
        }
        check(ctx, chiml_gpu_sync(ctx), "sync");
        drain();                                          // nSteps == 0: the t = 0 samples
        if(!myBlob.empty()) std::remove(myBlob.c_str());  // every neighbour has bound long ago: the slabs step in lock-step
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::cout << std::setw(9) << sec << "\t" << rank << "\t" << P.grid.y_start << std::endl;   // main.cpp:59-65 (wall clock here)

        // ---- detector files (DTC/parallelDTC_TXT.cpp:25-55); a detector cut by a slab boundary is written by each slab for its part
        for(size_t d = 0; d < P.detectors.size(); ++d)
        {
            if(detSlot[d] < 0) continue;
            const PlanDetector& pd = P.detectors[d];
            const DetectorInput& di = IP.detectors_[pd.detector];
            if(di.cls != DTCCLASS::TXT && di.cls != DTCCLASS::BIN && di.cls != DTCCLASS::COUT) continue;
            int32_t loc[3], sz[3];
            local_box(P, pd, loc, sz);
            const size_t len = (size_t)sz[0] * sz[1] * sz[2];
            // outputCollectFunction_ (DTC/parallelDTCOutputFxn.hpp): field value or |field|^2, added twice with half the factor each (the cell
            // and its Yee-offset partner: single-component detectors have offset 0)
            const bool power = di.type == DTCTYPE::EPOW || di.type == DTCTYPE::HPOW;
            auto collect = [&](const double f) {
                double point = 0.0;
                const double v = power ? std::pow(std::abs(f), 2.0) : f;
                point = point + (pd.conv / 2.0) * v;
                point = point + (pd.conv / 2.0) * v;
                return point;
            };
            const std::vector<double>& data = detData[d];
            const size_t ns = data.size() / len;
            std::string name = di.name;
            if(nranks > 1) name += ".rank" + std::to_string(rank);
            make_dirs(name);
            // sample times: the reference passes tcur_, which it accumulates step by step (tcur_ += dt_, parallelFDTDField.hpp:1290)
            std::vector<double> times(1, 0.0);
            {
                double tcur = 0.0;
                for(long st = 1; times.size() < ns; ++st) { tcur += P.grid.desc.dt; if(st % pd.every == 0) times.push_back(tcur); }
            }
            if(di.cls == DTCCLASS::BIN)
            {
                // DTC/parallelDTC_BIN.cpp:12-52: int sz[3], int loc[3], then per sample: double t, rows of sz[0] doubles, z outer, y inner
                std::ofstream out(name.c_str(), std::ios::out | std::ios::binary);
                const int32_t hsz[3] = {sz[0], sz[1], sz[2]}, hloc[3] = {pd.loc[0], pd.loc[1] + (loc[1] - (pd.loc[1] - P.grid.y_start + 1)), pd.loc[2]};
                out.write(reinterpret_cast<const char*>(hsz), sizeof(hsz));
                out.write(reinterpret_cast<const char*>(hloc), sizeof(hloc));
                std::vector<double> rowv((size_t)sz[0]);
                for(size_t s = 0; s < ns; ++s)
                {
                    const double tt = times[s] * pd.t_conv;
                    out.write(reinterpret_cast<const char*>(&tt), sizeof(tt));
                    for(int kk = 0; kk < sz[2]; ++kk)
                        for(int jj = 0; jj < sz[1]; ++jj)
                        {
                            for(int ii = 0; ii < sz[0]; ++ii)
                            {
                                rowv[(size_t)ii] = collect(data[s * len + (size_t)ii + (size_t)sz[0] * ((size_t)kk + (size_t)sz[2] * (size_t)jj)]);
                            }
                            out.write(reinterpret_cast<const char*>(rowv.data()), (std::streamsize)(rowv.size() * sizeof(double)));
                        }
                }
                continue;
            }
            double rsl[3];
            for(int k = 0; k < 3; ++k)
            {
                // realSpaceLoc_ (DTC/parallelDTC.hpp:61): d * (loc - (n_vec - 2 np - n_vec % 2) / 2) with the ghost-inclusive global n_vec;
                // a 2-D grid has n_vec(z) = 1 and np(z) = 1, which makes the bracket -1
                const int n = P.grid.n_global[k];
                const int half = (k == 2 && P.grid.desc.ln[2] == 1) ? -1 : (n - n % 2) / 2;
                rsl[k] = P.grid.desc.d[k] * (pd.loc[k] - half);
            }
            if(di.SI) for(int k = 0; k < 3; ++k) rsl[k] *= IP.a_ * IP.a_;     // scaled twice in the reference (parallelDTC.hpp:69,82)
            if(di.cls == DTCCLASS::COUT)
            {
                // DTC/parallelDTC_COUT.cpp:18-41: to the console with the stream's default formatting, z and y descending, one line per row
                // (the reference prints while it steps; here the samples of the whole run follow the run)
                for(size_t s = 0; s < ns; ++s)
                {
                    std::cout << times[s] * pd.t_conv << "\t" << rsl[0] << "\t" << rsl[1] << '\t' << rsl[2] << '\t' << std::endl;
                    for(int kk = sz[2] - 1; kk >= 0; --kk)
                        for(int jj = sz[1] - 1; jj >= 0; --jj)
                        {
                            for(int ii = 0; ii < sz[0]; ++ii)
                                std::cout << "\t" << collect(data[s * len + (size_t)ii + (size_t)sz[0] * ((size_t)kk + (size_t)sz[2] * (size_t)jj)]);
                            std::cout << std::endl;
                        }
                }
                continue;
            }
            std::ofstream out(name.c_str());
            out << "# time\tx\ty\tz\tfield" << std::endl;
            for(size_t s = 0; s < ns; ++s)
            {
                const double t = times[s];
                out << std::setprecision(6) << t * pd.t_conv << "\t" << rsl[0] << "\t" << rsl[1] << "\t" << rsl[2];
                // sample layout: x fastest, then z, then y; the reference prints y outermost, then z, then x
                // (the complex-field writer prints the real part of the collected value at the stream's default 6 digits, unpadded:
                // DTC/parallelDTC_TXT.cpp:76-93 against :25-55)
                if(P.cplx) for(size_t i = 0; i < len; ++i) out << "\t" << collect(data[s * len + i]);
                else for(size_t i = 0; i < len; ++i) out << "\t" << std::setw(24) << std::setprecision(18) << collect(data[s * len + i]);
                out << '\n';
            }
        }
        // ---- frequency-domain fields of the flux regions: one file per region and slab holding every stored field's accumulators
        // (fInReal_ / fInCplx_ of parallelStorageFreqDTCReal) in the order parallelFluxDTC::fieldIn walks them.  The Poynting-vector
        // integration of getFlux() is post-processing on these arrays and is not part of the time-stepping path.
        std::vector<std::vector<double>> dftRe(P.dfts.size()), dftIm(P.dfts.size());
        // (the frequency detectors' groups follow the flux regions'; their accumulator files are <detector file>.dft)
        for(size_t ff = 0; ff < IP.fluxes_.size() + IP.freqDtcs_.size(); ++ff)
        {
            const bool isFlux = ff < IP.fluxes_.size();
            const std::vector<double>& freqs = isFlux ? IP.fluxes_[ff].freqs : IP.freqDtcs_[ff - IP.fluxes_.size()].freqs;
            std::string name = (isFlux ? IP.fluxes_[ff].name : IP.freqDtcs_[ff - IP.fluxes_.size()].name) + ".dft";
            if(nranks > 1) name += ".rank" + std::to_string(rank);
            make_dirs(name);
            std::ofstream out(name.c_str(), std::ios::out | std::ios::binary);
            const char magic[8] = {'C', 'H', 'I', 'M', 'L', 'D', 'F', 'T'};
            out.write(magic, 8);
            int32_t nsets = 0;
            for(const PlanDft& d : P.dfts) nsets += d.group == (int)ff;
            const int32_t hdr[2] = {nsets, (int32_t)freqs.size()};
            out.write(reinterpret_cast<const char*>(hdr), sizeof(hdr));
            out.write(reinterpret_cast<const char*>(freqs.data()), (std::streamsize)(freqs.size() * sizeof(double)));
            for(size_t q = 0; q < P.dfts.size(); ++q)
            {
                const PlanDft& d = P.dfts[q];
                if(d.group != (int)ff) continue;
                std::vector<double> re(d.acc_len), im(d.acc_len);
                check(ctx, chiml_gpu_download_dft(ctx, dftSlot[q], re.data(), im.data()), "download_dft");
                const int32_t sh[4] = {d.field, d.npts, (int32_t)d.lines.size(), d.every};
                const uint64_t len = d.acc_len;
                out.write(reinterpret_cast<const char*>(sh), sizeof(sh));
                out.write(reinterpret_cast<const char*>(&len), sizeof(len));
                out.write(reinterpret_cast<const char*>(re.data()), (std::streamsize)(len * sizeof(double)));
                out.write(reinterpret_cast<const char*>(im.data()), (std::streamsize)(len * sizeof(double)));
                dftRe[q].swap(re); dftIm[q].swap(im);
            }
        }
        // flux spectra (parallelFluxDTC::getFlux): one process holds whole surfaces; several slabs leave their accumulator files
        if(nranks == 1 && !IP.fluxes_.empty()) write_flux_files(IP, P, dftRe, dftIm, nSteps);
        if(nranks == 1) write_freq_detector_files(IP, P, dftRe, dftIm, nSteps);
        // ---- level populations (ML/QEPopDtc.cpp:37-61)
        for(size_t q = 0; q < P.emitters.size(); ++q)
        {
            const PlanEmitter& e = P.emitters[q];
            const QEInput& qi = IP.qes_[e.object];
            for(size_t dd = 0; dd < e.pop_level.size(); ++dd)
            {
                const std::vector<double>& pop = popData[q][dd];
                const size_t ns = pop.size() / 2;
                std::string name = qi.dtcPopFiles[dd];
                if(nranks > 1) name += ".rank" + std::to_string(rank);
                make_dirs(name);
                std::ofstream out(name.c_str());
                for(size_t tt = 0; tt < ns; ++tt)
                {
                    const double re = pop[2 * tt], im = pop[2 * tt + 1];
                    out << std::setw(9) << std::setprecision(9) << tt * e.pop_every * e.dt << "\t" << std::setw(16) << std::setprecision(16) << re << "\t"
                        << std::setw(16) << std::setprecision(16) << im << "\t" << std::setw(16) << std::setprecision(16) << std::abs(std::complex<double>(re, im)) << std::endl;
                }
            }
        }
        if(ctxIm) chiml_gpu_destroy(ctxIm);
        chiml_gpu_destroy(ctx);
    }
    catch(std::exception& e)
    {
        std::fprintf(stderr, "chiml: %s\n", e.what());
        if(ctxIm) chiml_gpu_destroy(ctxIm);
        if(ctx) chiml_gpu_destroy(ctx);
        return 1;
    }
    return 0;
}
