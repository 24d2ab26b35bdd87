// chiml_flux: the flux spectra files of a single-rank run from the accumulator files the driver wrote.
// usage: chiml_flux <input.json> [--steps N] [--ranks R] [--incd FILE]   (FILE: the incident-field series of a TFSF run, "CHIMLINC", int32 n, then six
// series Ex Ey Ez Hx Hy Hz of n complex values, as the drop-in / the reference driver's --incd-dump writes them)   (run in the directory the relative output names of the input refer to)
// Reads <flux name>.dft of every flux region (format: chiml_b200/host/main.cpp), writes <flux name>.dat like the reference's
// parallelFluxDTC::getFlux.  The driver `chiml` calls the same function at the end of a single-rank run; this tool exists so that
// accumulators of several slabs can be merged first, and so that the post-processing is testable without a GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <array>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "setup.hpp"

namespace {
using namespace chiml_host;

// accumulators of the stored fields of flux region ff from one accumulator file, in the order of P.dfts
void read_region(const std::string& name, const SlabPlan& P, int ff, std::vector<std::vector<double>>& re, std::vector<std::vector<double>>& im)
{
    std::ifstream in(name.c_str(), std::ios::binary);
    if(!in) throw std::runtime_error("cannot read " + name);
    char magic[8]; int32_t hdr[2];
    in.read(magic, 8); in.read(reinterpret_cast<char*>(hdr), sizeof(hdr));
    if(std::memcmp(magic, "CHIMLDFT", 8) != 0) throw std::runtime_error(name + " is not an accumulator file");
    in.seekg((std::streamoff)hdr[1] * 8, std::ios::cur);
    for(size_t q = 0; q < P.dfts.size(); ++q)
    {
        if(P.dfts[q].group != ff) continue;
        int32_t sh[4]; uint64_t len = 0;
        in.read(reinterpret_cast<char*>(sh), sizeof(sh)); in.read(reinterpret_cast<char*>(&len), sizeof(len));
        if(!in || len != P.dfts[q].acc_len || sh[0] != P.dfts[q].field) throw std::runtime_error(name + " does not match the input's flux regions");
        re[q].resize(len); im[q].resize(len);
        in.read(reinterpret_cast<char*>(re[q].data()), (std::streamsize)(len * 8));
        in.read(reinterpret_cast<char*>(im[q].data()), (std::streamsize)(len * 8));
    }
}

// global grid point of (line l, point i) of a stored field
std::array<long, 3> point_of(const SlabPlan& P, const PlanDft& d, size_t l, int i)
{
    const long lnx = P.grid.desc.ln[0], lnz = P.grid.desc.ln[2];
    const long gidx = (long)d.lines[l].ind + (long)i * d.stride;
    const long x = gidx % lnx, row = gidx / lnx;
    return {x, row / lnz + P.grid.y_start, row % lnz};
}
} // namespace

int main(int argc, char** argv)
{
    if(argc < 2) { std::fprintf(stderr, "usage: chiml_flux <input.json> [--steps N] [--ranks R]\n"); return 2; }
    long steps = -1;
    int nranks = 1;
    std::string incdFile;
    for(int a = 2; a + 1 < argc; a += 2)
    {
        if(std::string(argv[a]) == "--steps") steps = std::atol(argv[a + 1]);
        else if(std::string(argv[a]) == "--incd") incdFile = argv[a + 1];
        else if(std::string(argv[a]) == "--ranks") nranks = std::atoi(argv[a + 1]);
    }
    try
    {
        Json root = read_input_file(argv[1]);
        Inputs IP(root, true);
        SlabPlan P = build_plan(IP, 0, 1);
        std::vector<std::vector<std::complex<double>>> incd;
        if(!incdFile.empty())
        {
            std::ifstream in(incdFile.c_str(), std::ios::binary);
            char magic[8]; int32_t n = 0;
            in.read(magic, 8); in.read(reinterpret_cast<char*>(&n), 4);
            if(!in || std::memcmp(magic, "CHIMLINC", 8) != 0 || n < 0) throw std::runtime_error(incdFile + " is not an incident-series file");
            incd.assign(6, std::vector<std::complex<double>>((size_t)n));
            for(auto& v : incd) in.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(v.size() * sizeof(std::complex<double>)));
            if(!in) throw std::runtime_error(incdFile + " is truncated");
        }
        std::vector<std::vector<double>> re(P.dfts.size()), im(P.dfts.size());
        if(nranks == 1)
        {
            for(size_t ff = 0; ff < IP.fluxes_.size(); ++ff) read_region(IP.fluxes_[ff].name + ".dft", P, (int)ff, re, im);
            for(size_t k = 0; k < IP.freqDtcs_.size(); ++k) read_region(IP.freqDtcs_[k].name + ".dft", P, (int)(IP.fluxes_.size() + k), re, im);
        }
        else
        {
            // several slabs: every slab wrote the accumulators of its parts of the surfaces (<name>.dft.rank<r>); an accumulator is
            // identified by (region, surface, role, box of the stored field) and the global grid point, whatever the decomposition
            for(size_t q = 0; q < P.dfts.size(); ++q) { re[q].assign(P.dfts[q].acc_len, 0.0); im[q].assign(P.dfts[q].acc_len, 0.0); }
            typedef std::array<long, 8> SetKey;      // group, surface, role, gloc[3], pad
            auto key_of = [](const PlanDft& d) { return SetKey{{d.group, d.surface, d.role, d.gloc[0], d.gloc[1], d.gloc[2], d.gsz[0] + 1000L * d.gsz[1], d.gsz[2]}}; };
            std::map<SetKey, size_t> whole;
            std::vector<std::map<std::array<long, 3>, size_t>> where(P.dfts.size());     // global point -> accumulator index of frequency 0
            for(size_t q = 0; q < P.dfts.size(); ++q)
            {
                const PlanDft& d = P.dfts[q];
                whole[key_of(d)] = q;
                for(size_t l = 0; l < (size_t)d.nlines && l < d.lines.size(); ++l)
                    for(int i = 0; i < d.npts; ++i) where[q][point_of(P, d, l, i)] = (size_t)d.lines[l].out + (size_t)d.nfreq * i;
            }
            std::vector<size_t> filled(P.dfts.size(), 0);
            for(int r = 0; r < nranks; ++r)
            {
                SlabPlan S = build_plan(IP, r, nranks);
                std::vector<std::vector<double>> sre(S.dfts.size()), sim(S.dfts.size());
                for(size_t ff = 0; ff < IP.fluxes_.size(); ++ff)
                {
                    bool here = false;
                    for(const PlanDft& d : S.dfts) here = here || d.group == (int)ff;
                    if(here) read_region(IP.fluxes_[ff].name + ".dft.rank" + std::to_string(r), S, (int)ff, sre, sim);
                }
                for(size_t q = 0; q < S.dfts.size(); ++q)
                {
                    const PlanDft& d = S.dfts[q];
                    auto it = whole.find(key_of(d));
                    if(it == whole.end()) throw std::runtime_error("a slab holds a stored field the whole grid does not have");
                    const size_t Q = it->second;
                    for(size_t l = 0; l < (size_t)d.nlines && l < d.lines.size(); ++l)
                        for(int i = 0; i < d.npts; ++i)
                        {
                            auto w = where[Q].find(point_of(S, d, l, i));
                            if(w == where[Q].end()) throw std::runtime_error("a slab holds an accumulator the whole grid does not have");
                            const size_t src = (size_t)d.lines[l].out + (size_t)d.nfreq * i;
                            for(int f = 0; f < d.nfreq; ++f) { re[Q][w->second + f] = sre[q][src + f]; im[Q][w->second + f] = sim[q][src + f]; }
                            ++filled[Q];
                        }
                }
            }
            for(size_t q = 0; q < P.dfts.size(); ++q)
                if(filled[q] != where[q].size()) throw std::runtime_error("the slabs' accumulator files do not cover a flux surface");
        }
        write_flux_files(IP, P, re, im, steps >= 0 ? steps : P.grid.n_steps, incd.empty() ? nullptr : &incd);
        if(nranks == 1) write_freq_detector_files(IP, P, re, im, steps >= 0 ? steps : P.grid.n_steps);
        else if(!IP.freqDtcs_.empty()) throw std::runtime_error("frequency-detector files are written by single-rank runs");
    }
    catch(std::exception& e) { std::fprintf(stderr, "chiml_flux: %s\n", e.what()); return 1; }
    return 0;
}
