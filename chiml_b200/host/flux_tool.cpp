// chiml_flux: the flux spectra files of a single-rank run from the accumulator files the driver wrote.
// usage: chiml_flux <input.json> [--steps N]      (run in the directory the relative output names of the input refer to)
// Reads <flux name>.dft of every flux region (format: chiml_b200/host/main.cpp), writes <flux name>.dat like the reference's
// parallelFluxDTC::getFlux.  The driver `chiml` calls the same function at the end of a single-rank run; this tool exists so that
// accumulators of several slabs can be merged first, and so that the post-processing is testable without a GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "setup.hpp"

int main(int argc, char** argv)
{
    if(argc < 2) { std::fprintf(stderr, "usage: chiml_flux <input.json> [--steps N]\n"); return 2; }
    long steps = -1;
    for(int a = 2; a + 1 < argc; a += 2) if(std::string(argv[a]) == "--steps") steps = std::atol(argv[a + 1]);
    try
    {
        using namespace chiml_host;
        Json root = read_input_file(argv[1]);
        Inputs IP(root);
        SlabPlan P = build_plan(IP, 0, 1);
        std::vector<std::vector<double>> re(P.dfts.size()), im(P.dfts.size());
        for(size_t ff = 0; ff < IP.fluxes_.size(); ++ff)
        {
            const std::string name = IP.fluxes_[ff].name + ".dft";
            std::ifstream in(name.c_str(), std::ios::binary);
            if(!in) throw std::runtime_error("cannot read " + name);
            char magic[8]; int32_t hdr[2];
            in.read(magic, 8); in.read(reinterpret_cast<char*>(hdr), sizeof(hdr));
            if(std::memcmp(magic, "CHIMLDFT", 8) != 0) throw std::runtime_error(name + " is not an accumulator file");
            in.seekg((std::streamoff)hdr[1] * 8, std::ios::cur);
            for(size_t q = 0; q < P.dfts.size(); ++q)
            {
                if(P.dfts[q].group != (int)ff) continue;
                int32_t sh[4]; uint64_t len = 0;
                in.read(reinterpret_cast<char*>(sh), sizeof(sh)); in.read(reinterpret_cast<char*>(&len), sizeof(len));
                if(!in || len != P.dfts[q].acc_len || sh[0] != P.dfts[q].field) throw std::runtime_error(name + " does not match the input's flux regions");
                re[q].resize(len); im[q].resize(len);
                in.read(reinterpret_cast<char*>(re[q].data()), (std::streamsize)(len * 8));
                in.read(reinterpret_cast<char*>(im[q].data()), (std::streamsize)(len * 8));
            }
        }
        write_flux_files(IP, P, re, im, steps >= 0 ? steps : P.grid.n_steps);
    }
    catch(std::exception& e) { std::fprintf(stderr, "chiml_flux: %s\n", e.what()); return 1; }
    return 0;
}
