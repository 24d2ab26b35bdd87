// chiml_plan: JSON input -> plan file(s) (include/chiml_plan.h) without touching a GPU.
// usage: chiml_plan <input.json> <out_prefix> [--ranks N] [--only R] [--threads T] [--split equal|reference]
#include <cstdio>
#include <cstdlib>
#include <string>

#include "setup.hpp"

int main(int argc, char** argv)
{
    if(argc < 3) { std::fprintf(stderr, "usage: chiml_plan <input.json> <out_prefix> [--ranks N] [--only R] [--threads T] [--split equal|reference]\n"); return 2; }
    int nranks = 1, threads = 0, only = -1;
    bool refSplit = false;
    for(int a = 3; a + 1 < argc; a += 2)
    {
        const std::string s = argv[a];
        if(s == "--ranks") nranks = std::atoi(argv[a + 1]);
        else if(s == "--threads") threads = std::atoi(argv[a + 1]);
        else if(s == "--only") only = std::atoi(argv[a + 1]);
        else if(s == "--split") refSplit = std::string(argv[a + 1]) == "reference";
    }
    try
    {
        chiml_host::Json root = chiml_host::read_input_file(argv[1]);
        chiml_host::Inputs IP(root);
        for(int r = 0; r < nranks; ++r)
        {
            if(only >= 0 && r != only) continue;
            chiml_host::SlabPlan P = chiml_host::build_plan(IP, r, nranks, threads, refSplit);
            P.write(std::string(argv[2]) + ".rank" + std::to_string(r) + ".plan");
        }
    }
    catch(std::exception& e) { std::fprintf(stderr, "chiml_plan: %s\n", e.what()); return 1; }
    return 0;
}
