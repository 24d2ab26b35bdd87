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
// usage: chiml_ref <input.json> [--ranks R] [--steps N] [--dump FILE] [--no-output] [--quiet]
//
// dump file layout (little endian): magic "CHIMLDMP" | int32 nranks | then per rank, per grid:
//   int32 rank | char name[16] | int32 lnx, lny, lnz | int32 yStart(global row of local row 1) | float64 data[lnx*lny*lnz]
// with the reference's own index order x + lnx*(z + lnz*y), ghost cells included.
#include <FDTD_MANAGER/parallelFDTDField.hpp>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>

namespace mpi = boost::mpi;

struct Options
{
    std::string input;
    int ranks = 1;
    int steps = -1;
    std::string dump;
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
        g_cells = long(FF.E_[0] ? FF.E_[0]->x() - 2 : FF.E_[2]->x() - 2) * long(FF.E_[0] ? FF.E_[0]->y() - 2 : FF.E_[2]->y() - 2)
                * long( (FF.E_[0] ? FF.E_[0]->z() : FF.E_[2]->z()) == 1 ? 1 : (FF.E_[0] ? FF.E_[0]->z() - 2 : FF.E_[2]->z() - 2) );
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

int main(int argc, char** argv)
{
    Options opt;
    for(int a = 1; a < argc; ++a)
    {
        std::string s = argv[a];
        if(s == "--ranks" && a + 1 < argc) opt.ranks = std::atoi(argv[++a]);
        else if(s == "--steps" && a + 1 < argc) opt.steps = std::atoi(argv[++a]);
        else if(s == "--dump" && a + 1 < argc) opt.dump = argv[++a];
        else if(s == "--no-output") opt.output = false;
        else if(s == "--quiet") opt.quiet = true;
        else if(opt.input.empty()) opt.input = s;
        else { std::fprintf(stderr, "chiml_ref: unknown argument %s\n", s.c_str()); return 2; }
    }
    if(opt.input.empty() || opt.ranks < 1)
    {
        std::fprintf(stderr, "usage: chiml_ref <input.json> [--ranks R] [--steps N] [--dump FILE] [--no-output] [--quiet]\n");
        return 2;
    }
    std::streambuf* oldCout = nullptr;
    std::ostringstream sink;
    if(opt.quiet) oldCout = std::cout.rdbuf(sink.rdbuf());

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
