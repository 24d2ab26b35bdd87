// y-slab ghost-row exchange over NVLink peer memory (protocol: chiml_b200/slab.py; reference GRID/parallelGrid.hpp:738-770,
// ML/parallelQE.hpp:618-645,694-715).  One process per GPU; the neighbours' buffers are mapped with CUDA IPC.  The sender pushes:
// a kernel on the halo stream copies its boundary row straight into the neighbour's ghost row with peer stores and, when the
// last block is through, publishes the step number in a flag word in the neighbour's memory.  The receiver's compute stream
// blocks on that flag with a one-thread wait kernel just before the boundary tiles that read the ghost row; everything else of
// the half step does not depend on it and keeps the SMs busy meanwhile.
#pragma once

namespace chiml {

struct HaloSeg { const double* src; double* dst; long n; };   // one contiguous run of doubles

struct HaloPushArgs
{
    HaloSeg seg[4];
    int nseg;
    int* peer_flag[2];         // flag words (in the receiver's memory) to publish, and how many
    int nflag;
    int value;
    unsigned* counter;         // zero-initialised block counter (reset by the last block)
};

__global__ void __launch_bounds__(256) k_halo_push(const __grid_constant__ HaloPushArgs a)
{
    for(int s = 0; s < a.nseg; ++s)
    {
        const long n = a.seg[s].n;
        const bool vec = ((reinterpret_cast<uintptr_t>(a.seg[s].src) | reinterpret_cast<uintptr_t>(a.seg[s].dst)) & 15) == 0 && (n & 1) == 0;
        if(vec)
        {
            const double2* __restrict__ src = reinterpret_cast<const double2*>(a.seg[s].src);
            double2* __restrict__ dst = reinterpret_cast<double2*>(a.seg[s].dst);
            for(long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n / 2; i += (long)gridDim.x * blockDim.x) dst[i] = src[i];
        }
        else
            for(long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) a.seg[s].dst[i] = a.seg[s].src[i];
    }
    __threadfence_system();
    __syncthreads();
    if(threadIdx.x == 0)
    {
        const unsigned done = atomicAdd(a.counter, 1u) + 1u;
        if(done == gridDim.x)
        {
            *a.counter = 0u;
            __threadfence_system();
            for(int f = 0; f < a.nflag; ++f) *reinterpret_cast<volatile int*>(a.peer_flag[f]) = a.value;
            __threadfence_system();
        }
    }
}

// node-centred oriented-dipole P_y, row 1 of this slab, expanded from the compact row spans into the dense ghost row of the slab below
struct NodePushArgs
{
    const double* pool[MAX_POLES];
    double* dst[MAX_POLES];
    int npoles;
    const int32_t* sp_xmin; const int32_t* sp_xmax; const int64_t* sp_base;
    int lx, lz;
    int* peer_flag; int value; unsigned* counter;
};
__global__ void __launch_bounds__(256) k_halo_push_nodes(const __grid_constant__ NodePushArgs a)
{
    const long n = (long)a.lx * a.lz;
    for(long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    {
        const int x = (int)(i % a.lx), z = (int)(i / a.lx);
        const long row = z + (long)a.lz * 1;
        // a slab without node cells (no span table) pushes zeros: the neighbour waits for this row whatever it holds
        const int xmin = a.sp_xmin ? a.sp_xmin[row] : -1;
        const bool in = xmin >= 0 && x >= xmin && x <= a.sp_xmax[row];
        const long ip = in ? a.sp_base[row] + (x - xmin) : 0;
        for(int p = 0; p < a.npoles; ++p) a.dst[p][i] = (in && a.pool[p]) ? a.pool[p][ip] : 0.0;
    }
    __threadfence_system();
    __syncthreads();
    if(threadIdx.x == 0)
    {
        const unsigned done = atomicAdd(a.counter, 1u) + 1u;
        if(done == gridDim.x)
        {
            *a.counter = 0u;
            __threadfence_system();
            *reinterpret_cast<volatile int*>(a.peer_flag) = a.value;
            __threadfence_system();
        }
    }
}

// blocks the stream until every listed flag has reached its value (or ~30 s have passed: then the error flag is raised so that
// a lost neighbour surfaces as an error instead of a hang)
struct HaloWaitArgs { const int* flag[4]; int value[4]; int n; int* error; };
__global__ void k_halo_wait(const HaloWaitArgs a)
{
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for(int f = 0; f < a.n; ++f)
    {
        const volatile int* p = a.flag[f];
        while(*p < a.value[f])
        {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if(t - t0 > 30000000000ull) { *a.error = 1; return; }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

} // namespace chiml
