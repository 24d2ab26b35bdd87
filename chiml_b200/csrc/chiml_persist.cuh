// k_steps_2d: ALL time steps of one chiml_gpu_step_n call in ONE cooperative launch, for 2-D grids.
//
// A 2-D grid of BASELINE's sizes is a few MB to a few hundred MB of state: a 512 x 512 step moves 19 MB, which the kernels of the
// launch-per-phase path finish in ~3 us each -- the step then costs seven launch latencies (0.066 ms), 5 % of what the memory system
// can do.  Here the grid of thread blocks is resident for the whole call (one block per SM slot, cooperative launch) and walks the
// reference's step order (FDTD_MANAGER/parallelFDTDField.hpp:1228-1303) with grid-wide barriers where the order carries a dependency:
//
//     H half step   | barrier |  sources (E and H)  | barrier |  E half step  | barrier | next step ...
//     (+ detector / running-DFT samples of E fields    (+ samples of H fields: the E half
//        taken after the previous step: the H half        step only reads H)
//        step only reads E)
//
// (periodic boundaries add the wrap copies of H after the sources and of E after the E half step, each behind its own barrier)
// Three barriers per step instead of seven launches; small grids stay in L2 between the phases.  The work items are the tile records
// of the launch-per-phase kernels (same device functions, same arithmetic, same results): a warp takes one record (k_fast), one
// (record, component) pair (k_uniform_rows / k_general) at a time.
// No emitters, no oriented-dipole media (the host setup refuses them in 2-D), single slab: everything else takes the launch path.
#pragma once

#include <cooperative_groups.h>

namespace chiml {

constexpr int P2D_MAX_DET = 8, P2D_MAX_DFT = 24;

struct P2DDetector { int field; int loc[3], sz[3]; int every; double* ring; unsigned long long cap, count0; unsigned long long sample_len; };
struct P2DDft { int field, group, every, nfreq, npts, stride; const ChimlDftLine* lines; unsigned long long nlines; double* re; double* im; unsigned long long tw_off; };

struct Persist2DArgs
{
    const StepArgs* sa;            // device: [0] H half step, [1] / [2] E half step with pole buffer 0 / 1 holding the current P
    int pcur0;                     // which of the two the first step of this call uses
    const TileRec* tiles[2][3];    // [0] E, [1] H; fast, uniform, general
    unsigned ntiles[2][3];
    int nsteps, nsrc, ndet, ndft;
    long long step0;               // steps taken before this call
    SourceDev src[MAX_SOURCES];
    const double* src_amp;         // [nsteps][nsrc]
    double* field[CHIML_NFIELDS];
    P2DDetector det[P2D_MAX_DET];
    P2DDft dft[P2D_MAX_DFT];
    const double* tw; unsigned long long tw_per_step;
    int lx, lz; long px;
    int periodic;                  // wrap copies (chiml_gpu_set_periodic) of the components that exist
    ChimlWrap wrap[6]; int has_wrap[6];
};

// applyBC1Proc on a 2-D grid (k_wrap, chiml_kernels.cuh): rows 0 and ymax over x in [1, xmax-1], columns 0 and xmax over y in [1, ymax],
// each ghost cell from its periodic image inside
__device__ __forceinline__ void p2d_wraps(const Persist2DArgs& p, const bool isE, const unsigned long long gt, const unsigned long long nt)
{
    for(int i = 0; i < 3; ++i)
    {
        const int comp = (isE ? 0 : 3) + i;
        if(!p.has_wrap[comp]) continue;
        const ChimlWrap w = p.wrap[comp];
        double* F = p.field[comp];
        const unsigned long long nR = 2ull * (w.xmax - 1), nC = 2ull * w.ymax;
        for(unsigned long long e = gt; e < nR + nC; e += nt)
        {
            int x, y;
            if(e < nR) { x = (int)(e % (w.xmax - 1)) + 1; y = (e / (w.xmax - 1)) ? w.ymax : 0; }
            else       { const unsigned long long j = e - nR; y = (int)(j % w.ymax) + 1; x = (j / w.ymax) ? w.xmax : 0; }
            const int sx = x == 0 ? w.xmax - 1 : (x == w.xmax ? 1 : x);
            const int sy = y == 0 ? w.ymax - 1 : (y == w.ymax ? 1 : y);
            F[x + p.px * (long)y] = F[sx + p.px * (long)sy];
        }
    }
}

template <bool IS_E, int MODE>
__device__ __forceinline__ void p2d_family(const StepArgs& a, const Persist2DArgs& p, const unsigned gw, const unsigned nw)
{
    const int fam = IS_E ? 0 : 1;
    const int xl = 2 * threadIdx.x;
    for(unsigned u = gw; u < p.ntiles[fam][0]; u += nw) fast_tile<IS_E, MODE>(a, p.tiles[fam][0][u], xl, 0);
    for(unsigned u = gw; u < 3u * p.ntiles[fam][1]; u += nw) uniform_body<IS_E, MODE>(a, p.tiles[fam][1][u / 3], xl, 0, (int)(u % 3));
    for(unsigned u = gw; u < 3u * p.ntiles[fam][2]; u += nw)
    {
        const TileRec& t = p.tiles[fam][2][u / 3];
        const int x = t.x0 + xl, z = t.z0, comp = (int)(u % 3);
        if(x >= a.px || z >= a.lz) continue;
        if(comp == 0)      general_comp<IS_E, MODE, 0>(a, t, x, z);
        else if(comp == 1) general_comp<IS_E, MODE, 1>(a, t, x, z);
        else               general_comp<IS_E, MODE, 2>(a, t, x, z);
    }
}

// is any detector / running-DFT sample of the fields of one family due after step number `count`?
__device__ __forceinline__ bool p2d_samples_due(const Persist2DArgs& p, const long long count, const bool wantH)
{
    for(int d = 0; d < p.ndet; ++d)
        if((p.det[d].field >= CHIML_HX && p.det[d].field <= CHIML_HZ) == wantH && count % p.det[d].every == 0) return true;
    for(int q = 0; q < p.ndft; ++q)
        if((p.dft[q].field >= CHIML_HX && p.dft[q].field <= CHIML_HZ) == wantH && count % p.dft[q].every == 0 && p.dft[q].nlines) return true;
    return false;
}

// detector / running-DFT samples due after step number `count` (counted from 1 over the whole run) of the fields of one family,
// taken by threads gt, gt + nt, ... of nt
__device__ __forceinline__ void p2d_samples(const Persist2DArgs& p, const long long count, const int k, const bool wantH,
                                            const unsigned long long gt, const unsigned long long nt)
{
    for(int d = 0; d < p.ndet; ++d)
    {
        const P2DDetector& dt = p.det[d];
        const bool isH = dt.field >= CHIML_HX && dt.field <= CHIML_HZ;
        if(isH != wantH || count % dt.every != 0) continue;
        const unsigned long long s = dt.count0 + (unsigned long long)(count / dt.every - p.step0 / dt.every) - 1ull;   // sample number
        double* out = dt.ring + (s % dt.cap) * dt.sample_len;
        const double* f = p.field[dt.field];
        for(unsigned long long i = gt; i < dt.sample_len; i += nt)
        {
            const int ix = (int)(i % dt.sz[0]), iz = (int)((i / dt.sz[0]) % dt.sz[2]), iy = (int)(i / ((unsigned long long)dt.sz[0] * dt.sz[2]));
            out[i] = f[(dt.loc[0] + ix) + p.px * ((dt.loc[2] + iz) + (long)p.lz * (dt.loc[1] + iy))];
        }
    }
    for(int q = 0; q < p.ndft; ++q)
    {
        const P2DDft& d = p.dft[q];
        const bool isH = d.field >= CHIML_HX && d.field <= CHIML_HZ;
        if(isH != wantH || count % d.every != 0 || d.nlines == 0) continue;
        const double* tw = p.tw + (unsigned long long)k * p.tw_per_step + d.tw_off;
        const double* f = p.field[d.field];
        const unsigned long long n = d.nlines * (unsigned long long)d.npts * (unsigned long long)d.nfreq;
        for(unsigned long long e = gt; e < n; e += nt)
        {
            // k_dft_batch (chiml_kernels.cuh): acc[out + f + nfreq*i] += tw[f] * field[ind + i*stride]
            const int fq = (int)(e % d.nfreq), i = (int)((e / d.nfreq) % d.npts);
            const unsigned long long l = e / ((unsigned long long)d.nfreq * d.npts);
            const long lg = (long)d.lines[l].ind + (long)i * d.stride;
            const double u = f[(lg % p.lx) + p.px * (lg / p.lx)];
            const unsigned long long o = (unsigned long long)d.lines[l].out + fq + (unsigned long long)d.nfreq * i;
            d.re[o] = da(d.re[o], dm(tw[2 * fq], u));
            d.im[o] = da(d.im[o], dm(tw[2 * fq + 1], u));
        }
    }
}

// Grid-wide barrier between two phases.  Every thread first waits for its OWN outstanding memory operations (__threadfence): the tile
// bodies load whole tile rows, and a lane whose cells belong to another record never uses what it loaded -- such a load can still be in
// flight when its warp reaches the barrier, and could then install its (by then stale) line in L1 AFTER the barrier's L1 invalidation;
// the next phase would read old field values from it.
__device__ __forceinline__ void p2d_barrier(cooperative_groups::grid_group& grid)
{
    __threadfence();
    grid.sync();
}

template <int MODE>
__global__ void __launch_bounds__(256) k_steps_2d(const __grid_constant__ Persist2DArgs p)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ StepArgs sh[2];            // [0] H half step, [1] E half step of the current step
    const unsigned tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    const unsigned gw = blockIdx.x * blockDim.y + threadIdx.y, nw = gridDim.x * blockDim.y;
    constexpr unsigned WORDS = sizeof(StepArgs) / sizeof(unsigned long long);
    static_assert(sizeof(StepArgs) % sizeof(unsigned long long) == 0, "StepArgs is copied in 8-byte words");
    for(unsigned i = tid; i < WORDS; i += nthr)
        reinterpret_cast<unsigned long long*>(&sh[0])[i] = reinterpret_cast<const unsigned long long*>(&p.sa[0])[i];
    for(int k = 0; k < p.nsteps; ++k)
    {
        __syncthreads();
        for(unsigned i = tid; i < WORDS; i += nthr)
            reinterpret_cast<unsigned long long*>(&sh[1])[i] = reinterpret_cast<const unsigned long long*>(&p.sa[1 + ((p.pcur0 + k) & 1)])[i];
        __syncthreads();
        const long long count = p.step0 + k + 1;
        // ---- H half step: updateH + updateHPML_ (step() items 4, 6); E-field samples of the step before.  When samples are due the
        // last warp of every block takes them while the other seven work on the tiles, otherwise all eight take tiles
        const unsigned lastw = blockDim.y - 1, slane = blockIdx.x * 32u + threadIdx.x, nslane = gridDim.x * 32u;
        {
            const bool due = k > 0 && p2d_samples_due(p, count - 1, false);
            if(!due) p2d_family<false, MODE>(sh[0], p, gw, nw);
            else if(threadIdx.y == lastw) p2d_samples(p, count - 1, k - 1, false, slane, nslane);
            else p2d_family<false, MODE>(sh[0], p, blockIdx.x * lastw + threadIdx.y, gridDim.x * lastw);
        }
        p2d_barrier(grid);
        // ---- sources (item 7): field[box] += dt * Re sum pulse(t), E and H alike
        {
            const unsigned long long gt = (unsigned long long)blockIdx.x * nthr + tid, nt = (unsigned long long)gridDim.x * nthr;
            for(int q = 0; q < p.nsrc; ++q)
            {
                const SourceDev& s = p.src[q];
                const double av = p.src_amp[(size_t)k * p.nsrc + q];
                double* f = p.field[s.field];
                const unsigned long long n = (unsigned long long)s.sz[0] * s.sz[1] * s.sz[2];
                for(unsigned long long i = gt; i < n; i += nt)
                {
                    const int ix = (int)(i % s.sz[0]), iz = (int)((i / s.sz[0]) % s.sz[2]), iy = (int)(i / ((unsigned long long)s.sz[0] * s.sz[2]));
                    const long r = (s.loc[0] + ix) + p.px * ((s.loc[2] + iz) + (long)p.lz * (s.loc[1] + iy));
                    f[r] = da(f[r], av);
                }
            }
        }
        p2d_barrier(grid);
        // ---- periodic boundaries of H (item 9)
        if(p.periodic)
        {
            p2d_wraps(p, false, (unsigned long long)blockIdx.x * nthr + tid, (unsigned long long)gridDim.x * nthr);
            p2d_barrier(grid);
        }
        // ---- E half step: isotropic poles, updateD / updateE, updateEPML_, D2E (items 10-15); H-field samples of this step
        {
            const bool due = p2d_samples_due(p, count, true);
            if(!due) p2d_family<true, MODE>(sh[1], p, gw, nw);
            else if(threadIdx.y == lastw) p2d_samples(p, count, k, true, slane, nslane);
            else p2d_family<true, MODE>(sh[1], p, blockIdx.x * lastw + threadIdx.y, gridDim.x * lastw);
        }
        p2d_barrier(grid);
        // ---- periodic boundaries of E (item 17)
        if(p.periodic)
        {
            p2d_wraps(p, true, (unsigned long long)blockIdx.x * nthr + tid, (unsigned long long)gridDim.x * nthr);
            p2d_barrier(grid);
        }
    }
    if(p.nsteps > 0) p2d_samples(p, p.step0 + p.nsteps, p.nsteps - 1, false, (unsigned long long)blockIdx.x * nthr + tid, (unsigned long long)gridDim.x * nthr);
}

} // namespace chiml
