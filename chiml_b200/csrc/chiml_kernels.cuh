// Device kernels of the B200 FDTD engine.  sm_100a only.
//
// Arithmetic contract: every update performs the same individually rounded multiplies and adds, in
// the same order, as the reference's BLAS level-1 call chains (one daxpy = one rounded product plus
// one rounded sum per element), so results are bit-identical to the restated oracle.  That is why
// the code below uses __dmul_rn / __dadd_rn instead of letting the compiler contract to FMA; the
// kernels are HBM-bound, so the extra FP64 issue slots are free (DESIGN.md "Arithmetic order").
#pragma once

#include "chiml_ctx.hpp"

namespace chiml {

struct PmlArgs
{
    const double* F;      // DbField per coordinate along `axis`
    const double* b;
    const double* c;
    const int32_t* cmap;  // coordinate -> compact psi coordinate
    double* psi;
    double Db;
    long psi_pitch;
    int axis;
    int nact;
    int present;
    int has_psi;
};

struct CompArgs
{
    const uint16_t* info;
    const ClassEntry* cls;
    const double2* pf;    // compact copy of {pf1, pf2} per class for the interior fast path
    double* U;            // E_c or H_c
    double* D;            // D_c (E components of dispersive runs) or nullptr
    const double* Vj;     // grid_j, used with offK  (UTIL/FDTD_up_eq.cpp:29-30)
    const double* Vk;     // grid_k, used with offJ  (UTIL/FDTD_up_eq.cpp:32-33)
    long offJ, offK;      // physical offsets of ind_j / ind_k relative to ind
    PmlArgs pml[2];
    // isotropic poles (compact row spans)
    const int32_t* sp_xmin;
    const int64_t* sp_base;
    const double* Pcur[MAX_POLES];
    double* Pnew[MAX_POLES];     // the buffer that held prevP receives the new P
    // oriented-dipole D->E
    const double* oP[MAX_POLES]; // node-centred pole state AFTER this step's node update
    const double* oPg[MAX_POLES];// dense ghost row ny+1 of that state (written by the slab above), or nullptr
    int nordip;
    int ord_dx, ord_dy, ord_dz;  // node offset r + e_c of orDipDtoU
    int ord_zvariant;            // orDipDtoUZ (2-D TM Ez)
    // chiral poles (pools over the same spans as the isotropic ones); the eight-point stencil on the other family's same component
    const double* chiCur[MAX_CHI];
    double* chiNew[MAX_CHI];
    const double* oppPrev;       // previous value of the other family's component (prevH_[c] for E_c, prevE_[c] for H_c)
    long chi_oi, chi_oj, chi_ok; // physical offsets of ind_i, ind_j, ind_k of the chiral list relative to ind
};

struct StepArgs
{
    CompArgs c[3];
    const double* fam[3];        // the three components of the OTHER family (H for the E half step, E for the H half step)
    // node span table (oriented dipoles)
    const int32_t* nsp_xmin;
    const int32_t* nsp_xmax;
    const int64_t* nsp_base;
    int lx, ly, lz;
    long px;
    int pml_on_D;
    // TMA descriptors (CUtensorMap, 128 bytes each, in global memory) of one-row boxes (64 x 1 x 1 doubles) of the nine field arrays
    // [0..8] and of the psi arrays [9 + 2*comp + part]; nullptr = none (2-D grids, CHIML_B200_NO_TMA): chiml_update.cuh tma_prefetch_row
    const unsigned char* tmaps;
};
constexpr int TMAP_BYTES = 128, TMAP_PSI0 = 9, TMAP_COUNT = 9 + 12;

constexpr unsigned REC_WIDE2 = 0x80000000u;   // TileRec::part flag: the record spans two z-adjacent tiles, half a warp per row, x origin pad4
constexpr int TILE_X = 64;   // cells per tile row (32 lanes x 2 cells)
constexpr int TILE_Z = 8;    // rows per tile (3-D); 2-D grids use 1

// one work item of k_fast / k_uniform / k_general (built at commit time)
struct TileRec
{
    int x0, z0, y, ny;           // ny: number of consecutive y planes the block marches over (k_fast); 1 elsewhere
    unsigned rect[3];            // per component: xlo | xhi<<8 | zlo<<16 | zhi<<24 (tile-local, hi exclusive); 0 = no cell of this component
    unsigned part;               // which record of its tile this is (a tile with more than two info values per component has several)
    unsigned info[3];            // per component: the one info value of its cells (k_uniform)
    unsigned np;                 // isotropic poles of the classes of rect (bits 0-7, 8-15, 16-23: components 0, 1, 2)
    double2 pf[3];               // per component: {pf1, pf2} of its class
    double inv_eps[3];           // per component: 1/eps of its class (pole-free D->E)
    double pad3;
    // second rectangle of a component (k_uniform only; rectB == 0 when the tile has one info value)
    unsigned rectB[3]; unsigned pad4;  // pad4: tile-local x of lane 0 in a REC_WIDE2 record
    unsigned infoB[3]; unsigned npB;   // npB: as np, for rectB
    double2 pfB[3];
    double inv_epsB[3];
    double pad6;
};

struct NodeArgs
{
    const uint16_t* info;
    const ClassEntry* cls;
    const double* E[3];
    long eoff[3];                // physical offsets of the second averaging point per component
    const int32_t* sp_xmin;
    const int32_t* sp_xmax;
    const int64_t* sp_base;
    const int32_t* rows;         // compact list of the grid rows (z + lz*y) that hold node cells
    const double* Pcur[3][MAX_POLES];
    double* Pnew[3][MAX_POLES];
    const double* dipg[3][MAX_POLES];   // position-dependent dipole grids over the node spans (REL_TO_NORM), or nullptr: the class's direction
    int lx, ly, lz;
    long px;
};

__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }
// y <- y + a*x with separately rounded product and sum (daxpy semantics)
__device__ __forceinline__ double axpy1(double y, double a, double x) { return __dadd_rn(y, __dmul_rn(a, x)); }

__device__ __forceinline__ double node_value(const StepArgs& a, const double* pool, const double* ghost, int x, int y, int z)
{
    if(ghost && y == a.ly - 1) return ghost[x + (long)a.lx * z];
    if(!pool) return 0.0;           // a pole grid of the whole grid that this slab holds no cell of
    const long row = z + (long)a.lz * y;
    const int xmin = a.nsp_xmin[row];
    if(xmin < 0 || x < xmin || x > a.nsp_xmax[row]) return 0.0;
    return pool[a.nsp_base[row] + (x - xmin)];
}

// the same for the two x-adjacent cells of a thread: one span lookup per row
__device__ __forceinline__ double2 node_pair(const StepArgs& a, const double* pool, const double* ghost, int x, int y, int z)
{
    if(ghost && y == a.ly - 1) return make_double2(ghost[x + (long)a.lx * z], ghost[x + 1 + (long)a.lx * z]);
    if(!pool) return make_double2(0.0, 0.0);
    const long row = z + (long)a.lz * y;
    const int xmin = a.nsp_xmin[row];
    double2 v = make_double2(0.0, 0.0);
    if(xmin < 0) return v;
    const int xmax = a.nsp_xmax[row];
    const double* p = pool + a.nsp_base[row] - xmin;
    if(x >= xmin && x <= xmax) v.x = p[x];
    if(x + 1 >= xmin && x + 1 <= xmax) v.y = p[x + 1];
    return v;
}

} // namespace chiml
#include "chiml_tiles.hpp"
#include "chiml_update.cuh"
#include "chiml_emitters.cuh"
#include "chiml_halo.cuh"
#include "chiml_persist.cuh"
namespace chiml {

// updatePolE, oriented-dipole poles at the integer nodes
// (FDTD_MANAGER/parallelFDTDField.hpp:1350-1354 -> UTIL/FDTD_up_eq.cpp:450-631)
// DIPG: the dipole vector of a node comes from the grids of setupDipMoments (parallelFDTDField.hpp:960-1048) instead of the object's class
template <bool DIPG>
__global__ void __launch_bounds__(256) k_ordip_poles(const __grid_constant__ NodeArgs a)
{
    // one block per 256-cell chunk of one row of the compact row list: only rows that hold node cells are visited.
    // (Two nodes per thread with 16-byte accesses was measured: 94 registers, 0.71 ms against 0.43 ms on the C5 slab -- the kernel
    // lives on many resident warps, and its DRAM traffic beyond the pole state is the E field it averages, 24 B per node.)
    const long row = a.rows[blockIdx.y];
    const int xmin = a.sp_xmin[row];
    const int x = xmin + blockIdx.x * blockDim.x + threadIdx.x;
    if(x > a.sp_xmax[row]) return;
    const long r = x + a.px * row;
    const uint16_t info = a.info[r];
    if(info == 0) return;
    const ClassEntry& ce = a.cls[info & CLS_MASK];
    const long ip = a.sp_base[row] + (x - xmin);
    double e0[3] = {0.0, 0.0, 0.0}, e1[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for(int c = 0; c < 3; ++c)
        if(a.E[c]) { e0[c] = a.E[c][r]; e1[c] = a.E[c][r + a.eoff[c]]; }
    const bool planar = a.E[0] != nullptr;     // 3-D or TE; otherwise the TM (Ez only) variant
    for(int p = 0; p < ce.npoles; ++p)
    {
        double pc[3] = {0.0, 0.0, 0.0}, dp3[3];
#pragma unroll
        for(int c = 0; c < 3; ++c)
        {
            dp3[c] = ce.dip[p][c];
            if(a.E[c])
            {
                pc[c] = dm(ce.alpha[p], a.Pcur[c][p][ip]);
                pc[c] = axpy1(pc[c], ce.xi[p], a.Pnew[c][p][ip]);
                if(DIPG && a.dipg[c][p]) dp3[c] = a.dipg[c][p][ip];
            }
        }
        if(planar)
        {
            double dotU = 0.0;
            bool first = true;
#pragma unroll
            for(int c = 0; c < 3; ++c)
            {
                if(!a.E[c]) continue;
                const double dp = dp3[c];
                const double t0 = __ddiv_rn(dm(dp, e0[c]), 2.0);   // multAvg: x*y/2.0 (UTIL/utilityFxns.hpp:38)
                const double t1 = __ddiv_rn(dm(dp, e1[c]), 2.0);
                if(first) { dotU = da(t0, t1); first = false; }
                else { dotU = da(dotU, t0); dotU = da(dotU, t1); }
            }
#pragma unroll
            for(int c = 0; c < 3; ++c)
                if(a.E[c]) pc[c] = axpy1(pc[c], ce.gamma[p], dm(dp3[c], dotU));
        }
        else
        {
            const double dotU = dm(dp3[2], e0[2]);
            pc[2] = axpy1(pc[2], ce.gamma[p], dotU);
        }
#pragma unroll
        for(int c = 0; c < 3; ++c)
            if(a.E[c]) a.Pnew[c][p][ip] = pc[c];
    }
}

// src->addPul (SOURCE/parallelSourceNormal.cpp:15-37): grid[box] += dt*Re(sum pulse(t)); the product is formed on the host
__global__ void k_source(double* field, int lx0, int lz0, int ly0, int sx, int sz, int sy, int lz, long px, const double* amp, const uint16_t* skipInfo = nullptr)
{
    const long n = (long)sx * sz * sy;
    const double av = *amp;
    for(long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    {
        const int ix = (int)(i % sx);
        const int iz = (int)((i / sx) % sz);
        const int iy = (int)(i / ((long)sx * sz));
        const long r = (lx0 + ix) + px * ((lz0 + iz) + (long)lz * (ly0 + iy));
        if(skipInfo && (skipInfo[r] & F_D2E)) continue;       // H = (B - sum M) / mu_inf has already replaced what the reference would add to
        field[r] = da(field[r], av);
    }
}

// TFSF surface corrections (tfsfUpdateFxnReal::addIncdFields / addIncdFieldsEPChange, SOURCE/parallelTFSF.cpp:77-105): one launch applies one
// "wave" of surface lists -- at most one per target array, so no two entries touch the same cell -- blockIdx.y = entry.  For pair l and
// element i:  target[main_l + i * stride_main] += prefactor * incd[incd_l + ix0 + i * stride_incd] (/ ep_mu[same index]); a negative
// incident stride starts at the far end like the BLAS call it replaces (ix0 = (1 - n) * stride).  Main-grid indices are the reference's
// logical ones (x + lx * row) and are mapped to the padded rows here.
constexpr int TFSF_WAVE = 12;
struct TfsfEntry { double* target; const int32_t* pairs; const double* ep_mu; int npairs, n, stride_incd, stride_main, incd_offset; double prefactor; };
struct TfsfArgs { TfsfEntry e[TFSF_WAVE]; int n; const double* incd; int lx; long px; };
__global__ void k_tfsf(TfsfArgs a)
{
    if((int)blockIdx.y >= a.n) return;
    const TfsfEntry& t = a.e[blockIdx.y];
    const double* incd = a.incd + t.incd_offset;
    const long total = (long)t.npairs * t.n;
    const long ix0 = t.stride_incd < 0 ? (long)(1 - t.n) * t.stride_incd : 0;
    for(long q = blockIdx.x * (long)blockDim.x + threadIdx.x; q < total; q += (long)gridDim.x * blockDim.x)
    {
        const long l = q / t.n, i = q % t.n;
        const long ii = t.pairs[2 * l] + ix0 + i * t.stride_incd;
        const long m = t.pairs[2 * l + 1] + i * (long)t.stride_main;
        double v = incd[ii];
        if(t.ep_mu) v = v / t.ep_mu[ii];
        const long r = (m % a.lx) + a.px * (m / a.lx);
        t.target[r] = da(t.target[r], dm(t.prefactor, v));
    }
}
// commit-time check: a surface cell inside the CPML (any of the four CPML flags of its component) cannot be reproduced, because the
// reference adds the incident field between the curl and the CPML terms, which are one pass here
__global__ void k_tfsf_check(const int32_t* pairs, int npairs, int n, int stride_main, const uint16_t* info, int lx, long px, long ncell, int* err)
{
    const long total = (long)npairs * n;
    for(long q = blockIdx.x * (long)blockDim.x + threadIdx.x; q < total; q += (long)gridDim.x * blockDim.x)
    {
        const long m = pairs[2 * (q / n) + 1] + (q % n) * (long)stride_main;
        if(m < 0 || m >= ncell) { atomicOr(err, 4); continue; }
        if(info[(m % lx) + px * (m / lx)] & (F_PS0 | F_PS1 | F_PG0 | F_PG1)) atomicOr(err, 8);
    }
}

// copy2PrevFields_ (FDTD_MANAGER/parallelFDTDField.hpp:1411-1416, 1441-1446): rows {length, x, y, z} of the three components of one family into
// their prev grids, after the chiral update that read them
struct PrevCopyArgs { const double* src[3]; double* dst[3]; const int4* rows; unsigned nrows; int lz; long px; };
__global__ void k_prev_copy(PrevCopyArgs a)
{
    const unsigned q = blockIdx.x;
    if(q >= a.nrows) return;
    const int4 b = a.rows[q];
    const long off = b.y + a.px * (b.w + (long)a.lz * b.z);
    for(int c = 0; c < 3; ++c)
        if(a.src[c] && a.dst[c])
            for(int i = threadIdx.x; i < b.x; i += blockDim.x) a.dst[c][off + i] = a.src[c][off + i];
}

// Periodic wrap copies of up to three components in one launch (applyBC1Proc, UTIL/FDTD_up_eq.cpp:1058-1116; blockIdx.y = component).
// The reference's sequence of dcopy_ calls amounts to: every ghost cell of the box [0, xmax] x [0, ymax] x [zmin-1, zmax] takes the
// value of its periodic image inside the box (x = 0 <- xmax-1, x = xmax <- 1, likewise y and z); every source is an inner cell, so
// the copies are independent.  3-D: the shell is walked as y faces (whole planes), z faces (rows 1 .. ymax-1), x faces (the rest).
// 2-D (zmin = 0): rows 0 and ymax over x in [1, xmax-1], columns 0 and xmax over y in [1, ymax] (the cells (0, 0), (xmax, 0) stay as they
// are, as in the reference).
// A slab of several (ymax < 0; applyBCProcMid, UTIL/FDTD_up_eq.cpp:1036-1061): the x / z ghost cells of the owned rows 1 .. ly - 2 only, each from
// its image in the same row; the y direction is the ring of ghost-row pushes between the slabs.
struct WrapArgs { double* f[3]; ChimlWrap w[3]; int n; int lz; long px; int ly; };
__global__ void k_wrap(WrapArgs a)
{
    const int c = blockIdx.y;
    if(c >= a.n) return;
    double* F = a.f[c];
    const ChimlWrap w = a.w[c];
    const long px = a.px, lz = a.lz;
    const long i0 = blockIdx.x * (long)blockDim.x + threadIdx.x, istep = (long)gridDim.x * blockDim.x;
    if(w.ymax < 0)
    {
        const long R = a.ly - 2;                                   // owned rows
        if(w.zmin != 0)
        {
            const long X = w.xmax + 1, Z = w.zmax - w.zmin + 2;
            const long nB = 2 * X * R, nC = 2 * R * (Z - 2);
            for(long i = i0; i < nB + nC; i += istep)
            {
                int x, y, z;
                if(i < nB) { x = (int)(i % X); const long r = i / X; y = (int)(r % R) + 1; z = (r / R) ? w.zmax : w.zmin - 1; }
                else       { const long j = i - nB; z = (int)(j % (Z - 2)) + w.zmin; const long r = j / (Z - 2); y = (int)(r % R) + 1; x = (r / R) ? w.xmax : 0; }
                const int sx = x == 0 ? w.xmax - 1 : (x == w.xmax ? 1 : x);
                const int sz = z == w.zmin - 1 ? w.zmax - 1 : (z == w.zmax ? w.zmin : z);
                F[x + px * (z + lz * y)] = F[sx + px * (sz + lz * y)];
            }
        }
        else
            for(long i = i0; i < 2 * R; i += istep)
            {
                const int y = (int)(i % R) + 1, x = (i / R) ? w.xmax : 0;
                F[x + px * (long)y] = F[(x == 0 ? w.xmax - 1 : 1) + px * (long)y];
            }
        return;
    }
    if(w.zmin != 0)
    {
        const long X = w.xmax + 1, Y = w.ymax + 1, Z = w.zmax - w.zmin + 2;
        const long nA = 2 * X * Z, nB = 2 * X * (Y - 2), nC = 2 * (Y - 2) * (Z - 2);
        for(long i = i0; i < nA + nB + nC; i += istep)
        {
            int x, y, z;
            if(i < nA)           { x = (int)(i % X); const long r = i / X; z = (int)(r % Z) + w.zmin - 1; y = (r / Z) ? w.ymax : 0; }
            else if(i < nA + nB) { const long j = i - nA; x = (int)(j % X); const long r = j / X; y = (int)(r % (Y - 2)) + 1; z = (r / (Y - 2)) ? w.zmax : w.zmin - 1; }
            else                 { const long j = i - nA - nB; z = (int)(j % (Z - 2)) + w.zmin; const long r = j / (Z - 2); y = (int)(r % (Y - 2)) + 1; x = (r / (Y - 2)) ? w.xmax : 0; }
            const int sx = x == 0 ? w.xmax - 1 : (x == w.xmax ? 1 : x);
            const int sy = y == 0 ? w.ymax - 1 : (y == w.ymax ? 1 : y);
            const int sz = z == w.zmin - 1 ? w.zmax - 1 : (z == w.zmax ? w.zmin : z);
            F[x + px * (z + lz * y)] = F[sx + px * (sz + lz * sy)];
        }
    }
    else
    {
        const long nR = 2L * (w.xmax - 1), nC = 2L * w.ymax;
        for(long i = i0; i < nR + nC; i += istep)
        {
            int x, y;
            if(i < nR) { x = (int)(i % (w.xmax - 1)) + 1; y = (i / (w.xmax - 1)) ? w.ymax : 0; }
            else       { const long j = i - nR; y = (int)(j % w.ymax) + 1; x = (j / w.ymax) ? w.xmax : 0; }
            const int sx = x == 0 ? w.xmax - 1 : (x == w.xmax ? 1 : x);
            const int sy = y == 0 ? w.ymax - 1 : (y == w.ymax ? 1 : y);
            F[x + px * (long)y] = F[sx + px * (long)sy];
        }
    }
}

// Bloch-periodic wrap copies of complex fields held as a real and an imaginary array (applyBC1Proc, complex fields, UTIL/FDTD_up_eq.cpp:1248-1324):
// the same ghost shell as k_wrap, every ghost cell = phase * image with phase = exp(i (+-kx dx xmax +- ky dy ymax +- kz dz zmax)) over the axes
// that wrapped, ONE exponential of the summed argument as the reference evaluates it (table ph, index (sx+1) + 3 (sy+1) + 9 (sz+1), computed on
// the host), product in netlib order.  The reference's corner assignments read row ymax AFTER it received the phased image of row 1: a corner
// is phase_corner * (phase_y+ * F(x', 1, z')), for the corners at y = 0 too.  Likewise the 2-D columns run over rows 1 .. ymax, so their cell in
// row ymax is phase_x * (phase_y+ * F(x', 1)).
struct BlochArgs { double* fr[3]; double* fi[3]; ChimlWrap w[3]; double2 ph[3][27]; int n; int lz; long px; };
__device__ __forceinline__ double2 cmul_rn(const double2 a, const double2 b)
{ return make_double2(da(dm(a.x, b.x), -dm(a.y, b.y)), da(dm(a.x, b.y), dm(a.y, b.x))); }
__global__ void k_wrap_bloch(const __grid_constant__ BlochArgs a)
{
    const int c = blockIdx.y;
    if(c >= a.n) return;
    double* R = a.fr[c]; double* I = a.fi[c];
    const ChimlWrap w = a.w[c];
    const long px = a.px, lz = a.lz;
    const long i0 = blockIdx.x * (long)blockDim.x + threadIdx.x, istep = (long)gridDim.x * blockDim.x;
    const double2 phyPlus = a.ph[c][1 + 3 * 2 + 9 * 1];
    if(w.zmin != 0)
    {
        const long X = w.xmax + 1, Y = w.ymax + 1, Z = w.zmax - w.zmin + 2;
        const long nA = 2 * X * Z, nB = 2 * X * (Y - 2), nC = 2 * (Y - 2) * (Z - 2);
        for(long i = i0; i < nA + nB + nC; i += istep)
        {
            int x, y, z;
            if(i < nA)           { x = (int)(i % X); const long r = i / X; z = (int)(r % Z) + w.zmin - 1; y = (r / Z) ? w.ymax : 0; }
            else if(i < nA + nB) { const long j = i - nA; x = (int)(j % X); const long r = j / X; y = (int)(r % (Y - 2)) + 1; z = (r / (Y - 2)) ? w.zmax : w.zmin - 1; }
            else                 { const long j = i - nA - nB; z = (int)(j % (Z - 2)) + w.zmin; const long r = j / (Z - 2); y = (int)(r % (Y - 2)) + 1; x = (r / (Y - 2)) ? w.xmax : 0; }
            const int cx = x == 0 ? -1 : (x == w.xmax ? 1 : 0), cy = y == 0 ? -1 : (y == w.ymax ? 1 : 0), cz = z == w.zmin - 1 ? -1 : (z == w.zmax ? 1 : 0);
            const int sx = cx < 0 ? w.xmax - 1 : (cx > 0 ? 1 : x);
            const int sz = cz < 0 ? w.zmax - 1 : (cz > 0 ? w.zmin : z);
            const bool corner = cx != 0 && cy != 0 && cz != 0;
            const int sy = corner ? 1 : (cy < 0 ? w.ymax - 1 : (cy > 0 ? 1 : y));
            const long s = sx + px * (sz + lz * sy);
            double2 v = make_double2(R[s], I[s]);
            if(corner) v = cmul_rn(phyPlus, v);
            v = cmul_rn(a.ph[c][(cx + 1) + 3 * (cy + 1) + 9 * (cz + 1)], v);
            const long d = x + px * (z + lz * y);
            R[d] = v.x; I[d] = v.y;
        }
    }
    else
    {
        const long nR = 2L * (w.xmax - 1), nC = 2L * w.ymax;
        for(long i = i0; i < nR + nC; i += istep)
        {
            int x, y;
            if(i < nR) { x = (int)(i % (w.xmax - 1)) + 1; y = (i / (w.xmax - 1)) ? w.ymax : 0; }
            else       { const long j = i - nR; y = (int)(j % w.ymax) + 1; x = (j / w.ymax) ? w.xmax : 0; }
            const int cx = x == 0 ? -1 : (x == w.xmax ? 1 : 0), cy = y == 0 ? -1 : (y == w.ymax ? 1 : 0);
            const int sx = cx < 0 ? w.xmax - 1 : (cx > 0 ? 1 : x);
            const int sy = cy < 0 ? w.ymax - 1 : (cy > 0 ? 1 : y);
            const long s = sx + px * (long)sy;
            double2 v = make_double2(R[s], I[s]);
            if(cx != 0 && cy != 0) { v = cmul_rn(phyPlus, v); v = cmul_rn(a.ph[c][(cx + 1) + 3 * 1 + 9 * 1], v); }       // column cell of row ymax: y first, then x
            else v = cmul_rn(a.ph[c][(cx + 1) + 3 * (cy + 1) + 9 * 1], v);
            const long d = x + px * (long)y;
            R[d] = v.x; I[d] = v.y;
        }
    }
}

// detector sampling (DTC/parallelStorageDTC.cpp:17-44): copy the box into the ring, x fastest, then z, then y
__global__ void k_detector(const double* field, int lx0, int lz0, int ly0, int sx, int sz, int sy, int lz, long px, double* out)
{
    const long n = (long)sx * sz * sy;
    for(long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    {
        const int ix = (int)(i % sx);
        const int iz = (int)((i / sx) % sz);
        const int iy = (int)(i / ((long)sx * sz));
        out[i] = field[(lx0 + ix) + px * ((lz0 + iz) + (long)lz * (ly0 + iy))];
    }
}

// running DFT (DTC/parallelStorageFreqDTC.cpp:21-30): acc[out + f + nfreq*i] += tw[f] * field[ind + i*stride], one rounded product and
// one rounded sum per entry like the dger_ it replaces; f fastest so that the accumulator traffic is coalesced.  Here:
// all running-DFT sets that are due after a step in one launch (blockIdx.y = set): a flux box is 8 (2-D) to 24 (3-D) stored fields, each a
// few thousand entries -- one launch per set cost 0.098 ms per step on C2 (4 edges, 64 frequencies), more than the field update itself
constexpr int DFT_BATCH = 24;
struct DftBatchSet { const double* field; const ChimlDftLine* lines; unsigned long long nlines; int npts, stride, nfreq; const double* tw; double* re; double* im; };
struct DftBatchArgs { DftBatchSet s[DFT_BATCH]; int n; int lx; long px; };
__global__ void k_dft_batch(const __grid_constant__ DftBatchArgs a)
{
    if((int)blockIdx.y >= a.n) return;
    const DftBatchSet& d = a.s[blockIdx.y];
    const size_t n = (size_t)d.nlines * (size_t)d.npts * (size_t)d.nfreq;
    for(size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
    {
        const int f = (int)(e % d.nfreq);
        const int i = (int)((e / d.nfreq) % d.npts);
        const size_t l = e / ((size_t)d.nfreq * d.npts);
        const long lg = (long)d.lines[l].ind + (long)i * d.stride;
        const double u = d.field[(lg % a.lx) + a.px * (lg / a.lx)];
        const size_t o = (size_t)d.lines[l].out + f + (size_t)d.nfreq * i;
        d.re[o] = da(d.re[o], dm(d.tw[2 * f], u));
        d.im[o] = da(d.im[o], dm(d.tw[2 * f + 1], u));
    }
}

// ---- setup-time painting of the cell-info planes from the reference's lists ------------------------
// one warp per run
__global__ void k_paint_runs(const ChimlRun* runs, const uint8_t* cls, size_t nruns, uint16_t flags, uint16_t* info, int lx, long px, int* err)
{
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x & 31;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) / 32;
    for(size_t e = warp; e < nruns; e += nwarps)
    {
        const ChimlRun rr = runs[e];
        const long row = rr.ind / lx;
        const int x0 = rr.ind % lx;
        const long base = x0 + px * row;
        const uint16_t cv = cls[e];
        for(int i = lane; i < rr.n; i += 32)
        {
            uint16_t v = info[base + i];
            if((v & CLS_MASK) != 0 && cv != 0 && (v & CLS_MASK) != cv) atomicExch(err, 1);
            if(v & flags) atomicExch(err, 2);   // the same operation listed twice for one cell
            info[base + i] = (uint16_t)(v | flags | cv);
        }
    }
}

// one warp per strided line (CPML lists)
__global__ void k_paint_lines(const int4* lines /* n, stride, ind, unused */, size_t nlines, uint16_t flags, uint16_t* info, int lx, long px, int* err)
{
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x & 31;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) / 32;
    for(size_t e = warp; e < nlines; e += nwarps)
    {
        const int4 ln = lines[e];
        for(int i = lane; i < ln.x; i += 32)
        {
            const long l = (long)ln.z + (long)i * ln.y;
            const long r = (l % lx) + px * (l / lx);
            const uint16_t v = info[r];
            if(v & flags) atomicExch(err, 3);
            info[r] = (uint16_t)(v | flags);
        }
    }
}

} // namespace chiml
