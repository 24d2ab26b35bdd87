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
    const double* V;      // field driving this part (grid_k for part 0, grid_j for part 1)
    const double* F;      // DbField per coordinate along `axis`
    const double* b;
    const double* c;
    const int32_t* cmap;  // coordinate -> compact psi coordinate
    double* psi;
    double Db;
    long off;             // physical offset of the second stencil point
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
    int nordip;
    int ord_dx, ord_dy, ord_dz;  // node offset r + e_c of orDipDtoU
    int ord_zvariant;            // orDipDtoUZ (2-D TM Ez)
};

struct StepArgs
{
    CompArgs c[3];
    // node span table (oriented dipoles)
    const int32_t* nsp_xmin;
    const int32_t* nsp_xmax;
    const int64_t* nsp_base;
    int lx, ly, lz;
    long px;
    int pml_on_D;
};

struct NodeArgs
{
    const uint16_t* info;
    const ClassEntry* cls;
    const double* E[3];
    long eoff[3];                // physical offsets of the second averaging point per component
    const int32_t* sp_xmin;
    const int64_t* sp_base;
    const double* Pcur[3][MAX_POLES];
    double* Pnew[3][MAX_POLES];
    int lx, ly, lz;
    long px;
};

__device__ __forceinline__ double dm(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double da(double a, double b) { return __dadd_rn(a, b); }
// y <- y + a*x with separately rounded product and sum (daxpy semantics)
__device__ __forceinline__ double axpy1(double y, double a, double x) { return __dadd_rn(y, __dmul_rn(a, x)); }

__device__ __forceinline__ double node_value(const StepArgs& a, const double* pool, int x, int y, int z)
{
    const long row = z + (long)a.lz * y;
    const int xmin = a.nsp_xmin[row];
    if(xmin < 0 || x < xmin || x > a.nsp_xmax[row]) return 0.0;
    return pool[a.nsp_base[row] + (x - xmin)];
}

template <bool IS_E>
__device__ __forceinline__ void update_component(const StepArgs& a, const CompArgs& ca, long r, long row, int x, int y, int z)
{
    const uint16_t info = ca.info[r];
    if(info == 0) return;
    const ClassEntry& ce = ca.cls[info & CLS_MASK];

    double u = ca.U[r];
    double pn[MAX_POLES];
    int np = 0;

    // updatePolE, isotropic poles (FDTD_MANAGER/parallelFDTDField.hpp:1355-1361 -> UTIL/FDTD_up_eq.cpp:435-446):
    // tmp = P; P = alpha*P; P += xi*Pprev; P += gamma*E^n; Pprev = tmp
    if(IS_E && (info & F_D2E))
    {
        np = ce.npoles;
        if(np > 0)
        {
            const long ip = ca.sp_base[row] + (x - ca.sp_xmin[row]);
#pragma unroll
            for(int p = 0; p < MAX_POLES; ++p)
            {
                if(p < np)
                {
                    double t = dm(ce.alpha[p], ca.Pcur[p][ip]);
                    t = axpy1(t, ce.xi[p], ca.Pnew[p][ip]);
                    t = axpy1(t, ce.gamma[p], u);
                    ca.Pnew[p][ip] = t;
                    pn[p] = t;
                }
            }
        }
    }

    const bool pmlCell = (info & (F_PG0 | F_PS0 | F_PG1 | F_PS1)) != 0;
    const bool pmlOnD = IS_E && a.pml_on_D;
    const bool needD = IS_E && ((info & (F_ISD | F_D2E | F_ORD2E)) || (pmlOnD && pmlCell));
    double dv = needD ? ca.D[r] : 0.0;
    bool dDirty = false;

    // updateD / updateE / updateH: TwoCompCurl, OneCompCurlJ, OneCompCurlK (UTIL/FDTD_up_eq.cpp:10-35)
    if(info & F_CURL)
    {
        double t = (IS_E && (info & F_ISD)) ? dv : u;
        if(ca.Vj)
        {
            t = axpy1(t,  ce.pf2, ca.Vj[r]);
            t = axpy1(t, -ce.pf2, ca.Vj[r + ca.offK]);
        }
        if(ca.Vk)
        {
            t = axpy1(t, -ce.pf1, ca.Vk[r]);
            t = axpy1(t,  ce.pf1, ca.Vk[r + ca.offJ]);
        }
        if(IS_E && (info & F_ISD)) { dv = t; dDirty = true; } else u = t;
    }

    // parallelCPML<T>::updateGrid (PML/parallelPML.hpp:693-697): part 0 then part 1; each part is
    // updatePsiField then the grid daxpys (PML/parallelPML.cpp:12-40)
    if(pmlCell)
    {
        double t = pmlOnD ? dv : u;
#pragma unroll
        for(int part = 0; part < 2; ++part)
        {
            const PmlArgs& pp = ca.pml[part];
            const uint16_t fg = part == 0 ? F_PG0 : F_PG1;
            const uint16_t fs = part == 0 ? F_PS0 : F_PS1;
            if(!(info & (fg | fs))) continue;
            const double vr = pp.V[r];
            const double vo = pp.V[r + pp.off];
            const int coord = pp.axis == 0 ? x : (pp.axis == 1 ? y : z);
            double ps = 0.0;
            if(info & fs)
            {
                const int cc = pp.cmap[coord];
                long ip;
                if(pp.axis == 0)      ip = cc + pp.psi_pitch * (z + (long)a.lz * y);
                else if(pp.axis == 1) ip = x + a.px * (z + (long)a.lz * cc);
                else                  ip = x + a.px * (cc + (long)pp.nact * y);
                const double cv = pp.c[coord];
                ps = dm(pp.b[coord], pp.psi[ip]);
                ps = axpy1(ps,  cv, vr);
                ps = axpy1(ps, -cv, vo);
                pp.psi[ip] = ps;
            }
            if(info & fg)
            {
                const double Fv = pp.F[coord];
                t = axpy1(t,  Fv, vr);
                t = axpy1(t, -Fv, vo);
                if(info & fs) t = axpy1(t, pp.Db, ps);
            }
        }
        if(pmlOnD) { dv = t; dDirty = true; } else u = t;
    }

    // D2E (FDTD_MANAGER/parallelFDTDField.hpp:1452-1473)
    if(IS_E && (info & F_D2E))
    {
        // DtoU (UTIL/FDTD_up_eq.cpp:838-848): E = D; E *= 1/eps; E += (-1/eps) P_p for every pole grid
        u = dm(ce.inv_eps, dv);
#pragma unroll
        for(int p = 0; p < MAX_POLES; ++p)
            if(p < np) u = axpy1(u, ce.neg_inv_eps, pn[p]);
    }
    else if(IS_E && (info & F_ORD2E))
    {
        // orDipDtoU / orDipDtoUZ (UTIL/FDTD_up_eq.cpp:862-889)
        u = dm(ce.inv_eps, dv);
        for(int p = 0; p < ca.nordip; ++p)
        {
            const double p0 = node_value(a, ca.oP[p], x, y, z);
            if(ca.ord_zvariant)
                u = axpy1(u, ce.neg_inv_eps, p0);
            else
            {
                const double p1 = node_value(a, ca.oP[p], x + ca.ord_dx, y + ca.ord_dy, z + ca.ord_dz);
                u = axpy1(u, ce.neg_half_inv_eps, p0);
                u = axpy1(u, ce.neg_half_inv_eps, p1);
            }
        }
    }

    ca.U[r] = u;
    if(dDirty) ca.D[r] = dv;
}

// One thread per cell; the three components that share a cell index share their neighbour loads
// through L1.  grid = (ceil(lx/BX), ceil(lz/BZ), ly).
template <bool IS_E>
__global__ void __launch_bounds__(256) k_update(const __grid_constant__ StepArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = blockIdx.y * blockDim.y + threadIdx.y;
    const int y = blockIdx.z;
    if(x >= a.lx || z >= a.lz) return;
    const long row = z + (long)a.lz * y;
    const long r = x + a.px * row;
#pragma unroll
    for(int c = 0; c < 3; ++c)
        if(a.c[c].U) update_component<IS_E>(a, a.c[c], r, row, x, y, z);
}

// updatePolE, oriented-dipole poles at the integer nodes
// (FDTD_MANAGER/parallelFDTDField.hpp:1350-1354 -> UTIL/FDTD_up_eq.cpp:450-631)
__global__ void __launch_bounds__(256) k_ordip_poles(const __grid_constant__ NodeArgs a)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int z = blockIdx.y * blockDim.y + threadIdx.y;
    const int y = blockIdx.z;
    if(x >= a.lx || z >= a.lz) return;
    const long row = z + (long)a.lz * y;
    const long r = x + a.px * row;
    const uint16_t info = a.info[r];
    if(info == 0) return;
    const ClassEntry& ce = a.cls[info & CLS_MASK];
    const long ip = a.sp_base[row] + (x - a.sp_xmin[row]);
    double e0[3] = {0.0, 0.0, 0.0}, e1[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for(int c = 0; c < 3; ++c)
        if(a.E[c]) { e0[c] = a.E[c][r]; e1[c] = a.E[c][r + a.eoff[c]]; }
    const bool planar = a.E[0] != nullptr;     // 3-D or TE; otherwise the TM (Ez only) variant
    for(int p = 0; p < ce.npoles; ++p)
    {
        double pc[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for(int c = 0; c < 3; ++c)
            if(a.E[c])
            {
                pc[c] = dm(ce.alpha[p], a.Pcur[c][p][ip]);
                pc[c] = axpy1(pc[c], ce.xi[p], a.Pnew[c][p][ip]);
            }
        if(planar)
        {
            double dotU = 0.0;
            bool first = true;
#pragma unroll
            for(int c = 0; c < 3; ++c)
            {
                if(!a.E[c]) continue;
                const double dp = ce.dip[p][c];
                const double t0 = __ddiv_rn(dm(dp, e0[c]), 2.0);   // multAvg: x*y/2.0 (UTIL/utilityFxns.hpp:38)
                const double t1 = __ddiv_rn(dm(dp, e1[c]), 2.0);
                if(first) { dotU = da(t0, t1); first = false; }
                else { dotU = da(dotU, t0); dotU = da(dotU, t1); }
            }
#pragma unroll
            for(int c = 0; c < 3; ++c)
                if(a.E[c]) pc[c] = axpy1(pc[c], ce.gamma[p], dm(ce.dip[p][c], dotU));
        }
        else
        {
            const double dotU = dm(ce.dip[p][2], e0[2]);
            pc[2] = axpy1(pc[2], ce.gamma[p], dotU);
        }
#pragma unroll
        for(int c = 0; c < 3; ++c)
            if(a.E[c]) a.Pnew[c][p][ip] = pc[c];
    }
}

// src->addPul (SOURCE/parallelSourceNormal.cpp:15-37): grid[box] += dt*Re(sum pulse(t)); the product is formed on the host
__global__ void k_source(double* field, int lx0, int lz0, int ly0, int sx, int sz, int sy, int lz, long px, const double* amp)
{
    const long n = (long)sx * sz * sy;
    const double av = *amp;
    for(long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    {
        const int ix = (int)(i % sx);
        const int iz = (int)((i / sx) % sz);
        const int iy = (int)(i / ((long)sx * sz));
        const long r = (lx0 + ix) + px * ((lz0 + iz) + (long)lz * (ly0 + iy));
        field[r] = da(field[r], av);
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
