// The fused half-step kernels (H half step: updateH + updateHPML_; E half step: updatePolE (isotropic),
// updateD, updateE, updateEPML_, D2E) -- reference FDTD_MANAGER/parallelFDTDField.hpp:1228-1303.
// Included by chiml_kernels.cuh.
//
// Work decomposition.  The grid is cut into tiles of TILE_X x TILE_Z cells of one y plane; one thread
// block processes one tile, each thread two x-adjacent cells and every component of the family at
// them.  At commit time every tile of every family is classified from the painted cell-info planes:
//   FAST     every component's updated cells form a rectangle of plain interior curl cells of one class
//   UNIFORM  every component's updated cells form a rectangle with ONE info value whose flags need no
//            per-cell data beyond the CPML coordinate tables (curl, D target, CPML parts, pole-free D->E)
//   GENERAL  anything else (dispersive cells, material boundaries, ...): per-cell info is read
// and three compact tile lists per family are built.  k_fast / k_uniform never read the cell-info
// planes and run block-uniform straight-line code with all loads issued up front; they are small
// (about 60 registers) so that enough warps are resident to cover HBM latency.  The stencil is the
// reference's (derivOff of FDTD_MANAGER/parallelFDTDField.cpp:80-82,248-250): component c reads
// grid_j = other[(c+1)%3] one cell along axis (c+2)%3 and grid_k = other[(c+2)%3] one cell along
// axis (c+1)%3, backwards for E, forwards for H; chiml_gpu_commit verifies that the lists agree.
// Field arrays carry one plane of zeroed slack on either side, so the unconditional neighbour loads
// of ghost / padding cells stay in bounds.
#pragma once

namespace chiml {

// which components exist (FDTD_MANAGER/parallelFDTDField.hpp:391-443): TE = Ex,Ey,Hz; TM = Ez,Hx,Hy
template <int MODE> __host__ __device__ constexpr bool has_E(int c) { return MODE == CHIML_MODE_3D || (MODE == CHIML_MODE_TE ? c != 2 : c == 2); }
template <int MODE> __host__ __device__ constexpr bool has_H(int c) { return MODE == CHIML_MODE_3D || (MODE == CHIML_MODE_TE ? c == 2 : c != 2); }
template <bool IS_E, int MODE> __host__ __device__ constexpr bool has_own(int c) { return IS_E ? has_E<MODE>(c) : has_H<MODE>(c); }
template <bool IS_E, int MODE> __host__ __device__ constexpr bool has_other(int c) { return IS_E ? has_H<MODE>(c) : has_E<MODE>(c); }

// ---------------------------------------------------------------------------------------------------
// general path: one field component C of one cell, driven by its cell-info word
// ---------------------------------------------------------------------------------------------------
// u = current value of U[r]; (vj_r, vj_n) = grid_j at ind and ind_k, (vk_r, vk_n) = grid_k at ind and ind_j --
// the four stencil values of TwoCompCurl, which are also the stencil values of the two CPML parts (part 0:
// grid_k, derivative along j = (C+1)%3; part 1: grid_j, derivative along k = (C+2)%3).
// Returns true when U[r] must be written back.
// DP: the family has D-like arrays and pole pools (always for E; for H only with magnetic-dispersive media: B, M, mu_inf)
template <bool IS_E, int MODE, int C, bool DP = IS_E>
__device__ __forceinline__ bool update_cell(const StepArgs& a, const CompArgs& ca, const unsigned info, double& u,
                                            const double vj_r, const double vj_n, const double vk_r, const double vk_n,
                                            const long r, const long row, const int x, const int y, const int z)
{
    if(info == 0) return false;
    // D is needed by almost every cell of a general E tile: fetch it before anything that depends on the class table
    const double dIn = (DP && ca.D) ? ca.D[r] : 0.0;
    const ClassEntry& ce = ca.cls[info & CLS_MASK];
    constexpr bool HAS_VJ = has_other<IS_E, MODE>((C + 1) % 3);
    constexpr bool HAS_VK = has_other<IS_E, MODE>((C + 2) % 3);

    double pn[MAX_POLES];
    int np = 0;

    // updatePolE, isotropic poles (parallelFDTDField.hpp:1355-1361 -> UTIL/FDTD_up_eq.cpp:435-446):
    // tmp = P; P = alpha*P; P += xi*Pprev; P += gamma*E^n; Pprev = tmp
    if(DP && (info & F_D2E))
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

    // updateChiE / updateChiH (:1392-1447 -> UpdateChiral, UTIL/FDTD_up_eq.cpp:64-111): chiral poles driven by the other family's component C at
    // the eight corners r, j, k, j+k-r, i, i+j-r, i+k-r, i+j+k-2r of the list entry, current values then previous ones
    double cn[MAX_CHI];
    int nc = 0;
    if(DP && (info & F_D2E))
    {
        nc = ce.nchi;
        if(nc > 0)
        {
            const long ip = ca.sp_base[row] + (x - ca.sp_xmin[row]);
            const long o2 = ca.chi_oi, o3 = ca.chi_oj, o4 = ca.chi_ok;
            const long off8[8] = {0, o3, o4, o3 + o4, o2, o2 + o3, o2 + o4, o2 + o3 + o4};
            const double* __restrict__ opp = a.fam[C];
            double ov[8], pv[8];
#pragma unroll
            for(int k = 0; k < 8; ++k) { ov[k] = opp[r + off8[k]]; pv[k] = ca.oppPrev[r + off8[k]]; }
#pragma unroll
            for(int p = 0; p < MAX_CHI; ++p)
            {
                if(p < nc)
                {
                    double t = dm(ce.chi_alpha[p], ca.chiCur[p][ip]);
                    t = axpy1(t, ce.chi_xi[p], ca.chiNew[p][ip]);
#pragma unroll
                    for(int k = 0; k < 8; ++k) t = axpy1(t, ce.chi_g8[p], ov[k]);
#pragma unroll
                    for(int k = 0; k < 8; ++k) t = axpy1(t, ce.chi_gp8[p], pv[k]);
                    ca.chiNew[p][ip] = t;
                    cn[p] = t;
                }
            }
        }
    }

    const bool pmlCell = (info & (F_PG0 | F_PS0 | F_PG1 | F_PS1)) != 0;
    const bool pmlOnD = DP && a.pml_on_D;
    const bool needD = DP && ((info & (F_ISD | F_D2E | F_ORD2E)) || (pmlOnD && pmlCell));
    double dv = needD ? dIn : 0.0;
    bool dDirty = false;

    // updateD / updateE / updateH: TwoCompCurl, OneCompCurlJ, OneCompCurlK (UTIL/FDTD_up_eq.cpp:10-35)
    if(info & F_CURL)
    {
        double t = (DP && (info & F_ISD)) ? dv : u;
        if(HAS_VJ)
        {
            t = axpy1(t,  ce.pf2, vj_r);
            t = axpy1(t, -ce.pf2, vj_n);
        }
        if(HAS_VK)
        {
            t = axpy1(t, -ce.pf1, vk_r);
            t = axpy1(t,  ce.pf1, vk_n);
        }
        if(DP && (info & F_ISD)) { dv = t; dDirty = true; } else u = t;
    }

    // parallelCPML<T>::updateGrid (PML/parallelPML.hpp:693-697): part 0 then part 1; each part is
    // updatePsiField then the grid daxpys (PML/parallelPML.cpp:12-40)
    if(pmlCell)
    {
        double t = pmlOnD ? dv : u;
#pragma unroll
        for(int part = 0; part < 2; ++part)
        {
            if(part == 0 ? !HAS_VK : !HAS_VJ) continue;
            const PmlArgs& pp = ca.pml[part];
            const unsigned fg = part == 0 ? F_PG0 : F_PG1;
            const unsigned fs = part == 0 ? F_PS0 : F_PS1;
            if(!(info & (fg | fs))) continue;
            constexpr int AX0 = (C + 1) % 3, AX1 = (C + 2) % 3;
            const int axis = part == 0 ? AX0 : AX1;          // compile-time after unrolling
            const double vr = part == 0 ? vk_r : vj_r;
            const double vo = part == 0 ? vk_n : vj_n;
            const int coord = axis == 0 ? x : (axis == 1 ? y : z);
            double ps = 0.0;
            if(info & fs)
            {
                const int cc = pp.cmap[coord];
                long ip;
                if(axis == 0)      ip = cc + pp.psi_pitch * row;
                else if(axis == 1) ip = x + a.px * (z + (long)a.lz * cc);
                else               ip = x + a.px * (cc + (long)pp.nact * y);
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

    // D2E (parallelFDTDField.hpp:1452-1473)
    if(DP && (info & F_D2E))
    {
        // DtoU (UTIL/FDTD_up_eq.cpp:838-848): E = D; E *= 1/eps; E += (-1/eps) P_p for every pole grid
        u = dm(ce.inv_eps, dv);
#pragma unroll
        for(int p = 0; p < MAX_POLES; ++p)
            if(p < np) u = axpy1(u, ce.neg_inv_eps, pn[p]);
        // chiDtoU (:920-925) with epMuInfty = -eps (E) / +mu (H)
#pragma unroll
        for(int p = 0; p < MAX_CHI; ++p)
            if(p < nc) u = axpy1(u, ce.chi_fac, cn[p]);
    }
    else if(IS_E && (info & F_ORD2E))
    {
        // orDipDtoU / orDipDtoUZ (UTIL/FDTD_up_eq.cpp:862-889)
        u = dm(ce.inv_eps, dv);
        for(int p = 0; p < ca.nordip; ++p)
        {
            const double p0 = node_value(a, ca.oP[p], ca.oPg[p], x, y, z);
            if(ca.ord_zvariant)
                u = axpy1(u, ce.neg_inv_eps, p0);
            else
            {
                const double p1 = node_value(a, ca.oP[p], ca.oPg[p], x + ca.ord_dx, y + ca.ord_dy, z + ca.ord_dz);
                u = axpy1(u, ce.neg_half_inv_eps, p0);
                u = axpy1(u, ce.neg_half_inv_eps, p1);
            }
        }
    }

    if(dDirty) ca.D[r] = dv;
    return true;
}

// ---------------------------------------------------------------------------------------------------
// shared pieces of the three tile kernels
// ---------------------------------------------------------------------------------------------------
// Values of V at the stencil neighbour of cells (x, x+1): one cell along AXIS in direction SIGN.
// own2 = V[r], V[r+1].  Row pitch and plane stride are multiples of 16 doubles, so the y / z
// neighbours are aligned 16-byte loads; the x neighbour needs one extra scalar.
template <int AXIS, int SIGN>
__device__ __forceinline__ double2 neighbour2(const double* V, const long r, const long px, const long plane, const double2 own2)
{
    if(AXIS == 0) return SIGN > 0 ? make_double2(own2.y, V[r + 2]) : make_double2(V[r - 1], own2.x);
    const long off = (AXIS == 2 ? px : plane) * SIGN;
    return *reinterpret_cast<const double2*>(V + r + off);
}

// all loads of one thread: own family u, other family v at r, and the six stencil neighbours
template <bool IS_E, int MODE>
struct PairLoads
{
    double2 u[3], v[3], nj[3], nk[3];
    __device__ __forceinline__ void load(const StepArgs& a, const long r)
    {
        constexpr int S = IS_E ? -1 : 1;
        const long plane = a.px * a.lz;
#pragma unroll
        for(int c = 0; c < 3; ++c) u[c] = v[c] = nj[c] = nk[c] = make_double2(0.0, 0.0);
#pragma unroll
        for(int c = 0; c < 3; ++c)
        {
            if(has_own<IS_E, MODE>(c)) u[c] = *reinterpret_cast<const double2*>(a.c[c].U + r);
            if(has_other<IS_E, MODE>(c)) v[c] = *reinterpret_cast<const double2*>(a.fam[c] + r);
        }
        // component c: grid_j = other[(c+1)%3] along axis (c+2)%3; grid_k = other[(c+2)%3] along axis (c+1)%3
        if(has_own<IS_E, MODE>(0))
        {
            if(has_other<IS_E, MODE>(1)) nj[0] = neighbour2<2, S>(a.fam[1], r, a.px, plane, v[1]);
            if(has_other<IS_E, MODE>(2)) nk[0] = neighbour2<1, S>(a.fam[2], r, a.px, plane, v[2]);
        }
        if(has_own<IS_E, MODE>(1))
        {
            if(has_other<IS_E, MODE>(2)) nj[1] = neighbour2<0, S>(a.fam[2], r, a.px, plane, v[2]);
            if(has_other<IS_E, MODE>(0)) nk[1] = neighbour2<2, S>(a.fam[0], r, a.px, plane, v[0]);
        }
        if(has_own<IS_E, MODE>(2))
        {
            if(has_other<IS_E, MODE>(0)) nj[2] = neighbour2<1, S>(a.fam[0], r, a.px, plane, v[0]);
            if(has_other<IS_E, MODE>(1)) nk[2] = neighbour2<0, S>(a.fam[1], r, a.px, plane, v[1]);
        }
    }
};

// Which tile and which tile row a thread works on.  3-D: one block per tile, threadIdx.y = z row.  2-D grids have one-row tiles: a
// block takes ROWS_2D consecutive tiles of the list, one per threadIdx.y (the warps of these kernels never synchronise).
constexpr int ROWS_2D = 8;
template <int MODE>
__device__ __forceinline__ bool tile_of_thread(const unsigned ntiles, unsigned& tile, int& zl)
{
    if(MODE == CHIML_MODE_3D) { tile = blockIdx.x; zl = threadIdx.y; return true; }
    tile = blockIdx.x * ROWS_2D + threadIdx.y; zl = 0;
    return tile < ntiles;
}

// rectangle of updated cells of one component inside the tile: bytes xlo, xhi, zlo, zhi (hi exclusive)
__device__ __forceinline__ void rect_mask(const unsigned rect, const int xl, const int zl, bool& m0, bool& m1)
{
    const int xlo = rect & 0xFF, xhi = (rect >> 8) & 0xFF, zlo = (rect >> 16) & 0xFF, zhi = rect >> 24;
    const bool zin = zl >= zlo && zl < zhi;
    m0 = zin && xl >= xlo && xl < xhi;
    m1 = zin && xl + 1 >= xlo && xl + 1 < xhi;
}
__device__ __forceinline__ void store_pair(double* p, const double2 t, const bool m0, const bool m1)
{
    if(m0 && m1) *reinterpret_cast<double2*>(p) = t;
    else if(m0) p[0] = t.x;
    else if(m1) p[1] = t.y;
}

// ---------------------------------------------------------------------------------------------------
// k_fast: tiles whose updated cells are plain interior curl cells (TwoCompCurl / OneCompCurlJ / K)
// ---------------------------------------------------------------------------------------------------
// The block marches along y over the t.ny planes of its tile column.  The two stencil neighbours that lie one plane away in y
// (E half step: H_z[y-1] for E_x, H_x[y-1] for E_z; H half step: E_z[y+1] for H_x, E_x[y+1] for H_z) are carried in registers
// from one plane to the next, so every array crosses HBM exactly once per half step however far apart in time the tiles of
// neighbouring planes would otherwise run (a whole y plane of tiles is ~150 MB of traffic: more than L2 holds).
template <bool IS_E, int MODE>
__device__ __forceinline__ void fast_tile(const StepArgs& a, const TileRec& t, const int xl, const int zl)
{
    const int x = t.x0 + xl, z = t.z0 + zl;
    if(x >= a.px || z >= a.lz) return;
    constexpr int S = IS_E ? -1 : 1;
    const long plane = a.px * a.lz;
    long r = x + a.px * (z + (long)a.lz * t.y);
    const int ny = t.ny;
    bool m0[3], m1[3];
    double2 pf[3];
#pragma unroll
    for(int c = 0; c < 3; ++c)
    {
        m0[c] = m1[c] = false;
        if(has_own<IS_E, MODE>(c) && t.rect[c] != 0) rect_mask(t.rect[c], xl, zl, m0[c], m1[c]);
        pf[c] = t.pf[c];
    }
    // restrict WITHOUT const: the no-alias promise lets the loads move above the stores, but the loads stay coherent ld.global (a
    // const restrict pointer makes them ld.global.nc, which the persistent multi-step kernel k_steps_2d must not use: there the
    // arrays read in one half step are written in the other half step of the same launch)
    double* __restrict__ f0 = const_cast<double*>(a.fam[0]);
    double* __restrict__ f1 = const_cast<double*>(a.fam[1]);
    double* __restrict__ f2 = const_cast<double*>(a.fam[2]);
    // carried planes of the y-coupled arrays: E half step: the plane below; H half step: the current plane (loaded as "next" before)
    constexpr bool Y0 = has_own<IS_E, MODE>(2) && has_other<IS_E, MODE>(0);   // own z reads other x one plane away
    constexpr bool Y2 = has_own<IS_E, MODE>(0) && has_other<IS_E, MODE>(2);   // own x reads other z one plane away
    double2 c0 = make_double2(0.0, 0.0), c2 = make_double2(0.0, 0.0);
    if(IS_E)
    {
        if(Y0) c0 = *reinterpret_cast<const double2*>(f0 + r - plane);
        if(Y2) c2 = *reinterpret_cast<const double2*>(f2 + r - plane);
    }
    else
    {
        if(Y0) c0 = *reinterpret_cast<const double2*>(f0 + r);
        if(Y2) c2 = *reinterpret_cast<const double2*>(f2 + r);
    }
    for(int iy = 0; iy < ny; ++iy, r += plane)
    {
        double2 v[3], nj[3], nk[3], u[3];
#pragma unroll
        for(int c = 0; c < 3; ++c) v[c] = nj[c] = nk[c] = u[c] = make_double2(0.0, 0.0);
#pragma unroll
        for(int c = 0; c < 3; ++c)
            if(has_own<IS_E, MODE>(c)) u[c] = *reinterpret_cast<const double2*>(a.c[c].U + r);
        double2 n0 = make_double2(0.0, 0.0), n2 = make_double2(0.0, 0.0);   // H half step: the plane above
        if(IS_E)
        {
            if(has_other<IS_E, MODE>(0)) v[0] = *reinterpret_cast<const double2*>(f0 + r);
            if(has_other<IS_E, MODE>(2)) v[2] = *reinterpret_cast<const double2*>(f2 + r);
        }
        else
        {
            if(Y0) { v[0] = c0; n0 = *reinterpret_cast<const double2*>(f0 + r + plane); }
            else if(has_other<IS_E, MODE>(0)) v[0] = *reinterpret_cast<const double2*>(f0 + r);
            if(Y2) { v[2] = c2; n2 = *reinterpret_cast<const double2*>(f2 + r + plane); }
            else if(has_other<IS_E, MODE>(2)) v[2] = *reinterpret_cast<const double2*>(f2 + r);
        }
        if(has_other<IS_E, MODE>(1)) v[1] = *reinterpret_cast<const double2*>(f1 + r);
        // component c: grid_j = other[(c+1)%3] along axis (c+2)%3; grid_k = other[(c+2)%3] along axis (c+1)%3
        if(has_own<IS_E, MODE>(0))
        {
            if(has_other<IS_E, MODE>(1)) nj[0] = neighbour2<2, S>(f1, r, a.px, plane, v[1]);
            if(has_other<IS_E, MODE>(2)) nk[0] = IS_E ? c2 : n2;
        }
        if(has_own<IS_E, MODE>(1))
        {
            if(has_other<IS_E, MODE>(2)) nj[1] = neighbour2<0, S>(f2, r, a.px, plane, v[2]);
            if(has_other<IS_E, MODE>(0)) nk[1] = neighbour2<2, S>(f0, r, a.px, plane, v[0]);
        }
        if(has_own<IS_E, MODE>(2))
        {
            if(has_other<IS_E, MODE>(0)) nj[2] = IS_E ? c0 : n0;
            if(has_other<IS_E, MODE>(1)) nk[2] = neighbour2<0, S>(f1, r, a.px, plane, v[1]);
        }
#pragma unroll
        for(int c = 0; c < 3; ++c)
        {
            if(!has_own<IS_E, MODE>(c)) continue;
            if(!(m0[c] || m1[c])) continue;
            const double2 vj = v[(c + 1) % 3], vk = v[(c + 2) % 3];
            double2 w = u[c];
            if(has_other<IS_E, MODE>((c + 1) % 3))
            {
                w.x = axpy1(w.x,  pf[c].y, vj.x);    w.y = axpy1(w.y,  pf[c].y, vj.y);
                w.x = axpy1(w.x, -pf[c].y, nj[c].x); w.y = axpy1(w.y, -pf[c].y, nj[c].y);
            }
            if(has_other<IS_E, MODE>((c + 2) % 3))
            {
                w.x = axpy1(w.x, -pf[c].x, vk.x);    w.y = axpy1(w.y, -pf[c].x, vk.y);
                w.x = axpy1(w.x,  pf[c].x, nk[c].x); w.y = axpy1(w.y,  pf[c].x, nk[c].y);
            }
            store_pair(a.c[c].U + r, w, m0[c], m1[c]);
        }
        if(IS_E) { c0 = v[0]; c2 = v[2]; } else { if(Y0) c0 = n0; if(Y2) c2 = n2; }
    }
}

template <bool IS_E, int MODE>
__global__ void __launch_bounds__(256, 2) k_fast(const __grid_constant__ StepArgs a, const TileRec* __restrict__ tiles, const unsigned ntiles)
{
    unsigned ti; int zl;
    if(!tile_of_thread<MODE>(ntiles, ti, zl)) return;
    fast_tile<IS_E, MODE>(a, tiles[ti], 2 * threadIdx.x, zl);
}

// ---------------------------------------------------------------------------------------------------
// k_uniform: tiles where each component has ONE info value over a rectangle: curl into E/H or D, CPML parts
// (psi recursion + grid terms) on E/H or D, pole-free D->E.  Block-uniform control flow, no cell-info reads.
// ---------------------------------------------------------------------------------------------------
constexpr unsigned FL_RUNTIME = 0xFFFFFFFFu;

// FL: the flag byte of the rectangle's info value as a compile-time constant (every `flags & X` below folds away), or FL_RUNTIME
// CPML coefficients that do not change from plane to plane: per-x pairs (derivative along x), z scalars (along z)
struct PmlCoef { double2 F[2], b[2], c[2]; int2 cc[2]; int cmz[2]; };

template <bool IS_E, int MODE, int C>
__device__ __forceinline__ void pml_coef_static(const StepArgs& a, const unsigned info, const int x, const int z, PmlCoef& k)
{
    const CompArgs& ca = a.c[C];
#pragma unroll
    for(int part = 0; part < 2; ++part)
    {
        k.F[part] = k.b[part] = k.c[part] = make_double2(0.0, 0.0);
        k.cc[part] = make_int2(0, 0); k.cmz[part] = 0;
        if(part == 0 ? !has_other<IS_E, MODE>((C + 2) % 3) : !has_other<IS_E, MODE>((C + 1) % 3)) continue;
        const PmlArgs& pp = ca.pml[part];
        const unsigned fg = part == 0 ? F_PG0 : F_PG1, fs = part == 0 ? F_PS0 : F_PS1;
        if(!(info & (fg | fs))) continue;
        constexpr int AX0 = (C + 1) % 3, AX1 = (C + 2) % 3;
        const int axis = part == 0 ? AX0 : AX1;
        if(axis == 0)
        {
            k.F[part] = *reinterpret_cast<const double2*>(pp.F + x);
            if(info & fs)
            {
                k.b[part] = *reinterpret_cast<const double2*>(pp.b + x);
                k.c[part] = *reinterpret_cast<const double2*>(pp.c + x);
                k.cc[part] = *reinterpret_cast<const int2*>(pp.cmap + x);
            }
        }
        else if(axis == 2)
        {
            const double f = pp.F[z];
            k.F[part] = make_double2(f, f);
            if(info & fs)
            {
                const double bb = pp.b[z], cc = pp.c[z];
                k.b[part] = make_double2(bb, bb); k.c[part] = make_double2(cc, cc);
                k.cmz[part] = pp.cmap[z];
            }
        }
    }
}

// HOISTED: the caller marches a column of planes and passes the masks and the plane-independent coefficients it computed once
// PT: isotropic poles known at compile time (0 none, 1 some) or decided by `np` at run time (-1)
template <bool IS_E, int MODE, int C, unsigned FL, bool HOISTED = false, int PT = -1>
__device__ __forceinline__ void uniform_rect(const StepArgs& a, const unsigned rect, const unsigned info_rt, const double2 pfc, const double inv_eps,
                                             const PairLoads<IS_E, MODE>& L,
                                             const long r, const long row, const int x, const int y, const int z, const int xl, const int zl,
                                             const int np = 0, const bool hm0 = false, const bool hm1 = false, const PmlCoef* hc = nullptr, const int ycm = -2)
{
    // info_rt: the rectangle's info value (class bits always valid); its flag byte equals FL when FL is a compile-time constant
    const unsigned info = FL == FL_RUNTIME ? info_rt : FL;
    const CompArgs& ca = a.c[C];
    constexpr bool HAS_VJ = has_other<IS_E, MODE>((C + 1) % 3);
    constexpr bool HAS_VK = has_other<IS_E, MODE>((C + 2) % 3);
    bool m0 = hm0, m1 = hm1;
    PmlCoef own;
    if(!HOISTED)
    {
        rect_mask(rect, xl, zl, m0, m1);
        if(!(m0 || m1)) return;
        pml_coef_static<IS_E, MODE, C>(a, info, x, z, own);
    }
    const PmlCoef& kc = HOISTED ? *hc : own;
    const double2 vj = L.v[(C + 1) % 3], vk = L.v[(C + 2) % 3];
    const double2 nj = L.nj[C], nk = L.nk[C];
    double2 u = L.u[C];
    const double2 uOld = u;          // E^n: what the isotropic poles are driven by (they are updated before E in the reference)

    const bool pmlCell = (info & (F_PG0 | F_PS0 | F_PG1 | F_PS1)) != 0;
    const bool pmlOnD = IS_E && a.pml_on_D;
    const bool needD = IS_E && ((info & (F_ISD | F_D2E | F_ORD2E)) || (pmlOnD && pmlCell));

    // ---- every load of this component is issued here, before any arithmetic waits on one of them -------------------------
    double2 dv = make_double2(0.0, 0.0);
    if(needD) dv = *reinterpret_cast<const double2*>(ca.D + r);
    double2 psv[2], Fv[2], bv[2], cv[2];
    long pip[2] = {0, 0};
    int2 pcc[2];
#pragma unroll
    for(int part = 0; part < 2; ++part)
    {
        psv[part] = make_double2(0.0, 0.0);
        Fv[part] = kc.F[part]; bv[part] = kc.b[part]; cv[part] = kc.c[part]; pcc[part] = kc.cc[part];
        if(part == 0 ? !HAS_VK : !HAS_VJ) continue;
        const PmlArgs& pp = ca.pml[part];
        const unsigned fg = part == 0 ? F_PG0 : F_PG1;
        const unsigned fs = part == 0 ? F_PS0 : F_PS1;
        if(!(info & (fg | fs))) continue;
        constexpr int AX0 = (C + 1) % 3, AX1 = (C + 2) % 3;
        const int axis = part == 0 ? AX0 : AX1;
        if(axis == 0)
        {
            if(info & fs)
            {
                pip[part] = pp.psi_pitch * row;
                if(m0) psv[part].x = pp.psi[pip[part] + pcc[part].x];
                if(m1) psv[part].y = pp.psi[pip[part] + pcc[part].y];
            }
        }
        else if(axis == 1)
        {
            const double f = pp.F[y];
            Fv[part] = make_double2(f, f);
            if(info & fs)
            {
                const double bb = pp.b[y], cc = pp.c[y];
                bv[part] = make_double2(bb, bb); cv[part] = make_double2(cc, cc);
                // ycm: the compact y coordinate of this plane, fetched by the marching caller one plane ahead (so that the psi load does
                // not wait for a table lookup first)
                const int cm = ycm != -2 ? ycm : pp.cmap[y];
                pip[part] = x + a.px * (z + (long)a.lz * cm);
                psv[part] = *reinterpret_cast<const double2*>(pp.psi + pip[part]);
            }
        }
        else if(info & fs)
        {
            pip[part] = x + a.px * (kc.cmz[part] + (long)pp.nact * y);
            psv[part] = *reinterpret_cast<const double2*>(pp.psi + pip[part]);
        }
    }

    // oriented-dipole node polarisation of the first two poles, issued with the other loads when the flag byte is a compile-time
    // constant (span lookup -> pool value is a dependent chain of two memory latencies; left to the end it is paid on top of
    // everything else; in the generic body the extra live registers cost more than they save)
    constexpr bool HOIST_NODE = IS_E && FL != FL_RUNTIME && (FL & F_ORD2E) != 0;
    constexpr int NODE_HOIST = 2;
    double2 np0[NODE_HOIST], np1[NODE_HOIST];
    if constexpr(HOIST_NODE)
    {
#pragma unroll
        for(int p = 0; p < NODE_HOIST; ++p)
        {
            np0[p] = np1[p] = make_double2(0.0, 0.0);
            if(p >= ca.nordip) continue;
            np0[p] = node_pair(a, ca.oP[p], ca.oPg[p], x, y, z);
            if(!ca.ord_zvariant) np1[p] = node_pair(a, ca.oP[p], ca.oPg[p], x + ca.ord_dx, y + ca.ord_dy, z + ca.ord_dz);
        }
    }

    bool dDirty = false;
    if(info & F_CURL)
    {
        const double2 pf = pfc;
        double2 w = (IS_E && (info & F_ISD)) ? dv : u;
        if(HAS_VJ)
        {
            w.x = axpy1(w.x,  pf.y, vj.x); w.y = axpy1(w.y,  pf.y, vj.y);
            w.x = axpy1(w.x, -pf.y, nj.x); w.y = axpy1(w.y, -pf.y, nj.y);
        }
        if(HAS_VK)
        {
            w.x = axpy1(w.x, -pf.x, vk.x); w.y = axpy1(w.y, -pf.x, vk.y);
            w.x = axpy1(w.x,  pf.x, nk.x); w.y = axpy1(w.y,  pf.x, nk.y);
        }
        if(IS_E && (info & F_ISD)) { dv = w; dDirty = true; } else u = w;
    }
    if(pmlCell)
    {
        double2 w = pmlOnD ? dv : u;
#pragma unroll
        for(int part = 0; part < 2; ++part)
        {
            if(part == 0 ? !HAS_VK : !HAS_VJ) continue;
            const PmlArgs& pp = ca.pml[part];
            const unsigned fg = part == 0 ? F_PG0 : F_PG1;
            const unsigned fs = part == 0 ? F_PS0 : F_PS1;
            if(!(info & (fg | fs))) continue;
            constexpr int AX0 = (C + 1) % 3, AX1 = (C + 2) % 3;
            const int axis = part == 0 ? AX0 : AX1;
            const double2 vr = part == 0 ? vk : vj;
            const double2 vo = part == 0 ? nk : nj;
            double2 ps = make_double2(0.0, 0.0);
            if(info & fs)
            {
                double2 p = psv[part];
                p.x = dm(bv[part].x, p.x);            p.y = dm(bv[part].y, p.y);
                p.x = axpy1(p.x,  cv[part].x, vr.x);  p.y = axpy1(p.y,  cv[part].y, vr.y);
                p.x = axpy1(p.x, -cv[part].x, vo.x);  p.y = axpy1(p.y, -cv[part].y, vo.y);
                if(axis == 0)
                {
                    if(m0) pp.psi[pip[part] + pcc[part].x] = p.x;
                    if(m1) pp.psi[pip[part] + pcc[part].y] = p.y;
                }
                else store_pair(pp.psi + pip[part], p, m0, m1);
                ps = p;
            }
            if(info & fg)
            {
                w.x = axpy1(w.x,  Fv[part].x, vr.x); w.y = axpy1(w.y,  Fv[part].y, vr.y);
                w.x = axpy1(w.x, -Fv[part].x, vo.x); w.y = axpy1(w.y, -Fv[part].y, vo.y);
                if(info & fs) { w.x = axpy1(w.x, pp.Db, ps.x); w.y = axpy1(w.y, pp.Db, ps.y); }
            }
        }
        if(pmlOnD) { dv = w; dDirty = true; } else u = w;
    }
    if(IS_E && (info & F_D2E))
    {
        // DtoU (UTIL/FDTD_up_eq.cpp:838-848): E = (1/eps) * D, then E += (-1/eps) P_p for every isotropic pole p, in pole order
        const double ie = inv_eps;
        u.x = dm(ie, dv.x); u.y = dm(ie, dv.y);
        if(PT == 1 || (PT == -1 && np > 0))
        {
            // updatePolE (parallelFDTDField.hpp:1355-1361 -> UTIL/FDTD_up_eq.cpp:435-446): P = alpha P + xi P_prev + gamma E^n, written into
            // the buffer that held P_prev.  The reference runs this before the D update; nothing it reads is written by that update, so
            // doing it here, pole by pole straight into the D->E sum, gives the same bits without holding the new P values in registers
            const ClassEntry& ce = ca.cls[info_rt & CLS_MASK];
            const long ip = ca.sp_base[row] + (x - ca.sp_xmin[row]);
            const double nie = ce.neg_inv_eps;
#pragma unroll 2
            for(int p = 0; p < np; ++p)
            {
                const double al = ce.alpha[p], xi = ce.xi[p], ga = ce.gamma[p];
                double* __restrict__ pc = const_cast<double*>(ca.Pcur[p]) + ip;     // not const: see k_fast
                double* __restrict__ pn = ca.Pnew[p] + ip;
                // the row spans of the pools start at an even x with an even length (build_spans): ip is even, both cells of the pair lie
                // inside the span whenever one of them is updated -> one aligned 16-byte access per pool
                const double2 cc = *reinterpret_cast<const double2*>(pc);
                const double2 oo = *reinterpret_cast<const double2*>(pn);
                double t0 = dm(al, cc.x), t1 = dm(al, cc.y);
                t0 = axpy1(t0, xi, oo.x);     t1 = axpy1(t1, xi, oo.y);
                t0 = axpy1(t0, ga, uOld.x);   t1 = axpy1(t1, ga, uOld.y);
                store_pair(pn, make_double2(t0, t1), m0, m1);
                u.x = axpy1(u.x, nie, t0);    u.y = axpy1(u.y, nie, t1);
            }
        }
    }
    else if(IS_E && (info & F_ORD2E))
    {
        // orDipDtoU / orDipDtoUZ (UTIL/FDTD_up_eq.cpp:862-889): E = D/eps - (1/2eps) sum_p (P_p[r] + P_p[r + e_c]), P at the nodes.
        // -1/eps and -0.5/eps are -(1/eps) and -(0.5 * (1/eps)) exactly (sign flip and scaling by a power of two commute with rounding)
        const double ie = inv_eps, nie = -inv_eps, nhie = -0.5 * inv_eps;
        u.x = dm(ie, dv.x); u.y = dm(ie, dv.y);
        for(int p = 0; p < ca.nordip; ++p)
        {
            double2 p0, p1 = make_double2(0.0, 0.0);
            if(HOIST_NODE && p < NODE_HOIST) { p0 = np0[p & (NODE_HOIST - 1)]; p1 = np1[p & (NODE_HOIST - 1)]; }
            else
            {
                p0 = node_pair(a, ca.oP[p], ca.oPg[p], x, y, z);
                if(!ca.ord_zvariant) p1 = node_pair(a, ca.oP[p], ca.oPg[p], x + ca.ord_dx, y + ca.ord_dy, z + ca.ord_dz);
            }
            if(ca.ord_zvariant) { u.x = axpy1(u.x, nie, p0.x); u.y = axpy1(u.y, nie, p0.y); }
            else
            {
                u.x = axpy1(u.x, nhie, p0.x); u.y = axpy1(u.y, nhie, p0.y);
                u.x = axpy1(u.x, nhie, p1.x); u.y = axpy1(u.y, nhie, p1.y);
            }
        }
    }
    store_pair(ca.U + r, u, m0, m1);
    if(dDirty) store_pair(ca.D + r, dv, m0, m1);
}

// a UNIFORM tile holds, per component, one or two rectangles of cells with one info value each (two: a tile cut by a CPML or
// material boundary); a thread whose two cells lie in different rectangles runs the body once per rectangle
template <bool IS_E, int MODE, int C>
__device__ __forceinline__ void uniform_comp(const StepArgs& a, const TileRec& t, const PairLoads<IS_E, MODE>& L,
                                             const long r, const long row, const int x, const int y, const int z, const int xl, const int zl)
{
    if constexpr(has_own<IS_E, MODE>(C))
    {
#pragma unroll 1
        for(int w = 0; w < 2; ++w)
        {
            const unsigned rect = w == 0 ? t.rect[C] : t.rectB[C];
            if(rect == 0) continue;
            const unsigned info = w == 0 ? t.info[C] : t.infoB[C];
            const double2 pfc = w == 0 ? t.pf[C] : t.pfB[C];
            const double ie = w == 0 ? t.inv_eps[C] : t.inv_epsB[C];
            const int np = (int)(((w == 0 ? t.np : t.npB) >> (8 * C)) & 0xFFu);
            // the flag combinations the reference's lists produce for pole-free cells get a body compiled for exactly that combination
            // (the flag tests, dead branches and zero-initialised operands of the generic body are most of its instructions)
#define CHIML_UCASE(F) case (F): uniform_rect<IS_E, MODE, C, (F)>(a, rect, info, pfc, ie, L, r, row, x, y, z, xl, zl, np); break;
            constexpr bool VJ = has_other<IS_E, MODE>((C + 1) % 3), VK = has_other<IS_E, MODE>((C + 2) % 3);
            switch(info & 0xFF00u)
            {
                CHIML_UCASE(F_CURL)
                default:
                    if constexpr(IS_E)
                    {
                        switch(info & 0xFF00u)
                        {
                            CHIML_UCASE(F_CURL | F_ISD | F_D2E)
                            default:
                                if constexpr(VJ && VK)
                                    switch(info & 0xFF00u)
                                    {
                                        CHIML_UCASE(F_PG0 | F_PG1 | F_D2E)
                                        CHIML_UCASE(F_PG0 | F_PS0 | F_PG1 | F_D2E)
                                        CHIML_UCASE(F_PG0 | F_PG1 | F_PS1 | F_D2E)
                                        CHIML_UCASE(F_PG0 | F_PS0 | F_PG1 | F_PS1 | F_D2E)
                                        CHIML_UCASE(F_PG0 | F_PG1)
                                        CHIML_UCASE(F_PG0 | F_PS0 | F_PG1)
                                        CHIML_UCASE(F_PG0 | F_PG1 | F_PS1)
                                        CHIML_UCASE(F_PG0 | F_PS0 | F_PG1 | F_PS1)
                                        default: uniform_rect<IS_E, MODE, C, FL_RUNTIME>(a, rect, info, pfc, ie, L, r, row, x, y, z, xl, zl, np);
                                    }
                                else uniform_rect<IS_E, MODE, C, FL_RUNTIME>(a, rect, info, pfc, ie, L, r, row, x, y, z, xl, zl, np);
                        }
                    }
                    else if constexpr(VJ && VK)
                    {
                        switch(info & 0xFF00u)
                        {
                            CHIML_UCASE(F_PG0 | F_PG1)
                            CHIML_UCASE(F_PG0 | F_PS0 | F_PG1)
                            CHIML_UCASE(F_PG0 | F_PG1 | F_PS1)
                            CHIML_UCASE(F_PG0 | F_PS0 | F_PG1 | F_PS1)
                            default: uniform_rect<IS_E, MODE, C, FL_RUNTIME>(a, rect, info, pfc, ie, L, r, row, x, y, z, xl, zl, np);
                        }
                    }
                    else uniform_rect<IS_E, MODE, C, FL_RUNTIME>(a, rect, info, pfc, ie, L, r, row, x, y, z, xl, zl, np);
            }
#undef CHIML_UCASE
        }
    }
}

// k_uniform geometry per half step: blocks per 3-D tile along z, and resident blocks per SM the register budget is set for
// (measured on the C5 slab, profiles/README.md r1x/r1y; 2-D grids have one row per tile and use k_uniform_rows)
#ifndef CHIML_SPLIT_E
#define CHIML_SPLIT_E 1
#endif
#ifndef CHIML_OCC_E
#define CHIML_OCC_E 4
#endif
#ifndef CHIML_SPLIT_H
#define CHIML_SPLIT_H 1
#endif
#ifndef CHIML_OCC_H
#define CHIML_OCC_H 4
#endif
#ifndef CHIML_PIPE_Y
#define CHIML_PIPE_Y 0      // measured equal on the C5 slab and on one-axis CPML probes (profiles/README.md r2_05): the L2 prefetch already covers it
#endif
#ifndef CHIML_PREFETCH_PLANES
#define CHIML_PREFETCH_PLANES 1
#endif
template <bool IS_E> __host__ __device__ constexpr int uniform_split() { return IS_E ? CHIML_SPLIT_E : CHIML_SPLIT_H; }
template <bool IS_E> __host__ __device__ constexpr int uniform_occ() { return IS_E ? CHIML_OCC_E : CHIML_OCC_H; }

// L2 prefetch of one 128-byte line (fire and forget: holds no register and no shared memory).  The marching kernels touch every
// array at a fixed plane stride, so the lines of the plane two steps ahead are requested while the current plane is computed; the
// demand loads then pay L2 latency instead of HBM latency, which is what the UNIFORM kernels (few warps, many arrays) are short of.
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
// The same through the TMA unit: ONE instruction of ONE lane asks for a whole 64-cell row (512 bytes, four L2 lines) of a tile,
// described by a tensor map of the array (chiml_gpu_commit builds them); coordinates are (x, z-like, y-like) of the array's own
// layout, out-of-range parts of the box are skipped by the hardware.  Replaces sixteen prefetch.global.L2 plus their address
// arithmetic per row and array.
__device__ __forceinline__ void tma_prefetch_row(const unsigned char* tmaps, const int map, const int c0, const int c1, const int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 :: "l"(tmaps + (size_t)map * TMAP_BYTES), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
constexpr int PREFETCH_PLANES = CHIML_PREFETCH_PLANES;

// psi lines of component C of a UNIFORM tile at plane y (address arithmetic of uniform_rect), for the rectangle with this info
template <bool IS_E, int MODE, int C>
__device__ __forceinline__ void prefetch_psi(const StepArgs& a, const unsigned info, const int x, const int y, const int z)
{
    const CompArgs& ca = a.c[C];
#pragma unroll
    for(int part = 0; part < 2; ++part)
    {
        const unsigned fs = part == 0 ? F_PS0 : F_PS1;
        if(!(info & fs)) continue;
        constexpr int AX0 = (C + 1) % 3, AX1 = (C + 2) % 3;
        const int axis = part == 0 ? AX0 : AX1;
        const PmlArgs& pp = ca.pml[part];
        if(axis == 1)      { const int cm = pp.cmap[y]; if(cm >= 0) prefetch_l2(pp.psi + x + a.px * (z + (long)a.lz * cm)); }
        else if(axis == 2) { const int cm = pp.cmap[z]; if(cm >= 0) prefetch_l2(pp.psi + x + a.px * (cm + (long)pp.nact * y)); }
    }
}

// the same by TMA row prefetch: x0 = first cell of the tile row; the psi arrays' tensor maps follow their compact layouts
// (y-normal slabs: (x, z, compact y); z-normal: (x, compact z, y); x-normal: (compact x, z + lz * y, -))
template <bool IS_E, int MODE, int C>
__device__ __forceinline__ void tma_prefetch_psi(const StepArgs& a, const unsigned info, const int x0, const int y, const int z)
{
    const CompArgs& ca = a.c[C];
#pragma unroll
    for(int part = 0; part < 2; ++part)
    {
        const unsigned fs = part == 0 ? F_PS0 : F_PS1;
        if(!(info & fs)) continue;
        constexpr int AX0 = (C + 1) % 3, AX1 = (C + 2) % 3;
        const int axis = part == 0 ? AX0 : AX1;
        const PmlArgs& pp = ca.pml[part];
        const int map = TMAP_PSI0 + 2 * ((IS_E ? 0 : 3) + C) + part;
        if(axis == 1)      { const int cm = pp.cmap[y]; if(cm >= 0) tma_prefetch_row(a.tmaps, map, x0, z, cm); }
        else if(axis == 2) { const int cm = pp.cmap[z]; if(cm >= 0) tma_prefetch_row(a.tmaps, map, x0, cm, y); }
        else               tma_prefetch_row(a.tmaps, map, 0, z + a.lz * y, 0);
    }
}

// One component per thread (k_uniform: the block's component, blockIdx.x % 3; k_uniform_rows / k_general<E>: threadIdx.z): each
// thread issues the <= 8 independent loads of ITS component at once -- own value, the two driving arrays at the cell, their two
// stencil neighbours, D, psi -- so a plane costs one memory latency instead of one per component.  The driving arrays are shared
// between components; the other readers hit L1 / L2.  y-coupled neighbours are carried in registers while the block marches
// along y (see k_fast).
template <bool IS_E, int MODE, int C>
__device__ __forceinline__ void comp_march_init(const StepArgs& a, const long r, const long plane, double2& carry)
{
    // component 0 reads other[2] one plane away in y (grid_k), component 2 reads other[0] (grid_j); component 1 has no y neighbour
    carry = make_double2(0.0, 0.0);
    constexpr int SRC = C == 0 ? 2 : 0;
    if(C != 1 && has_other<IS_E, MODE>(SRC)) carry = *reinterpret_cast<const double2*>(a.fam[SRC] + (IS_E ? r - plane : r));
}
template <bool IS_E, int MODE, int C>
__device__ __forceinline__ void comp_march_load(const StepArgs& a, const long r, const long plane, double2& carry, PairLoads<IS_E, MODE>& L, const bool needU = true)
{
    constexpr int S = IS_E ? -1 : 1;
    constexpr int J = (C + 1) % 3, K = (C + 2) % 3;           // grid_j = other[J] (neighbour along axis K), grid_k = other[K] (along axis J)
    double* __restrict__ fj = const_cast<double*>(a.fam[J]);      // not const: see k_fast
    double* __restrict__ fk = const_cast<double*>(a.fam[K]);
    L.u[C] = make_double2(0.0, 0.0);
    if(needU) L.u[C] = *reinterpret_cast<const double2*>(a.c[C].U + r);   // not needed where D->E overwrites E
    L.v[J] = L.v[K] = L.nj[C] = L.nk[C] = make_double2(0.0, 0.0);
    // which of the two driving arrays is the y-coupled one: axis K == 1 -> grid_j (C == 2); axis J == 1 -> grid_k (C == 0)
    constexpr bool JY = K == 1, KY = J == 1;
    double2 nextPlane = make_double2(0.0, 0.0);
    if(has_other<IS_E, MODE>(J))
    {
        if(JY && !IS_E) { L.v[J] = carry; nextPlane = *reinterpret_cast<const double2*>(fj + r + plane); }
        else L.v[J] = *reinterpret_cast<const double2*>(fj + r);
    }
    if(has_other<IS_E, MODE>(K))
    {
        if(KY && !IS_E) { L.v[K] = carry; nextPlane = *reinterpret_cast<const double2*>(fk + r + plane); }
        else L.v[K] = *reinterpret_cast<const double2*>(fk + r);
    }
    if(has_other<IS_E, MODE>(J)) L.nj[C] = JY ? (IS_E ? carry : nextPlane) : neighbour2<K, S>(fj, r, a.px, plane, L.v[J]);
    if(has_other<IS_E, MODE>(K)) L.nk[C] = KY ? (IS_E ? carry : nextPlane) : neighbour2<J, S>(fk, r, a.px, plane, L.v[K]);
    if(JY && has_other<IS_E, MODE>(J)) carry = IS_E ? L.v[J] : nextPlane;
    if(KY && has_other<IS_E, MODE>(K)) carry = IS_E ? L.v[K] : nextPlane;
}

// a column of planes of a single-rectangle tile with the flag byte known at compile time: masks, prefactors and the plane-independent
// CPML coefficients are set up once, the plane loop holds only loads, the reference's arithmetic and stores
template <bool IS_E, int MODE, int C, unsigned FL, bool POLES>
__device__ __forceinline__ void uniform_column(const StepArgs& a, const TileRec& t, const unsigned rect, const unsigned info, const double2 pfc, const double ie,
                                               const int np, const int xl, const int zl, const int x, const int z)
{
    bool m0, m1;
    rect_mask(rect, xl, zl, m0, m1);
    if(!(m0 || m1)) return;
    PmlCoef kc;
    pml_coef_static<IS_E, MODE, C>(a, FL, x, z, kc);
    const long plane = a.px * a.lz;
    long r = x + a.px * (z + (long)a.lz * t.y);
    double2 carry;
    comp_march_init<IS_E, MODE, C>(a, r, plane, carry);
    constexpr bool needU = !IS_E || !(FL & (F_D2E | F_ORD2E)) || POLES;      // poles are driven by E^n
    const bool anyD = IS_E && a.c[C].D && ((FL & (F_ISD | F_D2E | F_ORD2E)) || (a.pml_on_D && (FL & (F_PG0 | F_PS0 | F_PG1 | F_PS1))));
    const bool leader = (threadIdx.x & 1) == 0;      // one lane per 32-byte sector (the prefetch unit of L2)
    // the first live lane of this lane's tile row (live = updates a cell: the others have left above)
    const unsigned liveLanes = __activemask() & ((t.part & REC_WIDE2) ? (threadIdx.x < 16 ? 0x0000FFFFu : 0xFFFF0000u) : 0xFFFFFFFFu);
    const bool rowLeader = (int)(threadIdx.x & 31) == __ffs(liveLanes) - 1;
    const int ny = t.ny, y0 = t.y;
    // psi of the y-normal slabs is stored under a compact y coordinate: it is looked up one plane ahead (CHIML_PIPE_Y)
    constexpr int YPART = C == 0 ? 0 : 1;
    constexpr bool YPSI = CHIML_PIPE_Y && C != 1 && ((FL & (YPART == 0 ? F_PS0 : F_PS1)) != 0);
    int cmNext = -2;
    if constexpr(YPSI) cmNext = a.c[C].pml[YPART].cmap[y0];
    for(int iy = 0; iy < ny; ++iy, r += plane)
    {
        const int y = y0 + iy;
        const int cmCur = cmNext;
        if constexpr(YPSI) if(iy + 1 < ny) cmNext = a.c[C].pml[YPART].cmap[y + 1];
        if(MODE == CHIML_MODE_3D && !POLES && a.tmaps != nullptr)
        {
            // one lane per tile row (a warp of a REC_WIDE2 record covers two rows: one lane per half-warp) asks the TMA unit for the
            // whole next row of every array the column streams
            if(rowLeader && iy + PREFETCH_PLANES < ny)
            {
                const int yp = y + PREFETCH_PLANES, x0 = t.x0;
                constexpr int FB = IS_E ? CHIML_HX : CHIML_EX, OB = IS_E ? CHIML_EX : CHIML_HX;
                if(has_other<IS_E, MODE>((C + 1) % 3)) tma_prefetch_row(a.tmaps, FB + (C + 1) % 3, x0, z, yp);
                if(has_other<IS_E, MODE>((C + 2) % 3)) tma_prefetch_row(a.tmaps, FB + (C + 2) % 3, x0, z, yp);
                if(needU) tma_prefetch_row(a.tmaps, OB + C, x0, z, yp);
                if(anyD) tma_prefetch_row(a.tmaps, CHIML_DX + C, x0, z, yp);
                tma_prefetch_psi<IS_E, MODE, C>(a, FL, x0, yp, z);
            }
        }
        else if(leader && iy + PREFETCH_PLANES < ny)
        {
            const long rp = r + PREFETCH_PLANES * plane;
            if(has_other<IS_E, MODE>((C + 1) % 3)) prefetch_l2(a.fam[(C + 1) % 3] + rp);
            if(has_other<IS_E, MODE>((C + 2) % 3)) prefetch_l2(a.fam[(C + 2) % 3] + rp);
            if(needU) prefetch_l2(a.c[C].U + rp);
            if(anyD) prefetch_l2(a.c[C].D + rp);
            prefetch_psi<IS_E, MODE, C>(a, FL, x, y + PREFETCH_PLANES, z);
            if constexpr(POLES)
            {
                // the pole pools are read pole after pole (two poles' loads in flight at a time): have them wait in L2
                const long rowp = z + (long)a.lz * (y + PREFETCH_PLANES);
                const long ipp = a.c[C].sp_base[rowp] + (x - a.c[C].sp_xmin[rowp]);
                for(int p = 0; p < np; ++p) { prefetch_l2(a.c[C].Pcur[p] + ipp); prefetch_l2(a.c[C].Pnew[p] + ipp); }
            }
        }
        PairLoads<IS_E, MODE> L;
        comp_march_load<IS_E, MODE, C>(a, r, plane, carry, L, needU);
        uniform_rect<IS_E, MODE, C, FL, true, POLES ? 1 : 0>(a, 0u, info, pfc, ie, L, r, z + (long)a.lz * y, x, y, z, xl, zl, np, m0, m1, &kc, YPSI ? cmCur : -2);
    }
}

template <bool IS_E, int MODE, int C>
__device__ __forceinline__ void uniform_march(const StepArgs& a, const TileRec& t, const int xl, const int zl, const int x, const int z)
{
    if(!has_own<IS_E, MODE>(C) || (t.rect[C] == 0 && t.rectB[C] == 0)) return;
    if constexpr(has_own<IS_E, MODE>(C) && has_other<IS_E, MODE>((C + 1) % 3) && has_other<IS_E, MODE>((C + 2) % 3))
    {
        // a warp is one z row of the tile.  When all its cells lie in ONE of the tile's rectangles (every warp of a one-rectangle tile;
        // in a tile cut along z every warp) it takes the column path of that rectangle; a row cut along x stays on the generic path
        // (letting the two halves of a warp run two column bodies one after the other costs more than it saves: r2b, profiles/README.md)
        unsigned rect = t.rect[C], info = t.info[C];
        double2 pfc = t.pf[C];
        double ie = t.inv_eps[C];
        int np = (int)((t.np >> (8 * C)) & 0xFFu);
        bool single = t.rectB[C] == 0;
        if(!single)
        {
            bool a0, a1, b0, b1;
            rect_mask(t.rect[C], xl, zl, a0, a1);
            rect_mask(t.rectB[C], xl, zl, b0, b1);
            const unsigned lanes = __activemask();
            const bool inA = a0 || a1, inB = b0 || b1;
            if(__all_sync(lanes, !inB)) { if(!inA) return; single = true; }
            else if(__all_sync(lanes, !inA)) { if(!inB) return; rect = t.rectB[C]; info = t.infoB[C]; pfc = t.pfB[C]; ie = t.inv_epsB[C]; np = (int)((t.npB >> (8 * C)) & 0xFFu); single = true; }
        }
        if(single)
        {
#define CHIML_COL(F) case (F): \
                if constexpr(IS_E && (((F) & F_D2E) != 0)) { if(np > 0) { uniform_column<IS_E, MODE, C, (F), true>(a, t, rect, info, pfc, ie, np, xl, zl, x, z); return; } } \
                uniform_column<IS_E, MODE, C, (F), false>(a, t, rect, info, pfc, ie, np, xl, zl, x, z); return;
            switch(info & 0xFF00u)
            {
                CHIML_COL(F_CURL)
                CHIML_COL(F_PG0 | F_PG1 | F_D2E)
                CHIML_COL(F_PG0 | F_PS0 | F_PG1 | F_D2E)
                CHIML_COL(F_PG0 | F_PG1 | F_PS1 | F_D2E)
                CHIML_COL(F_PG0 | F_PS0 | F_PG1 | F_PS1 | F_D2E)
                CHIML_COL(F_CURL | F_ISD | F_D2E)
                CHIML_COL(F_CURL | F_ISD | F_ORD2E)
                CHIML_COL(F_PG0 | F_PG1 | F_ORD2E)
                CHIML_COL(F_PG0 | F_PS0 | F_PG1 | F_ORD2E)
                CHIML_COL(F_PG0 | F_PG1 | F_PS1 | F_ORD2E)
                CHIML_COL(F_PG0 | F_PS0 | F_PG1 | F_PS1 | F_ORD2E)
                CHIML_COL(F_PG0 | F_PG1)
                CHIML_COL(F_PG0 | F_PS0 | F_PG1)
                CHIML_COL(F_PG0 | F_PG1 | F_PS1)
                CHIML_COL(F_PG0 | F_PS0 | F_PG1 | F_PS1)
                default: break;
            }
#undef CHIML_COL
        }
    }
    const long plane = a.px * a.lz;
    long r = x + a.px * (z + (long)a.lz * t.y);
    double2 carry;
    comp_march_init<IS_E, MODE, C>(a, r, plane, carry);
    const bool needU = !IS_E || (t.rect[C] != 0 && !(t.info[C] & (F_D2E | F_ORD2E))) || (t.rectB[C] != 0 && !(t.infoB[C] & (F_D2E | F_ORD2E))) ||
                       (((t.np | t.npB) >> (8 * C)) & 0xFFu) != 0;
    const unsigned anyInfo = (t.rect[C] ? t.info[C] : 0u) | (t.rectB[C] ? t.infoB[C] : 0u);
    const bool anyD = IS_E && a.c[C].D && ((anyInfo & (F_ISD | F_D2E | F_ORD2E)) || (a.pml_on_D && (anyInfo & (F_PG0 | F_PS0 | F_PG1 | F_PS1))));
    const bool leader = (threadIdx.x & 1) == 0;      // one lane per 32-byte sector (the prefetch unit of L2)
    for(int iy = 0; iy < t.ny; ++iy, r += plane)
    {
        const int y = t.y + iy;
        if(leader && iy + PREFETCH_PLANES < t.ny)
        {
            const long rp = r + PREFETCH_PLANES * plane;
            if(has_other<IS_E, MODE>((C + 1) % 3)) prefetch_l2(a.fam[(C + 1) % 3] + rp);
            if(has_other<IS_E, MODE>((C + 2) % 3)) prefetch_l2(a.fam[(C + 2) % 3] + rp);
            if(needU) prefetch_l2(a.c[C].U + rp);
            if(anyD) prefetch_l2(a.c[C].D + rp);
            if(t.rect[C])  prefetch_psi<IS_E, MODE, C>(a, t.info[C], x, y + PREFETCH_PLANES, z);
            if(t.rectB[C]) prefetch_psi<IS_E, MODE, C>(a, t.infoB[C], x, y + PREFETCH_PLANES, z);
        }
        PairLoads<IS_E, MODE> L;
        comp_march_load<IS_E, MODE, C>(a, r, plane, carry, L, needU);
        uniform_comp<IS_E, MODE, C>(a, t, L, r, z + (long)a.lz * y, x, y, z, xl, zl);
    }
}

// One component per thread for both half steps: with the flag-specialised column bodies the split wins for H as well (2.33 vs
// 2.78 ms, profiles/README.md), although the three component threads re-read the shared driving arrays through L1.
template <bool IS_E, int MODE>
__device__ __forceinline__ void uniform_body(const StepArgs& a, const TileRec& t, const int xl, const int zl, const int comp)
{
    const int x = t.x0 + xl, z = t.z0 + zl;
    if(x >= a.px || z >= a.lz) return;
    if(comp == 0)      uniform_march<IS_E, MODE, 0>(a, t, xl, zl, x, z);
    else if(comp == 1) uniform_march<IS_E, MODE, 1>(a, t, xl, zl, x, z);
    else               uniform_march<IS_E, MODE, 2>(a, t, xl, zl, x, z);
}

// One component per BLOCK: blockIdx.x % 3 selects it, so a block runs a single code body (the flag-specialised column of its
// component) and the three blocks of a tile are neighbours in the grid -- they run at about the same time on different SMs and
// share the driving arrays through L2.  Against one component per thread z-index inside one block (12 warps marching in
// lock-step through three different bodies) this is 8-15 % faster: smaller independent blocks drift apart and fill each other's
// memory-latency gaps, and each scheduler holds one body instead of three.  A block owns TILE_Z / split z rows of the tile.
template <bool IS_E, int MODE>
__global__ void __launch_bounds__(32 * TILE_Z / uniform_split<IS_E>(), uniform_occ<IS_E>()) k_uniform(const __grid_constant__ StepArgs a, const TileRec* __restrict__ tiles)
{
    constexpr int SPLIT = uniform_split<IS_E>(), ROWS = TILE_Z / SPLIT;
    const unsigned b = blockIdx.x / 3;
    const TileRec& t = tiles[b / SPLIT];
    int xl = 2 * threadIdx.x, zl = threadIdx.y + ROWS * (b % SPLIT);
    // narrow records (CPML faces normal to x: 20 of the 64 cells of a row): two z-adjacent tiles in one record, half a warp per row,
    // so that 10 of 16 lanes work instead of 10 of 32 (the x-face columns ran at 3.5 TB/s against 5.2 for the other faces)
    if(SPLIT == 1 && (t.part & REC_WIDE2)) { xl = (int)t.pad4 + 2 * (threadIdx.x & 15); zl = 2 * threadIdx.y + (threadIdx.x >> 4); }
    uniform_body<IS_E, MODE>(a, t, xl, zl, blockIdx.x % 3);
}
// 2-D grids: a tile is one row of 64 cells
template <bool IS_E, int MODE>
__global__ void __launch_bounds__(32 * ROWS_2D * 3) k_uniform_rows(const __grid_constant__ StepArgs a, const TileRec* __restrict__ tiles, const unsigned ntiles)
{
    const unsigned ti = blockIdx.x * ROWS_2D + threadIdx.y;
    if(ti < ntiles) uniform_body<IS_E, MODE>(a, tiles[ti], 2 * threadIdx.x, 0, threadIdx.z);
}

// ---------------------------------------------------------------------------------------------------
// k_general: everything else, per cell
// ---------------------------------------------------------------------------------------------------
template <bool IS_E, int MODE, int C, bool DP = IS_E>
__device__ __forceinline__ void general_pair(const StepArgs& a, double2 u, const double2 vj, const double2 nj, const double2 vk, const double2 nk,
                                             const long r, const long row, const int x, const int y, const int z)
{
    if(!has_own<IS_E, MODE>(C)) return;
    const CompArgs& ca = a.c[C];
    const ushort2 info = *reinterpret_cast<const ushort2*>(ca.info + r);
    const bool w0 = update_cell<IS_E, MODE, C, DP>(a, ca, info.x, u.x, vj.x, nj.x, vk.x, nk.x, r, row, x, y, z);
    const bool w1 = update_cell<IS_E, MODE, C, DP>(a, ca, info.y, u.y, vj.y, nj.y, vk.y, nk.y, r + 1, row, x + 1, y, z);
    store_pair(ca.U + r, u, w0, w1);
}

// one component per thread (threadIdx.z): the loads of a component do not queue behind the arithmetic of another
template <bool IS_E, int MODE, int C, bool DP = IS_E>
__device__ __forceinline__ void general_comp(const StepArgs& a, const TileRec& t, const int x, const int z)
{
    if(!has_own<IS_E, MODE>(C)) return;
    const long plane = a.px * a.lz;
    const int y = t.y;
    const long row = z + (long)a.lz * y;
    const long r = x + a.px * row;
    double2 carry;
    comp_march_init<IS_E, MODE, C>(a, r, plane, carry);
    PairLoads<IS_E, MODE> L;
    comp_march_load<IS_E, MODE, C>(a, r, plane, carry, L);
    general_pair<IS_E, MODE, C, DP>(a, L.u[C], L.v[(C + 1) % 3], L.nj[C], L.v[(C + 2) % 3], L.nk[C], r, row, x, y, z);
}

template <bool IS_E, int MODE, bool SPLIT, bool DP = IS_E>
__global__ void __launch_bounds__(SPLIT ? 768 : 256, SPLIT ? 1 : 2) k_general(const __grid_constant__ StepArgs a, const TileRec* __restrict__ tiles, const unsigned ntiles)
{
    unsigned ti; int zl;
    if(!tile_of_thread<MODE>(ntiles, ti, zl)) return;
    const TileRec& t = tiles[ti];
    const int x = t.x0 + 2 * threadIdx.x, z = t.z0 + zl;
    if(x >= a.px || z >= a.lz) return;
    if(!SPLIT)
    {
        const int y = t.y;
        const long row = z + (long)a.lz * y;
        const long r = x + a.px * row;
        PairLoads<IS_E, MODE> L;
        L.load(a, r);
        general_pair<IS_E, MODE, 0, DP>(a, L.u[0], L.v[1], L.nj[0], L.v[2], L.nk[0], r, row, x, y, z);
        general_pair<IS_E, MODE, 1, DP>(a, L.u[1], L.v[2], L.nj[1], L.v[0], L.nk[1], r, row, x, y, z);
        general_pair<IS_E, MODE, 2, DP>(a, L.u[2], L.v[0], L.nj[2], L.v[1], L.nk[2], r, row, x, y, z);
        return;
    }
    if(threadIdx.z == 0)      general_comp<IS_E, MODE, 0, DP>(a, t, x, z);
    else if(threadIdx.z == 1) general_comp<IS_E, MODE, 1, DP>(a, t, x, z);
    else                      general_comp<IS_E, MODE, 2, DP>(a, t, x, z);
}

// ---------------------------------------------------------------------------------------------------
// commit-time tile summary: per tile and component the bounding rectangle of non-zero info cells, their
// count, the first non-zero info value and whether all non-zero values are equal.  One block per tile.
// ---------------------------------------------------------------------------------------------------
// TileSummary: chiml_tiles.hpp

// algorithmic bytes one field-component cell moves per step: 16 B RW + 8 B cross-read when it is
// updated, +16 B when its D value is read and written, +24 B per isotropic pole, +16 B per psi value touched
__device__ __forceinline__ unsigned cell_alg_bytes(const unsigned info, const ClassEntry* cls, const bool isE, const bool pmlOnD)
{
    if(info == 0) return 0;
    const bool pml = (info & (F_PG0 | F_PS0 | F_PG1 | F_PS1)) != 0;
    unsigned b = 0;
    if((info & F_CURL) || (info & (F_PG0 | F_PG1))) b += 24;
    if(isE && ((info & (F_ISD | F_D2E | F_ORD2E)) || (pmlOnD && pml))) b += 16;
    if(isE && (info & F_D2E)) b += 24u * (unsigned)cls[info & CLS_MASK].npoles;
    if(info & F_PS0) b += 16;
    if(info & F_PS1) b += 16;
    return b;
}

__global__ void k_tile_summary(const uint16_t* i0, const uint16_t* i1, const uint16_t* i2, TileSummary* out,
                               unsigned nxt, unsigned nzt, int lz, long px,
                               const ClassEntry* c0, const ClassEntry* c1, const ClassEntry* c2, int isE, int pmlOnD)
{
    __shared__ unsigned s_bytes;
    const ClassEntry* cp[3] = {c0, c1, c2};
    if(threadIdx.x == 0 && threadIdx.y == 0) s_bytes = 0;
    const unsigned tile = blockIdx.x;
    const unsigned xt = tile % nxt, zt = (tile / nxt) % nzt, y = tile / (nxt * nzt);
    const int xl = 2 * threadIdx.x, zl = threadIdx.y;
    const int x = xt * TILE_X + xl, z = zt * blockDim.y + zl;
    __shared__ unsigned s_lo[TS_NV][3][2], s_hi[TS_NV][3][2], s_cnt[TS_NV][3], s_val[TS_NV][3], s_tot[3], s_other[3];
    const uint16_t* ip[3] = {i0, i1, i2};
    if(threadIdx.x < 3 && threadIdx.y == 0)
    {
        const int c = threadIdx.x;
        for(int k = 0; k < TS_NV; ++k) { s_lo[k][c][0] = s_lo[k][c][1] = 255; s_hi[k][c][0] = s_hi[k][c][1] = 0; s_cnt[k][c] = 0; s_val[k][c] = 0; }
        s_tot[c] = 0; s_other[c] = 0;
    }
    __syncthreads();
    unsigned v[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    if(z < lz && x < px)
    {
        const long r = x + px * (z + (long)lz * y);
        for(int c = 0; c < 3; ++c)
        {
            if(!ip[c]) continue;
            v[c][0] = ip[c][r]; v[c][1] = ip[c][r + 1];
            for(int k = 0; k < 2; ++k)
                if(v[c][k])
                {
                    atomicAdd(&s_bytes, cell_alg_bytes(v[c][k], cp[c], isE != 0, pmlOnD != 0));
                    atomicAdd(&s_tot[c], 1u);
                }
        }
    }
    // distinct values, largest first: value w is the maximum of the values below value w-1
    for(int w = 0; w < TS_NV; ++w)
    {
        for(int c = 0; c < 3; ++c)
            for(int k = 0; k < 2; ++k)
                if(v[c][k] && (w == 0 || v[c][k] < s_val[w - 1][c])) atomicMax(&s_val[w][c], v[c][k]);
        __syncthreads();
        for(int c = 0; c < 3; ++c)
            for(int k = 0; k < 2; ++k)
            {
                if(!v[c][k] || v[c][k] != s_val[w][c]) continue;
                atomicMin(&s_lo[w][c][0], (unsigned)(xl + k)); atomicMax(&s_hi[w][c][0], (unsigned)(xl + k + 1));
                atomicMin(&s_lo[w][c][1], (unsigned)zl);       atomicMax(&s_hi[w][c][1], (unsigned)(zl + 1));
                atomicAdd(&s_cnt[w][c], 1u);
            }
        __syncthreads();
    }
    for(int c = 0; c < 3; ++c)
        for(int k = 0; k < 2; ++k)
            if(v[c][k] && v[c][k] < s_val[TS_NV - 1][c]) s_other[c] = 1;
    __syncthreads();
    if(threadIdx.x < 3 && threadIdx.y == 0)
    {
        const int c = threadIdx.x;
        TileSummary& o = out[tile];
        o.total[c] = s_tot[c]; o.other[c] = s_other[c];
        for(int w = 0; w < TS_NV; ++w)
        {
            o.count[c][w] = s_cnt[w][c]; o.info[c][w] = s_cnt[w][c] ? s_val[w][c] : 0u;
            o.rect[c][w] = s_cnt[w][c] ? (s_lo[w][c][0] | (s_hi[w][c][0] << 8) | (s_lo[w][c][1] << 16) | (s_hi[w][c][1] << 24)) : 0u;
        }
        if(c == 0) o.bytes = s_bytes;
    }
}

} // namespace chiml
