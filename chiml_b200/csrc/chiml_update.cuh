// The fused half-step kernel (H half step: updateH + updateHPML_; E half step: updatePolE (isotropic),
// updateD, updateE, updateEPML_, D2E) -- reference FDTD_MANAGER/parallelFDTDField.hpp:1228-1303.
// Included by chiml_kernels.cuh.
#pragma once

namespace chiml {

constexpr int TILE_X = 64;   // cells per tile row (32 lanes x 2 cells)
constexpr int TILE_Z = 8;    // rows per tile (3-D); 2-D grids use 1

// tile descriptor (one uint32 per tile, per family, per plane): status in the top byte, the
// uniform class of each component in the three low bytes
constexpr unsigned TD_GENERAL = 0u, TD_UNIFORM = 1u, TD_EMPTY = 2u;

// which components exist (FDTD_MANAGER/parallelFDTDField.hpp:391-443): TE = Ex,Ey,Hz; TM = Ez,Hx,Hy
template <int MODE> __host__ __device__ constexpr bool has_E(int c) { return MODE == CHIML_MODE_3D || (MODE == CHIML_MODE_TE ? c != 2 : c == 2); }
template <int MODE> __host__ __device__ constexpr bool has_H(int c) { return MODE == CHIML_MODE_3D || (MODE == CHIML_MODE_TE ? c == 2 : c != 2); }
template <bool IS_E, int MODE> __host__ __device__ constexpr bool has_own(int c) { return IS_E ? has_E<MODE>(c) : has_H<MODE>(c); }
template <bool IS_E, int MODE> __host__ __device__ constexpr bool has_other(int c) { return IS_E ? has_H<MODE>(c) : has_E<MODE>(c); }

// General path: one field component C of one cell.  u = current value of U[r]; (vj_r, vj_n) = grid_j at
// ind and ind_k, (vk_r, vk_n) = grid_k at ind and ind_j -- the four stencil values of TwoCompCurl, which
// are also the stencil values of the two CPML parts (part 0: grid_k, derivative along j = (C+1)%3;
// part 1: grid_j, derivative along k = (C+2)%3).  Returns true when U[r] must be written back.
template <bool IS_E, int MODE, int C>
__device__ __forceinline__ bool update_cell(const StepArgs& a, const CompArgs& ca, const unsigned info, double& u,
                                            const double vj_r, const double vj_n, const double vk_r, const double vk_n,
                                            const long r, const long row, const int x, const int y, const int z)
{
    if(info == 0) return false;
    const ClassEntry& ce = ca.cls[info & CLS_MASK];
    constexpr bool HAS_VJ = has_other<IS_E, MODE>((C + 1) % 3);
    constexpr bool HAS_VK = has_other<IS_E, MODE>((C + 2) % 3);

    double pn[MAX_POLES];
    int np = 0;

    // updatePolE, isotropic poles (parallelFDTDField.hpp:1355-1361 -> UTIL/FDTD_up_eq.cpp:435-446):
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

    if(dDirty) ca.D[r] = dv;
    return true;
}

// Values of V at the stencil neighbour of cells (x, x+1): one cell along AXIS in direction SIGN.
// own2 = V[r], V[r+1].  Row pitch and plane stride are multiples of 16 doubles, so the y / z
// neighbours are aligned 16-byte loads; the x neighbour needs one extra scalar.
template <int AXIS, int SIGN>
__device__ __forceinline__ double2 neighbour2(const double* __restrict__ V, const long r, const long px, const long plane, const double2 own2)
{
    if(AXIS == 0) return SIGN > 0 ? make_double2(own2.y, V[r + 2]) : make_double2(V[r - 1], own2.x);
    const long off = (AXIS == 2 ? px : plane) * SIGN;
    return *reinterpret_cast<const double2*>(V + r + off);
}

template <bool IS_E, int MODE, int C>
__device__ __forceinline__ void general_pair(const StepArgs& a, double2 u, const double2 vj, const double2 nj, const double2 vk, const double2 nk,
                                             const long r, const long row, const int x, const int y, const int z)
{
    if(!has_own<IS_E, MODE>(C)) return;
    const CompArgs& ca = a.c[C];
    const ushort2 info = *reinterpret_cast<const ushort2*>(ca.info + r);
    const bool w0 = update_cell<IS_E, MODE, C>(a, ca, info.x, u.x, vj.x, nj.x, vk.x, nk.x, r, row, x, y, z);
    const bool w1 = update_cell<IS_E, MODE, C>(a, ca, info.y, u.y, vj.y, nj.y, vk.y, nk.y, r + 1, row, x + 1, y, z);
    if(w0 && w1) *reinterpret_cast<double2*>(ca.U + r) = u;
    else if(w0) ca.U[r] = u.x;
    else if(w1) ca.U[r + 1] = u.y;
}

// TwoCompCurl / OneCompCurlJ / OneCompCurlK on both cells of the pair (UTIL/FDTD_up_eq.cpp:10-35)
template <bool IS_E, int MODE, int C>
__device__ __forceinline__ void fast_pair(const StepArgs& a, const unsigned cls, double2 t, const double2 vj, const double2 nj, const double2 vk, const double2 nk, const long r)
{
    if(!has_own<IS_E, MODE>(C)) return;
    const CompArgs& ca = a.c[C];
    const double2 pf = ca.pf[cls];          // {pf1, pf2}
    if(has_other<IS_E, MODE>((C + 1) % 3))
    {
        t.x = axpy1(t.x,  pf.y, vj.x); t.y = axpy1(t.y,  pf.y, vj.y);
        t.x = axpy1(t.x, -pf.y, nj.x); t.y = axpy1(t.y, -pf.y, nj.y);
    }
    if(has_other<IS_E, MODE>((C + 2) % 3))
    {
        t.x = axpy1(t.x, -pf.x, vk.x); t.y = axpy1(t.y, -pf.x, vk.y);
        t.x = axpy1(t.x,  pf.x, nk.x); t.y = axpy1(t.y,  pf.x, nk.y);
    }
    *reinterpret_cast<double2*>(ca.U + r) = t;
}

// Fused half step, y-marching.  A block owns one (x, z) tile column (TILE_X x TILE_Z cells) and marches
// through a chunk of y planes; each thread owns two x-adjacent cells and all components at them.
//   * Register pipelining: while plane y is being computed, every load of plane y+1 (own fields, the
//     other family, its x and z stencil neighbours) is already in flight, so HBM latency is hidden by
//     design rather than by occupancy.  The y stencil neighbour never touches memory again: the E
//     half step keeps the previous plane of H in registers, the H half step reads E one plane ahead.
//     Each field value therefore crosses HBM once per half step (plus a one-row z halo per tile).
//   * A per-tile, per-plane descriptor built at commit time says whether every cell of the tile is a
//     plain interior curl cell of one material class: such tiles (the bulk of the domain) never read
//     the cell-info planes and run straight-line code; CPML / dispersive / boundary tiles take the
//     general path (update_cell).
// The stencil is the reference's (derivOff of FDTD_MANAGER/parallelFDTDField.cpp:80-82,248-250):
// component c reads grid_j = other[(c+1)%3] one cell along axis (c+2)%3 and grid_k = other[(c+2)%3] one
// cell along axis (c+1)%3, backwards for E, forwards for H; chiml_gpu_commit verifies the lists agree.
// Field arrays carry one plane of zeroed slack on either side, so the look-ahead loads and the
// neighbour loads of ghost / padding cells stay in bounds.
template <bool IS_E, int MODE>
struct Stage
{
    double2 u[3];      // own family at the two cells
    double2 v[3];      // other family at the two cells (E half step only; the H half step rotates v separately)
    double  xs[2];     // the extra scalar of the two x-neighbour pairs: other[2] (for c=1), other[1] (for c=2)
    double2 zn[2];     // the two z neighbours: other[1] (for c=0), other[0] (for c=1)
};

template <bool IS_E, int MODE>
__device__ __forceinline__ void load_other(const StepArgs& a, const long r, double2 (&v)[3])
{
#pragma unroll
    for(int c = 0; c < 3; ++c)
        if(has_other<IS_E, MODE>(c)) v[c] = *reinterpret_cast<const double2*>(a.fam[c] + r);
}

template <bool IS_E, int MODE>
__device__ __forceinline__ void load_stage(const StepArgs& a, const long r, Stage<IS_E, MODE>& st)
{
    constexpr int S = IS_E ? -1 : 1;
    constexpr bool IS3D = MODE == CHIML_MODE_3D;
#pragma unroll
    for(int c = 0; c < 3; ++c)
        if(has_own<IS_E, MODE>(c)) st.u[c] = *reinterpret_cast<const double2*>(a.c[c].U + r);
    if(IS_E) load_other<IS_E, MODE>(a, r, st.v);
    const long xo = S > 0 ? 2 : -1;
    if(has_own<IS_E, MODE>(1) && has_other<IS_E, MODE>(2)) st.xs[0] = a.fam[2][r + xo];
    if(has_own<IS_E, MODE>(2) && has_other<IS_E, MODE>(1)) st.xs[1] = a.fam[1][r + xo];
    if(IS3D)
    {
        st.zn[0] = *reinterpret_cast<const double2*>(a.fam[1] + r + S * a.px);
        st.zn[1] = *reinterpret_cast<const double2*>(a.fam[0] + r + S * a.px);
    }
}

template <bool IS_E, int MODE>
__global__ void __launch_bounds__(256, 2) k_update(const __grid_constant__ StepArgs a)
{
    constexpr int S = IS_E ? -1 : 1;
    // block -> (x tile, z tile, y chunk); x tiles fastest so that concurrently running blocks stream neighbouring rows
    unsigned b = blockIdx.x;
    const unsigned xt = b % a.nxt;  b /= a.nxt;
    const unsigned zt = b % a.nzt;
    const int y0 = (int)(b / a.nzt) * a.ychunk;
    const int y1 = min(y0 + a.ychunk, a.ly);
    const int x = 2 * (xt * 32 + threadIdx.x);
    const int z = zt * blockDim.y + threadIdx.y;
    if(x >= a.px || z >= a.lz) return;
    const long plane = a.px * a.lz;
    long r = x + a.px * (z + (long)a.lz * y0);
    const unsigned* tdp = a.tiledesc + ((size_t)y0 * a.nzt + zt) * a.nxt + xt;
    const size_t tdStride = (size_t)a.nzt * a.nxt;

    Stage<IS_E, MODE> cur, nxt;
    double2 vy[3];     // E: other family one plane back (y-1); H: other family one plane ahead (y+1)
    double2 vy2[3];    // H: other family two planes ahead, in flight
#pragma unroll
    for(int c = 0; c < 3; ++c)
    {
        cur.u[c] = cur.v[c] = nxt.u[c] = nxt.v[c] = vy[c] = vy2[c] = make_double2(0.0, 0.0);
    }
    cur.xs[0] = cur.xs[1] = nxt.xs[0] = nxt.xs[1] = 0.0;
    cur.zn[0] = cur.zn[1] = nxt.zn[0] = nxt.zn[1] = make_double2(0.0, 0.0);

    // prologue
    load_stage<IS_E, MODE>(a, r, cur);
    if(IS_E) load_other<IS_E, MODE>(a, r - plane, vy);
    else { load_other<IS_E, MODE>(a, r, cur.v); load_other<IS_E, MODE>(a, r + plane, vy); }
    unsigned td = *tdp;

#pragma unroll 1
    for(int y = y0; y < y1; ++y)
    {
        // ---- issue every load of plane y+1 (and, for H, the other family of plane y+2) -------------------------
        const long rn = r + plane;
        unsigned tdn = TD_EMPTY << 24;
        if(y + 1 < y1)
        {
            load_stage<IS_E, MODE>(a, rn, nxt);
            if(!IS_E) load_other<IS_E, MODE>(a, rn + plane, vy2);
            tdn = tdp[tdStride];
        }
        // ---- compute plane y -----------------------------------------------------------------------------------
        const unsigned status = td >> 24;
        if(status != TD_EMPTY)
        {
            // component c: grid_j = other[(c+1)%3] along axis (c+2)%3 ; grid_k = other[(c+2)%3] along axis (c+1)%3
            const double2 nj0 = cur.zn[0];                                                     // other[1] along z
            const double2 nk0 = vy[2];                                                         // other[2] along y
            const double2 nj1 = S > 0 ? make_double2(cur.v[2].y, cur.xs[0]) : make_double2(cur.xs[0], cur.v[2].x);   // other[2] along x
            const double2 nk1 = cur.zn[1];                                                     // other[0] along z
            const double2 nj2 = vy[0];                                                         // other[0] along y
            const double2 nk2 = S > 0 ? make_double2(cur.v[1].y, cur.xs[1]) : make_double2(cur.xs[1], cur.v[1].x);   // other[1] along x
            if(status == TD_UNIFORM)
            {
                fast_pair<IS_E, MODE, 0>(a, td & 0xFFu,         cur.u[0], cur.v[1], nj0, cur.v[2], nk0, r);
                fast_pair<IS_E, MODE, 1>(a, (td >> 8) & 0xFFu,  cur.u[1], cur.v[2], nj1, cur.v[0], nk1, r);
                fast_pair<IS_E, MODE, 2>(a, (td >> 16) & 0xFFu, cur.u[2], cur.v[0], nj2, cur.v[1], nk2, r);
            }
            else
            {
                const long row = z + (long)a.lz * y;
                general_pair<IS_E, MODE, 0>(a, cur.u[0], cur.v[1], nj0, cur.v[2], nk0, r, row, x, y, z);
                general_pair<IS_E, MODE, 1>(a, cur.u[1], cur.v[2], nj1, cur.v[0], nk1, r, row, x, y, z);
                general_pair<IS_E, MODE, 2>(a, cur.u[2], cur.v[0], nj2, cur.v[1], nk2, r, row, x, y, z);
            }
        }
        // ---- rotate ----------------------------------------------------------------------------------------------
        if(IS_E)
        {
#pragma unroll
            for(int c = 0; c < 3; ++c) vy[c] = cur.v[c];
            cur = nxt;
        }
        else
        {
#pragma unroll
            for(int c = 0; c < 3; ++c) { nxt.v[c] = vy[c]; vy[c] = vy2[c]; }
            cur = nxt;
        }
        td = tdn;
        tdp += tdStride;
        r = rn;
    }
}

// Commit-time classification of the tiles of one family (one block per tile, same shape as k_update).
__global__ void k_tile_desc(const uint16_t* i0, const uint16_t* i1, const uint16_t* i2, unsigned* desc,
                            unsigned nxt, unsigned nzt, int lx, int lz, long px)
{
    const unsigned tile = blockIdx.x;
    const unsigned xt = tile % nxt, zt = (tile / nxt) % nzt, y = tile / (nxt * nzt);
    const int x = 2 * (xt * 32 + threadIdx.x);
    const int z = zt * blockDim.y + threadIdx.y;
    __shared__ unsigned first[3];
    __shared__ int mixed, nonzero;
    const uint16_t* ip[3] = {i0, i1, i2};
    if(threadIdx.x == 0 && threadIdx.y == 0)
    {
        mixed = 0; nonzero = 0;
        const long r0 = (long)(2 * xt * 32) + px * ((long)zt * blockDim.y + (long)lz * y);
        for(int c = 0; c < 3; ++c) first[c] = ip[c] ? ip[c][r0] : 0u;
    }
    __syncthreads();
    if(z < lz && x < px)
    {
        const long r = x + px * (z + (long)lz * y);
        for(int c = 0; c < 3; ++c)
        {
            if(!ip[c]) continue;
            const unsigned v0 = ip[c][r], v1 = ip[c][r + 1];
            if(v0 != first[c] || v1 != first[c]) mixed = 1;
            if(v0 | v1) nonzero = 1;
        }
    }
    else if(z < lz) mixed = 1;      // tile reaches past the padded row: cannot be uniform
    __syncthreads();
    if(threadIdx.x == 0 && threadIdx.y == 0)
    {
        unsigned d;
        bool simple = !mixed;
        for(int c = 0; c < 3; ++c)
            if(ip[c] && (first[c] & 0xFF00u) != F_CURL) simple = false;
        // a tile whose last rows lie beyond lz is still uniform: those threads exit in k_update
        if(!nonzero) d = TD_EMPTY << 24;
        else if(simple) d = (TD_UNIFORM << 24) | (first[0] & 0xFFu) | ((first[1] & 0xFFu) << 8) | ((first[2] & 0xFFu) << 16);
        else d = TD_GENERAL << 24;
        desc[tile] = d;
    }
}

} // namespace chiml
