// k_uniform_cells: the E half step of UNIFORM tiles, one component per thread, with the tile's one or two rectangles resolved ONCE per
// thread into "what each of my two cells does" (info, prefactors): the plane loop is then straight-line code per cell instead of a
// loop over rectangles with masks -- fewer instructions per plane than k_uniform<true, MODE, true> and the same loads.
#pragma once

namespace chiml {

struct PipeCell                    // what one cell of a UNIFORM tile needs besides the staged operands
{
    unsigned info; double pf1, pf2, inv_eps;
};

// one cell, the reference's operation order (curl; CPML part 0, part 1; D -> E), operands in registers
template <bool HAS_VJ, bool HAS_VK>
__device__ __forceinline__ void pipe_cell(const PipeCell& k, const bool pmlOnD, double& u, double& dv, bool& dDirty,
                                          const double vj, const double nj, const double vk, const double nk,
                                          double (&ps)[2], bool (&psDirty)[2], const double (&F)[2], const double (&b)[2], const double (&c)[2], const double (&Db)[2])
{
    const unsigned info = k.info;
    if(info & F_CURL)
    {
        double t = (info & F_ISD) ? dv : u;
        if(HAS_VJ) { t = axpy1(t,  k.pf2, vj); t = axpy1(t, -k.pf2, nj); }
        if(HAS_VK) { t = axpy1(t, -k.pf1, vk); t = axpy1(t,  k.pf1, nk); }
        if(info & F_ISD) { dv = t; dDirty = true; } else u = t;
    }
    if(info & (F_PG0 | F_PS0 | F_PG1 | F_PS1))
    {
        double t = pmlOnD ? dv : u;
#pragma unroll
        for(int part = 0; part < 2; ++part)
        {
            if(part == 0 ? !HAS_VK : !HAS_VJ) continue;
            const unsigned fg = part == 0 ? F_PG0 : F_PG1, fs = part == 0 ? F_PS0 : F_PS1;
            if(!(info & (fg | fs))) continue;
            const double vr = part == 0 ? vk : vj, vo = part == 0 ? nk : nj;
            double p = 0.0;
            if(info & fs)
            {
                p = dm(b[part], ps[part]);
                p = axpy1(p,  c[part], vr);
                p = axpy1(p, -c[part], vo);
                ps[part] = p; psDirty[part] = true;
            }
            if(info & fg)
            {
                t = axpy1(t,  F[part], vr);
                t = axpy1(t, -F[part], vo);
                if(info & fs) t = axpy1(t, Db[part], p);
            }
        }
        if(pmlOnD) { dv = t; dDirty = true; } else u = t;
    }
    if(info & F_D2E) u = dm(k.inv_eps, dv);      // DtoU without pole grids (UTIL/FDTD_up_eq.cpp:838-843)
}

template <int MODE, int C>
__device__ __forceinline__ void cell_comp(const StepArgs& a, const TileRec& t, const int xl, const int zl, const int x, const int z)
{
    constexpr bool IS_E = true;
    if constexpr(has_own<IS_E, MODE>(C))
    {
        if(t.rect[C] == 0 && t.rectB[C] == 0) return;
        constexpr int J = (C + 1) % 3, K = (C + 2) % 3;       // grid_j = H_J, neighbour along axis K; grid_k = H_K, neighbour along axis J
        constexpr bool HAS_VJ = has_other<IS_E, MODE>(J), HAS_VK = has_other<IS_E, MODE>(K);
        const CompArgs& ca = a.c[C];
        // which rectangle each of my two cells belongs to: decided once, the plane loop below is straight-line per cell
        bool a0 = false, a1 = false, b0 = false, b1 = false;
        if(t.rect[C])  rect_mask(t.rect[C], xl, zl, a0, a1);
        if(t.rectB[C]) rect_mask(t.rectB[C], xl, zl, b0, b1);
        const bool my0 = a0 || b0, my1 = a1 || b1;
        if(!(my0 || my1)) return;
        PipeCell k0, k1;
        k0.info = a0 ? t.info[C] : (b0 ? t.infoB[C] : 0u); k1.info = a1 ? t.info[C] : (b1 ? t.infoB[C] : 0u);
        k0.pf1 = a0 ? t.pf[C].x : t.pfB[C].x; k0.pf2 = a0 ? t.pf[C].y : t.pfB[C].y; k0.inv_eps = a0 ? t.inv_eps[C] : t.inv_epsB[C];
        k1.pf1 = a1 ? t.pf[C].x : t.pfB[C].x; k1.pf2 = a1 ? t.pf[C].y : t.pfB[C].y; k1.inv_eps = a1 ? t.inv_eps[C] : t.inv_epsB[C];
        const unsigned any = k0.info | k1.info;
        const bool pmlOnD = a.pml_on_D != 0;
        const bool anyPml = (any & (F_PG0 | F_PS0 | F_PG1 | F_PS1)) != 0;
        const bool needD = ca.D && ((any & (F_ISD | F_D2E)) || (pmlOnD && anyPml));
        const bool needU = (my0 && !(k0.info & F_D2E)) || (my1 && !(k1.info & F_D2E));
        const long plane = a.px * a.lz;
        long r = x + a.px * (z + (long)a.lz * t.y);
        const double* __restrict__ fj = a.fam[J];
        const double* __restrict__ fk = a.fam[K];
        // CPML coefficients: along x per cell and the same for every plane, along z one scalar for every plane, along y per plane
        constexpr int AX[2] = {J, K};                          // part 0 differentiates along J, part 1 along K
        bool hasPs[2], hasPg[2];
        double F0[2] = {0.0, 0.0}, F1[2] = {0.0, 0.0}, B0[2] = {0.0, 0.0}, B1[2] = {0.0, 0.0}, C0[2] = {0.0, 0.0}, C1[2] = {0.0, 0.0}, Db[2] = {0.0, 0.0};
        int cm0[2] = {0, 0}, cm1[2] = {0, 0};                  // axis x: compact psi columns of my two cells; axis z: compact row
#pragma unroll
        for(int part = 0; part < 2; ++part)
        {
            const unsigned fg = part == 0 ? F_PG0 : F_PG1, fs = part == 0 ? F_PS0 : F_PS1;
            hasPs[part] = (any & fs) != 0 && (part == 0 ? HAS_VK : HAS_VJ);
            hasPg[part] = (any & (fg | fs)) != 0 && (part == 0 ? HAS_VK : HAS_VJ);
            if(!hasPg[part]) continue;
            const PmlArgs& pp = ca.pml[part];
            Db[part] = pp.Db;
            if(AX[part] == 0)
            {
                F0[part] = pp.F[x]; F1[part] = pp.F[x + 1];
                if(hasPs[part]) { B0[part] = pp.b[x]; B1[part] = pp.b[x + 1]; C0[part] = pp.c[x]; C1[part] = pp.c[x + 1]; cm0[part] = pp.cmap[x]; cm1[part] = pp.cmap[x + 1]; }
            }
            else if(AX[part] == 2)
            {
                F0[part] = F1[part] = pp.F[z];
                if(hasPs[part]) { B0[part] = B1[part] = pp.b[z]; C0[part] = C1[part] = pp.c[z]; cm0[part] = pp.cmap[z]; }
            }
        }
        const unsigned fsBit[2] = {F_PS0, F_PS1};
        // psi address of (part, plane y) for y / z slabs (a pair); x slabs address two scalars
        auto psi_pair = [&](const int part, const int y) -> double* {
            const PmlArgs& pp = ca.pml[part];
            if(AX[part] == 1) { const int cm = pp.cmap[y]; return pp.psi + x + a.px * (z + (long)a.lz * cm); }
            return pp.psi + x + a.px * (cm0[part] + (long)pp.nact * y);
        };
        // carried y-neighbour planes (register): H_J when K == 1, H_K when J == 1
        double2 carryJ = make_double2(0.0, 0.0), carryK = make_double2(0.0, 0.0);
        if(HAS_VJ && K == 1) carryJ = *reinterpret_cast<const double2*>(fj + r - plane);
        if(HAS_VK && J == 1) carryK = *reinterpret_cast<const double2*>(fk + r - plane);
        const bool leader = (threadIdx.x & 7) == 0;      // one lane per 128-byte line prefetches
        for(int iy = 0; iy < t.ny; ++iy, r += plane)
        {
            const int y = t.y + iy;
            if(leader && iy + PREFETCH_PLANES < t.ny)
            {
                const long rp = r + PREFETCH_PLANES * plane;
                if(HAS_VJ) prefetch_l2(fj + rp);
                if(HAS_VK) prefetch_l2(fk + rp);
                if(needU) prefetch_l2(ca.U + rp);
                if(needD) prefetch_l2(ca.D + rp);
#pragma unroll
                for(int part = 0; part < 2; ++part)
                    if(hasPs[part] && AX[part] != 0) prefetch_l2(psi_pair(part, y + PREFETCH_PLANES));
            }
            // ---- every load of the plane, then the arithmetic
            double2 u = make_double2(0.0, 0.0), vj = u, vk = u, nj = u, nk = u, dv = u;
            if(needU) u = *reinterpret_cast<const double2*>(ca.U + r);
            if(HAS_VJ) vj = *reinterpret_cast<const double2*>(fj + r);
            if(HAS_VK) vk = *reinterpret_cast<const double2*>(fk + r);
            if(HAS_VJ) { if(K == 2) nj = *reinterpret_cast<const double2*>(fj + r - a.px); else if(K == 0) nj = make_double2(fj[r - 1], vj.x); else nj = carryJ; }
            if(HAS_VK) { if(J == 2) nk = *reinterpret_cast<const double2*>(fk + r - a.px); else if(J == 0) nk = make_double2(fk[r - 1], vk.x); else nk = carryK; }
            if(needD) dv = *reinterpret_cast<const double2*>(ca.D + r);
            double ps0[2] = {0.0, 0.0}, ps1[2] = {0.0, 0.0};
            double Fy0[2] = {F0[0], F0[1]}, Fy1[2] = {F1[0], F1[1]}, By0[2] = {B0[0], B0[1]}, By1[2] = {B1[0], B1[1]}, Cy0[2] = {C0[0], C0[1]}, Cy1[2] = {C1[0], C1[1]};
#pragma unroll
            for(int part = 0; part < 2; ++part)
            {
                if(!hasPg[part]) continue;
                const PmlArgs& pp = ca.pml[part];
                if(AX[part] == 1)
                {
                    Fy0[part] = Fy1[part] = pp.F[y];
                    if(hasPs[part]) { By0[part] = By1[part] = pp.b[y]; Cy0[part] = Cy1[part] = pp.c[y]; }
                }
                if(!hasPs[part]) continue;
                if(AX[part] == 0)
                {
                    const long base = pp.psi_pitch * (z + (long)a.lz * y);
                    if(my0 && (k0.info & fsBit[part])) ps0[part] = pp.psi[base + cm0[part]];
                    if(my1 && (k1.info & fsBit[part])) ps1[part] = pp.psi[base + cm1[part]];
                }
                else { const double2 p = *reinterpret_cast<const double2*>(psi_pair(part, y)); ps0[part] = p.x; ps1[part] = p.y; }
            }
            bool d0 = false, d1 = false, pd0[2] = {false, false}, pd1[2] = {false, false};
            if(my0) pipe_cell<HAS_VJ, HAS_VK>(k0, pmlOnD, u.x, dv.x, d0, vj.x, nj.x, vk.x, nk.x, ps0, pd0, Fy0, By0, Cy0, Db);
            if(my1) pipe_cell<HAS_VJ, HAS_VK>(k1, pmlOnD, u.y, dv.y, d1, vj.y, nj.y, vk.y, nk.y, ps1, pd1, Fy1, By1, Cy1, Db);
            store_pair(ca.U + r, u, my0, my1);
            if(d0 || d1) store_pair(ca.D + r, dv, d0, d1);
#pragma unroll
            for(int part = 0; part < 2; ++part)
            {
                if(!hasPs[part] || !(pd0[part] || pd1[part])) continue;
                const PmlArgs& pp = ca.pml[part];
                if(AX[part] == 0)
                {
                    const long base = pp.psi_pitch * (z + (long)a.lz * y);
                    if(pd0[part]) pp.psi[base + cm0[part]] = ps0[part];
                    if(pd1[part]) pp.psi[base + cm1[part]] = ps1[part];
                }
                else store_pair(psi_pair(part, y), make_double2(ps0[part], ps1[part]), pd0[part], pd1[part]);
            }
            if(HAS_VJ && K == 1) carryJ = vj;
            if(HAS_VK && J == 1) carryK = vk;
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(768, 1) k_uniform_cells(const __grid_constant__ StepArgs a, const TileRec* __restrict__ tiles)
{
    const TileRec& t = tiles[blockIdx.x];
    const int xl = 2 * threadIdx.x, zl = threadIdx.y;
    const int x = t.x0 + xl, z = t.z0 + zl;
    if(x >= a.px || z >= a.lz) return;
    if(threadIdx.z == 0)      cell_comp<MODE, 0>(a, t, xl, zl, x, z);
    else if(threadIdx.z == 1) cell_comp<MODE, 1>(a, t, xl, zl, x, z);
    else                      cell_comp<MODE, 2>(a, t, xl, zl, x, z);
}

} // namespace chiml
