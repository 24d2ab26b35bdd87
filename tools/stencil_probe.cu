// Development probe: the bare interior H half step (curl of E into H, 3-D, uniform vacuum) in a few
// structural variants, to find which ingredient of k_update costs bandwidth.  Not part of the product.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

struct A { double* H[3]; const double* E[3]; const unsigned* td; const double2* pf; int lx, ly, lz; long px; unsigned nxt, nzt; };

__device__ __forceinline__ double ax(double y, double a, double x) { return __dadd_rn(y, __dmul_rn(a, x)); }

// variant bits: 1 = tile-descriptor load gates everything; 2 = pf from table (dependent on td) instead of constants
template <int VAR, int MINB>
__global__ void __launch_bounds__(256, MINB) k_h(const __grid_constant__ A a)
{
    unsigned b = blockIdx.x;
    const unsigned xt = b % a.nxt; b /= a.nxt;
    const unsigned zt = b % a.nzt;
    const int y = b / a.nzt;
    unsigned td = 1u << 24 | 0x010101u;
    if(VAR & 1) { td = a.td[((size_t)y * a.nzt + zt) * a.nxt + xt]; if((td >> 24) == 2u) return; }
    const int x = 2 * (xt * 32 + threadIdx.x);
    const int z = zt * 8 + threadIdx.y;
    if(x >= a.px || z >= a.lz) return;
    const long plane = a.px * a.lz;
    const long r = x + a.px * (z + (long)a.lz * y);
    double2 u[3], v[3];
#pragma unroll
    for(int c = 0; c < 3; ++c) { u[c] = *(const double2*)(a.H[c] + r); v[c] = *(const double2*)(a.E[c] + r); }
    const double2 nj0 = *(const double2*)(a.E[1] + r + a.px);     // Ey z+1
    const double2 nk0 = *(const double2*)(a.E[2] + r + plane);    // Ez y+1
    const double2 nj1 = make_double2(v[2].y, a.E[2][r + 2]);      // Ez x+1
    const double2 nk1 = *(const double2*)(a.E[0] + r + a.px);     // Ex z+1
    const double2 nj2 = *(const double2*)(a.E[0] + r + plane);    // Ex y+1
    const double2 nk2 = make_double2(v[1].y, a.E[1][r + 2]);      // Ey x+1
    double2 pf[3];
    if(VAR & 2) { pf[0] = a.pf[td & 0xFF]; pf[1] = a.pf[(td >> 8) & 0xFF]; pf[2] = a.pf[(td >> 16) & 0xFF]; }
    else pf[0] = pf[1] = pf[2] = make_double2(-0.25, -0.25);
    const double2 vj[3] = {v[1], v[2], v[0]}, vk[3] = {v[2], v[0], v[1]};
    const double2 nj[3] = {nj0, nj1, nj2}, nk[3] = {nk0, nk1, nk2};
#pragma unroll
    for(int c = 0; c < 3; ++c)
    {
        double2 t = u[c];
        t.x = ax(t.x, pf[c].y, vj[c].x);  t.y = ax(t.y, pf[c].y, vj[c].y);
        t.x = ax(t.x, -pf[c].y, nj[c].x); t.y = ax(t.y, -pf[c].y, nj[c].y);
        t.x = ax(t.x, -pf[c].x, vk[c].x); t.y = ax(t.y, -pf[c].x, vk[c].y);
        t.x = ax(t.x, pf[c].x, nk[c].x);  t.y = ax(t.y, pf[c].x, nk[c].y);
        *(double2*)(a.H[c] + r) = t;
    }
}

int main()
{
    const int lx = 258, ly = 258, lz = 258;
    A a; a.lx = lx; a.ly = ly; a.lz = lz; a.px = 272;
    const size_t plane = (size_t)a.px * lz, n = plane * ly, guard = plane + 32;
    for(int c = 0; c < 3; ++c)
    {
        double* d; cudaMalloc(&d, (n + 2 * guard) * 8); cudaMemset(d, 0, (n + 2 * guard) * 8); a.H[c] = d + guard;
        cudaMalloc(&d, (n + 2 * guard) * 8); cudaMemset(d, 0, (n + 2 * guard) * 8); a.E[c] = d + guard;
    }
    a.nxt = (lx + 63) / 64; a.nzt = (lz + 7) / 8;
    const size_t ntiles = (size_t)a.nxt * a.nzt * ly;
    std::vector<unsigned> td(ntiles, 1u << 24 | 0x010101u);
    unsigned* dtd; cudaMalloc(&dtd, ntiles * 4); cudaMemcpy(dtd, td.data(), ntiles * 4, cudaMemcpyHostToDevice); a.td = dtd;
    std::vector<double2> pf(256, make_double2(-0.25, -0.25));
    double2* dpf; cudaMalloc(&dpf, 256 * 16); cudaMemcpy(dpf, pf.data(), 256 * 16, cudaMemcpyHostToDevice); a.pf = dpf;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = 9.0 * (double)(lx - 2) * (ly - 2) * (lz - 2) * 8;
    auto run = [&](const char* name, auto launch) {
        for(int w = 0; w < 3; ++w) launch();
        cudaEventRecord(e0);
        for(int it = 0; it < 20; ++it) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-40s %8.1f GB/s algorithmic  %.3f ms  (%s)\n", name, bytes * 20 / (ms * 1e-3) / 1e9, ms / 20, cudaGetErrorString(cudaGetLastError()));
    };
    const dim3 blk(32, 8);
    const unsigned grid = (unsigned)ntiles;
    run("bare stencil, consts, minb2", [&] { k_h<0, 2><<<grid, blk>>>(a); });
    run("bare stencil, consts, minb4", [&] { k_h<0, 4><<<grid, blk>>>(a); });
    run("bare stencil, consts, minb6", [&] { k_h<0, 6><<<grid, blk>>>(a); });
    run("+ tile desc gate, minb4", [&] { k_h<1, 4><<<grid, blk>>>(a); });
    run("+ tile desc + pf table, minb4", [&] { k_h<3, 4><<<grid, blk>>>(a); });
    run("+ tile desc + pf table, minb2", [&] { k_h<3, 2><<<grid, blk>>>(a); });
    run("+ tile desc + pf table, minb6", [&] { k_h<3, 6><<<grid, blk>>>(a); });
    return 0;
}
