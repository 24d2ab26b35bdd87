// Maxwell-Liouville quantum-emitter kernels: reference ML/parallelQE.hpp (addQE :682-718, updateDensity :614-678, PCABAM4 :751-770,
// denDeriv :727-744), ML/Hamiltonian.cpp:59-69, UTIL/FDTD_up_eq.cpp:1367-1430, UTIL/FDTD_up_eq.hpp:620, ML/QEPopDtc.{hpp,cpp}.
//
// One thread per emitter node; the thread walks every level system of its node in the reference's order, so the polarisation it
// writes is summed in the same order.  State is SoA over emitters -- element (sys, k, re/im) of emitter e lives at
// ((sys*N*N + k)*2 + ri)*nemit + e -- so that a warp's loads are contiguous.  The four derivative histories rotate through four
// slots (no copies): the kernel reads all four and overwrites the slot of the oldest with the new derivative.
// These are tiny non-GEMM complex contractions (N = 2..8): CUDA cores, FP64, no tensor cores.
#pragma once

namespace chiml {

constexpr int EMIT_MAX_POP = 8;

struct EmitArgs
{
    // main-grid fields
    double* E[3];              // nullptr when the component does not exist in this mode
    int lz; long px;
    int threeD, tm;            // tm: Ez-only mode (getE_TM copies instead of averaging)
    // the emitter set
    int nemit, nsys, nlevel;
    int box_lo[3];
    int bx, bz;                // P box pitch: x extent (n0+2) and z extent (pz)
    double dt, inv_hbar, na;
    const double* h0;          // nsys * n2 complex
    const double* mu;          // 3 * n2 complex
    int mu_present[3];
    const int32_t* gam_ptr; const int32_t* gam_col; const double* gam_val;
    const int32_t* loc;        // 3 * nemit
    const double* eps;         // P-box shaped
    double* P[3];
    double* rho;               // state slot 0
    double* f[4];              // f_n, f_{n-1}, f_{n-2}, f_{n-3} BEFORE this step; f[3] receives the new f_n
    // population detectors
    int npop; int sample;
    int pop_level[EMIT_MAX_POP];
    double* pop_partial;       // [npop][gridDim.x][2]
};

struct cxd { double re, im; };
__device__ __forceinline__ cxd cmul(const cxd a, const cxd b)
{
    cxd r;
    r.re = __dsub_rn(__dmul_rn(a.re, b.re), __dmul_rn(a.im, b.im));
    r.im = __dadd_rn(__dmul_rn(a.re, b.im), __dmul_rn(a.im, b.re));
    return r;
}
__device__ __forceinline__ cxd cadd(const cxd a, const cxd b) { cxd r; r.re = __dadd_rn(a.re, b.re); r.im = __dadd_rn(a.im, b.im); return r; }
// y += cplx(a, 0) * x  (zaxpy_ with a real scalar promoted to complex)
__device__ __forceinline__ cxd zaxpy_r(const cxd y, const double a, const cxd x)
{
    cxd ca; ca.re = a; ca.im = 0.0;
    return cadd(y, cmul(ca, x));
}

// parallelQEBase::denDeriv, MKL branch: T = (i/hbar) H rho as a column-major zgemm on the row-major arrays, out = T + T^H, then gam_
template <int N>
__device__ __forceinline__ void den_deriv(const EmitArgs& a, const cxd* H, const cxd* den, cxd* out)
{
    cxd T[N * N];
    cxd alpha; alpha.re = 0.0; alpha.im = a.inv_hbar;
#pragma unroll
    for(int j = 0; j < N; ++j)
    {
#pragma unroll
        for(int i = 0; i < N; ++i) { T[i + j * N].re = 0.0; T[i + j * N].im = 0.0; }
#pragma unroll
        for(int l = 0; l < N; ++l)
        {
            const cxd temp = cmul(alpha, den[l + j * N]);
#pragma unroll
            for(int i = 0; i < N; ++i) T[i + j * N] = cadd(T[i + j * N], cmul(temp, H[i + l * N]));
        }
    }
    cxd one; one.re = 1.0; one.im = 0.0;
#pragma unroll
    for(int i = 0; i < N; ++i)
#pragma unroll
        for(int j = 0; j < N; ++j)
        {
            cxd b; b.re = T[j * N + i].re; b.im = -T[j * N + i].im;
            out[i * N + j] = cadd(cmul(one, T[i * N + j]), cmul(one, b));
        }
#pragma unroll
    for(int ii = 0; ii < N * N; ++ii)
        for(int k = a.gam_ptr[ii]; k < a.gam_ptr[ii + 1]; ++k)
        {
            const cxd v = den[a.gam_col[k]];
            const double g = a.gam_val[k];
            out[ii].re = __dadd_rn(out[ii].re, __dmul_rn(v.re, g));
            out[ii].im = __dadd_rn(out[ii].im, __dmul_rn(v.im, g));
        }
}

template <int N>
__global__ void __launch_bounds__(128) k_emit_density(const __grid_constant__ EmitArgs a)
{
    constexpr int N2 = N * N;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = e < a.nemit;
    double popv[EMIT_MAX_POP][2];
#pragma unroll
    for(int p = 0; p < EMIT_MAX_POP; ++p) { popv[p][0] = 0.0; popv[p][1] = 0.0; }
    if(valid)
    {
        const int lx0 = a.loc[3 * e], ly0 = a.loc[3 * e + 1], lz0 = a.loc[3 * e + 2];
        const int gx = a.box_lo[0] + 1 + lx0, gy = a.box_lo[1] + 1 + ly0, gz = a.threeD ? a.box_lo[2] + 1 + lz0 : 0;
        const long r = gx + a.px * (gz + (long)a.lz * gy);
        const long plane = a.px * a.lz;
        // node field: e_c = 0.5 E_c[r] + 0.5 E_c[r - e_c]  (getE_TE); Ez is copied in TM mode (getE_TM)
        double ev[3] = {0.0, 0.0, 0.0};
        if(a.E[0]) ev[0] = __dadd_rn(__dmul_rn(0.5, a.E[0][r]), __dmul_rn(0.5, a.E[0][r - 1]));
        if(a.E[1]) ev[1] = __dadd_rn(__dmul_rn(0.5, a.E[1][r]), __dmul_rn(0.5, a.E[1][r - plane]));
        if(a.E[2]) ev[2] = a.tm ? a.E[2][r] : __dadd_rn(__dmul_rn(0.5, a.E[2][r]), __dmul_rn(0.5, a.E[2][r - (a.threeD ? a.px : 0)]));
        double Pacc[3] = {0.0, 0.0, 0.0};
        const double dt = a.dt;
        const double c55 = 55.0 * dt / 24.0, c59 = -59.0 * dt / 24.0, c37 = 37.0 * dt / 24.0, c9m = -9.0 * dt / 24.0;
        const double c9 = 9.0 * dt / 24.0, c19 = 19.0 * dt / 24.0, c5m = -5.0 * dt / 24.0, c1 = dt / 24.0;
        for(int sy = 0; sy < a.nsys; ++sy)
        {
            const size_t base = (size_t)sy * N2 * 2 * a.nemit + e;
            cxd rho[N2], f0[N2], f1[N2], f2[N2], pred[N2], H[N2], fp[N2];
#pragma unroll
            for(int k = 0; k < N2; ++k)
            {
                const size_t i0 = base + (size_t)(2 * k) * a.nemit, i1 = i0 + a.nemit;
                rho[k].re = a.rho[i0];  rho[k].im = a.rho[i1];
                f0[k].re = a.f[0][i0];  f0[k].im = a.f[0][i1];
                f1[k].re = a.f[1][i0];  f1[k].im = a.f[1][i1];
                f2[k].re = a.f[2][i0];  f2[k].im = a.f[2][i1];
                cxd f3; f3.re = a.f[3][i0]; f3.im = a.f[3][i1];
                // PCABAM4 predictor (:754-758)
                cxd p = rho[k];
                p = zaxpy_r(p, c55, f0[k]);
                p = zaxpy_r(p, c59, f1[k]);
                p = zaxpy_r(p, c37, f2[k]);
                p = zaxpy_r(p, c9m, f3);
                pred[k] = p;
            }
            // Hamiltonian::getHam
            const cxd* h0 = reinterpret_cast<const cxd*>(a.h0) + (size_t)sy * N2;
#pragma unroll
            for(int k = 0; k < N2; ++k) H[k] = h0[k];
#pragma unroll
            for(int c = 0; c < 3; ++c)
            {
                if(!a.mu_present[c]) continue;
                const cxd* mu = reinterpret_cast<const cxd*>(a.mu) + (size_t)c * N2;
                cxd ce; ce.re = ev[c]; ce.im = 0.0;
#pragma unroll
                for(int k = 0; k < N2; ++k)
                {
                    cxd neg; neg.re = -mu[k].re; neg.im = -mu[k].im;
                    H[k] = cadd(H[k], cmul(ce, neg));
                }
            }
            den_deriv<N>(a, H, pred, fp);
            // corrector (:761-764)
#pragma unroll
            for(int k = 0; k < N2; ++k)
            {
                cxd v = rho[k];
                v = zaxpy_r(v, c9, fp[k]);
                v = zaxpy_r(v, c19, f0[k]);
                v = zaxpy_r(v, c5m, f1[k]);
                v = zaxpy_r(v, c1, f2[k]);
                rho[k] = v;
            }
            // new derivative (:769) into the slot of the oldest history
            den_deriv<N>(a, H, rho, fp);
#pragma unroll
            for(int k = 0; k < N2; ++k)
            {
                const size_t i0 = base + (size_t)(2 * k) * a.nemit, i1 = i0 + a.nemit;
                a.rho[i0] = rho[k].re;   a.rho[i1] = rho[k].im;
                a.f[3][i0] = fp[k].re;   a.f[3][i1] = fp[k].im;
            }
            // QEPopDtc::inPop
            if(a.sample)
#pragma unroll
                for(int p = 0; p < EMIT_MAX_POP; ++p)
                    if(p < a.npop)
                    {
                        cxd v = rho[0];
#pragma unroll
                        for(int k = 1; k < N2; ++k) if(k == a.pop_level[p]) v = rho[k];
                        popv[p][0] = __dadd_rn(popv[p][0], v.re);
                        popv[p][1] = __dadd_rn(popv[p][1], v.im);
                    }
            // updateQEPol: P_c += na * Re(zdotc(rho, mu_c))
#pragma unroll
            for(int c = 0; c < 3; ++c)
            {
                if(!a.E[c]) continue;
                const cxd* mu = reinterpret_cast<const cxd*>(a.mu) + (size_t)c * N2;
                cxd acc; acc.re = 0.0; acc.im = 0.0;
#pragma unroll
                for(int k = 0; k < N2; ++k)
                {
                    cxd cj; cj.re = rho[k].re; cj.im = -rho[k].im;
                    acc = cadd(acc, cmul(cj, mu[k]));
                }
                Pacc[c] = __dadd_rn(Pacc[c], __dmul_rn(a.na, acc.re));
            }
        }
        const long pi = (lx0 + 1) + (long)a.bx * ((lz0 + (a.threeD ? 1 : 0)) + (long)a.bz * (ly0 + 1));
#pragma unroll
        for(int c = 0; c < 3; ++c)
            if(a.E[c]) a.P[c][pi] = Pacc[c];
    }
    // deterministic per-block partial sums of the sampled populations
    if(a.sample && a.npop > 0)
    {
        __shared__ double sh[128];
        for(int p = 0; p < a.npop; ++p)
            for(int ri = 0; ri < 2; ++ri)
            {
                sh[threadIdx.x] = popv[p][ri];
                __syncthreads();
                for(int s = 64; s > 0; s >>= 1)
                {
                    if(threadIdx.x < s) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + s]);
                    __syncthreads();
                }
                if(threadIdx.x == 0) a.pop_partial[((size_t)p * gridDim.x + blockIdx.x) * 2 + ri] = sh[0];
                __syncthreads();
            }
    }
}

// second stage of the population sum: one block, fixed order; appends curPop / npoints (QEPopDtc::accumPop)
__global__ void k_emit_pop_reduce(const double* partial, int nblocks, int npop, double inv_npoints_unused, double npoints, double* ring, size_t cap, size_t sample)
{
    __shared__ double sh[256];
    for(int p = 0; p < npop; ++p)
        for(int ri = 0; ri < 2; ++ri)
        {
            double s = 0.0;
            for(int b = threadIdx.x; b < nblocks; b += blockDim.x) s = __dadd_rn(s, partial[((size_t)p * nblocks + b) * 2 + ri]);
            sh[threadIdx.x] = s;
            __syncthreads();
            for(int k = 128; k > 0; k >>= 1)
            {
                if(threadIdx.x < k) sh[threadIdx.x] = __dadd_rn(sh[threadIdx.x], sh[threadIdx.x + k]);
                __syncthreads();
            }
            if(threadIdx.x == 0) ring[((size_t)p * cap + sample) * 2 + ri] = __ddiv_rn(sh[0], npoints);
            __syncthreads();
        }
}

// addP (UTIL/FDTD_up_eq.cpp:1367-1380): E[r] += -0.5 P[n]/eps[n]; E[r] += -0.5 P[n+off]/eps[n+off] over the box of (n+1)^3 cells
struct AddPArgs
{
    double* E[3];
    const double* P[3];
    const double* eps;
    int box_lo[3];
    int nx, ny, nz;            // box extents (n0+1, n1+1, n2+1 or 1)
    int bx, bz;
    int zoff;                  // zOff_
    int lz; long px;
    int iy0, iy1;              // box rows [iy0, iy1) handled by this launch (slab-boundary rows go first, see chiml_halo.cuh)
};
__global__ void k_emit_addP(const __grid_constant__ AddPArgs a)
{
    const long n = (long)a.nx * (a.iy1 - a.iy0) * a.nz;
    for(long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    {
        const int ix = (int)(i % a.nx);
        const int iz = (int)((i / a.nx) % a.nz);
        const int iy = a.iy0 + (int)(i / ((long)a.nx * a.nz));
        const long g = (a.box_lo[0] + ix) + a.px * ((a.lz > 1 ? a.box_lo[2] + iz : 0) + (long)a.lz * (a.box_lo[1] + iy));
        const long p0 = ix + (long)a.bx * (iz + (long)a.bz * iy);
        const double ep0 = a.eps[p0];
#pragma unroll
        for(int c = 0; c < 3; ++c)
        {
            if(!a.E[c]) continue;
            const long p1 = (ix + (c == 0)) + (long)a.bx * ((iz + (c == 2 ? a.zoff : 0)) + (long)a.bz * (iy + (c == 1)));
            double ev = a.E[c][g];
            ev = __dadd_rn(ev, __ddiv_rn(__dmul_rn(-0.5, a.P[c][p0]), ep0));
            ev = __dadd_rn(ev, __ddiv_rn(__dmul_rn(-0.5, a.P[c][p1]), a.eps[p1]));
            a.E[c][g] = ev;
        }
    }
}

} // namespace chiml
