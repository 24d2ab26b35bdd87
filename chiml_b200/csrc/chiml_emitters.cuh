// Maxwell-Liouville quantum-emitter kernels: reference ML/parallelQE.hpp (addQE :682-718, updateDensity :614-678, PCABAM4 :751-770,
// denDeriv :727-744), ML/Hamiltonian.cpp:59-69, UTIL/FDTD_up_eq.cpp:1367-1430, UTIL/FDTD_up_eq.hpp:620, ML/QEPopDtc.{hpp,cpp}.
//
// Two kernels, chosen per emitter set by the number of levels N (EmitterDev::group):
//   k_emit_density<N>       one THREAD per emitter node: all N x N elements of rho and of its four derivative histories in registers.
//                           State is SoA over emitters -- element (sys, k, re/im) of emitter e at ((sys*N*N + k)*2 + ri)*nemit + e -- so
//                           that a warp's loads are contiguous.  Best for N = 2, 3 (255 registers, no spills for N = 2); for N >= 4 the
//                           10 N^2 doubles of state spill (2.6 KB of stack at N = 4).
//   k_emit_density_g<N, G>  a GROUP of G = 4 / 16 / 16 / 32 lanes per emitter node (N = 2..5), lane k holding element k of every
//                           matrix: 10 doubles of state per lane, no spills.  The small complex products H rho exchange their operands
//                           with warp shuffles -- the column of rho and this lane's row of H for the zgemm, the transposed element for
//                           T + T^H, the gam_ column for the relaxation term, lane 0 gathering rho for Re<rho|mu>.  State is AoS:
//                           element (sys, e, k) is the complex number at ((sys * nemit + e) * N*N + k), a warp reads 32 consecutive
//                           complex numbers (512 B) per array.  Used for N = 4, 5 (measured on C4's two-level sheet the group kernel
//                           is 36 % SLOWER than the thread kernel: four lanes per emitter do redundant H rows and 40 shuffles).
// Either way every element is accumulated in the reference's order (results are bit-identical between the two), the level systems
// of a node are walked in the reference's order so that the polarisation is summed in the same order, and the four derivative
// histories rotate through four slots (no copies).  Tiny non-GEMM complex contractions: CUDA cores, FP64, no tensor cores.
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
    int gam_maxrow;            // longest row of gam_ (the group kernel walks every row that far, in lock-step)
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
            const size_t base = (size_t)sy * N2 * 2 * a.nemit + e;        // SoA over emitters: a warp's loads are contiguous
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


// ---- group-per-emitter kernel ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ cxd shfl_cx(const cxd v, const int src)
{
    cxd r;
    r.re = __shfl_sync(0xffffffffu, v.re, src);
    r.im = __shfl_sync(0xffffffffu, v.im, src);
    return r;
}

// lane m of the group (m = this lane's element, gb = first lane of the group in the warp): element m of d rho/dt for the density `den`
// (every lane passes ITS element), given this lane's row of H: Hrow[l] = H[(m % N) + l*N].  Same operations in the same order as
// den_deriv<N> above, element by element.
template <int N>
__device__ __forceinline__ cxd den_deriv_g(const EmitArgs& a, const cxd* Hrow, const cxd den, const int m, const int gb, const bool live)
{
    constexpr int N2 = N * N;
    const int col = m / N;                      // T is indexed i + j*N with i = m % N, j = m / N
    cxd alpha; alpha.re = 0.0; alpha.im = a.inv_hbar;
    cxd T; T.re = 0.0; T.im = 0.0;
#pragma unroll
    for(int l = 0; l < N; ++l)
    {
        const cxd dl = shfl_cx(den, gb + (live ? l + col * N : 0));
        const cxd temp = cmul(alpha, dl);
        T = cadd(T, cmul(temp, Hrow[l]));
    }
    // out[i*N + j] = 1 * T[i*N + j] + 1 * conj(T[j*N + i]): the partner is the transposed element
    const int tr = (m % N) * N + m / N;
    const cxd Tt = shfl_cx(T, gb + (live ? tr : 0));
    cxd one; one.re = 1.0; one.im = 0.0;
    cxd b; b.re = Tt.re; b.im = -Tt.im;
    cxd out = cadd(cmul(one, T), cmul(one, b));
    // relaxation: out[m] += gam_val[k] * den[gam_col[k]] over row m, in the row's stored order; every lane walks gam_maxrow entries so
    // that the shuffles stay in lock-step
    const int k0 = live ? a.gam_ptr[m] : 0, k1 = live ? a.gam_ptr[m + 1] : 0;
    for(int kk = 0; kk < a.gam_maxrow; ++kk)
    {
        const bool has = k0 + kk < k1;
        const int c = has ? a.gam_col[k0 + kk] : 0;
        const cxd v = shfl_cx(den, gb + c);
        if(has)
        {
            const double g = a.gam_val[k0 + kk];
            out.re = __dadd_rn(out.re, __dmul_rn(v.re, g));
            out.im = __dadd_rn(out.im, __dmul_rn(v.im, g));
        }
    }
    (void)N2;
    return out;
}

template <int N, int G>
__global__ void __launch_bounds__(128, 4) k_emit_density_g(const __grid_constant__ EmitArgs a)
{
    constexpr int N2 = N * N, EPB = 128 / G;          // emitters per block
    static_assert(N2 <= G && G <= 32 && (G & (G - 1)) == 0, "one element per lane, a group inside one warp");
    const int m = threadIdx.x % G;                    // this lane's element of the matrices
    const int gb = (threadIdx.x % 32) / G * G;        // first lane of the group inside its warp
    const int eb = threadIdx.x / G;                   // emitter of the block
    const int e = blockIdx.x * EPB + eb;
    const bool live = e < a.nemit && m < N2;
    const int ee = e < a.nemit ? e : 0;               // idle groups shadow emitter 0 (they take part in the shuffles, store nothing)
    const cxd* rhoA = reinterpret_cast<const cxd*>(a.rho);
    const cxd* fA[4] = {reinterpret_cast<const cxd*>(a.f[0]), reinterpret_cast<const cxd*>(a.f[1]), reinterpret_cast<const cxd*>(a.f[2]),
                        reinterpret_cast<const cxd*>(a.f[3])};
    const int lx0 = a.loc[3 * ee], ly0 = a.loc[3 * ee + 1], lz0 = a.loc[3 * ee + 2];
    const int gx = a.box_lo[0] + 1 + lx0, gy = a.box_lo[1] + 1 + ly0, gz = a.threeD ? a.box_lo[2] + 1 + lz0 : 0;
    const long r = gx + a.px * (gz + (long)a.lz * gy);
    const long plane = a.px * a.lz;
    // node field: e_c = 0.5 E_c[r] + 0.5 E_c[r - e_c]  (getE_TE); Ez is copied in TM mode (getE_TM)
    double ev[3] = {0.0, 0.0, 0.0};
    if(a.E[0]) ev[0] = __dadd_rn(__dmul_rn(0.5, a.E[0][r]), __dmul_rn(0.5, a.E[0][r - 1]));
    if(a.E[1]) ev[1] = __dadd_rn(__dmul_rn(0.5, a.E[1][r]), __dmul_rn(0.5, a.E[1][r - plane]));
    if(a.E[2]) ev[2] = a.tm ? a.E[2][r] : __dadd_rn(__dmul_rn(0.5, a.E[2][r]), __dmul_rn(0.5, a.E[2][r - (a.threeD ? a.px : 0)]));
    const double dt = a.dt;
    const double c55 = 55.0 * dt / 24.0, c59 = -59.0 * dt / 24.0, c37 = 37.0 * dt / 24.0, c9m = -9.0 * dt / 24.0;
    const double c9 = 9.0 * dt / 24.0, c19 = 19.0 * dt / 24.0, c5m = -5.0 * dt / 24.0, c1 = dt / 24.0;
    const int mm = m < N2 ? m : 0;
    double Pacc[3] = {0.0, 0.0, 0.0};
    double popv[EMIT_MAX_POP][2];
#pragma unroll
    for(int p = 0; p < EMIT_MAX_POP; ++p) { popv[p][0] = 0.0; popv[p][1] = 0.0; }
    for(int sy = 0; sy < a.nsys; ++sy)
    {
        const size_t idx = ((size_t)sy * a.nemit + ee) * N2 + mm;
        cxd rho = rhoA[idx];
        const cxd f0 = fA[0][idx], f1 = fA[1][idx], f2 = fA[2][idx], f3 = fA[3][idx];
        // PCABAM4 predictor (:754-758)
        cxd pred = rho;
        pred = zaxpy_r(pred, c55, f0);
        pred = zaxpy_r(pred, c59, f1);
        pred = zaxpy_r(pred, c37, f2);
        pred = zaxpy_r(pred, c9m, f3);
        // Hamiltonian::getHam, the row of H this element's sums run over: H[(m % N) + l*N], l < N
        cxd Hrow[N];
        const cxd* h0 = reinterpret_cast<const cxd*>(a.h0) + (size_t)sy * N2;
#pragma unroll
        for(int l = 0; l < N; ++l)
        {
            const int q = (mm % N) + l * N;
            cxd h = h0[q];
#pragma unroll
            for(int c = 0; c < 3; ++c)
            {
                if(!a.mu_present[c]) continue;
                const cxd mu = (reinterpret_cast<const cxd*>(a.mu) + (size_t)c * N2)[q];
                cxd ce; ce.re = ev[c]; ce.im = 0.0;
                cxd neg; neg.re = -mu.re; neg.im = -mu.im;
                h = cadd(h, cmul(ce, neg));
            }
            Hrow[l] = h;
        }
        cxd fp = den_deriv_g<N>(a, Hrow, pred, mm, gb, m < N2);
        // corrector (:761-764)
        rho = zaxpy_r(rho, c9, fp);
        rho = zaxpy_r(rho, c19, f0);
        rho = zaxpy_r(rho, c5m, f1);
        rho = zaxpy_r(rho, c1, f2);
        // new derivative (:769) into the slot of the oldest history
        fp = den_deriv_g<N>(a, Hrow, rho, mm, gb, m < N2);
        if(live)
        {
            reinterpret_cast<cxd*>(a.rho)[idx] = rho;
            reinterpret_cast<cxd*>(a.f[3])[idx] = fp;
        }
        // QEPopDtc::inPop (lane 0 of the group keeps the sums)
        if(a.sample)
#pragma unroll
            for(int p = 0; p < EMIT_MAX_POP; ++p)
                if(p < a.npop)
                {
                    const cxd v = shfl_cx(rho, gb + a.pop_level[p]);
                    popv[p][0] = __dadd_rn(popv[p][0], v.re);
                    popv[p][1] = __dadd_rn(popv[p][1], v.im);
                }
        // updateQEPol: P_c += na * Re(zdotc(rho, mu_c)), the dot product summed over k in order (every lane runs it, lane 0 keeps it)
#pragma unroll
        for(int c = 0; c < 3; ++c)
        {
            if(!a.E[c]) continue;
            const cxd* mu = reinterpret_cast<const cxd*>(a.mu) + (size_t)c * N2;
            cxd acc; acc.re = 0.0; acc.im = 0.0;
#pragma unroll
            for(int k = 0; k < N2; ++k)
            {
                const cxd rk = shfl_cx(rho, gb + k);
                cxd cj; cj.re = rk.re; cj.im = -rk.im;
                acc = cadd(acc, cmul(cj, mu[k]));
            }
            Pacc[c] = __dadd_rn(Pacc[c], __dmul_rn(a.na, acc.re));
        }
    }
    if(m == 0 && e < a.nemit)
    {
        const long pi = (lx0 + 1) + (long)a.bx * ((lz0 + (a.threeD ? 1 : 0)) + (long)a.bz * (ly0 + 1));
#pragma unroll
        for(int c = 0; c < 3; ++c)
            if(a.E[c]) a.P[c][pi] = Pacc[c];
    }
    // deterministic per-block partial sums of the sampled populations (one value per emitter of the block, fixed tree)
    if(a.sample && a.npop > 0)
    {
        __shared__ double sh[EPB];
        for(int p = 0; p < a.npop; ++p)
            for(int ri = 0; ri < 2; ++ri)
            {
                if(m == 0) sh[eb] = e < a.nemit ? popv[p][ri] : 0.0;
                __syncthreads();
                for(int s = EPB / 2; s > 0; s >>= 1)
                {
                    if(m == 0 && eb < s) sh[eb] = __dadd_rn(sh[eb], sh[eb + s]);
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
