// TEST INFRASTRUCTURE ONLY (oracle/): plain C++ definitions of the Fortran BLAS / LAPACK / MKL
// symbols the reference declares in UTIL/utilities_MKL.hpp:17-98, so that its translation units link
// without MKL.  Each routine follows the public (netlib reference) BLAS semantics and loop order:
// one rounded multiply and one rounded add per element (build with -ffp-contract=off), complex
// products by the conventional 4-multiply formula (MKL's zgemm3m uses the 3-multiply variant, which
// differs in the last bits; the 1e-10 parity tolerance absorbs that).  Quirk kept from the
// reference's header: the single-precision names saxpy_/scopy_/sscal_/isamax_/isamin_/sasum_ are
// declared on int data, and are implemented here as integer operations.
#include <cmath>
#include <complex>
#include <cstdlib>
#include <stdexcept>
#include <vector>

typedef std::complex<double> cplx;

static inline cplx cmul(const cplx& a, const cplx& b)
{
    return cplx(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
}
static inline bool isN(char c) { return c == 'N' || c == 'n'; }
static inline bool isT(char c) { return c == 'T' || c == 't'; }
static inline bool isC(char c) { return c == 'C' || c == 'c'; }
static inline bool isR(char c) { return c == 'R' || c == 'r'; }

template <typename T> static inline T conjIf(const T& v, bool) { return v; }
template <> inline cplx conjIf<cplx>(const cplx& v, bool c) { return c ? std::conj(v) : v; }
template <typename T> static inline T mul(const T& a, const T& b) { return a * b; }
template <> inline cplx mul<cplx>(const cplx& a, const cplx& b) { return cmul(a, b); }

template <typename T> static void axpy(int n, T a, const T* x, int incx, T* y, int incy)
{
    if(n <= 0) return;
    int ix = incx < 0 ? (1 - n) * incx : 0;
    int iy = incy < 0 ? (1 - n) * incy : 0;
    for(int i = 0; i < n; ++i, ix += incx, iy += incy)
        y[iy] = y[iy] + mul<T>(a, x[ix]);
}
template <typename T> static void copy(int n, const T* x, int incx, T* y, int incy)
{
    if(n <= 0) return;
    int ix = incx < 0 ? (1 - n) * incx : 0;
    int iy = incy < 0 ? (1 - n) * incy : 0;
    for(int i = 0; i < n; ++i, ix += incx, iy += incy)
        y[iy] = x[ix];
}
template <typename T> static void scal(int n, T a, T* x, int incx)
{
    if(n <= 0 || incx <= 0) return;
    for(int i = 0, ix = 0; i < n; ++i, ix += incx)
        x[ix] = mul<T>(a, x[ix]);
}

// C <- alpha op(A) op(B) + beta C, column major, netlib loop order
template <typename T> static void gemm(char ta, char tb, int m, int n, int k, T alpha, const T* a, int lda, const T* b, int ldb, T beta, T* c, int ldc)
{
    const bool na = isN(ta), nb = isN(tb);
    const bool ca = isC(ta), cb = isC(tb);
    for(int j = 0; j < n; ++j)
    {
        for(int i = 0; i < m; ++i)
            c[i + j * ldc] = (beta == T(0)) ? T(0) : mul<T>(beta, c[i + j * ldc]);
        if(na)
        {
            for(int l = 0; l < k; ++l)
            {
                T bv = nb ? b[l + j * ldb] : conjIf<T>(b[j + l * ldb], cb);
                T temp = mul<T>(alpha, bv);
                for(int i = 0; i < m; ++i)
                    c[i + j * ldc] = c[i + j * ldc] + mul<T>(temp, a[i + l * lda]);
            }
        }
        else
        {
            for(int i = 0; i < m; ++i)
            {
                T temp = T(0);
                for(int l = 0; l < k; ++l)
                {
                    T av = conjIf<T>(a[l + i * lda], ca);
                    T bv = nb ? b[l + j * ldb] : conjIf<T>(b[j + l * ldb], cb);
                    temp = temp + mul<T>(av, bv);
                }
                c[i + j * ldc] = c[i + j * ldc] + mul<T>(alpha, temp);
            }
        }
    }
}

// y <- alpha op(A) x + beta y, column major
template <typename T> static void gemv(char tr, int m, int n, T alpha, const T* a, int lda, const T* x, int incx, T beta, T* y, int incy)
{
    const bool nt = isN(tr);
    const int leny = nt ? m : n, lenx = nt ? n : m;
    int kx = incx > 0 ? 0 : (1 - lenx) * incx;
    int ky = incy > 0 ? 0 : (1 - leny) * incy;
    for(int i = 0, iy = ky; i < leny; ++i, iy += incy)
        y[iy] = (beta == T(0)) ? T(0) : mul<T>(beta, y[iy]);
    if(nt)
    {
        for(int j = 0, jx = kx; j < n; ++j, jx += incx)
        {
            T temp = mul<T>(alpha, x[jx]);
            for(int i = 0, iy = ky; i < m; ++i, iy += incy)
                y[iy] = y[iy] + mul<T>(temp, a[i + j * lda]);
        }
    }
    else
    {
        const bool cj = isC(tr);
        for(int j = 0, jy = ky; j < n; ++j, jy += incy)
        {
            T temp = T(0);
            for(int i = 0, ix = kx; i < m; ++i, ix += incx)
                temp = temp + mul<T>(conjIf<T>(a[i + j * lda], cj), x[ix]);
            y[jy] = y[jy] + mul<T>(alpha, temp);
        }
    }
}

// A <- alpha x op(y)^T + A, column major
template <typename T> static void ger(int m, int n, T alpha, const T* x, int incx, const T* y, int incy, T* a, int lda, bool conjy)
{
    int jy = incy > 0 ? 0 : (1 - n) * incy;
    int kx = incx > 0 ? 0 : (1 - m) * incx;
    for(int j = 0; j < n; ++j, jy += incy)
    {
        if(y[jy] != T(0))
        {
            T temp = mul<T>(alpha, conjIf<T>(y[jy], conjy));
            for(int i = 0, ix = kx; i < m; ++i, ix += incx)
                a[i + j * lda] = a[i + j * lda] + mul<T>(x[ix], temp);
        }
    }
}

template <typename T> static void omatcopy(char ordering, char trans, int rows, int cols, T alpha, const T* A, int lda, T* B, int ldb)
{
    // row-major view; column-major swaps the roles of rows and cols
    if(ordering == 'C' || ordering == 'c') std::swap(rows, cols);
    const bool tr = isT(trans) || isC(trans);
    const bool cj = isC(trans) || isR(trans);
    for(int r = 0; r < rows; ++r)
        for(int c = 0; c < cols; ++c)
        {
            T v = mul<T>(alpha, conjIf<T>(A[r * lda + c], cj));
            if(tr) B[c * ldb + r] = v;
            else   B[r * ldb + c] = v;
        }
}

extern "C"
{
void dgemv_(const char* trans, const int* m, const int* n, const double* alpha, const double* a, const int* lda, const double* x, const int* incx, const double* beta, double* y, const int* incy)
{ gemv<double>(*trans, *m, *n, *alpha, a, *lda, x, *incx, *beta, y, *incy); }

void zgemv_(const char* trans, const int* m, const int* n, const cplx* alpha, const cplx* a, const int* lda, const cplx* x, const int* incx, const cplx* beta, cplx* y, const int* incy)
{ gemv<cplx>(*trans, *m, *n, *alpha, a, *lda, x, *incx, *beta, y, *incy); }

void dgemm_(char* transa, char* transb, const int* m, const int* n, const int* k, const double* alpha, const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c, const int* ldc)
{ gemm<double>(*transa, *transb, *m, *n, *k, *alpha, a, *lda, b, *ldb, *beta, c, *ldc); }

void zgemm3m_(char* transa, char* transb, const int* m, const int* n, const int* k, const cplx* alpha, const cplx* a, const int* lda, const cplx* b, const int* ldb, const cplx* beta, cplx* c, const int* ldc)
{ gemm<cplx>(*transa, *transb, *m, *n, *k, *alpha, a, *lda, b, *ldb, *beta, c, *ldc); }

void zhemm_(char* side, char*, const int* m, const int* n, const cplx* alpha, const cplx* a, const int* lda, const cplx* b, const int* ldb, const cplx* beta, const cplx* c, const int* ldc)
{
    // full storage assumed Hermitian: equivalent to a general product
    char N = 'N';
    if(*side == 'L' || *side == 'l') gemm<cplx>(N, N, *m, *n, *m, *alpha, a, *lda, b, *ldb, *beta, const_cast<cplx*>(c), *ldc);
    else                             gemm<cplx>(N, N, *m, *n, *n, *alpha, b, *ldb, a, *lda, *beta, const_cast<cplx*>(c), *ldc);
}

double ddot_(const int* n, const double* x, const int* incx, const double* y, const int* incy)
{
    double s = 0.0;
    int ix = *incx < 0 ? (1 - *n) * *incx : 0;
    int iy = *incy < 0 ? (1 - *n) * *incy : 0;
    for(int i = 0; i < *n; ++i, ix += *incx, iy += *incy) s = s + x[ix] * y[iy];
    return s;
}

#ifndef ZDOT_RETURN
void zdotc_(cplx* res, const int* n, const cplx* x, const int* incx, const cplx* y, const int* incy)
{
    cplx s(0.0, 0.0);
    int ix = *incx < 0 ? (1 - *n) * *incx : 0;
    int iy = *incy < 0 ? (1 - *n) * *incy : 0;
    for(int i = 0; i < *n; ++i, ix += *incx, iy += *incy) s = s + cmul(std::conj(x[ix]), y[iy]);
    *res = s;
}
#else
cplx zdotc_(const int* n, const cplx* x, const int* incx, const cplx* y, const int* incy)
{
    cplx s(0.0, 0.0);
    int ix = *incx < 0 ? (1 - *n) * *incx : 0;
    int iy = *incy < 0 ? (1 - *n) * *incy : 0;
    for(int i = 0; i < *n; ++i, ix += *incx, iy += *incy) s = s + cmul(std::conj(x[ix]), y[iy]);
    return s;
}
#endif

void saxpy_(const int* n, const int* a, const int* x, const int* incx, int* y, const int* incy) { axpy<int>(*n, *a, x, *incx, y, *incy); }
void scopy_(const int* n, const int* x, const int* incx, int* y, const int* incy) { copy<int>(*n, x, *incx, y, *incy); }
void sscal_(const int* n, const int* a, int* x, const int* incx) { scal<int>(*n, *a, x, *incx); }
void daxpy_(const int* n, const double* a, const double* x, const int* incx, double* y, const int* incy) { axpy<double>(*n, *a, x, *incx, y, *incy); }
void dcopy_(const int* n, const double* x, const int* incx, double* y, const int* incy) { copy<double>(*n, x, *incx, y, *incy); }
void dscal_(const int* n, const double* a, double* x, const int* incx) { scal<double>(*n, *a, x, *incx); }
void zaxpy_(const int* n, const cplx* a, const cplx* x, const int* incx, const cplx* y, const int* incy) { axpy<cplx>(*n, *a, x, *incx, const_cast<cplx*>(y), *incy); }
void zcopy_(const int* n, const cplx* x, const int* incx, cplx* y, const int* incy) { copy<cplx>(*n, x, *incx, y, *incy); }
void zscal_(const int* n, const cplx* a, cplx* x, const int* incx) { scal<cplx>(*n, *a, x, *incx); }

void dger_(const int* m, const int* n, const double* alpha, const double* x, const int* incx, const double* y, const int* incy, double* a, const int* lda)
{ ger<double>(*m, *n, *alpha, x, *incx, y, *incy, a, *lda, false); }
void zgeru_(const int* m, const int* n, const cplx* alpha, const cplx* x, const int* incx, const cplx* y, const int* incy, cplx* a, const int* lda)
{ ger<cplx>(*m, *n, *alpha, x, *incx, y, *incy, a, *lda, false); }
void zgerc_(const int* m, const int* n, const cplx* alpha, const cplx* x, const int* incx, const cplx* y, const int* incy, cplx* a, const int* lda)
{ ger<cplx>(*m, *n, *alpha, x, *incx, y, *incy, a, *lda, true); }

void mkl_domatcopy_(const char* ordering, const char* trans, const int* rows, const int* cols, const double* alpha, const double* A, const int* lda, double* B, const int* ldb)
{ omatcopy<double>(*ordering, *trans, *rows, *cols, *alpha, A, *lda, B, *ldb); }
void mkl_zomatcopy_(const char* ordering, const char* trans, const int* rows, const int* cols, const cplx* alpha, const cplx* A, const int* lda, cplx* B, const int* ldb)
{ omatcopy<cplx>(*ordering, *trans, *rows, *cols, *alpha, A, *lda, B, *ldb); }

// C <- alpha op(A) + beta op(B); m x n result.  In-place use (C aliasing A or B) is only safe for 'N'.
void mkl_zomatadd(const char* ordering, const char* transa, const char* transb, const int* m, const int* n, const cplx* alpha, const cplx* A, const int* lda, const cplx* beta, const cplx* B, const int* ldb, cplx* C, const int* ldc)
{
    const bool rowMajor = (*ordering == 'R' || *ordering == 'r');
    const bool ta = isT(*transa) || isC(*transa), ca = isC(*transa) || isR(*transa);
    const bool tb = isT(*transb) || isC(*transb), cb = isC(*transb) || isR(*transb);
    std::vector<cplx> out(size_t(*m) * size_t(*n));
    for(int i = 0; i < *m; ++i)
        for(int j = 0; j < *n; ++j)
        {
            // element (i,j) of op(X): X(i,j) or X(j,i)
            auto at = [&](const cplx* X, int ld, bool t, int r, int c) -> cplx {
                if(t) std::swap(r, c);
                return rowMajor ? X[r * ld + c] : X[r + c * ld];
            };
            cplx av = conjIf<cplx>(at(A, *lda, ta, i, j), ca);
            cplx bv = conjIf<cplx>(at(B, *ldb, tb, i, j), cb);
            out[size_t(i) * size_t(*n) + j] = cmul(*alpha, av) + cmul(*beta, bv);
        }
    for(int i = 0; i < *m; ++i)
        for(int j = 0; j < *n; ++j)
        {
            if(rowMajor) C[i * *ldc + j] = out[size_t(i) * size_t(*n) + j];
            else         C[i + j * *ldc] = out[size_t(i) * size_t(*n) + j];
        }
}

// index (1-based) of max / min |x|
int idamax_(const int* n, const double* x, const int* inc)
{
    if(*n < 1 || *inc <= 0) return 0;
    int best = 0; double bv = std::abs(x[0]);
    for(int i = 1; i < *n; ++i) { double v = std::abs(x[i * *inc]); if(v > bv) { bv = v; best = i; } }
    return best + 1;
}
int idamin_(const int* n, const double* x, const int* inc)
{
    if(*n < 1 || *inc <= 0) return 0;
    int best = 0; double bv = std::abs(x[0]);
    for(int i = 1; i < *n; ++i) { double v = std::abs(x[i * *inc]); if(v < bv) { bv = v; best = i; } }
    return best + 1;
}
int izamax_(const int* n, const cplx* x, const int* inc)
{
    if(*n < 1 || *inc <= 0) return 0;
    int best = 0; double bv = std::abs(x[0].real()) + std::abs(x[0].imag());
    for(int i = 1; i < *n; ++i) { const cplx& c = x[i * *inc]; double v = std::abs(c.real()) + std::abs(c.imag()); if(v > bv) { bv = v; best = i; } }
    return best + 1;
}
int izamin_(const int* n, const cplx* x, const int* inc)
{
    if(*n < 1 || *inc <= 0) return 0;
    int best = 0; double bv = std::abs(x[0].real()) + std::abs(x[0].imag());
    for(int i = 1; i < *n; ++i) { const cplx& c = x[i * *inc]; double v = std::abs(c.real()) + std::abs(c.imag()); if(v < bv) { bv = v; best = i; } }
    return best + 1;
}
int isamax_(const int* n, const int* x, const int* inc)
{
    if(*n < 1 || *inc <= 0) return 0;
    int best = 0; long bv = std::labs(x[0]);
    for(int i = 1; i < *n; ++i) { long v = std::labs(x[i * *inc]); if(v > bv) { bv = v; best = i; } }
    return best + 1;
}
int isamin_(const int* n, const int* x, const int* inc)
{
    if(*n < 1 || *inc <= 0) return 0;
    int best = 0; long bv = std::labs(x[0]);
    for(int i = 1; i < *n; ++i) { long v = std::labs(x[i * *inc]); if(v < bv) { bv = v; best = i; } }
    return best + 1;
}
int sasum_(const int* n, const int* x, const int* inc)
{
    long s = 0;
    for(int i = 0; i < *n; ++i) s += std::labs(x[i * *inc]);
    return int(s);
}
double dasum_(const int* n, const double* x, const int* inc)
{
    double s = 0.0;
    for(int i = 0; i < *n; ++i) s = s + std::abs(x[i * *inc]);
    return s;
}

// LU with partial pivoting (column major), LAPACK conventions (1-based pivots)
void dgetrf_(const int* m, const int* n, double* a, const int* lda, int* ipiv, int* info)
{
    *info = 0;
    const int mn = *m < *n ? *m : *n;
    for(int j = 0; j < mn; ++j)
    {
        int p = j; double pv = std::abs(a[j + j * *lda]);
        for(int i = j + 1; i < *m; ++i) { double v = std::abs(a[i + j * *lda]); if(v > pv) { pv = v; p = i; } }
        ipiv[j] = p + 1;
        if(a[p + j * *lda] != 0.0)
        {
            if(p != j) for(int c = 0; c < *n; ++c) std::swap(a[j + c * *lda], a[p + c * *lda]);
            double inv = 1.0 / a[j + j * *lda];
            for(int i = j + 1; i < *m; ++i) a[i + j * *lda] = a[i + j * *lda] * inv;
        }
        else if(*info == 0) *info = j + 1;
        for(int c = j + 1; c < *n; ++c)
            for(int i = j + 1; i < *m; ++i)
                a[i + c * *lda] = a[i + c * *lda] - a[i + j * *lda] * a[j + c * *lda];
    }
}

// inverse from the LU factors of dgetrf_
void dgetri_(const int* n, double* a, const int* lda, const int* ipiv, double* work, const int* lwork, int* info)
{
    *info = 0;
    const int N = *n, ld = *lda;
    if(*lwork == -1) { work[0] = double(N > 1 ? N : 1); return; }
    for(int i = 0; i < N; ++i) if(a[i + i * ld] == 0.0) { *info = i + 1; return; }
    // inv(U) in place (upper triangle)
    for(int j = 0; j < N; ++j)
    {
        a[j + j * ld] = 1.0 / a[j + j * ld];
        double ajj = -a[j + j * ld];
        for(int i = 0; i < j; ++i)
        {
            double s = 0.0;
            for(int l = i; l < j; ++l) s = s + a[i + l * ld] * a[l + j * ld];
            work[i] = s;
        }
        for(int i = 0; i < j; ++i) a[i + j * ld] = work[i] * ajj;
    }
    // solve inv(A) L = inv(U)
    std::vector<double> col(N);
    for(int j = N - 2; j >= 0; --j)
    {
        for(int i = j + 1; i < N; ++i) { col[i] = a[i + j * ld]; a[i + j * ld] = 0.0; }
        for(int l = j + 1; l < N; ++l)
            for(int i = 0; i < N; ++i)
                a[i + j * ld] = a[i + j * ld] - a[i + l * ld] * col[l];
    }
    for(int j = N - 2; j >= 0; --j)
    {
        int jp = ipiv[j] - 1;
        if(jp != j) for(int i = 0; i < N; ++i) std::swap(a[i + j * ld], a[i + jp * ld]);
    }
}

void dsyev_(const char*, const char*, const int*, double*, const int*, double*, double*, const int*, int*)
{ throw std::runtime_error("blas shim: dsyev_ is not on the time-stepping path and is not provided"); }
void dgesvd_(const char*, const char*, const int*, const int*, double*, const int*, double*, double*, const int*, double*, const int*, double*, const int*, int*)
{ throw std::runtime_error("blas shim: dgesvd_ is not on the time-stepping path and is not provided"); }
} // extern "C"
