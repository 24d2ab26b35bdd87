// Development probe: what HBM throughput do 6-read / 3-write FP64 streams reach on this GPU for different
// thread/work shapes?  (Not part of the product; used to choose the structure of k_update.)
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

struct Ptrs { const double* in[6]; double* out[3]; };

// A: one-shot, each thread 2 doubles of every stream (like k_update v4), RMW on out
template <int VEC>
__global__ void k_oneshot(Ptrs p, size_t n)
{
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if(i >= n) return;
    double acc[VEC];
#pragma unroll
    for(int v = 0; v < VEC; ++v) acc[v] = 0.0;
#pragma unroll
    for(int k = 0; k < 3; ++k)
#pragma unroll
        for(int v = 0; v < VEC; v += 2)
        {
            double2 a = *reinterpret_cast<const double2*>(p.in[k] + i + v);
            acc[v] += a.x; acc[v + 1] += a.y;
        }
#pragma unroll
    for(int k = 0; k < 3; ++k)
#pragma unroll
        for(int v = 0; v < VEC; v += 2)
        {
            double2 o = *reinterpret_cast<const double2*>(p.out[k] + i + v);
            o.x += acc[v]; o.y += acc[v + 1];
            *reinterpret_cast<double2*>(p.out[k] + i + v) = o;
        }
}

// B: persistent grid-stride with one-iteration register prefetch
__global__ void k_persistent(Ptrs p, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x * 2;
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if(i >= n) return;
    double2 a[3], o[3];
#pragma unroll
    for(int k = 0; k < 3; ++k) { a[k] = *reinterpret_cast<const double2*>(p.in[k] + i); o[k] = *reinterpret_cast<const double2*>(p.out[k] + i); }
    for(; i < n; i += stride)
    {
        double2 an[3], on[3];
        const size_t j = i + stride;
        if(j < n)
        {
#pragma unroll
            for(int k = 0; k < 3; ++k) { an[k] = *reinterpret_cast<const double2*>(p.in[k] + j); on[k] = *reinterpret_cast<const double2*>(p.out[k] + j); }
        }
        const double sx = a[0].x + a[1].x + a[2].x, sy = a[0].y + a[1].y + a[2].y;
#pragma unroll
        for(int k = 0; k < 3; ++k) { o[k].x += sx; o[k].y += sy; *reinterpret_cast<double2*>(p.out[k] + i) = o[k]; }
#pragma unroll
        for(int k = 0; k < 3; ++k) { a[k] = an[k]; o[k] = on[k]; }
    }
}

int main()
{
    const size_t n = (size_t)272 * 258 * 258;   // one padded 256^3 field
    Ptrs p;
    for(int k = 0; k < 6; ++k) { double* d; cudaMalloc(&d, n * 8); cudaMemset(d, 0, n * 8); p.in[k] = d; }
    for(int k = 0; k < 3; ++k) { cudaMalloc(&p.out[k], n * 8); cudaMemset(p.out[k], 0, n * 8); }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = 9.0 * n * 8;   // 3 in read + 3 out read + 3 out written
    auto run = [&](const char* name, auto launch) {
        for(int w = 0; w < 3; ++w) launch();
        cudaEventRecord(e0);
        for(int it = 0; it < 20; ++it) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-34s %8.1f GB/s   (%s)\n", name, bytes * 20 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    for(int tb : {128, 256, 512, 1024})
    {
        char nm[64];
        snprintf(nm, 64, "oneshot vec2 tb=%d", tb); run(nm, [&] { k_oneshot<2><<<(unsigned)((n / 2 + tb - 1) / tb), tb>>>(p, n); });
        snprintf(nm, 64, "oneshot vec4 tb=%d", tb); run(nm, [&] { k_oneshot<4><<<(unsigned)((n / 4 + tb - 1) / tb), tb>>>(p, n); });
        snprintf(nm, 64, "oneshot vec8 tb=%d", tb); run(nm, [&] { k_oneshot<8><<<(unsigned)((n / 8 + tb - 1) / tb), tb>>>(p, n); });
    }
    for(int bps : {1, 2, 4, 8})
        for(int tb : {256, 512, 1024})
        {
            char nm[64];
            snprintf(nm, 64, "persistent tb=%d blocks/SM=%d", tb, bps);
            run(nm, [&] { k_persistent<<<148 * bps, tb>>>(p, n); });
        }
    return 0;
}
