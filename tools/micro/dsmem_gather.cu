// Microbenchmark: random 8-byte gathers from distributed shared memory (cluster of CS CTAs, one slice each)
// versus random 32-byte sector gathers from global/L2.  Decides whether a DSMEM-resident vector beats the
// L1TEX divergent-gather limit for SpMV on an expander graph.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

template <int UNROLL>
__global__ void k_dsmem(int slice, int iters, double* out, long long* cyc) {
    extern __shared__ double sm[];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned cs = cl.num_blocks();
    for (int i = threadIdx.x; i < slice; i += blockDim.x) sm[i] = (double)(i + cl.block_rank());
    cl.sync();
    const unsigned n = slice * cs;
    unsigned r = hash32(blockIdx.x * blockDim.x + threadIdx.x + 1);
    double acc = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            r = hash32(r + u);
            unsigned idx = r % n;
            unsigned rank = idx / slice, off = idx % slice;
            const double* p = cl.map_shared_rank(sm, rank);
            v[u] = p[off];
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u];
    }
    long long t1 = clock64();
    cl.sync();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 123.456) out[0] = acc;
}

template <int UNROLL>
__global__ void k_global(const double* __restrict__ vec, unsigned n, int iters, double* out, long long* cyc) {
    unsigned r = hash32(blockIdx.x * blockDim.x + threadIdx.x + 1);
    double acc = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double a[UNROLL], b[UNROLL], c[UNROLL], d[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            r = hash32(r + u);
            unsigned idx = r % n;
            asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a[u]), "=d"(b[u]), "=d"(c[u]), "=d"(d[u]) : "l"(vec + 4 * (size_t)idx));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += a[u] + b[u] + c[u];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 123.456) out[0] = acc;
}

int main() {
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    printf("%s SMs=%d\n", prop.name, prop.multiProcessorCount);
    double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * 4096);
    const int threads = 1024, iters = 64;
    for (int cs : {4, 8, 16}) {
        int n = 100000; int slice = (n + cs - 1) / cs;
        size_t smem = slice * sizeof(double);
        auto kern = k_dsmem<4>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cs > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim = {(unsigned)cs, 1, 1};
        cfg.attrs = attr; cfg.numAttrs = 1;
        int maxc = 0; cfg.gridDim = dim3(cs);
        cudaOccupancyMaxActiveClusters(&maxc, kern, &cfg);
        int nclusters = maxc; if (nclusters < 1) { printf("cs=%d: cannot launch (%s)\n", cs, cudaGetErrorString(cudaGetLastError())); continue; }
        cfg.gridDim = dim3(cs * nclusters);
        for (int rep = 0; rep < 2; ++rep) {
            cudaError_t e = cudaLaunchKernelEx(&cfg, kern, slice, iters, out, cyc);
            if (e != cudaSuccess) { printf("cs=%d launch failed %s\n", cs, cudaGetErrorString(e)); break; }
            e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("cs=%d run failed %s\n", cs, cudaGetErrorString(e)); break; }
        }
        std::vector<long long> h(cs * nclusters); cudaMemcpy(h.data(), cyc, 8 * h.size(), cudaMemcpyDeviceToHost);
        double mean = 0; for (auto v : h) mean += v; mean /= h.size();
        double gathers_per_cta = (double)threads * iters * 4;
        printf("DSMEM cs=%2d clusters=%3d (SMs used %3d) slice=%6d doubles: %.0f cycles -> %.3f gathers/clk/SM\n", cs, nclusters, cs * nclusters, slice, mean, gathers_per_cta / mean);
    }
    {   // global sector gathers, 1 CTA of 1024 threads per SM
        unsigned n = 100000; double* vec; cudaMalloc(&vec, 32 * (size_t)n); cudaMemset(vec, 0, 32 * (size_t)n);
        for (int rep = 0; rep < 2; ++rep) { k_global<4><<<prop.multiProcessorCount, threads>>>(vec, n, iters, out, cyc); cudaDeviceSynchronize(); }
        std::vector<long long> h(prop.multiProcessorCount); cudaMemcpy(h.data(), cyc, 8 * h.size(), cudaMemcpyDeviceToHost);
        double mean = 0; for (auto v : h) mean += v; mean /= h.size();
        printf("global 32B sector gathers (L2-resident 3.2 MB): %.0f cycles -> %.3f gathers/clk/SM\n", mean, (double)threads * iters * 4 / mean);
        for (int rep = 0; rep < 2; ++rep) { k_global<4><<<2 * prop.multiProcessorCount, threads / 2>>>(vec, n, iters * 2, out, cyc); cudaDeviceSynchronize(); }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
