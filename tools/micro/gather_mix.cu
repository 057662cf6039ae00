// Microbenchmark: random 32-byte sector gathers (ld.global.cg.v4.f64) mixed with a re-read streaming region per SM
// (12 bytes per gather: the col/val stream of a CSR SpMV), streamed either from global/L1 or from shared memory.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// MODE 0: gathers only (indices from ALU). MODE 1: indices loaded from a per-CTA global region (ld.global.nc) + 8-byte value.
// MODE 2: indices + values from shared memory (region copied once).
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(const double* __restrict__ vec, unsigned n, const int* __restrict__ cols, const double* __restrict__ vals,
                  int per_cta, int phases, double* out, long long* cyc) {
    extern __shared__ unsigned char smraw[];
    int* scol = (int*)smraw; double* sval = (double*)(smraw + ((per_cta * 4 + 15) / 16) * 16);
    const int* mycol = cols + (size_t)blockIdx.x * per_cta; const double* myval = vals + (size_t)blockIdx.x * per_cta;
    if (MODE == 2) { for (int i = threadIdx.x; i < per_cta; i += blockDim.x) { scol[i] = mycol[i]; sval[i] = myval[i]; } __syncthreads(); }
    double acc = 0.0; unsigned r = hash32(blockIdx.x * 1024 + threadIdx.x + 1);
    long long t0 = clock64();
    for (int ph = 0; ph < phases; ++ph) {
        for (int base = threadIdx.x; base < per_cta; base += 4 * 1024) {
            int c[4]; double w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int i = base + u * 1024; bool ok = i < per_cta; int ii = ok ? i : base;
                if (MODE == 0) { r = hash32(r + u); c[u] = r % n; w[u] = ok ? 1.0 : 0.0; }
                else if (MODE == 1) { c[u] = __ldg(mycol + ii); w[u] = ok ? __ldg(myval + ii) : 0.0; }
                else { c[u] = scol[ii]; w[u] = ok ? sval[ii] : 0.0; }
            }
            double a[4], b[4], cc[4], d[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a[u]), "=d"(b[u]), "=d"(cc[u]), "=d"(d[u]) : "l"(vec + 4 * (size_t)c[u]));
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += w[u] * (a[u] + b[u] + cc[u]);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 123.456) out[0] = acc;
}

int main() {
    cudaSetDevice(0); cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount; const unsigned n = 100000; const int per_cta = 14900, phases = 40;
    double *vec, *vals, *out; int* cols; long long* cyc;
    cudaMalloc(&vec, 32 * (size_t)n); cudaMemset(vec, 0, 32 * (size_t)n);
    std::vector<int> hc((size_t)sms * per_cta); std::vector<double> hv(hc.size(), 1.0);
    unsigned r = 12345; for (auto& c : hc) { r = r * 1664525u + 1013904223u; c = (r >> 8) % n; }
    cudaMalloc(&cols, 4 * hc.size()); cudaMalloc(&vals, 8 * hv.size()); cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * sms);
    cudaMemcpy(cols, hc.data(), 4 * hc.size(), cudaMemcpyHostToDevice); cudaMemcpy(vals, hv.data(), 8 * hv.size(), cudaMemcpyHostToDevice);
    size_t smem = ((per_cta * 4 + 15) / 16) * 16 + (size_t)per_cta * 8;
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) k<0><<<sms, 1024>>>(vec, n, cols, vals, per_cta, phases, out, cyc);
            if (mode == 1) k<1><<<sms, 1024>>>(vec, n, cols, vals, per_cta, phases, out, cyc);
            if (mode == 2) k<2><<<sms, 1024, smem>>>(vec, n, cols, vals, per_cta, phases, out, cyc);
            cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) printf("mode %d: %s\n", mode, cudaGetErrorString(e));
        }
        std::vector<long long> h(sms); cudaMemcpy(h.data(), cyc, 8 * sms, cudaMemcpyDeviceToHost);
        double mean = 0; for (auto v : h) mean += v; mean /= sms;
        printf("mode %d (%s): %.0f cycles/phase -> %.3f gathers/clk/SM\n", mode, mode == 0 ? "ALU indices" : mode == 1 ? "col/val from global (L1/L2)" : "col/val from shared memory",
               mean / phases, (double)per_cta * phases / mean);
    }
    return 0;
}
