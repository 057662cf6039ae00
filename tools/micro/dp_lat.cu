// Latency of dependent double-precision operations and 64-bit shuffles on one warp (and with 22 warps in flight).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/dp_lat tools/micro/dp_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b, int n) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = fma(x, b, a);
    long long t1 = clock64();
    double y = x;
    for (int i = 0; i < n; ++i) y = y + __shfl_xor_sync(0xffffffffu, y, 1);
    long long t2 = clock64();
    double z = y;
    for (int i = 0; i < n; ++i) z = rsqrt(z + 2.0);
    long long t3 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = z;
    if (threadIdx.x == 0) { cyc[3 * blockIdx.x] = t1 - t0; cyc[3 * blockIdx.x + 1] = t2 - t1; cyc[3 * blockIdx.x + 2] = t3 - t2; }
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 2048); cudaMalloc(&cyc, 8 * 8);
    for (int threads : {32, 704, 1024}) {
        k<<<1, threads>>>(out, cyc, 1.0, 0.999, 1000);
        k<<<1, threads>>>(out, cyc, 1.0, 0.999, 1000);
        long long h[3]; cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
        printf("threads %4d: cycles per dependent DFMA %.1f | SHFL.64 + DADD %.1f | rsqrt(double)+DADD %.1f\n", threads, h[0] / 1000.0, h[1] / 1000.0, h[2] / 1000.0);
    }
    return 0;
}
