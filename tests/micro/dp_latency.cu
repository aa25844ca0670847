// Dependent-chain latencies of the FP64 operations the resident GCMC chain's serial section is made of (one thread).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dp_latency dp_latency.cu && ./dp_latency
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <class F> __device__ long long chain(F f, double& x) {
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = f(x);
    return clock64() - t0;
}
__global__ void k(double* out, long long* cyc, double seed) {
    double x = seed; int j = 0;
    cyc[j++] = chain([](double v) { return fma(v, 1.0000001, 1e-9); }, x);
    cyc[j++] = chain([](double v) { return __dadd_rn(v, 1e-9); }, x);
    cyc[j++] = chain([](double v) { return __dmul_rn(v, 1.0000001); }, x);
    cyc[j++] = chain([](double v) { return rint(v * 1.5) * 0.7 + 0.1; }, x);       // DMUL + FRND + DFMA
    cyc[j++] = chain([](double v) { return 1.0 / (v + 2.0); }, x);                 // DADD + division
    cyc[j++] = chain([](double v) { return exp(-v * 0.5 - 0.1); }, x);
    cyc[j++] = chain([](double v) { return (double)(__double2ll_rz(v * 1000.0) & 1023) * 0.001 + 0.3; }, x);  // D2I + I2D
    cyc[j++] = chain([](double v) { return floor(v * 3.3) * 0.1 + 0.05; }, x);
    float y = (float)x;
    { long long t0 = clock64();
#pragma unroll 16
      for (int i = 0; i < N; ++i) y = fmaf(y, 1.0000001f, 1e-9f);
      cyc[j++] = clock64() - t0; }
    unsigned u = (unsigned)(y * 1000.f) + 12345u;
    { long long t0 = clock64();
#pragma unroll 16
      for (int i = 0; i < N; ++i) { u ^= u >> 11; u ^= (u << 7) & 0x9d2c5680u; u ^= (u << 15) & 0xefc60000u; u ^= u >> 18; }
      cyc[j++] = clock64() - t0; }
    extern __shared__ unsigned sm[];
    for (int i = 0; i < 1024; ++i) sm[i] = (i * 7 + 1) & 1023;
    { long long t0 = clock64();
      unsigned p = u & 1023;
#pragma unroll 16
      for (int i = 0; i < N; ++i) p = sm[p];
      cyc[j++] = clock64() - t0; u += p; }
    { long long t0 = clock64();
      for (int i = 0; i < 256; ++i) __syncthreads();
      cyc[j++] = clock64() - t0; }
    out[0] = x + y + u;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * 16);
    const char* names[] = {"DFMA", "DADD", "DMUL", "DMUL+rint+DFMA", "DADD+1/x", "exp(DFMA)", "DMUL+D2I+I2D+DFMA", "DMUL+floor+DFMA", "FFMA", "MT temper (12 int ops)", "LDS chase", "__syncthreads x256 (1 thread)"};
    for (int threads : {1, 544}) {
        k<<<1, threads, 4096>>>(out, cyc, 0.37);
        long long h[16];
        cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed\n"); return 1; }
        printf("threads per CTA = %d\n", threads);
        for (int j = 0; j < 12; ++j) printf("  %-32s %8.1f cycles per step\n", names[j], (double)h[j] / (j == 11 ? 256 : N));
    }
    return 0;
}
