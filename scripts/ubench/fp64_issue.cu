// Does a DFMA occupy the warp scheduler's issue slot for two cycles, or can integer / fp32 work issue in
// its shadow?  16 warps per SM (4 per scheduler), fully independent instruction streams per thread:
//   A: 8 DFMA                per iteration
//   B: 8 DFMA + 8 IMAD       per iteration
//   C: 8 DFMA + 8 FFMA       per iteration
//   D: 16 IMAD               per iteration
// cycles per iteration per scheduler tell which resource is shared.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(double *out, long long *cyc, int iters, double c, int ci, float cf)
{
    double d[8]; int q[16]; float f[8];
    for (int i = 0; i < 8; ++i) { d[i] = threadIdx.x * 1e-3 + i; f[i] = threadIdx.x * 1e-3f + i; }
    for (int i = 0; i < 16; ++i) q[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE != 3) d[i] = fma(d[i], c, 1.0);
            if (MODE == 1) q[i] = q[i] * ci + 3;
            if (MODE == 2) f[i] = fmaf(f[i], cf, 1.0f);
            if (MODE == 3) { q[i] = q[i] * ci + 3; q[i + 8] = q[i + 8] * ci + 5; }
        }
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; ++i) s += d[i] + f[i]; for (int i = 0; i < 16; ++i) s += q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *out; long long *cyc, h; cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    const char *names[] = {"8 DFMA", "8 DFMA + 8 IMAD", "8 DFMA + 8 FFMA", "16 IMAD"};
#define RUN(M) k<M><<<148, 512>>>(out, cyc, iters, 0.999, 3, 0.999f); k<M><<<148, 512>>>(out, cyc, iters, 0.999, 3, 0.999f); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-18s %.2f cycles per iteration (4 warps per scheduler: %.2f per warp-iteration)\n", names[M], (double)h / iters, (double)h / iters / 4);
    RUN(0) RUN(1) RUN(2) RUN(3)
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
