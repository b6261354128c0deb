// Microbenchmark: fp64 / fp32 dependent-chain latency and per-SM throughput on B200.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP> __global__ void lat(double *out, long long *cyc, double a, double b)
{
    double x = a + threadIdx.x * 1e-9, y = b;
    float xf = (float)x, yf = (float)b;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = fma(x, y, b);                       // DFMA chain
        if (OP == 1) x = x + y;                              // DADD chain
        if (OP == 2) x = fmin(x + y, b);                     // DADD + DMNMX
        if (OP == 3) { bool p = x <= y; double u = fmin(x + 1e-3, y); double d = fmax(x - 1e-4, 0.0); x = p ? u : d; } // recurrence step
        if (OP == 4) xf = fmaf(xf, yf, yf);                  // FFMA chain
        if (OP == 5) x = x * y;                              // DMUL chain
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x + xf;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// throughput: ILP independent chains per thread, many warps
template <int ILP> __global__ void thr(double *out, long long *cyc, double a, double b)
{
    double x[ILP];
    for (int k = 0; k < ILP; ++k) x[k] = a + k + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < N / 4; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], b, a);
    }
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < ILP; ++k) s += x[k];
    out[threadIdx.x + blockIdx.x * blockDim.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
    const char *names[] = {"DFMA chain", "DADD chain", "DADD+fmin chain", "recur step", "FFMA chain", "DMUL chain"};
#define RUN_LAT(OP) lat<OP><<<1, 32>>>(out, cyc, 1.0, 0.999); lat<OP><<<1, 32>>>(out, cyc, 1.0, 0.999); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-18s %.2f cycles/iter\n", names[OP], (double)h / N);
    RUN_LAT(0) RUN_LAT(1) RUN_LAT(2) RUN_LAT(3) RUN_LAT(4) RUN_LAT(5)
    for (int warps = 1; warps <= 32; warps *= 2) {
        thr<4><<<1, 32 * warps>>>(out, cyc, 1.0, 0.999); thr<4><<<1, 32 * warps>>>(out, cyc, 1.0, 0.999);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("1 SM, %2d warps x ILP4: %.2f DFMA lanes/clk/SM\n", warps, (double)N * 4 * 32 * warps / h);
    }
    thr<8><<<1, 1024>>>(out, cyc, 1.0, 0.999); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("1 SM, 32 warps x ILP8: %.2f DFMA lanes/clk/SM\n", (double)N * 8 * 1024 / h);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    thr<8><<<148 * 2, 1024>>>(out, cyc, 1.0, 0.999);
    cudaEventRecord(a); thr<8><<<148 * 2, 1024>>>(out, cyc, 1.0, 0.999); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("full chip: %.2f TFLOP/s fp64 (FMA=2)\n", 2.0 * N * 8 * 1024 * 148 * 2 / (ms * 1e-3) / 1e12);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
