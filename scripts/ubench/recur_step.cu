// Microbenchmark: candidate formulations of the compressor recurrence step (dependent chain).
#include <cstdio>
#include <cuda_runtime.h>
#define N 8192
struct E { double M, inc, dec, T; };
template <int OP> __global__ void lat(double *out, long long *cyc, const E *tab, double a0)
{
    double a = a0 + threadIdx.x * 1e-9;
    E e[8];
    for (int k = 0; k < 8; ++k) e[k] = tab[(threadIdx.x + k) & 63];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N / 8; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const E &x = e[k];
            if (OP == 0) {   // v1: fmin/fmax
                bool p = a <= x.M; double u = fmin(a + x.inc, x.M); double d = fmax(a - x.dec, 0.0); a = p ? u : d;
            }
            if (OP == 1) {   // thresholds on a, double compares
                bool p = a <= x.M, q1 = a >= x.T, q2 = a < x.dec;
                double u = a + x.inc, d = a - x.dec;
                double vu = q1 ? x.M : u, vd = q2 ? 0.0 : d;
                a = p ? vu : vd;
            }
            if (OP == 2) {   // thresholds on a, integer compares on the bit patterns
                long long ab = __double_as_longlong(a);
                bool p = ab <= __double_as_longlong(x.M), q1 = ab >= __double_as_longlong(x.T), q2 = ab < __double_as_longlong(x.dec);
                double u = a + x.inc, d = a - x.dec;
                double vu = q1 ? x.M : u, vd = q2 ? 0.0 : d;
                a = p ? vu : vd;
            }
            if (OP == 3) {   // compare after add, no fmin
                bool p = a <= x.M; double u = a + x.inc, d = a - x.dec;
                u = u > x.M ? x.M : u; d = d < 0.0 ? 0.0 : d; a = p ? u : d;
            }
            if (OP == 4) {   // single select: precombined predicate, constants selected early
                long long ab = __double_as_longlong(a);
                bool p = ab <= __double_as_longlong(x.M), q1 = ab >= __double_as_longlong(x.T), q2 = ab < __double_as_longlong(x.dec);
                double addend = p ? x.inc : -x.dec;            // independent of the add
                double s = a + addend;
                bool clampv = p ? q1 : q2;
                double cv = p ? x.M : 0.0;
                a = clampv ? cv : s;
            }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *out; long long *cyc, h; E *tab, ht[64];
    for (int i = 0; i < 64; ++i) { double M = 1.0 + 0.01 * (i % 7); ht[i] = {M, M / 480, M / 9600, M - M / 480}; }
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8); cudaMalloc(&tab, sizeof ht); cudaMemcpy(tab, ht, sizeof ht, cudaMemcpyHostToDevice);
    const char *names[] = {"fmin/fmax", "thresholds, DSETP", "thresholds, int cmp", "cmp after add", "single add + select"};
#define RUN(OP) lat<OP><<<1, 32>>>(out, cyc, tab, 0.5); lat<OP><<<1, 32>>>(out, cyc, tab, 0.5); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-22s %.2f cycles/step\n", names[OP], (double)h / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4)
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
