// What do the conversions around the fp64 filters cost?  The chain converts int16 -> double (I2F.F64), double ->
// int (F2I.F64, the quantisers and audioop.mul's floor) and float <-> double (F2F) several times per sample; if
// those run at a fraction of the DFMA rate they, not the FMAs, set the fp64 roof.  16 warps per SM (4 per
// scheduler), eight independent streams per thread; every mode is timed next to a baseline with the same
// surrounding instructions, so the difference is the conversion alone.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o cvt_rate cvt_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
enum { DADD_ONLY, I2F64, MAGIC_I2F64, F2I64_BASE, F2I64, F2I64_FLOOR, MAGIC_F2I64, F2F_32_64_BASE, F2F_32_64, F2F_64_32_BASE, F2F_64_32, DSETP_SEL, I2F32, F2I32, NMODES };
template <int MODE> __global__ void k(double *out, long long *cyc, int iters, double c, int ci, float cf)
{
    double d[8], acc[8]; int q[8]; float f[8];
    for (int i = 0; i < 8; ++i) { d[i] = threadIdx.x * 1e-3 + i; acc[i] = i; f[i] = threadIdx.x * 1e-3f + i; q[i] = threadIdx.x + i; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == DADD_ONLY) acc[i] = __dadd_rn(acc[i], __hiloint2double(0x3ff00000, q[i] ^ it));
            if (MODE == I2F64) acc[i] = __dadd_rn(acc[i], (double)(q[i] ^ it));
            if (MODE == MAGIC_I2F64) acc[i] = __dadd_rn(acc[i], __dsub_rn(__hiloint2double(0x43300000, q[i] ^ it), 4503601774854144.0));
            if (MODE == F2I64_BASE) { d[i] = __dadd_rn(d[i], c); q[i] += __double2loint(d[i]); }
            if (MODE == F2I64) { d[i] = __dadd_rn(d[i], c); q[i] += __double2int_rz(d[i]); }
            if (MODE == F2I64_FLOOR) { d[i] = __dadd_rn(d[i], c); q[i] += __double2int_rd(d[i]); }
            if (MODE == MAGIC_F2I64) { d[i] = __dadd_rn(d[i], c); q[i] += __double2loint(__dadd_rd(d[i], 6755399441055744.0)); }
            if (MODE == F2F_32_64_BASE) { f[i] = __fadd_rn(f[i], cf); acc[i] = __dadd_rn(acc[i], __hiloint2double(0x3ff00000, __float_as_int(f[i]))); }
            if (MODE == F2F_32_64) { f[i] = __fadd_rn(f[i], cf); acc[i] = __dadd_rn(acc[i], (double)f[i]); }
            if (MODE == F2F_64_32_BASE) { d[i] = __dadd_rn(d[i], c); f[i] = __fadd_rn(f[i], __int_as_float(__double2hiint(d[i]))); }
            if (MODE == F2F_64_32) { d[i] = __dadd_rn(d[i], c); f[i] = __fadd_rn(f[i], (float)d[i]); }
            if (MODE == DSETP_SEL) { d[i] = __dadd_rn(d[i], c); q[i] += (d[i] <= acc[i]) ? ci : it; }
            if (MODE == I2F32) f[i] = __fadd_rn(f[i], (float)(q[i] ^ it));
            if (MODE == F2I32) { f[i] = __fadd_rn(f[i], cf); q[i] += __float2int_rz(f[i]); }
        }
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; ++i) s += d[i] + f[i] + acc[i] + q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int M> void run(const char *name, double *out, long long *cyc)
{
    const int iters = 20000; long long h;
    k<M><<<148, 512>>>(out, cyc, iters, 0.999, 3, 0.999f); k<M><<<148, 512>>>(out, cyc, iters, 0.999, 3, 0.999f);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-52s %7.2f cycles per iteration of 8 (4 warps per scheduler)\n", name, (double)h / iters);
}
int main()
{
    double *out; long long *cyc; cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 8);
    run<DADD_ONLY>("LOP3 + DADD (baseline for the next two)", out, cyc);
    run<I2F64>("LOP3 + I2F.F64.S32 + DADD", out, cyc);
    run<MAGIC_I2F64>("LOP3 + DADD(2^52 magic) + DADD", out, cyc);
    run<F2I64_BASE>("DADD + IADD (baseline for the next three)", out, cyc);
    run<F2I64>("DADD + F2I.S32.F64.TRUNC + IADD", out, cyc);
    run<F2I64_FLOOR>("DADD + F2I.S32.F64.FLOOR + IADD", out, cyc);
    run<MAGIC_F2I64>("DADD + DADD.RM(1.5 * 2^52) + IADD", out, cyc);
    run<F2F_32_64_BASE>("FADD + DADD (baseline)", out, cyc);
    run<F2F_32_64>("FADD + F2F.F64.F32 + DADD", out, cyc);
    run<F2F_64_32_BASE>("DADD + FADD (baseline)", out, cyc);
    run<F2F_64_32>("DADD + F2F.F32.F64 + FADD", out, cyc);
    run<DSETP_SEL>("DADD + DSETP + SEL + IADD", out, cyc);
    run<I2F32>("LOP3 + I2F.F32.S32 + FADD", out, cyc);
    run<F2I32>("FADD + F2I.S32.F32 + IADD", out, cyc);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
