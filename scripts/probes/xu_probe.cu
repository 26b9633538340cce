// Probe (GPU box only): throughput of the 64-bit special-function / conversion instructions the f64 blobness kernel
// uses (MUFU.RSQ64H, MUFU.RCP64H, I2F.F64, F2F) next to DFMA, per SM, 16 resident warps, 4 independent chains.
// Not part of the library.
#include <cuda_runtime.h>
#include <cstdio>
enum { OP_DFMA, OP_RSQ64, OP_RCP64, OP_I2D, OP_MAGIC, OP_D2F2D, OP_MIX };
template <int OP>
__global__ void k(double* out, int iters, double a, double b) {
    double x[4];
    unsigned u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { x[j] = 1.0 + threadIdx.x * 1e-3 + j; u[j] = threadIdx.x + j; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (OP == OP_DFMA) x[j] = __fma_rn(x[j], a, b);
            if (OP == OP_RSQ64) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(x[j]));
            if (OP == OP_RCP64) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(x[j]));
            if (OP == OP_I2D) { asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(x[j]) : "r"(u[j])); u[j] = __double2loint(x[j]) + 1; }
            if (OP == OP_MAGIC) { x[j] = __hiloint2double(0x43300000, u[j]) - 4503599627370496.0; u[j] = __double2loint(x[j]) + 1; }
            if (OP == OP_D2F2D) { float f; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(x[j])); asm volatile("cvt.f64.f32 %0, %1;" : "=d"(x[j]) : "f"(f)); }
            if (OP == OP_MIX) {       // 16 DFMA per MUFU.RSQ64H: the kernel's mix in stage D
                double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[j]));
#pragma unroll
                for (int q = 0; q < 16; ++q) x[j] = __fma_rn(x[j], a, b);
                x[j] += y * 1e-30;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, double per_iter, double* d, int sms, double clk) {
    const int iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<sms, 512>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e0);
    k<OP><<<sms, 512>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cycles = ms * 1e-3 * clk * 1e9;
    printf("%-28s %.3f warp-instr / clk / SM  (%.1f cycles per warp-instr per SM)\n", name,
           iters * 4.0 * 16 * per_iter / cycles, cycles / (iters * 4.0 * 16 * per_iter));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double clk = p.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz, 16 warps / SM, 4 chains per thread\n", p.name, sms, clk);
    double* d; cudaMalloc(&d, sizeof(double) * sms * 512);
    run<OP_DFMA>("DFMA", 1, d, sms, clk);
    run<OP_RSQ64>("MUFU.RSQ64H", 1, d, sms, clk);
    run<OP_RCP64>("MUFU.RCP64H", 1, d, sms, clk);
    run<OP_I2D>("I2F.F64.U32 (+F2I back)", 1, d, sms, clk);
    run<OP_MAGIC>("2^52 magic u32->f64 (DADD)", 1, d, sms, clk);
    run<OP_D2F2D>("F2F f64->f32->f64 (pair)", 1, d, sms, clk);
    run<OP_MIX>("16 DFMA + 1 RSQ64H (17 instr)", 17, d, sms, clk);
    return 0;
}
