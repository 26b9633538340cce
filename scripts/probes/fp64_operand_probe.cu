// Probe (GPU box only): does the FP64 pipe's issue rate depend on where the operands come from?  DFMA / DADD / DMUL
// with one, two or three distinct 64-bit REGISTER operands (the blobness kernel's instructions read two or three),
// 16 resident warps per SM, 8 independent chains per thread.  Not part of the library.
#include <cuda_runtime.h>
#include <cstdio>
enum { FMA_1R, FMA_2R, FMA_3R, ADD_2R, MUL_2R, MIX_3R };
template <int OP>
__global__ void k(double* out, const double* in, int iters) {
    double x[8], y[8], z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { x[j] = in[threadIdx.x + j]; y[j] = in[64 + threadIdx.x + j]; z[j] = in[128 + threadIdx.x + j]; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (OP == FMA_1R) x[j] = __fma_rn(x[j], 0.999999, 1e-9);
            if (OP == FMA_2R) x[j] = __fma_rn(x[j], y[j], 1e-9);
            if (OP == FMA_3R) x[j] = __fma_rn(y[j], z[j], x[j]);
            if (OP == ADD_2R) x[j] = __dadd_rn(x[j], y[j]);
            if (OP == MUL_2R) x[j] = __dmul_rn(x[j], y[j]);
            if (OP == MIX_3R) { x[j] = __fma_rn(y[j], z[j], x[j]); y[j] = __dadd_rn(y[j], z[(j + 1) & 7]); z[j] = __dmul_rn(z[j], x[(j + 3) & 7]); }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j] + y[j] + z[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, double per_iter, double* d, const double* in, int sms, double clk) {
    const int iters = 1024;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps : {8, 16, 32}) {
        dim3 block(32 * (warps > 16 ? 16 : warps)), grid(sms * (warps > 16 ? warps / 16 : 1));
        k<OP><<<grid, block>>>(d, in, iters);
        cudaEventRecord(e0);
        k<OP><<<grid, block>>>(d, in, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double cycles = ms * 1e-3 * clk * 1e9;
        printf("%-34s warps/SM %2d : %.3f warp-instr / clk / SM\n", name, warps, iters * 8.0 * warps * per_iter / cycles);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double clk = p.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz, 8 chains per thread\n", p.name, sms, clk);
    double *d, *in; cudaMalloc(&d, sizeof(double) * sms * 2 * 512); cudaMalloc(&in, sizeof(double) * 1024);
    cudaMemset(in, 0, sizeof(double) * 1024);
    run<FMA_1R>("DFMA  x = fma(x, c, c)", 1, d, in, sms, clk);
    run<FMA_2R>("DFMA  x = fma(x, y, c)", 1, d, in, sms, clk);
    run<FMA_3R>("DFMA  x = fma(y, z, x)", 1, d, in, sms, clk);
    run<ADD_2R>("DADD  x = x + y", 1, d, in, sms, clk);
    run<MUL_2R>("DMUL  x = x * y", 1, d, in, sms, clk);
    run<MIX_3R>("DFMA + DADD + DMUL, register ops", 3, d, in, sms, clk);
    return 0;
}
