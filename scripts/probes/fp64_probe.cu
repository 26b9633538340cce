// Probe (GPU box only): FP64 pipe of one SM -- DFMA throughput as a function of resident warps and independent
// chains per thread.  Not part of the library.
#include <cuda_runtime.h>
#include <cstdio>
template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 1e-3 + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) x[j] = __fma_rn(x[j], a, b);
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
void run(int warps_per_sm, double* d, int sms, double clk_ghz) {
    const int iters = 4096;
    dim3 block(32 * (warps_per_sm > 32 ? 32 : warps_per_sm));
    dim3 grid(sms * (warps_per_sm > 32 ? warps_per_sm / 32 : 1));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_kernel<ILP><<<grid, block>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e0);
    dfma_kernel<ILP><<<grid, block>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cycles = ms * 1e-3 * clk_ghz * 1e9;
    double warp_instr_per_sm = (double)iters * ILP * warps_per_sm;
    printf("warps/SM %2d  ILP %d : %.3f warp-DFMA / clk / SM   (%.1f cycles per dependent DFMA per warp)\n", warps_per_sm, ILP,
           warp_instr_per_sm / cycles, cycles / iters);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double clk = p.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz\n", p.name, sms, clk);
    double* d; cudaMalloc(&d, sizeof(double) * sms * 2048 * 2);
    for (int w : {4, 8, 16, 24, 32, 64}) { run<1>(w, d, sms, clk); run<2>(w, d, sms, clk); run<4>(w, d, sms, clk); run<8>(w, d, sms, clk); }
    return 0;
}
