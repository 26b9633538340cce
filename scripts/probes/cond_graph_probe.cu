// Probe (GPU box only): CUDA conditional graph nodes on this driver -- WHILE with a cooperative kernel inside,
// nested IF inside WHILE, per-iteration overhead vs unrolled kernel nodes.  Not part of the library.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void coop_body(int* c) {
    cg::grid_group grid = cg::this_grid();
    if (blockIdx.x == 0 && threadIdx.x == 0) c[1] += 1;
    grid.sync();
    if (threadIdx.x == 0) atomicAdd(&c[2], 1);
    grid.sync();
}
__global__ void ctl_top(int* c, cudaGraphConditionalHandle h_if) {
    if (threadIdx.x == 0) cudaGraphSetConditional(h_if, (c[0] % 3 == 1) ? 1u : 0u);
}
__global__ void if_body(int* c) { if (threadIdx.x == 0) c[3] += 1; }
__global__ void ctl_end(int* c, int n, cudaGraphConditionalHandle h_while) {
    if (threadIdx.x == 0) { int v = ++c[0]; cudaGraphSetConditional(h_while, v < n ? 1u : 0u); }
}
__global__ void tiny(int* c) { if (threadIdx.x == 0 && c[7] == 12345) c[6] += 1; }
__global__ void set_cond(cudaGraphConditionalHandle h, unsigned v) { if (threadIdx.x == 0) cudaGraphSetConditional(h, v); }

static int add_cond(cudaStream_t s, cudaGraphConditionalNodeType type, cudaGraph_t handle_graph_override,
                    cudaGraphConditionalHandle* h, cudaGraph_t* body, unsigned defval) {
    cudaStreamCaptureStatus st; cudaGraph_t g; const cudaGraphNode_t* deps; size_t nd;
    CK(cudaStreamGetCaptureInfo(s, &st, nullptr, &g, &deps, &nd));
    CK(cudaGraphConditionalHandleCreate(h, handle_graph_override ? handle_graph_override : g, defval, cudaGraphCondAssignDefault));
    cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional; p.conditional.handle = *h;
    p.conditional.type = type; p.conditional.size = 1;
    cudaGraphNode_t node;
    CK(cudaGraphAddNode(&node, g, deps, nd, &p));
    CK(cudaStreamUpdateCaptureDependencies(s, &node, 1, cudaStreamSetCaptureDependencies));
    *body = p.conditional.phGraph_out[0];
    return 0;
}

int run_variant(int variant, int N) {
    // variant 0: nested IF handle created on the body graph; 1: on the top-level graph
    int* d; CK(cudaMalloc(&d, 64)); CK(cudaMemset(d, 0, 64));
    cudaStream_t s, s2, s3; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking));
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    tiny<<<1, 32, 0, s>>>(d);
    cudaGraph_t top; { cudaStreamCaptureStatus st; CK(cudaStreamGetCaptureInfo(s, &st, nullptr, &top, nullptr, nullptr)); }
    cudaGraphConditionalHandle hw, hi; cudaGraph_t wbody, ibody;
    if (add_cond(s, cudaGraphCondTypeWhile, nullptr, &hw, &wbody, 1)) return 1;
    CK(cudaStreamBeginCaptureToGraph(s2, wbody, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    {
        int nblk = 296; void* args[] = {&d};
        CK(cudaLaunchCooperativeKernel((void*)coop_body, dim3(nblk), dim3(256), args, 0, s2));
        if (variant == 0) { if (add_cond(s2, cudaGraphCondTypeIf, nullptr, &hi, &ibody, 0)) return 1; }
        else { if (add_cond(s2, cudaGraphCondTypeIf, top, &hi, &ibody, 0)) return 1; }
        // ctl_top must run BEFORE the IF node: re-do in the right order below
    }
    // the IF node has been added after the coop kernel; its handle is set by ctl_top which must precede it: so
    // capture order is coop -> IF.  ctl_top is put at the very top of the NEXT iteration's path instead: simpler to
    // set the IF condition from ctl_end of the previous iteration and from a pre-loop kernel.
    CK(cudaStreamBeginCaptureToGraph(s3, ibody, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    if_body<<<1, 32, 0, s3>>>(d);
    CK(cudaStreamEndCapture(s3, nullptr));
    ctl_top<<<1, 32, 0, s2>>>(d, hi);        // sets the IF condition for the NEXT iteration (c[0] not yet incremented)
    ctl_end<<<1, 32, 0, s2>>>(d, N, hw);
    CK(cudaStreamEndCapture(s2, nullptr));
    tiny<<<1, 32, 0, s>>>(d);
    cudaGraph_t g; CK(cudaStreamEndCapture(s, &g));
    cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemsetAsync(d, 0, 64, s));
        CK(cudaEventRecord(e0, s)); CK(cudaGraphLaunch(ex, s)); CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        int h[8]; CK(cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost));
        printf("variant %d rep %d: it=%d coop_runs=%d blocks_seen=%d if_runs=%d  %.1f us total, %.2f us/iter\n",
               variant, rep, h[0], h[1], h[2], h[3], ms * 1e3, ms * 1e3 / N);
    }
    cudaGraphExecDestroy(ex); cudaGraphDestroy(g);
    return 0;
}

int overhead(int N) {
    int* d; CK(cudaMalloc(&d, 64)); CK(cudaMemset(d, 0, 64));
    cudaStream_t s, s2; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int kernels_per_iter : {1, 8}) {
        // (a) unrolled
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < N * kernels_per_iter; ++i) tiny<<<148, 256, 0, s>>>(d);
        cudaGraph_t g; CK(cudaStreamEndCapture(s, &g));
        cudaGraphExec_t ex; CK(cudaGraphInstantiate(&ex, g, 0));
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0, s)); CK(cudaGraphLaunch(ex, s)); CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("unrolled   %d iters x %d kernels: %.1f us (%.2f us/kernel)\n", N, kernels_per_iter, best * 1e3, best * 1e3 / (N * kernels_per_iter));
        cudaGraphExecDestroy(ex); cudaGraphDestroy(g);
        // (b) WHILE
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        tiny<<<1, 32, 0, s>>>(d);
        cudaGraphConditionalHandle hw; cudaGraph_t wbody;
        if (add_cond(s, cudaGraphCondTypeWhile, nullptr, &hw, &wbody, 1)) return 1;
        CK(cudaStreamBeginCaptureToGraph(s2, wbody, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < kernels_per_iter - 1; ++i) tiny<<<148, 256, 0, s2>>>(d);
        ctl_end<<<1, 32, 0, s2>>>(d, N, hw);
        CK(cudaStreamEndCapture(s2, nullptr));
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ex, g, 0));
        best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaMemsetAsync(d, 0, 64, s));
            CK(cudaEventRecord(e0, s)); CK(cudaGraphLaunch(ex, s)); CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("WHILE node %d iters x %d kernels: %.1f us (%.2f us/iter)\n", N, kernels_per_iter, best * 1e3, best * 1e3 / N);
        cudaGraphExecDestroy(ex); cudaGraphDestroy(g);
        // (c) WHILE + 2 IF nodes per iteration (both false)
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        tiny<<<1, 32, 0, s>>>(d);
        if (add_cond(s, cudaGraphCondTypeWhile, nullptr, &hw, &wbody, 1)) return 1;
        CK(cudaStreamBeginCaptureToGraph(s2, wbody, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < kernels_per_iter - 1; ++i) tiny<<<148, 256, 0, s2>>>(d);
        for (int j = 0; j < 2; ++j) {
            cudaGraphConditionalHandle hi; cudaGraph_t ib;
            if (add_cond(s2, cudaGraphCondTypeIf, nullptr, &hi, &ib, 0)) return 1;
            cudaStream_t s3; CK(cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking));
            CK(cudaStreamBeginCaptureToGraph(s3, ib, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
            tiny<<<148, 256, 0, s3>>>(d);
            CK(cudaStreamEndCapture(s3, nullptr));
        }
        ctl_end<<<1, 32, 0, s2>>>(d, N, hw);
        CK(cudaStreamEndCapture(s2, nullptr));
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ex, g, 0));
        best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaMemsetAsync(d, 0, 64, s));
            CK(cudaEventRecord(e0, s)); CK(cudaGraphLaunch(ex, s)); CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("WHILE+2 IF(false) %d iters x %d kernels: %.1f us (%.2f us/iter)\n", N, kernels_per_iter, best * 1e3, best * 1e3 / N);
        cudaGraphExecDestroy(ex); cudaGraphDestroy(g);
    }
    return 0;
}

int main() {
    int rv = 0;
    rv |= run_variant(0, 10);
    cudaGetLastError();
    rv |= run_variant(1, 10);
    cudaGetLastError();
    rv |= overhead(30);
    printf("probe rc %d\n", rv);
    return 0;
}
