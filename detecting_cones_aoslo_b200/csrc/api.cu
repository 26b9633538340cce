// api.cu -- context management, the C ABI wrappers of the stage kernels and the whole-run
// driver `aoslo_run` (reference pipeline_with_delitions.py:12-403 after crop / pad).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace aoslo {

static thread_local char g_last_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_last_err, sizeof(g_last_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
    cudaGetLastError();
    return AOSLO_ERR_CUDA;
}

template <typename T>
static int dmalloc(T** p, size_t count) {
    AOSLO_CUDA(cudaMalloc((void**)p, sizeof(T) * (count ? count : 1)));
    return AOSLO_OK;
}

static int alloc_cell_list(CellList& cl, int H, int W, int cap, int shift) {
    cl.shift = shift;
    cl.gw = (W + (1 << shift) - 1) >> shift;
    cl.gh = (H + (1 << shift) - 1) >> shift;
    cl.ncell = cl.gw * cl.gh;
    cl.cap = cap;
    AOSLO_TRY(dmalloc(&cl.count, cl.ncell));
    AOSLO_TRY(dmalloc(&cl.start, cl.ncell + 1));
    AOSLO_TRY(dmalloc(&cl.slot, cap));
    AOSLO_TRY(dmalloc(&cl.cell_of, cap));
    AOSLO_TRY(dmalloc(&cl.sorted_idx, cap));
    AOSLO_TRY(dmalloc(&cl.sorted_pos, cap));
    AOSLO_CUDA(cudaMemset(cl.start, 0, sizeof(int) * (cl.ncell + 1)));
    AOSLO_CUDA(cudaMemset(cl.count, 0, sizeof(int) * cl.ncell));
    return AOSLO_OK;
}

}  // namespace aoslo
// device + pinned-host mirrors of the sklearn-order KD tree; only the sklearn tie mode needs them
static int kd_alloc(aoslo::KdTreeDev& kd, int cap) {
    using namespace aoslo;
    size_t nn = (size_t)2 * cap + 2;
    AOSLO_TRY(dmalloc(&kd.idx_array, cap));
    AOSLO_TRY(dmalloc(&kd.node_start, nn));
    AOSLO_TRY(dmalloc(&kd.node_end, nn));
    AOSLO_TRY(dmalloc(&kd.node_bounds, nn * 4));
    AOSLO_CUDA(cudaHostAlloc((void**)&kd.h_points, sizeof(double) * 2 * cap, cudaHostAllocDefault));
    AOSLO_CUDA(cudaHostAlloc((void**)&kd.h_idx_array, sizeof(int32_t) * cap, cudaHostAllocDefault));
    AOSLO_CUDA(cudaHostAlloc((void**)&kd.h_node_start, sizeof(int32_t) * nn, cudaHostAllocDefault));
    AOSLO_CUDA(cudaHostAlloc((void**)&kd.h_node_end, sizeof(int32_t) * nn, cudaHostAllocDefault));
    AOSLO_CUDA(cudaHostAlloc((void**)&kd.h_node_bounds, sizeof(double) * 4 * nn, cudaHostAllocDefault));
    kd.cap = cap;
    kd.cached_n = -1;
    return AOSLO_OK;
}
static void kd_free(aoslo::KdTreeDev& kd) {
    cudaFree(kd.idx_array); cudaFree(kd.node_start); cudaFree(kd.node_end); cudaFree(kd.node_bounds);
    cudaFreeHost(kd.h_points); cudaFreeHost(kd.h_idx_array); cudaFreeHost(kd.h_node_start);
    cudaFreeHost(kd.h_node_end); cudaFreeHost(kd.h_node_bounds);
}

int aoslo::kd_reserve(aoslo_ctx* c) {
    if (c->kd.cap >= c->cap) return AOSLO_OK;
    AOSLO_TRY(kd_alloc(c->kd, c->cap));
    AOSLO_TRY(kd_alloc(c->kd2, c->cap));
    AOSLO_CUDA(cudaHostAlloc((void**)&c->kd_stage, sizeof(double) * 2 * c->cap, cudaHostAllocDefault));
    return AOSLO_OK;
}
namespace aoslo {

static int alloc_links(CellLinks& lk, int H, int W, int cap, int fixed_shift) {
    lk.fixed_shift = fixed_shift;
    lk.min_shift = fixed_shift > 0 ? fixed_shift : 1;
    lk.shift = lk.min_shift;
    lk.gw = (W + (1 << lk.shift) - 1) >> lk.shift;
    lk.gh = (H + (1 << lk.shift) - 1) >> lk.shift;
    lk.ncell = lk.gw * lk.gh;                       // the finest grid: head[] is sized for it
    lk.cap = cap;
    AOSLO_TRY(dmalloc(&lk.head, lk.ncell));
    AOSLO_CUDA(cudaMemset(lk.head, 0, sizeof(unsigned long long) * lk.ncell));     // tag 0 is never a valid build
    AOSLO_TRY(dmalloc(&lk.next, cap));
    return AOSLO_OK;
}

static void free_cell_list(CellList& cl) {
    cudaFree(cl.count); cudaFree(cl.start); cudaFree(cl.slot); cudaFree(cl.cell_of);
    cudaFree(cl.sorted_idx); cudaFree(cl.sorted_pos);
}

// host waits for everything enqueued on `s`.  With many concurrent lanes (one host thread each)
// spinning in cudaStreamSynchronize burns a core per lane; a blocking-sync event sleeps instead.
int stream_wait(aoslo_ctx* ctx, cudaStream_t s) {
    if (ctx->blocking_sync && ctx->ev_block) {
        AOSLO_CUDA(cudaEventRecord(ctx->ev_block, s));
        AOSLO_CUDA(cudaEventSynchronize(ctx->ev_block));
    } else {
        AOSLO_CUDA(cudaStreamSynchronize(s));
    }
    return AOSLO_OK;
}

// read a device int through the pinned mailbox (one stream sync)
static int read_int(aoslo_ctx* ctx, cudaStream_t s, const int* d, int* out) {
    AOSLO_CUDA(cudaMemcpyAsync(ctx->h_pinned, d, sizeof(int), cudaMemcpyDeviceToHost, s));
    AOSLO_TRY(stream_wait(ctx, s));
    *out = ctx->h_pinned[0];
    return AOSLO_OK;
}

}  // namespace aoslo

using namespace aoslo;

extern "C" const char* aoslo_version(void) { return "aoslo_b200 0.1 (sm_100a)"; }
extern "C" const char* aoslo_last_cuda_error(void) { return g_last_err; }
extern "C" long long aoslo_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" int aoslo_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// Host wait policy of the device's primary context: 0 = auto, 1 = spin, 2 = yield, 3 = blocking sync
// (cudaSetDeviceFlags).  Yield keeps oversubscribed hosts (more waiting lane threads than cores) moving.
extern "C" int aoslo_set_wait_policy(int device, int policy) {
    unsigned flags = cudaDeviceScheduleAuto;
    if (policy == 1) flags = cudaDeviceScheduleSpin;
    else if (policy == 2) flags = cudaDeviceScheduleYield;
    else if (policy == 3) flags = cudaDeviceScheduleBlockingSync;
    int cur = 0;
    cudaGetDevice(&cur);
    AOSLO_CUDA(cudaSetDevice(device));
    cudaError_t e = cudaSetDeviceFlags(flags);
    cudaSetDevice(cur);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaSetDeviceFlags");
    return AOSLO_OK;
}

extern "C" int aoslo_ctx_create(aoslo_ctx** out, int device, int H, int W, int particle_capacity, int gt_capacity) {
    if (!out || H < 2 || W < 2 || particle_capacity < 2 || gt_capacity < 1) return AOSLO_ERR_INVALID;
    if ((long long)H * W > 0x7fffffffLL) return AOSLO_ERR_INVALID;
    int ndev = 0;
    AOSLO_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return AOSLO_ERR_INVALID;
    AOSLO_CUDA(cudaSetDevice(device));
    aoslo_ctx* c = new aoslo_ctx();
    c->device = device; c->H = H; c->W = W; c->cap = particle_capacity; c->gt_cap = gt_capacity;
    cudaDeviceProp prop;
    AOSLO_CUDA(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    c->nb = c->num_sms * 4;
    c->coop_blocks = coop_blocks_for_delete(c->num_sms);
    int shift = 2;
    if (const char* e = getenv("AOSLO_CELL_SHIFT")) { int v = atoi(e); if (v >= 1 && v <= 6) shift = v; }
    int cap = c->cap, gcap = c->gt_cap;
    size_t kn = (size_t)(cap > gcap ? cap : gcap) * AOSLO_KMAX;
    int st = AOSLO_OK;
#define A(x) do { if (st == AOSLO_OK) st = (x); } while (0)
    A(dmalloc(&c->d_status, 1));
    A(dmalloc(&c->d_counters, 32));
    if (st == AOSLO_OK && cudaHostAlloc((void**)&c->h_pinned, sizeof(int) * 32, cudaHostAllocDefault) != cudaSuccess)
        st = AOSLO_ERR_CUDA;
    A(alloc_cell_list(c->cl_g, H, W, gcap, shift));
    {
        int lshift = 0;                              // 0 = density adaptive
        if (const char* e = getenv("AOSLO_LINK_SHIFT")) { int v = atoi(e); if (v >= 1 && v <= 6) lshift = v; }
        A(alloc_links(c->lk_p, H, W, cap, lshift));
    }
    A(dmalloc(&c->d_pos_tmp, (size_t)cap * 2));
    A(dmalloc(&c->d_link_geom, 8));
    A(dmalloc(&c->d_link_epoch, 1));
    A(dmalloc(&c->d_link_epoch_ticket, 1));
    A(dmalloc(&c->d_link_epoch_n, 1));
    A(dmalloc(&c->d_head_n, c->lk_p.ncell));
    if (st == AOSLO_OK && cudaMemset(c->d_link_epoch_n, 0, sizeof(unsigned int)) != cudaSuccess) st = AOSLO_ERR_CUDA;
    if (st == AOSLO_OK && cudaMemset(c->d_head_n, 0, sizeof(unsigned long long) * c->lk_p.ncell) != cudaSuccess)
        st = AOSLO_ERR_CUDA;
    if (st == AOSLO_OK && cudaMemset(c->d_link_epoch, 0, sizeof(unsigned int)) != cudaSuccess) st = AOSLO_ERR_CUDA;
    if (st == AOSLO_OK && cudaMemset(c->d_link_epoch_ticket, 0, sizeof(int)) != cudaSuccess) st = AOSLO_ERR_CUDA;
    A(dmalloc(&c->d_ctl, 1));
    A(dmalloc(&c->d_params, 1));
    if (st == AOSLO_OK && cudaHostAlloc((void**)&c->h_ctl, sizeof(RunCtl), cudaHostAllocDefault) != cudaSuccess)
        st = AOSLO_ERR_CUDA;
    if (st == AOSLO_OK && cudaHostAlloc((void**)&c->h_params, sizeof(RunParams), cudaHostAllocDefault) != cudaSuccess)
        st = AOSLO_ERR_CUDA;
    // environment knobs are read once here, never on the launch path
    c->no_graph = getenv("AOSLO_NO_GRAPH") ? 1 : 0;
    if (const char* e = getenv("AOSLO_K1_TMA")) c->tune.k1_tma = atoi(e);
    if (const char* e = getenv("AOSLO_PAIR_TH")) c->tune.pair_th = atoi(e);
    if (const char* e = getenv("AOSLO_PAIR_OCC")) c->tune.pair_occ = atoi(e);
    if (const char* e = getenv("AOSLO_DTILE_TH")) c->tune.dtile_th = atoi(e);
    if (const char* e = getenv("AOSLO_BLOBNESS_KERNEL"))
        c->tune.k1_kernel = !strcmp(e, "tile") ? 1 : !strcmp(e, "band") ? 2 : !strcmp(e, "dtile") ? 3 : 0;
    A(dmalloc(&c->d_epoch, 1));
    A(dmalloc(&c->d_active, cap));
    A(dmalloc(&c->d_ns_list, cap));
    A(dmalloc(&c->d_knn_idx, kn));
    A(dmalloc(&c->d_knn_d2, kn));
    A(dmalloc(&c->d_block_sums, kMaxScanBlocks + 1));
    A(dmalloc(&c->d_scan_ticket, 1));
    if (st == AOSLO_OK && cudaMemset(c->d_scan_ticket, 0, sizeof(int)) != cudaSuccess) st = AOSLO_ERR_CUDA;
    A(dmalloc(&c->d_touch, cap));
    A(dmalloc(&c->d_del, cap));
    A(dmalloc(&c->d_rowstate, cap));
    A(dmalloc(&c->d_force, (size_t)cap * 2));
    A(dmalloc(&c->d_pos[0], (size_t)cap * 2));
    A(dmalloc(&c->d_pos[1], (size_t)cap * 2));
    A(dmalloc(&c->d_stable[0], cap));
    A(dmalloc(&c->d_stable[1], cap));
    A(dmalloc(&c->d_dist_p2g, cap));
    A(dmalloc(&c->d_dist_g2p, gcap));
    A(dmalloc(&c->d_partials, (size_t)c->nb * 16));
    A(dmalloc(&c->d_record, 1));
    c->max_records = 1024;
    A(dmalloc(&c->d_records, c->max_records));
    if (st == AOSLO_OK && cudaHostAlloc((void**)&c->h_records, sizeof(aoslo_metrics_record) * c->max_records, cudaHostAllocDefault) != cudaSuccess)
        st = AOSLO_ERR_CUDA;
    if (st == AOSLO_OK && cudaHostAlloc((void**)&c->h_record, sizeof(aoslo_metrics_record), cudaHostAllocDefault) != cudaSuccess)
        st = AOSLO_ERR_CUDA;
    A(dmalloc(&c->d_lut, 64 * 64));
    A(dmalloc(&c->d_lut_radial, 512));           // >= 2 * 15 * 15 + 1
    A(dmalloc(&c->d_lut_flag, 1));
    A(gravity_init());
    A(dmalloc(&c->d_occ, (size_t)H * ((W + 31) / 32 + 2) + 4));   // + 4: the clear kernel stores whole uint4
    A(dmalloc(&c->d_occ_gt, (size_t)H * ((W + 31) / 32 + 2) + 4));
    A(dmalloc(&c->d_occ_ok, 2));
    A(dmalloc(&c->d_thin, (size_t)7 * H * ((W + 31) / 32 + 2)));
    A(dmalloc(&c->d_thin_ctr, 8));
    A(dmalloc(&c->d_thin_touch, (size_t)H * W));
    c->thin_threads = thin_threads_default(H, W, c->num_sms);
    c->coop_blocks_thin = coop_blocks_for_thin(c->num_sms, c->thin_threads);
    A(dmalloc(&c->d_field, (size_t)H * W));
    A(dmalloc(&c->d_win, (size_t)H * W));
    c->kd.cap = 0;   // sklearn-order KD tree mirrors are allocated on first use (aoslo::kd_reserve)
    for (int i = 0; i < 12 && st == AOSLO_OK; ++i)
        if (cudaEventCreate(&c->ev[i]) != cudaSuccess) st = AOSLO_ERR_CUDA;
    if (st == AOSLO_OK && cudaEventCreateWithFlags(&c->ev_block, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess)
        st = AOSLO_ERR_CUDA;
    if (st == AOSLO_OK && cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking) != cudaSuccess) st = AOSLO_ERR_CUDA;
    if (st == AOSLO_OK && cudaStreamCreateWithFlags(&c->cap_stream2, cudaStreamNonBlocking) != cudaSuccess) st = AOSLO_ERR_CUDA;
    if (st == AOSLO_OK && cudaStreamCreateWithFlags(&c->cap_stream3, cudaStreamNonBlocking) != cudaSuccess) st = AOSLO_ERR_CUDA;
#undef A
    if (st == AOSLO_OK && cudaMemset(c->d_status, 0, sizeof(int)) != cudaSuccess) st = AOSLO_ERR_CUDA;
    if (st == AOSLO_OK && cudaMemset(c->d_touch, 0, sizeof(unsigned long long) * cap) != cudaSuccess) st = AOSLO_ERR_CUDA;
    if (st == AOSLO_OK) { unsigned int one = 1u; if (cudaMemcpy(c->d_epoch, &one, sizeof(one), cudaMemcpyHostToDevice) != cudaSuccess) st = AOSLO_ERR_CUDA; }
    if (st == AOSLO_OK && cudaMemset(c->d_counters, 0, sizeof(int) * 32) != cudaSuccess) st = AOSLO_ERR_CUDA;
    if (st != AOSLO_OK) { aoslo_ctx_destroy(c); return st; }
    *out = c;
    return AOSLO_OK;
}

extern "C" int aoslo_ctx_destroy(aoslo_ctx* c) {
    if (!c) return AOSLO_OK;
    cudaSetDevice(c->device);
    cudaFree(c->d_status); cudaFree(c->d_counters); cudaFreeHost(c->h_pinned);
    free_cell_list(c->cl_g);
    cudaFree(c->lk_p.head); cudaFree(c->lk_p.next); cudaFree(c->d_pos_tmp); cudaFree(c->d_active); cudaFree(c->d_ns_list);
    cudaFree(c->d_knn_idx); cudaFree(c->d_knn_d2); cudaFree(c->d_block_sums); cudaFree(c->d_scan_ticket);
    cudaFree(c->d_touch); cudaFree(c->d_del); cudaFree(c->d_rowstate); cudaFree(c->d_force);
    cudaFree(c->d_pos[0]); cudaFree(c->d_pos[1]); cudaFree(c->d_stable[0]); cudaFree(c->d_stable[1]);
    cudaFree(c->d_dist_p2g); cudaFree(c->d_dist_g2p); cudaFree(c->d_partials); cudaFree(c->d_record);
    cudaFreeHost(c->h_record); cudaFree(c->d_records); cudaFreeHost(c->h_records); cudaFree(c->d_lut); cudaFree(c->d_lut_radial); cudaFree(c->d_lut_flag); cudaFree(c->d_occ); cudaFree(c->d_occ_gt); cudaFree(c->d_occ_ok); cudaFree(c->d_thin); cudaFree(c->d_thin_ctr); cudaFree(c->d_thin_touch); cudaFree(c->d_field); cudaFree(c->d_win);
    kd_free(c->kd); kd_free(c->kd2); cudaFreeHost(c->kd_stage);
    for (int i = 0; i < 12; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    if (c->ev_block) cudaEventDestroy(c->ev_block);
    if (c->ev_upload) cudaEventDestroy(c->ev_upload);
    for (RunGraph& g : c->graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
    }
    if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
    if (c->cap_stream2) cudaStreamDestroy(c->cap_stream2);
    if (c->cap_stream3) cudaStreamDestroy(c->cap_stream3);
    cudaFree(c->d_link_geom); cudaFree(c->d_link_epoch); cudaFree(c->d_link_epoch_ticket);
    cudaFree(c->d_link_epoch_n); cudaFree(c->d_head_n); cudaFree(c->d_ctl); cudaFree(c->d_params);
    cudaFreeHost(c->h_ctl); cudaFreeHost(c->h_params);
    cudaFree(c->d_epoch);
    cudaGetLastError();
    delete c;
    return AOSLO_OK;
}

extern "C" int aoslo_ctx_capacity(aoslo_ctx* ctx) { return ctx ? ctx->cap : 0; }

extern "C" int aoslo_ctx_set_tuning(aoslo_ctx* ctx, const char* key, int value) {
    if (!ctx || !key) return AOSLO_ERR_INVALID;
    if (!strcmp(key, "k1_tma")) ctx->tune.k1_tma = value;
    else if (!strcmp(key, "pair_th")) ctx->tune.pair_th = value;
    else if (!strcmp(key, "pair_occ")) ctx->tune.pair_occ = value;
    else if (!strcmp(key, "dtile_th")) ctx->tune.dtile_th = value;
    else if (!strcmp(key, "dtile_split")) ctx->tune.dtile_split = value;
    else if (!strcmp(key, "dtile_occ")) ctx->tune.dtile_occ = value;
    else if (!strcmp(key, "k1_kernel")) ctx->tune.k1_kernel = value;
    else if (!strcmp(key, "no_graph")) ctx->no_graph = value ? 1 : 0;
    else return AOSLO_ERR_INVALID;
    return AOSLO_OK;
}

extern "C" int aoslo_ctx_set_blocking_sync(aoslo_ctx* ctx, int enable) {
    if (!ctx) return AOSLO_ERR_INVALID;
    ctx->blocking_sync = enable ? 1 : 0;
    return AOSLO_OK;
}

extern "C" int aoslo_ctx_status(aoslo_ctx* ctx, void* stream, int* h_status_out, int clear) {
    if (!ctx || !h_status_out) return AOSLO_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    AOSLO_TRY(read_int(ctx, s, ctx->d_status, h_status_out));
    if (clear) AOSLO_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(int), s));
    return AOSLO_OK;
}

extern "C" int aoslo_ctx_clear_status(aoslo_ctx* ctx, void* stream) {
    if (!ctx) return AOSLO_ERR_INVALID;
    AOSLO_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(int), (cudaStream_t)stream));
    return AOSLO_OK;
}

extern "C" int aoslo_prepare_inputs(aoslo_ctx* ctx, void* stream, const uint8_t* d_raw_img, int H0, int W0,
                                    int pitch0_bytes, int x_min, int x_max, int y_min, int y_max, int bx, int by,
                                    uint8_t* d_img_out, int out_pitch_bytes, const double* d_gt, int n_gt_raw,
                                    double* d_gt_out, int* h_n_gt_out, int gt_is_raw) {
    if (!ctx || !d_raw_img || !d_img_out) return AOSLO_ERR_INVALID;
    if (n_gt_raw > 0 && (!d_gt || !d_gt_out)) return AOSLO_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    int* d_count = ctx->d_counters + 30;             // dedicated: aoslo_run(n_gt = -1) reads it on the device
    AOSLO_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int), s));
    AOSLO_TRY(stage_prepare(ctx, s, d_raw_img, H0, W0, pitch0_bytes, x_min, x_max, y_min, y_max, bx, by, d_img_out,
                            out_pitch_bytes, d_gt, n_gt_raw, d_gt_out, d_count, gt_is_raw));
    if (!h_n_gt_out) return AOSLO_OK;                // the caller does not need the count on the host: no sync
    return read_int(ctx, s, d_count, h_n_gt_out);
}

// Deferred host work: `h[i] += addend` on the caller's array, once the upload that read it has completed.
// aoslo_run calls it between enqueueing its device work and waiting for it.
static int run_deferred_host_work(aoslo_ctx* ctx) {
    if (!ctx->defer_ptr) return AOSLO_OK;
    double* h = ctx->defer_ptr;
    const long long n = ctx->defer_n;
    const double a = ctx->defer_add;
    ctx->defer_ptr = nullptr;
    AOSLO_CUDA(cudaEventSynchronize(ctx->ev_upload));
    for (long long i = 0; i < n; ++i) h[i] += a;
    return AOSLO_OK;
}

extern "C" int aoslo_stage_gt(aoslo_ctx* ctx, void* stream, double* h_gt, int n_gt_raw, double* d_gt_raw,
                              int shift_in_place) {
    if (!ctx || n_gt_raw < 0 || (n_gt_raw > 0 && (!h_gt || !d_gt_raw))) return AOSLO_ERR_INVALID;
    AOSLO_TRY(run_deferred_host_work(ctx));          // an earlier staging that no run picked up
    if (n_gt_raw == 0) return AOSLO_OK;
    cudaStream_t s = (cudaStream_t)stream;
    AOSLO_CUDA(cudaMemcpyAsync(d_gt_raw, h_gt, sizeof(double) * 2 * (size_t)n_gt_raw, cudaMemcpyHostToDevice, s));
    if (!shift_in_place) return AOSLO_OK;
    if (!ctx->ev_upload) AOSLO_CUDA(cudaEventCreateWithFlags(&ctx->ev_upload, cudaEventDisableTiming));
    AOSLO_CUDA(cudaEventRecord(ctx->ev_upload, s));
    ctx->defer_ptr = h_gt;
    ctx->defer_n = 2LL * n_gt_raw;
    ctx->defer_add = -1.0;
    return AOSLO_OK;
}

extern "C" int aoslo_ctx_flush_host_work(aoslo_ctx* ctx) {
    if (!ctx) return AOSLO_ERR_INVALID;
    return run_deferred_host_work(ctx);
}

extern "C" int aoslo_seed(aoslo_ctx* ctx, void* stream, const void* d_Rb, int field_dtype, int H, int W, int bx,
                          int by, double threshold, double* d_particles, int capacity, int* h_n_out) {
    if (!ctx || !d_Rb || !d_particles || !h_n_out || H != ctx->H || W != ctx->W) return AOSLO_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    AOSLO_TRY(stage_seed(ctx, s, d_Rb, field_dtype, H, W, bx, by, threshold, d_particles, capacity, ctx->d_counters));
    AOSLO_TRY(read_int(ctx, s, ctx->d_counters, h_n_out));
    return (*h_n_out > capacity) ? AOSLO_ERR_CAPACITY : AOSLO_OK;
}

extern "C" int aoslo_knn(aoslo_ctx* ctx, void* stream, const double* d_points, int n_points, const double* d_query,
                         int n_query, int k, int tie_mode, int32_t* d_idx_out, double* d_dist_out) {
    if (!ctx || !d_points || !d_query || !d_idx_out || !d_dist_out || n_points < 1 || n_query < 0) return AOSLO_ERR_INVALID;
    if (k < 1 || k > AOSLO_KMAX) return AOSLO_ERR_UNSUPPORTED;
    if (n_points < k) return AOSLO_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    // distances come back squared from the kernels; take the root in place (rdist_to_dist)
    AOSLO_TRY(stage_knn(ctx, s, d_points, n_points, d_query, n_query, k, tie_mode, d_idx_out, d_dist_out, true, nullptr));
    return sqrt_inplace(s, d_dist_out, (size_t)n_query * k);
}

extern "C" int aoslo_delete_close(aoslo_ctx* ctx, void* stream, const double* d_particles, const uint8_t* d_stable,
                                  int n, const void* d_Rb, int field_dtype, int H, int W, double threshold,
                                  int tie_mode, double* d_particles_out, uint8_t* d_stable_out,
                                  uint8_t* d_deleted_out, int* h_n_out) {
    if (!ctx || !d_particles || !d_stable || !d_Rb || !d_particles_out || !d_stable_out || !h_n_out) return AOSLO_ERR_INVALID;
    if (H != ctx->H || W != ctx->W) return AOSLO_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    AOSLO_TRY(stage_delete(ctx, s, d_particles, d_stable, n, d_Rb, field_dtype, H, W, threshold, tie_mode,
                           d_particles_out, d_stable_out, d_deleted_out, ctx->d_counters, ctx->d_counters + 1, nullptr, 0));
    return read_int(ctx, s, ctx->d_counters, h_n_out);
}

extern "C" int aoslo_dist_lut(aoslo_ctx* ctx, void* stream, int rfs, double mu, double sigma, double lambda_,
                              double* d_lut) {
    if (!ctx || !d_lut) return AOSLO_ERR_INVALID;
    return stage_lut(ctx, (cudaStream_t)stream, rfs, mu, sigma, lambda_, d_lut);
}

extern "C" int aoslo_gravity_field(aoslo_ctx* ctx, void* stream, const double* d_particles, int n,
                                   const double* d_lut, int rfs, int H, int W, double* d_field) {
    if (!ctx || (!d_particles && n > 0) || !d_lut || !d_field || H != ctx->H || W != ctx->W || n < 0)
        return AOSLO_ERR_INVALID;
    if (n > ctx->cap) return AOSLO_ERR_CAPACITY;
    cudaStream_t s = (cudaStream_t)stream;
    return stage_gravity(ctx, s, d_particles, n, nullptr, d_lut, rfs, H, W, d_field, false);
}

extern "C" int aoslo_add_particles(aoslo_ctx* ctx, void* stream, const double* d_field, int H, int W, int bx, int by,
                                   int window, double threshold, double* d_particles, uint8_t* d_stable, int n,
                                   int capacity, int* h_n_out) {
    if (!ctx || !d_field || !d_particles || !d_stable || !h_n_out || H != ctx->H || W != ctx->W) return AOSLO_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    AOSLO_TRY(stage_add(ctx, s, d_field, H, W, bx, by, window, threshold, d_particles, d_stable, n, nullptr, capacity,
                        ctx->d_counters));
    AOSLO_TRY(read_int(ctx, s, ctx->d_counters, h_n_out));
    return (*h_n_out > capacity) ? AOSLO_ERR_CAPACITY : AOSLO_OK;
}

extern "C" int aoslo_dist_force(aoslo_ctx* ctx, void* stream, const double* d_particles, const uint8_t* d_stable,
                                int n, int n_neighbours, int energy_type, double p0, double p1, double p2,
                                int tie_mode, double* d_force) {
    if (!ctx || !d_particles || !d_force) return AOSLO_ERR_INVALID;
    if (energy_type != AOSLO_ENERGY_PIT && energy_type != AOSLO_ENERGY_EXP) return AOSLO_ERR_UNSUPPORTED;
    return stage_force(ctx, (cudaStream_t)stream, d_particles, d_stable, n, n_neighbours, energy_type, p0, p1, p2,
                       tie_mode, d_force, nullptr);
}

extern "C" int aoslo_move(aoslo_ctx* ctx, void* stream, double* d_particles, const uint8_t* d_stable, int n,
                          const void* d_grad, int field_dtype, int H, int W, const double* d_force,
                          double lambda_blob, double lambda_dist, int step_mode) {
    if (!ctx || !d_particles || !d_stable || !d_grad || !d_force || n < 0) return AOSLO_ERR_INVALID;
    if (step_mode != AOSLO_STEP_DISCRETE && step_mode != AOSLO_STEP_CONTIN) return AOSLO_ERR_UNSUPPORTED;
    return stage_move(ctx, (cudaStream_t)stream, d_particles, d_stable, n, d_grad, field_dtype, H, W, d_force,
                      lambda_blob, lambda_dist, step_mode, nullptr);
}

extern "C" int aoslo_metrics(aoslo_ctx* ctx, void* stream, const double* d_particles, const uint8_t* d_stable, int n,
                             const double* d_gt, int n_gt, const void* d_Rb, int field_dtype, int H, int W, int bx,
                             int by, aoslo_metrics_record* h_record_out, double* d_dist_p2g, double* d_dist_g2p) {
    if (!ctx || !d_particles || !d_stable || !d_gt || !d_Rb || !h_record_out) return AOSLO_ERR_INVALID;
    if (H != ctx->H || W != ctx->W) return AOSLO_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    double* p2g = d_dist_p2g ? d_dist_p2g : ctx->d_dist_p2g;
    double* g2p = d_dist_g2p ? d_dist_g2p : ctx->d_dist_g2p;
    AOSLO_TRY(stage_metrics(ctx, s, d_particles, d_stable, n, d_gt, n_gt, false, false, d_Rb, field_dtype, H, W, bx,
                            by, 0, ctx->d_record, p2g, g2p, nullptr));
    AOSLO_CUDA(cudaMemcpyAsync(ctx->h_record, ctx->d_record, sizeof(aoslo_metrics_record), cudaMemcpyDeviceToHost, s));
    AOSLO_TRY(stream_wait(ctx, s));
    *h_record_out = *ctx->h_record;
    return AOSLO_OK;
}

extern "C" int aoslo_classify(aoslo_ctx* ctx, void* stream, const double* d_dist_g2p, int n_gt,
                              const double* d_dist_p2g, const double* d_particles, int n, int bx, int by, int H,
                              int W, double threshold, uint8_t* d_gt_is_tp, uint8_t* d_particle_is_fp) {
    if (!ctx || !d_dist_g2p || !d_dist_p2g || !d_particles || !d_gt_is_tp || !d_particle_is_fp || n < 0 || n_gt < 0)
        return AOSLO_ERR_INVALID;
    return stage_classify(ctx, (cudaStream_t)stream, d_dist_g2p, n_gt, d_dist_p2g, d_particles, n, bx, by, W, H,
                          threshold, d_gt_is_tp, d_particle_is_fp);
}

// One loop iteration without the resample / metrics tail: distance force, move, deletion
// (reference pipeline_with_delitions.py:173-253), on the kernels aoslo_run executes per epoch.
extern "C" int aoslo_iterate(aoslo_ctx* ctx, void* stream, const aoslo_config* cfg, const void* d_Rb,
                             const void* d_grad, int field_dtype, const double* d_particles,
                             const uint8_t* d_stable, int n, double* d_particles_out, uint8_t* d_stable_out,
                             int* h_n_out) {
    if (!ctx || !cfg || !d_Rb || !d_grad || !d_particles || !d_stable || !d_particles_out || !d_stable_out || !h_n_out)
        return AOSLO_ERR_INVALID;
    if (cfg->H != ctx->H || cfg->W != ctx->W || n > ctx->cap) return AOSLO_ERR_INVALID;
    const int k = cfg->dist_n_neighbours;
    if (k < 1 || k + 1 > AOSLO_KMAX) return AOSLO_ERR_UNSUPPORTED;
    if (n < k + 1 || n < 2) return AOSLO_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const bool pit = cfg->dist_energy_type == AOSLO_ENERGY_PIT;
    const double e0 = pit ? cfg->mu_dist_func : cfg->dist_alpha;
    const double e1 = pit ? cfg->sigma_dist_func : cfg->dist_sigma;
    const double e2 = pit ? cfg->lambda_dist_func : 0.0;
    AOSLO_TRY(epoch_guard(ctx, s, 2u));              // the 20-bit touch epoch must not wrap (one resolve per call)
    // a call that failed between the force step and the deletion must not leak its non-stable list
    AOSLO_CUDA(cudaMemsetAsync(ctx->d_counters + 12, 0, sizeof(int) * 4, s));
    if (cfg->tie_mode == AOSLO_TIE_CANONICAL) {
        // the code path of aoslo_run: particles into buffer 0, one resample period (S / N structures, non-stable
        // list), one step on the tombstone structure, order-preserving compaction of the survivors
        int* cnt0 = ctx->d_counters + 4;
        int* cnt1 = ctx->d_counters + 5;
        ctx->h_pinned[8] = n;
        AOSLO_CUDA(cudaMemcpyAsync(cnt0, ctx->h_pinned + 8, sizeof(int), cudaMemcpyHostToDevice, s));
        AOSLO_CUDA(cudaMemcpyAsync(ctx->d_pos[0], d_particles, sizeof(double) * 2 * n, cudaMemcpyDeviceToDevice, s));
        AOSLO_CUDA(cudaMemcpyAsync(ctx->d_stable[0], d_stable, n, cudaMemcpyDeviceToDevice, s));
        AOSLO_TRY(stage_period_begin(ctx, s, n, cnt0));
        AOSLO_TRY(stage_dual_step(ctx, s, n, cnt0, k, cfg->dist_energy_type, e0, e1, e2, d_grad, d_Rb, field_dtype,
                                  cfg->H, cfg->W, cfg->lambda_blob, cfg->lambda_dist, cfg->step_mode,
                                  cfg->threshhold_dist_del, nullptr, 3));
        ++ctx->epoch;
        AOSLO_TRY(stage_compact(ctx, s, n, cnt0, cnt1, 0));
        AOSLO_TRY(read_int(ctx, s, cnt0, h_n_out));
        const int m = *h_n_out;
        AOSLO_CUDA(cudaMemcpyAsync(d_particles_out, ctx->d_pos[0], sizeof(double) * 2 * m, cudaMemcpyDeviceToDevice, s));
        AOSLO_CUDA(cudaMemcpyAsync(d_stable_out, ctx->d_stable[0], m, cudaMemcpyDeviceToDevice, s));
        return AOSLO_OK;
    }
    // sklearn tie order: the stage kernels of run_sklearn (host-built tree per neighbour search)
    AOSLO_CUDA(cudaMemcpyAsync(ctx->d_pos_tmp, d_particles, sizeof(double) * 2 * n, cudaMemcpyDeviceToDevice, s));
    AOSLO_TRY(stage_force(ctx, s, ctx->d_pos_tmp, d_stable, n, k, cfg->dist_energy_type, e0, e1, e2, cfg->tie_mode,
                          ctx->d_force, nullptr));
    AOSLO_TRY(stage_move(ctx, s, ctx->d_pos_tmp, d_stable, n, d_grad, field_dtype, cfg->H, cfg->W, ctx->d_force,
                         cfg->lambda_blob, cfg->lambda_dist, cfg->step_mode, nullptr));
    AOSLO_TRY(stage_delete(ctx, s, ctx->d_pos_tmp, d_stable, n, d_Rb, field_dtype, cfg->H, cfg->W,
                           cfg->threshhold_dist_del, cfg->tie_mode, d_particles_out, d_stable_out, nullptr,
                           ctx->d_counters, nullptr, nullptr, 1));
    return read_int(ctx, s, ctx->d_counters, h_n_out);
}

// ---------------------------------------------------------------------------------------
// whole run
// ---------------------------------------------------------------------------------------
// Control flow = reference pipeline_with_delitions.py:101-382.
//
// Canonical tie order (run_canonical): the WHOLE run -- seeds, thinning to the fixpoint, LUT / gravity /
// adding, the epoch loop, the final reads -- is ONE CUDA graph with no host decision inside.  The
// particles always live in buffer 0 (the compaction of a step reads the moved copies); the count is a
// device scalar; every decision the reference takes in Python is taken by a one-thread kernel:
//   * seeds + thinning `while deleted: del_close_particles(thr=3)` (:101-120) -> one cooperative kernel on a bitmap of
//     alive pixels (stage_thin_pixels); thinning thresholds above 4 px: a WHILE conditional graph node of thinning rounds;
//   * the epoch loop -> unrolled kernel nodes when the early stop is off; with the early stop (:381-382) a
//     WHILE node whose body carries IF nodes for the resample (:256) and metrics (:320) branches;
//   * all value parameters (thresholds, lambdas, energy parameters, frequencies, GT count) are read from a
//     device-resident RunParams block, so one cached graph serves every configuration of a sweep -- only
//     structural parameters (shapes, neighbour-count class, loop layout, buffer addresses) are baked in.
// The host synchronises exactly once, at the end, to read the summary / records / status.
//
// sklearn tie order (run_sklearn): every kNN needs the host-built tree of the current points, so the run
// stays a sequence of stage calls with the count on the host.
namespace aoslo {

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void ctl_stamp(RunCtl* ctl, int bucket) {
    const unsigned long long t = global_ns();
    if (bucket >= 0) ctl->acc_ns[bucket] += t - ctl->t_last;
    ctl->t_last = t;
}

__global__ void run_begin_kernel(RunCtl* ctl, int* status, int* counters) {
    if (threadIdx.x != 0) return;
    RunCtl z = {};
    *ctl = z;
    (void)status;             // sticky: conditions raised before the run (the field's eigenvalue assert) are kept
    counters[12] = 0; counters[13] = 0; counters[14] = 0; counters[15] = 0;
    counters[4] = 0; counters[5] = 0;
}

__global__ void after_seed_kernel(RunCtl* ctl, const int* cnt, cudaGraphConditionalHandle h_thin, int use_handle) {
    if (threadIdx.x != 0) return;
    const int n = *cnt;
    ctl->n_seeds = n;
    ctl->n_after_thinning = n;
    ctl->thin_cond = n >= 2 ? 1 : 0;
    if (use_handle) cudaGraphSetConditional(h_thin, (unsigned)ctl->thin_cond);
}

// end of one thinning round: the survivors (buffer 1) become buffer 0; another round follows iff something
// was deleted (reference :116-120)
__global__ void __launch_bounds__(kThreads) thin_swap_kernel(const double2* __restrict__ pos1,
                                                             const uint8_t* __restrict__ st1, const int* cnt1,
                                                             double2* __restrict__ pos0, uint8_t* __restrict__ st0,
                                                             int* cnt0, RunCtl* ctl, cudaGraphConditionalHandle h_thin,
                                                             int use_handle) {
    const int n1 = *cnt1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += gridDim.x * blockDim.x) {
        pos0[i] = pos1[i];
        st0[i] = st1[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int n0 = *cnt0;            // no other thread of this kernel reads cnt0
        *cnt0 = n1;
        ctl->thin_rounds += 1;
        ctl->n_after_thinning = n1;
        ctl->thin_cond = (n1 != n0 && n1 >= 2) ? 1 : 0;
        if (use_handle) cudaGraphSetConditional(h_thin, (unsigned)ctl->thin_cond);
    }
}

__global__ void loop_begin_kernel(RunCtl* ctl, const RunParams* rp, const int* cnt, const int* n_dead, int* status,
                                  cudaGraphConditionalHandle h_loop, int use_handle, int stamp) {
    if (threadIdx.x != 0) return;
    const int n = *cnt - *n_dead;
    int go = rp->epoch_num > 0 ? 1 : 0;
    if (n < rp->kk || n < 2) { atomicOr(status, AOSLO_DEV_KNN_SHORT); go = 0; }   // sklearn: n_neighbors > n_samples
    ctl->it = 0;
    ctl->loop_cond = go;
    if (use_handle) cudaGraphSetConditional(h_loop, (unsigned)go);
    if (stamp) ctl_stamp(ctl, -1);
}

// after the deletion of an iteration: the two branch conditions of this iteration are fixed (resample :256,
// metrics :320)
__global__ void iter_mid_kernel(RunCtl* ctl, const RunParams* rp, cudaGraphConditionalHandle h_res,
                                cudaGraphConditionalHandle h_met, int use_res, int use_met, int stamp) {
    if (threadIdx.x != 0) return;
    const int it = ctl->it;
    const int do_res = (it % rp->fix_freq == 3) ? 1 : 0;
    const int do_met = (it % rp->metric_freq == 0 && ctl->n_rec < rp->max_records) ? 1 : 0;
    ctl->do_resample = do_res;
    ctl->do_metrics = do_met;
    if (use_res) cudaGraphSetConditional(h_res, (unsigned)do_res);
    if (use_met) cudaGraphSetConditional(h_met, (unsigned)do_met);
    if (stamp) ctl_stamp(ctl, 2);
}

__global__ void iter_end_kernel(RunCtl* ctl, const RunParams* rp, const int* cnt, const int* n_dead, int* status,
                                cudaGraphConditionalHandle h_loop, int use_handle, int stamp) {
    if (threadIdx.x != 0) return;
    const int it = ctl->it + 1;
    ctl->it = it;
    const int n = *cnt - *n_dead;                    // survivors (tombstones are compacted at the next resample)
    int go = it < rp->epoch_num ? 1 : 0;
    // sum(stable) / len(particles) > 0.997 (:381) -- Python true division in float64
    if (rp->early_stop && (double)ctl->n_stable / (double)n > 0.997) go = 0;
    if (go && (n < rp->kk || n < 2)) { atomicOr(status, AOSLO_DEV_KNN_SHORT); go = 0; }
    ctl->loop_cond = go;
    if (use_handle) cudaGraphSetConditional(h_loop, (unsigned)go);
    if (stamp) ctl_stamp(ctl, 4);
}

__global__ void stamp_kernel(RunCtl* ctl, int bucket) {
    if (threadIdx.x == 0) ctl_stamp(ctl, bucket);
}

// n_gt = -1: the GT count of the run is the device scalar left by aoslo_prepare_inputs
__global__ void params_ngt_kernel(RunParams* rp, const int* d_n_gt) {
    if (threadIdx.x == 0) rp->n_gt = *d_n_gt;
}

__global__ void run_end_kernel(RunCtl* ctl, const int* cnt) {
    if (threadIdx.x == 0) ctl->n_final = *cnt;
}

static thread_local int* tl_capture_counter = nullptr;   // kernel launches recorded into a graph body

}  // namespace aoslo

void aoslo::count_launch() {
    if (tl_capture_counter) ++*tl_capture_counter;
    else g_launches.fetch_add(1, std::memory_order_relaxed);
}

namespace {

struct CondNode {
    cudaGraphConditionalHandle handle = 0;
    cudaGraph_t body = nullptr;
};

// WHILE / IF node after everything captured so far on `q`; the capture continues after the node.  The handle is
// created by the caller BEFORE the kernel that sets it is recorded.
int add_cond_node(cudaStream_t q, cudaGraphConditionalNodeType type, CondNode* node) {
    cudaStreamCaptureStatus cs;
    cudaGraph_t g = nullptr;
    const cudaGraphNode_t* deps = nullptr;
    size_t ndeps = 0;
    AOSLO_CUDA(cudaStreamGetCaptureInfo(q, &cs, nullptr, &g, &deps, &ndeps));
    if (cs != cudaStreamCaptureStatusActive) return AOSLO_ERR_INVALID;
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = node->handle;
    p.conditional.type = type;
    p.conditional.size = 1;
    cudaGraphNode_t gn;
    AOSLO_CUDA(cudaGraphAddNode(&gn, g, deps, ndeps, &p));
    AOSLO_CUDA(cudaStreamUpdateCaptureDependencies(q, &gn, 1, cudaStreamSetCaptureDependencies));
    node->body = p.conditional.phGraph_out[0];
    return AOSLO_OK;
}

int new_cond_handle(cudaStream_t q, cudaGraphConditionalHandle* h) {
    cudaStreamCaptureStatus cs;
    cudaGraph_t g = nullptr;
    AOSLO_CUDA(cudaStreamGetCaptureInfo(q, &cs, nullptr, &g, nullptr, nullptr));
    if (cs != cudaStreamCaptureStatusActive) return AOSLO_ERR_INVALID;
    AOSLO_CUDA(cudaGraphConditionalHandleCreate(h, g, 0, cudaGraphCondAssignDefault));
    return AOSLO_OK;
}

struct RunKey {                      // everything baked into the nodes of a cached run graph
    int H, W, bx, by, fdt, k_class, energy, step_mode, rfs, window;
    int dynamic_loop, timing, epoch_num, fix_freq, metric_freq, max_rec, hint, has_records, thin_pixels;
    const void *Rb, *grad, *gt, *records;
};

struct RunPlan {
    aoslo_ctx* ctx;
    const aoslo_config* cfg;
    const void* d_Rb;
    const void* d_grad;
    int fdt;
    const double* d_gt;
    int n_gt;
    bool want_records;
    int max_records;
    bool timing;
    bool dynamic_loop;               // WHILE / IF nodes (early stop) instead of the unrolled loop
    bool thin_pixels;                // seeds + thinning in the pixel domain (thinning threshold <= 4 px)
    int inner;                       // inner pixels: upper bound of the seed count
    int hint;                        // launch-size hint of the in-loop compaction
    int* cnt0;
    int* cnt1;
    // kernel launches per body, counted while capturing (gpu_launches bookkeeping)
    int k_static = 0, k_thin = 0, k_iter = 0, k_res = 0, k_met = 0;
};

int enqueue_thin_round(RunPlan& P, cudaStream_t q, cudaGraphConditionalHandle h, int use_handle) {
    aoslo_ctx* c = P.ctx;
    const aoslo_config* cfg = P.cfg;
    AOSLO_TRY(stage_delete(c, q, c->d_pos[0], c->d_stable[0], P.inner, P.d_Rb, P.fdt, cfg->H, cfg->W,
                           cfg->thinning_threshold, AOSLO_TIE_CANONICAL, c->d_pos[1], c->d_stable[1], nullptr, P.cnt1,
                           nullptr, P.cnt0, 0, &c->d_params->thin_thr, 0));
    thin_swap_kernel<<<c->nb, kThreads, 0, q>>>((const double2*)c->d_pos[1], c->d_stable[1], P.cnt1,
                                                (double2*)c->d_pos[0], c->d_stable[0], P.cnt0, c->d_ctl, h, use_handle);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

// resample (:256-264 / :130-152) on buffer 0: [compaction of the survivors,] all stable, gravity field, adding,
// then a new resample period (S / N structures, non-stable list)
int enqueue_resample(RunPlan& P, cudaStream_t q, bool first) {
    aoslo_ctx* c = P.ctx;
    const aoslo_config* cfg = P.cfg;
    const int bound = c->cap;
    if (!first) AOSLO_TRY(stage_compact(c, q, bound, P.cnt0, P.cnt1, P.hint));
    AOSLO_TRY(stage_set_stable(c, q, c->d_stable[0], bound, 1, P.cnt0));
    AOSLO_TRY(stage_gravity(c, q, c->d_pos[0], bound, P.cnt0, c->d_lut, cfg->reception_field_size, cfg->H, cfg->W,
                            c->d_field, !first));
    // the spine kernel reads the base count before it writes the total, so both may be the same scalar
    AOSLO_TRY(stage_add(c, q, c->d_field, cfg->H, cfg->W, cfg->bx, cfg->by, cfg->particle_call_window,
                        cfg->particle_call_threshold, c->d_pos[0], c->d_stable[0], bound, P.cnt0, c->cap, P.cnt0,
                        &c->d_params->add_thr, true));
    return stage_period_begin(c, q, bound, P.cnt0);
}

int enqueue_metrics(RunPlan& P, cudaStream_t q, int static_it, int static_slot) {
    aoslo_ctx* c = P.ctx;
    const aoslo_config* cfg = P.cfg;
    const bool dyn = static_slot < 0;
    return stage_dual_metrics(c, q, c->cap, P.cnt0, P.d_gt, P.n_gt, P.d_Rb, P.fdt, cfg->H, cfg->W, cfg->bx, cfg->by,
                              dyn ? 0 : static_it, c->d_records + (dyn ? 0 : static_slot), &c->d_params->n_gt,
                              dyn ? &c->d_ctl->n_rec : nullptr, dyn ? &c->d_ctl->it : nullptr);
}

// force + move + deletion of one iteration (:185-253) on buffer 0; deleted particles become tombstones
int enqueue_step(RunPlan& P, cudaStream_t q) {
    aoslo_ctx* c = P.ctx;
    const aoslo_config* cfg = P.cfg;
    const bool pit = cfg->dist_energy_type == AOSLO_ENERGY_PIT;
    const double e0 = pit ? cfg->mu_dist_func : cfg->dist_alpha;
    const double e1 = pit ? cfg->sigma_dist_func : cfg->dist_sigma;
    const double e2 = pit ? cfg->lambda_dist_func : 0.0;
    for (int phase = 1; phase <= 2; ++phase) {
        AOSLO_TRY(stage_dual_step(c, q, c->cap, P.cnt0, cfg->dist_n_neighbours, cfg->dist_energy_type, e0, e1, e2,
                                  P.d_grad, P.d_Rb, P.fdt, cfg->H, cfg->W, cfg->lambda_blob, cfg->lambda_dist,
                                  cfg->step_mode, cfg->threshhold_dist_del, c->d_params, phase));
        if (phase == 1 && P.timing) {
            stamp_kernel<<<1, 32, 0, q>>>(c->d_ctl, 0);   // force and move are one kernel: the move bucket stays 0
            AOSLO_LAUNCH_CHECK();
        }
    }
    return AOSLO_OK;
}

// captures `fn` into the body graph of a conditional node; launches are counted into *counter
template <class F>
int capture_body(cudaStream_t q, cudaGraph_t body, int* counter, F fn) {
    AOSLO_CUDA(cudaStreamBeginCaptureToGraph(q, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    int* saved = tl_capture_counter;
    tl_capture_counter = counter;
    int st = fn();
    tl_capture_counter = saved;
    cudaError_t ce = cudaStreamEndCapture(q, nullptr);
    if (st != AOSLO_OK) return st;
    if (ce != cudaSuccess) return set_cuda_error(ce, "cudaStreamEndCapture(body)");
    return AOSLO_OK;
}

// Enqueues the whole run on `q`.  graph = true: `q` is capturing and the loops become conditional nodes;
// graph = false (AOSLO_NO_GRAPH debug path): the same kernels run eagerly and the host evaluates the loop
// conditions from the control block (one synchronisation per decision).
int enqueue_run(RunPlan& P, cudaStream_t q, bool graph) {
    aoslo_ctx* c = P.ctx;
    const aoslo_config* cfg = P.cfg;
    const int H = cfg->H, W = cfg->W;
    RunCtl* ctl = c->d_ctl;
    const RunParams* rp = c->d_params;
    auto read_ctl = [&]() -> int {
        AOSLO_CUDA(cudaMemcpyAsync(c->h_ctl, ctl, sizeof(RunCtl), cudaMemcpyDeviceToHost, q));
        return stream_wait(c, q);
    };
    run_begin_kernel<<<1, 32, 0, q>>>(ctl, c->d_status, c->d_counters);
    AOSLO_LAUNCH_CHECK();
    if (P.thin_pixels) {
        // ---- seeds (:101) and thinning to the fixpoint (:115-120) in the pixel domain: one cooperative kernel + the
        // ordered listing of the survivors
        AOSLO_TRY(stage_thin_pixels(c, q, P.d_Rb, P.fdt, H, W, cfg->bx, cfg->by, cfg->init_blob_threshold,
                                    cfg->thinning_threshold, rp, ctl, c->d_pos[0], c->d_stable[0], P.cnt0));
    } else {
        // ---- seeds (:101), flags (:113)
        AOSLO_TRY(stage_seed(c, q, P.d_Rb, P.fdt, H, W, cfg->bx, cfg->by, cfg->init_blob_threshold, c->d_pos[0], c->cap,
                             P.cnt0, &rp->seed_thr, true));
        AOSLO_TRY(stage_set_stable(c, q, c->d_stable[0], P.inner, 0, P.cnt0));
        // ---- thinning to the fixpoint (:115-120)
        if (graph) {
            CondNode thin;
            AOSLO_TRY(new_cond_handle(q, &thin.handle));
            after_seed_kernel<<<1, 32, 0, q>>>(ctl, P.cnt0, thin.handle, 1);
            AOSLO_LAUNCH_CHECK();
            AOSLO_TRY(add_cond_node(q, cudaGraphCondTypeWhile, &thin));
            AOSLO_TRY(capture_body(c->cap_stream2, thin.body, &P.k_thin,
                                   [&]() { return enqueue_thin_round(P, c->cap_stream2, thin.handle, 1); }));
        } else {
            after_seed_kernel<<<1, 32, 0, q>>>(ctl, P.cnt0, 0, 0);
            AOSLO_LAUNCH_CHECK();
            for (;;) {
                AOSLO_TRY(read_ctl());
                if (!c->h_ctl->thin_cond) break;
                AOSLO_TRY(enqueue_thin_round(P, q, 0, 0));
            }
        }
    }
    // ---- all stable, LUT, gravity field, adding (:130-152), first resample period; GT cell list once per run
    int* n_dead = c->d_counters + 7;
    AOSLO_TRY(stage_lut(c, q, cfg->reception_field_size, cfg->mu_dist_func, cfg->sigma_dist_func,
                        cfg->lambda_dist_func, c->d_lut, rp));
    AOSLO_TRY(enqueue_resample(P, q, true));
    AOSLO_TRY(build_cell_list(c, q, c->cl_g, P.d_gt, P.n_gt, &rp->n_gt));
    AOSLO_TRY(stage_gt_bitmap(c, q, P.d_gt, P.n_gt, &rp->n_gt));
    // ---- the epoch loop (:173-382)
    if (!P.dynamic_loop && graph) {
        // unrolled: the branches are static, no control kernels
        if (P.timing) {
            stamp_kernel<<<1, 32, 0, q>>>(ctl, -1);
            AOSLO_LAUNCH_CHECK();
        }
        int n_rec = 0;
        for (int it = 0; it < cfg->epoch_num; ++it) {
            AOSLO_TRY(enqueue_step(P, q));
            if (P.timing) { stamp_kernel<<<1, 32, 0, q>>>(ctl, 2); AOSLO_LAUNCH_CHECK(); }
            if (it % cfg->fix_particles_frequency == 3) AOSLO_TRY(enqueue_resample(P, q, false));
            if (P.timing) { stamp_kernel<<<1, 32, 0, q>>>(ctl, 3); AOSLO_LAUNCH_CHECK(); }
            if (it % cfg->metric_measure_freq == 0 && P.want_records && n_rec < P.max_records) {
                AOSLO_TRY(enqueue_metrics(P, q, it, n_rec));
                ++n_rec;
            }
            if (P.timing) { stamp_kernel<<<1, 32, 0, q>>>(ctl, 4); AOSLO_LAUNCH_CHECK(); }
        }
    } else {
        auto step_and_mid = [&](cudaStream_t b, cudaGraphConditionalHandle h_res, cudaGraphConditionalHandle h_met,
                                int use_res, int use_met) -> int {
            AOSLO_TRY(enqueue_step(P, b));
            iter_mid_kernel<<<1, 32, 0, b>>>(ctl, rp, h_res, h_met, use_res, use_met, P.timing ? 1 : 0);
            AOSLO_LAUNCH_CHECK();
            return AOSLO_OK;
        };
        auto iter_tail = [&](cudaStream_t b, cudaGraphConditionalHandle h_loop, int use) -> int {
            if (cfg->early_stop) AOSLO_TRY(stage_dual_count_stable(c, b, c->cap, P.cnt0, &ctl->n_stable));
            iter_end_kernel<<<1, 32, 0, b>>>(ctl, rp, P.cnt0, n_dead, c->d_status, h_loop, use, P.timing ? 1 : 0);
            AOSLO_LAUNCH_CHECK();
            return AOSLO_OK;
        };
        if (graph) {
            CondNode loop;
            AOSLO_TRY(new_cond_handle(q, &loop.handle));
            loop_begin_kernel<<<1, 32, 0, q>>>(ctl, rp, P.cnt0, n_dead, c->d_status, loop.handle, 1, P.timing ? 1 : 0);
            AOSLO_LAUNCH_CHECK();
            AOSLO_TRY(add_cond_node(q, cudaGraphCondTypeWhile, &loop));
            cudaStream_t b = c->cap_stream2;
            AOSLO_TRY(capture_body(b, loop.body, &P.k_iter, [&]() -> int {
                // the IF handles belong to the body graph and are set by iter_mid_kernel, upstream of both nodes
                CondNode res, met;
                AOSLO_TRY(new_cond_handle(b, &res.handle));
                if (P.want_records) AOSLO_TRY(new_cond_handle(b, &met.handle));
                AOSLO_TRY(step_and_mid(b, res.handle, met.handle, 1, P.want_records ? 1 : 0));
                cudaStream_t b3 = c->cap_stream3;
                AOSLO_TRY(add_cond_node(b, cudaGraphCondTypeIf, &res));
                AOSLO_TRY(capture_body(b3, res.body, &P.k_res, [&]() -> int {
                    AOSLO_TRY(enqueue_resample(P, b3, false));
                    if (P.timing) { stamp_kernel<<<1, 32, 0, b3>>>(ctl, 3); AOSLO_LAUNCH_CHECK(); }
                    return AOSLO_OK;
                }));
                if (P.want_records) {
                    AOSLO_TRY(add_cond_node(b, cudaGraphCondTypeIf, &met));
                    AOSLO_TRY(capture_body(b3, met.body, &P.k_met, [&]() { return enqueue_metrics(P, b3, 0, -1); }));
                }
                return iter_tail(b, loop.handle, 1);
            }));
        } else {
            loop_begin_kernel<<<1, 32, 0, q>>>(ctl, rp, P.cnt0, n_dead, c->d_status, 0, 0, P.timing ? 1 : 0);
            AOSLO_LAUNCH_CHECK();
            for (;;) {
                AOSLO_TRY(read_ctl());
                if (!c->h_ctl->loop_cond) break;
                AOSLO_TRY(step_and_mid(q, 0, 0, 0, 0));
                AOSLO_TRY(read_ctl());
                if (c->h_ctl->do_resample) {
                    AOSLO_TRY(enqueue_resample(P, q, false));
                    if (P.timing) { stamp_kernel<<<1, 32, 0, q>>>(ctl, 3); AOSLO_LAUNCH_CHECK(); }
                }
                if (c->h_ctl->do_metrics && P.want_records) AOSLO_TRY(enqueue_metrics(P, q, 0, -1));
                AOSLO_TRY(iter_tail(q, 0, 0));
            }
        }
    }
    // ---- the survivors, compacted in order, are the result
    AOSLO_TRY(stage_compact(c, q, c->cap, P.cnt0, P.cnt1, P.hint));
    // ---- final reads: stable count, control block, status, records -> pinned memory
    AOSLO_TRY(stage_count_stable(c, q, c->d_stable[0], c->cap, &ctl->n_stable, P.cnt0));
    run_end_kernel<<<1, 32, 0, q>>>(ctl, P.cnt0);
    AOSLO_LAUNCH_CHECK();
    AOSLO_CUDA(cudaMemcpyAsync(c->h_ctl, ctl, sizeof(RunCtl), cudaMemcpyDeviceToHost, q));
    AOSLO_CUDA(cudaMemcpyAsync(c->h_pinned + 2, c->d_status, sizeof(int), cudaMemcpyDeviceToHost, q));
    if (P.want_records)
        AOSLO_CUDA(cudaMemcpyAsync(c->h_records, c->d_records, sizeof(aoslo_metrics_record) * P.max_records,
                                   cudaMemcpyDeviceToHost, q));
    return AOSLO_OK;
}

int k_class_of(int kk) { return kk <= 2 ? 2 : (kk <= 4 ? 4 : 8); }

int run_canonical(aoslo_ctx* ctx, cudaStream_t s, const aoslo_config* cfg, const void* d_Rb, const void* d_grad,
                  int field_dtype, const double* d_gt, int n_gt, aoslo_metrics_record* h_records, int max_records,
                  aoslo_run_summary* h_summary, float* h_stage_ms) {
    const int H = cfg->H, W = cfg->W;
    RunPlan P;
    P.ctx = ctx; P.cfg = cfg; P.d_Rb = d_Rb; P.d_grad = d_grad; P.fdt = field_dtype; P.d_gt = d_gt;
    P.n_gt = n_gt < 0 ? ctx->gt_cap : n_gt;          // launch-size bound only: the kernels read the device count
    P.want_records = h_records != nullptr && max_records > 0;
    P.max_records = P.want_records ? max_records : 0;
    P.timing = h_stage_ms != nullptr;
    P.dynamic_loop = cfg->early_stop != 0 || ctx->no_graph;
    P.thin_pixels = cfg->thinning_threshold <= 4.0;
    P.inner = (H - 2 * cfg->by) * (W - 2 * cfg->bx);
    if (P.inner < 1) return AOSLO_ERR_INVALID;
    if (P.inner > ctx->cap) P.inner = ctx->cap;
    if (ctx->loop_hint <= 0) {
        int h = 1024;
        while (h < P.inner / 16) h <<= 1;
        ctx->loop_hint = h;
    }
    P.hint = ctx->loop_hint < ctx->cap ? ctx->loop_hint : ctx->cap;
    P.cnt0 = ctx->d_counters + 4;
    P.cnt1 = ctx->d_counters + 5;
    // ---- value parameters -> device (stream ordered; the pinned staging copy is free again after the final sync)
    {
        const bool pit = cfg->dist_energy_type == AOSLO_ENERGY_PIT;
        RunParams& r = *ctx->h_params;
        memset(&r, 0, sizeof(r));
        r.seed_thr = cfg->init_blob_threshold; r.thin_thr = cfg->thinning_threshold;
        r.del_thr = cfg->threshhold_dist_del; r.add_thr = cfg->particle_call_threshold;
        r.lut_mu = cfg->mu_dist_func; r.lut_sigma = cfg->sigma_dist_func; r.lut_lam = cfg->lambda_dist_func;
        r.e0 = pit ? cfg->mu_dist_func : cfg->dist_alpha;
        r.e1 = pit ? cfg->sigma_dist_func : cfg->dist_sigma;
        r.e2 = pit ? cfg->lambda_dist_func : 0.0;
        r.lb = cfg->lambda_blob; r.ld = cfg->lambda_dist;
        r.kk = cfg->dist_n_neighbours + 1;
        r.fix_freq = cfg->fix_particles_frequency; r.metric_freq = cfg->metric_measure_freq;
        r.epoch_num = cfg->epoch_num; r.early_stop = cfg->early_stop; r.max_records = P.max_records; r.n_gt = n_gt;
        AOSLO_CUDA(cudaMemcpyAsync(ctx->d_params, ctx->h_params, sizeof(RunParams), cudaMemcpyHostToDevice, s));
        if (n_gt < 0) {
            params_ngt_kernel<<<1, 32, 0, s>>>(ctx->d_params, ctx->d_counters + 30);
            AOSLO_LAUNCH_CHECK();
        }
    }
    AOSLO_TRY(epoch_guard(ctx, s, (unsigned)cfg->epoch_num + 1u));
    ctx->epoch += (unsigned)cfg->epoch_num;          // upper estimate of the resolve calls of this run
    if (ctx->no_graph) {
        AOSLO_TRY(enqueue_run(P, s, false));
    } else {
        RunKey key;
        memset(&key, 0, sizeof(key));
        key.H = H; key.W = W; key.bx = cfg->bx; key.by = cfg->by; key.fdt = field_dtype;
        key.k_class = k_class_of(cfg->dist_n_neighbours + 1); key.energy = cfg->dist_energy_type;
        key.step_mode = cfg->step_mode; key.rfs = cfg->reception_field_size; key.window = cfg->particle_call_window;
        key.dynamic_loop = P.dynamic_loop; key.timing = P.timing; key.hint = P.hint;
        key.has_records = P.want_records; key.thin_pixels = P.thin_pixels;
        if (!P.dynamic_loop) {
            key.epoch_num = cfg->epoch_num; key.fix_freq = cfg->fix_particles_frequency;
            key.metric_freq = cfg->metric_measure_freq; key.max_rec = P.max_records;
        } else {
            key.max_rec = P.max_records;             // size of the record read-back node
        }
        key.Rb = d_Rb; key.grad = d_grad; key.gt = d_gt; key.records = ctx->d_records;
        RunGraph* hit = nullptr;
        for (RunGraph& g : ctx->graphs)
            if (g.exec && g.key.size() == sizeof(key) && memcmp(g.key.data(), &key, sizeof(key)) == 0) hit = &g;
        if (!hit) {
            const size_t kMaxGraphs = 24;          // a sweep cycles through (neighbour class) x (window) combinations
            if (ctx->graphs.size() < kMaxGraphs) {
                ctx->graphs.emplace_back();
                hit = &ctx->graphs.back();
            } else {
                hit = &ctx->graphs[0];
                for (RunGraph& g : ctx->graphs) if (g.stamp < hit->stamp) hit = &g;
                if (hit->exec) cudaGraphExecDestroy(hit->exec);
                if (hit->graph) cudaGraphDestroy(hit->graph);
                *hit = RunGraph();
            }
            // record on the context's own streams (the caller's may be the legacy default stream, which cannot
            // capture); nothing executes during capture
            cudaStream_t q = ctx->cap_stream;
            AOSLO_CUDA(cudaStreamBeginCapture(q, cudaStreamCaptureModeThreadLocal));
            int* saved = tl_capture_counter;
            tl_capture_counter = &P.k_static;
            int st = enqueue_run(P, q, true);
            tl_capture_counter = saved;
            cudaGraph_t graph = nullptr;
            cudaError_t ce = cudaStreamEndCapture(q, &graph);
            if (st != AOSLO_OK || ce != cudaSuccess) {
                // a failed body capture may have left the helper streams capturing
                cudaStreamEndCapture(ctx->cap_stream2, nullptr);
                cudaStreamEndCapture(ctx->cap_stream3, nullptr);
                cudaGetLastError();
                if (graph) cudaGraphDestroy(graph);
                hit->exec = nullptr;
                if (st != AOSLO_OK) return st;
                return set_cuda_error(ce, "cudaStreamEndCapture");
            }
            ce = cudaGraphInstantiate(&hit->exec, graph, 0);
            if (ce != cudaSuccess) {
                cudaGraphDestroy(graph);
                hit->exec = nullptr;
                return set_cuda_error(ce, "cudaGraphInstantiate");
            }
            hit->graph = graph;                      // kept alive with its conditional handles
            hit->key.assign((const unsigned char*)&key, (const unsigned char*)&key + sizeof(key));
            hit->static_kernels = P.k_static; hit->thin_kernels = P.k_thin; hit->iter_kernels = P.k_iter;
            hit->resample_kernels = P.k_res; hit->metrics_kernels = P.k_met;
        }
        hit->stamp = ++ctx->graph_clock;
        AOSLO_CUDA(cudaGraphLaunch(hit->exec, s));
        P.k_static = hit->static_kernels; P.k_thin = hit->thin_kernels; P.k_iter = hit->iter_kernels;
        P.k_res = hit->resample_kernels; P.k_met = hit->metrics_kernels;
    }
    // ---- host work that was waiting for a busy device (aoslo_stage_gt), then the one synchronisation of the run
    AOSLO_TRY(run_deferred_host_work(ctx));
    AOSLO_TRY(stream_wait(ctx, s));
    const RunCtl& r = *ctx->h_ctl;
    h_summary->status |= ctx->h_pinned[2];
    h_summary->n_seeds = r.n_seeds;
    h_summary->n_after_thinning = r.n_after_thinning;
    h_summary->n_thinning_rounds = r.thin_rounds;
    h_summary->n_stable = r.n_stable;
    h_summary->n_particles = r.n_final;
    int n_rec = 0, it_done = 0;
    if (P.dynamic_loop) {
        n_rec = r.n_rec;
        it_done = r.it;
    } else {
        for (int it = 0; it < cfg->epoch_num; ++it)
            if (it % cfg->metric_measure_freq == 0 && P.want_records && n_rec < P.max_records) ++n_rec;
        it_done = cfg->epoch_num;
    }
    if (r.n_seeds < 2 || r.n_after_thinning < 2) h_summary->status |= AOSLO_DEV_KNN_SHORT;
    h_summary->n_iterations = it_done;
    h_summary->n_records = n_rec;
    for (int i = 0; i < n_rec; ++i) h_records[i] = ctx->h_records[i];
    if (h_stage_ms)
        for (int i = 0; i < 5; ++i) h_stage_ms[i] = (float)((double)r.acc_ns[i] * 1e-6);
    if (!ctx->no_graph) {
        // kernels the graph executed: static nodes + the bodies as often as the control block says they ran
        long long k = P.k_static + (long long)r.thin_rounds * P.k_thin;
        if (P.dynamic_loop) {
            int n_res = 0;
            for (int it = 0; it < it_done; ++it) if (it % cfg->fix_particles_frequency == 3) ++n_res;
            k += (long long)it_done * P.k_iter + (long long)n_res * P.k_res + (long long)n_rec * P.k_met;
        }
        count_launches((int)k);
    }
    // launch-size hint of the next run on this context (performance only): ~1.5 x the final count, power of two,
    // changed only when it is off by a class so that similar runs keep sharing the cached graph
    {
        int want = 1024;
        while (want < r.n_final + r.n_final / 2 && want < ctx->cap) want <<= 1;
        if (want > ctx->loop_hint || want * 4 <= ctx->loop_hint) ctx->loop_hint = want;
    }
    ctx->last_n = r.n_final;
    ctx->last_buf = 0;
    return AOSLO_OK;
}

// sklearn tie order: stage calls with the count on the host (every kNN builds sklearn's tree on the host)
int run_sklearn(aoslo_ctx* ctx, cudaStream_t s, const aoslo_config* cfg, const void* d_Rb, const void* d_grad,
                int field_dtype, const double* d_gt, int n_gt, aoslo_metrics_record* h_records, int max_records,
                aoslo_run_summary* h_summary, float* h_stage_ms) {
    const int H = cfg->H, W = cfg->W, cap = ctx->cap;
    float acc_ms[5] = {0, 0, 0, 0, 0};
    // stage timing: CUDA events on the launch stream, resolved once at the end (no extra syncs)
    std::vector<cudaEvent_t>& pool = ctx->ev_pool;
    size_t ev_used = 0;
    std::vector<int> ev_bucket;
    auto mark = [&](int bucket) -> int {            // records an event; bucket = stage that just ended (-1: start)
        if (!h_stage_ms) return AOSLO_OK;
        if (ev_used == pool.size()) {
            cudaEvent_t e;
            AOSLO_CUDA(cudaEventCreate(&e));
            pool.push_back(e);
        }
        AOSLO_CUDA(cudaEventRecord(pool[ev_used], s));
        ev_bucket.push_back(bucket);
        ++ev_used;
        return AOSLO_OK;
    };
    int cur = 0, n = 0;
    int* d_cnt[2] = {ctx->d_counters + 4, ctx->d_counters + 5};   // count of buffer 0 / 1
    AOSLO_TRY(epoch_guard(ctx, s, (unsigned)cfg->epoch_num + 1u));
    // ---- seeds (:101) and flags (:113)
    AOSLO_TRY(stage_seed(ctx, s, d_Rb, field_dtype, H, W, cfg->bx, cfg->by, cfg->init_blob_threshold, ctx->d_pos[cur],
                         cap, d_cnt[cur]));
    AOSLO_TRY(read_int(ctx, s, d_cnt[cur], &n));
    if (n > cap) return AOSLO_ERR_CAPACITY;
    h_summary->n_seeds = n;
    // too few points for a neighbour search (sklearn raises): report it together with whatever the status word
    // already holds (the field's eigenvalue assert)
    auto short_exit = [&]() -> int {
        int word = 0;
        AOSLO_TRY(read_int(ctx, s, ctx->d_status, &word));
        h_summary->status = word | AOSLO_DEV_KNN_SHORT;
        ctx->last_n = 0;
        return AOSLO_OK;
    };
    if (n < 2) return short_exit();
    AOSLO_TRY(stage_set_stable(ctx, s, ctx->d_stable[cur], n, 0, nullptr));
    // ---- thinning to a fixpoint (:115-120)
    int rounds = 0;
    while (true) {
        int n_new = 0;
        AOSLO_TRY(stage_delete(ctx, s, ctx->d_pos[cur], ctx->d_stable[cur], n, d_Rb, field_dtype, H, W,
                               cfg->thinning_threshold, cfg->tie_mode, ctx->d_pos[cur ^ 1], ctx->d_stable[cur ^ 1],
                               nullptr, d_cnt[cur ^ 1], nullptr, nullptr, 0));
        AOSLO_TRY(read_int(ctx, s, d_cnt[cur ^ 1], &n_new));
        cur ^= 1;
        ++rounds;
        bool none = (n_new == n);
        n = n_new;
        if (none || n < 2) break;
    }
    h_summary->n_after_thinning = n;
    h_summary->n_thinning_rounds = rounds;
    if (n < 2) return short_exit();
    // ---- all stable, gravity field, adding (:130-152)
    AOSLO_TRY(stage_lut(ctx, s, cfg->reception_field_size, cfg->mu_dist_func, cfg->sigma_dist_func,
                        cfg->lambda_dist_func, ctx->d_lut));
    bool lut_ready = false;                          // radial table / flag of the LUT: derived once per run
    auto resample = [&]() -> int {
        AOSLO_TRY(stage_set_stable(ctx, s, ctx->d_stable[cur], n, 1, nullptr));
        AOSLO_TRY(stage_gravity(ctx, s, ctx->d_pos[cur], n, nullptr, ctx->d_lut, cfg->reception_field_size, H, W,
                                ctx->d_field, lut_ready));
        lut_ready = true;
        AOSLO_TRY(stage_add(ctx, s, ctx->d_field, H, W, cfg->bx, cfg->by, cfg->particle_call_window,
                            cfg->particle_call_threshold, ctx->d_pos[cur], ctx->d_stable[cur], n, nullptr, cap,
                            d_cnt[cur]));
        AOSLO_TRY(read_int(ctx, s, d_cnt[cur], &n));
        if (n > cap) return AOSLO_ERR_CAPACITY;
        return AOSLO_OK;
    };
    AOSLO_TRY(resample());
    AOSLO_TRY(build_cell_list(ctx, s, ctx->cl_g, d_gt, n_gt, nullptr));      // GT cell list once

    const int k = cfg->dist_n_neighbours;
    const bool pit = cfg->dist_energy_type == AOSLO_ENERGY_PIT;
    const double e0 = pit ? cfg->mu_dist_func : cfg->dist_alpha;
    const double e1 = pit ? cfg->sigma_dist_func : cfg->dist_sigma;
    const double e2 = pit ? cfg->lambda_dist_func : 0.0;
    int n_rec = 0, it_done = 0;
    AOSLO_TRY(mark(-1));
    for (int it = 0; it < cfg->epoch_num; ++it) {
        ++it_done;
        if (n < k + 1) { h_summary->status |= AOSLO_DEV_KNN_SHORT; break; }
        AOSLO_TRY(stage_force(ctx, s, ctx->d_pos[cur], ctx->d_stable[cur], n, k, cfg->dist_energy_type, e0, e1, e2,
                              cfg->tie_mode, ctx->d_force, nullptr));
        AOSLO_TRY(mark(0));
        AOSLO_TRY(stage_move(ctx, s, ctx->d_pos[cur], ctx->d_stable[cur], n, d_grad, field_dtype, H, W, ctx->d_force,
                             cfg->lambda_blob, cfg->lambda_dist, cfg->step_mode, nullptr));
        AOSLO_TRY(mark(1));
        AOSLO_TRY(stage_delete(ctx, s, ctx->d_pos[cur], ctx->d_stable[cur], n, d_Rb, field_dtype, H, W,
                               cfg->threshhold_dist_del, cfg->tie_mode, ctx->d_pos[cur ^ 1], ctx->d_stable[cur ^ 1],
                               nullptr, d_cnt[cur ^ 1], nullptr, nullptr, 1));
        cur ^= 1;
        AOSLO_TRY(read_int(ctx, s, d_cnt[cur], &n));
        AOSLO_TRY(mark(2));
        if (it % cfg->fix_particles_frequency == 3) AOSLO_TRY(resample());
        AOSLO_TRY(mark(3));
        if (it % cfg->metric_measure_freq == 0 && h_records && n_rec < max_records) {
            AOSLO_TRY(stage_metrics(ctx, s, ctx->d_pos[cur], ctx->d_stable[cur], n, d_gt, n_gt, true, false, d_Rb,
                                    field_dtype, H, W, cfg->bx, cfg->by, it, ctx->d_records + n_rec, ctx->d_dist_p2g,
                                    ctx->d_dist_g2p, nullptr));
            ++n_rec;
        }
        if (cfg->early_stop) {
            AOSLO_TRY(stage_count_stable(ctx, s, ctx->d_stable[cur], n, ctx->d_counters + 2, nullptr));
            int n_stable = 0;
            AOSLO_TRY(read_int(ctx, s, ctx->d_counters + 2, &n_stable));
            h_summary->n_stable = n_stable;
            // sum(stable) / len(particles) > 0.997 (:381) -- Python true division in float64
            if ((double)n_stable / (double)n > 0.997) { AOSLO_TRY(mark(4)); break; }
        }
        AOSLO_TRY(mark(4));
    }
    // ---- results: records, status -- one synchronisation
    if (n_rec)
        AOSLO_CUDA(cudaMemcpyAsync(ctx->h_records, ctx->d_records, sizeof(aoslo_metrics_record) * n_rec,
                                   cudaMemcpyDeviceToHost, s));
    if (!cfg->early_stop) {
        AOSLO_TRY(stage_count_stable(ctx, s, ctx->d_stable[cur], n, ctx->d_counters + 2, nullptr));
        AOSLO_CUDA(cudaMemcpyAsync(ctx->h_pinned + 1, ctx->d_counters + 2, sizeof(int), cudaMemcpyDeviceToHost, s));
    }
    AOSLO_CUDA(cudaMemcpyAsync(ctx->h_pinned + 2, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, s));
    AOSLO_TRY(stream_wait(ctx, s));
    if (!cfg->early_stop) h_summary->n_stable = ctx->h_pinned[1];
    h_summary->status |= ctx->h_pinned[2];
    for (int i = 0; i < n_rec; ++i) h_records[i] = ctx->h_records[i];
    h_summary->n_iterations = it_done;
    h_summary->n_particles = n;
    h_summary->n_records = n_rec;
    if (h_stage_ms) {
        for (size_t i = 1; i < ev_used; ++i) {
            if (ev_bucket[i] < 0) continue;
            float ms = 0.f;
            AOSLO_CUDA(cudaEventElapsedTime(&ms, pool[i - 1], pool[i]));
            acc_ms[ev_bucket[i]] += ms;
        }
        for (int i = 0; i < 5; ++i) h_stage_ms[i] = acc_ms[i];
    }
    ctx->last_n = n;
    ctx->last_buf = cur;
    return AOSLO_OK;
}

}  // namespace

extern "C" int aoslo_run(aoslo_ctx* ctx, void* stream, const aoslo_config* cfg, const void* d_Rb, const void* d_grad,
                         int field_dtype, const double* d_gt, int n_gt, aoslo_metrics_record* h_records,
                         int max_records, double* d_particles_out, uint8_t* d_stable_out,
                         aoslo_run_summary* h_summary, float* h_stage_ms) {
    if (!ctx || !cfg || !d_Rb || !d_grad || !d_gt || !h_summary) return AOSLO_ERR_INVALID;
    if (cfg->H != ctx->H || cfg->W != ctx->W || n_gt == 0 || n_gt > ctx->gt_cap) return AOSLO_ERR_INVALID;
    if (n_gt < 0 && cfg->tie_mode != AOSLO_TIE_CANONICAL) return AOSLO_ERR_INVALID;   // the host-tree path needs it
    if (cfg->dist_n_neighbours < 1 || cfg->dist_n_neighbours + 1 > AOSLO_KMAX) return AOSLO_ERR_UNSUPPORTED;
    if (cfg->particle_call_window < 1 || cfg->reception_field_size < 1 || cfg->reception_field_size > 64)
        return AOSLO_ERR_UNSUPPORTED;
    if (cfg->fix_particles_frequency < 1 || cfg->metric_measure_freq < 1 || cfg->epoch_num < 0) return AOSLO_ERR_INVALID;
    if (cfg->dist_energy_type != AOSLO_ENERGY_PIT && cfg->dist_energy_type != AOSLO_ENERGY_EXP) return AOSLO_ERR_UNSUPPORTED;
    if (max_records < 0) return AOSLO_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    AOSLO_CUDA(cudaSetDevice(ctx->device));
    if (h_records && max_records > ctx->max_records) {
        // grow the record buffers (wandb-style runs log every iteration); cached graphs refer to the old ones
        AOSLO_CUDA(cudaStreamSynchronize(s));
        for (RunGraph& g : ctx->graphs) {
            if (g.exec) cudaGraphExecDestroy(g.exec);
            if (g.graph) cudaGraphDestroy(g.graph);
        }
        ctx->graphs.clear();
        cudaFree(ctx->d_records);
        cudaFreeHost(ctx->h_records);
        ctx->d_records = nullptr; ctx->h_records = nullptr;
        ctx->max_records = 0;
        AOSLO_CUDA(cudaMalloc((void**)&ctx->d_records, sizeof(aoslo_metrics_record) * (size_t)max_records));
        AOSLO_CUDA(cudaHostAlloc((void**)&ctx->h_records, sizeof(aoslo_metrics_record) * (size_t)max_records,
                                 cudaHostAllocDefault));
        ctx->max_records = max_records;
    }
    memset(h_summary, 0, sizeof(*h_summary));
    int st;
    if (cfg->tie_mode == AOSLO_TIE_CANONICAL)
        st = run_canonical(ctx, s, cfg, d_Rb, d_grad, field_dtype, d_gt, n_gt, h_records, max_records, h_summary,
                           h_stage_ms);
    else
        st = run_sklearn(ctx, s, cfg, d_Rb, d_grad, field_dtype, d_gt, n_gt, h_records, max_records, h_summary,
                         h_stage_ms);
    {
        const int hw = run_deferred_host_work(ctx);      // error exits / the host-tree path: nothing stays pending
        if (st == AOSLO_OK) st = hw;
    }
    if (st != AOSLO_OK) return st;
    if (d_particles_out || d_stable_out)
        return aoslo_read_particles(ctx, stream, d_particles_out, d_stable_out, h_summary->n_particles);
    return AOSLO_OK;
}

extern "C" int aoslo_sweep(aoslo_ctx* ctx, void* stream, const aoslo_config* h_cfgs, int n_trials, const void* d_Rb,
                           const void* d_grad, int field_dtype, const double* d_gt, int n_gt,
                           aoslo_trial_result* h_results) {
    if (!ctx || !h_cfgs || !h_results || n_trials < 0) return AOSLO_ERR_INVALID;
    std::vector<aoslo_metrics_record> recs;
    for (int t = 0; t < n_trials; ++t) {
        const aoslo_config& cfg = h_cfgs[t];
        aoslo_trial_result& r = h_results[t];
        memset(&r, 0, sizeof(r));
        r.dice_coeff_1 = -INFINITY; r.dice_coeff_2 = -INFINITY;
        r.best_lsa_mean = INFINITY; r.best_blobEn_mean = -INFINITY;
        if (cfg.metric_measure_freq < 1 || cfg.epoch_num < 0) { r.error = AOSLO_ERR_INVALID; continue; }
        const int max_rec = cfg.epoch_num / cfg.metric_measure_freq + 2;
        recs.resize((size_t)max_rec);
        AOSLO_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(int), (cudaStream_t)stream));
        r.error = aoslo_run(ctx, stream, &cfg, d_Rb, d_grad, field_dtype, d_gt, n_gt, recs.data(), max_rec, nullptr,
                            nullptr, &r.summary, nullptr);
        if (r.error == AOSLO_ERR_CUDA) return r.error;          // a CUDA error is not a property of the trial
        if (r.error != AOSLO_OK) continue;
        for (int i = 0; i < r.summary.n_records; ++i) {
            const aoslo_metrics_record& m = recs[i];
            // calc_acc_metrics (utils.py:378-402) + the tracking of :335-359
            const double d1 = 2.0 * m.tp[0] / (2.0 * m.tp[0] + m.fp[0] + m.fn[0]);
            const double d2 = 2.0 * m.tp[1] / (2.0 * m.tp[1] + m.fp[1] + m.fn[1]);
            if (r.dice_coeff_1 < d1 * 0.995) r.dice_coeff_1 = d1;
            if (r.dice_coeff_2 < d2 * 0.995) r.dice_coeff_2 = d2;
            const double lsa_mean = m.sum_p2g / m.n_particles + m.sum_g2p / m.n_gt;
            const double blob_mean = m.blob_energy_sum / m.n_particles;
            if (r.best_lsa_mean > lsa_mean) r.best_lsa_mean = lsa_mean;
            if (r.best_blobEn_mean < blob_mean) r.best_blobEn_mean = blob_mean;
        }
        if (r.summary.n_records > 0) r.last = recs[r.summary.n_records - 1];
    }
    return AOSLO_OK;
}

extern "C" int aoslo_read_particles(aoslo_ctx* ctx, void* stream, double* d_particles_out, uint8_t* d_stable_out,
                                    int n) {
    if (!ctx || n < 0 || n > ctx->last_n) return AOSLO_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    const int b = ctx->last_buf;
    if (d_particles_out && n)
        AOSLO_CUDA(cudaMemcpyAsync(d_particles_out, ctx->d_pos[b], sizeof(double) * 2 * n, cudaMemcpyDeviceToDevice, s));
    if (d_stable_out && n)
        AOSLO_CUDA(cudaMemcpyAsync(d_stable_out, ctx->d_stable[b], n, cudaMemcpyDeviceToDevice, s));
    return AOSLO_OK;
}
