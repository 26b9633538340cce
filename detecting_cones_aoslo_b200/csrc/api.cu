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

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_last_err, sizeof(g_last_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
    cudaGetLastError();
    return AOSLO_ERR_CUDA;
}

// stage drivers defined in particles.cu (trailing d_n: optional device-resident particle count)
int stage_knn(aoslo_ctx*, cudaStream_t, const double*, int, const double*, int, int, int, int32_t*, double*, bool,
              const int* d_n);
int stage_delete(aoslo_ctx*, cudaStream_t, const double*, const uint8_t*, int, const void*, int, int, int, double, int,
                 double*, uint8_t*, uint8_t*, int*, int*, const int* d_n, int mode);
int stage_force_move(aoslo_ctx*, cudaStream_t, const double*, const uint8_t*, int, int, int, double, double, double,
                     const void*, int, int, int, double, double, int, double*, const int* d_n, bool links_ready);
int stage_seed(aoslo_ctx*, cudaStream_t, const void*, int, int, int, int, int, double, double*, int, int*);
int stage_lut(aoslo_ctx*, cudaStream_t, int, double, double, double, double*);
int stage_gravity(aoslo_ctx*, cudaStream_t, const double*, int, int, int, double*, bool packed_ready);
int stage_add(aoslo_ctx*, cudaStream_t, const double*, int, int, int, int, int, double, double*, uint8_t*, int,
              const int*, int, int*);
int stage_force(aoslo_ctx*, cudaStream_t, const double*, const uint8_t*, int, int, int, double, double, double, int,
                double*, const int* d_n);
int stage_move(aoslo_ctx*, cudaStream_t, double*, const uint8_t*, int, const void*, int, int, int, const double*,
               double, double, int, const int* d_n);
int stage_metrics(aoslo_ctx*, cudaStream_t, const double*, const uint8_t*, int, const double*, int, bool, bool,
                  const void*, int, int, int, int, int, int, aoslo_metrics_record*, double*, double*, const int* d_n);
int stage_set_stable(aoslo_ctx*, cudaStream_t, uint8_t*, int, uint8_t, const int* d_n);
int stage_count_stable(aoslo_ctx*, cudaStream_t, const uint8_t*, int, int*, const int* d_n);
int coop_blocks_for_delete(int num_sms);
int stage_classify(aoslo_ctx*, cudaStream_t, const double*, int, const double*, const double*, int, int, int, int, int,
                   double, uint8_t*, uint8_t*);
int stage_prepare(aoslo_ctx*, cudaStream_t, const uint8_t*, int, int, int, int, int, int, int, int, int, uint8_t*, int,
                  const double*, int, double*, int*);
int sqrt_inplace(cudaStream_t, double*, size_t);

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
    A(alloc_cell_list(c->cl_p, H, W, cap, shift));
    A(alloc_cell_list(c->cl_g, H, W, gcap, shift));
    {
        int lshift = 0;                              // 0 = density adaptive
        if (const char* e = getenv("AOSLO_LINK_SHIFT")) { int v = atoi(e); if (v >= 1 && v <= 6) lshift = v; }
        A(alloc_links(c->lk_p, H, W, cap, lshift));
    }
    A(dmalloc(&c->d_pos_tmp, (size_t)cap * 2));
    A(dmalloc(&c->d_epoch, 1));
    A(dmalloc(&c->d_active, cap));
    A(dmalloc(&c->d_active_b, cap));
    A(dmalloc(&c->d_ns_list, cap));
    A(dmalloc(&c->d_knn_idx, kn));
    A(dmalloc(&c->d_knn_d2, kn));
    A(dmalloc(&c->d_block_sums, kMaxScanBlocks + 1));
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
    A(dmalloc(&c->d_lut_packed, 12288));         // >= grav_table_words(63): rank table of the gravity gather
    A(dmalloc(&c->d_field, (size_t)H * W));
    A(dmalloc(&c->d_win, (size_t)H * W));
    c->kd.cap = 0;   // sklearn-order KD tree mirrors are allocated on first use (aoslo::kd_reserve)
    for (int i = 0; i < 12 && st == AOSLO_OK; ++i)
        if (cudaEventCreate(&c->ev[i]) != cudaSuccess) st = AOSLO_ERR_CUDA;
    if (st == AOSLO_OK && cudaEventCreateWithFlags(&c->ev_block, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess)
        st = AOSLO_ERR_CUDA;
    if (st == AOSLO_OK && cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking) != cudaSuccess) st = AOSLO_ERR_CUDA;
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
    free_cell_list(c->cl_p); free_cell_list(c->cl_g);
    cudaFree(c->lk_p.head); cudaFree(c->lk_p.next); cudaFree(c->d_pos_tmp); cudaFree(c->d_active); cudaFree(c->d_active_b); cudaFree(c->d_ns_list);
    cudaFree(c->d_knn_idx); cudaFree(c->d_knn_d2); cudaFree(c->d_block_sums);
    cudaFree(c->d_touch); cudaFree(c->d_del); cudaFree(c->d_rowstate); cudaFree(c->d_force);
    cudaFree(c->d_pos[0]); cudaFree(c->d_pos[1]); cudaFree(c->d_stable[0]); cudaFree(c->d_stable[1]);
    cudaFree(c->d_dist_p2g); cudaFree(c->d_dist_g2p); cudaFree(c->d_partials); cudaFree(c->d_record);
    cudaFreeHost(c->h_record); cudaFree(c->d_records); cudaFreeHost(c->h_records); cudaFree(c->d_lut); cudaFree(c->d_lut_packed); cudaFree(c->d_field); cudaFree(c->d_win);
    kd_free(c->kd); kd_free(c->kd2); cudaFreeHost(c->kd_stage);
    for (int i = 0; i < 12; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    if (c->ev_block) cudaEventDestroy(c->ev_block);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
    cudaFree(c->d_epoch);
    cudaGetLastError();
    delete c;
    return AOSLO_OK;
}

extern "C" int aoslo_ctx_capacity(aoslo_ctx* ctx) { return ctx ? ctx->cap : 0; }

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

extern "C" int aoslo_prepare_inputs(aoslo_ctx* ctx, void* stream, const uint8_t* d_raw_img, int H0, int W0,
                                    int pitch0_bytes, int x_min, int x_max, int y_min, int y_max, int bx, int by,
                                    uint8_t* d_img_out, int out_pitch_bytes, const double* d_gt_minus1, int n_gt_raw,
                                    double* d_gt_out, int* h_n_gt_out) {
    if (!ctx || !d_raw_img || !d_img_out || !h_n_gt_out) return AOSLO_ERR_INVALID;
    if (n_gt_raw > 0 && (!d_gt_minus1 || !d_gt_out)) return AOSLO_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    AOSLO_CUDA(cudaMemsetAsync(ctx->d_counters, 0, sizeof(int), s));
    AOSLO_TRY(stage_prepare(ctx, s, d_raw_img, H0, W0, pitch0_bytes, x_min, x_max, y_min, y_max, bx, by, d_img_out,
                            out_pitch_bytes, d_gt_minus1, n_gt_raw, d_gt_out, ctx->d_counters));
    return read_int(ctx, s, ctx->d_counters, h_n_gt_out);
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
    if (!ctx || !d_particles || !d_lut || !d_field || H != ctx->H || W != ctx->W || n < 0) return AOSLO_ERR_INVALID;
    if (n > ctx->cap) return AOSLO_ERR_CAPACITY;
    cudaStream_t s = (cudaStream_t)stream;
    AOSLO_TRY(build_cell_list(ctx, s, ctx->cl_p, d_particles, n, nullptr));
    return stage_gravity(ctx, s, d_lut, rfs, H, W, d_field, false);
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
// (reference pipeline_with_delitions.py:173-253).  Same code path as inside aoslo_run, incl. the
// non-stable-list fast path and its all-rows fallback in the canonical tie order.
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
    const double* moved;
    int mode;
    if (cfg->tie_mode == AOSLO_TIE_CANONICAL) {
        AOSLO_TRY(stage_force_move(ctx, s, d_particles, d_stable, n, k, cfg->dist_energy_type, e0, e1, e2, d_grad,
                                   field_dtype, cfg->H, cfg->W, cfg->lambda_blob, cfg->lambda_dist, cfg->step_mode,
                                   ctx->d_pos_tmp, nullptr, false));
        moved = ctx->d_pos_tmp;
        mode = 2;
    } else {
        AOSLO_CUDA(cudaMemcpyAsync(ctx->d_pos_tmp, d_particles, sizeof(double) * 2 * n, cudaMemcpyDeviceToDevice, s));
        AOSLO_TRY(stage_force(ctx, s, ctx->d_pos_tmp, d_stable, n, k, cfg->dist_energy_type, e0, e1, e2,
                              cfg->tie_mode, ctx->d_force, nullptr));
        AOSLO_TRY(stage_move(ctx, s, ctx->d_pos_tmp, d_stable, n, d_grad, field_dtype, cfg->H, cfg->W, ctx->d_force,
                             cfg->lambda_blob, cfg->lambda_dist, cfg->step_mode, nullptr));
        moved = ctx->d_pos_tmp;
        mode = 1;
    }
    AOSLO_TRY(stage_delete(ctx, s, moved, d_stable, n, d_Rb, field_dtype, cfg->H, cfg->W, cfg->threshhold_dist_del,
                           cfg->tie_mode, d_particles_out, d_stable_out, nullptr, ctx->d_counters, nullptr, nullptr,
                           mode));
    return read_int(ctx, s, ctx->d_counters, h_n_out);
}

// ---------------------------------------------------------------------------------------
// whole run
// ---------------------------------------------------------------------------------------
// Control flow = reference pipeline_with_delitions.py:101-382.  The particle count lives in a
// device scalar (double-buffered with the particle arrays); in the canonical tie mode the
// host only synchronises where it needs a decision: the thinning fixpoint test, after every
// resample (to refresh its upper bound of the count), for the early-stop test when that is
// enabled, and once at the end for the records.  The sklearn-order tie mode needs the count
// on the host for every kNN (host tree build) and synchronises per stage.
extern "C" int aoslo_run(aoslo_ctx* ctx, void* stream, const aoslo_config* cfg, const void* d_Rb, const void* d_grad,
                         int field_dtype, const double* d_gt, int n_gt, aoslo_metrics_record* h_records,
                         int max_records, double* d_particles_out, uint8_t* d_stable_out,
                         aoslo_run_summary* h_summary, float* h_stage_ms) {
    if (!ctx || !cfg || !d_Rb || !d_grad || !d_gt || !h_summary) return AOSLO_ERR_INVALID;
    if (cfg->H != ctx->H || cfg->W != ctx->W || n_gt < 1 || n_gt > ctx->gt_cap) return AOSLO_ERR_INVALID;
    if (cfg->dist_n_neighbours < 1 || cfg->dist_n_neighbours + 1 > AOSLO_KMAX) return AOSLO_ERR_UNSUPPORTED;
    if (cfg->particle_call_window < 1 || cfg->reception_field_size < 1 || cfg->reception_field_size > 64)
        return AOSLO_ERR_UNSUPPORTED;
    if (cfg->fix_particles_frequency < 1 || cfg->metric_measure_freq < 1) return AOSLO_ERR_INVALID;
    if (max_records > ctx->max_records) max_records = ctx->max_records;
    cudaStream_t s = (cudaStream_t)stream;
    const int H = cfg->H, W = cfg->W, cap = ctx->cap;
    const bool on_device = (cfg->tie_mode == AOSLO_TIE_CANONICAL);   // count stays on the device
    memset(h_summary, 0, sizeof(*h_summary));
    AOSLO_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(int), s));
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
    if (n < 2) { h_summary->status = AOSLO_DEV_KNN_SHORT; return AOSLO_OK; }
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
    if (n < 2) { h_summary->status = AOSLO_DEV_KNN_SHORT; return AOSLO_OK; }
    // ---- all stable, gravity field, adding (:130-152)
    AOSLO_TRY(stage_lut(ctx, s, cfg->reception_field_size, cfg->mu_dist_func, cfg->sigma_dist_func,
                        cfg->lambda_dist_func, ctx->d_lut));
    bool packed_ready = false;                       // rank-packed LUT: derived once per run
    // resample; with `sync` the exact count comes back to the host (it bounds later launch sizes),
    // without it `n` stays an upper-bound hint -- every kernel takes the real count from the device
    cudaStream_t qs = s;                             // stream the loop is enqueued on (the capture stream while recording)
    auto resample = [&](const int* d_n_in, bool sync) -> int {
        AOSLO_TRY(stage_set_stable(ctx, qs, ctx->d_stable[cur], n, 1, d_n_in));
        AOSLO_TRY(build_cell_list(ctx, qs, ctx->cl_p, ctx->d_pos[cur], n, d_n_in));
        AOSLO_TRY(stage_gravity(ctx, qs, ctx->d_lut, cfg->reception_field_size, H, W, ctx->d_field, packed_ready));
        packed_ready = true;
        AOSLO_TRY(stage_add(ctx, qs, ctx->d_field, H, W, cfg->bx, cfg->by, cfg->particle_call_window,
                            cfg->particle_call_threshold, ctx->d_pos[cur], ctx->d_stable[cur], n, d_n_in, cap,
                            ctx->d_counters + 6));
        // the appended count lands in a scratch scalar, then becomes this buffer's count
        AOSLO_CUDA(cudaMemcpyAsync(d_cnt[cur], ctx->d_counters + 6, sizeof(int), cudaMemcpyDeviceToDevice, qs));
        if (sync) {
            AOSLO_TRY(read_int(ctx, qs, d_cnt[cur], &n));
            if (n > cap) return AOSLO_ERR_CAPACITY;
        }
        return AOSLO_OK;
    };
    AOSLO_TRY(resample(nullptr, true));
    AOSLO_TRY(build_cell_list(ctx, qs, ctx->cl_g, d_gt, n_gt, nullptr));      // GT cell list once

    const int k = cfg->dist_n_neighbours;
    const bool pit = cfg->dist_energy_type == AOSLO_ENERGY_PIT;
    int n_rec = 0, it_done = 0;
    // The epoch loop.  In the canonical tie order without the early stop it contains no host decision:
    // it is then captured once into a CUDA graph (cached in the context, keyed by everything baked into
    // the nodes) and replayed -- one graph launch instead of ~350 kernel launches per run.
    const bool graphable = on_device && !cfg->early_stop && !h_stage_ms && n >= k + 1 &&
                           !getenv("AOSLO_NO_GRAPH");
    if (graphable && cur != 0) {                     // the graph is recorded for buffer parity 0
        AOSLO_CUDA(cudaMemcpyAsync(ctx->d_pos[0], ctx->d_pos[1], sizeof(double) * 2 * n, cudaMemcpyDeviceToDevice, qs));
        AOSLO_CUDA(cudaMemcpyAsync(ctx->d_stable[0], ctx->d_stable[1], n, cudaMemcpyDeviceToDevice, qs));
        AOSLO_CUDA(cudaMemcpyAsync(d_cnt[0], d_cnt[1], sizeof(int), cudaMemcpyDeviceToDevice, qs));
        cur = 0;
    }
    auto enqueue_loop = [&](bool syncfree) -> int {
        bool links_valid = false;                    // ctx->lk_p describes d_pos[cur]
        if (!syncfree) AOSLO_TRY(mark(-1));
        for (int it = 0; it < cfg->epoch_num; ++it) {
            ++it_done;
            if (!syncfree && n < k + 1) { h_summary->status |= AOSLO_DEV_KNN_SHORT; break; }
            // `n` is exact in the synchronising mode and an upper-bound hint of the device count otherwise
            const int* dn = on_device ? d_cnt[cur] : nullptr;
            const double e0 = pit ? cfg->mu_dist_func : cfg->dist_alpha;
            const double e1 = pit ? cfg->sigma_dist_func : cfg->dist_sigma;
            const double e2 = pit ? cfg->lambda_dist_func : 0.0;
            const double* moved;
            if (on_device) {
                // non-stable list, warp-kNN on the linked cells + force + Jacobi move into d_pos_tmp
                AOSLO_TRY(stage_force_move(ctx, qs, ctx->d_pos[cur], ctx->d_stable[cur], n, k, cfg->dist_energy_type,
                                           e0, e1, e2, d_grad, field_dtype, H, W, cfg->lambda_blob, cfg->lambda_dist,
                                           cfg->step_mode, ctx->d_pos_tmp, dn, links_valid));
                links_valid = false;
                moved = ctx->d_pos_tmp;
                if (!syncfree) { AOSLO_TRY(mark(0)); AOSLO_TRY(mark(1)); }
            } else {
                AOSLO_TRY(stage_force(ctx, qs, ctx->d_pos[cur], ctx->d_stable[cur], n, k, cfg->dist_energy_type, e0,
                                      e1, e2, cfg->tie_mode, ctx->d_force, dn));
                AOSLO_TRY(mark(0));
                AOSLO_TRY(stage_move(ctx, qs, ctx->d_pos[cur], ctx->d_stable[cur], n, d_grad, field_dtype, H, W,
                                     ctx->d_force, cfg->lambda_blob, cfg->lambda_dist, cfg->step_mode, dn));
                AOSLO_TRY(mark(1));
                moved = ctx->d_pos[cur];
            }
            AOSLO_TRY(stage_delete(ctx, qs, moved, ctx->d_stable[cur], n, d_Rb, field_dtype, H, W,
                                   cfg->threshhold_dist_del, cfg->tie_mode, ctx->d_pos[cur ^ 1],
                                   ctx->d_stable[cur ^ 1], nullptr, d_cnt[cur ^ 1], nullptr, dn, on_device ? 2 : 1));
            cur ^= 1;
            if (!on_device) AOSLO_TRY(read_int(ctx, qs, d_cnt[cur], &n));
            dn = on_device ? d_cnt[cur] : nullptr;
            if (!syncfree) AOSLO_TRY(mark(2));
            if (it % cfg->fix_particles_frequency == 3) AOSLO_TRY(resample(dn, !syncfree));
            if (!syncfree) AOSLO_TRY(mark(3));
            if (it % cfg->metric_measure_freq == 0 && h_records && n_rec < max_records) {
                AOSLO_TRY(stage_metrics(ctx, qs, ctx->d_pos[cur], ctx->d_stable[cur], n, d_gt, n_gt, true, false, d_Rb,
                                        field_dtype, H, W, cfg->bx, cfg->by, it, ctx->d_records + n_rec,
                                        ctx->d_dist_p2g, ctx->d_dist_g2p, dn));
                links_valid = on_device;             // the metrics pass left links of d_pos[cur] behind
                ++n_rec;
            }
            if (cfg->early_stop) {
                int n_stable = 0;
                AOSLO_TRY(stage_count_stable(ctx, qs, ctx->d_stable[cur], n, ctx->d_counters + 2, dn));
                AOSLO_CUDA(cudaMemcpyAsync(ctx->h_pinned + 1, ctx->d_counters + 2, sizeof(int), cudaMemcpyDeviceToHost, qs));
                AOSLO_TRY(read_int(ctx, qs, d_cnt[cur], &n));
                n_stable = ctx->h_pinned[1];
                h_summary->n_stable = n_stable;
                // sum(stable) / len(particles) > 0.997 (:381) -- Python true division in float64
                if ((double)n_stable / (double)n > 0.997) break;
            }
            if (!syncfree) AOSLO_TRY(mark(4));
        }
        return AOSLO_OK;
    };
    if (!graphable) {
        AOSLO_TRY(enqueue_loop(false));
    } else {
        // launch-size hint: next power of two above 1.25 n, so that runs of similar size share the graph
        int hint = 1024;
        while (hint < n + n / 4 && hint < cap) hint <<= 1;
        if (hint > cap) hint = cap;
        struct Key {
            aoslo_config cfg; const void* Rb; const void* grad; const void* gt; int n_gt, fdt, hint, max_rec;
            const void* recs;
        } key;
        memset(&key, 0, sizeof(key));
        key.cfg = *cfg; key.Rb = d_Rb; key.grad = d_grad; key.gt = d_gt; key.n_gt = n_gt; key.fdt = field_dtype;
        key.hint = hint; key.max_rec = max_records; key.recs = h_records ? (const void*)ctx->d_records : nullptr;
        const bool hit = ctx->graph_exec && ctx->graph_key.size() == sizeof(key) &&
                         memcmp(ctx->graph_key.data(), &key, sizeof(key)) == 0;
        const int n_exact = n;
        if (!hit) {
            if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
            n = hint;
            // record on the context's own stream (the caller's may be the legacy default stream, which
            // cannot capture); nothing executes during capture, the graph is launched on the caller's stream
            qs = ctx->cap_stream;
            AOSLO_CUDA(cudaStreamBeginCapture(qs, cudaStreamCaptureModeThreadLocal));
            int st_loop = enqueue_loop(true);
            cudaGraph_t graph = nullptr;
            cudaError_t ce = cudaStreamEndCapture(qs, &graph);
            qs = s;
            if (st_loop != AOSLO_OK) { if (graph) cudaGraphDestroy(graph); return st_loop; }
            if (ce != cudaSuccess) return set_cuda_error(ce, "cudaStreamEndCapture");
            size_t n_nodes = 0;
            cudaGraphGetNodes(graph, nullptr, &n_nodes);
            std::vector<cudaGraphNode_t> nodes(n_nodes);
            cudaGraphGetNodes(graph, nodes.data(), &n_nodes);
            int kernels = 0;
            for (size_t i = 0; i < n_nodes; ++i) {
                cudaGraphNodeType t;
                if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ++kernels;
            }
            ce = cudaGraphInstantiate(&ctx->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) { ctx->graph_exec = nullptr; return set_cuda_error(ce, "cudaGraphInstantiate"); }
            ctx->graph_kernel_nodes = kernels;
            ctx->graph_key.assign((const unsigned char*)&key, (const unsigned char*)&key + sizeof(key));
            // capture ran the bookkeeping of the loop once; replaying needs the same values
        } else {
            // bookkeeping the captured loop would have done
            for (int it = 0; it < cfg->epoch_num; ++it)
                if (it % cfg->metric_measure_freq == 0 && h_records && n_rec < max_records) ++n_rec;
            it_done = cfg->epoch_num;
            if (cfg->epoch_num & 1) cur ^= 1;
            ctx->epoch += (unsigned)cfg->epoch_num;
        }
        n = n_exact;
        AOSLO_CUDA(cudaGraphLaunch(ctx->graph_exec, s));
        count_launches(ctx->graph_kernel_nodes);
    }
    // ---- results: records, final particles, status -- one synchronisation
    if (n_rec)
        AOSLO_CUDA(cudaMemcpyAsync(ctx->h_records, ctx->d_records, sizeof(aoslo_metrics_record) * n_rec,
                                   cudaMemcpyDeviceToHost, s));
    if (!cfg->early_stop) {
        AOSLO_TRY(stage_count_stable(ctx, s, ctx->d_stable[cur], n, ctx->d_counters + 2, on_device ? d_cnt[cur] : nullptr));
        AOSLO_CUDA(cudaMemcpyAsync(ctx->h_pinned + 1, ctx->d_counters + 2, sizeof(int), cudaMemcpyDeviceToHost, s));
    }
    AOSLO_CUDA(cudaMemcpyAsync(ctx->h_pinned + 2, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, s));
    AOSLO_TRY(read_int(ctx, s, d_cnt[cur], &n));
    if (!cfg->early_stop) h_summary->n_stable = ctx->h_pinned[1];
    h_summary->status |= ctx->h_pinned[2];
    for (int i = 0; i < n_rec; ++i) h_records[i] = ctx->h_records[i];
    if (d_particles_out)
        AOSLO_CUDA(cudaMemcpyAsync(d_particles_out, ctx->d_pos[cur], sizeof(double) * 2 * n, cudaMemcpyDeviceToDevice, s));
    if (d_stable_out)
        AOSLO_CUDA(cudaMemcpyAsync(d_stable_out, ctx->d_stable[cur], n, cudaMemcpyDeviceToDevice, s));
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
    return AOSLO_OK;
}
