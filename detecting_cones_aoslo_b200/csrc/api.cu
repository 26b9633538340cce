// api.cu -- context management, the C ABI wrappers of the stage kernels and the whole-run
// driver `aoslo_run` (reference pipeline_with_delitions.py:12-403 after crop / pad).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace aoslo {

static thread_local char g_last_err[512] = "";

int set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_last_err, sizeof(g_last_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
    cudaGetLastError();
    return AOSLO_ERR_CUDA;
}

// stage drivers defined in particles.cu
int stage_knn(aoslo_ctx*, cudaStream_t, const double*, int, const double*, int, int, int, int32_t*, double*, bool);
int stage_delete(aoslo_ctx*, cudaStream_t, const double*, const uint8_t*, int, const void*, int, int, int, double, int,
                 double*, uint8_t*, uint8_t*, int*, int*);
int stage_seed(aoslo_ctx*, cudaStream_t, const void*, int, int, int, int, int, double, double*, int, int*);
int stage_lut(aoslo_ctx*, cudaStream_t, int, double, double, double, double*);
int stage_gravity(aoslo_ctx*, cudaStream_t, const double*, int, int, int, double*);
int stage_add(aoslo_ctx*, cudaStream_t, const double*, int, int, int, int, int, double, double*, uint8_t*, int,
              const int*, int, int*);
int stage_force(aoslo_ctx*, cudaStream_t, const double*, const uint8_t*, int, int, int, double, double, double, int,
                double*);
int stage_move(aoslo_ctx*, cudaStream_t, double*, const uint8_t*, int, const void*, int, int, int, const double*,
               double, double, int);
int stage_metrics(aoslo_ctx*, cudaStream_t, const double*, const uint8_t*, int, const double*, int, bool, bool,
                  const void*, int, int, int, int, int, int, aoslo_metrics_record*, double*, double*);
int stage_set_stable(aoslo_ctx*, cudaStream_t, uint8_t*, int, uint8_t);
int stage_count_stable(aoslo_ctx*, cudaStream_t, const uint8_t*, int, int*);
int coop_blocks_for_delete(int num_sms);
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
    return AOSLO_OK;
}

static void free_cell_list(CellList& cl) {
    cudaFree(cl.count); cudaFree(cl.start); cudaFree(cl.slot); cudaFree(cl.cell_of);
    cudaFree(cl.sorted_idx); cudaFree(cl.sorted_pos);
}

// read a device int through the pinned mailbox (one stream sync)
static int read_int(aoslo_ctx* ctx, cudaStream_t s, const int* d, int* out) {
    AOSLO_CUDA(cudaMemcpyAsync(ctx->h_pinned, d, sizeof(int), cudaMemcpyDeviceToHost, s));
    AOSLO_CUDA(cudaStreamSynchronize(s));
    *out = ctx->h_pinned[0];
    return AOSLO_OK;
}

}  // namespace aoslo

using namespace aoslo;

extern "C" const char* aoslo_version(void) { return "aoslo_b200 0.1 (sm_100a)"; }
extern "C" const char* aoslo_last_cuda_error(void) { return g_last_err; }
extern "C" int aoslo_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
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
    A(dmalloc(&c->d_partials, 4096));
    A(dmalloc(&c->d_record, 1));
    if (st == AOSLO_OK && cudaHostAlloc((void**)&c->h_record, sizeof(aoslo_metrics_record), cudaHostAllocDefault) != cudaSuccess)
        st = AOSLO_ERR_CUDA;
    A(dmalloc(&c->d_lut, 64 * 64));
    A(dmalloc(&c->d_field, (size_t)H * W));
    A(dmalloc(&c->d_win, (size_t)H * W));
    // sklearn-order KD tree mirrors
    KdTreeDev& kd = c->kd;
    kd.cap = cap;
    size_t nn = (size_t)2 * cap + 2;
    A(dmalloc(&kd.idx_array, cap));
    A(dmalloc(&kd.node_start, nn));
    A(dmalloc(&kd.node_end, nn));
    A(dmalloc(&kd.node_bounds, nn * 4));
    if (st == AOSLO_OK) {
        bool ok = cudaHostAlloc((void**)&kd.h_points, sizeof(double) * 2 * cap, cudaHostAllocDefault) == cudaSuccess &&
                  cudaHostAlloc((void**)&kd.h_idx_array, sizeof(int32_t) * cap, cudaHostAllocDefault) == cudaSuccess &&
                  cudaHostAlloc((void**)&kd.h_node_start, sizeof(int32_t) * nn, cudaHostAllocDefault) == cudaSuccess &&
                  cudaHostAlloc((void**)&kd.h_node_end, sizeof(int32_t) * nn, cudaHostAllocDefault) == cudaSuccess &&
                  cudaHostAlloc((void**)&kd.h_node_bounds, sizeof(double) * 4 * nn, cudaHostAllocDefault) == cudaSuccess;
        if (!ok) st = set_cuda_error(cudaGetLastError(), "cudaHostAlloc(kd)");
    }
    for (int i = 0; i < 12 && st == AOSLO_OK; ++i)
        if (cudaEventCreate(&c->ev[i]) != cudaSuccess) st = AOSLO_ERR_CUDA;
#undef A
    if (st == AOSLO_OK && cudaMemset(c->d_status, 0, sizeof(int)) != cudaSuccess) st = AOSLO_ERR_CUDA;
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
    cudaFree(c->d_knn_idx); cudaFree(c->d_knn_d2); cudaFree(c->d_block_sums);
    cudaFree(c->d_touch); cudaFree(c->d_del); cudaFree(c->d_rowstate); cudaFree(c->d_force);
    cudaFree(c->d_pos[0]); cudaFree(c->d_pos[1]); cudaFree(c->d_stable[0]); cudaFree(c->d_stable[1]);
    cudaFree(c->d_dist_p2g); cudaFree(c->d_dist_g2p); cudaFree(c->d_partials); cudaFree(c->d_record);
    cudaFreeHost(c->h_record); cudaFree(c->d_lut); cudaFree(c->d_field); cudaFree(c->d_win);
    KdTreeDev& kd = c->kd;
    cudaFree(kd.idx_array); cudaFree(kd.node_start); cudaFree(kd.node_end); cudaFree(kd.node_bounds);
    cudaFreeHost(kd.h_points); cudaFreeHost(kd.h_idx_array); cudaFreeHost(kd.h_node_start);
    cudaFreeHost(kd.h_node_end); cudaFreeHost(kd.h_node_bounds);
    for (int i = 0; i < 12; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    cudaGetLastError();
    delete c;
    return AOSLO_OK;
}

extern "C" int aoslo_ctx_capacity(aoslo_ctx* ctx) { return ctx ? ctx->cap : 0; }

extern "C" int aoslo_ctx_status(aoslo_ctx* ctx, void* stream, int* h_status_out, int clear) {
    if (!ctx || !h_status_out) return AOSLO_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    AOSLO_TRY(read_int(ctx, s, ctx->d_status, h_status_out));
    if (clear) AOSLO_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(int), s));
    return AOSLO_OK;
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
    AOSLO_TRY(stage_knn(ctx, s, d_points, n_points, d_query, n_query, k, tie_mode, d_idx_out, d_dist_out, true));
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
                           d_particles_out, d_stable_out, d_deleted_out, ctx->d_counters, ctx->d_counters + 1));
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
    return stage_gravity(ctx, s, d_lut, rfs, H, W, d_field);
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
                       tie_mode, d_force);
}

extern "C" int aoslo_move(aoslo_ctx* ctx, void* stream, double* d_particles, const uint8_t* d_stable, int n,
                          const void* d_grad, int field_dtype, int H, int W, const double* d_force,
                          double lambda_blob, double lambda_dist, int step_mode) {
    if (!ctx || !d_particles || !d_stable || !d_grad || !d_force || n < 0) return AOSLO_ERR_INVALID;
    if (step_mode != AOSLO_STEP_DISCRETE && step_mode != AOSLO_STEP_CONTIN) return AOSLO_ERR_UNSUPPORTED;
    return stage_move(ctx, (cudaStream_t)stream, d_particles, d_stable, n, d_grad, field_dtype, H, W, d_force,
                      lambda_blob, lambda_dist, step_mode);
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
                            by, 0, ctx->d_record, p2g, g2p));
    AOSLO_CUDA(cudaMemcpyAsync(ctx->h_record, ctx->d_record, sizeof(aoslo_metrics_record), cudaMemcpyDeviceToHost, s));
    AOSLO_CUDA(cudaStreamSynchronize(s));
    *h_record_out = *ctx->h_record;
    return AOSLO_OK;
}

// ---------------------------------------------------------------------------------------
// whole run
// ---------------------------------------------------------------------------------------
namespace {
struct StageTimer {
    aoslo_ctx* ctx;
    cudaStream_t s;
    float acc[5] = {0, 0, 0, 0, 0};
    std::chrono::steady_clock::time_point t0;
    void begin() { t0 = std::chrono::steady_clock::now(); }
    // v1 run loop synchronises at the end of every stage (count read-back), so host time == device time
    void end(int bucket) {
        acc[bucket] += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};
}  // namespace

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
    cudaStream_t s = (cudaStream_t)stream;
    const int H = cfg->H, W = cfg->W, cap = ctx->cap;
    memset(h_summary, 0, sizeof(*h_summary));
    AOSLO_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(int), s));
    StageTimer tm{ctx, s};

    int cur = 0, n = 0;
    int* d_n = ctx->d_counters;          // scratch scalar
    // ---- seeds (:101) and flags (:113)
    AOSLO_TRY(stage_seed(ctx, s, d_Rb, field_dtype, H, W, cfg->bx, cfg->by, cfg->init_blob_threshold, ctx->d_pos[cur],
                         cap, d_n));
    AOSLO_TRY(read_int(ctx, s, d_n, &n));
    if (n > cap) return AOSLO_ERR_CAPACITY;
    h_summary->n_seeds = n;
    if (n < 2) { h_summary->status = AOSLO_DEV_KNN_SHORT; return AOSLO_OK; }
    AOSLO_TRY(stage_set_stable(ctx, s, ctx->d_stable[cur], n, 0));
    // ---- thinning to a fixpoint (:115-120)
    int rounds = 0;
    while (true) {
        int n_new = 0;
        AOSLO_TRY(stage_delete(ctx, s, ctx->d_pos[cur], ctx->d_stable[cur], n, d_Rb, field_dtype, H, W,
                               cfg->thinning_threshold, cfg->tie_mode, ctx->d_pos[cur ^ 1], ctx->d_stable[cur ^ 1],
                               nullptr, d_n, nullptr));
        AOSLO_TRY(read_int(ctx, s, d_n, &n_new));
        cur ^= 1;
        ++rounds;
        bool none = (n_new == n);
        n = n_new;
        if (none) break;
        if (n < 2) break;
    }
    h_summary->n_after_thinning = n;
    h_summary->n_thinning_rounds = rounds;
    if (n < 2) { h_summary->status = AOSLO_DEV_KNN_SHORT; return AOSLO_OK; }
    // ---- all stable, gravity field, adding (:130-152)
    AOSLO_TRY(stage_lut(ctx, s, cfg->reception_field_size, cfg->mu_dist_func, cfg->sigma_dist_func,
                        cfg->lambda_dist_func, ctx->d_lut));
    auto resample = [&]() -> int {
        AOSLO_TRY(stage_set_stable(ctx, s, ctx->d_stable[cur], n, 1));
        AOSLO_TRY(build_cell_list(ctx, s, ctx->cl_p, ctx->d_pos[cur], n, nullptr));
        AOSLO_TRY(stage_gravity(ctx, s, ctx->d_lut, cfg->reception_field_size, H, W, ctx->d_field));
        AOSLO_TRY(stage_add(ctx, s, ctx->d_field, H, W, cfg->bx, cfg->by, cfg->particle_call_window,
                            cfg->particle_call_threshold, ctx->d_pos[cur], ctx->d_stable[cur], n, nullptr, cap, d_n));
        AOSLO_TRY(read_int(ctx, s, d_n, &n));
        if (n > cap) return AOSLO_ERR_CAPACITY;
        return AOSLO_OK;
    };
    AOSLO_TRY(resample());
    // GT cell list once
    AOSLO_TRY(build_cell_list(ctx, s, ctx->cl_g, d_gt, n_gt, nullptr));

    const int k = cfg->dist_n_neighbours;
    const bool pit = cfg->dist_energy_type == AOSLO_ENERGY_PIT;
    int n_rec = 0, it_done = 0;
    for (int it = 0; it < cfg->epoch_num; ++it) {
        ++it_done;
        if (n < k + 1) { h_summary->status |= AOSLO_DEV_KNN_SHORT; break; }
        tm.begin();
        AOSLO_TRY(stage_force(ctx, s, ctx->d_pos[cur], ctx->d_stable[cur], n, k, cfg->dist_energy_type,
                              pit ? cfg->mu_dist_func : cfg->dist_alpha, pit ? cfg->sigma_dist_func : cfg->dist_sigma,
                              pit ? cfg->lambda_dist_func : 0.0, cfg->tie_mode, ctx->d_force));
        AOSLO_CUDA(cudaStreamSynchronize(s));
        tm.end(0);
        tm.begin();
        AOSLO_TRY(stage_move(ctx, s, ctx->d_pos[cur], ctx->d_stable[cur], n, d_grad, field_dtype, H, W, ctx->d_force,
                             cfg->lambda_blob, cfg->lambda_dist, cfg->step_mode));
        AOSLO_CUDA(cudaStreamSynchronize(s));
        tm.end(1);
        tm.begin();
        AOSLO_TRY(stage_delete(ctx, s, ctx->d_pos[cur], ctx->d_stable[cur], n, d_Rb, field_dtype, H, W,
                               cfg->threshhold_dist_del, cfg->tie_mode, ctx->d_pos[cur ^ 1], ctx->d_stable[cur ^ 1],
                               nullptr, d_n, nullptr));
        AOSLO_TRY(read_int(ctx, s, d_n, &n));
        cur ^= 1;
        tm.end(2);
        tm.begin();
        if (it % cfg->fix_particles_frequency == 3) AOSLO_TRY(resample());
        AOSLO_CUDA(cudaStreamSynchronize(s));
        tm.end(3);
        tm.begin();
        if (it % cfg->metric_measure_freq == 0 && h_records && n_rec < max_records) {
            AOSLO_TRY(stage_metrics(ctx, s, ctx->d_pos[cur], ctx->d_stable[cur], n, d_gt, n_gt, true, false, d_Rb,
                                    field_dtype, H, W, cfg->bx, cfg->by, it, ctx->d_record, ctx->d_dist_p2g,
                                    ctx->d_dist_g2p));
            AOSLO_CUDA(cudaMemcpyAsync(ctx->h_record, ctx->d_record, sizeof(aoslo_metrics_record),
                                       cudaMemcpyDeviceToHost, s));
            AOSLO_CUDA(cudaStreamSynchronize(s));
            h_records[n_rec++] = *ctx->h_record;
        }
        int n_stable = 0;
        AOSLO_TRY(stage_count_stable(ctx, s, ctx->d_stable[cur], n, ctx->d_counters + 2));
        AOSLO_TRY(read_int(ctx, s, ctx->d_counters + 2, &n_stable));
        h_summary->n_stable = n_stable;
        // sum(stable) / len(particles) > 0.997 (:381) -- Python true division in float64
        if (cfg->early_stop && ((double)n_stable / (double)n > 0.997)) break;
        tm.end(4);
    }
    if (d_particles_out)
        AOSLO_CUDA(cudaMemcpyAsync(d_particles_out, ctx->d_pos[cur], sizeof(double) * 2 * n, cudaMemcpyDeviceToDevice, s));
    if (d_stable_out)
        AOSLO_CUDA(cudaMemcpyAsync(d_stable_out, ctx->d_stable[cur], n, cudaMemcpyDeviceToDevice, s));
    int status = 0;
    AOSLO_TRY(read_int(ctx, s, ctx->d_status, &status));
    h_summary->status |= status;
    h_summary->n_iterations = it_done;
    h_summary->n_particles = n;
    h_summary->n_records = n_rec;
    if (h_stage_ms) for (int i = 0; i < 5; ++i) h_stage_ms[i] = tm.acc[i];
    return AOSLO_OK;
}
