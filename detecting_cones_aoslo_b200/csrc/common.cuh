// common.cuh -- shared declarations of libaoslo_b200 (sm_100a).
//
// Data layout in HBM:
//   particles   double2[n]  rows [x, y]  (the reference's (n,2) float64 ndarray)
//   stable      uint8[n]
//   fields      T[H*W] row-major, gradient T[H*W*2] interleaved (T = float | double)
//   cell list   count[ncell], start[ncell+1], sorted_pos double2[n], sorted_idx int[n]
//   kNN table   idx int32[n*k], d2 double[n*k]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include <vector>

#include "aoslo_b200.h"

namespace aoslo {

int set_cuda_error(cudaError_t e, const char* where);
void count_launch();
void count_launches(int n);

#define AOSLO_CUDA(call)                                                         \
    do {                                                                         \
        cudaError_t e__ = (call);                                                \
        if (e__ != cudaSuccess) return ::aoslo::set_cuda_error(e__, #call);      \
    } while (0)
#define AOSLO_TRY(call)                                                          \
    do {                                                                         \
        int s__ = (call);                                                        \
        if (s__ != AOSLO_OK) return s__;                                         \
    } while (0)
// every kernel launch is followed by this: counts the launch (aoslo_launch_count) and checks it
#define AOSLO_LAUNCH_CHECK()                                                     \
    do {                                                                         \
        ::aoslo::count_launch();                                                 \
        AOSLO_CUDA(cudaGetLastError());                                          \
    } while (0)

// Row pitch (in pixels) of every blobness / gradient FIELD buffer: rows are padded to a multiple of 4 pixels so
// that a row starts 16-byte aligned in f32 (32-byte in f64) whatever the frame width -- the alignment TMA tensor
// stores need.  R_b is (H, fpitch(W)), the gradient (H, fpitch(W), 2); only the first W pixels of a row are data.
__host__ __device__ inline int fpitch(int W) { return (W + 3) & ~3; }

constexpr int kThreads = 256;          // block size of the grid-stride kernels
constexpr int kScanThreads = 1024;     // block size of the ordered-compaction kernels
constexpr int kMaxScanBlocks = 1024;   // spine length of the 3-phase scan

// Uniform-grid cell list over [0,W) x [0,H); cell edge = 1 << shift pixels.
struct CellList {
    int shift = 2;
    int gw = 0, gh = 0, ncell = 0;
    int* count = nullptr;        // [ncell]
    int* start = nullptr;        // [ncell + 1]
    int* slot = nullptr;         // [cap]   rank of the particle inside its cell
    int* cell_of = nullptr;      // [cap]
    int* sorted_idx = nullptr;   // [cap]
    double2* sorted_pos = nullptr;  // [cap]
    int cap = 0;
};

struct CellView {                // what kernels take by value
    int shift, gw, gh;
    const int* start;
    const int* sorted_idx;
    const double2* sorted_pos;
};

inline CellView view_of(const CellList& c) {
    return CellView{c.shift, c.gw, c.gh, c.start, c.sorted_idx, c.sorted_pos};
}

// Scan-free cell structure for kNN: per-cell singly linked lists (head[cell], next[particle]).
// Built by one memset + one kernel (atomicExch); chain order is arbitrary, which is fine because
// neighbour selection uses the strict total order (d^2, index).
struct CellLinks {
    int fixed_shift = 0;         // > 0: pinned cell edge (log2); 0: chosen per build from the density
    int min_shift = 1;           // head[] is allocated for this finest grid
    int shift = 3;
    int gw = 0, gh = 0, ncell = 0;
    unsigned long long* head = nullptr;   // [ncell] (tag << 32) | first index; stale tags = empty
    int* next = nullptr;         // [cap]
    int cap = 0;
};

// geom = {shift, gw, gh, ncell, tag} lives in DEVICE memory: the cell edge is chosen from the point count, which the
// run loop only knows on the device (the last block of link_kernel publishes it, every later kernel reads it);
// head words are (tag << 32) | index, valid only with the tag of the current build (no clearing pass)
struct LinkView {
    const int* geom;
    const unsigned long long* head;
    const int* next;
    const double2* pos;
};
constexpr int kLinkBiasDense = 6;     // link_shift_for bias: thinning rounds (dense seed sets)
constexpr int kLinkBiasSteady = 3;    // steady state of the loop / metrics

// ---- whole-run control block (device resident, mirrored to pinned memory once at the end of a run) ----
// aoslo_run enqueues seeding + thinning (one cooperative kernel on the alive-pixel bitmap), resampling and the epoch loop
// (unrolled, or a WHILE node with IF nodes for the resample / metrics branches when the early stop is on)
// as ONE CUDA graph; every decision the reference takes on the host is taken by a 1-thread kernel here.
struct RunCtl {
    int it;                  // iterations executed so far
    int n_rec;               // metric records written
    int n_seeds, n_after_thinning, thin_rounds;
    int n_stable;            // stable flags counted last (early stop / summary)
    int thin_cond;           // mirrors of the conditional-node values (the eager debug path reads them)
    int loop_cond;
    int do_resample, do_metrics;
    int n_final;
    int pad_;
    unsigned long long t_last;       // %globaltimer of the last stage boundary
    unsigned long long acc_ns[5];    // the reference's five time_spended buckets
};

// value parameters of a run, read by the kernels from device memory so that one cached graph serves every
// configuration of a sweep (only structural parameters -- shapes, template classes, loop layout -- are baked)
struct RunParams {
    double seed_thr, thin_thr, del_thr, add_thr;
    double lut_mu, lut_sigma, lut_lam;
    double e0, e1, e2;       // energy parameters (pit: mu, sigma, lambda; exp: alpha, sigma, 0)
    double lb, ld;           // lambda_blob, lambda_dist
    int kk;                  // dist_n_neighbours + 1
    int fix_freq, metric_freq, epoch_num, early_stop, max_records, n_gt, pad_;
};

// K1 implementation knobs: defaults from the environment ONCE at context creation (AOSLO_K1_TMA, AOSLO_PAIR_TH,
// AOSLO_PAIR_OCC, AOSLO_DTILE_TH, AOSLO_BLOBNESS_KERNEL = tile | band | dtile), changed with aoslo_ctx_set_tuning;
// nothing on the launch path calls getenv
struct K1Tuning {
    int k1_tma = 1;          // 0: cp.async staging, plain stores; 1: TMA tensor loads; 2: TMA tensor stores too
    int pair_th = 0;         // tile height of the f32 pair kernel (0 = chosen per image)
    int pair_occ = 0;        // CTAs per SM of the f32 pair kernel (0 = by tile height)
    int dtile_th = 0;        // tile height of the f64 tile kernel (0 = 32, 16 for frames of less than half a wave)
    int dtile_occ = 0;       // CTAs per SM of the f64 tile kernel (0 = by tile height)
    int dtile_split = 1;     // 16-row remainder tiles instead of a nearly empty last wave of full tiles
    int k1_kernel = 0;       // 0 auto, 1 generic tile kernel, 2 band kernel, 3 f64 tile kernel
};

struct RunGraph {            // one cached instantiation
    std::vector<unsigned char> key;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int static_kernels = 0, thin_kernels = 0, iter_kernels = 0, resample_kernels = 0, metrics_kernels = 0;
    unsigned long long stamp = 0;    // LRU
};

// sklearn-order KD tree mirrored on the device (built on the host, kdtree_host.cpp)
struct KdTreeDev {
    int n = 0, n_nodes = 0;
    int32_t* idx_array = nullptr;     // [cap]
    int32_t* node_start = nullptr;    // [2*cap]
    int32_t* node_end = nullptr;      // [2*cap]
    double* node_bounds = nullptr;    // [2*cap*4] lo_x lo_y hi_x hi_y
    // pinned host staging
    double* h_points = nullptr;       // [cap*2]
    int32_t* h_idx_array = nullptr;
    int32_t* h_node_start = nullptr;
    int32_t* h_node_end = nullptr;
    double* h_node_bounds = nullptr;
    int cap = 0;
    int cached_n = -1;                // the device arrays hold the tree of the cached_n points in h_points
};

}  // namespace aoslo

struct aoslo_ctx {
    int device = 0;
    int H = 0, W = 0;
    int cap = 0, gt_cap = 0;
    int num_sms = 0;
    int nb = 0;                       // blocks of the grid-stride kernels (multiple of the SM count)
    int coop_blocks = 0;              // co-resident blocks of the cooperative deletion kernel
    int* d_status = nullptr;
    int* d_counters = nullptr;        // [32] device scalars (counts, round counters)
    int* h_pinned = nullptr;          // [32] pinned readback
    aoslo::CellList cl_g;             // ground truth
    aoslo::CellLinks lk_p;            // particles (kNN)
    double* d_pos_tmp = nullptr;      // [cap*2] moved positions (Jacobi move writes here)
    int* d_active = nullptr;          // [cap] compact list of active deletion rows (non-stable path)
    int* d_ns_list = nullptr;         // [cap] compact list of non-stable particles
    int blocking_sync = 0;            // 1: host waits sleep on a blocking event instead of spinning
    cudaEvent_t ev_block = nullptr;
    // host work deferred behind the next run (aoslo_stage_gt: the reference's in-place GT shift)
    cudaEvent_t ev_upload = nullptr;
    double* defer_ptr = nullptr;
    long long defer_n = 0;
    double defer_add = 0.0;
    unsigned int epoch = 1;           // host-side upper estimate of *d_epoch
    unsigned int* d_epoch = nullptr;  // device call counter of the in-loop deletion (stamps `touch`)
    // cached CUDA graphs of the whole run (canonical tie order), keyed by the structural parameters
    std::vector<aoslo::RunGraph> graphs;
    unsigned long long graph_clock = 0;
    cudaStream_t cap_stream = nullptr;   // recording streams of the capture (top level, WHILE body, IF body)
    cudaStream_t cap_stream2 = nullptr;
    cudaStream_t cap_stream3 = nullptr;
    aoslo::RunCtl* d_ctl = nullptr;
    aoslo::RunCtl* h_ctl = nullptr;      // pinned
    aoslo::RunParams* d_params = nullptr;
    aoslo::RunParams* h_params = nullptr;   // pinned staging of the per-run upload
    int* d_link_geom = nullptr;          // [8] shift, gw, gh, ncell, tag of lk_p
    unsigned int* d_link_epoch = nullptr;   // build counter = tag of the last build
    unsigned int* d_link_epoch_n = nullptr; // same for the non-stable structure N of the run loop
    unsigned long long* d_head_n = nullptr; // [ncell of the finest grid] heads of N
    int* d_link_epoch_ticket = nullptr;
    int loop_hint = 0;                   // launch-size hint of the in-loop compaction (performance only)
    aoslo::K1Tuning tune;
    int last_n = 0, last_buf = 0;        // where the final particles of the last run live (aoslo_read_particles)
    int no_graph = 0;                    // AOSLO_NO_GRAPH read once at context creation (eager debug path)
    int coop_cap_blocks = 0;
    int32_t* d_knn_idx = nullptr;     // [max(cap,gt_cap) * KMAX]
    double* d_knn_d2 = nullptr;
    int* d_block_sums = nullptr;      // [kMaxScanBlocks + 1]
    int* d_scan_ticket = nullptr;     // last-block ticket of scan_count_kernel
    unsigned long long* d_touch = nullptr;  // [cap]
    uint8_t* d_del = nullptr;         // [cap]
    uint8_t* d_rowstate = nullptr;    // [cap]
    double* d_force = nullptr;        // [cap*2]
    double* d_pos[2] = {nullptr, nullptr};
    uint8_t* d_stable[2] = {nullptr, nullptr};
    double* d_dist_p2g = nullptr;     // [cap]
    double* d_dist_g2p = nullptr;     // [gt_cap]
    double* d_partials = nullptr;     // per-block metrics partials (nb * 9 doubles)
    aoslo_metrics_record* d_record = nullptr;
    aoslo_metrics_record* h_record = nullptr;   // pinned
    aoslo_metrics_record* d_records = nullptr;  // [max_records] one per metrics evaluation of a run
    aoslo_metrics_record* h_records = nullptr;  // pinned
    int max_records = 0;
    double* d_lut = nullptr;          // [64*64]
    double* d_lut_radial = nullptr;   // [2*15*15+1] LUT entry by squared distance (lut_radial_kernel)
    int* d_lut_flag = nullptr;        // 1: the LUT is radial and non-increasing (nearest-particle gravity kernel)
    uint32_t* d_occ = nullptr;        // [H][(W+31)/32 + 2] occupancy bitmap of the particles' pixels
    uint32_t* d_occ_gt = nullptr;     // same for the ground truth of the current run (stage_gt_bitmap)
    int* d_occ_ok = nullptr;          // [2] particles / GT: every point is an in-image integer pixel
    uint32_t* d_thin = nullptr;       // [7][H * wpr] pixel-domain thinning: six stencil bit planes + pending rows
    int* d_thin_ctr = nullptr;        // [8] its counters
    unsigned long long* d_thin_touch = nullptr;   // [H * W] second wavefront stamp buffer (the first lives in d_field)
    int coop_blocks_thin = 0;         // co-resident blocks of thin_pixels_kernel
    int thin_threads = 128;           // its block size (AOSLO_THIN_THREADS, read once at context creation)
    double* d_field = nullptr;        // [H*W]
    int32_t* d_win = nullptr;         // [H*W] window candidates of adding_particles
    aoslo::KdTreeDev kd, kd2;         // two cached trees (sklearn-order tie mode): typically particles and GT
    double* kd_stage = nullptr;       // pinned [cap*2]: the points of the current call, compared with the caches
    int kd_lru = 0;                   // slot to overwrite on a miss
    cudaEvent_t ev[12] = {};
    std::vector<cudaEvent_t> ev_pool;   // stage-timing events of aoslo_run
};

namespace aoslo {

// ---------------------------------------------------------------------------------------
// Ordered compaction / exclusive scan in two launches (count + spine by the last block, emit).  `Op` provides
//   __device__ int  value(int i) const      -- the integer to be scanned (0/1 for a compaction)
//   __device__ void emit(int i, int v, int exclusive_prefix) const
// The element count may live on the device (d_m != nullptr) so that callers never
// have to synchronise to learn it.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += t;
    }
    return v;
}

// block-wide exclusive scan of one int per thread; returns exclusive prefix, total in `total`
__device__ __forceinline__ int block_excl_scan(int v, int* smem_warp /*[32]*/, int& total) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = warp_incl_scan(v);
    if (lane == 31) smem_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        int w = (lane < nw) ? smem_warp[lane] : 0;
        int wi = warp_incl_scan(w);
        smem_warp[lane] = wi - w;          // exclusive warp offsets
        if (lane == 31) smem_warp[32] = wi;  // total
    }
    __syncthreads();
    int excl = incl - v + smem_warp[wid];
    total = smem_warp[32];
    __syncthreads();
    return excl;
}

__device__ __forceinline__ void scan_span(int m, int& lo, int& hi) {
    int per = (m + gridDim.x - 1) / gridDim.x;
    per = ((per + blockDim.x - 1) / blockDim.x) * blockDim.x;
    long long l = (long long)blockIdx.x * per;
    lo = (int)(l < m ? l : m);
    long long h = l + per;
    hi = (int)(h < m ? h : m);
}

// Phase 1 + 2: every block sums its span; the LAST block to finish (ticket) turns the block sums into exclusive
// offsets (+ base) and stores the total -- the former one-block spine kernel, without its launch.
// `clamp` > 0: the stored total never exceeds it (an over-full append raises AOSLO_DEV_CAPACITY in `status` and later
// kernels iterate over at most `clamp` rows -- the buffers' capacity).
template <class Op>
__global__ void __launch_bounds__(kScanThreads) scan_count_kernel(Op op, int m_host, const int* d_m, int* block_sums,
                                                                 int* ticket, int* d_total, const int* d_base,
                                                                 int base_host, int* d_total_added, int clamp,
                                                                 int* status) {
    __shared__ int sw[33];
    __shared__ int s_last;
    int m = d_m ? *d_m : m_host;
    const int base = d_base ? *d_base : base_host;       // read before the last block may overwrite it (d_total)
    int lo, hi;
    scan_span(m, lo, hi);
    int acc = 0;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += op.value(i);
    int total;
    block_excl_scan(acc, sw, total);
    if (threadIdx.x == 0) {
        __stcg(&block_sums[blockIdx.x], total);
        __threadfence();
        s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int nblocks = (int)gridDim.x;                  // <= kMaxScanBlocks = blockDim.x
    int v = (threadIdx.x < nblocks) ? __ldcg(&block_sums[threadIdx.x]) : 0;
    int grand;
    int excl = block_excl_scan(v, sw, grand);
    if (threadIdx.x < nblocks) block_sums[threadIdx.x] = base + excl;
    if (threadIdx.x == 0) {
        int t = base + grand;
        if (clamp > 0 && t > clamp) {                    // the emit phase skips the rows beyond the capacity
            t = clamp;
            if (status) atomicOr(status, AOSLO_DEV_CAPACITY);
        }
        if (d_total) *d_total = t;
        if (d_total_added) *d_total_added = grand;
        *ticket = 0;
    }
}

template <class Op>
__global__ void __launch_bounds__(kScanThreads) scan_emit_kernel(Op op, int m_host, const int* d_m,
                                                                const int* block_sums) {
    __shared__ int sw[33];
    int m = d_m ? *d_m : m_host;
    int lo, hi;
    scan_span(m, lo, hi);
    int carry = block_sums[blockIdx.x];
    for (int base = lo; base < hi; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = (i < hi) ? op.value(i) : 0;
        int total;
        int excl = block_excl_scan(v, sw, total);
        if (i < hi) op.emit(i, v, carry + excl);
        carry += total;
    }
}

int scan_blocks_for(int m_capacity);

// host driver: two launches on `stream` (count + spine, emit).
//   d_total      <- base + sum(value)      (may be nullptr)
//   d_total_added<- sum(value)             (may be nullptr)
// `base` (offset added to every prefix handed to emit) comes from *d_base if non-null.
template <class Op>
int ordered_scan(aoslo_ctx* ctx, cudaStream_t s, Op op, int m_capacity, int m_host, const int* d_m,
                 int* d_total, const int* d_base, int base_host, int* d_total_added, int clamp = 0) {
    int nb = scan_blocks_for(m_capacity);
    scan_count_kernel<Op><<<nb, kScanThreads, 0, s>>>(op, m_host, d_m, ctx->d_block_sums, ctx->d_scan_ticket, d_total,
                                                      d_base, base_host, d_total_added, clamp, ctx->d_status);
    AOSLO_LAUNCH_CHECK();
    scan_emit_kernel<Op><<<nb, kScanThreads, 0, s>>>(op, m_host, d_m, ctx->d_block_sums);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

// field element access (T = float | double), always widened to double for decisions
template <typename T>
__device__ __forceinline__ double fld(const T* p, long long i) {
    return (double)p[i];
}

// ---------------------------------------------------------------------------------------
// stage entry points implemented in particles.cu / blobness.cu (host side, stream-ordered)
// (kd_reserve: api.cu)
// ---------------------------------------------------------------------------------------
int kd_reserve(aoslo_ctx* ctx);
int epoch_guard(aoslo_ctx* ctx, cudaStream_t s, unsigned int calls_ahead);
int stream_wait(aoslo_ctx* ctx, cudaStream_t s);   // host waits for the stream (spin or blocking event)
int build_links(aoslo_ctx* ctx, cudaStream_t s, CellLinks& lk, const double* d_pos, int n_host, const int* d_n,
                int bias_x4);
int build_cell_list(aoslo_ctx* ctx, cudaStream_t s, CellList& cl, const double* d_pos, int n_host, const int* d_n);
int knn_sklearn(aoslo_ctx* ctx, cudaStream_t s, const double* d_points, int n_points, const double* d_query,
                int n_query, int k, int32_t* d_idx, double* d_d2);

// stage drivers (particles.cu).  Trailing d_n: optional device-resident particle count (n is then only a
// launch-size bound); d_thr / rp: device-resident value parameters of the graph path (override the host values)
int stage_knn(aoslo_ctx*, cudaStream_t, const double* d_points, int n_points, const double* d_query, int n_query,
              int k, int tie_mode, int32_t* d_idx, double* d_d2, bool tree_is_particles, const int* d_n = nullptr);
int stage_delete(aoslo_ctx*, cudaStream_t, const double* d_pos, const uint8_t* d_stable, int n, const void* d_Rb,
                 int field_dtype, int H, int W, double thr, int tie_mode, double* d_pos_out, uint8_t* d_stable_out,
                 uint8_t* d_deleted_out, int* d_n_out, int* d_rounds_out, const int* d_n, int mode,
                 const double* d_thr = nullptr, int scan_hint = 0);
int stage_seed(aoslo_ctx*, cudaStream_t, const void* d_Rb, int field_dtype, int H, int W, int bx, int by, double thr,
               double* d_pos, int cap, int* d_n_out, const double* d_thr = nullptr, bool clamp_count = false);
int stage_lut(aoslo_ctx*, cudaStream_t, int rfs, double mu, double sigma, double lam, double* d_lut,
              const RunParams* rp = nullptr);
int gravity_init();
int stage_gravity(aoslo_ctx*, cudaStream_t, const double* d_pos, int n_host, const int* d_n, const double* d_lut,
                  int rfs, int H, int W, double* d_field, bool lut_ready);
int stage_add(aoslo_ctx*, cudaStream_t, const double* d_field, int H, int W, int bx, int by, int window, double thr,
              double* d_pos, uint8_t* d_stable, int n_host, const int* d_n, int cap, int* d_n_out,
              const double* d_thr = nullptr, bool clamp_count = false);
int stage_force(aoslo_ctx*, cudaStream_t, const double* d_pos, const uint8_t* d_stable, int n, int k, int energy_type,
                double p0, double p1, double p2, int tie_mode, double* d_force, const int* d_n);
int stage_move(aoslo_ctx*, cudaStream_t, double* d_pos, const uint8_t* d_stable, int n, const void* d_grad,
               int field_dtype, int H, int W, const double* d_force, double lb, double ld, int step_mode,
               const int* d_n);
int stage_metrics(aoslo_ctx*, cudaStream_t, const double* d_pos, const uint8_t* d_stable, int n, const double* d_gt,
                  int n_gt, bool gt_list_ready, bool p_list_ready, const void* d_Rb, int field_dtype, int H, int W,
                  int bx, int by, int iteration, aoslo_metrics_record* d_rec, double* d_p2g, double* d_g2p,
                  const int* d_n, const int* d_n_gt = nullptr, int* d_rec_idx = nullptr,
                  const int* d_iteration = nullptr);
// run loop, tombstone / dual-structure scheme (particles.cu): state lives in buffer 0, tombstone count in
// d_counters[7]
int stage_period_begin(aoslo_ctx*, cudaStream_t, int n_bound, const int* d_n);
int stage_dual_step(aoslo_ctx*, cudaStream_t, int n_bound, const int* d_n, int k, int energy_type, double p0, double p1,
                    double p2, const void* d_grad, const void* d_Rb, int field_dtype, int H, int W, double lb, double ld,
                    int step_mode, double del_thr, const RunParams* rp, int phases);
int stage_compact(aoslo_ctx*, cudaStream_t, int n_bound, int* d_n, int* d_n_scratch, int scan_hint);
int stage_gt_bitmap(aoslo_ctx*, cudaStream_t, const double* d_gt, int n_gt, const int* d_n_gt);
int stage_dual_metrics(aoslo_ctx*, cudaStream_t, int n_bound, const int* d_n, const double* d_gt, int n_gt,
                       const void* d_Rb, int field_dtype, int H, int W, int bx, int by, int iteration,
                       aoslo_metrics_record* d_rec, const int* d_n_gt, int* d_rec_idx, const int* d_iteration);
int stage_dual_count_stable(aoslo_ctx*, cudaStream_t, int n_bound, const int* d_n, int* d_out);
int stage_set_stable(aoslo_ctx*, cudaStream_t, uint8_t* d_stable, int n, uint8_t v, const int* d_n);
int stage_count_stable(aoslo_ctx*, cudaStream_t, const uint8_t* d_stable, int n, int* d_out, const int* d_n);
int coop_blocks_for_delete(int num_sms);
int thin_threads_default(int H, int W, int num_sms);
int coop_blocks_for_thin(int num_sms, int threads);
int stage_thin_pixels(aoslo_ctx*, cudaStream_t, const void* d_Rb, int field_dtype, int H, int W, int bx, int by,
                      double seed_thr, double thin_thr, const RunParams* rp, RunCtl* ctl, double* d_pos,
                      uint8_t* d_stable, int* d_n_out);
int stage_classify(aoslo_ctx*, cudaStream_t, const double* d_g2p, int n_gt, const double* d_p2g, const double* d_pos,
                   int n, int bx, int by, int W, int H, double thr, uint8_t* gt_is_tp, uint8_t* particle_is_fp);
int stage_prepare(aoslo_ctx*, cudaStream_t, const uint8_t* d_raw, int H0, int W0, int pitch0, int x_min, int x_max,
                  int y_min, int y_max, int bx, int by, uint8_t* d_img_out, int out_pitch, const double* d_gt_raw,
                  int n_gt_raw, double* d_gt_out, int* d_n_gt_out, int gt_is_raw);
int sqrt_inplace(cudaStream_t, double* a, size_t n);

}  // namespace aoslo
