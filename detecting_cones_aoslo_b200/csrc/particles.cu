// particles.cu -- the particle system of the detection path on sm_100a (compiled with
// -fmad=false: every FP64 product/sum is rounded separately, like the reference's NumPy code,
// because the loop's decisions are discrete and amplify rounding differences).
//
// Kernels (reference file:line each one replaces):
//   cell_count / cell_scatter        uniform-grid cell list of the ground truth (replaces sklearn KDTree.fit, utils.py:218)
//   knn_cells_kernel                 exact kNN, canonical (d^2, index) order          (utils.py:30, :219)
//   knn_sklearn_kernel               exact kNN in sklearn's depth-first tie order     (same call sites)
//   seed (ordered compaction)        utils.initialize_high_prop_location             (utils.py:359-375)
//   delete_resolve_kernel            utils.del_close_particles greedy loop           (utils.py:249-263)
//   keep-compaction                  utils.del_close_particles compaction            (utils.py:265-273)
//   dist_lut_kernel                  utils.get_precomputed_dist_func                 (utils.py:95-109)
//   gravity_nearest / gravity_bits   utils.calc_gravity_field_optim                  (utils.py:310-333)
//   window_argmin_kernel + append    utils.adding_particles                          (utils.py:336-356)
//   dist_force_kernel                utils.calc_dist_energy_pit(_optim) / calc_dist_energy (utils.py:112-134, :43-58)
//   move_kernel                      the move step                 (pipeline_with_delitions.py:215-245)
//   metrics kernels                  calc_dist_to_neares x2, get_particles_in_image, calc_acc_metrics,
//                                    calc_metrics('Chamfer'), calc_blob_energy (utils.py:214-242, :378-418, :137-144)
// The whole-run loop of aoslo_run (canonical tie order) works on a tombstone + two-structure layout of the same
// arithmetic (period_begin / dual_force_move / ns_relink / dual_rows / dual_metrics, "The run loop's particle
// structure" below); its seeds + thinning rounds run in the pixel domain (thin_pixels_kernel), its gravity field and
// metrics distances come from occupancy bitmaps (gravity_nearest_kernel, occ_nearest_d2).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace aoslo {

// =======================================================================================
// ordered scan helpers
// =======================================================================================
int scan_blocks_for(int m_capacity) {
    int nb = (m_capacity + kScanThreads - 1) / kScanThreads;
    if (nb < 1) nb = 1;
    if (nb > kMaxScanBlocks) nb = kMaxScanBlocks;
    return nb;
}

// =======================================================================================
// cell list
// =======================================================================================
__device__ __forceinline__ int cell_coord(double v, int shift, int gmax, bool& bad) {
    // positions must lie in [0, extent): anything else is reported (the reference would raise
    // IndexError or silently wrap a negative index) and binned into the border cell
    if (!(v >= 0.0)) { bad = true; return 0; }
    int c = (int)v >> shift;       // v >= 0: trunc == floor
    if (v >= 2147483000.0) { bad = true; return gmax - 1; }
    if (c >= gmax) { c = gmax - 1; }
    return c;
}

__global__ void __launch_bounds__(kThreads) cell_count_kernel(const double2* __restrict__ pos, int n_host,
                                                              const int* d_n, int shift, int gw, int gh, int W,
                                                              int H, int* count, int* cell_of, int* slot,
                                                              int* status) {
    int n = d_n ? *d_n : n_host;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double2 p = pos[i];
        bool bad = false;
        int cx = cell_coord(p.x, shift, gw, bad);
        int cy = cell_coord(p.y, shift, gh, bad);
        if (p.x != p.x || p.y != p.y) atomicOr(status, AOSLO_DEV_NAN);
        else if (bad || p.x >= (double)W || p.y >= (double)H) atomicOr(status, AOSLO_DEV_ESCAPED);
        int c = cy * gw + cx;
        cell_of[i] = c;
        slot[i] = atomicAdd(&count[c], 1);
    }
}

struct CellScanOp {
    int* count;
    int* start;
    __device__ int value(int i) const { return count[i]; }
    // the count is consumed here, so it is re-zeroed for the next build (no memset per build)
    __device__ void emit(int i, int, int excl) const { start[i] = excl; count[i] = 0; }
};

__global__ void __launch_bounds__(kThreads) cell_scatter_kernel(const double2* __restrict__ pos, int n_host,
                                                                const int* d_n, const int* __restrict__ cell_of,
                                                                const int* __restrict__ slot,
                                                                const int* __restrict__ start, int* sorted_idx,
                                                                double2* sorted_pos) {
    int n = d_n ? *d_n : n_host;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int dst = start[cell_of[i]] + slot[i];
        sorted_idx[dst] = i;
        sorted_pos[dst] = pos[i];
    }
}

int build_cell_list(aoslo_ctx* ctx, cudaStream_t s, CellList& cl, const double* d_pos, int n_host, const int* d_n) {
    // cl.count is all zero here: zeroed at allocation and re-zeroed by every scan (CellScanOp::emit)
    cell_count_kernel<<<ctx->nb, kThreads, 0, s>>>((const double2*)d_pos, n_host, d_n, cl.shift, cl.gw, cl.gh,
                                                   ctx->W, ctx->H, cl.count, cl.cell_of, cl.slot, ctx->d_status);
    AOSLO_LAUNCH_CHECK();
    CellScanOp op{cl.count, cl.start};
    // start[ncell] = total
    AOSLO_TRY(ordered_scan(ctx, s, op, cl.ncell, cl.ncell, nullptr, cl.start + cl.ncell, nullptr, 0, nullptr));
    cell_scatter_kernel<<<ctx->nb, kThreads, 0, s>>>((const double2*)d_pos, n_host, d_n, cl.cell_of, cl.slot,
                                                     cl.start, cl.sorted_idx, cl.sorted_pos);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

// =======================================================================================
// exact kNN on the cell list, canonical (d^2, index) order
// =======================================================================================
template <int K>
struct TopK {
    double d2[K];
    int idx[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < K; ++j) { d2[j] = INFINITY; idx[j] = 0x7fffffff; }
    }
    __device__ __forceinline__ void push(double d, int id) {
        // strict total order (d2, idx); NaN distances never enter
        if (!(d < d2[K - 1] || (d == d2[K - 1] && id < idx[K - 1]))) return;
        d2[K - 1] = d; idx[K - 1] = id;
#pragma unroll
        for (int j = K - 1; j > 0; --j) {
            bool less = d2[j] < d2[j - 1] || (d2[j] == d2[j - 1] && idx[j] < idx[j - 1]);
            if (less) {
                double td = d2[j]; d2[j] = d2[j - 1]; d2[j - 1] = td;
                int ti = idx[j]; idx[j] = idx[j - 1]; idx[j - 1] = ti;
            }
        }
    }
};

template <int K>
__device__ __forceinline__ void knn_cells_query(const CellView& cl, double qx, double qy, TopK<K>& top) {
    top.init();
    const int c = 1 << cl.shift;
    int cx = (qx >= 0.0) ? ((int)fmin(qx, 1.0e9) >> cl.shift) : 0;
    int cy = (qy >= 0.0) ? ((int)fmin(qy, 1.0e9) >> cl.shift) : 0;
    cx = min(cx, cl.gw - 1);
    cy = min(cy, cl.gh - 1);
    const int rmax = max(cl.gw, cl.gh);
    for (int r = 0; r <= rmax; ++r) {
        int ylo = cy - r, yhi = cy + r, xlo = cx - r, xhi = cx + r;
        for (int y = max(ylo, 0); y <= min(yhi, cl.gh - 1); ++y) {
            bool edge_row = (y == ylo) || (y == yhi);
            if (edge_row) {
                // the whole row segment is one contiguous span of the sorted arrays
                int a = max(xlo, 0), b = min(xhi, cl.gw - 1);
                if (a > b) continue;
                int p0 = cl.start[y * cl.gw + a], p1 = cl.start[y * cl.gw + b + 1];
                for (int p = p0; p < p1; ++p) {
                    double2 t = cl.sorted_pos[p];
                    double dx = qx - t.x, dy = qy - t.y;
                    top.push(dx * dx + dy * dy, cl.sorted_idx[p]);
                }
            } else {
                if (xlo >= 0) {
                    int p0 = cl.start[y * cl.gw + xlo], p1 = cl.start[y * cl.gw + xlo + 1];
                    for (int p = p0; p < p1; ++p) {
                        double2 t = cl.sorted_pos[p];
                        double dx = qx - t.x, dy = qy - t.y;
                        top.push(dx * dx + dy * dy, cl.sorted_idx[p]);
                    }
                }
                if (xhi < cl.gw && r > 0) {
                    int p0 = cl.start[y * cl.gw + xhi], p1 = cl.start[y * cl.gw + xhi + 1];
                    for (int p = p0; p < p1; ++p) {
                        double2 t = cl.sorted_pos[p];
                        double dx = qx - t.x, dy = qy - t.y;
                        top.push(dx * dx + dy * dy, cl.sorted_idx[p]);
                    }
                }
            }
        }
        // every unseen point lies outside the (2r+1)^2 block: lower-bound its distance
        double bound = INFINITY;
        if (xlo > 0) bound = fmin(bound, qx - (double)(xlo * c));
        if (ylo > 0) bound = fmin(bound, qy - (double)(ylo * c));
        if (xhi < cl.gw - 1) bound = fmin(bound, (double)((xhi + 1) * c) - qx);
        if (yhi < cl.gh - 1) bound = fmin(bound, (double)((yhi + 1) * c) - qy);
        if (bound == INFINITY) break;                         // whole grid visited
        if (bound > 0.0 && top.d2[K - 1] < bound * bound) break;   // strict: ties at the bound stay open
    }
}

// =======================================================================================
// scan-free linked cells: build (memset + 1 kernel) and the same exact ring search
// =======================================================================================
// Cell edge from the point density, evaluated on the device from the device-resident count (the run loop never
// brings it to the host).  bias_x4 / 4 scales the target cell area: 6 (about two points per cell, 2..16 px) for
// the dense seed sets of the thinning rounds, 3 (about one point per two cells: 4 px cells at one particle per
// ~30 px^2) for the steady state, where the queries' nearest neighbours are close and small cells mean fewer
// candidates (measured: metrics pass 38 us with 4 px cells, 45 us with 8 px).  AOSLO_LINK_SHIFT pins it.
__host__ __device__ inline int link_shift_for(int n, int H, int W, int min_shift, int fixed_shift, int bias_x4) {
    int shift = fixed_shift;
    if (shift <= 0) {
        const double area = 2.0 * (double)H * (double)W / (double)(n > 0 ? n : 1);
        shift = 1;
        while (shift < 4 && (double)(1 << (shift + 1)) * (double)(1 << (shift + 1)) <= area * 0.25 * (double)bias_x4) ++shift;
    }
    if (shift < min_shift) shift = min_shift;
    return shift;
}

// Build of the linked cells in ONE kernel, without clearing the heads: a head word is (tag << 32) | index and only
// words carrying the tag of the current build are valid.  The tag is a device counter: every thread reads it at
// kernel start, the LAST block to finish (ticket) publishes tag + geometry for the kernels that follow.
__global__ void __launch_bounds__(kThreads) link_kernel(const double2* __restrict__ pos, int n_host,
                                                        const int* d_n, int W, int H, int min_shift, int fixed_shift,
                                                        int bias_x4, int* geom, unsigned int* epoch, int* ticket,
                                                        unsigned long long* head, int* next, int* status) {
    const int n = d_n ? *d_n : n_host;
    const int shift = link_shift_for(n, H, W, min_shift, fixed_shift, bias_x4);
    const int gw = (W + (1 << shift) - 1) >> shift, gh = (H + (1 << shift) - 1) >> shift;
    const unsigned int tag = *epoch + 1u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double2 p = pos[i];
        bool bad = false;
        int cx = cell_coord(p.x, shift, gw, bad);
        int cy = cell_coord(p.y, shift, gh, bad);
        if (p.x != p.x || p.y != p.y) atomicOr(status, AOSLO_DEV_NAN);
        else if (bad || p.x >= (double)W || p.y >= (double)H) atomicOr(status, AOSLO_DEV_ESCAPED);
        const unsigned long long old =
            atomicExch(&head[cy * gw + cx], ((unsigned long long)tag << 32) | (unsigned long long)(unsigned int)i);
        next[i] = ((unsigned int)(old >> 32) == tag) ? (int)(unsigned int)old : -1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {      // every block has read the old epoch by now
            geom[0] = shift; geom[1] = gw; geom[2] = gh; geom[3] = gw * gh; geom[4] = (int)tag;
            *epoch = tag;
            *ticket = 0;
        }
    }
}

// n_host is the count, or (with d_n) only its launch-size upper bound
int build_links(aoslo_ctx* ctx, cudaStream_t s, CellLinks& lk, const double* d_pos, int n_host, const int* d_n,
                int bias_x4) {
    link_kernel<<<ctx->nb, kThreads, 0, s>>>((const double2*)d_pos, n_host, d_n, ctx->W, ctx->H, lk.min_shift,
                                             lk.fixed_shift, bias_x4, ctx->d_link_geom, ctx->d_link_epoch,
                                             ctx->d_link_epoch_ticket, lk.head, lk.next, ctx->d_status);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

inline LinkView link_view(aoslo_ctx* ctx, const CellLinks& lk, const double* d_pos) {
    return LinkView{ctx->d_link_geom, lk.head, lk.next, (const double2*)d_pos};
}

__device__ __forceinline__ int link_first(const LinkView& lv, int cell, unsigned int tag) {
    const unsigned long long h = lv.head[cell];
    return ((unsigned int)(h >> 32) == tag) ? (int)(unsigned int)h : -1;
}

template <int K>
__device__ __forceinline__ void knn_links_cell(const LinkView& lv, int cell, unsigned int tag, double qx, double qy,
                                               TopK<K>& top) {
    for (int p = link_first(lv, cell, tag); p >= 0; p = lv.next[p]) {
        double2 t = lv.pos[p];
        double dx = qx - t.x, dy = qy - t.y;
        top.push(dx * dx + dy * dy, p);
    }
}

// Thread-per-query ring search.  (A variant that walked the nine chains of the 3 x 3 block in lock step -- nine
// independent loads per step -- was measured: 80 instead of 32-64 registers, a third fewer resident threads, metrics
// pass 45 -> 56 us, thinning kNN unchanged: these kernels are bound by the number of resident queries, not by the
// length of one query's chain.)
template <int K>
__device__ __forceinline__ void knn_links_query(const LinkView& lv, double qx, double qy, TopK<K>& top) {
    top.init();
    const int shift = lv.geom[0], gw = lv.geom[1], gh = lv.geom[2];
    const unsigned int tag = (unsigned int)lv.geom[4];
    const int c = 1 << shift;
    int cx = (qx >= 0.0) ? ((int)fmin(qx, 1.0e9) >> shift) : 0;
    int cy = (qy >= 0.0) ? ((int)fmin(qy, 1.0e9) >> shift) : 0;
    cx = min(cx, gw - 1);
    cy = min(cy, gh - 1);
    const int rmax = max(gw, gh);
    for (int r = 0; r <= rmax; ++r) {
        const int ylo = cy - r, yhi = cy + r, xlo = cx - r, xhi = cx + r;
        for (int y = max(ylo, 0); y <= min(yhi, gh - 1); ++y) {
            if (y == ylo || y == yhi) {
                for (int x = max(xlo, 0); x <= min(xhi, gw - 1); ++x)
                    knn_links_cell<K>(lv, y * gw + x, tag, qx, qy, top);
            } else {
                if (xlo >= 0) knn_links_cell<K>(lv, y * gw + xlo, tag, qx, qy, top);
                if (xhi < gw) knn_links_cell<K>(lv, y * gw + xhi, tag, qx, qy, top);
            }
        }
        double bound = INFINITY;
        if (xlo > 0) bound = fmin(bound, qx - (double)(xlo * c));
        if (ylo > 0) bound = fmin(bound, qy - (double)(ylo * c));
        if (xhi < gw - 1) bound = fmin(bound, (double)((xhi + 1) * c) - qx);
        if (yhi < gh - 1) bound = fmin(bound, (double)((yhi + 1) * c) - qy);
        if (bound == INFINITY) break;
        if (bound > 0.0 && top.d2[K - 1] < bound * bound) break;
    }
}

template <int K>
__global__ void __launch_bounds__(kThreads) knn_links_kernel(LinkView lv, const double2* __restrict__ q,
                                                             int nq_host, const int* d_nq, int k,
                                                             int32_t* out_idx, double* out_d2, int* status) {
    int nq = d_nq ? *d_nq : nq_host;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += gridDim.x * blockDim.x) {
        double2 p = q[i];
        TopK<K> top;
        knn_links_query<K>(lv, p.x, p.y, top);
        bool shortfall = false;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (j < k) {
                bool empty = top.idx[j] == 0x7fffffff;
                shortfall |= empty;
                out_idx[(size_t)i * k + j] = empty ? -1 : top.idx[j];
                out_d2[(size_t)i * k + j] = top.d2[j];
            }
        }
        if (shortfall) atomicOr(status, AOSLO_DEV_KNN_SHORT);
    }
}

int knn_links(aoslo_ctx* ctx, cudaStream_t s, const LinkView& v, const double* d_query, int nq_host,
              const int* d_nq, int k, int32_t* d_idx, double* d_d2) {
    const double2* q = (const double2*)d_query;
    if (k < 1 || k > AOSLO_KMAX) return AOSLO_ERR_UNSUPPORTED;
    if (k == 1) knn_links_kernel<1><<<ctx->nb, kThreads, 0, s>>>(v, q, nq_host, d_nq, k, d_idx, d_d2, ctx->d_status);
    else if (k == 2) knn_links_kernel<2><<<ctx->nb, kThreads, 0, s>>>(v, q, nq_host, d_nq, k, d_idx, d_d2, ctx->d_status);
    else if (k <= 4) knn_links_kernel<4><<<ctx->nb, kThreads, 0, s>>>(v, q, nq_host, d_nq, k, d_idx, d_d2, ctx->d_status);
    else knn_links_kernel<8><<<ctx->nb, kThreads, 0, s>>>(v, q, nq_host, d_nq, k, d_idx, d_d2, ctx->d_status);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

// =======================================================================================
// warp-cooperative exact kNN: the 32 lanes split the cells of a ring, keep private top-K lists
// and merge them with shuffles; the dependent-load chain of a query shrinks from ~25 cells to
// 1-2, which is what bounds the in-loop kernels (only the few non-stable particles are queried).
// =======================================================================================
template <int K>
__device__ __forceinline__ void warp_merge_topk(const TopK<K>& mine, TopK<K>& out) {
    // K rounds of warp arg-min over the lanes' list heads by the strict order (d2, idx)
    TopK<K> t = mine;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double bd = t.d2[0];
        int bi = t.idx[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double od = __shfl_xor_sync(0xffffffffu, bd, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        out.d2[j] = bd;
        out.idx[j] = bi;
        if (t.idx[0] == bi && bi != 0x7fffffff) {          // this lane owned the winner: pop it
#pragma unroll
            for (int m = 0; m < K - 1; ++m) { t.d2[m] = t.d2[m + 1]; t.idx[m] = t.idx[m + 1]; }
            t.d2[K - 1] = INFINITY;
            t.idx[K - 1] = 0x7fffffff;
        }
    }
}

// every lane passes the same query; on return `out` (replicated in all lanes) holds the K nearest
template <int K>
__device__ __forceinline__ void knn_links_query_warp(const LinkView& lv, double qx, double qy, TopK<K>& out) {
    const int lane = threadIdx.x & 31;
    TopK<K> mine;
    mine.init();
    out.init();
    const int shift = lv.geom[0], gw = lv.geom[1], gh = lv.geom[2];
    const unsigned int tag = (unsigned int)lv.geom[4];
    const int c = 1 << shift;
    int cx = (qx >= 0.0) ? ((int)fmin(qx, 1.0e9) >> shift) : 0;
    int cy = (qy >= 0.0) ? ((int)fmin(qy, 1.0e9) >> shift) : 0;
    cx = min(cx, gw - 1);
    cy = min(cy, gh - 1);
    const int rmax = max(gw, gh);
    // rings 0 and 1 (the 3x3 block) are visited together -- lane j takes cell j -- so the common case costs
    // one round of dependent loads and one merge; further rings follow one at a time
    for (int r = 1; r <= rmax; ++r) {
        if (r == 1) {
            if (lane < 9) {
                const int x = cx + (lane % 3) - 1, y = cy + (lane / 3) - 1;
                if (x >= 0 && x < gw && y >= 0 && y < gh) knn_links_cell<K>(lv, y * gw + x, tag, qx, qy, mine);
            }
        } else {
            const int ncells = 8 * r;
            for (int j = lane; j < ncells; j += 32) {
                int dx, dy;
                if (j <= 2 * r) { dx = -r + j; dy = -r; }
                else if (j <= 4 * r + 1) { dx = -r + (j - 2 * r - 1); dy = r; }
                else if (j <= 6 * r) { dx = -r; dy = -r + 1 + (j - 4 * r - 2); }
                else { dx = r; dy = -r + 1 + (j - 6 * r - 1); }
                const int x = cx + dx, y = cy + dy;
                if (x >= 0 && x < gw && y >= 0 && y < gh) knn_links_cell<K>(lv, y * gw + x, tag, qx, qy, mine);
            }
        }
        warp_merge_topk<K>(mine, out);
        const int ylo = cy - r, yhi = cy + r, xlo = cx - r, xhi = cx + r;
        double bound = INFINITY;
        if (xlo > 0) bound = fmin(bound, qx - (double)(xlo * c));
        if (ylo > 0) bound = fmin(bound, qy - (double)(ylo * c));
        if (xhi < gw - 1) bound = fmin(bound, (double)((xhi + 1) * c) - qx);
        if (yhi < gh - 1) bound = fmin(bound, (double)((yhi + 1) * c) - qy);
        if (bound == INFINITY) break;
        if (bound > 0.0 && out.d2[K - 1] < bound * bound) break;
    }
}

// =======================================================================================
// exact kNN in sklearn's order: depth-first single-tree query of KDTree(leaf_size=1)
// (sklearn/neighbors/_binary_tree.pxi.tp:1603-1656, _kd_tree.pyx.tp:128-145,
//  sklearn/utils/_heap.pyx:heap_push, sklearn/utils/_sorting.pyx: insertion sort for n <= 15)
// =======================================================================================
struct KdView {
    int n, n_nodes;
    const int32_t* idx_array;
    const int32_t* node_start;
    const int32_t* node_end;
    const double* bounds;          // [n_nodes][4] lo_x lo_y hi_x hi_y
    const double2* data;
};

__device__ __forceinline__ double kd_min_rdist(const KdView& t, int node, double qx, double qy) {
    const double* b = t.bounds + (size_t)node * 4;
    double r = 0.0;
    {
        double d_lo = b[0] - qx, d_hi = qx - b[2];
        double d = (d_lo + fabs(d_lo)) + (d_hi + fabs(d_hi));
        double h = 0.5 * d;
        r += h * h;
    }
    {
        double d_lo = b[1] - qy, d_hi = qy - b[3];
        double d = (d_lo + fabs(d_lo)) + (d_hi + fabs(d_hi));
        double h = 0.5 * d;
        r += h * h;
    }
    return r;
}

__device__ __forceinline__ void sk_heap_push(double* vals, int* idxs, int size, double val, int val_idx) {
    if (val >= vals[0]) return;
    vals[0] = val; idxs[0] = val_idx;
    int cur = 0;
    while (true) {
        int l = 2 * cur + 1, r = l + 1, sw;
        if (l >= size) break;
        else if (r >= size) {
            if (vals[l] > val) sw = l; else break;
        } else if (vals[l] >= vals[r]) {
            if (val < vals[l]) sw = l; else break;
        } else {
            if (val < vals[r]) sw = r; else break;
        }
        vals[cur] = vals[sw]; idxs[cur] = idxs[sw];
        cur = sw;
    }
    vals[cur] = val; idxs[cur] = val_idx;
}

__global__ void __launch_bounds__(128) knn_sklearn_kernel(KdView t, const double2* __restrict__ q, int nq, int k,
                                                          int32_t* out_idx, double* out_d2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    double qx = q[i].x, qy = q[i].y;
    double vals[AOSLO_KMAX];
    int idxs[AOSLO_KMAX];
    for (int j = 0; j < k; ++j) { vals[j] = INFINITY; idxs[j] = 0; }
    int st_node[48];
    double st_lb[48];
    int sp = 0;
    st_node[0] = 0; st_lb[0] = kd_min_rdist(t, 0, qx, qy); sp = 1;
    while (sp > 0) {
        --sp;
        int node = st_node[sp];
        double lb = st_lb[sp];
        if (lb > vals[0]) continue;                      // trim
        if (2 * node + 1 >= t.n_nodes) {                 // leaf
            for (int a = t.node_start[node]; a < t.node_end[node]; ++a) {
                int id = t.idx_array[a];
                double2 p = t.data[id];
                double dx = qx - p.x, dy = qy - p.y;
                double d = dx * dx + dy * dy;            // euclidean_rdist: sequential sum of squares
                sk_heap_push(vals, idxs, k, d, id);
            }
        } else {
            int i1 = 2 * node + 1, i2 = i1 + 1;
            double lb1 = kd_min_rdist(t, i1, qx, qy), lb2 = kd_min_rdist(t, i2, qx, qy);
            if (lb1 <= lb2) {                            // nearer child first, left on ties
                st_node[sp] = i2; st_lb[sp] = lb2; ++sp;
                st_node[sp] = i1; st_lb[sp] = lb1; ++sp;
            } else {
                st_node[sp] = i1; st_lb[sp] = lb1; ++sp;
                st_node[sp] = i2; st_lb[sp] = lb2; ++sp;
            }
        }
    }
    // simultaneous_sort for n <= 15: insertion sort by value, stable w.r.t. the heap array
    for (int a = 1; a < k; ++a) {
        double tv = vals[a]; int ti = idxs[a];
        int b = a;
        while (b > 0 && vals[b - 1] > tv) { vals[b] = vals[b - 1]; idxs[b] = idxs[b - 1]; --b; }
        vals[b] = tv; idxs[b] = ti;
    }
    for (int j = 0; j < k; ++j) {
        out_idx[(size_t)i * k + j] = idxs[j];
        out_d2[(size_t)i * k + j] = vals[j];
    }
}

int knn_sklearn(aoslo_ctx* ctx, cudaStream_t s, const double* d_points, int n_points, const double* d_query,
                int n_query, int k, int32_t* d_idx, double* d_d2) {
    if (n_points > ctx->cap) return AOSLO_ERR_CAPACITY;
    AOSLO_TRY(kd_reserve(ctx));
    if (k < 1 || k > AOSLO_KMAX) return AOSLO_ERR_UNSUPPORTED;
    if (n_points < k) return AOSLO_ERR_INVALID;   // sklearn: ValueError("Expected n_neighbors <= n_samples_fit")
    // host round trip: sklearn's idx_array is the residue of libstdc++ std::nth_element.  Two trees stay cached
    // (typically the particles and the GT): the positions do not change between the metrics of one iteration and the
    // force step of the next, and the GT never changes
    AOSLO_CUDA(cudaMemcpyAsync(ctx->kd_stage, d_points, sizeof(double) * 2 * n_points, cudaMemcpyDeviceToHost, s));
    AOSLO_TRY(stream_wait(ctx, s));
    KdTreeDev* slots[2] = {&ctx->kd, &ctx->kd2};
    KdTreeDev* hit = nullptr;
    for (KdTreeDev* c : slots)
        if (c->cached_n == n_points && memcmp(c->h_points, ctx->kd_stage, sizeof(double) * 2 * n_points) == 0) hit = c;
    int n_nodes = aoslo_kdtree_num_nodes(n_points);
    if (hit) {
        ctx->kd_lru = (hit == slots[0]) ? 1 : 0;
    } else {
        hit = slots[ctx->kd_lru];
        ctx->kd_lru ^= 1;
        hit->cached_n = -1;
        memcpy(hit->h_points, ctx->kd_stage, sizeof(double) * 2 * n_points);
        AOSLO_TRY(aoslo_kdtree_build_host(hit->h_points, n_points, hit->h_idx_array, hit->h_node_start, hit->h_node_end,
                                          hit->h_node_bounds, n_nodes));
        AOSLO_CUDA(cudaMemcpyAsync(hit->idx_array, hit->h_idx_array, sizeof(int32_t) * n_points, cudaMemcpyHostToDevice, s));
        AOSLO_CUDA(cudaMemcpyAsync(hit->node_start, hit->h_node_start, sizeof(int32_t) * n_nodes, cudaMemcpyHostToDevice, s));
        AOSLO_CUDA(cudaMemcpyAsync(hit->node_end, hit->h_node_end, sizeof(int32_t) * n_nodes, cudaMemcpyHostToDevice, s));
        AOSLO_CUDA(cudaMemcpyAsync(hit->node_bounds, hit->h_node_bounds, sizeof(double) * 4 * n_nodes,
                                   cudaMemcpyHostToDevice, s));
        hit->cached_n = n_points;
    }
    KdTreeDev& kd = *hit;
    KdView v{n_points, n_nodes, kd.idx_array, kd.node_start, kd.node_end, kd.node_bounds, (const double2*)d_points};
    int blocks = (n_query + 127) / 128;
    if (blocks > 0) {
        knn_sklearn_kernel<<<blocks, 128, 0, s>>>(v, (const double2*)d_query, n_query, k, d_idx, d_d2);
        AOSLO_LAUNCH_CHECK();
    }
    // the pinned staging buffers are reused by the next call
    AOSLO_TRY(stream_wait(ctx, s));
    return AOSLO_OK;
}

// =======================================================================================
// seeding: raster-order compaction of the inner-region mask R_b > threshold
// =======================================================================================
template <typename T>
struct SeedOp {
    const T* Rb;
    int W, bx, by, wi;
    double thr;
    const double* d_thr;        // device-resident threshold of the graph path (overrides thr)
    double2* out;
    int cap;
    int* status;
    __device__ int value(int m) const {
        int y = by + m / wi, x = bx + m % wi;
        const double t = d_thr ? *d_thr : thr;
        return ((double)Rb[(size_t)y * fpitch(W) + x] > t) ? 1 : 0;
    }
    __device__ void emit(int m, int v, int dst) const {
        if (!v) return;
        if (dst >= cap) { atomicOr(status, AOSLO_DEV_CAPACITY); return; }
        int y = by + m / wi, x = bx + m % wi;
        out[dst] = make_double2((double)x, (double)y);
    }
};

// =======================================================================================
// deletion (utils.py:245-278)
// =======================================================================================
// Row i of the kNN(k=2) table is (from, to, dist).  The reference walks the rows in order and
// deletes the endpoint with the lower blobness unless one endpoint is already deleted; the
// outcome of a row therefore depends only on earlier rows sharing an endpoint.  Wavefront
// rounds: in every round each unresolved row stamps both endpoints with (round, ~row) by
// atomicMax -- the lowest unresolved row wins an endpoint -- and a row that owns both of its
// endpoints is resolved.  Rows resolved in one round never share an endpoint, so the result is
// exactly the sequential one.
struct DeleteParams {
    const double2* pos;
    const uint8_t* stable;
    int n_host;
    const int* d_n;
    const int32_t* knn_idx;     // [n][2]
    const double* knn_d2;       // [n][2]
    double thr;
    const double* d_thr;        // overrides thr (graph path)
    int W, H;
    unsigned long long* touch;
    uint8_t* del;
    uint8_t* rowstate;          // 0 inactive / resolved, 1 pending
    int* counters;              // [3] rotating "rows still pending"
    int* rounds_out;
    int* status;
};

constexpr int kCoopThreads = 1024;     // block size of the cooperative (grid-synchronising) deletion kernel

template <typename T>
__device__ __forceinline__ double energy_at(const T* Rb, double2 p, int W, int H, int* status) {
    int ix = (int)p.x, iy = (int)p.y;
    if (ix < 0 || ix >= W || iy < 0 || iy >= H || p.x != p.x || p.y != p.y) {
        atomicOr(status, (p.x != p.x || p.y != p.y) ? AOSLO_DEV_NAN : AOSLO_DEV_ESCAPED);
        return 0.0;
    }
    return (double)Rb[(size_t)iy * fpitch(W) + ix];
}

template <typename T>
__global__ void __launch_bounds__(kCoopThreads) delete_resolve_kernel(DeleteParams p, const T* Rb) {
    cg::grid_group grid = cg::this_grid();
    const int n = p.d_n ? *p.d_n : p.n_host;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthr = gridDim.x * blockDim.x;
    const double thr = p.d_thr ? *p.d_thr : p.thr;
    __shared__ int s_pending;

    for (int i = tid; i < n; i += nthr) {
        int from = p.knn_idx[2 * i], to = p.knn_idx[2 * i + 1];
        double dist = sqrt(p.knn_d2[2 * i + 1]);
        bool active = (to >= 0) && (from >= 0) && !p.stable[from] && (dist < thr);
        p.rowstate[i] = active ? 1 : 0;
        p.del[i] = 0;
        p.touch[i] = 0ull;
    }
    if (tid < 3) p.counters[tid] = 0;
    grid.sync();

    int round = 1;
    for (;; ++round) {
        const unsigned long long hi = (unsigned long long)round << 32;
        for (int i = tid; i < n; i += nthr) {
            if (p.rowstate[i] != 1) continue;
            unsigned long long key = hi | (unsigned long long)(0xffffffffu - (unsigned)i);
            atomicMax(&p.touch[p.knn_idx[2 * i]], key);
            atomicMax(&p.touch[p.knn_idx[2 * i + 1]], key);
        }
        if (tid == 0) p.counters[(round + 1) % 3] = 0;
        grid.sync();
        if (threadIdx.x == 0) s_pending = 0;
        __syncthreads();
        int pending = 0;
        for (int i = tid; i < n; i += nthr) {
            if (p.rowstate[i] != 1) continue;
            int from = p.knn_idx[2 * i], to = p.knn_idx[2 * i + 1];
            unsigned long long key = hi | (unsigned long long)(0xffffffffu - (unsigned)i);
            if (__ldcg(&p.touch[from]) == key && __ldcg(&p.touch[to]) == key) {
                if (!__ldcg(&p.del[from]) && !__ldcg(&p.del[to])) {
                    double e1 = energy_at<T>(Rb, p.pos[from], p.W, p.H, p.status);
                    double e2 = energy_at<T>(Rb, p.pos[to], p.W, p.H, p.status);
                    p.del[(e1 < e2) ? from : to] = 1;
                }
                p.rowstate[i] = 0;
            } else {
                ++pending;
            }
        }
        if (pending) atomicAdd(&s_pending, pending);
        __syncthreads();
        if (threadIdx.x == 0 && s_pending) atomicAdd(&p.counters[round % 3], s_pending);
        grid.sync();
        if (p.counters[round % 3] == 0) break;
    }
    if (tid == 0 && p.rounds_out) *p.rounds_out = round;
}

// ---- thinning to the fixpoint (aoslo_run, canonical order) in the PIXEL DOMAIN -----------------------------------
// reference :101-120: seeds = inner pixels with R_b > threshold in raster order, then `while deleted:
// del_close_particles(thr = 3)`.  The seeds are distinct integer pixels, the order-preserving compaction keeps them in
// raster order, and a round only ever looks at neighbours closer than the threshold -- so the whole loop can be stated
// on a bitmap of alive pixels, without particle indices and without a compaction per round:
//   * row order = raster order of the pixels; the canonical (d^2, index) order among equidistant neighbours is the
//     raster order of their pixels, i.e. the stencil offsets sorted by (d^2, dy, dx);
//   * phase R (per round): the nearest alive neighbour of all 32 pixels of a bitmap word at once -- for every stencil
//     offset, in order, `center & shifted(row y+dy by dx) & ~found`; the offset's number goes into six bit planes;
//   * wavefront resolution of the greedy walk (utils.py:249-263) exactly as in delete_resolve_kernel: every pending
//     row stamps both endpoints with (wave, ~pixel) by atomicMax, a row owning both endpoints acts, the victim's bit
//     is cleared; the waves of all rounds run inside ONE cooperative kernel (one grid barrier per wave, stamp
//     buffers alternating), a round without a deletion ends the loop;
//   * the survivors are listed once, at the end (ordered scan over the bitmap words).
// Replaces, per run: the seed compaction (2 kernels over the R_b field), and per round (5 on the benchmark frame)
// a pixel-map scatter, a row kernel, the cooperative resolve, a compaction (2 kernels) and a copy.
constexpr int kThinOffsets = 44;               // |d| < 4: thresholds up to 4 px
constexpr int kThinPlanes = 6;
constexpr int kThinThreads = 512;
constexpr int kThinMaxWaves = 20000;
__constant__ signed char c_thin_dx[kThinOffsets] = {0, -1, 1, 0, -1, 1, -1, 1, 0, -2, 2, 0, -1, 1, -2, 2, -2, 2, -1, 1, -2, 2,
                                                    -2, 2, 0, -3, 3, 0, -1, 1, -3, 3, -3, 3, -1, 1, -2, 2, -3, 3, -3, 3, -2, 2};
__constant__ signed char c_thin_dy[kThinOffsets] = {-1, 0, 0, 1, -1, -1, 1, 1, -2, 0, 0, 2, -2, -2, -1, -1, 1, 1, 2, 2, -2, -2,
                                                    2, 2, -3, 0, 0, 3, -3, -3, -1, -1, 1, 1, 3, 3, -3, -3, -2, -2, 2, 2, 3, 3};
struct ThinOff { int dx, dy, d2; };
__host__ __device__ constexpr ThinOff thin_off(int k) {
    constexpr int dx[kThinOffsets] = {0, -1, 1, 0, -1, 1, -1, 1, 0, -2, 2, 0, -1, 1, -2, 2, -2, 2, -1, 1, -2, 2,
                                      -2, 2, 0, -3, 3, 0, -1, 1, -3, 3, -3, 3, -1, 1, -2, 2, -3, 3, -3, 3, -2, 2};
    constexpr int dy[kThinOffsets] = {-1, 0, 0, 1, -1, -1, 1, 1, -2, 0, 0, 2, -2, -2, -1, -1, 1, 1, 2, 2, -2, -2,
                                      2, 2, -3, 0, 0, 3, -3, -3, -1, -1, 1, 1, 3, 3, -3, -3, -2, -2, 2, 2, 3, 3};
    return ThinOff{dx[k], dy[k], dx[k] * dx[k] + dy[k] * dy[k]};
}

struct ThinParams {
    int H, W, bx, by, wpr, cap;
    double seed_thr, thin_thr;
    const RunParams* rp;            // device-resident thresholds of the graph path (override the two above)
    uint32_t* alive;                // [H][wpr] bitmap of the alive pixels, bit x at word 1 + x / 32
    uint32_t* planes;               // [kThinPlanes][H * wpr] stencil number of every pending pixel's neighbour
    uint32_t* pend;                 // [H * wpr] rows still pending in the current round
    unsigned long long* touch[2];   // [H * W] wavefront stamps per pixel, waves alternate between the two
    int* ctr;                       // [8], zero at launch: 0..3 rotating pending counts, 4 deletions, 5 seeds
    RunCtl* ctl;
    int* status;
};

// neighbour mask of a bitmap word: bit b = pixel (x + dx) of the row whose words around this column are L, C, R
template <int DX>
__device__ __forceinline__ uint32_t thin_shift(uint32_t L, uint32_t C, uint32_t R) {
    if (DX > 0) return __funnelshift_r(C, R, DX);
    if (DX < 0) return __funnelshift_l(L, C, -DX);
    return C;
}

template <int K>
struct ThinScan {      // offsets K .. kThinOffsets-1, in order
    __device__ static __forceinline__ void run(const uint32_t (&L)[7], const uint32_t (&C)[7], const uint32_t (&R)[7],
                                               int lim, uint32_t center, uint32_t& found, uint32_t (&pl)[kThinPlanes]) {
        constexpr ThinOff o = thin_off(K);
        if (K < lim) {
            const uint32_t m = center & thin_shift<o.dx>(L[o.dy + 3], C[o.dy + 3], R[o.dy + 3]) & ~found;
            found |= m;
#pragma unroll
            for (int b = 0; b < kThinPlanes; ++b)
                if ((K >> b) & 1) pl[b] |= m;
        }
        ThinScan<K + 1>::run(L, C, R, lim, center, found, pl);
    }
};
template <>
struct ThinScan<kThinOffsets> {
    __device__ static __forceinline__ void run(const uint32_t (&)[7], const uint32_t (&)[7], const uint32_t (&)[7], int,
                                               uint32_t, uint32_t&, uint32_t (&)[kThinPlanes]) {}
};

// One bitmap word of a wave.  CHECK = false: every pending row of the word stamps its two endpoints for wave `wave`.
// CHECK = true: every pending row tests whether it owns both endpoints in wave `wave`; an owner acts (the endpoint with
// the lower blobness dies unless one is dead already, utils.py:257-263) and leaves the pending set, the others stamp
// for wave + 1.
template <typename T, bool CHECK>
__device__ __forceinline__ void thin_wave_word(const ThinParams& p, const T* __restrict__ Rb, int t, int wave,
                                               int* pending, int* deleted) {
    const uint32_t pm0 = p.pend[t];
    if (!pm0) return;
    const int W = p.W, wpr = p.wpr;
    const size_t nwords = (size_t)p.H * wpr, fp = fpitch(W);
    uint32_t pm = pm0, keep = pm0;
    const int y = t / wpr, x0 = (t - y * wpr - 1) * 32;
    uint32_t pl[kThinPlanes];
#pragma unroll
    for (int b = 0; b < kThinPlanes; ++b) pl[b] = p.planes[(size_t)b * nwords + t];
    const unsigned long long hi = (unsigned long long)(unsigned)wave << 32;
    unsigned long long* cur = (wave & 1) ? p.touch[1] : p.touch[0];      // no dynamic index into the parameters
    unsigned long long* nxt = (wave & 1) ? p.touch[0] : p.touch[1];
    while (pm) {
        const int b = __ffs((int)pm) - 1;
        pm &= pm - 1u;
        int code = 0;
#pragma unroll
        for (int k = 0; k < kThinPlanes; ++k) code |= (int)((pl[k] >> b) & 1u) << k;
        const int dx = c_thin_dx[code], dy = c_thin_dy[code];
        const int x = x0 + b, pix = y * W + x, q = pix + dy * W + dx;
        const unsigned long long low = (unsigned long long)(0xffffffffu - (unsigned)pix);
        if (!CHECK) {
            atomicMax(&cur[pix], hi | low);
            atomicMax(&cur[q], hi | low);
        } else if (__ldcg(&cur[pix]) == (hi | low) && __ldcg(&cur[q]) == (hi | low)) {
            const int tq = (y + dy) * wpr + 1 + ((x + dx) >> 5), bq = (x + dx) & 31;
            const bool a_from = (__ldcg(&p.alive[t]) >> b) & 1u;
            const bool a_to = (__ldcg(&p.alive[tq]) >> bq) & 1u;
            if (a_from && a_to) {
                const double e1 = (double)Rb[(size_t)y * fp + x];
                const double e2 = (double)Rb[(size_t)(y + dy) * fp + x + dx];
                if (e1 < e2) atomicAnd(&p.alive[t], ~(1u << b));
                else atomicAnd(&p.alive[tq], ~(1u << bq));
                ++*deleted;
            }
            keep &= ~(1u << b);
        } else {
            const unsigned long long key = ((unsigned long long)(unsigned)(wave + 1) << 32) | low;
            atomicMax(&nxt[pix], key);
            atomicMax(&nxt[q], key);
            ++*pending;
        }
    }
    if (CHECK) p.pend[t] = keep;
}

template <typename T>
__global__ void __launch_bounds__(kThinThreads) thin_pixels_kernel(ThinParams p, const T* __restrict__ Rb) {
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int H = p.H, W = p.W, wpr = p.wpr, nwords = H * wpr;
    const size_t fp = fpitch(W);
    const double seed_thr = p.rp ? p.rp->seed_thr : p.seed_thr;
    const double thin_thr = p.rp ? p.rp->thin_thr : p.thin_thr;
    __shared__ int s_pending, s_deleted;

    // ---- seeds (utils.py:359-375): a warp per bitmap word, a lane per pixel
    {
        int nseed = 0;
        for (int t = tid >> 5; t < nwords; t += nthr >> 5) {
            const int y = t / wpr, x = (t - y * wpr - 1) * 32 + lane;
            bool sd = false;
            if (x >= p.bx && x < W - p.bx && y >= p.by && y < H - p.by) {
                sd = (double)Rb[(size_t)y * fp + x] > seed_thr;
                if (sd) { p.touch[0][(size_t)y * W + x] = 0ull; p.touch[1][(size_t)y * W + x] = 0ull; }
            }
            const uint32_t m = __ballot_sync(0xffffffffu, sd);
            if (lane == 0) { p.alive[t] = m; nseed += __popc(m); }
        }
        if (nseed) atomicAdd(&p.ctr[5], nseed);
    }
    grid.sync();
    const int n_seeds = __ldcg(&p.ctr[5]);
    int lim = 0;                                  // offsets with d^2 < thr^2 (exact: small integers)
    {
        const double thr2 = thin_thr * thin_thr;
        for (int k = 0; k < kThinOffsets; ++k) {
            const int dx = c_thin_dx[k], dy = c_thin_dy[k];
            if ((double)(dx * dx + dy * dy) < thr2) lim = k + 1;
        }
    }
    int rounds = 0, wave = 0, deleted_before = 0;
    bool overflow = false;
    while (n_seeds >= 2) {
        // ---- phase R: nearest alive neighbour of every alive pixel, 32 pixels at a time
        for (int t = tid; t < nwords; t += nthr) {
            const uint32_t center = __ldcg(&p.alive[t]);     // L2: cleared by other SMs' atomics in earlier waves
            uint32_t found = 0u, pl[kThinPlanes];
#pragma unroll
            for (int b = 0; b < kThinPlanes; ++b) pl[b] = 0u;
            if (center) {                          // never a pad word: 1 <= t % wpr <= wpr - 2
                const int y = t / wpr;
                uint32_t L[7], C[7], R[7];
#pragma unroll
                for (int d = 0; d < 7; ++d) {
                    const int yy = y + d - 3;
                    const bool in = yy >= 0 && yy < H;
                    const uint32_t* row = p.alive + t + (d - 3) * wpr;
                    L[d] = in ? __ldcg(row - 1) : 0u;
                    C[d] = in ? __ldcg(row) : 0u;
                    R[d] = in ? __ldcg(row + 1) : 0u;
                }
                ThinScan<0>::run(L, C, R, lim, center, found, pl);
            }
            p.pend[t] = found;
#pragma unroll
            for (int b = 0; b < kThinPlanes; ++b) p.planes[(size_t)b * nwords + t] = pl[b];
        }
        // ---- wavefront rounds of this thinning round.  One grid barrier per wave: a row that is still pending after
        // its check of wave w stamps for wave w+1 right away, into the OTHER stamp buffer (checks of wave w read buffer
        // w & 1, which nobody writes between the two barriers); "rows pending after wave w" is read after the next
        // barrier.  Stamps carry the wave number, so the stale ones of wave w-1 in a buffer lose against wave w+1.
        const int w_first = wave + 1;
        for (int t = tid; t < nwords; t += nthr) thin_wave_word<T, false>(p, Rb, t, wave + 1, nullptr, nullptr);
        for (;;) {
            ++wave;
            if (tid == 0) p.ctr[(wave + 1) & 3] = 0;     // added to during wave - 3, read during wave - 2
            if (threadIdx.x == 0) { s_pending = 0; s_deleted = 0; }
            grid.sync();
            if (wave > w_first && __ldcg(&p.ctr[(wave - 1) & 3]) == 0) break;   // nothing was stamped for this wave
            if (wave - w_first >= kThinMaxWaves) { overflow = true; break; }
            int pending = 0, deleted = 0;
            for (int t = tid; t < nwords; t += nthr) thin_wave_word<T, true>(p, Rb, t, wave, &pending, &deleted);
            if (pending) atomicAdd(&s_pending, pending);
            if (deleted) atomicAdd(&s_deleted, deleted);
            __syncthreads();
            if (threadIdx.x == 0) {
                if (s_pending) atomicAdd(&p.ctr[wave & 3], s_pending);
                if (s_deleted) atomicAdd(&p.ctr[4], s_deleted);
            }
        }
        ++rounds;
        const int deleted_total = __ldcg(&p.ctr[4]);
        const bool none = deleted_total == deleted_before;
        deleted_before = deleted_total;
        if (none || overflow) break;
    }
    if (tid == 0) {
        p.ctl->n_seeds = n_seeds;
        p.ctl->n_after_thinning = n_seeds - deleted_before;
        p.ctl->thin_rounds = rounds;
        p.ctl->thin_cond = 0;
        // the reference's particle list would not fit (seeds), or the wavefront did not terminate (never observed)
        if (n_seeds > p.cap || overflow) atomicOr(p.status, AOSLO_DEV_CAPACITY);
    }
}

struct AliveEmitOp {               // the survivors in raster order -> particle rows, not stable
    const uint32_t* alive;
    int wpr;
    double2* pos;
    uint8_t* stable;
    int cap;
    int* status;
    __device__ int value(int t) const { return __popc(alive[t]); }
    __device__ void emit(int t, int v, int dst) const {
        if (!v) return;
        const int y = t / wpr, x0 = (t - y * wpr - 1) * 32;
        uint32_t m = alive[t];
        while (m) {
            const int b = __ffs((int)m) - 1;
            m &= m - 1u;
            if (dst >= cap) { atomicOr(status, AOSLO_DEV_CAPACITY); return; }
            pos[dst] = make_double2((double)(x0 + b), (double)y);
            stable[dst] = 0;
            ++dst;
        }
    }
};

// kNN(k=2) rows of the deletion step on the linked cells + compact list of the active rows
// (rows whose `from` is not stable and whose neighbour is closer than the threshold)
__global__ void __launch_bounds__(kThreads) knn2_rows_kernel(LinkView lv, const uint8_t* __restrict__ stable,
                                                             int n_host, const int* d_n, double thr,
                                                             const double* d_thr, int32_t* knn_idx, double* knn_d2,
                                                             uint8_t* del, int* active, int* n_active, int* status) {
    int n = d_n ? *d_n : n_host;
    if (d_thr) thr = *d_thr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double2 p = lv.pos[i];
        TopK<2> top;
        knn_links_query<2>(lv, p.x, p.y, top);
        bool ok = top.idx[1] != 0x7fffffff;
        if (!ok) atomicOr(status, AOSLO_DEV_KNN_SHORT);
        int from = top.idx[0], to = ok ? top.idx[1] : -1;
        knn_idx[2 * (size_t)i] = from;
        knn_idx[2 * (size_t)i + 1] = to;
        knn_d2[2 * (size_t)i] = top.d2[0];
        knn_d2[2 * (size_t)i + 1] = top.d2[1];
        del[i] = 0;
        if (ok && !stable[from] && sqrt(top.d2[1]) < thr) active[atomicAdd(n_active, 1)] = i;
    }
}

// same collection from an existing kNN table (sklearn-order path)
__global__ void __launch_bounds__(kThreads) collect_active_kernel(const int32_t* __restrict__ knn_idx,
                                                                  const double* __restrict__ knn_d2,
                                                                  const uint8_t* __restrict__ stable, int n_host,
                                                                  const int* d_n, double thr, uint8_t* del,
                                                                  int* active, int* n_active) {
    int n = d_n ? *d_n : n_host;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int from = knn_idx[2 * (size_t)i], to = knn_idx[2 * (size_t)i + 1];
        del[i] = 0;
        if (to >= 0 && from >= 0 && !stable[from] && sqrt(knn_d2[2 * (size_t)i + 1]) < thr)
            active[atomicAdd(n_active, 1)] = i;
    }
}

// The same wavefront resolution as delete_resolve_kernel, for the in-loop case where only the
// rows of the few non-stable particles are active: ONE block walks the compact active list, so it
// needs no cooperative launch / grid sync and does not monopolise the GPU (other frames keep
// running next to it).  `touch` is never reset: keys carry a per-call epoch.
template <typename T>
__global__ void __launch_bounds__(kScanThreads) delete_resolve_small_kernel(
    const double2* __restrict__ pos, int32_t* knn_idx, double* knn_d2, const int* __restrict__ active,
    int* n_active, int n_host, const int* d_n, unsigned int* d_epoch, int W, int H, const T* __restrict__ Rb,
    unsigned long long* touch, uint8_t* del, uint8_t* astate, int* block_sums, int nb_scan, int* d_n_out,
    int* status) {
    __shared__ int s_pending;
    __shared__ int s_hist[kMaxScanBlocks];      // deletions per block of the compaction that follows
    __shared__ int sw[33];
    const int m = d_n ? *d_n : n_host;
    // per-call epoch from a device counter (so the call can be replayed from a CUDA graph)
    const unsigned int epoch = *d_epoch;
    for (int b = threadIdx.x; b < kMaxScanBlocks; b += blockDim.x) s_hist[b] = 0;
    __syncthreads();
    const int na = *n_active;
    // span of one compaction block (scan_span with blockDim = kScanThreads, gridDim = nb_scan)
    int per = 1;
    if (block_sums) {
        per = (m + nb_scan - 1) / nb_scan;
        per = ((per + kScanThreads - 1) / kScanThreads) * kScanThreads;
        if (per < 1) per = kScanThreads;
    }
    for (int a = threadIdx.x; a < na; a += blockDim.x) astate[a] = 1;
    __syncthreads();
    for (unsigned int round = 1; na > 0; ++round) {
        const unsigned long long hi = ((unsigned long long)epoch << 44) | ((unsigned long long)round << 32);
        for (int a = threadIdx.x; a < na; a += blockDim.x) {
            if (!astate[a]) continue;
            int i = active[a];
            unsigned long long key = hi | (unsigned long long)(0xffffffffu - (unsigned)i);
            atomicMax(&touch[__ldcg(&knn_idx[2 * (size_t)i])], key);
            atomicMax(&touch[__ldcg(&knn_idx[2 * (size_t)i + 1])], key);
        }
        if (threadIdx.x == 0) s_pending = 0;
        __threadfence();
        __syncthreads();
        int pending = 0;
        for (int a = threadIdx.x; a < na; a += blockDim.x) {
            if (!astate[a]) continue;
            int i = active[a];
            int from = __ldcg(&knn_idx[2 * (size_t)i]), to = __ldcg(&knn_idx[2 * (size_t)i + 1]);
            unsigned long long key = hi | (unsigned long long)(0xffffffffu - (unsigned)i);
            if (__ldcg(&touch[from]) == key && __ldcg(&touch[to]) == key) {
                if (!__ldcg(&del[from]) && !__ldcg(&del[to])) {
                    double e1 = energy_at<T>(Rb, pos[from], W, H, status);
                    double e2 = energy_at<T>(Rb, pos[to], W, H, status);
                    const int victim = (e1 < e2) ? from : to;
                    __stcg(&del[victim], (uint8_t)1);     // every deletion is performed exactly once
                    if (block_sums) atomicAdd(&s_hist[victim / per], 1);
                }
                astate[a] = 0;
            } else {
                ++pending;
            }
        }
        if (pending) atomicAdd(&s_pending, pending);
        __threadfence();
        __syncthreads();
        int left = s_pending;
        __syncthreads();
        if (left == 0) break;
        if (round >= 4000u) { if (threadIdx.x == 0) atomicOr(status, AOSLO_DEV_CAPACITY); break; }
    }
    __syncthreads();
    // offsets of the order-preserving compaction that follows (scan_emit_kernel<KeepOp> with nb_scan
    // blocks): kept elements before block b = min(b * per, m) - deletions before it.  Replaces the count and
    // spine launches of the generic three-phase scan.
    if (block_sums) {
        const int v = (threadIdx.x < nb_scan) ? s_hist[threadIdx.x] : 0;
        int total;
        const int excl = block_excl_scan(v, sw, total);
        if (threadIdx.x < nb_scan) {
            long long lo = (long long)threadIdx.x * per;
            if (lo > m) lo = m;
            block_sums[threadIdx.x] = (int)lo - excl;
        }
        if (threadIdx.x == 0 && d_n_out) *d_n_out = m - total;
    }
    if (threadIdx.x == 0) {                   // ready for the next call
        *d_epoch = epoch + 1;
        *n_active = 0;
    }
}

struct KeepOp {
    const double2* pos;
    const uint8_t* stable;
    const uint8_t* del;
    double2* pos_out;
    uint8_t* stable_out;
    __device__ int value(int i) const { return del[i] ? 0 : 1; }
    __device__ void emit(int i, int v, int dst) const {
        if (!v) return;
        pos_out[dst] = pos[i];
        stable_out[dst] = stable[i];
    }
};

// =======================================================================================
// pit energy (utils.py:82-87 with scipy.stats written out) and the LUT (utils.py:95-109)
// =======================================================================================
__device__ __forceinline__ double energy_pit2(double dist, double mu, double sigma, double lam) {
    const double kSqrt2Pi = 2.5066282746310002;   // np.sqrt(2 * np.pi)
    double z = dist - lam;
    double e = (z >= 0.0) ? exp(-z) : 0.0;        // expon(loc=lam).pdf: 0 left of loc
    double y = (dist - mu) / sigma;
    double nrm = exp(-(y * y) / 2.0) / kSqrt2Pi / sigma;
    return (dist <= mu) ? (e - nrm + 0.48) : (e + nrm - 0.48);
}

__global__ void dist_lut_kernel(int rfs, double mu, double sigma, double lam, double* lut, const RunParams* rp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rfs * rfs) return;
    if (rp) { mu = rp->lut_mu; sigma = rp->lut_sigma; lam = rp->lut_lam; }
    int y = i / rfs, x = i - y * rfs;
    double dx = (double)(x - rfs / 2), dy = (double)(y - rfs / 2);
    lut[i] = energy_pit2(sqrt(dx * dx + dy * dy), mu, sigma, lam);
}

// =======================================================================================
// gravity field (utils.py:310-333) from an occupancy bitmap of the particles' pixels
// =======================================================================================
// field[y][x] = max(-0.48, max over particles p with |x-px|<=h, |y-py|<=h of LUT[y-py+h][x-px+h]), px = int(p[0]).
// Only the pixel a particle sits on matters, so the particles are scattered into a bitmap (one bit per pixel, one
// zero word on each side of a row) and a tile reads the (TY + 2h) x (128 + 2h) bits around it from shared memory.
//
// Nearest-particle path (gravity_nearest_kernel).  The reference's LUT is a function of the distance alone and, for
// most parameters (all named configurations), non-increasing in it.  lut_radial_kernel checks both properties on
// the LUT ENTRIES THEMSELVES -- bit-equal values for equal dx^2+dy^2, T[a] >= T[b] for consecutive attainable
// a < b -- and then the maximum over the window is exactly the entry of the smallest squared distance: a
// windowed squared-distance transform.  Row pass: the nearest set bit of a row within |dx| <= h is two bit scans
// on the 2h+1 window bits.  Column pass: d2(y) = min over dy of r2(y+dy) + dy^2, two output rows per register as
// 16-bit halves, one VIADDMNMX.U16x2 (min(a + imm, c) per halfword) per (source row, output row pair).
// ~40 instructions per pixel instead of the 190 of the former per-candidate gather over a sorted cell list.
//
// General path (gravity_bits_kernel): any other LUT (a caller's table, sigma_dist_func < 0.83 where the pit
// function jumps upwards at mu, h > 15).  Every set bit of a window row is a candidate; its TY entries for the
// thread's output rows are consecutive in a column-major copy of the LUT padded with -inf rows.
// Both kernels are always launched; the device flag written by lut_radial_kernel decides which one works.
constexpr int kGravCols = 128;         // columns per tile = threads per CTA
constexpr int kGravOccWords = 8;       // staged words per bitmap row: 6 cover [x0-32, x0+160), 2 zero words
constexpr int kNearRows = 32;          // output rows per thread, nearest-particle path
constexpr int kBitsRows = 16;          // output rows per thread, general path
constexpr int kNearMaxH = 15;          // 2h+1 <= 31 window bits: one 32-bit funnel shift
constexpr unsigned kD2Inf = 0x4000u;   // "no particle" in a 16-bit half; kD2Inf + kD2Inf does not wrap

__host__ __device__ inline int occ_words_per_row(int W) { return (W + 31) / 32 + 2; }

// `ok` (optional): set to 1 here; the scatter that follows clears it when a point is not an in-image integer pixel
__global__ void __launch_bounds__(kThreads) occ_clear_kernel(uint4* occ, int n16, int* ok) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x)
        occ[i] = make_uint4(0u, 0u, 0u, 0u);
    if (ok && blockIdx.x == 0 && threadIdx.x == 0) *ok = 1;
}

// Gravity use (ok == nullptr): positions must lie in [0, W) x [0, H); anything else is reported (the reference
// would raise IndexError or silently wrap a negative index) and leaves no bit; the pixel is int(particle[0]).
// Nearest-distance use (ok != nullptr, metrics): the bitmap answers distance queries only if every point IS an
// integer pixel inside the image; otherwise *ok = 0 and the caller keeps to its cell-list search.  Tombstones
// (del[i] != 0) are skipped.
__global__ void __launch_bounds__(kThreads) occ_scatter_kernel(const double2* __restrict__ pos,
                                                               const uint8_t* __restrict__ del, int n_host,
                                                               const int* d_n, int W, int H, int wpr, uint32_t* occ,
                                                               int* status, int* ok) {
    int n = d_n ? *d_n : n_host;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (del && del[i]) continue;
        double2 p = pos[i];
        const bool nan = p.x != p.x || p.y != p.y;
        const bool out = !(p.x >= 0.0) || !(p.y >= 0.0) || p.x >= (double)W || p.y >= (double)H;
        if (nan || out) {
            if (ok) *ok = 0;
            else atomicOr(status, nan ? AOSLO_DEV_NAN : AOSLO_DEV_ESCAPED);
            continue;
        }
        const int ix = (int)p.x, iy = (int)p.y;        // int(particle[0]) truncation
        if (ok && ((double)ix != p.x || (double)iy != p.y)) *ok = 0;
        atomicOr(&occ[(size_t)iy * wpr + 1 + (ix >> 5)], 1u << (ix & 31));
    }
}

// Squared distance from the integer pixel (x, y) to the nearest set bit, searched in the rows y, y-+1, y-+2, ...
// through the window |dx| <= 15; a result <= 15^2 is the global minimum (every point outside the window is farther
// than 15), anything else returns -1 and the caller searches its cell list.  Distances between integer pixels are
// exact in both searches, so sqrt() of either gives the same double.
__device__ __forceinline__ int occ_nearest_d2(const uint32_t* __restrict__ occ, int wpr, int H, int x, int y) {
    constexpr int R = 15;
    const int b = x + 32 - R;                          // padded bit of pixel x - R
    const int word = b >> 5, sh = b & 31;              // word + 1 <= wpr - 1 for every x < W
    int best = 0x7fffffff;
    for (int dy = 0; dy <= R; ++dy) {
        if (dy * dy >= best) break;
        for (int side = 0; side < (dy ? 2 : 1); ++side) {
            const int yy = side ? y + dy : y - dy;
            if (yy < 0 || yy >= H) continue;
            const uint32_t* row = occ + (size_t)yy * wpr + word;
            const uint32_t w = __funnelshift_r(__ldg(row), __ldg(row + 1), sh);      // bits 0..30 = x-15 .. x+15
            const uint32_t wr = (w >> R) & 0xffffu, wl = w & 0x7fffu;
            const unsigned rr = (unsigned)(__ffs((int)wr) - 1);
            const unsigned rl = (unsigned)(R - 31 + __clz((int)wl));
            const unsigned r = rr < rl ? rr : rl;
            if (r <= (unsigned)R) {
                const int c = (int)(r * r) + dy * dy;
                best = c < best ? c : best;
            }
        }
    }
    return best <= R * R ? best : -1;
}

// T[d2] = the LUT entry of squared distance d2 (d2 <= 2 h^2); *flag = 1 iff the LUT is radial and non-increasing
// over the attainable d2 and h fits the 32-bit window of the nearest-particle kernel
__global__ void __launch_bounds__(1024) lut_radial_kernel(const double* __restrict__ lut, int rfs, double* T,
                                                          int* flag) {
    constexpr int kMaxT = 2 * kNearMaxH * kNearMaxH + 1;
    __shared__ unsigned long long s_bits[kMaxT];
    __shared__ unsigned char s_has[kMaxT];
    __shared__ int s_ok;
    const int h = rfs / 2, n = rfs * rfs, nT = 2 * h * h + 1;
    if (h < 1 || h > kNearMaxH || (rfs & 1) == 0) {
        if (threadIdx.x == 0) *flag = 0;
        return;
    }
    for (int i = threadIdx.x; i < nT; i += blockDim.x) { s_bits[i] = 0ull; s_has[i] = 0; }
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ky = i / rfs, kx = i - ky * rfs;
        const int d2 = (kx - h) * (kx - h) + (ky - h) * (ky - h);
        s_has[d2] = 1;
        s_bits[d2] = (unsigned long long)__double_as_longlong(lut[i]);   // racing writers: checked below
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ky = i / rfs, kx = i - ky * rfs;
        const int d2 = (kx - h) * (kx - h) + (ky - h) * (ky - h);
        if (s_bits[d2] != (unsigned long long)__double_as_longlong(lut[i])) s_ok = 0;
    }
    for (int a = threadIdx.x; a < nT - 1; a += blockDim.x) {
        if (!s_has[a]) continue;
        int b = a + 1;
        while (b < nT && !s_has[b]) ++b;
        if (b < nT && !(__longlong_as_double((long long)s_bits[a]) >= __longlong_as_double((long long)s_bits[b])))
            s_ok = 0;                                   // NaN entries fail here too
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nT; i += blockDim.x)
        T[i] = s_has[i] ? __longlong_as_double((long long)s_bits[i]) : -0.48;
    if (threadIdx.x == 0) *flag = s_ok;
}

// the bitmap rows y0-h .. y0+rows-1+h of the tile, words (x0>>5)-1 .. (x0>>5)+4 of each (bit 0 of the staged row
// is pixel x0-32); rows / words outside the image are zero
__device__ __forceinline__ void stage_occ_rows(const uint32_t* __restrict__ occ, int wpr, int H, int x0, int y_first,
                                               int rows, uint32_t* s_occ) {
    const int gw0 = x0 >> 5;                            // padded row: word (x0>>5)-1 lives at index x0>>5
    for (int i = threadIdx.x; i < rows * kGravOccWords; i += blockDim.x) {
        const int r = i / kGravOccWords, w = i - r * kGravOccWords;
        const int yy = y_first + r, gw = gw0 + w;
        uint32_t v = 0u;
        if (w < 6 && yy >= 0 && yy < H && gw < wpr) v = __ldg(occ + (size_t)yy * wpr + gw);
        s_occ[i] = v;
    }
}

template <int HH>
__global__ void __launch_bounds__(kGravCols) gravity_nearest_kernel(const uint32_t* __restrict__ occ, int wpr,
                                                                   const double* __restrict__ T,
                                                                   const int* __restrict__ flag, int H, int W,
                                                                   double* __restrict__ field) {
    constexpr int TY = kNearRows, SR = TY + 2 * HH, NT = 2 * HH * HH + 1;
    __shared__ uint32_t s_occ[SR * kGravOccWords];
    __shared__ double s_T[NT + 1];
    if (*flag == 0) return;
    const int x0 = blockIdx.x * kGravCols, y0 = blockIdx.y * TY;
    stage_occ_rows(occ, wpr, H, x0, y0 - HH, SR, s_occ);
    for (int i = threadIdx.x; i <= NT; i += blockDim.x) {
        const double t = (i < NT) ? T[i] : -0.48;
        s_T[i] = (t > -0.48) ? t : -0.48;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= W) return;
    uint32_t best[TY / 2];
#pragma unroll
    for (int j = 0; j < TY / 2; ++j) best[j] = kD2Inf * 0x10001u;
    const int b = threadIdx.x + 32 - HH;               // staged bit of pixel x - HH
    const uint32_t* row = s_occ + (b >> 5);
    const int sh = b & 31;
#pragma unroll
    for (int s = 0; s < SR; ++s) {
        // window bits 0 .. 2 HH = pixels x-HH .. x+HH of source row y0 - HH + s
        const uint32_t w = __funnelshift_r(row[s * kGravOccWords], row[s * kGravOccWords + 1], sh);
        const uint32_t wr = (w >> HH) & ((1u << (HH + 1)) - 1u);       // dx = 0 .. HH at bits 0 .. HH
        const uint32_t wl = w & ((1u << HH) - 1u);                     // dx = -HH .. -1 at bits 0 .. HH-1
        const unsigned rr = (unsigned)(__ffs((int)wr) - 1);            // none: 0xffffffff
        const unsigned rl = (unsigned)(HH - 31 + __clz((int)wl));      // none: HH + 1
        const unsigned r = rr < rl ? rr : rl;
        const uint32_t r2 = (r <= (unsigned)HH) ? r * r : kD2Inf;
        const uint32_t v = r2 * 0x10001u;
#pragma unroll
        for (int j = 0; j < TY / 2; ++j) {
            // low half: output row y0 + 2j, high half: y0 + 2j + 1
            const int da = s - HH - 2 * j, db = da - 1;
            const bool ina = da >= -HH && da <= HH, inb = db >= -HH && db <= HH;
            if (ina || inb) {
                const uint32_t c = (ina ? (uint32_t)(da * da) : kD2Inf) | ((inb ? (uint32_t)(db * db) : kD2Inf) << 16);
                best[j] = __viaddmin_u16x2(v, c, best[j]);
            }
        }
    }
    // s_T[NT] = -0.48 serves every "no particle in the window" value (>= kD2Inf)
    double* out = field + (size_t)y0 * W + x;
    if (y0 + TY <= H) {
#pragma unroll
        for (int r = 0; r < TY; ++r) {
            const uint32_t d2 = (best[r >> 1] >> (16 * (r & 1))) & 0xffffu;
            *out = s_T[d2 < (uint32_t)NT ? d2 : (uint32_t)NT];
            out += W;
        }
    } else {
#pragma unroll
        for (int r = 0; r < TY; ++r) {
            if (y0 + r < H) {
                const uint32_t d2 = (best[r >> 1] >> (16 * (r & 1))) & 0xffffu;
                out[(size_t)r * W] = s_T[d2 < (uint32_t)NT ? d2 : (uint32_t)NT];
            }
        }
    }
}

__host__ __device__ inline int grav_bits_rows_stride(int rfs) { return rfs + 2 * (kBitsRows - 1); }   // odd
inline size_t grav_bits_smem(int rfs) {
    return sizeof(double) * (size_t)rfs * grav_bits_rows_stride(rfs) +
           sizeof(uint32_t) * (size_t)(kBitsRows + rfs - 1) * kGravOccWords;
}

constexpr int kBitsTilesPerCta = 4;    // row tiles per CTA of the general kernel (the LUT copy is staged once)

__global__ void __launch_bounds__(kGravCols) gravity_bits_kernel(const uint32_t* __restrict__ occ, int wpr,
                                                                const double* __restrict__ lut, int rfs,
                                                                const int* __restrict__ flag, int H, int W,
                                                                double* __restrict__ field) {
    constexpr int TY = kBitsRows, PAD = kBitsRows - 1;
    extern __shared__ __align__(16) unsigned char s_raw[];
    if (*flag != 0) return;
    const int h = rfs / 2, ROWS = grav_bits_rows_stride(rfs), SR = TY + 2 * h;
    double* s_tab = reinterpret_cast<double*>(s_raw);                  // [rfs][ROWS]: column kx, padded row PAD + ky
    uint32_t* s_occ = reinterpret_cast<uint32_t*>(s_tab + rfs * ROWS);
    for (int i = threadIdx.x; i < rfs * ROWS; i += blockDim.x) {
        const int kx = i / ROWS, ky = i - kx * ROWS - PAD;
        s_tab[i] = (ky >= 0 && ky < rfs) ? lut[ky * rfs + kx] : -INFINITY;
    }
    const int x0 = blockIdx.x * kGravCols, x = x0 + threadIdx.x;
    const int b = threadIdx.x + 32 - h;                // staged bit of pixel x - h
    const uint32_t* row = s_occ + (b >> 5);
    const int sh = b & 31;
    const unsigned long long mask = (1ull << (2 * h + 1)) - 1ull;      // 2h+1 <= 63
    for (int t = 0; t < kBitsTilesPerCta; ++t) {
        const int y0 = (blockIdx.y * kBitsTilesPerCta + t) * TY;
        if (y0 >= H) break;
        __syncthreads();                               // the previous tile's bitmap rows are no longer read
        stage_occ_rows(occ, wpr, H, x0, y0 - h, SR, s_occ);
        __syncthreads();
        if (x >= W) continue;
        double best[TY];
#pragma unroll
        for (int r = 0; r < TY; ++r) best[r] = -INFINITY;
        for (int s = 0; s < SR; ++s) {
            const uint32_t w0 = row[s * kGravOccWords], w1 = row[s * kGravOccWords + 1],
                           w2 = row[s * kGravOccWords + 2];
            unsigned long long win =
                ((unsigned long long)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);
            win &= mask;
            // a particle at bit q sits on pixel x + q - h: kx = x - px + h = 2h - q; ky of output row r = 2h - s + r
            const double* base = s_tab + PAD + 2 * h - s;
            while (win) {
                const int q = __ffsll((long long)win) - 1;
                win &= win - 1ull;
                const double* col = base + (2 * h - q) * ROWS;
#pragma unroll
                for (int r = 0; r < TY; ++r) {
                    const double v = col[r];
                    best[r] = (v > best[r]) ? v : best[r];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < TY; ++r) {
            const int y = y0 + r;
            if (y < H) field[(size_t)y * W + x] = (best[r] > -0.48) ? best[r] : -0.48;
        }
    }
}

// =======================================================================================
// adding particles (utils.py:336-356): first-occurrence row-major argmin per window
// =======================================================================================
__global__ void __launch_bounds__(kThreads) window_argmin_kernel(const double* __restrict__ field, int H, int W,
                                                                int bx, int by, int step, double thr, int nwx,
                                                                int nwy, int32_t* cand, const double* d_thr) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwx * nwy) return;
    if (d_thr) thr = *d_thr;
    int wy = w / nwx, wx = w - wy * nwx;
    int px = bx + wx * step, py = by + wy * step;
    int x1 = min(px + step, W), y1 = min(py + step, H);
    double best = INFINITY;
    int bxy = -1;
    bool first = true;
    for (int y = py; y < y1; ++y)
        for (int x = px; x < x1; ++x) {
            double v = field[(size_t)y * W + x];
            if (first || v < best) { best = v; bxy = y * W + x; first = false; }
        }
    cand[w] = (bxy >= 0 && fabs(best) < thr) ? bxy : -1;
}

struct AppendOp {
    const int32_t* cand;
    int W;
    double2* pos;
    uint8_t* stable;
    int cap;
    int* status;
    __device__ int value(int w) const { return cand[w] >= 0 ? 1 : 0; }
    __device__ void emit(int w, int v, int dst) const {
        if (!v) return;
        if (dst >= cap) { atomicOr(status, AOSLO_DEV_CAPACITY); return; }
        int c = cand[w];
        pos[dst] = make_double2((double)(c % W), (double)(c / W));
        stable[dst] = 0;
    }
};

// =======================================================================================
// distance force (utils.py:112-134 'pit', :43-58 'exp')
// =======================================================================================
// force of one particle from its neighbour list (column 0 = "self" is dropped, utils.py:32)
template <int KMAX>
__device__ __forceinline__ double2 force_from_neighbours(double2 p, const int* nb_idx, int kk,
                                                         const double2* __restrict__ pos, int energy_type,
                                                         double p0, double p1, double p2) {
    double fx = 0.0, fy = 0.0;
#pragma unroll
    for (int j = 1; j < KMAX; ++j) {
        if (j >= kk) break;
        int nb = nb_idx[j];
        double2 q = (nb >= 0 && nb != 0x7fffffff) ? pos[nb] : make_double2(NAN, NAN);
        double vx = p.x - q.x, vy = p.y - q.y;
        double nrm = sqrt(vx * vx + vy * vy);
        double dx, dy;
        if (energy_type == AOSLO_ENERGY_PIT) {
            double f = energy_pit2(nrm, p0, p1, p2);
            dx = (vx / nrm) * f;
            dy = (vy / nrm) * f;
        } else {                                // 'exp': vec / (norm / sigma^2); alpha applied after the sum
            double den = nrm / (p1 * p1);
            dx = vx / den;
            dy = vy / den;
        }
        if (j == 1) { fx = dx; fy = dy; } else { fx += dx; fy += dy; }
    }
    if (energy_type == AOSLO_ENERGY_EXP) { fx = p0 * fx; fy = p0 * fy; }
    return make_double2(fx, fy);
}

__global__ void __launch_bounds__(kThreads) dist_force_kernel(const double2* __restrict__ pos,
                                                              const uint8_t* __restrict__ stable, int n_host,
                                                              const int* d_n, const int32_t* __restrict__ knn_idx,
                                                              int kk /*row stride = k+1*/, int energy_type,
                                                              double p0, double p1, double p2, double2* force) {
    int n = d_n ? *d_n : n_host;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        bool skip = (energy_type == AOSLO_ENERGY_PIT) && stable && stable[i];
        double2 f = make_double2(0.0, 0.0);
        if (!skip) f = force_from_neighbours<AOSLO_KMAX>(pos[i], knn_idx + (size_t)i * kk, kk, pos, energy_type, p0, p1, p2);
        force[i] = f;
    }
}

// =======================================================================================
// move step (pipeline_with_delitions.py:215-245)
// =======================================================================================
__device__ __forceinline__ double np_sign(double v) {
    if (v > 0.0) return 1.0;
    if (v < 0.0) return -1.0;
    return v;   // 0 -> 0, NaN -> NaN
}

// returns false (and flags the status word) when the particle cannot be moved
template <typename T>
__device__ __forceinline__ bool move_one(double2& p, double2 f, const T* __restrict__ grad, int H, int W, double lb,
                                         double ld, int step_mode, int* status) {
    if (p.x != p.x || p.y != p.y) { atomicOr(status, AOSLO_DEV_NAN); return false; }
    int ix = (int)p.x, iy = (int)p.y;
    if (ix < 0 || ix >= W || iy < 0 || iy >= H) {   // reference: IndexError / silent negative wrap
        atomicOr(status, AOSLO_DEV_ESCAPED);
        return false;
    }
    size_t o = ((size_t)iy * fpitch(W) + ix) * 2;
    double g0 = (double)grad[o], g1 = (double)grad[o + 1];
    if (step_mode == AOSLO_STEP_CONTIN) {
        double dy = (-1.0 * lb) * g0 + (-ld) * f.x;
        double dx = (-1.0 * lb) * g1 + ld * f.y;
        dy = -dy;
        p.x += dx;
        p.y += dy;
    } else {
        double delta_y = (-1.0 * lb) * g0 + (-ld) * f.y;
        double delta_x = (-1.0 * lb) * g1 + ld * f.x;
        delta_y = -delta_y;
        double ax = fabs(delta_x), ay = fabs(delta_y);
        if (ax > ay) {
            p.x += np_sign(delta_x);
            if (ax < ay * 2.0) p.y += np_sign(delta_y);
        } else {
            p.y += np_sign(delta_y);
            if (ay < ax * 2.0) p.x += np_sign(delta_x);
        }
    }
    return true;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) move_kernel(double2* pos, const uint8_t* __restrict__ stable,
                                                        int n_host, const int* d_n, const T* __restrict__ grad,
                                                        int H, int W, const double2* __restrict__ force,
                                                        double lb, double ld, int step_mode, int* status) {
    int n = d_n ? *d_n : n_host;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (stable[i]) continue;
        double2 p = pos[i];
        if (move_one<T>(p, force[i], grad, H, W, lb, ld, step_mode, status)) pos[i] = p;
    }
}

// =======================================================================================
// metrics
// =======================================================================================
// Moments of the nearest-distance arrays as plain sums: sum d, sum (d - kShift), sum (d - kShift)^2 with a fixed
// shift near the typical distance (1 px), so that var = (q - l^2 / n) / n does not cancel (worst case, distances of
// tens of pixels: a few digits of 16, tolerance of the parity tests 1e-12).  Every reduction level is three adds per
// array -- the Chan / Welford merge this replaces cost two FP64 divisions per level and made the shuffle tree
// (executed by EVERY thread) four times as long as the queries themselves (ncu: 1 330 instructions per query).
// The tree order is fixed, so the sums are deterministic.
constexpr double kMomentShift = 1.0;
struct Moments {
    double sum, l, q;
    __device__ __forceinline__ void init() { sum = 0.0; l = 0.0; q = 0.0; }
    __device__ __forceinline__ void add(double x) {
        const double d = x - kMomentShift;
        sum += x;
        l += d;
        q += d * d;
    }
    __device__ __forceinline__ void merge(const Moments& o) { sum += o.sum; l += o.l; q += o.q; }
    __device__ __forceinline__ double var(double n) const { return (q - l * l / n) / n; }    // np.var (ddof = 0)
};

__device__ __forceinline__ Moments shfl_down_moments(const Moments& m, int o) {
    Moments r;
    r.sum = __shfl_down_sync(0xffffffffu, m.sum, o);
    r.l = __shfl_down_sync(0xffffffffu, m.l, o);
    r.q = __shfl_down_sync(0xffffffffu, m.q, o);
    return r;
}

struct MetricsPartial {          // one per block
    Moments p2g, g2p;
    double energy;
    double pad_;
};

// L2-coherent read of another block's partial (written before its __threadfence + ticket)
__device__ __forceinline__ MetricsPartial __ldcg_partial(const MetricsPartial* p) {
    const double* d = reinterpret_cast<const double*>(p);
    MetricsPartial r;
    r.p2g.sum = __ldcg(d + 0); r.p2g.l = __ldcg(d + 1); r.p2g.q = __ldcg(d + 2);
    r.g2p.sum = __ldcg(d + 3); r.g2p.l = __ldcg(d + 4); r.g2p.q = __ldcg(d + 5);
    r.energy = __ldcg(d + 6);
    r.pad_ = 0.0;
    return r;
}

// One grid pass over [0, n) particles (nearest GT, in-image test, FP counts, blob energy) and
// [n, n + n_gt) ground-truth points (nearest particle, TP / FN counts).  Integer counts go to
// global atomics (order independent); FP64 moments are merged in a fixed tree per block and
// written as per-block partials for the finalize kernel.
template <typename T>
__global__ void __launch_bounds__(kThreads) metrics_pass_kernel(
    CellView cl_g, LinkView lk_p, const double2* __restrict__ pos, const uint8_t* __restrict__ stable, int n_host,
    const int* d_n, const double2* __restrict__ gt, int n_gt, const T* __restrict__ Rb, int H, int W, int bx, int by,
    double* d_p2g, double* d_g2p, int* icount /*[9]: 8 counts + block ticket*/, MetricsPartial* partials,
    int iteration, aoslo_metrics_record* rec, int* status, const int* d_n_gt, int* d_rec_idx, const int* d_iteration) {
    // graph path: GT count, record slot and iteration number are device scalars (d_rec_idx is advanced by the
    // thread that writes the record)
    const int n = d_n ? *d_n : n_host;
    if (d_n_gt) n_gt = *d_n_gt;
    const int total = n + n_gt;
    int c[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // tp1 tp2 fn1 fn2 fp1 fp2 nin nst
    Moments mp, mg;
    mp.init(); mg.init();
    double energy = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        TopK<1> top;
        if (i < n) {
            double2 p = pos[i];
            knn_cells_query<1>(cl_g, p.x, p.y, top);
            if (top.idx[0] == 0x7fffffff) atomicOr(status, AOSLO_DEV_KNN_SHORT);
            double d = sqrt(top.d2[0]);
            if (d_p2g) d_p2g[i] = d;
            bool inside = ((double)bx < p.x) && (p.x < (double)(W - bx)) && ((double)by < p.y) &&
                          (p.y < (double)(H - by));
            c[6] += inside;
            c[4] += inside && (d > 1.42);
            c[5] += inside && (d > 2.83);
            c[7] += stable[i] ? 1 : 0;
            mp.add(d);
            energy += energy_at<T>(Rb, p, W, H, status);
        } else {
            double2 g = gt[i - n];
            knn_links_query<1>(lk_p, g.x, g.y, top);
            if (top.idx[0] == 0x7fffffff) atomicOr(status, AOSLO_DEV_KNN_SHORT);
            double d = sqrt(top.d2[0]);
            if (d_g2p) d_g2p[i - n] = d;
            c[0] += (d <= 1.42); c[2] += (d > 1.42);
            c[1] += (d <= 2.83); c[3] += (d > 2.83);
            mg.add(d);
        }
    }
    // ---- block reduction
    __shared__ MetricsPartial s_part[kThreads / 32];
    __shared__ int s_c[8];
    if (threadIdx.x < 8) s_c[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int v = c[j];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_c[j], v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        Moments a = shfl_down_moments(mp, o), b = shfl_down_moments(mg, o);
        double e = __shfl_down_sync(0xffffffffu, energy, o);
        mp.merge(a); mg.merge(b); energy += e;
    }
    if ((threadIdx.x & 31) == 0) {
        s_part[threadIdx.x >> 5].p2g = mp;
        s_part[threadIdx.x >> 5].g2p = mg;
        s_part[threadIdx.x >> 5].energy = energy;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        MetricsPartial acc = s_part[0];
        for (int w = 1; w < kThreads / 32; ++w) {
            acc.p2g.merge(s_part[w].p2g);
            acc.g2p.merge(s_part[w].g2p);
            acc.energy += s_part[w].energy;
        }
        partials[blockIdx.x] = acc;
    }
    if (threadIdx.x < 8 && s_c[threadIdx.x]) atomicAdd(&icount[threadIdx.x], s_c[threadIdx.x]);
    // ---- the last block to finish merges the per-block partials (fixed order: independent of WHICH block
    // is last) and writes the record -- no separate finalize launch
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&icount[8], 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    MetricsPartial acc;
    acc.p2g.init(); acc.g2p.init(); acc.energy = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
        const MetricsPartial pb = __ldcg_partial(partials + b);
        acc.p2g.merge(pb.p2g);
        acc.g2p.merge(pb.g2p);
        acc.energy += pb.energy;
    }
    for (int o = 16; o > 0; o >>= 1) {
        Moments a = shfl_down_moments(acc.p2g, o), g = shfl_down_moments(acc.g2p, o);
        double e = __shfl_down_sync(0xffffffffu, acc.energy, o);
        acc.p2g.merge(a); acc.g2p.merge(g); acc.energy += e;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        if (threadIdx.x < kThreads / 32) acc = s_part[threadIdx.x];
        else { acc.p2g.init(); acc.g2p.init(); acc.energy = 0.0; }
        for (int o = 16; o > 0; o >>= 1) {
            Moments a = shfl_down_moments(acc.p2g, o), g = shfl_down_moments(acc.g2p, o);
            double e = __shfl_down_sync(0xffffffffu, acc.energy, o);
            acc.p2g.merge(a); acc.g2p.merge(g); acc.energy += e;
        }
    }
    if (threadIdx.x != 0) return;
    if (d_rec_idx) { const int slot = *d_rec_idx; rec += slot; *d_rec_idx = slot + 1; }
    if (d_iteration) iteration = *d_iteration;
    rec->iteration = iteration;
    rec->n_particles = n;
    rec->n_stable = __ldcg(&icount[7]);
    rec->n_gt = n_gt;
    rec->tp[0] = __ldcg(&icount[0]); rec->tp[1] = __ldcg(&icount[1]);
    rec->fn[0] = __ldcg(&icount[2]); rec->fn[1] = __ldcg(&icount[3]);
    rec->fp[0] = __ldcg(&icount[4]); rec->fp[1] = __ldcg(&icount[5]);
    rec->n_in_image = __ldcg(&icount[6]);
    rec->pad_ = 0;
    rec->sum_p2g = acc.p2g.sum; rec->sum_g2p = acc.g2p.sum;
    rec->var_p2g = acc.p2g.var((double)n); rec->var_g2p = acc.g2p.var((double)n_gt);
    rec->blob_energy_sum = acc.energy;
}

// =======================================================================================
// The run loop's particle structure (aoslo_run, canonical tie order): tombstones + two cell structures
// =======================================================================================
// Between two resamples only the non-stable particles move (:216) -- a few thousand of ~145 k on the benchmark
// frame -- yet rebuilding the cell structure and compacting the arrays after every deletion touches all of them,
// twice per iteration.  So inside a resample period
//   * deleted particles stay in the arrays as TOMBSTONES (del[i] = 1, position NaN): indices do not shift, the
//     relative order of the survivors -- all that the (d^2, index) neighbour order and the row-order deletion rule
//     depend on -- is unchanged, and the order-preserving compaction runs once per period (before the resample,
//     and once at the end of the run);
//   * the STABLE particles are linked once per period (structure S); only the NON-STABLE ones are re-linked after
//     they move (structure N, a few thousand entries, one small kernel).  A query walks both chains of a cell.
//     A tombstone needs no test in the candidate loops: its NaN distance never enters a top-k list.
struct DualView {
    const int* geom;                 // shift, gw, gh, ncell, tag of S, tag of N
    const unsigned long long* headS;
    const unsigned long long* headN;
    const int* next;                 // shared: a particle is in exactly one structure
    const double2* pos;
};

template <int K>
__device__ __forceinline__ void knn_dual_cell(const DualView& dv, int cell, unsigned int tagS, unsigned int tagN,
                                              double qx, double qy, TopK<K>& top) {
    const unsigned long long hs = dv.headS[cell], hn = dv.headN[cell];
    for (int p = ((unsigned int)(hs >> 32) == tagS) ? (int)(unsigned int)hs : -1; p >= 0; p = dv.next[p]) {
        const double2 t = dv.pos[p];
        const double dx = qx - t.x, dy = qy - t.y;
        top.push(dx * dx + dy * dy, p);
    }
    for (int p = ((unsigned int)(hn >> 32) == tagN) ? (int)(unsigned int)hn : -1; p >= 0; p = dv.next[p]) {
        const double2 t = dv.pos[p];
        const double dx = qx - t.x, dy = qy - t.y;
        top.push(dx * dx + dy * dy, p);
    }
}

// thread-per-query ring search over S and N (same stopping rule as knn_links_query)
template <int K>
__device__ __forceinline__ void knn_dual_query(const DualView& dv, double qx, double qy, TopK<K>& top) {
    top.init();
    const int shift = dv.geom[0], gw = dv.geom[1], gh = dv.geom[2];
    const unsigned int tagS = (unsigned int)dv.geom[4], tagN = (unsigned int)dv.geom[5];
    const int c = 1 << shift;
    int cx = (qx >= 0.0) ? ((int)fmin(qx, 1.0e9) >> shift) : 0;
    int cy = (qy >= 0.0) ? ((int)fmin(qy, 1.0e9) >> shift) : 0;
    cx = min(cx, gw - 1);
    cy = min(cy, gh - 1);
    const int rmax = max(gw, gh);
    for (int r = 0; r <= rmax; ++r) {
        const int ylo = cy - r, yhi = cy + r, xlo = cx - r, xhi = cx + r;
        for (int y = max(ylo, 0); y <= min(yhi, gh - 1); ++y) {
            if (y == ylo || y == yhi) {
                for (int x = max(xlo, 0); x <= min(xhi, gw - 1); ++x)
                    knn_dual_cell<K>(dv, y * gw + x, tagS, tagN, qx, qy, top);
            } else {
                if (xlo >= 0) knn_dual_cell<K>(dv, y * gw + xlo, tagS, tagN, qx, qy, top);
                if (xhi < gw) knn_dual_cell<K>(dv, y * gw + xhi, tagS, tagN, qx, qy, top);
            }
        }
        double bound = INFINITY;
        if (xlo > 0) bound = fmin(bound, qx - (double)(xlo * c));
        if (ylo > 0) bound = fmin(bound, qy - (double)(ylo * c));
        if (xhi < gw - 1) bound = fmin(bound, (double)((xhi + 1) * c) - qx);
        if (yhi < gh - 1) bound = fmin(bound, (double)((yhi + 1) * c) - qy);
        if (bound == INFINITY) break;
        if (bound > 0.0 && top.d2[K - 1] < bound * bound) break;
    }
}

// warp-per-query variant (lanes split the cells; see knn_links_query_warp).  Rings 0-2 -- the 5 x 5 block, one cell
// per lane -- are visited TOGETHER: with the 4 px cells of the steady state the 4th neighbour (~1 cone pitch away)
// lies beyond the 3 x 3 block's guaranteed radius, so ring 2 was needed for nearly every query and cost a second
// round of dependent loads + a second merge; the extra cells are free (idle lanes).  Further rings one at a time.
template <int K>
__device__ __forceinline__ void knn_dual_query_warp(const DualView& dv, double qx, double qy, TopK<K>& out) {
    const int lane = threadIdx.x & 31;
    TopK<K> mine;
    mine.init();
    out.init();
    const int shift = dv.geom[0], gw = dv.geom[1], gh = dv.geom[2];
    const unsigned int tagS = (unsigned int)dv.geom[4], tagN = (unsigned int)dv.geom[5];
    const int c = 1 << shift;
    int cx = (qx >= 0.0) ? ((int)fmin(qx, 1.0e9) >> shift) : 0;
    int cy = (qy >= 0.0) ? ((int)fmin(qy, 1.0e9) >> shift) : 0;
    cx = min(cx, gw - 1);
    cy = min(cy, gh - 1);
    const int rmax = max(gw, gh);
    for (int r = 2; r <= rmax; ++r) {
        if (r == 2) {
            if (lane < 25) {
                const int x = cx + (lane % 5) - 2, y = cy + (lane / 5) - 2;
                if (x >= 0 && x < gw && y >= 0 && y < gh) knn_dual_cell<K>(dv, y * gw + x, tagS, tagN, qx, qy, mine);
            }
        } else {
            const int ncells = 8 * r;
            for (int j = lane; j < ncells; j += 32) {
                int dx, dy;
                if (j <= 2 * r) { dx = -r + j; dy = -r; }
                else if (j <= 4 * r + 1) { dx = -r + (j - 2 * r - 1); dy = r; }
                else if (j <= 6 * r) { dx = -r; dy = -r + 1 + (j - 4 * r - 2); }
                else { dx = r; dy = -r + 1 + (j - 6 * r - 1); }
                const int x = cx + dx, y = cy + dy;
                if (x >= 0 && x < gw && y >= 0 && y < gh) knn_dual_cell<K>(dv, y * gw + x, tagS, tagN, qx, qy, mine);
            }
        }
        warp_merge_topk<K>(mine, out);
        const int ylo = cy - r, yhi = cy + r, xlo = cx - r, xhi = cx + r;
        double bound = INFINITY;
        if (xlo > 0) bound = fmin(bound, qx - (double)(xlo * c));
        if (ylo > 0) bound = fmin(bound, qy - (double)(ylo * c));
        if (xhi < gw - 1) bound = fmin(bound, (double)((xhi + 1) * c) - qx);
        if (yhi < gh - 1) bound = fmin(bound, (double)((yhi + 1) * c) - qy);
        if (bound == INFINITY) break;
        if (bound > 0.0 && out.d2[K - 1] < bound * bound) break;
    }
}

__device__ __forceinline__ void link_push(unsigned long long* head, int* next, int cell, unsigned int tag, int i) {
    const unsigned long long old =
        atomicExch(&head[cell], ((unsigned long long)tag << 32) | (unsigned long long)(unsigned int)i);
    next[i] = ((unsigned int)(old >> 32) == tag) ? (int)(unsigned int)old : -1;
}

// Start of a resample period: clears the tombstones, links the stable particles into S and the non-stable ones into
// N (+ the non-stable list), fixes the geometry from the count.  The last block publishes geometry and both tags.
__global__ void __launch_bounds__(kThreads) period_begin_kernel(
    const double2* __restrict__ pos, const uint8_t* __restrict__ stable, int n_host, const int* d_n, int W, int H,
    int min_shift, int fixed_shift, int bias_x4, int* geom, unsigned int* epochS, unsigned int* epochN, int* ticket,
    unsigned long long* headS, unsigned long long* headN, int* next, uint8_t* del, int* ns_list, int* n_ns, int* n_dead,
    int* status) {
    const int n = d_n ? *d_n : n_host;
    const int shift = link_shift_for(n, H, W, min_shift, fixed_shift, bias_x4);
    const int gw = (W + (1 << shift) - 1) >> shift, gh = (H + (1 << shift) - 1) >> shift;
    const unsigned int tagS = *epochS + 1u, tagN = *epochN + 1u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double2 p = pos[i];
        const uint8_t st = stable[i];
        del[i] = 0;
        bool bad = false;
        const int cx = cell_coord(p.x, shift, gw, bad);
        const int cy = cell_coord(p.y, shift, gh, bad);
        if (p.x != p.x || p.y != p.y) atomicOr(status, AOSLO_DEV_NAN);
        else if (bad || p.x >= (double)W || p.y >= (double)H) atomicOr(status, AOSLO_DEV_ESCAPED);
        if (st) {
            link_push(headS, next, cy * gw + cx, tagS, i);
        } else {
            link_push(headN, next, cy * gw + cx, tagN, i);
            ns_list[atomicAdd(n_ns, 1)] = i;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
            geom[0] = shift; geom[1] = gw; geom[2] = gh; geom[3] = gw * gh; geom[4] = (int)tagS; geom[5] = (int)tagN;
            *epochS = tagS;
            *epochN = tagN;
            *n_dead = 0;
            *ticket = 0;
        }
    }
}

// After the move: the non-stable survivors take their moved positions (Jacobi: every force was computed from the
// old ones) and are linked into a fresh N structure.
__global__ void __launch_bounds__(kThreads) ns_relink_kernel(const int* __restrict__ ns_list,
                                                             const int* __restrict__ n_ns,
                                                             const double2* __restrict__ moved, double2* pos,
                                                             const uint8_t* __restrict__ del, int W, int H, int* geom,
                                                             unsigned int* epochN, int* ticket,
                                                             unsigned long long* headN, int* next, int* status) {
    const int total = *n_ns;
    const int shift = geom[0], gw = geom[1], gh = geom[2];
    const unsigned int tagN = *epochN + 1u;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const int i = ns_list[q];
        if (del[i]) continue;
        const double2 p = moved[i];
        pos[i] = p;
        bool bad = false;
        const int cx = cell_coord(p.x, shift, gw, bad);
        const int cy = cell_coord(p.y, shift, gh, bad);
        if (p.x != p.x || p.y != p.y) atomicOr(status, AOSLO_DEV_NAN);
        else if (bad || p.x >= (double)W || p.y >= (double)H) atomicOr(status, AOSLO_DEV_ESCAPED);
        link_push(headN, next, cy * gw + cx, tagN, i);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
            geom[5] = (int)tagN;
            *epochN = tagN;
            *ticket = 0;
        }
    }
}

// force + move of the non-stable survivors (one warp per particle; utils.py:112-134 / :43-58 and
// pipeline_with_delitions.py:215-245), moved positions into `moved[i]`
template <int K, typename T>
__global__ void __launch_bounds__(kThreads) dual_force_move_kernel(
    DualView dv, const int* __restrict__ ns_list, const int* __restrict__ n_ns, const uint8_t* __restrict__ del,
    int kk, int energy_type, double p0, double p1, double p2, const T* __restrict__ grad, int H, int W, double lb,
    double ld, int step_mode, double2* __restrict__ moved, int* status, const RunParams* rp) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int total = *n_ns;
    if (rp) { kk = rp->kk; p0 = rp->e0; p1 = rp->e1; p2 = rp->e2; lb = rp->lb; ld = rp->ld; }
    for (int q = warp; q < total; q += nwarps) {
        const int i = ns_list[q];
        if (del[i]) continue;
        double2 p = dv.pos[i];
        TopK<K> top;
        knn_dual_query_warp<K>(dv, p.x, p.y, top);
        int nb = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (lane == j) nb = top.idx[j];
        double tx = 0.0, ty = 0.0;
        if (lane >= 1 && lane < kk) {
            double2 qn = (nb != 0x7fffffff) ? dv.pos[nb] : make_double2(NAN, NAN);
            double vx = p.x - qn.x, vy = p.y - qn.y;
            double nrm = sqrt(vx * vx + vy * vy);
            if (energy_type == AOSLO_ENERGY_PIT) {
                double f = energy_pit2(nrm, p0, p1, p2);
                tx = (vx / nrm) * f;
                ty = (vy / nrm) * f;
            } else {
                double den = nrm / (p1 * p1);
                tx = vx / den;
                ty = vy / den;
            }
        }
        double fx = __shfl_sync(0xffffffffu, tx, 1), fy = __shfl_sync(0xffffffffu, ty, 1);
        for (int j = 2; j < kk; ++j) {
            fx += __shfl_sync(0xffffffffu, tx, j);
            fy += __shfl_sync(0xffffffffu, ty, j);
        }
        if (energy_type == AOSLO_ENERGY_EXP) { fx = p0 * fx; fy = p0 * fy; }
        if (lane == 0) {
            bool shortfall = false;
#pragma unroll
            for (int j = 0; j < K; ++j)
                if (j == kk - 1 && top.idx[j] == 0x7fffffff) shortfall = true;
            if (shortfall) atomicOr(status, AOSLO_DEV_KNN_SHORT);
            if (step_mode >= 0) move_one<T>(p, make_double2(fx, fy), grad, H, W, lb, ld, step_mode, status);
            moved[i] = p;
        }
    }
}

// Wavefront resolution of the active rows by ONE block (see delete_resolve_small_kernel); a victim becomes a
// tombstone: del = 1, position NaN (it drops out of every later neighbour search by itself).  Called by the last
// block of dual_rows_kernel.
template <typename T>
__device__ void dual_resolve_block(double2* pos, const int32_t* knn_idx, const int* active, int* n_active,
                                   unsigned int* d_epoch, int W, int H, const T* __restrict__ Rb,
                                   unsigned long long* touch, uint8_t* del, uint8_t* astate, int* n_dead, int* status) {
    __shared__ int s_pending, s_deleted;
    const unsigned int epoch = *d_epoch;
    const int na = __ldcg(n_active);
    if (threadIdx.x == 0) s_deleted = 0;
    for (int a = threadIdx.x; a < na; a += blockDim.x) astate[a] = 1;
    __syncthreads();
    for (unsigned int round = 1; na > 0; ++round) {
        const unsigned long long hi = ((unsigned long long)epoch << 44) | ((unsigned long long)round << 32);
        for (int a = threadIdx.x; a < na; a += blockDim.x) {
            if (!astate[a]) continue;
            const int i = __ldcg(&active[a]);
            const unsigned long long key = hi | (unsigned long long)(0xffffffffu - (unsigned)i);
            atomicMax(&touch[__ldcg(&knn_idx[2 * (size_t)i])], key);
            atomicMax(&touch[__ldcg(&knn_idx[2 * (size_t)i + 1])], key);
        }
        if (threadIdx.x == 0) s_pending = 0;
        __threadfence();
        __syncthreads();
        int pending = 0;
        for (int a = threadIdx.x; a < na; a += blockDim.x) {
            if (!astate[a]) continue;
            const int i = __ldcg(&active[a]);
            const int from = __ldcg(&knn_idx[2 * (size_t)i]), to = __ldcg(&knn_idx[2 * (size_t)i + 1]);
            const unsigned long long key = hi | (unsigned long long)(0xffffffffu - (unsigned)i);
            if (__ldcg(&touch[from]) == key && __ldcg(&touch[to]) == key) {
                if (!__ldcg(&del[from]) && !__ldcg(&del[to])) {
                    const double2 pf = make_double2(__ldcg(&pos[from].x), __ldcg(&pos[from].y));
                    const double2 pt = make_double2(__ldcg(&pos[to].x), __ldcg(&pos[to].y));
                    const double e1 = energy_at<T>(Rb, pf, W, H, status);
                    const double e2 = energy_at<T>(Rb, pt, W, H, status);
                    const int victim = (e1 < e2) ? from : to;
                    __stcg(&del[victim], (uint8_t)1);
                    pos[victim] = make_double2(NAN, NAN);
                    atomicAdd(&s_deleted, 1);
                }
                astate[a] = 0;
            } else {
                ++pending;
            }
        }
        if (pending) atomicAdd(&s_pending, pending);
        __threadfence();
        __syncthreads();
        const int left = s_pending;
        __syncthreads();
        if (left == 0) break;
        if (round >= 4000u) { if (threadIdx.x == 0) atomicOr(status, AOSLO_DEV_CAPACITY); break; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {                   // ready for the next call
        *d_epoch = epoch + 1;
        *n_dead += s_deleted;
        *n_active = 0;
    }
}

// Deletion of one iteration (utils.py:245-278 on the tombstone structure): rows of the non-stable survivors
// (kNN(2) at the moved positions) -> compact active list -> the LAST block to finish resolves it (wavefront rounds,
// one block; no second launch).
//
// Why the rows of the non-stable particles suffice.  A row is active only if its `from` -- the first entry of its
// (d^2, index)-ordered neighbour list -- is non-stable and its neighbour is closer than the threshold.  `from` is the
// row's own particle unless other particles sit at exactly the same position; then EVERY row of that coincident group
// has the same two endpoints, the group's two lowest indices a < b, and only row a can act: the reference walks the
// rows in index order, row a comes first and either deletes one of {a, b} or finds one of them already deleted, so the
// later rows of the group find an endpoint deleted and do nothing.  Row a is active iff a is non-stable -- and then a
// is in the non-stable list and `from == a` in its own row.  So: rows whose `from` is another particle are skipped,
// stable rows never need to be looked at, and the result equals the sequential walk over all rows.
template <typename T>
__global__ void __launch_bounds__(kThreads) dual_rows_kernel(DualView dv, double2* pos,
                                                             const uint8_t* __restrict__ stable, uint8_t* del,
                                                             const int* __restrict__ ns_list,
                                                             const int* __restrict__ n_ns, double thr,
                                                             const double* d_thr, int32_t* knn_idx, double* knn_d2,
                                                             int* active, int* n_active, int* ticket,
                                                             unsigned int* d_epoch, int W, int H,
                                                             const T* __restrict__ Rb, unsigned long long* touch,
                                                             uint8_t* astate, int* n_dead, int* status) {
    const int lane = threadIdx.x & 31;
    if (d_thr) thr = *d_thr;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int total = *n_ns;
    for (int q = warp; q < total; q += nwarps) {
        const int i = ns_list[q];
        if (del[i]) continue;
        double2 p = dv.pos[i];
        TopK<2> top;
        knn_dual_query_warp<2>(dv, p.x, p.y, top);
        if (lane == 0) {
            const bool ok = top.idx[1] != 0x7fffffff;
            const int from = top.idx[0], to = ok ? top.idx[1] : -1;
            if (!ok) atomicOr(status, AOSLO_DEV_KNN_SHORT);
            knn_idx[2 * (size_t)i] = from;
            knn_idx[2 * (size_t)i + 1] = to;
            knn_d2[2 * (size_t)i] = top.d2[0];
            knn_d2[2 * (size_t)i + 1] = top.d2[1];
            if (ok && from == i && !stable[from] && sqrt(top.d2[1]) < thr) active[atomicAdd(n_active, 1)] = i;
        }
    }
    __shared__ int s_last;
    __threadfence();                           // this block's rows / list entries are visible before its ticket
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) *ticket = 0;
    dual_resolve_block<T>(pos, knn_idx, active, n_active, d_epoch, W, H, Rb, touch, del, astate, n_dead, status);
}

struct KeepAliveOp {                           // order-preserving compaction of the survivors
    const double2* pos;
    const uint8_t* stable;
    const uint8_t* del;
    double2* pos_out;
    uint8_t* stable_out;
    __device__ int value(int i) const { return del[i] ? 0 : 1; }
    __device__ void emit(int i, int v, int dst) const {
        if (!v) return;
        pos_out[dst] = pos[i];
        stable_out[dst] = stable[i];
    }
};

__global__ void __launch_bounds__(kThreads) copy_particles_kernel(const double2* __restrict__ pos_in,
                                                                  const uint8_t* __restrict__ st_in, const int* cnt_in,
                                                                  double2* __restrict__ pos_out,
                                                                  uint8_t* __restrict__ st_out, int* cnt_out,
                                                                  int* n_dead) {
    const int n = *cnt_in;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        pos_out[i] = pos_in[i];
        st_out[i] = st_in[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *cnt_out = n;                              // no other thread of this kernel reads it
        if (n_dead) *n_dead = 0;
    }
}

// metrics over the survivors (see metrics_pass_kernel); GT -> nearest particle searches S and N
template <typename T>
__global__ void __launch_bounds__(kThreads) dual_metrics_kernel(
    CellView cl_g, DualView dv, const uint8_t* __restrict__ stable, const uint8_t* __restrict__ del, int n_host,
    const int* d_n, const int* d_n_dead, const double2* __restrict__ gt, int n_gt, const T* __restrict__ Rb, int H, int W,
    int bx, int by, int* icount /*[9]: 8 counts + block ticket*/, MetricsPartial* partials, int iteration,
    aoslo_metrics_record* rec, int* status, const int* d_n_gt, int* d_rec_idx, const int* d_iteration,
    const uint32_t* __restrict__ occ_p, const uint32_t* __restrict__ occ_g, int wpr, const int* __restrict__ occ_ok) {
    const int n = d_n ? *d_n : n_host;
    if (d_n_gt) n_gt = *d_n_gt;
    const int total = n + n_gt;
    // nearest distances between integer pixels come from the occupancy bitmaps (occ_nearest_d2); a set that is not
    // all integer pixels (occ_ok = 0), a query that is not one, or a neighbour beyond 15 px takes the cell search
    const bool bits_p = occ_ok[0] != 0, bits_g = occ_ok[1] != 0;
    int c[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // tp1 tp2 fn1 fn2 fp1 fp2 nin nst
    Moments mp, mg;
    mp.init(); mg.init();
    double energy = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        TopK<1> top;
        if (i < n) {
            if (del[i]) continue;
            const double2 p = dv.pos[i];
            int q2 = -1;
            if (bits_g) {
                const int ix = (int)p.x, iy = (int)p.y;
                if ((double)ix == p.x && (double)iy == p.y && ix >= 0 && ix < W && iy >= 0 && iy < H)
                    q2 = occ_nearest_d2(occ_g, wpr, H, ix, iy);
            }
            double d;
            if (q2 >= 0) {
                d = sqrt((double)q2);
            } else {
                knn_cells_query<1>(cl_g, p.x, p.y, top);
                if (top.idx[0] == 0x7fffffff) atomicOr(status, AOSLO_DEV_KNN_SHORT);
                d = sqrt(top.d2[0]);
            }
            const bool inside = ((double)bx < p.x) && (p.x < (double)(W - bx)) && ((double)by < p.y) &&
                                (p.y < (double)(H - by));
            c[6] += inside;
            c[4] += inside && (d > 1.42);
            c[5] += inside && (d > 2.83);
            c[7] += stable[i] ? 1 : 0;
            mp.add(d);
            energy += energy_at<T>(Rb, p, W, H, status);
        } else {
            const double2 g = gt[i - n];
            int q2 = -1;
            if (bits_p) {
                const int ix = (int)g.x, iy = (int)g.y;
                if ((double)ix == g.x && (double)iy == g.y && ix >= 0 && ix < W && iy >= 0 && iy < H)
                    q2 = occ_nearest_d2(occ_p, wpr, H, ix, iy);
            }
            double d;
            if (q2 >= 0) {
                d = sqrt((double)q2);
            } else {
                knn_dual_query<1>(dv, g.x, g.y, top);
                if (top.idx[0] == 0x7fffffff) atomicOr(status, AOSLO_DEV_KNN_SHORT);
                d = sqrt(top.d2[0]);
            }
            c[0] += (d <= 1.42); c[2] += (d > 1.42);
            c[1] += (d <= 2.83); c[3] += (d > 2.83);
            mg.add(d);
        }
    }
    __shared__ MetricsPartial s_part[kThreads / 32];
    __shared__ int s_c[8];
    if (threadIdx.x < 8) s_c[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int v = c[j];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_c[j], v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        Moments a = shfl_down_moments(mp, o), b = shfl_down_moments(mg, o);
        double e = __shfl_down_sync(0xffffffffu, energy, o);
        mp.merge(a); mg.merge(b); energy += e;
    }
    if ((threadIdx.x & 31) == 0) {
        s_part[threadIdx.x >> 5].p2g = mp;
        s_part[threadIdx.x >> 5].g2p = mg;
        s_part[threadIdx.x >> 5].energy = energy;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        MetricsPartial acc = s_part[0];
        for (int w = 1; w < kThreads / 32; ++w) {
            acc.p2g.merge(s_part[w].p2g);
            acc.g2p.merge(s_part[w].g2p);
            acc.energy += s_part[w].energy;
        }
        acc.pad_ = 0.0;
        partials[blockIdx.x] = acc;
    }
    if (threadIdx.x < 8 && s_c[threadIdx.x]) atomicAdd(&icount[threadIdx.x], s_c[threadIdx.x]);
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&icount[8], 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    MetricsPartial acc;
    acc.p2g.init(); acc.g2p.init(); acc.energy = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
        const MetricsPartial pb = __ldcg_partial(partials + b);
        acc.p2g.merge(pb.p2g);
        acc.g2p.merge(pb.g2p);
        acc.energy += pb.energy;
    }
    for (int o = 16; o > 0; o >>= 1) {
        Moments a = shfl_down_moments(acc.p2g, o), g = shfl_down_moments(acc.g2p, o);
        double e = __shfl_down_sync(0xffffffffu, acc.energy, o);
        acc.p2g.merge(a); acc.g2p.merge(g); acc.energy += e;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        if (threadIdx.x < kThreads / 32) acc = s_part[threadIdx.x];
        else { acc.p2g.init(); acc.g2p.init(); acc.energy = 0.0; }
        for (int o = 16; o > 0; o >>= 1) {
            Moments a = shfl_down_moments(acc.p2g, o), g = shfl_down_moments(acc.g2p, o);
            double e = __shfl_down_sync(0xffffffffu, acc.energy, o);
            acc.p2g.merge(a); acc.g2p.merge(g); acc.energy += e;
        }
    }
    if (threadIdx.x != 0) return;
    const int n_alive = n - *d_n_dead;
    if (d_rec_idx) { const int slot = *d_rec_idx; rec += slot; *d_rec_idx = slot + 1; }
    if (d_iteration) iteration = *d_iteration;
    rec->iteration = iteration;
    rec->n_particles = n_alive;
    rec->n_stable = __ldcg(&icount[7]);
    rec->n_gt = n_gt;
    rec->tp[0] = __ldcg(&icount[0]); rec->tp[1] = __ldcg(&icount[1]);
    rec->fn[0] = __ldcg(&icount[2]); rec->fn[1] = __ldcg(&icount[3]);
    rec->fp[0] = __ldcg(&icount[4]); rec->fp[1] = __ldcg(&icount[5]);
    rec->n_in_image = __ldcg(&icount[6]);
    rec->pad_ = 0;
    rec->sum_p2g = acc.p2g.sum; rec->sum_g2p = acc.g2p.sum;
    rec->var_p2g = acc.p2g.var((double)n_alive); rec->var_g2p = acc.g2p.var((double)n_gt);
    rec->blob_energy_sum = acc.energy;
}

__global__ void __launch_bounds__(kThreads) dual_count_stable_kernel(const uint8_t* __restrict__ stable,
                                                                     const uint8_t* __restrict__ del, int n_host,
                                                                     const int* d_n, int* out) {
    const int n = d_n ? *d_n : n_host;
    int c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        c += (stable[i] && !del[i]) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

__global__ void set_all_stable_kernel(uint8_t* stable, int n_host, const int* d_n, uint8_t v) {
    int n = d_n ? *d_n : n_host;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) stable[i] = v;
}

// count stable flags into *out (zeroed by the caller); used by the early-stop test (:381)
__global__ void __launch_bounds__(kThreads) count_stable_kernel(const uint8_t* __restrict__ stable, int n_host,
                                                                const int* d_n, int* out) {
    int n = d_n ? *d_n : n_host;
    int c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += stable[i] ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// =======================================================================================
// host-side stage drivers (stream ordered; used by the C ABI wrappers and by aoslo_run)
// =======================================================================================
// d_n (optional): device-resident point/query count of a self-query (canonical mode only; the
// sklearn-order path needs the count on the host to build the tree)
int stage_knn(aoslo_ctx* ctx, cudaStream_t s, const double* d_points, int n_points, const double* d_query,
              int n_query, int k, int tie_mode, int32_t* d_idx, double* d_d2, bool tree_is_particles,
              const int* d_n) {
    if (tie_mode == AOSLO_TIE_SKLEARN) {
        if (d_n) return AOSLO_ERR_INVALID;
        return knn_sklearn(ctx, s, d_points, n_points, d_query, n_query, k, d_idx, d_d2);
    }
    (void)tree_is_particles;
    if (n_points > ctx->lk_p.cap) return AOSLO_ERR_CAPACITY;
    AOSLO_TRY(build_links(ctx, s, ctx->lk_p, d_points, n_points, d_n, kLinkBiasDense));
    return knn_links(ctx, s, link_view(ctx, ctx->lk_p, d_points), d_query, n_query, d_n, k, d_idx, d_d2);
}

// The 20-bit epoch field of the touch keys must never wrap inside a run: called at the start of a run
// (stream ordered) when the host-side estimate gets close; `calls_ahead` = resolve calls the run may make.
int epoch_guard(aoslo_ctx* ctx, cudaStream_t s, unsigned int calls_ahead) {
    if (ctx->epoch + calls_ahead + 16u < (1u << 20)) return AOSLO_OK;
    AOSLO_CUDA(cudaMemsetAsync(ctx->d_touch, 0, sizeof(unsigned long long) * ctx->cap, s));
    const unsigned int one = 1u;
    AOSLO_CUDA(cudaMemcpyAsync(ctx->d_epoch, &one, sizeof(one), cudaMemcpyHostToDevice, s));
    AOSLO_CUDA(cudaStreamSynchronize(s));
    ctx->epoch = 1;
    return AOSLO_OK;
}

// small = true: in-loop variant (compact active list + one resolving block, no cooperative launch)
template <typename T>
int stage_delete_t(aoslo_ctx* ctx, cudaStream_t s, const double* d_pos, const uint8_t* d_stable, int n,
                   const T* d_Rb, int H, int W, double thr, int tie_mode, double* d_pos_out, uint8_t* d_stable_out,
                   uint8_t* d_deleted_out, int* d_n_out, int* d_rounds_out, const int* d_n, int mode,
                   const double* d_thr, int scan_hint) {
    // mode 0: cooperative all-rows resolve; 1: compact active list + one resolving block (the sklearn-order run loop)
    if (n > ctx->cap) return AOSLO_ERR_CAPACITY;
    if (n < 2) return AOSLO_ERR_INVALID;   // sklearn: n_neighbors=2 > n_samples
    if (d_n && d_n == d_n_out) return AOSLO_ERR_INVALID;   // the count is double-buffered
    int* n_active = ctx->d_counters + 12;
    const bool small = mode == 1;
    if (small) {
        if (tie_mode == AOSLO_TIE_CANONICAL) {
            AOSLO_TRY(build_links(ctx, s, ctx->lk_p, d_pos, n, d_n, kLinkBiasSteady));
            knn2_rows_kernel<<<ctx->nb, kThreads, 0, s>>>(link_view(ctx, ctx->lk_p, d_pos), d_stable, n, d_n, thr, d_thr,
                                                          ctx->d_knn_idx, ctx->d_knn_d2, ctx->d_del, ctx->d_active,
                                                          n_active, ctx->d_status);
            AOSLO_LAUNCH_CHECK();
        } else {
            AOSLO_TRY(stage_knn(ctx, s, d_pos, n, d_pos, n, 2, tie_mode, ctx->d_knn_idx, ctx->d_knn_d2, true, d_n));
            collect_active_kernel<<<ctx->nb, kThreads, 0, s>>>(ctx->d_knn_idx, ctx->d_knn_d2, d_stable, n, d_n, thr,
                                                               ctx->d_del, ctx->d_active, n_active);
            AOSLO_LAUNCH_CHECK();
        }
        // the resolving block also prepares the compaction offsets, so only the emit phase of the
        // ordered scan has to be launched afterwards
        // the number of compaction blocks is a launch-size choice only (any count is correct for any n)
        const int nb_scan = scan_blocks_for(scan_hint > 0 ? scan_hint : n);
        delete_resolve_small_kernel<T><<<1, kScanThreads, 0, s>>>(
            (const double2*)d_pos, ctx->d_knn_idx, ctx->d_knn_d2, ctx->d_active, n_active, n, d_n, ctx->d_epoch, W, H,
            d_Rb, ctx->d_touch, ctx->d_del, ctx->d_rowstate,
            ctx->d_block_sums, nb_scan, d_n_out, ctx->d_status);
        AOSLO_LAUNCH_CHECK();
        ++ctx->epoch;                               // host-side upper estimate of the device epoch
        if (d_deleted_out) AOSLO_CUDA(cudaMemcpyAsync(d_deleted_out, ctx->d_del, n, cudaMemcpyDeviceToDevice, s));
        KeepOp op{(const double2*)d_pos, d_stable, ctx->d_del, (double2*)d_pos_out, d_stable_out};
        scan_emit_kernel<KeepOp><<<nb_scan, kScanThreads, 0, s>>>(op, n, d_n, ctx->d_block_sums);
        AOSLO_LAUNCH_CHECK();
        return AOSLO_OK;
    } else {
        AOSLO_TRY(stage_knn(ctx, s, d_pos, n, d_pos, n, 2, tie_mode, ctx->d_knn_idx, ctx->d_knn_d2, true, d_n));
        DeleteParams p;
        p.pos = (const double2*)d_pos; p.stable = d_stable; p.n_host = n; p.d_n = d_n;
        p.knn_idx = ctx->d_knn_idx; p.knn_d2 = ctx->d_knn_d2; p.thr = thr; p.d_thr = d_thr; p.W = W; p.H = H;
        p.touch = ctx->d_touch; p.del = ctx->d_del; p.rowstate = ctx->d_rowstate;
        p.counters = ctx->d_counters + 8; p.rounds_out = d_rounds_out; p.status = ctx->d_status;
        void* args[] = {&p, (void*)&d_Rb};
        int blocks = ctx->coop_blocks;
        int need = (n + kCoopThreads - 1) / kCoopThreads;   // n is an upper bound when the count lives on the device
        if (need < blocks) blocks = need < 1 ? 1 : need;
        AOSLO_CUDA(cudaLaunchCooperativeKernel((void*)delete_resolve_kernel<T>, dim3(blocks), dim3(kCoopThreads), args, 0, s));
        count_launch();
    }
    if (d_deleted_out) AOSLO_CUDA(cudaMemcpyAsync(d_deleted_out, ctx->d_del, n, cudaMemcpyDeviceToDevice, s));
    KeepOp op{(const double2*)d_pos, d_stable, ctx->d_del, (double2*)d_pos_out, d_stable_out};
    return ordered_scan(ctx, s, op, n, n, d_n, d_n_out, nullptr, 0, nullptr);
}

int stage_delete(aoslo_ctx* ctx, cudaStream_t s, const double* d_pos, const uint8_t* d_stable, int n,
                 const void* d_Rb, int field_dtype, int H, int W, double thr, int tie_mode, double* d_pos_out,
                 uint8_t* d_stable_out, uint8_t* d_deleted_out, int* d_n_out, int* d_rounds_out, const int* d_n,
                 int mode, const double* d_thr, int scan_hint) {
    if (field_dtype == AOSLO_F32)
        return stage_delete_t<float>(ctx, s, d_pos, d_stable, n, (const float*)d_Rb, H, W, thr, tie_mode, d_pos_out,
                                     d_stable_out, d_deleted_out, d_n_out, d_rounds_out, d_n, mode, d_thr, scan_hint);
    return stage_delete_t<double>(ctx, s, d_pos, d_stable, n, (const double*)d_Rb, H, W, thr, tie_mode, d_pos_out,
                                  d_stable_out, d_deleted_out, d_n_out, d_rounds_out, d_n, mode, d_thr, scan_hint);
}

// clamp_count: the stored count never exceeds cap (graph path); otherwise the caller compares it with cap
int stage_seed(aoslo_ctx* ctx, cudaStream_t s, const void* d_Rb, int field_dtype, int H, int W, int bx, int by,
               double thr, double* d_pos, int cap, int* d_n_out, const double* d_thr, bool clamp_count) {
    int wi = W - 2 * bx, hi = H - 2 * by;
    if (wi <= 0 || hi <= 0) return AOSLO_ERR_INVALID;
    int m = wi * hi;
    const int clamp = clamp_count ? cap : 0;
    if (field_dtype == AOSLO_F32) {
        SeedOp<float> op{(const float*)d_Rb, W, bx, by, wi, thr, d_thr, (double2*)d_pos, cap, ctx->d_status};
        return ordered_scan(ctx, s, op, m, m, nullptr, d_n_out, nullptr, 0, nullptr, clamp);
    }
    SeedOp<double> op{(const double*)d_Rb, W, bx, by, wi, thr, d_thr, (double2*)d_pos, cap, ctx->d_status};
    return ordered_scan(ctx, s, op, m, m, nullptr, d_n_out, nullptr, 0, nullptr, clamp);
}

int stage_lut(aoslo_ctx*, cudaStream_t s, int rfs, double mu, double sigma, double lam, double* d_lut,
              const RunParams* rp) {
    if (rfs < 1 || rfs > 64) return AOSLO_ERR_UNSUPPORTED;
    dist_lut_kernel<<<(rfs * rfs + 255) / 256, 256, 0, s>>>(rfs, mu, sigma, lam, d_lut, rp);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

// once per context (device): 63 x 93 doubles + the bitmap rows of the general kernel exceed the 48 KB default
int gravity_init() {
    AOSLO_CUDA(cudaFuncSetAttribute(gravity_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)grav_bits_smem(63)));
    return AOSLO_OK;
}

template <int HH>
static void launch_nearest(dim3 grid, cudaStream_t s, const aoslo_ctx* ctx, int wpr, int H, int W, double* d_field) {
    gravity_nearest_kernel<HH><<<grid, kGravCols, 0, s>>>(ctx->d_occ, wpr, ctx->d_lut_radial, ctx->d_lut_flag, H, W,
                                                          d_field);
}

// lut_ready: the radial table / flag of this LUT are already in the context (derived once per run)
int stage_gravity(aoslo_ctx* ctx, cudaStream_t s, const double* d_pos, int n_host, const int* d_n,
                  const double* d_lut, int rfs, int H, int W, double* d_field, bool lut_ready) {
    if (rfs < 1 || rfs > 64) return AOSLO_ERR_UNSUPPORTED;
    if ((rfs & 1) == 0) return AOSLO_ERR_INVALID;     // even rfs: shape mismatch in the reference (utils.py:329)
    if (H != ctx->H || W != ctx->W) return AOSLO_ERR_INVALID;
    const int h = rfs / 2, wpr = occ_words_per_row(W);
    if (!lut_ready) {
        lut_radial_kernel<<<1, 1024, 0, s>>>(d_lut, rfs, ctx->d_lut_radial, ctx->d_lut_flag);
        AOSLO_LAUNCH_CHECK();
    }
    const int n16 = (int)(((size_t)H * wpr + 3) / 4);
    occ_clear_kernel<<<ctx->num_sms, kThreads, 0, s>>>(reinterpret_cast<uint4*>(ctx->d_occ), n16, nullptr);
    AOSLO_LAUNCH_CHECK();
    occ_scatter_kernel<<<ctx->nb, kThreads, 0, s>>>((const double2*)d_pos, nullptr, n_host, d_n, W, H, wpr,
                                                    ctx->d_occ, ctx->d_status, nullptr);
    AOSLO_LAUNCH_CHECK();
    if (h >= 1 && h <= kNearMaxH) {                   // otherwise the flag is 0 and the general kernel works
        dim3 grid((W + kGravCols - 1) / kGravCols, (H + kNearRows - 1) / kNearRows);
        switch (h) {
#define AOSLO_NEAR_CASE(HH) case HH: launch_nearest<HH>(grid, s, ctx, wpr, H, W, d_field); break;
            AOSLO_NEAR_CASE(1) AOSLO_NEAR_CASE(2) AOSLO_NEAR_CASE(3) AOSLO_NEAR_CASE(4) AOSLO_NEAR_CASE(5)
            AOSLO_NEAR_CASE(6) AOSLO_NEAR_CASE(7) AOSLO_NEAR_CASE(8) AOSLO_NEAR_CASE(9) AOSLO_NEAR_CASE(10)
            AOSLO_NEAR_CASE(11) AOSLO_NEAR_CASE(12) AOSLO_NEAR_CASE(13) AOSLO_NEAR_CASE(14) AOSLO_NEAR_CASE(15)
#undef AOSLO_NEAR_CASE
            default: break;
        }
        AOSLO_LAUNCH_CHECK();
    }
    {
        const int row_tiles = (H + kBitsRows - 1) / kBitsRows;
        dim3 grid((W + kGravCols - 1) / kGravCols, (row_tiles + kBitsTilesPerCta - 1) / kBitsTilesPerCta);
        gravity_bits_kernel<<<grid, kGravCols, grav_bits_smem(rfs), s>>>(ctx->d_occ, wpr, d_lut, rfs, ctx->d_lut_flag,
                                                                         H, W, d_field);
        AOSLO_LAUNCH_CHECK();
    }
    return AOSLO_OK;
}

int stage_add(aoslo_ctx* ctx, cudaStream_t s, const double* d_field, int H, int W, int bx, int by, int window,
              double thr, double* d_pos, uint8_t* d_stable, int n_host, const int* d_n, int cap, int* d_n_out,
              const double* d_thr, bool clamp_count) {
    if (window < 1) return AOSLO_ERR_INVALID;
    int nwx = (W - 2 * bx + window - 1) / window, nwy = (H - 2 * by + window - 1) / window;
    if (nwx <= 0 || nwy <= 0) return AOSLO_ERR_INVALID;
    int nw = nwx * nwy;
    window_argmin_kernel<<<(nw + kThreads - 1) / kThreads, kThreads, 0, s>>>(d_field, H, W, bx, by, window, thr,
                                                                           nwx, nwy, ctx->d_win, d_thr);
    AOSLO_LAUNCH_CHECK();
    AppendOp op{ctx->d_win, W, (double2*)d_pos, d_stable, cap, ctx->d_status};
    return ordered_scan(ctx, s, op, nw, nw, nullptr, d_n_out, d_n, n_host, nullptr, clamp_count ? cap : 0);
}

int stage_force(aoslo_ctx* ctx, cudaStream_t s, const double* d_pos, const uint8_t* d_stable, int n, int k,
                int energy_type, double p0, double p1, double p2, int tie_mode, double* d_force, const int* d_n) {
    if (k < 1 || k + 1 > AOSLO_KMAX) return AOSLO_ERR_UNSUPPORTED;
    if (n > ctx->cap) return AOSLO_ERR_CAPACITY;
    if (n < k + 1) return AOSLO_ERR_INVALID;
    AOSLO_TRY(stage_knn(ctx, s, d_pos, n, d_pos, n, k + 1, tie_mode, ctx->d_knn_idx, ctx->d_knn_d2, true, d_n));
    dist_force_kernel<<<ctx->nb, kThreads, 0, s>>>((const double2*)d_pos, d_stable, n, d_n, ctx->d_knn_idx, k + 1,
                                                   energy_type, p0, p1, p2, (double2*)d_force);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

int stage_move(aoslo_ctx* ctx, cudaStream_t s, double* d_pos, const uint8_t* d_stable, int n, const void* d_grad,
               int field_dtype, int H, int W, const double* d_force, double lb, double ld, int step_mode,
               const int* d_n) {
    if (field_dtype == AOSLO_F32)
        move_kernel<float><<<ctx->nb, kThreads, 0, s>>>((double2*)d_pos, d_stable, n, d_n, (const float*)d_grad, H,
                                                        W, (const double2*)d_force, lb, ld, step_mode, ctx->d_status);
    else
        move_kernel<double><<<ctx->nb, kThreads, 0, s>>>((double2*)d_pos, d_stable, n, d_n, (const double*)d_grad,
                                                         H, W, (const double2*)d_force, lb, ld, step_mode,
                                                         ctx->d_status);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

// particles cell list must NOT be assumed: builds both lists (GT list can be cached by the caller)
int stage_metrics(aoslo_ctx* ctx, cudaStream_t s, const double* d_pos, const uint8_t* d_stable, int n,
                  const double* d_gt, int n_gt, bool gt_list_ready, bool p_list_ready, const void* d_Rb,
                  int field_dtype, int H, int W, int bx, int by, int iteration, aoslo_metrics_record* d_rec,
                  double* d_p2g, double* d_g2p, const int* d_n, const int* d_n_gt, int* d_rec_idx,
                  const int* d_iteration) {
    if (n > ctx->cap || n_gt > ctx->gt_cap) return AOSLO_ERR_CAPACITY;
    if (n < 1 || n_gt < 1) return AOSLO_ERR_INVALID;
    if (!gt_list_ready) AOSLO_TRY(build_cell_list(ctx, s, ctx->cl_g, d_gt, n_gt, d_n_gt));
    if (!p_list_ready) AOSLO_TRY(build_links(ctx, s, ctx->lk_p, d_pos, n, d_n, kLinkBiasSteady));
    // one pass: particle -> nearest GT ("gt_to_particles", :322) and GT -> nearest particle
    // ("particles_to_gt", :323), counts, moments and blob energy
    int* icount = ctx->d_counters + 16;
    AOSLO_CUDA(cudaMemsetAsync(icount, 0, sizeof(int) * 9, s));
    MetricsPartial* partials = (MetricsPartial*)ctx->d_partials;
    if (field_dtype == AOSLO_F32)
        metrics_pass_kernel<float><<<ctx->nb, kThreads, 0, s>>>(
            view_of(ctx->cl_g), link_view(ctx, ctx->lk_p, d_pos), (const double2*)d_pos, d_stable, n, d_n,
            (const double2*)d_gt, n_gt, (const float*)d_Rb, H, W, bx, by, d_p2g, d_g2p, icount, partials, iteration,
            d_rec, ctx->d_status, d_n_gt, d_rec_idx, d_iteration);
    else
        metrics_pass_kernel<double><<<ctx->nb, kThreads, 0, s>>>(
            view_of(ctx->cl_g), link_view(ctx, ctx->lk_p, d_pos), (const double2*)d_pos, d_stable, n, d_n,
            (const double2*)d_gt, n_gt, (const double*)d_Rb, H, W, bx, by, d_p2g, d_g2p, icount, partials, iteration,
            d_rec, ctx->d_status, d_n_gt, d_rec_idx, d_iteration);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

int stage_set_stable(aoslo_ctx* ctx, cudaStream_t s, uint8_t* d_stable, int n, uint8_t v, const int* d_n) {
    set_all_stable_kernel<<<ctx->nb, kThreads, 0, s>>>(d_stable, n, d_n, v);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

int stage_count_stable(aoslo_ctx* ctx, cudaStream_t s, const uint8_t* d_stable, int n, int* d_out, const int* d_n) {
    AOSLO_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int), s));
    count_stable_kernel<<<ctx->nb, kThreads, 0, s>>>(d_stable, n, d_n, d_out);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

// ---- host drivers of the run loop's tombstone / dual-structure scheme (aoslo_run, canonical tie order) ----
static DualView dual_view(aoslo_ctx* ctx) {
    return DualView{ctx->d_link_geom, ctx->lk_p.head, ctx->d_head_n, ctx->lk_p.next, (const double2*)ctx->d_pos[0]};
}

// start of a resample period on buffer 0: tombstones cleared, S / N linked, non-stable list collected
int stage_period_begin(aoslo_ctx* ctx, cudaStream_t s, int n_bound, const int* d_n) {
    int* n_ns = ctx->d_counters + 13;
    AOSLO_CUDA(cudaMemsetAsync(n_ns, 0, sizeof(int), s));
    period_begin_kernel<<<ctx->nb, kThreads, 0, s>>>(
        (const double2*)ctx->d_pos[0], ctx->d_stable[0], n_bound, d_n, ctx->W, ctx->H, ctx->lk_p.min_shift,
        ctx->lk_p.fixed_shift, kLinkBiasSteady, ctx->d_link_geom, ctx->d_link_epoch, ctx->d_link_epoch_n,
        ctx->d_link_epoch_ticket, ctx->lk_p.head, ctx->d_head_n, ctx->lk_p.next, ctx->d_del, ctx->d_ns_list, n_ns,
        ctx->d_counters + 7, ctx->d_status);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

// one iteration: force + move of the non-stable survivors, re-link, deletion rows, resolve (no compaction)
int stage_dual_step(aoslo_ctx* ctx, cudaStream_t s, int n_bound, const int* d_n, int k, int energy_type, double p0,
                    double p1, double p2, const void* d_grad, const void* d_Rb, int field_dtype, int H, int W, double lb,
                    double ld, int step_mode, double del_thr, const RunParams* rp, int phases) {
    // phases: bit 0 = force + move, bit 1 = re-link + deletion (the caller may put a time stamp between them)
    if (k < 1 || k + 1 > AOSLO_KMAX) return AOSLO_ERR_UNSUPPORTED;
    int* n_ns = ctx->d_counters + 13;
    int* n_active = ctx->d_counters + 12;
    int* n_dead = ctx->d_counters + 7;
    const DualView dv = dual_view(ctx);
    const int kk = k + 1;
    if (phases & 1) {
#define AOSLO_DFM(KT, TT)                                                                                         \
    dual_force_move_kernel<KT, TT><<<ctx->nb, kThreads, 0, s>>>(dv, ctx->d_ns_list, n_ns, ctx->d_del, kk,          \
                                                                energy_type, p0, p1, p2, (const TT*)d_grad, H, W,  \
                                                                lb, ld, step_mode, (double2*)ctx->d_pos_tmp,       \
                                                                ctx->d_status, rp)
    if (field_dtype == AOSLO_F32) {
        if (kk <= 2) AOSLO_DFM(2, float); else if (kk <= 4) AOSLO_DFM(4, float); else AOSLO_DFM(8, float);
    } else {
        if (kk <= 2) AOSLO_DFM(2, double); else if (kk <= 4) AOSLO_DFM(4, double); else AOSLO_DFM(8, double);
    }
#undef AOSLO_DFM
    AOSLO_LAUNCH_CHECK();
    }
    if (!(phases & 2)) return AOSLO_OK;
    // the non-stable list is short (thousands): a quarter of the grid keeps the ticket chain short
    ns_relink_kernel<<<ctx->num_sms, kThreads, 0, s>>>(ctx->d_ns_list, n_ns, (const double2*)ctx->d_pos_tmp,
                                                  (double2*)ctx->d_pos[0], ctx->d_del, W, H, ctx->d_link_geom,
                                                  ctx->d_link_epoch_n, ctx->d_link_epoch_ticket, ctx->d_head_n,
                                                  ctx->lk_p.next, ctx->d_status);
    AOSLO_LAUNCH_CHECK();
    if (field_dtype == AOSLO_F32)
        dual_rows_kernel<float><<<ctx->nb, kThreads, 0, s>>>(
            dv, (double2*)ctx->d_pos[0], ctx->d_stable[0], ctx->d_del, ctx->d_ns_list, n_ns, del_thr,
            rp ? &rp->del_thr : nullptr, ctx->d_knn_idx, ctx->d_knn_d2, ctx->d_active, n_active,
            ctx->d_link_epoch_ticket, ctx->d_epoch, W, H, (const float*)d_Rb, ctx->d_touch, ctx->d_rowstate, n_dead,
            ctx->d_status);
    else
        dual_rows_kernel<double><<<ctx->nb, kThreads, 0, s>>>(
            dv, (double2*)ctx->d_pos[0], ctx->d_stable[0], ctx->d_del, ctx->d_ns_list, n_ns, del_thr,
            rp ? &rp->del_thr : nullptr, ctx->d_knn_idx, ctx->d_knn_d2, ctx->d_active, n_active,
            ctx->d_link_epoch_ticket, ctx->d_epoch, W, H, (const double*)d_Rb, ctx->d_touch, ctx->d_rowstate, n_dead,
            ctx->d_status);
    AOSLO_LAUNCH_CHECK();
    (void)n_bound; (void)d_n;
    return AOSLO_OK;
}

// order-preserving compaction of the survivors of buffer 0 (through buffer 1 and back); *d_n becomes the survivor
// count, the tombstone counter is reset
int stage_compact(aoslo_ctx* ctx, cudaStream_t s, int n_bound, int* d_n, int* d_n_scratch, int scan_hint) {
    KeepAliveOp op{(const double2*)ctx->d_pos[0], ctx->d_stable[0], ctx->d_del, (double2*)ctx->d_pos[1],
                   ctx->d_stable[1]};
    AOSLO_TRY(ordered_scan(ctx, s, op, scan_hint > 0 ? scan_hint : n_bound, n_bound, d_n, d_n_scratch, nullptr, 0,
                           nullptr));
    copy_particles_kernel<<<ctx->nb, kThreads, 0, s>>>((const double2*)ctx->d_pos[1], ctx->d_stable[1], d_n_scratch,
                                                       (double2*)ctx->d_pos[0], ctx->d_stable[0], d_n,
                                                       ctx->d_counters + 7);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

// occupancy bitmap for nearest-distance queries: *ok = 1 iff every (surviving) point is an in-image integer pixel
static int occ_build(aoslo_ctx* ctx, cudaStream_t s, uint32_t* occ, const double2* pos, const uint8_t* del, int n_host,
                     const int* d_n, int H, int W, int* ok) {
    const int wpr = occ_words_per_row(W);
    const int n16 = (int)(((size_t)H * wpr + 3) / 4);
    occ_clear_kernel<<<ctx->num_sms, kThreads, 0, s>>>(reinterpret_cast<uint4*>(occ), n16, ok);
    AOSLO_LAUNCH_CHECK();
    occ_scatter_kernel<<<ctx->nb, kThreads, 0, s>>>(pos, del, n_host, d_n, W, H, wpr, occ, ctx->d_status, ok);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

// once per run: the ground truth's bitmap for the particle -> nearest GT distances of stage_dual_metrics
int stage_gt_bitmap(aoslo_ctx* ctx, cudaStream_t s, const double* d_gt, int n_gt, const int* d_n_gt) {
    return occ_build(ctx, s, ctx->d_occ_gt, (const double2*)d_gt, nullptr, n_gt, d_n_gt, ctx->H, ctx->W,
                     ctx->d_occ_ok + 1);
}

int stage_dual_metrics(aoslo_ctx* ctx, cudaStream_t s, int n_bound, const int* d_n, const double* d_gt, int n_gt,
                       const void* d_Rb, int field_dtype, int H, int W, int bx, int by, int iteration,
                       aoslo_metrics_record* d_rec, const int* d_n_gt, int* d_rec_idx, const int* d_iteration) {
    if (n_gt > ctx->gt_cap) return AOSLO_ERR_CAPACITY;
    if (n_gt < 1) return AOSLO_ERR_INVALID;
    int* icount = ctx->d_counters + 16;
    AOSLO_CUDA(cudaMemsetAsync(icount, 0, sizeof(int) * 9, s));
    MetricsPartial* partials = (MetricsPartial*)ctx->d_partials;
    const DualView dv = dual_view(ctx);
    // bitmap of the survivors (the GT bitmap is per run: stage_gt_bitmap)
    const int wpr = occ_words_per_row(W);
    AOSLO_TRY(occ_build(ctx, s, ctx->d_occ, dv.pos, ctx->d_del, n_bound, d_n, H, W, ctx->d_occ_ok));
    if (field_dtype == AOSLO_F32)
        dual_metrics_kernel<float><<<ctx->nb, kThreads, 0, s>>>(
            view_of(ctx->cl_g), dv, ctx->d_stable[0], ctx->d_del, n_bound, d_n, ctx->d_counters + 7,
            (const double2*)d_gt, n_gt, (const float*)d_Rb, H, W, bx, by, icount, partials, iteration, d_rec,
            ctx->d_status, d_n_gt, d_rec_idx, d_iteration, ctx->d_occ, ctx->d_occ_gt, wpr, ctx->d_occ_ok);
    else
        dual_metrics_kernel<double><<<ctx->nb, kThreads, 0, s>>>(
            view_of(ctx->cl_g), dv, ctx->d_stable[0], ctx->d_del, n_bound, d_n, ctx->d_counters + 7,
            (const double2*)d_gt, n_gt, (const double*)d_Rb, H, W, bx, by, icount, partials, iteration, d_rec,
            ctx->d_status, d_n_gt, d_rec_idx, d_iteration, ctx->d_occ, ctx->d_occ_gt, wpr, ctx->d_occ_ok);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

int stage_dual_count_stable(aoslo_ctx* ctx, cudaStream_t s, int n_bound, const int* d_n, int* d_out) {
    AOSLO_CUDA(cudaMemsetAsync(d_out, 0, sizeof(int), s));
    dual_count_stable_kernel<<<ctx->nb, kThreads, 0, s>>>(ctx->d_stable[0], ctx->d_del, n_bound, d_n, d_out);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

__global__ void sqrt_inplace_kernel(double* a, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        a[i] = sqrt(a[i]);
}

int sqrt_inplace(cudaStream_t s, double* a, size_t n) {
    if (n == 0) return AOSLO_OK;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    sqrt_inplace_kernel<<<blocks, 256, 0, s>>>(a, n);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

// =======================================================================================
// input preparation on the device (reference pipeline_with_delitions.py:27-33):
// np.round + crop_image (image_transformation.py:6-21) + mirror_padding_image (:23-35)
// =======================================================================================
__global__ void __launch_bounds__(kThreads) crop_pad_kernel(const uint8_t* __restrict__ raw, int pitch0, int x_min,
                                                            int y_min, int wc, int hc, int bx, int by,
                                                            uint8_t* __restrict__ out, int out_pitch, int W, int H) {
    int x = blockIdx.x * 64 + (threadIdx.x & 63);
    int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= W || y >= H) return;
    int cx = x - bx, cy = y - by;                       // symmetric (edge-including) mirror
    cx = cx < 0 ? -1 - cx : (cx >= wc ? 2 * wc - 1 - cx : cx);
    cy = cy < 0 ? -1 - cy : (cy >= hc ? 2 * hc - 1 - cy : cy);
    out[(size_t)y * out_pitch + x] = raw[(size_t)(y_min + cy) * pitch0 + (x_min + cx)];
}

struct GtPrepOp {
    const double2* gt;          // shift = 0: already shifted by -1 on the host (the reference mutates the caller's
    double shift;               // array); shift = 1: the caller's raw 1-based array, GT_enumerate_from_zero applied here
    double x_min, x_max, y_min, y_max, bx, by;
    double2* out;
    __device__ int value(int i) const {
        double gx = rint(gt[i].x - shift), gy = rint(gt[i].y - shift);   // np.round: half to even
        return (gy > y_min && gy < y_max && gx > x_min && gx < x_max) ? 1 : 0;
    }
    __device__ void emit(int i, int v, int dst) const {
        if (!v) return;
        double gx = rint(gt[i].x - shift), gy = rint(gt[i].y - shift);
        out[dst] = make_double2(gx - x_min + bx, gy - y_min + by);
    }
};

int stage_prepare(aoslo_ctx* ctx, cudaStream_t s, const uint8_t* d_raw, int H0, int W0, int pitch0, int x_min,
                  int x_max, int y_min, int y_max, int bx, int by, uint8_t* d_img_out, int out_pitch,
                  const double* d_gt_raw, int n_gt_raw, double* d_gt_out, int* d_n_gt_out, int gt_is_raw) {
    if (x_min < 0 || y_min < 0 || x_max >= W0 || y_max >= H0) return AOSLO_ERR_INVALID;   // the reference's asserts
    const int wc = x_max - x_min, hc = y_max - y_min;
    if (wc < 1 || hc < 1 || bx < 0 || by < 0 || bx > wc || by > hc) return AOSLO_ERR_INVALID;
    const int W = wc + 2 * bx, H = hc + 2 * by;
    if (W != ctx->W || H != ctx->H || out_pitch < W) return AOSLO_ERR_INVALID;
    dim3 grid((W + 63) / 64, (H + 3) / 4);
    crop_pad_kernel<<<grid, kThreads, 0, s>>>(d_raw, pitch0, x_min, y_min, wc, hc, bx, by, d_img_out, out_pitch, W, H);
    AOSLO_LAUNCH_CHECK();
    if (d_gt_raw && n_gt_raw > 0) {
        GtPrepOp op{(const double2*)d_gt_raw, gt_is_raw ? 1.0 : 0.0, (double)x_min, (double)x_max, (double)y_min, (double)y_max,
                    (double)bx, (double)by, (double2*)d_gt_out};
        AOSLO_TRY(ordered_scan(ctx, s, op, n_gt_raw, n_gt_raw, nullptr, d_n_gt_out, nullptr, 0, nullptr));
    }
    return AOSLO_OK;
}

// classification arrays of utils.visualize_all (utils.py:174-203): per GT point TP (nearest particle
// strictly closer than the threshold) / FN, per particle FP (nearest GT farther than the threshold AND
// strictly inside the original image, get_particles_in_image)
__global__ void __launch_bounds__(kThreads) classify_kernel(const double* __restrict__ d_g2p, int n_gt,
                                                            const double* __restrict__ d_p2g,
                                                            const double2* __restrict__ pos, int n, int bx, int by,
                                                            int W, int H, double thr, uint8_t* gt_is_tp,
                                                            uint8_t* particle_is_fp) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_gt + n; i += gridDim.x * blockDim.x) {
        if (i < n_gt) {
            gt_is_tp[i] = (d_g2p[i] < thr) ? 1 : 0;
        } else {
            const int j = i - n_gt;
            const double2 p = pos[j];
            const bool inside = ((double)bx < p.x) && (p.x < (double)(W - bx)) && ((double)by < p.y) &&
                                (p.y < (double)(H - by));
            particle_is_fp[j] = (inside && d_p2g[j] > thr) ? 1 : 0;
        }
    }
}

int stage_classify(aoslo_ctx* ctx, cudaStream_t s, const double* d_g2p, int n_gt, const double* d_p2g,
                   const double* d_pos, int n, int bx, int by, int W, int H, double thr, uint8_t* gt_is_tp,
                   uint8_t* particle_is_fp) {
    classify_kernel<<<ctx->nb, kThreads, 0, s>>>(d_g2p, n_gt, d_p2g, (const double2*)d_pos, n, bx, by, W, H, thr,
                                                 gt_is_tp, particle_is_fp);
    AOSLO_LAUNCH_CHECK();
    return AOSLO_OK;
}

int thin_threads_default(int H, int W, int num_sms) {
    // Two blocks per SM; block size by bitmap words per SM, about four words per thread.  The smaller the footprint
    // of the cooperative kernel, the more of the other lanes' kernels run beside it (2077^2, 8 lanes, runs/s
    // resident: 512 x 1: 1 308, 256 x 2: 1 293, 256 x 1: 1 350, 128 x 2: 1 353; single-run latency 2.03 / 2.02 /
    // 2.15 / 2.16 ms) -- but a 4126^2 frame on 2 x 128 threads per SM walks 14 words per thread (4126^2, 2 lanes,
    // block size 128 / 256 / 512: 235 / 265 / 254 runs/s, 6.03 / 5.09 / 4.70 ms; 4 lanes: 256 -> 282, 512 -> 269).
    const long long words_per_sm = ((long long)H * occ_words_per_row(W) + num_sms - 1) / (num_sms > 0 ? num_sms : 1);
    int t = words_per_sm <= 1024 ? 128 : 256;
    if (const char* e = getenv("AOSLO_THIN_THREADS")) { int v = atoi(e); if (v == 128 || v == 256 || v == 512) t = v; }
    return t;
}

int coop_blocks_for_thin(int num_sms, int threads) {
    int per_sm_f = 0, per_sm_d = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_f, thin_pixels_kernel<float>, threads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_d, thin_pixels_kernel<double>, threads, 0);
    int per_sm = per_sm_f < per_sm_d ? per_sm_f : per_sm_d;
    if (per_sm < 1) per_sm = 1;
    int want = 2;
    if (const char* e = getenv("AOSLO_THIN_BLOCKS_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 4) want = v; }
    if (per_sm > want) per_sm = want;
    return per_sm * num_sms;
}

// seeds + thinning to the fixpoint of aoslo_run (canonical order, threshold <= 4 px): the survivors land in
// d_pos / d_stable, their count in *d_n_out (clamped to cap), n_seeds / n_after_thinning / thin_rounds in *ctl
int stage_thin_pixels(aoslo_ctx* ctx, cudaStream_t s, const void* d_Rb, int field_dtype, int H, int W, int bx, int by,
                      double seed_thr, double thin_thr, const RunParams* rp, RunCtl* ctl, double* d_pos,
                      uint8_t* d_stable, int* d_n_out) {
    if (H != ctx->H || W != ctx->W || W - 2 * bx <= 0 || H - 2 * by <= 0) return AOSLO_ERR_INVALID;
    if (!(thin_thr <= 4.0)) return AOSLO_ERR_UNSUPPORTED;
    ThinParams p;
    p.H = H; p.W = W; p.bx = bx; p.by = by; p.wpr = occ_words_per_row(W); p.cap = ctx->cap;
    p.seed_thr = seed_thr; p.thin_thr = thin_thr; p.rp = rp;
    const size_t nwords = (size_t)H * p.wpr;
    p.alive = ctx->d_occ;                                   // free until the first resample
    p.planes = ctx->d_thin;
    p.pend = ctx->d_thin + (size_t)kThinPlanes * nwords;
    p.touch[0] = reinterpret_cast<unsigned long long*>(ctx->d_field);   // free until the first resample
    p.touch[1] = ctx->d_thin_touch;
    p.ctr = ctx->d_thin_ctr;
    p.ctl = ctl;
    p.status = ctx->d_status;
    AOSLO_CUDA(cudaMemsetAsync(p.ctr, 0, sizeof(int) * 8, s));
    int blocks = ctx->coop_blocks_thin;
    const int threads = ctx->thin_threads;
    const int need = (int)((nwords + threads - 1) / threads);
    if (need < blocks) blocks = need < 1 ? 1 : need;
    void* args[] = {&p, (void*)&d_Rb};
    if (field_dtype == AOSLO_F32)
        AOSLO_CUDA(cudaLaunchCooperativeKernel((void*)thin_pixels_kernel<float>, dim3(blocks), dim3(threads), args, 0, s));
    else
        AOSLO_CUDA(cudaLaunchCooperativeKernel((void*)thin_pixels_kernel<double>, dim3(blocks), dim3(threads), args, 0, s));
    count_launch();
    AliveEmitOp op{ctx->d_occ, p.wpr, (double2*)d_pos, d_stable, ctx->cap, ctx->d_status};
    return ordered_scan(ctx, s, op, (int)nwords, (int)nwords, nullptr, d_n_out, nullptr, 0, nullptr, ctx->cap);
}

int coop_blocks_for_delete(int num_sms) {
    int per_sm_f = 0, per_sm_d = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_f, delete_resolve_kernel<float>, kCoopThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_d, delete_resolve_kernel<double>, kCoopThreads, 0);
    int per_sm = per_sm_f < per_sm_d ? per_sm_f : per_sm_d;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 1) per_sm = 1;        // one 1024-thread block per SM: the grid barrier costs per block
    return per_sm * num_sms;
}

}  // namespace aoslo
