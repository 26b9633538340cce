// hungarian.cu -- metric_algo = 'hangarian' of the reference on the device (sm_100a; compiled with -fmad=false).
//
// Replaces utils.calc_metrics(..., mode='hangarian') (utils.py:228-234):
//     dist_matr = scipy.spatial.distance_matrix(GT_particles, particles)          -> dist_matrix_kernel
//     row_ind, col_ind = scipy.optimize.linear_sum_assignment(dist_matr)          -> lsap_kernel
//     dist_matr[row_ind, col_ind].sum() / .mean() / .var()                        -> lsap_kernel (tail)
//
// linear_sum_assignment is scipy's rectangular LSAP solver (shortest augmenting paths, Crouse 2016; scipy is a
// third-party dependency of the reference, installed here as 1.18.1).  The distances of lattice points tie all the
// time, so WHICH optimal assignment comes out depends on the solver's scan order and tie rule, and the variance of the
// assigned costs depends on the assignment.  lsap_kernel therefore restates the solver step by step:
//   * rows = the smaller point set (scipy transposes a tall matrix), processed in index order;
//   * per row one shortest-path search: every step relaxes all columns that are still `remaining` with
//     r = minVal + cost[i][j] - u[i] - v[j] (in this order of operations) and picks the column with the lowest
//     shortest-path cost; among equal ones the LAST unassigned column in the order of scipy's `remaining` vector if
//     there is one, else the FIRST assigned one; `remaining` starts as nc-1 .. 0 and loses the picked entry by
//     swap-with-last, which is why each column's current position is tracked (inside its tie key `tkey`);
//   * dual update u[i] += minVal - spc[col4row[i]] over the visited rows, v[j] -= minVal - spc[j] over the visited
//     columns, then the augmentation along `path`.
// tests/test_host_logic.py holds the same restatement in NumPy and checks it against scipy on tie-heavy instances
// (identical assignments); tests/test_gpu_parity.py checks the kernel against scipy on lattice point sets.
//
// One search step is a scan over <= nc columns followed by an arg-min: it is spread over the 1024 threads of ONE CTA
// (column j belongs to thread j mod 1024) with a shuffle / shared-memory reduction on the key (cost, tie class,
// position); the steps of a search and the rows are inherently sequential (~120 steps per row on the shipped image's
// 7 602 x 8 156 problem).  Independent runs (lanes) solve their problems concurrently on different SMs.
#include <cfloat>

#include "common.cuh"

namespace aoslo {

__global__ void dist_matrix_kernel(const double2* __restrict__ a, int nr, const double2* __restrict__ b, int nc,
                                   double* __restrict__ cost) {
    // scipy.spatial.distance_matrix -> minkowski_distance(p=2): sum(|x - y| ** 2) ** 0.5 (square, sum, sqrt)
    const size_t total = (size_t)nr * nc;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / nc), j = (int)(t - (size_t)i * nc);
        const double dx = fabs(b[j].x - a[i].x), dy = fabs(b[j].y - a[i].y);
        cost[t] = sqrt(dx * dx + dy * dy);
    }
}

struct LsapWork {
    int nr, nc;
    const double* cost;      // [nr * nc]
    double* u;               // [nr]
    double* v;               // [nc]
    double* spc;             // [nc] shortestPathCosts
    int* path;               // [nc]
    int* col4row;            // [nr]
    int* row4col;            // [nc]
    int* remaining;          // [nc]
    int* tkey;               // [nc] -1: taken out of `remaining` in this search; else the tie key of the column:
                             //      unassigned -> nc-1-pos (later position first), assigned -> nc+pos (earlier first)
    int* sr_list;            // [nr] rows visited by this search
    int* sc_list;            // [nc] columns visited by this search
    double* sel;             // [nr] cost of the assigned pairs, row order
    double* out;             // [4] sum, mean, var, number of search steps
    int* status;
};

constexpr int kLsapThreads = 1024;
constexpr int kLsapBatch = 4;      // columns per thread and batch: 16 independent loads in flight

struct LsapKey {
    double s;
    int t;      // tie-break: unassigned columns first, the later position first; then assigned, earlier first
    int j;
};
__device__ __forceinline__ bool lsap_less(const LsapKey& a, const LsapKey& b) {
    return a.s < b.s || (a.s == b.s && a.t < b.t);
}

__global__ void __launch_bounds__(kLsapThreads) lsap_kernel(LsapWork w) {
    __shared__ double s_s[kLsapThreads / 32];
    __shared__ int s_t[kLsapThreads / 32], s_j[kLsapThreads / 32];
    __shared__ double s_min_val, s_ui;
    __shared__ int s_i, s_sink, s_num, s_nsr, s_nsc;
    __shared__ double s_red[kLsapThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nr = w.nr, nc = w.nc;
    for (int i = tid; i < nr; i += kLsapThreads) { w.u[i] = 0.0; w.col4row[i] = -1; }
    for (int j = tid; j < nc; j += kLsapThreads) { w.v[j] = 0.0; w.path[j] = -1; w.row4col[j] = -1; }
    double n_steps = 0.0;
    __syncthreads();
    for (int cur = 0; cur < nr; ++cur) {
        for (int j = tid; j < nc; j += kLsapThreads) {
            w.spc[j] = INFINITY;
            w.remaining[nc - 1 - j] = j;                           // remaining[it] = nc - it - 1: pos(j) = nc-1-j
            w.tkey[j] = (w.row4col[j] == -1) ? j : (2 * nc - 1 - j);
        }
        if (tid == 0) { s_min_val = 0.0; s_i = cur; s_ui = w.u[cur]; s_sink = -1; s_num = nc; s_nsr = 0; s_nsc = 0; }
        __syncthreads();
        while (true) {
            const int i = s_i;
            const double min_val = s_min_val;
            const double ui = s_ui;
            const double* crow = w.cost + (size_t)i * nc;
            LsapKey best{INFINITY, 0x7fffffff, -1};
            for (int j0 = tid; j0 < nc; j0 += kLsapThreads * kLsapBatch) {
                double c[kLsapBatch], vv[kLsapBatch], ss[kLsapBatch];
                int tk[kLsapBatch];
#pragma unroll
                for (int k = 0; k < kLsapBatch; ++k) {
                    const int j = j0 + k * kLsapThreads;
                    const bool ok = j < nc;
                    tk[k] = ok ? w.tkey[j] : -1;
                    c[k] = ok ? crow[j] : 0.0;
                    vv[k] = ok ? w.v[j] : 0.0;
                    ss[k] = ok ? w.spc[j] : INFINITY;
                }
#pragma unroll
                for (int k = 0; k < kLsapBatch; ++k) {
                    if (tk[k] < 0) continue;
                    const int j = j0 + k * kLsapThreads;
                    const double r = ((min_val + c[k]) - ui) - vv[k];
                    double sv = ss[k];
                    if (r < sv) { w.path[j] = i; w.spc[j] = r; sv = r; }
                    const LsapKey key{sv, tk[k], j};
                    if (lsap_less(key, best)) best = key;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                LsapKey k{__shfl_down_sync(0xffffffffu, best.s, o), __shfl_down_sync(0xffffffffu, best.t, o),
                          __shfl_down_sync(0xffffffffu, best.j, o)};
                if (lsap_less(k, best)) best = k;
            }
            if (lane == 0) { s_s[warp] = best.s; s_t[warp] = best.t; s_j[warp] = best.j; }
            __syncthreads();
            if (warp == 0) {
                best = LsapKey{s_s[lane], s_t[lane], s_j[lane]};
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    LsapKey k{__shfl_down_sync(0xffffffffu, best.s, o), __shfl_down_sync(0xffffffffu, best.t, o),
                              __shfl_down_sync(0xffffffffu, best.j, o)};
                    if (lsap_less(k, best)) best = k;
                }
                if (lane == 0) {
                    w.sr_list[s_nsr++] = i;                        // SR[i] = true
                    if (best.j < 0 || !(best.s < INFINITY)) {      // infeasible cost matrix (never: costs are finite)
                        atomicOr(w.status, AOSLO_DEV_NAN);
                        s_sink = -2;
                    } else {
                        const int js = best.j;
                        const int num = --s_num;                   // remaining[index] = remaining[--num_remaining]
                        const int jm = w.remaining[num];
                        const int r4c = w.row4col[js];
                        const int r4m = w.row4col[jm];
                        const int index = (best.t < nc) ? (nc - 1 - best.t) : (best.t - nc);   // position of js
                        s_min_val = best.s;
                        if (r4c == -1) {
                            s_sink = js;
                        } else {
                            s_i = r4c;
                            s_ui = w.u[r4c];
                        }
                        w.sc_list[s_nsc++] = js;                   // SC[j] = true
                        w.remaining[index] = jm;
                        if (jm != js) w.tkey[jm] = (r4m == -1) ? (nc - 1 - index) : (nc + index);
                        w.tkey[js] = -1;
                    }
                }
            }
            n_steps += 1.0;
            __syncthreads();
            if (s_sink != -1) break;
        }
        const int sink = s_sink;
        if (sink < 0) break;
        // dual variables
        const double min_val = s_min_val;
        const int nsr = s_nsr, nsc = s_nsc;
        if (tid == 0) w.u[cur] = w.u[cur] + min_val;
        for (int k = tid; k < nsr; k += kLsapThreads) {
            const int i2 = w.sr_list[k];
            if (i2 != cur) w.u[i2] = w.u[i2] + (min_val - w.spc[w.col4row[i2]]);
        }
        for (int k = tid; k < nsc; k += kLsapThreads) {
            const int j2 = w.sc_list[k];
            w.v[j2] = w.v[j2] - (min_val - w.spc[j2]);
        }
        __syncthreads();
        // augment the previous solution along the path
        if (tid == 0) {
            int j = sink;
            while (true) {
                const int i2 = w.path[j];
                w.row4col[j] = i2;
                const int tmp = w.col4row[i2];
                w.col4row[i2] = j;
                j = tmp;
                if (i2 == cur) break;
            }
        }
        __syncthreads();
    }
    // ---- dist_matr[row_ind, col_ind]: sum, mean, var (two passes like np.var; fixed reduction tree)
    auto block_sum = [&](double x) -> double {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        __syncthreads();
        if (lane == 0) s_red[warp] = x;
        __syncthreads();
        double t = 0.0;
        if (warp == 0) {
            t = s_red[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
            if (lane == 0) s_red[0] = t;
        }
        __syncthreads();
        return s_red[0];
    };
    double part = 0.0;
    for (int i = tid; i < nr; i += kLsapThreads) {
        const int j = w.col4row[i];
        const double c = j >= 0 ? w.cost[(size_t)i * nc + j] : 0.0;
        w.sel[i] = c;
        part += c;
    }
    const double sum = block_sum(part);
    const double mean = sum / (double)nr;
    part = 0.0;
    for (int i = tid; i < nr; i += kLsapThreads) {
        const double d = w.sel[i] - mean;
        part += d * d;
    }
    const double ss = block_sum(part);
    if (tid == 0) { w.out[0] = sum; w.out[1] = mean; w.out[2] = ss / (double)nr; w.out[3] = n_steps; }
}

}  // namespace aoslo

using namespace aoslo;

extern "C" int aoslo_hungarian(aoslo_ctx* ctx, void* stream, const double* d_gt, int n_gt, const double* d_particles,
                               int n, double* h_out4, double* d_assigned_cost) {
    if (!ctx || !d_gt || !d_particles || !h_out4 || n_gt < 1 || n < 1) return AOSLO_ERR_INVALID;
    cudaStream_t s = (cudaStream_t)stream;
    // dist_matr has one row per GT point; linear_sum_assignment transposes a tall matrix
    const bool transpose = n < n_gt;
    const int nr = transpose ? n : n_gt, nc = transpose ? n_gt : n;
    const double2* a = reinterpret_cast<const double2*>(transpose ? d_particles : d_gt);
    const double2* b = reinterpret_cast<const double2*>(transpose ? d_gt : d_particles);
    const size_t cells = (size_t)nr * nc;
    if (cells * sizeof(double) > ((size_t)16 << 30)) return AOSLO_ERR_CAPACITY;   // the reference's dense matrix
    unsigned char* pool = nullptr;
    const size_t off_cost = 0, off_u = off_cost + cells * 8, off_v = off_u + (size_t)nr * 8, off_spc = off_v + (size_t)nc * 8,
                 off_sel = off_spc + (size_t)nc * 8, off_out = off_sel + (size_t)nr * 8, off_path = off_out + 64,
                 off_c4r = off_path + (size_t)nc * 4, off_r4c = off_c4r + (size_t)nr * 4, off_rem = off_r4c + (size_t)nc * 4,
                 off_tk = off_rem + (size_t)nc * 4, off_srl = off_tk + (size_t)nc * 4, off_scl = off_srl + (size_t)nr * 4,
                 total = off_scl + (size_t)nc * 4;
    AOSLO_CUDA(cudaMalloc((void**)&pool, total));
    LsapWork w;
    w.nr = nr; w.nc = nc;
    w.cost = reinterpret_cast<double*>(pool + off_cost);
    w.u = reinterpret_cast<double*>(pool + off_u);
    w.v = reinterpret_cast<double*>(pool + off_v);
    w.spc = reinterpret_cast<double*>(pool + off_spc);
    w.sel = reinterpret_cast<double*>(pool + off_sel);
    w.out = reinterpret_cast<double*>(pool + off_out);
    w.path = reinterpret_cast<int*>(pool + off_path);
    w.col4row = reinterpret_cast<int*>(pool + off_c4r);
    w.row4col = reinterpret_cast<int*>(pool + off_r4c);
    w.remaining = reinterpret_cast<int*>(pool + off_rem);
    w.tkey = reinterpret_cast<int*>(pool + off_tk);
    w.sr_list = reinterpret_cast<int*>(pool + off_srl);
    w.sc_list = reinterpret_cast<int*>(pool + off_scl);
    w.status = ctx->d_status;
    int st = AOSLO_OK;
    dist_matrix_kernel<<<ctx->nb, kThreads, 0, s>>>(a, nr, b, nc, reinterpret_cast<double*>(pool + off_cost));
    count_launch();
    lsap_kernel<<<1, kLsapThreads, 0, s>>>(w);
    count_launch();
    { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) st = set_cuda_error(e, "lsap_kernel"); }
    double h[4] = {0, 0, 0, 0};
    if (st == AOSLO_OK && cudaMemcpyAsync(h, w.out, sizeof(h), cudaMemcpyDeviceToHost, s) != cudaSuccess) st = AOSLO_ERR_CUDA;
    if (st == AOSLO_OK && d_assigned_cost &&
        cudaMemcpyAsync(d_assigned_cost, w.sel, sizeof(double) * nr, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
        st = AOSLO_ERR_CUDA;
    if (cudaStreamSynchronize(s) != cudaSuccess) st = AOSLO_ERR_CUDA;
    cudaFree(pool);
    for (int k = 0; k < 4; ++k) h_out4[k] = h[k];
    return st;
}
