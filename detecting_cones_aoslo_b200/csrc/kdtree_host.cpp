// kdtree_host.cpp -- host-side builder of scikit-learn's KDTree(leaf_size=1) layout.
//
// The reference's neighbour search is sklearn.neighbors.NearestNeighbors(algorithm='kd_tree',
// leaf_size=1) (reference utils.py:29, :218).  Particles live on the integer lattice, so exact
// distance ties are the norm and the answer depends on sklearn's traversal order, which depends on
// the tree's idx_array -- the residue of libstdc++ std::nth_element with sklearn's
// IndexComparator (sklearn/neighbors/_partition_nodes.pyx) applied recursively
// (sklearn/neighbors/_binary_tree.pxi.tp:1033-1083).  That residue is reproduced by calling the same
// std::nth_element here; the device then walks the tree (particles.cu: knn_sklearn_kernel).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <thread>
#include <vector>

#include "aoslo_b200.h"

namespace {

struct IndexComparator {
    const double* data;
    int split_dim;
    bool operator()(const int32_t& a, const int32_t& b) const {
        double av = data[(size_t)a * 2 + split_dim];
        double bv = data[(size_t)b * 2 + split_dim];
        return av == bv ? a < b : av < bv;
    }
};

struct Builder {
    const double* data;
    int32_t* idx;
    int32_t* node_start;
    int32_t* node_end;
    double* bounds;
    int n_nodes;

    // depth < kParallelDepth: the two subtrees (disjoint ranges of idx, disjoint nodes) are built by two threads --
    // the result is the same sequence of std::nth_element calls on the same ranges, only not one after the other
    int parallel_depth = 3;              // 8 threads; 16 on hosts with >= 16 hardware threads (set by the caller)
    static constexpr int kParallelMinPoints = 2048;

    void build(int i_node, int idx_start, int idx_end, int depth = 0) {
        // init_node (sklearn/neighbors/_kd_tree.pyx.tp:72-104): bounds = min/max of the members
        double lo0 = INFINITY, lo1 = INFINITY, hi0 = -INFINITY, hi1 = -INFINITY;
        for (int i = idx_start; i < idx_end; ++i) {
            const double* r = data + (size_t)idx[i] * 2;
            lo0 = std::fmin(lo0, r[0]); hi0 = std::fmax(hi0, r[0]);
            lo1 = std::fmin(lo1, r[1]); hi1 = std::fmax(hi1, r[1]);
        }
        double* b = bounds + (size_t)i_node * 4;
        b[0] = lo0; b[1] = lo1; b[2] = hi0; b[3] = hi1;
        node_start[i_node] = idx_start;
        node_end[i_node] = idx_end;
        if (2 * i_node + 1 >= n_nodes) return;          // leaf
        if (idx_end - idx_start < 2) return;            // "too many nodes allocated" guard: leaf
        int n_points = idx_end - idx_start;
        int n_mid = n_points / 2;
        // find_node_split_dim (:600-644): first dimension with the strictly largest spread; max - min of the
        // members per dimension is exactly the bounds just computed (fmax / fmin of the same values)
        int j_max = 0;
        double max_spread = 0;
        const double spread[2] = {hi0 - lo0, hi1 - lo1};
        for (int j = 0; j < 2; ++j)
            if (spread[j] > max_spread) { max_spread = spread[j]; j_max = j; }
        IndexComparator cmp{data, j_max};
        std::nth_element(idx + idx_start, idx + idx_start + n_mid, idx + idx_end, cmp);
        if (depth < parallel_depth && n_points >= kParallelMinPoints) {
            std::thread left([=] { build(2 * i_node + 1, idx_start, idx_start + n_mid, depth + 1); });
            build(2 * i_node + 2, idx_start + n_mid, idx_end, depth + 1);
            left.join();
        } else {
            build(2 * i_node + 1, idx_start, idx_start + n_mid, depth + 1);
            build(2 * i_node + 2, idx_start + n_mid, idx_end, depth + 1);
        }
    }
};

}  // namespace

extern "C" int aoslo_kdtree_num_nodes(int n) {
    if (n < 1) return 0;
    // n_levels = int(log2(max(1, (n - 1) / leaf_size)) + 1), leaf_size = 1 (:875-877)
    double v = std::fmax(1.0, (double)(n - 1));
    int n_levels = (int)(std::log2(v) + 1);
    return (1 << n_levels) - 1;
}

extern "C" int aoslo_kdtree_build_host(const double* h_points, int n, int32_t* h_idx_array, int32_t* h_node_start,
                                       int32_t* h_node_end, double* h_node_bounds, int n_nodes) {
    if (!h_points || !h_idx_array || !h_node_start || !h_node_end || !h_node_bounds || n < 1) return AOSLO_ERR_INVALID;
    if (n_nodes != aoslo_kdtree_num_nodes(n)) return AOSLO_ERR_INVALID;
    for (int i = 0; i < n; ++i) h_idx_array[i] = i;
    for (int i = 0; i < n_nodes; ++i) {
        h_node_start[i] = 0; h_node_end[i] = 0;
        h_node_bounds[4 * (size_t)i + 0] = 0; h_node_bounds[4 * (size_t)i + 1] = 0;
        h_node_bounds[4 * (size_t)i + 2] = 0; h_node_bounds[4 * (size_t)i + 3] = 0;
    }
    Builder b{h_points, h_idx_array, h_node_start, h_node_end, h_node_bounds, n_nodes};
    static const unsigned hw = std::thread::hardware_concurrency();
    b.parallel_depth = hw >= 16 ? 4 : 3;
    b.build(0, 0, n);
    return AOSLO_OK;
}
