/*
 * aoslo_b200.h -- C ABI of libaoslo_b200.so: the B200 (sm_100a) implementation of the
 * cone-detection hot path of EgorovEV/Detecting_cones_AOSLO (`find_blobs`).
 *
 * Every entry point replaces one reference function (cited as file:line of the
 * reference tree).  The reference is pure Python, so the binding a maintainer adds
 * is a ctypes stub; the signatures use only plain pointers,
 * sizes and POD structs -- no torch / numpy types.
 *
 * Conventions
 *   - every function returns an aoslo_status (0 = ok, negative = error);
 *   - pointers named d_* are DEVICE pointers, h_* are HOST pointers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - particles are (n,2) float64 rows [x, y] exactly like the reference's ndarray;
 *   - blobness / gradient fields (every d_Rb, d_grad, and the d_field of aoslo_grad_field) are row-major
 *     (H,W) with a ROW PITCH of AOSLO_FIELD_PITCH(W) pixels -- W rounded up to a multiple of 4, so that
 *     rows start 16-byte aligned (the blobness kernel's TMA stores); the pad pixels are never read; the
 *     gradient field is (H,W,2) interleaved with [...,0] = d/drow and [...,1] = -d/dcol (reference
 *     utils.py:152-163), pitch AOSLO_FIELD_PITCH(W) pixel pairs.  The gravity field of aoslo_gravity_field
 *     / aoslo_add_particles is dense (H,W);
 *   - field element type is selected by `field_dtype`: AOSLO_F32 or AOSLO_F64;
 *   - device-side conditions (particle left the padded image, NaN position, capacity
 *     overflow, eigenvalue assert, too few points for kNN) are accumulated in the
 *     context's status word, read with aoslo_ctx_status().
 *   - there is NO CPU fallback: every compute entry point launches sm_100a kernels.
 */
#ifndef AOSLO_B200_H
#define AOSLO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    AOSLO_OK = 0,
    AOSLO_ERR_INVALID = -1,     /* bad argument                                   */
    AOSLO_ERR_CUDA = -2,        /* CUDA runtime error (see aoslo_last_cuda_error) */
    AOSLO_ERR_CAPACITY = -3,    /* particle capacity of the context exceeded      */
    AOSLO_ERR_UNSUPPORTED = -4  /* enumerated option not implemented              */
} aoslo_status;

/* bits of the device status word (aoslo_ctx_status) */
enum {
    AOSLO_DEV_ESCAPED = 1,      /* a particle index left [0,W)x[0,H): reference raises IndexError / wraps */
    AOSLO_DEV_NAN = 2,          /* NaN position (0/0 force on coincident particles, utils.py:123)          */
    AOSLO_DEV_CAPACITY = 4,     /* append beyond capacity                                                   */
    AOSLO_DEV_EIG_ASSERT = 8,   /* e0 > e1 violated somewhere (pipeline_with_delitions.py:43)              */
    AOSLO_DEV_KNN_SHORT = 16    /* fewer points than requested neighbours (sklearn ValueError)             */
};

enum { AOSLO_F32 = 0, AOSLO_F64 = 1 };
enum { AOSLO_FORMULA_CUSTOM = 0, AOSLO_FORMULA_SIMPLE_DIV = 1 };   /* pipeline_with_delitions.py:60-68 */
enum { AOSLO_GRAD_NP = 0, AOSLO_GRAD_SOBEL = 1 };                  /* utils.py:153-157 */
enum { AOSLO_STEP_DISCRETE = 0, AOSLO_STEP_CONTIN = 1 };           /* pipeline_with_delitions.py:219/:229 */
enum { AOSLO_ENERGY_PIT = 0, AOSLO_ENERGY_EXP = 1 };               /* utils.py:112 / :43 */
enum { AOSLO_TIE_CANONICAL = 0, AOSLO_TIE_SKLEARN = 1 };           /* kNN tie order, SURVEY.md 0-3 */

#define AOSLO_KMAX 8            /* max neighbours per query incl. self (dist_n_neighbours <= 7) */
#define AOSLO_MAX_SCALES 8
#define AOSLO_MAX_RADIUS 16
#define AOSLO_FIELD_PITCH(W) (((W) + 3) & ~3)   /* row pitch of R_b / gradient fields, in pixels */

typedef struct aoslo_ctx aoslo_ctx;

/* The reference's 29-key config dict (pipeline_with_delitions.py:411-436) reduced to
 * the keys that reach the numeric path, validated once on the host. */
typedef struct {
    int32_t H, W;                   /* padded image shape                                      */
    int32_t bx, by;                 /* boundary_size.x / .y                                    */
    double lambda_dist, lambda_blob;
    double dist_alpha, dist_sigma;  /* 'exp' energy                                            */
    int32_t dist_n_neighbours;
    int32_t epoch_num;
    int32_t metric_measure_freq;
    int32_t step_mode;              /* AOSLO_STEP_*                                            */
    int32_t dist_energy_type;       /* AOSLO_ENERGY_*                                          */
    double mu_dist_func, sigma_dist_func, lambda_dist_func;
    double threshhold_dist_del;     /* (sic) reference spelling                                */
    int32_t particle_call_window;
    double particle_call_threshold;
    double init_blob_threshold;
    int32_t reception_field_size;
    int32_t fix_particles_frequency;
    int32_t tie_mode;               /* AOSLO_TIE_*                                             */
    int32_t early_stop;             /* 1 = reference behaviour (:381-382), 0 = run all epochs  */
    double thinning_threshold;      /* 3.0 in the reference (:118)                             */
} aoslo_config;

/* One record per metrics evaluation (pipeline_with_delitions.py:320-359).  Integer
 * counts are exact; dice / precision / recall are derived from them by the caller. */
typedef struct {
    int32_t iteration;
    int32_t n_particles;
    int32_t n_stable;
    int32_t n_gt;
    int32_t tp[2], fp[2], fn[2];    /* thresholds 1.42 px, 2.83 px (utils.py:379)              */
    int32_t n_in_image;
    int32_t pad_;
    double sum_p2g, sum_g2p;        /* sum of nearest distances particle->GT, GT->particle      */
    double var_p2g, var_g2p;        /* np.var of each                                          */
    double blob_energy_sum;         /* sum R_b[int(y)][int(x)] (utils.py:137)                  */
} aoslo_metrics_record;

/* Summary of one whole detection run (aoslo_run). */
typedef struct {
    int32_t status;                 /* device status word                                      */
    int32_t n_iterations;           /* iterations executed (early stop counted)                */
    int32_t n_particles;            /* final count                                             */
    int32_t n_stable;
    int32_t n_seeds;
    int32_t n_after_thinning;
    int32_t n_thinning_rounds;
    int32_t n_records;
} aoslo_run_summary;

/* ---- library ------------------------------------------------------------------------ */
const char* aoslo_version(void);
const char* aoslo_last_cuda_error(void);
int aoslo_device_count(void);
/* number of kernels this library has launched in this process (monotonic counter) */
long long aoslo_launch_count(void);

/* host wait policy of `device` (cudaSetDeviceFlags): 0 auto, 1 spin, 2 yield, 3 blocking sync */
int aoslo_set_wait_policy(int device, int policy);

/* ---- context: owns all scratch (cell lists, kNN tables, scan spines, status word) ------ */
int aoslo_ctx_create(aoslo_ctx** out, int device, int H, int W, int particle_capacity, int gt_capacity);
int aoslo_ctx_destroy(aoslo_ctx* ctx);
int aoslo_ctx_status(aoslo_ctx* ctx, void* stream, int* h_status_out, int clear);
/* zeroes the status word, stream ordered (no host synchronisation) */
int aoslo_ctx_clear_status(aoslo_ctx* ctx, void* stream);
int aoslo_ctx_capacity(aoslo_ctx* ctx);
/* implementation knobs (defaults come from the environment once, at aoslo_ctx_create): "k1_tma" 0|1|2,
 * "pair_th", "pair_occ", "dtile_th", "dtile_occ" (0 = automatic), "dtile_split" 0|1,
 * "k1_kernel" 0 auto | 1 generic tile | 2 band | 3 f64 tile,
 * "no_graph" 1 = run aoslo_run's canonical path eagerly (debugging / profiling).  Results do not depend on them.
 * Environment only, read at aoslo_ctx_create: AOSLO_THIN_THREADS (128 | 256 | 512) and AOSLO_THIN_BLOCKS_PER_SM (1..4),
 * the shape of the cooperative thinning kernel; AOSLO_CELL_SHIFT / AOSLO_LINK_SHIFT, the cell edges of the searches. */
int aoslo_ctx_set_tuning(aoslo_ctx* ctx, const char* key, int value);
/* 1: host-side waits of this context sleep on a blocking-sync event instead of spinning (use when many
 * contexts run concurrently, one host thread each); 0 (default): lowest-latency spinning sync */
int aoslo_ctx_set_blocking_sync(aoslo_ctx* ctx, int enable);

/* ---- L1 pre-processing on the device: GT_enumerate_from_zero (utils.py:15-17) when gt_is_raw = 1, np.round of the
 * GT, crop_image (image_transformation.py:6-21; the reference's region asserts -> AOSLO_ERR_INVALID) and
 * mirror_padding_image (:23-35).  d_raw_img: (H0,W0) u8; d_img_out: (H,W) u8 with the given pitch,
 * H = (y_max-y_min) + 2*by, W = (x_max-x_min) + 2*bx (must equal the context's shape);
 * d_gt: (n_gt_raw,2) f64, gt_is_raw = 0: already shifted by -1 (the caller did the reference's in-place shift),
 * gt_is_raw = 1: the caller's 1-based array as uploaded by aoslo_stage_gt;
 * GT rows kept when strictly inside the region, order preserved; count -> *h_n_gt_out (one host synchronisation),
 * or h_n_gt_out = NULL: the count stays on the device for aoslo_run(..., n_gt = -1) and nothing synchronises. */
int aoslo_prepare_inputs(aoslo_ctx* ctx, void* stream, const uint8_t* d_raw_img, int H0, int W0, int pitch0_bytes,
                         int x_min, int x_max, int y_min, int y_max, int bx, int by, uint8_t* d_img_out,
                         int out_pitch_bytes, const double* d_gt, int n_gt_raw, double* d_gt_out,
                         int* h_n_gt_out, int gt_is_raw);

/* Upload of the caller's (n_gt_raw,2) f64 ground truth (pinned or pageable host memory) into d_gt_raw.
 * shift_in_place = 1: the reference shifts the CALLER's array by -1 (pipeline_with_delitions.py:26 -> utils.py:16);
 * that host pass (one read + write of the array) is deferred: the next aoslo_run of this context performs it after
 * it has enqueued its device work and before it waits (the upload has completed by then, checked with an event), so
 * it costs no latency.  aoslo_ctx_flush_host_work performs it immediately if no run picked it up (error paths);
 * aoslo_stage_gt itself flushes an older pending shift first.  The array must stay valid until then. */
int aoslo_stage_gt(aoslo_ctx* ctx, void* stream, double* h_gt, int n_gt_raw, double* d_gt_raw, int shift_in_place);
int aoslo_ctx_flush_host_work(aoslo_ctx* ctx);

/* ---- L2 field: replaces skimage.hessian_matrix + hessian_matrix_eigvals + the blobness
 * formula + utils.calc_grad_field (pipeline_with_delitions.py:36-38, :60-68, :91;
 * utils.py:152-163).  u8 image in, R_b and gradient out, one fused kernel.
 * h_weights: n_scales rows of AOSLO_MAX_RADIUS+1 half-kernel weights w[0..radius] (w[0] centre);
 * h_radius / h_sigma2: per scale.  d_grad may be NULL (R_b only). */
int aoslo_blobness(aoslo_ctx* ctx, void* stream, const uint8_t* d_img, int H, int W, int pitch_bytes,
                   const double* h_weights, const int* h_radius, const double* h_sigma2, int n_scales,
                   int formula, int grad_type, int field_dtype, void* d_Rb, void* d_grad);

/* utils.calc_grad_field on an arbitrary field (utils.py:152-163). */
int aoslo_grad_field(aoslo_ctx* ctx, void* stream, const void* d_field, int H, int W, int grad_type,
                     int field_dtype, void* d_grad);

/* ---- L3 particle system --------------------------------------------------------------- */
/* utils.initialize_high_prop_location (utils.py:359-375): raster-order seeds of the inner region. */
int aoslo_seed(aoslo_ctx* ctx, void* stream, const void* d_Rb, int field_dtype, int H, int W, int bx, int by,
               double threshold, double* d_particles, int capacity, int* h_n_out);

/* kNN of `d_query` rows against `d_points` (utils.py:29-31, :218-219).  Outputs (nq,k):
 * indices int32 and Euclidean distances float64, ordered like the reference's sklearn call
 * (tie_mode = AOSLO_TIE_SKLEARN) or by (d^2, index) (AOSLO_TIE_CANONICAL). */
int aoslo_knn(aoslo_ctx* ctx, void* stream, const double* d_points, int n_points, const double* d_query,
              int n_query, int k, int tie_mode, int32_t* d_idx_out, double* d_dist_out);

/* utils.del_close_particles (utils.py:245-278): greedy deletion in row order + order-preserving
 * compaction.  In: particles/stable (n).  Out: compacted particles/stable, per-input deleted mask. */
int aoslo_delete_close(aoslo_ctx* ctx, void* stream, const double* d_particles, const uint8_t* d_stable, int n,
                       const void* d_Rb, int field_dtype, int H, int W, double threshold, int tie_mode,
                       double* d_particles_out, uint8_t* d_stable_out, uint8_t* d_deleted_out, int* h_n_out);

/* utils.get_precomputed_dist_func (utils.py:95-109): rfs x rfs float64 LUT of energy_pit_func2. */
int aoslo_dist_lut(aoslo_ctx* ctx, void* stream, int rfs, double mu, double sigma, double lambda_, double* d_lut);

/* utils.calc_gravity_field_optim (utils.py:310-333): (H,W) float64 field. */
int aoslo_gravity_field(aoslo_ctx* ctx, void* stream, const double* d_particles, int n, const double* d_lut,
                        int rfs, int H, int W, double* d_field);

/* utils.adding_particles (utils.py:336-356): appends to d_particles/d_stable at row n. */
int aoslo_add_particles(aoslo_ctx* ctx, void* stream, const double* d_field, int H, int W, int bx, int by,
                        int window, double threshold, double* d_particles, uint8_t* d_stable, int n,
                        int capacity, int* h_n_out);

/* utils.calc_dist_energy_pit_optim / calc_dist_energy (utils.py:132-134, :112-129, :43-58):
 * (n,2) float64 force.  For the 'pit' energy stable rows are 0 (d_stable may be NULL = none). */
int aoslo_dist_force(aoslo_ctx* ctx, void* stream, const double* d_particles, const uint8_t* d_stable, int n,
                     int n_neighbours, int energy_type, double p0, double p1, double p2, int tie_mode,
                     double* d_force);

/* the move step (pipeline_with_delitions.py:215-245), in place on d_particles. */
int aoslo_move(aoslo_ctx* ctx, void* stream, double* d_particles, const uint8_t* d_stable, int n,
               const void* d_grad, int field_dtype, int H, int W, const double* d_force, double lambda_blob,
               double lambda_dist, int step_mode);

/* ---- L4 metrics: utils.calc_dist_to_neares x2 + get_particles_in_image + calc_acc_metrics +
 * calc_metrics('Chamfer') + calc_blob_energy (utils.py:214-242, :378-418, :137-144).
 * d_dist_p2g (n) / d_dist_g2p (n_gt) may be NULL. */
int aoslo_metrics(aoslo_ctx* ctx, void* stream, const double* d_particles, const uint8_t* d_stable, int n,
                  const double* d_gt, int n_gt, const void* d_Rb, int field_dtype, int H, int W, int bx, int by,
                  aoslo_metrics_record* h_record_out, double* d_dist_p2g, double* d_dist_g2p);

/* classification arrays of utils.visualize_all (utils.py:174-203) from the two nearest-distance arrays of
 * aoslo_metrics: per GT point TP (d < threshold) / FN, per particle FP (d > threshold and strictly inside
 * the original image).  Plotting stays on the host. */
int aoslo_classify(aoslo_ctx* ctx, void* stream, const double* d_dist_g2p, int n_gt, const double* d_dist_p2g,
                   const double* d_particles, int n, int bx, int by, int H, int W, double threshold,
                   uint8_t* d_gt_is_tp, uint8_t* d_particle_is_fp);

/* ---- calc_metrics(mode='hangarian') (utils.py:228-234): scipy.spatial.distance_matrix(GT, particles) +
 * scipy.optimize.linear_sum_assignment restated step by step (same optimal assignment, incl. scipy's tie rule) +
 * sum / mean / var of the assigned distances.  h_out4: sum, mean, var, number of search steps (diagnostic).
 * d_assigned_cost (may be NULL): min(n_gt, n) assigned distances in the order of the smaller point set.
 * Synchronous; allocates the reference's dense n_gt x n matrix for the duration of the call. */
int aoslo_hungarian(aoslo_ctx* ctx, void* stream, const double* d_gt, int n_gt, const double* d_particles, int n,
                    double* h_out4, double* d_assigned_cost);

/* ---- one loop iteration: distance force + move + deletion (pipeline_with_delitions.py:173-253),
 * the code path aoslo_run executes per epoch; out-of-place. */
int aoslo_iterate(aoslo_ctx* ctx, void* stream, const aoslo_config* cfg, const void* d_Rb, const void* d_grad,
                  int field_dtype, const double* d_particles, const uint8_t* d_stable, int n,
                  double* d_particles_out, uint8_t* d_stable_out, int* h_n_out);

/* ---- L5 whole run: pipeline_with_delitions.find_blobs (:12-403) after crop/pad.
 * Seeds, thins, resamples and runs the epoch loop with every stage on the device; returns the
 * per-evaluation metric records, the final particles and a summary.  Stage timings (ms) in
 * h_stage_ms[5] (may be NULL): dist_energy_calc, moving_particles, deleting_particles, fix_resample,
 * metrics_logging (the reference's `time_spended` keys, :165-171).
 * AOSLO_TIE_CANONICAL: the whole run is one cached CUDA graph (seeds + thinning to the fixpoint in one cooperative
 * kernel on a bitmap of alive pixels; conditional WHILE / IF nodes for the early stop and the resample / metrics
 * branches); the host synchronises once, at the end.
 * All value parameters of `cfg` are uploaded per call, so runs that differ only in values (a sweep) share the
 * graph; d_Rb, d_grad, d_gt must be stable addresses for the graph to be reused.
 * AOSLO_TIE_SKLEARN: stage calls with one host round trip per kNN (host-built tree).
 * n_gt = -1 (canonical tie order only): the GT count is the device scalar the last aoslo_prepare_inputs of this
 * context left behind (d_gt must be its d_gt_out); the count comes back in the records' n_gt field.
 * summary.status = the context's status word at the end of the run; it is NOT cleared at the start, so a condition
 * raised by aoslo_blobness on the same stream (AOSLO_DEV_EIG_ASSERT) is reported with the run's own. */
int aoslo_run(aoslo_ctx* ctx, void* stream, const aoslo_config* cfg, const void* d_Rb, const void* d_grad,
              int field_dtype, const double* d_gt, int n_gt, aoslo_metrics_record* h_records, int max_records,
              double* d_particles_out, uint8_t* d_stable_out, aoslo_run_summary* h_summary, float* h_stage_ms);

/* ---- L6 sweep: `n_trials` configurations on ONE blobness field (replaces the per-trial body of the wandb agent loop,
 * main_with_delitions.py:81-132: every trial of the sweep re-runs find_blobs on the same image, and no swept
 * parameter changes the field).  Trials run back to back on `stream` (one aoslo_run each: in the canonical tie order
 * trials that differ only in value parameters share one cached CUDA graph).  Per trial: the run summary, the last
 * metrics record and the best-result bookkeeping of find_blobs' non-wandb branch (:335-359: dice with the 0.5 %
 * hysteresis, min Chamfer mean, max mean blob energy).  A trial that fails (argument error, device condition) gets
 * its code in `error` / `summary.status` and does not stop the sweep (the reference's driver catches and continues,
 * :125-132).  Run several contexts on several streams for concurrency (sweep.UnitRunner). */
typedef struct {
    int32_t error;                  /* aoslo_status of the trial's aoslo_run (0 = ok)          */
    int32_t pad_;
    aoslo_run_summary summary;
    aoslo_metrics_record last;      /* last metrics evaluation of the trial                    */
    double dice_coeff_1, dice_coeff_2;          /* result['dice_coeff_1' / '2']                */
    double best_lsa_mean, best_blobEn_mean;     /* result['best_lsa_mean' / 'best_blobEn_mean'] */
} aoslo_trial_result;

int aoslo_sweep(aoslo_ctx* ctx, void* stream, const aoslo_config* h_cfgs, int n_trials, const void* d_Rb,
                const void* d_grad, int field_dtype, const double* d_gt, int n_gt, aoslo_trial_result* h_results);

/* final particles / stable flags of the last aoslo_run on this context (the first n rows, n <= summary.n_particles),
 * copied device -> device on `stream`: lets the caller size its output after the run instead of passing
 * capacity-sized buffers to aoslo_run (whose d_particles_out / d_stable_out may be NULL). */
int aoslo_read_particles(aoslo_ctx* ctx, void* stream, double* d_particles_out, uint8_t* d_stable_out, int n);

/* ---- host-side input readers (replace cv2.imread(path, -1) / pd.read_csv(path).to_numpy() of the reference's
 * drivers, pipeline_with_delitions.py:443-446, main_with_delitions.py:153-155); they decode into caller buffers
 * (pinned host memory in the batched loader).  TIFF subset: baseline 8-bit grayscale strips, uncompressed or
 * PackBits, either byte order; anything else returns AOSLO_ERR_UNSUPPORTED. */
int aoslo_tiff_info(const char* path, int* h_H, int* h_W, int* h_bits, int* h_samples);
int aoslo_tiff_read_gray8(const char* path, uint8_t* h_out, int out_pitch_bytes);
/* "x,y" per line.  skip_first_row = 1: the reference's pd.read_csv default, which consumes line 1 of the header-less
 * annotation file as column names.  *h_n_out = points in the file (AOSLO_ERR_CAPACITY if more than `capacity`). */
int aoslo_csv_read_points(const char* path, int skip_first_row, double* h_xy_out, int capacity, int* h_n_out);

/* ---- host helper used by the sklearn-order kNN: builds sklearn's KDTree(leaf_size=1)
 * idx_array / node table on the HOST (std::nth_element, sklearn/neighbors/_partition_nodes.pyx). */
int aoslo_kdtree_build_host(const double* h_points, int n, int32_t* h_idx_array, int32_t* h_node_start,
                            int32_t* h_node_end, double* h_node_bounds, int n_nodes);
int aoslo_kdtree_num_nodes(int n);

#ifdef __cplusplus
}
#endif
#endif /* AOSLO_B200_H */
