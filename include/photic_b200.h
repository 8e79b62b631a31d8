/*
 * photic_b200.h -- C ABI of libphotic_b200.so: the B200 (sm_100a, FP64) replacement for the
 * per-pixel semi-analytical inversion of stblake/photic.
 *
 * Drop-in boundary. The reference's entry point for this path is
 *
 *     void samodel(scene scene_data[], geogrid gridded_data[], int *scene_indexes, int nscenes,
 *                  bool empirical_depth_present, geogrid empirical_depths,
 *                  int n_smoothing_radius, int n_spatial, int n_bottoms,
 *                  float **depth, ... float **index_optical_depth, ...);      model/samodel.h:8-19
 *
 * called once from model/bam.c:3236-3241. photic_b200/host/samodel_b200.c keeps that exact symbol
 * and signature (it replaces model/samodel.c in the reference's link line) and forwards to the
 * functions below. Everything here is extern "C", plain pointers and sizes, no torch/C++ types.
 * All functions return 0 on success or a PHB_E* code; phb_error_string() describes it (the
 * reference itself has no error channel: it printf()s and exit(1)s, model/common.h:62-67 -- the
 * shim reproduces that).
 *
 * There is NO CPU fallback: every compute entry point fails with PHB_ENODEVICE when no CUDA
 * device is usable.
 */
#ifndef PHOTIC_B200_H_
#define PHOTIC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHB_MAX_SCENES 16  /* acquisition dates per inversion           */
#define PHB_MAX_BANDS 8    /* spectral bands per scene                  */
#define PHB_MAX_BOTTOMS 8  /* substrate classes (table of samodel.c:464) */
#define PHB_MAX_SPATIAL 3  /* NSPATIAL: regions = (2*NSPATIAL-1)^2 <= 25 */

enum {
  PHB_OK = 0,
  PHB_EINVAL = 1,    /* bad descriptor / argument                         */
  PHB_ENODEVICE = 2, /* no usable CUDA device (there is no CPU fallback)   */
  PHB_ECUDA = 3,     /* a CUDA call failed; see phb_error_string           */
  PHB_ENOMEM = 4,    /* host or device allocation failed                   */
  PHB_ENOFIT = 5,    /* phb_jerlov_*: the reference's `return false` paths  */
  PHB_ENOPEER = 6    /* a peer band cannot be mapped (no peer access between the devices / IPC refused) */
};

/*
 * What SCENE ... / SET NSMOOTH|NSPATIAL|NBOTTOMS leave in the reference's globals
 * (scene: model/common.h:194-218; defaults model/common.c:152), flattened to a POD.
 * Only the fields samodel() reads are present.
 */
typedef struct phb_scene_desc {
  int32_t n_scenes;
  int32_t n_bands[PHB_MAX_SCENES];
  int32_t wavelengths[PHB_MAX_SCENES][PHB_MAX_BANDS]; /* integer nm: scene.wavelengths is int[] */
  double theta_view[PHB_MAX_SCENES];                  /* degrees, scene.theta_v */
  double theta_sun[PHB_MAX_SCENES];                   /* degrees, scene.theta_w */
  double h_tide[PHB_MAX_SCENES];                      /* metres,  scene.H_tide  */
  int32_t n_smoothing_radius;                         /* SET NSMOOTH  (default 1) */
  int32_t n_spatial;                                  /* SET NSPATIAL (default 2) */
  int32_t n_bottoms;                                  /* SET NBOTTOMS (default 3) */
  int32_t nrows, ncols;                               /* raster held by THIS call (shard incl. halo rows) */
  float nodata;                                       /* geogrid.nodata_value of the reflectance grids */
  int32_t prior_present;                              /* MODEL ... DEPTHS grid given (bam.c:3077)   */
  float prior_nodata;
  double r_sigma[PHB_MAX_SCENES][PHB_MAX_BANDS];      /* scene.R_sigma (SCENE ... RSIGMA, bam.c:1474): only the
                                                         depth-error trials read it (samodel.c:3004-3006)  */
  int32_t nodata_per_band;                            /* != 0: nodata_band[][] holds every grid's own nodata_value --
                                                         the reference tests band k against gridded_data[k].nodata_value
                                                         (samodel.c:683, 941, 2999-3003); 0: all grids share `nodata` */
  float nodata_band[PHB_MAX_SCENES][PHB_MAX_BANDS];
} phb_scene_desc;

/*
 * Output planes, each [nrows][ncols] float32 unless noted; any pointer may be NULL (not wanted).
 * Cell values follow samodel(): defaults of samodel.c:819-829 where nothing is inverted, depth
 * NEGATED (samodel.c:1486-1490), model_error holds the Rrs error (samodel.c:1121).
 * Rows outside [row_begin,row_end) are left untouched.
 */
typedef struct phb_outputs {
  float *depth, *model_error, *bottom_albedo, *bottom_sand, *bottom_seagrass, *bottom_coral;
  float *K_min, *bottom_type, *index_optical_depth;
  float *K;          /* [n_scenes][max_bands][nrows][ncols]  (<scene>_K_<band>.nc, samodel.c:1513-1560) */
  float *P, *G, *X;  /* [n_scenes][nrows][ncols]             (<scene>_{P,G,X}.nc, samodel.c:1562-1596) */
  uint8_t *converged; /* md->converged  (samodel.c:2398-2402)                                   */
  int32_t *n_evals;   /* md->n_iterations = nelmin icount (samodel.c:2396)                      */
} phb_outputs;

/* Counters of one inversion call (for throughput / roofline accounting). */
typedef struct phb_stats {
  int64_t n_valid;      /* pixels inverted                                               */
  int64_t n_shallow;    /* of which with all NBOTTOMS substrates (h_prior <= 8 m)        */
  int64_t n_evals;      /* objective evaluations, incl. the 2 outside nelmin per start   */
  int64_t n_iters;      /* Nelder-Mead iterations                                        */
  int64_t n_converged;  /* pixels with ifault == 0                                       */
  double alg_flops;     /* sum_px evals*F_eval(Nr,Ns,Nb) + iters*(n^2+9n): the reference's literal op count */
  float ms_classify, ms_solve; /* CUDA-event times of the two kernels on the launch stream */
  float ms_h2d, ms_d2h;        /* host entry point only                                      */
  int32_t warps_per_cta, ctas, smem_bytes, regs; /* launch geometry actually used          */
} phb_stats;

typedef struct phb_ctx phb_ctx; /* one per device; owns scratch, queues, simplex slabs */

int phb_version(void);
const char *phb_error_string(int code);
int phb_device_count(void);

int phb_ctx_create(int device, phb_ctx **out);
void phb_ctx_destroy(phb_ctx *ctx);

/*
 * Scene-level constants samodel() derives once (samodel.c:505-618): a0, a1, a_w, b_bw and the
 * bottom spectra interpolated at each band, secants, a_w(640). Host-only, no device needed.
 * out layout: [n_scenes][PHB_MAX_BANDS][4 + PHB_MAX_BOTTOMS] then a_w640, sec_view[], sec_sun[].
 */
int phb_band_tables(const phb_scene_desc *desc, double *out_tables, double *out_aux);

/*
 * Inversion of rows [row_begin,row_end) of the raster, inputs and outputs in DEVICE memory.
 *   d_planes : [sum_s n_bands[s]][nrows][ncols] float32, scene-major then band
 *   d_prior  : [nrows][ncols] float32 DEPTHS grid (negative-down) or NULL
 *   d_debug  : optional [n_valid_capacity][phb_debug_record_len()] doubles + int32 pixel index
 *              per record, full-precision per-pixel results for parity tests; NULL in production
 * Edge clamping of the (2*n_spatial-1)^2 neighbourhood happens at the raster's own edges
 * (samodel.c:2971-2989): give a shard its halo rows and the result equals the unsharded one.
 * Asynchronous on `stream` (a cudaStream_t) except for the final stats read-back.
 */
int phb_invert_device(phb_ctx *ctx, const phb_scene_desc *desc, const float *d_planes, const float *d_prior,
                      int row_begin, int row_end, const phb_outputs *d_out, void *stream, phb_stats *stats);

/* Same with HOST buffers: copies in (through a pinned staging ring, so pageable caller memory streams at PCIe speed),
 * inverts, copies out. */
int phb_invert_host(phb_ctx *ctx, const phb_scene_desc *desc, const float *const *h_planes, const float *h_prior,
                    int row_begin, int row_end, const phb_outputs *h_out, phb_stats *stats);

/*
 * One process, several GPUs: what the reference's OpenMP team is to its cores. The scene is cut into
 * contiguous row bands, one per context -- one context per device -- each on its own host thread, which copies its
 * rows plus (n_spatial-1)+(n_smoothing_radius-1) halo rows straight from the caller's host rasters to its device and
 * its rows of the results back. Pixels are independent given the read-only halo, so the result equals phb_invert_host()
 * on one device bit for bit. Load balance: where the devices can map each other's memory (NVLink / NVSwitch peer
 * access) the bands are EQUAL row counts and the solve kernels share the work at run time -- a device whose own queue is
 * empty takes pixels from its neighbours' queues, reads their neighbourhoods from the owner's planes and stores the
 * results into the owner's planes over NVLink (phb_shard_* below) -- as the reference's `omp schedule(dynamic)` does
 * for cores (samodel.c:902). Without peer access the bands are planned by estimated cost instead (valid pixels weighted
 * by the depth bin of their DEPTHS prior; phb_plan_row_bands) and every device works on its own band only.
 *   ctxs      n_ctx contexts from phb_ctx_create (normally one per device; several on one device also work)
 *   stats     sums over the devices, times = the slowest device; per_ctx (nullable) [n_ctx]: what each DEVICE did
 *             (incl. pixels taken from neighbours); edges_out (nullable) [n_ctx + 1]: band k = rows [edges[k], edges[k+1])
 * phb_plan_row_bands is the planner on its own (host only, no device): edges [n_parts + 1], row_cost (nullable) [nrows].
 */
int phb_plan_row_bands(const phb_scene_desc *desc, const float *const *h_planes, const float *h_prior, int n_parts,
                       int32_t *edges, double *row_cost);
int phb_invert_host_multi(phb_ctx *const *ctxs, int n_ctx, const phb_scene_desc *desc, const float *const *h_planes,
                          const float *h_prior, const phb_outputs *h_out, phb_stats *stats, phb_stats *per_ctx,
                          int32_t *edges_out);

/*
 * The same two entry points for rasters held the way the reference holds them: `float **array`, one malloc per row
 * (geogrid.array, common.c:562-572). Rows are copied straight from / to the caller's row pointers through the pinned
 * staging ring -- no packed intermediate copy of the scene on the host (Pilbara: 26 GB of reflectance planes).
 *   plane_rows [sum_s n_bands[s]][nrows] row pointers, scene-major then band; prior_rows [nrows] or NULL
 *   out        row pointers per result grid; any member may be NULL
 */
typedef struct phb_row_outputs {
  float *const *depth, *const *model_error, *const *bottom_albedo, *const *bottom_sand, *const *bottom_seagrass,
      *const *bottom_coral, *const *K_min, *const *bottom_type, *const *index_optical_depth; /* [nrows] */
  float *const *const *K;                           /* [n_scenes * max_bands][nrows] */
  float *const *const *P, *const *const *G, *const *const *X; /* [n_scenes][nrows] */
} phb_row_outputs;
int phb_invert_rows(phb_ctx *const *ctxs, int n_ctx, const phb_scene_desc *desc, const float *const *const *plane_rows,
                    const float *const *prior_rows, const phb_row_outputs *out, phb_stats *stats, phb_stats *per_ctx,
                    int32_t *edges_out);

/*
 * A row band of a scene resident on one device, shareable with the other devices of the box. This is
 * the device-resident form of the multi-GPU path: one band per device -- one process per device (torch.distributed) or
 * all in one process -- each exported as a POD handle (CUDA IPC across processes, peer access inside one) that the
 * other devices open. phb_shard_solve() then runs the persistent solve kernel over the device's OWN queue first and
 * its peers' queues after: work moves between devices pixel by pixel over NVLink, with no collective on the data path
 * and no cost model.
 *   create    desc->nrows = rows of the band's raster INCLUDING its halo rows; [row_begin,row_end) = the rows of that
 *             raster this band owns (are inverted / carry results). scene_planes != 0: K/P/G/X planes too.
 *   buffers   device pointers of the band's own rasters: the caller fills d_planes [SB][nrows][ncols] and d_prior
 *             (NULL when desc->prior_present == 0) before prepare, and reads d_out after solve.
 *   prepare   validity scan, output defaults, work queues (asynchronous on `stream`).
 *   solve     peers [n_peers]: handles of the other bands in the order this device should take work from them
 *             (exclude its own). EVERY band must have finished prepare before ANY solve starts, and results are complete
 *             only after EVERY device's solve has finished: the caller synchronises (barrier / events) at both points.
 *             stats: what THIS device computed (n_valid counts the pixels it inverted, its own or its neighbours').
 * PHB_ENOPEER when a peer cannot be mapped; the caller then runs the bands independently (n_peers = 0).
 */
typedef struct phb_shard phb_shard;
typedef struct phb_shard_handle { unsigned char bytes[512]; } phb_shard_handle;
int phb_shard_create(phb_ctx *ctx, const phb_scene_desc *desc, int row_begin, int row_end, int scene_planes,
                     phb_shard **out);
int phb_shard_buffers(phb_shard *s, float **d_planes, float **d_prior, phb_outputs *d_out);
int phb_shard_export(phb_shard *s, phb_shard_handle *h);
int phb_shard_prepare(phb_shard *s, void *stream);
int phb_shard_solve(phb_shard *s, const phb_shard_handle *peers, int n_peers, void *stream, phb_stats *stats);
int64_t phb_shard_valid(phb_shard *s); /* valid pixels of the band's own rows (after prepare; synchronises) */
void phb_shard_destroy(phb_shard *s);

/*
 * Depth-error estimate, the last phase of samodel() (samodel.c:1376-1477): for every 0.25 m depth interval
 * up to min(floor(max depth), 30 m), `n_samples` (128 in the reference) pixels are drawn at random inside the
 * interval (up to sqrt(nrows*ncols) probes each), re-inverted with every reflectance shifted by
 * n_sigma * R_sigma[band], n_sigma ~ U(-1, 1), hot-started from the previous trial's P, G, X; the interval's
 * sigma is the population standard deviation of the trial depths and every cell gets the sigma of its interval.
 * The draws are libc rand() calls in the reference's order (random_in_range common.c:527, frand2 common.c:220);
 * the reference seeds with time(NULL) (samodel.c:371), here the seed is an argument, so the reference run with
 * srand(seed) produces the same table (tests pin that against oracle/_ref).
 *   h_depth     [nrows][ncols] depth plane as phb_invert_* leaves it (NEGATED, -0 where nothing was inverted)
 *   chain_mode  PHB_SIGMA_CHAIN_REFERENCE: one hot-start chain through all trials, as the reference runs them
 *               (inherently serial: one warp works, ~10 ms per trial);
 *               PHB_SIGMA_CHAIN_PER_INTERVAL: the chain restarts cold at every depth interval, intervals run
 *               in parallel (the shim's PHOTIC_B200_SIGMA_CHAIN=interval; equals the reference modified the same way)
 *   h_depth_sigma [nrows][ncols] out; table (nullable) [PHB_SIGMA_MAX_INTERVALS] out; trials (nullable)
 *               [PHB_SIGMA_MAX_INTERVALS][n_samples] out: the trial depths (0 = no pixel found / no prior)
 * A DEPTHS prior is required (PHB_EINVAL otherwise): without one the reference starts a hot trial from
 * md->depth_prev, which depends on the order its pixel loop happened to run in.
 */
#define PHB_SIGMA_MAX_INTERVALS 120
#define PHB_SIGMA_CHAIN_REFERENCE 0
#define PHB_SIGMA_CHAIN_PER_INTERVAL 1
int phb_depth_sigma_host(phb_ctx *ctx, const phb_scene_desc *desc, const float *const *h_planes, const float *h_prior,
                         const float *h_depth, unsigned seed, int n_samples, int chain_mode, int max_intervals,
                         float *h_depth_sigma, double *table, int32_t *n_intervals, double *trials, phb_stats *stats);

/* The same for rasters held as row pointers (geogrid.array): what the samodel() shim calls. Called directly after
 * phb_invert_rows() on one device with the same plane_rows array it reuses the rasters that call left on the device. */
int phb_depth_sigma_rows(phb_ctx *ctx, const phb_scene_desc *desc, const float *const *const *plane_rows,
                         const float *const *prior_rows, const float *const *depth_rows, unsigned seed, int n_samples,
                         int chain_mode, int max_intervals, float *const *sigma_rows, double *table, int32_t *n_intervals,
                         double *trials, phb_stats *stats);

/* Parity-test hook: like phb_invert_host but also returns the full-precision per-pixel record
 * (layout of oracle/ref_harness.c: 16 + n_scenes*max_bands + 3*n_scenes doubles) for every
 * inverted pixel. rec: [capacity][reclen]; pix: [capacity] linear index i*ncols+j; n_iters (nullable): [capacity][3]
 * = nelmin's icount of the best H start (md->n_iterations), converged | iterations << 1 of that start, and nelmin's
 * restart count `numres` (asa047.c:493) summed over the H starts of the pixel. */
int phb_debug_record_len(const phb_scene_desc *desc);
int phb_invert_host_debug(phb_ctx *ctx, const phb_scene_desc *desc, const float *const *h_planes,
                          const float *h_prior, int row_begin, int row_end, const phb_outputs *h_out,
                          double *rec, int32_t *pix, int32_t *n_iters, int64_t capacity, phb_stats *stats);

/* Test hook (host only): the scene-level constants block the kernels receive (struct ModelConst of
 * csrc/device_model.cuh), for the CPU emulation of the solve kernel in tests/emu. Returns the block's size in
 * bytes (-1 for a bad descriptor); fills `out` when capacity suffices. */
int64_t phb_debug_model_const(const phb_scene_desc *desc, void *out, int64_t capacity);

/* Known-answer hooks (device evaluation of the forward model / objective on caller-supplied
 * parameter vectors, mirroring oracle/ref_harness.c:ref_error_kat) and of the exact libm port. */
int phb_kat_objective(phb_ctx *ctx, const phb_scene_desc *desc, int n_bottoms_active, int n_regions, int origin,
                      const double *rrs_measured, int nparams, int nvec, const double *params, double *out6);
/* Mapping study: closed-loop objective evaluations on every SM, 16 warps per SM, one pixel per
 * team of `team_warps` warps -- 1 is the product's mapping (one warp per pixel), 2 / 4 / 8 spread the forward-model terms
 * of one pixel over that many warps and leave the ordered sum and the penalties to one of them. same_smsp != 0 puts the
 * warps of a team on one scheduler; skew_cycles > 0 starts team t of an SM t * skew_cycles late, so that the teams stay
 * out of phase as the warps of the solve kernel are (0: in step). Returns the value of the first evaluation
 * (bit-identical between mappings), the rate and the kernel time. */
int phb_eval_bench(phb_ctx *ctx, const phb_scene_desc *desc, int n_bottoms_active, int n_regions, int origin,
                   const double *rrs_measured, int nparams, const double *params, int team_warps, int same_smsp,
                   int skew_cycles, int reps, double *first_value, double *evals_per_s, float *ms);
int phb_kat_math(phb_ctx *ctx, int fn /*0 exp,1 log,2 pow,3 fast_div,4 fast_sqrt,5 a/b,6 sqrt,7 exp_main,8 x/pi fast,9-11 range predicates,12 log10,13/15 guarded a/b,14 guarded sqrt*/, const double *x,
                 const double *y, int64_t n, double *out);

/*
 * REFINE (model/refine.c:12-302): point-wise depth remap.
 * args: [0] clip_min [1] clip_max [2] scale_min [3] scale_max [4] shape [5] linear_m [6] linear_c
 *       [7] scrap_min [8] scrap_max [9] power_a [10] power_b
 * Without PHB_REFINE_CLIP the grid's own min/max (nodata skipped) are used (refine.c:215-225);
 * minmax_io (2 floats, nullable) lets a sharded caller all-reduce them: in = local, out = used.
 */
#define PHB_REFINE_CLIP 1
#define PHB_REFINE_SCALE 2
#define PHB_REFINE_LINEAR 4
#define PHB_REFINE_SCRAP 8
#define PHB_REFINE_POWER 16
int phb_refine_minmax_device(phb_ctx *ctx, const float *d_in, int64_t n, float nodata, float *h_minmax, void *stream);
int phb_refine_device(phb_ctx *ctx, const float *d_in, float nodata, const float *d_land, float land_nodata,
                      const float *d_shallow, float shallow_nodata, int64_t n, int flags, const float *args,
                      const float *minmax, float *d_out, void *stream);
int phb_refine_host(phb_ctx *ctx, const float *h_in, float nodata, const float *h_land, float land_nodata,
                    const float *h_shallow, float shallow_nodata, int nrows, int ncols, int flags, const float *args,
                    float *h_out);

/*
 * int16 scale/offset packing of a grid, the form every photic grid takes in its NetCDF file:
 * compress_2d (model/nc.c:271-320, called by write_nc nc.c:14) and decompress_2d (nc.c:247-266, called by read_nc
 * nc.c:138). Packing on the device before the copy to the host halves the bytes that cross PCIe; NetCDF I/O itself stays
 * on the host. Bit-identical to the reference including its order-dependent range (the running maximum starts at
 * FLT_MIN and is only tested when a value did not lower the running minimum, nc.c:286-298) and the x86 `(short) rint()`.
 *   pack    add_offset = the minimum, scale_factor = (max - min) / 32767, missing_value = SHRT_MIN, cells == (float)spval
 *           -> missing_value. Synchronises the stream once (two floats come back before the packing pass).
 *   unpack  missing_value -> (float)spval, else (float)packed * scale_factor + add_offset.
 */
int phb_nc_pack_device(phb_ctx *ctx, const float *d_grid, int64_t n, double spval, int16_t *d_packed, float *add_offset,
                       float *scale_factor, int16_t *missing_value, void *stream);
int phb_nc_unpack_device(phb_ctx *ctx, const int16_t *d_packed, int64_t n, float add_offset, float scale_factor,
                         int16_t missing_value, double spval, float *d_grid, void *stream);
int phb_nc_pack_host(phb_ctx *ctx, const float *h_grid, int nrows, int ncols, double spval, int16_t *h_packed,
                     float *add_offset, float *scale_factor, int16_t *missing_value);
int phb_nc_unpack_host(phb_ctx *ctx, const int16_t *h_packed, int nrows, int ncols, float add_offset, float scale_factor,
                       int16_t missing_value, double spval, float *h_grid);

/*
 * MODEL Lee_Kd_LS8 / MODEL Lee_Secchi_LS8 (bam.c:3250-3610 -> Kd_LS8 secchi.c:13, secchi_disk_depth secchi.c:59):
 * Lee et al. (2016) diffuse attenuation (mode 0: min over Kd(443,481,530,554,656)) and Secchi-disk depth (mode 1)
 * from the coastal / blue / green / red Landsat-8 reflectance planes; these feed the scene-level priors upstream
 * of samodel(). spv: nodata of the four planes; a cell with any nodata gets spv[0].
 * theta_s: solar zenith in degrees as the reference passes it. Point-wise, bit-identical to the CPU.
 */
int phb_lee_ls8_device(phb_ctx *ctx, int mode, const float *d_coastal, const float *d_blue, const float *d_green,
                       const float *d_red, const float *spv, float theta_s, int64_t n, float *d_out, void *stream);
int phb_lee_ls8_host(phb_ctx *ctx, int mode, const float *h_coastal, const float *h_blue, const float *h_green,
                     const float *h_red, const float *spv, float theta_s, int nrows, int ncols, float *h_out);

/*
 * COMPUTE K (bam.c:2362-2392 -> model/jerlov.c): Jerlov water type and spectral attenuation coefficients, the
 * scene-level prior the K penalties of samodel_error are tuned around. HOST functions by design:
 * a few hundred transect points in, a handful of scalars out, sequential FP64 regression sums (jerlov_host.h);
 * no device and no phb_ctx needed. Bit-identical to the reference (tests/test_jerlov.py, golden from oracle/_ref).
 *   phb_jerlov_fit          `jerlov` jerlov.c:75-210 (+ `linear_fit` common.c:418): out6 = K_i, K_j, slope m,
 *                           intercept c, correlation r, fractional water-type index (0 = OI .. 9 = C9);
 *                           PHB_ENOFIT where the reference returns false (outputs computed so far are kept,
 *                           the others are 0). The reference's progress printf()s are not reproduced.
 *   phb_jerlov_k            `compute_k` / `compute_k_from_jerlov` jerlov.c:274-316: K at each wavelength for a
 *                           water-type index in [0, 9) (the reference reads past its table beyond that: PHB_EINVAL);
 *                           0.0 for a wavelength outside 400..700 nm, as there.
 *   phb_jerlov_k_from_ratio `compute_k_from_ratio` jerlov.c:214-270.
 */
int phb_jerlov_fit(float wlen_i, float wlen_j, float lsm_i, float lsm_j, const float *Li, const float *Lj, int npoints,
                   float manual_ratio, float *out6, int32_t *n_shallow);
int phb_jerlov_k(float water_type, const float *wavelengths, int n, float *k);
int phb_jerlov_k_from_ratio(float ratio, float wlen_i, float wlen_j, const float *wavelengths, int n, float *water_type,
                            float *k);

/* FP64 pipe peak of this device, measured with a dependent-free DFMA chain kernel (MEASURED_PEAKS.json
 * has no FP64 entry). Returns TFLOP/s (2 flops per DFMA) and the kernel time. */
int phb_fp64_peak(phb_ctx *ctx, double *tflops, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* PHOTIC_B200_H_ */
