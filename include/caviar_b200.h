/*
 * caviar_b200.h -- C-ABI of the B200-native CAVIAR pair-feature hot path.
 *
 * This is the drop-in boundary.  Every entry point takes
 * plain pointers and sizes; no torch / numpy / C++ types cross it.  All
 * functions return 0 (CV_OK) or a negative CV_E* code and never throw; the
 * message of the last failure on the calling thread is caviar_last_error().
 *
 * What each entry point replaces in the reference (/root/reference):
 *
 *   caviar_pair_distances_host   CAVIAR-1.0.py:110-122  compute_distances_parallel()
 *                                (called at :337, :338, :671) -- host (T,n,3) f32
 *                                in, host (T,P) f32 out, no PBC, bit-exact.
 *   caviar_features_host         CAVIAR_Distances-Extractor.py:143,159 -- the
 *                                `distance dNNNN selA selB` series the cpptraj
 *                                subprocess evaluates per frame (imaging on),
 *                                plus the north_star extensions (angles, flags,
 *                                occupancy, scores).
 *   caviar_run (aggregates)      CAVIAR-1.0.py:342,380,496,674,783 (per-pair
 *                                mean over frames = sum/T), :539-545 (mean/var
 *                                from sum, sum of squares), :186 (threshold).
 *   caviar_plan_atom_order /     CAVIAR-1.0.py:335-336,670  atom_slice(ca).xyz -- the compaction of the selected atoms.
 *   caviar_stage_frames          The plan publishes the ORDER in which it wants them (per-tile blocks, an atom repeated
 *                                once per tile that references it); whoever slices the trajectory hands the frames over
 *                                in that order (raw float32 xyz, nothing else is done to them) and the fused kernel
 *                                streams them as they are.  caviar_stage_frames is the same permutation on the device
 *                                for callers whose frames are already compact in their own order.
 *   caviar_compact_frames_host   the same selection at the trajectory-file boundary
 *                                (cpptraj's masks, Extractor:133-138), on the host.
 *   caviar_residue_components    CAVIAR-1.0.py:183-199 (mean-distance threshold :186 -> residue graph -> components).
 *   caviar_cov_split / _gemm     CAVIAR-1.0.py:396-402 (_cov_blocks: X0.T @ X0 / n ... of the VAMP-2 scan) on tcgen05.
 *   caviar_weighted_scores       north_star extension: per-frame scores weighted by the
 *                                abs_loading CAVIAR-1.0.py:141-147 publishes per pair.
 *   caviar_format_csv            CAVIAR_Distances-Extractor.py:196-198 (CSV rows).
 *
 * Ownership: the caller owns every buffer it passes.  The library keeps no
 * pointer after a call returns, except inside the opaque CvPlan (device copies
 * of the index tables and cut-offs), which lives until caviar_plan_destroy(),
 * and the plan caviar_pair_distances_host keeps for its next call
 * (caviar_release_cache() frees it).
 * A plan may be used from several streams at once as long as each stream has
 * its own workspace.
 */
#ifndef CAVIAR_B200_H
#define CAVIAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CV_ABI_VERSION 3

/* status codes */
#define CV_OK              0
#define CV_E_INVALID      -1   /* bad argument (null pointer, negative size, index out of range) */
#define CV_E_CUDA         -2   /* a CUDA runtime call failed; see caviar_last_error() */
#define CV_E_NO_DEVICE    -3   /* no CUDA device / wrong architecture: there is no CPU fallback */
#define CV_E_UNSUPPORTED  -4   /* e.g. a cell too skewed for the image search */
#define CV_E_NOMEM        -5

/* periodic box kinds */
#define CV_BOX_NONE       0    /* plain Euclidean distance (the reference's compute_distances_parallel) */
#define CV_BOX_ORTHO      1    /* per frame 3 floats: edge lengths */
#define CV_BOX_TRICLINIC  2    /* per frame 9 floats: rows = lattice vectors, lower triangular
                                  (a along x, b in the xy plane -- the Amber/cpptraj setting) */

/* plan flags */
#define CV_PLAN_FORCE_DIRECT  1  /* never stage: gather straight from the caller's frames */
#define CV_PLAN_FORCE_STAGED  2  /* always stage through the compaction kernel */

/* kernel paths a plan can resolve to */
#define CV_MODE_DIRECT    0    /* gather from global memory, any frame stride */
#define CV_MODE_ZEROCOPY  1    /* TMA-stream the caller's frames as they are (small, aligned frames) */
#define CV_MODE_STAGED    2    /* frames in plan order (per-tile blocks of raw float32 xyz), TMA-streamed */
#define CV_MODE_ALLPAIRS  3    /* the reference's own list -- all (i, j >= i + minsep), i-major (CAVIAR-1.0.py:331) -- without
                                  a box: 2-D tiles of the pair matrix, each streaming only its rows' and columns' atoms
                                  of the caller's frames: 16-byte aligned, a multiple of 4 floats apart and at least
                                  3 * ceil4(n_atoms) floats long -- caviar_features_host pads on upload when they are not */

/* atom order of the frames handed to caviar_features_host */
#define CV_ORDER_CALLER   0    /* (T, n_atoms, 3) as the pair / triplet lists index it */
#define CV_ORDER_PLAN     1    /* (T, staged_floats_per_frame / 3, 3): atom caviar_plan_atom_order()[k] in slot k
                                  (plans whose mode is not CV_MODE_STAGED have no plan order: treated as CALLER) */

typedef struct CvPlan CvPlan;

typedef struct CvPlanDesc {
    uint32_t struct_size;        /* sizeof(CvPlanDesc), for ABI evolution */
    int32_t  n_atoms;            /* atoms per input frame (N) */
    int64_t  n_pairs;            /* P */
    const int32_t* pairs;        /* host, [P][2] atom indices in [0, N) */
    int64_t  n_triplets;         /* A */
    const int32_t* triplets;     /* host, [A][3] (a, b, c): angle at b, end distance a..c */
    int32_t  box_kind;           /* CV_BOX_* */
    int32_t  flags;              /* CV_PLAN_* */
    double   pair_cutoff;        /* contact flag:  (double)d < pair_cutoff   (CAVIAR-1.0.py:186 semantics) */
    double   hbond_dist_cutoff;  /* triplet flag:  (double)d(a,c) < hbond_dist_cutoff  and            */
    double   hbond_angle_cutoff; /*                (double)angle_deg > hbond_angle_cutoff             */
    int32_t  device;             /* CUDA device ordinal, -1 = current */
    int32_t  n_ctas;             /* 0 = auto (SM count x resident CTAs) -- for tests */
} CvPlanDesc;

typedef struct CvPlanInfo {
    int32_t  mode;                    /* CV_MODE_* */
    int32_t  n_tiles;                 /* column tiles (feature slices handled by one CTA at a time) */
    int32_t  n_ctas;                  /* persistent grid size */
    int32_t  threads_per_cta;
    int32_t  stages;                  /* TMA pipeline depth (0 in direct mode) */
    int32_t  frames_per_stage;
    int32_t  smem_bytes;              /* dynamic shared memory per CTA */
    int32_t  partial_slots;           /* max frame chunks per run = rows of the partial-aggregate buffer */
    int64_t  n_features;              /* P + A */
    int64_t  n_flag_words;            /* ceil(P/32) + ceil(A/32) uint32 words per frame */
    int64_t  n_distinct_atoms;        /* U */
    int64_t  staged_floats_per_frame; /* floats per frame of the plan-order layout (0 unless CV_MODE_STAGED) */
    int64_t  boxrec_floats_per_frame; /* floats per frame of a prepared box record */
    int64_t  workspace_bytes;         /* caller-provided scratch for caviar_run */
    int64_t  algorithmic_bytes_per_frame; /* 12U + box + 4P + 4A + 4*flag_words + 4 */
    int64_t  staged_fixedpoint_floats; /* always 0 since ABI 3: plan-order frames are raw float32 xyz throughout */
    int64_t  gather_bank_conflicts;   /* staged mode: gather references the block layout could not keep off a used
                                         shared-memory bank (0 = every warp gather is conflict-free) */
} CvPlanInfo;

/* Per-frame outputs of one caviar_run call (device pointers, any may be NULL). */
typedef struct CvFrameOutputs {
    float*    dist;    /* [n_frames][P]  row-major, column k <-> pairs[k] (CAVIAR-1.0.py:122) */
    float*    angle;   /* [n_frames][A]  degrees */
    uint32_t* flags;   /* [n_frames][n_flag_words]; pair words first, then triplet words;
                          bit b of word w = feature 32*w + b of that kind */
    int32_t*  score;   /* [n_frames] number of set flags in the frame */
} CvFrameOutputs;

/* Running per-feature aggregates (device pointers): occupancy alone, or occupancy + sum + sumsq.  caviar_run ADDS
 * the block's totals, so a trajectory processed in several blocks accumulates.
 * A run that asks for no dist / angle matrix and no moments -- a contact map: flags, scores, occupancy -- takes the
 * lean kernel flavour (no moment accumulation; without a box the cut-off is applied to the squared distance,
 * bit-identical flags, no square root). */
typedef struct CvAggregates {
    uint64_t* occupancy;  /* [P + A] frames in which the flag was set */
    double*   sum;        /* [P + A] sum over frames of d (pairs) / angle (triplets) */
    double*   sumsq;      /* [P + A] sum of squares */
    double*   occupancy_f64; /* alternative to `occupancy` (set exactly one of the two): the same counts accumulated as
                                float64 -- exact below 2^53 -- so that occupancy, sum and sumsq can live in ONE
                                contiguous float64 buffer and the multi-GPU exchange is a single sum all-reduce of it */
} CvAggregates;

/* Host-side planning only (no CUDA call): what plan_create would decide. */
int caviar_plan_describe(const CvPlanDesc* desc, int32_t n_sms, CvPlanInfo* info);

/* Host-side planning only (no CUDA call): how caviar_run would cut a run of n_frames frames into the frame chunks of its
 * work queue (generic kernels; the all-pairs flavour chunks on its own).  out[4] = {chunk_frames, tail_frames, n_big_chunks,
 * n_chunks}: chunks [0, n_big_chunks) hold chunk_frames frames each, the rest tail_frames each (the last one clipped at
 * n_frames); n_chunks never exceeds CvPlanInfo.partial_slots.  For CPU-side verification that every frame is covered once. */
int caviar_plan_chunks(const CvPlanDesc* desc, int32_t n_sms, int64_t n_frames, int32_t* out);

/* Host-side planning only: the layout tables a plan would upload, for CPU-side verification of the planner.
 *   tile_desc [n_tiles][8] int32 = {kind, f_begin, n_feat, idx_off, grp_off, grp_floats, 0, 0}
 *   idx       [3][n_tiles*896] int32: float offset (3*slot) of atom role 0/1/2 inside the tile's block
 *   src_off   [staged_floats_per_frame] int32: staged float j <- float src_off[j] of the input frame
 * Any buffer may be NULL. */
int caviar_plan_layout(const CvPlanDesc* desc, int32_t n_sms, int32_t* tile_desc, int32_t* idx, int32_t* src_off);

/* Host-side check of the all-pairs geometry (CV_MODE_ALLPAIRS), no CUDA call: runs the kernel's own per-thread set-up for
 * every tile, warp and lane and verifies that every pair is evaluated exactly once from the right atoms.  Returns the
 * number of tiles (0: the plan does not use that flavour), < 0 on a violation.  stats (optional, 8 values): windows, fast
 * windows, largest tile block in atoms, frames per stage, stages, shared-memory bytes, mean atoms per tile, CTAs. */
int64_t caviar_plan_allpairs_check(const CvPlanDesc* desc, int32_t n_sms, int64_t* stats);

int  caviar_plan_create(const CvPlanDesc* desc, CvPlan** plan);
void caviar_plan_destroy(CvPlan* plan);
int  caviar_plan_info(const CvPlan* plan, CvPlanInfo* info);

/* Raw boxes (device, float32: [T][3] ortho or [T][9] triclinic) -> per-frame box
 * records (device, boxrec_floats_per_frame each): inverse edges, safe radius and
 * the list of lattice translations the image search has to try. */
int caviar_prepare_box(const CvPlan* plan, const float* box_dev, int64_t n_frames,
                       float* boxrec_dev, void* stream);

/* The plan order (host-side planning only, no CUDA call): slot k of a plan-order frame holds atom order[k] of the
 * caller's numbering, k in [0, staged_floats_per_frame / 3).  Blocks of consecutive slots belong to one column tile;
 * an atom referenced by several tiles appears once per tile, block tails are padded with a repeated atom.  Slicing a
 * trajectory with this list -- frames[:, order, :] -- IS the hand-over format of CV_MODE_STAGED plans (the reference
 * slices with its own list at CAVIAR-1.0.py:335).  Returns the number of slots (0 for plans that are not staged), or a
 * negative CV_E_* code; order may be NULL to query the size. */
int64_t caviar_plan_atom_order(const CvPlanDesc* desc, int32_t n_sms, int32_t* order, int64_t cap);

/* The same permutation on the device (CV_MODE_STAGED only): frames (device, frame t at xyz + t*stride floats, AoS xyz,
 * caller's atom order) -> plan-order layout (device, staged_floats_per_frame each, 16-byte aligned).  Pure data
 * movement, coordinates are copied bit for bit whatever the box. */
int caviar_stage_frames(const CvPlan* plan, const float* xyz_dev, int64_t frame_stride_floats,
                        int64_t n_frames, float* staged_dev, void* stream);

/* The fused hot path over n_frames frames resident on the device.
 *   coords: CV_MODE_STAGED -> frames in plan order, raw float32 xyz (frame_stride_floats apart, 0 = densely packed,
 *           otherwise >= staged_floats_per_frame and a multiple of 4); otherwise the caller's frames themselves
 *           (frame_stride_floats apart; CV_MODE_ZEROCOPY needs stride == 3*n_atoms; CV_MODE_ALLPAIRS takes frames that
 *           are 16-byte aligned, a multiple of 4 floats apart and >= 3*ceil4(n_atoms) floats long, adds the aggregates in
 *           place -- no finalize launch -- and otherwise falls back to the generic kernel of the same list where that
 *           kernel takes the same frames, or fails with CV_E_INVALID asking for the padding).
 *           With a box, atoms may sit anywhere (unwrapped trajectories included): displacements are minimum-imaged.
 *   boxrec: prepared records, NULL iff box_kind == CV_BOX_NONE.
 * Launches on `stream` (a cudaStream_t) and returns without synchronising. */
int caviar_run(const CvPlan* plan, const float* coords_dev, int64_t frame_stride_floats,
               const float* boxrec_dev, int64_t n_frames,
               const CvFrameOutputs* out, const CvAggregates* agg,
               void* workspace_dev, void* stream);

/* Number of kernels the last caviar_run / *_host call on this thread launched. */
int64_t caviar_last_launch_count(void);

/* Weighted frame scores from the packed flags of a run:  score[t] = sum over the set flags of frame t of
 * weights_q32[f], f counting pairs first, then triplets (the weights CAVIAR publishes per pair
 * are the abs_loading column of save_tic_csv, CAVIAR-1.0.py:141-147).  Weights and scores are 32.32 fixed point
 * (int64 = value * 2^32, rounded by the caller), so the result is exact and independent of the summation order.
 * flags_dev: [n_frames][n_flag_words] as written by caviar_run; all pointers are device pointers. */
int caviar_weighted_scores(const CvPlan* plan, const uint32_t* flags_dev, int64_t n_frames,
                           const int64_t* weights_q32_dev, int64_t* score_q32_dev, void* stream);

/* Residue clusters from the aggregates, on the device (CAVIAR-1.0.py:183-199 build_residue_components, fed at :496-499
 * and :783-786 with the per-pair time-mean distances and nm_cut = cluster_cut / 10): residues = the plan's atoms, an
 * edge (i, j) for every pair whose float32 time mean is below the cut-off, compared in double and strictly (:186):
 *     mean_p = fl32(sum[p] / n_frames)                                            single run (:783)
 *     mean_p = 0.5f * (fl32(sum[p] / n_frames) + fl32(sum_b[p] / n_frames_b))     pooled run (:496), sum_b_dev != NULL
 * labels_dev [n_atoms] int32 receives the component id of every residue, ids numbered in the order of each
 * component's lowest residue index -- the order in which the reference's DFS discovers them, so the label vector is
 * the reference's.  sum_dev / sum_b_dev: the fp64 `sum` aggregates of caviar_run (first P entries are used). */
int caviar_residue_components(const CvPlan* plan, const double* sum_dev, int64_t n_frames,
                              const double* sum_b_dev, int64_t n_frames_b, double cutoff,
                              int32_t* labels_dev, void* stream);

/* ---- lagged covariance blocks on the tensor cores (CAVIAR-1.0.py:396-402 _cov_blocks) ----
 * C = A^T B / n for (n, K) float32 matrices of centred per-frame features, as three TF32 tensor-core products per
 * block (hi.hi + hi.lo + lo.hi, float32 accumulation in TMEM): float32-class accuracy, not TF32's.
 *
 *   caviar_cov_padded_columns(K)   Kp = K rounded up to the 128-column tile: row pitch of the hi / lo planes.
 *   caviar_cov_split               x (n rows, ld_x floats apart, K columns), optional mu[K] (the per-column mean to
 *                                  subtract: dist - mu_pool, CAVIAR-1.0.py:381-382) -> hi, lo planes (n, Kp) float32,
 *                                  128-byte aligned, hi + lo == x - mu exactly, both exactly representable in TF32 up
 *                                  to the truncation of lo.  One pass, HBM-bound.
 *   caviar_cov_gemm                c[i][j] (+)= scale * sum_t a[t][i] * b[t][j], i, j in [0, K): one 128 x 128 tile per
 *                                  CTA, TMA -> shared memory -> tcgen05.mma.kind::tf32 -> TMEM -> global.
 *                                  Row t of a / b is at plane + t * row_stride_floats (>= Kp, multiple of 4), so
 *                                  a_* / b_* may be views of the SAME planes: X0 = X[:-tau], Xt = X[tau:] (pass
 *                                  plane + tau * Kp), the even / odd rows of the reference's cross-validation
 *                                  (stride 2 * Kp, :409-421).  accumulate != 0 adds to c (stacked ensembles).
 *                                  scale is usually 1 / max(1, n_total) (:398-401).
 * All pointers are device pointers; asynchronous on `stream`. */
int64_t caviar_cov_padded_columns(int64_t n_cols);
int caviar_cov_split(const float* x_dev, int64_t ld_x, const float* mu_dev, int64_t n_rows, int64_t n_cols,
                     float* hi_dev, float* lo_dev, void* stream);
int caviar_cov_gemm(const float* a_hi_dev, const float* a_lo_dev, const float* b_hi_dev, const float* b_lo_dev,
                    int64_t n_rows, int64_t row_stride_floats, int64_t n_cols, float scale, int32_t accumulate,
                    float* c_dev, int64_t ld_c, void* stream);

/* Reference-facing, HOST buffers in and out (pageable or pinned).  Streams the
 * frames through pinned staging, double buffered.  Any output may be NULL.
 *   box_host: NULL | [T][3] | [T][9] float32 according to the plan's box_kind.
 *   aggregates are written (not accumulated) as host arrays of P + A entries. */
typedef struct CvHostOutputs {
    float*    dist;       /* [T][P] */
    float*    angle;      /* [T][A] */
    uint32_t* flags;      /* [T][n_flag_words] */
    int32_t*  score;      /* [T] */
    uint64_t* occupancy;  /* [P + A] */
    double*   sum;        /* [P + A] */
    double*   sumsq;      /* [P + A] */
    const int64_t* weights_q32;  /* in, optional: [P + A] feature weights, 32.32 fixed point (see caviar_weighted_scores) */
    int64_t*  score_q32;  /* [T] weighted frame scores, 32.32 fixed point; needs weights_q32 */
} CvHostOutputs;

int caviar_features_host(const CvPlan* plan, const float* xyz_host, int64_t frame_stride_floats,
                         const float* box_host, int64_t n_frames, const CvHostOutputs* out,
                         int64_t frames_per_block /* 0 = auto */, int32_t input_order /* CV_ORDER_* */);

/* compute_distances_parallel(X, pairs) in one call (no box, no cut-offs): fills dist_host [T][P].
 * CAVIAR-1.0.py:110-122.  The plan and the device / bounce buffers of the LAST call are kept and reused when the
 * next call brings the same n_atoms and pair list on the same device (the reference calls the operator once per
 * trajectory with one list, :337-338); caviar_release_cache() frees them.  Thread-safe (calls are serialised). */
int caviar_pair_distances_host(const float* xyz_host, int64_t n_frames, int32_t n_atoms,
                               const int32_t* pairs, int64_t n_pairs, float* dist_host);
void caviar_release_cache(void);

/* Host-side compaction at the trajectory-file boundary (the atom_slice of CAVIAR-1.0.py:335; what cpptraj's masks
 * select at Extractor:133-138):  out[t][k] = atom atoms[k] of frame t, native float32 xyz.  `src` is the file's
 * coordinate block as mapped into memory: frames frame_stride_bytes apart, 12 bytes per atom, big-endian when
 * big_endian != 0 (Amber NetCDF is XDR).  Multi-threaded; touches only the selected atoms. */
int caviar_compact_frames_host(const void* src, int64_t n_frames, int64_t frame_stride_bytes, int32_t n_atoms,
                               const int32_t* atoms, int64_t n_sel, int32_t big_endian, float* out);

/* The same for formats that keep a frame as three separate coordinate arrays (DCD: X[N], Y[N], Z[N], each inside a
 * Fortran record): x_off / y_off / z_off = byte offsets of those arrays inside a frame; byte_swap != 0 when the file's
 * byte order is not the host's. */
int caviar_compact_frames_soa_host(const void* src, int64_t n_frames, int64_t frame_stride_bytes, int32_t n_atoms,
                                   int64_t x_off, int64_t y_off, int64_t z_off, const int32_t* atoms,
                                   int64_t n_sel, int32_t byte_swap, float* out);

/* GROMACS XTC trajectories (the reference CLI advertises them, CAVIAR_Distances-Extractor.py:99, and leaves them to
 * cpptraj): compressed integer coordinates, decoded on the host threads (cv_xtc.cu restates the published xdr3dfcoord
 * algorithm of the xdrfile library).  `data` / `size`: the file image (e.g. a read-only memory map).
 *   caviar_xtc_index: number of frames (< 0: CV_E_*), atom count in *n_atoms, byte offset of every frame in offsets[]
 *                     when it has room for them (cap entries; NULL to count only).
 *   caviar_xtc_decode_host: frames at offsets[0 .. n_frames) -> xyz_out [n_frames][n_sel][3] float32 = positions of the
 *                     selected atoms (atoms == NULL: all) * scale, box_out [n_frames][9] = box vectors * scale or NULL
 *                     (XTC is in nm: scale = 10 gives Angstrom). */
int64_t caviar_xtc_index(const void* data, int64_t size, int64_t* offsets, int64_t cap, int32_t* n_atoms);
int caviar_xtc_decode_host(const void* data, int64_t size, const int64_t* offsets, int64_t n_frames,
                           const int32_t* atoms, int64_t n_sel, float scale, float* xyz_out, float* box_out);

/* Amber ASCII trajectories (mdcrd; the reference CLI advertises them, CAVIAR_Distances-Extractor.py:99, and leaves them
 * to cpptraj): 3N numbers per frame in Fortran 10F8.3 -- fixed positions, so only the selected atoms are parsed, frames in
 * parallel on the host threads (cv_mdcrd.cu).  `data` / `size`: the file image; `first`: byte of the first frame (after the
 * title line); `frame_bytes`: distance between frames (coordinates + optional box line); `eol`: 1 = LF, 2 = CR LF;
 * frames[0 .. n_frames): frame indices -> xyz_out [n_frames][n_sel][3] float32 (atoms == NULL: all n_atoms), in the
 * file's unit (Angstrom).  A field that is not a number is CV_E_INVALID naming the frame. */
int caviar_mdcrd_decode_host(const void* data, int64_t size, int64_t first, int64_t frame_bytes, int32_t eol,
                             int32_t n_atoms, const int64_t* frames, int64_t n_frames, const int32_t* atoms,
                             int64_t n_sel, float* xyz_out);

/* Host-side CSV row formatter for the Extractor output (CAVIAR_Distances-Extractor.py:147-151,196-198: one frame
 * column, then one distance column per pair).  Formats rows x cols float32 values as
 *     "<first_index + r*index_step>,<v>,<v>,...<EOL>"    with `decimals` fixed decimals (printf "%.Nf" rounding)
 * into `out` (capacity `cap` bytes) using the host cores; <EOL> is "\r\n" when crlf != 0 (what Python's csv.writer,
 * through which the reference writes its rows, emits) and "\n" otherwise.  Returns the number of bytes written, or a
 * negative CV_E_* code (CV_E_NOMEM when `cap` is too small: 26 + 65*cols bytes per row suffice for any float; rows of
 * ordinary magnitudes take 3 + digits + decimals bytes per value). */
int64_t caviar_format_csv(const float* values, int64_t rows, int64_t cols, int64_t row_stride_floats,
                          int64_t first_index, int64_t index_step, int32_t decimals, int32_t crlf,
                          char* out, int64_t cap);

/* float64 flavour of compute_distances_parallel: the reference returns float64 when X is float64 (numpy type
 * propagation at CAVIAR-1.0.py:118-119).  Same arithmetic order in IEEE double, bit-exact; HBM-bound gather kernel. */
int caviar_pair_distances_host_f64(const double* xyz_host, int64_t n_frames, int32_t n_atoms,
                                   const int32_t* pairs, int64_t n_pairs, double* dist_host);

const char* caviar_last_error(void);
int caviar_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CAVIAR_B200_H */
