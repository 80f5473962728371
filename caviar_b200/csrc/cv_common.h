// Shared host/device definitions of the CAVIAR pair-feature path (sm_100a).
#pragma once
#include <stdint.h>

namespace cv {

constexpr int kConsumerThreads = 224;   // 7 warps evaluate features (+1 producer warp = 8 warps: 2 per SM sub-partition)
constexpr int kProducerThreads = 32;    // 1 warp drives the TMA pipeline (staged / zero-copy modes)
constexpr int kFeatPerThread   = 4;     // features whose atom indices a thread keeps in registers
constexpr int kTileWidth       = kConsumerThreads * kFeatPerThread;   // 896 features per column tile
// Feature -> thread map inside a tile of n_feat columns: with ks = ceil(n_feat / 224) slots per thread, warp w owns the
// 32 * ks CONSECUTIVE columns [32 ks w, 32 ks (w + 1)) and lane l evaluates columns 32 ks w + 32 k + l, k < ks.  Every warp
// instruction still covers 32 consecutive columns (coalesced stores, one ballot = one flag word), partial tiles keep all
// seven warps busy, and a thread's features are neighbours in the list: in an i-major all-pairs list
// (CAVIAR-1.0.py:331) they usually share atom i, which is then loaded once.
constexpr int kFeatStride      = 32;
constexpr int kMaxShifts       = 16;    // lattice translations the image search may have to try
constexpr int kRecFloats       = 12 + 4 * kMaxShifts + kMaxShifts + 12;  // 104 floats = 416 B per frame
constexpr int kMaxStages       = 8;

// box record layout (floats)
enum : int {
    REC_AX = 0, REC_BY = 1, REC_CZ = 2, REC_BX = 3, REC_CX = 4, REC_CY = 5,
    REC_IAX = 6, REC_IBY = 7, REC_ICZ = 8, REC_RSAFE2 = 9, REC_NSHIFT = 10,
    REC_FAR = 11,                          // |component| of a raw displacement beyond which it is pre-reduced in fp64
    REC_SHIFTS = 12,                       // kMaxShifts x float4 {vx, vy, vz, |v|^2 / 2}
    REC_COEF = REC_SHIFTS + 4 * kMaxShifts,  // kMaxShifts x int: (sa+2) | (sb+2)<<4 | (sc+2)<<8, v = sa*a + sb*b + sc*c
    REC_DBL = REC_COEF + kMaxShifts        // 6 doubles: ax, by, cz, bx, cx, cy (8-byte aligned: 92 floats in)
};

enum : int { KIND_PAIR = 0, KIND_TRIPLET = 1 };
enum : int { AP_SHARED = 0, AP_FAST = 1, AP_GENERAL = 2 };   // flavours of allpairs_kernel, tiles listed in this order

// One column tile = a contiguous slice of features evaluated by one CTA at a time.
struct TileDesc {
    int32_t kind;        // KIND_PAIR / KIND_TRIPLET
    int32_t f_begin;     // first feature, counted within its kind (multiple of 32)
    int32_t n_feat;      // features in the tile (<= kTileWidth)
    int32_t idx_off;     // first slot of the tile in the index tables (tile * kTileWidth)
    int32_t grp_off;     // float offset of the tile's atom block inside one staged frame
    int32_t grp_floats;  // floats of that block (multiple of 4)
    int32_t reserved[2];
};

struct RunParams {
    const float* coords;        // plan-order frames (staged layout), or the caller's frames
    long long    frame_stride;  // floats between consecutive frames of `coords`
    const float* boxrec;        // kRecFloats per frame, or nullptr
    long long    n_frames;
    const TileDesc* tiles;
    const int32_t*  idx0;       // per tile slot: 3*atom (float offset) of atom role 0 / 1 / 2
    const int32_t*  idx1;
    const int32_t*  idx2;
    int32_t n_tiles;
    int32_t frames_per_stage;
    int32_t n_stages;
    int32_t stage_floats;       // floats reserved per pipeline stage for coordinates
    int32_t chunk_frames;       // frames per work item (multiple of frames_per_stage)
    int32_t tail_frames;        // frames per item of the last chunks (chunk_frames / 4): guided scheduling, small tail
    int32_t n_big_chunks;       // chunks [0, n_big_chunks) have chunk_frames frames, the rest tail_frames
    int32_t n_chunks;           // work items = n_chunks * n_tiles, chunk-major
    unsigned int* queue;        // work counter, zeroed before the launch
    float*    dist;
    float*    angle;
    uint32_t* flags;
    int32_t*  score;
    long long P, A, W, Wp;      // pairs, triplets, flag words per frame, pair flag words
    unsigned long long* occ_part;   // [n_chunks][P+A]: every (chunk, tile) item writes its own row segment
    double* sum_part;
    double* sq_part;
    float pair_thr;             // d < pair_thr
    float pair_thr2;            // no box, lean runs: d*d-before-sqrt < pair_thr2  <=>  sqrt_rn(.) < pair_thr
    float hb_dthr;              // d(a,c) < hb_dthr
    float hb_athr;              // angle > hb_athr
    int32_t whole_frame;        // 1: a stage is frames_per_stage consecutive whole frames (one bulk copy)
};

// ------------------------------------------------------------------------------------------------
// All-pairs path (cv_allpairs.cuh): the i-major list of CAVIAR-1.0.py:331 / :667 without a box
// ------------------------------------------------------------------------------------------------
// The list is cut into WINDOWS of 128 consecutive pairs (= 4 flag words, window b covers pairs [128 b, 128 b + 128)).
// A warp owns TWO windows for a whole work item (8 pairs per lane: lane l evaluates pairs 128 b + 32 m + l, m < 4, of
// both), a tile is 7 warps = 14 windows.  Windows that lie inside one row of the pair matrix ("fast": atom i is
// warp-uniform, atom j is affine in the slot) are grouped 2-D -- 14 consecutive rows of the same 128-column band --
// so that a tile needs ~270 atoms of the frame instead of all of them; windows that contain a row boundary ("general":
// every lane keeps its own (i, j) per slot) are grouped by row.  A tile's atoms are one or two contiguous ranges of
// the frame, streamed by TMA as they are.
constexpr int kApSlots  = 4;                       // flag words per window
constexpr int kApWindow = 32 * kApSlots;           // pairs per window
#ifndef CV_AP_WARPS
#define CV_AP_WARPS 7
#endif
// Consumer warps of an all-pairs CTA (+ 1 producer warp).  Seven like features_kernel.  (Warp w runs on SM sub-partition
// w % 4, so 7 + 1 warps leave one scheduler per CTA with half the work; 8 + 1 warps -- -DCV_AP_WARPS=8 -- balance them but
// measured slower: 288 threads x 108 registers fit only ONE CTA per SM at the 4-warp allocation granularity (full flavour
// 4.3 against 6.1e11 evals/s), and the 80-register contact-map flavour, which keeps two, lost 3 %.)
constexpr int kApWarps  = CV_AP_WARPS;             // window pairs per tile: one per consumer warp
constexpr int kApConsumerThreads = 32 * kApWarps;
constexpr int kApThreads = kApConsumerThreads + kProducerThreads;
constexpr int kApMinGap = 32;                      // atoms: a hole at least this wide splits a tile's atoms into two ranges

struct ApTile {
    int32_t blk[2][kApWarps];   // window index of warp w (first / second window), -1 = none
    int32_t fast_mask;          // bit w: both windows of warp w are whole single-row windows
    int32_t a0, n0;             // atoms [a0, a0 + n0) of the frame (multiples of 4) ...
    int32_t a1, n1;             // ... [a1, a1 + n1) ...
    int32_t a2, n2;             // ... and [a2, a2 + n2); n = 0: range unused (ranges ascending, disjoint)
    int32_t reserved;
};

struct ApParams {
    const float* coords;        // the caller's frames, AoS xyz
    long long    frame_stride;  // floats between frames (multiple of 4)
    long long    n_frames;
    const ApTile*  tiles;
    const int32_t* row_of_word; // [ceil(P/32)]: row (atom i) of pair 32 w
    const int32_t* row_off;     // [n_rows + 1]: first pair of row i; row_off[n_rows] = P
    int32_t n_tiles;
    int32_t minsep;             // j >= i + minsep
    int32_t frames_per_stage;
    int32_t n_stages;
    int32_t stage_floats;       // floats reserved per pipeline stage
    int32_t chunk_frames, tail_frames, n_big_chunks, n_chunks;     // work queue, as in RunParams
    unsigned int* queue;
    float*    dist;
    uint32_t* flags;
    int32_t*  score;
    long long P, W;
    // Aggregates are accumulated IN PLACE, in chunk order: item (chunk c, tile) adds its registers to the totals after item
    // (c - 1, tile) has (seq[tile] == c; the queue is chunk-major, so that item was handed out earlier).  A fixed order,
    // hence bit-reproducible sums, without per-chunk partial rows -- which lets chunks be short (see caviar_run).
    unsigned int* seq;              // [n_tiles] zeroed before the launch
    unsigned long long* occ;        // [P] or nullptr
    double* occ_f64;                // [P] or nullptr (occupancy as float64, see CvAggregates)
    double* sum;                    // [P] or nullptr
    double* sq;
    float pair_thr, pair_thr2;
};

constexpr int kQueueBytes = 256;        // the work counter lives at the head of the caller's workspace
constexpr int kMinChunkFrames = 32;     // ... but never fewer frames per item than this (short blocks of the host path)
constexpr int kMinTailFrames = 8;       // ... except the quarter-size items of the tail (per-item flush = 24 B per feature)
constexpr int kFinalizeSlicesFew = 8, kFinalizeSlicesMany = 32;   // chunk slices summed in parallel, combined in a fixed order
constexpr int kFinalizeManyRows = 128;   // from this many chunk rows on, the 32-slice shape
constexpr int kTinyChunkFrames = 4;     // ... and runs with fewer items than CTAs are cut down to this (latency chain of one CTA, config 1)
constexpr int kItemsPerCta = 32;        // work items per CTA the chunking aims for (tail imbalance <= ~1/16)

}  // namespace cv
