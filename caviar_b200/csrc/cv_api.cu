// C-ABI of the CAVIAR pair-feature path (see include/caviar_b200.h).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC cv_api.cu -o libcaviar_b200.so
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/caviar_b200.h"
#include "cv_device.cuh"
#include "cv_allpairs.cuh"
#include "cv_plan.hpp"

using namespace cv;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local int64_t g_launches = 0;

static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CV_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(CV_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));            \
    } while (0)

// CAVIAR_TRACE=1: wall-clock of the host entry points' phases on stderr (development aid, off by default)
namespace {
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    Trace() : on(std::getenv("CAVIAR_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[caviar_b200] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};
}  // namespace

extern "C" const char* caviar_last_error(void) { return g_err.c_str(); }
// other translation units of the library (cv_cov.cu) report through the same thread-local slot
extern "C" int caviar_set_error(int code, const char* msg) { return fail(code, msg ? msg : ""); }
extern "C" int caviar_abi_version(void) { return CV_ABI_VERSION; }
extern "C" int64_t caviar_last_launch_count(void) { return g_launches; }

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
namespace {
struct DevBuf {                       // grow-only device buffer
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t alloc(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    ~DevBuf() { if (p) cudaFree(p); }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct PinBuf {                       // grow-only pinned host buffer
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t alloc(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    ~PinBuf() { if (p) cudaFreeHost(p); }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct Stream {
    cudaStream_t s = nullptr;
    cudaError_t create() { return s ? cudaSuccess : cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); }
    ~Stream() { if (s) cudaStreamDestroy(s); }
};
struct Event {
    cudaEvent_t e = nullptr;
    cudaError_t create() { return e ? cudaSuccess : cudaEventCreateWithFlags(&e, cudaEventDisableTiming); }
    ~Event() { if (e) cudaEventDestroy(e); }
};
// memcpy between pageable user memory and the pinned bounce buffers, spread over a few host threads: one core
// moves ~10 GB/s, PCIe 5 x16 wants ~55 GB/s.
void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    const size_t kMin = 4u << 20;
    static const size_t kCopyThreads = [] {                // CAVIAR_COPY_THREADS overrides (A/B on the GPU box)
        const char* e = std::getenv("CAVIAR_COPY_THREADS");
        return (size_t)std::max(1, e ? atoi(e) : 16);
    }();
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = std::min<size_t>(std::min<size_t>(hw ? hw : 4, kCopyThreads), bytes / kMin);
    if (nt <= 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> pool;
    const size_t step = (bytes / nt + 4095) & ~(size_t)4095;
    for (size_t k = 1; k < nt; ++k) {
        const size_t lo = k * step, hi = std::min(bytes, lo + step);
        if (lo < hi) pool.emplace_back([=] { memcpy((char*)dst + lo, (const char*)src + lo, hi - lo); });
    }
    memcpy(dst, src, std::min(bytes, step));
    for (auto& t : pool) t.join();
}
bool is_pinned(const void* p) {
    if (!p) return true;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
// Everything caviar_features_host needs between calls: allocated on first use, grown on demand, freed with the plan.
struct HostSet {
    DevBuf xyz, staged, dist, angle, flags, score, wscore;
    PinBuf h_in, h_dist, h_angle, h_flags, h_score, h_wscore;
    Event uploaded, computed, downloaded;
};
struct HostCtx {
    HostSet set[2];
    DevBuf ws, occ, sum, sq, box_all, rec_all, wq;
    Stream s_up, s_run, s_down;
    std::mutex mu;
};
}  // namespace

struct CvPlan {
    HostPlan hp;
    int device = 0;
    TileDesc* d_tiles = nullptr;
    int32_t *d_idx0 = nullptr, *d_idx1 = nullptr, *d_idx2 = nullptr;
    int32_t* d_src_off = nullptr;
    int32_t* d_pairs = nullptr;        // [P][2] in the caller's numbering, uploaded on first use (caviar_residue_components)
    ApTile* d_ap_tiles = nullptr;      // all-pairs flavour (hp.ap.on): tiles, row of every flag word, first pair of every row
    int32_t *d_ap_row_of_word = nullptr, *d_ap_row_off = nullptr;
    int* d_err = nullptr;
    mutable HostCtx host;     // cached resources of the host entry point
};

extern "C" int caviar_plan_describe(const CvPlanDesc* desc, int32_t n_sms, CvPlanInfo* info) {
    if (!desc || !info) return fail(CV_E_INVALID, "null argument");
    HostPlan hp;
    std::string why = build_plan(*desc, n_sms, hp);
    if (!why.empty()) return fail(CV_E_INVALID, why);
    fill_info(hp, *info);
    return CV_OK;
}

// Host-side planning only: the layout tables themselves, for tests that check the planner without a GPU.
//   tile_desc : [n_tiles][8] int32 = {kind, f_begin, n_feat, idx_off, grp_off, grp_floats, 0, 0}
//   idx       : [3][n_tiles * 896] int32, float offset (3 * slot) of atom role 0 / 1 / 2 inside the tile's block
//   src_off   : [staged_floats_per_frame] int32, staged float j <- float src_off[j] of the input frame
// Pass NULL buffers to query the sizes (n_tiles via CvPlanInfo, src_off length = staged_floats_per_frame).
extern "C" int caviar_plan_layout(const CvPlanDesc* desc, int32_t n_sms, int32_t* tile_desc, int32_t* idx, int32_t* src_off) {
    if (!desc) return fail(CV_E_INVALID, "null argument");
    HostPlan hp;
    std::string why = build_plan(*desc, n_sms, hp);
    if (!why.empty()) return fail(CV_E_INVALID, why);
    if (hp.ap.on) return fail(CV_E_INVALID, "all-pairs plans keep no index tables: caviar_plan_allpairs_check verifies their geometry");
    const size_t C = hp.tiles.size();
    if (tile_desc)
        for (size_t c = 0; c < C; ++c) {
            const TileDesc& t = hp.tiles[c];
            const int32_t row[8] = {t.kind, t.f_begin, t.n_feat, t.idx_off, t.grp_off, t.grp_floats, t.reserved[0], 0};
            memcpy(tile_desc + 8 * c, row, sizeof row);
        }
    if (idx) {
        memcpy(idx, hp.idx0.data(), hp.idx0.size() * 4);
        memcpy(idx + hp.idx0.size(), hp.idx1.data(), hp.idx1.size() * 4);
        memcpy(idx + 2 * hp.idx0.size(), hp.idx2.data(), hp.idx2.size() * 4);
    }
    if (src_off && !hp.src_off.empty()) memcpy(src_off, hp.src_off.data(), hp.src_off.size() * 4);
    return CV_OK;
}


// Host-side check of the all-pairs geometry (no GPU): runs the kernel's own per-thread set-up (ap_thread_setup) for
// every tile, warp and lane and verifies that every pair of the list is evaluated exactly once, that the shared-memory
// offsets resolve to the pair's atoms inside the tile's atom ranges, that fast warps really are warp-uniform in i and
// affine in j, and that the TMA alignment rules hold.  Returns the number of tiles (0: the plan does not use the
// all-pairs flavour), < 0 on a violation (caviar_last_error says which).
//   stats (optional, 8 values): windows, fast windows, largest tile block in atoms, frames per stage, stages,
//   shared-memory bytes, mean atoms per tile, CTAs
extern "C" int64_t caviar_plan_allpairs_check(const CvPlanDesc* desc, int32_t n_sms, int64_t* stats) {
    if (!desc) return fail(CV_E_INVALID, "null argument");
    HostPlan hp;
    std::string why = build_plan(*desc, n_sms, hp);
    if (!why.empty()) return fail(CV_E_INVALID, why);
    const ApHost& ap = hp.ap;
    if (stats) for (int q = 0; q < 8; ++q) stats[q] = 0;
    if (!ap.on) return 0;
    const int64_t P = hp.P;
    const int n = (hp.desc.n_atoms + 3) / 4 * 4;          // tiles round their atom ranges to multiples of 4 atoms
    std::vector<uint8_t> seen((size_t)P, 0);
    int64_t max_atoms = 0, sum_atoms = 0;
    for (size_t c = 0; c < ap.tiles.size(); ++c) {
        const ApTile& t = ap.tiles[c];
        const std::string at = " (tile " + std::to_string(c) + ")";
        if (t.a0 % 4 || t.n0 % 4 || t.a1 % 4 || t.n1 % 4 || t.a2 % 4 || t.n2 % 4 || t.n0 < 4 || t.a0 < 0 || t.a0 + t.n0 > n)
            return fail(CV_E_INVALID, "atom range 0 misaligned or out of the frame" + at);
        if (t.n1 > 0 && (t.a1 < t.a0 + t.n0 || t.a1 + t.n1 > n)) return fail(CV_E_INVALID, "atom range 1 overlaps range 0 or leaves the frame" + at);
        if (t.n2 > 0 && (t.n1 <= 0 || t.a2 < t.a1 + t.n1 || t.a2 + t.n2 > n)) return fail(CV_E_INVALID, "atom range 2 overlaps range 1 or leaves the frame" + at);
        const int tile_atoms = t.n0 + t.n1 + t.n2;
        if (3 * tile_atoms * ap.frames_per_stage > ap.stage_floats) return fail(CV_E_INVALID, "stage too small" + at);
        max_atoms = std::max<int64_t>(max_atoms, tile_atoms); sum_atoms += tile_atoms;
        if (getenv("CAVIAR_AP_DUMP") && tile_atoms > atoi(getenv("CAVIAR_AP_DUMP")))
            fprintf(stderr, "tile %zu fast_mask %x atoms [%d,+%d) [%d,+%d) [%d,+%d) blk0 %d..%d blk1 %d..%d\n", c, t.fast_mask, t.a0, t.n0, t.a1, t.n1,
                    t.a2, t.n2, t.blk[0][0], t.blk[0][6], t.blk[1][0], t.blk[1][6]);
        auto atom_of = [&](int off) {
            const int a = off / 3;
            return a < t.n0 ? t.a0 + a : (a < t.n0 + t.n1 ? t.a1 + (a - t.n0) : t.a2 + (a - t.n0 - t.n1));
        };
        const bool shared = (int)c < ap.n_shared_tiles;
        for (int w = 0; w < kApWarps; ++w) {
            const bool fast = (t.fast_mask >> w) & 1;
            // the fast kernel runs tiles [0, n_fast_tiles): there a warp has two whole windows or none at all
            if ((int)c < ap.n_fast_tiles && !(fast ? (t.blk[0][w] >= 0 && t.blk[1][w] >= 0) : (t.blk[0][w] < 0 && t.blk[1][w] < 0)))
                return fail(CV_E_INVALID, "fast tile with a half-filled or general warp" + at);
            if ((int)c >= ap.n_fast_tiles && fast) return fail(CV_E_INVALID, "general tile with a fast warp" + at);
            int i_uniform[2] = {-1, -1};
            for (int lane = 0; lane < 32; ++lane) {
                ApThread th;
                ap_thread_setup(t, w, lane, ap.row_of_word.data(), ap.row_off.data(), ap.minsep, P, th);
                if (fast && th.vmask != 0xFFu) return fail(CV_E_INVALID, "fast warp with a missing slot" + at);
                for (int q = 0; q < 2 * kApSlots; ++q) {
                    if (!((th.vmask >> q) & 1u)) continue;
                    const int win = q / kApSlots, m = q % kApSlots;
                    const int64_t k = (int64_t)th.kbase[win] + 32 * m + lane;
                    if (k < 0 || k >= P) return fail(CV_E_INVALID, "slot beyond the list" + at);
                    if (seen[k]++) return fail(CV_E_INVALID, "pair " + std::to_string(k) + " evaluated twice" + at);
                    int io = th.ioff[q], jo = th.joff[q];
                    if (fast) {                       // what ap_fast_frame / ap_shared_frame read
                        io = th.ioff[kApSlots * win]; jo = th.joff[shared ? 0 : kApSlots * win] + 96 * m;
                        if (i_uniform[win] < 0) i_uniform[win] = io;
                        if (io != i_uniform[win]) return fail(CV_E_INVALID, "fast warp is not uniform in i" + at);
                    }
                    if (io < 0 || jo < 0 || io % 3 || jo % 3 || io / 3 >= tile_atoms || jo / 3 >= tile_atoms)
                        return fail(CV_E_INVALID, "offset outside the tile's block" + at);
                    if (atom_of(io) != hp.pairs[2 * k] || atom_of(jo) != hp.pairs[2 * k + 1])
                        return fail(CV_E_INVALID, "pair " + std::to_string(k) + " resolves to the wrong atoms" + at);
                }
            }
        }
    }
    for (int64_t k = 0; k < P; ++k) if (!seen[k]) return fail(CV_E_INVALID, "pair " + std::to_string(k) + " is never evaluated");
    if (stats) {
        stats[0] = ap.n_windows; stats[1] = ap.n_fast_windows; stats[2] = max_atoms; stats[3] = ap.frames_per_stage;
        stats[4] = ap.n_stages; stats[5] = ap.smem_bytes; stats[6] = sum_atoms / (int64_t)ap.tiles.size(); stats[7] = ap.n_ctas;
    }
    return (int64_t)ap.tiles.size();
}

template <typename T>
static cudaError_t upload(T** dst, const std::vector<T>& src) {
    *dst = nullptr;
    if (src.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)dst, src.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

extern "C" void caviar_plan_destroy(CvPlan* plan) {
    if (!plan) return;
    cudaFree(plan->d_tiles);
    cudaFree(plan->d_idx0); cudaFree(plan->d_idx1); cudaFree(plan->d_idx2);
    cudaFree(plan->d_src_off); cudaFree(plan->d_err); cudaFree(plan->d_pairs);
    cudaFree(plan->d_ap_tiles); cudaFree(plan->d_ap_row_of_word); cudaFree(plan->d_ap_row_off);
    delete plan;
}

extern "C" int caviar_plan_create(const CvPlanDesc* desc, CvPlan** out) {
    if (!desc || !out) return fail(CV_E_INVALID, "null argument");
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(CV_E_NO_DEVICE, "no CUDA device: caviar_b200 has no CPU fallback");
    int dev = desc->device;
    if (dev < 0) CV_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CV_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10)
        return fail(CV_E_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100: this library ships sm_100a code only");
    CV_CUDA(cudaSetDevice(dev));
    CvPlan* p = new (std::nothrow) CvPlan();
    if (!p) return fail(CV_E_NOMEM, "out of host memory");
    p->device = dev;
    std::string why = build_plan(*desc, prop.multiProcessorCount, p->hp);
    if (!why.empty()) { delete p; return fail(CV_E_INVALID, why); }
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = upload(&p->d_tiles, p->hp.tiles);
    if (e == cudaSuccess) e = upload(&p->d_idx0, p->hp.idx0);
    if (e == cudaSuccess) e = upload(&p->d_idx1, p->hp.idx1);
    if (e == cudaSuccess) e = upload(&p->d_idx2, p->hp.idx2);
    if (e == cudaSuccess) e = upload(&p->d_src_off, p->hp.src_off);
    if (e == cudaSuccess && p->hp.ap.on) e = upload(&p->d_ap_tiles, p->hp.ap.tiles);
    if (e == cudaSuccess && p->hp.ap.on) e = upload(&p->d_ap_row_of_word, p->hp.ap.row_of_word);
    if (e == cudaSuccess && p->hp.ap.on) e = upload(&p->d_ap_row_off, p->hp.ap.row_off);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_err, sizeof(int));          // box set-up errors
    if (e == cudaSuccess) e = cudaMemset(p->d_err, 0, sizeof(int));
    if (e != cudaSuccess) {
        std::string msg = std::string("plan upload: ") + cudaGetErrorString(e);
        caviar_plan_destroy(p);
        return fail(CV_E_CUDA, msg);
    }
    *out = p;
    return CV_OK;
}

extern "C" int caviar_plan_info(const CvPlan* plan, CvPlanInfo* info) {
    if (!plan || !info) return fail(CV_E_INVALID, "null argument");
    fill_info(plan->hp, *info);
    return CV_OK;
}

// The plan order: slot k of a plan-order (staged) frame holds atom order[k] of the caller's frame.  Host-side only.
extern "C" int64_t caviar_plan_atom_order(const CvPlanDesc* desc, int32_t n_sms, int32_t* order, int64_t cap) {
    if (!desc) return fail(CV_E_INVALID, "null argument");
    HostPlan hp;
    std::string why = build_plan(*desc, n_sms, hp);
    if (!why.empty()) return fail(CV_E_INVALID, why);
    const int64_t n = hp.staged_floats / 3;
    if (order) {
        if (cap < n) return fail(CV_E_INVALID, "atom order buffer too small");
        for (int64_t k = 0; k < n; ++k) order[k] = hp.src_off[3 * k] / 3;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------
// box records / staging
// ------------------------------------------------------------------------------------------------
extern "C" int caviar_prepare_box(const CvPlan* plan, const float* box_dev, int64_t n_frames, float* boxrec_dev,
                                  void* stream) {
    if (!plan) return fail(CV_E_INVALID, "null plan");
    g_launches = 0;
    if (plan->hp.desc.box_kind == CV_BOX_NONE) return CV_OK;
    if (n_frames <= 0) return CV_OK;
    if (!box_dev || !boxrec_dev) return fail(CV_E_INVALID, "null box pointer");
    cudaStream_t st = (cudaStream_t)stream;
    CV_CUDA(cudaMemsetAsync(plan->d_err, 0, sizeof(int), st));
    const int threads = 128;
    const unsigned blocks = (unsigned)((n_frames + threads - 1) / threads);
    box_prepare_kernel<<<blocks, threads, 0, st>>>(box_dev, n_frames, plan->hp.desc.box_kind, boxrec_dev, plan->d_err);
    CV_CUDA(cudaGetLastError());
    g_launches = 1;
    int err = 0;
    CV_CUDA(cudaMemcpyAsync(&err, plan->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CV_CUDA(cudaStreamSynchronize(st));
    if (err & 1) return fail(CV_E_UNSUPPORTED, "triclinic cell is not lower triangular (a along x, b in the xy plane)");
    if (err & 2) return fail(CV_E_INVALID, "box with a non-positive edge");
    if (err & 4) return fail(CV_E_UNSUPPORTED, "cell too skewed: more than 16 lattice translations would have to be searched");
    return CV_OK;
}

extern "C" int caviar_stage_frames(const CvPlan* plan, const float* xyz_dev, int64_t frame_stride_floats,
                                   int64_t n_frames, float* staged_dev, void* stream) {
    if (!plan) return fail(CV_E_INVALID, "null plan");
    g_launches = 0;
    const HostPlan& hp = plan->hp;
    if (hp.mode != CV_MODE_STAGED) return fail(CV_E_INVALID, "plan does not use the staged layout");
    if (n_frames <= 0) return CV_OK;
    if (!xyz_dev || !staged_dev) return fail(CV_E_INVALID, "null coordinate pointer");
    if (frame_stride_floats < 3ll * hp.desc.n_atoms) return fail(CV_E_INVALID, "frame stride smaller than 3*n_atoms");
    if ((reinterpret_cast<uintptr_t>(staged_dev) & 15) != 0) return fail(CV_E_INVALID, "staged buffer must be 16-byte aligned");
    const int S4 = (int)(hp.staged_floats / 4);
    const int64_t chunk = 65535ll * kStageFramesPerThread;     // grid.y is limited to 65535 blocks
    for (int64_t t0 = 0; t0 < n_frames; t0 += chunk) {
        const int64_t n = std::min<int64_t>(chunk, n_frames - t0);
        const unsigned gy = (unsigned)((n + kStageFramesPerThread - 1) / kStageFramesPerThread);
        stage_frames_kernel<<<dim3((S4 + 255) / 256, gy), 256, 0, (cudaStream_t)stream>>>(
            xyz_dev + t0 * frame_stride_floats, frame_stride_floats, n,
            reinterpret_cast<const int4*>(plan->d_src_off), S4, hp.staged_floats / 4,
            reinterpret_cast<float4*>(staged_dev + t0 * hp.staged_floats));
        CV_CUDA(cudaGetLastError());
        g_launches += 1;
    }
    return CV_OK;
}

// ------------------------------------------------------------------------------------------------
// the fused run
// ------------------------------------------------------------------------------------------------
template <int BOX, bool PIPE, bool LEAN>
static cudaError_t launch_features(const CvPlan* plan, const RunParams& rp, cudaStream_t st) {
    auto kern = features_kernel<BOX, PIPE, LEAN>;
    const int threads = PIPE ? kConsumerThreads + kProducerThreads : kConsumerThreads;
    const int smem = PIPE ? plan->hp.smem_bytes : 0;
    if (PIPE) {
        // opt in to > 48 KB of dynamic shared memory once per kernel flavour and device (grow-only, benign race)
        static int granted[64] = {0};
        const int dev = plan->device & 63;
        if (smem > granted[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            granted[dev] = smem;
        }
    }
    kern<<<plan->hp.n_ctas, threads, smem, st>>>(rp);
    return cudaGetLastError();
}


template <bool LEAN, int KIND>
static cudaError_t launch_allpairs(const CvPlan* plan, const ApParams& ap, cudaStream_t st) {
    if (ap.n_tiles <= 0) return cudaSuccess;
    auto kern = allpairs_kernel<LEAN, KIND>;
    const int occ = (LEAN && KIND != AP_GENERAL) ? kApLeanOccupancy : 2;
    const int smem = ap.n_stages * ap.stage_floats * 4 + (2 * kMaxStages * 8 + kMaxStages * 24);
    static int granted[64] = {0};
    const int dev = plan->device & 63;
    if (smem > granted[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        granted[dev] = smem;
    }
    kern<<<plan->hp.ap.n_ctas / 2 * occ, kApThreads, smem, st>>>(ap);
    return cudaGetLastError();
}

// Frame chunks of the work queue: about kItemsPerCta items per CTA, whole pipeline stages each.  The last
// quarter of the frames is cut into chunks a quarter the size, so that the CTAs run dry within a quarter item
// of each other (the queue is chunk-major: small chunks are handed out last).
struct Chunking { int32_t chunk_frames, tail_frames, n_big_chunks, n_chunks; };
static bool make_chunks(int64_t n_frames, int64_t fb, int max_chunks, int64_t n_tiles, int n_ctas, Chunking& c) {
    const int64_t unit = 4 * fb;
    int64_t cf = std::max<int64_t>((7 * n_frames / 4 + max_chunks - 1) / max_chunks, kMinChunkFrames);
    cf = std::max<int64_t>(unit, (cf + unit - 1) / unit * unit);
    if (cf > (1ll << 30)) return false;
    const int64_t tf = cf / 4;
    int64_t n_big = (3 * n_frames / 4) / cf;
    int64_t n_small = (n_frames - n_big * cf + tf - 1) / tf;
    // Short runs (tail items below kMinChunkFrames frames): every item costs a flush of 24 B per feature and a
    // pipeline restart, so the small tail only pays when a CTA gets so few items that imbalance dominates
    // (config 2: 4 items per CTA, +4 %; config 5 at 4 096 frames: 16 per CTA, -3.5 % -> uniform chunks there).
    const int64_t uniform_items = (n_frames + cf - 1) / cf * n_tiles;
    const bool small_tail_pays = uniform_items < 8ll * n_ctas;
    if (tf < kMinTailFrames || (tf < kMinChunkFrames && !small_tail_pays) || n_big + n_small > max_chunks) {
        n_big = (n_frames + cf - 1) / cf; n_small = 0;
    }
    c.chunk_frames = (int32_t)cf;
    c.tail_frames = (int32_t)tf;
    c.n_big_chunks = (int32_t)n_big;
    c.n_chunks = (int32_t)(n_big + n_small);
    // Very short runs (fewer items than CTAs: config 1 has 42 on 296): the pass is ONE CTA's chain of stage round
    // trips, so the frames are cut finer -- uniform chunks, down to kTinyChunkFrames -- until every CTA has an item.
    int64_t tiny_floor = kTinyChunkFrames;
    if (const char* e = std::getenv("CAVIAR_TINY_CHUNK")) tiny_floor = std::max(1, atoi(e));       // A/B: tools/gpu_short_runs.py
    if (uniform_items < n_ctas) {                       // (counted before the quarter-size tail, which only adds small items)
        const int64_t want = (n_ctas + n_tiles - 1) / n_tiles;                   // chunks that make one item per CTA
        const int64_t tiny = std::max<int64_t>(tiny_floor, (n_frames + want - 1) / want);
        const int64_t n_tiny = (n_frames + tiny - 1) / tiny;
        if (tiny < cf && n_tiny <= max_chunks) {
            c.chunk_frames = (int32_t)tiny; c.tail_frames = (int32_t)tiny;
            c.n_big_chunks = (int32_t)n_tiny; c.n_chunks = (int32_t)n_tiny;
        }
    }
    return true;
}

extern "C" int caviar_plan_chunks(const CvPlanDesc* desc, int32_t n_sms, int64_t n_frames, int32_t* out) {
    if (!desc || !out) return fail(CV_E_INVALID, "null argument");
    if (n_frames <= 0) return fail(CV_E_INVALID, "n_frames must be positive");
    HostPlan hp;
    std::string why = build_plan(*desc, n_sms, hp);
    if (!why.empty()) return fail(CV_E_INVALID, why);
    Chunking ck;
    if (!make_chunks(n_frames, hp.frames_per_stage, hp.max_chunks, (int64_t)hp.tiles.size(), hp.n_ctas, ck))
        return fail(CV_E_INVALID, "too many frames for one run: split the call");
    out[0] = ck.chunk_frames; out[1] = ck.tail_frames; out[2] = ck.n_big_chunks; out[3] = ck.n_chunks;
    return CV_OK;
}

extern "C" int caviar_run(const CvPlan* plan, const float* coords_dev, int64_t frame_stride_floats,
                          const float* boxrec_dev, int64_t n_frames, const CvFrameOutputs* out,
                          const CvAggregates* agg, void* workspace_dev, void* stream) {
    g_launches = 0;
    if (!plan) return fail(CV_E_INVALID, "null plan");
    if (n_frames < 0) return fail(CV_E_INVALID, "negative frame count");
    if (n_frames == 0) return CV_OK;
    const HostPlan& hp = plan->hp;
    {
        int cur = -1;
        CV_CUDA(cudaGetDevice(&cur));
        if (cur != plan->device)
            return fail(CV_E_INVALID, "plan was created on device " + std::to_string(plan->device) +
                                          " but the current device is " + std::to_string(cur));
    }
    if (!coords_dev) return fail(CV_E_INVALID, "null coordinate pointer");
    if (hp.desc.box_kind != CV_BOX_NONE && !boxrec_dev) return fail(CV_E_INVALID, "plan has a box but boxrec is null");
    const bool want_agg = agg && (agg->occupancy || agg->occupancy_f64 || agg->sum || agg->sumsq);
    const bool want_moments = want_agg && (agg->sum || agg->sumsq);
    if (want_agg && !agg->occupancy && !agg->occupancy_f64) return fail(CV_E_INVALID, "aggregates: occupancy alone, or occupancy + sum + sumsq");
    if (want_agg && agg->occupancy && agg->occupancy_f64) return fail(CV_E_INVALID, "aggregates: occupancy as uint64 OR as float64, not both");
    if (want_moments && !(agg->sum && agg->sumsq)) return fail(CV_E_INVALID, "aggregates: occupancy alone, or occupancy + sum + sumsq");
    if (!workspace_dev) return fail(CV_E_INVALID, "null workspace (CvPlanInfo.workspace_bytes of device memory)");
    cudaStream_t st = (cudaStream_t)stream;

    RunParams rp{};
    rp.coords = coords_dev;
    rp.boxrec = boxrec_dev;
    rp.n_frames = n_frames;
    // The all-pairs flavour streams ranges of the caller's frames as they are: it needs them a multiple of 16 bytes apart
    // and -- tiles round their atom ranges to multiples of 4 atoms -- 3 * ceil4(n_atoms) floats long (frames of a selection
    // whose size is not a multiple of 4 must be padded: caviar_features_host does that on the way to the device).
    const long long ap_min_stride = 3ll * ((hp.desc.n_atoms + 3) / 4 * 4);
    const bool use_ap = hp.ap.on && frame_stride_floats % 4 == 0 && (reinterpret_cast<uintptr_t>(coords_dev) & 15) == 0 &&
                        frame_stride_floats >= ap_min_stride;
    if (use_ap) {
        rp.frame_stride = frame_stride_floats;
    } else if (hp.ap.on && hp.mode == CV_MODE_STAGED) {
        // (no generic fall-back with the same hand-over format: the generic plan of this list wants plan-order frames)
        return fail(CV_E_INVALID, "all-pairs plan: frames must be 16-byte aligned, a multiple of 4 floats apart and at least "
                                  "3*ceil4(n_atoms) = " + std::to_string(ap_min_stride) + " floats long (pad the frame stride)");
    } else if (hp.mode == CV_MODE_STAGED) {
        // plan-order frames: densely packed unless the caller says otherwise (0 = dense)
        rp.frame_stride = frame_stride_floats > 0 ? frame_stride_floats : hp.staged_floats;
        if (rp.frame_stride < hp.staged_floats || rp.frame_stride % 4 != 0)
            return fail(CV_E_INVALID, "plan-order frames: stride must be >= staged_floats_per_frame and a multiple of 4 floats");
    } else {
        if (frame_stride_floats < 3ll * hp.desc.n_atoms) return fail(CV_E_INVALID, "frame stride smaller than 3*n_atoms");
        if (hp.mode == CV_MODE_ZEROCOPY && frame_stride_floats != 3ll * hp.desc.n_atoms)
            return fail(CV_E_INVALID, "zero-copy mode needs densely packed frames (stride == 3*n_atoms)");
        rp.frame_stride = frame_stride_floats;
    }
    if (!use_ap && hp.mode != CV_MODE_DIRECT && (reinterpret_cast<uintptr_t>(coords_dev) & 15) != 0)
        return fail(CV_E_INVALID, "coordinates must be 16-byte aligned for TMA");
    // every flavour reads the record's head with 128-bit loads (and TMA copies it whole in the ring flavours)
    if (boxrec_dev && (reinterpret_cast<uintptr_t>(boxrec_dev) & 15) != 0) return fail(CV_E_INVALID, "box records must be 16-byte aligned");
    rp.tiles = plan->d_tiles;
    rp.idx0 = plan->d_idx0; rp.idx1 = plan->d_idx1; rp.idx2 = plan->d_idx2;
    rp.n_tiles = (int)hp.tiles.size();
    rp.frames_per_stage = hp.frames_per_stage; rp.n_stages = hp.n_stages; rp.stage_floats = hp.stage_floats;
    rp.whole_frame = hp.whole_frame;
    if (out) { rp.dist = out->dist; rp.angle = out->angle; rp.flags = out->flags; rp.score = out->score; }
    rp.P = hp.P; rp.A = hp.A; rp.W = hp.W; rp.Wp = hp.Wp;
    const long long F = hp.P + hp.A;
    Chunking ck;
    if (use_ap) {
        // The all-pairs flavour adds its aggregates in place, in chunk order (ApParams::seq): no partial rows, so chunks can
        // be SHORT -- and that is what a run that emits the distance matrix wants.  The queue is chunk-major, so short items
        // keep all CTAs within a few hundred frames of each other and the 2 MB rows of the matrix are completed in the L2
        // within microseconds; with ~1 500-frame items the CTAs drift thousands of rows apart, every 512-byte piece reaches
        // HBM on its own and the write rate collapses (config 4, 50 000 frames: 5.3e11 -> 6.3e11 evals/s from 1 470- to
        // 96-frame items; a fill kernel writes 7.5 TB/s on this box, this pattern 3).  Contact-map runs write 1/32 of that
        // and prefer longer items (fewer pipeline restarts).
        const bool emits = out && out->dist;
        int64_t cf = emits ? 128 : 512;
        if (const char* e = std::getenv("CAVIAR_AP_CHUNK")) cf = std::max(1, atoi(e));
        const int64_t fb = hp.ap.frames_per_stage, nt = (int64_t)hp.ap.tiles.size();
        cf = std::max(fb, (cf + fb - 1) / fb * fb);
        while (cf > 4 * fb && nt * ((n_frames + cf - 1) / cf) < 8ll * hp.ap.n_ctas) cf = std::max(4 * fb, cf / 2 / fb * fb);   // short runs: enough items to balance
        const int64_t nc = (n_frames + cf - 1) / cf;
        if (nc >= (1ll << 31) / std::max<int64_t>(nt, 1)) return fail(CV_E_INVALID, "too many frames for one run: split the call");
        ck.chunk_frames = (int32_t)cf; ck.tail_frames = (int32_t)cf; ck.n_big_chunks = (int32_t)nc; ck.n_chunks = (int32_t)nc;
    } else if (!make_chunks(n_frames, hp.frames_per_stage, hp.max_chunks, (int64_t)hp.tiles.size(), hp.n_ctas, ck)) {
        return fail(CV_E_INVALID, "too many frames for one run: split the call");
    }
    rp.chunk_frames = ck.chunk_frames; rp.tail_frames = ck.tail_frames; rp.n_big_chunks = ck.n_big_chunks; rp.n_chunks = ck.n_chunks;
    rp.queue = reinterpret_cast<unsigned int*>(workspace_dev);
    // Lean run (a contact map: flags / scores / occupancy, no feature matrix, no moments): the kernel flavour that
    // skips the moment accumulation and, without a box, the IEEE square root (flag decided on the squared distance).
    const bool lean = !rp.dist && !rp.angle && !want_moments;
    if (want_agg && !use_ap) {
        rp.occ_part = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace_dev) + kQueueBytes);
        if (want_moments) {
            rp.sum_part = reinterpret_cast<double*>(rp.occ_part + (long long)rp.n_chunks * F);
            rp.sq_part = rp.sum_part + (long long)rp.n_chunks * F;
        }
    }
    rp.pair_thr = hp.pair_thr; rp.pair_thr2 = hp.pair_thr2; rp.hb_dthr = hp.hb_dthr; rp.hb_athr = hp.hb_athr;

    CV_CUDA(cudaMemsetAsync(rp.queue, 0, kQueueBytes, st));       // work counters (the all-pairs flavour keeps three, 64 bytes apart)
    if (rp.score) CV_CUDA(cudaMemsetAsync(rp.score, 0, sizeof(int32_t) * n_frames, st));

    cudaError_t e;
    if (use_ap) {
        ApParams ap{};
        ap.coords = coords_dev; ap.frame_stride = frame_stride_floats; ap.n_frames = n_frames;
        ap.tiles = plan->d_ap_tiles; ap.row_of_word = plan->d_ap_row_of_word; ap.row_off = plan->d_ap_row_off;
        ap.n_tiles = (int)hp.ap.tiles.size(); ap.minsep = hp.ap.minsep;
        ap.frames_per_stage = hp.ap.frames_per_stage; ap.n_stages = hp.ap.n_stages; ap.stage_floats = hp.ap.stage_floats;
        ap.chunk_frames = ck.chunk_frames; ap.tail_frames = ck.tail_frames; ap.n_big_chunks = ck.n_big_chunks; ap.n_chunks = ck.n_chunks;
        ap.queue = rp.queue; ap.dist = rp.dist; ap.flags = rp.flags; ap.score = rp.score;
        ap.P = hp.P; ap.W = hp.W;
        ap.seq = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(workspace_dev) + kQueueBytes);
        CV_CUDA(cudaMemsetAsync(ap.seq, 0, sizeof(unsigned int) * (size_t)ap.n_tiles, st));
        if (want_agg) {
            ap.occ = reinterpret_cast<unsigned long long*>(agg->occupancy); ap.occ_f64 = agg->occupancy_f64;
            ap.sum = agg->sum; ap.sq = agg->sumsq;
        }
        ap.pair_thr = hp.pair_thr; ap.pair_thr2 = hp.pair_thr2;
        // tiles are listed [shared columns | fast | general], one instantiation each; the first two store the distance
        // matrix unconditionally, so a run that wants moments without it takes the general flavour throughout
        const bool direct = lean || rp.dist;
        const int n_shared = direct ? hp.ap.n_shared_tiles : 0, n_fast = direct ? hp.ap.n_fast_tiles : 0;
        ApParams part[3] = {ap, ap, ap};
        const int first[4] = {0, n_shared, n_fast, ap.n_tiles};
        for (int k = 0; k < 3; ++k) {
            part[k].tiles = ap.tiles + first[k]; part[k].n_tiles = first[k + 1] - first[k];
            part[k].queue = ap.queue + 16 * k; part[k].seq = ap.seq + first[k];
        }
        if (lean && kApLeanOccupancy > 2) {
            // three CTAs per SM: three rings must fit in shared memory
            for (int k = 0; k < 2; ++k)
                while (part[k].frames_per_stage > 1 && kApLeanOccupancy * (part[k].n_stages * part[k].stage_floats * 4 + 1024) > 220 * 1024) {
                    part[k].stage_floats = part[k].stage_floats / part[k].frames_per_stage * (part[k].frames_per_stage / 2);
                    part[k].frames_per_stage /= 2;
                }
        }
        e = lean ? launch_allpairs<true, AP_SHARED>(plan, part[0], st) : launch_allpairs<false, AP_SHARED>(plan, part[0], st);
        if (e == cudaSuccess) e = lean ? launch_allpairs<true, AP_FAST>(plan, part[1], st) : launch_allpairs<false, AP_FAST>(plan, part[1], st);
        if (e == cudaSuccess) e = lean ? launch_allpairs<true, AP_GENERAL>(plan, part[2], st) : launch_allpairs<false, AP_GENERAL>(plan, part[2], st);
        for (int k = 0; k < 3; ++k) g_launches += part[k].n_tiles > 0 ? 1 : 0;
        g_launches -= 1;                                   // (one launch is counted below)
        if (e != cudaSuccess) return fail(CV_E_CUDA, std::string("allpairs_kernel launch: ") + cudaGetErrorString(e));
    } else {
    // TMA ring (plan-order blocks or whole small frames) unless the plan gathers from the caller's frames
    const bool pipe = hp.mode != CV_MODE_DIRECT;
#define CV_LAUNCH(B, P_) (lean ? launch_features<B, P_, true>(plan, rp, st) : launch_features<B, P_, false>(plan, rp, st))
    switch (hp.desc.box_kind * 2 + (pipe ? 1 : 0)) {
        case 0: e = CV_LAUNCH(0, false); break;
        case 1: e = CV_LAUNCH(0, true); break;
        case 2: e = CV_LAUNCH(1, false); break;
        case 3: e = CV_LAUNCH(1, true); break;
        case 4: e = CV_LAUNCH(2, false); break;
        case 5: e = CV_LAUNCH(2, true); break;
        default: return fail(CV_E_INVALID, "internal: no kernel flavour for this plan");
    }
#undef CV_LAUNCH
    if (e != cudaSuccess) return fail(CV_E_CUDA, std::string("features_kernel launch: ") + cudaGetErrorString(e));
    }
    g_launches += 1;
    if (want_agg && !use_ap) {
        const unsigned blocks = (unsigned)((F + 31) / 32);
        unsigned long long* occ_out = reinterpret_cast<unsigned long long*>(agg->occupancy);
        const bool many = rp.n_chunks >= kFinalizeManyRows;      // (a function of the run's shape only: reproducible)
#define CV_FINALIZE(M, S)                                                                                          \
        finalize_kernel<M, S><<<blocks, dim3(32, S), 0, st>>>(rp.occ_part, rp.sum_part, rp.sq_part, F, rp.n_chunks, \
                                                              occ_out, agg->occupancy_f64, agg->sum, agg->sumsq)
        if (want_moments) { if (many) CV_FINALIZE(true, kFinalizeSlicesMany); else CV_FINALIZE(true, kFinalizeSlicesFew); }
        else              { if (many) CV_FINALIZE(false, kFinalizeSlicesMany); else CV_FINALIZE(false, kFinalizeSlicesFew); }
#undef CV_FINALIZE
        CV_CUDA(cudaGetLastError());
        g_launches += 1;
    }
    return CV_OK;
}

// Thresholded mean-distance graph -> residue components, on the device (CAVIAR-1.0.py:183-199 with the means of :496 / :783).
extern "C" int caviar_residue_components(const CvPlan* plan, const double* sum_dev, int64_t n_frames,
                                         const double* sum_b_dev, int64_t n_frames_b, double cutoff,
                                         int32_t* labels_dev, void* stream) {
    g_launches = 0;
    if (!plan) return fail(CV_E_INVALID, "null plan");
    const HostPlan& hp = plan->hp;
    if (hp.P <= 0) return fail(CV_E_INVALID, "the plan has no pairs");
    if (!sum_dev || !labels_dev) return fail(CV_E_INVALID, "null argument");
    if (n_frames <= 0 || (sum_b_dev && n_frames_b <= 0)) return fail(CV_E_INVALID, "frame counts must be positive");
    const int n = hp.desc.n_atoms, threads = 1024;
    const size_t smem = (size_t)(n + threads) * sizeof(int);
    if (smem > 200 * 1024) return fail(CV_E_UNSUPPORTED, "more than ~50 000 residues: the label table does not fit in shared memory");
    cudaStream_t st = (cudaStream_t)stream;
    CvPlan* mp = const_cast<CvPlan*>(plan);
    if (!mp->d_pairs) {
        std::lock_guard<std::mutex> lock(mp->host.mu);
        if (!mp->d_pairs) {
            int32_t* d = nullptr;
            CV_CUDA(cudaMalloc((void**)&d, hp.pairs.size() * sizeof(int32_t)));
            cudaError_t e = cudaMemcpy(d, hp.pairs.data(), hp.pairs.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { cudaFree(d); return fail(CV_E_CUDA, std::string("pair upload: ") + cudaGetErrorString(e)); }
            mp->d_pairs = d;
        }
    }
    static int granted[64] = {0};
    const int dev = plan->device & 63;
    if ((int)smem > granted[dev] && smem > 48 * 1024) {
        CV_CUDA(cudaFuncSetAttribute(components_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        granted[dev] = (int)smem;
    }
    components_kernel<<<1, threads, smem, st>>>(sum_dev, (double)n_frames, sum_b_dev, (double)std::max<int64_t>(n_frames_b, 1),
                                                plan->d_pairs, hp.P, n, cutoff, labels_dev);
    CV_CUDA(cudaGetLastError());
    g_launches = 1;
    return CV_OK;
}

extern "C" int caviar_weighted_scores(const CvPlan* plan, const uint32_t* flags_dev, int64_t n_frames,
                                      const int64_t* weights_q32_dev, int64_t* score_q32_dev, void* stream) {
    g_launches = 0;
    if (!plan) return fail(CV_E_INVALID, "null plan");
    if (n_frames < 0) return fail(CV_E_INVALID, "negative frame count");
    if (n_frames == 0) return CV_OK;
    if (!flags_dev || !weights_q32_dev || !score_q32_dev) return fail(CV_E_INVALID, "null argument");
    const HostPlan& hp = plan->hp;
    const long long blocks = (n_frames * 32 + 255) / 256;
    if (blocks > 0x7fffffffll) return fail(CV_E_INVALID, "too many frames for one call");
    weighted_score_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        flags_dev, n_frames, hp.W, hp.Wp, hp.P, reinterpret_cast<const long long*>(weights_q32_dev),
        reinterpret_cast<long long*>(score_q32_dev));
    CV_CUDA(cudaGetLastError());
    g_launches = 1;
    return CV_OK;
}

// ------------------------------------------------------------------------------------------------
// reference-facing host entry points
// ------------------------------------------------------------------------------------------------

// Block loop with two device buffer sets: while block b computes, block b+1 is uploaded and block b-1 is
// downloaded.  Pageable host memory goes through pinned bounce buffers (one memcpy each way); memory the
// caller pinned (cudaHostAlloc / cudaHostRegister / torch pin_memory) is DMA'd in place.
namespace {
// the *_host entry points select the plan's device for their own calls and hand the caller's device back on return
struct DeviceGuard {
    int prev = -1;
    cudaError_t enter(int dev) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) { prev = -1; return e; }
        return prev == dev ? cudaSuccess : cudaSetDevice(dev);
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

extern "C" int caviar_features_host(const CvPlan* plan, const float* xyz_host, int64_t frame_stride_floats,
                                    const float* box_host, int64_t n_frames, const CvHostOutputs* out,
                                    int64_t frames_per_block, int32_t input_order) {
    g_launches = 0;
    if (!plan || !out) return fail(CV_E_INVALID, "null argument");
    if (n_frames < 0) return fail(CV_E_INVALID, "negative frame count");
    const HostPlan& hp = plan->hp;
    if (input_order != CV_ORDER_CALLER && input_order != CV_ORDER_PLAN) return fail(CV_E_INVALID, "unknown input_order");
    // frames already sliced in plan order (caviar_plan_atom_order): nothing to permute on the device
    const bool ap = hp.ap.on;                      // all-pairs flavour: the caller's frames, padded to 16-byte multiples if need be
    const bool plan_order = input_order == CV_ORDER_PLAN && hp.mode == CV_MODE_STAGED && !ap;
    const bool staged = hp.mode == CV_MODE_STAGED && !ap;
    const int64_t min_stride = plan_order ? hp.staged_floats : 3ll * hp.desc.n_atoms;
    const int64_t P = hp.P, A = hp.A, W = hp.W, F = P + A;
    const int box_kind = hp.desc.box_kind;
    const int box_floats = box_kind == CV_BOX_ORTHO ? 3 : (box_kind == CV_BOX_TRICLINIC ? 9 : 0);
    if (n_frames > 0 && !xyz_host) return fail(CV_E_INVALID, "null coordinates");
    if (box_kind != CV_BOX_NONE && n_frames > 0 && !box_host) return fail(CV_E_INVALID, "plan has a box but box_host is null");
    if (frame_stride_floats < min_stride)
        return fail(CV_E_INVALID, plan_order ? "frame stride smaller than the plan-order frame (staged_floats_per_frame)"
                                             : "frame stride smaller than 3*n_atoms");
    if (plan_order && frame_stride_floats % 4 != 0) return fail(CV_E_INVALID, "plan-order frames must be a multiple of 16 bytes apart");
    const bool want_agg = out->occupancy || out->sum || out->sumsq;
    const bool want_moments = out->sum || out->sumsq;
    if (want_agg && (!out->occupancy || (want_moments && !(out->sum && out->sumsq))))
        return fail(CV_E_INVALID, "aggregates: occupancy alone, or occupancy + sum + sumsq");
    const bool want_ws = out->score_q32 != nullptr;
    if (want_ws && !out->weights_q32) return fail(CV_E_INVALID, "score_q32 needs weights_q32");
    const bool need_flags = out->flags || want_ws;          // the weighted scores are computed from the flag words
    DeviceGuard guard;
    CV_CUDA(guard.enter(plan->device));

    // block size: ~256 MB of input + output per buffer set
    const int64_t in_row = frame_stride_floats * 4 + box_floats * 4;
    const int64_t out_row = (out->dist ? 4 * P : 0) + (out->angle ? 4 * A : 0) + (out->flags ? 4 * W : 0) + (out->score ? 4 : 0) + (want_ws ? 8 : 0);
    int64_t TB = frames_per_block > 0 ? frames_per_block : std::max<int64_t>(64, (256ll << 20) / std::max<int64_t>(1, in_row + out_row));
    TB = std::max<int64_t>(1, std::min<int64_t>(TB, std::max<int64_t>(n_frames, 1)));
    const bool in_pinned = is_pinned(xyz_host);
    const bool out_pinned = is_pinned(out->dist) && is_pinned(out->angle) && is_pinned(out->flags) && is_pinned(out->score) &&
                            is_pinned(out->score_q32);

    // device-side frame stride: the caller's, unless the all-pairs flavour needs padding (selection size not a multiple of
    // 4, or an odd stride): then the upload is a pitched copy into frames of 3 * ceil4(n_atoms) floats
    const int64_t ap_stride = 3ll * ((hp.desc.n_atoms + 3) / 4 * 4);
    const bool repitch = ap && (frame_stride_floats % 4 != 0 || frame_stride_floats < ap_stride);
    const int64_t dev_stride = repitch ? ap_stride : frame_stride_floats;

    Trace tr;
    HostCtx& hc = plan->host;
    std::lock_guard<std::mutex> lock(hc.mu);           // one host call per plan at a time
    HostSet (&set)[2] = hc.set;
    DevBuf &ws = hc.ws, &occ = hc.occ, &sum = hc.sum, &sq = hc.sq, &box_all = hc.box_all, &rec_all = hc.rec_all, &wq = hc.wq;
    Stream &s_up = hc.s_up, &s_run = hc.s_run, &s_down = hc.s_down;
    typedef HostSet Set;
    CV_CUDA(s_up.create()); CV_CUDA(s_run.create()); CV_CUDA(s_down.create());
    CV_CUDA(ws.alloc((size_t)kQueueBytes + (size_t)hp.max_chunks * F * 24));
    for (Set& s : set) {
        CV_CUDA(s.xyz.alloc((size_t)TB * dev_stride * 4));
        CV_CUDA(s.staged.alloc(staged && !plan_order ? (size_t)TB * hp.staged_floats * 4 : 0));
        CV_CUDA(s.dist.alloc(out->dist ? (size_t)TB * P * 4 : 0));
        CV_CUDA(s.angle.alloc(out->angle ? (size_t)TB * A * 4 : 0));
        CV_CUDA(s.flags.alloc(need_flags ? (size_t)TB * W * 4 : 0));
        CV_CUDA(s.wscore.alloc(want_ws ? (size_t)TB * 8 : 0));
        CV_CUDA(s.score.alloc(out->score ? (size_t)TB * 4 : 0));
        if (!in_pinned) {
            CV_CUDA(s.h_in.alloc((size_t)TB * frame_stride_floats * 4));
        }
        if (!out_pinned) {
            CV_CUDA(s.h_dist.alloc(out->dist ? (size_t)TB * P * 4 : 0));
            CV_CUDA(s.h_angle.alloc(out->angle ? (size_t)TB * A * 4 : 0));
            CV_CUDA(s.h_flags.alloc(out->flags ? (size_t)TB * W * 4 : 0));
            CV_CUDA(s.h_score.alloc(out->score ? (size_t)TB * 4 : 0));
            CV_CUDA(s.h_wscore.alloc(want_ws ? (size_t)TB * 8 : 0));
        }
        CV_CUDA(s.uploaded.create()); CV_CUDA(s.computed.create()); CV_CUDA(s.downloaded.create());
    }
    if (want_agg) {
        CV_CUDA(occ.alloc((size_t)F * 8)); CV_CUDA(sum.alloc((size_t)F * 8)); CV_CUDA(sq.alloc((size_t)F * 8));
        CV_CUDA(cudaMemsetAsync(occ.p, 0, (size_t)F * 8, s_run.s));
        CV_CUDA(cudaMemsetAsync(sum.p, 0, (size_t)F * 8, s_run.s));
        CV_CUDA(cudaMemsetAsync(sq.p, 0, (size_t)F * 8, s_run.s));
    }
    // occupancy alone + no feature matrix = a lean (contact-map) run, see caviar_run
    const bool dev_moments = want_moments;

    tr.mark("features_host: buffers");
    if (want_ws) {
        CV_CUDA(wq.alloc((size_t)F * 8));
        CV_CUDA(cudaMemcpyAsync(wq.p, out->weights_q32, (size_t)F * 8, cudaMemcpyHostToDevice, s_run.s));
    }
    int64_t launches = 0;
    if (box_floats && n_frames > 0) {
        // boxes are tiny (<= 36 B per frame): upload and prepare all records once, outside the block loop
        CV_CUDA(box_all.alloc((size_t)n_frames * box_floats * 4));
        CV_CUDA(rec_all.alloc((size_t)n_frames * kRecFloats * 4));
        CV_CUDA(cudaMemcpyAsync(box_all.p, box_host, (size_t)n_frames * box_floats * 4, cudaMemcpyHostToDevice, s_run.s));
        int rc = caviar_prepare_box(plan, box_all.as<float>(), n_frames, rec_all.as<float>(), s_run.s);
        if (rc != CV_OK) return rc;
        launches += g_launches;
    }
    const int64_t n_blocks = (n_frames + TB - 1) / TB;
    struct Pending { int64_t t0 = 0, n = 0; bool live = false; } pending[2];

    auto drain = [&](int k) -> int {   // copy bounce buffers of set k out to the caller's arrays
        Pending& pd = pending[k];
        if (!pd.live) return CV_OK;
        CV_CUDA(cudaEventSynchronize(set[k].downloaded.e));
        if (!out_pinned) {
            if (out->dist) parallel_memcpy(out->dist + pd.t0 * P, set[k].h_dist.p, (size_t)pd.n * P * 4);
            if (out->angle) parallel_memcpy(out->angle + pd.t0 * A, set[k].h_angle.p, (size_t)pd.n * A * 4);
            if (out->flags) memcpy(out->flags + pd.t0 * W, set[k].h_flags.p, (size_t)pd.n * W * 4);
            if (out->score) memcpy(out->score + pd.t0, set[k].h_score.p, (size_t)pd.n * 4);
            if (want_ws) memcpy(out->score_q32 + pd.t0, set[k].h_wscore.p, (size_t)pd.n * 8);
        }
        pd.live = false;
        return CV_OK;
    };

    // On any early error return, in-flight copies into the caller's buffers must have finished first.
    struct Quiesce {
        Stream &a, &b, &c;
        bool armed = true;
        ~Quiesce() { if (armed) { cudaStreamSynchronize(a.s); cudaStreamSynchronize(b.s); cudaStreamSynchronize(c.s); } }
    } quiesce{s_up, s_run, s_down};

    for (int64_t b = 0; b < n_blocks; ++b) {
        const int k = (int)(b & 1);
        Set& s = set[k];
        const int64_t t0 = b * TB, n = std::min<int64_t>(TB, n_frames - t0);
        int rc = drain(k);                       // set k is free again once its previous download finished
        if (rc != CV_OK) return rc;
        // upload
        const float* src = xyz_host + t0 * frame_stride_floats;
        if (!in_pinned) {
            parallel_memcpy(s.h_in.p, src, (size_t)n * frame_stride_floats * 4);
            src = s.h_in.as<float>();
        }
        if (repitch)
            CV_CUDA(cudaMemcpy2DAsync(s.xyz.p, (size_t)dev_stride * 4, src, (size_t)frame_stride_floats * 4, (size_t)hp.desc.n_atoms * 12, (size_t)n,
                                      cudaMemcpyHostToDevice, s_up.s));
        else
            CV_CUDA(cudaMemcpyAsync(s.xyz.p, src, (size_t)n * frame_stride_floats * 4, cudaMemcpyHostToDevice, s_up.s));
        CV_CUDA(cudaEventRecord(s.uploaded.e, s_up.s));
        // compute
        CV_CUDA(cudaStreamWaitEvent(s_run.s, s.uploaded.e, 0));
        const float* coords = s.xyz.as<float>();
        if (staged && !plan_order) {
            rc = caviar_stage_frames(plan, s.xyz.as<float>(), frame_stride_floats, n, s.staged.as<float>(), s_run.s);
            if (rc != CV_OK) return rc;
            launches += g_launches;
            coords = s.staged.as<float>();
        }
        CvFrameOutputs fo{out->dist ? s.dist.as<float>() : nullptr, out->angle ? s.angle.as<float>() : nullptr,
                          need_flags ? s.flags.as<uint32_t>() : nullptr, out->score ? s.score.as<int32_t>() : nullptr};
        CvAggregates ag{occ.as<uint64_t>(), dev_moments ? sum.as<double>() : nullptr, dev_moments ? sq.as<double>() : nullptr};
        const int64_t run_stride = (staged && !plan_order) ? hp.staged_floats : dev_stride;
        rc = caviar_run(plan, coords, run_stride, box_floats ? rec_all.as<float>() + t0 * kRecFloats : nullptr, n, &fo, want_agg ? &ag : nullptr, ws.p, s_run.s);
        if (rc != CV_OK) return rc;
        launches += g_launches;
        if (want_ws) {
            rc = caviar_weighted_scores(plan, s.flags.as<uint32_t>(), n, wq.as<int64_t>(), s.wscore.as<int64_t>(), s_run.s);
            if (rc != CV_OK) return rc;
            launches += g_launches;
        }
        CV_CUDA(cudaEventRecord(s.computed.e, s_run.s));
        // download
        CV_CUDA(cudaStreamWaitEvent(s_down.s, s.computed.e, 0));
        float* ddst = out_pinned ? out->dist + t0 * P : s.h_dist.as<float>();
        float* adst = out_pinned ? out->angle + t0 * A : s.h_angle.as<float>();
        uint32_t* fdst = out_pinned ? out->flags + t0 * W : s.h_flags.as<uint32_t>();
        int32_t* sdst = out_pinned ? out->score + t0 : s.h_score.as<int32_t>();
        if (out->dist) CV_CUDA(cudaMemcpyAsync(ddst, s.dist.p, (size_t)n * P * 4, cudaMemcpyDeviceToHost, s_down.s));
        if (out->angle) CV_CUDA(cudaMemcpyAsync(adst, s.angle.p, (size_t)n * A * 4, cudaMemcpyDeviceToHost, s_down.s));
        if (out->flags) CV_CUDA(cudaMemcpyAsync(fdst, s.flags.p, (size_t)n * W * 4, cudaMemcpyDeviceToHost, s_down.s));
        if (out->score) CV_CUDA(cudaMemcpyAsync(sdst, s.score.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s_down.s));
        if (want_ws) CV_CUDA(cudaMemcpyAsync(out_pinned ? out->score_q32 + t0 : s.h_wscore.as<int64_t>(), s.wscore.p, (size_t)n * 8,
                                             cudaMemcpyDeviceToHost, s_down.s));
        CV_CUDA(cudaEventRecord(s.downloaded.e, s_down.s));
        pending[k] = Pending{t0, n, true};
        // the next upload into the other set may only start once that set's kernels are done: drain() above
        // waits for its download, which is ordered after its compute.
    }
    for (int k = 0; k < 2; ++k) { int rc = drain(k); if (rc != CV_OK) return rc; }
    tr.mark("features_host: block loop");
    if (want_agg) {
        CV_CUDA(cudaStreamSynchronize(s_run.s));
        CV_CUDA(cudaMemcpy(out->occupancy, occ.p, (size_t)F * 8, cudaMemcpyDeviceToHost));
        if (want_moments) {
            CV_CUDA(cudaMemcpy(out->sum, sum.p, (size_t)F * 8, cudaMemcpyDeviceToHost));
            CV_CUDA(cudaMemcpy(out->sumsq, sq.p, (size_t)F * 8, cudaMemcpyDeviceToHost));
        }
    }
    CV_CUDA(cudaStreamSynchronize(s_run.s));
    quiesce.armed = false;
    g_launches = launches;
    return CV_OK;
}

// The reference calls its operator several times with the same pair list (trajectory A, trajectory B:
// CAVIAR-1.0.py:337-338), so the last plan -- index tables on the device, the host entry point's device buffers,
// bounce buffers, streams -- is kept between calls.  One entry; the mutex also serialises concurrent callers.
namespace {
struct OperatorCache {
    std::mutex mu;
    CvPlan* plan = nullptr;
    int32_t n_atoms = 0;
    int device = -1;
    std::vector<int32_t> pairs;
};
OperatorCache& operator_cache() {
    static OperatorCache* c = new OperatorCache();     // never destroyed: the CUDA runtime may be gone at exit
    return *c;
}
}  // namespace

extern "C" void caviar_release_cache(void) {
    OperatorCache& oc = operator_cache();
    std::lock_guard<std::mutex> lock(oc.mu);
    if (oc.plan) { caviar_plan_destroy(oc.plan); oc.plan = nullptr; }
    oc.pairs.clear(); oc.pairs.shrink_to_fit();
}

extern "C" int caviar_pair_distances_host(const float* xyz_host, int64_t n_frames, int32_t n_atoms,
                                          const int32_t* pairs, int64_t n_pairs, float* dist_host) {
    if (!dist_host && n_frames > 0) return fail(CV_E_INVALID, "null output");
    if (n_pairs > 0 && !pairs) return fail(CV_E_INVALID, "null index array");
    Trace tr;
    OperatorCache& oc = operator_cache();
    std::lock_guard<std::mutex> lock(oc.mu);
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); cur = -1; }
    const bool hit = oc.plan && oc.n_atoms == n_atoms && oc.device == cur && (int64_t)oc.pairs.size() == 2 * n_pairs &&
                     (n_pairs == 0 || memcmp(oc.pairs.data(), pairs, (size_t)n_pairs * 8) == 0);
    if (!hit) {
        if (oc.plan) { caviar_plan_destroy(oc.plan); oc.plan = nullptr; }
        CvPlanDesc d{};
        d.struct_size = sizeof(CvPlanDesc);
        d.n_atoms = n_atoms; d.n_pairs = n_pairs; d.pairs = pairs;
        d.n_triplets = 0; d.triplets = nullptr;
        d.box_kind = CV_BOX_NONE;
        d.pair_cutoff = 0.0; d.hbond_dist_cutoff = 0.0; d.hbond_angle_cutoff = 0.0;
        d.device = -1;
        int rc = caviar_plan_create(&d, &oc.plan);
        if (rc != CV_OK) { oc.plan = nullptr; return rc; }
        oc.n_atoms = n_atoms; oc.device = oc.plan->device;
        oc.pairs.assign(pairs, pairs + 2 * n_pairs);
    }
    tr.mark(hit ? "pair_distances: cached plan" : "pair_distances: plan_create");
    CvHostOutputs o{};
    o.dist = dist_host;
    int rc = caviar_features_host(oc.plan, xyz_host, 3ll * n_atoms, nullptr, n_frames, &o, 0, CV_ORDER_CALLER);
    tr.mark("pair_distances: features_host");
    return rc;
}

extern "C" int caviar_pair_distances_host_f64(const double* xyz_host, int64_t n_frames, int32_t n_atoms,
                                              const int32_t* pairs, int64_t n_pairs, double* dist_host) {
    g_launches = 0;
    if (n_frames < 0 || n_atoms <= 0 || n_pairs <= 0) return fail(CV_E_INVALID, "bad shape");
    if (n_frames == 0) return CV_OK;
    if (!xyz_host || !pairs || !dist_host) return fail(CV_E_INVALID, "null argument");
    for (int64_t k = 0; k < 2 * n_pairs; ++k)
        if (pairs[k] < 0 || pairs[k] >= n_atoms) return fail(CV_E_INVALID, "pair index out of range");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(CV_E_NO_DEVICE, "no CUDA device: caviar_b200 has no CPU fallback");
    // grow-only device buffers and a stream kept between calls (the operator is called once per trajectory, :337-338);
    // blocks of frames are double-buffered: the upload of block b + 1 overlaps the kernel and download of block b
    struct F64Ctx { std::mutex mu; DevBuf pairs, xyz[2], out[2]; Stream up, run; Event uploaded[2], done[2]; };
    static F64Ctx* ctx = new F64Ctx();                     // never destroyed: the CUDA runtime may be gone at exit
    std::lock_guard<std::mutex> lock(ctx->mu);
    CV_CUDA(ctx->up.create()); CV_CUDA(ctx->run.create());
    CV_CUDA(ctx->pairs.alloc((size_t)n_pairs * 8));
    CV_CUDA(cudaMemcpyAsync(ctx->pairs.p, pairs, (size_t)n_pairs * 8, cudaMemcpyHostToDevice, ctx->run.s));
    const int64_t row_in = 24ll * n_atoms, row_out = 8ll * n_pairs;
    const int64_t TB = std::max<int64_t>(1, std::min<int64_t>(n_frames, (256ll << 20) / (row_in + row_out)));
    for (int k = 0; k < 2; ++k) {
        CV_CUDA(ctx->xyz[k].alloc((size_t)TB * row_in)); CV_CUDA(ctx->out[k].alloc((size_t)TB * row_out));
        CV_CUDA(ctx->uploaded[k].create()); CV_CUDA(ctx->done[k].create());
    }
    struct Quiesce { F64Ctx* c; ~Quiesce() { cudaStreamSynchronize(c->up.s); cudaStreamSynchronize(c->run.s); } } quiesce{ctx};
    int64_t b = 0;
    for (int64_t t0 = 0; t0 < n_frames; t0 += TB, ++b) {
        const int k = (int)(b & 1);
        const int64_t n = std::min<int64_t>(TB, n_frames - t0);
        if (b >= 2) CV_CUDA(cudaStreamWaitEvent(ctx->up.s, ctx->done[k].e, 0));       // buffer set k is free again
        CV_CUDA(cudaMemcpyAsync(ctx->xyz[k].p, xyz_host + t0 * 3ll * n_atoms, (size_t)n * row_in, cudaMemcpyHostToDevice, ctx->up.s));
        CV_CUDA(cudaEventRecord(ctx->uploaded[k].e, ctx->up.s));
        CV_CUDA(cudaStreamWaitEvent(ctx->run.s, ctx->uploaded[k].e, 0));
        dim3 grid((unsigned)((n_pairs + 255) / 256), (unsigned)std::min<int64_t>(n, 4096));
        pair_distances_f64_kernel<<<grid, 256, 0, ctx->run.s>>>(ctx->xyz[k].as<double>(), n, n_atoms, ctx->pairs.as<int2>(), n_pairs,
                                                                ctx->out[k].as<double>());
        CV_CUDA(cudaGetLastError());
        g_launches += 1;
        CV_CUDA(cudaMemcpyAsync(dist_host + t0 * n_pairs, ctx->out[k].p, (size_t)n * row_out, cudaMemcpyDeviceToHost, ctx->run.s));
        CV_CUDA(cudaEventRecord(ctx->done[k].e, ctx->run.s));
    }
    CV_CUDA(cudaStreamSynchronize(ctx->up.s));
    CV_CUDA(cudaStreamSynchronize(ctx->run.s));
    return CV_OK;
}

// ------------------------------------------------------------------------------------------------
// Host-side compaction of trajectory frames (the atom_slice of CAVIAR-1.0.py:335, at the file boundary)
// ------------------------------------------------------------------------------------------------
// out[t][k][:] = frame t of src, atom atoms[k], converted to native float32.  src is what a trajectory file maps
// into memory: frames frame_stride_bytes apart, atoms as 3 x 4-byte floats, big-endian for Amber NetCDF (XDR).
// Only the cache lines of the selected atoms are touched, frames are spread over the host threads: numpy needs a
// byte-swapping copy of the WHOLE frame block plus a fancy-index pass (1.5 GB/s, 0.8 ms per 100k-atom frame each).
extern "C" int caviar_compact_frames_host(const void* src, int64_t n_frames, int64_t frame_stride_bytes, int32_t n_atoms,
                                          const int32_t* atoms, int64_t n_sel, int32_t big_endian, float* out) {
    if (n_frames < 0 || n_sel < 0 || n_atoms <= 0) return fail(CV_E_INVALID, "bad shape");
    if (n_frames == 0 || n_sel == 0) return CV_OK;
    if (!src || !atoms || !out) return fail(CV_E_INVALID, "null argument");
    if (frame_stride_bytes < 12ll * n_atoms) return fail(CV_E_INVALID, "frame stride smaller than 12*n_atoms bytes");
    for (int64_t k = 0; k < n_sel; ++k)
        if (atoms[k] < 0 || atoms[k] >= n_atoms) return fail(CV_E_INVALID, "atom index out of range");
    unsigned hw = std::thread::hardware_concurrency();
    const int64_t work = n_frames * n_sel;
    int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(hw ? hw : 4, 32), std::min<int64_t>(n_frames, work / 65536 + 1)));
    auto run = [&](int k) {
        const int64_t t0 = n_frames * k / nt, t1 = n_frames * (k + 1) / nt;
        for (int64_t t = t0; t < t1; ++t) {
            const unsigned char* frame = static_cast<const unsigned char*>(src) + t * frame_stride_bytes;
            uint32_t* dst = reinterpret_cast<uint32_t*>(out + t * 3 * n_sel);
            for (int64_t a = 0; a < n_sel; ++a) {
                uint32_t v[3];
                memcpy(v, frame + 12ll * atoms[a], 12);
                if (big_endian) { v[0] = __builtin_bswap32(v[0]); v[1] = __builtin_bswap32(v[1]); v[2] = __builtin_bswap32(v[2]); }
                dst[3 * a] = v[0]; dst[3 * a + 1] = v[1]; dst[3 * a + 2] = v[2];
            }
        }
    };
    std::vector<std::thread> pool;
    for (int k = 1; k < nt; ++k) pool.emplace_back(run, k);
    run(0);
    for (auto& th : pool) th.join();
    return CV_OK;
}

// The same for trajectory formats that keep a frame as three separate coordinate arrays (DCD: X[N], Y[N], Z[N], each a
// Fortran record): x_off / y_off / z_off are the byte offsets of those arrays inside a frame.
extern "C" int caviar_compact_frames_soa_host(const void* src, int64_t n_frames, int64_t frame_stride_bytes, int32_t n_atoms,
                                              int64_t x_off, int64_t y_off, int64_t z_off, const int32_t* atoms,
                                              int64_t n_sel, int32_t byte_swap, float* out) {
    if (n_frames < 0 || n_sel < 0 || n_atoms <= 0 || x_off < 0 || y_off < 0 || z_off < 0) return fail(CV_E_INVALID, "bad shape");
    if (n_frames == 0 || n_sel == 0) return CV_OK;
    if (!src || !atoms || !out) return fail(CV_E_INVALID, "null argument");
    const int64_t last = std::max(x_off, std::max(y_off, z_off)) + 4ll * n_atoms;
    if (frame_stride_bytes < last) return fail(CV_E_INVALID, "frame stride smaller than the coordinate arrays");
    for (int64_t k = 0; k < n_sel; ++k)
        if (atoms[k] < 0 || atoms[k] >= n_atoms) return fail(CV_E_INVALID, "atom index out of range");
    unsigned hw = std::thread::hardware_concurrency();
    const int64_t work = n_frames * n_sel;
    int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(hw ? hw : 4, 32), std::min<int64_t>(n_frames, work / 65536 + 1)));
    auto run = [&](int k) {
        const int64_t t0 = n_frames * k / nt, t1 = n_frames * (k + 1) / nt;
        const int64_t off[3] = {x_off, y_off, z_off};
        for (int64_t t = t0; t < t1; ++t) {
            const unsigned char* frame = static_cast<const unsigned char*>(src) + t * frame_stride_bytes;
            uint32_t* dst = reinterpret_cast<uint32_t*>(out + t * 3 * n_sel);
            for (int c = 0; c < 3; ++c) {
                const unsigned char* arr = frame + off[c];
                for (int64_t a = 0; a < n_sel; ++a) {
                    uint32_t v;
                    memcpy(&v, arr + 4ll * atoms[a], 4);
                    dst[3 * a + c] = byte_swap ? __builtin_bswap32(v) : v;
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int k = 1; k < nt; ++k) pool.emplace_back(run, k);
    run(0);
    for (auto& th : pool) th.join();
    return CV_OK;
}

// ------------------------------------------------------------------------------------------------
// CSV rows (host): fixed-point formatting without printf
// ------------------------------------------------------------------------------------------------
namespace {
inline char* put_u64(char* p, unsigned long long v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
// "%.{dec}f" of a float.  (double)v * 10^dec is exact for dec <= 8 (24 + 27 bits < 53), and nearbyint() rounds
// half-to-even exactly like printf does on the exact binary value.  Values whose scaled magnitude does not fit 64 bits
// (and non-finite ones) go through snprintf; kFixedChars bounds what either path can write.
constexpr int kFixedChars = 64;       // sign + 39 integer digits of FLT_MAX + '.' + 17 decimals, rounded up
inline char* put_fixed(char* p, float v, int dec, double scale) {
    const double mag = std::fabs((double)v) * scale;
    if (!(mag < 1.8e19) || dec > 8) {
        const int n = snprintf(p, kFixedChars, "%.*f", dec, (double)v);
        return p + std::min(std::max(n, 0), kFixedChars - 1);
    }
    const unsigned long long q = (unsigned long long)std::nearbyint(mag);
    if (std::signbit(v)) *p++ = '-';                     // printf prints "-0.0000" for small negatives as well
    unsigned long long pw = 1;
    for (int i = 0; i < dec; ++i) pw *= 10;
    const unsigned long long ip = q / pw, fp = q % pw;
    p = put_u64(p, ip);
    if (dec > 0) {
        *p++ = '.';
        for (unsigned long long d = pw / 10; d; d /= 10) *p++ = (char)('0' + (fp / d) % 10);
    }
    return p;
}
}  // namespace

extern "C" int64_t caviar_format_csv(const float* values, int64_t rows, int64_t cols, int64_t row_stride_floats,
                                     int64_t first_index, int64_t index_step, int32_t decimals, int32_t crlf,
                                     char* out, int64_t cap) {
    if (rows < 0 || cols < 0 || decimals < 0 || decimals > 17) return fail(CV_E_INVALID, "bad csv shape");
    if (rows == 0) return 0;
    if (!values || !out) return fail(CV_E_INVALID, "null argument");
    const int64_t per_row = 24 + (int64_t)(kFixedChars + 1) * cols;     // worst case of one row
    const double scale = std::pow(10.0, decimals);
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(hw ? hw : 4, 32), rows * cols / 65536 + 1));
    // per-thread buffers grow geometrically from an estimate of the typical row (a few integer digits + the decimals);
    // only the worst case of ONE row is ever reserved ahead of the cursor
    struct Part { std::vector<char> buf; size_t len = 0; size_t size() const { return len; } const char* data() const { return buf.data(); } };
    std::vector<Part> parts(nt);
    std::vector<std::thread> pool;
    const int64_t typical_row = 12 + (int64_t)(decimals + 8) * cols;
    auto work = [&](int k) {
        const int64_t r0 = rows * k / nt, r1 = rows * (k + 1) / nt;
        Part& part = parts[k];
        part.buf.resize((size_t)((r1 - r0) * typical_row + per_row));
        for (int64_t r = r0; r < r1; ++r) {
            if (part.len + (size_t)per_row > part.buf.size()) part.buf.resize(part.buf.size() + part.buf.size() / 2 + (size_t)per_row);
            char* p = part.buf.data() + part.len;
            long long idx = first_index + r * index_step;
            if (idx < 0) { *p++ = '-'; idx = -idx; }
            p = put_u64(p, (unsigned long long)idx);
            const float* row = values + r * row_stride_floats;
            for (int64_t c = 0; c < cols; ++c) { *p++ = ','; p = put_fixed(p, row[c], decimals, scale); }
            if (crlf) *p++ = '\r';                               // csv.writer's line terminator (Extractor:196-198)
            *p++ = '\n';
            part.len = (size_t)(p - part.buf.data());
        }
    };
    for (int k = 1; k < nt; ++k) pool.emplace_back(work, k);
    work(0);
    for (auto& t : pool) t.join();
    int64_t total = 0;
    for (auto& b : parts) total += (int64_t)b.size();
    if (total > cap) return fail(CV_E_NOMEM, "csv buffer too small");
    char* o = out;
    for (auto& b : parts) { memcpy(o, b.data(), b.size()); o += b.size(); }
    return total;
}
