// Lagged covariance blocks of CAVIAR-1.0.py:396-402 (_cov_blocks: X0.T @ X0 / n, Xt.T @ Xt / n, X0.T @ Xt / n over
// centred per-frame distances) on the 5th-generation tensor cores: hand-written tcgen05 / TMEM / TMA for sm_100a.
//
// This is the one dense contraction next to the hot path.  The reference multiplies float32
// matrices with numpy's sgemm; a single TF32 pass (10-bit mantissa) would lose three decimal digits against that, so
// every operand is split once into two TF32-exact halves, x = hi + lo with hi = rna_tf32(x) and lo = x - hi (exact),
// and each product block is accumulated as  hi.hi + hi.lo + lo.hi  in the float32 accumulator held in TMEM ("3xTF32":
// the dropped lo.lo term and the truncation of lo are ~2^-21 relative, the class of sgemm's own rounding).
//
//   cov_split_kernel   X (n, K) float32 [- mu] -> HI, LO (n, Kp) float32 planes (centring fused; HBM-bound, once per X)
//   cov_gemm_kernel    C[i, j] (+)= scale * sum_t A[t, i] * B[t, j]  for one 128 x 128 tile of C per CTA; the rows t of
//                      A and B are views of the planes with any row stride (lagged pairs X[:-tau] / X[tau:], the even /
//                      odd frames of the reference's cross-validation, CAVIAR-1.0.py:409-421 -- no copies):
//                        warp 0   TMA producer: 3-D tensor maps stream [32 frames x 128 features] boxes of the four
//                                 planes (A.hi, A.lo, B.hi, B.lo) into a 3-stage shared-memory ring, 128-byte rows swizzled
//                                 in 32-byte chunks (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B = UMMA SWIZZLE_128B_BASE32B)
//                                 -- the operands are MN-major (features contiguous), which tcgen05 takes as they are:
//                                 no transposed copy of the (n, K) matrix is ever made;
//                        warp 1   MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::tf32
//                                 (M = 128, N = 128, K = 8) x 4 k-steps x 3 terms per stage, accumulating in TMEM;
//                                 tcgen05.commit releases the stage to the producer and, at the end, the accumulator;
//                        warps 2-5 epilogue: every few stages the partial tile is promoted TMEM -> float32 registers
//                                 (tcgen05.ld 32 lanes x 32 columns, round-to-nearest adds; two TMEM accumulators
//                                 alternate so this overlaps the next MMAs), then scale by 1/n and 512-byte row stores.
//
// Rows beyond n and columns beyond K are zero-filled by TMA (out-of-bounds boxes), so ragged shapes need no masks in
// the main loop; the epilogue masks its stores.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <string>

#include "../../include/caviar_b200.h"

namespace cvcov {

constexpr int kTileM = 128;                     // C tile per CTA = UMMA M x N; N is 128 or 256 (template parameter TN)
constexpr int kBlockK = 32;                     // frames per pipeline stage (4 UMMA k-steps of 8)
constexpr int kUmmaK = 8;                       // tf32: 32 bytes of K per instruction
constexpr int kPanel = 32;                      // features per 128-byte swizzle row
constexpr int kPanelBytes = kBlockK * 128;      // one [32 frames x 32 features] panel: 4 KB = 8 swizzle atoms of 4 rows x 128 B
constexpr int kPlaneBytes = (kTileM / kPanel) * kPanelBytes;     // one A plane of a stage: 16 KB (a B plane: TN / 128 times that)
// Per tile width TN: a stage holds A.hi, A.lo (16 KB each) and B.hi, B.lo (TN / 128 x 16 KB each).  With TN = 128 every MMA
// re-reads 4 KB of A and 4 KB of B from shared memory for 67 clocks of math -- shared-memory bandwidth (128 B/clk) and
// the tensor core are then balanced; TN = 256 halves the A traffic per flop and is tensor-bound.
template <int TN> struct Geo {
    static constexpr int kBPlaneBytes = (TN / kPanel) * kPanelBytes;
    static constexpr int kStageBytes = 2 * kPlaneBytes + 2 * kBPlaneBytes;          // 64 KB / 96 KB
    static constexpr int kStages = TN == 128 ? 3 : 2;
    static constexpr int kTmemCols = 2 * TN;                                        // two alternating accumulators
    static constexpr int kEpiWarps = 4 * (TN / 128);                                // 128 accumulator columns per thread
    static constexpr int kThreads = 64 + 32 * kEpiWarps;                            // producer + issuer + epilogue warps
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;
    // Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 [4,6), A = B = TF32 [7,10) [10,13), A and B MN-major
    // (bits 15, 16), N >> 3 at [17,23), M >> 4 at [24,29)
    static constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                           ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// "all MMAs issued so far are done" -> one arrival on an mbarrier (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Shared-memory matrix descriptor of an MN-major operand plane (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 = distance between 32-feature panels,
//   [32,46) stride byte offset >> 4 = distance between swizzle atoms along K, [46,48) version = 1,
//   [61,64) layout type 1 = SWIZZLE_128B_BASE32B: 128-byte rows swizzled in 32-byte chunks, 4 rows per atom -- the only
//   swizzle the tensor core accepts for MN-major 32-bit operands (plain SWIZZLE_128B silently yields zeros: measured),
//   written by TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
#ifdef CV_COV_DEBUG
__constant__ uint32_t dbg_lbo = kPanelBytes, dbg_sbo = 512;
#define CV_LBO dbg_lbo
#define CV_SBO dbg_sbo
#else
#define CV_LBO ((uint32_t)kPanelBytes)
#define CV_SBO 512u
#endif
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(CV_LBO >> 4) << 16) | ((uint64_t)(CV_SBO >> 4) << 32) |
           (1ull << 46) | (1ull << 61);
}

__global__ void __launch_bounds__(256)
cov_split_kernel(const float* __restrict__ x, long long ld_x, const float* __restrict__ mu, long long n, int K, int Kp,
                 float* __restrict__ hi, float* __restrict__ lo) {
    const int c4 = blockIdx.x * blockDim.x + threadIdx.x;            // group of 4 columns
    if (4 * c4 >= Kp) return;
    float m[4] = {0.f, 0.f, 0.f, 0.f};
    if (mu != nullptr)
        for (int k = 0; k < 4; ++k) if (4 * c4 + k < K) m[k] = __ldg(mu + 4 * c4 + k);
    for (long long t = blockIdx.y; t < n; t += gridDim.y) {
        float4 h, l;
        float* hp = &h.x;
        float* lp = &l.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = 4 * c4 + k;
            const float v = (c < K) ? __fsub_rn(__ldg(x + t * ld_x + c), m[k]) : 0.f;     // dist - mu (CAVIAR-1.0.py:381-382)
            uint32_t hb;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
            const float hv = __uint_as_float(hb & 0xFFFFE000u);
            hp[k] = hv;
            lp[k] = __fsub_rn(v, hv);                                                     // exact: |lo| <= 2^-11 |v|
        }
        reinterpret_cast<float4*>(hi + t * (long long)Kp)[c4] = h;
        reinterpret_cast<float4*>(lo + t * (long long)Kp)[c4] = l;
    }
}

struct GemmParams {
    long long n;            // frames (reduction length)
    int K;                  // valid columns of C in both directions
    long long ld_c;         // floats between rows of C
    float scale;            // e.g. 1 / max(1, n)
    int accumulate;         // 1: C += scale * A^T B (stacked ensembles, CAVIAR-1.0.py:421), 0: C = ...
    int flush_kblocks;      // k-blocks (of 32 frames) accumulated in TMEM before the partial tile is promoted to registers
    float* c;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// The tensor core adds into its float32 accumulator with truncation: over a 14 000-frame reduction that alone costs
// 1e-4 relative (cuBLAS's TF32 path shows the same figure), whatever the operand precision.  So the accumulator chain
// in TMEM is kept SHORT -- `flush_kblocks` stages of 32 frames -- and the partial tiles are promoted into float32
// registers of the epilogue warps with round-to-nearest adds, exactly what a float32 sgemm does.  Two TMEM accumulators
// alternate, so draining partial tile c overlaps the MMAs of partial tile c + 1.
template <int TN>
__global__ void __launch_bounds__(Geo<TN>::kThreads, 1)
cov_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                const GemmParams p) {
    extern __shared__ unsigned char smem_raw[];
    // swizzle atoms must sit on their natural boundaries: align the ring to 1 KB
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int kStages = Geo<TN>::kStages, kStageBytes = Geo<TN>::kStageBytes, kBPlaneBytes = Geo<TN>::kBPlaneBytes;
    constexpr int kTmemCols = Geo<TN>::kTmemCols, kTileN = TN;
    constexpr uint32_t kInstrDesc = Geo<TN>::kInstrDesc;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* bar_empty = bar_full + kStages;
    uint64_t* bar_acc_full = bar_empty + kStages;          // [2] partial tile complete in TMEM buffer b
    uint64_t* bar_acc_empty = bar_acc_full + 2;            // [2] TMEM buffer b drained by the 4 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * kTileM, n0 = blockIdx.x * kTileN;
    const int n_kblocks = (int)((p.n + kBlockK - 1) / kBlockK);
    const int G = p.flush_kblocks;
    const int n_chunks = (n_kblocks + G - 1) / G;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_acc_full + b, 1); mbar_init(bar_acc_empty + b, Geo<TN>::kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {            // one warp allocates the accumulator columns and later frees them
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int kb = 0; kb < n_kblocks; ++kb) {
                const int s = kb % kStages;
                if (kb >= kStages) mbar_wait(bar_empty + s, ((kb / kStages) - 1) & 1);
                unsigned char* st = smem + s * kStageBytes;
                mbar_expect_tx(bar_full + s, kStageBytes);
                // box = {32 features, 32 frames, 4 (A) or TN / 32 (B) panels}: coordinates (0, first frame, first panel)
                tma_load_3d(st, &map_a_hi, bar_full + s, 0, kb * kBlockK, m0 / kPanel);
                tma_load_3d(st + kPlaneBytes, &map_a_lo, bar_full + s, 0, kb * kBlockK, m0 / kPanel);
                tma_load_3d(st + 2 * kPlaneBytes, &map_b_hi, bar_full + s, 0, kb * kBlockK, n0 / kPanel);
                tma_load_3d(st + 2 * kPlaneBytes + kBPlaneBytes, &map_b_lo, bar_full + s, 0, kb * kBlockK, n0 / kPanel);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one elected thread) =====================
        if (lane == 0) {
            int kb = 0;
            for (int c = 0; c < n_chunks; ++c) {
                const int buf = c & 1;
                if (c >= 2) { mbar_wait(bar_acc_empty + buf, ((c >> 1) - 1) & 1); tcgen05_fence_after(); }
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * kTileN);
                const int kb_end = min(kb + G, n_kblocks);
                for (int first = 1; kb < kb_end; ++kb) {
                    const int s = kb % kStages;
                    mbar_wait(bar_full + s, (kb / kStages) & 1);
                    tcgen05_fence_after();
                    const uint32_t base = smem_u32(smem + s * kStageBytes);
#pragma unroll
                    for (int kk = 0; kk < kBlockK / kUmmaK; ++kk) {
                        // one k-step = 8 frames = two 4-row swizzle atoms: 1 KB further into every panel
                        const uint64_t a_hi = smem_desc(base + kk * 1024), a_lo = smem_desc(base + kPlaneBytes + kk * 1024);
                        const uint64_t b_hi = smem_desc(base + 2 * kPlaneBytes + kk * 1024);
                        const uint64_t b_lo = smem_desc(base + 2 * kPlaneBytes + kBPlaneBytes + kk * 1024);
                        umma_tf32(tmem_d, a_lo, b_hi, kInstrDesc, first ? 0u : 1u);      // small terms first
                        umma_tf32(tmem_d, a_hi, b_lo, kInstrDesc, 1);
                        umma_tf32(tmem_d, a_hi, b_hi, kInstrDesc, 1);
                        first = 0;
                    }
                    umma_commit(bar_empty + s);              // the stage may be refilled once these MMAs have read it
                }
                umma_commit(bar_acc_full + buf);             // partial tile complete
            }
        }
    } else {
        // ===================== epilogue: TMEM -> float32 registers (promotion), then global =====================
        // a warp may only touch the TMEM lanes of its own quarter: lanes [32 * (warp % 4), + 32)
        // (two warps share a quarter when TN = 256: each takes 128 of the accumulator's columns)
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;                  // 0 for TN = 128
        constexpr int kCols = 128;
        float acc[kCols];
#pragma unroll
        for (int j = 0; j < kCols; ++j) acc[j] = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
            const int buf = c & 1;
            mbar_wait(bar_acc_full + buf, (c >> 1) & 1);
            tcgen05_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < kCols; c0 += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * kTileN + half * kCols + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[c0 + j] = __fadd_rn(acc[c0 + j], __uint_as_float(v[j]));
            }
            tcgen05_fence_before();
            if (lane == 0) mbar_arrive(bar_acc_empty + buf);     // this warp's quarter of the buffer may be overwritten
        }
        const int row = m0 + 32 * q + lane;
        if (row < p.K) {
            const int nc0 = n0 + half * kCols;
            float* dst = p.c + (long long)row * p.ld_c + nc0;
            const bool vec = (p.ld_c & 3) == 0 && (reinterpret_cast<uintptr_t>(p.c) & 15) == 0;
#pragma unroll
            for (int j = 0; j < kCols; j += 4) {
                if (vec && nc0 + j + 4 <= p.K) {
                    float4 o = p.accumulate ? *reinterpret_cast<const float4*>(dst + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    o.x = fmaf(acc[j], p.scale, o.x); o.y = fmaf(acc[j + 1], p.scale, o.y);
                    o.z = fmaf(acc[j + 2], p.scale, o.z); o.w = fmaf(acc[j + 3], p.scale, o.w);
                    *reinterpret_cast<float4*>(dst + j) = o;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (nc0 + j + i < p.K) dst[j + i] = fmaf(acc[j + i], p.scale, p.accumulate ? dst[j + i] : 0.f);
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            sym = nullptr;
        return reinterpret_cast<EncodeTiledFn>(sym);
    }();
    return fn;
}

// (n, Kp) row-major float32 plane seen as {32 features, n frames, Kp / 32 panels}: a box {32, 32, 4} lands in shared
// memory panel by panel, each panel 32 rows of 128 bytes, 128-byte swizzled -- the canonical MN-major UMMA operand.
static bool make_map(CUtensorMap* map, const float* plane, long long n, long long row_stride_floats, int Kp, int box_panels) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)kPanel, (cuuint64_t)n, (cuuint64_t)(Kp / kPanel)};
    const cuuint64_t strides[2] = {(cuuint64_t)row_stride_floats * 4, (cuuint64_t)kPanel * 4};
    const cuuint32_t box[3] = {(cuuint32_t)kPanel, (cuuint32_t)kBlockK, (cuuint32_t)box_panels};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(plane), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace cvcov

extern "C" int caviar_set_error(int code, const char* msg);     // cv_api.cu: thread-local last error

extern "C" int64_t caviar_cov_padded_columns(int64_t n_cols) {
    return (n_cols + cvcov::kTileM - 1) / cvcov::kTileM * cvcov::kTileM;
}

extern "C" int caviar_cov_split(const float* x_dev, int64_t ld_x, const float* mu_dev, int64_t n_rows, int64_t n_cols,
                                float* hi_dev, float* lo_dev, void* stream) {
    if (n_rows < 0 || n_cols <= 0 || ld_x < n_cols) return caviar_set_error(CV_E_INVALID, "cov_split: bad shape");
    if (n_rows == 0) return CV_OK;
    if (!x_dev || !hi_dev || !lo_dev) return caviar_set_error(CV_E_INVALID, "cov_split: null argument");
    if ((reinterpret_cast<uintptr_t>(hi_dev) | reinterpret_cast<uintptr_t>(lo_dev)) & 127)
        return caviar_set_error(CV_E_INVALID, "cov_split: planes must be 128-byte aligned");
    if (n_cols > (1 << 30)) return caviar_set_error(CV_E_INVALID, "cov_split: too many columns");
    const int Kp = (int)caviar_cov_padded_columns(n_cols);
    dim3 grid((Kp / 4 + 255) / 256, (unsigned)std::min<int64_t>(n_rows, 16384));
    cvcov::cov_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, ld_x, mu_dev, n_rows, (int)n_cols, Kp, hi_dev, lo_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return caviar_set_error(CV_E_CUDA, cudaGetErrorString(e));
    return CV_OK;
}

extern "C" int caviar_cov_gemm(const float* a_hi_dev, const float* a_lo_dev, const float* b_hi_dev, const float* b_lo_dev,
                               int64_t n_rows, int64_t row_stride_floats, int64_t n_cols, float scale, int32_t accumulate,
                               float* c_dev, int64_t ld_c, void* stream) {
    using namespace cvcov;
    if (n_rows <= 0 || n_cols <= 0 || ld_c < n_cols) return caviar_set_error(CV_E_INVALID, "cov_gemm: bad shape");
    if (row_stride_floats < caviar_cov_padded_columns(n_cols) || (row_stride_floats & 3))
        return caviar_set_error(CV_E_INVALID, "cov_gemm: row stride must be >= the padded column count and a multiple of 4 floats");
    if (!a_hi_dev || !a_lo_dev || !b_hi_dev || !b_lo_dev || !c_dev) return caviar_set_error(CV_E_INVALID, "cov_gemm: null argument");
    if ((reinterpret_cast<uintptr_t>(a_hi_dev) | reinterpret_cast<uintptr_t>(a_lo_dev) | reinterpret_cast<uintptr_t>(b_hi_dev) |
         reinterpret_cast<uintptr_t>(b_lo_dev)) & 15)
        return caviar_set_error(CV_E_INVALID, "cov_gemm: planes must be 16-byte aligned");
    if (n_rows >= (1ll << 31)) return caviar_set_error(CV_E_INVALID, "cov_gemm: too many rows");
    const int Kp = (int)caviar_cov_padded_columns(n_cols);
    // 128 x 256 tiles (tensor-bound) when C is wide enough to fill the SMs with them, else 128 x 128; CAVIAR_COV_TILE_N overrides
    static const int forced_tn = [] { const char* e = getenv("CAVIAR_COV_TILE_N"); return e ? atoi(e) : 0; }();
    const int tn = forced_tn == 128 || forced_tn == 256 ? forced_tn : ((long long)(Kp / 256) * (Kp / 128) >= 148 ? 256 : 128);
    CUtensorMap maps[4];
    const float* planes[4] = {a_hi_dev, a_lo_dev, b_hi_dev, b_lo_dev};
    for (int i = 0; i < 4; ++i)
        if (!make_map(&maps[i], planes[i], n_rows, row_stride_floats, Kp, i < 2 ? kTileM / kPanel : tn / kPanel))
            return caviar_set_error(CV_E_CUDA, "cov_gemm: cuTensorMapEncodeTiled failed (driver entry point missing or bad plane)");
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(cov_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<128>::kSmemBytes);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(cov_gemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<256>::kSmemBytes);
    });
    if (attr_err != cudaSuccess) return caviar_set_error(CV_E_CUDA, cudaGetErrorString(attr_err));
    GemmParams p;
    p.n = n_rows; p.K = (int)n_cols; p.ld_c = ld_c; p.scale = scale; p.accumulate = accumulate ? 1 : 0; p.c = c_dev;
    // frames per TMEM accumulation chain = 32 * flush_kblocks (CAVIAR_COV_FLUSH overrides for experiments).  Measured at
    // n = 14 000, K = 4 096 (max error relative to the diagonal, vs float64): 1 -> 1.0e-6 / 2.31 ms, 2 -> 1.3e-6 / 2.09 ms,
    // 4 -> 1.5e-6 / 1.96 ms, 8 -> 2.5e-6 / 1.89 ms, 64 -> 2.1e-5 / 1.82 ms; cuBLAS float32 6.8e-6 / 7.85 ms, TF32 1.1e-4.
    static const int flush = [] { const char* e = getenv("CAVIAR_COV_FLUSH"); return std::max(1, e ? atoi(e) : 4); }();
    p.flush_kblocks = flush;
    if (tn == 256) {
        dim3 grid((Kp + 255) / 256, Kp / kTileM);
        cov_gemm_kernel<256><<<grid, Geo<256>::kThreads, Geo<256>::kSmemBytes, (cudaStream_t)stream>>>(maps[0], maps[1], maps[2], maps[3], p);
    } else {
        dim3 grid(Kp / 128, Kp / kTileM);
        cov_gemm_kernel<128><<<grid, Geo<128>::kThreads, Geo<128>::kSmemBytes, (cudaStream_t)stream>>>(maps[0], maps[1], maps[2], maps[3], p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return caviar_set_error(CV_E_CUDA, cudaGetErrorString(e));
    return CV_OK;
}
