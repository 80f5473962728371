// Host-side planning for the CAVIAR pair-feature path (pure C++, no CUDA calls).
//
// A plan cuts the feature list (pairs first, then triplets; column order = input order, CAVIAR-1.0.py:122)
// into column tiles of <= 896 features, decides how coordinates reach the SMs (per-tile blocks of raw float32 xyz in
// PLAN ORDER streamed by TMA / one shared block or the caller's frame streamed whole / gathers from global memory for
// large overlapping selections), lays out the plan order (bank-conflict-free atom order inside every block) and sizes
// the TMA ring and the work queue.  The plan order is published (src_off / caviar_plan_atom_order) so that whoever
// slices the atoms out of the trajectory (CAVIAR-1.0.py:335) can hand the frames over in that order.  (tile, frame chunk) work items are handed to the persistent
// CTAs at run time, so that the atom offsets of a tile stay in registers for a whole item.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/caviar_b200.h"
#include "cv_common.h"

namespace cv {

// The all-pairs flavour (cv_allpairs.cuh) of a plan whose pair list is the reference's own i-major list.
struct ApHost {
    bool on = false;
    int minsep = 0, n_rows = 0;
    std::vector<int32_t> row_off;       // [n_rows + 1] first pair of row i
    std::vector<int32_t> row_of_word;   // [ceil(P / 32)] row of pair 32 w
    std::vector<ApTile> tiles;          // fast tiles first, general tiles last (the queue hands the last ones out first)
    int n_fast_windows = 0, n_windows = 0, n_shared_tiles = 0, n_fast_tiles = 0;   // tiles: [shared | fast | general], n_fast_tiles counts the first two
    int frames_per_stage = 0, n_stages = 0, stage_floats = 0, smem_bytes = 0, occupancy = 2, n_ctas = 0;
};

struct HostPlan {
    CvPlanDesc desc{};
    std::vector<int32_t> pairs, triplets;
    int mode = CV_MODE_DIRECT;
    int n_sms = 148;
    int occupancy = 2;
    int n_ctas = 0;
    int n_pair_tiles = 0, n_trip_tiles = 0, pair_tile_w = kTileWidth, trip_tile_w = kTileWidth;
    std::vector<TileDesc> tiles;
    std::vector<int32_t> idx0, idx1, idx2;
    std::vector<int32_t> src_off;      // staged float j of a frame <- float src_off[j] of the input frame
    int64_t staged_floats = 0;
    int64_t n_distinct = 0;
    int64_t bank_conflict_refs = 0;    // gather references left on an already-used shared-memory bank (staged mode)
    int frames_per_stage = 1, n_stages = 0, stage_floats = 0, smem_bytes = 0, whole_frame = 0;
    int max_chunks = 1;                // upper bound of frame chunks per run (sizes the workspace)
    int64_t P = 0, A = 0, W = 0, Wp = 0;
    float pair_thr = 0.f, pair_thr2 = 0.f, hb_dthr = 0.f, hb_athr = 0.f;
    ApHost ap;
};

inline float f32_at_or_above(double c) {     // smallest float t with (double)t >= c:  (double)d < c  <=>  d < t
    if (std::isnan(c)) return -INFINITY;     // never true
    float t = (float)c;
    if ((double)t < c) t = std::nextafterf(t, INFINITY);
    return t;
}
inline float f32_at_or_below(double c) {     // largest float t with (double)t <= c:  (double)a > c  <=>  a > t
    if (std::isnan(c)) return INFINITY;
    float t = (float)c;
    if ((double)t > c) t = std::nextafterf(t, -INFINITY);
    return t;
}

// smallest float s with sqrtf(s) >= t (sqrtf is correctly rounded and monotone):  s < result  <=>  sqrtf(s) < t
inline float sq_threshold(float t) {
    if (!(t > 0.f)) return 0.f;                       // sqrt(s) < t never holds for t <= 0 (or NaN): s < 0 never does
    if (std::isinf(t)) return INFINITY;
    float s = t * t;                                  // close; walk to the exact boundary
    while (std::sqrt(s) >= t && s > 0.f) s = std::nextafterf(s, 0.f);
    while (std::sqrt(s) < t) s = std::nextafterf(s, INFINITY);
    return s;
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

inline void split_tiles(int64_t n, int& n_tiles, int& width) {
    if (n <= 0) { n_tiles = 0; width = kTileWidth; return; }
    int64_t c = (n + kTileWidth - 1) / kTileWidth;
    int64_t w = (n + c - 1) / c;
    width = round_up((int)w, 32);
    n_tiles = (int)((n + width - 1) / width);
}

// Where does each distinct atom of a tile sit inside the tile's staged block?  The consumers gather with LDS from
// 3*slot (+0,+1,+2): the 32 lanes of one gather instruction read the atoms of 32 consecutive tile-local features in
// one role, and two lanes collide on a shared-memory bank exactly when their slots are congruent mod 32 (3 is
// invertible mod 32).  The block order is ours to choose, so atoms get a "colour" = slot mod 32 such that the atoms
// of one gather have distinct colours wherever possible (greedy, least-used colour first to keep the classes
// balanced); slot = 32*row + colour with each class sorted by atom index.  A block sorted by atom index (the first
// version) measured 55 % of all shared-memory wavefronts as bank-conflict replays.
// Returns the number of references that still collide (atoms shared between gathers that had been coloured before).
inline int64_t assign_bank_slots(const HostPlan& hp, const TileDesc& t, const std::vector<int32_t>& ua,
                                 std::vector<int32_t>& slot_of, int& n_slots) {
    const int n = (int)ua.size();
    const int roles = t.kind == KIND_PAIR ? 2 : 3;
    const int32_t* src = t.kind == KIND_PAIR ? hp.pairs.data() + 2ll * t.f_begin : hp.triplets.data() + 3ll * t.f_begin;
    std::vector<int8_t> colour((size_t)n, -1);
    int count[32] = {0};
    int64_t collisions = 0;
    int32_t members[32];
    for (int g0 = 0; g0 < t.n_feat; g0 += 32) {
        const int g1 = std::min(t.n_feat, g0 + 32);
        for (int r = 0; r < roles; ++r) {
            uint32_t used = 0;
            int nm = 0;
            for (int s = g0; s < g1; ++s) {
                const int32_t a = (int32_t)(std::lower_bound(ua.begin(), ua.end(), src[(size_t)roles * s + r]) - ua.begin());
                bool dup = false;
                for (int m = 0; m < nm; ++m) dup |= members[m] == a;
                if (dup) continue;                               // same atom = broadcast, no conflict
                members[nm++] = a;
                if (colour[a] >= 0) {
                    if (used & (1u << colour[a])) ++collisions;
                    used |= 1u << colour[a];
                }
            }
            for (int m = 0; m < nm; ++m) {
                const int32_t a = members[m];
                if (colour[a] >= 0) continue;
                int best = -1;
                for (int cc = 0; cc < 32; ++cc)
                    if (!(used & (1u << cc)) && (best < 0 || count[cc] < count[best])) best = cc;
                if (best < 0) {                                  // cannot happen with <= 32 members; keep it total
                    best = 0;
                    for (int cc = 1; cc < 32; ++cc) if (count[cc] < count[best]) best = cc;
                    ++collisions;
                }
                colour[a] = (int8_t)best; used |= 1u << best; ++count[best];
            }
        }
    }
    // rename the colours so that the fullest classes come first: the last row of the block is then filled from the left
    int order[32], name[32];
    for (int cc = 0; cc < 32; ++cc) order[cc] = cc;
    std::stable_sort(order, order + 32, [&](int x, int y) { return count[x] > count[y]; });
    for (int k = 0; k < 32; ++k) name[order[k]] = k;
    int row[32] = {0};
    slot_of.assign((size_t)n, 0);
    n_slots = 0;
    for (int i = 0; i < n; ++i) {                                // ua is sorted: each class stays sorted by atom index
        const int cc = name[colour[i] < 0 ? 0 : colour[i]];
        slot_of[i] = 32 * row[cc]++ + cc;
        n_slots = std::max(n_slots, slot_of[i] + 1);
    }
    return collisions;
}

inline void build_allpairs(HostPlan& hp);

// Returns "" on success, otherwise the reason.
inline std::string build_plan(const CvPlanDesc& d, int n_sms, HostPlan& hp) {
    if (d.struct_size != sizeof(CvPlanDesc)) return "CvPlanDesc.struct_size mismatch (ABI)";
    if (d.n_atoms <= 0) return "n_atoms must be positive";
    if (d.n_pairs < 0 || d.n_triplets < 0) return "negative feature count";
    if (d.n_pairs + d.n_triplets == 0) return "need at least one pair or triplet";
    if ((d.n_pairs > 0 && !d.pairs) || (d.n_triplets > 0 && !d.triplets)) return "null index array";
    if (d.box_kind < CV_BOX_NONE || d.box_kind > CV_BOX_TRICLINIC) return "unknown box_kind";
    if ((long long)d.n_atoms * 3 >= (1ll << 31)) return "n_atoms too large";
    if (d.n_pairs + d.n_triplets >= (1ll << 31) - kTileWidth) return "too many features";
    hp.desc = d;
    hp.P = d.n_pairs; hp.A = d.n_triplets;
    hp.pairs.assign(d.pairs, d.pairs + 2 * d.n_pairs);
    hp.triplets.assign(d.triplets, d.triplets + 3 * d.n_triplets);
    hp.desc.pairs = nullptr; hp.desc.triplets = nullptr;
    for (int32_t v : hp.pairs) if (v < 0 || v >= d.n_atoms) return "pair index out of range";
    for (int32_t v : hp.triplets) if (v < 0 || v >= d.n_atoms) return "triplet index out of range";
    hp.Wp = (hp.P + 31) / 32;
    hp.W = hp.Wp + (hp.A + 31) / 32;
    hp.pair_thr = f32_at_or_above(d.pair_cutoff);
    hp.pair_thr2 = sq_threshold(hp.pair_thr);
    hp.hb_dthr = f32_at_or_above(d.hbond_dist_cutoff);
    hp.hb_athr = f32_at_or_below(d.hbond_angle_cutoff);
    hp.n_sms = n_sms > 0 ? n_sms : 148;

    split_tiles(hp.P, hp.n_pair_tiles, hp.pair_tile_w);
    split_tiles(hp.A, hp.n_trip_tiles, hp.trip_tile_w);
    const int C = hp.n_pair_tiles + hp.n_trip_tiles;
    hp.tiles.resize(C);
    for (int c = 0; c < C; ++c) {
        TileDesc& t = hp.tiles[c];
        const bool trip = c >= hp.n_pair_tiles;
        const int ci = trip ? c - hp.n_pair_tiles : c;
        const int w = trip ? hp.trip_tile_w : hp.pair_tile_w;
        const int64_t n = trip ? hp.A : hp.P;
        t.kind = trip ? KIND_TRIPLET : KIND_PAIR;
        t.f_begin = ci * w;
        t.n_feat = (int)std::min<int64_t>(w, n - (int64_t)ci * w);
        t.idx_off = c * kTileWidth;
        t.grp_off = 0; t.grp_floats = 0; t.reserved[0] = t.reserved[1] = 0;
    }

    // distinct atoms per tile and overall
    std::vector<std::vector<int32_t>> tile_atoms(C);
    std::vector<uint8_t> seen((size_t)d.n_atoms, 0);
    int64_t sum_u = 0;
    int max_u = 0;
    for (int c = 0; c < C; ++c) {
        const TileDesc& t = hp.tiles[c];
        std::vector<int32_t>& a = tile_atoms[c];
        if (t.kind == KIND_PAIR) a.assign(hp.pairs.begin() + 2ll * t.f_begin, hp.pairs.begin() + 2ll * (t.f_begin + t.n_feat));
        else a.assign(hp.triplets.begin() + 3ll * t.f_begin, hp.triplets.begin() + 3ll * (t.f_begin + t.n_feat));
        std::sort(a.begin(), a.end());
        a.erase(std::unique(a.begin(), a.end()), a.end());
        for (int32_t v : a) seen[v] = 1;
        sum_u += (int64_t)a.size();
        max_u = std::max(max_u, (int)a.size());
    }
    hp.n_distinct = 0;
    for (uint8_t s : seen) hp.n_distinct += s;

    // ---- how do coordinates reach the SMs? ----
    // Three cases.  (1) Tiles that share few atoms get one coordinate block each (staged, TMA-streamed; an atom is
    // duplicated once per tile that references it).  (2) Tiles that overlap heavily -- all-pairs lists: every tile of
    // 896 columns needs ~900 atoms of the same selection -- share ONE block: the frame itself (zero-copy) or the
    // compacted selection, TMA-streamed whole while it is small.  (3) Beyond ~16 KB the whole-block refill per tile
    // and frame costs more than it saves (measured: all pairs of 3 000 / 5 000 atoms run 1.3x / 1.9x faster gathered
    // with LDG through L1/L2 than through a 36 / 60 KB whole-frame ring), so large overlapping selections are
    // gathered from the caller's frames in global memory.  Per-tile blocks for such lists would multiply the frame
    // thousands of times.
    const int64_t frame_bytes = 12ll * d.n_atoms;
    const int64_t u_pad_all = (hp.n_distinct + 3) / 4 * 4;
    constexpr int64_t kWholeBlockBytes = 16 * 1024;
    const bool heavy = sum_u > hp.n_distinct + hp.n_distinct / 2;
    const bool small_all = 12 * u_pad_all <= kWholeBlockBytes;
    // per-tile blocks would multiply the input bytes by more than ~8 (moderate sharing, e.g. random lists over few
    // atoms, is better served by per-tile blocks than by uncoalesced LDG gathers: 7x faster on config 3)
    const bool blow_up = sum_u > 8 * (hp.n_distinct + 1024);
    const bool zc_ok = (d.n_atoms % 4 == 0) && frame_bytes <= (C <= 2 ? 48 * 1024 : kWholeBlockBytes);
    bool shared_group = false;
    if (d.flags & CV_PLAN_FORCE_DIRECT) {
        hp.mode = CV_MODE_DIRECT;
    } else if (d.flags & CV_PLAN_FORCE_STAGED) {
        hp.mode = CV_MODE_STAGED;
        shared_group = heavy && small_all;
    } else if (zc_ok && ((int64_t)C * d.n_atoms <= 2 * sum_u + 1024 || (heavy && 4 * hp.n_distinct >= 3ll * d.n_atoms))) {
        hp.mode = CV_MODE_ZEROCOPY;          // tiles need most of the (small) frame anyway: stream it as it is
    } else if (heavy && small_all) {
        hp.mode = CV_MODE_STAGED; shared_group = true;
    } else if (blow_up) {
        hp.mode = CV_MODE_DIRECT;
    } else {
        hp.mode = CV_MODE_STAGED;
    }

    // ---- index tables (float offsets = 3 * atom slot) and the staged layout ----
    hp.idx0.assign((size_t)C * kTileWidth, 0);
    hp.idx1.assign((size_t)C * kTileWidth, 0);
    hp.idx2.assign((size_t)C * kTileWidth, 0);
    hp.src_off.clear();
    std::vector<int32_t> global_slot;
    if (hp.mode == CV_MODE_STAGED && shared_group) {
        global_slot.assign((size_t)d.n_atoms, -1);
        int32_t k = 0, last = 0;
        for (int32_t a = 0; a < d.n_atoms; ++a)
            if (seen[a]) { global_slot[a] = k++; last = a; hp.src_off.insert(hp.src_off.end(), {3 * a, 3 * a + 1, 3 * a + 2}); }
        for (; k < u_pad_all; ++k) hp.src_off.insert(hp.src_off.end(), {3 * last, 3 * last + 1, 3 * last + 2});
    }
    int max_grp_floats = 0;
    hp.bank_conflict_refs = 0;
    std::vector<int32_t> slot_of;      // per-tile: slot of ua[i] inside the tile's staged block
    for (int c = 0; c < C; ++c) {
        TileDesc& t = hp.tiles[c];
        const std::vector<int32_t>& ua = tile_atoms[c];
        auto rank = [&](int32_t atom) -> int32_t { return (int32_t)(std::lower_bound(ua.begin(), ua.end(), atom) - ua.begin()); };
        auto local = [&](int32_t atom) -> int32_t {
            if (hp.mode != CV_MODE_STAGED) return 3 * atom;
            if (shared_group) return 3 * global_slot[atom];
            return 3 * slot_of[rank(atom)];
        };
        if (hp.mode == CV_MODE_STAGED && !shared_group) {
            int n_slots = 0;
            hp.bank_conflict_refs += assign_bank_slots(hp, t, ua, slot_of, n_slots);
            const int u_pad = round_up(n_slots, 4);
            t.grp_off = (int32_t)hp.src_off.size();
            t.grp_floats = 3 * u_pad;
            std::vector<int32_t> atom_at((size_t)u_pad, ua.back());      // holes / padding repeat a valid atom
            for (size_t i = 0; i < ua.size(); ++i) atom_at[slot_of[i]] = ua[i];
            for (int k = 0; k < u_pad; ++k) {
                const int32_t a = atom_at[k];
                hp.src_off.insert(hp.src_off.end(), {3 * a, 3 * a + 1, 3 * a + 2});
            }
        } else if (hp.mode == CV_MODE_STAGED) {
            t.grp_off = 0; t.grp_floats = (int32_t)(3 * u_pad_all);
        } else {
            t.grp_off = 0; t.grp_floats = 3 * d.n_atoms;
        }
        max_grp_floats = std::max(max_grp_floats, t.grp_floats);
        for (int s = 0; s < t.n_feat; ++s) {
            const size_t o = (size_t)t.idx_off + s;
            const int64_t f = (int64_t)t.f_begin + s;
            if (t.kind == KIND_PAIR) {
                hp.idx0[o] = local(hp.pairs[2 * f]); hp.idx1[o] = local(hp.pairs[2 * f + 1]);
            } else {
                hp.idx0[o] = local(hp.triplets[3 * f]); hp.idx1[o] = local(hp.triplets[3 * f + 1]);
                hp.idx2[o] = local(hp.triplets[3 * f + 2]);
            }
        }
    }
    hp.staged_floats = (hp.mode == CV_MODE_STAGED) ? (int64_t)hp.src_off.size() : 0;
    hp.whole_frame = (hp.mode == CV_MODE_ZEROCOPY) || (hp.mode == CV_MODE_STAGED && shared_group);

    // ---- pipeline geometry ----
    if (hp.mode == CV_MODE_DIRECT) {
        hp.frames_per_stage = 1; hp.n_stages = 0; hp.stage_floats = 0; hp.smem_bytes = 0; hp.occupancy = 2;
    } else {
        const int gb = max_grp_floats * 4;
        const int rec_b = d.box_kind != CV_BOX_NONE ? kRecFloats * 4 : 0;
        const int bar_bytes = 2 * kMaxStages * 8 + kMaxStages * 24;   // mbarriers + per-stage work metadata
        const int frame_b = std::max(gb + rec_b, 1);
        // Ring geometry.  Per-stage costs (barrier wait, metadata, release) are paid per warp, so few LARGE stages win:
        // two stages of as many frames as fit beat four single-frame stages by 5 % on config 3.  2 CTAs/SM when two
        // stages of >= 1 frame fit in half of the shared memory (110 KB), else 1 CTA/SM.  CAVIAR_OCC / CAVIAR_FB /
        // CAVIAR_STAGES override for experiments.
        int want_occ = 2;
        if (const char* e = std::getenv("CAVIAR_OCC")) want_occ = std::max(1, std::min(2, atoi(e)));
        const int budget[3] = {0, 220 * 1024, 110 * 1024};
        int fb = 0, ns = 0;
        hp.occupancy = 0;
        for (int occ = want_occ; occ >= 1 && hp.occupancy == 0; --occ) {
            const int room = budget[occ] - bar_bytes;
            int f = std::min(16, room / (2 * frame_b));
            if (const char* e = std::getenv("CAVIAR_FB")) f = std::min(f, std::max(1, atoi(e)));
            if (f >= 1) {
                hp.occupancy = occ; fb = f;
                ns = std::max(2, std::min(f == 1 ? 4 : 2, room / (f * frame_b)));
                if (const char* e = std::getenv("CAVIAR_STAGES")) ns = std::max(2, std::min(std::min(atoi(e), kMaxStages), room / (f * frame_b)));
            }
        }
        if (hp.occupancy == 0) return "a column tile's coordinate block does not fit in shared memory";
        const int stage_total = fb * frame_b;
        hp.frames_per_stage = fb; hp.n_stages = ns; hp.stage_floats = fb * max_grp_floats;
        hp.smem_bytes = ns * stage_total + bar_bytes;
    }

    // ---- persistent grid + dynamic work queue ----
    // CTAs pull (frame chunk, tile) items from an atomic counter; chunks are sized so that every CTA gets about
    // kItemsPerCta items, which bounds the tail imbalance without inflating the partial-aggregate buffers.
    hp.n_ctas = d.n_ctas > 0 ? d.n_ctas : hp.n_sms * hp.occupancy;
    hp.max_chunks = std::max(1, (kItemsPerCta * hp.n_ctas + C - 1) / C);
    build_allpairs(hp);
    return "";
}


// ------------------------------------------------------------------------------------------------
// All-pairs plans
// ------------------------------------------------------------------------------------------------
// Is the pair list exactly [(i, j) for i in range(n) for j in range(i + minsep, n)] (CAVIAR-1.0.py:331)?
inline bool is_all_pairs(const HostPlan& hp, int& minsep) {
    const int64_t n = hp.desc.n_atoms, P = hp.P;
    if (P < 1 || hp.pairs[0] != 0) return false;
    const int64_t m = hp.pairs[1];
    if (m < 1 || m >= n) return false;
    const int64_t rows = n - m;
    if (P != rows * (rows + 1) / 2) return false;
    int64_t k = 0;
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = i + m; j < n; ++j, ++k)
            if (hp.pairs[2 * k] != i || hp.pairs[2 * k + 1] != j) return false;
    minsep = (int)m;
    return true;
}

// Cuts an all-pairs list into windows and 2-D tiles (cv_common.h).  Leaves hp.ap.on false when the list is not an
// all-pairs list, or when too few windows lie
// inside a single row for the geometry to pay (short rows: small selections stay on the generic kernel).
inline void build_allpairs(HostPlan& hp) {
    ApHost& ap = hp.ap;
    ap = ApHost();
    const CvPlanDesc& d = hp.desc;
    if (d.box_kind != CV_BOX_NONE || hp.A != 0 || (d.flags & (CV_PLAN_FORCE_DIRECT | CV_PLAN_FORCE_STAGED))) return;
    if (hp.P >= (1ll << 30)) return;              // (4 P bytes per row must fit 32 bits)
    // (a generic plan of the same list that is STAGED -- it would be, for some selection sizes that are not multiples of 4 --
    // wants plan-order frames; the all-pairs flavour takes the caller's frames, and caviar_run / caviar_features_host give
    // it precedence: see there)
    if (const char* e = std::getenv("CAVIAR_NO_ALLPAIRS")) if (atoi(e) != 0) return;
    int minsep = 0;
    if (!is_all_pairs(hp, minsep)) return;
    const int n = d.n_atoms;
    const int64_t P = hp.P;
    ap.minsep = minsep;
    ap.n_rows = n - minsep;
    ap.row_off.resize((size_t)ap.n_rows + 1);
    {
        int64_t o = 0;
        for (int i = 0; i < ap.n_rows; ++i) { ap.row_off[i] = (int32_t)o; o += n - minsep - i; }
        ap.row_off[ap.n_rows] = (int32_t)o;
    }
    ap.row_of_word.resize((size_t)hp.Wp);
    for (int64_t w = 0, r = 0; w < hp.Wp; ++w) {
        while (32 * w >= ap.row_off[r + 1]) ++r;
        ap.row_of_word[w] = (int32_t)r;
    }
    auto row_of = [&](int64_t k) { int r = ap.row_of_word[k >> 5]; while (k >= ap.row_off[r + 1]) ++r; return r; };

    // windows: fast = whole and inside one row
    const int nb = (int)((P + kApWindow - 1) / kApWindow);
    ap.n_windows = nb;
    struct Fast { int band, row, b, j0; };
    std::vector<Fast> fast;
    std::vector<int> gen;
    for (int b = 0; b < nb; ++b) {
        const int64_t k0 = (int64_t)b * kApWindow, k1 = std::min<int64_t>(P, k0 + kApWindow) - 1;
        const int i0 = row_of(k0), i1 = row_of(k1);
        if (i0 == i1 && k0 + kApWindow <= P) {
            const int j0 = i0 + minsep + (int)(k0 - ap.row_off[i0]);
            fast.push_back({j0 / kApWindow, i0, b, j0});
        } else {
            gen.push_back(b);
        }
    }
    // Windows -> warps (two each) -> tiles (7 warps).
    //  * SHARED warps (optional, see below): the same 128 columns of rows i and i + 256.  The first pairs of two rows 256 apart are a multiple
    //    of 128 pairs apart (256 L - 255 * 128), so their window grids coincide and atom j is loaded once for both rows.
    //    A tile is 7 such warps from consecutive rows of one band: atoms = 7 + 7 rows' i and ~270 j (up to three ranges).
    //  * FAST warps: what has no partner 256 rows away (the windows next to the diagonal, the last rows) pairs up with the
    //    window of a neighbouring row in the same band; tiles as above, 14 rows' i.
    //  * what is left over joins the general windows.  Fast tiles never straddle two bands.
    struct Warp { int b0, b1; bool fast; };
    struct FastWarp { int band, row, b0, b1; };
    std::vector<FastWarp> shared_w, fw;
    {
        // OFF by default: measured 7 % SLOWER on config 4 (contact map 1.131 against 1.216e12 evals/s, everything emitted
        // 5.63 against 6.11e11).  The kernel issues at ~68 % of its slots whatever the mix -- 40 % fewer LDS in two thirds of
        // the fast windows did not shorten it, the third launch and the extra half-filled tiles at the band ends cost more.
        // CAVIAR_AP_SHARE=1 builds the geometry (tests keep it correct).
        bool share = false;
        if (const char* e = std::getenv("CAVIAR_AP_SHARE")) share = atoi(e) != 0;
        // (adjacent windows of one row -- 1 KB pieces of the matrix -- were tried and lost 4 % to the wider column bands:
        // CAVIAR_AP_PAIRING=h builds that geometry)
        bool horizontal = false;
        if (const char* e = std::getenv("CAVIAR_AP_PAIRING")) horizontal = e[0] == 'h';
        std::vector<uint8_t> used(fast.size(), 0);
        if (share) {
            // fast is in list order (row-major, columns ascending): index of the first window of every row
            std::vector<int> row_first((size_t)ap.n_rows + 1, -1), row_cnt((size_t)ap.n_rows + 1, 0);
            for (size_t q = 0; q < fast.size(); ++q) { if (row_first[fast[q].row] < 0) row_first[fast[q].row] = (int)q; ++row_cnt[fast[q].row]; }
            for (size_t q = 0; q < fast.size(); ++q) {
                if (used[q]) continue;
                const int r2 = fast[q].row + 256;
                if (r2 >= ap.n_rows || row_first[r2] < 0) continue;
                for (int c = 0; c < row_cnt[r2]; ++c) {
                    const size_t q2 = (size_t)row_first[r2] + c;
                    if (!used[q2] && fast[q2].j0 == fast[q].j0) {
                        shared_w.push_back({fast[q].band, fast[q].row, fast[q].b, fast[q2].b});
                        used[q] = used[q2] = 1;
                        break;
                    }
                }
            }
            std::sort(shared_w.begin(), shared_w.end(), [](const FastWarp& x, const FastWarp& y) { return x.band != y.band ? x.band < y.band : x.row < y.row; });
        }
        std::vector<Fast> odd;
        for (size_t q = 0; q < fast.size();) {
            if (used[q]) { ++q; continue; }
            if (horizontal && q + 1 < fast.size() && !used[q + 1] && fast[q + 1].row == fast[q].row && fast[q + 1].b == fast[q].b + 1) {
                fw.push_back({1000 + fast[q].band / 2, fast[q].row, fast[q].b, fast[q + 1].b});
                q += 2;
            } else {
                odd.push_back(fast[q]); ++q;
            }
        }
        std::sort(odd.begin(), odd.end(), [](const Fast& x, const Fast& y) { return x.band != y.band ? x.band < y.band : (x.row != y.row ? x.row < y.row : x.b < y.b); });
        for (size_t q = 0; q < odd.size();) {
            if (q + 1 < odd.size() && odd[q + 1].band == odd[q].band && odd[q + 1].row - odd[q].row <= 8) {
                fw.push_back({odd[q].band, odd[q].row, odd[q].b, odd[q + 1].b});
                q += 2;
            } else {
                gen.push_back(odd[q].b); ++q;
            }
        }
        std::sort(fw.begin(), fw.end(), [](const FastWarp& x, const FastWarp& y) { return x.band != y.band ? x.band < y.band : (x.row != y.row ? x.row < y.row : x.b0 < y.b0); });
    }
    std::vector<std::vector<Warp>> tiles;
    {
        auto make_tiles = [&](const std::vector<FastWarp>& list) {
            size_t q = 0;
            while (q < list.size()) {
                size_t e = q;
                while (e < list.size() && list[e].band == list[q].band) ++e;
                std::vector<Warp> cur;
                int first_row = 0;
                for (; q < e; ++q) {
                    // (sparse bands -- the last one only holds the rows whose alignment puts a whole window at its very start
                    // -- would otherwise make tiles that span hundreds of rows, i.e. hundreds of atoms i)
                    if (!cur.empty() && list[q].row - first_row > 96) { tiles.push_back(cur); cur.clear(); }
                    if (cur.empty()) first_row = list[q].row;
                    cur.push_back({list[q].b0, list[q].b1, true});
                    if ((int)cur.size() == kApWarps) { tiles.push_back(cur); cur.clear(); }
                }
                if (!cur.empty()) tiles.push_back(cur);
            }
        };
        make_tiles(shared_w);
        ap.n_shared_tiles = (int)tiles.size();
        make_tiles(fw);
        ap.n_fast_windows = 0;
        ap.n_fast_tiles = (int)tiles.size();
        for (const auto& t : tiles) ap.n_fast_windows += 2 * (int)t.size();
        std::sort(gen.begin(), gen.end());
        std::vector<Warp> cur;
        for (size_t g = 0; g < gen.size(); g += 2) {
            cur.push_back({gen[g], g + 1 < gen.size() ? gen[g + 1] : -1, false});
            if ((int)cur.size() == kApWarps) { tiles.push_back(cur); cur.clear(); }
        }
        if (!cur.empty()) tiles.push_back(cur);
    }
    // not worth it when most windows cross rows (rows shorter than ~3 windows), or when there is too little work to tile
    if (10 * (int64_t)ap.n_fast_windows < 6 * (int64_t)nb) return;

    // atoms of every tile: one or two contiguous ranges of the frame
    std::vector<uint8_t> need((size_t)n);
    int max_atoms = 0;
    ap.tiles.resize(tiles.size());
    for (size_t c = 0; c < tiles.size(); ++c) {
        ApTile& t = ap.tiles[c];
        for (int w = 0; w < kApWarps; ++w) t.blk[0][w] = t.blk[1][w] = -1;
        t.fast_mask = 0; t.reserved = 0;
        std::fill(need.begin(), need.end(), 0);
        for (size_t w = 0; w < tiles[c].size(); ++w) {
            const Warp& wp = tiles[c][w];
            t.blk[0][w] = wp.b0; t.blk[1][w] = wp.b1;
            if (wp.fast) t.fast_mask |= 1 << w;
            for (int b : {wp.b0, wp.b1}) {
                if (b < 0) continue;
                int64_t k = (int64_t)b * kApWindow;
                const int64_t k1 = std::min<int64_t>(P, k + kApWindow);
                while (k < k1) {                                   // row by row
                    const int i = row_of(k);
                    const int64_t ke = std::min<int64_t>(k1, ap.row_off[i + 1]);
                    need[i] = 1;
                    const int j0 = i + minsep + (int)(k - ap.row_off[i]);
                    std::fill(need.begin() + j0, need.begin() + j0 + (ke - k), 1);
                    k = ke;
                }
            }
        }
        int lo = 0, hi = n - 1;
        while (!need[lo]) ++lo;
        while (!need[hi]) --hi;
        // the two widest holes inside [lo, hi] (at least kApMinGap atoms wide) split the atoms into up to three ranges
        int g_lo[2] = {-1, -1}, g_len[2] = {0, 0};
        for (int a = lo; a <= hi;) {
            if (need[a]) { ++a; continue; }
            int e = a;
            while (!need[e]) ++e;
            const int len = e - a;
            if (len >= kApMinGap) {
                if (len > g_len[0]) { g_lo[1] = g_lo[0]; g_len[1] = g_len[0]; g_lo[0] = a; g_len[0] = len; }
                else if (len > g_len[1]) { g_lo[1] = a; g_len[1] = len; }
            }
            a = e;
        }
        if (g_len[1] > 0 && g_lo[1] < g_lo[0]) { std::swap(g_lo[0], g_lo[1]); std::swap(g_len[0], g_len[1]); }   // ascending
        if (g_len[0] == 0 && g_len[1] > 0) { g_lo[0] = g_lo[1]; g_len[0] = g_len[1]; g_len[1] = 0; }
        int rb[3], re[3], nr = 1;                                  // ranges [rb, re), multiples of 4
        rb[0] = lo / 4 * 4; re[0] = round_up(hi + 1, 4);
        for (int g = 0; g < 2; ++g) {
            if (g_len[g] == 0) continue;
            const int end_prev = round_up(g_lo[g], 4), beg_next = (g_lo[g] + g_len[g]) / 4 * 4;
            if (beg_next <= end_prev) continue;
            re[nr] = re[nr - 1]; re[nr - 1] = end_prev; rb[nr] = beg_next; ++nr;
        }
        t.a0 = rb[0]; t.n0 = re[0] - rb[0];
        t.a1 = nr > 1 ? rb[1] : 0; t.n1 = nr > 1 ? re[1] - rb[1] : 0;
        t.a2 = nr > 2 ? rb[2] : 0; t.n2 = nr > 2 ? re[2] - rb[2] : 0;
        max_atoms = std::max(max_atoms, t.n0 + t.n1 + t.n2);
    }

    // ring geometry: 8 frames per stage, 3 stages, 2 CTAs per SM
    const int bar_bytes = 2 * kMaxStages * 8 + kMaxStages * 24;
    const int frame_b = 12 * max_atoms;
    int fb = 8, ns = 3;
    if (const char* e = std::getenv("CAVIAR_AP_FB")) fb = std::max(1, std::min(32, atoi(e)));
    if (const char* e = std::getenv("CAVIAR_AP_STAGES")) ns = std::max(2, std::min(kMaxStages, atoi(e)));
    while (fb > 1 && ns * fb * frame_b + bar_bytes > 96 * 1024) fb /= 2;       // (8 or 4 frames per stage: no measurable difference)
    if (ns * fb * frame_b + bar_bytes > 110 * 1024) return;
    ap.frames_per_stage = fb; ap.n_stages = ns; ap.stage_floats = fb * 3 * max_atoms;
    ap.smem_bytes = ns * fb * frame_b + bar_bytes;
    ap.occupancy = 2;
    ap.n_ctas = d.n_ctas > 0 ? d.n_ctas : hp.n_sms * ap.occupancy;
    ap.on = true;
}

inline void fill_info(const HostPlan& hp, CvPlanInfo& info) {
    info.mode = hp.mode;
    info.n_tiles = (int)hp.tiles.size();
    info.n_ctas = hp.n_ctas;
    info.threads_per_cta = kConsumerThreads + ((hp.mode == CV_MODE_DIRECT) ? 0 : kProducerThreads);
    info.stages = hp.n_stages;
    info.frames_per_stage = hp.frames_per_stage;
    info.smem_bytes = hp.smem_bytes;
    info.partial_slots = hp.max_chunks;
    info.n_features = hp.P + hp.A;
    info.n_flag_words = hp.W;
    info.n_distinct_atoms = hp.n_distinct;
    info.staged_floats_per_frame = hp.staged_floats;
    info.boxrec_floats_per_frame = hp.desc.box_kind != CV_BOX_NONE ? kRecFloats : 0;
    info.workspace_bytes = kQueueBytes + (int64_t)hp.max_chunks * (hp.P + hp.A) * 24;
    const int64_t box_b = hp.desc.box_kind == CV_BOX_NONE ? 0 : (hp.desc.box_kind == CV_BOX_ORTHO ? 12 : 36);
    info.gather_bank_conflicts = hp.bank_conflict_refs;
    info.staged_fixedpoint_floats = 0;       // (ABI 2 field) the staged layout is raw float32 throughout
    info.algorithmic_bytes_per_frame = 12 * hp.n_distinct + box_b + 4 * hp.P + 4 * hp.A + 4 * hp.W + 4;
    if (hp.ap.on) {
        info.mode = CV_MODE_ALLPAIRS;
        info.n_tiles = (int)hp.ap.tiles.size();
        info.n_ctas = hp.ap.n_ctas;
        info.threads_per_cta = kApThreads;
        info.stages = hp.ap.n_stages;
        info.frames_per_stage = hp.ap.frames_per_stage;
        info.smem_bytes = hp.ap.smem_bytes;
    }
}

}  // namespace cv
