// GROMACS XTC trajectories on the host (row f1: the reference CLI advertises nc / dcd / xtc / trr / mdcrd,
// CAVIAR_Distances-Extractor.py:99, and hands them to cpptraj).  XTC stores every frame as XDR (big-endian) header +
// lossily compressed coordinates: positions are rounded to integers (x * precision), and the integers of an atom are
// packed either as one mixed-radix number over the frame's bounding box ("large") or, for runs of atoms close to their
// predecessor (water molecules), as small offsets in an adaptive radix taken from a fixed table.  This file restates the
// published decoding algorithm (the xdrfile library's xdr3dfcoord, LGPL; not vendored, written from the format's
// description) as a bit reader + mixed-radix unpacker, decodes frames in parallel on the host threads and hands back only
// the selected atoms, in Angstrom.  Host code only; compiled by nvcc with the rest of the library.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/caviar_b200.h"

extern "C" int caviar_set_error(int code, const char* msg);

namespace {

constexpr int kFirstIdx = 9;
// radix table of the small-offset coder: magic[i]^3 < 2^i, entries below kFirstIdx unused
const int kMagic[] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406,
                      512, 645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642,
                      26007, 32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127,
                      524287, 660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491,
                      6658042, 8388607, 10568983, 13316085, 16777216};
constexpr int kLastIdx = (int)(sizeof(kMagic) / sizeof(kMagic[0]));

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline float bef(const uint8_t* p) { uint32_t u = be32(p); float f; memcpy(&f, &u, 4); return f; }
inline int bit_length(unsigned long long v) { int n = 0; while (v) { ++n; v >>= 1; } return n; }

struct BitReader {                      // most significant bit first
    const uint8_t* p; const uint8_t* end;
    uint64_t acc = 0; int have = 0; bool overrun = false;
    uint32_t get(int n) {               // n <= 32
        while (have < n) {
            uint8_t b = 0;
            if (p < end) b = *p++; else overrun = true;
            acc = (acc << 8) | b; have += 8;
        }
        have -= n;
        return (uint32_t)((acc >> have) & ((n == 32) ? 0xffffffffull : ((1ull << n) - 1)));
    }
};

// three integers packed as one number in the mixed radix (s0, s1, s2), sent as 8-bit groups, least significant first
inline void unpack3(BitReader& br, int bits, const unsigned s[3], int out[3]) {
    unsigned __int128 v = 0;
    int shift = 0;
    while (bits > 8) { v |= (unsigned __int128)br.get(8) << shift; shift += 8; bits -= 8; }
    if (bits > 0) v |= (unsigned __int128)br.get(bits) << shift;
    out[2] = (int)(v % s[2]); v /= s[2];
    out[1] = (int)(v % s[1]); v /= s[1];
    out[0] = (int)v;
}

struct FrameHead {
    int64_t next = 0;           // offset of the next frame
    int32_t natoms = 0;
    int64_t coord_off = 0;      // offset of the coordinate section (the second natoms field)
    int64_t box_off = 0;
};

// Parses one frame header at `pos`; returns "" or the reason.
std::string frame_head(const uint8_t* d, int64_t size, int64_t pos, FrameHead& h) {
    if (pos + 56 > size) return "truncated frame header";
    const uint32_t magic = be32(d + pos);
    if (magic != 1995 && magic != 2023) return "magic number 1995 / 2023 missing";
    h.natoms = (int32_t)be32(d + pos + 4);
    if (h.natoms <= 0) return "bad atom count";
    h.box_off = pos + 16;
    h.coord_off = pos + 52;
    if ((int32_t)be32(d + h.coord_off) != h.natoms) return "atom count mismatch inside a frame";
    int64_t p = h.coord_off + 4;
    if (h.natoms <= 9) {
        h.next = p + 12ll * h.natoms;
    } else {
        p += 4 + 24 + 4;        // precision, minint[3], maxint[3], smallidx
        if (p + 8 > size) return "truncated frame";
        int64_t nbytes;
        if (magic == 2023) { nbytes = (int64_t)(((uint64_t)be32(d + p) << 32) | be32(d + p + 4)); p += 8; }
        else { nbytes = (int64_t)be32(d + p); p += 4; }
        if (nbytes < 0) return "bad byte count";
        h.next = p + (nbytes + 3) / 4 * 4;
    }
    if (h.next > size) return "truncated frame";
    return "";
}

// Decodes the frame at `pos`: `sel` atoms (nullptr: all), scaled, into out[n_sel][3]; box (9 floats, scaled) into box_out.
std::string decode_frame(const uint8_t* d, int64_t size, int64_t pos, const int32_t* sel, int64_t n_sel, float scale,
                         float* out, float* box_out, std::vector<float>& tmp) {
    FrameHead h;
    std::string why = frame_head(d, size, pos, h);
    if (!why.empty()) return why;
    const int n = h.natoms;
    if (box_out) for (int k = 0; k < 9; ++k) box_out[k] = bef(d + h.box_off + 4 * k) * scale;
    tmp.resize((size_t)3 * n);
    float* xyz = tmp.data();
    int64_t p = h.coord_off + 4;
    if (n <= 9) {
        for (int k = 0; k < 3 * n; ++k) xyz[k] = bef(d + p + 4 * k);
    } else {
        const float precision = bef(d + p); p += 4;
        int minint[3], maxint[3];
        for (int k = 0; k < 3; ++k) minint[k] = (int32_t)be32(d + p + 4 * k);
        for (int k = 0; k < 3; ++k) maxint[k] = (int32_t)be32(d + p + 12 + 4 * k);
        p += 24;
        int smallidx = (int32_t)be32(d + p); p += 4;
        const uint32_t magic = be32(d + pos);
        int64_t nbytes;
        if (magic == 2023) { nbytes = (int64_t)(((uint64_t)be32(d + p) << 32) | be32(d + p + 4)); p += 8; }
        else { nbytes = (int64_t)be32(d + p); p += 4; }
        if (!(precision > 0.f)) return "non-positive precision";
        if (smallidx < kFirstIdx || smallidx >= kLastIdx) return "small-offset radix index out of range";
        unsigned sizeint[3];
        for (int k = 0; k < 3; ++k) {
            const long long s = (long long)maxint[k] - (long long)minint[k] + 1;
            if (s <= 0 || s > 0xffffffffll) return "bad coordinate range";
            sizeint[k] = (unsigned)s;
        }
        int bitsize, bitsint[3] = {0, 0, 0};
        if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) {
            for (int k = 0; k < 3; ++k) bitsint[k] = bit_length(sizeint[k]);
            bitsize = 0;                                        // ranges too wide for one mixed-radix number
        } else {
            // bits of the product, taken 8 at a time like the writer's multi-byte multiplication does
            unsigned __int128 prod = (unsigned __int128)sizeint[0] * sizeint[1] * sizeint[2];
            int nb = 0;
            while ((prod >> (8 * (nb + 1))) != 0) ++nb;
            bitsize = 8 * nb + bit_length((unsigned long long)(prod >> (8 * nb)));
        }
        int smaller = kMagic[std::max(kFirstIdx, smallidx - 1)] / 2;
        int smallnum = kMagic[smallidx] / 2;
        unsigned sizesmall[3] = {(unsigned)kMagic[smallidx], (unsigned)kMagic[smallidx], (unsigned)kMagic[smallidx]};
        BitReader br{d + p, d + p + nbytes};
        const float inv = 1.0f / precision;
        int i = 0, run = 0;
        float* w = xyz;
        while (i < n) {
            int cur[3], prev[3];
            if (bitsize == 0) { for (int k = 0; k < 3; ++k) cur[k] = (int)br.get(bitsint[k]); }
            else unpack3(br, bitsize, sizeint, cur);
            ++i;
            for (int k = 0; k < 3; ++k) { cur[k] += minint[k]; prev[k] = cur[k]; }
            int is_smaller = 0;
            if (br.get(1) == 1) {
                run = (int)br.get(5);
                is_smaller = run % 3;
                run -= is_smaller;
                --is_smaller;
            }
            if (run > 0) {
                if (i + run / 3 > n) return "a run of small offsets leaves the frame";
                for (int k = 0; k < run; k += 3) {
                    int s[3];
                    unpack3(br, smallidx, sizesmall, s);
                    ++i;
                    for (int c = 0; c < 3; ++c) s[c] += prev[c] - smallnum;
                    if (k == 0) {
                        // the writer swapped the first two atoms of a run (oxygen between its hydrogens): undo it
                        for (int c = 0; c < 3; ++c) std::swap(s[c], prev[c]);
                        for (int c = 0; c < 3; ++c) *w++ = prev[c] * inv;
                    } else {
                        for (int c = 0; c < 3; ++c) prev[c] = s[c];
                    }
                    for (int c = 0; c < 3; ++c) *w++ = s[c] * inv;
                }
            } else {
                for (int c = 0; c < 3; ++c) *w++ = cur[c] * inv;
            }
            smallidx += is_smaller;
            if (smallidx < kFirstIdx || smallidx >= kLastIdx) return "small-offset radix index out of range";
            if (is_smaller < 0) {
                smallnum = smaller;
                smaller = smallidx > kFirstIdx ? kMagic[smallidx - 1] / 2 : 0;
            } else if (is_smaller > 0) {
                smaller = smallnum;
                smallnum = kMagic[smallidx] / 2;
            }
            sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)kMagic[smallidx];
            if (br.overrun) return "compressed coordinates end before the last atom";
        }
    }
    if (sel) {
        for (int64_t a = 0; a < n_sel; ++a) {
            const int32_t s = sel[a];
            if (s < 0 || s >= n) return "selected atom outside the frame";
            out[3 * a] = xyz[3 * s] * scale; out[3 * a + 1] = xyz[3 * s + 1] * scale; out[3 * a + 2] = xyz[3 * s + 2] * scale;
        }
    } else {
        for (int k = 0; k < 3 * n; ++k) out[k] = xyz[k] * scale;
    }
    return "";
}

}  // namespace

// Walks the frames of an XTC file image: returns the number of frames (or < 0), the atom count, and -- if `offsets` has
// room (cap entries) -- the byte offset of every frame.  Frames of different atom counts are refused.
extern "C" int64_t caviar_xtc_index(const void* data, int64_t size, int64_t* offsets, int64_t cap, int32_t* n_atoms) {
    if (!data || size < 0) return caviar_set_error(CV_E_INVALID, "null XTC image");
    const uint8_t* d = static_cast<const uint8_t*>(data);
    int64_t pos = 0, n = 0;
    int32_t natoms = 0;
    while (pos < size) {
        FrameHead h;
        std::string why = frame_head(d, size, pos, h);
        if (!why.empty()) return caviar_set_error(CV_E_INVALID, ("XTC frame " + std::to_string(n) + ": " + why).c_str());
        if (n == 0) natoms = h.natoms;
        else if (h.natoms != natoms) return caviar_set_error(CV_E_UNSUPPORTED, "XTC frames with different atom counts");
        if (offsets && n < cap) offsets[n] = pos;
        ++n;
        pos = h.next;
    }
    if (n_atoms) *n_atoms = natoms;
    return n;
}

// Decodes the frames at `offsets[0 .. n_frames)` (from caviar_xtc_index) on the host threads:
//   xyz_out [n_frames][n_sel][3] float32 = coordinates of the selected atoms (atoms == NULL: all n_atoms) * scale
//   box_out [n_frames][9] float32 = box vectors * scale, or NULL            (scale = 10: nm -> Angstrom)
extern "C" int caviar_xtc_decode_host(const void* data, int64_t size, const int64_t* offsets, int64_t n_frames,
                                      const int32_t* atoms, int64_t n_sel, float scale, float* xyz_out, float* box_out) {
    if (n_frames <= 0) return CV_OK;
    if (!data || !offsets || !xyz_out) return caviar_set_error(CV_E_INVALID, "null argument");
    const uint8_t* d = static_cast<const uint8_t*>(data);
    FrameHead h0;
    std::string why = frame_head(d, size, offsets[0], h0);
    if (!why.empty()) return caviar_set_error(CV_E_INVALID, ("XTC: " + why).c_str());
    const int64_t row = atoms ? n_sel : h0.natoms;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int n_thr = (int)std::min<int64_t>(std::min<unsigned>(hw, 32u), n_frames);
    std::vector<std::string> errs((size_t)n_thr);
    auto work = [&](int t) {
        std::vector<float> tmp;
        for (int64_t f = t; f < n_frames; f += n_thr) {
            if (offsets[f] < 0 || offsets[f] >= size) { errs[t] = "frame offset outside the file"; return; }
            std::string e = decode_frame(d, size, offsets[f], atoms, n_sel, scale, xyz_out + f * row * 3,
                                         box_out ? box_out + f * 9 : nullptr, tmp);
            if (!e.empty()) { errs[t] = "frame at byte " + std::to_string(offsets[f]) + ": " + e; return; }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_thr; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (const std::string& e : errs) if (!e.empty()) return caviar_set_error(CV_E_INVALID, ("XTC " + e).c_str());
    return CV_OK;
}
