// Amber ASCII trajectories (mdcrd), host side: the reference CLI advertises them (CAVIAR_Distances-Extractor.py:99) and
// leaves the parsing to cpptraj.  A frame is 3N numbers in Fortran 10F8.3 -- eight characters per value, ten values per
// line, no separator guaranteed ("-123.456-234.567") -- so the position of every value is known without scanning: only
// the SELECTED atoms are parsed, frames in parallel on the host threads.
//
// A field is read as [spaces][sign][digits][.digits][spaces]: mantissa / 10^decimals in double.  With the mantissa below
// 2^53 and a power of ten that is exact in double the quotient is the correctly rounded double of the decimal string --
// the value strtod (and numpy's string -> float64 conversion, which trajio used before) returns -- and it is rounded to
// float32 once, as an assignment into a float32 array does.  Anything else in a field (the "********" of a Fortran
// overflow, an empty field) is an error naming the frame, as the numpy conversion raised ValueError; "NaN" / "Infinity"
// and exponents go through strtod.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "../../include/caviar_b200.h"
#include "cv_common.h"

extern "C" int caviar_set_error(int code, const char* msg);     // cv_api.cu: sets the thread-local message, returns code

namespace {

inline bool parse_f8(const uint8_t* p, float& out) {
    int i = 0;
    while (i < 8 && p[i] == ' ') ++i;
    bool neg = false;
    if (i < 8 && (p[i] == '-' || p[i] == '+')) { neg = p[i] == '-'; ++i; }
    uint64_t mant = 0;
    int digits = 0, dec = 0;
    while (i < 8 && p[i] >= '0' && p[i] <= '9') { mant = mant * 10 + (uint64_t)(p[i] - '0'); ++i; ++digits; }
    if (i < 8 && p[i] == '.') {
        ++i;
        while (i < 8 && p[i] >= '0' && p[i] <= '9') { mant = mant * 10 + (uint64_t)(p[i] - '0'); ++i; ++digits; ++dec; }
    }
    while (i < 8 && p[i] == ' ') ++i;
    if (i != 8 || digits == 0) {
        // not plain fixed point ("     NaN", "Infinity", an exponent): let strtod decide, as the numpy conversion did
        char buf[9];
        int b = 0, e = 8;
        while (b < 8 && p[b] == ' ') ++b;
        while (e > b && p[e - 1] == ' ') --e;
        if (e == b) return false;
        for (int k = b; k < e; ++k) buf[k - b] = (char)p[k];
        buf[e - b] = 0;
        char* end = nullptr;
        const double v = std::strtod(buf, &end);
        if (end != buf + (e - b)) return false;
        out = (float)v;
        return true;
    }
    static const double pow10[8] = {1.0, 10.0, 100.0, 1000.0, 1e4, 1e5, 1e6, 1e7};
    const double v = (double)mant / pow10[dec];          // at most 8 digits: exact operands, one correctly rounded division
    out = (float)(neg ? -v : v);
    return true;
}

}  // namespace

// Frames `frames[0 .. n_frames)` (indices) of an mdcrd image whose first frame starts at byte `first`, frames
// `frame_bytes` apart, lines ending in `eol` bytes (1 = LF, 2 = CR LF):
//   xyz_out [n_frames][n_sel][3] float32 = coordinates of the selected atoms (atoms == NULL: all n_atoms), in the file's unit
extern "C" int caviar_mdcrd_decode_host(const void* data, int64_t size, int64_t first, int64_t frame_bytes, int32_t eol,
                                        int32_t n_atoms, const int64_t* frames, int64_t n_frames, const int32_t* atoms,
                                        int64_t n_sel, float* xyz_out) {
    if (n_frames <= 0) return CV_OK;
    if (!data || !frames || !xyz_out) return caviar_set_error(CV_E_INVALID, "null argument");
    if (n_atoms <= 0 || first < 0 || frame_bytes <= 0 || (eol != 1 && eol != 2))
        return caviar_set_error(CV_E_INVALID, "bad mdcrd geometry");
    const int64_t n_val = 3ll * n_atoms;
    const int64_t coord_bytes = 8 * n_val + (int64_t)eol * ((n_val + 9) / 10);
    if (coord_bytes > frame_bytes) return caviar_set_error(CV_E_INVALID, "mdcrd frame shorter than its coordinates");
    const int64_t row = atoms ? n_sel : (int64_t)n_atoms;
    if (atoms)
        for (int64_t a = 0; a < n_sel; ++a)
            if (atoms[a] < 0 || atoms[a] >= n_atoms) return caviar_set_error(CV_E_INVALID, "selected atom outside the frame");
    const uint8_t* d = static_cast<const uint8_t*>(data);
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int n_thr = (int)std::min<int64_t>(std::min<unsigned>(hw, 32u), n_frames);
    std::vector<std::string> errs((size_t)n_thr);
    auto work = [&](int t) {
        for (int64_t f = t; f < n_frames; f += n_thr) {
            const int64_t base = first + frames[f] * frame_bytes;
            if (frames[f] < 0 || base + coord_bytes > size) { errs[t] = "frame " + std::to_string(frames[f]) + " outside the file"; return; }
            float* out = xyz_out + f * row * 3;
            for (int64_t a = 0; a < row; ++a) {
                const int64_t v0 = 3ll * (atoms ? atoms[a] : (int32_t)a);
                for (int c = 0; c < 3; ++c) {
                    const int64_t v = v0 + c;
                    if (!parse_f8(d + base + 8 * v + (v / 10) * eol, out[3 * a + c])) {
                        errs[t] = "frame " + std::to_string(frames[f]) + ", value " + std::to_string(v) + ": not a number in F8.3 ('" +
                                  std::string(reinterpret_cast<const char*>(d + base + 8 * v + (v / 10) * eol), 8) + "')";
                        return;
                    }
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_thr; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (const std::string& e : errs) if (!e.empty()) return caviar_set_error(CV_E_INVALID, ("mdcrd " + e).c_str());
    return CV_OK;
}
