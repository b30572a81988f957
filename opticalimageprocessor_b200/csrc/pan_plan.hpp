// pan_plan.hpp -- host-side geometry of the sectioned sub-pixel shift (pure C++, no CUDA).
//
// The reference shifts a strip in 30000-row sections through ONE reused buffer
// (ref stitcher.h:83-139, imageop.h:230-275).  Everything that makes that observable in the
// output file -- which section an output row comes from, which source rows are zero border,
// which are stale rows left over from the previous section -- is a closed-form function of the
// global row index.  This header evaluates it; the kernel only sees uniform "segments".
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace oip {

// one run of output rows that share a section buffer
struct ShiftSegment {
    int64_t g0, g1;      // output (file) rows [g0,g1)
    int64_t j0;          // section-local row index of g0
    int64_t sec_off;     // global source row held in buffer row 0
    int32_t rows_s;      // buffer rows [0,rows_s) hold fresh rows of this section
    int64_t stale_off;   // buffer rows [rows_s,hbuf) hold global row stale_off+t (or -1: none)
    int32_t hbuf;        // buffer height seen by cv::remap (border / interior decisions)
};

inline int shift_ucut(double dY) { return dY >= 0.0 ? 0 : (int)(-dY) + 1; } // ref stitcher.h:122
inline int shift_bcut(double dY) { return dY >= 0.0 ? (int)dY + 1 : 0; }    // ref stitcher.h:123

// returns false if the arguments are rejected (the reference would assert / misbehave)
inline bool plan_shift_segments(int64_t total_rows, int section_rows, int row_guard, double dY,
                                std::vector<ShiftSegment> &out)
{
    out.clear();
    if (total_rows <= 0) return true;
    if (total_rows <= row_guard) {
        // ref imageop.h:242-244 throws "please use cv::remap()": we do exactly that, one remap
        out.push_back({0, total_rows, 0, 0, (int32_t)total_rows, -1, (int32_t)total_rows});
        return true;
    }
    if (section_rows > row_guard || section_rows < 8) return false;
    const int u = shift_ucut(dY), b = shift_bcut(dY), c = u + b;      // imageop.h:246
    if (c >= section_rows) return false;
    struct Sec { int64_t off; int32_t rows; };
    std::vector<Sec> secs;
    int64_t row_offset = 0;
    for (;;) {                                                         // imageop.h:249-267
        int64_t left = total_rows - row_offset;
        int32_t rows = (int32_t)std::min<int64_t>(section_rows, left);
        if (rows <= c) break;
        secs.push_back({row_offset, rows});
        row_offset += rows - c;
    }
    const int L = (int)secs.size() - 1;
    for (int s = 0; s <= L; ++s) {
        ShiftSegment g;
        g.sec_off = secs[s].off;
        g.rows_s = secs[s].rows;
        g.hbuf = section_rows;                                         // dst/buff are always S rows tall
        g.stale_off = (secs[s].rows < section_rows && s > 0) ? secs[s - 1].off : -1;
        g.g0 = secs[s].off + (s == 0 ? 0 : u);                         // :260-265 (upper cut only once)
        g.g1 = secs[s].off + secs[s].rows - b;
        g.j0 = g.g0 - secs[s].off;
        if (g.g1 > g.g0) out.push_back(g);
    }
    if (b > 0 && L >= 0) {                                             // imageop.h:269-272
        ShiftSegment g;
        g.sec_off = secs[L].off;
        g.rows_s = secs[L].rows;
        g.hbuf = section_rows;
        g.stale_off = (secs[L].rows < section_rows && L > 0) ? secs[L - 1].off : -1;
        g.g0 = total_rows - b;
        g.g1 = total_rows;
        g.j0 = section_rows - b;
        out.push_back(g);
    }
    return true;
}

// OpenCV's fixed-point map conversion for one coordinate: cvRound(float(i + d) * 32)
// (ref stitcher.h:96-97 builds float(i + d); cv::remap quantises to 1/32 px)
inline int map_fixed(int64_t i, double d)
{
    float m = (float)((double)i + d);
    float v = m * 32.0f;
    return (int)lrintf(v);
}
inline int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }
// first tap row/col (XY - 1) for integer coordinate i
inline int tap_base(int64_t i, double d) { return sat_short(map_fixed(i, d) >> 5) - 1; }

} // namespace oip
