// OpenCV / CLI11 stand-ins so that the reference's headers compile unmodified for the checker build.
// cv::Mat is a minimal functional matrix (what PreStitch/SectionaryRemap touch at run time);
// cv::remap forwards to the oracle's restatement (oipo_remap_cubic_u16), which tests pin bit-for-bit
// against the real cv2.remap.  Everything else (imread, imwrite, imdecode, phaseCorrelate, ...) only
// has to PARSE: those paths are out of scope and abort if reached.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "oip_oracle.h"

#define CV_16U 2
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16UC4 CV_MAKETYPE(CV_16U, 4)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

namespace cv {
enum { INTER_CUBIC = 2, BORDER_CONSTANT = 0, IMREAD_UNCHANGED = -1 };
struct Scalar { double v[4] = {0, 0, 0, 0}; };
struct Range { int start, end; Range(int s = 0, int e = 0) : start(s), end(e) {} };
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
struct Point2d { double x = 0, y = 0; };
struct NoArray {};
inline NoArray noArray() { return NoArray(); }

class Mat {
public:
    int rows = 0, cols = 0;
    uint8_t *data = nullptr;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void *ext) : rows(r), cols(c), data((uint8_t *)ext), type_(type) { step = (size_t)c * elemSize(); }
    void create(int r, int c, int type) {
        if (r == rows && c == cols && type == type_ && data) return;
        rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
        // zero-filled (the real cv::Mat::create leaves memory uninitialised): rows the reference never writes
        // (SURVEY C-4) then read 0 in the golden files instead of garbage
        store_ = std::shared_ptr<uint8_t>(new uint8_t[(size_t)r * step + 64](), std::default_delete<uint8_t[]>());
        data = store_.get();
    }
    void release() { store_.reset(); data = nullptr; rows = cols = 0; }
    int type() const { return type_; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return (size_t)channels() * ((type_ & 7) == CV_32F ? 4 : 2); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    bool empty() const { return data == nullptr; }
    uint8_t *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T &at(int i) { return reinterpret_cast<T *>(data)[i]; }
    Mat rowRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * elemSize(); m.cols = b - a; return m; }
    Mat operator()(const Range &r, const Range &c) const { return rowRange(r.start, r.end).colRange(c.start, c.end); }
    Mat clone() const { Mat m(rows, cols, type_); for (int r = 0; r < rows; ++r) memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize()); return m; }
    void copyTo(Mat dst) const { for (int r = 0; r < rows; ++r) memcpy(dst.ptr(r), ptr(r), (size_t)cols * elemSize()); }
protected:
    int type_ = 0;
    std::shared_ptr<uint8_t> store_;
};
template <typename T> struct MatType;
template <> struct MatType<uint16_t> { enum { value = CV_16UC1 }; };
template <> struct MatType<float> { enum { value = CV_32FC1 }; };
template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, MatType<T>::value) {}
    Mat_(const Mat &) { abort(); } // type-converting construction (phase-correlation path): out of scope
};
typedef Mat_<uint16_t> Mat1w;
typedef Mat_<float> Mat1f;

inline void remap(const Mat &src, Mat &dst, const Mat &mapx, const Mat &mapy, int interp, int border, const Scalar & = Scalar()) {
    if (src.type() != CV_16UC1 || mapx.type() != CV_32FC1 || mapy.type() != CV_32FC1 || interp != INTER_CUBIC || border != BORDER_CONSTANT) abort();
    dst.create(mapx.rows, mapx.cols, CV_16UC1);
    oipo_remap_cubic_u16((const uint16_t *)src.data, src.cols, src.rows, (int64_t)(src.step / 2), (uint16_t *)dst.data,
                         mapx.cols, mapx.rows, (const float *)mapx.data, (const float *)mapy.data);
}
inline Mat imread(const std::string &, int) { abort(); }
// imwrite stand-in: the matrix bytes as they lie in memory (rows x cols x channels u16), no container.  The real
// cv::imwrite stores 4-channel data as a TIFF in RGBA order (host/tiff_io.hpp, tests/test_tiff_cpu.py)
inline bool imwrite(const std::string &path, const Mat &m) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = true;
    for (int r = 0; r < m.rows && ok; ++r) ok = fwrite(m.ptr(r), m.elemSize(), (size_t)m.cols, f) == (size_t)m.cols;
    fclose(f);
    return ok;
}
inline Mat imdecode(const Mat &, int, Mat * = nullptr) { abort(); }
inline void split(const Mat &, Mat *) { abort(); }
// cv::merge of n single-channel CV_16U planes -> interleaved rows x cols x n (verified against cv2.merge)
inline void merge(const Mat *mv, size_t n, Mat &dst) {
    if (n < 1 || n > 4) abort();
    for (size_t c = 0; c < n; ++c)
        if (mv[c].type() != CV_16UC1 || mv[c].rows != mv[0].rows || mv[c].cols != mv[0].cols) abort();
    dst.create(mv[0].rows, mv[0].cols, CV_MAKETYPE(CV_16U, (int)n));
    for (int r = 0; r < dst.rows; ++r) {
        uint16_t *o = (uint16_t *)dst.ptr(r);
        for (size_t c = 0; c < n; ++c) {
            const uint16_t *s = (const uint16_t *)mv[c].ptr(r);
            for (int x = 0; x < dst.cols; ++x) o[(size_t)x * n + c] = s[x];
        }
    }
}
inline void resize(const Mat &, Mat &, Size, double, double, int) { abort(); }
template <typename A, typename B> inline Point2d phaseCorrelate(const A &, const B &, NoArray, double *) { abort(); }
} // namespace cv

namespace CLI { namespace detail {
inline std::string to_lower(std::string s) { std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); }); return s; }
} }
