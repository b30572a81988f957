// NumCpp stand-in (dpilger26/NumCpp, absent here, unpinned in the reference) so that the reference's preproc.h compiles
// UNMODIFIED for the checker build.  Only what preproc.h names: nc::NdArray views with 2-D slicing and astype, and
// nc::polynomial::Poly1d<double>::fit.  The estimation path that uses them (CalcInterBandCorrelation,
// DoCorrelationPolynomialFitting: SURVEY 8f N2, floating point, pinned by tolerance against cv2 / numpy.polyfit) is not
// run through this shim: fit() aborts.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdlib>
#include <vector>

namespace nc {
struct Slice { int start, stop; Slice(int a, int b) : start(a), stop(b) {} };
template <typename T> class NdArray {
public:
    NdArray() {}
    NdArray(T *p, int rows, int cols, bool /*takeOwnership*/) : rows_(rows), cols_(cols) { v_.assign(p, p + (size_t)rows * cols); }
    NdArray operator()(const Slice &r, const Slice &c) const {
        NdArray o; o.rows_ = r.stop - r.start; o.cols_ = c.stop - c.start; o.v_.resize((size_t)o.rows_ * o.cols_);
        for (int y = 0; y < o.rows_; ++y)
            for (int x = 0; x < o.cols_; ++x) o.v_[(size_t)y * o.cols_ + x] = v_[(size_t)(y + r.start) * cols_ + x + c.start];
        return o;
    }
    template <typename U> NdArray<U> astype() const {
        NdArray<U> o; o.resize(rows_, cols_);
        for (size_t i = 0; i < v_.size(); ++i) o.data()[i] = (U)v_[i];
        return o;
    }
    void resize(int r, int c) { rows_ = r; cols_ = c; v_.resize((size_t)r * c); }
    T *data() { return v_.data(); }
private:
    int rows_ = 0, cols_ = 0;
    std::vector<T> v_;
};
namespace polynomial {
template <typename T> class Poly1d {
public:
    static Poly1d fit(const NdArray<T> &, const NdArray<T> &, int) { abort(); }
    std::vector<T> coefficients() const { return c_; }
private:
    std::vector<T> c_;
};
} // namespace polynomial
} // namespace nc
