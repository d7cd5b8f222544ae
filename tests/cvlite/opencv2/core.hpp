// cvlite — TEST-ONLY stand-in for <opencv2/core.hpp>.
//
// This container (and the GPU boxes) have no OpenCV C++ headers. The reference's hot-path
// interface is expressed in cv::Point2f / cv::Mat (include/KDTree.h:15,25,30 and
// include/RansacFilter.h:19-21 of the reference), so two things need *some* definition of
// those types to compile here:
//   1. the reference's own sources, built unmodified into oracle/_ref/ as the checker
//      (src/KDTree.cpp, src/RansacFilter.cpp, tests/test_kdtree.cpp);
//   2. this repo's drop-in adapter headers (include/KDTree.h, include/RansacFilter.h) when
//      their tests are compiled without a real OpenCV.
// A maintainer integrating the library uses real OpenCV; this file is never shipped in the
// product library.
//
// Arithmetic contract of the cv::Mat subset (only what src/RansacFilter.cpp:69-140 calls).
// Each rule was checked bit-for-bit against Python cv2 4.13.0 (tests/golden/gen_golden.py):
//   Mat * Mat   (gemm, no flags, inner dim 3)   fp32, ((a0*b0 + a1*b1) + a2*b2), every op rounded
//   Mat.t() * Mat (GEMM_1_T)                    products and sums in fp64, result rounded to fp32
//   mul / operator/ / operator+                 element-wise fp32, IEEE (x/0 = inf, 0/0 = NaN)
//   reduce(.., 0, REDUCE_SUM)                   fp32, (r0 + r1) + r2
//   sum()                                       fp64 accumulation; order is SIMD-dependent in OpenCV,
//                                               so the order used here is the oracle's definition
//                                               (oracle/vb_oracle.h: vbo_score_sum)
//   SVDecomp                                    OpenCV's result depends on its LAPACK/Jacobi build
//                                               ("unpinned"); here it is the oracle's defined solve
//                                               (vbo_null_vector_8x9 / vbo_svd3x3)
#ifndef CVLITE_OPENCV2_CORE_HPP
#define CVLITE_OPENCV2_CORE_HPP

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <random>
#include <utility>
#include <vector>

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U> &o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)) {}
    // OpenCV: saturate_cast<T>(x*pt.x + y*pt.y); for float that is plain float arithmetic.
    T dot(const Point_ &o) const { return static_cast<T>(x * o.x + y * o.y); }
};
template <typename T> inline Point_<T> operator-(const Point_<T> &a, const Point_<T> &b) {
    return Point_<T>(static_cast<T>(a.x - b.x), static_cast<T>(a.y - b.y));
}
template <typename T> inline Point_<T> operator+(const Point_<T> &a, const Point_<T> &b) {
    return Point_<T>(static_cast<T>(a.x + b.x), static_cast<T>(a.y + b.y));
}
template <typename T> inline bool operator==(const Point_<T> &a, const Point_<T> &b) { return a.x == b.x && a.y == b.y; }
template <typename T> inline bool operator!=(const Point_<T> &a, const Point_<T> &b) { return !(a == b); }
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;

}  // namespace cv

#if defined(CVLITE_WITH_ORACLE_HOOKS) && !defined(CVLITE_WITH_MAT)
#define CVLITE_WITH_MAT
#endif

#ifdef CVLITE_WITH_MAT
#ifdef CVLITE_WITH_ORACLE_HOOKS
// ---- oracle hooks (C, oracle/vb_oracle.c) used by SVDecomp / sum / the seed stand-in -------------
extern "C" {
void vbo_null_vector_8x9(const float *A /*8x9 row-major*/, float *f9);
void vbo_svd3x3(const float *F /*3x3*/, float *U /*3x3*/, float *D /*3*/, float *Vt /*3x3*/);
double vbo_score_sum(const float *e, int n);
unsigned vbo_ref_seed_next(void);
}

// Seed hook for src/RansacFilter.cpp:15-16 ("std::random_device rd; std::mt19937 gen(rd());").
// The reference seeds from the OS; a parity test needs a chosen seed, so inside the reference's
// translation unit the name random_device resolves to this fixed-value device.
namespace std {
struct cvlite_seeded_random_device {
    typedef unsigned int result_type;
    result_type operator()() { return vbo_ref_seed_next(); }
};
}  // namespace std
#define random_device cvlite_seeded_random_device
#endif  // CVLITE_WITH_ORACLE_HOOKS

#define CV_32FC1 5
#define CV_32F 5

namespace cv {

enum ReduceTypes { REDUCE_SUM = 0 };

struct Scalar_ {
    double v[4];
    double operator[](int i) const { return v[i]; }
};

class Mat;
struct MatT {  // result of Mat::t(): remembers that the left gemm operand is transposed
    const Mat &m;
    explicit MatT(const Mat &m_) : m(m_) {}
};

class Mat {
   public:
    int rows, cols;
    Mat() : rows(0), cols(0), step_(0), off_(0) {}
    Mat(int r, int c, int /*type*/) : rows(r), cols(c), step_(c), off_(0), buf_(new std::vector<float>(size_t(r) * c)) {}
    bool empty() const { return rows == 0 || cols == 0 || !buf_; }
    template <typename T> T &at(int i, int j) { return (*buf_)[off_ + size_t(i) * step_ + j]; }
    template <typename T> const T &at(int i, int j) const { return (*buf_)[off_ + size_t(i) * step_ + j]; }
    template <typename T> T &at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i / cols, i % cols); }
    template <typename T> const T &at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i / cols, i % cols); }
    Mat row(int i) const {  // shares storage, like cv::Mat::row
        Mat r;
        r.rows = 1; r.cols = cols; r.step_ = step_; r.off_ = off_ + size_t(i) * step_; r.buf_ = buf_;
        return r;
    }
    Mat reshape(int /*cn*/, int new_rows) const {  // copies; the reference only reads the result
        Mat r(new_rows, rows * cols / new_rows, CV_32F);
        for (int i = 0; i < rows * cols; i++) r.at<float>(i / r.cols, i % r.cols) = at<float>(i / cols, i % cols);
        return r;
    }
    Mat clone() const {
        Mat r(rows, cols, CV_32F);
        for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) r.at<float>(i, j) = at<float>(i, j);
        return r;
    }
    void copyTo(Mat &dst) const { dst = clone(); }
    MatT t() const { return MatT(*this); }
    Mat mul(const Mat &o) const {
        Mat r(rows, cols, CV_32F);
        for (int i = 0; i < rows; i++) for (int j = 0; j < cols; j++) r.at<float>(i, j) = at<float>(i, j) * o.at<float>(i, j);
        return r;
    }
    static Mat diag(const Mat &d) {
        int n = d.rows * d.cols;
        Mat r(n, n, CV_32F);
        for (int i = 0; i < n; i++) r.at<float>(i, i) = d.at<float>(i);
        return r;
    }

   private:
    size_t step_, off_;
    std::shared_ptr<std::vector<float> > buf_;
};

// gemm, no flags: fp32 ((a0*b0 + a1*b1) + a2*b2) — only inner dimension 3 is exercised by the reference.
inline Mat operator*(const Mat &a, const Mat &b) {
    if (a.cols != 3 || b.rows != 3) { std::fprintf(stderr, "cvlite: gemm inner dim %d unsupported\n", a.cols); std::abort(); }
    Mat r(a.rows, b.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            float p0 = a.at<float>(i, 0) * b.at<float>(0, j);
            float p1 = a.at<float>(i, 1) * b.at<float>(1, j);
            float p2 = a.at<float>(i, 2) * b.at<float>(2, j);
            float s = p0 + p1;
            r.at<float>(i, j) = s + p2;
        }
    return r;
}
// gemm with GEMM_1_T: fp64 products and sums, rounded to fp32 at the end.
inline Mat operator*(const MatT &at, const Mat &b) {
    const Mat &a = at.m;
    if (a.rows != 3 || b.rows != 3) { std::fprintf(stderr, "cvlite: gemm(T) inner dim unsupported\n"); std::abort(); }
    Mat r(a.cols, b.cols, CV_32F);
    for (int i = 0; i < a.cols; i++)
        for (int j = 0; j < b.cols; j++) {
            double s = double(a.at<float>(0, i)) * double(b.at<float>(0, j));
            s = s + double(a.at<float>(1, i)) * double(b.at<float>(1, j));
            s = s + double(a.at<float>(2, i)) * double(b.at<float>(2, j));
            r.at<float>(i, j) = float(s);
        }
    return r;
}
inline Mat operator/(const Mat &a, const Mat &b) {
    Mat r(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) r.at<float>(i, j) = a.at<float>(i, j) / b.at<float>(i, j);
    return r;
}
inline Mat operator+(const Mat &a, const Mat &b) {
    Mat r(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++) for (int j = 0; j < a.cols; j++) r.at<float>(i, j) = a.at<float>(i, j) + b.at<float>(i, j);
    return r;
}
inline void reduce(const Mat &src, Mat &dst, int /*dim = 0*/, int /*REDUCE_SUM*/) {
    Mat r(1, src.cols, CV_32F);
    for (int j = 0; j < src.cols; j++) {
        float s = src.at<float>(0, j);
        for (int i = 1; i < src.rows; i++) s = s + src.at<float>(i, j);
        r.at<float>(0, j) = s;
    }
    dst = r;
}
#ifdef CVLITE_WITH_ORACLE_HOOKS
inline Scalar_ sum(const Mat &m) {
    Mat c = (m.rows == 1) ? m.clone() : m.reshape(0, 1);
    Scalar_ s;
    s.v[0] = vbo_score_sum(&c.at<float>(0, 0), c.cols);
    s.v[1] = s.v[2] = s.v[3] = 0;
    return s;
}

struct SVD {
    enum Flags { MODIFY_A = 1, NO_UV = 2, FULL_UV = 4 };
};
// Only the two shapes src/RansacFilter.cpp:94,98 uses. For the 8x9 system only Vt.row(8) — the
// null vector — is defined; the other rows, U and D are zero.
inline void SVDecomp(const Mat &A, Mat &D, Mat &U, Mat &Vt, int /*flags*/) {
    if (A.rows == 8 && A.cols == 9) {
        Mat a = A.clone();
        Mat vt(9, 9, CV_32F);
        vbo_null_vector_8x9(&a.at<float>(0, 0), &vt.at<float>(8, 0));
        D = Mat(8, 1, CV_32F); U = Mat(8, 8, CV_32F); Vt = vt;
    } else if (A.rows == 3 && A.cols == 3) {
        Mat a = A.clone();
        Mat u(3, 3, CV_32F), d(3, 1, CV_32F), vt(3, 3, CV_32F);
        vbo_svd3x3(&a.at<float>(0, 0), &u.at<float>(0, 0), &d.at<float>(0, 0), &vt.at<float>(0, 0));
        D = d; U = u; Vt = vt;
    } else {
        std::fprintf(stderr, "cvlite: SVDecomp %dx%d unsupported\n", A.rows, A.cols);
        std::abort();
    }
}
#endif  // CVLITE_WITH_ORACLE_HOOKS

}  // namespace cv
#endif  // CVLITE_WITH_MAT

#endif  // CVLITE_OPENCV2_CORE_HPP
