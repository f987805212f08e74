/* Minimal stand-in for the OpenCV API the reference's hot-path sources use (TEST INFRASTRUCTURE; oracle/_ref/libref_path.so).
 * OpenCV is not in this image and not vendored by the reference, so the reference's own sources are compiled UNMODIFIED against
 * this header instead.  What is modelled is exactly the arithmetic the path reaches through cv::Mat — and it is the arithmetic
 * the committed cv2 4.13 goldens pinned (tests/golden/make_golden_cv2.py, tests/test_oracle_golden.py):
 *   Mat * Mat      CV_64F: products accumulated sequentially in double.  CV_32F: OpenCV's hand-unrolled small-matrix branch
 *                  (no transposed operand, 2 <= len <= 4, len == rows or cols of the result): float products, float left-to-right
 *                  sum; every other CV_32F product: (double)a*(double)b accumulated sequentially in double, narrowed to float.
 *   .t()           lazy, as OpenCV's MatExpr folds it into the gemm flags (a transposed operand selects the general branch).
 *   determinant / .inv() of a 3x3: cofactor formula in double, inverse = cofactors * (1/det), narrowed for CV_32F.
 *   cv::triangulatePoints, cv::computeCorrespondEpilines: the bit-exact restatements of the oracle (eg3d_oracle_match.cpp /
 *                  eg3d_oracle_geom.cpp), themselves checked against real cv2 calls.
 * Everything else (image I/O, drawing, findFundamentalMat) is declared so that the sources compile and aborts if ever called. */
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <sstream>
#include <fstream>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_LOAD_IMAGE_COLOR 1
#define CV_LOAD_IMAGE_GRAYSCALE 0
#define CV_BGR2GRAY 6
#define CV_GRAY2BGR 8
#define CV_AA 16
#define CV_FILLED -1

namespace cv {
typedef unsigned char uchar;
typedef std::string String;
[[noreturn]] inline void eg3d_stub_unreachable(const char* what) { fprintf(stderr, "eg3d cv stub: %s is not modelled (off the hot path)\n", what); abort(); }

template <typename T, int N> struct Vec {
  T val[N];
  Vec() { for (int i = 0; i < N; i++) val[i] = T(0); }
  Vec(T a, T b) { static_assert(N >= 2, ""); for (int i = 0; i < N; i++) val[i] = T(0); val[0] = a; val[1] = b; }
  Vec(T a, T b, T c) { static_assert(N >= 3, ""); for (int i = 0; i < N; i++) val[i] = T(0); val[0] = a; val[1] = b; val[2] = c; }
  Vec(T a, T b, T c, T d) { static_assert(N >= 4, ""); for (int i = 0; i < N; i++) val[i] = T(0); val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
  T dot(const Vec& o) const { T s = T(0); for (int i = 0; i < N; i++) s += val[i] * o.val[i]; return s; }
  bool operator==(const Vec& o) const { for (int i = 0; i < N; i++) if (val[i] != o.val[i]) return false; return true; }
  bool operator!=(const Vec& o) const { return !(*this == o); }
};
typedef Vec<uchar, 3> Vec3b; typedef Vec<float, 2> Vec2f; typedef Vec<float, 3> Vec3f; typedef Vec<float, 4> Vec4f;
typedef Vec<double, 2> Vec2d; typedef Vec<double, 3> Vec3d; typedef Vec<double, 4> Vec4d; typedef Vec<int, 2> Vec2i;
template <typename T, int N> Vec<T, N> operator*(const Vec<T, N>& a, double s) { Vec<T, N> r; for (int i = 0; i < N; i++) r[i] = T(a[i] * s); return r; }
template <typename T, int N> Vec<T, N> operator*(double s, const Vec<T, N>& a) { return a * s; }
template <typename T, int N> Vec<T, N> operator+(const Vec<T, N>& a, const Vec<T, N>& b) { Vec<T, N> r; for (int i = 0; i < N; i++) r[i] = T(a[i] + b[i]); return r; }

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T a, T b) : x(a), y(b) {}
  template <typename U> Point_(const Point_<U>& o) : x(T(o.x)), y(T(o.y)) {}
};
typedef Point_<int> Point; typedef Point_<int> Point2i; typedef Point_<float> Point2f; typedef Point_<double> Point2d;
template <typename T> struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
};
typedef Point3_<float> Point3f; typedef Point3_<double> Point3d; typedef Point3_<int> Point3i;
template <typename T> struct Size_ { T width, height; Size_() : width(0), height(0) {} Size_(T w, T h) : width(w), height(h) {} };
typedef Size_<int> Size;
struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  template <typename T, int N> Scalar(const Vec<T, N>& v) { for (int i = 0; i < 4; i++) val[i] = i < N ? (double)v[i] : 0.0; }
  double& operator[](int i) { return val[i]; }
  const double& operator[](int i) const { return val[i]; }
  Scalar& operator+=(const Scalar& o) { for (int i = 0; i < 4; i++) val[i] += o.val[i]; return *this; }
};
struct Range { int start, end; Range(int s = 0, int e = 0) : start(s), end(e) {} };
struct Rect { int x, y, width, height; Rect(int a = 0, int b = 0, int c = 0, int d = 0) : x(a), y(b), width(c), height(d) {} };
struct Exception : public std::exception { const char* what() const noexcept override { return "cv::Exception (stub)"; } };
enum { BORDER_REFLECT_101 = 4, FM_LMEDS = 4, FM_RANSAC = 8, FM_8POINT = 2, LINE_8 = 8, LINE_AA = 16, FONT_HERSHEY_SIMPLEX = 0, IMREAD_COLOR = 1, COLOR_BGR2GRAY = 6, COLOR_GRAY2BGR = 8, COLOR_BGR2RGB = 4 };
inline int borderInterpolate(int, int, int) { eg3d_stub_unreachable("borderInterpolate"); }

inline int cvRound(double v) { return (int)lrint(v); }
class Mat {
 public:
  int rows, cols;
  uchar* data;   /* non-null iff allocated */
  bool lazy_t;   /* result of .t(): an operand of the next product that OpenCV would pass as a GEMM_x_T flag */
  Mat() : rows(0), cols(0), data(nullptr), lazy_t(false), type_(0) {}
  Mat(int r, int c, int type) : lazy_t(false) { create(r, c, type); }
  Mat(Size s, int type) : lazy_t(false) { create(s.height, s.width, type); }
  Mat(int r, int c, int type, const Scalar& s) : lazy_t(false) { create(r, c, type); fill(s); }
  Mat(Size sz, int type, const Scalar& s) : lazy_t(false) { create(sz.height, sz.width, type); fill(s); }
  explicit Mat(const std::vector<Point2f>& v) : lazy_t(false) {   /* N x 1, CV_32FC2 */
    create((int)v.size(), 1, CV_32FC2);
    for (size_t i = 0; i < v.size(); i++) { ((float*)ptr())[2 * i] = v[i].x; ((float*)ptr())[2 * i + 1] = v[i].y; }
  }
  void create(int r, int c, int type) { rows = r; cols = c; type_ = type; data_ = std::make_shared<std::vector<uchar>>((size_t)r * c * elemSize(), 0); data = data_->data(); }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
  size_t elemSize() const { const int d = depth(); return (size_t)channels() * (d == CV_8U ? 1 : d == CV_32F ? 4 : 8); }
  bool empty() const { return !data_ || rows * cols == 0; }
  Size size() const { return Size(cols, rows); }
  uchar* ptr() { return data_->data(); }
  const uchar* ptr() const { return data_->data(); }
  template <typename T> T& at(int r, int c) { return *(T*)(data_->data() + ((size_t)r * cols + c) * sizeof(T)); }
  template <typename T> const T& at(int r, int c) const { return *(const T*)(data_->data() + ((size_t)r * cols + c) * sizeof(T)); }
  template <typename T> T& at(int i) { return *(T*)(data_->data() + (size_t)i * sizeof(T)); }
  template <typename T> const T& at(int i) const { return *(const T*)(data_->data() + (size_t)i * sizeof(T)); }
  template <typename T> T& at(Point p) { return at<T>(p.y, p.x); }
  template <typename T> const T& at(Point p) const { return at<T>(p.y, p.x); }
  Mat clone() const { Mat m; m.rows = rows; m.cols = cols; m.type_ = type_; m.lazy_t = lazy_t; if (data_) { m.data_ = std::make_shared<std::vector<uchar>>(*data_); m.data = m.data_->data(); } return m; }
  void copyTo(Mat& o) const { o = clone(); }
  void convertTo(Mat& o, int rtype) const {
    Mat r(rows, cols, CV_MAKETYPE(rtype & 7, channels()));
    const size_t n = (size_t)rows * cols * channels();
    for (size_t i = 0; i < n; i++) r.set_elem(i, get_elem(i));
    o = r;
  }
  double get_elem(size_t i) const { const int d = depth(); return d == CV_64F ? ((const double*)ptr())[i] : d == CV_32F ? (double)((const float*)ptr())[i] : (double)ptr()[i]; }
  void set_elem(size_t i, double v) { const int d = depth(); if (d == CV_64F) ((double*)ptr())[i] = v; else if (d == CV_32F) ((float*)ptr())[i] = (float)v; else ptr()[i] = (uchar)v; }
  void fill(const Scalar& s) { const int cn = channels(); const size_t n = (size_t)rows * cols; for (size_t i = 0; i < n; i++) for (int c = 0; c < cn; c++) set_elem(i * cn + c, s.val[c & 3]); }
  static Mat header_only(int r, int c, int type) { Mat m; m.rows = r; m.cols = c; m.type_ = type; return m; }   /* size without pixels (stub only) */
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  static Mat zeros(Size s, int type) { return Mat(s, type); }
  static Mat eye(int r, int c, int type) { Mat m(r, c, type); for (int i = 0; i < r && i < c; i++) m.set_elem((size_t)i * c + i, 1.0); return m; }
  /* element (r, c) of the matrix this object stands for (the transpose of the stored one when lazy_t) */
  double get(int r, int c) const { return lazy_t ? get_elem((size_t)c * cols + r) : get_elem((size_t)r * cols + c); }
  int lrows() const { return lazy_t ? cols : rows; }
  int lcols() const { return lazy_t ? rows : cols; }
  void release() { rows = cols = 0; data_.reset(); data = nullptr; }
  Mat t() const { Mat m = *this; m.lazy_t = !lazy_t; return m; }   /* shares the data, as a MatExpr would */
  Mat inv() const;
  Mat& operator+=(const Mat& o) {
    assert(!lazy_t && !o.lazy_t && rows == o.rows && cols == o.cols && type_ == o.type_);
    const size_t n = (size_t)rows * cols;
    if (depth() == CV_64F) for (size_t i = 0; i < n; i++) ((double*)ptr())[i] += ((const double*)o.ptr())[i];
    else if (depth() == CV_32F) for (size_t i = 0; i < n; i++) ((float*)ptr())[i] += ((const float*)o.ptr())[i];
    else eg3d_stub_unreachable("Mat += on this depth");
    return *this;
  }
 private:
  int type_;
  std::shared_ptr<std::vector<uchar>> data_;
};
template <typename T> class Mat_ : public Mat { public: Mat_() {} Mat_(int r, int c) : Mat(r, c, sizeof(T) == 8 ? CV_64F : CV_32F) {} Mat_(int r, int c, int) : Mat(r, c, sizeof(T) == 8 ? CV_64F : CV_32F) {} T& operator()(int r, int c) { return this->template at<T>(r, c); } };
typedef const Mat& InputArray; typedef Mat& OutputArray;

inline Mat operator*(const Mat& a, const Mat& b) {
  const int M = a.lrows(), K = a.lcols(), N = b.lcols();
  assert(K == b.lrows() && a.depth() == b.depth());
  Mat d(M, N, a.type());
  if (a.depth() == CV_64F) {
    for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) { double s = 0; for (int k = 0; k < K; k++) s += a.get(i, k) * b.get(k, j); d.at<double>(i, j) = s; }
  } else if (a.depth() == CV_32F) {
    const bool small = !a.lazy_t && !b.lazy_t && 2 <= K && K <= 4 && (K == N || K == M);   /* OpenCV gemm: flags == 0 && 2 <= len <= 4 && (len == d_size.width || len == d_size.height) */
    for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) {
      if (small) { float s = (float)a.get(i, 0) * (float)b.get(0, j); for (int k = 1; k < K; k++) s += (float)a.get(i, k) * (float)b.get(k, j); d.at<float>(i, j) = s; }
      else { double s = 0; for (int k = 0; k < K; k++) s += a.get(i, k) * b.get(k, j); d.at<float>(i, j) = (float)s; }
    }
  } else eg3d_stub_unreachable("Mat * Mat on this depth");
  return d;
}
inline Mat operator-(const Mat&, const Mat&) { eg3d_stub_unreachable("Mat - Mat"); }
inline double determinant(const Mat& m) {
  assert(m.rows == 3 && m.cols == 3);
  double a[9]; for (int i = 0; i < 9; i++) a[i] = m.get(i / 3, i % 3);
  return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
}
inline Mat Mat::inv() const {
  assert(rows == 3 && cols == 3);
  double m[9]; for (int i = 0; i < 9; i++) m[i] = get(i / 3, i % 3);
  double d = determinant(*this);
  Mat r(3, 3, type());
  if (d != 0.) {
    d = 1. / d;
    const double t[9] = {(m[4] * m[8] - m[5] * m[7]) * d, (m[2] * m[7] - m[1] * m[8]) * d, (m[1] * m[5] - m[2] * m[4]) * d,
                         (m[5] * m[6] - m[3] * m[8]) * d, (m[0] * m[8] - m[2] * m[6]) * d, (m[2] * m[3] - m[0] * m[5]) * d,
                         (m[3] * m[7] - m[4] * m[6]) * d, (m[1] * m[6] - m[0] * m[7]) * d, (m[0] * m[4] - m[1] * m[3]) * d};
    for (int i = 0; i < 9; i++) r.set_elem(i, t[i]);
  }
  return r;
}

/* bit-exact restatements, defined in oracle/ref_path_wrapper.cpp on top of the oracle's functions */
void triangulatePoints(const Mat& P1, const Mat& P2, const std::vector<Point2f>& x1, const std::vector<Point2f>& x2, Vec4f& out);
void computeCorrespondEpilines(const Mat& points, int whichImage, const Mat& F, std::vector<Vec3f>& lines);
inline Mat findFundamentalMat(const std::vector<Point2f>&, const std::vector<Point2f>&, int = 0, double = 3., double = 0.99) { eg3d_stub_unreachable("findFundamentalMat (F is an input of the path)"); }
inline Mat imread(const std::string&, int = 1) { eg3d_stub_unreachable("imread"); }
inline bool imwrite(const std::string&, const Mat&) { eg3d_stub_unreachable("imwrite"); }
inline void cvtColor(const Mat& in, Mat& out, int) { out = Mat::header_only(in.rows, in.cols, CV_8UC1); }   /* PLGEdgeManager's ctor keeps greyscale copies nobody reads */
inline void line(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) { eg3d_stub_unreachable("line"); }
inline void circle(Mat&, Point, int, const Scalar&, int = 1, int = 8, int = 0) { eg3d_stub_unreachable("circle"); }
inline void rectangle(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) { eg3d_stub_unreachable("rectangle"); }
inline void putText(Mat&, const std::string&, Point, int, double, Scalar, int = 1, int = 8, bool = false) { eg3d_stub_unreachable("putText"); }
}  // namespace cv
