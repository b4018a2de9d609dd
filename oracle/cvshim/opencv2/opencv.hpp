// A stand-in for the OpenCV C++ headers, just wide enough to compile the reference's processing block VERBATIM
// (oracle/build_ref.py cuts it out of /root/reference/BscanFFT.cpp at build time; nothing of it is stored in this repository).
// TEST INFRASTRUCTURE ONLY - see the header of oracle/abcoct_oracle.py.
//
// This image has OpenCV only as the Python module cv2 (no C++ headers, no linkable libopencv), so cv::Mat here is a header over a
// NumPy array and every OpenCV *function* (dft, normalize, resize, medianBlur, magnitude, log, accumulate, ...) is forwarded to the
// very same OpenCV kernels through cv2 (cvshim_ops.py).  What this file restates of OpenCV is the C++ surface semantics the block
// relies on:
//   * Mat headers share data; row / colRange / rowRange / Mat(m, Rect) are views; at<T>(r, c) is unchecked pointer arithmetic
//     (the block indexes N x 1 column vectors as (0, i) and a transposed-shape `slopes` as (p, q) - both only work that way);
//   * an assignment or OutputArray write into a Mat of the same size and type goes INTO the existing buffer (Mat::create is a
//     no-op then) - `data_y.row(p) = data_y.row(p) - mean` depends on it; anything else re-binds the header;
//   * lazy MatExpr folding (modules/core/src/matop.cpp): `Mat / s` is a convertTo with alpha = 1 / s, `20.0 * m / 2.303` folds to
//     ONE multiplication by 20.0 * (1 / 2.303), `(a - b) / c` is cv::divide(cv::subtract(a, b), c), `s / m` is cv::divide(s, m),
//     `m - s` is cv::add(m, -s), `max(m, s)` is cv::max.
#pragma once
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <string>
#include <sys/types.h>
#include <vector>

namespace cv {
namespace py = pybind11;

enum { CV_8U = 0, CV_8S = 1, CV_16U = 2, CV_16S = 3, CV_32S = 4, CV_32F = 5, CV_64F = 6 };
enum { CV_8UC1 = 0, CV_16UC1 = 2, CV_32SC1 = 4, CV_32FC1 = 5, CV_64FC1 = 6, CV_32FC2 = 5 + 8, CV_64FC2 = 6 + 8 };
enum { DFT_INVERSE = 1, DFT_SCALE = 2, DFT_ROWS = 4, DFT_COMPLEX_OUTPUT = 16, DFT_REAL_OUTPUT = 32 };
enum { NORM_MINMAX = 32 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { BORDER_CONSTANT = 0 };
enum { FONT_HERSHEY_SIMPLEX = 0 };
enum { THRESH_BINARY = 0 };
enum { COLORMAP_JET = 2 };

inline py::module_& ops() {
  static py::module_* m = new py::module_(py::module_::import("cvshim_ops"));
  return *m;
}

struct Size {
  int width = 0, height = 0;
  Size() {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
  int x, y, width, height;
  Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};
struct Point {
  int x, y;
  Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Scalar {
  double v[4] = {0, 0, 0, 0};
  Scalar() {}
  Scalar(double a) { v[0] = a; }
  Scalar(double a, double b, double c = 0, double d = 0) {
    v[0] = a, v[1] = b, v[2] = c, v[3] = d;
  }
  static Scalar all(double a) { return Scalar(a, a, a, a); }
  double operator()(int i) const { return v[i]; }
  double operator[](int i) const { return v[i]; }
};

class Mat;
struct MatExpr;
struct OutputArray;

class Mat {
 public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  size_t step = 0;
  py::object arr;  // numpy.ndarray (rows, cols) or (rows, cols, channels); None when empty

  Mat() {}
  Mat(int r, int c, int type) { bind(ops().attr("zeros")(r, c, type & 7, (type >> 3) + 1)); }
  Mat(const Mat& m, const Rect& r) { bind(ops().attr("view")(m.arr, r.y, r.x, r.height, r.width)); }
  Mat(const MatExpr& e);
  explicit Mat(py::object a) { bind(std::move(a)); }

  static Mat zeros(Size s, int type) { return Mat(s.height, s.width, type); }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }

  void bind(py::object a) {
    arr = std::move(a);
    py::array na = py::reinterpret_borrow<py::array>(arr);
    rows = (int)na.shape(0);
    cols = (int)na.shape(1);
    data = static_cast<unsigned char*>(const_cast<void*>(na.data()));
    step = (size_t)na.strides(0);
  }
  bool empty() const { return data == nullptr; }
  int depth() const { return ops().attr("depth_of")(arr).cast<int>(); }
  int channels() const {
    py::array na = py::reinterpret_borrow<py::array>(arr);
    return na.ndim() == 3 ? (int)na.shape(2) : 1;
  }
  int type() const { return depth() + ((channels() - 1) << 3); }
  Size size() const { return Size(cols, rows); }
  bool same_layout(const py::object& other) const {
    if (empty()) return false;
    py::array a = py::reinterpret_borrow<py::array>(arr), b = py::reinterpret_borrow<py::array>(other);
    if (a.ndim() != b.ndim() || !a.dtype().is(b.dtype())) return false;
    for (py::ssize_t i = 0; i < a.ndim(); ++i)
      if (a.shape(i) != b.shape(i)) return false;
    return true;
  }
  // Mat::create + write: into the existing buffer when size and type match, a fresh buffer otherwise
  void put(py::object result) {
    if (same_layout(result))
      ops().attr("assign")(arr, result);
    else
      bind(ops().attr("copy")(result));
  }

  template <class T>
  T& at(int r, int c) {
    return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T));
  }
  template <class T>
  T& at(int i) {  // continuous storage in every use the block makes of it
    return reinterpret_cast<T*>(data)[i];
  }
  template <class T>
  T* ptr(int r) {
    return reinterpret_cast<T*>(data + (size_t)r * step);
  }

  Mat view(int y, int x, int h, int w) const { return Mat(ops().attr("view")(arr, y, x, h, w)); }
  Mat row(int i) const { return view(i, 0, 1, cols); }
  Mat col(int i) const { return view(0, i, rows, 1); }
  Mat rowRange(int a, int b) const { return view(a, 0, b - a, cols); }
  Mat colRange(int a, int b) const { return view(0, a, rows, b - a); }
  Mat operator()(const Rect& r) const { return view(r.y, r.x, r.height, r.width); }

  inline void copyTo(const OutputArray& dst) const;
  inline void convertTo(const OutputArray& dst, int rtype, double alpha = 1, double beta = 0) const;

  Mat& operator=(const MatExpr& e);
  Mat& operator=(const Scalar& s) {  // setTo
    ops().attr("set_all")(arr, s.v[0]);
    return *this;
  }
  Mat& operator+=(const Scalar& s) {  // cv::add(m, s, m)
    put(ops().attr("add_scalar")(arr, s.v[0]));
    return *this;
  }
};

// const _OutputArray&: binds to an lvalue Mat or to a temporary header (m.row(i)) alike
struct OutputArray {
  Mat* p;
  mutable Mat tmp;
  OutputArray(Mat& m) : p(&m) {}
  OutputArray(Mat&& m) : p(nullptr), tmp(m) {}
  Mat& mat() const { return p ? *p : tmp; }
};

inline void Mat::copyTo(const OutputArray& dst) const { dst.mat().put(arr); }
inline void Mat::convertTo(const OutputArray& dst, int rtype, double alpha, double beta) const {
  dst.mat().put(ops().attr("convert")(arr, rtype & 7, alpha, beta));
}

template <class T>
struct DepthOf;
template <>
struct DepthOf<float> {
  enum { value = CV_32F };
};
template <>
struct DepthOf<double> {
  enum { value = CV_64F };
};
template <class T>
class Mat_ : public Mat {
 public:
  Mat_(const Mat& m) { m.convertTo(*this, DepthOf<T>::value); }
};

// ---- lazy expressions (modules/core/src/matop.cpp), only the shapes the block uses
struct MatExpr {
  enum Kind { ADDEX, DIV_MM, DIV_SM, MAX_MS } kind = ADDEX;
  Mat a, b;
  double alpha = 1, beta = 0, s = 0;  // ADDEX: alpha * a + beta * b + s; DIV_SM: s / a; MAX_MS: max(a, s)
  py::object eval(const Mat* into) const {
    switch (kind) {
      case DIV_MM: return ops().attr("divide")(a.arr, b.arr);
      case DIV_SM: return ops().attr("divide_scalar_by")(s, a.arr);
      case MAX_MS: return ops().attr("max_scalar")(a.arr, s);
      default: break;
    }
    if (!b.empty()) {
      if (s == 0 && alpha == 1 && beta == -1) return ops().attr("subtract")(a.arr, b.arr);
      if (s == 0 && alpha == 1 && beta == 1) return ops().attr("add")(a.arr, b.arr);
      throw std::runtime_error("cvshim: a MatExpr shape the reference block does not use");
    }
    // MatOp_AddEx::assign: a real scalar with |alpha| != 1 (or a destination of another buffer) is ONE convertTo(alpha, s);
    // alpha == 1 written over its own operand is cv::add(a, s)
    const bool own = into && into->data == a.data;
    if (alpha == 1 && (own || s != 0)) return ops().attr("add_scalar")(a.arr, s);
    return ops().attr("convert")(a.arr, a.depth(), alpha, s);
  }
  bool scaled() const { return kind == ADDEX && b.empty() && s == 0; }
};
inline Mat::Mat(const MatExpr& e) { bind(ops().attr("copy")(e.eval(nullptr))); }
inline Mat& Mat::operator=(const MatExpr& e) {
  put(e.eval(this));
  return *this;
}

inline MatExpr operator+(const Mat& a, const Mat& b) {  // AddEx(a, b, 1, 1) -> cv::add
  MatExpr e;
  e.a = a, e.b = b, e.alpha = 1, e.beta = 1;
  return e;
}
inline MatExpr operator-(const Mat& a, const Mat& b) {
  MatExpr e;
  e.a = a, e.b = b, e.alpha = 1, e.beta = -1;
  return e;
}
inline MatExpr operator+(const MatExpr& x, const MatExpr& y) {  // MatOp::add: operands with a second matrix are evaluated first
  MatExpr e;
  e.a = Mat(x), e.b = Mat(y), e.alpha = 1, e.beta = 1;
  return e;
}
inline MatExpr operator-(const Mat& a, double s) {  // operator-(const Mat&, const Scalar&): AddEx(a, 1, 0, -s)
  MatExpr e;
  e.a = a, e.s = -s;
  return e;
}
inline MatExpr operator/(const Mat& a, double s) {  // AddEx(a, 1. / s)
  MatExpr e;
  e.a = a, e.alpha = 1. / s;
  return e;
}
inline MatExpr operator*(double s, const Mat& a) {
  MatExpr e;
  e.a = a, e.alpha = s;
  return e;
}
inline MatExpr operator*(const Mat& a, double s) { return s * a; }
inline MatExpr operator/(const MatExpr& x, double s) {  // MatOp_AddEx::multiply(e, 1. / s)
  if (x.kind != MatExpr::ADDEX) throw std::runtime_error("cvshim: expr / scalar on a non-linear expression");
  MatExpr e = x;
  const double f = 1. / s;
  e.alpha *= f, e.beta *= f, e.s *= f;
  return e;
}
inline MatExpr operator/(const MatExpr& x, const Mat& m) {  // MatOp::divide: evaluate the left side, then cv::divide
  MatExpr e;
  e.kind = MatExpr::DIV_MM;
  e.a = Mat(x), e.b = m;
  return e;
}
inline MatExpr operator/(double s, const Mat& a) {  // MatOp_Bin '/', cv::divide(s, a)
  MatExpr e;
  e.kind = MatExpr::DIV_SM;
  e.a = a, e.s = s;
  return e;
}
inline MatExpr max(const Mat& a, double s) {
  MatExpr e;
  e.kind = MatExpr::MAX_MS;
  e.a = a, e.s = s;
  return e;
}

// ---- functions, forwarded to OpenCV through cv2
inline void max(const Mat& a, double s, const OutputArray& dst) { dst.mat().put(ops().attr("max_scalar")(a.arr, s)); }
inline void normalize(const Mat& src, const OutputArray& dst, double a, double b, int norm_type) {
  dst.mat().put(ops().attr("normalize")(src.arr, a, b, norm_type));
}
inline Scalar mean(const Mat& m) { return Scalar(ops().attr("mean0")(m.arr).cast<double>()); }
inline void multiply(const Mat& a, const Mat& b, const OutputArray& dst) { dst.mat().put(ops().attr("multiply")(a.arr, b.arr)); }
inline void merge(const Mat* mv, size_t n, const OutputArray& dst) {
  py::list l;
  for (size_t i = 0; i < n; ++i) l.append(mv[i].arr);
  dst.mat().put(ops().attr("merge")(l));
}
inline void split(const Mat& m, Mat* mv) {
  py::list l = ops().attr("split")(m.arr);
  for (size_t i = 0; i < l.size(); ++i) mv[i].put(py::reinterpret_borrow<py::object>(l[i]));
}
inline void dft(const Mat& src, const OutputArray& dst, int flags = 0) { dst.mat().put(ops().attr("dft")(src.arr, flags)); }
inline void magnitude(const Mat& x, const Mat& y, const OutputArray& dst) { dst.mat().put(ops().attr("magnitude")(x.arr, y.arr)); }
inline void accumulate(const Mat& src, Mat& dst) { ops().attr("accumulate")(src.arr, dst.arr); }
inline void transpose(const Mat& src, const OutputArray& dst) { dst.mat().put(ops().attr("transpose")(src.arr)); }
inline void log(const Mat& src, const OutputArray& dst) { dst.mat().put(ops().attr("log")(src.arr)); }
inline void copyMakeBorder(const Mat& src, const OutputArray& dst, int t, int b, int l, int r, int btype, const Scalar& v = Scalar()) {
  dst.mat().put(ops().attr("copy_make_border")(src.arr, t, b, l, r, btype, v.v[0]));
}
inline void medianBlur(const Mat& src, const OutputArray& dst, int k) { dst.mat().put(ops().attr("median_blur")(src.arr, k)); }
inline void resize(const Mat& src, const OutputArray& dst, Size, double fx, double fy, int interp) {
  dst.mat().put(ops().attr("resize")(src.arr, fx, fy, interp));
}
inline double threshold(const Mat& src, const OutputArray& dst, double thresh, double maxval, int type) {
  dst.mat().put(ops().attr("threshold")(src.arr, thresh, maxval, type));
  return thresh;
}
inline void applyColorMap(const Mat& src, const OutputArray& dst, int colormap) {
  dst.mat().put(ops().attr("apply_color_map")(src.arr, colormap));
}
// display calls of the camera loop: nothing to show here
inline void putText(Mat&, const char*, Point, int, double, Scalar, int = 1, int = 8) {}
inline void imshow(const char*, const Mat&) {}
inline void resizeWindow(const char*, int, int) {}

}  // namespace cv

#define CV_Assert(x)                                                  \
  do {                                                                \
    if (!(x)) throw std::runtime_error("CV_Assert failed: " #x);      \
  } while (0)
