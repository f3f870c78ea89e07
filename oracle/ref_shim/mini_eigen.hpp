// mini_eigen.hpp -- a small, eager stand-in for the subset of Eigen 3 that the REFERENCE's own
// state-estimator/src/mav_state_est/rbis.{hpp,cpp} uses, so that those files can be compiled here UNMODIFIED
// (Eigen is not installed and there is no network).  TEST INFRASTRUCTURE ONLY (oracle/_ref), never product code.
// Every expression evaluates immediately into a concrete Matrix with plain double loops; Eigen's lazy
// evaluation, vectorisation and summation order are not reproduced (differences are at rounding level).
// Semantics restated from Eigen's documentation: column-major storage, Quaternion (w,x,y,z) with
// inverse() = conjugate / squaredNorm, q * v by the 2*cross form, LDLT solve, determinant.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <type_traits>
#include <sstream>
#include <string>
#include <vector>

namespace Eigen {

const int Dynamic = -1;

template <typename T, int R, int C> class Matrix;

template <class D> struct traits;

// ---- CRTP base: anything with rows(), cols(), coeff(i, j) ----
template <class D>
class MatrixBase {
 public:
  const D& derived() const { return *static_cast<const D*>(this); }
  D& derived() { return *static_cast<D*>(this); }
  typedef typename traits<D>::Scalar Scalar;
  enum { RowsAtCompileTime = traits<D>::Rows, ColsAtCompileTime = traits<D>::Cols };
  typedef Matrix<Scalar, traits<D>::Rows, traits<D>::Cols> PlainObject;

  int rows() const { return derived().rows_(); }
  int cols() const { return derived().cols_(); }
  int size() const { return rows() * cols(); }
  Scalar coeff(int i, int j) const { return derived().coeff_(i, j); }
  Scalar coeff(int i) const { return cols() == 1 ? coeff(i, 0) : coeff(0, i); }
  Scalar operator()(int i, int j) const { return coeff(i, j); }
  Scalar operator()(int i) const { return cols() == 1 ? coeff(i, 0) : coeff(0, i); }
  PlainObject eval() const {
    PlainObject m(rows(), cols());
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) m.ref(i, j) = coeff(i, j);
    return m;
  }
  Matrix<Scalar, traits<D>::Cols, traits<D>::Rows> transpose() const {
    Matrix<Scalar, traits<D>::Cols, traits<D>::Rows> m(cols(), rows());
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) m.ref(j, i) = coeff(i, j);
    return m;
  }
  Scalar squaredNorm() const {
    Scalar s = 0;
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) s += coeff(i, j) * coeff(i, j);
    return s;
  }
  Scalar norm() const { return std::sqrt(squaredNorm()); }
  template <class O>
  Matrix<Scalar, 3, 1> cross(const MatrixBase<O>& o) const {
    Matrix<Scalar, 3, 1> r;
    const Scalar a0 = (*this)(0), a1 = (*this)(1), a2 = (*this)(2), b0 = o(0), b1 = o(1), b2 = o(2);
    r.ref(0, 0) = a1 * b2 - a2 * b1;
    r.ref(1, 0) = a2 * b0 - a0 * b2;
    r.ref(2, 0) = a0 * b1 - a1 * b0;
    return r;
  }
  template <class O>
  Scalar dot(const MatrixBase<O>& o) const {
    Scalar s = 0;
    for (int i = 0; i < size(); i++) s += (*this)(i) * o(i);
    return s;
  }
  Scalar determinant() const;  // defined after Matrix
  Matrix<Scalar, (traits<D>::Rows == 1 ? traits<D>::Cols : traits<D>::Rows), 1> diagonal() const {
    const int n = rows() < cols() ? rows() : cols();
    Matrix<Scalar, (traits<D>::Rows == 1 ? traits<D>::Cols : traits<D>::Rows), 1> d(n, 1);
    for (int i = 0; i < n; i++) d.ref(i, 0) = coeff(i, i);
    return d;
  }
  // 1x1 results convert to their scalar, as in Eigen
  template <int R_ = traits<D>::Rows, int C_ = traits<D>::Cols, typename = typename std::enable_if<R_ == 1 && C_ == 1>::type>
  operator Scalar() const { return coeff(0, 0); }
};

// ---- writable mixin (Matrix, Block, Map, TransposeRef) ----
template <class D>
class WritableBase : public MatrixBase<D> {
 public:
  typedef typename traits<D>::Scalar Scalar;
  using MatrixBase<D>::derived;
  using MatrixBase<D>::rows;
  using MatrixBase<D>::cols;
  using MatrixBase<D>::operator();
  Scalar& operator()(int i, int j) { return derived().ref(i, j); }
  Scalar& operator()(int i) { return cols() == 1 ? derived().ref(i, 0) : derived().ref(0, i); }
  D& noalias() { return derived(); }
  template <class O>
  D& assign(const MatrixBase<O>& o) {
    const typename MatrixBase<O>::PlainObject v = o.eval();  // aliasing-safe
    derived().resize_like(v.rows(), v.cols());
    assert(rows() == v.rows() && cols() == v.cols());
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) derived().ref(i, j) = v.coeff(i, j);
    return derived();
  }
  template <class O>
  D& operator+=(const MatrixBase<O>& o) {
    const typename MatrixBase<O>::PlainObject v = o.eval();
    assert(rows() == v.rows() && cols() == v.cols());
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) derived().ref(i, j) += v.coeff(i, j);
    return derived();
  }
  template <class O>
  D& operator-=(const MatrixBase<O>& o) {
    const typename MatrixBase<O>::PlainObject v = o.eval();
    assert(rows() == v.rows() && cols() == v.cols());
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) derived().ref(i, j) -= v.coeff(i, j);
    return derived();
  }
  D& operator*=(Scalar s) {
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) derived().ref(i, j) *= s;
    return derived();
  }
  D& setZero() {
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) derived().ref(i, j) = 0;
    return derived();
  }
  D& setIdentity() {
    for (int j = 0; j < cols(); j++)
      for (int i = 0; i < rows(); i++) derived().ref(i, j) = (i == j) ? 1 : 0;
    return derived();
  }
};

// ---- blocks ----
template <class P, int R, int C> class Block;
template <class P, int R, int C>
struct traits<Block<P, R, C> > {
  typedef typename traits<typename std::remove_const<P>::type>::Scalar Scalar;
  enum { Rows = R, Cols = C };
};
template <class P, int R, int C>
class Block : public WritableBase<Block<P, R, C> > {
 public:
  typedef typename traits<Block>::Scalar Scalar;
  Block(P& p, int i0, int j0, int r, int c) : p_(&p), i0_(i0), j0_(j0), r_(r), c_(c) {}
  Block(const Block&) = default;
  int rows_() const { return r_; }
  int cols_() const { return c_; }
  Scalar coeff_(int i, int j) const { return p_->coeff(i0_ + i, j0_ + j); }
  Scalar& ref(int i, int j) { return p_->ref(i0_ + i, j0_ + j); }
  void resize_like(int, int) {}
  template <class O>
  Block& operator=(const MatrixBase<O>& o) { return this->assign(o); }
  Block& operator=(const Block& o) { return this->assign(o); }

 private:
  P* p_;
  int i0_, j0_, r_, c_;
};

// ---- lvalue transpose: K.transpose() = X ----
template <class P> class TransposeRef;
template <class P>
struct traits<TransposeRef<P> > {
  typedef typename traits<P>::Scalar Scalar;
  enum { Rows = traits<P>::Cols, Cols = traits<P>::Rows };
};
template <class P>
class TransposeRef : public WritableBase<TransposeRef<P> > {
 public:
  typedef typename traits<P>::Scalar Scalar;
  explicit TransposeRef(P& p) : p_(&p) {}
  int rows_() const { return p_->cols(); }
  int cols_() const { return p_->rows(); }
  Scalar coeff_(int i, int j) const { return p_->coeff(j, i); }
  Scalar& ref(int i, int j) { return p_->ref(j, i); }
  void resize_like(int r, int c) { p_->resize_like(c, r); }
  template <class O>
  TransposeRef& operator=(const MatrixBase<O>& o) { return this->assign(o); }

 private:
  P* p_;
};

template <class V> class DiagonalWrapper;
template <class V> class ArrayWrapper;
template <class M> class LDLT;
template <class M> class LLT;
template <typename T> class Quaternion;

// ---- dense matrix ----
template <typename T, int R, int C>
struct traits<Matrix<T, R, C> > {
  typedef T Scalar;
  enum { Rows = R, Cols = C };
};
template <typename T, int R, int C>
class Matrix : public WritableBase<Matrix<T, R, C> > {
 public:
  typedef T Scalar;
  typedef WritableBase<Matrix> Base;
  using Base::operator();
  Matrix() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C), d_((size_t)r_ * c_, T(0)) {}
  explicit Matrix(int n) : r_(R == Dynamic ? n : R), c_(C == Dynamic ? (R == Dynamic ? 1 : n) : C), d_((size_t)r_ * c_, T(0)) {
    if (R != Dynamic && C != Dynamic && R * C == 1) d_[0] = T(n);  // Matrix<T,1,1>(value)
  }
  Matrix(int r, int c) : r_(R == Dynamic ? r : R), c_(C == Dynamic ? c : C), d_((size_t)r_ * c_, T(0)) {
    if (R != Dynamic && C != Dynamic && R * C == 2) { d_[0] = T(r); d_[1] = T(c); }  // Vector2(x, y)
  }
  Matrix(T x, T y, T z) : r_(R), c_(C), d_(3) {
    static_assert(R * C == 3, "3-coefficient constructor");
    d_[0] = x; d_[1] = y; d_[2] = z;
  }
  explicit Matrix(const T* p) : r_(R), c_(C), d_((size_t)R * C) {  // fixed-size matrix from a coefficient array (Eigen: Vector3d(ptr))
    static_assert(R != Dynamic && C != Dynamic, "pointer constructor needs a fixed size");
    for (size_t k = 0; k < d_.size(); k++) d_[k] = p[k];
  }
  template <class O>
  Matrix(const MatrixBase<O>& o) : r_(o.rows()), c_(o.cols()), d_((size_t)o.rows() * o.cols()) {
    assert((R == Dynamic || R == r_) && (C == Dynamic || C == c_));
    for (int j = 0; j < c_; j++)
      for (int i = 0; i < r_; i++) ref(i, j) = o.coeff(i, j);
  }
  template <class V>
  Matrix(const DiagonalWrapper<V>& dw);
  Matrix(const Matrix&) = default;
  Matrix& operator=(const Matrix&) = default;
  template <class O>
  Matrix& operator=(const MatrixBase<O>& o) { return this->assign(o); }

  int rows_() const { return r_; }
  int cols_() const { return c_; }
  T coeff_(int i, int j) const { return d_[(size_t)i + (size_t)r_ * j]; }
  T& ref(int i, int j) { return d_[(size_t)i + (size_t)r_ * j]; }
  void resize_like(int r, int c) {
    if (r == r_ && c == c_) return;
    assert((R == Dynamic || R == r) && (C == Dynamic || C == c));
    r_ = r; c_ = c;
    d_.assign((size_t)r * c, T(0));
  }
  void resize(int r, int c = 1) { resize_like(r, c); }
  T* data() { return d_.data(); }
  const T* data() const { return d_.data(); }
  T& operator[](int i) { return d_[(size_t)i]; }   // vectors only
  const T& operator[](int i) const { return d_[(size_t)i]; }
  T x() const { return d_[0]; }
  T y() const { return d_[1]; }
  T z() const { return d_[2]; }
  T& x() { return d_[0]; }
  T& y() { return d_[1]; }
  T& z() { return d_[2]; }

  static Matrix Zero() { return Matrix(); }
  static Matrix Zero(int r, int c = 1) { return Matrix(r, c); }
  static Matrix Ones() {
    Matrix m;
    for (auto& v : m.d_) v = T(1);
    return m;
  }
  static Matrix Identity() {
    Matrix m;
    m.setIdentity();
    return m;
  }
  static Matrix Identity(int r, int c) {
    Matrix m(r, c);
    m.setIdentity();
    return m;
  }
  static Matrix LinSpaced(T low, T high) {  // fixed-size vector
    Matrix m;
    const int n = m.size();
    for (int i = 0; i < n; i++) m(i) = (n == 1) ? high : (T)(low + (high - low) * i / (n - 1));
    return m;
  }

  // blocks (mutable and const)
  template <int BR, int BC> Block<Matrix, BR, BC> block(int i, int j) { return Block<Matrix, BR, BC>(*this, i, j, BR, BC); }
  template <int BR, int BC> Block<const Matrix, BR, BC> block(int i, int j) const { return Block<const Matrix, BR, BC>(*this, i, j, BR, BC); }
  Block<Matrix, Dynamic, Dynamic> block(int i, int j, int r, int c) { return Block<Matrix, Dynamic, Dynamic>(*this, i, j, r, c); }
  template <int N> Block<Matrix, N, 1> segment(int i) { return Block<Matrix, N, 1>(*this, i, 0, N, 1); }
  template <int N> Block<const Matrix, N, 1> segment(int i) const { return Block<const Matrix, N, 1>(*this, i, 0, N, 1); }
  template <int N> Block<Matrix, N, 1> head() { return segment<N>(0); }
  template <int N> Block<const Matrix, N, 1> head() const { return segment<N>(0); }
  template <int N> Block<Matrix, N, 1> tail() { return segment<N>(r_ - N); }
  template <int N> Block<const Matrix, N, 1> tail() const { return segment<N>(r_ - N); }
  Block<Matrix, Dynamic, 1> head(int n) { return Block<Matrix, Dynamic, 1>(*this, 0, 0, n, 1); }
  Block<Matrix, Dynamic, 1> segment(int i, int n) { return Block<Matrix, Dynamic, 1>(*this, i, 0, n, 1); }
  Block<Matrix, Dynamic, 1> tail(int n) { return Block<Matrix, Dynamic, 1>(*this, r_ - n, 0, n, 1); }

  Block<Matrix, R, 1> col(int j) { return Block<Matrix, R, 1>(*this, 0, j, r_, 1); }
  Block<const Matrix, R, 1> col(int j) const { return Block<const Matrix, R, 1>(*this, 0, j, r_, 1); }
  static Matrix UnitZ() {
    Matrix m;
    m(2) = T(1);
    return m;
  }
  Matrix& operator=(const Quaternion<T>& q);  // 3x3 = rotation matrix of q (Eigen RotationBase assignment)
  LLT<Matrix> llt() const;

  using MatrixBase<Matrix>::transpose;
  TransposeRef<Matrix> transpose() { return TransposeRef<Matrix>(*this); }
  DiagonalWrapper<Matrix> asDiagonal() const;
  ArrayWrapper<Matrix> array() const;
  LDLT<Matrix> ldlt() const;

 private:
  int r_, c_;
  std::vector<T> d_;
};

// ---- arithmetic (eager) ----
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> m(a.rows(), a.cols());
  assert(a.rows() == b.rows() && a.cols() == b.cols());
  for (int j = 0; j < a.cols(); j++)
    for (int i = 0; i < a.rows(); i++) m.ref(i, j) = a.coeff(i, j) + b.coeff(i, j);
  return m;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> m(a.rows(), a.cols());
  assert(a.rows() == b.rows() && a.cols() == b.cols());
  for (int j = 0; j < a.cols(); j++)
    for (int i = 0; i < a.rows(); i++) m.ref(i, j) = a.coeff(i, j) - b.coeff(i, j);
  return m;
}
template <class A>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator-(const MatrixBase<A>& a) {
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> m(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); j++)
    for (int i = 0; i < a.rows(); i++) m.ref(i, j) = -a.coeff(i, j);
  return m;
}
template <class A>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator*(const MatrixBase<A>& a, typename traits<A>::Scalar s) {
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> m(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); j++)
    for (int i = 0; i < a.rows(); i++) m.ref(i, j) = a.coeff(i, j) * s;
  return m;
}
template <class A>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator*(typename traits<A>::Scalar s, const MatrixBase<A>& a) {
  return a * s;
}
template <class A>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> operator/(const MatrixBase<A>& a, typename traits<A>::Scalar s) {
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Cols> m(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); j++)
    for (int i = 0; i < a.rows(); i++) m.ref(i, j) = a.coeff(i, j) / s;
  return m;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  assert(a.cols() == b.rows());
  const typename MatrixBase<A>::PlainObject av = a.eval();
  const typename MatrixBase<B>::PlainObject bv = b.eval();
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> m(a.rows(), b.cols());
  for (int j = 0; j < bv.cols(); j++)
    for (int i = 0; i < av.rows(); i++) {
      typename traits<A>::Scalar s = 0;
      for (int k = 0; k < av.cols(); k++) s += av.coeff(i, k) * bv.coeff(k, j);
      m.ref(i, j) = s;
    }
  return m;
}
// scalar (op) 1x1 matrix, e.g. `-log(det) - r.transpose() * S.solve(r)` (rbis.cpp:142)
inline double operator-(double s, const Matrix<double, 1, 1>& m) { return s - m.coeff(0, 0); }
inline double operator+(double s, const Matrix<double, 1, 1>& m) { return s + m.coeff(0, 0); }
template <class A>
std::ostream& operator<<(std::ostream& os, const MatrixBase<A>& a) {
  for (int i = 0; i < a.rows(); i++) {
    for (int j = 0; j < a.cols(); j++) os << (j ? " " : "") << a.coeff(i, j);
    if (i + 1 < a.rows()) os << "\n";
  }
  return os;
}

// ---- comma initialiser for vectors: `x << a, b;` (estimate_tools/iir_notch.cpp:11,53) ----
template <class M>
struct CommaInitializer {
  M& m;
  int k;
  CommaInitializer(M& m_, double v) : m(m_), k(0) { m.data()[k++] = v; }
  CommaInitializer& operator,(double v) { m.data()[k++] = v; return *this; }
};
template <class T, int R>
CommaInitializer<Matrix<T, R, 1> > operator<<(Matrix<T, R, 1>& m, double v) { return CommaInitializer<Matrix<T, R, 1> >(m, v); }

// ---- asDiagonal ----
template <class V>
class DiagonalWrapper {
 public:
  explicit DiagonalWrapper(const V& v) : v_(v) {}
  const V& vec() const { return v_; }

 private:
  V v_;
};
template <typename T, int R, int C>
DiagonalWrapper<Matrix<T, R, C> > Matrix<T, R, C>::asDiagonal() const { return DiagonalWrapper<Matrix>(*this); }
template <typename T, int R, int C>
template <class V>
Matrix<T, R, C>::Matrix(const DiagonalWrapper<V>& dw) : r_(dw.vec().size()), c_(dw.vec().size()), d_((size_t)r_ * c_, T(0)) {
  for (int i = 0; i < r_; i++) ref(i, i) = dw.vec()(i);
}
template <class A, class V>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Rows == Dynamic ? Dynamic : traits<V>::Rows * traits<V>::Cols>
operator*(const MatrixBase<A>& a, const DiagonalWrapper<V>& d) {
  Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<A>::Rows == Dynamic ? Dynamic : traits<V>::Rows * traits<V>::Cols> m(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); j++)
    for (int i = 0; i < a.rows(); i++) m.ref(i, j) = a.coeff(i, j) * d.vec()(j);
  return m;
}

// ---- .array() < x  ->  .any() ----
struct BoolArray {
  std::vector<bool> b;
  bool any() const { for (bool v : b) if (v) return true; return false; }
  bool all() const { for (bool v : b) if (!v) return false; return true; }
};
template <class V>
class ArrayWrapper {
 public:
  explicit ArrayWrapper(const V& v) : v_(v) {}
  BoolArray operator!=(const ArrayWrapper& o) const {
    BoolArray r;
    for (int j = 0; j < v_.cols(); j++)
      for (int i = 0; i < v_.rows(); i++) r.b.push_back(v_.coeff(i, j) != o.v_.coeff(i, j));
    return r;
  }
  BoolArray operator<(typename traits<V>::Scalar x) const {
    BoolArray r;
    for (int j = 0; j < v_.cols(); j++)
      for (int i = 0; i < v_.rows(); i++) r.b.push_back(v_.coeff(i, j) < x);
    return r;
  }

 private:
  V v_;
};
template <typename T, int R, int C>
ArrayWrapper<Matrix<T, R, C> > Matrix<T, R, C>::array() const { return ArrayWrapper<Matrix>(*this); }

// ---- Map ----
template <class M> class Map;
template <class M>
struct traits<Map<M> > {
  typedef typename traits<typename std::remove_const<M>::type>::Scalar Scalar;
  enum { Rows = traits<typename std::remove_const<M>::type>::Rows, Cols = traits<typename std::remove_const<M>::type>::Cols };
};
template <class M>
class Map : public WritableBase<Map<M> > {
 public:
  typedef typename traits<Map>::Scalar Scalar;
  typedef typename std::conditional<std::is_const<M>::value, const Scalar*, Scalar*>::type Ptr;
  explicit Map(Ptr p) : p_(p), r_(traits<Map>::Rows), c_(traits<Map>::Cols) {}
  Map(Ptr p, int n) : p_(p), r_(traits<Map>::Rows == Dynamic ? n : (int)traits<Map>::Rows), c_(traits<Map>::Cols == Dynamic ? 1 : (int)traits<Map>::Cols) {}
  Map(Ptr p, int r, int c) : p_(p), r_(r), c_(c) {}
  int rows_() const { return r_; }
  int cols_() const { return c_; }
  Scalar coeff_(int i, int j) const { return p_[(size_t)i + (size_t)r_ * j]; }
  Scalar& ref(int i, int j) { return const_cast<Scalar*>(p_)[(size_t)i + (size_t)r_ * j]; }
  void resize_like(int, int) {}
  template <class O>
  Map& operator=(const MatrixBase<O>& o) { return this->assign(o); }

 private:
  Ptr p_;
  int r_, c_;
};

// ---- LDLT (no pivoting; the matrices of the reference are symmetric positive definite) ----
template <class M>
class LDLT {
 public:
  typedef typename traits<M>::Scalar Scalar;
  LDLT() : n_(0) {}
  template <class O>
  explicit LDLT(const MatrixBase<O>& a) { compute(a); }
  template <class O>
  void compute(const MatrixBase<O>& a) {
    n_ = a.rows();
    L_.assign((size_t)n_ * n_, Scalar(0));
    D_.assign((size_t)n_, Scalar(0));
    for (int k = 0; k < n_; k++) {
      Scalar d = a.coeff(k, k);
      for (int p = 0; p < k; p++) d -= l(k, p) * l(k, p) * D_[(size_t)p];
      D_[(size_t)k] = d;
      l(k, k) = 1;
      for (int i = k + 1; i < n_; i++) {
        Scalar v = a.coeff(i, k);
        for (int p = 0; p < k; p++) v -= l(i, p) * l(k, p) * D_[(size_t)p];
        l(i, k) = v / d;
      }
    }
  }
  template <class B>
  Matrix<Scalar, traits<B>::Rows, traits<B>::Cols> solve(const MatrixBase<B>& b) const {
    Matrix<Scalar, traits<B>::Rows, traits<B>::Cols> x = b.eval();
    for (int c = 0; c < x.cols(); c++) {
      for (int i = 0; i < n_; i++)
        for (int k = 0; k < i; k++) x.ref(i, c) -= lc(i, k) * x.coeff(k, c);
      for (int i = 0; i < n_; i++) x.ref(i, c) /= D_[(size_t)i];
      for (int i = n_ - 1; i >= 0; i--)
        for (int k = i + 1; k < n_; k++) x.ref(i, c) -= lc(k, i) * x.coeff(k, c);
    }
    return x;
  }
  Scalar determinant() const {
    Scalar d = 1;
    for (int i = 0; i < n_; i++) d *= D_[(size_t)i];
    return d;
  }

 private:
  Scalar& l(int i, int j) { return L_[(size_t)i + (size_t)n_ * j]; }
  Scalar lc(int i, int j) const { return L_[(size_t)i + (size_t)n_ * j]; }
  int n_;
  std::vector<Scalar> L_, D_;
};
template <typename T, int R, int C>
LDLT<Matrix<T, R, C> > Matrix<T, R, C>::ldlt() const { return LDLT<Matrix>(*this); }

// determinant by Gaussian elimination with partial pivoting (as Eigen's PartialPivLU for general sizes)
template <class D>
typename MatrixBase<D>::Scalar MatrixBase<D>::determinant() const {
  const int n = rows();
  assert(n == cols());
  std::vector<Scalar> a((size_t)n * n);
  for (int j = 0; j < n; j++)
    for (int i = 0; i < n; i++) a[(size_t)i + (size_t)n * j] = coeff(i, j);
  Scalar det = 1;
  for (int k = 0; k < n; k++) {
    int piv = k;
    for (int i = k + 1; i < n; i++)
      if (std::fabs(a[(size_t)i + (size_t)n * k]) > std::fabs(a[(size_t)piv + (size_t)n * k])) piv = i;
    if (a[(size_t)piv + (size_t)n * k] == 0) return 0;
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(a[(size_t)k + (size_t)n * j], a[(size_t)piv + (size_t)n * j]);
      det = -det;
    }
    det *= a[(size_t)k + (size_t)n * k];
    for (int i = k + 1; i < n; i++) {
      const Scalar f = a[(size_t)i + (size_t)n * k] / a[(size_t)k + (size_t)n * k];
      for (int j = k + 1; j < n; j++) a[(size_t)i + (size_t)n * j] -= f * a[(size_t)k + (size_t)n * j];
    }
  }
  return det;
}

// ---- quaternion (coefficients w, x, y, z) ----
template <typename T>
class Quaternion {
 public:
  Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
  Quaternion(T w, T x, T y, T z) : w_(w), x_(x), y_(y), z_(z) {}
  // from a rotation matrix: Eigen's quaternionbase_assign_impl<Matrix3> (trace branch, else the largest diagonal element)
  explicit Quaternion(const Matrix<T, 3, 3>& m) {
    T t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > T(0)) {
      t = std::sqrt(t + T(1));
      w_ = T(0.5) * t;
      t = T(0.5) / t;
      x_ = (m(2, 1) - m(1, 2)) * t; y_ = (m(0, 2) - m(2, 0)) * t; z_ = (m(1, 0) - m(0, 1)) * t;
    } else {
      int i = 0;
      if (m(1, 1) > m(0, 0)) i = 1;
      if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + T(1));
      T v[3];
      v[i] = T(0.5) * t;
      t = T(0.5) / t;
      w_ = (m(k, j) - m(j, k)) * t;
      v[j] = (m(j, i) + m(i, j)) * t;
      v[k] = (m(k, i) + m(i, k)) * t;
      x_ = v[0]; y_ = v[1]; z_ = v[2];
    }
  }
  static Quaternion Identity() { return Quaternion(1, 0, 0, 0); }
  T w() const { return w_; }
  T x() const { return x_; }
  T y() const { return y_; }
  T z() const { return z_; }
  T& w() { return w_; }
  T& x() { return x_; }
  T& y() { return y_; }
  T& z() { return z_; }
  Matrix<T, 3, 1> vec() const { return Matrix<T, 3, 1>(x_, y_, z_); }
  T squaredNorm() const { return w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_; }
  T norm() const { return std::sqrt(squaredNorm()); }
  void normalize() { const T n = norm(); w_ /= n; x_ /= n; y_ /= n; z_ /= n; }
  Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
  Quaternion inverse() const {  // Eigen: conjugate / squaredNorm, zero quaternion if the norm is zero
    const T n2 = squaredNorm();
    if (n2 > T(0)) return Quaternion(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2);
    return Quaternion(0, 0, 0, 0);
  }
  Quaternion operator*(const Quaternion& b) const {
    return Quaternion(w_ * b.w_ - x_ * b.x_ - y_ * b.y_ - z_ * b.z_, w_ * b.x_ + x_ * b.w_ + y_ * b.z_ - z_ * b.y_,
                      w_ * b.y_ + y_ * b.w_ + z_ * b.x_ - x_ * b.z_, w_ * b.z_ + z_ * b.w_ + x_ * b.y_ - y_ * b.x_);
  }
  template <class V>
  Matrix<T, 3, 1> operator*(const MatrixBase<V>& v) const {  // Eigen _transformVector
    const Matrix<T, 3, 1> u = vec();
    Matrix<T, 3, 1> uv = u.cross(v);
    uv += uv;
    return Matrix<T, 3, 1>(v(0), v(1), v(2)) + w_ * uv + u.cross(uv);
  }
  Matrix<T, 3, 3> toRotationMatrix() const {
    Matrix<T, 3, 3> res;
    const T tx = T(2) * x_, ty = T(2) * y_, tz = T(2) * z_;
    const T twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_, tyy = ty * y_, tyz = tz * y_,
            tzz = tz * z_;
    res(0, 0) = T(1) - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
    res(1, 0) = txy + twz; res(1, 1) = T(1) - (txx + tzz); res(1, 2) = tyz - twx;
    res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = T(1) - (txx + tyy);
    return res;
  }

 private:
  T w_, x_, y_, z_;
};

template <typename T, int R, int C>
Matrix<T, R, C>& Matrix<T, R, C>::operator=(const Quaternion<T>& q) {
  return this->assign(q.toRotationMatrix());
}

// ---- LLT (Cholesky, lower factor) ----
template <class M>
class LLT {
 public:
  typedef typename traits<M>::Scalar Scalar;
  explicit LLT(const M& a) : L_(a.rows(), a.cols()) {
    const int n = a.rows();
    for (int j = 0; j < n; j++) {
      Scalar d = a.coeff(j, j);
      for (int k = 0; k < j; k++) d -= L_.coeff(j, k) * L_.coeff(j, k);
      d = std::sqrt(d);
      L_.ref(j, j) = d;
      for (int i = j + 1; i < n; i++) {
        Scalar v = a.coeff(i, j);
        for (int k = 0; k < j; k++) v -= L_.coeff(i, k) * L_.coeff(j, k);
        L_.ref(i, j) = v / d;
      }
    }
  }
  const M& matrixL() const { return L_; }

 private:
  M L_;
};
template <typename T, int R, int C>
LLT<Matrix<T, R, C> > Matrix<T, R, C>::llt() const { return LLT<Matrix>(*this); }

// ---- Transform<T, 3, Isometry>: rotation + translation, the few members pronto_conversions_* use ----
template <typename T>
class Isometry3 {
 public:
  Isometry3() { setIdentity(); }
  void setIdentity() { R_.setIdentity(); t_ = Matrix<T, 3, 1>(); }
  Matrix<T, 3, 1>& translation() { return t_; }
  const Matrix<T, 3, 1>& translation() const { return t_; }
  Matrix<T, 3, 3> rotation() const { return R_; }
  Matrix<T, 3, 3>& linear() { return R_; }
  const Matrix<T, 3, 3>& linear() const { return R_; }
  Isometry3& rotate(const Quaternion<T>& q) {  // applies the rotation on the right of the current transform
    R_ = R_ * q.toRotationMatrix();
    return *this;
  }

 private:
  Matrix<T, 3, 3> R_;
  Matrix<T, 3, 1> t_;
};
typedef Isometry3<double> Isometry3d;
typedef Isometry3<float> Isometry3f;
#ifndef EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#endif

template <class T> using aligned_allocator = std::allocator<T>;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<int, Dynamic, 1> VectorXi;
typedef Matrix<int, 3, 1> Vector3i;
typedef Quaternion<double> Quaterniond;

}  // namespace Eigen
