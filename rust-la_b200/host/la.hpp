// la.hpp -- C++ host-side mirror of the rust-la public API for the dense hot path, above the C ABI (include/la_cabi.h).
//
// The reference is a compiled (Rust) crate and the image has no Rust toolchain, so this header is the compiled-language
// stand-in for `rust-la_b200/rust/src/{matrix,lu}.rs`: same type names, method names, argument meaning and error
// behaviour, so tests written against it read like the reference's own tests.
//
//   la::Matrix<T>            reference src/matrix/mod.rs:26-30; new :207-211, rows :256, cols :260, get_data :264,
//                            get :557-560, id :416-426, operator* :957-998, mmul (mmatrix.rs:82-98),
//                            det/solve/inverse/is_singular/is_non_singular :1025-1047
//   la::LUDecomposition<T>   reference src/decomp/lu.rs:95-278
//   la::m<T>({{..},{..}})    the m! macro, src/macros.rs:39-42
//   la::Panic                the reference `assert!`s (panics); thrown BEFORE any FFI call
//   std::optional            Option<Matrix<T>> (None on numerical singularity, lu.rs:241-243)
#pragma once
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/la_cabi.h"

namespace la {

struct Panic : std::logic_error {
  explicit Panic(const std::string& what) : std::logic_error("assertion failed: " + what) {}
};
struct LaError : std::runtime_error {
  int status;
  LaError(int st, const char* text) : std::runtime_error(std::string("la_b200: ") + text), status(st) {}
};
inline void check(int status) {
  if (status != LA_OK) throw LaError(status, la_last_error());
}
#define LA_ASSERT(cond) \
  do {                  \
    if (!(cond)) throw ::la::Panic(#cond); \
  } while (0)

// typed dispatch onto the C ABI
template <typename T> struct Abi;
template <> struct Abi<double> {
  static int gemm_host(const double* a, const double* b, double* c, size_t m, size_t k, size_t n) { return la_gemm_f64_host(a, b, c, m, k, n); }
  static int factor(la_buf* lu, size_t m, size_t n, uint64_t* piv, int* sign) { return la_lu_factor_f64(lu, m, n, piv, sign); }
  static int nonsingular(const la_buf* lu, size_t n, int* out) { return la_lu_is_nonsingular_f64(lu, n, out); }
  static int det(const la_buf* lu, size_t n, int pos, double* out) { return la_lu_det_f64(lu, n, pos, out); }
  static int solve(const la_buf* lu, size_t m, size_t n, const uint64_t* piv, const la_buf* b, size_t nx, la_buf* x) { return la_lu_solve_f64(lu, m, n, piv, b, nx, x); }
  static int chol_factor(la_buf* a, size_t n, int* ok) { return la_chol_factor_f64(a, n, ok); }
  static int chol_solve(const la_buf* l, size_t n, const la_buf* b, size_t nx, la_buf* x) { return la_chol_solve_f64(l, n, b, nx, x); }
};
template <> struct Abi<float> {
  static int gemm_host(const float* a, const float* b, float* c, size_t m, size_t k, size_t n) { return la_gemm_f32_host(a, b, c, m, k, n); }
  static int factor(la_buf* lu, size_t m, size_t n, uint64_t* piv, int* sign) { return la_lu_factor_f32(lu, m, n, piv, sign); }
  static int nonsingular(const la_buf* lu, size_t n, int* out) { return la_lu_is_nonsingular_f32(lu, n, out); }
  static int det(const la_buf* lu, size_t n, int pos, float* out) { return la_lu_det_f32(lu, n, pos, out); }
  static int solve(const la_buf* lu, size_t m, size_t n, const uint64_t* piv, const la_buf* b, size_t nx, la_buf* x) { return la_lu_solve_f32(lu, m, n, piv, b, nx, x); }
  static int chol_factor(la_buf* a, size_t n, int* ok) { return la_chol_factor_f32(a, n, ok); }
  static int chol_solve(const la_buf* l, size_t n, const la_buf* b, size_t nx, la_buf* x) { return la_chol_solve_f32(l, n, b, nx, x); }
};
template <> struct Abi<int64_t> {
  static int gemm_host(const int64_t* a, const int64_t* b, int64_t* c, size_t m, size_t k, size_t n) { return la_gemm_i64_host(a, b, c, m, k, n); }
};

template <typename T> class LUDecomposition;

template <typename T>
class Matrix {
 public:
  // Matrix::new, mod.rs:207-211
  Matrix(size_t no_rows, size_t no_cols, std::vector<T> data) : no_rows_(no_rows), data_(std::move(data)) {
    LA_ASSERT(no_rows * no_cols == data_.size());
    LA_ASSERT(no_rows > 0 && no_cols > 0);
  }
  static Matrix id(size_t m, size_t n) {  // mod.rs:416-426
    std::vector<T> d(m * n, T(0));
    for (size_t i = 0; i < (m < n ? m : n); ++i) d[i * n + i] = T(1);
    return Matrix(m, n, std::move(d));
  }
  size_t rows() const { return no_rows_; }
  size_t cols() const { return data_.size() / no_rows_; }
  const std::vector<T>& get_data() const { return data_; }
  std::vector<T>& get_mut_data() { return data_; }
  T get(size_t row, size_t col) const {
    LA_ASSERT(row < no_rows_ && col < cols());
    return data_[row * cols() + col];
  }
  bool operator==(const Matrix& o) const { return no_rows_ == o.no_rows_ && data_ == o.data_; }  // derive(PartialEq)
  bool approx_eq(const Matrix& o) const {  // mod.rs:1141-1147, ApproxEq absolute 1e-6
    if (rows() != o.rows() || cols() != o.cols()) return false;
    for (size_t i = 0; i < data_.size(); ++i) {
      T d = data_[i] - o.data_[i];
      if (!((d < 0 ? -d : d) < T(1.0e-6))) return false;
    }
    return true;
  }
  Matrix t() const {
    std::vector<T> d(data_.size());
    for (size_t r = 0; r < rows(); ++r)
      for (size_t c = 0; c < cols(); ++c) d[c * rows() + r] = data_[r * cols() + c];
    return Matrix(cols(), rows(), std::move(d));
  }

  // impl Mul, mod.rs:957-980: shape assert first (panics before the FFI call), output is a fresh "dirty" buffer
  Matrix operator*(const Matrix& m) const {
    LA_ASSERT(cols() == m.no_rows_);
    std::vector<T> d(no_rows_ * m.cols());
    check(Abi<T>::gemm_host(data_.data(), m.data_.data(), d.data(), no_rows_, cols(), m.cols()));
    return Matrix(no_rows_, m.cols(), std::move(d));
  }
  // Matrix::mmul, mmatrix.rs:82-98
  Matrix& mmul(const Matrix& m, Matrix& dst) const {
    LA_ASSERT(cols() == m.no_rows_);
    LA_ASSERT(dst.rows() == no_rows_);
    LA_ASSERT(dst.cols() == m.cols());
    check(Abi<T>::gemm_host(data_.data(), m.data_.data(), dst.data_.data(), no_rows_, cols(), m.cols()));
    return dst;
  }

  // LU callers, mod.rs:1025-1047 (each re-factorises, like the reference)
  T det() const {
    LA_ASSERT(cols() == no_rows_);
    return LUDecomposition<T>(*this).det();
  }
  std::optional<Matrix> solve(const Matrix& b) const { return LUDecomposition<T>(*this).solve(b); }
  std::optional<Matrix> inverse() const {
    LA_ASSERT(no_rows_ == cols());
    return LUDecomposition<T>(*this).solve(Matrix::id(no_rows_, no_rows_));
  }
  bool is_singular() const { return !is_non_singular(); }
  bool is_non_singular() const {
    LA_ASSERT(no_rows_ == cols());
    return LUDecomposition<T>(*this).is_non_singular();
  }

 private:
  size_t no_rows_;
  std::vector<T> data_;
};

// the m! macro: la::m<double>({{1, 2}, {3, 4}})
template <typename T>
Matrix<T> m(std::initializer_list<std::initializer_list<T>> rows) {
  std::vector<T> d;
  size_t nr = 0;
  for (auto& r : rows) {
    ++nr;
    for (auto v : r) d.push_back(v);
  }
  const size_t nc = nr ? d.size() / nr : 0;  // before the move: argument evaluation order is unspecified
  return Matrix<T>(nr, nc, std::move(d));
}

namespace detail {
struct DevBuf {
  la_buf* h = nullptr;
  explicit DevBuf(size_t bytes) { check(la_buf_alloc(bytes, 0, &h)); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : h(o.h) { o.h = nullptr; }
  ~DevBuf() {
    if (h) la_buf_free(h);
  }
};
}  // namespace detail

// LUDecomposition<T>, lu.rs:95-101.  The packed factors stay resident in HBM; host copies are made on demand.
template <typename T>
class LUDecomposition {
 public:
  explicit LUDecomposition(const Matrix<T>& a)  // LUDecomposition::new, lu.rs:104-168
      : m_(a.rows()), n_(a.cols()), lu_(a.rows() * a.cols() * sizeof(T)), piv_(a.rows()) {
    check(la_buf_upload(lu_.h, 0, a.get_data().data(), m_ * n_ * sizeof(T)));  // ludata = a.get_data().clone()
    int sign = 1;
    check(Abi<T>::factor(lu_.h, m_, n_, piv_.data(), &sign));
    pospivsign_ = sign != 0;
  }
  bool is_singular() const { return !is_non_singular(); }
  bool is_non_singular() const {  // lu.rs:174-182 (out-of-bounds panic of the reference for m < n kept as a Panic)
    LA_ASSERT(m_ >= n_);
    int out = 0;
    check(Abi<T>::nonsingular(lu_.h, n_, &out));
    return out != 0;
  }
  Matrix<T> get_lu() const {
    std::vector<T> d(m_ * n_);
    check(la_buf_download(lu_.h, 0, d.data(), d.size() * sizeof(T)));
    return Matrix<T>(m_, n_, std::move(d));
  }
  Matrix<T> get_l() const {  // lu.rs:184-202
    auto lu = get_lu();
    size_t nn = m_ >= n_ ? n_ : m_;
    std::vector<T> l(m_ * nn);
    for (size_t i = 0; i < m_; ++i)
      for (size_t j = 0; j < nn; ++j) l[i * nn + j] = i > j ? lu.get_data()[i * n_ + j] : (i == j ? T(1) : T(0));
    return Matrix<T>(m_, nn, std::move(l));
  }
  Matrix<T> get_u() const {  // lu.rs:204-215
    auto lu = get_lu();
    size_t mm = m_ >= n_ ? n_ : m_;
    std::vector<T> u(mm * n_);
    for (size_t i = 0; i < mm; ++i)
      for (size_t j = 0; j < n_; ++j) u[i * n_ + j] = i <= j ? lu.get_data()[i * n_ + j] : T(0);
    return Matrix<T>(mm, n_, std::move(u));
  }
  Matrix<T> get_p() const {  // lu.rs:217-220
    size_t len = piv_.size();
    std::vector<T> p(len * len, T(0));
    for (size_t i = 0; i < len; ++i) p[i * len + piv_[i]] = T(1);
    return Matrix<T>(len, len, std::move(p));
  }
  const std::vector<uint64_t>& get_piv() const { return piv_; }
  bool pospivsign() const { return pospivsign_; }
  T det() const {  // lu.rs:224-232
    LA_ASSERT(m_ == n_);
    T out = T(0);
    check(Abi<T>::det(lu_.h, n_, pospivsign_ ? 1 : 0, &out));
    return out;
  }
  std::optional<Matrix<T>> solve(const Matrix<T>& b) const {  // lu.rs:237-278
    LA_ASSERT(b.rows() == m_);
    if (!is_non_singular()) return std::nullopt;
    size_t nx = b.cols(), bytes = m_ * nx * sizeof(T);
    detail::DevBuf db(bytes), dx(bytes);
    check(la_buf_upload(db.h, 0, b.get_data().data(), bytes));
    check(Abi<T>::solve(lu_.h, m_, n_, piv_.data(), db.h, nx, dx.h));
    std::vector<T> x(m_ * nx);
    check(la_buf_download(dx.h, 0, x.data(), bytes));
    return Matrix<T>(m_, nx, std::move(x));
  }

 private:
  size_t m_, n_;
  detail::DevBuf lu_;
  std::vector<uint64_t> piv_;
  bool pospivsign_ = true;
};

// CholeskyDecomposition<T>, cholesky.rs:52-144.  `make` is the reference's `new`: empty unless the matrix is square,
// exactly symmetric and positive definite.  L stays resident in HBM.
template <typename T>
class CholeskyDecomposition {
 public:
  static std::optional<CholeskyDecomposition<T>> make(const Matrix<T>& a) {
    if (a.rows() != a.cols()) return std::nullopt;  // cholesky.rs:57-59
    CholeskyDecomposition<T> c(a.rows());
    check(la_buf_upload(c.l_.h, 0, a.get_data().data(), c.n_ * c.n_ * sizeof(T)));
    int ok = 0;
    check(Abi<T>::chol_factor(c.l_.h, c.n_, &ok));
    if (!ok) return std::nullopt;  // not symmetric (:91-93) or not positive definite (:99-102)
    return std::optional<CholeskyDecomposition<T>>(std::move(c));
  }
  Matrix<T> get_l() const {
    std::vector<T> d(n_ * n_);
    check(la_buf_download(l_.h, 0, d.data(), d.size() * sizeof(T)));
    return Matrix<T>(n_, n_, std::move(d));
  }
  Matrix<T> solve(const Matrix<T>& b) const {  // cholesky.rs:116-144
    LA_ASSERT(b.rows() == n_);
    size_t nx = b.cols(), bytes = n_ * nx * sizeof(T);
    detail::DevBuf db(bytes), dx(bytes);
    check(la_buf_upload(db.h, 0, b.get_data().data(), bytes));
    check(Abi<T>::chol_solve(l_.h, n_, db.h, nx, dx.h));
    std::vector<T> x(n_ * nx);
    check(la_buf_download(dx.h, 0, x.data(), bytes));
    return Matrix<T>(n_, nx, std::move(x));
  }
  CholeskyDecomposition(CholeskyDecomposition&&) noexcept = default;

 private:
  explicit CholeskyDecomposition(size_t n) : n_(n), l_(n * n * sizeof(T)) {}
  size_t n_;
  detail::DevBuf l_;
};

}  // namespace la
