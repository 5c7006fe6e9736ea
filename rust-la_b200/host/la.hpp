// la.hpp -- C++ host-side mirror of the rust-la public API for the dense hot path, above the C ABI (include/la_cabi.h).
//
// The reference is a compiled (Rust) crate and the image has no Rust toolchain, so this header is the compiled-language
// stand-in for `rust-la_b200/rust/src/{matrix,lu}.rs`: same type names, method names, argument meaning and error
// behaviour, so tests written against it read like the reference's own tests.
//
//   la::Matrix<T>            reference src/matrix/mod.rs:26-30; new :207-211, rows :256, cols :260, get_data :264,
//                            get :557-560, id :416-426, operator* :957-998, mmul (mmatrix.rs:82-98),
//                            det/solve/inverse/is_singular/is_non_singular :1025-1047
//                            Device-backed: a Matrix owns a host Vec AND / OR a device buffer, each valid or
//                            stale; operators run on the device buffers and leave the result there, `get_data()`
//                            (mod.rs:264) downloads on first use, `get_mut_data()` (mmatrix.rs:11) invalidates the device
//                            copy -- so `a * b * c`, `a.t() * b`, `a.inverse()` chains keep their intermediates in HBM.
//   la::LUDecomposition<T>   reference src/decomp/lu.rs:95-278
//   la::QRDecomposition<T>   reference src/decomp/qr.rs:20-238; Matrix::pinverse mod.rs:1049-1057
//   la::m<T>({{..},{..}})    the m! macro, src/macros.rs:39-42
//   la::Panic                the reference `assert!`s (panics); thrown BEFORE any FFI call
//   std::optional            Option<Matrix<T>> (None on numerical singularity, lu.rs:241-243)
#pragma once
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <memory>
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
  static constexpr bool device_backed = true;
  static int gemm(const la_buf* a, const la_buf* b, la_buf* c, size_t m, size_t k, size_t n) { return la_gemm_f64(a, b, c, m, k, n); }
  static int transpose(const la_buf* s, la_buf* d, size_t r, size_t c) { return la_transpose_f64(s, d, r, c); }
  static int identity(la_buf* d, size_t n) { return la_identity_f64(d, n); }
  static int elementwise(int op, const la_buf* a, const la_buf* b, double s, la_buf* c, size_t n) { return la_elementwise_f64(op, a, b, s, c, n); }
  static int reduce(int kind, const la_buf* a, const la_buf* b, size_t n, double* out) { return la_reduce_f64(kind, a, b, n, out); }
  static int qr_factor(la_buf* qr, size_t m, size_t n, la_buf* rd, la_buf* tm) { return la_qr_factor_f64(qr, m, n, rd, tm); }
  static int qr_get_r(const la_buf* qr, size_t m, size_t n, const la_buf* rd, la_buf* r) { return la_qr_get_r_f64(qr, m, n, rd, r); }
  static int qr_get_q(const la_buf* qr, size_t m, size_t n, const la_buf* tm, la_buf* q) { return la_qr_get_q_f64(qr, m, n, tm, q); }
  static int qr_solve(const la_buf* qr, size_t m, size_t n, const la_buf* rd, const la_buf* b, size_t nx, la_buf* x) { return la_qr_solve_f64(qr, m, n, rd, b, nx, x); }
};
template <> struct Abi<float> {
  static int gemm_host(const float* a, const float* b, float* c, size_t m, size_t k, size_t n) { return la_gemm_f32_host(a, b, c, m, k, n); }
  static int factor(la_buf* lu, size_t m, size_t n, uint64_t* piv, int* sign) { return la_lu_factor_f32(lu, m, n, piv, sign); }
  static int nonsingular(const la_buf* lu, size_t n, int* out) { return la_lu_is_nonsingular_f32(lu, n, out); }
  static int det(const la_buf* lu, size_t n, int pos, float* out) { return la_lu_det_f32(lu, n, pos, out); }
  static int solve(const la_buf* lu, size_t m, size_t n, const uint64_t* piv, const la_buf* b, size_t nx, la_buf* x) { return la_lu_solve_f32(lu, m, n, piv, b, nx, x); }
  static int chol_factor(la_buf* a, size_t n, int* ok) { return la_chol_factor_f32(a, n, ok); }
  static int chol_solve(const la_buf* l, size_t n, const la_buf* b, size_t nx, la_buf* x) { return la_chol_solve_f32(l, n, b, nx, x); }
  static constexpr bool device_backed = true;
  static int gemm(const la_buf* a, const la_buf* b, la_buf* c, size_t m, size_t k, size_t n) { return la_gemm_f32(a, b, c, m, k, n); }
  static int transpose(const la_buf* s, la_buf* d, size_t r, size_t c) { return la_transpose_f32(s, d, r, c); }
  static int identity(la_buf* d, size_t n) { return la_identity_f32(d, n); }
  static int elementwise(int op, const la_buf* a, const la_buf* b, float s, la_buf* c, size_t n) { return la_elementwise_f32(op, a, b, s, c, n); }
  static int reduce(int kind, const la_buf* a, const la_buf* b, size_t n, float* out) { return la_reduce_f32(kind, a, b, n, out); }
  static int qr_factor(la_buf* qr, size_t m, size_t n, la_buf* rd, la_buf* tm) { return la_qr_factor_f32(qr, m, n, rd, tm); }
  static int qr_get_r(const la_buf* qr, size_t m, size_t n, const la_buf* rd, la_buf* r) { return la_qr_get_r_f32(qr, m, n, rd, r); }
  static int qr_get_q(const la_buf* qr, size_t m, size_t n, const la_buf* tm, la_buf* q) { return la_qr_get_q_f32(qr, m, n, tm, q); }
  static int qr_solve(const la_buf* qr, size_t m, size_t n, const la_buf* rd, const la_buf* b, size_t nx, la_buf* x) { return la_qr_solve_f32(qr, m, n, rd, b, nx, x); }
};
template <> struct Abi<int64_t> {
  static int gemm_host(const int64_t* a, const int64_t* b, int64_t* c, size_t m, size_t k, size_t n) { return la_gemm_i64_host(a, b, c, m, k, n); }
  static constexpr bool device_backed = false;  // integers: host-pointer entry point only
  static int gemm(const la_buf*, const la_buf*, la_buf*, size_t, size_t, size_t) { return LA_ERR_UNSUPPORTED; }
  static int transpose(const la_buf*, la_buf*, size_t, size_t) { return LA_ERR_UNSUPPORTED; }
  static int elementwise(int, const la_buf*, const la_buf*, int64_t, la_buf*, size_t) { return LA_ERR_UNSUPPORTED; }
};

template <typename T> class LUDecomposition;
template <typename T> class QRDecomposition;

namespace detail {
struct DevBuf {
  la_buf* h = nullptr;
  size_t bytes = 0;
  explicit DevBuf(size_t nbytes, int device = 0) : bytes(nbytes) { check(la_buf_alloc(nbytes, device, &h)); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : h(o.h), bytes(o.bytes) { o.h = nullptr; }
  ~DevBuf() {
    if (h) la_buf_free(h);
  }
};
}  // namespace detail

// Products below this many multiply-adds between two host-only operands go through the host-pointer entry point (one call,
// no buffer objects); anything larger, and anything that already has a device copy, stays on the device.
constexpr size_t kDeviceResidentMinWork = size_t(1) << 21;

template <typename T>
class Matrix {
 public:
  // Matrix::new, mod.rs:207-211
  Matrix(size_t no_rows, size_t no_cols, std::vector<T> data) : no_rows_(no_rows), no_cols_(no_cols), host_(std::move(data)) {
    LA_ASSERT(no_rows * no_cols == host_.size());
    LA_ASSERT(no_rows > 0 && no_cols > 0);
  }
  static Matrix id(size_t m, size_t n) {  // mod.rs:416-426
    std::vector<T> d(m * n, T(0));
    for (size_t i = 0; i < (m < n ? m : n); ++i) d[i * n + i] = T(1);
    return Matrix(m, n, std::move(d));
  }
  size_t rows() const { return no_rows_; }
  size_t cols() const { return no_cols_; }
  // get_data, mod.rs:264: the host Vec, downloaded from HBM the first time it is asked for after a device operation
  const std::vector<T>& get_data() const {
    sync_host();
    return host_;
  }
  // get_mut_data, mmatrix.rs:11: the caller may change the host Vec, so the device copy becomes stale
  std::vector<T>& get_mut_data() {
    sync_host();
    dev_valid_ = false;
    return host_;
  }
  T get(size_t row, size_t col) const {
    LA_ASSERT(row < no_rows_ && col < no_cols_);
    return get_data()[row * no_cols_ + col];
  }
  bool operator==(const Matrix& o) const {  // derive(PartialEq)
    return no_rows_ == o.no_rows_ && no_cols_ == o.no_cols_ && get_data() == o.get_data();
  }
  bool approx_eq(const Matrix& o) const {  // mod.rs:1141-1147, ApproxEq absolute 1e-6
    if (rows() != o.rows() || cols() != o.cols()) return false;
    const auto &x = get_data(), &y = o.get_data();
    for (size_t i = 0; i < x.size(); ++i) {
      T d = x[i] - y[i];
      if (!((d < 0 ? -d : d) < T(1.0e-6))) return false;
    }
    return true;
  }
  Matrix t() const {  // mod.rs:653-669
    if (Abi<T>::device_backed && dev_valid_) {
      Matrix out(no_cols_, no_rows_, Dirty{});
      check(Abi<T>::transpose(dev_->h, out.dev_->h, no_rows_, no_cols_));
      return out;
    }
    const auto& src = get_data();
    std::vector<T> d(src.size());
    for (size_t r = 0; r < no_rows_; ++r)
      for (size_t c = 0; c < no_cols_; ++c) d[c * no_rows_ + r] = src[r * no_cols_ + c];
    return Matrix(no_cols_, no_rows_, std::move(d));
  }

  // impl Mul, mod.rs:957-980: shape assert first (panics before the FFI call), output is a fresh "dirty" buffer
  Matrix operator*(const Matrix& m) const {
    LA_ASSERT(no_cols_ == m.no_rows_);
    if (use_device(m, no_rows_ * no_cols_ * m.no_cols_)) {
      Matrix out(no_rows_, m.no_cols_, Dirty{});
      check(Abi<T>::gemm(device()->h, m.device()->h, out.dev_->h, no_rows_, no_cols_, m.no_cols_));
      return out;
    }
    std::vector<T> d(no_rows_ * m.no_cols_);
    check(Abi<T>::gemm_host(get_data().data(), m.get_data().data(), d.data(), no_rows_, no_cols_, m.no_cols_));
    return Matrix(no_rows_, m.no_cols_, std::move(d));
  }
  // Matrix::mmul, mmatrix.rs:82-98
  Matrix& mmul(const Matrix& m, Matrix& dst) const {
    LA_ASSERT(no_cols_ == m.no_rows_);
    LA_ASSERT(dst.rows() == no_rows_);
    LA_ASSERT(dst.cols() == m.cols());
    if (use_device(m, no_rows_ * no_cols_ * m.no_cols_)) {
      if (!dst.dev_) dst.dev_ = std::make_shared<detail::DevBuf>(dst.no_rows_ * dst.no_cols_ * sizeof(T));
      check(Abi<T>::gemm(device()->h, m.device()->h, dst.dev_->h, no_rows_, no_cols_, m.no_cols_));
      dst.dev_valid_ = true;
      dst.host_valid_ = false;
      return dst;
    }
    check(Abi<T>::gemm_host(get_data().data(), m.get_data().data(), dst.get_mut_data().data(), no_rows_, no_cols_, m.no_cols_));
    return dst;
  }
  // elementwise operators (mod.rs:487-527, :853-929) on the device copies when either operand has one
  Matrix operator+(const Matrix& m) const { return elementwise(LA_EW_ADD, &m, T(0)); }
  Matrix operator-(const Matrix& m) const { return elementwise(LA_EW_SUB, &m, T(0)); }
  Matrix operator-() const { return elementwise(LA_EW_NEG, nullptr, T(0)); }
  Matrix scale(T factor) const { return elementwise(LA_EW_SCALE, nullptr, factor); }
  Matrix elem_mul(const Matrix& m) const { return elementwise(LA_EW_MUL, &m, T(0)); }
  Matrix elem_div(const Matrix& m) const { return elementwise(LA_EW_DIV, &m, T(0)); }
  T frobenius_norm() const {  // mod.rs:1094-1101
    T out = T(0);
    check(Abi<T>::reduce(LA_RED_SUMSQ, device()->h, nullptr, no_rows_ * no_cols_, &out));
    return out;
  }

  // LU callers, mod.rs:1025-1047 (each re-factorises, like the reference)
  T det() const {
    LA_ASSERT(cols() == no_rows_);
    return LUDecomposition<T>(*this).det();
  }
  std::optional<Matrix> solve(const Matrix& b) const { return LUDecomposition<T>(*this).solve(b); }
  std::optional<Matrix> inverse() const {
    LA_ASSERT(no_rows_ == cols());
    if (Abi<T>::device_backed && dev_valid_) {  // the identity is generated in HBM as well
      Matrix eye(no_rows_, no_rows_, Dirty{});
      check(Abi<T>::identity(eye.dev_->h, no_rows_));
      return LUDecomposition<T>(*this).solve(eye);
    }
    return LUDecomposition<T>(*this).solve(Matrix::id(no_rows_, no_rows_));
  }
  bool is_singular() const { return !is_non_singular(); }
  bool is_non_singular() const {
    LA_ASSERT(no_rows_ == cols());
    return LUDecomposition<T>(*this).is_non_singular();
  }
  // mod.rs:1049-1057: (r.t() * &r).inverse().unwrap() * &self.t(), r = QRDecomposition::new(self).get_r()
  Matrix pinverse() const {
    Matrix r = QRDecomposition<T>(*this).get_r();
    auto inv = (r.t() * r).inverse();
    if (!inv.has_value()) throw Panic("called `Option::unwrap()` on a `None` value");
    return *inv * t();
  }

  // ---- device backing (not part of the reference API) ----
  bool on_device() const { return dev_valid_; }          // a current copy lives in HBM
  bool host_materialised() const { return host_valid_; }  // a current copy lives in the host Vec
  // current device copy, uploading the host Vec if necessary (kept for later operations)
  const std::shared_ptr<detail::DevBuf>& device() const {
    if (!dev_valid_) {
      if (!dev_) dev_ = std::make_shared<detail::DevBuf>(no_rows_ * no_cols_ * sizeof(T));
      check(la_buf_upload(dev_->h, 0, host_.data(), host_.size() * sizeof(T)));
      dev_valid_ = true;
    }
    return dev_;
  }
  // result of a device operation: a fresh device buffer (alloc_dirty_vec, internalutil.rs:7-13), no host copy yet
  struct Dirty {};
  Matrix(size_t no_rows, size_t no_cols, Dirty)
      : no_rows_(no_rows), no_cols_(no_cols), host_valid_(false),
        dev_(std::make_shared<detail::DevBuf>(no_rows * no_cols * sizeof(T))), dev_valid_(true) {
    LA_ASSERT(no_rows > 0 && no_cols > 0);
  }

 private:
  bool use_device(const Matrix& other, size_t work) const {
    return Abi<T>::device_backed && (dev_valid_ || other.dev_valid_ || work >= kDeviceResidentMinWork);
  }
  void sync_host() const {
    if (!host_valid_) {
      host_.resize(no_rows_ * no_cols_);
      check(la_buf_download(dev_->h, 0, host_.data(), host_.size() * sizeof(T)));
      host_valid_ = true;
    }
  }
  Matrix elementwise(int op, const Matrix* m, T scalar) const {
    if (m) {
      LA_ASSERT(no_rows_ == m->no_rows_);
      LA_ASSERT(no_cols_ == m->no_cols_);
    }
    if (Abi<T>::device_backed) {
      Matrix out(no_rows_, no_cols_, Dirty{});
      check(Abi<T>::elementwise(op, device()->h, m ? m->device()->h : nullptr, scalar, out.dev_->h, no_rows_ * no_cols_));
      return out;
    }
    const auto& x = get_data();
    std::vector<T> d(x.size());
    for (size_t i = 0; i < x.size(); ++i) {
      const T y = m ? m->get_data()[i] : T(0);
      d[i] = op == LA_EW_ADD ? x[i] + y : op == LA_EW_SUB ? x[i] - y : op == LA_EW_MUL ? x[i] * y
             : op == LA_EW_DIV ? x[i] / y : op == LA_EW_SCALE ? scalar * x[i] : -x[i];
    }
    return Matrix(no_rows_, no_cols_, std::move(d));
  }

  size_t no_rows_, no_cols_;
  mutable std::vector<T> host_;
  mutable bool host_valid_ = true;
  mutable std::shared_ptr<detail::DevBuf> dev_;
  mutable bool dev_valid_ = false;
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


// LUDecomposition<T>, lu.rs:95-101.  The packed factors stay resident in HBM; host copies are made on demand.
template <typename T>
class LUDecomposition {
 public:
  explicit LUDecomposition(const Matrix<T>& a)  // LUDecomposition::new, lu.rs:104-168
      : m_(a.rows()), n_(a.cols()), lu_(a.rows() * a.cols() * sizeof(T)), piv_(a.rows()) {
    // ludata = a.get_data().clone(): device to device when `a` already lives in HBM
    if (a.on_device()) check(la_buf_copy(lu_.h, a.device()->h, m_ * n_ * sizeof(T)));
    else check(la_buf_upload(lu_.h, 0, a.get_data().data(), m_ * n_ * sizeof(T)));
    int sign = 1;
    check(Abi<T>::factor(lu_.h, m_, n_, piv_.data(), &sign));
    pospivsign_ = sign != 0;
  }
  // The same factorisation spread over several GPUs (la_lu_factor_f64_mg: 128-column blocks dealt round-robin, the panel
  // owner's block column copied to every device); the packed factors land on device 0, so solve / det / get_l work as usual.
  LUDecomposition(const Matrix<T>& a, const std::vector<int>& devices)
      : m_(a.rows()), n_(a.cols()), lu_(a.rows() * a.cols() * sizeof(T)), piv_(a.rows()) {
    static_assert(sizeof(T) == sizeof(double), "the multi-device factorisation is fp64 only");
    LA_ASSERT(m_ == n_ && !devices.empty());
    std::vector<T> packed(m_ * n_);
    int sign = 1;
    check(la_lu_factor_f64_mg(int(devices.size()), devices.data(), a.get_data().data(), packed.data(), n_, piv_.data(), &sign));
    check(la_buf_upload(lu_.h, 0, packed.data(), packed.size() * sizeof(T)));
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
    LA_ASSERT(m_ == n_);  // lu.rs:257-275 index X with n
    Matrix<T> x(m_, b.cols(), typename Matrix<T>::Dirty{});  // stays in HBM until somebody asks for get_data()
    check(Abi<T>::solve(lu_.h, m_, n_, piv_.data(), b.device()->h, b.cols(), x.device()->h));
    return x;
  }

 private:
  size_t m_, n_;
  detail::DevBuf lu_;
  std::vector<uint64_t> piv_;
  bool pospivsign_ = true;
};

// QRDecomposition<T>, qr.rs:20-238.  The packed factors, rdiag and the block factors of the compact-WY form stay in HBM.
template <typename T>
class QRDecomposition {
 public:
  explicit QRDecomposition(const Matrix<T>& a)  // QRDecomposition::new, qr.rs:26-43
      : m_(a.rows()), n_(a.cols()), qr_(a.rows() * a.cols() * sizeof(T)),
        rd_(((a.rows() < a.cols() ? a.rows() : a.cols()) + 1) * sizeof(T)), tm_(tmat_bytes(a.rows(), a.cols())),
        rdiag_(a.rows() < a.cols() ? a.rows() : a.cols()) {
    if (a.on_device()) check(la_buf_copy(qr_.h, a.device()->h, m_ * n_ * sizeof(T)));
    else check(la_buf_upload(qr_.h, 0, a.get_data().data(), m_ * n_ * sizeof(T)));
    check(Abi<T>::qr_factor(qr_.h, m_, n_, rd_.h, tm_.h));
    check(la_buf_download(rd_.h, 0, rdiag_.data(), rdiag_.size() * sizeof(T)));
  }
  bool is_full_rank() const {  // qr.rs:110-117: rdiag[j] for j < cols -- out of bounds when m < n
    for (size_t j = 0; j < n_; ++j) {
      LA_ASSERT(j < rdiag_.size());
      if (rdiag_[j] == T(0)) return false;
    }
    return true;
  }
  const std::vector<T>& get_rdiag() const { return rdiag_; }
  Matrix<T> get_qr() const {
    std::vector<T> d(m_ * n_);
    check(la_buf_download(qr_.h, 0, d.data(), d.size() * sizeof(T)));
    return Matrix<T>(m_, n_, std::move(d));
  }
  Matrix<T> get_h() const {  // qr.rs:121-135
    auto d = get_qr().get_data();
    for (size_t i = 0; i < m_; ++i)
      for (size_t j = i + 1; j < n_; ++j) d[i * n_ + j] = T(0);
    return Matrix<T>(m_, n_, std::move(d));
  }
  Matrix<T> get_r() const {  // qr.rs:138-152, left in HBM
    Matrix<T> r(m_, n_, typename Matrix<T>::Dirty{});
    check(Abi<T>::qr_get_r(qr_.h, m_, n_, rd_.h, r.device()->h));
    return r;
  }
  Matrix<T> get_q() const {  // qr.rs:155-194
    Matrix<T> q(m_, m_, typename Matrix<T>::Dirty{});
    check(Abi<T>::qr_get_q(qr_.h, m_, n_, tm_.h, q.device()->h));
    return q;
  }
  std::optional<Matrix<T>> solve(const Matrix<T>& b) const {  // qr.rs:199-238, quirks included
    LA_ASSERT(b.rows() == m_);
    if (!is_full_rank()) return std::nullopt;
    LA_ASSERT(n_ * b.cols() == m_ * b.cols());  // Matrix::new(cols, nx, <m * nx values>), :237
    Matrix<T> x(m_, b.cols(), typename Matrix<T>::Dirty{});
    check(Abi<T>::qr_solve(qr_.h, m_, n_, rd_.h, b.device()->h, b.cols(), x.device()->h));
    return x;
  }

 private:
  static size_t tmat_bytes(size_t m, size_t n) {
    size_t e = 0;
    check(la_qr_tmat_elems(m, n, 0, sizeof(T), &e));
    return e * sizeof(T);
  }
  size_t m_, n_;
  detail::DevBuf qr_, rd_, tm_;
  std::vector<T> rdiag_;
};

// CholeskyDecomposition<T>, cholesky.rs:52-144.  `make` is the reference's `new`: empty unless the matrix is square,
// exactly symmetric and positive definite.  L stays resident in HBM.
template <typename T>
class CholeskyDecomposition {
 public:
  static std::optional<CholeskyDecomposition<T>> make(const Matrix<T>& a) {
    if (a.rows() != a.cols()) return std::nullopt;  // cholesky.rs:57-59
    CholeskyDecomposition<T> c(a.rows());
    if (a.on_device()) check(la_buf_copy(c.l_.h, a.device()->h, c.n_ * c.n_ * sizeof(T)));
    else check(la_buf_upload(c.l_.h, 0, a.get_data().data(), c.n_ * c.n_ * sizeof(T)));
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
    Matrix<T> x(n_, b.cols(), typename Matrix<T>::Dirty{});
    check(Abi<T>::chol_solve(l_.h, n_, b.device()->h, b.cols(), x.device()->h));
    return x;
  }
  CholeskyDecomposition(CholeskyDecomposition&&) noexcept = default;

 private:
  explicit CholeskyDecomposition(size_t n) : n_(n), l_(n * n * sizeof(T)) {}
  size_t n_;
  detail::DevBuf l_;
};

}  // namespace la
