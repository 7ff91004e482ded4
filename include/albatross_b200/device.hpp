// albatross_b200 C++ trait layer — RAII ownership of the C-ABI objects and DeviceLDLT, the
// CovarianceRepresentation (reference concept: src/models/gp.hpp:42-45, instance
// Eigen::SerializableLDLT src/eigen/serializable_ldlt.hpp:22-215) whose factor lives in HBM.
//
// Nothing here computes on the host: every member is one call into include/albatross_b200.h.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <map>
#include <limits>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../albatross_b200.h"
#include "linalg_types.hpp"

namespace albatross_b200 {

// The reference has no error channel (ALBATROSS_ASSERT -> assert, src/details/error_handling.hpp:37-45).
// A non-zero C-ABI status aborts with the library's message, or throws when the translation unit
// defines ALBATROSS_B200_EXCEPTIONS.
struct device_error : public std::runtime_error {
  device_error(int status_, const std::string &what_) : std::runtime_error(what_), status(status_) {}
  int status;
};

inline void check_status(int status, const char *what) {
  if (status == AB_OK) {
    return;
  }
  std::string msg = std::string(what) + ": " + ab_last_error();
#ifdef ALBATROSS_B200_EXCEPTIONS
  throw device_error(status, msg);
#else
  std::fprintf(stderr, "albatross_b200: %s (status %d)\n", msg.c_str(), status);
  std::abort();
#endif
}

#define ALBATROSS_B200_CHECK(call) ::albatross_b200::check_status((call), #call)

// Policy for matrices that are not positive definite: the reference's diagonally pivoted
// LDLT never fails — on a semi-definite or indefinite K it returns zero / negative pivots, and everything
// computed from them (log-determinant, solves) is NaN or +-inf, which GenericTuner maps to an infinite
// objective (src/tune/tune.hpp:164-166, :204-206).  The device factorisation is unpivoted and reports
// AB_ERR_NOT_PD instead; the trait layer turns that status into the same observable result — NaN outputs
// and is_positive_definite() == false — rather than aborting, so a tuner step into a bad region continues.
// true: the call failed with AB_ERR_NOT_PD (the caller fills its outputs with NaN); any other non-zero
// status is fatal as before.
inline bool is_not_positive_definite(int status, const char *what) {
  if (status == AB_ERR_NOT_PD) {
    return true;
  }
  check_status(status, what);
  return false;
}
#define ALBATROSS_B200_NOT_PD(call) ::albatross_b200::is_not_positive_definite((call), #call)

inline double quiet_nan() { return std::numeric_limits<double>::quiet_NaN(); }
inline void fill_nan(double *p, std::size_t n) {
  for (std::size_t i = 0; p != nullptr && i < n; ++i) {
    p[i] = quiet_nan();
  }
}

// One handle per (process, device).  Models are copied liberally by the reference (FitModel stores
// the model by value, the tuner copies it per evaluation: src/core/fit_model.hpp:112,
// src/tune/tune.hpp:278), so the handle lives outside the models and is shared.
class Device {
public:
  explicit Device(int device_index) {
    ALBATROSS_B200_CHECK(ab_create(&h_, device_index));
  }
  Device(const Device &) = delete;
  Device &operator=(const Device &) = delete;
  ~Device() {
    if (h_) {
      ab_destroy(h_);
    }
  }
  ab_handle get() const { return h_; }

  ab_phase_times timings() const {
    ab_phase_times t;
    ALBATROSS_B200_CHECK(ab_timings(h_, &t));
    return t;
  }

  // The process-wide default: device ALBATROSS_B200_DEVICE (or LOCAL_RANK under torchrun / mpirun
  // style launchers, else 0).
  static std::shared_ptr<Device> &default_device() {
    static std::shared_ptr<Device> dev;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!dev) {
      int index = 0;
      if (const char *e = std::getenv("ALBATROSS_B200_DEVICE")) {
        index = std::atoi(e);
      } else if (const char *r = std::getenv("LOCAL_RANK")) {
        index = std::atoi(r);
      }
      dev = std::make_shared<Device>(index);
    }
    return dev;
  }

private:
  ab_handle h_ = nullptr;
};

inline void set_default_device(int device_index) {
  Device::default_device() = std::make_shared<Device>(device_index);
}

// Device-resident column-major matrix (opaque ab_matrix), move-only.
class DeviceMatrix {
public:
  DeviceMatrix() = default;
  DeviceMatrix(std::shared_ptr<Device> dev, ab_matrix m) : dev_(std::move(dev)), m_(m) {}
  explicit DeviceMatrix(const MatrixXd &host, std::shared_ptr<Device> dev = Device::default_device())
      : dev_(std::move(dev)) {
    ALBATROSS_B200_CHECK(ab_matrix_upload(dev_->get(), host.data(), host.rows(), host.cols(), &m_));
  }
  DeviceMatrix(DeviceMatrix &&o) noexcept : dev_(std::move(o.dev_)), m_(o.m_) { o.m_ = nullptr; }
  DeviceMatrix &operator=(DeviceMatrix &&o) noexcept {
    if (this != &o) {
      reset();
      dev_ = std::move(o.dev_);
      m_ = o.m_;
      o.m_ = nullptr;
    }
    return *this;
  }
  DeviceMatrix(const DeviceMatrix &) = delete;
  DeviceMatrix &operator=(const DeviceMatrix &) = delete;
  ~DeviceMatrix() { reset(); }

  void reset() {
    if (m_) {
      ab_matrix_free(dev_->get(), m_);
      m_ = nullptr;
    }
  }
  // Hands the matrix to a consumer that takes ownership (ab_potrf).
  ab_matrix release() {
    ab_matrix m = m_;
    m_ = nullptr;
    return m;
  }
  ab_matrix get() const { return m_; }
  const std::shared_ptr<Device> &device() const { return dev_; }

  Index rows() const {
    int64_t r = 0, c = 0;
    ALBATROSS_B200_CHECK(ab_matrix_dims(m_, &r, &c));
    return static_cast<Index>(r);
  }
  Index cols() const {
    int64_t r = 0, c = 0;
    ALBATROSS_B200_CHECK(ab_matrix_dims(m_, &r, &c));
    return static_cast<Index>(c);
  }
  MatrixXd to_host() const {
    MatrixXd out(rows(), cols());
    ALBATROSS_B200_CHECK(ab_matrix_download(dev_->get(), m_, out.data()));
    return out;
  }
  void add_diagonal(const VectorXd &d) {
    ALBATROSS_B200_CHECK(ab_matrix_add_diag(dev_->get(), m_, d.data()));
  }

private:
  std::shared_ptr<Device> dev_;
  ab_matrix m_ = nullptr;
};

using GroupIndices = std::vector<std::size_t>;

// CSR form of a GroupIndexer in std::map key order: what the C ABI takes wherever the reference
// iterates a std::map<Key, GroupIndices> (src/indexing/group_by.hpp:349-376).
struct GroupCSR {
  std::vector<int64_t> indices;
  std::vector<int64_t> offsets;
  int64_t ngroups() const { return static_cast<int64_t>(offsets.size()) - 1; }
};

template <typename GroupKey>
inline GroupCSR to_csr(const std::map<GroupKey, GroupIndices> &indexer) {
  GroupCSR csr;
  csr.offsets.push_back(0);
  for (const auto &pair : indexer) {
    for (std::size_t i : pair.second) {
      csr.indices.push_back(static_cast<int64_t>(i));
    }
    csr.offsets.push_back(static_cast<int64_t>(csr.indices.size()));
  }
  return csr;
}

inline GroupCSR to_csr(const std::vector<GroupIndices> &blocks) {
  GroupCSR csr;
  csr.offsets.push_back(0);
  for (const auto &block : blocks) {
    for (std::size_t i : block) {
      csr.indices.push_back(static_cast<int64_t>(i));
    }
    csr.offsets.push_back(static_cast<int64_t>(csr.indices.size()));
  }
  return csr;
}

/*
 * DeviceLDLT — drop-in for Eigen::SerializableLDLT as the CovarianceRepresentation of Fit<GPFit<..>>.
 * Copies share the device factor (shared_ptr), like the value-semantic host type but without a
 * 32 GiB copy.  Member names and meaning follow src/eigen/serializable_ldlt.hpp.
 */
class DeviceLDLT {
  struct Holder {
    std::shared_ptr<Device> dev;
    ab_factor f = nullptr;
    ~Holder() {
      if (f) {
        ab_factor_free(dev->get(), f);
      }
    }
  };

public:
  DeviceLDLT() = default;

  // SerializableLDLT(const MatrixXd&), serializable_ldlt.hpp:27: upload + in-place blocked factorisation.
  explicit DeviceLDLT(const MatrixXd &cov, std::shared_ptr<Device> dev = Device::default_device()) {
    DeviceMatrix m(cov, dev);
    adopt(std::move(m));
  }

  // From a Gram matrix that is already in HBM (what the device models do; the matrix is consumed).
  explicit DeviceLDLT(DeviceMatrix &&cov) { adopt(std::move(cov)); }

  // From a factor produced by a fused C-ABI call (ab_gp_fit).
  DeviceLDLT(std::shared_ptr<Device> dev, ab_factor f) : holder_(std::make_shared<Holder>()) {
    holder_->dev = std::move(dev);
    holder_->f = f;
  }

  ab_factor get() const { return holder_ ? holder_->f : nullptr; }
  const std::shared_ptr<Device> &device() const { return holder_->dev; }

  Index rows() const {
    int64_t n = 0;
    ALBATROSS_B200_CHECK(ab_factor_rows(get(), &n));
    return static_cast<Index>(n);
  }
  Index cols() const { return rows(); }

  // serializable_ldlt.hpp:36 (isPositive()): every pivot was > 0.
  bool is_positive_definite() const {
    if (get() == nullptr) {
      return false;
    }
    int64_t bad = -1;
    ALBATROSS_B200_CHECK(ab_factor_info(get(), &bad));
    return bad < 0;
  }

  // LDLT::solve, third_party/eigen/Eigen/src/Cholesky/LDLT.h:558-592.
  MatrixXd solve(const MatrixXd &rhs) const {
    MatrixXd out(rhs.rows(), rhs.cols());
    ALBATROSS_B200_CHECK(ab_factor_solve(h(), get(), rhs.data(), rhs.cols(), out.data()));
    return out;
  }
  VectorXd solve(const VectorXd &rhs) const {
    VectorXd out(rhs.size());
    ALBATROSS_B200_CHECK(ab_factor_solve(h(), get(), rhs.data(), 1, out.data()));
    return out;
  }

  // serializable_ldlt.hpp:100-109: D^-1/2 L^-1 P rhs.
  MatrixXd sqrt_solve(const MatrixXd &rhs) const {
    MatrixXd out(rhs.rows(), rhs.cols());
    ALBATROSS_B200_CHECK(ab_factor_sqrt_solve(h(), get(), rhs.data(), rhs.cols(), out.data()));
    return out;
  }

  // serializable_ldlt.hpp:91-94: P^T L D^1/2 rhs.
  MatrixXd sqrt_product(const MatrixXd &rhs) const {
    MatrixXd out(rhs.rows(), rhs.cols());
    ALBATROSS_B200_CHECK(ab_factor_sqrt_product(h(), get(), rhs.data(), rhs.cols(), out.data()));
    return out;
  }

  // serializable_ldlt.hpp:123-126: P^T L^-T D^-1/2 rhs.
  MatrixXd sqrt_transpose_solve(const MatrixXd &rhs) const {
    MatrixXd out(rhs.rows(), rhs.cols());
    ALBATROSS_B200_CHECK(ab_factor_sqrt_transpose_solve(h(), get(), rhs.data(), rhs.cols(), out.data()));
    return out;
  }

  // serializable_ldlt.hpp:111-115: D^1/2 (P^T L)^T as a dense upper-triangular matrix.
  MatrixXd sqrt_transpose() const {
    const Index n = rows();
    MatrixXd out(n, n);
    ALBATROSS_B200_CHECK(ab_factor_sqrt_transpose(h(), get(), out.data()));
    return out;
  }

  // serializable_ldlt.hpp:74-84 / :58-69, returned as the diagonal's vector (the reference returns an
  // Eigen::DiagonalMatrix; non-positive pivots, which it clamps to 0, cannot occur in a usable factor).
  VectorXd diagonal_sqrt() const {
    VectorXd out(rows());
    ALBATROSS_B200_CHECK(ab_factor_diagonal_sqrt(h(), get(), out.data()));
    return out;
  }
  VectorXd diagonal_sqrt_inverse() const {
    VectorXd out = diagonal_sqrt();
    for (Index i = 0; i < out.size(); ++i) {
      out[i] = out[i] > 0. ? 1. / out[i] : 0.;
    }
    return out;
  }

  // serializable_ldlt.hpp:128-135.
  double log_determinant() const {
    double out = 0.;
    ALBATROSS_B200_CHECK(ab_factor_logdet(h(), get(), &out));
    return out;
  }

  // src/evaluation/likelihood.hpp:38-47.
  double negative_log_likelihood(const VectorXd &deviation) const {
    double out = 0.;
    ALBATROSS_B200_CHECK(ab_factor_nll(h(), get(), deviation.data(), &out));
    return out;
  }

  // serializable_ldlt.hpp:181-199.
  VectorXd inverse_diagonal() const {
    VectorXd out(rows());
    ALBATROSS_B200_CHECK(ab_factor_inverse_diagonal(h(), get(), out.data()));
    return out;
  }

  // serializable_ldlt.hpp:137-175.
  std::vector<MatrixXd> inverse_blocks(const std::vector<GroupIndices> &blocks) const {
    const GroupCSR csr = to_csr(blocks);
    std::size_t total = 0;
    for (const auto &b : blocks) {
      total += b.size() * b.size();
    }
    std::vector<double> flat(total);
    ALBATROSS_B200_CHECK(ab_factor_inverse_blocks(h(), get(), csr.indices.data(), csr.offsets.data(),
                                                  csr.ngroups(), flat.data()));
    std::vector<MatrixXd> out;
    std::size_t at = 0;
    for (const auto &b : blocks) {
      const Index g = static_cast<Index>(b.size());
      MatrixXd m(g, g);
      for (Index k = 0; k < g * g; ++k) {
        m.data()[k] = flat[at + static_cast<std::size_t>(k)];
      }
      at += b.size() * b.size();
      out.push_back(std::move(m));
    }
    return out;
  }

  // Host materialisation in Eigen::SerializableLDLT's packed layout (strict lower = unit L,
  // diagonal = D, identity transpositions): src/cereal/serializable_ldlt.hpp:18-32.
  void export_packed(MatrixXd *LD, std::vector<int64_t> *transpositions) const {
    const Index n = rows();
    *LD = MatrixXd(n, n);
    transpositions->assign(static_cast<std::size_t>(n), 0);
    ALBATROSS_B200_CHECK(ab_factor_export_packed(h(), get(), LD->data(), transpositions->data()));
  }

  bool operator==(const DeviceLDLT &other) const { return get() == other.get(); }

private:
  ab_handle h() const { return holder_->dev->get(); }

  void adopt(DeviceMatrix &&m) {
    holder_ = std::make_shared<Holder>();
    holder_->dev = m.device();
    const int status = ab_potrf(holder_->dev->get(), m.release(), &holder_->f);
    if (status != AB_ERR_NOT_PD) { // a non-PD factor is kept so that is_positive_definite() can report it
      check_status(status, "ab_potrf");
    }
  }

  std::shared_ptr<Holder> holder_;
};

} // namespace albatross_b200
