// albatross_b200 C++ trait layer — block-diagonal factorisations and the QR concept of the sparse GP.
//
//   BlockDiagonal / BlockDiagonalLDLT   src/linalg/block_diagonal.hpp:24-91, :96-331: per-group blocks, each
//                                       factored by a DeviceLDLT (one blocked factorisation per block on the
//                                       device); solve / sqrt_solve / log_determinant apply block by block.
//   DenseQRImplementation / DeviceQR    the QRImplementation concept of src/models/sparse_gp.hpp:72-89 with the
//                                       helpers of src/linalg/qr_utils.hpp:18-53: compute(), get_R(), get_P(),
//                                       sqrt_solve(R, P, rhs).  The reference's QRType is Eigen's column-pivoted
//                                       Householder QR; here R comes from CholQR2 on the device (ab_qr_r), the
//                                       permutation is the identity and R's diagonal is positive.  R^T R = B^T B
//                                       in both, which is all the sparse model uses R for (sparse_gp.hpp:394-403,
//                                       :577-593).  SparseGaussianProcessRegression itself never builds B on the
//                                       host (sparse_gp.hpp of this layer -> ab_sparse_fit); these types serve
//                                       callers that compose the pieces themselves.
// Only the common member subset of linalg_types.hpp is used (works with Eigen and with the stand-ins).
#pragma once

#include <cassert>
#include <memory>
#include <numeric>

#include "device.hpp"

namespace albatross_b200 {

namespace details {
inline MatrixXd copy_rows(const MatrixXd &m, Index row0, Index rows) {
  MatrixXd out(rows, m.cols());
  for (Index j = 0; j < m.cols(); ++j) {
    for (Index i = 0; i < rows; ++i) {
      out(i, j) = m(row0 + i, j);
    }
  }
  return out;
}
inline void paste_rows(const MatrixXd &block, Index row0, MatrixXd *into) {
  for (Index j = 0; j < block.cols(); ++j) {
    for (Index i = 0; i < block.rows(); ++i) {
      (*into)(row0 + i, j) = block(i, j);
    }
  }
}
} // namespace details

struct BlockDiagonalLDLT { // block_diagonal.hpp:96-218
  std::vector<DeviceLDLT> blocks;

  Index rows() const {
    Index n = 0;
    for (const auto &b : blocks) {
      n += b.rows();
    }
    return n;
  }
  Index cols() const { return rows(); }

  MatrixXd solve(const MatrixXd &rhs) const { return apply(rhs, false); }           // :96-108
  MatrixXd sqrt_solve(const MatrixXd &rhs) const { return apply(rhs, true); }       // :110-122
  double log_determinant() const {                                                  // :180-186
    double out = 0.;
    for (const auto &b : blocks) {
      out += b.log_determinant();
    }
    return out;
  }
  bool is_positive_definite() const {
    for (const auto &b : blocks) {
      if (!b.is_positive_definite()) {
        return false;
      }
    }
    return true;
  }

private:
  MatrixXd apply(const MatrixXd &rhs, bool sqrt_only) const {
    assert(cols() == rhs.rows());
    MatrixXd out(rows(), rhs.cols());
    Index i = 0;
    for (const auto &b : blocks) {
      const MatrixXd chunk = details::copy_rows(rhs, i, b.rows());
      details::paste_rows(sqrt_only ? b.sqrt_solve(chunk) : b.solve(chunk), i, &out);
      i += b.rows();
    }
    return out;
  }
};

struct BlockDiagonal { // block_diagonal.hpp:24-91
  std::vector<MatrixXd> blocks;

  Index rows() const {
    Index n = 0;
    for (const auto &b : blocks) {
      n += b.rows();
    }
    return n;
  }
  Index cols() const { return rows(); }

  BlockDiagonalLDLT ldlt(std::shared_ptr<Device> dev = Device::default_device()) const { // :310-317
    BlockDiagonalLDLT out;
    for (const auto &b : blocks) {
      out.blocks.emplace_back(b, dev);
    }
    return out;
  }
  BlockDiagonal operator-(const BlockDiagonal &rhs) const { // :278-290
    assert(blocks.size() == rhs.blocks.size());
    BlockDiagonal out;
    for (std::size_t k = 0; k < blocks.size(); ++k) {
      MatrixXd d(blocks[k].rows(), blocks[k].cols());
      for (Index j = 0; j < d.cols(); ++j) {
        for (Index i = 0; i < d.rows(); ++i) {
          d(i, j) = blocks[k](i, j) - rhs.blocks[k](i, j);
        }
      }
      out.blocks.push_back(std::move(d));
    }
    return out;
  }
  MatrixXd toDense() const { // :319-331
    const Index n = rows();
    MatrixXd out(n, n);
    for (Index j = 0; j < n; ++j) {
      for (Index i = 0; i < n; ++i) {
        out(i, j) = 0.;
      }
    }
    Index at = 0;
    for (const auto &b : blocks) {
      for (Index j = 0; j < b.cols(); ++j) {
        for (Index i = 0; i < b.rows(); ++i) {
          out(at + i, at + j) = b(i, j);
        }
      }
      at += b.rows();
    }
    return out;
  }
};

// The object DenseQRImplementation::compute returns: the member subset the reference uses of
// Eigen::ColPivHouseholderQR (sparse_gp.hpp:394-403, :577, :593; qr_utils.hpp:18-53).
class DeviceQR {
public:
  DeviceQR(const MatrixXd &m, std::shared_ptr<Device> dev = Device::default_device())
      : rows_(m.rows()), cols_(m.cols()), R_(m.cols(), m.cols()) {
    ALBATROSS_B200_CHECK(ab_qr_r(dev->get(), m.data(), m.rows(), m.cols(), R_.data()));
  }
  Index rows() const { return rows_; }
  Index cols() const { return cols_; }
  Index rank() const { return cols_; } // CholQR2 succeeds only at full numerical rank
  const MatrixXd &matrixR() const { return R_; }
  // identity: P.indices()[i] == i (the reference's colsPermutation())
  std::vector<Index> colsPermutationIndices() const {
    std::vector<Index> p(static_cast<std::size_t>(cols_));
    std::iota(p.begin(), p.end(), Index(0));
    return p;
  }

private:
  Index rows_, cols_;
  MatrixXd R_;
};

struct DenseQRImplementation { // sparse_gp.hpp:81-89
  using QRType = DeviceQR;
  // ThreadPool* of the reference signature is accepted as an opaque pointer and ignored
  static std::unique_ptr<QRType> compute(const MatrixXd &m, const void * /*threads*/ = nullptr) {
    return std::make_unique<QRType>(m);
  }
};
using DeviceQRImplementation = DenseQRImplementation;

inline MatrixXd get_R(const DeviceQR &qr) { return qr.matrixR(); }                       // qr_utils.hpp:18-27
inline std::vector<Index> get_P(const DeviceQR &qr) { return qr.colsPermutationIndices(); } // :29-33

// R^-T P^T rhs (qr_utils.hpp:35-45); P as its index vector: (P^T rhs)(i, :) = rhs(P[i], :).
inline MatrixXd sqrt_solve(const MatrixXd &R, const std::vector<Index> &P, const MatrixXd &rhs) {
  const Index m = R.rows();
  MatrixXd out(m, rhs.cols());
  for (Index c = 0; c < rhs.cols(); ++c) {
    for (Index i = 0; i < m; ++i) { // forward substitution with R^T (lower triangular)
      double acc = rhs(P[static_cast<std::size_t>(i)], c);
      for (Index k = 0; k < i; ++k) {
        acc -= R(k, i) * out(k, c);
      }
      out(i, c) = acc / R(i, i);
    }
  }
  return out;
}
inline MatrixXd sqrt_solve(const DeviceQR &qr, const MatrixXd &rhs) { // :47-53
  return sqrt_solve(qr.matrixR(), qr.colsPermutationIndices(), rhs);
}

} // namespace albatross_b200
