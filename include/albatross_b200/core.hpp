// albatross_b200 C++ trait layer — distributions, datasets and the integer (indexing) contract.
//
// Host-side value types with the reference's names and members: MarginalDistribution /
// JointDistribution (src/core/distribution.hpp:27-234), RegressionDataset (src/core/dataset.hpp:24-84),
// GroupIndexer / group_by / LeaveOneOutGrouper / KFoldGrouper (src/indexing/group_by.hpp:349-435),
// subset / set_subset / indices_complement (src/indexing/subset.hpp:21-205).  The index outputs are the
// bit-exact part of the contract; nothing here touches the device.
#pragma once

#include <algorithm>
#include <cassert>
#include <map>
#include <numeric>
#include <set>
#include <type_traits>
#include <vector>

#include "linalg_types.hpp"

namespace albatross_b200 {

using GroupIndices = std::vector<std::size_t>;
template <typename GroupKey> using GroupIndexer = std::map<GroupKey, GroupIndices>;

// ---- subset (src/indexing/subset.hpp) ---------------------------------------------------------

template <typename SizeType, typename X>
inline std::vector<X> subset(const std::vector<X> &v, const std::vector<SizeType> &indices) {
  std::vector<X> out;
  out.reserve(indices.size());
  for (const auto &i : indices) {
    out.push_back(v[static_cast<std::size_t>(i)]);
  }
  return out;
}

template <typename SizeType>
inline VectorXd subset(const VectorXd &v, const std::vector<SizeType> &indices) {
  VectorXd out(static_cast<Index>(indices.size()));
  for (std::size_t i = 0; i < indices.size(); ++i) {
    out[static_cast<Index>(i)] = v[static_cast<Index>(indices[i])];
  }
  return out;
}

template <typename SizeType>
inline void set_subset(const VectorXd &from, const std::vector<SizeType> &indices, VectorXd *to) {
  assert(static_cast<std::size_t>(from.size()) == indices.size());
  for (std::size_t i = 0; i < indices.size(); ++i) {
    (*to)[static_cast<Index>(indices[i])] = from[static_cast<Index>(i)];
  }
}

// symmetric_subset, subset.hpp:118-130
template <typename SizeType>
inline MatrixXd symmetric_subset(const MatrixXd &m, const std::vector<SizeType> &indices) {
  const Index k = static_cast<Index>(indices.size());
  MatrixXd out(k, k);
  for (Index j = 0; j < k; ++j) {
    for (Index i = 0; i < k; ++i) {
      out(i, j) = m(static_cast<Index>(indices[static_cast<std::size_t>(i)]),
                    static_cast<Index>(indices[static_cast<std::size_t>(j)]));
    }
  }
  return out;
}

// indices_complement, subset.hpp:186-205: [0, n) minus `indices`, ascending.
inline GroupIndices indices_complement(const GroupIndices &indices, std::size_t n) {
  std::set<std::size_t> held(indices.begin(), indices.end());
  GroupIndices out;
  for (std::size_t i = 0; i < n; ++i) {
    if (held.find(i) == held.end()) {
      out.push_back(i);
    }
  }
  return out;
}

// ---- distributions ----------------------------------------------------------------------------

struct MarginalDistribution {
  VectorXd mean;
  DiagonalMatrixXd covariance;

  MarginalDistribution() = default;
  // distribution.hpp:71-75: no variance given = zeros
  MarginalDistribution(const VectorXd &mean_) : mean(mean_), covariance(zeros(mean_.size())) {}
  MarginalDistribution(const VectorXd &mean_, const DiagonalMatrixXd &covariance_)
      : mean(mean_), covariance(covariance_) {
    assert_valid();
  }
  MarginalDistribution(const VectorXd &mean_, const VectorXd &variance_)
      : mean(mean_), covariance(variance_) {
    assert_valid();
  }
  MarginalDistribution(double mean_, double variance_) : mean(1), covariance(1) {
    mean[0] = mean_;
    covariance.diagonal()[0] = variance_;
  }

  std::size_t size() const { return static_cast<std::size_t>(mean.size()); }
  void assert_valid() const { assert(mean.size() == covariance.diagonal().size()); }
  double get_diagonal(Index i) const { return covariance.diagonal()[i]; }
  bool has_covariance() const {
    for (Index i = 0; i < covariance.diagonal().size(); ++i) {
      if (covariance.diagonal()[i] != 0.) {
        return true;
      }
    }
    return false;
  }
  MarginalDistribution operator[](std::size_t i) const {
    return MarginalDistribution(mean[static_cast<Index>(i)], get_diagonal(static_cast<Index>(i)));
  }
  bool operator==(const MarginalDistribution &o) const {
    return mean == o.mean && covariance.diagonal() == o.covariance.diagonal();
  }

  template <typename SizeType> MarginalDistribution subset(const std::vector<SizeType> &indices) const {
    return MarginalDistribution(albatross_b200::subset(mean, indices),
                                albatross_b200::subset(VectorXd(covariance.diagonal()), indices));
  }
  template <typename SizeType>
  void set_subset(const MarginalDistribution &from, const std::vector<SizeType> &indices) {
    albatross_b200::set_subset(from.mean, indices, &mean);
    VectorXd d(covariance.diagonal());
    albatross_b200::set_subset(VectorXd(from.covariance.diagonal()), indices, &d);
    covariance = DiagonalMatrixXd(d);
  }

private:
  static VectorXd zeros(Index n) {
    VectorXd z(n);
    for (Index i = 0; i < n; ++i) {
      z[i] = 0.;
    }
    return z;
  }
};

struct JointDistribution {
  VectorXd mean;
  MatrixXd covariance;

  JointDistribution() = default;
  JointDistribution(double mean_, double variance_) : mean(1), covariance(1, 1) {
    mean[0] = mean_;
    covariance(0, 0) = variance_;
  }
  JointDistribution(const VectorXd &mean_, const MatrixXd &covariance_)
      : mean(mean_), covariance(covariance_) {
    assert_valid();
  }
  std::size_t size() const { return static_cast<std::size_t>(mean.size()); }
  void assert_valid() const {
    assert(mean.size() == covariance.rows() && covariance.rows() == covariance.cols());
  }
  double get_diagonal(Index i) const { return covariance(i, i); }
  bool operator==(const JointDistribution &o) const { return mean == o.mean && covariance == o.covariance; }
  MarginalDistribution marginal() const { // distribution.hpp:229-232
    VectorXd var(mean.size());
    for (Index i = 0; i < mean.size(); ++i) {
      var[i] = covariance(i, i);
    }
    return MarginalDistribution(mean, var);
  }
  template <typename SizeType> JointDistribution subset(const std::vector<SizeType> &indices) const {
    return JointDistribution(albatross_b200::subset(mean, indices), symmetric_subset(covariance, indices));
  }
};

// ---- grouping -----------------------------------------------------------------------------------

struct LeaveOneOutGrouper {};                 // group_by.hpp:389-403: index i -> {i}
struct KFoldGrouper {                         // group_by.hpp:420-435: index i -> fold i % k
  KFoldGrouper(std::size_t k_ = 2) : k(k_) {}
  std::size_t k;
};

// IndexerBuilder<>::build, group_by.hpp:349-376: keys ordered by std::map, member indices ascending
// in encounter order.
template <typename GrouperFunction, typename Iterable>
inline auto build_indexer(const GrouperFunction &grouper_function, const Iterable &iterable) {
  using Value = typename Iterable::value_type;
  using GroupKey = typename std::decay<decltype(grouper_function(std::declval<const Value &>()))>::type;
  GroupIndexer<GroupKey> output;
  std::size_t i = 0;
  for (const auto &value : iterable) {
    output[grouper_function(value)].push_back(i);
    ++i;
  }
  return output;
}

template <typename Iterable>
inline GroupIndexer<std::size_t> build_indexer(const LeaveOneOutGrouper &, const Iterable &iterable) {
  GroupIndexer<std::size_t> output;
  std::size_t i = 0;
  for (auto it = iterable.begin(); it != iterable.end(); ++it, ++i) {
    output.emplace_hint(output.end(), i, GroupIndices{i});
  }
  return output;
}

template <typename Iterable>
inline GroupIndexer<std::size_t> build_indexer(const KFoldGrouper &grouper, const Iterable &iterable) {
  GroupIndexer<std::size_t> output;
  std::size_t i = 0;
  for (auto it = iterable.begin(); it != iterable.end(); ++it, ++i) {
    output[i % grouper.k].push_back(i);
  }
  return output;
}

// Grouped<Key, Value>: the map the reference's group_by / cross-validation calls return
// (group_by.hpp:60-346), trimmed to the members the GP path uses.
template <typename KeyType, typename ValueType> class Grouped {
public:
  using MapType = std::map<KeyType, ValueType>;
  Grouped() = default;
  Grouped(const MapType &m) : map_(m) {}
  Grouped(MapType &&m) : map_(std::move(m)) {}

  const MapType &get_map() const { return map_; }
  operator const MapType &() const { return map_; }
  std::size_t size() const { return map_.size(); }
  const ValueType &at(const KeyType &k) const { return map_.at(k); }
  ValueType &operator[](const KeyType &k) { return map_[k]; }
  auto begin() const { return map_.begin(); }
  auto end() const { return map_.end(); }
  auto find(const KeyType &k) const { return map_.find(k); }
  std::vector<KeyType> keys() const {
    std::vector<KeyType> out;
    for (const auto &pair : map_) {
      out.push_back(pair.first);
    }
    return out;
  }
  std::vector<ValueType> values() const {
    std::vector<ValueType> out;
    for (const auto &pair : map_) {
      out.push_back(pair.second);
    }
    return out;
  }
  ValueType first_value() const { return map_.begin()->second; }
  template <typename F> auto apply(F &&f) const {
    using Out = typename std::decay<decltype(f(std::declval<const KeyType &>(),
                                               std::declval<const ValueType &>()))>::type;
    Grouped<KeyType, Out> out;
    for (const auto &pair : map_) {
      out[pair.first] = f(pair.first, pair.second);
    }
    return out;
  }

private:
  MapType map_;
};

// ---- dataset --------------------------------------------------------------------------------------

template <typename FeatureType> struct RegressionDataset;

template <typename FeatureType, typename GrouperFunction> class GroupBy {
public:
  GroupBy(const RegressionDataset<FeatureType> &parent, const GrouperFunction &grouper)
      : parent_(parent), indexers_(build_indexer(grouper, parent.features)) {}
  using IndexerType = decltype(build_indexer(std::declval<const GrouperFunction &>(),
                                             std::declval<const std::vector<FeatureType> &>()));
  using KeyType = typename IndexerType::key_type;

  const IndexerType &indexers() const { return indexers_; }
  std::size_t size() const { return indexers_.size(); }
  std::vector<KeyType> keys() const {
    std::vector<KeyType> out;
    for (const auto &pair : indexers_) {
      out.push_back(pair.first);
    }
    return out;
  }
  Grouped<KeyType, RegressionDataset<FeatureType>> groups() const {
    Grouped<KeyType, RegressionDataset<FeatureType>> out;
    for (const auto &pair : indexers_) {
      out[pair.first] = parent_.subset(pair.second);
    }
    return out;
  }
  Grouped<KeyType, std::size_t> counts() const {
    Grouped<KeyType, std::size_t> out;
    for (const auto &pair : indexers_) {
      out[pair.first] = pair.second.size();
    }
    return out;
  }

private:
  RegressionDataset<FeatureType> parent_;
  IndexerType indexers_;
};

template <typename FeatureType> struct RegressionDataset {
  std::vector<FeatureType> features;
  MarginalDistribution targets;

  RegressionDataset() = default;
  RegressionDataset(const std::vector<FeatureType> &features_, const MarginalDistribution &targets_)
      : features(features_), targets(targets_) {
    assert(features.size() == targets.size());
  }
  RegressionDataset(const std::vector<FeatureType> &features_, const VectorXd &targets_)
      : RegressionDataset(features_, MarginalDistribution(targets_)) {}

  std::size_t size() const { return features.size(); }

  template <typename SizeType> RegressionDataset subset(const std::vector<SizeType> &indices) const {
    return RegressionDataset(albatross_b200::subset(features, indices), targets.subset(indices));
  }

  template <typename GrouperFunction>
  GroupBy<FeatureType, GrouperFunction> group_by(GrouperFunction grouper) const {
    return GroupBy<FeatureType, GrouperFunction>(*this, grouper);
  }
};

template <typename FeatureType>
inline RegressionDataset<FeatureType> create_dataset(const std::vector<FeatureType> &features,
                                                     const MarginalDistribution &targets) {
  return RegressionDataset<FeatureType>(features, targets);
}

// linspace, src/utils/ (used by UniformlySpacedInducingPoints): n points from a to b inclusive.
inline std::vector<double> linspace(double a, double b, std::size_t n) {
  std::vector<double> xs(n);
  const double step = (b - a) / static_cast<double>(n - 1);
  double val = a;
  for (std::size_t i = 0; i < n; ++i) {
    xs[i] = val;
    val += step;
  }
  return xs;
}

} // namespace albatross_b200
