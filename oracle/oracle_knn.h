/*
 * oracle_knn.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Exact k-nearest-neighbour search standing in for nanoflann::KdTreeFLANN<PointType>
 * (include/lego_loam/nanoflann_pcl.h:54-152).  Two backends behind one class:
 *
 *  - ORACLE_WITH_NANOFLANN: the reference's own vendored nanoflann.hpp (1.3.0),
 *    compiled from where it lies in /root/reference (see oracle/Makefile, output only in
 *    oracle/_ref/), with the exact metric typedef of nanoflann_pcl.h:100-102
 *    (SO3_Adaptor == L2_Simple) and nearestKSearch semantics of nanoflann_pcl.h:140-152.
 *  - a self-contained median-split kd-tree (port) used when the reference is absent and
 *    to cross-check the first.
 *
 * Semantics shared by both: squared L2 in float accumulated ((0+dx^2)+dy^2)+dz^2 with
 * d = query - point (nanoflann.hpp:432-440); results ascending by distance; a candidate
 * replaces a stored one only if strictly closer (nanoflann.hpp:184).
 */
#ifndef ORACLE_KNN_H
#define ORACLE_KNN_H

#include <algorithm>
#include <cfloat>
#include <cstddef>
#include <memory>
#include <vector>

struct P4 {
  float x, y, z, i;
};

#ifdef ORACLE_WITH_NANOFLANN
#include <nanoflann.hpp>
#endif

namespace oknn {

inline float dist2(const float* q, const P4& p) {
  float r = 0.f;
  float d = q[0] - p.x;
  r += d * d;
  d = q[1] - p.y;
  r += d * d;
  d = q[2] - p.z;
  r += d * d;
  return r;
}

/* Same insertion rule as nanoflann::KNNResultSet::addPoint (nanoflann.hpp:175-202). */
struct ResultSet {
  int* idx;
  float* d2;
  int cap, count;
  void init(int* i, float* d, int k) {
    idx = i; d2 = d; cap = k; count = 0;
    if (cap) d2[cap - 1] = FLT_MAX;
  }
  float worst() const { return d2[cap - 1]; }
  void add(float dist, int index) {
    int i;
    for (i = count; i > 0; --i) {
      if (d2[i - 1] > dist) {
        if (i < cap) { d2[i] = d2[i - 1]; idx[i] = idx[i - 1]; }
      } else {
        break;
      }
    }
    if (i < cap) { d2[i] = dist; idx[i] = index; }
    if (count < cap) count++;
  }
};

/* Port backend: kd-tree with middle split on the widest bbox dimension, leaf <= 10. */
class PortTree {
 public:
  void build(const P4* pts, int n) {
    pts_ = pts; n_ = n;
    ind_.resize(n);
    for (int i = 0; i < n; ++i) ind_[i] = i;
    nodes_.clear();
    if (n > 0) root_ = divide(0, n); else root_ = -1;
  }
  int knn(const float* q, int k, int* idx, float* d2) const {
    ResultSet rs; rs.init(idx, d2, k);
    if (root_ >= 0) search(root_, q, rs);
    return rs.count;
  }
 private:
  struct Node { int left, right, lo, hi, dim; float split; float bmin[3], bmax[3]; };
  static float coord(const P4& p, int d) { return d == 0 ? p.x : (d == 1 ? p.y : p.z); }
  int divide(int lo, int hi) {
    Node nd; nd.lo = lo; nd.hi = hi; nd.left = nd.right = -1; nd.dim = 0; nd.split = 0.f;
    for (int d = 0; d < 3; ++d) { nd.bmin[d] = FLT_MAX; nd.bmax[d] = -FLT_MAX; }
    for (int i = lo; i < hi; ++i)
      for (int d = 0; d < 3; ++d) {
        const float c = coord(pts_[ind_[i]], d);
        nd.bmin[d] = std::min(nd.bmin[d], c); nd.bmax[d] = std::max(nd.bmax[d], c);
      }
    const int me = (int)nodes_.size();
    nodes_.push_back(nd);
    if (hi - lo > 10) {
      int dim = 0; float span = -1.f;
      for (int d = 0; d < 3; ++d) if (nd.bmax[d] - nd.bmin[d] > span) { span = nd.bmax[d] - nd.bmin[d]; dim = d; }
      const int mid = (lo + hi) / 2;
      std::nth_element(ind_.begin() + lo, ind_.begin() + mid, ind_.begin() + hi,
                       [&](int a, int b) { return coord(pts_[a], dim) < coord(pts_[b], dim); });
      const int l = divide(lo, mid);
      const int r = divide(mid, hi);
      nodes_[me].left = l; nodes_[me].right = r; nodes_[me].dim = dim;
    }
    return me;
  }
  float box_dist2(const Node& nd, const float* q) const {
    float r = 0.f;
    for (int d = 0; d < 3; ++d) {
      float diff = 0.f;
      if (q[d] < nd.bmin[d]) diff = nd.bmin[d] - q[d];
      else if (q[d] > nd.bmax[d]) diff = q[d] - nd.bmax[d];
      r += diff * diff;
    }
    return r;
  }
  void search(int ni, const float* q, ResultSet& rs) const {
    const Node& nd = nodes_[ni];
    if (nd.left < 0) {
      /* visit leaf points in ascending original index so that equal distances keep the lowest index */
      int tmp[16]; int m = 0;
      for (int i = nd.lo; i < nd.hi; ++i) tmp[m++] = ind_[i];
      std::sort(tmp, tmp + m);
      for (int i = 0; i < m; ++i) {
        const float d = dist2(q, pts_[tmp[i]]);
        if (d < rs.worst()) rs.add(d, tmp[i]);
      }
      return;
    }
    const Node& a = nodes_[nd.left];
    const Node& b = nodes_[nd.right];
    const float da = box_dist2(a, q), db = box_dist2(b, q);
    const int first = da <= db ? nd.left : nd.right;
    const int second = da <= db ? nd.right : nd.left;
    const float dsecond = da <= db ? db : da;
    search(first, q, rs);
    /* bbox distances are lower bounds up to rounding; keep a 1e-6 relative slack so pruning stays exact */
    if (dsecond * 0.999999f <= rs.worst()) search(second, q, rs);
  }
  const P4* pts_ = nullptr;
  int n_ = 0, root_ = -1;
  std::vector<int> ind_;
  std::vector<Node> nodes_;
};

#ifdef ORACLE_WITH_NANOFLANN
struct NfAdaptor {
  const P4* pts = nullptr;
  size_t n = 0;
  inline size_t kdtree_get_point_count() const { return n; }
  inline float kdtree_get_pt(const size_t idx, int dim) const {
    const P4& p = pts[idx];
    if (dim == 0) return p.x;
    else if (dim == 1) return p.y;
    else if (dim == 2) return p.z;
    else return 0.0f;
  }
  template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
};
typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::SO3_Adaptor<float, NfAdaptor>, NfAdaptor, 3, int> NfTree;
#endif

extern int g_use_nanoflann;

class KdTree {
 public:
  KdTree()
#ifdef ORACLE_WITH_NANOFLANN
      : nf_(3, adaptor_)
#endif
  {}
  /* setInputCloud: the tree keeps its own copy of the cloud, like the shared_ptr the adapter holds. */
  void setInputCloud(const std::vector<P4>& cloud) {
    cloud_ = cloud;
    built_ = true;
#ifdef ORACLE_WITH_NANOFLANN
    if (g_use_nanoflann) {
      adaptor_.pts = cloud_.data();
      adaptor_.n = cloud_.size();
      nf_.buildIndex();
      used_nf_ = true;
      return;
    }
#endif
    used_nf_ = false;
    port_.build(cloud_.data(), (int)cloud_.size());
  }
  bool built() const { return built_; }
  const std::vector<P4>& cloud() const { return cloud_; }
  /* nearestKSearch: fills up to k results; unfilled slots keep d2 = FLT_MAX, idx = -1. */
  int nearestKSearch(const P4& q, int k, int* idx, float* d2) const {
    for (int i = 0; i < k; ++i) { idx[i] = -1; d2[i] = FLT_MAX; }
    if (!built_ || cloud_.empty()) return 0;
    const float qq[3] = {q.x, q.y, q.z};
#ifdef ORACLE_WITH_NANOFLANN
    if (used_nf_) {
      nanoflann::KNNResultSet<float, int> rs(k);
      rs.init(idx, d2);
      nf_.findNeighbors(rs, qq, nanoflann::SearchParams());
      return (int)rs.size();
    }
#endif
    return port_.knn(qq, k, idx, d2);
  }
  /* radiusSearch of nanoflann_pcl.h:155-175: squared radius computed in double and cast to float, points accepted
   * while d2 < r2 (strict, nanoflann.hpp:249-253), results sorted ascending by distance (sorted = true is the
   * wrapper's default, nanoflann_pcl.h:69).  The sort is std::sort on distance only, so the order of equal
   * distances is unspecified in the reference; the port backend keeps ascending index for them. */
  int radiusSearch(const P4& q, double radius, std::vector<int>& idx, std::vector<float>& d2) const {
    idx.clear(); d2.clear();
    if (!built_ || cloud_.empty()) return 0;
    const float qq[3] = {q.x, q.y, q.z};
    const float r2 = static_cast<float>(radius * radius);
#ifdef ORACLE_WITH_NANOFLANN
    if (used_nf_) {
      std::vector<std::pair<int, float> > id;
      id.reserve(128);
      nanoflann::RadiusResultSet<float, int> rs(r2, id);
      nf_.findNeighbors(rs, qq, nanoflann::SearchParams());
      std::sort(id.begin(), id.end(), nanoflann::IndexDist_Sorter());
      for (size_t i = 0; i < id.size(); ++i) { idx.push_back(id[i].first); d2.push_back(id[i].second); }
      return (int)id.size();
    }
#endif
    std::vector<std::pair<float, int> > id;
    for (size_t i = 0; i < cloud_.size(); ++i) {
      const float d = dist2(qq, cloud_[i]);
      if (d < r2) id.push_back(std::make_pair(d, (int)i));
    }
    std::stable_sort(id.begin(), id.end(),
                     [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first < b.first; });
    for (size_t i = 0; i < id.size(); ++i) { idx.push_back(id[i].second); d2.push_back(id[i].first); }
    return (int)id.size();
  }
 private:
  std::vector<P4> cloud_;
  bool built_ = false;
  bool used_nf_ = false;
  PortTree port_;
#ifdef ORACLE_WITH_NANOFLANN
  NfAdaptor adaptor_;
  NfTree nf_;
#endif
};

}  // namespace oknn

#endif
