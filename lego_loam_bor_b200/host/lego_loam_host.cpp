// lego_loam_host.cpp -- see lego_loam_host.h.  Host glue only: everything per-point runs on the GPU
// through the C ABI, except the sub-map assembly of MapOptimization::extractSurroundingKeyFrames,
// which the survey scopes as host work for this round.
#include "lego_loam_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

namespace lego_loam {

// ------------------------------------------------------------------------------------------ Device

struct Handshake {
  std::mutex m;
  std::condition_variable cv;
  uint64_t projected = 0, associated = 0;  // frames through ImageProjection / FeatureAssociation
  bool mapping_busy = false;               // MapOptimization still owns the down-sampled scan buffers
};
static Handshake& hs(Device* d) {
  static std::mutex g;
  static std::vector<std::pair<Device*, std::unique_ptr<Handshake>>> all;
  std::lock_guard<std::mutex> lk(g);
  for (auto& e : all)
    if (e.first == d) return *e.second;
  all.emplace_back(d, std::unique_ptr<Handshake>(new Handshake()));
  return *all.back().second;
}

Device::Device(const LegoLoamParams& params, int max_points, int cuda_device) : _params(params) {
  const int n = params.num_vertical_scans * params.num_horizontal_scans;
  const int rc = ll_create(&params, 1, max_points > 0 ? max_points : n, cuda_device, nullptr, &_h);
  if (rc != LL_OK) throw std::runtime_error("ll_create failed (" + std::to_string(rc) + "): no usable CUDA device or bad parameters");
}
Device::~Device() {
  if (_h) ll_destroy(_h);
}
void Device::check(int rc, const char* what) const {
  if (rc < 0) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + ll_last_error(_h));
}
void Device::waitIdle() {
  Handshake& k = hs(this);
  std::unique_lock<std::mutex> lk(k.m);
  k.cv.wait(lk, [&]() { return k.associated == k.projected && !k.mapping_busy; });
}
template <typename T>
std::vector<T> Device::download(int buffer) {
  size_t n = 0;
  check(ll_download(_h, 0, buffer, nullptr, 0, &n), "ll_download");
  std::vector<T> out(n);
  if (n) check(ll_download(_h, 0, buffer, out.data(), n * sizeof(T), &n), "ll_download");
  return out;
}
Cloud Device::download_cloud(int buffer) { return download<PointXYZI>(buffer); }
template std::vector<float> Device::download<float>(int);
template std::vector<int32_t> Device::download<int32_t>(int);
template std::vector<uint32_t> Device::download<uint32_t>(int);
template std::vector<uint8_t> Device::download<uint8_t>(int);

// --------------------------------------------------------------------------------- ImageProjection

ImageProjection::ImageProjection(const LegoLoamParams&, std::shared_ptr<Device> dev, Channel<ProjectionOut>& output_channel)
    : _dev(dev), _output_channel(output_channel) {}

void ImageProjection::cloudHandler(const float* xyzi, int n_points, double stamp) {
  Handshake& k = hs(_dev.get());
  {
    // device state is shared: wait until FeatureAssociation has consumed the previous projection
    std::unique_lock<std::mutex> lk(k.m);
    k.cv.wait(lk, [&]() { return k.associated == k.projected; });
  }
  ProjectionOut out;
  out.frame = _frame++;
  out.seg_msg.stamp = stamp;
  {
    std::lock_guard<std::mutex> lk(_dev->mutex());
    const int32_t n = n_points;
    _dev->check(ll_set_scans_host(_dev->h(), xyzi, &n, std::max(1, n_points)), "ll_set_scans_host");
    _dev->check(ll_image_projection(_dev->h()), "ll_image_projection");
    if (_dev->download_payloads) {
      out.segmented_cloud = _dev->download_cloud(LL_BUF_SEG_CLOUD);
      out.outlier_cloud = _dev->download_cloud(LL_BUF_OUTLIER_CLOUD);
      out.seg_msg.startRingIndex = _dev->download<int32_t>(LL_BUF_START_RING_INDEX);
      out.seg_msg.endRingIndex = _dev->download<int32_t>(LL_BUF_END_RING_INDEX);
      out.seg_msg.segmentedCloudGroundFlag = _dev->download<uint8_t>(LL_BUF_SEG_GROUND_FLAG);
      out.seg_msg.segmentedCloudColInd = _dev->download<uint32_t>(LL_BUF_SEG_COL_IND);
      out.seg_msg.segmentedCloudRange = _dev->download<float>(LL_BUF_SEG_RANGE);
      const std::vector<float> o = _dev->download<float>(LL_BUF_ORIENTATION);
      out.seg_msg.startOrientation = o[0]; out.seg_msg.endOrientation = o[1]; out.seg_msg.orientationDiff = o[2];
    }
  }
  {
    std::lock_guard<std::mutex> lk(k.m);
    k.projected++;
  }
  _output_channel.send(std::move(out));
}

// ------------------------------------------------------------------------------ FeatureAssociation

FeatureAssociation::FeatureAssociation(const LegoLoamParams&, std::shared_ptr<Device> dev, Channel<ProjectionOut>& input_channel,
                                       Channel<AssociationOut>& output_channel)
    : _dev(dev), _input_channel(input_channel), _output_channel(output_channel) {
  _run_thread = std::thread(&FeatureAssociation::runFeatureAssociation, this);  // featureAssociation.cpp:87
}

FeatureAssociation::~FeatureAssociation() {
  ProjectionOut stop;
  stop.shutdown = true;  // featureAssociation.cpp:90-94 sends an empty item
  _input_channel.send(std::move(stop));
  _run_thread.join();
}

void FeatureAssociation::transformSum(float out6[6]) {
  std::lock_guard<std::mutex> lk(_dev->mutex());
  _dev->check(ll_get_poses(_dev->h(), out6, nullptr, nullptr), "ll_get_poses");
}

void FeatureAssociation::runFeatureAssociation() {
  Handshake& k = hs(_dev.get());
  while (true) {
    ProjectionOut projection;
    _input_channel.receive(projection);
    if (projection.shutdown) break;
    AssociationOut out;
    bool hand_over = false;
    {
      std::unique_lock<std::mutex> dl(_dev->mutex());
      const int rc = ll_feature_association(_dev->h());
      _dev->check(rc, "ll_feature_association");
      hand_over = rc == 1;  // every mapping_frequency_divider-th odometry frame (featureAssociation.cpp:1432)
      if (hand_over) {
        dl.unlock();
        {
          std::unique_lock<std::mutex> lk(k.m);  // the down-sampled scan buffers belong to MapOptimization until it is done
          k.cv.wait(lk, [&]() { return !k.mapping_busy; });
          k.mapping_busy = true;
        }
        dl.lock();
        _dev->check(ll_map_downsample_current_scan(_dev->h()), "ll_map_downsample_current_scan");
        out.cloud_corner_last = _dev->download_cloud(LL_BUF_SCAN_CORNER_DS);
        out.cloud_surf_last = _dev->download_cloud(LL_BUF_SCAN_SURF_DS);
        out.cloud_outlier_last = _dev->download_cloud(LL_BUF_SCAN_OUTLIER_DS);
        _dev->check(ll_get_poses(_dev->h(), out.laser_odometry, nullptr, nullptr), "ll_get_poses");
        out.stamp = projection.seg_msg.stamp;
        out.frame = projection.frame;
      }
    }
    {
      std::lock_guard<std::mutex> lk(k.m);
      k.associated++;
    }
    k.cv.notify_all();
    if (hand_over) _output_channel.send(std::move(out));
  }
}

// --------------------------------------------------------------------------------- MapOptimization

void voxelGridFilter(const Cloud& in, float leaf, Cloud& out) {
  out.clear();
  if (in.empty()) return;
  const float inv = 1.0f / leaf;
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (const PointXYZI& p : in) {
    const float c[3] = {p.x, p.y, p.z};
    if (!std::isfinite(c[0]) || !std::isfinite(c[1]) || !std::isfinite(c[2])) continue;
    for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], c[d]); mx[d] = std::max(mx[d], c[d]); }
  }
  long long ext[3];
  int lo[3], span[3];
  for (int d = 0; d < 3; ++d) {
    ext[d] = (long long)((mx[d] - mn[d]) * inv) + 1;
    lo[d] = (int)std::floor(mn[d] * inv);
    span[d] = (int)std::floor(mx[d] * inv) - lo[d] + 1;
  }
  if (ext[0] * ext[1] * ext[2] > 2147483647LL) { out = in; return; }  // PCL gives the input back
  struct Key { int vox; int src; };
  std::vector<Key> keys;
  keys.reserve(in.size());
  for (size_t i = 0; i < in.size(); ++i) {
    const PointXYZI& p = in[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    const int a = (int)std::floor(p.x * inv) - lo[0], b = (int)std::floor(p.y * inv) - lo[1], c = (int)std::floor(p.z * inv) - lo[2];
    keys.push_back(Key{a + b * span[0] + c * span[0] * span[1], (int)i});
  }
  std::stable_sort(keys.begin(), keys.end(), [](const Key& l, const Key& r) { return l.vox < r.vox; });
  for (size_t s = 0; s < keys.size();) {
    size_t e = s;
    float sx = 0, sy = 0, sz = 0, si = 0;
    while (e < keys.size() && keys[e].vox == keys[s].vox) {
      const PointXYZI& p = in[keys[e].src];
      sx += p.x; sy += p.y; sz += p.z; si += p.intensity;
      ++e;
    }
    const float n = (float)(e - s);
    out.push_back(PointXYZI{sx / n, sy / n, sz / n, si / n});
    s = e;
  }
}

static void transformCloud(const Cloud& in, const float pose[6], Cloud& out) {
  // mapOptmization.cpp:428-473 (roll = pose[0], pitch = pose[1], yaw = pose[2])
  const float cr = std::cos(pose[0]), sr = std::sin(pose[0]), cp = std::cos(pose[1]), sp = std::sin(pose[1]);
  const float cy = std::cos(pose[2]), sy = std::sin(pose[2]);
  out.resize(in.size());
  for (size_t i = 0; i < in.size(); ++i) {
    const PointXYZI& p = in[i];
    const float x1 = cy * p.x - sy * p.y, y1 = sy * p.x + cy * p.y, z1 = p.z;
    const float x2 = x1, y2 = cr * y1 - sr * z1, z2 = sr * y1 + cr * z1;
    out[i] = PointXYZI{cp * x2 + sp * z2 + pose[3], y2 + pose[4], -sp * x2 + cp * z2 + pose[5], p.intensity};
  }
}

MapOptimization::MapOptimization(const LegoLoamParams& params, std::shared_ptr<Device> dev, Channel<AssociationOut>& input_channel)
    : _dev(dev), _input_channel(input_channel), _search_radius(params.surrounding_keyframe_search_radius) {
  _run_thread = std::thread(&MapOptimization::run, this);  // mapOptmization.cpp:122
}

MapOptimization::~MapOptimization() {
  AssociationOut stop;
  stop.shutdown = true;  // mapOptmization.cpp:126-129
  _input_channel.send(std::move(stop));
  _run_thread.join();
}

void MapOptimization::transformAftMapped(float out6[6]) {
  std::lock_guard<std::mutex> lk(_pose_mtx);
  std::memcpy(out6, _aft, sizeof(_aft));
}

void MapOptimization::extractSurroundingKeyFrames(const float pos[3]) {
  _corner_from_map_ds.clear();
  _surf_from_map_ds.clear();
  if (_key_frames.empty()) return;
  // radius search over the key poses, nearest first (kdtreeSurroundingKeyPoses.radiusSearch, sorted)
  std::vector<std::pair<float, int>> near;
  for (size_t i = 0; i < _key_frames.size(); ++i) {
    const float* kp = _key_frames[i].pose;
    const float d2 = (kp[3] - pos[0]) * (kp[3] - pos[0]) + (kp[4] - pos[1]) * (kp[4] - pos[1]) + (kp[5] - pos[2]) * (kp[5] - pos[2]);
    if (d2 <= _search_radius * _search_radius) near.emplace_back(d2, (int)i);
  }
  std::sort(near.begin(), near.end());
  Cloud poses, poses_ds;
  for (auto& n : near) {
    const float* kp = _key_frames[n.second].pose;
    poses.push_back(PointXYZI{kp[3], kp[4], kp[5], (float)n.second});  // intensity = key-frame index
  }
  voxelGridFilter(poses, 1.0f, poses_ds);  // downSizeFilterSurroundingKeyPoses; the id is (int)mean intensity (sic)
  // drop key frames that left the surrounding set (mapOptmization.cpp:935-955)
  for (size_t i = 0; i < _surrounding_ids.size();) {
    bool keep = false;
    for (const PointXYZI& q : poses_ds) keep = keep || (_surrounding_ids[i] == (int)q.intensity);
    if (keep) { ++i; continue; }
    _surrounding_ids.erase(_surrounding_ids.begin() + i);
    _surrounding_corner.erase(_surrounding_corner.begin() + i);
    _surrounding_surf.erase(_surrounding_surf.begin() + i);
    _surrounding_outlier.erase(_surrounding_outlier.begin() + i);
  }
  // add the new ones, transformed into the map frame (mapOptmization.cpp:957-980)
  for (const PointXYZI& q : poses_ds) {
    const int id = (int)q.intensity;
    if (std::find(_surrounding_ids.begin(), _surrounding_ids.end(), id) != _surrounding_ids.end()) continue;
    if (id < 0 || id >= (int)_key_frames.size()) continue;
    const KeyFrame& kf = _key_frames[id];
    Cloud c, s, o;
    transformCloud(kf.corner, kf.pose, c);
    transformCloud(kf.surf, kf.pose, s);
    transformCloud(kf.outlier, kf.pose, o);
    _surrounding_ids.push_back(id);
    _surrounding_corner.push_back(std::move(c));
    _surrounding_surf.push_back(std::move(s));
    _surrounding_outlier.push_back(std::move(o));
  }
  Cloud corner_all, surf_all;
  for (size_t i = 0; i < _surrounding_ids.size(); ++i) {
    corner_all.insert(corner_all.end(), _surrounding_corner[i].begin(), _surrounding_corner[i].end());
    surf_all.insert(surf_all.end(), _surrounding_surf[i].begin(), _surrounding_surf[i].end());
    surf_all.insert(surf_all.end(), _surrounding_outlier[i].begin(), _surrounding_outlier[i].end());
  }
  voxelGridFilter(corner_all, 0.2f, _corner_from_map_ds);  // downSizeFilterCorner
  voxelGridFilter(surf_all, 0.4f, _surf_from_map_ds);      // downSizeFilterSurf
}

void MapOptimization::saveKeyFramesAndFactor(const AssociationOut& in) {
  float aft[6], tobe[6];
  _dev->check(ll_download(_dev->h(), 0, LL_BUF_TRANSFORM_AFT_MAPPED, aft, sizeof(aft), nullptr), "ll_download");
  _dev->check(ll_download(_dev->h(), 0, LL_BUF_TRANSFORM_TOBE_MAPPED, tobe, sizeof(tobe), nullptr), "ll_download");
  const float cur[3] = {aft[3], aft[4], aft[5]};  // currentRobotPosPoint (mapOptmization.cpp:1336-1338)
  const float d = std::sqrt((_previous_pos[0] - cur[0]) * (_previous_pos[0] - cur[0]) + (_previous_pos[1] - cur[1]) * (_previous_pos[1] - cur[1]) +
                            (_previous_pos[2] - cur[2]) * (_previous_pos[2] - cur[2]));
  const bool save = !(d < 0.3);
  if (!save && !_key_frames.empty()) return;
  std::memcpy(_previous_pos, cur, sizeof(cur));
  KeyFrame kf;
  // first key frame: prior on transformTobeMapped; later: transformAftMapped (== tobe after transformUpdate).
  // iSAM2 of a pure odometry chain returns its initial values: identity.
  std::memcpy(kf.pose, _key_frames.empty() ? tobe : aft, sizeof(kf.pose));
  kf.corner = in.cloud_corner_last;
  kf.surf = in.cloud_surf_last;
  kf.outlier = in.cloud_outlier_last;
  _key_frames.push_back(std::move(kf));
  _n_key_frames = _key_frames.size();
}

void MapOptimization::run() {
  Handshake& k = hs(_dev.get());
  while (true) {
    AssociationOut association;
    _input_channel.receive(association);
    if (association.shutdown) break;
    {
      std::lock_guard<std::mutex> dl(_dev->mutex());
      _dev->check(ll_map_predict_pose(_dev->h()), "ll_map_predict_pose");  // transformAssociateToMap
      float pos[3];
      {
        std::lock_guard<std::mutex> pl(_pose_mtx);
        pos[0] = _aft[3]; pos[1] = _aft[4]; pos[2] = _aft[5];
      }
      extractSurroundingKeyFrames(pos);
      // downsampleCurrentScan already ran on the device at hand-over (FeatureAssociation thread)
      _dev->check(ll_map_set_local(_dev->h(), 0, (const float*)_corner_from_map_ds.data(), (int)_corner_from_map_ds.size(),
                                   (const float*)_surf_from_map_ds.data(), (int)_surf_from_map_ds.size()), "ll_map_set_local");
      _dev->check(ll_scan_to_map(_dev->h()), "ll_scan_to_map");  // guards + transformUpdate inside
      saveKeyFramesAndFactor(association);
      float aft[6];
      _dev->check(ll_download(_dev->h(), 0, LL_BUF_TRANSFORM_AFT_MAPPED, aft, sizeof(aft), nullptr), "ll_download");
      std::lock_guard<std::mutex> pl(_pose_mtx);
      std::memcpy(_aft, aft, sizeof(aft));
    }
    _cycles++;
    {
      std::lock_guard<std::mutex> lk(k.m);
      k.mapping_busy = false;
    }
    k.cv.notify_all();
  }
}

}  // namespace lego_loam
