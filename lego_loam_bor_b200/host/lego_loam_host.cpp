// lego_loam_host.cpp -- see lego_loam_host.h.  Host glue only: everything per-point runs on the GPU
// through the C ABI, including MapOptimization's key frames and local map.
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
void Device::fail(const std::string& what) {
  {
    std::lock_guard<std::mutex> lk(_err_mtx);
    if (!_failed.load()) _error = what;
  }
  _failed.store(true);
  fprintf(stderr, "lego_loam: %s\n", what.c_str());
}
std::string Device::error() {
  std::lock_guard<std::mutex> lk(_err_mtx);
  return _error;
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
  handle(stamp, [&]() {
    const int32_t n = n_points;
    _dev->check(ll_set_scans_host(_dev->h(), xyzi, &n, std::max(1, n_points)), "ll_set_scans_host");
  });
}

void ImageProjection::cloudHandler(const ll_pointcloud2_view& msg) {
  // pcl::fromROSMsg needs x, y, z FLOAT32 fields; big-endian messages are not produced by any supported driver
  if (msg.is_bigendian || msg.off_x < 0 || msg.off_y < 0 || msg.off_z < 0 || msg.point_step < 12 ||
      (uint64_t)msg.width * (uint64_t)msg.height > 0x7fffffffull ||
      (uint64_t)msg.width * msg.height * msg.point_step > msg.data_len)
    throw std::runtime_error("ImageProjection::cloudHandler: unsupported or inconsistent sensor_msgs/PointCloud2");
  handle((double)msg.stamp_sec + 1e-9 * (double)msg.stamp_nsec, [&]() {
    const int32_t n = (int32_t)(msg.width * msg.height);
    static const uint8_t none = 0;
    _dev->check(ll_set_scans_pointcloud2_host(_dev->h(), msg.data ? msg.data : &none, &n, (size_t)msg.data_len, (int)msg.point_step, msg.off_x,
                                              msg.off_y, msg.off_z, msg.off_intensity, msg.is_dense),
                "ll_set_scans_pointcloud2_host");
  });
}

template <class SetScans>
void ImageProjection::handle(double stamp, SetScans set_scans) {
  if (_dev->failed()) throw std::runtime_error("ImageProjection::cloudHandler: a stage failed earlier: " + _dev->error());
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
    set_scans();
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
    if (!_dev->failed()) try {
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
        if (_dev->download_payloads) {  // MapOptimization reads them on the device
          out.cloud_corner_last = _dev->download_cloud(LL_BUF_SCAN_CORNER_DS);
          out.cloud_surf_last = _dev->download_cloud(LL_BUF_SCAN_SURF_DS);
          out.cloud_outlier_last = _dev->download_cloud(LL_BUF_SCAN_OUTLIER_DS);
        }
        _dev->check(ll_get_poses(_dev->h(), out.laser_odometry, nullptr, nullptr), "ll_get_poses");
        out.stamp = projection.seg_msg.stamp;
        out.frame = projection.frame;
      }
    } catch (const std::exception& e) {
      _dev->fail(e.what());
      if (hand_over) {  // the mapping cycle of this scan will not run
        std::lock_guard<std::mutex> lk(k.m);
        k.mapping_busy = false;
        hand_over = false;
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

MapOptimization::MapOptimization(const LegoLoamParams& params, std::shared_ptr<Device> dev, Channel<AssociationOut>& input_channel)
    : _dev(dev), _input_channel(input_channel) {
  {
    // cloudKeyPoses3D/6D, the key-frame clouds and the local maps live on the device (allocateMemory, mapOptmization.cpp:146-245)
    std::lock_guard<std::mutex> dl(_dev->mutex());
    const int n = params.num_vertical_scans * params.num_horizontal_scans;
    const int max_keyframes = 4096;  // ~34 minutes of mapping cycles at 2 Hz; outgrowing it stops the mapping cleanly (Device::failed)
    _dev->check(ll_map_enable_keyframes(_dev->h(), max_keyframes, max_keyframes * (n / 8 + 512), n, 2 * n), "ll_map_enable_keyframes");
  }
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

void MapOptimization::run() {
  Handshake& k = hs(_dev.get());
  while (true) {
    AssociationOut association;
    _input_channel.receive(association);
    if (association.shutdown) break;
    if (!_dev->failed()) try {
      // mapOptmization.cpp:1545-1560, every step on the device.  downsampleCurrentScan already ran at the hand-over
      // (FeatureAssociation thread), before the next frame could overwrite the last-frame clouds; it does not
      // depend on extractSurroundingKeyFrames, so the order of the two is immaterial.
      std::lock_guard<std::mutex> dl(_dev->mutex());
      // OdometryToTransform(association.laser_odometry, transformSum), mapOptmization.cpp:1539: the pose that belongs to
      // THIS scan -- FeatureAssociation may have integrated further scans since the hand-over
      _dev->check(ll_map_set_odometry(_dev->h(), association.laser_odometry), "ll_map_set_odometry");
      _dev->check(ll_map_predict_pose(_dev->h()), "ll_map_predict_pose");                            // transformAssociateToMap
      _dev->check(ll_map_extract_surrounding_keyframes(_dev->h()), "ll_map_extract_surrounding_keyframes");
      _dev->check(ll_scan_to_map(_dev->h()), "ll_scan_to_map");                                      // guards + transformUpdate inside
      _dev->check(ll_map_save_keyframe(_dev->h()), "ll_map_save_keyframe");                          // saveKeyFramesAndFactor
      float aft[6];
      int32_t state[4] = {0, 0, 0, 0};
      _dev->check(ll_download(_dev->h(), 0, LL_BUF_TRANSFORM_AFT_MAPPED, aft, sizeof(aft), nullptr), "ll_download");
      _dev->check(ll_download(_dev->h(), 0, LL_BUF_KEYFRAME_STATE, state, sizeof(state), nullptr), "ll_download");
      if (state[3] != 0) _dev->fail("MapOptimization: key-frame capacity exceeded (LL_BUF_KEYFRAME_STATE bits " + std::to_string(state[3]) + "); mapping stops");
      _n_key_frames = (size_t)state[0];
      std::lock_guard<std::mutex> pl(_pose_mtx);
      std::memcpy(_aft, aft, sizeof(aft));
    } catch (const std::exception& e) {
      _dev->fail(e.what());
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
