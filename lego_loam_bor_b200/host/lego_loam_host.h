// lego_loam_host.h -- host-side C++ mirror of the reference's stage classes on top of the C ABI.
//
// Same class names and method shapes as the reference (LeGO-LOAM/src/imageProjection.h:8-16,
// featureAssociation.h:10-19, mapOptimization.h:36-43), the same payload structs (utility.h:64-80,
// cloud_msgs/msg/cloud_info.msg) and the same Channel (include/lego_loam/channel.h), with
//   ros::NodeHandle&                       ->  const LegoLoamParams& (the 21 yaml keys)
//   sensor_msgs::PointCloud2ConstPtr       ->  (const float* xyzi, int n, double stamp)
//   pcl::PointCloud<PointXYZI>::Ptr        ->  std::vector<PointXYZI> (16-byte points)
// All heavy work is enqueued on the GPU through include/lego_loam_b200.h; the payloads that travel
// through the channels carry host copies only when Device::download_payloads asks for them.  Device
// state is shared by the three stages, so the hand-offs are synchronous (blocking channels = the
// reference's bag mode, main.cpp:37-38) and GPU calls are serialised by one mutex; the stage threads
// exist for interface parity, not for overlap.
//
// MapOptimization's key frames, surrounding-key-frame selection and local-map assembly run on the
// device too (ll_map_save_keyframe / ll_map_extract_surrounding_keyframes);
// iSAM2 is the identity because loop closure is off (loam_config.yaml:24).
#pragma once

#include <atomic>
#include <string>
#include <condition_variable>
#include <cstdint>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/lego_loam_b200.h"

namespace lego_loam {

struct PointXYZI {
  float x, y, z, intensity;
};
typedef std::vector<PointXYZI> Cloud;

// include/lego_loam/channel.h:11-56, restated
template <class T>
class Channel {
 public:
  explicit Channel(bool blocking_send) : _empty(true), _blocking_send(blocking_send) {}
  void send(T&& item) {
    std::unique_lock<std::mutex> lock(_m);
    if (_blocking_send) _cv.wait(lock, [&]() { return _empty; });
    _item = std::move(item);
    _empty = false;
    _cv.notify_all();
  }
  void send(const T& item) {
    std::unique_lock<std::mutex> lock(_m);
    if (_blocking_send) _cv.wait(lock, [&]() { return _empty; });
    _item = item;
    _empty = false;
    _cv.notify_all();
  }
  void receive(T& item) {
    std::unique_lock<std::mutex> lock(_m);
    _cv.wait(lock, [&]() { return !_empty; });
    item = std::move(_item);
    _empty = true;
    _cv.notify_all();
  }

 private:
  T _item;
  bool _empty;
  bool _blocking_send;
  std::mutex _m;
  std::condition_variable _cv;
};

// cloud_msgs/msg/cloud_info.msg:1-13
struct cloud_info {
  double stamp = 0;
  std::vector<int32_t> startRingIndex, endRingIndex;
  float startOrientation = 0, endOrientation = 0, orientationDiff = 0;
  std::vector<uint8_t> segmentedCloudGroundFlag;
  std::vector<uint32_t> segmentedCloudColInd;
  std::vector<float> segmentedCloudRange;
};

// utility.h:64-70.  `shutdown` replaces the reference's "empty item + !ros::ok()" stop signal.
struct ProjectionOut {
  Cloud segmented_cloud, outlier_cloud;
  cloud_info seg_msg;
  bool shutdown = false;
  uint64_t frame = 0;
};

// utility.h:73-80; laser_odometry carries transformSum (rx, ry, rz, tx, ty, tz) instead of a quaternion
struct AssociationOut {
  Cloud cloud_outlier_last, cloud_corner_last, cloud_surf_last;
  float laser_odometry[6] = {0, 0, 0, 0, 0, 0};
  double stamp = 0;
  bool shutdown = false;
  uint64_t frame = 0;
};

// One sequence on the GPU (batch = 1) shared by the three stages.
class Device {
 public:
  explicit Device(const LegoLoamParams& params, int max_points = 0, int cuda_device = 0);
  ~Device();
  ll_handle* h() const { return _h; }
  std::mutex& mutex() { return _mtx; }
  const LegoLoamParams& params() const { return _params; }
  Cloud download_cloud(int buffer);
  template <typename T>
  std::vector<T> download(int buffer);
  void check(int rc, const char* what) const;
  // blocks until every frame handed to ImageProjection has left FeatureAssociation and MapOptimization is idle
  void waitIdle();
  // when true the stages also download their full payloads (segmented cloud, cloud_info arrays) to the host
  bool download_payloads = false;
  // A stage thread never throws (that would terminate the process): the first failure is recorded here, the stages
  // stop touching the device and only drain their channels; cloudHandler (the caller's thread) reports it.
  void fail(const std::string& what);
  bool failed() const { return _failed.load(); }
  std::string error();

 private:
  LegoLoamParams _params;
  ll_handle* _h = nullptr;
  std::mutex _mtx;
  std::atomic<bool> _failed{false};
  std::mutex _err_mtx;
  std::string _error;
};

class ImageProjection {
 public:
  ImageProjection(const LegoLoamParams& params, std::shared_ptr<Device> dev, Channel<ProjectionOut>& output_channel);
  ~ImageProjection() = default;
  // imageProjection.cpp:153-174; NaN points must already be removed by the caller
  void cloudHandler(const float* xyzi, int n_points, double stamp);
  // the same from a sensor_msgs/PointCloud2 (e.g. ll_bag_get_pointcloud2): pcl::fromROSMsg + removeNaNFromPointCloud
  // (imageProjection.cpp:159-161) run on the device from the message bytes
  void cloudHandler(const ll_pointcloud2_view& msg);

 private:
  template <class SetScans>
  void handle(double stamp, SetScans set_scans);
  std::shared_ptr<Device> _dev;
  Channel<ProjectionOut>& _output_channel;
  uint64_t _frame = 0;
};

class FeatureAssociation {
 public:
  FeatureAssociation(const LegoLoamParams& params, std::shared_ptr<Device> dev, Channel<ProjectionOut>& input_channel,
                     Channel<AssociationOut>& output_channel);
  ~FeatureAssociation();
  void runFeatureAssociation();  // featureAssociation.cpp:1386-1450
  void transformSum(float out6[6]);

 private:
  std::shared_ptr<Device> _dev;
  Channel<ProjectionOut>& _input_channel;
  Channel<AssociationOut>& _output_channel;
  std::thread _run_thread;
};

class MapOptimization {
 public:
  MapOptimization(const LegoLoamParams& params, std::shared_ptr<Device> dev, Channel<AssociationOut>& input_channel);
  ~MapOptimization();
  void run();  // mapOptmization.cpp:1521-1570
  size_t keyFrames() const { return _n_key_frames.load(); }
  size_t cycles() const { return _cycles.load(); }
  void transformAftMapped(float out6[6]);

 private:
  std::shared_ptr<Device> _dev;
  Channel<AssociationOut>& _input_channel;
  std::thread _run_thread;
  float _aft[6] = {0, 0, 0, 0, 0, 0};
  std::mutex _pose_mtx;
  std::atomic<size_t> _n_key_frames{0}, _cycles{0};
};

}  // namespace lego_loam
