// sequence_driver.cpp -- bag-less replacement of the reference's main.cpp bag loop (main.cpp:37-47,72-76):
// builds the two channels and the three stage objects and feeds a recorded sequence of scans to
// ImageProjection::cloudHandler, one message at a time, with blocking channels (the reference's
// deterministic rosbag mode).
//
//   sequence_driver <config: A|B|C|T> <scans.bin | recording.bag[:topic]> <poses.out> [--stream]
//
// --stream: the scans are pushed without waiting for the stages in between (the three stage threads overlap like in the
// reference's live mode); only the line of the last frame is written, after everything has drained.
//
// scans.bin: int32 n_frames, then per frame: int32 n_points, n_points * 4 float32 (x, y, z, intensity).
// recording.bag: a rosbag v2.0 file (main.cpp:26-35,60-76); the sensor_msgs/PointCloud2 messages of `topic` (default: the
// first such topic) go to cloudHandler as raw message bytes, decoded on the device.
// poses.out: per frame one text line: frame, transformSum[6], transformAftMapped[6], key frames, map cycles.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "lego_loam_host.h"

using namespace lego_loam;

static LegoLoamParams config(const char* name) {
  LegoLoamParams p;
  ll_default_params(&p);
  if (!strcmp(name, "B")) { p.num_vertical_scans = 32; p.ground_scan_index = 15; }
  else if (!strcmp(name, "C")) { p.num_vertical_scans = 64; p.num_horizontal_scans = 2048; p.ground_scan_index = 31; p.vertical_angle_bottom = -16.6f; p.vertical_angle_top = 16.6f; }
  else if (!strcmp(name, "T")) { p.num_horizontal_scans = 450; }
  return p;
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s <A|B|C|T> <scans.bin> <poses.out>\n", argv[0]); return 2; }
  const LegoLoamParams params = config(argv[1]);
  const bool stream_mode = argc > 4 && !strcmp(argv[4], "--stream");
  std::string in_path = argv[2], topic;
  const size_t bag_ext = in_path.find(".bag");
  ll_bag* bag = nullptr;
  if (bag_ext != std::string::npos) {
    if (bag_ext + 4 < in_path.size() && in_path[bag_ext + 4] == ':') { topic = in_path.substr(bag_ext + 5); in_path.resize(bag_ext + 4); }
    if (ll_bag_open(in_path.c_str(), topic.c_str(), &bag) != LL_OK) {
      fprintf(stderr, "Unable to open rosbag [%s]: %s\n", in_path.c_str(), ll_bag_last_error());  // main.cpp:32-33
      return 1;
    }
  }
  FILE* f = bag ? nullptr : fopen(argv[2], "rb");
  if (bag) {
    FILE* out = fopen(argv[3], "w");
    if (!out) return 1;
    try {
      std::shared_ptr<Device> dev(new Device(params));
      Channel<ProjectionOut> projection_out_channel(true);
      Channel<AssociationOut> association_out_channel(true);
      ImageProjection IP(params, dev, projection_out_channel);
      FeatureAssociation FA(params, dev, projection_out_channel, association_out_channel);
      MapOptimization MO(params, dev, association_out_channel);
      const int n_msgs = ll_bag_num_messages(bag);
      for (int i = 0; i < n_msgs; ++i) {
        ll_pointcloud2_view msg;
        if (ll_bag_get_pointcloud2(bag, i, &msg) != LL_OK) { fprintf(stderr, "message %d: %s\n", i, ll_bag_last_error()); break; }
        IP.cloudHandler(msg);
        dev->waitIdle();
        float ts[6], am[6];
        FA.transformSum(ts);
        MO.transformAftMapped(am);
        fprintf(out, "%d %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %zu %zu\n", i, ts[0], ts[1], ts[2], ts[3], ts[4],
                ts[5], am[0], am[1], am[2], am[3], am[4], am[5], MO.keyFrames(), MO.cycles());
      }
    } catch (const std::exception& e) {
      fprintf(stderr, "fatal: %s\n", e.what());
      return 1;
    }
    fclose(out);
    ll_bag_close(bag);
    return 0;
  }
  if (!f) { fprintf(stderr, "Unable to open [%s]\n", argv[2]); return 1; }  // main.cpp:32-33
  int32_t n_frames = 0;
  if (fread(&n_frames, 4, 1, f) != 1) return 1;
  FILE* out = fopen(argv[3], "w");
  if (!out) return 1;
  try {
    std::shared_ptr<Device> dev(new Device(params));
    Channel<ProjectionOut> projection_out_channel(true);
    Channel<AssociationOut> association_out_channel(true);  // bag mode: blocking (main.cpp:38)
    ImageProjection IP(params, dev, projection_out_channel);
    FeatureAssociation FA(params, dev, projection_out_channel, association_out_channel);
    MapOptimization MO(params, dev, association_out_channel);
    std::vector<float> scan;
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n_frames; ++i) {
      int32_t n = 0;
      if (fread(&n, 4, 1, f) != 1) break;
      scan.resize((size_t)n * 4);
      if (n && fread(scan.data(), 16, n, f) != (size_t)n) break;
      IP.cloudHandler(scan.data(), n, 0.1 * i);
      if (stream_mode && i + 1 < n_frames) continue;
      dev->waitIdle();
      if (dev->failed()) { fprintf(stderr, "fatal: %s\n", dev->error().c_str()); return 1; }
      float ts[6], am[6];
      FA.transformSum(ts);
      MO.transformAftMapped(am);
      fprintf(out, "%d %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %zu %zu\n", i, ts[0], ts[1], ts[2], ts[3], ts[4],
              ts[5], am[0], am[1], am[2], am[3], am[4], am[5], MO.keyFrames(), MO.cycles());
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "Entire sequence processed at %.1fX speed (%d scans, %.3f s)\n", 0.1 * n_frames / sec, n_frames, sec);  // main.cpp:99-102
  } catch (const std::exception& e) {
    fprintf(stderr, "fatal: %s\n", e.what());
    return 1;
  }
  fclose(out);
  fclose(f);
  return 0;
}
