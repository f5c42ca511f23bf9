/*
 * lego_oracle.cpp -- TEST INFRASTRUCTURE, not product code.  "parity unpinned" for the
 * arithmetic that lives in un-vendored Eigen / PCL / tf / libm (see DESIGN.md and
 * SURVEY.md section 8c); pinned by the hand-derivable known answers of SURVEY.md section 4 and, for
 * k-NN, by the reference's own vendored nanoflann.hpp (tests/test_oracle_pins.py).
 *
 * CPU restatement of the reference hot path, one sequence per object.  Paths below are
 * relative to /root/reference/LeGO-LOAM/src.  Precision / promotion follows SURVEY.md
 * section 10: float overloads of libm on float arguments, every expression that mixes in a
 * double literal or M_PI is evaluated in double and rounded on assignment.
 * Compiled with the reference's flags (-std=c++11 -O3 -g, CMakeLists.txt:4) plus
 * -ffp-contract=off (a no-op on x86-64 without -march).
 */
#include "lego_oracle.h"

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../include/ll_smallmat.h"
#include "oracle_knn.h"
#include "oracle_math.h"

namespace om { int g_use_libm = 0; }
static int g_float_accum = 0;
static int g_stable_sort = 0;
namespace oknn {
#ifdef ORACLE_WITH_NANOFLANN
int g_use_nanoflann = 1;
#else
int g_use_nanoflann = 0;
#endif
}

using om::asin_;
using om::atan2_;
using om::cos_;
using om::sin_;
using om::tan_;

namespace {

const double DEG_TO_RAD = M_PI / 180.0;          /* utility.h:50 */
const float RAD2DEG = 180.0 / M_PI;              /* featureAssociation.cpp:39 */

struct smoothness_t { float value; size_t ind; }; /* utility.h:53-56 */
struct by_value { bool operator()(smoothness_t const& l, smoothness_t const& r) { return l.value < r.value; } };

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

/* pcl::VoxelGrid<PointXYZI>::applyFilter restated (PCL is not vendored; SURVEY.md section 8 f1,
 * section 11.3): float inverse leaf, floor() of float products, ascending voxel index, centroid
 * of all four fields summed in float in input order, divided by the count. */
void voxel_grid(const std::vector<P4>& in, float leaf, std::vector<P4>& out) {
  out.clear();
  if (in.empty()) return;
  const float inv = 1.0f / leaf;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (const P4& p : in) {
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x);
    mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y);
    mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z);
  }
  const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1;
  const int64_t dy = (int64_t)((mx[1] - mn[1]) * inv) + 1;
  const int64_t dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)INT_MAX) { out = in; return; } /* PCL warns and copies the input */
  int minb[3], maxb[3], divb[3];
  for (int d = 0; d < 3; ++d) {
    minb[d] = (int)std::floor(mn[d] * inv);
    maxb[d] = (int)std::floor(mx[d] * inv);
    divb[d] = maxb[d] - minb[d] + 1;
  }
  const int mul1 = divb[0], mul2 = divb[0] * divb[1];
  std::vector<std::pair<int, int> > iv; /* (voxel idx, point idx) */
  iv.reserve(in.size());
  for (size_t k = 0; k < in.size(); ++k) {
    const P4& p = in[k];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    const int i0 = (int)std::floor(p.x * inv) - minb[0];
    const int i1 = (int)std::floor(p.y * inv) - minb[1];
    const int i2 = (int)std::floor(p.z * inv) - minb[2];
    iv.push_back(std::make_pair(i0 + i1 * mul1 + i2 * mul2, (int)k));
  }
  std::stable_sort(iv.begin(), iv.end(),
                   [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first < b.first; });
  size_t s = 0;
  while (s < iv.size()) {
    size_t e = s + 1;
    while (e < iv.size() && iv[e].first == iv[s].first) ++e;
    float cx = 0.f, cy = 0.f, cz = 0.f, ci = 0.f;
    for (size_t k = s; k < e; ++k) {
      const P4& p = in[iv[k].second];
      cx += p.x; cy += p.y; cz += p.z; ci += p.i;
    }
    const float cnt = (float)(e - s);
    P4 c; c.x = cx / cnt; c.y = cy / cnt; c.z = cz / cnt; c.i = ci / cnt;
    out.push_back(c);
    s = e;
  }
}

}  // namespace

struct lo_handle {
  LegoLoamParams prm;
  int V, H, N;
  /* ---- ImageProjection members (imageProjection.h:45-81) ---- */
  float ang_bottom, ang_res_x, ang_res_y, segment_theta, sensor_mount_angle;
  int gsi, seg_valid_point_num, seg_valid_line_num;
  std::vector<float> range_mat;
  std::vector<int8_t> ground_mat;
  std::vector<int> label_mat;
  int label_count;
  std::vector<P4> full_cloud, full_info_cloud;
  std::vector<P4> laser_cloud_in;
  /* ProjectionOut */
  std::vector<P4> segmented_cloud, outlier_cloud;
  std::vector<int> start_ring, end_ring;
  float start_ori, end_ori, ori_diff;
  std::vector<uint8_t> seg_ground_flag;
  std::vector<uint32_t> seg_col_ind;
  std::vector<float> seg_range;
  /* ---- FeatureAssociation members (featureAssociation.h:24-120) ---- */
  float scan_period, edge_threshold, surf_threshold, nearest_feature_dist_sqr;
  int mapping_frequency_div;
  std::vector<smoothness_t> cloudSmoothness;
  std::vector<float> cloudCurvature;
  std::vector<int> cloudNeighborPicked, cloudLabel;
  std::vector<P4> cornerSharp, cornerLessSharp, surfFlat, surfLessFlat;
  std::vector<int> cornerSharpInd, cornerLessSharpInd, surfFlatInd;
  std::vector<int> lessFlatRawCount;
  std::vector<P4> cornerLast, surfLast, outlierLast;
  int cornerLastNum, surfLastNum;
  std::vector<float> searchCornerInd1, searchCornerInd2, searchSurfInd1, searchSurfInd2, searchSurfInd3;
  float transformCur[6], transformSum[6];
  std::unique_ptr<oknn::KdTree> kdCornerLast, kdSurfLast;
  std::vector<P4> laserCloudOri, coeffSel;
  bool isDegenerate, systemInitedLM;
  float matP3[9];
  size_t cycle_count;
  int odom_iters[2];
  /* ---- MapOptimization scan-to-map slice (mapOptimization.h) ---- */
  std::vector<P4> mapCorner, mapSurf, scanCornerDS, scanSurfTotalDS;
  std::unique_ptr<oknn::KdTree> kdCornerMap, kdSurfMap;
  float transformTobeMapped[6], transformBefMapped[6], transformAftMapped[6], transformIncre[6];
  bool mapDegenerate;
  float matP6[36];
  int map_iters[2];
  double map_trace[340];
  std::vector<P4> moOri, moCoeff;
  double timers[5];
  /* index traces for parity tests: scan-to-scan correspondences of every search round (LM iterations 0, 5, 10, 15, 20)
   * as [stage][round][feature][3] (closest, ind2, ind3; -1 = none / not run), and the 5-NN of every scan-to-map
   * iteration as [iteration][query][5] (corner queries first; -1 x 5 when the fifth neighbour is not closer than 1 m) */
  std::vector<int> odomSearchTrace, mapKnnTrace;
  int stable_sort; /* 1: equal curvatures keep their position order in extractFeatures (see there) */
  int float_accum; /* 0: exact double accumulation of J^T J (default); 1: float, row by row; 2: float, 8 partial sums (SIMD-like) */

  explicit lo_handle(const LegoLoamParams& p) : prm(p) {
    V = p.num_vertical_scans; H = p.num_horizontal_scans; N = V * H;
    /* imageProjection.cpp:57-84 */
    float bottom = p.vertical_angle_bottom;
    const float top = p.vertical_angle_top;
    ang_res_x = (M_PI * 2) / (H);
    ang_res_y = DEG_TO_RAD * (top - bottom) / float(V - 1);
    ang_bottom = -(bottom - 0.1) * DEG_TO_RAD;
    segment_theta = p.segment_theta; segment_theta *= DEG_TO_RAD;
    seg_valid_point_num = p.segment_valid_point_num;
    seg_valid_line_num = p.segment_valid_line_num;
    gsi = p.ground_scan_index;
    sensor_mount_angle = p.sensor_mount_angle; sensor_mount_angle *= DEG_TO_RAD;
    full_cloud.resize(N); full_info_cloud.resize(N);
    /* featureAssociation.cpp:69-81 */
    scan_period = p.scan_period;
    edge_threshold = p.edge_threshold; surf_threshold = p.surf_threshold;
    mapping_frequency_div = p.mapping_frequency_divider;
    const float nd = p.nearest_feature_search_distance;
    nearest_feature_dist_sqr = nd * nd;
    float_accum = g_float_accum;
    stable_sort = g_stable_sort;
    reset();
  }

  /* a freshly constructed FeatureAssociation (featureAssociation.cpp:96-157 initializationValue); MapOptimization untouched */
  void resetFeatureAssociation() {
    cloudSmoothness.assign(N, smoothness_t{0.f, 0});
    cloudCurvature.assign(N, 0.f); cloudNeighborPicked.assign(N, 0); cloudLabel.assign(N, 0);
    searchCornerInd1.assign(N, 0.f); searchCornerInd2.assign(N, 0.f);
    searchSurfInd1.assign(N, 0.f); searchSurfInd2.assign(N, 0.f); searchSurfInd3.assign(N, 0.f);
    for (int i = 0; i < 6; ++i) { transformCur[i] = 0; transformSum[i] = 0; }
    systemInitedLM = false; isDegenerate = false; cycle_count = 0;
    cornerLast.clear(); surfLast.clear(); outlierLast.clear(); cornerLastNum = surfLastNum = 0;
    kdCornerLast.reset(new oknn::KdTree()); kdSurfLast.reset(new oknn::KdTree());
    for (int i = 0; i < 9; ++i) matP3[i] = 0.f;
    odom_iters[0] = odom_iters[1] = 0;
    lessFlatRawCount.assign(V, 0);
  }

  void reset() {
    resetFeatureAssociation();
    for (int i = 0; i < 6; ++i) {
      transformTobeMapped[i] = 0; transformBefMapped[i] = 0; transformAftMapped[i] = 0; transformIncre[i] = 0;
    }
    kdCornerMap.reset(new oknn::KdTree()); kdSurfMap.reset(new oknn::KdTree());
    for (int i = 0; i < 36; ++i) matP6[i] = 0.f; /* mapOptmization.cpp:223 matP.setZero() */
    mapDegenerate = false;
    map_iters[0] = map_iters[1] = 0;
    for (int i = 0; i < 340; ++i) map_trace[i] = 0.0;
    for (int i = 0; i < 5; ++i) timers[i] = 0;
    resetKeyFrames();
  }

  /* ================= ImageProjection ================= */

  void resetParameters() { /* imageProjection.cpp:107-150 */
    P4 nanPoint;
    nanPoint.x = nanPoint.y = nanPoint.z = std::numeric_limits<float>::quiet_NaN();
    nanPoint.i = 0.f; /* pcl::PointXYZI() default intensity */
    segmented_cloud.clear(); outlier_cloud.clear();
    range_mat.assign(N, FLT_MAX); ground_mat.assign(N, 0); label_mat.assign(N, 0);
    label_count = 1;
    std::fill(full_cloud.begin(), full_cloud.end(), nanPoint);
    std::fill(full_info_cloud.begin(), full_info_cloud.end(), nanPoint);
    start_ring.assign(V, 0); end_ring.assign(V, 0);
    seg_ground_flag.assign(N, 0); seg_col_ind.assign(N, 0); seg_range.assign(N, 0.f);
  }

  void findStartEndAngle() { /* imageProjection.cpp:234-249 */
    if (laser_cloud_in.empty()) { start_ori = end_ori = ori_diff = 0.f; return; } /* reference: UB on empty input */
    P4 point = laser_cloud_in.front();
    start_ori = -atan2_(point.y, point.x);
    point = laser_cloud_in.back();
    end_ori = -atan2_(point.y, point.x) + 2 * M_PI;
    if (end_ori - start_ori > 3 * M_PI) {
      end_ori -= 2 * M_PI;
    } else if (end_ori - start_ori < M_PI) {
      end_ori += 2 * M_PI;
    }
    ori_diff = end_ori - start_ori;
  }

  void projectPointCloud() { /* imageProjection.cpp:178-224 */
    const size_t cloudSize = laser_cloud_in.size();
    for (size_t i = 0; i < cloudSize; ++i) {
      P4 thisPoint = laser_cloud_in[i];
      float range = sqrtf(thisPoint.x * thisPoint.x + thisPoint.y * thisPoint.y + thisPoint.z * thisPoint.z);
      float verticalAngle = asin_(thisPoint.z / range);
      const float rowf = (verticalAngle + ang_bottom) / ang_res_y;
      if (!(rowf == rowf)) continue;              /* x86 cvttss2si(NaN) = INT_MIN, rejected by the test below */
      if (!(rowf > -2147483648.f && rowf < 2147483648.f)) continue; /* out-of-range conversions are INT_MIN too */
      int rowIdn = rowf;
      if (rowIdn < 0 || rowIdn >= V) continue;
      float horizonAngle = atan2_(thisPoint.x, thisPoint.y);
      int columnIdn = -round((horizonAngle - M_PI_2) / ang_res_x) + H * 0.5;
      if (columnIdn >= H) columnIdn -= H;
      if (columnIdn < 0 || columnIdn >= H) continue;
      if (range < 0.1) continue;
      range_mat[columnIdn + rowIdn * H] = range;
      thisPoint.i = (float)rowIdn + (float)columnIdn / 10000.0;
      size_t index = columnIdn + rowIdn * H;
      full_cloud[index] = thisPoint;
      full_info_cloud[index] = thisPoint;
      full_info_cloud[index].i = range;
    }
  }

  void groundRemoval() { /* imageProjection.cpp:254-308 */
    for (int j = 0; j < H; ++j) {
      for (int i = 0; i < gsi && i + 1 < V; ++i) {
        size_t lowerInd = j + (i)*H;
        size_t upperInd = j + (i + 1) * H;
        if (full_cloud[lowerInd].i == -1 || full_cloud[upperInd].i == -1) {
          ground_mat[lowerInd] = -1;
          continue;
        }
        float dX = full_cloud[upperInd].x - full_cloud[lowerInd].x;
        float dY = full_cloud[upperInd].y - full_cloud[lowerInd].y;
        float dZ = full_cloud[upperInd].z - full_cloud[lowerInd].z;
        float vertical_angle = atan2_(dZ, sqrtf(dX * dX + dY * dY + dZ * dZ));
        if ((vertical_angle - sensor_mount_angle) <= 10 * DEG_TO_RAD) {
          ground_mat[lowerInd] = 1;
          ground_mat[upperInd] = 1;
        }
      }
    }
    for (int i = 0; i < V; ++i)
      for (int j = 0; j < H; ++j)
        if (ground_mat[j + i * H] == 1 || range_mat[j + i * H] == FLT_MAX) label_mat[j + i * H] = -1;
  }

  void labelComponents(int row, int col) { /* imageProjection.cpp:412-496 */
    const float segmentThetaThreshold = tan_(segment_theta);
    std::vector<bool> lineCountFlag(V, false);
    /* two boost::circular_buffer<Vector2i>(N): allocated (not initialised) per seed, as the reference does */
    std::unique_ptr<int[]> queue(new int[2 * (size_t)N]);
    std::unique_ptr<int[]> all_pushed(new int[2 * (size_t)N]);
    size_t q_head = 0, q_tail = 0, n_pushed = 0;
    queue[0] = row; queue[1] = col; q_tail = 1;
    all_pushed[0] = row; all_pushed[1] = col; n_pushed = 1;
    const int nb[4][2] = {{0, -1}, {-1, 0}, {1, 0}, {0, 1}};
    while (q_tail > q_head) {
      const int fx = queue[2 * q_head], fy = queue[2 * q_head + 1];
      ++q_head;
      label_mat[fy + fx * H] = label_count;
      for (int k = 0; k < 4; ++k) {
        int thisIndX = fx + nb[k][0];
        int thisIndY = fy + nb[k][1];
        if (thisIndX < 0 || thisIndX >= V) continue;
        if (thisIndY < 0) thisIndY = H - 1;
        if (thisIndY >= H) thisIndY = 0;
        if (label_mat[thisIndY + thisIndX * H] != 0) continue;
        float d1 = std::max(range_mat[fy + fx * H], range_mat[thisIndY + thisIndX * H]);
        float d2 = std::min(range_mat[fy + fx * H], range_mat[thisIndY + thisIndX * H]);
        float alpha = (nb[k][0] == 0) ? ang_res_x : ang_res_y;
        float tang = (d2 * sin_(alpha) / (d1 - d2 * cos_(alpha)));
        if (tang > segmentThetaThreshold) {
          queue[2 * q_tail] = thisIndX; queue[2 * q_tail + 1] = thisIndY; ++q_tail;
          label_mat[thisIndY + thisIndX * H] = label_count;
          lineCountFlag[thisIndX] = true;
          all_pushed[2 * n_pushed] = thisIndX; all_pushed[2 * n_pushed + 1] = thisIndY; ++n_pushed;
        }
      }
    }
    bool feasibleSegment = false;
    if (n_pushed >= 30) {
      feasibleSegment = true;
    } else if ((int)n_pushed >= seg_valid_point_num) {
      int lineCount = 0;
      for (int i = 0; i < V; ++i)
        if (lineCountFlag[i] == true) ++lineCount;
      if (lineCount >= seg_valid_line_num) feasibleSegment = true;
    }
    if (feasibleSegment == true) {
      ++label_count;
    } else {
      for (size_t i = 0; i < n_pushed; ++i) label_mat[all_pushed[2 * i + 1] + all_pushed[2 * i] * H] = 999999;
    }
  }

  void cloudSegmentation() { /* imageProjection.cpp:352-396 */
    for (int i = 0; i < V; ++i)
      for (int j = 0; j < H; ++j)
        if (label_mat[j + i * H] == 0) labelComponents(i, j);
    int sizeOfSegCloud = 0;
    for (int i = 0; i < V; ++i) {
      start_ring[i] = sizeOfSegCloud - 1 + 5;
      for (int j = 0; j < H; ++j) {
        if (label_mat[j + i * H] > 0 || ground_mat[j + i * H] == 1) {
          if (label_mat[j + i * H] == 999999) {
            if (i > gsi && j % 5 == 0) {
              outlier_cloud.push_back(full_cloud[j + i * H]);
              continue;
            } else {
              continue;
            }
          }
          if (ground_mat[j + i * H] == 1) {
            if (j % 5 != 0 && j > 5 && j < H - 5) continue;
          }
          seg_ground_flag[sizeOfSegCloud] = (ground_mat[j + i * H] == 1);
          seg_col_ind[sizeOfSegCloud] = j;
          seg_range[sizeOfSegCloud] = range_mat[j + i * H];
          segmented_cloud.push_back(full_cloud[j + i * H]);
          ++sizeOfSegCloud;
        }
      }
      end_ring[i] = sizeOfSegCloud - 1 - 5;
    }
  }

  void cloudHandler(const float* xyzi, int n) { /* imageProjection.cpp:153-174 */
    resetParameters();
    laser_cloud_in.resize(n);
    if (n) std::memcpy(laser_cloud_in.data(), xyzi, sizeof(P4) * (size_t)n);
    findStartEndAngle();
    projectPointCloud();
    groundRemoval();
    cloudSegmentation();
  }

  /* ================= FeatureAssociation ================= */

  void adjustDistortion() { /* featureAssociation.cpp:161-197 */
    bool halfPassed = false;
    int cloudSize = segmented_cloud.size();
    P4 point;
    for (int i = 0; i < cloudSize; i++) {
      point.x = segmented_cloud[i].y;
      point.y = segmented_cloud[i].z;
      point.z = segmented_cloud[i].x;
      float ori = -atan2_(point.x, point.z);
      if (!halfPassed) {
        if (ori < start_ori - M_PI / 2)
          ori += 2 * M_PI;
        else if (ori > start_ori + M_PI * 3 / 2)
          ori -= 2 * M_PI;
        if (ori - start_ori > M_PI) halfPassed = true;
      } else {
        ori += 2 * M_PI;
        if (ori < end_ori - M_PI * 3 / 2)
          ori += 2 * M_PI;
        else if (ori > end_ori + M_PI / 2)
          ori -= 2 * M_PI;
      }
      float relTime = (ori - start_ori) / ori_diff;
      point.i = int(segmented_cloud[i].i) + scan_period * relTime;
      segmented_cloud[i] = point;
    }
  }

  void calculateSmoothness() { /* featureAssociation.cpp:200-223 */
    int cloudSize = segmented_cloud.size();
    const std::vector<float>& r = seg_range;
    for (int i = 5; i < cloudSize - 5; i++) {
      float diffRange = r[i - 5] + r[i - 4] + r[i - 3] + r[i - 2] + r[i - 1] - r[i] * 10 + r[i + 1] + r[i + 2] +
                        r[i + 3] + r[i + 4] + r[i + 5];
      cloudCurvature[i] = diffRange * diffRange;
      cloudNeighborPicked[i] = 0;
      cloudLabel[i] = 0;
      cloudSmoothness[i].value = cloudCurvature[i];
      cloudSmoothness[i].ind = i;
    }
  }

  void markOccludedPoints() { /* featureAssociation.cpp:226-262 */
    int cloudSize = segmented_cloud.size();
    const std::vector<float>& r = seg_range;
    for (int i = 5; i < cloudSize - 6; ++i) {
      float depth1 = r[i];
      float depth2 = r[i + 1];
      int columnDiff = std::abs(int(seg_col_ind[i + 1] - seg_col_ind[i]));
      if (columnDiff < 10) {
        if (depth1 - depth2 > 0.3) {
          cloudNeighborPicked[i - 5] = 1; cloudNeighborPicked[i - 4] = 1; cloudNeighborPicked[i - 3] = 1;
          cloudNeighborPicked[i - 2] = 1; cloudNeighborPicked[i - 1] = 1; cloudNeighborPicked[i] = 1;
        } else if (depth2 - depth1 > 0.3) {
          cloudNeighborPicked[i + 1] = 1; cloudNeighborPicked[i + 2] = 1; cloudNeighborPicked[i + 3] = 1;
          cloudNeighborPicked[i + 4] = 1; cloudNeighborPicked[i + 5] = 1; cloudNeighborPicked[i + 6] = 1;
        }
      }
      float diff1 = std::abs(r[i - 1] - r[i]);
      float diff2 = std::abs(r[i + 1] - r[i]);
      if (diff1 > 0.02 * r[i] && diff2 > 0.02 * r[i]) cloudNeighborPicked[i] = 1;
    }
  }

  void extractFeatures() { /* featureAssociation.cpp:265-383 */
    cornerSharp.clear(); cornerLessSharp.clear(); surfFlat.clear(); surfLessFlat.clear();
    cornerSharpInd.clear(); cornerLessSharpInd.clear(); surfFlatInd.clear();
    std::vector<P4> lessFlatScan, lessFlatScanDS;
    const size_t colIndSize = seg_col_ind.size(); /* == N, imageProjection.cpp:138 */
    for (int i = 0; i < V; i++) {
      lessFlatScan.clear();
      for (int j = 0; j < 6; j++) {
        int sp = (start_ring[i] * (6 - j) + end_ring[i] * j) / 6;
        int ep = (start_ring[i] * (5 - j) + end_ring[i] * (j + 1)) / 6 - 1;
        if (sp >= ep) continue;
        /* std::sort leaves the order of equal curvatures unspecified (libstdc++'s introsort decides); stable_sort is one
         * of the permitted outcomes and the one a (value, position) sort on the device produces */
        if (stable_sort) std::stable_sort(cloudSmoothness.begin() + sp, cloudSmoothness.begin() + ep, by_value());
        else std::sort(cloudSmoothness.begin() + sp, cloudSmoothness.begin() + ep, by_value());
        int largestPickedNum = 0;
        for (int k = ep; k >= sp; k--) {
          int ind = cloudSmoothness[k].ind;
          if (cloudNeighborPicked[ind] == 0 && cloudCurvature[ind] > edge_threshold && seg_ground_flag[ind] == false) {
            largestPickedNum++;
            if (largestPickedNum <= 2) {
              cloudLabel[ind] = 2;
              cornerSharp.push_back(segmented_cloud[ind]); cornerSharpInd.push_back(ind);
              cornerLessSharp.push_back(segmented_cloud[ind]); cornerLessSharpInd.push_back(ind);
            } else if (largestPickedNum <= 20) {
              cloudLabel[ind] = 1;
              cornerLessSharp.push_back(segmented_cloud[ind]); cornerLessSharpInd.push_back(ind);
            } else {
              break;
            }
            cloudNeighborPicked[ind] = 1;
            for (int l = 1; l <= 5; l++) {
              if ((size_t)(ind + l) >= colIndSize) continue;
              int columnDiff = std::abs(int(seg_col_ind[ind + l] - seg_col_ind[ind + l - 1]));
              if (columnDiff > 10) break;
              cloudNeighborPicked[ind + l] = 1;
            }
            for (int l = -1; l >= -5; l--) {
              if (ind + l < 0) continue;
              int columnDiff = std::abs(int(seg_col_ind[ind + l] - seg_col_ind[ind + l + 1]));
              if (columnDiff > 10) break;
              cloudNeighborPicked[ind + l] = 1;
            }
          }
        }
        int smallestPickedNum = 0;
        for (int k = sp; k <= ep; k++) {
          int ind = cloudSmoothness[k].ind;
          if (cloudNeighborPicked[ind] == 0 && cloudCurvature[ind] < surf_threshold && seg_ground_flag[ind] == true) {
            cloudLabel[ind] = -1;
            surfFlat.push_back(segmented_cloud[ind]); surfFlatInd.push_back(ind);
            smallestPickedNum++;
            if (smallestPickedNum >= 4) break;
            cloudNeighborPicked[ind] = 1;
            for (int l = 1; l <= 5; l++) {
              if ((size_t)(ind + l) >= colIndSize) continue;
              int columnDiff = std::abs(int(seg_col_ind[ind + l] - seg_col_ind[ind + l - 1]));
              if (columnDiff > 10) break;
              cloudNeighborPicked[ind + l] = 1;
            }
            for (int l = -1; l >= -5; l--) {
              if (ind + l < 0) continue;
              int columnDiff = std::abs(int(seg_col_ind[ind + l] - seg_col_ind[ind + l + 1]));
              if (columnDiff > 10) break;
              cloudNeighborPicked[ind + l] = 1;
            }
          }
        }
        for (int k = sp; k <= ep; k++)
          if (cloudLabel[k] <= 0) lessFlatScan.push_back(segmented_cloud[k]);
      }
      lessFlatRawCount[i] = (int)lessFlatScan.size();
      voxel_grid(lessFlatScan, 0.2f, lessFlatScanDS); /* downSizeFilter.setLeafSize(0.2,...) featureAssociation.cpp:101 */
      surfLessFlat.insert(surfLessFlat.end(), lessFlatScanDS.begin(), lessFlatScanDS.end());
    }
  }

  void TransformToStart(const P4* pi, P4* po) { /* featureAssociation.cpp:388-418 */
    float s = 10 * (pi->i - int(pi->i));
    float ry = s * transformCur[1];
    float rx = s * transformCur[0];
    float rz = s * transformCur[2];
    float tx = s * transformCur[3];
    float ty = s * transformCur[4];
    float tz = s * transformCur[5];
    float x1 = cos_(rz) * (pi->x - tx) + sin_(rz) * (pi->y - ty);
    float y1 = -sin_(rz) * (pi->x - tx) + cos_(rz) * (pi->y - ty);
    float z1 = (pi->z - tz);
    float x2 = x1;
    float y2 = cos_(rx) * y1 + sin_(rx) * z1;
    float z2 = -sin_(rx) * y1 + cos_(rx) * z1;
    po->x = cos_(ry) * x2 - sin_(ry) * z2;
    po->y = y2;
    po->z = sin_(ry) * x2 + cos_(ry) * z2;
    po->i = pi->i;
  }

  void TransformToEnd(const P4* pi, P4* po) { /* featureAssociation.cpp:422-471 */
    float s = 10 * (pi->i - int(pi->i));
    float rx = s * transformCur[0];
    float ry = s * transformCur[1];
    float rz = s * transformCur[2];
    float tx = s * transformCur[3];
    float ty = s * transformCur[4];
    float tz = s * transformCur[5];
    float x1 = cos_(rz) * (pi->x - tx) + sin_(rz) * (pi->y - ty);
    float y1 = -sin_(rz) * (pi->x - tx) + cos_(rz) * (pi->y - ty);
    float z1 = (pi->z - tz);
    float x2 = x1;
    float y2 = cos_(rx) * y1 + sin_(rx) * z1;
    float z2 = -sin_(rx) * y1 + cos_(rx) * z1;
    float x3 = cos_(ry) * x2 - sin_(ry) * z2;
    float y3 = y2;
    float z3 = sin_(ry) * x2 + cos_(ry) * z2;
    rx = transformCur[0]; ry = transformCur[1]; rz = transformCur[2];
    tx = transformCur[3]; ty = transformCur[4]; tz = transformCur[5];
    float x4 = cos_(ry) * x3 + sin_(ry) * z3;
    float y4 = y3;
    float z4 = -sin_(ry) * x3 + cos_(ry) * z3;
    float x5 = x4;
    float y5 = cos_(rx) * y4 - sin_(rx) * z4;
    float z5 = sin_(rx) * y4 + cos_(rx) * z4;
    float x6 = cos_(rz) * x5 - sin_(rz) * y5 + tx;
    float y6 = sin_(rz) * x5 + cos_(rz) * y5 + ty;
    float z6 = z5 + tz;
    const float inten = int(pi->i);
    po->x = x6; po->y = y6; po->z = z6; po->i = inten;
  }

  void AccumulateRotation(float cx, float cy, float cz, float lx, float ly, float lz, float& ox, float& oy,
                          float& oz) { /* featureAssociation.cpp:474-500 */
    float srx = cos_(lx) * cos_(cx) * sin_(ly) * sin_(cz) - cos_(cx) * cos_(cz) * sin_(lx) - cos_(lx) * cos_(ly) * sin_(cx);
    ox = -asin_(srx);
    float srycrx = sin_(lx) * (cos_(cy) * sin_(cz) - cos_(cz) * sin_(cx) * sin_(cy)) +
                   cos_(lx) * sin_(ly) * (cos_(cy) * cos_(cz) + sin_(cx) * sin_(cy) * sin_(cz)) +
                   cos_(lx) * cos_(ly) * cos_(cx) * sin_(cy);
    float crycrx = cos_(lx) * cos_(ly) * cos_(cx) * cos_(cy) -
                   cos_(lx) * sin_(ly) * (cos_(cz) * sin_(cy) - cos_(cy) * sin_(cx) * sin_(cz)) -
                   sin_(lx) * (sin_(cy) * sin_(cz) + cos_(cy) * cos_(cz) * sin_(cx));
    oy = atan2_(srycrx / cos_(ox), crycrx / cos_(ox));
    float srzcrx = sin_(cx) * (cos_(lz) * sin_(ly) - cos_(ly) * sin_(lx) * sin_(lz)) +
                   cos_(cx) * sin_(cz) * (cos_(ly) * cos_(lz) + sin_(lx) * sin_(ly) * sin_(lz)) +
                   cos_(lx) * cos_(cx) * cos_(cz) * sin_(lz);
    float crzcrx = cos_(lx) * cos_(lz) * cos_(cx) * cos_(cz) -
                   cos_(cx) * sin_(cz) * (cos_(ly) * sin_(lz) - cos_(lz) * sin_(lx) * sin_(ly)) -
                   sin_(cx) * (sin_(ly) * sin_(lz) + cos_(ly) * cos_(lz) * sin_(lx));
    oz = atan2_(srzcrx / cos_(ox), crzcrx / cos_(ox));
  }

  static float sqd(const P4& a, const P4& b) {
    return (a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z);
  }

  void findCorrespondingCornerFeatures(int iterCount) { /* featureAssociation.cpp:503-637 */
    int cornerPointsSharpNum = cornerSharp.size();
    const int lastN = (int)cornerLast.size();
    for (int i = 0; i < cornerPointsSharpNum; i++) {
      P4 pointSel;
      TransformToStart(&cornerSharp[i], &pointSel);
      if (iterCount % 5 == 0) {
        int sidx[1]; float sd2[1];
        kdCornerLast->nearestKSearch(pointSel, 1, sidx, sd2);
        int closestPointInd = -1, minPointInd2 = -1;
        if (sd2[0] < nearest_feature_dist_sqr && sidx[0] < lastN) { /* 2nd test: reference reads out of bounds */
          closestPointInd = sidx[0];
          int closestPointScan = int(cornerLast[closestPointInd].i);
          float pointSqDis, minPointSqDis2 = nearest_feature_dist_sqr;
          /* loop bound is the CURRENT frame's sharp count (sic); min() only avoids the reference's out-of-bounds read */
          for (int j = closestPointInd + 1; j < cornerPointsSharpNum && j < lastN; j++) {
            if (int(cornerLast[j].i) > closestPointScan + 2.5) break;
            pointSqDis = sqd(cornerLast[j], pointSel);
            if (int(cornerLast[j].i) > closestPointScan) {
              if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
            }
          }
          for (int j = closestPointInd - 1; j >= 0; j--) {
            if (int(cornerLast[j].i) < closestPointScan - 2.5) break;
            pointSqDis = sqd(cornerLast[j], pointSel);
            if (int(cornerLast[j].i) < closestPointScan) {
              if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
            }
          }
        }
        searchCornerInd1[i] = closestPointInd;
        searchCornerInd2[i] = minPointInd2;
        traceOdom(1, iterCount / 5, i, closestPointInd, minPointInd2, -1);
      }
      if (searchCornerInd2[i] >= 0) {
        P4 tripod1 = cornerLast[(int)searchCornerInd1[i]];
        P4 tripod2 = cornerLast[(int)searchCornerInd2[i]];
        float x0 = pointSel.x, y0 = pointSel.y, z0 = pointSel.z;
        float x1 = tripod1.x, y1 = tripod1.y, z1 = tripod1.z;
        float x2 = tripod2.x, y2 = tripod2.y, z2 = tripod2.z;
        float m11 = ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1));
        float m22 = ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1));
        float m33 = ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1));
        float a012 = sqrtf(m11 * m11 + m22 * m22 + m33 * m33);
        float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
        float la = ((y1 - y2) * m11 + (z1 - z2) * m22) / a012 / l12;
        float lb = -((x1 - x2) * m11 - (z1 - z2) * m33) / a012 / l12;
        float lc = -((x1 - x2) * m22 + (y1 - y2) * m33) / a012 / l12;
        float ld2 = a012 / l12;
        float s = 1;
        if (iterCount >= 5) s = 1 - 1.8 * fabsf(ld2);
        if (s > 0.1 && ld2 != 0) {
          P4 coeff;
          coeff.x = s * la; coeff.y = s * lb; coeff.z = s * lc; coeff.i = s * ld2;
          laserCloudOri.push_back(cornerSharp[i]);
          coeffSel.push_back(coeff);
        }
      }
    }
  }

  void findCorrespondingSurfFeatures(int iterCount) { /* featureAssociation.cpp:640-779 */
    int surfPointsFlatNum = surfFlat.size();
    const int lastN = (int)surfLast.size();
    for (int i = 0; i < surfPointsFlatNum; i++) {
      P4 pointSel;
      TransformToStart(&surfFlat[i], &pointSel);
      if (iterCount % 5 == 0) {
        int sidx[1]; float sd2[1];
        kdSurfLast->nearestKSearch(pointSel, 1, sidx, sd2);
        int closestPointInd = -1, minPointInd2 = -1, minPointInd3 = -1;
        if (sd2[0] < nearest_feature_dist_sqr && sidx[0] < lastN) {
          closestPointInd = sidx[0];
          int closestPointScan = int(surfLast[closestPointInd].i);
          float pointSqDis, minPointSqDis2 = nearest_feature_dist_sqr, minPointSqDis3 = nearest_feature_dist_sqr;
          for (int j = closestPointInd + 1; j < surfPointsFlatNum && j < lastN; j++) {
            if (int(surfLast[j].i) > closestPointScan + 2.5) break;
            pointSqDis = sqd(surfLast[j], pointSel);
            if (int(surfLast[j].i) <= closestPointScan) {
              if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
            } else {
              if (pointSqDis < minPointSqDis3) { minPointSqDis3 = pointSqDis; minPointInd3 = j; }
            }
          }
          for (int j = closestPointInd - 1; j >= 0; j--) {
            if (int(surfLast[j].i) < closestPointScan - 2.5) break;
            pointSqDis = sqd(surfLast[j], pointSel);
            if (int(surfLast[j].i) >= closestPointScan) {
              if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
            } else {
              if (pointSqDis < minPointSqDis3) { minPointSqDis3 = pointSqDis; minPointInd3 = j; }
            }
          }
        }
        searchSurfInd1[i] = closestPointInd;
        traceOdom(0, iterCount / 5, i, closestPointInd, minPointInd2, minPointInd3);
        searchSurfInd2[i] = minPointInd2;
        searchSurfInd3[i] = minPointInd3;
      }
      if (searchSurfInd2[i] >= 0 && searchSurfInd3[i] >= 0) {
        P4 tripod1 = surfLast[(int)searchSurfInd1[i]];
        P4 tripod2 = surfLast[(int)searchSurfInd2[i]];
        P4 tripod3 = surfLast[(int)searchSurfInd3[i]];
        float pa = (tripod2.y - tripod1.y) * (tripod3.z - tripod1.z) - (tripod3.y - tripod1.y) * (tripod2.z - tripod1.z);
        float pb = (tripod2.z - tripod1.z) * (tripod3.x - tripod1.x) - (tripod3.z - tripod1.z) * (tripod2.x - tripod1.x);
        float pc = (tripod2.x - tripod1.x) * (tripod3.y - tripod1.y) - (tripod3.x - tripod1.x) * (tripod2.y - tripod1.y);
        float pd = -(pa * tripod1.x + pb * tripod1.y + pc * tripod1.z);
        float ps = sqrtf(pa * pa + pb * pb + pc * pc);
        pa /= ps; pb /= ps; pc /= ps; pd /= ps;
        float pd2 = pa * pointSel.x + pb * pointSel.y + pc * pointSel.z + pd;
        float s = 1;
        if (iterCount >= 5) {
          s = 1 - 1.8 * fabsf(pd2) /
                      sqrtf(sqrtf(pointSel.x * pointSel.x + pointSel.y * pointSel.y + pointSel.z * pointSel.z));
        }
        if (s > 0.1 && pd2 != 0) {
          P4 coeff;
          coeff.x = s * pa; coeff.y = s * pb; coeff.z = s * pc; coeff.i = s * pd2;
          laserCloudOri.push_back(surfFlat[i]);
          coeffSel.push_back(coeff);
        }
      }
    }
  }

  /* AtA = At*A and AtB = At*B (featureAssociation.cpp:860-863, mapOptmization.cpp:1257-1259).  Eigen's
   * float GEMM summation order is not reproducible; restated as exact double products accumulated in
   * double, rounded to float once (closer to the true value than any float order). */
  template <int C>
  void normal_equations(const std::vector<float>& A, const std::vector<float>& B, int n, float* AtA, float* AtB) {
    if (float_accum) {
      /* what the reference's Eigen float GEMM does up to its (unreproducible) summation order: float products summed in
       * float -- row by row (1) or as eight interleaved partial sums added pairwise at the end, like a SIMD kernel (2).
       * tests/test_oracle_pins.py shows that poses stay inside the parity tolerance and no LM exit flips. */
      const int L = float_accum == 2 ? 8 : 1;
      float acc[8][C * C], accb[8][C];
      for (int l = 0; l < L; ++l) { for (int i = 0; i < C * C; ++i) acc[l][i] = 0.f; for (int i = 0; i < C; ++i) accb[l][i] = 0.f; }
      for (int r = 0; r < n; ++r) {
        const int l = r % L;
        for (int a = 0; a < C; ++a) {
          for (int b = a; b < C; ++b) acc[l][a * C + b] += A[r * C + a] * A[r * C + b];
          accb[l][a] += A[r * C + a] * B[r];
        }
      }
      for (int step = 1; step < L; step *= 2)
        for (int l = 0; l + step < L; l += 2 * step) {
          for (int i = 0; i < C * C; ++i) acc[l][i] += acc[l + step][i];
          for (int i = 0; i < C; ++i) accb[l][i] += accb[l + step][i];
        }
      for (int a = 0; a < C; ++a) {
        for (int b = a; b < C; ++b) AtA[a * C + b] = AtA[b * C + a] = acc[0][a * C + b];
        AtB[a] = accb[0][a];
      }
      return;
    }
    double acc[C * C], accb[C];
    for (int i = 0; i < C * C; ++i) acc[i] = 0.0;
    for (int i = 0; i < C; ++i) accb[i] = 0.0;
    for (int r = 0; r < n; ++r) {
      for (int a = 0; a < C; ++a) {
        for (int b = a; b < C; ++b) acc[a * C + b] += (double)A[r * C + a] * (double)A[r * C + b];
        accb[a] += (double)A[r * C + a] * (double)B[r];
      }
    }
    for (int a = 0; a < C; ++a) {
      for (int b = a; b < C; ++b) AtA[a * C + b] = AtA[b * C + a] = (float)acc[a * C + b];
      AtB[a] = (float)accb[a];
    }
  }

  /* parity aid: the correspondences of search round `round` of stage (0 surf, 1 corner) */
  void traceOdom(int stage, int round, int i, int a, int b, int c) {
    const int cap = 24 * V;
    if (round < 0 || round >= 5 || i < 0 || i >= cap) return;
    int* t = odomSearchTrace.data() + (((size_t)stage * 5 + round) * cap + i) * 3;
    t[0] = a; t[1] = b; t[2] = c;
  }
  void traceKnn(int iter, int q, const int* ind, const float* dis) {
    const size_t Q = scanCornerDS.size() + scanSurfTotalDS.size();
    if (mapKnnTrace.size() != 10 * Q * 5) return;
    int* t = mapKnnTrace.data() + ((size_t)iter * Q + q) * 5;
    for (int j = 0; j < 5; ++j) t[j] = dis[4] < 1.0 ? ind[j] : -1;
  }

  bool calculateTransformationSurf(int iterCount) { /* featureAssociation.cpp:785-921 */
    int pointSelNum = laserCloudOri.size();
    std::vector<float> matA(pointSelNum * 3), matB(pointSelNum);
    float srx = sin_(transformCur[0]); float crx = cos_(transformCur[0]);
    float sry = sin_(transformCur[1]); float cry = cos_(transformCur[1]);
    float srz = sin_(transformCur[2]); float crz = cos_(transformCur[2]);
    float tx = transformCur[3]; float ty = transformCur[4]; float tz = transformCur[5];
    float a1 = crx * sry * srz; float a2 = crx * crz * sry; float a3 = srx * sry;
    float a4 = tx * a1 - ty * a2 - tz * a3;
    float a5 = srx * srz; float a6 = crz * srx;
    float a7 = ty * a6 - tz * crx - tx * a5;
    float a8 = crx * cry * srz; float a9 = crx * cry * crz; float a10 = cry * srx;
    float a11 = tz * a10 + ty * a9 - tx * a8;
    float b1 = -crz * sry - cry * srx * srz;
    float b2 = cry * crz * srx - sry * srz;
    float b5 = cry * crz - srx * sry * srz;
    float b6 = cry * srz + crz * srx * sry;
    float c1 = -b6; float c2 = b5;
    float c3 = tx * b6 - ty * b5;
    float c4 = -crx * crz; float c5 = crx * srz;
    float c6 = ty * c5 + tx * -c4;
    float c7 = b2; float c8 = -b1;
    float c9 = tx * -b2 - ty * -b1;
    for (int i = 0; i < pointSelNum; i++) {
      P4 pointOri = laserCloudOri[i];
      P4 coeff = coeffSel[i];
      float arx = (-a1 * pointOri.x + a2 * pointOri.y + a3 * pointOri.z + a4) * coeff.x +
                  (a5 * pointOri.x - a6 * pointOri.y + crx * pointOri.z + a7) * coeff.y +
                  (a8 * pointOri.x - a9 * pointOri.y - a10 * pointOri.z + a11) * coeff.z;
      float arz = (c1 * pointOri.x + c2 * pointOri.y + c3) * coeff.x + (c4 * pointOri.x - c5 * pointOri.y + c6) * coeff.y +
                  (c7 * pointOri.x + c8 * pointOri.y + c9) * coeff.z;
      float aty = -b6 * coeff.x + c4 * coeff.y + b2 * coeff.z;
      float d2 = coeff.i;
      matA[i * 3 + 0] = arx; matA[i * 3 + 1] = arz; matA[i * 3 + 2] = aty;
      matB[i] = -0.05 * d2;
    }
    float AtA[9], AtB[3], X[3], AtAc[9];
    normal_equations<3>(matA, matB, pointSelNum, AtA, AtB);
    for (int i = 0; i < 9; ++i) AtAc[i] = AtA[i];
    llm::colpiv_qr_solve<3, 3>(AtAc, AtB, X);
    if (iterCount == 0) isDegenerate = llm::degeneracy_projector<3>(AtA, 10.f, matP3);
    if (isDegenerate) {
      float X2[3] = {X[0], X[1], X[2]};
      for (int r = 0; r < 3; ++r) X[r] = matP3[r * 3 + 0] * X2[0] + matP3[r * 3 + 1] * X2[1] + matP3[r * 3 + 2] * X2[2];
    }
    transformCur[0] += X[0];
    transformCur[2] += X[1];
    transformCur[4] += X[2];
    for (int i = 0; i < 6; i++)
      if (std::isnan(transformCur[i])) transformCur[i] = 0;
    float deltaR = sqrt(pow(RAD2DEG * (X[0]), 2) + pow(RAD2DEG * (X[1]), 2));
    float deltaT = sqrt(pow(X[2] * 100, 2));
    if (deltaR < 0.1 && deltaT < 0.1) return false;
    return true;
  }

  bool calculateTransformationCorner(int iterCount) { /* featureAssociation.cpp:928-1032 */
    int pointSelNum = laserCloudOri.size();
    std::vector<float> matA(pointSelNum * 3), matB(pointSelNum);
    float srx = sin_(transformCur[0]); float crx = cos_(transformCur[0]);
    float sry = sin_(transformCur[1]); float cry = cos_(transformCur[1]);
    float srz = sin_(transformCur[2]); float crz = cos_(transformCur[2]);
    float tx = transformCur[3]; float ty = transformCur[4]; float tz = transformCur[5];
    float b1 = -crz * sry - cry * srx * srz;
    float b2 = cry * crz * srx - sry * srz;
    float b3 = crx * cry;
    float b4 = tx * -b1 + ty * -b2 + tz * b3;
    float b5 = cry * crz - srx * sry * srz;
    float b6 = cry * srz + crz * srx * sry;
    float b7 = crx * sry;
    float b8 = tz * b7 - ty * b6 - tx * b5;
    float c5 = crx * srz;
    for (int i = 0; i < pointSelNum; i++) {
      P4 pointOri = laserCloudOri[i];
      P4 coeff = coeffSel[i];
      float ary = (b1 * pointOri.x + b2 * pointOri.y - b3 * pointOri.z + b4) * coeff.x +
                  (b5 * pointOri.x + b6 * pointOri.y - b7 * pointOri.z + b8) * coeff.z;
      float atx = -b5 * coeff.x + c5 * coeff.y + b1 * coeff.z;
      float atz = b7 * coeff.x - srx * coeff.y - b3 * coeff.z;
      float d2 = coeff.i;
      matA[i * 3 + 0] = ary; matA[i * 3 + 1] = atx; matA[i * 3 + 2] = atz;
      matB[i] = -0.05 * d2;
    }
    float AtA[9], AtB[3], X[3], AtAc[9];
    normal_equations<3>(matA, matB, pointSelNum, AtA, AtB);
    for (int i = 0; i < 9; ++i) AtAc[i] = AtA[i];
    llm::colpiv_qr_solve<3, 3>(AtAc, AtB, X);
    if (iterCount == 0) isDegenerate = llm::degeneracy_projector<3>(AtA, 10.f, matP3);
    if (isDegenerate) {
      float X2[3] = {X[0], X[1], X[2]};
      for (int r = 0; r < 3; ++r) X[r] = matP3[r * 3 + 0] * X2[0] + matP3[r * 3 + 1] * X2[1] + matP3[r * 3 + 2] * X2[2];
    }
    transformCur[1] += X[0];
    transformCur[3] += X[1];
    transformCur[5] += X[2];
    for (int i = 0; i < 6; i++)
      if (std::isnan(transformCur[i])) transformCur[i] = 0;
    float deltaR = sqrt(pow(RAD2DEG * (X[0]), 2));
    float deltaT = sqrt(pow(X[1] * 100, 2) + pow(X[2] * 100, 2));
    if (deltaR < 0.1 && deltaT < 0.1) return false;
    return true;
  }

  void checkSystemInitialization() { /* featureAssociation.cpp:1181-1209 */
    cornerLessSharp.swap(cornerLast);
    surfLessFlat.swap(surfLast);
    kdCornerLast->setInputCloud(cornerLast);
    kdSurfLast->setInputCloud(surfLast);
    cornerLastNum = cornerLast.size();
    surfLastNum = surfLast.size();
    systemInitedLM = true;
  }

  void updateTransformation() { /* featureAssociation.cpp:1213-1235 */
    odomSearchTrace.assign((size_t)2 * 5 * 24 * V * 3, -1);
    odom_iters[0] = odom_iters[1] = 0;
    if (cornerLastNum < 10 || surfLastNum < 100) return;
    for (int iterCount1 = 0; iterCount1 < 25; iterCount1++) {
      laserCloudOri.clear(); coeffSel.clear();
      findCorrespondingSurfFeatures(iterCount1);
      odom_iters[0] = iterCount1 + 1;
      if (laserCloudOri.size() < 10) continue;
      if (calculateTransformationSurf(iterCount1) == false) break;
    }
    for (int iterCount2 = 0; iterCount2 < 25; iterCount2++) {
      laserCloudOri.clear(); coeffSel.clear();
      findCorrespondingCornerFeatures(iterCount2);
      odom_iters[1] = iterCount2 + 1;
      if (laserCloudOri.size() < 10) continue;
      if (calculateTransformationCorner(iterCount2) == false) break;
    }
  }

  void integrateTransformation() { /* featureAssociation.cpp:1241-1270 */
    float rx, ry, rz, tx, ty, tz;
    AccumulateRotation(transformSum[0], transformSum[1], transformSum[2], -transformCur[0], -transformCur[1],
                       -transformCur[2], rx, ry, rz);
    float x1 = cos_(rz) * (transformCur[3]) - sin_(rz) * (transformCur[4]);
    float y1 = sin_(rz) * (transformCur[3]) + cos_(rz) * (transformCur[4]);
    float z1 = transformCur[5];
    float x2 = x1;
    float y2 = cos_(rx) * y1 - sin_(rx) * z1;
    float z2 = sin_(rx) * y1 + cos_(rx) * z1;
    tx = transformSum[3] - (cos_(ry) * x2 + sin_(ry) * z2);
    ty = transformSum[4] - y2;
    tz = transformSum[5] - (-sin_(ry) * x2 + cos_(ry) * z2);
    transformSum[0] = rx; transformSum[1] = ry; transformSum[2] = rz;
    transformSum[3] = tx; transformSum[4] = ty; transformSum[5] = tz;
  }

  void adjustOutlierCloud() { /* featureAssociation.cpp:1273-1283 */
    outlierLast.resize(outlier_cloud.size());
    for (size_t i = 0; i < outlier_cloud.size(); ++i) {
      P4 point;
      point.x = outlier_cloud[i].y; point.y = outlier_cloud[i].z; point.z = outlier_cloud[i].x;
      point.i = outlier_cloud[i].i;
      outlierLast[i] = point;
    }
  }

  void publishCloudsLast() { /* featureAssociation.cpp:1329-1359 */
    for (size_t i = 0; i < cornerLessSharp.size(); i++) TransformToEnd(&cornerLessSharp[i], &cornerLessSharp[i]);
    for (size_t i = 0; i < surfLessFlat.size(); i++) TransformToEnd(&surfLessFlat[i], &surfLessFlat[i]);
    cornerLessSharp.swap(cornerLast);
    surfLessFlat.swap(surfLast);
    cornerLastNum = cornerLast.size();
    surfLastNum = surfLast.size();
    if (cornerLastNum > 10 && surfLastNum > 100) {
      kdCornerLast->setInputCloud(cornerLast);
      kdSurfLast->setInputCloud(surfLast);
    }
    adjustOutlierCloud();
  }

  int runFeatureAssociationOnce() { /* featureAssociation.cpp:1386-1450, one loop body */
    double t0 = now_s();
    adjustDistortion();
    calculateSmoothness();
    markOccludedPoints();
    extractFeatures();
    double t1 = now_s();
    timers[1] += t1 - t0;
    if (!systemInitedLM) {
      checkSystemInitialization();
      timers[2] += now_s() - t1;
      return 0;
    }
    updateTransformation();
    integrateTransformation();
    publishCloudsLast();
    timers[2] += now_s() - t1;
    cycle_count++;
    if ((int)cycle_count == mapping_frequency_div) {
      cycle_count = 0;
      return 1;
    }
    return 0;
  }

  /* ================= MapOptimization: scan-to-map ================= */

  float cRoll, sRoll, cPitch, sPitch, cYaw, sYaw, tX, tY, tZ;
  void updatePointAssociateToMapSinCos() { /* mapOptmization.cpp:397-410 */
    cRoll = cos_(transformTobeMapped[0]); sRoll = sin_(transformTobeMapped[0]);
    cPitch = cos_(transformTobeMapped[1]); sPitch = sin_(transformTobeMapped[1]);
    cYaw = cos_(transformTobeMapped[2]); sYaw = sin_(transformTobeMapped[2]);
    tX = transformTobeMapped[3]; tY = transformTobeMapped[4]; tZ = transformTobeMapped[5];
  }
  void pointAssociateToMap(const P4* pi, P4* po) { /* mapOptmization.cpp:412-426 */
    float x1 = cYaw * pi->x - sYaw * pi->y;
    float y1 = sYaw * pi->x + cYaw * pi->y;
    float z1 = pi->z;
    float x2 = x1;
    float y2 = cRoll * y1 - sRoll * z1;
    float z2 = sRoll * y1 + cRoll * z1;
    po->x = cPitch * x2 + sPitch * z2 + tX;
    po->y = y2 + tY;
    po->z = -sPitch * x2 + cPitch * z2 + tZ;
    po->i = pi->i;
  }


  /* mapOptmization.cpp:264-387.  MapOptimization::transformSum is the odometry pose that arrives through
   * nav_msgs::Odometry and a tf RPY->quaternion->RPY round trip (featureAssociation.cpp:1287-1297,
   * utility.h:96-110; tf is not vendored): restated as the identity on FeatureAssociation's transformSum. */
  void transformAssociateToMap() {
    const float* tS = transformSum;
    float x1 = cos_(tS[1]) * (transformBefMapped[3] - tS[3]) - sin_(tS[1]) * (transformBefMapped[5] - tS[5]);
    float y1 = transformBefMapped[4] - tS[4];
    float z1 = sin_(tS[1]) * (transformBefMapped[3] - tS[3]) + cos_(tS[1]) * (transformBefMapped[5] - tS[5]);
    float x2 = x1;
    float y2 = cos_(tS[0]) * y1 + sin_(tS[0]) * z1;
    float z2 = -sin_(tS[0]) * y1 + cos_(tS[0]) * z1;
    transformIncre[3] = cos_(tS[2]) * x2 + sin_(tS[2]) * y2;
    transformIncre[4] = -sin_(tS[2]) * x2 + cos_(tS[2]) * y2;
    transformIncre[5] = z2;
    float sbcx = sin_(tS[0]), cbcx = cos_(tS[0]), sbcy = sin_(tS[1]), cbcy = cos_(tS[1]), sbcz = sin_(tS[2]), cbcz = cos_(tS[2]);
    float sblx = sin_(transformBefMapped[0]), cblx = cos_(transformBefMapped[0]);
    float sbly = sin_(transformBefMapped[1]), cbly = cos_(transformBefMapped[1]);
    float sblz = sin_(transformBefMapped[2]), cblz = cos_(transformBefMapped[2]);
    float salx = sin_(transformAftMapped[0]), calx = cos_(transformAftMapped[0]);
    float saly = sin_(transformAftMapped[1]), caly = cos_(transformAftMapped[1]);
    float salz = sin_(transformAftMapped[2]), calz = cos_(transformAftMapped[2]);
    float srx = -sbcx * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz) -
                cbcx * sbcy * (calx * calz * (cbly * sblz - cblz * sblx * sbly) - calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) -
                cbcx * cbcy * (calx * salz * (cblz * sbly - cbly * sblx * sblz) - calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx);
    transformTobeMapped[0] = -asin_(srx);
    float srycrx = sbcx * (cblx * cblz * (caly * salz - calz * salx * saly) - cblx * sblz * (caly * calz + salx * saly * salz) + calx * saly * sblx) -
                   cbcx * cbcy * ((caly * calz + salx * saly * salz) * (cblz * sbly - cbly * sblx * sblz) +
                                  (caly * salz - calz * salx * saly) * (sbly * sblz + cbly * cblz * sblx) - calx * cblx * cbly * saly) +
                   cbcx * sbcy * ((caly * calz + salx * saly * salz) * (cbly * cblz + sblx * sbly * sblz) +
                                  (caly * salz - calz * salx * saly) * (cbly * sblz - cblz * sblx * sbly) + calx * cblx * saly * sbly);
    float crycrx = sbcx * (cblx * sblz * (calz * saly - caly * salx * salz) - cblx * cblz * (saly * salz + caly * calz * salx) + calx * caly * sblx) +
                   cbcx * cbcy * ((saly * salz + caly * calz * salx) * (sbly * sblz + cbly * cblz * sblx) +
                                  (calz * saly - caly * salx * salz) * (cblz * sbly - cbly * sblx * sblz) + calx * caly * cblx * cbly) -
                   cbcx * sbcy * ((saly * salz + caly * calz * salx) * (cbly * sblz - cblz * sblx * sbly) +
                                  (calz * saly - caly * salx * salz) * (cbly * cblz + sblx * sbly * sblz) - calx * caly * cblx * sbly);
    transformTobeMapped[1] = atan2_(srycrx / cos_(transformTobeMapped[0]), crycrx / cos_(transformTobeMapped[0]));
    float srzcrx = (cbcz * sbcy - cbcy * sbcx * sbcz) * (calx * salz * (cblz * sbly - cbly * sblx * sblz) - calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx) -
                   (cbcy * cbcz + sbcx * sbcy * sbcz) * (calx * calz * (cbly * sblz - cblz * sblx * sbly) - calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) +
                   cbcx * sbcz * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz);
    float crzcrx = (cbcy * sbcz - cbcz * sbcx * sbcy) * (calx * calz * (cbly * sblz - cblz * sblx * sbly) - calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) -
                   (sbcy * sbcz + cbcy * cbcz * sbcx) * (calx * salz * (cblz * sbly - cbly * sblx * sblz) - calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx) +
                   cbcx * cbcz * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz);
    transformTobeMapped[2] = atan2_(srzcrx / cos_(transformTobeMapped[0]), crzcrx / cos_(transformTobeMapped[0]));
    x1 = cos_(transformTobeMapped[2]) * transformIncre[3] - sin_(transformTobeMapped[2]) * transformIncre[4];
    y1 = sin_(transformTobeMapped[2]) * transformIncre[3] + cos_(transformTobeMapped[2]) * transformIncre[4];
    z1 = transformIncre[5];
    x2 = x1;
    y2 = cos_(transformTobeMapped[0]) * y1 - sin_(transformTobeMapped[0]) * z1;
    z2 = sin_(transformTobeMapped[0]) * y1 + cos_(transformTobeMapped[0]) * z1;
    transformTobeMapped[3] = transformAftMapped[3] - (cos_(transformTobeMapped[1]) * x2 + sin_(transformTobeMapped[1]) * z2);
    transformTobeMapped[4] = transformAftMapped[4] - y2;
    transformTobeMapped[5] = transformAftMapped[5] - (-sin_(transformTobeMapped[1]) * x2 + cos_(transformTobeMapped[1]) * z2);
  }

  void transformUpdate() { /* mapOptmization.cpp:389-395 */
    for (int i = 0; i < 6; i++) {
      transformBefMapped[i] = transformSum[i];
      transformAftMapped[i] = transformTobeMapped[i];
    }
  }

  void cornerOptimization(int iterCount) { /* mapOptmization.cpp:1028-1134 */
    updatePointAssociateToMapSinCos();
    const std::vector<P4>& mp = kdCornerMap->cloud();
    for (size_t i = 0; i < scanCornerDS.size(); i++) {
      P4 pointOri = scanCornerDS[i], pointSel;
      pointAssociateToMap(&pointOri, &pointSel);
      int ind[5]; float dis[5];
      kdCornerMap->nearestKSearch(pointSel, 5, ind, dis);
      traceKnn(iterCount, (int)i, ind, dis);
      if (dis[4] < 1.0) {
        float cx = 0, cy = 0, cz = 0;
        for (int j = 0; j < 5; j++) { cx += mp[ind[j]].x; cy += mp[ind[j]].y; cz += mp[ind[j]].z; }
        cx /= 5; cy /= 5; cz /= 5;
        float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
        for (int j = 0; j < 5; j++) {
          float ax = mp[ind[j]].x - cx;
          float ay = mp[ind[j]].y - cy;
          float az = mp[ind[j]].z - cz;
          a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
          a22 += ay * ay; a23 += ay * az; a33 += az * az;
        }
        a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
        float matA1[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33};
        float matD1[3], matV1[9];
        llm::self_adjoint_eigen<3>(matA1, matD1, matV1);
        if (matD1[2] > 3 * matD1[1]) {
          float x0 = pointSel.x, y0 = pointSel.y, z0 = pointSel.z;
          /* row 0 of the eigenvector matrix (sic), mapOptmization.cpp:1086-1091 */
          float x1 = cx + 0.1 * matV1[0 * 3 + 0];
          float y1 = cy + 0.1 * matV1[0 * 3 + 1];
          float z1 = cz + 0.1 * matV1[0 * 3 + 2];
          float x2 = cx - 0.1 * matV1[0 * 3 + 0];
          float y2 = cy - 0.1 * matV1[0 * 3 + 1];
          float z2 = cz - 0.1 * matV1[0 * 3 + 2];
          float a012 = sqrtf(((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) +
                             ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) +
                             ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)));
          float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
          float la = ((y1 - y2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) +
                      (z1 - z2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))) / a012 / l12;
          float lb = -((x1 - x2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) -
                       (z1 - z2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
          float lc = -((x1 - x2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) +
                       (y1 - y2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
          float ld2 = a012 / l12;
          float s = 1 - 0.9 * fabsf(ld2);
          P4 coeff;
          coeff.x = s * la; coeff.y = s * lb; coeff.z = s * lc; coeff.i = s * ld2;
          if (s > 0.1) { moOri.push_back(pointOri); moCoeff.push_back(coeff); }
        }
      }
    }
  }

  void surfOptimization(int iterCount) { /* mapOptmization.cpp:1136-1197 */
    updatePointAssociateToMapSinCos();
    const std::vector<P4>& mp = kdSurfMap->cloud();
    for (size_t i = 0; i < scanSurfTotalDS.size(); i++) {
      P4 pointOri = scanSurfTotalDS[i], pointSel;
      pointAssociateToMap(&pointOri, &pointSel);
      int ind[5]; float dis[5];
      kdSurfMap->nearestKSearch(pointSel, 5, ind, dis);
      traceKnn(iterCount, (int)(scanCornerDS.size() + i), ind, dis);
      if (dis[4] < 1.0) {
        float matA0[15], matB0[5] = {-1, -1, -1, -1, -1}, matX0[3];
        for (int j = 0; j < 5; j++) { matA0[j * 3 + 0] = mp[ind[j]].x; matA0[j * 3 + 1] = mp[ind[j]].y; matA0[j * 3 + 2] = mp[ind[j]].z; }
        llm::colpiv_qr_solve<5, 3>(matA0, matB0, matX0);
        float pa = matX0[0], pb = matX0[1], pc = matX0[2], pd = 1;
        float ps = sqrtf(pa * pa + pb * pb + pc * pc);
        pa /= ps; pb /= ps; pc /= ps; pd /= ps;
        bool planeValid = true;
        for (int j = 0; j < 5; j++) {
          if (fabsf(pa * mp[ind[j]].x + pb * mp[ind[j]].y + pc * mp[ind[j]].z + pd) > 0.2) { planeValid = false; break; }
        }
        if (planeValid) {
          float pd2 = pa * pointSel.x + pb * pointSel.y + pc * pointSel.z + pd;
          float s = 1 - 0.9 * fabsf(pd2) /
                            sqrtf(sqrtf(pointSel.x * pointSel.x + pointSel.y * pointSel.y + pointSel.z * pointSel.z));
          P4 coeff;
          coeff.x = s * pa; coeff.y = s * pb; coeff.z = s * pc; coeff.i = s * pd2;
          if (s > 0.1) { moOri.push_back(pointOri); moCoeff.push_back(coeff); }
        }
      }
    }
  }

  bool LMOptimization(int iterCount) { /* mapOptmization.cpp:1199-1312 */
    float srx = sin_(transformTobeMapped[0]); float crx = cos_(transformTobeMapped[0]);
    float sry = sin_(transformTobeMapped[1]); float cry = cos_(transformTobeMapped[1]);
    float srz = sin_(transformTobeMapped[2]); float crz = cos_(transformTobeMapped[2]);
    int laserCloudSelNum = moOri.size();
    map_iters[1] = laserCloudSelNum;
    double* trace = map_trace + 34 * iterCount;
    for (int i = 0; i < 34; ++i) trace[i] = 0.0;
    trace[27] = laserCloudSelNum;
    if (laserCloudSelNum < 50) return false;
    std::vector<float> matA(laserCloudSelNum * 6), matB(laserCloudSelNum);
    for (int i = 0; i < laserCloudSelNum; i++) {
      P4 pointOri = moOri[i];
      P4 coeff = moCoeff[i];
      float arx = (crx * sry * srz * pointOri.x + crx * crz * sry * pointOri.y - srx * sry * pointOri.z) * coeff.x +
                  (-srx * srz * pointOri.x - crz * srx * pointOri.y - crx * pointOri.z) * coeff.y +
                  (crx * cry * srz * pointOri.x + crx * cry * crz * pointOri.y - cry * srx * pointOri.z) * coeff.z;
      float ary = ((cry * srx * srz - crz * sry) * pointOri.x + (sry * srz + cry * crz * srx) * pointOri.y +
                   crx * cry * pointOri.z) * coeff.x +
                  ((-cry * crz - srx * sry * srz) * pointOri.x + (cry * srz - crz * srx * sry) * pointOri.y -
                   crx * sry * pointOri.z) * coeff.z;
      float arz = ((crz * srx * sry - cry * srz) * pointOri.x + (-cry * crz - srx * sry * srz) * pointOri.y) * coeff.x +
                  (crx * crz * pointOri.x - crx * srz * pointOri.y) * coeff.y +
                  ((sry * srz + cry * crz * srx) * pointOri.x + (crz * sry - cry * srx * srz) * pointOri.y) * coeff.z;
      matA[i * 6 + 0] = arx; matA[i * 6 + 1] = ary; matA[i * 6 + 2] = arz;
      matA[i * 6 + 3] = coeff.x; matA[i * 6 + 4] = coeff.y; matA[i * 6 + 5] = coeff.z;
      matB[i] = -coeff.i;
    }
    float AtA[36], AtB[6], X[6], AtAc[36];
    normal_equations<6>(matA, matB, laserCloudSelNum, AtA, AtB);
    { int k = 0; for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) trace[k++] = AtA[r * 6 + c];
      for (int r = 0; r < 6; ++r) trace[21 + r] = AtB[r]; }
    for (int i = 0; i < 36; ++i) AtAc[i] = AtA[i];
    llm::colpiv_qr_solve<6, 6>(AtAc, AtB, X);
    if (iterCount == 0) mapDegenerate = llm::degeneracy_projector<6>(AtA, 100.f, matP6);
    if (mapDegenerate) {
      float X2[6];
      for (int i = 0; i < 6; ++i) X2[i] = X[i];
      for (int r = 0; r < 6; ++r) {
        float s = 0.f;
        for (int c = 0; c < 6; ++c) s += matP6[r * 6 + c] * X2[c];
        X[r] = s;
      }
    }
    for (int i = 0; i < 6; ++i) transformTobeMapped[i] += X[i];
    for (int i = 0; i < 6; ++i) trace[28 + i] = X[i];
    const float r2d = 57.29578f; /* pcl::rad2deg(float) */
    float deltaR = sqrt(pow(X[0] * r2d, 2) + pow(X[1] * r2d, 2) + pow(X[2] * r2d, 2));
    float deltaT = sqrt(pow(X[3] * 100, 2) + pow(X[4] * 100, 2) + pow(X[5] * 100, 2));
    if (deltaR < 0.05 && deltaT < 0.05) return true;
    return false;
  }

  void scan2MapOptimization() { /* mapOptmization.cpp:1315-1332 (transformUpdate is host glue) */
    map_iters[0] = 0; map_iters[1] = 0;
    mapKnnTrace.assign((size_t)10 * (scanCornerDS.size() + scanSurfTotalDS.size()) * 5, -1);
    if (mapCorner.size() > 10 && mapSurf.size() > 100) {
      kdCornerMap->setInputCloud(mapCorner);
      kdSurfMap->setInputCloud(mapSurf);
      for (int iterCount = 0; iterCount < 10; iterCount++) {
        moOri.clear(); moCoeff.clear();
        cornerOptimization(iterCount);
        surfOptimization(iterCount);
        map_iters[0] = iterCount + 1;
        if (LMOptimization(iterCount) == true) break;
      }
      transformUpdate();
    }
  }

  void downsampleCurrentScan() { /* mapOptmization.cpp:999-1026 */
    std::vector<P4>& surfDS = surfLastDS;
    std::vector<P4>& outDS = outlierLastDS;
    std::vector<P4> total;
    voxel_grid(cornerLast, 0.2f, scanCornerDS);
    voxel_grid(surfLast, 0.4f, surfDS);
    voxel_grid(outlierLast, 0.4f, outDS);
    total = surfDS;
    total.insert(total.end(), outDS.begin(), outDS.end());
    voxel_grid(total, 0.4f, scanSurfTotalDS);
  }

  /* ================= MapOptimization: key frames and the local map, loop closure off =================
   *   cloudKeyPoses3D/6D, the three key-frame cloud stores, the surrounding-key-frame
   * cache and the two robot positions are the members of mapOptimization.h; iSAM2 is replaced by the identity
   * (SURVEY.md sections 8c, 11.5: a pure odometry chain's optimum is its initial values, and the double
   * RzRyRx -> rpy round trip returns the float it was given away from gimbal lock). */
  struct Pose6 { float roll, pitch, yaw, x, y, z; };
  std::vector<P4> surfLastDS, outlierLastDS;    /* laserCloudSurfLastDS, laserCloudOutlierLastDS */
  std::vector<P4> cloudKeyPoses3D;               /* intensity = key-frame index */
  std::vector<Pose6> cloudKeyPoses6D;
  std::vector<std::vector<P4> > cornerCloudKeyFrames, surfCloudKeyFrames, outlierCloudKeyFrames;
  std::vector<int> surroundingExistingKeyPosesID;
  std::vector<std::vector<P4> > surroundingCornerCloudKeyFrames, surroundingSurfCloudKeyFrames, surroundingOutlierCloudKeyFrames;
  std::unique_ptr<oknn::KdTree> kdSurroundingKeyPoses;
  P4 currentRobotPosPoint, previousRobotPosPoint; /* PointType(): zeros */
  float transformLast[6];
  int last_rebuild; /* test aid: 1 when the last extractSurroundingKeyFrames erased a key frame */

  void resetKeyFrames() {
    surfLastDS.clear(); outlierLastDS.clear();
    cloudKeyPoses3D.clear(); cloudKeyPoses6D.clear();
    cornerCloudKeyFrames.clear(); surfCloudKeyFrames.clear(); outlierCloudKeyFrames.clear();
    surroundingExistingKeyPosesID.clear();
    surroundingCornerCloudKeyFrames.clear(); surroundingSurfCloudKeyFrames.clear(); surroundingOutlierCloudKeyFrames.clear();
    kdSurroundingKeyPoses.reset(new oknn::KdTree());
    currentRobotPosPoint.x = currentRobotPosPoint.y = currentRobotPosPoint.z = currentRobotPosPoint.i = 0.f;
    previousRobotPosPoint = currentRobotPosPoint;
    for (int i = 0; i < 6; ++i) transformLast[i] = 0;
    last_rebuild = 0;
    timer_map_assembly = 0;
  }

  /* updateTransformPointCloudSinCos + transformPointCloud (mapOptmization.cpp:428-473) */
  std::vector<P4> transformPointCloud(const std::vector<P4>& cloudIn, const Pose6& tIn) {
    const float ctRoll = cos_(tIn.roll), stRoll = sin_(tIn.roll);
    const float ctPitch = cos_(tIn.pitch), stPitch = sin_(tIn.pitch);
    const float ctYaw = cos_(tIn.yaw), stYaw = sin_(tIn.yaw);
    const float tInX = tIn.x, tInY = tIn.y, tInZ = tIn.z;
    std::vector<P4> cloudOut(cloudIn.size());
    for (size_t i = 0; i < cloudIn.size(); ++i) {
      const P4* pointFrom = &cloudIn[i];
      float x1 = ctYaw * pointFrom->x - stYaw * pointFrom->y;
      float y1 = stYaw * pointFrom->x + ctYaw * pointFrom->y;
      float z1 = pointFrom->z;
      float x2 = x1;
      float y2 = ctRoll * y1 - stRoll * z1;
      float z2 = stRoll * y1 + ctRoll * z1;
      P4 pointTo;
      pointTo.x = ctPitch * x2 + stPitch * z2 + tInX;
      pointTo.y = y2 + tInY;
      pointTo.z = -stPitch * x2 + ctPitch * z2 + tInZ;
      pointTo.i = pointFrom->i;
      cloudOut[i] = pointTo;
    }
    return cloudOut;
  }

  void clearCloud() { mapCorner.clear(); mapSurf.clear(); } /* mapOptmization.cpp:1518-1523 */

  void extractSurroundingKeyFrames() { /* mapOptmization.cpp:856-996, the enable_loop_closure == false branch :915-987 */
    last_rebuild = 0;
    clearCloud(); /* the reference clears at the end of the previous cycle (:1560) */
    if (cloudKeyPoses3D.empty()) return;
    std::vector<P4> surroundingKeyPoses, surroundingKeyPosesDS;
    std::vector<int> pointSearchInd;
    std::vector<float> pointSearchSqDis;
    kdSurroundingKeyPoses->setInputCloud(cloudKeyPoses3D);
    kdSurroundingKeyPoses->radiusSearch(currentRobotPosPoint, (double)prm.surrounding_keyframe_search_radius, pointSearchInd,
                                        pointSearchSqDis);
    for (size_t i = 0; i < pointSearchInd.size(); ++i) surroundingKeyPoses.push_back(cloudKeyPoses3D[pointSearchInd[i]]);
    voxel_grid(surroundingKeyPoses, 1.0f, surroundingKeyPosesDS); /* downSizeFilterSurroundingKeyPoses, :78 */
    /* delete key frames that are not in the surrounding region (:935-955) */
    const int numSurroundingPosesDS = (int)surroundingKeyPosesDS.size();
    for (int i = 0; i < (int)surroundingExistingKeyPosesID.size(); ++i) {
      bool existingFlag = false;
      for (int j = 0; j < numSurroundingPosesDS; ++j) {
        if (surroundingExistingKeyPosesID[i] == (int)surroundingKeyPosesDS[j].i) { existingFlag = true; break; }
      }
      if (existingFlag == false) {
        surroundingExistingKeyPosesID.erase(surroundingExistingKeyPosesID.begin() + i);
        surroundingCornerCloudKeyFrames.erase(surroundingCornerCloudKeyFrames.begin() + i);
        surroundingSurfCloudKeyFrames.erase(surroundingSurfCloudKeyFrames.begin() + i);
        surroundingOutlierCloudKeyFrames.erase(surroundingOutlierCloudKeyFrames.begin() + i);
        --i;
        last_rebuild = 1;
      }
    }
    /* add new key frames (:957-980) */
    for (int i = 0; i < numSurroundingPosesDS; ++i) {
      bool existingFlag = false;
      for (size_t k = 0; k < surroundingExistingKeyPosesID.size(); ++k) {
        if (surroundingExistingKeyPosesID[k] == (int)surroundingKeyPosesDS[i].i) { existingFlag = true; break; }
      }
      if (existingFlag == true) continue;
      const int thisKeyInd = (int)surroundingKeyPosesDS[i].i;
      const Pose6 thisTransformation = cloudKeyPoses6D[thisKeyInd];
      surroundingExistingKeyPosesID.push_back(thisKeyInd);
      surroundingCornerCloudKeyFrames.push_back(transformPointCloud(cornerCloudKeyFrames[thisKeyInd], thisTransformation));
      surroundingSurfCloudKeyFrames.push_back(transformPointCloud(surfCloudKeyFrames[thisKeyInd], thisTransformation));
      surroundingOutlierCloudKeyFrames.push_back(transformPointCloud(outlierCloudKeyFrames[thisKeyInd], thisTransformation));
    }
    std::vector<P4> laserCloudCornerFromMap, laserCloudSurfFromMap;
    for (size_t i = 0; i < surroundingExistingKeyPosesID.size(); ++i) { /* :982-986 */
      laserCloudCornerFromMap.insert(laserCloudCornerFromMap.end(), surroundingCornerCloudKeyFrames[i].begin(), surroundingCornerCloudKeyFrames[i].end());
      laserCloudSurfFromMap.insert(laserCloudSurfFromMap.end(), surroundingSurfCloudKeyFrames[i].begin(), surroundingSurfCloudKeyFrames[i].end());
      laserCloudSurfFromMap.insert(laserCloudSurfFromMap.end(), surroundingOutlierCloudKeyFrames[i].begin(), surroundingOutlierCloudKeyFrames[i].end());
    }
    voxel_grid(laserCloudCornerFromMap, 0.2f, mapCorner); /* downSizeFilterCorner :989-991 */
    voxel_grid(laserCloudSurfFromMap, 0.4f, mapSurf);     /* downSizeFilterSurf :993-995 */
  }

  void saveKeyFramesAndFactor() { /* mapOptmization.cpp:1335-1474 with iSAM2 == identity */
    currentRobotPosPoint.x = transformAftMapped[3];
    currentRobotPosPoint.y = transformAftMapped[4];
    currentRobotPosPoint.z = transformAftMapped[5];
    bool saveThisKeyFrame = true;
    if (om::sqrt_((previousRobotPosPoint.x - currentRobotPosPoint.x) * (previousRobotPosPoint.x - currentRobotPosPoint.x) +
                  (previousRobotPosPoint.y - currentRobotPosPoint.y) * (previousRobotPosPoint.y - currentRobotPosPoint.y) +
                  (previousRobotPosPoint.z - currentRobotPosPoint.z) * (previousRobotPosPoint.z - currentRobotPosPoint.z)) < 0.3) {
      saveThisKeyFrame = false;
    }
    if (saveThisKeyFrame == false && !cloudKeyPoses3D.empty()) return;
    previousRobotPosPoint = currentRobotPosPoint;
    /* first key frame: PriorFactor on transformTobeMapped (:1362-1376); later: transformAftMapped (:1391-1400).
     * latestEstimate == the inserted initial value. */
    const float* est = cloudKeyPoses3D.empty() ? transformTobeMapped : transformAftMapped;
    if (cloudKeyPoses3D.empty()) for (int i = 0; i < 6; ++i) transformLast[i] = transformTobeMapped[i];
    P4 thisPose3D;
    thisPose3D.x = est[3]; thisPose3D.y = est[4]; thisPose3D.z = est[5]; /* translation().y(), .z(), .x() of Point3(T[5],T[3],T[4]) */
    thisPose3D.i = (float)cloudKeyPoses3D.size();
    cloudKeyPoses3D.push_back(thisPose3D);
    Pose6 thisPose6D;
    thisPose6D.x = thisPose3D.x; thisPose6D.y = thisPose3D.y; thisPose6D.z = thisPose3D.z;
    thisPose6D.roll = est[0]; thisPose6D.pitch = est[1]; thisPose6D.yaw = est[2]; /* rotation().pitch(), .yaw(), .roll() of RzRyRx(T[2],T[0],T[1]) */
    cloudKeyPoses6D.push_back(thisPose6D);
    if (cloudKeyPoses3D.size() > 1) { /* :1440-1452 */
      for (int i = 0; i < 6; ++i) { transformLast[i] = transformAftMapped[i]; transformTobeMapped[i] = transformAftMapped[i]; }
    }
    cornerCloudKeyFrames.push_back(scanCornerDS); /* copies of laserCloudCornerLastDS / SurfLastDS / OutlierLastDS :1461-1470 */
    surfCloudKeyFrames.push_back(surfLastDS);
    outlierCloudKeyFrames.push_back(outlierLastDS);
  }

  /* one body of MapOptimization::run (mapOptmization.cpp:1526-1562); transformSum arrives as FeatureAssociation's */
  double timer_map_assembly;
  void mappingCycle() {
    transformAssociateToMap();
    double t0 = now_s();
    extractSurroundingKeyFrames();
    double t1 = now_s();
    downsampleCurrentScan();
    double t2 = now_s();
    scan2MapOptimization();
    double t3 = now_s();
    saveKeyFramesAndFactor();
    timer_map_assembly += (t1 - t0) + (now_s() - t3);
    timers[4] += t2 - t1;
    timers[3] += t3 - t2;
    /* correctPoses: nothing without a closed loop; clearCloud() is left to the next cycle's extract so that the
     * local map of this cycle stays downloadable */
  }
};

/* ============================ C interface ============================ */

template <typename T>
static int copy_out(const T* src, size_t n, void* dst, size_t dst_bytes, size_t* n_elems) {
  if (n_elems) *n_elems = n;
  if (!dst) return 0;
  if (dst_bytes < n * sizeof(T)) return LL_ERR_CAPACITY;
  if (n) std::memcpy(dst, src, n * sizeof(T));
  return 0;
}

extern "C" {

void lo_set_math_backend(int use_libm) { om::g_use_libm = use_libm ? 1 : 0; }
void lo_set_accum_backend(int float_accum) { g_float_accum = float_accum; }
void lo_set_sort_backend(int stable) { g_stable_sort = stable ? 1 : 0; }
void lo_set_knn_backend(int use_nanoflann) {
#ifdef ORACLE_WITH_NANOFLANN
  oknn::g_use_nanoflann = use_nanoflann ? 1 : 0;
#else
  (void)use_nanoflann;
  oknn::g_use_nanoflann = 0;
#endif
}
int lo_has_nanoflann(void) {
#ifdef ORACLE_WITH_NANOFLANN
  return 1;
#else
  return 0;
#endif
}

lo_handle* lo_create(const LegoLoamParams* p) {
  if (!p || p->num_vertical_scans < 2 || p->num_horizontal_scans < 1) return nullptr;
  return new lo_handle(*p);
}
void lo_destroy(lo_handle* h) { delete h; }
void lo_reset(lo_handle* h) { h->reset(); }
void lo_reset_feature_association(lo_handle* h) { h->resetFeatureAssociation(); }

int lo_image_projection(lo_handle* h, const float* xyzi, int n) {
  double t0 = now_s();
  h->cloudHandler(xyzi, n);
  h->timers[0] += now_s() - t0;
  return 0;
}
int lo_feature_association(lo_handle* h) { return h->runFeatureAssociationOnce(); }

int lo_map_set_local(lo_handle* h, const float* corner, int nc, const float* surf, int ns) {
  h->mapCorner.resize(nc); h->mapSurf.resize(ns);
  if (nc) std::memcpy(h->mapCorner.data(), corner, sizeof(P4) * (size_t)nc);
  if (ns) std::memcpy(h->mapSurf.data(), surf, sizeof(P4) * (size_t)ns);
  return 0;
}
int lo_map_set_scan(lo_handle* h, const float* corner, int nc, const float* surf, int ns) {
  h->scanCornerDS.resize(nc); h->scanSurfTotalDS.resize(ns);
  if (nc) std::memcpy(h->scanCornerDS.data(), corner, sizeof(P4) * (size_t)nc);
  if (ns) std::memcpy(h->scanSurfTotalDS.data(), surf, sizeof(P4) * (size_t)ns);
  return 0;
}
int lo_map_downsample_current_scan(lo_handle* h) {
  double t0 = now_s();
  h->downsampleCurrentScan();
  h->timers[4] += now_s() - t0;
  return 0;
}
int lo_map_set_initial_guess(lo_handle* h, const float* t6) {
  for (int i = 0; i < 6; ++i) h->transformTobeMapped[i] = t6[i];
  return 0;
}
int lo_map_set_poses(lo_handle* h, const float* aft6, const float* bef6) {
  for (int i = 0; i < 6; ++i) { h->transformAftMapped[i] = aft6[i]; h->transformBefMapped[i] = bef6[i]; }
  return 0;
}
int lo_map_predict_pose(lo_handle* h) { h->transformAssociateToMap(); return 0; }
int lo_scan_to_map(lo_handle* h) {
  double t0 = now_s();
  h->scan2MapOptimization();
  h->timers[3] += now_s() - t0;
  return 0;
}

/* ---- key frames and the local map ---- */
int lo_map_extract_surrounding_keyframes(lo_handle* h) { h->extractSurroundingKeyFrames(); return 0; }
int lo_map_save_keyframe(lo_handle* h) { h->saveKeyFramesAndFactor(); return 0; }
int lo_mapping_cycle(lo_handle* h) { h->mappingCycle(); return 0; }
double lo_get_timer_map_assembly(lo_handle* h) { return h->timer_map_assembly; }
/* which: 0 corner, 1 surf, 2 outlier; the key frame's cloud transformed by its own pose (transformPointCloud) */
int lo_map_download_keyframe(lo_handle* h, int kf, int which, void* dst, size_t dst_bytes, size_t* n) {
  if (kf < 0 || kf >= (int)h->cloudKeyPoses6D.size() || which < 0 || which > 2) return LL_ERR_INVALID_ARG;
  const std::vector<P4>& src = which == 0 ? h->cornerCloudKeyFrames[kf] : (which == 1 ? h->surfCloudKeyFrames[kf] : h->outlierCloudKeyFrames[kf]);
  const std::vector<P4> t = h->transformPointCloud(src, h->cloudKeyPoses6D[kf]);
  return copy_out(t.data(), t.size(), dst, dst_bytes, n);
}

int lo_download(lo_handle* h, int buffer, void* dst, size_t dst_bytes, size_t* n) {
  const size_t S = h->segmented_cloud.size();
  switch (buffer) {
    case LL_BUF_RANGE_MAT: return copy_out(h->range_mat.data(), h->range_mat.size(), dst, dst_bytes, n);
    case LL_BUF_FULL_CLOUD: return copy_out(h->full_cloud.data(), h->full_cloud.size(), dst, dst_bytes, n);
    case LL_BUF_GROUND_MAT: return copy_out(h->ground_mat.data(), h->ground_mat.size(), dst, dst_bytes, n);
    case LL_BUF_LABEL_MAT: return copy_out(h->label_mat.data(), h->label_mat.size(), dst, dst_bytes, n);
    case LL_BUF_SEG_CLOUD: return copy_out(h->segmented_cloud.data(), S, dst, dst_bytes, n);
    case LL_BUF_SEG_GROUND_FLAG: return copy_out(h->seg_ground_flag.data(), S, dst, dst_bytes, n);
    case LL_BUF_SEG_COL_IND: return copy_out(h->seg_col_ind.data(), S, dst, dst_bytes, n);
    case LL_BUF_SEG_RANGE: return copy_out(h->seg_range.data(), S, dst, dst_bytes, n);
    case LL_BUF_START_RING_INDEX: return copy_out(h->start_ring.data(), h->start_ring.size(), dst, dst_bytes, n);
    case LL_BUF_END_RING_INDEX: return copy_out(h->end_ring.data(), h->end_ring.size(), dst, dst_bytes, n);
    case LL_BUF_ORIENTATION: { float o[3] = {h->start_ori, h->end_ori, h->ori_diff}; return copy_out(o, 3, dst, dst_bytes, n); }
    case LL_BUF_OUTLIER_CLOUD: return copy_out(h->outlier_cloud.data(), h->outlier_cloud.size(), dst, dst_bytes, n);
    case LL_BUF_CLOUD_CURVATURE: return copy_out(h->cloudCurvature.data(), h->cloudCurvature.size(), dst, dst_bytes, n);
    case LL_BUF_NEIGHBOR_PICKED: return copy_out(h->cloudNeighborPicked.data(), h->cloudNeighborPicked.size(), dst, dst_bytes, n);
    case LL_BUF_CLOUD_LABEL: return copy_out(h->cloudLabel.data(), h->cloudLabel.size(), dst, dst_bytes, n);
    case LL_BUF_CORNER_SHARP: return copy_out(h->cornerSharp.data(), h->cornerSharp.size(), dst, dst_bytes, n);
    case LL_BUF_CORNER_LESS_SHARP: return copy_out(h->cornerLessSharp.data(), h->cornerLessSharp.size(), dst, dst_bytes, n);
    case LL_BUF_SURF_FLAT: return copy_out(h->surfFlat.data(), h->surfFlat.size(), dst, dst_bytes, n);
    case LL_BUF_SURF_LESS_FLAT: return copy_out(h->surfLessFlat.data(), h->surfLessFlat.size(), dst, dst_bytes, n);
    case LL_BUF_CORNER_SHARP_IND: return copy_out(h->cornerSharpInd.data(), h->cornerSharpInd.size(), dst, dst_bytes, n);
    case LL_BUF_CORNER_LESS_SHARP_IND: return copy_out(h->cornerLessSharpInd.data(), h->cornerLessSharpInd.size(), dst, dst_bytes, n);
    case LL_BUF_SURF_FLAT_IND: return copy_out(h->surfFlatInd.data(), h->surfFlatInd.size(), dst, dst_bytes, n);
    case LL_BUF_CORNER_LAST: return copy_out(h->cornerLast.data(), h->cornerLast.size(), dst, dst_bytes, n);
    case LL_BUF_SURF_LAST: return copy_out(h->surfLast.data(), h->surfLast.size(), dst, dst_bytes, n);
    case LL_BUF_TRANSFORM_CUR: return copy_out(h->transformCur, 6, dst, dst_bytes, n);
    case LL_BUF_TRANSFORM_SUM: return copy_out(h->transformSum, 6, dst, dst_bytes, n);
    case LL_BUF_ODOM_ITERS: return copy_out(h->odom_iters, 2, dst, dst_bytes, n);
    case LL_BUF_MAP_CORNER: return copy_out(h->mapCorner.data(), h->mapCorner.size(), dst, dst_bytes, n);
    case LL_BUF_MAP_SURF: return copy_out(h->mapSurf.data(), h->mapSurf.size(), dst, dst_bytes, n);
    case LL_BUF_SCAN_CORNER_DS: return copy_out(h->scanCornerDS.data(), h->scanCornerDS.size(), dst, dst_bytes, n);
    case LL_BUF_SCAN_SURF_TOTAL_DS: return copy_out(h->scanSurfTotalDS.data(), h->scanSurfTotalDS.size(), dst, dst_bytes, n);
    case LL_BUF_TRANSFORM_TOBE_MAPPED: return copy_out(h->transformTobeMapped, 6, dst, dst_bytes, n);
    case LL_BUF_MAP_ITERS: return copy_out(h->map_iters, 2, dst, dst_bytes, n);
    case LL_BUF_MAP_TRACE: return copy_out(h->map_trace, 340, dst, dst_bytes, n);
    case LL_BUF_TRANSFORM_BEF_MAPPED: return copy_out(h->transformBefMapped, 6, dst, dst_bytes, n);
    case LL_BUF_TRANSFORM_AFT_MAPPED: return copy_out(h->transformAftMapped, 6, dst, dst_bytes, n);
    case LL_BUF_OUTLIER_LAST: return copy_out(h->outlierLast.data(), h->outlierLast.size(), dst, dst_bytes, n);
    case LL_BUF_SURF_LESS_FLAT_RAW_COUNT: return copy_out(h->lessFlatRawCount.data(), h->lessFlatRawCount.size(), dst, dst_bytes, n);
    case LL_BUF_SCAN_SURF_DS: return copy_out(h->surfLastDS.data(), h->surfLastDS.size(), dst, dst_bytes, n);
    case LL_BUF_SCAN_OUTLIER_DS: return copy_out(h->outlierLastDS.data(), h->outlierLastDS.size(), dst, dst_bytes, n);
    case LL_BUF_KEYFRAME_STATE: {
      int v[4] = {(int)h->cloudKeyPoses6D.size(), (int)h->surroundingExistingKeyPosesID.size(), h->last_rebuild, 0};
      return copy_out(v, 4, dst, dst_bytes, n);
    }
    case LL_BUF_KEY_POSES_6D: return copy_out(h->cloudKeyPoses6D.data(), h->cloudKeyPoses6D.size(), dst, dst_bytes, n);
    case LL_BUF_SURROUNDING_KEY_IDS: return copy_out(h->surroundingExistingKeyPosesID.data(), h->surroundingExistingKeyPosesID.size(), dst, dst_bytes, n);
    case LL_BUF_MAP_KNN_IDX: { /* elements are rows of 5 indices */
      struct R5 { int v[5]; };
      return copy_out(reinterpret_cast<const R5*>(h->mapKnnTrace.data()), h->mapKnnTrace.size() / 5, dst, dst_bytes, n);
    }
    case LL_BUF_ODOM_SEARCH_IDX: { /* rows of 3 indices */
      struct R3 { int v[3]; };
      return copy_out(reinterpret_cast<const R3*>(h->odomSearchTrace.data()), h->odomSearchTrace.size() / 3, dst, dst_bytes, n);
    }
    default: return LL_ERR_INVALID_ARG;
  }
}

int lo_upload(lo_handle* h, int buffer, const void* src, size_t n_elems) {
  const float* f = (const float*)src;
  switch (buffer) {
    case LL_BUF_TRANSFORM_CUR: if (n_elems != 6) return LL_ERR_INVALID_ARG; for (int i = 0; i < 6; ++i) h->transformCur[i] = f[i]; return 0;
    case LL_BUF_TRANSFORM_SUM: if (n_elems != 6) return LL_ERR_INVALID_ARG; for (int i = 0; i < 6; ++i) h->transformSum[i] = f[i]; return 0;
    case LL_BUF_TRANSFORM_TOBE_MAPPED: if (n_elems != 6) return LL_ERR_INVALID_ARG; for (int i = 0; i < 6; ++i) h->transformTobeMapped[i] = f[i]; return 0;
    case LL_BUF_TRANSFORM_BEF_MAPPED: if (n_elems != 6) return LL_ERR_INVALID_ARG; for (int i = 0; i < 6; ++i) h->transformBefMapped[i] = f[i]; return 0;
    case LL_BUF_TRANSFORM_AFT_MAPPED: if (n_elems != 6) return LL_ERR_INVALID_ARG; for (int i = 0; i < 6; ++i) h->transformAftMapped[i] = f[i]; return 0;
    default: return LL_ERR_INVALID_ARG;
  }
}

/* pcl::fromROSMsg + pcl::removeNaNFromPointCloud (imageProjection.cpp:159-161) restated: FLOAT32 fields x, y, z,
 * intensity of every point_step-sized record (absent intensity field: the PointXYZI default 0); a dense cloud is
 * copied unchecked, otherwise points with a non-finite x, y or z are dropped and the order is kept.  PCL is not
 * vendored: parity unpinned.  Returns the number of points written to out_xyzi. */
int lo_decode_pointcloud2(const unsigned char* data, int n_points, int point_step, int off_x, int off_y, int off_z,
                          int off_intensity, int is_dense, float* out_xyzi) {
  int m = 0;
  for (int i = 0; i < n_points; ++i) {
    const unsigned char* rec = data + (size_t)i * point_step;
    float x, y, z, in = 0.f;
    std::memcpy(&x, rec + off_x, 4);
    std::memcpy(&y, rec + off_y, 4);
    std::memcpy(&z, rec + off_z, 4);
    if (off_intensity >= 0) std::memcpy(&in, rec + off_intensity, 4);
    if (!is_dense && (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z))) continue;
    out_xyzi[4 * m + 0] = x; out_xyzi[4 * m + 1] = y; out_xyzi[4 * m + 2] = z; out_xyzi[4 * m + 3] = in;
    ++m;
  }
  return m;
}

int lo_voxel_grid(const float* xyzi, int n, float leaf, float* out_xyzi) {
  std::vector<P4> in(n), out;
  if (n) std::memcpy(in.data(), xyzi, sizeof(P4) * (size_t)n);
  voxel_grid(in, leaf, out);
  if (!out.empty()) std::memcpy(out_xyzi, out.data(), sizeof(P4) * out.size());
  return (int)out.size();
}

int lo_knn(const float* cloud, int n, const float* query, int nq, int k, int* idx, float* d2) {
  std::vector<P4> c(n);
  if (n) std::memcpy(c.data(), cloud, sizeof(P4) * (size_t)n);
  oknn::KdTree t;
  t.setInputCloud(c);
  for (int q = 0; q < nq; ++q) {
    P4 p; p.x = query[4 * q]; p.y = query[4 * q + 1]; p.z = query[4 * q + 2]; p.i = 0;
    t.nearestKSearch(p, k, idx + (size_t)q * k, d2 + (size_t)q * k);
  }
  return 0;
}

/* ---- the reference's own threading (main.cpp:37-47, channel.h): ImageProjection on the caller's thread,
 * FeatureAssociation and MapOptimization on a thread each, joined by one-slot channels whose send blocks until the slot
 * is free (bag mode, main.cpp:37-38).  Three oracle objects play the three stage objects; what crosses a channel is a
 * copy of exactly the members ProjectionOut / AssociationOut carry (utility.h:64-80). ---- */
}  /* extern "C" */
namespace {

struct ProjectionPayload {
  std::vector<P4> segmented_cloud, outlier_cloud;
  std::vector<int> start_ring, end_ring;
  float start_ori, end_ori, ori_diff;
  std::vector<uint8_t> seg_ground_flag;
  std::vector<uint32_t> seg_col_ind;
  std::vector<float> seg_range;
  int frame;
};
struct AssociationPayload {
  std::vector<P4> corner_last, surf_last, outlier_last;
  float laser_odometry[6];
  int frame;
};

template <typename T>
class SlotChannel { /* channel.h:11-56 with blocking send */
 public:
  void send(T&& item) {
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&] { return empty_; });
    item_ = std::move(item);
    empty_ = false;
    cv_.notify_all();
  }
  void receive(T& item) {
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [&] { return !empty_; });
    item = std::move(item_);
    empty_ = true;
    cv_.notify_all();
  }
 private:
  T item_;
  bool empty_ = true;
  std::mutex m_;
  std::condition_variable cv_;
};

}  // namespace

extern "C" {

/* Runs frames [0, n_frames) of one sequence through three stage threads.  h_ip / h_fa / h_mo: the three stage objects
 * (h_mo may already hold key frames and poses).  scans[f]: xyzi of frame f, counts[f] points.  stage_ms[f*3 + k]:
 * milliseconds stage k (0 IP, 1 FA, 2 MO; MO = 0 on frames that are not handed over) spent on frame f; *wall_s: seconds from
 * the first frame >= first_timed_frame entering ImageProjection until the last stage is idle. */
int lo_run_pipeline(lo_handle* h_ip, lo_handle* h_fa, lo_handle* h_mo, const float* const* scans, const int* counts, int n_frames,
                    int first_timed_frame, double* stage_ms, double* wall_s) {
  SlotChannel<ProjectionPayload> ch1;
  SlotChannel<AssociationPayload> ch2;
  for (int i = 0; i < n_frames * 3; ++i) stage_ms[i] = 0.0;
  std::thread fa([&] {
    for (;;) {
      ProjectionPayload in;
      ch1.receive(in);
      if (in.frame < 0) break;
      const double t0 = now_s();
      h_fa->segmented_cloud = std::move(in.segmented_cloud); h_fa->outlier_cloud = std::move(in.outlier_cloud);
      h_fa->start_ring = std::move(in.start_ring); h_fa->end_ring = std::move(in.end_ring);
      h_fa->start_ori = in.start_ori; h_fa->end_ori = in.end_ori; h_fa->ori_diff = in.ori_diff;
      h_fa->seg_ground_flag = std::move(in.seg_ground_flag); h_fa->seg_col_ind = std::move(in.seg_col_ind);
      h_fa->seg_range = std::move(in.seg_range);
      const int handed = h_fa->runFeatureAssociationOnce();
      AssociationPayload out;
      if (handed == 1) { /* featureAssociation.cpp:1434-1447: deep copies */
        out.corner_last = h_fa->cornerLast; out.surf_last = h_fa->surfLast; out.outlier_last = h_fa->outlierLast;
        for (int i = 0; i < 6; ++i) out.laser_odometry[i] = h_fa->transformSum[i];
        out.frame = in.frame;
      }
      stage_ms[in.frame * 3 + 1] = 1e3 * (now_s() - t0);
      if (handed == 1) ch2.send(std::move(out));
    }
    AssociationPayload stop;
    stop.frame = -1;
    ch2.send(std::move(stop));
  });
  std::thread mo([&] {
    for (;;) {
      AssociationPayload in;
      ch2.receive(in);
      if (in.frame < 0) break;
      const double t0 = now_s();
      h_mo->cornerLast = std::move(in.corner_last); h_mo->surfLast = std::move(in.surf_last); h_mo->outlierLast = std::move(in.outlier_last);
      for (int i = 0; i < 6; ++i) h_mo->transformSum[i] = in.laser_odometry[i];  /* OdometryToTransform, mapOptmization.cpp:1539 */
      h_mo->mappingCycle();
      stage_ms[in.frame * 3 + 2] = 1e3 * (now_s() - t0);
    }
  });
  double t_start = 0.0;
  for (int f = 0; f < n_frames; ++f) {
    if (f == first_timed_frame) t_start = now_s();
    const double t0 = now_s();
    h_ip->cloudHandler(scans[f], counts[f]);
    ProjectionPayload out; /* imageProjection.cpp:538-547: the filled clouds travel, the stage gets fresh ones */
    out.segmented_cloud = h_ip->segmented_cloud; out.outlier_cloud = h_ip->outlier_cloud;
    out.start_ring = h_ip->start_ring; out.end_ring = h_ip->end_ring;
    out.start_ori = h_ip->start_ori; out.end_ori = h_ip->end_ori; out.ori_diff = h_ip->ori_diff;
    out.seg_ground_flag = h_ip->seg_ground_flag; out.seg_col_ind = h_ip->seg_col_ind; out.seg_range = h_ip->seg_range;
    out.frame = f;
    stage_ms[f * 3 + 0] = 1e3 * (now_s() - t0);
    ch1.send(std::move(out));
  }
  ProjectionPayload stop;
  stop.frame = -1;
  ch1.send(std::move(stop));
  fa.join();
  mo.join();
  *wall_s = now_s() - t_start;
  return 0;
}

void lo_get_timers(lo_handle* h, double* sec5) { for (int i = 0; i < 5; ++i) sec5[i] = h->timers[i]; }
void lo_reset_timers(lo_handle* h) { for (int i = 0; i < 5; ++i) h->timers[i] = 0; }

}  // extern "C"
