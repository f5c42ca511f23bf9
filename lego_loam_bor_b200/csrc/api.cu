// api.cu -- the C ABI of include/lego_loam_b200.h: handle lifetime, device memory, stage sequencing.
// No CPU fallback anywhere: without a usable CUDA device ll_create fails with LL_ERR_NO_DEVICE.
#include <math.h>
#include <cmath>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "ll_device.cuh"
#include "ll_kernels.h"

struct ll_handle {
  LegoLoamParams prm;
  DevState st;
  LaunchCtx ctx;
  // MapOptimization runs on a stream of its own, like the reference's MapOptimization thread (mapOptmization.cpp:122):
  // a mapping cycle overlaps the next scans of FeatureAssociation; events order the two where they share data
  LaunchCtx ctx_map;
  cudaEvent_t ev_a2b = nullptr, ev_ds_done = nullptr, ev_b_tail = nullptr;
  cudaEvent_t pose_ev_b[2] = {nullptr, nullptr};
  int device = 0;
  bool own_stream = false;
  std::vector<void*> allocs;
  std::string err;
  int64_t frames = 0;        // frames seen by ll_feature_association
  int64_t odom_cycles = 0;   // _cycle_count of featureAssociation.cpp:1429-1433
  bool map_set = false;
  bool handed_to_mapping = false;
  bool labels_pending = false;  // _label_mat of the last frame still lacks the numbers of its non-root cells (k_label_final runs on demand)
  // pinned staging ring for the per-sequence point counts (a slot is reused only after its copy ran)
  int32_t* h_n_in = nullptr;
  cudaEvent_t slot_ev[16] = {};
  bool slot_used[16] = {false};
  int slot = 0;
  // ll_set_scans_host double-buffers the input on its own copy stream, so the H2D copy of scan f+1
  // overlaps the kernels of scan f when the caller sets the next scan before reading back the pose
  float4* in_buf[2] = {nullptr, nullptr};
  int* n_in_buf[2] = {nullptr, nullptr};
  int* n_in_default = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copied[2] = {}, consumed[2] = {};
  bool consumed_valid[2] = {false, false};
  uint8_t* raw_buf[2] = {nullptr, nullptr};  // ll_set_scans_pointcloud2_host: raw message bytes, [B][raw_stride]
  size_t raw_stride = 0;
  int* n_raw_buf[2] = {nullptr, nullptr};
  int* pc2_tile_cnt = nullptr;
  bool buf_xyz3[2] = {false, false};  // the staged scans are packed 12-byte points (ll_set_scans_xyz_host)
  int wr = 0;        // buffer the next ll_set_scans_host writes
  int pending = -1;  // buffer waiting to be consumed by ll_image_projection
  bool timing = false;
  cudaEvent_t ev[6] = {};
  cudaEvent_t pose_ev[2] = {nullptr, nullptr};  // ll_get_poses_async / ll_wait_poses: up to two read-backs in flight
  int pose_head = 0, pose_pending = 0;
  float stage_ms[5];
  bool ev_valid = false;
  int* trace_knn = nullptr;   // ll_enable_index_trace
  int* trace_odom = nullptr;
  // capacity error bits of the key-frame store, read back one mapping cycle late without waiting (ll_mapping_cycle)
  int32_t* h_kf_err = nullptr;
  cudaEvent_t kf_err_ev = nullptr;
  bool kf_err_pending = false;
};

namespace {

#define CK(call)                                                                     \
  do {                                                                               \
    const cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                        \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e__);                  \
      return LL_ERR_CUDA;                                                            \
    }                                                                                \
  } while (0)

// the mapping stream sees everything enqueued on the frame stream so far
int map_waits_frames(ll_handle* h) {
  CK(cudaEventRecord(h->ev_a2b, h->ctx.stream));
  CK(cudaStreamWaitEvent(h->ctx_map.stream, h->ev_a2b, 0));
  return LL_OK;
}
// the frame stream waits for everything enqueued on the mapping stream so far
int frames_wait_map(ll_handle* h) {
  CK(cudaEventRecord(h->ev_b_tail, h->ctx_map.stream));
  CK(cudaStreamWaitEvent(h->ctx.stream, h->ev_b_tail, 0));
  return LL_OK;
}
int sync_all(ll_handle* h) {
  CK(cudaStreamSynchronize(h->ctx.stream));
  CK(cudaStreamSynchronize(h->ctx_map.stream));
  return LL_OK;
}

template <typename T>
cudaError_t dev_alloc(ll_handle* h, T** p, size_t n, bool zero = true) {
  void* q = nullptr;
  const size_t bytes = (n ? n : 1) * sizeof(T);
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess) return e;
  h->allocs.push_back(q);
  if (zero) e = cudaMemsetAsync(q, 0, bytes, h->ctx.stream);
  *p = (T*)q;
  return e;
}

int next_pow2(int v) {
  int n = 1;
  while (n < v) n <<= 1;
  return n;
}

// release one device allocation of the handle early (a superseded map / grid)
template <typename T>
void dev_free(ll_handle* h, T*& p) {
  if (!p) return;
  for (size_t i = 0; i < h->allocs.size(); ++i)
    if (h->allocs[i] == (void*)p) { h->allocs.erase(h->allocs.begin() + i); break; }
  cudaFree((void*)p);
  p = nullptr;
}

void free_grid(ll_handle* h, HashGrid* g) {
  dev_free(h, g->cell_start); dev_free(h, g->cnt); dev_free(h, g->occ);
  dev_free(h, g->tile_tot); dev_free(h, g->sorted); dev_free(h, g->count); dev_free(h, g->sig);
}

int alloc_grid(ll_handle* h, HashGrid* g, int B, int cap, float cell, int expected_points = 0, bool ring_sig = false) {
  g->cell = cell;
  g->inv_cell = 1.0f / cell;
  g->cap = cap;
  // buckets ~ 1..2x the points actually indexed: a low load factor keeps empty cells cheap, but every
  // build scans the whole table, so it is sized from the expected fill rather than from the capacity
  const int want = expected_points > 0 ? expected_points : cap;
  g->tbl = next_pow2(want < 4096 ? 4096 : want);
  g->ntiles = g->tbl / 4096;
  CK(dev_alloc(h, &g->cell_start, (size_t)B * grid_cs_stride(*g)));
  CK(dev_alloc(h, &g->cnt, (size_t)B * g->tbl));
  CK(dev_alloc(h, &g->occ, (size_t)B * (g->tbl / 32)));
  CK(dev_alloc(h, &g->tile_tot, (size_t)B * g->ntiles));
  CK(dev_alloc(h, &g->sorted, (size_t)B * cap));
  CK(dev_alloc(h, &g->count, (size_t)B));
  g->sig = nullptr;
  if (ring_sig) CK(dev_alloc(h, &g->sig, (size_t)B * g->tbl));
  return LL_OK;
}

int stage_counts(ll_handle* h, const int32_t* n_points, int stride_points, const char* who, int* dst, cudaStream_t stream) {
  DevState& st = h->st;
  const int B = st.p.B;
  const int slot = h->slot;
  h->slot = (h->slot + 1) % 16;
  if (h->slot_used[slot]) CK(cudaEventSynchronize(h->slot_ev[slot]));
  int32_t* stage = h->h_n_in + (size_t)slot * B;
  for (int s = 0; s < B; ++s) {
    if (n_points[s] < 0 || n_points[s] > stride_points) { h->err = std::string(who) + ": n_points out of range"; return LL_ERR_INVALID_ARG; }
    stage[s] = n_points[s];
  }
  CK(cudaMemcpyAsync(dst, stage, sizeof(int32_t) * B, cudaMemcpyHostToDevice, stream));
  CK(cudaEventRecord(h->slot_ev[slot], stream));
  h->slot_used[slot] = true;
  return LL_OK;
}

// cloudKeyPoses3D/6D, the key-frame clouds and the surrounding cache start empty; PointType() positions are zero
int reset_keyframes(ll_handle* h) {
  KeyframeStore& kf = h->st.kf;
  const int B = h->st.p.B;
  cudaStream_t sm = h->ctx.stream;
  CK(cudaMemsetAsync(kf.kf_count, 0, (size_t)B * 4, sm)); CK(cudaMemsetAsync(kf.pool_used, 0, (size_t)B * 4, sm));
  CK(cudaMemsetAsync(kf.kf_new, 0xff, (size_t)B * 4, sm)); CK(cudaMemsetAsync(kf.robot_pos, 0, (size_t)B * 32, sm));
  CK(cudaMemsetAsync(kf.transform_last, 0, (size_t)B * 24, sm)); CK(cudaMemsetAsync(kf.sur_n, 0, (size_t)B * 4, sm));
  CK(cudaMemsetAsync(kf.sur_first, 0, (size_t)B * 4, sm)); CK(cudaMemsetAsync(kf.sur_rebuild, 0, (size_t)B * 4, sm));
  CK(cudaMemsetAsync(kf.sur_valid, 0, (size_t)B * 4, sm)); CK(cudaMemsetAsync(kf.err, 0, (size_t)B * 4, sm));
  CK(cudaMemsetAsync(kf.sur_n_erased, 0, (size_t)B * 4, sm)); CK(cudaMemsetAsync(kf.sur_last_erased, 0, (size_t)B * 4, sm));
  for (int m = 0; m < 2; ++m) {
    const VoxTable& t = kf.tbl[m];
    CK(cudaMemsetAsync(t.key, 0xff, (size_t)B * t.parts * t.sub_cap * 8, sm));
    CK(cudaMemsetAsync(t.list_n, 0, (size_t)B * t.parts * 4, sm));
    CK(cudaMemsetAsync(t.list_done, 0, (size_t)B * t.parts * 4, sm));
  }
  CK(cudaMemsetAsync(kf.xs_cur, 0, (size_t)B * 2 * 4, sm)); CK(cudaMemsetAsync(kf.xs_n, 0, (size_t)B * 2 * 4, sm));
  CK(cudaMemsetAsync(kf.xs_merge, 0, (size_t)B * 2 * 16, sm)); CK(cudaMemsetAsync(kf.xs_bbox, 0, (size_t)B * 2 * 32, sm));
  CK(cudaMemsetAsync(h->st.map_counts, 0, (size_t)B * 8, sm));
  h->kf_err_pending = false;
  return LL_OK;
}

int check_stream(ll_handle* h, const char* where) {
  for (const LaunchCtx* c : {&h->ctx, &h->ctx_map})
    if (c->first_error != cudaSuccess) {
      h->err = std::string(where) + ": launch of " + (c->first_error_kernel ? c->first_error_kernel : "?") +
               " failed: " + cudaGetErrorString(c->first_error);
      return LL_ERR_CUDA;
    }
  return LL_OK;
}

}  // namespace

extern "C" {

void ll_default_params(LegoLoamParams* p) {
  // LeGO-LOAM/config/loam_config.yaml:5-35
  p->num_vertical_scans = 16; p->num_horizontal_scans = 1800; p->ground_scan_index = 7;
  p->vertical_angle_bottom = -15.f; p->vertical_angle_top = 15.f; p->sensor_mount_angle = 0.f; p->scan_period = 0.1f;
  p->segment_valid_point_num = 5; p->segment_valid_line_num = 3; p->segment_theta = 60.0f;
  p->edge_threshold = 0.1f; p->surf_threshold = 0.1f; p->nearest_feature_search_distance = 5.f;
  p->enable_loop_closure = 0; p->mapping_frequency_divider = 5;
  p->surrounding_keyframe_search_radius = 50.0f; p->surrounding_keyframe_search_num = 50;
  p->history_keyframe_search_radius = 7.0f; p->history_keyframe_search_num = 25;
  p->history_keyframe_fitness_score = 0.3f; p->global_map_visualization_search_radius = 500.0f;
}

int ll_create(const LegoLoamParams* prm, int batch, int max_points, int device, void* cuda_stream, ll_handle** out) {
  if (!prm || !out || batch < 1 || max_points < 1) return LL_ERR_INVALID_ARG;
  if (prm->num_vertical_scans < 2 || prm->num_vertical_scans > LL_MAX_RINGS || prm->num_horizontal_scans < 16 ||
      prm->ground_scan_index < 0 || prm->ground_scan_index >= prm->num_vertical_scans)
    return LL_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return LL_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return LL_ERR_NO_DEVICE;
  ll_handle* h = new ll_handle();
  // (every failure below releases what has been allocated so far)
#undef CK
#define CK(call) do { const cudaError_t e__ = (call); if (e__ != cudaSuccess) { ll_destroy(h); return LL_ERR_CUDA; } } while (0)
  h->prm = *prm;
  h->device = device;
  if (cuda_stream) {
    h->ctx.stream = (cudaStream_t)cuda_stream;
  } else {
    if (cudaStreamCreateWithFlags(&h->ctx.stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return LL_ERR_CUDA; }
    h->own_stream = true;
  }
  if (cudaStreamCreateWithFlags(&h->ctx_map.stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return LL_ERR_CUDA; }
  if (cudaEventCreateWithFlags(&h->ev_a2b, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&h->ev_ds_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_b_tail, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&h->pose_ev_b[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->pose_ev_b[1], cudaEventDisableTiming) != cudaSuccess) { delete h; return LL_ERR_CUDA; }
  memset(&h->st, 0, sizeof(DevState));
  DevState& st = h->st;
  DevParams& p = st.p;
  const int B = batch, V = prm->num_vertical_scans, H = prm->num_horizontal_scans, N = V * H;
  p.B = B; p.V = V; p.H = H; p.N = N; p.max_pts = max_points;
  // imageProjection.cpp:64-84 with the promotion rules of SURVEY.md section 10
  const double DEG_TO_RAD = LL_PI / 180.0;
  const float bottom = prm->vertical_angle_bottom, top = prm->vertical_angle_top;
  p.ang_res_x = (float)((LL_PI * 2) / (H));
  p.ang_res_y = (float)(DEG_TO_RAD * (double)(top - bottom) / (double)(float)(V - 1));
  p.ang_bottom = (float)(-((double)bottom - 0.1) * DEG_TO_RAD);
  p.sensor_mount_angle = (float)((double)prm->sensor_mount_angle * DEG_TO_RAD);
  const float seg_theta = (float)((double)prm->segment_theta * DEG_TO_RAD);
  p.seg_tan_theta = ll_tanf(seg_theta);              // imageProjection.cpp:414
  p.sin_ax = ll_sinf(p.ang_res_x); p.cos_ax = ll_cosf(p.ang_res_x);  // :462-463
  p.sin_ay = ll_sinf(p.ang_res_y); p.cos_ay = ll_cosf(p.ang_res_y);
  p.gsi = prm->ground_scan_index;
  p.seg_valid_point_num = prm->segment_valid_point_num;
  p.seg_valid_line_num = prm->segment_valid_line_num;
  p.scan_period = prm->scan_period;
  p.edge_threshold = prm->edge_threshold;
  p.surf_threshold = prm->surf_threshold;
  p.nearest_feature_dist_sqr = prm->nearest_feature_search_distance * prm->nearest_feature_search_distance;
  p.cap_sharp = 12 * V; p.cap_less_sharp = 120 * V; p.cap_flat = 24 * V;
  st.frame_tag = 0;
  st.cap_outlier = V * ((H + 4) / 5);
  st.cap_map_corner = 0; st.cap_map_surf = 0;
  st.map_max_blocks = 96;
  const size_t BN = (size_t)B * N;
  CK(dev_alloc(h, &h->n_in_default, B));
  st.n_in = h->n_in_default;
  for (int b = 0; b < 2; ++b) {
    CK(dev_alloc(h, &h->in_buf[b], (size_t)B * max_points, false));
    CK(dev_alloc(h, &h->n_in_buf[b], B));
    CK(cudaEventCreateWithFlags(&h->copied[b], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->consumed[b], cudaEventDisableTiming));
  }
  CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  CK(cudaMallocHost((void**)&h->h_n_in, sizeof(int32_t) * B * 16));
  for (int i = 0; i < 16; ++i) CK(cudaEventCreateWithFlags(&h->slot_ev[i], cudaEventDisableTiming));
  CK(dev_alloc(h, &st.winner, BN));
  CK(dev_alloc(h, &st.range_mat, BN)); CK(dev_alloc(h, &st.full_cloud, BN));
  CK(dev_alloc(h, &st.ground_mat, BN)); CK(dev_alloc(h, &st.label_mat, BN));
  CK(dev_alloc(h, &st.parent, BN)); CK(dev_alloc(h, &st.comp_size, BN)); CK(dev_alloc(h, &st.comp_rows, BN));
  CK(dev_alloc(h, &st.tile_counts, (size_t)B * V * 4));
  CK(dev_alloc(h, &st.orientation, (size_t)B * 4)); CK(dev_alloc(h, &st.half_idx, B));
  CK(dev_alloc(h, &st.seg_cloud, BN)); CK(dev_alloc(h, &st.seg_range, BN));
  CK(dev_alloc(h, &st.seg_col, BN)); CK(dev_alloc(h, &st.seg_ground, BN)); CK(dev_alloc(h, &st.seg_ori, BN)); CK(dev_alloc(h, &st.seg_class, BN));
  CK(dev_alloc(h, &st.start_ring, (size_t)B * V)); CK(dev_alloc(h, &st.end_ring, (size_t)B * V));
  CK(dev_alloc(h, &st.seg_count, B));
  CK(dev_alloc(h, &st.outlier_cloud, (size_t)B * st.cap_outlier)); CK(dev_alloc(h, &st.outlier_count, B));
  CK(dev_alloc(h, &st.curvature, BN)); CK(dev_alloc(h, &st.picked, BN)); CK(dev_alloc(h, &st.cloud_label, BN));
  CK(dev_alloc(h, &st.slot4, (size_t)B * 2)); CK(dev_alloc(h, &st.sext_off, (size_t)B * V * 16));
  CK(dev_alloc(h, &st.st_sharp_ind, (size_t)B * V * 12)); CK(dev_alloc(h, &st.scan_list, BN));
  CK(dev_alloc(h, &st.st_less_sharp_ind, (size_t)B * V * 120));
  CK(dev_alloc(h, &st.st_flat_ind, (size_t)B * V * 24));
  CK(dev_alloc(h, &st.st_less_flat, BN)); CK(dev_alloc(h, &st.ring_counts, (size_t)B * V * 8));
  CK(dev_alloc(h, &st.corner_sharp, (size_t)B * p.cap_sharp)); CK(dev_alloc(h, &st.corner_sharp_ind, (size_t)B * p.cap_sharp));
  CK(dev_alloc(h, &st.corner_less_sharp, (size_t)B * p.cap_less_sharp));
  CK(dev_alloc(h, &st.corner_less_sharp_ind, (size_t)B * p.cap_less_sharp));
  CK(dev_alloc(h, &st.surf_flat, (size_t)B * p.cap_flat)); CK(dev_alloc(h, &st.surf_flat_ind, (size_t)B * p.cap_flat));
  CK(dev_alloc(h, &st.surf_less_flat, BN)); CK(dev_alloc(h, &st.feat_counts, (size_t)B * 4));
  CK(dev_alloc(h, &st.corner_last, (size_t)B * p.cap_less_sharp)); CK(dev_alloc(h, &st.surf_last, BN));
  CK(dev_alloc(h, &st.last_counts, (size_t)B * 2));
  CK(dev_alloc(h, &st.win_first, (size_t)2 * B * 2 * (LL_MAX_RINGS + 8))); CK(dev_alloc(h, &st.win_last, (size_t)2 * B * 2 * (LL_MAX_RINGS + 8)));
  CK(cudaMemsetAsync(st.win_first, 0x7f, (size_t)2 * B * 2 * (LL_MAX_RINGS + 8) * 4, h->ctx.stream));
  CK(cudaMemsetAsync(st.win_last, 0xff, (size_t)2 * B * 2 * (LL_MAX_RINGS + 8) * 4, h->ctx.stream));
  CK(dev_alloc(h, &st.odom_ga, (size_t)B * p.cap_flat)); CK(dev_alloc(h, &st.odom_ok, (size_t)B * p.cap_flat));
  CK(dev_alloc(h, &st.stage_clocks, (size_t)B * 16)); CK(dev_alloc(h, &st.ring_clocks, (size_t)B * V * 10)); CK(dev_alloc(h, &st.odom_cl, (size_t)B * p.cap_flat));
  CK(dev_alloc(h, &st.odom_gb, (size_t)B * p.cap_sharp)); CK(dev_alloc(h, &st.odom_s0, (size_t)B * p.cap_flat));
  CK(dev_alloc(h, &st.outlier_last, (size_t)B * st.cap_outlier));
  // (a roomy table for the small corner cloud: an EMPTY cell that hashes onto an occupied bucket costs the search a useless
  // bucket visit, and a feature with no neighbour within 5 m looks at all 1331 cells around it)
  { const int rc = alloc_grid(h, &st.grid_corner_last, B, p.cap_less_sharp, 1.0f, p.cap_less_sharp * 4, true); if (rc) { ll_destroy(h); return rc; } }
  { const int rc = alloc_grid(h, &st.grid_surf_last, B, N, 1.0f, N / 3, true); if (rc) { ll_destroy(h); return rc; } }
  CK(dev_alloc(h, &st.transform_cur, (size_t)B * 6)); CK(dev_alloc(h, &st.transform_sum, (size_t)B * 6));
  CK(dev_alloc(h, &st.odom_iters, (size_t)B * 2)); CK(dev_alloc(h, &st.odom_flags, (size_t)B * 4));
  CK(dev_alloc(h, &st.odom_matP, (size_t)B * 9));
  CK(dev_alloc(h, &st.map_counts, (size_t)B * 2));
  CK(dev_alloc(h, &st.scan_corner_ds, (size_t)B * p.cap_less_sharp)); CK(dev_alloc(h, &st.scan_surf_ds, BN));
  CK(dev_alloc(h, &st.scan_ds_counts, (size_t)B * 2));
  CK(dev_alloc(h, &st.transform_tobe_mapped, (size_t)B * 6)); CK(dev_alloc(h, &st.map_odom, (size_t)B * 6));
  CK(dev_alloc(h, &st.transform_bef_mapped, (size_t)B * 6)); CK(dev_alloc(h, &st.transform_aft_mapped, (size_t)B * 6));
  CK(dev_alloc(h, &st.map_iters, (size_t)B * 2)); CK(dev_alloc(h, &st.map_flags, (size_t)B * 4));
  CK(dev_alloc(h, &st.map_matP, (size_t)B * 36));
  CK(dev_alloc(h, &st.map_partials, (size_t)B * st.map_max_blocks * 28));
  CK(dev_alloc(h, &st.map_trace, (size_t)B * 10 * 34));
  st.map_knn_cap = p.cap_less_sharp + N;
  CK(dev_alloc(h, &st.map_knn_rec, (size_t)B * st.map_knn_cap * (LL_KNN_K + 1), false));
  CK(dev_alloc(h, &st.map_knn_sel, (size_t)B * st.map_knn_cap, false)); CK(dev_alloc(h, &st.map_fit, (size_t)B * st.map_knn_cap * 2, false)); CK(dev_alloc(h, &st.map_ticket, (size_t)B));
  st.vox_cap = N + st.cap_outlier;
  CK(dev_alloc(h, &st.vox_key0, (size_t)B * 3 * st.vox_cap, false)); CK(dev_alloc(h, &st.vox_key1, (size_t)B * 3 * st.vox_cap, false));
  CK(dev_alloc(h, &st.vox_val0, (size_t)B * 3 * st.vox_cap, false)); CK(dev_alloc(h, &st.vox_val1, (size_t)B * 3 * st.vox_cap, false));
  CK(dev_alloc(h, &st.vox_tmp_surf, BN, false)); CK(dev_alloc(h, &st.vox_tmp_out, (size_t)B * st.cap_outlier, false));
  CK(dev_alloc(h, &st.vox_tmp_counts, (size_t)B * 2));
  for (int i = 0; i < 6; ++i) CK(cudaEventCreate(&h->ev[i]));
  for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&h->pose_ev[i], cudaEventDisableTiming));
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  *out = h;
  return LL_OK;
#undef CK
#define CK(call)                                                                     \
  do {                                                                               \
    const cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                        \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e__);                  \
      return LL_ERR_CUDA;                                                            \
    }                                                                                \
  } while (0)
}

int ll_destroy(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  cudaSetDevice(h->device);
  if (h->ctx.stream) cudaStreamSynchronize(h->ctx.stream);
  for (void* q : h->allocs) cudaFree(q);
  if (h->h_n_in) cudaFreeHost(h->h_n_in);
  if (h->h_kf_err) cudaFreeHost(h->h_kf_err);
  if (h->kf_err_ev) cudaEventDestroy(h->kf_err_ev);
  for (int i = 0; i < 16; ++i) if (h->slot_ev[i]) cudaEventDestroy(h->slot_ev[i]);
  for (int b = 0; b < 2; ++b) { if (h->copied[b]) cudaEventDestroy(h->copied[b]); if (h->consumed[b]) cudaEventDestroy(h->consumed[b]); }
  if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
  for (int i = 0; i < 6; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  for (int i = 0; i < 2; ++i) if (h->pose_ev[i]) cudaEventDestroy(h->pose_ev[i]);
  if (h->ctx.ev_start) {
    for (int i = 0; i < LaunchCtx::kMaxTimed; ++i) { cudaEventDestroy(h->ctx.ev_start[i]); cudaEventDestroy(h->ctx.ev_stop[i]); }
    delete[] h->ctx.ev_start;
    delete[] h->ctx.ev_stop;
    delete[] h->ctx.ev_name;
  }
  if (h->ctx_map.stream) { cudaStreamSynchronize(h->ctx_map.stream); cudaStreamDestroy(h->ctx_map.stream); }
  for (cudaEvent_t e : {h->ev_a2b, h->ev_ds_done, h->ev_b_tail, h->pose_ev_b[0], h->pose_ev_b[1]}) if (e) cudaEventDestroy(e);
  if (h->ctx_map.ev_start) {
    for (int i = 0; i < LaunchCtx::kMaxTimed; ++i) { cudaEventDestroy(h->ctx_map.ev_start[i]); cudaEventDestroy(h->ctx_map.ev_stop[i]); }
    delete[] h->ctx_map.ev_start;
    delete[] h->ctx_map.ev_stop;
    delete[] h->ctx_map.ev_name;
  }
  if (h->own_stream) cudaStreamDestroy(h->ctx.stream);
  delete h;
  return LL_OK;
}

const char* ll_last_error(const ll_handle* h) { return h ? h->err.c_str() : "null handle"; }
int64_t ll_kernel_launches(const ll_handle* h) { return h ? h->ctx.launches + h->ctx_map.launches : 0; }

int ll_reset_feature_association(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc_s0 = sync_all(h); if (rc_s0) return rc_s0; }  // both streams idle before host-driven reads / writes of device state
  DevState& st = h->st;
  const DevParams& p = st.p;
  const size_t BN = (size_t)p.B * p.N;
  cudaStream_t sm = h->ctx.stream;
  // featureAssociation.cpp:96-157 initializationValue
  CK(cudaMemsetAsync(st.curvature, 0, BN * 4, sm)); CK(cudaMemsetAsync(st.picked, 0, BN * 4, sm));
  CK(cudaMemsetAsync(st.cloud_label, 0, BN * 4, sm)); CK(cudaMemsetAsync(st.slot4, 0, (size_t)p.B * 8, sm));
  CK(cudaMemsetAsync(st.transform_cur, 0, (size_t)p.B * 24, sm)); CK(cudaMemsetAsync(st.transform_sum, 0, (size_t)p.B * 24, sm));
  CK(cudaMemsetAsync(st.last_counts, 0, (size_t)p.B * 8, sm)); CK(cudaMemsetAsync(st.odom_flags, 0, (size_t)p.B * 16, sm));
  CK(cudaMemsetAsync(st.odom_iters, 0, (size_t)p.B * 8, sm)); CK(cudaMemsetAsync(st.odom_matP, 0, (size_t)p.B * 36, sm));
  CK(cudaMemsetAsync(st.grid_corner_last.cell_start, 0, (size_t)p.B * grid_cs_stride(st.grid_corner_last) * 4, sm));
  CK(cudaMemsetAsync(st.grid_corner_last.occ, 0, (size_t)p.B * (st.grid_corner_last.tbl / 32) * 4, sm));
  CK(cudaMemsetAsync(st.grid_surf_last.occ, 0, (size_t)p.B * (st.grid_surf_last.tbl / 32) * 4, sm));
  CK(cudaMemsetAsync(st.grid_surf_last.cell_start, 0, (size_t)p.B * grid_cs_stride(st.grid_surf_last) * 4, sm));
  CK(cudaMemsetAsync(st.win_first, 0x7f, (size_t)2 * p.B * 2 * (LL_MAX_RINGS + 8) * 4, sm));
  CK(cudaMemsetAsync(st.win_last, 0xff, (size_t)2 * p.B * 2 * (LL_MAX_RINGS + 8) * 4, sm));
  CK(cudaMemsetAsync(st.grid_corner_last.sig, 0, (size_t)p.B * st.grid_corner_last.tbl * 4, sm));
  CK(cudaMemsetAsync(st.grid_surf_last.sig, 0, (size_t)p.B * st.grid_surf_last.tbl * 4, sm));
  h->frames = 0;
  h->odom_cycles = 0;
  h->handed_to_mapping = false;
  return LL_OK;
}

int ll_reset(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc = ll_reset_feature_association(h); if (rc) return rc; }
  DevState& st = h->st;
  const DevParams& p = st.p;
  cudaStream_t sm = h->ctx.stream;
  CK(cudaMemsetAsync(st.transform_tobe_mapped, 0, (size_t)p.B * 24, sm)); CK(cudaMemsetAsync(st.map_odom, 0, (size_t)p.B * 24, sm));
  CK(cudaMemsetAsync(st.transform_bef_mapped, 0, (size_t)p.B * 24, sm)); CK(cudaMemsetAsync(st.transform_aft_mapped, 0, (size_t)p.B * 24, sm));
  CK(cudaMemsetAsync(st.map_flags, 0, (size_t)p.B * 16, sm)); CK(cudaMemsetAsync(st.map_matP, 0, (size_t)p.B * 144, sm));
  CK(cudaMemsetAsync(st.map_iters, 0, (size_t)p.B * 8, sm));
  if (st.kf.enabled) { const int rc = reset_keyframes(h); if (rc) return rc; }
  return LL_OK;
}

int ll_set_scans_host(ll_handle* h, const float* xyzi, const int32_t* n_points, int stride_points) {
  if (!h || !xyzi || !n_points || stride_points < 1) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  DevState& st = h->st;
  const int B = st.p.B;
  if (stride_points > st.p.max_pts) { h->err = "ll_set_scans_host: stride_points exceeds max_points"; return LL_ERR_CAPACITY; }
  const int b = h->wr;
  // the buffer may still be read by the projection kernels of two scans ago
  if (h->consumed_valid[b]) CK(cudaStreamWaitEvent(h->copy_stream, h->consumed[b], 0));
  { const int rc = stage_counts(h, n_points, stride_points, "ll_set_scans_host", h->n_in_buf[b], h->copy_stream); if (rc) return rc; }
  // ONE strided copy for the whole batch: every row carries the longest scan's bytes (scans of a batch are of similar
  // length, so little padding crosses PCIe, and the DMA engine gets one descriptor instead of one per sequence)
  {
    int max_n = 0;
    for (int s = 0; s < B; ++s) max_n = n_points[s] > max_n ? n_points[s] : max_n;
    if (max_n > 0)
      CK(cudaMemcpy2DAsync(h->in_buf[b], (size_t)st.p.max_pts * 16, xyzi, (size_t)stride_points * 16, (size_t)max_n * 16, B,
                           cudaMemcpyHostToDevice, h->copy_stream));
  }
  CK(cudaEventRecord(h->copied[b], h->copy_stream));
  h->buf_xyz3[b] = false;
  h->pending = b;
  h->wr ^= 1;
  return LL_OK;
}

int ll_set_scans_xyz_host(ll_handle* h, const float* xyz, const int32_t* n_points, int stride_points) {
  if (!h || !xyz || !n_points || stride_points < 1) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  DevState& st = h->st;
  const int B = st.p.B;
  if (stride_points > st.p.max_pts) { h->err = "ll_set_scans_xyz_host: stride_points exceeds max_points"; return LL_ERR_CAPACITY; }
  const int b = h->wr;
  if (h->consumed_valid[b]) CK(cudaStreamWaitEvent(h->copy_stream, h->consumed[b], 0));
  { const int rc = stage_counts(h, n_points, stride_points, "ll_set_scans_xyz_host", h->n_in_buf[b], h->copy_stream); if (rc) return rc; }
  // the same input buffer, packed: sequence s starts at float 3 * s * max_pts
  float* dst = reinterpret_cast<float*>(h->in_buf[b]);
  {
    int max_n = 0;
    for (int s = 0; s < B; ++s) max_n = n_points[s] > max_n ? n_points[s] : max_n;
    if (max_n > 0)
      CK(cudaMemcpy2DAsync(dst, (size_t)st.p.max_pts * 12, xyz, (size_t)stride_points * 12, (size_t)max_n * 12, B,
                           cudaMemcpyHostToDevice, h->copy_stream));
  }
  CK(cudaEventRecord(h->copied[b], h->copy_stream));
  h->buf_xyz3[b] = true;
  h->pending = b;
  h->wr ^= 1;
  return LL_OK;
}

int ll_set_scans_pointcloud2_host(ll_handle* h, const uint8_t* data, const int32_t* n_points, size_t stride_bytes, int point_step,
                                  int off_x, int off_y, int off_z, int off_intensity, int is_dense) {
  if (!h || !data || !n_points || point_step < 12) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  const int offs[4] = {off_x, off_y, off_z, off_intensity};
  for (int k = 0; k < 4; ++k) {
    if (k == 3 && offs[k] < 0) continue;  // no intensity field in the message
    if (offs[k] < 0 || offs[k] + 4 > point_step) { h->err = "ll_set_scans_pointcloud2_host: field offset outside the point record"; return LL_ERR_INVALID_ARG; }
  }
  DevState& st = h->st;
  const int B = st.p.B;
  for (int s = 0; s < B; ++s) {
    if (n_points[s] < 0 || (size_t)n_points[s] * point_step > stride_bytes) { h->err = "ll_set_scans_pointcloud2_host: n_points * point_step exceeds stride_bytes"; return LL_ERR_INVALID_ARG; }
    if (n_points[s] > st.p.max_pts) { h->err = "ll_set_scans_pointcloud2_host: message larger than max_points"; return LL_ERR_CAPACITY; }
  }
  const size_t need = (size_t)st.p.max_pts * point_step;
  if (need > h->raw_stride) {
    CK(cudaStreamSynchronize(h->copy_stream));
    for (int b = 0; b < 2; ++b) {
      CK(dev_alloc(h, &h->raw_buf[b], (size_t)B * need, false));
      if (!h->n_raw_buf[b]) CK(dev_alloc(h, &h->n_raw_buf[b], B));
    }
    if (!h->pc2_tile_cnt) CK(dev_alloc(h, &h->pc2_tile_cnt, (size_t)B * ((st.p.max_pts + pc2_tile_points() - 1) / pc2_tile_points())));
    h->raw_stride = need;
  }
  const int b = h->wr;
  if (h->consumed_valid[b]) CK(cudaStreamWaitEvent(h->copy_stream, h->consumed[b], 0));
  { const int rc = stage_counts(h, n_points, st.p.max_pts, "ll_set_scans_pointcloud2_host", h->n_raw_buf[b], h->copy_stream); if (rc) return rc; }
  {
    int max_n = 0;
    for (int s = 0; s < B; ++s) max_n = n_points[s] > max_n ? n_points[s] : max_n;
    if (max_n > 0)
      CK(cudaMemcpy2DAsync(h->raw_buf[b], h->raw_stride, data, stride_bytes, (size_t)max_n * point_step, B, cudaMemcpyHostToDevice,
                           h->copy_stream));
  }
  Pc2Args a;
  a.raw = h->raw_buf[b]; a.raw_stride = h->raw_stride; a.n_raw = h->n_raw_buf[b];
  a.point_step = point_step; a.off_x = off_x; a.off_y = off_y; a.off_z = off_z; a.off_intensity = off_intensity; a.is_dense = is_dense ? 1 : 0;
  a.out = h->in_buf[b]; a.out_stride = st.p.max_pts; a.n_out = h->n_in_buf[b];
  a.tile_cnt = h->pc2_tile_cnt; a.ntiles = (st.p.max_pts + pc2_tile_points() - 1) / pc2_tile_points();
  launch_decode_pointcloud2(h->ctx, h->copy_stream, B, a);
  CK(cudaEventRecord(h->copied[b], h->copy_stream));
  h->buf_xyz3[b] = false;
  h->pending = b;
  h->wr ^= 1;
  return check_stream(h, "ll_set_scans_pointcloud2_host");
}

int ll_set_scans_device(ll_handle* h, const float* xyzi_dev, const int32_t* n_points, int stride_points) {
  if (!h || !xyzi_dev || !n_points || stride_points < 1) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  DevState& st = h->st;
  if (stride_points > st.p.max_pts) { h->err = "ll_set_scans_device: stride_points exceeds max_points"; return LL_ERR_CAPACITY; }
  { const int rc = stage_counts(h, n_points, stride_points, "ll_set_scans_device", h->n_in_default, h->ctx.stream); if (rc) return rc; }
  st.n_in = h->n_in_default;
  st.in_pts = (const float4*)xyzi_dev;
  st.in_stride = stride_points;
  st.in_xyz3 = 0;
  h->pending = -1;
  return LL_OK;
}

int ll_image_projection(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  const int pb = h->pending;
  if (pb >= 0) {
    // scans staged by ll_set_scans_host: the kernels wait for that copy only
    CK(cudaStreamWaitEvent(h->ctx.stream, h->copied[pb], 0));
    h->st.in_pts = h->in_buf[pb];
    h->st.n_in = h->n_in_buf[pb];
    h->st.in_stride = h->st.p.max_pts;
    h->st.in_xyz3 = h->buf_xyz3[pb] ? 1 : 0;
    h->pending = -1;
  }
  if (!h->st.in_pts) { h->err = "ll_image_projection: no scans set"; return LL_ERR_STATE; }
  h->st.frame_tag += 1;
  if (h->timing) cudaEventRecord(h->ev[0], h->ctx.stream);
  launch_projection(h->ctx, h->st);
  if (pb >= 0) {
    CK(cudaEventRecord(h->consumed[pb], h->ctx.stream));  // the input buffer is only read by the projection kernels
    h->consumed_valid[pb] = true;
  }
  if (h->timing) cudaEventRecord(h->ev[1], h->ctx.stream);
  launch_segmentation(h->ctx, h->st);
  h->labels_pending = true;
  if (h->timing) cudaEventRecord(h->ev[2], h->ctx.stream);
  return check_stream(h, "ll_image_projection");
}

int ll_feature_association(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  const bool first = (h->frames == 0);
  launch_feature_extraction(h->ctx, h->st);
  if (h->timing) cudaEventRecord(h->ev[3], h->ctx.stream);
  launch_odometry(h->ctx, h->st, first);
  if (h->timing) cudaEventRecord(h->ev[4], h->ctx.stream);
  h->frames += 1;
  h->handed_to_mapping = false;
  if (!first) {
    // featureAssociation.cpp:1429-1433
    h->odom_cycles += 1;
    if (h->odom_cycles == h->prm.mapping_frequency_divider) {
      h->odom_cycles = 0;
      h->handed_to_mapping = true;
    }
  }
  const int rc = check_stream(h, "ll_feature_association");
  if (rc) return rc;
  return h->handed_to_mapping ? 1 : 0;
}

static int ensure_map_capacity(ll_handle* h, int nc, int ns) {
  DevState& st = h->st;
  const int B = st.p.B;
  if (nc <= st.cap_map_corner && ns <= st.cap_map_surf) return LL_OK;
  // (re)allocate with headroom; maps of other sequences are preserved
  const int ncap = nc > st.cap_map_corner ? (int)(nc * 1.25) + 1024 : st.cap_map_corner;
  const int scap = ns > st.cap_map_surf ? (int)(ns * 1.25) + 1024 : st.cap_map_surf;
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  float4 *mc = nullptr, *ms = nullptr;
  CK(dev_alloc(h, &mc, (size_t)B * ncap));
  CK(dev_alloc(h, &ms, (size_t)B * scap));
  if (st.map_corner) CK(cudaMemcpy2DAsync(mc, (size_t)ncap * 16, st.map_corner, (size_t)st.cap_map_corner * 16, (size_t)st.cap_map_corner * 16, B, cudaMemcpyDeviceToDevice, h->ctx.stream));
  if (st.map_surf) CK(cudaMemcpy2DAsync(ms, (size_t)scap * 16, st.map_surf, (size_t)st.cap_map_surf * 16, (size_t)st.cap_map_surf * 16, B, cudaMemcpyDeviceToDevice, h->ctx.stream));
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }   // the copies are done: the superseded maps and grids can go
  dev_free(h, st.map_corner); dev_free(h, st.map_surf);
  free_grid(h, &st.grid_map_corner); free_grid(h, &st.grid_map_surf);
  st.map_corner = mc; st.map_surf = ms;
  st.cap_map_corner = ncap; st.cap_map_surf = scap;
  { const int rc = alloc_grid(h, &st.grid_map_corner, B, ncap, 1.0f); if (rc) return rc; }
  { const int rc = alloc_grid(h, &st.grid_map_surf, B, scap, 1.0f); if (rc) return rc; }
  return LL_OK;
}

int ll_map_set_local(ll_handle* h, int seq, const float* corner, int nc, const float* surf, int ns) {
  if (!h || seq < 0 || seq >= h->st.p.B || nc < 0 || ns < 0) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc_s0 = sync_all(h); if (rc_s0) return rc_s0; }  // both streams idle before host-driven reads / writes of device state
  const int rc = ensure_map_capacity(h, nc, ns);
  if (rc) return rc;
  DevState& st = h->st;
  if (nc) CK(cudaMemcpyAsync(st.map_corner + (size_t)seq * st.cap_map_corner, corner, (size_t)nc * 16, cudaMemcpyHostToDevice, h->ctx.stream));
  if (ns) CK(cudaMemcpyAsync(st.map_surf + (size_t)seq * st.cap_map_surf, surf, (size_t)ns * 16, cudaMemcpyHostToDevice, h->ctx.stream));
  const int cnt[2] = {nc, ns};
  CK(cudaMemcpyAsync(st.map_counts + seq * 2, cnt, 8, cudaMemcpyHostToDevice, h->ctx.stream));
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  h->map_set = true;
  return LL_OK;
}

int ll_map_set_scan(ll_handle* h, int seq, const float* corner, int nc, const float* surf, int ns) {
  if (!h || seq < 0 || seq >= h->st.p.B || nc < 0 || ns < 0) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc_s0 = sync_all(h); if (rc_s0) return rc_s0; }  // both streams idle before host-driven reads / writes of device state
  DevState& st = h->st;
  if (nc > st.p.cap_less_sharp || ns > st.p.N) { h->err = "ll_map_set_scan: cloud larger than capacity"; return LL_ERR_CAPACITY; }
  if (nc) CK(cudaMemcpyAsync(st.scan_corner_ds + (size_t)seq * st.p.cap_less_sharp, corner, (size_t)nc * 16, cudaMemcpyHostToDevice, h->ctx.stream));
  if (ns) CK(cudaMemcpyAsync(st.scan_surf_ds + (size_t)seq * st.p.N, surf, (size_t)ns * 16, cudaMemcpyHostToDevice, h->ctx.stream));
  const int cnt[2] = {nc, ns};
  CK(cudaMemcpyAsync(st.scan_ds_counts + seq * 2, cnt, 8, cudaMemcpyHostToDevice, h->ctx.stream));
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  return LL_OK;
}

int ll_map_downsample_current_scan(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  // the hand-over to MapOptimization: the scan's clouds are down-sampled into buffers of their own and the odometry
  // pose that belongs to this scan is kept with them (AssociationOut::laser_odometry, mapOptmization.cpp:1539), so that
  // FeatureAssociation may integrate further frames before the mapping cycle runs
  CK(cudaMemcpyAsync(h->st.map_odom, h->st.transform_sum, (size_t)h->st.p.B * 24, cudaMemcpyDeviceToDevice, h->ctx.stream));
  // from here on the cycle runs on the mapping stream, behind everything enqueued for the scan so far; the next
  // publishCloudsLast (frame stream) waits until the last-frame clouds have been read
  { const int rc = map_waits_frames(h); if (rc) return rc; }
  launch_downsample_current_scan(h->ctx_map, h->st);
  CK(cudaEventRecord(h->ev_ds_done, h->ctx_map.stream));
  h->ctx.wait_before_publish = h->ev_ds_done;
  return check_stream(h, "ll_map_downsample_current_scan");
}

int ll_map_set_initial_guess(ll_handle* h, const float* t) {
  if (!h || !t) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc_s0 = sync_all(h); if (rc_s0) return rc_s0; }  // both streams idle before host-driven reads / writes of device state
  CK(cudaMemcpyAsync(h->st.transform_tobe_mapped, t, (size_t)h->st.p.B * 24, cudaMemcpyHostToDevice, h->ctx.stream));
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  return LL_OK;
}

int ll_map_set_initial_guess_async(ll_handle* h, const float* t) {
  if (!h || !t) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  CK(cudaMemcpyAsync(h->st.transform_tobe_mapped, t, (size_t)h->st.p.B * 24, cudaMemcpyHostToDevice, h->ctx_map.stream));
  return LL_OK;
}

int ll_map_set_poses(ll_handle* h, const float* aft, const float* bef) {
  if (!h || !aft || !bef) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc_s0 = sync_all(h); if (rc_s0) return rc_s0; }  // both streams idle before host-driven reads / writes of device state
  CK(cudaMemcpyAsync(h->st.transform_aft_mapped, aft, (size_t)h->st.p.B * 24, cudaMemcpyHostToDevice, h->ctx.stream));
  CK(cudaMemcpyAsync(h->st.transform_bef_mapped, bef, (size_t)h->st.p.B * 24, cudaMemcpyHostToDevice, h->ctx.stream));
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  return LL_OK;
}

int ll_map_set_odometry(ll_handle* h, const float* transform_sum) {
  if (!h || !transform_sum) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc_s0 = sync_all(h); if (rc_s0) return rc_s0; }  // both streams idle before host-driven reads / writes of device state
  CK(cudaMemcpyAsync(h->st.map_odom, transform_sum, (size_t)h->st.p.B * 24, cudaMemcpyHostToDevice, h->ctx.stream));
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  return LL_OK;
}

int ll_map_predict_pose(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc = map_waits_frames(h); if (rc) return rc; }
  launch_map_predict_pose(h->ctx_map, h->st);
  return check_stream(h, "ll_map_predict_pose");
}

int ll_scan_to_map(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  if (!h->map_set) { h->err = "ll_scan_to_map: no local map set"; return LL_ERR_STATE; }
  if (h->st.cap_map_corner == 0 || h->st.cap_map_surf == 0) return LL_OK;  // empty maps: the guard of :1316 fails for every sequence
  { const int rc = map_waits_frames(h); if (rc) return rc; }
  launch_scan_to_map(h->ctx_map, h->st);
  return check_stream(h, "ll_scan_to_map");
}

int ll_map_enable_keyframes(ll_handle* h, int max_keyframes, int pool_points, int max_map_corner, int max_map_surf) {
  if (!h || max_keyframes < 1 || max_keyframes > 32768 || pool_points < 1 || max_map_corner < 1 || max_map_surf < 1) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc_s0 = sync_all(h); if (rc_s0) return rc_s0; }  // both streams idle before host-driven reads / writes of device state
  if (h->prm.enable_loop_closure) { h->err = "ll_map_enable_keyframes: the loop-closure branch of extractSurroundingKeyFrames is not built"; return LL_ERR_STATE; }
  DevState& st = h->st;
  KeyframeStore& kf = st.kf;
  if (kf.enabled) { h->err = "ll_map_enable_keyframes: already enabled"; return LL_ERR_STATE; }
  const int B = st.p.B;
  {
    // at least the requested capacities (ensure_map_capacity adds headroom when it has to grow)
    const int rc = ensure_map_capacity(h, max_map_corner > st.cap_map_corner ? max_map_corner : st.cap_map_corner,
                                       max_map_surf > st.cap_map_surf ? max_map_surf : st.cap_map_surf);
    if (rc) return rc;
  }
  kf.kf_cap = max_keyframes;
  kf.pool_cap = pool_points;
  const double r = (double)h->prm.surrounding_keyframe_search_radius;
  kf.radius2 = (float)(r * r);  // nanoflann_pcl.h:163
  CK(dev_alloc(h, &kf.kf_count, (size_t)B)); CK(dev_alloc(h, &kf.kf_pose, (size_t)B * kf.kf_cap * 6));
  CK(dev_alloc(h, &kf.kf_off, (size_t)B * kf.kf_cap * 4)); CK(dev_alloc(h, &kf.pool_used, (size_t)B));
  CK(dev_alloc(h, &kf.pool_pts, (size_t)B * kf.pool_cap, false)); CK(dev_alloc(h, &kf.pool_key, (size_t)B * kf.pool_cap, false));
  CK(dev_alloc(h, &kf.pool_perm, (size_t)B * kf.pool_cap, false)); CK(dev_alloc(h, &kf.kf_new, (size_t)B));
  CK(dev_alloc(h, &kf.pool_link, (size_t)B * kf.pool_cap, false));
  CK(dev_alloc(h, &kf.sur_erased, (size_t)B * kf.kf_cap)); CK(dev_alloc(h, &kf.sur_n_erased, (size_t)B));
  CK(dev_alloc(h, &kf.sur_last_erased, (size_t)B));
  CK(dev_alloc(h, &kf.sel_k0, (size_t)B * kf.kf_cap, false)); CK(dev_alloc(h, &kf.sel_k1, (size_t)B * kf.kf_cap, false));
  CK(dev_alloc(h, &kf.sel_v0, (size_t)B * kf.kf_cap, false)); CK(dev_alloc(h, &kf.sel_v1, (size_t)B * kf.kf_cap, false));
  CK(dev_alloc(h, &kf.sel_rank_idx, (size_t)B * kf.kf_cap, false)); CK(dev_alloc(h, &kf.sel_ds_ids, (size_t)B * kf.kf_cap, false));
  CK(dev_alloc(h, &kf.sel_first_pos, (size_t)B * kf.kf_cap, false));
  CK(dev_alloc(h, &kf.robot_pos, (size_t)B * 8)); CK(dev_alloc(h, &kf.transform_last, (size_t)B * 6));
  CK(dev_alloc(h, &kf.sur_ids, (size_t)B * kf.kf_cap)); CK(dev_alloc(h, &kf.sur_n, (size_t)B));
  CK(dev_alloc(h, &kf.sur_first, (size_t)B)); CK(dev_alloc(h, &kf.sur_rebuild, (size_t)B));
  CK(dev_alloc(h, &kf.sur_valid, (size_t)B)); CK(dev_alloc(h, &kf.err, (size_t)B));
  // voxel tables: twice the map capacity in slots (they are filled to at most 3/4)
  const float leaf[2] = {0.2f, 0.4f};  // downSizeFilterCorner, downSizeFilterSurf (mapOptmization.cpp:71-72)
  const int parts[2] = {1, 4};
  const int want[2] = {st.cap_map_corner, st.cap_map_surf};
  for (int m = 0; m < 2; ++m) {
    VoxTable& t = kf.tbl[m];
    t.parts = parts[m];
    t.sub_cap = next_pow2((2 * want[m] + parts[m] - 1) / parts[m]);
    if (t.sub_cap < 1024) t.sub_cap = 1024;
    t.leaf = leaf[m];
    const size_t slots = (size_t)B * t.parts * t.sub_cap;
    CK(dev_alloc(h, &t.key, slots, false)); CK(dev_alloc(h, &t.sum, slots, false)); CK(dev_alloc(h, &t.cnt, slots, false));
    CK(dev_alloc(h, &t.list, slots, false)); CK(dev_alloc(h, &t.list_n, (size_t)B * t.parts));
    CK(dev_alloc(h, &t.head, slots, false)); CK(dev_alloc(h, &t.tail, slots, false)); CK(dev_alloc(h, &t.claim, slots)); CK(dev_alloc(h, &t.list_done, (size_t)B * t.parts));
  }
  kf.sort_cap = st.cap_map_corner > st.cap_map_surf ? st.cap_map_corner : st.cap_map_surf;
  CK(dev_alloc(h, &kf.sk0, (size_t)B * 2 * kf.sort_cap, false)); CK(dev_alloc(h, &kf.sk1, (size_t)B * 2 * kf.sort_cap, false));
  CK(dev_alloc(h, &kf.sv0, (size_t)B * 2 * kf.sort_cap, false)); CK(dev_alloc(h, &kf.sv1, (size_t)B * 2 * kf.sort_cap, false));
  for (int b = 0; b < 2; ++b) {
    CK(dev_alloc(h, &kf.xs_key[b], (size_t)B * 2 * kf.sort_cap, false)); CK(dev_alloc(h, &kf.xs_slot[b], (size_t)B * 2 * kf.sort_cap, false));
  }
  CK(dev_alloc(h, &kf.xn_key, (size_t)B * 2 * kf.sort_cap, false)); CK(dev_alloc(h, &kf.xn_slot, (size_t)B * 2 * kf.sort_cap, false));
  CK(dev_alloc(h, &kf.xs_cur, (size_t)B * 2)); CK(dev_alloc(h, &kf.xs_n, (size_t)B * 2)); CK(dev_alloc(h, &kf.xs_merge, (size_t)B * 2 * 4));
  CK(dev_alloc(h, &kf.xs_tile, (size_t)B * 2 * (kf.sort_cap / 1024 + 1))); CK(dev_alloc(h, &kf.xs_bbox, (size_t)B * 2 * 8));
  if (!h->h_kf_err) {
    CK(cudaMallocHost((void**)&h->h_kf_err, sizeof(int32_t) * B));
    memset(h->h_kf_err, 0, sizeof(int32_t) * B);
    CK(cudaEventCreateWithFlags(&h->kf_err_ev, cudaEventDisableTiming));
  }
  kf.enabled = 1;
  { const int rc = reset_keyframes(h); if (rc) return rc; }
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  h->map_set = true;
  return LL_OK;
}

int ll_map_extract_surrounding_keyframes(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  if (!h->st.kf.enabled) { h->err = "ll_map_extract_surrounding_keyframes: call ll_map_enable_keyframes first"; return LL_ERR_STATE; }
  { const int rc = map_waits_frames(h); if (rc) return rc; }
  launch_extract_surrounding_keyframes(h->ctx_map, h->st);
  return check_stream(h, "ll_map_extract_surrounding_keyframes");
}

int ll_map_save_keyframe(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  if (!h->st.kf.enabled) { h->err = "ll_map_save_keyframe: call ll_map_enable_keyframes first"; return LL_ERR_STATE; }
  { const int rc = map_waits_frames(h); if (rc) return rc; }
  launch_save_keyframe(h->ctx_map, h->st);
  return check_stream(h, "ll_map_save_keyframe");
}

int ll_mapping_cycle(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  if (!h->st.kf.enabled) { h->err = "ll_mapping_cycle: call ll_map_enable_keyframes first"; return LL_ERR_STATE; }
  // capacity errors of the previous cycle (a sequence that outgrew the key-frame slots, the point pool, a voxel table or
  // the local-map capacity) are reported here, one cycle late, so that no call ever waits for the device
  if (h->kf_err_pending && cudaEventQuery(h->kf_err_ev) == cudaSuccess) {
    h->kf_err_pending = false;
    for (int s = 0; s < h->st.p.B; ++s)
      if (h->h_kf_err[s]) {
        char msg[160];
        snprintf(msg, sizeof(msg), "ll_mapping_cycle: key-frame store of sequence %d out of capacity (LL_BUF_KEYFRAME_STATE bits %d)", s, h->h_kf_err[s]);
        h->err = msg;
        return LL_ERR_CAPACITY;
      }
  }
  // MapOptimization::run, mapOptmization.cpp:1545-1560 (downsampleCurrentScan first: it is the hand-over that fixes which
  // odometry pose the cycle works with, and it does not depend on the two steps the reference runs before it)
  int rc = ll_map_downsample_current_scan(h);
  if (rc < 0) return rc;
  rc = ll_map_predict_pose(h);
  if (rc < 0) return rc;
  rc = ll_map_extract_surrounding_keyframes(h);
  if (rc < 0) return rc;
  rc = ll_scan_to_map(h);
  if (rc < 0) return rc;
  rc = ll_map_save_keyframe(h);
  if (rc < 0) return rc;
  if (!h->kf_err_pending) {
    CK(cudaMemcpyAsync(h->h_kf_err, h->st.kf.err, sizeof(int32_t) * h->st.p.B, cudaMemcpyDeviceToHost, h->ctx_map.stream));
    CK(cudaEventRecord(h->kf_err_ev, h->ctx_map.stream));
    h->kf_err_pending = true;
  }
  return LL_OK;
}

int ll_map_download_keyframe(ll_handle* h, int seq, int keyframe, int which, void* dst, size_t dst_bytes, size_t* n_elems) {
  if (!h || seq < 0 || seq >= h->st.p.B || which < 0 || which > 2) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc_s0 = sync_all(h); if (rc_s0) return rc_s0; }  // both streams idle before host-driven reads / writes of device state
  KeyframeStore& kf = h->st.kf;
  if (!kf.enabled) { h->err = "ll_map_download_keyframe: call ll_map_enable_keyframes first"; return LL_ERR_STATE; }
  int count = 0, off[4];
  CK(cudaMemcpyAsync(&count, kf.kf_count + seq, 4, cudaMemcpyDeviceToHost, h->ctx.stream));
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  if (keyframe < 0 || keyframe >= count) return LL_ERR_INVALID_ARG;
  CK(cudaMemcpyAsync(off, kf.kf_off + ((size_t)seq * kf.kf_cap + keyframe) * 4, 16, cudaMemcpyDeviceToHost, h->ctx.stream));
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  const size_t n = (size_t)(off[which + 1] - off[which]);
  if (n_elems) *n_elems = n;
  if (!dst) return LL_OK;
  if (dst_bytes < n * 16) return LL_ERR_CAPACITY;
  if (!n) return LL_OK;
  // the pool holds the cloud sorted by voxel; undo that with the stored permutation
  std::vector<float> pts(n * 4);
  std::vector<int> perm(n);
  CK(cudaMemcpyAsync(pts.data(), kf.pool_pts + (size_t)seq * kf.pool_cap + off[which], n * 16, cudaMemcpyDeviceToHost, h->ctx.stream));
  CK(cudaMemcpyAsync(perm.data(), kf.pool_perm + (size_t)seq * kf.pool_cap + off[which], n * 4, cudaMemcpyDeviceToHost, h->ctx.stream));
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  float* out = (float*)dst;
  for (size_t u = 0; u < n; ++u) {
    if (perm[u] < 0 || (size_t)perm[u] >= n) { h->err = "ll_map_download_keyframe: corrupt permutation"; return LL_ERR_STATE; }
    memcpy(out + (size_t)perm[u] * 4, pts.data() + u * 4, 16);
  }
  return LL_OK;
}

int ll_process_scans(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  int rc = ll_image_projection(h);
  if (rc < 0) return rc;
  rc = ll_feature_association(h);
  if (rc < 0) return rc;
  if (rc == 1 && h->map_set) {
    if (h->st.kf.enabled) {
      rc = ll_mapping_cycle(h);
      if (rc < 0) return rc;
    } else {
      rc = ll_map_downsample_current_scan(h);
      if (rc < 0) return rc;
      rc = ll_map_predict_pose(h);
      if (rc < 0) return rc;
      rc = ll_scan_to_map(h);
      if (rc < 0) return rc;
    }
    if (h->timing) { frames_wait_map(h); cudaEventRecord(h->ev[5], h->ctx.stream); }  // (timing serialises the two streams)
    h->ev_valid = true;
    return 1;
  }
  if (h->timing) cudaEventRecord(h->ev[5], h->ctx.stream);
  h->ev_valid = true;
  return 0;
}

int ll_join_mapping(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  return frames_wait_map(h);
}

int ll_synchronize(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  return LL_OK;
}

int ll_get_poses(ll_handle* h, float* tsum, float* tcur, float* tmap) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc_s0 = sync_all(h); if (rc_s0) return rc_s0; }  // both streams idle before host-driven reads / writes of device state
  const size_t bytes = (size_t)h->st.p.B * 24;
  if (tsum) CK(cudaMemcpyAsync(tsum, h->st.transform_sum, bytes, cudaMemcpyDeviceToHost, h->ctx.stream));
  if (tcur) CK(cudaMemcpyAsync(tcur, h->st.transform_cur, bytes, cudaMemcpyDeviceToHost, h->ctx.stream));
  if (tmap) CK(cudaMemcpyAsync(tmap, h->st.transform_tobe_mapped, bytes, cudaMemcpyDeviceToHost, h->ctx.stream));
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  return LL_OK;
}

// tf::Quaternion::setRPY (tf/LinearMath/Quaternion.h) on (roll, pitch, yaw) = (t[2], -t[0], -t[1]), then the axis
// shuffle of featureAssociation.cpp:1290-1296 / mapOptmization.cpp:516-522
void ll_transform_to_odometry(const float* t, double* o) {
  const double roll = (double)t[2], pitch = (double)(-t[0]), yaw = (double)(-t[1]);
  const double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
  const double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
  const double qx = sr * cp * cy - cr * sp * sy;
  const double qy = cr * sp * cy + sr * cp * sy;
  const double qz = cr * cp * sy - sr * sp * cy;
  const double qw = cr * cp * cy + sr * sp * sy;
  o[0] = t[3]; o[1] = t[4]; o[2] = t[5];
  o[3] = -qy; o[4] = -qz; o[5] = qx; o[6] = qw;
}

// utility.h:96-110: tf::Matrix3x3(tf::Quaternion(o.z, -o.x, -o.y, o.w)).getRPY(roll, pitch, yaw) (Matrix3x3::setRotation
// + getEulerYPR, solution 1), transform = (-pitch, -yaw, roll, position)
void ll_odometry_to_transform(const double* o, float* t) {
  const double x = o[5], y = -o[3], z = -o[4], w = o[6];
  const double d = x * x + y * y + z * z + w * w;
  const double s = 2.0 / d;
  const double xs = x * s, ys = y * s, zs = z * s;
  const double wx = w * xs, wy = w * ys, wz = w * zs;
  const double xx = x * xs, xy = x * ys, xz = x * zs;
  const double yy = y * ys, yz = y * zs, zz = z * zs;
  const double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
  const double m01 = xy - wz, m02 = xz + wy;
  double roll, pitch, yaw;
  if (std::fabs(m20) >= 1.0) {  // gimbal lock branch of getEulerYPR
    yaw = 0.0;
    const double delta = std::atan2(m21, m22);
    if (m20 < 0.0) { pitch = 3.14159265358979323846 / 2.0; roll = delta; }
    else { pitch = -3.14159265358979323846 / 2.0; roll = delta; }
    (void)m01; (void)m02;
  } else {
    pitch = -std::asin(m20);
    roll = std::atan2(m21 / std::cos(pitch), m22 / std::cos(pitch));
    yaw = std::atan2(m10 / std::cos(pitch), m00 / std::cos(pitch));
  }
  t[0] = (float)(-pitch); t[1] = (float)(-yaw); t[2] = (float)roll;
  t[3] = (float)o[0]; t[4] = (float)o[1]; t[5] = (float)o[2];
}

int ll_get_odometry(ll_handle* h, double* laser_odometry, double* odom_aft_mapped) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc_s0 = sync_all(h); if (rc_s0) return rc_s0; }  // both streams idle before host-driven reads / writes of device state
  const int B = h->st.p.B;
  std::vector<float> ts((size_t)B * 6), ta((size_t)B * 6), tb((size_t)B * 6);
  CK(cudaMemcpyAsync(ts.data(), h->st.transform_sum, (size_t)B * 24, cudaMemcpyDeviceToHost, h->ctx.stream));
  CK(cudaMemcpyAsync(ta.data(), h->st.transform_aft_mapped, (size_t)B * 24, cudaMemcpyDeviceToHost, h->ctx.stream));
  CK(cudaMemcpyAsync(tb.data(), h->st.transform_bef_mapped, (size_t)B * 24, cudaMemcpyDeviceToHost, h->ctx.stream));
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  for (int s = 0; s < B; ++s) {
    if (laser_odometry) ll_transform_to_odometry(&ts[(size_t)s * 6], laser_odometry + (size_t)s * 7);
    if (odom_aft_mapped) {
      double* o = odom_aft_mapped + (size_t)s * 13;
      ll_transform_to_odometry(&ta[(size_t)s * 6], o);
      for (int k = 0; k < 6; ++k) o[7 + k] = (double)tb[(size_t)s * 6 + k];
    }
  }
  return LL_OK;
}

int ll_get_poses_async(ll_handle* h, float* tsum, float* tcur, float* tmap) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  if (h->pose_pending >= 2) { h->err = "ll_get_poses_async: two read-backs already in flight, call ll_wait_poses"; return LL_ERR_STATE; }
  const size_t bytes = (size_t)h->st.p.B * 24;
  if (tsum) CK(cudaMemcpyAsync(tsum, h->st.transform_sum, bytes, cudaMemcpyDeviceToHost, h->ctx.stream));
  if (tcur) CK(cudaMemcpyAsync(tcur, h->st.transform_cur, bytes, cudaMemcpyDeviceToHost, h->ctx.stream));
  // transformTobeMapped belongs to MapOptimization: read behind the mapping cycles enqueued so far (it is the pose of the
  // last cycle that has been handed over, like a subscriber of /aft_mapped_to_init sees it)
  if (tmap) CK(cudaMemcpyAsync(tmap, h->st.transform_tobe_mapped, bytes, cudaMemcpyDeviceToHost, h->ctx_map.stream));
  CK(cudaEventRecord(h->pose_ev[(h->pose_head + h->pose_pending) & 1], h->ctx.stream));
  CK(cudaEventRecord(h->pose_ev_b[(h->pose_head + h->pose_pending) & 1], h->ctx_map.stream));
  h->pose_pending += 1;
  return LL_OK;
}

int ll_wait_poses(ll_handle* h) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  if (h->pose_pending == 0) { h->err = "ll_wait_poses: no ll_get_poses_async pending"; return LL_ERR_STATE; }
  CK(cudaEventSynchronize(h->pose_ev[h->pose_head]));  // the oldest one
  CK(cudaEventSynchronize(h->pose_ev_b[h->pose_head]));
  h->pose_head ^= 1;
  h->pose_pending -= 1;
  return LL_OK;
}

int ll_time_kernel(ll_handle* h, const char* kernel_name) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc = sync_all(h); if (rc) return rc; }
  for (LaunchCtx* cp : {&h->ctx, &h->ctx_map}) {
    LaunchCtx& c = *cp;
    if (!c.ev_start) {
      c.ev_start = new cudaEvent_t[LaunchCtx::kMaxTimed];
      c.ev_stop = new cudaEvent_t[LaunchCtx::kMaxTimed];
      c.ev_name = new const char*[LaunchCtx::kMaxTimed];
      for (int i = 0; i < LaunchCtx::kMaxTimed; ++i) { CK(cudaEventCreate(&c.ev_start[i])); CK(cudaEventCreate(&c.ev_stop[i])); }
    }
    c.timed_used = 0;
    memset(c.timed_name, 0, sizeof(c.timed_name));
    if (kernel_name) strncpy(c.timed_name, kernel_name, sizeof(c.timed_name) - 1);
  }
  return LL_OK;
}

int ll_get_kernel_time(ll_handle* h, double* total_ms, int* launches) {
  if (!h || !total_ms || !launches) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc = sync_all(h); if (rc) return rc; }
  double tot = 0.0;
  int n = 0;
  for (LaunchCtx* cp : {&h->ctx, &h->ctx_map}) {
    LaunchCtx& c = *cp;
    for (int i = 0; i < c.timed_used; ++i) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, c.ev_start[i], c.ev_stop[i]));
      tot += ms;
    }
    n += c.timed_used;
  }
  *total_ms = tot;
  *launches = n;
  return LL_OK;
}

int ll_get_kernel_time_table(ll_handle* h, char* buf, size_t cap) {
  if (!h || !buf || cap < 2) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc = sync_all(h); if (rc) return rc; }
  std::vector<std::string> names;
  std::vector<double> ms;
  std::vector<int> cnt;
  for (LaunchCtx* cp : {&h->ctx, &h->ctx_map}) {
    LaunchCtx& c = *cp;
    for (int i = 0; i < c.timed_used; ++i) {
      float t = 0.f;
      CK(cudaEventElapsedTime(&t, c.ev_start[i], c.ev_stop[i]));
      size_t k = 0;
      for (; k < names.size(); ++k) if (names[k] == c.ev_name[i]) break;
      if (k == names.size()) { names.push_back(c.ev_name[i]); ms.push_back(0.0); cnt.push_back(0); }
      ms[k] += t;
      cnt[k] += 1;
    }
  }
  std::string out;
  char line[160];
  for (size_t k = 0; k < names.size(); ++k) {
    snprintf(line, sizeof(line), "%s %.6f %d\n", names[k].c_str(), ms[k], cnt[k]);
    out += line;
  }
  if (out.size() + 1 > cap) return LL_ERR_CAPACITY;
  memcpy(buf, out.c_str(), out.size() + 1);
  return LL_OK;
}

int ll_enable_index_trace(ll_handle* h, int enable) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  DevState& st = h->st;
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  if (enable) {
    if (!h->trace_knn) {
      CK(dev_alloc(h, &h->trace_knn, (size_t)st.p.B * 10 * st.map_knn_cap * 5, false));
      CK(dev_alloc(h, &h->trace_odom, (size_t)st.p.B * 2 * 5 * st.p.cap_flat * 3, false));
      CK(cudaMemsetAsync(h->trace_knn, 0xff, (size_t)st.p.B * 10 * st.map_knn_cap * 5 * 4, h->ctx.stream));
      CK(cudaMemsetAsync(h->trace_odom, 0xff, (size_t)st.p.B * 2 * 5 * st.p.cap_flat * 3 * 4, h->ctx.stream));
    }
    st.map_knn_trace = h->trace_knn;
    st.odom_trace = h->trace_odom;
  } else {
    st.map_knn_trace = nullptr;
    st.odom_trace = nullptr;
  }
  return LL_OK;
}

int ll_enable_stage_timing(ll_handle* h, int enable) {
  if (!h) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  h->timing = enable != 0;
  h->st.stage_clocks_on = enable != 0 ? 1 : 0;
  h->ev_valid = false;
  return LL_OK;
}

int ll_get_stage_times_ms(ll_handle* h, float* ms5) {
  if (!h || !ms5) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  if (!h->timing || !h->ev_valid) { h->err = "ll_get_stage_times_ms: timing not enabled or no frame processed"; return LL_ERR_STATE; }
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  for (int i = 0; i < 5; ++i) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]));
    ms5[i] = ms;
  }
  return LL_OK;
}

static int fetch_count(ll_handle* h, const int* dev, int* out) {
  CK(cudaMemcpyAsync(out, dev, sizeof(int), cudaMemcpyDeviceToHost, h->ctx.stream));
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  return LL_OK;
}

int ll_download(ll_handle* h, int seq, int buffer, void* dst, size_t dst_bytes, size_t* n_elems) {
  if (!h || seq < 0 || seq >= h->st.p.B) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc_s0 = sync_all(h); if (rc_s0) return rc_s0; }  // both streams idle before host-driven reads / writes of device state
  DevState& st = h->st;
  const DevParams& p = st.p;
  const size_t N = p.N;
  const void* src = nullptr;
  size_t elem = 0, n = 0;
  int cnt = 0;
#define FIXED(ptr, esz, count) do { src = (const char*)(ptr) + (size_t)seq * (count) * (esz); elem = (esz); n = (count); } while (0)
#define COUNTED(ptr, esz, stride, cntptr) do { const int rc__ = fetch_count(h, (cntptr), &cnt); if (rc__) return rc__; \
    src = (const char*)(ptr) + (size_t)seq * (stride) * (esz); elem = (esz); n = (size_t)cnt; } while (0)
  switch (buffer) {
    case LL_BUF_RANGE_MAT: FIXED(st.range_mat, 4, N); break;
    case LL_BUF_FULL_CLOUD: FIXED(st.full_cloud, 16, N); break;
    case LL_BUF_GROUND_MAT: FIXED(st.ground_mat, 1, N); break;
    case LL_BUF_LABEL_MAT:
      // labelMat is private to ImageProjection and cloudSegmentation only tests it for > 0 / == 999999 (imageProjection.cpp:
      // 363-383), which the device decides from the union-find forest; the NUMBERS of the non-root cells are filled in here
      if (h->labels_pending) {
        launch_label_final(h->ctx, st);
        h->labels_pending = false;
        { const int rc_l = sync_all(h); if (rc_l) return rc_l; }
      }
      FIXED(st.label_mat, 4, N); break;
    case LL_BUF_SEG_CLOUD: COUNTED(st.seg_cloud, 16, N, st.seg_count + seq); break;
    case LL_BUF_SEG_GROUND_FLAG: COUNTED(st.seg_ground, 1, N, st.seg_count + seq); break;
    case LL_BUF_SEG_COL_IND: COUNTED(st.seg_col, 4, N, st.seg_count + seq); break;
    case LL_BUF_SEG_RANGE: COUNTED(st.seg_range, 4, N, st.seg_count + seq); break;
    case LL_BUF_START_RING_INDEX: FIXED(st.start_ring, 4, (size_t)p.V); break;
    case LL_BUF_END_RING_INDEX: FIXED(st.end_ring, 4, (size_t)p.V); break;
    case LL_BUF_ORIENTATION: src = st.orientation + seq * 4; elem = 4; n = 3; break;
    case LL_BUF_OUTLIER_CLOUD: COUNTED(st.outlier_cloud, 16, (size_t)st.cap_outlier, st.outlier_count + seq); break;
    case LL_BUF_CLOUD_CURVATURE: FIXED(st.curvature, 4, N); break;
    case LL_BUF_NEIGHBOR_PICKED: FIXED(st.picked, 4, N); break;
    case LL_BUF_CLOUD_LABEL: FIXED(st.cloud_label, 4, N); break;
    case LL_BUF_CORNER_SHARP: COUNTED(st.corner_sharp, 16, (size_t)p.cap_sharp, st.feat_counts + seq * 4 + 0); break;
    case LL_BUF_CORNER_LESS_SHARP: COUNTED(st.corner_less_sharp, 16, (size_t)p.cap_less_sharp, st.feat_counts + seq * 4 + 1); break;
    case LL_BUF_SURF_FLAT: COUNTED(st.surf_flat, 16, (size_t)p.cap_flat, st.feat_counts + seq * 4 + 2); break;
    case LL_BUF_SURF_LESS_FLAT: COUNTED(st.surf_less_flat, 16, N, st.feat_counts + seq * 4 + 3); break;
    case LL_BUF_CORNER_SHARP_IND: COUNTED(st.corner_sharp_ind, 4, (size_t)p.cap_sharp, st.feat_counts + seq * 4 + 0); break;
    case LL_BUF_CORNER_LESS_SHARP_IND: COUNTED(st.corner_less_sharp_ind, 4, (size_t)p.cap_less_sharp, st.feat_counts + seq * 4 + 1); break;
    case LL_BUF_SURF_FLAT_IND: COUNTED(st.surf_flat_ind, 4, (size_t)p.cap_flat, st.feat_counts + seq * 4 + 2); break;
    case LL_BUF_CORNER_LAST: COUNTED(st.corner_last, 16, (size_t)p.cap_less_sharp, st.last_counts + seq * 2 + 0); break;
    case LL_BUF_SURF_LAST: COUNTED(st.surf_last, 16, N, st.last_counts + seq * 2 + 1); break;
    case LL_BUF_TRANSFORM_CUR: src = st.transform_cur + seq * 6; elem = 4; n = 6; break;
    case LL_BUF_TRANSFORM_SUM: src = st.transform_sum + seq * 6; elem = 4; n = 6; break;
    case LL_BUF_ODOM_ITERS: src = st.odom_iters + seq * 2; elem = 4; n = 2; break;
    case LL_BUF_MAP_CORNER: if (!st.map_corner) { n = 0; elem = 16; src = nullptr; break; }
      COUNTED(st.map_corner, 16, (size_t)st.cap_map_corner, st.map_counts + seq * 2 + 0); break;
    case LL_BUF_MAP_SURF: if (!st.map_surf) { n = 0; elem = 16; src = nullptr; break; }
      COUNTED(st.map_surf, 16, (size_t)st.cap_map_surf, st.map_counts + seq * 2 + 1); break;
    case LL_BUF_SCAN_CORNER_DS: COUNTED(st.scan_corner_ds, 16, (size_t)p.cap_less_sharp, st.scan_ds_counts + seq * 2 + 0); break;
    case LL_BUF_SCAN_SURF_TOTAL_DS: COUNTED(st.scan_surf_ds, 16, N, st.scan_ds_counts + seq * 2 + 1); break;
    case LL_BUF_TRANSFORM_TOBE_MAPPED: src = st.transform_tobe_mapped + seq * 6; elem = 4; n = 6; break;
    case LL_BUF_MAP_ITERS: src = st.map_iters + seq * 2; elem = 4; n = 2; break;
    case LL_BUF_STAGE_CLOCKS: src = st.stage_clocks + (size_t)seq * 16; elem = 8; n = 16; break;
    case LL_BUF_RING_CLOCKS: src = st.ring_clocks + (size_t)seq * p.V * 10; elem = 80; n = (size_t)p.V; break;
    case LL_BUF_MAP_TRACE: src = st.map_trace + (size_t)seq * 340; elem = 8; n = 340; break;
    case LL_BUF_SCAN_SURF_DS: COUNTED(st.vox_tmp_surf, 16, N, st.vox_tmp_counts + seq * 2 + 0); break;
    case LL_BUF_SCAN_OUTLIER_DS: COUNTED(st.vox_tmp_out, 16, (size_t)st.cap_outlier, st.vox_tmp_counts + seq * 2 + 1); break;
    case LL_BUF_TRANSFORM_BEF_MAPPED: src = st.transform_bef_mapped + seq * 6; elem = 4; n = 6; break;
    case LL_BUF_TRANSFORM_AFT_MAPPED: src = st.transform_aft_mapped + seq * 6; elem = 4; n = 6; break;
    case LL_BUF_OUTLIER_LAST: COUNTED(st.outlier_last, 16, (size_t)st.cap_outlier, st.odom_flags + seq * 4 + 3); break;
    case LL_BUF_INPUT_CLOUD: {
      // the scans staged by the last ll_set_scans_* call (or consumed by the last ll_image_projection)
      const float4* buf = h->pending >= 0 ? h->in_buf[h->pending] : st.in_pts;
      const int* cntp = h->pending >= 0 ? h->n_in_buf[h->pending] : st.n_in;
      const int stride = h->pending >= 0 ? st.p.max_pts : st.in_stride;
      if (!buf) return LL_ERR_STATE;
      if (h->pending >= 0) CK(cudaStreamSynchronize(h->copy_stream));
      // packed scans (ll_set_scans_xyz_host) come back as they went in: 3 floats per point
      const bool xyz3 = h->pending >= 0 ? h->buf_xyz3[h->pending] : st.in_xyz3 != 0;
      if (xyz3) COUNTED(buf, 12, (size_t)stride, cntp + seq);
      else COUNTED(buf, 16, (size_t)stride, cntp + seq);
      break;
    }
    case LL_BUF_ODOM_SEARCH_IDX:
      if (!st.odom_trace) { h->err = "ll_download: call ll_enable_index_trace first"; return LL_ERR_STATE; }
      src = st.odom_trace + (size_t)seq * 2 * 5 * p.cap_flat * 3; elem = 12; n = (size_t)2 * 5 * p.cap_flat; break;
    case LL_BUF_MAP_KNN_IDX: {
      if (!st.map_knn_trace) { h->err = "ll_download: call ll_enable_index_trace first"; return LL_ERR_STATE; }
      int qn[2];
      CK(cudaMemcpyAsync(qn, st.scan_ds_counts + seq * 2, 8, cudaMemcpyDeviceToHost, h->ctx.stream));
      { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
      const size_t Q = (size_t)qn[0] + (size_t)qn[1];
      if (n_elems) *n_elems = 10 * Q;
      if (!dst) return LL_OK;
      if (dst_bytes < 10 * Q * 20) return LL_ERR_CAPACITY;
      for (int it = 0; it < 10 && Q; ++it)
        CK(cudaMemcpyAsync((char*)dst + (size_t)it * Q * 20, st.map_knn_trace + (((size_t)seq * 10 + it) * st.map_knn_cap) * 5, Q * 20,
                           cudaMemcpyDeviceToHost, h->ctx.stream));
      { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
      return LL_OK;
    }
    case LL_BUF_KEYFRAME_STATE: {
      if (!st.kf.enabled) return LL_ERR_STATE;
      int v[4];
      CK(cudaMemcpyAsync(&v[0], st.kf.kf_count + seq, 4, cudaMemcpyDeviceToHost, h->ctx.stream));
      CK(cudaMemcpyAsync(&v[1], st.kf.sur_n + seq, 4, cudaMemcpyDeviceToHost, h->ctx.stream));
      CK(cudaMemcpyAsync(&v[2], st.kf.sur_last_erased + seq, 4, cudaMemcpyDeviceToHost, h->ctx.stream));
      CK(cudaMemcpyAsync(&v[3], st.kf.err + seq, 4, cudaMemcpyDeviceToHost, h->ctx.stream));
      { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
      if (n_elems) *n_elems = 4;
      if (!dst) return LL_OK;
      if (dst_bytes < 16) return LL_ERR_CAPACITY;
      memcpy(dst, v, 16);
      return LL_OK;
    }
    case LL_BUF_KEY_POSES_6D: if (!st.kf.enabled) return LL_ERR_STATE;
      { const int rc__ = fetch_count(h, st.kf.kf_count + seq, &cnt); if (rc__) return rc__; }
      src = st.kf.kf_pose + (size_t)seq * st.kf.kf_cap * 6; elem = 24; n = (size_t)cnt; break;
    case LL_BUF_SURROUNDING_KEY_IDS: if (!st.kf.enabled) return LL_ERR_STATE;
      COUNTED(st.kf.sur_ids, 4, (size_t)st.kf.kf_cap, st.kf.sur_n + seq); break;
    case LL_BUF_SURF_LESS_FLAT_RAW_COUNT: {
      // gathered from the per-ring counters (stride 8)
      std::vector<int> rc((size_t)p.V * 8);
      CK(cudaMemcpyAsync(rc.data(), st.ring_counts + (size_t)seq * p.V * 8, rc.size() * 4, cudaMemcpyDeviceToHost, h->ctx.stream));
      { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
      if (n_elems) *n_elems = p.V;
      if (!dst) return LL_OK;
      if (dst_bytes < (size_t)p.V * 4) return LL_ERR_CAPACITY;
      for (int i = 0; i < p.V; ++i) ((int*)dst)[i] = rc[(size_t)i * 8 + 4];
      return LL_OK;
    }
    default: return LL_ERR_INVALID_ARG;
  }
#undef FIXED
#undef COUNTED
  if (n_elems) *n_elems = n;
  if (!dst) return LL_OK;
  if (dst_bytes < n * elem) return LL_ERR_CAPACITY;
  if (n) {
    CK(cudaMemcpyAsync(dst, src, n * elem, cudaMemcpyDeviceToHost, h->ctx.stream));
    { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  }
  return LL_OK;
}

int ll_upload(ll_handle* h, int seq, int buffer, const void* src, size_t n_elems) {
  if (!h || seq < 0 || seq >= h->st.p.B || !src) return LL_ERR_INVALID_ARG;
  cudaSetDevice(h->device);  // the handle's device, whatever the caller's current one is
  { const int rc_s0 = sync_all(h); if (rc_s0) return rc_s0; }  // both streams idle before host-driven reads / writes of device state
  DevState& st = h->st;
  float* dst = nullptr;
  switch (buffer) {
    case LL_BUF_TRANSFORM_CUR: dst = st.transform_cur + seq * 6; break;
    case LL_BUF_TRANSFORM_SUM: dst = st.transform_sum + seq * 6; break;
    case LL_BUF_TRANSFORM_TOBE_MAPPED: dst = st.transform_tobe_mapped + seq * 6; break;
    case LL_BUF_TRANSFORM_BEF_MAPPED: dst = st.transform_bef_mapped + seq * 6; break;
    case LL_BUF_TRANSFORM_AFT_MAPPED: dst = st.transform_aft_mapped + seq * 6; break;
    default: return LL_ERR_INVALID_ARG;
  }
  if (n_elems != 6) return LL_ERR_INVALID_ARG;
  CK(cudaMemcpyAsync(dst, src, 24, cudaMemcpyHostToDevice, h->ctx.stream));
  { const int rc_s = sync_all(h); if (rc_s) return rc_s; }
  return LL_OK;
}

}  // extern "C"
