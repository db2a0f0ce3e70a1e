// test_shim.cpp -- the reference's per-frame flow written against include/cloud_merger_shim.hpp (GPU) and checked,
// bit for bit, against the CPU oracle (oracle/cm_oracle.h; test infrastructure, linked here only as the checker).
//
// The call sequence is the reference's own: callbackX { transformPointCloud; getROI; getCloudPart } per sensor,
// fusePointclouds (operator+=), voxelgrid (pc_preprocessing_main.cpp:318-337, :20-59, :131-177) -- first function by
// function, then through the fused FusedFrame path that a node would use in production.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#include "cloud_merger_shim.hpp"
#include "cm_oracle.h"

using namespace cloud_merger;

static int g_fail = 0;
#define CHECK(cond, ...)                         \
  do {                                           \
    if (!(cond)) {                               \
      std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
      std::printf(__VA_ARGS__);                  \
      std::printf("\n");                         \
      ++g_fail;                                  \
    }                                            \
  } while (0)

static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static double urand() {  // splitmix64
  uint64_t z = (g_rng += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (z >> 11) * (1.0 / 9007199254740992.0);
}

static Cloud make_cloud(size_t n, bool with_nan, uint64_t stamp) {
  Cloud c;
  c.points.resize(n);
  for (size_t i = 0; i < n; ++i) {
    PointXYZI& p = c.points[i];
    const double r = 1.0 + 60.0 * urand(), az = 6.283185307179586 * urand(), el = -0.4 + 0.6 * urand();
    p.x = static_cast<float>(r * std::cos(el) * std::cos(az));
    p.y = static_cast<float>(r * std::cos(el) * std::sin(az));
    p.z = static_cast<float>(r * std::sin(el));
    p.intensity = static_cast<float>(255.0 * urand());
    if (with_nan && urand() < 0.01) p.x = std::numeric_limits<float>::quiet_NaN();
  }
  c.width = static_cast<uint32_t>(n); c.height = 1; c.is_dense = !with_nan;
  set_stamp(c, stamp);
  return c;
}

static std::vector<float> packed(const Cloud& c) {
  std::vector<float> v(c.points.size() * 4);
  for (size_t i = 0; i < c.points.size(); ++i) {
    v[i * 4 + 0] = c.points[i].x; v[i * 4 + 1] = c.points[i].y; v[i * 4 + 2] = c.points[i].z; v[i * 4 + 3] = c.points[i].intensity;
  }
  return v;
}

static bool same_bits(float a, float b) { return std::memcmp(&a, &b, 4) == 0 || (std::isnan(a) && std::isnan(b)); }

static void expect_cloud(const Cloud& got, const std::vector<float>& want_xyzi, size_t n, const char* what) {
  CHECK(got.points.size() == n, "%s: %zu points, oracle %zu", what, got.points.size(), n);
  if (got.points.size() != n) return;
  size_t bad = 0;
  for (size_t i = 0; i < n; ++i) {
    const PointXYZI& p = got.points[i];
    if (!same_bits(p.x, want_xyzi[i * 4]) || !same_bits(p.y, want_xyzi[i * 4 + 1]) || !same_bits(p.z, want_xyzi[i * 4 + 2]) ||
        !same_bits(p.intensity, want_xyzi[i * 4 + 3]) || p.data3 != 1.0f)
      ++bad;
  }
  CHECK(bad == 0, "%s: %zu of %zu points differ bitwise", what, bad, n);
  CHECK(got.width == n && got.height == 1, "%s: width/height", what);
}

// oracle: chained PassThrough on packed points (getROI / getCloudPart)
static std::vector<float> oracle_passes(const std::vector<float>& in, const std::vector<cmo_pass_t>& passes) {
  std::vector<float> cur = in;
  for (const cmo_pass_t& ps : passes) {
    const int64_t n = static_cast<int64_t>(cur.size() / 4);
    std::vector<int32_t> idx(static_cast<size_t>(n) + 1);
    const int64_t k = cmo_passthrough(cur.data(), n, ps.axis, ps.lo, ps.hi, ps.negative, idx.data());
    std::vector<float> next(static_cast<size_t>(k) * 4);
    for (int64_t i = 0; i < k; ++i) std::memcpy(&next[i * 4], &cur[static_cast<size_t>(idx[i]) * 4], 16);
    cur.swap(next);
  }
  return cur;
}

int main() {
  if (cm_device_count() == 0) {
    std::printf("no CUDA device: the shim has no CPU path (expected on the build box)\n");
    Context ctx(1024);
    CHECK(!ctx.ok(), "Context must fail without a GPU");
    return g_fail ? 1 : 77;  // 77 = skipped
  }
  const Params prm;
  const int S = 3;
  const size_t N[S] = {40000, 33333, 25001};
  Transform tf[S];
  for (int s = 0; s < S; ++s) {
    const double yaw = 2.0943951023931953 * s, half = yaw / 2;
    tf[s].q[0] = 0.01 * s; tf[s].q[1] = -0.02; tf[s].q[2] = std::sin(half); tf[s].q[3] = std::cos(half);
    const double nq = std::sqrt(tf[s].q[0] * tf[s].q[0] + tf[s].q[1] * tf[s].q[1] + tf[s].q[2] * tf[s].q[2] + tf[s].q[3] * tf[s].q[3]);
    for (double& v : tf[s].q) v /= nq;
    tf[s].origin[0] = 1.2 * std::cos(yaw); tf[s].origin[1] = 1.2 * std::sin(yaw); tf[s].origin[2] = 1.9;
  }
  std::vector<Cloud> raw;
  for (int s = 0; s < S; ++s) raw.push_back(make_cloud(N[s], s == 1, 100 + 7 * s));

  const std::vector<cmo_pass_t> roi = {{2, prm.roi_z_min, prm.roi_z_max, 0},
                                       {1, -prm.roi_width / 2, prm.roi_width / 2, 0},
                                       {0, -prm.roi_mid, prm.roi_length - prm.roi_mid, 0}};

  // ---- function by function, as the reference calls them --------------------------------------------------------------
  Context ctx(1 << 16, S);
  CHECK(ctx.ok(), "Context: %s", ctx.last_error().c_str());
  Cloud::Ptr fused(new Cloud);
  std::vector<float> fused_oracle;
  for (int s = 0; s < S; ++s) {
    float m12[12];
    cmo_tf_to_matrix(tf[s].q, tf[s].origin, m12);
    // callbackX: pcl_ros::transformPointCloud(input, *cloud_ptr, transform)
    Cloud::Ptr cloud_ptr(new Cloud);
    transformPointCloud(ctx, raw[s], *cloud_ptr, tf[s]);
    std::vector<float> in = packed(raw[s]), tr(in.size());
    cmo_transform(in.data(), static_cast<int64_t>(N[s]), m12, raw[s].is_dense ? 1 : 0, tr.data());
    expect_cloud(*cloud_ptr, tr, N[s], "transformPointCloud");
    CHECK(cloud_ptr->is_dense == raw[s].is_dense, "transformPointCloud keeps is_dense");
    // getROI(cloud_ptr, cloud_ROI_ptr)
    Cloud::Ptr roi_ptr(new Cloud);
    getROI(ctx, cloud_ptr, roi_ptr);
    const std::vector<float> roi_o = oracle_passes(tr, roi);
    expect_cloud(*roi_ptr, roi_o, roi_o.size() / 4, "getROI");
    CHECK(roi_ptr->is_dense, "PassThrough output is dense");
    // getCloudPart(cloud_ROI_ptr, part, length, deviation)
    Cloud::Ptr part(new Cloud);
    getCloudPart(ctx, roi_ptr, part, 30.0f, -prm.roi_mid + 11.0f);
    const std::vector<float> part_o = oracle_passes(roi_o, {{0, -prm.roi_mid + 11.0f, -prm.roi_mid + 11.0f + 30.0f, 0}});
    expect_cloud(*part, part_o, part_o.size() / 4, "getCloudPart");
    CHECK(part->points.size() > 100 && part->points.size() < roi_ptr->points.size(), "getCloudPart is a proper slice");
    // proceedFront's zone loop: getCloudPart x5, each followed by the two z windows of removeGround (:228-270, :80-92)
    {
      const float rm = prm.roi_mid;
      const std::vector<ZonePart> parts = {{30.0f, -rm + 11.0f + 8.0f + 15.0f + 11.0f, 2.5f}, {11.0f, -rm + 11.0f + 8.0f + 15.0f, 2.0f},
                                           {15.0f, -rm + 11.0f + 8.0f, 1.5f}, {8.0f, -rm + 11.0f, 0.3f}, {11.0f, -rm, 0.5f}};
      std::vector<Cloud::Ptr> ground_parts, no_ground_parts;
      getCloudPartsZSplit(ctx, roi_ptr, parts, prm.roi_z_max, ground_parts, no_ground_parts);
      CHECK(ground_parts.size() == parts.size() && no_ground_parts.size() == parts.size(), "zone split: %s", ctx.last_error().c_str());
      size_t total = 0;
      for (size_t k = 0; k < parts.size() && k < ground_parts.size(); ++k) {
        const cmo_pass_t x = {0, parts[k].deviation, parts[k].deviation + parts[k].length, 0};
        const std::vector<float> g_o = oracle_passes(roi_o, {x, {2, -parts[k].z_max_ground, parts[k].z_max_ground, 0}});
        const std::vector<float> n_o = oracle_passes(roi_o, {x, {2, static_cast<float>(parts[k].z_max_ground + 0.01), prm.roi_z_max, 0}});
        expect_cloud(*ground_parts[k], g_o, g_o.size() / 4, "zone ground part");
        expect_cloud(*no_ground_parts[k], n_o, n_o.size() / 4, "zone no-ground part");
        total += g_o.size() / 4 + n_o.size() / 4;
      }
      CHECK(total > roi_ptr->points.size() / 2 && total <= roi_ptr->points.size() + 16, "zones cover most of the ROI cloud (%zu of %zu)",
            total, roi_ptr->points.size());
      // proceedX after getROI (:228-312): per zone getCloudPart + removeGround, appended zone after zone
      Cloud::Ptr ng_all(new Cloud), g_all(new Cloud);
      proceedZones(ctx, roi_ptr, parts, ng_all, g_all);
      std::vector<float> want_g, want_ng;
      for (size_t k = 0; k < parts.size(); ++k) {
        const cmo_pass_t x = {0, parts[k].deviation, parts[k].deviation + parts[k].length, 0};
        const std::vector<float> low = oracle_passes(roi_o, {x, {2, -parts[k].z_max_ground, parts[k].z_max_ground, 0}});
        const std::vector<float> high = oracle_passes(roi_o, {x, {2, static_cast<float>(parts[k].z_max_ground + 0.01), prm.roi_z_max, 0}});
        const int64_t n_low = static_cast<int64_t>(low.size() / 4);
        std::vector<int32_t> inl(static_cast<size_t>(n_low) + 1);
        const int64_t n_in = cmo_plane_ransac(low.data(), n_low, static_cast<double>(prm.distance_threshold), static_cast<double>(prm.prob),
                                              prm.max_iterations, 1, 12345u, prm.sum_order, nullptr, nullptr, nullptr, inl.data());
        std::vector<float> rest;
        for (int64_t i = 0, j = 0; i < n_low; ++i) {
          if (j < n_in && inl[static_cast<size_t>(j)] == i) { want_g.insert(want_g.end(), &low[i * 4], &low[i * 4] + 4); ++j; }
          else rest.insert(rest.end(), &low[i * 4], &low[i * 4] + 4);
        }
        const int64_t n_rest = static_cast<int64_t>(rest.size() / 4);
        std::vector<int32_t> keep(static_cast<size_t>(n_rest) + 1);
        const int64_t kk = cmo_radius_outlier(rest.data(), n_rest, static_cast<double>(prm.radius), static_cast<int32_t>(prm.min_neighbor), 0, keep.data());
        for (int64_t i = 0; i < kk; ++i) want_ng.insert(want_ng.end(), &rest[static_cast<size_t>(keep[i]) * 4], &rest[static_cast<size_t>(keep[i]) * 4] + 4);
        want_ng.insert(want_ng.end(), high.begin(), high.end());
      }
      expect_cloud(*g_all, want_g, want_g.size() / 4, "proceedZones ground");
      expect_cloud(*ng_all, want_ng, want_ng.size() / 4, "proceedZones no_ground");
    }
    // outlierRemoval(cloud_ptr) on a copy of the ROI cloud (in the reference it runs on the RANSAC outliers, :119)
    {
      Cloud::Ptr filtered(new Cloud(*roi_ptr));
      outlierRemoval(ctx, filtered);
      const int64_t n_roi = static_cast<int64_t>(roi_o.size() / 4);
      std::vector<int32_t> keep(static_cast<size_t>(n_roi) + 1);
      const int64_t k = cmo_radius_outlier(roi_o.data(), n_roi, static_cast<double>(prm.radius), static_cast<int32_t>(prm.min_neighbor), 0, keep.data());
      std::vector<float> want(static_cast<size_t>(k) * 4);
      for (int64_t i = 0; i < k; ++i) std::memcpy(&want[i * 4], &roi_o[static_cast<size_t>(keep[i]) * 4], 16);
      expect_cloud(*filtered, want, static_cast<size_t>(k), "outlierRemoval");
      CHECK(k > 100 && k < n_roi, "outlierRemoval removes some but not all points (%lld of %lld kept)", (long long)k, (long long)n_roi);
    }
    // removeGround(cloud, no_ground, ground, z_min, z_max, max_angle) -- :71-122 -- on the whole ROI cloud of this sensor
    {
      Cloud::Ptr no_ground(new Cloud), ground(new Cloud);
      const float zg = 0.4f;
      removeGround(ctx, roi_ptr, no_ground, ground, -zg, zg, 0.05f);
      const int64_t n_roi = static_cast<int64_t>(roi_o.size() / 4);
      std::vector<int32_t> idx(static_cast<size_t>(n_roi) + 1);
      auto gather = [&](const std::vector<float>& src, const int32_t* ix, int64_t k) {
        std::vector<float> o(static_cast<size_t>(k) * 4);
        for (int64_t i = 0; i < k; ++i) std::memcpy(&o[i * 4], &src[static_cast<size_t>(ix[i]) * 4], 16);
        return o;
      };
      int64_t k = cmo_passthrough(roi_o.data(), n_roi, 2, -zg, zg, 0, idx.data());
      const std::vector<float> low = gather(roi_o, idx.data(), k);
      k = cmo_passthrough(roi_o.data(), n_roi, 2, static_cast<float>(zg + 0.01), prm.roi_z_max, 0, idx.data());
      const std::vector<float> high = gather(roi_o, idx.data(), k);
      const int64_t n_low = static_cast<int64_t>(low.size() / 4);
      std::vector<int32_t> inl(static_cast<size_t>(n_low) + 1);
      int32_t info[7];
      const int64_t n_in = cmo_plane_ransac(low.data(), n_low, static_cast<double>(prm.distance_threshold), static_cast<double>(prm.prob),
                                            prm.max_iterations, 1, 12345u, prm.sum_order, info, nullptr, nullptr, inl.data());
      const std::vector<float> want_ground = gather(low, inl.data(), n_in);
      std::vector<int32_t> rest;
      for (int64_t i = 0, j = 0; i < n_low; ++i) {
        if (j < n_in && inl[static_cast<size_t>(j)] == i) { ++j; continue; }
        rest.push_back(static_cast<int32_t>(i));
      }
      const std::vector<float> not_ground = gather(low, rest.data(), static_cast<int64_t>(rest.size()));
      std::vector<int32_t> keep(rest.size() + 1);
      const int64_t kk = cmo_radius_outlier(not_ground.data(), static_cast<int64_t>(rest.size()), static_cast<double>(prm.radius),
                                            static_cast<int32_t>(prm.min_neighbor), 0, keep.data());
      std::vector<float> want_ng = gather(not_ground, keep.data(), kk);
      want_ng.insert(want_ng.end(), high.begin(), high.end());
      expect_cloud(*ground, want_ground, static_cast<size_t>(n_in), "removeGround ground");
      expect_cloud(*no_ground, want_ng, want_ng.size() / 4, "removeGround no_ground");
      CHECK(info[0] == 1 && n_in > 50 && n_in < n_low, "removeGround found a plane (%lld of %lld inliers)", (long long)n_in, (long long)n_low);
    }
    // fusePointclouds: *no_ground_ptr = first; *no_ground_ptr += rest
    if (s == 0) *fused = *roi_ptr; else *fused += *roi_ptr;
    fused_oracle.insert(fused_oracle.end(), roi_o.begin(), roi_o.end());
  }
  CHECK(stamp_of(*fused) == 100 + 7 * (S - 1), "operator+= keeps the newest stamp");
  // voxelgrid(no_ground_ptr, voxel_cloud_ptr)
  Cloud::Ptr voxel(new Cloud);
  voxelgrid(ctx, fused, voxel);
  {
    const int64_t m = static_cast<int64_t>(fused_oracle.size() / 4);
    const float leaf[3] = {prm.voxel_size, prm.voxel_size, prm.voxel_size};
    std::vector<float> cen(static_cast<size_t>(m) * 4 + 4);
    int32_t flags = 0;
    const int64_t v = cmo_voxelgrid(fused_oracle.data(), m, 1, leaf, prm.points_per_voxel, 1, 0, cen.data(), nullptr, nullptr,
                                    nullptr, nullptr, nullptr, nullptr, &flags);
    expect_cloud(*voxel, cen, static_cast<size_t>(v), "voxelgrid");
    CHECK(v > 50, "voxelgrid produced %lld voxels", static_cast<long long>(v));
  }

  // ---- the fused production path: callbacks submit, the main loop merges ----------------------------------------------------
  {
    FusedFrame ff(S, 1 << 16, /*required: sensors 0 and 1; sensor 2 optional like the top Velodyne*/ 0b011);
    CHECK(ff.ok(), "FusedFrame");
    for (int s = 0; s < S; ++s) ff.setTransform(s, tf[s]);
    Cloud f2, v2;
    ff.onCloud(0, raw[0]);
    CHECK(!ff.fuseAndVoxel(f2, v2), "fuse must wait for the required sensors (flag gate)");
    ff.onCloud(0, raw[0]);
    ff.onCloud(1, raw[1]);
    ff.onCloud(2, raw[2]);
    CHECK(ff.fuseAndVoxel(f2, v2), "fuseAndVoxel");
    std::vector<cmo_cloud_t> oc(S);
    std::vector<std::vector<float>> keep;
    for (int s = 0; s < S; ++s) {
      keep.push_back(packed(raw[s]));
      oc[s].data = reinterpret_cast<const uint8_t*>(keep.back().data());
      oc[s].n_points = static_cast<int64_t>(N[s]); oc[s].point_step = 16;
      oc[s].off_x = 0; oc[s].off_y = 4; oc[s].off_z = 8; oc[s].off_i = 12; oc[s].is_dense = raw[s].is_dense ? 1 : 0;
      cmo_tf_to_matrix(tf[s].q, tf[s].origin, oc[s].m);
    }
    const size_t tot = N[0] + N[1] + N[2];
    std::vector<float> sx(tot * 4), cen(tot * 4);
    std::vector<uint32_t> ssrc(tot);
    int64_t nsurv = 0;
    int32_t flags = 0;
    const float leaf[3] = {prm.voxel_size, prm.voxel_size, prm.voxel_size};
    const int64_t v = cmo_merge_frame(oc.data(), S, roi.data(), 3, leaf, prm.points_per_voxel, 1, 0, 1, sx.data(), ssrc.data(), &nsurv,
                                      cen.data(), nullptr, nullptr, nullptr, nullptr, nullptr, &flags);
    expect_cloud(f2, sx, static_cast<size_t>(nsurv), "FusedFrame fused cloud");
    expect_cloud(v2, cen, static_cast<size_t>(v), "FusedFrame voxel cloud");
    CHECK(stamp_of(v2) == 100 + 7 * (S - 1), "fused stamp");
    // the function-by-function result and the fused result are the same clouds
    CHECK(f2.points.size() == fused->points.size() && v2.points.size() == voxel->points.size(), "fused == stepwise");
  }
  std::printf("%s: %d failure(s)\n", g_fail ? "FAILED" : "shim ok", g_fail);
  return g_fail ? 1 : 0;
}
