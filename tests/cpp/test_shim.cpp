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
#include <thread>
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

// oracle: what a proceedX does after getROI (pc_preprocessing_main.cpp:228-312) -- per zone getCloudPart, the two z windows,
// RANSAC plane on the lower one, outlierRemoval of its non-ground points, the upper window appended; zone after zone
static void oracle_proceed(const std::vector<float>& roi_o, const std::vector<ZonePart>& parts, const Params& prm,
                           std::vector<float>& want_g, std::vector<float>& want_ng) {
  want_g.clear(); want_ng.clear();
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
      oracle_proceed(roi_o, parts, prm, want_g, want_ng);
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
    // the four zone tables of the reference (proceedFront, proceedRear, callbackTopMiddle, callbackFrontMiddle; :228-312, :428-446,
    // :474-497 with Parameter.h:45-81) on the TRANSFORMED cloud: getROI inside, zones, plain parts appended
    {
      const ProceedTable tables[4] = {frontTable(prm), rearTable(prm), topTable(prm), livoxTable(prm)};
      const char* names[4] = {"proceedFront", "proceedRear", "proceedTop", "proceedLivox"};
      CHECK(tables[0].parts.size() == 5 && tables[1].parts.size() == 4 && tables[2].parts.size() == 1 && tables[2].plain.size() == 1 &&
            tables[3].parts.size() == 4, "zone tables");
      // deviations as the reference writes them: -roi_mid + vr_rear + vr_veh + vr_mid = 30, vt: 20, plain part [-15, 20], livox front 34
      CHECK(tables[1].parts[0].deviation == 30.0f && tables[1].parts[0].length == 30.0f && tables[2].parts[0].deviation == 20.0f &&
            tables[2].plain[0].deviation == -15.0f && tables[2].plain[0].length == 35.0f && tables[3].parts[0].deviation == 34.0f,
            "zone table values");
      for (int t = 0; t < 4; ++t) {
        if (s != (t % S)) continue;  // one table per sensor cloud keeps the test short
        Cloud::Ptr ng(new Cloud), g(new Cloud);
        proceedTable(ctx, cloud_ptr, tables[t], ng, g);
        std::vector<float> want_g, want_ng;
        oracle_proceed(roi_o, tables[t].parts, prm, want_g, want_ng);
        for (const ZonePart& pl : tables[t].plain) {
          const std::vector<float> part_plain = oracle_passes(roi_o, {{0, pl.deviation, pl.deviation + pl.length, 0}});
          want_ng.insert(want_ng.end(), part_plain.begin(), part_plain.end());
        }
        expect_cloud(*g, want_g, want_g.size() / 4, names[t]);
        expect_cloud(*ng, want_ng, want_ng.size() / 4, names[t]);
        CHECK(want_ng.size() > 400, "%s: the table selects points (%zu)", names[t], want_ng.size() / 4);
      }
    }
    // my_cloud_fusion's variant (CloudFusionNode.h:87-216, 327-493; my_cloud_fusion/src/Parameter.h): double constants
    {
      const FusionParams fp;
      const FusionParams::Group& gr = fp.group[s % 4];
      auto f = [](double v) { return static_cast<float>(v); };
      const cmo_pass_t z = {2, f(fp.z_min), f(fp.z_max), 0}, y = {1, f(-fp.lane_width / 2), f(fp.lane_width / 2), 0};
      const cmo_pass_t xr[3] = {{0, f(gr.mid_length / 2), f((gr.mid_length / 2) + gr.front_length), 0},
                                {0, f(-gr.mid_length / 2), f(gr.mid_length / 2), 0},
                                {0, f(-(gr.mid_length / 2) - gr.rear_length), f(gr.rear_length), 0}};
      Cloud::Ptr fr(new Cloud), mi(new Cloud), re(new Cloud);
      filter_ROI_R(ctx, fp, cloud_ptr, fr, mi, re, gr.front_length, gr.mid_length, gr.rear_length);
      const Cloud::Ptr got3[3] = {fr, mi, re};
      std::vector<float> zone_o[3];
      for (int k = 0; k < 3; ++k) {
        zone_o[k] = oracle_passes(tr, {z, y, xr[k]});
        expect_cloud(*got3[k], zone_o[k], zone_o[k].size() / 4, "filter_ROI_R");
      }
      CHECK(zone_o[0].size() + zone_o[1].size() > 400, "filter_ROI_R selects points");
      // remove_ground: the two z windows (RANSAC is commented out in this package)
      Cloud::Ptr ng(new Cloud), g(new Cloud);
      remove_ground(ctx, fp, mi, ng, g, gr.z_min_ground_mid, gr.z_max_ground_mid, 0.01);
      const std::vector<float> g_o = oracle_passes(zone_o[1], {{2, f(gr.z_min_ground_mid), f(gr.z_max_ground_mid), 0}});
      const std::vector<float> ng_o = oracle_passes(zone_o[1], {{2, f(gr.z_max_ground_mid + 0.01), f(fp.z_max), 0}});
      expect_cloud(*g, g_o, g_o.size() / 4, "remove_ground ground");
      expect_cloud(*ng, ng_o, ng_o.size() / 4, "remove_ground no_ground");
      // proceed_pointcloud: all of it in one zone-slicing pass
      Cloud::Ptr png(new Cloud), pg(new Cloud);
      proceed_pointcloud(ctx, fp, cloud_ptr, png, pg, s % 4);
      std::vector<float> want_g, want_ng;
      const double zlo[3] = {gr.z_min_ground_front, gr.z_min_ground_mid, gr.z_min_ground_rear};
      const double zhi[3] = {gr.z_max_ground_front, gr.z_max_ground_mid, gr.z_max_ground_rear};
      for (int k = 0; k < 3; ++k) {
        const std::vector<float> a = oracle_passes(zone_o[k], {{2, f(zlo[k]), f(zhi[k]), 0}});
        const std::vector<float> b = oracle_passes(zone_o[k], {{2, f(zhi[k] + 0.01), f(fp.z_max), 0}});
        want_g.insert(want_g.end(), a.begin(), a.end());
        want_ng.insert(want_ng.end(), b.begin(), b.end());
      }
      expect_cloud(*pg, want_g, want_g.size() / 4, "proceed_pointcloud ground");
      expect_cloud(*png, want_ng, want_ng.size() / 4, "proceed_pointcloud no_ground");
      // filter_ROI_T: longitudinal bar += transverse bar
      Cloud::Ptr tshape(new Cloud);
      filter_ROI_T(ctx, fp, cloud_ptr, tshape);
      std::vector<float> t_o = oracle_passes(tr, {{0, f(-fp.x_longitudinal / 2), f(fp.x_longitudinal / 2), 0},
                                                  {1, f(-fp.y_longitudinal / 2), f(fp.y_longitudinal / 2), 0}, z});
      const std::vector<float> t2 = oracle_passes(tr, {{0, f(fp.x_longitudinal / 2), f((fp.x_longitudinal / 2) + fp.x_transverse), 0},
                                                       {1, f(-fp.y_transverse / 2), f(fp.y_transverse / 2), 0}, z});
      t_o.insert(t_o.end(), t2.begin(), t2.end());
      expect_cloud(*tshape, t_o, t_o.size() / 4, "filter_ROI_T");
      // remove_outliers: radius 0.1 m in this package
      Cloud::Ptr ro(new Cloud(*mi));
      remove_outliers(ctx, fp, ro);
      const int64_t n_mi = static_cast<int64_t>(zone_o[1].size() / 4);
      std::vector<int32_t> keep(static_cast<size_t>(n_mi) + 1);
      const int64_t kk = cmo_radius_outlier(zone_o[1].data(), n_mi, fp.radius, static_cast<int32_t>(fp.min_neighbor), 0, keep.data());
      std::vector<float> want(static_cast<size_t>(kk) * 4);
      for (int64_t i = 0; i < kk; ++i) std::memcpy(&want[i * 4], &zone_o[1][static_cast<size_t>(keep[i]) * 4], 16);
      expect_cloud(*ro, want, static_cast<size_t>(kk), "remove_outliers");
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

    // ---- wire adapters: the same frame delivered as sensor_msgs::PointCloud2 in the sensors' own layouts, extrinsics as
    // Eigen::Matrix4f; published clouds described the way pcl::toROSMsg does (pc_preprocessing_main.cpp:199-220, :520-525)
#if defined(CM_SHIM_HAVE_ROS_MSG) && defined(CM_SHIM_HAVE_EIGEN)
    {
      struct Lay { uint32_t step, ox, oy, oz, oi; const char* extra; uint32_t oextra; uint8_t textra; };
      const Lay lays[3] = {{32, 0, 4, 8, 16, "ring", 20, sensor_msgs::PointField::UINT16},     // Velodyne (Melodic driver)
                           {22, 0, 4, 8, 12, "ring", 16, sensor_msgs::PointField::UINT16},     // newer Velodyne: + time f32 @ 18
                           {18, 0, 4, 8, 12, "tag", 16, sensor_msgs::PointField::UINT8}};      // Livox
      FusedFrame fm(S, 1 << 16, 0b111);
      CHECK(fm.ok(), "FusedFrame (msg)");
      for (int s = 0; s < S; ++s) {
        float m12[12];
        cmo_tf_to_matrix(tf[s].q, tf[s].origin, m12);
        Eigen::Matrix4f m = Eigen::Matrix4f::Identity();
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) m(r, c) = m12[r * 4 + c];
        fm.setTransform(s, m);
        sensor_msgs::PointCloud2 msg;
        const Lay& L = lays[s];
        msg.height = 1; msg.width = static_cast<uint32_t>(N[s]); msg.point_step = L.step; msg.row_step = L.step * msg.width;
        msg.is_dense = raw[s].is_dense; msg.is_bigendian = 0;
        msg.header.stamp.sec = 100 + 7 * s; msg.header.stamp.nsec = 0;
        const char* nm[4] = {"x", "y", "z", "intensity"};
        const uint32_t off[4] = {L.ox, L.oy, L.oz, L.oi};
        for (int k = 0; k < 4; ++k) {
          sensor_msgs::PointField f; f.name = nm[k]; f.offset = off[k]; f.datatype = sensor_msgs::PointField::FLOAT32; f.count = 1;
          msg.fields.push_back(f);
        }
        sensor_msgs::PointField fx; fx.name = L.extra; fx.offset = L.oextra; fx.datatype = L.textra; fx.count = 1;
        msg.fields.insert(msg.fields.begin() + 1, fx);  // field order in the message does not matter: lookup is by name
        msg.data.assign(static_cast<size_t>(msg.row_step), 0xA5);
        for (size_t i = 0; i < N[s]; ++i) {
          uint8_t* r = &msg.data[i * L.step];
          std::memcpy(r + L.ox, &raw[s].points[i].x, 4); std::memcpy(r + L.oy, &raw[s].points[i].y, 4);
          std::memcpy(r + L.oz, &raw[s].points[i].z, 4); std::memcpy(r + L.oi, &raw[s].points[i].intensity, 4);
        }
        cm_layout_t got;
        CHECK(layoutFromMsg(msg, &got) == CM_OK && got.point_step == (int)L.step && got.off_intensity == (int)L.oi &&
              got.is_dense == (raw[s].is_dense ? 1 : 0), "layoutFromMsg sensor %d", s);
        CHECK(fm.onCloudMsg(s, msg), "onCloudMsg sensor %d", s);
      }
      Cloud f3, v3;
      CHECK(fm.fuseAndVoxel(f3, v3), "fuseAndVoxel (msg)");
      expect_cloud(f3, sx, static_cast<size_t>(nsurv), "PointCloud2 path fused cloud");
      expect_cloud(v3, cen, static_cast<size_t>(v), "PointCloud2 path voxel cloud");
      CHECK(stamp_of(v3) == (100ull + 7 * (S - 1)) * 1000000ull, "stamp converted like pcl_conversions (microseconds)");
      // publish side: header + memcpy
      sensor_msgs::PointCloud2 out_msg;
      CHECK(toROSMsg(v3, out_msg), "toROSMsg");
      CHECK(out_msg.height == 1 && out_msg.width == v3.points.size() && out_msg.point_step == 32 && out_msg.row_step == 32 * out_msg.width &&
            !out_msg.is_bigendian && out_msg.is_dense && out_msg.fields.size() == 4 && out_msg.fields[3].name == "intensity" &&
            out_msg.fields[3].offset == 16 && out_msg.fields[3].datatype == sensor_msgs::PointField::FLOAT32 && out_msg.fields[2].offset == 8,
            "published message header as pcl::toROSMsg writes it");
      CHECK(out_msg.data.size() == v3.points.size() * 32 && (v3.points.empty() || std::memcmp(out_msg.data.data(), v3.points.data(), out_msg.data.size()) == 0),
            "published message data");
      // a message pcl_ros could not map (x as FLOAT64, or big-endian) is refused; an integer intensity reads as absent
      sensor_msgs::PointCloud2 bad;
      bad.point_step = 32;
      sensor_msgs::PointField bx; bx.name = "x"; bx.offset = 0; bx.datatype = sensor_msgs::PointField::FLOAT64; bx.count = 1;
      bad.fields.push_back(bx);
      cm_layout_t lb;
      CHECK(layoutFromMsg(bad, &lb) == CM_E_INVALID, "FLOAT64 x must be refused");
    }
#else
    CHECK(false, "the wire adapters were not compiled (mock ROS / Eigen headers missing from the include path)");
#endif
  }
  // ---- one whole main-loop iteration of pcl_preprocessing, device-resident, callbacks on their own threads ---------------------
  {
    const std::vector<ProceedTable> tables = {frontTable(prm), livoxTable(prm), topTable(prm)};
    PreprocessingFrame pf(tables, 1 << 16, /*required*/ 0b011);
    CHECK(pf.ok(), "PreprocessingFrame");
    for (int s = 0; s < S; ++s) pf.setTransform(s, tf[s]);
    Cloud ng, g, vx;
    CHECK(!pf.fuseAndVoxel(ng, g, vx), "fuse must wait for the required sensors");
    auto oracle_node = [&](const std::vector<const Cloud*>& in, std::vector<float>& want_ng, std::vector<float>& want_g) {
      want_ng.clear(); want_g.clear();
      for (int s = 0; s < S; ++s) {
        if (!in[static_cast<size_t>(s)]) continue;
        float m12[12];
        cmo_tf_to_matrix(tf[s].q, tf[s].origin, m12);
        const Cloud& c = *in[static_cast<size_t>(s)];
        std::vector<float> pk = packed(c), tr(pk.size());
        cmo_transform(pk.data(), static_cast<int64_t>(c.points.size()), m12, c.is_dense ? 1 : 0, tr.data());
        const std::vector<float> roi_o = oracle_passes(tr, roi);
        std::vector<float> a, b;
        oracle_proceed(roi_o, tables[static_cast<size_t>(s)].parts, prm, a, b);
        for (const ZonePart& pl : tables[static_cast<size_t>(s)].plain) {
          const std::vector<float> part_plain = oracle_passes(roi_o, {{0, pl.deviation, pl.deviation + pl.length, 0}});
          b.insert(b.end(), part_plain.begin(), part_plain.end());
        }
        want_g.insert(want_g.end(), a.begin(), a.end());
        want_ng.insert(want_ng.end(), b.begin(), b.end());
      }
    };
    auto check_frame = [&](const std::vector<const Cloud*>& in, const char* what) {
      std::vector<float> want_ng, want_g;
      oracle_node(in, want_ng, want_g);
      expect_cloud(ng, want_ng, want_ng.size() / 4, what);
      expect_cloud(g, want_g, want_g.size() / 4, what);
      const int64_t m = static_cast<int64_t>(want_ng.size() / 4);
      const float leaf[3] = {prm.voxel_size, prm.voxel_size, prm.voxel_size};
      std::vector<float> cen(static_cast<size_t>(m) * 4 + 4);
      int32_t flags = 0;
      const int64_t v = cmo_voxelgrid(want_ng.data(), m, 1, leaf, prm.points_per_voxel, 1, 0, cen.data(), nullptr, nullptr, nullptr,
                                      nullptr, nullptr, nullptr, &flags);
      expect_cloud(vx, cen, static_cast<size_t>(v), what);
      CHECK(v > 0 && want_g.size() > 400 && want_ng.size() > 400, "%s: non-trivial frame (%lld voxels, %zu ground, %zu no-ground points)", what,
            (long long)v, want_g.size() / 4, want_ng.size() / 4);
    };
    // frame 1: the two required sensors only (the optional one has never delivered: its stored clouds are empty)
    {
      std::thread t0([&] { pf.onCloud(0, raw[0]); }), t1([&] { pf.onCloud(1, raw[1]); });
      t0.join(); t1.join();
    }
    CHECK(pf.fuseAndVoxel(ng, g, vx), "PreprocessingFrame frame 1");
    check_frame({&raw[0], &raw[1], nullptr}, "node frame 1");
    // frame 2: all three from their own threads; sensor 0 delivers twice -- the first cloud after the fusion wins (:330)
    Cloud other = make_cloud(30000, false, 500);
    {
      pf.onCloud(0, raw[0]);
      std::thread t0([&] { pf.onCloud(0, other); }), t1([&] { pf.onCloud(1, raw[1]); }), t2([&] { pf.onCloud(2, raw[2]); });
      t0.join(); t1.join(); t2.join();
    }
    CHECK(pf.fuseAndVoxel(ng, g, vx), "PreprocessingFrame frame 2");
    check_frame({&raw[0], &raw[1], &raw[2]}, "node frame 2");
    // frame 3: the optional sensor stays silent -- its STORED clouds are appended again, as the reference's globals are
    pf.onCloud(0, other);
    pf.onCloud(1, raw[1]);
    CHECK(pf.fuseAndVoxel(ng, g, vx), "PreprocessingFrame frame 3");
    check_frame({&other, &raw[1], &raw[2]}, "node frame 3 (stale optional sensor)");
    CHECK(stamp_of(vx) == 500, "newest stamp");
  }
  std::printf("%s: %d failure(s)\n", g_fail ? "FAILED" : "shim ok", g_fail);
  return g_fail ? 1 : 0;
}
