// cm_pcl_ref.cpp -- C-ABI wrapper around the REAL PCL classes the reference calls, with the reference's exact settings.
// TEST INFRASTRUCTURE ONLY: it exists so that someone who has PCL (1.8.x, the ROS Melodic pin) can pin the CPU oracle
// (oracle/cm_oracle.cpp) -- and through it the CUDA path -- against the library the reference actually runs. PCL is not
// installed in the build image of this repository, so nothing here has been compiled there; tests/test_oracle_vs_pcl.py
// skips (and says so) when the library is absent.
//
// One function per PCL call site of the hot path, same argument meaning and output convention as the cmo_* function of
// oracle/cm_oracle.h it is diffed against:
//   cmp_transform        pcl::transformPointCloud(cloud_in, cloud_out, Eigen::Affine3f)  (what pcl_ros::transformPointCloud(in, out,
//                        tf::Transform) calls; pc_preprocessing_main.cpp:322, CloudFusionNode.h:508)
//   cmp_passthrough      pcl::PassThrough<PointXYZI>: setFilterFieldName / setFilterLimits / filter   (pc_preprocessing_main.cpp:20-59)
//   cmp_concat           pcl::PointCloud::operator+=                                               (:137-149)
//   cmp_voxelgrid        pcl::VoxelGrid<PointXYZI>: setLeafSize, setDownsampleAllData(true),
//                        setMinimumPointsNumberPerVoxel                                            (:168-177)
//   cmp_radius_outlier   pcl::RadiusOutlierRemoval<PointXYZI>: setRadiusSearch, setMinNeighborsInRadius,
//                        setKeepOrganized(false)                                                   (:184-192)
//   cmp_plane_ransac     pcl::SACSegmentation<PointXYZI>: SACMODEL_PLANE, SAC_RANSAC, setMaxIterations, setAxis, setEpsAngle,
//                        setDistanceThreshold, setOptimizeCoefficients(true), setProbability      (:95-108)
//   cmp_tf_to_matrix     Eigen::Quaternionf / Translation3f -> Affine3f as pcl_ros builds it from a tf::Transform
#include <pcl/common/transforms.h>
#include <pcl/filters/extract_indices.h>
#include <pcl/filters/passthrough.h>
#include <pcl/filters/radius_outlier_removal.h>
#include <pcl/filters/voxel_grid.h>
#include <pcl/pcl_config.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/segmentation/sac_segmentation.h>

#include <Eigen/Geometry>
#include <cstdint>
#include <cstdio>
#include <cstring>

using Cloud = pcl::PointCloud<pcl::PointXYZI>;

namespace {
Cloud::Ptr to_cloud(const float* xyzi, int64_t n, bool is_dense) {
  Cloud::Ptr c(new Cloud);
  c->points.resize(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) {
    pcl::PointXYZI& p = c->points[static_cast<size_t>(i)];
    p.x = xyzi[i * 4 + 0]; p.y = xyzi[i * 4 + 1]; p.z = xyzi[i * 4 + 2]; p.intensity = xyzi[i * 4 + 3];
  }
  c->width = static_cast<uint32_t>(n); c->height = 1; c->is_dense = is_dense;
  return c;
}
void from_cloud(const Cloud& c, float* xyzi) {
  for (size_t i = 0; i < c.points.size(); ++i) {
    xyzi[i * 4 + 0] = c.points[i].x; xyzi[i * 4 + 1] = c.points[i].y; xyzi[i * 4 + 2] = c.points[i].z;
    xyzi[i * 4 + 3] = c.points[i].intensity;
  }
}
const char* axis_name(int axis) { return axis == 0 ? "x" : axis == 1 ? "y" : axis == 2 ? "z" : "intensity"; }
}  // namespace

extern "C" {

const char* cmp_version(void) {
  static char buf[64];
  std::snprintf(buf, sizeof(buf), "PCL %d.%d.%d", PCL_MAJOR_VERSION, PCL_MINOR_VERSION, PCL_REVISION_VERSION);
  return buf;
}

void cmp_tf_to_matrix(const double* q_xyzw, const double* origin, float* m12) {
  // pcl_ros::transformPointCloud(in, out, tf::Transform): Eigen::Quaternionf(w, x, y, z), Eigen::Vector3f(origin),
  // Eigen::Affine3f t(Eigen::Translation3f(origin) * rotation)
  const Eigen::Quaternionf rot(static_cast<float>(q_xyzw[3]), static_cast<float>(q_xyzw[0]), static_cast<float>(q_xyzw[1]),
                               static_cast<float>(q_xyzw[2]));
  const Eigen::Vector3f org(static_cast<float>(origin[0]), static_cast<float>(origin[1]), static_cast<float>(origin[2]));
  const Eigen::Affine3f t(Eigen::Translation3f(org) * rot);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) m12[r * 4 + c] = t.matrix()(r, c);
}

void cmp_transform(const float* in_xyzi, int64_t n, const float* m12, int32_t is_dense, float* out_xyzi) {
  Eigen::Affine3f t = Eigen::Affine3f::Identity();
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) t.matrix()(r, c) = m12[r * 4 + c];
  Cloud::Ptr in = to_cloud(in_xyzi, n, is_dense != 0);
  Cloud out;
  pcl::transformPointCloud(*in, out, t);
  from_cloud(out, out_xyzi);
}

int64_t cmp_passthrough(const float* in_xyzi, int64_t n, int32_t axis, float lo, float hi, int32_t negative,
                        int32_t* out_indices) {
  Cloud::Ptr in = to_cloud(in_xyzi, n, true);
  pcl::PassThrough<pcl::PointXYZI> pass;
  pass.setInputCloud(in);
  pass.setFilterFieldName(axis_name(axis));
  pass.setFilterLimits(lo, hi);
  pass.setFilterLimitsNegative(negative != 0);
  std::vector<int> idx;
  pass.filter(idx);
  for (size_t i = 0; i < idx.size(); ++i) out_indices[i] = idx[i];
  return static_cast<int64_t>(idx.size());
}

// a += b; returns the size; writes width, height, is_dense of the result to meta[3]
int64_t cmp_concat(const float* a_xyzi, int64_t na, int32_t a_dense, uint64_t a_stamp, const float* b_xyzi, int64_t nb,
                   int32_t b_dense, uint64_t b_stamp, float* out_xyzi, uint64_t* out_stamp, int32_t* meta) {
  Cloud::Ptr a = to_cloud(a_xyzi, na, a_dense != 0), b = to_cloud(b_xyzi, nb, b_dense != 0);
  a->header.stamp = a_stamp; b->header.stamp = b_stamp;
  *a += *b;
  from_cloud(*a, out_xyzi);
  if (out_stamp) *out_stamp = a->header.stamp;
  if (meta) { meta[0] = static_cast<int32_t>(a->width); meta[1] = static_cast<int32_t>(a->height); meta[2] = a->is_dense ? 1 : 0; }
  return static_cast<int64_t>(a->points.size());
}

// returns the number of output points; grid[9] = getMinBoxCoordinates, getMaxBoxCoordinates, getNrDivisions
int64_t cmp_voxelgrid(const float* xyzi, int64_t n, int32_t is_dense, const float* leaf, uint32_t min_points,
                      int32_t downsample_all, float* out_xyzi, int32_t* grid) {
  Cloud::Ptr in = to_cloud(xyzi, n, is_dense != 0);
  pcl::VoxelGrid<pcl::PointXYZI> vg;
  vg.setInputCloud(in);
  vg.setLeafSize(leaf[0], leaf[1], leaf[2]);
  vg.setDownsampleAllData(downsample_all != 0);
  vg.setMinimumPointsNumberPerVoxel(min_points);
  Cloud out;
  vg.filter(out);
  from_cloud(out, out_xyzi);
  if (grid) {
    const Eigen::Vector3i mn = vg.getMinBoxCoordinates(), mx = vg.getMaxBoxCoordinates(), dv = vg.getNrDivisions();
    for (int k = 0; k < 3; ++k) { grid[k] = mn[k]; grid[3 + k] = mx[k]; grid[6 + k] = dv[k]; }
  }
  return static_cast<int64_t>(out.points.size());
}

int64_t cmp_radius_outlier(const float* xyzi, int64_t n, double radius, int32_t min_pts, int32_t negative, int32_t* out_indices) {
  Cloud::Ptr in = to_cloud(xyzi, n, true);
  pcl::RadiusOutlierRemoval<pcl::PointXYZI> outrem;
  outrem.setInputCloud(in);
  outrem.setRadiusSearch(radius);
  outrem.setMinNeighborsInRadius(min_pts);
  outrem.setKeepOrganized(false);
  outrem.setNegative(negative != 0);
  std::vector<int> idx;
  outrem.filter(idx);
  for (size_t i = 0; i < idx.size(); ++i) out_indices[i] = idx[i];
  return static_cast<int64_t>(idx.size());
}

// coeff_out[4]; out_inliers ascending; returns the number of inliers (0: segment() failed)
int64_t cmp_plane_ransac(const float* xyzi, int64_t n, double threshold, double probability, int32_t max_iterations,
                         int32_t optimize, float eps_angle, float* coeff_out, int32_t* out_inliers) {
  Cloud::Ptr in = to_cloud(xyzi, n, true);
  pcl::SACSegmentation<pcl::PointXYZI> seg;
  pcl::PointIndices::Ptr inliers(new pcl::PointIndices);
  pcl::ModelCoefficients::Ptr coefficients(new pcl::ModelCoefficients);
  seg.setModelType(pcl::SACMODEL_PLANE);
  seg.setMethodType(pcl::SAC_RANSAC);
  seg.setMaxIterations(max_iterations);
  seg.setAxis(Eigen::Vector3f(0.0f, 0.0f, 1.0f));
  seg.setEpsAngle(eps_angle);
  seg.setDistanceThreshold(threshold);
  seg.setOptimizeCoefficients(optimize != 0);
  seg.setProbability(probability);
  seg.setInputCloud(in);
  seg.segment(*inliers, *coefficients);
  for (size_t i = 0; i < inliers->indices.size(); ++i) out_inliers[i] = inliers->indices[i];
  if (coeff_out)
    for (size_t k = 0; k < 4; ++k) coeff_out[k] = k < coefficients->values.size() ? coefficients->values[k] : 0.0f;
  return static_cast<int64_t>(inliers->indices.size());
}

}  // extern "C"
