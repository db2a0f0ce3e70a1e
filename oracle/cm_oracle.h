/*
 * cm_oracle.h -- CPU oracle for the cloud_merger merge hot path (TEST INFRASTRUCTURE ONLY).
 *
 * PARITY UNPINNED: the reference (timspilak/cloud_merger) ships no tests, fixtures or golden
 * vectors, and the arithmetic of the path lives in third-party libraries that are not under
 * /root/reference and are not installed in this image: PCL 1.8.1 (pcl::PassThrough,
 * pcl::VoxelGrid, pcl::PointCloud::operator+=, pcl::transformPointCloud), pcl_ros 1.7.x
 * (pcl_ros::transformPointCloud) and Eigen 3.3.4 -- the versions pinned by ROS Melodic /
 * Ubuntu 18.04 (reference: my_cloud_fusion/package.xml:55, my_cloud_fusion/CMakeLists.txt:18).
 * This file restates the published PCL 1.8.1 algorithms; it is anchored on the reference's
 * call sites and on the hand-checkable known-answer vector of SURVEY.md section 8c
 * (tests/golden/known_answer.json), and cross-checked against an independent numpy
 * restatement (oracle/np_oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The product path (cloud_merger_b200/) never does.
 *
 * Every float operation below is an individually rounded IEEE-754 binary32 operation
 * (compile with -ffp-contract=off, no -ffast-math, no -march=native): the reference builds with
 * "-std=c++14" only (my_cloud_fusion/CMakeLists.txt:5), i.e. baseline x86-64 without FMA.
 */
#ifndef CM_ORACLE_H_
#define CM_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One PassThrough stage: pcl::PassThrough<PointXYZI> with setFilterFieldName / setFilterLimits /
 * setFilterLimitsNegative. axis: 0 = "x", 1 = "y", 2 = "z", 3 = "intensity". */
typedef struct {
  int32_t axis;
  float lo;
  float hi;
  int32_t negative;
} cmo_pass_t;

/* Description of one sensor cloud as it arrives in a sensor_msgs/PointCloud2 (data[], point_step and the
 * byte offsets of the FLOAT32 fields x, y, z, intensity). off_i < 0: no intensity field (read as 0). */
typedef struct {
  const uint8_t* data;
  int64_t n_points;
  int32_t point_step;
  int32_t off_x, off_y, off_z, off_i;
  int32_t is_dense;
  float m[12]; /* sensor -> base extrinsic, rows 0..2 of the 4x4, row-major: m[r*4+c] */
} cmo_cloud_t;

/* Result flags of cmo_voxelgrid */
#define CMO_FLAG_PCL_OVERFLOW 1 /* PCL 1.8.1 would warn "Leaf size is too small" and return the input unchanged */

/* a1: pcl_ros subscriber deserialisation (pc_preprocessing_main.cpp:520-525): PointCloud2 bytes -> packed xyzi. */
void cmo_unpack(const uint8_t* data, int64_t n, int32_t point_step, int32_t off_x, int32_t off_y, int32_t off_z,
                int32_t off_i, float* out_xyzi);

/* a2: pcl_ros::transformPointCloud (pc_preprocessing_main.cpp:322). In place allowed. m = 12 floats row-major 3x4. */
void cmo_transform(const float* in_xyzi, int64_t n, const float* m, int32_t is_dense, float* out_xyzi);

/* a3-a6: one pcl::PassThrough (pc_preprocessing_main.cpp:20-59). Writes the indices (into the input) of the kept
 * points in input order; returns how many. */
int64_t cmo_passthrough(const float* in_xyzi, int64_t n, int32_t axis, float lo, float hi, int32_t negative,
                        int32_t* out_indices);

/* a8: pcl::VoxelGrid<PointXYZI>::applyFilter (pc_preprocessing_main.cpp:168-177).
 *   out_xyzi      [n*4]  centroid, float accumulation in ascending point-index order (CentroidPoint order is
 *                        unspecified in PCL because std::sort is unstable; ascending index is one valid order)
 *   out_xyzi_sort [n*4]  centroid, float accumulation in the order std::sort left the pairs (what a PCL run does)
 *   out_xyzi_f64  [n*4]  centroid accumulated in double, ascending point-index order (tolerance anchor)
 *   out_count     [n]    points per emitted voxel
 *   out_idx       [n]    voxel index of each emitted voxel (PCL's idx; 64-bit so it also covers the extension)
 *   point_idx     [n]    voxel index of every input point (-1: point skipped as non-finite)
 *   grid          [9]    min_b[3], max_b[3], div_b[3]
 *   flags                CMO_FLAG_*
 * Any out pointer may be NULL. force64: 0 = behave exactly like PCL 1.8.1 (on overflow: copy the input to out_xyzi,
 * return n, set CMO_FLAG_PCL_OVERFLOW); 1 = the 64-bit extension (same formulas in int64, never refuses).
 * Returns the number of voxels emitted. */
int64_t cmo_voxelgrid(const float* xyzi, int64_t n, int32_t is_dense, const float* leaf, uint32_t min_points,
                      int32_t downsample_all, int32_t force64, float* out_xyzi, float* out_xyzi_sort,
                      double* out_xyzi_f64, uint32_t* out_count, int64_t* out_idx, int64_t* point_idx, int32_t* grid,
                      int32_t* flags);

/* The whole per-frame path exactly in the reference's order (callbackX -> getROI chain -> fusePointclouds -> voxelgrid):
 * per sensor unpack + transform + chained PassThrough (each pass copies its survivors, as copyPointCloud does),
 * concat with operator+=, VoxelGrid. threads > 1 runs the per-sensor part on that many std::threads (the reference
 * runs its six callbacks on ros::AsyncSpinner(6), pc_preprocessing_main.cpp:513); concat + VoxelGrid stay on one.
 *   out_survivor_xyzi [sum n * 4], out_survivor_src [sum n]: the merged cropped cloud and, per point, its index in the
 *   un-cropped concatenation. The voxel outputs are as in cmo_voxelgrid (out_xyzi = ascending-index float order).
 * Returns voxels emitted; *n_survivors receives the merged cloud size. */
int64_t cmo_merge_frame(const cmo_cloud_t* clouds, int32_t n_clouds, const cmo_pass_t* passes, int32_t n_passes,
                        const float* leaf, uint32_t min_points, int32_t downsample_all, int32_t force64,
                        int32_t threads, float* out_survivor_xyzi, uint32_t* out_survivor_src, int64_t* n_survivors,
                        float* out_xyzi, double* out_xyzi_f64, uint32_t* out_count, int64_t* out_idx,
                        int64_t* point_idx, int32_t* grid, int32_t* flags);

/* pcl::RadiusOutlierRemoval<PointXYZI> (pc_preprocessing_main.cpp:184-192, called from removeGround :119): keeps a
 * point iff more than min_pts points (itself included) lie strictly inside `radius` (float squared distance, FLANN
 * L2_Simple order); negative inverts. Writes the kept indices in input order, returns how many. */
int64_t cmo_radius_outlier(const float* xyzi, int64_t n, double radius, int32_t min_pts, int32_t negative,
                           int32_t* out_indices);

/* RANSAC ground plane: pcl::SACSegmentation<PointXYZI>, SACMODEL_PLANE + SAC_RANSAC, as removeGround() configures it
 * (pc_preprocessing_main.cpp:95-108; Parameter.h:38-42), restated from PCL 1.8.1 / Eigen 3.3.4 / Boost.Random (see the
 * block comment in cm_oracle.cpp). sum_order: order of Eigen's 4-wide float reductions, 0 = SSE2, 1 = SSE3, 2 = scalar.
 *  cmo_mt19937_at     i-th output of the sampler's engine (known-answer hook);
 *  cmo_plane_score    one externally given sample: coefficients + inlier count (-1: rejected by isSampleGood);
 *  cmo_plane_ransac   the whole segment() call; out_info[7] = found, iterations, draws, best count, best sample[3];
 *                     out_inliers = the indices pcl::ExtractIndices is then given (ascending). Returns their number. */
uint32_t cmo_mt19937_at(uint32_t seed, int64_t i);
int64_t cmo_plane_score(const float* xyzi, int64_t n, const int32_t* sample, double threshold, int32_t sum_order,
                        float* coeff);
int64_t cmo_plane_ransac(const float* xyzi, int64_t n, double threshold, double probability, int32_t max_iterations,
                         int32_t optimize, uint32_t seed, int32_t sum_order, int32_t* out_info, float* coeff_ransac,
                         float* coeff_out, int32_t* out_inliers);

/* Eigen 3.3.4 Quaternionf::toRotationMatrix + Translation, as pcl_ros builds the Affine3f from a tf::Transform
 * (double quaternion xyzw + origin narrowed to float). Writes 12 floats row-major. Host glue, kept here so the tests can
 * pin it too. */
void cmo_tf_to_matrix(const double* quat_xyzw, const double* origin_xyz, float* m);

const char* cmo_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CM_ORACLE_H_ */
