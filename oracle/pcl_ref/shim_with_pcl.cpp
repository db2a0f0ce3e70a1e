// shim_with_pcl.cpp -- compiles include/cloud_merger_shim.hpp against REAL PCL (CM_SHIM_HAVE_PCL): proves that the
// reference-named functions take pcl::PointCloud<pcl::PointXYZI>::Ptr as the reference passes them and that the stand-in
// record layout of the tests (32 bytes, x y z @ 0 4 8, 1.0f @ 12, intensity @ 16) is PCL's own. Built only where PCL exists.
#include "cloud_merger_shim.hpp"

#ifndef CM_SHIM_HAVE_PCL
#error "this translation unit must see <pcl/point_cloud.h>"
#endif
#include <cstddef>

static_assert(sizeof(pcl::PointXYZI) == 32, "pcl::PointXYZI is a 32-byte record");
static_assert(offsetof(pcl::PointXYZI, x) == 0 && offsetof(pcl::PointXYZI, y) == 4 && offsetof(pcl::PointXYZI, z) == 8,
              "xyz at 0, 4, 8");
static_assert(offsetof(pcl::PointXYZI, intensity) == 16, "intensity at 16");

// the reference's call sites, written against the shim (pc_preprocessing_main.cpp:228-269, :131-177)
extern "C" int cm_shim_pcl_compiles(void) {
  using namespace cloud_merger;
  Context ctx(1024, 6);
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud_ptr(new pcl::PointCloud<pcl::PointXYZI>), roi(new pcl::PointCloud<pcl::PointXYZI>),
      no_ground(new pcl::PointCloud<pcl::PointXYZI>), ground(new pcl::PointCloud<pcl::PointXYZI>), voxel(new pcl::PointCloud<pcl::PointXYZI>);
  pcl::PointXYZI p;
  p.x = 1.f; p.y = 2.f; p.z = 0.5f; p.intensity = 7.f;
  cloud_ptr->push_back(p);
  if (p.data[3] != 1.0f) return 2;  // PCL_ADD_POINT4D pads xyz with 1.0f
  if (!ctx.ok()) return 77;         // no GPU here: the calls below would only report CM_E_NO_DEVICE
  getROI(ctx, cloud_ptr, roi);
  proceedFront(ctx, cloud_ptr, no_ground, ground);
  voxelgrid(ctx, no_ground, voxel);
  return 0;
}
