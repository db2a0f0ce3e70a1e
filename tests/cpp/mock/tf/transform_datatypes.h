// Test stand-in for <tf/transform_datatypes.h>: tf::Quaternion / tf::Vector3 / tf::Transform with the accessors the shim uses.
#pragma once
namespace tf {
struct Quaternion {
  double v[4] = {0, 0, 0, 1};
  Quaternion() = default;
  Quaternion(double x, double y, double z, double w) : v{x, y, z, w} {}
  double x() const { return v[0]; } double y() const { return v[1]; } double z() const { return v[2]; } double w() const { return v[3]; }
};
struct Vector3 {
  double v[3] = {0, 0, 0};
  Vector3() = default;
  Vector3(double x, double y, double z) : v{x, y, z} {}
  double x() const { return v[0]; } double y() const { return v[1]; } double z() const { return v[2]; }
};
struct Transform {
  Quaternion q; Vector3 o;
  Transform() = default;
  Transform(const Quaternion& q_, const Vector3& o_) : q(q_), o(o_) {}
  Quaternion getRotation() const { return q; }
  const Vector3& getOrigin() const { return o; }
};
}  // namespace tf
