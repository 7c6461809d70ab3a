// SHIM (test infrastructure, ours): drake::Vector3<T> for the reference's bvh.cpp.  Drake is not installed here.
#pragma once
#include <Eigen/Dense>
namespace drake {
template <typename T>
using Vector3 = Eigen::ShimVector3<T>;
}
