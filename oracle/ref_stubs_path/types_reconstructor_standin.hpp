/* Stand-in for external/manifoldReconstructor/include/types_reconstructor.hpp (TEST INFRASTRUCTURE): that vendored header pulls
 * CGAL's Delaunay triangulation and Eigen for types the hot path never touches.  Force-included with its include guard
 * pre-defined (-DTYPES_RECONSTR_HPP_), so that the reference's own SfMData.h is used unmodified.  CameraType as declared at
 * types_reconstructor.hpp:68-82. */
#pragma once
#include <string>
#include <vector>
#include <set>
#include <glm.hpp>
struct CameraType {
  glm::mat3 intrinsics;
  glm::mat3 rotation;
  glm::vec3 translation;
  glm::mat4 cameraMatrix;
  glm::vec3 center;
  glm::mat4 mvp;
  std::string pathImage;
  int imageWidth;
  int imageHeight;
  std::vector<int> visiblePoints;
};
