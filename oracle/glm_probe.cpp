/*
 * glm_probe.cpp — the reference's camera-matrix arithmetic evaluated with the reference's OWN vendored glm (test infrastructure).
 *
 * Built by `make -C oracle ref` into oracle/_ref/libglm_probe.so with -I/root/reference/external/glm (header-only glm 0.9.6.3,
 * compiled where it lies; nothing of it is copied).  The few lines below repeat, with real glm types, what
 * external/manifoldReconstructor/src/OpenMvgParser.cpp does with the numbers of an OpenMVG sfm_data JSON:
 *   :252-256 intrinsics (focal, principal point -> glm::mat3), :280-289 rotation[r][c] = json[r][c], translation = -center * rotation,
 *   :107-125 eMatrix / kMatrix and cameraMatrix = eMatrix * kMatrix,
 * and src/edgegraph3d/utils/geometry/geometric_utilities.cpp:973-977 (compute_projection: vec4 * mat4, divide by the third
 * component).  tests/golden/make_golden_glm.py turns its answers into committed golden vectors that pin
 * edgegraph3d_b200/openmvg_io.py (camera construction) and the oracle's / the kernels' projection.
 */
#include <glm.hpp>

extern "C" {

/* rot9: json rotation row-major; out12: rows 0..2 of cameraMatrix read as [row][col] (convert_glm_mat4_to_cv_Mat34,
 * src/edgegraph3d/utils/edge_graph_3d_utilities.cpp:190-206); out_t3: translation */
void eg3d_ref_glm_camera(const float* rot9, const float* center3, float focal, float ppx, float ppy, float* out12, float* out_t3) {
  glm::mat3 intr(0.0);
  intr[0][0] = focal; intr[1][1] = focal; intr[0][2] = ppx; intr[1][2] = ppy; intr[2][2] = 1.0;
  glm::mat3 rotation; glm::vec3 center;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) rotation[r][c] = rot9[3 * r + c];
  for (int r = 0; r < 3; r++) center[r] = center3[r];
  glm::vec3 translation = -center * rotation;
  glm::mat4 eMatrix(0.0), kMatrix(0.0);
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) eMatrix[r][c] = rotation[r][c];
  eMatrix[0][3] = translation[0]; eMatrix[1][3] = translation[1]; eMatrix[2][3] = translation[2]; eMatrix[3][3] = 1.0;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) kMatrix[r][c] = intr[r][c];
  glm::mat4 cameraMatrix = eMatrix * kMatrix;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) out12[4 * r + c] = cameraMatrix[r][c];
  for (int r = 0; r < 3; r++) out_t3[r] = translation[r];
}

/* cam12: rows 0..2 of cameraMatrix as above (row 3 of the glm object is (0,0,0,*): kMatrix[3][3] = 0) */
void eg3d_ref_glm_project(const float* cam12, const float* x3, float* out2) {
  glm::mat4 m(0.0);
  for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) m[r][c] = cam12[4 * r + c];
  glm::vec4 p = glm::vec4(x3[0], x3[1], x3[2], 1.0) * m;
  glm::vec2 q(p[0] / p[2], p[1] / p[2]);
  out2[0] = q[0]; out2[1] = q[1];
}

}
