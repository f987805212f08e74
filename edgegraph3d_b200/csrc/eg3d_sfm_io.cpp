// eg3d_sfm_io.cpp — row f3 of SURVEY §8 in C++ (host code of libeg3d.so, no Python, no rapidjson): the data formats either side of
// the hot path.
//   eg3d_sfm_load   OpenMVG sfm_data JSON -> SfMData  (external/manifoldReconstructor/src/OpenMvgParser.cpp:75-153, 241-301;
//                   camera conventions: SURVEY A.1 — view index = position in `extrinsics`, translation = -center * rotation and
//                   cameraMatrix = eMatrix * kMatrix in the operation order of the reference's vendored glm 0.9.6, radial distortion ignored)
//   eg3d_sfm_save   output_sfm_data (src/edgegraph3d/io/output/output_sfm_data.cpp:186-229): views / intrinsics / extrinsics of the
//                   original file kept verbatim, `structure` rewritten from the given points (id_feat = OUTPUT_SFMD_FEATURE_ID = 0)
//   eg3d_write_ply  output_point_cloud.cpp
// edgegraph3d_b200/openmvg_io.py is the same thing in Python; tests/test_sfm_io.py holds the two against each other bit for bit
// (cameras, tracks) and through a save -> load round trip.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/eg3d.h"

namespace {

// ---- a small JSON document: values keep the [begin, end) span of their source text so that sub-trees can be copied verbatim
struct JVal {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  double num = 0; bool b = false; std::string str;
  std::vector<JVal> arr;
  std::vector<std::pair<std::string, JVal>> obj;
  size_t begin = 0, end = 0;
  const JVal* get(const char* key) const {
    if (kind != Obj) return nullptr;
    for (const auto& kv : obj) if (kv.first == key) return &kv.second;
    return nullptr;
  }
};

struct JParser {
  const std::string& s; size_t p = 0; bool ok = true;
  explicit JParser(const std::string& src) : s(src) {}
  void ws() { while (p < s.size() && (s[p] == ' ' || s[p] == '\n' || s[p] == '\t' || s[p] == '\r')) p++; }
  bool lit(const char* t) { size_t n = std::strlen(t); if (s.compare(p, n, t) == 0) { p += n; return true; } return false; }
  std::string string() {
    std::string out;
    if (p >= s.size() || s[p] != '"') { ok = false; return out; }
    p++;
    while (p < s.size() && s[p] != '"') {
      if (s[p] == '\\' && p + 1 < s.size()) {
        const char c = s[p + 1]; p += 2;
        switch (c) {
          case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break; case 'b': out += '\b'; break; case 'f': out += '\f'; break;
          case 'u': { if (p + 4 <= s.size()) { unsigned cp = (unsigned)std::strtoul(s.substr(p, 4).c_str(), nullptr, 16); p += 4; if (cp < 0x80) out += (char)cp; else { out += '?'; } } break; }
          default: out += c;
        }
      } else out += s[p++];
    }
    if (p >= s.size()) { ok = false; return out; }
    p++;
    return out;
  }
  JVal value() {
    JVal v; ws(); v.begin = p;
    if (p >= s.size()) { ok = false; return v; }
    const char c = s[p];
    if (c == '{') {
      v.kind = JVal::Obj; p++; ws();
      if (p < s.size() && s[p] == '}') { p++; v.end = p; return v; }
      while (ok) {
        ws(); std::string k = string(); ws();
        if (!ok || p >= s.size() || s[p] != ':') { ok = false; break; }
        p++;
        v.obj.emplace_back(std::move(k), value());
        ws();
        if (p < s.size() && s[p] == ',') { p++; continue; }
        if (p < s.size() && s[p] == '}') { p++; break; }
        ok = false;
      }
    } else if (c == '[') {
      v.kind = JVal::Arr; p++; ws();
      if (p < s.size() && s[p] == ']') { p++; v.end = p; return v; }
      while (ok) {
        v.arr.push_back(value());
        ws();
        if (p < s.size() && s[p] == ',') { p++; continue; }
        if (p < s.size() && s[p] == ']') { p++; break; }
        ok = false;
      }
    } else if (c == '"') { v.kind = JVal::Str; v.str = string(); }
    else if (lit("true")) { v.kind = JVal::Bool; v.b = true; }
    else if (lit("false")) { v.kind = JVal::Bool; v.b = false; }
    else if (lit("null")) { v.kind = JVal::Null; }
    else {
      char* e = nullptr;
      v.num = std::strtod(s.c_str() + p, &e);            // correctly rounded, as Python's float()
      if (e == s.c_str() + p) { ok = false; return v; }
      v.kind = JVal::Num; p = (size_t)(e - s.c_str());
    }
    v.end = p;
    return v;
  }
};

bool read_file(const char* path, std::string& out) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END); long n = std::ftell(f); std::fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? (size_t)n : 0);
  const size_t got = n > 0 ? std::fread(&out[0], 1, (size_t)n, f) : 0;
  std::fclose(f);
  return got == out.size();
}

inline float f32(const JVal* v) { return v && v->kind == JVal::Num ? (float)v->num : 0.f; }      // rapidjson GetFloat: double narrowed
inline long long i64(const JVal* v) { return v && v->kind == JVal::Num ? (long long)v->num : 0; }

// translation = -center * rotation (OpenMvgParser.cpp:289) and cameraMatrix = eMatrix * kMatrix (:107-125) in float, glm 0.9.6
// operation order: vec3 * mat3 -> m[i][0] v.x + m[i][1] v.y + m[i][2] v.z; mat4 * mat4 -> column c = m1[0] m2[c][0] + m1[1] m2[c][1]
// + m1[2] m2[c][2] + m1[3] m2[c][3], left to right, the zero terms included (bit-identical to real glm: tests/golden/glm_golden.npz
// through openmvg_io.glm_camera_matrix, which this repeats).  The file is compiled with -ffp-contract=off.
void glm_camera_matrix(const float R[9], const float C[3], const float K[9], float P[12], float t[3]) {
  const float nc[3] = {-C[0], -C[1], -C[2]};
  for (int i = 0; i < 3; i++) {
    const float a = R[3 * i] * nc[0], b = R[3 * i + 1] * nc[1], c = R[3 * i + 2] * nc[2];
    const float ab = a + b;
    t[i] = ab + c;
  }
  float E[4][4] = {{0}}, Km[4][4] = {{0}}, res[4][4];
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) { E[r][c] = R[3 * r + c]; Km[r][c] = K[3 * r + c]; } E[r][3] = t[r]; }
  E[3][3] = 1.f;
  for (int c = 0; c < 4; c++)
    for (int j = 0; j < 4; j++) {
      float acc = E[0][j] * Km[c][0];
      for (int k = 1; k < 4; k++) { const float term = E[k][j] * Km[c][k]; acc = acc + term; }
      res[c][j] = acc;
    }
  for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) P[4 * r + c] = res[r][c];
}

void put_double(std::string& out, double v) {           // shortest text that round-trips, as Python's repr(float)
  if (!std::isfinite(v)) { out += "null"; return; }
  char buf[40];
  auto r = std::to_chars(buf, buf + sizeof buf, v);
  std::string s(buf, r.ptr);
  if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
  out += s;
}

}  // namespace

struct eg3d_sfm {
  int32_t n_views = 0, width = 0, height = 0;
  std::vector<float> cameras, K, R, center, t;
  std::vector<int64_t> view_keys;
  std::vector<float> track_xyz, track_xy; std::vector<int64_t> track_off; std::vector<int32_t> track_view;
};

extern "C" {

eg3d_status eg3d_sfm_load(const char* path, eg3d_sfm** out) {
  if (!path || !out) return EG3D_ERR_INVALID_ARG;
  std::string text;
  if (!read_file(path, text)) return EG3D_ERR_INVALID_ARG;
  JParser jp(text);
  const JVal doc = jp.value();
  if (!jp.ok || doc.kind != JVal::Obj) return EG3D_ERR_INVALID_ARG;
  const JVal *jintr = doc.get("intrinsics"), *jviews = doc.get("views"), *jext = doc.get("extrinsics"), *jstruct = doc.get("structure");
  if (!jintr || !jviews || !jext || jintr->kind != JVal::Arr || jviews->kind != JVal::Arr || jext->kind != JVal::Arr || jintr->arr.empty()) return EG3D_ERR_INVALID_ARG;
  struct Intr { float K[9]; int w, h; };
  std::map<long long, Intr> intr; long long first_intr = 0; bool have_first = false;
  for (const JVal& it : jintr->arr) {
    const JVal* val = it.get("value"); const JVal* pw = val ? val->get("ptr_wrapper") : nullptr; const JVal* data = pw ? pw->get("data") : nullptr;
    if (!data) return EG3D_ERR_INVALID_ARG;
    const JVal* pp = data->get("principal_point");
    if (!pp || pp->kind != JVal::Arr || pp->arr.size() < 2) return EG3D_ERR_INVALID_ARG;
    Intr in; std::memset(&in, 0, sizeof in);
    const float f = f32(data->get("focal_length"));
    in.K[0] = f; in.K[4] = f; in.K[2] = f32(&pp->arr[0]); in.K[5] = f32(&pp->arr[1]); in.K[8] = 1.f;   // radial distortion is ignored (:252-256)
    in.w = (int)i64(data->get("width")); in.h = (int)i64(data->get("height"));
    const long long key = i64(it.get("key"));
    intr[key] = in;
    if (!have_first) { first_intr = key; have_first = true; }
  }
  std::map<long long, long long> intr_of_pose;          // views keyed by id_pose
  for (const JVal& v : jviews->arr) {
    const JVal* val = v.get("value"); const JVal* pw = val ? val->get("ptr_wrapper") : nullptr; const JVal* data = pw ? pw->get("data") : nullptr;
    if (!data) return EG3D_ERR_INVALID_ARG;
    intr_of_pose[i64(data->get("id_pose"))] = i64(data->get("id_intrinsic"));
  }
  std::unique_ptr<eg3d_sfm> s(new eg3d_sfm());
  std::map<long long, int> pos_of_pose;
  for (size_t pos = 0; pos < jext->arr.size(); pos++) {   // view index = position in `extrinsics` (:292)
    const JVal& ex = jext->arr[pos];
    const long long key = i64(ex.get("key"));
    pos_of_pose[key] = (int)pos;
    const JVal* val = ex.get("value"); const JVal* rot = val ? val->get("rotation") : nullptr; const JVal* cen = val ? val->get("center") : nullptr;
    if (!rot || !cen || rot->kind != JVal::Arr || rot->arr.size() != 3 || cen->kind != JVal::Arr || cen->arr.size() != 3) return EG3D_ERR_INVALID_ARG;
    float R[9], C[3];
    for (int r = 0; r < 3; r++) {
      if (rot->arr[r].kind != JVal::Arr || rot->arr[r].arr.size() != 3) return EG3D_ERR_INVALID_ARG;
      for (int c = 0; c < 3; c++) R[3 * r + c] = f32(&rot->arr[r].arr[c]);
      C[r] = f32(&cen->arr[r]);
    }
    auto io = intr_of_pose.find(key);
    const Intr& in = (io != intr_of_pose.end() && intr.count(io->second)) ? intr[io->second] : intr[first_intr];
    float P[12], t[3];
    glm_camera_matrix(R, C, in.K, P, t);
    s->cameras.insert(s->cameras.end(), P, P + 12); s->K.insert(s->K.end(), in.K, in.K + 9); s->R.insert(s->R.end(), R, R + 9);
    s->center.insert(s->center.end(), C, C + 3); s->t.insert(s->t.end(), t, t + 3); s->view_keys.push_back(key);
  }
  s->n_views = (int32_t)jext->arr.size();
  s->width = intr[first_intr].w; s->height = intr[first_intr].h;
  s->track_off.push_back(0);
  if (jstruct && jstruct->kind == JVal::Arr)
    for (const JVal& pt : jstruct->arr) {
      const JVal* val = pt.get("value"); const JVal* X = val ? val->get("X") : nullptr; const JVal* obs = val ? val->get("observations") : nullptr;
      if (!X || X->kind != JVal::Arr || X->arr.size() != 3) return EG3D_ERR_INVALID_ARG;
      for (int k = 0; k < 3; k++) s->track_xyz.push_back(f32(&X->arr[k]));
      if (obs && obs->kind == JVal::Arr)
        for (const JVal& ob : obs->arr) {
          const JVal* ov = ob.get("value"); const JVal* x = ov ? ov->get("x") : nullptr;
          auto pp = pos_of_pose.find(i64(ob.get("key")));
          if (!x || x->kind != JVal::Arr || x->arr.size() != 2 || pp == pos_of_pose.end()) return EG3D_ERR_INVALID_ARG;
          s->track_view.push_back(pp->second);
          s->track_xy.push_back(f32(&x->arr[0])); s->track_xy.push_back(f32(&x->arr[1]));
        }
      s->track_off.push_back((int64_t)s->track_view.size());
    }
  *out = s.release();
  return EG3D_OK;
}

eg3d_status eg3d_sfm_get(const eg3d_sfm* s, eg3d_sfm_view* v) {
  if (!s || !v) return EG3D_ERR_INVALID_ARG;
  v->n_views = s->n_views; v->width = s->width; v->height = s->height;
  v->cameras = s->cameras.data(); v->K = s->K.data(); v->R = s->R.data(); v->center = s->center.data(); v->t = s->t.data();
  v->view_keys = s->view_keys.data();
  v->n_tracks = (int64_t)s->track_off.size() - 1;
  v->track_xyz = s->track_xyz.data(); v->track_off = s->track_off.data(); v->track_view = s->track_view.data(); v->track_xy = s->track_xy.data();
  return EG3D_OK;
}

void eg3d_sfm_free(eg3d_sfm* s) { delete s; }

eg3d_status eg3d_sfm_save(const char* path, const char* original_path, int64_t n_points, const float* xyz, const int64_t* obs_off,
                          const int32_t* obs_view, const float* obs_xy, const uint8_t* keep, int64_t* n_written) {
  if (!path || !original_path || n_points < 0 || (n_points > 0 && (!xyz || !obs_off || !obs_view || !obs_xy))) return EG3D_ERR_INVALID_ARG;
  std::string text;
  if (!read_file(original_path, text)) return EG3D_ERR_INVALID_ARG;
  JParser jp(text);
  const JVal doc = jp.value();
  if (!jp.ok || doc.kind != JVal::Obj) return EG3D_ERR_INVALID_ARG;
  const JVal* jext = doc.get("extrinsics");
  if (!jext || jext->kind != JVal::Arr || !doc.get("views") || !doc.get("intrinsics")) return EG3D_ERR_INVALID_ARG;
  std::vector<long long> keys;
  for (const JVal& ex : jext->arr) keys.push_back(i64(ex.get("key")));
  auto raw = [&](const char* key, const char* dflt) { const JVal* v = doc.get(key); return v ? text.substr(v->begin, v->end - v->begin) : std::string(dflt); };
  std::string out;
  out.reserve(text.size() + (size_t)n_points * 256);
  out += "{\"sfm_data_version\": " + raw("sfm_data_version", "\"0.3\"") + ", \"root_path\": " + raw("root_path", "\"\"");
  out += ", \"views\": " + raw("views", "[]") + ", \"intrinsics\": " + raw("intrinsics", "[]") + ", \"extrinsics\": " + raw("extrinsics", "[]");
  out += ", \"structure\": [";
  int64_t written = 0;
  for (int64_t i = 0; i < n_points; i++) {
    if (keep && !keep[i]) continue;
    if (written) out += ", ";
    out += "{\"key\": " + std::to_string(i) + ", \"value\": {\"X\": [";
    for (int k = 0; k < 3; k++) { if (k) out += ", "; put_double(out, (double)xyz[3 * i + k]); }
    out += "], \"observations\": [";
    for (int64_t o = obs_off[i]; o < obs_off[i + 1]; o++) {
      const int v = obs_view[o];
      if (v < 0 || (size_t)v >= keys.size()) return EG3D_ERR_INVALID_ARG;
      if (o > obs_off[i]) out += ", ";
      out += "{\"key\": " + std::to_string(keys[(size_t)v]) + ", \"value\": {\"id_feat\": 0, \"x\": [";
      put_double(out, (double)obs_xy[2 * o]); out += ", "; put_double(out, (double)obs_xy[2 * o + 1]);
      out += "]}}";
    }
    out += "]}}";
    written++;
  }
  out += "], \"control_points\": " + raw("control_points", "[]") + "}";
  FILE* f = std::fopen(path, "wb");
  if (!f) return EG3D_ERR_INVALID_ARG;
  const size_t w = std::fwrite(out.data(), 1, out.size(), f);
  std::fclose(f);
  if (w != out.size()) return EG3D_ERR_INVALID_ARG;
  if (n_written) *n_written = written;
  return EG3D_OK;
}

eg3d_status eg3d_write_ply(const char* path, int64_t n, const float* xyz, const uint8_t* rgb) {
  if (!path || n < 0 || (n > 0 && !xyz)) return EG3D_ERR_INVALID_ARG;
  FILE* f = std::fopen(path, "w");
  if (!f) return EG3D_ERR_INVALID_ARG;
  std::fprintf(f, "ply\nformat ascii 1.0\nelement vertex %lld\nproperty float x\nproperty float y\nproperty float z\n", (long long)n);
  if (rgb) std::fprintf(f, "property uchar red\nproperty uchar green\nproperty uchar blue\n");
  std::fprintf(f, "end_header\n");
  for (int64_t i = 0; i < n; i++) {
    std::fprintf(f, "%g %g %g", xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    if (rgb) std::fprintf(f, " %d %d %d", rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    std::fprintf(f, "\n");
  }
  std::fclose(f);
  return EG3D_OK;
}

}  // extern "C"
