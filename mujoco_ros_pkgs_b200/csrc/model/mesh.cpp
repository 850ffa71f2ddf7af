// mesh.cpp — mesh assets for the model compiler: file readers (binary / ASCII STL, OBJ), 3-D convex hull, mass
// properties of the hull, centring and principal-axis alignment.
//
// The reference loads whatever `mj_loadXML` accepts (mujoco_ros/src/mujoco_env.cpp:771-911), meshes included; MuJoCo's
// compiler replaces every collision mesh by its convex hull (qhull), moves the vertices to the centre of mass, rotates
// them into the principal axes of inertia and folds that frame into the geom's pose.  Same steps here; the hull is an
// incremental construction of our own (no qhull in this image).  Inertia comes from the hull as a solid of uniform
// density (MuJoCo integrates the original triangle surface, identical for convex input).
#include "mesh.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>

namespace b2mj {

namespace {

typedef std::array<double, 3> V3;
V3 sub(const V3& a, const V3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
V3 crs(const V3& a, const V3& b) { return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}; }
double dt(const V3& a, const V3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

struct Face {
  int v[3];
  V3 n;      // outward normal (not normalised)
  double off;
  bool alive;
};

// signed distance-like quantity of point p above face f (positive = outside)
double above(const Face& f, const V3& p) { return dt(f.n, p) - f.off; }

Face make_face(const std::vector<V3>& P, int a, int b, int c) {
  Face f;
  f.v[0] = a; f.v[1] = b; f.v[2] = c;
  f.n = crs(sub(P[b], P[a]), sub(P[c], P[a]));
  f.off = dt(f.n, P[a]);
  f.alive = true;
  return f;
}

}  // namespace

// incremental convex hull; returns triangles over the indices of P, empty on degenerate input
bool convex_hull(const std::vector<double>& pts, std::vector<int>& tri, std::string& err) {
  const int n = (int)pts.size() / 3;
  std::vector<V3> P(n);
  double scale = 0;
  for (int i = 0; i < n; i++) {
    P[i] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    for (int k = 0; k < 3; k++) scale = std::max(scale, std::fabs(P[i][k]));
  }
  if (n < 4) { err = "a mesh needs at least 4 vertices"; return false; }
  const double eps = 1e-10 * std::max(scale, 1e-300);
  // initial tetrahedron: extreme points along x, farthest from that line, farthest from that plane
  int i0 = 0, i1 = 0;
  for (int i = 0; i < n; i++) { if (P[i][0] < P[i0][0]) i0 = i; if (P[i][0] > P[i1][0]) i1 = i; }
  if (i0 == i1) { for (int i = 0; i < n; i++) if (dt(sub(P[i], P[i0]), sub(P[i], P[i0])) > dt(sub(P[i1], P[i0]), sub(P[i1], P[i0]))) i1 = i; }
  int i2 = -1;
  double best = eps * eps;
  for (int i = 0; i < n; i++) {
    const V3 c = crs(sub(P[i1], P[i0]), sub(P[i], P[i0]));
    if (dt(c, c) > best) { best = dt(c, c); i2 = i; }
  }
  if (i2 < 0) { err = "mesh vertices are collinear"; return false; }
  int i3 = -1;
  best = 0;
  const V3 nrm = crs(sub(P[i1], P[i0]), sub(P[i2], P[i0]));
  const double nl = std::sqrt(dt(nrm, nrm));
  for (int i = 0; i < n; i++) {
    const double d = std::fabs(dt(nrm, sub(P[i], P[i0]))) / nl;
    if (d > best) { best = d; i3 = i; }
  }
  if (i3 < 0 || best <= eps) { err = "mesh vertices are coplanar"; return false; }
  if (dt(nrm, sub(P[i3], P[i0])) > 0) std::swap(i1, i2);  // make (i0, i1, i2) face away from i3
  std::vector<Face> F;
  F.push_back(make_face(P, i0, i1, i2));
  F.push_back(make_face(P, i0, i3, i1));
  F.push_back(make_face(P, i1, i3, i2));
  F.push_back(make_face(P, i2, i3, i0));
  for (int i = 0; i < n; i++) {
    if (i == i0 || i == i1 || i == i2 || i == i3) continue;
    // faces visible from P[i]
    std::vector<int> vis;
    for (int f = 0; f < (int)F.size(); f++) {
      if (!F[f].alive) continue;
      const double len = std::sqrt(dt(F[f].n, F[f].n));
      if (above(F[f], P[i]) > eps * len) vis.push_back(f);
    }
    if (vis.empty()) continue;  // inside (or on) the current hull
    // horizon: directed edges of visible faces whose reverse is not an edge of a visible face
    std::set<std::pair<int, int>> edges;
    for (int f : vis)
      for (int k = 0; k < 3; k++) edges.insert({F[f].v[k], F[f].v[(k + 1) % 3]});
    std::vector<std::pair<int, int>> horizon;
    for (auto& e : edges)
      if (!edges.count({e.second, e.first})) horizon.push_back(e);
    for (int f : vis) F[f].alive = false;
    for (auto& e : horizon) F.push_back(make_face(P, e.first, e.second, i));
  }
  tri.clear();
  for (auto& f : F)
    if (f.alive) { tri.push_back(f.v[0]); tri.push_back(f.v[1]); tri.push_back(f.v[2]); }
  if (tri.size() < 12) { err = "convex hull construction failed"; return false; }
  return true;
}

// symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations: A = V diag(w) V'
static void eig3(const double A[9], double w[3], double V[9]) {
  double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) a[i][j] = A[3 * i + j];
  for (int sweep = 0; sweep < 64; sweep++) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    if (off < 1e-300) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (std::fabs(a[p][q]) < 1e-300) continue;
        const double th = (a[q][q] - a[p][p]) / (2 * a[p][q]);
        const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1));
        const double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 3; k++) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 3; i++) {
    w[i] = a[i][i];
    for (int j = 0; j < 3; j++) V[3 * i + j] = v[i][j];
  }
}

static void mat2quat(double q[4], const double R[9]) {
  // columns of R are the axes; standard branch on the largest diagonal term
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    const double s = std::sqrt(tr + 1) * 2;
    q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    const double s = std::sqrt(1 + R[0] - R[4] - R[8]) * 2;
    q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    const double s = std::sqrt(1 + R[4] - R[0] - R[8]) * 2;
    q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
  } else {
    const double s = std::sqrt(1 + R[8] - R[0] - R[4]) * 2;
    q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
  }
  const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; k++) q[k] /= nq;
}

bool mesh_process(const std::vector<double>& raw, const double scale[3], MeshData& out, std::string& err) {
  std::vector<double> pts(raw);
  for (size_t i = 0; i < pts.size(); i++) pts[i] *= scale[i % 3];
  // merge exact duplicates (STL repeats every vertex per facet)
  {
    std::map<std::array<double, 3>, int> seen;
    std::vector<double> uniq;
    for (size_t i = 0; i + 2 < pts.size(); i += 3) {
      std::array<double, 3> k = {pts[i], pts[i + 1], pts[i + 2]};
      if (seen.emplace(k, (int)uniq.size() / 3).second) uniq.insert(uniq.end(), k.begin(), k.end());
    }
    pts.swap(uniq);
  }
  std::vector<int> tri;
  if (!convex_hull(pts, tri, err)) return false;
  // volume, centre of mass and covariance of the hull as signed tetrahedra over the origin
  double vol = 0, com[3] = {0, 0, 0}, C[9] = {0};
  for (size_t t = 0; t + 2 < tri.size(); t += 3) {
    const double *a = &pts[3 * tri[t]], *b = &pts[3 * tri[t + 1]], *c = &pts[3 * tri[t + 2]];
    const double det = a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
    vol += det / 6;
    double s[3];
    for (int k = 0; k < 3; k++) { s[k] = a[k] + b[k] + c[k]; com[k] += det / 24 * s[k]; }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) C[3 * i + j] += det / 120 * (s[i] * s[j] + a[i] * a[j] + b[i] * b[j] + c[i] * c[j]);
  }
  if (vol < 1e-300) { err = "mesh hull has no volume"; return false; }
  for (int k = 0; k < 3; k++) com[k] /= vol;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[3 * i + j] -= vol * com[i] * com[j];
  const double trC = C[0] + C[4] + C[8];
  double I[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) I[3 * i + j] = (i == j ? trC : 0.0) - C[3 * i + j];
  double w[3], V[9];
  eig3(I, w, V);
  // principal axes sorted by decreasing moment, right-handed
  int ord[3] = {0, 1, 2};
  std::sort(ord, ord + 3, [&](int x, int y) { return w[x] > w[y]; });
  double R[9];
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++) R[3 * i + j] = V[3 * i + ord[j]];
  {
    const double c0[3] = {R[0], R[3], R[6]}, c1[3] = {R[1], R[4], R[7]};
    const double cz[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2], c0[0] * c1[1] - c0[1] * c1[0]};
    if (cz[0] * R[2] + cz[1] * R[5] + cz[2] * R[8] < 0) { R[2] = -R[2]; R[5] = -R[5]; R[8] = -R[8]; }
  }
  for (int k = 0; k < 3; k++) { out.pos[k] = com[k]; out.inertia[k] = w[ord[k]]; }
  mat2quat(out.quat, R);
  out.volume = vol;
  // hull vertices in the mesh frame, in order of first appearance in the hull triangles
  std::map<int, int> remap;
  out.vert.clear();
  out.rbound = 0;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int idx : tri) {
    if (remap.count(idx)) continue;
    remap[idx] = (int)out.vert.size() / 3;
    const double d[3] = {pts[3 * idx] - com[0], pts[3 * idx + 1] - com[1], pts[3 * idx + 2] - com[2]};
    double l[3];
    for (int j = 0; j < 3; j++) l[j] = R[j] * d[0] + R[3 + j] * d[1] + R[6 + j] * d[2];  // R' d
    out.vert.insert(out.vert.end(), l, l + 3);
    out.rbound = std::max(out.rbound, std::sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]));
    for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], l[k]); hi[k] = std::max(hi[k], l[k]); }
  }
  for (int k = 0; k < 3; k++) out.aabb[k] = std::max(std::fabs(lo[k]), std::fabs(hi[k]));
  return true;
}

bool mesh_read_file(const std::string& path, std::vector<double>& pts, std::string& err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { err = "cannot open mesh file '" + path + "'"; return false; }
  std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::string ext = path.size() >= 4 ? path.substr(path.size() - 4) : "";
  for (auto& ch : ext) ch = (char)std::tolower((unsigned char)ch);
  pts.clear();
  if (ext == ".obj") {
    std::istringstream is(data);
    std::string line;
    while (std::getline(is, line)) {
      if (line.size() > 2 && line[0] == 'v' && (line[1] == ' ' || line[1] == '\t')) {
        std::istringstream ls(line.substr(2));
        double x, y, z;
        if (ls >> x >> y >> z) { pts.push_back(x); pts.push_back(y); pts.push_back(z); }
      }
    }
  } else if (ext == ".stl") {
    bool binary = data.size() >= 84;
    if (binary) {
      uint32_t ntri;
      std::memcpy(&ntri, data.data() + 80, 4);
      binary = data.size() == 84 + (size_t)ntri * 50;
      if (binary) {
        for (uint32_t t = 0; t < ntri; t++) {
          const char* p = data.data() + 84 + (size_t)t * 50 + 12;
          for (int k = 0; k < 9; k++) {
            float v;
            std::memcpy(&v, p + 4 * k, 4);
            pts.push_back(v);
          }
        }
      }
    }
    if (!binary) {  // ASCII
      std::istringstream is(data);
      std::string tok;
      while (is >> tok) {
        if (tok == "vertex") {
          double x, y, z;
          if (is >> x >> y >> z) { pts.push_back(x); pts.push_back(y); pts.push_back(z); }
        }
      }
    }
  } else {
    err = "mesh file '" + path + "': only .stl and .obj are supported";
    return false;
  }
  if (pts.size() < 12) { err = "mesh file '" + path + "' has fewer than 4 vertices"; return false; }
  return true;
}

static thread_local std::string g_model_dir;
void set_model_dir(const std::string& dir) { g_model_dir = dir; }
const std::string& model_dir() { return g_model_dir; }

}  // namespace b2mj
