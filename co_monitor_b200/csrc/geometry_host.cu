// Host-side geometry analysis (pure CPU, fp64): apertures -> plane + convex polygon, and the FP32
// filter tables derived from them.  No CUDA calls in this file.
//
// Reference semantics reduced here:
//   simulation.py:198-200   plane normal N = (p2-p1) x (p3-p2), plane point p1
//   simulation.py:226-231   slab = vertices +/- orientation_vector
//   simulation.py:260-261   Delaunay(slab).find_simplex(X) >= 0   with X on the plane
// => X passes iff it lies in the cross-section polygon of conv(slab) with the plane.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <cuda_runtime.h>

#include "scene.h"

namespace com {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_error; }

namespace {

struct V3 {
  double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 load(const double* p) { return {p[0], p[1], p[2]}; }
inline void store(double* p, V3 a) { p[0] = a.x, p[1] = a.y, p[2] = a.z; }

struct P2 {
  double x, y;
};
inline double cross2(P2 o, P2 a, P2 b) { return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x); }

// Andrew monotone chain, counter-clockwise, strictly convex (collinear points dropped).
std::vector<P2> convex_hull(std::vector<P2> pts, double merge_tol, double flat_tol) {
  std::sort(pts.begin(), pts.end(), [](P2 a, P2 b) { return a.x < b.x || (a.x == b.x && a.y < b.y); });
  std::vector<P2> uniq;
  for (auto p : pts) {
    bool dup = false;
    for (auto q : uniq)
      if (std::hypot(p.x - q.x, p.y - q.y) <= merge_tol) {
        dup = true;
        break;
      }
    if (!dup) uniq.push_back(p);
  }
  pts.swap(uniq);
  std::sort(pts.begin(), pts.end(), [](P2 a, P2 b) { return a.x < b.x || (a.x == b.x && a.y < b.y); });
  size_t n = pts.size();
  if (n < 3) return pts;
  std::vector<P2> h(2 * n);
  size_t k = 0;
  for (size_t i = 0; i < n; ++i) {
    while (k >= 2 && cross2(h[k - 2], h[k - 1], pts[i]) <= 0) --k;
    h[k++] = pts[i];
  }
  for (size_t i = n - 1, t = k + 1; i > 0; --i) {
    while (k >= t && cross2(h[k - 2], h[k - 1], pts[i - 1]) <= 0) --k;
    h[k++] = pts[i - 1];
  }
  h.resize(k - 1);
  // drop vertices that stick out of the line through their neighbours by less than flat_tol:
  // such an edge pair has no numerical meaning at the reference's own round-off level.
  bool changed = true;
  while (changed && h.size() > 3) {
    changed = false;
    for (size_t i = 0; i < h.size(); ++i) {
      P2 a = h[(i + h.size() - 1) % h.size()], b = h[i], c = h[(i + 1) % h.size()];
      double len = std::hypot(c.x - a.x, c.y - a.y);
      if (len == 0 || std::fabs(cross2(a, c, b)) / len < flat_tol) {
        h.erase(h.begin() + i);
        changed = true;
        break;
      }
    }
  }
  return h;
}

void edges_from_vertices(const std::vector<P2>& v, double (*edge)[3]) {
  size_t n = v.size();
  for (size_t k = 0; k < n; ++k) {
    P2 a = v[k], b = v[(k + 1) % n];
    double dx = b.x - a.x, dy = b.y - a.y, len = std::hypot(dx, dy);
    double ea = -dy / len, eb = dx / len;  // inward normal of a CCW polygon
    edge[k][0] = ea;
    edge[k][1] = eb;
    edge[k][2] = -(ea * a.x + eb * a.y);
  }
}

}  // namespace

int aperture_from_slab(const double* verts, int n, const double* offset, ComAperture* out) {
  if (!verts || !offset || !out || n < 3 || n > 16) {
    set_error("com_aperture_from_slab: need 3..16 vertices (got %d)", n);
    return COM_E_BADARG;
  }
  std::memset(out, 0, sizeof(*out));
  V3 p1 = load(verts), p2 = load(verts + 3), p3 = load(verts + 6), off = load(offset);
  V3 N = cross(p2 - p1, p3 - p2);
  double nn = norm(N), l12 = norm(p2 - p1);
  if (!(nn > 0) || !(l12 > 0)) {
    set_error("aperture plane is degenerate (first three vertices collinear)");
    return COM_E_GEOMETRY;
  }
  V3 nh = (1.0 / nn) * N, e1 = (1.0 / l12) * (p2 - p1), e2 = cross(nh, e1);
  store(out->p1, p1), store(out->normal, N), store(out->e1, e1), store(out->e2, e2);

  std::vector<V3> slab;
  for (int i = 0; i < n; ++i) slab.push_back(load(verts + 3 * i) + off);
  for (int i = 0; i < n; ++i) slab.push_back(load(verts + 3 * i) - off);
  std::vector<double> dist(slab.size());
  double scale = 0;
  for (size_t i = 0; i < slab.size(); ++i) {
    dist[i] = dot(nh, slab[i] - p1);
    scale = std::max(scale, norm(slab[i] - p1));
  }
  const double on_plane = 1e-12 * std::max(1.0, scale);
  std::vector<P2> pts;
  auto push = [&](V3 X) { pts.push_back({dot(X - p1, e1), dot(X - p1, e2)}); };
  for (size_t i = 0; i < slab.size(); ++i) {
    if (std::fabs(dist[i]) <= on_plane) push(slab[i]);
    for (size_t j = i + 1; j < slab.size(); ++j)
      if ((dist[i] < -on_plane && dist[j] > on_plane) || (dist[i] > on_plane && dist[j] < -on_plane)) {
        double t = dist[i] / (dist[i] - dist[j]);
        push(slab[i] + t * (slab[j] - slab[i]));
      }
  }
  std::vector<P2> hull = convex_hull(pts, 1e-10, 1e-10);
  if (hull.size() < 3) {
    set_error("aperture slab does not cut its own plane in a polygon (%zu section points)", hull.size());
    return COM_E_GEOMETRY;
  }
  if (hull.size() > COM_MAX_EDGES) {
    set_error("aperture section has %zu edges (max %d)", hull.size(), COM_MAX_EDGES);
    return COM_E_GEOMETRY;
  }
  out->n_edges = (int)hull.size();
  for (size_t k = 0; k < hull.size(); ++k) out->vertex[k][0] = hull[k].x, out->vertex[k][1] = hull[k].y;
  edges_from_vertices(hull, out->edge);
  return COM_OK;
}

int aperture_contains(const ComAperture* ap, const double* point) {
  V3 r = load(point) - load(ap->p1);
  double x = dot(r, load(ap->e1)), y = dot(r, load(ap->e2));
  for (int k = 0; k < ap->n_edges; ++k)
    if (!(ap->edge[k][0] * x + ap->edge[k][1] * y + ap->edge[k][2] >= 0)) return 0;
  return 1;
}

namespace {

std::vector<P2> vertices_of(const std::vector<std::array<double, 3>>& e) {
  std::vector<P2> v;
  for (size_t k = 0; k < e.size(); ++k) {  // vertex k = edge k-1 /\ edge k  (start of edge k)
    auto& a = e[(k + e.size() - 1) % e.size()];
    auto& b = e[k];
    double det = a[0] * b[1] - a[1] * b[0];
    v.push_back({(-a[2] * b[1] + b[2] * a[1]) / det, (-a[0] * b[2] + b[0] * a[2]) / det});
  }
  return v;
}

}  // namespace

int build_host_scene(const ComSceneDesc& d, HostScene& hs) {
  if (d.n_crystal <= 0 || d.n_slits <= 0 || !d.crystal_xyz || !d.crystal_normal || !d.slit_crys_side ||
      !d.slit_plasma_side || !d.port_vertices || !d.detector_vertices || d.n_port_vertices < 3 ||
      d.n_port_vertices > 16 || d.n_detector_vertices < 3 || d.n_detector_vertices > 16 ||
      d.n_voxel_corners < 0 || d.n_voxel_corners > 8 || d.n_shields < 0 || d.n_shields > COM_MAX_SHIELDS ||
      (d.n_shields > 0 && !d.shields)) {
    set_error("com_scene_create: incomplete or inconsistent scene description");
    return COM_E_BADARG;
  }
  const int C = d.n_crystal, S = d.n_slits;
  DeviceScene& dv = hs.dev;
  dv.n_crystal = C;
  dv.n_crystal_padded = (C + 31) / 32 * 32;
  dv.n_slits = S;
  dv.n_apertures = 2 * S + 2;
  for (int i = 0; i < 3; ++i) dv.origin[i] = d.origin[i];
  dv.area_pt = d.area_pt, dv.refl_max = d.refl_max, dv.aoi0 = d.aoi0, dv.refl_sigma = d.refl_sigma;
  hs.crystal_xyz.assign(d.crystal_xyz, d.crystal_xyz + 3 * (size_t)C);
  hs.crystal_normal.assign(d.crystal_normal, d.crystal_normal + 3 * (size_t)C);

  // ---- apertures: crys_0, plasma_0, crys_1, ..., port, detector -------------------------------------
  dv.near_edge_mm = d.near_edge_mm > 0 ? d.near_edge_mm : 1e-9;
  dv.near_band_mm = std::max(dv.near_edge_mm, d.near_band_mm > 0 ? d.near_band_mm : 1e-4);
  dv.physics_flags = d.physics_flags;
  dv.n_shields = d.n_shields;
  for (int s = 0; s < d.n_shields; ++s) {
    const ComShield& sh = d.shields[s];
    const V3 a = load(sh.axis), u = load(sh.u), v = load(sh.v);
    if (!(sh.radius > 0) || sh.n_sides < 0 || sh.n_sides == 1 || sh.n_sides == 2 || std::fabs(norm(a) - 1) > 1e-9 ||
        std::fabs(norm(u) - 1) > 1e-9 || std::fabs(norm(v) - 1) > 1e-9 || std::fabs(dot(a, u)) > 1e-9 ||
        std::fabs(dot(a, v)) > 1e-9 || std::fabs(dot(u, v)) > 1e-9) {
      set_error("shield %d: radius must be positive and (axis, u, v) orthonormal", s);
      return COM_E_GEOMETRY;
    }
    dv.shields[s] = sh;
  }
  hs.apertures.assign(dv.n_apertures, ComAperture{});
  if (d.apertures_override) {
    for (int a = 0; a < dv.n_apertures; ++a) {
      hs.apertures[a] = d.apertures_override[a];
      ComAperture& ap = hs.apertures[a];
      if (ap.n_edges < 3 || ap.n_edges > COM_MAX_EDGES) {
        set_error("apertures_override[%d] has %d edges", a, ap.n_edges);
        return COM_E_GEOMETRY;
      }
      std::vector<std::array<double, 3>> e;
      for (int k = 0; k < ap.n_edges; ++k) e.push_back({ap.edge[k][0], ap.edge[k][1], ap.edge[k][2]});
      std::vector<P2> v = vertices_of(e);  // keep the vertex list consistent with the (possibly shifted) edges
      for (int k = 0; k < ap.n_edges; ++k) ap.vertex[k][0] = v[k].x, ap.vertex[k][1] = v[k].y;
    }
  } else
  for (int k = 0; k < S; ++k) {
    int rc = aperture_from_slab(d.slit_crys_side + 12 * k, 4, d.slit_depth, &hs.apertures[2 * k]);
    if (rc) return rc;
    rc = aperture_from_slab(d.slit_plasma_side + 12 * k, 4, d.slit_depth, &hs.apertures[2 * k + 1]);
    if (rc) return rc;
  }
  if (!d.apertures_override) {
    int rc = aperture_from_slab(d.port_vertices, d.n_port_vertices, d.port_depth, &hs.apertures[2 * S]);
    if (rc) return rc;
    rc = aperture_from_slab(d.detector_vertices, d.n_detector_vertices, d.detector_depth, &hs.apertures[2 * S + 1]);
    if (rc) return rc;
  }

  // ---- compact fp64 apertures for the exact kernel ------------------------------------------------------
  hs.dev_apertures.clear();
  hs.edges.clear();
  for (const ComAperture& ap : hs.apertures) {
    DevAperture da{};
    for (int i = 0; i < 3; ++i) da.p1[i] = ap.p1[i], da.normal[i] = ap.normal[i], da.e1[i] = ap.e1[i], da.e2[i] = ap.e2[i];
    da.n_edges = ap.n_edges;
    da.edge_offset = (int)(hs.edges.size() / 3);
    for (int k = 0; k < ap.n_edges; ++k)
      for (int i = 0; i < 3; ++i) hs.edges.push_back(ap.edge[k][i]);
    hs.dev_apertures.push_back(da);
  }
  dv.n_edges_total = (int)(hs.edges.size() / 3);

  // ---- FP32 filter ----------------------------------------------------------------------------------
  // Every aperture test of the filter is "bounding rectangle of the polygon, grown by eps": a superset of the
  // polygon, so the filter can only over-accept; the exact kernel decides.
  double eps = d.filter_eps_mm > 0 ? d.filter_eps_mm : 5e-3;
  dv.filter_disabled = eps >= 1e30 ? 1 : 0;
  dv.no_run_cull = (d.filter_flags & COM_FILTER_NO_RUN_CULL) ? 1 : 0;
  if (dv.filter_disabled) eps = 5e-3;

  // slits: one polygon per side, repeated with period T along e1.  Verify the periodicity the collimator
  // builder guarantees (collimator.py:103-119); otherwise the slit stage of the filter is switched off.
  const ComAperture& c0 = hs.apertures[0];
  const ComAperture& p0 = hs.apertures[1];
  double period = 0;
  bool periodic = true;
  if (S > 1) {
    period = dot(load(hs.apertures[2].p1) - load(c0.p1), load(c0.e1));
    for (int k = 1; k < S && periodic; ++k)
      for (int side = 0; side < 2 && periodic; ++side) {
        const ComAperture& a0 = hs.apertures[side];
        const ComAperture& ak = hs.apertures[2 * k + side];
        V3 sh = load(ak.p1) - load(a0.p1);
        if (std::fabs(dot(sh, load(a0.e1)) - k * period) > 1e-6 || std::fabs(dot(sh, load(a0.e2))) > 1e-6 ||
            std::fabs(dot(sh, load(a0.normal))) / norm(load(a0.normal)) > 1e-6 || ak.n_edges != a0.n_edges)
          periodic = false;
        for (int e = 0; e < a0.n_edges && periodic; ++e)
          if (std::hypot(ak.vertex[e][0] - a0.vertex[e][0], ak.vertex[e][1] - a0.vertex[e][1]) > 1e-4)
            periodic = false;
      }
    if (!(period > 1e-3)) periodic = false;
  } else {
    period = 1e9;
  }
  dv.slit_filter = periodic ? 1 : 0;
  {
    const V3 nc = (1.0 / norm(load(c0.normal))) * load(c0.normal), np = (1.0 / norm(load(p0.normal))) * load(p0.normal);
    dv.slit_same_w = norm(nc - np) < 1e-9 ? 1 : 0;
  }
  dv.slit_period = (float)period;
  dv.slit_inv_period = (float)(1.0 / period);
  dv.slit_period_d = periodic ? period : 0.0;
  dv.slit_xlo_d = 1e300, dv.slit_xhi_d = -1e300;
  for (int k = 0; k < c0.n_edges; ++k) {
    dv.slit_xlo_d = std::min(dv.slit_xlo_d, c0.vertex[k][0]);
    dv.slit_xhi_d = std::max(dv.slit_xhi_d, c0.vertex[k][0]);
  }
  auto rect_of = [&](const ComAperture& ap, double grow, double& xlo, double& xhi, double& ylo, double& yhi) {
    xlo = ylo = 1e300, xhi = yhi = -1e300;
    for (int k = 0; k < ap.n_edges; ++k) {
      xlo = std::min(xlo, ap.vertex[k][0]), xhi = std::max(xhi, ap.vertex[k][0]);
      ylo = std::min(ylo, ap.vertex[k][1]), yhi = std::max(yhi, ap.vertex[k][1]);
    }
    xlo -= grow, xhi += grow, ylo -= grow, yhi += grow;
  };
  double xlo, xhi, ylo, yhi;
  rect_of(c0, eps, xlo, xhi, ylo, yhi);
  dv.rect_crys = SlitRect{(float)xlo, (float)xhi, (float)ylo, (float)yhi};
  rect_of(p0, eps, xlo, xhi, ylo, yhi);
  dv.rect_plasma = SlitRect{(float)xlo, (float)xhi, (float)ylo, (float)yhi};

  // Inner rectangles: the bounding rectangle shrunk until its four corners (hence, by convexity, all of it) are at
  // least `safe` inside every edge of the polygon.  FP32 evaluation of the in-plane coordinates is good to ~1e-4 mm
  //  and the slit polygons agree with slit 0's to 1e-4 mm (checked above): 2e-2 mm is ample.
  // A ray the filter calls "certain" skips the fp64 aperture tests altogether, so the extensions that add tests of their
  // own (segment check, shields) and the detector image (which needs the fp64 detector hit) switch the shortcut off.
  const double safe = std::max(5e-3, 2.0 * dv.near_band_mm);
  dv.certain_mm = (float)safe;
  bool okc_slits = false;
  auto inner_of = [&](const ComAperture& ap, double& ixlo, double& ixhi, double& iylo, double& iyhi) {
    rect_of(ap, -safe, ixlo, ixhi, iylo, iyhi);
    for (int it = 0; it < 4000 && ixlo < ixhi && iylo < iyhi; ++it) {
      bool inside = true;
      const double cx[4] = {ixlo, ixhi, ixhi, ixlo}, cy[4] = {iylo, iylo, iyhi, iyhi};
      for (int k = 0; k < ap.n_edges && inside; ++k)
        for (int c = 0; c < 4 && inside; ++c)
          inside = ap.edge[k][0] * cx[c] + ap.edge[k][1] * cy[c] + ap.edge[k][2] >= safe;
      if (inside) return true;
      ixlo += 5e-3, ixhi -= 5e-3, iylo += 5e-3, iyhi -= 5e-3;
    }
    ixlo = 1.0, ixhi = -1.0, iylo = 1.0, iyhi = -1.0;  // no inner rectangle: shortcut off
    return false;
  };
  {
    double a = 1, b = -1, c2 = 1, d2 = -1;
    // every slit must share slit 0's inner rectangle: take it from slit 0 and verify it on the others
    bool okc = periodic && inner_of(c0, a, b, c2, d2);
    dv.inner_crys = SlitRect{(float)a, (float)b, (float)c2, (float)d2};
    bool okp = periodic && inner_of(p0, a, b, c2, d2);
    dv.inner_plasma = SlitRect{(float)a, (float)b, (float)c2, (float)d2};
    for (int k = 1; k < S && okc && okp; ++k)
      for (int side = 0; side < 2; ++side) {
        const ComAperture& ak = hs.apertures[2 * k + side];
        const SlitRect& r = side ? dv.inner_plasma : dv.inner_crys;
        const double cx[4] = {r.xlo, r.xhi, r.xhi, r.xlo}, cy[4] = {r.ylo, r.ylo, r.yhi, r.yhi};
        for (int e = 0; e < ak.n_edges; ++e)
          for (int c = 0; c < 4; ++c)
            if (!(ak.edge[e][0] * cx[c] + ak.edge[e][1] * cy[c] + ak.edge[e][2] >= 0.5 * safe)) okc = false;
      }
    if (!okc || !okp || dv.physics_flags != 0 || dv.n_shields > 0)
      dv.inner_crys = dv.inner_plasma = SlitRect{1.f, -1.f, 1.f, -1.f};
    okc_slits = okc && okp;
  }

  hs.refl_table.resize(18001 + 18002);
  for (int m = 0; m <= 18000; ++m) {  // the expressions of the fp64 kernel, operation by operation
    const double aoi = (180.0 - (double)m / 100.0) / 2.0;
    const double da = aoi - dv.aoi0;
    hs.refl_table[m] = dv.refl_max * std::exp(-(da * da) / (2.0 * (dv.refl_sigma * dv.refl_sigma)));
  }
  // ... followed by the cosines of the rounding boundaries: entry 18001 + j = cos((j - 0.5) / 100 degrees), j = 0..18001;
  // rint(100 deg) = m  <=>  cos_bound[m + 1] < cos <= cos_bound[m]  (the cosine decreases with the angle)
  for (int j = 0; j <= 18001; ++j)
    hs.refl_table[18001 + j] = std::cos(((double)j - 0.5) * (3.141592653589793 / 18000.0));
  hs.refl_table[18001] = 2.0, hs.refl_table[18001 + 18001] = -2.0;  // the angle lies in [0, 180] degrees: no boundary beyond
  const int Cp = dv.n_crystal_padded;
  hs.det_forms.assign((size_t)4 * 3 * Cp, 0.f);
  hs.slit_forms.assign((size_t)4 * 6 * Cp, 0.f);
  hs.port_forms.assign((size_t)4 * 3 * Cp, 0.f);
  const ComAperture& det = hs.apertures[2 * S + 1];
  const ComAperture& port = hs.apertures[2 * S];
  double dxlo, dxhi, dylo, dyhi;
  rect_of(det, eps, dxlo, dxhi, dylo, dyhi);
  const double xc = 0.5 * (dxlo + dxhi), yc = 0.5 * (dylo + dyhi), hw = 0.5 * (dxhi - dxlo), hh = 0.5 * (dyhi - dylo);
  {  // certain regions of detector and port, and whether the column-interval filter applies
    double a, b, c2, d2;
    const bool okd = inner_of(det, a, b, c2, d2);
    dv.det_inner = okd ? SlitRect{(float)((a - xc) / hw), (float)((b - xc) / hw), (float)((c2 - yc) / hh), (float)((d2 - yc) / hh)}
                       : SlitRect{1.f, -1.f, 1.f, -1.f};
    rect_of(port, eps, a, b, c2, d2);
    dv.port_outer = SlitRect{(float)a, (float)b, (float)c2, (float)d2};
    // the port is a wide 14-gon: its certain region is the bounding rectangle scaled about its centre until the four
    // corners are inside by `safe`
    auto inner_scaled = [&](const ComAperture& ap, double& ixlo, double& ixhi, double& iylo, double& iyhi) {
      double bxlo, bxhi, bylo, byhi;
      rect_of(ap, 0.0, bxlo, bxhi, bylo, byhi);
      const double cx = 0.5 * (bxlo + bxhi), cy = 0.5 * (bylo + byhi), hx = 0.5 * (bxhi - bxlo), hy = 0.5 * (byhi - bylo);
      for (double f = 0.99; f > 0.02; f -= 0.01) {
        ixlo = cx - f * hx, ixhi = cx + f * hx, iylo = cy - f * hy, iyhi = cy + f * hy;
        bool inside = true;
        const double qx[4] = {ixlo, ixhi, ixhi, ixlo}, qy[4] = {iylo, iylo, iyhi, iyhi};
        for (int k = 0; k < ap.n_edges && inside; ++k)
          for (int c = 0; c < 4 && inside; ++c) inside = ap.edge[k][0] * qx[c] + ap.edge[k][1] * qy[c] + ap.edge[k][2] >= safe;
        if (inside) return true;
      }
      return false;
    };
    const bool okq = inner_scaled(port, a, b, c2, d2);
    dv.port_inner = okq ? SlitRect{(float)a, (float)b, (float)c2, (float)d2} : SlitRect{1.f, -1.f, 1.f, -1.f};
    // the interval filter visits every slit a column can reach and emits the rays of each: adjacent slit rectangles
    // must not overlap (a ray in two of them would be emitted twice)
    for (int side = 0; side < 2; ++side) {
      const ComAperture& ap = hs.apertures[side];
      dv.n_slit_edge[side] = std::min(ap.n_edges, 8);
      for (int k = 0; k < dv.n_slit_edge[side]; ++k)
        dv.slit_edge[side][k] = make_float4((float)ap.edge[k][0], (float)ap.edge[k][1], (float)(ap.edge[k][2] + eps),
                                            (float)(ap.edge[k][2] - safe));
    }
    dv.n_det_edge = (okd && det.n_edges <= 8) ? det.n_edges : 0;
    for (int k = 0; k < dv.n_det_edge; ++k) {  // x = xc + hw Xs / W, y = yc + hh Ys / W
      const double A = det.edge[k][0] * hw, B = det.edge[k][1] * hh, Cc = det.edge[k][0] * xc + det.edge[k][1] * yc + det.edge[k][2];
      dv.det_edge[k] = make_float4((float)A, (float)B, (float)(Cc + eps), (float)(Cc - safe));
    }
    dv.n_port_edge = (okq && port.n_edges <= 16) ? port.n_edges : 0;
    for (int k = 0; k < dv.n_port_edge; ++k)
      dv.port_edge[k] = make_float4((float)port.edge[k][0], (float)port.edge[k][1], (float)(port.edge[k][2] + eps),
                                    (float)(port.edge[k][2] - safe));
    const bool few_edges = c0.n_edges <= 8 && p0.n_edges <= 8;
    const bool disjoint = S == 1 || (dv.rect_crys.xhi < dv.slit_period + dv.rect_crys.xlo &&
                                     dv.rect_plasma.xhi < dv.slit_period + dv.rect_plasma.xlo);
    dv.interval_filter = (periodic && dv.slit_same_w && disjoint && few_edges && !dv.filter_disabled) ? 1 : 0;
    dv.certain_off = (!okc_slits || dv.physics_flags != 0 || dv.n_shields > 0) ? 1 : 0;
    if (d.filter_flags & (COM_FILTER_NO_RUN_CULL | COM_FILTER_NO_INTERVALS)) dv.interval_filter = 0;
  }
  const V3 O = load(d.origin);
  auto unit_normal = [&](const ComAperture& ap) {
    V3 N = load(ap.normal);
    return (1.0 / norm(N)) * N;
  };
  // value(p') = F.(c' - p')  ->  g = -F, h = F.c'
  auto put_at = [&](float* f, V3 F, V3 cl) {
    f[0] = (float)-F.x, f[1] = (float)-F.y, f[2] = (float)-F.z, f[3] = (float)dot(F, cl);
  };
  // detector forms are read by the lane that owns the crystal point: form-major keeps the loads coalesced;
  // slit forms are gathered by arbitrary lanes: crystal-major keeps one point's six forms in 96 contiguous bytes
  auto put = [&](std::vector<float>& tab, int form, int c, V3 F, V3 cl) {
    put_at(&tab[((size_t)form * Cp + c) * 4], F, cl);
  };
  auto put_slit = [&](int form, int c, V3 F, V3 cl) { put_at(&hs.slit_forms[((size_t)c * 6 + form) * 4], F, cl); };
  for (int c = 0; c < Cp; ++c) {
    if (c >= C) {  // padding lanes: |Xs| = 1e30 can never be <= |W| = 0
      hs.det_forms[((size_t)0 * Cp + c) * 4 + 3] = 1e30f;
      continue;
    }
    const V3 cp = load(&hs.crystal_xyz[3 * (size_t)c]), n = load(&hs.crystal_normal[3 * (size_t)c]);
    const V3 cl = cp - O;
    auto mirror = [&](V3 v) { return v - (2.0 * dot(v, n)) * n; };
    {  // detector seen in the crystal mirror: the line voxel -> crystal point, continued, meets the mirrored
       // detector plane at in-plane coordinates x = X'/W, y = Y'/W (homogeneous, linear in the voxel position)
      const V3 nd = unit_normal(det), q = load(det.p1) - cp;
      const double a = dot(nd, q), cx0 = dot(load(det.e1), q), cy0 = dot(load(det.e2), q);
      const V3 Nm = mirror(nd), e1m = mirror(load(det.e1)), e2m = mirror(load(det.e2));
      put(hs.det_forms, 0, c, (1.0 / hw) * (a * e1m - (cx0 + xc) * Nm), cl);
      put(hs.det_forms, 1, c, (1.0 / hh) * (a * e2m - (cy0 + yc) * Nm), cl);
      put(hs.det_forms, 2, c, Nm, cl);
    }
    for (int side = 0; side < 2; ++side) {
      const ComAperture& ap = hs.apertures[side];
      const V3 nh = unit_normal(ap), q = load(ap.p1) - cp;
      const double a = dot(nh, q), cx0 = dot(load(ap.e1), q), cy0 = dot(load(ap.e2), q);
      put_slit(3 * side + 0, c, a * load(ap.e1) - cx0 * nh, cl);
      put_slit(3 * side + 1, c, a * load(ap.e2) - cy0 * nh, cl);
      put_slit(3 * side + 2, c, nh, cl);
    }
    {  // port plane, same construction: hit of the line voxel -> crystal point at x = X'/W, y = Y'/W (mm, port frame)
      const V3 nh = unit_normal(port), q = load(port.p1) - cp;
      const double a = dot(nh, q), cx0 = dot(load(port.e1), q), cy0 = dot(load(port.e2), q);
      put_at(&hs.port_forms[((size_t)c * 3 + 0) * 4], a * load(port.e1) - cx0 * nh, cl);
      put_at(&hs.port_forms[((size_t)c * 3 + 1) * 4], a * load(port.e2) - cy0 * nh, cl);
      put_at(&hs.port_forms[((size_t)c * 3 + 2) * 4], nh, cl);
    }
  }
  return COM_OK;
}

}  // namespace com

extern "C" {

int com_aperture_from_slab(const double* verts_h, int32_t n, const double* offset_h, ComAperture* out_h) {
  return com::aperture_from_slab(verts_h, n, offset_h, out_h);
}

int com_aperture_contains(const ComAperture* ap_h, const double* point_h) {
  if (!ap_h || !point_h) return 0;
  return com::aperture_contains(ap_h, point_h);
}

const char* com_last_error(void) { return com::last_error(); }

const char* com_error_string(int code) {
  switch (code) {
    case COM_OK: return "ok";
    case COM_E_BADARG: return "bad argument";
    case COM_E_CUDA: return "CUDA error";
    case COM_E_NOMEM: return "out of memory / workspace too small";
    case COM_E_OVERFLOW: return "candidate list overflow";
    case COM_E_GEOMETRY: return "degenerate geometry";
    case COM_E_EMPTY: return "Not enough points to perform calculations!";
    case COM_E_IO: return "result file could not be opened or written";
    default: return "unknown error";
  }
}

int com_abi_version(void) { return COM_ABI_VERSION; }

}  // extern "C"
