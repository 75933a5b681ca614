// b2_convex.h — convex narrowphase for mesh and height-field geoms ("heightfield/mesh
// collision (GJK/EPA or MPR)"; reference call sites: terrains/heightfield_terrains.py builds <hfield> assets,
// utils/spec_config.py:245-276 sets the collision bits of arbitrary robot geoms, meshes included).
//
// MuJoCo collides meshes through their convex hull with a GJK/EPA pair routine (engine_collision_gjk.c; libccd
// MPR before 3.2) and height fields as one triangular prism per half cell against the other geom's convex routine
// (engine_collision_convex.c: mjc_ConvexHField); plane-mesh picks vertices in the support plane
// (mjc_PlaneConvex).  Those sources are not in /root/reference, so — like the box routines of b2_kernel.cuh — the
// rules below are OWN DEFINITIONS of the same geometric quantities (closest points / minimum translation of the
// convex hulls, one contact per convex pair as with MuJoCo's default `multiccd` off):
//   * a geom is a convex CORE (point, segment, box, vertex set, prism) inflated by a radius; separated cores give
//     dist = |closest points| - r1 - r2 (GJK), overlapping cores give dist = -(penetration depth) - r1 - r2 (EPA on
//     the Minkowski difference); normal from geom 1 to geom 2, position midway between the two surface points;
//   * cylinders and ellipsoids are smooth convex cores (radius 0) with closed-form support points; against a plane
//     they use MuJoCo's primitives (mjc_PlaneCylinder: the deepest rim point, the rim point at the other end of the
//     same generator, and two more points of the near cap at +-120 degrees; mjc_PlaneEllipsoid: the support point
//     along the plane normal) [UPSTREAM-MEMORY];
//   * plane - mesh: up to four vertices within 1e-3 of the lowest one: the lowest, the one farthest from it, the
//     one farthest from that line, and the farthest one on the other side of the line;
//   * height field: cells under the geom's bounding sphere, two prisms per cell (diagonal (r,c)-(r+1,c+1)), top
//     faces at the sample heights, bottom at -size[3]; at most B2C_MAXOUT contacts per geom pair, deepest kept.
// SINGLE SOURCE: this file is compiled as fp32 device code into libb2sim.so and — the same text with B2C_REAL =
// double — into the CPU oracle (oracle/b2_oracle.c includes it), so oracle-vs-kernel parity checks the fp32
// arithmetic and the integration, not the algorithm.  The algorithm itself is pinned independently by
// tests/test_convex.py: depth / distance / normal of random polytope pairs against the exact values obtained from the
// convex hull of the Minkowski difference (scipy.spatial.ConvexHull), and statics of resting bodies; cylinders and
// ellipsoids by tests/test_smooth_shapes.py against -min_u [h_A(u) + h_B(-u)] from closed-form support functions.
#pragma once

#ifndef B2C_REAL
#error "define B2C_REAL (float | double), B2C_FN / B2C_INL (function qualifiers) and B2C_SQRT before including b2_convex.h"
#endif

#define B2C_POINT 0     // sphere core
#define B2C_SEGMENT 1   // capsule core: pos +- axis_z * size[1]
#define B2C_BOX 2       // half sizes size[0..2]
#define B2C_VERTS 3     // mesh: nvert local-frame vertices
#define B2C_PRISM 4     // 6 world-frame points in pts
#define B2C_CYLINDER 5  // radius size[0], half height size[1] along local z
#define B2C_ELLIPSOID 6 // semi-axes size[0..2]
#define B2C_MAXOUT 8
#define B2C_GJK_ITER 40
#define B2C_EPA_ITER 24
#define B2C_EPA_NV (4 + B2C_EPA_ITER)
#define B2C_EPA_NF (4 + 2 * B2C_EPA_ITER + 8)

typedef B2C_REAL b2c_real;

typedef struct {
  int type, nvert;
  const b2c_real* pos;   // world position [3] (unused by B2C_PRISM)
  const b2c_real* mat;   // row-major world rotation [9]
  const b2c_real* size;  // geom_size
  const b2c_real* vert;  // mesh vertices, 3 per vertex, geom frame
  b2c_real pts[18];      // prism corners, world frame
} B2CShape;

typedef struct { b2c_real dist, pos[3], n[3]; } B2CCon;
typedef struct { b2c_real w[3], a[3]; } B2CVert;  // Minkowski-difference point w = a - b and its witness on shape A

B2C_INL b2c_real b2c_dot(const b2c_real* a, const b2c_real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
B2C_INL void b2c_cross(b2c_real* r, const b2c_real* a, const b2c_real* b) {
  b2c_real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
B2C_INL b2c_real b2c_eps(void) { return sizeof(b2c_real) == 4 ? (b2c_real)1e-6 : (b2c_real)1e-12; }

// farthest point of the core in world direction d
B2C_INL void b2c_support(const B2CShape* s, const b2c_real* d, b2c_real* out) {
  if (s->type == B2C_PRISM) {
    // corners 3..5 lie straight below corners 0..2 along the field's z axis: the sign of d along that axis picks
    // the triangle (top on a tie, as a scan of the six corners in order would)
    const b2c_real* R = s->mat;
    const b2c_real* q = s->pts + ((R[2] * d[0] + R[5] * d[1] + R[8] * d[2]) >= 0 ? 0 : 9);
    const b2c_real v0 = b2c_dot(q, d), v1 = b2c_dot(q + 3, d), v2 = b2c_dot(q + 6, d);
    const int best = v1 > v0 ? (v2 > v1 ? 6 : 3) : (v2 > v0 ? 6 : 0);
    out[0] = q[best]; out[1] = q[best + 1]; out[2] = q[best + 2];
    return;
  }
  const b2c_real* p = s->pos; const b2c_real* R = s->mat;
  if (s->type == B2C_POINT) { out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; return; }
  if (s->type == B2C_SEGMENT) {
    b2c_real ax[3] = {R[2], R[5], R[8]};
    b2c_real h = b2c_dot(ax, d) >= 0 ? s->size[1] : -s->size[1];
    out[0] = p[0] + ax[0] * h; out[1] = p[1] + ax[1] * h; out[2] = p[2] + ax[2] * h;
    return;
  }
  b2c_real dl[3] = {R[0] * d[0] + R[3] * d[1] + R[6] * d[2], R[1] * d[0] + R[4] * d[1] + R[7] * d[2],
                    R[2] * d[0] + R[5] * d[1] + R[8] * d[2]};
  b2c_real v[3];
  if (s->type == B2C_BOX) {
    for (int k = 0; k < 3; k++) v[k] = dl[k] >= 0 ? s->size[k] : -s->size[k];
  } else if (s->type == B2C_CYLINDER) {
    const b2c_real rho = B2C_SQRT(dl[0] * dl[0] + dl[1] * dl[1]);
    const b2c_real k = rho > b2c_eps() * B2C_SQRT(b2c_dot(dl, dl)) ? s->size[0] / rho : 0;  // (along the axis: the cap's centre)
    v[0] = k * dl[0]; v[1] = k * dl[1]; v[2] = dl[2] >= 0 ? s->size[1] : -s->size[1];
  } else if (s->type == B2C_ELLIPSOID) {
    const b2c_real a = s->size[0] * dl[0], b = s->size[1] * dl[1], c = s->size[2] * dl[2];
    const b2c_real nn = B2C_SQRT(a * a + b * b + c * c), k = nn > 0 ? 1 / nn : 0;
    v[0] = s->size[0] * a * k; v[1] = s->size[1] * b * k; v[2] = s->size[2] * c * k;
  } else {  // B2C_VERTS
    int best = 0;
    b2c_real bv = b2c_dot(s->vert, dl);
    for (int i = 1; i < s->nvert; i++) {
      b2c_real t = b2c_dot(s->vert + 3 * i, dl);
      if (t > bv) { bv = t; best = i; }
    }
    v[0] = s->vert[3 * best]; v[1] = s->vert[3 * best + 1]; v[2] = s->vert[3 * best + 2];
  }
  out[0] = p[0] + R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  out[1] = p[1] + R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  out[2] = p[2] + R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
}
B2C_FN void b2c_support_diff(const B2CShape* A, const B2CShape* B, const b2c_real* d, B2CVert* o) {
  b2c_real nd[3] = {-d[0], -d[1], -d[2]}, b[3];
  b2c_support(A, d, o->a);
  b2c_support(B, nd, b);
  o->w[0] = o->a[0] - b[0]; o->w[1] = o->a[1] - b[1]; o->w[2] = o->a[2] - b[2];
}

// Closest point of triangle (a, b, c) to the origin: barycentric weights l[3] (zero weight = vertex not needed).
B2C_FN void b2c_closest_tri(const b2c_real* a, const b2c_real* b, const b2c_real* c, b2c_real* l) {
  b2c_real ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
  b2c_real d1 = -b2c_dot(ab, a), d2 = -b2c_dot(ac, a);
  l[0] = l[1] = l[2] = 0;
  if (d1 <= 0 && d2 <= 0) { l[0] = 1; return; }
  b2c_real d3 = -b2c_dot(ab, b), d4 = -b2c_dot(ac, b);
  if (d3 >= 0 && d4 <= d3) { l[1] = 1; return; }
  b2c_real vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { b2c_real v = d1 / (d1 - d3); l[0] = 1 - v; l[1] = v; return; }
  b2c_real d5 = -b2c_dot(ab, c), d6 = -b2c_dot(ac, c);
  if (d6 >= 0 && d5 <= d6) { l[2] = 1; return; }
  b2c_real vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { b2c_real w = d2 / (d2 - d6); l[0] = 1 - w; l[2] = w; return; }
  b2c_real va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    b2c_real w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    l[1] = 1 - w; l[2] = w;
    return;
  }
  b2c_real den = 1 / (va + vb + vc);
  l[1] = vb * den; l[2] = vc * den; l[0] = 1 - l[1] - l[2];
}

// Closest point of the simplex (n = 1..4 vertices) to the origin.  Writes weights, removes unused vertices
// (order preserved), returns the new size; `v` receives the closest point.  Returns 4 with v = 0 when a
// tetrahedron contains the origin.
B2C_FN int b2c_simplex(B2CVert* S, int n, b2c_real* lam, b2c_real* v) {
  b2c_real l[4] = {0, 0, 0, 0};
  if (n == 1) l[0] = 1;
  else if (n == 2) {
    b2c_real ab[3] = {S[1].w[0] - S[0].w[0], S[1].w[1] - S[0].w[1], S[1].w[2] - S[0].w[2]};
    b2c_real den = b2c_dot(ab, ab), t = den > 0 ? -b2c_dot(S[0].w, ab) / den : 0;
    t = t < 0 ? 0 : (t > 1 ? 1 : t);
    l[0] = 1 - t; l[1] = t;
  } else if (n == 3) b2c_closest_tri(S[0].w, S[1].w, S[2].w, l);
  else {
    // faces (i, j, k) opposite vertex o; the origin is outside a face when it lies on the other side than o
    const int F[4][4] = {{0, 1, 2, 3}, {0, 1, 3, 2}, {0, 2, 3, 1}, {1, 2, 3, 0}};
    b2c_real best = -1;
    int inside = 1;
    for (int f = 0; f < 4; f++) {
      const b2c_real *a = S[F[f][0]].w, *b = S[F[f][1]].w, *c = S[F[f][2]].w, *o = S[F[f][3]].w;
      b2c_real ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]}, nn[3];
      b2c_cross(nn, ab, ac);
      b2c_real so = (o[0] - a[0]) * nn[0] + (o[1] - a[1]) * nn[1] + (o[2] - a[2]) * nn[2];
      b2c_real sp = -b2c_dot(a, nn);
      // degenerate (flat) tetrahedron: every face is treated as "outside" so the result is the closest point
      // of the flat figure
      int out = (so * so <= b2c_eps() * b2c_eps() * b2c_dot(nn, nn) * b2c_dot(ab, ab)) ? 1 : (sp * so < 0);
      if (!out) continue;
      inside = 0;
      b2c_real t[3];
      b2c_closest_tri(a, b, c, t);
      b2c_real p[3];
      for (int k = 0; k < 3; k++) p[k] = t[0] * a[k] + t[1] * b[k] + t[2] * c[k];
      b2c_real d2 = b2c_dot(p, p);
      if (best < 0 || d2 < best) {
        best = d2;
        l[0] = l[1] = l[2] = l[3] = 0;
        l[F[f][0]] = t[0]; l[F[f][1]] = t[1]; l[F[f][2]] = t[2];
      }
    }
    if (inside) {
      v[0] = v[1] = v[2] = 0;
      lam[0] = lam[1] = lam[2] = lam[3] = (b2c_real)0.25;
      return 4;
    }
  }
  int m = 0;
  v[0] = v[1] = v[2] = 0;
  for (int i = 0; i < n; i++) {
    if (l[i] <= 0) continue;
    if (m != i) S[m] = S[i];
    lam[m] = l[i];
    for (int k = 0; k < 3; k++) v[k] += l[i] * S[m].w[k];
    m++;
  }
  return m;
}

// Grow a simplex that touches the origin (n < 4) into a tetrahedron for EPA.  Returns 0 when the Minkowski
// difference is flat (no volume): the caller reports a touching contact.
B2C_FN int b2c_blowup(const B2CShape* A, const B2CShape* B, B2CVert* S, int n, b2c_real scale) {
  const b2c_real tiny = (sizeof(b2c_real) == 4 ? (b2c_real)1e-5 : (b2c_real)1e-10) * scale;
  if (n == 1) {
    for (int k = 0; k < 6 && n == 1; k++) {
      b2c_real d[3] = {0, 0, 0};
      d[k >> 1] = (k & 1) ? -1 : 1;
      b2c_support_diff(A, B, d, &S[1]);
      b2c_real e[3] = {S[1].w[0] - S[0].w[0], S[1].w[1] - S[0].w[1], S[1].w[2] - S[0].w[2]};
      if (b2c_dot(e, e) > tiny * tiny) n = 2;
    }
    if (n == 1) return 0;
  }
  if (n == 2) {
    b2c_real e[3] = {S[1].w[0] - S[0].w[0], S[1].w[1] - S[0].w[1], S[1].w[2] - S[0].w[2]};
    b2c_real ax[3] = {0, 0, 0};
    int k0 = (e[0] * e[0] <= e[1] * e[1] && e[0] * e[0] <= e[2] * e[2]) ? 0 : (e[1] * e[1] <= e[2] * e[2] ? 1 : 2);
    ax[k0] = 1;
    b2c_real d0[3], d1[3];
    b2c_cross(d0, e, ax);
    b2c_cross(d1, e, d0);
    b2c_real n0 = B2C_SQRT(b2c_dot(d0, d0)), n1 = B2C_SQRT(b2c_dot(d1, d1));
    for (int k = 0; k < 3; k++) { d0[k] /= n0; d1[k] /= n1; }
    const b2c_real cs[6] = {1, (b2c_real)0.5, (b2c_real)-0.5, -1, (b2c_real)-0.5, (b2c_real)0.5};
    const b2c_real sn[6] = {0, (b2c_real)0.8660254, (b2c_real)0.8660254, 0, (b2c_real)-0.8660254, (b2c_real)-0.8660254};
    for (int k = 0; k < 6 && n == 2; k++) {
      b2c_real d[3] = {cs[k] * d0[0] + sn[k] * d1[0], cs[k] * d0[1] + sn[k] * d1[1], cs[k] * d0[2] + sn[k] * d1[2]};
      b2c_support_diff(A, B, d, &S[2]);
      b2c_real f[3] = {S[2].w[0] - S[0].w[0], S[2].w[1] - S[0].w[1], S[2].w[2] - S[0].w[2]}, c[3];
      b2c_cross(c, e, f);
      if (b2c_dot(c, c) > tiny * tiny * b2c_dot(e, e)) n = 3;
    }
    if (n == 2) return 0;
  }
  if (n == 3) {
    b2c_real e[3] = {S[1].w[0] - S[0].w[0], S[1].w[1] - S[0].w[1], S[1].w[2] - S[0].w[2]};
    b2c_real f[3] = {S[2].w[0] - S[0].w[0], S[2].w[1] - S[0].w[1], S[2].w[2] - S[0].w[2]}, nn[3];
    b2c_cross(nn, e, f);
    b2c_real nl = B2C_SQRT(b2c_dot(nn, nn));
    for (int sgn = 0; sgn < 2 && n == 3; sgn++) {
      b2c_real d[3] = {sgn ? -nn[0] : nn[0], sgn ? -nn[1] : nn[1], sgn ? -nn[2] : nn[2]};
      b2c_support_diff(A, B, d, &S[3]);
      b2c_real g[3] = {S[3].w[0] - S[0].w[0], S[3].w[1] - S[0].w[1], S[3].w[2] - S[0].w[2]};
      b2c_real vol = b2c_dot(g, nn);
      if (vol * vol > tiny * tiny * nl * nl) n = 4;
    }
    if (n == 3) return 0;
  }
  return 1;
}

typedef struct { unsigned char i[3]; b2c_real n[3], d; } B2CFace;

B2C_FN int b2c_face_plane(const B2CVert* V, B2CFace* f) {
  const b2c_real *a = V[f->i[0]].w, *b = V[f->i[1]].w, *c = V[f->i[2]].w;
  b2c_real ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
  b2c_cross(f->n, ab, ac);
  b2c_real l2 = b2c_dot(f->n, f->n);
  if (l2 <= 0) { f->d = (b2c_real)1e30; return 0; }
  b2c_real inv = 1 / B2C_SQRT(l2);
  f->n[0] *= inv; f->n[1] *= inv; f->n[2] *= inv;
  f->d = b2c_dot(f->n, a);
  return 1;
}

// Penetration depth of overlapping cores: expanding polytope on the Minkowski difference, started from the
// tetrahedron S that contains the origin.  Returns the depth (>= 0), the normal n (from A to B) and the witness
// points pa, pb.
B2C_FN b2c_real b2c_epa(const B2CShape* A, const B2CShape* B, const B2CVert* S, b2c_real scale, b2c_real* n,
                        b2c_real* pa, b2c_real* pb) {
  B2CVert V[B2C_EPA_NV];
  B2CFace F[B2C_EPA_NF];
  unsigned char E[2 * B2C_EPA_NF];
  int nv = 4, nf = 4;
  for (int i = 0; i < 4; i++) V[i] = S[i];
  const unsigned char T[4][4] = {{0, 1, 2, 3}, {0, 3, 1, 2}, {0, 2, 3, 1}, {1, 3, 2, 0}};  // face + opposite vertex
  for (int f = 0; f < 4; f++) {
    F[f].i[0] = T[f][0]; F[f].i[1] = T[f][1]; F[f].i[2] = T[f][2];
    b2c_face_plane(V, &F[f]);
    const b2c_real *a = V[T[f][0]].w, *o = V[T[f][3]].w;
    if (F[f].n[0] * (o[0] - a[0]) + F[f].n[1] * (o[1] - a[1]) + F[f].n[2] * (o[2] - a[2]) > 0) {  // outward = away from the 4th vertex
      unsigned char t = F[f].i[1]; F[f].i[1] = F[f].i[2]; F[f].i[2] = t;
      b2c_face_plane(V, &F[f]);
    }
  }
  const b2c_real tol = (sizeof(b2c_real) == 4 ? (b2c_real)2e-6 : (b2c_real)1e-11) * scale;
  int best = 0;
  B2CFace good = F[0];  // closest face of the last consistent polytope
  b2c_real lastd = -(b2c_real)1e30;
  for (int it = 0;; it++) {
    best = 0;
    for (int f = 1; f < nf; f++) if (F[f].d < F[best].d) best = f;
    // The closest face can only move outwards as the polytope grows.  When it comes back in, the last expansion
    // corrupted the polytope (a triangle built across a nearly flat region of the horizon came out flipped - curved
    // shapes in fp32 produce such regions near convergence): the answer is the closest face before it.
    if (F[best].d < lastd - 4 * tol) { F[0] = good; nf = 1; best = 0; break; }
    good = F[best];
    lastd = F[best].d;
    if (it >= B2C_EPA_ITER || nv >= B2C_EPA_NV || nf + 8 > B2C_EPA_NF) break;
    B2CVert w;
    b2c_support_diff(A, B, F[best].n, &w);
    if (b2c_dot(F[best].n, w.w) - F[best].d <= tol) break;
    // faces that see the new point; their boundary (edges not shared by two visible faces) is the horizon
    int ne = 0, keep = 0;
    for (int f = 0; f < nf; f++) {
      const b2c_real* a = V[F[f].i[0]].w;
      b2c_real side = F[f].n[0] * (w.w[0] - a[0]) + F[f].n[1] * (w.w[1] - a[1]) + F[f].n[2] * (w.w[2] - a[2]);
      // (a face the new point is coplanar with - within rounding, an eighth of the margin by which the best face
      // sees it - stays: deciding such faces by the sign of a rounding error lets the horizon run through a flat
      // facet, and the triangle built on an edge inside it comes out flipped; rim points of a cylinder cap are all
      // coplanar)
      if (F[f].d < (b2c_real)1e29 && side <= tol * (b2c_real)0.125) { if (keep != f) F[keep] = F[f]; keep++; continue; }
      for (int e = 0; e < 3; e++) {
        unsigned char p = F[f].i[e], q = F[f].i[(e + 1) % 3];
        int found = -1;
        for (int k = 0; k < ne; k++) if (E[2 * k] == q && E[2 * k + 1] == p) { found = k; break; }
        if (found >= 0) { E[2 * found] = E[2 * (ne - 1)]; E[2 * found + 1] = E[2 * (ne - 1) + 1]; ne--; }
        else if (ne < B2C_EPA_NF) { E[2 * ne] = p; E[2 * ne + 1] = q; ne++; }
      }
    }
    if (keep == nf || ne < 3 || keep + ne > B2C_EPA_NF) break;  // numerical trouble: keep the current best face
    nf = keep;
    V[nv] = w;
    for (int k = 0; k < ne; k++) {
      F[nf].i[0] = E[2 * k]; F[nf].i[1] = E[2 * k + 1]; F[nf].i[2] = (unsigned char)nv;
      b2c_face_plane(V, &F[nf]);
      nf++;
    }
    nv++;
  }
  if (F[best].d > (b2c_real)1e29) {  // (cannot happen with a proper start simplex)
    n[0] = 0; n[1] = 0; n[2] = 1;
    for (int k = 0; k < 3; k++) { pa[k] = S[0].a[k]; pb[k] = S[0].a[k] - S[0].w[k]; }
    return 0;
  }
  // Witness points: barycentric coordinates of the origin's projection on the closest face.  A flat facet of the
  // Minkowski difference is covered by several coplanar triangles with the same plane distance: the one that
  // contains the projection is the one whose closest POINT is nearest.
  {
    const b2c_real dsel = F[best].d + 4 * tol;
    b2c_real q2 = -1;
    int sel = best;
    for (int f = 0; f < nf; f++) {
      if (F[f].d > dsel) continue;
      b2c_real t[3], q[3];
      b2c_closest_tri(V[F[f].i[0]].w, V[F[f].i[1]].w, V[F[f].i[2]].w, t);
      for (int k = 0; k < 3; k++) q[k] = t[0] * V[F[f].i[0]].w[k] + t[1] * V[F[f].i[1]].w[k] + t[2] * V[F[f].i[2]].w[k];
      b2c_real d2 = b2c_dot(q, q);
      if (q2 < 0 || d2 < q2) { q2 = d2; sel = f; }
    }
    best = sel;
  }
  const B2CVert *a = &V[F[best].i[0]], *b = &V[F[best].i[1]], *c = &V[F[best].i[2]];
  b2c_real l[3];
  b2c_closest_tri(a->w, b->w, c->w, l);
  for (int k = 0; k < 3; k++) {
    pa[k] = l[0] * a->a[k] + l[1] * b->a[k] + l[2] * c->a[k];
    pb[k] = pa[k] - (l[0] * a->w[k] + l[1] * b->w[k] + l[2] * c->w[k]);
    n[k] = F[best].n[k];
  }
  return F[best].d > 0 ? F[best].d : 0;
}

// One contact between two inflated convex cores (A, ra) and (B, rb); normal from A to B.  `scale` = a length of
// the order of the geoms' size (tolerances are relative to it).  Returns 0 or 1.
B2C_FN int b2c_pair(B2CCon* out, b2c_real margin, const B2CShape* A, b2c_real ra, const B2CShape* B, b2c_real rb,
                    const b2c_real* ca, const b2c_real* cb, b2c_real scale) {
  const b2c_real cutoff = margin + ra + rb;
  B2CVert S[4];
  b2c_real lam[4] = {1, 0, 0, 0}, v[3], d[3] = {ca[0] - cb[0], ca[1] - cb[1], ca[2] - cb[2]};
  if (b2c_dot(d, d) <= 0) { d[0] = 1; d[1] = 0; d[2] = 0; }
  b2c_support_diff(A, B, d, &S[0]);
  int n = 1, pen = 0;
  v[0] = S[0].w[0]; v[1] = S[0].w[1]; v[2] = S[0].w[2];
  const b2c_real rel = sizeof(b2c_real) == 4 ? (b2c_real)1e-6 : (b2c_real)1e-12;
  const b2c_real abs2 = (sizeof(b2c_real) == 4 ? (b2c_real)1e-10 : (b2c_real)1e-22) * scale * scale;
  for (int it = 0; it < B2C_GJK_ITER; it++) {
    b2c_real vv = b2c_dot(v, v);
    if (vv <= abs2) { pen = 1; break; }
    b2c_real nd[3] = {-v[0], -v[1], -v[2]};
    B2CVert w;
    b2c_support_diff(A, B, nd, &w);
    b2c_real vw = b2c_dot(v, w.w);
    if (vw > 0 && vw * vw > cutoff * cutoff * vv) return 0;  // separating plane beyond the cutoff
    if (vv - vw <= rel * vv) break;                          // no progress: v is the closest point
    int dup = 0;
    for (int i = 0; i < n; i++) {
      b2c_real e[3] = {w.w[0] - S[i].w[0], w.w[1] - S[i].w[1], w.w[2] - S[i].w[2]};
      if (b2c_dot(e, e) <= abs2) dup = 1;
    }
    if (dup) break;
    S[n++] = w;
    b2c_real vold = vv;
    n = b2c_simplex(S, n, lam, v);
    if (n == 4 && v[0] == 0 && v[1] == 0 && v[2] == 0) { pen = 1; break; }
    if (b2c_dot(v, v) >= vold) break;  // rounding: no decrease
  }
  b2c_real nrm[3], pa[3], pb[3], dist;
  if (!pen) {
    b2c_real dc = B2C_SQRT(b2c_dot(v, v));
    if (dc - ra - rb > margin) return 0;
    pa[0] = pa[1] = pa[2] = 0;
    for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) pa[k] += lam[i] * S[i].a[k];
    for (int k = 0; k < 3; k++) { pb[k] = pa[k] - v[k]; nrm[k] = -v[k] / dc; }
    dist = dc - ra - rb;
  } else {
    if (n < 4 && !b2c_blowup(A, B, S, n, scale)) {
      // flat Minkowski difference touching the origin: zero-depth contact along the centre line
      b2c_real l = B2C_SQRT(b2c_dot(d, d));
      for (int k = 0; k < 3; k++) { nrm[k] = -d[k] / l; pa[k] = S[0].a[k]; pb[k] = S[0].a[k] - S[0].w[k]; }
      dist = -ra - rb;
    } else {
      b2c_real depth = b2c_epa(A, B, S, scale, nrm, pa, pb);
      dist = -depth - ra - rb;
    }
    if (dist > margin) return 0;
  }
  out->dist = dist;
  for (int k = 0; k < 3; k++) {
    out->n[k] = nrm[k];
    out->pos[k] = (b2c_real)0.5 * ((pa[k] + nrm[k] * ra) + (pb[k] - nrm[k] * rb));
  }
  return 1;
}

// plane (pos pp, normal pn) against a mesh: up to 4 contacts
B2C_FN int b2c_plane_mesh(B2CCon* out, b2c_real margin, const b2c_real* pp, const b2c_real* pn, const B2CShape* M) {
  const b2c_real* R = M->mat; const b2c_real* p = M->pos;
  // plane normal in the mesh frame, and the plane offset along it
  b2c_real nl[3] = {R[0] * pn[0] + R[3] * pn[1] + R[6] * pn[2], R[1] * pn[0] + R[4] * pn[1] + R[7] * pn[2],
                    R[2] * pn[0] + R[5] * pn[1] + R[8] * pn[2]};
  b2c_real off = (p[0] - pp[0]) * pn[0] + (p[1] - pp[1]) * pn[1] + (p[2] - pp[2]) * pn[2];  // height of the mesh origin
  int ia = 0;
  b2c_real lo = b2c_dot(M->vert, nl);
  for (int i = 1; i < M->nvert; i++) { b2c_real h = b2c_dot(M->vert + 3 * i, nl); if (h < lo) { lo = h; ia = i; } }
  if (off + lo > margin) return 0;
  const b2c_real thr = lo + (b2c_real)1e-3;
  int idx[4] = {ia, -1, -1, -1};
  const b2c_real* va = M->vert + 3 * ia;
  b2c_real bd = 0;
  for (int i = 0; i < M->nvert; i++) {
    const b2c_real* vi = M->vert + 3 * i;
    if (b2c_dot(vi, nl) > thr) continue;
    b2c_real e[3] = {vi[0] - va[0], vi[1] - va[1], vi[2] - va[2]}, d2 = b2c_dot(e, e);
    if (d2 > bd) { bd = d2; idx[1] = i; }
  }
  if (idx[1] >= 0 && bd > (b2c_real)1e-12) {
    const b2c_real* vb = M->vert + 3 * idx[1];
    b2c_real ab[3] = {vb[0] - va[0], vb[1] - va[1], vb[2] - va[2]}, side[3];
    b2c_cross(side, nl, ab);  // in-plane direction perpendicular to ab
    b2c_real bpos = 0, bneg = 0;
    for (int i = 0; i < M->nvert; i++) {
      const b2c_real* vi = M->vert + 3 * i;
      if (b2c_dot(vi, nl) > thr) continue;
      b2c_real s = (vi[0] - va[0]) * side[0] + (vi[1] - va[1]) * side[1] + (vi[2] - va[2]) * side[2];
      if (s > bpos) { bpos = s; idx[2] = i; }
      if (s < bneg) { bneg = s; idx[3] = i; }
    }
    const b2c_real lim = (b2c_real)1e-6 * bd;  // (side has length |ab|: s / |ab| is a distance)
    if (bpos * bpos <= lim * bd) idx[2] = -1;
    if (bneg * bneg <= lim * bd) idx[3] = -1;
  } else idx[1] = -1;
  int n = 0;
  for (int k = 0; k < 4; k++) {
    if (idx[k] < 0) continue;
    const b2c_real* v = M->vert + 3 * idx[k];
    b2c_real dist = off + b2c_dot(v, nl);
    if (dist > margin) continue;
    b2c_real wv[3] = {p[0] + R[0] * v[0] + R[1] * v[1] + R[2] * v[2], p[1] + R[3] * v[0] + R[4] * v[1] + R[5] * v[2],
                      p[2] + R[6] * v[0] + R[7] * v[1] + R[8] * v[2]};
    out[n].dist = dist;
    for (int j = 0; j < 3; j++) { out[n].n[j] = pn[j]; out[n].pos[j] = wv[j] - pn[j] * ((b2c_real)0.5 * dist); }
    n++;
  }
  return n;
}

// Height field H (pose hp/hR, size = [rx, ry, zmax, zbase], nrow x ncol samples in [0, 1], row-major with the
// row index along y) against an inflated convex core G.  gc / rbound: G's bounding sphere.  Normal from the
// height field to G.  At most B2C_MAXOUT contacts per pair, deepest kept, grid order otherwise.
//
// The pair is processed in stages so that a caller can spread the work of several pairs over SIMD lanes (the
// kernel does; b2c_hfield below runs them in sequence):
//   b2c_hfield_range   cells under the geom's footprint -> `count` items, item j = (row, column, triangle)
//   b2c_hfield_cull    cheap rejects of one item (height of the prism, plane of its top triangle)
//   b2c_hfield_prism   GJK / EPA of one prism against G
//   b2c_hfield_keep    merge one contact into the pair's list (items in ascending j), b2c_hfield_sort at the end
typedef struct { int r0, c0, ncc, count; b2c_real lo2; } B2CHfRange;

B2C_FN int b2c_hfield_range(B2CHfRange* rg_out, b2c_real margin, const b2c_real* hp, const b2c_real* hR,
                            const b2c_real* hsize, int nrow, int ncol, const B2CShape* G, b2c_real rg,
                            const b2c_real* gc, b2c_real rbound) {
  rg_out->count = 0;
  b2c_real dlt[3] = {gc[0] - hp[0], gc[1] - hp[1], gc[2] - hp[2]};
  b2c_real c[3] = {hR[0] * dlt[0] + hR[3] * dlt[1] + hR[6] * dlt[2], hR[1] * dlt[0] + hR[4] * dlt[1] + hR[7] * dlt[2],
                   hR[2] * dlt[0] + hR[5] * dlt[1] + hR[8] * dlt[2]};
  const b2c_real reach = rbound + margin;
  if (c[2] - reach > hsize[2] || c[2] + reach < -hsize[3]) return 0;
  if (c[0] - reach > hsize[0] || c[0] + reach < -hsize[0] || c[1] - reach > hsize[1] || c[1] + reach < -hsize[1]) return 0;
  // extent of the inflated geom along the field's axes (six support points): only the cells under its footprint
  // can touch it, and only prisms that rise to within `margin` of its lowest point
  b2c_real lo[3], hi[3];
  for (int k = 0; k < 3; k++) {
    b2c_real ax[3] = {hR[k], hR[3 + k], hR[6 + k]}, nax[3] = {-ax[0], -ax[1], -ax[2]}, p[3];
    b2c_support(G, ax, p);
    hi[k] = (p[0] - hp[0]) * ax[0] + (p[1] - hp[1]) * ax[1] + (p[2] - hp[2]) * ax[2] + rg + margin;
    b2c_support(G, nax, p);
    lo[k] = (p[0] - hp[0]) * ax[0] + (p[1] - hp[1]) * ax[1] + (p[2] - hp[2]) * ax[2] - rg - margin;
  }
  if (lo[2] > hsize[2] || hi[2] < -hsize[3] || lo[0] > hsize[0] || hi[0] < -hsize[0] || lo[1] > hsize[1] || hi[1] < -hsize[1]) return 0;
  const b2c_real dx = 2 * hsize[0] / (b2c_real)(ncol - 1), dy = 2 * hsize[1] / (b2c_real)(nrow - 1);
  int c0 = (int)((lo[0] + hsize[0]) / dx), c1 = (int)((hi[0] + hsize[0]) / dx);
  int r0 = (int)((lo[1] + hsize[1]) / dy), r1 = (int)((hi[1] + hsize[1]) / dy);
  if (lo[0] + hsize[0] < 0) c0 = 0;
  if (lo[1] + hsize[1] < 0) r0 = 0;
  if (c1 > ncol - 2) c1 = ncol - 2;
  if (r1 > nrow - 2) r1 = nrow - 2;
  if (c1 < c0 || r1 < r0) return 0;
  rg_out->r0 = r0; rg_out->c0 = c0; rg_out->ncc = c1 - c0 + 1; rg_out->lo2 = lo[2];
  rg_out->count = 2 * (r1 - r0 + 1) * (c1 - c0 + 1);
  return rg_out->count;
}

// Corners of the top triangle of item j in the field's frame.  t = 0: (r,c) (r,c+1) (r+1,c+1);  t = 1: (r,c) (r+1,c+1) (r+1,c)
B2C_INL void b2c_hfield_tri(const B2CHfRange* R, int j, const b2c_real* hsize, int nrow, int ncol, const b2c_real* hdata,
                            b2c_real* lx, b2c_real* ly, b2c_real* lz) {
  const int t = j & 1, cell = j >> 1, r = R->r0 + cell / R->ncc, cc = R->c0 + cell % R->ncc;
  const b2c_real dx = 2 * hsize[0] / (b2c_real)(ncol - 1), dy = 2 * hsize[1] / (b2c_real)(nrow - 1);
  const b2c_real z00 = hdata[r * ncol + cc] * hsize[2], z01 = hdata[r * ncol + cc + 1] * hsize[2];
  const b2c_real z10 = hdata[(r + 1) * ncol + cc] * hsize[2], z11 = hdata[(r + 1) * ncol + cc + 1] * hsize[2];
  const b2c_real x0 = -hsize[0] + dx * (b2c_real)cc, y0 = -hsize[1] + dy * (b2c_real)r;
  lx[0] = x0; lx[1] = x0 + dx; lx[2] = t ? x0 : x0 + dx;
  ly[0] = y0; ly[1] = t ? y0 + dy : y0; ly[2] = y0 + dy;
  lz[0] = z00; lz[1] = t ? z11 : z01; lz[2] = t ? z10 : z11;
}

// 1 when item j survives the cheap rejects and needs b2c_hfield_prism
B2C_FN int b2c_hfield_cull(const B2CHfRange* R, int j, b2c_real margin, const b2c_real* hp, const b2c_real* hR,
                           const b2c_real* hsize, int nrow, int ncol, const b2c_real* hdata, const B2CShape* G, b2c_real rg) {
  b2c_real lx[3], ly[3], lz[3];
  b2c_hfield_tri(R, j, hsize, nrow, ncol, hdata, lx, ly, lz);
  b2c_real zmax = lz[0] > lz[1] ? (lz[0] > lz[2] ? lz[0] : lz[2]) : (lz[1] > lz[2] ? lz[1] : lz[2]);
  if (R->lo2 > zmax) return 0;
  // the prism lies below the plane of its top triangle: a geom whose lowest point along that plane's normal
  // clears it by more than the margin cannot touch (rejects the prisms under a limb that hovers over a slope)
  b2c_real e1[3] = {lx[1] - lx[0], ly[1] - ly[0], lz[1] - lz[0]}, e2[3] = {lx[2] - lx[0], ly[2] - ly[0], lz[2] - lz[0]}, nl[3];
  b2c_cross(nl, e1, e2);
  if (nl[2] < 0) { nl[0] = -nl[0]; nl[1] = -nl[1]; nl[2] = -nl[2]; }
  b2c_real nw[3], dn[3], q[3];
  for (int k = 0; k < 3; k++) nw[k] = hR[3 * k] * nl[0] + hR[3 * k + 1] * nl[1] + hR[3 * k + 2] * nl[2];
  dn[0] = -nw[0]; dn[1] = -nw[1]; dn[2] = -nw[2];
  b2c_support(G, dn, q);
  b2c_real p0[3];
  for (int k = 0; k < 3; k++) p0[k] = hp[k] + hR[3 * k] * lx[0] + hR[3 * k + 1] * ly[0] + hR[3 * k + 2] * lz[0];
  const b2c_real nn = B2C_SQRT(b2c_dot(nw, nw));
  const b2c_real h = (q[0] - p0[0]) * nw[0] + (q[1] - p0[1]) * nw[1] + (q[2] - p0[2]) * nw[2];
  return h > (rg + margin) * nn ? 0 : 1;
}

// the prism under item j against G; 1 and *cn when they touch (within margin)
B2C_FN int b2c_hfield_prism(B2CCon* cn, const B2CHfRange* R, int j, b2c_real margin, const b2c_real* hp, const b2c_real* hR,
                            const b2c_real* hsize, int nrow, int ncol, const b2c_real* hdata, const B2CShape* G,
                            b2c_real rg, const b2c_real* gc, b2c_real rbound) {
  b2c_real lx[3], ly[3], lz[3];
  b2c_hfield_tri(R, j, hsize, nrow, ncol, hdata, lx, ly, lz);
  const b2c_real dx = 2 * hsize[0] / (b2c_real)(ncol - 1), dy = 2 * hsize[1] / (b2c_real)(nrow - 1);
  B2CShape P;
  P.type = B2C_PRISM; P.nvert = 6; P.pos = hp; P.mat = hR; P.size = hsize; P.vert = hdata;
  b2c_real pc[3] = {0, 0, 0};
  for (int i = 0; i < 6; i++) {
    b2c_real l[3] = {lx[i % 3], ly[i % 3], i < 3 ? lz[i] : -hsize[3]};
    for (int k = 0; k < 3; k++) {
      P.pts[3 * i + k] = hp[k] + hR[3 * k] * l[0] + hR[3 * k + 1] * l[1] + hR[3 * k + 2] * l[2];
      pc[k] += P.pts[3 * i + k] * ((b2c_real)1 / 6);
    }
  }
  return b2c_pair(cn, margin, &P, 0, G, rg, pc, gc, rbound + dx + dy);
}

// `tie` of b2c_hfield_keep: depths closer than this count as equal
B2C_INL b2c_real b2c_hfield_tie(const b2c_real* hsize, int nrow, int ncol, b2c_real rbound) {
  const b2c_real dx = 2 * hsize[0] / (b2c_real)(ncol - 1), dy = 2 * hsize[1] / (b2c_real)(nrow - 1);
  return (b2c_real)1e-5 * (rbound + dx + dy);
}

// contact cn of item j joins the pair's list (out / order / *n); items arrive in ascending j
B2C_INL void b2c_hfield_keep(B2CCon* out, int* order, int* n, const B2CCon* cn, int j, b2c_real tie) {
  if (*n < B2C_MAXOUT) { order[*n] = j; out[(*n)++] = *cn; return; }
  // evict the shallowest; neighbouring prisms often report the same contact, so depths within `tie` count
  // as equal and the later prism goes (the choice must not hinge on the last bit of a depth)
  b2c_real dmax = out[0].dist;
  for (int i = 1; i < *n; i++) if (out[i].dist > dmax) dmax = out[i].dist;
  int worst = -1;
  for (int i = 0; i < *n; i++) if (out[i].dist >= dmax - tie && (worst < 0 || order[i] > order[worst])) worst = i;
  if (cn->dist < dmax - tie) { out[worst] = *cn; order[worst] = j; }
}

// back to grid order whatever was evicted (insertion sort, n <= 8)
B2C_INL void b2c_hfield_sort(B2CCon* out, int* order, int n) {
  for (int a = 1; a < n; a++) {
    B2CCon cv = out[a];
    int ov = order[a], b = a - 1;
    while (b >= 0 && order[b] > ov) { out[b + 1] = out[b]; order[b + 1] = order[b]; b--; }
    out[b + 1] = cv; order[b + 1] = ov;
  }
}

B2C_FN int b2c_hfield(B2CCon* out, b2c_real margin, const b2c_real* hp, const b2c_real* hR, const b2c_real* hsize,
                      int nrow, int ncol, const b2c_real* hdata, const B2CShape* G, b2c_real rg,
                      const b2c_real* gc, b2c_real rbound) {
  B2CHfRange R;
  if (!b2c_hfield_range(&R, margin, hp, hR, hsize, nrow, ncol, G, rg, gc, rbound)) return 0;
  int n = 0, order[B2C_MAXOUT];
  const b2c_real tie = b2c_hfield_tie(hsize, nrow, ncol, rbound);
  for (int j = 0; j < R.count; j++) {
    B2CCon cn;
    if (!b2c_hfield_cull(&R, j, margin, hp, hR, hsize, nrow, ncol, hdata, G, rg)) continue;
    if (!b2c_hfield_prism(&cn, &R, j, margin, hp, hR, hsize, nrow, ncol, hdata, G, rg, gc, rbound)) continue;
    b2c_hfield_keep(out, order, &n, &cn, j, tie);
  }
  b2c_hfield_sort(out, order, n);
  return n;
}

// Shape of a geom of MuJoCo type `gtype` (2 sphere, 3 capsule, 6 box, 7 mesh); returns its inflation radius.
B2C_FN b2c_real b2c_shape(B2CShape* s, int gtype, const b2c_real* pos, const b2c_real* mat, const b2c_real* size,
                          const b2c_real* vert, int nvert) {
  s->pos = pos; s->mat = mat; s->size = size; s->vert = vert; s->nvert = nvert;
  if (gtype == 2) { s->type = B2C_POINT; return size[0]; }
  if (gtype == 3) { s->type = B2C_SEGMENT; return size[0]; }
  if (gtype == 6) { s->type = B2C_BOX; return 0; }
  if (gtype == 5) { s->type = B2C_CYLINDER; return 0; }
  if (gtype == 4) { s->type = B2C_ELLIPSOID; return 0; }
  s->type = B2C_VERTS;
  return 0;
}

// plane (pos pp, normal pn) against a cylinder: up to 4 contacts (mjc_PlaneCylinder)
B2C_FN int b2c_plane_cylinder(B2CCon* out, b2c_real margin, const b2c_real* pp, const b2c_real* pn, const b2c_real* pos,
                              const b2c_real* mat, const b2c_real* size) {
  b2c_real axis[3] = {mat[2], mat[5], mat[8]};
  b2c_real prjaxis = b2c_dot(pn, axis);
  if (prjaxis > 0) { axis[0] = -axis[0]; axis[1] = -axis[1]; axis[2] = -axis[2]; prjaxis = -prjaxis; }  // axis towards the plane
  const b2c_real dist0 = (pos[0] - pp[0]) * pn[0] + (pos[1] - pp[1]) * pn[1] + (pos[2] - pp[2]) * pn[2];
  // radial direction towards the plane: -normal without its component along the axis
  b2c_real vec[3] = {axis[0] * prjaxis - pn[0], axis[1] * prjaxis - pn[1], axis[2] * prjaxis - pn[2]};
  const b2c_real len2 = b2c_dot(vec, vec);
  if (len2 >= (b2c_real)1e-10) {  // (MuJoCo: mjMINVAL^2; 1e-10 = an angle of 1e-5 rad, above fp32 rounding of a flat cap)
    const b2c_real k = size[0] / B2C_SQRT(len2);
    vec[0] *= k; vec[1] *= k; vec[2] *= k;
  } else {  // cap parallel to the plane: the cylinder's x axis
    vec[0] = mat[0] * size[0]; vec[1] = mat[3] * size[0]; vec[2] = mat[6] * size[0];
  }
  const b2c_real prjvec = b2c_dot(vec, pn);
  axis[0] *= size[1]; axis[1] *= size[1]; axis[2] *= size[1];
  prjaxis *= size[1];
  int n = 0;
  if (dist0 + prjaxis + prjvec > margin) return 0;
  out[n].dist = dist0 + prjaxis + prjvec;
  for (int k = 0; k < 3; k++) { out[n].pos[k] = pos[k] + vec[k] + axis[k] - pn[k] * out[n].dist * (b2c_real)0.5; out[n].n[k] = pn[k]; }
  n++;
  if (dist0 - prjaxis + prjvec <= margin) {
    out[n].dist = dist0 - prjaxis + prjvec;
    for (int k = 0; k < 3; k++) { out[n].pos[k] = pos[k] + vec[k] - axis[k] - pn[k] * out[n].dist * (b2c_real)0.5; out[n].n[k] = pn[k]; }
    n++;
  }
  const b2c_real prjvec1 = -prjvec * (b2c_real)0.5;
  if (dist0 + prjaxis + prjvec1 <= margin) {
    b2c_real vec1[3];
    b2c_cross(vec1, vec, axis);
    const b2c_real l1 = B2C_SQRT(b2c_dot(vec1, vec1));
    if (l1 > 0) {
      const b2c_real k = size[0] * (b2c_real)0.8660254037844386 / l1;
      const b2c_real d = dist0 + prjaxis + prjvec1;
      for (int sgn = 0; sgn < 2; sgn++) {
        out[n].dist = d;
        for (int c = 0; c < 3; c++) {
          out[n].pos[c] = pos[c] + (sgn ? -k : k) * vec1[c] + axis[c] - vec[c] * (b2c_real)0.5 - pn[c] * d * (b2c_real)0.5;
          out[n].n[c] = pn[c];
        }
        n++;
      }
    }
  }
  return n;
}

// plane against an ellipsoid: the support point along -normal (mjc_PlaneEllipsoid)
B2C_FN int b2c_plane_ellipsoid(B2CCon* out, b2c_real margin, const b2c_real* pp, const b2c_real* pn, const B2CShape* E) {
  const b2c_real nd[3] = {-pn[0], -pn[1], -pn[2]};
  b2c_real q[3];
  b2c_support(E, nd, q);
  const b2c_real dist = (q[0] - pp[0]) * pn[0] + (q[1] - pp[1]) * pn[1] + (q[2] - pp[2]) * pn[2];
  if (dist > margin) return 0;
  out[0].dist = dist;
  for (int k = 0; k < 3; k++) { out[0].pos[k] = q[k] - pn[k] * dist * (b2c_real)0.5; out[0].n[k] = pn[k]; }
  return 1;
}
