// emul.cpp — C entry points that run warp-level device helpers of b2_kernel.cuh on one emulated warp (32 fibers).
#define B2_HOST_EMULATION
#include "warp_emul.h"

#include <functional>
#include <vector>

#include "../../mjlab_b200/csrc/b2_kernel.cuh"
#include "../../mjlab_b200/csrc/b2_tables.h"

namespace {
void run_warp(std::function<void(int)> fn) {
  warp_emul::ctx().w[0].bar = warp_emul::Bar{0, 32, 0};
  warp_emul::run_fibers(32, [&]() { fn(warp_emul::lane()); });
}
}  // namespace

extern "C" {
// schedules exactly as b2_create builds them; returns the number of sparse words
int emul_schedules(int nv, const int* dof_parentid, unsigned* dense, unsigned* sparse, int* start) {
  std::vector<unsigned> d, s;
  b2_build_ldl_schedules(nv, dof_parentid, d, s, start);
  if (dense) memcpy(dense, d.data(), d.size() * 4);
  if (sparse) memcpy(sparse, s.data(), s.size() * 4);
  return (int)s.size();
}
// factor the packed matrix A in place (device routine ldl_factor) and solve for x (ldl_solve)
void emul_ldl(int n, const int* dof_parentid, float* A, float* invdiag, float* x, int use_sparse) {
  std::vector<unsigned> d, s;
  int start[18];
  b2_build_ldl_schedules(n, dof_parentid, d, s, start);
  s.push_back(0);  // never empty
  run_warp([&](int lane) {
    b2::ldl_factor(A, invdiag, n, s.data(), start, d.data(), use_sparse != 0, lane);
    b2::ldl_solve(A, invdiag, x, n, lane);
  });
}
void emul_symv(int n, const float* M, const float* x, float* y) {
  run_warp([&](int lane) { b2::symv(M, x, y, n, lane); });
}
// narrowphase primitives (no warp intrinsics inside: called on the calling thread).  Poses are 12 floats
// (pos[3], row-major mat[9]); out: 7 floats per contact = dist, pos[3], normal[3].
static int prim_out(const b2::RawCon* c, int n, float* out) {
  for (int i = 0; i < n; i++) {
    out[7 * i] = c[i].dist;
    for (int k = 0; k < 3; k++) { out[7 * i + 1 + k] = c[i].pos[k]; out[7 * i + 4 + k] = c[i].n[k]; }
  }
  return n;
}
int emul_sphere_box(const float* sp, float r, const float* box, const float* h, float margin, float* out) {
  b2::RawCon c[1];
  return prim_out(c, b2::sphere_box(c[0], margin, sp, r, box, h), out);
}
int emul_capsule_box(const float* cap, const float* size, const float* box, const float* h, float margin, float* out) {
  b2::RawCon c[2];
  return prim_out(c, b2::capsule_box(c, margin, cap, size, box, h), out);
}
int emul_box_box(const float* b1, const float* h1, const float* b2_, const float* h2, float* out) {
  b2::RawCon c[8];
  return prim_out(c, b2::box_box(c, 0.f, b1, h1, b2_, h2), out);
}
// J x for contact rows (4 pyramid rows per contact, written to CJV0..3) and limit rows (LJV): device mulJ on
// caller-built SoA blocks.  layout[] returns the enum values the caller needs to fill them.
void emul_layout(int* out) {
  int v[] = {CS0, CMU, CINFO, CJV0, C_NFIELD, LINFO, LJV, L_NFIELD, SD};
  memcpy(out, v, sizeof(v));
}
void emul_mulJ(const float* x, float* con, float* lim, const int* gstart, const float* cdof,
               const unsigned long long* dofmask, int ncon, int nlim, int ngroup, int MC, int NLC) {
  std::vector<float> gV(6 * (size_t)MC + 6, 0.f);
  run_warp([&](int lane) {
    b2::mulJ(x, CJV0, LJV, false, con, lim, gstart, gV.data(), cdof, dofmask, ncon, nlim, ngroup, MC, NLC, lane);
  });
}
float emul_wsum(const float* v) {
  float out[32];
  run_warp([&](int lane) { out[lane] = b2::wsum(v[lane]); });
  for (int l = 1; l < 32; l++) if (out[l] != out[0]) return NAN;  // every lane must hold the same bits
  return out[0];
}
}
