// b2_tables.h — host-side derived tables shared by the library (b2sim.cu) and the CPU emulation of the
// warp routines (tests/emul): plain C++, no CUDA.
#pragma once
#include <algorithm>
#include <vector>

// Trailing-update schedules of the bottom-up blocked L^T D L (b2_kernel.cuh: ldl_factor), one word per target
// entry: p | i << 12 | j << 18 with p = tri(i) + j.  `dense`: every packed index in order (a block with m
// leading rows takes the first tri(m) words); `sparse`: per block (pivots kt, kt-1, .., four at a time from
// the last dof) the entries whose row and column are both ancestors of one of the block's pivots - the only
// ones the dof tree lets the update touch; start[b] .. start[b+1] delimits block b (start has 18 slots).
inline void b2_build_ldl_schedules(int nv, const int* dof_parentid, std::vector<unsigned>& dense,
                                   std::vector<unsigned>& sparse, int* start) {
  std::vector<unsigned long long> danc(std::max(nv, 1), 0ull);  // strict dof ancestors
  for (int k = 0; k < nv; k++)
    for (int a = dof_parentid[k]; a >= 0; a = dof_parentid[a]) danc[k] |= 1ull << a;
  dense.clear();
  sparse.clear();
  for (int i = 0; i < nv; i++)
    for (int j = 0; j <= i; j++) dense.push_back((unsigned)(i * (i + 1) / 2 + j) | ((unsigned)i << 12) | ((unsigned)j << 18));
  int blk = 0;
  for (int kt = nv - 1; kt >= 0; kt -= 4, blk++) {
    int nbk = std::min(4, kt + 1), lead = kt - nbk + 1;
    start[blk] = (int)sparse.size();
    for (int i = 0; i < lead; i++)
      for (int j = 0; j <= i; j++) {
        bool hit = false;
        for (int t = 0; t < nbk; t++) {
          unsigned long long a = danc[kt - t];
          if ((a >> i & 1ull) && (a >> j & 1ull)) hit = true;
        }
        if (hit) sparse.push_back((unsigned)(i * (i + 1) / 2 + j) | ((unsigned)i << 12) | ((unsigned)j << 18));
      }
  }
  for (int b = blk; b < 18; b++) start[b] = (int)sparse.size();
}
