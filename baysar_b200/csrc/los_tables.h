// Tables the LOS stage derives from the model's interpolation axes at model creation (host side, plain C++; also used by
// the CPU emulation of the kernels in tests/emu).  None of it changes what is computed:
//   * reciprocals of the grid spacings / knot differences, so that the per-point cell weights and the Cox-de Boor
//     recurrence multiply instead of divide (an FP64 division is ~25 dependent instructions);
//   * bucket tables over the leading bits of a double (sign, exponent, three mantissa bits: eight buckets per octave)
//     that give a search a starting index at most a step or two below the answer, replacing five-step bisections whose
//     every step is a dependent load.
#pragma once
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "device_common.cuh"

namespace los_tables {

inline int key_of(double x) {
  int64_t b;
  std::memcpy(&b, &x, 8);
  return (int)((int32_t)(b >> 32) >> BAYSAR_KEY_SHIFT);
}
inline double value_of_key(int key) {  // smallest double with this key
  const int64_t b = (int64_t)((uint64_t)(uint32_t)(key << BAYSAR_KEY_SHIFT) << 32);
  double x;
  std::memcpy(&x, &b, 8);
  return x;
}

struct Host {
  std::vector<double> d;  // all double tables, 16-byte aligned segments
  std::vector<int> i;     // all bucket tables
  size_t recip_off[AX_COUNT];
  GridKey key[AX_COUNT];
};

// upper = false: table of lower_bound (first index with g[idx] >= edge); true: upper_bound (first index with g[idx] > edge)
inline void add_axis(Host& h, int ax, const double* g, int n, bool spline) {
  // reciprocals
  while (h.d.size() % 2) h.d.push_back(0.0);
  h.recip_off[ax] = h.d.size();
  if (!spline) {
    for (int k = 0; k + 1 < n; ++k) h.d.push_back(1.0 / (g[k + 1] - g[k]));
    if (n < 2) h.d.push_back(0.0);
  } else {
    for (int j = 1; j <= 3; ++j)
      for (int k = 0; k < n; ++k) {
        const double den = k + j < n ? g[k + j] - g[k] : 0.0;
        h.d.push_back(den != 0.0 ? 1.0 / den : 0.0);  // zero differences (repeated end knots) are never used
      }
  }
  // bucket table
  GridKey gk{(int)h.i.size(), 0, 1, 0};
  bool ok = n >= 1;
  for (int k = 0; k < n; ++k) ok = ok && g[k] > 0.0 && std::isfinite(g[k]) && (k == 0 || g[k] >= g[k - 1]);
  if (ok && key_of(g[n - 1]) - key_of(g[0]) < 4096) {
    gk.key0 = key_of(g[0]) - 1;
    gk.nb = key_of(g[n - 1]) + 1 - gk.key0 + 1;
    h.i.push_back(0);  // bucket 0: everything below the axis
    for (int b = 1; b < gk.nb; ++b) {
      const double edge = value_of_key(gk.key0 + b);
      int idx = 0;
      if (!spline) while (idx < n && g[idx] < edge) ++idx;
      else while (idx < n && g[idx] <= edge) ++idx;
      h.i.push_back(idx);
    }
  } else {
    h.i.push_back(0);  // degenerate axis: a single bucket, the search scans forward from 0
  }
  h.key[ax] = gk;
}

inline Host build(const int64_t* I, const double* a11_ne, const double* a11_te, const double* h_ne, const double* h_te,
                  const double* spl_tx, const double* spl_ty) {
  Host h;
  add_axis(h, AX_A11_NE, a11_ne, (int)I[I_A11_NNE], false);
  add_axis(h, AX_A11_TE, a11_te, (int)I[I_A11_NTE], false);
  add_axis(h, AX_H_NE, h_ne, (int)I[I_H_NNE], false);
  add_axis(h, AX_H_TE, h_te, (int)I[I_H_NTE], false);
  add_axis(h, AX_SPL_X, spl_tx, (int)I[I_SPL_NX], true);
  add_axis(h, AX_SPL_Y, spl_ty, (int)I[I_SPL_NY], true);
  return h;
}

// view over buffers holding Host::d / Host::i (device copies, or the host vectors themselves in the emulator)
inline LosTablesDev view(const Host& h, const double* d_base, const int* i_base) {
  LosTablesDev t;
  for (int a = 0; a < AX_COUNT; ++a) { t.recip[a] = d_base + h.recip_off[a]; t.key[a] = h.key[a]; }
  t.key_tbl = i_base;
  return t;
}

}  // namespace los_tables
