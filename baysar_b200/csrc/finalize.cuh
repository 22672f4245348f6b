// K3 — logP = sum of posterior_components in list order (reference baysar/posterior.py:228) and the output checks
// of check_call_output (posterior.py:249-258), one thread per theta.
#pragma once
#include "device_common.cuh"

__global__ void finalize_kernel(const double* comps, int ncols, int64_t n, double* logp, uint32_t* status) {
  pdl_trigger();
  pdl_wait();  // the components of the chord / pixel kernels
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* c = comps + i * ncols;
  double s = 0.0;
  for (int k = 0; k < ncols; ++k) s = s + c[k];
  uint32_t st = 0;
  if (s != s) st |= 1u << 13;            // BAYSAR_ST_LOGP_NAN
  else if (s - s != 0.0) st |= 1u << 14; // BAYSAR_ST_LOGP_INF
  else if (s >= 0.0) st |= 1u << 15;     // BAYSAR_ST_LOGP_POSITIVE
  logp[i] = s;
  if (st) status[i] |= st;
}
