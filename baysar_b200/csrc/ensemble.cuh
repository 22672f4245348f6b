// Affine-invariant ensemble sampler step (Goodman & Weare stretch move, red/blue split) on the device — the caller of
// the batched posterior for the 1M-walker workload (reference caller: emcee's EnsembleSampler driven from
// tulasa/fitting.py:1275-1304; emcee's StretchMove: z = ((a - 1) u + 1)^2 / a, proposal = c + z (x - c),
// accept with probability min(1, z^(n-1) p(proposal) / p(x))).  Walkers, proposals and log-posteriors never leave HBM.
// Random numbers: Philox4x32-10, counter = (global walker index: 64 bits, step: low 32 bits, draw + 2 * high 32 bits of
// step), key = seed — reproducible for any grid shape; `walker_offset` is the global index of this launch's first walker,
// so ranks that share a seed still draw distinct streams.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ens {

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

// uniform in the open interval (0, 1): 52 random bits from two 32-bit words, (bits + 1/2) / 2^52 is exact in float64
__host__ __device__ inline double u01(uint32_t hi, uint32_t lo) {
  const uint64_t bits = (((uint64_t)hi << 32) | lo) >> 12;
  return ((double)bits + 0.5) * (1.0 / 4503599627370496.0);
}

// proposal[s] = c + z (x_s - c), c a uniformly drawn walker of the complementary ensemble; log_factor = (n - 1) ln z
__global__ void stretch_propose_kernel(const double* __restrict__ walkers, const double* __restrict__ complement, int n_params,
                                       int64_t S, int64_t C, double a, uint64_t seed, uint64_t step, int64_t walker_offset,
                                       double* __restrict__ proposal, double* __restrict__ log_factor) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const uint64_t g = (uint64_t)(s + walker_offset);
  const Philox4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)step, 2u * (uint32_t)(step >> 32),
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
  const double u = u01(r.x, r.y);
  const double t = (a - 1.0) * u + 1.0;
  const double z = t * t / a;
  const int64_t pick = (int64_t)(u01(r.z, r.w) * (double)C);
  const double* x = walkers + s * n_params;
  const double* c = complement + (pick < C ? pick : C - 1) * n_params;
  double* p = proposal + s * n_params;
  for (int k = 0; k < n_params; ++k) p[k] = c[k] + z * (x[k] - c[k]);
  log_factor[s] = (double)(n_params - 1) * log(z);
}

// accept / reject in place; a proposal whose evaluation failed (status != 0, NaN or +inf logP) is rejected
__global__ void stretch_accept_kernel(double* __restrict__ walkers, double* __restrict__ logp, const double* __restrict__ proposal,
                                      const double* __restrict__ logp_new, const double* __restrict__ log_factor,
                                      const uint32_t* __restrict__ status_new, int n_params, int64_t S, uint64_t seed,
                                      uint64_t step, int64_t walker_offset, unsigned long long* __restrict__ n_accepted) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const uint64_t g = (uint64_t)(s + walker_offset);
  const Philox4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)step, 1u + 2u * (uint32_t)(step >> 32),
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
  const double lnu = log(u01(r.x, r.y));
  const double lp = logp_new[s];
  const bool ok = (status_new == nullptr || status_new[s] == 0u) && lp == lp && lp < 1.7976931348623157e308;
  if (ok && lnu < log_factor[s] + lp - logp[s]) {
    double* x = walkers + s * n_params;
    const double* p = proposal + s * n_params;
    for (int k = 0; k < n_params; ++k) x[k] = p[k];
    logp[s] = lp;
    if (n_accepted) atomicAdd(n_accepted, 1ull);
  }
}

}  // namespace ens
