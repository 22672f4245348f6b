// Issue-rate probe of kind::tf32 tcgen05.mma (measurement hook, no reference counterpart): one CTA per SM, one elected
// thread issues `iters` k-blocks of the contraction kernels' MMA pattern (two 8-deep k-steps x {hi*hi, hi*lo, lo*hi})
// back to back on constant operands and reports cycles per MMA.  Used to separate the tensor pipe's own pace from
// the feeding pipelines of if_contract.cuh / pix_contract.cuh.
#pragma once
#include "if_contract.cuh"

namespace mmaprobe {

__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}

// form: 0 = A from tensor memory (.ts), 1 = A from shared memory (.ss)
// pattern: 0 = the kernels' pattern (hi*hi accumulators rotate per k-block, one cross-term accumulator),
//          1 = every MMA into one accumulator, 2 = one MMA per k-step only (hi*hi)
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int tn, int form, int n_acc_hi, int b_stages, int pattern,
                                                           int iters, long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* b_ring = base;                       // b_stages x tn x 128 B
  unsigned char* a_tile = base + (size_t)4 * 256 * 128;  // 128 x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (4 * 256 * 128 + 128 * 128) / 4; i += 128) reinterpret_cast<float*>(base)[i] = 1.0f + (float)(i & 7) * 0.125f;
  if (tid == 0) {
    ifc::mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ifc::smem_u32(&tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = tmem_slot;
  long long cycles = -1;
  if (warp == 0) {
    const uint32_t idesc = ifc::make_idesc(tn);
    const uint32_t a_cols = (uint32_t)((n_acc_hi + 1) * tn);  // A stages behind the accumulators (when they fit)
    const uint64_t adesc = ifc::make_desc(ifc::smem_u32(a_tile));
    if (ifc::elect_one()) {
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint64_t d_bhi = ifc::make_desc(ifc::smem_u32(b_ring + (size_t)(it % b_stages) * tn * 128));
        const uint64_t d_blo = d_bhi + 4;
        const uint32_t acc0 = tmem_d + (uint32_t)((pattern == 1 ? 0 : it % n_acc_hi) * tn);
        const uint32_t acc1 = pattern == 1 ? acc0 : tmem_d + (uint32_t)(n_acc_hi * tn);
        const uint32_t a_hi = tmem_d + (a_cols + 32u * (uint32_t)(it & 1) <= 480u ? a_cols + 32u * (uint32_t)(it & 1) : 0u);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t adv = (uint64_t)(2 * ks);
          if (form == 0) {
            ifc::mma_tf32_ts(acc0, a_hi + 8 * ks, d_bhi + adv, idesc, 1u);
            if (pattern != 2) {
              ifc::mma_tf32_ts(acc1, a_hi + 8 * ks, d_blo + adv, idesc, 1u);
              ifc::mma_tf32_ts(acc1, a_hi + 16 + 8 * ks, d_bhi + adv, idesc, 1u);
            }
          } else {
            mma_tf32_ss(acc0, adesc + adv, d_bhi + adv, idesc, 1u);
            if (pattern != 2) {
              mma_tf32_ss(acc1, adesc + adv, d_blo + adv, idesc, 1u);
              mma_tf32_ss(acc1, adesc + 4 + adv, d_bhi + adv, idesc, 1u);
            }
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ifc::smem_u32(&bar)) : "memory");
      // bounded wait: a probe must never hang the device
      uint32_t done = 0;
      const long long t_lim = clock64() + 4000000000LL;
      while (!done && clock64() < t_lim) {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(done) : "r"(ifc::smem_u32(&bar)), "r"(0u) : "memory");
      }
      cycles = done ? clock64() - t0 : -2;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(512));
    if (cycles != -1) out[blockIdx.x] = cycles;
  }
}

}  // namespace mmaprobe
