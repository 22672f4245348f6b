// TEST INFRASTRUCTURE ONLY — never part of the product.
//
// A minimal execution model that lets the CUDA kernel sources in baysar_b200/csrc/*.cuh be compiled with g++ and run
// on CPU threads, so that kernel logic (indexing, barriers, reductions) can be checked against the golden fixtures
// in the GPU-less build container before GPU minutes are spent.  Every CUDA thread of a block is a std::thread;
// __syncthreads / warp shuffles are std::barrier phases.  Blocks run one after another.  It is slow (seconds per
// block) and is only ever built by tests/test_emulated_kernels.py into a scratch directory.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(x) alignas(x)
#define BAYSAR_DEV inline

struct EmuDim3 { unsigned x = 1, y = 1, z = 1; };
struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct float2 { float x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }

namespace emu {
struct Block {
  unsigned nthreads;
  std::barrier<> bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<uint64_t> slots;     // one 8-byte exchange slot per thread
  std::vector<int> preds;
  std::vector<unsigned char> smem;
  explicit Block(unsigned nt, size_t smem_bytes) : nthreads(nt), bar(nt), slots(nt), preds(nt), smem(smem_bytes + 64) {
    for (unsigned w = 0; w < (nt + 31) / 32; ++w) {
      unsigned cnt = (w + 1) * 32 <= nt ? 32 : nt - w * 32;
      warp_bar.emplace_back(new std::barrier<>(cnt));
    }
  }
};
inline thread_local Block* cur = nullptr;
}  // namespace emu

inline thread_local EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

inline unsigned char* emu_dynamic_smem() {
  uintptr_t p = reinterpret_cast<uintptr_t>(emu::cur->smem.data());
  return reinterpret_cast<unsigned char*>((p + 15) & ~uintptr_t(15));
}

inline void __syncthreads() { emu::cur->bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::cur->warp_bar[threadIdx.x >> 5]->arrive_and_wait(); }
inline int __syncthreads_or(int pred) {
  emu::Block* b = emu::cur;
  b->preds[threadIdx.x] = pred;
  b->bar.arrive_and_wait();
  int r = 0;
  for (unsigned i = 0; i < b->nthreads; ++i) r |= (b->preds[i] != 0);
  b->bar.arrive_and_wait();
  return r;
}

template <typename T>
inline T emu_shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  emu::Block* b = emu::cur;
  const unsigned w0 = threadIdx.x & ~31u;
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  b->slots[threadIdx.x] = raw;
  b->warp_bar[threadIdx.x >> 5]->arrive_and_wait();
  T out = v;
  if (src_lane >= 0 && src_lane < 32 && w0 + src_lane < b->nthreads) {
    uint64_t o = b->slots[w0 + src_lane];
    std::memcpy(&out, &o, sizeof(T));
  }
  b->warp_bar[threadIdx.x >> 5]->arrive_and_wait();
  return out;
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_shfl(v, (int)((threadIdx.x & 31) ^ m)); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl(v, src & 31); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, int d) {
  const int lane = (int)(threadIdx.x & 31);
  return emu_shfl(v, lane + d < 32 ? lane + d : lane);
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, int d) {
  int lane = threadIdx.x & 31;
  T o = emu_shfl(v, lane - d >= 0 ? lane - d : lane);
  return o;
}
inline int __any_sync(unsigned, int pred) {
  int r = pred != 0;
  for (int o = 16; o > 0; o >>= 1) r |= __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

// atomics: threads of one block run concurrently -> use compiler builtins
inline int atomicMin(int* p, int v) { int o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
inline int atomicMax(int* p, int v) { int o = __atomic_load_n(p, __ATOMIC_RELAXED); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
inline uint32_t atomicOr(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }

inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __powf(float a, float b) { return std::pow(a, b); }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
using std::ceil; using std::exp; using std::fabs; using std::floor; using std::log; using std::log10; using std::pow;
using std::rint; using std::sqrt;

template <typename Kernel, typename... Args>
void emu_launch(Kernel kernel, EmuDim3 grid, unsigned nthreads, size_t smem_bytes, Args... args) {
  for (unsigned by = 0; by < grid.y; ++by) {
    for (unsigned bx = 0; bx < grid.x; ++bx) {
      emu::Block block(nthreads, smem_bytes);
      std::vector<std::thread> ts;
      ts.reserve(nthreads);
      for (unsigned t = 0; t < nthreads; ++t) {
        ts.emplace_back([&, t]() {
          emu::cur = &block;
          threadIdx.x = t; blockIdx.x = bx; blockIdx.y = by; blockDim.x = nthreads; gridDim.x = grid.x; gridDim.y = grid.y;
          kernel(args...);
        });
      }
      for (auto& th : ts) th.join();
    }
  }
}
