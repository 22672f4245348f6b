// K2b' — instrument function AND wavelength map as one tensor-core contraction (tcgen05 + TMEM), sm_100a only.
//
// For a chord with a static wavelength map and a non-negative instrument function, everything
// SpectrometerChord.forward_model does between the summed line shapes and the pixel spectrum is linear except the
// clip at the pre-convolution minimum (reference baysar/spectrometers.py:269-310):
//
//     conv   = fftconvolve(S, w, "same").clip(min S)        # the clip can only bite where the tap window leaves the grid
//     pixel_p = (1 - a_p) conv[i_p] + a_p conv[i_p + 1]      # interp1d onto the pixel wavelengths (i_p, a_p constant)
//
// so the (theta x fine grid) . (fine grid x pixel) product the north star asks for is taken literally:
//
//     C[theta, c] = sum_s S[theta, s] * B[s, c],   B[s, c] = (1 - a_c) w[i_c + o - s] + a_c w[i_c + 1 + o - s]
//
// with one column per pixel (the interpolation folded into B) plus the explicit fine-grid columns of the two edge
// zones (a_c = 0), where the pixel stage applies the clip and the trapz correction itself.  Compared with
// if_contract.cuh (every fine-grid column, N = F) this multiplies ~3.4x fewer elements on the Balmer demo.
//
// * M = 128 thetas, N = TN columns (64 or 128) per tile, K in blocks of 16 over the tile's own band
//   [k0, k0 + 16 n_kb) — tiles differ in length, so the (column tile, theta block) pairs are dealt to the CTAs by a
//   longest-first schedule computed on the host.
// * A operand: exactly the pipeline of if_contract.cuh (coalesced cp.async ring -> converter threads split each value
//   into TF32 hi + lo -> tensor memory; .ts MMAs).
// * B operand: no longer Toeplitz (the pixel pitch is not a whole number of fine-grid points), but still cheap to
//   GENERATE: four builder warps (each owning every fourth k-block) compute the TN x 16 tile from the zero-padded tap
//   table in shared memory (five LDS per four elements, one FMA and a three-operation TF32 split each) straight into a
//   K-major SWIZZLE_128B ring stage (hi parts in the first 64 bytes of a row, lo parts in the second), so B costs no
//   L2 / HBM traffic at all.
// * 3xTF32 with split accumulators: TMEM accumulation truncates, and here the area renormalisation no longer
//   cancels a common bias (the post-convolution area comes from an exact identity, see pixel_fold_kernel), so the
//   hi*hi products rotate over N_ACC_HI accumulators (4 at TN = 64, 2 at TN = 128: a column sees 1/4 or 1/2 of the
//   truncating additions per accumulator) and the 2^-11-sized cross terms get their own; the epilogue adds them.
// * Warp roles: 0-7 A converters, 8-11 epilogue, 12 MMA issue, 13-14 A loaders, 15-22 B builders (four teams of two).
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "if_contract.cuh"

namespace pxc {

using ifc::A_TILE_BYTES;
using ifc::RING_MAX;
using ifc::TILE_K;
using ifc::TILE_M;

constexpr int CONVERTER_WARPS = 8, EPILOGUE_WARPS = 4, BUILDER_WARPS = 8, BUILDER_TEAMS = 4;
constexpr int TSUM_SLOTS = 3;    // partial column sums per (row, tile): one per converter group
constexpr int LOADER_WARPS = 2;  // cp.async mode: each loader warp copies half of the rows of a k-block
constexpr int WARP_MMA = CONVERTER_WARPS + EPILOGUE_WARPS, WARP_LOADER = WARP_MMA + 1, WARP_BUILDER0 = WARP_LOADER + LOADER_WARPS;
constexpr int THREADS = (WARP_BUILDER0 + BUILDER_WARPS) * 32;

template <int TN>
struct Cfg {
#ifndef BAYSAR_PXC_NACC128
#define BAYSAR_PXC_NACC128 2
#endif
  static constexpr int N_ACC_HI = TN == 64 ? 4 : BAYSAR_PXC_NACC128;
  static constexpr int ACC_COLS = (N_ACC_HI + 1) * TN;        // 320 / 384 TMEM columns of accumulators
  static constexpr int A_STAGES = (512 - ACC_COLS) / 32;      // 6 / 4 tensor-memory stages of the split A operand
#ifndef BAYSAR_PXC_BSTAGES128
#define BAYSAR_PXC_BSTAGES128 6
#endif
  static constexpr int B_STAGES = TN == 64 ? 8 : BAYSAR_PXC_BSTAGES128;  // shared-memory stages of B tiles
  static constexpr int B_STAGE_BYTES = TN * 128;              // TN rows x (16 hi + 16 lo) floats
  static constexpr int STAGING_BYTES = TILE_M * TN * 4;
  static_assert(A_STAGES % 2 == 0, "the two converter groups take alternate stages");
  static_assert(TN % 64 == 0 && B_STAGES > BUILDER_TEAMS, "every builder team owns a stage in flight, the MMAs read another");
};
constexpr int MIN_KB = 4;  // every accumulator of a tile is initialised by its own first k-block
constexpr int MAX_FOLD_TILES = 64;  // folded (pixel) column tiles per chord: pixel_fold_kernel keeps one correction / scale per tile in shared memory

struct Params {
  const float* s;          // [rows][pitch] FP32 spectra, zero padded
  int64_t pitch;
  const int* tile_kcol0;   // [n_col_tiles] first A column of the tile's band (pad_left + k0, multiple of 16)
  const int* tile_nkb;     // [n_col_tiles] k-blocks in the band (>= MIN_KB)
  const int* col_idx;      // [n_col_tiles * TN]: B[k-block kb, element e][c] = lerp(wtab[x], wtab[x + 1], col_frac[c]),
  const float* col_frac;   //                     x = col_idx[c] - 16 kb - e
  const float* wtab;       // zero-padded instrument taps
  int wlen;                // floats, multiple of 4
  const int* sched;        // [ctas][sched_slots]: (column tile << 16) | theta block
  const int* sched_cnt;    // [ctas]
  int sched_slots;
  float* out;              // [rows][out_pitch], column tile ct at ct * TN
  int64_t out_pitch;
  int ring;                // shared-memory stages of raw A tiles
  unsigned sleep_epi, sleep_build;  // nanoseconds between polls of the long waits (0: spin)
  int use_tma;             // A loader: 1 = one TMA box per k-block, 0 = sixteen cp.async per lane (fallback / experiments)
  // column-sum identity (bias correction in pixel_fold_kernel): tsum[row][ct][grp] = this converter group's share of
  // sum_s S[row, s] * bbar[tile_boff[ct] + (s - k0)], bbar = sum over the tile's folded columns of B[s, c]
  const float* bbar;       // may be nullptr (no sums)
  const int* tile_boff;    // [n_col_tiles] offset of the tile's 16 n_kb floats in bbar, -1: none (explicit edge columns)
  double* tsum;            // [rows][n_col_tiles][TSUM_SLOTS]
  int n_col_tiles;
  long long* dbg_cycles;   // TRACE instantiation: [ctas][16] cycles spent waiting per role
  // B from a precomputed image (b_image_kernel, model creation): stage-sized, already swizzled and TF32-split tiles, one
  // per (column tile, position i in the tile's k-block sequence); nullptr: the builder warps generate B
  const unsigned char* bimg;
  const int* tile_bimg;    // [n_col_tiles] index of the tile's first stage image
};

template <int TN>
inline size_t smem_bytes(int ring, int wlen) {
  using C = Cfg<TN>;
  return (size_t)C::B_STAGES * C::B_STAGE_BYTES + C::STAGING_BYTES + (size_t)ring * A_TILE_BYTES + (size_t)wlen * 4 +
         (2 * C::A_STAGES + 2 * C::B_STAGES + 2 * RING_MAX + 2) * sizeof(uint64_t) + 16 + 1024;
}
template <int TN>
inline int ring_stages(int wlen) {
  const long room = (long)ifc::SMEM_LIMIT - (long)smem_bytes<TN>(0, wlen);
  const int r = (int)(room / A_TILE_BYTES);
  return r > RING_MAX ? RING_MAX : r;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

// Shared-memory accesses by 32-bit shared-window address: the carve-up below goes through an integer alignment, after
// which the compiler no longer knows the address space and would emit generic LD / ST with 64-bit address arithmetic.
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
// read of the tap table (constant after the prologue): not volatile, so that the compiler may batch the loads of
// several unrolled iterations ahead of the arithmetic
__device__ __forceinline__ float lds32_const(uint32_t addr) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d));
}
__device__ __forceinline__ void mbar_arrive32(uint32_t bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait32(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// wait with back-off: a warp that expects to wait long (the epilogue for a whole tile, a builder team that is several
// stages ahead) sleeps between polls instead of spending issue slots of its scheduler on the spin loop (28 % of the
// kernel's executed instructions were try_wait / branch pairs)
__device__ __forceinline__ void mbar_wait32_sleep(uint32_t bar, uint32_t parity, unsigned ns) {
  uint32_t done;
  for (;;) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(ns);
  }
}
template <bool TRACE>
__device__ __forceinline__ void mbar_wait32_timed(uint32_t bar, uint32_t parity, long long& acc) {
  if (TRACE) {
    const long long t0 = clock64();
    mbar_wait32(bar, parity);
    acc += clock64() - t0;
  } else {
    mbar_wait32(bar, parity);
  }
}
// v ~ hi + lo: hi = v rounded to TF32 (nearest, ties away: add half an ulp of the 10-bit mantissa, clear the low 13
// bits — three ALU operations where cvt.rna.tf32.f32 compiles to four per conversion on sm_100), lo = v - hi exactly;
// the tensor core ignores lo's low 13 bits, a truncation that is symmetric about zero because hi is rounded.
__device__ __forceinline__ void split_fast(float v, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}

// TRACE: per-role wait-cycle counters into p.dbg_cycles[cta][16] (measurement builds only)
// tmap: the zero-padded spectra as a 2-D float32 tensor {pitch, rows}, box {16, 128}, SWIZZLE_64B — the swizzle is
// exactly ifc::ring_off (16-byte chunk ^ ((row >> 1) & 3)), so one TMA instruction per k-block replaces the sixteen
// cp.async instructions per lane with which a single loader warp could not keep up (measured: ~790 cycles per k-block).
template <int TN, bool TRACE>
__global__ void __launch_bounds__(THREADS, 1) pix_contract_kernel(Params p, const __grid_constant__ CUtensorMap tmap) {
  using C = Cfg<TN>;
  extern __shared__ unsigned char smem_dyn[];
  // shared-window addresses (1024-byte aligned carve): B ring | epilogue staging | A ring | tap table | barriers | TMEM slot
  const uint32_t sm0 = (ifc::smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t bring = sm0;
  const uint32_t staging = bring + C::B_STAGES * C::B_STAGE_BYTES;
  const uint32_t ring = staging + C::STAGING_BYTES;
  const uint32_t wtab = ring + (uint32_t)p.ring * A_TILE_BYTES;
  const uint32_t full = wtab + (uint32_t)p.wlen * 4;       // [A_STAGES]
  const uint32_t empty = full + 8 * C::A_STAGES;           // [A_STAGES]
  const uint32_t bfull = empty + 8 * C::A_STAGES;          // [B_STAGES]
  const uint32_t bempty = bfull + 8 * C::B_STAGES;         // [B_STAGES]
  const uint32_t landed = bempty + 8 * C::B_STAGES;        // [RING_MAX]
  const uint32_t consumed = landed + 8 * RING_MAX;         // [RING_MAX]
  const uint32_t acc_full = consumed + 8 * RING_MAX;
  const uint32_t acc_free = acc_full + 8;
  const uint32_t tmem_slot = acc_free + 8;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long t_entry = TRACE ? clock64() : 0;
  // programmatic dependent launch: this prologue (tap table, barriers, tensor-memory allocation; model constants only)
  // overlaps with the tail of the spectral stage; the spectra are touched only after griddepcontrol.wait below
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int my_tiles = p.sched_cnt[blockIdx.x];
  const int* my_sched = p.sched + (int64_t)blockIdx.x * p.sched_slots;

  for (int i = tid; i < p.wlen; i += THREADS) asm volatile("st.shared.f32 [%0], %1;" ::"r"(wtab + 4 * i), "f"(p.wtab[i]) : "memory");
  if (tid == 0) {
    auto init = [](uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); };
    for (int s = 0; s < C::A_STAGES; ++s) { init(full + 8 * s, CONVERTER_WARPS / 2); init(empty + 8 * s, 1); }
    for (int s = 0; s < C::B_STAGES; ++s) { init(bfull + 8 * s, p.bimg ? 1 : BUILDER_WARPS / BUILDER_TEAMS); init(bempty + 8 * s, 1); }
    init(acc_full, 1);
    init(acc_free, EPILOGUE_WARPS);
    for (int r = 0; r < RING_MAX; ++r) { init(landed + 8 * r, p.use_tma ? 1 : 32 * LOADER_WARPS); init(consumed + 8 * r, CONVERTER_WARPS / 2); }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tmem_d;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_d) : "r"(tmem_slot));
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the previous kernel (the spectral stage) has completed
  if (TRACE && tid == 0) p.dbg_cycles[(int64_t)blockIdx.x * 16 + 13] = clock64() - t_entry;

  // With the B image the builder warps are free: four of them (16..19, tensor-memory lane quadrants 0..3 like every group
  // of four consecutive warps) form a third converter group -- the converters are latency-bound (wait for the ring, split,
  // wait for the tensor-memory stage, store) and set the pace of the kernel
  const bool third_group = p.bimg != nullptr && warp >= WARP_BUILDER0 + 1 && warp < WARP_BUILDER0 + 5;
  if (warp < CONVERTER_WARPS || third_group) {
    // ===================== A converters: ring (shared memory) -> TF32 split -> tensor memory =====================
    // (if_contract.cuh: group `grp` takes the k-block iterations grp, grp + n_grp, ... of this CTA's tiles, back to back)
    const int n_grp = p.bimg ? 3 : 2;
    const int quad = warp & 3, grp = third_group ? 2 : warp >> 2;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t lane_base = tmem_d + ((uint32_t)(quad * 32) << 16) + C::ACC_COLS;
    int r = grp % p.ring;
    uint32_t r_phase = (uint32_t)((grp / p.ring) & 1);
    int d_pending = -1;
    long long w0 = 0, w1 = 0;
    const long long t_begin = TRACE ? clock64() : 0;
    int64_t g0 = 0;  // index of the tile's first k-block in this CTA's sequence
    for (int q = 0; q < my_tiles; ++q) {
      const int ct = my_sched[q] >> 16, rb = my_sched[q] & 0xffff;
      const int n_kb = p.tile_nkb[ct];
      const int boff = p.bbar ? p.tile_boff[ct] : -1;
      double tacc = 0.0;
      for (int i = (int)(((grp - g0) % n_grp + n_grp) % n_grp); i < n_kb; i += n_grp) {
        const int64_t g = g0 + i;
        const int d = (int)(g % C::A_STAGES);
        float4 bb[4];
        if (boff >= 0) {  // warp-uniform addresses: one L1 transaction per load
          const float4* bp = reinterpret_cast<const float4*>(p.bbar + boff + TILE_K * ifc::kb_order(i, n_kb));
#pragma unroll
          for (int c = 0; c < 4; ++c) bb[c] = __ldg(bp + c);
        }
        mbar_wait32_timed<TRACE>(landed + 8 * r, r_phase, w0);
        float4 raw[4];
        const uint32_t row_ptr = ring + (uint32_t)r * A_TILE_BYTES;
#pragma unroll
        for (int c = 0; c < 4; ++c) raw[c] = lds128(row_ptr + ifc::ring_off(row_in_tile, c));
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          split_fast(raw[c].x, hi[4 * c + 0], lo[4 * c + 0]);
          split_fast(raw[c].y, hi[4 * c + 1], lo[4 * c + 1]);
          split_fast(raw[c].z, hi[4 * c + 2], lo[4 * c + 2]);
          split_fast(raw[c].w, hi[4 * c + 3], lo[4 * c + 3]);
        }
        // TMA loader: the stage is refilled through the async proxy, and nothing orders this warp's generic-proxy
        // reads before that write except a cross-proxy fence (without it single elements of a row came back with the
        // NEXT ring cycle's data in a few of a thousand launches: write-after-read across proxies)
        if (p.use_tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive32(consumed + 8 * r);
        if (d_pending >= 0) {  // the group's previous k-block is in tensor memory: hand it over
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;");
          __syncwarp();
          if (lane == 0) mbar_arrive32(full + 8 * d_pending);
        }
        if (g >= C::A_STAGES) {  // TMEM stage d: the MMAs of its previous user are done
          mbar_wait32_timed<TRACE>(empty + 8 * d, (uint32_t)(((g / C::A_STAGES) - 1) & 1), w1);
          asm volatile("tcgen05.fence::after_thread_sync;");
        }
        ifc::tmem_st16(lane_base + 32 * d, hi);
        ifc::tmem_st16(lane_base + 32 * d + TILE_K, lo);
        d_pending = d;
        r += n_grp;
        if (r >= p.ring) { r -= p.ring; r_phase ^= 1u; }
        if (boff >= 0) {  // sixteen float32 products per k-block, float64 across k-blocks
          float part = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            part = fmaf(raw[c].x, bb[c].x, part);
            part = fmaf(raw[c].y, bb[c].y, part);
            part = fmaf(raw[c].z, bb[c].z, part);
            part = fmaf(raw[c].w, bb[c].w, part);
          }
          tacc += (double)part;
        }
      }
      if (boff >= 0) {
        double* ts = p.tsum + ((int64_t)(rb * TILE_M + row_in_tile) * p.n_col_tiles + ct) * TSUM_SLOTS;
        ts[grp] = tacc;
        if (n_grp == 2 && grp == 0) ts[2] = 0.0;
      }
      g0 += n_kb;
    }
    if (d_pending >= 0) {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) mbar_arrive32(full + 8 * d_pending);
    }
    if (TRACE && tid == 0) {
      long long* o = p.dbg_cycles + (int64_t)blockIdx.x * 16;
      o[0] = w0; o[1] = w1; o[2] = clock64() - t_begin;
    }
  } else if (warp >= WARP_LOADER && warp < WARP_BUILDER0) {
    // ===================== A loader: one TMA box (128 rows x 16 floats, SWIZZLE_64B) per k-block =====================
    int r = 0;
    uint32_t r_phase = 0;
    int64_t g = 0;
    long long w0 = 0;
    if (p.use_tma && lane == 0 && warp == WARP_LOADER) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
    for (int q = 0; q < my_tiles; ++q) {
      const int ct = my_sched[q] >> 16, rb = my_sched[q] & 0xffff;
      const int n_kb = p.tile_nkb[ct];
      const int col0 = p.tile_kcol0[ct], row0 = rb * TILE_M;
      for (int i = 0; i < n_kb; ++i, ++g) {
        if (g >= p.ring) mbar_wait32_timed<TRACE>(consumed + 8 * r, r_phase ^ 1u, w0);
        if (!p.use_tma) {  // loader warp lw: rows 64 lw .. 64 lw + 63, lane = (row & 7, 16-byte chunk)
          const int lw = warp - WARP_LOADER;
          const float* src = p.s + (int64_t)(row0 + 64 * lw + (lane >> 2)) * p.pitch + col0 + 4 * (lane & 3) + TILE_K * ifc::kb_order(i, n_kb);
          const uint32_t dst = ring + (uint32_t)r * A_TILE_BYTES;
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) {
            const int row = 64 * lw + (lane >> 2) + 8 * qq;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + ifc::ring_off(row, lane & 3)),
                         "l"(src + (int64_t)(8 * qq) * p.pitch) : "memory");
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(landed + 8 * r) : "memory");
        } else if (lane == 0 && warp == WARP_LOADER) {
          const uint32_t bar = landed + 8 * r, dst = ring + (uint32_t)r * A_TILE_BYTES;
          const int col = col0 + TILE_K * ifc::kb_order(i, n_kb);
          asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(A_TILE_BYTES) : "memory");
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
              "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(col), "r"(row0), "r"(bar) : "memory");
        }
        __syncwarp();
        if (++r == p.ring) { r = 0; r_phase ^= 1u; }
      }
    }
    if (!p.use_tma) asm volatile("cp.async.wait_all;" ::: "memory");
    if (TRACE && lane == 0 && warp == WARP_LOADER) p.dbg_cycles[(int64_t)blockIdx.x * 16 + 3] = w0;
  } else if (warp >= WARP_BUILDER0 && p.bimg) {  // (warps 16..19 took the converter branch above)
    // ===================== B loader: one bulk copy (16 KB at TN = 128) of a precomputed stage image per k-block ========
    // B does not depend on theta: b_image_kernel writes every (column tile, k-block) tile once per model, in exactly the
    // bytes the builder warps would store, and the image stays in L2.  The copy is written through the async proxy and
    // read by the tensor core through the async proxy: no cross-proxy fence; the builders' tap reads and tile stores
    // (368 of the ~690 shared-memory wavefronts per k-block) are gone.
    if (warp == WARP_BUILDER0 && lane == 0) {
      int64_t g = 0;
      long long w0 = 0;
      const long long t_begin = TRACE ? clock64() : 0;
      for (int q = 0; q < my_tiles; ++q) {
        const int ct = my_sched[q] >> 16;
        const int n_kb = p.tile_nkb[ct];
        const unsigned char* src = p.bimg + (size_t)p.tile_bimg[ct] * C::B_STAGE_BYTES;
        for (int i = 0; i < n_kb; ++i, ++g) {
          const int sb = (int)(g % C::B_STAGES);
          if (g >= C::B_STAGES) mbar_wait32_timed<TRACE>(bempty + 8 * sb, (uint32_t)(((g / C::B_STAGES) - 1) & 1), w0);
          const uint32_t bar = bfull + 8 * sb, dst = bring + (uint32_t)sb * C::B_STAGE_BYTES;
          asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(C::B_STAGE_BYTES) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                       "l"(src + (size_t)i * C::B_STAGE_BYTES), "r"(C::B_STAGE_BYTES), "r"(bar) : "memory");
        }
      }
      if (TRACE) {
        long long* o = p.dbg_cycles + (int64_t)blockIdx.x * 16;
        o[4] = w0; o[5] = clock64() - t_begin;
      }
    }
  } else if (warp >= WARP_BUILDER0) {
    // ===================== B builders: tap table -> interpolated, TF32-split k-block tiles =====================
    // Builder team tm (two warps, each half of the columns) owns the k-blocks g = tm (mod 4) of this CTA's sequence:
    // the generic -> async proxy fence a stage needs before the tensor core may read it (a MEMBAR) is paid once per
    // warp and four k-blocks, and a single warp's dependent-issue latency is spread over eight warps.
    // Lane = (column j of a group of eight, 16-byte k chunk qc): five tap reads give the four interpolated elements
    // 4 qc .. 4 qc + 3 of column n, stored as one 16-byte vector of hi parts and one of lo parts.  The eight rows of a
    // group are one swizzle atom: a warp's 32 stores spread evenly over the eight bank groups.
    constexpr int ITS = TN / 16;  // column groups of this warp's half
    const int wb = warp - WARP_BUILDER0, tm = wb >> 1, it_base = (wb & 1) * ITS, j = lane >> 2, qc = lane & 3;
    const uint32_t soff0 = (uint32_t)(j * 128 + ((qc ^ j) << 4));
    int64_t g = 0;
    long long w0 = 0;
    const long long t_begin = TRACE ? clock64() : 0;
    for (int q = 0; q < my_tiles; ++q) {
      const int ct = my_sched[q] >> 16;
      const int n_kb = p.tile_nkb[ct];
      uint32_t cx[ITS];
      float cf[ITS];
      bool loaded = false;
      for (int i = 0; i < n_kb; ++i, ++g) {
        if ((int)(g & 3) != tm) continue;
        if (!loaded) {
#pragma unroll
          for (int it = 0; it < ITS; ++it) {
            const int n = (it_base + it) * 8 + j;
            cx[it] = wtab + 4u * (uint32_t)(p.col_idx[ct * TN + n] - 4 * qc - 3);
            cf[it] = p.col_frac[ct * TN + n];
          }
          loaded = true;
        }
        const int kb = ifc::kb_order(i, n_kb);
        const int sb = (int)(g % C::B_STAGES);
        if (g >= C::B_STAGES) {
          if (TRACE || p.sleep_build == 0) mbar_wait32_timed<TRACE>(bempty + 8 * sb, (uint32_t)(((g / C::B_STAGES) - 1) & 1), w0);
          else mbar_wait32_sleep(bempty + 8 * sb, (uint32_t)(((g / C::B_STAGES) - 1) & 1), p.sleep_build);
        }
        const uint32_t st = bring + (uint32_t)sb * C::B_STAGE_BYTES + soff0 + (uint32_t)(it_base * 1024);
        const uint32_t koff = (uint32_t)(4 * TILE_K * kb);
        // groups of four column groups: all twenty tap reads first (ptxas keeps shared loads behind earlier shared
        // stores), then the arithmetic, then the eight stores
#pragma unroll
        for (int it0 = 0; it0 < ITS; it0 += 4) {
          float L[4][5];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t wp = cx[it0 + u] - koff;  // tap at x - 3 + m lives at wp + 4 m, x = col_idx - 16 kb - 4 qc
#pragma unroll
            for (int m = 0; m < 5; ++m) L[u][m] = lds32_const(wp + 4 * m);
          }
          uint32_t hi[4][4], lo[4][4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int t = 0; t < 4; ++t)  // element e = 4 qc + t: lerp(tap[x - t], tap[x - t + 1], frac)
              split_fast(fmaf(cf[it0 + u], L[u][4 - t] - L[u][3 - t], L[u][3 - t]), hi[u][t], lo[u][t]);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            sts128(st + (uint32_t)((it0 + u) * 1024), hi[u][0], hi[u][1], hi[u][2], hi[u][3]);
            sts128((st + (uint32_t)((it0 + u) * 1024)) ^ 64u, lo[u][0], lo[u][1], lo[u][2], lo[u][3]);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) mbar_arrive32(bfull + 8 * sb);
      }
    }
    if (TRACE && wb == 0 && lane == 0) {
      long long* o = p.dbg_cycles + (int64_t)blockIdx.x * 16;
      o[4] = w0; o[5] = clock64() - t_begin;
    }
  } else if (warp < CONVERTER_WARPS + EPILOGUE_WARPS) {
    // ===================== epilogue (warps 8-11): add the accumulators, store the tile =====================
    const int quad = warp & 3;
    const int my_row = quad * 32 + lane;
    constexpr int CHUNKS = TN / 4;  // 16-byte chunks per staging row
    long long w0 = 0, t_drain = 0;
    for (int q = 0; q < my_tiles; ++q) {
      const int ct = my_sched[q] >> 16, rb = my_sched[q] & 0xffff;
      if (TRACE) mbar_wait32_timed<TRACE>(acc_full, (uint32_t)(q & 1), w0);
      else mbar_wait32_sleep(acc_full, (uint32_t)(q & 1), p.sleep_epi);
      asm volatile("tcgen05.fence::after_thread_sync;");
      const long long t_d0 = TRACE ? clock64() : 0;
      const uint32_t srow = staging + (uint32_t)my_row * (TN * 4);
#pragma unroll 2
      for (int c0 = 0; c0 < TN; c0 += 16) {
        uint32_t v[C::N_ACC_HI + 1][16];
        const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
#pragma unroll
        for (int a = 0; a <= C::N_ACC_HI; ++a) tmem_ld16(taddr + (uint32_t)(a * TN), v[a]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e4 = 0; e4 < 16; e4 += 4) {
          float f[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float hh = __uint_as_float(v[0][e4 + e]);
            if (C::N_ACC_HI >= 2) hh += __uint_as_float(v[C::N_ACC_HI >= 2 ? 1 : 0][e4 + e]);
            if (C::N_ACC_HI == 4) hh += __uint_as_float(v[C::N_ACC_HI == 4 ? 2 : 0][e4 + e]) + __uint_as_float(v[C::N_ACC_HI - 1][e4 + e]);
            f[e] = hh + __uint_as_float(v[C::N_ACC_HI][e4 + e]);
          }
          const int chunk = (c0 + e4) >> 2;
          sts128(srow + ((uint32_t)(chunk ^ (my_row & (CHUNKS - 1))) << 4), __float_as_uint(f[0]), __float_as_uint(f[1]),
                 __float_as_uint(f[2]), __float_as_uint(f[3]));
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();  // also orders this warp's staging writes before its reads below
      if (lane == 0) mbar_arrive32(acc_free);
      if (TRACE) t_drain += clock64() - t_d0;
      // coalesced stores of this warp's 32 rows from the staging tile
      constexpr int ROWS_PER_IT = 32 / CHUNKS == 0 ? 1 : 32 / CHUNKS;  // 1 (TN = 128) or 2 (TN = 64)
#pragma unroll 4
      for (int rr = 0; rr < 32; rr += ROWS_PER_IT) {
        const int trow = quad * 32 + rr + (ROWS_PER_IT == 2 ? (lane >> 4) : 0);
        const int chunk = lane & (CHUNKS - 1);
        const float4 f4 = lds128(staging + (uint32_t)trow * (TN * 4) + ((uint32_t)(chunk ^ (trow & (CHUNKS - 1))) << 4));
        *reinterpret_cast<float4*>(p.out + (int64_t)(rb * TILE_M + trow) * p.out_pitch + ct * TN + 4 * chunk) = f4;
      }
      __syncwarp();  // staging rows are rewritten by the next tile
    }
    if (TRACE && warp == CONVERTER_WARPS && lane == 0) {
      long long* o = p.dbg_cycles + (int64_t)blockIdx.x * 16;
      o[6] = w0; o[7] = t_drain;
    }
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = ifc::make_idesc(TN);
    int64_t g = 0;
    long long w0 = 0, w1 = 0, w2 = 0;
    const long long t_begin = TRACE ? clock64() : 0;
    for (int q = 0; q < my_tiles; ++q) {
      const int n_kb = p.tile_nkb[my_sched[q] >> 16];
      if (q >= 1) {  // the epilogue has drained the accumulators of the previous tile
        mbar_wait32_timed<TRACE>(acc_free, (uint32_t)((q - 1) & 1), w2);
        asm volatile("tcgen05.fence::after_thread_sync;");
      }
      for (int i = 0; i < n_kb; ++i, ++g) {
        const int sa = (int)(g % C::A_STAGES), sb = (int)(g % C::B_STAGES);
        mbar_wait32_timed<TRACE>(full + 8 * sa, (uint32_t)((g / C::A_STAGES) & 1), w0);
        mbar_wait32_timed<TRACE>(bfull + 8 * sb, (uint32_t)((g / C::B_STAGES) & 1), w1);
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (ifc::elect_one()) {
          const uint64_t d_bhi = ifc::make_desc(bring + (uint32_t)sb * C::B_STAGE_BYTES);
          const uint64_t d_blo = d_bhi + (uint64_t)(64 >> 4);
          const uint32_t acc_hi = tmem_d + (uint32_t)((i % C::N_ACC_HI) * TN), acc_x = tmem_d + (uint32_t)(C::N_ACC_HI * TN);
          const uint32_t a_hi = tmem_d + C::ACC_COLS + 32 * sa, a_lo = a_hi + TILE_K;
#pragma unroll
          for (int ks = 0; ks < TILE_K / 8; ++ks) {
            const uint64_t adv = (uint64_t)((ks * 8 * 4) >> 4);
            ifc::mma_tf32_ts(acc_hi, a_hi + 8 * ks, d_bhi + adv, idesc, (i >= C::N_ACC_HI || ks) ? 1u : 0u);
            ifc::mma_tf32_ts(acc_x, a_hi + 8 * ks, d_blo + adv, idesc, (i | ks) ? 1u : 0u);
            ifc::mma_tf32_ts(acc_x, a_lo + 8 * ks, d_bhi + adv, idesc, 1u);
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty + 8 * sa) : "memory");
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bempty + 8 * sb) : "memory");
          if (i == n_kb - 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(acc_full) : "memory");
        }
        __syncwarp();
      }
    }
    if (TRACE && lane == 0) {
      long long* o = p.dbg_cycles + (int64_t)blockIdx.x * 16;
      o[8] = w0; o[9] = w1; o[10] = w2; o[11] = clock64() - t_begin; o[12] = g;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (TRACE && tid == 0) p.dbg_cycles[(int64_t)blockIdx.x * 16 + 14] = clock64() - t_entry;
  if (warp == WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(512));
}

// One stage image per (column tile, sequence position): the builder warps' arithmetic and store pattern (see above) with
// the tap table read from global memory and the stores going to the image instead of a ring stage.  64 threads = the two
// warps of a builder team; grid (n_col_tiles, max k-blocks of a tile).
template <int TN>
__global__ void __launch_bounds__(64) b_image_kernel(const int* __restrict__ tile_nkb, const int* __restrict__ tile_bimg,
                                                     const int* __restrict__ col_idx, const float* __restrict__ col_frac,
                                                     const float* __restrict__ wtab, int wlen, unsigned char* __restrict__ bimg) {
  using C = Cfg<TN>;
  constexpr int ITS = TN / 16;
  const int ct = blockIdx.x, i = blockIdx.y;
  const int n_kb = tile_nkb[ct];
  if (i >= n_kb) return;
  const int wb = threadIdx.x >> 5, lane = threadIdx.x & 31, it_base = wb * ITS, j = lane >> 2, qc = lane & 3;
  const uint32_t soff0 = (uint32_t)(j * 128 + ((qc ^ j) << 4));
  const int kb = ifc::kb_order(i, n_kb);
  unsigned char* img = bimg + (size_t)(tile_bimg[ct] + i) * C::B_STAGE_BYTES;
  for (int it = 0; it < ITS; ++it) {
    const int n = (it_base + it) * 8 + j;
    const int x0 = col_idx[ct * TN + n] - 4 * qc - 3 - TILE_K * kb;  // tap at x - 3 + m is wtab[x0 + m]
    const float cf = col_frac[ct * TN + n];
    float L[5];
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      const int e = x0 + m;
      L[m] = (e >= 0 && e < wlen) ? wtab[e] : 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) split_fast(fmaf(cf, L[4 - t] - L[3 - t], L[3 - t]), hi[t], lo[t]);
    const uint32_t off = soff0 + (uint32_t)((it_base + it) * 1024);
    *reinterpret_cast<uint4*>(img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(img + (off ^ 64u)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

}  // namespace pxc
