// K2b — instrument-function contraction on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Replaces `fftconvolve(spectra, instrument_function_fm, "same").clip(spectra.min())` of
// SpectrometerChord.forward_model (reference baysar/spectrometers.py:269-271) for 128 thetas at a time:
//
//     C[theta, j] = sum_k' S[theta, k'] * T[k', j],     T[k', j] = w[j + o - k']   (banded Toeplitz, o = (K_I - 1) / 2)
//
// * M = 128 thetas (one TMEM lane each), N = 128 fine-grid outputs per CTA, K walks the band in blocks of 32.
// * A operand: the pre-convolution spectra, split into two TF32 terms (hi, lo) by the spectral kernel and stored
//   zero-padded in HBM/L2 as [N][pitch]; tiles are copied global -> shared with 16-byte cp.async into the canonical
//   K-major SWIZZLE_128B layout (8 rows x 128 B atoms, 16-byte chunk index XOR row-in-atom).
// * B operand: the Toeplitz tile depends only on (j0 - k0), so it is GENERATED in shared memory from the taps (kept
//   reversed, TF32-split, in four element-shifted copies so that every 4-tap chunk is one aligned 16-byte load);
//   streaming it from L2 as well would double the operand traffic and make the kernel L2-bound.
// * 3xTF32: hi*hi accumulates in one FP32 TMEM tile, hi*lo + lo*hi in a second one; the epilogue adds them.
//   TMEM accumulation truncates (measured: signed error mean -2.5e-6, never positive, with one accumulator and
//   k-blocks in natural order), so (a) the 2^-11-sized cross terms get their own accumulator and (b) k-blocks are
//   visited from the edges of the band towards its centre: the big contributions arrive last and most roundings
//   happen while the accumulator is still small.
// * Warp-specialised pipeline, 3 shared-memory stages: warps 0-7 produce (cp.async A tiles, generate B tiles) and
//   later run the epilogue; warp 8 issues the MMAs.  full[s] (producers -> issuer) and empty[s] (tcgen05.commit ->
//   producers) mbarriers; no block-wide barrier inside the k loop.
// * Epilogue: tcgen05.ld (32 lanes x 32 columns), clip at the per-theta pre-convolution minimum, trapz partial per
//   (theta, tile) in float64 (deterministic: summed in order by the pixel kernel), FP32 store.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ifc {

constexpr int TILE_M = 128;   // thetas per CTA
constexpr int TILE_N = 128;   // fine-grid outputs per CTA
constexpr int TILE_K = 32;    // TF32 elements per k-block = one 128-byte swizzle row
constexpr int STAGES = 3;
constexpr int PRODUCER_THREADS = 256;
constexpr int THREADS = PRODUCER_THREADS + 32;              // + one MMA-issuing warp
constexpr int OPERAND_BYTES = TILE_M * TILE_K * 4;          // 16 KB per operand per stage
constexpr int STAGE_BYTES = 4 * OPERAND_BYTES;              // A_hi, A_lo, B_hi, B_lo
constexpr int TMEM_COLS = 2 * TILE_N;                       // two FP32 accumulators

struct Params {
  const float* s_hi;      // [n_rows][pitch]  TF32-rounded spectra, zero padded: column pad_left + k'
  const float* s_lo;      // [n_rows][pitch]  TF32-rounded remainder
  int64_t pitch;          // floats per row (multiple of 32)
  int pad_left;           // multiple of 32
  int F;                  // valid outputs per row
  int n_rows;             // multiple of 128 (workspace rows beyond the batch hold stale but finite data)
  int m_lo, m_hi;         // k-blocks k0 = j0 + 32 * m for m in [m_lo, m_hi]
  const float* taps_rev;  // [2][4][taps_len]: rev[i] = w[x_max - i], x_max = 127 + o - 32 m_lo (zero outside the tap
                          // window), TF32 hi then lo, each in 4 element-shifted copies copy_s[i] = rev[i + s]
  int taps_len;           // multiple of 4, >= 164 + 32 (m_hi - m_lo)
  const float* row_min;   // [n_rows] clip level (pre-convolution minimum), may be nullptr
  const double* aux;      // alternative source of the clip level: aux[row][8], element 1 (rows < aux_rows), may be nullptr
  int aux_rows;
  float* out;             // [n_rows][out_pitch]
  int64_t out_pitch;
  double* trapz_part;     // [n_rows][n_tiles]  sum over the tile of w_j * out[j] (w = 1/2 at j = 0, F-1), may be nullptr
  int n_tiles;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), leading byte offset (16 B units, 1 for swizzled K-major) in [16,30), stride byte offset (8 rows x 128 B =
// 1024 B -> 64) in [32,46), version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10),
// both K-major, N >> 3 in [17,23), M >> 4 in [24,29).
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}

// byte offset of (row, 16-byte chunk) inside a K-major SWIZZLE_128B operand tile whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

// k-block visiting order: from the two ends of the band towards its centre (see header)
__device__ __forceinline__ int kb_order(int i, int n_kb) { return (i & 1) ? n_kb - 1 - (i >> 1) : (i >> 1); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

__global__ void __launch_bounds__(THREADS, 1) if_contract_kernel(Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: [STAGES][A_hi | A_lo | B_hi | B_lo] (1024-byte aligned), then taps, then barriers
  unsigned char* stage_base = smem_raw;
  float* taps = reinterpret_cast<float*>(smem_raw + STAGES * STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(taps + 8 * p.taps_len);
  uint64_t* empty = full + STAGES;
  uint64_t* all_done = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(all_done + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j0 = blockIdx.x * TILE_N;
  const int row0 = blockIdx.y * TILE_M;
  const int n_kb = p.m_hi - p.m_lo + 1;

  for (int i = tid; i < 8 * p.taps_len; i += THREADS) taps[i] = p.taps_rev[i];
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 2 * PRODUCER_THREADS); mbar_init(&empty[s], 1); }
    mbar_init(all_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = *tmem_slot;

  if (warp < 8) {
    // ===================== producers =====================
    for (int i = 0; i < n_kb; ++i) {
      const int s = i % STAGES;
      if (i >= STAGES) mbar_wait(&empty[s], (uint32_t)(((i / STAGES) - 1) & 1));  // MMAs of the previous user are done
      const int kb = kb_order(i, n_kb);
      unsigned char* st = stage_base + s * STAGE_BYTES;
      const int64_t col0 = (int64_t)p.pad_left + j0 + 32 * (p.m_lo + kb);
      // ---- A: 128 rows x 8 chunks for hi and lo, 16-byte cp.async, coalesced along the row
#pragma unroll
      for (int r = 0; r < (TILE_M * 8) / PRODUCER_THREADS; ++r) {
        const int c = tid + r * PRODUCER_THREADS;
        const int row = c >> 3, chunk = c & 7;
        const int64_t g = (int64_t)(row0 + row) * p.pitch + col0 + chunk * 4;
        const uint32_t off = swz(row, chunk);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(st + off)), "l"(p.s_hi + g) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(st + OPERAND_BYTES + off)), "l"(p.s_lo + g) : "memory");
      }
      // ---- B: Toeplitz tile, row n (output j0 + n), chunk q -> taps w[x], x = n - kk + o - 32 m, kk = 4q..4q+3.
      // In the reversed array the four taps are ascending: rev index = x_max - x = 127 - n + 4q + 32 kb + e.
#pragma unroll
      for (int r = 0; r < (TILE_N * 8) / PRODUCER_THREADS; ++r) {
        const int c = tid + r * PRODUCER_THREADS;
        const int n = c >> 3, q = c & 7;
        const int idx = 127 - n + 4 * q + 32 * kb;
        const int sh = idx & 3;
        const float4 hi = *reinterpret_cast<const float4*>(taps + sh * p.taps_len + (idx - sh));
        const float4 lo = *reinterpret_cast<const float4*>(taps + (4 + sh) * p.taps_len + (idx - sh));
        const uint32_t off = swz(n, q);
        *reinterpret_cast<float4*>(st + 2 * OPERAND_BYTES + off) = hi;
        *reinterpret_cast<float4*>(st + 3 * OPERAND_BYTES + off) = lo;
      }
      // two arrivals per thread on full[s]: one fires when this thread's cp.async copies have landed, the other is a
      // plain (release) arrive that publishes its st.shared writes of the B tile
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[s])) : "memory");
      mbar_arrive(&full[s]);
    }
  } else {
    // ===================== MMA issuer (warp 8) =====================
    // B_hi and B_lo sit back to back in a stage, i.e. they form ONE 256-row K-major operand: a single N = 256 MMA
    // computes A_hi*B_hi into accumulator 0 (columns 0-127) and A_hi*B_lo into accumulator 1 (columns 128-255) while
    // reading A_hi from shared memory once; A_lo*B_hi then adds into accumulator 1 with an N = 128 MMA.
    const uint32_t idesc256 = make_idesc(2 * TILE_N), idesc128 = make_idesc(TILE_N);
    for (int i = 0; i < n_kb; ++i) {
      const int s = i % STAGES;
      mbar_wait(&full[s], (uint32_t)((i / STAGES) & 1));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // producers' generic-proxy writes -> tensor-core proxy
      asm volatile("tcgen05.fence::after_thread_sync;");
      if (lane == 0) {
        const uint32_t a_hi = smem_u32(stage_base + s * STAGE_BYTES);
        const uint64_t d_ahi = make_desc(a_hi), d_alo = make_desc(a_hi + OPERAND_BYTES);
        const uint64_t d_bhi = make_desc(a_hi + 2 * OPERAND_BYTES);
#pragma unroll
        for (int ks = 0; ks < TILE_K / 8; ++ks) {
          const uint64_t adv = (uint64_t)((ks * 8 * 4) >> 4);  // 32 bytes per k-step inside the 128-byte swizzle row
          mma_tf32(tmem_d, d_ahi + adv, d_bhi + adv, idesc256, (i | ks) ? 1u : 0u);
          mma_tf32(tmem_d + TILE_N, d_alo + adv, d_bhi + adv, idesc128, 1u);
        }
        // frees this stage when the MMAs above (and all earlier ones) have completed
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        if (i == n_kb - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(all_done)) : "memory");
      }
      __syncwarp();
    }
  }

  // ---- epilogue (warps 0-7): TMEM -> registers -> add accumulators, clip, trapz partial -> global
  if (warp < 8) {
    mbar_wait(all_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int chalf = warp >> 2;               // columns [64 * chalf, 64 * chalf + 64)
    const int row = row0 + quad * 32 + lane;
    float clip = -3.402823466e38f;
    if (p.row_min) clip = p.row_min[row];
    else if (p.aux && row < p.aux_rows) clip = (float)p.aux[(int64_t)row * 8 + 1];
    double part = 0.0;
    float* orow = p.out + (int64_t)row * p.out_pitch;
#pragma unroll 1
    for (int c0 = 64 * chalf; c0 < 64 * chalf + 64; c0 += 32) {
      uint32_t v[32], u[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
      tmem_ld32(taddr, v);
      tmem_ld32(taddr + TILE_N, u);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 32; q += 4) {
        float f[4];
        float psum = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float x = __uint_as_float(v[q + e]) + __uint_as_float(u[q + e]);
          x = x < clip ? clip : x;
          f[e] = x;
          const int j = j0 + c0 + q + e;
          if (j < p.F) {
            if (j == 0 || j == p.F - 1) part += 0.5 * (double)x; else psum += x;
          }
        }
        part += (double)psum;
        if (j0 + c0 + q < p.F) *reinterpret_cast<float4*>(orow + j0 + c0 + q) = make_float4(f[0], f[1], f[2], f[3]);
      }
    }
    if (p.trapz_part) {
      // the two column halves of a row live in warps w and w + 4: combine through shared memory (taps are dead now)
      double* xch = reinterpret_cast<double*>(taps);
      if (chalf == 1) xch[quad * 32 + lane] = part;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (chalf == 0) p.trapz_part[(int64_t)row * p.n_tiles + blockIdx.x] = part + xch[quad * 32 + lane];
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS));
}

// Split FP32 values into two TF32 terms (round-to-nearest): v ~ hi + lo with |v - hi - lo| <= 2^-22 |v|.
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
  hi = __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hi));
  lo = __uint_as_float(l);
}

// test / utility: rows of FP32 spectra [n][F] -> padded TF32-split rows
__global__ void split_rows_kernel(const float* s, int64_t n, int F, float* hi, float* lo, int64_t pitch, int pad_left) {
  const int64_t r = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < F; j += gridDim.x * blockDim.x) {
    float h, l;
    split_tf32(s[r * F + j], h, l);
    hi[r * pitch + pad_left + j] = h;
    lo[r * pitch + pad_left + j] = l;
  }
}

inline size_t smem_bytes(int taps_len) {
  return (size_t)STAGES * STAGE_BYTES + (size_t)8 * taps_len * sizeof(float) + (2 * STAGES + 1) * sizeof(uint64_t) + 16;
}

}  // namespace ifc
