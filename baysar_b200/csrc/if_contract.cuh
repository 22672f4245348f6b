// K2b — instrument-function contraction on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Replaces `fftconvolve(spectra, instrument_function_fm, "same").clip(spectra.min())` of
// SpectrometerChord.forward_model (reference baysar/spectrometers.py:269-271) for 128 thetas at a time:
//
//     C[theta, j] = sum_k' S[theta, k'] * T[k', j],     T[k', j] = w[j + o - k']   (banded Toeplitz, o = (K_I - 1) / 2)
//
// * M = 128 thetas (one TMEM lane each), N = 128 fine-grid outputs per tile, K walks the band in blocks of 16.
// * Persistent: one CTA per SM loops over (theta block, output tile) pairs; neighbouring CTAs work on neighbouring
//   output tiles of the same thetas, so the overlapping parts of their k-ranges are served by L2.
// * B operand — the Toeplitz tile depends only on (j0 - k0): ALL k-blocks of the band are windows of ONE tall
//   "master" tile (128 + 16 (n_kb - 1) rows, K-major SWIZZLE_128B; each 128-byte row holds 16 taps as TF32 "hi" parts
//   and the same 16 taps' TF32 "lo" parts), which every CTA builds once in shared memory (90 KB for a 455-tap
//   instrument function).  The B descriptor of k-block kb is the master's base plus 2048 (n_kb - 1 - kb) bytes
//   (+ 64 bytes for the lo half); nothing is generated or copied per k-block.
// * A operand — the pre-convolution spectra, plain FP32 rows [theta][pitch] in HBM/L2 (zero padded).  A k-block of a
//   tile (128 rows x 64 bytes) is copied global -> shared with coalesced 16-byte cp.async (a row-per-thread LDG
//   pattern was measured L1-tag bound at 75 % l1tex utilisation) into a ring of up to 8 stages, XOR-swizzled so that
//   the row-per-thread reads that follow are conflict-free.  Converter threads then own one theta row each: they
//   read 64 bytes, split the values into two TF32 terms (hi, lo) and write them to TENSOR MEMORY with tcgen05.st;
//   the MMAs take A from TMEM (the .ts form), so the tensor pipe's only shared-memory reads are the B windows.
// * 3xTF32: hi*hi accumulates in one FP32 TMEM tile, hi*lo + lo*hi in a second one; the epilogue adds them.
//   TMEM accumulation truncates (measured: signed error mean -2.5e-6, never positive, with one accumulator and
//   k-blocks in natural order), so (a) the 2^-11-sized cross terms get their own accumulator and (b) k-blocks are
//   visited from the edges of the band towards its centre: the big contributions arrive last and most roundings
//   happen while the accumulator is still small.
// * The tensor pipe is the bound: a 128 x 128 x 8 TF32 MMA was measured at ~108 cycles on B200 (the issuing thread
//   spends its whole time in tcgen05.mma with every other stage removed; N = 96 costs the same as N = 128).  What
//   the kernel can still lose is the gap between two tiles, while the accumulators are read out: the epilogue
//   therefore only DRAINS them (tcgen05.ld, add, clip, swizzled st.shared into a 64 KB staging tile) before it
//   releases them to the MMAs of the next tile, and does the trapz partial sums and the global stores afterwards,
//   from shared memory, coalesced along the rows.
// * TMEM map (512 columns): accumulator hi*hi [0,128), accumulator cross terms [128,256), eight A stages of 16 hi +
//   16 lo columns from 256.
// * Warp roles: warps 0-7 convert A (two groups of four warps take alternate k-blocks; warp w: TMEM lanes
//   32 (w % 4)..), warps 8-11 run the epilogue, warp 12 issues the MMAs, warp 13 issues the cp.async copies.
//   landed[r] (cp.async -> converters), consumed[r] (converters -> loader), full[s] (converters -> issuer),
//   empty[s] (tcgen05.commit -> converters), acc_full (commit -> epilogue), acc_free (epilogue -> issuer)
//   mbarriers; no block-wide barrier after the prologue.
// * Epilogue: clip at the per-theta pre-convolution minimum, trapz partial per (theta, tile) in float64
//   (deterministic: fixed shuffle tree, summed in a fixed order by the pixel kernel), FP32 store — every column, or
//   only the columns a static wavelength map reads (compact rows).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ifc {

constexpr int TILE_M = 128;   // thetas per tile
constexpr int TILE_N = 128;   // fine-grid outputs per tile
constexpr int TILE_K = 16;    // TF32 elements per k-block (two 8-deep MMA steps)
constexpr int A_STAGES = 8;     // tensor-memory stages of the split A operand
constexpr int RING_MAX = 8;     // shared-memory stages of the raw FP32 A tiles (8 KB each): 2..8, what fits
constexpr int A_TILE_BYTES = TILE_M * TILE_K * 4;
constexpr int CONVERTER_WARPS = 8, EPILOGUE_WARPS = 4;
constexpr int THREADS = (CONVERTER_WARPS + EPILOGUE_WARPS + 2) * 32;  // + MMA issuer + loader
constexpr int TMEM_COLS = 512;
constexpr int TMEM_ACC1 = TILE_N;          // cross-term accumulator
constexpr int TMEM_A = 2 * TILE_N;         // first A stage
constexpr int TMEM_A_STAGE = 2 * TILE_K;   // columns per A stage (hi, then lo)
constexpr int MAX_KB = 64;                 // master tile <= 1136 rows x 128 B (+ 64 KB staging + ring >= 2 stages)
constexpr int STAGING_BYTES = TILE_M * TILE_N * 4;  // epilogue staging tile (64 KB)
constexpr size_t SMEM_LIMIT = 232448;      // 227 KB of dynamic shared memory per CTA

struct Params {
  const float* s;         // [n_rows][pitch]  FP32 spectra, zero padded: column pad_left + k'
  int64_t pitch;          // floats per row (multiple of 32)
  int pad_left;           // multiple of 16
  int F;                  // valid outputs per row
  int n_row_blocks;       // blocks of 128 rows (workspace rows beyond the batch hold stale but finite data)
  int m_lo, m_hi;         // k-blocks k0 = j0 + 16 * m for m in [m_lo, m_hi]
  const float* taps_rev;  // [2][4][taps_len]: rev[i] = w[x_max - i], x_max = TILE_N - 1 + o - 16 m_lo (zero outside the
                          // tap window), TF32 hi then lo, each in 4 element-shifted copies copy_s[i] = rev[i + s]
  int taps_len;           // multiple of 4, >= TILE_N + 16 (m_hi - m_lo) + 20
  const float* row_min;   // [n_rows] clip level (pre-convolution minimum), may be nullptr
  const double* aux;      // alternative source of the clip level: aux[row][8], element 1 (rows < aux_rows), may be nullptr
  int aux_rows;
  float* out;             // [n_rows][out_pitch]
  int64_t out_pitch;
  double* trapz_part;     // [n_rows][n_tiles]  sum over the tile of w_j * out[j] (w = 1/2 at j = 0, F-1), may be nullptr
  int n_tiles;
  int ring;               // shared-memory A stages (ring_stages(n_kb))
  // compact output (static wavelength map): only the columns the pixel stage reads are stored, packed in column order
  const uint32_t* cmask;  // [n_tiles * 4] bit e of word w: column 32 w + e is needed; nullptr = store every column to `out`
  const int* cbase;       // [n_tiles * 4] number of needed columns before word w
  float* out_c;           // [n_rows][out_c_pitch]
  int64_t out_c_pitch;
  int debug;              // timing experiments only (wrong results): 1: one MMA per k-step; 4: no A loads; 8: no output
                          // stores; 16: no tcgen05.st
  long long* dbg_cycles;  // optional [gridDim.x][8] cycles spent waiting per role (timing experiments)
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

template <bool DEBUG>
__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, long long& acc) {
  if (DEBUG) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
  } else {
    mbar_wait(bar, parity);
  }
}

// one elected lane of a converged warp (the compiler then issues tcgen05 instructions without a per-lane loop)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
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

// D[tmem] (+)= A[tmem] * B[smem]: the A-from-tensor-memory form (cute::SM100_MMA_TF32_TS)
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate));
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

// 32 lanes x 16 columns, registers -> tensor memory
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// Split an FP32 value into two TF32 terms (round-to-nearest): v ~ hi + lo with |v - hi - lo| <= 2^-22 |v|.
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(v - __uint_as_float(hi)));
}

inline int master_rows(int n_kb) { return TILE_N + TILE_K * (n_kb - 1); }
inline size_t smem_fixed(int n_kb) {
  return (size_t)master_rows(n_kb) * 128 + STAGING_BYTES + (2 * A_STAGES + 2 * RING_MAX + 2) * sizeof(uint64_t) + 16 +
         1024;  // + alignment slack
}
inline int ring_stages(int n_kb) {
  const long room = (long)SMEM_LIMIT - (long)smem_fixed(n_kb);
  const int r = (int)(room / A_TILE_BYTES);
  return r > RING_MAX ? RING_MAX : r;
}
inline size_t smem_bytes(int n_kb) { return smem_fixed(n_kb) + (size_t)ring_stages(n_kb) * A_TILE_BYTES; }

// byte offset of (row, 16-byte chunk 0..3) inside a ring stage (128 rows x 64 bytes): rows two apart share a bank
// quad, the XOR spreads a quarter-warp's eight row-per-thread reads over all eight 16-byte bank groups
__device__ __forceinline__ uint32_t ring_off(int row, int chunk) { return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }

template <bool DEBUG>
__global__ void __launch_bounds__(THREADS, 1) if_contract_kernel(Params p) {
  const int dbg_mode = DEBUG ? p.debug : 0;  // timing experiments compile away in the production instantiation
  extern __shared__ unsigned char smem_dyn[];
  // carve (1024-byte aligned): master | epilogue staging | A ring | barriers | TMEM slot
  unsigned char* smem_raw = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int n_kb = p.m_hi - p.m_lo + 1;
  const int R = TILE_N + TILE_K * (n_kb - 1);
  unsigned char* master = smem_raw;
  unsigned char* staging = smem_raw + (size_t)R * 128;
  unsigned char* ring = staging + STAGING_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)p.ring * A_TILE_BYTES);
  uint64_t* empty = full + A_STAGES;
  uint64_t* acc_full = empty + A_STAGES;
  uint64_t* acc_free = acc_full + 1;
  uint64_t* landed = acc_free + 1;         // [RING_MAX]
  uint64_t* consumed = landed + RING_MAX;  // [RING_MAX]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(consumed + RING_MAX);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t total_tiles = (int64_t)p.n_row_blocks * p.n_tiles;
  const int64_t my_tiles = total_tiles > blockIdx.x ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  // ---- prologue: master Toeplitz tile.  Row rho: chunks 0-3 hold the hi parts of the taps rev[base - rho + (0..15)],
  // chunks 4-7 the lo parts of the same taps.
  {
    const int base = TILE_N - 1 + TILE_K * (n_kb - 1);
    for (int c = tid; c < R * 4; c += THREADS) {
      const int rho = c >> 2, q = c & 3;
      const int idx = base - rho + 4 * q;
      const int sh = idx & 3;
      const float4 hi = *reinterpret_cast<const float4*>(p.taps_rev + (size_t)sh * p.taps_len + (idx - sh));
      const float4 lo = *reinterpret_cast<const float4*>(p.taps_rev + (size_t)(4 + sh) * p.taps_len + (idx - sh));
      *reinterpret_cast<float4*>(master + swz(rho, q)) = hi;
      *reinterpret_cast<float4*>(master + swz(rho, 4 + q)) = lo;
    }
  }
  if (tid == 0) {
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(&full[s], CONVERTER_WARPS / 2); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_free, EPILOGUE_WARPS);
    for (int r = 0; r < RING_MAX; ++r) { mbar_init(&landed[r], 32); mbar_init(&consumed[r], CONVERTER_WARPS / 2); }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == CONVERTER_WARPS + EPILOGUE_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // master tile (generic proxy) -> tensor-core proxy
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_d = *tmem_slot;

  if (warp < CONVERTER_WARPS) {
    // ===================== A converters: ring (shared memory) -> TF32 split -> tensor memory =====================
    // Group `grp` (warps 4 grp .. 4 grp + 3) takes the k-block iterations g = grp, grp + 2, ...  Software pipelined:
    // the tcgen05.st of an iteration completes (wait::st) while the group's next k-block is read and split, so full[]
    // is signalled one group iteration late; the four TMEM stages absorb the delay.
    const int quad = warp & 3, grp = warp >> 2;
    const int row_in_tile = quad * 32 + lane;
    const int64_t iters = my_tiles * n_kb;
    const uint32_t lane_base = tmem_d + ((uint32_t)(quad * 32) << 16) + TMEM_A;
    int r = grp % p.ring;
    uint32_t r_phase = (uint32_t)((grp / p.ring) & 1);
    int d_pending = -1;
    long long w_landed = 0, w_empty = 0, w_st = 0;
    const long long t_begin = DEBUG ? clock64() : 0;
    for (int64_t g = grp; g < iters; g += 2) {
      const int d = (int)(g % A_STAGES);
      mbar_wait_timed<DEBUG>(&landed[r], r_phase, w_landed);
      // this thread's 16 floats of its theta row (the whole 64-byte row of the stage)
      float4 raw[4];
      const unsigned char* row_ptr = ring + (size_t)r * A_TILE_BYTES;
#pragma unroll
      for (int c = 0; c < 4; ++c) raw[c] = *reinterpret_cast<const float4*>(row_ptr + ring_off(row_in_tile, c));
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        split_tf32(raw[c].x, hi[4 * c + 0], lo[4 * c + 0]);
        split_tf32(raw[c].y, hi[4 * c + 1], lo[4 * c + 1]);
        split_tf32(raw[c].z, hi[4 * c + 2], lo[4 * c + 2]);
        split_tf32(raw[c].w, hi[4 * c + 3], lo[4 * c + 3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&consumed[r]);  // this warp has read ring stage r (the loader may refill it)
      if (d_pending >= 0) {                      // the group's previous k-block is in tensor memory: hand it over
        const long long t0 = DEBUG ? clock64() : 0;
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (DEBUG) w_st += clock64() - t0;
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[d_pending]);
      }
      if (g >= A_STAGES) {  // TMEM stage d: the MMAs of its previous user are done
        mbar_wait_timed<DEBUG>(&empty[d], (uint32_t)(((g / A_STAGES) - 1) & 1), w_empty);
        asm volatile("tcgen05.fence::after_thread_sync;");
      }
      if (!(dbg_mode & 16)) {
        tmem_st16(lane_base + TMEM_A_STAGE * d, hi);
        tmem_st16(lane_base + TMEM_A_STAGE * d + TILE_K, lo);
      }
      d_pending = d;
      r += 2;
      if (r >= p.ring) { r -= p.ring; r_phase ^= 1u; }
    }
    if (d_pending >= 0) {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[d_pending]);
    }
    if (DEBUG && p.dbg_cycles && tid == 0) {
      long long* o = p.dbg_cycles + (int64_t)blockIdx.x * 8;
      o[0] = w_landed; o[1] = w_empty; o[2] = w_st; o[3] = clock64() - t_begin;
    }
  } else if (warp == CONVERTER_WARPS + EPILOGUE_WARPS + 1) {
    // ===================== A loader (warp 13): global -> ring with coalesced 16-byte cp.async =====================
    // Lane l copies the 16-byte chunk (l & 3) of rows (l >> 2) + 8 q, q = 0..15: every warp instruction covers eight
    // whole 64-byte row segments.  k-block iterations of this CTA's tiles run back to back.
    int r = 0;
    uint32_t r_phase = 0;
    int64_t g = 0;
    long long w_consumed = 0;
    for (int64_t q = 0; q < my_tiles; ++q) {
      const int64_t t = blockIdx.x + q * gridDim.x;
      const int rb = (int)(t / p.n_tiles), nt = (int)(t - (int64_t)rb * p.n_tiles);
      const float* base = p.s + (int64_t)(rb * TILE_M + (lane >> 2)) * p.pitch + p.pad_left + nt * TILE_N + TILE_K * p.m_lo + 4 * (lane & 3);
      for (int i = 0; i < n_kb; ++i, ++g) {
        if (g >= p.ring) mbar_wait_timed<DEBUG>(&consumed[r], r_phase ^ 1u, w_consumed);  // previous use has been read
        const float* src = base + TILE_K * kb_order(i, n_kb);
        unsigned char* dst = ring + (size_t)r * A_TILE_BYTES;
#pragma unroll 8
        for (int qq = 0; qq < ((dbg_mode & 4) ? 0 : 16); ++qq) {
          const int row = (lane >> 2) + 8 * qq;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + ring_off(row, lane & 3))),
                       "l"(src + (int64_t)(8 * qq) * p.pitch) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&landed[r])) : "memory");
        if (++r == p.ring) { r = 0; r_phase ^= 1u; }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (DEBUG && p.dbg_cycles && lane == 0) p.dbg_cycles[(int64_t)blockIdx.x * 8 + 4] = w_consumed;
  } else if (warp < CONVERTER_WARPS + EPILOGUE_WARPS) {
    // ===================== epilogue (warps 8-11) =====================
    // Phase 1 (accumulators busy): TMEM -> registers, add the two accumulators, clip, st.shared into the staging tile
    // (row r, 16-byte chunk c at chunk position c ^ (r & 31): conflict-free for the row-per-thread writes and for the
    // row-per-warp reads), then release the accumulators.  Phase 2: trapz partial and global stores from shared memory.
    const int quad = warp & 3;
    long long w_accfull = 0;
    for (int64_t q = 0; q < my_tiles; ++q) {
      const int64_t t = blockIdx.x + q * gridDim.x;
      const int rb = (int)(t / p.n_tiles), nt = (int)(t - (int64_t)rb * p.n_tiles);
      const int j0 = nt * TILE_N;
      const int my_row = quad * 32 + lane;
      {
        const int row = rb * TILE_M + my_row;
        float clip = -3.402823466e38f;
        if (p.row_min) clip = p.row_min[row];
        else if (p.aux && row < p.aux_rows) clip = (float)p.aux[(int64_t)row * 8 + 1];
        mbar_wait_timed<DEBUG>(acc_full, (uint32_t)(q & 1), w_accfull);
        asm volatile("tcgen05.fence::after_thread_sync;");
        unsigned char* srow = staging + (size_t)my_row * (TILE_N * 4);
#pragma unroll 1
        for (int c0 = 0; c0 < TILE_N; c0 += 32) {
          uint32_t v[32], u[32];
          const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
          tmem_ld32(taddr, v);
          tmem_ld32(taddr + TMEM_ACC1, u);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int e4 = 0; e4 < 32; e4 += 4) {
            float f[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float x = __uint_as_float(v[e4 + e]) + __uint_as_float(u[e4 + e]);
              f[e] = x < clip ? clip : x;
            }
            const int chunk = (c0 + e4) >> 2;
            *reinterpret_cast<float4*>(srow + (((chunk ^ (my_row & 31))) << 4)) = make_float4(f[0], f[1], f[2], f[3]);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncwarp();  // also orders this warp's staging writes before its reads below
        if (lane == 0) mbar_arrive(acc_free);  // the MMAs of the next tile may overwrite the accumulators
      }
      // ---- phase 2: this warp's 32 rows, one row per iteration, lane l holds columns 4 l .. 4 l + 3
      const bool interior = j0 > 0 && j0 + TILE_N < p.F;  // unit trapz weights, no bounds
      uint32_t need = 0xffffffffu;
      int cb = 0;
      if (p.cmask) {
        const int w = (j0 + 4 * lane) >> 5;
        need = p.cmask[w];
        cb = p.cbase[w];
      }
      const int bit0 = (4 * lane) & 31;
      for (int rr = 0; rr < 32; ++rr) {
        const int trow = quad * 32 + rr;
        const int row = rb * TILE_M + trow;
        const float4 f4 = *reinterpret_cast<const float4*>(staging + (size_t)trow * (TILE_N * 4) + ((lane ^ (trow & 31)) << 4));
        const float f[4] = {f4.x, f4.y, f4.z, f4.w};
        const int jc = j0 + 4 * lane;
        double ps;
        if (interior) {
          ps = (double)((f[0] + f[1]) + (f[2] + f[3]));
        } else {
          ps = 0.0;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = jc + e;
            if (j < p.F) ps += ((j == 0 || j == p.F - 1) ? 0.5 : 1.0) * (double)f[e];
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
        if (lane == 0 && p.trapz_part) p.trapz_part[(int64_t)row * p.n_tiles + nt] = ps;
        if (dbg_mode & 8) {
        } else if (p.cmask) {
          float* crow = p.out_c + (int64_t)row * p.out_c_pitch + cb;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (need & (1u << (bit0 + e))) crow[__popc(need & ((1u << (bit0 + e)) - 1u))] = f[e];
        } else if (interior || jc < p.F) {
          *reinterpret_cast<float4*>(p.out + (int64_t)row * p.out_pitch + jc) = f4;
        }
      }
      __syncwarp();  // staging rows are rewritten by the next tile's phase 1
    }
    if (DEBUG && p.dbg_cycles && warp == CONVERTER_WARPS && lane == 0) p.dbg_cycles[(int64_t)blockIdx.x * 8 + 5] = w_accfull;
  } else if (warp == CONVERTER_WARPS + EPILOGUE_WARPS) {
    // ===================== MMA issuer (warp 12) =====================
    const uint32_t idesc = make_idesc(TILE_N);
    const uint32_t mb = smem_u32(master);
    int64_t g = 0;
    long long w_full = 0, w_accfree = 0;
    for (int64_t q = 0; q < my_tiles; ++q) {
      const uint32_t acc0 = tmem_d, acc1 = tmem_d + TMEM_ACC1;
      if (q >= 1) {  // the epilogue has drained the accumulators of the previous tile
        mbar_wait_timed<DEBUG>(acc_free, (uint32_t)((q - 1) & 1), w_accfree);
        asm volatile("tcgen05.fence::after_thread_sync;");
      }
      for (int i = 0; i < n_kb; ++i, ++g) {
        const int s = (int)(g % A_STAGES);
        mbar_wait_timed<DEBUG>(&full[s], (uint32_t)((g / A_STAGES) & 1), w_full);
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (elect_one()) {
          const int kb = kb_order(i, n_kb);
          const uint64_t d_bhi = make_desc(mb + (uint32_t)(2048 * (n_kb - 1 - kb)));
          const uint64_t d_blo = d_bhi + (uint64_t)(64 >> 4);  // lo parts: second half of each 128-byte row
          const uint32_t a_hi = tmem_d + TMEM_A + TMEM_A_STAGE * s, a_lo = a_hi + TILE_K;
#pragma unroll
          for (int ks = 0; ks < TILE_K / 8; ++ks) {
            const uint64_t adv = (uint64_t)((ks * 8 * 4) >> 4);  // 32 bytes per k-step inside the 128-byte swizzle row
            mma_tf32_ts(acc0, a_hi + 8 * ks, d_bhi + adv, idesc, (i | ks) ? 1u : 0u);
            if ((dbg_mode & 3) != 1) {
              mma_tf32_ts(acc1, a_hi + 8 * ks, d_blo + adv, idesc, (i | ks) ? 1u : 0u);
              mma_tf32_ts(acc1, a_lo + 8 * ks, d_bhi + adv, idesc, 1u);
            }
          }
          // frees this A stage when the MMAs above (and all earlier ones) have completed
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
          if (i == n_kb - 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(acc_full)) : "memory");
        }
        __syncwarp();
      }
    }
    if (DEBUG && p.dbg_cycles && lane == 0) {
      p.dbg_cycles[(int64_t)blockIdx.x * 8 + 6] = w_full;
      p.dbg_cycles[(int64_t)blockIdx.x * 8 + 7] = w_accfree;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == CONVERTER_WARPS + EPILOGUE_WARPS)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS));
}

// test / utility: rows of FP32 spectra [n][F] -> zero-padded rows [n][pitch]
__global__ void pad_rows_kernel(const float* s, int64_t n, int F, float* out, int64_t pitch, int pad_left) {
  const int64_t r = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < F; j += gridDim.x * blockDim.x) out[r * pitch + pad_left + j] = s[r * F + j];
}

}  // namespace ifc
