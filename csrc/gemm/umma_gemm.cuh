// Persistent warp-specialised tcgen05 GEMM for sm_100a:
//
//     D[M,N] (+)= sum_{src} sum_k A_src[M,k] * B_src[N,k]       bf16 x bf16 -> fp32 (TMEM)
//
// * operands staged by TMA into 128B-swizzled shared memory, 64-wide K blocks, N-stage mbarrier ring
// * one elected thread issues tcgen05.mma (UMMA 128 x BN x 16), accumulators double-buffered in TMEM so
//   the epilogue of tile i overlaps the main loop of tile i+1
// * either operand may be K-major ([rows][K]) or MN-major ([K][rows]) — the latter is what makes the
//   weight-gradient / sufficient-factor outer product  dW = Uᵀ·V  a plain TMA GEMM with no transposes
// * the reduction may span several *sources* (one tensor map per source): for SFB the sources are the
//   peers' symmetric (u, v) buffers read over NVLink, gated by per-peer epoch flags
// * for implicit-GEMM convolution one operand is fetched by the TMA engine in im2col mode (IM2COL_A / IM2COL_B) or, where
//   the channel count does not allow that, built by cp.async gather warps (GATHER_A / GATHER_B, see conv_gather.cuh)
// * the TMA-producer and MMA-issuer loops run warp-uniform with one elected issuing lane (operands in uniform registers)
// * fused epilogues: bias + ReLU (+ mask) -> bf16 ; fp32 store / atomic split-K ; in-place SGD update
//
// Replaces the reference's cublasSgemm call sites (src/caffe/util/math_functions.cu:15-45) used by
// InnerProduct fwd/bwd (layers/inner_product_layer.cu:13-52), the per-image conv GEMMs
// (layers/conv_layer.cu:23-119) and ComputeGradientFromSV_gpu (layers/inner_product_layer.cu:55-64).
#pragma once
#include "sm100_prims.cuh"
#include "conv_gather.cuh"

namespace psd {

constexpr int kMaxSrc = 8;
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;              // two warps per TMEM lane quadrant, alternating 32-column chunks
constexpr int kNumThreads = 128 + 32 * kEpiWarps;   // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps4-11 epilogue
constexpr int kEpiWarp0 = 4;
constexpr int kGatherWarp0 = kEpiWarp0 + kEpiWarps;  // conv kernels: 8 more warps gather the implicit-im2col operand
constexpr int kGatherThreads = 256;          // 8 warps: the producers are instruction-latency bound, not bandwidth bound
// GATHER_A / GATHER_B: operand built by the cp.async gather warps (any channel count, ROW mode).
// IM2COL_A / IM2COL_B: operand fetched by the TMA engine in im2col mode (C_g % 64 == 0): no gather warps at all.
enum GatherMode : int { GATHER_NONE = 0, GATHER_A = 1, GATHER_B = 2, IM2COL_A = 3, IM2COL_B = 4 };
__host__ __device__ constexpr bool has_gather_warps(int g) { return g == GATHER_A || g == GATHER_B; }

enum EpiMode : int { EPI_BF16 = 0, EPI_F32 = 1, EPI_SGD = 2 };

struct TmapSet {
  CUtensorMap a[kMaxSrc];
  CUtensorMap b[kMaxSrc];
};

struct GemmParams {
  int M, N;              // output extent
  int kb_per_src;        // number of 64-wide k blocks per source
  int num_src;
  int split_k;           // >1: partial sums accumulated with atomics (EPI_F32 only)
  int src_rot;           // first source to visit (own rank for SFB: local data needs no flag wait)
  int cluster;           // CTAs per cluster sharing the TMA operand by multicast (1 = no clusters); conv kernels only
  int max_stages;        // experiment knob: use only this many ring stages (0 = all that fit)
  int no_bulk_epi;       // experiment knob: bit 0 / bit 1 = fp32 / bf16 output through the per-warp 32x32 walk instead of bulk row stores
  // --- epilogue operands
  __nv_bfloat16* c_bf16; // EPI_BF16 output [M, ldc]
  float* c_f32;          // EPI_F32 output [M, ldc]
  long ldc;
  const float* bias;     // optional, per output column (N)
  const __nv_bfloat16* mask;  // optional [M, ldc]: zero the output where mask <= 0 (fused ReLU backward)
  int relu;              // EPI_BF16: apply max(0, x) (negative_slope below)
  float relu_slope;
  int atomic;            // EPI_F32: atomicAdd instead of store
  int col_cg, col_cgk;   // EPI_F32 (conv wgrad over channel-padded K): output column (tap*cgk + c) -> tap*cg + c, c >= cg dropped
  float alpha;           // scale applied to the accumulator
  // --- EPI_SGD: W[M,N] fp32 master, H history, Wb bf16 shadow, in place
  float* w;
  float* h;
  __nv_bfloat16* wb;
  float lr, momentum, decay;
  const float* lr_dev;   // optional device-resident global learning rate (multiplies lr)
  int rule;              // 0 SGD, 1 Nesterov, 2 AdaGrad
  int l1;
  float delta;
  // --- peer gating (SFB): flag[src] must reach `epoch` before src's tiles are read
  const uint32_t* flags;
  uint32_t epoch;
  const uint32_t* epoch_dev;   // optional device-resident step counter added to `epoch` (CUDA-graph replays)
};

// CG = CTAs per UMMA (1, or 2 = paired CTAs: each CTA stages its own 128 rows of A but only BN/2 rows of B).
template <int BN, int CG = 1>
struct GemmSmem {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;     // 16 KB
  static constexpr int kBBytes = (BN / CG) * BLOCK_K * 2;   // this CTA's share of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);   // power of two >= 2 accumulators
  static constexpr int kBarBytes = 256;
  static constexpr int kEpiStageBytes = kEpiWarps * 32 * 33 * 4;    // per epilogue warp: 32x32 fp32 transpose tile (+1 pad)
  // as many ring stages as fit beside the barriers and the epilogue staging (at most 8: 4 @ BN 256, 6 @ 128, 8 @ <= 64;
  // paired CTAs: 6 @ 256, 8 @ 128)
  static constexpr int kFit = (232448 - 1024 - kBarBytes - kEpiStageBytes) / kStageBytes;
  static constexpr int kStages = kFit > 8 ? 8 : kFit;
  static constexpr int kTotal = kStages * kStageBytes + kBarBytes + kEpiStageBytes + 1024;  // +1024 alignment slack
  static_assert(kTotal <= 232448, "shared memory budget (227 KB per CTA)");
};

__device__ __forceinline__ float sgd_apply(float acc, float& w, float& h, const GemmParams& p, float lr) {
  float g = acc;
  if (p.decay != 0.f) g += p.decay * (p.l1 ? (w > 0.f ? 1.f : (w < 0.f ? -1.f : 0.f)) : w);
  float step;
  if (p.rule == 0) {
    h = lr * g + p.momentum * h;
    step = h;
  } else if (p.rule == 1) {
    float h_old = h;
    h = lr * g + p.momentum * h;
    step = (1.f + p.momentum) * h - p.momentum * h_old;
  } else {
    h = h + g * g;
    step = lr * g / (sqrtf(h) + p.delta);
  }
  w -= step;
  return w;
}

// Epilogue of one 32x32 accumulator sub-tile.  `r` holds, per lane, 32 consecutive columns of ONE row
// (tcgen05.ld 32x32b layout).  The sub-tile is transposed through a per-warp padded smem tile so that every
// global access of the warp is row-contiguous (64-128 B segments) instead of 32 scattered pieces.
template <int EPI>
__device__ __forceinline__ void epilogue_tile32(const GemmParams& p, const uint32_t (&r)[32], float* stage, int lane,
                                                int row0, int col0) {
  if (row0 >= p.M || col0 >= p.N) return;          // warp-uniform
#pragma unroll
  for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = __uint_as_float(r[j]) * p.alpha;
  __syncwarp();
  const int nrows = min(32, p.M - row0);
  const int ncols = min(32, p.N - col0);
  if constexpr (EPI == EPI_BF16) {
    // two rows per pass: lanes 0-15 -> row 2i, lanes 16-31 -> row 2i+1, each lane a column pair
    const int half = lane >> 4, cp = (lane & 15) * 2;
    const bool c0 = cp < ncols, c1 = cp + 1 < ncols;
    float b0 = 0.f, b1 = 0.f;
    if (p.bias != nullptr) {
      if (c0) b0 = __ldg(p.bias + col0 + cp);
      if (c1) b1 = __ldg(p.bias + col0 + cp + 1);
    }
    const bool pair_ok = c1 && ((p.ldc & 1) == 0) && (((col0 + cp) & 1) == 0);
    // fused ReLU-backward mask: issue all 16 loads of this lane up front (they are independent), then combine
    uint32_t mk[16];
    if (p.mask != nullptr) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int rr = 2 * i + half;
        const long off = static_cast<long>(row0 + rr) * p.ldc + col0 + cp;
        uint32_t v = 0x3f803f80u;                                      // (1.0, 1.0): keep
        if (rr < nrows) {
          if (pair_ok) v = *reinterpret_cast<const uint32_t*>(p.mask + off);
          else {
            const uint32_t lo = c0 ? __bfloat16_as_ushort(p.mask[off]) : 0x3f80u;
            const uint32_t hi = c1 ? __bfloat16_as_ushort(p.mask[off + 1]) : 0x3f80u;
            v = lo | (hi << 16);
          }
        }
        mk[i] = v;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int rr = 2 * i + half;
      if (rr >= nrows) continue;
      float x0 = stage[rr * 33 + cp] + b0, x1 = stage[rr * 33 + cp + 1] + b1;
      if (p.relu) {
        x0 = x0 > 0.f ? x0 : x0 * p.relu_slope;
        x1 = x1 > 0.f ? x1 : x1 * p.relu_slope;
      }
      const long off = static_cast<long>(row0 + rr) * p.ldc + col0 + cp;
      if (p.mask != nullptr) {
        const float m0 = __uint_as_float(mk[i] << 16), m1 = __uint_as_float(mk[i] & 0xffff0000u);
        if (!(m0 > 0.f)) x0 *= p.relu_slope;
        if (!(m1 > 0.f)) x1 *= p.relu_slope;
      }
      if (pair_ok) {
        *reinterpret_cast<__nv_bfloat162*>(p.c_bf16 + off) = __floats2bfloat162_rn(x0, x1);
      } else {
        if (c0) p.c_bf16[off] = __float2bfloat16(x0);
        if (c1) p.c_bf16[off + 1] = __float2bfloat16(x1);
      }
    }
  } else if constexpr (EPI == EPI_F32) {
    int col = col0 + lane;
    bool col_ok = lane < ncols;
    if (p.col_cgk > 0) {
      const int tap = col / p.col_cgk, c = col - tap * p.col_cgk;
      col_ok = col_ok && c < p.col_cg;
      col = tap * p.col_cg + c;
    }
    if (col_ok) {
#pragma unroll 4
      for (int rr = 0; rr < nrows; ++rr) {
        float* dst = p.c_f32 + static_cast<long>(row0 + rr) * p.ldc + col;
        const float v = stage[rr * 33 + lane];
        if (p.atomic) atomicAdd(dst, v);
        else *dst = v;
      }
    }
  }
  // EPI_SGD does not come through here: see sgd_rows_from_slab (CTA-cooperative, row-contiguous).
  __syncwarp();
}

// EPI_SGD, CTA-cooperative form.  The optimizer step is pure HBM streaming (18 B per weight), and DRAM only streams
// at full rate when whole rows of the tile (BN * 4 B = 1 KB) are touched back to back: the per-warp 32x32 sub-tile walk
// above issues isolated 128-byte pieces 36 KB apart and measured 2.3 TB/s against 4.9 TB/s for the same update as a
// linear stream.  Here the two warps that can read a TMEM lane quadrant dump its 32 x BN accumulators into one shared
// slab, then ALL epilogue warps update 4 full rows each: every warp issues BN/32 consecutive 128-byte accesses per row
// and array (one contiguous 1 KB burst), with 8 * BN/32 independent loads in flight.
template <int BN>
__device__ __forceinline__ void sgd_rows_from_slab(const GemmParams& p, const float* slab, int e, int lane, int row0,
                                                   int col0, float lr_eff) {
  constexpr int LD = BN + 1;
  constexpr int NC = BN / 32;
  constexpr int RW = 32 / kEpiWarps;                 // rows of the slab per warp
  float wv[RW][NC], hv[RW][NC];
#pragma unroll
  for (int k = 0; k < RW; ++k) {
    const int row = row0 + e * RW + k;
    const long off0 = static_cast<long>(row) * p.ldc + col0 + lane;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const bool ok = row < p.M && col0 + lane + 32 * i < p.N;
      wv[k][i] = ok ? p.w[off0 + 32 * i] : 0.f;
      hv[k][i] = ok ? p.h[off0 + 32 * i] : 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < RW; ++k) {
    const int row = row0 + e * RW + k;
    const long off0 = static_cast<long>(row) * p.ldc + col0 + lane;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      if (row < p.M && col0 + lane + 32 * i < p.N) {
        sgd_apply(slab[(e * RW + k) * LD + lane + 32 * i], wv[k][i], hv[k][i], p, lr_eff);
        p.w[off0 + 32 * i] = wv[k][i];
        p.h[off0 + 32 * i] = hv[k][i];
        if (p.wb != nullptr) p.wb[off0 + 32 * i] = __float2bfloat16(wv[k][i]);
      }
    }
  }
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

// Producer policy: both operands via TMA.  CG == 2: the loads complete on the LEADER CTA's full barrier (`bar` is then its
// shared::cluster address with the rank bit cleared) and this CTA fetches only its half of the B tile (rows / columns
// [crank * BN/2, (crank + 1) * BN/2) of the tile).
template <int CG>
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y) {
  if constexpr (CG == 2) tma_load_2d_cg2(dst, map, bar, x, y);
  else tma_load_2d_u32(dst, map, bar, x, y);
}
template <int BN, bool A_MN, bool B_MN, int CG = 1>
struct TmaProducer {
  static_assert(CG == 1 || !B_MN || (BN / CG) % 64 == 0, "MN-major B: each CTA's share must be whole 64-column chunks");
  __device__ static void load_a(const TmapSet& tm, int src, int kb, int m_blk, uint32_t sa, uint32_t bar) {
    const int k0 = kb * BLOCK_K;
    if constexpr (!A_MN) {
      tma_2d<CG>(sa, &tm.a[src], bar, k0, m_blk * BLOCK_M);
    } else {
#pragma unroll
      for (int c = 0; c < BLOCK_M / 64; ++c) tma_2d<CG>(sa + c * 8192, &tm.a[src], bar, m_blk * BLOCK_M + 64 * c, k0);
    }
  }
  __device__ static void load_b(const TmapSet& tm, int src, int kb, int n_blk, int crank, uint32_t sb, uint32_t bar) {
    const int k0 = kb * BLOCK_K;
    const int n0 = n_blk * BN + crank * (BN / CG);
    if constexpr (!B_MN) {
      tma_2d<CG>(sb, &tm.b[src], bar, k0, n0);                 // tensor-map box: BN / CG rows
    } else {
#pragma unroll
      for (int c = 0; c < BN / CG / 64; ++c) tma_2d<CG>(sb + c * 8192, &tm.b[src], bar, n0 + 64 * c, k0);
    }
  }
  __device__ static void load_stage(const TmapSet& tm, int src, int kb, int m_blk, int n_blk, int crank, uint32_t sa,
                                    uint32_t sb, uint32_t bar) {
    load_a(tm, src, kb, m_blk, sa, bar);
    load_b(tm, src, kb, n_blk, crank, sb, bar);
  }
};

// Tile enumeration.  Without clusters: tile -> (m fastest, then n, then split).  With a cluster of C CTAs the
// cluster walks "cluster tiles"; its CTAs take C consecutive blocks along the dimension the multicast (TMA)
// operand does NOT depend on: m for fprop/dgrad (weights shared), n for wgrad (dY slice shared).  Blocks past
// the edge are harmless: loads zero-fill / are out of range and the epilogue skips them.
struct TileCoord {
  int m_blk, n_blk, split;
};
// NFAST (optimizer-step epilogue): n fastest, so that the CTAs running at the same time update whole rows of W —
// one contiguous region of memory — instead of a 1 KB column slab with a row-sized stride between pieces.
template <int GATHER>
__device__ __forceinline__ TileCoord tile_coord(int t, int m_blocks, int n_blocks, int C, int crank, bool nfast) {
  TileCoord tc;
  if (nfast) {
    tc.n_blk = t % n_blocks;
    const int rest = t / n_blocks;
    tc.m_blk = rest % m_blocks;
    tc.split = rest / m_blocks;
  } else if (GATHER == GATHER_B || GATHER == IM2COL_A) {
    const int n_groups = (n_blocks + C - 1) / C;
    tc.m_blk = t % m_blocks;
    const int rest = t / m_blocks;
    tc.n_blk = (rest % n_groups) * C + crank;
    tc.split = rest / n_groups;
  } else {
    const int m_groups = (m_blocks + C - 1) / C;
    tc.m_blk = (t % m_groups) * C + crank;
    const int rest = t / m_groups;
    tc.n_blk = rest % n_blocks;
    tc.split = rest / n_blocks;
  }
  return tc;
}

// Paired CTAs: the pair walks (pair of m-blocks, n-block, split) tiles; CTA `crank` of the pair owns m-block 2*mp + crank.
// A phantom m-block past the edge (odd m_blocks) is harmless: its loads are out of range (zero-filled) and its
// epilogue rows are skipped.
__device__ __forceinline__ TileCoord tile_coord_pair(int t, int m_pairs, int n_blocks, int crank, bool nfast) {
  TileCoord tc;
  int mp;
  if (nfast) {
    tc.n_blk = t % n_blocks;
    const int rest = t / n_blocks;
    mp = rest % m_pairs;
    tc.split = rest / m_pairs;
  } else {
    mp = t % m_pairs;
    const int rest = t / m_pairs;
    tc.n_blk = rest % n_blocks;
    tc.split = rest / n_blocks;
  }
  tc.m_blk = 2 * mp + crank;
  return tc;
}

template <int BN, bool A_MN, bool B_MN, int EPI, int GATHER = GATHER_NONE, int CG = 1>
__global__ void __launch_bounds__(kNumThreads + (has_gather_warps(GATHER) ? kGatherThreads : 0), 1)
umma_gemm_kernel(const __grid_constant__ TmapSet tm, const GemmParams p, const ConvGeom cg) {
  static_assert(CG == 1 || CG == 2, "one CTA or a CTA pair per UMMA");
  static_assert(CG == 1 || !has_gather_warps(GATHER), "paired CTAs need TMA producers (cp.async cannot signal the leader)");
  static_assert(GATHER != GATHER_A || !A_MN, "gathered A is produced K-major");
  static_assert(GATHER != GATHER_B || B_MN, "gathered B is produced MN-major");
  static_assert(GATHER != IM2COL_A || !A_MN, "im2col A is K-major");
  static_assert(GATHER != IM2COL_B || B_MN, "im2col B is MN-major");
  using S = GemmSmem<BN, CG>;
  constexpr int kStages = S::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + kStages * S::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(bar_base + S::kBarBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nst = (p.max_stages > 0 && p.max_stages < kStages) ? p.max_stages : kStages;     // ring depth in use

  const int m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_blocks = (p.N + BN - 1) / BN;
  const int total_kb = p.kb_per_src * p.num_src;
  // CTAs per cluster sharing one operand by TMA multicast (CG == 1 only): the cp.async gather modes share the TMA (weight /
  // dY) operand; the im2col-TMA modes share the IM2COL operand itself — its delivery rate, not its bytes, is what bounds
  // the convolution kernels, so each CTA of the cluster fetches 1/C of the pixel rows for everybody.
  const int C = (CG == 1 && GATHER != GATHER_NONE && p.cluster > 1) ? p.cluster : 1;
  const int crank = (C > 1 || CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = CG == 1 || crank == 0;            // CG == 2: cluster rank 0 issues the MMAs for the pair
  const int prank = CG == 2 ? crank : 0;                // which half of the B tile this CTA stages (paired CTAs only)
  const uint16_t cmask = static_cast<uint16_t>((1u << C) - 1);
  // cluster-level tile space and stride
  const int m_pairs = (m_blocks + 1) / 2;
  const int num_tiles = CG == 2 ? m_pairs * n_blocks * p.split_k
                                : ((GATHER == GATHER_B || GATHER == IM2COL_A) ? m_blocks * ((n_blocks + C - 1) / C)
                                                                              : ((m_blocks + C - 1) / C) * n_blocks) * p.split_k;
  const int tile0 = blockIdx.x / (C * CG), tile_step = gridDim.x / (C * CG);
  // n fastest for the epilogues that stream whole fp32 rows (optimizer step; plain fp32 output of the inner-product weight
  // gradient): the CTAs running at the same time then cover complete rows — one contiguous region of memory
  // (not for the atomic split-K form: there the m-blocks of one weight tile should run back to back)
  const bool nfast = EPI == EPI_SGD || (EPI == EPI_F32 && GATHER == GATHER_NONE && !p.atomic);
  auto coord = [&](int t) -> TileCoord {
    if constexpr (CG == 2) return tile_coord_pair(t, m_pairs, n_blocks, crank, nfast);
    else return tile_coord<GATHER>(t, m_blocks, n_blocks, C, crank, nfast);
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_src; ++s) {
      tma_prefetch_desc(&tm.a[s]);
      tma_prefetch_desc(&tm.b[s]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1 + (has_gather_warps(GATHER) ? kGatherThreads : 0));   // CG == 2: only the leader's is used
      mbar_init(&empty_bar[s], C);        // released by the MMA thread of every CTA sharing the multicast operand
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps * CG);   // one arrive per epilogue warp (CG == 2: of both CTAs, on the leader's)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CG == 2) {
      tmem_alloc_cg2(tmem_slot, S::kTmemCols);
    } else {
      tmem_alloc(tmem_slot, S::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (C > 1 || CG == 2) cluster_sync_all();   // peers' barriers / TMEM exist before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop with uniform control flow and ONE elected lane issues the TMA / barrier
    // instructions: the operands then live in uniform registers (UTMALDG straight from UR), whereas a loop under
    // `if (lane == 0)` is divergent code in which every issue needs an R2UR waterfall.  ncu: the single-lane version
    // cost ~600 cycles per k-block, the floor of every BN <= 128 kernel.
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const TileCoord tc = coord(tile);
        const int m_blk = tc.m_blk, n_blk = tc.n_blk, split = tc.split;
        const int g0 = static_cast<int>(static_cast<long>(total_kb) * split / p.split_k);
        const int g1 = static_cast<int>(static_cast<long>(total_kb) * (split + 1) / p.split_k);
        // im2col modes: coordinates of the tile's first base pixel (A) / tap decode of the tile's k-columns (B)
        [[maybe_unused]] int im_w = 0, im_h = 0, im_n = 0;
        [[maybe_unused]] const int im_org_w = cg.off_w - (cg.dr < 0 ? cg.S - 1 : 0);
        [[maybe_unused]] const int im_org_h = cg.off_h - (cg.dr < 0 ? cg.R - 1 : 0);
        [[maybe_unused]] int im_c0[BN / CG / 64 > 0 ? BN / CG / 64 : 1];
        [[maybe_unused]] uint16_t im_offw[BN / CG / 64 > 0 ? BN / CG / 64 : 1], im_offh[BN / CG / 64 > 0 ? BN / CG / 64 : 1];
        if constexpr (GATHER == IM2COL_A) {
          // first base pixel this CTA fetches: the tile's, or (cluster multicast) that of its 128 / C pixel slice
          const uint32_t m0 = static_cast<uint32_t>(m_blk) * BLOCK_M + (C > 1 ? static_cast<uint32_t>(crank * (BLOCK_M / C)) : 0u);
          const uint32_t n_img = fdiv(m0, cg.div_ohow);
          const uint32_t rem = m0 - n_img * static_cast<uint32_t>(cg.OH * cg.OW);
          const uint32_t oh = fdiv(rem, cg.div_ow);
          const uint32_t ow = rem - oh * static_cast<uint32_t>(cg.OW);
          im_w = static_cast<int>(ow) * cg.sw + im_org_w;
          im_h = static_cast<int>(oh) * cg.sh + im_org_h;
          im_n = static_cast<int>(n_img);
        }
        if constexpr (GATHER == IM2COL_B) {
#pragma unroll
          for (int c = 0; c < BN / CG / 64; ++c) {
            int kc = n_blk * BN + prank * (BN / CG) + c * 64;
            if (kc >= cg.K) kc = 0;                      // columns past K are dropped by the epilogue: load anything valid
            const int tap = static_cast<int>(fdiv(static_cast<uint32_t>(kc), cg.div_cg));
            im_c0[c] = kc - tap * cg.Cgk;
            const int r = static_cast<int>(fdiv(static_cast<uint32_t>(tap), cg.div_s));
            im_offh[c] = static_cast<uint16_t>(r);
            im_offw[c] = static_cast<uint16_t>(tap - r * cg.S);
          }
        }
        // (source, k-block) advance incrementally: no divisions in the single-thread producer loop
        int src_lin = g0 / p.kb_per_src;
        int kb = g0 - src_lin * p.kb_per_src;
        int src = (src_lin + p.src_rot) % p.num_src;
        bool new_src = true;
        for (int g = g0; g < g1; ++g) {
          if (new_src) {
            if (p.flags != nullptr) {
              if (lane == 0) {
                const uint32_t ep = p.epoch + (p.epoch_dev != nullptr ? *reinterpret_cast<const volatile uint32_t*>(p.epoch_dev) : 0u);
                wait_flag_ge(p.flags + src, ep);
              }
              __syncwarp();
            }
            new_src = false;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStageBytes;
          if (elect_one()) {
          // completion barrier of this stage's loads: own full barrier, or (paired CTAs) the leader's
          const uint32_t fbar = CG == 2 ? (smem_u32(&full_bar[stage]) & kPeerBitMask) : smem_u32(&full_bar[stage]);
          const uint32_t sa32 = smem_u32(sa), sb32 = sa32 + S::kABytes;
          if constexpr (GATHER == GATHER_NONE) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], CG * S::kStageBytes);
            TmaProducer<BN, A_MN, B_MN, CG>::load_stage(tm, src, kb, m_blk, n_blk, prank, sa32, sb32, fbar);
          } else if constexpr (GATHER == IM2COL_A) {
            // A tile = 128 output pixels x 64 channels of tap (r, s): one im2col-mode TMA instruction
            const int k0 = g * BLOCK_K;
            const int tap = static_cast<int>(fdiv(static_cast<uint32_t>(k0), cg.div_cg));
            const int c0 = k0 - tap * cg.Cgk;
            const int r = static_cast<int>(fdiv(static_cast<uint32_t>(tap), cg.div_s));
            const int sx = tap - r * cg.S;
            const bool in_k = k0 < cg.K;
            const int offw = cg.dr > 0 ? sx : cg.S - 1 - sx, offh = cg.dr > 0 ? r : cg.R - 1 - r;
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], CG * S::kStageBytes);
            // k-blocks past K (K % 64 != 0 never happens here: C_g % 64 == 0) — keep the guard cheap
            if constexpr (CG == 2)
              tma_load_im2col_4d_cg2(sa32, &tm.a[0], fbar, in_k ? c0 : 0, im_w, im_h, im_n,
                                     static_cast<uint16_t>(in_k ? offw : 0), static_cast<uint16_t>(in_k ? offh : 0));
            else if (C > 1)       // this CTA's 128 / C pixel rows of the shared A tile, delivered to the whole cluster
              tma_load_im2col_4d_mcast(sa32 + crank * (BLOCK_M / C) * 128, &tm.a[0], fbar, in_k ? c0 : 0, im_w, im_h, im_n,
                                       static_cast<uint16_t>(in_k ? offw : 0), static_cast<uint16_t>(in_k ? offh : 0), cmask);
            else
              tma_load_im2col_4d(sa32, &tm.a[0], fbar, in_k ? c0 : 0, im_w, im_h, im_n,
                                 static_cast<uint16_t>(in_k ? offw : 0), static_cast<uint16_t>(in_k ? offh : 0));
            if constexpr (B_MN) {
              // dgrad straight from the FPROP weight layout [co][tap][ci] (no packed copy): the B tile of k-block
              // (tap, co0..co0+63) is BN/64 boxes {64 ci, 64 co, 1 tap} of a 3-D map — ci contiguous = MN-major B;
              // output channels past Cout_g and input channels past C_g are zero-filled by the TMA
              const int n0 = n_blk * BN + prank * (BN / CG);
#pragma unroll
              for (int c = 0; c < BN / CG / 64; ++c) {
                if constexpr (CG == 2) tma_load_3d_cg2(sb32 + c * 8192, &tm.b[0], fbar, n0 + 64 * c, in_k ? c0 : 0, in_k ? tap : 0);
                else tma_load_3d_u32(sb32 + c * 8192, &tm.b[0], fbar, n0 + 64 * c, in_k ? c0 : 0, in_k ? tap : 0);
              }
            } else {
              TmaProducer<BN, A_MN, B_MN, CG>::load_b(tm, src, kb, n_blk, prank, sb32, fbar);
            }
          } else if constexpr (GATHER == IM2COL_B) {
            // B tile = BN/64 chunks of [64 reduction pixels][64 channels of one tap]; the pixels advance with g.
            // Paired CTAs: this CTA fetches chunks [crank * BN/128, (crank + 1) * BN/128) of the tile.
            const uint32_t m0 = static_cast<uint32_t>(g) * BLOCK_K;
            const uint32_t n_img = fdiv(m0, cg.div_ohow);
            const uint32_t rem = m0 - n_img * static_cast<uint32_t>(cg.OH * cg.OW);
            const uint32_t oh = fdiv(rem, cg.div_ow);
            const uint32_t ow = rem - oh * static_cast<uint32_t>(cg.OW);
            const int bw = static_cast<int>(ow) * cg.sw + im_org_w, bh = static_cast<int>(oh) * cg.sh + im_org_h;
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], CG * S::kStageBytes);
#pragma unroll
            for (int c = 0; c < BN / CG / 64; ++c) {
              if constexpr (CG == 2) {
                tma_load_im2col_4d_cg2(sb32 + c * 8192, &tm.b[0], fbar, im_c0[c], bw, bh, static_cast<int>(n_img), im_offw[c],
                                       im_offh[c]);
              } else if (C > 1) {
                // cluster along Cout sharing the im2col B tile: chunk c is fetched by CTA (c % C) for everybody
                if (c % C == crank)
                  tma_load_im2col_4d_mcast(sb32 + c * 8192, &tm.b[0], fbar, im_c0[c], bw, bh, static_cast<int>(n_img),
                                           im_offw[c], im_offh[c], cmask);
              } else {
                tma_load_im2col_4d(sb32 + c * 8192, &tm.b[0], fbar, im_c0[c], bw, bh, static_cast<int>(n_img), im_offw[c],
                                   im_offh[c]);
              }
            }
            TmaProducer<BN, A_MN, B_MN, CG>::load_a(tm, src, kb, m_blk, sa32, fbar);
          } else if constexpr (GATHER == GATHER_A) {
            mbar_arrive_expect_tx(&full_bar[stage], S::kBBytes);     // the whole B tile lands here (C slices)
            if (C == 1) {
              TmaProducer<BN, A_MN, B_MN>::load_b(tm, src, kb, n_blk, 0, sb32, fbar);
            } else {
              // this CTA fetches rows [crank*BN/C, (crank+1)*BN/C) of the K-major weight tile for the whole cluster
              const int rows = BN / C;
              tma_load_2d_mcast(sa + S::kABytes + crank * rows * 128, &tm.b[src], &full_bar[stage], kb * BLOCK_K,
                                n_blk * BN + crank * rows, cmask);
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], S::kABytes);
            if (C == 1) {
              TmaProducer<BN, A_MN, B_MN>::load_a(tm, src, kb, m_blk, sa32, fbar);
            } else {
              // dY^T tile (MN-major: 2 chunks of [64 k-rows][64 cout]); CTA `crank` fetches k-rows [crank*64/C, ...)
              const int krows = BLOCK_K / C;
#pragma unroll
              for (int c = 0; c < BLOCK_M / 64; ++c)
                tma_load_2d_mcast(sa + c * 8192 + crank * krows * 128, &tm.a[src], &full_bar[stage],
                                  m_blk * BLOCK_M + 64 * c, kb * BLOCK_K + crank * krows, cmask);
            }
          }
          }  // elect_one
          __syncwarp();
          if (++stage == nst) { stage = 0; phase ^= 1; }
          if (++kb == p.kb_per_src) {
            kb = 0;
            if (++src == p.num_src) src = 0;
            new_src = true;
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues; paired CTAs: leader only) ==========
    {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M * CG, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
        const int split = coord(tile).split;
        const int g0 = static_cast<int>(static_cast<long>(total_kb) * split / p.split_k);
        const int g1 = static_cast<int>(static_cast<long>(total_kb) * (split + 1) / p.split_k);
        const int as = it & 1;
        mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int g = g0; g < g1; ++g) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * S::kStageBytes);
          const uint32_t b_addr = a_addr + S::kABytes;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t da = A_MN ? make_smem_desc(a_addr + k * (UMMA_K * 128), 8192, 1024)
                                       : make_smem_desc(a_addr + k * (UMMA_K * 2), 16, 1024);
              const uint64_t db = B_MN ? make_smem_desc(b_addr + k * (UMMA_K * 128), 8192, 1024)
                                       : make_smem_desc(b_addr + k * (UMMA_K * 2), 16, 1024);
              if constexpr (CG == 2) umma_bf16_cg2(d_tmem, da, db, idesc, (g > g0 || k > 0) ? 1u : 0u);
              else umma_bf16(d_tmem, da, db, idesc, (g > g0 || k > 0) ? 1u : 0u);
            }
            // frees the smem slot when these MMAs retire — in every CTA that shares the multicast operand / the pair
            if constexpr (CG == 2) {
              umma_commit_cg2(&empty_bar[stage], 0x3);
              if (g == g1 - 1) umma_commit_cg2(&tmem_full[as], 0x3);   // accumulator complete -> both CTAs' epilogues
            } else {
              if (C == 1) umma_commit(&empty_bar[stage]);
              else umma_commit_mcast(&empty_bar[stage], cmask);
              if (g == g1 - 1) umma_commit(&tmem_full[as]);      // accumulator complete -> epilogue
            }
          }
          __syncwarp();
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
        if (g0 >= g1 && elect_one()) {                                  // (never: every split owns >= 1 k-block)
          if constexpr (CG == 2) umma_commit_cg2(&tmem_full[as], 0x3);
          else umma_commit(&tmem_full[as]);
        }
      }
    }
  } else if (has_gather_warps(GATHER) && warp >= kGatherWarp0) {
    // ===================== gather producers: implicit im2col -> swizzled smem via cp.async =====================
    const int gt = threadIdx.x - kGatherWarp0 * 32;      // 0..kGatherThreads-1
    const int j = gt & 7;                                 // this thread's 16-byte chunk column
    const int rg = gt >> 3;                               // first row; further rows every kGatherThreads/8
    constexpr int kRowStep = kGatherThreads / 8;      // 32: row r and r + 32 share (r & 7), i.e. the same swizzle
    const GatherRegs gr = load_gather_regs(cg);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const TileCoord tc = coord(tile);
      const int m_blk = tc.m_blk, n_blk = tc.n_blk, split = tc.split;
      const int g0 = static_cast<int>(static_cast<long>(total_kb) * split / p.split_k);
      const int g1 = static_cast<int>(static_cast<long>(total_kb) * (split + 1) / p.split_k);
      if constexpr (GATHER == GATHER_A) {
        // rows of the A tile are fixed along K: fetch their row-table entries once per tile
        constexpr int kRows = BLOCK_M / kRowStep;
        RowPos pos[kRows];
#pragma unroll
        for (int i = 0; i < kRows; ++i) pos[i] = load_row(gr, static_cast<long>(m_blk) * BLOCK_M + rg + i * kRowStep);
        for (int g = g0; g < g1; ++g) {
          const ChunkOff t = decode_chunk_off(gr, g * BLOCK_K + j * 8);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes) + rg * 128 + ((j ^ (rg & 7)) << 4);
          if (gr.mode == 0) {
#pragma unroll
            for (int i = 0; i < kRows; ++i) gather_chunk_tap(gr, sa + i * (kRowStep * 128), pos[i], t);
          } else {
#pragma unroll
            for (int i = 0; i < kRows; ++i) gather_chunk_row(gr, sa + i * (kRowStep * 128), pos[i], t);
          }
          cp_async_mbar_arrive_noinc(&full_bar[stage]);
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
      } else {
        // B tile (MN-major): BN/64 chunk-columns of [64 reduction rows (m)][64 k-columns]; the k-columns are
        // fixed per tile (decode the taps once), the rows advance with the reduction index g — their row-table
        // entries for k-block g+1 are fetched before waiting for the slot of k-block g
        constexpr int kChunks = BN / 64 > 0 ? BN / 64 : 1;      // (BN < 64 is never instantiated with gather warps)
        constexpr int kRows = 64 / kRowStep;
        ChunkOff taps[kChunks];
#pragma unroll
        for (int c = 0; c < kChunks; ++c) taps[c] = decode_chunk_off(gr, n_blk * BN + c * 64 + j * 8);
        RowPos pos[kRows], nxt[kRows];
#pragma unroll
        for (int i = 0; i < kRows; ++i) nxt[i] = load_row(gr, static_cast<long>(g0) * BLOCK_K + rg + i * kRowStep);
        for (int g = g0; g < g1; ++g) {
#pragma unroll
          for (int i = 0; i < kRows; ++i) {
            pos[i] = nxt[i];
            nxt[i] = load_row(gr, static_cast<long>(g + 1) * BLOCK_K + rg + i * kRowStep);
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t sb = smem_u32(smem + stage * S::kStageBytes) + S::kABytes + rg * 128 + ((j ^ (rg & 7)) << 4);
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
#pragma unroll
            for (int i = 0; i < kRows; ++i) {
              if (gr.mode == 0) gather_chunk_tap(gr, sb + c * 8192 + i * (kRowStep * 128), pos[i], taps[c]);
              else gather_chunk_row(gr, sb + c * 8192 + i * (kRowStep * 128), pos[i], taps[c]);
            }
          }
          cp_async_mbar_arrive_noinc(&full_bar[stage]);
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + kEpiWarps) {
    // ===================== epilogue warps: TMEM -> registers -> global =====================
    const int e = warp - kEpiWarp0;
    const int q = e & 3;                          // TMEM lane quadrant == warp_id % 4
    const int half = e >> 2;                      // which of the two warps of this quadrant
    int it = 0;
    // EPI_SGD streams W and H (8 B read + 10 B written per element) — pure HBM traffic with nothing to reuse.
    // Optional: each epilogue thread asks the L2 for one row segment of the NEXT tile's W or H while the current
    // tile is processed.  Measured on B200 (AlexNet fc6+fc7 update): 336 us without vs 383 us with the prefetch —
    // the extra DRAM stream collides with the write-back of the previous tile — so it is compiled out.
    constexpr bool kSgdPrefetch = false;
    auto prefetch_sgd_tile = [&](int t) {
      if (EPI != EPI_SGD || t >= num_tiles || (p.N & 3) != 0 || (p.ldc & 3) != 0) return;
      const TileCoord pc = coord(t);
      const int te = (warp - kEpiWarp0) * 32 + lane;      // 0..255: row = te >> 1, W or H = te & 1
      const int row = pc.m_blk * BLOCK_M + (te >> 1);
      const int col = pc.n_blk * BN;
      const int cols = min(BN, p.N - col);
      if (row < p.M && cols > 0) {
        const float* base = (te & 1) ? p.h : p.w;
        if ((reinterpret_cast<uintptr_t>(base) & 15) == 0)
          prefetch_l2_bulk(base + static_cast<long>(row) * p.ldc + col, static_cast<uint32_t>(cols) * 4u);
      }
    };
    if (kSgdPrefetch) prefetch_sgd_tile(tile0);
    [[maybe_unused]] const bool bulk_f32 = EPI == EPI_F32 && !p.atomic && !(p.no_bulk_epi & 1) && p.col_cgk == 0 && (p.N & 3) == 0 &&
                                           (p.ldc & 3) == 0 &&
                                           (reinterpret_cast<uintptr_t>(p.c_f32) & 15) == 0;
    [[maybe_unused]] const bool bulk_bf16 = EPI == EPI_BF16 && !(p.no_bulk_epi & 2) && (p.N & 7) == 0 && (p.ldc & 7) == 0 &&
                                            (reinterpret_cast<uintptr_t>(p.c_bf16) & 15) == 0 &&
                                            (p.mask == nullptr || (reinterpret_cast<uintptr_t>(p.mask) & 15) == 0);
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
      const TileCoord tc = coord(tile);
      const int m_blk = tc.m_blk, n_blk = tc.n_blk;
      const int as = it & 1;
      if (kSgdPrefetch) prefetch_sgd_tile(tile + tile_step);
      mbar_wait(&tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
      const int row0 = m_blk * BLOCK_M + q * 32;
      if constexpr (EPI == EPI_SGD) {
        static_assert(32 * (BN + 1) * 4 <= S::kEpiStageBytes, "SGD slab must fit the epilogue staging buffer");
        const float lr_eff = p.lr_dev != nullptr ? p.lr * __ldg(p.lr_dev) : p.lr;
#pragma unroll 1
        for (int qq = 0; qq < 4; ++qq) {
          if (q == qq) {                                  // the two warps that own this TMEM lane quadrant
#pragma unroll 1
            for (int c = half; c < BN / 32; c += kEpiWarps / 4) {
              uint32_t r[32];
              tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) epi_stage[lane * (BN + 1) + c * 32 + j] = __uint_as_float(r[j]) * p.alpha;
            }
          }
          epi_bar_sync();
          sgd_rows_from_slab<BN>(p, epi_stage, e, lane, m_blk * BLOCK_M + qq * 32, n_blk * BN, lr_eff);
          epi_bar_sync();                                 // slab free for the next quadrant
        }
      } else {
        bool stored = false;
        if constexpr (EPI == EPI_F32) {
          // Plain fp32 output (inner-product weight gradients, the SFB reconstruct in its two-pass form): pure HBM
          // streaming — a 128 x BN tile is 128 KB of output for as little as 4 k-blocks of MMA.  The per-warp 32x32 walk
          // of epilogue_tile32 issues isolated 128-byte stores and measured 33 % of the HBM copy rate
          // (profiles/r1_roofline.md).  Here a TMEM lane quadrant (32 rows x BN columns) is laid down in ONE shared slab
          // (row pitch BN + 4 floats: 16-byte aligned rows, conflict-free float4 stores from the lane = row layout) and
          // leaves as 32 bulk-async (TMA) row stores of up to 1 KB each; the issuing lanes only wait until the copy
          // engine has READ the slab, so the global writes of one quadrant overlap the fill of the next.
          if (bulk_f32) {
            constexpr int LDS = BN + 4;
            static_assert(32 * LDS * 4 <= S::kEpiStageBytes, "fp32 slab must fit the epilogue staging buffer");
            constexpr int RW = 32 / kEpiWarps;
            const int col0 = n_blk * BN;
            const int ncols = min(BN, p.N - col0);
#pragma unroll 1
            for (int qq = 0; qq < 4; ++qq) {
              if (q == qq) {
#pragma unroll 1
                for (int c = half; c < BN / 32; c += kEpiWarps / 4) {
                  uint32_t r[32];
                  tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c * 32, r);
                  tmem_ld_wait();
                  float4* dst = reinterpret_cast<float4*>(epi_stage + lane * LDS + c * 32);
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    dst[j] = make_float4(__uint_as_float(r[4 * j]) * p.alpha, __uint_as_float(r[4 * j + 1]) * p.alpha,
                                         __uint_as_float(r[4 * j + 2]) * p.alpha, __uint_as_float(r[4 * j + 3]) * p.alpha);
                }
                fence_proxy_async_smem();               // generic-proxy writes -> visible to the bulk-copy engine
              }
              epi_bar_sync();
              if (lane < RW) {
                const int rr = e * RW + lane;
                const int row = m_blk * BLOCK_M + qq * 32 + rr;
                if (row < p.M && ncols > 0)
                  bulk_store_row(p.c_f32 + static_cast<long>(row) * p.ldc + col0, epi_stage + rr * LDS,
                                 static_cast<uint32_t>(ncols) * 4u);
                bulk_commit_wait_read();                 // the slab may be overwritten once the engine has read it
              }
              epi_bar_sync();
            }
            stored = true;
          }
        }
        if constexpr (EPI == EPI_BF16) {
          // bf16 activations / data gradients: the same idea with TWO half-size slabs (32 rows x BN bf16, row pitch
          // BN * 2 + 16 bytes: conflict-free 16-byte stores from the lane = row layout).  Bias, (leaky) ReLU and the
          // producer's ReLU mask are applied in registers in that layout — the mask row segment of a lane is 64
          // contiguous bytes.  Quadrant qq + 1 is converted into the other slab while the copy engine still reads
          // quadrant qq; one CTA-wide barrier per quadrant, rows leave as up-to-512-byte bulk stores instead of the
          // 64-byte pieces of the per-warp walk.  MEASURED SLOWER than the walk in every network (AlexNet 3.65 vs
          // 3.27 ms, VGG-16 12.7 vs 9.97 ms: only two of the eight epilogue warps convert at a time) — kept behind
          // set_bulk_epilogue(3) with its numerics test, OFF by default (profiles/r2_conv_experiments.md, E7).
          if (bulk_bf16) {
            constexpr int LDB = BN * 2 + 16;
            static_assert(2 * 32 * LDB <= S::kEpiStageBytes, "two bf16 slabs must fit the epilogue staging buffer");
            constexpr int RW = 32 / kEpiWarps;
            uint8_t* slabs = reinterpret_cast<uint8_t*>(epi_stage);
            const int col0 = n_blk * BN;
            const int ncols = min(BN, p.N - col0);
#pragma unroll 1
            for (int qq = 0; qq < 4; ++qq) {
              uint8_t* slab = slabs + (qq & 1) * (32 * LDB);
              if (q == qq) {
                const long grow = static_cast<long>(m_blk) * BLOCK_M + qq * 32 + lane;
#pragma unroll 1
                for (int c = half; c < BN / 32; c += kEpiWarps / 4) {
                  const int cb = col0 + c * 32;
                  if (cb >= p.N) break;                    // warp-uniform: no valid column in this chunk
                  uint32_t r[32];
                  tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c * 32, r);
                  uint4 mk[4];
                  const bool use_mask = p.mask != nullptr;
                  if (use_mask) {
                    // (1.0, 1.0) = keep; rows past M and 8-column groups past N are never stored
                    const uint4* mrow = reinterpret_cast<const uint4*>(p.mask + (grow < p.M ? grow : 0) * p.ldc + cb);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      mk[k] = (cb + 8 * k < p.N) ? __ldg(mrow + k) : make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
                  }
                  tmem_ld_wait();
                  uint32_t out[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    float x0 = __uint_as_float(r[2 * j]) * p.alpha, x1 = __uint_as_float(r[2 * j + 1]) * p.alpha;
                    if (p.bias != nullptr) {               // warp-uniform addresses: one broadcast transaction each
                      x0 += __ldg(p.bias + min(cb + 2 * j, p.N - 1));
                      x1 += __ldg(p.bias + min(cb + 2 * j + 1, p.N - 1));
                    }
                    if (p.relu) {
                      x0 = x0 > 0.f ? x0 : x0 * p.relu_slope;
                      x1 = x1 > 0.f ? x1 : x1 * p.relu_slope;
                    }
                    if (use_mask) {
                      const uint32_t m = reinterpret_cast<const uint32_t*>(mk)[j];
                      if (!(__uint_as_float(m << 16) > 0.f)) x0 *= p.relu_slope;
                      if (!(__uint_as_float(m & 0xffff0000u) > 0.f)) x1 *= p.relu_slope;
                    }
                    const __nv_bfloat162 v = __floats2bfloat162_rn(x0, x1);
                    out[j] = *reinterpret_cast<const uint32_t*>(&v);
                  }
                  uint4* dst = reinterpret_cast<uint4*>(slab + lane * LDB + c * 64);
#pragma unroll
                  for (int k = 0; k < 4; ++k) dst[k] = make_uint4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
                }
                fence_proxy_async_smem();                 // generic-proxy writes -> visible to the bulk-copy engine
              }
              // every bulk group issued so far has been read (the other slab, quadrant qq - 1): after the barrier
              // below that slab is free for quadrant qq + 1
              if (lane < RW) bulk_wait_read_all();
              epi_bar_sync();
              if (lane < RW) {
                const int rr = e * RW + lane;
                const long row = static_cast<long>(m_blk) * BLOCK_M + qq * 32 + rr;
                if (row < p.M && ncols > 0)
                  bulk_store_row(p.c_bf16 + row * p.ldc + col0, slab + rr * LDB, static_cast<uint32_t>(ncols) * 2u);
                bulk_commit();
              }
            }
            stored = true;
          }
        }
        if (!stored) {
#pragma unroll 1
          for (int c = half; c < BN / 32; c += kEpiWarps / 4) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c * 32, r);
            tmem_ld_wait();
            epilogue_tile32<EPI>(p, r, epi_stage + e * (32 * 33), lane, row0, n_blk * BN + c * 32);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[as]);
        else mbar_arrive_remote(&tmem_empty[as], 0);      // the pair's accumulator buffer is released on the leader
      }
    }
    if constexpr (EPI != EPI_SGD) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // bulk row stores have landed
  }

  tc_fence_before();
  __syncthreads();
  if (C > 1 || CG == 2) cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer may still touch it
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, S::kTmemCols);
    else tmem_dealloc(tmem_base, S::kTmemCols);
  }
}

}  // namespace psd
