// Implicit-GEMM convolution operand gather (im2col never materialised).
//
// Four producer warps (128 threads) build, directly in the 128B-swizzled shared-memory layout the UMMA
// descriptors expect, 64-element K slices of the virtual im2col matrix
//        A[m, k],   m = (n, oh, ow),   k = (r, s, c)          (NHWC input, channels innermost)
// with 16-byte cp.async copies (zero-fill for padding taps / tails).  Two addressing modes:
//   TAP : C_g % 8 == 0.  A 16-byte chunk is 8 channels of one tap; per-tap bounds checks.
//   ROW : first layers (C padded to 4): the S*C elements of one kernel row are contiguous in memory, so
//         K is organised as [R][Lp] (Lp = S*C rounded up to 8; the pad multiplies zero weights).
//
// Thread mapping (coalescing): thread t owns chunk column j = t & 7 and rows (t >> 3) + 16*i.  A warp-level
// cp.async therefore covers 4 rows x 128 contiguous bytes (4-8 L1 wavefronts) instead of 32 scattered 16-byte
// pieces (32 wavefronts) — the difference between a gather-bound and an MMA-bound main loop.  The k -> (r,s,c)
// split is per thread per k-block (same for all its rows); the m -> (n,oh,ow) split uses precomputed magic
// multipliers.
//
// The same routine feeds fprop (A operand, K-major: 128 rows x 64 k), dgrad (gathers dY with mirrored taps)
// and wgrad (B operand, MN-major: 64 reduction rows x 64 k-columns per chunk).
//
// Replaces im2col_gpu_kernel/col2im_gpu_kernel + per-image cublasSgemm loops
// (reference: src/caffe/util/im2col.cu:12-132, src/caffe/layers/conv_layer.cu:13-119).
#pragma once
#include "sm100_prims.cuh"

namespace psd {

struct FastDiv {
  uint32_t mul, shift, d;
};
inline FastDiv make_fastdiv(uint32_t d) {
  // exact for all 32-bit n when computed with a 64-bit product (n < 2^31 here)
  FastDiv f;
  f.d = d;
  if (d == 1) { f.mul = 0; f.shift = 0; return f; }
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.shift = s;
  f.mul = static_cast<uint32_t>(((1ull << (32 + s)) + d - 1) / d - (1ull << 32));
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  if (f.d == 1) return n;
  const uint32_t hi = __umulhi(n, f.mul);
  return (hi + ((n - hi) >> 1)) >> (f.shift - 1);
}

struct ConvGeom {
  const __nv_bfloat16* x;   // gathered tensor (activations for fprop/wgrad, dY for dgrad), group offset applied
  int N, H, W;              // its spatial extent
  long pitch;               // pixel pitch in elements
  int Cg;                   // channels per group (TAP) / padded channels (ROW)
  int OH, OW;               // extent of the row index m = (n, oh, ow)
  int R, S, sh, sw;
  int off_h, off_w;         // ih = oh*sh + off_h + r*dr  (fprop: off=-pad, dr=+1 ; dgrad: off=+pad, dr=-1)
  int dr;
  int mode;                 // 0 TAP, 1 ROW
  int L, Lp;                // ROW: valid / padded elements per kernel row
  int K;                    // total reduction length (multiple of 8)
  long M;                   // N*OH*OW
  FastDiv div_ow, div_ohow, div_cg, div_s, div_lp;
};

// Per-thread, per-k-block decode of this thread's chunk column.
struct ChunkTap {
  int r, s, c;      // TAP: tap row / col / first channel ; ROW: r = kernel row, c = element offset in the padded row
  bool in_k;
};
__device__ __forceinline__ ChunkTap decode_chunk(const ConvGeom& g, int k) {
  ChunkTap t;
  t.in_k = k < g.K;
  if (g.mode == 0) {
    const uint32_t tap = fdiv(static_cast<uint32_t>(k), g.div_cg);
    t.c = k - static_cast<int>(tap) * g.Cg;
    t.r = static_cast<int>(fdiv(tap, g.div_s));
    t.s = static_cast<int>(tap) - t.r * g.S;
  } else {
    t.r = static_cast<int>(fdiv(static_cast<uint32_t>(k), g.div_lp));
    t.c = k - t.r * g.Lp;
    t.s = 0;
  }
  return t;
}

// Per-thread decoded row position m -> (image base, ih0, iw0); computed once per tile (fprop / dgrad: the rows of
// a tile do not change along K) or once per k-block (wgrad: the reduction runs over m).
struct RowPos {
  long base;        // element offset of image n
  int ih0, iw0;     // input coordinates of tap (0,0)
  int lim;          // ROW mode: valid elements of a kernel row inside the image row; TAP: unused
  bool valid;
};
__device__ __forceinline__ RowPos decode_row(const ConvGeom& g, long m) {
  RowPos p;
  p.valid = m < g.M;
  const uint32_t mm = p.valid ? static_cast<uint32_t>(m) : 0u;
  const uint32_t n = fdiv(mm, g.div_ohow);
  const uint32_t rem = mm - n * static_cast<uint32_t>(g.OH * g.OW);
  const uint32_t oh = fdiv(rem, g.div_ow);
  const uint32_t ow = rem - oh * static_cast<uint32_t>(g.OW);
  p.base = static_cast<long>(n) * g.H * g.W * g.pitch;
  p.ih0 = static_cast<int>(oh) * g.sh + g.off_h;
  p.iw0 = static_cast<int>(ow) * g.sw + g.off_w;
  p.lim = min(g.L, (g.W - p.iw0) * g.Cg);
  return p;
}

// Copy this thread's 16-byte chunk of the row at `pos` (or zeros) to `dst`: no divisions on this path.
__device__ __forceinline__ void gather_chunk(const ConvGeom& g, uint32_t dst, const RowPos& pos, const ChunkTap& t) {
  const __nv_bfloat16* src = g.x;
  uint32_t bytes = 0;
  const int ih = pos.ih0 + t.r * g.dr;
  if (g.mode == 0) {
    const int iw = pos.iw0 + t.s * g.dr;
    if (pos.valid && t.in_k && static_cast<unsigned>(ih) < static_cast<unsigned>(g.H) &&
        static_cast<unsigned>(iw) < static_cast<unsigned>(g.W)) {
      src = g.x + pos.base + (static_cast<long>(ih) * g.W + iw) * g.pitch + t.c;
      bytes = 16;
    }
  } else {
    if (pos.valid && t.in_k && static_cast<unsigned>(ih) < static_cast<unsigned>(g.H)) {
      const int valid = max(0, min(8, pos.lim - t.c));
      if (valid > 0) {
        src = g.x + pos.base + (static_cast<long>(ih) * g.W + pos.iw0) * g.pitch + t.c;
        bytes = static_cast<uint32_t>(valid * 2);
      }
    }
  }
  cp_async_16(dst, src, bytes);
}

}  // namespace psd
