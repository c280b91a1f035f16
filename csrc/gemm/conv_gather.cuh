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

// Per-thread decoded row position m -> pointer to tap (0,0) + its input coordinates; computed once per tile
// (fprop / dgrad: the rows of a tile do not change along K) or once per k-block (wgrad: the reduction runs over m).
struct RowPos {
  const __nv_bfloat16* p00;   // address of input element (n, ih0, iw0, 0) — may point outside the image (never dereferenced then)
  int ih0, iw0;               // input coordinates of tap (0,0); ih0 = INT_MIN/2 marks an invalid (out-of-range) row
  int lim;                    // ROW mode: valid elements of a kernel row inside the image row
};

// The geometry fields the inner loops need, hoisted into registers once per role (the asm memory clobbers of
// cp.async would otherwise make the compiler re-read the kernel-parameter bank on every use).
struct GatherRegs {
  const __nv_bfloat16* x;
  int H, W, S, dr, Cg, Lp, K, OHOW, OW, sh, sw, off_h, off_w, L, mode;
  long pitch, img;            // pixel pitch, elements per image
  long M;
  int row_step;               // elements between consecutive input rows (W * pitch)
  FastDiv div_ow, div_ohow, div_cg, div_s, div_lp;
};
__device__ __forceinline__ GatherRegs load_gather_regs(const ConvGeom& g) {
  GatherRegs r;
  r.x = g.x; r.H = g.H; r.W = g.W; r.S = g.S; r.dr = g.dr; r.Cg = g.Cg; r.Lp = g.Lp; r.K = g.K;
  r.OHOW = g.OH * g.OW; r.OW = g.OW; r.sh = g.sh; r.sw = g.sw; r.off_h = g.off_h; r.off_w = g.off_w; r.L = g.L;
  r.mode = g.mode; r.pitch = g.pitch; r.img = static_cast<long>(g.H) * g.W * g.pitch; r.M = g.M;
  r.row_step = static_cast<int>(g.W * g.pitch);
  r.div_ow = g.div_ow; r.div_ohow = g.div_ohow; r.div_cg = g.div_cg; r.div_s = g.div_s; r.div_lp = g.div_lp;
  return r;
}

__device__ __forceinline__ RowPos decode_row(const GatherRegs& g, long m) {
  RowPos p;
  if (m >= g.M) {
    p.p00 = g.x; p.ih0 = -(1 << 29); p.iw0 = -(1 << 29); p.lim = 0;
    return p;
  }
  const uint32_t mm = static_cast<uint32_t>(m);
  const uint32_t n = fdiv(mm, g.div_ohow);
  const uint32_t rem = mm - n * static_cast<uint32_t>(g.OHOW);
  const uint32_t oh = fdiv(rem, g.div_ow);
  const uint32_t ow = rem - oh * static_cast<uint32_t>(g.OW);
  p.ih0 = static_cast<int>(oh) * g.sh + g.off_h;
  p.iw0 = static_cast<int>(ow) * g.sw + g.off_w;
  p.p00 = g.x + static_cast<long>(n) * g.img + static_cast<long>(p.ih0) * g.row_step + static_cast<long>(p.iw0) * g.pitch;
  p.lim = min(g.L, (g.W - p.iw0) * g.Cg);
  return p;
}

// Per-thread, per-k-block chunk decode with everything the row loop needs precomputed.
struct ChunkOff {
  int dh, dw;        // tap displacement in input rows / columns (already multiplied by the direction)
  int off;           // element offset from the row's p00 to this chunk
  int c;             // ROW mode: element offset inside the padded kernel row
  bool in_k;
};
__device__ __forceinline__ ChunkOff decode_chunk_off(const GatherRegs& g, int k) {
  ChunkOff t;
  t.in_k = k < g.K;
  if (g.mode == 0) {
    const uint32_t tap = fdiv(static_cast<uint32_t>(k), g.div_cg);
    const int c = k - static_cast<int>(tap) * g.Cg;
    const int r = static_cast<int>(fdiv(tap, g.div_s));
    const int s = static_cast<int>(tap) - r * g.S;
    t.dh = r * g.dr;
    t.dw = s * g.dr;
    t.off = t.dh * g.row_step + t.dw * static_cast<int>(g.pitch) + c;
    t.c = c;
  } else {
    const int r = static_cast<int>(fdiv(static_cast<uint32_t>(k), g.div_lp));
    t.c = k - r * g.Lp;
    t.dh = r;
    t.dw = 0;
    t.off = r * g.row_step + t.c;
  }
  return t;
}

// One 16-byte chunk of one row: ~8 integer instructions + the cp.async.
__device__ __forceinline__ void gather_chunk_tap(const GatherRegs& g, uint32_t dst, const RowPos& pos, const ChunkOff& t) {
  const int ih = pos.ih0 + t.dh, iw = pos.iw0 + t.dw;
  const bool ok = t.in_k && static_cast<unsigned>(ih) < static_cast<unsigned>(g.H) &&
                  static_cast<unsigned>(iw) < static_cast<unsigned>(g.W);
  cp_async_16(dst, ok ? pos.p00 + t.off : g.x, ok ? 16u : 0u);
}
__device__ __forceinline__ void gather_chunk_row(const GatherRegs& g, uint32_t dst, const RowPos& pos, const ChunkOff& t) {
  const int ih = pos.ih0 + t.dh;
  int valid = min(8, pos.lim - t.c);
  if (!(t.in_k && static_cast<unsigned>(ih) < static_cast<unsigned>(g.H)) || valid < 0) valid = 0;
  cp_async_16(dst, valid > 0 ? pos.p00 + t.off : g.x, static_cast<uint32_t>(valid * 2));
}

}  // namespace psd
