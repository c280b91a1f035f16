// Implicit-GEMM convolution operand gather (im2col never materialised).
//
// Four producer warps (128 threads) build, directly in the 128B-swizzled shared-memory layout the UMMA
// descriptors expect, 64-element K slices of the virtual im2col matrix
//        A[m, k],   m = (n, oh, ow),   k = (r, s, c)          (NHWC input, channels innermost)
// with 16-byte cp.async copies (zero-fill for padding taps / tails).  Two addressing modes:
//   TAP : C_g % 8 == 0.  A 16-byte chunk is 8 channels of one tap; per-tap bounds checks.
//   ROW : first layers (C padded to 4): the S*C elements of one kernel row are contiguous in memory, so
//         K is organised as [R][Lp] (Lp = S*C rounded up to 8; the pad multiplies zero weights).
// The same routine feeds fprop (A operand, K-major: 128 rows x 64 k), dgrad (gathers dY with mirrored
// taps) and wgrad (B operand, MN-major: 64 reduction rows x 64 k-columns per chunk).
//
// Replaces im2col_gpu_kernel/col2im_gpu_kernel + per-image cublasSgemm loops
// (reference: src/caffe/util/im2col.cu:12-132, src/caffe/layers/conv_layer.cu:13-119).
#pragma once
#include "sm100_prims.cuh"

namespace psd {

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
};

// Fill `nrows` rows x 64 k-elements (128 B each, swizzled) at smem `dst` for rows m0.. and k-range k0..k0+63.
// Called by 128 threads with (row, chunk-group) assignment passed in: thread handles row `row` (< nrows)
// and all 8 chunks of it.
__device__ __forceinline__ void gather_row(const ConvGeom& g, uint32_t dst_row_addr, int row_in_tile, long m, int k0) {
  const uint32_t sw = static_cast<uint32_t>(row_in_tile & 7);
  if (m >= g.M) {
#pragma unroll
    for (int j = 0; j < 8; ++j) cp_async_16(dst_row_addr + ((j ^ sw) << 4), g.x, 0);
    return;
  }
  const int ow = static_cast<int>(m % g.OW);
  const long t = m / g.OW;
  const int oh = static_cast<int>(t % g.OH);
  const int n = static_cast<int>(t / g.OH);
  const int ih0 = oh * g.sh + g.off_h, iw0 = ow * g.sw + g.off_w;
  const __nv_bfloat16* img = g.x + static_cast<long>(n) * g.H * g.W * g.pitch;
  if (g.mode == 0) {
    int tap = k0 / g.Cg;
    int c = k0 - tap * g.Cg;
    int r = tap / g.S;
    int s = tap - r * g.S;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ih = ih0 + r * g.dr, iw = iw0 + s * g.dr;
      const bool ok = (k0 + j * 8 < g.K) && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W;
      const __nv_bfloat16* src = ok ? img + (static_cast<long>(ih) * g.W + iw) * g.pitch + c : g.x;
      cp_async_16(dst_row_addr + ((j ^ sw) << 4), src, ok ? 16u : 0u);
      c += 8;
      if (c >= g.Cg) { c = 0; if (++s == g.S) { s = 0; ++r; } }
    }
  } else {
    int r = k0 / g.Lp;
    int e = k0 - r * g.Lp;          // element offset inside the padded kernel row
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ih = ih0 + r;
      // elements [e, e+8) of the row starting at pixel iw0; valid while e < L and inside the image row
      int valid = 0;
      if (k0 + j * 8 < g.K && ih >= 0 && ih < g.H) {
        const int row_elems = (g.W - iw0) * g.Cg;            // elements left in this image row
        const int lim = min(g.L, row_elems);
        valid = max(0, min(8, lim - e));
      }
      const __nv_bfloat16* src = valid > 0 ? img + (static_cast<long>(ih) * g.W + iw0) * g.pitch + e : g.x;
      cp_async_16(dst_row_addr + ((j ^ sw) << 4), src, static_cast<uint32_t>(valid * 2));
      e += 8;
      if (e >= g.Lp) { e = 0; ++r; }
    }
  }
}

}  // namespace psd
