// MAX / AVE pooling, NHWC bf16, 8 channels per thread (16-byte accesses), Caffe window semantics
// (ceil output size computed by the caller; AVE divisor = window clipped to the padded extent).
//
// MAX forward stores the arg-max as the in-window tap index (uint8) so backward is a pure gather:
// every input pixel sums dy over the output windows whose recorded tap points at it — no atomics,
// deterministic.  AVE backward is the matching gather.
//
// reference: src/caffe/layers/pooling_layer.cu:12-47 (MaxPoolForward), :50-78 (AvePoolForward),
// :213-256 (MaxPoolBackward), :259-293 (AvePoolBackward).
#include "nhwc_common.cuh"

namespace psd {

struct PoolGeom {
  int N, C, H, W, OH, OW, kh, kw, sh, sw, ph, pw;
  long xpitch, ypitch;   // pixel pitches of the input-side and output-side tensors
};

// KT > 0: compile-time square window (fully unrolled, all window loads issued before use); KT == 0: generic.
template <bool MAXP, int KT>
__global__ void __launch_bounds__(256)
pool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx,
                PoolGeom g) {
  const int c8 = g.C / 8;
  const long total = static_cast<long>(g.N) * g.OH * g.OW * c8;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % c8);
    long t = i / c8;
    const int ow = static_cast<int>(t % g.OW); t /= g.OW;
    const int oh = static_cast<int>(t % g.OH);
    const int n = static_cast<int>(t / g.OH);
    int hs = oh * g.sh - g.ph, ws = ow * g.sw - g.pw;
    const int he_pad = min(hs + g.kh, g.H + g.ph), we_pad = min(ws + g.kw, g.W + g.pw);
    const int he = min(he_pad, g.H), we = min(we_pad, g.W);
    const int h0 = max(hs, 0), w0 = max(ws, 0);
    float acc[8];
    int best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j] = MAXP ? -3.402823466e38f : 0.f; best[j] = 0; }
    const __nv_bfloat16* xb = x + (static_cast<long>(n) * g.H * g.W) * g.xpitch + v * 8;
    if constexpr (KT > 0) {
      bf16x8 win[KT * KT];
      bool ok[KT * KT];
#pragma unroll
      for (int dh = 0; dh < KT; ++dh) {
#pragma unroll
        for (int dw = 0; dw < KT; ++dw) {
          const int h = hs + dh, w = ws + dw;
          ok[dh * KT + dw] = h >= 0 && h < g.H && w >= 0 && w < g.W;
          if (ok[dh * KT + dw]) win[dh * KT + dw] = ld8(xb + (static_cast<long>(h) * g.W + w) * g.xpitch);
        }
      }
#pragma unroll
      for (int t = 0; t < KT * KT; ++t) {
        if (!ok[t]) continue;
        float f[8];
        unpack8(win[t], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (MAXP) {
            if (f[j] > acc[j]) { acc[j] = f[j]; best[j] = t; }
          } else {
            acc[j] += f[j];
          }
        }
      }
    } else {
      for (int h = h0; h < he; ++h) {
        for (int w = w0; w < we; ++w) {
          float f[8];
          unpack8(ld8(xb + (static_cast<long>(h) * g.W + w) * g.xpitch), f);
          const int tap = (h - hs) * g.kw + (w - ws);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (MAXP) {
              if (f[j] > acc[j]) { acc[j] = f[j]; best[j] = tap; }
            } else {
              acc[j] += f[j];
            }
          }
        }
      }
    }
    if (!MAXP) {
      const float inv = 1.f / static_cast<float>((he_pad - hs) * (we_pad - ws));
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= inv;
    }
    const long opix = (static_cast<long>(n) * g.OH + oh) * g.OW + ow;
    st8(y + opix * g.ypitch + v * 8, pack8(acc));
    if (MAXP && idx != nullptr) {
      uint2 pk;
      pk.x = best[0] | (best[1] << 8) | (best[2] << 16) | (best[3] << 24);
      pk.y = best[4] | (best[5] << 8) | (best[6] << 16) | (best[7] << 24);
      *reinterpret_cast<uint2*>(idx + opix * g.C + v * 8) = pk;
    }
  }
}

// NC > 0: at most NC x NC output windows cover an input pixel (NC = ceil(k / stride)), loops unrolled with
// predicates so every dy / index load is in flight before the accumulation; NC == 0: generic.
template <bool MAXP, int NC>
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx, __nv_bfloat16* __restrict__ dx,
                PoolGeom g) {
  const int c8 = g.C / 8;
  const long total = static_cast<long>(g.N) * g.H * g.W * c8;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % c8);
    long t = i / c8;
    const int w = static_cast<int>(t % g.W); t /= g.W;
    const int h = static_cast<int>(t % g.H);
    const int n = static_cast<int>(t / g.H);
    // output windows that cover (h, w)
    const int hp = h + g.ph, wp = w + g.pw;
    const int oh0 = hp < g.kh ? 0 : (hp - g.kh) / g.sh + 1;
    const int oh1 = min(hp / g.sh + 1, g.OH);
    const int ow0 = wp < g.kw ? 0 : (wp - g.kw) / g.sw + 1;
    const int ow1 = min(wp / g.sw + 1, g.OW);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    auto accumulate = [&](int oh, int ow, const bf16x8& dv, const uint2& pk) {
      float d[8];
      unpack8(dv, d);
      const int hs = oh * g.sh - g.ph, ws = ow * g.sw - g.pw;
      if (MAXP) {
        const int tap = (h - hs) * g.kw + (w - ws);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int b = ((j < 4 ? pk.x : pk.y) >> (8 * (j & 3))) & 0xff;
          if (b == tap) acc[j] += d[j];
        }
      } else {
        const int he_pad = min(hs + g.kh, g.H + g.ph), we_pad = min(ws + g.kw, g.W + g.pw);
        const float inv = 1.f / static_cast<float>((he_pad - hs) * (we_pad - ws));
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += d[j] * inv;
      }
    };
    if constexpr (NC > 0) {
      bf16x8 dv[NC * NC];
      uint2 pk[NC * NC];
      bool ok[NC * NC];
#pragma unroll
      for (int a = 0; a < NC; ++a) {
#pragma unroll
        for (int b = 0; b < NC; ++b) {
          const int oh = oh0 + a, ow = ow0 + b;
          const int t = a * NC + b;
          ok[t] = oh < oh1 && ow < ow1;
          pk[t] = make_uint2(0, 0);
          if (ok[t]) {
            const long opix = (static_cast<long>(n) * g.OH + oh) * g.OW + ow;
            dv[t] = ld8(dy + opix * g.ypitch + v * 8);
            if (MAXP) pk[t] = *reinterpret_cast<const uint2*>(idx + opix * g.C + v * 8);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < NC; ++a) {
#pragma unroll
        for (int b = 0; b < NC; ++b)
          if (ok[a * NC + b]) accumulate(oh0 + a, ow0 + b, dv[a * NC + b], pk[a * NC + b]);
      }
    } else {
      for (int oh = oh0; oh < oh1; ++oh) {
        for (int ow = ow0; ow < ow1; ++ow) {
          const long opix = (static_cast<long>(n) * g.OH + oh) * g.OW + ow;
          const bf16x8 dv = ld8(dy + opix * g.ypitch + v * 8);
          uint2 pk = make_uint2(0, 0);
          if (MAXP) pk = *reinterpret_cast<const uint2*>(idx + opix * g.C + v * 8);
          accumulate(oh, ow, dv, pk);
        }
      }
    }
    const long ipix = (static_cast<long>(n) * g.H + h) * g.W + w;
    st8(dx + ipix * g.xpitch + v * 8, pack8(acc));
  }
}

static PoolGeom make_geom(const NhwcView& x, int64_t oh, int64_t ow, at::IntArrayRef k, at::IntArrayRef s,
                          at::IntArrayRef p) {
  PoolGeom g;
  g.N = x.N; g.C = x.C; g.H = x.H; g.W = x.W; g.OH = oh; g.OW = ow;
  g.kh = k[0]; g.kw = k[1]; g.sh = s[0]; g.sw = s[1]; g.ph = p[0]; g.pw = p[1];
  g.xpitch = x.pitch;
  g.ypitch = x.C;
  return g;
}

std::tuple<at::Tensor, at::Tensor> pool_fwd(const at::Tensor& x, bool is_max, at::IntArrayRef k, at::IntArrayRef s,
                                            at::IntArrayRef p, int64_t oh, int64_t ow, bool want_idx) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16);
  c10::cuda::CUDAGuard guard(x.device());
  NhwcView xv = nhwc_view(x);
  TORCH_CHECK(xv.C % 8 == 0, "pool: channels must be a multiple of 8");
  TORCH_CHECK(k[0] * k[1] <= 255, "pool: window too large for uint8 tap index");
  at::Tensor y = empty_nhwc(xv.N, xv.C, oh, ow, x.options());
  at::Tensor idx = (is_max && want_idx) ? at::empty({xv.N, oh, ow, xv.C}, x.options().dtype(at::kByte)) : at::Tensor();
  PoolGeom g = make_geom(xv, oh, ow, k, s, p);
  const long total = static_cast<long>(g.N) * g.OH * g.OW * (g.C / 8);
  auto stream = at::cuda::getCurrentCUDAStream();
  auto xp = reinterpret_cast<const __nv_bfloat16*>(x.data_ptr());
  auto yp = reinterpret_cast<__nv_bfloat16*>(y.data_ptr());
  uint8_t* ip = idx.defined() ? idx.data_ptr<uint8_t>() : nullptr;
  const int grid = grid_for(total, 256, 148 * 32);
  const int kt = (g.kh == g.kw && (g.kh == 2 || g.kh == 3)) ? g.kh : 0;
#define PSD_PF(MX, KT) pool_fwd_kernel<MX, KT><<<grid, 256, 0, stream>>>(xp, yp, MX ? ip : nullptr, g)
  if (is_max) { if (kt == 3) PSD_PF(true, 3); else if (kt == 2) PSD_PF(true, 2); else PSD_PF(true, 0); }
  else        { if (kt == 3) PSD_PF(false, 3); else if (kt == 2) PSD_PF(false, 2); else PSD_PF(false, 0); }
#undef PSD_PF
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {y, idx.defined() ? idx : at::empty({0}, x.options().dtype(at::kByte))};
}

at::Tensor pool_bwd(const at::Tensor& dy, const at::Tensor& idx, bool is_max, at::IntArrayRef in_hw, at::IntArrayRef k,
                    at::IntArrayRef s, at::IntArrayRef p) {
  TORCH_CHECK(dy.is_cuda() && dy.scalar_type() == at::kBFloat16);
  c10::cuda::CUDAGuard guard(dy.device());
  NhwcView dv = nhwc_view(dy);
  at::Tensor dx = empty_nhwc(dv.N, dv.C, in_hw[0], in_hw[1], dy.options());
  NhwcView xv = nhwc_view(dx);
  PoolGeom g = make_geom(xv, dv.H, dv.W, k, s, p);
  g.ypitch = dv.pitch;
  const long total = static_cast<long>(g.N) * g.H * g.W * (g.C / 8);
  auto stream = at::cuda::getCurrentCUDAStream();
  auto dyp = reinterpret_cast<const __nv_bfloat16*>(dy.data_ptr());
  auto dxp = reinterpret_cast<__nv_bfloat16*>(dx.data_ptr());
  const int grid = grid_for(total, 256, 148 * 32);
  const int ncand = std::max((g.kh + g.sh - 1) / g.sh, (g.kw + g.sw - 1) / g.sw);
  const int nc = ncand <= 3 ? ncand : 0;
  const uint8_t* ip = nullptr;
  if (is_max) {
    TORCH_CHECK(idx.numel() == dy.numel(), "pool_bwd: index tensor missing");
    ip = idx.data_ptr<uint8_t>();
  }
#define PSD_PB(MX, NCV) pool_bwd_kernel<MX, NCV><<<grid, 256, 0, stream>>>(dyp, ip, dxp, g)
  if (is_max) { if (nc == 1) PSD_PB(true, 1); else if (nc == 2) PSD_PB(true, 2); else if (nc == 3) PSD_PB(true, 3); else PSD_PB(true, 0); }
  else        { if (nc == 1) PSD_PB(false, 1); else if (nc == 2) PSD_PB(false, 2); else if (nc == 3) PSD_PB(false, 3); else PSD_PB(false, 0); }
#undef PSD_PB
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return dx;
}

}  // namespace psd

TORCH_LIBRARY_FRAGMENT(poseidon, m) {
  m.def("pool_fwd(Tensor x, bool is_max, int[] k, int[] s, int[] p, int oh, int ow, bool want_idx) -> (Tensor, Tensor)",
        &psd::pool_fwd);
  m.def("pool_bwd(Tensor dy, Tensor idx, bool is_max, int[] in_hw, int[] k, int[] s, int[] p) -> Tensor", &psd::pool_bwd);
}
