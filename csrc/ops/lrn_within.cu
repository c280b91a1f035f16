// Within-channel LRN and stochastic pooling for the sm100 engine (NHWC bf16, 8 channels per thread, fp32 math).
//
// WITHIN_CHANNEL LRN — the reference composes Split -> Power(x^2) -> AVE pool(size, pad (size-1)/2) -> Power(1 + alpha*s)^-beta
// -> Eltwise PROD (src/caffe/layers/lrn_layer.cpp:20-69, 159-165): five layers, five HBM round trips.  Here:
//   forward   scale = 1 + alpha * (sum_{window} x^2) / pool_size ;  y = x * scale^-beta             (one kernel)
//   backward  r_i = dy_i * y_i / (scale_i * pool_size_i)                                             (kernel 1, fp32 r and p = scale^-beta)
//             dx_j = dy_j * p_j - 2 * alpha * beta * x_j * sum_{i : j in window(i)} r_i              (kernel 2)
// pool_size follows Caffe's AVE pooling: the window is clipped to the PADDED extent for the divisor and to the image for
// the sum (src/caffe/layers/pooling_layer.cu:50-78).
//
// STOCHASTIC pooling (src/caffe/layers/pooling_layer.cu:81-150, 295-330; no padding):
//   train  pick the first window element whose running sum reaches u * sum(window), u ~ U(0,1) per output element;
//          the tap index is stored exactly like MAX pooling's arg-max, so the backward is pool_bwd's MAX gather
//   test   y = sum(x^2) / sum(x)
// The uniform numbers come from a counter-based hash of (seed, device-resident iteration counter, element index) — the
// same generator as the dropout kernel, so a captured CUDA graph draws fresh numbers on every replay.
#include "nhwc_common.cuh"

namespace psd {

struct WinGeom {
  int N, C, H, W, size, pre;
  long pitch_x, pitch_y;
};

__device__ __forceinline__ float pow_neg_beta(float scale, float beta) { return fast_ex2(-beta * fast_lg2(scale)); }

// window of output pixel (h, w): rows [hs, he) x cols [ws, we) clipped to the image; pool_size from the padded extent
__device__ __forceinline__ void lrn_window(const WinGeom& g, int h, int w, int& hs, int& he, int& ws, int& we, float& inv_ps) {
  hs = h - g.pre; ws = w - g.pre;
  const int hpe = min(hs + g.size, g.H + g.pre), wpe = min(ws + g.size, g.W + g.pre);
  inv_ps = 1.f / static_cast<float>((hpe - hs) * (wpe - ws));
  he = min(hpe, g.H); we = min(wpe, g.W);
  hs = max(hs, 0); ws = max(ws, 0);
}

// MODE 0: forward (out = y).  MODE 1: backward pass 1 (r, p fp32 out; needs dy).
template <int MODE>
__global__ void __launch_bounds__(256)
lrn_within_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ y,
                  float* __restrict__ r, float* __restrict__ p, WinGeom g, float alpha, float beta, long total8, long pitch_dy) {
  const int c8n = g.C >> 3;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long pix = i / c8n;
    const int c0 = static_cast<int>(i - pix * c8n) * 8;
    const int w = static_cast<int>(pix % g.W);
    const long t = pix / g.W;
    const int h = static_cast<int>(t % g.H);
    const long n = t / g.H;
    int hs, he, ws, we;
    float inv_ps;
    lrn_window(g, h, w, hs, he, ws, we, inv_ps);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, v[8];
    const __nv_bfloat16* xb = x + n * g.H * g.W * g.pitch_x + c0;
    for (int hh = hs; hh < he; ++hh)
      for (int ww = ws; ww < we; ++ww) {
        unpack8(ld8(xb + static_cast<long>(hh * g.W + ww) * g.pitch_x), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[k] * v[k];
      }
    unpack8(ld8(xb + static_cast<long>(h * g.W + w) * g.pitch_x), v);
    if constexpr (MODE == 0) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = v[k] * pow_neg_beta(1.f + alpha * acc[k] * inv_ps, beta);
      st8(y + pix * g.pitch_y + c0, pack8(o));
    } else {
      float gy[8];
      unpack8(ld8(dy + pix * pitch_dy + c0), gy);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float scale = 1.f + alpha * acc[k] * inv_ps;
        const float pk = pow_neg_beta(scale, beta);
        p[pix * g.C + c0 + k] = pk;
        r[pix * g.C + c0 + k] = gy[k] * v[k] * pk / scale * inv_ps;       // dy * y / (scale * pool_size)
      }
    }
  }
}

__global__ void __launch_bounds__(256)
lrn_within_bwd2_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, const float* __restrict__ r,
                       const float* __restrict__ p, __nv_bfloat16* __restrict__ dx, WinGeom g, float alpha, float beta, long total8,
                       long pitch_dy) {
  const int c8n = g.C >> 3;
  const float k2 = 2.f * alpha * beta;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long pix = i / c8n;
    const int c0 = static_cast<int>(i - pix * c8n) * 8;
    const int w = static_cast<int>(pix % g.W);
    const long t = pix / g.W;
    const int h = static_cast<int>(t % g.H);
    const long n = t / g.H;
    // output pixels i whose window contains (h, w):  i - pre <= h < i - pre + size  (window end clipped to H + pre, which
    // never excludes an in-image pixel)
    const int is = max(h + g.pre - g.size + 1, 0), ie = min(h + g.pre, g.H - 1);
    const int js = max(w + g.pre - g.size + 1, 0), je = min(w + g.pre, g.W - 1);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* rb = r + n * g.H * g.W * g.C + c0;
    for (int ii = is; ii <= ie; ++ii)
      for (int jj = js; jj <= je; ++jj) {
        const float4 a = *reinterpret_cast<const float4*>(rb + static_cast<long>(ii * g.W + jj) * g.C);
        const float4 b = *reinterpret_cast<const float4*>(rb + static_cast<long>(ii * g.W + jj) * g.C + 4);
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
        acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
      }
    float xv[8], gy[8], o[8];
    unpack8(ld8(x + pix * g.pitch_x + c0), xv);
    unpack8(ld8(dy + pix * pitch_dy + c0), gy);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = gy[k] * p[pix * g.C + c0 + k] - k2 * xv[k] * acc[k];
    st8(dx + pix * g.pitch_y + c0, pack8(o));
  }
}

static WinGeom win_geom(const NhwcView& v, int size, long pitch_out) {
  WinGeom g;
  g.N = v.N; g.C = v.C; g.H = v.H; g.W = v.W; g.size = size; g.pre = (size - 1) / 2;
  g.pitch_x = v.pitch; g.pitch_y = pitch_out;
  return g;
}

at::Tensor lrn_within_fwd(const at::Tensor& x, int64_t size, double alpha, double beta) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16, "lrn_within_fwd: bf16 CUDA tensor expected");
  c10::cuda::CUDAGuard guard(x.device());
  NhwcView v = nhwc_view(x);
  TORCH_CHECK(v.C % 8 == 0 && v.pitch % 8 == 0, "lrn_within: channels must be a multiple of 8");
  at::Tensor y = empty_nhwc(v.N, v.C, v.H, v.W, x.options());
  const long total8 = static_cast<long>(v.N) * v.H * v.W * (v.C / 8);
  lrn_within_kernel<0><<<grid_for(total8, 256), 256, 0, at::cuda::getCurrentCUDAStream()>>>(
      reinterpret_cast<const __nv_bfloat16*>(x.data_ptr()), nullptr, reinterpret_cast<__nv_bfloat16*>(y.data_ptr()), nullptr,
      nullptr, win_geom(v, static_cast<int>(size), v.C), static_cast<float>(alpha), static_cast<float>(beta), total8, 0);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return y;
}

at::Tensor lrn_within_bwd(const at::Tensor& x, const at::Tensor& dy, int64_t size, double alpha, double beta) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && dy.scalar_type() == at::kBFloat16);
  c10::cuda::CUDAGuard guard(x.device());
  NhwcView v = nhwc_view(x), vd = nhwc_view(dy);
  TORCH_CHECK(vd.N == v.N && vd.C == v.C && vd.H == v.H && vd.W == v.W, "lrn_within_bwd: shape mismatch");
  TORCH_CHECK(v.C % 8 == 0 && v.pitch % 8 == 0 && vd.pitch % 8 == 0, "lrn_within: channels must be a multiple of 8");
  const long pixels = static_cast<long>(v.N) * v.H * v.W;
  at::Tensor r = at::empty({pixels, v.C}, x.options().dtype(at::kFloat)), p = at::empty_like(r);
  at::Tensor dx = empty_nhwc(v.N, v.C, v.H, v.W, x.options());
  const long total8 = pixels * (v.C / 8);
  auto st = at::cuda::getCurrentCUDAStream();
  WinGeom g = win_geom(v, static_cast<int>(size), v.C);
  lrn_within_kernel<1><<<grid_for(total8, 256), 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x.data_ptr()), reinterpret_cast<const __nv_bfloat16*>(dy.data_ptr()), nullptr,
      r.data_ptr<float>(), p.data_ptr<float>(), g, static_cast<float>(alpha), static_cast<float>(beta), total8, vd.pitch);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  lrn_within_bwd2_kernel<<<grid_for(total8, 256), 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x.data_ptr()), reinterpret_cast<const __nv_bfloat16*>(dy.data_ptr()),
      r.data_ptr<float>(), p.data_ptr<float>(), reinterpret_cast<__nv_bfloat16*>(dx.data_ptr()), g, static_cast<float>(alpha),
      static_cast<float>(beta), total8, vd.pitch);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return dx;
}

// ------------------------------------------------------------------------------------------------ stochastic pooling
__device__ __forceinline__ uint32_t mix32(uint64_t z) {      // splitmix64 finaliser -> 32 uniform bits
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<uint32_t>(z >> 16);
}

struct StoGeom {
  int N, C, H, W, OH, OW, kh, kw, sh, sw;
  long pitch_x;
};

template <bool TRAIN>
__global__ void __launch_bounds__(256)
stochastic_pool_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ idx, StoGeom g,
                       long total8, uint64_t seed, const long* __restrict__ iter_dev) {
  const int c8n = g.C >> 3;
  const uint64_t it = iter_dev != nullptr ? static_cast<uint64_t>(*iter_dev) : 0ull;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total8; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long opix = i / c8n;
    const int c0 = static_cast<int>(i - opix * c8n) * 8;
    const int ow = static_cast<int>(opix % g.OW);
    const long t = opix / g.OW;
    const int oh = static_cast<int>(t % g.OH);
    const long n = t / g.OH;
    const int hs = oh * g.sh, ws = ow * g.sw;
    const int he = min(hs + g.kh, g.H), we = min(ws + g.kw, g.W);
    const __nv_bfloat16* xb = x + n * g.H * g.W * g.pitch_x + c0;
    float sum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, sq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, v[8];
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) {
        unpack8(ld8(xb + static_cast<long>(h * g.W + w) * g.pitch_x), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) { sum[k] += v[k]; sq[k] += v[k] * v[k]; }
      }
    float o[8];
    if constexpr (!TRAIN) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = sum[k] > 0.f ? sq[k] / sum[k] : 0.f;
    } else {
      float thr[8], run[8];
      alignas(8) uint8_t tap[8];
      bool done[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t u = mix32(seed ^ (it * 0x9E3779B97F4A7C15ull) ^ (static_cast<uint64_t>(opix * g.C + c0 + k) * 0xD1B54A32D192ED03ull));
        thr[k] = (static_cast<float>(u >> 8) * (1.f / 16777216.f)) * sum[k];
        run[k] = 0.f; done[k] = false; o[k] = 0.f;
        tap[k] = static_cast<uint8_t>((he - 1 - hs) * g.kw + (we - 1 - ws));     // default: the last element of the window
      }
      for (int h = hs; h < he; ++h)
        for (int w = ws; w < we; ++w) {
          unpack8(ld8(xb + static_cast<long>(h * g.W + w) * g.pitch_x), v);
          const bool last = (h == he - 1) && (w == we - 1);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            run[k] += v[k];
            if (!done[k] && (run[k] >= thr[k] || last)) {
              done[k] = true;
              o[k] = v[k];
              tap[k] = static_cast<uint8_t>((h - hs) * g.kw + (w - ws));
            }
          }
        }
      if (idx != nullptr) *reinterpret_cast<uint2*>(idx + opix * g.C + c0) = *reinterpret_cast<const uint2*>(tap);
    }
    st8(y + opix * g.C + c0, pack8(o));
  }
}

// returns (y, idx) — idx (uint8 tap index, pool_bwd's MAX format) only in training
std::tuple<at::Tensor, at::Tensor> stochastic_pool_fwd(const at::Tensor& x, at::IntArrayRef k, at::IntArrayRef s, int64_t oh,
                                                       int64_t ow, bool train, int64_t seed,
                                                       const c10::optional<at::Tensor>& iter_dev) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16, "stochastic_pool: bf16 CUDA tensor expected");
  c10::cuda::CUDAGuard guard(x.device());
  NhwcView v = nhwc_view(x);
  TORCH_CHECK(v.C % 8 == 0 && v.pitch % 8 == 0, "stochastic_pool: channels must be a multiple of 8");
  TORCH_CHECK(k[0] * k[1] <= 255, "stochastic_pool: window too large for the uint8 tap index");
  StoGeom g{v.N, v.C, v.H, v.W, static_cast<int>(oh), static_cast<int>(ow), static_cast<int>(k[0]), static_cast<int>(k[1]),
            static_cast<int>(s[0]), static_cast<int>(s[1]), v.pitch};
  at::Tensor y = empty_nhwc(v.N, v.C, oh, ow, x.options());
  const long opix = static_cast<long>(v.N) * oh * ow;
  at::Tensor idx = train ? at::empty({opix, v.C}, x.options().dtype(at::kByte)) : at::empty({0}, x.options().dtype(at::kByte));
  const long total8 = opix * (v.C / 8);
  const long* itp = iter_dev.has_value() ? reinterpret_cast<const long*>(iter_dev->data_ptr()) : nullptr;
  auto st = at::cuda::getCurrentCUDAStream();
  if (train)
    stochastic_pool_kernel<true><<<grid_for(total8, 256), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(x.data_ptr()), reinterpret_cast<__nv_bfloat16*>(y.data_ptr()),
        idx.data_ptr<uint8_t>(), g, total8, static_cast<uint64_t>(seed), itp);
  else
    stochastic_pool_kernel<false><<<grid_for(total8, 256), 256, 0, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(x.data_ptr()), reinterpret_cast<__nv_bfloat16*>(y.data_ptr()), nullptr, g, total8,
        0ull, nullptr);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return {y, idx};
}

}  // namespace psd

TORCH_LIBRARY_FRAGMENT(poseidon, m) {
  m.def("lrn_within_fwd(Tensor x, int size, float alpha, float beta) -> Tensor", &psd::lrn_within_fwd);
  m.def("lrn_within_bwd(Tensor x, Tensor dy, int size, float alpha, float beta) -> Tensor", &psd::lrn_within_bwd);
  m.def("stochastic_pool_fwd(Tensor x, int[] k, int[] s, int oh, int ow, bool train, int seed, Tensor? iter_dev) -> (Tensor, Tensor)",
        &psd::stochastic_pool_fwd);
}
