// Fused gradient all-reduce + optimizer step over NVLink peer memory — the replacement for the whole
// Bösen push/pull path (worker D2H -> oplog -> ZeroMQ -> server add -> ServerPushRow -> H2D).
//
//   world == 1 : fused_update         W,H (+bf16 shadow) <- step(W, H, G)      one pass, all rules
//   world  > 1 : allreduce_sgd        every rank owns 1/P of each bucket ("two-shot"):
//        phase 0  signal "my gradients for epoch e are in my symmetric G buffer", wait for all peers
//        phase 1  g = sum_p G_p[slice]  (P2P vector loads, or one multimem.ld_reduce through the NVSwitch)
//                 apply weight decay + momentum/Nesterov/AdaGrad on the owned slice (history is sharded)
//                 store the new fp32 weights (+ bf16 shadow) into EVERY rank's W (P2P stores / multimem.st)
//        phase 2  last CTA: signal "slice written / done reading", wait for all peers
//      small buckets use the one-shot variant (every rank reduces the full bucket, no weight broadcast).
//   No NCCL call, no host round trip, one kernel per bucket launched from the backward hooks (DWBP).
//
// reference: src/caffe/solver.cpp:455-473 (ThreadSyncWithPS), :815-892 (ComputeUpdateValue: 4-5 cuBLAS L1
// passes per blob), src/caffe/blob.cpp:208-286 (UpdatePSTable / SyncWithPSTable via host memory).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

#include "../gemm/sm100_prims.cuh"

namespace psd {

constexpr int kMaxRanks = 8;

struct UpdateHyper {
  float lr, momentum, decay, delta, gscale;
  int rule, l1;
};

__device__ __forceinline__ float step_rule(float g, float& w, float& h, const UpdateHyper& p) {
  g *= p.gscale;
  if (p.decay != 0.f) g += p.decay * (p.l1 ? (w > 0.f ? 1.f : (w < 0.f ? -1.f : 0.f)) : w);
  float step;
  if (p.rule == 0) {
    h = p.lr * g + p.momentum * h;
    step = h;
  } else if (p.rule == 1) {
    const float h_old = h;
    h = p.lr * g + p.momentum * h;
    step = (1.f + p.momentum) * h - p.momentum * h_old;
  } else {
    h = h + g * g;
    step = p.lr * g / (sqrtf(h) + p.delta);
  }
  w -= step;
  return w;
}

__device__ __forceinline__ uint2 pack_bf16x4(float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  return u;
}

// ------------------------------------------------------------------------------------ single GPU
// REARM: the gradient buffer is a persistent accumulation target (conv wgrad adds into it with split-K atomics): leave it
// zeroed for the next step instead of paying a memset launch per layer and step.
template <bool REARM>
__global__ void __launch_bounds__(256)
fused_update_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ h, __nv_bfloat16* __restrict__ wb,
                    long n, UpdateHyper hp, const float* __restrict__ lr_dev) {
  if (lr_dev != nullptr) hp.lr *= __ldg(lr_dev);     // global learning rate lives on the device (CUDA-graph replay safe)
  const long n4 = n >> 2;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float4 wv = reinterpret_cast<float4*>(w)[i], hv = reinterpret_cast<float4*>(h)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    if (REARM) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    step_rule(gv.x, wv.x, hv.x, hp);
    step_rule(gv.y, wv.y, hv.y, hp);
    step_rule(gv.z, wv.z, hv.z, hp);
    step_rule(gv.w, wv.w, hv.w, hp);
    reinterpret_cast<float4*>(w)[i] = wv;
    reinterpret_cast<float4*>(h)[i] = hv;
    if (wb != nullptr) reinterpret_cast<uint2*>(wb)[i] = pack_bf16x4(wv);
  }
  // tail
  for (long i = (n4 << 2) + blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float wv = w[i], hv = h[i];
    step_rule(g[i], wv, hv, hp);
    if (REARM) g[i] = 0.f;
    w[i] = wv;
    h[i] = hv;
    if (wb != nullptr) wb[i] = __float2bfloat16(wv);
  }
}

static UpdateHyper make_hyper(double lr, double momentum, double decay, int64_t rule, bool l1, double delta, double gscale) {
  UpdateHyper hp;
  hp.lr = static_cast<float>(lr);
  hp.momentum = static_cast<float>(momentum);
  hp.decay = static_cast<float>(decay);
  hp.delta = static_cast<float>(delta);
  hp.gscale = static_cast<float>(gscale);
  hp.rule = static_cast<int>(rule);
  hp.l1 = l1;
  return hp;
}

// Same element order in memory: equal sizes, and equal strides on every dimension that has extent > 1
// (size-1 dimensions carry arbitrary strides, e.g. 1x1 convolution weights in channels-last form).
static bool same_dense_layout(const at::Tensor& a, const at::Tensor& b) {
  if (a.numel() != b.numel() || a.sizes() != b.sizes()) return false;
  for (int64_t d = 0; d < a.dim(); ++d)
    if (a.size(d) > 1 && a.stride(d) != b.stride(d)) return false;
  return true;
}

// W, G, H: fp32 tensors with identical (dense) layout; wb: optional bf16 shadow in the same storage order.
void fused_update(at::Tensor w, at::Tensor g, at::Tensor h, c10::optional<at::Tensor> wb, double lr, double momentum,
                  double decay, int64_t rule, bool l1, double delta, double gscale, const c10::optional<at::Tensor>& lr_dev,
                  bool rearm) {
  TORCH_CHECK(w.is_cuda() && w.scalar_type() == at::kFloat && g.scalar_type() == at::kFloat && h.scalar_type() == at::kFloat);
  TORCH_CHECK(same_dense_layout(w, g) && same_dense_layout(w, h), "fused_update: W, G, H must share one dense layout");
  TORCH_CHECK(w.is_non_overlapping_and_dense(), "fused_update: dense tensors expected");
  c10::cuda::CUDAGuard guard(w.device());
  const long n = w.numel();
  __nv_bfloat16* wbp = nullptr;
  if (wb.has_value()) {
    TORCH_CHECK(wb->scalar_type() == at::kBFloat16 && wb->numel() == n && wb->is_non_overlapping_and_dense());
    wbp = reinterpret_cast<__nv_bfloat16*>(wb->data_ptr());
  }
  const bool aligned = (reinterpret_cast<uintptr_t>(w.data_ptr()) | reinterpret_cast<uintptr_t>(g.data_ptr()) |
                        reinterpret_cast<uintptr_t>(h.data_ptr())) % 16 == 0 &&
                       (wbp == nullptr || reinterpret_cast<uintptr_t>(wbp) % 8 == 0);
  TORCH_CHECK(aligned, "fused_update: 16-byte aligned buffers expected");
  const int grid = static_cast<int>(std::max<long>(1, std::min<long>((n / 4 + 255) / 256, 148 * 8)));
  const UpdateHyper hp = make_hyper(lr, momentum, decay, rule, l1, delta, gscale);
  const float* lrp = lr_dev.has_value() ? lr_dev->data_ptr<float>() : nullptr;
  auto st = at::cuda::getCurrentCUDAStream();
  if (rearm) fused_update_kernel<true><<<grid, 256, 0, st>>>(w.data_ptr<float>(), g.data_ptr<float>(), h.data_ptr<float>(), wbp, n, hp, lrp);
  else fused_update_kernel<false><<<grid, 256, 0, st>>>(w.data_ptr<float>(), g.data_ptr<float>(), h.data_ptr<float>(), wbp, n, hp, lrp);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------ single GPU, many tensors, one launch
// GoogLeNet steps 128 parameter tensors, most of them a few KB: at ~3 us per launch the optimizer was 7 % of the step
// (profiles/r2_kernels_googlenet_1gpu.txt).  One launch walks up to kMaxMulti tensors: every block owns a fixed-size
// chunk of one tensor (block -> tensor through a prefix table in the kernel parameters).
constexpr int kMaxMulti = 48;
constexpr int kMultiChunk4 = 2048;          // float4 per block-chunk (32 KB of weights)
struct MultiSeg {
  float* w[kMaxMulti];
  float* g[kMaxMulti];
  float* h[kMaxMulti];
  __nv_bfloat16* wb[kMaxMulti];
  long n[kMaxMulti];
  float lr[kMaxMulti], decay[kMaxMulti];
  int first_block[kMaxMulti + 1];
  unsigned char rearm[kMaxMulti];
  int count;
};

__global__ void __launch_bounds__(256)
fused_update_multi_kernel(const __grid_constant__ MultiSeg ms, UpdateHyper hp, const float* __restrict__ lr_dev) {
  const float lr_glob = lr_dev != nullptr ? __ldg(lr_dev) : 1.f;
  int s = 0;
  while (s + 1 < ms.count && static_cast<int>(blockIdx.x) >= ms.first_block[s + 1]) ++s;
  hp.lr = ms.lr[s] * lr_glob;
  hp.decay = ms.decay[s];
  float* __restrict__ w = ms.w[s];
  float* __restrict__ g = ms.g[s];
  float* __restrict__ h = ms.h[s];
  __nv_bfloat16* __restrict__ wb = ms.wb[s];
  const bool rearm = ms.rearm[s] != 0;
  const long n = ms.n[s], n4 = n >> 2;
  const long lo = static_cast<long>(blockIdx.x - ms.first_block[s]) * kMultiChunk4;
  const long hi = min(n4, lo + kMultiChunk4);
  for (long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    float4 wv = reinterpret_cast<float4*>(w)[i], hv = reinterpret_cast<float4*>(h)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    if (rearm) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    step_rule(gv.x, wv.x, hv.x, hp);
    step_rule(gv.y, wv.y, hv.y, hp);
    step_rule(gv.z, wv.z, hv.z, hp);
    step_rule(gv.w, wv.w, hv.w, hp);
    reinterpret_cast<float4*>(w)[i] = wv;
    reinterpret_cast<float4*>(h)[i] = hv;
    if (wb != nullptr) reinterpret_cast<uint2*>(wb)[i] = pack_bf16x4(wv);
  }
  if (hi == n4) {                               // the tensor's last chunk also takes its (< 4 element) tail
    for (long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      float wv = w[i], hv = h[i];
      step_rule(g[i], wv, hv, hp);
      if (rearm) g[i] = 0.f;
      w[i] = wv;
      h[i] = hv;
      if (wb != nullptr) wb[i] = __float2bfloat16(wv);
    }
  }
}

// ws / gs / hs: fp32 tensors of pairwise identical dense layout, 16-byte aligned; wbs[i] may be undefined (numel 0).
void fused_update_multi(std::vector<at::Tensor> ws, std::vector<at::Tensor> gs, std::vector<at::Tensor> hs,
                        std::vector<at::Tensor> wbs, std::vector<double> lrs, std::vector<double> decays,
                        std::vector<int64_t> rearms, double momentum, int64_t rule, bool l1, double delta, double gscale,
                        const c10::optional<at::Tensor>& lr_dev) {
  const size_t total = ws.size();
  TORCH_CHECK(gs.size() == total && hs.size() == total && wbs.size() == total && lrs.size() == total && decays.size() == total &&
              rearms.size() == total, "fused_update_multi: list lengths differ");
  if (total == 0) return;
  c10::cuda::CUDAGuard guard(ws[0].device());
  auto st = at::cuda::getCurrentCUDAStream();
  const float* lrp = lr_dev.has_value() ? lr_dev->data_ptr<float>() : nullptr;
  const UpdateHyper hp = make_hyper(1.0, momentum, 0.0, rule, l1, delta, gscale);
  for (size_t base = 0; base < total; base += kMaxMulti) {
    MultiSeg ms{};
    ms.count = static_cast<int>(std::min<size_t>(kMaxMulti, total - base));
    int blocks = 0;
    for (int i = 0; i < ms.count; ++i) {
      const at::Tensor &w = ws[base + i], &g = gs[base + i], &h = hs[base + i], &wb = wbs[base + i];
      TORCH_CHECK(w.is_cuda() && w.scalar_type() == at::kFloat && g.scalar_type() == at::kFloat && h.scalar_type() == at::kFloat);
      TORCH_CHECK(same_dense_layout(w, g) && same_dense_layout(w, h) && w.is_non_overlapping_and_dense(),
                  "fused_update_multi: W, G, H must share one dense layout");
      ms.w[i] = w.data_ptr<float>(); ms.g[i] = g.data_ptr<float>(); ms.h[i] = h.data_ptr<float>();
      ms.wb[i] = wb.numel() ? reinterpret_cast<__nv_bfloat16*>(wb.data_ptr()) : nullptr;
      if (wb.numel()) TORCH_CHECK(wb.scalar_type() == at::kBFloat16 && wb.numel() == w.numel());
      TORCH_CHECK((reinterpret_cast<uintptr_t>(ms.w[i]) | reinterpret_cast<uintptr_t>(ms.g[i]) | reinterpret_cast<uintptr_t>(ms.h[i])) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(ms.wb[i]) % 8 == 0, "fused_update_multi: 16-byte aligned buffers expected");
      ms.n[i] = w.numel();
      ms.lr[i] = static_cast<float>(lrs[base + i]);
      ms.decay[i] = static_cast<float>(decays[base + i]);
      ms.rearm[i] = rearms[base + i] ? 1 : 0;
      ms.first_block[i] = blocks;
      blocks += static_cast<int>(std::max<long>(1, ((w.numel() >> 2) + kMultiChunk4 - 1) / kMultiChunk4));
    }
    ms.first_block[ms.count] = blocks;
    fused_update_multi_kernel<<<blocks, 256, 0, st>>>(ms, hp, lrp);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
  }
}

// ------------------------------------------------------------------------------------ multi GPU
struct PeerPtrs {
  const float* g[kMaxRanks];      // every rank's gradient bucket (symmetric)
  float* w[kMaxRanks];            // every rank's fp32 master weights for this bucket (symmetric)
  __nv_bfloat16* wb[kMaxRanks];   // every rank's bf16 shadow (may be null)
  uint32_t* flags[kMaxRanks];     // every rank's flag block for this bucket: [2][kMaxRanks] u32
  const float* g_mc;              // multicast (NVLS) address of the gradient bucket, or null
  float* w_mc;                    // multicast address of the weights, or null
};

__device__ __forceinline__ float4 multimem_ld_reduce_f32x4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f32x4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// Cross-rank barrier on this bucket's flag block.  phase 0/1 use separate words so epochs never alias.
__device__ __forceinline__ void peer_barrier_flags(uint32_t* const* flags, int rank, int world, int phase, uint32_t epoch) {
  if (threadIdx.x < world) {
    // tell peer t that `rank` reached `epoch`
    st_release_sys(flags[threadIdx.x] + phase * kMaxRanks + rank, epoch);
  }
  if (threadIdx.x < world) wait_flag_ge(flags[rank] + phase * kMaxRanks + threadIdx.x, epoch);
  __syncthreads();
}
__device__ __forceinline__ void peer_barrier(const PeerPtrs& pp, int rank, int world, int phase, uint32_t epoch) {
  peer_barrier_flags(pp.flags, rank, world, phase, epoch);
}

template <bool ONE_SHOT>
__global__ void __launch_bounds__(512)
allreduce_sgd_kernel(PeerPtrs pp, float* __restrict__ h, long n, int rank, int world, uint32_t epoch, UpdateHyper hp,
                     unsigned int* __restrict__ done_counter, const float* __restrict__ lr_dev,
                     const uint32_t* __restrict__ epoch_dev) {
  if (lr_dev != nullptr) hp.lr *= __ldg(lr_dev);
  // the step counter lives in device memory (bumped on this stream once per iteration) so that a captured CUDA graph
  // of the whole training step sees a fresh epoch on every replay; `epoch` is then the offset inside the step
  if (epoch_dev != nullptr) epoch += *reinterpret_cast<const volatile uint32_t*>(epoch_dev);
  // ---- phase 0: all ranks' gradients for this epoch are in place
  peer_barrier(pp, rank, world, 0, epoch);

  const long n4 = n >> 2;           // buckets are padded to multiples of 4 floats by the arena
  long lo = 0, hi = n4;
  if (!ONE_SHOT) {
    const long per = (n4 + world - 1) / world;
    lo = min(n4, per * rank);
    hi = min(n4, lo + per);
  }
  for (long i = lo + blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < hi;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    float4 gv;
    if (pp.g_mc != nullptr) {
      gv = multimem_ld_reduce_f32x4(pp.g_mc + 4 * i);
    } else {
      // ONE_SHOT: fixed order 0..world-1 (every rank sums the same bucket: identical order = bit-identical replicas);
      // two-shot: staggered from the own rank so the links are used evenly (one owner per shard: order is irrelevant)
      gv = reinterpret_cast<const float4*>(pp.g[ONE_SHOT ? 0 : rank])[i];
#pragma unroll 1
      for (int q = 1; q < world; ++q) {
        const int p = ONE_SHOT ? q : (rank + q) % world;
        const float4 t = reinterpret_cast<const float4*>(pp.g[p])[i];
        gv.x += t.x; gv.y += t.y; gv.z += t.z; gv.w += t.w;
      }
    }
    float4 wv = reinterpret_cast<float4*>(pp.w[rank])[i];
    float4 hv = reinterpret_cast<float4*>(h)[i];
    step_rule(gv.x, wv.x, hv.x, hp);
    step_rule(gv.y, wv.y, hv.y, hp);
    step_rule(gv.z, wv.z, hv.z, hp);
    step_rule(gv.w, wv.w, hv.w, hp);
    reinterpret_cast<float4*>(h)[i] = hv;
    if (ONE_SHOT) {
      reinterpret_cast<float4*>(pp.w[rank])[i] = wv;
      if (pp.wb[rank] != nullptr) reinterpret_cast<uint2*>(pp.wb[rank])[i] = pack_bf16x4(wv);
    } else {
      const uint2 b = pack_bf16x4(wv);
      if (pp.w_mc != nullptr) {
        multimem_st_f32x4(pp.w_mc + 4 * i, wv);
      } else {
#pragma unroll 1
        for (int q = 0; q < world; ++q) reinterpret_cast<float4*>(pp.w[(rank + q) % world])[i] = wv;
      }
#pragma unroll 1
      for (int q = 0; q < world; ++q) {
        const int p = (rank + q) % world;
        if (pp.wb[p] != nullptr) reinterpret_cast<uint2*>(pp.wb[p])[i] = b;
      }
    }
  }
  // ---- phase 1: everyone finished reading my G and (two-shot) writing my W
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *done_counter = 0;       // ready for the next launch on this stream
    peer_barrier(pp, rank, world, 1, epoch);
  }
}

// g_ptrs / w_ptrs / wb_ptrs / flag_ptrs: per-rank device addresses of this bucket inside each rank's symmetric arena.
void allreduce_sgd(std::vector<int64_t> g_ptrs, std::vector<int64_t> w_ptrs, std::vector<int64_t> wb_ptrs,
                   std::vector<int64_t> flag_ptrs, int64_t g_mc, int64_t w_mc, at::Tensor h, int64_t n, int64_t rank,
                   int64_t epoch, bool one_shot, at::Tensor done_counter, double lr, double momentum, double decay,
                   int64_t rule, bool l1, double delta, double gscale, int64_t max_ctas,
                   const c10::optional<at::Tensor>& lr_dev, const c10::optional<at::Tensor>& epoch_dev) {
  const int world = static_cast<int>(g_ptrs.size());
  const uint32_t* edp = epoch_dev.has_value() ? reinterpret_cast<const uint32_t*>(epoch_dev->data_ptr()) : nullptr;
  const float* lrp = lr_dev.has_value() ? lr_dev->data_ptr<float>() : nullptr;
  TORCH_CHECK(world >= 1 && world <= kMaxRanks && w_ptrs.size() == g_ptrs.size() && flag_ptrs.size() == g_ptrs.size());
  TORCH_CHECK(h.is_cuda() && h.scalar_type() == at::kFloat && h.numel() >= n && n % 4 == 0);
  TORCH_CHECK(done_counter.scalar_type() == at::kInt && done_counter.numel() >= 1);
  c10::cuda::CUDAGuard guard(h.device());
  PeerPtrs pp{};
  for (int p = 0; p < world; ++p) {
    pp.g[p] = reinterpret_cast<const float*>(g_ptrs[p]);
    pp.w[p] = reinterpret_cast<float*>(w_ptrs[p]);
    pp.wb[p] = wb_ptrs.empty() ? nullptr : reinterpret_cast<__nv_bfloat16*>(wb_ptrs[p]);
    pp.flags[p] = reinterpret_cast<uint32_t*>(flag_ptrs[p]);
  }
  pp.g_mc = reinterpret_cast<const float*>(g_mc);
  pp.w_mc = reinterpret_cast<float*>(w_mc);
  UpdateHyper hp = make_hyper(lr, momentum, decay, rule, l1, delta, gscale);
  const long work4 = one_shot ? n / 4 : (n / 4 + world - 1) / world;
  int grid = static_cast<int>(std::max<long>(1, std::min<long>((work4 + 511) / 512, max_ctas > 0 ? max_ctas : 64)));
  auto stream = at::cuda::getCurrentCUDAStream();
  auto* dc = reinterpret_cast<unsigned int*>(done_counter.data_ptr());
  if (one_shot)
    allreduce_sgd_kernel<true><<<grid, 512, 0, stream>>>(pp, h.data_ptr<float>(), n, static_cast<int>(rank), world,
                                                         static_cast<uint32_t>(epoch), hp, dc, lrp, edp);
  else
    allreduce_sgd_kernel<false><<<grid, 512, 0, stream>>>(pp, h.data_ptr<float>(), n, static_cast<int>(rank), world,
                                                          static_cast<uint32_t>(epoch), hp, dc, lrp, edp);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------ one launch per BUCKET
// Round 2: the per-tensor launch above cost two system-scope flag barriers and a separate zero_() launch per parameter
// tensor (weight and bias separately) — the fixed multi-GPU tax the 1 -> 2 GPU step showed.  This
// form reduces + steps + broadcasts up to kMaxSegs tensors of a layer behind ONE barrier pair, and re-arms the gradient
// staging inside the kernel:
//   * two-shot segments: only the OWNER of a shard ever reads it, so the owner writes zeros back over the shard in
//     every rank's G right after reading it (one multimem.st through the switch, or P2P stores) — made data-dependent on
//     the loaded value so the store cannot overtake the load;
//   * one-shot segments: every rank reads every rank's G, so each rank zeroes ITS OWN copy after the closing barrier
//     (the last CTA; these segments are <= 256 KB).
constexpr int kMaxSegs = 4;
struct SegSet {
  long g_off[kMaxSegs], w_off[kMaxSegs], wb_off[kMaxSegs];    // byte offsets inside the (symmetric) arena
  float* h[kMaxSegs];                                         // local optimizer history (sharded by rank for two-shot)
  long n[kMaxSegs];                                           // floats, multiple of 4
  float lr[kMaxSegs], decay[kMaxSegs];
  int one_shot[kMaxSegs];
  int nseg;
};
struct ArenaPtrs {
  char* base[kMaxRanks];
  char* mc;                       // multicast mapping of the arena (0 = none: P2P loads / stores)
  uint32_t* flags[kMaxRanks];     // this bucket's flag block on every rank
};

__device__ __forceinline__ void multimem_st_b64(void* mc, uint2 v) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(mc), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y))
               : "memory");
}
// 0.0f that the hardware cannot produce before `v` has been loaded
__device__ __forceinline__ float dependent_zero(float v) {
  uint32_t z;
  asm volatile("and.b32 %0, %1, 0;" : "=r"(z) : "r"(__float_as_uint(v)));
  return __uint_as_float(z);
}

__global__ void __launch_bounds__(512)
allreduce_sgd_multi_kernel(const __grid_constant__ ArenaPtrs ap, const __grid_constant__ SegSet ss, int rank, int world, uint32_t epoch,
                           UpdateHyper hp,
                           unsigned int* __restrict__ done_counter, const float* __restrict__ lr_dev,
                           const uint32_t* __restrict__ epoch_dev) {
  const float lr_glob = lr_dev != nullptr ? __ldg(lr_dev) : 1.f;
  if (epoch_dev != nullptr) epoch += *reinterpret_cast<const volatile uint32_t*>(epoch_dev);
  peer_barrier_flags(ap.flags, rank, world, 0, epoch);
  for (int s = 0; s < ss.nseg; ++s) {
    UpdateHyper h = hp;
    h.lr = ss.lr[s] * lr_glob;
    h.decay = ss.decay[s];
    const bool one = ss.one_shot[s] != 0;
    const long n4 = ss.n[s] >> 2;
    long lo = 0, hi = n4;
    if (!one) {
      const long per = (n4 + world - 1) / world;
      lo = min(n4, per * rank);
      hi = min(n4, lo + per);
    }
    float4* wl = reinterpret_cast<float4*>(ap.base[rank] + ss.w_off[s]);
    float4* hl = reinterpret_cast<float4*>(ss.h[s]);
    const bool has_wb = ss.wb_off[s] >= 0;
    for (long i = lo + blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < hi;
         i += static_cast<long>(gridDim.x) * blockDim.x) {
      float4 gv;
      if (ap.mc != nullptr) {
        gv = multimem_ld_reduce_f32x4(reinterpret_cast<const float*>(ap.mc + ss.g_off[s]) + 4 * i);
      } else {
        // fixed peer order 0..world-1: identical fp32 summation order on every rank, so one-shot replicas stay
        // bit-identical (the rank-staggered order let them drift for world > 2)
        gv = reinterpret_cast<const float4*>(ap.base[0] + ss.g_off[s])[i];
#pragma unroll 1
        for (int q = 1; q < world; ++q) {
          const float4 t = reinterpret_cast<const float4*>(ap.base[q] + ss.g_off[s])[i];
          gv.x += t.x; gv.y += t.y; gv.z += t.z; gv.w += t.w;
        }
      }
      if (!one) {
        // re-arm the shard in every rank's staging buffer (only this rank reads it)
        const float z = dependent_zero(gv.x + gv.y + gv.z + gv.w);
        const float4 zv = make_float4(z, z, z, z);
        if (ap.mc != nullptr) {
          multimem_st_f32x4(reinterpret_cast<float*>(ap.mc + ss.g_off[s]) + 4 * i, zv);
        } else {
#pragma unroll 1
          for (int q = 0; q < world; ++q) reinterpret_cast<float4*>(ap.base[(rank + q) % world] + ss.g_off[s])[i] = zv;
        }
      }
      float4 wv = wl[i], hv = hl[i];
      step_rule(gv.x, wv.x, hv.x, h);
      step_rule(gv.y, wv.y, hv.y, h);
      step_rule(gv.z, wv.z, hv.z, h);
      step_rule(gv.w, wv.w, hv.w, h);
      hl[i] = hv;
      const uint2 b = pack_bf16x4(wv);
      if (one) {
        wl[i] = wv;
        if (has_wb) reinterpret_cast<uint2*>(ap.base[rank] + ss.wb_off[s])[i] = b;
      } else if (ap.mc != nullptr) {
        multimem_st_f32x4(reinterpret_cast<float*>(ap.mc + ss.w_off[s]) + 4 * i, wv);
        if (has_wb) multimem_st_b64(reinterpret_cast<uint2*>(ap.mc + ss.wb_off[s]) + i, b);
      } else {
#pragma unroll 1
        for (int q = 0; q < world; ++q) {
          const int p = (rank + q) % world;
          reinterpret_cast<float4*>(ap.base[p] + ss.w_off[s])[i] = wv;
          if (has_wb) reinterpret_cast<uint2*>(ap.base[p] + ss.wb_off[s])[i] = b;
        }
      }
    }
  }
  // ---- everyone finished reading my G and (two-shot) writing my W
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *done_counter = 0;
    peer_barrier_flags(ap.flags, rank, world, 1, epoch);
    for (int s = 0; s < ss.nseg; ++s) {
      if (!ss.one_shot[s]) continue;
      float4* gl = reinterpret_cast<float4*>(ap.base[rank] + ss.g_off[s]);
      for (long i = threadIdx.x; i < (ss.n[s] >> 2); i += blockDim.x) gl[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

void allreduce_sgd_multi(std::vector<int64_t> base_ptrs, int64_t mc_base, std::vector<int64_t> flag_ptrs,
                         std::vector<int64_t> g_offs, std::vector<int64_t> w_offs, std::vector<int64_t> wb_offs,
                         std::vector<at::Tensor> hists, std::vector<int64_t> ns, std::vector<int64_t> one_shots,
                         std::vector<double> lrs, std::vector<double> decays, int64_t rank, int64_t epoch,
                         at::Tensor done_counter, double momentum, int64_t rule, bool l1, double delta, double gscale,
                         int64_t max_ctas, const c10::optional<at::Tensor>& lr_dev, const c10::optional<at::Tensor>& epoch_dev) {
  const int world = static_cast<int>(base_ptrs.size());
  const int nseg = static_cast<int>(g_offs.size());
  TORCH_CHECK(world >= 1 && world <= kMaxRanks && flag_ptrs.size() == base_ptrs.size(), "allreduce_sgd_multi: 1..8 ranks");
  TORCH_CHECK(nseg >= 1 && nseg <= kMaxSegs && w_offs.size() == g_offs.size() && wb_offs.size() == g_offs.size() &&
              hists.size() == g_offs.size() && ns.size() == g_offs.size() && one_shots.size() == g_offs.size() &&
              lrs.size() == g_offs.size() && decays.size() == g_offs.size(), "allreduce_sgd_multi: 1..4 segments");
  TORCH_CHECK(done_counter.scalar_type() == at::kInt && done_counter.numel() >= 1);
  c10::cuda::CUDAGuard guard(done_counter.device());
  ArenaPtrs ap{};
  for (int p = 0; p < world; ++p) {
    ap.base[p] = reinterpret_cast<char*>(base_ptrs[p]);
    ap.flags[p] = reinterpret_cast<uint32_t*>(flag_ptrs[p]);
  }
  ap.mc = reinterpret_cast<char*>(mc_base);
  SegSet ss{};
  ss.nseg = nseg;
  long work4 = 0;
  for (int s = 0; s < nseg; ++s) {
    TORCH_CHECK(hists[s].is_cuda() && hists[s].scalar_type() == at::kFloat && hists[s].numel() >= ns[s] && ns[s] % 4 == 0);
    TORCH_CHECK(g_offs[s] % 16 == 0 && w_offs[s] % 16 == 0 && (wb_offs[s] < 0 || wb_offs[s] % 8 == 0), "segment alignment");
    ss.g_off[s] = g_offs[s]; ss.w_off[s] = w_offs[s]; ss.wb_off[s] = wb_offs[s];
    ss.h[s] = hists[s].data_ptr<float>();
    ss.n[s] = ns[s];
    ss.lr[s] = static_cast<float>(lrs[s]);
    ss.decay[s] = static_cast<float>(decays[s]);
    ss.one_shot[s] = one_shots[s] ? 1 : 0;
    work4 = std::max<long>(work4, one_shots[s] ? ns[s] / 4 : (ns[s] / 4 + world - 1) / world);
  }
  UpdateHyper hp = make_hyper(1.0, momentum, 0.0, rule, l1, delta, gscale);
  const int grid = static_cast<int>(std::max<long>(1, std::min<long>((work4 + 511) / 512, max_ctas > 0 ? max_ctas : 64)));
  allreduce_sgd_multi_kernel<<<grid, 512, 0, at::cuda::getCurrentCUDAStream()>>>(
      ap, ss, static_cast<int>(rank), world, static_cast<uint32_t>(epoch), hp,
      reinterpret_cast<unsigned int*>(done_counter.data_ptr()), lr_dev.has_value() ? lr_dev->data_ptr<float>() : nullptr,
      epoch_dev.has_value() ? reinterpret_cast<const uint32_t*>(epoch_dev->data_ptr()) : nullptr);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------ bounded staleness (SSP) on the arena
// Bösen's SSP / SSPPush consistency (ps/src/petuum_ps/consistency/ssp_push_consistency_controller.cpp:70-115,
// ssp_consistency_controller.cpp:37-161; src/caffe/solver.cpp:441-443 "clock") without a server, a host thread or NCCL:
//
//   clock c, rank p, per bucket:
//     ssp_delta   Δ_p(c) = optimizer step on the LOCAL gradient (own momentum history, like the reference's workers);
//                 applied to the local weights at once (read-my-writes), stored in slot c % R of p's delta ring inside the
//                 symmetric arena (R = s + 1), gradient staging re-armed; then "ready" flag c+1 raised on every peer.
//                 Before overwriting slot c % R the kernel waits until every peer has CONSUMED Δ_p(c - R) — which is
//                 exactly the SSP blocking rule: nobody runs more than s clocks ahead of the slowest reader.
//                 The last CTA then PLANS the fold: for each peer q it must fold q's clocks <= c - s (waits for them) and
//                 may fold whatever q has published beyond that (opportunistic freshness) — one decision per launch, so all
//                 CTAs of the fold kernel act on the same clock ranges.
//     ssp_fold    W_p -= Σ_q Σ_{c' in plan(q)} Δ_q(c')  read straight from the peers' rings over NVLink (P2P loads), bf16
//                 shadow refreshed; then "consumed" counters raised on the producers.  Every delta is folded exactly once
//                 by every peer; s = 0 degenerates to BSP with summed per-worker updates.
// All clocks live in device memory (the step counter that also drives the BSP kernels), so the whole step — including the
// waits — is one CUDA graph.
constexpr int kSspReadySlot = 5, kSspConsumedSlot = 6;

struct SspSegs {
  long g_off[kMaxSegs], w_off[kMaxSegs], wb_off[kMaxSegs], d_off[kMaxSegs];   // byte offsets in the arena; d_off = slot 0 of the ring
  float* h[kMaxSegs];          // local (per-worker) optimizer history
  long n[kMaxSegs];            // floats, multiple of 4
  float lr[kMaxSegs], decay[kMaxSegs];
  int nseg;
};
struct SspState {
  uint32_t* folded;      // [kMaxRanks] clocks of peer q already folded into my weights
  uint32_t* plan;        // [kMaxRanks] fold up to (exclusive) this clock in the next ssp_fold
  uint32_t* max_lag;     // [1] largest number of a peer's clocks (<= my clock) left unfolded after planning
};

__global__ void __launch_bounds__(512)
ssp_delta_kernel(ArenaPtrs ap, SspSegs ss, long ring_stride, int rank, int world, int R, int staleness, UpdateHyper hp,
                 unsigned int* __restrict__ done_counter, const float* __restrict__ lr_dev, const uint32_t* __restrict__ clock_dev,
                 SspState st) {
  const float lr_glob = lr_dev != nullptr ? __ldg(lr_dev) : 1.f;
  const uint32_t c = *reinterpret_cast<const volatile uint32_t*>(clock_dev);
  // slot reuse: every peer has folded my clock c - R
  if (c >= static_cast<uint32_t>(R)) {
    if (threadIdx.x < world && static_cast<int>(threadIdx.x) != rank)
      wait_flag_ge(ap.flags[rank] + kSspConsumedSlot * kMaxRanks + threadIdx.x, c - R + 1);
    __syncthreads();
  }
  const long slot_off = static_cast<long>(c % R) * ring_stride;
  for (int s = 0; s < ss.nseg; ++s) {
    UpdateHyper h = hp;
    h.lr = ss.lr[s] * lr_glob;
    h.decay = ss.decay[s];
    const long n4 = ss.n[s] >> 2;
    float4* gl = reinterpret_cast<float4*>(ap.base[rank] + ss.g_off[s]);
    float4* wl = reinterpret_cast<float4*>(ap.base[rank] + ss.w_off[s]);
    uint2* bl = reinterpret_cast<uint2*>(ap.base[rank] + ss.wb_off[s]);
    float4* hl = reinterpret_cast<float4*>(ss.h[s]);
    float4* dl = reinterpret_cast<float4*>(ap.base[rank] + ss.d_off[s] + slot_off);
    for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
      const float4 gv = gl[i];
      gl[i] = make_float4(0.f, 0.f, 0.f, 0.f);                       // re-arm the gradient staging
      float4 wv = wl[i], hv = hl[i];
      const float4 w0 = wv;
      step_rule(gv.x, wv.x, hv.x, h);
      step_rule(gv.y, wv.y, hv.y, h);
      step_rule(gv.z, wv.z, hv.z, h);
      step_rule(gv.w, wv.w, hv.w, h);
      hl[i] = hv;
      wl[i] = wv;                                                     // read-my-writes
      bl[i] = pack_bf16x4(wv);
      dl[i] = make_float4(w0.x - wv.x, w0.y - wv.y, w0.z - wv.z, w0.w - wv.w);
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) *done_counter = 0;
  __threadfence_system();
  if (threadIdx.x < world && static_cast<int>(threadIdx.x) != rank) {
    const int q = threadIdx.x;
    st_release_sys(ap.flags[q] + kSspReadySlot * kMaxRanks + rank, c + 1);            // Δ_rank(c) is readable
    // plan the fold of q's deltas: clocks [0, must) are due now, anything q has published beyond is folded if it is there
    const uint32_t lo = st.folded[q];
    const uint32_t must = (c + 1 > static_cast<uint32_t>(staleness)) ? c + 1 - staleness : 0u;
    const uint32_t* ready = ap.flags[rank] + kSspReadySlot * kMaxRanks + q;
    if (must > lo) wait_flag_ge(ready, must);
    const uint32_t avail = ld_acquire_sys(ready);
    uint32_t hi = avail > lo ? avail : lo;
    if (hi > lo + static_cast<uint32_t>(R)) hi = lo + R;              // (cannot happen: q blocks on my consumption)
    st.plan[q] = hi;
    atomicMax(st.max_lag, (c + 1 > hi) ? c + 1 - hi : 0u);
  }
}

__global__ void __launch_bounds__(512)
ssp_fold_kernel(ArenaPtrs ap, SspSegs ss, long ring_stride, int rank, int world, int R, unsigned int* __restrict__ done_counter,
                SspState st, int drain) {
  __shared__ uint32_t lo_s[kMaxRanks], hi_s[kMaxRanks];
  if (threadIdx.x < kMaxRanks) {
    const int q = threadIdx.x;
    uint32_t lo = 0, hi = 0;
    if (q < world && q != rank) {
      lo = st.folded[q];
      // drain: the host barrier guarantees that every rank has published its last clock — the flag is final, so every
      // CTA reads the same value
      hi = drain ? ld_acquire_sys(ap.flags[rank] + kSspReadySlot * kMaxRanks + q) : st.plan[q];
      if (hi < lo) hi = lo;
    }
    lo_s[q] = lo;
    hi_s[q] = hi;
  }
  __syncthreads();
  bool any = false;
  for (int q = 0; q < world; ++q) any = any || hi_s[q] > lo_s[q];
  if (any) {
    for (int s = 0; s < ss.nseg; ++s) {
      const long n4 = ss.n[s] >> 2;
      float4* wl = reinterpret_cast<float4*>(ap.base[rank] + ss.w_off[s]);
      uint2* bl = reinterpret_cast<uint2*>(ap.base[rank] + ss.wb_off[s]);
      for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long>(gridDim.x) * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < world; ++q) {
          for (uint32_t cc = lo_s[q]; cc < hi_s[q]; ++cc) {
            const float4 d = reinterpret_cast<const float4*>(ap.base[q] + ss.d_off[s] + static_cast<long>(cc % R) * ring_stride)[i];
            acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
          }
        }
        float4 wv = wl[i];
        wv.x -= acc.x; wv.y -= acc.y; wv.z -= acc.z; wv.w -= acc.w;
        wl[i] = wv;
        bl[i] = pack_bf16x4(wv);
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) *done_counter = 0;
  if (threadIdx.x < world && static_cast<int>(threadIdx.x) != rank) {
    const int q = threadIdx.x;
    st.folded[q] = hi_s[q];
    st_release_sys(ap.flags[q] + kSspConsumedSlot * kMaxRanks + rank, hi_s[q]);       // q may reuse the slots I have read
  }
}

static void fill_ssp(ArenaPtrs& ap, SspSegs& ss, const std::vector<int64_t>& base_ptrs, const std::vector<int64_t>& flag_ptrs,
                     const std::vector<int64_t>& g_offs, const std::vector<int64_t>& w_offs, const std::vector<int64_t>& wb_offs,
                     const std::vector<int64_t>& d_offs, const std::vector<at::Tensor>& hists, const std::vector<int64_t>& ns,
                     const std::vector<double>& lrs, const std::vector<double>& decays, long* work4) {
  const int world = static_cast<int>(base_ptrs.size());
  const int nseg = static_cast<int>(g_offs.size());
  TORCH_CHECK(world >= 1 && world <= kMaxRanks && flag_ptrs.size() == base_ptrs.size(), "ssp: 1..8 ranks");
  TORCH_CHECK(nseg >= 1 && nseg <= kMaxSegs && w_offs.size() == g_offs.size() && wb_offs.size() == g_offs.size() &&
              d_offs.size() == g_offs.size() && ns.size() == g_offs.size(), "ssp: 1..4 segments");
  for (int p = 0; p < world; ++p) {
    ap.base[p] = reinterpret_cast<char*>(base_ptrs[p]);
    ap.flags[p] = reinterpret_cast<uint32_t*>(flag_ptrs[p]);
  }
  ap.mc = nullptr;
  ss.nseg = nseg;
  *work4 = 0;
  for (int s = 0; s < nseg; ++s) {
    TORCH_CHECK(ns[s] % 4 == 0 && g_offs[s] % 16 == 0 && w_offs[s] % 16 == 0 && wb_offs[s] % 8 == 0 && d_offs[s] % 16 == 0,
                "ssp: segment alignment");
    ss.g_off[s] = g_offs[s]; ss.w_off[s] = w_offs[s]; ss.wb_off[s] = wb_offs[s]; ss.d_off[s] = d_offs[s];
    ss.n[s] = ns[s];
    if (!hists.empty()) {
      TORCH_CHECK(hists[s].is_cuda() && hists[s].scalar_type() == at::kFloat && hists[s].numel() >= ns[s]);
      ss.h[s] = hists[s].data_ptr<float>();
      ss.lr[s] = static_cast<float>(lrs[s]);
      ss.decay[s] = static_cast<float>(decays[s]);
    }
    *work4 = std::max<long>(*work4, ns[s] / 4);
  }
}

static SspState ssp_state(at::Tensor& state) {
  TORCH_CHECK(state.is_cuda() && state.scalar_type() == at::kInt && state.numel() >= 2 * kMaxRanks + 1, "ssp: state tensor");
  auto* p = reinterpret_cast<uint32_t*>(state.data_ptr());
  return SspState{p, p + kMaxRanks, p + 2 * kMaxRanks};
}

void ssp_delta(std::vector<int64_t> base_ptrs, std::vector<int64_t> flag_ptrs, std::vector<int64_t> g_offs,
               std::vector<int64_t> w_offs, std::vector<int64_t> wb_offs, std::vector<int64_t> d_offs, int64_t ring_stride,
               std::vector<at::Tensor> hists, std::vector<int64_t> ns, std::vector<double> lrs, std::vector<double> decays,
               int64_t rank, int64_t ring, int64_t staleness, at::Tensor done_counter, at::Tensor state, double momentum,
               int64_t rule, bool l1, double delta, double gscale, int64_t max_ctas, const c10::optional<at::Tensor>& lr_dev,
               const at::Tensor& clock_dev) {
  c10::cuda::CUDAGuard guard(done_counter.device());
  ArenaPtrs ap{};
  SspSegs ss{};
  long work4 = 0;
  fill_ssp(ap, ss, base_ptrs, flag_ptrs, g_offs, w_offs, wb_offs, d_offs, hists, ns, lrs, decays, &work4);
  TORCH_CHECK(ring >= 1 && ring == staleness + 1 && ring_stride % 16 == 0, "ssp: ring = staleness + 1 slots");
  UpdateHyper hp = make_hyper(1.0, momentum, 0.0, rule, l1, delta, gscale);
  const int grid = static_cast<int>(std::max<long>(1, std::min<long>((work4 + 511) / 512, max_ctas > 0 ? max_ctas : 96)));
  ssp_delta_kernel<<<grid, 512, 0, at::cuda::getCurrentCUDAStream()>>>(
      ap, ss, ring_stride, static_cast<int>(rank), static_cast<int>(base_ptrs.size()), static_cast<int>(ring),
      static_cast<int>(staleness), hp, reinterpret_cast<unsigned int*>(done_counter.data_ptr()),
      lr_dev.has_value() ? lr_dev->data_ptr<float>() : nullptr, reinterpret_cast<const uint32_t*>(clock_dev.data_ptr()),
      ssp_state(state));
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void ssp_fold(std::vector<int64_t> base_ptrs, std::vector<int64_t> flag_ptrs, std::vector<int64_t> w_offs,
              std::vector<int64_t> wb_offs, std::vector<int64_t> d_offs, int64_t ring_stride, std::vector<int64_t> ns, int64_t rank,
              int64_t ring, at::Tensor done_counter, at::Tensor state, bool drain, int64_t max_ctas) {
  c10::cuda::CUDAGuard guard(done_counter.device());
  ArenaPtrs ap{};
  SspSegs ss{};
  long work4 = 0;
  fill_ssp(ap, ss, base_ptrs, flag_ptrs, w_offs, w_offs, wb_offs, d_offs, {}, ns, {}, {}, &work4);
  const int grid = static_cast<int>(std::max<long>(1, std::min<long>((work4 + 511) / 512, max_ctas > 0 ? max_ctas : 96)));
  ssp_fold_kernel<<<grid, 512, 0, at::cuda::getCurrentCUDAStream()>>>(
      ap, ss, ring_stride, static_cast<int>(rank), static_cast<int>(base_ptrs.size()), static_cast<int>(ring),
      reinterpret_cast<unsigned int*>(done_counter.data_ptr()), ssp_state(state), drain ? 1 : 0);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// Publish a payload (sufficient factors) into EVERY rank's arena at the same offset — one multimem.st through
// the NVSwitch when a multicast mapping exists, P2P stores otherwise — then (optionally) raise this rank's
// epoch flag on every peer.  The payload crosses NVLink exactly once per peer; consumers read it locally.
__global__ void __launch_bounds__(512)
peer_push_kernel(const uint4* __restrict__ src, PeerPtrs dst, uint4* __restrict__ dst_mc, long n16, int rank, int world,
                 int slot, uint32_t epoch, int signal, unsigned int* __restrict__ done_counter, int wait_slot,
                 const uint32_t* __restrict__ epoch_dev) {
  if (epoch_dev != nullptr) epoch += *reinterpret_cast<const volatile uint32_t*>(epoch_dev);
  if (wait_slot >= 0) {
    // single-buffered slots: every peer must have consumed the previous step's payload (it says so on MY flag block)
    if (threadIdx.x < world) wait_flag_ge(dst.flags[rank] + wait_slot * kMaxRanks + threadIdx.x, epoch - 1);
    __syncthreads();
  }
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n16; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const uint4 v = src[i];
    if (dst_mc != nullptr) {
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_mc + i),
                   "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                   : "memory");
    } else {
#pragma unroll 1
      for (int q = 0; q < world; ++q) reinterpret_cast<uint4*>(dst.w[(rank + q) % world])[i] = v;
    }
  }
  if (!signal) return;
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *done_counter = 0;
    __threadfence_system();
    if (threadIdx.x < world) st_release_sys(dst.flags[threadIdx.x] + slot * kMaxRanks + rank, epoch);
  }
}

// src: dense tensor (nbytes % 16 == 0).  dst_ptrs[p]: destination address inside rank p's arena.
void peer_push(const at::Tensor& src, std::vector<int64_t> dst_ptrs, int64_t dst_mc, std::vector<int64_t> flag_ptrs,
               int64_t rank, int64_t slot, int64_t epoch, bool signal, at::Tensor done_counter, int64_t wait_slot,
               const c10::optional<at::Tensor>& epoch_dev) {
  const int world = static_cast<int>(dst_ptrs.size());
  const uint32_t* edp = epoch_dev.has_value() ? reinterpret_cast<const uint32_t*>(epoch_dev->data_ptr()) : nullptr;
  TORCH_CHECK(world >= 1 && world <= kMaxRanks && src.is_cuda() && src.is_contiguous());
  const long nbytes = src.numel() * src.element_size();
  TORCH_CHECK(nbytes % 16 == 0 && reinterpret_cast<uintptr_t>(src.data_ptr()) % 16 == 0, "peer_push: 16-byte granularity");
  c10::cuda::CUDAGuard guard(src.device());
  PeerPtrs pp{};
  for (int p = 0; p < world; ++p) {
    pp.w[p] = reinterpret_cast<float*>(dst_ptrs[p]);
    pp.flags[p] = (signal || wait_slot >= 0) ? reinterpret_cast<uint32_t*>(flag_ptrs[p]) : nullptr;
  }
  const long n16 = nbytes / 16;
  const int grid = static_cast<int>(std::max<long>(1, std::min<long>((n16 + 511) / 512, 32)));
  peer_push_kernel<<<grid, 512, 0, at::cuda::getCurrentCUDAStream()>>>(
      reinterpret_cast<const uint4*>(src.data_ptr()), pp, reinterpret_cast<uint4*>(dst_mc), n16, static_cast<int>(rank), world,
      static_cast<int>(slot), static_cast<uint32_t>(epoch), signal ? 1 : 0,
      reinterpret_cast<unsigned int*>(done_counter.data_ptr()), static_cast<int>(wait_slot), edp);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// Raise this rank's `slot` flag to `epoch` on every peer (stream-ordered after the kernels that consumed the data).
__global__ void peer_signal_kernel(PeerPtrs dst, int rank, int world, int slot, uint32_t epoch,
                                   const uint32_t* __restrict__ epoch_dev) {
  if (epoch_dev != nullptr) epoch += *reinterpret_cast<const volatile uint32_t*>(epoch_dev);
  __threadfence_system();
  if (threadIdx.x < world) st_release_sys(dst.flags[threadIdx.x] + slot * kMaxRanks + rank, epoch);
}
void peer_signal(std::vector<int64_t> flag_ptrs, int64_t rank, int64_t slot, int64_t epoch,
                 const c10::optional<at::Tensor>& epoch_dev) {
  const int world = static_cast<int>(flag_ptrs.size());
  TORCH_CHECK(world >= 1 && world <= kMaxRanks);
  PeerPtrs pp{};
  for (int p = 0; p < world; ++p) pp.flags[p] = reinterpret_cast<uint32_t*>(flag_ptrs[p]);
  const uint32_t* edp = epoch_dev.has_value() ? reinterpret_cast<const uint32_t*>(epoch_dev->data_ptr()) : nullptr;
  peer_signal_kernel<<<1, 32, 0, at::cuda::getCurrentCUDAStream()>>>(pp, static_cast<int>(rank), world, static_cast<int>(slot),
                                                                     static_cast<uint32_t>(epoch), edp);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace psd

TORCH_LIBRARY_FRAGMENT(poseidon, m) {
  m.def("fused_update(Tensor(a!) w, Tensor(d!) g, Tensor(b!) h, Tensor(c!)? wb, float lr, float momentum, float decay, int rule, "
        "bool l1, float delta, float gscale, Tensor? lr_dev, bool rearm=False) -> ()", &psd::fused_update);
  m.def("fused_update_multi(Tensor[] ws, Tensor[] gs, Tensor[] hs, Tensor[] wbs, float[] lrs, float[] decays, int[] rearms, "
        "float momentum, int rule, bool l1, float delta, float gscale, Tensor? lr_dev) -> ()", &psd::fused_update_multi);
  m.def("allreduce_sgd(int[] g_ptrs, int[] w_ptrs, int[] wb_ptrs, int[] flag_ptrs, int g_mc, int w_mc, Tensor(a!) h, int n, "
        "int rank, int epoch, bool one_shot, Tensor(b!) done_counter, float lr, float momentum, float decay, int rule, "
        "bool l1, float delta, float gscale, int max_ctas, Tensor? lr_dev, Tensor? epoch_dev) -> ()", &psd::allreduce_sgd);
  m.def("allreduce_sgd_multi(int[] base_ptrs, int mc_base, int[] flag_ptrs, int[] g_offs, int[] w_offs, int[] wb_offs, "
        "Tensor[] hists, int[] ns, int[] one_shots, float[] lrs, float[] decays, int rank, int epoch, Tensor(a!) done_counter, "
        "float momentum, int rule, bool l1, float delta, float gscale, int max_ctas, Tensor? lr_dev, Tensor? epoch_dev) -> ()",
        &psd::allreduce_sgd_multi);
  m.def("ssp_delta(int[] base_ptrs, int[] flag_ptrs, int[] g_offs, int[] w_offs, int[] wb_offs, int[] d_offs, int ring_stride, "
        "Tensor[] hists, int[] ns, float[] lrs, float[] decays, int rank, int ring, int staleness, Tensor(a!) done_counter, "
        "Tensor(b!) state, float momentum, int rule, bool l1, float delta, float gscale, int max_ctas, Tensor? lr_dev, "
        "Tensor clock_dev) -> ()", &psd::ssp_delta);
  m.def("ssp_fold(int[] base_ptrs, int[] flag_ptrs, int[] w_offs, int[] wb_offs, int[] d_offs, int ring_stride, int[] ns, "
        "int rank, int ring, Tensor(a!) done_counter, Tensor(b!) state, bool drain, int max_ctas) -> ()", &psd::ssp_fold);
  m.def("peer_push(Tensor src, int[] dst_ptrs, int dst_mc, int[] flag_ptrs, int rank, int slot, int epoch, bool signal, "
        "Tensor(a!) done_counter, int wait_slot, Tensor? epoch_dev) -> ()", &psd::peer_push);
  m.def("peer_signal(int[] flag_ptrs, int rank, int slot, int epoch, Tensor? epoch_dev) -> ()", &psd::peer_signal);
}
