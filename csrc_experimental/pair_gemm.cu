// EXPERIMENTAL — paired-CTA (cta_group::2) tcgen05 GEMM.  NOT part of the shipped extension: it has been compiled and its
// SASS inspected, but it has not yet run on a GPU (round 1 ran out of GPU budget).  Build + test on a B200 with
//     POSEIDON_EXPERIMENTAL=1 python -m pytest tests/test_pair_gemm_gpu.py
// Purpose: validate the 2-CTA protocol of DESIGN.md §8 on the simplest shape (C[M,N] bf16 = A[M,K] · B[N,K]ᵀ, both K-major,
// BN = 256) before porting it into umma_gemm.cuh for the im2col-TMA convolution modes.
//
// Why: a single-CTA UMMA tile writes (16 KB + BN*128 B) into shared memory per k-block and reads the same bytes back —
// 96 KB at BN = 256 against 512 cycles of MMA — and is shared-memory-bandwidth bound (profiles/r1_conv_ncu_summary.md §4).
// With cta_group::2 a pair of CTAs computes a 256 x BN tile: each CTA stages its own 128 rows of A and HALF of B
// (BN/2 rows); the tensor cores of both SMs read both halves.  Per SM and k-block: 32 KB written + 32 KB read.
//
// Roles per CTA (384 threads): warp 0 TMA producer (both CTAs), warp 1 MMA issuer (leader CTA only), warp 2 TMEM
// allocation, warps 4-11 epilogue (each CTA drains its own 128 TMEM lanes).
// Barriers: full[s] lives in the LEADER (one expect_tx arrival for the bytes of both CTAs; both CTAs' TMA loads complete
// on it through the .cta_group::2 form with the peer bit of the barrier address cleared); empty[s] and tmem_full[a] live
// in each CTA and are arrived by the leader's multicast tcgen05.commit; tmem_empty[a] lives in the leader and collects one
// arrival per epilogue warp of BOTH CTAs (the follower arrives remotely through mapa).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>

#include "../csrc/gemm/sm100_prims.cuh"

namespace psd_exp {
using namespace psd;

constexpr int BM = 128;            // rows per CTA (256 per pair)
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int kStages = 5;
constexpr int kABytes = BM * BK * 2;            // 16 KB
constexpr int kBHalfBytes = (BN / 2) * BK * 2;  // 16 KB: this CTA's half of the B tile
constexpr int kStageBytes = kABytes + kBHalfBytes;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int kEpiStageBytes = kEpiWarps * 32 * 33 * 4;
constexpr int kSmemBytes = kStages * kStageBytes + 256 + kEpiStageBytes + 1024;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank-in-pair bit of a shared::cluster address

struct Maps {
  CUtensorMap a, b;
};

__device__ __forceinline__ void tma_load_2d_cg2(uint32_t smem_dst, const CUtensorMap* map, uint32_t leader_bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
pair_gemm_kernel(const __grid_constant__ Maps tm, __nv_bfloat16* __restrict__ c, long ldc, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + kStages * kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(bar_base + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();          // 0 = leader of the pair
  const bool leader = crank == 0;
  const int m_pairs = (M + 2 * BM - 1) / (2 * BM);
  const int n_blocks = (N + BN - 1) / BN;
  const int num_tiles = m_pairs * n_blocks;
  const int kbs = (K + BK - 1) / BK;
  const int tile0 = blockIdx.x >> 1, tile_step = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a);
    tma_prefetch_desc(&tm.b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);                    // leader: one expect_tx arrival (follower's copy is unused)
      mbar_init(&empty_bar[s], 1);                   // one multicast commit per use
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * kEpiWarps);      // leader: epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_cg2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // both CTAs' barriers and TMEM exist before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs, warp-uniform, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int mp = tile % m_pairs, n_blk = tile / m_pairs;
      const int m_row = (2 * mp + static_cast<int>(crank)) * BM;
      const int n_row = n_blk * BN + static_cast<int>(crank) * (BN / 2);
      for (int kb = 0; kb < kbs; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);     // own barrier, released by the leader's multicast commit
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t lbar = smem_u32(&full_bar[stage]) & kPeerBitMask;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
          tma_load_2d_cg2(sa, &tm.a, lbar, kb * BK, m_row);
          tma_load_2d_cg2(sa + kABytes, &tm.b, lbar, kb * BK, n_row);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, false, false);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
      const int as = it & 1;
      mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < kbs; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
        const uint32_t b_addr = a_addr + kABytes;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16_cg2(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_cg2(&empty_bar[stage], 0x3);             // frees this stage in both CTAs
          if (kb == kbs - 1) umma_commit_cg2(&tmem_full[as], 0x3);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 4 + kEpiWarps) {
    // ===================== epilogue (each CTA: its own 128 TMEM lanes) =====================
    const int e = warp - 4, q = e & 3, half = e >> 2;
    float* stg = epi_stage + e * (32 * 33);
    int it = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
      const int mp = tile % m_pairs, n_blk = tile / m_pairs;
      const int as = it & 1;
      mbar_wait(&tmem_full[as], (it >> 1) & 1);
      tc_fence_after();
      const int row0 = (2 * mp + static_cast<int>(crank)) * BM + q * 32;
#pragma unroll 1
      for (int cc = half; cc < BN / 32; cc += 2) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(r[j]);
        __syncwarp();
        const int col = n_blk * BN + cc * 32 + lane;
        if (col < N) {
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr)
            if (row0 + rr < M) c[static_cast<long>(row0 + rr) * ldc + col] = __float2bfloat16(stg[rr * 33 + lane]);
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[as]);
        else mbar_arrive_remote(&tmem_empty[as], 0);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                // nobody leaves (or frees TMEM) while the peer may still touch it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 512);
  }
}

void encode_2d(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer) {
  using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Fn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    TORCH_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && ptr != nullptr);
    fn = reinterpret_cast<Fn>(ptr);
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  TORCH_CHECK(fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
}

// C[M,N] bf16 = A[M,K] · B[N,K]ᵀ  (both K-major, K % 8 == 0)
at::Tensor pair_gemm_bf16(const at::Tensor& a, const at::Tensor& b) {
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && a.scalar_type() == at::kBFloat16 && b.scalar_type() == at::kBFloat16);
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && a.stride(1) == 1 && b.stride(1) == 1 && a.size(1) == b.size(1));
  c10::cuda::CUDAGuard guard(a.device());
  const int64_t M = a.size(0), N = b.size(0), K = a.size(1);
  at::Tensor c = at::empty({M, N}, a.options());
  Maps tm;
  encode_2d(&tm.a, a.data_ptr(), K, M, a.stride(0), BK, BM);
  encode_2d(&tm.b, b.data_ptr(), K, N, b.stride(0), BK, BN / 2);
  auto kern = pair_gemm_kernel;
  static bool configured = false;
  if (!configured) {
    C10_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int64_t pairs = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  const int grid = static_cast<int>(2 * std::max<int64_t>(1, std::min<int64_t>(pairs, sms / 2)));
  kern<<<grid, kThreads, kSmemBytes, at::cuda::getCurrentCUDAStream()>>>(
      tm, reinterpret_cast<__nv_bfloat16*>(c.data_ptr()), c.stride(0), static_cast<int>(M), static_cast<int>(N),
      static_cast<int>(K));
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return c;
}

}  // namespace psd_exp

TORCH_LIBRARY_FRAGMENT(poseidon_exp, m) {
  m.def("pair_gemm_bf16(Tensor a, Tensor b) -> Tensor", &psd_exp::pair_gemm_bf16);
}
