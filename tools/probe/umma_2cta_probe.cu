// Probe: tcgen05.mma.cta_group::2 mechanics on sm_100a, as the halo conv would use them.
//   cluster of 2 CTAs; CTA r holds rows [128 r, 128 r + 128) of A (its own M half) and rows [64 r, 64 r + 64) of B
//   (its half of N = 128) at the SAME smem offsets; both CTAs' TMA loads complete_tx on the LEADER's mbarrier
//   (cp.async.bulk.tensor ... .cta_group::2 with a mapa'd barrier address); the leader's elected thread issues the
//   M = 256 MMAs and commits with .multicast::cluster to a barrier in both CTAs; each CTA reads its 128 TMEM lanes.
// Prints max |err| against a host reference for D = A B^T (256 x 128, K = 64).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../pgtformer_b200/csrc/ptx.cuh"
using namespace pgt;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// TMA load whose complete_tx lands on a barrier given as a shared::cluster address (may be the peer CTA's)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2cta(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 128 rows x 128 B = 16 KB
  uint8_t* sB = smem + 16384;         // 64 rows x 128 B = 8 KB
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
  uint64_t* done = full + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(full + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) { mbar_init(full, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();                 // both CTAs' barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tb = *tptr;
  if (threadIdx.x == 0) {
    const uint32_t leader_full = mapa_u32(smem_u32(full), 0);
    if (rank == 0) mbar_arrive_expect_tx(full, 2 * (16384 + 8192));
    tma_load_2d_2sm(sA, &tmA, leader_full, 0, rank * 128);
    tma_load_2d_2sm(sB, &tmB, leader_full, 0, rank * 64);
    if (rank == 0) {
      mbar_wait(full, 0);
      tc_fence_after();
      const uint64_t da = umma_desc_k_sw128(smem_u32(sA));
      const uint64_t db = umma_desc_k_sw128(smem_u32(sB));
      for (int k = 0; k < 4; ++k) umma_bf16_ss_2cta(tb, da + 2 * k, db + 2 * k, umma_idesc_bf16(256, 128), k ? 1u : 0u);
      umma_commit_2cta(done, 3);
    }
  }
  mbar_wait(done, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < 128; c += 32) {
    tmem_ld_32x32(tb + (uint32_t(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(size_t)(rank * 128 + warp * 32 + lane) * 128 + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 128;" ::"r"(tb) : "memory");
}

int main() {
  const int M = 256, K = 64, N = 128;
  std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K);
  srand(1);
  for (int i = 0; i < M * K; ++i) { float v = (rand() % 17 - 8) / 8.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 8.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dO;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dO, M * N * 4);
  cudaMemset(dO, 0xff, M * N * 4);
  cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
  CUtensorMap tA, tB;
  cuuint64_t gs[1] = {128}; cuuint32_t es[2] = {1, 1};
  cuuint64_t gdA[2] = {64, (cuuint64_t)M}; cuuint32_t bxA[2] = {64, 128};
  enc(&tA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, gdA, gs, bxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cuuint64_t gdB[2] = {64, (cuuint64_t)N}; cuuint32_t bxB[2] = {64, 64};
  enc(&tB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, gdB, gs, bxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int smem = 16384 + 8192 + 256;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<2, 128, smem>>>(tA, tB, dO);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> hO(M * N);
  cudaMemcpy(hO.data(), dO, M * N * 4, cudaMemcpyDeviceToHost);
  double me = 0;
  for (int r = 0; r < M; ++r)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)fA[r * K + k] * fB[n * K + k];
      me = fmax(me, fabs(ref - hO[r * N + n]));
    }
  printf("2-CTA MMA M256 N128 K64: max err %.5f %s\n", me, me < 1e-3 ? "OK" : "MISMATCH");
  return 0;
}
