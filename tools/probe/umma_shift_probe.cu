// Probe: can a K-major SWIZZLE_128B UMMA operand start at a 128-byte (one row) offset inside a TMA-written slab?
// Loads a (128+8)-row x 64-col bf16 slab with TMA (128B swizzle), then for shift s = 0..7 runs D = A[s:s+128] * B^T
// with descriptor start = slab + 128*s and base_offset = {0, s}; prints the max error of each variant.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../pgtformer_b200/csrc/ptx.cuh"
using namespace pgt;

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int shift, int boff,
      int sbo_bytes) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 256 rows x 128 B = 32 KB
  uint8_t* sB = smem + 32768;         // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint64_t* done = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc<64>(tptr); tc_fence_before(); }
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 32768 + 8192);
    tma_load_2d(sA, &tmA, bar, 0, 0);
    tma_load_2d(sB, &tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    uint64_t da = umma_desc_k_sw128(smem_u32(sA) + 128 * shift);
    da &= ~(uint64_t(0x3FFF) << 32);
    da |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
    da |= uint64_t(boff & 7) << 49;
    const uint64_t db = umma_desc_k_sw128(smem_u32(sB));
    for (int k = 0; k < 4; ++k) umma_bf16_ss(tb, da + 2 * k, db + 2 * k, umma_idesc_bf16(128, 64), k ? 1u : 0u);
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < 64; c += 32) {
    tmem_ld_32x32(tb + (uint32_t(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tb);
}

int main() {
  const int RA = 256, K = 64, N = 64;
  std::vector<__nv_bfloat16> hA(RA * K), hB(N * K);
  std::vector<float> fA(RA * K), fB(N * K);
  srand(1);
  for (int i = 0; i < RA * K; ++i) { float v = (rand() % 17 - 8) / 8.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { float v = (rand() % 13 - 6) / 8.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dO;
  cudaMalloc(&dA, RA * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), RA * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
  CUtensorMap tA, tB;
  cuuint64_t gd[2] = {64, (cuuint64_t)RA}; cuuint64_t gs[1] = {128}; cuuint32_t bx[2] = {64, 256}; cuuint32_t es[2] = {1, 1};
  enc(&tA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cuuint64_t gd2[2] = {64, 64}; cuuint32_t bx2[2] = {64, 64};
  enc(&tB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, gd2, gs, bx2, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024);
  std::vector<float> hO(128 * 64);
  // (shift, base_offset, sbo): contiguous rows (sbo 1024) and a pitched slab (8-row groups 1280 B apart)
  for (int sbo : {1024, 1280}) {
    for (int s = 0; s < 12; ++s) {
      for (int boff : {0, s & 7}) {
        probe<<<1, 128, 49152 + 1024>>>(tA, tB, dO, s, boff, sbo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sbo %d shift %d boff %d: CUDA error %s\n", sbo, s, boff, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hO.data(), dO, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        double me = 0;
        for (int r = 0; r < 128; ++r) {
          const int src = s + (r / 8) * (sbo / 128) + (r % 8);       // slab row feeding tile row r
          for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)fA[src * K + k] * fB[n * K + k];
            me = fmax(me, fabs(ref - hO[r * 64 + n]));
          }
        }
        printf("sbo %4d shift %2d base_offset %d : max err %.4f %s\n", sbo, s, boff, me, me < 1e-3 ? "OK" : "MISMATCH");
        if (boff == (s & 7) && boff == 0) break;
      }
    }
  }
  return 0;
}
