// Fused Swin MLP half-block on tcgen05 (C = 256):   out = x + W2 * GELU(W1 * LN(x) + b1) + b2
// Replaces LayerNorm(norm2) + Mlp(fc1, exact GELU, fc2) + residual of VSTSREncoderTransformerBlock
// (modules/rstt_layers.py:116-132,335-336): five HBM passes over the token matrix (LN r/w, fc1 r/w, fc2 r+r/w)
// become two (read x, write out); the 128 x 256 hidden tile never leaves the SM.
//
// Per CTA, persistent over 128-token tiles (352 threads):
//   warp 0      TMA producer of the weight k-blocks (W1 then W2, 256 x 64 each, 3-deep ring)
//   warp 1      tcgen05.mma issuer:  acc1 = LN(x) W1^T  (TMEM cols 0..255),  acc2 = H W2^T  (cols 256..511)
//   warps 2..9  compute: LayerNorm of the x tile from smem (two threads per row), GELU epilogue of acc1 written
//               straight back into the A-operand tile, final epilogue acc2 + b2 + x written over the x tile
//   warp 10     DMA: TMA load of the x tile (kept intact as the residual), TMA store of the finished tile
#include <cudaTypedefs.h>

#include "common.cuh"
#include "ptx.cuh"
#include "epi_common.cuh"

namespace pgt {

constexpr int SM_C = 256;
constexpr int SM_BM = 128;
constexpr int SM_SUB = SM_BM * 128;            // one [128 x 64] bf16 sub-tile: 16 KB
constexpr int SM_WST = 3;                      // weight ring depth
constexpr int SM_WBYTES = SM_C * 128;          // one [256 x 64] weight k-block: 32 KB
constexpr int SM_THREADS = 352;
constexpr int SM_SMEM = 4 * SM_SUB /*R*/ + 4 * SM_SUB /*A*/ + SM_WST * SM_WBYTES + 128 * 2 * 8 /*xch*/ + 256;

struct SwinMlpParams {
  int T, m_tiles;
  const float* ln_g;
  const float* ln_b;
  float eps;
  const float* b1;
  const float* b2;
  float* gn_stats;      // optional [m_tiles][4][32][2]
};

__global__ void __launch_bounds__(SM_THREADS, 1)
swin_mlp_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmO,
                const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                const SwinMlpParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];     // no static smem in this kernel: the window starts 1024-aligned
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sR = smem;                           // x tile (residual), later the output tile: 4 swizzled sub-tiles
  uint8_t* sA = sR + 4 * SM_SUB;                // A operand: LN(x), later GELU(hidden)
  uint8_t* sW = sA + 4 * SM_SUB;                // weight ring
  float2* xch = reinterpret_cast<float2*>(sW + SM_WST * SM_WBYTES);     // [2 halves][128 rows] (sum, sumsq)
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 256);
  uint64_t* x_full = bars;          // DMA -> compute
  uint64_t* r_free = bars + 1;      // DMA (after its store drained) -> DMA next load   (kept as a barrier for symmetry)
  uint64_t* y_ready = bars + 2;     // compute -> MMA   (256)
  uint64_t* acc1_full = bars + 3;   // MMA -> compute
  uint64_t* h_ready = bars + 4;     // compute -> MMA   (256)
  uint64_t* acc2_full = bars + 5;   // MMA -> compute
  uint64_t* out_ready = bars + 6;   // compute -> DMA   (256)
  uint64_t* w_full = bars + 7;      // [SM_WST]
  uint64_t* w_empty = w_full + SM_WST;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_empty + SM_WST);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmO); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    mbar_init(x_full, 1); mbar_init(r_free, 1);
    mbar_init(y_ready, 256); mbar_init(acc1_full, 1); mbar_init(h_ready, 256); mbar_init(acc2_full, 1);
    mbar_init(out_ready, 256);
    for (int i = 0; i < SM_WST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      for (int j = 0; j < 8; ++j) {
        mbar_wait(&w_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&w_full[st], SM_WBYTES);
          tma_load_2d(sW + st * SM_WBYTES, j < 4 ? &tmW1 : &tmW2, &w_full[st], (j & 3) * 64, 0);
        }
        __syncwarp();
        if (++st == SM_WST) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(SM_BM, SM_C);
    int st = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      for (int gemm = 0; gemm < 2; ++gemm) {
        mbar_wait(gemm == 0 ? y_ready : h_ready, par);
        tc_fence_after();
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(&w_full[st], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = umma_desc_k_sw128(smem_u32(sA + kb * SM_SUB));
            const uint64_t db = umma_desc_k_sw128(smem_u32(sW + st * SM_WBYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + gemm * SM_C, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&w_empty[st]);
            if (kb == 3) umma_commit(gemm == 0 ? acc1_full : acc2_full);
          }
          __syncwarp();
          if (++st == SM_WST) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ compute warps
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                    // which 128-column half of the row this thread owns
    const int r = quad * 32 + lane;
    const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16);
    const int c_lo = half * 128;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      // ---- LayerNorm(x) -> A operand (two threads per row: partial sums exchanged through smem)
      mbar_wait(x_full, par);
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const uint8_t* src = sR + (half * 2 + sub) * SM_SUB + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = *reinterpret_cast<const uint4*>(src + ((c ^ (r & 7)) << 4));
          const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
          s += (a.x + a.y) + (b.x + b.y) + (cc.x + cc.y) + (d.x + d.y);
          q += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + cc.x * cc.x + cc.y * cc.y + d.x * d.x + d.y * d.y;
        }
      }
      xch[half * 128 + r] = make_float2(s, q);
      named_bar_sync(5, 256);
      const float2 other = xch[(half ^ 1) * 128 + r];
      const float mean = (s + other.x) * (1.f / SM_C);
      const float var = fmaxf((q + other.y) * (1.f / SM_C) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int kb = half * 2 + sub;
        const uint8_t* src = sR + kb * SM_SUB + r * 128;
        uint8_t* dst = sA + kb * SM_SUB + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int off = (c ^ (r & 7)) << 4;
          const uint4 u = *reinterpret_cast<const uint4*>(src + off);
          const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
          const float v[8] = {a.x, a.y, b.x, b.y, cc.x, cc.y, d.x, d.y};
          const int col = kb * 64 + c * 8;
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_g + col)), g1 = __ldg(reinterpret_cast<const float4*>(p.ln_g + col + 4));
          const float4 e0 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col)), e1 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col + 4));
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = (v[j] - mean) * rstd * gg[j] + ee[j];
          uint4 o;
          o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]);
          o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
          *reinterpret_cast<uint4*>(dst + off) = o;
        }
      }
      fence_proxy_async();
      mbar_arrive(y_ready);
      // ---- hidden = GELU(acc1 + b1) -> A operand (in place: GEMM1 has finished reading it)
      mbar_wait(acc1_full, par);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = c_lo; c0 < c_lo + 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c0, v);
        tmem_ld_wait();
        float f[32];
        const float4* b4 = reinterpret_cast<const float4*>(p.b1 + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = __ldg(b4 + i);
          f[4 * i + 0] = gelu_erf(__uint_as_float(v[4 * i + 0]) + bb.x);
          f[4 * i + 1] = gelu_erf(__uint_as_float(v[4 * i + 1]) + bb.y);
          f[4 * i + 2] = gelu_erf(__uint_as_float(v[4 * i + 2]) + bb.z);
          f[4 * i + 3] = gelu_erf(__uint_as_float(v[4 * i + 3]) + bb.w);
        }
        uint8_t* dst = sA + (c0 >> 6) * SM_SUB + r * 128;
        const int ch0 = (c0 & 63) >> 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack_bf16x2(f[8 * i + 0], f[8 * i + 1]); o.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
          o.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]); o.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
          *reinterpret_cast<uint4*>(dst + (((ch0 + i) ^ (r & 7)) << 4)) = o;
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(h_ready);
      // ---- out = acc2 + b2 + x, written over the x tile (then TMA-stored by the DMA warp)
      mbar_wait(acc2_full, par);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = c_lo; c0 < c_lo + 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + SM_C + c0, v);
        tmem_ld_wait();
        float f[32];
        const float4* b4 = reinterpret_cast<const float4*>(p.b2 + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = __ldg(b4 + i);
          f[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + bb.x;
          f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + bb.y;
          f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + bb.z;
          f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + bb.w;
        }
        uint8_t* row = sR + (c0 >> 6) * SM_SUB + r * 128;
        const int ch0 = (c0 & 63) >> 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4* dst = reinterpret_cast<uint4*>(row + (((ch0 + i) ^ (r & 7)) << 4));
          const uint4 u = *dst;
          const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
          f[8 * i + 0] += a.x; f[8 * i + 1] += a.y; f[8 * i + 2] += b.x; f[8 * i + 3] += b.y;
          f[8 * i + 4] += cc.x; f[8 * i + 5] += cc.y; f[8 * i + 6] += d.x; f[8 * i + 7] += d.y;
          uint4 o;
          o.x = pack_bf16x2(f[8 * i + 0], f[8 * i + 1]); o.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
          o.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]); o.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
          *dst = o;
        }
        if (p.gn_stats != nullptr)
          gn_chunk_stats<8>(f, p.gn_stats + (((size_t)tile * 4 + quad) * 32 + c0 / 8) * 2, 0, lane);
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(out_ready);
    }
  } else {
    // ------------------------------------------------------------------ DMA warp: x tile in, finished tile out
    int it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      if (lane == 0) {
        mbar_arrive_expect_tx(x_full, 4 * SM_SUB);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sR + kb * SM_SUB, &tmX, x_full, kb * 64, tile * SM_BM);
      }
      __syncwarp();
      mbar_wait(out_ready, par);
      if (lane == 0) {
        for (int kb = 0; kb < 4; ++kb) tma_store_2d(&tmO, sR + kb * SM_SUB, kb * 64, tile * SM_BM);
        bulk_commit();
        bulk_wait_read<0>();          // the x / out buffer may be refilled
      }
      __syncwarp();
    }
    if (lane == 0) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static int enc2d(CUtensorMap* map, const void* base, int ld, long long rows, int cols, int box_rows) {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return PGT_ERR_DRIVER;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PGT_OK : PGT_ERR_DRIVER;
}

}  // namespace pgt

using namespace pgt;

extern "C" int pgt_swin_mlp_bf16(const void* x, int ldx, int T, int C, const float* ln_g, const float* ln_b, float eps,
                                 const void* W1, const float* b1, const void* W2, const float* b2, void* out, int ldo,
                                 float* gn_stats, void* stream) {
  PGT_CHECK_ARG(x && out && ln_g && ln_b && W1 && W2 && b1 && b2 && T > 0);
  if (C != SM_C) return PGT_ERR_UNSUPPORTED;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  PGT_CHECK_ARG(al(x) && al(out) && al(W1) && al(W2) && ldx % 8 == 0 && ldo % 8 == 0);
  CUtensorMap tx, to, t1, t2;
  int rc = enc2d(&tx, x, ldx, T, C, SM_BM);
  if (rc == PGT_OK) rc = enc2d(&to, out, ldo, T, C, SM_BM);
  if (rc == PGT_OK) rc = enc2d(&t1, W1, C, C, C, SM_C);
  if (rc == PGT_OK) rc = enc2d(&t2, W2, C, C, C, SM_C);
  if (rc != PGT_OK) return rc;
  static bool attr = false;
  if (!attr) {
    PGT_CUDA_OK(cudaFuncSetAttribute(swin_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM));
    attr = true;
  }
  SwinMlpParams p{};
  p.T = T; p.m_tiles = ceil_div(T, SM_BM);
  p.ln_g = ln_g; p.ln_b = ln_b; p.eps = eps; p.b1 = b1; p.b2 = b2; p.gn_stats = gn_stats;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
  {
    ProfScope ps(PGT_PROF_GEMM, 4.0 * (double)T * C * C, st, "swin_mlp_fused C256");
    swin_mlp_kernel<<<grid, SM_THREADS, SM_SMEM, st>>>(tx, to, t1, t2, p);
  }
  PGT_LAUNCH_OK();
  return PGT_OK;
}
