// LayerNorm fused into the GEMM that consumes it (C = 256):   out[T, N] = LN(x)[T, 256] * W[N, 256]^T + b
// Replaces norm1 + the q / kv projections of WindowAttention3D (modules/rstt_layers.py:116-132,176-188, fused into one
// [3C, C] weight at load): the normalised token matrix is never written to HBM — x is read once, LN happens in shared
// memory in place, and the same tile is the A operand of all N / 256 column blocks.
//
// Per CTA, persistent over 128-token tiles (352 threads):
//   warp 0      TMA producer of the weight k-blocks ([256 x 64] each, 3-deep ring; N/256 x 4 per tile)
//   warp 1      tcgen05.mma issuer: column block j accumulates in TMEM columns [256 (j & 1), +256) — the epilogue of
//               block j overlaps the MMAs of block j + 1
//   warps 2..9  compute: LayerNorm of the x tile in place (two threads per row), then per column block the epilogue
//               acc + bias -> bf16 -> swizzled staging tile
//   warp 10     DMA: TMA load of the x tile (as soon as the last MMA of the previous tile has consumed the buffer),
//               TMA store of each finished [128 x 256] block
#include <cudaTypedefs.h>

#include <cstdio>

#include "common.cuh"
#include "tmap.cuh"
#include "ptx.cuh"

namespace pgt {

constexpr int LL_C = 256;
constexpr int LL_BM = 128;
constexpr int LL_SUB = LL_BM * 128;            // one [128 x 64] bf16 sub-tile: 16 KB
constexpr int LL_WST = 3;                      // weight ring depth
constexpr int LL_WBYTES = LL_C * 128;          // one [256 x 64] weight k-block: 32 KB
constexpr int LL_THREADS = 352;
constexpr int LL_SMEM = 4 * LL_SUB /*x / A*/ + 4 * LL_SUB /*staging*/ + LL_WST * LL_WBYTES + 128 * 8 /*xch*/ + 256;

struct LnLinearParams {
  int T, m_tiles, nb;   // nb = N / 256 column blocks
  const float* ln_g;
  const float* ln_b;
  float eps;
  const float* bias;    // [N]
};

__global__ void __launch_bounds__(LL_THREADS, 1)
ln_linear_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmO,
                 const __grid_constant__ CUtensorMap tmW, const LnLinearParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];     // no static smem in this kernel: the window starts 1024-aligned
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sX = smem;                           // x tile -> LN(x): the A operand of every column block
  uint8_t* sO = sX + 4 * LL_SUB;                // finished [128 x 256] block, 4 swizzled sub-tiles
  uint8_t* sW = sO + 4 * LL_SUB;                // weight ring
  float2* xch = reinterpret_cast<float2*>(sW + LL_WST * LL_WBYTES);     // [128 rows] (sum, sumsq) exchange
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 128);
  uint64_t* x_full = bars;           // DMA -> compute
  uint64_t* sx_free = bars + 1;      // MMA commit (last column block) -> DMA
  uint64_t* y_ready = bars + 2;      // compute (256) -> MMA
  uint64_t* acc_full = bars + 3;     // [2] MMA commit -> compute
  uint64_t* acc_free = bars + 5;     // [2] compute (256) -> MMA
  uint64_t* so_free = bars + 7;      // [2] DMA (store has read staging half h) -> compute
  uint64_t* out_ready = bars + 9;    // [2] compute (256) -> DMA
  uint64_t* w_full = bars + 11;      // [LL_WST]
  uint64_t* w_empty = w_full + LL_WST;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_empty + LL_WST);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmO); tma_prefetch_desc(&tmW);
    mbar_init(x_full, 1); mbar_init(sx_free, 1); mbar_init(y_ready, 256);
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_free[i], 256); }
    for (int i = 0; i < 2; ++i) { mbar_init(&so_free[i], 1); mbar_init(&out_ready[i], 256); }
    for (int i = 0; i < LL_WST; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int NB = p.nb;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      for (int j = 0; j < NB; ++j) {
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(&w_empty[st], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&w_full[st], LL_WBYTES);
            tma_load_2d(sW + st * LL_WBYTES, &tmW, &w_full[st], kb * 64, j * LL_C);
          }
          __syncwarp();
          if (++st == LL_WST) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(LL_BM, LL_C);
    int st = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      mbar_wait(y_ready, it & 1);
      tc_fence_after();
      if (elect_one()) {
        int s = st;
        uint32_t sp = ph;
        const uint64_t da0 = umma_desc_k_sw128(smem_u32(sX));
        const uint64_t db0 = umma_desc_k_sw128(smem_u32(sW));
        for (int j = 0; j < NB; ++j) {
          const int g = it * NB + j;
          mbar_wait(&acc_free[g & 1], ((g >> 1) & 1) ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(&w_full[s], sp);
            tc_fence_after();
            const uint64_t da = da0 + (uint64_t)(kb * (LL_SUB >> 4));
            const uint64_t db = db0 + (uint64_t)(s * (LL_WBYTES >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + (g & 1) * LL_C, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&w_empty[s]);
            if (++s == LL_WST) { s = 0; sp ^= 1; }
          }
          umma_commit(&acc_full[g & 1]);
        }
        umma_commit(sx_free);                       // every MMA that reads the x tile has retired
      }
      __syncwarp();
      const int ns = st + 4 * NB;
      ph ^= (ns / LL_WST) & 1;
      st = ns % LL_WST;
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ compute warps
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                    // which 128-column half of the row this thread owns
    const int r = quad * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      // ---- LayerNorm(x) in place (two threads per row; the halves meet through smem)
      mbar_wait(x_full, it & 1);
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const uint8_t* src = sX + (half * 2 + sub) * LL_SUB + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = *reinterpret_cast<const uint4*>(src + ((c ^ (r & 7)) << 4));
          const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
          s += (a.x + a.y) + (b.x + b.y) + (cc.x + cc.y) + (d.x + d.y);
          q += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + cc.x * cc.x + cc.y * cc.y + d.x * d.x + d.y * d.y;
        }
      }
      if (half == 0) xch[r] = make_float2(s, q);
      named_bar_sync(5, 256);
      if (half == 1) {
        const float2 o = xch[r];
        s += o.x; q += o.y;
        xch[r] = make_float2(s, q);
      }
      named_bar_sync(5, 256);
      if (half == 0) {
        const float2 o = xch[r];
        s = o.x; q = o.y;
      }
      const float mean = s * (1.f / LL_C);
      const float var = fmaxf(q * (1.f / LL_C) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
      const float nm = -mean * rstd;
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int kb = half * 2 + sub;
        uint8_t* row = sX + kb * LL_SUB + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4* ptr = reinterpret_cast<uint4*>(row + ((c ^ (r & 7)) << 4));
          const uint4 u = *ptr;
          const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
          const float v[8] = {a.x, a.y, b.x, b.y, cc.x, cc.y, d.x, d.y};
          const int col = kb * 64 + c * 8;
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_g + col)), g1 = __ldg(reinterpret_cast<const float4*>(p.ln_g + col + 4));
          const float4 e0 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col)), e1 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col + 4));
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = fmaf(fmaf(v[j], rstd, nm), gg[j], ee[j]);
          uint4 o;
          o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]);
          o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
          *ptr = o;
        }
      }
      fence_proxy_async();
      mbar_arrive(y_ready);
      // ---- per column block: acc + bias -> bf16 -> staging tile
      for (int j = 0; j < NB; ++j) {
        const int g = it * NB + j;
        const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16) + (g & 1) * LL_C;
        mbar_wait(&acc_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        // the [128 x 256] block leaves in two 128-column halves through a ping-pong pair of 32 KB staging buffers:
        // the TMA store of one half reads its buffer while the warps fill the other
#pragma unroll 1
        for (int hp = 0; hp < 2; ++hp) {
          mbar_wait(&so_free[hp], (g & 1) ^ 1);         // the store that last used this half has read it
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            const int c0 = hp * 128 + half * 64 + cc * 32;
            uint32_t v[32];
            tmem_ld_32x32(t_row + c0, v);
            tmem_ld_wait();
            if (hp == 1 && cc == 1) {                    // last TMEM read of this thread: the accumulator may be reused
              tc_fence_before();
              mbar_arrive(&acc_free[g & 1]);
            }
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + j * LL_C + c0);
            uint8_t* dst = sO + (c0 >> 6) * LL_SUB + r * 128;
            const int ch0 = (c0 & 63) >> 3;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 ba = __ldg(b4 + 2 * i), bb = __ldg(b4 + 2 * i + 1);
              uint4 o;
              o.x = pack_bf16x2(__uint_as_float(v[8 * i + 0]) + ba.x, __uint_as_float(v[8 * i + 1]) + ba.y);
              o.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) + ba.z, __uint_as_float(v[8 * i + 3]) + ba.w);
              o.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) + bb.x, __uint_as_float(v[8 * i + 5]) + bb.y);
              o.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) + bb.z, __uint_as_float(v[8 * i + 7]) + bb.w);
              *reinterpret_cast<uint4*>(dst + (((ch0 + i) ^ (r & 7)) << 4)) = o;
            }
          }
          fence_proxy_async();
          mbar_arrive(&out_ready[hp]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ DMA warp
    int it = 0;
    int pending = -1;                                  // staging half of the newest committed store (not yet released)
    if (lane == 0 && (int)blockIdx.x < p.m_tiles) {
      mbar_arrive_expect_tx(x_full, 4 * LL_SUB);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(sX + kb * LL_SUB, &tmX, x_full, kb * 64, blockIdx.x * LL_BM);
    }
    __syncwarp();
    for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
      for (int j = 0; j < NB; ++j) {
        const int g = it * NB + j;
        if (j == NB - 1) {
          // the x buffer is free once the last column block's MMAs have retired: fetch the next tile now
          const int nt = tile + gridDim.x;
          if (nt < p.m_tiles) {
            mbar_wait(sx_free, it & 1);
            if (lane == 0) {
              mbar_arrive_expect_tx(x_full, 4 * LL_SUB);
              for (int kb = 0; kb < 4; ++kb) tma_load_2d(sX + kb * LL_SUB, &tmX, x_full, kb * 64, nt * LL_BM);
            }
            __syncwarp();
          }
        }
        for (int hp = 0; hp < 2; ++hp) {
          mbar_wait(&out_ready[hp], g & 1);
          if (lane == 0) {
            for (int q = 0; q < 2; ++q)
              tma_store_2d(&tmO, sO + (2 * hp + q) * LL_SUB, j * LL_C + hp * 128 + q * 64, tile * LL_BM);
            bulk_commit();
            if (pending >= 0) {
              bulk_wait_read<1>();                       // the store before this one has read its half
              mbar_arrive(&so_free[pending]);
            }
          }
          pending = hp;
          __syncwarp();
        }
      }
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

static int ll_enc2d(CUtensorMap* map, const void* base, int ld, long long rows, int cols, int box_rows) {
  return tmap_rows_bf16(map, base, ld, rows, cols, box_rows);
}

}  // namespace pgt

using namespace pgt;

extern "C" int pgt_ln_linear_bf16(const void* x, int ldx, int T, int C, const float* ln_g, const float* ln_b, float eps,
                                  const void* W, int ldw, int N, const float* bias, void* out, int ldo, void* stream) {
  PGT_CHECK_ARG(x && out && ln_g && ln_b && W && bias && T > 0 && N > 0);
  if (C != LL_C || (N % LL_C) != 0) return PGT_ERR_UNSUPPORTED;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  PGT_CHECK_ARG(al(x) && al(out) && al(W) && ldx % 8 == 0 && ldo % 8 == 0 && ldw % 8 == 0 && ldw >= C && ldo >= N);
  CUtensorMap tx, to, tw;
  int rc = ll_enc2d(&tx, x, ldx, T, C, LL_BM);
  if (rc == PGT_OK) rc = ll_enc2d(&to, out, ldo, T, N, LL_BM);
  if (rc == PGT_OK) rc = ll_enc2d(&tw, W, ldw, N, C, LL_C);
  if (rc != PGT_OK) return rc;
  static PerDeviceOnce once;
  PGT_CUDA_OK(once.run([] { return cudaFuncSetAttribute(ln_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LL_SMEM); }));
  LnLinearParams p{};
  p.T = T; p.m_tiles = ceil_div(T, LL_BM); p.nb = N / LL_C;
  p.ln_g = ln_g; p.ln_b = ln_b; p.eps = eps; p.bias = bias;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
  {
    char desc[64];
    if (prof_enabled()) snprintf(desc, sizeof(desc), "ln_linear M%d N%d K256", T, N);
    ProfScope ps(PGT_PROF_GEMM, 2.0 * (double)T * N * C, st, desc);
    ln_linear_kernel<<<grid, LL_THREADS, LL_SMEM, st>>>(tx, to, tw, p);
  }
  PGT_LAUNCH_OK();
  return PGT_OK;
}
