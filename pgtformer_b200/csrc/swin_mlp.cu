// Fused Swin MLP half-block on tcgen05 (C = 256):   out = x + W2 * GELU(W1 * LN(x) + b1) + b2
// Replaces LayerNorm(norm2) + Mlp(fc1, exact GELU, fc2) + residual of VSTSREncoderTransformerBlock
// (modules/rstt_layers.py:116-132,335-336): five HBM passes over the token matrix (LN r/w, fc1 r/w, fc2 r+r/w)
// become two (read x, write out); the 128 x 256 hidden tile never leaves the SM.
//
// The per-tile chain  load -> LN -> GEMM1 -> GELU -> GEMM2 -> +x -> store  is strictly serial, and its ALU phases
// (LN, GELU, residual) cost several times the two MMAs, so each CTA runs TWO tile pipelines ("groups") side by side:
// while one group's warps wait for a TMA load or an MMA, the other group's warps own the issue slots.
//
// Per CTA, persistent over 128-token tiles (640 threads):
//   warp 0        TMA producer of the weight k-blocks (W1 / W2, 256 x 64 each, 3-deep ring shared by both groups)
//   warp 1        tcgen05.mma issuer for both groups in the fixed order  G2(b') G1(a) G2(a) G1(b)  (b half a tile behind
//                 a);  group g accumulates
//                 in TMEM columns [256 g, 256 g + 256) — GEMM2 reuses GEMM1's columns once the GELU pass has read them
//   warp 2+g      DMA of group g: TMA load of the x tile, re-fetch of x (an L2 hit) for the residual once GEMM2 has
//                 consumed the buffer, TMA store of the finished tile
//   warps 4..19   compute, 8 per group: LayerNorm of the tile in place, GELU epilogue written back in place as the
//                 A operand of GEMM2, final epilogue acc + b2 + x in place.  One 64 KB buffer per group carries
//                 x -> LN(x) -> hidden -> x (again) -> out.
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tmap.cuh"
#include "ptx.cuh"
#include "epi_common.cuh"

namespace pgt {

constexpr int SM_C = 256;
constexpr int SM_BM = 128;
constexpr int SM_SUB = SM_BM * 128;            // one [128 x 64] bf16 sub-tile: 16 KB
constexpr int SM_WST = 3;                      // weight ring depth
constexpr int SM_WBYTES = SM_C * 128;          // one [256 x 64] weight k-block: 32 KB
constexpr int SM_THREADS = 640;
constexpr int SM_SMEM = 2 * 4 * SM_SUB /*tiles*/ + SM_WST * SM_WBYTES + 2 * 128 * 8 /*xch*/ + 256;

struct SwinMlpParams {
  int T, m_tiles;
  const float* ln_g;
  const float* ln_b;
  float eps;
  const float* b1;
  const float* b2;
  float* gn_stats;      // optional [m_tiles][4][32][2]
};

struct SmGroupBars {
  uint64_t x_full, res_full, y_ready, acc1_full, h_ready, acc2_full, out_ready;
};

// compute warps of one group
__device__ __forceinline__ void swin_mlp_compute(const SwinMlpParams& p, uint8_t* sX, float2* xch, SmGroupBars* gb,
                                                 uint32_t t_acc, int g, int n_it, int quad, int half, int lane) {
  const int r = quad * 32 + lane;
  const uint32_t t_row = t_acc + (uint32_t(quad * 32) << 16);
  const int c_lo = half * 128;
  for (int it = 0; it < n_it; ++it) {
    const int tile = blockIdx.x + (2 * it + g) * gridDim.x;
    const uint32_t par = it & 1;
    // ---- LayerNorm(x) in place -> A operand of GEMM1 (two threads per row; the halves meet through smem)
    mbar_wait(&gb->x_full, par);
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      const uint8_t* src = sX + (half * 2 + sub) * SM_SUB + r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 u = *reinterpret_cast<const uint4*>(src + ((c ^ (r & 7)) << 4));
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        s += (a.x + a.y) + (b.x + b.y) + (cc.x + cc.y) + (d.x + d.y);
        q += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + cc.x * cc.x + cc.y * cc.y + d.x * d.x + d.y * d.y;
      }
    }
    if (half == 0) xch[r] = make_float2(s, q);
    named_bar_sync(5 + g, 256);
    if (half == 1) {
      const float2 o = xch[r];
      s += o.x; q += o.y;
      xch[r] = make_float2(s, q);
    }
    named_bar_sync(5 + g, 256);
    if (half == 0) {
      const float2 o = xch[r];
      s = o.x; q = o.y;
    }
    const float mean = s * (1.f / SM_C);
    const float var = fmaxf(q * (1.f / SM_C) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + p.eps);
    const float nm = -mean * rstd;
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      const int kb = half * 2 + sub;
      uint8_t* row = sX + kb * SM_SUB + r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4* ptr = reinterpret_cast<uint4*>(row + ((c ^ (r & 7)) << 4));
        const uint4 u = *ptr;
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        const float v[8] = {a.x, a.y, b.x, b.y, cc.x, cc.y, d.x, d.y};
        float y[8];
        if (p.ln_g != nullptr) {
          const int col = kb * 64 + c * 8;
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_g + col)), g1 = __ldg(reinterpret_cast<const float4*>(p.ln_g + col + 4));
          const float4 e0 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col)), e1 = __ldg(reinterpret_cast<const float4*>(p.ln_b + col + 4));
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = fmaf(fmaf(v[j], rstd, nm), gg[j], ee[j]);
        } else {                               // gamma / beta folded into W1 / b1 by the caller
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = fmaf(v[j], rstd, nm);
        }
        uint4 o;
        o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]);
        o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
        *ptr = o;
      }
    }
    fence_proxy_async();
    mbar_arrive(&gb->y_ready);
    // ---- hidden = GELU(acc + b1) -> A operand of GEMM2 (in place: GEMM1 has finished reading the buffer)
    mbar_wait(&gb->acc1_full, par);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = c_lo; c0 < c_lo + 128; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(t_row + c0, v);
      tmem_ld_wait();
      const float4* b4 = reinterpret_cast<const float4*>(p.b1 + c0);
      uint8_t* dst = sX + (c0 >> 6) * SM_SUB + r * 128;
      const int ch0 = (c0 & 63) >> 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 ba = __ldg(b4 + 2 * i), bb = __ldg(b4 + 2 * i + 1);
        uint4 o;
        o.x = pack_bf16x2(gelu_erf(__uint_as_float(v[8 * i + 0]) + ba.x), gelu_erf(__uint_as_float(v[8 * i + 1]) + ba.y));
        o.y = pack_bf16x2(gelu_erf(__uint_as_float(v[8 * i + 2]) + ba.z), gelu_erf(__uint_as_float(v[8 * i + 3]) + ba.w));
        o.z = pack_bf16x2(gelu_erf(__uint_as_float(v[8 * i + 4]) + bb.x), gelu_erf(__uint_as_float(v[8 * i + 5]) + bb.y));
        o.w = pack_bf16x2(gelu_erf(__uint_as_float(v[8 * i + 6]) + bb.z), gelu_erf(__uint_as_float(v[8 * i + 7]) + bb.w));
        *reinterpret_cast<uint4*>(dst + (((ch0 + i) ^ (r & 7)) << 4)) = o;
      }
    }
    tc_fence_before();
    fence_proxy_async();
    mbar_arrive(&gb->h_ready);
    // ---- out = acc + b2 + x in place over the re-fetched x tile (then TMA-stored by the group's DMA warp)
    mbar_wait(&gb->acc2_full, par);
    tc_fence_after();
    mbar_wait(&gb->res_full, par);
#pragma unroll 1
    for (int c0 = c_lo; c0 < c_lo + 128; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(t_row + c0, v);
      tmem_ld_wait();
      float f[32];
      const float4* b4 = reinterpret_cast<const float4*>(p.b2 + c0);
      uint8_t* row = sX + (c0 >> 6) * SM_SUB + r * 128;
      const int ch0 = (c0 & 63) >> 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 ba = __ldg(b4 + 2 * i), bb = __ldg(b4 + 2 * i + 1);
        uint4* dst = reinterpret_cast<uint4*>(row + (((ch0 + i) ^ (r & 7)) << 4));
        const uint4 u = *dst;
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        f[8 * i + 0] = __uint_as_float(v[8 * i + 0]) + ba.x + a.x;  f[8 * i + 1] = __uint_as_float(v[8 * i + 1]) + ba.y + a.y;
        f[8 * i + 2] = __uint_as_float(v[8 * i + 2]) + ba.z + b.x;  f[8 * i + 3] = __uint_as_float(v[8 * i + 3]) + ba.w + b.y;
        f[8 * i + 4] = __uint_as_float(v[8 * i + 4]) + bb.x + cc.x; f[8 * i + 5] = __uint_as_float(v[8 * i + 5]) + bb.y + cc.y;
        f[8 * i + 6] = __uint_as_float(v[8 * i + 6]) + bb.z + d.x;  f[8 * i + 7] = __uint_as_float(v[8 * i + 7]) + bb.w + d.y;
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
    mbar_arrive(&gb->out_ready);
  }
}

__global__ void __launch_bounds__(SM_THREADS, 1)
swin_mlp_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmO,
                const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                const SwinMlpParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];     // no static smem in this kernel: the window starts 1024-aligned
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sW = smem + 2 * 4 * SM_SUB;                  // weight ring behind the two tile buffers
  float2* xch = reinterpret_cast<float2*>(sW + SM_WST * SM_WBYTES);     // [2 groups][128 rows] (sum, sumsq)
  SmGroupBars* gbar = reinterpret_cast<SmGroupBars*>(xch + 256);         // [2]
  uint64_t* w_full = reinterpret_cast<uint64_t*>(gbar + 2);              // [SM_WST]
  uint64_t* w_empty = w_full + SM_WST;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_empty + SM_WST);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmO); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&gbar[g].x_full, 1); mbar_init(&gbar[g].res_full, 1);
      mbar_init(&gbar[g].y_ready, 256); mbar_init(&gbar[g].acc1_full, 1);
      mbar_init(&gbar[g].h_ready, 256); mbar_init(&gbar[g].acc2_full, 1);
      mbar_init(&gbar[g].out_ready, 256);
    }
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

  // this CTA's tiles: blockIdx.x + j * gridDim.x; group g owns the j = 2 i + g
  const int n_loc = (int)blockIdx.x < p.m_tiles ? (p.m_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int n_grp[2] = {(n_loc + 1) >> 1, n_loc >> 1};

  // GEMM order of the CTA (weight producer and MMA issuer walk the same list).  Group b runs half a tile behind group a:
  //   ... G2(b, i-1)  G1(a, i)  G2(a, i)  G1(b, i)  G2(b, i) ...
  // so that one group's TMA load / LayerNorm / store phases fall into the other's GEMM + GELU phases.  With the two groups
  // in phase (G1a G1b G2a G2b) every CTA of the grid loaded, normalised and stored at the same moments, and each x load
  // queued behind 19 MB of simultaneous requests.
  auto for_each_gemm = [&](auto&& fn) {
    for (int it = 0; it < n_grp[0]; ++it) {
      if (it >= 1 && it - 1 < n_grp[1]) fn(1, 1, it - 1);
      fn(0, 0, it);
      fn(1, 0, it);
      if (it < n_grp[1]) fn(0, 1, it);
    }
    if (n_grp[1] >= 1 && n_grp[1] == n_grp[0]) fn(1, 1, n_grp[1] - 1);
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer (same order as the MMA warp)
    int st = 0;
    uint32_t ph = 0;
    for_each_gemm([&](int gemm, int /*g*/, int /*it*/) {
      for (int kb = 0; kb < 4; ++kb) {
        mbar_wait(&w_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&w_full[st], SM_WBYTES);
          tma_load_2d(sW + st * SM_WBYTES, gemm == 0 ? &tmW1 : &tmW2, &w_full[st], kb * 64, 0);
        }
        __syncwarp();
        if (++st == SM_WST) { st = 0; ph ^= 1; }
      }
    });
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(SM_BM, SM_C);
    int st = 0;
    uint32_t ph = 0;
    for_each_gemm([&](int gemm, int g, int it) {
      const uint32_t par = it & 1;
      uint8_t* sX = smem + g * 4 * SM_SUB;
      mbar_wait(gemm == 0 ? &gbar[g].y_ready : &gbar[g].h_ready, par);
      tc_fence_after();
      for (int kb = 0; kb < 4; ++kb) {
        mbar_wait(&w_full[st], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(sX + kb * SM_SUB));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sW + st * SM_WBYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + g * SM_C, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&w_empty[st]);
          if (kb == 3) umma_commit(gemm == 0 ? &gbar[g].acc1_full : &gbar[g].acc2_full);
        }
        __syncwarp();
        if (++st == SM_WST) { st = 0; ph ^= 1; }
      }
    });
  } else if (warp < 4) {
    // ------------------------------------------------------------------ DMA warp of group g
    const int g = warp - 2;
    uint8_t* sX = smem + g * 4 * SM_SUB;
    for (int it = 0; it < n_grp[g]; ++it) {
      const int tile = blockIdx.x + (2 * it + g) * gridDim.x;
      const uint32_t par = it & 1;
      if (lane == 0) {
        mbar_arrive_expect_tx(&gbar[g].x_full, 4 * SM_SUB);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sX + kb * SM_SUB, &tmX, &gbar[g].x_full, kb * 64, tile * SM_BM);
      }
      __syncwarp();
      mbar_wait(&gbar[g].acc2_full, par);        // GEMM2 has consumed the buffer: bring x back for the residual
      if (lane == 0) {
        mbar_arrive_expect_tx(&gbar[g].res_full, 4 * SM_SUB);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sX + kb * SM_SUB, &tmX, &gbar[g].res_full, kb * 64, tile * SM_BM);
      }
      __syncwarp();
      mbar_wait(&gbar[g].out_ready, par);
      if (lane == 0) {
        for (int kb = 0; kb < 4; ++kb) tma_store_2d(&tmO, sX + kb * SM_SUB, kb * 64, tile * SM_BM);
        bulk_commit();
        bulk_wait_read<0>();          // the buffer may be refilled
      }
      __syncwarp();
    }
    if (lane == 0) bulk_wait0();
  } else {
    // ------------------------------------------------------------------ compute warps
    const int g = (warp - 4) >> 3;
    swin_mlp_compute(p, smem + g * 4 * SM_SUB, xch + g * 128, &gbar[g], tmem_base + g * SM_C, g, n_grp[g], warp & 3,
                     ((warp - 4) & 7) >> 2, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static int enc2d(CUtensorMap* map, const void* base, int ld, long long rows, int cols, int box_rows) {
  return tmap_rows_bf16(map, base, ld, rows, cols, box_rows);
}

}  // namespace pgt

using namespace pgt;

extern "C" int pgt_swin_mlp_bf16(const void* x, int ldx, int T, int C, const float* ln_g, const float* ln_b, float eps,
                                 const void* W1, const float* b1, const void* W2, const float* b2, void* out, int ldo,
                                 float* gn_stats, void* stream) {
  PGT_CHECK_ARG(x && out && W1 && W2 && b1 && b2 && T > 0 && ((ln_g == nullptr) == (ln_b == nullptr)));   // both null: no affine
  if (C != SM_C) return PGT_ERR_UNSUPPORTED;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  PGT_CHECK_ARG(al(x) && al(out) && al(W1) && al(W2) && ldx % 8 == 0 && ldo % 8 == 0);
  CUtensorMap tx, to, t1, t2;
  int rc = enc2d(&tx, x, ldx, T, C, SM_BM);
  if (rc == PGT_OK) rc = enc2d(&to, out, ldo, T, C, SM_BM);
  if (rc == PGT_OK) rc = enc2d(&t1, W1, C, C, C, SM_C);
  if (rc == PGT_OK) rc = enc2d(&t2, W2, C, C, C, SM_C);
  if (rc != PGT_OK) return rc;
  static PerDeviceOnce once;
  PGT_CUDA_OK(once.run([] { return cudaFuncSetAttribute(swin_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM); }));
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
