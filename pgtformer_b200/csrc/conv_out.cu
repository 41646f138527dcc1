// decoder.norm_out -> SiLU -> conv_out (3x3, 64 -> 3 channels, fp32 NCHW result) as ONE tcgen05 kernel
// (`archs/pgtformer_arch.py:707-710`, `archs/tdcrqvae3_arch.py:700-706`: GroupNorm(32, eps 1e-6), swish, Conv2d).
//
// Why its own kernel: with Cout = 3 the generic halo conv computes a 64-wide tile for 3 useful columns and stores
// fp32 NCHW through its slow path (2.3 ms per 16 clips of 512^2), after a separate GroupNorm apply pass over the
// largest tensor of the model (0.5 ms, 3.2 GB).  Here:
//   * 8 builder warps read the RAW conv input once (128-bit loads, a pixel's 64 channels = one 128-byte row), apply
//     y = silu(x * a[f,c] + b[f,c]) (GroupNorm folded into per-(frame, channel) affine terms) and write the bf16
//     (16+2) x (8+2)-pixel halo slab straight into the 128B-swizzled layout — the normalised tensor never exists in HBM;
//   * one elected lane issues, per 16x8-pixel tile, 9 taps x 4 k-steps of tcgen05.mma 128x16x16: tap (dy,dx) is the
//     same slab viewed from row dy*10+dx (SBO = slab pitch, as in the halo conv of gemm_tc.cu), the 9 x 16 x 64 weight
//     tile is resident in shared memory;
//   * 4 epilogue warps read the 128 x 16 accumulator from TMEM, add the bias and store the 3 planes (fp32 NCHW).
// The tile is bound by the tensor core's shared-memory operand reads (4.5 KB per MMA), ~3.5x faster than before.
#include <cstdio>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace pgt {

constexpr int CO_TW = 8, CO_TH = 16;                         // output tile: 16 rows x 8 pixels = 128 GEMM rows
constexpr int CO_SW = CO_TW + 2, CO_SH = CO_TH + 2;          // halo slab
constexpr int CO_PITCH = CO_SW * 128;                        // bytes between slab image rows
constexpr int CO_SLAB = ((CO_SH * CO_PITCH + 1023) / 1024) * 1024;   // 23552
constexpr int CO_NB = 16;                                    // padded Cout (UMMA N)
constexpr int CO_WBYTES = 9 * CO_NB * 128;                   // 18 KB: [tap][16 rows][64 ch] bf16, K-major SW128
constexpr int CO_BUILDERS = 256;
constexpr int CO_THREADS = 32 + CO_BUILDERS + 128;
constexpr int CO_SMEM = 2 * CO_SLAB + CO_WBYTES + 2 * 128 * 4 /*ab of the tile's frame, double buffered*/ + 256 + 1024;

__device__ __forceinline__ uint64_t co_desc_sbo(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void tmem_ld_32x16b(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// silu(v) = v * sigmoid(v) = h + h * tanh(h), h = v / 2: one MUFU (tanh.approx, rel. error 2^-11, below the bf16 rounding
// of the result) instead of the ex2 + rcp pair — the builders are MUFU-bound
__device__ __forceinline__ float silu_tanh(float v) {
  const float h = 0.5f * v;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

struct ConvOutParams {
  const __nv_bfloat16* x;      // [F, H, W, 64] raw (pre-norm) input, pixel pitch ldx elements
  int ldx, F, H, W, cout;
  const float* ab;             // [F][2][64]
  const __nv_bfloat16* w;      // [cout][9 * 64] packed (tap-major), row pitch ldw
  int ldw;
  const float* bias;           // [cout] or null
  float* out;                  // [F, cout, H, W]
  int tiles_x, tiles_y, num_tiles;
};

__global__ void __launch_bounds__(CO_THREADS, 2)
conv_out_gn_kernel(const ConvOutParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* slab = smem;                                         // [2][CO_SLAB]
  uint8_t* sW = smem + 2 * CO_SLAB;                             // [9][16 x 128 B]
  float* sAB = reinterpret_cast<float*>(sW + CO_WBYTES);        // [2][2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sAB + 2 * 128);
  uint64_t* slab_full = bars;                                   // [2]
  uint64_t* slab_empty = bars + 2;                              // [2]
  uint64_t* acc_full = bars + 4;                                // [2]
  uint64_t* acc_empty = bars + 6;                               // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&slab_full[i], CO_BUILDERS);
        mbar_init(&slab_empty[i], 1);
        mbar_init(&acc_full[i], 1);
        mbar_init(&acc_empty[i], 128);
      }
      fence_barrier_init();
    }
    tmem_alloc<32>(tmem_ptr);
    tc_fence_before();
  }
  // weights -> swizzled K-major tiles (rows >= cout are zero)
  for (int i = threadIdx.x; i < 9 * CO_NB * 8; i += CO_THREADS) {
    const int tap = i / (CO_NB * 8), r = (i / 8) % CO_NB, ch = i % 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < p.cout) v = __ldg(reinterpret_cast<const uint4*>(p.w + (size_t)r * p.ldw + tap * 64 + ch * 8));
    *reinterpret_cast<uint4*>(sW + tap * (CO_NB * 128) + r * 128 + ((ch ^ (r & 7)) << 4)) = v;
  }
  fence_proxy_async();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int per_frame = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, CO_NB);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      mbar_wait(&slab_full[buf], ph);
      mbar_wait(&acc_empty[buf], ph ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(slab + buf * CO_SLAB);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint64_t da = co_desc_sbo(sa + ((tap / 3) * CO_SW + (tap % 3)) * 128, CO_PITCH);
          const uint64_t db = umma_desc_k_sw128(smem_u32(sW + tap * (CO_NB * 128)));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + buf * CO_NB, da + 2 * k, db + 2 * k, idesc, (tap | k) != 0 ? 1u : 0u);
        }
        umma_commit(&slab_empty[buf]);
        umma_commit(&acc_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp <= 8) {
    // ------------------------------------------------------------------ slab builders (GroupNorm + SiLU on the way in)
    const int bt = threadIdx.x - 32;                             // 0..255
    const int chunk = bt & 7;                                    // 8 channels = one 16-byte chunk
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int f = tile / per_frame, r = tile - f * per_frame;
      const int y0 = (r / p.tiles_x) * CO_TH - 1, x0 = (r % p.tiles_x) * CO_TW - 1;
      float a[8], b[8];
      {
        const float4* pa = reinterpret_cast<const float4*>(p.ab + (size_t)f * 128 + chunk * 8);
        const float4 a0 = __ldg(pa), a1 = __ldg(pa + 1), b0 = __ldg(pa + 16), b1 = __ldg(pa + 17);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
      }
      // loads first (6 independent 128-bit requests per thread), then wait for the slab, then transform + store
      constexpr int NPIX = CO_SH * CO_SW;                        // 180
      constexpr int ITERS = (NPIX * 8 + CO_BUILDERS - 1) / CO_BUILDERS;   // 6
      uint4 raw[ITERS];
      bool inside[ITERS];
#pragma unroll
      for (int i = 0; i < ITERS; ++i) {
        const int pix = (bt >> 3) + i * (CO_BUILDERS / 8);
        const int sy = pix / CO_SW, sx = pix - sy * CO_SW;
        const int y = y0 + sy, x = x0 + sx;
        inside[i] = pix < NPIX && y >= 0 && y < p.H && x >= 0 && x < p.W;
        raw[i] = make_uint4(0, 0, 0, 0);
        if (inside[i]) raw[i] = __ldg(reinterpret_cast<const uint4*>(p.x + ((size_t)(f * p.H + y) * p.W + x) * p.ldx + chunk * 8));
      }
      mbar_wait(&slab_empty[buf], ph ^ 1);
      uint8_t* sl = slab + buf * CO_SLAB;
#pragma unroll
      for (int i = 0; i < ITERS; ++i) {
        const int pix = (bt >> 3) + i * (CO_BUILDERS / 8);
        if (pix < NPIX) {
          uint4 o = make_uint4(0, 0, 0, 0);                      // zero padding applies to the activated tensor
          if (inside[i]) {
            const uint32_t u[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
            uint32_t q[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 xv = unpack_bf16x2(u[e]);
              const float v0 = fmaf(xv.x, a[2 * e], b[2 * e]), v1 = fmaf(xv.y, a[2 * e + 1], b[2 * e + 1]);
              q[e] = pack_bf16x2(silu_tanh(v0), silu_tanh(v1));
            }
            o = make_uint4(q[0], q[1], q[2], q[3]);
          }
          *reinterpret_cast<uint4*>(sl + pix * 128 + ((chunk ^ (pix & 7)) << 4)) = o;
        }
      }
      fence_proxy_async();
      mbar_arrive(&slab_full[buf]);
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> + bias -> fp32 NCHW planes
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                            // GEMM row = pixel (ty = row / 8, tx = row % 8)
    const uint32_t tacc = tmem_base + (uint32_t(quad * 32) << 16);
    float bias[3] = {0.f, 0.f, 0.f};
    for (int c = 0; c < p.cout && c < 3; ++c) bias[c] = p.bias ? __ldg(p.bias + c) : 0.f;
    const size_t plane = (size_t)p.H * p.W;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int f = tile / per_frame, r = tile - f * per_frame;
      const int y = (r / p.tiles_x) * CO_TH + (row >> 3), x = (r % p.tiles_x) * CO_TW + (row & 7);
      mbar_wait(&acc_full[buf], ph);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld_32x16b(tacc + buf * CO_NB, v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
      float* o = p.out + (size_t)f * p.cout * plane + (size_t)y * p.W + x;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (c < p.cout) o[c * plane] = __uint_as_float(v[c]) + bias[c];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<32>(tmem_base);
  }
}

}  // namespace pgt

using namespace pgt;

extern "C" int pgt_conv_out_gn(const void* x, int F, int H, int W, int Cin, int ldx, const float* gn_ab, const void* Wp,
                               int ldw, int Cout, const float* bias, float* out, void* stream) {
  PGT_CHECK_ARG(x && gn_ab && Wp && out && F > 0);
  if (Cin != 64 || Cout < 1 || Cout > 3 || H % CO_TH != 0 || W % CO_TW != 0 || ldx % 8 != 0 || ldw % 8 != 0)
    return PGT_ERR_UNSUPPORTED;
  PGT_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wp) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(gn_ab) & 15) == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ConvOutParams p{};
  p.x = reinterpret_cast<const __nv_bfloat16*>(x); p.ldx = ldx; p.F = F; p.H = H; p.W = W; p.cout = Cout;
  p.ab = gn_ab; p.w = reinterpret_cast<const __nv_bfloat16*>(Wp); p.ldw = ldw; p.bias = bias; p.out = out;
  p.tiles_x = W / CO_TW; p.tiles_y = H / CO_TH; p.num_tiles = F * p.tiles_x * p.tiles_y;
  static PerDeviceOnce once;
  PGT_CUDA_OK(once.run([] { return cudaFuncSetAttribute(conv_out_gn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CO_SMEM); }));
  const int grid = p.num_tiles < 2 * num_sms() ? p.num_tiles : 2 * num_sms();   // 2 CTAs per SM (66 KB smem, 32 TMEM columns each)
  char desc[64];
  snprintf(desc, sizeof(desc), "conv_out_gn F%d H%d W%d N%d", F, H, W, Cout);
  ProfScope ps(PGT_PROF_GEMM, 2.0 * F * (double)H * W * Cout * 9 * Cin, st, desc);
  conv_out_gn_kernel<<<grid, CO_THREADS, CO_SMEM, st>>>(p);
  PGT_LAUNCH_OK();
  return PGT_OK;
}
