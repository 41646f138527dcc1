// Nearest-codebook L2 argmin (VQEmbedding.compute_distances + find_nearest_embedding, archs/tdcrqvae3_arch.py:99-126)
// on tcgen05, exact by construction.
//
//   argmin_k ||z - e_k||^2  =  argmin_k ( ||e_k||^2 - 2 z.e_k )        (||z||^2 is constant per token)
//
// Pass 1 (tensor cores): the scores  s~ = z~ . e~_k  of the bf16-rounded operands for all K codes, 128 tokens per CTA:
//   * warps 2..9 read the CTA's 128 z rows ONCE from HBM with 128-bit loads, round to bf16 and write the K-major
//     128B-swizzled A tile (128 x E, resident in shared memory for the whole sweep); the same pass yields ||z|| and
//     ||z - z~|| per token (needed for the certificate below);
//   * warp 0 streams the bf16 codebook (N-tile = 256 codes, k-block = 64) through a TMA ring, warp 1 issues
//     tcgen05.mma 128x256x16 into a double-buffered 128 x 256 fp32 accumulator in TMEM;
//   * the two epilogue warpgroups alternate N-tiles: thread = token row, d~_k = ||e_k||^2 - 2 s~ for its 256 codes,
//     running minimum in a register; every code with d~_k <= running min + W is appended to the row's candidate list,
//     which is emptied whenever the minimum drops by more than W (a superset of the final window
//     {k : d~_k <= min d~ + W}; ~ln K appends per row, so the scan stays 3-4 instructions per code, warp divergence is
//     rare and the list holds a handful of entries).
// Certificate: |d~_k - d_k| <= D := 2 (||z - z~|| max||e~|| + ||z|| max||e - e~||) + slack (Cauchy-Schwarz on the two
//   rounding-error dot products + fp32 accumulation slack), hence the true argmin lies in {k : d~_k <= min d~ + 2D},
//   W = 2D.  The window's members (usually ONE) are re-evaluated exactly — fp32 direct sums, whose relative error is
//   bounded by 23 ulp, decide unless two candidates are closer than that bound, in which case fp64 decides (lowest
//   index on ties).  Rows whose window holds more than 16 codes or whose list overflowed (degenerate codebooks: many
//   duplicated / zero rows) are appended to a list for the exhaustive fp32+fp64 kernel in codebook.cu.
// The result therefore equals an fp64 argmin of ||z - e_k||^2 with first-index tie-break for every input.
#include <float.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace pgt {

constexpr int LT_BM = 128;                       // tokens per CTA
constexpr int LT_BN = 256;                       // codes per N-tile
constexpr int LT_BK = 64;                        // k-block (one 128-byte swizzled row)
constexpr int LT_NST = 3;                        // codebook ring depth
constexpr int LT_EMAX = 512;                     // A tile resident: 128 x 512 bf16 = 128 KB
constexpr int LT_A_KB_BYTES = LT_BM * 128;       // 16 KB per k-block of A
constexpr int LT_B_STAGE = LT_BN * 128;          // 32 KB
constexpr int LT_THREADS = 64 + 256;
constexpr int LT_TOP = 16;                       // window members resolved in-kernel (more -> exhaustive kernel)
constexpr int LT_LIST = 16;                      // candidate-list capacity per (token, warpgroup)
constexpr int LT_SMEM = LT_BM * LT_EMAX * 2 + LT_NST * LT_B_STAGE + 1280 /*barriers, z norms*/ + 1024 /*align*/;

__device__ __forceinline__ float4 ld_nc_f4(const float4* p) {       // streaming 128-bit load (read once, keep out of L1)
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// ------------------------------------------------------------------------------ codebook pack (load time)
// bf16 copy of the codebook, ||e_k||^2 (fp64 sum rounded to fp32) and the two maxima the certificate needs:
// norm[K] = max_k ||e~_k||, norm[K+1] = max_k ||e_k - e~_k|| (written as squared maxima, finalised by the caller kernel).
__global__ void __launch_bounds__(256)
codebook_pack_kernel(const float* __restrict__ cb, int K, int E, __nv_bfloat16* __restrict__ cb16, float* __restrict__ norm) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= K) return;
  double s = 0.0;
  float sr = 0.f, sd = 0.f;
  for (int e = lane; e < E; e += 32) {
    const float v = cb[(size_t)row * E + e];
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    const float vb = __bfloat162float(b);
    cb16[(size_t)row * E + e] = b;
    s += (double)v * (double)v;
    sr = fmaf(vb, vb, sr);
    const float dd = v - vb;
    sd = fmaf(dd, dd, sd);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
    sd += __shfl_xor_sync(0xffffffffu, sd, o);
  }
  if (lane == 0) {
    norm[row] = (float)s;
    // non-negative floats order like their bit patterns
    atomicMax(reinterpret_cast<int*>(norm + K), __float_as_int(sr * 1.0001f));
    atomicMax(reinterpret_cast<int*>(norm + K + 1), __float_as_int(sd * 1.0001f));
  }
}

// ------------------------------------------------------------------------------ the sweep
__global__ void __launch_bounds__(LT_THREADS, 1)
l2_argmin_tc_kernel(const __grid_constant__ CUtensorMap tmB, const float* __restrict__ z, int T, int E,
                    const float* __restrict__ cb, const float* __restrict__ norm, int K, int64_t* __restrict__ idx,
                    float* __restrict__ quant, int* __restrict__ fb_count, int* __restrict__ fb_list,
                    float2* __restrict__ scratch /*[T][2][LT_LIST] (d~, code)*/) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                          // [E/64][128 rows][128 B]
  uint8_t* sB = sA + LT_BM * LT_EMAX * 2;                      // [NST][256 rows][128 B]
  uint8_t* tail = sB + LT_NST * LT_B_STAGE;
  uint64_t* b_full = reinterpret_cast<uint64_t*>(tail);        // [NST]
  uint64_t* b_empty = b_full + LT_NST;                         // [NST]
  uint64_t* t_full = b_empty + LT_NST;                         // [2]
  uint64_t* t_empty = t_full + 2;                              // [2]
  uint64_t* a_full = t_empty + 2;                              // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_full + 1);
  float* zn = reinterpret_cast<float*>(tail + 128);            // [128] ||z||^2
  float* dzn = zn + LT_BM;                                     // [128] ||z - z~||^2
  // after the sweep the codebook ring is dead: the warpgroups exchange their scan results through it
  float* xmin = reinterpret_cast<float*>(sB);                  // [2][128] running minima
  int* xcnt = reinterpret_cast<int*>(sB + 1024);               // [2][128] list lengths (> LT_LIST: overflowed)
  int* mi = reinterpret_cast<int*>(sB + 2048);                 // [128][LT_TOP] window members
  int* ncand = reinterpret_cast<int*>(sB + 2048 + LT_BM * LT_TOP * 4);   // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * LT_BM;
  const int KB = E / LT_BK;                                    // k-blocks
  const int NT = K / LT_BN;                                    // N-tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < LT_NST; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 128); }
    mbar_init(a_full, 256);
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
    // ---------------------------------------------------------------- TMA producer: codebook tiles
    int st = 0;
    uint32_t ph = 0;
    for (int nt = 0; nt < NT; ++nt) {
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&b_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&b_full[st], LT_B_STAGE);
          tma_load_2d(sB + st * LT_B_STAGE, &tmB, &b_full[st], kb * LT_BK, nt * LT_BN);
        }
        __syncwarp();
        if (++st == LT_NST) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(LT_BM, LT_BN);
    mbar_wait(a_full, 0);
    tc_fence_after();
    int st = 0;
    uint32_t ph = 0;
    for (int nt = 0; nt < NT; ++nt) {
      const int buf = nt & 1, use = nt >> 1;
      if (use > 0) { mbar_wait(&t_empty[buf], (use - 1) & 1); tc_fence_after(); }
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&b_full[st], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = umma_desc_k_sw128(smem_u32(sA + kb * LT_A_KB_BYTES));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sB + st * LT_B_STAGE));
#pragma unroll
          for (int k = 0; k < LT_BK / 16; ++k)
            umma_bf16_ss(tmem_base + buf * LT_BN, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&b_empty[st]);
          if (kb == KB - 1) umma_commit(&t_full[buf]);
        }
        __syncwarp();
        if (++st == LT_NST) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // ---------------------------------------------------------------- 8 worker warps
    const int w8 = warp - 2;                                   // 0..7
    // (1) z rows -> bf16 A tile.  Warp w8 converts rows w8, w8+8, ...; a lane covers 8 consecutive floats per
    //     256-float step (two 128-bit loads), i.e. exactly one 16-byte chunk of the swizzled row.  Four rows are
    //     loaded before any is converted: 16 independent 128-bit loads in flight per lane.
    constexpr int RB = 4;
    const int nstep = (E + 255) / 256;                         // 1 or 2
    for (int rb = w8; rb < LT_BM; rb += 8 * RB) {
      float4 va[RB][2][2];
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int t = t0 + rb + 8 * i;
#pragma unroll
        for (int sp = 0; sp < 2; ++sp) {
          const int c = sp * 256 + lane * 8;
          if (sp < nstep && t < T && c < E) {
            const float4* p = reinterpret_cast<const float4*>(z + (size_t)t * E + c);
            va[i][sp][0] = ld_nc_f4(p);
            va[i][sp][1] = ld_nc_f4(p + 1);
          } else {
            va[i][sp][0] = va[i][sp][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int r = rb + 8 * i;
        float s2 = 0.f, d2 = 0.f;
#pragma unroll
        for (int sp = 0; sp < 2; ++sp) {
          const int c = sp * 256 + lane * 8;
          if (sp < nstep && c < E) {
            const float4 a = va[i][sp][0], b = va[i][sp][1];
            const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
            uint32_t u[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              u[q] = pack_bf16x2(f[2 * q], f[2 * q + 1]);
              const float2 back = unpack_bf16x2(u[q]);
              s2 = fmaf(f[2 * q], f[2 * q], s2); s2 = fmaf(f[2 * q + 1], f[2 * q + 1], s2);
              const float e0 = f[2 * q] - back.x, e1 = f[2 * q + 1] - back.y;      // exact in fp32
              d2 = fmaf(e0, e0, d2); d2 = fmaf(e1, e1, d2);
            }
            const int kb = c >> 6, chunk = (c & 63) >> 3;
            *reinterpret_cast<uint4*>(sA + kb * LT_A_KB_BYTES + r * 128 + ((chunk ^ (r & 7)) << 4)) = make_uint4(u[0], u[1], u[2], u[3]);
          }
        }
        s2 = warp_sum(s2); d2 = warp_sum(d2);
        if (lane == 0) { zn[r] = s2; dzn[r] = d2; }
      }
    }
    fence_proxy_async();                                       // generic-proxy writes of A -> visible to tcgen05.mma
    mbar_arrive(a_full);
    named_bar_sync(1, 256);                                    // zn / dzn of every row are in shared memory

    // (2) scan: warpgroup g owns accumulator buffer g (N-tiles g, g+2, ..); thread = token row
    const int g = w8 >> 2;
    const int quad = warp & 3;                                 // TMEM lane quadrant this warp may read
    const int r = quad * 32 + lane;
    const uint32_t tacc = tmem_base + (uint32_t(quad * 32) << 16) + g * LT_BN;
    float W;
    {
      const float emax = sqrtf(__ldg(norm + K)), demax = sqrtf(__ldg(norm + K + 1));
      const float zl = sqrtf(zn[r]) * 1.0001f, dzl = sqrtf(dzn[r]) * 1.0001f;
      const float D = 2.f * (dzl * emax + zl * demax) + zl * emax * (1.f / 4096.f) + 1e-30f;
      W = 2.f * D * 1.001f;
    }
    float2* mylist = scratch + ((size_t)(t0 + r) * 2 + g) * LT_LIST;
    const bool live = t0 + r < T;
    float runmin = FLT_MAX, thr = FLT_MAX;
    int cnt = 0;
    for (int nt = g, use = 0; nt < NT; nt += 2, ++use) {
      mbar_wait(&t_full[g], use & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < LT_BN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tacc + c, v);
        const float4* np = reinterpret_cast<const float4*>(norm + nt * LT_BN + c);
        float nv[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 n4 = __ldg(np + q);                     // same address in every lane: one broadcast transaction
          nv[4 * q] = n4.x; nv[4 * q + 1] = n4.y; nv[4 * q + 2] = n4.z; nv[4 * q + 3] = n4.w;
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = fmaf(-2.f, __uint_as_float(v[i]), nv[i]);
          if (d <= thr) {
            if (d < runmin) {
              if (d + W < runmin) cnt = 0;                     // every earlier entry is now outside any final window
              runmin = d; thr = d + W;
            }
            if (cnt < LT_LIST && live) mylist[cnt] = make_float2(d, __int_as_float(nt * LT_BN + c + i));
            ++cnt;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&t_empty[g]);
    }
    // (3) exchange between the warpgroups (the ring is dead: every MMA has completed once the last t_full fired)
    named_bar_sync(1, 256);
    xmin[g * LT_BM + r] = runmin;
    xcnt[g * LT_BM + r] = cnt;
    named_bar_sync(1, 256);                                    // also orders the candidate-list stores (CTA scope)
    if (g == 0) {
      const float m = fminf(xmin[r], xmin[LT_BM + r]);
      const float win = m + W;
      int n = 0;
      bool ovf = false;
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
        const int cn = xcnt[gg * LT_BM + r];
        if (cn > LT_LIST) ovf = true;
        const float2* lst = scratch + ((size_t)(t0 + r) * 2 + gg) * LT_LIST;
        for (int j = 0; j < min(cn, LT_LIST) && live; ++j) {
          const float2 e = lst[j];
          if (e.x <= win) {
            if (n < LT_TOP) mi[r * LT_TOP + n] = __float_as_int(e.y);
            ++n;
          }
        }
      }
      ncand[r] = (ovf || n > LT_TOP) ? -1 : n;
    }
    named_bar_sync(1, 256);
    // (4) exact resolution: one warp per token
    for (int rr = w8; rr < LT_BM; rr += 8) {
      const int t = t0 + rr;
      if (t >= T) continue;
      const int n = ncand[rr];
      int best = -1;
      if (n < 0) {
        if (lane == 0) fb_list[atomicAdd(fb_count, 1)] = t;
        continue;
      } else if (n == 1) {
        best = mi[rr * LT_TOP];
      } else {
        float zr[LT_EMAX / 32];
#pragma unroll
        for (int q = 0; q < LT_EMAX / 128; ++q) {
          if (q * 128 + lane * 4 < E) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(z + (size_t)t * E) + q * 32 + lane);
            zr[4 * q] = a.x; zr[4 * q + 1] = a.y; zr[4 * q + 2] = a.z; zr[4 * q + 3] = a.w;
          } else {
            zr[4 * q] = zr[4 * q + 1] = zr[4 * q + 2] = zr[4 * q + 3] = 0.f;
          }
        }
        float b1 = FLT_MAX, b2 = FLT_MAX;                      // best and second-best fp32 distances
        for (int j = 0; j < n; ++j) {
          const int k = mi[rr * LT_TOP + j];
          float s = 0.f;
#pragma unroll
          for (int q = 0; q < LT_EMAX / 128; ++q) {
            if (q * 128 + lane * 4 < E) {
              const float4 e = __ldg(reinterpret_cast<const float4*>(cb + (size_t)k * E) + q * 32 + lane);
              float d;
              d = zr[4 * q] - e.x; s = fmaf(d, d, s);
              d = zr[4 * q + 1] - e.y; s = fmaf(d, d, s);
              d = zr[4 * q + 2] - e.z; s = fmaf(d, d, s);
              d = zr[4 * q + 3] - e.w; s = fmaf(d, d, s);
            }
          }
          s = warp_sum(s);
          if (s < b1 || (s == b1 && k < best)) { b2 = b1; b1 = s; best = k; }
          else if (s < b2) b2 = s;
        }
        // fp32 sums of non-negative terms: relative error <= (2 + 16 + 5) ulp ~ 1.4e-6 each; closer than that -> fp64
        if (!(b1 * (1.f + 4e-6f) < b2)) {
          double bd = 0.0;
          best = -1;
          for (int j = 0; j < n; ++j) {
            const int k = mi[rr * LT_TOP + j];
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < LT_EMAX / 128; ++q) {
              if (q * 128 + lane * 4 < E) {
                const float4 e = __ldg(reinterpret_cast<const float4*>(cb + (size_t)k * E) + q * 32 + lane);
                double d;
                d = (double)zr[4 * q] - (double)e.x; s += d * d;
                d = (double)zr[4 * q + 1] - (double)e.y; s += d * d;
                d = (double)zr[4 * q + 2] - (double)e.z; s += d * d;
                d = (double)zr[4 * q + 3] - (double)e.w; s += d * d;
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (best < 0 || s < bd || (s == bd && k < best)) { bd = s; best = k; }
          }
        }
      }
      if (lane == 0) idx[t] = best;
      if (quant != nullptr) {
        const float4* src = reinterpret_cast<const float4*>(cb + (size_t)best * E);
        float4* dst = reinterpret_cast<float4*>(quant + (size_t)t * E);
        for (int e = lane; e < (E >> 2); e += 32) dst[e] = __ldg(src + e);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

int l2_argmin_list_launch(const float* z, int T, int E, const float* codebook, int K, int64_t* idx, float* quant,
                          const int* list, const int* count, int grid, cudaStream_t st);      // codebook.cu

}  // namespace pgt

using namespace pgt;

extern "C" int pgt_codebook_pack(const float* codebook, int K, int E, void* cb_bf16, float* cb_norm, void* stream) {
  PGT_CHECK_ARG(codebook && cb_bf16 && cb_norm && K > 0 && E > 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PGT_CUDA_OK(cudaMemsetAsync(cb_norm + K, 0, 2 * sizeof(float), st));
  codebook_pack_kernel<<<ceil_div(K, 8), 256, 0, st>>>(codebook, K, E, reinterpret_cast<__nv_bfloat16*>(cb_bf16), cb_norm);
  PGT_LAUNCH_OK();
  return PGT_OK;
}

extern "C" int64_t pgt_l2_argmin_ws_ints(int T) {
  return 2 + (int64_t)((T + 1) / 2) * 2 + (int64_t)T * 2 * LT_LIST * 2;
}

extern "C" int pgt_l2_argmin_tc(const float* z, int T, int E, const float* codebook, const void* cb_bf16,
                                const float* cb_norm, int K, int64_t* idx, float* quant, int32_t* workspace,
                                void* stream) {
  PGT_CHECK_ARG(z && codebook && cb_bf16 && cb_norm && idx && workspace && T > 0);
  if (K % LT_BN != 0 || E % LT_BK != 0 || E > LT_EMAX || E % 128 != 0) return PGT_ERR_UNSUPPORTED;
  PGT_CHECK_ARG((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(codebook) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(cb_bf16) & 15) == 0 && (quant == nullptr || (reinterpret_cast<uintptr_t>(quant) & 15) == 0));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tmB;
  const uint64_t dims[2] = {(uint64_t)E, (uint64_t)K};
  const uint64_t strides[1] = {(uint64_t)E * 2};
  const uint32_t box[2] = {LT_BK, LT_BN};
  int rc = tmap_encode(&tmB, cb_bf16, 2, dims, strides, box);
  if (rc != PGT_OK) return rc;
  static PerDeviceOnce once;
  PGT_CUDA_OK(once.run([] { return cudaFuncSetAttribute(l2_argmin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM); }));
  PGT_CUDA_OK(cudaMemsetAsync(workspace, 0, sizeof(int32_t), st));
  {
    ProfScope ps(PGT_PROF_ARGMIN, 2.0 * T * (double)K * E, st, "l2_argmin_tc");
    l2_argmin_tc_kernel<<<ceil_div(T, LT_BM), LT_THREADS, LT_SMEM, st>>>(
        tmB, z, T, E, codebook, cb_norm, K, idx, quant, workspace, workspace + 2,
        reinterpret_cast<float2*>(workspace + 2 + ((T + 1) / 2) * 2));
    PGT_LAUNCH_OK();
  }
  // tokens whose certificate window did not fit the shortlist (degenerate codebooks): exhaustive exact kernel
  return l2_argmin_list_launch(z, T, E, codebook, K, idx, quant, workspace + 2, workspace, num_sms(), st);
}
