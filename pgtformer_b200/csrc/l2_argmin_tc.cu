// Nearest-codebook L2 argmin (VQEmbedding.compute_distances + find_nearest_embedding, archs/tdcrqvae3_arch.py:99-126)
// on tcgen05, exact by construction.
//
//   argmin_k ||z - e_k||^2  =  argmin_k ( ||e_k||^2 - 2 z.e_k )        (||z||^2 is constant per token)
//
// Pass 0 (z_pack_kernel, HBM-bound): the fp32 z rows are read ONCE with 128-bit streaming loads and written back as bf16
//   rows z~ together with (||z||^2, ||z - z~||^2) per token (needed for the certificate below).
// Pass 1 (l2_argmin_pair_kernel, tensor cores): the scores  s~ = z~ . e~_k  of the bf16-rounded operands for all K codes on
//   CTA PAIRS — M = 256 tokens per tcgen05.mma.cta_group::2, N-tile = 256 codes, k-block = 64:
//   * each CTA keeps its 128 tokens' A tile (128 x E bf16) resident in shared memory and loads HALF of every codebook
//     tile through its TMA ring, so one pass over the bf16 codebook (K x E x 2 B = 1 MB) serves 256 tokens: 4 KB of
//     L2 -> SM traffic per token (a single-CTA sweep pulls 8 KB per token — 384 MB at T = 49 152 — and is bound by it);
//   * persistent: a pair walks over its 256-token tiles; the A k-blocks of the next tile are loaded by their own producer
//     warp (per-k-block barriers) as soon as the last N-tile of the current one has consumed them, and the codebook ring
//     never drains;
//   * two scan warpgroups alternate over the double-buffered 128 x 256 fp32 accumulator in TMEM: thread = token row,
//     d~_k = ||e_k||^2 - 2 s~ for 32 codes at a time, BRANCH-FREE chunk minimum, and ONE test per chunk
//     (chunk min <= running min + W); only then the chunk's codes within W of the (updated) running minimum are appended,
//     with predicated stores, to the row's candidate list in shared memory.  The list is emptied whenever the minimum drops
//     by more than W, so it is a superset of the final window {k : d~_k <= min d~ + W}.  (A test-and-branch per code costs
//     ~46 clk of dependent latency each: measured 23-28 us per tile against 8.6 us of MMAs.)
//   * merge (all 256 scan threads: each filters its own list against the row's global minimum) and exact resolution
//     (below) of tile i run while the tensor pipe already works on tile i + 1.
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

constexpr int LT_BM = 128;                       // tokens per CTA (256 per pair)
constexpr int LT_BN = 256;                       // codes per N-tile
constexpr int LT_BK = 64;                        // k-block (one 128-byte swizzled row)
constexpr int LT_EMAX = 512;                     // A tile resident: 128 x 512 bf16 = 128 KB
constexpr int LT_A_KB_BYTES = LT_BM * 128;       // 16 KB per k-block of A
constexpr int LT_KBMAX = LT_EMAX / LT_BK;        // 8
constexpr int LT_BST = (LT_BN / 2) * 128;        // codebook stage per CTA: 128 codes x 64 k = 16 KB
constexpr int LT_NST = 3;                        // codebook ring depth
constexpr int LT_THREADS = 11 * 32;              // warps: 0 codebook producer, 1 MMA issuer, 2..9 scan / resolve, 10 A producer
constexpr int LT_TOP = 16;                       // window members resolved in-kernel (more -> exhaustive kernel)
constexpr int LT_LIST = 16;                      // candidate-list capacity per (token, warpgroup)
constexpr int LT_XCH = 2 * LT_BM * LT_LIST * 8 /*lists*/ + 2 * LT_BM * 4 /*xmin*/ + LT_BM * LT_TOP * 4 /*mi*/ + 2 * LT_BM * 4 /*ncand, ovf*/;
constexpr int LT_SMEM = LT_BM * LT_EMAX * 2 + LT_NST * LT_BST + LT_XCH + 512 /*barriers*/ + 1024 /*align*/;

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

// ------------------------------------------------------------------------------ z pack (pass 0)
__global__ void __launch_bounds__(256)
z_pack_kernel(const float* __restrict__ z, int T, int E, __nv_bfloat16* __restrict__ zb, float2* __restrict__ zn2) {
  constexpr int RB = 4;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int nstep = (E + 255) / 256;
  for (int t0 = gw * RB; t0 < T; t0 += nw * RB) {
    float4 va[RB][2][2];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
#pragma unroll
      for (int sp = 0; sp < 2; ++sp) {
        const int c = sp * 256 + lane * 8;
        if (sp < nstep && t0 + i < T && c < E) {
          const float4* p = reinterpret_cast<const float4*>(z + (size_t)(t0 + i) * E + c);
          va[i][sp][0] = ld_nc_f4(p);
          va[i][sp][1] = ld_nc_f4(p + 1);
        } else {
          va[i][sp][0] = va[i][sp][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int t = t0 + i;
      if (t >= T) break;
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
          *reinterpret_cast<uint4*>(zb + (size_t)t * E + c) = make_uint4(u[0], u[1], u[2], u[3]);
        }
      }
      s2 = warp_sum(s2); d2 = warp_sum(d2);
      if (lane == 0) zn2[t] = make_float2(s2, d2);
    }
  }
}

// ------------------------------------------------------------------------------ scan / merge / exact resolution
// Scan of one 128 x 256 accumulator buffer, thread = token row (see the header comment).  lst: this thread's candidate
// list in shared memory (lst_s: its shared-window address), entry j 2 * LT_BM entries further (entry-major: the lanes of
// a warp hit distinct banks).
__device__ __forceinline__ void argmin_scan_ntile(uint32_t tacc, const float* __restrict__ norm_nt, int code0, float W,
                                                  uint32_t lst_s, float& runmin, float& thr, int& cnt) {
  uint32_t laddr = lst_s + (uint32_t)cnt * (2 * LT_BM * 8);   // shared-window address of the next list entry
#pragma unroll 1
  for (int c = 0; c < LT_BN; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tacc + c, v);
    const float4* np = reinterpret_cast<const float4*>(norm_nt + c);
    float d[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 n4 = __ldg(np + q);                       // same address in every lane: one broadcast transaction
      d[4 * q] = n4.x; d[4 * q + 1] = n4.y; d[4 * q + 2] = n4.z; d[4 * q + 3] = n4.w;
    }
    tmem_ld_wait();
    float m0 = FLT_MAX, m1 = FLT_MAX, m2 = FLT_MAX, m3 = FLT_MAX;
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      d[i] = fmaf(-2.f, __uint_as_float(v[i]), d[i]);             m0 = fminf(m0, d[i]);
      d[i + 1] = fmaf(-2.f, __uint_as_float(v[i + 1]), d[i + 1]); m1 = fminf(m1, d[i + 1]);
      d[i + 2] = fmaf(-2.f, __uint_as_float(v[i + 2]), d[i + 2]); m2 = fminf(m2, d[i + 2]);
      d[i + 3] = fmaf(-2.f, __uint_as_float(v[i + 3]), d[i + 3]); m3 = fminf(m3, d[i + 3]);
    }
    const float cm = fminf(fminf(m0, m1), fminf(m2, m3));
    if (cm <= thr) {                                         // the only branch of the chunk (dead rows: thr = -FLT_MAX)
      if (cm < runmin) {
        if (cm + W < runmin) { cnt = 0; laddr = lst_s; }     // every earlier entry is now outside any final window
        runmin = cm; thr = cm + W;
      }
      // predicated append of every code of the chunk within W of the running minimum — no branch per code: a
      // test-and-branch costs ~46 clk of dependent latency each (measured), five predicated instructions do not
      const int cbase = code0 + c;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        asm volatile(
            "{\n\t.reg .pred h, p;\n\t"
            "setp.le.f32 h, %2, %3;\n\t"
            "setp.lt.and.s32 p, %0, %5, h;\n\t"
            "@p st.shared.v2.b32 [%1], {%6, %4};\n\t"
            "@h add.s32 %0, %0, 1;\n\t"
            "@h add.u32 %1, %1, %7;\n\t}"
            : "+r"(cnt), "+r"(laddr)
            : "f"(d[i]), "f"(thr), "r"(cbase + i), "n"(LT_LIST), "r"(__float_as_uint(d[i])), "n"(2 * LT_BM * 8)
            : "memory");
      }
    }
  }
}

__device__ __forceinline__ void argmin_load_row(const float* __restrict__ p, int E, int lane, float4 (&r)[LT_EMAX / 128]) {
#pragma unroll
  for (int q = 0; q < LT_EMAX / 128; ++q)
    r[q] = (q * 128 + lane * 4 < E) ? __ldg(reinterpret_cast<const float4*>(p) + q * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// Exact resolution of one token by one warp: the n (> 1) members of its window re-evaluated with fp32 direct sums (the rows
// of up to four candidates are in flight together), fp64 when two of them are closer than the fp32 error bound; lowest
// index on ties.  Returns the code in every lane.
__device__ __forceinline__ int argmin_resolve_exact(const float* __restrict__ z, const float* __restrict__ cb, int E, int t, int n,
                                                    const int* mi_row, int lane) {
  float4 zr[LT_EMAX / 128];
  argmin_load_row(z + (size_t)t * E, E, lane, zr);
  float b1 = FLT_MAX, b2 = FLT_MAX;                          // best and second-best fp32 distances
  int best = -1;
  for (int j0 = 0; j0 < n; j0 += 4) {
    float4 e[4][LT_EMAX / 128];
    int k[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      k[u] = mi_row[min(j0 + u, n - 1)];
      argmin_load_row(cb + (size_t)k[u] * E, E, lane, e[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (j0 + u < n) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < LT_EMAX / 128; ++q) {
          float dd;
          dd = zr[q].x - e[u][q].x; s = fmaf(dd, dd, s);
          dd = zr[q].y - e[u][q].y; s = fmaf(dd, dd, s);
          dd = zr[q].z - e[u][q].z; s = fmaf(dd, dd, s);
          dd = zr[q].w - e[u][q].w; s = fmaf(dd, dd, s);
        }
        s = warp_sum(s);
        if (s < b1 || (s == b1 && k[u] < best)) { b2 = b1; b1 = s; best = k[u]; }
        else if (s < b2) b2 = s;
      }
    }
  }
  // fp32 sums of non-negative terms: relative error <= (2 + 16 + 5) ulp ~ 1.4e-6 each; closer than that -> fp64
  if (!(b1 * (1.f + 4e-6f) < b2)) {
    double bd = 0.0;
    best = -1;
    for (int j = 0; j < n; ++j) {
      const int k = mi_row[j];
      float4 e[LT_EMAX / 128];
      argmin_load_row(cb + (size_t)k * E, E, lane, e);
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < LT_EMAX / 128; ++q) {
        double dd;
        dd = (double)zr[q].x - (double)e[q].x; s += dd * dd;
        dd = (double)zr[q].y - (double)e[q].y; s += dd * dd;
        dd = (double)zr[q].z - (double)e[q].z; s += dd * dd;
        dd = (double)zr[q].w - (double)e[q].w; s += dd * dd;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (best < 0 || s < bd || (s == bd && k < best)) { bd = s; best = k; }
    }
  }
  return best;
}

// ------------------------------------------------------------------------------ the pair sweep
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LT_THREADS, 1)
l2_argmin_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const float* __restrict__ z, const float2* __restrict__ zn2, int T, int E,
                      const float* __restrict__ cb, const float* __restrict__ norm, int K, int64_t* __restrict__ idx,
                      float* __restrict__ quant, int* __restrict__ fb_count, int* __restrict__ fb_list) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                          // [E/64][128 rows][128 B]: this CTA's 128 tokens
  uint8_t* sB = sA + LT_BM * LT_EMAX * 2;                      // [NST][128 rows][128 B]: this CTA's half of a codebook tile
  uint8_t* tail = sB + LT_NST * LT_BST;
  float2* lists = reinterpret_cast<float2*>(tail);             // [LT_LIST][2 groups][128 rows] (d~, code)
  float* xmin = reinterpret_cast<float*>(lists + LT_LIST * 2 * LT_BM);   // [2][128] running minima of the two warpgroups
  int* mi = reinterpret_cast<int*>(xmin + 2 * LT_BM);          // [128][LT_TOP] window members
  int* ncand = mi + LT_BM * LT_TOP;                            // [128] members found (zero between tiles)
  int* ovf = ncand + LT_BM;                                    // [128] a list overflowed
  uint64_t* b_full = reinterpret_cast<uint64_t*>(tail + LT_XCH);       // [NST]   (used in the leader)
  uint64_t* b_empty = b_full + LT_NST;                         // [NST]
  uint64_t* a_full = b_empty + LT_NST;                         // [8]     (leader)
  uint64_t* a_empty = a_full + LT_KBMAX;                       // [8]
  uint64_t* t_full = a_empty + LT_KBMAX;                       // [2]
  uint64_t* t_empty = t_full + 2;                              // [2]     (leader)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = E / LT_BK;                                    // k-blocks
  const int NT = K / LT_BN;                                    // N-tiles
  const int n_pt = (T + 2 * LT_BM - 1) / (2 * LT_BM);          // 256-token pair tiles
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < LT_NST; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < LT_KBMAX; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 8); }   // 4 scan warps in each CTA
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta<512>(tmem_ptr);
    tc_fence_before();
  }
  if (threadIdx.x < LT_BM) { ncand[threadIdx.x] = 0; ovf[threadIdx.x] = 0; }
  __syncthreads();
  cluster_sync_all();                                          // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0) {
    // ---------------------------------------------------------------- codebook producer: this CTA's half of every tile
    int st = 0;
    uint32_t ph = 0;
    for (int pt = pair; pt < n_pt; pt += npairs) {
      for (int nt = 0; nt < NT; ++nt) {
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&b_empty[st], ph ^ 1);
          if (elect_one()) {
            const uint32_t fb = mapa_u32(smem_u32(&b_full[st]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&b_full[st], 2 * LT_BST);
            tma_load_2d_2sm(sB + st * LT_BST, &tmB, fb, kb * LT_BK, nt * LT_BN + (int)rank * (LT_BN / 2));
          }
          __syncwarp();
          if (++st == LT_NST) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 10) {
    // ---------------------------------------------------------------- A producer: the bf16 z rows of this CTA's 128 tokens
    int it = 0;
    for (int pt = pair; pt < n_pt; pt += npairs, ++it) {
      const int t0 = pt * 2 * LT_BM + (int)rank * LT_BM;
      for (int kb = 0; kb < KB; ++kb) {
        if (it > 0) mbar_wait(&a_empty[kb], (it - 1) & 1);       // the previous tile's last N-tile has read this k-block
        if (elect_one()) {
          const uint32_t fa = mapa_u32(smem_u32(&a_full[kb]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&a_full[kb], 2 * LT_A_KB_BYTES);
          tma_load_2d_2sm(sA + kb * LT_A_KB_BYTES, &tmA, fa, kb * LT_BK, t0);       // rows >= T: zero fill
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (pair leader; one elected lane)
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * LT_BM, LT_BN);
      const uint64_t da0 = umma_desc_k_sw128(smem_u32(sA));
      const uint64_t db0 = umma_desc_k_sw128(smem_u32(sB));
      int st = 0, it = 0, gc = 0;                              // gc: N-tiles issued so far (accumulator buffer = gc & 1)
      uint32_t ph = 0;
      for (int pt = pair; pt < n_pt; pt += npairs, ++it) {
        for (int nt = 0; nt < NT; ++nt, ++gc) {
          const int buf = gc & 1, use = gc >> 1;
          if (use > 0) { mbar_wait(&t_empty[buf], (use - 1) & 1); tc_fence_after(); }
          for (int kb = 0; kb < KB; ++kb) {
            if (nt == 0) mbar_wait(&a_full[kb], it & 1);
            mbar_wait(&b_full[st], ph);
            tc_fence_after();
            const uint64_t da = da0 + (uint64_t)(kb * (LT_A_KB_BYTES >> 4));
            const uint64_t db = db0 + (uint64_t)(st * (LT_BST >> 4));
#pragma unroll
            for (int k = 0; k < LT_BK / 16; ++k)
              umma_bf16_ss_2cta(tmem_base + buf * LT_BN, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2cta(&b_empty[st]);                    // frees the stage in both CTAs
            if (nt == NT - 1) umma_commit_2cta(&a_empty[kb]);  // the tile is done with this A k-block
            if (kb == KB - 1) umma_commit_2cta(&t_full[buf]);
            if (++st == LT_NST) { st = 0; ph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- 8 worker warps: scan, merge, exact resolution
    const int w8 = warp - 2;                                   // 0..7
    const int g = w8 >> 2;                                     // warpgroup g scans the N-tiles with (global count & 1) == g
    const int quad = warp & 3;                                 // TMEM lane quadrant this warp may read
    const int r = quad * 32 + lane;
    const uint32_t tacc = tmem_base + (uint32_t(quad * 32) << 16) + g * LT_BN;
    const uint32_t te = mapa_u32(smem_u32(&t_empty[g]), 0);
    const float emax = sqrtf(__ldg(norm + K)), demax = sqrtf(__ldg(norm + K + 1));
    float2* lst = lists + g * LT_BM + r;
    int it = 0, use = 0;
    for (int pt = pair; pt < n_pt; pt += npairs, ++it) {
      const int t0 = pt * 2 * LT_BM + (int)rank * LT_BM;
      const bool live = t0 + r < T;
      float W;
      {
        const float2 nn = live ? __ldg(zn2 + t0 + r) : make_float2(0.f, 0.f);
        const float zl = sqrtf(nn.x) * 1.0001f, dzl = sqrtf(nn.y) * 1.0001f;
        const float D = 2.f * (dzl * emax + zl * demax) + zl * emax * (1.f / 4096.f) + 1e-30f;
        W = 2.f * D * 1.001f;
      }
      float runmin = live ? FLT_MAX : -FLT_MAX, thr = runmin;  // rows past T never pass the chunk test
      int cnt = 0;
      for (int nt = 0; nt < NT; ++nt) {
        if (((it * NT + nt) & 1) != g) continue;
        mbar_wait(&t_full[g], use & 1);
        ++use;
        tc_fence_after();
        argmin_scan_ntile(tacc, norm + nt * LT_BN, nt * LT_BN, W, smem_u32(lst), runmin, thr, cnt);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(te);                // one cluster-scope arrive per warp
      }
      // merge: every thread filters its own list against the row's minimum over both warpgroups
      xmin[g * LT_BM + r] = runmin;
      named_bar_sync(1, 256);
      {
        const float win = fminf(xmin[r], xmin[LT_BM + r]) + W;
        if (cnt > LT_LIST) ovf[r] = 1;
        const int nl = min(cnt, LT_LIST);
        for (int j = 0; j < nl; ++j) {
          const float2 e = lst[j * 2 * LT_BM];
          if (e.x <= win) {
            const int pos = atomicAdd(&ncand[r], 1);
            if (pos < LT_TOP) mi[r * LT_TOP + pos] = __float_as_int(e.y);
          }
        }
      }
      named_bar_sync(1, 256);
      // resolution: warp w8 owns rows w8 + 8 i; lane i < 16 looks at row i, windows of one code (the usual case) are done
      // there and then, the others go through the warp-wide exact evaluation one after the other
      {
        const int rr = w8 + 8 * (lane & 15);
        const int t = t0 + rr;
        const bool mine = lane < 16 && t < T;
        int n = 0, best = -1;
        if (mine) {
          n = ncand[rr];
          if (ovf[rr] != 0 || n > LT_TOP || n == 0) n = -1;
          if (n == 1) best = mi[rr * LT_TOP];
          if (n < 0) fb_list[atomicAdd(fb_count, 1)] = t;      // exhaustive kernel
        }
        __syncwarp();
        if (lane < 16) { ncand[rr] = 0; ovf[rr] = 0; }         // only this warp reads these rows: ready for the next tile
        unsigned multi = __ballot_sync(0xffffffffu, mine && n > 1);
        while (multi != 0) {
          const int l = __ffs(multi) - 1;
          multi &= multi - 1;
          const int n_l = __shfl_sync(0xffffffffu, n, l);
          const int rr_l = w8 + 8 * l;
          const int b = argmin_resolve_exact(z, cb, E, t0 + rr_l, n_l, mi + rr_l * LT_TOP, lane);
          if (lane == l) best = b;
        }
        if (mine && n > 0) idx[t] = best;
        if (quant != nullptr) {
          unsigned have = __ballot_sync(0xffffffffu, mine && n > 0);
          while (have != 0) {
            const int l = __ffs(have) - 1;
            have &= have - 1;
            const int b = __shfl_sync(0xffffffffu, best, l);
            const float4* src = reinterpret_cast<const float4*>(cb + (size_t)b * E);
            float4* dst = reinterpret_cast<float4*>(quant + (size_t)(t0 + w8 + 8 * l) * E);
            for (int e = lane; e < (E >> 2); e += 32) dst[e] = __ldg(src + e);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                          // neither CTA retires while the peer may still signal it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
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

// workspace layout (int32 units): [fb_count, pad x3][fb_list: T][zn2: T float2][zb: T x LT_EMAX bf16], every section
// 16-byte aligned
static inline int64_t ws_align4(int64_t v) { return (v + 3) & ~int64_t(3); }
static inline int64_t ws_off_list(int T) { (void)T; return 4; }
static inline int64_t ws_off_zn2(int T) { return ws_align4(ws_off_list(T) + T); }
static inline int64_t ws_off_zb(int T) { return ws_align4(ws_off_zn2(T) + (int64_t)T * 2); }

extern "C" int64_t pgt_l2_argmin_ws_ints(int T) {
  return ws_off_zb(T) + (int64_t)T * (LT_EMAX / 2);
}

extern "C" int pgt_l2_argmin_tc(const float* z, int T, int E, const float* codebook, const void* cb_bf16,
                                const float* cb_norm, int K, int64_t* idx, float* quant, int32_t* workspace,
                                void* stream) {
  PGT_CHECK_ARG(z && codebook && cb_bf16 && cb_norm && idx && workspace && T > 0);
  if (K % LT_BN != 0 || E % LT_BK != 0 || E > LT_EMAX || E % 128 != 0) return PGT_ERR_UNSUPPORTED;
  PGT_CHECK_ARG((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(codebook) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(cb_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0 &&
                (quant == nullptr || (reinterpret_cast<uintptr_t>(quant) & 15) == 0));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* fb_count = workspace;
  int* fb_list = workspace + ws_off_list(T);
  float2* zn2 = reinterpret_cast<float2*>(workspace + ws_off_zn2(T));
  __nv_bfloat16* zb = reinterpret_cast<__nv_bfloat16*>(workspace + ws_off_zb(T));
  CUtensorMap tmA, tmB;
  int rc = tmap_rows_bf16(&tmA, zb, E, T, E, LT_BM);
  if (rc == PGT_OK) rc = tmap_rows_bf16(&tmB, cb_bf16, E, K, E, LT_BN / 2);      // each CTA of a pair loads half a tile
  if (rc != PGT_OK) return rc;
  static PerDeviceOnce once;
  PGT_CUDA_OK(once.run([] { return cudaFuncSetAttribute(l2_argmin_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM); }));
  PGT_CUDA_OK(cudaMemsetAsync(workspace, 0, sizeof(int32_t), st));
  const int n_pt = ceil_div(T, 2 * LT_BM);
  const int npairs = n_pt < num_sms() / 2 ? n_pt : num_sms() / 2;
  {
    ProfScope ps(PGT_PROF_ARGMIN, 2.0 * T * (double)K * E, st, "l2_argmin_tc");
    // pack grid: warps take 4 rows per step; pick the CTA count in [SMs, 2 SMs] that leaves the smallest ragged last step
    int pack_ctas = num_sms() * 2;
    {
      const long long items = ceil_div(T, 4);
      long long best_waste = -1;
      for (int c = num_sms() * 2; c >= num_sms(); --c) {
        const long long w = (long long)c * 8;
        const long long waste = ((items + w - 1) / w) * w - items;
        if (best_waste < 0 || waste < best_waste) { best_waste = waste; pack_ctas = c; }
      }
    }
    z_pack_kernel<<<pack_ctas, 256, 0, st>>>(z, T, E, zb, zn2);
    l2_argmin_pair_kernel<<<2 * npairs, LT_THREADS, LT_SMEM, st>>>(tmA, tmB, z, zn2, T, E, codebook, cb_norm, K, idx, quant,
                                                                   fb_count, fb_list);
    PGT_LAUNCH_OK();
  }
  // tokens whose certificate window did not fit the shortlist (degenerate codebooks): exhaustive exact kernel
  return l2_argmin_list_launch(z, T, E, codebook, K, idx, quant, fb_list, fb_count, num_sms(), st);
}
