// Shifted-window spatio-temporal attention core (WindowAttention3D, modules/rstt_layers.py:195-234, with the roll /
// window_partition / window_reverse / shift mask of VSTSREncoderTransformerBlock :301-329 and EncoderLayer :552-568)
// on TMA + tcgen05, sm_100a.
//
// Work unit: a PAIR of 3x4x4 windows (2 x 48 tokens) and one 64-column chunk of the heads (2 heads of d = 32, or one of
// d = 64).  Persistent CTAs, warp-specialised:
//   warp 0     TMA producer.  A window's q / k / v rows of one chunk are ONE 5-D box [64 ch, 4 x, 4 y, 3 frames, clip] of
//              the [T, 3C] qkv matrix (128B swizzle) — the roll by -shift is a coordinate offset, window_partition is the
//              box shape.  Windows in the last window row / column of a shifted block wrap around the frame: they are 2
//              (or 4) half (quarter) boxes, landing one after the other, so their rows sit in a permuted order; attention
//              does not care as long as bias and mask are permuted alike, which the host does once per layer (below).
//   warp 1     tcgen05.mma issuer.  Per head and pair:  S = Q K^T as two 128x48xd MMAs (window 0 from tile row 0, window
//              1 from tile row 48-64, so that its rows land on TMEM lanes 64..111 and every warp of the softmax group sees
//              ONE window), then O = P V as two 128 x d x 48 MMAs (V consumed MN-major straight from its TMA tile).
//   warps 2-5 / 6-9   two softmax groups that alternate chunks (ping-pong against the tensor pipe): thread = query row,
//              tcgen05.ld of its 48 scores, t = s * scale*log2e + table, exp2, row sum, bf16 P into the swizzled A tile,
//              then O * (1/sum) from TMEM -> bf16 -> swizzled staging tile -> TMA store (window_reverse + roll back are
//              the same box coordinates as the load).
// Bias / mask table (built at load time by the engine, fp16, already multiplied by log2 e):
//   tab[type][head][j = key/8][row][8]   type 0 interior, 1 x-wrapped (right edge), 2 y-wrapped (bottom edge), 3 corner;
//   entry = relative_position_bias[pi_t(row)][pi_t(key)] + (-100 where the reference's shift mask separates the two
//   tokens), pi_t = the row permutation of the wrapped box order.  Type 0 lives in shared memory, the others in L2.
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace pgt {

constexpr int WT_N = 48;                          // tokens per window
constexpr int WT_TILE = 2 * WT_N * 128;           // 12 KB: one operand chunk of a window pair (96 rows x 128 B)
constexpr int WT_STAGE = 3 * WT_TILE;             // k | v | q
constexpr int WT_NST = 3;
constexpr int WT_PTILE = 128 * 128;               // 16 KB: P tile in TMEM-lane space
constexpr int WT_THREADS = 64 + 256;
constexpr int WT_HEADS = 8;
constexpr int WT_TAB_BYTES = WT_HEADS * 6 * WT_N * 16;            // 36 KB: one type of the fp16 table
constexpr int WT_SMEM = 2048 /*pad read by the row -16 view*/ + WT_NST * WT_STAGE + 2 * WT_PTILE + 2 * WT_TILE + WT_TAB_BYTES +
                        512 /*barriers*/ + 1024 /*align*/;

struct WinParams {
  int clips, H, W, C, heads, d, shift;
  int nwx, nwy, n_windows, n_pairs, n_chunks, hpc;   // hpc: heads per 64-column chunk
  int mode_n64;                                      // d = 32: compute P V with N = 64 (both heads' columns) instead of a half-atom N = 32 view
  float sl2;                                         // d^-1/2 * log2(e)
  const uint4* tab;                                  // [4][heads][6][48] x 16 B
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

// MN-major SW128 operand (V: key rows x d contiguous): 8 key rows per 1024-byte atom.
__device__ __forceinline__ uint64_t wt_desc_mn_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((16384 >> 4) & 0x3FFF) << 16;    // LBO: next 64-column atom (never used: N <= 64)
  d |= static_cast<uint64_t>((1024 >> 4) & 0x3FFF) << 32;     // SBO: next 8 key rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

struct WinCoord {
  int clip, x0, y0, xs, ys;        // xs / ys: 1 when the window wraps in x / y
};

__device__ __forceinline__ WinCoord win_coord(const WinParams& p, int w) {
  WinCoord c;
  const int per_clip = p.nwx * p.nwy;
  c.clip = w / per_clip;
  const int r = w - c.clip * per_clip;
  const int wy = r / p.nwx, wx = r - wy * p.nwx;
  c.xs = (p.shift > 0 && wx == p.nwx - 1) ? 1 : 0;
  c.ys = (p.shift > 0 && wy == p.nwy - 1) ? 1 : 0;
  c.x0 = wx * 4 + p.shift;
  c.y0 = wy * 4 + p.shift;
  return c;
}

__global__ void __launch_bounds__(WT_THREADS, 1)
window_attn_tc_kernel(const __grid_constant__ CUtensorMap tmI0, const __grid_constant__ CUtensorMap tmI1,
                      const __grid_constant__ CUtensorMap tmI2, const __grid_constant__ CUtensorMap tmI3,
                      const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                      const __grid_constant__ CUtensorMap tmO2, const __grid_constant__ CUtensorMap tmO3,
                      const WinParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem + 2048;                                  // [NST][k | v | q]
  uint8_t* sP = ring + WT_NST * WT_STAGE;                       // [2 groups][128 rows x 128 B]
  uint8_t* sO = sP + 2 * WT_PTILE;                              // [2 groups][96 rows x 128 B]
  uint8_t* sTab = sO + 2 * WT_TILE;                             // type-0 table
  uint64_t* bars = reinterpret_cast<uint64_t*>(sTab + WT_TAB_BYTES);
  uint64_t* st_full = bars;                                     // [NST]
  uint64_t* st_empty = st_full + WT_NST;                        // [NST]
  uint64_t* s_full = st_empty + WT_NST;                         // [2]
  uint64_t* p_full = s_full + 2;                                // [2]
  uint64_t* o_full = p_full + 2;                                // [2]
  uint64_t* o_empty = o_full + 2;                               // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const CUtensorMap* tmI[4] = {&tmI0, &tmI1, &tmI2, &tmI3};
  const CUtensorMap* tmO[4] = {&tmO0, &tmO1, &tmO2, &tmO3};

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) { tma_prefetch_desc(tmI[i]); tma_prefetch_desc(tmO[i]); }
    for (int i = 0; i < WT_NST; ++i) { mbar_init(&st_full[i], 1); mbar_init(&st_empty[i], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 128);
      mbar_init(&o_full[g], 1);
      mbar_init(&o_empty[g], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<512>(tmem_ptr);
    tc_fence_before();
  }
  // type-0 table -> shared memory (all threads)
  for (int i = threadIdx.x; i < WT_TAB_BYTES / 16; i += WT_THREADS) reinterpret_cast<uint4*>(sTab)[i] = __ldg(p.tab + i);
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int my_pairs = (p.n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // pairs blockIdx.x, +gridDim.x, ..
  const int NCH = p.n_chunks;
  const int HPC = p.hpc;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int st = 0;
    uint32_t ph = 0;
    for (int i = 0; i < my_pairs; ++i) {
      const int pair = blockIdx.x + i * gridDim.x;
      for (int c = 0; c < NCH; ++c) {
        mbar_wait(&st_empty[st], ph ^ 1);
        if (elect_one()) {
          uint8_t* sK = ring + st * WT_STAGE;
          uint8_t* sV = sK + WT_TILE;
          uint8_t* sQ = sV + WT_TILE;
          const int nwin = (2 * pair + 1 < p.n_windows) ? 2 : 1;
          mbar_arrive_expect_tx(&st_full[st], nwin * 3 * WT_N * 128);
          for (int wi = 0; wi < nwin; ++wi) {
            const WinCoord wc = win_coord(p, 2 * pair + wi);
            const int nx = wc.xs ? 2 : 1, ny = wc.ys ? 2 : 1;
            const CUtensorMap* m = tmI[wc.ys * 2 + wc.xs];
            const int part_bytes = (WT_N / (nx * ny)) * 128;
            int off = wi * WT_N * 128;
            for (int py = 0; py < ny; ++py) {
              const int y = wc.ys ? (py == 0 ? p.H - 2 : 0) : wc.y0;
              for (int px = 0; px < nx; ++px) {
                const int x = wc.xs ? (px == 0 ? p.W - 2 : 0) : wc.x0;
                tma_load_5d(sQ + off, m, &st_full[st], c * 64, x, y, 0, wc.clip);
                tma_load_5d(sK + off, m, &st_full[st], p.C + c * 64, x, y, 0, wc.clip);
                tma_load_5d(sV + off, m, &st_full[st], 2 * p.C + c * 64, x, y, 0, wc.clip);
                off += part_bytes;
              }
            }
          }
        }
        __syncwarp();
        if (++st == WT_NST) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // Group g works on chunks c = g (mod 2) of every pair; its items are (pair, chunk, head in chunk).
    const int items = my_pairs * (NCH / 2) * HPC;               // per group
    const int d = p.d;
    const uint32_t idesc_s = umma_idesc_bf16(128, WT_N);
    const int n_o = (d == 32 && !p.mode_n64) ? 32 : 64;
    const uint32_t idesc_o = umma_idesc_bf16(128, n_o) | (1u << 16);
    auto stage_of = [&](int g, int n, int& st, uint32_t& ph) {    // global chunk index of item n of group g
      const int q = (n / HPC) * 2 + g;
      st = q % WT_NST; ph = (q / WT_NST) & 1;
    };
    auto issue_s = [&](int g, int n) {
      int st; uint32_t ph;
      stage_of(g, n, st, ph);
      const int hh = n % HPC;
      if (hh == 0) { mbar_wait(&st_full[st], ph); tc_fence_after(); }
      if (elect_one()) {
        uint8_t* sK = ring + st * WT_STAGE;
        uint8_t* sQ = sK + 2 * WT_TILE;
        const uint32_t tS = tmem_base + g * 256;
        const uint32_t koff = hh * 64;                           // second head of a d = 32 chunk: +64 B inside the row
#pragma unroll
        for (int wi = 0; wi < 2; ++wi) {
          // window 1 through the view that starts 16 rows before the tile: its rows land on lanes 64..111
          const uint64_t da = umma_desc_k_sw128(smem_u32(sQ) + koff - (wi ? 16 * 128 : 0));
          const uint64_t db = umma_desc_k_sw128(smem_u32(sK) + koff + wi * WT_N * 128);
          for (int k = 0; k < d / 16; ++k) umma_bf16_ss(tS + wi * WT_N, da + 2 * k, db + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[g]);
      }
      __syncwarp();
    };
    auto issue_o = [&](int g, int n) {
      int st; uint32_t ph;
      stage_of(g, n, st, ph);
      const int hh = n % HPC;
      if (elect_one()) {
        uint8_t* sV = ring + st * WT_STAGE + WT_TILE;
        const uint32_t tO = tmem_base + g * 256 + 96;
        const uint32_t voff = (d == 32 && !p.mode_n64) ? hh * 64 : 0;
        const uint64_t da = umma_desc_k_sw128(smem_u32(sP + g * WT_PTILE));
#pragma unroll
        for (int wi = 0; wi < 2; ++wi) {
#pragma unroll
          for (int k = 0; k < WT_N / 16; ++k) {
            const uint64_t db = wt_desc_mn_sw128(smem_u32(sV) + voff + (wi * WT_N + k * 16) * 128);
            umma_bf16_ss(tO + wi * 64, da + 2 * k, db, idesc_o, k != 0 ? 1u : 0u);
          }
        }
        umma_commit(&o_full[g]);
        if (hh == HPC - 1) umma_commit(&st_empty[st]);           // every MMA reading this stage has been issued
      }
      __syncwarp();
    };
    if (items > 0) {
      issue_s(0, 0);
      issue_s(1, 0);
      for (int n = 0; n < items; ++n) {
        const uint32_t par = n & 1;
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&p_full[g], par);                           // P_g(n) is in smem and S_g has been read out
          mbar_wait(&o_empty[g], par ^ 1);                      // O_g of item n-1 has been drained
          tc_fence_after();
          issue_o(g, n);
          if (n + 1 < items) issue_s(g, n + 1);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / output groups
    const int g = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int L = quad * 32 + lane;                              // TMEM lane = row of the P tile
    const int wi = quad >> 1;                                    // window of the pair this warp sees
    const int rw_raw = L - wi * 64;                              // row inside the window (valid < 48)
    const bool valid_row = rw_raw < WT_N;
    const int rw = valid_row ? rw_raw : WT_N - 1;
    const int tr = wi * WT_N + rw;                               // row in token space (staging tile)
    const uint32_t lane_base = uint32_t(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + g * 256 + wi * WT_N;
    const uint32_t tO = tmem_base + lane_base + g * 256 + 96 + wi * 64;
    uint8_t* prow = sP + g * WT_PTILE + L * 128;
    uint8_t* orow = sO + g * WT_TILE + tr * 128;
    const int d = p.d;
    const int items = my_pairs * (NCH / 2) * HPC;
    const bool leader = (warp - 2) % 4 == 0 && lane == 0;        // issues this group's TMA stores
    for (int n = 0; n < items; ++n) {
      const uint32_t par = n & 1;
      const int hh = n % HPC;
      const int cidx = n / HPC;                                  // this group's chunk counter
      const int pair = blockIdx.x + (cidx / (NCH / 2)) * gridDim.x;
      const int chunk = (cidx % (NCH / 2)) * 2 + g;
      const int head = chunk * HPC + hh;
      const int w = 2 * pair + wi;
      const bool win_ok = w < p.n_windows;
      int type = 0;
      if (win_ok) { const WinCoord wc = win_coord(p, w); type = wc.ys * 2 + wc.xs; }
      // bias / mask row of this thread: type 0 from shared memory, wrapped types from L2 (generic pointer)
      const uint4* trow = (type == 0 ? reinterpret_cast<const uint4*>(sTab) : p.tab + (size_t)type * (WT_TAB_BYTES / 16)) +
                          (size_t)head * 6 * WT_N + rw;
      mbar_wait(&s_full[g], par);
      tc_fence_after();
      uint32_t s0[32], s1[16];
      tmem_ld_32x32(tS, s0);
      tmem_ld_32x16(tS + 32, s1);
      float t[WT_N];
      uint4 bq[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) bq[j] = trow[j * WT_N];
      tmem_ld_wait();
      float mx = -1e30f;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const uint32_t u[4] = {bq[j].x, bq[j].y, bq[j].z, bq[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&u[e]));
          const int c = j * 8 + 2 * e;
          const float sa = __uint_as_float(c < 32 ? s0[c] : s1[c - 32]);
          const float sb = __uint_as_float(c + 1 < 32 ? s0[c + 1] : s1[c + 1 - 32]);
          t[c] = fmaf(sa, p.sl2, b2.x);
          t[c + 1] = fmaf(sb, p.sl2, b2.y);
          mx = fmaxf(mx, fmaxf(t[c], t[c + 1]));
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < WT_N; ++c) { t[c] = ex2_approx(t[c] - mx); sum += t[c]; }
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        uint4 u;
        u.x = pack_bf16x2(t[8 * j + 0], t[8 * j + 1]);
        u.y = pack_bf16x2(t[8 * j + 2], t[8 * j + 3]);
        u.z = pack_bf16x2(t[8 * j + 4], t[8 * j + 5]);
        u.w = pack_bf16x2(t[8 * j + 6], t[8 * j + 7]);
        *reinterpret_cast<uint4*>(prow + ((j ^ (L & 7)) << 4)) = u;
      }
      const float inv = 1.f / sum;
      tc_fence_before();
      fence_proxy_async();                                       // P (generic writes) -> visible to the tensor core
      mbar_arrive(&p_full[g]);
      // the staging tile of the previous chunk must have been read by its TMA store before anyone overwrites it
      if (hh == 0 && n > 0) {
        if (leader) bulk_wait_read<0>();
        named_bar_sync(2 + g, 128);
      }
      // O = (P V) / sum
      mbar_wait(&o_full[g], par);
      tc_fence_after();
      const int ocol = (d == 32 && p.mode_n64) ? hh * 32 : 0;
      if (d == 32) {
        uint32_t v[32];
        tmem_ld_32x32(tO + ocol, v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&o_empty[g]);
        if (valid_row) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(v[8 * q + 0]) * inv, __uint_as_float(v[8 * q + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv, __uint_as_float(v[8 * q + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv, __uint_as_float(v[8 * q + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv, __uint_as_float(v[8 * q + 7]) * inv);
            *reinterpret_cast<uint4*>(orow + (((hh * 4 + q) ^ (tr & 7)) << 4)) = u;
          }
        }
      } else {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          tmem_ld_32x32(tO + half * 32, v);
          tmem_ld_wait();
          if (half == 1) { tc_fence_before(); mbar_arrive(&o_empty[g]); }
          if (valid_row) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(v[8 * q + 0]) * inv, __uint_as_float(v[8 * q + 1]) * inv);
              u.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv, __uint_as_float(v[8 * q + 3]) * inv);
              u.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv, __uint_as_float(v[8 * q + 5]) * inv);
              u.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv, __uint_as_float(v[8 * q + 7]) * inv);
              *reinterpret_cast<uint4*>(orow + (((half * 4 + q) ^ (tr & 7)) << 4)) = u;
            }
          }
        }
      }
      if (hh == HPC - 1) {
        // chunk complete: window_reverse + roll back = the same boxes as the load, as TMA stores
        fence_proxy_async();
        named_bar_sync(2 + g, 128);
        if (leader) {
          const int nwin = (2 * pair + 1 < p.n_windows) ? 2 : 1;
          for (int w2 = 0; w2 < nwin; ++w2) {
            const WinCoord wc = win_coord(p, 2 * pair + w2);
            const int nx = wc.xs ? 2 : 1, ny = wc.ys ? 2 : 1;
            const CUtensorMap* m = tmO[wc.ys * 2 + wc.xs];
            const int part_bytes = (WT_N / (nx * ny)) * 128;
            int off = w2 * WT_N * 128;
            for (int py = 0; py < ny; ++py) {
              const int y = wc.ys ? (py == 0 ? p.H - 2 : 0) : wc.y0;
              for (int px = 0; px < nx; ++px) {
                const int x = wc.xs ? (px == 0 ? p.W - 2 : 0) : wc.x0;
                tma_store_5d(m, sO + g * WT_TILE + off, chunk * 64, x, y, 0, wc.clip);
                off += part_bytes;
              }
            }
          }
          bulk_commit();
        }
      }
    }
    if (leader) bulk_wait0();                                    // stores complete before the CTA retires its smem
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace pgt

using namespace pgt;

// 5-D view [ch, x, y, frame, clip] of a [T, ld] token matrix whose rows are ordered (clip, frame, y, x).
static int win_map(CUtensorMap* map, const void* base, int ld, int cols, int clips, int H, int W, int bx, int by) {
  const uint64_t dims[5] = {(uint64_t)cols, (uint64_t)W, (uint64_t)H, 3, (uint64_t)clips};
  const uint64_t row = (uint64_t)ld * 2;
  const uint64_t strides[4] = {row, row * W, row * W * H, row * W * H * 3};
  const uint32_t box[5] = {64, (uint32_t)bx, (uint32_t)by, 3, 1};
  return tmap_encode(map, base, 5, dims, strides, box);
}

extern "C" int pgt_window_attention_tc(const void* qkv, int ldqkv, int clips, int H, int W, int C, int heads, int shift,
                                       const void* tab, void* out, int ldo, int mode_n64, void* stream) {
  PGT_CHECK_ARG(qkv && tab && out && clips > 0 && H > 0 && W > 0 && heads > 0);
  PGT_CHECK_ARG(H % 4 == 0 && W % 4 == 0 && C % heads == 0 && ldqkv % 8 == 0 && ldo % 8 == 0 && ldqkv >= 3 * C);
  if (H <= 4 || W <= 4) shift = 0;                         // get_window_size(): no shift when the map is one window
  const int d = C / heads;
  if ((d != 32 && d != 64) || heads != WT_HEADS || C % 128 != 0 || (shift != 0 && shift != 2)) return PGT_ERR_UNSUPPORTED;
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al(qkv) || !al(out) || !al(tab)) return PGT_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap mi[4], mo[4];
  for (int t = 0; t < 4; ++t) {
    const int bx = (t & 1) ? 2 : 4, by = (t & 2) ? 2 : 4;
    int rc = win_map(&mi[t], qkv, ldqkv, 3 * C, clips, H, W, bx, by);
    if (rc == PGT_OK) rc = win_map(&mo[t], out, ldo, C, clips, H, W, bx, by);
    if (rc != PGT_OK) return rc;
  }
  WinParams p{};
  p.clips = clips; p.H = H; p.W = W; p.C = C; p.heads = heads; p.d = d; p.shift = shift;
  p.nwx = W / 4; p.nwy = H / 4;
  p.n_windows = clips * p.nwx * p.nwy;
  p.n_pairs = (p.n_windows + 1) / 2;
  p.n_chunks = C / 64;
  p.hpc = 64 / d;
  p.mode_n64 = mode_n64 ? 1 : 0;
  p.sl2 = (1.0f / sqrtf((float)d)) * 1.4426950408889634f;
  p.tab = reinterpret_cast<const uint4*>(tab);
  static PerDeviceOnce once;
  PGT_CUDA_OK(once.run([] { return cudaFuncSetAttribute(window_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM); }));
  const int grid = p.n_pairs < num_sms() ? p.n_pairs : num_sms();
  ProfScope ps(PGT_PROF_WINDOW_ATTN, 4.0 * WT_N * WT_N * C * (double)p.n_windows, st, "window_attn_tc");
  window_attn_tc_kernel<<<grid, WT_THREADS, WT_SMEM, st>>>(mi[0], mi[1], mi[2], mi[3], mo[0], mo[1], mo[2], mo[3], p);
  PGT_LAUNCH_OK();
  return PGT_OK;
}
