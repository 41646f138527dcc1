// RQ-VAE codebook kernels: row argmax + embedding gather (the path PGTFormer.forward takes) and the
// nearest-codebook L2 argmin (the path TDCRQVAE3.forward / get_codes takes).  Index results are exact:
// argmax is a pure fp32 compare; argmin is an fp32 shortlist (top-4 per token) re-evaluated in fp64, lowest
// index on ties — it matches an fp64 argmin of ||z-e||^2, which is stricter than the reference's own fp32 addmm.
#include <float.h>

#include "common.cuh"

namespace pgt {

// ------------------------------------------------------------------------------ argmax + gather
// One warp per token row: K fp32 logits streamed with 128-bit loads (lane-interleaved float4), running
// (max, first index), warp-shuffle reduction, then the codebook row is gathered cooperatively.
__global__ void __launch_bounds__(256)
argmax_gather_kernel(const float* __restrict__ logits, int T, int K, const float* __restrict__ codebook, int E,
                     const int64_t* __restrict__ idx_in, int64_t* __restrict__ idx_out, void* __restrict__ quant,
                     int ldq, int quant_dtype) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= T) return;
  int best_i;
  if (idx_in != nullptr) {
    best_i = (int)idx_in[row];
  } else {
    const float4* p = reinterpret_cast<const float4*>(logits + (size_t)row * K);
    float best = -FLT_MAX;
    best_i = 0x7fffffff;
    const int nvec = K >> 2;
    for (int i = lane; i < nvec; i += 32) {
      float4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
      const int b = i * 4;
      // strictly-greater keeps the first maximum inside the lane's ascending scan
      if (v.x > best) { best = v.x; best_i = b; }
      if (v.y > best) { best = v.y; best_i = b + 1; }
      if (v.z > best) { best = v.z; best_i = b + 2; }
      if (v.w > best) { best = v.w; best_i = b + 3; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
  }
  if (lane == 0 && idx_out != nullptr) idx_out[row] = best_i;
  if (quant != nullptr) {
    const float4* src = reinterpret_cast<const float4*>(codebook + (size_t)best_i * E);
    if (quant_dtype == PGT_BF16) {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(quant) + (size_t)row * ldq;
      for (int i = lane; i < (E >> 2); i += 32) {
        const float4 v = __ldg(src + i);
        uint2 u;
        u.x = pack_bf16x2(v.x, v.y);
        u.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(dst + i * 4) = u;
      }
    } else {
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(quant) + (size_t)row * ldq);
      for (int i = lane; i < (E >> 2); i += 32) dst[i] = __ldg(src + i);
    }
  }
}

// ------------------------------------------------------------------------------ L2 argmin
constexpr int AM_TT = 64;      // tokens per CTA
constexpr int AM_TC = 64;      // codes per tile
constexpr int AM_KC = 32;      // feature chunk
constexpr int AM_LD = AM_TT + 4;
constexpr int AM_TOP = 4;

struct Cand {
  float v;
  int i;
};
__device__ __forceinline__ bool cand_less(float v, int i, const Cand& c) { return v < c.v || (v == c.v && i < c.i); }
__device__ __forceinline__ void cand_insert(Cand (&top)[AM_TOP], float v, int i) {
  if (!cand_less(v, i, top[AM_TOP - 1])) return;
  top[AM_TOP - 1].v = v; top[AM_TOP - 1].i = i;
#pragma unroll
  for (int k = AM_TOP - 1; k > 0; --k) {
    if (cand_less(top[k].v, top[k].i, top[k - 1])) {
      const Cand t = top[k]; top[k] = top[k - 1]; top[k - 1] = t;
    }
  }
}

// Exhaustive and exact: every code is scored in fp32 (direct sum of squared differences), the per-token top-4 is
// re-evaluated in fp64.  Runs on all T tokens (list == nullptr), or on the tokens `list[0 .. *count)` that the
// tcgen05 kernel (l2_argmin_tc.cu) could not certify — normally none, then every CTA returns at once.
__global__ void __launch_bounds__(256)
l2_argmin_kernel(const float* __restrict__ z, int T, int E, const float* __restrict__ cb, int K,
                 int64_t* __restrict__ idx, float* __restrict__ quant, const int* __restrict__ list,
                 const int* __restrict__ count) {
  // staging tiles (17 KB) and the post-loop merge buffer (32 KB) share the same storage
  __shared__ __align__(16) unsigned char smraw[AM_TT * 16 * AM_TOP * sizeof(Cand)];
  __shared__ int short_list[AM_TT][AM_TOP];
  float (*xs)[AM_LD] = reinterpret_cast<float (*)[AM_LD]>(smraw);
  float (*es)[AM_LD] = reinterpret_cast<float (*)[AM_LD]>(smraw + AM_KC * AM_LD * sizeof(float));
  Cand (*merge)[16][AM_TOP] = reinterpret_cast<Cand (*)[16][AM_TOP]>(smraw);
  static_assert(2 * AM_KC * AM_LD * sizeof(float) <= sizeof(smraw), "staging tiles must fit the merge buffer");
  const int tx = threadIdx.x & 15;          // code micro-column
  const int ty = threadIdx.x >> 4;          // token micro-row
  const int n_tok = list != nullptr ? *count : T;
  for (int t0 = blockIdx.x * AM_TT; t0 < n_tok; t0 += gridDim.x * AM_TT) {
  auto token = [&](int i) { return list != nullptr ? list[i] : i; };
  __syncthreads();                          // the previous chunk's merge buffer / shortlist are no longer read
  Cand top[4][AM_TOP];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int k = 0; k < AM_TOP; ++k) { top[a][k].v = FLT_MAX; top[a][k].i = 0x7fffffff; }

  const int lrow = threadIdx.x >> 2;        // 0..63: token / code row this thread stages
  const int lcol = (threadIdx.x & 3) * 8;   // 8 consecutive features
  for (int c0 = 0; c0 < K; c0 += AM_TC) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int k0 = 0; k0 < E; k0 += AM_KC) {
      {
        const int ti = t0 + lrow;
        float4 a = make_float4(0, 0, 0, 0), b = a;
        if (ti < n_tok) {
          const float4* p = reinterpret_cast<const float4*>(z + (size_t)token(ti) * E + k0 + lcol);
          a = __ldg(p); b = __ldg(p + 1);
        }
        xs[lcol + 0][lrow] = a.x; xs[lcol + 1][lrow] = a.y; xs[lcol + 2][lrow] = a.z; xs[lcol + 3][lrow] = a.w;
        xs[lcol + 4][lrow] = b.x; xs[lcol + 5][lrow] = b.y; xs[lcol + 6][lrow] = b.z; xs[lcol + 7][lrow] = b.w;
        const int c = c0 + lrow;
        a = make_float4(0, 0, 0, 0); b = a;
        if (c < K) {
          const float4* p = reinterpret_cast<const float4*>(cb + (size_t)c * E + k0 + lcol);
          a = __ldg(p); b = __ldg(p + 1);
        }
        es[lcol + 0][lrow] = a.x; es[lcol + 1][lrow] = a.y; es[lcol + 2][lrow] = a.z; es[lcol + 3][lrow] = a.w;
        es[lcol + 4][lrow] = b.x; es[lcol + 5][lrow] = b.y; es[lcol + 6][lrow] = b.z; es[lcol + 7][lrow] = b.w;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < AM_KC; ++k) {
        const float4 xv = *reinterpret_cast<const float4*>(&xs[k][ty * 4]);
        const float4 ev = *reinterpret_cast<const float4*>(&es[k][tx * 4]);
        const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
        const float ea[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const float d = xa[a] - ea[b];
            acc[a][b] = fmaf(d, d, acc[a][b]);
          }
      }
      __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int c = c0 + tx * 4 + b;
        if (c < K) cand_insert(top[a], acc[a][b], c);
      }
  }
  // merge the 16 per-thread shortlists of each token
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int k = 0; k < AM_TOP; ++k) merge[ty * 4 + a][tx][k] = top[a][k];
  __syncthreads();
  if (threadIdx.x < AM_TT) {
    Cand best[AM_TOP];
#pragma unroll
    for (int k = 0; k < AM_TOP; ++k) { best[k].v = FLT_MAX; best[k].i = 0x7fffffff; }
    for (int j = 0; j < 16; ++j)
#pragma unroll
      for (int k = 0; k < AM_TOP; ++k) cand_insert(best, merge[threadIdx.x][j][k].v, merge[threadIdx.x][j][k].i);
#pragma unroll
    for (int k = 0; k < AM_TOP; ++k) short_list[threadIdx.x][k] = best[k].i;
  }
  __syncthreads();
  // fp64 re-evaluation of the shortlisted codes: one warp per token
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int tt = warp; tt < AM_TT; tt += 8) {
    if (t0 + tt >= n_tok) continue;
    const int t = token(t0 + tt);
    double bd = 0.0;
    int bi = -1;
    for (int k = 0; k < AM_TOP; ++k) {
      const int c = short_list[tt][k];
      if (c < 0 || c >= K) continue;
      double s = 0.0;
      for (int e = lane; e < E; e += 32) {
        const double d = (double)z[(size_t)t * E + e] - (double)cb[(size_t)c * E + e];
        s += d * d;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (bi < 0 || s < bd || (s == bd && c < bi)) { bd = s; bi = c; }
    }
    if (lane == 0) idx[t] = bi;
    if (quant != nullptr)
      for (int e = lane; e < E; e += 32) quant[(size_t)t * E + e] = cb[(size_t)bi * E + e];
  }
  }
}

int l2_argmin_list_launch(const float* z, int T, int E, const float* codebook, int K, int64_t* idx, float* quant,
                          const int* list, const int* count, int grid, cudaStream_t st) {
  if (E % AM_KC != 0) return PGT_ERR_UNSUPPORTED;
  l2_argmin_kernel<<<grid, 256, 0, st>>>(z, T, E, codebook, K, idx, quant, list, count);
  PGT_LAUNCH_OK();
  return PGT_OK;
}

}  // namespace pgt

using namespace pgt;

extern "C" int pgt_argmax_gather(const float* logits, int T, int K, const float* codebook, int E,
                                 const int64_t* idx_in, int64_t* idx, void* quant, int ldq, int quant_dtype,
                                 void* stream) {
  PGT_CHECK_ARG((logits || idx_in) && T > 0 && K > 0 && K % 4 == 0);
  PGT_CHECK_ARG(quant == nullptr || (codebook != nullptr && E % 4 == 0 && ldq % 4 == 0));
  PGT_CHECK_ARG((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  const int warps = 8;
  ProfScope ps(PGT_PROF_ARGMAX, (double)T * K * 4 + (double)T * 8 + (quant ? (double)T * E * (quant_dtype == PGT_BF16 ? 2 : 4) : 0.0),
               static_cast<cudaStream_t>(stream));
  argmax_gather_kernel<<<ceil_div(T, warps), warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, T, K, codebook, E, idx_in, idx, quant, ldq, quant_dtype);
  PGT_LAUNCH_OK();
  return PGT_OK;
}

extern "C" int pgt_l2_argmin(const float* z, int T, int E, const float* codebook, int K, int64_t* idx, float* quant,
                             void* stream) {
  PGT_CHECK_ARG(z && codebook && idx && T > 0 && K > 0 && E > 0 && E % AM_KC == 0);
  PGT_CHECK_ARG((reinterpret_cast<uintptr_t>(z) & 15) == 0 && (reinterpret_cast<uintptr_t>(codebook) & 15) == 0);
  ProfScope ps(PGT_PROF_ARGMIN, 2.0 * T * (double)K * E, static_cast<cudaStream_t>(stream));
  l2_argmin_kernel<<<ceil_div(T, AM_TT), 256, 0, static_cast<cudaStream_t>(stream)>>>(z, T, E, codebook, K, idx, quant, nullptr, nullptr);
  PGT_LAUNCH_OK();
  return PGT_OK;
}
