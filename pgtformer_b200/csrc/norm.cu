// GroupNorm(32)+SiLU, LayerNorm and AdaIN on channels-last bf16 activations (HBM-bound kernels:
// 128-bit vector loads, fp32 statistics, deterministic two-level reductions — no atomics).
#include "common.cuh"

namespace pgt {

constexpr int GN_GROUPS = 32;
constexpr int GN_MAX_CHUNKS = 64;
constexpr int GN_THREADS = 512;

__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ---------------------------------------------------------------------------------- GroupNorm
// Pass 1: per (frame, pixel-chunk) partial (sum, sumsq) of every group.  Thread (prow, vcol) owns 8 fixed
// channels and strides over the chunk's pixels; channel sums are folded to groups at the end.
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int HW, int C, int pix_per_chunk,
                float* __restrict__ partial /*[F][chunks][32][2]*/) {
  extern __shared__ float sm[];              // [2][C] channel sums, then reused
  const int vc = C >> 3;
  const int rows_par = GN_THREADS / vc;
  const int vcol = threadIdx.x % vc;
  const int prow = threadIdx.x / vc;
  const int chunk = blockIdx.x, f = blockIdx.y, nchunks = gridDim.x;
  const int p0 = chunk * pix_per_chunk;
  const int p1 = min(HW, p0 + pix_per_chunk);
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  if (prow < rows_par) {
    const __nv_bfloat16* base = x + ((size_t)f * HW) * ldx + vcol * 8;
    for (int p = p0 + prow; p < p1; p += rows_par) {
      float v[8];
      load8_bf16(base + (size_t)p * ldx, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += v[i]; q[i] += v[i] * v[i]; }
    }
  }
  for (int i = threadIdx.x; i < 2 * C; i += GN_THREADS) sm[i] = 0.f;
  __syncthreads();
  // serialised accumulation over prow keeps the order fixed (deterministic)
  for (int rr = 0; rr < rows_par; ++rr) {
    if (prow == rr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sm[vcol * 8 + i] += s[i];
        sm[C + vcol * 8 + i] += q[i];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < GN_GROUPS) {
    const int cpg = C / GN_GROUPS;
    float gs = 0.f, gq = 0.f;
    for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) { gs += sm[c]; gq += sm[C + c]; }
    float* o = partial + (((size_t)f * nchunks + chunk) * GN_GROUPS + threadIdx.x) * 2;
    o[0] = gs; o[1] = gq;
  }
}

// Pass 2: y = act(x * a[c] + b[c]) with a = rstd*gamma, b = beta - mean*rstd*gamma.
__global__ void __launch_bounds__(256)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int HW, int C, int pix_per_block, int nchunks,
                const float* __restrict__ partial, const float* __restrict__ gamma, const float* __restrict__ beta,
                float eps, int apply_silu, __nv_bfloat16* __restrict__ y, int ldy) {
  extern __shared__ float sm[];              // a[C], b[C], mean[32], rstd[32]
  float* sa = sm;
  float* sb = sm + C;
  float* smean = sm + 2 * C;
  float* srstd = smean + GN_GROUPS;
  const int f = blockIdx.y;
  const int cpg = C / GN_GROUPS;
  if (threadIdx.x < GN_GROUPS) {
    double s = 0.0, q = 0.0;
    for (int ch = 0; ch < nchunks; ++ch) {
      const float* o = partial + (((size_t)f * nchunks + ch) * GN_GROUPS + threadIdx.x) * 2;
      s += (double)o[0]; q += (double)o[1];
    }
    const double n = (double)HW * cpg;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    smean[threadIdx.x] = (float)mean;
    srstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float a = srstd[g] * gamma[c];
    sa[c] = a;
    sb[c] = beta[c] - smean[g] * a;
  }
  __syncthreads();
  const int vc = C >> 3;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const size_t total = (size_t)(p1 - p0) * vc;
  for (size_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int p = p0 + (int)(i / vc);
    const int c0 = (int)(i % vc) * 8;
    float v[8];
    load8_bf16(x + ((size_t)f * HW + p) * ldx + c0, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = v[j] * sa[c0 + j] + sb[c0 + j];
      v[j] = apply_silu ? t / (1.0f + __expf(-t)) : t;
    }
    store8_bf16(y + ((size_t)f * HW + p) * ldy + c0, v);
  }
}

// ---------------------------------------------------------------------------------- LayerNorm
// One warp per token row; two-pass (mean, then centred variance) entirely in registers.
template <int MAXV>   // max 8-element vectors per lane: C <= 256 * MAXV
__global__ void __launch_bounds__(256)
layernorm_kernel(const void* __restrict__ xin, int ldx, int x_dtype, int T, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y, int ldy,
                 const __nv_bfloat16* __restrict__ pos, int ldpos, __nv_bfloat16* __restrict__ y2, int ldy2) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= T) return;
  const int nv = C >> 8;                       // vectors of 8 per lane (C multiple of 256)
  float v[MAXV][8];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv) {
      const int c0 = (i * 32 + lane) * 8;
      if (x_dtype == PGT_BF16) {
        load8_bf16(reinterpret_cast<const __nv_bfloat16*>(xin) + (size_t)warp * ldx + c0, v[i]);
      } else {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(xin) + (size_t)warp * ldx + c0);
        const float4 a = __ldg(p), b = __ldg(p + 1);
        v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
        v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (i < nv)
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (i < nv) {
      const int c0 = (i * 32 + lane) * 8;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * __ldg(gamma + c0 + j) + __ldg(beta + c0 + j);
      store8_bf16(y + (size_t)warp * ldy + c0, o);
      if (y2 != nullptr) {
        float pv[8];
        load8_bf16(pos + (size_t)warp * ldpos + c0, pv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += pv[j];
        store8_bf16(y2 + (size_t)warp * ldy2 + c0, o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------- AdaIN
// Block = (frame, 64-channel slab).  Thread (prow, vcol) strides over pixels accumulating fp32 (sum, sumsq)
// for content q and style l; block reduce; second sweep applies the affine map.
constexpr int ADAIN_THREADS = 256;
__global__ void __launch_bounds__(ADAIN_THREADS)
adain_kernel(const void* __restrict__ qin, int ldq, int q_dtype, const __nv_bfloat16* __restrict__ l, int ldl, int HW,
             int C, float eps, __nv_bfloat16* __restrict__ y, int ldy) {
  __shared__ float red[4][ADAIN_THREADS / 8][64];   // [stat][prow][channel]
  __shared__ float sa[64], sb[64];
  const int f = blockIdx.y;
  const int cbase = blockIdx.x * 64;
  const int vcol = threadIdx.x & 7;
  const int prow = threadIdx.x >> 3;
  const int rows_par = ADAIN_THREADS / 8;
  const int c0 = cbase + vcol * 8;
  float acc[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  auto load_q = [&](int p, float (&v)[8]) {
    if (q_dtype == PGT_BF16) {
      load8_bf16(reinterpret_cast<const __nv_bfloat16*>(qin) + ((size_t)f * HW + p) * ldq + c0, v);
    } else {
      const float4* pp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(qin) + ((size_t)f * HW + p) * ldq + c0);
      const float4 a = __ldg(pp), b = __ldg(pp + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
  };
  for (int p = prow; p < HW; p += rows_par) {
    float vq[8], vl[8];
    load_q(p, vq);
    load8_bf16(l + ((size_t)f * HW + p) * ldl + c0, vl);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[0][j] += vq[j]; acc[1][j] += vq[j] * vq[j];
      acc[2][j] += vl[j]; acc[3][j] += vl[j] * vl[j];
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[k][prow][vcol * 8 + j] = acc[k][j];
  __syncthreads();
  if (threadIdx.x < 64) {
    double s[4] = {0, 0, 0, 0};
    for (int r = 0; r < rows_par; ++r)
#pragma unroll
      for (int k = 0; k < 4; ++k) s[k] += (double)red[k][r][threadIdx.x];
    const double n = (double)HW;
    const double mq = s[0] / n, ml = s[2] / n;
    double vq = (s[1] - n * mq * mq) / (n - 1.0), vl = (s[3] - n * ml * ml) / (n - 1.0);   // unbiased (torch.var)
    if (vq < 0) vq = 0;
    if (vl < 0) vl = 0;
    const double sq = sqrt(vq + (double)eps), sl = sqrt(vl + (double)eps);
    const double a = sl / sq;
    sa[threadIdx.x] = (float)a;
    sb[threadIdx.x] = (float)(ml - mq * a);
  }
  __syncthreads();
  for (int p = prow; p < HW; p += rows_par) {
    float vq[8];
    load_q(p, vq);
#pragma unroll
    for (int j = 0; j < 8; ++j) vq[j] = vq[j] * sa[vcol * 8 + j] + sb[vcol * 8 + j];
    store8_bf16(y + ((size_t)f * HW + p) * ldy + c0, vq);
  }
}

}  // namespace pgt

using namespace pgt;

static int gn_chunks(int HW) {
  int c = ceil_div(HW, 1024);
  return c < 1 ? 1 : (c > GN_MAX_CHUNKS ? GN_MAX_CHUNKS : c);
}

extern "C" int64_t pgt_groupnorm_ws_floats(int F, int HW, int C) {
  (void)C;
  return (int64_t)F * gn_chunks(HW) * GN_GROUPS * 2;
}

extern "C" int pgt_groupnorm_silu(const void* x, int ldx, int F, int HW, int C, const float* gamma, const float* beta,
                                  float eps, int apply_silu, void* y, int ldy, float* ws, void* stream) {
  PGT_CHECK_ARG(x && y && ws && gamma && beta && F > 0 && HW > 0);
  PGT_CHECK_ARG(C % 32 == 0 && C % 8 == 0 && C / 8 <= GN_THREADS && ldx % 8 == 0 && ldy % 8 == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(PGT_PROF_NORM, 3.0 * F * (double)HW * C * 2, st);     // read, read, write (bf16)
  const int nchunks = gn_chunks(HW);
  const int ppc = ceil_div(HW, nchunks);
  gn_stats_kernel<<<dim3(nchunks, F), GN_THREADS, 2 * C * sizeof(float), st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, HW, C, ppc, ws);
  PGT_LAUNCH_OK();
  // apply: ~16 KB of activations per block
  int ppb = (16384 / (C * 2));
  if (ppb < 8) ppb = 8;
  const int nblk = ceil_div(HW, ppb);
  gn_apply_kernel<<<dim3(nblk, F), 256, (2 * C + 2 * GN_GROUPS) * sizeof(float), st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, HW, C, ppb, nchunks, ws, gamma, beta, eps, apply_silu,
      reinterpret_cast<__nv_bfloat16*>(y), ldy);
  PGT_LAUNCH_OK();
  return PGT_OK;
}

extern "C" int pgt_layernorm(const void* x, int ldx, int x_dtype, int T, int C, const float* gamma, const float* beta,
                             float eps, void* y, int ldy, const void* pos, int ldpos, void* y2, int ldy2,
                             void* stream) {
  PGT_CHECK_ARG(x && y && gamma && beta && T > 0);
  PGT_CHECK_ARG(C % 256 == 0 && C <= 1024 && ldx % 8 == 0 && ldy % 8 == 0);
  PGT_CHECK_ARG(y2 == nullptr || (pos != nullptr && ldpos % 8 == 0 && ldy2 % 8 == 0));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int warps_per_block = 8;
  const int grid = ceil_div(T, warps_per_block);
  layernorm_kernel<4><<<grid, warps_per_block * 32, 0, st>>>(
      x, ldx, x_dtype, T, C, gamma, beta, eps, reinterpret_cast<__nv_bfloat16*>(y), ldy,
      reinterpret_cast<const __nv_bfloat16*>(pos), ldpos, reinterpret_cast<__nv_bfloat16*>(y2), ldy2);
  PGT_LAUNCH_OK();
  return PGT_OK;
}

extern "C" int pgt_adain(const void* q, int ldq, int q_dtype, const void* l, int ldl, int F, int HW, int C, float eps,
                         void* y, int ldy, void* stream) {
  PGT_CHECK_ARG(q && l && y && F > 0 && HW > 1 && C % 64 == 0 && ldq % 8 == 0 && ldl % 8 == 0 && ldy % 8 == 0);
  adain_kernel<<<dim3(C / 64, F), ADAIN_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      q, ldq, q_dtype, reinterpret_cast<const __nv_bfloat16*>(l), ldl, HW, C, eps,
      reinterpret_cast<__nv_bfloat16*>(y), ldy);
  PGT_LAUNCH_OK();
  return PGT_OK;
}
