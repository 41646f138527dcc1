// tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   out[rows, N] = epilogue( A[rows, K] * W[N, K]^T )          bf16 operands, fp32 accumulate in TMEM
//
// One persistent, warp-specialised kernel (192 threads):
//   warp 0      TMA producer: A tile (128 rows x 64 K) and W tile (BN rows x 64 K) per k-block, 128B swizzle
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, cta_group::1)
//   warps 2..5  epilogue: tcgen05.ld (32x32b) -> bias / activation / residual / SFT -> global store
// Pipelines: smem full/empty ring (TMA <-> MMA) and a 2-deep TMEM accumulator ring (MMA <-> epilogue), so
// the epilogue of tile i overlaps the main loop of tile i+1.
//
// The A operand is produced by TMA in three addressing modes:
//   LINEAR   2-D map [K, rows]
//   CONV_S1  4-D map [C, W, H, F]; the 128-row tile is a (tn x th x tw) pixel patch and every 3x3 tap is the
//            same box shifted by (dx-1, dy-1); out-of-bounds pixels / channels are zero-filled by TMA, which
//            is exactly the conv zero padding — no im2col buffer, no halo handling in software
//   CONV_S2  5-D map [2C, W/2, 2, H/2, F] (row / column parity split) for stride-2 convs
#include <cudaTypedefs.h>

#include "common.cuh"
#include "ptx.cuh"

namespace pgt {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int A_STAGE_BYTES = BM * BK * 2;

enum { MODE_LINEAR = 0, MODE_CONV_S1 = 1, MODE_CONV_S2 = 2 };

struct GemmParams {
  int mode;
  int M, N, K;
  int num_kb, m_tiles, n_tiles;
  // conv geometry in OUTPUT space
  int F, H, W;
  int tw, th, tn, tiles_x, tiles_y;
  int cin_blocks, ksize, pad_lo, cin_ld;
  // epilogue
  const float* bias;
  int act, epi_mode;
  const void* residual;
  int ldr, res_dtype;
  const void* aux;
  int ldaux;
  float sft_w;
  void* out;
  int ldo, out_dtype, out_layout;
  double flops;      // algorithmic 2*M*N*K with the un-padded K (host-side accounting only)
};

template <int BN>
struct GemmCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (196608 / STAGE_BYTES) > 8 ? 8 : (196608 / STAGE_BYTES);
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void decode_conv_tile(const GemmParams& p, int m_blk, int& n0, int& y0, int& x0) {
  const int tx = m_blk % p.tiles_x;
  const int t2 = m_blk / p.tiles_x;
  const int ty = t2 % p.tiles_y;
  const int tf = t2 / p.tiles_y;
  x0 = tx * p.tw;
  y0 = ty * p.th;
  n0 = tf * p.tn;
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr);
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.n_tiles;
        const int m_blk = tile / p.n_tiles;
        int n0 = 0, y0 = 0, x0 = 0;
        if (p.mode != MODE_LINEAR) decode_conv_tile(p, m_blk, n0, y0, x0);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          void* dst_a = smem_a + stage * A_STAGE_BYTES;
          void* dst_b = smem_b + stage * Cfg::B_STAGE_BYTES;
          if (p.mode == MODE_LINEAR) {
            tma_load_2d(dst_a, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          } else {
            const int tap = kb / p.cin_blocks;
            const int cb = kb - tap * p.cin_blocks;
            const int dy = tap / p.ksize;
            const int dx = tap - dy * p.ksize;
            if (p.mode == MODE_CONV_S1) {
              tma_load_4d(dst_a, &tmA, &full_bar[stage], cb * BK, x0 + dx - p.pad_lo, y0 + dy - p.pad_lo, n0);
            } else {
              const int oy = dy - p.pad_lo, ox = dx - p.pad_lo;
              const int qy = (oy < 0) ? -((1 - oy) >> 1) : (oy >> 1);
              const int qx = (ox < 0) ? -((1 - ox) >> 1) : (ox >> 1);
              const int py = oy - 2 * qy, px = ox - 2 * qx;
              tma_load_5d(dst_a, &tmA, &full_bar[stage], px * p.cin_ld + cb * BK, x0 + qx, py, y0 + qy, n0);
            }
          }
          tma_load_2d(dst_b, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = umma_desc_k_sw128(smem_u32(smem_a + stage * A_STAGE_BYTES));
          const uint64_t db = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes (2 x 16 B units) per UMMA_K=16 step inside the 128 B swizzle row
            umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);              // frees the smem stage when the MMAs retire
          if (kb == p.num_kb - 1) umma_commit(&tmem_full[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int r = quad * 32 + lane;            // row of the 128-row tile
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n_blk = tile % p.n_tiles;
      const int m_blk = tile / p.n_tiles;
      bool valid;
      long long orow;                          // output row (pixel / token) index
      int pn = 0, py = 0, px = 0;
      if (p.mode == MODE_LINEAR) {
        orow = (long long)m_blk * BM + r;
        valid = orow < p.M;
      } else {
        int n0, y0, x0;
        decode_conv_tile(p, m_blk, n0, y0, x0);
        const int ix = r % p.tw;
        const int t2 = r / p.tw;
        const int iy = t2 % p.th;
        const int in = t2 / p.th;
        pn = n0 + in; py = y0 + iy; px = x0 + ix;
        valid = (pn < p.F) && (py < p.H) && (px < p.W);
        orow = ((long long)pn * p.H + py) * p.W + px;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(quad * 32) << 16) + acc * BN;
      const int col_base = n_blk * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (col_base + c0 >= p.N) break;       // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(t_row + c0, v);
        tmem_ld_wait();
        if (valid) {
        const int col0 = col_base + c0;
        const int ncol = min(32, p.N - col0);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncol) f[j] += __ldg(p.bias + col0 + j);
        }
        if (p.act != PGT_ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
        }
        const bool vec_ok = (ncol == 32);
        if (p.residual != nullptr) {
          if (p.res_dtype == PGT_BF16) {
            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + orow * p.ldr + col0;
            const bool rvec = vec_ok && ((p.ldr & 7) == 0);
            float rr[32];
            if (rvec) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(rp) + q);
                float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
                rr[q * 8 + 0] = a.x; rr[q * 8 + 1] = a.y; rr[q * 8 + 2] = b.x; rr[q * 8 + 3] = b.y;
                rr[q * 8 + 4] = c.x; rr[q * 8 + 5] = c.y; rr[q * 8 + 6] = d.x; rr[q * 8 + 7] = d.y;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) rr[j] = (j < ncol) ? __bfloat162float(rp[j]) : 0.f;
            }
            if (p.epi_mode == PGT_EPI_SFT) {
              const __nv_bfloat16* ap = reinterpret_cast<const __nv_bfloat16*>(p.aux) + orow * p.ldaux + col0;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float s = (j < ncol) ? __bfloat162float(ap[j]) : 0.f;
                f[j] = rr[j] + p.sft_w * (rr[j] * s + f[j]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] += rr[j];
            }
          } else {
            const float* rp = reinterpret_cast<const float*>(p.residual) + orow * p.ldr + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) f[j] += __ldg(rp + j);
          }
        }
        if (p.out_layout == PGT_OUT_NCHW) {
          float* op = reinterpret_cast<float*>(p.out);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncol) op[(((long long)pn * p.N + (col0 + j)) * p.H + py) * p.W + px] = f[j];
        } else if (p.out_dtype == PGT_BF16) {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + col0;
          if (vec_ok && ((p.ldo & 7) == 0)) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 u;
              u.x = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
              u.y = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
              u.z = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
              u.w = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
              reinterpret_cast<uint4*>(op)[q] = u;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) op[j] = __float2bfloat16_rn(f[j]);
          }
        } else {
          float* op = reinterpret_cast<float*>(p.out) + orow * p.ldo + col0;
          if (vec_ok && ((p.ldo & 3) == 0)) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              reinterpret_cast<float4*>(op)[q] = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) op[j] = f[j];
          }
        }
        }  // valid
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------- host side
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
  }
  return fn;
}

static int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
  auto fn = get_encode_fn();
  if (fn == nullptr) return PGT_ERR_DRIVER;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PGT_OK : PGT_ERR_DRIVER;
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const void* W, int ldw, GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmB;
  {
    // rows beyond N (weight matrices are packed to a multiple of 16 rows, not of BN) are zero-filled
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t str[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {BK, (uint32_t)BN};
    int rc = encode_map(&tmB, W, 2, dims, str, box);
    if (rc != PGT_OK) return rc;
  }
  p.n_tiles = ceil_div(p.N, BN);
  static bool attr_set = false;
  if (!attr_set) {
    PGT_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  {
    ProfScope ps(PGT_PROF_GEMM, p.flops, stream);
    gemm_tc_kernel<BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  }
  PGT_LAUNCH_OK();
  return PGT_OK;
}

static int pick_bn(int N, int m_tiles) {
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  // wide N: prefer 256-wide tiles when they still fill the machine
  const int t256 = m_tiles * ceil_div(N, 256);
  return (t256 >= num_sms()) ? 256 : 128;
}

static int dispatch_gemm(const CUtensorMap& tmA, const void* W, int ldw, GemmParams& p, cudaStream_t stream) {
  switch (pick_bn(p.N, p.m_tiles)) {
    case 32: return launch_gemm<32>(tmA, W, ldw, p, stream);
    case 64: return launch_gemm<64>(tmA, W, ldw, p, stream);
    case 128: return launch_gemm<128>(tmA, W, ldw, p, stream);
    default: return launch_gemm<256>(tmA, W, ldw, p, stream);
  }
}

static int fill_epilogue(GemmParams& p, const pgt_epilogue* ep) {
  if (ep == nullptr || ep->out == nullptr) return PGT_ERR_INVALID;
  p.bias = ep->bias;
  p.act = ep->act;
  p.epi_mode = ep->mode;
  p.residual = ep->residual;
  p.ldr = ep->ldr;
  p.res_dtype = ep->res_dtype;
  p.aux = ep->aux;
  p.ldaux = ep->ldaux;
  p.sft_w = ep->sft_w;
  p.out = ep->out;
  p.ldo = ep->ldo;
  p.out_dtype = ep->out_dtype;
  p.out_layout = ep->out_layout;
  if (p.epi_mode == PGT_EPI_SFT && (p.residual == nullptr || p.aux == nullptr || p.res_dtype != PGT_BF16))
    return PGT_ERR_INVALID;
  if (p.out_layout == PGT_OUT_NCHW && (p.out_dtype != PGT_F32 || p.mode == MODE_LINEAR)) return PGT_ERR_INVALID;
  return PGT_OK;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace pgt

using namespace pgt;

extern "C" int pgt_linear_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                               const pgt_epilogue* ep, void* stream) {
  PGT_CHECK_ARG(A && W && M > 0 && N > 0 && K > 0);
  PGT_CHECK_ARG((lda % 8) == 0 && (ldw % 8) == 0 && aligned16(A) && aligned16(W));
  GemmParams p{};
  p.mode = MODE_LINEAR;
  p.M = M; p.N = N; p.K = K;
  p.num_kb = ceil_div(K, BK);
  p.m_tiles = ceil_div(M, BM);
  p.flops = 2.0 * M * (double)N * K;
  int rc = fill_epilogue(p, ep);
  if (rc != PGT_OK) return rc;
  CUtensorMap tmA;
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
  uint64_t str[1] = {(uint64_t)lda * 2};
  uint32_t box[2] = {BK, BM};
  rc = encode_map(&tmA, A, 2, dims, str, box);
  if (rc != PGT_OK) return rc;
  return dispatch_gemm(tmA, W, ldw, p, static_cast<cudaStream_t>(stream));
}

extern "C" int pgt_conv_bf16(const void* x, int F, int Hin, int Win, int Cin, int ldx, const void* Wp, int ldw,
                             int Cout, int ksize, int stride, int pad_lo, const pgt_epilogue* ep, void* stream) {
  PGT_CHECK_ARG(x && Wp && F > 0 && Hin > 0 && Win > 0 && Cin > 0 && Cout > 0);
  PGT_CHECK_ARG((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2) && pad_lo >= 0 && pad_lo <= 1);
  PGT_CHECK_ARG((ldx % 8) == 0 && (ldw % 8) == 0 && aligned16(x) && aligned16(Wp) && ldx >= Cin);
  const int cin_pad = ceil_div(Cin, BK) * BK;
  PGT_CHECK_ARG(ldw >= ksize * ksize * cin_pad);
  GemmParams p{};
  p.N = Cout;
  p.ksize = ksize;
  p.pad_lo = pad_lo;
  p.cin_blocks = cin_pad / BK;
  p.K = ksize * ksize * cin_pad;
  p.num_kb = ksize * ksize * p.cin_blocks;
  p.F = F;
  CUtensorMap tmA;
  int rc;
  if (stride == 1) {
    p.mode = MODE_CONV_S1;
    p.H = Hin; p.W = Win;
  } else {
    if ((Hin & 1) || (Win & 1) || (Cin % BK) != 0 || ldx != Cin) return PGT_ERR_UNSUPPORTED;
    p.mode = MODE_CONV_S2;
    p.H = Hin / 2; p.W = Win / 2;
    p.cin_ld = ldx;
  }
  rc = fill_epilogue(p, ep);
  if (rc != PGT_OK) return rc;
  // 128-pixel tile = tn frames x th rows x tw columns
  int tw = 1;
  while (tw * 2 <= p.W && tw * 2 <= BM) tw *= 2;
  int th = 1;
  while (th * 2 <= p.H && tw * th * 2 <= BM) th *= 2;
  int tn = BM / (tw * th);
  p.tw = tw; p.th = th; p.tn = tn;
  p.tiles_x = ceil_div(p.W, tw);
  p.tiles_y = ceil_div(p.H, th);
  p.m_tiles = p.tiles_x * p.tiles_y * ceil_div(F, tn);
  p.M = F * p.H * p.W;
  p.flops = 2.0 * p.M * (double)Cout * (ksize * ksize * Cin);
  if (p.mode == MODE_CONV_S1) {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Win, (uint64_t)Hin, (uint64_t)F};
    uint64_t str[3] = {(uint64_t)ldx * 2, (uint64_t)Win * ldx * 2, (uint64_t)Hin * Win * ldx * 2};
    uint32_t box[4] = {BK, (uint32_t)tw, (uint32_t)th, (uint32_t)tn};
    rc = encode_map(&tmA, x, 4, dims, str, box);
  } else {
    uint64_t dims[5] = {(uint64_t)2 * ldx, (uint64_t)Win / 2, 2, (uint64_t)Hin / 2, (uint64_t)F};
    uint64_t str[4] = {(uint64_t)2 * ldx * 2, (uint64_t)Win * ldx * 2, (uint64_t)2 * Win * ldx * 2,
                       (uint64_t)Hin * Win * ldx * 2};
    uint32_t box[5] = {BK, (uint32_t)tw, 1, (uint32_t)th, (uint32_t)tn};
    rc = encode_map(&tmA, x, 5, dims, str, box);
  }
  if (rc != PGT_OK) return rc;
  return dispatch_gemm(tmA, Wp, ldw, p, static_cast<cudaStream_t>(stream));
}
