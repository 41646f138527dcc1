// Layout conversion and small data-movement kernels (HBM-bound, 128-bit vectorised), and the streaming video
// front / back end.
#include <algorithm>

#include "common.cuh"
#include "tmap.cuh"

namespace pgt {

__global__ void copy2d_kernel(const __nv_bfloat16* __restrict__ x, int ldx, size_t T, int C,
                              __nv_bfloat16* __restrict__ y, int ldy) {
  const int vc = C >> 3;
  const size_t total = T * vc;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < total; i += 4 * stride) {       // four independent 16-byte copies in flight
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t j = i + k * stride;
      u[k] = __ldg(reinterpret_cast<const uint4*>(x + (j / vc) * ldx) + (j % vc));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t j = i + k * stride;
      reinterpret_cast<uint4*>(y + (j / vc) * ldy)[j % vc] = u[k];
    }
  }
  for (; i < total; i += stride)
    reinterpret_cast<uint4*>(y + (i / vc) * ldy)[i % vc] = __ldg(reinterpret_cast<const uint4*>(x + (i / vc) * ldx) + (i % vc));
}

// fp32 NCHW -> bf16 NHWC through a 32x32 shared-memory transpose (coalesced on both sides).
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int C, int HW, const float* __restrict__ mean,
                                    const float* __restrict__ stdv, __nv_bfloat16* __restrict__ y, int ldy) {
  __shared__ float tile[32][33];
  const int f = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    float v = 0.f;
    if (c < C && p < HW) {
      v = x[((size_t)f * C + c) * HW + p];
      if (mean != nullptr) v = (v - mean[c]) / stdv[c];
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    if (p < HW && c < ldy && c0 + 32 <= ((C + 31) / 32) * 32) {
      if (c < C) y[((size_t)f * HW + p) * ldy + c] = __float2bfloat16_rn(tile[threadIdx.x][j]);
    }
  }
}

__global__ void nhwc_to_f32_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int HW, int C, float* __restrict__ y,
                                   int to_nchw) {
  __shared__ float tile[32][33];
  const int f = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (p < HW && c < C) ? __bfloat162float(x[((size_t)f * HW + p) * ldx + c]) : 0.f;
  }
  __syncthreads();
  if (to_nchw) {
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const int c = c0 + j, p = p0 + threadIdx.x;
      if (c < C && p < HW) y[((size_t)f * C + c) * HW + p] = tile[threadIdx.x][j];
    }
  } else {
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const int p = p0 + j, c = c0 + threadIdx.x;
      if (c < C && p < HW) y[((size_t)f * HW + p) * C + c] = tile[j][threadIdx.x];
    }
  }
}

}  // namespace pgt

using namespace pgt;

static int ew_grid(size_t total, int threads) {
  size_t b = (total + threads - 1) / threads;
  const size_t cap = (size_t)num_sms() * 32;
  return (int)(b < cap ? (b ? b : 1) : cap);
}

extern "C" int pgt_copy2d(const void* x, int ldx, int T, int C, void* y, int ldy, void* stream) {
  PGT_CHECK_ARG(x && y && T > 0 && C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0);
  ProfScope ps(PGT_PROF_MOVE, 2.0 * (double)T * C * 2, static_cast<cudaStream_t>(stream), "pgt_copy2d");
  const size_t total = (size_t)T * (C / 8);
  copy2d_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, (size_t)T, C, reinterpret_cast<__nv_bfloat16*>(y), ldy);
  PGT_LAUNCH_OK();
  return PGT_OK;
}

extern "C" int pgt_nchw_f32_to_nhwc_bf16(const float* x, int F, int C, int HW, const float* mean, const float* stdv,
                                         void* y, int ldy, void* stream) {
  PGT_CHECK_ARG(x && y && F > 0 && C > 0 && HW > 0 && ldy >= C && (mean == nullptr) == (stdv == nullptr));
  ProfScope ps(PGT_PROF_MOVE, 6.0 * F * (double)C * HW, static_cast<cudaStream_t>(stream), "pgt_nchw_f32_to_nhwc_bf16");
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), F);
  nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      x, C, HW, mean, stdv, reinterpret_cast<__nv_bfloat16*>(y), ldy);
  PGT_LAUNCH_OK();
  return PGT_OK;
}

extern "C" int pgt_nhwc_bf16_to_f32(const void* x, int ldx, int F, int HW, int C, float* y, int to_nchw, void* stream) {
  PGT_CHECK_ARG(x && y && F > 0 && C > 0 && HW > 0 && ldx >= C);
  ProfScope ps(PGT_PROF_MOVE, 6.0 * F * (double)C * HW, static_cast<cudaStream_t>(stream), "pgt_nhwc_bf16_to_f32");
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), F);
  nhwc_to_f32_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, HW, C, y, to_nchw);
  PGT_LAUNCH_OK();
  return PGT_OK;
}

// ---- temporal regroup used by the SFT fusion block's cross-frame 1x1 mixers
// dir 0: x [b,3,P,C] -> y [b,P,3*C] (channel = frame*C + c);   dir 1: the inverse.
namespace pgt {
__global__ void regroup_frames_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int b, int P, int C,
                                      __nv_bfloat16* __restrict__ y, int ldy, int dir) {
  const int vc = C >> 3;
  const size_t total = (size_t)b * 3 * P * vc;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % vc);
    size_t r = i / vc;
    const int p = (int)(r % P); r /= P;
    const int fr = (int)(r % 3);
    const int clip = (int)(r / 3);
    const size_t frame_row = ((size_t)clip * 3 + fr) * P + p;          // [b,3,P] row
    const size_t clip_row = (size_t)clip * P + p;                      // [b,P] row
    if (dir == 0)
      reinterpret_cast<uint4*>(y + clip_row * ldy + fr * C)[v] = __ldg(reinterpret_cast<const uint4*>(x + frame_row * ldx) + v);
    else
      reinterpret_cast<uint4*>(y + frame_row * ldy)[v] = __ldg(reinterpret_cast<const uint4*>(x + clip_row * ldx + fr * C) + v);
  }
}
}  // namespace pgt

extern "C" int pgt_regroup_frames(const void* x, int ldx, int clips, int P, int C, void* y, int ldy, int dir,
                                  void* stream) {
  PGT_CHECK_ARG(x && y && clips > 0 && P > 0 && C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && (dir == 0 || dir == 1));
  ProfScope ps(PGT_PROF_MOVE, 4.0 * clips * 3.0 * P * C, static_cast<cudaStream_t>(stream), "pgt_regroup_frames");
  const size_t total = (size_t)clips * 3 * P * (C / 8);
  regroup_frames_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, clips, P, C, reinterpret_cast<__nv_bfloat16*>(y), ldy, dir);
  PGT_LAUNCH_OK();
  return PGT_OK;
}


// ---------------------------------------------------------------- streaming video front / back end
// The reference's loop (inference.py:6-19,37-76) turns rgb24 frames into float tensors with numpy
// (`np.array(rgb / 255.0, np.float32)`: a double division rounded to fp32), runs the model on the window
// (f[i-1], f[i], f[i+1]) and writes `clamp(out[0][1], 0, 1) * 255` truncated to uint8.  These kernels are those two
// conversions, and the frame gather that lets per-frame work (BiSeNet, the attention-free encoder levels) be computed
// once per distinct frame and handed to every window that contains the frame.
namespace pgt {
__constant__ float c_u8_to_unit[256];             // (float)(v / 255.0) evaluated in double on the host

__global__ void u8hwc_to_f32nchw_kernel(const uint8_t* __restrict__ x, size_t HW, size_t total, float* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t f = i / HW, p = i - f * HW;
    const uint8_t* s = x + (f * HW + p) * 3;
    float* d = y + f * 3 * HW + p;
    d[0] = c_u8_to_unit[s[0]];
    d[HW] = c_u8_to_unit[s[1]];
    d[2 * HW] = c_u8_to_unit[s[2]];
  }
}

__global__ void f32nchw_to_u8hwc_kernel(const float* __restrict__ x, size_t HW, int first, int step, size_t total,
                                        uint8_t* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t j = i / HW, p = i - j * HW;
    const float* s = x + ((size_t)first + j * step) * 3 * HW + p;
    uint8_t* d = y + (j * HW + p) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = fminf(fmaxf(s[c * HW], 0.f), 1.f) * 255.0f;     // torch.clamp, then the fp32 product numpy forms
      d[c] = (uint8_t)v;                                             // astype(uint8) of a value in [0, 255]: truncation
    }
  }
}

__global__ void gather_frames_kernel(const uint4* __restrict__ x, size_t vec_per_frame, const int* __restrict__ idx,
                                     uint4* __restrict__ y) {
  const uint4* s = x + (size_t)idx[blockIdx.y] * vec_per_frame;
  uint4* d = y + (size_t)blockIdx.y * vec_per_frame;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < vec_per_frame; i += (size_t)gridDim.x * blockDim.x)
    d[i] = __ldg(s + i);
}
}  // namespace pgt

extern "C" int pgt_u8hwc_to_f32nchw(const void* x_u8, int F, int H, int W, float* y, void* stream) {
  PGT_CHECK_ARG(x_u8 && y && F > 0 && H > 0 && W > 0);
  static pgt::PerDeviceOnce once;               // __constant__ memory is per device
  PGT_CUDA_OK(once.run([] {
    float t[256];
    for (int v = 0; v < 256; ++v) t[v] = (float)((double)v / 255.0);
    return cudaMemcpyToSymbol(pgt::c_u8_to_unit, t, sizeof(t));
  }));
  const size_t HW = (size_t)H * W, total = (size_t)F * HW;
  ProfScope ps(PGT_PROF_MOVE, 15.0 * (double)total, static_cast<cudaStream_t>(stream), "pgt_u8hwc_to_f32nchw");
  pgt::u8hwc_to_f32nchw_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(x_u8), HW, total, y);
  PGT_LAUNCH_OK();
  return PGT_OK;
}

extern "C" int pgt_f32nchw_to_u8hwc(const float* x, int first, int step, int n, int H, int W, void* y_u8, void* stream) {
  PGT_CHECK_ARG(x && y_u8 && n > 0 && H > 0 && W > 0 && first >= 0 && step >= 1);
  const size_t HW = (size_t)H * W, total = (size_t)n * HW;
  ProfScope ps(PGT_PROF_MOVE, 15.0 * (double)total, static_cast<cudaStream_t>(stream), "pgt_f32nchw_to_u8hwc");
  pgt::f32nchw_to_u8hwc_kernel<<<ew_grid(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, HW, first, step, total, static_cast<uint8_t*>(y_u8));
  PGT_LAUNCH_OK();
  return PGT_OK;
}

extern "C" int pgt_gather_frames(const void* x, long long frame_bytes, const int* idx_dev, int n, void* y, void* stream) {
  PGT_CHECK_ARG(x && y && idx_dev && n > 0 && frame_bytes > 0 && frame_bytes % 16 == 0);
  PGT_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0);
  const size_t vec = (size_t)frame_bytes / 16;
  ProfScope ps(PGT_PROF_MOVE, 2.0 * (double)n * frame_bytes, static_cast<cudaStream_t>(stream), "pgt_gather_frames");
  unsigned bx = (unsigned)std::min<size_t>((vec + 255) / 256, 1024);
  pgt::gather_frames_kernel<<<dim3(bx, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), vec, idx_dev, static_cast<uint4*>(y));
  PGT_LAUNCH_OK();
  return PGT_OK;
}
