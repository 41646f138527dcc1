// TORCH_LIBRARY shim over the C ABI (include/pgt_b200.h): `torch.ops.pgt.*` for the hot-path kernels, the PyTorch-side
// binding SURVEY 8(b) sketches next to the ctypes one (pgtformer_b200/_lib.py).  Nothing is computed here: every op
// checks dtypes / devices, takes raw device pointers and the current CUDA stream of the tensor's device, and calls the
// same extern "C" entry point the ctypes binding calls.  Built by pgtformer_b200/build.py into lib/libpgt_torch.so
// (plain g++, links libpgt_b200.so); loaded on demand by pgtformer_b200/torch_ops.py.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include "../../include/pgt_b200.h"

namespace {

void check(int rc, const char* what) {
  if (rc == PGT_OK) return;
  std::string msg = std::string("libpgt_b200 (") + what + "): " + pgt_strerror(rc);
  if (rc == PGT_ERR_CUDA) msg += std::string(": ") + pgt_last_cuda_error();
  TORCH_CHECK(false, msg);
}

void* stream_of(const at::Tensor& t) { return at::cuda::getCurrentCUDAStream(t.device().index()).stream(); }

int ld(const at::Tensor& t) {           // row pitch (elements) of a channels-last [..., C] tensor / view
  TORCH_CHECK(t.stride(-1) == 1, "expected a channels-last tensor");
  return t.dim() > 1 ? (int)t.stride(-2) : (int)t.size(-1);
}

void window_attention(const at::Tensor& qkv, int64_t clips, int64_t H, int64_t W, int64_t C, int64_t heads, int64_t shift,
                      const at::Tensor& tab16, at::Tensor out) {
  TORCH_CHECK(qkv.is_cuda() && qkv.scalar_type() == at::kBFloat16 && out.scalar_type() == at::kBFloat16 &&
              tab16.scalar_type() == at::kHalf && tab16.is_contiguous(), "window_attention: bf16 qkv / out, fp16 table");
  c10::cuda::CUDAGuard guard(qkv.device());
  check(pgt_window_attention_tc(qkv.data_ptr(), ld(qkv), (int)clips, (int)H, (int)W, (int)C, (int)heads, (int)shift,
                                tab16.data_ptr(), out.data_ptr(), ld(out), 0, stream_of(qkv)), "window_attention");
}

void mha_fwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, int64_t clips, int64_t L, int64_t heads, int64_t d,
             at::Tensor out) {
  TORCH_CHECK(q.is_cuda() && q.scalar_type() == at::kBFloat16 && k.scalar_type() == at::kBFloat16 &&
              v.scalar_type() == at::kBFloat16 && out.scalar_type() == at::kBFloat16, "mha_fwd: bf16 tensors");
  c10::cuda::CUDAGuard guard(q.device());
  check(pgt_mha_fwd(q.data_ptr(), ld(q), k.data_ptr(), ld(k), v.data_ptr(), ld(v), (int)clips, (int)L, (int)heads, (int)d,
                    out.data_ptr(), ld(out), stream_of(q)), "mha_fwd");
}

void argmax_gather(const at::Tensor& logits, const at::Tensor& codebook, at::Tensor idx, at::Tensor quant) {
  TORCH_CHECK(logits.is_cuda() && logits.scalar_type() == at::kFloat && logits.is_contiguous() &&
              codebook.scalar_type() == at::kFloat && idx.scalar_type() == at::kLong, "argmax_gather: fp32 logits / codebook, int64 idx");
  c10::cuda::CUDAGuard guard(logits.device());
  const int qd = quant.scalar_type() == at::kBFloat16 ? PGT_BF16 : PGT_F32;
  check(pgt_argmax_gather(logits.data_ptr<float>(), (int)logits.size(0), (int)logits.size(1), codebook.data_ptr<float>(),
                          (int)codebook.size(1), nullptr, idx.data_ptr<int64_t>(), quant.data_ptr(), ld(quant), qd,
                          stream_of(logits)), "argmax_gather");
}

std::tuple<at::Tensor, at::Tensor> codebook_pack(const at::Tensor& codebook, int64_t K) {
  TORCH_CHECK(codebook.is_cuda() && codebook.scalar_type() == at::kFloat && codebook.is_contiguous() && codebook.size(0) >= K);
  c10::cuda::CUDAGuard guard(codebook.device());
  auto cb16 = at::empty({K, codebook.size(1)}, codebook.options().dtype(at::kBFloat16));
  auto norm = at::empty({K + 2}, codebook.options());
  check(pgt_codebook_pack(codebook.data_ptr<float>(), (int)K, (int)codebook.size(1), cb16.data_ptr(), norm.data_ptr<float>(),
                          stream_of(codebook)), "codebook_pack");
  return std::make_tuple(cb16, norm);
}

void l2_argmin(const at::Tensor& z, const at::Tensor& codebook, const at::Tensor& cb16, const at::Tensor& norm, int64_t K,
               at::Tensor idx, const c10::optional<at::Tensor>& quant) {
  TORCH_CHECK(z.is_cuda() && z.scalar_type() == at::kFloat && z.is_contiguous() && codebook.is_contiguous() &&
              idx.scalar_type() == at::kLong, "l2_argmin: fp32 contiguous z / codebook, int64 idx");
  c10::cuda::CUDAGuard guard(z.device());
  const int T = (int)z.size(0), E = (int)z.size(1);
  auto ws = at::empty({pgt_l2_argmin_ws_ints(T)}, z.options().dtype(at::kInt));
  float* qp = quant.has_value() ? quant->data_ptr<float>() : nullptr;
  int rc = pgt_l2_argmin_tc(z.data_ptr<float>(), T, E, codebook.data_ptr<float>(), cb16.data_ptr(), norm.data_ptr<float>(),
                            (int)K, idx.data_ptr<int64_t>(), qp, ws.data_ptr<int>(), stream_of(z));
  if (rc == PGT_ERR_UNSUPPORTED)
    rc = pgt_l2_argmin(z.data_ptr<float>(), T, E, codebook.data_ptr<float>(), (int)K, idx.data_ptr<int64_t>(), qp, stream_of(z));
  check(rc, "l2_argmin");
}

void linear(const at::Tensor& a, const at::Tensor& w, const c10::optional<at::Tensor>& bias, int64_t act,
            const c10::optional<at::Tensor>& residual, at::Tensor out) {
  TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16 && w.stride(1) == 1,
              "linear: bf16 a [M, K], w [N, K]");
  c10::cuda::CUDAGuard guard(a.device());
  pgt_epilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.bias = bias.has_value() ? bias->data_ptr<float>() : nullptr;
  ep.act = (int)act;
  ep.mode = PGT_EPI_PLAIN;
  if (residual.has_value()) {
    ep.residual = residual->data_ptr();
    ep.ldr = ld(*residual);
    ep.res_dtype = residual->scalar_type() == at::kBFloat16 ? PGT_BF16 : PGT_F32;
  }
  ep.out = out.data_ptr();
  ep.ldo = ld(out);
  ep.out_dtype = out.scalar_type() == at::kBFloat16 ? PGT_BF16 : PGT_F32;
  ep.out_layout = PGT_OUT_NHWC;
  const int64_t M = a.numel() / a.size(-1);
  check(pgt_linear_bf16(a.data_ptr(), ld(a), w.data_ptr(), (int)w.stride(0), (int)M, (int)w.size(0), (int)a.size(-1), &ep,
                        stream_of(a)), "linear");
}

}  // namespace

TORCH_LIBRARY(pgt, m) {
  m.def("window_attention(Tensor qkv, int clips, int H, int W, int C, int heads, int shift, Tensor tab16, Tensor(a!) out) -> ()");
  m.def("mha_fwd(Tensor q, Tensor k, Tensor v, int clips, int L, int heads, int d, Tensor(a!) out) -> ()");
  m.def("argmax_gather(Tensor logits, Tensor codebook, Tensor(a!) idx, Tensor(b!) quant) -> ()");
  m.def("codebook_pack(Tensor codebook, int K) -> (Tensor, Tensor)");
  m.def("l2_argmin(Tensor z, Tensor codebook, Tensor cb16, Tensor norm, int K, Tensor(a!) idx, Tensor(b!)? quant) -> ()");
  m.def("linear(Tensor a, Tensor w, Tensor? bias, int act, Tensor? residual, Tensor(a!) out) -> ()");
}

TORCH_LIBRARY_IMPL(pgt, CUDA, m) {
  m.impl("window_attention", window_attention);
  m.impl("mha_fwd", mha_fwd);
  m.impl("argmax_gather", argmax_gather);
  m.impl("codebook_pack", codebook_pack);
  m.impl("l2_argmin", l2_argmin);
  m.impl("linear", linear);
}
