// The C-ABI of include/rcu_b200.h exposed to PyTorch as a torch extension: custom operators in the `rcu_b200` namespace
// (torch.ops.rcu_b200.*) that take and return tensors, launch on PyTorch's current CUDA stream and raise the exceptions
// the reference would (RCU_EINVAL -> ValueError, RCU_ENOTSUP -> NotImplementedError).  No pybind, no Python.h: the
// library is loaded with torch.ops.load_library.  It links against librcu_b200.so next to it ($ORIGIN rpath); the
// kernels live there, this file is plumbing.  The ctypes binding (_lib.py) stays the reference binding of the ABI —
// _torch_ext.py says which calls go through here and tools/binding_overhead.py measures the difference.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <optional>
#include <vector>

#include "../../include/rcu_b200.h"

namespace {

void check(int rc) {
  if (rc == RCU_OK) return;
  const char* msg = rcu_last_error();
  if (rc == RCU_EINVAL) TORCH_CHECK_VALUE(false, msg ? msg : "rcu_b200: invalid argument");
  if (rc == RCU_ENOTSUP) TORCH_CHECK_NOT_IMPLEMENTED(false, msg ? msg : "rcu_b200: not supported");
  TORCH_CHECK(false, "rcu_b200 error ", rc, ": ", msg ? msg : "");
}

void* stream_of(const at::Tensor& t) { return c10::cuda::getCurrentCUDAStream(t.device().index()).stream(); }

void need_cuda(const at::Tensor& t, at::ScalarType dtype, const char* name) {
  TORCH_CHECK_VALUE(t.is_cuda() && t.scalar_type() == dtype && t.is_contiguous(), name, " must be a contiguous CUDA tensor of dtype ", dtype);
}

// ECE tables + U-E joint table in one pass (rcu_eval_fused).  Returns the flat int64 table
// [count S x (B+1) | positives | conf_sum (float64 bits) | ue S x 4 x K | invalid S].
at::Tensor eval_fused(const at::Tensor& p, const at::Tensor& prediction, const at::Tensor& target, const std::optional<at::Tensor>& mask,
                      const at::Tensor& edges_f32, const at::Tensor& breaks_f32, const at::Tensor& seg_class, int64_t n_subjects,
                      int64_t n_classes, at::Tensor workspace) {
  need_cuda(p, at::kFloat, "p");
  need_cuda(prediction, at::kByte, "prediction");
  need_cuda(target, at::kByte, "target");
  if (mask) need_cuda(*mask, at::kByte, "mask");
  need_cuda(workspace, at::kByte, "workspace");
  TORCH_CHECK_VALUE(!edges_f32.is_cuda() && edges_f32.scalar_type() == at::kFloat && edges_f32.is_contiguous() && edges_f32.numel() >= 2,
                    "edges_f32 must be a contiguous float32 host tensor");
  TORCH_CHECK_VALUE(!breaks_f32.is_cuda() && breaks_f32.scalar_type() == at::kFloat && breaks_f32.is_contiguous(), "breaks_f32 must be a host float32 tensor");
  TORCH_CHECK_VALUE(!seg_class.is_cuda() && seg_class.scalar_type() == at::kByte && seg_class.numel() == breaks_f32.numel() + 1,
                    "seg_class must be a host uint8 tensor with one entry more than breaks_f32");
  TORCH_CHECK_VALUE(n_subjects >= 1 && p.numel() % n_subjects == 0, p.numel(), " elements do not split into ", n_subjects, " subjects");
  const int64_t n = p.numel(), vps = n / n_subjects, n_bins = edges_f32.numel() - 1, nb1 = n_bins + 1;
  TORCH_CHECK_VALUE(prediction.numel() == n && target.numel() == n && (!mask || mask->numel() == n), "input sizes differ");
  c10::cuda::CUDAGuard guard(p.device());
  at::Tensor flat = at::empty({n_subjects * (3 * nb1 + 4 * n_classes + 1)}, p.options().dtype(at::kLong));
  if (n == 0) return flat.zero_();
  int64_t* base = flat.data_ptr<int64_t>();
  const int64_t o1 = n_subjects * nb1, o3 = 3 * o1, o4 = o3 + n_subjects * 4 * n_classes;
  check(rcu_eval_fused(p.data_ptr<float>(), prediction.data_ptr<uint8_t>(), target.data_ptr<uint8_t>(), mask ? mask->data_ptr<uint8_t>() : nullptr,
                       vps, (int)n_subjects, edges_f32.data_ptr<float>(), (int)n_bins, breaks_f32.data_ptr<float>(), (int)breaks_f32.numel(),
                       seg_class.data_ptr<uint8_t>(), (int)n_classes, reinterpret_cast<uint64_t*>(base), reinterpret_cast<uint64_t*>(base + o1),
                       reinterpret_cast<double*>(base + 2 * o1), reinterpret_cast<uint64_t*>(base + o3), reinterpret_cast<uint64_t*>(base + o4),
                       workspace.data_ptr(), (size_t)workspace.numel(), stream_of(p)));
  return flat;
}

// MultiPredictionSummary on interleaved logits (T, N, H, W, 2) [+ the weight-scaling sample (N, H, W, 2)] in one launch.
// Returns [mean (N,2,H,W), entropy (N,1,H,W), mutual_info | empty, variance | empty, prediction | empty, foreground | empty, ws | empty].
std::vector<at::Tensor> aggregate(const at::Tensor& logits, const std::optional<at::Tensor>& ws_logits, bool want_mi, bool want_var,
                                  bool emit_prediction, bool emit_foreground) {
  need_cuda(logits, at::kFloat, "logits");
  TORCH_CHECK_VALUE(logits.dim() == 5 && logits.size(4) == 2, "logits must have shape (T, N, H, W, 2)");
  const int64_t t = logits.size(0), n = logits.size(1), h = logits.size(2), w = logits.size(3);
  c10::cuda::CUDAGuard guard(logits.device());
  auto f32 = logits.options();
  at::Tensor none = at::empty({0}, f32);
  at::Tensor mean = at::empty({n, 2, h, w}, f32), entropy = at::empty({n, 1, h, w}, f32);
  at::Tensor mi = want_mi ? at::empty({n, 1, h, w}, f32) : none, var = want_var ? at::empty({n, 1, h, w}, f32) : none;
  at::Tensor pred = emit_prediction ? at::empty({n, h, w}, f32.dtype(at::kByte)) : at::empty({0}, f32.dtype(at::kByte));
  at::Tensor fg = emit_foreground ? at::empty({n, h, w}, f32) : none;
  at::Tensor ws = none;
  auto opt = [](const at::Tensor& x) -> float* { return x.numel() ? x.data_ptr<float>() : nullptr; };
  if (ws_logits) {
    need_cuda(*ws_logits, at::kFloat, "ws_logits");
    TORCH_CHECK_VALUE(ws_logits->dim() == 4 && ws_logits->size(0) == n && ws_logits->size(1) == h && ws_logits->size(2) == w && ws_logits->size(3) == 2,
                      "ws_logits must have shape (N, H, W, 2)");
    ws = at::empty({n, 2, h, w}, f32);
    check(rcu_aggregate_ws(logits.data_ptr<float>(), (int)t, n, h * w, ws_logits->data_ptr<float>(), ws.data_ptr<float>(), mean.data_ptr<float>(),
                           entropy.data_ptr<float>(), opt(mi), opt(var), pred.numel() ? pred.data_ptr<uint8_t>() : nullptr, opt(fg), stream_of(logits)));
  } else {
    check(rcu_aggregate(logits.data_ptr<float>(), 0, (int)t, n, h * w, mean.data_ptr<float>(), entropy.data_ptr<float>(), opt(mi), opt(var),
                        pred.numel() ? pred.data_ptr<uint8_t>() : nullptr, opt(fg), nullptr, stream_of(logits)));
  }
  return {mean, entropy, mi, var, pred, fg, ws};
}

// The same on logit differences l0 - l1: (T, N, H, W) [+ the weight-scaling sample (N, H, W)] (rcu_aggregate_ws_diff / input_kind 3).
std::vector<at::Tensor> aggregate_diff(const at::Tensor& diff, const std::optional<at::Tensor>& ws_diff, bool want_mi, bool want_var,
                                       bool emit_prediction, bool emit_foreground) {
  need_cuda(diff, at::kFloat, "logit_diff");
  TORCH_CHECK_VALUE(diff.dim() == 4, "logit_diff must have shape (T, N, H, W)");
  const int64_t t = diff.size(0), n = diff.size(1), h = diff.size(2), w = diff.size(3);
  c10::cuda::CUDAGuard guard(diff.device());
  auto f32 = diff.options();
  at::Tensor none = at::empty({0}, f32);
  at::Tensor mean = at::empty({n, 2, h, w}, f32), entropy = at::empty({n, 1, h, w}, f32);
  at::Tensor mi = want_mi ? at::empty({n, 1, h, w}, f32) : none, var = want_var ? at::empty({n, 1, h, w}, f32) : none;
  at::Tensor pred = emit_prediction ? at::empty({n, h, w}, f32.dtype(at::kByte)) : at::empty({0}, f32.dtype(at::kByte));
  at::Tensor fg = emit_foreground ? at::empty({n, h, w}, f32) : none;
  at::Tensor ws = none;
  auto opt = [](const at::Tensor& x) -> float* { return x.numel() ? x.data_ptr<float>() : nullptr; };
  if (ws_diff) {
    need_cuda(*ws_diff, at::kFloat, "ws_logit_diff");
    TORCH_CHECK_VALUE(ws_diff->dim() == 3 && ws_diff->size(0) == n && ws_diff->size(1) == h && ws_diff->size(2) == w, "ws_logit_diff must have shape (N, H, W)");
    ws = at::empty({n, 2, h, w}, f32);
    check(rcu_aggregate_ws_diff(diff.data_ptr<float>(), (int)t, n, h * w, ws_diff->data_ptr<float>(), ws.data_ptr<float>(), mean.data_ptr<float>(),
                                entropy.data_ptr<float>(), opt(mi), opt(var), pred.numel() ? pred.data_ptr<uint8_t>() : nullptr, opt(fg), stream_of(diff)));
  } else {
    check(rcu_aggregate(diff.data_ptr<float>(), 3, (int)t, n, h * w, mean.data_ptr<float>(), entropy.data_ptr<float>(), opt(mi), opt(var),
                        pred.numel() ? pred.data_ptr<uint8_t>() : nullptr, opt(fg), nullptr, stream_of(diff)));
  }
  return {mean, entropy, mi, var, pred, fg, ws};
}

// rcu_unet_forward on an engine handle (the integer value of the rcu_unet* the Python side owns; its workspace is bound).
at::Tensor unet_forward(int64_t handle, const at::Tensor& images, int64_t n_samples, int64_t dropout_mode, bool det_first, int64_t seed,
                        int64_t slice_index0, int64_t sample0, const std::optional<at::Tensor>& scale) {
  need_cuda(images, at::kFloat, "images");
  TORCH_CHECK_VALUE(images.dim() == 4, "images must have shape (N, C, H, W)");
  if (scale) need_cuda(*scale, at::kFloat, "scale");
  c10::cuda::CUDAGuard guard(images.device());
  at::Tensor logits = at::empty({n_samples, images.size(0), images.size(2), images.size(3), 2}, images.options());
  check(rcu_unet_forward(reinterpret_cast<rcu_unet*>(handle), images.data_ptr<float>(), images.size(0), (int)n_samples, (int)dropout_mode,
                         det_first ? 1 : 0, (uint64_t)seed, slice_index0, (int)sample0, scale ? scale->data_ptr<float>() : nullptr,
                         logits.data_ptr<float>(), stream_of(images)));
  return logits;
}

int64_t abi_version() { return rcu_abi_version(); }

}  // namespace

TORCH_LIBRARY(rcu_b200, m) {
  m.def("abi_version() -> int", &abi_version);
  m.def("eval_fused(Tensor p, Tensor prediction, Tensor target, Tensor? mask, Tensor edges_f32, Tensor breaks_f32, Tensor seg_class, "
        "int n_subjects, int n_classes, Tensor(a!) workspace) -> Tensor", &eval_fused);
  m.def("aggregate(Tensor logits, Tensor? ws_logits, bool want_mi, bool want_var, bool emit_prediction, bool emit_foreground) -> Tensor[]", &aggregate);
  m.def("aggregate_diff(Tensor logit_diff, Tensor? ws_logit_diff, bool want_mi, bool want_var, bool emit_prediction, bool emit_foreground) -> Tensor[]",
        &aggregate_diff);
  m.def("unet_forward(int handle, Tensor images, int n_samples, int dropout_mode, bool det_first, int seed, int slice_index0, int sample0, "
        "Tensor? scale) -> Tensor", &unet_forward);
}
