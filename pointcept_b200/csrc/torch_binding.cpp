// Thin compiled PyTorch binding over the C ABI (include/b2pc.h): allocation, the current CUDA stream and the autograd graph
// live here; every kernel is reached through the extern "C" entry points of libb2pc.so, exactly like the ctypes binding in
// pointcept_b200/_lib.py (which stays the reference binding and the one the parity tests can force with B2PC_BINDING=ctypes).
// Why it exists: the PT-v3 step at 2 scenes per GPU is host-bound; a Python autograd.Function costs ~25 us per direction,
// the same node in C++ a few.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "../../include/b2pc.h"

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

namespace {

inline b2pc_stream_t cur_stream() { return (b2pc_stream_t)at::cuda::getCurrentCUDAStream().stream(); }

inline int dt(const Tensor& t) {
  switch (t.scalar_type()) {
    case at::kFloat: return B2PC_F32;
    case at::kHalf: return B2PC_F16;
    case at::kBFloat16: return B2PC_BF16;
    default: TORCH_CHECK(false, "b2pc: unsupported dtype ", t.scalar_type());
  }
}
inline at::ScalarType st(int code) { return code == B2PC_F32 ? at::kFloat : (code == B2PC_F16 ? at::kHalf : at::kBFloat16); }

inline void check(int status, const char* what) {
  TORCH_CHECK(status == 0, "b2pc ", what, " failed (status ", status, "): ", b2pc_last_error());
}
inline Tensor workspace(size_t bytes, const Tensor& like) {
  return at::empty({(int64_t)(bytes > 0 ? bytes : 1)}, like.options().dtype(at::kByte));
}
inline void need_cuda(const Tensor& t) { TORCH_CHECK(t.is_cuda(), "pointcept_b200 operators run on CUDA tensors only; there is no CPU fallback"); }
// every entry point pins the device of its tensors: the launch goes to the GPU that owns the data and to that GPU's current stream
#define B2PC_GUARD(t) need_cuda(t); const c10::cuda::CUDAGuard device_guard__((t).device())
inline const void* optr(const Tensor& t) { return t.defined() ? t.data_ptr() : nullptr; }
inline const float* fptr(const Tensor& t) { return t.defined() ? t.data_ptr<float>() : nullptr; }
inline Tensor f32c(const c10::optional<Tensor>& t) {
  if (!t.has_value() || !t->defined()) return Tensor();
  Tensor v = t->detach();
  if (v.scalar_type() != at::kFloat || !v.is_contiguous()) v = v.to(at::kFloat).contiguous();
  return v;
}

// ---- LayerNorm ----------------------------------------------------------------------------------------------------------
struct LayerNormFn : public torch::autograd::Function<LayerNormFn> {
  static Tensor forward(AutogradContext* ctx, Tensor x, Tensor weight, c10::optional<Tensor> bias, double eps, int64_t out_code) {
    B2PC_GUARD(x);
    x = x.contiguous();
    const int64_t n = x.size(0), c = x.size(1);
    Tensor w = weight.detach();
    if (w.scalar_type() != at::kFloat || !w.is_contiguous()) w = w.to(at::kFloat).contiguous();
    Tensor b;
    if (bias.has_value() && bias->defined()) {
      b = bias->detach();
      if (b.scalar_type() != at::kFloat || !b.is_contiguous()) b = b.to(at::kFloat).contiguous();
    }
    Tensor y = at::empty({n, c}, x.options().dtype(st((int)out_code)));
    Tensor mean = at::empty({n}, x.options().dtype(at::kFloat)), rstd = at::empty({n}, x.options().dtype(at::kFloat));
    check(b2pc_layer_norm_fwd(x.data_ptr(), dt(x), w.data_ptr<float>(), b.defined() ? b.data_ptr<float>() : nullptr, n, (int)c,
                              (float)eps, y.data_ptr(), (int)out_code, mean.data_ptr<float>(), rstd.data_ptr<float>(), cur_stream()),
          "layer_norm_fwd");
    ctx->save_for_backward({x, w, mean, rstd});
    ctx->saved_data["has_bias"] = b.defined();
    ctx->saved_data["pdtype"] = (int64_t)weight.scalar_type();
    return y;
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    auto saved = ctx->get_saved_variables();
    Tensor x = saved[0], w = saved[1], mean = saved[2], rstd = saved[3];
    B2PC_GUARD(x);
    Tensor dy = grads[0].contiguous();
    const int64_t n = x.size(0), c = x.size(1);
    const bool has_bias = ctx->saved_data["has_bias"].toBool();
    Tensor dx = at::empty_like(x);
    Tensor dg = at::empty({c}, x.options().dtype(at::kFloat));
    Tensor db = has_bias ? at::empty({c}, x.options().dtype(at::kFloat)) : Tensor();
    Tensor ws = workspace(b2pc_layer_norm_bwd_workspace_bytes(n, (int)c), x);
    check(b2pc_layer_norm_bwd(dy.data_ptr(), dt(dy), x.data_ptr(), dt(x), w.data_ptr<float>(), mean.data_ptr<float>(),
                              rstd.data_ptr<float>(), n, (int)c, dx.data_ptr(), dg.data_ptr<float>(),
                              has_bias ? db.data_ptr<float>() : nullptr, ws.data_ptr(), (size_t)ws.numel(), cur_stream()),
          "layer_norm_bwd");
    const auto pd = (at::ScalarType)ctx->saved_data["pdtype"].toInt();
    if (pd != at::kFloat) { dg = dg.to(pd); if (has_bias) db = db.to(pd); }
    return {dx, dg, db, Tensor(), Tensor()};
  }
};

// ---- serialized gathers ----------------------------------------------------------------------------------------------------
struct SerializedGatherFn : public torch::autograd::Function<SerializedGatherFn> {
  static Tensor forward(AutogradContext* ctx, Tensor x, Tensor order_pad, Tensor primary_pos, Tensor dup_slots, Tensor dup_points) {
    ctx->save_for_backward({primary_pos, dup_slots, dup_points});
    return x.index_select(0, order_pad);
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    auto s = ctx->get_saved_variables();
    Tensor dy = grads[0];
    Tensor dx = dy.index_select(0, s[0]);
    if (s[1].numel() > 0) dx.index_add_(0, s[2], dy.index_select(0, s[1]));
    return {dx, Tensor(), Tensor(), Tensor(), Tensor()};
  }
};
struct SerializedScatterBackFn : public torch::autograd::Function<SerializedScatterBackFn> {
  static Tensor forward(AutogradContext* ctx, Tensor x_pad, Tensor primary_pos) {
    ctx->save_for_backward({primary_pos});
    ctx->saved_data["t_pad"] = x_pad.size(0);
    return x_pad.index_select(0, primary_pos);
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    auto s = ctx->get_saved_variables();
    Tensor dy = grads[0];
    auto sizes = dy.sizes().vec();
    sizes[0] = ctx->saved_data["t_pad"].toInt();
    Tensor dx = at::zeros(sizes, dy.options());
    dx.index_copy_(0, s[0], dy);
    return {dx, Tensor()};
  }
};

// ---- patch attention ----------------------------------------------------------------------------------------------------------
struct PatchAttentionFn : public torch::autograd::Function<PatchAttentionFn> {
  static Tensor forward(AutogradContext* ctx, Tensor qkv, Tensor cu, int64_t max_seqlen, double scale, int64_t impl) {
    B2PC_GUARD(qkv);
    TORCH_CHECK(qkv.scalar_type() == at::kHalf || qkv.scalar_type() == at::kBFloat16, "patch attention takes fp16 or bf16 qkv");
    qkv = qkv.contiguous();
    if (cu.scalar_type() != at::kInt) cu = cu.to(at::kInt);
    cu = cu.contiguous();
    const int64_t T = qkv.size(0), H = qkv.size(2), D = qkv.size(3);
    Tensor out = at::empty({T, H, D}, qkv.options());
    Tensor lse = at::empty({H, T}, qkv.options().dtype(at::kFloat));
    check(b2pc_patch_attn_fwd(qkv.data_ptr(), dt(qkv), cu.data_ptr<int32_t>(), (int)cu.numel() - 1, (int)max_seqlen, T, (int)H, (int)D,
                              (float)scale, out.data_ptr(), lse.data_ptr<float>(), (int)impl, cur_stream()),
          "patch_attn_fwd");
    ctx->save_for_backward({qkv, out, lse, cu});
    ctx->saved_data["max_seqlen"] = max_seqlen;
    ctx->saved_data["scale"] = scale;
    ctx->saved_data["impl"] = impl;
    return out;
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    auto s = ctx->get_saved_variables();
    Tensor qkv = s[0], out = s[1], lse = s[2], cu = s[3];
    B2PC_GUARD(qkv);
    Tensor dout = grads[0].contiguous();
    const int64_t T = qkv.size(0), H = qkv.size(2), D = qkv.size(3);
    Tensor dqkv = at::empty_like(qkv);
    Tensor ws = workspace(b2pc_patch_attn_bwd_workspace_bytes(T, (int)H, (int)D), qkv);
    check(b2pc_patch_attn_bwd(dout.data_ptr(), qkv.data_ptr(), out.data_ptr(), lse.data_ptr<float>(), dt(qkv), cu.data_ptr<int32_t>(),
                              (int)cu.numel() - 1, (int)ctx->saved_data["max_seqlen"].toInt(), T, (int)H, (int)D,
                              (float)ctx->saved_data["scale"].toDouble(), dqkv.data_ptr(), ws.data_ptr(), (size_t)ws.numel(),
                              (int)ctx->saved_data["impl"].toInt(), cur_stream()),
          "patch_attn_bwd");
    return {dqkv, Tensor(), Tensor(), Tensor(), Tensor()};
  }
};

// ---- serialized attention: [order] gather -> patch attention -> [inverse] gather as ONE operator (point rows in, point rows out) ----------
struct SerializedAttentionFn : public torch::autograd::Function<SerializedAttentionFn> {
  static Tensor forward(AutogradContext* ctx, Tensor qkv, Tensor gidx, Tensor sidx, Tensor dup_point, Tensor cu, int64_t max_seqlen,
                        int64_t heads, double scale) {
    B2PC_GUARD(qkv);
    TORCH_CHECK(qkv.scalar_type() == at::kHalf || qkv.scalar_type() == at::kBFloat16, "serialized attention takes fp16 or bf16 qkv");
    qkv = qkv.contiguous();
    const int64_t n = qkv.size(0), c3 = qkv.size(1), H = heads, D = c3 / (3 * H), t_pad = gidx.size(0);
    TORCH_CHECK(c3 == 3 * H * D && gidx.scalar_type() == at::kInt && sidx.scalar_type() == at::kInt && cu.scalar_type() == at::kInt,
                "serialized attention: qkv [N, 3*H*D], int32 index tables");
    Tensor out = at::empty({n, H * D}, qkv.options());
    Tensor lse = at::empty({H, t_pad}, qkv.options().dtype(at::kFloat));
    check(b2pc_serialized_attn_fwd(qkv.data_ptr(), dt(qkv), gidx.data_ptr<int32_t>(), sidx.data_ptr<int32_t>(), cu.data_ptr<int32_t>(),
                                   (int)cu.numel() - 1, (int)max_seqlen, t_pad, (int)H, (int)D, (float)scale, out.data_ptr(),
                                   lse.data_ptr<float>(), cur_stream()),
          "serialized_attn_fwd");
    ctx->save_for_backward({qkv, out, lse, gidx, sidx, dup_point, cu});
    ctx->saved_data["max_seqlen"] = max_seqlen;
    ctx->saved_data["heads"] = heads;
    ctx->saved_data["scale"] = scale;
    return out;
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    auto s = ctx->get_saved_variables();
    Tensor qkv = s[0], out = s[1], lse = s[2], gidx = s[3], sidx = s[4], dup_point = s[5], cu = s[6];
    B2PC_GUARD(qkv);
    Tensor dout = grads[0].contiguous();
    if (dout.scalar_type() != qkv.scalar_type()) dout = dout.to(qkv.scalar_type());
    const int64_t H = ctx->saved_data["heads"].toInt(), D = qkv.size(1) / (3 * H), t_pad = gidx.size(0), n_dup = dup_point.numel();
    Tensor dqkv = at::empty_like(qkv);
    Tensor ws = workspace(b2pc_serialized_attn_bwd_workspace_bytes(t_pad, (int)H, (int)D, n_dup), qkv);
    check(b2pc_serialized_attn_bwd(dout.data_ptr(), qkv.data_ptr(), out.data_ptr(), lse.data_ptr<float>(), dt(qkv), gidx.data_ptr<int32_t>(),
                                   sidx.data_ptr<int32_t>(), n_dup ? dup_point.data_ptr<int32_t>() : nullptr, n_dup, cu.data_ptr<int32_t>(),
                                   (int)cu.numel() - 1, (int)ctx->saved_data["max_seqlen"].toInt(), t_pad, (int)H, (int)D,
                                   (float)ctx->saved_data["scale"].toDouble(), dqkv.data_ptr(), ws.data_ptr(), (size_t)ws.numel(), cur_stream()),
          "serialized_attn_bwd");
    return {dqkv, Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
  }
};

// ---- sparse convolution -----------------------------------------------------------------------------------------------------------
Tensor gather_gemm(const Tensor& feat, const Tensor& w, const Tensor& bias, const Tensor& pair, int64_t n_out, int c_in, int c_out,
                   int kv, bool transpose_w, bool flip, int impl) {
  Tensor out = at::empty({n_out, c_out}, feat.options());
  const size_t wsb = b2pc_spconv_gather_gemm_workspace_bytes(n_out, c_in, c_out, kv);
  Tensor ws = wsb ? workspace(wsb, feat) : Tensor();
  check(b2pc_spconv_gather_gemm(feat.data_ptr(), w.data_ptr(), bias.defined() ? bias.data_ptr() : nullptr, pair.data_ptr<int32_t>(),
                                pair.size(1), feat.size(0), n_out, c_in, c_out, kv, transpose_w, flip, dt(feat), out.data_ptr(),
                                wsb ? ws.data_ptr() : nullptr, wsb, impl, cur_stream()),
        "spconv_gather_gemm");
  return out;
}

struct SparseConvFn : public torch::autograd::Function<SparseConvFn> {
  // w16 / b16: optional half-precision shadows of weight / bias kept fresh by the cast plan (no per-call cast kernels)
  static Tensor forward(AutogradContext* ctx, Tensor feat, Tensor weight, c10::optional<Tensor> bias, Tensor table_fwd, Tensor table_bwd,
                        bool flip_bwd, int64_t impl, c10::optional<Tensor> w16, c10::optional<Tensor> b16) {
    B2PC_GUARD(feat);
    feat = feat.contiguous();
    const int c_out = (int)weight.size(0), kv = (int)weight.size(1), c_in = (int)weight.size(2);
    TORCH_CHECK(feat.size(1) == c_in, "sparse conv: feature width ", feat.size(1), " != weight input channels ", c_in);
    Tensor w;
    if (w16.has_value() && w16->defined() && w16->scalar_type() == feat.scalar_type()) w = w16->view({c_out, kv, c_in});
    else w = weight.detach().to(feat.scalar_type()).contiguous();
    Tensor b;
    const bool has_bias = bias.has_value() && bias->defined();
    if (has_bias) {
      if (b16.has_value() && b16->defined() && b16->scalar_type() == feat.scalar_type()) b = *b16;
      else b = bias->detach().to(feat.scalar_type()).contiguous();
    }
    Tensor out = gather_gemm(feat, w, b, table_fwd, table_fwd.size(1), c_in, c_out, kv, false, false, (int)impl);
    ctx->save_for_backward({feat, w, table_fwd, table_bwd});
    ctx->saved_data["flip_bwd"] = flip_bwd;
    ctx->saved_data["has_bias"] = has_bias;
    ctx->saved_data["wdtype"] = (int64_t)weight.scalar_type();
    ctx->saved_data["impl"] = impl;
    return out;
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    auto s = ctx->get_saved_variables();
    Tensor feat = s[0], w = s[1], table_fwd = s[2], table_bwd = s[3];
    B2PC_GUARD(feat);
    const int c_out = (int)w.size(0), kv = (int)w.size(1), c_in = (int)w.size(2);
    const int impl = (int)ctx->saved_data["impl"].toInt();
    Tensor dout = grads[0].contiguous();
    if (dout.scalar_type() != feat.scalar_type()) dout = dout.to(feat.scalar_type());
    Tensor dfeat, dweight, dbias;
    if (ctx->needs_input_grad(0))
      dfeat = gather_gemm(dout, w, Tensor(), table_bwd, feat.size(0), c_out, c_in, kv, true, ctx->saved_data["flip_bwd"].toBool(), impl);
    if (ctx->needs_input_grad(1)) {
      dweight = at::empty({c_out, kv, c_in}, feat.options().dtype(at::kFloat));
      const int64_t n_out = table_fwd.size(1);
      Tensor ws = workspace(b2pc_spconv_bwd_weight_workspace_bytes(n_out, c_in, c_out, kv), feat);
      check(b2pc_spconv_bwd_weight(feat.data_ptr(), dout.data_ptr(), table_fwd.data_ptr<int32_t>(), table_fwd.size(1), feat.size(0), n_out,
                                   c_in, c_out, kv, dt(feat), dweight.data_ptr<float>(), ws.data_ptr(), (size_t)ws.numel(), impl,
                                   cur_stream()),
            "spconv_bwd_weight");
      const auto wd = (at::ScalarType)ctx->saved_data["wdtype"].toInt();
      if (wd != at::kFloat) dweight = dweight.to(wd);
    }
    if (ctx->saved_data["has_bias"].toBool() && ctx->needs_input_grad(2)) {
      if (c_out % 4 == 0 && dout.size(0) > 0) {
        dbias = at::empty({c_out}, dout.options().dtype(at::kFloat));
        Tensor ws = workspace(b2pc_colsum_workspace_bytes(dout.size(0), c_out), dout);
        check(b2pc_colsum(dout.data_ptr(), dt(dout), dout.size(0), c_out, dbias.data_ptr<float>(), ws.data_ptr(), (size_t)ws.numel(), cur_stream()),
              "colsum");
      } else {
        dbias = dout.to(at::kFloat).sum(0);
      }
    }
    return {dfeat, dweight, dbias, Tensor(), Tensor(), Tensor(), Tensor(), Tensor(), Tensor()};
  }
};

// ---- pooling ------------------------------------------------------------------------------------------------------------------------
struct SegmentMaxFn : public torch::autograd::Function<SegmentMaxFn> {
  static Tensor forward(AutogradContext* ctx, Tensor x, Tensor order, Tensor seg_start, Tensor seg_len) {
    B2PC_GUARD(x);
    x = x.contiguous();
    const int64_t n = x.size(0), c = x.size(1), m = seg_start.size(0);
    Tensor out = at::empty({m, c}, x.options());
    Tensor arg = at::empty({m, c}, x.options().dtype(at::kInt));
    check(b2pc_segment_max_fwd(x.data_ptr(), dt(x), order.data_ptr<int64_t>(), seg_start.data_ptr<int64_t>(), seg_len.data_ptr<int64_t>(), m,
                               (int)c, out.data_ptr(), arg.data_ptr<int32_t>(), cur_stream()),
          "segment_max_fwd");
    ctx->save_for_backward({arg});
    ctx->saved_data["n"] = n;
    return out;
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    Tensor arg = ctx->get_saved_variables()[0];
    B2PC_GUARD(arg);
    Tensor dout = grads[0].contiguous();
    const int64_t n = ctx->saved_data["n"].toInt(), m = arg.size(0), c = arg.size(1);
    Tensor dx = at::empty({n, c}, dout.options());
    check(b2pc_segment_max_bwd(dout.data_ptr(), dt(dout), arg.data_ptr<int32_t>(), m, (int)c, n, dx.data_ptr(), cur_stream()),
          "segment_max_bwd");
    return {dx, Tensor(), Tensor(), Tensor()};
  }
};

struct UnpoolAddFn : public torch::autograd::Function<UnpoolAddFn> {
  static Tensor forward(AutogradContext* ctx, Tensor parent, Tensor child, Tensor cluster, Tensor order, Tensor seg_len) {
    ctx->save_for_backward({order, seg_len});
    ctx->saved_data["pd"] = (int64_t)parent.scalar_type();
    ctx->saved_data["cd"] = (int64_t)child.scalar_type();
    return parent + child.index_select(0, cluster);
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    auto s = ctx->get_saved_variables();
    Tensor dy = grads[0];
    Tensor dchild = at::segment_reduce(dy.index_select(0, s[0]), "sum", s[1], c10::nullopt, c10::nullopt, 0, true, c10::nullopt);
    return {dy.to((at::ScalarType)ctx->saved_data["pd"].toInt()), dchild.to((at::ScalarType)ctx->saved_data["cd"].toInt()), Tensor(),
            Tensor(), Tensor()};
  }
};

// ---- Linear with the fused fp32 bias-gradient reduction --------------------------------------------------------------------------------
struct LinearFn : public torch::autograd::Function<LinearFn> {
  static Tensor forward(AutogradContext* ctx, Tensor x, Tensor weight, c10::optional<Tensor> bias, int64_t ccode, c10::optional<Tensor> w16,
                        c10::optional<Tensor> b16) {
    B2PC_GUARD(x);
    const auto cd = st((int)ccode);
    Tensor xc = x.scalar_type() == cd ? x : x.to(cd);
    Tensor wc;
    if (w16.has_value() && w16->defined() && w16->scalar_type() == cd) wc = *w16;
    else wc = weight.scalar_type() == cd ? weight.detach() : weight.detach().to(cd);
    const bool has_bias = bias.has_value() && bias->defined();
    Tensor bc;
    if (has_bias) {
      if (b16.has_value() && b16->defined() && b16->scalar_type() == cd) bc = *b16;
      else bc = bias->scalar_type() == cd ? bias->detach() : bias->detach().to(cd);
    }
    Tensor y;
    {
      at::AutoDispatchBelowADInplaceOrView guard;
      y = has_bias ? at::addmm(bc, xc, wc.t()) : at::mm(xc, wc.t());
    }
    ctx->save_for_backward({xc, wc});
    ctx->saved_data["xd"] = (int64_t)x.scalar_type();
    ctx->saved_data["wd"] = (int64_t)weight.scalar_type();
    ctx->saved_data["bd"] = has_bias ? (int64_t)bias->scalar_type() : (int64_t)-1;
    return y;
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    auto s = ctx->get_saved_variables();
    Tensor xc = s[0], wc = s[1];
    B2PC_GUARD(xc);
    Tensor dy = grads[0].contiguous();
    Tensor dx, dw, db;
    const auto xd = (at::ScalarType)ctx->saved_data["xd"].toInt(), wd = (at::ScalarType)ctx->saved_data["wd"].toInt();
    const int64_t bd = ctx->saved_data["bd"].toInt();
    if (ctx->needs_input_grad(0)) { dx = at::mm(dy, wc); if (dx.scalar_type() != xd) dx = dx.to(xd); }
    if (ctx->needs_input_grad(1)) {
      // half-precision operands, fp32 accumulate AND fp32 result straight from the GEMM (no bf16 rounding of dW, no cast kernel)
      static bool mm_dtype_ok = true;
      if (wd == at::kFloat && dy.scalar_type() != at::kFloat && mm_dtype_ok) {
        try { dw = at::mm(dy.t(), xc, at::kFloat); } catch (const c10::Error&) { mm_dtype_ok = false; }
      }
      if (!dw.defined()) { dw = at::mm(dy.t(), xc); if (dw.scalar_type() != wd) dw = dw.to(wd); }
    }
    if (bd >= 0 && ctx->needs_input_grad(2)) {
      const int64_t n = dy.size(0), c = dy.size(1);
      db = at::empty({c}, dy.options().dtype(at::kFloat));
      Tensor ws = workspace(b2pc_colsum_workspace_bytes(n, (int)c), dy);
      check(b2pc_colsum(dy.data_ptr(), dt(dy), n, (int)c, db.data_ptr<float>(), ws.data_ptr(), (size_t)ws.numel(), cur_stream()), "colsum");
      if ((at::ScalarType)bd != at::kFloat) db = db.to((at::ScalarType)bd);
    }
    return {dx, dw, db, Tensor(), Tensor(), Tensor()};
  }
};

// ---- fused residual glue (csrc/fused.cuh): [LayerNorm_a] -> [DropPath scale] -> + shortcut -> [half copy] -> [LayerNorm_b] ---------------
struct FusedResidualFn : public torch::autograd::Function<FusedResidualFn> {
  // returns {r} + ({r16} if emit_half) + ({y} if LayerNorm_b); flags are recoverable from the argument list
  static variable_list forward(AutogradContext* ctx, Tensor shortcut, Tensor x, c10::optional<Tensor> u, double keep,
                               c10::optional<Tensor> ga, c10::optional<Tensor> ba, double eps_a, c10::optional<Tensor> gb,
                               c10::optional<Tensor> bb, double eps_b, bool emit_half, c10::optional<Tensor> x_bias) {
    // x_bias: the bias parameter of the Linear that produced x (already added there); it only routes its gradient -- the
    // column sums of dx -- through this node, so that Linear needs no separate bias-gradient reduction
    B2PC_GUARD(x);
    TORCH_CHECK(shortcut.scalar_type() == at::kFloat, "fused_residual: the residual stream is fp32");
    shortcut = shortcut.contiguous();
    x = x.contiguous();
    const int64_t n = x.size(0), c = x.size(1);
    Tensor gaf = f32c(ga), baf = f32c(ba), gbf = f32c(gb), bbf = f32c(bb);
    Tensor uf = (u.has_value() && u->defined()) ? u->contiguous() : Tensor();
    Tensor r = at::empty({n, c}, x.options().dtype(at::kFloat));
    Tensor r16 = emit_half ? at::empty({n, c}, x.options()) : Tensor();
    Tensor y = gbf.defined() ? at::empty({n, c}, x.options()) : Tensor();
    Tensor sa = gaf.defined() ? at::empty({2, n}, x.options().dtype(at::kFloat)) : Tensor();
    Tensor sb = gbf.defined() ? at::empty({2, n}, x.options().dtype(at::kFloat)) : Tensor();
    check(b2pc_fused_residual_fwd(shortcut.data_ptr<float>(), x.data_ptr(), dt(x), fptr(uf), (float)keep, fptr(gaf), fptr(baf), (float)eps_a,
                                  fptr(gbf), fptr(bbf), (float)eps_b, n, (int)c, r.data_ptr<float>(), r16.defined() ? r16.data_ptr() : nullptr,
                                  y.defined() ? y.data_ptr() : nullptr, sa.defined() ? sa.data_ptr<float>() : nullptr,
                                  sb.defined() ? sb.data_ptr<float>() : nullptr, cur_stream()),
          "fused_residual_fwd");
    ctx->save_for_backward({x, r, uf, gaf, gbf, sa, sb});
    ctx->saved_data["keep"] = keep;
    ctx->saved_data["emit_half"] = emit_half;
    ctx->saved_data["has_ba"] = baf.defined();
    ctx->saved_data["has_bb"] = bbf.defined();
    ctx->saved_data["xb_dtype"] = (x_bias.has_value() && x_bias->defined()) ? (int64_t)x_bias->scalar_type() : (int64_t)-1;
    variable_list outs{r};
    if (emit_half) outs.push_back(r16);
    if (y.defined()) outs.push_back(y);
    return outs;
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    auto s = ctx->get_saved_variables();
    Tensor x = s[0], r = s[1], uf = s[2], gaf = s[3], gbf = s[4], sa = s[5], sb = s[6];
    B2PC_GUARD(x);
    const bool emit_half = ctx->saved_data["emit_half"].toBool();
    const int64_t n = x.size(0), c = x.size(1);
    size_t gi = 0;
    Tensor dr_out = grads[gi++], dr16, dy;
    if (emit_half) dr16 = grads[gi++];
    if (gbf.defined()) dy = grads[gi++];
    if (dr_out.defined()) dr_out = dr_out.to(at::kFloat).contiguous();
    if (dr16.defined()) dr16 = dr16.to(x.scalar_type()).contiguous();
    if (dy.defined()) dy = dy.to(x.scalar_type()).contiguous();
    Tensor d_shortcut = at::empty({n, c}, x.options().dtype(at::kFloat));
    Tensor dx = at::empty_like(x);
    Tensor dga, dba, dgb, dbb;
    if (gaf.defined()) { dga = at::empty({c}, r.options()); dba = at::empty({c}, r.options()); }
    if (gbf.defined()) { dgb = at::empty({c}, r.options()); dbb = at::empty({c}, r.options()); }
    const int64_t xbd = ctx->saved_data["xb_dtype"].toInt();
    // (needs_input_grad() indexes the VARIABLE inputs only -- absent optional tensors shift it -- so the request is keyed on the
    // argument having been passed; the bias is a leaf parameter whenever it is)
    Tensor dxb = xbd >= 0 ? at::empty({c}, r.options()) : Tensor();
    Tensor ws = workspace(b2pc_fused_residual_bwd_workspace_bytes(n, (int)c), x);
    check(b2pc_fused_residual_bwd(fptr(dr_out), optr(dr16), optr(dy), dt(x), r.data_ptr<float>(), x.data_ptr(), fptr(uf),
                                  (float)ctx->saved_data["keep"].toDouble(), fptr(gaf), fptr(gbf), fptr(sa), fptr(sb), n, (int)c,
                                  d_shortcut.data_ptr<float>(), dx.data_ptr(), dga.defined() ? dga.data_ptr<float>() : nullptr,
                                  dba.defined() ? dba.data_ptr<float>() : nullptr, dgb.defined() ? dgb.data_ptr<float>() : nullptr,
                                  dbb.defined() ? dbb.data_ptr<float>() : nullptr, dxb.defined() ? dxb.data_ptr<float>() : nullptr, ws.data_ptr(),
                                  (size_t)ws.numel(), cur_stream()),
          "fused_residual_bwd");
    if (!ctx->saved_data["has_ba"].toBool()) dba = Tensor();
    if (!ctx->saved_data["has_bb"].toBool()) dbb = Tensor();
    if (dxb.defined() && (at::ScalarType)xbd != at::kFloat) dxb = dxb.to((at::ScalarType)xbd);
    return {d_shortcut, dx, Tensor(), Tensor(), dga, dba, Tensor(), dgb, dbb, Tensor(), Tensor(), dxb};
  }
};

// ---- exact GELU --------------------------------------------------------------------------------------------------------------------------
struct GeluFn : public torch::autograd::Function<GeluFn> {
  // x_bias: bias of the Linear that produced x (entered that Linear detached); its gradient = column sums of dx, produced here
  static Tensor forward(AutogradContext* ctx, Tensor x, c10::optional<Tensor> x_bias) {
    B2PC_GUARD(x);
    x = x.contiguous();
    Tensor y = at::empty_like(x);
    check(b2pc_gelu_fwd(x.data_ptr(), dt(x), x.numel(), y.data_ptr(), cur_stream()), "gelu_fwd");
    ctx->save_for_backward({x});
    ctx->saved_data["xb_dtype"] = (x_bias.has_value() && x_bias->defined() && x.dim() == 2) ? (int64_t)x_bias->scalar_type() : (int64_t)-1;
    return y;
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    Tensor x = ctx->get_saved_variables()[0];
    B2PC_GUARD(x);
    Tensor dy = grads[0].contiguous();
    if (dy.scalar_type() != x.scalar_type()) dy = dy.to(x.scalar_type());
    Tensor dx = at::empty_like(x);
    const int64_t xbd = ctx->saved_data["xb_dtype"].toInt();
    Tensor db;
    if (xbd >= 0) {
      const int64_t n = x.size(0), c = x.size(1);
      db = at::empty({c}, x.options().dtype(at::kFloat));
      Tensor ws = workspace(b2pc_gelu_bwd_colsum_workspace_bytes(n, (int)c), x);
      check(b2pc_gelu_bwd_colsum(dy.data_ptr(), x.data_ptr(), dt(x), n, (int)c, dx.data_ptr(), db.data_ptr<float>(), ws.data_ptr(),
                                 (size_t)ws.numel(), cur_stream()),
            "gelu_bwd_colsum");
      if ((at::ScalarType)xbd != at::kFloat) db = db.to((at::ScalarType)xbd);
    } else {
      check(b2pc_gelu_bwd(dy.data_ptr(), x.data_ptr(), dt(x), x.numel(), dx.data_ptr(), cur_stream()), "gelu_bwd");
    }
    return {dx, db};
  }
};

// ---- half-precision parameter shadows: one launch per step instead of one cast kernel per weight ------------------------------------------
struct CastRecord { const float* src; void* dst; long long count; long long first_block; };
// -> (plan tensor on the parameters' device, n_items, total_blocks)
std::tuple<Tensor, int64_t, int64_t> make_cast_plan(std::vector<Tensor> params, std::vector<Tensor> shadows) {
  TORCH_CHECK(params.size() == shadows.size() && !params.empty(), "cast plan: parameter / shadow lists differ");
  std::vector<CastRecord> rec(params.size());
  long long blocks = 0;
  for (size_t i = 0; i < params.size(); ++i) {
    TORCH_CHECK(params[i].is_cuda() && params[i].scalar_type() == at::kFloat && params[i].is_contiguous() && shadows[i].is_contiguous() &&
                    params[i].numel() == shadows[i].numel(), "cast plan: parameters must be contiguous fp32 CUDA tensors with same-size shadows");
    rec[i] = CastRecord{params[i].data_ptr<float>(), shadows[i].data_ptr(), (long long)params[i].numel(), blocks};
    blocks += (params[i].numel() + 2047) / 2048;
  }
  Tensor host = at::empty({(int64_t)(rec.size() * sizeof(CastRecord))}, at::TensorOptions().dtype(at::kByte));
  memcpy(host.data_ptr(), rec.data(), rec.size() * sizeof(CastRecord));
  return std::make_tuple(host.to(params[0].device()), (int64_t)rec.size(), (int64_t)blocks);
}
void run_cast_plan(Tensor plan, int64_t n_items, int64_t total_blocks, int64_t dst_code) {
  B2PC_GUARD(plan);
  check(b2pc_multi_cast(plan.data_ptr(), (int)n_items, total_blocks, (int)dst_code, cur_stream()), "multi_cast");
}

// ---- stochastic depth + residual ----------------------------------------------------------------------------------------------------------
struct DropPathAddFn : public torch::autograd::Function<DropPathAddFn> {
  static Tensor forward(AutogradContext* ctx, Tensor shortcut, Tensor x, double keep) {
    B2PC_GUARD(x);
    shortcut = shortcut.contiguous();
    x = x.contiguous();
    const int64_t n = x.size(0), c = x.size(1);
    Tensor rs = at::empty({n}, x.options().dtype(at::kFloat)).bernoulli_(keep);
    if (keep > 0.0) rs.div_(keep);
    Tensor out = at::empty_like(shortcut);
    check(b2pc_rowscale_add(shortcut.data_ptr(), dt(shortcut), x.data_ptr(), dt(x), rs.data_ptr<float>(), n, (int)c, out.data_ptr(),
                            cur_stream()),
          "rowscale_add");
    ctx->save_for_backward({rs});
    ctx->saved_data["xd"] = (int64_t)x.scalar_type();
    return out;
  }
  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    Tensor rs = ctx->get_saved_variables()[0];
    B2PC_GUARD(rs);
    Tensor dy = grads[0].contiguous();
    const int64_t n = dy.size(0), c = dy.size(1);
    const auto xd = (at::ScalarType)ctx->saved_data["xd"].toInt();
    Tensor dx = at::empty({n, c}, dy.options().dtype(xd));
    check(b2pc_rowscale(dy.data_ptr(), dt(dy), rs.data_ptr<float>(), n, (int)c, dx.data_ptr(), dt(dx), cur_stream()), "rowscale");
    return {dy, dx, Tensor()};
  }
};

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("layer_norm", [](Tensor x, Tensor w, c10::optional<Tensor> b, double eps, int64_t out_code) {
    return LayerNormFn::apply(x, w, b, eps, out_code);
  });
  m.def("serialized_gather", [](Tensor x, Tensor a, Tensor b, Tensor c, Tensor d) { return SerializedGatherFn::apply(x, a, b, c, d); });
  m.def("serialized_scatter_back", [](Tensor x, Tensor p) { return SerializedScatterBackFn::apply(x, p); });
  m.def("patch_attention", [](Tensor qkv, Tensor cu, int64_t max_seqlen, double scale, int64_t impl) {
    return PatchAttentionFn::apply(qkv, cu, max_seqlen, scale, impl);
  });
  m.def("sparse_conv", [](Tensor feat, Tensor weight, c10::optional<Tensor> bias, Tensor tf, Tensor tb, bool flip, int64_t impl,
                          c10::optional<Tensor> w16, c10::optional<Tensor> b16) {
    return SparseConvFn::apply(feat, weight, bias, tf, tb, flip, impl, w16, b16);
  }, py::arg("feat"), py::arg("weight"), py::arg("bias"), py::arg("table_fwd"), py::arg("table_bwd"), py::arg("flip"), py::arg("impl"),
     py::arg("w16") = py::none(), py::arg("b16") = py::none());
  m.def("segment_max", [](Tensor x, Tensor order, Tensor start, Tensor len) { return SegmentMaxFn::apply(x, order, start, len); });
  m.def("unpool_add", [](Tensor p, Tensor c, Tensor cl, Tensor order, Tensor len) { return UnpoolAddFn::apply(p, c, cl, order, len); });
  m.def("linear", [](Tensor x, Tensor w, c10::optional<Tensor> b, int64_t ccode, c10::optional<Tensor> w16, c10::optional<Tensor> b16) {
    return LinearFn::apply(x, w, b, ccode, w16, b16);
  }, py::arg("x"), py::arg("weight"), py::arg("bias"), py::arg("ccode"), py::arg("w16") = py::none(), py::arg("b16") = py::none());
  m.def("fused_residual", [](Tensor shortcut, Tensor x, c10::optional<Tensor> u, double keep, c10::optional<Tensor> ga, c10::optional<Tensor> ba,
                             double eps_a, c10::optional<Tensor> gb, c10::optional<Tensor> bb, double eps_b, bool emit_half,
                             c10::optional<Tensor> x_bias) {
    return FusedResidualFn::apply(shortcut, x, u, keep, ga, ba, eps_a, gb, bb, eps_b, emit_half, x_bias);
  });
  m.def("gelu", [](Tensor x, c10::optional<Tensor> x_bias) { return GeluFn::apply(x, x_bias); }, py::arg("x"), py::arg("x_bias") = py::none());
  m.def("serialized_attention", [](Tensor qkv, Tensor gidx, Tensor sidx, Tensor dup_point, Tensor cu, int64_t max_seqlen, int64_t heads,
                                   double scale) { return SerializedAttentionFn::apply(qkv, gidx, sidx, dup_point, cu, max_seqlen, heads, scale); });
  m.def("make_cast_plan", &make_cast_plan);
  m.def("run_cast_plan", &run_cast_plan);
  m.def("drop_path_add", [](Tensor s, Tensor x, double keep) { return DropPathAddFn::apply(s, x, keep); });
  m.def("launch_count", []() { return (int64_t)b2pc_launch_count(); });
}
