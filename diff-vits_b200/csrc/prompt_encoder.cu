// PromptEncoder (reference: model3.py:382-433) as a plan over the engine's kernels: the speech-prompt encoder whose output
// is the UNet's encoder_hidden_states. Channel-last bf16 [B, Tp, 128] throughout (the layout cross-attention reads), fp32
// result. Launch program for n_layers = 4: 1 layout + 11 LayerNorm + 18 tap-GEMM (tcgen05) + 4 attention launches.
//
//   pre      : masked_fill(pad, 0) -> LayerNorm(100) -> ConvTBC k=1 (Linear 100 -> 128) -> * (1 - pad)      (model.py:72-97, model3.py:420-421)
//   layer x4 : x += MHA(LN1(x), key padding) ; x *= (1 - pad) ; x += FFN(LN2(x)) ; x *= (1 - pad)           (operations.py:798-821)
//              MHA = fairseq/torch multi_head_attention_forward: in_proj [3C, C] without bias, 8 heads, q scaled by d^-1/2,
//              padded keys get zero weight, out_proj without bias (operations.py:304-470, 792-794)
//              FFN = sum_i Linear_i(x shifted by i - 4) * 9^-1/2 -> ReLU -> Linear(4C -> C) with the reference's quirk that
//              tap 0 reads the UNSHIFTED rows (`padded[i:T+i] if i else x`, operations.py:678-684); rows outside [0, Tp) are
//              zero, rows of the padding region hold LayerNorm(0) = beta, exactly as in the reference. The 9^-1/2 is folded
//              into the packed tap weights and bias (ReLU is positively homogeneous).
//   out      : LayerNorm(128) -> ConvTBC k=1 -> * (1 - pad) -> LayerNorm(128) -> * (1 - pad)                  (model3.py:426-431)
#include "engine.h"

#include <cstring>

namespace dv {

namespace {
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
}  // namespace

int PromptPlan::build(std::string& err) {
  const dv_prompt_cfg& c = cfg;
  if (c.hidden != 128) { err = "prompt encoder: only hidden = 128 is built"; return DV_ERR_UNSUPPORTED; }
  if (c.in_channels < 1 || c.in_channels > 128) { err = "prompt encoder: in_channels must be in [1, 128]"; return DV_ERR_UNSUPPORTED; }
  if (c.num_heads != 8) { err = "prompt encoder: only 8 heads (head dim 16) are built"; return DV_ERR_UNSUPPORTED; }
  if (c.ffn_kernel != 9 || c.ffn_mult != 4) { err = "prompt encoder: only the 9-tap, 4x conv feed-forward is built"; return DV_ERR_UNSUPPORTED; }
  if (c.n_layers < 1 || c.n_layers > 16 || B < 1 || Tp < 1) { err = "prompt encoder: bad n_layers / B / Tp"; return DV_ERR_INVALID; }
  const int C = 128, F = c.ffn_mult * C;
  size_t cur = 0;
  auto alloc = [&](size_t bytes) { size_t o = align_up(cur, 1024); cur = o + bytes; return o; };
  auto slot = [&](const std::string& name, const std::string& recipe, std::vector<std::string> keys, const std::string& dtype,
                  std::vector<int> shape) -> long long {
    ParamSlot s;
    s.name = name; s.recipe = recipe; s.keys = std::move(keys); s.dtype = dtype; s.shape = std::move(shape);
    size_t n = 1;
    for (int d : s.shape) n *= static_cast<size_t>(d);
    s.bytes = n * (dtype == "bf16" ? 2 : 4);
    s.off = align_up(param_bytes, 256);
    param_bytes = s.off + s.bytes;
    slots.push_back(s);
    return static_cast<long long>(s.off);
  };
  const size_t rowsC = static_cast<size_t>(B) * Tp * C * 2;
  off_xin = alloc(rowsC);
  off_len = alloc(static_cast<size_t>(B) * 4);
  const size_t x = alloc(rowsC), n = alloc(rowsC), ao = alloc(rowsC), y = alloc(rowsC);
  const size_t qkv = alloc(rowsC * 3), h = alloc(rowsC * c.ffn_mult);
  off_out32 = alloc(static_cast<size_t>(B) * Tp * C * 4);

  auto ln = [&](const std::string& key, size_t src, size_t dst, int c_real, bool mask_in, bool mask_out, bool f32) {
    Step s;
    s.kind = 1; s.tag = key; s.x = src; s.out = dst;
    s.gamma = slot(key + ".weight", "vec", {key + ".weight"}, "f32", {c_real});
    s.beta = slot(key + ".bias", "vec", {key + ".bias"}, "f32", {c_real});
    memset(&s.ln, 0, sizeof s.ln);
    s.ln.c_real = c_real; s.ln.mask_in = mask_in; s.ln.mask_out = mask_out; s.ln.rows_per_batch = Tp;
    s.ln_f32 = f32;
    steps.push_back(s);
  };
  auto gemm = [&](const std::string& tag, size_t a, int a_c, const int* row_offs, int ntaps, const std::string& recipe,
                  std::vector<std::string> wkeys, const std::string& bkey, const std::string& brecipe, int N, size_t out,
                  bool res, bool relu, bool mask) {
    Step s;
    s.kind = 0; s.tag = tag; s.a_off = a; s.a_c = a_c; s.nseg = ntaps; s.n = N; s.out = out;
    for (int i = 0; i < ntaps; ++i) s.row_off[i] = row_offs[i];
    s.k_tot = ntaps * a_c;
    s.w_off = slot(tag + ".W", recipe, std::move(wkeys), "bf16", {N, s.k_tot});
    if (!bkey.empty()) s.b_off = slot(tag + ".b", brecipe, {bkey}, "f32", {N});
    s.has_res = res; s.res_off = out; s.relu = relu; s.mask_rows = mask;
    steps.push_back(s);
  };
  const int zero[1] = {0};
  // operations.py:678-684: tap 0 reads the unshifted rows, taps 1..8 rows t + i - 4
  const int taps[9] = {0, -3, -2, -1, 0, 1, 2, 3, 4};

  ln("pre.layer_norm", off_xin, n, c.in_channels, true, false, false);
  gemm("pre.conv", n, C, zero, 1, "convtbc1_pad", {"pre.conv.weight"}, "pre.conv.bias", "vec", C, x, false, false, true);
  for (int l = 0; l < c.n_layers; ++l) {
    const std::string p = "layers." + std::to_string(l) + ".op.";
    ln(p + "layer_norm1", x, n, C, false, false, false);
    gemm(p + "self_attn.in_proj", n, C, zero, 1, "linear", {p + "self_attn.in_proj_weight"}, "", "", 3 * C, qkv, false, false, false);
    {
      Step s;
      s.kind = 2; s.tag = p + "self_attn";
      s.q = qkv; s.k = qkv + static_cast<size_t>(C) * 2; s.v = qkv + static_cast<size_t>(C) * 4; s.o = ao;
      steps.push_back(s);
    }
    gemm(p + "self_attn.out_proj", ao, C, zero, 1, "linear", {p + "self_attn.out_proj.weight"}, "", "", C, x, true, false, true);
    ln(p + "layer_norm2", x, n, C, false, false, false);
    std::vector<std::string> wk;
    for (int i = 0; i < 9; ++i) wk.push_back(p + "ffn.ffn_1." + std::to_string(i) + ".weight");
    gemm(p + "ffn.ffn_1", n, C, taps, 9, "ffn_taps", wk, p + "ffn.ffn_1.0.bias", "vec_ffn_scale", F, h, false, true, false);
    gemm(p + "ffn.ffn_2", h, F, zero, 1, "linear", {p + "ffn.ffn_2.weight"}, p + "ffn.ffn_2.bias", "vec", C, x, true, false, true);
  }
  ln("out_proj.layer_norm", x, n, C, true, false, false);
  gemm("out_proj.conv", n, C, zero, 1, "convtbc1", {"out_proj.conv.weight"}, "out_proj.conv.bias", "vec", C, y, false, false, true);
  ln("layer_norm", y, n, C, true, true, true);
  ws_bytes = align_up(cur, 1024);
  param_bytes = align_up(param_bytes, 256);
  manifest = manifest_json(slots);
  return DV_OK;
}

int PromptPlan::bind(void* workspace, const void* prm, cudaStream_t st, std::string& err) {
  if (!workspace || !prm) { err = "null workspace/params"; return DV_ERR_INVALID; }
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(prm) & 255)) {
    err = "workspace and params must be 256-byte aligned";
    return DV_ERR_INVALID;
  }
  ws = static_cast<uint8_t*>(workspace);
  params = static_cast<const uint8_t*>(prm);
  cudaError_t ce = gemm_tc_init();
  if (ce == cudaSuccess) ce = attention_tc_init();
  if (ce == cudaSuccess) ce = cudaMemsetAsync(ws, 0, ws_bytes, st);
  if (ce != cudaSuccess) { err = cudaGetErrorString(ce); return DV_ERR_CUDA; }
  const int C = 128;
  auto pf = [&](long long off) -> const float* { return off < 0 ? nullptr : reinterpret_cast<const float*>(params + off); };
  for (Step& s : steps) {
    if (s.kind == 0) {
      GemmParams& p = s.gp;
      memset(&p, 0, sizeof p);
      for (int i = 0; i < GEMM_MAX_SRC; ++i) {
        p.map_rows[i] = 128;
        if (!make_tmap_3d(&p.mapA[i], ws + s.a_off, s.a_c, s.a_c, Tp, B, static_cast<size_t>(Tp) * s.a_c, err)) return DV_ERR_CUDA;
      }
      const int n_pad = (s.n + 127) / 128 * 128;
      if (!make_tmap_2d(&p.mapW, params + s.w_off, s.k_tot, n_pad, 64, err)) return DV_ERR_CUDA;
      p.mapW2 = p.mapW;
      p.nseg = s.nseg;
      for (int i = 0; i < s.nseg; ++i) {
        GemmSeg& g = p.seg[i];
        g.src = 0; g.row_off = static_cast<int16_t>(s.row_off[i]); g.c0 = 0; g.npanels = static_cast<int16_t>(s.a_c / 64);
        g.wsel = 0; g.wk0 = static_cast<int16_t>(i * (s.a_c / 64)); g.ntap = 1; g.tap_stride = g.npanels; g.xf = 0; g.xf_c0 = 0;
      }
      p.B = B; p.L = Tp; p.N = s.n; p.n_pad = n_pad; p.tiles_per_batch = (Tp + 127) / 128;
      p.out_mode = GEMM_OUT_BF16;
      p.bias = pf(s.b_off);
      p.res = s.has_res ? ws + s.res_off : nullptr; p.ld_res = s.has_res ? s.n : 0;
      p.out = ws + s.out; p.ld_out = s.n;
      p.relu = s.relu ? 1 : 0;
      p.row_len = s.mask_rows ? reinterpret_cast<const int*>(ws + off_len) : nullptr;
    } else if (s.kind == 1) {
      s.ln.row_len = reinterpret_cast<const int*>(ws + off_len);
      s.ln.out_f32 = s.ln_f32 ? reinterpret_cast<float*>(ws + off_out32) : nullptr;
    } else {
      AttnParams& a = s.at;
      memset(&a, 0, sizeof a);
      a.q = ws + s.q; a.k = ws + s.k; a.v = ws + s.v; a.o = ws + s.o;
      a.ld_q = a.ld_k = a.ld_v = 3 * C; a.ld_o = C;
      a.bs_q = a.bs_k = a.bs_v = static_cast<long long>(Tp) * 3 * C; a.bs_o = static_cast<long long>(Tp) * C;
      a.B = B; a.H = cfg.num_heads; a.Lq = Tp; a.Lk = Tp; a.D = C / cfg.num_heads;
      a.scale = 1.0f / sqrtf(static_cast<float>(a.D));
      a.kv_len = reinterpret_cast<const int*>(ws + off_len);
      const char* ae = nullptr;
      if (!attention_tc_prepare(a, &ae)) { err = std::string("attention_tc_prepare: ") + (ae ? ae : "?"); return DV_ERR_CUDA; }
    }
  }
  bound = true;
  return DV_OK;
}

int PromptPlan::encode(const float* prompt, const int32_t* lengths, float* ehs_out, cudaStream_t st, std::string& err) {
  if (!bound) { err = "prompt plan not bound"; return DV_ERR_STATE; }
  if (!prompt || !lengths || !ehs_out) { err = "null argument"; return DV_ERR_INVALID; }
#define DV_PCK(call)                                                                                  \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e__); return DV_ERR_CUDA; } \
  } while (0)
  DV_PCK(cudaMemcpyAsync(ws + off_len, lengths, static_cast<size_t>(B) * 4, cudaMemcpyDeviceToDevice, st));
  DV_PCK(launch_nct_to_ntc(prompt, ws + off_xin, B, cfg.in_channels, Tp, 128, 0, 1 << 30, 0, st));
  for (Step& s : steps) {
    cudaError_t e = cudaSuccess;
    if (s.kind == 0) e = launch_gemm_tc(s.gp, st);
    else if (s.kind == 1)
      e = launch_layernorm_masked(ws + s.x, ws + s.out, reinterpret_cast<const float*>(params + s.gamma),
                                  reinterpret_cast<const float*>(params + s.beta), B * Tp, 1e-5f, s.ln, st);
    else e = launch_attention(s.at, st);
    if (e != cudaSuccess) { err = "launch " + s.tag + ": " + cudaGetErrorString(e); return DV_ERR_CUDA; }
  }
  DV_PCK(cudaMemcpyAsync(ehs_out, ws + off_out32, static_cast<size_t>(B) * Tp * 128 * 4, cudaMemcpyDeviceToDevice, st));
#undef DV_PCK
  return DV_OK;
}

}  // namespace dv
