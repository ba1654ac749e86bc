// Plan builder + executor for the Diff-VITS denoiser (see engine.h).
#include "engine.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include <cudaTypedefs.h>

namespace dv {

const KernelApi& kernel_api(bool f16) {
  static const KernelApi bf{dv::gemm_tc_init,    dv::chain_tc_init,    dv::attention_tc_init, dv::attention_tc_prepare,
                            dv::launch_gemm_tc,  dv::launch_chain_tc,  dv::launch_attention,  dv::launch_gn_apply,
                            dv::launch_layernorm, dv::launch_nct_to_ntc, dv::launch_f32_to_bf16};
  static const KernelApi hf{dv16::gemm_tc_init,    dv16::chain_tc_init,    dv16::attention_tc_init, dv16::attention_tc_prepare,
                            dv16::launch_gemm_tc,  dv16::launch_chain_tc,  dv16::launch_attention,  dv16::launch_gn_apply,
                            dv16::launch_layernorm, dv16::launch_nct_to_ntc, dv16::launch_f32_to_bf16};
  return f16 ? hf : bf;
}

// --------------------------------------------------------------------------------------------- TMA descriptors
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn(std::string& err) {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
    err = std::string("cuTensorMapEncodeTiled unavailable: ") + cudaGetErrorString(e);
    return nullptr;
  }
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  return fn;
}

bool make_tmap_3d(CUtensorMap* m, const void* base, int view_c, int ld, int L, int B, size_t batch_stride_elems,
                  std::string& err, int box_rows) {
  auto fn = get_encode_fn(err);
  if (!fn) return false;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(view_c), static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(batch_stride_elems) * 2};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(3d c=%d ld=%d L=%d B=%d) failed: %d", view_c, ld, L, B, (int)r);
    err = buf;
    return false;
  }
  return true;
}

bool make_tmap_2d(CUtensorMap* m, const void* base, int k_tot, int rows, int box_rows, std::string& err) {
  auto fn = get_encode_fn(err);
  if (!fn) return false;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(k_tot), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(k_tot) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(2d k=%d rows=%d) failed: %d", k_tot, rows, (int)r);
    err = buf;
    return false;
  }
  return true;
}

// --------------------------------------------------------------------------------------------- builder
namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

struct Builder {
  Plan& P;
  size_t ws_cur = 0, scratch_cur = 0, scratch_max = 0;
  std::vector<std::pair<size_t*, int>> scratch_fix;  // offsets that live in the scratch arena (patched at the end)
  std::vector<std::string> wall_keys, ball_keys;
  int ss_cur = 0;
  int cross_layers = 0;
  std::vector<TRef> kv_bufs;
  TRef ehs_bf16;
  std::string err;

  explicit Builder(Plan& p) : P(p) {}

  size_t ws_alloc(size_t bytes) {
    size_t o = align_up(ws_cur, 1024);
    ws_cur = o + bytes;
    return o;
  }
  // bf16 tensors are followed by their lo plane in split precision (Plan::plane)
  size_t tensor_bytes(int L, int C, bool lo) const { return static_cast<size_t>(P.B) * L * C * 2 * (lo ? 2 : 1); }
  // block outputs: the residual stream (lo plane in both split modes)
  TRef act(int L, int C) {
    TRef t;
    t.L = L; t.C = C; t.lo = P.split;
    t.off = ws_alloc(tensor_bytes(L, C, t.lo));
    return t;
  }
  // scratch tensors: lifetime = one block; offsets are relative to the scratch arena until finalize()
  static constexpr size_t SCRATCH_TAG = size_t(1) << 62;
  // stream = part of the residual stream (the transformer's z); everything else is read by the tensor core / attention only
  TRef tmp(int L, int C, bool stream = false) {
    TRef t;
    t.L = L; t.C = C; t.lo = P.triple || (P.split && stream);
    size_t o = align_up(scratch_cur, 1024);
    scratch_cur = o + tensor_bytes(L, C, t.lo);
    if (scratch_cur > scratch_max) scratch_max = scratch_cur;
    t.off = SCRATCH_TAG | o;
    return t;
  }
  void scratch_reset() { scratch_cur = 0; }
  // Normalised tensors written by the LAST GEMM of one block and read by the FIRST GEMM of the next (fused output GroupNorm
  // across a block boundary) cannot live in the per-block scratch arena (the producer's block still uses it): they
  // alternate between two buffers sized for the largest of them. Buffer k is read at the start of block k and next written
  // at the end of block k + 1.
  static constexpr size_t BOUND_TAG = size_t(1) << 61, BOUND_PAR = size_t(1) << 60;
  size_t bound_max = 0;
  int bound_n = 0;
  TRef boundary(int L, int C) {
    TRef t;
    t.L = L; t.C = C; t.lo = P.triple;           // GEMM operands carry a lo plane in FP32X3 only (as tmp() does)
    const size_t bytes = tensor_bytes(L, C, t.lo);
    if (bytes > bound_max) bound_max = bytes;
    t.off = BOUND_TAG | ((bound_n++ & 1) ? BOUND_PAR : 0);
    return t;
  }
  // Can GroupNorm(x0) be applied by the kernel that produces x0? (single source written by a plain bf16 tap-GEMM with
  // statistics; every tile of a batch item inside one round of that kernel's persistent tile loop)
  bool gn_fusable(const TRef& x0, const TRef& x1) const {
    if (!P.gn_out || x1.valid() || x0.prod < 0 || x0.prod >= static_cast<int>(P.ops.size())) return false;
    const Op& pr = P.ops[x0.prod];
    if (pr.kind != OP_GEMM) return false;
    const GemmDesc& g = pr.g;
    if (g.out_mode != GEMM_OUT_BF16 || g.gnout.on || g.out_bs != 0 || g.trim_col != 0 || g.out_dynamic || g.stats_off < 0 ||
        g.out.off != x0.off || g.n != x0.C)
      return false;
    return ((g.l_out + 127) / 128) * (g.n_pad / 64) <= 128;
  }
  // the tensor a GroupNorm writes: scratch of the current block, or a boundary buffer when the producer of x0 (in the previous
  // block) is going to write it
  TRef gn_target(const TRef& x0, const TRef& x1, int L, int C, bool same_block = false) {
    return (!same_block && gn_fusable(x0, x1)) ? boundary(L, C) : tmp(L, C);
  }

  long long slot(const std::string& name, const std::string& recipe, std::vector<std::string> keys,
                 const std::string& dtype, std::vector<int> shape) {
    ParamSlot s;
    s.name = name; s.recipe = recipe; s.keys = std::move(keys); s.dtype = dtype; s.shape = std::move(shape);
    size_t n = 1;
    for (int d : s.shape) n *= static_cast<size_t>(d);
    if (dtype == "bf16" && P.triple) s.dtype = "bf16x2";    // hi matrix followed by the lo matrix
    if (dtype == "bf16" && P.f16) s.dtype = "f16";
    s.bytes = n * ((s.dtype == "bf16" || s.dtype == "f16") ? 2 : 4);
    s.off = align_up(P.param_bytes, 256);
    P.param_bytes = s.off + s.bytes;
    P.slots.push_back(s);
    return static_cast<long long>(s.off);
  }
  long long vec(const std::string& key, int n) { return slot(key, "vec", {key}, "f32", {n}); }

  // GroupNorm partials of a GEMM output: [B][ceil(L/128)*4 row chunks][N/16 blocks][2] floats
  void attach_stats(GemmDesc& g, TRef& out) {
    const int chunks = ((g.l_out + 127) / 128) * 4;
    g.stats_off = static_cast<long long>(P.stats_bytes);
    g.stats_nblk = g.n / 16;
    P.stats_bytes += align_up(static_cast<size_t>(P.B) * chunks * g.stats_nblk * 2 * sizeof(float), 256);
    out.st_off = g.stats_off; out.st_chunks = chunks; out.st_nblk = g.stats_nblk; out.st_fold = 0;
  }

  // ---- op emitters
  void emit_gn(std::vector<Op>& list, const std::string& tag, TRef x0, TRef x1, const std::string& norm_key, float eps,
               bool silu, long long ss_off, TRef out) {
    const int C = x0.C + (x1.valid() ? x1.C : 0);
    if (x0.st_off < 0 || (x1.valid() && x1.st_off < 0)) { err = "internal: GroupNorm input without statistics: " + tag; return; }
    if (&list == &P.ops && gn_fusable(x0, x1)) {          // applied by the producer's epilogue: no launch of its own
      GnOutDesc& d = P.ops[x0.prod].g.gnout;
      d.on = true; d.out = out; d.eps = eps; d.silu = silu; d.ss = ss_off; d.tag = tag;
      P.ops[x0.prod].tag += " +gn:" + tag;          // shows up in the op table (dv_plan_op_info)
      d.gamma = vec(norm_key + ".weight", C);
      d.beta = vec(norm_key + ".bias", C);
      return;
    }
    Op ap;
    ap.kind = OP_GN_APPLY; ap.tag = tag;
    memset(&ap.gn, 0, sizeof ap.gn);
    ap.gn.c0 = x0.C; ap.gn.c1 = x1.valid() ? x1.C : 0;
    ap.gn.B = P.B; ap.gn.L = x0.L; ap.gn.G = P.cfg.norm_num_groups;
    ap.gn.eps = eps; ap.gn.silu = silu ? 1 : 0;
    ap.gn.nst = x1.valid() ? 2 : 1;
    const TRef* xs[2] = {&x0, &x1};
    for (int i = 0; i < ap.gn.nst; ++i) {
      ap.gn.st[i].chunks = xs[i]->st_chunks; ap.gn.st[i].nblk = xs[i]->st_nblk;
      ap.gn.st[i].fold = xs[i]->st_fold; ap.gn.st[i].cover = xs[i]->C;
      ap.gno.st[i] = xs[i]->st_off;
    }
    ap.gno.x0 = x0.off; ap.gno.x1 = x1.valid() ? x1.off : 0; ap.gno.out = out.off;
    ap.gno.x_lo = x0.lo && (!x1.valid() || x1.lo); ap.gno.out_lo = out.lo;
    ap.gno.gamma = vec(norm_key + ".weight", C);
    ap.gno.beta = vec(norm_key + ".bias", C);
    ap.gno.ss = ss_off;
    list.push_back(ap);
  }
  void emit_ln(std::vector<Op>& list, const std::string& tag, TRef x, const std::string& key, TRef out) {
    Op o;
    o.kind = OP_LN; o.tag = tag;
    o.lno.x = x.off; o.lno.out = out.off;
    o.lno.x_lo = x.lo; o.lno.out_lo = out.lo;
    o.lno.gamma = vec(key + ".weight", x.C);
    o.lno.beta = vec(key + ".bias", x.C);
    o.ln.rows = P.B * x.L; o.ln.C = x.C; o.ln.eps = 1e-5f;
    list.push_back(o);
  }
  Op& emit_gemm(std::vector<Op>& list, const std::string& tag) {
    Op o;
    o.kind = OP_GEMM; o.tag = tag;
    for (int i = 0; i < GEMM_MAX_LSRC; ++i) { o.g.view_l[i] = 0; o.g.view_c[i] = 0; }
    list.push_back(o);
    return list.back();
  }
  static void add_src(GemmDesc& g, TRef t) {
    g.src[g.nsrc] = t; g.view_l[g.nsrc] = t.L; g.view_c[g.nsrc] = t.C; g.nsrc++;
  }
  // ntap = 3: conv k3 with shared rows (row_off = -1). wk0 < 0: the segment's weights follow the previous segment's in K;
  // tap_stride 0: the taps of this segment are `channels` apart (single-source conv).
  static void add_seg(GemmDesc& g, int src, int row_off, int c0, int channels, int ntap = 1, int wk0 = -1, int tap_stride = 0,
                      bool xf = false, int xf_c0 = 0) {
    GemmSeg s;
    memset(&s, 0, sizeof s);
    s.src = static_cast<int16_t>(src); s.row_off = static_cast<int16_t>(row_off);
    s.c0 = static_cast<int16_t>(c0); s.npanels = static_cast<int16_t>(channels / 64);
    s.ntap = static_cast<int16_t>(ntap);
    s.wk0 = static_cast<int16_t>(wk0 >= 0 ? wk0 : g.k_tot / 64);
    s.tap_stride = static_cast<int16_t>(tap_stride > 0 ? tap_stride : channels / 64);
    s.xf = xf ? 1 : 0; s.xf_c0 = static_cast<int16_t>(xf_c0);
    g.seg[g.nseg++] = s;
    g.k_tot += channels * ntap;
  }
  // GroupNorm parameters of a fused-norm GEMM
  void set_xf(GemmDesc& g, TRef x0, TRef x1, const std::string& norm_key, float eps, bool silu, long long ss_off) {
    const int C = x0.C + (x1.valid() ? x1.C : 0);
    if (x0.st_off < 0 || (x1.valid() && x1.st_off < 0)) { err = "internal: GroupNorm input without statistics: " + norm_key; return; }
    g.xf.on = true; g.xf.x0 = x0; g.xf.x1 = x1;
    g.xf.gamma = vec(norm_key + ".weight", C);
    g.xf.beta = vec(norm_key + ".bias", C);
    g.xf.ss = ss_off; g.xf.eps = eps; g.xf.silu = silu;
  }

  // Linear / 1x1 conv: out[L, N] = a[L, K] W^T (+bias) (+res). norm_key != "": `a` is the RAW input of a GroupNorm (eps,
  // no SiLU) that is applied on the way into the tensor core.
  void linear(std::vector<Op>& list, const std::string& tag, TRef a, const std::string& recipe,
              std::vector<std::string> wkeys, const std::string& bias_key, int N, TRef res, TRef& out, int out_mode,
              bool stats = false, const std::string& norm_key = "", float norm_eps = 1e-5f) {
    Op& o = emit_gemm(list, tag);
    GemmDesc& g = o.g;
    add_src(g, a);
    add_seg(g, 0, 0, 0, a.C, 1, -1, 0, !norm_key.empty(), 0);
    if (!norm_key.empty()) set_xf(g, a, TRef(), norm_key, norm_eps, false, -1);
    g.n = N; g.n_pad = round_up(N, 128);
    g.w_off = slot(tag + ".W", recipe, std::move(wkeys), "bf16", {g.n_pad, g.k_tot});
    if (!bias_key.empty())
      g.bias_off = slot(tag + ".b", recipe == "geglu" ? "geglu_vec" : "vec", {bias_key}, "f32", {g.n_pad});
    g.res = res; g.l_out = a.L; g.out_mode = out_mode; g.ld_out = out.C;
    if (stats) attach_stats(g, out);
    if (&list == &P.ops) out.prod = static_cast<int>(list.size()) - 1;
    g.out = out;
  }

  // conv k3 s1 p1 over (a0 | a1) (+ optional 1x1 shortcut over xs fused as extra K panels) (+ optional identity residual).
  // Each 64-channel panel is staged once (136 rows) and read by the three taps through row-shifted descriptors.
  // norm_key != "": (a0 | a1) is the RAW input of GroupNorm(norm_key) + SiLU (+ time scale/shift at ss_off), applied in
  // shared memory on the way into the tensor core; otherwise a0 is the already normalised operand.
  void conv3(std::vector<Op>& list, const std::string& tag, TRef a0, TRef a1, int cin_real, const std::string& wkey,
             const std::string& bkey, int N, const TRef* sc_src, int n_sc, const std::string& sc_key, TRef res,
             TRef& out, int out_mode, bool dynamic_out, const std::string& norm_key = "", long long ss_off = -1) {
    Op& o = emit_gemm(list, tag);
    GemmDesc& g = o.g;
    const bool xf = !norm_key.empty();
    const int cin = a0.C + (a1.valid() ? a1.C : 0);
    const int np = cin / 64;                       // K panels per tap: weights are tap-major over the concatenated channels
    add_src(g, a0);
    if (a1.valid()) add_src(g, a1);
    if (P.gemm2) {
      add_seg(g, 0, -1, 0, a0.C, 3, 0, np, xf, 0);
      if (a1.valid()) add_seg(g, 1, -1, 0, a1.C, 3, a0.C / 64, np, xf, a0.C);
    } else {                                       // one-ring kernel: a segment (= its own A loads) per tap and source
      for (int tap = 0; tap < 3; ++tap) {
        add_seg(g, 0, tap - 1, 0, a0.C, 1, tap * np);
        if (a1.valid()) add_seg(g, 1, tap - 1, 0, a1.C, 1, tap * np + a0.C / 64);
      }
    }
    if (xf) set_xf(g, a0, a1, norm_key, P.cfg.norm_eps, true, ss_off);
    std::vector<std::string> wk = {wkey};
    std::vector<std::string> bk = {bkey};
    for (int i = 0; i < n_sc; ++i) {
      add_src(g, sc_src[i]);
      add_seg(g, g.nsrc - 1, 0, 0, sc_src[i].C);
    }
    if (n_sc) { wk.push_back(sc_key + ".weight"); bk.push_back(sc_key + ".bias"); }
    g.n = N; g.n_pad = round_up(N, 128);
    long long wo = slot(tag + ".W", "conv3", wk, "bf16", {g.n_pad, g.k_tot});
    P.slots.back().shape.push_back(cin_real);  // [n_pad, k_tot, cin_real]: packer pads each tap from cin_real to cin
    P.slots.back().bytes = static_cast<size_t>(g.n_pad) * g.k_tot * 2 * (P.triple ? 2 : 1);
    g.w_off = wo;
    g.bias_off = slot(tag + ".b", "vec_sum", bk, "f32", {g.n_pad});
    g.res = res; g.l_out = a0.L; g.out_mode = out_mode; g.ld_out = out.C; g.out_dynamic = dynamic_out;
    if (out_mode == GEMM_OUT_BF16) attach_stats(g, out);
    if (&list == &P.ops) out.prod = static_cast<int>(list.size()) - 1;
    g.out = out;
  }

  TRef resnet(const std::string& pfx, TRef x0, TRef x1, int cout) {
    scratch_reset();
    const int L = x0.L;
    const int cin = x0.C + (x1.valid() ? x1.C : 0);
    const float eps = P.cfg.norm_eps;
    TRef u = tmp(L, cout);
    if (P.fuse_gn) {
      conv3(P.ops, pfx + ".conv1", x0, x1, cin, pfx + ".conv1.weight", pfx + ".conv1.bias", cout, nullptr, 0, "", TRef(), u,
            GEMM_OUT_BF16, false, pfx + ".norm1", -1);
    } else {
      TRef a = gn_target(x0, x1, L, cin);
      emit_gn(P.ops, pfx + ".norm1", x0, x1, pfx + ".norm1", eps, true, -1, a);
      conv3(P.ops, pfx + ".conv1", a, TRef(), cin, pfx + ".conv1.weight", pfx + ".conv1.bias", cout, nullptr, 0, "", TRef(), u,
            GEMM_OUT_BF16, false);
    }
    // time scale/shift: rows [ss_cur, ss_cur + 2*cout) of the stacked time_emb_proj
    wall_keys.push_back(pfx + ".time_emb_proj.weight");
    ball_keys.push_back(pfx + ".time_emb_proj.bias");
    const long long ss_off = ss_cur;
    ss_cur += 2 * cout;
    TRef v = u;
    std::string n2 = pfx + ".norm2";
    if (!P.fuse_gn) {
      v = tmp(L, cout);
      emit_gn(P.ops, pfx + ".norm2", u, TRef(), pfx + ".norm2", eps, true, ss_off, v);
      n2.clear();
    }
    TRef y = act(L, cout);
    if (cin != cout) {
      TRef sc[2] = {x0, x1};
      conv3(P.ops, pfx + ".conv2", v, TRef(), cout, pfx + ".conv2.weight", pfx + ".conv2.bias", cout, sc, x1.valid() ? 2 : 1,
            pfx + ".conv_shortcut", TRef(), y, GEMM_OUT_BF16, false, n2, ss_off);
      P.named.emplace_back(pfx, y);
    } else {
      if (x1.valid()) { err = "identity shortcut over a channel concat is not supported"; return y; }
      conv3(P.ops, pfx + ".conv2", v, TRef(), cout, pfx + ".conv2.weight", pfx + ".conv2.bias", cout, nullptr, 0, "", x0, y,
            GEMM_OUT_BF16, false, n2, ss_off);
      P.named.emplace_back(pfx, y);
    }
    return y;
  }

  TRef transformer(const std::string& pfx, TRef y) {
    if (P.fuse && !P.split && y.C % 128 == 0 && y.C <= 384 && y.C <= P.fuse_max_c) return transformer_fused(pfx, y);
    scratch_reset();
    const int L = y.L, C = y.C, H = P.cfg.num_heads, D = C / H;
    const std::string tb = pfx + ".transformer_blocks.0";
    const bool ln_fused = P.fuse_ln && !P.split && C % 128 == 0 && C <= 384;
    const bool gn_in_gemm = P.fuse_gn && !ln_fused;   // (the one-stage chain of DV_FUSE_LN has no operand transform)
    TRef g = y;                                  // fused: proj_in reads the raw block input and normalises it in flight
    std::string gn_key = pfx + ".norm";
    if (!gn_in_gemm) {
      g = gn_target(y, TRef(), L, C);
      emit_gn(P.ops, pfx + ".norm", y, TRef(), pfx + ".norm", 1e-6f, false, -1, g);
      gn_key.clear();
    }
    TRef z = tmp(L, C, true);
    TRef nrm = tmp(L, C);
    TRef qkv = tmp(L, 3 * C);
    TRef ao = tmp(L, C);
    // A projection whose output row fits one CTA's TMEM accumulator (C <= 384) takes the LayerNorm that follows it
    // along as a one-stage chain (csrc/chain_tc.cu, CH_ROW with a global LayerNorm output): one launch instead of two.
    auto proj_ln = [&](const std::string& tag, TRef a, const std::string& wkey, const std::string& bkey, TRef res,
                       const std::string& lnkey) {
      if (!ln_fused) {
        linear(P.ops, tag, a, "linear", {wkey}, bkey, C, res, z, GEMM_OUT_BF16, false,
               (tag == pfx + ".proj_in") ? gn_key : std::string(), 1e-6f);
        emit_ln(P.ops, lnkey, z, lnkey, nrm);
        return;
      }
      Op& o = emit_chain(tag + "+ln", L, C);
      ChainDesc& c = o.ch;
      c.nstages = 1;
      c.st[0].kind = CH_ROW; c.st[0].a = a; c.st[0].res = res; c.st[0].out = z; c.st[0].out2 = nrm;
      set_w(c.st[0], tag, "linear", {wkey}, bkey, C, C);
      set_ln(c.st[0], lnkey, C);
    };
    // --- self attention
    proj_ln(pfx + ".proj_in", g, pfx + ".proj_in.weight", pfx + ".proj_in.bias", TRef(), tb + ".norm1");
    linear(P.ops, tb + ".attn1.qkv", nrm, "linear",
           {tb + ".attn1.to_q.weight", tb + ".attn1.to_k.weight", tb + ".attn1.to_v.weight"}, "", 3 * C, TRef(), qkv,
           GEMM_OUT_BF16);
    {
      Op o;
      o.kind = OP_ATTN; o.tag = tb + ".attn1";
      memset(&o.at, 0, sizeof o.at);
      o.ato.q = qkv.off; o.ato.k = qkv.off + static_cast<size_t>(C) * 2; o.ato.v = qkv.off + static_cast<size_t>(C) * 4;
      o.ato.o = ao.off; o.ato.cross = false;
      o.at.ld_q = o.at.ld_k = o.at.ld_v = 3 * C; o.at.ld_o = C;
      o.at.bs_q = o.at.bs_k = o.at.bs_v = static_cast<long long>(L) * 3 * C; o.at.bs_o = static_cast<long long>(L) * C;
      o.at.B = P.B; o.at.H = H; o.at.Lq = L; o.at.Lk = L; o.at.D = D;
      o.at.scale = 1.0f / sqrtf(static_cast<float>(D));
      P.ops.push_back(o);
    }
    proj_ln(tb + ".attn1.to_out", ao, tb + ".attn1.to_out.0.weight", tb + ".attn1.to_out.0.bias", z, tb + ".norm2");
    // --- cross attention to the prompt (K/V hoisted into prep_ops)
    TRef kv;
    kv.L = P.Tp; kv.C = 2 * C;
    kv.lo = P.triple;
    kv.off = ws_alloc(tensor_bytes(P.Tp, 2 * C, kv.lo));
    linear(P.prep_ops, tb + ".attn2.kv", ehs_bf16, "linear", {tb + ".attn2.to_k.weight", tb + ".attn2.to_v.weight"}, "",
           2 * C, TRef(), kv, GEMM_OUT_BF16);
    TRef q = qkv;  // reuse the first C columns' storage as a dense [L, C] tensor
    q.C = C;
    if (P.triple) q = tmp(L, C);  // (an alias would put its lo plane inside qkv's hi plane: keep the two-plane rule simple)
    linear(P.ops, tb + ".attn2.to_q", nrm, "linear", {tb + ".attn2.to_q.weight"}, "", C, TRef(), q, GEMM_OUT_BF16);
    {
      Op o;
      o.kind = OP_ATTN; o.tag = tb + ".attn2";
      memset(&o.at, 0, sizeof o.at);
      o.ato.q = q.off; o.ato.k = kv.off; o.ato.v = kv.off + static_cast<size_t>(C) * 2; o.ato.o = ao.off;
      o.ato.cross = true;
      o.at.ld_q = C; o.at.ld_k = o.at.ld_v = 2 * C; o.at.ld_o = C;
      o.at.bs_q = static_cast<long long>(L) * C; o.at.bs_k = o.at.bs_v = static_cast<long long>(P.Tp) * 2 * C;
      o.at.bs_o = static_cast<long long>(L) * C;
      o.at.B = P.B; o.at.H = H; o.at.Lq = L; o.at.Lk = P.Tp; o.at.D = D;
      o.at.scale = 1.0f / sqrtf(static_cast<float>(D));
      P.ops.push_back(o);
    }
    proj_ln(tb + ".attn2.to_out", ao, tb + ".attn2.to_out.0.weight", tb + ".attn2.to_out.0.bias", z, tb + ".norm3");
    // --- GEGLU feed-forward
    TRef gg = tmp(L, 4 * C);
    linear(P.ops, tb + ".ff.geglu", nrm, "geglu", {tb + ".ff.net.0.proj.weight"}, tb + ".ff.net.0.proj.bias", 8 * C,
           TRef(), gg, GEMM_OUT_GEGLU);
    linear(P.ops, tb + ".ff.out", gg, "linear", {tb + ".ff.net.2.weight"}, tb + ".ff.net.2.bias", C, z, z, GEMM_OUT_BF16);
    // --- proj_out + residual
    TRef yo = act(L, C);
    linear(P.ops, pfx + ".proj_out", z, "linear", {pfx + ".proj_out.weight"}, pfx + ".proj_out.bias", C, y, yo,
           GEMM_OUT_BF16, true);
    P.named.emplace_back(pfx, yo);
    return yo;
  }

  // ---- the same block as chained-GEMM launches (csrc/chain_tc.cu): 6 launches instead of 15
  //   A : proj_in -> z ; LN1 ; qkv                      B1: to_out(attn1) + z -> z ; LN2 ; to_q
  //   B2: to_out(attn2) + z -> z ; LN3 ; GEGLU ; ff.out + z ; proj_out + y -> block output (+GroupNorm partials)
  Op& emit_chain(const std::string& tag, int L, int C) {
    Op o;
    o.kind = OP_CHAIN; o.tag = tag;
    o.ch.L = L; o.ch.C = C;
    P.ops.push_back(o);
    return P.ops.back();
  }
  void set_w(ChainStageDesc& sd, const std::string& tag, const std::string& recipe, std::vector<std::string> keys,
             const std::string& bias_key, int n, int k, bool second = false) {
    const long long wo = slot(tag + ".W", recipe, std::move(keys), "bf16", {n, k});
    long long bo = -1;
    if (!bias_key.empty()) bo = slot(tag + ".b", recipe == "geglu" ? "geglu_vec" : "vec", {bias_key}, "f32", {n});
    if (second) { sd.w2_off = wo; sd.w2_rows = n; sd.w2_k = k; sd.bias2_off = bo; }
    else { sd.w_off = wo; sd.w_rows = n; sd.w_k = k; sd.bias_off = bo; }
  }
  void set_ln(ChainStageDesc& sd, const std::string& key, int C) {
    sd.gamma = vec(key + ".weight", C);
    sd.beta = vec(key + ".bias", C);
    sd.eps = 1e-5f;
  }

  TRef transformer_fused(const std::string& pfx, TRef y) {
    scratch_reset();
    const int L = y.L, C = y.C, H = P.cfg.num_heads, D = C / H;
    const std::string tb = pfx + ".transformer_blocks.0";
    TRef g = gn_target(y, TRef(), L, C);
    emit_gn(P.ops, pfx + ".norm", y, TRef(), pfx + ".norm", 1e-6f, false, -1, g);
    TRef z = tmp(L, C);
    TRef qkv = tmp(L, 3 * C);
    TRef ao = tmp(L, C);
    {
      Op& o = emit_chain(pfx + ".chain_a", L, C);
      ChainDesc& c = o.ch;
      c.nstages = 2;
      c.st[0].kind = CH_ROW; c.st[0].a = g; c.st[0].out = z;
      set_w(c.st[0], pfx + ".proj_in", "linear", {pfx + ".proj_in.weight"}, pfx + ".proj_in.bias", C, C);
      set_ln(c.st[0], tb + ".norm1", C);
      c.st[1].kind = CH_STREAM; c.st[1].n = 3 * C; c.st[1].out = qkv;
      set_w(c.st[1], tb + ".attn1.qkv", "linear",
            {tb + ".attn1.to_q.weight", tb + ".attn1.to_k.weight", tb + ".attn1.to_v.weight"}, "", 3 * C, C);
    }
    {
      Op o;
      o.kind = OP_ATTN; o.tag = tb + ".attn1";
      memset(&o.at, 0, sizeof o.at);
      o.ato.q = qkv.off; o.ato.k = qkv.off + static_cast<size_t>(C) * 2; o.ato.v = qkv.off + static_cast<size_t>(C) * 4;
      o.ato.o = ao.off; o.ato.cross = false;
      o.at.ld_q = o.at.ld_k = o.at.ld_v = 3 * C; o.at.ld_o = C;
      o.at.bs_q = o.at.bs_k = o.at.bs_v = static_cast<long long>(L) * 3 * C; o.at.bs_o = static_cast<long long>(L) * C;
      o.at.B = P.B; o.at.H = H; o.at.Lq = L; o.at.Lk = L; o.at.D = D;
      o.at.scale = 1.0f / sqrtf(static_cast<float>(D));
      P.ops.push_back(o);
    }
    TRef kv;
    kv.L = P.Tp; kv.C = 2 * C;
    kv.lo = P.triple;
    kv.off = ws_alloc(tensor_bytes(P.Tp, 2 * C, kv.lo));
    linear(P.prep_ops, tb + ".attn2.kv", ehs_bf16, "linear", {tb + ".attn2.to_k.weight", tb + ".attn2.to_v.weight"}, "",
           2 * C, TRef(), kv, GEMM_OUT_BF16);
    TRef q = qkv;  // the first [L, C] worth of the qkv scratch, dense
    q.C = C;
    {
      Op& o = emit_chain(pfx + ".chain_b1", L, C);
      ChainDesc& c = o.ch;
      c.nstages = 2;
      c.st[0].kind = CH_ROW; c.st[0].a = ao; c.st[0].res = z; c.st[0].out = z;
      set_w(c.st[0], tb + ".attn1.to_out", "linear", {tb + ".attn1.to_out.0.weight"}, tb + ".attn1.to_out.0.bias", C, C);
      set_ln(c.st[0], tb + ".norm2", C);
      c.st[1].kind = CH_STREAM; c.st[1].n = C; c.st[1].out = q;
      set_w(c.st[1], tb + ".attn2.to_q", "linear", {tb + ".attn2.to_q.weight"}, "", C, C);
    }
    {
      Op o;
      o.kind = OP_ATTN; o.tag = tb + ".attn2";
      memset(&o.at, 0, sizeof o.at);
      o.ato.q = q.off; o.ato.k = kv.off; o.ato.v = kv.off + static_cast<size_t>(C) * 2; o.ato.o = ao.off;
      o.ato.cross = true;
      o.at.ld_q = C; o.at.ld_k = o.at.ld_v = 2 * C; o.at.ld_o = C;
      o.at.bs_q = static_cast<long long>(L) * C; o.at.bs_k = o.at.bs_v = static_cast<long long>(P.Tp) * 2 * C;
      o.at.bs_o = static_cast<long long>(L) * C;
      o.at.B = P.B; o.at.H = H; o.at.Lq = L; o.at.Lk = P.Tp; o.at.D = D;
      o.at.scale = 1.0f / sqrtf(static_cast<float>(D));
      P.ops.push_back(o);
    }
    TRef yo = act(L, C);
    {
      Op& o = emit_chain(pfx + ".chain_b2", L, C);
      ChainDesc& c = o.ch;
      c.nstages = 3;
      c.st[0].kind = CH_ROW; c.st[0].a = ao; c.st[0].res = z; c.st[0].out = z;
      set_w(c.st[0], tb + ".attn2.to_out", "linear", {tb + ".attn2.to_out.0.weight"}, tb + ".attn2.to_out.0.bias", C, C);
      set_ln(c.st[0], tb + ".norm3", C);
      c.st[1].kind = CH_FF; c.st[1].n = 4 * C; c.st[1].res = z;
      set_w(c.st[1], tb + ".ff.geglu", "geglu", {tb + ".ff.net.0.proj.weight"}, tb + ".ff.net.0.proj.bias", 8 * C, C);
      set_w(c.st[1], tb + ".ff.out", "linear", {tb + ".ff.net.2.weight"}, tb + ".ff.net.2.bias", C, 4 * C, true);
      c.st[2].kind = CH_STREAM; c.st[2].n = C; c.st[2].res = y; c.st[2].out = yo;
      set_w(c.st[2], pfx + ".proj_out", "linear", {pfx + ".proj_out.weight"}, pfx + ".proj_out.bias", C, C);
      // GroupNorm partials of the block output, same layout as a GEMM epilogue's
      const int chunks = ((L + 127) / 128) * 4;
      c.st[2].stats_off = static_cast<long long>(P.stats_bytes);
      c.st[2].stats_nblk = C / 16;
      P.stats_bytes += align_up(static_cast<size_t>(P.B) * chunks * (C / 16) * 2 * sizeof(float), 256);
      yo.st_off = c.st[2].stats_off; yo.st_chunks = chunks; yo.st_nblk = C / 16; yo.st_fold = 0;
    }
    P.named.emplace_back(pfx, yo);
    return yo;
  }

  // conv k3 s2 p1 (resnet.py:214-223): L -> ceil(L / 2). The input is read through the pair view [L/2, 2C] (row l =
  // x[2l] | x[2l+1]). For odd L the last pair has only its first half: the first-half tap uses a view with ceil(L/2)
  // rows, the second-half taps a view with floor(L/2) rows, so the missing x[L] is TMA zero fill = the conv padding.
  TRef downsample(const std::string& key, TRef x) {
    const int C = x.C, Lo = (x.L + 1) / 2;
    TRef out = act(Lo, C);
    Op& o = emit_gemm(P.ops, key);
    GemmDesc& g = o.g;
    add_src(g, x);
    g.view_l[0] = Lo; g.view_c[0] = 2 * C;
    add_src(g, x);
    g.view_l[1] = x.L / 2; g.view_c[1] = 2 * C;
    g.alg_nsrc = 1;
    add_seg(g, 1, -1, C, C);  // tap 0: x[2l-1] = second half of view row l-1
    add_seg(g, 0, 0, 0, C);   // tap 1: x[2l]
    add_seg(g, 1, 0, C, C);   // tap 2: x[2l+1]
    g.n = C; g.n_pad = round_up(C, 128);
    g.w_off = slot(key + ".W", "conv3", {key + ".weight"}, "bf16", {g.n_pad, g.k_tot});
    P.slots.back().shape.push_back(C);
    P.slots.back().bytes = static_cast<size_t>(g.n_pad) * g.k_tot * 2 * (P.triple ? 2 : 1);
    g.bias_off = slot(key + ".b", "vec_sum", {key + ".bias"}, "f32", {g.n_pad});
    g.l_out = Lo; g.out_mode = GEMM_OUT_BF16; g.ld_out = C;
    attach_stats(g, out);
    out.prod = static_cast<int>(P.ops.size()) - 1;
    g.out = out;
    P.named.emplace_back(key, out);
    return out;
  }

  // nearest resize to the skip's length Ls (2L or 2L - 1; resnet.py:138-173, `size=` path) + conv k3 p1, folded into
  // ONE GEMM over the low-resolution rows with N = 2C (columns [0, C) = output row 2p, [C, 2C) = row 2p + 1). With
  // u[i] = x[i / 2] for i < Ls (nearest with size = Ls is exactly that, also for Ls = 2L - 1) and zero outside:
  //   row 2p     = W0 x[p-1] + W1 x[p] + W2 u[2p+1]        u[2p+1] = x[p], except that u[Ls] is conv padding when Ls is odd
  //   row 2p + 1 = (W0 + W1) x[p] + W2 x[p+1]
  // Ls even: three K segments (x, -1) (x, 0) (x, +1) with weights [W0 | W1+W2 | 0 ; 0 | W0+W1 | W2] ("conv3_up2").
  // Ls odd:  four segments (x, -1) (x, 0) (x', 0) (x, +1), [W0 | W1 | W2 | 0 ; 0 | W0+W1 | 0 | W2] ("conv3_up2_odd"), x' = the
  //          same tensor behind a map one row shorter (TMA zero fill removes the W2 term of the last even row); the odd
  //          half of the last GEMM row does not exist: it is neither stored nor counted in the GroupNorm partials.
  TRef upsample(const std::string& key, TRef x, int Ls) {
    const int C = x.C;
    if (Ls != 2 * x.L && Ls != 2 * x.L - 1) { err = "internal: upsample target length"; return x; }
    TRef out = act(Ls, C);
    Op& o = emit_gemm(P.ops, key);
    GemmDesc& g = o.g;
    const bool odd = (Ls & 1) != 0;
    add_src(g, x);
    if (!odd && P.gemm2) {
      add_seg(g, 0, -1, 0, C, 3);                 // (x, -1) (x, 0) (x, +1): one staged box, three row-shifted reads
    } else if (!odd) {
      add_seg(g, 0, -1, 0, C);
      add_seg(g, 0, 0, 0, C);
      add_seg(g, 0, 1, 0, C);
    } else {
      add_seg(g, 0, -1, 0, C);
      add_seg(g, 0, 0, 0, C);
      add_src(g, x);
      g.view_l[1] = x.L - 1;
      add_seg(g, 1, 0, 0, C);
      add_seg(g, 0, 1, 0, C);
    }
    g.n = 2 * C; g.n_pad = round_up(2 * C, 128);
    // both packings live in the parameter blob so that its layout does not depend on (B, T, Tp): one packed blob per
    // module serves every plan (engine.py: DenoiserGroup shares it)
    const long long w3 = slot(key + ".W", "conv3_up2", {key + ".weight"}, "bf16", {g.n_pad, 3 * C});
    const long long w4 = slot(key + ".Wodd", "conv3_up2_odd", {key + ".weight"}, "bf16", {g.n_pad, 4 * C});
    g.w_off = odd ? w4 : w3;
    g.bias_off = slot(key + ".b", "vec_dup2", {key + ".bias"}, "f32", {g.n_pad});
    g.out = out; g.l_out = x.L; g.out_mode = GEMM_OUT_BF16; g.ld_out = 2 * C;
    g.out_bs = static_cast<long long>(Ls) * C;
    g.trim_col = (Ls & 1) ? C : 0;
    g.alg_k = 3 * C; g.alg_nsrc = 1;   // the reference convolves 2L rows with 3 taps; the 4th K segment is bookkeeping
    attach_stats(g, out);
    out.st_fold = 1;             // partials are indexed by low-res row chunk and 2C/16 blocks: block kb also at kb + C/16
    P.named.emplace_back(key, out);
    return out;
  }
};

std::string json_escape(const std::string& s) { return s; }  // keys are [A-Za-z0-9_.] only

}  // namespace

std::string manifest_json(const std::vector<ParamSlot>& slots) {
  std::ostringstream js;
  js << "[";
  for (size_t i = 0; i < slots.size(); ++i) {
    const ParamSlot& s = slots[i];
    if (i) js << ",";
    js << "{\"name\":\"" << json_escape(s.name) << "\",\"recipe\":\"" << s.recipe << "\",\"dtype\":\"" << s.dtype
       << "\",\"offset\":" << s.off << ",\"nbytes\":" << s.bytes << ",\"shape\":[";
    for (size_t k = 0; k < s.shape.size(); ++k) js << (k ? "," : "") << s.shape[k];
    js << "],\"keys\":[";
    for (size_t k = 0; k < s.keys.size(); ++k) js << (k ? "," : "") << "\"" << json_escape(s.keys[k]) << "\"";
    js << "]}";
  }
  js << "]";
  return js.str();
}

int Plan::build(std::string& err) {
  const dv_unet_cfg& c = cfg;
  n_levels = c.n_levels;
  if (n_levels < 2 || n_levels > 8) { err = "n_levels must be in [2, 8]"; return DV_ERR_INVALID; }
  if (c.layers_per_block < 1 || c.layers_per_block > 4) { err = "layers_per_block must be in [1, 4]"; return DV_ERR_INVALID; }
  if (B < 1 || T < 1 || Tp < 1) { err = "B, T, Tp must be positive"; return DV_ERR_INVALID; }
  if ((T >> (n_levels - 1)) < 2) { err = "T too short for the number of levels (need T >= 2^n_levels)"; return DV_ERR_INVALID; }
  for (int i = 0; i < n_levels; ++i) {
    const int ch = c.block_out_channels[i];
    if (ch % 64 != 0 || ch > 1024) { err = "block_out_channels must be multiples of 64 (<= 1024)"; return DV_ERR_UNSUPPORTED; }
    if (ch % c.num_heads != 0) { err = "channels not divisible by heads"; return DV_ERR_INVALID; }
    const int d = ch / c.num_heads;
    if (i < n_levels - 1 || true) {
      if (d != 16 && d != 32 && d != 48 && d != 64) { err = "head dim must be 16/32/48/64"; return DV_ERR_UNSUPPORTED; }
    }
    // GroupNorm statistics are kept per 16-channel block (GEMM epilogue partials, gn_apply's kb0 / nkb): a group must be a
    // whole number of blocks, for every level and therefore for every concat of two levels
    if (c.norm_num_groups < 1 || ch % (16 * c.norm_num_groups) != 0) {
      err = "block_out_channels must be multiples of 16 * norm_num_groups";
      return DV_ERR_UNSUPPORTED;
    }
  }
  if (c.cross_attention_dim % 64 != 0) { err = "cross_attention_dim must be a multiple of 64"; return DV_ERR_UNSUPPORTED; }
  if (c.x_channels < 1 || c.x_channels > c.in_channels) { err = "x_channels out of range"; return DV_ERR_INVALID; }
  if (c.out_channels > 128) { err = "out_channels > 128 not supported"; return DV_ERR_UNSUPPORTED; }
  if (c.precision != DV_PRECISION_BF16 && c.precision != DV_PRECISION_FP32X3 && c.precision != DV_PRECISION_FP16S &&
      c.precision != DV_PRECISION_FP16) {
    err = "unknown precision";
    return DV_ERR_INVALID;
  }
  triple = c.precision == DV_PRECISION_FP32X3;
  f16 = c.precision == DV_PRECISION_FP16S || c.precision == DV_PRECISION_FP16;
  split = triple || c.precision == DV_PRECISION_FP16S;
  K = &kernel_api(f16);
  emb_dim = 4 * c.block_out_channels[0];
  if (emb_dim != 512) { err = "time_embed_dim must be 512 (block_out_channels[0] == 128)"; return DV_ERR_UNSUPPORTED; }
  cin_pad = round_up(c.in_channels, 64);

  {
    const char* fe = getenv("DV_FUSE");
    fuse = !(fe && fe[0] == '0');
    { const char* fl = getenv("DV_FUSE_LN"); fuse_ln = fl && fl[0] == '1'; }
    const char* fm = getenv("DV_FUSE_MAXC");
    fuse_max_c = fm ? atoi(fm) : 128;
    const char* fg = getenv("DV_FUSE_GN");
    fuse_gn = !split && fg && fg[0] == '1';
    const char* g2 = getenv("DV_GEMM2");
    gemm2 = fuse_gn || (g2 && g2[0] == '1');
    const char* go = getenv("DV_GN_OUT");
    // default: on in the single-plane modes (BF16 / FP16: C2 step -0.8 %), off in the split modes where the tail re-reads and
    // rewrites two planes and measured slower (fp16s 239.6 vs 235.8 ms, fp32x3 414.3 vs 412.7); DV_GN_OUT=1 / 0 forces it
    gn_out = !gemm2 && (go ? go[0] != '0' : !split);
  }
  Builder b(*this);
  const int C0 = c.block_out_channels[0];

  // ---- persistent regions
  off_xin = b.ws_alloc(b.tensor_bytes(T, cin_pad, triple));
  b.ehs_bf16.L = Tp; b.ehs_bf16.C = c.cross_attention_dim; b.ehs_bf16.lo = triple;
  b.ehs_bf16.off = b.ws_alloc(b.tensor_bytes(Tp, c.cross_attention_dim, triple));
  off_ehs = b.ehs_bf16.off;
  off_aug = b.ws_alloc(static_cast<size_t>(B) * emb_dim * 4);
  off_embact = b.ws_alloc(static_cast<size_t>(B) * emb_dim * 4 * 2);          // two time-table slots
  off_kvlen = b.ws_alloc(static_cast<size_t>(B) * 4);
  const size_t xbytes = static_cast<size_t>(B) * c.x_channels * T * 4;
  const size_t obytes = static_cast<size_t>(B) * c.out_channels * T * 4;
  off_x = b.ws_alloc(xbytes);
  off_xp = b.ws_alloc(xbytes);
  off_base = b.ws_alloc(xbytes);
  off_h0 = b.ws_alloc(obytes);
  off_h1 = b.ws_alloc(obytes);
  off_ttab = b.ws_alloc(4096 * 4);
  off_gncnt = b.ws_alloc(static_cast<size_t>(B + 1) * 4);   // + the sticky "a wait gave up" flag
  diag_ctas = sampler_diag_ctas(static_cast<size_t>(B) * c.x_channels * T);
  off_diag_part = b.ws_alloc(static_cast<size_t>(DIAG_MAX_ROWS) * diag_ctas * 16);
  off_diag_rows = b.ws_alloc(static_cast<size_t>(DIAG_MAX_ROWS) * 16);

  // ---- time embedding parameters (fp32; tiny GEMVs on CUDA cores)
  p_freqs = b.slot("time.freqs", "time_freqs", {}, "f32", {C0 / 2});
  p_w1 = b.slot("time.w1", "rows", {"time_embedding.linear_1.weight"}, "f32", {emb_dim, C0});
  p_b1 = b.vec("time_embedding.linear_1.bias", emb_dim);
  p_w2 = b.slot("time.w2", "rows", {"time_embedding.linear_2.weight"}, "f32", {emb_dim, emb_dim});
  p_b2 = b.vec("time_embedding.linear_2.bias", emb_dim);

  // ---- the network (unet_1d_condition.py:943-1032)
  TRef xin;
  xin.off = off_xin; xin.L = T; xin.C = cin_pad; xin.lo = triple;
  std::vector<TRef> skips;
  TRef h = b.act(T, C0);
  b.conv3(ops, "conv_in", xin, TRef(), c.in_channels, "conv_in.weight", "conv_in.bias", C0, nullptr, 0, "", TRef(), h,
          GEMM_OUT_BF16, false);
  named.emplace_back("conv_in", h);
  skips.push_back(h);
  int ch = C0;
  for (int i = 0; i < n_levels; ++i) {
    const int co = c.block_out_channels[i];
    const bool last = (i == n_levels - 1);
    const std::string bp = "down_blocks." + std::to_string(i);
    for (int j = 0; j < c.layers_per_block; ++j) {
      h = b.resnet(bp + ".resnets." + std::to_string(j), h, TRef(), co);
      if (!last) h = b.transformer(bp + ".attentions." + std::to_string(j), h);
      skips.push_back(h);
      ch = co;
    }
    if (!last) {
      h = b.downsample(bp + ".downsamplers.0.conv", h);
      skips.push_back(h);
    }
  }
  h = b.resnet("mid_block.resnets.0", h, TRef(), ch);
  h = b.transformer("mid_block.attentions.0", h);
  h = b.resnet("mid_block.resnets.1", h, TRef(), ch);
  for (int i = 0; i < n_levels; ++i) {
    const int co = c.block_out_channels[n_levels - 1 - i];
    const bool final_block = (i == n_levels - 1);
    const std::string bp = "up_blocks." + std::to_string(i);
    for (int j = 0; j < c.layers_per_block + 1; ++j) {
      TRef skip = skips.back();
      skips.pop_back();
      if (skip.L != h.L) { err = "internal: skip length mismatch"; return DV_ERR_INVALID; }
      h = b.resnet(bp + ".resnets." + std::to_string(j), h, skip, co);
      if (i > 0) h = b.transformer(bp + ".attentions." + std::to_string(j), h);
    }
    if (!final_block) h = b.upsample(bp + ".upsamplers.0.conv", h, skips.back().L);
  }
  if (!b.err.empty()) { err = b.err; return DV_ERR_UNSUPPORTED; }
  {
    b.scratch_reset();
    TRef dummy;
    dummy.L = T; dummy.C = c.out_channels;
    if (fuse_gn) {
      b.conv3(ops, "conv_out", h, TRef(), C0, "conv_out.weight", "conv_out.bias", c.out_channels, nullptr, 0, "", TRef(), dummy,
              GEMM_OUT_F32_NCL, true, "conv_norm_out", -1);
    } else {
      TRef a = b.gn_target(h, TRef(), T, C0);
      b.emit_gn(ops, "conv_norm_out", h, TRef(), "conv_norm_out", c.norm_eps, true, -1, a);
      b.conv3(ops, "conv_out", a, TRef(), C0, "conv_out.weight", "conv_out.bias", c.out_channels, nullptr, 0, "", TRef(), dummy,
              GEMM_OUT_F32_NCL, true);
    }
  }

  // ---- stacked time_emb_proj
  n_all = b.ss_cur;
  p_wall = b.slot("time.w_all_t", "rows_t", b.wall_keys, "f32", {emb_dim, n_all});
  p_ball = b.slot("time.b_all", "rows", b.ball_keys, "f32", {n_all});
  off_ss = b.ws_alloc(static_cast<size_t>(B) * n_all * 4 * 2);                // two time-table slots
  off_stats = b.ws_alloc(stats_bytes);
  const size_t scratch_base = b.ws_alloc(b.scratch_max);
  const size_t bound_stride = align_up(b.bound_max, 1024);
  const size_t bound_base = b.ws_alloc(2 * bound_stride);
  ws_bytes = align_up(b.ws_cur, 1024);

  // ---- resolve scratch-relative offsets and n_all
  auto fix = [&](size_t& o) {
    if (o & Builder::SCRATCH_TAG) o = (o & ~Builder::SCRATCH_TAG) + scratch_base;
    else if (o & Builder::BOUND_TAG) o = (o & ~(Builder::BOUND_TAG | Builder::BOUND_PAR)) + bound_base + ((o & Builder::BOUND_PAR) ? bound_stride : 0);
  };
  auto fix_list = [&](std::vector<Op>& list) {
    for (Op& o : list) {
      switch (o.kind) {
        case OP_GEMM:
          for (int i = 0; i < o.g.nsrc; ++i) fix(o.g.src[i].off);
          if (o.g.res.valid()) fix(o.g.res.off);
          fix(o.g.out.off);
          if (o.g.xf.on) { fix(o.g.xf.x0.off); if (o.g.xf.x1.valid()) fix(o.g.xf.x1.off); }
          if (o.g.gnout.on) fix(o.g.gnout.out.off);
          break;
        case OP_GN_APPLY:
          fix(o.gno.x0); fix(o.gno.x1); fix(o.gno.out);
          o.gn.ss_ld = n_all;
          break;
        case OP_LN: fix(o.lno.x); fix(o.lno.out); break;
        case OP_CHAIN:
          for (int i = 0; i < o.ch.nstages; ++i) {
            ChainStageDesc& sd = o.ch.st[i];
            if (sd.a.valid()) fix(sd.a.off);
            if (sd.res.valid()) fix(sd.res.off);
            if (sd.out.valid()) fix(sd.out.off);
            if (sd.out2.valid()) fix(sd.out2.off);
          }
          break;
        case OP_ATTN: {
          // q/k/v offsets may carry an added column offset: fix preserves low bits
          fix(o.ato.q); fix(o.ato.k); fix(o.ato.v); fix(o.ato.o);
          break;
        }
      }
    }
  };
  fix_list(ops);
  fix_list(prep_ops);

  // ---- algorithmic work per launch (roofline accounting in bench.py)
  for (Op& o : ops) {
    const double Bd = B;
    switch (o.kind) {
      case OP_GEMM: {
        double abytes = 0;
        for (int i = 0; i < (o.g.alg_nsrc ? o.g.alg_nsrc : o.g.nsrc); ++i) abytes += Bd * o.g.src[i].L * o.g.src[i].C * 2;
        o.flops = 2.0 * Bd * o.g.l_out * o.g.n * (o.g.alg_k ? o.g.alg_k : o.g.k_tot);
        const double outw = o.g.out_mode == GEMM_OUT_F32_NCL ? 4.0 * o.g.n : (o.g.out_mode == GEMM_OUT_GEGLU ? o.g.n : 2.0 * o.g.n);
        o.bytes = abytes + 2.0 * o.g.n_pad * o.g.k_tot + Bd * o.g.l_out * outw +
                  (o.g.res.valid() ? Bd * o.g.l_out * o.g.n * 2.0 : 0.0) +
                  (o.g.gnout.on ? Bd * o.g.l_out * o.g.n * 2.0 : 0.0);   // fused output GroupNorm: the normalised copy
        break;
      }
      case OP_ATTN:
        o.flops = 4.0 * Bd * o.at.H * o.at.Lq * static_cast<double>(o.at.Lk) * o.at.D;
        o.bytes = 2.0 * Bd * o.at.H * o.at.D * (2.0 * o.at.Lq + 2.0 * o.at.Lk);
        break;
      case OP_CHAIN: {
        const double rows = Bd * o.ch.L;
        for (int i = 0; i < o.ch.nstages; ++i) {
          const ChainStageDesc& sd = o.ch.st[i];
          o.flops += 2.0 * rows * sd.w_rows * sd.w_k + 2.0 * rows * sd.w2_rows * sd.w2_k;
          o.bytes += 2.0 * sd.w_rows * sd.w_k + 2.0 * sd.w2_rows * sd.w2_k;
          if (sd.a.valid()) o.bytes += rows * sd.a.C * 2.0;
          if (sd.res.valid()) o.bytes += rows * o.ch.C * 2.0;
          if (sd.out.valid()) o.bytes += rows * (sd.kind == CH_STREAM ? sd.n : o.ch.C) * 2.0;
        }
        break;
      }
      case OP_GN_APPLY: o.bytes = 4.0 * Bd * o.gn.L * (o.gn.c0 + o.gn.c1); break;
      case OP_LN: o.bytes = 4.0 * o.ln.rows * static_cast<double>(o.ln.C); break;
    }
  }

  // ---- manifest
  manifest = manifest_json(slots);
  param_bytes = align_up(param_bytes, 256);
  return DV_OK;
}

// --------------------------------------------------------------------------------------------- bind
int Plan::bind(void* workspace, const void* prm, cudaStream_t st, std::string& err) {
  if (!workspace || !prm) { err = "null workspace/params"; return DV_ERR_INVALID; }
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(prm) & 255)) {
    err = "workspace and params must be 256-byte aligned";
    return DV_ERR_INVALID;
  }
  ws = static_cast<uint8_t*>(workspace);
  params = static_cast<const uint8_t*>(prm);
  {
    cudaError_t ie = K->gemm_tc_init();
    if (ie != cudaSuccess) { err = std::string("gemm_tc_init: ") + cudaGetErrorString(ie); return DV_ERR_CUDA; }
    ie = K->chain_tc_init();
    if (ie != cudaSuccess) { err = std::string("chain_tc_init: ") + cudaGetErrorString(ie); return DV_ERR_CUDA; }
    ie = K->attention_tc_init();
    if (ie != cudaSuccess) { err = std::string("attention_tc_init: ") + cudaGetErrorString(ie); return DV_ERR_CUDA; }
  }
  cudaError_t ce = cudaMemsetAsync(ws, 0, ws_bytes, st);
  if (ce != cudaSuccess) { err = cudaGetErrorString(ce); return DV_ERR_CUDA; }

  auto pf = [&](long long off) -> const float* {
    return off < 0 ? nullptr : reinterpret_cast<const float*>(params + off);
  };
  auto bind_list = [&](std::vector<Op>& list) -> bool {
    for (Op& o : list) {
      if (o.kind == OP_GEMM) {
        GemmDesc& g = o.g;
        GemmParams& p = o.gp;
        memset(&p, 0, sizeof p);
        // kernel map 2 * s = hi plane of logical source s, 2 * s + 1 = its lo plane (split precision only)
        for (int i = 0; i < GEMM_MAX_SRC; ++i) {
          const int s = (i >> 1) < g.nsrc ? (i >> 1) : 0;  // unused maps alias source 0 so prefetch.tensormap stays valid
          const TRef& t = g.src[s];
          const size_t lo = (triple && (i & 1)) ? plane(t) : 0;
          int rows = 128;                                  // sources read with shared conv taps stage a 136-row box
          for (int k = 0; k < g.nseg; ++k) if (g.seg[k].src == s && g.seg[k].ntap == 3) rows = 136;
          p.map_rows[i] = static_cast<int16_t>(rows);
          if (!make_tmap_3d(&p.mapA[i], ws + t.off + lo, g.view_c[s], g.view_c[s], g.view_l[s], B,
                            static_cast<size_t>(t.L) * t.C, err, rows))
            return false;
        }
        if (!make_tmap_2d(&p.mapW, params + g.w_off, g.k_tot, g.n_pad, 64, err)) return false;
        p.mapW2 = p.mapW;
        if (triple && !make_tmap_2d(&p.mapW2, params + g.w_off + static_cast<size_t>(g.n_pad) * g.k_tot * 2, g.k_tot, g.n_pad, 64, err))
          return false;
        // logical segment -> kernel segments: a.w (BF16) or a_hi.w_hi + a_hi.w_lo + a_lo.w_hi (split: three K segments
        // over the same weight panels of the hi / lo matrix)
        p.nseg = 0;
        for (int i = 0; i < g.nseg; ++i) {
          GemmSeg sg = g.seg[i];
          const int lsrc = sg.src;
          sg.wsel = 0;
          sg.src = static_cast<int16_t>(2 * lsrc);
          p.seg[p.nseg++] = sg;
          if (triple) {
            sg.wsel = 1;
            p.seg[p.nseg++] = sg;                         // a_hi . w_lo
            sg.wsel = 0;
            sg.src = static_cast<int16_t>(2 * lsrc + 1);
            p.seg[p.nseg++] = sg;                         // a_lo . w_hi
          }
        }
        if (g.xf.on) {                                    // GroupNorm fused into the operand path
          GemmXf& x = p.xf;
          const TRef* xs[2] = {&g.xf.x0, &g.xf.x1};
          x.nst = g.xf.x1.valid() ? 2 : 1;
          for (int i = 0; i < x.nst; ++i) {
            x.st_ptr[i] = reinterpret_cast<const float*>(ws + off_stats + xs[i]->st_off);
            x.st_chunks[i] = xs[i]->st_chunks; x.st_nblk[i] = xs[i]->st_nblk; x.st_fold[i] = xs[i]->st_fold;
            x.st_cover[i] = xs[i]->C;
          }
          x.C = g.xf.x0.C + (g.xf.x1.valid() ? g.xf.x1.C : 0);
          x.L = g.xf.x0.L; x.G = cfg.norm_num_groups; x.eps = g.xf.eps; x.silu = g.xf.silu ? 1 : 0;
          x.gamma = pf(g.xf.gamma); x.beta = pf(g.xf.beta);
          x.scale_shift = g.xf.ss < 0 ? nullptr : reinterpret_cast<const float*>(ws + off_ss) + g.xf.ss;
          x.ss_ld = n_all;
        }
        p.B = B; p.L = g.l_out; p.N = g.n; p.n_pad = g.n_pad;
        p.tiles_per_batch = (g.l_out + 127) / 128;
        p.out_mode = g.out_mode;
        p.bias = pf(g.bias_off);
        p.res = g.res.valid() ? ws + g.res.off : nullptr;
        p.res_lo = (g.res.valid() && g.res.lo) ? ws + g.res.off + plane(g.res) : nullptr;
        p.ld_res = g.res.valid() ? g.res.C : 0;
        p.out = g.out_dynamic ? nullptr : ws + g.out.off;
        p.out_lo = (g.out.lo && !g.out_dynamic && g.out_mode != GEMM_OUT_F32_NCL) ? ws + g.out.off + plane(g.out) : nullptr;
        p.ld_out = g.ld_out;
        p.out_bs = g.out_bs;
        p.trim_col = g.trim_col;
        p.stats = g.stats_off >= 0 ? reinterpret_cast<float*>(ws + off_stats + g.stats_off) : nullptr;
        p.stats_nblk = g.stats_nblk;
        p.bn = 0;
        if (g.gnout.on) {
          GemmGnOut& n = p.gn;
          n.out = ws + g.gnout.out.off;
          n.out_lo = g.gnout.out.lo ? ws + g.gnout.out.off + plane(g.gnout.out) : nullptr;
          n.counter = reinterpret_cast<unsigned*>(ws + off_gncnt);
          n.G = cfg.norm_num_groups; n.eps = g.gnout.eps; n.silu = g.gnout.silu ? 1 : 0;
          n.gamma = pf(g.gnout.gamma); n.beta = pf(g.gnout.beta);
          n.scale_shift = g.gnout.ss < 0 ? nullptr : reinterpret_cast<const float*>(ws + off_ss) + g.gnout.ss;
          n.ss_ld = n_all;
          { const char* se = getenv("DV_GN_STASH"); n.no_stash = (se && se[0] == '0') ? 1 : 0; }
        }
#ifdef DV_DEBUG_HOOKS
        { const char* de = getenv("DV_GEMM_DBG"); p.dbg_flags = de ? atoi(de) : 0; }   // timing experiments only
#endif
      } else if (o.kind == OP_CHAIN) {
        ChainParams& p = o.cp;
        memset(&p, 0, sizeof p);
        p.nstages = o.ch.nstages;
        p.B = B; p.L = o.ch.L; p.C = o.ch.C;
        p.tiles_per_batch = (o.ch.L + 127) / 128;
#ifdef DV_DEBUG_HOOKS
        { const char* de = getenv("DV_CHAIN_DBG"); p.dbg_flags = de ? atoi(de) : 0; }
#endif
        int na = 0, nw = 0;
        for (int i = 0; i < o.ch.nstages; ++i) {
          const ChainStageDesc& sd = o.ch.st[i];
          ChainStage& cs = p.st[i];
          cs.kind = sd.kind;
          cs.a_map = -1;
          if (sd.a.valid()) {
            if (na >= 2) { err = "internal: too many streamed operands in a chain"; return false; }
            if (!make_tmap_3d(&p.mapA[na], ws + sd.a.off, sd.a.C, sd.a.C, sd.a.L, B, static_cast<size_t>(sd.a.L) * sd.a.C, err))
              return false;
            cs.a_map = na++;
          }
          if (nw + (sd.w2_rows ? 2 : 1) > CHAIN_MAX_W) { err = "internal: too many weight matrices in a chain"; return false; }
          if (!make_tmap_2d(&p.mapW[nw], params + sd.w_off, sd.w_k, sd.w_rows, 128, err)) return false;
          cs.w_map = nw++;
          cs.w2_map = 0;
          if (sd.w2_rows) {
            if (!make_tmap_2d(&p.mapW[nw], params + sd.w2_off, sd.w2_k, sd.w2_rows, 128, err)) return false;
            cs.w2_map = nw++;
          }
          cs.n = sd.n;
          cs.bias = pf(sd.bias_off); cs.bias2 = pf(sd.bias2_off);
          cs.res = sd.res.valid() ? ws + sd.res.off : nullptr;
          cs.ld_res = sd.res.valid() ? sd.res.C : 0;
          cs.out = sd.out.valid() ? ws + sd.out.off : nullptr;
          cs.ld_out = sd.out.valid() ? sd.out.C : 0;
          cs.out2 = sd.out2.valid() ? ws + sd.out2.off : nullptr;
          cs.ld_out2 = sd.out2.valid() ? sd.out2.C : 0;
          cs.gamma = pf(sd.gamma); cs.beta = pf(sd.beta); cs.eps = sd.eps;
          cs.stats = sd.stats_off >= 0 ? reinterpret_cast<float*>(ws + off_stats + sd.stats_off) : nullptr;
          cs.stats_nblk = sd.stats_nblk;
        }
        // unused descriptors alias valid ones so prefetch.tensormap stays legal
        for (int i = na; i < 2; ++i) p.mapA[i] = p.mapA[0];
        for (int i = nw; i < CHAIN_MAX_W; ++i) p.mapW[i] = p.mapW[0];
      } else if (o.kind == OP_GN_APPLY) {
        o.gn.x0 = ws + o.gno.x0;
        o.gn.x1 = o.gn.c1 ? ws + o.gno.x1 : nullptr;
        o.gn.out = ws + o.gno.out;
        o.gn.x0_lo = o.gn.x1_lo = nullptr; o.gn.out_lo = nullptr;
        {
          const size_t rows = static_cast<size_t>(B) * o.gn.L * 2;
          if (o.gno.x_lo) {
            o.gn.x0_lo = ws + o.gno.x0 + rows * o.gn.c0;
            if (o.gn.c1) o.gn.x1_lo = ws + o.gno.x1 + rows * o.gn.c1;
          }
          if (o.gno.out_lo) o.gn.out_lo = ws + o.gno.out + rows * (o.gn.c0 + o.gn.c1);
        }
        for (int i = 0; i < o.gn.nst; ++i) o.gn.st[i].ptr = reinterpret_cast<const float*>(ws + off_stats + o.gno.st[i]);
        o.gn.gamma = pf(o.gno.gamma);
        o.gn.beta = pf(o.gno.beta);
        o.gn.scale_shift = o.gno.ss < 0 ? nullptr : reinterpret_cast<const float*>(ws + off_ss) + o.gno.ss;
      } else if (o.kind == OP_LN) {
        o.ln.x = ws + o.lno.x; o.ln.out = ws + o.lno.out;
        o.ln.x_lo = nullptr; o.ln.out_lo = nullptr;
        {
          const size_t pl = static_cast<size_t>(o.ln.rows) * o.ln.C * 2;
          if (o.lno.x_lo) o.ln.x_lo = ws + o.lno.x + pl;
          if (o.lno.out_lo) o.ln.out_lo = ws + o.lno.out + pl;
        }
        o.ln.gamma = pf(o.lno.gamma); o.ln.beta = pf(o.lno.beta);
      } else if (o.kind == OP_ATTN) {
        o.at.q = ws + o.ato.q; o.at.k = ws + o.ato.k; o.at.v = ws + o.ato.v; o.at.o = ws + o.ato.o;
        if (triple) {  // lo plane of the tensor each operand lives in: + B * (batch stride) bf16 elements
          o.at.q_lo = ws + o.ato.q + static_cast<size_t>(B) * o.at.bs_q * 2;
          o.at.k_lo = ws + o.ato.k + static_cast<size_t>(B) * o.at.bs_k * 2;
          o.at.v_lo = ws + o.ato.v + static_cast<size_t>(B) * o.at.bs_v * 2;
          o.at.o_lo = ws + o.ato.o + static_cast<size_t>(B) * o.at.bs_o * 2;
        }
        o.at.kv_len = nullptr;  // set in prompt_prepare
        const char* ae = nullptr;
        if (!K->attention_tc_prepare(o.at, &ae)) { err = std::string("attention_tc_prepare: ") + (ae ? ae : "?"); return false; }
      }
    }
    return true;
  };
  if (!bind_list(ops) || !bind_list(prep_ops)) return DV_ERR_CUDA;
  bound = true;
  prompt_ready = false;
  ++graph_epoch;
  if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; graph_key = 0; }
  return DV_OK;
}

// --------------------------------------------------------------------------------------------- execution
#define DV_CK(call)                                                                     \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
      return DV_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

int& skip_kinds_mask() {
#ifdef DV_DEBUG_HOOKS
  static int mask = [] { const char* e = getenv("DV_SKIP_KINDS"); return e ? atoi(e) : 0; }();
#else
  static int mask = 0;   // product build: every op is always launched (dv_debug_skip_kinds exists only in the debug library)
#endif
  return mask;
}

int Plan::enqueue_ops(std::vector<Op>& list, float* dyn_out, cudaStream_t st, std::string& err, int* launches,
                      cudaEvent_t* events, int ss_slot) {
  int idx = 0;
  // DV_SKIP_KINDS (timing experiments only, results are garbage): bit k set = do not launch ops of OpKind k, so the
  // step-time difference is that kernel class's marginal cost inside the captured loop
  const int skip_kinds = skip_kinds_mask();
  for (Op& o : list) {
    if (events) cudaEventRecordWithFlags(events[idx++], st, cudaEventRecordExternal);
    cudaError_t e = cudaSuccess;
    if (skip_kinds & (1 << o.kind)) { if (launches) ++*launches; continue; }
    switch (o.kind) {
      case OP_GEMM:
        if (o.g.out_dynamic || (ss_slot && (o.gp.xf.scale_shift || o.gp.gn.scale_shift))) {
          GemmParams p = o.gp;
          if (o.g.out_dynamic) p.out = dyn_out;
          if (ss_slot && p.xf.scale_shift) p.xf.scale_shift += static_cast<size_t>(ss_slot) * B * n_all;
          if (ss_slot && p.gn.scale_shift) p.gn.scale_shift += static_cast<size_t>(ss_slot) * B * n_all;
          e = K->launch_gemm_tc(p, st);
        } else {
          e = K->launch_gemm_tc(o.gp, st);
        }
        break;
      case OP_CHAIN: e = K->launch_chain_tc(o.cp, st); break;
      case OP_GN_APPLY:
        if (ss_slot && o.gn.scale_shift) {
          GnParams g = o.gn;
          g.scale_shift += static_cast<size_t>(ss_slot) * B * n_all;
          e = K->launch_gn_apply(g, st);
        } else {
          e = K->launch_gn_apply(o.gn, st);
        }
        break;
      case OP_LN:
        e = K->launch_layernorm(o.ln.x, o.ln.out, o.ln.gamma, o.ln.beta, o.ln.rows, o.ln.C, o.ln.eps, st, o.ln.x_lo, o.ln.out_lo);
        break;
      case OP_ATTN: {
        AttnParams a = o.at;
        if (o.ato.cross && has_kvlen) a.kv_len = reinterpret_cast<const int*>(ws + off_kvlen);
        e = K->launch_attention(a, st);
        break;
      }
    }
    if (e != cudaSuccess) {
      err = "launch " + o.tag + ": " + cudaGetErrorString(e);
      return DV_ERR_CUDA;
    }
    if (launches) ++*launches;
  }
  return DV_OK;
}

int Plan::prompt_prepare(const float* ehs, const float* aug, const int32_t* kv_len, cudaStream_t st, std::string& err) {
  if (!bound) { err = "plan not bound"; return DV_ERR_STATE; }
  if (!ehs) { err = "ehs is null"; return DV_ERR_INVALID; }
  const size_t n = static_cast<size_t>(B) * Tp * cfg.cross_attention_dim;
  DV_CK(K->launch_f32_to_bf16(ehs, ws + off_ehs, n, st, triple ? ws + off_ehs + n * 2 : nullptr));
  // The captured loop reads the prompt-side tensors at fixed workspace addresses, so a new batch of the same shape
  // keeps the graph; only a change of WHICH optional inputs exist (their pointers are baked into the launches) drops it.
  const bool topo_changed = prompt_ready && (has_aug != (aug != nullptr) || has_kvlen != (kv_len != nullptr));
  has_aug = aug != nullptr;
  if (aug) DV_CK(cudaMemcpyAsync(ws + off_aug, aug, static_cast<size_t>(B) * emb_dim * 4, cudaMemcpyDeviceToDevice, st));
  has_kvlen = kv_len != nullptr;
  if (kv_len) DV_CK(cudaMemcpyAsync(ws + off_kvlen, kv_len, static_cast<size_t>(B) * 4, cudaMemcpyDeviceToDevice, st));
  int rc = enqueue_ops(prep_ops, nullptr, st, err, nullptr, nullptr);
  if (rc != DV_OK) return rc;
  prompt_ready = true;
  if (topo_changed) {
    ++graph_epoch;
    if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; graph_key = 0; }
  }
  return DV_OK;
}

int Plan::set_cond(const float* cond, cudaStream_t st, std::string& err) {
  if (!bound) { err = "plan not bound"; return DV_ERR_STATE; }
  const int cc = cfg.in_channels - cfg.x_channels;
  if (cc <= 0) return DV_OK;
  DV_CK(K->launch_nct_to_ntc(cond, ws + off_xin, B, cc, T, cin_pad, cfg.x_channels, 1 << 30, 0, st,
                          triple ? ws + off_xin + static_cast<size_t>(B) * T * cin_pad * 2 : nullptr));
  return DV_OK;
}

int Plan::profile_forward(const float* sample, int x_only, const float* t, int t_stride, float* out, cudaStream_t st,
                          float* ms, int cap, std::string& err) {
  // One denoiser call captured into a CUDA graph with an external event-record node between consecutive launches, so
  // the intervals are device-side times of the call AS IT RUNS IN THE PRODUCT (graph replay), free of CPU launch cost.
  const int n = static_cast<int>(ops.size());
  if (cap < n + 1) { err = "ms buffer too small (need ops + 1)"; return DV_ERR_INVALID; }
  if (!cap_stream) DV_CK(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
  std::vector<cudaEvent_t> ev(n + 2);
  for (auto& e : ev) DV_CK(cudaEventCreate(&e));
  int launches = 0;
  DV_CK(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
  cudaError_t ce = cudaEventRecordWithFlags(ev[n + 1], cap_stream, cudaEventRecordExternal);   // prologue start
  int rc = ce == cudaSuccess ? forward(sample, x_only, t, t_stride, out, cap_stream, err, &launches, ev.data()) : DV_ERR_CUDA;
  if (rc == DV_OK) ce = cudaEventRecordWithFlags(ev[n], cap_stream, cudaEventRecordExternal);
  cudaGraph_t graph = nullptr;
  cudaError_t ce2 = cudaStreamEndCapture(cap_stream, &graph);
  cudaGraphExec_t exec = nullptr;
  if (rc == DV_OK && ce == cudaSuccess && ce2 == cudaSuccess) ce2 = cudaGraphInstantiate(&exec, graph, 0);
  if (rc == DV_OK && ce == cudaSuccess && ce2 == cudaSuccess) {
    for (int rep = 0; rep < 2 && ce2 == cudaSuccess; ++rep) ce2 = cudaGraphLaunch(exec, st);   // second replay is timed
    if (ce2 == cudaSuccess) ce2 = cudaStreamSynchronize(st);
    if (ce2 == cudaSuccess) {
      for (int i = 0; i < n; ++i) cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
      cudaEventElapsedTime(&ms[n], ev[n + 1], ev[0]);    // prologue (layout change + time embedding)
    }
  }
  if (exec) cudaGraphExecDestroy(exec);
  if (graph) cudaGraphDestroy(graph);
  for (auto& e : ev) cudaEventDestroy(e);
  if (rc != DV_OK) return rc;
  if (ce != cudaSuccess || ce2 != cudaSuccess) {
    err = std::string("profile_forward: ") + cudaGetErrorString(ce != cudaSuccess ? ce : ce2);
    return DV_ERR_CUDA;
  }
  return DV_OK;
}

int Plan::enqueue_time_tables(const float* t, int t_stride, int slot, cudaStream_t st, std::string& err) {
  TembParams tp;
  tp.t = t; tp.t_stride = t_stride;
  tp.freqs = reinterpret_cast<const float*>(params + p_freqs);
  tp.w1 = reinterpret_cast<const float*>(params + p_w1); tp.b1 = reinterpret_cast<const float*>(params + p_b1);
  tp.w2 = reinterpret_cast<const float*>(params + p_w2); tp.b2 = reinterpret_cast<const float*>(params + p_b2);
  tp.aug = has_aug ? reinterpret_cast<const float*>(ws + off_aug) : nullptr;
  tp.emb_base = reinterpret_cast<float*>(ws + off_embact) + static_cast<size_t>(slot) * B * emb_dim;
  tp.w_all = reinterpret_cast<const float*>(params + p_wall); tp.b_all = reinterpret_cast<const float*>(params + p_ball);
  tp.ss_all = reinterpret_cast<float*>(ws + off_ss) + static_cast<size_t>(slot) * B * n_all;
  tp.B = B; tp.n_all = n_all; tp.sin_dim = cfg.block_out_channels[0]; tp.emb_dim = emb_dim;
  DV_CK(launch_temb(tp, st));
  return DV_OK;
}

int Plan::forward(const float* sample, int x_only, const float* t, int t_stride, float* out, cudaStream_t st,
                  std::string& err, int* launches, cudaEvent_t* events, int tables_slot) {
  if (!bound) { err = "plan not bound"; return DV_ERR_STATE; }
  if (!prompt_ready) { err = "dv_prompt_prepare has not been called"; return DV_ERR_STATE; }
  DV_CK(K->launch_nct_to_ntc(sample, ws + off_xin, B, x_only ? cfg.x_channels : cfg.in_channels, T, cin_pad, 0, 1 << 30, 0, st,
                          triple ? ws + off_xin + static_cast<size_t>(B) * T * cin_pad * 2 : nullptr));
  if (tables_slot < 0) {
    int rc = enqueue_time_tables(t, t_stride, 0, st, err);
    if (rc != DV_OK) return rc;
    tables_slot = 0;
  }
  if (launches) *launches += 3;
  return enqueue_ops(ops, out, st, err, launches, events, tables_slot);
}

static uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
  const uint8_t* p = static_cast<const uint8_t*>(data);
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

// One model evaluation per step (NFE == nsteps); history slots rotate by pointer at enqueue time.
int Plan::enqueue_loop(int sampler, int nsteps, const float* coef, bool rt, cudaStream_t s, int* launches,
                       std::string& err, int* diag_rows) {
  const size_t n = static_cast<size_t>(B) * cfg.x_channels * T;
  float* X = reinterpret_cast<float*>(ws + off_x);
  float* XP = reinterpret_cast<float*>(ws + off_xp);
  float* BASE = reinterpret_cast<float*>(ws + off_base);
  float* H[2] = {reinterpret_cast<float*>(ws + off_h0), reinterpret_cast<float*>(ws + off_h1)};
  const float* ttab = reinterpret_cast<const float*>(ws + off_ttab);
  int cur = 0;  // H[cur] = newest model output m_{-1}, H[cur^1] = m_{-2}
  // sampler diagnostics: update launch j writes its per-CTA partials into row j (rows past DIAG_MAX_ROWS are not recorded)
  int drow = 0;
  auto diag = [&]() -> float* {
    if (!diag_rows || drow >= DIAG_MAX_ROWS) return nullptr;
    return reinterpret_cast<float*>(ws + off_diag_part) + static_cast<size_t>(drow++) * diag_ctas * 4;
  };
  // Denoiser calls of the loop in order: call 0 at ttab[0], then one at ttab[i+1] for every step that evaluates the
  // model. The time tables of call k+1 depend on t only: they are computed on a side stream while call k runs.
  std::vector<int> tcall(1, 0);
  for (int i = 0; i < nsteps; ++i) {
    const float* c = coef + static_cast<size_t>(i) * 8;
    const bool evals = sampler == DV_SAMPLER_DPMPP_2M ? (i != nsteps - 1) : (c[6] != 0.f);
    if (evals) tcall.push_back(i + 1);
  }
  static const bool ahead = [] { const char* e = getenv("DV_TABLES_AHEAD"); return !(e && e[0] == '0'); }();
  if (ahead) {
    if (!side_stream) DV_CK(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
    if (!ev_tfork) DV_CK(cudaEventCreateWithFlags(&ev_tfork, cudaEventDisableTiming));
    if (!ev_tjoin) DV_CK(cudaEventCreateWithFlags(&ev_tjoin, cudaEventDisableTiming));
  }
  size_t kcall = 0;
  int rc = DV_OK;
  // model(x, call k) -> dst. Tables of call k are in slot k & 1 (call 0: computed inline first).
  auto model = [&](const float* xin, float* dst) -> int {
    if (!ahead) return forward(xin, 1, ttab + tcall[kcall++], 0, dst, s, err, launches);
    const int slot = static_cast<int>(kcall & 1);
    if (kcall == 0) {
      int r0 = enqueue_time_tables(ttab + tcall[0], 0, 0, s, err);
      if (r0 != DV_OK) return r0;
    } else {
      DV_CK(cudaStreamWaitEvent(s, ev_tjoin, 0));            // tables of this call (side stream) are complete
    }
    if (kcall + 1 < tcall.size()) {                         // fork: next call's tables into the other slot
      DV_CK(cudaEventRecord(ev_tfork, s));                   // after the previous call's last reader of that slot
      DV_CK(cudaStreamWaitEvent(side_stream, ev_tfork, 0));
      int r1 = enqueue_time_tables(ttab + tcall[kcall + 1], 0, slot ^ 1, side_stream, err);
      if (r1 != DV_OK) return r1;
      DV_CK(cudaEventRecord(ev_tjoin, side_stream));
    }
    ++kcall;
    return forward(xin, 1, nullptr, 0, dst, s, err, launches, nullptr, slot);
  };
  rc = model(X, H[cur]);
  if (rc != DV_OK) return rc;
  for (int i = 0; i < nsteps; ++i) {
    const float* c = coef + static_cast<size_t>(i) * 8;
    const bool last = (i == nsteps - 1);
    if (sampler == DV_SAMPLER_DPMPP_2M) {
      const bool has_prev = c[3] != 0.f;
      DV_CK(launch_axpbypcz(X, X, H[cur], has_prev ? H[cur ^ 1] : nullptr, c[0], c[1], c[2], rt ? 1 : 0, c[4], c[5], n, s, diag()));
      ++*launches;
      if (!last) {
        cur ^= 1;
        rc = model(X, H[cur]);
        if (rc != DV_OK) return rc;
      }
    } else {
      const bool has_prev = c[5] != 0.f, use_corr = c[6] != 0.f;
      if (use_corr) {
        if (last) { err = "UniPC table: corrector on the last step"; return DV_ERR_INVALID; }
        DV_CK(launch_unipc_pred(XP, BASE, X, H[cur], H[cur ^ 1], c[0], c[1], c[2], c[3], has_prev ? 1 : 0, n, s, diag()));
        ++*launches;
        rc = model(XP, H[cur ^ 1]);  // m_t overwrites the consumed m_{-2}
        if (rc != DV_OK) return rc;
        DV_CK(launch_unipc_corr(X, BASE, H[cur ^ 1], H[cur], c[4], n, s, diag()));
        ++*launches;
        cur ^= 1;
      } else {
        DV_CK(launch_unipc_pred(X, nullptr, X, H[cur], H[cur ^ 1], c[0], c[1], c[2], c[3], has_prev ? 1 : 0, n, s, diag()));
        ++*launches;
      }
    }
  }
  if (diag_rows) {
    if (drow) {
      DV_CK(launch_diag_reduce(reinterpret_cast<const float*>(ws + off_diag_part), reinterpret_cast<float*>(ws + off_diag_rows), drow,
                               diag_ctas, s));
      ++*launches;
    }
    *diag_rows = drow;
  }
  return DV_OK;
}

int Plan::sample(int sampler, int nsteps, const float* t_in, const float* coef, float* x, int flags, cudaStream_t st,
                 std::string& err) {
  Plan* one[1] = {this};
  float* xs[1] = {x};
  return sample_group(one, 1, sampler, nsteps, t_in, coef, xs, flags, st, err);
}

// The whole loop of `n` plans ("lanes": disjoint slices of one batch of independent utterances, same T / Tp / sampler)
// in ONE CUDA graph: lane i runs on its own captured stream forked from / joined to lane 0's, so the GPU overlaps one
// lane's launch latencies, tails and MUFU-bound attention with the other lane's tensor-bound GEMMs. The graph is
// cached on plans[0].
int sample_group(Plan* const* plans, int n, int sampler, int nsteps, const float* t_in, const float* coef,
                 float* const* x, int flags, cudaStream_t st, std::string& err) {
  if (n < 1 || n > 16 || !plans || !x) { err = "bad lane group"; return DV_ERR_INVALID; }
  if (nsteps < 1 || nsteps > 4096) { err = "nsteps out of range"; return DV_ERR_INVALID; }
  if (sampler != DV_SAMPLER_DPMPP_2M && sampler != DV_SAMPLER_UNIPC_BH2) { err = "unknown sampler"; return DV_ERR_INVALID; }
  if (!t_in || !coef) { err = "null argument"; return DV_ERR_INVALID; }
  for (int i = 0; i < n; ++i) {
    Plan* p = plans[i];
    if (!p || !x[i]) { err = "null lane"; return DV_ERR_INVALID; }
    if (!p->bound || !p->prompt_ready) { err = "plan not bound / prompt not prepared"; return DV_ERR_STATE; }
    if (p->cfg.x_channels != p->cfg.out_channels) { err = "sampler needs x_channels == out_channels"; return DV_ERR_INVALID; }
  }
  const bool rt = (flags & DV_SAMPLE_ROUNDTRIP) != 0;
  const bool dg = (flags & DV_SAMPLE_NO_DIAG) == 0;
  Plan* lead = plans[0];
  for (int i = 0; i < n; ++i) {
    Plan* p = plans[i];
    const size_t nel = static_cast<size_t>(p->B) * p->cfg.x_channels * p->T;
    DV_CK(cudaMemcpyAsync(p->ws + p->off_ttab, t_in, static_cast<size_t>(nsteps) * 4, cudaMemcpyHostToDevice, st));
    DV_CK(cudaMemcpyAsync(p->ws + p->off_x, x[i], nel * 4, cudaMemcpyDeviceToDevice, st));
  }
  // The loop of a (plans, sampler, table) combination runs eagerly the FIRST time it is seen and is captured into a CUDA
  // graph when it comes back: capture + instantiate of ~16 k kernel nodes costs ~45 ms, as much as a third of a one-shot
  // synthesis (tts_infer.py sees a new T per call), while eager enqueue is as fast as a replay at those sizes and is
  // bit-identical (tests). DV_GRAPH_FIRST=1 captures at first sight.
  uint64_t key = 1469598103934665603ull;
  bool eager = (flags & DV_SAMPLE_NO_GRAPH) != 0;
  if (!eager) {
    key = fnv1a(key, &sampler, sizeof sampler);
    key = fnv1a(key, &nsteps, sizeof nsteps);
    key = fnv1a(key, &rt, sizeof rt);
    key = fnv1a(key, &dg, sizeof dg);
    { const int sk = skip_kinds_mask(); key = fnv1a(key, &sk, sizeof sk); }
    key = fnv1a(key, coef, static_cast<size_t>(nsteps) * 8 * 4);
    key = fnv1a(key, plans, static_cast<size_t>(n) * sizeof(Plan*));
    for (int i = 0; i < n; ++i) key = fnv1a(key, &plans[i]->graph_epoch, sizeof(uint64_t));
    const bool have = lead->graph_exec && key == lead->graph_key;
    const char* gf = getenv("DV_GRAPH_FIRST");
    if (!have && key != lead->pending_key && !(gf && gf[0] == '1')) {
      lead->pending_key = key;
      eager = true;
    }
  }
  if (eager) {
    int launches = 0;
    for (int i = 0; i < n; ++i) {
      int rows = 0;
      int rc = plans[i]->enqueue_loop(sampler, nsteps, coef, rt, st, &launches, err, dg ? &rows : nullptr);
      if (rc != DV_OK) return rc;
      plans[i]->last_diag_rows = rows;
    }
    lead->last_sample_launches = launches;
  } else {
    if (!lead->graph_exec || key != lead->graph_key) {
      if (lead->graph_exec) { cudaGraphExecDestroy(lead->graph_exec); lead->graph_exec = nullptr; }
      for (int i = 0; i < n; ++i) {
        Plan* p = plans[i];
        if (!p->cap_stream) DV_CK(cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking));
        if (!p->ev_fork) DV_CK(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
        if (!p->ev_join) DV_CK(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
      }
      cudaStream_t s0 = lead->cap_stream;
      DV_CK(cudaStreamBeginCapture(s0, cudaStreamCaptureModeThreadLocal));
      int launches = 0;
      int rc = DV_OK;
      cudaError_t ce = cudaSuccess;
      if (n > 1) ce = cudaEventRecord(lead->ev_fork, s0);
      for (int i = 1; i < n && ce == cudaSuccess; ++i) ce = cudaStreamWaitEvent(plans[i]->cap_stream, lead->ev_fork, 0);
      for (int i = 0; i < n && rc == DV_OK && ce == cudaSuccess; ++i)
        rc = plans[i]->enqueue_loop(sampler, nsteps, coef, rt, plans[i]->cap_stream, &launches, err, dg ? &plans[i]->graph_diag_rows : nullptr);
      for (int i = 1; i < n && ce == cudaSuccess; ++i) {   // always join, also on failure, so the capture can end
        ce = cudaEventRecord(plans[i]->ev_join, plans[i]->cap_stream);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(s0, plans[i]->ev_join, 0);
      }
      cudaGraph_t graph = nullptr;
      cudaError_t ce2 = cudaStreamEndCapture(s0, &graph);
      if (rc != DV_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (ce != cudaSuccess || ce2 != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        err = std::string("graph capture: ") + cudaGetErrorString(ce != cudaSuccess ? ce : ce2);
        return DV_ERR_CUDA;
      }
      ce = cudaGraphInstantiate(&lead->graph_exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) { lead->graph_exec = nullptr; err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce); return DV_ERR_CUDA; }
      lead->graph_key = key;
      lead->graph_launches = launches;
    }
    DV_CK(cudaGraphLaunch(lead->graph_exec, st));
    lead->last_sample_launches = lead->graph_launches;
    for (int i = 0; i < n; ++i) plans[i]->last_diag_rows = dg ? plans[i]->graph_diag_rows : 0;
  }
  for (int i = 0; i < n; ++i) {
    Plan* p = plans[i];
    const size_t nel = static_cast<size_t>(p->B) * p->cfg.x_channels * p->T;
    DV_CK(cudaMemcpyAsync(x[i], p->ws + p->off_x, nel * 4, cudaMemcpyDeviceToDevice, st));
  }
  return DV_OK;
}

Plan::~Plan() {
  if (graph_exec) cudaGraphExecDestroy(graph_exec);
  if (cap_stream) cudaStreamDestroy(cap_stream);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_join) cudaEventDestroy(ev_join);
  if (side_stream) cudaStreamDestroy(side_stream);
  if (ev_tfork) cudaEventDestroy(ev_tfork);
  if (ev_tjoin) cudaEventDestroy(ev_tjoin);
}

}  // namespace dv
