// extern "C" surface of libdiffvits_b200.so (declared in include/diffvits_b200.h).
#include <cstring>
#include <new>
#include <string>

#include "engine.h"
#ifdef DV_DEBUG_HOOKS
#include "../../include/diffvits_b200_debug.h"
#endif

struct dv_plan {
  dv::Plan impl;
};

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
int ck(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return DV_OK;
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return DV_ERR_CUDA;
}
cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

// Every entry point that takes a plan runs on the plan's device (the one its workspace lives on, recorded by dv_plan_bind),
// whatever device is current in the calling thread; the caller's current device is restored on return.
struct DevGuard {
  int prev = -1;
  bool switched = false;
  explicit DevGuard(int dev) {
    if (dev < 0) return;
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DevGuard() { if (switched) cudaSetDevice(prev); }
};

// There is no CPU path: refuse loudly when no sm_100 device is current.
int require_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(DV_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return fail(DV_ERR_CUDA, cudaGetErrorString(e));
  if (major != 10) return fail(DV_ERR_UNSUPPORTED, "diffvits_b200 kernels are built for sm_100a only (compute capability 10.x required)");
  return DV_OK;
}
}  // namespace

extern "C" {

int dv_version(void) { return DV_VERSION; }
const char* dv_last_error(void) { return g_err.c_str(); }

int dv_plan_create(const dv_unet_cfg* cfg, int B, int T, int Tp, dv_plan** out) {
  if (!cfg || !out) return fail(DV_ERR_INVALID, "null argument");
  dv_plan* p = new (std::nothrow) dv_plan();
  if (!p) return fail(DV_ERR_INVALID, "out of host memory");
  p->impl.cfg = *cfg;
  p->impl.B = B; p->impl.T = T; p->impl.Tp = Tp;
  std::string err;
  int rc = p->impl.build(err);
  if (rc != DV_OK) {
    delete p;
    return fail(rc, err);
  }
  *out = p;
  return DV_OK;
}
void dv_plan_destroy(dv_plan* plan) {
  if (!plan) return;
  DevGuard guard(plan->impl.device);   // streams / events / graph were created on the plan's device
  delete plan;
}
size_t dv_plan_workspace_bytes(const dv_plan* plan) { return plan ? plan->impl.ws_bytes : 0; }
size_t dv_plan_param_bytes(const dv_plan* plan) { return plan ? plan->impl.param_bytes : 0; }
const char* dv_plan_manifest_json(const dv_plan* plan) { return plan ? plan->impl.manifest.c_str() : ""; }
int dv_plan_launches_per_forward(const dv_plan* plan) { return plan ? static_cast<int>(plan->impl.ops.size()) + 3 : 0; }
int dv_plan_last_sample_launches(const dv_plan* plan) { return plan ? plan->impl.last_sample_launches : 0; }
int dv_plan_sample_diag(const dv_plan* plan, size_t* offset, int* rows) {
  if (!plan || !offset || !rows) return fail(DV_ERR_INVALID, "null argument");
  *offset = plan->impl.off_diag_rows;
  *rows = plan->impl.last_diag_rows;
  return DV_OK;
}
int dv_plan_health_flag(const dv_plan* plan, size_t* offset) {
  if (!plan || !offset) return fail(DV_ERR_INVALID, "null argument");
  *offset = plan->impl.off_gncnt + static_cast<size_t>(plan->impl.B) * 4;
  return DV_OK;
}

int dv_plan_bind(dv_plan* plan, void* workspace, const void* params, void* stream) {
  if (!plan) return fail(DV_ERR_INVALID, "null plan");
  {
    cudaPointerAttributes at;
    cudaError_t pe = cudaPointerGetAttributes(&at, workspace);
    if (pe != cudaSuccess || at.type != cudaMemoryTypeDevice) {
      cudaGetLastError();
      return fail(DV_ERR_INVALID, "workspace is not a device allocation");
    }
    plan->impl.device = at.device;
  }
  DevGuard guard(plan->impl.device);
  int rc = require_device();
  if (rc != DV_OK) return rc;
  std::string err;
  rc = plan->impl.bind(workspace, params, S(stream), err);
  return rc == DV_OK ? rc : fail(rc, err);
}
int dv_prompt_prepare(dv_plan* plan, const float* ehs, const float* aug, const int32_t* kv_len, void* stream) {
  if (!plan) return fail(DV_ERR_INVALID, "null plan");
  DevGuard guard(plan->impl.device);
  std::string err;
  int rc = plan->impl.prompt_prepare(ehs, aug, kv_len, S(stream), err);
  return rc == DV_OK ? rc : fail(rc, err);
}
int dv_set_cond(dv_plan* plan, const float* cond, void* stream) {
  if (!plan) return fail(DV_ERR_INVALID, "null plan");
  DevGuard guard(plan->impl.device);
  std::string err;
  int rc = plan->impl.set_cond(cond, S(stream), err);
  return rc == DV_OK ? rc : fail(rc, err);
}
int dv_unet_forward(dv_plan* plan, const float* sample, int x_only, const float* t, int t_stride, float* out,
                    void* stream) {
  if (!plan || !sample || !t || !out) return fail(DV_ERR_INVALID, "null argument");
  DevGuard guard(plan->impl.device);
  std::string err;
  int launches = 0;
  int rc = plan->impl.forward(sample, x_only, t, t_stride, out, S(stream), err, &launches);
  return rc == DV_OK ? rc : fail(rc, err);
}
int dv_sample(dv_plan* plan, int sampler, int nsteps, const float* t_in, const float* coef, float* x, int flags,
              void* stream) {
  if (!plan) return fail(DV_ERR_INVALID, "null plan");
  DevGuard guard(plan->impl.device);
  std::string err;
  int rc = plan->impl.sample(sampler, nsteps, t_in, coef, x, flags, S(stream), err);
  return rc == DV_OK ? rc : fail(rc, err);
}
int dv_sample_group(dv_plan* const* plans, int nplans, int sampler, int nsteps, const float* t_in, const float* coef,
                    float* const* x, int flags, void* stream) {
  if (!plans || nplans < 1 || nplans > 16) return fail(DV_ERR_INVALID, "bad lane group");
  dv::Plan* ps[16];
  for (int i = 0; i < nplans; ++i) {
    if (!plans[i]) return fail(DV_ERR_INVALID, "null plan");
    ps[i] = &plans[i]->impl;
    if (ps[i]->device != ps[0]->device) return fail(DV_ERR_INVALID, "lanes of one group must live on one device");
  }
  DevGuard guard(ps[0]->device);
  std::string err;
  int rc = dv::sample_group(ps, nplans, sampler, nsteps, t_in, coef, x, flags, S(stream), err);
  return rc == DV_OK ? rc : fail(rc, err);
}

int dv_plan_tensor_count(const dv_plan* plan) { return plan ? static_cast<int>(plan->impl.named.size()) : 0; }
int dv_plan_tensor_info(const dv_plan* plan, int i, char* name, int name_cap, size_t* offset, int* L, int* C) {
  if (!plan || i < 0 || i >= static_cast<int>(plan->impl.named.size())) return fail(DV_ERR_INVALID, "tensor index out of range");
  const auto& t = plan->impl.named[i];
  if (name && name_cap > 0) {
    strncpy(name, t.first.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (offset) *offset = t.second.off;
  if (L) *L = t.second.L;
  if (C) *C = t.second.C;
  return DV_OK;
}
int dv_plan_op_count(const dv_plan* plan) { return plan ? static_cast<int>(plan->impl.ops.size()) : 0; }
int dv_plan_op_info(const dv_plan* plan, int i, char* tag, int tag_cap, int* kind, double* flops, double* bytes) {
  if (!plan || i < 0 || i >= static_cast<int>(plan->impl.ops.size())) return fail(DV_ERR_INVALID, "op index out of range");
  const dv::Op& o = plan->impl.ops[i];
  if (tag && tag_cap > 0) {
    strncpy(tag, o.tag.c_str(), tag_cap - 1);
    tag[tag_cap - 1] = 0;
  }
  if (kind) *kind = static_cast<int>(o.kind);
  if (flops) *flops = o.flops;
  if (bytes) *bytes = o.bytes;
  return DV_OK;
}
int dv_plan_profile_forward(dv_plan* plan, const float* sample, int x_only, const float* t, int t_stride, float* out,
                            float* ms, int cap, void* stream) {
  if (!plan || !sample || !t || !out || !ms) return fail(DV_ERR_INVALID, "null argument");
  DevGuard guard(plan->impl.device);
  std::string err;
  int rc = plan->impl.profile_forward(sample, x_only, t, t_stride, out, S(stream), ms, cap, err);
  return rc == DV_OK ? rc : fail(rc, err);
}

// ------------------------------------------------------------------------------------------ prompt encoder
}  // extern "C"
struct dv_prompt_plan {
  dv::PromptPlan impl;
};
extern "C" {
int dv_prompt_plan_create(const dv_prompt_cfg* cfg, int B, int Tp, dv_prompt_plan** out) {
  if (!cfg || !out) return fail(DV_ERR_INVALID, "null argument");
  dv_prompt_plan* p = new (std::nothrow) dv_prompt_plan();
  if (!p) return fail(DV_ERR_INVALID, "out of host memory");
  p->impl.cfg = *cfg;
  p->impl.B = B; p->impl.Tp = Tp;
  std::string err;
  int rc = p->impl.build(err);
  if (rc != DV_OK) {
    delete p;
    return fail(rc, err);
  }
  *out = p;
  return DV_OK;
}
void dv_prompt_plan_destroy(dv_prompt_plan* plan) { delete plan; }
size_t dv_prompt_plan_workspace_bytes(const dv_prompt_plan* plan) { return plan ? plan->impl.ws_bytes : 0; }
size_t dv_prompt_plan_param_bytes(const dv_prompt_plan* plan) { return plan ? plan->impl.param_bytes : 0; }
const char* dv_prompt_plan_manifest_json(const dv_prompt_plan* plan) { return plan ? plan->impl.manifest.c_str() : ""; }
int dv_prompt_plan_bind(dv_prompt_plan* plan, void* workspace, const void* params, void* stream) {
  if (!plan) return fail(DV_ERR_INVALID, "null plan");
  {
    cudaPointerAttributes at;
    cudaError_t pe = cudaPointerGetAttributes(&at, workspace);
    if (pe != cudaSuccess || at.type != cudaMemoryTypeDevice) {
      cudaGetLastError();
      return fail(DV_ERR_INVALID, "workspace is not a device allocation");
    }
    plan->impl.device = at.device;
  }
  DevGuard guard(plan->impl.device);
  int rc = require_device();
  if (rc != DV_OK) return rc;
  std::string err;
  rc = plan->impl.bind(workspace, params, S(stream), err);
  return rc == DV_OK ? rc : fail(rc, err);
}
int dv_prompt_encode(dv_prompt_plan* plan, const float* prompt, const int32_t* lengths, float* ehs_out, void* stream) {
  if (!plan) return fail(DV_ERR_INVALID, "null plan");
  DevGuard guard(plan->impl.device);
  std::string err;
  int rc = plan->impl.encode(prompt, lengths, ehs_out, S(stream), err);
  return rc == DV_OK ? rc : fail(rc, err);
}

// ------------------------------------------------------------------------------------------ single kernels
#ifdef DV_DEBUG_HOOKS
// Debug library only (include/diffvits_b200_debug.h): timing ablations and CTA timelines.
static thread_local unsigned long long* g_test_trace = nullptr;
static thread_local int g_test_flags = 0;
void dv_gemm_test_trace(unsigned long long* buf) { g_test_trace = buf; }
void dv_gemm_test_flags(int flags) { g_test_flags = flags; }
int dv_debug_skip_kinds(int mask) { int& m = dv::skip_kinds_mask(); const int old = m; m = mask; return old; }
// CTA timeline of op `op` of the plan's denoiser call (tap-GEMM ops only): every later launch of that op records into `buf`
int dv_debug_trace_op(dv_plan* plan, int op, unsigned long long* buf) {
  if (!plan || op < 0 || op >= static_cast<int>(plan->impl.ops.size()) || plan->impl.ops[op].kind != dv::OP_GEMM) return DV_ERR_INVALID;
  plan->impl.ops[op].gp.trace = buf;
  return DV_OK;
}
#endif

int dv_gemm(const dv_gemm_src* srcs, int nsrc, const dv_gemm_seg* segs, int nseg, const void* w, int k_tot, int n,
            int n_pad, const float* bias, const void* res, int ld_res, void* out, int ld_out, int out_mode, int B,
            int L_out, int impl, const dv_gemm_opts* opts, void* stream) {
  if (!srcs || !segs || !w || !out) return fail(DV_ERR_INVALID, "null argument");
  if (nsrc < 1 || nsrc > dv::GEMM_MAX_LSRC || nseg < 1 || nseg > dv::GEMM_MAX_LSEG) return fail(DV_ERR_INVALID, "bad nsrc/nseg");
  if (n_pad % 128 || k_tot % 64) return fail(DV_ERR_INVALID, "n_pad must be a multiple of 128 and k_tot of 64");
  int rc = require_device();
  if (rc != DV_OK) return rc;
  dv::GemmParams p;
  memset(&p, 0, sizeof p);
  std::string err;
  dv::GemmRefSrc rs[dv::GEMM_MAX_SRC];
  for (int i = 0; i < dv::GEMM_MAX_SRC; ++i) {
    const int si = i < nsrc ? i : 0;
    const dv_gemm_src& s = srcs[si];
    if (s.C % 64) return fail(DV_ERR_INVALID, "source channels must be a multiple of 64");
    int rows = 128;                                        // sources read with shared conv taps stage a 136-row box
    for (int k = 0; k < nseg; ++k) if (segs[k].src == si && segs[k].ntap == 3) rows = 136;
    p.map_rows[i] = static_cast<int16_t>(rows);
    if (!dv::make_tmap_3d(&p.mapA[i], s.ptr, s.C, s.ld, s.L, B, s.bs > 0 ? static_cast<size_t>(s.bs) : static_cast<size_t>(s.L) * s.ld, err, rows))
      return fail(DV_ERR_CUDA, err);
    rs[i] = dv::GemmRefSrc{s.ptr, s.L, s.C, s.ld};
  }
  if (!dv::make_tmap_2d(&p.mapW, w, k_tot, n_pad, 64, err)) return fail(DV_ERR_CUDA, err);
  p.mapW2 = p.mapW;
  int ksum = 0;
  for (int i = 0; i < nseg; ++i) {
    if (segs[i].src < 0 || segs[i].src >= nsrc) return fail(DV_ERR_INVALID, "segment source out of range");
    p.seg[i].src = static_cast<int16_t>(segs[i].src);
    p.seg[i].row_off = static_cast<int16_t>(segs[i].row_off);
    p.seg[i].c0 = static_cast<int16_t>(segs[i].c0);
    p.seg[i].npanels = static_cast<int16_t>(segs[i].npanels);
    p.seg[i].wsel = 0;
    const int ntap = segs[i].ntap == 3 ? 3 : 1;
    if (segs[i].ntap != 0 && segs[i].ntap != 1 && segs[i].ntap != 3) return fail(DV_ERR_INVALID, "ntap must be 1 or 3");
    p.seg[i].ntap = static_cast<int16_t>(ntap);
    p.seg[i].wk0 = static_cast<int16_t>(segs[i].wk0 >= 0 ? segs[i].wk0 : ksum / 64);
    p.seg[i].tap_stride = static_cast<int16_t>(segs[i].tap_stride > 0 ? segs[i].tap_stride : segs[i].npanels);
    p.seg[i].xf = static_cast<int16_t>(segs[i].xf);
    p.seg[i].xf_c0 = static_cast<int16_t>(segs[i].xf_c0);
    if (segs[i].xf && !(opts && opts->norm)) return fail(DV_ERR_INVALID, "xf segment without dv_gemm_opts.norm");
    ksum += segs[i].npanels * 64 * ntap;
  }
  if (ksum != k_tot) return fail(DV_ERR_INVALID, "segments do not add up to k_tot");
  p.nseg = nseg;
  p.B = B; p.L = L_out; p.N = n; p.n_pad = n_pad;
  p.tiles_per_batch = (L_out + 127) / 128;
  p.out_mode = out_mode;
  p.bias = bias; p.res = res; p.ld_res = ld_res; p.out = out; p.ld_out = ld_out;
  if (opts) { p.out_bs = opts->out_bs; p.trim_col = opts->trim_col; p.stats = opts->stats; p.bn = opts->bn; }
  if (opts && opts->norm) {
    const dv_gemm_norm& nm = *opts->norm;
    if (nm.nst < 1 || nm.nst > 2) return fail(DV_ERR_INVALID, "norm.nst must be 1 or 2");
    for (int i = 0; i < nm.nst; ++i) {
      p.xf.st_ptr[i] = nm.stats[i]; p.xf.st_chunks[i] = nm.chunks[i]; p.xf.st_nblk[i] = nm.nblk[i];
      p.xf.st_fold[i] = nm.fold[i]; p.xf.st_cover[i] = nm.cover[i];
    }
    p.xf.nst = nm.nst; p.xf.C = nm.C; p.xf.L = nm.L; p.xf.G = nm.G; p.xf.eps = nm.eps; p.xf.silu = nm.silu;
    p.xf.gamma = nm.gamma; p.xf.beta = nm.beta; p.xf.scale_shift = nm.scale_shift; p.xf.ss_ld = nm.ss_ld;
  }
  p.stats_nblk = n / 16;
#ifdef DV_DEBUG_HOOKS
  p.trace = g_test_trace; p.dbg_flags = g_test_flags;
#endif
  if (impl == 1) return ck(dv::launch_gemm_ref(p, rs, w, k_tot, S(stream)), "gemm_ref");
  return ck(dv::launch_gemm_tc(p, S(stream)), "gemm_tc");
}

int dv_attention(const void* q, const void* k, const void* v, void* o, int ld_q, int ld_k, int ld_v, int ld_o,
                 long long bs_q, long long bs_k, long long bs_v, long long bs_o, int B, int H, int Lq, int Lk, int D,
                 float scale, const int32_t* kv_len, void* stream) {
  int rc = require_device();
  if (rc != DV_OK) return rc;
  dv::AttnParams p;
  memset(&p, 0, sizeof p);
  p.q = q; p.k = k; p.v = v; p.o = o;
  p.ld_q = ld_q; p.ld_k = ld_k; p.ld_v = ld_v; p.ld_o = ld_o;
  p.bs_q = bs_q; p.bs_k = bs_k; p.bs_v = bs_v; p.bs_o = bs_o;
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.D = D; p.scale = scale; p.kv_len = kv_len;
#ifdef DV_DEBUG_HOOKS
  p.trace = g_test_trace;   // the dv_gemm_test_trace buffer doubles as the attention kernel's debug timeline
#endif
  const char* ae = nullptr;
  if (!dv::attention_tc_prepare(p, &ae)) return fail(DV_ERR_CUDA, ae ? ae : "attention_tc_prepare");
  return ck(dv::launch_attention(p, S(stream)), "attention");
}

int dv_groupnorm(const void* x0, int c0, const void* x1, int c1, int B, int L, int G, float* scratch,
                 const float* gamma, const float* beta, const float* scale_shift, int ss_ld, float eps, int silu,
                 void* out, void* stream) {
  int rc = require_device();
  if (rc != DV_OK) return rc;
  dv::GnParams p;
  memset(&p, 0, sizeof p);
  p.x0 = x0; p.c0 = c0; p.x1 = x1; p.c1 = x1 ? c1 : 0;
  p.B = B; p.L = L; p.G = G; p.gamma = gamma; p.beta = beta;
  p.scale_shift = scale_shift; p.ss_ld = ss_ld; p.eps = eps; p.silu = silu; p.out = out;
  p.nst = 1;
  p.st[0].ptr = scratch; p.st[0].chunks = (L + 31) / 32; p.st[0].nblk = (p.c0 + p.c1) / 16; p.st[0].fold = 0;
  p.st[0].cover = p.c0 + p.c1;
  rc = ck(dv::launch_gn_stats(p, scratch, S(stream)), "gn_stats");
  if (rc != DV_OK) return rc;
  return ck(dv::launch_gn_apply(p, S(stream)), "gn_apply");
}

size_t dv_groupnorm_scratch_floats(int B, int L, int C) { return dv::gn_scratch_floats(B, L, C); }

int dv_layernorm(const void* x, void* out, const float* gamma, const float* beta, int rows, int C, float eps,
                 void* stream) {
  int rc = require_device();
  if (rc != DV_OK) return rc;
  return ck(dv::launch_layernorm(x, out, gamma, beta, rows, C, eps, S(stream)), "layernorm");
}

int dv_dpmpp_update(float* x_out, const float* x, float* m1, const float* m0, float a, float b, float c,
                    int roundtrip, float alpha_prev, float sigma_prev, size_t n, void* stream) {
  int rc = require_device();
  if (rc != DV_OK) return rc;
  return ck(dv::launch_axpbypcz(x_out, x, m1, m0, a, b, c, roundtrip, alpha_prev, sigma_prev, n, S(stream)), "dpmpp_update");
}
int dv_unipc_predict(float* xpred, float* base, const float* x, const float* m1, const float* m0, float a, float b,
                     float cp, float cc, int has_prev, size_t n, void* stream) {
  int rc = require_device();
  if (rc != DV_OK) return rc;
  return ck(dv::launch_unipc_pred(xpred, base, x, m1, m0, a, b, cp, cc, has_prev, n, S(stream)), "unipc_predict");
}
int dv_unipc_correct(float* x, const float* base, const float* mt, const float* m1, float d, size_t n, void* stream) {
  int rc = require_device();
  if (rc != DV_OK) return rc;
  return ck(dv::launch_unipc_corr(x, base, mt, m1, d, n, S(stream)), "unipc_correct");
}
// ---- multi-GPU gather (csrc/comm.cu)
int dv_comm_unique_id(const char* nccl_path, void* id128) {
  if (!id128) return fail(DV_ERR_INVALID, "null id buffer");
  std::string err;
  int rc = dv::comm_unique_id(nccl_path, id128, err);
  return rc == DV_OK ? rc : fail(rc, err);
}
int dv_comm_init(const char* nccl_path, const void* id128, int rank, int nranks, dv_comm** out) {
  std::string err;
  dv::Comm* c = nullptr;
  int rc = dv::comm_init(nccl_path, id128, rank, nranks, &c, err);
  if (rc != DV_OK) return fail(rc, err);
  *out = reinterpret_cast<dv_comm*>(c);
  return DV_OK;
}
int dv_comm_all_gather(dv_comm* comm, const float* send, float* recv, size_t count, void* stream) {
  std::string err;
  int rc = dv::comm_all_gather(reinterpret_cast<dv::Comm*>(comm), send, recv, count, S(stream), err);
  return rc == DV_OK ? rc : fail(rc, err);
}
void dv_comm_destroy(dv_comm* comm) { dv::comm_destroy(reinterpret_cast<dv::Comm*>(comm)); }

int dv_nct_to_ntc(const float* src, void* dst, int B, int C, int T, int ld, int col0, void* stream) {
  int rc = require_device();
  if (rc != DV_OK) return rc;
  return ck(dv::launch_nct_to_ntc(src, dst, B, C, T, ld, col0, 1 << 30, 0, S(stream)), "nct_to_ntc");
}

}  // extern "C"
