// Plan: the UNet1DConditionModel forward pass at a fixed (B, T, Tp), flattened at build time into a linear
// program of kernel launches over one caller-owned workspace. Mirrors the module order of
// unet1d/unet_1d_condition.py:743-1037 with step-invariant work hoisted.
#pragma once
#include <string>
#include <vector>

#include "../../include/diffvits_b200.h"
#include "kernels.h"

namespace dv {

// The launch functions of one operand format (kernels_api.inc): BF16 (namespace dv) or FP16 (namespace dv16).
struct KernelApi {
  cudaError_t (*gemm_tc_init)();
  cudaError_t (*chain_tc_init)();
  cudaError_t (*attention_tc_init)();
  bool (*attention_tc_prepare)(AttnParams&, const char**);
  cudaError_t (*launch_gemm_tc)(const GemmParams&, cudaStream_t);
  cudaError_t (*launch_chain_tc)(const ChainParams&, cudaStream_t);
  cudaError_t (*launch_attention)(const AttnParams&, cudaStream_t);
  cudaError_t (*launch_gn_apply)(const GnParams&, cudaStream_t);
  cudaError_t (*launch_layernorm)(const void*, void*, const float*, const float*, int, int, float, cudaStream_t, const void*, void*);
  cudaError_t (*launch_nct_to_ntc)(const float*, void*, int, int, int, int, int, int, int, cudaStream_t, void*);
  cudaError_t (*launch_f32_to_bf16)(const float*, void*, size_t, cudaStream_t, void*);
};
const KernelApi& kernel_api(bool f16);

struct TRef {  // 16-bit channel-last tensor [B, L, C] at byte offset `off` of the workspace
  size_t off = 0;
  int L = 0, C = 0;
  bool lo = false;   // followed by its lo plane (same shape, at + B*L*C*2 bytes): FP32X3 every tensor, FP16S the residual stream
  // GroupNorm partial statistics written by the producing GEMM's epilogue (offset inside the stats region)
  long long st_off = -1;
  int st_chunks = 0, st_nblk = 0, st_fold = 0;
  int prod = -1;     // index in Plan::ops of the tap-GEMM that writes this tensor (-1: another kind of producer)
  bool valid() const { return C > 0; }
};

struct ParamSlot {
  std::string name;               // engine-side slot name
  std::string recipe;             // packing recipe implemented by the host packer
  std::vector<std::string> keys;  // reference state_dict keys consumed
  std::string dtype;              // "bf16" | "f32"
  std::vector<int> shape;
  size_t off = 0, bytes = 0;
};

enum OpKind { OP_GEMM = 0, OP_GN_APPLY = 2, OP_LN = 3, OP_ATTN = 4, OP_CHAIN = 5 };
int& skip_kinds_mask();   // measurement hook (dv_debug_skip_kinds): op kinds that are not launched

// GroupNorm fused into a GEMM's operand path (GemmSeg::xf): the normalised tensor is (x0 | x1), never materialised
struct XfDesc {
  bool on = false;
  TRef x0, x1;
  long long gamma = -1, beta = -1, ss = -1;   // parameter slots / column offset into the time scale-shift table
  float eps = 1e-5f;
  bool silu = false;
};

// GroupNorm of a GEMM's output applied by the producing kernel (kernels.h: GemmGnOut) instead of a stand-alone gn_apply launch
struct GnOutDesc {
  bool on = false;
  TRef out;                                   // the normalised tensor (what the consumer GEMM reads)
  long long gamma = -1, beta = -1, ss = -1;   // parameter slots / column offset into the time scale-shift table
  float eps = 1e-5f;
  bool silu = false;
  std::string tag;                            // reference module path of the norm
};

struct GemmDesc {   // logical description: bind() expands it to the kernel's maps / segments (x3 in split precision)
  int nsrc = 0;
  TRef src[GEMM_MAX_LSRC];
  int view_l[GEMM_MAX_LSRC], view_c[GEMM_MAX_LSRC];  // TMA view (stride-2 conv reads [L/2, 2C])
  int nseg = 0;
  GemmSeg seg[GEMM_MAX_LSEG];
  size_t w_off = 0; int k_tot = 0, n = 0, n_pad = 0;
  long long bias_off = -1;
  TRef res; TRef out;
  int l_out = 0, out_mode = GEMM_OUT_BF16, ld_out = 0;
  long long out_bs = 0;      // elements between batch items of the output (0 = dense l_out * ld_out)
  int trim_col = 0;          // > 0: columns >= trim_col of the last GEMM row do not exist (odd-length upsample)
  int alg_k = 0, alg_nsrc = 0;  // roofline accounting: the reference's K / distinct inputs when they differ from k_tot / nsrc
  bool out_dynamic = false;  // conv_out: destination supplied per call
  long long stats_off = -1;  // >= 0: emit GroupNorm partials of the output
  int stats_nblk = 0;
  XfDesc xf;
  GnOutDesc gnout;
};

// One stage of a chained-GEMM launch (kernels.h: ChainStage), offsets instead of pointers until bind()
struct ChainStageDesc {
  int kind = CH_ROW;
  TRef a;                          // CH_ROW: A operand streamed from global (invalid = X buffer)
  size_t w_off = 0, w2_off = 0;    // packed weights [n_pad, k] / [n_pad2, k2]
  int w_rows = 0, w_k = 0, w2_rows = 0, w2_k = 0;
  int n = 0;
  long long bias_off = -1, bias2_off = -1;
  TRef res, out;
  TRef out2;                       // CH_ROW: LayerNorm result to global instead of the X buffer
  long long gamma = -1, beta = -1;
  float eps = 1e-5f;
  long long stats_off = -1;
  int stats_nblk = 0;
};
struct ChainDesc {
  int nstages = 0;
  ChainStageDesc st[CHAIN_MAX_STAGES];
  int L = 0, C = 0;
};

struct Op {
  OpKind kind;
  std::string tag;
  double flops = 0, bytes = 0;  // algorithmic work of one launch (for roofline accounting)
  GemmDesc g; GemmParams gp;
  ChainDesc ch; ChainParams cp;
  GnParams gn;
  struct { const void* x; void* out; const float* gamma; const float* beta; int rows, C; float eps; const void* x_lo; void* out_lo; } ln;
  AttnParams at;
  // deferred (offset-based) descriptions resolved in bind()
  struct { size_t x0, x1, out; long long st[2]; long long gamma, beta, ss; bool x_lo, out_lo; } gno;
  struct { size_t x, out; long long gamma, beta; bool x_lo, out_lo; } lno;
  struct { size_t q, k, v, o; bool cross; } ato;
};

struct Plan {
  dv_unet_cfg cfg;
  int B, T, Tp;
  int n_levels, emb_dim, cin_pad;
  bool split = false;        // FP32X3 / FP16S: tensors may carry a lo plane (TRef::lo): FP32X3 all of them, FP16S the residual stream
                             // (block outputs, the transformer's z): epilogues and norms read hi + lo where it exists
  bool triple = false;       // FP32X3 only: weights carry a lo matrix too and every product is three MMAs (hi.hi + hi.lo + lo.hi);
                             // attention core in FP32. FP16S: one MMA on the hi planes (FP16 operands), tensor-core attention
  bool f16 = false;          // FP16S: the dv16 kernel set (FP16 operand format)
  const KernelApi* K = nullptr;
  size_t plane(const TRef& t) const { return static_cast<size_t>(B) * t.L * t.C * 2; }
  // Widest level whose transformer blocks run as chained kernels (DV_FUSE_MAXC). Measured on C2 (B200, 32 x 1000):
  // the chain wins or ties while a level has >= ~1.5 row tiles per SM (C = 128: 250 tiles); deeper levels have too few
  // 128-row tiles (125 / 63) to fill 148 SMs with one CTA per tile, there the one-kernel-per-op program is faster.
  int fuse_max_c = 128;
  bool fuse_ln = false;      // DV_FUSE_LN=1: projection + following LayerNorm as a one-stage chain where C <= 384. Measured
                             // slower on C2 (4.62 vs 4.34 ms per call: one 128-row tile per CTA with a three-pass epilogue loses to
                             // an N-split GEMM + a streaming LayerNorm), so it is off; kept parity-tested for wider-N shapes.
  bool fuse = true;          // transformer sections as chained-GEMM kernels (DV_FUSE=0 selects the one-kernel-per-op program)
  bool gemm2 = false;        // DV_GEMM2=1: k=3 convs as ONE shared-tap segment on the two-ring kernel (csrc/gemm_tc2.cu) instead of
                             // three single-tap segments on the one-ring kernel (measured 4 % slower per C2 step: off)
  bool gn_out = true;        // single-source GroupNorms applied by the PRODUCING GEMM's epilogue (per-batch-item arrival counter, GemmGnOut):
                             // 47 of the 61 gn_apply launches of a call disappear; DV_GN_OUT=0 keeps them stand-alone. Default in the single-plane modes; parity-tested in all.
  bool fuse_gn = false;      // DV_FUSE_GN=1 (implies gemm2): GroupNorm (+SiLU, +scale/shift) applied inside the consuming conv / 1x1 GEMM (operand transform in
                             // shared memory) instead of a stand-alone gn_apply pass. Parity-tested; measured 13 % slower per C2 step (gemm_tc2.cu header): off
  std::vector<ParamSlot> slots;
  size_t param_bytes = 0, ws_bytes = 0;
  std::vector<Op> ops;       // one denoiser call
  std::vector<Op> prep_ops;  // prompt-side hoisted projections
  std::string manifest;
  std::vector<std::pair<std::string, TRef>> named;  // block outputs that stay intact after a call (debug/tests)

  // workspace regions (byte offsets)
  size_t off_xin = 0, off_ehs = 0, off_aug = 0, off_embact = 0, off_ss = 0, off_stats = 0, off_kvlen = 0;
  size_t off_x = 0, off_xp = 0, off_base = 0, off_h0 = 0, off_h1 = 0, off_ttab = 0;
  size_t off_gncnt = 0;      // [B] arrival counters of the fused output GroupNorm (zero between launches)
  size_t off_diag_part = 0, off_diag_rows = 0;   // sampler diagnostics: per-CTA partials [DIAG_MAX_ROWS][diag_ctas][4], folded rows [DIAG_MAX_ROWS][4]
  int diag_ctas = 0;
  size_t stats_bytes = 0;
  int n_all = 0;  // total time_emb_proj outputs
  long long p_freqs = -1, p_w1 = -1, p_b1 = -1, p_w2 = -1, p_b2 = -1, p_wall = -1, p_ball = -1;

  // bound state
  int device = -1;           // CUDA device of the workspace (dv_plan_bind); every entry point runs there
  uint8_t* ws = nullptr;
  const uint8_t* params = nullptr;
  bool bound = false, prompt_ready = false, has_aug = false, has_kvlen = false;
  int last_sample_launches = 0;
  int last_diag_rows = 0;    // rows of sampler diagnostics the most recent dv_sample wrote (0: switched off)

  // graph cache for dv_sample
  cudaGraphExec_t graph_exec = nullptr;
  cudaStream_t cap_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // lane fork / join inside a captured group loop
  // time tables one call ahead: the time-embedding MLP + the 22 time_emb_proj of call k+1 run on a side stream while
  // call k computes (they depend on t only), into the other of two table slots
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_tfork = nullptr, ev_tjoin = nullptr;
  uint64_t graph_key = 0;
  uint64_t pending_key = 0;   // loop seen once (run eagerly): it is captured when it comes back
  uint64_t graph_epoch = 0;   // bumped whenever launches baked into a captured graph become stale
  int graph_launches = 0;
  int graph_diag_rows = 0;

  int build(std::string& err);
  int bind(void* workspace, const void* params, cudaStream_t st, std::string& err);
  int enqueue_ops(std::vector<Op>& list, float* dyn_out, cudaStream_t st, std::string& err, int* launches,
                  cudaEvent_t* events = nullptr, int ss_slot = 0);
  // time-embedding MLP + all time_emb_proj for one call into table slot `slot` (2 launches)
  int enqueue_time_tables(const float* t, int t_stride, int slot, cudaStream_t st, std::string& err);
  // tables_slot < 0: compute the time tables inline (slot 0) before the ops; >= 0: they are already in that slot
  int forward(const float* sample, int x_only, const float* t, int t_stride, float* out, cudaStream_t st,
              std::string& err, int* launches, cudaEvent_t* events = nullptr, int tables_slot = -1);
  int prompt_prepare(const float* ehs, const float* aug, const int32_t* kv_len, cudaStream_t st, std::string& err);
  int set_cond(const float* cond, cudaStream_t st, std::string& err);
  // Debug/measurement: one eager forward with a CUDA event between consecutive launches; synchronises.
  int profile_forward(const float* sample, int x_only, const float* t, int t_stride, float* out, cudaStream_t st,
                      float* ms, int cap, std::string& err);
  int sample(int sampler, int nsteps, const float* t_in, const float* coef, float* x, int flags, cudaStream_t st,
             std::string& err);
  // the N-step loop (denoiser calls + sampler updates) of this plan, enqueued on `s` (no copies, no graph handling)
  // diag_rows != nullptr: sampler diagnostics on, *diag_rows = rows written (one per update launch, <= DIAG_MAX_ROWS)
  int enqueue_loop(int sampler, int nsteps, const float* coef, bool roundtrip, cudaStream_t s, int* launches,
                   std::string& err, int* diag_rows = nullptr);
  ~Plan();
};

int sample_group(Plan* const* plans, int n, int sampler, int nsteps, const float* t_in, const float* coef,
                 float* const* x, int flags, cudaStream_t st, std::string& err);

std::string manifest_json(const std::vector<ParamSlot>& slots);

// PromptEncoder (model3.py:382-433): the speech-prompt encoder that produces encoder_hidden_states, as a plan of the same
// kernels (csrc/prompt_encoder.cu). Channel-last [B, Tp, 128] throughout = the layout the UNet's cross-attention reads.
constexpr int DIAG_MAX_ROWS = 512;   // update launches of one loop that get a diagnostics row (50-step DPM++: 50, 25-step UniPC: 49)

struct PromptPlan {
  dv_prompt_cfg cfg;
  int B = 0, Tp = 0, device = -1;
  std::vector<ParamSlot> slots;
  std::string manifest;
  size_t param_bytes = 0, ws_bytes = 0;
  uint8_t* ws = nullptr;
  const uint8_t* params = nullptr;
  bool bound = false;
  struct Step {
    int kind = 0;                   // 0 GEMM, 1 LayerNorm (masked variant), 2 attention
    GemmParams gp; AttnParams at;
    size_t x = 0, out = 0; long long gamma = -1, beta = -1; LnMaskParams ln;
    bool ln_f32 = false;
    // GEMM description resolved at bind
    size_t a_off = 0; int a_c = 0; int nseg = 0; int row_off[9] = {0}; long long w_off = -1, b_off = -1; int k_tot = 0, n = 0;
    size_t res_off = 0; bool has_res = false, relu = false, mask_rows = false;
    size_t q = 0, k = 0, v = 0, o = 0;
    std::string tag;
  };
  std::vector<Step> steps;
  size_t off_xin = 0, off_len = 0, off_out32 = 0;
  int build(std::string& err);
  int bind(void* workspace, const void* params, cudaStream_t st, std::string& err);
  int encode(const float* prompt, const int32_t* lengths, float* ehs_out, cudaStream_t st, std::string& err);
};

// ---- multi-GPU: the single collective of the path (csrc/comm.cu), NCCL resolved at run time
struct Comm;
int comm_unique_id(const char* nccl_path, void* id128, std::string& err);
int comm_init(const char* nccl_path, const void* id128, int rank, int nranks, Comm** out, std::string& err);
int comm_all_gather(Comm* c, const float* send, float* recv, size_t count, cudaStream_t st, std::string& err);
void comm_destroy(Comm* c);

bool make_tmap_3d(CUtensorMap* m, const void* base, int view_c, int ld, int L, int B, size_t batch_stride_elems,
                  std::string& err, int box_rows = 128);
bool make_tmap_2d(CUtensorMap* m, const void* base, int k_tot, int rows, int box_rows, std::string& err);

}  // namespace dv
