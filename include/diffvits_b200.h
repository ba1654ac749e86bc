/*
 * diffvits_b200 - C ABI of the B200-native Diff-VITS denoising loop (libdiffvits_b200.so).
 *
 * The reference (adelacvg/diff-vits) is pure PyTorch and has no FFI layer; its boundary for this path is the
 * Python class API. Each entry point below names the reference interface it replaces (file:line relative to the
 * reference tree). The Python shims in diff-vits_b200/dropin/{unet1d,sampler} bind these with ctypes and keep
 * the reference's class names, constructor kwargs, state_dict layout and call signatures.
 *
 * Conventions
 *  - return 0 on success, negative dv_status on failure; dv_last_error() gives a thread-local message.
 *  - the caller owns every device buffer (workspace, parameter blob, inputs, outputs); the library never
 *    allocates device memory, never synchronises the device and never calls back. All work is enqueued on the
 *    given CUDA stream (cudaStream_t passed as void*) and is CUDA-graph-capture safe.
 *  - one plan per (device, B, T, Tp); a plan is not thread safe, distinct plans are independent.
 *  - there is no CPU path: every compute entry point requires a CUDA device of compute capability 10.x.
 *  - tensors: "NCT" = [B, C, T] fp32 contiguous (the reference's layout), "NTC" = channel-last.
 */
#ifndef DIFFVITS_B200_H
#define DIFFVITS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DV_VERSION 103

#if defined(__GNUC__)
#define DV_API __attribute__((visibility("default")))
#else
#define DV_API
#endif

typedef enum dv_status {
  DV_OK = 0,
  DV_ERR_INVALID = -1,     /* bad argument / unsupported configuration (reference: ValueError, unet_1d_condition.py:208-255) */
  DV_ERR_CUDA = -2,        /* CUDA runtime / driver error, message has the cudaError string */
  DV_ERR_STATE = -3,       /* call order violated (e.g. forward before bind / prompt_prepare) */
  DV_ERR_UNSUPPORTED = -4  /* valid for the reference but outside the accelerated path */
} dv_status;

/* UNet1DConditionModel constructor arguments that shape the network
 * (reference: unet1d/unet_1d_condition.py:151-203; call site model3.py:887-896). */
typedef struct dv_unet_cfg {
  int in_channels;            /* 228 = 100 latent + 128 cond on the diffusion path */
  int out_channels;           /* 100 */
  int n_levels;               /* len(block_out_channels), 2..8 */
  int block_out_channels[8];  /* (128, 256, 384, 512) */
  int layers_per_block;       /* 2 */
  int norm_num_groups;        /* 8 */
  int cross_attention_dim;    /* 128 */
  int num_heads;              /* `attention_head_dim` in the reference ctor (it is the head COUNT): 8 */
  float norm_eps;             /* 1e-5 */
  int x_channels;             /* leading input channels that change every step (the noisy latent): 100 */
  int precision;              /* DV_PRECISION_BF16 (0), _FP32X3 (1), _FP16S (2) or _FP16 (3), see below */
} dv_unet_cfg;

/* Arithmetic of the contractions. The reference computes in FP32 (unet1d/lora.py:99-104, attention_processor.py:1032).
 *  DV_PRECISION_BF16   : BF16 operands, FP32 accumulation - the fast path (per-call rel-L2 vs the FP32 reference ~8e-3,
 *                        the same floor as the reference under torch.autocast(bfloat16)).
 *  DV_PRECISION_FP32X3 : every BF16 operand carries a second "lo" plane (x = hi + lo, 16 mantissa bits) and every product
 *                        is a_hi.w_hi + a_hi.w_lo + a_lo.w_hi on the same tcgen05 kernel (three K segments per logical
 *                        segment); norms / SiLU / GELU / attention core in FP32. Meets the FP32 gate (1e-4 rel-L2). */
#define DV_PRECISION_BF16 0
#define DV_PRECISION_FP32X3 1
/*  DV_PRECISION_FP16S  : FP16 operands (same tcgen05 kind::f16 rate and bytes as BF16, 3 more mantissa bits, saturating
 *                        conversions) with the residual stream and every norm input carried as hi + lo planes (the lo plane
 *                        is read by epilogues and norms only, never by the tensor core: ONE MMA per product). Measured by
 *                        emulation at 8.0e-4 rel-L2 per call vs the FP32 reference (plain FP16 operands: 1.0e-3): the mode
 *                        that meets north_star's 1e-3 gate of the 16-bit tier. */
#define DV_PRECISION_FP16S 2
/*  DV_PRECISION_FP16   : FP16 operands everywhere the BF16 mode has BF16 ones (single planes, the same launches, chains and
 *                        fused norms: the BF16 mode's speed) - 3 more mantissa bits per rounding: ~1e-3 rel-L2 per call instead
 *                        of 8e-3. Conversions saturate at +-65504; the narrower exponent range is the price (the sampler
 *                        diagnostics report non-finite values). */
#define DV_PRECISION_FP16 3

typedef struct dv_plan dv_plan;

DV_API int dv_version(void);
DV_API const char* dv_last_error(void);

/* Build the per-call kernel program for UNet1DConditionModel.forward at a fixed (B, T, Tp).
 * Replaces: nn.Module construction + eager dispatch of unet_1d_condition.py:743-1037. Requires T % 2^(n_levels-1) == 0. */
DV_API int dv_plan_create(const dv_unet_cfg* cfg, int B, int T, int Tp, dv_plan** out);
DV_API void dv_plan_destroy(dv_plan* plan);

DV_API size_t dv_plan_workspace_bytes(const dv_plan* plan);
DV_API size_t dv_plan_param_bytes(const dv_plan* plan);
/* JSON list describing every slot of the packed parameter blob: which state_dict tensors
 * (reference key names, e.g. "down_blocks.0.resnets.0.conv1.weight") go where and in which packed layout. */
DV_API const char* dv_plan_manifest_json(const dv_plan* plan);
/* Number of kernels one dv_unet_forward enqueues (used for the bench's gpu_launches claim). */
DV_API int dv_plan_launches_per_forward(const dv_plan* plan);

/* Attach caller-owned device memory (256-byte aligned): workspace (zeroed by this call) and the packed parameters. */
DV_API int dv_plan_bind(dv_plan* plan, void* workspace, const void* params, void* stream);

/* Step-invariant prompt work, hoisted out of the loop: bf16 copy of encoder_hidden_states, the 16 cross-attention
 * K/V projections (attention_processor.py:1014-1015), the key-padding lengths and the text-time embedding `aug`
 * (unet_1d_condition.py:869-870; computed by the caller).   ehs: [B, Tp, cross_attention_dim] fp32,
 * aug: [B, 4*C0] fp32 or NULL, kv_len: [B] int32 or NULL (= no mask). */
DV_API int dv_prompt_prepare(dv_plan* plan, const float* ehs, const float* aug, const int32_t* kv_len, void* stream);

/* Step-invariant conditioning channels of the UNet input (model3.py:908 torch.cat([x, cond])): cond NCT [B, in-x, T]. */
DV_API int dv_set_cond(dv_plan* plan, const float* cond, void* stream);

/* UNet1DConditionModel.forward (unet_1d_condition.py:743-1037) for one denoiser call.
 * sample: NCT [B, in_channels, T] (full concat) if x_only == 0, else NCT [B, x_channels, T] with dv_set_cond done.
 * t: device fp32 model times, t[b * t_stride] (t_stride 0 broadcasts one value).  out: NCT [B, out_channels, T]. */
DV_API int dv_unet_forward(dv_plan* plan, const float* sample, int x_only, const float* t, int t_stride, float* out,
                    void* stream);

/* Whole multistep sampling loop (DPM_Solver.sample dpm_solver.py:1171-1213 / UniPC.sample uni_pc.py:606-658) with
 * the denoiser, captured in one CUDA graph and replayed. Coefficients are computed on the host by the caller with
 * the reference's own NoiseScheduleVP formulas.  t_in: host [nsteps] model times. coef: host [nsteps * 8]:
 *   DV_SAMPLER_DPMPP_2M  row i (update to t_{i+1}): {A, B, C, has_prev, alpha_prev, sigma_prev, 0, 0}
 *   DV_SAMPLER_UNIPC_BH2 row i: {a, b, cp, cc, d, has_prev, use_corrector, 0}
 * x: device NCT [B, x_channels, T] fp32, noise in, latent out. flags: DV_SAMPLE_*
 * A loop (plan, sampler, tables) is enqueued directly the first time it is seen and captured when the same loop comes
 * back (capture + instantiate cost ~45 ms, a third of a one-shot synthesis); DV_GRAPH_FIRST=1 in the environment
 * captures at first sight. Both paths issue the same kernels and give identical bits. */
#define DV_SAMPLER_DPMPP_2M 0
#define DV_SAMPLER_UNIPC_BH2 1
#define DV_SAMPLE_NO_GRAPH 1     /* enqueue directly instead of capturing/replaying a graph */
#define DV_SAMPLE_ROUNDTRIP 2    /* DPM++ only: reproduce the reference's x0->eps->x0 round trip bit for bit */
#define DV_SAMPLE_NO_DIAG 4      /* do not record the sampler diagnostics (dv_plan_sample_diag) */
DV_API int dv_sample(dv_plan* plan, int sampler, int nsteps, const float* t_in, const float* coef, float* x, int flags,
              void* stream);
/* The same loop for `nplans` plans at once ("lanes": disjoint slices of one batch of independent utterances - the path
 * has no cross-sample reduction - with the same T, Tp, sampler and tables). All lanes are captured into ONE graph,
 * lane i on its own stream forked from lane 0's, so launch latencies and kernel tails of one lane overlap with the
 * other lanes' kernels. x[i]: device NCT [B_i, x_channels, T] of lane i. The graph is cached on plans[0]. */
DV_API int dv_sample_group(dv_plan* const* plans, int nplans, int sampler, int nsteps, const float* t_in,
                           const float* coef, float* const* x, int flags, void* stream);
/* Kernel launches enqueued by the most recent dv_sample / dv_sample_group (reported on plans[0]). */
DV_API int dv_plan_last_sample_launches(const dv_plan* plan);
/* Sampler diagnostics of the most recent dv_sample on this plan. Replaces the reference's per-call, device-synchronising
 * `assert torch.isnan(x).any() == False` (model3.py:903): every sampler-update kernel also reduces (warp shuffles -> CTA ->
 * one fixed-order fold after the loop; no atomics, no synchronisation) the denoiser output it consumes and the state it
 * writes. Row j (fp32 x 4, at byte `offset + 16 j` of the workspace, j < rows <= 512) belongs to the j-th update launch of the
 * loop: {count of non-finite values, sum of squares of the denoiser output, sum of squares of the new state, kernel
 * (0 DPM-Solver++ update, 1 UniPC predictor, 2 UniPC corrector)}. Valid once the stream has passed the loop; the caller reads
 * them with the latent (the drop-in raises the reference's AssertionError from them). */
DV_API int dv_plan_sample_diag(const dv_plan* plan, size_t* offset, int* rows);
/* One uint32 at byte `offset` of the workspace, zero while the plan is healthy. The GroupNorm that the producing GEMM applies
 * in its epilogue waits (bounded) for the other tiles of a batch item; if such a wait ever gives up (the CTAs of one launch
 * were not co-scheduled: not observed, but the hardware gives no guarantee when several plans run concurrently on one
 * device) the kernel sets this flag instead of hanging the GPU and the plan's results must be discarded (rebind to clear). */
DV_API int dv_plan_health_flag(const dv_plan* plan, size_t* offset);

/* ---- PromptEncoder (model3.py:382-433; model.py:72-97, 137-171; operations.py:644-714, 784-821) ---------------------------
 * The speech-prompt encoder whose output is the UNet's encoder_hidden_states: ConvLayer pre (LayerNorm + 1x1), n_layers x
 * {LayerNorm, 8-head self-attention with key padding, LayerNorm, 9-tap conv feed-forward (ReLU, 9^-1/2)}, ConvLayer out,
 * final LayerNorm, padding rows zeroed where the reference multiplies by (1 - mask). The reference evaluates it once per
 * denoiser call (model3.py:906); it depends on the prompt only, so the drop-in evaluates it once per job with this plan. */
typedef struct dv_prompt_cfg {
  int in_channels;   /* 100 mel bins */
  int hidden;        /* 128 (= out channels = cross_attention_dim); only 128 is built */
  int n_layers;      /* 4 */
  int num_heads;     /* 8 (head dim 16) */
  int ffn_kernel;    /* 9 */
  int ffn_mult;      /* 4: filter size = 4 * hidden */
} dv_prompt_cfg;
typedef struct dv_prompt_plan dv_prompt_plan;
DV_API int dv_prompt_plan_create(const dv_prompt_cfg* cfg, int B, int Tp, dv_prompt_plan** out);
DV_API void dv_prompt_plan_destroy(dv_prompt_plan* plan);
DV_API size_t dv_prompt_plan_workspace_bytes(const dv_prompt_plan* plan);
DV_API size_t dv_prompt_plan_param_bytes(const dv_prompt_plan* plan);
DV_API const char* dv_prompt_plan_manifest_json(const dv_prompt_plan* plan);
DV_API int dv_prompt_plan_bind(dv_prompt_plan* plan, void* workspace, const void* params, void* stream);
/* PromptEncoder.forward(prompt, lengths) * mask, transposed: prompt NCT [B, in_channels, Tp] fp32, lengths [B] int32
 * (device), ehs_out [B, Tp, hidden] fp32 = the encoder_hidden_states Diffusion_Encoder.forward passes to the UNet
 * (model3.py:905-912), rows >= lengths[b] zero. */
DV_API int dv_prompt_encode(dv_prompt_plan* plan, const float* prompt, const int32_t* lengths, float* ehs_out, void* stream);

/* ---- multi-GPU -----------------------------------------------------------------------------------------
 * Independent utterances are batch-sharded, one process per GPU; the only collective of the path is ONE all-gather of the
 * output latents per job (the reference has no cross-sample reduction). It is issued in-stream, right behind dv_sample, on a
 * communicator this library owns. NCCL is resolved at run time (`nccl_path`, else $DV_NCCL_LIB, else libnccl.so.2 - pass the
 * path of the libnccl the host framework already loaded so that both use one instance); nothing is linked.
 *   rank 0: dv_comm_unique_id(path, id)  ->  the 128 bytes travel to every rank by any means (the drop-in broadcasts them
 *   through torch.distributed)  ->  every rank: dv_comm_init(path, id, rank, nranks, &comm) (collective call)
 *   dv_comm_all_gather: recv[r * count .. (r + 1) * count) = rank r's `send` (count fp32 elements per rank, equal on all). */
typedef struct dv_comm dv_comm;
DV_API int dv_comm_unique_id(const char* nccl_path, void* id128);
DV_API int dv_comm_init(const char* nccl_path, const void* id128, int rank, int nranks, dv_comm** out);
DV_API int dv_comm_all_gather(dv_comm* comm, const float* send, float* recv, size_t count, void* stream);
DV_API void dv_comm_destroy(dv_comm* comm);

/* ---- measurement support ------------------------------------------------------------------------------------- */
/* Block outputs (bf16 [B, L, C] at byte `offset` of the workspace, named by reference module path) that are still
 * intact after dv_unet_forward returns: lets tests compare the network block by block with the oracle. */
DV_API int dv_plan_tensor_count(const dv_plan* plan);
DV_API int dv_plan_tensor_info(const dv_plan* plan, int i, char* name, int name_cap, size_t* offset, int* L, int* C);
/* The per-call program: op i has a tag (reference module path + role), a kind (0 GEMM/conv on tcgen05, 1 GroupNorm
 * statistics, 2 GroupNorm apply, 3 LayerNorm, 4 attention) and its algorithmic FLOPs / HBM bytes per launch. */
DV_API int dv_plan_op_count(const dv_plan* plan);
DV_API int dv_plan_op_info(const dv_plan* plan, int i, char* tag, int tag_cap, int* kind, double* flops, double* bytes);
/* One dv_unet_forward captured into a CUDA graph with an event-record node between consecutive launches and
 * replayed twice; the second replay is reported. SYNCHRONISES the stream (the only entry point that does).
 * ms[i] = device time of op i (i < op_count) incl. the gap to the next launch, ms[op_count] = prologue (layout
 * change + time MLP). */
DV_API int dv_plan_profile_forward(dv_plan* plan, const float* sample, int x_only, const float* t, int t_stride,
                                   float* out, float* ms, int cap, void* stream);

/* ---- single-kernel entry points (used by the parity tests and micro-benchmarks) ------------------------------- */
/* bf16 rows of `ld` elements, `L` rows visible per batch item (rows outside [0, L) read as zero), `C` channels used,
 * `bs` elements between batch items (0 = L * ld) */
typedef struct dv_gemm_src { const void* ptr; int L; int C; int ld; long long bs; } dv_gemm_src;
/* One K segment: `npanels` 64-channel panels of source `src` starting at channel c0. ntap = 1 (or 0): rows l + row_off.
 * ntap = 3: conv k3 with shared rows - row_off is the first tap's offset (-1) and tap t reads rows l + row_off + t from the
 * same staged box. wk0 = first 64-wide K panel of (tap 0, panel 0) in `w` (< 0: the running K position), tap_stride = K panels
 * between taps (0: npanels). xf = 1: the segment's rows are GroupNorm-normalised in shared memory on the way to the tensor
 * core (dv_gemm_opts.norm), xf_c0 = channel of c0 inside the normalised tensor. */
typedef struct dv_gemm_seg { int src; int row_off; int c0; int npanels; int ntap; int wk0; int tap_stride; int xf; int xf_c0; } dv_gemm_seg;
/* GroupNorm fused into the consuming GEMM (nn.GroupNorm + SiLU + (1+scale)*x+shift, resnet.py:591-631): statistics as
 * partial sums [B][chunks][nblk][2] per source (what dv_groupnorm's scratch / the GEMM epilogue hold). */
typedef struct dv_gemm_norm {
  const float* stats[2]; int chunks[2]; int nblk[2]; int fold[2]; int cover[2]; int nst;
  int C; int L; int G; float eps; int silu;
  const float* gamma; const float* beta; const float* scale_shift; int ss_ld;
} dv_gemm_norm;
/* Optional per-call settings of dv_gemm (NULL = defaults): bn = force the N tile (64 / 128 / 192 / 256, 0 = heuristic);
 * stats != NULL = the epilogue also emits GroupNorm partial sums [B][ceil(L/128)*4][n/16][2] of the stored tensor (what
 * the engine's GroupNorm consumes instead of re-reading it); out_bs = elements between batch items of `out` (0 = dense
 * L_out * ld_out); trim_col > 0 = columns >= trim_col of the last row are neither stored nor counted in the partials.
 * out_bs + trim_col are how the folded nearest-resize conv writes 2L - 1 rows (resnet.py:138-173). */
typedef struct dv_gemm_opts { int bn; float* stats; long long out_bs; int trim_col; const dv_gemm_norm* norm; } dv_gemm_opts;
/* out_mode: 0 bf16 NTC, 1 fp32 NCT, 2 GEGLU bf16. impl: 0 = tcgen05 kernel, 1 = CUDA-core reference (writes raw
 * fp32 [B, L, n_pad] accumulators + bias into `out`). Replaces F.conv1d (unet1d/lora.py:102) / nn.Linear. */
DV_API int dv_gemm(const dv_gemm_src* srcs, int nsrc, const dv_gemm_seg* segs, int nseg, const void* w, int k_tot, int n,
            int n_pad, const float* bias, const void* res, int ld_res, void* out, int ld_out, int out_mode, int B,
            int L_out, int impl, const dv_gemm_opts* opts, void* stream);
/* F.scaled_dot_product_attention (attention_processor.py:1032); bf16 q/k/v/o with element strides. */
DV_API int dv_attention(const void* q, const void* k, const void* v, void* o, int ld_q, int ld_k, int ld_v, int ld_o,
                 long long bs_q, long long bs_k, long long bs_v, long long bs_o, int B, int H, int Lq, int Lk, int D,
                 float scale, const int32_t* kv_len, void* stream);
/* nn.GroupNorm (+SiLU, + (1+scale)*x+shift) on channel-last bf16 (resnet.py:591-631).
 * scratch: dv_groupnorm_scratch_floats(B, L, c0 + c1) floats: partial sums per (32-row chunk, 16-channel block),
 * the layout the GEMM epilogue emits inside the engine; deterministic fixed-order reduction. */
DV_API size_t dv_groupnorm_scratch_floats(int B, int L, int C);
DV_API int dv_groupnorm(const void* x0, int c0, const void* x1, int c1, int B, int L, int G, float* scratch,
                 const float* gamma, const float* beta, const float* scale_shift, int ss_ld, float eps, int silu,
                 void* out, void* stream);
/* nn.LayerNorm over the last dim of bf16 [rows, C] (attention.py:83,102,118). */
DV_API int dv_layernorm(const void* x, void* out, const float* gamma, const float* beta, int rows, int C, float eps,
                 void* stream);
/* DPM-Solver++ multistep update x_out = a*x + b*m1 + c*m0 (dpm_solver.py:547-580, 796-852). m0 may be NULL. */
DV_API int dv_dpmpp_update(float* x_out, const float* x, float* m1, const float* m0, float a, float b, float c,
                    int roundtrip, float alpha_prev, float sigma_prev, size_t n, void* stream);
/* UniPC bh predictor / corrector (uni_pc.py:471-588). */
DV_API int dv_unipc_predict(float* xpred, float* base, const float* x, const float* m1, const float* m0, float a, float b,
                     float cp, float cc, int has_prev, size_t n, void* stream);
DV_API int dv_unipc_correct(float* x, const float* base, const float* mt, const float* m1, float d, size_t n, void* stream);
/* [B, C, T] fp32 -> channel-last bf16 [B, T, ld] at column col0. */
DV_API int dv_nct_to_ntc(const float* src, void* dst, int B, int C, int T, int ld, int col0, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFVITS_B200_H */
