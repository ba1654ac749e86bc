// Launch-level interface of the sm_100a kernels (host side). All launches are stream-ordered, never
// synchronise, never allocate: safe inside CUDA-graph capture.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#include <utility>

namespace dv {

// Timing-ablation switches and CTA timelines inside the kernels exist only in the debug build (-DDV_DEBUG_HOOKS,
// libdiffvits_b200_dbg.so); the product library compiles them out.
#ifdef DV_DEBUG_HOOKS
#define DV_DBG(p, mask) ((p).dbg_flags & (mask))
#else
#define DV_DBG(p, mask) 0
#endif

// Launch with the programmatic-stream-serialization attribute (PDL): consecutive kernels of the denoiser overlap
// their prologue / launch latency with the previous kernel's tail. Disabled with DV_NO_PDL=1 in the environment.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ------------------------------------------------------------------ tap-GEMM (conv1d k3 / 1x1 / linear)
// out[b, l, n] = epi( sum over segments, panels p, taps t   A_seg[b, l + row_off + t, c0 + 64*p : +64] . W[n, panel(t, p)] )
// A sources are channel-last bf16 tensors [B, L, C] behind 3-D TMA maps (zero fill outside [0, L) gives the
// conv padding); W is bf16 [N_pad, K_tot] K-major behind a 2-D TMA map.
// Precision modes. BF16 (default): activations and weights are single bf16 tensors. SPLIT ("fp32x3"): every bf16 tensor has
// a second "lo" plane holding bf16(x - float(hi)) (hi + lo carries 16 mantissa bits) and every contraction is
// a_hi.w_hi + a_hi.w_lo + a_lo.w_hi in the fp32 accumulator - for the tap-GEMM simply three K segments per logical
// segment, over two A maps (hi / lo plane) and two W maps. Epilogues write both planes; norms, layout kernels and the
// attention core read hi + lo and write both planes.
constexpr int GEMM_MAX_LSRC = 3;                   // logical sources (torch.cat inputs / pair views)
constexpr int GEMM_MAX_LSEG = 9;                   // logical K segments (conv taps, shortcut sources; 9-tap conv-FFN of the prompt encoder)
constexpr int GEMM_MAX_SRC = 2 * GEMM_MAX_LSRC;    // maps in the kernel parameters (hi and lo plane of each source)
constexpr int GEMM_MAX_SEG = 3 * GEMM_MAX_LSEG;

enum GemmOut : int {
  GEMM_OUT_BF16 = 0,     // out bf16 [B, L, ld_out], columns [0, N)
  GEMM_OUT_F32_NCL = 1,  // out fp32 [B, N, L]  (transposed store, used by conv_out)
  GEMM_OUT_GEGLU = 2,    // tile columns [0,64) = u, [64,128) = gate: out bf16 [B, L, ld_out], N/2 columns
};

// One K segment = `npanels` 64-channel panels of one source. Every panel is ONE staged activation box ("A slot": rows
// l0 + row_off ... of the source, TMA zero fill outside [0, L) = the conv padding) that feeds `ntap` MMAs:
//   ntap = 1 : D += A[rows 0..127] . W[panel]                                   (1x1 conv / Linear / one conv tap)
//   ntap = 3 : conv k3 with SHARED rows: the box holds 136 rows starting at l0 - 1 and tap t reads rows t .. t+127 through
//              a row-shifted UMMA descriptor (+128 B per row; profiles/r2b_umma_row_shift_probe.txt) against
//              W[wk0 + t * tap_stride + panel]: the activation tile is loaded (and transformed) once instead of three times.
// xf != 0: the box is normalised IN SHARED MEMORY before the MMA reads it (GemmXf below): GroupNorm (+scale/shift, +SiLU)
// of the resnet blocks / transformer input fused into the consuming conv instead of a stand-alone pass over HBM.
struct GemmSeg {
  int16_t src;         // index into mapA
  int16_t row_off;     // first row of the A slot relative to the tile's first output row (-1 for ntap = 3)
  int16_t c0;          // first channel of the segment in the source tensor
  int16_t npanels;     // number of 64-channel panels
  int16_t wsel;        // weight matrix: 0 = mapW (hi plane), 1 = mapW2 (lo plane, split precision)
  int16_t wk0;         // 64-wide K panel of (tap 0, panel 0) inside that weight matrix
  int16_t ntap;        // 1 or 3
  int16_t tap_stride;  // K panels between consecutive taps of the same channel panel
  int16_t xf;          // 0 = operand as loaded, 1 = GroupNorm affine (+SiLU) applied in shared memory
  int16_t xf_c0;       // channel of c0 inside the normalised (possibly concatenated) tensor: coefficient index
};

// GroupNorm fused into the consumer (xf segments): y = SiLU?( ((x - mean_g) rstd_g gamma_c + beta_c) (1 + scale[b,c]) + shift[b,c] ),
// statistics from the producers' epilogue partials (same layout / reduction order as gn_apply, csrc/norm.cu).
struct GemmXf {
  const float* st_ptr[2];            // partial statistics of the (up to two) sources of the normalised tensor
  int st_chunks[2], st_nblk[2], st_fold[2], st_cover[2];
  int nst;
  int C, L, G;                       // channels of the normalised tensor, rows per batch item, groups
  float eps;
  int silu;
  const float* gamma; const float* beta;      // [C]
  const float* scale_shift; int ss_ld;        // optional per-(b, c): scale at [b*ss_ld + c], shift at [b*ss_ld + C + c]
};

// GroupNorm of this GEMM's OUTPUT applied by the producing CTAs themselves (out != nullptr), instead of a stand-alone
// gn_apply launch: after a tile's raw store + partial statistics, the CTA arrives on its batch item's counter, waits until
// every tile of that item (tiles_per_batch x N tiles, a contiguous range of the static tile schedule) has arrived, folds the
// item's partials into mean / rstd (same order as gn_apply: bit-identical) and writes the normalised (+scale/shift, +SiLU)
// copy of its own tile. Only the epilogue warps wait; the producer / MMA warps go on with the CTA's next tile.
struct GemmGnOut {
  void* out;                                  // normalised 16-bit [B, L, ld_out] (same geometry as the raw output)
  void* out_lo;                               // split precision: lo plane of the normalised tensor (FP32X3), or nullptr
  unsigned* counter;                          // [B], zero between launches (arrivals then departures; the last departure resets it);
                                              // counter[B] is a sticky flag: set when a wait gave up (results of that plan are then invalid)
  int G; float eps; int silu;
  const float* gamma; const float* beta;      // [N]
  const float* scale_shift; int ss_ld;        // optional per-(b, c): scale at [b*ss_ld + c], shift at [b*ss_ld + N + c]
  int no_stash;                               // A/B switch (DV_GN_STASH=0): normalisation pass re-reads the tile from global memory
};

struct alignas(64) GemmParams {
  CUtensorMap mapA[GEMM_MAX_SRC];
  CUtensorMap mapW;
  CUtensorMap mapW2;    // split precision: lo plane of the weights (aliases mapW otherwise)
  GemmSeg seg[GEMM_MAX_SEG];
  int nseg;
  int16_t map_rows[GEMM_MAX_SRC];   // rows of the TMA box of each A map (128, or 136 for sources read with shared taps)
  GemmXf xf;                        // used when any segment has xf != 0
  GemmGnOut gn;                     // gn.out != nullptr: GroupNorm of the output applied in the epilogue (GEMM_OUT_BF16, dense, with stats)
  int B, L, N;          // L = output rows per batch item, N = real output columns
  int n_pad;            // N rounded up to the N tile
  int tiles_per_batch;  // ceil(L / 128)
  int out_mode;
  const float* bias;    // [n_pad] or nullptr
  const void* res;      // bf16 [B, L, ld_res] residual added before the store, or nullptr
  const void* res_lo;   // split precision: lo plane of the residual
  int ld_res;
  void* out;
  void* out_lo;         // split precision: lo plane of the output (non-null selects the split epilogue)
  int ld_out;
  long long out_bs;     // GEMM_OUT_BF16: elements between batch items of `out` (0 = L * ld_out, dense)
  int trim_col;         // GEMM_OUT_BF16: if > 0, columns >= trim_col of the LAST row (l = L - 1) are neither stored nor counted in
                        // the GroupNorm partials (odd-length nearest-2x upsample: that half row does not exist)
  int relu;             // GEMM_OUT_BF16: max(v, 0) before the store (conv-FFN of the prompt encoder, operations.py:689)
  const int* row_len;   // GEMM_OUT_BF16: optional [B] valid rows per batch item; rows >= row_len[b] are stored as zeros
                        // (the reference's `x * (1 - padding_mask)`, operations.py:813, 820; model3.py:421-431)
  float* stats;         // optional GroupNorm partials of the stored tensor: [B][tiles_per_batch*4][stats_nblk][2]
  int stats_nblk;       // 16-channel blocks per row chunk (= N / 16)
  int bn;               // N tile (128 / 256); 0 = choose
  int dbg_flags;        // debug: 1 = skip the epilogue body (release TMEM only), 2 = skip MMA issue, 4 = skip TMA loads, 32 = no residual preload, 64 = no W prefetch before griddepcontrol.wait
  unsigned long long* trace;  // debug: CTA 0 appends (role, tile, event, clock) records; [0] = record count
};

// Debug/reference implementation of exactly the same contract on CUDA cores (plain loads, fp32 FMA).
// Used by tests to cross-check the tcgen05 kernel; never used by the engine.
struct GemmRefSrc { const void* ptr; int L, C, ld; };

// ------------------------------------------------------------------ chained GEMMs (csrc/chain_tc.cu)
// One persistent kernel runs a short program of row-local GEMMs of a transformer block over 128-row tiles, keeping
// the activations between them in shared memory / TMEM instead of HBM:
//   CH_ROW    : D[128, C] = A . W^T (+bias) (+res)  -> bf16 global (optional) -> LayerNorm (or plain copy) -> X (smem)
//               (or -> global `out2`: the last stage of a chain, e.g. proj_in + LayerNorm as ONE launch)
//               A is streamed from global by TMA (a_map >= 0) or is the X buffer (a_map < 0).
//   CH_STREAM : out[128, n] = X . W^T (+bias) (+res) -> bf16 global (+GroupNorm partials), 128 columns at a time
//   CH_FF     : GEGLU feed-forward: for each 64 hidden units  h = u * gelu(g), [u|g] = X . W1_chunk^T + b1 ; the bf16
//               h chunk goes to shared memory and is immediately contracted with W2[:, chunk] into the [128, C]
//               accumulator; result (+bias2) (+res) -> X (smem). The [L, 8C] / [L, 4C] intermediates never exist.
// C = 128 * U, U in {1, 2, 3}. Weights are the same packed [n_pad, K] K-major matrices the single GEMMs use.
constexpr int CHAIN_MAX_STAGES = 4;
constexpr int CHAIN_MAX_W = 5;
enum ChainKind : int { CH_ROW = 0, CH_STREAM = 1, CH_FF = 2 };

struct ChainStage {
  int kind;
  int a_map;              // CH_ROW: index into mapA of the streamed A operand, or -1 = X buffer
  int w_map, w2_map;      // indices into mapW (w2: CH_FF second matrix)
  int n;                  // CH_STREAM: output columns (multiple of 128); CH_FF: hidden width (multiple of 64)
  const float* bias;      // [n] (CH_ROW: [C]; CH_FF: [2n] in the interleaved u/g packing), may be null
  const float* bias2;     // CH_FF: [C]
  const void* res;        // bf16 [B, L, ld_res] added before the store, or null
  int ld_res;
  void* out;              // bf16 [B, L, ld_out] or null
  int ld_out;
  const float* gamma;     // CH_ROW: LayerNorm affine (null = copy the bf16 result to X unchanged)
  const float* beta;
  float eps;
  void* out2;             // CH_ROW: if set, the (normalised) row goes to global bf16 [B, L, ld_out2] instead of X:
  int ld_out2;            //         a GEMM with the LayerNorm of its own output fused in, as a one-stage chain
  float* stats;           // CH_STREAM: GroupNorm partials of the stored tensor (layout as GemmParams::stats) or null
  int stats_nblk;
};

struct alignas(64) ChainParams {
  CUtensorMap mapA[2];
  CUtensorMap mapW[CHAIN_MAX_W];
  ChainStage st[CHAIN_MAX_STAGES];
  int nstages;
  int B, L, C;
  int tiles_per_batch;    // ceil(L / 128)
  int dbg_flags;          // timing experiments only (DV_CHAIN_DBG): 1 skip GELU math, 2 skip proxy fences, 4 skip bias loads
};

// ------------------------------------------------------------------ attention (flash style, bf16 in, fp32 softmax)
struct alignas(64) AttnParams {
  CUtensorMap mapQ, mapK, mapV;                            // tcgen05 kernel: 3-D [H*D, L, B] maps (attention_tc_prepare)
  const void* q; const void* k; const void* v; void* o;   // bf16
  int ld_q, ld_k, ld_v, ld_o;                              // row strides in elements
  long long bs_q, bs_k, bs_v, bs_o;                        // batch strides in elements
  int B, H, Lq, Lk, D;
  float scale;                                             // D^-0.5
  unsigned long long* trace;   // debug: CTA (0,0,0) appends (role, tile, event, clock) records; [0] = record count
  const int* kv_len;   // per-batch number of un-masked keys (keys >= kv_len[b] get the -10000 bias), or nullptr
  // split precision: lo planes (same strides); q_lo != nullptr selects the FP32 CUDA-core kernel (csrc/attention_split.cu)
  const void* q_lo; const void* k_lo; const void* v_lo; void* o_lo;
};
// Encodes mapQ/mapK/mapV from the pointers / strides / shape already stored in p (host only, no stream work). Must be
// called whenever those change and before launch_attention_tc. Returns false with a message on failure.

// ------------------------------------------------------------------ norms
// Partial statistics layout (written by the GEMM epilogue or by gn_stats): float [B][chunks][nblk][2] holding the
// sum / sum of squares of a 32-row x 16-channel block. A GroupNorm group is any union of such blocks, also across
// the two sources of a channel concat.
struct GnStatSrc {
  const float* ptr;   // partials
  int chunks;         // 32-row chunks per batch item in the producer's row space
  int nblk;           // 16-channel blocks per chunk in the producer's column space
  int fold;           // 1: producer wrote [L/2, 2C] (folded nearest-2x upsample): block kb also lives at kb + nblk/2
  int cover;          // channels of the normalised tensor this source covers (starting where the previous one ended)
};
struct GnParams {
  const void* x0; int c0;      // first source  bf16 [B, L, c0]
  const void* x1; int c1;      // optional second source (channel concat), bf16 [B, L, c1]
  GnStatSrc st[2];
  int nst;
  int B, L, G;
  const float* gamma; const float* beta;   // [c0 + c1]
  const float* scale_shift; int ss_ld;     // optional per-(b, c): scale at [b*ss_ld + c], shift at [b*ss_ld + C + c]
  float eps; int silu;
  void* out;                   // bf16 [B, L, c0 + c1]
  const void* x0_lo; const void* x1_lo; void* out_lo;   // split precision: lo planes (out_lo != nullptr selects it)
};
// standalone statistics pass over (x0 | x1) into `scratch` (same partial layout, one source covering all channels)
// x_lo / out_lo: lo planes in split precision (both null otherwise)
// LayerNorm variants of the prompt encoder (C = 128 columns stored, statistics over the first c_real of them; the rest
// of the output row is zero): rows >= row_len[row / rows_per_batch] read as zeros when mask_in (masked_fill before the
// norm, model.py:95-97) and are written as zeros when mask_out (x * (1 - mask) after it, model3.py:429-431);
// out_f32 != nullptr: the result also goes there in fp32.
struct LnMaskParams { const int* row_len; int rows_per_batch; int c_real; int mask_in; int mask_out; float* out_f32; };

// ------------------------------------------------------------------ elementwise / small
// [B, Csrc, T] fp32 -> channel-last bf16 [B, T, ld_dst]; channel c goes to column c + (c >= split ? shift : 0).
// dst_lo != nullptr (split precision): the lo plane bf16(x - float(hi)) goes there.
// [B, T, C] fp32 -> bf16 (same layout)

struct TembParams {
  const float* t; int t_stride;          // model time per batch item (t_stride 0 = broadcast)
  const float* freqs;                    // [sin_dim/2] sinusoid frequencies exp(-ln(1e4) j / half)
  const float* w1; const float* b1;      // [512,128] (nn.Linear layout [out][in]), [512]
  const float* w2; const float* b2;      // [512,512], [512]
  const float* aug;                      // [B,512] addition embedding (or nullptr)
  float* emb_base;                       // scratch [B or 1, 512]: time MLP output before the text embedding is added
  const float* w_all; const float* b_all;  // [512, n_all] (all time_emb_proj stacked, TRANSPOSED), [n_all]
  float* ss_all;                         // out [B, n_all]
  int B, n_all, sin_dim, emb_dim;
};

// sampler updates (fp32, [n] elements)
// x_out = a*x + b*m1 + c*m0 ; if roundtrip: m1 is first replaced by the reference's x0->eps->x0 round trip
// m1' = (x - sig*((x - alp*m1)/sig))/alp  and written back.
// UniPC predictor: xpred = a*x + b*m1 + cp*(m0 - m1) ; base = a*x + b*m1 + cc*(m0 - m1)
// UniPC corrector: x = base + d*(mt - m1)

}  // namespace dv

// Launch-level API of the kernel sources (csrc/kernels_api.inc). Every kernel source is compiled twice: BF16 operands
// (namespace dv) and, with -DDV_F16, FP16 operands (namespace dv16, precision mode "fp16s"); the parameter structs above are
// shared, the engine picks the set of functions by precision (engine.h: KernelApi).
namespace dv {
#include "kernels_api.inc"
}
namespace dv16 {
using namespace dv;
#include "kernels_api.inc"
}
