/*
 * cleanumamba_b200 -- C ABI of libcleanumamba_sm100.so (sm_100a only).
 *
 * This is the drop-in boundary for the CleanUMamba hot path (forward / streaming of
 * /root/reference/src/network/CleanUMamba.py:252-490).  The reference has NO native code: below its
 * Python it calls ATen (cuDNN/cuBLAS) and the third-party operators of mamba-ssm==1.2.2 /
 * causal-conv1d==1.1.0 (environment.yml:29-30).  Each entry point below names the reference-side
 * operator(s) it replaces.  Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's allocator); the library never
 *     allocates persistent device memory;
 *   - every call is asynchronous and ordered on the cudaStream_t passed as `stream` (a `void*` here
 *     so the header needs no CUDA include);
 *   - return 0 on success, a negative CUM_E* code otherwise; never throws; cum_last_error() returns a
 *     thread-local message for the most recent failure on the calling thread;
 *   - callable from any host thread; every call works on the device that is CURRENT in the calling thread
 *     (cudaSetDevice / torch.cuda.device guard) -- the library never changes the current device.  Per-device
 *     state (SM count, raised shared-memory limits) is keyed by device ordinal, so one process may drive
 *     several GPUs; the only process-wide mutable state is the tensor-map cache and that attribute set,
 *     both behind a mutex and emptied by cum_shutdown().
 *
 * DATA LAYOUT.  Activations are channels-last: a (batch, time, channels) fp32 array whose channel
 * count is padded to a multiple of 8 (pad lanes are kept at 0 by zero-padded packed weights).  The
 * reference's NCL tensors only exist at the waveform ends, where C == 1 and both layouts coincide.
 * The carried caches of feed() (CleanUMamba.py:432-442,476-484) are kept time-major, (column, stream, channel), by the
 * streaming host (cleanumamba_b200/stream_tm.py): see "ABI v3" below.
 *
 * ABI HISTORY.  v1: offline forward, streaming, backward.  v2: cum_shutdown, fused first / last U-Net blocks, segment-parallel scan
 * workspace, device-side gradient scales, fused MR-STFT loss kernels, channel importances.  v3: plane-major tap-GEMM operands
 * (a_planes / n_half) and the small-M path (small_m_path) in cum_gemm_desc, fp16 carried state (state_f16) in cum_scan_desc, strided
 * conv_in / convt_out / dwconv_silu, cum_stream_shift_fwd.  Descriptors only grow at their end; zero-initialised new fields select
 * the previous behaviour.
 */
#ifndef CLEANUMAMBA_B200_H_
#define CLEANUMAMBA_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define CUM_ABI_VERSION 3

/* error codes */
#define CUM_OK            0
#define CUM_EINVAL      (-22)   /* bad argument (shape / alignment / enum) */
#define CUM_ENOTSUP     (-95)   /* valid request this build cannot serve */
#define CUM_ECUDA       (-5)    /* CUDA runtime / driver error; see cum_last_error() */

/* epilogues of cum_gemm_bias_act_fwd: v = acc + bias[n] */
#define CUM_EPI_NONE         0  /* out[n]   = v[n]                                   */
#define CUM_EPI_RELU         1  /* out[n]   = max(v[n], 0)                            */
#define CUM_EPI_SILU         2  /* out[n]   = v[n] * sigmoid(v[n])                    */
#define CUM_EPI_GLU_SIGMOID  8  /* out[n/2] = v[2c] * sigmoid(v[2c+1])   (layers.py:26-34; W rows interleaved) */
#define CUM_EPI_GLU_RELU     9
#define CUM_EPI_GLU_SILU    10
#define CUM_EPI_GLU_GELU    11

/* arithmetic of the contraction */
#define CUM_MATH_FP32        0  /* CUDA-core FFMA, exact fp32 products (reference arithmetic) */
#define CUM_MATH_TF32X3      1  /* tcgen05 kind::tf32, hi/lo split, 3 MMAs per product (~2^-21 rel. error) */
#define CUM_MATH_TF32        2  /* tcgen05 kind::tf32, single pass (10-bit mantissa; NOT within the fp32 tolerance) */
#define CUM_MATH_BF16X3      3  /* tcgen05 kind::f16 on bf16 hi/lo halves, 3 MMAs per product (~2^-16 rel. error per
                                   product; 2x the TF32X3 tensor rate; marginal (1.1e-4) at full-scale amplitude) */
#define CUM_MATH_BF16        5  /* REDUCED PRECISION variant (reported separately): `a` and `w` are bf16 arrays, one
                                   kind::f16 MMA pass, fp32 accumulate; ~3e-3 relative error per layer */
#define CUM_MATH_F16X3       4  /* tcgen05 kind::f16 on fp16 hi/lo halves (11+11 bits, ~2^-21 per product like TF32X3)
                                   at the bf16 tensor rate; weights pre-scaled by a power of two (cum_split_f16),
                                   undone by acc_scale; activations converted with saturation at +-65504 */

typedef void* cum_stream_t;    /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------------- */
int         cum_abi_version(void);
int         cum_init(int device);              /* probes `device` (fails unless it is sm_100); does not make it current */
int         cum_shutdown(void);                /* drops the library's caches (TMA descriptor cache, per-device kernel
                                                  attributes); the library stays usable, caches refill lazily */
const char* cum_last_error(void);

/* ---- waveform ends -------------------------------------------------------------------------- */
/* Replaces `std = x.std(dim=2, keepdim=True) + 1e-3; x /= std` (CleanUMamba.py:260-262, in place on the
 * caller's tensor, unbiased std).  x: (batch, length) contiguous; std_out: (batch). */
int cum_wave_normalize_fwd(float* x, float* std_out, int batch, int length, cum_stream_t stream);

/* Replaces F.pad (CleanUMamba.py:219-223) + encoder[0][0] Conv1d(1,H,K,S) + ReLU (:109-110).
 * x: (batch, length) [row stride x_stride]; samples t >= length read as 0.  w: (K, c_pad) taps-major;
 * bias: (c_pad).  y: (batch, rows_out, c_pad) channels-last,
 *   y[b,t,c] = relu(b[c] + sum_k w[k,c] * x[b,S t+k] / in_scale[b, max(0, t + row_offset) / group_rows]).
 * in_scale (batch, ceil(max(1, rows_out + row_offset) / group_rows)) may be NULL (no scaling); it carries the per-hop
 * running std of the streaming path (`frame / self.input_std`, CleanUMamba.py:399-401); row_offset < 0 lets the
 * first (longer) frame of a stream map to group 0. */
int cum_conv_in_fwd(const float* x, long long x_stride, int batch, int length, const float* w,
                    const float* bias, float* y, int rows_out, int c_pad, int kernel, int stride,
                    const float* in_scale, int group_rows, int row_offset, cum_stream_t stream);

/* Replaces the per-frame running std of feed() (CleanUMamba.py:399-401): for frame j of this call,
 *   s = std(x[b, j*hop : j*hop+frame_len], unbiased) + 1e-3;  f = frames_before + j + 1;
 *   running[b] = s / f + (1 - 1/f) * running[b];  scale_out[b, j] = running[b].
 * x: (batch, >= (frames-1)*hop + frame_len) raw pending samples [row stride x_stride]. */
int cum_stream_std_fwd(const float* x, long long x_stride, int batch, int frames, int frame_len, int hop,
                       int frames_before, float* running, float* scale_out, cum_stream_t stream);
/* Same, with the frame count kept on the device: frames_before = *frames_counter, and *frames_counter += frames afterwards
 * (stream-ordered).  No host-side argument changes between calls, so a captured CUDA graph of a streaming step can be replayed. */
int cum_stream_std_counter_fwd(const float* x, long long x_stride, int batch, int frames, int frame_len, int hop,
                               int* frames_counter, float* running, float* scale_out, cum_stream_t stream);

/* Replaces decoder[-1][2] ConvTranspose1d(H,1,K,S) (CleanUMamba.py:124) + the crop and `* std` (:318-319; per hop
 * `out *= self.input_std` :406-407 when streaming).  g: (batch, rows_in, c_pad) channels-last; w: (K, c_pad);
 * out: (batch, length) [row stride out_stride].  With full[m] = bias + sum_{j,k: S j + k = m} <g[b,j,:], w[k,:]>:
 *   out[b,i] = scale[b, i / scale_group] * full[first + i],  i < length.
 * scale (batch, ceil(length / scale_group)) may be NULL.  `first` > 0 skips the outputs already emitted when g
 * carries one column of history (streaming overlap-add, :476-484). */
int cum_convt_out_fwd(const float* g, int batch, int rows_in, int c_pad, const float* w, float bias,
                      const float* scale, int scale_group, float* out, long long out_stride, int first,
                      int length, int kernel, int stride, cum_stream_t stream);

/* bf16-storage variants of the two kernels above for the reduced-precision model variant (BASELINE.json configs[1]
 * "fp32 and bf16"): cum_conv_in_bf16_fwd writes y as bf16, cum_convt_out_bf16_fwd reads g as bf16; fp32 arithmetic. */
int cum_conv_in_bf16_fwd(const float* x, long long x_stride, int batch, int length, const float* w,
                         const float* bias, void* y_bf16, int rows_out, int c_pad, int kernel, int stride,
                         cum_stream_t stream);
int cum_convt_out_bf16_fwd(const void* g_bf16, int batch, int rows_in, int c_pad, const float* w, float bias,
                           const float* scale, int scale_group, float* out, long long out_stride, int first,
                           int length, int kernel, int stride, cum_stream_t stream);
/* cum_conv_in_fwd writing its output as fp16 hi / lo planes ("hl16", see cum_gemm_desc.a_lo): the first encoder GEMM then needs
 * no operand splitter.  Same arithmetic as cum_conv_in_fwd; y_hi[i] + y_lo[i] carries 22 bits of the fp32 result. */
int cum_conv_in_hl16_fwd(const float* x, long long x_stride, int batch, int length, const float* w,
                         const float* bias, void* y_hi, void* y_lo, int rows_out, int c_pad, int kernel, int stride,
                         cum_stream_t stream);

/* ---- ABI v3: the time-major streaming session (cleanumamba_b200/stream_tm.py) ------------------------------------
 * feed() for MANY concurrent streams (CleanUMamba.py:371-490 once per stream and hop in the reference) keeps every carried
 * buffer as (column, stream, channel): all streams of a hop sit in the M dimension of each GEMM (full 128-row tiles), a FIFO is
 * appended / consumed by whole planes, and no per-call gather / scatter copies are left.  The kernels below are the strided
 * forms of cum_conv_in_fwd / cum_convt_out_fwd / cum_dwconv_silu_fwd (element [b, t, :] of the channels-last side at
 * base + b * batch_stride + t * row_stride) and the one FIFO-maintenance launch that ends a call. */
int cum_conv_in_strided_fwd(const float* x, long long x_stride, int batch, int length, const float* w, const float* bias,
                            float* y, long long y_batch_stride, long long y_row_stride, int rows_out, int c_pad, int kernel,
                            int stride, const float* in_scale, int group_rows, int row_offset, cum_stream_t stream);
int cum_convt_out_strided_fwd(const float* g, long long g_batch_stride, long long g_row_stride, int batch, int rows_in, int c_pad,
                              const float* w, float bias, const float* scale, int scale_group, float* out, long long out_stride,
                              int first, int length, int kernel, int stride, cum_stream_t stream);
/* For every entry and row r < rows: base[r * row_stride + i] = base[r * row_stride + src_off + i], i < count (elements; the
 * ranges may overlap) -- the unconsumed tail of a carried FIFO moves to its front.  Up to 24 entries per launch. */
typedef struct cum_shift_entry { float* base; long long row_stride; long long src_off; long long count; int rows; int reserved; } cum_shift_entry;
int cum_stream_shift_fwd(const cum_shift_entry* entries, int n_entries, cum_stream_t stream);

/* ---- the tap-GEMM: every dense contraction of the path -------------------------------------- */
/* out[b, m, :] = EPI( bias + sum_{s<taps} W_s . a[b, m + tap_shift[s], 0:k] ) (+ addend[b, m, :])
 * Rows of `a` outside [0, a_rows) read as zero.  One descriptor covers:
 *   - nn.Conv1d(k=1) / nn.Linear (taps=1): encoder[i][2], decoder[j][0], tsfm_conv1/2 (CleanUMamba.py:111,122,
 *     139,194) and Mamba in_proj / x_proj / dt_proj / out_proj (mamba_ssm Mamba.forward);
 *   - nn.Conv1d(k=4, s=2) (taps=2, shifts {0,+1} over the (L/2, 2C) view of the input; :109);
 *   - nn.ConvTranspose1d(k=4, s=2) (taps=2, shifts {0,-1}, output seen as (Lin+1, 2Cout); :124);
 *   with ReLU / GLU (layers.py) / U-Net skip add (:315) fused as epilogue + addend. */
typedef struct cum_gemm_desc {
    const float* a;  long long a_batch_stride; long long a_row_stride;  /* elements */
    int a_rows;      /* valid rows per batch item */
    int k;           /* contraction length per tap (multiple of 4) */
    int taps;        /* 1 or 2 */
    int tap_shift[2];
    const float* w;  /* packed (taps, n, ldw), K contiguous */
    int ldw;         /* >= k, multiple of 4 */
    const float* bias;       /* (n) or NULL */
    float* c;        long long c_batch_stride; long long c_row_stride;
    int m;           /* output rows per batch item */
    int n;           /* weight rows (multiple of 8); GLU epilogues write n/2 columns */
    int batch;
    int epilogue;    /* CUM_EPI_* */
    const float* addend; long long add_batch_stride; long long add_row_stride;  /* optional */
    int math;        /* CUM_MATH_* */
    const float* w_lo;       /* TF32X3 / BF16X3 only: low halves of the weights; `w` must then hold the high halves (both
                                produced by cum_split_tf32 / cum_split_bf16 / cum_split_f16; 16-bit arrays for *16X3) */
    float acc_scale;         /* CUM_MATH_F16X3 only: accumulators are multiplied by this before the bias (1 / weight scale) */
    int out_bf16;            /* BF16 / F16X3 only: `c` and `addend` are bf16 arrays (strides still in elements) */
    int w_lo_is_zero;        /* split modes: caller asserts every element of w_lo is exactly 0 (e.g. weights of a checkpoint
                                shipped in fp16 under F16X3): the a_hi*w_lo pass is skipped -- identical result, 2 MMAs / product */
    /* CUM_MATH_F16X3 only -- "hl16" activations: a tensor stored as TWO fp16 planes, value = hi + lo with hi = fp16(x) and
     * lo = fp16(x - hi) (what the kernel's operand splitter computes from fp32 on the fly).  A layer that writes hl16 (c_lo set)
     * splits each output ONCE in its epilogue; the next layer (a_lo set) then feeds both planes to the tensor cores straight from
     * TMA: same products as with fp32 activations, no in-kernel splitter.  Strides are in elements and shared by both planes. */
    const void* a_lo;        /* NULL, or the low-half plane of `a` (then `a` is the high-half plane, both fp16) */
    void* c_lo;              /* NULL, or the low-half plane of `c` (then `c` is the high-half plane, both fp16) */
    const void* addend_lo;   /* NULL, or the low-half plane of `addend` (then `addend` is the high-half plane, both fp16) */
    const float* a_scale_dev;/* CUM_MATH_F16X3 with fp32 `a` only: NULL, or a DEVICE pointer to {s, 1/s} (cum_grad_scale_fwd): `a` is multiplied
                                by s while it is split into fp16 halves and the accumulator by 1/s -- back-propagated gradients
                                (~1e-6) are lifted into fp16's range by a power of two computed on the device, no host sync */
    int addend_is_mask;      /* tensor-core modes, fp32 output: `addend` is a ReLU MASK instead -- out = addend > 0 ? value : 0 (the ReLU backward
                                of a data-gradient GEMM; with `aux` the unmasked value is kept as well) */
    float* aux; long long aux_batch_stride; long long aux_row_stride;
                             /* optional SECOND fp32 output of a tensor-core call (training keeps what the backward needs without a
                                separate elementwise kernel): with a GLU epilogue the (m, n) PRE-ACTIVATION (bias added, before the
                                gate; row stride aux_row_stride >= n); with NONE / RELU the (m, n) value BEFORE the addend.  fp32
                                output only; NULL = off */
    int cta_pair;            /* tiles wider than 128 columns can run on CTA pairs (tcgen05 cta_group::2: 256-row tiles, each CTA
                                stages half of the weight tile).  0 = automatic (pairs whenever a problem has more than 128 rows),
                                1 = same, -1 = never.  Same products, same accumulation order: bit-identical results */
    /* ABI v3 -- PLANE-MAJOR `a` (time-major streaming: the carried encoder / decoder FIFOs of feed(), CleanUMamba.py:432-442,476-484,
     * are stored as (column, stream, channel) so that every GEMM of a hop has all streams in its M dimension and a FIFO is appended /
     * consumed by whole contiguous planes).  a_planes > 0: `a` is a stack of a_planes planes of (a_rows, a_plane_k) values
     * [plane stride a_batch_stride, row stride a_row_stride]; output row m of batch item b reads, for tap s and K offset kk < k,
     *     plane  a_plane0 + a_plane_step * (b' + tap_shift[s]) + kk / a_plane_k,  row m,  channel kk % a_plane_k      (b' = b, or b >> 1 with n_half)
     * (planes outside [0, a_planes) read as zero): the tap shifts move across PLANES instead of rows, and the 2C-wide input row of a
     * strided conv (k = 2 a_plane_k, a_plane_step = 2) is two neighbouring planes.  a_plane_k must be a multiple of 32 (64 for CUM_MATH_BF16).
     * n_half (plane mode, non-GLU epilogues): the packed weight holds 2 n rows per tap and batch item b uses rows
     * (b & 1) * n .. + n, `bias` (n entries) is shared by both halves: a transposed conv writes its even / odd output columns as
     * separate planes (c_batch_stride apart). */
    int a_planes; int a_plane_k; int a_plane0; int a_plane_step; int n_half;
    int small_m_path;        /* problems of a few output rows in total (a handful of streams fed hop by hop -- the reference's real-time
                                use) run on a CUDA-core kernel instead of the persistent tensor-core pipeline, whose fixed cost (~14 us
                                per launch) dominates there: exact fp32 FMAs on the full-precision weights (hi + lo halves in the split
                                modes).  0 = automatic: at most 4 output rows in total and 16 M multiply-adds;
                                1 = force (any size), -1 = never.  Not for bf16 / hl16 operands */
} cum_gemm_desc;
int cum_gemm_bias_act_fwd(const cum_gemm_desc* desc, cum_stream_t stream);

/* TF32 hi/lo split of a packed weight array for CUM_MATH_TF32X3: hi = w with the 13 low mantissa bits cleared,
 * lo = w - hi (exact in fp32).  Same bit arithmetic as the in-kernel split of the activations. */
int cum_split_tf32(const float* w, float* hi, float* lo, long long count, cum_stream_t stream);
/* bf16 hi/lo split for CUM_MATH_BF16X3: hi = bf16_rn(w), lo = bf16_rn(w - hi); hi / lo are bf16 arrays (2 bytes per
 * element) that are then passed as `w` / `w_lo` (ldw in elements, multiple of 8). */
int cum_split_bf16(const float* w, void* hi, void* lo, long long count, cum_stream_t stream);
/* fp16 hi/lo split of scale*w for CUM_MATH_F16X3 (scale = a power of two, typically 2^floor(log2(8/max|w|))); pass
 * acc_scale = 1/scale in the descriptor. */
int cum_split_f16(const float* w, void* hi, void* lo, long long count, float scale, cum_stream_t stream);

/* ---- fused U-Net blocks (CUM_MATH_F16X3 arithmetic) --------------------------------------------------------------- */
/* First encoder block as ONE kernel (CleanUMamba.py:108-113 for i = 0, with F.pad of :263):
 *   y[t,c]   = relu(conv_b[c] + sum_k conv_w[k,c] x[b, 2t+k])            Conv1d(1, 64, 4, 2) + ReLU, x read as 0 beyond `length`
 *   out[t,:] = GLU(glu_b + W y[t,:])                                       Conv1d(64, 128, 1) + layers.Activation("Sigmoid")
 * The 64-channel intermediate never reaches HBM (it was a 2 x 1.3 GB round trip per 64 x 10 s batch).  Same products and
 * accumulation order as cum_conv_in_fwd followed by cum_gemm_bias_act_fwd(CUM_MATH_F16X3, CUM_EPI_GLU_SIGMOID).
 * channels (conv outputs) and channels_out (GLU outputs) are padded counts <= 64, multiples of 8 (64 / 64 for the shipped
 * channels_H; pruned checkpoints are narrower): the kernel works on a zero-padded 64 x 128 tile.  Wider: CUM_EINVAL -- use the
 * two calls. */
typedef struct cum_enc0_block_desc {
    const float* x; long long x_stride; int batch; int length;   /* (batch, length) waveform, already normalised */
    const float* conv_w;     /* (4, channels) taps-major */
    const float* conv_b;     /* (channels) */
    const void* glu_w_hi;    /* (2 channels_out, channels) fp16, rows interleaved (a_c, b_c), scaled: cum_split_f16 */
    const void* glu_w_lo;    /* low halves (may be NULL when w_lo_is_zero) */
    const float* glu_b;      /* (2 channels_out) interleaved */
    float acc_scale;         /* 1 / weight scale */
    int w_lo_is_zero;
    float* out;              /* (batch, rows_out, channels_out) fp32 channels-last: the level-0 skip */
    int rows_out;            /* (padded_length - 4) / 2 + 1 */
    int channels;            /* conv output channels, padded (<= 64, multiple of 8) */
    int channels_out;        /* GLU output channels, padded (<= 64, multiple of 8) */
} cum_enc0_block_desc;
int cum_enc0_block_fwd(const cum_enc0_block_desc* desc, cum_stream_t stream);

/* Last decoder block as ONE kernel (CleanUMamba.py:121-128 for the last level, crop + de-normalisation of :318-319):
 *   g[p,:]      = GLU(glu_b + W a[b,p,:])                                  Conv1d(64, 128, 1) + layers.Activation("Sigmoid")
 *   out[b,2p+k] = (convt_bias + <g[p], convt_w[k]> + <g[p-1], convt_w[k+2]>) * scale[b]     ConvTranspose1d(64, 1, 4, 2), k = 0, 1
 * for 0 <= 2p+k < out_length (<= 2 rows_in + 2).  Replaces cum_gemm_bias_act_fwd + cum_convt_out_fwd; the gated 64-channel
 * tensor never reaches HBM. */
typedef struct cum_dec_last_block_desc {
    const float* a; int batch; int rows_in;   /* (batch, rows_in, channels) fp32 channels-last (skip already added) */
    const void* glu_w_hi; const void* glu_w_lo; const float* glu_b; float acc_scale; int w_lo_is_zero;   /* (2 channels_gated, channels) */
    const float* convt_w;    /* (4, channels_gated) taps-major */
    float convt_bias;
    const float* scale;      /* (batch) per-clip std, or NULL */
    float* out; long long out_stride; int out_length;
    int channels;            /* input channels, padded (<= 64, multiple of 8) */
    int channels_gated;      /* GLU output channels, padded (<= 64, multiple of 8) */
} cum_dec_last_block_desc;
int cum_dec_last_block_fwd(const cum_dec_last_block_desc* desc, cum_stream_t stream);

/* ---- Mamba block operators ------------------------------------------------------------------- */
/* Replaces Block.forward's `residual = h + residual; h = LayerNorm(residual)` (mamba_ssm Block, non-fused
 * branch) and the final add + norm_f (CleanUMamba.py:292-294).  All (rows, c_pad) with `c` real channels.
 * residual_in may be NULL (first block); residual_out may alias residual_in. */
int cum_ln_residual_fwd(const float* h, const float* residual_in, float* residual_out, float* normed,
                        const float* gamma, const float* beta, float eps, long long rows, int c, int c_pad,
                        cum_stream_t stream);

/* Replaces causal_conv1d_fn(x, w(d,width), b, "silu") == silu(conv1d(x)[..., :L]) (Mamba.forward) and, with
 * state pointers, causal_conv1d_update / the roll-and-sum of Mamba.step.  x: rows of stride x_row_stride inside a
 * (batch, len, *) array; w: (width, d_pad) taps-major; y: (batch, len, d_pad).  conv_state (batch, width-1, d_pad):
 * the inputs preceding t=0 (NULL = zeros); conv_state_out receives the last width-1 inputs (may alias). */
int cum_dwconv_silu_fwd(const float* x, long long x_batch_stride, long long x_row_stride, const float* w,
                        const float* bias, float* y, const float* conv_state, float* conv_state_out,
                        int batch, int len, int d_pad, int width, cum_stream_t stream);
int cum_dwconv_silu_strided_fwd(const float* x, long long x_batch_stride, long long x_row_stride, const float* w,
                                const float* bias, float* y, long long y_batch_stride, long long y_row_stride,
                                const float* conv_state, float* conv_state_out, int batch, int len, int d_pad, int width,
                                cum_stream_t stream);

/* Replaces selective_scan_fn(u, delta, A, B, C, D, z, delta_bias, delta_softplus=True) (mamba_ssm; oracle =
 * selective_scan_ref) and, with h0/h_out, selective_state_update / Mamba.step's recurrence:
 *   dl = softplus(delta + delta_bias);  h_t = exp(dl A) h_{t-1} + (dl u_t) B_t;  y_t = (<h_t, C_t> + D u_t) silu(z_t)
 * Channels-last operands, each described by (pointer, batch stride, row stride) in elements:
 *   u, delta, z, y: (batch, len, d);  Bm, Cm: (batch, len, n_state);  a2 = -exp(A_log) * log2(e): (d, n_state);
 *   h0 / h_out: (batch, d, n_state) or NULL (zero initial state / state not returned; may alias). */
typedef struct cum_scan_desc {
    const float* u;     long long u_bs, u_rs;
    const float* delta; long long dl_bs, dl_rs;
    const float* z;     long long z_bs, z_rs;      /* z may be NULL (no gate) */
    const float* Bm;    long long B_bs, B_rs;
    const float* Cm;    long long C_bs, C_rs;
    float* y;           long long y_bs, y_rs;
    const float* a2;    /* (d, n_state) row-major */
    const float* Dskip; /* (d) or NULL */
    const float* delta_bias; /* (d) or NULL */
    const float* h0; float* h_out;
    int batch, len, d, n_state;
    int delta_softplus;
    float* h_ckpt;      /* optional (training): (batch, ceil(len/16), n_state, d) -- h at the start of every 16-step chunk (channel
                           fastest: a warp's 32 channels store / load one 128-byte run per state),
                           consumed by cum_selective_scan_bwd */
    void* workspace;    /* optional scratch (16-byte aligned) of cum_selective_scan_workspace_bytes(desc) bytes: with it, SMALL
                           batches (fewer than SM-count/2 CTAs of 64 channels, len >= 256) run segment-parallel -- the clip is cut
                           into time segments scanned concurrently from h = 0, a carry pass chains their end states
                           (h is linear in its start state), a second pass writes y from the true start states: 2x the
                           arithmetic, up to 64-fold parallelism.  NULL: always the time-sequential kernel */
    long long workspace_bytes;
    int state_f16;      /* ABI v3 -- streaming variant with a REDUCED-PRECISION carried state (reported separately): h0 / h_out point to
                           fp16 arrays of the same shape (half the state traffic that bounds a 1-hop call); the recurrence itself
                           stays fp32.  Needs n_state = 64, d % 16 == 0, both h0 and h_out */
} cum_scan_desc;
int cum_selective_scan_fwd(const cum_scan_desc* desc, cum_stream_t stream);
/* 0 when the problem would not run segment-parallel (large batch, short sequence, training checkpoints requested) */
long long cum_selective_scan_workspace_bytes(const cum_scan_desc* desc);

/* ---- backward (training; autograd of CleanUMamba.forward, src/training/train.py:278-285) ------------------- */
/* In training the GLU gate and the skip add run as separate kernels so the pre-activation is saved once:
 *   glu_fwd : out[r,c] = z[r,2c] * sigmoid(z[r,2c+1]) (+ addend[r,c]);  z: (rows, 2 h_pad) interleaved
 *   glu_bwd : dz from dout (rows, h_pad); dbias (2 h_pad) += column sums of dz (atomics; may be NULL)
 *   relu_bwd: dz = dy * (y > 0); dbias (cols) += column sums.     colsum: dbias += column sums of d.
 *   add     : out = a + b (count elements, multiple of 4). */
int cum_glu_fwd(const float* z, const float* addend, float* out, long long rows, int h_pad, cum_stream_t stream);
/* dz_scale4: NULL, or 4 floats on the device that receive {s, 1/s, -, -} of dz like cum_grad_scale_fwd (the max is taken while dz is
 * produced: no separate pass before the f16x3 gradient GEMMs). */
int cum_glu_bwd(const float* z, const float* dout, float* dz, float* dbias, long long rows, int h_pad, float* dz_scale4,
                cum_stream_t stream);
int cum_relu_bwd(const float* y, const float* dy, float* dz, float* dbias, long long rows, int cols, float* dz_scale4,
                 cum_stream_t stream);
int cum_colsum(const float* d, float* dbias, long long rows, int cols, float* d_scale4, cum_stream_t stream);
int cum_add_fwd(const float* a, const float* b, float* out, long long count, cum_stream_t stream);

/* Weight gradient of the tap-GEMM:  dw[s, n, k] += sum_{b, row < m} dz[b,row,n] * a[b, row + tap_shift[s], k]
 * (rows of `a` outside [0, a_rows) read as zero; dw is accumulated with atomics -> zero it first).  The DATA gradient
 * of the tap-GEMM is cum_gemm_bias_act_fwd itself with transposed packed weights and negated shifts. */
typedef struct cum_wgrad_desc {
    const float* dz; long long dz_batch_stride; long long dz_row_stride;
    const float* a;  long long a_batch_stride;  long long a_row_stride;
    int a_rows;
    float* dw; int ldw;          /* (taps, n, ldw) */
    int m, n, k, taps;
    int tap_shift[2];
    int batch;
    int math;            /* CUM_MATH_FP32: CUDA-core FFMA (no workspace).  Any tensor-core mode: tcgen05 split-K GEMM over the
                            rows: MN-major fp16 hi / lo operands of the device-scaled gradient (no transposes, f16 tensor rate);
                            CUM_WGRAD_MN=0 selects the previous transposing TF32X3 path */
    void* workspace;     /* tensor-core mode: >= cum_gemm_wgrad_workspace_bytes(desc) bytes of device scratch */
    const float* dz_scale_dev;   /* optional DEVICE pointer to {s, 1/s} of dz from cum_grad_scale_fwd (saves the call its own amax pass) */
} cum_wgrad_desc;
int       cum_gemm_wgrad(const cum_wgrad_desc* desc, cum_stream_t stream);
/* scale4[0] = s = 2^(15 - e) with max|x| = f 2^e, f in [0.5, 1) (1 when x is all zero), scale4[1] = 1 / s; scale4[2..3] scratch.
 * x: (batch, rows, cols) with the given strides, cols % 4 == 0, rows 16-byte aligned.  The power-of-two "loss scale" of ONE
 * gradient tensor, left on the device for cum_gemm_desc.a_scale_dev / cum_wgrad_desc.dz_scale_dev. */
int cum_grad_scale_fwd(const float* x, long long batch_stride, long long row_stride, int batch, int rows, int cols, float* scale4,
                       cum_stream_t stream);
long long cum_gemm_wgrad_workspace_bytes(const cum_wgrad_desc* desc);

/* LayerNorm backward + residual-stream add: x = saved LN input (rows, c_pad); dy = grad of the normalised output;
 * dres_in (may be NULL) = gradient already in the residual stream; dx = LN_bwd(dy) + dres_in;
 * dgamma / dbeta (c_pad) accumulated with atomics. */
int cum_ln_residual_bwd(const float* x, const float* dy, const float* dres_in, const float* gamma, float* dx,
                        float* dgamma, float* dbeta, float eps, long long rows, int c, int c_pad, cum_stream_t stream);

/* Backward of cum_dwconv_silu_fwd (width 4, zero initial state): dx (strided like x), dw (width, d_pad) and db (d_pad)
 * accumulated with atomics. */
int cum_dwconv_silu_bwd(const float* x, long long x_batch_stride, long long x_row_stride, const float* w,
                        const float* bias, const float* dy, float* dx, long long dx_batch_stride,
                        long long dx_row_stride, float* dw, float* db, int batch, int len, int d_pad, int width,
                        cum_stream_t stream);

/* Weight / bias gradients of cum_conv_in_fwd (no data gradient: the waveform is not a parameter); y = saved output. */
int cum_conv_in_bwd(const float* x, long long x_stride, int batch, int length, const float* y, const float* dy,
                    float* dw, float* db, int rows_out, int c_pad, int kernel, int stride, cum_stream_t stream);

/* Backward of cum_convt_out_fwd (first = 0, scale per batch item): dg (batch, rows_in, c_pad) written; dw (K, c_pad)
 * and dbias (1) accumulated with atomics. */
int cum_convt_out_bwd(const float* g, int batch, int rows_in, int c_pad, const float* w, const float* scale,
                      const float* dout, long long dout_stride, int length, float* dg, float* dw, float* dbias,
                      int kernel, int stride, cum_stream_t stream);

/* Reverse selective scan.  `fwd` = the descriptor of the forward call (y / h0 / h_out ignored); h_ckpt = the
 * checkpoints that call wrote.  du, ddelta, dz are written; dB, dC (batch, len, n_state), dA_log (d, n_state), dD (d),
 * ddelta_bias (d) are accumulated with atomics -> zero them first.  ddelta is the gradient w.r.t. the RAW delta
 * (before bias and softplus). */
typedef struct cum_scan_bwd_desc {
    cum_scan_desc fwd;
    const float* h_ckpt;
    const float* dout; long long dout_bs; long long dout_rs;
    float* du;     long long du_bs;  long long du_rs;
    float* ddelta; long long ddl_bs; long long ddl_rs;
    float* dz;     long long dz_bs;  long long dz_rs;
    float* dB;     long long dB_bs;  long long dB_rs;
    float* dC;     long long dC_bs;  long long dC_rs;
    float* dA_log; float* dD; float* ddelta_bias;
} cum_scan_bwd_desc;
int cum_selective_scan_bwd(const cum_scan_bwd_desc* desc, cum_stream_t stream);

/* ---- multi-resolution STFT loss (src/util/stft_loss.py:16-184, util.py:322) --------------------------------- */
/* One resolution of the loss = a windowed DFT run as a tap-GEMM on the tensor cores (cum_gemm_bias_act_fwd with the hann window
 * folded into a (2F, win_length) cos / -sin basis, re / im interleaved) between the kernels below.
 * frames[(s, b, fr), k] = sig_s[b, reflect(fr hop + k + (n_fft - win)/2 - n_fft/2)]: torch.stft's center=True reflect padding by index
 * arithmetic; s = 0 predicted (x), 1 target (y; NULL = one signal); frames: ((1|2) batch n_frames, win) fp32. */
int cum_stft_frames_fwd(const float* x, const float* y, long long sig_stride, int length, int batch, int n_frames, int hop, int win,
                        int n_fft, float* frames, cum_stream_t stream);
/* sums[0..2] += sum (ym - xm)^2, sum ym^2, sum |log ym - log xm| over rows x bins, m = sqrt(max(re^2 + im^2, 1e-7)) (stft_loss.py:35,
 * :57, :80).  sx / sy: (rows, ld) spectrograms, (re, im) of bin f at columns 2f, 2f+1.  sums: 3 doubles, zeroed by the caller. */
int cum_stft_loss_reduce_fwd(const float* sx, const float* sy, long long rows, int bins, int ld, double* sums, cum_stream_t stream);
/* dsx = coef[0] d(sum (ym-xm)^2 / 2)/dsx ... : gradient of the two loss terms w.r.t. the predicted spectrogram,
 * dxm = -coef[0] (ym - xm) - coef[1] sign(log ym - log xm) / xm, d(re, im) = dxm (re, im) / xm (0 under the floor); coef: 2 floats on
 * the device (coef[0] = k_sc / (sqrt(A) sqrt(B)), coef[1] = k_mag / count).  dsx may alias sx.  Pad columns are zeroed. */
int cum_stft_loss_bwd(const float* sx, const float* sy, long long rows, int bins, int ld, const float* coef, float* dsx,
                      cum_stream_t stream);
/* dx[b, reflect(...)] += dframes[(b, fr), k]: transpose of cum_stft_frames_fwd for the predicted signal (atomic accumulation). */
int cum_stft_overlap_add(const float* dframes, int length, int batch, int n_frames, int hop, int win, int n_fft, float* dx,
                         long long dx_stride, cum_stream_t stream);

/* ---- pruning support ---------------------------------------------------------------------------- */
/* Channel-importance statistics of a weight matrix w (rows, cols; row strides ldw / ldg) and its gradient g, the inputs of the
 * reference's pruning criteria (src/pruning/pruninggroup.py:160-226 `channel_importances`, importance.py:39): for every row and
 * every column, in ONE pass over w and g,
 *   [0] sum w^2 ("weight")   [1] sum g^2 ("grad")   [2] sum |w g| ("taylor_individual")
 *   [3] sum (w g)^2 ("taylor_squared_individual")   [4] sum w g (|.| of it = "taylor_group")
 * out_rows: (5, rows) or NULL; out_cols: (5, cols) or NULL (zeroed by the call).  The caller regroups rows / columns into
 * channels (n_heads consecutive rows per channel, K taps per input channel of a conv weight). */
int cum_channel_importance_fwd(const float* w, const float* g, int rows, int cols, long long ldw, long long ldg,
                               float* out_rows, float* out_cols, cum_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CLEANUMAMBA_B200_H_ */
