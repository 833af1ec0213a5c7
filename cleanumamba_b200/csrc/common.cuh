// cleanumamba_b200 -- shared device/host helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/cleanumamba_b200.h"

namespace cum {

// thread-local last-error string (cum_last_error)
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define CUM_REQUIRE(cond, ...)                         \
    do {                                               \
        if (!(cond)) {                                 \
            ::cum::set_error(__VA_ARGS__);             \
            return CUM_EINVAL;                         \
        }                                              \
    } while (0)

#define CUM_LAUNCH_CHECK(what)                                          \
    do {                                                                \
        cudaError_t e__ = cudaGetLastError();                           \
        if (e__ != cudaSuccess) return ::cum::cuda_fail(e__, what);     \
    } while (0)

// ---- programmatic dependent launch (PDL): a kernel of the forward path is launched with "programmatic stream serialization", so
// its CTAs may start -- barrier / TMEM / tensor-map set-up, no global-memory access -- while the previous kernel of the stream is
// still draining; pdl_wait() (griddepcontrol.wait) blocks until that kernel has completed and its writes are visible, and must
// precede the first global access.  pdl_trigger() lets the NEXT kernel launch as soon as every CTA of this one has started.
// The forward of a pruned checkpoint (57 launches of ~10 us) and a streaming step (260 launches) are bound by launch gaps.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();       // CUM_PDL=0 disables (api.cu)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }
__device__ __forceinline__ float siluf_(float v) { return v / (1.0f + expf(-v)); }
__device__ __forceinline__ float geluf_(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }

// GLU gate selected by the CUM_EPI_GLU_* code (layers.py:17-24)
__device__ __forceinline__ float glu_gate(int epi, float b) {
    switch (epi) {
        case CUM_EPI_GLU_RELU: return fmaxf(b, 0.0f);
        case CUM_EPI_GLU_SILU: return siluf_(b);
        case CUM_EPI_GLU_GELU: return geluf_(b);
        default:               return sigmoidf_(b);
    }
}
__device__ __forceinline__ float unary_act(int epi, float v) {
    switch (epi) {
        case CUM_EPI_RELU: return fmaxf(v, 0.0f);
        case CUM_EPI_SILU: return siluf_(v);
        default:           return v;
    }
}
static inline bool epi_is_glu(int epi) { return epi >= CUM_EPI_GLU_SIGMOID && epi <= CUM_EPI_GLU_GELU; }
static inline bool epi_valid(int epi) {
    return epi == CUM_EPI_NONE || epi == CUM_EPI_RELU || epi == CUM_EPI_SILU || epi_is_glu(epi);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// entry points implemented per translation unit (called by api.cu)
int wave_normalize_fwd(float* x, float* std_out, int batch, int length, cudaStream_t st);
int conv_in_fwd(const float* x, long long x_stride, int batch, int length, const float* w, const float* bias,
                float* y, int rows_out, int c_pad, int kernel, int stride, const float* in_scale, int group_rows,
                int row_offset, cudaStream_t st, int out_fmt = 0 /* 0 fp32, 1 bf16, 2 fp16 hi/lo planes */, void* y_lo = nullptr,
                long long y_bs = 0, long long y_rs = 0 /* 0, 0 = contiguous (batch, rows_out, c_pad) */);
int stream_std_fwd(const float* x, long long x_stride, int batch, int frames, int frame_len, int hop,
                   int frames_before, float* running, float* scale_out, cudaStream_t st, int* frames_counter = nullptr);
int convt_out_fwd(const float* g, int batch, int rows_in, int c_pad, const float* w, float bias,
                  const float* scale, int scale_group, float* out, long long out_stride, int first, int length,
                  int kernel, int stride, cudaStream_t st, bool in_bf16 = false, long long g_bs = 0, long long g_rs = 0);
int gemm_simt_fwd(const cum_gemm_desc& d, cudaStream_t st);
bool gemm_skinny_ok(const cum_gemm_desc& d);        // a few output rows in total (single-stream streaming): CUDA-core path
int gemm_skinny_fwd(const cum_gemm_desc& d, cudaStream_t st);
int gemm_tc_fwd(const cum_gemm_desc& d, cudaStream_t st);
int enc0_block_fwd(const cum_enc0_block_desc& d, cudaStream_t st);
int dec_last_block_fwd(const cum_dec_last_block_desc& d, cudaStream_t st);
int split_tf32(const float* w, float* hi, float* lo, long long n, cudaStream_t st);
int split_bf16(const float* w, void* hi, void* lo, long long n, cudaStream_t st);
int split_f16(const float* w, void* hi, void* lo, long long n, float scale, cudaStream_t st);
int ln_residual_fwd(const float* h, const float* residual_in, float* residual_out, float* normed,
                    const float* gamma, const float* beta, float eps, long long rows, int c, int c_pad,
                    cudaStream_t st);
int dwconv_silu_fwd(const float* x, long long x_bs, long long x_rs, const float* w, const float* bias, float* y,
                    const float* conv_state, float* conv_state_out, int batch, int len, int d_pad, int width,
                    cudaStream_t st, long long y_bs = 0, long long y_rs = 0);
struct StreamShiftEntry { float* base; long long row_stride; long long src_off; long long count; int rows; int pad; };
int stream_shift_fwd(const StreamShiftEntry* entries, int n_entries, cudaStream_t st);
int selective_scan_fwd(const cum_scan_desc& d, cudaStream_t st);
long long selective_scan_workspace_bytes(const cum_scan_desc& d);

int glu_fwd(const float* z, const float* addend, float* out, long long rows, int h_pad, cudaStream_t st);
int rowblock_bwd(int mode, const float* z, const float* dout, float* dz, float* dbias, long long rows, int cols, cudaStream_t st,
                 float* scale4 = nullptr);
int grad_scale_finalize(float* scale4, cudaStream_t st);
int add_fwd(const float* a, const float* b, float* out, long long count, cudaStream_t st);
int wgrad_fwd(const cum_wgrad_desc& d, cudaStream_t st);
int wgrad_tc_fwd(const cum_wgrad_desc& d, cudaStream_t st);
long long wgrad_tc_workspace_bytes(const cum_wgrad_desc& d);
int grad_scale_fwd(const float* x, long long bs, long long rs, int batch, int rows, int cols, float* scale4, cudaStream_t st);
int ln_bwd(const float* x, const float* dy, const float* dres_in, const float* gamma, float* dx, float* dgamma,
           float* dbeta, float eps, long long rows, int c, int c_pad, cudaStream_t st);
int dwconv_silu_bwd(const float* x, long long x_bs, long long x_rs, const float* w, const float* bias, const float* dy,
                    float* dx, long long dx_bs, long long dx_rs, float* dw, float* db, int batch, int len, int d_pad,
                    int width, cudaStream_t st);
int conv_in_bwd(const float* x, long long x_stride, int batch, int length, const float* y, const float* dy, float* dw,
                float* db, int rows_out, int c_pad, int kernel, int stride, cudaStream_t st);
int convt_out_bwd(const float* g, int batch, int rows_in, int c_pad, const float* w, const float* scale,
                  const float* dout, long long dout_stride, int length, float* dg, float* dw, float* dbias, int kernel,
                  int stride, cudaStream_t st);
int selective_scan_bwd(const cum_scan_bwd_desc& d, cudaStream_t st);
int stft_frames_fwd(const float* x, const float* y, long long sig_stride, int length, int batch, int n_frames, int hop, int win,
                    int n_fft, float* frames, cudaStream_t st);
int stft_loss_reduce_fwd(const float* sx, const float* sy, long long rows, int bins, int ld, double* sums, cudaStream_t st);
int stft_loss_bwd(const float* sx, const float* sy, long long rows, int bins, int ld, const float* coef, float* dsx, cudaStream_t st);
int stft_overlap_add(const float* dframes, int length, int batch, int n_frames, int hop, int win, int n_fft, float* dx,
                     long long dx_stride, cudaStream_t st);
int channel_importance_fwd(const float* w, const float* g, int rows, int cols, long long ldw, long long ldg, float* out_rows,
                           float* out_cols, cudaStream_t st);

constexpr int CI_MAXK = 8;   // conv_in / conv_in_bwd
constexpr int CT_MAXK = 8;   // convt_out / convt_out_bwd

int  sm_count();                                                          // of the calling thread's current device
int  ensure_dyn_smem(const void* kernel, int bytes, const char* what);    // per (kernel, device), thread-safe
void forget_func_attrs();
void tensor_map_cache_clear();                                             // gemm_tc.cu

}  // namespace cum
