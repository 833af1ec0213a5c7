// extern "C" surface of libcleanumamba_sm100.so: argument validation, error reporting, dispatch.
#include "common.cuh"

#include <string.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <set>
#include <utility>

namespace cum {

static thread_local char g_err[512] = "";
// per-device state: one process may drive several GPUs (teacher / student on two devices, DataParallel), from several threads
constexpr int MAX_DEVICES = 64;
static std::atomic<int> g_sm_count[MAX_DEVICES];
static std::mutex g_attr_mu;
static std::set<std::pair<const void*, int>> g_attr_done;      // (kernel, device) pairs whose dynamic-smem limit is raised

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return CUM_ECUDA;
}
static int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev < 0 || dev >= MAX_DEVICES) ? 0 : dev;
}
// SM count of the CURRENT device of the calling thread
int sm_count() {
    const int dev = current_device();
    int v = g_sm_count[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        g_sm_count[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: raise it once per (kernel, device)
int ensure_dyn_smem(const void* kernel, int bytes, const char* what) {
    const std::pair<const void*, int> key(kernel, current_device());
    std::lock_guard<std::mutex> lk(g_attr_mu);
    if (g_attr_done.count(key)) return CUM_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return cuda_fail(e, what);
    g_attr_done.insert(key);
    return CUM_OK;
}
bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CUM_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}
void forget_func_attrs() {
    std::lock_guard<std::mutex> lk(g_attr_mu);
    g_attr_done.clear();
}

static int validate_gemm(const cum_gemm_desc& d) {
    CUM_REQUIRE(d.a && d.w && d.c, "gemm: null a/w/c pointer");
    CUM_REQUIRE(d.batch > 0 && d.m > 0 && d.n > 0 && d.k > 0, "gemm: empty problem (batch=%d m=%d n=%d k=%d)", d.batch, d.m, d.n, d.k);
    CUM_REQUIRE(d.taps == 1 || d.taps == 2, "gemm: taps=%d (must be 1 or 2)", d.taps);
    CUM_REQUIRE(d.k % 4 == 0 && d.ldw % 4 == 0 && d.ldw >= d.k, "gemm: k=%d / ldw=%d must be multiples of 4 with ldw >= k", d.k, d.ldw);
    CUM_REQUIRE(d.n % 8 == 0, "gemm: n=%d must be a multiple of 8", d.n);
    CUM_REQUIRE(d.a_row_stride % 4 == 0 && d.a_batch_stride % 4 == 0 && d.a_row_stride >= 0, "gemm: a strides must be multiples of 4 elements");
    CUM_REQUIRE(d.c_row_stride % 4 == 0 && d.c_batch_stride % 4 == 0, "gemm: c strides must be multiples of 4 elements");
    CUM_REQUIRE(aligned16(d.a) && aligned16(d.w) && aligned16(d.c) && (!d.bias || aligned16(d.bias)), "gemm: a/w/c/bias must be 16-byte aligned");
    CUM_REQUIRE(epi_valid(d.epilogue), "gemm: unknown epilogue %d", d.epilogue);
    if (d.addend) {
        CUM_REQUIRE(aligned16(d.addend) && d.add_row_stride % 4 == 0 && d.add_batch_stride % 4 == 0, "gemm: addend must be 16-byte aligned with strides multiple of 4");
    }
    CUM_REQUIRE(d.a_rows > 0, "gemm: a_rows=%d", d.a_rows);
    return CUM_OK;
}

}  // namespace cum

using namespace cum;

extern "C" {

int cum_abi_version(void) { return CUM_ABI_VERSION; }

const char* cum_last_error(void) { return g_err; }

// Checks that `device` can run this library.  Does NOT change the caller's current device: every entry point works on the
// device that is current in the calling thread (the Python host wraps its calls in a device guard).
int cum_init(int device) {
    CUM_REQUIRE(device >= 0 && device < MAX_DEVICES, "cum_init: device %d out of range", device);
    int major = 0, minor = 0, sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (major != 10) {
        set_error("cum_init: device %d is sm_%d%d; this library contains sm_100a code only (no fallback)", device, major, minor);
        return CUM_ENOTSUP;
    }
    g_sm_count[device].store(sms, std::memory_order_relaxed);
    return CUM_OK;
}

// Drops every process-wide cache of the library (tensor-map cache, per-device kernel attributes).  Stream-ordered work already
// submitted is not affected; the library can be used again afterwards (caches refill lazily).
int cum_shutdown(void) {
    tensor_map_cache_clear();
    forget_func_attrs();
    return CUM_OK;
}

int cum_wave_normalize_fwd(float* x, float* std_out, int batch, int length, cum_stream_t stream) {
    return wave_normalize_fwd(x, std_out, batch, length, (cudaStream_t)stream);
}

int cum_conv_in_fwd(const float* x, long long x_stride, int batch, int length, const float* w, const float* bias,
                    float* y, int rows_out, int c_pad, int kernel, int stride, const float* in_scale, int group_rows,
                    int row_offset, cum_stream_t stream) {
    return conv_in_fwd(x, x_stride, batch, length, w, bias, y, rows_out, c_pad, kernel, stride, in_scale, group_rows,
                       row_offset, (cudaStream_t)stream);
}

int cum_conv_in_bf16_fwd(const float* x, long long x_stride, int batch, int length, const float* w, const float* bias,
                         void* y_bf16, int rows_out, int c_pad, int kernel, int stride, cum_stream_t stream) {
    return conv_in_fwd(x, x_stride, batch, length, w, bias, reinterpret_cast<float*>(y_bf16), rows_out, c_pad, kernel, stride,
                       nullptr, 0, 0, (cudaStream_t)stream, 1);
}

int cum_conv_in_hl16_fwd(const float* x, long long x_stride, int batch, int length, const float* w, const float* bias,
                         void* y_hi, void* y_lo, int rows_out, int c_pad, int kernel, int stride, cum_stream_t stream) {
    return conv_in_fwd(x, x_stride, batch, length, w, bias, reinterpret_cast<float*>(y_hi), rows_out, c_pad, kernel, stride,
                       nullptr, 0, 0, (cudaStream_t)stream, 2, y_lo);
}

int cum_convt_out_bf16_fwd(const void* g_bf16, int batch, int rows_in, int c_pad, const float* w, float bias,
                           const float* scale, int scale_group, float* out, long long out_stride, int first, int length,
                           int kernel, int stride, cum_stream_t stream) {
    return convt_out_fwd(reinterpret_cast<const float*>(g_bf16), batch, rows_in, c_pad, w, bias, scale, scale_group, out,
                         out_stride, first, length, kernel, stride, (cudaStream_t)stream, true);
}

int cum_stream_std_fwd(const float* x, long long x_stride, int batch, int frames, int frame_len, int hop,
                       int frames_before, float* running, float* scale_out, cum_stream_t stream) {
    return stream_std_fwd(x, x_stride, batch, frames, frame_len, hop, frames_before, running, scale_out,
                          (cudaStream_t)stream);
}

int cum_stream_std_counter_fwd(const float* x, long long x_stride, int batch, int frames, int frame_len, int hop,
                               int* frames_counter, float* running, float* scale_out, cum_stream_t stream) {
    CUM_REQUIRE(frames_counter, "stream_std: null frames_counter");
    return stream_std_fwd(x, x_stride, batch, frames, frame_len, hop, 0, running, scale_out, (cudaStream_t)stream, frames_counter);
}

int cum_convt_out_fwd(const float* g, int batch, int rows_in, int c_pad, const float* w, float bias,
                      const float* scale, int scale_group, float* out, long long out_stride, int first, int length,
                      int kernel, int stride, cum_stream_t stream) {
    return convt_out_fwd(g, batch, rows_in, c_pad, w, bias, scale, scale_group, out, out_stride, first, length, kernel,
                         stride, (cudaStream_t)stream);
}

// ---- ABI v3: strided variants for the time-major streaming session + its FIFO maintenance
int cum_conv_in_strided_fwd(const float* x, long long x_stride, int batch, int length, const float* w, const float* bias,
                            float* y, long long y_batch_stride, long long y_row_stride, int rows_out, int c_pad, int kernel, int stride,
                            const float* in_scale, int group_rows, int row_offset, cum_stream_t stream) {
    CUM_REQUIRE(y_batch_stride > 0 && y_row_stride > 0, "conv_in_strided: strides must be positive");
    return conv_in_fwd(x, x_stride, batch, length, w, bias, y, rows_out, c_pad, kernel, stride, in_scale, group_rows,
                       row_offset, (cudaStream_t)stream, 0, nullptr, y_batch_stride, y_row_stride);
}

int cum_convt_out_strided_fwd(const float* g, long long g_batch_stride, long long g_row_stride, int batch, int rows_in, int c_pad,
                              const float* w, float bias, const float* scale, int scale_group, float* out, long long out_stride,
                              int first, int length, int kernel, int stride, cum_stream_t stream) {
    CUM_REQUIRE(g_batch_stride > 0 && g_row_stride > 0, "convt_out_strided: strides must be positive");
    return convt_out_fwd(g, batch, rows_in, c_pad, w, bias, scale, scale_group, out, out_stride, first, length, kernel,
                         stride, (cudaStream_t)stream, false, g_batch_stride, g_row_stride);
}

int cum_dwconv_silu_strided_fwd(const float* x, long long x_batch_stride, long long x_row_stride, const float* w,
                                const float* bias, float* y, long long y_batch_stride, long long y_row_stride,
                                const float* conv_state, float* conv_state_out, int batch, int len, int d_pad, int width,
                                cum_stream_t stream) {
    CUM_REQUIRE(y_batch_stride > 0 && y_row_stride > 0, "dwconv_silu_strided: strides must be positive");
    return dwconv_silu_fwd(x, x_batch_stride, x_row_stride, w, bias, y, conv_state, conv_state_out, batch, len, d_pad,
                           width, (cudaStream_t)stream, y_batch_stride, y_row_stride);
}

int cum_stream_shift_fwd(const cum_shift_entry* entries, int n_entries, cum_stream_t stream) {
    static_assert(sizeof(cum_shift_entry) == sizeof(StreamShiftEntry), "cum_shift_entry mirrors StreamShiftEntry");
    return stream_shift_fwd(reinterpret_cast<const StreamShiftEntry*>(entries), n_entries, (cudaStream_t)stream);
}

int cum_gemm_bias_act_fwd(const cum_gemm_desc* desc, cum_stream_t stream) {
    if (!desc) { set_error("gemm: null descriptor"); return CUM_EINVAL; }
    int rc = validate_gemm(*desc);
    if (rc != CUM_OK) return rc;
    if (gemm_skinny_ok(*desc)) return gemm_skinny_fwd(*desc, (cudaStream_t)stream);
    switch (desc->math) {
        case CUM_MATH_FP32: return gemm_simt_fwd(*desc, (cudaStream_t)stream);
        case CUM_MATH_TF32X3:
        case CUM_MATH_BF16X3:
        case CUM_MATH_F16X3:
        case CUM_MATH_BF16:
        case CUM_MATH_TF32: return gemm_tc_fwd(*desc, (cudaStream_t)stream);
        default: set_error("gemm: unknown math mode %d", desc->math); return CUM_EINVAL;
    }
}

int cum_enc0_block_fwd(const cum_enc0_block_desc* desc, cum_stream_t stream) {
    if (!desc) { set_error("enc0_block: null descriptor"); return CUM_EINVAL; }
    return enc0_block_fwd(*desc, (cudaStream_t)stream);
}

int cum_dec_last_block_fwd(const cum_dec_last_block_desc* desc, cum_stream_t stream) {
    if (!desc) { set_error("dec_last_block: null descriptor"); return CUM_EINVAL; }
    return dec_last_block_fwd(*desc, (cudaStream_t)stream);
}

int cum_split_tf32(const float* w, float* hi, float* lo, long long count, cum_stream_t stream) {
    return split_tf32(w, hi, lo, count, (cudaStream_t)stream);
}

int cum_split_bf16(const float* w, void* hi, void* lo, long long count, cum_stream_t stream) {
    return split_bf16(w, hi, lo, count, (cudaStream_t)stream);
}

int cum_split_f16(const float* w, void* hi, void* lo, long long count, float scale, cum_stream_t stream) {
    return split_f16(w, hi, lo, count, scale, (cudaStream_t)stream);
}

int cum_ln_residual_fwd(const float* h, const float* residual_in, float* residual_out, float* normed,
                        const float* gamma, const float* beta, float eps, long long rows, int c, int c_pad,
                        cum_stream_t stream) {
    return ln_residual_fwd(h, residual_in, residual_out, normed, gamma, beta, eps, rows, c, c_pad, (cudaStream_t)stream);
}

int cum_dwconv_silu_fwd(const float* x, long long x_batch_stride, long long x_row_stride, const float* w,
                        const float* bias, float* y, const float* conv_state, float* conv_state_out, int batch,
                        int len, int d_pad, int width, cum_stream_t stream) {
    return dwconv_silu_fwd(x, x_batch_stride, x_row_stride, w, bias, y, conv_state, conv_state_out, batch, len, d_pad,
                           width, (cudaStream_t)stream);
}

int cum_selective_scan_fwd(const cum_scan_desc* desc, cum_stream_t stream) {
    if (!desc) { set_error("selective_scan: null descriptor"); return CUM_EINVAL; }
    return selective_scan_fwd(*desc, (cudaStream_t)stream);
}

long long cum_selective_scan_workspace_bytes(const cum_scan_desc* desc) {
    return desc ? selective_scan_workspace_bytes(*desc) : 0;
}

int cum_glu_fwd(const float* z, const float* addend, float* out, long long rows, int h_pad, cum_stream_t stream) {
    return glu_fwd(z, addend, out, rows, h_pad, (cudaStream_t)stream);
}
int cum_glu_bwd(const float* z, const float* dout, float* dz, float* dbias, long long rows, int h_pad, float* dz_scale4,
                cum_stream_t stream) {
    return rowblock_bwd(0, z, dout, dz, dbias, rows, 2 * h_pad, (cudaStream_t)stream, dz_scale4);
}
int cum_relu_bwd(const float* y, const float* dy, float* dz, float* dbias, long long rows, int cols, float* dz_scale4,
                 cum_stream_t stream) {
    return rowblock_bwd(1, y, dy, dz, dbias, rows, cols, (cudaStream_t)stream, dz_scale4);
}
int cum_colsum(const float* d, float* dbias, long long rows, int cols, float* d_scale4, cum_stream_t stream) {
    return rowblock_bwd(2, nullptr, d, nullptr, dbias, rows, cols, (cudaStream_t)stream, d_scale4);
}
int cum_add_fwd(const float* a, const float* b, float* out, long long count, cum_stream_t stream) {
    return add_fwd(a, b, out, count, (cudaStream_t)stream);
}
int cum_gemm_wgrad(const cum_wgrad_desc* desc, cum_stream_t stream) {
    if (!desc) { set_error("wgrad: null descriptor"); return CUM_EINVAL; }
    if (desc->math == CUM_MATH_FP32) return wgrad_fwd(*desc, (cudaStream_t)stream);
    return wgrad_tc_fwd(*desc, (cudaStream_t)stream);
}
int cum_grad_scale_fwd(const float* x, long long batch_stride, long long row_stride, int batch, int rows, int cols, float* scale4,
                       cum_stream_t stream) {
    return grad_scale_fwd(x, batch_stride, row_stride, batch, rows, cols, scale4, (cudaStream_t)stream);
}
long long cum_gemm_wgrad_workspace_bytes(const cum_wgrad_desc* desc) {
    if (!desc || desc->math == CUM_MATH_FP32) return 0;
    return wgrad_tc_workspace_bytes(*desc);
}
int cum_ln_residual_bwd(const float* x, const float* dy, const float* dres_in, const float* gamma, float* dx,
                        float* dgamma, float* dbeta, float eps, long long rows, int c, int c_pad, cum_stream_t stream) {
    return ln_bwd(x, dy, dres_in, gamma, dx, dgamma, dbeta, eps, rows, c, c_pad, (cudaStream_t)stream);
}
int cum_dwconv_silu_bwd(const float* x, long long x_batch_stride, long long x_row_stride, const float* w,
                        const float* bias, const float* dy, float* dx, long long dx_batch_stride,
                        long long dx_row_stride, float* dw, float* db, int batch, int len, int d_pad, int width,
                        cum_stream_t stream) {
    return dwconv_silu_bwd(x, x_batch_stride, x_row_stride, w, bias, dy, dx, dx_batch_stride, dx_row_stride, dw, db,
                           batch, len, d_pad, width, (cudaStream_t)stream);
}
int cum_conv_in_bwd(const float* x, long long x_stride, int batch, int length, const float* y, const float* dy,
                    float* dw, float* db, int rows_out, int c_pad, int kernel, int stride, cum_stream_t stream) {
    return conv_in_bwd(x, x_stride, batch, length, y, dy, dw, db, rows_out, c_pad, kernel, stride, (cudaStream_t)stream);
}
int cum_convt_out_bwd(const float* g, int batch, int rows_in, int c_pad, const float* w, const float* scale,
                      const float* dout, long long dout_stride, int length, float* dg, float* dw, float* dbias,
                      int kernel, int stride, cum_stream_t stream) {
    return convt_out_bwd(g, batch, rows_in, c_pad, w, scale, dout, dout_stride, length, dg, dw, dbias, kernel, stride,
                         (cudaStream_t)stream);
}
int cum_stft_frames_fwd(const float* x, const float* y, long long sig_stride, int length, int batch, int n_frames, int hop, int win,
                        int n_fft, float* frames, cum_stream_t stream) {
    return stft_frames_fwd(x, y, sig_stride, length, batch, n_frames, hop, win, n_fft, frames, (cudaStream_t)stream);
}
int cum_stft_loss_reduce_fwd(const float* sx, const float* sy, long long rows, int bins, int ld, double* sums, cum_stream_t stream) {
    return stft_loss_reduce_fwd(sx, sy, rows, bins, ld, sums, (cudaStream_t)stream);
}
int cum_stft_loss_bwd(const float* sx, const float* sy, long long rows, int bins, int ld, const float* coef, float* dsx,
                      cum_stream_t stream) {
    return stft_loss_bwd(sx, sy, rows, bins, ld, coef, dsx, (cudaStream_t)stream);
}
int cum_stft_overlap_add(const float* dframes, int length, int batch, int n_frames, int hop, int win, int n_fft, float* dx,
                         long long dx_stride, cum_stream_t stream) {
    return stft_overlap_add(dframes, length, batch, n_frames, hop, win, n_fft, dx, dx_stride, (cudaStream_t)stream);
}
int cum_channel_importance_fwd(const float* w, const float* g, int rows, int cols, long long ldw, long long ldg, float* out_rows,
                               float* out_cols, cum_stream_t stream) {
    return channel_importance_fwd(w, g, rows, cols, ldw, ldg, out_rows, out_cols, (cudaStream_t)stream);
}
int cum_selective_scan_bwd(const cum_scan_bwd_desc* desc, cum_stream_t stream) {
    if (!desc) { set_error("selective_scan_bwd: null descriptor"); return CUM_EINVAL; }
    return selective_scan_bwd(*desc, (cudaStream_t)stream);
}

}  // extern "C"
