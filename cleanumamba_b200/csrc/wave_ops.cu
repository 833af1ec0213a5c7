// Waveform-end kernels: input normalisation, first encoder conv (Cin = 1), last decoder transposed conv (Cout = 1).
// All three are HBM-bound streaming kernels (a few FLOP per byte); see DESIGN.md "kernels".
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace cum {

// ---------------------------------------------------------------------------------------------------------
// std (unbiased) + 1e-3, in-place divide.  Reference: CleanUMamba.py:260-262.  Two-pass mean/variance in fp64 (torch's CPU std
// accumulates in double), then the divide.
// ---------------------------------------------------------------------------------------------------------
__device__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    if (wid == 0) {
        t = warp_sum(t);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}

// A clip is shared by a CLUSTER of WN_CL CTAs (one CTA per clip walked 640 KB three times with 1024 threads: 96 us at batch 1 -- 13 %
// of the whole forward of a pruned checkpoint -- and 64 CTAs on 148 SMs at batch 64).  Each CTA reduces its slice, the CTA totals meet
// through distributed shared memory (every CTA adds the WN_CL partials in the same order, so all of them hold the same fp64 value).
constexpr int WN_CL = 8, WN_THREADS = 512;

__device__ __forceinline__ double cluster_sum(double v, double* red, double* part) {
    const double t = block_sum(v, red);
    if (threadIdx.x == 0) *part = t;
    cluster_sync_all();                                  // release / acquire: every CTA's partial is visible cluster-wide
    double tot = 0.0;
#pragma unroll
    for (int r = 0; r < WN_CL; ++r) {
        double pv;
        asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(pv) : "r"(mapa_rank(smem_u32(part), (uint32_t)r)));
        tot += pv;
    }
    return tot;
}

__global__ void __cluster_dims__(WN_CL, 1, 1) __launch_bounds__(WN_THREADS)
wave_normalize_kernel(float* __restrict__ x, float* __restrict__ std_out, int length) {
    pdl_trigger();
    pdl_wait();          // PDL: nothing of the previous kernel is touched before this point

    __shared__ double red[32];
    __shared__ double part[2];
    const int clip = blockIdx.x / WN_CL;
    const int rank = (int)cluster_ctarank();
    float* row = x + (long long)clip * length;
    const bool vec = (length % 4 == 0) && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
    const int t = rank * WN_THREADS + threadIdx.x, nt = WN_CL * WN_THREADS;      // this thread among the clip's threads
    const int n4 = vec ? length >> 2 : 0;
    float4* row4 = reinterpret_cast<float4*>(row);
    double s = 0.0;
    if (vec) for (int i = t; i < n4; i += nt) { const float4 v = row4[i]; s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w); }
    else for (int i = t; i < length; i += nt) s += (double)row[i];
    const double mean = cluster_sum(s, red, &part[0]) / (double)length;
    double q = 0.0;
    if (vec) for (int i = t; i < n4; i += nt) {
        const float4 v = row4[i];
        const double d0 = (double)v.x - mean, d1 = (double)v.y - mean, d2 = (double)v.z - mean, d3 = (double)v.w - mean;
        q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    else for (int i = t; i < length; i += nt) { const double d = (double)row[i] - mean; q += d * d; }
    const double var = cluster_sum(q, red, &part[1]) / (double)(length > 1 ? length - 1 : 1);
    const float sd = (float)sqrt(var) + 1e-3f;
    if (t == 0) std_out[clip] = sd;
    if (vec) for (int i = t; i < n4; i += nt) { float4 v = row4[i]; v.x = v.x / sd; v.y = v.y / sd; v.z = v.z / sd; v.w = v.w / sd; row4[i] = v; }
    else for (int i = t; i < length; i += nt) row[i] = row[i] / sd;
    cluster_sync_all();                                  // no CTA exits while a peer may still read its partials
}

int wave_normalize_fwd(float* x, float* std_out, int batch, int length, cudaStream_t st) {
    CUM_REQUIRE(x && std_out && batch > 0 && length > 0, "wave_normalize: bad arguments");
    cudaError_t e = launch_kernel(wave_normalize_kernel, dim3((unsigned)batch * WN_CL), dim3(WN_THREADS), 0, st, x, std_out, length);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(wave_normalize_kernel)");
    return CUM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// stream_std: per-frame std of the streaming path.  Reference: CleanUMamba.py:399-401
//   s_f = std(frame_f, unbiased) + 1e-3 ;  input_std <- s_f / f + (1 - 1/f) * input_std   (running mean over frames)
// One CTA per stream walks the `frames` frames of this call in order; frame j covers x[b, j*hop : j*hop + frame_len].
// scale_out[b, j] = the running value after frame j; running[b] is updated in place.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stream_std_kernel(const float* __restrict__ x, long long x_stride, int frames,
                                                          int frame_len, int hop, int frames_before, const int* __restrict__ frames_counter,
                                                          float* __restrict__ running, float* __restrict__ scale_out) {
    pdl_trigger();
    pdl_wait();          // PDL: nothing of the previous kernel is touched before this point

    __shared__ double red[32];
    const int b = blockIdx.x;
    if (frames_counter) frames_before = *frames_counter;       // device-side frame count (CUDA-graph replays: no host argument changes)
    const float* xb = x + (long long)b * x_stride;
    float run = running[b];
    for (int j = 0; j < frames; ++j) {
        const float* fr = xb + (long long)j * hop;
        double s = 0.0;
        for (int i = threadIdx.x; i < frame_len; i += blockDim.x) s += (double)fr[i];
        const double mean = block_sum(s, red) / (double)frame_len;
        double q = 0.0;
        for (int i = threadIdx.x; i < frame_len; i += blockDim.x) {
            const double d = (double)fr[i] - mean;
            q += d * d;
        }
        const double var = block_sum(q, red) / (double)(frame_len > 1 ? frame_len - 1 : 1);
        const float sf = (float)sqrt(var) + 1e-3f;
        const float f = (float)(frames_before + j + 1);
        run = sf / f + (1.0f - 1.0f / f) * run;
        if (threadIdx.x == 0) scale_out[(long long)b * frames + j] = run;
    }
    if (threadIdx.x == 0) running[b] = run;
}

__global__ void add_int_kernel(int* counter, int v) {
    pdl_wait();
    *counter += v;
}

int stream_std_fwd(const float* x, long long x_stride, int batch, int frames, int frame_len, int hop,
                   int frames_before, float* running, float* scale_out, cudaStream_t st, int* frames_counter) {
    CUM_REQUIRE(x && running && scale_out, "stream_std: null pointer");
    CUM_REQUIRE(batch > 0 && frames > 0 && frame_len > 0 && hop > 0 && frames_before >= 0, "stream_std: bad shape");
    cudaError_t e = launch_kernel(stream_std_kernel, dim3(batch), dim3(256), 0, st, x, x_stride, frames, frame_len, hop, frames_before,
                                  (const int*)frames_counter, running, scale_out);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(stream_std_kernel)");
    if (frames_counter) {       // every block has read the counter: stream order
        e = launch_kernel(add_int_kernel, dim3(1), dim3(1), 0, st, frames_counter, frames);
        if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(add_int_kernel)");
    }
    return CUM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// conv_in: y[b,t,c] = relu(bias[c] + sum_k w[k,c] * x[b, S t + k]),  x read as 0 beyond `length` (F.pad).
// CTA = 64 output rows; the 64*S+K input samples are staged in smem, every thread then produces float4s of
// channels for one row so the stores are fully coalesced (the kernel is store-bound: 4*Cp bytes per row).
// ---------------------------------------------------------------------------------------------------------
constexpr int CI_ROWS = 64;

// fp32 -> fp16 hi / lo ("hl16", see cum_gemm_desc): hi = fp16(x) saturated, lo = fp16(x - hi)
__device__ __forceinline__ uint32_t cvt_h2_sat(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
// store 4 consecutive channels at element offset `o`: OUTF 0 = fp32 (y), 1 = bf16 (y), 2 = hl16 (y = hi plane, y_lo = lo plane)
template <int OUTF>
__device__ __forceinline__ void store4(float* __restrict__ y, void* __restrict__ y_lo, long long o, float4 acc) {
    if (OUTF == 1) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + o) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    } else if (OUTF == 2) {
        const uint32_t h0 = cvt_h2_sat(acc.x, acc.y), h1 = cvt_h2_sat(acc.z, acc.w);
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h0)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
        const uint32_t l0 = cvt_h2_sat(acc.x - f0.x, acc.y - f0.y), l1 = cvt_h2_sat(acc.z - f1.x, acc.w - f1.y);
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(y) + o) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(y_lo) + o) = make_uint2(l0, l1);
    } else {
        *reinterpret_cast<float4*>(y + o) = acc;
    }
}

template <int OUTF>
__global__ void __launch_bounds__(256) conv_in_kernel(const float* __restrict__ x, long long x_stride, int length,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       float* __restrict__ y, void* __restrict__ y_lo, int rows_out, int c_pad, int kernel,
                                                       int stride, const float* __restrict__ in_scale, int scale_groups,
                                                       int group_rows, int row_offset, long long y_bs, long long y_rs) {
    pdl_trigger();
    pdl_wait();          // PDL: nothing of the previous kernel is touched before this point

    extern __shared__ float xs[];
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * CI_ROWS;
    const int nin = CI_ROWS * stride + kernel;
    const float* xb = x + (long long)b * x_stride;
    for (int i = threadIdx.x; i < nin; i += blockDim.x) {
        const long long g = (long long)t0 * stride + i;
        xs[i] = (g < length) ? xb[g] : 0.0f;
    }
    __syncthreads();
    const int c4n = c_pad >> 2;
    const int rows = min(CI_ROWS, rows_out - t0);
    for (int idx = threadIdx.x; idx < rows * c4n; idx += blockDim.x) {
        const int t = idx / c4n, c4 = idx - t * c4n;
        float4 acc = __ldg(reinterpret_cast<const float4*>(bias) + c4);
        // streaming: the samples feeding output row t are divided by the running std of the hop that row belongs to
        const float sdiv = in_scale ? __ldg(in_scale + (long long)b * scale_groups + max(0, t0 + t + row_offset) / group_rows) : 1.0f;
#pragma unroll 4
        for (int k = 0; k < kernel; ++k) {
            const float xv = in_scale ? xs[t * stride + k] / sdiv : xs[t * stride + k];
            const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (long long)k * c_pad) + c4);
            acc.x = fmaf(wv.x, xv, acc.x); acc.y = fmaf(wv.y, xv, acc.y);
            acc.z = fmaf(wv.z, xv, acc.z); acc.w = fmaf(wv.w, xv, acc.w);
        }
        acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
        store4<OUTF>(y, y_lo, (long long)b * y_bs + (long long)(t0 + t) * y_rs + 4 * c4, acc);
    }
}

// Fast path for the shipped geometry (64 channels, K = 4, S = 2, no per-hop scale): 256 rows per CTA, each thread owns one
// float4 of channels for 16 rows, weights and bias live in registers, every pass stores 16 rows x 256 B = 4 KB contiguous.
constexpr int CIF_ROWS = 256;
template <int OUTF>
__global__ void __launch_bounds__(256) conv_in_c64_kernel(const float* __restrict__ x, long long x_stride, int length,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           float* __restrict__ y, void* __restrict__ y_lo, int rows_out) {
    pdl_trigger();
    pdl_wait();          // PDL: nothing of the previous kernel is touched before this point

    __shared__ __align__(16) float xs[CIF_ROWS * 2 + 4];
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * CIF_ROWS;
    const float* xb = x + (long long)b * x_stride;
    for (int i = threadIdx.x; i < CIF_ROWS * 2 + 4; i += 256) {
        const long long g = 2ll * t0 + i;
        xs[i] = (g < length) ? __ldg(xb + g) : 0.0f;
    }
    const int c4 = threadIdx.x & 15, tr = threadIdx.x >> 4;
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + c4);
    float4 wv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wv[k] = __ldg(reinterpret_cast<const float4*>(w + k * 64) + c4);
    __syncthreads();
    const int rows = min(CIF_ROWS, rows_out - t0);
#pragma unroll 4
    for (int t = tr; t < rows; t += 16) {
        const float2 xa = *reinterpret_cast<const float2*>(xs + 2 * t), xc = *reinterpret_cast<const float2*>(xs + 2 * t + 2);
        const float xv[4] = {xa.x, xa.y, xc.x, xc.y};
        float4 acc = bv;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc.x = fmaf(wv[k].x, xv[k], acc.x); acc.y = fmaf(wv[k].y, xv[k], acc.y);
            acc.z = fmaf(wv[k].z, xv[k], acc.z); acc.w = fmaf(wv[k].w, xv[k], acc.w);
        }
        acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
        store4<OUTF>(y, y_lo, ((long long)b * rows_out + t0 + t) * 64 + 4 * c4, acc);
    }
}

int conv_in_fwd(const float* x, long long x_stride, int batch, int length, const float* w, const float* bias,
                float* y, int rows_out, int c_pad, int kernel, int stride, const float* in_scale, int group_rows,
                int row_offset, cudaStream_t st, int out_fmt, void* y_lo, long long y_bs, long long y_rs) {
    const bool strided = y_bs != 0 || y_rs != 0;       // streaming (time-major): y[b, t, :] at y + b * y_bs + t * y_rs
    if (!strided) { y_bs = (long long)rows_out * c_pad; y_rs = c_pad; }
    CUM_REQUIRE(y_bs % 4 == 0 && y_rs % 4 == 0, "conv_in: output strides must be multiples of 4 elements");
    CUM_REQUIRE(out_fmt != 2 || (y_lo && aligned16(y_lo)), "conv_in: hi/lo output needs a 16-byte aligned y_lo");
    CUM_REQUIRE(x && w && bias && y, "conv_in: null pointer");
    CUM_REQUIRE(batch > 0 && length > 0 && rows_out > 0, "conv_in: empty problem");
    CUM_REQUIRE(c_pad > 0 && c_pad % 4 == 0, "conv_in: c_pad=%d must be a positive multiple of 4", c_pad);
    CUM_REQUIRE(kernel >= 1 && kernel <= CI_MAXK && stride >= 1 && stride <= kernel, "conv_in: kernel=%d stride=%d unsupported", kernel, stride);
    CUM_REQUIRE(aligned16(w) && aligned16(bias) && aligned16(y), "conv_in: w/bias/y must be 16-byte aligned");
    CUM_REQUIRE(!in_scale || group_rows > 0, "conv_in: group_rows must be positive when in_scale is given");
    if (c_pad == 64 && kernel == 4 && stride == 2 && !in_scale && batch <= 65535 && !strided) {
        dim3 gridf((unsigned)cdiv(rows_out, CIF_ROWS), batch);
        if (out_fmt == 2) conv_in_c64_kernel<2><<<gridf, 256, 0, st>>>(x, x_stride, length, w, bias, y, y_lo, rows_out);
        else if (out_fmt == 1) conv_in_c64_kernel<1><<<gridf, 256, 0, st>>>(x, x_stride, length, w, bias, y, y_lo, rows_out);
        else conv_in_c64_kernel<0><<<gridf, 256, 0, st>>>(x, x_stride, length, w, bias, y, y_lo, rows_out);
        CUM_LAUNCH_CHECK("conv_in_c64_kernel");
        return CUM_OK;
    }
    dim3 grid((unsigned)cdiv(rows_out, CI_ROWS), batch);
    const size_t smem = (size_t)(CI_ROWS * stride + kernel) * sizeof(float);
    const int groups = in_scale ? (int)cdiv(max(1, rows_out + row_offset), group_rows) : 0;
#define CI_LAUNCH(F) conv_in_kernel<F><<<grid, 256, smem, st>>>(x, x_stride, length, w, bias, y, y_lo, rows_out, c_pad, kernel, stride, in_scale, groups, group_rows, row_offset, y_bs, y_rs)
    if (out_fmt == 2) CI_LAUNCH(2); else if (out_fmt == 1) CI_LAUNCH(1); else CI_LAUNCH(0);
#undef CI_LAUNCH
    CUM_LAUNCH_CHECK("conv_in_kernel");
    return CUM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// stream_shift: end-of-call FIFO maintenance of the time-major streaming session -- every carried buffer (pending samples, the
// encoder levels' unconsumed columns, the decoder levels' carried GLU column) moves its unconsumed tail to the front, ALL of them
// in one launch (the round-1 session did this with ~30 torch copy / cat / clone kernels per call).  Entry: for every row r < rows,
// base[r * row_stride + i] = base[r * row_stride + src_off + i], i < count.  Source and destination may overlap (src_off < count):
// positions congruent modulo src_off form independent chains, one thread walks one chain front to back.
// ---------------------------------------------------------------------------------------------------------
constexpr int SHIFT_MAX_ENTRIES = 24;
struct StreamShiftTable { StreamShiftEntry e[SHIFT_MAX_ENTRIES]; };

template <typename V>
__device__ __forceinline__ void shift_entry(const StreamShiftEntry& e, long long first, long long step) {
    constexpr int W = sizeof(V) / 4;
    const long long s = e.src_off / W, n = e.count / W;
    const long long par = s < n ? s : n;                 // independent chains per row
    const long long total = par * e.rows;
    for (long long w = first; w < total; w += step) {
        const long long r = w / par, j = w - r * par;
        V* row = reinterpret_cast<V*>(e.base + r * e.row_stride);
        for (long long i = j; i < n; i += s) row[i] = row[i + s];
    }
}

__global__ void __launch_bounds__(256) stream_shift_kernel(const StreamShiftTable tab) {
    pdl_trigger();
    pdl_wait();
    const StreamShiftEntry& e = tab.e[blockIdx.y];
    if (e.count <= 0 || e.src_off <= 0) return;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x, step = (long long)gridDim.x * blockDim.x;
    const bool v4 = ((e.src_off | e.count | e.row_stride) & 3) == 0 && (reinterpret_cast<uintptr_t>(e.base) & 15) == 0;
    if (v4) shift_entry<float4>(e, first, step); else shift_entry<float>(e, first, step);
}

int stream_shift_fwd(const StreamShiftEntry* entries, int n_entries, cudaStream_t st) {
    CUM_REQUIRE(entries && n_entries > 0 && n_entries <= SHIFT_MAX_ENTRIES, "stream_shift: 1..%d entries", SHIFT_MAX_ENTRIES);
    StreamShiftTable tab;
    memset(&tab, 0, sizeof(tab));
    long long most = 0;
    for (int i = 0; i < n_entries; ++i) {
        const StreamShiftEntry& e = entries[i];
        CUM_REQUIRE(e.count == 0 || (e.base && e.rows > 0 && e.src_off >= 0 && e.count > 0), "stream_shift: bad entry %d", i);
        tab.e[i] = e;
        const long long par = (e.src_off < e.count ? e.src_off : e.count) * (long long)e.rows;
        if (par > most) most = par;
    }
    if (most == 0) return CUM_OK;
    long long blocks = cdiv(cdiv(most, 4), 256);
    const long long cap = 8LL * sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cudaError_t e = launch_kernel(stream_shift_kernel, dim3((unsigned)blocks, (unsigned)n_entries), dim3(256), 0, st, tab);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(stream_shift_kernel)");
    CUM_LAUNCH_CHECK("stream_shift_kernel");
    return CUM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// convt_out: out[b,m] = scale[b] * (bias + sum_{j,k: S j + k = m} <g[b,j,:], w[k,:]>)  for m < length.
// CTA = 64 input rows (+ halo).  Phase 1: a group of LPR lanes computes the K tap-dots of one input row with
// float4 loads (coalesced, read-once: the kernel is load-bound at 4*Cp bytes per row).  Phase 2: overlap-add
// of the dots from smem into S*64 output samples.
// ---------------------------------------------------------------------------------------------------------
constexpr int CT_ROWS = 64;

template <int LPR, bool IN16>
__global__ void __launch_bounds__(256) convt_out_kernel(const float* __restrict__ g, int rows_in, int c_pad,
                                                         const float* __restrict__ w, float bias,
                                                         const float* __restrict__ scale, int scale_groups,
                                                         int scale_group, float* __restrict__ out, long long out_stride,
                                                         int first, int length, int kernel, int stride, int halo,
                                                         long long g_bs, long long g_rs) {
    pdl_trigger();
    pdl_wait();          // PDL: nothing of the previous kernel is touched before this point

    extern __shared__ float dots[];  // [(CT_ROWS + halo)][kernel]
    const int b = blockIdx.y;
    const long long mbase = (long long)first + (long long)blockIdx.x * CT_ROWS * stride;  // first output of this CTA
    const int jbase = (int)(mbase / stride);
    const int j0 = jbase - halo;                 // first staged input row (may be negative)
    const int nrows = CT_ROWS + halo + 1;        // +1: mbase need not be a multiple of stride
    const int c4n = c_pad >> 2;
    const int sub = threadIdx.x % LPR, grp = threadIdx.x / LPR, ngrp = blockDim.x / LPR;
    // NOTE: the trip count must be warp-uniform (full-mask shuffles below): iterate on r0, predicate on r
    for (int r0 = 0; r0 < nrows; r0 += ngrp) {
        const int r = r0 + grp;
        const int j = j0 + r;
        float acc[CT_MAXK];
#pragma unroll
        for (int k = 0; k < CT_MAXK; ++k) acc[k] = 0.f;
        if (r < nrows && j >= 0 && j < rows_in) {
            const float4* row = reinterpret_cast<const float4*>(g + (long long)b * g_bs + (long long)j * g_rs);
            const uint2* row16 = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(g) + (long long)b * g_bs + (long long)j * g_rs);
            for (int c4 = sub; c4 < c4n; c4 += LPR) {
                float4 v;
                if (IN16) {
                    const uint2 raw = __ldg(row16 + c4);
                    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
                    const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
                    v = make_float4(a.x, a.y, c.x, c.y);
                } else {
                    v = __ldg(row + c4);
                }
#pragma unroll
                for (int k = 0; k < CT_MAXK; ++k) {
                    if (k < kernel) {
                        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (long long)k * c_pad) + c4);
                        acc[k] = fmaf(v.x, wv.x, fmaf(v.y, wv.y, fmaf(v.z, wv.z, fmaf(v.w, wv.w, acc[k]))));
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < CT_MAXK; ++k) {
            if (k < kernel) {
                float v = acc[k];
#pragma unroll
                for (int o = LPR >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (sub == 0 && r < nrows) dots[r * kernel + k] = v;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CT_ROWS * stride; i += blockDim.x) {
        const long long m = mbase + i;           // index in the full transposed-conv output
        const long long o = m - first;           // index in `out`
        if (o >= length) break;
        float acc = bias;
        for (int k = (int)(m % stride); k < kernel; k += stride) {
            const long long j = (m - k) / stride;
            if (m - k >= 0 && j < rows_in) acc += dots[(int)(j - j0) * kernel + k];
        }
        const float sc = scale ? __ldg(scale + (long long)b * scale_groups + o / scale_group) : 1.0f;
        out[(long long)b * out_stride + o] = acc * sc;
    }
}

// Fast path for the shipped geometry (64 channels, K = 4, S = 2): 128 staged rows per CTA (126 + halo + 1), 16 lanes per row with the
// four tap weights of their channels in registers, FOUR rows in flight per lane group (the generic kernel had one 16-byte load
// outstanding per thread and ran at 1.2 TB/s), and a 5-shuffle transposing reduction of the 4 tap sums instead of 16.
constexpr int CTF_ROWS = 126;
template <bool IN16>
__global__ void __launch_bounds__(256) convt_out_c64_kernel(const float* __restrict__ g, int rows_in, const float* __restrict__ w,
                                                             float bias, const float* __restrict__ scale, int scale_groups,
                                                             int scale_group, float* __restrict__ out, long long out_stride,
                                                             int first, int length, long long g_bs, long long g_rs) {
    pdl_trigger();
    pdl_wait();          // PDL: nothing of the previous kernel is touched before this point

    __shared__ float dots[(CTF_ROWS + 2) * 4];
    const int b = blockIdx.y;
    const long long mbase = (long long)first + (long long)blockIdx.x * CTF_ROWS * 2;
    const int j0 = (int)(mbase / 2) - 1;
    constexpr int nrows = CTF_ROWS + 2;          // 128
    const int sub = threadIdx.x & 15, grp = threadIdx.x >> 4;
    float4 wv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wv[k] = __ldg(reinterpret_cast<const float4*>(w + k * 64) + sub);
#pragma unroll
    for (int r0 = 0; r0 < nrows; r0 += 64) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + r0 + 16 * u + grp;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j >= 0 && j < rows_in) {
                if (IN16) {
                    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(g) + (long long)b * g_bs + (long long)j * g_rs) + sub);
                    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
                    const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
                    v[u] = make_float4(a.x, a.y, c.x, c.y);
                } else {
                    v[u] = __ldg(reinterpret_cast<const float4*>(g + (long long)b * g_bs + (long long)j * g_rs) + sub);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float a[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = fmaf(v[u].x, wv[k].x, fmaf(v[u].y, wv[k].y, fmaf(v[u].z, wv[k].z, v[u].w * wv[k].w)));
            // transposing reduction over the 16 lanes: after the xor-8 and xor-4 steps every lane carries ONE tap
            const bool up8 = sub & 8, up4 = sub & 4;
            const float s0 = __shfl_xor_sync(0xffffffffu, up8 ? a[0] : a[2], 8);
            const float s1 = __shfl_xor_sync(0xffffffffu, up8 ? a[1] : a[3], 8);
            const float k0 = (up8 ? a[2] : a[0]) + s0, k1 = (up8 ? a[3] : a[1]) + s1;
            float t = (up4 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, up4 ? k0 : k1, 4);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            if ((sub & 3) == 0) dots[(r0 + 16 * u + grp) * 4 + (sub >> 2)] = t;      // tap = 2 * bit3 + bit2
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CTF_ROWS * 2; i += 256) {
        const long long m = mbase + i;
        const long long o = m - first;
        if (o >= length) break;
        float acc = bias;
        for (int k = (int)(m & 1); k < 4; k += 2) {
            const long long j = (m - k) / 2;
            if (m - k >= 0 && j < rows_in) acc += dots[(int)(j - j0) * 4 + k];
        }
        const float sc = scale ? __ldg(scale + (long long)b * scale_groups + o / scale_group) : 1.0f;
        out[(long long)b * out_stride + o] = acc * sc;
    }
}

int convt_out_fwd(const float* g, int batch, int rows_in, int c_pad, const float* w, float bias,
                  const float* scale, int scale_group, float* out, long long out_stride, int first, int length,
                  int kernel, int stride, cudaStream_t st, bool in_bf16, long long g_bs, long long g_rs) {
    if (g_bs == 0 && g_rs == 0) { g_bs = (long long)rows_in * c_pad; g_rs = c_pad; }    // streaming (time-major): g[b, j, :] at g + b * g_bs + j * g_rs
    CUM_REQUIRE(g_bs % 4 == 0 && g_rs % 4 == 0, "convt_out: input strides must be multiples of 4 elements");
    CUM_REQUIRE(g && w && out, "convt_out: null pointer");
    CUM_REQUIRE(batch > 0 && batch <= 65535 && rows_in > 0 && length > 0 && first >= 0, "convt_out: empty problem");
    CUM_REQUIRE(c_pad > 0 && c_pad % 4 == 0, "convt_out: c_pad=%d must be a positive multiple of 4", c_pad);
    CUM_REQUIRE(kernel >= 1 && kernel <= CT_MAXK && stride >= 1 && stride <= kernel, "convt_out: kernel=%d stride=%d unsupported", kernel, stride);
    CUM_REQUIRE(aligned16(g) && aligned16(w), "convt_out: g/w must be 16-byte aligned");
    CUM_REQUIRE(!scale || scale_group > 0, "convt_out: scale_group must be positive when scale is given");
    const int halo = (kernel - 1) / stride;
    const long long out_rows = (long long)(rows_in - 1) * stride + kernel;
    CUM_REQUIRE((long long)first + length <= out_rows, "convt_out: first+length=%lld exceeds the transposed-conv output (%lld)",
                (long long)first + length, out_rows);
    const int groups = scale ? (int)cdiv(length, scale_group) : 0;
    if (c_pad == 64 && kernel == 4 && stride == 2) {
        dim3 gridf((unsigned)cdiv(length, (long long)CTF_ROWS * 2), batch);
        if (in_bf16) convt_out_c64_kernel<true><<<gridf, 256, 0, st>>>(g, rows_in, w, bias, scale, groups, scale_group, out, out_stride, first, length, g_bs, g_rs);
        else convt_out_c64_kernel<false><<<gridf, 256, 0, st>>>(g, rows_in, w, bias, scale, groups, scale_group, out, out_stride, first, length, g_bs, g_rs);
        CUM_LAUNCH_CHECK("convt_out_c64_kernel");
        return CUM_OK;
    }
    dim3 grid((unsigned)cdiv(length, (long long)CT_ROWS * stride), batch);
    const size_t smem = (size_t)(CT_ROWS + halo + 1) * kernel * sizeof(float);
#define CT_LAUNCH(L, I) convt_out_kernel<L, I><<<grid, 256, smem, st>>>(g, rows_in, c_pad, w, bias, scale, groups, scale_group, out, out_stride, first, length, kernel, stride, halo, g_bs, g_rs)
    if (c_pad <= 64) { if (in_bf16) CT_LAUNCH(16, true); else CT_LAUNCH(16, false); }
    else             { if (in_bf16) CT_LAUNCH(32, true); else CT_LAUNCH(32, false); }
#undef CT_LAUNCH
    CUM_LAUNCH_CHECK("convt_out_kernel");
    return CUM_OK;
}

}  // namespace cum
