// Multi-resolution STFT loss -- the kernels around the windowed-DFT-as-GEMM:
//   /root/reference/src/util/stft_loss.py:16-35  torch.stft (center=True, reflect padding, hann window) -> magnitude, floored at sqrt(1e-7)
//   :38-82   spectral convergence ||y_mag - x_mag||_F / ||y_mag||_F  and  L1(log y_mag, log x_mag)
// One resolution = one GEMM on the tensor cores: S[(signal, clip, frame), (f, re|im)] = frames . basis^T with the hann window folded
// into the (win_length x 2F) cos / -sin basis (only the win_length non-zero taps of the n_fft window are contracted).  The kernels
// here build the frame matrix (reflect padding resolved by index arithmetic, no padded copy), reduce the two spectrograms to the
// three sums the loss needs, turn them into dL/dS of the predicted signal, and overlap-add dL/dframes back onto the waveform.
#include "common.cuh"

namespace cum {

// frames[(s * batch + b) * n_frames + fr, k] = sig_s[b, reflect(fr * hop + k + first)],  first = (n_fft - win) / 2 - n_fft / 2
__global__ void __launch_bounds__(256) stft_frames_kernel(const float* __restrict__ x, const float* __restrict__ y, long long sig_stride,
                                                           int length, int batch, int n_frames, int hop, int win, int first,
                                                           float* __restrict__ frames) {
    const long long row = blockIdx.x;                       // (signal, clip, frame)
    const int fr = (int)(row % n_frames);
    const long long sb = row / n_frames;
    const int b = (int)(sb % batch);
    const float* src = (sb / batch ? y : x) + (long long)b * sig_stride;
    float* dst = frames + row * win;
    const int base = fr * hop + first;
    for (int k = threadIdx.x; k < win; k += blockDim.x) {
        int i = base + k;
        if (i < 0) i = -i;
        if (i >= length) i = 2 * (length - 1) - i;
        dst[k] = __ldg(src + i);
    }
}

int stft_frames_fwd(const float* x, const float* y, long long sig_stride, int length, int batch, int n_frames, int hop, int win,
                    int n_fft, float* frames, cudaStream_t st) {
    CUM_REQUIRE(x && frames && length > n_fft / 2 && batch > 0 && n_frames > 0 && hop > 0 && win > 0 && win <= n_fft,
                "stft_frames: bad arguments (reflect padding needs length > n_fft / 2)");
    const long long rows = (long long)(y ? 2 : 1) * batch * n_frames;
    CUM_REQUIRE(rows < (1ll << 31), "stft_frames: too many frames");
    stft_frames_kernel<<<(unsigned)rows, 256, 0, st>>>(x, y ? y : x, sig_stride, length, batch, n_frames, hop, win,
                                                        (n_fft - win) / 2 - n_fft / 2, frames);
    CUM_LAUNCH_CHECK("stft_frames_kernel");
    return CUM_OK;
}

__device__ __forceinline__ float block_sum_f(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    if (wid == 0) t = warp_sum(t);
    return t;       // valid in warp 0
}

// sums[0] += sum (ym - xm)^2, sums[1] += sum ym^2, sums[2] += sum |log ym - log xm|   over rows x bins (fp64 accumulators)
__global__ void __launch_bounds__(256) stft_loss_reduce_kernel(const float* __restrict__ sx, const float* __restrict__ sy, long long rows,
                                                                int bins, int ld, double* __restrict__ sums) {
    __shared__ float red[8];
    float a = 0.f, b2 = 0.f, c = 0.f;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const float2* px = reinterpret_cast<const float2*>(sx + r * ld);
        const float2* py = reinterpret_cast<const float2*>(sy + r * ld);
        for (int f = threadIdx.x; f < bins; f += blockDim.x) {
            const float2 vx = __ldg(px + f), vy = __ldg(py + f);
            const float xm = sqrtf(fmaxf(fmaf(vx.x, vx.x, vx.y * vx.y), 1e-7f));
            const float ym = sqrtf(fmaxf(fmaf(vy.x, vy.x, vy.y * vy.y), 1e-7f));
            const float d = ym - xm;
            a = fmaf(d, d, a);
            b2 = fmaf(ym, ym, b2);
            c += fabsf(logf(ym) - logf(xm));
        }
    }
    a = block_sum_f(a, red); b2 = block_sum_f(b2, red); c = block_sum_f(c, red);
    if (threadIdx.x == 0) { atomicAdd(sums, (double)a); atomicAdd(sums + 1, (double)b2); atomicAdd(sums + 2, (double)c); }
}

int stft_loss_reduce_fwd(const float* sx, const float* sy, long long rows, int bins, int ld, double* sums, cudaStream_t st) {
    CUM_REQUIRE(sx && sy && sums && rows > 0 && bins > 0 && ld >= 2 * bins && ld % 2 == 0, "stft_loss_reduce: bad arguments");
    const int grid = (int)(rows < 8LL * sm_count() ? rows : 8LL * sm_count());
    stft_loss_reduce_kernel<<<grid, 256, 0, st>>>(sx, sy, rows, bins, ld, sums);
    CUM_LAUNCH_CHECK("stft_loss_reduce_kernel");
    return CUM_OK;
}

// dS_x = d(k_sc * sqrt(A) / sqrt(B) + k_mag * C) / dS_x with coef = {k_sc / (sqrt(A) sqrt(B)), k_mag} read from device memory:
//   dxm = -coef0 (ym - xm) - coef1 sign(log ym - log xm) / xm ;  d(re, im) = dxm (re, im) / xm  (0 under the 1e-7 floor)
__global__ void __launch_bounds__(256) stft_loss_bwd_kernel(const float* __restrict__ sx, const float* __restrict__ sy, long long rows,
                                                             int bins, int ld, const float* __restrict__ coef, float* __restrict__ dsx) {
    const float k0 = __ldg(coef), k1 = __ldg(coef + 1);
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const float2* px = reinterpret_cast<const float2*>(sx + r * ld);
        const float2* py = reinterpret_cast<const float2*>(sy + r * ld);
        float2* pd = reinterpret_cast<float2*>(dsx + r * ld);
        for (int f = threadIdx.x; f < ld / 2; f += blockDim.x) {
            float2 o = make_float2(0.f, 0.f);
            if (f < bins) {
                const float2 vx = __ldg(px + f), vy = __ldg(py + f);
                const float pw = fmaf(vx.x, vx.x, vx.y * vx.y);
                if (pw > 1e-7f) {
                    const float xm = sqrtf(pw);
                    const float ym = sqrtf(fmaxf(fmaf(vy.x, vy.x, vy.y * vy.y), 1e-7f));
                    const float dl = logf(ym) - logf(xm);
                    const float sg = dl > 0.f ? 1.f : (dl < 0.f ? -1.f : 0.f);
                    const float dxm = -k0 * (ym - xm) - k1 * sg / xm;
                    const float s = dxm / xm;
                    o = make_float2(s * vx.x, s * vx.y);
                }
            }
            pd[f] = o;          // pad columns of the GEMM operand are written as zeros
        }
    }
}

int stft_loss_bwd(const float* sx, const float* sy, long long rows, int bins, int ld, const float* coef, float* dsx, cudaStream_t st) {
    CUM_REQUIRE(sx && sy && coef && dsx && rows > 0 && bins > 0 && ld >= 2 * bins && ld % 2 == 0, "stft_loss_bwd: bad arguments");
    const int grid = (int)(rows < 8LL * sm_count() ? rows : 8LL * sm_count());
    stft_loss_bwd_kernel<<<grid, 256, 0, st>>>(sx, sy, rows, bins, ld, coef, dsx);
    CUM_LAUNCH_CHECK("stft_loss_bwd_kernel");
    return CUM_OK;
}

// dx[b, reflect(fr * hop + k + first)] += dframes[(b, fr), k]   (transpose of stft_frames_kernel for the predicted signal)
__global__ void __launch_bounds__(256) stft_overlap_add_kernel(const float* __restrict__ dframes, int length, int n_frames, int hop,
                                                                int win, int first, float* __restrict__ dx, long long dx_stride) {
    const long long row = blockIdx.x;
    const int fr = (int)(row % n_frames);
    const int b = (int)(row / n_frames);
    const float* src = dframes + row * win;
    float* dst = dx + (long long)b * dx_stride;
    const int base = fr * hop + first;
    for (int k = threadIdx.x; k < win; k += blockDim.x) {
        int i = base + k;
        if (i < 0) i = -i;
        if (i >= length) i = 2 * (length - 1) - i;
        atomicAdd(dst + i, __ldg(src + k));
    }
}

int stft_overlap_add(const float* dframes, int length, int batch, int n_frames, int hop, int win, int n_fft, float* dx,
                     long long dx_stride, cudaStream_t st) {
    CUM_REQUIRE(dframes && dx && length > n_fft / 2 && batch > 0 && n_frames > 0, "stft_overlap_add: bad arguments");
    const long long rows = (long long)batch * n_frames;
    CUM_REQUIRE(rows < (1ll << 31), "stft_overlap_add: too many frames");
    stft_overlap_add_kernel<<<(unsigned)rows, 256, 0, st>>>(dframes, length, n_frames, hop, win, (n_fft - win) / 2 - n_fft / 2, dx, dx_stride);
    CUM_LAUNCH_CHECK("stft_overlap_add_kernel");
    return CUM_OK;
}

}  // namespace cum
