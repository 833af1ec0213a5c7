// Channel-importance statistics of one weight matrix and its gradient -- the quantities the reference's pruning criteria are
// built from (/root/reference/src/pruning/pruninggroup.py:160-226, importance.py:39): per channel
//   weight = sum w^2, grad = sum g^2, taylor_individual = sum |w g|, taylor_squared_individual = sum (w g)^2, taylor_sum = sum w g
// (taylor_group = |taylor_sum|).  ONE pass over w and g (HBM-bound: 8 bytes per element) yields the five statistics for every
// row (dim-0 channels) AND every column (dim-1 channels) of the (rows, cols) matrix; the reference makes ~10 elementwise /
// reduction passes per (module, dim).  The host (importance.py) regroups rows / columns into channels (heads, conv taps).
#include "common.cuh"

namespace cum {

constexpr int IMP_ROWS = 32;      // rows per CTA strip
constexpr int IMP_THREADS = 256;

struct Stat5 {
    float w2, g2, ti, t2, ts;
    __device__ __forceinline__ void add(float w, float g) {
        const float p = w * g;
        w2 = fmaf(w, w, w2); g2 = fmaf(g, g, g2); ti += fabsf(p); t2 = fmaf(p, p, t2); ts += p;
    }
};

__global__ void __launch_bounds__(IMP_THREADS) channel_importance_kernel(const float* __restrict__ w, const float* __restrict__ g,
                                                                          int rows, int cols, long long ldw, long long ldg,
                                                                          float* __restrict__ out_rows, float* __restrict__ out_cols) {
    __shared__ float red[IMP_THREADS / 32][5];
    const int r0 = blockIdx.x * IMP_ROWS;
    const int nr = min(IMP_ROWS, rows - r0);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // columns: thread c accumulates its column over the strip's rows (coalesced row reads), one atomic per statistic per strip
    for (int c = threadIdx.x; c < cols; c += IMP_THREADS) {
        Stat5 s = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int r = 0; r < nr; ++r) s.add(__ldg(w + (long long)(r0 + r) * ldw + c), __ldg(g + (long long)(r0 + r) * ldg + c));
        if (out_cols) {
            atomicAdd(out_cols + c, s.w2); atomicAdd(out_cols + cols + c, s.g2); atomicAdd(out_cols + 2ll * cols + c, s.ti);
            atomicAdd(out_cols + 3ll * cols + c, s.t2); atomicAdd(out_cols + 4ll * cols + c, s.ts);
        }
    }
    if (!out_rows) return;
    // rows: the strip's data is L1 / L2 resident from the column pass; block-reduce each row
    for (int r = 0; r < nr; ++r) {
        Stat5 s = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int c = threadIdx.x; c < cols; c += IMP_THREADS) s.add(__ldg(w + (long long)(r0 + r) * ldw + c), __ldg(g + (long long)(r0 + r) * ldg + c));
        float v[5] = {s.w2, s.g2, s.ti, s.t2, s.ts};
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] = warp_sum(v[k]);
        __syncthreads();
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < 5; ++k) red[wid][k] = v[k];
        __syncthreads();
        if (threadIdx.x < 5) {
            float t = 0.f;
            for (int i = 0; i < IMP_THREADS / 32; ++i) t += red[i][threadIdx.x];
            out_rows[(long long)threadIdx.x * rows + r0 + r] = t;
        }
    }
}

int channel_importance_fwd(const float* w, const float* g, int rows, int cols, long long ldw, long long ldg, float* out_rows,
                           float* out_cols, cudaStream_t st) {
    CUM_REQUIRE(w && g && rows > 0 && cols > 0 && ldw >= cols && ldg >= cols, "channel_importance: bad arguments");
    CUM_REQUIRE(out_rows || out_cols, "channel_importance: no output requested");
    if (out_cols) {
        cudaError_t e = cudaMemsetAsync(out_cols, 0, sizeof(float) * 5ull * cols, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(channel_importance)");
    }
    channel_importance_kernel<<<(unsigned)cdiv(rows, IMP_ROWS), IMP_THREADS, 0, st>>>(w, g, rows, cols, ldw, ldg, out_rows, out_cols);
    CUM_LAUNCH_CHECK("channel_importance_kernel");
    return CUM_OK;
}

}  // namespace cum
