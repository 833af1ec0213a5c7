// Exact-fp32 tap-GEMM on the CUDA cores (CUM_MATH_FP32).
//
//   out[b, m, :] = EPI( bias + sum_{s<taps} W_s . a[b, m + shift_s, 0:k] ) (+ addend[b, m, :])
//
// Products and accumulation are plain fp32 FFMA, i.e. the reference's arithmetic (its CPU path), so this kernel is
// (a) the fp32-exact mode of the product and (b) the on-device checker for the tcgen05 modes at sizes the CPU
// oracle cannot reach.  128x128x16 CTA tile, 256 threads, 8x8 register micro-tile split as 2x2 blocks of 4x4 so
// that every shared-memory read is a conflict-free float4, register-staged double buffering of the global loads.
#include "common.cuh"

namespace cum {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_THREADS = 256;

struct SimtParams {
    const float* a; long long a_bs, a_rs; int a_rows, k, taps, shift0, shift1;
    const float* w; int ldw; long long w_tap_stride;
    const float* bias;
    float* c; long long c_bs, c_rs; int m, n, epi;
    const float* addend; long long add_bs, add_rs;
    int a_planes, a_plane_k, a_plane0, a_plane_step, n_half;     // plane-major A (cum_gemm_desc.a_planes; 0 = off)
};

__global__ void __launch_bounds__(SG_THREADS) gemm_simt_kernel(const SimtParams p) {
    __shared__ __align__(16) float As[2][SG_BK][SG_BM + 4];
    __shared__ __align__(16) float Ws[2][SG_BK][SG_BN + 4];

    const int b = blockIdx.z;
    const int m0 = blockIdx.x * SG_BM;
    const int n0 = blockIdx.y * SG_BN;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 thread grid

    // global->smem staging: each thread moves 2 float4 of A and 2 float4 of W per k-block
    // tile element (row r, k-quad q): r = (tid >> 2) + 64*i, q = tid & 3
    const int lr = tid >> 2, lq = tid & 3;
    const float* ab = p.a + (long long)b * p.a_bs;
    const int nh_off = (p.n_half && (b & 1)) ? p.n : 0;       // n_half: weight rows of this batch item's column half (the bias is shared)

    const int kblocks = (p.k + SG_BK - 1) / SG_BK;
    const int total = kblocks * p.taps;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rw[2];
    auto load_global = [&](int it) {
        const int tap = it / kblocks, kb = it - tap * kblocks;
        const int shift = tap == 0 ? p.shift0 : p.shift1;
        const int kk = kb * SG_BK + lq * 4;
        // plane-major A: the tap shift and the upper part of K select the plane (see cum_gemm_desc.a_planes)
        int ka = kk, rshift = shift;
        const float* abase = ab;
        bool plane_ok = true;
        if (p.a_planes) {
            const int kp = kk / p.a_plane_k;
            const int pl = p.a_plane0 + p.a_plane_step * ((p.n_half ? b >> 1 : b) + shift) + kp;
            ka = kk - kp * p.a_plane_k;
            rshift = 0;
            plane_ok = pl >= 0 && pl < p.a_planes;
            abase = p.a + (long long)pl * p.a_bs;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int row = m0 + lr + 64 * i + rshift;
            ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (plane_ok && row >= 0 && row < p.a_rows && (m0 + lr + 64 * i) < p.m && kk < p.k)
                ra[i] = __ldg(reinterpret_cast<const float4*>(abase + (long long)row * p.a_rs + ka));
            const int nn = n0 + lr + 64 * i;
            rw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (nn < p.n && kk < p.k)
                rw[i] = __ldg(reinterpret_cast<const float4*>(p.w + (long long)tap * p.w_tap_stride + (long long)(nn + nh_off) * p.ldw + kk));
        }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = lr + 64 * i;
            As[buf][lq * 4 + 0][r] = ra[i].x; As[buf][lq * 4 + 1][r] = ra[i].y;
            As[buf][lq * 4 + 2][r] = ra[i].z; As[buf][lq * 4 + 3][r] = ra[i].w;
            Ws[buf][lq * 4 + 0][r] = rw[i].x; Ws[buf][lq * 4 + 1][r] = rw[i].y;
            Ws[buf][lq * 4 + 2][r] = rw[i].z; Ws[buf][lq * 4 + 3][r] = rw[i].w;
        }
    };

    load_global(0);
    store_smem(0);
    __syncthreads();
    for (int it = 0; it < total; ++it) {
        const int buf = it & 1;
        if (it + 1 < total) load_global(it + 1);
#pragma unroll
        for (int kk = 0; kk < SG_BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
            const float4 w1 = *reinterpret_cast<const float4*>(&Ws[buf][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        if (it + 1 < total) {
            store_smem(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue: rows m0 + ty*4 + {0..3} (+64), column quads n0 + tx*4 (+64)
    const bool glu = p.epi >= CUM_EPI_GLU_SIGMOID;
#pragma unroll
    for (int ih = 0; ih < 2; ++ih) {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int m = m0 + ih * 64 + ty * 4 + ii;
            if (m >= p.m) continue;
            float* crow = p.c + (long long)b * p.c_bs + (long long)m * p.c_rs;
            const float* arow = p.addend ? p.addend + (long long)b * p.add_bs + (long long)m * p.add_rs : nullptr;
#pragma unroll
            for (int jh = 0; jh < 2; ++jh) {
                const int n = n0 + jh * 64 + tx * 4;
                if (n >= p.n) continue;  // n is a multiple of 8: quads are all-in or all-out
                float v[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) v[jj] = acc[ih * 4 + ii][jh * 4 + jj];
                if (p.bias) {
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                    v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w;
                }
                if (glu) {
                    float2 o;
                    o.x = v[0] * glu_gate(p.epi, v[1]);
                    o.y = v[2] * glu_gate(p.epi, v[3]);
                    const int oc = n >> 1;
                    if (arow) {
                        const float2 ad = __ldg(reinterpret_cast<const float2*>(arow + oc));
                        o.x += ad.x; o.y += ad.y;
                    }
                    *reinterpret_cast<float2*>(crow + oc) = o;
                } else {
                    float4 o = make_float4(unary_act(p.epi, v[0]), unary_act(p.epi, v[1]), unary_act(p.epi, v[2]),
                                           unary_act(p.epi, v[3]));
                    if (arow) {
                        const float4 ad = __ldg(reinterpret_cast<const float4*>(arow + n));
                        o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w;
                    }
                    *reinterpret_cast<float4*>(crow + n) = o;
                }
            }
        }
    }
}

int gemm_simt_fwd(const cum_gemm_desc& d, cudaStream_t st) {
    SimtParams p;
    p.a = d.a; p.a_bs = d.a_batch_stride; p.a_rs = d.a_row_stride; p.a_rows = d.a_rows; p.k = d.k; p.taps = d.taps;
    p.shift0 = d.tap_shift[0]; p.shift1 = d.tap_shift[1];
    p.w = d.w; p.ldw = d.ldw; p.w_tap_stride = (long long)d.n * d.ldw;
    p.bias = d.bias; p.c = d.c; p.c_bs = d.c_batch_stride; p.c_rs = d.c_row_stride; p.m = d.m; p.n = d.n;
    p.epi = d.epilogue; p.addend = d.addend; p.add_bs = d.add_batch_stride; p.add_rs = d.add_row_stride;
    p.a_planes = d.a_planes > 0 ? d.a_planes : 0; p.a_plane_k = d.a_plane_k; p.a_plane0 = d.a_plane0; p.a_plane_step = d.a_plane_step;
    p.n_half = (d.a_planes > 0 && d.n_half) ? 1 : 0;
    if (p.a_planes) {
        CUM_REQUIRE(d.a_plane_k > 0 && d.a_plane_k % 4 == 0, "gemm: plane-major a needs a_plane_k %% 4 == 0");
        CUM_REQUIRE(!p.n_half || !(d.epilogue >= CUM_EPI_GLU_SIGMOID), "gemm: n_half needs a non-GLU epilogue");
        if (p.n_half) p.w_tap_stride = 2LL * d.n * d.ldw;
    }
    dim3 grid((unsigned)cdiv(d.m, SG_BM), (unsigned)cdiv(d.n, SG_BN), (unsigned)d.batch);
    CUM_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm: n=%d or batch=%d too large for the grid", d.n, d.batch);
    gemm_simt_kernel<<<grid, SG_THREADS, 0, st>>>(p);
    CUM_LAUNCH_CHECK("gemm_simt_kernel");
    return CUM_OK;
}

}  // namespace cum
