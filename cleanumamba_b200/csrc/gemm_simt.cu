// Exact-fp32 tap-GEMM on the CUDA cores (CUM_MATH_FP32).
//
//   out[b, m, :] = EPI( bias + sum_{s<taps} W_s . a[b, m + shift_s, 0:k] ) (+ addend[b, m, :])
//
// Products and accumulation are plain fp32 FFMA, i.e. the reference's arithmetic (its CPU path), so this kernel is
// (a) the fp32-exact mode of the product and (b) the on-device checker for the tcgen05 modes at sizes the CPU
// oracle cannot reach.  128x128x16 CTA tile, 256 threads, 8x8 register micro-tile split as 2x2 blocks of 4x4 so
// that every shared-memory read is a conflict-free float4, register-staged double buffering of the global loads.
#include "common.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

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

// ---------------------------------------------------------------------------------------------------------
// Skinny tap-GEMM: the whole problem is a few output rows (one stream fed hop by hop -- the reference's real-time use,
// CleanUMamba.py:371-418 / examples/streaming_demo.py -- or a handful of streams: every GEMM of a call is 1-33 rows).  The tcgen05
// kernel is a persistent pipeline (TMEM allocation, barrier ring, tensor-map prefetch, 128-row tiles): ~14 us per launch whatever the
// size, 36 launches per hop.  Here a warp owns TWO output columns (one GLU pair) for up to 8 rows, its lanes split K (coalesced
// 16-byte weight loads, the few A rows come from L1), a butterfly reduction ends the K loop and lane r finishes row r (bias,
// activation / gate, addend).  Products are exact fp32 FMAs on the full-precision weights (split modes: hi + lo halves), so this
// path is at least as accurate as the tensor-core modes it stands in for.  Same descriptor semantics as the kernels above.
// ---------------------------------------------------------------------------------------------------------
constexpr int SK_AUTO_ROWS = 4;                 // automatic choice: at most 4 output rows in total (m * batch) ...
constexpr long long SK_MAX_MACS = 16000000;     // ... and at most 16 M multiply-adds.  The kernel runs at ~0.4 TMAC/s with ~4 us of fixed cost
                                               // against ~14 us for the tensor-core launch, and it loses efficiency with the rows per pass.
                                               // Measured (E6 full, 1 hop per call from the graph; every GEMM here vs none): 1 stream 0.40 vs
                                               // 0.73 ms, 2 streams 0.45 vs 0.73, 4 streams 0.59 vs 0.73, 8 streams 0.83 vs 0.74, 16 streams
                                               // 1.37 vs 0.75; mixed per-GEMM rules by bytes or work were worse than all-or-nothing from 8 streams
constexpr int SK_M = 8;             // rows per warp pass
constexpr int SK_WARPS = 4;

// four consecutive weights w[k .. k+3] of one weight row; WF: 0 fp32, 1 fp32 hi + lo (TF32X3), 2 fp16 hi + lo (F16X3, scaled), 3 bf16 hi + lo
template <int WF>
__device__ __forceinline__ float4 sk_load_w(const void* w, const void* w_lo, long long off) {
    if (WF == 0) return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(w) + off));
    if (WF == 1) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(w) + off));
        const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(w_lo) + off));
        return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(w) + off));
    const uint2 l = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(w_lo) + off));
    float2 h0, h1, l0, l1;
    if (WF == 2) {
        h0 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)); h1 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
        l0 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)); l1 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
    } else {
        h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.x)); h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.y));
        l0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l.x)); l1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l.y));
    }
    return make_float4(h0.x + l0.x, h0.y + l0.y, h1.x + l1.x, h1.y + l1.y);
}

struct SkinnyParams { SimtParams s; const void* w_lo; float acc_scale; };

// RM: rows per pass (1, 2, 4 or 8: the few-row cases keep more weight loads in flight -- a lane's K loop is a chain of L2 round trips)
template <int WF, int RM>
__global__ void __launch_bounds__(SK_WARPS * 32) gemm_skinny_kernel(const SkinnyParams q) {
    constexpr int SK_M = RM;
    pdl_trigger();
    pdl_wait();          // PDL: nothing of the previous kernel is touched before this point
    const SimtParams& p = q.s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int n = (blockIdx.x * SK_WARPS + warp) * 2;          // this warp's two output columns (weight rows n, n + 1)
    if (n >= p.n) return;
    const int nh_off = (p.n_half && (b & 1)) ? p.n : 0;
    const bool glu = p.epi >= CUM_EPI_GLU_SIGMOID;
    for (int m0 = 0; m0 < p.m; m0 += SK_M) {
        float acc[SK_M][2];
#pragma unroll
        for (int r = 0; r < SK_M; ++r) acc[r][0] = acc[r][1] = 0.f;
        for (int tap = 0; tap < p.taps; ++tap) {
            const int shift = tap == 0 ? p.shift0 : p.shift1;
            const long long wrow = (long long)tap * p.w_tap_stride + (long long)(n + nh_off) * p.ldw;
#pragma unroll (RM <= 2 ? 8 : (RM == 4 ? 4 : 2))
            for (int kk = lane * 4; kk < p.k; kk += 128) {
                const float4 w0 = sk_load_w<WF>(p.w, q.w_lo, wrow + kk);
                const float4 w1 = sk_load_w<WF>(p.w, q.w_lo, wrow + p.ldw + kk);
                // A addressing as in gemm_simt_kernel (plane-major: the tap shift and the upper part of K select the plane)
                int ka = kk, rshift = shift;
                const float* abase = p.a + (long long)b * p.a_bs;
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
                for (int r = 0; r < SK_M; ++r) {
                    const int row = m0 + r + rshift;
                    if (plane_ok && m0 + r < p.m && row >= 0 && row < p.a_rows) {
                        const float4 av = __ldg(reinterpret_cast<const float4*>(abase + (long long)row * p.a_rs + ka));
                        acc[r][0] = fmaf(av.x, w0.x, fmaf(av.y, w0.y, fmaf(av.z, w0.z, fmaf(av.w, w0.w, acc[r][0]))));
                        acc[r][1] = fmaf(av.x, w1.x, fmaf(av.y, w1.y, fmaf(av.z, w1.z, fmaf(av.w, w1.w, acc[r][1]))));
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < SK_M; ++r) { acc[r][0] = warp_sum(acc[r][0]); acc[r][1] = warp_sum(acc[r][1]); }
        // lane r finishes row m0 + r (every lane holds every total after the butterfly)
        float v0 = 0.f, v1 = 0.f;
#pragma unroll
        for (int r = 0; r < SK_M; ++r) if (lane == r) { v0 = acc[r][0]; v1 = acc[r][1]; }
        const int m = m0 + lane;
        if (lane < SK_M && m < p.m) {
            v0 *= q.acc_scale; v1 *= q.acc_scale;
            if (p.bias) { v0 += __ldg(p.bias + n); v1 += __ldg(p.bias + n + 1); }
            float* crow = p.c + (long long)b * p.c_bs + (long long)m * p.c_rs;
            const float* arow = p.addend ? p.addend + (long long)b * p.add_bs + (long long)m * p.add_rs : nullptr;
            if (glu) {
                float o = v0 * glu_gate(p.epi, v1);
                if (arow) o += __ldg(arow + (n >> 1));
                crow[n >> 1] = o;
            } else {
                float o0 = unary_act(p.epi, v0), o1 = unary_act(p.epi, v1);
                if (arow) { o0 += __ldg(arow + n); o1 += __ldg(arow + n + 1); }
                *reinterpret_cast<float2*>(crow + n) = make_float2(o0, o1);
            }
        }
    }
}

bool gemm_skinny_ok(const cum_gemm_desc& d) {
    static int env = -1;
    if (env < 0) { const char* e = getenv("CUM_GEMM_SKINNY"); env = (e && e[0] == '0') ? 0 : 1; }
    if (!env || d.math == CUM_MATH_BF16 || d.math == CUM_MATH_TF32 || d.a_lo || d.c_lo || d.addend_lo || d.out_bf16 || d.aux || d.addend_is_mask ||
        d.a_scale_dev || d.batch > 65535)
        return false;
    if ((d.math == CUM_MATH_TF32X3 || d.math == CUM_MATH_BF16X3 || d.math == CUM_MATH_F16X3) && !d.w_lo) return false;
    if (d.math != CUM_MATH_FP32 && d.math != CUM_MATH_TF32X3 && d.ldw % 8) return false;       // 16-bit weight rows: 8-byte aligned quads
    if (d.small_m_path) return d.small_m_path > 0;
    return (long long)d.batch * d.m <= SK_AUTO_ROWS && (long long)d.batch * d.m * d.n * d.k * d.taps <= SK_MAX_MACS;
}

int gemm_skinny_fwd(const cum_gemm_desc& d, cudaStream_t st) {
    SkinnyParams q;
    SimtParams& p = q.s;
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
    q.w_lo = d.w_lo;
    q.acc_scale = d.math == CUM_MATH_F16X3 ? d.acc_scale : 1.0f;
    CUM_REQUIRE(d.math != CUM_MATH_F16X3 || d.acc_scale > 0.f, "gemm: F16X3 needs acc_scale = 1 / (weight scale passed to cum_split_f16)");
    const dim3 grid((unsigned)cdiv(d.n / 2, SK_WARPS), (unsigned)d.batch);
    cudaError_t e;
    const int wf = d.math == CUM_MATH_TF32X3 ? 1 : d.math == CUM_MATH_F16X3 ? 2 : d.math == CUM_MATH_BF16X3 ? 3 : 0;
    // rows per pass; m > 4 keeps 8 rows per pass (two 4-row passes measured slower: 0.96 vs 0.82 ms per call at 8 streams -- and both
    // slower than the tensor-core path there, 0.74 ms: the sessions use this kernel up to 4 streams)
    const int rm = d.m <= 1 ? 1 : d.m <= 2 ? 2 : d.m <= 4 ? 4 : 8;
#define SK_LAUNCH(WF, RM) e = launch_kernel(gemm_skinny_kernel<WF, RM>, grid, dim3(SK_WARPS * 32), 0, st, q)
#define SK_ROWS(WF) do { if (rm == 1) SK_LAUNCH(WF, 1); else if (rm == 2) SK_LAUNCH(WF, 2); else if (rm == 4) SK_LAUNCH(WF, 4); else SK_LAUNCH(WF, 8); } while (0)
    switch (wf) {
        case 1: SK_ROWS(1); break;
        case 2: SK_ROWS(2); break;
        case 3: SK_ROWS(3); break;
        default: SK_ROWS(0); break;
    }
#undef SK_ROWS
#undef SK_LAUNCH
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(gemm_skinny_kernel)");
    return CUM_OK;
}

}  // namespace cum
