// Backward kernels of the CleanUMamba path (training: SURVEY.md §8 row a14).
//
//   glu_fwd / glu_bwd           GLU gate kept separate in training so the pre-activation Z is saved once
//   relu_bwd, colsum            activation masks + bias gradients (column sums with one atomic per CTA column)
//   add                         U-Net skip add kept separate in training (the ReLU mask needs the un-added value)
//   wgrad_simt                  dW_s[n,k] = sum_rows dZ[row,n] * A[row+shift_s,k]   (fp32 FFMA, split over rows)
//   ln_bwd                      LayerNorm backward + residual-stream gradient add, dgamma / dbeta
//   dwconv_silu_bwd             depthwise causal conv + SiLU backward (dx, dw, db)
//   conv_in_bwd / convt_out_bwd weight gradients of the waveform-end layers (+ dg for the last transposed conv)
// The data-gradient of every dense layer is the forward tap-GEMM with transposed packed weights (same kernel).
// The reverse selective scan lives in scan_bwd.cu.
#include "common.cuh"

namespace cum {

// ---------------------------------------------------------------------------------------------------------
// GLU on an interleaved pre-activation Z (rows, 2H): out[r,c] = Z[r,2c] * sigmoid(Z[r,2c+1]) (+ addend)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) glu_fwd_kernel(const float* __restrict__ z, const float* __restrict__ addend,
                                                       float* __restrict__ out, long long n_out4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;    // float4 of outputs
    if (i >= n_out4) return;
    const float4 z0 = reinterpret_cast<const float4*>(z)[2 * i], z1 = reinterpret_cast<const float4*>(z)[2 * i + 1];
    float4 o = make_float4(z0.x * sigmoidf_(z0.y), z0.z * sigmoidf_(z0.w), z1.x * sigmoidf_(z1.y), z1.z * sigmoidf_(z1.w));
    if (addend) {
        const float4 a = reinterpret_cast<const float4*>(addend)[i];
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    reinterpret_cast<float4*>(out)[i] = o;
}

int glu_fwd(const float* z, const float* addend, float* out, long long rows, int h_pad, cudaStream_t st) {
    CUM_REQUIRE(z && out && rows > 0 && h_pad > 0 && h_pad % 4 == 0, "glu_fwd: bad arguments");
    const long long n4 = rows * h_pad / 4;
    glu_fwd_kernel<<<(unsigned)cdiv(n4, 256), 256, 0, st>>>(z, addend, out, n4);
    CUM_LAUNCH_CHECK("glu_fwd_kernel");
    return CUM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Row-block kernels with fused column sums.  CTA = RB_ROWS rows x all columns; a thread owns one float4 column group
// (looping when there are more than 256 groups) and walks the CTA's rows, so the bias gradient costs one atomicAdd
// per column per CTA.
//   MODE 0: glu_bwd   dz[r,2c] = dout*sig(b) ; dz[r,2c+1] = dout*a*sig(b)*(1-sig(b))     (z: (rows,2H), dout: (rows,H))
//   MODE 1: relu_bwd  dz = dy * (y > 0)
//   MODE 2: colsum    no elementwise output, only the column sums of `dout`
// ---------------------------------------------------------------------------------------------------------
constexpr int RB_ROWS = 128;      // minimum rows per CTA; large problems use more so the column atomics stay ~1 per SM wave

// Column partial sums of the `slots` row slots of a CTA meet in shared memory (red: 256 float4) so that ONE thread per column group
// issues the atomics.  Must be called by every thread of the CTA; returns true for the threads (slot 0) that hold the CTA total.
__device__ __forceinline__ bool cta_slot_reduce(float4& acc, float4* red, int slots, int gpr, int tg, int ts) {
    if (slots <= 1) return ts == 0;
    __syncthreads();                 // a previous use of `red` is over
    red[threadIdx.x] = acc;
    __syncthreads();
    if (ts != 0) return false;
    for (int sl = 1; sl < slots; ++sl) {
        const float4 o = red[sl * gpr + tg];
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    return true;
}

template <int MODE>
__global__ void __launch_bounds__(256) rowblock_bwd_kernel(const float* __restrict__ z, const float* __restrict__ dout,
                                                            float* __restrict__ dz, float* __restrict__ dbias,
                                                            long long rows, int cols /* of dz / z */, int rows_per_cta,
                                                            unsigned* __restrict__ amax_bits /* optional: max |dz| for the gradient scale */) {
    const int groups = cols >> 2;                       // float4 groups per row
    const int gpr = groups < 256 ? groups : 256;        // groups handled per pass
    const int slots = 256 / gpr;                        // row slots per pass
    const int tg = threadIdx.x % gpr, ts = threadIdx.x / gpr;
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const long long r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
    // slots > 1 (fewer than 129 column groups): every column is summed by `slots` threads of the CTA.  Their partial sums meet in
    // shared memory so that ONE atomic per column and CTA reaches the bias gradient -- with 16 slots x ~1200 CTAs the same 64
    // addresses took 19 k serialised atomics each, a third of the kernel's time on the 64-channel levels.
    __shared__ float4 red[256];
    const bool active = ts < slots;
    float amax = 0.f;
    for (int g = tg; g < groups; g += gpr) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active) {
        auto grad = [&](const float4& zv, const float4& dv) {
            if (MODE == 0) {
                const float s0 = sigmoidf_(zv.y), s1 = sigmoidf_(zv.w);
                return make_float4(dv.x * s0, dv.x * zv.x * s0 * (1.f - s0), dv.y * s1, dv.y * zv.z * s1 * (1.f - s1));
            } else if (MODE == 1) {
                return make_float4(zv.x > 0.f ? dv.x : 0.f, zv.y > 0.f ? dv.y : 0.f, zv.z > 0.f ? dv.z : 0.f, zv.w > 0.f ? dv.w : 0.f);
            }
            return dv;
        };
        auto load_z = [&](long long r) { return MODE == 2 ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<const float4*>(z + r * cols)[g]; };
        auto load_d = [&](long long r) {
            if (MODE == 0) { const float2 d2 = reinterpret_cast<const float2*>(dout + r * (cols >> 1))[g]; return make_float4(d2.x, d2.y, 0.f, 0.f); }
            return reinterpret_cast<const float4*>(dout + r * cols)[g];
        };
        auto finish = [&](long long r, const float4& d) {
            if (MODE != 2) reinterpret_cast<float4*>(dz + r * cols)[g] = d;
            acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w))));
        };
        // four rows per iteration, every load issued before the first use: the kernel is a pure stream (48-80 bytes per thread in
        // flight instead of 16-24 -- one row at a time ran at 1.4 (colsum) to 3.5 TB/s (GLU backward))
        constexpr int UNR = 4;
        long long r = r0 + ts;
        for (; r + (UNR - 1) * slots < r1; r += UNR * slots) {
            float4 zv[UNR], dv[UNR];
#pragma unroll
            for (int i = 0; i < UNR; ++i) { zv[i] = load_z(r + i * slots); dv[i] = load_d(r + i * slots); }
#pragma unroll
            for (int i = 0; i < UNR; ++i) finish(r + i * slots, grad(zv[i], dv[i]));
        }
        for (; r < r1; r += slots) finish(r, grad(load_z(r), load_d(r)));
        }
        if (!cta_slot_reduce(acc, red, slots, gpr, tg, ts)) continue;      // uniform call; the g loop has a single iteration when slots > 1
        if (dbias && active) {
            atomicAdd(dbias + 4 * g + 0, acc.x); atomicAdd(dbias + 4 * g + 1, acc.y);
            atomicAdd(dbias + 4 * g + 2, acc.z); atomicAdd(dbias + 4 * g + 3, acc.w);
        }
    }
    if (amax_bits && active) {            // one atomic per warp (non-negative floats order like their bits)
        const unsigned act = __activemask();
        const unsigned m = __reduce_max_sync(act, __float_as_uint(amax));
        if ((threadIdx.x & 31) == (unsigned)(__ffs(act) - 1) && m != 0u) atomicMax(amax_bits, m);
    }
}

int rowblock_bwd(int mode, const float* z, const float* dout, float* dz, float* dbias, long long rows, int cols,
                 cudaStream_t st, float* scale4) {
    CUM_REQUIRE(dout && rows > 0 && cols > 0 && cols % 4 == 0, "rowblock_bwd: bad arguments");
    CUM_REQUIRE(mode == 2 || (z && dz), "rowblock_bwd: z/dz required");
    CUM_REQUIRE(mode != 0 || cols % 8 == 0, "glu_bwd: cols must be a multiple of 8");
    long long rpc = cdiv(rows, 8LL * sm_count());      // ~8 CTAs per SM at most -> few thousand atomics per column
    if (rpc < RB_ROWS) rpc = RB_ROWS;
    const unsigned grid = (unsigned)cdiv(rows, rpc);
    unsigned* amax_bits = scale4 ? reinterpret_cast<unsigned*>(scale4) + 2 : nullptr;
    if (scale4) {
        cudaError_t e = cudaMemsetAsync(scale4, 0, 16, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(rowblock_bwd scale)");
    }
    if (mode == 0) rowblock_bwd_kernel<0><<<grid, 256, 0, st>>>(z, dout, dz, dbias, rows, cols, (int)rpc, amax_bits);
    else if (mode == 1) rowblock_bwd_kernel<1><<<grid, 256, 0, st>>>(z, dout, dz, dbias, rows, cols, (int)rpc, amax_bits);
    else rowblock_bwd_kernel<2><<<grid, 256, 0, st>>>(z, dout, dz, dbias, rows, cols, (int)rpc, amax_bits);
    CUM_LAUNCH_CHECK("rowblock_bwd_kernel");
    if (scale4) return grad_scale_finalize(scale4, st);      // {s, 1/s} of dz for the f16x3 gradient GEMMs: no separate amax pass
    return CUM_OK;
}

__global__ void __launch_bounds__(256) add_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                   float* __restrict__ out, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    reinterpret_cast<float4*>(out)[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
}

int add_fwd(const float* a, const float* b, float* out, long long count, cudaStream_t st) {
    CUM_REQUIRE(a && b && out && count > 0 && count % 4 == 0, "add: bad arguments");
    add_kernel<<<(unsigned)cdiv(count / 4, 256), 256, 0, st>>>(a, b, out, count / 4);
    CUM_LAUNCH_CHECK("add_kernel");
    return CUM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// wgrad: dW_s[n, k] += sum_{b, row < m} dZ[b,row,n] * A[b, row + shift_s, k]      (rows of A outside [0,a_rows) = 0)
// 128(n) x 128(k) tile per CTA, 16 rows per smem step, the row range split over gridDim.z CTAs that finish with
// atomicAdd into the zero-initialised fp32 gradient.  fp32 FFMA (tensor-core wgrad is future work: both operands are
// MN-major activations).
// ---------------------------------------------------------------------------------------------------------
struct WgradParams {
    const float* dz; long long dz_bs, dz_rs;
    const float* a;  long long a_bs, a_rs; int a_rows;
    float* dw; int ldw; long long w_tap_stride;
    int m, n, k, taps, shift0, shift1, batch, rows_per_cta;
};

__global__ void __launch_bounds__(256) wgrad_simt_kernel(const WgradParams p) {
    __shared__ __align__(16) float Zs[2][16][128 + 4];
    __shared__ __align__(16) float As[2][16][128 + 4];
    const int n0 = blockIdx.x * 128, k0 = blockIdx.y * 128;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int chunks_per_batch = (p.m + p.rows_per_cta - 1) / p.rows_per_cta;
    const int b = blockIdx.z / chunks_per_batch;
    const int rbeg = (blockIdx.z % chunks_per_batch) * p.rows_per_cta;
    const int rend = min(p.m, rbeg + p.rows_per_cta);
    const float* zb = p.dz + (long long)b * p.dz_bs;
    const float* ab = p.a + (long long)b * p.a_bs;
    // staging: thread moves 2 float4 per operand per step: row = (tid >> 5) + 8*i, column quad = tid & 31
    const int lr = tid >> 5, lq = tid & 31;

    for (int tap = 0; tap < p.taps; ++tap) {
        const int shift = tap == 0 ? p.shift0 : p.shift1;
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        float4 rz[2], ra[2];
        auto load_global = [&](int r) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = r + lr + 8 * i;
                rz[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < rend) {
                    if (n0 + lq * 4 < p.n) rz[i] = __ldg(reinterpret_cast<const float4*>(zb + (long long)row * p.dz_rs + n0 + lq * 4));
                    const int arow = row + shift;
                    if (arow >= 0 && arow < p.a_rows && k0 + lq * 4 < p.k)
                        ra[i] = __ldg(reinterpret_cast<const float4*>(ab + (long long)arow * p.a_rs + k0 + lq * 4));
                }
            }
        };
        auto store_smem = [&](int buf) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                *reinterpret_cast<float4*>(&Zs[buf][lr + 8 * i][lq * 4]) = rz[i];
                *reinterpret_cast<float4*>(&As[buf][lr + 8 * i][lq * 4]) = ra[i];
            }
        };
        const int steps = (rend - rbeg + 15) / 16;
        if (steps > 0) {
            load_global(rbeg);
            store_smem(0);
        }
        __syncthreads();
        for (int s = 0; s < steps; ++s) {
            const int buf = s & 1;
            if (s + 1 < steps) load_global(rbeg + (s + 1) * 16);
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                const float4 z0 = *reinterpret_cast<const float4*>(&Zs[buf][kk][ty * 4]);
                const float4 z1 = *reinterpret_cast<const float4*>(&Zs[buf][kk][64 + ty * 4]);
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + tx * 4]);
                const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(zv[i], av[j], acc[i][j]);
            }
            if (s + 1 < steps) store_smem(buf ^ 1);
            __syncthreads();
        }
        float* dw = p.dw + (long long)tap * p.w_tap_stride;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = n0 + (i >> 2) * 64 + ty * 4 + (i & 3);
            if (n >= p.n) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + (j >> 2) * 64 + tx * 4 + (j & 3);
                if (k < p.k) atomicAdd(dw + (long long)n * p.ldw + k, acc[i][j]);
            }
        }
        __syncthreads();
    }
}

int wgrad_fwd(const cum_wgrad_desc& d, cudaStream_t st) {
    CUM_REQUIRE(d.dz && d.a && d.dw, "wgrad: null pointer");
    CUM_REQUIRE(d.batch > 0 && d.m > 0 && d.n > 0 && d.k > 0 && (d.taps == 1 || d.taps == 2), "wgrad: bad shape");
    CUM_REQUIRE(d.n % 4 == 0 && d.k % 4 == 0 && d.ldw >= d.k, "wgrad: n, k must be multiples of 4 and ldw >= k");
    CUM_REQUIRE(d.dz_row_stride % 4 == 0 && d.dz_batch_stride % 4 == 0 && d.a_row_stride % 4 == 0 && d.a_batch_stride % 4 == 0 &&
                aligned16(d.dz) && aligned16(d.a), "wgrad: operands must be 16-byte aligned with strides multiple of 4");
    WgradParams p;
    p.dz = d.dz; p.dz_bs = d.dz_batch_stride; p.dz_rs = d.dz_row_stride;
    p.a = d.a; p.a_bs = d.a_batch_stride; p.a_rs = d.a_row_stride; p.a_rows = d.a_rows;
    p.dw = d.dw; p.ldw = d.ldw; p.w_tap_stride = (long long)d.n * d.ldw;
    p.m = d.m; p.n = d.n; p.k = d.k; p.taps = d.taps; p.shift0 = d.tap_shift[0]; p.shift1 = d.tap_shift[1]; p.batch = d.batch;
    // enough row-splits to give every SM a few CTAs, at least 256 rows each
    const long long tiles = cdiv(d.n, 128) * cdiv(d.k, 128);
    long long want = (4LL * sm_count() + tiles - 1) / tiles;           // CTAs along z
    long long per_batch = (want + d.batch - 1) / d.batch;
    if (per_batch < 1) per_batch = 1;
    int rows_per_cta = (int)cdiv(d.m, per_batch);
    if (rows_per_cta < 256) rows_per_cta = 256;
    rows_per_cta = (rows_per_cta + 15) / 16 * 16;
    p.rows_per_cta = rows_per_cta;
    const long long gz = (long long)d.batch * cdiv(d.m, rows_per_cta);
    CUM_REQUIRE(gz <= 65535, "wgrad: grid.z=%lld too large", gz);
    dim3 grid((unsigned)cdiv(d.n, 128), (unsigned)cdiv(d.k, 128), (unsigned)gz);
    wgrad_simt_kernel<<<grid, 256, 0, st>>>(p);
    CUM_LAUNCH_CHECK("wgrad_simt_kernel");
    return CUM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// ln_bwd: x = residual stream value that was normalised (saved), y = LN(x)*gamma+beta.
//   dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)) (+ dres_in: gradient already flowing in the residual stream)
//   dgamma += sum_rows dy*xhat ; dbeta += sum_rows dy.   One warp per row (row in registers), CTA-level column
//   partials in shared memory -> one atomicAdd per column per CTA.
// ---------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                      const float* __restrict__ dres_in, const float* __restrict__ gamma,
                                                      float* __restrict__ dx, float* __restrict__ dgamma,
                                                      float* __restrict__ dbeta, float eps, long long rows, int c,
                                                      int c_pad, int rows_per_cta) {
    extern __shared__ float part[];          // [2][c_pad]
    for (int i = threadIdx.x; i < 2 * c_pad; i += blockDim.x) part[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c4n = c_pad >> 2;
    float4 gsum[NCH], bsum[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) gsum[i] = bsum[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long rbeg = (long long)blockIdx.x * rows_per_cta;
    const long long rend = rbeg + rows_per_cta < rows ? rbeg + rows_per_cta : rows;
    for (long long row = rbeg + wid; row < rend; row += 8) {
        float4 xv[NCH], dv[NCH];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c4 = lane + 32 * i;
            xv[i] = dv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c4 < c4n) {
                xv[i] = reinterpret_cast<const float4*>(x + row * c_pad)[c4];
                dv[i] = reinterpret_cast<const float4*>(dy + row * c_pad)[c4];
                s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
            }
        }
        const float mean = warp_sum(s) / (float)c;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int cb = (lane + 32 * i) * 4;
            xv[i].x = cb + 0 < c ? xv[i].x - mean : 0.f; xv[i].y = cb + 1 < c ? xv[i].y - mean : 0.f;
            xv[i].z = cb + 2 < c ? xv[i].z - mean : 0.f; xv[i].w = cb + 3 < c ? xv[i].w - mean : 0.f;
            q += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)c + eps);
        float m1 = 0.f, m2 = 0.f;          // sum g*dy, sum g*dy*xhat
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c4 = lane + 32 * i;
            if (c4 < c4n) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
                xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;      // xhat
                gsum[i].x += dv[i].x * xv[i].x; gsum[i].y += dv[i].y * xv[i].y; gsum[i].z += dv[i].z * xv[i].z; gsum[i].w += dv[i].w * xv[i].w;
                bsum[i].x += dv[i].x; bsum[i].y += dv[i].y; bsum[i].z += dv[i].z; bsum[i].w += dv[i].w;
                dv[i].x *= g.x; dv[i].y *= g.y; dv[i].z *= g.z; dv[i].w *= g.w;          // g*dy
                m1 += (dv[i].x + dv[i].y) + (dv[i].z + dv[i].w);
                m2 += (dv[i].x * xv[i].x + dv[i].y * xv[i].y) + (dv[i].z * xv[i].z + dv[i].w * xv[i].w);
            }
        }
        m1 = warp_sum(m1) / (float)c;
        m2 = warp_sum(m2) / (float)c;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c4 = lane + 32 * i;
            if (c4 < c4n) {
                const int cb = c4 * 4;
                float4 o;
                o.x = cb + 0 < c ? rstd * (dv[i].x - m1 - xv[i].x * m2) : 0.f;
                o.y = cb + 1 < c ? rstd * (dv[i].y - m1 - xv[i].y * m2) : 0.f;
                o.z = cb + 2 < c ? rstd * (dv[i].z - m1 - xv[i].z * m2) : 0.f;
                o.w = cb + 3 < c ? rstd * (dv[i].w - m1 - xv[i].w * m2) : 0.f;
                if (dres_in) {
                    const float4 r = reinterpret_cast<const float4*>(dres_in + row * c_pad)[c4];
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                reinterpret_cast<float4*>(dx + row * c_pad)[c4] = o;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c4 = lane + 32 * i;
        if (c4 < c4n) {
            atomicAdd(&part[c4 * 4 + 0], gsum[i].x); atomicAdd(&part[c4 * 4 + 1], gsum[i].y);
            atomicAdd(&part[c4 * 4 + 2], gsum[i].z); atomicAdd(&part[c4 * 4 + 3], gsum[i].w);
            atomicAdd(&part[c_pad + c4 * 4 + 0], bsum[i].x); atomicAdd(&part[c_pad + c4 * 4 + 1], bsum[i].y);
            atomicAdd(&part[c_pad + c4 * 4 + 2], bsum[i].z); atomicAdd(&part[c_pad + c4 * 4 + 3], bsum[i].w);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < c_pad; i += blockDim.x) {
        atomicAdd(dgamma + i, part[i]);
        atomicAdd(dbeta + i, part[c_pad + i]);
    }
}

int ln_bwd(const float* x, const float* dy, const float* dres_in, const float* gamma, float* dx, float* dgamma,
           float* dbeta, float eps, long long rows, int c, int c_pad, cudaStream_t st) {
    CUM_REQUIRE(x && dy && gamma && dx && dgamma && dbeta, "ln_bwd: null pointer");
    CUM_REQUIRE(rows > 0 && c > 0 && c <= c_pad && c_pad % 4 == 0 && c_pad <= 1024, "ln_bwd: bad shape");
    const int rows_per_cta = 64;
    const unsigned grid = (unsigned)cdiv(rows, rows_per_cta);
    const size_t smem = 2 * (size_t)c_pad * sizeof(float);
    const int c4n = c_pad / 4;
    if (c4n <= 32)       ln_bwd_kernel<1><<<grid, 256, smem, st>>>(x, dy, dres_in, gamma, dx, dgamma, dbeta, eps, rows, c, c_pad, rows_per_cta);
    else if (c4n <= 64)  ln_bwd_kernel<2><<<grid, 256, smem, st>>>(x, dy, dres_in, gamma, dx, dgamma, dbeta, eps, rows, c, c_pad, rows_per_cta);
    else if (c4n <= 128) ln_bwd_kernel<4><<<grid, 256, smem, st>>>(x, dy, dres_in, gamma, dx, dgamma, dbeta, eps, rows, c, c_pad, rows_per_cta);
    else                 ln_bwd_kernel<8><<<grid, 256, smem, st>>>(x, dy, dres_in, gamma, dx, dgamma, dbeta, eps, rows, c, c_pad, rows_per_cta);
    CUM_LAUNCH_CHECK("ln_bwd_kernel");
    return CUM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// dwconv_silu_bwd (width 4): p[t] = b + sum_k w[k] x[t-3+k], y = silu(p).
//   dp = dy * silu'(p);  dx[t] = sum_k w[k] dp[t+3-k];  dw[k] += sum_t dp[t] x[t-3+k];  db += sum_t dp[t]
// Thread = 4 channels x DWB_T steps; p is recomputed from x (rolling window), dp for the 3 steps after the tile is
// recomputed too so dx needs no second pass.  Zero initial state (training).
// ---------------------------------------------------------------------------------------------------------
constexpr int DWB_T = 32;

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float silu_grad(float p) {
    const float s = sigmoidf_(p);
    return s * (1.f + p * (1.f - s));
}

__global__ void __launch_bounds__(128) dwconv_silu_bwd_kernel(const float* __restrict__ x, long long x_bs, long long x_rs,
                                                               const float* __restrict__ w, const float* __restrict__ bias,
                                                               const float* __restrict__ dy, float* __restrict__ dx,
                                                               long long dx_bs, long long dx_rs, float* __restrict__ dw,
                                                               float* __restrict__ db, int len, int d_pad) {
    const int c4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (c4 >= (d_pad >> 2)) return;
    const int b = blockIdx.z;
    const int t0 = blockIdx.y * DWB_T;
    const float* xb = x + (long long)b * x_bs;
    const float* dyb = dy + (long long)b * len * d_pad;
    float4 wv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) wv[k] = __ldg(reinterpret_cast<const float4*>(w + (long long)k * d_pad) + c4);
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + c4);
    auto ldx = [&](int t) { return (t >= 0 && t < len) ? *reinterpret_cast<const float4*>(xb + (long long)t * x_rs + c4 * 4) : f4_zero(); };
    // window of x: xw[j] = x[t - 3 + j]; window of dp: dpw[j] = dp[t + j] for the dx taps
    float4 xw[4], dpw[4], dwa[4], dba = f4_zero();
#pragma unroll
    for (int k = 0; k < 4; ++k) dwa[k] = f4_zero();
    // dp(t) helper needs x[t-3..t]
    auto dp_at = [&](int t, const float4* xwin) {
        if (t >= len) return f4_zero();
        float4 p = bv;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            p.x = fmaf(wv[k].x, xwin[k].x, p.x); p.y = fmaf(wv[k].y, xwin[k].y, p.y);
            p.z = fmaf(wv[k].z, xwin[k].z, p.z); p.w = fmaf(wv[k].w, xwin[k].w, p.w);
        }
        const float4 g = *reinterpret_cast<const float4*>(dyb + (long long)t * d_pad + c4 * 4);
        return make_float4(g.x * silu_grad(p.x), g.y * silu_grad(p.y), g.z * silu_grad(p.z), g.w * silu_grad(p.w));
    };
    // prime: dp[t0], dp[t0+1], dp[t0+2] and the x window ending at t0+2
#pragma unroll
    for (int j = 0; j < 4; ++j) xw[j] = ldx(t0 - 3 + j);       // x[t0-3 .. t0]
    const int tend = min(t0 + DWB_T, len);
    // rolling: at step t we own dp[t], dp[t+1], dp[t+2], dp[t+3]
    float4 xq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xq[j] = xw[j];
    dpw[0] = dp_at(t0, xq);
#pragma unroll
    for (int a = 1; a < 4; ++a) {
        xq[0] = xq[1]; xq[1] = xq[2]; xq[2] = xq[3]; xq[3] = ldx(t0 + a);
        dpw[a] = dp_at(t0 + a, xq);
    }
    // xq now holds x[t0 .. t0+3]; xw holds x[t0-3 .. t0]
    for (int t = t0; t < tend; ++t) {
        // weight / bias gradients use dp[t] (only the tile's own steps, so nothing is double counted)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dwa[k].x = fmaf(dpw[0].x, xw[k].x, dwa[k].x); dwa[k].y = fmaf(dpw[0].y, xw[k].y, dwa[k].y);
            dwa[k].z = fmaf(dpw[0].z, xw[k].z, dwa[k].z); dwa[k].w = fmaf(dpw[0].w, xw[k].w, dwa[k].w);
        }
        dba.x += dpw[0].x; dba.y += dpw[0].y; dba.z += dpw[0].z; dba.w += dpw[0].w;
        // dx[t] = sum_k w[k] dp[t + 3 - k]
        float4 o = f4_zero();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o.x = fmaf(wv[k].x, dpw[3 - k].x, o.x); o.y = fmaf(wv[k].y, dpw[3 - k].y, o.y);
            o.z = fmaf(wv[k].z, dpw[3 - k].z, o.z); o.w = fmaf(wv[k].w, dpw[3 - k].w, o.w);
        }
        *reinterpret_cast<float4*>(dx + (long long)b * dx_bs + (long long)t * dx_rs + c4 * 4) = o;
        // slide: x windows and dp window
        xw[0] = xw[1]; xw[1] = xw[2]; xw[2] = xw[3]; xw[3] = ldx(t + 1);
        xq[0] = xq[1]; xq[1] = xq[2]; xq[2] = xq[3]; xq[3] = ldx(t + 4);
        dpw[0] = dpw[1]; dpw[1] = dpw[2]; dpw[2] = dpw[3]; dpw[3] = dp_at(t + 4, xq);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        atomicAdd(dw + (long long)k * d_pad + c4 * 4 + 0, dwa[k].x); atomicAdd(dw + (long long)k * d_pad + c4 * 4 + 1, dwa[k].y);
        atomicAdd(dw + (long long)k * d_pad + c4 * 4 + 2, dwa[k].z); atomicAdd(dw + (long long)k * d_pad + c4 * 4 + 3, dwa[k].w);
    }
    atomicAdd(db + c4 * 4 + 0, dba.x); atomicAdd(db + c4 * 4 + 1, dba.y);
    atomicAdd(db + c4 * 4 + 2, dba.z); atomicAdd(db + c4 * 4 + 3, dba.w);
}

int dwconv_silu_bwd(const float* x, long long x_bs, long long x_rs, const float* w, const float* bias, const float* dy,
                    float* dx, long long dx_bs, long long dx_rs, float* dw, float* db, int batch, int len, int d_pad,
                    int width, cudaStream_t st) {
    CUM_REQUIRE(x && w && bias && dy && dx && dw && db, "dwconv_silu_bwd: null pointer");
    CUM_REQUIRE(width == 4, "dwconv_silu_bwd: width 4 only");
    CUM_REQUIRE(batch > 0 && batch <= 65535 && len > 0 && d_pad % 4 == 0 && x_rs % 4 == 0 && dx_rs % 4 == 0, "dwconv_silu_bwd: bad shape");
    dim3 grid((unsigned)cdiv(d_pad / 4, 128), (unsigned)cdiv(len, DWB_T), (unsigned)batch);
    dwconv_silu_bwd_kernel<<<grid, 128, 0, st>>>(x, x_bs, x_rs, w, bias, dy, dx, dx_bs, dx_rs, dw, db, len, d_pad);
    CUM_LAUNCH_CHECK("dwconv_silu_bwd_kernel");
    return CUM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// conv_in_bwd: y = relu(b + sum_k w[k,c] x[S t + k]) -> given dy and y:  dw[k,c] += sum dz x[S t+k], db[c] += sum dz,
// dz = dy * (y > 0).  (No data gradient: the waveform is not a parameter.)
// CTA = 256 rows; thread = one float4 channel group x row slots; one atomic per (k, c) per CTA.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_in_bwd_kernel(const float* __restrict__ x, long long x_stride, int length,
                                                           const float* __restrict__ y, const float* __restrict__ dy,
                                                           float* __restrict__ dw, float* __restrict__ db, int rows_out,
                                                           int c_pad, int kernel, int stride, int rows_per_cta) {
    const int groups = c_pad >> 2;
    const int gpr = groups < 256 ? groups : 256;
    const int slots = 256 / gpr;
    const int tg = threadIdx.x % gpr, ts = threadIdx.x / gpr;
    __shared__ float4 red[256];
    const bool active = ts < slots;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * rows_per_cta, t1 = min(rows_out, t0 + rows_per_cta);
    const float* xb = x + (long long)b * x_stride;
    for (int g = tg; g < groups; g += gpr) {          // (a single iteration whenever slots > 1)
        float4 acc[CI_MAXK], accb = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < CI_MAXK; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active)
        for (int t = t0 + ts; t < t1; t += slots) {
            const long long off = ((long long)b * rows_out + t) * c_pad;
            const float4 yv = reinterpret_cast<const float4*>(y + off)[g];
            float4 d = reinterpret_cast<const float4*>(dy + off)[g];
            d.x = yv.x > 0.f ? d.x : 0.f; d.y = yv.y > 0.f ? d.y : 0.f; d.z = yv.z > 0.f ? d.z : 0.f; d.w = yv.w > 0.f ? d.w : 0.f;
            accb.x += d.x; accb.y += d.y; accb.z += d.z; accb.w += d.w;
#pragma unroll
            for (int k = 0; k < CI_MAXK; ++k) {
                if (k < kernel) {
                    const long long si = (long long)t * stride + k;
                    const float xv = si < length ? __ldg(xb + si) : 0.f;
                    acc[k].x = fmaf(d.x, xv, acc[k].x); acc[k].y = fmaf(d.y, xv, acc[k].y);
                    acc[k].z = fmaf(d.z, xv, acc[k].z); acc[k].w = fmaf(d.w, xv, acc[k].w);
                }
            }
        }
        // one atomic per column and CTA (the row slots' partial sums meet in shared memory first)
#pragma unroll
        for (int k = 0; k < CI_MAXK; ++k) {
            if (k < kernel) {       // uniform
                if (cta_slot_reduce(acc[k], red, slots, gpr, tg, ts) && active) {
                    float* p = dw + (long long)k * c_pad + 4 * g;
                    atomicAdd(p + 0, acc[k].x); atomicAdd(p + 1, acc[k].y); atomicAdd(p + 2, acc[k].z); atomicAdd(p + 3, acc[k].w);
                }
            }
        }
        if (cta_slot_reduce(accb, red, slots, gpr, tg, ts) && active) {
            atomicAdd(db + 4 * g + 0, accb.x); atomicAdd(db + 4 * g + 1, accb.y);
            atomicAdd(db + 4 * g + 2, accb.z); atomicAdd(db + 4 * g + 3, accb.w);
        }
    }
}

int conv_in_bwd(const float* x, long long x_stride, int batch, int length, const float* y, const float* dy, float* dw,
                float* db, int rows_out, int c_pad, int kernel, int stride, cudaStream_t st) {
    CUM_REQUIRE(x && y && dy && dw && db, "conv_in_bwd: null pointer");
    CUM_REQUIRE(batch > 0 && batch <= 65535 && rows_out > 0 && c_pad % 4 == 0 && kernel >= 1 && kernel <= CI_MAXK, "conv_in_bwd: bad shape");
    long long rpc = cdiv((long long)rows_out * batch, 8LL * sm_count());
    if (rpc < 256) rpc = 256;
    if (rpc > rows_out) rpc = rows_out;
    dim3 grid((unsigned)cdiv(rows_out, rpc), (unsigned)batch);
    conv_in_bwd_kernel<<<grid, 256, 0, st>>>(x, x_stride, length, y, dy, dw, db, rows_out, c_pad, kernel, stride, (int)rpc);
    CUM_LAUNCH_CHECK("conv_in_bwd_kernel");
    return CUM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// convt_out_bwd: out[b,m] = scale[b] * (bias + sum_{j,k: S j + k = m} <g[b,j,:], w[k,:]>), m < length.
//   e[m] = scale[b] * dout[b,m] (0 beyond length);  dg[b,j,c] = sum_k e[S j + k] w[k,c];
//   dw[k,c] += sum_{b,j} g[b,j,c] e[S j + k];  dbias += sum e.
// CTA = 64 input rows; thread = float4 channel group x row slots (like conv_in_bwd).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) convt_out_bwd_kernel(const float* __restrict__ g, int rows_in, int c_pad,
                                                             const float* __restrict__ w, const float* __restrict__ scale,
                                                             const float* __restrict__ dout, long long dout_stride,
                                                             int length, float* __restrict__ dg, float* __restrict__ dw,
                                                             float* __restrict__ dbias, int kernel, int stride,
                                                             int blocks_per_cta) {
    extern __shared__ float es[];           // e[S*j0 .. S*(j0+64) + kernel)
    const int b = blockIdx.y;
    const float sc = scale ? scale[b] : 1.0f;
    const int ne = 64 * stride + kernel;
    const int groups = c_pad >> 2;
    const int gpr = groups < 256 ? groups : 256;
    const int slots = 256 / gpr;
    const int tg = threadIdx.x % gpr, ts = threadIdx.x / gpr;
    const int nblocks = (rows_in + 63) / 64;
    const int blk0 = blockIdx.x * blocks_per_cta, blk1 = min(nblocks, blk0 + blocks_per_cta);
    float esum = 0.f;
    // this kernel supports c_pad <= 1024 (one column group per thread) so the dw accumulators stay in registers
    float4 wv[CT_MAXK], acc[CT_MAXK];
#pragma unroll
    for (int k = 0; k < CT_MAXK; ++k) {
        wv[k] = (k < kernel && tg < groups && ts < slots) ? __ldg(reinterpret_cast<const float4*>(w + (long long)k * c_pad) + tg)
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int blk = blk0; blk < blk1; ++blk) {
        const int j0 = blk * 64, j1 = min(rows_in, j0 + 64);
        __syncthreads();
        for (int i = threadIdx.x; i < ne; i += blockDim.x) {
            const long long m = (long long)j0 * stride + i;
            const float v = m < length ? sc * dout[(long long)b * dout_stride + m] : 0.f;
            es[i] = v;
            // every output sample belongs to exactly one block's first 64*S window (the last block also owns the K-S tail)
            if (i < 64 * stride || blk == nblocks - 1) esum += v;
        }
        __syncthreads();
        if (ts < slots && tg < groups) {
            for (int j = j0 + ts; j < j1; j += slots) {
                const long long off = ((long long)b * rows_in + j) * c_pad;
                const float4 gv = reinterpret_cast<const float4*>(g + off)[tg];
                float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < CT_MAXK; ++k) {
                    if (k < kernel) {
                        const float e = es[(j - j0) * stride + k];
                        o.x = fmaf(e, wv[k].x, o.x); o.y = fmaf(e, wv[k].y, o.y); o.z = fmaf(e, wv[k].z, o.z); o.w = fmaf(e, wv[k].w, o.w);
                        acc[k].x = fmaf(e, gv.x, acc[k].x); acc[k].y = fmaf(e, gv.y, acc[k].y);
                        acc[k].z = fmaf(e, gv.z, acc[k].z); acc[k].w = fmaf(e, gv.w, acc[k].w);
                    }
                }
                reinterpret_cast<float4*>(dg + off)[tg] = o;
            }
        }
    }
    // CTA totals first (shared memory), then one atomic per address and CTA
    __shared__ float4 red[256];
    __shared__ float esum_w[8];
    esum = warp_sum(esum);
    if ((threadIdx.x & 31) == 0) esum_w[threadIdx.x >> 5] = esum;
    __syncthreads();
    if (threadIdx.x == 0 && dbias) {
        float e = 0.f;
        for (int i = 0; i < 8; ++i) e += esum_w[i];
        atomicAdd(dbias, e);
    }
    const bool active = ts < slots && tg < groups;
#pragma unroll
    for (int k = 0; k < CT_MAXK; ++k) {
        if (k < kernel) {           // uniform
            if (cta_slot_reduce(acc[k], red, slots, gpr, tg, ts) && active) {
                float* p = dw + (long long)k * c_pad + 4 * tg;
                atomicAdd(p + 0, acc[k].x); atomicAdd(p + 1, acc[k].y); atomicAdd(p + 2, acc[k].z); atomicAdd(p + 3, acc[k].w);
            }
        }
    }
}

int convt_out_bwd(const float* g, int batch, int rows_in, int c_pad, const float* w, const float* scale,
                  const float* dout, long long dout_stride, int length, float* dg, float* dw, float* dbias, int kernel,
                  int stride, cudaStream_t st) {
    CUM_REQUIRE(g && w && dout && dg && dw, "convt_out_bwd: null pointer");
    CUM_REQUIRE(batch > 0 && batch <= 65535 && rows_in > 0 && c_pad % 4 == 0 && kernel >= 1 && kernel <= CT_MAXK && stride >= 1, "convt_out_bwd: bad shape");
    CUM_REQUIRE(c_pad <= 1024, "convt_out_bwd: c_pad=%d > 1024 not supported", c_pad);
    const int nblocks = (rows_in + 63) / 64;
    long long bpc = cdiv((long long)nblocks * batch, 8LL * sm_count());      // 64-row blocks per CTA
    if (bpc < 1) bpc = 1;
    dim3 grid((unsigned)cdiv(nblocks, bpc), (unsigned)batch);
    const size_t smem = (size_t)(64 * stride + kernel) * sizeof(float);
    convt_out_bwd_kernel<<<grid, 256, smem, st>>>(g, rows_in, c_pad, w, scale, dout, dout_stride, length, dg, dw, dbias, kernel, stride, (int)bpc);
    CUM_LAUNCH_CHECK("convt_out_bwd_kernel");
    return CUM_OK;
}

}  // namespace cum
