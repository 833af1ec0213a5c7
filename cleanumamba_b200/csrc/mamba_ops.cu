// Mamba-block operators in channels-last layout: residual add + LayerNorm, depthwise causal conv + SiLU,
// selective scan (fwd) with optional carried state.  All HBM / MUFU-bound CUDA-core kernels.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda_fp16.h>
#include <stdlib.h>

#ifndef SCAN_T_UNROLL
#define SCAN_T_UNROLL 1   // time-loop unroll of the scan recurrence (2 trades occupancy for ILP)
#endif
constexpr int kScanUnroll = SCAN_T_UNROLL;

namespace cum {

// ---------------------------------------------------------------------------------------------------------
// ln_residual: r = h (+ residual_in); normed = LayerNorm_c(r) * gamma + beta.   One warp per row, the row lives in
// registers (NCH float4 per lane), two-pass mean / variance over the `c` real channels (pad lanes stay 0).
// Reference: mamba_ssm Block.forward (non-fused add+norm, fp32 residual) and CleanUMamba.py:292-294.
// ---------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(256) ln_residual_kernel(const float* __restrict__ h, const float* __restrict__ rin,
                                                           float* __restrict__ rout, float* __restrict__ normed,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, long long rows,
                                                           int c, int c_pad) {
    pdl_trigger();
    pdl_wait();          // PDL: nothing of the previous kernel is touched before this point

    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int c4n = c_pad >> 2;
    const float4* hp = reinterpret_cast<const float4*>(h + row * c_pad);
    const float4* rp = rin ? reinterpret_cast<const float4*>(rin + row * c_pad) : nullptr;
    float4 v[NCH];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c4 = lane + 32 * i;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 < c4n) {
            v[i] = hp[c4];
            if (rp) {
                const float4 r = rp[c4];
                v[i].x += r.x; v[i].y += r.y; v[i].z += r.z; v[i].w += r.w;
            }
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    const float mean = warp_sum(s) / (float)c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int cb = (lane + 32 * i) * 4;
        const float dx = cb + 0 < c ? v[i].x - mean : 0.f, dy = cb + 1 < c ? v[i].y - mean : 0.f;
        const float dz = cb + 2 < c ? v[i].z - mean : 0.f, dw = cb + 3 < c ? v[i].w - mean : 0.f;
        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)c + eps);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int c4 = lane + 32 * i;
        if (c4 < c4n) {
            if (rout) reinterpret_cast<float4*>(rout + row * c_pad)[c4] = v[i];
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
            const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c4);
            const int cb = c4 * 4;
            float4 o;
            o.x = cb + 0 < c ? (v[i].x - mean) * rstd * g.x + bt.x : 0.f;
            o.y = cb + 1 < c ? (v[i].y - mean) * rstd * g.y + bt.y : 0.f;
            o.z = cb + 2 < c ? (v[i].z - mean) * rstd * g.z + bt.z : 0.f;
            o.w = cb + 3 < c ? (v[i].w - mean) * rstd * g.w + bt.w : 0.f;
            reinterpret_cast<float4*>(normed + row * c_pad)[c4] = o;
        }
    }
}

int ln_residual_fwd(const float* h, const float* rin, float* rout, float* normed, const float* gamma,
                    const float* beta, float eps, long long rows, int c, int c_pad, cudaStream_t st) {
    CUM_REQUIRE(h && normed && gamma && beta, "ln_residual: null pointer");
    CUM_REQUIRE(rows > 0 && c > 0 && c <= c_pad && c_pad % 4 == 0, "ln_residual: bad shape rows=%lld c=%d c_pad=%d", rows, c, c_pad);
    CUM_REQUIRE(c_pad <= 1024, "ln_residual: c_pad=%d > 1024 not supported", c_pad);
    CUM_REQUIRE(aligned16(h) && aligned16(normed) && aligned16(gamma) && aligned16(beta) && (!rin || aligned16(rin)) && (!rout || aligned16(rout)),
                "ln_residual: pointers must be 16-byte aligned");
    const unsigned grid = (unsigned)cdiv(rows, 8);
    const int c4n = c_pad / 4;
    cudaError_t e;
    if (c4n <= 32)        e = launch_kernel(ln_residual_kernel<1>, dim3(grid), dim3(256), 0, st, h, rin, rout, normed, gamma, beta, eps, rows, c, c_pad);
    else if (c4n <= 64)   e = launch_kernel(ln_residual_kernel<2>, dim3(grid), dim3(256), 0, st, h, rin, rout, normed, gamma, beta, eps, rows, c, c_pad);
    else if (c4n <= 128)  e = launch_kernel(ln_residual_kernel<4>, dim3(grid), dim3(256), 0, st, h, rin, rout, normed, gamma, beta, eps, rows, c, c_pad);
    else                  e = launch_kernel(ln_residual_kernel<8>, dim3(grid), dim3(256), 0, st, h, rin, rout, normed, gamma, beta, eps, rows, c, c_pad);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(ln_residual_kernel)");
    return CUM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// dwconv_silu: y[b,t,c] = silu(bias[c] + sum_k w[k,c] * x[b, t-(W-1)+k, c]);  inputs before t=0 come from
// conv_state (or are 0).  Thread = 4 channels x DW_T consecutive steps with a rolling register window, so x is
// read once (plus a W-1 halo per tile) and all accesses are coalesced float4 over channels.
// Reference: causal_conv1d_fn(..., "silu") == act(conv1d(x)[..., :L]) in Mamba.forward; Mamba.step's roll+sum.
// ---------------------------------------------------------------------------------------------------------
constexpr int DW_T = 16;
constexpr int DW_MAXW = 4;

__device__ __forceinline__ float4 f4_fma(float4 w, float4 x, float4 a) {
    return make_float4(fmaf(w.x, x.x, a.x), fmaf(w.y, x.y, a.y), fmaf(w.z, x.z, a.z), fmaf(w.w, x.w, a.w));
}

__global__ void __launch_bounds__(128) dwconv_silu_kernel(const float* __restrict__ x, long long x_bs, long long x_rs,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           float* __restrict__ y, long long y_bs, long long y_rs,
                                                           const float* __restrict__ state, float* __restrict__ state_out,
                                                           int len, int d_pad, int width) {
    pdl_trigger();
    pdl_wait();          // PDL: nothing of the previous kernel is touched before this point

    const int c4 = blockIdx.x * blockDim.x + threadIdx.x;
    if (c4 >= (d_pad >> 2)) return;
    const int b = blockIdx.z;
    const int t0 = blockIdx.y * DW_T;
    const float* xb = x + (long long)b * x_bs;
    float4 wv[DW_MAXW];
#pragma unroll
    for (int k = 0; k < DW_MAXW; ++k)
        wv[k] = k < width ? __ldg(reinterpret_cast<const float4*>(w + (long long)k * d_pad) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + c4);
    // window[j] holds x[t - (W-1) + j]; slots are right-aligned in a DW_MAXW-wide window
    float4 win[DW_MAXW];
#pragma unroll
    for (int j = 0; j < DW_MAXW - 1; ++j) {
        const int back = (DW_MAXW - 1) - j;  // this slot is x[t0 - back]
        win[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (back <= width - 1) {
            const int t = t0 - back;
            if (t >= 0) win[j] = *reinterpret_cast<const float4*>(xb + (long long)t * x_rs + c4 * 4);
            else if (state) win[j] = *reinterpret_cast<const float4*>(state + ((long long)b * (width - 1) + (width - 1 + t)) * d_pad + c4 * 4);
        }
    }
    const int tend = min(t0 + DW_T, len);
    // all rows of the tile are requested before the first one is used (one dependent load per step left the kernel latency-bound
    // at 57 % of the HBM roofline)
    constexpr int DW_B = 8;
    for (int tb = t0; tb < tend; tb += DW_B) {
        float4 xin[DW_B];
#pragma unroll
        for (int i = 0; i < DW_B; ++i)
            xin[i] = (tb + i < tend) ? *reinterpret_cast<const float4*>(xb + (long long)(tb + i) * x_rs + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < DW_B; ++i) {
            if (tb + i >= tend) break;
            win[DW_MAXW - 1] = xin[i];
            float4 acc = bv;
#pragma unroll
            for (int k = 0; k < DW_MAXW; ++k) {
                acc = f4_fma(wv[k], win[k], acc);  // tap k multiplies x[t - (W-1) + k]
            }
            float4 o = make_float4(siluf_(acc.x), siluf_(acc.y), siluf_(acc.z), siluf_(acc.w));
            *reinterpret_cast<float4*>(y + (long long)b * y_bs + (long long)(tb + i) * y_rs + c4 * 4) = o;
#pragma unroll
            for (int j = 0; j < DW_MAXW - 1; ++j) win[j] = win[j + 1];
        }
    }
    // single-tile problems (streaming: a few tokens per call) update the carried state here: the window now holds the last
    // W-1 inputs (older entries from the previous state when len < W-1); this thread is the only reader of these state entries
    if (state_out) {
#pragma unroll
        for (int j = 0; j < DW_MAXW - 1; ++j)
            *reinterpret_cast<float4*>(state_out + ((long long)b * (DW_MAXW - 1) + j) * d_pad + c4 * 4) = win[j];
    }
}

// new conv_state = last (W-1) inputs (older entries come from the previous state when len < W-1)
__global__ void dwconv_state_kernel(const float* __restrict__ x, long long x_bs, long long x_rs,
                                    const float* __restrict__ state, float* __restrict__ state_out, int len, int d_pad,
                                    int width) {
    pdl_trigger();
    pdl_wait();          // PDL: nothing of the previous kernel is touched before this point

    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d_pad) return;
    const int b = blockIdx.y;
    float v[DW_MAXW - 1];
#pragma unroll
    for (int j = 0; j < DW_MAXW - 1; ++j) {
        v[j] = 0.f;
        if (j < width - 1) {
            const int t = len - (width - 1) + j;
            if (t >= 0) v[j] = x[(long long)b * x_bs + (long long)t * x_rs + c];
            else if (state) v[j] = state[((long long)b * (width - 1) + (width - 1 + t)) * d_pad + c];
        }
    }
#pragma unroll
    for (int j = 0; j < DW_MAXW - 1; ++j)
        if (j < width - 1) state_out[((long long)b * (width - 1) + j) * d_pad + c] = v[j];
}

int dwconv_silu_fwd(const float* x, long long x_bs, long long x_rs, const float* w, const float* bias, float* y,
                    const float* conv_state, float* conv_state_out, int batch, int len, int d_pad, int width,
                    cudaStream_t st, long long y_bs, long long y_rs) {
    if (y_bs == 0 && y_rs == 0) { y_bs = (long long)len * d_pad; y_rs = d_pad; }
    CUM_REQUIRE(y_bs % 4 == 0 && y_rs % 4 == 0, "dwconv_silu: output strides must be multiples of 4 elements");
    CUM_REQUIRE(x && w && bias && y, "dwconv_silu: null pointer");
    CUM_REQUIRE(batch > 0 && len > 0 && d_pad > 0 && d_pad % 4 == 0, "dwconv_silu: bad shape");
    CUM_REQUIRE(width >= 1 && width <= DW_MAXW, "dwconv_silu: width=%d unsupported (1..4)", width);
    CUM_REQUIRE(x_rs % 4 == 0 && x_bs % 4 == 0 && aligned16(x) && aligned16(w) && aligned16(bias) && aligned16(y),
                "dwconv_silu: strides/pointers must be 16-byte aligned");
    CUM_REQUIRE(batch <= 65535, "dwconv_silu: batch too large");
    dim3 grid((unsigned)cdiv(d_pad / 4, 128), (unsigned)cdiv(len, DW_T), (unsigned)batch);
    CUM_REQUIRE(width == DW_MAXW, "dwconv_silu: only width 4 is instantiated (d_conv=4, CleanUMamba.py:143)");
    const bool fused_state = conv_state_out && len <= DW_T;      // one time tile per (stream, channel): the kernel writes the new state itself
    cudaError_t e = launch_kernel(dwconv_silu_kernel, grid, dim3(128), 0, st, x, x_bs, x_rs, w, bias, y, y_bs, y_rs, conv_state,
                                  fused_state ? conv_state_out : (float*)nullptr, len, d_pad, width);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(dwconv_silu_kernel)");
    CUM_LAUNCH_CHECK("dwconv_silu_kernel");
    if (conv_state_out && !fused_state) {
        dim3 g2((unsigned)cdiv(d_pad, 128), (unsigned)batch);
        e = launch_kernel(dwconv_state_kernel, g2, dim3(128), 0, st, x, x_bs, x_rs, conv_state, conv_state_out, len, d_pad, width);
        if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(dwconv_state_kernel)");
        CUM_LAUNCH_CHECK("dwconv_state_kernel");
    }
    return CUM_OK;
}

// ---------------------------------------------------------------------------------------------------------
// selective_scan (forward).  Reference arithmetic: mamba_ssm selective_scan_ref / selective_scan_fwd_kernel.
//
// CTA = CH channels x SL state-slices (threads = CH*SL; a warp = 32 channels of one slice, so B_t / C_t reads
// are warp-wide shared-memory broadcasts and u / delta reads are conflict-free).  Time is processed in chunks of
// TC steps, double-buffered through shared memory with cp.async:
//   stage      u, delta (TC x CH) and B, C (TC x NP) tiles   (cp.async, next chunk in flight during compute)
//   discretise delta <- softplus(delta + bias) once per (t, channel), in place
//   recurrence every thread keeps NS states of one channel in registers: h = exp2(dl*A2) h + (dl u) B_t ;
//              partial y = sum_i h_i C_t,i  -> ypart[slice][t][ch]         (carry h stays in registers across chunks)
//   combine    y = (sum_slices ypart + D u) * silu(z)  -> coalesced store
// A2 = -exp(A_log) * log2(e) is prepared at weight-pack time so the decay is a single ex2.approx.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ float softplusf_(float x) {
    // torch softplus (beta=1, threshold=20) = max(x,0) + log1p(exp(-|x|)), evaluated with the fast intrinsics:
    // for e = exp(-|x|) < 1/16 a 5-term series of log1p (|rel err| < 2e-7), else log(1 + e) where 1 + e is exact enough
    if (x > 20.0f) return x;
    const float e = __expf(-fabsf(x));
    const float l = e < 0.0625f ? e * (1.0f - e * (0.5f - e * (0.33333333f - e * (0.25f - e * 0.2f)))) : __logf(1.0f + e);
    return fmaxf(x, 0.0f) + l;
}

template <int NS, int SL, int CH, int TC>
struct ScanSmem {
    static constexpr int NP = NS * SL;
    float u[2][TC][CH];
    float dl[2][TC][CH];
    float z[2][TC][CH];           // staged with u / delta so the combine phase never waits on a global load
    float Bm[2][TC][NP];
    float Cm[2][TC][NP];
    float ypart[SL][TC][CH];
};

// Kernel parameters: the public descriptor plus the SEGMENT-PARALLEL mode used for small batches (selective_scan_fwd below).
// nseg > 1: blockIdx.y = item * nseg + segment; a segment covers rows [segment * seg_len, ...) of its item; carried states
// (h0 / h_out) are indexed by blockIdx.y (one state block per segment); seg_sum receives sum_t softplus(delta_t + bias) of the
// segment per channel (its total decay is exp2(a2 * seg_sum)).
struct ScanK : cum_scan_desc {
    int nseg, seg_len;
    float* seg_sum;
};

template <int NS, int SL, int CH, int TC>
__global__ void __launch_bounds__(CH* SL, (CH * SL >= 256) ? 3 : 4) selective_scan_fwd_kernel(const ScanK p) {
    pdl_trigger();
    pdl_wait();          // PDL: nothing of the previous kernel is touched before this point

    constexpr int NP = NS * SL;
    constexpr int NT = CH * SL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScanSmem<NS, SL, CH, TC>& sm = *reinterpret_cast<ScanSmem<NS, SL, CH, TC>*>(smem_raw);

    const int bb = blockIdx.y;                                   // state block (h0 / h_out / seg_sum) index
    const int seg = p.nseg > 1 ? bb % p.nseg : 0;
    const int b = p.nseg > 1 ? bb / p.nseg : bb;                 // batch item
    const int row0 = seg * p.seg_len;                            // first row of this CTA's segment
    const int len = p.nseg > 1 ? min(p.seg_len, p.len - row0) : p.len;
    const int c0 = blockIdx.x * CH;
    const int tid = threadIdx.x;
    const int ch = tid % CH, slice = tid / CH;
    const int c = c0 + ch;
    const bool c_ok = c < p.d;
    const int nchunks = (len + TC - 1) / TC;

    const float* ub = p.u + (long long)b * p.u_bs + (long long)row0 * p.u_rs;
    const float* db = p.delta + (long long)b * p.dl_bs + (long long)row0 * p.dl_rs;
    const float* Bb = p.Bm + (long long)b * p.B_bs + (long long)row0 * p.B_rs;
    const float* Cb = p.Cm + (long long)b * p.C_bs + (long long)row0 * p.C_rs;
    const float* zb = (p.z && p.y) ? p.z + (long long)b * p.z_bs + (long long)row0 * p.z_rs : nullptr;
    const bool vec_ok = ((p.u_rs | p.dl_rs | p.B_rs | p.C_rs | p.u_bs | p.dl_bs | p.B_bs | p.C_bs) % 4 == 0) &&
                        (!p.z || ((p.z_rs | p.z_bs) % 4 == 0 && ((uintptr_t)p.z & 15) == 0)) &&
                        ((((uintptr_t)p.u | (uintptr_t)p.delta | (uintptr_t)p.Bm | (uintptr_t)p.Cm) & 15) == 0) &&
                        (p.d % 4 == 0) && (p.n_state % 4 == 0);

    auto stage = [&](int chunk, int buf) {
        const int t0 = chunk * TC;
        // u / delta tiles: TC x CH
        for (int i = tid; i < TC * (CH / 4); i += NT) {
            const int t = i / (CH / 4), q = i - t * (CH / 4);
            const int tt = t0 + t, cc = c0 + q * 4;
            float* su = &sm.u[buf][t][q * 4];
            float* sd = &sm.dl[buf][t][q * 4];
            float* sz = &sm.z[buf][t][q * 4];
            if (tt < len && cc + 3 < p.d && vec_ok) {
                cp_async16(su, ub + (long long)tt * p.u_rs + cc);
                cp_async16(sd, db + (long long)tt * p.dl_rs + cc);
                if (zb) cp_async16(sz, zb + (long long)tt * p.z_rs + cc);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool ok = tt < len && cc + e < p.d;
                    su[e] = ok ? ub[(long long)tt * p.u_rs + cc + e] : 0.f;
                    sd[e] = ok ? db[(long long)tt * p.dl_rs + cc + e] : 0.f;
                    if (zb) sz[e] = ok ? zb[(long long)tt * p.z_rs + cc + e] : 0.f;
                }
            }
        }
        // B / C tiles: TC x NP
        for (int i = tid; i < TC * (NP / 4); i += NT) {
            const int t = i / (NP / 4), q = i - t * (NP / 4);
            const int tt = t0 + t, nn = q * 4;
            float* sb = &sm.Bm[buf][t][nn];
            float* sc = &sm.Cm[buf][t][nn];
            if (tt < len && nn + 3 < p.n_state && vec_ok) {
                cp_async16(sb, Bb + (long long)tt * p.B_rs + nn);
                cp_async16(sc, Cb + (long long)tt * p.C_rs + nn);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const bool ok = tt < len && nn + e < p.n_state;
                    sb[e] = ok ? Bb[(long long)tt * p.B_rs + nn + e] : 0.f;
                    sc[e] = ok ? Cb[(long long)tt * p.C_rs + nn + e] : 0.f;
                }
            }
        }
        cp_async_commit();
    };

    // per-thread constants and carried state
    float2 a2p[NS / 2], h2[NS / 2];          // (state 2j, state 2j+1) pairs
    const bool state_vec = (p.n_state % 4 == 0) && (slice * NS + NS <= p.n_state) && c_ok;   // float4 path (64 B / thread)
    // Carried state (streaming): the (CH x NP) state block of this CTA is ONE contiguous 16 KB run of h0 / h_out.  Per-thread
    // float4 accesses touch it with 16 bytes per 256-byte-strided lane (32 LSU wavefronts per instruction, x4 instructions, x2 for
    // load + store: at one token per call that WAS the kernel).  Instead the block moves through shared memory (aliasing ypart,
    // 16-byte chunks XOR-swizzled by the channel) with fully coalesced global accesses.
    static_assert(NS <= TC, "the state staging tile aliases ypart");
    float* const hs = &sm.ypart[0][0][0];
    const bool state_tile = (p.n_state == NP) && (c0 + CH <= p.d) &&
                            (!p.h0 || ((uintptr_t)p.h0 & 15) == 0) && (!p.h_out || ((uintptr_t)p.h_out & 15) == 0);
    auto hs_chunk = [&](int chn, int chunk) { return hs + chn * NP + ((chunk ^ (chn & (NP / 4 - 1))) << 2); };
    if (state_tile) {
#pragma unroll
        for (int q = 0; q < NS / 4; ++q) {
            const float4 av = __ldg(reinterpret_cast<const float4*>(p.a2 + (long long)c * p.n_state + slice * NS) + q);
            a2p[2 * q] = make_float2(av.x, av.y); a2p[2 * q + 1] = make_float2(av.z, av.w);
            h2[2 * q] = make_float2(0.f, 0.f); h2[2 * q + 1] = make_float2(0.f, 0.f);
        }
        if (p.h0) {
            const float4* src = reinterpret_cast<const float4*>(p.h0 + ((long long)bb * p.d + c0) * NP);
            for (int i = tid; i < CH * NP / 4; i += NT) {
                const int chn = i / (NP / 4), chunk = i - chn * (NP / 4);
                *reinterpret_cast<float4*>(hs_chunk(chn, chunk)) = src[i];
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < NS / 4; ++q) {
                const float4 hv = *reinterpret_cast<const float4*>(hs_chunk(ch, slice * (NS / 4) + q));
                h2[2 * q] = make_float2(hv.x, hv.y); h2[2 * q + 1] = make_float2(hv.z, hv.w);
            }
            // (the chunk loop synchronises twice before ypart is written)
        }
    } else if (state_vec) {
#pragma unroll
        for (int q = 0; q < NS / 4; ++q) {
            const float4 av = __ldg(reinterpret_cast<const float4*>(p.a2 + (long long)c * p.n_state + slice * NS) + q);
            a2p[2 * q] = make_float2(av.x, av.y); a2p[2 * q + 1] = make_float2(av.z, av.w);
            float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.h0) hv = *(reinterpret_cast<const float4*>(p.h0 + ((long long)bb * p.d + c) * p.n_state + slice * NS) + q);
            h2[2 * q] = make_float2(hv.x, hv.y); h2[2 * q + 1] = make_float2(hv.z, hv.w);
        }
    } else {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int n = slice * NS + i;
            const bool ok = c_ok && n < p.n_state;
            const float av = ok ? p.a2[(long long)c * p.n_state + n] : 0.f;
            const float hv = (ok && p.h0) ? p.h0[((long long)bb * p.d + c) * p.n_state + n] : 0.f;
            if (i & 1) { a2p[i >> 1].y = av; h2[i >> 1].y = hv; } else { a2p[i >> 1].x = av; h2[i >> 1].x = hv; }
        }
    }

    float seg_dl = 0.f;          // sum of the discretised steps of this segment (segment-parallel mode)
    stage(0, 0);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int buf = chunk & 1;
        const int t0 = chunk * TC;
        const int tn = min(TC, len - t0);
        if (p.h_ckpt) {   // training: state at the start of this chunk, for the reverse scan
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const int n = slice * NS + i;
                if (c_ok && n < p.n_state)
                    // (batch, chunk, state, channel): the 32 lanes of a warp are 32 consecutive channels -> one 128-byte run per store
                    p.h_ckpt[(((long long)b * nchunks + chunk) * p.n_state + n) * p.d + c] = (i & 1) ? h2[i >> 1].y : h2[i >> 1].x;
            }
        }
        cp_async_wait<0>();
        __syncthreads();  // (A) tile `buf` landed; everybody is done with the previous chunk's combine
        if (chunk + 1 < nchunks) stage(chunk + 1, buf ^ 1);
        // discretise: delta <- softplus(delta + bias), once per (t, channel)
        for (int i = tid; i < TC * CH; i += NT) {
            const int t = i / CH, cc = i - t * CH;
            float v = sm.dl[buf][t][cc];
            if (p.delta_bias && c0 + cc < p.d) v += __ldg(p.delta_bias + c0 + cc);
            if (p.delta_softplus) v = softplusf_(v);
            sm.dl[buf][t][cc] = v;
        }
        __syncthreads();  // (B)
        // recurrence: packed 2-wide fp32 math (FMUL2 / FFMA2, sm_100) halves the FMA-pipe issue slots per state update, so the
        // loop is bounded by the MUFU ex2 rate alone (1 ex2 per update)
#pragma unroll kScanUnroll
        for (int t = 0; t < tn; ++t) {
            const float dl = sm.dl[buf][t][ch];
            seg_dl += dl;
            const float du = dl * sm.u[buf][t][ch];
            const float2 dl2 = make_float2(dl, dl), du2 = make_float2(du, du);
            const float4* bq = reinterpret_cast<const float4*>(&sm.Bm[buf][t][slice * NS]);
            const float4* cq = reinterpret_cast<const float4*>(&sm.Cm[buf][t][slice * NS]);
            float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < NS / 4; ++q) {
                const float4 bv = bq[q], cv = cq[q];
                float2 x0 = __fmul2_rn(dl2, a2p[2 * q]), x1 = __fmul2_rn(dl2, a2p[2 * q + 1]);
                x0.x = ex2_approx(x0.x); x0.y = ex2_approx(x0.y);
                x1.x = ex2_approx(x1.x); x1.y = ex2_approx(x1.y);
                h2[2 * q] = __ffma2_rn(x0, h2[2 * q], __fmul2_rn(du2, make_float2(bv.x, bv.y)));
                h2[2 * q + 1] = __ffma2_rn(x1, h2[2 * q + 1], __fmul2_rn(du2, make_float2(bv.z, bv.w)));
                acc0 = __ffma2_rn(h2[2 * q], make_float2(cv.x, cv.y), acc0);
                acc1 = __ffma2_rn(h2[2 * q + 1], make_float2(cv.z, cv.w), acc1);
            }
            const float2 a = __fadd2_rn(acc0, acc1);
            sm.ypart[slice][t][ch] = a.x + a.y;
        }
        __syncthreads();  // (C)
        // combine + gate + store (skipped by the first pass of the segment-parallel mode: only the end state is wanted)
        if (p.y)
        for (int i = tid; i < tn * CH; i += NT) {
            const int t = i / CH, cc = i - t * CH;
            const int cg = c0 + cc;
            if (cg >= p.d) continue;
            float yv = 0.f;
#pragma unroll
            for (int s = 0; s < SL; ++s) yv += sm.ypart[s][t][cc];
            if (p.Dskip) yv = fmaf(__ldg(p.Dskip + cg), sm.u[buf][t][cc], yv);
            if (p.z) {
                const float zz = sm.z[buf][t][cc];
                yv *= __fdividef(zz, 1.0f + __expf(-zz));
            }
            p.y[(long long)b * p.y_bs + (long long)(row0 + t0 + t) * p.y_rs + cg] = yv;
        }
    }
    if (p.seg_sum && slice == 0 && c_ok) p.seg_sum[(long long)bb * p.d + c] = seg_dl;
    if (p.h_out) {
        if (state_tile) {
            __syncthreads();          // the last combine phase has read ypart
#pragma unroll
            for (int q = 0; q < NS / 4; ++q)
                *reinterpret_cast<float4*>(hs_chunk(ch, slice * (NS / 4) + q)) = make_float4(h2[2 * q].x, h2[2 * q].y, h2[2 * q + 1].x, h2[2 * q + 1].y);
            __syncthreads();
            float4* dst = reinterpret_cast<float4*>(p.h_out + ((long long)bb * p.d + c0) * NP);
            for (int i = tid; i < CH * NP / 4; i += NT) {
                const int chn = i / (NP / 4), chunk = i - chn * (NP / 4);
                dst[i] = *reinterpret_cast<const float4*>(hs_chunk(chn, chunk));
            }
        } else if (state_vec) {
#pragma unroll
            for (int q = 0; q < NS / 4; ++q)
                *(reinterpret_cast<float4*>(p.h_out + ((long long)bb * p.d + c) * p.n_state + slice * NS) + q) =
                    make_float4(h2[2 * q].x, h2[2 * q].y, h2[2 * q + 1].x, h2[2 * q + 1].y);
        } else {
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const int n = slice * NS + i;
                if (c_ok && n < p.n_state) p.h_out[((long long)bb * p.d + c) * p.n_state + n] = (i & 1) ? h2[i >> 1].y : h2[i >> 1].x;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// selective_scan, few tokens per call with carried state (streaming: 1-2 tokens per stream per feed(); the reference's
// Mamba.step / selective_state_update).  The work is then the 2 x 512 KB of state per (stream, layer) that must be read and
// written, not the recurrence: one thread owns 4 states of one channel (16 lanes = one channel's 256-byte state row, fully
// coalesced), loops over the tokens, and the 16 lanes reduce <h, C_t> with shuffles.  No shared memory, 11 registers of
// state: occupancy hides the latency that bounded the chunked kernel (2.0 ms per layer at 4096 streams x 1 token).
// ---------------------------------------------------------------------------------------------------------
// CPT channels per thread (independent 16-byte state loads in flight per thread: one load per thread left the kernel at 80 % of the
// copy bandwidth); H16: the carried state is stored as fp16 (reduced-precision streaming state, half the traffic; recurrence in fp32)
template <int T, int CPT, bool H16>      // tokens per call (compile-time: every token's inputs are requested before the dependent recurrence starts)
__global__ void __launch_bounds__(256) selective_scan_step_kernel(const cum_scan_desc p) {
    pdl_trigger();
    pdl_wait();          // PDL: nothing of the previous kernel is touched before this point

    const int b = blockIdx.y;
    const int g = threadIdx.x & 15;
    const int cbase = blockIdx.x * (16 * CPT) + (threadIdx.x >> 4);      // d % (16 CPT) == 0 (host check): whole 16-lane groups stay or leave
    if (cbase >= p.d) return;
    float4 a[CPT], h[CPT];
    float bias[CPT], dk[CPT];
    float dlv[CPT][T], uv[CPT][T], zv[CPT][T];
    float4 bvv[T], cvv[T];
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
        const int c = cbase + 16 * q;
        const long long hoff = ((long long)b * p.d + c) * 64;
        if (!p.h0) h[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        else if (H16) {
            const uint2 raw = *(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.h0) + hoff) + g);
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
            h[q] = make_float4(lo.x, lo.y, hi.x, hi.y);
        } else h[q] = *(reinterpret_cast<const float4*>(p.h0 + hoff) + g);
    }
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
        const int c = cbase + 16 * q;
        a[q] = __ldg(reinterpret_cast<const float4*>(p.a2 + (long long)c * 64) + g);
        bias[q] = p.delta_bias ? __ldg(p.delta_bias + c) : 0.f;
        dk[q] = p.Dskip ? __ldg(p.Dskip + c) : 0.f;
        const float* ub = p.u + (long long)b * p.u_bs + c;
        const float* db = p.delta + (long long)b * p.dl_bs + c;
        const float* zb = p.z ? p.z + (long long)b * p.z_bs + c : nullptr;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            dlv[q][t] = db[(long long)t * p.dl_rs];
            uv[q][t] = ub[(long long)t * p.u_rs];
            zv[q][t] = zb ? zb[(long long)t * p.z_rs] : 0.f;
        }
    }
    const float* Bb = p.Bm + (long long)b * p.B_bs + 4 * g;
    const float* Cb = p.Cm + (long long)b * p.C_bs + 4 * g;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        bvv[t] = *reinterpret_cast<const float4*>(Bb + (long long)t * p.B_rs);
        cvv[t] = *reinterpret_cast<const float4*>(Cb + (long long)t * p.C_rs);
    }
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
        const int c = cbase + 16 * q;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            float dl = dlv[q][t] + bias[q];
            if (p.delta_softplus) dl = softplusf_(dl);
            const float u = uv[q][t];
            const float du = dl * u;
            const float4 bv = bvv[t], cv = cvv[t];
            h[q].x = fmaf(ex2_approx(dl * a[q].x), h[q].x, du * bv.x);
            h[q].y = fmaf(ex2_approx(dl * a[q].y), h[q].y, du * bv.y);
            h[q].z = fmaf(ex2_approx(dl * a[q].z), h[q].z, du * bv.z);
            h[q].w = fmaf(ex2_approx(dl * a[q].w), h[q].w, du * bv.w);
            float part = fmaf(h[q].x, cv.x, fmaf(h[q].y, cv.y, fmaf(h[q].z, cv.z, h[q].w * cv.w)));
            part += __shfl_xor_sync(0xffffffffu, part, 8);
            part += __shfl_xor_sync(0xffffffffu, part, 4);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            if (g == 0) {
                float yv = fmaf(dk[q], u, part);
                if (p.z) yv *= __fdividef(zv[q][t], 1.0f + __expf(-zv[q][t]));
                p.y[(long long)b * p.y_bs + (long long)t * p.y_rs + c] = yv;
            }
        }
        if (p.h_out) {
            const long long hoff = ((long long)b * p.d + c) * 64;
            if (H16) {
                const __half2 lo = __floats2half2_rn(h[q].x, h[q].y), hi = __floats2half2_rn(h[q].z, h[q].w);
                *(reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.h_out) + hoff) + g) =
                    make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
            } else *(reinterpret_cast<float4*>(p.h_out + hoff) + g) = h[q];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// TMA-staged state update (many streams: BASELINE configs[2], 4096 streams x 1 hop).  The per-thread-load kernel above is bound
// by WARP LIFETIME, not by bytes: a warp issues one 16-byte state load per thread, waits ~1 us for DRAM, computes for ~0.2 us and
// retires -- 64 resident warps x 512 B = 32 KB in flight per SM, 2.6 TB/s of reads whatever the state's element size (fp16 state:
// the same 2.4 ms).  Here the state moves through shared memory: a persistent CTA owns ONE block of 64 channels (its A2 rows, D and
// dt bias stay in registers) and walks over streams; per (stream, channel block) tile one elected thread requests the tile's
// 64 x 64 state block as TMA tensor boxes (128-byte swizzle: conflict-free reads with compile-time register indices) and the
// tile's 256-byte rows of delta, u, z, B_t, C_t as plain bulk copies onto one mbarrier, SB_NST stages deep; the 256 threads update
// the state IN PLACE in shared memory (thread = 16 states of one channel: softplus 4x instead of 16x redundant, 2 shuffles
// instead of 4, ~6 instructions per state instead of ~20) and TMA stores write the tile back.  In flight per SM: 3 CTAs x 2 tiles.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_3d_(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int SB_CH = 64;          // channels per tile
constexpr int SB_NST = 3;          // stages
constexpr int SB_MAXT = 8;         // tokens per call (runtime): the state is read and written once per call, whatever T.  Measured at 4096
                                   // streams x 3 layers vs the per-thread / chunked kernels: 1 token 2.23 vs 2.44 ms, 2: 2.30 vs 3.76, 4: 3.85 vs
                                   // 6.96, 8: 7.88 vs 8.77, 16: 15.1 vs 12.1 -> up to 8 tokens, the chunked kernel beyond
template <bool H16> struct StepBulkCfg {
    static constexpr uint32_t STATE_BYTES = SB_CH * 64 * (H16 ? 2 : 4);   // fp32: two boxes of 64 rows x 128 B; fp16: one
    static constexpr uint32_t ROW_BYTES = SB_CH * 4;                       // one token's delta / u / z / B / C row of the tile (64 floats)
    static uint32_t stage_bytes(int T) { return (STATE_BYTES + 5u * (uint32_t)T * ROW_BYTES + 1023u) & ~1023u; }     // swizzle atoms are 1 KB aligned
    static uint32_t smem_bytes(int T) { return SB_NST * stage_bytes(T) + 64 /*barriers*/ + 1024 /*alignment*/; }
};

template <bool H16>
__global__ void __launch_bounds__(256, 3) selective_scan_step_bulk_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut,
                                                                            const cum_scan_desc p, int groups, uint32_t stage_bytes) {
    using Cfg = StepBulkCfg<H16>;
    const int T = p.len;
    extern __shared__ uint8_t sb_raw[];
    const uint32_t base = (smem_u32(sb_raw) + 1023u) & ~1023u;
    uint8_t* gen = sb_raw + (base - smem_u32(sb_raw));
    const uint32_t bar0 = base + SB_NST * stage_bytes;
    const int tid = threadIdx.x;
    if (tid == 0) {
        tma_prefetch_desc(&tmIn);
        tma_prefetch_desc(&tmOut);
        for (int s = 0; s < SB_NST; ++s) mbar_init(bar0 + 8u * s, 1);
        fence_barrier_init();
    }
    pdl_trigger();
    __syncthreads();
    pdl_wait();          // PDL: nothing of the previous kernel is touched before this point

    const int c0 = blockIdx.x * SB_CH;
    const int ch = tid >> 2, sub = tid & 3;
    const int c = c0 + ch;
    // This thread's 16 states as four 16-byte pieces i = 0..3 (compile-time register indices).
    //   fp32: piece i = logical chunk 2 sub + (i & 1) of box (i >> 1): states 32 (i >> 1) + 8 sub + 4 (i & 1) + {0..3}
    //   fp16: pieces (2 e, 2 e + 1) = the 8 halves of logical chunk 2 sub + e: states 16 sub + 8 e + {0..7}
    // physical 16-byte chunk inside the 128-byte row of channel ch: logical ^ (ch & 7) -- a quarter-warp (2 channels x 4 subs) hits 8 banks groups
    auto state0 = [&](int i) { return H16 ? 16 * sub + 4 * i : 32 * (i >> 1) + 8 * sub + 4 * (i & 1); };
    float a[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.a2 + (long long)c * 64 + state0(i)));
        a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
    }
    const float bias = p.delta_bias ? __ldg(p.delta_bias + c) : 0.f;
    const float dk = p.Dskip ? __ldg(p.Dskip + c) : 0.f;
    const uint32_t x7 = (uint32_t)(ch & 7);
    uint32_t poff[H16 ? 2 : 4];              // byte offset of this thread's pieces inside the stage's state region
    if (H16) {
#pragma unroll
        for (int e = 0; e < 2; ++e) poff[e] = (uint32_t)ch * 128u + (((uint32_t)(2 * sub + e)) ^ x7) * 16u;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) poff[i] = (uint32_t)(i >> 1) * 8192u + (uint32_t)ch * 128u + (((uint32_t)(2 * sub + (i & 1))) ^ x7) * 16u;
    }

    auto issue = [&](int b, int s) {            // one thread: every byte of tile (b, this channel block) onto the stage's barrier
        const uint32_t st = base + (uint32_t)s * stage_bytes, bar = bar0 + 8u * s;
        mbar_arrive_expect_tx(bar, Cfg::STATE_BYTES + (p.z ? 5u : 4u) * (uint32_t)T * Cfg::ROW_BYTES);
        const int row = b * p.d + c0;
        tma_load_3d(st, &tmIn, bar, 0, row, 0);
        if (!H16) tma_load_3d(st + 8192u, &tmIn, bar, 32, row, 0);
        for (int t = 0; t < T; ++t) {
            const uint32_t r = st + Cfg::STATE_BYTES + (uint32_t)t * 5 * Cfg::ROW_BYTES;
            bulk_load(r, p.delta + (long long)b * p.dl_bs + (long long)t * p.dl_rs + c0, Cfg::ROW_BYTES, bar);
            bulk_load(r + Cfg::ROW_BYTES, p.u + (long long)b * p.u_bs + (long long)t * p.u_rs + c0, Cfg::ROW_BYTES, bar);
            if (p.z) bulk_load(r + 2 * Cfg::ROW_BYTES, p.z + (long long)b * p.z_bs + (long long)t * p.z_rs + c0, Cfg::ROW_BYTES, bar);
            bulk_load(r + 3 * Cfg::ROW_BYTES, p.Bm + (long long)b * p.B_bs + (long long)t * p.B_rs, Cfg::ROW_BYTES, bar);
            bulk_load(r + 4 * Cfg::ROW_BYTES, p.Cm + (long long)b * p.C_bs + (long long)t * p.C_rs, Cfg::ROW_BYTES, bar);
        }
    };

    const int b_first = blockIdx.y;
    const int n_tiles = (p.batch - b_first + groups - 1) / groups;          // streams b_first, b_first + groups, ...
    if (tid == 0) {
        for (int i = 0; i < SB_NST - 1 && i < n_tiles; ++i) issue(b_first + i * groups, i);
    }
    for (int it = 0; it < n_tiles; ++it) {
        const int s = it % SB_NST;
        const int b = b_first + it * groups;
        mbar_wait(bar0 + 8u * s, (uint32_t)(it / SB_NST) & 1u);
        uint8_t* stage = gen + (size_t)s * stage_bytes;
        float h[16];
        if (H16) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const uint4 raw = *reinterpret_cast<const uint4*>(stage + poff[e]);
                const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
                    h[8 * e + 2 * k] = f.x; h[8 * e + 2 * k + 1] = f.y;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(stage + poff[i]);
                h[4 * i] = v.x; h[4 * i + 1] = v.y; h[4 * i + 2] = v.z; h[4 * i + 3] = v.w;
            }
        }
#pragma unroll 1
        for (int t = 0; t < T; ++t) {
            const float* row = reinterpret_cast<const float*>(stage + Cfg::STATE_BYTES) + t * 5 * SB_CH;
            float dl = row[ch] + bias;
            if (p.delta_softplus) dl = softplusf_(dl);
            const float u = row[SB_CH + ch];
            const float du = dl * u;
            float part = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 bv = *reinterpret_cast<const float4*>(row + 3 * SB_CH + state0(i));
                const float4 cv = *reinterpret_cast<const float4*>(row + 4 * SB_CH + state0(i));
                h[4 * i] = fmaf(ex2_approx(dl * a[4 * i]), h[4 * i], du * bv.x);
                h[4 * i + 1] = fmaf(ex2_approx(dl * a[4 * i + 1]), h[4 * i + 1], du * bv.y);
                h[4 * i + 2] = fmaf(ex2_approx(dl * a[4 * i + 2]), h[4 * i + 2], du * bv.z);
                h[4 * i + 3] = fmaf(ex2_approx(dl * a[4 * i + 3]), h[4 * i + 3], du * bv.w);
                part = fmaf(h[4 * i], cv.x, fmaf(h[4 * i + 1], cv.y, fmaf(h[4 * i + 2], cv.z, fmaf(h[4 * i + 3], cv.w, part))));
            }
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            if (sub == 0) {
                float yv = fmaf(dk, u, part);
                if (p.z) { const float zv = row[2 * SB_CH + ch]; yv *= __fdividef(zv, 1.0f + __expf(-zv)); }
                p.y[(long long)b * p.y_bs + (long long)t * p.y_rs + c] = yv;
            }
        }
        // new state back into the stage, in place
        if (H16) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                uint32_t w[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const __half2 v = __floats2half2_rn(h[8 * e + 2 * k], h[8 * e + 2 * k + 1]);
                    w[k] = *reinterpret_cast<const uint32_t*>(&v);
                }
                *reinterpret_cast<uint4*>(stage + poff[e]) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(stage + poff[i]) = make_float4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
        }
        fence_proxy_async();            // generic-proxy writes of the state -> visible to the TMA store (async proxy)
        __syncthreads();
        if (tid == 0) {
            const uint32_t st = base + (uint32_t)s * stage_bytes;
            const int row = b * p.d + c0;
            tma_store_3d_(&tmOut, st, 0, row, 0);
            if (!H16) tma_store_3d_(&tmOut, st + 8192u, 32, row, 0);
            bulk_commit_group();
            // refill the stage of the PREVIOUS iteration: its store (the group before the one just committed) has read it
            const int nxt = it + SB_NST - 1;
            if (nxt < n_tiles) {
                bulk_wait_group_read<1>();
                issue(b_first + nxt * groups, nxt % SB_NST);
            }
        }
    }
    if (tid == 0) bulk_wait_group<0>();     // the stores must have left shared memory before the CTA exits
}

template <bool H16>
static int launch_step_bulk(const cum_scan_desc& d, cudaStream_t st) {
    using Cfg = StepBulkCfg<H16>;
    auto kern = selective_scan_step_bulk_kernel<H16>;
    const uint32_t stage = Cfg::stage_bytes(d.len), smem = Cfg::smem_bytes(d.len);
    { const int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(kern), (int)Cfg::smem_bytes(SB_MAXT), "cudaFuncSetAttribute(selective_scan_step_bulk_kernel)"); if (rc_attr) return rc_attr; }
    // the carried state as a 2-D tensor (batch * d rows of 64 states): boxes of 64 rows x 128 bytes, 128-byte swizzle.  The fp16 state
    // uses the 2-byte tensor type of the map helper (raw copies: no conversion takes place)
    CUtensorMap tmIn, tmOut;
    const uint64_t rows = (uint64_t)d.batch * (uint64_t)d.d;
    int rc = make_tensor_map(&tmIn, d.h0, 64, rows, 1, 64, rows * 64, H16 ? 64 : 32, SB_CH, "scan state in", H16, H16);
    if (rc) return rc;
    rc = make_tensor_map(&tmOut, d.h_out, 64, rows, 1, 64, rows * 64, H16 ? 64 : 32, SB_CH, "scan state out", H16, H16);
    if (rc) return rc;
    const int cblocks = d.d / SB_CH;
    // every CTA of the grid must be resident at once (one CTA more than fit would run as a second wave and double the kernel's
    // time): 3 CTAs per SM up to 4 tokens per call (56-67 KB of shared memory each), 2 beyond; each walks over batch / groups streams
    const int per_sm = (int)(232448u / (smem + 1024u)) >= 3 ? 3 : ((int)(232448u / (smem + 1024u)) >= 2 ? 2 : 1);
    int groups = (per_sm * sm_count()) / cblocks;
    if (groups < 1) groups = 1;
    if (groups > d.batch) groups = d.batch;
    if (groups > 65535) groups = 65535;
    cudaError_t e = launch_kernel(kern, dim3((unsigned)cblocks, (unsigned)groups), dim3(256), smem, st, tmIn, tmOut, d, groups, stage);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(selective_scan_step_bulk_kernel)");
    return CUM_OK;
}

// many streams with whole 64-channel blocks and 16-byte aligned rows: the bulk-staged kernel (CUM_SCAN_STEP_BULK=0 disables: A/B)
static bool step_bulk_ok(const cum_scan_desc& d) {
    static int env = -1;
    if (env < 0) { const char* e = getenv("CUM_SCAN_STEP_BULK"); env = (e && e[0] == '0') ? 0 : 1; }
    if (!env || d.d % SB_CH || d.n_state != 64 || !d.h0 || !d.h_out || d.len > SB_MAXT || d.h_ckpt) return false;
    if ((long long)d.batch * (d.d / SB_CH) < 4LL * sm_count()) return false;        // few tiles: the per-thread kernel has more parallelism
    const auto ok = [](const void* q, long long bs, long long rs) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0 && bs % 4 == 0 && rs % 4 == 0; };
    return ok(d.u, d.u_bs, d.u_rs) && ok(d.delta, d.dl_bs, d.dl_rs) && (!d.z || ok(d.z, d.z_bs, d.z_rs)) && ok(d.Bm, d.B_bs, d.B_rs) &&
           ok(d.Cm, d.C_bs, d.C_rs) && (reinterpret_cast<uintptr_t>(d.h0) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.h_out) & 15) == 0;
}

template <int T, bool H16>
static cudaError_t launch_step(const cum_scan_desc& d, cudaStream_t st) {
    // channels per thread = independent 16-byte (fp32 state) / 8-byte (fp16 state) loads in flight per thread.
    // CUM_SCAN_STEP_CPT=1|2|4 overrides (A/B measurements); default 2, and 4 for the fp16 state (half the bytes per load)
    static int cpt_env = -1;
    if (cpt_env < 0) { const char* e = getenv("CUM_SCAN_STEP_CPT"); cpt_env = e ? atoi(e) : 0; }
    int cpt = cpt_env > 0 ? cpt_env : (H16 ? 4 : 2);
    while (cpt > 1 && d.d % (16 * cpt)) cpt >>= 1;
    const dim3 grid((unsigned)(d.d / (16 * cpt)), (unsigned)d.batch);
    if (cpt >= 4) return launch_kernel(selective_scan_step_kernel<T, 4, H16>, grid, dim3(256), 0, st, d);
    if (cpt == 2) return launch_kernel(selective_scan_step_kernel<T, 2, H16>, grid, dim3(256), 0, st, d);
    return launch_kernel(selective_scan_step_kernel<T, 1, H16>, grid, dim3(256), 0, st, d);
}

template <int NS, int SL, int CH, int TC>
static int launch_scan(const ScanK& d, cudaStream_t st) {
    auto kern = selective_scan_fwd_kernel<NS, SL, CH, TC>;
    const size_t smem = sizeof(ScanSmem<NS, SL, CH, TC>);
    { const int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(kern), (int)smem, "cudaFuncSetAttribute(selective_scan_fwd_kernel)"); if (rc_attr) return rc_attr; }
    dim3 grid((unsigned)cdiv(d.d, CH), (unsigned)(d.batch * (d.nseg > 1 ? d.nseg : 1)));
    cudaError_t e = launch_kernel(kern, grid, dim3(CH * SL), smem, st, d);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(selective_scan_fwd_kernel)");
    return CUM_OK;
}

static int launch_scan_any(const ScanK& k, cudaStream_t st) {
    if (k.n_state > 32) return launch_scan<16, 4, 64, 16>(k, st);
    if (k.n_state > 16) return launch_scan<16, 2, 64, 16>(k, st);
    if (k.n_state > 8)  return launch_scan<16, 1, 64, 16>(k, st);
    return launch_scan<8, 1, 64, 16>(k, st);
}

// ---------------------------------------------------------------------------------------------------------
// Segment-parallel scan for SMALL batches (a single long clip: batch 1 x 60 s is 32 CTAs on 148 SMs with the plain kernel).
// The recurrence h_t = dA_t h_{t-1} + x_t is linear in h, so a clip is cut into `nseg` time segments that run concurrently:
//   pass 1  every segment scans from h = 0: its local end state E_s and sum_t dl_t (total decay P_s = exp2(a2 * sum dl))
//   pass 2  carry: H_0 = h0, H_{s+1} = P_s * H_s + E_s                      (tiny: nseg steps per (item, channel, state))
//   pass 3  every segment scans again from its true start state H_s and writes y
// Twice the arithmetic, nseg-fold parallelism: pays when cdiv(d, 64) * batch CTAs leave most SMs idle.
// ---------------------------------------------------------------------------------------------------------
struct SegPlan { int nseg, seg_len; size_t state_elems, sum_elems; };

static SegPlan plan_segments(const cum_scan_desc& d) {
    SegPlan s{1, d.len, 0, 0};
    const long long ctas = cdiv(d.d, 64) * d.batch;
    if (d.h_ckpt || d.len < 256 || 2 * ctas > sm_count()) return s;
    long long n = (2LL * sm_count() + ctas - 1) / ctas;
    // segments of at least 64 steps -- except for the handful of CTAs of a pruned checkpoint (d_inner 8-48 at batch 1-4), whose scan is
    // pure latency: 16-step segments (one staging chunk each) cut a 624-token pass from 21-31 us to one chunk's time
    const long long min_seg = ctas <= 8 ? 16 : 64;
    if (n > d.len / min_seg) n = d.len / min_seg;
    if (n > 64) n = 64;
    if (n < 2) return s;
    s.seg_len = (int)(cdiv(cdiv(d.len, n), 16) * 16);
    s.nseg = (int)cdiv(d.len, s.seg_len);
    if (s.nseg < 2 || (long long)d.batch * s.nseg > 65535) { s.nseg = 1; s.seg_len = d.len; return s; }
    s.state_elems = (size_t)d.batch * s.nseg * d.d * d.n_state;
    s.sum_elems = ((size_t)d.batch * s.nseg * d.d + 3) / 4 * 4;
    return s;
}

long long selective_scan_workspace_bytes(const cum_scan_desc& d) {
    if (d.batch <= 0 || d.len <= 0 || d.d <= 0 || d.n_state <= 0) return 0;
    const SegPlan s = plan_segments(d);
    return s.nseg > 1 ? (long long)((2 * s.state_elems + s.sum_elems) * sizeof(float)) : 0;
}

// H_0 = h0 (or 0); H_{s+1} = exp2(a2 * sum_s) * H_s + E_s; h_start[s] = H_s; h_out = H_nseg
__global__ void __launch_bounds__(256) scan_carry_kernel(const float* __restrict__ e, const float* __restrict__ sum, const float* __restrict__ a2,
                                                          const float* __restrict__ h0, float* __restrict__ h_start, float* __restrict__ h_out,
                                                          int batch, int nseg, int d, int n) {
    pdl_trigger();
    pdl_wait();          // PDL: nothing of the previous kernel is touched before this point

    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (item, channel, state)
    if (i >= (long long)batch * d * n) return;
    const int st = (int)(i % n);
    const int c = (int)((i / n) % d);
    const int b = (int)(i / ((long long)n * d));
    const float a = a2[(long long)c * n + st];
    float h = h0 ? h0[i] : 0.f;
    for (int s = 0; s < nseg; ++s) {
        const long long o = (((long long)b * nseg + s) * d + c) * n + st;
        h_start[o] = h;
        h = fmaf(ex2_approx(a * sum[((long long)b * nseg + s) * d + c]), h, e[o]);
    }
    if (h_out) h_out[i] = h;
}

int selective_scan_fwd(const cum_scan_desc& d, cudaStream_t st) {
    CUM_REQUIRE(d.u && d.delta && d.Bm && d.Cm && d.y && d.a2, "selective_scan: null pointer");
    CUM_REQUIRE(d.batch > 0 && d.batch <= 65535 && d.len > 0 && d.d > 0 && d.n_state > 0, "selective_scan: bad shape");
    CUM_REQUIRE(d.n_state <= 64, "selective_scan: n_state=%d > 64 not supported", d.n_state);
    const auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    // measured at 4096 streams (3 layers): 1 token 2.5 ms vs 6.0 ms chunked, 2 tokens 4.1 vs 6.2, 4 tokens 7.3 vs 6.4 -> up to 2 tokens
    const bool step_ok = d.n_state == 64 && d.d % 16 == 0 && (d.h0 || d.h_out) && !d.h_ckpt && al16(d.a2) && al16(d.Bm) && al16(d.Cm) &&
                         (!d.h0 || al16(d.h0)) && (!d.h_out || al16(d.h_out)) && (d.B_rs | d.C_rs | d.B_bs | d.C_bs) % 4 == 0;
    if (d.state_f16) {
        // reduced-precision carried state (streaming variant, reported separately): fp16 storage, fp32 recurrence; only the state-update
        // kernels read / write it, longer calls advance in pieces (8 tokens per launch TMA-staged, else 2)
        CUM_REQUIRE(step_ok && d.h0 && d.h_out, "selective_scan: state_f16 needs n_state = 64, d %% 16 == 0, aligned operands and both h0 and h_out");
        cum_scan_desc s = d;
        for (int t = 0; t < d.len;) {
            s.u = d.u + (long long)t * d.u_rs; s.delta = d.delta + (long long)t * d.dl_rs; s.z = d.z ? d.z + (long long)t * d.z_rs : nullptr;
            s.Bm = d.Bm + (long long)t * d.B_rs; s.Cm = d.Cm + (long long)t * d.C_rs; s.y = d.y + (long long)t * d.y_rs;
            if (t > 0) s.h0 = d.h_out;
            s.len = d.len - t < SB_MAXT ? d.len - t : SB_MAXT;
            if (step_bulk_ok(s)) {
                const int rc = launch_step_bulk<true>(s, st);
                if (rc) return rc;
                t += s.len;
                continue;
            }
            const int n = d.len - t >= 2 ? 2 : 1;
            s.len = n;
            cudaError_t e = n == 1 ? launch_step<1, true>(s, st) : launch_step<2, true>(s, st);
            if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(selective_scan_step_kernel, fp16 state)");
            t += n;
        }
        return CUM_OK;
    }
    // many streams, up to 8 tokens per call: the TMA-staged kernel (the state is read and written once per call; measured at 4096
    // streams x 3 layers: 1 token 2.18 ms vs 2.42 per-thread loads, 2 tokens 2.24 vs 3.72)
    if (step_bulk_ok(d)) return launch_step_bulk<false>(d, st);
    if (d.len <= 2 && step_ok) {
        cudaError_t e = d.len == 1 ? launch_step<1, false>(d, st) : launch_step<2, false>(d, st);
        if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(selective_scan_step_kernel)");
        return CUM_OK;
    }
    ScanK k;
    static_cast<cum_scan_desc&>(k) = d;
    k.nseg = 1; k.seg_len = d.len; k.seg_sum = nullptr;
    const SegPlan sp = plan_segments(d);
    if (sp.nseg > 1 && d.workspace && al16(d.workspace) &&
        d.workspace_bytes >= (long long)((2 * sp.state_elems + sp.sum_elems) * sizeof(float))) {
        float* e = reinterpret_cast<float*>(d.workspace);
        float* sum = e + sp.state_elems;
        float* hstart = sum + sp.sum_elems;
        // pass 1: local end states + per-segment decay sums (no y, no gate)
        ScanK a = k;
        a.nseg = sp.nseg; a.seg_len = sp.seg_len; a.seg_sum = sum;
        a.y = nullptr; a.h0 = nullptr; a.h_out = e; a.h_ckpt = nullptr;
        int rc = launch_scan_any(a, st);
        if (rc) return rc;
        // pass 2: carry across segments
        const long long n = (long long)d.batch * d.d * d.n_state;
        cudaError_t ce = launch_kernel(scan_carry_kernel, dim3((unsigned)cdiv(n, 256)), dim3(256), 0, st, (const float*)e, (const float*)sum, d.a2, d.h0, hstart, d.h_out, d.batch, sp.nseg, d.d, d.n_state);
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaLaunchKernelEx(scan_carry_kernel)");
        // pass 3: every segment from its true start state
        ScanK c = k;
        c.nseg = sp.nseg; c.seg_len = sp.seg_len; c.seg_sum = nullptr;
        c.h0 = hstart; c.h_out = nullptr; c.h_ckpt = nullptr;
        return launch_scan_any(c, st);
    }
    return launch_scan_any(k, st);
}

}  // namespace cum
