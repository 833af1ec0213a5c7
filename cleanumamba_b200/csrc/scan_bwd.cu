// Reverse selective scan (backward of cum_selective_scan_fwd).  Reference arithmetic: autograd through
// selective_scan_ref (mamba_ssm); upstream's CUDA kernel is selective_scan_bwd_kernel.cuh.
//
// Forward:  dl = softplus(delta + bias);  e_t = exp(dl A);  h_t = e_t h_{t-1} + (dl u_t) B_t;
//           y_t = <h_t, C_t> + D u_t;     out_t = y_t silu(z_t)
// Backward (g_t = dL/dh_t, carried in reverse):  g_t = dy_t C_t + e_{t+1} g_{t+1}
//   dC_t[n] += sum_d dy_t h_t[n]            dB_t[n] += sum_d g_t[n] dl u_t
//   du_t     = dy_t D + dl sum_n g_t B_t    ddl_t    = u_t sum_n g_t B_t + sum_n g_t h_{t-1} A e_t
//   dA[d,n] += sum_t g_t h_{t-1} dl e_t     dD[d]   += sum_t dy_t u_t     dz_t = dout_t y_t silu'(z_t)
//   ddelta_t = ddl_t sigmoid(delta_t + bias)   dbias[d] += sum_t ddelta_t   dA_log = dA * A
//
// Decomposition: CTA = 32 channels x SL state slices of 4 states (a warp = 32 channels of one slice, so B_t / C_t are
// shared-memory broadcasts and the d-reductions of dB / dC are warp shuffles).  Time runs in chunks of SCAN_TC = 16
// in REVERSE; the forward pass stored h at every chunk start (h_ckpt), so each chunk is recomputed forward once with
// its 16 x 4 state history held in registers, then walked backwards.  Cross-slice sums go through shared memory.
#include "common.cuh"

namespace cum {

constexpr int SB_TC = 16;     // must equal the forward kernel's chunk length (checkpoint spacing)
constexpr int SB_CH = 32;
constexpr float LN2F = 0.69314718055994530942f;

__device__ __forceinline__ float ex2_approx_b(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float softplus_b(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

// Sum 8 per-lane values over the 32 lanes of a warp with 9 shuffles instead of 8 x 5: each xor step halves the number of
// values a lane still owns.  On return lane L holds the warp total of value index ((L>>4)&1)*4 + ((L>>3)&1)*2 + ((L>>2)&1)
// (replicated over the 4 lanes that share those bits).
__device__ __forceinline__ float warp_sum8(const float (&v)[8], int lane) {
    float w4[4], w2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b4 ? v[i] : v[i + 4];
        const float keep = b4 ? v[i + 4] : v[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b3 ? w4[i] : w4[i + 2];
        const float keep = b3 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float send = b2 ? w2[0] : w2[1];
    const float keep = b2 ? w2[1] : w2[0];
    float r = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

template <int SL>
struct ScanBwdSmem {
    static constexpr int NP = 4 * SL;
    float u[SB_TC][SB_CH], dl[SB_TC][SB_CH], duv[SB_TC][SB_CH], draw[SB_TC][SB_CH], z[SB_TC][SB_CH], dout[SB_TC][SB_CH], dy[SB_TC][SB_CH];
    float Bm[SB_TC][NP], Cm[SB_TC][NP];
    float part0[SL][SB_TC][SB_CH];      // y partials, then sum_n g B partials
    float part1[SL][SB_TC][SB_CH];      // sum_n g h_prev A e partials
};

// The kernel is instruction-issue-bound (ncu round 2: 0.56 IPC per scheduler, MUFU 20 %, DRAM 5 %; 47 SASS instructions per state
// update).  Every per-state operation therefore runs as PACKED 2-wide fp32 math (FMUL2 / FFMA2 of sm_100: one issue slot per state
// PAIR), dl*u is discretised once per (t, channel) into shared memory, and the per-step address arithmetic of the dB / dC atomics is
// hoisted (each lane's reduction role -- which of the warp's 8 totals it ends up holding -- is fixed for the whole kernel).
template <int SL>
__global__ void __launch_bounds__(32 * SL) selective_scan_bwd_kernel(const cum_scan_bwd_desc p) {
    constexpr int NP = 4 * SL;
    constexpr int NT = 32 * SL;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ScanBwdSmem<SL>& sm = *reinterpret_cast<ScanBwdSmem<SL>*>(smem_raw);
    const cum_scan_desc& f = p.fwd;

    const int b = blockIdx.y, c0 = blockIdx.x * SB_CH;
    const int tid = threadIdx.x, ch = tid & 31, slice = tid >> 5;
    const int c = c0 + ch;
    const bool c_ok = c < f.d;
    const int nchunks = (f.len + SB_TC - 1) / SB_TC;

    float2 a2p[2], Alnp[2], G2[2], dA2[2];      // state pairs (4*slice + {0,1}) and (4*slice + {2,3})
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float av[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int n = slice * 4 + 2 * j + e;
            av[e] = (c_ok && n < f.n_state) ? f.a2[(long long)c * f.n_state + n] : 0.f;
        }
        a2p[j] = make_float2(av[0], av[1]);
        Alnp[j] = make_float2(av[0] * LN2F, av[1] * LN2F);           // A = a2 / log2(e)
        G2[j] = make_float2(0.f, 0.f);
        dA2[j] = make_float2(0.f, 0.f);
    }
    float accD = 0.f, accBias = 0.f;     // per (tid & 31) channel partials of the combine threads
    const float Dv = (f.Dskip && c_ok) ? f.Dskip[c] : 0.f;

    // d-reductions of dB / dC: after warp_sum8 lane L holds the warp total of value ((L>>4)&1)*4 + ((L>>3)&1)*2 + ((L>>2)&1)
    // (values 0..3 = dC of this slice's 4 states, 4..7 = dB); one lane of every four adds it to global memory
    const int ridx = ((ch >> 4) & 1) * 4 + ((ch >> 3) & 1) * 2 + ((ch >> 2) & 1);
    const int rn = slice * 4 + (ridx & 3);
    const bool red_lane = (ch & 3) == 0 && rn < f.n_state;
    float* const red_base = (ridx < 4 ? p.dC + (long long)b * p.dC_bs : p.dB + (long long)b * p.dB_bs) + rn;
    const long long red_rs = ridx < 4 ? p.dC_rs : p.dB_rs;

    // register prefetch buffers: every thread owns a fixed share of each tile
    constexpr int PF_E = (SB_TC * SB_CH + NT - 1) / NT;      // u / delta / z / dout elements per thread
    constexpr int PF_N = (SB_TC * NP + NT - 1) / NT;         // B / C elements per thread
    float pf_u[PF_E], pf_d[PF_E], pf_z[PF_E], pf_o[PF_E], pf_B[PF_N], pf_C[PF_N];
    const float* const ub = f.u + (long long)b * f.u_bs;
    const float* const db = f.delta + (long long)b * f.dl_bs;
    const float* const zb = f.z ? f.z + (long long)b * f.z_bs : nullptr;
    const float* const ob = p.dout + (long long)b * p.dout_bs;
    const float* const Bb = f.Bm + (long long)b * f.B_bs;
    const float* const Cb = f.Cm + (long long)b * f.C_bs;
    auto prefetch = [&](int chunk) {
        const int t0 = chunk * SB_TC;
        const int tn = min(SB_TC, f.len - t0);
        int e = 0;
        for (int i = tid; i < SB_TC * SB_CH; i += NT, ++e) {
            const int t = i >> 5, cc = i & 31;
            const int tt = t0 + t, cg = c0 + cc;
            const bool ok = t < tn && cg < f.d;
            pf_u[e] = ok ? ub[(long long)tt * f.u_rs + cg] : 0.f;
            pf_d[e] = ok ? db[(long long)tt * f.dl_rs + cg] : 0.f;
            pf_z[e] = (ok && zb) ? zb[(long long)tt * f.z_rs + cg] : 0.f;
            pf_o[e] = ok ? ob[(long long)tt * p.dout_rs + cg] : 0.f;
        }
        e = 0;
        for (int i = tid; i < SB_TC * NP; i += NT, ++e) {
            const int t = i / NP, n = i - t * NP;
            const bool ok = t < tn && n < f.n_state;
            pf_B[e] = ok ? Bb[(long long)(t0 + t) * f.B_rs + n] : 0.f;
            pf_C[e] = ok ? Cb[(long long)(t0 + t) * f.C_rs + n] : 0.f;
        }
    };
    prefetch(nchunks - 1);

    for (int chunk = nchunks - 1; chunk >= 0; --chunk) {
        const int t0 = chunk * SB_TC;
        const int tn = min(SB_TC, f.len - t0);
        // ---- tiles of this chunk were fetched into registers while the previous chunk was being processed
        {
            int e = 0;
            for (int i = tid; i < SB_TC * SB_CH; i += NT, ++e) {
                const int t = i >> 5, cc = i & 31;
                const int cg = c0 + cc;
                float dr = pf_d[e];
                if (f.delta_bias && cg < f.d) dr += f.delta_bias[cg];
                const float dlv = f.delta_softplus ? softplus_b(dr) : dr;
                sm.u[t][cc] = pf_u[e];
                sm.draw[t][cc] = dr;
                sm.dl[t][cc] = dlv;
                sm.duv[t][cc] = dlv * pf_u[e];
                sm.z[t][cc] = pf_z[e];
                sm.dout[t][cc] = pf_o[e];
            }
            e = 0;
            for (int i = tid; i < SB_TC * NP; i += NT, ++e) {
                const int t = i / NP, n = i - t * NP;
                sm.Bm[t][n] = pf_B[e];
                sm.Cm[t][n] = pf_C[e];
            }
        }
        if (chunk > 0) prefetch(chunk - 1);       // global latency of the next (earlier) chunk overlaps this chunk's math
        __syncthreads();
        // ---- forward recompute with history in registers (checkpoints: (batch, chunk, state, channel), coalesced over the lanes)
        float2 hs2[2], hist[SB_TC][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float hv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = slice * 4 + 2 * j + e;
                hv[e] = (c_ok && n < f.n_state) ? p.h_ckpt[(((long long)b * nchunks + chunk) * f.n_state + n) * f.d + c] : 0.f;
            }
            hs2[j] = make_float2(hv[0], hv[1]);
        }
        {
            float2 h0 = hs2[0], h1 = hs2[1];
#pragma unroll
            for (int t = 0; t < SB_TC; ++t) {
                const float dlv = sm.dl[t][ch], duv = sm.duv[t][ch];
                const float2 dl2 = make_float2(dlv, dlv), du2 = make_float2(duv, duv);
                const float4 bv = *reinterpret_cast<const float4*>(&sm.Bm[t][slice * 4]);
                const float4 cv = *reinterpret_cast<const float4*>(&sm.Cm[t][slice * 4]);
                float2 x0 = __fmul2_rn(dl2, a2p[0]), x1 = __fmul2_rn(dl2, a2p[1]);
                x0.x = ex2_approx_b(x0.x); x0.y = ex2_approx_b(x0.y);
                x1.x = ex2_approx_b(x1.x); x1.y = ex2_approx_b(x1.y);
                h0 = __ffma2_rn(x0, h0, __fmul2_rn(du2, make_float2(bv.x, bv.y)));
                h1 = __ffma2_rn(x1, h1, __fmul2_rn(du2, make_float2(bv.z, bv.w)));
                hist[t][0] = h0; hist[t][1] = h1;
                const float2 yp = __ffma2_rn(h1, make_float2(cv.z, cv.w), __fmul2_rn(h0, make_float2(cv.x, cv.y)));
                sm.part0[slice][t][ch] = yp.x + yp.y;
            }
        }
        __syncthreads();
        // ---- combine 1: y, dz, dy, dD
        for (int i = tid; i < SB_TC * SB_CH; i += NT) {
            const int t = i >> 5, cc = i & 31;       // cc == tid & 31 for every i of this thread
            const int cg = c0 + cc;
            float yv = 0.f;
#pragma unroll
            for (int s = 0; s < SL; ++s) yv += sm.part0[s][t][cc];
            const float uv = sm.u[t][cc];
            yv = fmaf(Dv, uv, yv);                    // Dv belongs to channel (tid & 31) == cc
            const float dov = sm.dout[t][cc];
            float dyv = dov;
            if (f.z) {
                const float zz = sm.z[t][cc];
                const float sg = sigmoidf_(zz);
                dyv = dov * zz * sg;
                if (t < tn && cg < f.d)
                    p.dz[(long long)b * p.dz_bs + (long long)(t0 + t) * p.dz_rs + cg] = dov * yv * sg * (1.f + zz * (1.f - sg));
            }
            sm.dy[t][cc] = (t < tn) ? dyv : 0.f;
            if (t < tn) accD = fmaf(dyv, uv, accD);
        }
        __syncthreads();
        // ---- reverse walk.  Steps beyond the sequence end (t >= tn, last chunk only -- the FIRST one walked, so G is still 0 there)
        // have dy = 0: every product below is an exact 0 and nothing has to be masked except the atomics' row bound.
        float* red_ptr = red_base + (long long)(t0 + SB_TC - 1) * red_rs;
#pragma unroll
        for (int t = SB_TC - 1; t >= 0; --t) {
            const float dyv = sm.dy[t][ch], dlv = sm.dl[t][ch], duv = sm.duv[t][ch];
            const float2 dy2 = make_float2(dyv, dyv), dl2 = make_float2(dlv, dlv), du2 = make_float2(duv, duv);
            const float4 bv = *reinterpret_cast<const float4*>(&sm.Bm[t][slice * 4]);
            const float4 cv = *reinterpret_cast<const float4*>(&sm.Cm[t][slice * 4]);
            const float2 b2[2] = {make_float2(bv.x, bv.y), make_float2(bv.z, bv.w)};
            const float2 c2[2] = {make_float2(cv.x, cv.y), make_float2(cv.z, cv.w)};
            float2 sg2 = make_float2(0.f, 0.f), sd2 = make_float2(0.f, 0.f), dC2[2], dB2[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float2 g = __ffma2_rn(dy2, c2[j], G2[j]);
                dC2[j] = __fmul2_rn(dy2, hist[t][j]);
                dB2[j] = __fmul2_rn(g, du2);
                sg2 = __ffma2_rn(g, b2[j], sg2);
                const float2 hprev = t > 0 ? hist[t > 0 ? t - 1 : 0][j] : hs2[j];
                float2 e = __fmul2_rn(dl2, a2p[j]);
                e.x = ex2_approx_b(e.x); e.y = ex2_approx_b(e.y);
                const float2 q = __fmul2_rn(__fmul2_rn(g, hprev), e);
                dA2[j] = __ffma2_rn(q, dl2, dA2[j]);
                sd2 = __ffma2_rn(q, Alnp[j], sd2);
                G2[j] = __fmul2_rn(g, e);
            }
            sm.part0[slice][t][ch] = sg2.x + sg2.y;
            sm.part1[slice][t][ch] = sd2.x + sd2.y;
            // d-reductions of dB / dC over the warp's 32 channels: 8 values, 9 shuffles (warp_sum8)
            const float v8[8] = {dC2[0].x, dC2[0].y, dC2[1].x, dC2[1].y, dB2[0].x, dB2[0].y, dB2[1].x, dB2[1].y};
            const float tot = warp_sum8(v8, ch);
            if (red_lane && t < tn) atomicAdd(red_ptr, tot);
            red_ptr -= red_rs;
        }
        __syncthreads();
        // ---- combine 2: du, ddelta, dbias
        for (int i = tid; i < SB_TC * SB_CH; i += NT) {
            const int t = i >> 5, cc = i & 31;
            const int cg = c0 + cc;
            if (t >= tn || cg >= f.d) continue;
            float sgB = 0.f, sdl = 0.f;
#pragma unroll
            for (int s = 0; s < SL; ++s) {
                sgB += sm.part0[s][t][cc];
                sdl += sm.part1[s][t][cc];
            }
            const float uv = sm.u[t][cc], dlv = sm.dl[t][cc], dyv = sm.dy[t][cc];
            const float duv = fmaf(dyv, Dv, dlv * sgB);
            float ddl = fmaf(uv, sgB, sdl);
            if (f.delta_softplus) {
                const float dr = sm.draw[t][cc];
                ddl *= dr > 20.0f ? 1.0f : sigmoidf_(dr);
            }
            p.du[(long long)b * p.du_bs + (long long)(t0 + t) * p.du_rs + cg] = duv;
            p.ddelta[(long long)b * p.ddl_bs + (long long)(t0 + t) * p.ddl_rs + cg] = ddl;
            accBias += ddl;
        }
        __syncthreads();
    }
    // ---- parameter gradients
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float dv[2] = {dA2[j].x * Alnp[j].x, dA2[j].y * Alnp[j].y};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int n = slice * 4 + 2 * j + e;
            if (c_ok && n < f.n_state && p.dA_log) atomicAdd(p.dA_log + (long long)c * f.n_state + n, dv[e]);
        }
    }
    if (c_ok) {
        if (p.dD) atomicAdd(p.dD + c, accD);
        if (p.ddelta_bias) atomicAdd(p.ddelta_bias + c, accBias);
    }
}

template <int SL>
static int launch_scan_bwd(const cum_scan_bwd_desc& d, cudaStream_t st) {
    auto kern = selective_scan_bwd_kernel<SL>;
    const size_t smem = sizeof(ScanBwdSmem<SL>);
    { const int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(kern), (int)smem, "cudaFuncSetAttribute(selective_scan_bwd_kernel)"); if (rc_attr) return rc_attr; }
    dim3 grid((unsigned)cdiv(d.fwd.d, SB_CH), (unsigned)d.fwd.batch);
    kern<<<grid, 32 * SL, smem, st>>>(d);
    CUM_LAUNCH_CHECK("selective_scan_bwd_kernel");
    return CUM_OK;
}

int selective_scan_bwd(const cum_scan_bwd_desc& d, cudaStream_t st) {
    const cum_scan_desc& f = d.fwd;
    CUM_REQUIRE(f.u && f.delta && f.Bm && f.Cm && f.a2, "selective_scan_bwd: null forward operand");
    CUM_REQUIRE(d.dout && d.h_ckpt && d.du && d.ddelta && d.dB && d.dC, "selective_scan_bwd: null gradient pointer");
    CUM_REQUIRE(!f.z || d.dz, "selective_scan_bwd: dz required when z is given");
    CUM_REQUIRE(f.batch > 0 && f.batch <= 65535 && f.len > 0 && f.d > 0 && f.n_state > 0 && f.n_state <= 64, "selective_scan_bwd: bad shape");
    CUM_REQUIRE(f.n_state % 4 == 0, "selective_scan_bwd: n_state must be a multiple of 4 (padded layout)");
    if (f.n_state > 32) return launch_scan_bwd<16>(d, st);
    if (f.n_state > 16) return launch_scan_bwd<8>(d, st);
    if (f.n_state > 8)  return launch_scan_bwd<4>(d, st);
    return launch_scan_bwd<2>(d, st);
}

}  // namespace cum
