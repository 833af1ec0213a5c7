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
    float u[SB_TC][SB_CH], dl[SB_TC][SB_CH], draw[SB_TC][SB_CH], z[SB_TC][SB_CH], dout[SB_TC][SB_CH], dy[SB_TC][SB_CH];
    float Bm[SB_TC][NP], Cm[SB_TC][NP];
    float part0[SL][SB_TC][SB_CH];      // y partials, then sum_n g B partials
    float part1[SL][SB_TC][SB_CH];      // sum_n g h_prev A e partials
};

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

    float a2[4], Aln[4], G[4], dA[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = slice * 4 + i;
        const bool ok = c_ok && n < f.n_state;
        a2[i] = ok ? f.a2[(long long)c * f.n_state + n] : 0.f;
        Aln[i] = a2[i] * LN2F;           // A = a2 / log2(e)
        G[i] = 0.f;
        dA[i] = 0.f;
    }
    float accD = 0.f, accBias = 0.f;     // per (tid & 31) channel partials of the combine threads
    const float Dv = (f.Dskip && c_ok) ? f.Dskip[c] : 0.f;

    // register prefetch buffers: every thread owns a fixed share of each tile
    constexpr int PF_E = (SB_TC * SB_CH + NT - 1) / NT;      // u / delta / z / dout elements per thread
    constexpr int PF_N = (SB_TC * NP + NT - 1) / NT;         // B / C elements per thread
    float pf_u[PF_E], pf_d[PF_E], pf_z[PF_E], pf_o[PF_E], pf_B[PF_N], pf_C[PF_N];
    auto prefetch = [&](int chunk) {
        const int t0 = chunk * SB_TC;
        const int tn = min(SB_TC, f.len - t0);
        int e = 0;
        for (int i = tid; i < SB_TC * SB_CH; i += NT, ++e) {
            const int t = i >> 5, cc = i & 31;
            const int tt = t0 + t, cg = c0 + cc;
            const bool ok = t < tn && cg < f.d;
            pf_u[e] = ok ? f.u[(long long)b * f.u_bs + (long long)tt * f.u_rs + cg] : 0.f;
            pf_d[e] = ok ? f.delta[(long long)b * f.dl_bs + (long long)tt * f.dl_rs + cg] : 0.f;
            pf_z[e] = (ok && f.z) ? f.z[(long long)b * f.z_bs + (long long)tt * f.z_rs + cg] : 0.f;
            pf_o[e] = ok ? p.dout[(long long)b * p.dout_bs + (long long)tt * p.dout_rs + cg] : 0.f;
        }
        e = 0;
        for (int i = tid; i < SB_TC * NP; i += NT, ++e) {
            const int t = i / NP, n = i - t * NP;
            const bool ok = t < tn && n < f.n_state;
            pf_B[e] = ok ? f.Bm[(long long)b * f.B_bs + (long long)(t0 + t) * f.B_rs + n] : 0.f;
            pf_C[e] = ok ? f.Cm[(long long)b * f.C_bs + (long long)(t0 + t) * f.C_rs + n] : 0.f;
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
                sm.u[t][cc] = pf_u[e];
                sm.draw[t][cc] = dr;
                sm.dl[t][cc] = f.delta_softplus ? softplus_b(dr) : dr;
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
        // ---- forward recompute with history in registers
        float hs[4], hist[SB_TC][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = slice * 4 + i;
            hs[i] = (c_ok && n < f.n_state) ? p.h_ckpt[(((long long)b * nchunks + chunk) * f.d + c) * f.n_state + n] : 0.f;
        }
        {
            float h[4] = {hs[0], hs[1], hs[2], hs[3]};
#pragma unroll
            for (int t = 0; t < SB_TC; ++t) {
                const float dlv = sm.dl[t][ch];
                const float duv = dlv * sm.u[t][ch];
                const float4 bv = *reinterpret_cast<const float4*>(&sm.Bm[t][slice * 4]);
                const float4 cv = *reinterpret_cast<const float4*>(&sm.Cm[t][slice * 4]);
                h[0] = fmaf(ex2_approx_b(dlv * a2[0]), h[0], duv * bv.x);
                h[1] = fmaf(ex2_approx_b(dlv * a2[1]), h[1], duv * bv.y);
                h[2] = fmaf(ex2_approx_b(dlv * a2[2]), h[2], duv * bv.z);
                h[3] = fmaf(ex2_approx_b(dlv * a2[3]), h[3], duv * bv.w);
                hist[t][0] = h[0]; hist[t][1] = h[1]; hist[t][2] = h[2]; hist[t][3] = h[3];
                sm.part0[slice][t][ch] = (h[0] * cv.x + h[1] * cv.y) + (h[2] * cv.z + h[3] * cv.w);
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
        // ---- reverse walk
#pragma unroll
        for (int t = SB_TC - 1; t >= 0; --t) {
            const float dyv = sm.dy[t][ch];
            const float dlv = sm.dl[t][ch];
            const float duv = dlv * sm.u[t][ch];
            const float4 bv = *reinterpret_cast<const float4*>(&sm.Bm[t][slice * 4]);
            const float4 cv = *reinterpret_cast<const float4*>(&sm.Cm[t][slice * 4]);
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w}, cc4[4] = {cv.x, cv.y, cv.z, cv.w};
            float sgB = 0.f, sdl = 0.f, dCp[4], dBp[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float g = fmaf(dyv, cc4[i], G[i]);
                dCp[i] = dyv * hist[t][i];
                dBp[i] = g * duv;
                sgB = fmaf(g, bb[i], sgB);
                const float hprev = t > 0 ? hist[t > 0 ? t - 1 : 0][i] : hs[i];
                const float e = ex2_approx_b(dlv * a2[i]);
                const float de = g * hprev;
                dA[i] = fmaf(de * dlv, e, dA[i]);
                sdl = fmaf(de * e, Aln[i], sdl);
                G[i] = t < tn ? g * e : G[i];        // steps beyond the sequence end carry nothing
            }
            sm.part0[slice][t][ch] = sgB;
            sm.part1[slice][t][ch] = sdl;
            // d-reductions of dB / dC over the warp's 32 channels: 8 values, 9 shuffles (warp_sum8)
            {
                const float v8[8] = {dCp[0], dCp[1], dCp[2], dCp[3], dBp[0], dBp[1], dBp[2], dBp[3]};
                const float tot = warp_sum8(v8, ch);
                const int idx = ((ch >> 4) & 1) * 4 + ((ch >> 3) & 1) * 2 + ((ch >> 2) & 1);    // value this lane holds
                const int n = slice * 4 + (idx & 3);
                if ((ch & 3) == 0 && t < tn && n < f.n_state) {
                    float* dst = idx < 4 ? p.dC + (long long)b * p.dC_bs + (long long)(t0 + t) * p.dC_rs
                                         : p.dB + (long long)b * p.dB_bs + (long long)(t0 + t) * p.dB_rs;
                    atomicAdd(dst + n, tot);
                }
            }
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
    for (int i = 0; i < 4; ++i) {
        const int n = slice * 4 + i;
        if (c_ok && n < f.n_state && p.dA_log) atomicAdd(p.dA_log + (long long)c * f.n_state + n, dA[i] * Aln[i]);
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
