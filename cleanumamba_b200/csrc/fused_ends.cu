// The two waveform-end blocks of the U-Net, each as ONE persistent tcgen05 kernel (f16x3 arithmetic, 64-channel geometry):
//
//   enc0_block      waveform -> [F.pad + Conv1d(1,64,4,2) + ReLU] -> [Conv1d(64,128,1) + GLU] -> skip0 (B, L1, 64)
//                   (/root/reference/src/network/CleanUMamba.py:108-113, :263-269)
//   dec_last_block  x (B, T, 64) -> [Conv1d(64,128,1) + GLU] -> [ConvTranspose1d(64,1,4,2)] -> crop to L, * std -> waveform
//                   (:121-128, :313-319)
//
// Unfused, each block wrote its 64-channel intermediate (1.3 GB per 64 x 10 s batch) to HBM and read it back; both blocks are
// HBM-bound (the 1x1 GEMM is 16 kFLOP per 256-byte row), so the round trip doubled their time.  Here the intermediate never
// leaves the SM:
//   * enc0: four "A generator" warps compute the Cin = 1 strided conv + ReLU on the CUDA cores straight from the waveform,
//     split each value into fp16 hi / lo halves and write the two K-major 128-byte-swizzled MMA operand tiles; one thread
//     issues the 12 tcgen05 MMAs of the 128 x 128 x 64 tile (a_hi w_hi + a_lo w_hi + a_hi w_lo) into TMEM; eight epilogue
//     warps apply bias + GLU and stage the 128 x 64 fp32 tile in swizzled shared memory for a TMA store.
//   * dec_last: the A generators load the fp32 input rows (coalesced 512-byte requests) and split them; the epilogue
//     warps contract the 64 gated channels with the four transposed-conv taps in registers (thread = row), exchange the
//     partial sums through shared memory, overlap-add neighbouring rows and write the cropped, re-scaled waveform.
// A tile, TMEM accumulator and staging tile are double-buffered: generator, MMA and epilogue of consecutive tiles overlap.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <string.h>

namespace cum {

constexpr int FE_ROWS = 128;                    // rows per tile (MMA M)
constexpr int FE_K = 64;                        // channels in (MMA K: one 128-byte swizzle row of fp16)
constexpr int FE_N = 128;                       // GLU GEMM columns (interleaved a_c, b_c)
// warp 0: MMA + TMEM; warp 1: TMA (weights, dec_last input tiles); warps 4-7 / 8-11: epilogue set 0 / 1 (alternating tiles);
// warps 12-19: A generators.  Every stage of a tile is latency-bound on its own (TMEM load -> gate -> staging -> store), so
// the throughput comes from running two epilogue sets and a double-buffered generator concurrently, not from wider warps.
constexpr int FE_THREADS = 640;
constexpr int FE_GEN_WARP0 = 12, FE_GEN_THREADS = 256;
constexpr uint32_t FE_W_BYTES = FE_N * FE_K * 2;          // 16 KB per weight half
constexpr uint32_t FE_A_BYTES = FE_ROWS * FE_K * 2;       // 16 KB per activation half
constexpr uint32_t FE_OUT_BYTES = FE_ROWS * 64 * 4;       // 32 KB fp32 tile: enc0 staging per epilogue set / dec_last raw input per buffer
constexpr uint32_t FE_OFF_W = 0;
constexpr uint32_t FE_OFF_A = 2 * FE_W_BYTES;                          // [buf][hi | lo]
constexpr uint32_t FE_OFF_OUT = FE_OFF_A + 2 * 2 * FE_A_BYTES;         // [buf]
constexpr uint32_t FE_OFF_SMALL = FE_OFF_OUT + 2 * FE_OUT_BYTES;       // bias (128 f) | taps (64 x float4) | b0 (64 f) | part (2 x 128 float4)
constexpr uint32_t FE_SMALL_BYTES = 512 + 1024 + 256 + 2 * 128 * 16;
constexpr uint32_t FE_OFF_BAR = FE_OFF_SMALL + FE_SMALL_BYTES;
constexpr uint32_t FE_SMEM_BYTES = FE_OFF_BAR + 128 + 1024 /*align slack*/;
constexpr uint32_t FE_TMEM_COLS = 256;          // 2 x 128 accumulator columns

struct FusedEndParams {
    // enc0: waveform in; dec_last: waveform out
    const float* x; long long x_stride; int length;
    int rows;                   // rows per clip: enc0 output rows L1; dec_last input rows T
    int batch, tiles_per_clip, total_tiles;
    const float* taps;          // (4, 64) taps-major: enc0 conv weights / dec_last transposed-conv weights
    const float* b0;            // enc0: (64) conv bias
    const float* bias;          // (128) interleaved GLU bias
    float acc_scale;            // 2^-k undoing the fp16 weight pre-scale
    int skip_wlo;               // low weight half is exactly zero: two MMA passes
    int c_in, n_out;            // padded channel counts actually present (<= 64 / <= 128): the rest of the 64 x 128 tile is zero
    float out_bias; const float* scale; float* out; long long out_stride; int out_length;
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void set_bar(int set) { asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory"); }

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
// tcgen05.ld is asynchronous: its destination registers are only valid after tcgen05.wait::ld.  The wait carries the 32 registers as
// in/out operands so that the compiler cannot schedule any use of them (or a copy) above it -- a plain `asm volatile` wait orders only
// against other volatile asm statements, not against the arithmetic that consumes the registers.
__device__ __forceinline__ void tmem_ld_wait(float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}

// fp32 x8 -> fp16 hi (saturating) and lo halves, packed for one 16-byte chunk of a K-major operand row
__device__ __forceinline__ void split8(const float* f, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        h[e] = cvt_f16x2_sat(f[2 * e], f[2 * e + 1]);
        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[e]));
        l[e] = cvt_f16x2_sat(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// KIND 0 = enc0_block, 1 = dec_last_block
template <int KIND>
__global__ void __launch_bounds__(FE_THREADS, 1)
fused_end_kernel(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
                 const __grid_constant__ CUtensorMap tmIO, const FusedEndParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    float* bias_s = reinterpret_cast<float*>(smem_gen + FE_OFF_SMALL);
    float4* taps_s = reinterpret_cast<float4*>(smem_gen + FE_OFF_SMALL + 512);
    float* b0_s = reinterpret_cast<float*>(smem_gen + FE_OFF_SMALL + 512 + 1024);
    float4* part_s = reinterpret_cast<float4*>(smem_gen + FE_OFF_SMALL + 512 + 1024 + 256);      // [set][row]
    const uint32_t bar_base = smem_base + FE_OFF_BAR;
    auto a_full = [&](int b) { return bar_base + 8u * b; };
    auto a_empty = [&](int b) { return bar_base + 8u * (2 + b); };
    auto acc_full = [&](int b) { return bar_base + 8u * (4 + b); };
    auto acc_empty = [&](int b) { return bar_base + 8u * (6 + b); };
    auto raw_full = [&](int b) { return bar_base + 8u * (8 + b); };
    auto raw_empty = [&](int b) { return bar_base + 8u * (10 + b); };
    const uint32_t w_full = bar_base + 96;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + FE_OFF_BAR + 112);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(a_full(b), FE_GEN_THREADS);
            mbar_init(a_empty(b), 1);
            mbar_init(acc_full(b), 1);
            mbar_init(acc_empty(b), 4);
            mbar_init(raw_full(b), 1);
            mbar_init(raw_empty(b), FE_GEN_THREADS);
        }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(FE_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    pdl_trigger();
    // small operands: GLU bias, the four taps per channel as one float4, conv bias (parameters: never written by a kernel of the step)
    // (narrower layers -- pruned checkpoints -- are zero-padded to the 64 x 128 tile: zero taps / biases here, zero-filled TMA boxes)
    for (int i = threadIdx.x; i < FE_N; i += FE_THREADS) bias_s[i] = i < p.n_out ? __ldg(p.bias + i) : 0.f;
    {
        const int ct = (KIND == 0) ? p.c_in : p.n_out / 2;          // channels the taps apply to: conv outputs / gated channels
        for (int i = threadIdx.x; i < 64; i += FE_THREADS) {
            taps_s[i] = i < ct ? make_float4(__ldg(p.taps + i), __ldg(p.taps + ct + i), __ldg(p.taps + 2 * ct + i), __ldg(p.taps + 3 * ct + i))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            b0_s[i] = (KIND == 0 && i < ct) ? __ldg(p.b0 + i) : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();          // the previous kernel's output (the waveform / the decoder activations) is read from here on

    // tile -> (clip, first row).  dec_last tiles overlap by one row: row 0 of a tile only provides g[p-1] for row 1
    constexpr int STEP = (KIND == 0) ? FE_ROWS : FE_ROWS - 1;
    auto tile_coords = [&](int tile, int& b, int& m0) {
        b = tile / p.tiles_per_clip;
        m0 = (tile - b * p.tiles_per_clip) * STEP - (KIND == 0 ? 0 : 1);
    };

    if (warp == 1 && lane == 0) {
        // ===================================================================== TMA: weights once, dec_last input tiles two ahead
        tma_prefetch_desc(&tmWh);
        mbar_arrive_expect_tx(w_full, p.skip_wlo ? FE_W_BYTES : 2 * FE_W_BYTES);
        tma_load_3d(smem_base + FE_OFF_W, &tmWh, w_full, 0, 0, 0);
        if (!p.skip_wlo) tma_load_3d(smem_base + FE_OFF_W + FE_W_BYTES, &tmWl, w_full, 0, 0, 0);
        if (KIND == 1) {
            tma_prefetch_desc(&tmIO);
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                int b, m0;
                tile_coords(tile, b, m0);
                mbar_wait(raw_empty(buf), ((it >> 1) & 1) ^ 1u);
                mbar_arrive_expect_tx(raw_full(buf), FE_OUT_BYTES);
                // two boxes of 32 channels x 128 rows (128-byte swizzle); rows outside [0, rows) arrive as zeros
                const uint32_t dst = smem_base + FE_OFF_OUT + buf * FE_OUT_BYTES;
                tma_load_3d(dst, &tmIO, raw_full(buf), 0, m0, b);
                tma_load_3d(dst + FE_OUT_BYTES / 2, &tmIO, raw_full(buf), 32, m0, b);
            }
        }
    } else if (warp == 0 && lane == 0) {
        // ===================================================================== MMA issuer
        mbar_wait(w_full, 0);
        // c = f32 (1 << 4); a / b = f16 (0); K-major both; N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | ((uint32_t)(FE_N >> 3) << 17) | ((uint32_t)(FE_ROWS >> 4) << 24);
        const uint64_t whi = umma_desc_sw128(smem_base + FE_OFF_W), wlo = umma_desc_sw128(smem_base + FE_OFF_W + FE_W_BYTES);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            mbar_wait(acc_empty(buf), ph ^ 1u);
            mbar_wait(a_full(buf), ph);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)buf * FE_N;
            const uint64_t ahi = umma_desc_sw128(smem_base + FE_OFF_A + buf * 2 * FE_A_BYTES);
            const uint64_t alo = umma_desc_sw128(smem_base + FE_OFF_A + buf * 2 * FE_A_BYTES + FE_A_BYTES);
#pragma unroll
            for (int kk = 0; kk < FE_K / 16; ++kk) {
                const uint64_t koff = (uint64_t)(kk * 2);          // 32 bytes per k-step
                umma_bf16(tmem_d, ahi + koff, whi + koff, idesc, kk != 0);
                umma_bf16(tmem_d, alo + koff, whi + koff, idesc, 1u);
                if (!p.skip_wlo) umma_bf16(tmem_d, ahi + koff, wlo + koff, idesc, 1u);
            }
            umma_commit(a_empty(buf));
            umma_commit(acc_full(buf));
        }
    } else if (warp >= FE_GEN_WARP0) {
        // ===================================================================== A generators (8 warps)
        const int t = threadIdx.x - FE_GEN_WARP0 * 32;
        int it = 0;
        if (KIND == 0) {
            // warp = one 8-channel chunk of the operand row (its 32 conv weights + 8 biases live in REGISTERS: the kernel is bound
            // by shared-memory bandwidth -- MMA operand reads + staging -- so the generator must not add broadcast loads), lane = rows
            // l, l+32, l+64, l+96 of the tile.  y[c] = relu(b0[c] + sum_k w[k][c] x[2 row + k]); samples beyond the clip read as 0 (F.pad)
            const int j = t >> 5;
            float4 wr[8];
            float br[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { wr[e] = taps_s[8 * j + e]; br[e] = b0_s[8 * j + e]; }
            float xn[4][4];
            auto fetch_x = [&](int tile) {          // the NEXT tile's samples are requested before this tile's arithmetic
                int b, m0;
                tile_coords(tile, b, m0);
                const float* xb = p.x + (long long)b * p.x_stride;
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const long long s0 = 2ll * (m0 + lane + 32 * rr);
#pragma unroll
                    for (int k = 0; k < 4; ++k) xn[rr][k] = (s0 + k < p.length) ? __ldg(xb + s0 + k) : 0.f;
                }
            };
            if ((int)blockIdx.x < p.total_tiles) fetch_x(blockIdx.x);
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                uint8_t* ahi = smem_gen + FE_OFF_A + buf * 2 * FE_A_BYTES;
                uint8_t* alo = ahi + FE_A_BYTES;
                float xv[4][4];
#pragma unroll
                for (int rr = 0; rr < 4; ++rr)
#pragma unroll
                    for (int k = 0; k < 4; ++k) xv[rr][k] = xn[rr][k];
                if (tile + (int)gridDim.x < p.total_tiles) fetch_x(tile + gridDim.x);
                mbar_wait(a_empty(buf), ph ^ 1u);
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const int row = lane + 32 * rr;
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float acc = br[e];
                        acc = fmaf(wr[e].x, xv[rr][0], acc); acc = fmaf(wr[e].y, xv[rr][1], acc);
                        acc = fmaf(wr[e].z, xv[rr][2], acc); acc = fmaf(wr[e].w, xv[rr][3], acc);
                        f[e] = fmaxf(acc, 0.f);
                    }
                    uint4 hi, lo;
                    split8(f, hi, lo);
                    const uint32_t off = (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4);
                    *reinterpret_cast<uint4*>(ahi + off) = hi;
                    *reinterpret_cast<uint4*>(alo + off) = lo;
                }
                fence_proxy_async();
                mbar_arrive(a_full(buf));
            }
        } else {
            // thread = (row, half of the channels): raw fp32 tile (TMA, two 128-byte-swizzled halves of 32 channels) -> fp16 hi / lo tiles
            const int row = t & 127, ch = t >> 7;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                uint8_t* ahi = smem_gen + FE_OFF_A + buf * 2 * FE_A_BYTES;
                uint8_t* alo = ahi + FE_A_BYTES;
                const uint8_t* raw = smem_gen + FE_OFF_OUT + buf * FE_OUT_BYTES + ch * (FE_OUT_BYTES / 2) + row * 128;
                mbar_wait(raw_full(buf), ph);
                float4 v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const float4*>(raw + ((c ^ (row & 7)) << 4));
                mbar_wait(a_empty(buf), ph ^ 1u);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = 4 * ch + jj;
                    const float f[8] = {v[2 * jj].x, v[2 * jj].y, v[2 * jj].z, v[2 * jj].w, v[2 * jj + 1].x, v[2 * jj + 1].y, v[2 * jj + 1].z, v[2 * jj + 1].w};
                    uint4 hi, lo;
                    split8(f, hi, lo);
                    const uint32_t off = (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4);
                    *reinterpret_cast<uint4*>(ahi + off) = hi;
                    *reinterpret_cast<uint4*>(alo + off) = lo;
                }
                // The raw tile is handed back to the TMA only now: an arrive issued right behind the loads does not wait for their data
                // (shared-memory loads complete asynchronously), and the async-proxy refill overtook them -- whole tiles computed from the
                // NEXT tile's input at three or more tiles per CTA.  Here every loaded value has been consumed by the conversion above.
                mbar_arrive(raw_empty(buf));
                fence_proxy_async();
                mbar_arrive(a_full(buf));
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ===================================================================== epilogue: set = tile parity, thread = tile row, all 128 columns
        const int q = warp & 3, set = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const bool leader = (q == 0 && lane == 0);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            if ((it & 1) != set) continue;
            const int buf = set;
            const uint32_t ph = (it >> 1) & 1;
            int b, m0;
            tile_coords(tile, b, m0);
            mbar_wait(acc_full(buf), ph);
            tc_fence_after();
            const uint32_t tsrc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * FE_N);
            uint8_t* st = smem_gen + FE_OFF_OUT + buf * FE_OUT_BYTES;
            float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
            float va[32], vb[32];
            __syncwarp();
            tmem_ld32_nowait(tsrc, va);
            if (KIND == 0) {
                if (leader) bulk_wait_read<0>();        // this set's previous store has finished reading the staging tile
                set_bar(set);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float* v = (c & 1) ? vb : va;
                tmem_ld_wait(v);
                if (c < 3) tmem_ld32_nowait(tsrc + 32 * (c + 1), (c & 1) ? va : vb);
                if (c == 3) {           // the accumulator is in registers: hand the TMEM buffer back to the MMA thread
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(buf));
                }
                float o[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float4 bq;
                    if ((i & 1) == 0) bq = *reinterpret_cast<const float4*>(bias_s + 32 * c + 2 * i);       // two column pairs per load
                    const float2 bv = (i & 1) ? make_float2(bq.z, bq.w) : make_float2(bq.x, bq.y);
                    const float xa = fmaf(v[2 * i], p.acc_scale, bv.x), xb = fmaf(v[2 * i + 1], p.acc_scale, bv.y);
                    o[i] = xa * fast_sigmoid(xb);
                    if (KIND == 1) {
                        const float4 w = taps_s[16 * c + i];
                        y0 = fmaf(o[i], w.x, y0); y1 = fmaf(o[i], w.y, y1); y2 = fmaf(o[i], w.z, y2); y3 = fmaf(o[i], w.w, y3);
                    }
                }
                if (KIND == 0) {
                    // staging tile: two halves of 32 channels (128-byte rows, 128-byte swizzle like the tensor map of the store)
                    uint8_t* half = st + (c >> 1) * (FE_OUT_BYTES / 2) + r * 128;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        *reinterpret_cast<float4*>(half + ((((c & 1) * 4 + k) ^ (r & 7)) << 4)) = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
                }
            }
            if (KIND == 0) {
                fence_proxy_async();
                set_bar(set);
                if (leader) {
                    const uint32_t src = smem_base + FE_OFF_OUT + buf * FE_OUT_BYTES;
                    tma_store_3d(&tmIO, src, 0, m0, b);
                    tma_store_3d(&tmIO, src + FE_OUT_BYTES / 2, 32, m0, b);
                    bulk_commit();
                }
            } else {
                const int pr = m0 + r;
                const bool valid = pr >= 0 && pr < p.rows;          // rows outside the clip contribute nothing (their A rows are zero, GLU(bias) is not)
                float4* part = part_s + set * 128;
                set_bar(set);                                       // the previous tile of this set has been read
                part[r] = valid ? make_float4(y0, y1, y2, y3) : make_float4(0.f, 0.f, 0.f, 0.f);
                set_bar(set);
                if (r >= 1 && pr <= p.rows) {
                    // out[2p + k] = bias + <g[p], w[k]> (k = 0, 1) + <g[p-1], w[k + 2]>
                    const float4 cur = part[r], prv = part[r - 1];
                    const float sc = p.scale ? __ldg(p.scale + b) : 1.0f;
                    const float e0 = (p.out_bias + cur.x + prv.z) * sc;
                    const float e1 = (p.out_bias + cur.y + prv.w) * sc;
                    const long long s = 2ll * pr;
                    float* ob = p.out + (long long)b * p.out_stride;
                    if (s + 1 < p.out_length && ((reinterpret_cast<uintptr_t>(ob + s) & 7u) == 0)) {
                        *reinterpret_cast<float2*>(ob + s) = make_float2(e0, e1);
                    } else {
                        if (s < p.out_length) ob[s] = e0;
                        if (s + 1 < p.out_length) ob[s + 1] = e1;
                    }
                }
            }
        }
        if (KIND == 0 && leader) bulk_wait_read<0>();
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(FE_TMEM_COLS));
    }
}

// ------------------------------------------------------------------------------------------------ host side
template <int KIND>
static int launch_fused_end(const CUtensorMap& tmWh, const CUtensorMap& tmWl, const CUtensorMap& tmOut, const FusedEndParams& p,
                            cudaStream_t st) {
    auto kern = fused_end_kernel<KIND>;
    const int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), (int)FE_SMEM_BYTES, "cudaFuncSetAttribute(fused_end_kernel)");
    if (rc) return rc;
    const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
    cudaError_t e = launch_kernel(kern, dim3(grid), dim3(FE_THREADS), FE_SMEM_BYTES, st, tmWh, tmWl, tmOut, p);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(fused_end_kernel)");
    return CUM_OK;
}

// (n_out, c_in) fp16 weight halves seen through a 64 x 128 box: rows / columns beyond the real extents arrive as zeros
static int weight_maps(const void* w_hi, const void* w_lo, int c_in, int n_out, CUtensorMap* tmWh, CUtensorMap* tmWl) {
    CUM_REQUIRE(w_hi && aligned16(w_hi) && (!w_lo || aligned16(w_lo)), "fused block: weight halves must be 16-byte aligned");
    int rc = make_tensor_map(tmWh, w_hi, (uint64_t)c_in, (uint64_t)n_out, 1, (uint64_t)c_in, (uint64_t)c_in * n_out, FE_K, FE_N, "W_hi", true, true);
    if (rc) return rc;
    if (w_lo) return make_tensor_map(tmWl, w_lo, (uint64_t)c_in, (uint64_t)n_out, 1, (uint64_t)c_in, (uint64_t)c_in * n_out, FE_K, FE_N, "W_lo", true, true);
    *tmWl = *tmWh;
    return CUM_OK;
}

int enc0_block_fwd(const cum_enc0_block_desc& d, cudaStream_t st) {
    CUM_REQUIRE(d.x && d.conv_w && d.conv_b && d.glu_w_hi && d.glu_b && d.out, "enc0_block: null pointer");
    CUM_REQUIRE(d.batch > 0 && d.length > 0 && d.rows_out > 0, "enc0_block: empty problem");
    CUM_REQUIRE(d.channels > 0 && d.channels <= 64 && d.channels % 8 == 0 && d.channels_out > 0 && d.channels_out <= 64 && d.channels_out % 8 == 0,
                "enc0_block: the fused kernel serves up to 64 channels, multiples of 8 (channels=%d, channels_out=%d): use conv_in + gemm",
                d.channels, d.channels_out);
    CUM_REQUIRE(d.glu_w_lo || d.w_lo_is_zero, "enc0_block: glu_w_lo missing");
    CUM_REQUIRE(aligned16(d.out), "enc0_block: out must be 16-byte aligned");
    CUM_REQUIRE(d.acc_scale > 0.f, "enc0_block: acc_scale = 1 / (weight scale passed to cum_split_f16)");
    CUtensorMap tmWh, tmWl, tmOut;
    int rc = weight_maps(d.glu_w_hi, d.w_lo_is_zero ? nullptr : d.glu_w_lo, d.channels, 2 * d.channels_out, &tmWh, &tmWl);
    if (rc) return rc;
    const uint64_t co = (uint64_t)d.channels_out;
    rc = make_tensor_map(&tmOut, d.out, co, (uint64_t)d.rows_out, (uint64_t)d.batch, co, (uint64_t)d.rows_out * co, 32, FE_ROWS, "enc0 out");
    if (rc) return rc;
    FusedEndParams p;
    memset(&p, 0, sizeof(p));
    p.x = d.x; p.x_stride = d.x_stride; p.length = d.length; p.rows = d.rows_out; p.batch = d.batch;
    p.tiles_per_clip = (int)cdiv(d.rows_out, FE_ROWS);
    const long long total = (long long)p.tiles_per_clip * d.batch;
    CUM_REQUIRE(total < (1ll << 31), "enc0_block: too many tiles");
    p.total_tiles = (int)total;
    p.taps = d.conv_w; p.b0 = d.conv_b; p.bias = d.glu_b; p.acc_scale = d.acc_scale; p.skip_wlo = d.w_lo_is_zero ? 1 : 0;
    p.c_in = d.channels; p.n_out = 2 * d.channels_out;
    return launch_fused_end<0>(tmWh, tmWl, tmOut, p, st);
}

int dec_last_block_fwd(const cum_dec_last_block_desc& d, cudaStream_t st) {
    CUM_REQUIRE(d.a && d.glu_w_hi && d.glu_b && d.convt_w && d.out, "dec_last_block: null pointer");
    CUM_REQUIRE(d.batch > 0 && d.rows_in > 0 && d.out_length > 0, "dec_last_block: empty problem");
    CUM_REQUIRE(d.channels > 0 && d.channels <= 64 && d.channels % 8 == 0 && d.channels_gated > 0 && d.channels_gated <= 64 && d.channels_gated % 8 == 0,
                "dec_last_block: the fused kernel serves up to 64 channels, multiples of 8 (channels=%d, channels_gated=%d): use gemm + convt_out",
                d.channels, d.channels_gated);
    CUM_REQUIRE(d.glu_w_lo || d.w_lo_is_zero, "dec_last_block: glu_w_lo missing");
    CUM_REQUIRE(aligned16(d.a), "dec_last_block: a must be 16-byte aligned");
    CUM_REQUIRE(d.acc_scale > 0.f, "dec_last_block: acc_scale = 1 / (weight scale passed to cum_split_f16)");
    CUM_REQUIRE(d.out_length <= 2 * d.rows_in + 2, "dec_last_block: out_length exceeds 2 rows_in + 2");
    CUtensorMap tmWh, tmWl;
    int rc = weight_maps(d.glu_w_hi, d.w_lo_is_zero ? nullptr : d.glu_w_lo, d.channels, 2 * d.channels_gated, &tmWh, &tmWl);
    if (rc) return rc;
    FusedEndParams p;
    memset(&p, 0, sizeof(p));
    CUtensorMap tmIn;
    const uint64_t ci = (uint64_t)d.channels;
    rc = make_tensor_map(&tmIn, d.a, ci, (uint64_t)d.rows_in, (uint64_t)d.batch, ci, (uint64_t)d.rows_in * ci, 32, FE_ROWS, "dec_last in");
    if (rc) return rc;
    p.rows = d.rows_in; p.batch = d.batch;
    p.tiles_per_clip = (int)cdiv((long long)d.rows_in + 1, FE_ROWS - 1);
    const long long total = (long long)p.tiles_per_clip * d.batch;
    CUM_REQUIRE(total < (1ll << 31), "dec_last_block: too many tiles");
    p.total_tiles = (int)total;
    p.taps = d.convt_w; p.bias = d.glu_b; p.acc_scale = d.acc_scale; p.skip_wlo = d.w_lo_is_zero ? 1 : 0;
    p.c_in = d.channels; p.n_out = 2 * d.channels_gated;
    p.out_bias = d.convt_bias; p.scale = d.scale; p.out = d.out; p.out_stride = d.out_stride; p.out_length = d.out_length;
    return launch_fused_end<1>(tmWh, tmWl, tmIn, p, st);
}

}  // namespace cum
