// tcgen05 / TMEM / TMA tap-GEMM for sm_100a (every CUM_MATH_* mode except the exact-fp32 SIMT kernel).
//
//   out[b, m, :] = EPI( bias + sum_{s<taps} W_s . a[b, m + shift_s, 0:k] ) (+ addend[b, m, :])
//
// One persistent CTA per SM.  Tiles wider than 128 columns run on CTA PAIRS (cluster of 2 = the two SMs of a TPC,
// tcgen05.mma.cta_group::2): a 256 x 256 tile, each CTA stages its own 128 rows of A and HALF of the W tile, the leader issues
// the MMAs for both CTAs and multicasts the completion barriers.  Accumulators live in TMEM (2 x 256 columns, double-buffered
// so the epilogue of tile i overlaps the main loop of tile i+1); operands are staged by TMA into swizzled K-major
// shared-memory tiles (224 KB ring, 2-7 stages).  The conv taps are extra K-blocks whose A-tile is the same tensor map fetched
// at a shifted row coordinate; rows outside [0, a_rows) are zero-filled by TMA, which is exactly the conv / transposed-conv
// boundary condition.
//
// Warp roles (640 threads, 768 with the splitter):
//   warp 0        TMA producer (one lane)
//   warp 1        TMEM allocator + MMA issuer (one lane; in a pair only the leader's -- the peer's lane relays "A is split")
//   warps 4..19   epilogue: tcgen05.ld (16x256b fragments) -> bias / ReLU / GLU / skip-add -> fp32, bf16 or fp16 hi/lo planes;
//                 neighbouring lanes exchange one value so that each lane stores 8-16 contiguous bytes of one row
//   warps 20..23  (modes with fp32 activations) operand splitter: the TMA-landed fp32 A tile is split into hi / lo halves in
//                 place between TMA arrival and MMA issue
//
// Split products: fp32 activations cannot be fed to one tensor-core pass within the 1e-4 waveform tolerance (10-11 mantissa
// bits, 22 stacked layers), so every product is expanded a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (three MMAs, ~2^-21 per
// product).  W is split once at pack time; A either by the splitter warps (F16X3 / BF16X3 / TF32X3) or -- TC_F16PS -- once by
// the PRODUCING layer's epilogue, which writes fp16 hi / lo planes ("hl16") that the consumer feeds to the MMA straight from TMA.
#include "common.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <type_traits>
#include <string.h>

#include <mutex>
#include <unordered_map>

namespace cum {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                     // fp32 elements per K-block = one 128-byte swizzle row
constexpr int TC_UMMA_K = 8;                  // kind::tf32
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
constexpr uint32_t TC_TMEM_COLS = 512;
constexpr int TC_EPI_WARPS = 16;              // warps 4..19: four per TMEM lane quarter, 32-column chunks interleaved.  The epilogue is
                                              // serial-issue-bound per warp (~0.25 IPC), so its speed scales with the warp count

// MODE: 0 = single-pass TF32, 1 = TF32X3 (hi/lo fp32 tiles, 3 kind::tf32 MMAs), 2 / 3 = BF16X3 / F16X3 (hi/lo 16-bit tiles,
// 3 kind::f16 MMAs at twice the TF32 rate; F16X3 keeps 22 mantissa bits, its weights carry a power-of-two scale undone in the epilogue)
// BN = tile width (256, or 128 for narrow layers: smaller W box -> deeper pipeline for the HBM-bound layers)
constexpr int TC_TF32 = 0, TC_TF32X3 = 1, TC_BF16X3 = 2, TC_F16X3 = 3;   // F16X3: like BF16X3 with fp16 halves (11+11 bits)
constexpr int TC_BF16 = 4;   // bf16 activations straight from HBM (no splitter), one kind::f16 pass -- the reduced-precision variant
constexpr int TC_F16PS = 5;  // F16X3 arithmetic on activations that are ALREADY stored as fp16 hi / lo planes (written by the producing
                             // layer's epilogue): both halves arrive by TMA in MMA-ready form, no splitter warps, no raw fp32 tile.  In-kernel
                             // splitting re-did the conversion once per N-tile (3-6x per element) and sat on the TMA -> MMA critical path:
                             // with the splitter's work removed the K >= 256 layers ran 8-21 % faster
// CTA2: two CTAs of a cluster (an SM pair) work on ONE 256 x BN tile with tcgen05.mma.cta_group::2: each CTA stages its own 128
// rows of A and only HALF of the W tile (the MMA reads the other half from the peer's shared memory), which cuts the
// L2 -> SM operand traffic per flop by a third (f16x3) to a half (bf16) -- the measured bound of the K >= 512 layers.
template <int MODE, int BN, bool CTA2 = false> struct TcCfg {
    static constexpr bool PASS3 = MODE != TC_TF32 && MODE != TC_BF16;   // hi/lo operands, three MMA passes per product
    static constexpr bool PRES = MODE == TC_F16PS;                      // A comes pre-split from HBM
    static constexpr bool SPLIT = PASS3 && !PRES;                       // operand-splitter warps present
    static constexpr bool PLAIN16 = MODE == TC_BF16;                    // 16-bit operands, 128-byte rows = 64 elements per K-block
    static constexpr int BK = PLAIN16 ? 64 : TC_BK;                     // elements per K-block (always 128 B of A per row)
    static constexpr bool HALF = (MODE == TC_BF16X3 || MODE == TC_F16X3 || MODE == TC_F16PS);    // 16-bit hi/lo MMA operands
    static constexpr int W_ROWS = CTA2 ? BN / 2 : BN;                   // W rows (output columns) staged by one CTA
    static constexpr uint32_t W_BYTES = W_ROWS * TC_BK * (HALF ? 2 : 4);
    static constexpr uint32_t AOP_BYTES = HALF ? TC_A_BYTES / 2 : TC_A_BYTES;   // one MMA A-operand tile
    // stage layout: [A raw fp32 | A_lo (TF32X3) | W_hi | W_lo].  Both kinds of split are done IN PLACE: TF32X3 masks the raw tile
    // into A_hi; the 16-bit modes overwrite the 16 KB raw tile with its 8 KB hi + 8 KB lo tiles (all reads, a named barrier,
    // then the writes).  Smaller stages = a deeper TMA ring: the main loop is bound by bytes in flight, not by the MMA rate
    static constexpr uint32_t AHI_OFF = 0;
    static constexpr uint32_t ALO_OFF = HALF ? AOP_BYTES : TC_A_BYTES;
    static constexpr uint32_t W_OFF = (SPLIT && !HALF) ? 2 * TC_A_BYTES : TC_A_BYTES;
    static constexpr uint32_t WLO_OFF = W_OFF + W_BYTES;
    static constexpr uint32_t STAGE_BYTES = PASS3 ? WLO_OFF + W_BYTES : W_OFF + W_BYTES;
    static constexpr int STAGES = (int)(229376u / STAGE_BYTES);          // 224 KB ring (+ 1.25 KB of barriers / alignment slack <= 227 KB)
    static constexpr uint32_t TX_BYTES = TC_A_BYTES + (PASS3 ? 2 : 1) * W_BYTES;     // PRES: A_hi + A_lo = 16 KB as well
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int THREADS = (4 + TC_EPI_WARPS + (SPLIT ? 4 : 0)) * 32;   // last 4 warps = operand splitter
    static constexpr int UMMA_K = (HALF || PLAIN16) ? 16 : 8;
};

struct TcParams {
    int m, n, k, taps, shift0, shift1, batch, epi;
    int m_tiles, n_tiles, k_blocks;
    const float* bias;
    float* c; long long c_bs, c_rs;
    const float* addend; long long add_bs, add_rs;
    float acc_scale;          // F16X3: 2^-k undoing the weight pre-scale (1 otherwise)
    int skip_wlo;             // split modes: the low half of the weights is exactly zero -> skip its load and its MMA pass
    int w_k_batch_stride;     // wgrad (split-K over rows): W k-coordinate += batch * this + w_k_off
    int w_k_off;
    int mn_splits;            // MN-major wgrad: splits per clip ("batch" = clip * mn_splits + split, k_blocks K-blocks of 32 rows per split)
    const float* acc_scale_ptr;   // optional device-side factor multiplied into acc_scale (1 / gradient scale computed on the device)
    const float* a_scale_ptr;     // optional device-side factor applied to A by the fp16 operand splitter (gradient scale)
    int add_mask;             // the addend is a ReLU MASK: out = addend > 0 ? value : 0 (ReLU backward in the data-gradient epilogue)
    float* aux; long long aux_bs, aux_rs;   // optional second fp32 output (training): the value BEFORE the gate (GLU epilogues: the
                                            // (rows, n) pre-activation) or before the addend (other epilogues: same shape as c)
    void* c_lo;               // OUTF == 2: low-half plane of the output (c is the high-half plane), fp16
    const void* addend_lo;    // OUTF == 2: low-half plane of the addend
    // plane-major A (time-major streaming FIFOs, cum_gemm_desc.a_planes): tensor-map dims are (a_plane_k, streams, planes); batch item b,
    // tap shift s, K-offset ak read plane a_plane0 + a_plane_step * ((n_half ? b >> 1 : b) + s) + ak / a_plane_k at channels ak % a_plane_k
    int a_plane_k;            // 0 = off
    int a_plane0, a_plane_step;
    int n_half;               // batch item b computes output-column half (b & 1): W rows (b & 1) * n ... (the bias is shared by both halves)
};

// ------------------------------------------------------------------------------------------------ kernel
// OUTF: output (and addend) format -- 0 fp32, 1 bf16, 2 fp16 hi / lo planes ("hl16": what TC_F16PS consumes)
// MNM (weight gradients, MODE = TC_F16PS): both operands are MN-major -- C[n, k] = sum_rows dZ[row, n] A[row + shift, k] contracts
// over the ROW dimension, along which neither dZ nor A is contiguous.  Instead of transposing both into K-major copies first, the
// TMA boxes (64 channels x 32 rows of an fp16 plane, 128-byte swizzle) land in shared memory exactly as the canonical MN-major
// UMMA layout (128-byte rows along M / N, 8-row groups 1024 B apart, 64-channel slabs 4 KB apart) and the instruction descriptor
// marks A and B as transposed.
template <int MODE, int BN, int EPI, int OUTF, bool CTA2, int MNM = 0>
__global__ void __launch_bounds__(TcCfg<MODE, BN, CTA2>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl, const TcParams p) {
    using Cfg = TcCfg<MODE, BN, CTA2>;
    static_assert(!MNM || MODE == TC_F16PS || MODE == TC_F16X3, "MN-major operands: fp16 hi / lo tiles (pre-split planes, or fp32 A split in the kernel)");
    constexpr int STAGES = Cfg::STAGES;
    constexpr bool X3 = Cfg::PASS3;              // three MMA passes, W_lo exists
    constexpr bool SPL = Cfg::SPLIT;             // splitter warps between TMA and MMA
    constexpr bool PRES = Cfg::PRES;             // A_hi / A_lo come from HBM
    constexpr bool OUT16 = OUTF == 1;
    constexpr bool OUTS = OUTF == 2;
    constexpr bool BF = Cfg::HALF;               // 16-bit SPLIT operand tiles (bf16 or fp16, 64-byte swizzle)
    constexpr bool F16 = MODE == TC_F16X3 || MODE == TC_F16PS;
    constexpr bool P16 = Cfg::PLAIN16;           // plain bf16 operands (128-byte swizzle)
    constexpr bool GLU = (EPI == CUM_EPI_GLU_SIGMOID || EPI == TC_EPI_GENERIC_GLU);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    auto a_off = [](int s) { return (uint32_t)s * Cfg::STAGE_BYTES; };                      // TMA destination (fp32)
    auto ahi_off = [](int s) { return (uint32_t)s * Cfg::STAGE_BYTES + Cfg::AHI_OFF; };     // MMA operand A (hi)
    auto alo_off = [](int s) { return (uint32_t)s * Cfg::STAGE_BYTES + Cfg::ALO_OFF; };
    auto w_off = [](int s) { return (uint32_t)s * Cfg::STAGE_BYTES + Cfg::W_OFF; };
    auto wlo_off = [](int s) { return (uint32_t)s * Cfg::STAGE_BYTES + Cfg::WLO_OFF; };
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto split_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (3 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (3 * STAGES + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 8 * (3 * STAGES + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.batch * p.m_tiles * p.n_tiles;      // CTA2: tiles are 256 rows tall (p.m_tiles counts those)
    const int k_iters = p.k_blocks * p.taps;
    // CTA pair: rank 0 (the leader) issues the MMAs for both; every barrier the MMA thread waits on lives in the leader
    const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
    const int tile0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tstep = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        if (PRES) tma_prefetch_desc(&tmAl);
        tma_prefetch_desc(&tmWh);
        if (X3) tma_prefetch_desc(&tmWl);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
            // CTA2: one arrival per local splitter warp; the leader's barrier also takes one from the peer's forwarder thread
            mbar_init(split_bar(s), CTA2 ? (rank == 0 ? 5 : 4) : 128);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), CTA2 ? 2 * TC_EPI_WARPS : TC_EPI_WARPS * 32);   // CTA2: one per epilogue warp of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (CTA2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TC_TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TC_TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    pdl_trigger();
    tc_fence_before();
    if (CTA2) cluster_sync_all(); else __syncthreads();     // CTA2: the peer's barriers must be initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();          // set-up above overlapped the previous kernel's tail; from here on its results are read

    auto tile_coords = [&](int tile, int& b, int& m0, int& n0) {
        const int nb = tile % p.n_tiles;
        const int r = tile / p.n_tiles;
        const int mb = r % p.m_tiles;
        b = r / p.m_tiles;
        m0 = CTA2 ? mb * (2 * TC_BM) + (int)rank * TC_BM : mb * TC_BM;      // this CTA's 128 rows
        n0 = nb * BN;
    };
    auto tile_umma_n = [&](int n0) {
        int n_rem = p.n - n0;
        if (n_rem > BN) n_rem = BN;
        return (uint32_t)((n_rem + 15) & ~15);
    };

    if (warp == 0 && lane == 0) {
        // ===================================================================== TMA producer
        int s = 0;
        uint32_t ph = 0;
        for (int tile = tile0; tile < total_tiles; tile += tstep) {
            int b, m0, n0;
            tile_coords(tile, b, m0, n0);
            // CTA2: this CTA stages the W rows of ITS half of the tile's columns (the MMA splits N across the pair)
            const int wn = (CTA2 ? n0 + (int)rank * (int)(tile_umma_n(n0) >> 1) : n0) + ((p.n_half && (b & 1)) ? p.n : 0);
            for (int it = 0; it < k_iters; ++it) {
                const int tap = it / p.k_blocks, kb = it - tap * p.k_blocks;
                const int shift = tap == 0 ? p.shift0 : p.shift1;
                mbar_wait(empty_bar(s), ph ^ 1u);
                const bool load_lo = X3 && !p.skip_wlo;
                // bytes this stage will receive: A + W_hi (+ W_lo unless it is skipped); non-split modes have no W_lo at all
                const uint32_t tx = (X3 && !load_lo) ? Cfg::TX_BYTES - Cfg::W_BYTES : Cfg::TX_BYTES;
                const int wk = kb * Cfg::BK + b * p.w_k_batch_stride + p.w_k_off;
                // A coordinates: (K offset, row, batch plane); plane-major A moves the tap shift and the upper part of K into the plane
                int ak = kb * Cfg::BK, arow = m0 + shift, ab = b;
                if (p.a_plane_k) {
                    const int kp = ak / p.a_plane_k;
                    ab = p.a_plane0 + p.a_plane_step * ((p.n_half ? b >> 1 : b) + shift) + kp;
                    ak -= kp * p.a_plane_k;
                    arow = m0;
                }
                if constexpr (MNM != 0) {
                    // split `sp` of clip `clip`: rows [r, r + 32) of both operands (the activation side shifted by the conv tap);
                    // rows / channels outside the tensors arrive as zeros
                    const int clip = b / p.mn_splits, sp = b - clip * p.mn_splits;
                    const int r = (sp * p.k_blocks + kb) * 32;
                    // with the in-kernel splitter every CTA's bytes are counted on its own barrier (the splitter relays "ready" to
                    // the leader); without it the leader's MMA thread waits for the bytes of both CTAs directly
                    const bool remote = CTA2 && !SPL;
                    const uint32_t bar = remote ? mapa_rank(full_bar(s), 0) : full_bar(s);
                    if (!remote || rank == 0) mbar_arrive_expect_tx(full_bar(s), (remote ? 2 : 1) * tx);
                    auto ld = [&](uint32_t dst, const CUtensorMap* m, int c0, int c1) {
                        if (remote) tma_load_3d_2sm(dst, m, bar, c0, c1, clip); else tma_load_3d(dst, m, bar, c0, c1, clip);
                    };
                    if (SPL) {          // raw fp32 gradient tile: four slabs of 32 channels x 32 rows (split in place by the splitter warps)
#pragma unroll
                        for (int i = 0; i < 4; ++i) ld(smem_base + a_off(s) + i * 4096, &tmA, m0 + 32 * i, r);
                    } else {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            ld(smem_base + ahi_off(s) + i * 4096, &tmA, m0 + 64 * i, r);
                            ld(smem_base + alo_off(s) + i * 4096, &tmAl, m0 + 64 * i, r);
                        }
                    }
                    if constexpr (MNM == 2) {   // raw fp32 activation tile: slabs of 32 channels x 32 rows, split in place by the splitter warps
#pragma unroll
                        for (int i = 0; i < Cfg::W_ROWS / 32; ++i) ld(smem_base + w_off(s) + i * 4096, &tmWh, wn + 32 * i, r + shift);
                    } else {
#pragma unroll
                        for (int i = 0; i < Cfg::W_ROWS / 64; ++i) {
                            ld(smem_base + w_off(s) + i * 4096, &tmWh, wn + 64 * i, r + shift);
                            if (load_lo) ld(smem_base + wlo_off(s) + i * 4096, &tmWl, wn + 64 * i, r + shift);
                        }
                    }
                } else if (CTA2 && !SPL) {
                    // no splitter in between: the leader's MMA thread waits for the bytes of BOTH CTAs on its own barrier
                    const uint32_t lbar = mapa_rank(full_bar(s), 0);
                    if (rank == 0) mbar_arrive_expect_tx(full_bar(s), 2 * tx);
                    tma_load_3d_2sm(smem_base + a_off(s), &tmA, lbar, ak, arow, ab);
                    if (PRES) tma_load_3d_2sm(smem_base + alo_off(s), &tmAl, lbar, ak, arow, ab);
                    tma_load_3d_2sm(smem_base + w_off(s), &tmWh, lbar, wk, wn, tap);
                    if (load_lo) tma_load_3d_2sm(smem_base + wlo_off(s), &tmWl, lbar, wk, wn, tap);
                } else {
                    mbar_arrive_expect_tx(full_bar(s), tx);
                    // wgrad (split-K over rows): the split index selects a column range of ONE 2-D operand instead of a batch plane
                    if (p.w_k_batch_stride) tma_load_3d(smem_base + a_off(s), &tmA, full_bar(s), kb * Cfg::BK + b * p.w_k_batch_stride, m0, 0);
                    else tma_load_3d(smem_base + a_off(s), &tmA, full_bar(s), ak, arow, ab);
                    if (PRES) tma_load_3d(smem_base + alo_off(s), &tmAl, full_bar(s), ak, arow, ab);
                    tma_load_3d(smem_base + w_off(s), &tmWh, full_bar(s), wk, wn, tap);
                    if (load_lo) tma_load_3d(smem_base + wlo_off(s), &tmWl, full_bar(s), wk, wn, tap);
                }
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        // ===================================================================== MMA issuer (CTA2: the leader, for both CTAs)
        int s = 0;
        uint32_t ph = 0;
        int tcount = 0;
        for (int tile = tile0; tile < total_tiles; tile += tstep, ++tcount) {
            int b, m0, n0;
            tile_coords(tile, b, m0, n0);
            const int acc = tcount & 1;
            const uint32_t acc_ph = (tcount >> 1) & 1;
            const uint32_t umma_n = tile_umma_n(n0);
            // c=f32 (1<<4); a/b format 2 = tf32, 1 = bf16 (bits 7, 10); K-major both; N>>3 at bit 17, M>>4 at bit 24
            const uint32_t fmt = F16 ? 0u : ((BF || P16) ? 1u : 2u);       // 0 = f16, 1 = bf16, 2 = tf32
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((umma_n >> 3) << 17) |
                                   ((uint32_t)((CTA2 ? 2 * TC_BM : TC_BM) >> 4) << 24) |
                                   (MNM ? ((1u << 15) | (1u << 16)) : 0u);           // a_major / b_major = MN
            const uint32_t tmem_d = tmem_base + (uint32_t)acc * BN;
            if (CTA2) mbar_wait_cluster(tempty_bar(acc), acc_ph ^ 1u); else mbar_wait(tempty_bar(acc), acc_ph ^ 1u);
            tc_fence_after();
            for (int it = 0; it < k_iters; ++it) {
                if (CTA2) mbar_wait_cluster(SPL ? split_bar(s) : full_bar(s), ph); else mbar_wait(SPL ? split_bar(s) : full_bar(s), ph);
                tc_fence_after();
                const uint64_t adesc = MNM ? umma_desc_mn_sw128(smem_base + ahi_off(s), 4096u) : BF ? umma_desc_sw64(smem_base + ahi_off(s)) : umma_desc_sw128(smem_base + ahi_off(s));
                // MNM == 2: the in-kernel split leaves [hi | lo] per 64-channel group (8 KB): slabs of one half are 8 KB apart
                const uint64_t bdesc = MNM == 2 ? umma_desc_mn_sw128(smem_base + w_off(s), 8192u) : MNM ? umma_desc_mn_sw128(smem_base + w_off(s), 4096u) : BF ? umma_desc_sw64(smem_base + w_off(s)) : umma_desc_sw128(smem_base + w_off(s));
                const uint64_t alo = MNM ? umma_desc_mn_sw128(smem_base + alo_off(s), 4096u) : BF ? umma_desc_sw64(smem_base + alo_off(s)) : umma_desc_sw128(smem_base + alo_off(s));
                const uint64_t blo = MNM == 2 ? umma_desc_mn_sw128(smem_base + w_off(s) + 4096u, 8192u) : MNM ? umma_desc_mn_sw128(smem_base + wlo_off(s), 4096u) : BF ? umma_desc_sw64(smem_base + wlo_off(s)) : umma_desc_sw128(smem_base + wlo_off(s));
#pragma unroll
                for (int kk = 0; kk < Cfg::BK / Cfg::UMMA_K; ++kk) {
                    // K-major: 32 bytes per k-step in both element types; MN-major: 16 rows of 128 bytes = two 1024-byte row groups
                    const uint64_t koff = MNM ? (uint64_t)(kk * 128) : (uint64_t)(kk * 2);
                    if (P16) {
                        umma_f16_any<CTA2>(tmem_d, adesc + koff, bdesc + koff, idesc, (it | kk) != 0);
                    } else if (BF) {
                        umma_f16_any<CTA2>(tmem_d, adesc + koff, bdesc + koff, idesc, (it | kk) != 0);
                        umma_f16_any<CTA2>(tmem_d, alo + koff, bdesc + koff, idesc, 1u);
                        if (!p.skip_wlo) umma_f16_any<CTA2>(tmem_d, adesc + koff, blo + koff, idesc, 1u);
                    } else {
                        umma_tf32_any<CTA2>(tmem_d, adesc + koff, bdesc + koff, idesc, (it | kk) != 0);
                        if (X3) {
                            umma_tf32_any<CTA2>(tmem_d, alo + koff, bdesc + koff, idesc, 1u);
                            if (!p.skip_wlo) umma_tf32_any<CTA2>(tmem_d, adesc + koff, blo + koff, idesc, 1u);
                        }
                    }
                }
                // the stage is free / the accumulator is complete in BOTH CTAs once these MMAs retire
                if (CTA2) umma_commit_2sm(empty_bar(s)); else umma_commit(empty_bar(s));
                if (it == k_iters - 1) { if (CTA2) umma_commit_2sm(tfull_bar(acc)); else umma_commit(tfull_bar(acc)); }
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (CTA2 && SPL && warp == 1 && lane == 0 && rank == 1) {
        // ===================================================================== peer CTA: forward "my A tile is split" to the leader
        // A release at cluster scope costs a full memory barrier (ERRBAR); issued by the splitter warps themselves it sat on the
        // split -> MMA critical path of every K-block (pair kernel 1.7x slower than single).  The splitters now signal a local
        // barrier (cta scope) and this otherwise idle thread relays it: wait (acquire.cta) -> arrive on the leader (release.cluster)
        int s = 0;
        uint32_t ph = 0;
        const uint32_t split_leader = mapa_rank(split_bar(0), 0);
        for (int tile = tile0; tile < total_tiles; tile += tstep) {
            for (int it = 0; it < k_iters; ++it) {
                mbar_wait(split_bar(s), ph);
                mbar_arrive_cluster(split_leader + 8u * s);
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp >= 4 && warp < 4 + TC_EPI_WARPS) {
        // ===================================================================== epilogue (16 warps)
        const int q = warp & 3;                 // TMEM lane quarter this warp may touch
        const int sub = (warp - 4) >> 2;        // which of every four 32-column chunks
        const float acc_scale = p.acc_scale_ptr ? p.acc_scale * __ldg(p.acc_scale_ptr) : p.acc_scale;
        const int tq = lane & 3, tr = lane >> 2;
        int tcount = 0;
        const uint32_t tempty_leader = CTA2 ? mapa_rank(tempty_bar(0), 0) : 0u;
        for (int tile = tile0; tile < total_tiles; tile += tstep, ++tcount) {
            int b, m0, n0;
            tile_coords(tile, b, m0, n0);
            const int acc = tcount & 1;
            const uint32_t acc_ph = (tcount >> 1) & 1;
            mbar_wait(tfull_bar(acc), acc_ph);
            tc_fence_after();
            int n_rem = p.n - n0;
            if (n_rem > BN) n_rem = BN;
            // Output / addend formats: OUTF 0 fp32, 1 bf16 (addend bf16 too), 2 fp16 hi/lo planes.  The addend's format is otherwise
            // independent of the output's (a U-Net skip keeps the format its encoder layer wrote): hl16 planes when addend_lo is
            // set (warp-uniform runtime switch), else fp32.  All element offsets inside one batch plane fit 32 bits (host check).
            const bool ADDS = !OUT16 && p.addend_lo != nullptr;
            const bool has_add = p.addend != nullptr && EPI != TC_EPI_ATOMIC_ADD;
            const long long cbo = (long long)b * p.c_bs, abo = (long long)b * p.add_bs;
            // The epilogue is instruction-issue-bound on the narrow layers (~45 SASS instructions per output before this layout):
            // row offsets are computed once per tile, interior tiles take a predicate-free path, addresses are base + 32-bit offset
            auto run = [&](auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
                // byte pointers of this thread's four rows at its first column of the first chunk; bumped by one chunk stride per
                // iteration, the k-dependent part of an address is a compile-time immediate
                constexpr int CES = OUTF == 0 ? 4 : 2;                          // output element size
                const int aes = (OUT16 || ADDS) ? 2 : 4;                        // addend element size
                // this thread's four rows are 8 rows apart: ONE pointer / offset + a stride (the 80-register cap of the 768-thread
                // kernels does not leave room for four of each)
                const char* const abase = reinterpret_cast<const char*>(p.addend) + abo * aes;
                bool rok[2][2];
                const int c00 = sub * 16;
                const int row0 = m0 + q * 32 + tr;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh) rok[h][rh] = FULL || row0 + h * 16 + rh * 8 < p.m;
                // fp32 outputs (PAIR): neighbouring lanes swap one value (pair) so that each lane stores 8 (GLU) / 16 bytes of ONE k
                // group -- a row's segment of the chunk leaves in one store instruction instead of two: half the LSU wavefronts,
                // which bound the 64- / 128-channel layers.  Even lanes own the columns of k = 0, odd lanes those of k = 1
                constexpr bool PAIR = ((OUTF == 0) || (OUTS && !GLU)) && EPI != TC_EPI_ATOMIC_ADD;
                const int col = GLU ? ((n0 + c00) >> 1) + ((OUTS || PAIR) ? (tq & ~1) + 4 * (tq & 1) : tq)
                                    : n0 + c00 + (PAIR ? ((tq & 1) ? 8 + 2 * (tq - 1) : 2 * tq) : 2 * tq);
                const int acol = GLU ? ((n0 + c00) >> 1) + tq : n0 + c00 + 2 * tq;
                char* cp0 = reinterpret_cast<char*>(p.c) + (cbo + (long long)row0 * p.c_rs + col) * CES;
                unsigned ao0 = ((unsigned)row0 * (unsigned)p.add_rs + (unsigned)acol) * (unsigned)aes;
                const unsigned cs8 = 8u * (unsigned)p.c_rs * CES, as8 = 8u * (unsigned)p.add_rs * (unsigned)aes;     // bytes per 8 rows
                auto cp = [&](int h, int rh) { return cp0 + (size_t)((2 * h + rh) * cs8); };
                auto ao = [&](int h, int rh) { return ao0 + (unsigned)(2 * h + rh) * as8; };
                float* const auxb = (OUTF == 0 && p.aux) ? p.aux + (long long)b * p.aux_bs : nullptr;
                const long long lo_delta = OUTS ? reinterpret_cast<char*>(p.c_lo) - reinterpret_cast<char*>(p.c) : 0;      // hi -> lo plane
                const long long alo_delta = ADDS ? reinterpret_cast<const char*>(p.addend_lo) - reinterpret_cast<const char*>(p.addend) : 0;
                constexpr int CSTEP = (GLU ? 32 : 64) * CES;                    // bytes per chunk stride (64 accumulator columns)
                const int astep = (GLU ? 32 : 64) * aes;
                for (int c0 = c00; c0 < n_rem; c0 += 64) {
                    bool kok[2];
                    float2 bv[2];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int n = n0 + c0 + 8 * k + 2 * tq;
                        kok[k] = FULL || n < p.n;
                        bv[k] = (p.bias && kok[k]) ? __ldg(reinterpret_cast<const float2*>(p.bias + n)) : make_float2(0.f, 0.f);
                    }
                    float v[2][8];
                    __syncwarp();       // tcgen05.ld is .sync.aligned: reconverge after the predicated stores
                    tmem_ld_16x256b_x2_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v[0]);
                    tmem_ld_16x256b_x2_nowait(tmem_base + ((uint32_t)(q * 32 + 16) << 16) + (uint32_t)(acc * BN + c0), v[1]);
                    // the U-Net skip tile (addend) comes from DRAM: all of this thread's addend values are requested BEFORE
                    // waiting on the TMEM loads so their latency overlaps
                    float2 ad[2][2][2];
                    if (has_add) {
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                                for (int k = 0; k < 2; ++k) {
                                    ad[h][rh][k] = make_float2(0.f, 0.f);
                                    if (rok[h][rh] && kok[k]) {
                                        const int ko = GLU ? 4 * k : 8 * k;                 // elements
                                        if (ADDS) {     // hl16 addend: value = hi + lo
                                            const char* ah = abase + ao(h, rh) + ko * 2;
                                            if (GLU) ad[h][rh][k].x = __half2float(*reinterpret_cast<const __half*>(ah)) +
                                                                      __half2float(*reinterpret_cast<const __half*>(ah + alo_delta));
                                            else {
                                                const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(ah));
                                                const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(ah + alo_delta));
                                                ad[h][rh][k] = make_float2(fh.x + fl.x, fh.y + fl.y);
                                            }
                                        } else if (OUT16) {
                                            const char* a16 = abase + ao(h, rh) + ko * 2;
                                            if (GLU) ad[h][rh][k].x = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(a16));
                                            else ad[h][rh][k] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a16));
                                        } else if (GLU) ad[h][rh][k].x = __ldg(reinterpret_cast<const float*>(abase + ao(h, rh) + ko * 4));
                                        else ad[h][rh][k] = __ldg(reinterpret_cast<const float2*>(abase + ao(h, rh) + ko * 4));
                                    }
                                }
                    }
                    {   // the wait carries the 16 destination registers of the two loads as in/out operands: no use of them (or copy)
                        // may be scheduled above it (a bare volatile asm orders only against other volatile asm statements)
                        uint32_t* r0 = reinterpret_cast<uint32_t*>(v[0]);
                        uint32_t* r1 = reinterpret_cast<uint32_t*>(v[1]);
                        asm volatile("tcgen05.wait::ld.sync.aligned;"
                                     : "+r"(r0[0]), "+r"(r0[1]), "+r"(r0[2]), "+r"(r0[3]), "+r"(r0[4]), "+r"(r0[5]), "+r"(r0[6]), "+r"(r0[7]),
                                       "+r"(r1[0]), "+r"(r1[1]), "+r"(r1[2]), "+r"(r1[3]), "+r"(r1[4]), "+r"(r1[5]), "+r"(r1[6]), "+r"(r1[7])
                                     :: "memory");
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
#pragma unroll
                        for (int rh = 0; rh < 2; ++rh) {
                            if constexpr (OUTS && GLU) {
                                // hl16 GLU output: every lane computes its two gated values, neighbouring lanes swap one so that each
                                // stores ONE packed pair (even lane: the columns of k = 0, odd lane: k = 1) to the hi and the lo plane.
                                // No early exit before the shuffle: all 32 lanes take part
                                float o[2];
#pragma unroll
                                for (int k = 0; k < 2; ++k) {
                                    const float x0 = fmaf(v[h][4 * k + 2 * rh + 0], acc_scale, bv[k].x);
                                    const float x1 = fmaf(v[h][4 * k + 2 * rh + 1], acc_scale, bv[k].y);
                                    o[k] = x0 * tc_gate<EPI>(p.epi, x1);
                                    if (has_add) o[k] += ad[h][rh][k].x;
                                }
                                const bool odd = tq & 1;
                                const float got = __shfl_xor_sync(0xffffffffu, odd ? o[0] : o[1], 1);
                                const float p0 = odd ? got : o[0], p1 = odd ? o[1] : got;
                                if (rok[h][rh] && (FULL || n0 + c0 + 8 * (int)odd < p.n)) {
                                    const uint32_t hi2 = cvt_f16x2_sat(p0, p1);
                                    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi2));
                                    const uint32_t lo2 = cvt_f16x2_sat(p0 - hf.x, p1 - hf.y);
                                    *reinterpret_cast<uint32_t*>(cp(h, rh)) = hi2;
                                    *reinterpret_cast<uint32_t*>(cp(h, rh) + lo_delta) = lo2;
                                }
                                continue;
                            }
                            if constexpr (PAIR) {
                                const bool odd = tq & 1;
                                const bool ok = rok[h][rh] && (FULL || n0 + c0 + 8 * (int)odd < p.n);
                                if constexpr (GLU) {
                                    float o[2];
#pragma unroll
                                    for (int k = 0; k < 2; ++k) {
                                        const float x0 = F16 ? fmaf(v[h][4 * k + 2 * rh + 0], acc_scale, bv[k].x) : v[h][4 * k + 2 * rh + 0] + bv[k].x;
                                        const float x1 = F16 ? fmaf(v[h][4 * k + 2 * rh + 1], acc_scale, bv[k].y) : v[h][4 * k + 2 * rh + 1] + bv[k].y;
                                        o[k] = x0 * tc_gate<EPI>(p.epi, x1);
                                        if (has_add) o[k] += ad[h][rh][k].x;
                                        if (OUTF == 0 && auxb && rok[h][rh] && kok[k])       // training: the pre-activation pair is saved for the GLU backward
                                            *reinterpret_cast<float2*>(auxb + (long long)(row0 + (2 * h + rh) * 8) * p.aux_rs + (n0 + c0 + 8 * k + 2 * tq)) = make_float2(x0, x1);
                                    }
                                    const float got = __shfl_xor_sync(0xffffffffu, odd ? o[0] : o[1], 1);
                                    if (ok) *reinterpret_cast<float2*>(cp(h, rh)) = odd ? make_float2(got, o[1]) : make_float2(o[0], got);
                                } else {
                                    float2 o[2];
#pragma unroll
                                    for (int k = 0; k < 2; ++k) {
                                        const float x0 = F16 ? fmaf(v[h][4 * k + 2 * rh + 0], acc_scale, bv[k].x) : v[h][4 * k + 2 * rh + 0] + bv[k].x;
                                        const float x1 = F16 ? fmaf(v[h][4 * k + 2 * rh + 1], acc_scale, bv[k].y) : v[h][4 * k + 2 * rh + 1] + bv[k].y;
                                        o[k] = make_float2(tc_act<EPI>(p.epi, x0), tc_act<EPI>(p.epi, x1));
                                        if (OUTF == 0 && auxb && rok[h][rh] && kok[k])       // training: the value before the skip add (its sign is the ReLU mask)
                                            *reinterpret_cast<float2*>(auxb + (long long)(row0 + (2 * h + rh) * 8) * p.aux_rs + (n0 + c0 + 8 * k + 2 * tq)) = o[k];
                                        if (has_add) {
                                            if (p.add_mask) { o[k].x = ad[h][rh][k].x > 0.f ? o[k].x : 0.f; o[k].y = ad[h][rh][k].y > 0.f ? o[k].y : 0.f; }
                                            else { o[k].x += ad[h][rh][k].x; o[k].y += ad[h][rh][k].y; }
                                        }
                                    }
                                    const float2 snd = odd ? o[0] : o[1];
                                    const float gx = __shfl_xor_sync(0xffffffffu, snd.x, 1), gy = __shfl_xor_sync(0xffffffffu, snd.y, 1);
                                    const float4 w = odd ? make_float4(gx, gy, o[1].x, o[1].y) : make_float4(o[0].x, o[0].y, gx, gy);
                                    if constexpr (OUTS) {       // hl16 planes: 4 consecutive columns = 8 bytes per plane
                                        const uint32_t h0 = cvt_f16x2_sat(w.x, w.y), h1 = cvt_f16x2_sat(w.z, w.w);
                                        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h0));
                                        const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
                                        const uint32_t l0 = cvt_f16x2_sat(w.x - f0.x, w.y - f0.y), l1 = cvt_f16x2_sat(w.z - f1.x, w.w - f1.y);
                                        if (ok) {
                                            *reinterpret_cast<uint2*>(cp(h, rh)) = make_uint2(h0, h1);
                                            *reinterpret_cast<uint2*>(cp(h, rh) + lo_delta) = make_uint2(l0, l1);
                                        }
                                    } else {
                                        if (ok) *reinterpret_cast<float4*>(cp(h, rh)) = w;
                                    }
                                }
                                continue;
                            }
                            if (!rok[h][rh]) continue;
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                if (!kok[k]) break;
                                char* const dst = cp(h, rh) + (GLU ? 4 * k : 8 * k) * CES;
                                const float x0 = F16 ? fmaf(v[h][4 * k + 2 * rh + 0], acc_scale, bv[k].x) : v[h][4 * k + 2 * rh + 0] + bv[k].x;
                                const float x1 = F16 ? fmaf(v[h][4 * k + 2 * rh + 1], acc_scale, bv[k].y) : v[h][4 * k + 2 * rh + 1] + bv[k].y;
                                if (EPI == TC_EPI_ATOMIC_ADD) {
                                    atomicAdd(reinterpret_cast<float*>(dst), x0);
                                    atomicAdd(reinterpret_cast<float*>(dst) + 1, x1);
                                } else if (GLU) {
                                    float o = x0 * tc_gate<EPI>(p.epi, x1);
                                    if (has_add) o += ad[h][rh][k].x;
                                    if (OUT16) *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(o);
                                    else *reinterpret_cast<float*>(dst) = o;
                                } else {
                                    float2 o = make_float2(tc_act<EPI>(p.epi, x0), tc_act<EPI>(p.epi, x1));
                                    if (has_add) { o.x += ad[h][rh][k].x; o.y += ad[h][rh][k].y; }
                                    if (OUTS) {
                                        const uint32_t hi2 = cvt_f16x2_sat(o.x, o.y);
                                        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi2));
                                        const uint32_t lo2 = cvt_f16x2_sat(o.x - hf.x, o.y - hf.y);
                                        *reinterpret_cast<uint32_t*>(dst) = hi2;
                                        *reinterpret_cast<uint32_t*>(dst + lo_delta) = lo2;
                                    } else if (OUT16) {
                                        *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(o.x, o.y);
                                    } else {
                                        *reinterpret_cast<float2*>(dst) = o;
                                    }
                                }
                            }
                        }
                    }
                    cp0 += CSTEP;
                    ao0 += (unsigned)astep;
                }
            };
            if (m0 + TC_BM <= p.m && n0 + BN <= p.n) run(std::true_type{}); else run(std::false_type{});
            tc_fence_before();
            if (CTA2) {
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_leader + 8u * acc);
            } else {
                mbar_arrive(tempty_bar(acc));
            }
        }
    } else if (SPL && warp >= 4 + TC_EPI_WARPS) {
        // ===================================================================== operand splitter (A tile)
        const int t = threadIdx.x - (4 + TC_EPI_WARPS) * 32;
        const float asc = (F16 && p.a_scale_ptr) ? __ldg(p.a_scale_ptr) : 1.0f;     // power-of-two gradient scale (dgrad in f16x3)
        int s = 0;
        uint32_t ph = 0;
        for (int tile = tile0; tile < total_tiles; tile += tstep) {
            for (int it = 0; it < k_iters; ++it) {
                mbar_wait(full_bar(s), ph);
                if constexpr (MNM != 0) {
                    // MN-major: raw = four slabs (32 channels x 32 rows, 128-byte rows of fp32, 128B swizzle) -> hi / lo = two slabs each
                    // (64 channels x 32 rows, 128-byte rows of fp16, 128B swizzle).  Item = one 16-byte fp16 chunk = 8 channels of one row
                    // = two raw 16-byte chunks; 512 items, four per thread, all reads before the writes (in-place overwrite).
                    const uint8_t* raw = smem_gen + a_off(s);
                    uint8_t* hi = smem_gen + ahi_off(s);
                    uint8_t* lo = smem_gen + alo_off(s);
                    float4 va[4], vb[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = t + 128 * j;
                        const int r = i >> 4, qg = i & 15;              // row, fp16 chunk index over the 128 channels
                        const uint8_t* src = raw + (qg >> 2) * 4096 + r * 128;
                        va[j] = *reinterpret_cast<const float4*>(src + (((2 * (qg & 3)) ^ (r & 7)) << 4));
                        vb[j] = *reinterpret_cast<const float4*>(src + (((2 * (qg & 3) + 1) ^ (r & 7)) << 4));
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");     // the four splitter warps only
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = t + 128 * j;
                        const int r = i >> 4, qg = i & 15;
                        const float4 v0 = va[j], v1 = vb[j];
                        const float f[8] = {v0.x * asc, v0.y * asc, v0.z * asc, v0.w * asc, v1.x * asc, v1.y * asc, v1.z * asc, v1.w * asc};
                        uint32_t h[4], l[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            h[e] = cvt_f16x2_sat(f[2 * e], f[2 * e + 1]);
                            const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[e]));
                            l[e] = cvt_f16x2_sat(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
                        }
                        const uint32_t off = (uint32_t)(qg >> 3) * 4096u + (uint32_t)r * 128u + (uint32_t)(((qg & 7) ^ (r & 7)) << 4);
                        *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
                    }
                    if constexpr (MNM == 2) {
                        // the activation tile, 64 channels (two raw slabs = 8 KB) at a time -> [hi slab | lo slab] in the same 8 KB;
                        // 256 items of 8 channels per group, two per thread, reads / barrier / writes
                        uint8_t* wb = smem_gen + w_off(s);
#pragma unroll 1
                        for (int g = 0; g < Cfg::W_ROWS / 64; ++g) {
                            float4 wa[2], wc[2];
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const int i = t + 128 * j;
                                const int r = i >> 3, q = i & 7;
                                const uint8_t* src = wb + (2 * g + (q >> 2)) * 4096 + r * 128;
                                wa[j] = *reinterpret_cast<const float4*>(src + (((2 * (q & 3)) ^ (r & 7)) << 4));
                                wc[j] = *reinterpret_cast<const float4*>(src + (((2 * (q & 3) + 1) ^ (r & 7)) << 4));
                            }
                            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const int i = t + 128 * j;
                                const int r = i >> 3, q = i & 7;
                                const float f[8] = {wa[j].x, wa[j].y, wa[j].z, wa[j].w, wc[j].x, wc[j].y, wc[j].z, wc[j].w};
                                uint32_t h[4], l[4];
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    h[e] = cvt_f16x2_sat(f[2 * e], f[2 * e + 1]);
                                    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[e]));
                                    l[e] = cvt_f16x2_sat(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
                                }
                                uint8_t* dst = wb + g * 8192 + r * 128 + ((q ^ (r & 7)) << 4);
                                *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
                                *reinterpret_cast<uint4*>(dst + 4096) = make_uint4(l[0], l[1], l[2], l[3]);
                            }
                        }
                    }
                } else if (BF) {
                    // fp32 tile (128 x 32, 128B-swizzled rows) -> bf16 hi / lo tiles (128 x 32, 64B-swizzled rows)
                    const uint8_t* raw = smem_gen + a_off(s);
                    uint8_t* hi = smem_gen + ahi_off(s);
                    uint8_t* lo = smem_gen + alo_off(s);
                    // hi / lo overwrite the raw tile: every splitter thread first pulls its 4 x 32 bytes into registers
                    float4 va[4], vb[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = t + 128 * j;
                        const int r = i >> 2, co = i & 3;              // row, 16-byte output chunk (8 bf16 = 8 k)
                        va[j] = *reinterpret_cast<const float4*>(raw + r * 128 + (((2 * co) ^ (r & 7)) << 4));
                        vb[j] = *reinterpret_cast<const float4*>(raw + r * 128 + (((2 * co + 1) ^ (r & 7)) << 4));
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");     // the four splitter warps only
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = t + 128 * j;
                        const int r = i >> 2, co = i & 3;
                        const float4 v0 = va[j], v1 = vb[j];
                        const float f[8] = {v0.x * asc, v0.y * asc, v0.z * asc, v0.w * asc, v1.x * asc, v1.y * asc, v1.z * asc, v1.w * asc};
                        uint32_t h[4], l[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (F16) {
                                // saturating conversion (cvt.rn.satfinite): |a| > 65504 clamps instead of becoming inf, and the
                                // low half then absorbs up to another 65504 of the remainder
                                h[e] = cvt_f16x2_sat(f[2 * e], f[2 * e + 1]);
                                const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[e]));
                                l[e] = cvt_f16x2_sat(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
                            } else {
                                const __nv_bfloat162 hh = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                                const float2 hf = __bfloat1622float2(hh);
                                const __nv_bfloat162 ll = __floats2bfloat162_rn(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
                                h[e] = *reinterpret_cast<const uint32_t*>(&hh);
                                l[e] = *reinterpret_cast<const uint32_t*>(&ll);
                            }
                        }
                        const uint32_t off = (uint32_t)r * 64u + (uint32_t)((co ^ ((r >> 1) & 3)) << 4);
                        *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
                    }
                } else {
                float4* hi = reinterpret_cast<float4*>(smem_gen + a_off(s));
                float4* lo = reinterpret_cast<float4*>(smem_gen + alo_off(s));
#pragma unroll
                for (int j = 0; j < (int)(TC_A_BYTES / 16 / 128); ++j) {
                    const int i = t + 128 * j;
                    const float4 v = hi[i];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                    hi[i] = h;
                    lo[i] = l;
                }
                }
                fence_proxy_async();
                if (CTA2) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(split_bar(s));          // local; the peer's forwarder relays it to the leader
                } else {
                    mbar_arrive(split_bar(s));
                }
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    }

    __syncwarp();
    tc_fence_before();
    if (CTA2) cluster_sync_all(); else __syncthreads();     // CTA2: neither CTA may exit while its peer still reads its smem / TMEM
    if (warp == 1) {
        tc_fence_after();
        if (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS));
    }
}

// hi/lo split used for the weights (same bit arithmetic as the in-kernel activation split)
__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float v = w[i];
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        hi[i] = h;
        lo[i] = v - h;
    }
}

int split_tf32(const float* w, float* hi, float* lo, long long n, cudaStream_t st) {
    CUM_REQUIRE(w && hi && lo && n > 0, "split_tf32: bad arguments");
    split_tf32_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(w, hi, lo, n);
    CUM_LAUNCH_CHECK("split_tf32_kernel");
    return CUM_OK;
}

// bf16 hi/lo split of the weights for BF16X3 (round-to-nearest, same arithmetic as the in-kernel activation split)
__global__ void split_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float v = w[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

int split_bf16(const float* w, void* hi, void* lo, long long n, cudaStream_t st) {
    CUM_REQUIRE(w && hi && lo && n > 0, "split_bf16: bad arguments");
    split_bf16_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(w, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
    CUM_LAUNCH_CHECK("split_bf16_kernel");
    return CUM_OK;
}

// fp16 hi/lo split of (scale * w) for F16X3; scale is a power of two chosen by the caller so that max|scale*w| ~ 8..16
__global__ void split_f16_kernel(const float* __restrict__ w, __half* __restrict__ hi, __half* __restrict__ lo, long long n, float scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float v = fminf(fmaxf(w[i] * scale, -65504.f), 65504.f);
        const __half h = __float2half_rn(v);
        hi[i] = h;
        lo[i] = __float2half_rn(v - __half2float(h));
    }
}

int split_f16(const float* w, void* hi, void* lo, long long n, float scale, cudaStream_t st) {
    CUM_REQUIRE(w && hi && lo && n > 0 && scale > 0.f, "split_f16: bad arguments");
    split_f16_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(w, (__half*)hi, (__half*)lo, n, scale);
    CUM_LAUNCH_CHECK("split_f16_kernel");
    return CUM_OK;
}

// ------------------------------------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

static int make_map(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_elems,
                    uint64_t s2_elems, uint32_t box0, uint32_t box1, const char* what, bool bf16 = false, bool sw128_16 = false) {
    return make_tensor_map(tm, base, d0, d1, d2, s1_elems, s2_elems, box0, box1, what, bf16, sw128_16);
}

static int encode_map(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_elems,
                      uint64_t s2_elems, uint32_t box0, uint32_t box1, const char* what, bool bf16, bool sw128_16);

// Tensor-map cache: a descriptor depends only on (base pointer, extents, strides, box, element type, swizzle), and a model
// re-issues the same few hundred of them every forward (PyTorch's caching allocator hands the same activation buffers back), so
// cuTensorMapEncodeTiled (a driver call, 3-4 per GEMM launch) runs once per distinct operand instead of once per launch.
// Process-wide, mutex-protected, bounded; cum_shutdown() empties it.  The encoded bytes do not depend on the device.
struct TmKey {
    uint64_t v[8];
    bool operator==(const TmKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmKeyHash {
    size_t operator()(const TmKey& k) const {
        uint64_t h = 0x9e3779b97f4a7c15ull;
        for (int i = 0; i < 8; ++i) { h ^= k.v[i] + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); }
        return (size_t)h;
    }
};
static std::mutex g_tm_mu;
static std::unordered_map<TmKey, CUtensorMap, TmKeyHash> g_tm_cache;
constexpr size_t TM_CACHE_MAX = 8192;
void tensor_map_cache_clear() {
    std::lock_guard<std::mutex> lk(g_tm_mu);
    g_tm_cache.clear();
}

int make_tensor_map(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_elems,
                    uint64_t s2_elems, uint32_t box0, uint32_t box1, const char* what, bool bf16, bool sw128_16) {
    TmKey key;
    key.v[0] = reinterpret_cast<uint64_t>(base); key.v[1] = d0; key.v[2] = d1; key.v[3] = d2; key.v[4] = s1_elems; key.v[5] = s2_elems;
    key.v[6] = ((uint64_t)box0 << 32) | box1; key.v[7] = (bf16 ? 1u : 0u) | (sw128_16 ? 2u : 0u);
    {
        std::lock_guard<std::mutex> lk(g_tm_mu);
        auto it = g_tm_cache.find(key);
        if (it != g_tm_cache.end()) { *tm = it->second; return CUM_OK; }
    }
    const int rc_enc = encode_map(tm, base, d0, d1, d2, s1_elems, s2_elems, box0, box1, what, bf16, sw128_16);
    if (rc_enc) return rc_enc;
    std::lock_guard<std::mutex> lk(g_tm_mu);
    if (g_tm_cache.size() >= TM_CACHE_MAX) g_tm_cache.clear();
    g_tm_cache.emplace(key, *tm);
    return CUM_OK;
}

static int encode_map(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_elems,
                      uint64_t s2_elems, uint32_t box0, uint32_t box1, const char* what, bool bf16, bool sw128_16) {
    auto enc = get_encode();
    if (!enc) { set_error("gemm_tc: cuTensorMapEncodeTiled unavailable"); return CUM_ECUDA; }
    cuuint64_t dims[3] = {d0, d1, d2};
    const uint64_t esz = bf16 ? 2 : 4;
    cuuint64_t strides[2] = {s1_elems * esz, s2_elems * esz};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base),
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     (bf16 && !sw128_16) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_tc: cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %llu,%llu,%llu strides %llu,%llu)", what,
                  (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                  (unsigned long long)strides[0], (unsigned long long)strides[1]);
        return CUM_ECUDA;
    }
    return CUM_OK;
}

// split-K (wgrad) launch context: set by wgrad_tc_fwd around its launches (host-side, per calling thread)
static thread_local int g_wgrad_kbs = 0, g_wgrad_koff = 0;

template <int MODE, int BN, int EPI, int OUTF = 0, bool CTA2 = false>
static int launch_tc(const cum_gemm_desc& d, cudaStream_t st) {
    using Cfg = TcCfg<MODE, BN, CTA2>;
    constexpr bool X3 = Cfg::PASS3;
    constexpr bool BF = Cfg::HALF;
    constexpr bool P16 = Cfg::PLAIN16;
    constexpr bool PRES = Cfg::PRES;
    auto kern = gemm_tc_kernel<MODE, BN, EPI, OUTF, CTA2>;
    { const int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(kern), (int)Cfg::SMEM_BYTES, "cudaFuncSetAttribute(gemm_tc_kernel)"); if (rc_attr) return rc_attr; }
    CUtensorMap tmA, tmAl, tmWh, tmWl;
    const bool planes = d.a_planes > 0 && !g_wgrad_kbs;        // plane-major A: dims (a_plane_k, a_rows, a_planes)
    const uint64_t a_bs = ((d.batch > 1 || planes) && !g_wgrad_kbs) ? (uint64_t)d.a_batch_stride : (uint64_t)d.a_rows * (uint64_t)d.a_row_stride;
    const uint64_t a_d0 = planes ? (uint64_t)d.a_plane_k : (uint64_t)d.k, a_d2 = planes ? (uint64_t)d.a_planes : (uint64_t)d.batch;
    int rc = make_map(&tmA, d.a, g_wgrad_kbs ? (uint64_t)d.a_row_stride : a_d0, (uint64_t)d.a_rows,
                      g_wgrad_kbs ? 1 : a_d2, (uint64_t)d.a_row_stride, a_bs, Cfg::BK, TC_BM, "A", P16 || PRES, P16);
    if (rc) return rc;
    if (PRES) {     // low-half plane of the activations: same geometry
        rc = make_map(&tmAl, d.a_lo, a_d0, (uint64_t)d.a_rows, a_d2, (uint64_t)d.a_row_stride, a_bs, Cfg::BK, TC_BM,
                      "A_lo", true, false);
        if (rc) return rc;
    } else {
        tmAl = tmA;
    }
    const int w_rows = (planes && d.n_half) ? 2 * d.n : d.n;        // n_half: the packed weight has 2 n rows per tap, one half per batch parity
    const uint64_t w_ts = (uint64_t)w_rows * (uint64_t)d.ldw;
    const uint64_t w_k_extent = g_wgrad_kbs ? (uint64_t)d.ldw : (uint64_t)d.k;     // wgrad: W columns span every split
    rc = make_map(&tmWh, d.w, w_k_extent, (uint64_t)w_rows, (uint64_t)d.taps, (uint64_t)d.ldw, w_ts, Cfg::BK, Cfg::W_ROWS, "W", BF || P16, P16);
    if (rc) return rc;
    if (X3) {
        rc = make_map(&tmWl, d.w_lo, w_k_extent, (uint64_t)w_rows, (uint64_t)d.taps, (uint64_t)d.ldw, w_ts, TC_BK, Cfg::W_ROWS, "W_lo", BF);
        if (rc) return rc;
    } else {
        tmWl = tmWh;
    }
    TcParams p;
    p.m = d.m; p.n = d.n; p.k = d.k; p.taps = d.taps; p.shift0 = d.tap_shift[0]; p.shift1 = d.tap_shift[1];
    p.batch = d.batch; p.epi = d.epilogue;
    p.m_tiles = (int)cdiv(d.m, CTA2 ? 2 * TC_BM : TC_BM); p.n_tiles = (int)cdiv(d.n, BN); p.k_blocks = (int)cdiv(d.k, Cfg::BK);
    p.bias = d.bias; p.c = d.c; p.c_bs = d.c_batch_stride; p.c_rs = d.c_row_stride;
    p.addend = d.addend; p.add_bs = d.add_batch_stride; p.add_rs = d.add_row_stride;
    p.acc_scale = (MODE == TC_F16X3 || MODE == TC_F16PS) ? d.acc_scale : 1.0f;
    p.c_lo = d.c_lo; p.addend_lo = d.addend_lo;
    p.skip_wlo = (X3 && d.w_lo_is_zero) ? 1 : 0;
    p.w_k_batch_stride = g_wgrad_kbs; p.w_k_off = g_wgrad_koff;
    p.mn_splits = 0;
    p.aux = (OUTF == 0) ? d.aux : nullptr; p.aux_bs = d.aux_batch_stride; p.aux_rs = d.aux_row_stride;
    p.add_mask = (OUTF == 0 && d.addend_is_mask) ? 1 : 0;
    p.a_plane_k = planes ? d.a_plane_k : 0; p.a_plane0 = d.a_plane0; p.a_plane_step = d.a_plane_step; p.n_half = (planes && d.n_half) ? 1 : 0;
    p.a_scale_ptr = (MODE == TC_F16X3) ? d.a_scale_dev : nullptr;
    p.acc_scale_ptr = (MODE == TC_F16X3 && d.a_scale_dev) ? d.a_scale_dev + 1 : nullptr;
    const long long total = (long long)p.batch * p.m_tiles * p.n_tiles;
    CUM_REQUIRE(total < (1ll << 31), "gemm_tc: too many tiles");
    if (CTA2) {
        // one cluster of two CTAs (an SM pair of one TPC) per 256-row tile
        const int pairs = sm_count() / 2;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)(2 * (total < pairs ? total : pairs)));
        cfg.blockDim = dim3(Cfg::THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_enabled() ? 2 : 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmAl, tmWh, tmWl, p);
        if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(gemm_tc_kernel, cluster 2)");
        return CUM_OK;
    }
    const int grid = (int)(total < sm_count() ? total : sm_count());
    cudaError_t e = launch_kernel(kern, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, tmA, tmAl, tmWh, tmWl, p);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(gemm_tc_kernel)");
    return CUM_OK;
}

// CTA-pair policy for the 256-wide tiles.  Measured on B200 (E8 full, batch 64 x 10 s, whole forward): pairs are faster in every
// mode -- f16x3 46.2 vs 48.9 ms, tf32x3 62.1 vs 71.2, bf16x3 43.0 vs 45.3, bf16 24.5 vs 25.5 -- once no cluster-scope fence is
// left on the per-K-block path (with .release.cluster arrives the pair kernel was 1.7x SLOWER; see mbar_arrive_cluster).
// CUM_GEMM_CTA2=0 disables pairs (A/B measurements); cum_gemm_desc.cta_pair overrides per call.
static int cta2_policy() {
    static int v = -2;
    if (v == -2) {
        const char* e = getenv("CUM_GEMM_CTA2");
        v = !e ? 0 : (e[0] == '0' ? -1 : 1);
    }
    return v;
}

static bool narrow_pairs() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CUM_GEMM_CTA2_N128");
        v = (e && e[0] == '0') ? 0 : 1;          // measured: 41.9-42.5 vs 42.5-42.7 ms per E8-full 64 x 10 s step
    }
    return v != 0;
}

template <int MODE, int BN, bool CTA2>
static int dispatch_epi2(const cum_gemm_desc& d, cudaStream_t st) {
    if (d.c_lo) {           // fp16 hi / lo output planes (and addend): the activation format TC_F16PS consumes
        if constexpr (MODE == TC_F16PS || MODE == TC_F16X3) {
            switch (d.epilogue) {
                case CUM_EPI_NONE:        return launch_tc<MODE, BN, CUM_EPI_NONE, 2, CTA2>(d, st);
                case CUM_EPI_RELU:        return launch_tc<MODE, BN, CUM_EPI_RELU, 2, CTA2>(d, st);
                case CUM_EPI_GLU_SIGMOID: return launch_tc<MODE, BN, CUM_EPI_GLU_SIGMOID, 2, CTA2>(d, st);
                default:
                    if (epi_is_glu(d.epilogue)) return launch_tc<MODE, BN, TC_EPI_GENERIC_GLU, 2, CTA2>(d, st);   // ReLU / SiLU / GELU gates
                    set_error("gemm_tc: hi/lo output supports NONE / RELU / GLU_* epilogues"); return CUM_ENOTSUP;
            }
        } else {
            set_error("gemm_tc: hi/lo output planes are available for the F16X3 mode");
            return CUM_ENOTSUP;
        }
    }
    if (d.out_bf16) {       // bf16 output / addend: only the epilogues the bf16 variant of the model uses
        if constexpr (MODE == TC_BF16 || MODE == TC_F16X3) {
            switch (d.epilogue) {
                case CUM_EPI_NONE:        return launch_tc<MODE, BN, CUM_EPI_NONE, 1, CTA2>(d, st);
                case CUM_EPI_RELU:        return launch_tc<MODE, BN, CUM_EPI_RELU, 1, CTA2>(d, st);
                case CUM_EPI_GLU_SIGMOID: return launch_tc<MODE, BN, CUM_EPI_GLU_SIGMOID, 1, CTA2>(d, st);
                default:
                    if (epi_is_glu(d.epilogue)) return launch_tc<MODE, BN, TC_EPI_GENERIC_GLU, 1, CTA2>(d, st);
                    set_error("gemm_tc: bf16 output supports NONE / RELU / GLU_* epilogues"); return CUM_ENOTSUP;
            }
        } else {
            set_error("gemm_tc: bf16 output is available for the BF16 and F16X3 modes");
            return CUM_ENOTSUP;
        }
    }
    switch (d.epilogue) {
        case CUM_EPI_NONE:        return launch_tc<MODE, BN, CUM_EPI_NONE, 0, CTA2>(d, st);
        case CUM_EPI_RELU:        return launch_tc<MODE, BN, CUM_EPI_RELU, 0, CTA2>(d, st);
        case CUM_EPI_GLU_SIGMOID: return launch_tc<MODE, BN, CUM_EPI_GLU_SIGMOID, 0, CTA2>(d, st);
        default:
            return epi_is_glu(d.epilogue) ? launch_tc<MODE, BN, TC_EPI_GENERIC_GLU, 0, CTA2>(d, st)
                                          : launch_tc<MODE, BN, TC_EPI_GENERIC_UNARY, 0, CTA2>(d, st);
    }
}

template <int MODE, int BN>
static int dispatch_epi(const cum_gemm_desc& d, cudaStream_t st) {
    {
        // a CTA pair per 256 x BN tile when there are enough rows to fill whole pairs (BN = 128: the 128-channel layers are bound by
        // the shared-memory datapath -- MMA operand reads + TMA writes -- and a pair stages only half of the weight tile per CTA;
        // CUM_GEMM_CTA2_N128=0 restores single CTAs for them)
        const int pol = d.cta_pair != 0 ? d.cta_pair : cta2_policy();
        const bool want = pol >= 0 && (BN == 256 || narrow_pairs());
        if (want && d.m > TC_BM && (sm_count() & 1) == 0) return dispatch_epi2<MODE, BN, true>(d, st);
    }
    return dispatch_epi2<MODE, BN, false>(d, st);
}

// ------------------------------------------------------------------------------------------------ tensor-core wgrad
// dW_s[n, k] = sum_r dZ[r, n] * A[r + shift_s, k] as a split-K GEMM over the ROW dimension: both operands are first
// transposed into (channels, rows) matrices whose columns are laid out with a per-clip pitch P (one zero column per clip
// where a tap would otherwise reach into the neighbouring clip); the activation side is TF32-split while it is transposed.
// Then C'[n, k] += dZ^T[n, rows] . A^T[k, rows + shift] runs on the forward kernel (TF32X3: gradients need fp32 range),
// `splits` CTAs per output tile, atomic accumulation into the zero-initialised gradient.
template <bool SPLIT>
__global__ void __launch_bounds__(256) transpose_pitch_kernel(const float* __restrict__ x, long long x_bs, long long x_rs,
                                                               int rows_src, int cols, int batch, int pitch, long long r_pad,
                                                               int shift, float* __restrict__ out_hi, float* __restrict__ out_lo) {
    __shared__ float tile[32][33];
    const long long j0 = (long long)blockIdx.x * 32;      // output column (= flattened row index with pitch)
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
    // (clip, row) of column j0: ONE 32-bit division per thread, the 4 columns this thread reads are found incrementally
    // (a 64-bit divide / modulo per element made this kernel integer-bound at ~1.5 TB/s)
    const unsigned j0u = (unsigned)j0, pu = (unsigned)pitch;
    int b = (int)(j0u / pu), t = (int)(j0u - (unsigned)b * pu) + ty;
    while (t >= pitch) { t -= pitch; ++b; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long j = j0 + ty + 8 * i;
        const int ts = t + shift;                          // the tap shift is baked in here: a TMA box start must stay
        const int c = c0 + tx;                             // 16-byte aligned, so it cannot be a +-1 coordinate
        float v = 0.f;
        if (j < r_pad && b < batch && ts >= 0 && ts < rows_src && c < cols) v = x[(long long)b * x_bs + (long long)ts * x_rs + c];
        tile[ty + 8 * i][tx] = v;
        t += 8;
        while (t >= pitch) { t -= pitch; ++b; }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i;
        const long long j = j0 + tx;
        if (c < cols && j < r_pad) {
            const float v = tile[tx][ty + 8 * i];
            if (SPLIT) {
                const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
                out_hi[(long long)c * r_pad + j] = h;
                out_lo[(long long)c * r_pad + j] = v - h;
            } else {
                out_hi[(long long)c * r_pad + j] = v;
            }
        }
    }
}

struct WgradPlan { long long R, r_pad, rps; int splits, pitch; size_t ws_bytes; };

static WgradPlan plan_wgrad(const cum_wgrad_desc& d) {
    WgradPlan w;
    w.pitch = d.taps == 1 ? d.m : (d.m > d.a_rows ? d.m : d.a_rows);
    w.R = (long long)d.batch * w.pitch;
    const long long tiles = cdiv(d.n, TC_BM) * cdiv(d.k, 256);
    long long s = (3LL * sm_count() + tiles - 1) / tiles;
    const long long max_s = cdiv(w.R, 1024);              // at least 1024 rows (32 K-blocks) per split
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    if (s > 65535) s = 65535;
    w.rps = cdiv(cdiv(w.R, s), 32) * 32;
    w.splits = (int)cdiv(w.R, w.rps);
    w.r_pad = w.rps * w.splits;
    w.ws_bytes = (size_t)w.r_pad * ((size_t)d.n + 2 * (size_t)d.k) * sizeof(float);
    return w;
}

struct WgradMnPlan { int clips; long long rows_z, rows_a; size_t z_elems, a_elems, ws_bytes; };
static WgradMnPlan plan_wgrad_mn(const cum_wgrad_desc& d);
long long wgrad_tc_workspace_bytes(const cum_wgrad_desc& d) {       // either path fits
    const size_t a = plan_wgrad(d).ws_bytes, b = plan_wgrad_mn(d).ws_bytes;
    return (long long)(a > b ? a : b);
}

template <int BN> static int launch_wgrad_gemm(const cum_gemm_desc& g, cudaStream_t st) {
    if constexpr (BN == 256) {      // CTA pairs (half a W^T tile per CTA) when the output has more than one 128-row tile
        if (cta2_policy() >= 0 && g.m > TC_BM && (sm_count() & 1) == 0) return launch_tc<TC_TF32X3, BN, TC_EPI_ATOMIC_ADD, 0, true>(g, st);
    }
    return launch_tc<TC_TF32X3, BN, TC_EPI_ATOMIC_ADD, 0>(g, st);
}

// ------------------------------------------------------------------------------------------------ MN-major weight gradients
// dW_s[n, k] = sum_{b, r} dZ[b, r, n] A[b, r + shift_s, k] on the tensor cores WITHOUT transposing either operand (the K-major path
// above spends ~40 % of its time in the two transposing pre-passes) and at the fp16 tensor rate:
//   1. amax(dZ) -> a power-of-two scale computed ON THE DEVICE that lifts the gradients into fp16's range (back-propagated values
//      are ~1e-6: unscaled they underflow; the reciprocal reaches the GEMM epilogue through TcParams::acc_scale_ptr)
//   2. elementwise split of scale * dZ and of A into fp16 hi / lo planes, row-major as they are
//   3. per tap one split-K GEMM (splits = (clip, row range)) whose TMA boxes are MN-major operand slabs (gemm_tc_kernel<..., MNM>),
//      three fp16 MMA passes, atomic accumulation into the zero-initialised gradient.
// row-major (batch, rows, cols) with strides; `contig`: the whole tensor is one flat run (no index arithmetic per element)
__device__ __forceinline__ const float4* elem4(const float* x, long long bs, long long rs, int rows, int cols4, bool contig, long long i) {
    if (contig) return reinterpret_cast<const float4*>(x) + i;
    const unsigned c = (unsigned)(i % (unsigned)cols4);
    const long long br = i / (unsigned)cols4;
    const unsigned r = (unsigned)(br % (unsigned)rows);
    const long long b = br / (unsigned)rows;
    return reinterpret_cast<const float4*>(x + b * bs + (long long)r * rs) + c;
}

__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, long long bs, long long rs, int batch, int rows, int cols4,
                                                    unsigned* __restrict__ amax_bits) {
    const long long total = (long long)batch * rows * cols4;
    const bool contig = rs == 4ll * cols4 && (batch == 1 || bs == (long long)rows * rs);
    float m0 = 0.f, m1 = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < total; i += 2 * stride) {          // two independent loads in flight per thread
        const float4 v = __ldg(elem4(x, bs, rs, rows, cols4, contig, i)), u = __ldg(elem4(x, bs, rs, rows, cols4, contig, i + stride));
        m0 = fmaxf(m0, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        m1 = fmaxf(m1, fmaxf(fmaxf(fabsf(u.x), fabsf(u.y)), fmaxf(fabsf(u.z), fabsf(u.w))));
    }
    if (i < total) {
        const float4 v = __ldg(elem4(x, bs, rs, rows, cols4, contig, i));
        m0 = fmaxf(m0, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    float m = fmaxf(m0, m1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));      // non-negative floats order like their bits
}

// scale4 = {s, 1 / s, amax bits, -}: s = 2^(15 - e) with amax = f 2^e, f in [0.5, 1); 1 for an all-zero tensor
__global__ void grad_scale_finalize_kernel(float* __restrict__ scale4) {
    const float am = __uint_as_float(reinterpret_cast<const unsigned*>(scale4)[2]);
    float s = 1.f;
    if (am > 0.f) {
        int e;
        frexpf(am, &e);
        s = ldexpf(1.f, 15 - e);
    }
    scale4[0] = s;
    scale4[1] = 1.f / s;
}

int grad_scale_finalize(float* scale4, cudaStream_t st) {
    grad_scale_finalize_kernel<<<1, 1, 0, st>>>(scale4);
    CUM_LAUNCH_CHECK("grad_scale_finalize_kernel");
    return CUM_OK;
}

int grad_scale_fwd(const float* x, long long bs, long long rs, int batch, int rows, int cols, float* scale4, cudaStream_t st) {
    CUM_REQUIRE(x && scale4 && batch > 0 && rows > 0 && cols > 0 && cols % 4 == 0 && aligned16(x) && bs % 4 == 0 && rs % 4 == 0,
                "grad_scale: bad arguments (cols and strides must be multiples of 4, x 16-byte aligned)");
    cudaError_t e = cudaMemsetAsync(scale4, 0, 16, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_scale)");
    const long long total = (long long)batch * rows * (cols / 4);
    const long long want = cdiv(total, 256);
    const int grid = (int)(want < 8LL * sm_count() ? want : 8LL * sm_count());
    amax_kernel<<<grid, 256, 0, st>>>(x, bs, rs, batch, rows, cols / 4, reinterpret_cast<unsigned*>(scale4) + 2);
    CUM_LAUNCH_CHECK("amax_kernel");
    grad_scale_finalize_kernel<<<1, 1, 0, st>>>(scale4);
    CUM_LAUNCH_CHECK("grad_scale_finalize_kernel");
    return CUM_OK;
}

// hi = fp16(s x) (saturating), lo = fp16(s x - hi), s read from the device (1 when `scale` is NULL); compact (batch, rows, cols) planes
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, long long bs, long long rs, int batch, int rows, int cols4,
                                                            __half* __restrict__ hi, __half* __restrict__ lo,
                                                            const float* __restrict__ scale) {
    const float s = scale ? __ldg(scale) : 1.f;
    const long long total = (long long)batch * rows * cols4;
    const bool contig = rs == 4ll * cols4 && (batch == 1 || bs == (long long)rows * rs);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldg(elem4(x, bs, rs, rows, cols4, contig, i));
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        const uint32_t h0 = cvt_f16x2_sat(v.x, v.y), h1 = cvt_f16x2_sat(v.z, v.w);
        const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h0)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
        const uint32_t l0 = cvt_f16x2_sat(v.x - f0.x, v.y - f0.y), l1 = cvt_f16x2_sat(v.z - f1.x, v.w - f1.y);
        reinterpret_cast<uint2*>(hi)[i] = make_uint2(h0, h1);
        reinterpret_cast<uint2*>(lo)[i] = make_uint2(l0, l1);
    }
}

static int wgrad_mn_policy() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CUM_WGRAD_MN");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v;
}

static WgradMnPlan plan_wgrad_mn(const cum_wgrad_desc& d) {
    WgradMnPlan w;
    const bool flat = d.taps == 1;                 // 1x1 layers: all clips are one run of rows (no tap can cross a clip boundary)
    w.clips = flat ? 1 : d.batch;
    w.rows_z = flat ? (long long)d.batch * d.m : d.m;
    w.rows_a = flat ? (long long)d.batch * d.m : d.a_rows;
    w.z_elems = (size_t)w.clips * w.rows_z * d.n;
    w.a_elems = (size_t)w.clips * w.rows_a * d.k;
    w.ws_bytes = 256 + 4 * (((w.z_elems + 127) / 128) * 128 + ((w.a_elems + 127) / 128) * 128);      // two fp16 planes each
    return w;
}

// ZSPLIT: the gradient operand is read as fp32 straight from dz and split by the kernel's splitter warps (scaled by *scale);
// otherwise z_hi / z_lo are pre-split fp16 planes
// ZSPLIT == 2: the activation operand as well (fp32 straight from `a`, no planes, no workspace beyond the scale)
template <int BN, bool CTA2, int ZSPLIT>
static int launch_wgrad_mn(const __half* z_hi, const __half* z_lo, const __half* a_hi, const __half* a_lo, const WgradMnPlan& w,
                           const cum_wgrad_desc& d, int tap, const float* scale, cudaStream_t st) {
    constexpr int MODE = ZSPLIT ? TC_F16X3 : TC_F16PS;
    using Cfg = TcCfg<MODE, BN, CTA2>;
    auto kern = gemm_tc_kernel<MODE, BN, TC_EPI_ATOMIC_ADD, 0, CTA2, (ZSPLIT == 2 ? 2 : 1)>;
    const float* inv_scale = scale + 1;
    { const int rc_attr = ensure_dyn_smem(reinterpret_cast<const void*>(kern), (int)Cfg::SMEM_BYTES, "cudaFuncSetAttribute(gemm_tc_kernel MN)"); if (rc_attr) return rc_attr; }
    CUtensorMap tmA, tmAl, tmWh, tmWl;
    int rc;
    if (ZSPLIT) {       // fp32 dz as it lies in memory: (n, rows, clips) with the caller's strides, boxes of 32 channels x 32 rows
        const uint64_t zbs = w.clips > 1 ? (uint64_t)d.dz_batch_stride : (uint64_t)w.rows_z * (uint64_t)d.dz_row_stride;
        rc = make_map(&tmA, d.dz, (uint64_t)d.n, (uint64_t)w.rows_z, (uint64_t)w.clips, (uint64_t)d.dz_row_stride, zbs, 32, 32, "dZ");
        if (rc) return rc;
        tmAl = tmA;
    } else {
        rc = make_map(&tmA, z_hi, (uint64_t)d.n, (uint64_t)w.rows_z, (uint64_t)w.clips, (uint64_t)d.n, (uint64_t)w.rows_z * d.n, 64, 32, "dZ_hi", true, true);
        if (rc) return rc;
        rc = make_map(&tmAl, z_lo, (uint64_t)d.n, (uint64_t)w.rows_z, (uint64_t)w.clips, (uint64_t)d.n, (uint64_t)w.rows_z * d.n, 64, 32, "dZ_lo", true, true);
        if (rc) return rc;
    }
    if (ZSPLIT == 2) {
        const uint64_t abs_ = w.clips > 1 ? (uint64_t)d.a_batch_stride : (uint64_t)w.rows_a * (uint64_t)d.a_row_stride;
        rc = make_map(&tmWh, d.a, (uint64_t)d.k, (uint64_t)w.rows_a, (uint64_t)w.clips, (uint64_t)d.a_row_stride, abs_, 32, 32, "A");
        if (rc) return rc;
        tmWl = tmWh;
    } else {
        rc = make_map(&tmWh, a_hi, (uint64_t)d.k, (uint64_t)w.rows_a, (uint64_t)w.clips, (uint64_t)d.k, (uint64_t)w.rows_a * d.k, 64, 32, "A_hi", true, true);
        if (rc) return rc;
        rc = make_map(&tmWl, a_lo, (uint64_t)d.k, (uint64_t)w.rows_a, (uint64_t)w.clips, (uint64_t)d.k, (uint64_t)w.rows_a * d.k, 64, 32, "A_lo", true, true);
        if (rc) return rc;
    }
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.m = d.n; p.n = d.k; p.k = 32; p.taps = 1; p.shift0 = d.tap_shift[tap]; p.shift1 = 0; p.epi = CUM_EPI_NONE;
    p.m_tiles = (int)cdiv(d.n, CTA2 ? 2 * TC_BM : TC_BM); p.n_tiles = (int)cdiv(d.k, BN);
    // split-K: ~3 waves of CTAs, at least 8 K-blocks (256 rows) per split
    const long long kb_clip = cdiv(w.rows_z, 32);
    const long long tiles = (long long)p.m_tiles * p.n_tiles;
    const long long units = CTA2 ? sm_count() / 2 : sm_count();
    long long want = cdiv(cdiv(3 * units, tiles), w.clips);
    const long long max_s = kb_clip / 8 > 0 ? kb_clip / 8 : 1;
    if (want > max_s) want = max_s;
    if (want < 1) want = 1;
    p.k_blocks = (int)cdiv(kb_clip, want);
    p.mn_splits = (int)cdiv(kb_clip, p.k_blocks);
    const long long batch = (long long)w.clips * p.mn_splits;
    CUM_REQUIRE(batch * tiles < (1ll << 31), "wgrad: too many tiles");
    p.batch = (int)batch;
    p.bias = nullptr; p.c = d.dw + (size_t)tap * d.n * d.ldw; p.c_bs = 0; p.c_rs = d.ldw;
    p.addend = nullptr; p.acc_scale = 1.0f; p.acc_scale_ptr = inv_scale; p.a_scale_ptr = ZSPLIT ? scale : nullptr; p.skip_wlo = 0;
    const long long total = batch * tiles;
    if (CTA2) {
        const int pairs = sm_count() / 2;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)(2 * (total < pairs ? total : pairs)));
        cfg.blockDim = dim3(Cfg::THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmAl, tmWh, tmWl, p);
        if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(gemm_tc_kernel MN, cluster 2)");
        return CUM_OK;
    }
    const int grid = (int)(total < sm_count() ? total : sm_count());
    kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmAl, tmWh, tmWl, p);
    CUM_LAUNCH_CHECK("gemm_tc_kernel (MN-major wgrad)");
    return CUM_OK;
}

static bool wgrad_mn_ok(const cum_wgrad_desc& d) {
    const auto s4 = [](long long v) { return v % 4 == 0; };
    return wgrad_mn_policy() && d.n % 8 == 0 && d.k % 8 == 0 && aligned16(d.dz) && aligned16(d.a) && s4(d.dz_row_stride) && s4(d.a_row_stride) &&
           s4(d.dz_batch_stride) && s4(d.a_batch_stride) &&
           (d.taps != 1 || d.batch == 1 || (d.dz_batch_stride == (long long)d.m * d.dz_row_stride && d.a_batch_stride == (long long)d.m * d.a_row_stride));
}

static int wgrad_mn_fwd(const cum_wgrad_desc& d, cudaStream_t st) {
    const WgradMnPlan w = plan_wgrad_mn(d);
    uint8_t* ws = reinterpret_cast<uint8_t*>(d.workspace);
    float* own_scale = reinterpret_cast<float*>(ws);                 // {s, 1/s, amax bits, -} when the caller brings no scale
    const size_t zpl = ((w.z_elems + 127) / 128) * 128, apl = ((w.a_elems + 127) / 128) * 128;
    __half* z_hi = reinterpret_cast<__half*>(ws + 256);
    __half* z_lo = z_hi + zpl;
    __half* a_hi = z_lo + zpl;
    __half* a_lo = a_hi + apl;
    const bool flat = d.taps == 1;
    const int zb = flat ? 1 : d.batch;
    const long long z_bs = flat ? 0 : d.dz_batch_stride, a_bs = flat ? 0 : d.a_batch_stride;
    const float* scale = d.dz_scale_dev;
    if (!scale) {
        const int rc = grad_scale_fwd(d.dz, z_bs, d.dz_row_stride, zb, (int)w.rows_z, d.n, own_scale, st);
        if (rc) return rc;
        scale = own_scale;
    }
    const int grid = 8 * sm_count();
    // the gradient is split by the GEMM's own splitter warps (read once, as fp32); CUM_WGRAD_ZSPLIT=0: a pre-pass writes fp16 planes
    // CUM_WGRAD_ZSPLIT: 1 (default) the gradient is split in the kernel, the activation planes by an elementwise pre-pass; 2 both
    // operands in the kernel (no pre-pass at all, but the splitter warps then sit on the critical path of every stage: measured
    // 13.9-14.3 vs 13.4-13.7 ms); 0 both by pre-passes
    static const int zsplit = getenv("CUM_WGRAD_ZSPLIT") ? atoi(getenv("CUM_WGRAD_ZSPLIT")) : 1;
    if (!zsplit) {
        split_planes_kernel<<<grid, 256, 0, st>>>(d.dz, z_bs, d.dz_row_stride, zb, (int)w.rows_z, d.n / 4, z_hi, z_lo, scale);
        CUM_LAUNCH_CHECK("split_planes_kernel(dz)");
    }
    if (zsplit != 2) {
        split_planes_kernel<<<grid, 256, 0, st>>>(d.a, a_bs, d.a_row_stride, zb, (int)w.rows_a, d.k / 4, a_hi, a_lo, nullptr);
        CUM_LAUNCH_CHECK("split_planes_kernel(a)");
    }
    const bool pair = cta2_policy() >= 0 && d.n > TC_BM && (sm_count() & 1) == 0;
    for (int s = 0; s < d.taps; ++s) {
        int rc;
#define WG_LAUNCH(Z)                                                                                                        \
        (d.k <= 128 ? (pair ? launch_wgrad_mn<128, true, Z>(z_hi, z_lo, a_hi, a_lo, w, d, s, scale, st)                           \
                            : launch_wgrad_mn<128, false, Z>(z_hi, z_lo, a_hi, a_lo, w, d, s, scale, st))                         \
                    : (pair ? launch_wgrad_mn<256, true, Z>(z_hi, z_lo, a_hi, a_lo, w, d, s, scale, st)                           \
                            : launch_wgrad_mn<256, false, Z>(z_hi, z_lo, a_hi, a_lo, w, d, s, scale, st)))
        rc = zsplit == 2 ? WG_LAUNCH(2) : zsplit == 1 ? WG_LAUNCH(1) : WG_LAUNCH(0);
#undef WG_LAUNCH
        if (rc) return rc;
    }
    return CUM_OK;
}

int wgrad_tc_fwd(const cum_wgrad_desc& d, cudaStream_t st) {
    CUM_REQUIRE(d.workspace, "wgrad: tensor-core mode needs a workspace (cum_gemm_wgrad_workspace_bytes)");
    CUM_REQUIRE(aligned16(d.workspace), "wgrad: workspace must be 16-byte aligned");
    CUM_REQUIRE(d.n % 8 == 0 && d.k % 8 == 0, "wgrad: n and k must be multiples of 8");
    if (wgrad_mn_ok(d)) return wgrad_mn_fwd(d, st);
    const WgradPlan w = plan_wgrad(d);
    float* zt = reinterpret_cast<float*>(d.workspace);            // dZ^T   (n, r_pad)
    float* at_hi = zt + (size_t)d.n * w.r_pad;                     // A^T hi (k, r_pad)
    float* at_lo = at_hi + (size_t)d.k * w.r_pad;
    const int cb = d.taps == 1 ? 1 : d.batch;                      // taps == 1: rows of all clips are one flat run
    const long long rows_z = d.taps == 1 ? (long long)d.batch * d.m : d.m;
    const long long rows_a = d.taps == 1 ? (long long)d.batch * d.m : d.a_rows;
    const int pitch = d.taps == 1 ? (int)(rows_z > 2147483647LL ? 0 : rows_z) : w.pitch;
    CUM_REQUIRE(pitch > 0, "wgrad: too many rows");
    if (d.taps == 1)
        CUM_REQUIRE(d.batch == 1 || (d.dz_batch_stride == (long long)d.m * d.dz_row_stride && d.a_batch_stride == (long long)d.m * d.a_row_stride),
                    "wgrad: taps=1 with batch>1 needs batch-contiguous operands");
    dim3 gz((unsigned)cdiv(w.r_pad, 32), (unsigned)cdiv(d.n, 32)), ga((unsigned)cdiv(w.r_pad, 32), (unsigned)cdiv(d.k, 32));
    transpose_pitch_kernel<false><<<gz, 256, 0, st>>>(d.dz, d.dz_batch_stride, d.dz_row_stride, (int)rows_z, d.n, cb, pitch, w.r_pad, 0, zt, nullptr);
    CUM_LAUNCH_CHECK("transpose_pitch_kernel(dz)");
    for (int s = 0; s < d.taps; ++s) {
        // A^T for this tap: column (b, t) <- A[b, t + shift_s] (zero outside the clip), TF32 hi / lo halves
        transpose_pitch_kernel<true><<<ga, 256, 0, st>>>(d.a, d.a_batch_stride, d.a_row_stride, (int)rows_a, d.k, cb, pitch, w.r_pad,
                                                         d.tap_shift[s], at_hi, at_lo);
        CUM_LAUNCH_CHECK("transpose_pitch_kernel(a)");
        cum_gemm_desc g;
        memset(&g, 0, sizeof(g));
        g.a = zt; g.a_batch_stride = w.rps; g.a_row_stride = w.r_pad; g.a_rows = d.n; g.k = (int)w.rps; g.taps = 1;
        g.w = at_hi; g.w_lo = at_lo; g.ldw = (int)w.r_pad; g.bias = nullptr;
        g.c = d.dw + (size_t)s * d.n * d.ldw; g.c_batch_stride = 0; g.c_row_stride = d.ldw;
        g.m = d.n; g.n = d.k; g.batch = w.splits; g.epilogue = CUM_EPI_NONE; g.math = CUM_MATH_TF32X3;
        CUM_REQUIRE(w.r_pad < 2147483647LL, "wgrad: padded row count exceeds the TMA coordinate range");
        g_wgrad_kbs = (int)w.rps;
        g_wgrad_koff = 0;
        const int rc = d.k <= 128 ? launch_wgrad_gemm<128>(g, st) : launch_wgrad_gemm<256>(g, st);
        g_wgrad_kbs = 0; g_wgrad_koff = 0;
        if (rc) return rc;
    }
    return CUM_OK;
}

int gemm_tc_fwd(const cum_gemm_desc& d, cudaStream_t st) {
    CUM_REQUIRE(d.a_row_stride >= (d.a_planes > 0 ? d.a_plane_k : d.k) || d.a_rows == 1, "gemm_tc: a_row_stride < k");
    CUM_REQUIRE(!d.a_scale_dev || (d.math == CUM_MATH_F16X3 && !d.a_lo), "gemm_tc: a_scale_dev needs CUM_MATH_F16X3 with fp32 activations");
    CUM_REQUIRE(!d.addend_is_mask || (d.addend && !d.out_bf16 && !d.c_lo && !d.addend_lo && (d.epilogue == CUM_EPI_NONE || d.epilogue == CUM_EPI_RELU)),
                "gemm_tc: addend_is_mask needs an fp32 addend / output and a NONE / RELU epilogue");
    CUM_REQUIRE(!d.aux || (!d.out_bf16 && !d.c_lo && aligned16(d.aux) && d.aux_row_stride % 2 == 0 && d.aux_batch_stride % 2 == 0 &&
                           (d.epilogue == CUM_EPI_NONE || d.epilogue == CUM_EPI_RELU || epi_is_glu(d.epilogue))),
                "gemm_tc: aux (second output) needs an fp32 output, even strides and a NONE / RELU / GLU epilogue");
    CUM_REQUIRE(d.out_bf16 || d.c_lo || (aligned16(d.c) && d.c_row_stride % 4 == 0 && (d.batch == 1 || d.c_batch_stride % 4 == 0)),
                "gemm_tc: an fp32 output must be 16-byte aligned with row / batch strides that are multiples of 4 elements (c_row_stride=%lld)",
                (long long)d.c_row_stride);
    CUM_REQUIRE(!d.c_lo || (aligned16(d.c) && aligned16(d.c_lo) && d.c_row_stride % 4 == 0 && (d.batch == 1 || d.c_batch_stride % 4 == 0)),
                "gemm_tc: hi/lo output planes must be 16-byte aligned with strides that are multiples of 4 elements");
    CUM_REQUIRE((long long)d.m * d.c_row_stride < (1ll << 31) && (!d.addend || (long long)d.m * d.add_row_stride < (1ll << 29)),
                "gemm_tc: one batch plane of the output / addend must span fewer than 2^31 elements (m=%d, c_row_stride=%lld)",
                d.m, (long long)d.c_row_stride);
    if (d.a_planes > 0) {
        const int bk = d.math == CUM_MATH_BF16 ? 64 : TC_BK;
        CUM_REQUIRE(d.a_plane_k > 0 && d.a_plane_k % bk == 0, "gemm_tc: plane-major a needs a_plane_k (= %d) to be a multiple of the K-block (%d elements)", d.a_plane_k, bk);
        CUM_REQUIRE(!d.n_half || (d.n % 16 == 0 && !epi_is_glu(d.epilogue)), "gemm_tc: n_half needs n %% 16 == 0 and a non-GLU epilogue");
    }
    // Tile width.  128-wide tiles for narrow layers, and for problems whose 256-wide tiling would fill less than HALF of the SMs
    // (streaming / small batches: e.g. out_proj of 4096 streams x 1 hop is 64 tiles of K = 2048 on 148 SMs): twice the tiles in the
    // same single wave, half the time per tile.  Same products and accumulation order per output element: bit-identical results.
    // Measured (E6 streaming from the graph): single stream 0.87 -> 0.75 ms per hop, 512 streams 1.18 -> 1.10 ms, 4096 streams
    // unchanged.  Going further (128-wide whenever the wave-quantised time waves x width is smaller, e.g. 384 tiles in 3 waves
    // instead of 192 in 2) measured SLOWER: 1.16 vs 1.01 ms for the GEMMs of a 4096-stream call -- a 128-wide tile re-reads its A
    // rows for twice as many column tiles and runs at ~2/3 of the efficiency.  CUM_GEMM_FILL=0 disables
    static int split_env = -1;
    if (split_env < 0) { const char* e = getenv("CUM_GEMM_FILL"); split_env = (e && e[0] == '0') ? 0 : 1; }
    const long long tiles256 = cdiv(d.m, TC_BM) * (long long)d.batch * cdiv(d.n, 256);
    const bool narrow = d.n <= 128 || (split_env && 2 * tiles256 <= sm_count());
    if (d.math == CUM_MATH_TF32X3) {
        CUM_REQUIRE(d.w_lo && aligned16(d.w_lo), "gemm_tc: TF32X3 needs w_lo (see cum_split_tf32)");
        return narrow ? dispatch_epi<TC_TF32X3, 128>(d, st) : dispatch_epi<TC_TF32X3, 256>(d, st);
    }
    if (d.math == CUM_MATH_BF16X3) {
        CUM_REQUIRE(d.w_lo && aligned16(d.w_lo), "gemm_tc: BF16X3 needs w_lo (see cum_split_bf16)");
        CUM_REQUIRE(d.ldw % 8 == 0, "gemm_tc: BF16X3 needs ldw %% 8 == 0 (ldw=%d)", d.ldw);
        return narrow ? dispatch_epi<TC_BF16X3, 128>(d, st) : dispatch_epi<TC_BF16X3, 256>(d, st);
    }
    if (d.math == CUM_MATH_BF16) {
        CUM_REQUIRE(d.k % 8 == 0 && d.ldw % 8 == 0 && d.a_row_stride % 8 == 0 && d.a_batch_stride % 8 == 0,
                    "gemm_tc: BF16 needs k, ldw and the a strides to be multiples of 8 elements");
        return narrow ? dispatch_epi<TC_BF16, 128>(d, st) : dispatch_epi<TC_BF16, 256>(d, st);
    }
    CUM_REQUIRE(!d.out_bf16 || d.math == CUM_MATH_F16X3, "gemm_tc: bf16 output is available for the BF16 and F16X3 modes");
    CUM_REQUIRE((!d.a_lo && !d.c_lo && !d.addend_lo) || d.math == CUM_MATH_F16X3, "gemm_tc: hi/lo activation planes need CUM_MATH_F16X3");
    CUM_REQUIRE(!(d.addend_lo && d.out_bf16), "gemm_tc: a bf16 output takes a bf16 addend");
    CUM_REQUIRE(!(d.c_lo && d.out_bf16), "gemm_tc: c_lo and out_bf16 are mutually exclusive");
    if (d.math == CUM_MATH_F16X3 && d.a_lo) {
        CUM_REQUIRE(d.w_lo && aligned16(d.w_lo) && aligned16(d.a_lo), "gemm_tc: F16X3 needs w_lo; a_lo must be 16-byte aligned");
        CUM_REQUIRE(d.ldw % 8 == 0 && d.k % 8 == 0 && d.a_row_stride % 8 == 0 && d.a_batch_stride % 8 == 0,
                    "gemm_tc: hi/lo activation planes need k, ldw and the a strides to be multiples of 8 elements");
        CUM_REQUIRE(d.acc_scale > 0.f, "gemm_tc: F16X3 needs acc_scale = 1 / (weight scale passed to cum_split_f16)");
        return narrow ? dispatch_epi<TC_F16PS, 128>(d, st) : dispatch_epi<TC_F16PS, 256>(d, st);
    }
    if (d.math == CUM_MATH_F16X3) {
        CUM_REQUIRE(d.w_lo && aligned16(d.w_lo), "gemm_tc: F16X3 needs w_lo (see cum_split_f16)");
        CUM_REQUIRE(d.ldw % 8 == 0, "gemm_tc: F16X3 needs ldw %% 8 == 0 (ldw=%d)", d.ldw);
        CUM_REQUIRE(d.acc_scale > 0.f, "gemm_tc: F16X3 needs acc_scale = 1 / (weight scale passed to cum_split_f16)");
        return narrow ? dispatch_epi<TC_F16X3, 128>(d, st) : dispatch_epi<TC_F16X3, 256>(d, st);
    }
    return narrow ? dispatch_epi<TC_TF32, 128>(d, st) : dispatch_epi<TC_TF32, 256>(d, st);
}

}  // namespace cum
