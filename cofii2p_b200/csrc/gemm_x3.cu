// gemm_x3.cu -- persistent 3xTF32 contraction engine (COFI_GEMM_TF32X3S): the fp32-grade tensor-core path of the parity
// engine for nn.Linear layers and NHWC convolutions.
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T ),  A fp32 in HBM (K-major), W PRE-SPLIT by cofi_split_tf32 into two tf32-exact
//   planes Wh = rn_tf32(W), Wl = rn_tf32(W - Wh);  the kernel forms  Ah*Wh + Al*Wh + Ah*Wl  with Ah/Al split on the fly.
//
// Why a second kernel next to gemm_tc.cu: the one-tile-per-CTA 3xTF32 kernel there splits both operands from shared
// memory back into shared memory and then lets tcgen05.mma read both operands from shared memory.  ncu (profiles/
// r2_ncu_gemm_tc_x3.md) shows 1387 LSU shared-memory wavefronts per k-block for the split on top of the 768 clk of
// operand reads the twelve MMAs need: the 128 B/clk shared-memory port is the bound (2090 clk per k-block measured, 40 %
// tensor-pipe).  Here
//   * W arrives already split (weights are constants: split once per weights epoch), so it never passes through the LSU;
//   * A is split in REGISTERS by eight splitter warps (thread = tile row) and written to TENSOR MEMORY with tcgen05.st:
//     the MMAs take A from TMEM (tcgen05.mma [d], [a_tmem], b_desc) and only W from shared memory -- 16 KB LSU reads +
//     48 KB operand reads per k-block instead of 176 KB;
//   * the kernel is persistent (one CTA per SM strides over the output tiles) with two TMEM accumulators, so the epilogue
//     of tile i overlaps the main loop of tile i+1 and the TMA rings (4 A stages, 3-6 W stages) run across tile
//     boundaries: the K <= 128 streaming contractions keep loads in flight all the time instead of one load -> split ->
//     MMA -> epilogue chain per CTA.
//
// Warp roles (18 warps): 0 TMA producer, 1 MMA issuer, 2-9 epilogue (two sets of four, one warp per TMEM lane quarter
// and set, alternating over the 32-column chunks), 10-17 splitters (two sets of four alternating over the k-blocks).
// TMEM (512 columns): [0, 2*BN) two accumulators, [256, 512) four A slots of 64 columns (hi | lo of a 128 x 32 tile).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace cofi {

struct Epilogue {  // must match gemm_simt.cu / gemm_tc.cu
    const float* bias;
    const float* rowdiv;
    const float* colscale;
    const float* colshift;
    const float* residual;
    int64_t ldres;
    int accumulate;
    int act;
    const float* ln_gamma;
    const float* ln_beta;
    float ln_eps;
};

namespace tc {

const CUtensorMap* get_tmap_f32_es(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                   const uint32_t* box, const uint32_t* elem_strides);  // gemm_tc.cu

namespace x3 {

constexpr int TM = 128, TK = 32, A_BYTES = TM * TK * 4;
constexpr int CONV_TW = 64, CONV_TH = 2;
constexpr int NT = 4;         // TMEM A slots
constexpr int A_COL0 = 256;   // first TMEM column of the A slots
constexpr int EPI_WARPS = 8, SPLIT_WARPS = 8;
constexpr int THREADS = 32 * (2 + EPI_WARPS + SPLIT_WARPS);
constexpr int STAGE_PER_WARP = 8192;  // two 4096 B swizzled TMA-store boxes (alternating), or one [32][36] fp32 transposing area

template <int BN>
struct Cfg {
    static constexpr int B_PLANE = BN * TK * 4;
    static constexpr int B_STAGE = 2 * B_PLANE;  // hi tile + lo tile
    static constexpr int NSB = BN == 128 ? 3 : (BN == 64 ? 4 : 6);
    // shared-memory A stages (raw fp32 tiles; a stage is free again as soon as the splitters have read it, the TMEM slots
    // buffer the rest).  MUST be even: the two splitter sets alternate over the k-blocks, and a set that saw only every
    // other phase of a barrier could not tell a completed phase from the one two before it (mbarrier parity).
    static constexpr int NSA = BN == 128 ? 2 : 4;
    static_assert(NSA % 2 == 0 && NT % 2 == 0, "splitter sets own the even / odd slots");
    static constexpr int OFF_B = NSA * A_BYTES;
    static constexpr int OFF_STG = OFF_B + NSB * B_STAGE;
    static constexpr int OFF_BAR = OFF_STG + EPI_WARPS * STAGE_PER_WARP;
    static constexpr int OFF_VEC = OFF_BAR + 512;
    static constexpr int SMEM = OFF_VEC + 20 * BN * 4 + 1024 /* alignment slack */;   // 4 epilogue vectors + 2 x [4][BN][2] statistics
};

struct Params {
    float* C;
    int64_t ldc;
    int64_t M;
    int N;
    int num_kb;
    int m_tiles, n_tiles;
    Epilogue ep;
    int Ho, Wo, Cin, cpt, KW, pad, tiles_w, tiles_per_img, cstride;  // convolution geometry (CONV only)
    int tma_store;
    float* stat_out;
    unsigned long long* prof;  // COFI_X3_PROFILE: per-role barrier wait clocks (see cofi_debug_x3_profile)
    int dbg;                   // COFI_X3_DEBUG bits (perf triage only, results become wrong): 1 no column statistics, 2 no stores,
                               // 4 no proxy fence, 8 no epilogue math, 16 no per-tile epilogue barriers
};

// wait with optional accounting (perf triage only): slot = which wait of which role
#define X3_WAIT(bar, parity, slot)                        \
    do {                                                  \
        if (p.prof) {                                     \
            const long long t0_ = clock64();              \
            mbar_wait(bar, parity);                       \
            pacc[slot] += (unsigned long long)(clock64() - t0_); \
        } else {                                          \
            mbar_wait(bar, parity);                       \
        }                                                 \
    } while (0)

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// warp-collective: lane i writes 16 consecutive 32-bit columns of TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
            taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// one thread: D[tmem] (+)= A[tmem] * B[smem], tf32
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc512(uint32_t* dst_in_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(dst_in_smem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

struct Tile {
    int64_t m0;
    int n0, cb, ch0, cw0;
};
template <bool CONV>
__device__ __forceinline__ Tile decode(const Params& p, int t) {
    Tile tl;
    const int mt = t / p.n_tiles;
    tl.n0 = (t - mt * p.n_tiles);
    tl.m0 = 0;
    tl.cb = tl.ch0 = tl.cw0 = 0;
    if (CONV) {
        tl.cb = mt / p.tiles_per_img;
        const int r = mt - tl.cb * p.tiles_per_img;
        tl.ch0 = (r / p.tiles_w) * CONV_TH;
        tl.cw0 = (r % p.tiles_w) * CONV_TW;
    } else {
        tl.m0 = (int64_t)mt * TM;
    }
    return tl;
}

// EPI selects the epilogue instantiation: 0 generic (every option a run-time branch), 1 plain (bias / per-channel affine,
// optional residual, relu / leaky-relu, column statistics: the point-branch Linear layers and every convolution),
// 2 fused LayerNorm (+ residual, relu: the transformer's merge and mlp[2] layers), 3 accumulate (+ activation: the second half
// of the transformer's mlp[0]).  The generic epilogue is 2100 SASS instructions per 32-column chunk, most of them branches not
// taken, against ~350 in the specialised ones -- and the chunk loop is what the K <= 128 contractions spend their time in
// (instruction fetch of that sparse walk: 163840 x 128 x 32 went from 30.9 to 21.8 us).
template <int BN, bool CONV, int EPI>
__global__ void __launch_bounds__(THREADS, 1)
gemm_x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const Params p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* a_full = bars;              // [NSA <= 4]  TMA -> splitters
    uint64_t* a_empty = bars + 4;         // [NSA]  splitters -> TMA
    uint64_t* b_full = bars + 8;          // [NSB<=6] TMA -> MMA
    uint64_t* b_empty = bars + 14;        // [NSB]  MMA commit -> TMA
    uint64_t* ta_full = bars + 20;        // [NT]   splitters -> MMA (A hi/lo in TMEM)
    uint64_t* ta_empty = bars + 24;       // [NT]   MMA commit -> splitters
    uint64_t* acc_full = bars + 28;       // [2]    MMA commit -> epilogue
    uint64_t* acc_empty = bars + 30;      // [2]    epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
    float* s_scale = reinterpret_cast<float*>(smem + C::OFF_VEC);
    float* s_shift = s_scale + BN;
    float* s_gamma = s_shift + BN;
    float* s_beta = s_gamma + BN;
    float* s_col_base = s_beta + BN;  // [2 tile parities][4 quarters][BN][2]

    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    unsigned long long pacc[3] = {0ull, 0ull, 0ull};
    const long long t_start = p.prof ? clock64() : 0;
    const int total = p.m_tiles * p.n_tiles;
    const int my_tiles = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < C::NSA; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 4);
        }
        for (int s = 0; s < C::NSB; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < NT; ++s) {
            mbar_init(&ta_full[s], 4);
            mbar_init(&ta_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], EPI_WARPS);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc512(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================ TMA producer ================================
        {
            const bool leader = elect_one();
            uint32_t g = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const Tile tl = decode<CONV>(p, (int)blockIdx.x + i * (int)gridDim.x);
                const int n0 = tl.n0 * BN;
                for (int kb = 0; kb < p.num_kb; ++kb, ++g) {
                    const uint32_t sa = g % C::NSA, sb = g % C::NSB;
                    X3_WAIT(&a_empty[sa], ((g / C::NSA) & 1u) ^ 1u, 0);
                    uint8_t* a_dst = smem + sa * A_BYTES;
                    int kcol;
                    if (CONV) {
                        const int tap = kb / p.cpt, cc = kb - tap * p.cpt;
                        const int kh = tap / p.KW, kw = tap - kh * p.KW;
                        if (leader) {
                            mbar_expect_tx(&a_full[sa], A_BYTES);
                            tma_load_4d(&tmA, &a_full[sa], a_dst, cc * TK, tl.cw0 * p.cstride + kw - p.pad,
                                        tl.ch0 * p.cstride + kh - p.pad, tl.cb);
                        }
                        kcol = tap * p.Cin + cc * TK;
                    } else {
                        if (leader) {
                            mbar_expect_tx(&a_full[sa], A_BYTES);
                            tma_load_2d(&tmA, &a_full[sa], a_dst, kb * TK, (int)tl.m0);
                        }
                        kcol = kb * TK;
                    }
                    X3_WAIT(&b_empty[sb], ((g / C::NSB) & 1u) ^ 1u, 1);
                    uint8_t* b_dst = smem + C::OFF_B + sb * C::B_STAGE;
                    if (leader) {
                        mbar_expect_tx(&b_full[sb], C::B_STAGE);
                        tma_load_3d(&tmB, &b_full[sb], b_dst, kcol, n0, 0);
                        tma_load_3d(&tmB, &b_full[sb], b_dst + C::B_PLANE, kcol, n0, 1);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        {
            const bool leader = elect_one();
            constexpr uint32_t idesc = umma_idesc(2 /*tf32*/, TM, BN);
            uint32_t g = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const uint32_t buf = (uint32_t)i & 1u;
                X3_WAIT(&acc_empty[buf], (((uint32_t)i >> 1) & 1u) ^ 1u, 2);
                tc_fence_after();
                const uint32_t d = tmem_base + buf * BN;
                for (int kb = 0; kb < p.num_kb; ++kb, ++g) {
                    const uint32_t sb = g % C::NSB, st = g % NT;
                    X3_WAIT(&b_full[sb], (g / C::NSB) & 1u, 0);
                    X3_WAIT(&ta_full[st], (g / NT) & 1u, 1);
                    tc_fence_after();
                    const uint32_t bh = smem_u32(smem + C::OFF_B + sb * C::B_STAGE);
                    const uint32_t bl = bh + C::B_PLANE;
                    const uint32_t ah = tmem_base + A_COL0 + st * 64;
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < TK / 8; ++k) {
                            const uint64_t bhd = umma_desc_k128(bh + k * 32);
                            const uint64_t bld = umma_desc_k128(bl + k * 32);
                            mma_tf32_ts(d, ah + k * 8, bhd, idesc, (kb | k) != 0 ? 1u : 0u);
                            mma_tf32_ts(d, ah + 32 + k * 8, bhd, idesc, 1u);
                            mma_tf32_ts(d, ah + k * 8, bld, idesc, 1u);
                        }
                        tc_commit(&b_empty[sb]);
                        tc_commit(&ta_empty[st]);
                    }
                    __syncwarp();
                }
                if (leader) tc_commit(&acc_full[buf]);
                __syncwarp();
            }
        }
    } else if (warp < 2 + EPI_WARPS) {
        // ================================ epilogue ====================================
        const int set = (warp - 2) >> 2;
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int et = threadIdx.x - 64;
        const Epilogue& ep = p.ep;
        const bool has_rd = EPI == 0 && ep.rowdiv != nullptr, has_res = ep.residual != nullptr;
        const bool has_acc = (EPI == 0 || EPI == 3) && ep.accumulate != 0;
        const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
        const bool rvec_ok = has_res && ((ep.ldres & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0);
        const bool has_ln = (EPI == 0 || EPI == 2) && ep.ln_gamma != nullptr;  // host guarantees n_tiles == 1 and N <= BN
        const bool ln_idle = has_ln && set != 0;     // a fused LayerNorm needs the whole row in one thread
        const bool tma_st = !CONV && p.tma_store != 0;
        float* stage_base = reinterpret_cast<float*>(smem + C::OFF_STG + (warp - 2) * STAGE_PER_WARP);
        int tbuf = 0;            // TMA-store path: the two staging boxes alternate, so a chunk never waits for the previous store
        int stores_pending = 0;
        const bool restage = p.n_tiles > 1;
        for (int i = 0; i < my_tiles; ++i) {
            const Tile tl = decode<CONV>(p, (int)blockIdx.x + i * (int)gridDim.x);
            const int n0 = tl.n0 * BN;
            const uint32_t buf = (uint32_t)i & 1u;
            // The per-column vectors only change with the column tile: one staging (and one barrier) for the whole kernel when
            // there is a single column tile.  The column statistics alternate between two buffers, so the reduction of tile i
            // (after the barrier at its end) never meets the writes of tile i + 1.
            float* s_col = s_col_base + (i & 1) * (8 * BN);
            if (i == 0 || restage) {
                if (et < BN) {  // per-column epilogue vectors of this tile
                    const int n = n0 + et;
                    float sc = 1.0f, sh = 0.0f;
                    if (n < p.N) {
                        if (ep.colscale) {
                            sc = __ldg(ep.colscale + n);
                            sh = __ldg(ep.colshift + n);
                        }
                        if (ep.bias) sh += __ldg(ep.bias + n);
                    }
                    s_scale[et] = sc;
                    s_shift[et] = sh;
                    s_gamma[et] = (has_ln && n < p.N) ? __ldg(ep.ln_gamma + n) : 0.0f;
                    s_beta[et] = (has_ln && n < p.N) ? __ldg(ep.ln_beta + n) : 0.0f;
                }
                if (!(p.dbg & 16)) epi_bar();
            }
            X3_WAIT(&acc_full[buf], ((uint32_t)i >> 1) & 1u, 0);
            tc_fence_after();
            const uint32_t tacc = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
            const int r = q * 32 + lane;
            int64_t grow;
            bool row_ok;
            if (CONV) {
                const int hl = r / CONV_TW, wl = r - hl * CONV_TW;
                grow = ((int64_t)tl.cb * p.Ho + tl.ch0 + hl) * p.Wo + tl.cw0 + wl;
                row_ok = true;
            } else {
                grow = tl.m0 + r;
                row_ok = grow < p.M;
            }
            const float rd = (has_rd && row_ok) ? __ldg(ep.rowdiv + grow) : 1.0f;
            float* crow = p.C + grow * p.ldc;
            const float* rrow = has_res ? ep.residual + grow * ep.ldres : nullptr;
            float ln_mean = 0.0f, ln_rstd = 1.0f;
            if constexpr (EPI == 0 || EPI == 2)
            if (has_ln && !ln_idle) {
                float sum = 0.0f;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t acc[32];
                    tmem_ld32(tacc + (uint32_t)c0, acc);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < p.N) sum += fmaf(__uint_as_float(acc[j]), s_scale[c0 + j], s_shift[c0 + j]);
                }
                ln_mean = sum / (float)p.N;
                float ssq = 0.0f;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t acc[32];
                    tmem_ld32(tacc + (uint32_t)c0, acc);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < p.N) {
                            const float dd = fmaf(__uint_as_float(acc[j]), s_scale[c0 + j], s_shift[c0 + j]) - ln_mean;
                            ssq = fmaf(dd, dd, ssq);
                        }
                }
                ln_rstd = rsqrtf(ssq / (float)p.N + ep.ln_eps);
            }
            const int c_begin = ln_idle ? BN : (has_ln ? 0 : set * 32), c_step = has_ln ? 32 : 64;
#pragma unroll 1
            for (int c0 = c_begin; c0 < BN; c0 += c_step) {
                uint32_t acc[32];
                tmem_ld32(tacc + (uint32_t)c0, acc);
                tmem_ld_wait();
                const int nb = n0 + c0;
                float* stage = stage_base + (tma_st ? tbuf * 1024 : 0);
                if (tma_st && stores_pending >= 2) {  // the store issued from THIS box two chunks ago must have read it
                    if (lane == 0) bulk_wait_read<1>();
                    stores_pending = 1;
                }
                __syncwarp();
                if (row_ok && nb < p.N && !(p.dbg & 8)) {
                    const bool full = nb + 32 <= p.N;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
                    if constexpr (EPI == 0) {
                        if (has_rd) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = v[j] / rd;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 sc = *reinterpret_cast<const float4*>(s_scale + c0 + j);
                        const float4 sh = *reinterpret_cast<const float4*>(s_shift + c0 + j);
                        v[j] = fmaf(v[j], sc.x, sh.x);
                        v[j + 1] = fmaf(v[j + 1], sc.y, sh.y);
                        v[j + 2] = fmaf(v[j + 2], sc.z, sh.z);
                        v[j + 3] = fmaf(v[j + 3], sc.w, sh.w);
                    }
                    if (has_ln) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = (v[j] - ln_mean) * ln_rstd * s_gamma[c0 + j] + s_beta[c0 + j];
                        if (ep.act == COFI_ACT_RELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                        }
                    }
                    if (has_res) {
                        if (rvec_ok && full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 r4 = __ldg(reinterpret_cast<const float4*>(rrow + nb + j));
                                v[j] += r4.x;
                                v[j + 1] += r4.y;
                                v[j + 2] += r4.z;
                                v[j + 3] += r4.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (nb + j < p.N) v[j] += __ldg(rrow + nb + j);
                        }
                    }
                    if (has_acc) {
                        if (vec_ok && full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 r4 = *reinterpret_cast<const float4*>(crow + nb + j);
                                v[j] += r4.x;
                                v[j + 1] += r4.y;
                                v[j + 2] += r4.z;
                                v[j + 3] += r4.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (nb + j < p.N) v[j] += crow[nb + j];
                        }
                    }
                    if (has_ln) {
                        // activation was applied right after the norm, before the residual
                    } else if (ep.act == COFI_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
                    } else if (ep.act == COFI_ACT_LRELU01) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.0f ? v[j] : v[j] * 0.1f;
                    } else if (EPI == 0 && ep.act == COFI_ACT_SIGMOID) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 1.0f / (1.0f + expf(-v[j]));
                    }
                    if (tma_st) {
                        float* trow = stage + lane * 32;  // row `lane` of the box; 16-byte chunk j lands at j ^ (row & 7)
                        const int swz = lane & 7;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<float4*>(trow + ((j ^ swz) << 2)) =
                                make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
                        float* st = stage + lane * 36;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(st + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
                if (tma_st && !(p.dbg & 4)) fence_proxy_async_smem();
                __syncwarp();
                if (tma_st) {   // the bulk store goes first: it reads the staged box while the statistics below read it too
                    if (lane == 0) {   // an empty group when the chunk lies past N keeps the per-thread group count in step
                        if (nb < p.N && !(p.dbg & 2)) tma_store_2d(&tmC, stage, nb, (int)tl.m0 + q * 32);
                        bulk_commit();
                    }
                    ++stores_pending;
                    tbuf ^= 1;
                }
                if (p.stat_out && !(p.dbg & 1)) {
                    // per-tile column statistics (host guarantees M % 128 == 0, act none): lane = column
                    float cs = 0.0f, cq = 0.0f;
                    if (nb + lane < p.N) {
                        // four independent partial sums (rows rr = 4i + u): the 32 shared-memory loads are issued back to
                        // back and the add / fma chains are 8 long instead of 32
                        float s4[4] = {0.0f, 0.0f, 0.0f, 0.0f}, q4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr) {
                            const float x = tma_st ? stage[rr * 32 + ((((lane >> 2) ^ (rr & 7)) << 2) | (lane & 3))] : stage[rr * 36 + lane];
                            s4[rr & 3] += x;
                            q4[rr & 3] = fmaf(x, x, q4[rr & 3]);
                        }
                        cs = (s4[0] + s4[1]) + (s4[2] + s4[3]);
                        cq = (q4[0] + q4[1]) + (q4[2] + q4[3]);
                    }
                    s_col[(q * BN + c0 + lane) * 2 + 0] = cs;
                    s_col[(q * BN + c0 + lane) * 2 + 1] = cq;
                }
                if (!tma_st && nb < p.N) {
                    const bool full = nb + 32 <= p.N;
                    const int cc = (lane & 7) * 4;
                    float* drow = CONV ? nullptr : p.C + (tl.m0 + q * 32 + (lane >> 3)) * p.ldc + nb + cc;
                    const int64_t rows_left = CONV ? 0 : p.M - (tl.m0 + q * 32 + (lane >> 3));
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int rr = it * 4 + (lane >> 3);
                        float* dst;
                        if (CONV) {
                            const int rt = q * 32 + rr;
                            const int hl = rt / CONV_TW, wl = rt - hl * CONV_TW;
                            dst = p.C + (((int64_t)tl.cb * p.Ho + tl.ch0 + hl) * p.Wo + tl.cw0 + wl) * p.ldc + nb + cc;
                        } else {
                            if (it * 4 >= rows_left) break;
                            dst = drow + (int64_t)it * 4 * p.ldc;
                        }
                        const float4 val = *reinterpret_cast<const float4*>(stage + rr * 36 + cc);
                        if (vec_ok && full) {
                            *reinterpret_cast<float4*>(dst) = val;
                        } else {
                            if (nb + cc < p.N) dst[0] = val.x;
                            if (nb + cc + 1 < p.N) dst[1] = val.y;
                            if (nb + cc + 2 < p.N) dst[2] = val.z;
                            if (nb + cc + 3 < p.N) dst[3] = val.w;
                        }
                    }
                }
                __syncwarp();
            }
            // this warp's reads of the accumulator are complete (tcgen05.wait::ld above): hand the buffer back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            if ((p.stat_out || restage) && !(p.dbg & 16)) epi_bar();
            if (p.stat_out && et < BN && n0 + et < p.N) {
                float cs = 0.0f, cq = 0.0f;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    cs += s_col[(qq * BN + et) * 2 + 0];
                    cq += s_col[(qq * BN + et) * 2 + 1];
                }
                const int mt = CONV ? 0 : (int)(tl.m0 / TM);
                float* o = p.stat_out + ((int64_t)mt * p.N + n0 + et) * 2;
                o[0] = cs;
                o[1] = cq;
            }
        }
        if (tma_st && stores_pending) {  // shared memory must outlive the reads of the last stores
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
        }
    } else {
        // ================================ A splitters =================================
        const int sset = (warp - (2 + EPI_WARPS)) >> 2;
        const int q = warp & 3;
        const int r = q * 32 + lane;  // tile row of this thread
        const uint32_t total_kb = (uint32_t)my_tiles * (uint32_t)p.num_kb;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        for (uint32_t g = (uint32_t)sset; g < total_kb; g += 2) {
            const uint32_t sa = g % C::NSA, st = g % NT;
            X3_WAIT(&a_full[sa], (g / C::NSA) & 1u, 0);
            const uint4* row = reinterpret_cast<const uint4*>(smem + sa * A_BYTES + r * 128);
            uint4 x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = row[j ^ (r & 7)];  // undo SWIZZLE_128B: logical 16-byte chunk j of row r
            X3_WAIT(&ta_empty[st], ((g / NT) & 1u) ^ 1u, 1);
            tc_fence_after();
            const uint32_t ta = tmem_base + A_COL0 + st * 64 + lane_addr;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 v = x[h * 4 + j];
                    const uint32_t e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t hh = (e[c] + 0x1000u) & 0xFFFFE000u;
                        hi[j * 4 + c] = hh;
                        lo[j * 4 + c] = (__float_as_uint(__uint_as_float(e[c]) - __uint_as_float(hh)) + 0x1000u) & 0xFFFFE000u;
                    }
                }
                tmem_st16(ta + h * 16, hi);
                tmem_st16(ta + 32 + h * 16, lo);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&a_empty[sa]);
                mbar_arrive(&ta_full[st]);
            }
        }
    }
    if (p.prof && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == 2 + EPI_WARPS)) {
        // slots: [0..2] TMA (a_empty, b_empty), [3..5] MMA (b_full, ta_full, acc_empty), [6] epilogue warp 2 (acc_full),
        // [9..10] splitter warp 10 (a_full, ta_empty), [14] k-blocks, [15] CTA clocks, [13] CTAs
        const int base = warp == 0 ? 0 : (warp == 1 ? 3 : (warp == 2 ? 6 : 9));
        for (int j = 0; j < 3; ++j) atomicAdd(p.prof + base + j, pacc[j]);
        if (warp == 0) {
            atomicAdd(p.prof + 15, (unsigned long long)(clock64() - t_start));
            atomicAdd(p.prof + 14, (unsigned long long)my_tiles * (unsigned long long)p.num_kb);
            atomicAdd(p.prof + 13, 1ull);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// element-wise split of a weight matrix into its two tf32-exact planes
__global__ void split_tf32_kernel(const float* __restrict__ w, int64_t ldw, int N, int K, float* __restrict__ out) {
    const int64_t n = (int64_t)N * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / K, col = i - row * K;
        const float x = w[row * ldw + col];
        const uint32_t h = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
        const uint32_t l = (__float_as_uint(x - __uint_as_float(h)) + 0x1000u) & 0xFFFFE000u;
        out[i] = __uint_as_float(h);
        out[n + i] = __uint_as_float(l);
    }
}

static unsigned long long* prof_buffer() {
    static unsigned long long* buf = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* e = getenv("COFI_X3_PROFILE");
        if (e && e[0] == '1' && cudaMalloc(&buf, 16 * sizeof(unsigned long long)) == cudaSuccess)
            cudaMemset(buf, 0, 16 * sizeof(unsigned long long));
        else
            buf = nullptr;
    }
    return buf;
}

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int BN, bool CONV, int EPI>
static int launch_one_v(const CUtensorMap* a, const CUtensorMap* b, const CUtensorMap* c, const Params& p_in, cudaStream_t st) {
    Params p = p_in;
    p.tma_store = (c != nullptr && !CONV) ? 1 : 0;
    p.prof = prof_buffer();
    {
        static int dbg = -1;
        if (dbg < 0) {
            const char* e = getenv("COFI_X3_DEBUG");
            dbg = e ? atoi(e) : 0;
        }
        p.dbg = dbg;
    }
    if (!c) c = a;
    using C = Cfg<BN>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gemm_x3_kernel<BN, CONV, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gemm_x3, smem=%d): %s", C::SMEM, cudaGetErrorString(e));
            return COFI_ECUDA;
        }
        attr_done = true;
    }
    const int total = p.m_tiles * p.n_tiles;
    const int grid = total < num_sms() ? total : num_sms();
    gemm_x3_kernel<BN, CONV, EPI><<<grid, THREADS, C::SMEM, st>>>(*a, *b, *c, p);
    return check_launch(CONV ? "cofi_conv2d_nhwc(3xTF32)" : "cofi_gemm(3xTF32)");
}

template <int BN, bool CONV>
static int launch_one(const CUtensorMap* a, const CUtensorMap* b, const CUtensorMap* c, const Params& p, cudaStream_t st) {
    const Epilogue& ep = p.ep;
    if (ep.rowdiv || ep.act == COFI_ACT_SIGMOID || (ep.ln_gamma && ep.accumulate)) return launch_one_v<BN, CONV, 0>(a, b, c, p, st);
    if (!CONV && ep.ln_gamma) return launch_one_v<BN, CONV, 2>(a, b, c, p, st);
    if (!CONV && ep.accumulate) return launch_one_v<BN, CONV, 3>(a, b, c, p, st);
    if (ep.ln_gamma || ep.accumulate) return launch_one_v<BN, CONV, 0>(a, b, c, p, st);
    return launch_one_v<BN, CONV, 1>(a, b, c, p, st);
}

// Tile width: the persistent grid runs ceil(tiles / SMs) rounds of one tile each.  Per tile the kernel is bound by the
// slowest of: twelve MMAs per k-block (bn/2 clk each), the TMA stream (A tile + two W planes per k-block at ~40 B/clk per
// SM from L2) and the epilogue (~12 clk per column); pick the width that minimises rounds x tile time, so that small problems
// (160 row tiles) are not quantised to two full rounds while long contractions keep the 128-wide tile.
static int pick_bn(int N, int64_t m_tiles, int num_kb, bool full_row) {
    const int top = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
    if (full_row) return top;
    int best = top;
    int64_t best_cost = -1;
    for (int bn = top; bn >= 32; bn >>= 1) {
        const int64_t tiles = m_tiles * ((N + bn - 1) / bn);
        const int64_t rounds = (tiles + num_sms() - 1) / num_sms();
        const int64_t mma = (int64_t)num_kb * 6 * bn, load = (int64_t)num_kb * (A_BYTES + 256 * bn) / 40, epi = 12 * bn;
        const int64_t tile = (mma > load ? (mma > epi ? mma : epi) : (load > epi ? load : epi)) + 600;
        const int64_t cost = rounds * tile;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = bn;
        }
    }
    return best;
}

}  // namespace x3
}  // namespace tc

// Output tensor map of the TMA-store epilogue (gemm_tc.cu)
const CUtensorMap* gemm_c_tmap(const float* C, int64_t ldc, int64_t M, int N);

// W2 = [2][N][K] dense (cofi_split_tf32)
int gemm_x3_launch(const float* A, int64_t lda, const float* W2, float* C, int64_t ldc, int64_t M, int N, int K,
                   const Epilogue& ep, cudaStream_t st, float* stat_out) {
    using namespace tc;
    using namespace tc::x3;
    const int64_t m_tiles = ceil_div(M, TM);
    const int bn = pick_bn(N, m_tiles, (K + TK - 1) / TK, ep.ln_gamma != nullptr);
    uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[1] = {(uint64_t)lda * 4};
    uint32_t bA[2] = {TK, TM};
    uint64_t dB[3] = {(uint64_t)K, (uint64_t)N, 2}, sB[2] = {(uint64_t)K * 4, (uint64_t)N * K * 4};
    uint32_t bB[3] = {TK, (uint32_t)bn, 1};
    const CUtensorMap* ta = get_tmap_f32(A, 2, dA, sA, bA);
    const CUtensorMap* tb = get_tmap_f32(W2, 3, dB, sB, bB);
    if (!ta || !tb) return COFI_ECUDA;
    Params p{};
    p.C = C;
    p.ldc = ldc;
    p.M = M;
    p.N = N;
    p.num_kb = (K + TK - 1) / TK;
    p.m_tiles = (int)m_tiles;
    p.n_tiles = (N + bn - 1) / bn;
    p.ep = ep;
    p.stat_out = stat_out;
    const CUtensorMap* tcm = gemm_c_tmap(C, ldc, M, N);
    if (bn == 32) return launch_one<32, false>(ta, tb, tcm, p, st);
    if (bn == 64) return launch_one<64, false>(ta, tb, tcm, p, st);
    return launch_one<128, false>(ta, tb, tcm, p, st);
}

// w2 = [2][Cout][KH*KW*Cin] dense
int conv_x3_launch(const float* x, int B, int H, int W, int Cin, const float* w2, int Cout, int KH, int KW, int stride, int pad,
                   float* y, const Epilogue& ep, cudaStream_t st) {
    using namespace tc;
    using namespace tc::x3;
    const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
    const int64_t m_tiles = (int64_t)B * (Wo / CONV_TW) * (Ho / CONV_TH);
    const int bn = pick_bn(Cout, m_tiles, KH * KW * ((Cin + TK - 1) / TK), false);
    const int Ktot = KH * KW * Cin;
    uint64_t dA[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t sA[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    uint32_t bA[4] = {TK, (uint32_t)(CONV_TW * stride), (uint32_t)(CONV_TH * stride), 1};
    uint32_t eA[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    uint64_t dB[3] = {(uint64_t)Ktot, (uint64_t)Cout, 2}, sB[2] = {(uint64_t)Ktot * 4, (uint64_t)Cout * Ktot * 4};
    uint32_t bB[3] = {TK, (uint32_t)bn, 1};
    const CUtensorMap* ta = stride == 1 ? get_tmap_f32(x, 4, dA, sA, bA) : get_tmap_f32_es(x, 4, dA, sA, bA, eA);
    const CUtensorMap* tb = get_tmap_f32(w2, 3, dB, sB, bB);
    if (!ta || !tb) return COFI_ECUDA;
    Params p{};
    p.C = y;
    p.ldc = Cout;
    p.M = (int64_t)B * Ho * Wo;
    p.N = Cout;
    p.cpt = (Cin + TK - 1) / TK;
    p.num_kb = KH * KW * p.cpt;
    p.m_tiles = (int)m_tiles;
    p.n_tiles = (Cout + bn - 1) / bn;
    p.ep = ep;
    p.Ho = Ho;
    p.Wo = Wo;
    p.Cin = Cin;
    p.KW = KW;
    p.pad = pad;
    p.cstride = stride;
    p.tiles_w = Wo / CONV_TW;
    p.tiles_per_img = p.tiles_w * (Ho / CONV_TH);
    if (bn == 32) return launch_one<32, true>(ta, tb, nullptr, p, st);
    if (bn == 64) return launch_one<64, true>(ta, tb, nullptr, p, st);
    return launch_one<128, true>(ta, tb, nullptr, p, st);
}

}  // namespace cofi

using namespace cofi;

// perf triage (COFI_X3_PROFILE=1): copies the 16 accumulated counters of gemm_x3_kernel to the host and clears them
extern "C" int cofi_debug_x3_profile(unsigned long long* out16) {
    unsigned long long* buf = tc::x3::prof_buffer();
    COFI_REQUIRE(buf && out16, "cofi_debug_x3_profile: set COFI_X3_PROFILE=1 before the first 3xTF32 launch");
    cudaDeviceSynchronize();
    cudaMemcpy(out16, buf, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaMemset(buf, 0, 16 * sizeof(unsigned long long));
    return COFI_OK;
}

extern "C" int cofi_split_tf32(const float* w, int64_t ldw, int N, int K, float* out, void* stream) {
    COFI_REQUIRE(w && out && N > 0 && K > 0 && ldw >= K, "cofi_split_tf32: bad argument");
    const int64_t n = (int64_t)N * K;
    const int blocks = (int)((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048);
    tc::x3::split_tf32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, ldw, N, K, out);
    return check_launch("cofi_split_tf32");
}
