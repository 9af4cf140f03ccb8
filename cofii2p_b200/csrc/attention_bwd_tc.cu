// attention_bwd_tc.cu -- tcgen05 backward of the LoFTR attention (training path of
// reference model/transformer/linear_attention.py:69-77; forward: attention_tc.cu).
//
// With P = exp(scale * Q K^T - lse) (the forward's log-sum-exp makes P final, no online rescaling) and
// D_i = dO_i . O_i:   dV = P^T dO,   dP = dO V^T,   G = scale * P o (dP - D),   dQ = G K,   dK = G^T Q.
// Two launches of one kernel template, both shaped like the forward (TMA producer warp, MMA issuer warp, four
// element-wise warps owning one TMEM lane = one row each, 128 resident rows per CTA, streamed tiles of 64 rows):
//   DKV = false, CTA = 128 queries, keys streamed:   X = Q K^T, Y = dO V^T in TMEM -> G to shared memory as the next
//                 A operand -> dQ += G K accumulated in TMEM over all key tiles (K^T chunks as the K-major B operand).
//   DKV = true,  CTA = 128 keys, queries streamed:   X = K Q^T (= S^T), Y = V dO^T (= dP^T) -> P^T and G^T to shared
//                 memory -> dV += P^T dO, dK += G^T Q accumulated in TMEM (dO^T / Q^T chunks as B operands).
// Every MMA operand is K-major SWIZZLE_128B exactly as in the forward; the transposed copies K^T, Q^T, dO^T are made
// by the tiled transpose kernel into a caller-provided workspace.  The [L,S] matrices never reach HBM.
#include "common.cuh"
#include "tc_common.cuh"

namespace cofi {
namespace tc {

constexpr int BR = 128;  // resident rows per CTA
constexpr int BT = 64;   // streamed rows per tile
constexpr int BD = 32;   // head dimension
constexpr int B_STAGES = 2;
constexpr int R_BYTES = BR * BD * 4;   // 16 KB
constexpr int T_BYTES = BT * BD * 4;   // 8 KB  [64 rows x 32]
constexpr int U_BYTES = BD * BT * 4;   // 8 KB  two chunks of [32 d-rows x 32 streamed rows]
constexpr int A_BYTES = BR * BT * 4;   // 32 KB two k-blocks of [128 rows x 32]
constexpr int BWD_TMEM_COLS = 256;     // X [0,64) Y [64,128) acc0 [128,160) acc1 [160,192)

template <bool DKV>
struct BCfg {
    static constexpr int STAGE = 2 * T_BYTES + (DKV ? 2 : 1) * U_BYTES;
    static constexpr int NA = DKV ? 2 : 1;
    static constexpr int SMEM = 2 * R_BYTES + B_STAGES * STAGE + NA * A_BYTES + 1024 /* column vectors */ + 256 + 1024;
};

struct BwdParams {
    const float* out;   // forward output   [frames*L, heads*32]   (dQ pass: D = dO . O)
    const float* dout;  // its gradient
    const float* lse;   // [frames*L, heads]
    float* dsum;        // [frames*L, heads]  written by the dQ pass, read by the dK/dV pass
    float* o0;          // dQ | dV
    float* o1;          // -- | dK
    int64_t NR, NT, L;  // resident / streamed rows per frame; queries per frame
    int heads;
    float scale;
    int num_tiles;
};

template <bool DKV>
__global__ void __launch_bounds__(192)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmR1, const __grid_constant__ CUtensorMap tmR2,
                        const __grid_constant__ CUtensorMap tmT1, const __grid_constant__ CUtensorMap tmT2,
                        const __grid_constant__ CUtensorMap tmU1, const __grid_constant__ CUtensorMap tmU2,
                        const BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int STAGE = BCfg<DKV>::STAGE, NA = BCfg<DKV>::NA;
    uint8_t* sR1 = smem;
    uint8_t* sR2 = smem + R_BYTES;
    uint8_t* sT = sR2 + R_BYTES;                    // [stage][T1 | T2 | U1 | (U2)]
    uint8_t* sA = sT + B_STAGES * STAGE;            // [NA][A_BYTES]
    float* colL = reinterpret_cast<float*>(sA + NA * A_BYTES);   // [2][64]
    float* colD = colL + 2 * BT;                                 // [2][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + NA * A_BYTES + 1024);
    uint64_t* r_full = bars;
    uint64_t* t_full = bars + 1;    // [2]
    uint64_t* t_empty = bars + 3;   // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;
    uint64_t* acc_full = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    const int head = blockIdx.y, frame = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.x * BR;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmR1);
        tma_prefetch_desc(&tmR2);
        tma_prefetch_desc(&tmT1);
        tma_prefetch_desc(&tmT2);
        tma_prefetch_desc(&tmU1);
        if (DKV) tma_prefetch_desc(&tmU2);
        mbar_init(r_full, 1);
        for (int s = 0; s < B_STAGES; ++s) {
            mbar_init(&t_full[s], 1);
            mbar_init(&t_empty[s], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(p_full, 4);
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, BWD_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_X = tmem_base, tmem_Y = tmem_base + BT, tmem_A0 = tmem_base + 2 * BT, tmem_A1 = tmem_A0 + BD;

    if (warp == 0) {
        {   // whole warp runs the loop, the elected lane issues (tc_common.cuh: elect_one)
            const bool leader = elect_one();
            if (leader) {
                mbar_expect_tx(r_full, 2 * R_BYTES);
                tma_load_2d(&tmR1, r_full, sR1, head * BD, (int)(frame * p.NR + r0));
                tma_load_2d(&tmR2, r_full, sR2, head * BD, (int)(frame * p.NR + r0));
            }
            __syncwarp();
            for (int t = 0; t < p.num_tiles; ++t) {
                const int s = t % B_STAGES;
                const uint32_t ph = (uint32_t)(t / B_STAGES) & 1u;
                mbar_wait(&t_empty[s], ph ^ 1u);
                uint8_t* d = sT + s * STAGE;
                const int row0 = (int)(frame * p.NT + (int64_t)t * BT);
                if (leader) {
                    mbar_expect_tx(&t_full[s], STAGE);
                    tma_load_2d(&tmT1, &t_full[s], d, head * BD, row0);
                    tma_load_2d(&tmT2, &t_full[s], d + T_BYTES, head * BD, row0);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        tma_load_2d(&tmU1, &t_full[s], d + 2 * T_BYTES + c * (BD * 32 * 4), row0 + c * 32, head * BD);
                        if (DKV)
                            tma_load_2d(&tmU2, &t_full[s], d + 2 * T_BYTES + U_BYTES + c * (BD * 32 * 4), row0 + c * 32, head * BD);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        {
            const bool leader = elect_one();
            constexpr uint32_t idesc_x = umma_idesc(2, BR, BT);   // 128 x 64
            constexpr uint32_t idesc_a = umma_idesc(2, BR, BD);   // 128 x 32
            mbar_wait(r_full, 0);
            const uint32_t r1 = smem_u32(sR1), r2 = smem_u32(sR2), a0 = smem_u32(sA), a1 = a0 + A_BYTES;
            for (int t = 0; t < p.num_tiles; ++t) {
                const int s = t % B_STAGES;
                const uint32_t ph = (uint32_t)(t / B_STAGES) & 1u;
                const uint32_t tp = (uint32_t)t & 1u;
                mbar_wait(&t_full[s], ph);
                tc_fence_after();
                const uint32_t t1 = smem_u32(sT + s * STAGE), t2 = t1 + T_BYTES, u1 = t2 + T_BYTES, u2 = u1 + U_BYTES;
                // X = R1 T1^T, Y = R2 T2^T   (their TMEM columns are free: the operands of tile t-1 were published,
                // i.e. X/Y of tile t-1 were fully read)
                if (leader) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_tf32(tmem_X, umma_desc_k128(r1 + k * 32), umma_desc_k128(t1 + k * 32), idesc_x, k != 0 ? 1u : 0u);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_tf32(tmem_Y, umma_desc_k128(r2 + k * 32), umma_desc_k128(t2 + k * 32), idesc_x, k != 0 ? 1u : 0u);
                    tc_commit(s_full);
                }
                __syncwarp();
                // accumulate: acc0 += A0 U1, (acc1 += A1 U2)
                mbar_wait(p_full, tp);
                tc_fence_after();
                if (leader) {
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_tf32(tmem_A0, umma_desc_k128(a0 + c * (BR * 32 * 4) + k * 32),
                                     umma_desc_k128(u1 + c * (BD * 32 * 4) + k * 32), idesc_a, (t | c | k) != 0 ? 1u : 0u);
                    if (DKV) {
#pragma unroll
                        for (int c = 0; c < 2; ++c)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                mma_tf32(tmem_A1, umma_desc_k128(a1 + c * (BR * 32 * 4) + k * 32),
                                         umma_desc_k128(u2 + c * (BD * 32 * 4) + k * 32), idesc_a, (t | c | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(&t_empty[s]);
                }
                __syncwarp();
            }
            if (leader) tc_commit(acc_full);
            __syncwarp();
        }
    } else {
        // ================================ element-wise warps ================================
        const int q = warp & 3;
        const int r = q * 32 + lane;                 // resident row inside the tile == TMEM lane
        const bool row_ok = r0 + r < p.NR;
        const int64_t grow = (int64_t)frame * p.NR + r0 + r;
        const int64_t ld = (int64_t)p.heads * BD;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        uint8_t* arow = sA + r * 128;
        const int sw = r & 7;
        const int et = threadIdx.x - 64;             // 0..127
        float lse_r = 0.0f, dsum_r = 0.0f;
        if (!DKV && row_ok) {
            const float4* go = reinterpret_cast<const float4*>(p.dout + grow * ld + head * BD);
            const float4* oo = reinterpret_cast<const float4*>(p.out + grow * ld + head * BD);
#pragma unroll
            for (int d = 0; d < BD / 4; ++d) {
                const float4 a = __ldg(go + d), b = __ldg(oo + d);
                dsum_r = fmaf(a.x, b.x, dsum_r);
                dsum_r = fmaf(a.y, b.y, dsum_r);
                dsum_r = fmaf(a.z, b.z, dsum_r);
                dsum_r = fmaf(a.w, b.w, dsum_r);
            }
            lse_r = __ldg(p.lse + grow * p.heads + head);
            p.dsum[grow * p.heads + head] = dsum_r;
        }
        for (int t = 0; t < p.num_tiles; ++t) {
            const uint32_t tp = (uint32_t)t & 1u;
            const int valid = (int)((p.NT - (int64_t)t * BT) < BT ? (p.NT - (int64_t)t * BT) : BT);
            if (DKV) {
                // per-column (= per-query) lse and D of this tile; double-buffered, see the barrier below
                const int col = et & 63;
                const int64_t gq = (int64_t)frame * p.NT + (int64_t)t * BT + col;
                if (et < 64) colL[tp * BT + col] = col < valid ? __ldg(p.lse + gq * p.heads + head) : INFINITY;
                else colD[tp * BT + col] = col < valid ? __ldg(p.dsum + gq * p.heads + head) : 0.0f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait(s_full, tp);   // X, Y ready; also: the accumulate MMAs of tile t-1 completed -> sA is free
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t rx[32], ry[32];
                tmem_ld32(tmem_X + lane_off + c * 32, rx);
                tmem_ld32(tmem_Y + lane_off + c * 32, ry);
                tmem_ld_wait();
                uint8_t* blk = arow + c * (BR * 32 * 4);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float pv[4], gv[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int col = c * 32 + j + e;
                        const float x = __uint_as_float(rx[j + e]) * p.scale;
                        const float y = __uint_as_float(ry[j + e]);
                        float pe;
                        if (DKV) {
                            pe = expf(x - colL[tp * BT + col]);            // +inf for columns past the end -> 0
                            gv[e] = pe * (y - colD[tp * BT + col]) * p.scale;
                        } else {
                            pe = col < valid ? expf(x - lse_r) : 0.0f;
                            gv[e] = pe * (y - dsum_r) * p.scale;
                        }
                        pv[e] = pe;
                    }
                    const int off = (((j >> 2) ^ sw) & 7) << 4;
                    if (DKV) {
                        *reinterpret_cast<float4*>(blk + off) = make_float4(pv[0], pv[1], pv[2], pv[3]);
                        *reinterpret_cast<float4*>(blk + A_BYTES + off) = make_float4(gv[0], gv[1], gv[2], gv[3]);
                    } else {
                        *reinterpret_cast<float4*>(blk + off) = make_float4(gv[0], gv[1], gv[2], gv[3]);
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(acc_full, 0);
        tc_fence_after();
        {
            uint32_t raw[32];
            tmem_ld32(tmem_A0 + lane_off, raw);
            tmem_ld_wait();
            if (row_ok) {
                float* op = p.o0 + grow * ld + head * BD;
#pragma unroll
                for (int d = 0; d < BD; d += 4)
                    *reinterpret_cast<float4*>(op + d) = make_float4(__uint_as_float(raw[d]), __uint_as_float(raw[d + 1]),
                                                                     __uint_as_float(raw[d + 2]), __uint_as_float(raw[d + 3]));
            }
        }
        if (DKV) {
            uint32_t raw[32];
            tmem_ld32(tmem_A1 + lane_off, raw);
            tmem_ld_wait();
            if (row_ok) {
                float* op = p.o1 + grow * ld + head * BD;
#pragma unroll
                for (int d = 0; d < BD; d += 4)
                    *reinterpret_cast<float4*>(op + d) = make_float4(__uint_as_float(raw[d]), __uint_as_float(raw[d + 1]),
                                                                     __uint_as_float(raw[d + 2]), __uint_as_float(raw[d + 3]));
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BWD_TMEM_COLS);
    }
}

}  // namespace tc

bool attention_bwd_tc_supported(int64_t L, int64_t S, int heads, int D) {
    return D == 32 && heads >= 1 && L >= 1 && S >= 1 && L % 4 == 0 && S % 4 == 0;
}

}  // namespace cofi

using namespace cofi;

extern "C" int64_t cofi_attention_bwd_tc_workspace(int64_t L, int64_t S, int frames, int heads, int D) {
    if (L < 1 || S < 1 || frames < 1 || heads < 1 || D < 1) return -1;
    return (int64_t)frames * (2 * L + S) * heads * D * 4;  // Q^T, dO^T, K^T
}

extern "C" int cofi_attention_bwd_tc(const float* q, const float* k, const float* v, const float* out, const float* dout,
                                     const float* lse, int64_t L, int64_t S, int frames, int heads, int D, float scale,
                                     float* dq, float* dk, float* dv, float* dsum_work, void* tr_work, void* stream) {
    using namespace tc;
    COFI_REQUIRE(q && k && v && out && dout && lse && dq && dk && dv && dsum_work && tr_work, "cofi_attention_bwd_tc: null pointer");
    COFI_REQUIRE(attention_bwd_tc_supported(L, S, heads, D), "cofi_attention_bwd_tc: needs D == 32 and L, S multiples of 4");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t C = (int64_t)heads * D, RL = (int64_t)frames * L, RS = (int64_t)frames * S;
    float* qt = (float*)tr_work;
    float* dot = qt + RL * C;
    float* kt = dot + RL * C;
    int rc;
    if ((rc = cofi_nhwc_to_nchw(q, 1, 1, (int)RL, (int)C, qt, stream))) return rc;
    if ((rc = cofi_nhwc_to_nchw(dout, 1, 1, (int)RL, (int)C, dot, stream))) return rc;
    if ((rc = cofi_nhwc_to_nchw(k, 1, 1, (int)RS, (int)C, kt, stream))) return rc;

    auto rows_map = [&](const float* base, int64_t rows, uint32_t box_rows) {
        uint64_t d[2] = {(uint64_t)C, (uint64_t)rows}, s[1] = {(uint64_t)C * 4};
        uint32_t b[2] = {32, box_rows};
        return get_tmap_f32(base, 2, d, s, b);
    };
    auto cols_map = [&](const float* base, int64_t rows) {  // transposed copy [C, rows]: box = 32 rows(inner) x 32 channels
        uint64_t d[2] = {(uint64_t)rows, (uint64_t)C}, s[1] = {(uint64_t)rows * 4};
        uint32_t b[2] = {32, 32};
        return get_tmap_f32(base, 2, d, s, b);
    };
    const CUtensorMap* mQ128 = rows_map(q, RL, BR);
    const CUtensorMap* mG128 = rows_map(dout, RL, BR);
    const CUtensorMap* mK64 = rows_map(k, RS, BT);
    const CUtensorMap* mV64 = rows_map(v, RS, BT);
    const CUtensorMap* mKt = cols_map(kt, RS);
    const CUtensorMap* mK128 = rows_map(k, RS, BR);
    const CUtensorMap* mV128 = rows_map(v, RS, BR);
    const CUtensorMap* mQ64 = rows_map(q, RL, BT);
    const CUtensorMap* mG64 = rows_map(dout, RL, BT);
    const CUtensorMap* mGt = cols_map(dot, RL);
    const CUtensorMap* mQt = cols_map(qt, RL);
    if (!mQ128 || !mG128 || !mK64 || !mV64 || !mKt || !mK128 || !mV128 || !mQ64 || !mG64 || !mGt || !mQt) return COFI_ECUDA;

    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e1 = cudaFuncSetAttribute(attention_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              BCfg<false>::SMEM);
        cudaError_t e2 = cudaFuncSetAttribute(attention_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              BCfg<true>::SMEM);
        if (e1 != cudaSuccess || e2 != cudaSuccess) {
            set_error("cudaFuncSetAttribute(attention backward): %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
            return COFI_ECUDA;
        }
        attr_done = true;
    }
    BwdParams pq{out, dout, lse, dsum_work, dq, nullptr, L, S, L, heads, scale, (int)ceil_div(S, BT)};
    dim3 g1((unsigned)ceil_div(L, BR), heads, frames);
    attention_bwd_tc_kernel<false><<<g1, 192, BCfg<false>::SMEM, st>>>(*mQ128, *mG128, *mK64, *mV64, *mKt, *mKt, pq);
    if ((rc = check_launch("cofi_attention_bwd_tc(dq)"))) return rc;
    BwdParams pk{out, dout, lse, dsum_work, dv, dk, S, L, L, heads, scale, (int)ceil_div(L, BT)};
    dim3 g2((unsigned)ceil_div(S, BR), heads, frames);
    attention_bwd_tc_kernel<true><<<g2, 192, BCfg<true>::SMEM, st>>>(*mK128, *mV128, *mQ64, *mG64, *mGt, *mQt, pk);
    return check_launch("cofi_attention_bwd_tc(dkv)");
}
