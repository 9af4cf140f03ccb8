// attention_tc.cu -- tcgen05 flash attention for the LoFTR encoder layer
// (reference model/transformer/linear_attention.py:69-77: softmax(Q K^T / sqrt(D)) V per head; D = 32).
//
// One CTA = 128 query rows of one (frame, head).  Keys are walked in tiles of AK = 64:
//   warp 0   TMA producer: Q tile once; per key tile the K tile [64 keys x 32] and the V^T tile
//            [32 d x 64 keys] (V^T comes straight out of the v_proj GEMM with swapped operands, so both MMA
//            operands are K-major and use the same SWIZZLE_128B descriptors as the GEMM engine); 3-stage ring.
//   warp 1   MMA issuer:  S[128x64] = Q K^T  (kind::tf32, 4 x K=8) into one of TWO S buffers in TMEM;
//            O_t[128x32] = P V  (8 x K=8), fresh accumulator per tile.  S(t+1) is issued before P V(t), so the next
//            score tile is computed while the softmax warps still work on tile t (software pipeline).
//   warps 2-5 softmax: one query row per thread. tcgen05.ld of the S row, online max / exp2 / sum in fp32
//            registers (log2 domain, scale folded), O_(t-1) is read back from TMEM and folded into the running output
//            in registers with the usual exp(m_old - m_new) correction (that read also proves P V(t-1) finished
//            with the P buffer), then P is written to shared memory in the swizzled K-major operand layout (tf32)
//            -- no TMEM stores and no rescaling of TMEM accumulators.
// The [L,S,heads] score tensor of the reference never exists; HBM traffic is Q, K, V once per CTA row.
#include "common.cuh"
#include "tc_common.cuh"

namespace cofi {
namespace tc {

constexpr int AQ = 128;   // queries per CTA
constexpr int AK = 64;    // keys per tile (64: 81 KB of smem -> 2 CTAs per SM hide the MMA->softmax->MMA latency chain)
constexpr int P_BYTES = AQ * AK * 4;        // AK/32 k-blocks of [128 rows x 32 keys]
constexpr int KC = AK / 32;                 // 32-key chunks per tile
constexpr int KV_STAGES = 3;                // S(t+1) reads K(t+1) while P V(t) still reads V(t): three stages keep one load ahead
constexpr int TMEM_COLS = 256;              // S double-buffered: [0,AK) and [AK,2AK);  O_t: [2AK, 2AK+AD), AD <= 64
template <int AD>
struct ACfg {                               // AD = head dimension (32: the reference's 128/4; 64: BASELINE config 4's 256/4)
    static constexpr int DB = AD / 32;                 // 32-float k-blocks of the head dimension
    static constexpr int Q_BYTES = AQ * AD * 4;        // DB k-blocks of [128 rows x 32]
    static constexpr int K_BYTES = AK * AD * 4;        // DB k-blocks of [AK keys x 32]
    static constexpr int VT_BYTES = AD * AK * 4;       // KC chunks of [AD d-rows x 32 keys]
    static constexpr int SMEM = Q_BYTES + KV_STAGES * (K_BYTES + VT_BYTES) + P_BYTES + 1024 + 256;
};

struct AttParams {
    float* out;
    int64_t L, S;
    int heads;
    float scale;
    int num_tiles;
    float* lse;  // optional [frames*L, heads] log-sum-exp of the scaled scores (training forward)
};

template <int AD>
__global__ void __launch_bounds__(192)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmVt, const AttParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int Q_BYTES = ACfg<AD>::Q_BYTES, K_BYTES = ACfg<AD>::K_BYTES, VT_BYTES = ACfg<AD>::VT_BYTES, DB = ACfg<AD>::DB;
    uint8_t* sQ = smem;
    uint8_t* sKV = smem + Q_BYTES;                                   // [stage][K | Vt]
    uint8_t* sP = sKV + KV_STAGES * (K_BYTES + VT_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + P_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;            // [KV_STAGES]
    uint64_t* kv_empty = bars + 4;           // [KV_STAGES]
    uint64_t* s_full = bars + 7;             // [2]  one per S buffer
    uint64_t* p_full = bars + 9;
    uint64_t* o_full = bars + 10;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    const int head = blockIdx.y, frame = blockIdx.z;
    const int64_t q0 = (int64_t)blockIdx.x * AQ;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmVt);
        mbar_init(q_full, 1);
        for (int s = 0; s < KV_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        mbar_init(p_full, 4);
        mbar_init(o_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 2 * AK;   // S buffer b at tmem_S + b * AK

    if (warp == 0) {
        {   // whole warp runs the loop, the elected lane issues (tc_common.cuh: elect_one)
            const bool leader = elect_one();
            if (leader) {
                mbar_expect_tx(q_full, Q_BYTES);
                for (int kb = 0; kb < DB; ++kb)
                    tma_load_2d(&tmQ, q_full, sQ + kb * (AQ * 128), head * AD + kb * 32, (int)(frame * p.L + q0));
            }
            __syncwarp();
            for (int t = 0; t < p.num_tiles; ++t) {
                const int s = t % KV_STAGES;
                const uint32_t ph = (uint32_t)(t / KV_STAGES) & 1u;
                mbar_wait(&kv_empty[s], ph ^ 1u);
                uint8_t* kd = sKV + s * (K_BYTES + VT_BYTES);
                const int key0 = (int)(frame * p.S + (int64_t)t * AK);
                if (leader) {
                    mbar_expect_tx(&kv_full[s], K_BYTES + VT_BYTES);
#pragma unroll
                    for (int kb = 0; kb < DB; ++kb)
                        tma_load_2d(&tmK, &kv_full[s], kd + kb * (AK * 128), head * AD + kb * 32, key0);
#pragma unroll
                    for (int c = 0; c < KC; ++c)  // V^T chunk c: rows head*AD..+AD-1, keys key0+32c..+31
                        tma_load_2d(&tmVt, &kv_full[s], kd + K_BYTES + c * (AD * 32 * 4), key0 + c * 32, head * AD);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        {
            // Software pipeline: S(t+1) = Q K(t+1)^T is issued BEFORE P V(t), into the other S buffer, so the tensor pipe
            // works on the next score tile while the softmax warps are still busy with tile t.
            const bool leader = elect_one();
            constexpr uint32_t idesc_s = umma_idesc(2, AQ, AK);   // 128 x AK
            constexpr uint32_t idesc_o = umma_idesc(2, AQ, AD);   // 128 x AD
            mbar_wait(q_full, 0);
            const uint32_t q_addr = smem_u32(sQ);
            const uint32_t p_addr = smem_u32(sP);
            auto issue_s = [&](int t) {
                const int s = t % KV_STAGES;
                mbar_wait(&kv_full[s], (uint32_t)(t / KV_STAGES) & 1u);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(sKV + s * (K_BYTES + VT_BYTES));
                const uint32_t d = tmem_S + (uint32_t)(t & 1) * AK;
                if (leader) {
#pragma unroll
                    for (int kb = 0; kb < DB; ++kb)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_tf32(d, umma_desc_k128(q_addr + kb * (AQ * 128) + k * 32),
                                     umma_desc_k128(k_addr + kb * (AK * 128) + k * 32), idesc_s, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(&s_full[t & 1]);
                }
                __syncwarp();
            };
            issue_s(0);
            for (int t = 0; t < p.num_tiles; ++t) {
                const int s = t % KV_STAGES;
                const uint32_t tp = (uint32_t)t & 1u;
                // S buffer (t+1)&1 held tile t-1, whose P was published one iteration ago: free
                if (t + 1 < p.num_tiles) issue_s(t + 1);
                // O_t = P V   (P(t) published also means O(t-1) was folded: the O columns are free)
                mbar_wait(p_full, tp);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(sKV + s * (K_BYTES + VT_BYTES)) + K_BYTES;
                if (leader) {
#pragma unroll
                    for (int c = 0; c < KC; ++c)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            mma_tf32(tmem_O, umma_desc_k128(p_addr + c * (AQ * 32 * 4) + k * 32),
                                     umma_desc_k128(v_addr + c * (AD * 32 * 4) + k * 32), idesc_o, (c | k) != 0 ? 1u : 0u);
                    tc_commit(o_full);
                    tc_commit(&kv_empty[s]);
                }
                __syncwarp();
            }
        }
    } else {
        // ================================ softmax / output ================================
        const int q = warp & 3;
        const int r = q * 32 + lane;                 // query row inside the tile == TMEM lane
        const bool row_ok = q0 + r < p.L;
        float o[AD];
#pragma unroll
        for (int d = 0; d < AD; ++d) o[d] = 0.0f;
        float mrun = -INFINITY, lrun = 0.0f;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        uint8_t* prow = sP + r * 128;                // row r of each [128 x 32] k-block (128-byte rows)
        const int sw = r & 7;                        // SWIZZLE_128B: 16-byte chunk index ^= (row & 7)
        // scores are kept in the log2 domain: x = S * scale * log2(e), p = exp2(x - m)  (one FMUL + one MUFU per score)
        const float scale2 = p.scale * 1.4426950408889634f;
        float corr_prev = 1.0f;
        auto fold_o = [&](int t, float corr) {   // o = o * corr + O_t, then hand the O columns (and the P buffer) back
            mbar_wait(o_full, (uint32_t)t & 1u);
            tc_fence_after();
#pragma unroll
            for (int db = 0; db < DB; ++db) {
                uint32_t raw[32];
                tmem_ld32(tmem_O + lane_off + db * 32, raw);
                tmem_ld_wait();
#pragma unroll
                for (int d = 0; d < 32; ++d) o[db * 32 + d] = fmaf(o[db * 32 + d], corr, __uint_as_float(raw[d]));
            }
            tc_fence_before();
        };
        for (int t = 0; t < p.num_tiles; ++t) {
            const int valid = (int)((p.S - (int64_t)t * AK) < AK ? (p.S - (int64_t)t * AK) : AK);
            mbar_wait(&s_full[t & 1], (uint32_t)(t >> 1) & 1u);
            tc_fence_after();
            uint32_t raw[KC][32];
#pragma unroll
            for (int c = 0; c < KC; ++c) tmem_ld32(tmem_S + (uint32_t)(t & 1) * AK + lane_off + c * 32, raw[c]);
            tmem_ld_wait();
            // max on the raw scores (the scale is positive), then p = exp2(raw * scale2 - m) as one FFMA + one MUFU per score;
            // only the last key tile of a sequence can be partial
            float tmax = -INFINITY;
            if (valid < AK) {
#pragma unroll
                for (int c = 0; c < KC; ++c)
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c * 32 + j >= valid) raw[c][j] = 0xff800000u;   // -inf
            }
#pragma unroll
            for (int c = 0; c < KC; ++c)
#pragma unroll
                for (int j = 0; j < 32; ++j) tmax = fmaxf(tmax, __uint_as_float(raw[c][j]));
            const float mnew = fmaxf(mrun, tmax * scale2);
            const float corr = exp2f(mrun - mnew);
            float psum = 0.0f;
#pragma unroll
            for (int c = 0; c < KC; ++c)
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float e = exp2f(fmaf(__uint_as_float(raw[c][j]), scale2, -mnew));
                    raw[c][j] = __float_as_uint(e);
                    psum += e;
                }
            // the P buffer is still the A operand of P V(t-1): fold O(t-1) (its completion) before overwriting it
            if (t > 0) fold_o(t - 1, corr_prev);
#pragma unroll
            for (int c = 0; c < KC; ++c) {
                uint8_t* blk = prow + c * (AQ * 32 * 4);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(blk + ((((j >> 2) ^ sw) & 7) << 4)) =
                        make_float4(__uint_as_float(raw[c][j]), __uint_as_float(raw[c][j + 1]), __uint_as_float(raw[c][j + 2]),
                                    __uint_as_float(raw[c][j + 3]));
            }
            lrun = lrun * corr + psum;
            mrun = mnew;
            corr_prev = corr;
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        fold_o(p.num_tiles - 1, corr_prev);
        if (row_ok) {
            const float inv = 1.0f / lrun;
            float* op = p.out + ((int64_t)frame * p.L + q0 + r) * ((int64_t)p.heads * AD) + head * AD;
#pragma unroll
            for (int d = 0; d < AD; d += 4)
                *reinterpret_cast<float4*>(op + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
            if (p.lse) p.lse[((int64_t)frame * p.L + q0 + r) * p.heads + head] = (mrun + log2f(lrun)) * 0.6931471805599453f;
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace tc

bool attention_tc_supported(int64_t L, int64_t S, int heads, int D) {
    uint64_t d[2] = {4, 4}, st[1] = {16};
    (void)d;
    (void)st;
    return (D == 32 || D == 64) && heads >= 1 && L >= 1 && S >= 1 && (S % 4 == 0);
}

// vt: V^T [heads*D, frames*S] row-major (keys contiguous): produced by the v_proj GEMM with swapped operands
int attention_tc_launch(const float* q, const float* k, const float* vt, int64_t L, int64_t S, int frames, int heads,
                        int D, float scale, float* out, float* lse, cudaStream_t st) {
    using namespace tc;
    const int64_t C = (int64_t)heads * D;
    uint64_t dq[2] = {(uint64_t)C, (uint64_t)(frames * L)}, sq[1] = {(uint64_t)C * 4};
    uint32_t bq[2] = {32, AQ};
    uint64_t dk[2] = {(uint64_t)C, (uint64_t)(frames * S)}, sk[1] = {(uint64_t)C * 4};
    uint32_t bk[2] = {32, AK};
    uint64_t dv[2] = {(uint64_t)(frames * S), (uint64_t)C}, sv[1] = {(uint64_t)(frames * S) * 4};
    uint32_t bv[2] = {32, (uint32_t)D};
    const CUtensorMap* tq = get_tmap_f32(q, 2, dq, sq, bq);
    const CUtensorMap* tk = get_tmap_f32(k, 2, dk, sk, bk);
    const CUtensorMap* tv = get_tmap_f32(vt, 2, dv, sv, bv);
    if (!tq || !tk || !tv) return COFI_ECUDA;
    AttParams p{out, L, S, heads, scale, (int)ceil_div(S, AK), lse};
    dim3 grid((unsigned)ceil_div(L, AQ), heads, frames);
    static bool attr_done[2] = {false, false};
    auto set_attr = [&](auto kern, int smem, int slot) -> int {
        if (attr_done[slot]) return COFI_OK;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(attention smem=%d): %s", smem, cudaGetErrorString(e));
            return COFI_ECUDA;
        }
        attr_done[slot] = true;
        return COFI_OK;
    };
    if (D == 32) {
        if (int rc = set_attr(attention_tc_kernel<32>, ACfg<32>::SMEM, 0)) return rc;
        attention_tc_kernel<32><<<grid, 192, ACfg<32>::SMEM, st>>>(*tq, *tk, *tv, p);
    } else {
        if (int rc = set_attr(attention_tc_kernel<64>, ACfg<64>::SMEM, 1)) return rc;
        attention_tc_kernel<64><<<grid, 192, ACfg<64>::SMEM, st>>>(*tq, *tk, *tv, p);
    }
    return check_launch("cofi_attention(tcgen05)");
}

}  // namespace cofi
