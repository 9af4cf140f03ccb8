// sim_tc.cu -- fused similarity + row arg-min on tcgen05 (throughput engine of cofi_sim_argmin).
// reference model/network.py:174-179: D = 1 - img^T pc ; argmin over the pixels for every selected point.
//
// The [Npt, Npx] similarity matrix is never written anywhere: a CTA owns 128 point rows (A operand, loaded once by
// TMA), streams the pixel features in tiles of 128 (B operand, k-block granular shared-memory ring), the tensor core
// writes each 128x128 fp32 score tile into one of two TMEM buffers, and four epilogue warps (one point row per
// thread) fold the tile into a running (best score, best index) pair straight out of TMEM while the next tile's MMAs
// run into the other buffer.  HBM/L2 traffic: (Npt + Npx * ceil(Npt/128)) * C * 4 bytes; output 12 bytes per point.
// Scores are tf32 products with fp32 accumulation: ranking of near-ties can differ from the fp32 engine, which stays
// the bit-exact path (match.cu).
#include "common.cuh"
#include "tc_common.cuh"

namespace cofi {
namespace tc {

constexpr int SM_M = 128, SM_N = 128, SM_K = 32;
constexpr int SM_SLOT = SM_N * SM_K * 4;  // 16 KB per k-block of B (and of A)
constexpr int SM_RING = 6;
constexpr int SM_MAXKB = 4;               // C <= 128
constexpr int SM_SMEM = SM_MAXKB * SM_SLOT + SM_RING * SM_SLOT + 1024 + 256;

struct SimParams {
    int64_t* best_idx;
    float* best_val;
    int64_t Npt, Npx;
    int kb;         // C / 32
    int num_tiles;  // ceil(Npx / 128)
};

__global__ void __launch_bounds__(192)
sim_argmin_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const SimParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + SM_MAXKB * SM_SLOT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + SM_RING * SM_SLOT);
    uint64_t* a_full = bars;
    uint64_t* full = bars + 1;               // [SM_RING]
    uint64_t* empty = full + SM_RING;        // [SM_RING]
    uint64_t* s_full = empty + SM_RING;      // [2]
    uint64_t* s_empty = s_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);

    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    const int frame = blockIdx.y;
    const int64_t m0 = (int64_t)blockIdx.x * SM_M;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        mbar_init(a_full, 1);
        for (int s = 0; s < SM_RING; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&s_full[b], 1);
            mbar_init(&s_empty[b], 4);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        {   // whole warp runs the loop, the elected lane issues (tc_common.cuh: elect_one)
            const bool leader = elect_one();
            if (leader) {
                mbar_expect_tx(a_full, p.kb * SM_SLOT);
                for (int kb = 0; kb < p.kb; ++kb)
                    tma_load_2d(&tmA, a_full, sA + kb * SM_SLOT, kb * SM_K, (int)(frame * p.Npt + m0));
            }
            __syncwarp();
            int it = 0;
            for (int j = 0; j < p.num_tiles; ++j)
                for (int kb = 0; kb < p.kb; ++kb, ++it) {
                    const int s = it % SM_RING;
                    const uint32_t ph = (uint32_t)(it / SM_RING) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    if (leader) {
                        mbar_expect_tx(&full[s], SM_SLOT);
                        tma_load_2d(&tmB, &full[s], sB + s * SM_SLOT, kb * SM_K, (int)(frame * p.Npx + (int64_t)j * SM_N));
                    }
                    __syncwarp();
                }
        }
    } else if (warp == 1) {
        {
            const bool leader = elect_one();
            constexpr uint32_t idesc = umma_idesc(2, SM_M, SM_N);
            mbar_wait(a_full, 0);
            const uint32_t a_addr = smem_u32(sA);
            int it = 0;
            for (int j = 0; j < p.num_tiles; ++j) {
                const int b = j & 1;
                mbar_wait(&s_empty[b], ((uint32_t)(j >> 1) & 1u) ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < p.kb; ++kb, ++it) {
                    const int s = it % SM_RING;
                    mbar_wait(&full[s], (uint32_t)(it / SM_RING) & 1u);
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(sB + s * SM_SLOT);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < SM_K / 8; ++k)
                            mma_tf32(tmem_base + b * SM_N, umma_desc_k128(a_addr + kb * SM_SLOT + k * 32),
                                     umma_desc_k128(b_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
                        tc_commit(&empty[s]);
                    }
                    __syncwarp();
                }
                if (leader) tc_commit(&s_full[b]);
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float best = -INFINITY;
        int bidx = 0;
        for (int j = 0; j < p.num_tiles; ++j) {
            const int b = j & 1;
            mbar_wait(&s_full[b], (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            const int base = j * SM_N;
            const int valid = (int)((p.Npx - base) < SM_N ? (p.Npx - base) : SM_N);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t raw[32];
                tmem_ld32(tmem_base + lane_off + b * SM_N + c * 32, raw);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float v = __uint_as_float(raw[i]);
                    if (c * 32 + i < valid && v > best) {  // strict: lowest pixel index wins ties
                        best = v;
                        bidx = base + c * 32 + i;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[b]);
        }
        if (m0 + r < p.Npt) {
            p.best_idx[(int64_t)frame * p.Npt + m0 + r] = bidx;
            p.best_val[(int64_t)frame * p.Npt + m0 + r] = 1.0f - best;
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace tc

bool sim_argmin_tc_supported(int64_t ldpt, int64_t ldpx, int64_t Npt, int64_t Npx, int C) {
    return (C % 32 == 0) && C <= 32 * tc::SM_MAXKB && (ldpt % 4 == 0) && (ldpx % 4 == 0) && Npt > 0 && Npx > 0;
}

int sim_argmin_tc_launch(const float* pt, int64_t ldpt, const float* px, int64_t ldpx, int64_t Npt, int64_t Npx, int C,
                         int frames, int64_t* best_idx, float* best_val, int engine, cudaStream_t st) {
    using namespace tc;
    (void)engine;
    uint64_t dA[2] = {(uint64_t)C, (uint64_t)(frames * Npt)}, sA[1] = {(uint64_t)ldpt * 4};
    uint32_t bA[2] = {SM_K, SM_M};
    uint64_t dB[2] = {(uint64_t)C, (uint64_t)(frames * Npx)}, sB[1] = {(uint64_t)ldpx * 4};
    uint32_t bB[2] = {SM_K, SM_N};
    const CUtensorMap* ta = get_tmap_f32(pt, 2, dA, sA, bA);
    const CUtensorMap* tb = get_tmap_f32(px, 2, dB, sB, bB);
    if (!ta || !tb) return COFI_ECUDA;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(sim_argmin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(sim smem=%d): %s", SM_SMEM, cudaGetErrorString(e));
            return COFI_ECUDA;
        }
        attr_done = true;
    }
    SimParams p{best_idx, best_val, Npt, Npx, C / SM_K, (int)ceil_div(Npx, SM_N)};
    dim3 grid((unsigned)ceil_div(Npt, SM_M), frames);
    sim_argmin_tc_kernel<<<grid, 192, SM_SMEM, st>>>(*ta, *tb, p);
    return check_launch("cofi_sim_argmin(tcgen05)");
}

}  // namespace cofi
