// sim_tc16.cu -- fused similarity + row arg-min, fp16 operands (kind::f16), 256-point CTAs.
// reference model/network.py:174-179 (D = 1 - img^T pc, argmin over pixels); north star: "single fused bf16 kernel
// that never materialises the full matrix".  Features are L2-normalised (|x| <= 1), so fp16 (11 significand bits)
// is the better 16-bit format here.
//
// Why this shape: with fp32/tf32 operands a 128x128 tile needs 32 KB of operands per 0.5 M MACs and the kernel is bound
// by L2->SM bandwidth (~6.3 KB/clk chip-wide) at ~35 % of the tensor peak.  Here a CTA keeps TWO 128-point operand
// tiles resident (A: 2 x C x 128 fp16), so every streamed 128-pixel B tile (16 KB per 64 channels) feeds two MMAs:
// operand traffic per MAC drops 4x vs the tf32 kernel, and kind::f16 runs at twice the tf32 rate.  TMEM is fully
// used: 2 point-halves x 2 buffers x 128 columns = 512.  Eight epilogue warps (one point row per thread) fold each
// finished 128x128 score tile into a running (best, index) pair with one FMNMX per score (chunk max first, index
// rescan only when the running best improves) while the tensor core fills the other buffer.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace cofi {
namespace tc {

constexpr int S16_MH = 128, S16_N = 128, S16_K = 64;       // fp16 elements per k-block (128 bytes)
constexpr int S16_SLOT = S16_N * 128;                      // 16 KB per k-block tile
constexpr int S16_RING = 6, S16_MAXKB = 2;                 // C <= 128
constexpr int S16_SMEM = 2 * S16_MAXKB * S16_SLOT + S16_RING * S16_SLOT + 1024 + 256;

struct Sim16Params {
    int64_t* best_idx;
    float* best_val;
    int64_t Npt, Npx;
    int kb;
    int num_tiles;       // pixel tiles in total
    int tiles_per_split; // pixel tiles handled by one CTA (blockIdx.z selects the range)
    int64_t split_stride; // rows of the output per split (frames * Npt) when partial results are written
};

__global__ void __launch_bounds__(320)
sim_argmin_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const Sim16Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                     // [half][kb] tiles
    uint8_t* sB = smem + 2 * S16_MAXKB * S16_SLOT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + S16_RING * S16_SLOT);
    uint64_t* a_full = bars;
    uint64_t* full = bars + 1;
    uint64_t* empty = full + S16_RING;
    uint64_t* s_full = empty + S16_RING;   // [2]
    uint64_t* s_empty = s_full + 2;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int frame = blockIdx.y;
    const int64_t m0 = (int64_t)blockIdx.x * (2 * S16_MH);
    const int t0 = blockIdx.z * p.tiles_per_split;
    const int t1 = (t0 + p.tiles_per_split < p.num_tiles) ? t0 + p.tiles_per_split : p.num_tiles;
    const int ntl = t1 - t0;  // tiles of this CTA

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        mbar_init(a_full, 1);
        for (int s = 0; s < S16_RING; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&s_full[b], 1);
            mbar_init(&s_empty[b], 8);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(a_full, 2 * p.kb * S16_SLOT);
            for (int h = 0; h < 2; ++h)
                for (int kb = 0; kb < p.kb; ++kb)
                    tma_load_2d(&tmA, a_full, sA + (h * S16_MAXKB + kb) * S16_SLOT, kb * S16_K,
                                (int)(frame * p.Npt + m0 + h * S16_MH));
            int it = 0;
            for (int j = 0; j < ntl; ++j)
                for (int kb = 0; kb < p.kb; ++kb, ++it) {
                    const int s = it % S16_RING;
                    mbar_wait(&empty[s], ((uint32_t)(it / S16_RING) & 1u) ^ 1u);
                    mbar_expect_tx(&full[s], S16_SLOT);
                    tma_load_2d(&tmB, &full[s], sB + s * S16_SLOT, kb * S16_K,
                                (int)(frame * p.Npx + (int64_t)(t0 + j) * S16_N));
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(0 /*f16*/, S16_MH, S16_N);
            mbar_wait(a_full, 0);
            const uint32_t a_addr = smem_u32(sA);
            int it = 0;
            for (int j = 0; j < ntl; ++j) {
                const int b = j & 1;
                mbar_wait(&s_empty[b], ((uint32_t)(j >> 1) & 1u) ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < p.kb; ++kb, ++it) {
                    const int s = it % S16_RING;
                    mbar_wait(&full[s], (uint32_t)(it / S16_RING) & 1u);
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(sB + s * S16_SLOT);
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int k = 0; k < 4; ++k)  // UMMA_K = 16 fp16 = 32 bytes per step
                            mma_f16(tmem_base + (uint32_t)((h * 2 + b) * S16_N),
                                    umma_desc_k128(a_addr + (h * S16_MAXKB + kb) * S16_SLOT + k * 32),
                                    umma_desc_k128(b_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(&empty[s]);
                }
                tc_commit(&s_full[b]);
            }
        }
    } else {
        const int q = warp & 3;            // TMEM lane quarter this warp may touch
        const int h = (warp - 2) >> 2;     // point half (warps 2..5 -> 0, 6..9 -> 1)
        const int r = h * S16_MH + q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float best = -INFINITY;
        int bidx = 0;
        for (int j = 0; j < ntl; ++j) {
            const int b = j & 1;
            mbar_wait(&s_full[b], (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            const int base = (t0 + j) * S16_N;
            const bool last = base + S16_N > p.Npx;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t raw[32];
                tmem_ld32(tmem_base + lane_off + (uint32_t)((h * 2 + b) * S16_N + c * 32), raw);
                tmem_ld_wait();
                if (last) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (base + c * 32 + i >= p.Npx) raw[i] = 0xff800000u;  // -inf
                }
                float m01[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) m01[i] = fmaxf(__uint_as_float(raw[2 * i]), __uint_as_float(raw[2 * i + 1]));
#pragma unroll
                for (int i = 0; i < 8; ++i) m01[i] = fmaxf(m01[i], m01[i + 8]);
                const float cm = fmaxf(fmaxf(fmaxf(m01[0], m01[1]), fmaxf(m01[2], m01[3])),
                                       fmaxf(fmaxf(m01[4], m01[5]), fmaxf(m01[6], m01[7])));
                if (cm > best) {  // rare once the running best has warmed up; strict > keeps the lowest index on ties
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float v = __uint_as_float(raw[i]);
                        if (v > best) {
                            best = v;
                            bidx = base + c * 32 + i;
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[b]);
        }
        if (m0 + r < p.Npt) {
            const int64_t o = (int64_t)blockIdx.z * p.split_stride + (int64_t)frame * p.Npt + m0 + r;
            p.best_idx[o] = bidx;
            p.best_val[o] = 1.0f - best;
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// merge of the per-split partial results: ascending split order + strict '<' keeps the lowest pixel index on ties
__global__ void __launch_bounds__(256)
sim_merge_kernel(const int64_t* __restrict__ pidx, const float* __restrict__ pval, int nsplit, int64_t rows,
                 int64_t* __restrict__ best_idx, float* __restrict__ best_val) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    float bv = pval[i];
    int64_t bi = pidx[i];
    for (int s = 1; s < nsplit; ++s) {
        const float v = pval[(int64_t)s * rows + i];
        if (v < bv) {
            bv = v;
            bi = pidx[(int64_t)s * rows + i];
        }
    }
    best_idx[i] = bi;
    best_val[i] = bv;
}

__global__ void __launch_bounds__(256)
cast_f16_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int C, __half* __restrict__ y, int64_t ldy) {
    const int64_t total = rows * (C >> 1);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / (C >> 1);
        const int c = (int)(t - row * (C >> 1)) * 2;
        const float2 v = *reinterpret_cast<const float2*>(x + row * ldx + c);
        *reinterpret_cast<__half2*>(y + row * ldy + c) = __floats2half2_rn(v.x, v.y);
    }
}

}  // namespace tc
}  // namespace cofi

using namespace cofi;

extern "C" int cofi_cast_f16(const float* x, int64_t ldx, int64_t rows, int C, void* y, int64_t ldy, void* stream) {
    COFI_REQUIRE(x && y && rows >= 0 && C > 0 && C % 2 == 0 && ldx % 2 == 0 && ldy % 2 == 0, "cofi_cast_f16: bad argument");
    if (rows == 0) return COFI_OK;
    const int64_t work = rows * (C / 2);
    int64_t blocks = ceil_div(work, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    tc::cast_f16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, rows, C, reinterpret_cast<__half*>(y), ldy);
    return check_launch("cofi_cast_f16");
}

extern "C" int cofi_sim_argmin_f16(const void* pt, int64_t ldpt, const void* px, int64_t ldpx, int64_t Npt, int64_t Npx,
                                   int C, int frames, int64_t* best_idx, float* best_val, int nsplit, int64_t* ws_idx,
                                   float* ws_val, void* stream) {
    using namespace cofi::tc;
    COFI_REQUIRE(pt && px && best_idx && best_val, "cofi_sim_argmin_f16: null pointer");
    COFI_REQUIRE(Npt > 0 && Npx > 0 && frames > 0, "cofi_sim_argmin_f16: bad shape");
    COFI_REQUIRE(C % 64 == 0 && C <= 64 * S16_MAXKB, "cofi_sim_argmin_f16: C=%d must be 64 or 128", C);
    COFI_REQUIRE(ldpt % 8 == 0 && ldpx % 8 == 0 && ((uintptr_t)pt % 16) == 0 && ((uintptr_t)px % 16) == 0,
                 "cofi_sim_argmin_f16: rows must be 16-byte aligned");
    uint64_t dA[2] = {(uint64_t)C, (uint64_t)(frames * Npt)}, sA[1] = {(uint64_t)ldpt * 2};
    uint32_t bA[2] = {S16_K, S16_MH};
    uint64_t dB[2] = {(uint64_t)C, (uint64_t)(frames * Npx)}, sB[1] = {(uint64_t)ldpx * 2};
    uint32_t bB[2] = {S16_K, S16_N};
    const CUtensorMap* ta = get_tmap_f16(pt, 2, dA, sA, bA);
    const CUtensorMap* tb = get_tmap_f16(px, 2, dB, sB, bB);
    if (!ta || !tb) return COFI_ECUDA;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(sim_argmin_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S16_SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(sim16 smem=%d): %s", S16_SMEM, cudaGetErrorString(e));
            return COFI_ECUDA;
        }
        attr_done = true;
    }
    const int num_tiles = (int)ceil_div(Npx, S16_N);
    if (nsplit < 1) nsplit = 1;
    if (nsplit > num_tiles) nsplit = num_tiles;
    COFI_REQUIRE(nsplit == 1 || (ws_idx && ws_val), "cofi_sim_argmin_f16: nsplit > 1 needs the partial-result workspace");
    const int tps = (int)ceil_div(num_tiles, nsplit);
    nsplit = (int)ceil_div(num_tiles, tps);
    const int64_t rows = (int64_t)frames * Npt;
    Sim16Params p{nsplit > 1 ? ws_idx : best_idx, nsplit > 1 ? ws_val : best_val, Npt, Npx, C / S16_K, num_tiles, tps, rows};
    dim3 grid((unsigned)ceil_div(Npt, 2 * S16_MH), frames, nsplit);
    sim_argmin_f16_kernel<<<grid, 320, S16_SMEM, (cudaStream_t)stream>>>(*ta, *tb, p);
    int rc = check_launch("cofi_sim_argmin_f16");
    if (rc || nsplit == 1) return rc;
    sim_merge_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, (cudaStream_t)stream>>>(ws_idx, ws_val, nsplit, rows, best_idx,
                                                                                  best_val);
    return check_launch("cofi_sim_argmin_f16(merge)");
}
