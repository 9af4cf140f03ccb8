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
// exact mode: every epilogue thread (= one point row) keeps the pixels whose fp16 score lies within `margin` of its running
// best -- the only pixels that can win the exact fp32 comparison -- in a shared-memory list of S16_CAP entries
constexpr int S16_CAP = 16;
constexpr int S16_ROWS = 2 * S16_MH;
constexpr int S16_SMEM_CAND = S16_SMEM + S16_CAP * S16_ROWS * 8;

struct Sim16Params {
    int64_t* best_idx;
    float* best_val;
    int64_t Npt, Npx;
    int kb;
    int num_tiles;       // pixel tiles in total
    int tiles_per_split; // pixel tiles handled by one CTA (blockIdx.z selects the range)
    int64_t split_stride; // rows of the output per split (frames * Npt) when partial results are written
    // exact mode (CAND): per (split, row) candidate lists for sim_rerank_kernel
    float margin;          // 2 * (bound of |fp16 score - exact fp32 score|) for unit-norm rows
    const float* bound2;   // optional device [2]: max |pt row|^2, max |px row|^2 (scales the margin); null = rows have norm <= 1
    int32_t* cand_idx;     // [nsplit, rows, S16_CAP]
    float* cand_val;       // [nsplit, rows, S16_CAP]
    int32_t* cand_cnt;     // [nsplit, rows]   (-1 = list overflowed: the re-rank scans the whole row exactly)
};

// margin = 2 x bound(|fp16-engine score - exact fp32 total|) + the resolution of d = fl(1 - total): for |total| < 0.5 the
// subtraction quantises to 2^-24 (2^-23 above 1), so pixels whose totals differ by less than that can TIE in d, and the
// exact engine then takes the lowest index -- they must all be candidates.
__device__ __forceinline__ float sim_margin(float base, const float* bound2) {
    float m = base;
    if (bound2) m = base * sqrtf(bound2[0] * bound2[1]);
    return m + 2.5e-7f;
}

// explicit shared-state-space accesses: the dynamic-smem base is re-aligned through integer arithmetic, after which the
// compiler only knows a GENERIC pointer and would emit ST.E / LD.E (64-bit generic addressing) in the hot candidate scan
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

template <bool CAND>
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
    // candidate lists (CAND only; past the 1 KB alignment slack): values [S16_CAP][S16_ROWS] then indices, as shared-space
    // byte addresses; slot k of row r lives at lv_s + k * SLOT_B + r * 4
    const uint32_t lv_s = smem_u32(smem + S16_SMEM - 1024);
    const uint32_t li_s = lv_s + S16_CAP * S16_ROWS * 4;
    constexpr uint32_t SLOT_B = S16_ROWS * 4;

    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
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
        {   // whole warp runs the loop, the elected lane issues (tc_common.cuh: elect_one)
            const bool leader = elect_one();
            if (leader) {
                mbar_expect_tx(a_full, 2 * p.kb * S16_SLOT);
                for (int h = 0; h < 2; ++h)
                    for (int kb = 0; kb < p.kb; ++kb)
                        tma_load_2d(&tmA, a_full, sA + (h * S16_MAXKB + kb) * S16_SLOT, kb * S16_K,
                                    (int)(frame * p.Npt + m0 + h * S16_MH));
            }
            __syncwarp();
            int it = 0;
            for (int j = 0; j < ntl; ++j)
                for (int kb = 0; kb < p.kb; ++kb, ++it) {
                    const int s = it % S16_RING;
                    mbar_wait(&empty[s], ((uint32_t)(it / S16_RING) & 1u) ^ 1u);
                    if (leader) {
                        mbar_expect_tx(&full[s], S16_SLOT);
                        tma_load_2d(&tmB, &full[s], sB + s * S16_SLOT, kb * S16_K,
                                    (int)(frame * p.Npx + (int64_t)(t0 + j) * S16_N));
                    }
                    __syncwarp();
                }
        }
    } else if (warp == 1) {
        {
            const bool leader = elect_one();
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
                    if (leader) {
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int k = 0; k < 4; ++k)  // UMMA_K = 16 fp16 = 32 bytes per step
                                mma_f16(tmem_base + (uint32_t)((h * 2 + b) * S16_N),
                                        umma_desc_k128(a_addr + (h * S16_MAXKB + kb) * S16_SLOT + k * 32),
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
        const int q = warp & 3;            // TMEM lane quarter this warp may touch
        const int h = (warp - 2) >> 2;     // point half (warps 2..5 -> 0, 6..9 -> 1)
        const int r = h * S16_MH + q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float best = -INFINITY;
        int bidx = 0;
        const float margin = CAND ? sim_margin(p.margin, p.bound2) : 0.0f;
        float thr = -INFINITY;   // best - margin
        int cnt = 0;
        bool ovf = false;
        for (int j = 0; j < ntl; ++j) {
            const int b = j & 1;
            mbar_wait(&s_full[b], (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            const int base = (t0 + j) * S16_N;
            const bool last = base + S16_N > p.Npx;
            // CAND: the candidate scan makes the chunk body large; unrolled four times it no longer fits the instruction
            // cache (ncu: stall_no_inst on every reconvergence point), so the chunk loop stays rolled in that variant
#pragma unroll(CAND ? 1 : 4)
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
                if (CAND) {
                    // A lane enters when its chunk holds a score within the margin of its running best (a handful of
                    // times per row).  Lanes diverge here, so the body is kept branch-light: the running best is updated
                    // from the chunk maximum first, then only the 4-score groups whose partial maximum (m01, already
                    // computed for cm) passes the threshold are visited, and a visited score is stored unconditionally
                    // at the list tail, which advances only when the score qualifies.
                    if (cm > thr) {
                        best = fmaxf(best, cm);
                        thr = best - margin;
                        if (cnt >= S16_CAP - 4) {  // make room: drop what fell out of the margin since it was appended
                            int k2 = 0;
                            for (int k = 0; k < cnt; ++k) {
                                const uint32_t ov = lds32(lv_s + k * SLOT_B + r * 4);
                                if (__uint_as_float(ov) > thr) {
                                    sts32(lv_s + k2 * SLOT_B + r * 4, ov);
                                    sts32(li_s + k2 * SLOT_B + r * 4, lds32(li_s + k * SLOT_B + r * 4));
                                    ++k2;
                                }
                            }
                            cnt = k2;
                        }
#pragma unroll
                        for (int g2 = 0; g2 < 8; ++g2) {
                            if (m01[g2] > thr) {  // m01[g2] = max of raw[2g2], raw[2g2+1], raw[2g2+16], raw[2g2+17]
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int i = 2 * g2 + (e & 1) + (e >> 1) * 16;
                                    const float v = __uint_as_float(raw[i]);
                                    const uint32_t off = (uint32_t)(cnt < S16_CAP ? cnt : S16_CAP - 1) * SLOT_B + r * 4;
                                    sts32(lv_s + off, raw[i]);
                                    sts32(li_s + off, (uint32_t)(base + c * 32 + i));
                                    cnt += (v > thr) ? 1 : 0;
                                }
                            }
                        }
                        if (cnt >= S16_CAP) {  // the tail slot may have been overwritten: the row is re-scanned exactly
                            ovf = true;
                            cnt = S16_CAP;
                        }
                    }
                } else if (cm > best) {  // rare once the running best has warmed up; strict > keeps the lowest index on ties
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
            p.best_val[o] = CAND ? best : 1.0f - best;   // CAND: the raw fp16-engine score (the re-rank thresholds on it)
            if (CAND) {
                int k2 = 0;
                for (int k = 0; k < cnt && !ovf; ++k) {
                    const float ov = __uint_as_float(lds32(lv_s + k * SLOT_B + r * 4));
                    if (ov > thr) {
                        p.cand_val[o * S16_CAP + k2] = ov;
                        p.cand_idx[o * S16_CAP + k2] = (int32_t)lds32(li_s + k * SLOT_B + r * 4);
                        ++k2;
                    }
                }
                p.cand_cnt[o] = ovf ? -1 : k2;
            }
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

// ---------------------------------------------------------------------------------------------- exact re-rank
// The arithmetic of sim_argmin_simt_kernel (match.cu), i.e. of the reference's `1 - torch.sum(img * pc, dim=0)`
// (model/network.py:174): rounded products, ATen's cascade sum over the channel axis (groups of 16, sequential inside and
// across groups), d = 1 - total.  One warp per point row; lanes take the row's candidates (or, after a list overflow / an
// empty list, every pixel) and the warp keeps the lexicographic minimum of (d, pixel index).
__device__ __forceinline__ float sim_exact_d(const float* __restrict__ spt, const float* __restrict__ r, int C) {
    // 64 channels (16 independent 16-byte loads) are fetched before the strictly ordered arithmetic starts: two L2 round
    // trips per 128-channel row instead of one per group of 16 (the re-rank is a latency chain, not a throughput problem)
    float tot = 0.0f;
    for (int c0 = 0; c0 < C; c0 += 64) {
        float4 a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = __ldg(reinterpret_cast<const float4*>(r + c0 + 4 * i));
#pragma unroll
        for (int gI = 0; gI < 4; ++gI) {
            float g = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = a[gI * 4 + i];
                const float* sp = spt + c0 + gI * 16 + i * 4;
                g = __fadd_rn(g, __fmul_rn(v.x, sp[0]));
                g = __fadd_rn(g, __fmul_rn(v.y, sp[1]));
                g = __fadd_rn(g, __fmul_rn(v.z, sp[2]));
                g = __fadd_rn(g, __fmul_rn(v.w, sp[3]));
            }
            tot = __fadd_rn(tot, g);
        }
    }
    return __fsub_rn(1.0f, tot);
}

constexpr int RR_WARPS = 8, RR_MAXC = 128;

__global__ void __launch_bounds__(RR_WARPS * 32)
sim_rerank_kernel(const float* __restrict__ pt, int64_t ldpt, const float* __restrict__ px, int64_t ldpx, int64_t Npt,
                  int64_t Npx, int C, int64_t rows, int nsplit, float margin_base, const float* __restrict__ bound2,
                  const float* __restrict__ part_best, const int32_t* __restrict__ cand_idx,
                  const float* __restrict__ cand_val, const int32_t* __restrict__ cand_cnt,
                  int64_t* __restrict__ best_idx, float* __restrict__ best_val, int32_t* __restrict__ stats) {
    __shared__ float spt_all[RR_WARPS][RR_MAXC];
    __shared__ int s_need[RR_WARPS];
    __shared__ float s_bd[RR_WARPS][RR_WARPS];
    __shared__ int s_bi[RR_WARPS][RR_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * RR_WARPS + warp;
    const bool live = row < rows;
    float* spt = spt_all[warp];
    const int64_t frame = live ? row / Npt : 0;
    const float* pxb = px + frame * Npx * ldpx;
    float bd = INFINITY;
    int bi = 0x7fffffff;
    int evaluated = 0;
    bool scan = false;
    if (live) {
        for (int c = lane; c < C; c += 32) spt[c] = __ldg(pt + row * ldpt + c);
        __syncwarp();
        float gmax = -INFINITY;
        bool ovf = false;
        for (int s = 0; s < nsplit; ++s) {
            gmax = fmaxf(gmax, part_best[(int64_t)s * rows + row]);
            ovf |= cand_cnt[(int64_t)s * rows + row] < 0;
        }
        const float thr = gmax - sim_margin(margin_base, bound2);
        if (!ovf) {
            for (int slot = lane; slot < nsplit * S16_CAP; slot += 32) {
                const int s = slot / S16_CAP, k = slot - s * S16_CAP;
                const int64_t o = (int64_t)s * rows + row;
                if (k < cand_cnt[o] && cand_val[o * S16_CAP + k] > thr) {
                    const int x = cand_idx[o * S16_CAP + k];
                    const float d = sim_exact_d(spt, pxb + (int64_t)x * ldpx, C);
                    ++evaluated;
                    if (d < bd || (d == bd && x < bi)) {
                        bd = d;
                        bi = x;
                    }
                }
            }
        }
        // nothing evaluated anywhere in the warp (non-finite scores) or an overflowed list: exact scan of the whole row
        scan = ovf || __ballot_sync(0xffffffffu, evaluated > 0) == 0u;
    }
    if (lane == 0) s_need[warp] = scan ? 1 : 0;
    __syncthreads();
    // full scans are rare (about one row in ten thousand) but long: the whole block shares each of them, warp w taking the
    // pixels w*32 + lane, +256, ...; the lexicographic minimum of (d, pixel) does not depend on how the pixels are dealt out
    for (int w = 0; w < RR_WARPS; ++w) {
        if (!s_need[w]) continue;   // block-uniform
        const int64_t srow = (int64_t)blockIdx.x * RR_WARPS + w;
        const float* sb = px + (srow / Npt) * Npx * ldpx;
        float d0 = INFINITY;
        int i0 = 0x7fffffff;
        for (int64_t x = warp * 32 + lane; x < Npx; x += RR_WARPS * 32) {
            const float d = sim_exact_d(spt_all[w], sb + x * ldpx, C);
            if (d < d0) {  // x ascending per lane: strict < keeps the lowest index
                d0 = d;
                i0 = (int)x;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, d0, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i0, o);
            if (ov < d0 || (ov == d0 && oi < i0)) {
                d0 = ov;
                i0 = oi;
            }
        }
        if (lane == 0) {
            s_bd[w][warp] = d0;
            s_bi[w][warp] = i0;
        }
    }
    __syncthreads();
    if (!live) return;
    if (scan) {   // combine the eight partial minima of this warp's row
        bd = lane < RR_WARPS ? s_bd[warp][lane] : INFINITY;
        bi = lane < RR_WARPS ? s_bi[warp][lane] : 0x7fffffff;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < bd || (ov == bd && oi < bi)) {
            bd = ov;
            bi = oi;
        }
    }
    if (stats) {
        const int ev = __reduce_add_sync(0xffffffffu, evaluated);
        if (lane == 0) {
            atomicAdd(stats + 0, scan ? 0 : ev);   // candidates re-ranked
            atomicAdd(stats + 1, scan ? 1 : 0);    // rows that needed the full exact scan
        }
    }
    if (lane == 0) {
        best_idx[row] = bi;
        best_val[row] = bd;
    }
}

// fp32 -> fp16 rows plus the running maximum of the squared row norms (atomicMax on the float's bit pattern; norms are
// non-negative so the unsigned order is the float order).  One warp per row.
__global__ void __launch_bounds__(128)
cast_f16_bound_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int C, __half* __restrict__ y, int64_t ldy,
                      float* __restrict__ bound2_slot) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float ss = 0.0f;
    for (int c = lane * 2; c < C; c += 64) {
        const float2 v = *reinterpret_cast<const float2*>(x + row * ldx + c);
        ss += v.x * v.x + v.y * v.y;
        *reinterpret_cast<__half2*>(y + row * ldy + c) = __floats2half2_rn(v.x, v.y);
    }
    ss = warp_sum(ss) * 1.0001f;
    if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(bound2_slot), __float_as_uint(ss));
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

static int sim16_attrs() {
    using namespace cofi::tc;
    static bool attr_done = false;
    if (attr_done) return COFI_OK;
    cudaError_t e = cudaFuncSetAttribute(sim_argmin_f16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S16_SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(sim_argmin_f16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S16_SMEM_CAND);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(sim16 smem=%d): %s", S16_SMEM_CAND, cudaGetErrorString(e));
        return COFI_ECUDA;
    }
    attr_done = true;
    return COFI_OK;
}

extern "C" int cofi_cast_f16_bound(const float* x, int64_t ldx, int64_t rows, int C, void* y, int64_t ldy, float* bound2_slot,
                                   void* stream) {
    COFI_REQUIRE(x && y && bound2_slot && rows >= 0 && C > 0 && C % 2 == 0 && ldx % 2 == 0 && ldy % 2 == 0,
                 "cofi_cast_f16_bound: bad argument");
    if (rows == 0) return COFI_OK;
    tc::cast_f16_bound_kernel<<<(unsigned)ceil_div(rows, 4), 128, 0, (cudaStream_t)stream>>>(x, ldx, rows, C,
                                                                                           reinterpret_cast<__half*>(y), ldy, bound2_slot);
    return check_launch("cofi_cast_f16_bound");
}

// pixel-range split of the tcgen05 pass: fill the machine when frames * Npt / 256 CTAs would not
static int sim_exact_nsplit(int64_t Npt, int64_t Npx, int frames) {
    using namespace cofi::tc;
    const int64_t ctas = ceil_div(Npt, 2 * S16_MH) * frames;
    const int64_t tiles = ceil_div(Npx, S16_N);
    int64_t ns = 1;
    if (ctas < 148) ns = (2 * 148 + ctas - 1) / ctas;
    if (ns > tiles) ns = tiles;
    if (ns < 1) ns = 1;
    const int64_t tps = ceil_div(tiles, ns);
    return (int)ceil_div(tiles, tps);
}

extern "C" int64_t cofi_sim_argmin_exact_workspace(int64_t Npt, int64_t Npx, int frames) {
    using namespace cofi::tc;
    if (Npt <= 0 || Npx <= 0 || frames <= 0) return 0;
    const int64_t rows = Npt * frames, ns = sim_exact_nsplit(Npt, Npx, frames);
    // per (split, row): best index (8) + best score (4) + count (4) + S16_CAP * (index 4 + score 4); + 2 stats counters
    return ns * rows * (16 + S16_CAP * 8) + 64;
}

extern "C" int cofi_sim_argmin_exact(const float* pt, int64_t ldpt, const float* px, int64_t ldpx, const void* pt_h,
                                     int64_t ldpth, const void* px_h, int64_t ldpxh, int64_t Npt, int64_t Npx, int C,
                                     int frames, const float* bound2, int64_t* best_idx, float* best_val, void* work,
                                     int32_t* stats, void* stream) {
    using namespace cofi::tc;
    COFI_REQUIRE(pt && px && pt_h && px_h && best_idx && best_val && work, "cofi_sim_argmin_exact: null pointer");
    COFI_REQUIRE(Npt > 0 && Npx > 0 && frames > 0, "cofi_sim_argmin_exact: bad shape");
    COFI_REQUIRE(C % 64 == 0 && C <= 64 * S16_MAXKB, "cofi_sim_argmin_exact: C=%d must be 64 or 128", C);
    COFI_REQUIRE(ldpth % 8 == 0 && ldpxh % 8 == 0 && ((uintptr_t)pt_h % 16) == 0 && ((uintptr_t)px_h % 16) == 0,
                 "cofi_sim_argmin_exact: fp16 rows must be 16-byte aligned");
    COFI_REQUIRE(ldpx % 4 == 0 && ((uintptr_t)px % 16) == 0 && ((uintptr_t)work % 16) == 0,
                 "cofi_sim_argmin_exact: px rows / workspace must be 16-byte aligned");
    COFI_REQUIRE(Npx < (1ll << 31) && frames * Npt < (1ll << 31), "cofi_sim_argmin_exact: shape too large");
    uint64_t dA[2] = {(uint64_t)C, (uint64_t)(frames * Npt)}, sA[1] = {(uint64_t)ldpth * 2};
    uint32_t bA[2] = {S16_K, S16_MH};
    uint64_t dB[2] = {(uint64_t)C, (uint64_t)(frames * Npx)}, sB[1] = {(uint64_t)ldpxh * 2};
    uint32_t bB[2] = {S16_K, S16_N};
    const CUtensorMap* ta = get_tmap_f16(pt_h, 2, dA, sA, bA);
    const CUtensorMap* tb = get_tmap_f16(px_h, 2, dB, sB, bB);
    if (!ta || !tb) return COFI_ECUDA;
    if (int rc = sim16_attrs()) return rc;
    const int num_tiles = (int)ceil_div(Npx, S16_N);
    const int nsplit = sim_exact_nsplit(Npt, Npx, frames);
    const int tps = (int)ceil_div(num_tiles, nsplit);
    const int64_t rows = (int64_t)frames * Npt;
    // workspace carve-up (all offsets 16-byte aligned: rows * nsplit * {8, 4, 4, 64, 64})
    uint8_t* w = reinterpret_cast<uint8_t*>(work);
    int64_t* part_idx = reinterpret_cast<int64_t*>(w);
    w += (int64_t)nsplit * rows * 8;
    int32_t* c_idx = reinterpret_cast<int32_t*>(w);
    w += (int64_t)nsplit * rows * S16_CAP * 4;
    float* c_val = reinterpret_cast<float*>(w);
    w += (int64_t)nsplit * rows * S16_CAP * 4;
    float* part_val = reinterpret_cast<float*>(w);
    w += (int64_t)nsplit * rows * 4;
    int32_t* c_cnt = reinterpret_cast<int32_t*>(w);
    const float margin = 2.5e-3f;  // 2 x (2^-10 operand rounding + accumulation slack), see DESIGN.md
    Sim16Params p{part_idx, part_val, Npt, Npx, C / S16_K, num_tiles, tps, rows, margin, bound2, c_idx, c_val, c_cnt};
    dim3 grid((unsigned)ceil_div(Npt, 2 * S16_MH), frames, nsplit);
    sim_argmin_f16_kernel<true><<<grid, 320, S16_SMEM_CAND, (cudaStream_t)stream>>>(*ta, *tb, p);
    int rc = check_launch("cofi_sim_argmin_exact(tcgen05 pass)");
    if (rc) return rc;
    sim_rerank_kernel<<<(unsigned)ceil_div(rows, RR_WARPS), RR_WARPS * 32, 0, (cudaStream_t)stream>>>(
        pt, ldpt, px, ldpx, Npt, Npx, C, rows, nsplit, margin, bound2, part_val, c_idx, c_val, c_cnt, best_idx, best_val, stats);
    return check_launch("cofi_sim_argmin_exact(re-rank)");
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
    if (int rc = sim16_attrs()) return rc;
    const int num_tiles = (int)ceil_div(Npx, S16_N);
    if (nsplit < 1) nsplit = 1;
    if (nsplit > num_tiles) nsplit = num_tiles;
    COFI_REQUIRE(nsplit == 1 || (ws_idx && ws_val), "cofi_sim_argmin_f16: nsplit > 1 needs the partial-result workspace");
    const int tps = (int)ceil_div(num_tiles, nsplit);
    nsplit = (int)ceil_div(num_tiles, tps);
    const int64_t rows = (int64_t)frames * Npt;
    Sim16Params p{nsplit > 1 ? ws_idx : best_idx, nsplit > 1 ? ws_val : best_val, Npt, Npx, C / S16_K, num_tiles, tps, rows,
                  0.0f, nullptr, nullptr, nullptr, nullptr};
    dim3 grid((unsigned)ceil_div(Npt, 2 * S16_MH), frames, nsplit);
    sim_argmin_f16_kernel<false><<<grid, 320, S16_SMEM, (cudaStream_t)stream>>>(*ta, *tb, p);
    int rc = check_launch("cofi_sim_argmin_f16");
    if (rc || nsplit == 1) return rc;
    sim_merge_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, (cudaStream_t)stream>>>(ws_idx, ws_val, nsplit, rows, best_idx,
                                                                                  best_val);
    return check_launch("cofi_sim_argmin_f16(merge)");
}
