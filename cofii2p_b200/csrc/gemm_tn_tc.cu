// gemm_tn_tc.cu -- tcgen05 engine for the weight-gradient contractions of the training path:
//
//   C[Mo,No] = sum_r A[r,Mo] * B[r,No]        (dW of a Linear / KPConv: A = dY or the aggregated features, B = X or dY)
//   dw[Cout,(kh,kw,ci)] = sum_p dy[p,Cout] * x[p shifted by the tap, ci]   (dW of an NHWC convolution)
//
// The reduction index r (samples / output pixels) is the UMMA K dimension, and both operands are stored with r as the
// OUTER index -- they are "MN-major" operands (UMMA instruction-descriptor bits 15/16), so no transpose pass exists:
// TMA drops [32 r][32 columns] fp32 boxes (128-byte rows) straight into the one shared-memory layout tcgen05 accepts for
// MN-major tf32 operands, SWIZZLE_128B with a 32-byte atom (UMMA layout type 1 = TMA SWIZZLE_128B_ATOM_32B; address bits
// [5,6] ^= [7,8], i.e. the pattern repeats every 4 rows): consecutive 32-column blocks LBO = 4096 bytes apart, 4-row
// groups SBO = 512 bytes apart; one kind::tf32 MMA (K = 8) consumes two 4-row groups.
// Convolutions: the shifted (and, for stride 2, strided -- TMA elementStrides) input window of a tap is a 4-D box; TMA's
// out-of-bounds zero fill is the padding.  Split over r across CTAs (grid.z); partial tiles go to a workspace that a
// deterministic reduction kernel folds (no atomics).
#include "common.cuh"
#include "tc_common.cuh"

namespace cofi {
namespace tc {

const CUtensorMap* get_tmap_f32_mn(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                   const uint32_t* box, const uint32_t* elem_strides);

constexpr int TN_KB = 32;              // reduction rows per pipeline stage
constexpr int TN_BLK = TN_KB * 128;    // bytes of one [32 r][32 fp32] box

__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;   // next 32-column block
    d |= (uint64_t)(512 >> 4) << 32;         // next 4-row group
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                  // SWIZZLE_128B_BASE32B
    return d;
}

struct TnParams {
    float* out;      // [splits][Mo][No]
    int Mo, No;
    int64_t R, per;  // reduction length, rows per split (multiple of TN_KB)
    int conv, Ho, wchunks, KW, Cin, stride, pad;
};

template <int BN>
struct TnCfg {
    static constexpr int A_BYTES = 4 * TN_BLK;
    static constexpr int B_BYTES = (BN / 32) * TN_BLK;
    static constexpr int STAGE = A_BYTES + B_BYTES;
    static constexpr int NS = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int SMEM = STAGE * NS + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(192)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TnParams p) {
    using C = TnCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGE * C::NS);
    uint64_t* empty = full + C::NS;
    uint64_t* tmem_full = empty + C::NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128, n0 = blockIdx.y * BN;
    const int64_t r_begin = (int64_t)blockIdx.z * p.per;
    const int64_t r_end = r_begin + p.per < p.R ? r_begin + p.per : p.R;
    const int num_kb = (int)((r_end - r_begin + TN_KB - 1) / TN_KB);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < C::NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        {   // whole warp runs the loop, the elected lane issues (tc_common.cuh: elect_one)
            const bool leader = elect_one();
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % C::NS;
                mbar_wait(&empty[s], (((uint32_t)(kb / C::NS)) & 1u) ^ 1u);
                uint8_t* a_dst = smem + s * C::STAGE;
                uint8_t* b_dst = a_dst + C::A_BYTES;
                const int64_t r = r_begin + (int64_t)kb * TN_KB;
                if (leader) {
                    mbar_expect_tx(&full[s], C::STAGE);
#pragma unroll
                    for (int mb = 0; mb < 4; ++mb) tma_load_2d(&tmA, &full[s], a_dst + mb * TN_BLK, m0 + 32 * mb, (int)r);
                }
                if (p.conv) {
                    const int64_t g = r / TN_KB;
                    const int wc = (int)(g % p.wchunks);
                    const int ho = (int)((g / p.wchunks) % p.Ho);
                    const int b = (int)(g / ((int64_t)p.wchunks * p.Ho));
                    if (leader) {
#pragma unroll
                        for (int nb = 0; nb < BN / 32; ++nb) {
                            const int n = n0 + 32 * nb;
                            const int tap = n / p.Cin, ci0 = n - tap * p.Cin;
                            const int kh = tap / p.KW, kw = tap - kh * p.KW;
                            tma_load_4d(&tmB, &full[s], b_dst + nb * TN_BLK, ci0, wc * TN_KB * p.stride - p.pad + kw,
                                        ho * p.stride - p.pad + kh, b);
                        }
                    }
                } else if (leader) {
#pragma unroll
                    for (int nb = 0; nb < BN / 32; ++nb)
                        tma_load_2d(&tmB, &full[s], b_dst + nb * TN_BLK, n0 + 32 * nb, (int)r);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        {
            const bool leader = elect_one();
            constexpr uint32_t idesc = umma_idesc(2 /*tf32*/, 128, BN) | (1u << 15) | (1u << 16);  // A and B MN-major
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % C::NS;
                mbar_wait(&full[s], ((uint32_t)(kb / C::NS)) & 1u);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * C::STAGE);
                const uint32_t b_addr = a_addr + C::A_BYTES;
                if (leader) {
#pragma unroll
                    for (int k = 0; k < TN_KB / 8; ++k)
                        mma_tf32(tmem_base, umma_desc_mn128(a_addr + k * 1024, TN_BLK), umma_desc_mn128(b_addr + k * 1024, TN_BLK),
                                 idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(&empty[s]);
                }
                __syncwarp();
            }
            if (leader) tc_commit(tmem_full);
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;
        if (num_kb > 0) mbar_wait(tmem_full, 0);
        tc_fence_after();
        float* orow = p.out + ((int64_t)blockIdx.z * p.Mo + m) * p.No;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t acc[32];
            if (num_kb > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = 0u;
            }
            const int nb = n0 + c0;
            if (m < p.Mo && nb < p.No) {
                if (nb + 32 <= p.No) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(orow + nb + j) = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]),
                                                                                __uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (nb + j < p.No) orow[nb + j] = __uint_as_float(acc[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, BN);
}

static int pick_bn(int No) { return No <= 64 ? 64 : (No <= 128 ? 128 : 256); }

// rows per split (multiple of TN_KB) such that about two waves of CTAs exist and every CTA gets >= 8 k-blocks
int64_t tn_tc_per(int64_t R, int Mo, int No, int* splits) {
    const int bn = pick_bn(No);
    const int64_t tiles = (int64_t)((Mo + 127) / 128) * ((No + bn - 1) / bn);
    int64_t s = (2 * 148 + tiles - 1) / tiles;
    const int64_t max_s = (R + TN_KB * 8 - 1) / (TN_KB * 8);
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    int64_t per = ((R + s - 1) / s + TN_KB - 1) / TN_KB * TN_KB;
    *splits = (int)((R + per - 1) / per);
    return per;
}

template <int BN>
static int launch_tn(const CUtensorMap* ta, const CUtensorMap* tb, const TnParams& p, int splits, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(gemm_tn_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TnCfg<BN>::SMEM) != cudaSuccess)
            return set_error("gemm_tn_tc: cannot raise dynamic shared memory"), COFI_ECUDA;
        attr = true;
    }
    dim3 grid((unsigned)((p.Mo + 127) / 128), (unsigned)((p.No + BN - 1) / BN), (unsigned)splits);
    gemm_tn_tc_kernel<BN><<<grid, 192, TnCfg<BN>::SMEM, st>>>(*ta, *tb, p);
    return check_launch("gemm_tn_tc");
}

static int dispatch_tn(const CUtensorMap* ta, const CUtensorMap* tb, const TnParams& p, int splits, cudaStream_t st) {
    switch (pick_bn(p.No)) {
        case 64: return launch_tn<64>(ta, tb, p, splits, st);
        case 128: return launch_tn<128>(ta, tb, p, splits, st);
        default: return launch_tn<256>(ta, tb, p, splits, st);
    }
}

// partial[splits][Mo][No] = per-split A^T B; returns the number of splits through *splits_out
int gemm_tn_tc(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t R, int Mo, int No, float* partial,
               int* splits_out, cudaStream_t st) {
    TnParams p{};
    p.out = partial; p.Mo = Mo; p.No = No; p.R = R;
    p.per = tn_tc_per(R, Mo, No, splits_out);
    const uint64_t da[2] = {(uint64_t)Mo, (uint64_t)R}, sa[1] = {(uint64_t)lda * 4};
    const uint64_t db[2] = {(uint64_t)No, (uint64_t)R}, sb[1] = {(uint64_t)ldb * 4};
    const uint32_t box[2] = {32, TN_KB};
    const CUtensorMap* ta = get_tmap_f32_mn(A, 2, da, sa, box, nullptr);
    const CUtensorMap* tb = get_tmap_f32_mn(B, 2, db, sb, box, nullptr);
    if (!ta || !tb) return COFI_ECUDA;
    return dispatch_tn(ta, tb, p, *splits_out, st);
}

bool conv_wgrad_tc_ok(int Cin, int Cout, int Wo, int stride) {
    return Cin % 32 == 0 && Cout % 4 == 0 && Wo % TN_KB == 0 && TN_KB * stride <= 256;
}

int conv_wgrad_tc(const float* x, int B, int H, int W, int Cin, const float* dy, int Cout, int KH, int KW, int stride, int pad,
                  int Ho, int Wo, float* partial, int* splits_out, cudaStream_t st) {
    TnParams p{};
    p.out = partial; p.Mo = Cout; p.No = KH * KW * Cin; p.R = (int64_t)B * Ho * Wo;
    p.per = tn_tc_per(p.R, p.Mo, p.No, splits_out);
    p.conv = 1; p.Ho = Ho; p.wchunks = Wo / TN_KB; p.KW = KW; p.Cin = Cin; p.stride = stride; p.pad = pad;
    const uint64_t da[2] = {(uint64_t)Cout, (uint64_t)p.R}, sa[1] = {(uint64_t)Cout * 4};
    const uint32_t boxa[2] = {32, TN_KB};
    const uint64_t db[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t sb[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    const uint32_t boxb[4] = {32, (uint32_t)(TN_KB * stride), 1, 1};
    const uint32_t es[4] = {1, (uint32_t)stride, 1, 1};
    const CUtensorMap* ta = get_tmap_f32_mn(dy, 2, da, sa, boxa, nullptr);
    const CUtensorMap* tb = get_tmap_f32_mn(x, 4, db, sb, boxb, es);
    if (!ta || !tb) return COFI_ECUDA;
    return dispatch_tn(ta, tb, p, *splits_out, st);
}

}  // namespace tc
}  // namespace cofi
