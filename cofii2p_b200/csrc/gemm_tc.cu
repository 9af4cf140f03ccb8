// gemm_tc.cu -- tcgen05 tensor-core contraction engine (COFI_GEMM_TF32 / COFI_GEMM_TF32X3) for every dense
// contraction of the hot path: nn.Linear layers, KPConv weight-apply, stride-1 NHWC convolutions.
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T ),   A, W fp32 in HBM, both K-major.
//
// Structure (one 128 x BN output tile per CTA, warp-specialised):
//   warp 0     TMA producer: cp.async.bulk.tensor loads of a 128x32 fp32 A tile and a BNx32 fp32 W tile per
//              k-block into a 3-4 stage shared-memory ring (SWIZZLE_128B, 128-byte rows), mbarrier complete_tx.
//              Convolutions use a 4-D tensor map over the NHWC activation: the A tile of a (kh,kw,ci-chunk)
//              k-block is the box {32 ch, 64 w, 2 h, 1 b} at the shifted coordinate; TMA's out-of-bounds zero
//              fill implements the padding, so im2col never exists.
//   warp 1     MMA issuer: one thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8) four times
//              per k-block on UMMA shared-memory descriptors; fp32 accumulator lives in TMEM (BN columns);
//              tcgen05.commit releases ring slots and finally signals the epilogue.
//   warps 2-9  epilogue (two warps per TMEM lane quarter, alternating over the 32-column chunks): tcgen05.ld 32x32b.x32 (one accumulator row per thread), fused rowdiv / per-channel
//              affine / bias / residual / accumulate / activation, 128-byte vector stores.
//   warps 10-15 (TF32X3 only) operand splitters: rewrite each landed tile in place as hi = tf32-truncated value and
//              write lo = x - hi into a second buffer; the issuer then runs hi*hi + lo*hi + hi*lo (3xTF32), which
//              restores fp32-grade accuracy on the tensor cores.  The split is element-wise, hence swizzle-agnostic.
#include <stdlib.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace cofi {

struct Epilogue {  // must match gemm_simt.cu
    const float* bias;
    const float* rowdiv;
    const float* colscale;
    const float* colshift;
    const float* residual;
    int64_t ldres;
    int accumulate;
    int act;
    const float* ln_gamma;  // fused row LayerNorm over the N columns (tcgen05 engine, N <= 128): y = act(LN(x))+residual
    const float* ln_beta;
    float ln_eps;
};

namespace tc {

// ------------------------------------------------------------------------------------- TMA descriptor cache
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

struct Key {
    uint64_t v[12];
    bool operator==(const Key& o) const {
        for (int i = 0; i < 12; ++i)
            if (v[i] != o.v[i]) return false;
        return true;
    }
};
struct KeyHash {
    size_t operator()(const Key& k) const {
        uint64_t h = 1469598103934665603ull;
        for (int i = 0; i < 12; ++i) h = (h ^ k.v[i]) * 1099511628211ull;
        return (size_t)h;
    }
};

static const CUtensorMap* get_tmap_any(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                       const uint32_t* box, int half, const uint32_t* elem_strides = nullptr,
                                       int atom32 = 0 /* SWIZZLE_128B with a 32-byte atom (MN-major tf32 operands) */) {
    static std::mutex mu;
    static std::unordered_map<Key, CUtensorMap*, KeyHash> cache;
    Key k{};
    k.v[0] = (uint64_t)(uintptr_t)base;
    k.v[1] = (uint64_t)rank | ((uint64_t)half << 8) | ((uint64_t)atom32 << 16);
    for (int i = 0; i < rank; ++i) {
        k.v[2 + i] = dims[i];
        k.v[6 + i] = ((uint64_t)box[i] << 40) | (i > 0 ? strides_bytes[i - 1] : 0);
        if (elem_strides) k.v[10] |= (uint64_t)elem_strides[i] << (8 * i);
    }
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(k);
    if (it != cache.end()) return it->second;
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return nullptr;
    }
    if (cache.size() > 8192) {  // pointers churn only outside CUDA graphs; bound the cache
        // descriptors handed out a moment ago (the A/W maps of the launch being assembled) must stay valid: retired maps
        // are freed one generation later
        static std::vector<CUtensorMap*> retired;
        for (CUtensorMap* m : retired) delete m;
        retired.clear();
        for (auto& kv : cache) retired.push_back(kv.second);
        cache.clear();
    }
    CUtensorMap* m = new CUtensorMap;
    cuuint64_t gdim[5];
    cuuint64_t gstr[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = elem_strides ? elem_strides[i] : 1;
        if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
    }
    CUresult r = fn(m, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        delete m;
        set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu box %u %u", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
        return nullptr;
    }
    cache.emplace(k, m);
    return m;
}

const CUtensorMap* get_tmap_f32(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                const uint32_t* box) {
    return get_tmap_any(base, rank, dims, strides_bytes, box, 0);
}
const CUtensorMap* get_tmap_f32_mn(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                   const uint32_t* box, const uint32_t* elem_strides) {
    return get_tmap_any(base, rank, dims, strides_bytes, box, 0, elem_strides, 1);
}
const CUtensorMap* get_tmap_f16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                const uint32_t* box) {
    return get_tmap_any(base, rank, dims, strides_bytes, box, 1);
}
// fp32, K-major, with element strides (strided convolutions); used by gemm_x3.cu
const CUtensorMap* get_tmap_f32_es(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                   const uint32_t* box, const uint32_t* elem_strides) {
    return get_tmap_any(base, rank, dims, strides_bytes, box, 0, elem_strides, 0);
}

// ---------------------------------------------------------------------------------------------- kernel
constexpr int TM = 128;   // tile rows (UMMA M)
constexpr int TK = 32;    // fp32 elements per k-block = 128 bytes = one swizzle row
constexpr int A_BYTES = TM * TK * 4;
constexpr int CONV_TW = 64, CONV_TH = 2;  // conv A tile = 2 image rows x 64 pixels

struct Params {
    float* C;
    int64_t ldc;
    int64_t M;
    int N;
    int num_kb;
    Epilogue ep;
    // convolution geometry (CONV only)
    int Ho, Wo, Cin, cpt /* 32-channel chunks per tap */, KW, pad, tiles_w, tiles_per_img, cstride /* conv stride (1 or 2) */;
    int tma_store;    // dense output with 16-byte aligned rows: tiles leave through TMA stores (tmC), else per-thread stores
    float* stat_out;  // optional [row tiles][N][2] per-tile column (sum, sum of squares) of the stored output
    int dbg;  // COFI_TC_DEBUG bits (perf triage only): 1 = skip global stores, 2 = skip TMA+MMA
};

__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t b = __float_as_uint(x);
    b += 0x00000FFFu + ((b >> 13) & 1u);
    return __uint_as_float(b & 0xFFFFE000u);
}

// LIGHT (3xTF32 only): the configuration for short contractions (K <= 256).  Those launches are one load -> split -> MMA ->
// epilogue chain per CTA with nothing to overlap inside the CTA, so what matters is how many CTAs an SM holds: a 2-stage
// ring (hi + lo copies: <= 96 KB for BN <= 64), one epilogue set and four splitter warps = 320 threads at <= 102 registers,
// two CTAs per SM -- the same residency as the plain tf32 kernel.  The deep configuration (3 stages, two epilogue sets, six
// splitter warps, one CTA per SM) serves the long, tensor-bound contractions.
template <int BN, bool X3, bool LIGHT = false>
struct Cfg {
    static constexpr int B_BYTES = BN * TK * 4;
    static constexpr int STAGE = A_BYTES + B_BYTES;
    static constexpr int NS = LIGHT ? 2 : ((BN == 128) ? 3 : 4);
    static constexpr int RING = STAGE * NS * (X3 ? 2 : 1);
    static constexpr int SMEM = RING + 1024 /* alignment slack */ + 256 /* barriers */ + 12 * BN * 4 /* epilogue vectors + column statistics */;
    static constexpr int EPI_SETS = LIGHT ? 1 : 2;  // epilogue warp sets (4 warps each, one per TMEM lane quarter); the sets
                                                    // interleave over the 32-column chunks of the tile
    static constexpr int SPLIT_WARPS = X3 ? (LIGHT ? 4 : 6) : 0;         // operand splitters (3xTF32 only)
    static constexpr int THREADS = 32 * (2 + 4 * EPI_SETS + SPLIT_WARPS);  // TMA warp, MMA warp, epilogue sets, splitters
    static constexpr int MIN_CTAS = (X3 && !LIGHT) ? 1 : 2;  // register budget: two resident CTAs except for the deep 3xTF32 config
    static_assert(!LIGHT || (X3 && BN <= 64), "LIGHT is the short-K 3xTF32 configuration");
};

// LEAN: epilogue variant without fused LayerNorm, residual, accumulate and sigmoid (row divisor, bias / per-channel affine,
// relu / leaky-relu and the column statistics stay): the generic epilogue carries every option as a not-taken branch, about
// 2000 SASS instructions per 32-column chunk, and the instruction fetch of that sparse walk is what the short contractions
// paid for (csrc/gemm_x3.cu, PLAIN: 30.9 -> 21.8 us on 163840 x 128 x 32).
template <int BN, bool CONV, bool X3, bool HALF = false, bool LIGHT = false, bool LEAN = false>
__global__ void __launch_bounds__(Cfg<BN, X3, LIGHT>::THREADS, Cfg<BN, X3, LIGHT>::MIN_CTAS)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const Params p) {
    using C = Cfg<BN, X3, LIGHT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* lo_base = smem + C::STAGE * C::NS;  // X3 only
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::RING);
    uint64_t* full = bars;                 // [NS]
    uint64_t* empty = bars + C::NS;        // [NS]
    uint64_t* ready = bars + 2 * C::NS;    // [NS] (X3)
    uint64_t* tmem_full = bars + 3 * C::NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * C::NS + 1);
    float* s_scale = reinterpret_cast<float*>(smem + C::RING + 256);  // [BN] per-column multiplier
    float* s_shift = s_scale + BN;                                   // [BN] per-column shift (+bias)
    float* s_gamma = s_shift + BN;                                   // [BN] fused-LayerNorm weight
    float* s_beta = s_gamma + BN;                                    // [BN] fused-LayerNorm bias
    float* s_col = s_beta + BN;                                      // [4 quarters][BN][2] column statistics

    const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * BN;
    // tile origin
    int64_t m0 = 0;
    int cb = 0, ch0 = 0, cw0 = 0;
    if (CONV) {
        cb = blockIdx.x / p.tiles_per_img;
        const int t = blockIdx.x - cb * p.tiles_per_img;
        ch0 = (t / p.tiles_w) * CONV_TH;
        cw0 = (t % p.tiles_w) * CONV_TW;
    } else {
        m0 = (int64_t)blockIdx.x * TM;
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < C::NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&ready[s], C::SPLIT_WARPS > 0 ? C::SPLIT_WARPS : 1);
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
        // ================================ TMA producer ================================
        if (!(p.dbg & 2)) {
            const bool leader = elect_one();  // whole warp runs the loop, the elected lane issues (tc_common.cuh)
            for (int kb = 0; kb < p.num_kb; ++kb) {
                const int s = kb % C::NS;
                const uint32_t ph = (uint32_t)(kb / C::NS) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                uint8_t* a_dst = smem + s * C::STAGE;
                uint8_t* b_dst = a_dst + A_BYTES;
                if (CONV) {
                    const int tap = kb / p.cpt, cc = kb - tap * p.cpt;
                    const int kh = tap / p.KW, kw = tap - kh * p.KW;
                    if (leader) {
                        mbar_expect_tx(&full[s], C::STAGE);
                        // strided conv: the tensor map carries elementStrides {1, s, s, 1}, coordinates are in input pixels
                        tma_load_4d(&tmA, &full[s], a_dst, cc * TK, cw0 * p.cstride + kw - p.pad, ch0 * p.cstride + kh - p.pad, cb);
                        tma_load_2d(&tmB, &full[s], b_dst, tap * p.Cin + cc * TK, n0);
                    }
                } else if (leader) {
                    // k-block = 128 bytes of K: 32 fp32 or 64 fp16 elements
                    mbar_expect_tx(&full[s], C::STAGE);
                    tma_load_2d(&tmA, &full[s], a_dst, kb * (HALF ? 64 : TK), (int)m0);
                    tma_load_2d(&tmB, &full[s], b_dst, kb * (HALF ? 64 : TK), n0);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (!(p.dbg & 2)) {
            constexpr uint32_t idesc = umma_idesc(HALF ? 0 /*f16*/ : 2 /*tf32*/, TM, BN);
            const bool leader = elect_one();
            for (int kb = 0; kb < p.num_kb; ++kb) {
                const int s = kb % C::NS;
                const uint32_t ph = (uint32_t)(kb / C::NS) & 1u;
                mbar_wait(X3 ? &ready[s] : &full[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * C::STAGE);
                const uint32_t b_addr = a_addr + A_BYTES;
                const uint32_t alo = X3 ? smem_u32(lo_base + s * C::STAGE) : 0u;
                if (leader) {
#pragma unroll
                    for (int k = 0; k < TK / 8; ++k) {
                        const uint64_t ad = umma_desc_k128(a_addr + k * 32);
                        const uint64_t bd = umma_desc_k128(b_addr + k * 32);
                        if (HALF)
                            mma_f16(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                        else
                            mma_tf32(tmem_base, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                        if (X3) {
                            const uint64_t ald = umma_desc_k128(alo + k * 32);
                            const uint64_t bld = umma_desc_k128(alo + A_BYTES + k * 32);
                            mma_tf32(tmem_base, ald, bd, idesc, 1u);
                            mma_tf32(tmem_base, ad, bld, idesc, 1u);
                        }
                    }
                    tc_commit(&empty[s]);
                }
                __syncwarp();
            }
            if (leader) tc_commit(tmem_full);
            __syncwarp();
        }
    } else if (warp < 2 + 4 * C::EPI_SETS) {
        // ================================ epilogue ====================================
        // The K <= 128 contractions of the path are epilogue-bound (tcgen05.ld -> math -> staged stores is one dependent
        // chain per warp), so two warps share every TMEM lane quarter and alternate over the column chunks.
        const int set = (warp - 2) >> 2;
        {   // stage the per-column epilogue vectors while the main loop runs
            const int t = threadIdx.x - 64;  // 0 .. 128 * EPI_SETS - 1
            if (t < BN) {
                const int n = n0 + t;
                float sc = 1.0f, sh = 0.0f;
                if (n < p.N) {
                    if (p.ep.colscale) {
                        sc = __ldg(p.ep.colscale + n);
                        sh = __ldg(p.ep.colshift + n);
                    }
                    if (p.ep.bias) sh += __ldg(p.ep.bias + n);
                }
                s_scale[t] = sc;
                s_shift[t] = sh;
                s_gamma[t] = (p.ep.ln_gamma && n < p.N) ? __ldg(p.ep.ln_gamma + n) : 0.0f;
                s_beta[t] = (p.ep.ln_gamma && n < p.N) ? __ldg(p.ep.ln_beta + n) : 0.0f;
            }
            if (C::EPI_SETS == 2) asm volatile("bar.sync 1, 256;" ::: "memory");
            else asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        if (!(p.dbg & 2)) mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        const int r = q * 32 + lane;
        int64_t grow;
        bool row_ok;
        if (CONV) {
            const int hl = r / CONV_TW, wl = r - hl * CONV_TW;
            grow = ((int64_t)cb * p.Ho + ch0 + hl) * p.Wo + cw0 + wl;
            row_ok = true;  // tiles divide the image exactly (checked on the host)
        } else {
            grow = m0 + r;
            row_ok = grow < p.M;
        }
        const Epilogue& ep = p.ep;
        const bool has_rd = ep.rowdiv != nullptr, has_res = !LEAN && ep.residual != nullptr, has_acc = !LEAN && ep.accumulate != 0;
        const float rd = (has_rd && row_ok) ? __ldg(ep.rowdiv + grow) : 1.0f;
        float* crow = p.C + grow * p.ldc;
        const float* rrow = has_res ? ep.residual + grow * ep.ldres : nullptr;
        const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
        float* stage = reinterpret_cast<float*>(smem) + (set * 4 + q) * (32 * 36);  // ring is idle once tmem_full fired
        // TMA-store path: two [32 rows][128 B] SWIZZLE_128B buffers per warp (1024-byte aligned), alternating per chunk
        float* tstage = reinterpret_cast<float*>(smem) + (set * 4 + q) * 2048;
        const bool tma_st = !CONV && p.tma_store != 0;
        int tbuf = 0;
        const bool rvec_ok = has_res && ((ep.ldres & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0);
        const bool has_ln = !LEAN && ep.ln_gamma != nullptr;  // host guarantees gridDim.y == 1 and N <= BN
        float ln_mean = 0.0f, ln_rstd = 1.0f;
        const bool ln_idle = has_ln && set != 0;  // a fused LayerNorm needs the whole row in one thread: set 0 does it alone
        if constexpr (!LEAN)
        if (has_ln && !ln_idle) {
            float sum = 0.0f;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t acc[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
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
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c0 + j < p.N) {
                        const float d = fmaf(__uint_as_float(acc[j]), s_scale[c0 + j], s_shift[c0 + j]) - ln_mean;
                        ssq = fmaf(d, d, ssq);
                    }
            }
            ln_rstd = rsqrtf(ssq / (float)p.N + ep.ln_eps);
        }
        const int c_begin = ln_idle ? BN : (has_ln ? 0 : set * 32), c_step = has_ln ? 32 : 32 * C::EPI_SETS;
#pragma unroll 1
        for (int c0 = c_begin; c0 < BN; c0 += c_step) {
            uint32_t acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
            tmem_ld_wait();
            const int nb = n0 + c0;
            float* tb = tstage + tbuf * 1024;
            if (tma_st) {  // the store issued from this buffer two chunks ago must have read it
                if (lane == 0) bulk_wait_read<1>();
                __syncwarp();
            }
            if (row_ok && nb < p.N && !(p.dbg & 4)) {
                const bool full = nb + 32 <= p.N;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
                if (has_rd) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = v[j] / rd;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) {  // per-column scale/shift(+bias) staged in shared memory
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
                } else if (!LEAN && ep.act == COFI_ACT_SIGMOID) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 1.0f / (1.0f + expf(-v[j]));
                }
                // fallthrough to the staged store below
                if (tma_st) {
                    float* trow = tb + lane * 32;  // row `lane` of the box; 16-byte chunk j lands at j ^ (row & 7)
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
            if (tma_st) fence_proxy_async_smem();  // staged rows -> visible to the TMA engine
            __syncwarp();
            if (p.stat_out) {
                // per-tile column statistics for the GroupNorm that follows (host guarantees M % 128 == 0, act none):
                // lane = column, 32 conflict-free shared loads over this warp's 32 staged rows
                float cs = 0.0f, cq = 0.0f;
                if (nb + lane < p.N) {   // four independent partial sums: loads back to back, chains 8 long
                    float s4[4] = {0.0f, 0.0f, 0.0f, 0.0f}, q4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) {
                        const float x = tma_st ? tb[rr * 32 + ((((lane >> 2) ^ (rr & 7)) << 2) | (lane & 3))] : stage[rr * 36 + lane];
                        s4[rr & 3] += x;
                        q4[rr & 3] = fmaf(x, x, q4[rr & 3]);
                    }
                    cs = (s4[0] + s4[1]) + (s4[2] + s4[3]);
                    cq = (q4[0] + q4[1]) + (q4[2] + q4[3]);
                }
                s_col[(q * BN + c0 + lane) * 2 + 0] = cs;
                s_col[(q * BN + c0 + lane) * 2 + 1] = cq;
            }
            // transposed write-out: 8 lanes cover the 32 columns (128 B) of one row, 4 rows per instruction, so every
            // store instruction writes four full 128-byte lines (the per-thread-row layout would scatter 16-byte pieces)
            if (tma_st) {
                // one bulk tensor store per 32 x 32 chunk: rows past M and columns past N are clipped by the TMA unit
                if (nb < p.N && !(p.dbg & 5) && lane == 0) {
                    tma_store_2d(&tmC, tb, nb, (int)m0 + q * 32);
                    bulk_commit();
                }
                tbuf ^= 1;
            } else if (nb < p.N && !(p.dbg & 5)) {
                const bool full = nb + 32 <= p.N;
                const int cc = (lane & 7) * 4;
                // dense case: row (q*32 + lane/8 + 4*it) of the tile -> one running pointer, no per-store address arithmetic
                float* drow = CONV ? nullptr : p.C + (m0 + q * 32 + (lane >> 3)) * p.ldc + nb + cc;
                const int64_t rows_left = CONV ? 0 : p.M - (m0 + q * 32 + (lane >> 3));
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + (lane >> 3);
                    float* dst;
                    if (CONV) {
                        const int rt = q * 32 + rr;  // row inside the 128-row tile
                        const int hl = rt / CONV_TW, wl = rt - hl * CONV_TW;
                        dst = p.C + (((int64_t)cb * p.Ho + ch0 + hl) * p.Wo + cw0 + wl) * p.ldc + nb + cc;
                    } else {
                        if (it * 4 >= rows_left) break;  // rows beyond M (rows grow with it)
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
            __syncwarp();  // tcgen05.ld is warp-collective: reconverge before the next chunk
        }
        if (p.stat_out) {
            if (C::EPI_SETS == 2) asm volatile("bar.sync 1, 256;" ::: "memory");
            else asm volatile("bar.sync 1, 128;" ::: "memory");
            const int t = threadIdx.x - 64;
            if (t < BN && n0 + t < p.N) {
                float cs = 0.0f, cq = 0.0f;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    cs += s_col[(qq * BN + t) * 2 + 0];
                    cq += s_col[(qq * BN + t) * 2 + 1];
                }
                float* o = p.stat_out + ((int64_t)blockIdx.x * p.N + n0 + t) * 2;
                o[0] = cs;
                o[1] = cq;
            }
        }
        if (tma_st) {  // shared memory must outlive the reads of the last stores
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
        }
        tc_fence_before();
    } else if (X3) {
        // ================================ operand splitters (3xTF32) ==================
        constexpr int NT = C::SPLIT_WARPS * 32;
        const int t = threadIdx.x - 32 * (2 + 4 * C::EPI_SETS);  // 0 .. NT-1
        for (int kb = 0; kb < p.num_kb; ++kb) {
            const int s = kb % C::NS;
            const uint32_t ph = (uint32_t)(kb / C::NS) & 1u;
            mbar_wait(&full[s], ph);
            uint4* hi = reinterpret_cast<uint4*>(smem + s * C::STAGE);
            uint4* lo = reinterpret_cast<uint4*>(lo_base + s * C::STAGE);
            // hi = x rounded to tf32 (half-up on the bit pattern: 2 integer ops), lo = (x - hi) rounded the same way: both are
            // exact inputs of the tensor core, so the only error left is the 2^-22-relative rounding of lo.  This loop sits
            // between the TMA landing and the MMA issue of every k-block: ~5 ALU operations per element.
#pragma unroll 4
            for (int i = t; i < C::STAGE / 16; i += NT) {
                const uint4 x = hi[i];
                uint4 h, l;
                h.x = (x.x + 0x1000u) & 0xFFFFE000u;
                h.y = (x.y + 0x1000u) & 0xFFFFE000u;
                h.z = (x.z + 0x1000u) & 0xFFFFE000u;
                h.w = (x.w + 0x1000u) & 0xFFFFE000u;
                l.x = (__float_as_uint(__uint_as_float(x.x) - __uint_as_float(h.x)) + 0x1000u) & 0xFFFFE000u;
                l.y = (__float_as_uint(__uint_as_float(x.y) - __uint_as_float(h.y)) + 0x1000u) & 0xFFFFE000u;
                l.z = (__float_as_uint(__uint_as_float(x.z) - __uint_as_float(h.z)) + 0x1000u) & 0xFFFFE000u;
                l.w = (__float_as_uint(__uint_as_float(x.w) - __uint_as_float(h.w)) + 0x1000u) & 0xFFFFE000u;
                hi[i] = h;
                lo[i] = l;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[s]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

template <int BN, bool CONV, bool X3, bool HALF, bool LIGHT, bool LEAN>
static int launch_one_v(const CUtensorMap* a, const CUtensorMap* b, const CUtensorMap* c, const Params& p_in, dim3 grid,
                        cudaStream_t st) {
    Params p = p_in;
    p.tma_store = (c != nullptr && !CONV) ? 1 : 0;
    if (!c) c = a;  // placeholder, never dereferenced by the kernel
    using C = Cfg<BN, X3, LIGHT>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, CONV, X3, HALF, LIGHT, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             C::SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(smem=%d): %s", C::SMEM, cudaGetErrorString(e));
            return COFI_ECUDA;
        }
        attr_done = true;
    }
    gemm_tc_kernel<BN, CONV, X3, HALF, LIGHT, LEAN><<<grid, C::THREADS, C::SMEM, st>>>(*a, *b, *c, p);
    return check_launch(CONV ? "cofi_conv2d_nhwc(tcgen05)" : "cofi_gemm(tcgen05)");
}

template <int BN, bool CONV, bool X3, bool HALF = false, bool LIGHT = false>
static int launch_one(const CUtensorMap* a, const CUtensorMap* b, const CUtensorMap* c, const Params& p, dim3 grid,
                      cudaStream_t st) {
    const Epilogue& ep = p.ep;
    const bool lean = !ep.residual && !ep.accumulate && !ep.ln_gamma && ep.act != COFI_ACT_SIGMOID;
    if (lean) return launch_one_v<BN, CONV, X3, HALF, LIGHT, true>(a, b, c, p, grid, st);
    return launch_one_v<BN, CONV, X3, HALF, LIGHT, false>(a, b, c, p, grid, st);
}

constexpr int X3_LIGHT_MAX_KB = 8;  // K <= 256

template <bool CONV>
static int dispatch(const CUtensorMap* a, const CUtensorMap* b, const CUtensorMap* c, const Params& p, int bn, bool x3,
                    dim3 grid, cudaStream_t st) {
    if (x3) {
        if (!CONV && bn <= 64 && p.num_kb <= X3_LIGHT_MAX_KB) {  // short contraction: two CTAs per SM (see Cfg)
            if (bn == 32) return launch_one<32, CONV, true, false, true>(a, b, c, p, grid, st);
            return launch_one<64, CONV, true, false, true>(a, b, c, p, grid, st);
        }
        if (bn == 32) return launch_one<32, CONV, true>(a, b, c, p, grid, st);
        if (bn == 64) return launch_one<64, CONV, true>(a, b, c, p, grid, st);
        return launch_one<128, CONV, true>(a, b, c, p, grid, st);
    }
    if (bn == 32) return launch_one<32, CONV, false>(a, b, c, p, grid, st);
    if (bn == 64) return launch_one<64, CONV, false>(a, b, c, p, grid, st);
    return launch_one<128, CONV, false>(a, b, c, p, grid, st);
}

static int tc_debug() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("COFI_TC_DEBUG");
        v = e ? atoi(e) : 0;
    }
    return v;
}

// Output-tile width: the widest of 128/64/32 that still yields at least one CTA per SM -- small problems (the
// transformer's 10240 x 128 projections are 80 row tiles) are latency-bound and want more, narrower CTAs.
static int pick_bn(int N, int64_t row_tiles = 1 << 30, bool full_row = false) {
    int bn = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
    if (full_row) return bn;
    while (bn > 32 && row_tiles * ((N + bn - 1) / bn) < 148) bn >>= 1;
    return bn;
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------- GEMM entry
bool gemm_tc_supported(int64_t lda, int64_t ldw, int64_t ldc, int64_t M, int N, int K, const void* A, const void* W,
                       const void* C) {
    (void)ldc;
    (void)C;
    if (N < 16 || M < 1 || K < 4) return false;
    if ((lda & 3) || (ldw & 3) || (K & 3)) return false;
    if (((uintptr_t)A & 15) || ((uintptr_t)W & 15)) return false;
    if (M > 0x7fffffffLL) return false;
    return tc::encode_fn() != nullptr;
}

// Output tensor map for the TMA-store epilogue: [M, N] fp32 with row pitch ldc, 32 x 32 boxes, SWIZZLE_128B.  nullptr when
// the rows are not 16-byte aligned (the kernel then stores from registers through its transposing staging area) or when
// COFI_TMA_STORE=0.
const CUtensorMap* gemm_c_tmap(const float* C, int64_t ldc, int64_t M, int N) {
    static const bool on = [] {
        const char* e = getenv("COFI_TMA_STORE");
        return !(e && e[0] == '0');
    }();
    if (!on || (ldc & 3) || ((uintptr_t)C & 15) || ldc < N) return nullptr;
    uint64_t d[2] = {(uint64_t)N, (uint64_t)M}, s[1] = {(uint64_t)ldc * 4};
    uint32_t b[2] = {32, 32};
    return tc::get_tmap_f32(C, 2, d, s, b);
}

int gemm_tc_launch(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t M, int N,
                   int K, const Epilogue& ep, int engine, cudaStream_t st, float* stat_out) {
    using namespace tc;
    int bn = pick_bn(N, ceil_div(M, TM), ep.ln_gamma != nullptr);
    // 3xTF32, short K: 64-column tiles keep the hi + lo ring under 96 KB = two CTAs per SM (a fused LayerNorm needs the
    // whole row in one tile and keeps 128)
    if (engine == COFI_GEMM_TF32X3 && bn == 128 && ep.ln_gamma == nullptr && (K + TK - 1) / TK <= X3_LIGHT_MAX_KB) bn = 64;
    uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[1] = {(uint64_t)lda * 4};
    uint32_t bA[2] = {TK, TM};
    uint64_t dB[2] = {(uint64_t)K, (uint64_t)N}, sB[1] = {(uint64_t)ldw * 4};
    uint32_t bB[2] = {TK, (uint32_t)bn};
    const CUtensorMap* ta = get_tmap_f32(A, 2, dA, sA, bA);
    const CUtensorMap* tb = get_tmap_f32(W, 2, dB, sB, bB);
    if (!ta || !tb) return COFI_ECUDA;
    Params p{};
    p.C = C;
    p.ldc = ldc;
    p.M = M;
    p.N = N;
    p.num_kb = (K + TK - 1) / TK;
    p.ep = ep;
    p.stat_out = stat_out;
    p.dbg = tc_debug();
    dim3 grid((unsigned)ceil_div(M, TM), (unsigned)ceil_div(N, bn));
    return dispatch<false>(ta, tb, gemm_c_tmap(C, ldc, M, N), p, bn, engine == COFI_GEMM_TF32X3, grid, st);
}

// fp16-operand GEMM (A [M,K] half, W [N,K] half, fp32 accumulate/output): KPConv weight-apply on the tf32 engine.
// fp16 keeps 11 significand bits -- the same operand precision as tf32 -- at half the bytes and twice the MMA rate.
int gemm_tc_f16_launch(const void* A, int64_t lda, const void* W, int64_t ldw, float* C, int64_t ldc, int64_t M, int N,
                       int K, const Epilogue& ep, cudaStream_t st, float* stat_out) {
    using namespace tc;
    const int bn = pick_bn(N, ceil_div(M, TM));
    uint64_t dA[2] = {(uint64_t)K, (uint64_t)M}, sA[1] = {(uint64_t)lda * 2};
    uint32_t bA[2] = {64, TM};
    uint64_t dB[2] = {(uint64_t)K, (uint64_t)N}, sB[1] = {(uint64_t)ldw * 2};
    uint32_t bB[2] = {64, (uint32_t)bn};
    const CUtensorMap* ta = get_tmap_f16(A, 2, dA, sA, bA);
    const CUtensorMap* tb = get_tmap_f16(W, 2, dB, sB, bB);
    if (!ta || !tb) return COFI_ECUDA;
    Params p{};
    p.C = C;
    p.ldc = ldc;
    p.M = M;
    p.N = N;
    p.num_kb = (K + 63) / 64;
    p.ep = ep;
    p.stat_out = stat_out;
    p.dbg = tc_debug();
    dim3 grid((unsigned)ceil_div(M, TM), (unsigned)ceil_div(N, bn));
    const CUtensorMap* tc_map = gemm_c_tmap(C, ldc, M, N);
    if (bn == 32) return launch_one<32, false, false, true>(ta, tb, tc_map, p, grid, st);
    if (bn == 64) return launch_one<64, false, false, true>(ta, tb, tc_map, p, grid, st);
    return launch_one<128, false, false, true>(ta, tb, tc_map, p, grid, st);
}

// ---------------------------------------------------------------------------------------------- conv entry
bool conv_tc_supported(int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad) {
    (void)B;
    if (stride != 1 && stride != 2) return false;
    const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
    if (Wo % tc::CONV_TW || Ho % tc::CONV_TH) return false;
    if (Cin % 4 || Cout < 16) return false;
    return tc::encode_fn() != nullptr;
}

int conv_tc_launch(const float* x, int B, int H, int W, int Cin, const float* w, int Cout, int KH, int KW, int stride, int pad,
                   float* y, const Epilogue& ep, int engine, cudaStream_t st) {
    using namespace tc;
    const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
    const int bn = pick_bn(Cout, (int64_t)B * (Wo / CONV_TW) * (Ho / CONV_TH));
    const int Ktot = KH * KW * Cin;
    uint64_t dA[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t sA[3] = {(uint64_t)Cin * 4, (uint64_t)W * Cin * 4, (uint64_t)H * W * Cin * 4};
    // stride 2 (the ResNet stem and layer2's first block, reference model/imagenet.py:199-212): the box spans 2x the
    // pixels and elementStrides {1, 2, 2, 1} keep every second one, so the tile lands dense in shared memory
    uint32_t bA[4] = {TK, (uint32_t)(CONV_TW * stride), (uint32_t)(CONV_TH * stride), 1};
    uint32_t eA[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    uint64_t dB[2] = {(uint64_t)Ktot, (uint64_t)Cout}, sB[1] = {(uint64_t)Ktot * 4};
    uint32_t bB[2] = {TK, (uint32_t)bn};
    const CUtensorMap* ta = stride == 1 ? get_tmap_f32(x, 4, dA, sA, bA) : get_tmap_any(x, 4, dA, sA, bA, 0, eA, 0);
    const CUtensorMap* tb = get_tmap_f32(w, 2, dB, sB, bB);
    if (!ta || !tb) return COFI_ECUDA;
    Params p{};
    p.C = y;
    p.ldc = Cout;
    p.M = (int64_t)B * Ho * Wo;
    p.N = Cout;
    p.cpt = (Cin + TK - 1) / TK;
    p.num_kb = KH * KW * p.cpt;
    p.ep = ep;
    p.Ho = Ho;
    p.Wo = Wo;
    p.Cin = Cin;
    p.KW = KW;
    p.pad = pad;
    p.cstride = stride;
    p.tiles_w = Wo / CONV_TW;
    p.tiles_per_img = p.tiles_w * (Ho / CONV_TH);
    dim3 grid((unsigned)(B * p.tiles_per_img), (unsigned)ceil_div(Cout, bn));
    return dispatch<true>(ta, tb, nullptr, p, bn, engine == COFI_GEMM_TF32X3, grid, st);
}

}  // namespace cofi
