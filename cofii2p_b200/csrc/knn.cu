// knn.cu -- exact KNN-k tables of the point pyramid, built on the device.
//
// Replaces the table builder that runs immediately before the hot path: the reference's dataset calls
// precompute_point_cloud_stack_mode (model/kpconv/preprocess_data.py:36-107; open3d KNNSearch: true squared distance,
// ascending, self first) and offers precompute_point_cloud_cuda (:131-203; `knn()` = expanded-form distance + topk).
// 13 tables per frame (5 x neighbors, 4 x subsampling, 4 x upsampling), k = 128, about 1.1 G point pairs brute force.
//
// Design (no tensor cores: integer/selection work, ALU + L1/L2 bound):
//   1. knn_sort_chunks_kernel + knn_sort_kernel: bounding box -> 30-bit Morton code -> bitonic sort of
//      (code << 32 | index) keys (4096-key chunks sorted by one CTA each in shared memory, then one CTA per set merges
//      them: 8192-key shared-memory chunks, the few long-stride stages through L2) -> the set re-ordered as float4
//      (x, y, z, original index), the sorted codes, one AABB per 32 consecutive points and one per 32 such tiles.
//   2. knn_query_kernel: one warp per query, queries taken in their own Morton order so that the 8 warps of a CTA touch
//      the same tiles.  The warp seeds its candidate list from the 4 tiles around the query's position in the source
//      order, tightens it on the 8 tiles next to those, then sweeps all remaining tiles 32 AABBs at a time (one per
//      lane) and opens only tiles whose box can still beat the current k-th key.  Candidates are 64-bit keys
//      (distance bits << 32 | index): ascending key order IS the reference order (distance, then index), so the result
//      does not depend on the visiting order, on the sort, or on the culling -- it is the exact brute-force table.
//      Survivors are appended to a 128-entry shared-memory buffer by ballot/popc and folded into the sorted best list,
//      which lives in registers (4 keys per lane), by a fully unrolled bitonic sort + half-cleaner + bitonic merge whose
//      long strides are warp shuffles.
// The culling test is exact: in COFI_KNN_DIRECT the box distance is evaluated with the same rounded operations as a
// point distance and IEEE rounding is monotonic, so box <= every point inside, in floating point; in
// COFI_KNN_EXPANDED (the |a|^2+|b|^2-2ab form, whose fp32 cancellation noise reaches 1e-3 m^2 at 80 m) a margin
// bounding that noise is added.
#include <math.h>

#include <type_traits>

#include "common.cuh"

namespace cofi {
namespace {

constexpr int KB = 128;  // capacity of the best list and of the candidate buffer (k <= 128)
constexpr int QWARPS = 8;
constexpr int SORT_THREADS = 1024;
constexpr int SORT_CHUNK = 8192;    // shared-memory chunk of the merge step
constexpr int SORT_CHUNK_A = 4096;  // chunk sorted by one CTA in the first step
constexpr int MAX_SETS = 8;
constexpr int MAX_JOBS = 16;
constexpr unsigned long long KMAX = ~0ull;

struct SetDesc {
    const float* pts;          // [frames*n, 3]
    unsigned long long* keys;  // [frames, P]
    float4* sorted;            // [frames, npad]  (x, y, z, original index as int bits; pad = +inf / -1)
    uint32_t* codes;           // [frames, npad]
    float4* tmin;              // [frames, npad/32]  (min x, min y, min z, max |s|^2)
    float4* tmax;              // [frames, npad/32]
    float4* smin;              // [frames, nsup]  boxes of 32 consecutive tiles (1024 points), same layout
    float4* smax;              // [frames, nsup]
    float4* info;              // [frames]  (bbox min xyz, cells per metre)
    int n, npad, P, nsup;
};
struct SortParams {
    SetDesc set[MAX_SETS];
    int chunk_begin[MAX_SETS + 1];
    int nsets;
};
struct JobDesc {
    int64_t* out;             // [frames*nq, k]
    const int64_t* copy_from; // optional: the finished [frames*ns, k] table of the SOURCE set queried by itself.  A query whose
                              // coordinates equal a source point's has, by definition, that point's row -- the sub-sampled levels
                              // of the pyramid are subsets of the level above, so the 4 `subsampling` tables are row copies
    int src, qry, block_begin, k;
};
struct QueryParams {
    SetDesc set[MAX_SETS];
    JobDesc job[MAX_JOBS];
    int njobs, cull, fast;
};

__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 1023u;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t morton_code(float x, float y, float z, const float4 info) {
    const float fx = fminf(fmaxf((x - info.x) * info.w, 0.0f), 1023.0f);
    const float fy = fminf(fmaxf((y - info.y) * info.w, 0.0f), 1023.0f);
    const float fz = fminf(fmaxf((z - info.z) * info.w, 0.0f), 1023.0f);
    return spread10((uint32_t)fx) | (spread10((uint32_t)fy) << 1) | (spread10((uint32_t)fz) << 2);
}

// ------------------------------------------------------------------------------------------------ sort
__device__ __forceinline__ void bitonic_stage_smem(unsigned long long* sm, int len, int base, int k, int j) {
    for (int p = threadIdx.x; p < (len >> 1); p += SORT_THREADS) {
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        const int l = i | j;
        const bool asc = (((base + i) & k) == 0);
        const unsigned long long a = sm[i], b = sm[l];
        if ((a > b) == asc) {
            sm[i] = b;
            sm[l] = a;
        }
    }
}

// Step 1 of the sort, one CTA per (4096-key chunk, frame): bounding box of the whole set (recomputed by every chunk: a
// few thousand loads), keys = (morton << 32 | index) padded with KMAX, full bitonic sort of the chunk in shared memory
// (ascending or descending by chunk parity, as the global bitonic network wants it).
__global__ void __launch_bounds__(SORT_THREADS) knn_sort_chunks_kernel(const __grid_constant__ SortParams P) {
    __shared__ unsigned long long sm[SORT_CHUNK_A];
    __shared__ float red[6][32];
    __shared__ float4 s_info;
    int si = 0;
    while (si + 1 < P.nsets && (int)blockIdx.x >= P.chunk_begin[si + 1]) ++si;
    const SetDesc& S = P.set[si];
    const int chunk = blockIdx.x - P.chunk_begin[si];
    const int frame = blockIdx.y, n = S.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* pts = S.pts + (size_t)frame * n * 3;

    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < n; i += SORT_THREADS) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(pts + (size_t)i * 3 + a);
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
        if (lane == 0) {
            red[a][warp] = mn[a];
            red[3 + a][warp] = mx[a];
        }
    }
    __syncthreads();
    if (tid == 0) {
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            lo[a] = red[a][0];
            hi[a] = red[3 + a][0];
            for (int w = 1; w < 32; ++w) {
                lo[a] = fminf(lo[a], red[a][w]);
                hi[a] = fmaxf(hi[a], red[3 + a][w]);
            }
        }
        const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
        s_info = make_float4(lo[0], lo[1], lo[2], ext > 0.0f ? 1023.5f / ext : 0.0f);
        if (chunk == 0) S.info[frame] = s_info;
    }
    __syncthreads();
    const float4 info = s_info;

    const int CH = S.P < SORT_CHUNK_A ? S.P : SORT_CHUNK_A, base = chunk * CH;
    for (int i = tid; i < CH; i += SORT_THREADS) {
        unsigned long long key = KMAX;
        const int g = base + i;
        if (g < n) {
            const float x = __ldg(pts + (size_t)g * 3), y = __ldg(pts + (size_t)g * 3 + 1), z = __ldg(pts + (size_t)g * 3 + 2);
            key = ((unsigned long long)morton_code(x, y, z, info) << 32) | (unsigned)g;
        }
        sm[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= CH; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            bitonic_stage_smem(sm, CH, base, k, j);
            __syncthreads();
        }
    unsigned long long* keys = S.keys + (size_t)frame * S.P;
    for (int i = tid; i < CH; i += SORT_THREADS) keys[base + i] = sm[i];
}

// Step 2, one CTA per (set, frame): the remaining bitonic merge levels (strides >= 8192 through L2, the rest in
// 8192-key shared-memory chunks), then the re-ordered set, its codes and the tile / group boxes.
__global__ void __launch_bounds__(SORT_THREADS) knn_sort_kernel(const __grid_constant__ SortParams P) {
    extern __shared__ unsigned long long sm[];  // SORT_CHUNK keys
    const SetDesc& S = P.set[blockIdx.x];
    const int frame = blockIdx.y, n = S.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* pts = S.pts + (size_t)frame * n * 3;
    unsigned long long* keys = S.keys + (size_t)frame * S.P;

    const int CH = S.P < SORT_CHUNK ? S.P : SORT_CHUNK;
    for (int k = SORT_CHUNK_A << 1; k <= S.P; k <<= 1) {
        for (int j = k >> 1; j >= CH; j >>= 1) {
            for (int p = tid; p < (S.P >> 1); p += SORT_THREADS) {
                const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                const int l = i | j;
                const bool asc = ((i & k) == 0);
                const unsigned long long a = keys[i], b = keys[l];
                if ((a > b) == asc) {
                    keys[i] = b;
                    keys[l] = a;
                }
            }
            __syncthreads();
        }
        for (int base = 0; base < S.P; base += CH) {
            for (int i = tid; i < CH; i += SORT_THREADS) sm[i] = keys[base + i];
            __syncthreads();
            for (int j = (k >> 1) < (CH >> 1) ? (k >> 1) : (CH >> 1); j > 0; j >>= 1) {
                bitonic_stage_smem(sm, CH, base, k, j);
                __syncthreads();
            }
            for (int i = tid; i < CH; i += SORT_THREADS) keys[base + i] = sm[i];
            __syncthreads();
        }
    }

    // emit the re-ordered set, its codes and one box per 32 points (npad is a multiple of 32: warps stay whole)
    float4* sorted = S.sorted + (size_t)frame * S.npad;
    uint32_t* codes = S.codes + (size_t)frame * S.npad;
    float4* tmin = S.tmin + (size_t)frame * (S.npad >> 5);
    float4* tmax = S.tmax + (size_t)frame * (S.npad >> 5);
    for (int i = tid; i < S.npad; i += SORT_THREADS) {
        float x = INFINITY, y = INFINITY, z = INFINITY;
        int idx = -1;
        uint32_t code = 0xffffffffu;
        if (i < n) {
            const unsigned long long key = keys[i];
            idx = (int)(unsigned)key;
            code = (uint32_t)(key >> 32);
            x = __ldg(pts + (size_t)idx * 3);
            y = __ldg(pts + (size_t)idx * 3 + 1);
            z = __ldg(pts + (size_t)idx * 3 + 2);
        }
        sorted[i] = make_float4(x, y, z, __int_as_float(idx));
        codes[i] = code;
        float lx = x, ly = y, lz = z;
        float hx = idx >= 0 ? x : -INFINITY, hy = idx >= 0 ? y : -INFINITY, hz = idx >= 0 ? z : -INFINITY;
        float ss = idx >= 0 ? x * x + y * y + z * z : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o));
            ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o));
            lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o));
            hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o));
            hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o));
            hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
            ss = fmaxf(ss, __shfl_xor_sync(0xffffffffu, ss, o));
        }
        if (lane == 0) {
            tmin[i >> 5] = make_float4(lx, ly, lz, ss);
            tmax[i >> 5] = make_float4(hx, hy, hz, 0.0f);
        }
    }
    __syncthreads();
    // boxes of 32 consecutive tiles: the sweep tests these first and skips whole groups
    const int T = S.npad >> 5;
    for (int g = warp; g < S.nsup; g += SORT_THREADS / 32) {
        const int t = (g << 5) + lane;
        float4 lo = make_float4(INFINITY, INFINITY, INFINITY, 0.0f), hi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.0f);
        if (t < T) {
            lo = tmin[t];
            hi = tmax[t];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o));
            lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
            lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
            lo.w = fmaxf(lo.w, __shfl_xor_sync(0xffffffffu, lo.w, o));
            hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
            hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
            hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
        }
        if (lane == 0) {
            S.smin[(size_t)frame * S.nsup + g] = lo;
            S.smax[(size_t)frame * S.nsup + g] = hi;
        }
    }
}

// ------------------------------------------------------------------------------------------------ query
template <int MODE>
__device__ __forceinline__ float dist2(float qx, float qy, float qz, float qq, const float4 s) {
    if (MODE == COFI_KNN_DIRECT) {
        const float dx = __fsub_rn(qx, s.x), dy = __fsub_rn(qy, s.y), dz = __fsub_rn(qz, s.z);
        return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    } else {
        // model/kpconv/preprocess_data.py:120-129: -2 q.s, += |q|^2, += |s|^2, clamp(min=1e-12)
        const float dot = __fadd_rn(__fadd_rn(__fmul_rn(qx, s.x), __fmul_rn(qy, s.y)), __fmul_rn(qz, s.z));
        const float ss = __fadd_rn(__fadd_rn(__fmul_rn(s.x, s.x), __fmul_rn(s.y, s.y)), __fmul_rn(s.z, s.z));
        float d = __fmul_rn(-2.0f, dot);
        d = __fadd_rn(d, qq);
        d = __fadd_rn(d, ss);
        return fmaxf(d, 1e-12f);
    }
}

typedef unsigned long long u64;

// 128 keys of a warp held in registers, element e = lane * 4 + r
struct Keys4 {
    u64 v[4];
};

// one compare-exchange stage (stride J) of the bitonic network of block size K over the 128 register-resident keys:
// strides 1, 2 stay inside a lane, strides 4..64 are lane-xor shuffles.  Fully unrolled: no index arithmetic, no
// shared memory.
template <int J, int K>
__device__ __forceinline__ void bitonic_stage_reg(Keys4& x, int lane) {
    if constexpr (J >= 4) {
        const bool keep_min = ((lane & (J >> 2)) == 0) == (((lane << 2) & K) == 0);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const u64 p = __shfl_xor_sync(0xffffffffu, x.v[r], J >> 2);
            x.v[r] = ((x.v[r] < p) == keep_min) ? x.v[r] : p;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if ((r & J) == 0) {
                const bool asc = ((((lane << 2) | r) & K) == 0);
                const u64 a = x.v[r], b = x.v[r | J];
                const bool sw = (a > b) == asc;
                x.v[r] = sw ? b : a;
                x.v[r | J] = sw ? a : b;
            }
    }
}
template <int J, int K>
__device__ __forceinline__ void bitonic_stages_reg(Keys4& x, int lane) {
    bitonic_stage_reg<J, K>(x, lane);
    if constexpr (J > 1) bitonic_stages_reg<J / 2, K>(x, lane);
}
template <int K>
__device__ __forceinline__ void bitonic_sort_reg(Keys4& x, int lane) {
    if constexpr (K > 2) bitonic_sort_reg<K / 2>(x, lane);
    bitonic_stages_reg<K / 2, K>(x, lane);
}

// fold `cnt` pending candidates (shared memory) into the sorted best list (registers): sort the candidates ascending,
// half-cleaner against the reversed candidates (the 128 smallest of both, as a bitonic sequence), bitonic merge.
__device__ __forceinline__ Keys4 knn_merge_keys(Keys4 best, const u64* cand, int cnt, int lane) {
    __syncwarp();
    Keys4 c;
#pragma unroll
    for (int r = 0; r < 4; ++r) c.v[r] = (lane * 4 + r) < cnt ? cand[lane * 4 + r] : KMAX;
    bitonic_sort_reg<KB>(c, lane);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const u64 p = __shfl_sync(0xffffffffu, c.v[3 - r], 31 - lane);  // element 127 - e
        best.v[r] = best.v[r] < p ? best.v[r] : p;
    }
    bitonic_stages_reg<KB / 2, KB>(best, lane);
    __syncwarp();
    return best;
}

struct WarpState {
    Keys4 best;  // ascending, element e = lane * 4 + r
    u64* cand;   // [KB] pending candidates (shared memory)
    u64 thresh;  // current k-th key
    int cnt, k, lane;
};

__device__ __forceinline__ u64 knn_best_at(const WarpState& w, int e) {  // e is warp-uniform
    const int r = e & 3;
    const u64 t = r == 0 ? w.best.v[0] : (r == 1 ? w.best.v[1] : (r == 2 ? w.best.v[2] : w.best.v[3]));
    return __shfl_sync(0xffffffffu, t, e >> 2);
}

__device__ __forceinline__ void knn_merge(WarpState& w) {
    w.best = knn_merge_keys(w.best, w.cand, w.cnt, w.lane);
    w.thresh = knn_best_at(w, w.k - 1);
    w.cnt = 0;
}

template <int MODE>
__device__ __forceinline__ void knn_open_tile(WarpState& w, const float4* __restrict__ ssort, int t, float qx, float qy,
                                              float qz, float qq) {
    const float4 s = __ldg(ssort + ((size_t)t << 5) + w.lane);
    const int si = __float_as_int(s.w);
    const float d = dist2<MODE>(qx, qy, qz, qq, s);
    const u64 key = ((u64)__float_as_uint(d) << 32) | (unsigned)si;
    const bool pass = si >= 0 && key < w.thresh;
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m) {
        if (pass) w.cand[w.cnt + __popc(m & ((1u << w.lane) - 1u))] = key;
        w.cnt += __popc(m);
    }
}

// squared distance from the query to the box of tile t, with the rounded operations of a point distance
__device__ __forceinline__ float knn_box_dist(const float4* __restrict__ tmin, const float4* __restrict__ tmax, int t,
                                              float qx, float qy, float qz, float& smax) {
    const float4 lo = __ldg(tmin + t), hi = __ldg(tmax + t);
    const float dx = fmaxf(fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)), 0.0f);
    const float dy = fmaxf(fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)), 0.0f);
    const float dz = fmaxf(fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)), 0.0f);
    smax = lo.w;
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// can a tile whose box is at squared distance d still hold a key below the current k-th key?
template <int MODE>
__device__ __forceinline__ bool knn_box_pass(const WarpState& w, float d, float qq, float smax) {
    if (w.thresh == KMAX) return true;
    const float td = __uint_as_float((unsigned)(w.thresh >> 32));
    if (MODE == COFI_KNN_DIRECT) return d <= td;  // rounding is monotonic: box <= any point inside, exactly
    return !(d > td + 4e-6f * (qq + smax) + 1e-12f);
}

// ---- 256 keys of a warp in registers (element e = lane * 8 + r): the final sort of the fast path ----
struct Keys8 {
    u64 v[8];
};
template <int J, int K>
__device__ __forceinline__ void bitonic_stage_reg8(Keys8& x, int lane) {
    if constexpr (J >= 8) {
        const bool keep_min = ((lane & (J >> 3)) == 0) == (((lane << 3) & K) == 0);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const u64 p = __shfl_xor_sync(0xffffffffu, x.v[r], J >> 3);
            x.v[r] = ((x.v[r] < p) == keep_min) ? x.v[r] : p;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if ((r & J) == 0) {
                const bool asc = ((((lane << 3) | r) & K) == 0);
                const u64 a = x.v[r], b = x.v[r | J];
                const bool sw = (a > b) == asc;
                x.v[r] = sw ? b : a;
                x.v[r | J] = sw ? a : b;
            }
    }
}
template <int J, int K>
__device__ __forceinline__ void bitonic_stages_reg8(Keys8& x, int lane) {
    bitonic_stage_reg8<J, K>(x, lane);
    if constexpr (J > 1) bitonic_stages_reg8<J / 2, K>(x, lane);
}
template <int K>
__device__ __forceinline__ void bitonic_sort_reg8(Keys8& x, int lane) {
    if constexpr (K > 2) bitonic_sort_reg8<K / 2>(x, lane);
    bitonic_stages_reg8<K / 2, K>(x, lane);
}

constexpr int FCAP = 512;    // candidate capacity of the fast path
constexpr int FSORT = 256;   // keys the final register sort takes (k <= 128 plus ties / slack)
constexpr int FSLACK = 48;   // a tightening stops as soon as k <= count <= k + FSLACK
constexpr int FTRIG = 96;    // tighten again once k + FTRIG candidates are pending

// Fast path of a k > 1 query: select by THRESHOLD instead of by repeated sorting.
//   * candidates are 64-bit keys (distance bits << 32 | index) appended to a shared-memory list when distance <= t1
//     (t1 = +inf at the start);
//   * `tighten`: a bisection on the distance bit pattern over the pending list finds the smallest window t with
//     k <= #{d <= t} <= k + FSLACK (one count per step: <= 16 compares per lane and one redux), the list is compacted to the
//     survivors and t1 = t.  t1 is an upper bound of the final k-th distance at all times, because the list holds every
//     point seen so far with d <= t1 and at least k of them survive;
//   * order of work: the <= 12 tiles around the query's position in the source order, tighten, then every other tile whose
//     box can hold a distance <= t1 (group boxes first), tightening whenever k + FTRIG candidates are pending;
//   * ONE register sort (128 or 256 keys) puts the survivors in (distance, index) order; the first k are the table row.
// Every point of the true result has d <= final k-th distance <= t1 when its tile is tested, and the box test is exact, so
// the survivor set contains the result: identical output to the iterative path.  Returns false (nothing written) when more
// than FSORT keys tie inside the last window or the frame has fewer than k points; the caller then runs the iterative path.
template <int MODE>
__device__ __forceinline__ bool knn_fast_path(const JobDesc& J, const SetDesc& S, u64* cand, int lane, int frame, int qi, int qn,
                                              float qx, float qy, float qz, float qq, int r0, int r1) {
    const int ns = S.n, T = S.npad >> 5, k = J.k;
    if (ns < k) return false;
    const float4* ssort = S.sorted + (size_t)frame * S.npad;
    const float4* tmin = S.tmin + (size_t)frame * T;
    const float4* tmax = S.tmax + (size_t)frame * T;
    const float4* gmin = S.smin + (size_t)frame * S.nsup;
    const float4* gmax = S.smax + (size_t)frame * S.nsup;
    int cnt = 0;
    unsigned t1 = 0xffffffffu;
    WarpState w;   // only .thresh is used (knn_box_pass)
    w.thresh = KMAX;

    auto open = [&](int t) {
        const float4 s = __ldg(ssort + ((size_t)t << 5) + lane);
        const int si = __float_as_int(s.w);
        const unsigned db = __float_as_uint(dist2<MODE>(qx, qy, qz, qq, s));
        const bool keep = si >= 0 && db <= t1;
        const unsigned pm = __ballot_sync(0xffffffffu, keep);
        if (keep) cand[cnt + __popc(pm & ((1u << lane) - 1u))] = ((u64)db << 32) | (unsigned)si;
        cnt += __popc(pm);
    };
    // shrink the list to the smallest prefix-by-distance holding >= k keys (<= hard_max), false if ties make that impossible.
    // NS = register slots per lane (8 covers lists of <= 256 keys, the common case; 16 the full capacity).
    auto tighten_n = [&](int hard_max, auto ns_tag) -> bool {
        constexpr int NS = decltype(ns_tag)::value;
        __syncwarp();
        unsigned d[NS];
        unsigned hi = 0u;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int e = i * 32 + lane;
            d[i] = e < cnt ? (unsigned)(cand[e] >> 32) : 0xffffffffu;
            if (e < cnt) hi = max(hi, d[i]);
        }
        hi = __reduce_max_sync(0xffffffffu, hi);
        auto count_le = [&](unsigned t) {
            int c = 0;
#pragma unroll
            for (int i = 0; i < NS; ++i) c += d[i] <= t ? 1 : 0;
            return (int)__reduce_add_sync(0xffffffffu, c);
        };
        // smallest-t search for count >= k, stopped inside the slack window; the search starts 2^-8 below the largest
        // pending distance and falls back to [0, that) when even that holds k keys
        unsigned lo = hi > (8u << 23) ? hi - (8u << 23) : 0u;
        int c_hi = cnt;
        if (lo > 0u) {
            const int c = count_le(lo);
            if (c >= k) {
                hi = lo;
                c_hi = c;
                lo = 0u;
            } else {
                ++lo;
            }
        }
        const int want = min(k + FSLACK, hard_max);
        while (lo < hi && c_hi > want) {
            const unsigned mid = lo + ((hi - lo) >> 1);
            const int c = count_le(mid);
            if (c >= k) {
                hi = mid;
                c_hi = c;
            } else {
                lo = mid + 1;
            }
        }
        if (c_hi > hard_max) return false;
        int ncnt = 0;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            if (i * 32 < cnt) {   // warp-uniform
                const int e = i * 32 + lane;
                const bool keep = d[i] <= hi;   // slots past cnt hold 0xffffffff > hi
                const u64 key = e < cnt ? cand[e] : 0ull;
                __syncwarp();   // every lane has read its slot of this chunk before any lane overwrites one
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                if (keep) cand[ncnt + __popc(m & ((1u << lane) - 1u))] = key;
                ncnt += __popc(m);
                __syncwarp();
            }
        }
        cnt = ncnt;
        t1 = hi;
        w.thresh = ((u64)t1 << 32) | 0xffffffffull;
        return true;
    };
    auto tighten = [&](int hard_max) -> bool {
        if (cnt <= 256) return tighten_n(hard_max, std::integral_constant<int, 8>());
        return tighten_n(hard_max, std::integral_constant<int, FCAP / 32>());
    };

    for (int t = r0; t < r1; ++t) open(t);   // <= 12 tiles, 384 keys
    if (cnt >= k && !tighten(FCAP - 64)) return false;
    for (int gb = 0; gb < S.nsup; gb += 32) {
        bool gok = gb + lane < S.nsup;
        if (gok) {
            float smax;
            const float d = knn_box_dist(gmin, gmax, gb + lane, qx, qy, qz, smax);
            gok = knn_box_pass<MODE>(w, d, qq, smax);
        }
        unsigned gm = __ballot_sync(0xffffffffu, gok);
        while (gm) {
            const int rd = gb + __ffs(gm) - 1;
            gm &= gm - 1;
            const int t = (rd << 5) + lane;
            bool ok = t < T && (t < r0 || t >= r1);
            if (ok) {
                float smax;
                const float d = knn_box_dist(tmin, tmax, t, qx, qy, qz, smax);
                ok = knn_box_pass<MODE>(w, d, qq, smax);
            }
            unsigned m = __ballot_sync(0xffffffffu, ok);
            while (m) {
                const int tt = (rd << 5) + __ffs(m) - 1;
                m &= m - 1;
                if (tt < r0 || tt >= r1) {
                    // the box test above used the threshold of the start of this batch of 32: a tightening in between only
                    // makes it conservative
                    open(tt);
                    if (cnt >= k + FTRIG || cnt > FCAP - 32) {
                        if (cnt < k || !tighten(FCAP - 64)) return false;
                    }
                }
            }
        }
    }
    if (cnt < k) return false;
    if (cnt > FSORT && !tighten(FSORT)) return false;
    __syncwarp();
    int64_t* out = J.out + ((size_t)frame * qn + qi) * k;
    if (cnt <= KB) {  // at most 128 survivors: the 128-key network is enough
        Keys4 c;
#pragma unroll
        for (int r = 0; r < 4; ++r) c.v[r] = (lane * 4 + r) < cnt ? cand[lane * 4 + r] : KMAX;
        bitonic_sort_reg<KB>(c, lane);
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (lane * 4 + r < k) out[lane * 4 + r] = (int64_t)(unsigned)c.v[r];
    } else {
        Keys8 c;
#pragma unroll
        for (int r = 0; r < 8; ++r) c.v[r] = (lane * 8 + r) < cnt ? cand[lane * 8 + r] : KMAX;
        bitonic_sort_reg8<FSORT>(c, lane);
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (lane * 8 + r < k) out[lane * 8 + r] = (int64_t)(unsigned)c.v[r];
    }
    __syncwarp();
    return true;
}

template <int MODE>
__global__ void __launch_bounds__(QWARPS * 32) knn_query_kernel(const __grid_constant__ QueryParams P) {
    __shared__ unsigned long long s_cand[QWARPS][FCAP];
    int j = 0;
    while (j + 1 < P.njobs && (int)blockIdx.x >= P.job[j + 1].block_begin) ++j;
    const JobDesc& J = P.job[j];
    const SetDesc& S = P.set[J.src];
    const SetDesc& Q = P.set[J.qry];
    const int frame = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qp = ((int)blockIdx.x - J.block_begin) * QWARPS + warp;
    if (qp >= Q.n) return;  // whole warp; no block-level barrier below

    const float4 q4 = __ldg(Q.sorted + (size_t)frame * Q.npad + qp);
    const int qi = __float_as_int(q4.w);
    const float qx = q4.x, qy = q4.y, qz = q4.z;
    const float qq = __fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(qz, qz));
    const int ns = S.n, T = S.npad >> 5;
    const float4* ssort = S.sorted + (size_t)frame * S.npad;
    const float4* tmin = S.tmin + (size_t)frame * T;
    const float4* tmax = S.tmax + (size_t)frame * T;

    WarpState w;
    w.cand = s_cand[warp];
    w.thresh = KMAX;
    w.cnt = 0;
    w.k = J.k;
    w.lane = lane;
#pragma unroll
    for (int r = 0; r < 4; ++r) w.best.v[r] = KMAX;

    // position of the query in the source's Morton order
    int home = qp;
    if (J.src != J.qry) {
        const uint32_t code = morton_code(qx, qy, qz, __ldg(S.info + frame));
        const uint32_t* codes = S.codes + (size_t)frame * S.npad;
        int lo = 0, hi = ns;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(codes + mid) < code) lo = mid + 1;
            else hi = mid;
        }
        home = lo;
    }
    const int ht = min(home >> 5, T - 1);
    const int s0 = max(0, min(ht - 1, T - 4)), s1 = min(T, s0 + 4);
    const int r0 = max(0, s0 - 4), r1 = min(T, s1 + 4);
    const float4* gmin = S.smin + (size_t)frame * S.nsup;
    const float4* gmax = S.smax + (size_t)frame * S.nsup;

    if (J.k == 1) {
        // Nearest neighbour only (the up-sampling tables of the engine: the model reads column 0, reference
        // model/kpconv/functional.py:20): no candidate buffer, no sorting network -- the running best key is a warp-wide
        // minimum (two redux.sync per tile), the seeds give the first bound, one culled sweep does the rest.
        auto open1 = [&](int t) {
            const float4 s = __ldg(ssort + ((size_t)t << 5) + lane);
            const int si = __float_as_int(s.w);
            const float d = dist2<MODE>(qx, qy, qz, qq, s);
            const unsigned hi = si >= 0 ? __float_as_uint(d) : 0xffffffffu;
            const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
            const unsigned ml = __reduce_min_sync(0xffffffffu, (hi == mh && si >= 0) ? (unsigned)si : 0xffffffffu);
            const u64 m = ((u64)mh << 32) | ml;
            if (m < w.thresh) w.thresh = m;
        };
        for (int t = s0; t < s1; ++t) open1(t);
        for (int gb = 0; gb < S.nsup; gb += 32) {
            bool gok = gb + lane < S.nsup;
            if (gok && P.cull) {
                float smax;
                const float d = knn_box_dist(gmin, gmax, gb + lane, qx, qy, qz, smax);
                gok = knn_box_pass<MODE>(w, d, qq, smax);
            }
            unsigned gm = __ballot_sync(0xffffffffu, gok);
            while (gm) {
                const int rd = gb + __ffs(gm) - 1;
                gm &= gm - 1;
                const int t = (rd << 5) + lane;
                bool ok = t < T && (t < s0 || t >= s1);
                if (ok && P.cull) {
                    float smax;
                    const float d = knn_box_dist(tmin, tmax, t, qx, qy, qz, smax);
                    ok = knn_box_pass<MODE>(w, d, qq, smax);
                }
                unsigned m = __ballot_sync(0xffffffffu, ok);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    open1((rd << 5) + b);
                }
            }
        }
        if (lane == 0)
            J.out[(size_t)frame * Q.n + qi] = w.thresh == KMAX ? (int64_t)ns : (int64_t)(unsigned)w.thresh;
        return;
    }

    // a query that coincides with a source point copies that point's row of the source's own table (see JobDesc::copy_from):
    // equal coordinates have equal Morton codes, `home` is the first of that run, so the twin sits in the seed tiles
    if (P.fast && J.copy_from != nullptr) {
        int twin = -1;
        for (int t = s0; t < s1 && twin < 0; ++t) {
            const float4 s = __ldg(ssort + ((size_t)t << 5) + lane);
            const bool same = __float_as_int(s.w) >= 0 && s.x == qx && s.y == qy && s.z == qz;
            const unsigned m = __ballot_sync(0xffffffffu, same);
            if (m) twin = __shfl_sync(0xffffffffu, __float_as_int(s.w), __ffs(m) - 1);
        }
        if (twin >= 0) {
            const int64_t* srow = J.copy_from + ((size_t)frame * ns + twin) * J.k;
            int64_t* orow = J.out + ((size_t)frame * Q.n + qi) * J.k;
            for (int e = lane; e < J.k; e += 32) orow[e] = __ldg(srow + e);
            return;
        }
    }
    // threshold selection + one sort (see knn_fast_path); the iterative path below is the general fallback
    if (P.fast) {
        if (knn_fast_path<MODE>(J, S, w.cand, lane, frame, qi, Q.n, qx, qy, qz, qq, r0, r1)) return;
    }

    // Four phases through one loop body (the unrolled merge network exists twice in the code: overflow and phase end):
    //   0  seeds: the 4 tiles around the query's position in the source order, no culling
    //   1  the 8 tiles next to the seeds, culled against the seed threshold
    //   2  sweep of every other tile -- first the boxes of 32-tile groups, one per lane, then 32 tile boxes per step inside
    //      the groups that pass -- restricted to the radius at which a uniform surface would hold
    //      k points (twice the radius of the current 32nd key, i.e. 4x its squared distance): after it the k-th key is
    //      close to final
    //   3  second sweep: whatever else the tightened k-th key still admits
    // Pending candidates are folded in when the buffer could overflow and at the end of every phase.
    float near_d = INFINITY;
    for (int phase = 0; phase < 4; ++phase) {
        if (phase == 2 && P.cull && w.thresh != KMAX)
            near_d = 4.0f * __uint_as_float((unsigned)(knn_best_at(w, min(31, w.k - 1)) >> 32));
        if (phase == 3 && near_d == INFINITY) break;  // the first sweep already covered every tile
        const int rounds = phase < 2 ? 1 : S.nsup;
        for (int gb = 0; gb < rounds; gb += 32) {
            // which groups of 32 tiles can matter at all (one group box per lane)
            bool gok = gb + lane < rounds;
            if (gok && phase >= 2 && P.cull) {
                float smax;
                const float d = knn_box_dist(gmin, gmax, gb + lane, qx, qy, qz, smax);
                gok = (phase == 3 || d <= near_d) && knn_box_pass<MODE>(w, d, qq, smax);
            }
            unsigned gm = __ballot_sync(0xffffffffu, gok);
            while (gm) {
                const int rd = gb + __ffs(gm) - 1;
                gm &= gm - 1;
                int t;
                bool ok;
                if (phase == 0) {
                    t = s0 + lane;
                    ok = t < s1;
                } else if (phase == 1) {
                    t = lane < 4 ? s0 - 4 + lane : s1 + lane - 4;
                    ok = lane < 8 && t >= 0 && t < T;
                } else {
                    t = (rd << 5) + lane;
                    ok = t < T && (t < r0 || t >= r1);
                }
                if (ok && phase > 0 && P.cull) {
                    float smax;
                    const float d = knn_box_dist(tmin, tmax, t, qx, qy, qz, smax);
                    ok = (phase == 1 || (phase == 2 ? d <= near_d : d > near_d)) && knn_box_pass<MODE>(w, d, qq, smax);
                }
                unsigned m = __ballot_sync(0xffffffffu, ok);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    knn_open_tile<MODE>(w, ssort, __shfl_sync(0xffffffffu, t, b), qx, qy, qz, qq);
                    if (w.cnt > KB - 32) knn_merge(w);
                }
            }
        }
        if (w.cnt) knn_merge(w);
    }

    int64_t* out = J.out + ((size_t)frame * Q.n + qi) * J.k;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const u64 key = w.best.v[r];
        if (lane * 4 + r < J.k) out[lane * 4 + r] = key == KMAX ? (int64_t)ns : (int64_t)(unsigned)key;  // shadow index
    }
}

inline int next_pow2(int64_t n) {
    int p = 32;
    while (p < n) p <<= 1;
    return p;
}
inline int64_t align256(int64_t b) { return (b + 255) / 256 * 256; }

int64_t set_bytes(int64_t n, int frames) {
    const int64_t P = next_pow2(n), npad = (n + 31) / 32 * 32;
    const int64_t nsup = (npad / 32 + 31) / 32;
    return align256(frames * P * 8) + align256(frames * npad * 16) + align256(frames * npad * 4) +
           2 * align256(frames * (npad / 32) * 16) + 2 * align256(frames * nsup * 16) + align256((int64_t)frames * 16);
}

char* carve_set(SetDesc& S, const float* pts, int64_t n, int frames, char* w) {
    S.pts = pts;
    S.n = (int)n;
    S.P = next_pow2(n);
    S.npad = (int)((n + 31) / 32 * 32);
    S.keys = (unsigned long long*)w;
    w += align256((int64_t)frames * S.P * 8);
    S.sorted = (float4*)w;
    w += align256((int64_t)frames * S.npad * 16);
    S.codes = (uint32_t*)w;
    w += align256((int64_t)frames * S.npad * 4);
    S.tmin = (float4*)w;
    w += align256((int64_t)frames * (S.npad / 32) * 16);
    S.tmax = (float4*)w;
    w += align256((int64_t)frames * (S.npad / 32) * 16);
    S.nsup = (S.npad / 32 + 31) / 32;
    S.smin = (float4*)w;
    w += align256((int64_t)frames * S.nsup * 16);
    S.smax = (float4*)w;
    w += align256((int64_t)frames * S.nsup * 16);
    S.info = (float4*)w;
    w += align256((int64_t)frames * 16);
    return w;
}

int run_sort(QueryParams& Q, int nsets, int frames, cudaStream_t st) {
    static bool attr_done = false;
    const int smem = SORT_CHUNK * 8;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(knn_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("cofi_knn: cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
            return COFI_ECUDA;
        }
        attr_done = true;
    }
    SortParams SP;
    SP.nsets = nsets;
    int chunks = 0;
    for (int s = 0; s < nsets; ++s) {
        SP.set[s] = Q.set[s];
        SP.chunk_begin[s] = chunks;
        chunks += Q.set[s].P <= SORT_CHUNK_A ? 1 : Q.set[s].P / SORT_CHUNK_A;
    }
    SP.chunk_begin[nsets] = chunks;
    knn_sort_chunks_kernel<<<dim3(chunks, frames), SORT_THREADS, 0, st>>>(SP);
    int rc = check_launch("cofi_knn: chunk sort");
    if (rc) return rc;
    knn_sort_kernel<<<dim3(nsets, frames), SORT_THREADS, smem, st>>>(SP);
    return check_launch("cofi_knn: sort");
}

int run_query(QueryParams& Q, int frames, int mode, cudaStream_t st) {
    if (Q.njobs == 0) return COFI_OK;
    int blocks = 0;
    for (int j = 0; j < Q.njobs; ++j) {
        Q.job[j].block_begin = blocks;
        blocks += (int)ceil_div(Q.set[Q.job[j].qry].n, QWARPS);
    }
    if ((mode & 0xff) == COFI_KNN_DIRECT)
        knn_query_kernel<COFI_KNN_DIRECT><<<dim3(blocks, frames), QWARPS * 32, 0, st>>>(Q);
    else
        knn_query_kernel<COFI_KNN_EXPANDED><<<dim3(blocks, frames), QWARPS * 32, 0, st>>>(Q);
    return check_launch("cofi_knn: query");
}

int run(QueryParams& Q, int nsets, int frames, int mode, cudaStream_t st) {
    int rc = run_sort(Q, nsets, frames, st);
    if (rc) return rc;
    return run_query(Q, frames, mode, st);
}

}  // namespace
}  // namespace cofi

using namespace cofi;

extern "C" int64_t cofi_knn_pyramid_workspace(const int64_t* n_per_level, int levels, int frames) {
    if (!n_per_level || levels < 1 || levels > MAX_SETS || frames < 1) return -1;
    int64_t b = 0;
    for (int l = 0; l < levels; ++l) b += set_bytes(n_per_level[l], frames);
    return b;
}

extern "C" int cofi_knn_pyramid(const float* const* points, const int64_t* n_per_level, int levels, int frames, int k,
                                int k_up, int mode, int64_t* const* neighbors, int64_t* const* subsampling,
                                int64_t* const* upsampling, void* workspace, void* stream) {
    COFI_REQUIRE(points && n_per_level && workspace && levels >= 1 && levels <= MAX_SETS && frames >= 1 && frames <= 65535,
                 "cofi_knn_pyramid: bad argument");
    COFI_REQUIRE(k >= 1 && k <= KB && k_up >= 1 && k_up <= KB, "cofi_knn_pyramid: k and k_up must be in 1..128");
    COFI_REQUIRE((mode & 0xff) == COFI_KNN_DIRECT || (mode & 0xff) == COFI_KNN_EXPANDED, "cofi_knn_pyramid: bad mode");
    QueryParams Q;
    char* w = (char*)workspace;
    for (int l = 0; l < levels; ++l) {
        COFI_REQUIRE(points[l] && n_per_level[l] >= 1 && n_per_level[l] <= (1 << 20),
                     "cofi_knn_pyramid: level %d must hold 1..2^20 points per frame", l);
        w = carve_set(Q.set[l], points[l], n_per_level[l], frames, w);
    }
    Q.njobs = 0;
    Q.cull = (mode & COFI_KNN_NOCULL) ? 0 : 1;
    Q.fast = (mode & (COFI_KNN_NOCULL | COFI_KNN_NOFAST)) ? 0 : 1;
    // Two query launches.  First: the same-level tables and the up-sampling tables.  Second: the sub-sampling tables, whose
    // queries (level l+1) are points of level l in every pyramid built by half-sampling -- they copy rows of the finished
    // `neighbors[l]` (a query without a twin in level l is searched as before, so arbitrary pyramids stay correct).
    QueryParams Q2 = Q;
    auto add = [&](QueryParams& T, int src, int qry, int64_t* out, int kk, const int64_t* copy_from) {
        if (!out) return;
        JobDesc& J = T.job[T.njobs++];
        J.src = src;
        J.qry = qry;
        J.out = out;
        J.copy_from = copy_from;
        J.block_begin = 0;
        J.k = kk;
    };
    for (int l = 0; l < levels; ++l) {
        if (neighbors) add(Q, l, l, neighbors[l], k, nullptr);
        if (l + 1 < levels) {
            if (upsampling) add(Q, l + 1, l, upsampling[l], k_up, nullptr);  // level l points look up level l+1
            if (subsampling) {                                                // level l+1 points look up level l
                const bool can_copy = Q.fast && neighbors && neighbors[l];
                add(can_copy ? Q2 : Q, l, l + 1, subsampling[l], k, can_copy ? neighbors[l] : nullptr);
            }
        }
    }
    if (Q.njobs == 0 && Q2.njobs == 0) return COFI_OK;
    int rc = run_sort(Q, levels, frames, (cudaStream_t)stream);
    if (rc) return rc;
    rc = run_query(Q, frames, mode, (cudaStream_t)stream);
    if (rc) return rc;
    return run_query(Q2, frames, mode, (cudaStream_t)stream);
}

extern "C" int64_t cofi_knn_table_workspace(int64_t ns, int64_t nq, int frames) {
    if (ns < 1 || nq < 1 || frames < 1) return -1;
    return set_bytes(ns, frames) + set_bytes(nq, frames);
}

extern "C" int cofi_knn_table(const float* src, int64_t ns, const float* qry, int64_t nq, int frames, int k, int mode,
                              int64_t* out, void* workspace, void* stream) {
    COFI_REQUIRE(src && qry && out && workspace && ns >= 1 && nq >= 1 && ns <= (1 << 20) && nq <= (1 << 20) &&
                     frames >= 1 && frames <= 65535,
                 "cofi_knn_table: bad argument");
    COFI_REQUIRE(k >= 1 && k <= KB, "cofi_knn_table: k must be in 1..128");
    COFI_REQUIRE((mode & 0xff) == COFI_KNN_DIRECT || (mode & 0xff) == COFI_KNN_EXPANDED, "cofi_knn_table: bad mode");
    QueryParams Q;
    char* w = carve_set(Q.set[0], src, ns, frames, (char*)workspace);
    const bool same = (src == qry && ns == nq);
    if (!same) carve_set(Q.set[1], qry, nq, frames, w);
    Q.njobs = 1;
    Q.cull = (mode & COFI_KNN_NOCULL) ? 0 : 1;
    Q.fast = (mode & (COFI_KNN_NOCULL | COFI_KNN_NOFAST)) ? 0 : 1;
    Q.job[0].src = 0;
    Q.job[0].qry = same ? 0 : 1;
    Q.job[0].out = out;
    Q.job[0].copy_from = nullptr;
    Q.job[0].block_begin = 0;
    Q.job[0].k = k;
    return run(Q, same ? 1 : 2, frames, mode, (cudaStream_t)stream);
}
