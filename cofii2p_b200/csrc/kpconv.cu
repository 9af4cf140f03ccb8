// kpconv.cu -- point-stream gather kernels (reference model/kpconv/kpconv.py:79-122, functional.py).
//
// KPConv on this workload is *sparse*: a 128-NN patch of a 20480-point KITTI cloud spans metres while a
// kernel point only reaches sigma = 0.2..3.2 m, so most of the 128x15 influences are exactly zero.
// The aggregate kernel therefore never forms the (M,128,15,3) tensor the reference materialises: one warp
// owns one query point, keeps the 128 relative neighbour positions in registers (4 per lane, one 16-byte
// gather each from the packed (x,y,z,flag) table), and for each kernel point ballots the lanes whose
// influence is non-zero; only those neighbours' feature rows are loaded (coalesced, float4 per lane) and
// accumulated with warp-uniform control flow.  No tensor cores: the contraction is data-dependent sparse.
#include <cuda_fp16.h>

#include "common.cuh"

namespace cofi {

constexpr int kMaxKP = 32;

// ------------------------------------------------------------------------------------------- pack_points
__global__ void __launch_bounds__(256) pack_points_kernel(const float* __restrict__ pts,
                                                          const float* __restrict__ feats, int64_t ldf, int C,
                                                          int64_t rows, float4* __restrict__ packed) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* f = feats + row * ldf;
    // fp64 accumulation: the sign of the row sum decides membership in KPConv's neighbour count
    double s = 0.0;
    for (int c = lane; c < C; c += 32) s += (double)__ldg(f + c);
    s = warp_sum_d(s);
    if (lane == 0) {
        const float* p = pts + row * 3;
        packed[row] = make_float4(p[0], p[1], p[2], s > 0.0 ? 1.0f : 0.0f);
    }
}

// ------------------------------------------------------------------------------------ kpconv_aggregate
template <int VEC>
struct VecT;
template <>
struct VecT<1> {
    using T = float;
};
template <>
struct VecT<2> {
    using T = float2;
};
template <>
struct VecT<4> {
    using T = float4;
};

// One warp per query point.
//   1. each lane loads 4 of the <=128 neighbours (one 16-byte gather of (x,y,z,flag) each), forms the relative position
//      and culls exactly: |rel| > max|kp| + sigma can not be influenced by any kernel point;
//   2. surviving ("near") neighbours are compacted in ascending-h order into a per-warp shared-memory list;
//   3. the n_near x K (neighbour, kernel point) pairs are enumerated kernel-point-major and dealt 32 at a time to the
//      lanes, so the sqrt/div influence evaluations are spread over the warp instead of 60 per lane; a ballot keeps the
//      non-zero influences, which are consumed in order (same accumulation order as a dense h loop): broadcast
//      (row, weight), every lane loads its float4 slice of that feature row and accumulates; when the kernel point
//      changes the finished [C] row (or zeros) is stored.
template <int VEC, int NCH, typename OutT = float>
__global__ void __launch_bounds__(128)
kpconv_aggregate_kernel(const float* __restrict__ feats, int64_t ldf, int C, const float4* __restrict__ s_packed,
                        const float* __restrict__ q_points, const int64_t* __restrict__ nbr, int H, int64_t Mq,
                        int64_t Ns, int64_t total_q, const float* __restrict__ kernel_points, int K, float sigma,
                        float reach2, OutT* __restrict__ agg, float* __restrict__ cnt_out) {
    __shared__ float skp[kMaxKP * 3];
    __shared__ float4 snear[4][128];  // per warp: (rx, ry, rz, __int_as_float(row index))
    if (threadIdx.x < K * 3) skp[threadIdx.x] = kernel_points[threadIdx.x];
    __syncthreads();

    const int wib = threadIdx.x >> 5;
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    const int lane = threadIdx.x & 31;
    if (m >= total_q) return;
    const int64_t frame = m / Mq;
    const float* fbase = feats + frame * Ns * ldf;
    const float4* sp = s_packed + frame * Ns;
    const float qx = __ldg(q_points + m * 3 + 0), qy = __ldg(q_points + m * 3 + 1), qz = __ldg(q_points + m * 3 + 2);
    float4* near = snear[wib];

    float cnt = 0.0f;
    int n_near = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int h = j * 32 + lane;
        const int64_t id = (h < H) ? __ldcs(nbr + m * H + h) : Ns;
        bool is_near = false;
        float rx = 0.f, ry = 0.f, rz = 0.f;
        if (id >= 0 && id < Ns) {
            const float4 p = __ldg(sp + id);
            rx = p.x - qx;  // neighbours - q_points, reference kpconv.py:93
            ry = p.y - qy;
            rz = p.z - qz;
            cnt += p.w;
            is_near = rx * rx + ry * ry + rz * rz <= reach2;
        }  // else shadow neighbour: point at 1e6, zero feature -> zero influence
        const unsigned bal = __ballot_sync(0xffffffffu, is_near);
        if (is_near) near[n_near + __popc(bal & ((1u << lane) - 1u))] = make_float4(rx, ry, rz, __int_as_float((int)id));
        n_near += __popc(bal);
    }
    __syncwarp();
    cnt = warp_sum(cnt);
    if (lane == 0) cnt_out[m] = fmaxf(cnt, 1.0f);

    using V = typename VecT<VEC>::T;
    OutT* out_row = agg + m * (int64_t)K * C;
    const bool lane_active = (lane * VEC) < C;  // only matters for C < 32*VEC (C = 4)

    float acc[NCH][VEC];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[c][v] = 0.0f;

    auto store_row = [&](int k) {  // writes acc as row k and clears it
        if (lane_active) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                OutT* dst = out_row + (int64_t)k * C + (c * 32 + lane) * VEC;
                if constexpr (sizeof(OutT) == 4) {
                    V o;
                    float* ov = reinterpret_cast<float*>(&o);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) ov[v] = acc[c][v];
                    *reinterpret_cast<V*>(dst) = o;
                } else if constexpr (VEC == 4) {
                    __half2 lo = __floats2half2_rn(acc[c][0], acc[c][1]), hi = __floats2half2_rn(acc[c][2], acc[c][3]);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(dst) = pk;
                } else if constexpr (VEC == 2) {
                    *reinterpret_cast<__half2*>(dst) = __floats2half2_rn(acc[c][0], acc[c][1]);
                } else {
                    dst[0] = __float2half_rn(acc[c][0]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[c][v] = 0.0f;
    };

    int cur_k = 0;  // kernel point whose row is being accumulated
    const int total_pairs = n_near * K;
    const float inv_near = 1.0f / (float)(n_near > 0 ? n_near : 1);
    for (int base = 0; base < total_pairs; base += 32) {
        const int pidx = base + lane;
        float w = 0.0f;
        int pk = 0, prow = 0;
        if (pidx < total_pairs) {
            // pidx / n_near without the integer-division sequence: reciprocal estimate + one correction (pidx < 4096)
            pk = (int)((float)pidx * inv_near);
            if (pk * n_near > pidx) --pk;
            else if ((pk + 1) * n_near <= pidx) ++pk;
            const float4 nb = near[pidx - pk * n_near];
            prow = __float_as_int(nb.w);
            // differences = neighbours - kernel_points; sq = sum(d^2); w = clamp(1 - sqrt(sq)/sigma, 0)
            // (reference kpconv.py:97-99), evaluated without FMA contraction
            const float dx = nb.x - skp[pk * 3 + 0], dy = nb.y - skp[pk * 3 + 1], dz = nb.z - skp[pk * 3 + 2];
            const float sq = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            w = fmaxf(__fsub_rn(1.0f, __fdiv_rn(__fsqrt_rn(sq), sigma)), 0.0f);
        }
        unsigned mask = __ballot_sync(0xffffffffu, w > 0.0f);
        // The influencing pairs are consumed in order (the accumulation order of a dense h loop), but their feature rows are
        // fetched UNR at a time: the gathers are the latency of this kernel (L2 round trips), the FMAs are not.
        constexpr int UNR = (NCH * VEC >= 16) ? 2 : 4;
        while (mask) {
            int ek[UNR], er[UNR];
            float ew[UNR];
            V f[UNR][NCH];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const bool valid = mask != 0u;
                const int l = valid ? __ffs(mask) - 1 : 0;
                mask &= mask - 1;   // 0 stays 0
                ek[u] = valid ? __shfl_sync(0xffffffffu, pk, l) : -1;
                er[u] = __shfl_sync(0xffffffffu, prow, l);
                ew[u] = __shfl_sync(0xffffffffu, w, l);
                if (valid && lane_active) {
                    const float* row = fbase + (int64_t)er[u] * ldf;
#pragma unroll
                    for (int c = 0; c < NCH; ++c) f[u][c] = __ldg(reinterpret_cast<const V*>(row + (c * 32 + lane) * VEC));
                }
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                if (ek[u] < 0) break;   // warp-uniform
                while (cur_k < ek[u]) store_row(cur_k++);  // finished rows (zeros when nothing influenced them)
                if (lane_active) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const float* fv = reinterpret_cast<const float*>(&f[u][c]);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) acc[c][v] = fmaf(ew[u], fv[v], acc[c][v]);
                    }
                }
            }
        }
    }
    while (cur_k < K) store_row(cur_k++);
}

// ------------------------------------------------------------------------------------------ maxpool_rows
// 4 channels (16 B) per lane; rows narrower than a warp (C < 128) are walked by 32 / (C/4) lane groups on interleaved
// neighbours and folded with shuffles (see the fp16 variant below).
__global__ void __launch_bounds__(128)
maxpool_rows_kernel(const float* __restrict__ x, int64_t ldx, int C, const int64_t* __restrict__ nbr, int H,
                    int64_t Mq, int64_t Ns, int64_t total_q, float* __restrict__ out, int64_t ldo) {
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= total_q) return;
    const int64_t frame = m / Mq;
    const float* xb = x + frame * Ns * ldx;
    int idx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int h = j * 32 + lane;
        int64_t id = (h < H) ? __ldcs(nbr + m * H + h) : -2;  // -2: beyond H (ignored), -1: shadow (value 0)
        if (id >= Ns) id = -1;
        idx[j] = (int)id;
    }
    const bool vec = (C % 4) == 0;
    const int rl = C >> 2;
    const int lpr = (vec && (rl == 4 || rl == 8 || rl == 16)) ? rl : 32;
    const int groups = 32 / lpr, grp = lane / lpr, gl = lane - grp * lpr;
    for (int c0 = 0; c0 < C; c0 += 128) {
        const int c = c0 + gl * 4;
        const bool act = c < C;
        float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            for (int l = 0; l < 32; l += 4 * groups) {  // four independent gathers in flight per lane
                int id[4];
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) id[u] = __shfl_sync(0xffffffffu, idx[j], l + u * groups + grp);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (id[u] >= 0 && act) {
                        const float* r = xb + (int64_t)id[u] * ldx;
                        if (vec) {
                            v[u] = __ldg(reinterpret_cast<const float4*>(r + c));
                        } else {
                            v[u].x = r[c];
                            if (c + 1 < C) v[u].y = r[c + 1];
                            if (c + 2 < C) v[u].z = r[c + 2];
                            if (c + 3 < C) v[u].w = r[c + 3];
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (id[u] == -2) continue;  // beyond H: not a neighbour at all
                    mx.x = fmaxf(mx.x, v[u].x);
                    mx.y = fmaxf(mx.y, v[u].y);
                    mx.z = fmaxf(mx.z, v[u].z);
                    mx.w = fmaxf(mx.w, v[u].w);
                }
            }
        }
        for (int off = lpr; off < 32; off <<= 1) {
            mx.x = fmaxf(mx.x, __shfl_xor_sync(0xffffffffu, mx.x, off));
            mx.y = fmaxf(mx.y, __shfl_xor_sync(0xffffffffu, mx.y, off));
            mx.z = fmaxf(mx.z, __shfl_xor_sync(0xffffffffu, mx.z, off));
            mx.w = fmaxf(mx.w, __shfl_xor_sync(0xffffffffu, mx.w, off));
        }
        if (act && grp == 0) {
            float* o = out + m * ldo + c;
            if (vec) {
                *reinterpret_cast<float4*>(o) = mx;
            } else {
                o[0] = mx.x;
                if (c + 1 < C) o[1] = mx.y;
                if (c + 2 < C) o[2] = mx.z;
                if (c + 3 < C) o[3] = mx.w;
            }
        }
    }
}

// Instruction-lean variant for the shapes of the model (C = 64, 128, 256, 512; 16-byte aligned rows).  ncu showed the generic
// kernel above issue-bound, not bandwidth-bound (77 % issue slots busy, 67 instructions per gathered 16 bytes: 64-bit address
// arithmetic, three validity branches and a vector / scalar branch per load).  Here: compile-time lane groups, 32-bit offsets
// in float4 units, slots beyond H replaced by the first neighbour (a duplicate cannot change a maximum), shadow neighbours
// as a predicated load of zeros -- about 10 instructions per gather.
template <int LPR>   // lanes per row group: 16 (C = 64) or 32 (C a multiple of 128)
__global__ void __launch_bounds__(128)
maxpool_rows_fast_kernel(const float4* __restrict__ x4, int ldx4, int C4, const int64_t* __restrict__ nbr, int H, int64_t Mq,
                         int64_t Ns, int64_t total_q, float4* __restrict__ out4, int ldo4) {
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= total_q) return;
    const float4* xb = x4 + (m / Mq) * Ns * ldx4;
    int idx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int h = j * 32 + lane;
        int64_t id = (h < H) ? __ldcs(nbr + m * H + h) : -2;
        if (id >= Ns) id = -1;                      // shadow neighbour: a row of zeros
        idx[j] = (int)id;
    }
    const int first = __shfl_sync(0xffffffffu, idx[0], 0);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (idx[j] == -2) idx[j] = first;           // beyond H: repeat neighbour 0
    constexpr int GROUPS = 32 / LPR;
    const int grp = lane / LPR, gl = lane - grp * LPR;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c4 = gl; c4 < C4; c4 += LPR) {
        float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int l = 0; l < 32; l += 4 * GROUPS) {  // four independent gathers in flight per lane
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int id = __shfl_sync(0xffffffffu, idx[j], l + u * GROUPS + grp);
                    v[u] = zero4;
                    if (id >= 0) v[u] = __ldg(xb + id * ldx4 + c4);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    mx.x = fmaxf(mx.x, v[u].x);
                    mx.y = fmaxf(mx.y, v[u].y);
                    mx.z = fmaxf(mx.z, v[u].z);
                    mx.w = fmaxf(mx.w, v[u].w);
                }
            }
        }
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1) {
            mx.x = fmaxf(mx.x, __shfl_xor_sync(0xffffffffu, mx.x, off));
            mx.y = fmaxf(mx.y, __shfl_xor_sync(0xffffffffu, mx.y, off));
            mx.z = fmaxf(mx.z, __shfl_xor_sync(0xffffffffu, mx.z, off));
            mx.w = fmaxf(mx.w, __shfl_xor_sync(0xffffffffu, mx.w, off));
        }
        if (grp == 0) out4[m * ldo4 + c4] = mx;
    }
}

// fp16-input variant (tf32 engine): rounding is monotonic, so max over fp16-rounded rows == fp16-rounded max; the
// gather moves half the bytes.  8 channels (16 B) per lane; when a row needs fewer than 32 lanes (C < 256) the warp
// splits into 32 / (C/8) groups that walk interleaved neighbours and are folded with shuffles at the end, so all lanes
// gather (C = 64: four neighbours per step instead of one with 24 idle lanes).
__global__ void __launch_bounds__(128)
maxpool_rows_f16_kernel(const __half* __restrict__ x, int64_t ldx, int C, const int64_t* __restrict__ nbr, int H,
                        int64_t Mq, int64_t Ns, int64_t total_q, float* __restrict__ out, int64_t ldo) {
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= total_q) return;
    const int64_t frame = m / Mq;
    const __half* xb = x + frame * Ns * ldx;
    int idx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int h = j * 32 + lane;
        int64_t id = (h < H) ? __ldcs(nbr + m * H + h) : -2;
        if (id >= Ns) id = -1;
        idx[j] = (int)id;
    }
    const int rl = C >> 3;                                                       // lanes one row needs
    const int lpr = (rl == 4 || rl == 8 || rl == 16) ? rl : 32;                  // lanes per row group
    const int groups = 32 / lpr, grp = lane / lpr, gl = lane - grp * lpr;
    for (int c0 = 0; c0 < C; c0 += 256) {
        const int c = c0 + gl * 8;
        const bool act = c < C;
        __half2 mx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) mx[i] = __float2half2_rn(-65504.0f);
        const __half2 zero2 = __float2half2_rn(0.0f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            for (int l = 0; l < 32; l += 4 * groups) {  // four independent gathers in flight per lane
                int id[4];
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) id[u] = __shfl_sync(0xffffffffu, idx[j], l + u * groups + grp);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    v[u] = make_uint4(0u, 0u, 0u, 0u);  // shadow row = zeros
                    if (id[u] >= 0 && act) v[u] = __ldg(reinterpret_cast<const uint4*>(xb + (int64_t)id[u] * ldx + c));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (id[u] == -2) continue;
                    const __half2* hv = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) mx[i] = __hmax2(mx[i], id[u] >= 0 ? hv[i] : zero2);
                }
            }
        }
        for (int off = lpr; off < 32; off <<= 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const unsigned o = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<unsigned*>(&mx[i]), off);
                mx[i] = __hmax2(mx[i], *reinterpret_cast<const __half2*>(&o));
            }
        }
        if (act && grp == 0) {
            float* o = out + m * ldo + c;
            const float2 a = __half22float2(mx[0]), b = __half22float2(mx[1]), cc = __half22float2(mx[2]), d = __half22float2(mx[3]);
            *reinterpret_cast<float4*>(o) = make_float4(a.x, a.y, b.x, b.y);
            *reinterpret_cast<float4*>(o + 4) = make_float4(cc.x, cc.y, d.x, d.y);
        }
    }
}

// ------------------------------------------------------------------------------------------- gather_rows
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ x, int64_t ldx, int C, const int64_t* __restrict__ idx,
                   int64_t idx_stride, int64_t Mq, int64_t Ns, int64_t total_q, float* __restrict__ out,
                   int64_t ldo, int vec4) {
    const int per_row = vec4 ? (C >> 2) : C;
    const int64_t total = total_q * per_row;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = t / per_row;
        const int c = (int)(t - m * per_row);
        const int64_t frame = m / Mq;
        int64_t src = m - frame * Mq;  // identity when idx == NULL (then Mq == Ns)
        if (idx != nullptr) src = __ldg(idx + m * idx_stride);
        const bool ok = src >= 0 && src < Ns;
        const float* r = x + (frame * Ns + (ok ? src : 0)) * ldx;
        if (vec4) {
            float4 v = ok ? __ldg(reinterpret_cast<const float4*>(r) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            reinterpret_cast<float4*>(out + m * ldo)[c] = v;
        } else {
            out[m * ldo + c] = ok ? __ldg(r + c) : 0.0f;
        }
    }
}

// instruction-lean fp16 variant (see maxpool_rows_fast_kernel): 8 channels (16 bytes) per lane
template <int LPR>   // lanes per row group: 8 (C = 64), 16 (C = 128), 32 (C a multiple of 256)
__global__ void __launch_bounds__(128)
maxpool_rows_f16_fast_kernel(const uint4* __restrict__ x8, int ldx8, int C8, const int64_t* __restrict__ nbr, int H, int64_t Mq,
                             int64_t Ns, int64_t total_q, float4* __restrict__ out4, int ldo4) {
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= total_q) return;
    const uint4* xb = x8 + (m / Mq) * Ns * ldx8;
    int idx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int h = j * 32 + lane;
        int64_t id = (h < H) ? __ldcs(nbr + m * H + h) : -2;
        if (id >= Ns) id = -1;
        idx[j] = (int)id;
    }
    const int first = __shfl_sync(0xffffffffu, idx[0], 0);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (idx[j] == -2) idx[j] = first;
    constexpr int GROUPS = 32 / LPR;
    const int grp = lane / LPR, gl = lane - grp * LPR;
    for (int c8 = gl; c8 < C8; c8 += LPR) {
        __half2 mx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) mx[i] = __float2half2_rn(-65504.0f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int l = 0; l < 32; l += 4 * GROUPS) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int id = __shfl_sync(0xffffffffu, idx[j], l + u * GROUPS + grp);
                    v[u] = make_uint4(0u, 0u, 0u, 0u);   // shadow row = zeros
                    if (id >= 0) v[u] = __ldg(xb + id * ldx8 + c8);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const __half2* hv = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) mx[i] = __hmax2(mx[i], hv[i]);
                }
            }
        }
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const unsigned o = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<unsigned*>(&mx[i]), off);
                mx[i] = __hmax2(mx[i], *reinterpret_cast<const __half2*>(&o));
            }
        }
        if (grp == 0) {
            const float2 a = __half22float2(mx[0]), b = __half22float2(mx[1]), c = __half22float2(mx[2]), d = __half22float2(mx[3]);
            out4[m * ldo4 + c8 * 2] = make_float4(a.x, a.y, b.x, b.y);
            out4[m * ldo4 + c8 * 2 + 1] = make_float4(c.x, c.y, d.x, d.y);
        }
    }
}

}  // namespace cofi

using namespace cofi;

extern "C" int cofi_pack_points(const float* points, const float* feats, int64_t ldf, int C, int64_t rows,
                                float* packed, void* stream) {
    COFI_REQUIRE(points && feats && packed, "cofi_pack_points: null pointer");
    COFI_REQUIRE(C > 0 && rows >= 0 && ldf >= C, "cofi_pack_points: bad shape C=%d rows=%lld ldf=%lld", C,
                 (long long)rows, (long long)ldf);
    if (rows == 0) return COFI_OK;
    const int wpb = 8;
    pack_points_kernel<<<(unsigned)ceil_div(rows, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        points, feats, ldf, C, rows, reinterpret_cast<float4*>(packed));
    return check_launch("cofi_pack_points");
}

extern "C" int cofi_kpconv_aggregate(const float* feats, int64_t ldf, int C, const float* s_packed,
                                     const float* q_points, const int64_t* nbr, int H, int64_t Mq, int64_t Ns,
                                     int frames, const float* kernel_points, int K, float sigma, float kp_reach,
                                     float* agg, float* cnt, void* stream) {
    COFI_REQUIRE(feats && s_packed && q_points && nbr && kernel_points && agg && cnt,
                 "cofi_kpconv_aggregate: null pointer");
    COFI_REQUIRE(H > 0 && H <= 128, "cofi_kpconv_aggregate: H=%d must be in 1..128", H);
    COFI_REQUIRE(K > 0 && K <= kMaxKP, "cofi_kpconv_aggregate: K=%d must be in 1..%d", K, kMaxKP);
    COFI_REQUIRE(sigma > 0.0f, "cofi_kpconv_aggregate: sigma must be positive");
    COFI_REQUIRE(Mq >= 0 && Ns > 0 && frames > 0 && ldf >= C, "cofi_kpconv_aggregate: bad sizes");
    const int64_t total = Mq * frames;
    if (total == 0) return COFI_OK;
    // cull radius: (max|kp| + sigma) with a 1e-3 relative margin; kp_reach <= 0 disables the cull
    const float reach = kp_reach > 0.0f ? (kp_reach + sigma) * 1.001f : 1e18f;
    const float reach2 = reach * reach;
    const int wpb = 4;
    const dim3 grid((unsigned)ceil_div(total, wpb)), block(wpb * 32);
    cudaStream_t st = (cudaStream_t)stream;
    const float4* sp = reinterpret_cast<const float4*>(s_packed);
#define LAUNCH(VEC, NCH)                                                                                          \
    kpconv_aggregate_kernel<VEC, NCH><<<grid, block, 0, st>>>(feats, ldf, C, sp, q_points, nbr, H, Mq, Ns, total, \
                                                              kernel_points, K, sigma, reach2, agg, cnt)
    if (C <= 32) {
        LAUNCH(1, 1);
    } else if (C == 64) {
        COFI_REQUIRE(ldf % 2 == 0, "cofi_kpconv_aggregate: ldf must be even for C=64");
        LAUNCH(2, 1);
    } else if (C % 128 == 0 && C <= 1024) {
        COFI_REQUIRE(ldf % 4 == 0, "cofi_kpconv_aggregate: ldf must be a multiple of 4");
        switch (C / 128) {
            case 1: LAUNCH(4, 1); break;
            case 2: LAUNCH(4, 2); break;
            case 3: LAUNCH(4, 3); break;
            case 4: LAUNCH(4, 4); break;
            case 8: LAUNCH(4, 8); break;
            default:
                set_error("cofi_kpconv_aggregate: unsupported channel count C=%d", C);
                return COFI_EUNSUPPORTED;
        }
    } else {
        set_error("cofi_kpconv_aggregate: unsupported channel count C=%d", C);
        return COFI_EUNSUPPORTED;
    }
#undef LAUNCH
    return check_launch("cofi_kpconv_aggregate");
}

extern "C" int cofi_kpconv_aggregate_f16(const float* feats, int64_t ldf, int C, const float* s_packed,
                                         const float* q_points, const int64_t* nbr, int H, int64_t Mq, int64_t Ns,
                                         int frames, const float* kernel_points, int K, float sigma, float kp_reach,
                                         void* agg_f16, float* cnt, void* stream) {
    COFI_REQUIRE(feats && s_packed && q_points && nbr && kernel_points && agg_f16 && cnt,
                 "cofi_kpconv_aggregate_f16: null pointer");
    COFI_REQUIRE(H > 0 && H <= 128 && K > 0 && K <= kMaxKP && sigma > 0.0f, "cofi_kpconv_aggregate_f16: bad argument");
    COFI_REQUIRE(Mq >= 0 && Ns > 0 && frames > 0 && ldf >= C, "cofi_kpconv_aggregate_f16: bad sizes");
    const int64_t total = Mq * frames;
    if (total == 0) return COFI_OK;
    const float reach = kp_reach > 0.0f ? (kp_reach + sigma) * 1.001f : 1e18f;
    const float reach2 = reach * reach;
    const int wpb = 4;
    const dim3 grid((unsigned)ceil_div(total, wpb)), block(wpb * 32);
    cudaStream_t st = (cudaStream_t)stream;
    const float4* sp = reinterpret_cast<const float4*>(s_packed);
    __half* agg = reinterpret_cast<__half*>(agg_f16);
#define LAUNCH(VEC, NCH)                                                                                         \
    kpconv_aggregate_kernel<VEC, NCH, __half><<<grid, block, 0, st>>>(feats, ldf, C, sp, q_points, nbr, H, Mq, Ns, \
                                                                      total, kernel_points, K, sigma, reach2, agg, cnt)
    if (C <= 32) {
        LAUNCH(1, 1);
    } else if (C == 64) {
        LAUNCH(2, 1);
    } else if (C % 128 == 0 && C <= 1024 && ldf % 4 == 0) {
        switch (C / 128) {
            case 1: LAUNCH(4, 1); break;
            case 2: LAUNCH(4, 2); break;
            case 3: LAUNCH(4, 3); break;
            case 4: LAUNCH(4, 4); break;
            case 8: LAUNCH(4, 8); break;
            default:
                set_error("cofi_kpconv_aggregate_f16: unsupported channel count C=%d", C);
                return COFI_EUNSUPPORTED;
        }
    } else {
        set_error("cofi_kpconv_aggregate_f16: unsupported channel count C=%d", C);
        return COFI_EUNSUPPORTED;
    }
#undef LAUNCH
    return check_launch("cofi_kpconv_aggregate_f16");
}

extern "C" int cofi_maxpool_rows(const float* x, int64_t ldx, int C, const int64_t* nbr, int H, int64_t Mq,
                                 int64_t Ns, int frames, float* out, int64_t ldo, void* stream) {
    COFI_REQUIRE(x && nbr && out, "cofi_maxpool_rows: null pointer");
    COFI_REQUIRE(H > 0 && H <= 128 && C > 0 && ldx >= C && ldo >= C, "cofi_maxpool_rows: bad shape");
    COFI_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0, "cofi_maxpool_rows: leading dimensions must be multiples of 4");
    const int64_t total = Mq * frames;
    if (total == 0) return COFI_OK;
    const int wpb = 4;
    if (C % 64 == 0 && (C == 64 || C % 128 == 0) && ldx % 4 == 0 && ldo % 4 == 0 && ((uintptr_t)x % 16) == 0 &&
        ((uintptr_t)out % 16) == 0 && Ns * ldx < (1ll << 31)) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        float4* o4 = reinterpret_cast<float4*>(out);
        if (C == 64)
            maxpool_rows_fast_kernel<16><<<(unsigned)ceil_div(total, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
                x4, (int)(ldx / 4), C / 4, nbr, H, Mq, Ns, total, o4, (int)(ldo / 4));
        else
            maxpool_rows_fast_kernel<32><<<(unsigned)ceil_div(total, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
                x4, (int)(ldx / 4), C / 4, nbr, H, Mq, Ns, total, o4, (int)(ldo / 4));
        return check_launch("cofi_maxpool_rows");
    }
    maxpool_rows_kernel<<<(unsigned)ceil_div(total, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        x, ldx, C, nbr, H, Mq, Ns, total, out, ldo);
    return check_launch("cofi_maxpool_rows");
}

extern "C" int cofi_maxpool_rows_f16(const void* x_f16, int64_t ldx, int C, const int64_t* nbr, int H, int64_t Mq,
                                     int64_t Ns, int frames, float* out, int64_t ldo, void* stream) {
    COFI_REQUIRE(x_f16 && nbr && out, "cofi_maxpool_rows_f16: null pointer");
    COFI_REQUIRE(H > 0 && H <= 128 && C > 0 && C % 8 == 0 && ldx % 8 == 0 && ldo % 4 == 0 && ldx >= C && ldo >= C,
                 "cofi_maxpool_rows_f16: C and ldx must be multiples of 8, ldo of 4");
    COFI_REQUIRE(((uintptr_t)x_f16 % 16) == 0 && ((uintptr_t)out % 16) == 0, "cofi_maxpool_rows_f16: 16-byte alignment");
    const int64_t total = Mq * frames;
    if (total == 0) return COFI_OK;
    const int wpb = 4;
    if ((C == 64 || C == 128 || C % 256 == 0) && Ns * ldx < (1ll << 31)) {
        const uint4* x8 = reinterpret_cast<const uint4*>(x_f16);
        float4* o4 = reinterpret_cast<float4*>(out);
        const unsigned blocks = (unsigned)ceil_div(total, wpb);
        if (C == 64)
            maxpool_rows_f16_fast_kernel<8><<<blocks, wpb * 32, 0, (cudaStream_t)stream>>>(x8, (int)(ldx / 8), C / 8, nbr, H, Mq, Ns, total, o4, (int)(ldo / 4));
        else if (C == 128)
            maxpool_rows_f16_fast_kernel<16><<<blocks, wpb * 32, 0, (cudaStream_t)stream>>>(x8, (int)(ldx / 8), C / 8, nbr, H, Mq, Ns, total, o4, (int)(ldo / 4));
        else
            maxpool_rows_f16_fast_kernel<32><<<blocks, wpb * 32, 0, (cudaStream_t)stream>>>(x8, (int)(ldx / 8), C / 8, nbr, H, Mq, Ns, total, o4, (int)(ldo / 4));
        return check_launch("cofi_maxpool_rows_f16");
    }
    maxpool_rows_f16_kernel<<<(unsigned)ceil_div(total, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __half*>(x_f16), ldx, C, nbr, H, Mq, Ns, total, out, ldo);
    return check_launch("cofi_maxpool_rows_f16");
}

extern "C" int cofi_gather_rows(const float* x, int64_t ldx, int C, const int64_t* idx, int64_t idx_stride,
                                int64_t Mq, int64_t Ns, int frames, float* out, int64_t ldo, void* stream) {
    COFI_REQUIRE(x && out, "cofi_gather_rows: null pointer");
    COFI_REQUIRE(C > 0 && ldx >= C && ldo >= C && frames > 0, "cofi_gather_rows: bad shape");
    COFI_REQUIRE(idx != nullptr || Mq == Ns, "cofi_gather_rows: identity copy needs Mq == Ns");
    const int64_t total = Mq * frames;
    if (total == 0) return COFI_OK;
    const int vec4 = (C % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && ((uintptr_t)x % 16 == 0) &&
                      ((uintptr_t)out % 16 == 0))
                         ? 1
                         : 0;
    const int64_t work = total * (vec4 ? C / 4 : C);
    const unsigned blocks = (unsigned)(ceil_div(work, 256) < 148 * 16 ? ceil_div(work, 256) : 148 * 16);
    gather_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, C, idx, idx_stride, Mq, Ns, total, out,
                                                                 ldo, vec4);
    return check_launch("cofi_gather_rows");
}
