// norm.cu -- normalisation kernels over row-major [rows, C] activations.
// GroupNorm over a cloud (reference model/kpconv/modules.py:45-49), affine-free InstanceNorm
// (model/imagenet.py:123, model/network.py:42-43), train-mode BatchNorm (model/imagenet.py:381-394) are the
// same computation here: statistics per (frame, group of channels) over the R rows of the frame.
// HBM-bound elementwise work: coalesced float loads along channels, fp64 statistics, two deterministic
// passes (partials -> apply), no atomics.
#include <stdlib.h>

#include <cuda_fp16.h>

#include "common.cuh"

namespace cofi {

constexpr int kStatChunks = 64;  // row chunks per frame for the partial statistics

// partials layout: [frames][kStatChunks][C][2] doubles (sum, sumsq)
__global__ void __launch_bounds__(256)
norm_stats_kernel(const float* __restrict__ x, int64_t ldx, int64_t R, int C, double* __restrict__ partials) {
    const int frame = blockIdx.z;
    const int chunk = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t rows_per = (R + kStatChunks - 1) / kStatChunks;
    const int64_t r0 = chunk * rows_per;
    const int64_t r1 = (r0 + rows_per < R) ? r0 + rows_per : R;
    if (c >= C) return;
    double s = 0.0, ss = 0.0;
    const float* p = x + ((int64_t)frame * R) * ldx + c;
    for (int64_t r = r0; r < r1; ++r) {
        const double v = (double)__ldg(p + r * ldx);
        s += v;
        ss += v * v;
    }
    double* o = partials + (((int64_t)frame * kStatChunks + chunk) * C + c) * 2;
    o[0] = s;
    o[1] = ss;
}

// one block per (frame, group): reduce partials -> mean, rstd
__global__ void __launch_bounds__(128)
norm_finalize_kernel(const double* __restrict__ partials, int nchunks, int64_t R, int C, int G, float eps,
                     float2* __restrict__ mean_rstd, float* __restrict__ mean_out, float* __restrict__ var_out) {
    const int frame = blockIdx.y, g = blockIdx.x;
    const int gs = C / G;
    double s = 0.0, ss = 0.0;
    for (int t = threadIdx.x; t < nchunks * gs; t += blockDim.x) {
        const int chunk = t / gs, c = g * gs + (t - chunk * gs);
        const double* o = partials + (((int64_t)frame * kStatChunks + chunk) * C + c) * 2;
        s += o[0];
        ss += o[1];
    }
    __shared__ double sh[2][4];
    s = warp_sum_d(s);
    ss = warp_sum_d(ss);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        sh[0][w] = s;
        sh[1][w] = ss;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3];
        ss = sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3];
        const double n = (double)R * gs;
        const double mean = s / n;
        double var = ss / n - mean * mean;
        if (var < 0.0) var = 0.0;
        const double rstd = 1.0 / sqrt(var + (double)eps);
        mean_rstd[(int64_t)frame * G + g] = make_float2((float)mean, (float)rstd);
        if (mean_out) mean_out[(int64_t)frame * G + g] = (float)mean;
        if (var_out) var_out[(int64_t)frame * G + g] = (float)var;
    }
}

__global__ void __launch_bounds__(256)
norm_apply_kernel(const float* __restrict__ x, int64_t ldx, int64_t R, int C, int G,
                  const float2* __restrict__ mean_rstd, const float* __restrict__ gamma,
                  const float* __restrict__ beta, const float* __restrict__ residual, int64_t ldr, int act,
                  float* __restrict__ y, int64_t ldy, int64_t total_rows) {
    const int gs = C / G;
    const int64_t total = total_rows * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / C;
        const int c = (int)(t - row * C);
        const int64_t frame = row / R;
        const float2 ms = __ldg(mean_rstd + frame * G + c / gs);
        float v = (__ldg(x + row * ldx + c) - ms.x) * ms.y;
        if (gamma) v = v * __ldg(gamma + c) + __ldg(beta + c);
        if (residual) v += __ldg(residual + row * ldr + c);
        y[row * ldy + c] = apply_act(v, act);
    }
}

__global__ void __launch_bounds__(256)
affine_rows_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int C, const float* __restrict__ scale,
                   const float* __restrict__ shift, const float* __restrict__ residual, int64_t ldr, int act,
                   float* __restrict__ y, int64_t ldy) {
    const int64_t total = rows * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / C;
        const int c = (int)(t - row * C);
        float v = __ldg(x + row * ldx + c);
        if (scale) v = v * __ldg(scale + c) + __ldg(shift + c);
        if (residual) v += __ldg(residual + row * ldr + c);
        y[row * ldy + c] = apply_act(v, act);
    }
}

// finalize from the fp32 per-tile column statistics a GEMM epilogue produced: tiles [frames][tiles_per_frame][C][2]
__global__ void __launch_bounds__(128)
norm_finalize_tiles_kernel(const float* __restrict__ tiles, int tiles_per_frame, int64_t R, int C, int G, float eps,
                           float2* __restrict__ mean_rstd) {
    const int frame = blockIdx.y, g = blockIdx.x;
    const int gs = C / G;
    double s = 0.0, ss = 0.0;
    for (int t = threadIdx.x; t < tiles_per_frame * gs; t += blockDim.x) {
        const int tile = t / gs, c = g * gs + (t - tile * gs);
        const float2 v = __ldg(reinterpret_cast<const float2*>(tiles) + ((int64_t)frame * tiles_per_frame + tile) * C + c);
        s += (double)v.x;
        ss += (double)v.y;
    }
    __shared__ double sh[2][4];
    s = warp_sum_d(s);
    ss = warp_sum_d(ss);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        sh[0][w] = s;
        sh[1][w] = ss;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3];
        ss = sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3];
        const double n = (double)R * gs;
        const double mean = s / n;
        double var = ss / n - mean * mean;
        if (var < 0.0) var = 0.0;
        mean_rstd[(int64_t)frame * G + g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
}

// ---- vectorised variants (C % 4 == 0): thread = (4-channel chunk, row lane); no integer divisions in the loop ----
// block = (TX, TY) with TX * TY = 256; grid = (row chunks <= kStatChunks, frames, channel-chunk groups)
// The last CTA of each (frame, channel-chunk group) column -- elected with a self-resetting global counter -- also
// reduces the partials of the groups that live entirely inside its channel range and writes (mean, rstd): the
// separate finalize launch disappears.  Requires group boundaries aligned to the CTA's channel span (host-checked).
__global__ void __launch_bounds__(256)
norm_stats_vec_kernel(const float* __restrict__ x, int64_t ldx, int64_t R, int C, double* __restrict__ partials,
                      int G, float eps, float2* __restrict__ mean_rstd, float* __restrict__ mean_out,
                      float* __restrict__ var_out, unsigned int* __restrict__ counters) {
    __shared__ double red[256][8 + 1];
    __shared__ int s_last;
    const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
    const int c4 = blockIdx.z * TX + tx;
    const int frame = blockIdx.y, chunk = blockIdx.x;
    const int64_t rows_per = (R + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = chunk * rows_per;
    const int64_t r1 = (r0 + rows_per < R) ? r0 + rows_per : R;
    double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    if (c4 * 4 < C) {
        const float* p = x + ((int64_t)frame * R) * ldx + c4 * 4;
        // four independent row loads in flight per thread; a 4-row batch is summed in fp32 (exact enough: 4 terms),
        // batches accumulate in fp64
        int64_t r = r0 + ty;
        for (; r + 3 * TY < r1; r += 4 * TY) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(p + (r + u * TY) * ldx));
            float fs[4] = {0.f, 0.f, 0.f, 0.f}, fq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                fs[0] += v[u].x; fq[0] = fmaf(v[u].x, v[u].x, fq[0]);
                fs[1] += v[u].y; fq[1] = fmaf(v[u].y, v[u].y, fq[1]);
                fs[2] += v[u].z; fq[2] = fmaf(v[u].z, v[u].z, fq[2]);
                fs[3] += v[u].w; fq[3] = fmaf(v[u].w, v[u].w, fq[3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                s[i] += (double)fs[i];
                ss[i] += (double)fq[i];
            }
        }
        for (; r < r1; r += TY) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p + r * ldx));
            s[0] += v.x; ss[0] += (double)v.x * v.x;
            s[1] += v.y; ss[1] += (double)v.y * v.y;
            s[2] += v.z; ss[2] += (double)v.z * v.z;
            s[3] += v.w; ss[3] += (double)v.w * v.w;
        }
    }
    const int tid = ty * TX + tx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        red[tid][i] = s[i];
        red[tid][4 + i] = ss[i];
    }
    __syncthreads();
    if (ty == 0 && c4 * 4 < C) {
        for (int y = 1; y < TY; ++y)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                s[i] += red[y * TX + tx][i];
                ss[i] += red[y * TX + tx][4 + i];
            }
        double* o = partials + (((int64_t)frame * kStatChunks + chunk) * C + c4 * 4) * 2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[2 * i] = s[i];
            o[2 * i + 1] = ss[i];
        }
    }
    if (counters == nullptr) return;
    // ---- elect the last CTA of this (frame, z) column
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int* ctr = counters + (frame * gridDim.z + blockIdx.z);
        const unsigned int prev = atomicAdd(ctr, 1u);
        s_last = (prev == gridDim.x - 1) ? 1 : 0;
        if (s_last) *ctr = 0u;  // self-reset for the next launch (stream-ordered)
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- finalize the groups inside channels [blockIdx.z*TX*4, +TX*4): one warp per group, round robin
    const int gs = C / G;
    const int ch0 = blockIdx.z * TX * 4;
    const int ch1 = (ch0 + TX * 4 < C) ? ch0 + TX * 4 : C;
    const int g0 = ch0 / gs, g1 = (ch1 + gs - 1) / gs;
    const int warp = tid >> 5, lane = tid & 31;
    const int nchunks = gridDim.x;
    const int items = nchunks * gs;
    if (items >= 512) {
        // few large groups (GroupNorm with wide groups): the whole CTA reduces one group at a time
        for (int g = g0; g < g1; ++g) {
            double a = 0.0, b = 0.0;
#pragma unroll 4
            for (int t = tid; t < items; t += 256) {
                const int ck = t / gs, c = g * gs + (t - ck * gs);
                const double* o = partials + (((int64_t)frame * kStatChunks + ck) * C + c) * 2;
                a += __ldcg(o);
                b += __ldcg(o + 1);
            }
            a = warp_sum_d(a);
            b = warp_sum_d(b);
            __syncthreads();
            if (lane == 0) {
                red[warp][0] = a;
                red[warp][1] = b;
            }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < 8; ++w) {
                    a += red[w][0];
                    b += red[w][1];
                }
                const double n = (double)R * gs;
                const double mean = a / n;
                double var = b / n - mean * mean;
                if (var < 0.0) var = 0.0;
                mean_rstd[(int64_t)frame * G + g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
                if (mean_out) mean_out[(int64_t)frame * G + g] = (float)mean;
                if (var_out) var_out[(int64_t)frame * G + g] = (float)var;
            }
        }
        return;
    }
    for (int g = g0 + warp; g < g1; g += 8) {  // many small groups (InstanceNorm): one warp per group
        double a = 0.0, b = 0.0;
        for (int t = lane; t < items; t += 32) {
            const int ck = t / gs, c = g * gs + (t - ck * gs);
            const double* o = partials + (((int64_t)frame * kStatChunks + ck) * C + c) * 2;
            a += __ldcg(o);
            b += __ldcg(o + 1);
        }
        a = warp_sum_d(a);
        b = warp_sum_d(b);
        if (lane == 0) {
            const double n = (double)R * gs;
            const double mean = a / n;
            double var = b / n - mean * mean;
            if (var < 0.0) var = 0.0;
            mean_rstd[(int64_t)frame * G + g] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
            if (mean_out) mean_out[(int64_t)frame * G + g] = (float)mean;
            if (var_out) var_out[(int64_t)frame * G + g] = (float)var;
        }
    }
}

__global__ void __launch_bounds__(256)
norm_apply_vec_kernel(const float* __restrict__ x, int64_t ldx, int64_t R, int C, int G,
                      const float2* __restrict__ mean_rstd, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const float* __restrict__ residual, int64_t ldr, int act,
                      float* __restrict__ y, int64_t ldy) {
    const int tx = threadIdx.x, ty = threadIdx.y, TX = blockDim.x, TY = blockDim.y;
    const int c4 = blockIdx.z * TX + tx;
    if (c4 * 4 >= C) return;
    const int frame = blockIdx.y;
    const int gs = C / G;
    float mean[4], a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c4 * 4 + i;
        const float2 ms = __ldg(mean_rstd + (int64_t)frame * G + c / gs);
        mean[i] = ms.x;
        a[i] = gamma ? ms.y * __ldg(gamma + c) : ms.y;
        b[i] = beta ? __ldg(beta + c) : 0.0f;
    }
    const int64_t rows_per = (R + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = blockIdx.x * rows_per;
    const int64_t r1 = (r0 + rows_per < R) ? r0 + rows_per : R;
    const int64_t base = (int64_t)frame * R;
    auto one = [&](const float4& v, const float4& rr, int64_t row) {
        float o[4] = {(v.x - mean[0]) * a[0] + b[0], (v.y - mean[1]) * a[1] + b[1], (v.z - mean[2]) * a[2] + b[2],
                      (v.w - mean[3]) * a[3] + b[3]};
        o[0] += rr.x; o[1] += rr.y; o[2] += rr.z; o[3] += rr.w;
        *reinterpret_cast<float4*>(y + row * ldy + c4 * 4) =
            make_float4(apply_act(o[0], act), apply_act(o[1], act), apply_act(o[2], act), apply_act(o[3], act));
    };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t r = r0 + ty;
    for (; r + 3 * TY < r1; r += 4 * TY) {  // four independent rows in flight per thread
        float4 v[4], rr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t row = base + r + u * TY;
            v[u] = __ldg(reinterpret_cast<const float4*>(x + row * ldx + c4 * 4));
            rr[u] = residual ? __ldg(reinterpret_cast<const float4*>(residual + row * ldr + c4 * 4)) : zero4;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) one(v[u], rr[u], base + r + u * TY);
    }
    for (; r < r1; r += TY) {
        const int64_t row = base + r;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + row * ldx + c4 * 4));
        const float4 rr = residual ? __ldg(reinterpret_cast<const float4*>(residual + row * ldr + c4 * 4)) : zero4;
        one(v, rr, row);
    }
}

// one warp per row; C <= 32*kLNMax
constexpr int kLNMax = 64;  // up to 2048 channels
__global__ void __launch_bounds__(128)
layer_norm_rows_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int C,
                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int act,
                       const float* __restrict__ residual, int64_t ldr, float* __restrict__ y, int64_t ldy) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* p = x + row * ldx;
    double s = 0.0;
    for (int c = lane; c < C; c += 32) s += (double)__ldg(p + c);
    s = warp_sum_d(s);
    const double mean = s / C;
    double ss = 0.0;
    for (int c = lane; c < C; c += 32) {
        const double d = (double)__ldg(p + c) - mean;
        ss += d * d;
    }
    ss = warp_sum_d(ss);
    const float rstd = (float)(1.0 / sqrt(ss / C + (double)eps));
    const float meanf = (float)mean;
    for (int c = lane; c < C; c += 32) {
        float v = (__ldg(p + c) - meanf) * rstd;
        if (gamma) v = v * __ldg(gamma + c) + __ldg(beta + c);
        v = apply_act(v, act);
        if (residual) v += __ldg(residual + row * ldr + c);
        y[row * ldy + c] = v;
    }
}

__global__ void __launch_bounds__(128)
l2norm_rows_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int C, const float* __restrict__ add,
                   int64_t ldadd, float* __restrict__ y, int64_t ldy, __half* __restrict__ yh = nullptr, int64_t ldh = 0) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* p = x + row * ldx;
    double ss = 0.0;
    for (int c = lane; c < C; c += 32) {
        const double v = (double)__ldg(p + c);
        ss += v * v;
    }
    ss = warp_sum_d(ss);
    // F.normalize: x / max(||x||, eps), eps = 1e-12
    const float denom = fmaxf((float)sqrt(ss), 1e-12f);
    for (int c = lane; c < C; c += 32) {
        float v = __ldg(p + c) / denom;
        if (add) v += __ldg(add + row * ldadd + c);
        y[row * ldy + c] = v;
        if (yh) yh[row * ldh + c] = __float2half_rn(v);
    }
}

// column sums of squares over the L rows of each frame -> partial [frames][chunks][C] then scale
__global__ void __launch_bounds__(256)
colsq_kernel(const float* __restrict__ x, int64_t ldx, int64_t L, int C, double* __restrict__ partials) {
    const int frame = blockIdx.z, chunk = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t rows_per = (L + kStatChunks - 1) / kStatChunks;
    const int64_t r0 = chunk * rows_per;
    const int64_t r1 = (r0 + rows_per < L) ? r0 + rows_per : L;
    if (c >= C) return;
    double ss = 0.0;
    const float* p = x + ((int64_t)frame * L) * ldx + c;
    for (int64_t r = r0; r < r1; ++r) {
        const double v = (double)__ldg(p + r * ldx);
        ss += v * v;
    }
    partials[((int64_t)frame * kStatChunks + chunk) * C + c] = ss;
}

// inv[frame, c] = 1 / max(sqrt(sum of the chunk partials), 1e-12)
__global__ void __launch_bounds__(128)
colnorm_finalize_kernel(const double* __restrict__ partials, int C, int total, float* __restrict__ inv) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;  // frame * C + c
    if (t >= total) return;
    const int frame = t / C, c = t - frame * C;
    double ss = 0.0;
#pragma unroll 8
    for (int k = 0; k < kStatChunks; ++k) ss += partials[((int64_t)frame * kStatChunks + k) * C + c];
    inv[t] = fmaxf((float)sqrt(ss), 1e-12f);
}

__global__ void __launch_bounds__(256)
colnorm_apply_kernel(const float* __restrict__ x, int64_t ldx, int64_t L, int C, int64_t total_rows,
                     const float* __restrict__ denom, float* __restrict__ y, int64_t ldy) {
    const int64_t total = total_rows * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / C;
        const int c = (int)(t - row * C);
        const int64_t frame = row / L;
        y[row * ldy + c] = __ldg(x + row * ldx + c) / __ldg(denom + frame * C + c);
    }
}

}  // namespace cofi

using namespace cofi;

static unsigned ew_blocks(int64_t total, int threads) {
    int64_t b = ceil_div(total, threads);
    const int64_t cap = 148 * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int64_t cofi_norm_rows_workspace(int frames, int C) {
    // partial sums + (mean, rstd) table
    return (int64_t)frames * kStatChunks * C * 2 * sizeof(double) + (int64_t)frames * C * sizeof(float2) + 256;
}

static unsigned int* norm_counters() {
    // one zero-initialised, self-resetting election counter array per device (4096 (frame, chunk-group) columns)
    static unsigned int* ctr[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!ctr[dev]) {
        void* p = nullptr;
        if (cudaMalloc(&p, 4096 * sizeof(unsigned int)) != cudaSuccess) return nullptr;
        cudaMemset(p, 0, 4096 * sizeof(unsigned int));
        ctr[dev] = reinterpret_cast<unsigned int*>(p);
    }
    return ctr[dev];
}

extern "C" int cofi_norm_rows_init(void) { return norm_counters() ? COFI_OK : COFI_ECUDA; }

extern "C" int cofi_norm_rows(const float* x, int64_t ldx, int64_t R, int C, int frames, int G, const float* gamma,
                              const float* beta, float eps, const float* residual, int64_t ldr, int act, float* y,
                              int64_t ldy, void* partials, float* mean_out, float* var_out, void* stream) {
    COFI_REQUIRE(x && y && partials, "cofi_norm_rows: null pointer");
    COFI_REQUIRE(R > 0 && C > 0 && frames > 0 && G > 0 && C % G == 0, "cofi_norm_rows: bad shape R=%lld C=%d G=%d",
                 (long long)R, C, G);
    COFI_REQUIRE((gamma == nullptr) == (beta == nullptr), "cofi_norm_rows: gamma and beta go together");
    COFI_REQUIRE(((uintptr_t)partials % 16) == 0, "cofi_norm_rows: workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    double* part = reinterpret_cast<double*>(partials);
    float2* mr = reinterpret_cast<float2*>(part + (int64_t)frames * kStatChunks * C * 2);
    const bool vec = (C % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && (!residual || ldr % 4 == 0) &&
                     ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && (!residual || (uintptr_t)residual % 16 == 0);
    const int C4 = C / 4;
    bool fused_finalize = false;
    int TX = 1;
    while (TX < C4 && TX < 32) TX <<= 1;
    const int TY = 256 / TX;
    const unsigned zgroups = (unsigned)ceil_div(C4 > 0 ? C4 : 1, TX);
    int stat_chunks = (int)(R / (4 * TY) < 1 ? 1 : (R / (4 * TY) > kStatChunks ? kStatChunks : R / (4 * TY)));
    if (vec) {
        dim3 grid((unsigned)stat_chunks, frames, zgroups), block(TX, TY);
        const int gs = C / G, span = TX * 4;
        // fused finalize needs every group inside one CTA's channel span and an allocated counter array
        // Measured on B200: the in-kernel election + finalize (MEMBAR.GPU per CTA, serial tail of the last CTA) costs
        // more than the 5.7 us finalize launch it removes (2.46 -> 3.16 ms/step), so it stays off by default.
        static const bool fuse = getenv("COFI_NORM_FUSED_FINALIZE") != nullptr;
        unsigned int* ctr = (fuse && (span % gs == 0 || gs % span == 0) && gs <= span && (int64_t)frames * zgroups <= 4096)
                                ? norm_counters() : nullptr;
        fused_finalize = ctr != nullptr;
        norm_stats_vec_kernel<<<grid, block, 0, st>>>(x, ldx, R, C, part, G, eps, mr, mean_out, var_out, ctr);
        int rc = check_launch("cofi_norm_rows(stats)");
        if (rc) return rc;
    } else {
        const int threads = C >= 256 ? 256 : (C >= 128 ? 128 : (C >= 64 ? 64 : 32));
        dim3 grid((unsigned)ceil_div(C, threads), kStatChunks, frames);
        norm_stats_kernel<<<grid, threads, 0, st>>>(x, ldx, R, C, part);
        int rc = check_launch("cofi_norm_rows(stats)");
        if (rc) return rc;
    }
    if (!fused_finalize) {
        dim3 grid(G, frames);
        norm_finalize_kernel<<<grid, 128, 0, st>>>(part, vec ? stat_chunks : kStatChunks, R, C, G, eps, mr, mean_out, var_out);
        int rc = check_launch("cofi_norm_rows(finalize)");
        if (rc) return rc;
    }
    const int64_t rows = R * frames;
    if (vec) {
        // enough row chunks to fill the machine: ~4 CTAs per SM overall
        int64_t want = (148 * 8) / ((int64_t)frames * zgroups);
        if (want < 1) want = 1;
        int64_t maxc = ceil_div(R, TY);
        if (want > maxc) want = maxc;
        dim3 grid((unsigned)want, frames, zgroups), block(TX, TY);
        norm_apply_vec_kernel<<<grid, block, 0, st>>>(x, ldx, R, C, G, mr, gamma, beta, residual, ldr, act, y, ldy);
        return check_launch("cofi_norm_rows(apply)");
    }
    norm_apply_kernel<<<ew_blocks(rows * C, 256), 256, 0, st>>>(x, ldx, R, C, G, mr, gamma, beta, residual, ldr, act,
                                                                y, ldy, rows);
    return check_launch("cofi_norm_rows(apply)");
}

extern "C" int cofi_affine_rows(const float* x, int64_t ldx, int64_t rows, int C, const float* scale,
                                const float* shift, const float* residual, int64_t ldr, int act, float* y,
                                int64_t ldy, void* stream) {
    COFI_REQUIRE(x && y, "cofi_affine_rows: null pointer");
    COFI_REQUIRE((scale == nullptr) == (shift == nullptr), "cofi_affine_rows: scale and shift go together");
    if (rows == 0) return COFI_OK;
    affine_rows_kernel<<<ew_blocks(rows * C, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, rows, C, scale, shift,
                                                                                  residual, ldr, act, y, ldy);
    return check_launch("cofi_affine_rows");
}

extern "C" int cofi_layer_norm_rows(const float* x, int64_t ldx, int64_t rows, int C, const float* gamma,
                                    const float* beta, float eps, int act, const float* residual, int64_t ldr,
                                    float* y, int64_t ldy, void* stream) {
    COFI_REQUIRE(x && y, "cofi_layer_norm_rows: null pointer");
    COFI_REQUIRE(C > 0 && C <= 32 * kLNMax, "cofi_layer_norm_rows: C=%d out of range", C);
    COFI_REQUIRE((gamma == nullptr) == (beta == nullptr), "cofi_layer_norm_rows: gamma and beta go together");
    if (rows == 0) return COFI_OK;
    const int wpb = 4;
    layer_norm_rows_kernel<<<(unsigned)ceil_div(rows, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        x, ldx, rows, C, gamma, beta, eps, act, residual, ldr, y, ldy);
    return check_launch("cofi_layer_norm_rows");
}

extern "C" int cofi_l2norm_rows(const float* x, int64_t ldx, int64_t rows, int C, const float* add, int64_t ldadd,
                                float* y, int64_t ldy, void* stream) {
    COFI_REQUIRE(x && y && C > 0, "cofi_l2norm_rows: bad argument");
    if (rows == 0) return COFI_OK;
    const int wpb = 4;
    l2norm_rows_kernel<<<(unsigned)ceil_div(rows, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(x, ldx, rows, C, add,
                                                                                            ldadd, y, ldy);
    return check_launch("cofi_l2norm_rows");
}

extern "C" int cofi_l2norm_rows_f16(const float* x, int64_t ldx, int64_t rows, int C, float* y, int64_t ldy, void* y_f16,
                                    int64_t ldh, void* stream) {
    COFI_REQUIRE(x && y && y_f16 && C > 0, "cofi_l2norm_rows_f16: bad argument");
    if (rows == 0) return COFI_OK;
    const int wpb = 4;
    l2norm_rows_kernel<<<(unsigned)ceil_div(rows, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        x, ldx, rows, C, nullptr, 0, y, ldy, reinterpret_cast<__half*>(y_f16), ldh);
    return check_launch("cofi_l2norm_rows_f16");
}

extern "C" int64_t cofi_colnorm_workspace(int frames, int C) {
    return (int64_t)frames * kStatChunks * C * sizeof(double) + (int64_t)frames * C * sizeof(float) + 64;
}

extern "C" int cofi_colnorm_rows(const float* x, int64_t ldx, int64_t L, int C, int frames, void* colsq, float* y,
                                 int64_t ldy, void* stream) {
    COFI_REQUIRE(x && y && colsq, "cofi_colnorm_rows: null pointer");
    COFI_REQUIRE(L > 0 && C > 0 && frames > 0, "cofi_colnorm_rows: bad shape");
    COFI_REQUIRE(((uintptr_t)colsq % 8) == 0, "cofi_colnorm_rows: workspace must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    double* part = reinterpret_cast<double*>(colsq);  // frames*kStatChunks*C doubles
    const int threads = C >= 128 ? 128 : (C >= 64 ? 64 : 32);
    dim3 grid((unsigned)ceil_div(C, threads), kStatChunks, frames);
    colsq_kernel<<<grid, threads, 0, st>>>(x, ldx, L, C, part);
    int rc = check_launch("cofi_colnorm_rows(colsq)");
    if (rc) return rc;
    float* denom = reinterpret_cast<float*>(part + (int64_t)frames * kStatChunks * C);
    colnorm_finalize_kernel<<<(unsigned)ceil_div((int64_t)frames * C, 128), 128, 0, st>>>(part, C, frames * C, denom);
    rc = check_launch("cofi_colnorm_rows(finalize)");
    if (rc) return rc;
    const int64_t rows = L * frames;
    colnorm_apply_kernel<<<ew_blocks(rows * C, 256), 256, 0, st>>>(x, ldx, L, C, rows, denom, y, ldy);
    return check_launch("cofi_colnorm_rows(apply)");
}

// GroupNorm whose statistics were produced by the preceding GEMM's epilogue (cofi_gemm_colstats): finalize + apply only.
extern "C" int cofi_norm_rows_pre(const float* x, int64_t ldx, int64_t R, int C, int frames, int G, const float* gamma,
                                  const float* beta, float eps, const float* residual, int64_t ldr, int act, float* y,
                                  int64_t ldy, const float* tile_stats, void* workspace, void* stream) {
    COFI_REQUIRE(x && y && tile_stats && workspace, "cofi_norm_rows_pre: null pointer");
    COFI_REQUIRE(R > 0 && R % 128 == 0 && C > 0 && C % 4 == 0 && frames > 0 && G > 0 && C % G == 0,
                 "cofi_norm_rows_pre: R must be a multiple of 128, C of 4");
    COFI_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && (!residual || ldr % 4 == 0), "cofi_norm_rows_pre: leading dimensions % 4");
    COFI_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0 && ((uintptr_t)workspace % 8) == 0 &&
                     (!residual || (uintptr_t)residual % 16 == 0), "cofi_norm_rows_pre: alignment");
    cudaStream_t st = (cudaStream_t)stream;
    float2* mr = reinterpret_cast<float2*>(workspace);  // frames*G float2
    {
        dim3 grid(G, frames);
        norm_finalize_tiles_kernel<<<grid, 128, 0, st>>>(tile_stats, (int)(R / 128), R, C, G, eps, mr);
        int rc = check_launch("cofi_norm_rows_pre(finalize)");
        if (rc) return rc;
    }
    const int C4 = C / 4;
    int TX = 1;
    while (TX < C4 && TX < 32) TX <<= 1;
    const int TY = 256 / TX;
    const unsigned zgroups = (unsigned)ceil_div(C4, TX);
    int64_t want = (148 * 8) / ((int64_t)frames * zgroups);
    if (want < 1) want = 1;
    const int64_t maxc = ceil_div(R, TY);
    if (want > maxc) want = maxc;
    dim3 grid((unsigned)want, frames, zgroups), block(TX, TY);
    norm_apply_vec_kernel<<<grid, block, 0, st>>>(x, ldx, R, C, G, mr, gamma, beta, residual, ldr, act, y, ldy);
    return check_launch("cofi_norm_rows_pre(apply)");
}

// statistics only: (mean, rstd) per (frame, group) as float2 -- consumed by cofi_norm_rows_bwd
extern "C" int cofi_norm_rows_stats(const float* x, int64_t ldx, int64_t R, int C, int frames, int G, float eps,
                                    void* partials, float* mean_rstd, void* stream) {
    COFI_REQUIRE(x && partials && mean_rstd && R > 0 && C > 0 && frames > 0 && G > 0 && C % G == 0,
                 "cofi_norm_rows_stats: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    double* part = reinterpret_cast<double*>(partials);
    const int threads = C >= 256 ? 256 : (C >= 128 ? 128 : (C >= 64 ? 64 : 32));
    dim3 grid((unsigned)ceil_div(C, threads), kStatChunks, frames);
    norm_stats_kernel<<<grid, threads, 0, st>>>(x, ldx, R, C, part);
    int rc = check_launch("cofi_norm_rows_stats(stats)");
    if (rc) return rc;
    dim3 g2(G, frames);
    norm_finalize_kernel<<<g2, 128, 0, st>>>(part, kStatChunks, R, C, G, eps, reinterpret_cast<float2*>(mean_rstd), nullptr,
                                            nullptr);
    return check_launch("cofi_norm_rows_stats(finalize)");
}
