// backward.cu -- gradient kernels of the hot path (training, BASELINE config 5: the reference trains with
// loss.backward() through ATen autograd, train.py:285; here every forward op has a hand-written backward).
// HBM-bound element-wise / gather / scatter work: coalesced along channels, fp64 statistics, float atomics only where a
// scatter is inherent (neighbour gathers, max-pool, bilinear up-sampling, patch extraction).
#include "common.cuh"

namespace cofi {

static unsigned ew_blocks_b(int64_t total, int threads) {
    int64_t b = ceil_div(total, threads);
    const int64_t cap = 148 * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------ element-wise
// dx = dy * act'(y)   (y = activation OUTPUT: relu / lrelu: sign(y) == sign(pre-activation); sigmoid: y(1-y))
__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, int64_t n, int act, float* __restrict__ dx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = dy[i], v = y[i];
        float r = g;
        if (act == COFI_ACT_RELU) r = v > 0.0f ? g : 0.0f;
        else if (act == COFI_ACT_LRELU01) r = v > 0.0f ? g : 0.1f * g;
        else if (act == COFI_ACT_SIGMOID) r = g * v * (1.0f - v);
        dx[i] = r;
    }
}

// y[row, c] = x[row, c] / rowdiv[row]
__global__ void __launch_bounds__(256)
rowscale_kernel(const float* __restrict__ x, int64_t rows, int C, const float* __restrict__ rowdiv,
                float* __restrict__ y) {
    const int64_t total = rows * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
        y[t] = x[t] / __ldg(rowdiv + t / C);
}

// column sums, deterministic two-stage: partial[chunk][c] (fp64) then out[c]
constexpr int kColChunks = 128;
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int C, double* __restrict__ part) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ty = threadIdx.x >> 5;  // 8 row lanes
    const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = blockIdx.y * per, r1 = (r0 + per < rows) ? r0 + per : rows;
    double s = 0.0;
    if (c < C)
        for (int64_t r = r0 + ty; r < r1; r += 8) s += (double)__ldg(x + r * ldx + c);
    __shared__ double sh[8][33];
    sh[ty][threadIdx.x & 31] = s;
    __syncthreads();
    if (ty == 0 && c < C) {
        for (int k = 1; k < 8; ++k) s += sh[k][threadIdx.x & 31];
        part[(int64_t)blockIdx.y * C + c] = s;
    }
}
// 4 columns per thread (float4), two rows in flight, fp32 partials folded into fp64 across threads
__global__ void __launch_bounds__(256)
colsum_partial_v4_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int C, double* __restrict__ part) {
    const int cl = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + cl) * 4;
    const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = blockIdx.y * per, r1 = (r0 + per < rows) ? r0 + per : rows;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (c < C) {
        int64_t r = r0 + ty;
        for (; r + 8 < r1; r += 16) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(x + r * ldx + c));
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + (r + 8) * ldx + c));
            a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
            b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
        }
        if (r < r1) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(x + r * ldx + c));
            a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
        }
    }
    __shared__ double sh[8][32][4];
    sh[ty][cl][0] = (double)a.x + (double)b.x;
    sh[ty][cl][1] = (double)a.y + (double)b.y;
    sh[ty][cl][2] = (double)a.z + (double)b.z;
    sh[ty][cl][3] = (double)a.w + (double)b.w;
    __syncthreads();
    if (ty < 4 && c < C) {  // thread (ty, cl) folds component ty of column group cl
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += sh[k][cl][ty];
        part[(int64_t)blockIdx.y * C + c + ty] = s;
    }
}
__global__ void __launch_bounds__(256)
colsum_final_kernel(const double* __restrict__ part, int chunks, int C, float* __restrict__ out, int accumulate) {
    // 32 columns per block, 8 row-lanes: lane ty folds chunks ty, ty+8, ... (independent loads in flight instead of one
    // serial chain of `chunks` L2 round trips), then a fixed-order fold over the 8 partials -- deterministic
    __shared__ double sh[8][33];
    const int cl = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    double s = 0.0;
    if (c < C) {
#pragma unroll 4
        for (int k = ty; k < chunks; k += 8) s += part[(int64_t)k * C + c];
    }
    sh[ty][cl] = s;
    __syncthreads();
    if (ty == 0 && c < C) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sh[k][cl];
        out[c] = (accumulate ? out[c] : 0.0f) + (float)t;
    }
}

// ------------------------------------------------------------------------------------------ norm_rows backward
// forward: xh = (x-mean)*rstd ; z = xh*gamma+beta (+res) ; y = act(z).
// pass 1: per (frame, chunk, channel) partial sums of dz and dz*xh  (dz = dy*act'(y))
// pass 2 (finalize): per (frame, group): s1 = sum_c gamma_c A_c, s2 = sum_c gamma_c B_c ; dgamma_c = sum_frames B_c ;
//                    dbeta_c = sum_frames A_c
// pass 3: dx = rstd*(dz*gamma - s1/n - xh*s2/n) ; dres = dz
constexpr int kBwdChunks = 64;
__device__ __forceinline__ float act_grad(float g, float y, int act) {
    if (act == COFI_ACT_RELU) return y > 0.0f ? g : 0.0f;
    if (act == COFI_ACT_LRELU01) return y > 0.0f ? g : 0.1f * g;
    if (act == COFI_ACT_SIGMOID) return g * y * (1.0f - y);
    return g;
}
__global__ void __launch_bounds__(256)
norm_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y,
                      int64_t R, int C, int G, const float2* __restrict__ mean_rstd, int act,
                      double* __restrict__ part /* [frames][chunks][C][2] */) {
    const int frame = blockIdx.z, chunk = blockIdx.y;
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ty = threadIdx.x >> 5;
    const int64_t per = (R + kBwdChunks - 1) / kBwdChunks;
    const int64_t r0 = chunk * per, r1 = (r0 + per < R) ? r0 + per : R;
    double a = 0.0, b = 0.0;
    if (c < C) {
        const float2 ms = __ldg(mean_rstd + (int64_t)frame * G + c / (C / G));
        const int64_t base = (int64_t)frame * R;
        for (int64_t r = r0 + ty; r < r1; r += 8) {
            const int64_t i = (base + r) * C + c;
            const float dz = act_grad(__ldg(dy + i), __ldg(y + i), act);
            const float xh = (__ldg(x + i) - ms.x) * ms.y;
            a += (double)dz;
            b += (double)dz * (double)xh;
        }
    }
    __shared__ double sh[2][8][33];
    sh[0][ty][threadIdx.x & 31] = a;
    sh[1][ty][threadIdx.x & 31] = b;
    __syncthreads();
    if (ty == 0 && c < C) {
        for (int k = 1; k < 8; ++k) {
            a += sh[0][k][threadIdx.x & 31];
            b += sh[1][k][threadIdx.x & 31];
        }
        double* o = part + (((int64_t)frame * kBwdChunks + chunk) * C + c) * 2;
        o[0] = a;
        o[1] = b;
    }
}
// float4 variants (C % 4 == 0): four channels per thread, 16-byte loads, fp32 partial sums over the <= 40 rows a thread
// sees, promoted to double for the cross-thread reduction; 2-D indexing instead of a 64-bit division per element.
__global__ void __launch_bounds__(256)
norm_bwd_stats_v4_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y,
                         int64_t R, int C, int G, const float2* __restrict__ mean_rstd, int act,
                         double* __restrict__ part /* [frames][chunks][C][2] */) {
    const int frame = blockIdx.z, chunk = blockIdx.y;
    const int cl = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + cl) * 4;
    const int64_t per = (R + kBwdChunks - 1) / kBwdChunks;
    const int64_t r0 = chunk * per, r1 = (r0 + per < R) ? r0 + per : R;
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C) {
        const int gs = C / G;
        float mean[4], rstd[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 ms = __ldg(mean_rstd + (int64_t)frame * G + (c + e) / gs);
            mean[e] = ms.x;
            rstd[e] = ms.y;
        }
        const int64_t base = (int64_t)frame * R;
        for (int64_t r = r0 + ty; r < r1; r += 8) {
            const int64_t i = (base + r) * C + c;
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(dy + i));
            const float4 y4 = __ldg(reinterpret_cast<const float4*>(y + i));
            const float4 x4 = __ldg(reinterpret_cast<const float4*>(x + i));
            const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, yv[4] = {y4.x, y4.y, y4.z, y4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float dz = act_grad(gv[e], yv[e], act);
                a[e] += dz;
                b[e] = fmaf(dz, (xv[e] - mean[e]) * rstd[e], b[e]);
            }
        }
    }
    __shared__ double sh[8][32][8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        sh[ty][cl][e] = (double)a[e];
        sh[ty][cl][4 + e] = (double)b[e];
    }
    __syncthreads();
    if (c < C) {  // thread (ty, cl) folds entry ty (0..3: sums of dz, 4..7: sums of dz*xh) of column group cl
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += sh[k][cl][ty];
        part[(((int64_t)frame * kBwdChunks + chunk) * C + c + (ty & 3)) * 2 + (ty >> 2)] = t;
    }
}
__global__ void __launch_bounds__(256)
norm_bwd_apply_v4_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y,
                         int64_t R, int C, int G, const float2* __restrict__ mean_rstd, const float2* __restrict__ s12,
                         const float* __restrict__ gamma, int act, float* __restrict__ dx, float* __restrict__ dres) {
    const int frame = blockIdx.z;
    const int cl = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + cl) * 4;
    if (c >= C) return;
    const int gs = C / G;
    const float inv_n = 1.0f / ((float)R * (float)gs);
    float mean[4], rstd[4], k1[4], k2[4], gm[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int g = (c + e) / gs;
        const float2 ms = __ldg(mean_rstd + (int64_t)frame * G + g);
        const float2 sv = __ldg(s12 + (int64_t)frame * G + g);
        mean[e] = ms.x;
        rstd[e] = ms.y;
        k1[e] = sv.x * inv_n;
        k2[e] = sv.y * inv_n;
        gm[e] = gamma ? __ldg(gamma + c + e) : 1.0f;
    }
    const int64_t per = (R + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = blockIdx.y * per, r1 = (r0 + per < R) ? r0 + per : R;
    const int64_t base = (int64_t)frame * R;
    for (int64_t r = r0 + ty; r < r1; r += 8) {
        const int64_t i = (base + r) * C + c;
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(dy + i));
        const float4 y4 = __ldg(reinterpret_cast<const float4*>(y + i));
        const float4 x4 = __ldg(reinterpret_cast<const float4*>(x + i));
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, yv[4] = {y4.x, y4.y, y4.z, y4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
        float o[4], z[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            z[e] = act_grad(gv[e], yv[e], act);
            const float xh = (xv[e] - mean[e]) * rstd[e];
            o[e] = rstd[e] * (z[e] * gm[e] - k1[e] - xh * k2[e]);
        }
        *reinterpret_cast<float4*>(dx + i) = make_float4(o[0], o[1], o[2], o[3]);
        if (dres) *reinterpret_cast<float4*>(dres + i) = make_float4(z[0], z[1], z[2], z[3]);
    }
}
// per (frame, channel): reduce chunks -> AB[frame][c][2] (double)
__global__ void __launch_bounds__(128)
norm_bwd_reduce_kernel(const double* __restrict__ part, int C, int total /* frames*C */, double* __restrict__ ab) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int frame = t / C, c = t - frame * C;
    double a = 0.0, b = 0.0;
#pragma unroll 8
    for (int k = 0; k < kBwdChunks; ++k) {
        const double* o = part + (((int64_t)frame * kBwdChunks + k) * C + c) * 2;
        a += o[0];
        b += o[1];
    }
    ab[(int64_t)t * 2] = a;
    ab[(int64_t)t * 2 + 1] = b;
}
// per (frame, group): s1, s2 ; and (block y == 0 only) dgamma/dbeta over frames
__global__ void __launch_bounds__(128)
norm_bwd_finalize_kernel(const double* __restrict__ ab, int C, int G, int frames, const float* __restrict__ gamma,
                         float2* __restrict__ s12 /* [frames][G] */, float* __restrict__ dgamma,
                         float* __restrict__ dbeta, int accumulate) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int gs = C / G;
    if (t < frames * G) {
        const int frame = t / G, g = t - frame * G;
        double s1 = 0.0, s2 = 0.0;
        for (int c = g * gs; c < (g + 1) * gs; ++c) {
            const double gm = gamma ? (double)gamma[c] : 1.0;
            s1 += gm * ab[((int64_t)frame * C + c) * 2];
            s2 += gm * ab[((int64_t)frame * C + c) * 2 + 1];
        }
        s12[t] = make_float2((float)s1, (float)s2);
    }
    if (dgamma && t < C) {
        double a = 0.0, b = 0.0;
        for (int f = 0; f < frames; ++f) {
            a += ab[((int64_t)f * C + t) * 2];
            b += ab[((int64_t)f * C + t) * 2 + 1];
        }
        dbeta[t] = (accumulate ? dbeta[t] : 0.0f) + (float)a;
        dgamma[t] = (accumulate ? dgamma[t] : 0.0f) + (float)b;
    }
}
__global__ void __launch_bounds__(256)
norm_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y,
                      int64_t R, int C, int G, int64_t total_rows, const float2* __restrict__ mean_rstd,
                      const float2* __restrict__ s12, const float* __restrict__ gamma, int act,
                      float* __restrict__ dx, float* __restrict__ dres) {
    const int gs = C / G;
    const float inv_n = 1.0f / ((float)R * (float)gs);
    const int64_t total = total_rows * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / C;
        const int c = (int)(t - row * C);
        const int64_t frame = row / R;
        const int g = c / gs;
        const float2 ms = __ldg(mean_rstd + frame * G + g);
        const float2 s = __ldg(s12 + frame * G + g);
        const float dz = act_grad(__ldg(dy + t), __ldg(y + t), act);
        const float xh = (__ldg(x + t) - ms.x) * ms.y;
        const float gm = gamma ? __ldg(gamma + c) : 1.0f;
        dx[t] = ms.y * (dz * gm - s.x * inv_n - xh * s.y * inv_n);
        if (dres) dres[t] = dz;
    }
}

// ------------------------------------------------------------------------------------------ row norms backward
// LayerNorm: y = act(xh*gamma+beta) + res ; outputs dx, t1 = dz*xh, t2 = dz (column sums of t1/t2 = dgamma/dbeta)
__global__ void __launch_bounds__(128)
layer_norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t rows, int C,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int act,
                      float* __restrict__ dx, float* __restrict__ t1, float* __restrict__ t2) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* p = x + row * C;
    const float* g = dy + row * C;
    double s = 0.0;
    for (int c = lane; c < C; c += 32) s += (double)p[c];
    s = warp_sum_d(s);
    const double mean = s / C;
    double ss = 0.0;
    for (int c = lane; c < C; c += 32) {
        const double d = (double)p[c] - mean;
        ss += d * d;
    }
    ss = warp_sum_d(ss);
    const float rstd = (float)(1.0 / sqrt(ss / C + (double)eps));
    const float meanf = (float)mean;
    double a = 0.0, b = 0.0;  // sum(gm*dz), sum(gm*dz*xh)
    for (int c = lane; c < C; c += 32) {
        const float xh = (p[c] - meanf) * rstd;
        const float z = xh * gamma[c] + beta[c];
        float dz = g[c];
        if (act == COFI_ACT_RELU) dz = z > 0.0f ? dz : 0.0f;
        a += (double)dz * gamma[c];
        b += (double)dz * gamma[c] * xh;
    }
    a = warp_sum_d(a) / C;
    b = warp_sum_d(b) / C;
    for (int c = lane; c < C; c += 32) {
        const float xh = (p[c] - meanf) * rstd;
        const float z = xh * gamma[c] + beta[c];
        float dz = g[c];
        if (act == COFI_ACT_RELU) dz = z > 0.0f ? dz : 0.0f;
        dx[row * C + c] = rstd * (dz * gamma[c] - (float)a - xh * (float)b);
        t1[row * C + c] = dz * xh;
        t2[row * C + c] = dz;
    }
}
// y = x / max(|x|, eps): dx = (dy - y (y . dy)) / n
__global__ void __launch_bounds__(128)
l2norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t rows, int C,
                  float* __restrict__ dx) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* p = x + row * C;
    const float* g = dy + row * C;
    double ss = 0.0, dot = 0.0;
    for (int c = lane; c < C; c += 32) {
        ss += (double)p[c] * p[c];
        dot += (double)p[c] * g[c];
    }
    ss = warp_sum_d(ss);
    dot = warp_sum_d(dot);
    const float n = fmaxf((float)sqrt(ss), 1e-12f);
    const float k = (float)(dot / ((double)n * n));  // (y . dy) / n with y = x / n
    for (int c = lane; c < C; c += 32) dx[row * C + c] = (g[c] - p[c] * k) / n;
}
// column-wise (sequence axis) normalisation backward: y[l,c] = x[l,c] / n_c
__global__ void __launch_bounds__(256)
colnorm_bwd_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t L, int C,
                           double* __restrict__ part /* [frames][chunks][C][2]: sum x^2, sum x*dy */) {
    const int frame = blockIdx.z, chunk = blockIdx.y;
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ty = threadIdx.x >> 5;
    const int64_t per = (L + kBwdChunks - 1) / kBwdChunks;
    const int64_t r0 = chunk * per, r1 = (r0 + per < L) ? r0 + per : L;
    double a = 0.0, b = 0.0;
    if (c < C)
        for (int64_t r = r0 + ty; r < r1; r += 8) {
            const int64_t i = ((int64_t)frame * L + r) * C + c;
            const double xv = (double)__ldg(x + i);
            a += xv * xv;
            b += xv * (double)__ldg(dy + i);
        }
    __shared__ double sh[2][8][33];
    sh[0][ty][threadIdx.x & 31] = a;
    sh[1][ty][threadIdx.x & 31] = b;
    __syncthreads();
    if (ty == 0 && c < C) {
        for (int k = 1; k < 8; ++k) {
            a += sh[0][k][threadIdx.x & 31];
            b += sh[1][k][threadIdx.x & 31];
        }
        double* o = part + (((int64_t)frame * kBwdChunks + chunk) * C + c) * 2;
        o[0] = a;
        o[1] = b;
    }
}
__global__ void __launch_bounds__(256)
colnorm_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t L, int C,
                         int64_t total_rows, const double* __restrict__ ab /* [frames][C][2] */, float* __restrict__ dx) {
    const int64_t total = total_rows * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / C;
        const int c = (int)(t - row * C);
        const int64_t frame = row / L;
        const double ss = ab[(frame * C + c) * 2], dot = ab[(frame * C + c) * 2 + 1];
        const float n = fmaxf((float)sqrt(ss), 1e-12f);
        dx[t] = (__ldg(dy + t) - __ldg(x + t) * (float)(dot / ((double)n * n))) / n;
    }
}

// ------------------------------------------------------------------------------------------ gathers backward
// dx[idx[i*stride], :] += dy[i, :]   (dx pre-zeroed; shadow indices dropped)
__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(const float* __restrict__ dy, int64_t ldy, int C, const int64_t* __restrict__ idx,
                        int64_t idx_stride, int64_t Mq, int64_t Ns, int64_t total_q, float* __restrict__ dx) {
    const int64_t total = total_q * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = t / C;
        const int c = (int)(t - m * C);
        const int64_t frame = m / Mq;
        const int64_t src = __ldg(idx + m * idx_stride);
        if (src >= 0 && src < Ns) atomicAdd(dx + (frame * Ns + src) * C + c, __ldg(dy + m * ldy + c));
    }
}
// maxpool over neighbours backward: gradient goes to the FIRST neighbour attaining the maximum (torch.max semantics).
// Same gather as the forward kernel (kpconv.cu): warp per query, the 128 neighbour ids held 4 per lane, float4 per lane
// over the channels, four independent gathers in flight, narrow rows walked by several lane groups on interleaved
// neighbours.  The arg-max is tracked per channel as (value, neighbour position): positions order ties across groups.
__device__ __forceinline__ void amax_take(float& bv, int& bh, int& bi, float v, int h, int id) {
    if (v > bv || (v == bv && h < bh)) {
        bv = v;
        bh = h;
        bi = id;
    }
}
__global__ void __launch_bounds__(128)
maxpool_rows_bwd_kernel(const float* __restrict__ x, int C, const int64_t* __restrict__ nbr, int H, int64_t Mq,
                        int64_t Ns, int64_t total_q, const float* __restrict__ dy, float* __restrict__ dx) {
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= total_q) return;
    const int64_t frame = m / Mq;
    const float* xb = x + frame * Ns * C;
    int idx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int h = j * 32 + lane;
        int64_t id = (h < H) ? __ldcs(nbr + m * H + h) : -2;  // -2: beyond H (ignored), -1: shadow (value 0, no gradient)
        if (id >= Ns || id < 0) id = (h < H) ? -1 : -2;
        idx[j] = (int)id;
    }
    const bool vec = (C % 4) == 0;
    const int rl = C >> 2;
    const int lpr = (vec && (rl == 4 || rl == 8 || rl == 16)) ? rl : 32;
    const int groups = 32 / lpr, grp = lane / lpr, gl = lane - grp * lpr;
    for (int c0 = 0; c0 < C; c0 += 128) {
        const int c = c0 + gl * 4;
        const bool act = c < C;
        float bv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        int bh[4] = {1 << 30, 1 << 30, 1 << 30, 1 << 30}, bi[4] = {-1, -1, -1, -1};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            for (int l = 0; l < 32; l += 4 * groups) {
                int id[4];
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) id[u] = __shfl_sync(0xffffffffu, idx[j], l + u * groups + grp);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (id[u] >= 0 && act) {
                        const float* r = xb + (int64_t)id[u] * C;
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
                    const int h = j * 32 + l + u * groups + grp;
                    amax_take(bv[0], bh[0], bi[0], v[u].x, h, id[u]);
                    amax_take(bv[1], bh[1], bi[1], v[u].y, h, id[u]);
                    amax_take(bv[2], bh[2], bi[2], v[u].z, h, id[u]);
                    amax_take(bv[3], bh[3], bi[3], v[u].w, h, id[u]);
                }
            }
        }
        for (int off = lpr; off < 32; off <<= 1) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv[e], off);
                const int oh = __shfl_xor_sync(0xffffffffu, bh[e], off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi[e], off);
                amax_take(bv[e], bh[e], bi[e], ov, oh, oi);
            }
        }
        if (act && grp == 0) {
            const float* g = dy + m * C + c;
            float* db = dx + frame * Ns * C + c;
            if (bi[0] >= 0) atomicAdd(db + (int64_t)bi[0] * C, __ldg(g));
            if (c + 1 < C && bi[1] >= 0) atomicAdd(db + (int64_t)bi[1] * C + 1, __ldg(g + 1));
            if (c + 2 < C && bi[2] >= 0) atomicAdd(db + (int64_t)bi[2] * C + 2, __ldg(g + 2));
            if (c + 3 < C && bi[3] >= 0) atomicAdd(db + (int64_t)bi[3] * C + 3, __ldg(g + 3));
        }
    }
}

// KPConv aggregate backward: dfeats[nbr[m,h], c] += sum_k w[m,h,k] * dagg[m,k,c]  (same influence arithmetic as forward)
template <int VEC>
struct BVec;
template <>
struct BVec<1> { using T = float; };
template <>
struct BVec<2> { using T = float2; };
template <>
struct BVec<4> { using T = float4; };

// Mirror of the forward kernel (kpconv.cu): one warp per query point.  (1) the 128 neighbour ids / packed points are loaded
// 4 per lane and the ones inside the kernel's reach are compacted into shared memory; (2) the (neighbour, kernel point)
// pairs are evaluated 32 at a time, neighbour-major; (3) for every pair with a non-zero influence the warp accumulates
// w * dagg[m, k, :] (a row of THIS query: coalesced, L1-resident), and when the neighbour changes the finished sum goes to
// dfeats[neighbour] with one vector atomic per lane (red.global.add.v4.f32).
template <int VEC, int NCH>
__global__ void __launch_bounds__(128)
kpconv_aggregate_bwd_kernel(const float* __restrict__ dagg, int C, const float4* __restrict__ s_packed,
                            const float* __restrict__ q_points, const int64_t* __restrict__ nbr, int H, int64_t Mq,
                            int64_t Ns, int64_t total_q, const float* __restrict__ kernel_points, int K, float sigma,
                            float reach2, float* __restrict__ dfeats) {
    __shared__ float skp[32 * 3];
    __shared__ float4 snear[4][128];
    if (threadIdx.x < K * 3) skp[threadIdx.x] = kernel_points[threadIdx.x];
    __syncthreads();
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    if (m >= total_q) return;
    const int64_t frame = m / Mq;
    const float4* sp = s_packed + frame * Ns;
    float* db = dfeats + frame * Ns * C;
    const float qx = __ldg(q_points + m * 3), qy = __ldg(q_points + m * 3 + 1), qz = __ldg(q_points + m * 3 + 2);
    float4* near = snear[wib];
    int n_near = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int h = j * 32 + lane;
        const int64_t id = (h < H) ? __ldcs(nbr + m * H + h) : Ns;
        bool is_near = false;
        float rx = 0.f, ry = 0.f, rz = 0.f;
        if (id >= 0 && id < Ns) {
            const float4 p = __ldg(sp + id);
            rx = p.x - qx;
            ry = p.y - qy;
            rz = p.z - qz;
            is_near = rx * rx + ry * ry + rz * rz <= reach2;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, is_near);
        if (is_near) near[n_near + __popc(bal & ((1u << lane) - 1u))] = make_float4(rx, ry, rz, __int_as_float((int)id));
        n_near += __popc(bal);
    }
    __syncwarp();

    using V = typename BVec<VEC>::T;
    const float* drow = dagg + m * (int64_t)K * C;
    const bool lane_active = (lane * VEC) < C;
    float acc[NCH][VEC];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[c][v] = 0.0f;

    auto flush = [&](int row) {  // dfeats[row] += acc; acc = 0
        if (lane_active) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float* dst = db + (int64_t)row * C + (c * 32 + lane) * VEC;
                if constexpr (VEC == 4) {
                    atomicAdd(reinterpret_cast<float4*>(dst), make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]));
                } else if constexpr (VEC == 2) {
                    atomicAdd(reinterpret_cast<float2*>(dst), make_float2(acc[c][0], acc[c][1]));
                } else {
                    atomicAdd(dst, acc[c][0]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[c][v] = 0.0f;
    };

    int cur_j = -1, cur_row = 0;
    const int total_pairs = n_near * K;
    for (int base = 0; base < total_pairs; base += 32) {
        const int pidx = base + lane;
        float w = 0.0f;
        int pj = 0, pk = 0, prow = 0;
        if (pidx < total_pairs) {
            pj = pidx / K;
            pk = pidx - pj * K;
            const float4 nb = near[pj];
            prow = __float_as_int(nb.w);
            const float dx = nb.x - skp[pk * 3 + 0], dy = nb.y - skp[pk * 3 + 1], dz = nb.z - skp[pk * 3 + 2];
            const float sq = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            w = fmaxf(__fsub_rn(1.0f, __fdiv_rn(__fsqrt_rn(sq), sigma)), 0.0f);
        }
        unsigned mask = __ballot_sync(0xffffffffu, w > 0.0f);
        while (mask) {
            const int l = __ffs(mask) - 1;
            mask &= mask - 1;
            const int ej = __shfl_sync(0xffffffffu, pj, l);
            const int ek = __shfl_sync(0xffffffffu, pk, l);
            const int er = __shfl_sync(0xffffffffu, prow, l);
            const float ew = __shfl_sync(0xffffffffu, w, l);
            if (ej != cur_j) {
                if (cur_j >= 0) flush(cur_row);
                cur_j = ej;
                cur_row = er;
            }
            if (lane_active) {
                const float* row = drow + (int64_t)ek * C;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const V f = __ldg(reinterpret_cast<const V*>(row + (c * 32 + lane) * VEC));
                    const float* fv = reinterpret_cast<const float*>(&f);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[c][v] = fmaf(ew, fv[v], acc[c][v]);
                }
            }
        }
    }
    if (cur_j >= 0) flush(cur_row);
}

// ------------------------------------------------------------------------------------------ image helpers backward
// dy [B,2H,2W,C1+C2] -> dx1 [B,H,W,C1] (pre-zeroed, atomics), dx2 [B,2H,2W,C2]
__global__ void __launch_bounds__(256)
upsample2x_cat_bwd_kernel(const float* __restrict__ dy, int B, int H, int W, int C1, int C2, float* __restrict__ dx1,
                          float* __restrict__ dx2) {
    const int Ho = 2 * H, Wo = 2 * W, Ct = C1 + C2;
    const int64_t total = (int64_t)B * Ho * Wo * Ct;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % Ct);
        int64_t p = t / Ct;
        const int wo = (int)(p % Wo);
        p /= Wo;
        const int ho = (int)(p % Ho);
        const int b = (int)(p / Ho);
        const float g = __ldg(dy + t);
        if (c >= C1) {
            dx2[(((int64_t)b * Ho + ho) * Wo + wo) * C2 + (c - C1)] = g;
        } else {
            float sh = ((float)ho + 0.5f) * 0.5f - 0.5f, sw = ((float)wo + 0.5f) * 0.5f - 0.5f;
            if (sh < 0.f) sh = 0.f;
            if (sw < 0.f) sw = 0.f;
            const int h0 = (int)sh, w0 = (int)sw;
            const int h1 = h0 + (h0 < H - 1 ? 1 : 0), w1 = w0 + (w0 < W - 1 ? 1 : 0);
            const float lh1 = sh - (float)h0, lw1 = sw - (float)w0, lh0 = 1.0f - lh1, lw0 = 1.0f - lw1;
            float* xb = dx1 + (int64_t)b * H * W * C1;
            atomicAdd(xb + ((int64_t)h0 * W + w0) * C1 + c, g * lh0 * lw0);
            atomicAdd(xb + ((int64_t)h0 * W + w1) * C1 + c, g * lh0 * lw1);
            atomicAdd(xb + ((int64_t)h1 * W + w0) * C1 + c, g * lh1 * lw0);
            atomicAdd(xb + ((int64_t)h1 * W + w1) * C1 + c, g * lh1 * lw1);
        }
    }
}
__global__ void __launch_bounds__(256)
maxpool2d_3x3s2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int B, int H, int W, int C,
                           int Ho, int Wo, float* __restrict__ dx) {
    const int64_t total = (int64_t)B * Ho * Wo * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        int64_t p = t / C;
        const int wo = (int)(p % Wo);
        p /= Wo;
        const int ho = (int)(p % Ho);
        const int b = (int)(p / Ho);
        float best = -INFINITY;
        int64_t arg = -1;
        for (int dh = 0; dh < 3; ++dh) {
            const int hi = ho * 2 + dh - 1;
            if (hi < 0 || hi >= H) continue;
            for (int dw = 0; dw < 3; ++dw) {
                const int wi = wo * 2 + dw - 1;
                if (wi < 0 || wi >= W) continue;
                const int64_t i = (((int64_t)b * H + hi) * W + wi) * C + c;
                const float v = __ldg(x + i);
                if (v > best) {
                    best = v;
                    arg = i;
                }
            }
        }
        if (arg >= 0) atomicAdd(dx + arg, __ldg(dy + t));
    }
}
// zero-insertion: y[b, 2h, 2w, c] = x[b, h, w, c], other positions 0  (input gradient of stride-2 convolutions)
__global__ void __launch_bounds__(256)
dilate2_nhwc_kernel(const float* __restrict__ x, int B, int H, int W, int C, float* __restrict__ y) {
    const int64_t total = (int64_t)B * 2 * H * 2 * W * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        int64_t p = t / C;
        const int wo = (int)(p % (2 * W));
        p /= 2 * W;
        const int ho = (int)(p % (2 * H));
        const int b = (int)(p / (2 * H));
        y[t] = ((ho | wo) & 1) ? 0.0f : __ldg(x + (((int64_t)b * H + (ho >> 1)) * W + (wo >> 1)) * C + c);
    }
}
// dmap[b, top+dy, left+dx, c] += dpatch[i, c, dy, dx]
__global__ void __launch_bounds__(256)
extract_patch_bwd_kernel(const float* __restrict__ dpatch, int H, int W, int C, int b, const float* __restrict__ centers,
                         int64_t n, float* __restrict__ dmap, int frames = 1) {
    const int64_t per = n * C * 16, total = per * frames;
    const float* centers0 = centers;
    const float* dpatch0 = dpatch;
    const int b0 = b;
    for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(t0 / per);      // batched form: frame f reads centres [f,2,n], dpatch [f,n,C,16], image b0 + f
        const int64_t t = t0 - (int64_t)f * per;
        centers = centers0 + (int64_t)f * 2 * n;
        dpatch = dpatch0 + (int64_t)f * per;
        b = b0 + f;
        const int c = (int)(t % C);
        const int64_t r = t / C;
        const int pix = (int)(r % 16);
        const int64_t i = r / 16;
        const int left = (int)floorf(centers[i] - 2.0f), top = (int)floorf(centers[n + i] - 2.0f);
        const int yy = top + (pix >> 2), xx = left + (pix & 3);
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        atomicAdd(dmap + (((int64_t)b * H + yy) * W + xx) * C + c, __ldg(dpatch + (i * C + c) * 16 + pix));
    }
}

// ------------------------------------------------------------------------------------------ TN contraction
// C[Mo, No] (+)= sum_r A[r, m] * B[r, n]: both operands are read along their contiguous (channel) axis, so the weight
// gradients dW = dY^T X of linears (r = rows) and of convolutions (r = output pixels, B = tap-shifted NHWC input) need no
// transposes.  Split over r into gridDim.z partial results (deterministic second-stage sum).
struct DenseB {
    const float* X;
    int64_t ldx;
    int64_t R;
    int No;
    __device__ __forceinline__ float4 load4(int64_t r, int n) const {
        if (r < R && n < No) return __ldg(reinterpret_cast<const float4*>(X + r * ldx + n));
        return make_float4(0.f, 0.f, 0.f, 0.f);
    }
};
struct ConvTapB {  // column n = (tap, ci) of the im2col view; row r = output pixel
    const float* x;
    int B, H, W, Cin, KH, KW, stride, pad, Ho, Wo;
    int64_t R;
    int No;
    __device__ __forceinline__ float4 load4(int64_t r, int n) const {
        if (r >= R || n >= No) return make_float4(0.f, 0.f, 0.f, 0.f);
        const int wo = (int)(r % Wo);
        const int64_t t = r / Wo;
        const int ho = (int)(t % Ho), b = (int)(t / Ho);
        const int ci = n % Cin, tap = n / Cin;
        const int kw = tap % KW, kh = tap / KW;
        const int hi = ho * stride + kh - pad, wi = wo * stride + kw - pad;
        if (hi < 0 || hi >= H || wi < 0 || wi >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
        return __ldg(reinterpret_cast<const float4*>(x + (((int64_t)b * H + hi) * W + wi) * Cin + ci));
    }
};
constexpr int TN_BM = 64, TN_BN = 64, TN_BK = 16;
template <class BLoader>
__global__ void __launch_bounds__(256)
gemm_tn_kernel(const float* __restrict__ A, int64_t lda, int Mo, BLoader bl, int No, int64_t R, int64_t r_per_split,
               float* __restrict__ part /* [splits][Mo][No] */) {
    __shared__ __align__(16) float As[TN_BK][TN_BM + 4];
    __shared__ __align__(16) float Bs[TN_BK][TN_BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * TN_BM, n0 = blockIdx.y * TN_BN;
    const int64_t rs = blockIdx.z * r_per_split;
    const int64_t re = (rs + r_per_split < R) ? rs + r_per_split : R;
    const int lr = tid >> 4;          // 0..15 : row inside the k-block
    const int lc = (tid & 15) * 4;    // 0..60 : column quad
    const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads, 4 x 4 micro tile
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int64_t r0 = rs; r0 < re; r0 += TN_BK) {
        const int64_t r = r0 + lr;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (r < re) {
            if (m0 + lc < Mo) a = __ldg(reinterpret_cast<const float4*>(A + r * lda + m0 + lc));
            b = bl.load4(r, n0 + lc);
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&As[lr][lc]) = a;
        *reinterpret_cast<float4*>(&Bs[lr][lc]) = b;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TN_BK; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }
    float* o = part + (int64_t)blockIdx.z * Mo * No;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= Mo) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < No) o[(int64_t)m * No + n] = acc[i][j];
        }
    }
}
__global__ void __launch_bounds__(256)
split_sum_kernel(const float* __restrict__ part, int splits, int64_t n, float* __restrict__ out, int accumulate) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
#pragma unroll 8
        for (int k = 0; k < splits; ++k) s += (double)part[(int64_t)k * n + i];
        out[i] = (accumulate ? out[i] : 0.0f) + (float)s;
    }
}

// ------------------------------------------------------------------------------------------ attention (SIMT) bwd
constexpr int AB_D = 32, AB_T = 128, AB_TK = 64;
// forward twin that also writes the log-sum-exp of every (row, head)
__global__ void __launch_bounds__(AB_T)
attention_fwd_lse_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                         int64_t L, int64_t S, int heads, float scale, float* __restrict__ out,
                         float* __restrict__ lse) {
    __shared__ __align__(16) float Ks[AB_TK][AB_D];
    __shared__ __align__(16) float Vs[AB_TK][AB_D];
    const int head = blockIdx.y, frame = blockIdx.z;
    const int64_t ld = (int64_t)heads * AB_D;
    const int64_t row = (int64_t)blockIdx.x * AB_T + threadIdx.x;
    const bool active = row < L;
    const float* qp = q + ((int64_t)frame * L + (active ? row : 0)) * ld + head * AB_D;
    const float* kb = k + ((int64_t)frame * S) * ld + head * AB_D;
    const float* vb = v + ((int64_t)frame * S) * ld + head * AB_D;
    float qr[AB_D], o[AB_D];
#pragma unroll
    for (int d = 0; d < AB_D; ++d) {
        qr[d] = qp[d] * scale;
        o[d] = 0.f;
    }
    float mrun = -INFINITY, lrun = 0.f;
    for (int64_t s0 = 0; s0 < S; s0 += AB_TK) {
        __syncthreads();
        for (int t = threadIdx.x; t < AB_TK * AB_D; t += AB_T) {
            const int j = t / AB_D, d = t % AB_D;
            Ks[j][d] = (s0 + j < S) ? kb[(s0 + j) * ld + d] : 0.f;
            Vs[j][d] = (s0 + j < S) ? vb[(s0 + j) * ld + d] : 0.f;
        }
        __syncthreads();
        const int nk = (int)((S - s0) < AB_TK ? (S - s0) : AB_TK);
        for (int j = 0; j < nk; ++j) {
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < AB_D; ++d) a = fmaf(qr[d], Ks[j][d], a);
            const float mnew = fmaxf(mrun, a);
            const float corr = expf(mrun - mnew), p = expf(a - mnew);
            lrun = lrun * corr + p;
#pragma unroll
            for (int d = 0; d < AB_D; ++d) o[d] = fmaf(p, Vs[j][d], o[d] * corr);
            mrun = mnew;
        }
    }
    if (active) {
        float* op = out + ((int64_t)frame * L + row) * ld + head * AB_D;
        const float inv = 1.0f / lrun;
#pragma unroll
        for (int d = 0; d < AB_D; ++d) op[d] = o[d] * inv;
        lse[((int64_t)frame * L + row) * heads + head] = mrun + logf(lrun);
    }
}
// dQ (thread per query row) and D_i = dO_i . O_i
__global__ void __launch_bounds__(AB_T)
attention_bwd_dq_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                        const float* __restrict__ out, const float* __restrict__ dout, const float* __restrict__ lse,
                        int64_t L, int64_t S, int heads, float scale, float* __restrict__ dq, float* __restrict__ dsum) {
    __shared__ __align__(16) float Ks[AB_TK][AB_D];
    __shared__ __align__(16) float Vs[AB_TK][AB_D];
    const int head = blockIdx.y, frame = blockIdx.z;
    const int64_t ld = (int64_t)heads * AB_D;
    const int64_t row = (int64_t)blockIdx.x * AB_T + threadIdx.x;
    const bool active = row < L;
    const int64_t ro = ((int64_t)frame * L + (active ? row : 0)) * ld + head * AB_D;
    const float* kb = k + ((int64_t)frame * S) * ld + head * AB_D;
    const float* vb = v + ((int64_t)frame * S) * ld + head * AB_D;
    float qr[AB_D], go[AB_D], acc[AB_D];
    float D = 0.f;
#pragma unroll
    for (int d = 0; d < AB_D; ++d) {
        qr[d] = q[ro + d] * scale;
        go[d] = dout[ro + d];
        D = fmaf(go[d], out[ro + d], D);
        acc[d] = 0.f;
    }
    const float l = lse[((int64_t)frame * L + (active ? row : 0)) * heads + head];
    for (int64_t s0 = 0; s0 < S; s0 += AB_TK) {
        __syncthreads();
        for (int t = threadIdx.x; t < AB_TK * AB_D; t += AB_T) {
            const int j = t / AB_D, d = t % AB_D;
            Ks[j][d] = (s0 + j < S) ? kb[(s0 + j) * ld + d] : 0.f;
            Vs[j][d] = (s0 + j < S) ? vb[(s0 + j) * ld + d] : 0.f;
        }
        __syncthreads();
        const int nk = (int)((S - s0) < AB_TK ? (S - s0) : AB_TK);
        for (int j = 0; j < nk; ++j) {
            float a = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < AB_D; ++d) {
                a = fmaf(qr[d], Ks[j][d], a);
                dp = fmaf(go[d], Vs[j][d], dp);
            }
            const float ds = expf(a - l) * (dp - D) * scale;
#pragma unroll
            for (int d = 0; d < AB_D; ++d) acc[d] = fmaf(ds, Ks[j][d], acc[d]);
        }
    }
    if (active) {
#pragma unroll
        for (int d = 0; d < AB_D; ++d) dq[ro + d] = acc[d];
        dsum[((int64_t)frame * L + row) * heads + head] = D;
    }
}
// dK, dV (thread per key row), queries streamed through shared memory
__global__ void __launch_bounds__(AB_T)
attention_bwd_dkv_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                         const float* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ dsum,
                         int64_t L, int64_t S, int heads, float scale, float* __restrict__ dk, float* __restrict__ dv) {
    __shared__ __align__(16) float Qs[AB_TK][AB_D];
    __shared__ __align__(16) float Gs[AB_TK][AB_D];
    __shared__ float Ls[AB_TK], Ds[AB_TK];
    const int head = blockIdx.y, frame = blockIdx.z;
    const int64_t ld = (int64_t)heads * AB_D;
    const int64_t key = (int64_t)blockIdx.x * AB_T + threadIdx.x;
    const bool active = key < S;
    const int64_t ko = ((int64_t)frame * S + (active ? key : 0)) * ld + head * AB_D;
    float kr[AB_D], vr[AB_D], ak[AB_D], av[AB_D];
#pragma unroll
    for (int d = 0; d < AB_D; ++d) {
        kr[d] = k[ko + d];
        vr[d] = v[ko + d];
        ak[d] = 0.f;
        av[d] = 0.f;
    }
    const float* qb = q + ((int64_t)frame * L) * ld + head * AB_D;
    const float* gb = dout + ((int64_t)frame * L) * ld + head * AB_D;
    for (int64_t l0 = 0; l0 < L; l0 += AB_TK) {
        __syncthreads();
        for (int t = threadIdx.x; t < AB_TK * AB_D; t += AB_T) {
            const int j = t / AB_D, d = t % AB_D;
            Qs[j][d] = (l0 + j < L) ? qb[(l0 + j) * ld + d] : 0.f;
            Gs[j][d] = (l0 + j < L) ? gb[(l0 + j) * ld + d] : 0.f;
        }
        if (threadIdx.x < AB_TK) {
            const int64_t i = l0 + threadIdx.x;
            Ls[threadIdx.x] = (i < L) ? lse[((int64_t)frame * L + i) * heads + head] : INFINITY;
            Ds[threadIdx.x] = (i < L) ? dsum[((int64_t)frame * L + i) * heads + head] : 0.f;
        }
        __syncthreads();
        const int nq = (int)((L - l0) < AB_TK ? (L - l0) : AB_TK);
        for (int j = 0; j < nq; ++j) {
            float a = 0.f, dp = 0.f;
#pragma unroll
            for (int d = 0; d < AB_D; ++d) {
                a = fmaf(Qs[j][d], kr[d], a);
                dp = fmaf(Gs[j][d], vr[d], dp);
            }
            const float p = expf(a * scale - Ls[j]);
            const float ds = p * (dp - Ds[j]) * scale;
#pragma unroll
            for (int d = 0; d < AB_D; ++d) {
                av[d] = fmaf(p, Gs[j][d], av[d]);
                ak[d] = fmaf(ds, Qs[j][d], ak[d]);
            }
        }
    }
    if (active) {
#pragma unroll
        for (int d = 0; d < AB_D; ++d) {
            dk[ko + d] = ak[d];
            dv[ko + d] = av[d];
        }
    }
}

// ------------------------------------------------------------------------------------------ optimiser
// Adam (torch.optim.Adam defaults, reference train.py:154): in-place update of p, m, v
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            float lr, float b1, float b2, float eps, float bc1, float bc2, float grad_scale) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i] * grad_scale;
        const float mi = b1 * m[i] + (1.0f - b1) * gi;
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
    }
}

}  // namespace cofi

using namespace cofi;
#define ST (cudaStream_t) stream

extern "C" int cofi_act_bwd(const float* dy, const float* y, int64_t n, int act, float* dx, void* stream) {
    COFI_REQUIRE(dy && y && dx && n >= 0, "cofi_act_bwd: bad argument");
    if (n == 0) return COFI_OK;
    act_bwd_kernel<<<ew_blocks_b(n, 256), 256, 0, ST>>>(dy, y, n, act, dx);
    return check_launch("cofi_act_bwd");
}

extern "C" int cofi_rowscale(const float* x, int64_t rows, int C, const float* rowdiv, float* y, void* stream) {
    COFI_REQUIRE(x && y && rowdiv && rows >= 0 && C > 0, "cofi_rowscale: bad argument");
    if (rows == 0) return COFI_OK;
    rowscale_kernel<<<ew_blocks_b(rows * C, 256), 256, 0, ST>>>(x, rows, C, rowdiv, y);
    return check_launch("cofi_rowscale");
}

extern "C" int64_t cofi_colsum_workspace(int C) { return (int64_t)kColChunks * C * sizeof(double); }
extern "C" int cofi_colsum(const float* x, int64_t ldx, int64_t rows, int C, float* out, int accumulate, void* work,
                           void* stream) {
    COFI_REQUIRE(x && out && work && rows >= 0 && C > 0, "cofi_colsum: bad argument");
    COFI_REQUIRE(((uintptr_t)work % 8) == 0, "cofi_colsum: workspace alignment");
    int64_t want = (rows + 7) / 8;
    const int chunks = (int)(want < 1 ? 1 : (want > kColChunks ? kColChunks : want));
    if (C % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)x % 16) == 0) {
        dim3 grid4((unsigned)ceil_div(C, 128), (unsigned)chunks);
        colsum_partial_v4_kernel<<<grid4, 256, 0, ST>>>(x, ldx, rows, C, (double*)work);
    } else {
        dim3 grid((unsigned)ceil_div(C, 32), (unsigned)chunks);
        colsum_partial_kernel<<<grid, 256, 0, ST>>>(x, ldx, rows, C, (double*)work);
    }
    int rc = check_launch("cofi_colsum(partial)");
    if (rc) return rc;
    colsum_final_kernel<<<(unsigned)ceil_div(C, 32), 256, 0, ST>>>((const double*)work, chunks, C, out, accumulate);
    return check_launch("cofi_colsum(final)");
}

extern "C" int64_t cofi_norm_rows_bwd_workspace(int frames, int C) {
    return (int64_t)frames * kBwdChunks * C * 2 * sizeof(double) + (int64_t)frames * C * 2 * sizeof(double) +
           (int64_t)frames * C * sizeof(float2) + 256;
}
extern "C" int cofi_norm_rows_bwd(const float* x, const float* dy, const float* y, int64_t R, int C, int frames, int G,
                                  const float* mean_rstd, const float* gamma, int act, float* dx, float* dres,
                                  float* dgamma, float* dbeta, int accumulate, void* work, void* stream) {
    COFI_REQUIRE(x && dy && y && mean_rstd && dx && work, "cofi_norm_rows_bwd: null pointer");
    COFI_REQUIRE(R > 0 && C > 0 && frames > 0 && G > 0 && C % G == 0, "cofi_norm_rows_bwd: bad shape");
    COFI_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "cofi_norm_rows_bwd: dgamma and dbeta go together");
    COFI_REQUIRE(((uintptr_t)work % 16) == 0, "cofi_norm_rows_bwd: workspace alignment");
    double* part = (double*)work;
    double* ab = part + (int64_t)frames * kBwdChunks * C * 2;
    float2* s12 = (float2*)(ab + (int64_t)frames * C * 2);
    const float2* mr = (const float2*)mean_rstd;
    const bool v4 = C % 4 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)dy % 16) == 0 && ((uintptr_t)y % 16) == 0 &&
                    ((uintptr_t)dx % 16) == 0 && (!dres || ((uintptr_t)dres % 16) == 0);
    if (v4) {
        dim3 g1((unsigned)ceil_div(C, 128), kBwdChunks, frames);
        norm_bwd_stats_v4_kernel<<<g1, 256, 0, ST>>>(x, dy, y, R, C, G, mr, act, part);
    } else {
        dim3 g1((unsigned)ceil_div(C, 32), kBwdChunks, frames);
        norm_bwd_stats_kernel<<<g1, 256, 0, ST>>>(x, dy, y, R, C, G, mr, act, part);
    }
    int rc = check_launch("cofi_norm_rows_bwd(stats)");
    if (rc) return rc;
    norm_bwd_reduce_kernel<<<(unsigned)ceil_div((int64_t)frames * C, 128), 128, 0, ST>>>(part, C, frames * C, ab);
    rc = check_launch("cofi_norm_rows_bwd(reduce)");
    if (rc) return rc;
    const int tmax = frames * G > C ? frames * G : C;
    norm_bwd_finalize_kernel<<<(unsigned)ceil_div(tmax, 128), 128, 0, ST>>>(ab, C, G, frames, gamma, s12, dgamma, dbeta,
                                                                          accumulate);
    rc = check_launch("cofi_norm_rows_bwd(finalize)");
    if (rc) return rc;
    const int64_t rows = R * frames;
    if (v4) {
        // about 8 resident CTAs per SM over the whole grid, at least 8 rows per CTA
        const int64_t cb = ceil_div(C, 128);
        int64_t chunks = ceil_div((int64_t)148 * 8, cb * frames);
        if (chunks > ceil_div(R, 8)) chunks = ceil_div(R, 8);
        if (chunks < 1) chunks = 1;
        dim3 g2((unsigned)cb, (unsigned)chunks, frames);
        norm_bwd_apply_v4_kernel<<<g2, 256, 0, ST>>>(x, dy, y, R, C, G, mr, s12, gamma, act, dx, dres);
    } else {
        norm_bwd_apply_kernel<<<ew_blocks_b(rows * C, 256), 256, 0, ST>>>(x, dy, y, R, C, G, rows, mr, s12, gamma, act, dx, dres);
    }
    return check_launch("cofi_norm_rows_bwd(apply)");
}

extern "C" int cofi_layer_norm_bwd(const float* x, const float* dy, int64_t rows, int C, const float* gamma,
                                   const float* beta, float eps, int act, float* dx, float* t1, float* t2, void* stream) {
    COFI_REQUIRE(x && dy && gamma && beta && dx && t1 && t2 && rows >= 0 && C > 0, "cofi_layer_norm_bwd: bad argument");
    COFI_REQUIRE(act == COFI_ACT_NONE || act == COFI_ACT_RELU, "cofi_layer_norm_bwd: activation must be none or relu");
    if (rows == 0) return COFI_OK;
    layer_norm_bwd_kernel<<<(unsigned)ceil_div(rows, 4), 128, 0, ST>>>(x, dy, rows, C, gamma, beta, eps, act, dx, t1, t2);
    return check_launch("cofi_layer_norm_bwd");
}

extern "C" int cofi_l2norm_bwd(const float* x, const float* dy, int64_t rows, int C, float* dx, void* stream) {
    COFI_REQUIRE(x && dy && dx && rows >= 0 && C > 0, "cofi_l2norm_bwd: bad argument");
    if (rows == 0) return COFI_OK;
    l2norm_bwd_kernel<<<(unsigned)ceil_div(rows, 4), 128, 0, ST>>>(x, dy, rows, C, dx);
    return check_launch("cofi_l2norm_bwd");
}

extern "C" int64_t cofi_colnorm_bwd_workspace(int frames, int C) {
    return (int64_t)frames * kBwdChunks * C * 2 * sizeof(double) + (int64_t)frames * C * 2 * sizeof(double);
}
extern "C" int cofi_colnorm_bwd(const float* x, const float* dy, int64_t L, int C, int frames, float* dx, void* work,
                                void* stream) {
    COFI_REQUIRE(x && dy && dx && work && L > 0 && C > 0 && frames > 0, "cofi_colnorm_bwd: bad argument");
    double* part = (double*)work;
    double* ab = part + (int64_t)frames * kBwdChunks * C * 2;
    dim3 g1((unsigned)ceil_div(C, 32), kBwdChunks, frames);
    colnorm_bwd_partial_kernel<<<g1, 256, 0, ST>>>(x, dy, L, C, part);
    int rc = check_launch("cofi_colnorm_bwd(partial)");
    if (rc) return rc;
    norm_bwd_reduce_kernel<<<(unsigned)ceil_div((int64_t)frames * C, 128), 128, 0, ST>>>(part, C, frames * C, ab);
    rc = check_launch("cofi_colnorm_bwd(reduce)");
    if (rc) return rc;
    const int64_t rows = L * frames;
    colnorm_bwd_apply_kernel<<<ew_blocks_b(rows * C, 256), 256, 0, ST>>>(x, dy, L, C, rows, ab, dx);
    return check_launch("cofi_colnorm_bwd(apply)");
}

extern "C" int cofi_scatter_add_rows(const float* dy, int64_t ldy, int C, const int64_t* idx, int64_t idx_stride,
                                     int64_t Mq, int64_t Ns, int frames, float* dx, void* stream) {
    COFI_REQUIRE(dy && idx && dx && C > 0 && frames > 0 && ldy >= C, "cofi_scatter_add_rows: bad argument");
    const int64_t total = Mq * frames;
    if (total == 0) return COFI_OK;
    scatter_add_rows_kernel<<<ew_blocks_b(total * C, 256), 256, 0, ST>>>(dy, ldy, C, idx, idx_stride, Mq, Ns, total, dx);
    return check_launch("cofi_scatter_add_rows");
}

extern "C" int cofi_maxpool_rows_bwd(const float* x, int C, const int64_t* nbr, int H, int64_t Mq, int64_t Ns, int frames,
                                     const float* dy, float* dx, void* stream) {
    COFI_REQUIRE(x && nbr && dy && dx && C > 0 && H > 0 && H <= 128 && frames > 0, "cofi_maxpool_rows_bwd: bad argument");
    const int64_t total = Mq * frames;
    if (total == 0) return COFI_OK;
    maxpool_rows_bwd_kernel<<<(unsigned)ceil_div(total, 4), 128, 0, ST>>>(x, C, nbr, H, Mq, Ns, total, dy, dx);
    return check_launch("cofi_maxpool_rows_bwd");
}

extern "C" int cofi_kpconv_aggregate_bwd(const float* dagg, int C, const float* s_packed, const float* q_points,
                                         const int64_t* nbr, int H, int64_t Mq, int64_t Ns, int frames,
                                         const float* kernel_points, int K, float sigma, float kp_reach, float* dfeats,
                                         void* stream) {
    COFI_REQUIRE(dagg && s_packed && q_points && nbr && kernel_points && dfeats, "cofi_kpconv_aggregate_bwd: null pointer");
    COFI_REQUIRE(H > 0 && K > 0 && K <= 32 && sigma > 0.0f && C > 0 && frames > 0, "cofi_kpconv_aggregate_bwd: bad argument");
    const int64_t total = Mq * frames;
    if (total == 0) return COFI_OK;
    const float reach = kp_reach > 0.0f ? (kp_reach + sigma) * 1.001f : 1e18f;
    COFI_REQUIRE(H <= 128, "cofi_kpconv_aggregate_bwd: H must be <= 128");
#define BLAUNCH(VEC, NCH)                                                                                                  \
    kpconv_aggregate_bwd_kernel<VEC, NCH><<<(unsigned)ceil_div(total, 4), 128, 0, ST>>>(                                   \
        dagg, C, (const float4*)s_packed, q_points, nbr, H, Mq, Ns, total, kernel_points, K, sigma, reach * reach, dfeats)
    if (C <= 32) {
        BLAUNCH(1, 1);
    } else if (C == 64) {
        BLAUNCH(2, 1);
    } else if (C % 128 == 0 && C <= 1024) {
        switch (C / 128) {
            case 1: BLAUNCH(4, 1); break;
            case 2: BLAUNCH(4, 2); break;
            case 3: BLAUNCH(4, 3); break;
            case 4: BLAUNCH(4, 4); break;
            case 8: BLAUNCH(4, 8); break;
            default: set_error("cofi_kpconv_aggregate_bwd: unsupported channel count C=%d", C); return COFI_EUNSUPPORTED;
        }
    } else {
        set_error("cofi_kpconv_aggregate_bwd: unsupported channel count C=%d", C);
        return COFI_EUNSUPPORTED;
    }
#undef BLAUNCH
    return check_launch("cofi_kpconv_aggregate_bwd");
}

extern "C" int cofi_upsample2x_cat_bwd(const float* dy, int B, int H, int W, int C1, int C2, float* dx1, float* dx2,
                                       void* stream) {
    COFI_REQUIRE(dy && dx1 && dx2 && B > 0 && H > 0 && W > 0 && C1 > 0 && C2 > 0, "cofi_upsample2x_cat_bwd: bad argument");
    const int64_t total = (int64_t)B * 4 * H * W * (C1 + C2);
    upsample2x_cat_bwd_kernel<<<ew_blocks_b(total, 256), 256, 0, ST>>>(dy, B, H, W, C1, C2, dx1, dx2);
    return check_launch("cofi_upsample2x_cat_bwd");
}

extern "C" int cofi_maxpool2d_3x3s2_bwd(const float* x, const float* dy, int B, int H, int W, int C, float* dx,
                                        void* stream) {
    COFI_REQUIRE(x && dy && dx && B > 0 && H > 0 && W > 0 && C > 0, "cofi_maxpool2d_3x3s2_bwd: bad argument");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    maxpool2d_3x3s2_bwd_kernel<<<ew_blocks_b((int64_t)B * Ho * Wo * C, 256), 256, 0, ST>>>(x, dy, B, H, W, C, Ho, Wo, dx);
    return check_launch("cofi_maxpool2d_3x3s2_bwd");
}

extern "C" int cofi_dilate2_nhwc(const float* x, int B, int H, int W, int C, float* y, void* stream) {
    COFI_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0, "cofi_dilate2_nhwc: bad argument");
    dilate2_nhwc_kernel<<<ew_blocks_b((int64_t)B * 4 * H * W * C, 256), 256, 0, ST>>>(x, B, H, W, C, y);
    return check_launch("cofi_dilate2_nhwc");
}

extern "C" int cofi_extract_patch_bwd(const float* dpatch, int H, int W, int C, int b, const float* centers, int64_t n,
                                      float* dmap, void* stream) {
    COFI_REQUIRE(dpatch && centers && dmap && H >= 4 && W >= 4 && C > 0 && n >= 0, "cofi_extract_patch_bwd: bad argument");
    if (n == 0) return COFI_OK;
    extract_patch_bwd_kernel<<<ew_blocks_b(n * C * 16, 256), 256, 0, ST>>>(dpatch, H, W, C, b, centers, n, dmap);
    return check_launch("cofi_extract_patch_bwd");
}

extern "C" int cofi_extract_patch_batched_bwd(const float* dpatch, int H, int W, int C, int frames, const float* centers,
                                              int64_t n, float* dmap, void* stream) {
    COFI_REQUIRE(dpatch && centers && dmap && H >= 4 && W >= 4 && C > 0 && frames > 0 && n >= 0,
                 "cofi_extract_patch_batched_bwd: bad argument");
    if (n == 0) return COFI_OK;
    extract_patch_bwd_kernel<<<ew_blocks_b(n * C * 16 * frames, 256), 256, 0, ST>>>(dpatch, H, W, C, 0, centers, n, dmap, frames);
    return check_launch("cofi_extract_patch_batched_bwd");
}

static int tn_splits(int64_t R) {
    int64_t s = R / 2048;
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    return (int)s;
}
namespace cofi { namespace tc {
int64_t tn_tc_per(int64_t R, int Mo, int No, int* splits);
int gemm_tn_tc(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t R, int Mo, int No, float* partial,
               int* splits_out, cudaStream_t st);
bool conv_wgrad_tc_ok(int Cin, int Cout, int Wo, int stride);
int conv_wgrad_tc(const float* x, int B, int H, int W, int Cin, const float* dy, int Cout, int KH, int KW, int stride, int pad,
                  int Ho, int Wo, float* partial, int* splits_out, cudaStream_t st);
}}
static int tn_splits_any(int64_t R, int Mo, int No) {  // workspace covers either engine
    int tcs = 1;
    cofi::tc::tn_tc_per(R, Mo, No, &tcs);
    const int s = tn_splits(R);
    return s > tcs ? s : tcs;
}
extern "C" int64_t cofi_gemm_tn_workspace(int64_t R, int Mo, int No) { return (int64_t)tn_splits_any(R, Mo, No) * Mo * No * 4; }

// C[Mo,No] (+)= A[R,Mo]^T B[R,No]   (weight gradient of a Linear: A = dY, B = X)
extern "C" int cofi_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t R, int Mo, int No,
                            int accumulate, int engine, void* work, void* stream) {
    COFI_REQUIRE(A && B && C && work && R > 0 && Mo > 0 && No > 0, "cofi_gemm_tn: bad argument");
    COFI_REQUIRE(Mo % 4 == 0 && No % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0, "cofi_gemm_tn: Mo, No, lda, ldb multiples of 4");
    COFI_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0, "cofi_gemm_tn: 16-byte alignment");
    if (engine == COFI_GEMM_TF32) {  // tcgen05, MN-major operands (gemm_tn_tc.cu)
        int sp = 1;
        int rc = cofi::tc::gemm_tn_tc(A, lda, B, ldb, R, Mo, No, (float*)work, &sp, ST);
        if (rc) return rc;
        split_sum_kernel<<<ew_blocks_b((int64_t)Mo * No, 256), 256, 0, ST>>>((const float*)work, sp, (int64_t)Mo * No, C, accumulate);
        return check_launch("cofi_gemm_tn(sum)");
    }
    const int splits = tn_splits(R);
    const int64_t per = ceil_div(ceil_div(R, splits), TN_BK) * TN_BK;
    DenseB bl{B, ldb, R, No};
    dim3 grid((unsigned)ceil_div(Mo, TN_BM), (unsigned)ceil_div(No, TN_BN), (unsigned)ceil_div(R, per));
    gemm_tn_kernel<DenseB><<<grid, 256, 0, ST>>>(A, lda, Mo, bl, No, R, per, (float*)work);
    int rc = check_launch("cofi_gemm_tn");
    if (rc) return rc;
    split_sum_kernel<<<ew_blocks_b((int64_t)Mo * No, 256), 256, 0, ST>>>((const float*)work, (int)grid.z, (int64_t)Mo * No, C,
                                                                         accumulate);
    return check_launch("cofi_gemm_tn(sum)");
}

extern "C" int64_t cofi_conv2d_wgrad_workspace(int B, int Ho, int Wo, int Cout, int KH, int KW, int Cin) {
    return (int64_t)tn_splits_any((int64_t)B * Ho * Wo, Cout, KH * KW * Cin) * Cout * KH * KW * Cin * 4;
}
// dw[Cout, KH*KW*Cin] (+)= sum over output pixels of dy[p, co] * x[p shifted by the tap, ci]
extern "C" int cofi_conv2d_wgrad_nhwc(const float* x, int B, int H, int W, int Cin, const float* dy, int Cout, int KH,
                                      int KW, int stride, int pad, float* dw, int accumulate, int engine, void* work,
                                      void* stream) {
    COFI_REQUIRE(x && dy && dw && work, "cofi_conv2d_wgrad_nhwc: null pointer");
    COFI_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0, "cofi_conv2d_wgrad_nhwc: channel counts must be multiples of 4");
    {
        const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
        if (engine == COFI_GEMM_TF32 && cofi::tc::conv_wgrad_tc_ok(Cin, Cout, Wo, stride)) {
            int sp = 1;
            int rc = cofi::tc::conv_wgrad_tc(x, B, H, W, Cin, dy, Cout, KH, KW, stride, pad, Ho, Wo, (float*)work, &sp, ST);
            if (rc) return rc;
            const int64_t n = (int64_t)Cout * KH * KW * Cin;
            split_sum_kernel<<<ew_blocks_b(n, 256), 256, 0, ST>>>((const float*)work, sp, n, dw, accumulate);
            return check_launch("cofi_conv2d_wgrad_nhwc(sum)");
        }
    }
    ConvTapB bl;
    bl.x = x; bl.B = B; bl.H = H; bl.W = W; bl.Cin = Cin; bl.KH = KH; bl.KW = KW; bl.stride = stride; bl.pad = pad;
    bl.Ho = (H + 2 * pad - KH) / stride + 1;
    bl.Wo = (W + 2 * pad - KW) / stride + 1;
    bl.R = (int64_t)B * bl.Ho * bl.Wo;
    bl.No = KH * KW * Cin;
    const int splits = tn_splits(bl.R);
    const int64_t per = ceil_div(ceil_div(bl.R, splits), TN_BK) * TN_BK;
    dim3 grid((unsigned)ceil_div(Cout, TN_BM), (unsigned)ceil_div(bl.No, TN_BN), (unsigned)ceil_div(bl.R, per));
    gemm_tn_kernel<ConvTapB><<<grid, 256, 0, ST>>>(dy, Cout, Cout, bl, bl.No, bl.R, per, (float*)work);
    int rc = check_launch("cofi_conv2d_wgrad_nhwc");
    if (rc) return rc;
    split_sum_kernel<<<ew_blocks_b((int64_t)Cout * bl.No, 256), 256, 0, ST>>>((const float*)work, (int)grid.z,
                                                                             (int64_t)Cout * bl.No, dw, accumulate);
    return check_launch("cofi_conv2d_wgrad_nhwc(sum)");
}

extern "C" int cofi_attention_fwd_lse(const float* q, const float* k, const float* v, int64_t L, int64_t S, int frames,
                                      int heads, int D, float scale, float* out, float* lse, void* stream) {
    COFI_REQUIRE(q && k && v && out && lse && D == AB_D && L > 0 && S > 0 && frames > 0 && heads > 0,
                 "cofi_attention_fwd_lse: bad argument (D must be 32)");
    dim3 grid((unsigned)ceil_div(L, AB_T), heads, frames);
    attention_fwd_lse_kernel<<<grid, AB_T, 0, ST>>>(q, k, v, L, S, heads, scale, out, lse);
    return check_launch("cofi_attention_fwd_lse");
}

extern "C" int cofi_attention_bwd(const float* q, const float* k, const float* v, const float* out, const float* dout,
                                  const float* lse, int64_t L, int64_t S, int frames, int heads, int D, float scale,
                                  float* dq, float* dk, float* dv, float* dsum_work, void* stream) {
    COFI_REQUIRE(q && k && v && out && dout && lse && dq && dk && dv && dsum_work && D == AB_D,
                 "cofi_attention_bwd: bad argument (D must be 32)");
    dim3 g1((unsigned)ceil_div(L, AB_T), heads, frames);
    attention_bwd_dq_kernel<<<g1, AB_T, 0, ST>>>(q, k, v, out, dout, lse, L, S, heads, scale, dq, dsum_work);
    int rc = check_launch("cofi_attention_bwd(dq)");
    if (rc) return rc;
    dim3 g2((unsigned)ceil_div(S, AB_T), heads, frames);
    attention_bwd_dkv_kernel<<<g2, AB_T, 0, ST>>>(q, k, v, dout, lse, dsum_work, L, S, heads, scale, dk, dv);
    return check_launch("cofi_attention_bwd(dkv)");
}

extern "C" int cofi_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                              float eps, int step, float grad_scale, void* stream) {
    COFI_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "cofi_adam_step: bad argument");
    if (n == 0) return COFI_OK;
    const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
    adam_kernel<<<ew_blocks_b(n, 256), 256, 0, ST>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2, grad_scale);
    return check_launch("cofi_adam_step");
}
