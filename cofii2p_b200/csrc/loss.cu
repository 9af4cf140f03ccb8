// loss.cu -- the reference's three training losses (model/loss.py:9-93) as fused forward + analytic-backward kernels,
// batched over the B stacked frames of a training step (one CTA per frame, deterministic reductions, no atomics in the
// loss values).  Each kernel gathers its supervised rows straight from the token-layout network outputs, evaluates the
// loss of every frame and leaves dLoss/dInput for the gathered rows in a compact buffer; cofi_scatter_scaled_rows then
// multiplies by the upstream gradient and scatters into the (zero-initialised) dense gradient of the network output.
//   desc_loss        (:69-93)  circle-style log-sum-exp loss over the n x n coarse descriptor distance matrix
//   overlap_loss     (:53-60)  binary cross-entropy of the super-point overlap scores (in-frustum -> 1, outside -> 0)
//   fine_circle_loss (:9-51)   circle loss (m = 0.2, gamma = 5) of each key point against its 4 x 4 pixel patch
// What replaces what: ~60 ATen launches per frame and per direction (broadcast product, masks, logsumexp x4, softplus,
// cosine_similarity, scatter, exp/log/sum chains and their autograd graph) become 3 + 3 launches per STEP.
#include "common.cuh"

namespace cofi {

__device__ __forceinline__ float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }   // F.softplus defaults
__device__ __forceinline__ float sigmoid_sp(float x) { return x > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-x)); }

// ------------------------------------------------------------------------------------------------ desc_loss
struct DescParams {
    const float* img_tok;
    const int64_t* pix;
    int64_t img_rows;
    const float* pc_tok;
    const int64_t* kpt;
    int64_t pc_rows;
    const float* mask;
    int n, C;
    float pos_margin, neg_margin, log_scale;
    float* loss;
    float* dists;
    float* d_img;
    float* d_pc;
};

__global__ void __launch_bounds__(256)
desc_loss_kernel(const DescParams p) {
    extern __shared__ float sm[];
    const int n = p.n, C = p.C, ldA = C + 1, ldD = n + 1;
    float* A = sm;                 // image rows   [n][C+1]
    float* P = A + n * ldA;        // point rows   [n][C+1]
    float* D = P + n * ldA;        // distances    [n][n+1]
    float* G = D + n * ldD;        // dLoss/dD     [n][n+1]
    float* Lpr = G + n * ldD;      // row / column log-sum-exps of the positive and negative terms
    float* Lnr = Lpr + n;
    float* Lpc = Lnr + n;
    float* Lnc = Lpc + n;
    float* red = Lnc + n;          // [n] per-row loss
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const float* mask = p.mask + (int64_t)f * n * n;
    for (int r = warp; r < n; r += nw) {
        const float* a = p.img_tok + ((int64_t)f * p.img_rows + p.pix[(int64_t)f * n + r]) * C;
        const float* b = p.pc_tok + ((int64_t)f * p.pc_rows + p.kpt[(int64_t)f * n + r]) * C;
        for (int c = lane; c < C; c += 32) {
            A[r * ldA + c] = __ldg(a + c);
            P[r * ldA + c] = __ldg(b + c);
        }
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += blockDim.x) {   // dists = 1 - sum_c img[c, i] * pc[c, j]          (:73)
        const int i = e / n, j = e - i * n;
        float dot = 0.0f;
        for (int c = 0; c < C; ++c) dot = fmaf(A[i * ldA + c], P[j * ldA + c], dot);
        const float d = 1.0f - dot;
        D[i * ldD + j] = d;
        if (p.dists) p.dists[(int64_t)f * n * n + e] = d;
    }
    __syncthreads();
    const float s = p.log_scale, pm = p.pos_margin, nm = p.neg_margin;
    // pos = dists - 1e5 * (1 - mask); term_p = s * (pos - pm) * max(0, pos - pm)      (:75-80)
    // neg = dists + 1e5 * mask;       term_n = s * (nm - neg) * max(0, nm - neg)      (:82-87)
    auto term_p = [&](float d, float m) { const float x = (d - 1e5f * (1.0f - m)) - pm; return s * x * fmaxf(x, 0.0f); };
    auto term_n = [&](float d, float m) { const float x = nm - (d + 1e5f * m); return s * x * fmaxf(x, 0.0f); };
    for (int pass = 0; pass < 2; ++pass) {            // pass 0: over rows (dim=-1), pass 1: over columns (dim=-2)
        for (int i = warp; i < n; i += nw) {
            float mp = -INFINITY, mn = -INFINITY;
            for (int j = lane; j < n; j += 32) {
                const float d = pass ? D[j * ldD + i] : D[i * ldD + j];
                const float m = pass ? mask[j * n + i] : mask[i * n + j];
                mp = fmaxf(mp, term_p(d, m));
                mn = fmaxf(mn, term_n(d, m));
            }
            mp = warp_max(mp);
            mn = warp_max(mn);
            float sp = 0.0f, sn = 0.0f;
            for (int j = lane; j < n; j += 32) {
                const float d = pass ? D[j * ldD + i] : D[i * ldD + j];
                const float m = pass ? mask[j * n + i] : mask[i * n + j];
                sp += expf(term_p(d, m) - mp);
                sn += expf(term_n(d, m) - mn);
            }
            sp = warp_sum(sp);
            sn = warp_sum(sn);
            if (lane == 0) {
                (pass ? Lpc : Lpr)[i] = mp + logf(sp);
                (pass ? Lnc : Lnr)[i] = mn + logf(sn);
            }
        }
    }
    __syncthreads();
    // loss_i = softplus(Lpr_i + Lnr_i) / s + softplus(Lpc_i + Lnc_i) / s, mean over i                 (:89-93)
    for (int i = tid; i < n; i += blockDim.x) red[i] = (softplus_f(Lpr[i] + Lnr[i]) + softplus_f(Lpc[i] + Lnc[i])) / s;
    __syncthreads();
    if (tid == 0) {
        float acc = 0.0f;
        for (int i = 0; i < n; ++i) acc += red[i];
        p.loss[f] = acc / (float)n;
    }
    if (!p.d_img) return;
    // dLoss/dD: softplus' = sigmoid, logsumexp' = softmax, d term_p / dD = s * max(0, pos - pm) (the weight is detached),
    // d term_n / dD = -s * max(0, nm - neg)
    const float inv = 1.0f / (s * (float)n);
    for (int e = tid; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e - i * n;
        const float d = D[i * ldD + j], m = mask[i * n + j];
        const float xp = (d - 1e5f * (1.0f - m)) - pm, xn = nm - (d + 1e5f * m);
        const float wp = fmaxf(xp, 0.0f), wn = fmaxf(xn, 0.0f);
        const float tp = s * xp * wp, tn = s * xn * wn;
        const float ga = sigmoid_sp(Lpr[i] + Lnr[i]) * inv, gb = sigmoid_sp(Lpc[j] + Lnc[j]) * inv;
        G[i * ldD + j] = ga * (expf(tp - Lpr[i]) * s * wp - expf(tn - Lnr[i]) * s * wn) +
                         gb * (expf(tp - Lpc[j]) * s * wp - expf(tn - Lnc[j]) * s * wn);
    }
    __syncthreads();
    // D = 1 - A P^T  =>  dA = -G P,  dP = -G^T A
    for (int e = tid; e < n * C; e += blockDim.x) {
        const int i = e / C, c = e - i * C;
        float da = 0.0f, dp = 0.0f;
        for (int j = 0; j < n; ++j) {
            da = fmaf(G[i * ldD + j], P[j * ldA + c], da);
            dp = fmaf(G[j * ldD + i], A[j * ldA + c], dp);
        }
        p.d_img[((int64_t)f * n + i) * C + c] = -da;
        p.d_pc[((int64_t)f * n + i) * C + c] = -dp;
    }
}

// --------------------------------------------------------------------------------------------- overlap_loss
// F.binary_cross_entropy: loss = -mean(y * max(log s, -100) + (1 - y) * max(log(1 - s), -100)),
// d/ds = (s - y) / max((1 - s) * s, 1e-12) / N     (ATen binary_cross_entropy_backward)
__global__ void __launch_bounds__(256)
overlap_loss_kernel(const float* __restrict__ score, const int64_t* __restrict__ idx, int64_t rows, int n_in, int n_out,
                    float* __restrict__ loss, float* __restrict__ d_score) {
    __shared__ float red[256];
    const int f = blockIdx.x, N = n_in + n_out;
    float acc = 0.0f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float s = __ldg(score + (int64_t)f * rows + idx[(int64_t)f * N + i]);
        const float y = i < n_in ? 1.0f : 0.0f;
        acc -= y * fmaxf(logf(s), -100.0f) + (1.0f - y) * fmaxf(log1pf(-s), -100.0f);
        if (d_score) d_score[(int64_t)f * N + i] = (s - y) / fmaxf((1.0f - s) * s, 1e-12f) / (float)N;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[f] = red[0] / (float)N;
}

// ----------------------------------------------------------------------------------------- fine_circle_loss
// One warp per key point.  sim_j = cosine(patch[:, j], y) for the 16 pixels j; the positive is pixel rel, the other 15 are
// negatives:  ap = relu(1 + m - sp), logit_p = -ap (sp - (1 - m)) gamma;  an_j = relu(sn_j + m), logit_n_j = an_j (sn_j - m) gamma;
// L = log(1 + sum_j exp(logit_n_j) * exp(logit_p)), loss = mean over key points (ap / an are detached weights).
constexpr int FC_MAXC = 128;

__global__ void __launch_bounds__(256)
fine_circle_loss_kernel(const float* __restrict__ patch, const float* __restrict__ fpc, const int64_t* __restrict__ rel,
                        int n, int C, float m, float gamma, float* __restrict__ loss, float* __restrict__ d_patch,
                        float* __restrict__ d_fpc, int32_t* __restrict__ bad) {
    __shared__ float s_y[8][FC_MAXC];
    __shared__ float s_g[8][16], s_inx[8][16], s_sim[8][16];
    __shared__ float s_L[8];
    const int f = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.0f;
    for (int i = warp; i < n; i += 8) {
        const int64_t row = (int64_t)f * n + i;
        const float* x = patch + row * C * 16;
        const float* y = fpc + row * C;
        int r = (int)rel[row];
        if (r < 0 || r > 15) {  // the reference's label[...] indexing raises here (model/loss.py:22)
            if (bad && lane == 0) *bad = 1;
            r = r < 0 ? 0 : 15;
        }
        float ny2 = 0.0f;
        for (int c = lane; c < C; c += 32) {
            const float v = __ldg(y + c);
            s_y[warp][c] = v;
            ny2 = fmaf(v, v, ny2);
        }
        ny2 = warp_sum(ny2);
        __syncwarp();
        // lanes (j, half): pixel j = lane & 15, channels [half * C/2, (half + 1) * C/2)
        const int j = lane & 15, c0 = (lane >> 4) * (C >> 1), c1 = c0 + (C >> 1);
        float dot = 0.0f, nx2 = 0.0f;
        for (int c = c0; c < c1; ++c) {
            const float v = __ldg(x + c * 16 + j);
            dot = fmaf(v, s_y[warp][c], dot);
            nx2 = fmaf(v, v, nx2);
        }
        dot += __shfl_xor_sync(0xffffffffu, dot, 16);
        nx2 += __shfl_xor_sync(0xffffffffu, nx2, 16);
        const float nx = fmaxf(sqrtf(nx2), 1e-8f), ny = fmaxf(sqrtf(ny2), 1e-8f);   // torch.cosine_similarity, eps = 1e-8
        const float sim = dot / (nx * ny);
        const bool is_pos = (j == r);
        // negatives
        const float an = fmaxf(sim + m, 0.0f);
        float en = is_pos ? 0.0f : expf(an * (sim - m) * gamma);
        float loss_n = en;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) loss_n += __shfl_xor_sync(0xffffffffu, loss_n, o);   // within each 16-lane half
        // positive (broadcast from lane r)
        const float sp = __shfl_sync(0xffffffffu, sim, r);
        const float ap = fmaxf(1.0f + m - sp, 0.0f);
        const float loss_p = expf(-ap * (sp - (1.0f - m)) * gamma);
        const float z = loss_n * loss_p;
        const float L = log1pf(z);
        acc += L;
        if (d_patch) {
            const float k = 1.0f / ((1.0f + z) * (float)n);                // dL/dz / n
            // g_j = d(mean L)/d sim_j
            const float g = is_pos ? k * z * (-ap * gamma) : k * loss_p * en * (an * gamma);
            if (lane < 16) {
                s_g[warp][j] = g;
                s_inx[warp][j] = 1.0f / nx;
                s_sim[warp][j] = sim;
            }
            __syncwarp();
            const float inx = 1.0f / nx, iny = 1.0f / ny;
            const bool x_live = nx2 > 1e-16f, y_live = ny2 > 1e-16f;  // below eps the clamped norm is a constant
            // d sim_j / d x_cj = y_c / (nx ny) - sim_j x_cj / nx^2
            for (int c = c0; c < c1; ++c) {
                const float v = __ldg(x + c * 16 + j);
                d_patch[row * C * 16 + c * 16 + j] = g * (s_y[warp][c] * inx * iny - (x_live ? sim * v * inx * inx : 0.0f));
            }
            // d sim_j / d y_c = x_cj / (nx ny) - sim_j y_c / ny^2, summed over j
            float gs = 0.0f;
            for (int jj = 0; jj < 16; ++jj) gs = fmaf(s_g[warp][jj], s_sim[warp][jj], gs);
            for (int c = lane; c < C; c += 32) {
                float a = 0.0f;
                for (int jj = 0; jj < 16; ++jj) a = fmaf(s_g[warp][jj] * s_inx[warp][jj], __ldg(x + c * 16 + jj), a);
                d_fpc[row * C + c] = a * iny - (y_live ? gs * s_y[warp][c] * iny * iny : 0.0f);
            }
        }
        __syncwarp();
    }
    if (lane == 0) s_L[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < 8; ++w) t += s_L[w];
        loss[f] = t / (float)n;
    }
}

// ------------------------------------------------------------------------------------ scaled row scatter
// dst[f * rows_dst + idx[f * R + r], :] += scale_f * src[f * R + r, :]   (idx NULL: dst[f * R + r, :] = scale_f * src[...]),
// scale_f = scale_host * scale_dev[f]
__global__ void __launch_bounds__(256)
scatter_scaled_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, int64_t R, int64_t rows_dst,
                           int frames, int C, const float* __restrict__ scale_dev, float scale_host,
                           float* __restrict__ dst) {
    const int64_t total = (int64_t)frames * R * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / C;
        const int c = (int)(t - row * C);
        const int64_t f = row / R;
        const float v = scale_host * (scale_dev ? __ldg(scale_dev + f) : 1.0f) * __ldg(src + t);
        if (idx) {
            atomicAdd(dst + (f * rows_dst + idx[row]) * C + c, v);
        } else {
            dst[t] = v;
        }
    }
}

}  // namespace cofi

using namespace cofi;

extern "C" int cofi_desc_loss(const float* img_tok, const int64_t* pix, int64_t img_rows, const float* pc_tok,
                              const int64_t* kpt, int64_t pc_rows, const float* mask, int n, int C, int frames,
                              float pos_margin, float neg_margin, float log_scale, float* loss, float* dists, float* d_img,
                              float* d_pc, void* stream) {
    COFI_REQUIRE(img_tok && pix && pc_tok && kpt && mask && loss, "cofi_desc_loss: null pointer");
    COFI_REQUIRE(n > 0 && n <= 128 && C > 0 && C <= 256 && frames > 0, "cofi_desc_loss: bad shape");
    COFI_REQUIRE((d_img == nullptr) == (d_pc == nullptr), "cofi_desc_loss: pass both gradient buffers or neither");
    const int smem = (2 * n * (C + 1) + 2 * n * (n + 1) + 5 * n) * (int)sizeof(float);
    COFI_REQUIRE(smem <= 200 * 1024, "cofi_desc_loss: n=%d, C=%d needs %d bytes of shared memory", n, C, smem);
    static int attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(desc_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(desc_loss smem=%d): %s", smem, cudaGetErrorString(e));
            return COFI_ECUDA;
        }
        attr = smem;
    }
    DescParams p{img_tok, pix, img_rows, pc_tok, kpt, pc_rows, mask, n, C, pos_margin, neg_margin, log_scale, loss, dists,
                 d_img, d_pc};
    desc_loss_kernel<<<frames, 256, smem, (cudaStream_t)stream>>>(p);
    return check_launch("cofi_desc_loss");
}

extern "C" int cofi_overlap_loss(const float* score, const int64_t* idx, int64_t rows, int n_in, int n_out, int frames,
                                 float* loss, float* d_score, void* stream) {
    COFI_REQUIRE(score && idx && loss && rows > 0 && n_in >= 0 && n_out >= 0 && n_in + n_out > 0 && frames > 0,
                 "cofi_overlap_loss: bad argument");
    overlap_loss_kernel<<<frames, 256, 0, (cudaStream_t)stream>>>(score, idx, rows, n_in, n_out, loss, d_score);
    return check_launch("cofi_overlap_loss");
}

extern "C" int cofi_fine_circle_loss(const float* patch, const float* fpc, const int64_t* rel, int n, int C, int frames,
                                     float m, float gamma, float* loss, float* d_patch, float* d_fpc, int32_t* bad_flag,
                                     void* stream) {
    COFI_REQUIRE(patch && fpc && rel && loss && n > 0 && frames > 0, "cofi_fine_circle_loss: bad argument");
    COFI_REQUIRE(C > 0 && C <= FC_MAXC && C % 2 == 0, "cofi_fine_circle_loss: C=%d must be even and <= %d", C, FC_MAXC);
    COFI_REQUIRE((d_patch == nullptr) == (d_fpc == nullptr), "cofi_fine_circle_loss: pass both gradient buffers or neither");
    fine_circle_loss_kernel<<<frames, 256, 0, (cudaStream_t)stream>>>(patch, fpc, rel, n, C, m, gamma, loss, d_patch, d_fpc,
                                                                      bad_flag);
    return check_launch("cofi_fine_circle_loss");
}

extern "C" int cofi_scatter_scaled_rows(const float* src, const int64_t* idx, int64_t R, int64_t rows_dst, int frames, int C,
                                        const float* scale_dev, float scale_host, float* dst, void* stream) {
    COFI_REQUIRE(src && dst && R > 0 && frames > 0 && C > 0 && (idx == nullptr || rows_dst > 0),
                 "cofi_scatter_scaled_rows: bad argument");
    const int64_t total = (int64_t)frames * R * C;
    int64_t blocks = ceil_div(total, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    scatter_scaled_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, idx, R, rows_dst, frames, C, scale_dev,
                                                                                 scale_host, dst);
    return check_launch("cofi_scatter_scaled_rows");
}
