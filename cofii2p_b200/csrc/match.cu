// match.cu -- positional encoding and the coarse/fine matching kernels
// (reference model/network.py:167-264, model/transformer/position_encoding.py:29-50,
//  evaluation/eval_all.py:99-105).  Integer results (arg-min / arg-max / selections) are produced in fp32
// with the reference's accumulation order and lowest-index tie-breaking so that they are bit-exact.
#include "common.cuh"

namespace cofi {

// ---------------------------------------------------------------------------------------------- posenc
__global__ void __launch_bounds__(256)
posenc_sine_kernel(const float* __restrict__ coords, int64_t rows, int n_dim, int npf, int d_model,
                   const float* __restrict__ dim_t, float scale, float* __restrict__ out) {
    const int64_t total = rows * d_model;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / d_model;
        const int c = (int)(t - row * d_model);
        float v = 0.0f;  // zero padding beyond n_dim*npf
        if (c < n_dim * npf) {
            const int d = c / npf, i = c - d * npf;
            const float a = __fmul_rn(__ldg(coords + row * n_dim + d), scale);
            const float pd = __fdiv_rn(a, __ldg(dim_t + i));
            v = (i & 1) ? cosf(pd) : sinf(pd);
        }
        out[t] = v;
    }
}

// ----------------------------------------------------------------------------------------- sim_argmin
// fp32 engine. Block = 256 threads, SP points per block staged in shared memory; each thread walks pixels
// tid, tid+256, ... and keeps (min distance, index) per point; block reduction with lowest-index ties.
// Accumulation order mirrors ATen's cascade sum over the channel axis (groups of 16, sequential inside and
// across groups; aten/src/ATen/native/cpu/SumKernel.cpp multi_row_sum) applied to rounded products, i.e. the
// arithmetic of `1 - torch.sum(img.unsqueeze(-1) * pc.unsqueeze(-2), dim=0)` (reference network.py:174).
constexpr int SP = 8;
constexpr int SIM_MAXC = 256;

__global__ void __launch_bounds__(256)
sim_argmin_simt_kernel(const float* __restrict__ pt, int64_t ldpt, const float* __restrict__ px, int64_t ldpx,
                       int64_t Npt, int64_t Npx, int C, int64_t* __restrict__ best_idx,
                       float* __restrict__ best_val) {
    __shared__ float spt[SP][SIM_MAXC];
    __shared__ float rv[SP][8];
    __shared__ int ri[SP][8];
    const int frame = blockIdx.y;
    const int64_t p0 = (int64_t)blockIdx.x * SP;
    const float* ptb = pt + (int64_t)frame * Npt * ldpt;
    const float* pxb = px + (int64_t)frame * Npx * ldpx;
    for (int t = threadIdx.x; t < SP * C; t += blockDim.x) {
        const int p = t / C, c = t - p * C;
        spt[p][c] = (p0 + p < Npt) ? __ldg(ptb + (p0 + p) * ldpt + c) : 0.0f;
    }
    __syncthreads();
    float bv[SP];
    int bi[SP];
#pragma unroll
    for (int p = 0; p < SP; ++p) {
        bv[p] = INFINITY;
        bi[p] = 0x7fffffff;
    }
    for (int64_t x = threadIdx.x; x < Npx; x += blockDim.x) {
        const float* r = pxb + x * ldpx;
        float tot[SP];
#pragma unroll
        for (int p = 0; p < SP; ++p) tot[p] = 0.0f;
        for (int c0 = 0; c0 < C; c0 += 16) {
            float g[SP];
#pragma unroll
            for (int p = 0; p < SP; ++p) g[p] = 0.0f;
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(r + c0 + c));
#pragma unroll
                for (int p = 0; p < SP; ++p) {
                    g[p] = __fadd_rn(g[p], __fmul_rn(a.x, spt[p][c0 + c + 0]));
                    g[p] = __fadd_rn(g[p], __fmul_rn(a.y, spt[p][c0 + c + 1]));
                    g[p] = __fadd_rn(g[p], __fmul_rn(a.z, spt[p][c0 + c + 2]));
                    g[p] = __fadd_rn(g[p], __fmul_rn(a.w, spt[p][c0 + c + 3]));
                }
            }
#pragma unroll
            for (int p = 0; p < SP; ++p) tot[p] = __fadd_rn(tot[p], g[p]);
        }
#pragma unroll
        for (int p = 0; p < SP; ++p) {
            const float d = __fsub_rn(1.0f, tot[p]);
            if (d < bv[p]) {  // strict: keeps the lowest index seen by this thread (x ascending)
                bv[p] = d;
                bi[p] = (int)x;
            }
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int p = 0; p < SP; ++p) {
        float v = bv[p];
        int i = bi[p];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (ov < v || (ov == v && oi < i)) {
                v = ov;
                i = oi;
            }
        }
        if (lane == 0) {
            rv[p][w] = v;
            ri[p][w] = i;
        }
    }
    __syncthreads();
    if (threadIdx.x < SP) {
        const int p = threadIdx.x;
        float v = rv[p][0];
        int i = ri[p][0];
        for (int k = 1; k < 8; ++k)
            if (rv[p][k] < v || (rv[p][k] == v && ri[p][k] < i)) {
                v = rv[p][k];
                i = ri[p][k];
            }
        if (p0 + p < Npt) {
            best_idx[(int64_t)frame * Npt + p0 + p] = i;
            best_val[(int64_t)frame * Npt + p0 + p] = v;
        }
    }
}

// --------------------------------------------------------------------------------------- select_matches
// reference model/network.py:145-151 (threshold loop 0.9, 0.88, ... until >= min_count matches survive) + :181-187
// (score >= thr, border mask on the arg-min pixel, compaction in point order).  One CTA per frame, no host round trip:
//   1. level(p) = first threshold index t with score[p] >= thr[t] (thresholds descend, so p passes every later one);
//      histogram of the levels of the border-masked points, prefix sum -> first t whose count reaches min_count;
//   2. ordered compaction (ballot + warp/block prefix) of the points passing thr[t];
//   3. rows n..Npt-1 of the fixed-size outputs are padded with a valid dummy (index 0, centre (2,2)*xy_scale) so that the
//      downstream fixed-shape kernels of a captured graph touch only valid memory; the count says how many rows are real.
constexpr int SEL_MAXTHR = 128;

__global__ void __launch_bounds__(1024)
select_matches_kernel(const float* __restrict__ score, const int64_t* __restrict__ best_idx, int64_t Npt, int gridH,
                      int gridW, int xmax, int ymax, const float* __restrict__ thresholds, int nthr, int min_count,
                      float xy_scale,
                      int32_t* __restrict__ out_count, int64_t* __restrict__ out_index, float* __restrict__ out_xy) {
    const int frame = blockIdx.x;
    const float* sc = score + (int64_t)frame * Npt;
    const int64_t* bi = best_idx + (int64_t)frame * Npt;
    __shared__ float s_thr[SEL_MAXTHR];
    __shared__ int s_hist[SEL_MAXTHR];
    __shared__ int s_warp[32];
    __shared__ int s_chosen, s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < nthr; t += blockDim.x) {
        s_thr[t] = thresholds[t];
        s_hist[t] = 0;
    }
    __syncthreads();
    for (int64_t p = tid; p < Npt; p += blockDim.x) {
        const int x = (int)(bi[p] % gridW), y = (int)(bi[p] / gridW);
        const bool m = (x >= 2) && (x <= xmax) && (y <= ymax) && (y >= 2);
        if (m) {
            const float v = sc[p];
            int t = 0;
            while (t < nthr && !(v >= s_thr[t])) ++t;
            if (t < nthr) atomicAdd(&s_hist[t], 1);
        }
    }
    __syncthreads();
    if (tid == 0) {
        int chosen = nthr - 1, acc = 0;
        for (int t = 0; t < nthr; ++t) {
            acc += s_hist[t];
            if (acc >= min_count) {
                chosen = t;
                break;
            }
        }
        s_chosen = chosen;
        s_base = 0;
    }
    __syncthreads();
    const int chosen = s_chosen;
    const float thr = s_thr[chosen];
    int64_t* oi = out_index + (int64_t)frame * Npt;
    float* ox = out_xy + (int64_t)frame * 2 * Npt;
    for (int64_t p0 = 0; p0 < Npt; p0 += blockDim.x) {
        const int64_t p = p0 + tid;
        bool keep = false;
        int x = 0, y = 0;
        if (p < Npt) {
            x = (int)(bi[p] % gridW);
            y = (int)(bi[p] / gridW);
            const bool m = (x >= 2) && (x <= xmax) && (y <= ymax) && (y >= 2);
            keep = m && (sc[p] >= thr);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (keep) {
            const int n = before + __popc(bal & ((1u << lane) - 1u));
            oi[n] = p;
            ox[n] = (float)x * xy_scale;
            ox[Npt + n] = (float)y * xy_scale;
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    const int n = s_base;
    for (int64_t p = n + tid; p < Npt; p += blockDim.x) {
        oi[p] = 0;
        ox[p] = 2.0f * xy_scale;
        ox[Npt + p] = 2.0f * xy_scale;
    }
    if (tid == 0) {
        out_count[frame * 2 + 0] = n;
        out_count[frame * 2 + 1] = chosen;
    }
}

// -------------------------------------------------------------------------------------------- nn_argmin
__global__ void __launch_bounds__(256)
nn_argmin_kernel(const float* __restrict__ points, int64_t n, const float* __restrict__ nodes, int64_t M,
                 int64_t* __restrict__ idx) {
    const int64_t i = blockIdx.x;
    if (i >= n) return;
    points += (int64_t)blockIdx.y * n * 3;   // frame-local clouds and indices
    nodes += (int64_t)blockIdx.y * M * 3;
    idx += (int64_t)blockIdx.y * n;
    const float px = points[i * 3], py = points[i * 3 + 1], pz = points[i * 3 + 2];
    const float sp = __fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz));
    float bv = INFINITY;
    int bi = 0x7fffffff;
    for (int64_t j = threadIdx.x; j < M; j += blockDim.x) {
        const float nx = __ldg(nodes + j * 3), ny = __ldg(nodes + j * 3 + 1), nz = __ldg(nodes + j * 3 + 2);
        const float dot = __fadd_rn(__fadd_rn(__fmul_rn(px, nx), __fmul_rn(py, ny)), __fmul_rn(pz, nz));
        const float sn = __fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz));
        float d = __fmul_rn(-2.0f, dot);  // reference network.py:239-246: -2ab, += |a|^2, += |b|^2, clamp
        d = __fadd_rn(d, sp);
        d = __fadd_rn(d, sn);
        d = fmaxf(d, 1e-12f);
        if (d < bv) {
            bv = d;
            bi = (int)j;
        }
    }
    __shared__ float rv[8];
    __shared__ int ri[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
        }
    }
    if (lane == 0) {
        rv[w] = bv;
        ri[w] = bi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k)
            if (rv[k] < bv || (rv[k] == bv && ri[k] < bi)) {
                bv = rv[k];
                bi = ri[k];
            }
        idx[i] = bi;
    }
}

// ------------------------------------------------------------------------------------------ extract_patch
__global__ void __launch_bounds__(256)
extract_patch_kernel(const float* __restrict__ map, int H, int W, int C, int b, const float* __restrict__ centers,
                     int64_t n, float* __restrict__ out, int32_t* __restrict__ err_flag, int frames = 1) {
    const int64_t per = n * C * 16, total = per * frames;
    const float* centers0 = centers;
    float* out0 = out;
    const int b0 = b;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(t / per);          // batched form: frame f reads centres [f,2,n], map image b0+f
        const int64_t tf = t - (int64_t)f * per;
        centers = centers0 + (int64_t)f * 2 * n;
        out = out0 + (int64_t)f * per;
        b = b0 + f;
        const int c = (int)(tf % C);           // channel fastest for coalesced NHWC reads
        const int64_t r = tf / C;
        const int pix = (int)(r % 16);
        const int64_t i = r / 16;
        const int dy = pix >> 2, dx = pix & 3;
        const int left = (int)floorf(centers[i] - 2.0f);
        const int top = (int)floorf(centers[n + i] - 2.0f);
        const int right = (int)floorf(centers[i] + 2.0f);
        const int bottom = (int)floorf(centers[n + i] + 2.0f);
        float v = 0.0f;
        if (left < 0 || top < 0 || right > W || bottom > H || right - left != 4 || bottom - top != 4) {
            if (err_flag) *err_flag = 1;  // the reference asserts patch.shape == (B,C,4,4), network.py:222
        } else {
            v = __ldg(map + (((int64_t)b * H + top + dy) * W + left + dx) * C + c);
        }
        out[(i * C + c) * 16 + pix] = v;
    }
}

// --------------------------------------------------------------------------------------------- fine_match
__global__ void __launch_bounds__(128)
fine_match_kernel(const float* __restrict__ patch, const float* __restrict__ pc, int64_t n, int C,
                  int64_t* __restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    float dot = 0.f, na = 0.f, nb = 0.f;
    if (lane < 16) {
        for (int c = 0; c < C; ++c) {
            const float a = __ldg(patch + (i * C + c) * 16 + lane);
            const float b = __ldg(pc + i * C + c);
            dot = fmaf(a, b, dot);
            na = fmaf(a, a, na);
            nb = fmaf(b, b, nb);
        }
    }
    // torch.cosine_similarity: x.y / (max(|x|,eps) * max(|y|,eps)), eps = 1e-8
    float sim = (lane < 16) ? dot / (fmaxf(sqrtf(na), 1e-8f) * fmaxf(sqrtf(nb), 1e-8f)) : -INFINITY;
    int bi = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, sim, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > sim || (ov == sim && oi < bi)) {
            sim = ov;
            bi = oi;
        }
    }
    if (lane == 0) idx[i] = bi;
}

int sim_argmin_tc_launch(const float* pt, int64_t ldpt, const float* px, int64_t ldpx, int64_t Npt, int64_t Npx,
                         int C, int frames, int64_t* best_idx, float* best_val, int engine, cudaStream_t st);
bool sim_argmin_tc_supported(int64_t ldpt, int64_t ldpx, int64_t Npt, int64_t Npx, int C);

}  // namespace cofi

using namespace cofi;

static unsigned ew_blocks(int64_t total, int threads) {
    int64_t b = ceil_div(total, threads);
    const int64_t cap = 148 * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int cofi_posenc_sine(const float* coords, int64_t rows, int n_dim, int d_model, const float* dim_t,
                                float* out, void* stream) {
    COFI_REQUIRE(coords && out && dim_t && rows >= 0 && n_dim > 0 && d_model >= n_dim * 2,
                 "cofi_posenc_sine: bad argument");
    if (rows == 0) return COFI_OK;
    const int npf = d_model / n_dim / 2 * 2;
    const float scale = (float)(2.0 * 3.14159265358979323846);
    posenc_sine_kernel<<<ew_blocks(rows * d_model, 256), 256, 0, (cudaStream_t)stream>>>(coords, rows, n_dim, npf,
                                                                                        d_model, dim_t, scale, out);
    return check_launch("cofi_posenc_sine");
}

extern "C" int cofi_sim_argmin(const float* pt, int64_t ldpt, const float* px, int64_t ldpx, int64_t Npt,
                               int64_t Npx, int C, int frames, int64_t* best_idx, float* best_val, int engine,
                               void* stream) {
    COFI_REQUIRE(pt && px && best_idx && best_val, "cofi_sim_argmin: null pointer");
    COFI_REQUIRE(Npt > 0 && Npx > 0 && frames > 0, "cofi_sim_argmin: bad shape");
    COFI_REQUIRE(C > 0 && C % 16 == 0 && C <= SIM_MAXC, "cofi_sim_argmin: C=%d must be a multiple of 16, <= %d", C,
                 SIM_MAXC);
    COFI_REQUIRE(ldpx % 4 == 0 && ((uintptr_t)px % 16) == 0, "cofi_sim_argmin: px must be 16-byte aligned rows");
    if (engine != COFI_GEMM_FP32 && sim_argmin_tc_supported(ldpt, ldpx, Npt, Npx, C))
        return sim_argmin_tc_launch(pt, ldpt, px, ldpx, Npt, Npx, C, frames, best_idx, best_val, engine,
                                    (cudaStream_t)stream);
    dim3 grid((unsigned)ceil_div(Npt, SP), frames);
    sim_argmin_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pt, ldpt, px, ldpx, Npt, Npx, C, best_idx,
                                                                   best_val);
    return check_launch("cofi_sim_argmin(fp32)");
}

extern "C" int cofi_select_matches(const float* score, const int64_t* best_idx, int64_t Npt, int frames, int gridH,
                                   int gridW, int xmax, int ymax, const float* thresholds, int nthr, int min_count,
                                   float xy_scale, int32_t* out_count, int64_t* out_index, float* out_xy, void* stream) {
    COFI_REQUIRE(score && best_idx && thresholds && out_count && out_index && out_xy, "cofi_select_matches: null pointer");
    COFI_REQUIRE(Npt > 0 && frames > 0 && nthr > 0 && nthr <= SEL_MAXTHR && gridH > 4 && gridW > 4,
                 "cofi_select_matches: bad shape (at most %d thresholds)", SEL_MAXTHR);
    select_matches_kernel<<<frames, 1024, 0, (cudaStream_t)stream>>>(score, best_idx, Npt, gridH, gridW, xmax, ymax,
                                                                    thresholds, nthr, min_count, xy_scale, out_count, out_index,
                                                                    out_xy);
    return check_launch("cofi_select_matches");
}

extern "C" int cofi_nn_argmin(const float* points, int64_t n, const float* nodes, int64_t M, int64_t* idx,
                              void* stream) {
    COFI_REQUIRE(points && nodes && idx && n >= 0 && M > 0, "cofi_nn_argmin: bad argument");
    if (n == 0) return COFI_OK;
    nn_argmin_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(points, n, nodes, M, idx);
    return check_launch("cofi_nn_argmin");
}

extern "C" int cofi_nn_argmin_batched(const float* points, int64_t n, const float* nodes, int64_t M, int frames,
                                      int64_t* idx, void* stream) {
    COFI_REQUIRE(points && nodes && idx && n >= 0 && M > 0 && frames > 0 && frames < 65536, "cofi_nn_argmin_batched: bad argument");
    if (n == 0) return COFI_OK;
    nn_argmin_kernel<<<dim3((unsigned)n, (unsigned)frames), 256, 0, (cudaStream_t)stream>>>(points, n, nodes, M, idx);
    return check_launch("cofi_nn_argmin_batched");
}

extern "C" int cofi_extract_patch_batched(const float* map, int H, int W, int C, int frames, const float* centers,
                                          int64_t n, float* out, int32_t* err_flag, void* stream) {
    COFI_REQUIRE(map && centers && out && H >= 4 && W >= 4 && C > 0 && frames > 0 && n >= 0,
                 "cofi_extract_patch_batched: bad argument");
    if (n == 0) return COFI_OK;
    extract_patch_kernel<<<ew_blocks(n * C * 16 * frames, 256), 256, 0, (cudaStream_t)stream>>>(map, H, W, C, 0, centers, n,
                                                                                               out, err_flag, frames);
    return check_launch("cofi_extract_patch_batched");
}

extern "C" int cofi_extract_patch(const float* map, int H, int W, int C, int b, const float* centers, int64_t n,
                                  float* out, int32_t* err_flag, void* stream) {
    COFI_REQUIRE(map && centers && out && H >= 4 && W >= 4 && C > 0 && b >= 0 && n >= 0,
                 "cofi_extract_patch: bad argument");
    if (n == 0) return COFI_OK;
    extract_patch_kernel<<<ew_blocks(n * C * 16, 256), 256, 0, (cudaStream_t)stream>>>(map, H, W, C, b, centers, n,
                                                                                      out, err_flag);
    return check_launch("cofi_extract_patch");
}

extern "C" int cofi_fine_match(const float* patch, const float* pc, int64_t n, int C, int64_t* idx, void* stream) {
    COFI_REQUIRE(patch && pc && idx && n >= 0 && C > 0, "cofi_fine_match: bad argument");
    if (n == 0) return COFI_OK;
    const int wpb = 4;
    fine_match_kernel<<<(unsigned)ceil_div(n, wpb), wpb * 32, 0, (cudaStream_t)stream>>>(patch, pc, n, C, idx);
    return check_launch("cofi_fine_match");
}
