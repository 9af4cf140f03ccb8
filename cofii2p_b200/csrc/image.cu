// image.cu -- NHWC helpers of the image stream (reference model/imagenet.py:145,204,433,441-443).
// All HBM-bound elementwise/gather work: channel index fastest so warps read/write contiguous 128-byte lines.
#include "common.cuh"

namespace cofi {

__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ x, int B, int C, int H, int W, int Cpad, float* __restrict__ y) {
    const int64_t total = (int64_t)B * H * W * Cpad;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % Cpad);
        const int64_t p = t / Cpad;  // b*H*W + h*W + w
        const int64_t hw = p % ((int64_t)H * W);
        const int64_t b = p / ((int64_t)H * W);
        y[t] = c < C ? __ldg(x + (b * C + c) * (int64_t)H * W + hw) : 0.0f;
    }
}

// tiled transpose per image: [HW, C] -> [C, HW]
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const float* __restrict__ x, int64_t HW, int C, float* __restrict__ y) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const float* xb = x + (int64_t)b * HW * C;
    float* yb = y + (int64_t)b * HW * C;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int64_t p = p0 + i;
        const int c = c0 + tx;
        tile[i][tx] = (p < HW && c < C) ? __ldg(xb + p * C + c) : 0.0f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i;
        const int64_t p = p0 + tx;
        if (p < HW && c < C) yb[(int64_t)c * HW + p] = tile[tx][i];
    }
}

__global__ void __launch_bounds__(256)
maxpool2d_3x3s2_kernel(const float* __restrict__ x, int B, int H, int W, int C, int Ho, int Wo,
                       float* __restrict__ y) {
    const int64_t total = (int64_t)B * Ho * Wo * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        int64_t p = t / C;
        const int wo = (int)(p % Wo);
        p /= Wo;
        const int ho = (int)(p % Ho);
        const int b = (int)(p / Ho);
        float m = -INFINITY;
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
            const int hi = ho * 2 + dh - 1;
            if (hi < 0 || hi >= H) continue;
#pragma unroll
            for (int dw = 0; dw < 3; ++dw) {
                const int wi = wo * 2 + dw - 1;
                if (wi < 0 || wi >= W) continue;
                m = fmaxf(m, __ldg(x + (((int64_t)b * H + hi) * W + wi) * C + c));
            }
        }
        y[t] = m;
    }
}

// y[b,ho,wo, 0:C1] = bilinear x2 of x1 (align_corners=False, PyTorch's source index = (dst+0.5)/2-0.5 clamped
// at 0), y[b,ho,wo, C1:C1+C2] = x2[b,ho,wo,:]
__global__ void __launch_bounds__(256)
upsample2x_cat_kernel(const float* __restrict__ x1, int B, int H, int W, int C1, const float* __restrict__ x2,
                      int C2, float* __restrict__ y) {
    const int Ho = 2 * H, Wo = 2 * W, Ct = C1 + C2;
    const int64_t total = (int64_t)B * Ho * Wo * Ct;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % Ct);
        int64_t p = t / Ct;
        const int wo = (int)(p % Wo);
        p /= Wo;
        const int ho = (int)(p % Ho);
        const int b = (int)(p / Ho);
        float v;
        if (c >= C1) {
            v = __ldg(x2 + (((int64_t)b * Ho + ho) * Wo + wo) * C2 + (c - C1));
        } else {
            // ATen upsample_bilinear2d (area_pixel_compute_source_index, align_corners=False)
            float sh = ((float)ho + 0.5f) * 0.5f - 0.5f;
            float sw = ((float)wo + 0.5f) * 0.5f - 0.5f;
            if (sh < 0.f) sh = 0.f;
            if (sw < 0.f) sw = 0.f;
            const int h0 = (int)sh, w0 = (int)sw;
            const int h1 = h0 + (h0 < H - 1 ? 1 : 0), w1 = w0 + (w0 < W - 1 ? 1 : 0);
            const float lh1 = sh - (float)h0, lw1 = sw - (float)w0;
            const float lh0 = 1.0f - lh1, lw0 = 1.0f - lw1;
            const float* xb = x1 + (int64_t)b * H * W * C1;
            const float v00 = __ldg(xb + ((int64_t)h0 * W + w0) * C1 + c);
            const float v01 = __ldg(xb + ((int64_t)h0 * W + w1) * C1 + c);
            const float v10 = __ldg(xb + ((int64_t)h1 * W + w0) * C1 + c);
            const float v11 = __ldg(xb + ((int64_t)h1 * W + w1) * C1 + c);
            v = lh0 * (lw0 * v00 + lw1 * v01) + lh1 * (lw0 * v10 + lw1 * v11);
        }
        y[t] = v;
    }
}

}  // namespace cofi

using namespace cofi;

static unsigned ew_blocks(int64_t total, int threads) {
    int64_t b = ceil_div(total, threads);
    const int64_t cap = 148 * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

extern "C" int cofi_nchw_to_nhwc(const float* x, int B, int C, int H, int W, int Cpad, float* y, void* stream) {
    COFI_REQUIRE(x && y && B > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C, "cofi_nchw_to_nhwc: bad argument");
    const int64_t total = (int64_t)B * H * W * Cpad;
    nchw_to_nhwc_kernel<<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(x, B, C, H, W, Cpad, y);
    return check_launch("cofi_nchw_to_nhwc");
}

extern "C" int cofi_nhwc_to_nchw(const float* x, int B, int H, int W, int C, float* y, void* stream) {
    COFI_REQUIRE(x && y && B > 0 && C > 0 && H > 0 && W > 0, "cofi_nhwc_to_nchw: bad argument");
    const int64_t HW = (int64_t)H * W;
    dim3 grid((unsigned)ceil_div(HW, 32), (unsigned)ceil_div(C, 32), B);
    nhwc_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, HW, C, y);
    return check_launch("cofi_nhwc_to_nchw");
}

extern "C" int cofi_maxpool2d_3x3s2_nhwc(const float* x, int B, int H, int W, int C, float* y, void* stream) {
    COFI_REQUIRE(x && y && B > 0 && C > 0 && H > 0 && W > 0, "cofi_maxpool2d_3x3s2_nhwc: bad argument");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const int64_t total = (int64_t)B * Ho * Wo * C;
    maxpool2d_3x3s2_kernel<<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, C, Ho, Wo, y);
    return check_launch("cofi_maxpool2d_3x3s2_nhwc");
}

extern "C" int cofi_upsample2x_cat_nhwc(const float* x1, int B, int H, int W, int C1, const float* x2, int C2,
                                        float* y, void* stream) {
    COFI_REQUIRE(x1 && x2 && y && B > 0 && C1 > 0 && C2 > 0 && H > 0 && W > 0, "cofi_upsample2x_cat_nhwc: bad argument");
    const int64_t total = (int64_t)B * 4 * H * W * (C1 + C2);
    upsample2x_cat_kernel<<<ew_blocks(total, 256), 256, 0, (cudaStream_t)stream>>>(x1, B, H, W, C1, x2, C2, y);
    return check_launch("cofi_upsample2x_cat_nhwc");
}
