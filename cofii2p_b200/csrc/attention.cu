// attention.cu -- multi-head softmax attention of the LoFTR encoder layer
// (reference model/transformer/linear_attention.py:69-77).  fp32 SIMT engine (COFI_GEMM_FP32): flash-style,
// one query row per thread, K/V tiles staged in shared memory and broadcast to the warp, online softmax in
// chunks of 16 keys; the [L,S,heads] score tensor the reference materialises never exists.
#include "common.cuh"

namespace cofi {

constexpr int AT_D = 32;       // head dimension
constexpr int AT_TQ = 128;     // queries per block (one per thread)
constexpr int AT_TK = 64;      // keys per shared-memory tile
constexpr int AT_CH = 16;      // keys per online-softmax chunk

__global__ void __launch_bounds__(AT_TQ)
attention_simt_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      int64_t L, int64_t S, int heads, float scale, float* __restrict__ out) {
    __shared__ __align__(16) float Ks[AT_TK][AT_D];
    __shared__ __align__(16) float Vs[AT_TK][AT_D];
    const int head = blockIdx.y, frame = blockIdx.z;
    const int64_t ld = (int64_t)heads * AT_D;
    const int64_t row = (int64_t)blockIdx.x * AT_TQ + threadIdx.x;
    const bool active = row < L;
    const float* qp = q + ((int64_t)frame * L + (active ? row : 0)) * ld + head * AT_D;
    const float* kb = k + ((int64_t)frame * S) * ld + head * AT_D;
    const float* vb = v + ((int64_t)frame * S) * ld + head * AT_D;

    float qr[AT_D], o[AT_D];
#pragma unroll
    for (int d = 0; d < AT_D; d += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(qp + d));
        qr[d] = t.x * scale;
        qr[d + 1] = t.y * scale;
        qr[d + 2] = t.z * scale;
        qr[d + 3] = t.w * scale;
        o[d] = o[d + 1] = o[d + 2] = o[d + 3] = 0.f;
    }
    float mrun = -INFINITY, lrun = 0.f;

    for (int64_t s0 = 0; s0 < S; s0 += AT_TK) {
        __syncthreads();
        // cooperative tile load: 64 keys x 32 floats = 512 float4 per matrix, 4 per thread
        for (int t = threadIdx.x; t < AT_TK * AT_D / 4; t += AT_TQ) {
            const int j = t / (AT_D / 4), d4 = t % (AT_D / 4);
            float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
            if (s0 + j < S) {
                kk = __ldg(reinterpret_cast<const float4*>(kb + (s0 + j) * ld) + d4);
                vv = __ldg(reinterpret_cast<const float4*>(vb + (s0 + j) * ld) + d4);
            }
            reinterpret_cast<float4*>(&Ks[j][0])[d4] = kk;
            reinterpret_cast<float4*>(&Vs[j][0])[d4] = vv;
        }
        __syncthreads();
        const int nk = (int)((S - s0) < AT_TK ? (S - s0) : AT_TK);
        for (int c0 = 0; c0 < nk; c0 += AT_CH) {
            float sc[AT_CH];
            float cmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < AT_CH; ++j) {
                float a = 0.f;
#pragma unroll
                for (int d = 0; d < AT_D; d += 4) {
                    const float4 kk = *reinterpret_cast<const float4*>(&Ks[c0 + j][d]);
                    a = fmaf(qr[d], kk.x, a);
                    a = fmaf(qr[d + 1], kk.y, a);
                    a = fmaf(qr[d + 2], kk.z, a);
                    a = fmaf(qr[d + 3], kk.w, a);
                }
                sc[j] = (c0 + j < nk) ? a : -INFINITY;
                cmax = fmaxf(cmax, sc[j]);
            }
            const float mnew = fmaxf(mrun, cmax);
            const float corr = expf(mrun - mnew);  // exp(-inf) = 0 on the first chunk
            lrun *= corr;
#pragma unroll
            for (int d = 0; d < AT_D; ++d) o[d] *= corr;
#pragma unroll
            for (int j = 0; j < AT_CH; ++j) {
                const float p = expf(sc[j] - mnew);
                lrun += p;
#pragma unroll
                for (int d = 0; d < AT_D; d += 4) {
                    const float4 vv = *reinterpret_cast<const float4*>(&Vs[c0 + j][d]);
                    o[d] = fmaf(p, vv.x, o[d]);
                    o[d + 1] = fmaf(p, vv.y, o[d + 1]);
                    o[d + 2] = fmaf(p, vv.z, o[d + 2]);
                    o[d + 3] = fmaf(p, vv.w, o[d + 3]);
                }
            }
            mrun = mnew;
        }
    }
    if (active) {
        float* op = out + ((int64_t)frame * L + row) * ld + head * AT_D;
        const float inv = 1.0f / lrun;
#pragma unroll
        for (int d = 0; d < AT_D; d += 4)
            *reinterpret_cast<float4*>(op + d) = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
    }
}

int attention_tc_launch(const float* q, const float* k, const float* vt, int64_t L, int64_t S, int frames, int heads,
                        int D, float scale, float* out, float* lse, cudaStream_t st);
bool attention_tc_supported(int64_t L, int64_t S, int heads, int D);

}  // namespace cofi

using namespace cofi;

extern "C" int cofi_attention(const float* q, const float* k, const float* v, int64_t L, int64_t S, int frames,
                              int heads, int D, float scale, float* out, int engine, void* stream) {
    COFI_REQUIRE(q && k && v && out, "cofi_attention: null pointer");
    COFI_REQUIRE(D == AT_D, "cofi_attention: head dimension %d unsupported (must be %d)", D, AT_D);
    COFI_REQUIRE(L > 0 && S > 0 && frames > 0 && heads > 0, "cofi_attention: bad shape");
    COFI_REQUIRE(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 &&
                     ((uintptr_t)out % 16) == 0,
                 "cofi_attention: 16-byte alignment required");
    (void)engine;  // row-major V: fp32 SIMT engine; the tcgen05 engine is cofi_attention_vt (K-major V^T operand)
    dim3 grid((unsigned)ceil_div(L, AT_TQ), heads, frames);
    attention_simt_kernel<<<grid, AT_TQ, 0, (cudaStream_t)stream>>>(q, k, v, L, S, heads, scale, out);
    return check_launch("cofi_attention(fp32)");
}

extern "C" int cofi_attention_vt(const float* q, const float* k, const float* vt, int64_t L, int64_t S, int frames,
                                 int heads, int D, float scale, float* out, void* stream) {
    COFI_REQUIRE(q && k && vt && out, "cofi_attention_vt: null pointer");
    COFI_REQUIRE(L > 0 && S > 0 && frames > 0 && heads > 0, "cofi_attention_vt: bad shape");
    COFI_REQUIRE(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)vt % 16) == 0 &&
                     ((uintptr_t)out % 16) == 0,
                 "cofi_attention_vt: 16-byte alignment required");
    if (!attention_tc_supported(L, S, heads, D)) {
        set_error("cofi_attention_vt: unsupported shape (D=%d must be 32 or 64, frames*S*4 bytes must be 16-byte aligned)", D);
        return COFI_EUNSUPPORTED;
    }
    return attention_tc_launch(q, k, vt, L, S, frames, heads, D, scale, out, nullptr, (cudaStream_t)stream);
}

extern "C" int cofi_attention_vt_lse(const float* q, const float* k, const float* vt, int64_t L, int64_t S, int frames,
                                     int heads, int D, float scale, float* out, float* lse, void* stream) {
    COFI_REQUIRE(q && k && vt && out && lse, "cofi_attention_vt_lse: null pointer");
    COFI_REQUIRE(L > 0 && S > 0 && frames > 0 && heads > 0, "cofi_attention_vt_lse: bad shape");
    COFI_REQUIRE(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)vt % 16) == 0 &&
                     ((uintptr_t)out % 16) == 0,
                 "cofi_attention_vt_lse: 16-byte alignment required");
    if (!attention_tc_supported(L, S, heads, D)) {
        set_error("cofi_attention_vt_lse: unsupported shape (D=%d must be 32 or 64, frames*S*4 bytes must be 16-byte aligned)", D);
        return COFI_EUNSUPPORTED;
    }
    return attention_tc_launch(q, k, vt, L, S, frames, heads, D, scale, out, lse, (cudaStream_t)stream);
}
