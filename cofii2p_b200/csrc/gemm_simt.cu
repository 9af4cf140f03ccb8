// gemm_simt.cu -- fp32 SIMT contraction engine (COFI_GEMM_FP32): the exact-arithmetic parity path for every
// dense contraction of the hot path (nn.Linear, KPConv weight-apply, 3x3/7x7/1x1 convolutions as implicit GEMM).
// The tensor-core engines (gemm_tc.cu) are checked against this one and against the CPU oracle.
//
// C[M,N] = A[M,K] * W[N,K]^T, both operands K-major.  128x64 tile, BK=16, 256 threads, 8x4 micro-tile,
// float4 global loads, double-buffered shared memory.  A is produced by a loader functor so the same
// kernel serves dense matrices and the im2col view of an NHWC image.
#include "common.cuh"

namespace cofi {

struct DenseA {
    const float* A;
    int64_t lda;
    int64_t M;
    int K;
    __device__ __forceinline__ float4 load4(int64_t m, int k) const {
        if (m < M && k < K) return __ldg(reinterpret_cast<const float4*>(A + m * lda + k));
        return make_float4(0.f, 0.f, 0.f, 0.f);
    }
};

// im2col view: row m = (b, ho, wo); column k = (kh, kw, ci) with ci fastest; Cin % 4 == 0
struct ConvA {
    const float* x;
    int B, H, W, Cin, KH, KW, stride, pad, Ho, Wo;
    int64_t M;
    int K;
    __device__ __forceinline__ float4 load4(int64_t m, int k) const {
        if (m >= M || k >= K) return make_float4(0.f, 0.f, 0.f, 0.f);
        const int wo = (int)(m % Wo);
        const int64_t t = m / Wo;
        const int ho = (int)(t % Ho);
        const int b = (int)(t / Ho);
        const int ci = k % Cin;
        const int tap = k / Cin;
        const int kw = tap % KW, kh = tap / KW;
        const int hi = ho * stride + kh - pad, wi = wo * stride + kw - pad;
        if (hi < 0 || hi >= H || wi < 0 || wi >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
        return __ldg(reinterpret_cast<const float4*>(x + (((int64_t)b * H + hi) * W + wi) * Cin + ci));
    }
};

struct Epilogue {
    const float* bias;      // [N] or null (added)
    const float* rowdiv;    // [M] or null (acc / rowdiv[m])
    const float* colscale;  // [N] or null (acc * colscale[n] + colshift[n])
    const float* colshift;
    const float* residual;  // [M, ldres] or null
    int64_t ldres;
    int accumulate;
    int act;
    const float* ln_gamma;  // fused row LayerNorm over the N columns (tcgen05 engine, N <= 128): y = act(LN(x))+residual
    const float* ln_beta;
    float ln_eps;
};

constexpr int BM = 128, BN = 64, BK = 16, TM = 8, TN = 4;

template <class ALoader>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(ALoader al, const float* __restrict__ Wt, int64_t ldw, float* __restrict__ C, int64_t ldc,
                 int64_t M, int N, int K, Epilogue ep) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Ws[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    // loaders: A tile 128 rows x 16 k = 512 float4 -> 2 per thread; W tile 64 x 16 = 256 float4 -> 1 per thread
    const int lrow = tid >> 2;        // 0..63
    const int lk = (tid & 3) * 4;     // 0,4,8,12
    // compute mapping: 16 x 16 threads; thread (ty, tx) owns rows ty*8..+7, cols tx*4..+3
    const int ty = tid >> 4, tx = tid & 15;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra0, ra1, rw;
    auto gload = [&](int k0) {
        ra0 = al.load4(m0 + lrow, k0 + lk);
        ra1 = al.load4(m0 + lrow + 64, k0 + lk);
        const int n = n0 + lrow;
        if (n < N && k0 + lk < K)
            rw = __ldg(reinterpret_cast<const float4*>(Wt + (int64_t)n * ldw + k0 + lk));
        else
            rw = make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto sstore = [&](int buf) {
        As[buf][lk + 0][lrow] = ra0.x;
        As[buf][lk + 1][lrow] = ra0.y;
        As[buf][lk + 2][lrow] = ra0.z;
        As[buf][lk + 3][lrow] = ra0.w;
        As[buf][lk + 0][lrow + 64] = ra1.x;
        As[buf][lk + 1][lrow + 64] = ra1.y;
        As[buf][lk + 2][lrow + 64] = ra1.z;
        As[buf][lk + 3][lrow + 64] = ra1.w;
        Ws[buf][lk + 0][lrow] = rw.x;
        Ws[buf][lk + 1][lrow] = rw.y;
        Ws[buf][lk + 2][lrow] = rw.z;
        Ws[buf][lk + 3][lrow] = rw.w;
    };

    const int nk = (K + BK - 1) / BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            const float4 w = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + ty * 8 + i;
        if (m >= M) continue;
        const float rd = ep.rowdiv ? __ldg(ep.rowdiv + m) : 1.0f;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (ep.rowdiv) v = v / rd;
            if (ep.colscale) v = v * __ldg(ep.colscale + n) + __ldg(ep.colshift + n);
            if (ep.bias) v += __ldg(ep.bias + n);
            if (ep.residual) v += __ldg(ep.residual + m * ep.ldres + n);
            if (ep.accumulate) v += C[m * ldc + n];
            C[m * ldc + n] = apply_act(v, ep.act);
        }
    }
}

int gemm_simt_launch(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t M,
                     int N, int K, const Epilogue& ep, cudaStream_t st) {
    DenseA al{A, lda, M, K};
    dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, BN));
    gemm_simt_kernel<DenseA><<<grid, 256, 0, st>>>(al, W, ldw, C, ldc, M, N, K, ep);
    return check_launch("cofi_gemm(fp32)");
}

int conv_simt_launch(const float* x, int B, int H, int W, int Cin, const float* w, int Cout, int KH, int KW,
                     int stride, int pad, float* y, const Epilogue& ep, cudaStream_t st) {
    ConvA al;
    al.x = x;
    al.B = B;
    al.H = H;
    al.W = W;
    al.Cin = Cin;
    al.KH = KH;
    al.KW = KW;
    al.stride = stride;
    al.pad = pad;
    al.Ho = (H + 2 * pad - KH) / stride + 1;
    al.Wo = (W + 2 * pad - KW) / stride + 1;
    al.M = (int64_t)B * al.Ho * al.Wo;
    al.K = KH * KW * Cin;
    dim3 grid((unsigned)ceil_div(al.M, BM), (unsigned)ceil_div(Cout, BN));
    gemm_simt_kernel<ConvA><<<grid, 256, 0, st>>>(al, w, al.K, y, Cout, al.M, Cout, al.K, ep);
    return check_launch("cofi_conv2d_nhwc(fp32)");
}

// tensor-core engines (gemm_tc.cu)
int gemm_tc_launch(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t M, int N,
                   int K, const Epilogue& ep, int engine, cudaStream_t st, float* stat_out = nullptr);
bool gemm_tc_supported(int64_t lda, int64_t ldw, int64_t ldc, int64_t M, int N, int K, const void* A, const void* W,
                       const void* C);
int gemm_tc_f16_launch(const void* A, int64_t lda, const void* W, int64_t ldw, float* C, int64_t ldc, int64_t M, int N,
                       int K, const Epilogue& ep, cudaStream_t st, float* stat_out = nullptr);
bool conv_tc_supported(int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
int conv_tc_launch(const float* x, int B, int H, int W, int Cin, const float* w, int Cout, int KH, int KW, int stride, int pad,
                   float* y, const Epilogue& ep, int engine, cudaStream_t st);

// persistent 3xTF32 engine with pre-split weights (gemm_x3.cu)
int gemm_x3_launch(const float* A, int64_t lda, const float* W2, float* C, int64_t ldc, int64_t M, int N, int K,
                   const Epilogue& ep, cudaStream_t st, float* stat_out);
int conv_x3_launch(const float* x, int B, int H, int W, int Cin, const float* w2, int Cout, int KH, int KW, int stride, int pad,
                   float* y, const Epilogue& ep, cudaStream_t st);

}  // namespace cofi

using namespace cofi;

extern "C" int cofi_gemm(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t M,
                         int N, int K, const float* bias, const float* rowdiv, int accumulate, int act, int engine,
                         void* stream) {
    COFI_REQUIRE(A && W && C, "cofi_gemm: null pointer");
    COFI_REQUIRE(M >= 0 && N > 0 && K > 0, "cofi_gemm: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
    COFI_REQUIRE(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0, "cofi_gemm: K, lda, ldw must be multiples of 4");
    COFI_REQUIRE(lda >= K && ldw >= K && ldc >= N, "cofi_gemm: leading dimension too small");
    COFI_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0, "cofi_gemm: A and W must be 16-byte aligned");
    if (M == 0) return COFI_OK;
    Epilogue ep{bias, rowdiv, nullptr, nullptr, nullptr, 0, accumulate, act, nullptr, nullptr, 0.0f};
    if (engine == COFI_GEMM_FP32) return gemm_simt_launch(A, lda, W, ldw, C, ldc, M, N, K, ep, (cudaStream_t)stream);
    if (engine == COFI_GEMM_TF32X3S) {
        COFI_REQUIRE(ldw == K && gemm_tc_supported(lda, ldw, ldc, M, N, K, A, W, C),
                     "cofi_gemm: COFI_GEMM_TF32X3S needs W = cofi_split_tf32 output (ldw == K), N >= 16, 16-byte aligned operands");
        return gemm_x3_launch(A, lda, W, C, ldc, M, N, K, ep, (cudaStream_t)stream, nullptr);
    }
    if (engine == COFI_GEMM_TF32 || engine == COFI_GEMM_TF32X3) {
        if (!gemm_tc_supported(lda, ldw, ldc, M, N, K, A, W, C))
            return gemm_simt_launch(A, lda, W, ldw, C, ldc, M, N, K, ep, (cudaStream_t)stream);
        return gemm_tc_launch(A, lda, W, ldw, C, ldc, M, N, K, ep, engine, (cudaStream_t)stream);
    }
    set_error("cofi_gemm: unknown engine %d", engine);
    return COFI_EINVAL;
}

extern "C" int cofi_conv2d_nhwc(const float* x, int B, int H, int W, int Cin, const float* w, int Cout, int KH,
                                int KW, int stride, int pad, const float* scale, const float* shift,
                                const float* residual, int act, float* y, int engine, void* stream) {
    COFI_REQUIRE(x && w && y, "cofi_conv2d_nhwc: null pointer");
    COFI_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0,
                 "cofi_conv2d_nhwc: bad shape");
    COFI_REQUIRE(Cin % 4 == 0, "cofi_conv2d_nhwc: Cin=%d must be a multiple of 4 (pad the input)", Cin);
    COFI_REQUIRE((scale == nullptr) == (shift == nullptr), "cofi_conv2d_nhwc: scale and shift go together");
    COFI_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)w % 16) == 0, "cofi_conv2d_nhwc: 16-byte alignment");
    Epilogue ep{nullptr, nullptr, scale, shift, residual, Cout, 0, act, nullptr, nullptr, 0.0f};
    if (engine == COFI_GEMM_TF32X3S) {
        COFI_REQUIRE(conv_tc_supported(B, H, W, Cin, Cout, KH, KW, stride, pad),
                     "cofi_conv2d_nhwc: shape unsupported by the COFI_GEMM_TF32X3S engine");
        return conv_x3_launch(x, B, H, W, Cin, w, Cout, KH, KW, stride, pad, y, ep, (cudaStream_t)stream);
    }
    if ((engine == COFI_GEMM_TF32 || engine == COFI_GEMM_TF32X3) && conv_tc_supported(B, H, W, Cin, Cout, KH, KW, stride, pad))
        return conv_tc_launch(x, B, H, W, Cin, w, Cout, KH, KW, stride, pad, y, ep, engine, (cudaStream_t)stream);
    return conv_simt_launch(x, B, H, W, Cin, w, Cout, KH, KW, stride, pad, y, ep, (cudaStream_t)stream);
}

extern "C" int cofi_layer_norm_rows(const float* x, int64_t ldx, int64_t rows, int C, const float* gamma,
                                    const float* beta, float eps, int act, const float* residual, int64_t ldr,
                                    float* y, int64_t ldy, void* stream);

extern "C" int cofi_gemm_ln(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t M,
                            int N, int K, const float* bias, const float* gamma, const float* beta, float eps, int act,
                            const float* residual, int64_t ldr, int engine, void* stream) {
    COFI_REQUIRE(A && W && C && gamma && beta, "cofi_gemm_ln: null pointer");
    COFI_REQUIRE(M >= 0 && N > 0 && K > 0, "cofi_gemm_ln: bad shape");
    COFI_REQUIRE(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0 && lda >= K && ldw >= K && ldc >= N, "cofi_gemm_ln: bad leading dimension");
    COFI_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0, "cofi_gemm_ln: A and W must be 16-byte aligned");
    if (M == 0) return COFI_OK;
    if (engine == COFI_GEMM_TF32X3S) {
        COFI_REQUIRE(ldw == K && N <= 128 && N % 32 == 0 && gemm_tc_supported(lda, ldw, ldc, M, N, K, A, W, C),
                     "cofi_gemm_ln: COFI_GEMM_TF32X3S needs split weights (ldw == K), N <= 128, N % 32 == 0");
        Epilogue ep{bias, nullptr, nullptr, nullptr, residual, ldr, 0, act, gamma, beta, eps};
        return gemm_x3_launch(A, lda, W, C, ldc, M, N, K, ep, (cudaStream_t)stream, nullptr);
    }
    if ((engine == COFI_GEMM_TF32 || engine == COFI_GEMM_TF32X3) && N <= 128 && N % 32 == 0 &&
        gemm_tc_supported(lda, ldw, ldc, M, N, K, A, W, C)) {
        Epilogue ep{bias, nullptr, nullptr, nullptr, residual, ldr, 0, act, gamma, beta, eps};
        return gemm_tc_launch(A, lda, W, ldw, C, ldc, M, N, K, ep, engine, (cudaStream_t)stream);
    }
    int rc = cofi_gemm(A, lda, W, ldw, C, ldc, M, N, K, bias, nullptr, 0, COFI_ACT_NONE, engine, stream);
    if (rc) return rc;
    return cofi_layer_norm_rows(C, ldc, M, N, gamma, beta, eps, act, residual, ldr, C, ldc, stream);
}

extern "C" int cofi_gemm_f16(const void* A, int64_t lda, const void* W, int64_t ldw, float* C, int64_t ldc, int64_t M,
                             int N, int K, const float* bias, const float* rowdiv, int act, void* stream) {
    COFI_REQUIRE(A && W && C, "cofi_gemm_f16: null pointer");
    COFI_REQUIRE(M >= 0 && N >= 16 && K >= 8, "cofi_gemm_f16: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
    COFI_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K && ldc >= N,
                 "cofi_gemm_f16: K, lda, ldw must be multiples of 8 (16-byte rows)");
    COFI_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0, "cofi_gemm_f16: 16-byte alignment");
    if (M == 0) return COFI_OK;
    Epilogue ep{bias, rowdiv, nullptr, nullptr, nullptr, 0, 0, act, nullptr, nullptr, 0.0f};
    return gemm_tc_f16_launch(A, lda, W, ldw, C, ldc, M, N, K, ep, (cudaStream_t)stream);
}

// GEMM whose epilogue also emits per-128-row-tile column statistics (sum, sum of squares) of the stored output, so the
// GroupNorm that follows (cofi_norm_rows_pre) skips its statistics pass.  tensor-core engines only; M % 128 == 0.
static int gemm_colstats_impl(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc,
                              int64_t M, int N, int K, const float* bias, const float* rowdiv, int accumulate, int engine,
                              float* stats /* [M/128, N, 2] */, void* stream) {
    COFI_REQUIRE(A && W && C && stats, "cofi_gemm_colstats: null pointer");
    COFI_REQUIRE(M > 0 && M % 128 == 0 && N >= 16 && K > 0, "cofi_gemm_colstats: M must be a positive multiple of 128, N >= 16");
    COFI_REQUIRE(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0 && lda >= K && ldw >= K && ldc >= N, "cofi_gemm_colstats: bad leading dimension");
    COFI_REQUIRE(engine == COFI_GEMM_TF32 || engine == COFI_GEMM_TF32X3 || engine == COFI_GEMM_TF32X3S,
                 "cofi_gemm_colstats: tensor-core engines only");
    COFI_REQUIRE(gemm_tc_supported(lda, ldw, ldc, M, N, K, A, W, C), "cofi_gemm_colstats: shape/alignment unsupported");
    Epilogue ep{bias, rowdiv, nullptr, nullptr, nullptr, 0, accumulate, COFI_ACT_NONE, nullptr, nullptr, 0.0f};
    if (engine == COFI_GEMM_TF32X3S) {
        COFI_REQUIRE(ldw == K, "cofi_gemm_colstats: COFI_GEMM_TF32X3S needs W = cofi_split_tf32 output (ldw == K)");
        return gemm_x3_launch(A, lda, W, C, ldc, M, N, K, ep, (cudaStream_t)stream, stats);
    }
    return gemm_tc_launch(A, lda, W, ldw, C, ldc, M, N, K, ep, engine, (cudaStream_t)stream, stats);
}

extern "C" int cofi_gemm_colstats(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc,
                                  int64_t M, int N, int K, const float* bias, const float* rowdiv, int engine,
                                  float* stats /* [M/128, N, 2] */, void* stream) {
    return gemm_colstats_impl(A, lda, W, ldw, C, ldc, M, N, K, bias, rowdiv, 0, engine, stats, stream);
}

// C += A W^T + bias (C is read and rewritten), statistics of the stored sum: the second half of a Linear over a concatenated
// input whose first half was evaluated elsewhere (the decoder: W [up(x_c) | x_f] = up(W_c x_c) + W_f x_f)
extern "C" int cofi_gemm_colstats_acc(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc,
                                      int64_t M, int N, int K, const float* bias, int engine, float* stats, void* stream) {
    return gemm_colstats_impl(A, lda, W, ldw, C, ldc, M, N, K, bias, nullptr, 1, engine, stats, stream);
}

extern "C" int cofi_gemm_f16_colstats(const void* A, int64_t lda, const void* W, int64_t ldw, float* C, int64_t ldc,
                                      int64_t M, int N, int K, const float* bias, const float* rowdiv, float* stats,
                                      void* stream) {
    COFI_REQUIRE(A && W && C && stats, "cofi_gemm_f16_colstats: null pointer");
    COFI_REQUIRE(M > 0 && M % 128 == 0 && N >= 16 && K >= 8, "cofi_gemm_f16_colstats: M must be a positive multiple of 128");
    COFI_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K && ldc >= N, "cofi_gemm_f16_colstats: bad leading dimension");
    COFI_REQUIRE(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0, "cofi_gemm_f16_colstats: 16-byte alignment");
    Epilogue ep{bias, rowdiv, nullptr, nullptr, nullptr, 0, 0, COFI_ACT_NONE, nullptr, nullptr, 0.0f};
    return gemm_tc_f16_launch(A, lda, W, ldw, C, ldc, M, N, K, ep, (cudaStream_t)stream, stats);
}
