// pnp.cu -- batched P3P-RANSAC on the device: the hypothesis loop of `cv2.solvePnPRansac(..., iterationsCount=10000)`
// that reference evaluation/eval_all.py:107 runs on the CPU for every frame.
//
// OpenCV draws minimal sets sequentially from cv::RNG and stops early on a confidence bound; here ALL `iterations`
// hypotheses of all frames are evaluated at once, one thread each, and the sampling is counter-based so that the host
// oracle (oracle/pnp.py) restates it exactly:
//   hypothesis h: indices i_k = (Philox4x32-10(h, k, 0, 0; seed)[0] * n) >> 32, k = 0..3; a repeated index skips h
//   P3P on points 0..2 (Grunert's distance equations; s2 = u s1, s3 = v s1; u eliminated -> quartic in v, coefficients
//   derived symbolically -- see oracle/pnp.py::quartic_coeffs), closed-form quartic (Ferrari) + Newton polish, pose from
//   aligning the two triangles; the root with the smallest reprojection error on point 3 is kept
//   score = #{points: squared reprojection error <= thr^2, depth > 0}; winner = max score, lowest h on ties
//   (one 64-bit atomicMax per CTA on (score << 32 | ~h): deterministic)
// A second tiny kernel re-solves the winning hypothesis (same code, same result) and writes its pose and inlier mask.
// The final refinement on the inliers stays OpenCV's own solvePnP(ITERATIVE), as in solvePnPRansac -- identical inlier
// sets therefore give the identical pose.  Double precision throughout: 1e4 hypotheses x a few hundred points is tiny.
#include "common.cuh"

namespace cofi {
namespace {

__host__ __device__ __forceinline__ uint32_t philox_w0(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c0;
}

struct Pose {
    double R[9];
    double t[3];
};

// largest real root of m^3 + A m^2 + B m + C
__device__ double cubic_largest_root(double A, double B, double C) {
    const double Q = (A * A - 3.0 * B) / 9.0, R = (2.0 * A * A * A - 9.0 * A * B + 27.0 * C) / 54.0;
    double m;
    if (R * R < Q * Q * Q) {
        const double th = acos(fmin(fmax(R / sqrt(Q * Q * Q), -1.0), 1.0));
        m = -2.0 * sqrt(Q) * cos((th + 6.283185307179586) / 3.0) - A / 3.0;
    } else {
        const double a = -copysign(cbrt(fabs(R) + sqrt(fmax(R * R - Q * Q * Q, 0.0))), R);
        const double b = a != 0.0 ? Q / a : 0.0;
        m = a + b - A / 3.0;
    }
    for (int it = 0; it < 3; ++it) {  // Newton polish
        const double f = ((m + A) * m + B) * m + C, df = (3.0 * m + 2.0 * A) * m + B;
        if (df != 0.0) m -= f / df;
    }
    return m;
}

// real roots of q4 x^4 + q3 x^3 + q2 x^2 + q1 x + q0 (q4 != 0); returns how many
__device__ int quartic_real_roots(const double* q, double* x) {
    const double a = q[3] / q[4], b = q[2] / q[4], c = q[1] / q[4], d = q[0] / q[4];
    const double a2 = a * a;
    const double p = b - 0.375 * a2, qq = c - 0.5 * a * b + 0.125 * a2 * a;
    const double r = d - 0.25 * a * c + 0.0625 * a2 * b - (3.0 / 256.0) * a2 * a2;
    int n = 0;
    const double scale = fabs(p) + fabs(r) + 1.0;
    if (fabs(qq) < 1e-14 * scale) {  // biquadratic
        const double disc = p * p - 4.0 * r;
        if (disc >= 0.0) {
            const double sd = sqrt(disc);
            for (int s = 0; s < 2; ++s) {
                const double y2 = 0.5 * (-p + (s ? -sd : sd));
                if (y2 >= 0.0) {
                    x[n++] = sqrt(y2) - 0.25 * a;
                    x[n++] = -sqrt(y2) - 0.25 * a;
                }
            }
        }
    } else {
        const double m = cubic_largest_root(p, 0.25 * p * p - r, -0.125 * qq * qq);
        if (m > 0.0) {
            const double s = sqrt(2.0 * m);
            const double tol = 1e-12 * (fabs(p) + fabs(m) + fabs(qq / s) + 1.0);
            const double Dp = -2.0 * p - 2.0 * m - 2.0 * qq / s;   // y^2 - s y + (p/2 + m + q/(2s)) = 0
            const double Dm = -2.0 * p - 2.0 * m + 2.0 * qq / s;   // y^2 + s y + (p/2 + m - q/(2s)) = 0
            if (Dp >= -tol) {
                const double sd = sqrt(fmax(Dp, 0.0));
                x[n++] = 0.5 * (s + sd) - 0.25 * a;
                x[n++] = 0.5 * (s - sd) - 0.25 * a;
            }
            if (Dm >= -tol) {
                const double sd = sqrt(fmax(Dm, 0.0));
                x[n++] = 0.5 * (-s + sd) - 0.25 * a;
                x[n++] = 0.5 * (-s - sd) - 0.25 * a;
            }
        }
    }
    for (int i = 0; i < n; ++i)  // Newton polish on the normalised quartic
        for (int it = 0; it < 2; ++it) {
            const double v = x[i];
            const double f = (((v + a) * v + b) * v + c) * v + d, df = ((4.0 * v + 3.0 * a) * v + 2.0 * b) * v + c;
            if (df != 0.0) x[i] = v - f / df;
        }
    return n;
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ bool normalize3(double* v) {
    const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (!(n > 1e-300)) return false;
    v[0] /= n, v[1] /= n, v[2] /= n;
    return true;
}
// orthonormal frame (columns e1, e2, e3) of a triangle T[3][3] (rows = points)
__device__ bool tri_frame(const double T[3][3], double E[3][3]) {
    double e1[3] = {T[1][0] - T[0][0], T[1][1] - T[0][1], T[1][2] - T[0][2]};
    double w[3] = {T[2][0] - T[0][0], T[2][1] - T[0][1], T[2][2] - T[0][2]};
    double e3[3], e2[3];
    if (!normalize3(e1)) return false;
    cross3(e1, w, e3);
    if (!normalize3(e3)) return false;
    cross3(e3, e1, e2);
    for (int i = 0; i < 3; ++i) {
        E[i][0] = e1[i];
        E[i][1] = e2[i];
        E[i][2] = e3[i];
    }
    return true;
}

struct Cam {
    double fx, fy, cx, cy;
};

// squared reprojection error of world point X under pose; +inf when behind the camera
__device__ __forceinline__ double reproj_err2(const Pose& P, const Cam& k, const float* X, const float* uv) {
    const double x = X[0], y = X[1], z = X[2];
    const double xc = P.R[0] * x + P.R[1] * y + P.R[2] * z + P.t[0];
    const double yc = P.R[3] * x + P.R[4] * y + P.R[5] * z + P.t[1];
    const double zc = P.R[6] * x + P.R[7] * y + P.R[8] * z + P.t[2];
    if (!(zc > 0.0)) return INFINITY;
    const double du = xc / zc * k.fx + k.cx - (double)uv[0], dv = yc / zc * k.fy + k.cy - (double)uv[1];
    return du * du + dv * dv;
}

// hypothesis h -> pose (false: skipped).  obj [n][3], img [n][2] (shared or global memory)
__device__ bool solve_hypothesis(uint32_t h, int n, uint32_t k0, uint32_t k1, const Cam& cam, const float* obj, const float* img,
                                 Pose& best) {
    int idx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) idx[k] = (int)(((uint64_t)philox_w0(h, (uint32_t)k, 0u, 0u, k0, k1) * (uint64_t)n) >> 32);
    if (idx[0] == idx[1] || idx[0] == idx[2] || idx[0] == idx[3] || idx[1] == idx[2] || idx[1] == idx[3] || idx[2] == idx[3])
        return false;
    double f[3][3], P[3][3];
    for (int i = 0; i < 3; ++i) {
        const double bx = ((double)img[idx[i] * 2] - cam.cx) / cam.fx, by = ((double)img[idx[i] * 2 + 1] - cam.cy) / cam.fy;
        const double nrm = sqrt(bx * bx + by * by + 1.0);
        f[i][0] = bx / nrm, f[i][1] = by / nrm, f[i][2] = 1.0 / nrm;
        for (int d = 0; d < 3; ++d) P[i][d] = (double)obj[idx[i] * 3 + d];
    }
    const double c12 = f[0][0] * f[1][0] + f[0][1] * f[1][1] + f[0][2] * f[1][2];
    const double c13 = f[0][0] * f[2][0] + f[0][1] * f[2][1] + f[0][2] * f[2][2];
    const double c23 = f[1][0] * f[2][0] + f[1][1] * f[2][1] + f[1][2] * f[2][2];
    auto d2 = [&](int i, int j) {
        const double x = P[i][0] - P[j][0], y = P[i][1] - P[j][1], z = P[i][2] - P[j][2];
        return x * x + y * y + z * z;
    };
    const double a = d2(0, 1), b = d2(0, 2), c = d2(1, 2);
    if (fmin(a, fmin(b, c)) < 1e-12) return false;
    // quartic in v (sympy: resultant of the two distance-ratio quadrics after eliminating u), oracle/pnp.py::quartic_coeffs
    double q[5];
    {
        const double x0 = 2 * c, x1 = b * x0, x2 = a * b, x3 = 2 * x2, x4 = c23 * c23, x5 = 4 * x2, x6 = x4 * x5;
        const double x7 = a * a, x8 = b * b, x9 = c * c, x10 = -a * x0 + x7 + x8 + x9, x11 = 4 * c13, x12 = x11 * x2;
        const double x13 = b * c, x14 = x11 * x13, x15 = 8 * c13, x16 = x15 * x2, x17 = c12 * c23, x18 = a * c, x19 = 8 * x18;
        const double x20 = 4 * x13, x21 = c13 * x19 - x11 * x7 - x11 * x9 + x17 * x20 + x17 * x5 - 4 * x17 * x8;
        const double x22 = c13 * c13, x23 = c12 * c12, x24 = x20 * x23, x25 = x13 * x15;
        q[4] = -x1 + x10 + x3 - x6;
        q[3] = -x12 + x14 + x16 * x4 + x21;
        q[2] = -x16 * x17 - x17 * x25 - 4 * x18 - x19 * x22 + 4 * x22 * x7 + 4 * x22 * x9 + 4 * x23 * x8 - x24 + 4 * x4 * x8 - x6 +
               2 * x7 - 2 * x8 + 2 * x9;
        q[1] = x12 - x14 + x21 + x23 * x25;
        q[0] = x1 + x10 - x24 - x3;
    }
    if (q[4] == 0.0) return false;
    double roots[4];
    const int nr = quartic_real_roots(q, roots);
    double Ew[3][3];
    if (!tri_frame(P, Ew)) return false;
    bool found = false;
    double best_err = INFINITY;
    for (int i = 0; i < nr; ++i) {
        const double v = roots[i];
        if (!(v > 0.0)) continue;
        const double L = 2 * b * c * (c23 * v - c12);
        if (!(fabs(L) > 1e-300)) continue;
        const double M = c * (-a * v * v + 2 * a * c13 * v - a - b * v * v + b + c * v * v + c - 2 * c * c13 * v);
        const double u = -M / L;
        if (!(u > 0.0)) continue;
        const double den = u * u + v * v - 2 * u * v * c23;
        if (!(den > 0.0)) continue;
        const double s1 = sqrt(c / den), s2 = u * s1, s3 = v * s1;
        double X[3][3];
        for (int d = 0; d < 3; ++d) {
            X[0][d] = s1 * f[0][d];
            X[1][d] = s2 * f[1][d];
            X[2][d] = s3 * f[2][d];
        }
        auto x2d = [&](int i0, int j0) {
            const double x = X[i0][0] - X[j0][0], y = X[i0][1] - X[j0][1], z = X[i0][2] - X[j0][2];
            return x * x + y * y + z * z;
        };
        if (fabs(x2d(0, 1) - a) > 1e-6 * a || fabs(x2d(0, 2) - b) > 1e-6 * b) continue;  // spurious root of the eliminated system
        double Ec[3][3];
        if (!tri_frame(X, Ec)) continue;
        Pose cand;
        for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc)
                cand.R[r * 3 + cc] = Ec[r][0] * Ew[cc][0] + Ec[r][1] * Ew[cc][1] + Ec[r][2] * Ew[cc][2];
        for (int r = 0; r < 3; ++r)
            cand.t[r] = X[0][r] - (cand.R[r * 3] * P[0][0] + cand.R[r * 3 + 1] * P[0][1] + cand.R[r * 3 + 2] * P[0][2]);
        const double e = reproj_err2(cand, cam, obj + idx[3] * 3, img + idx[3] * 2);
        if (e < best_err) {
            best_err = e;
            best = cand;
            found = true;
        }
    }
    return found;
}

constexpr int PNP_MAXN = 2048;  // points of one frame staged in shared memory (20 bytes each)

__global__ void __launch_bounds__(128)
pnp_score_kernel(const float* __restrict__ img, const float* __restrict__ obj, const int32_t* __restrict__ count, int count_stride,
                 int n_max, const float* __restrict__ Kc, int iterations, double thr2, uint32_t k0, uint32_t k1,
                 unsigned long long* __restrict__ best_key) {
    extern __shared__ float sm[];
    float* s_obj = sm;                  // [n][3]
    float* s_img = sm + 3 * n_max;      // [n][2]
    const int f = blockIdx.y;
    const int n = min(count ? count[f * count_stride] : n_max, n_max);
    for (int i = threadIdx.x; i < n * 3; i += blockDim.x) s_obj[i] = obj[(int64_t)f * n_max * 3 + i];
    for (int i = threadIdx.x; i < n * 2; i += blockDim.x) s_img[i] = img[(int64_t)f * n_max * 2 + i];
    __syncthreads();
    const Cam cam{(double)Kc[f * 4], (double)Kc[f * 4 + 1], (double)Kc[f * 4 + 2], (double)Kc[f * 4 + 3]};
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = 0ull;
    if (h < (uint32_t)iterations && n >= 4) {
        Pose P;
        if (solve_hypothesis(h, n, k0, k1, cam, s_obj, s_img, P)) {
            int cnt = 0;
            for (int p = 0; p < n; ++p) cnt += reproj_err2(P, cam, s_obj + p * 3, s_img + p * 2) <= thr2 ? 1 : 0;
            key = ((unsigned long long)cnt << 32) | (unsigned long long)(0xFFFFFFFFu - h);   // max count, then lowest h
        }
    }
    // block maximum, then one atomic per CTA
    __shared__ unsigned long long red[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 4; ++w) key = red[w] > key ? red[w] : key;
        if (key) atomicMax(best_key + f, key);
    }
}

__global__ void __launch_bounds__(128)
pnp_emit_kernel(const float* __restrict__ img, const float* __restrict__ obj, const int32_t* __restrict__ count, int count_stride,
                int n_max, const float* __restrict__ Kc, double thr2, uint32_t k0, uint32_t k1,
                const unsigned long long* __restrict__ best_key, int32_t* __restrict__ out_count, int32_t* __restrict__ out_hyp,
                double* __restrict__ out_pose, uint8_t* __restrict__ out_inlier) {
    __shared__ Pose s_pose;
    __shared__ int s_ok;
    const int f = blockIdx.x;
    const int n = min(count ? count[f * count_stride] : n_max, n_max);
    const float* fo = obj + (int64_t)f * n_max * 3;
    const float* fi = img + (int64_t)f * n_max * 2;
    const Cam cam{(double)Kc[f * 4], (double)Kc[f * 4 + 1], (double)Kc[f * 4 + 2], (double)Kc[f * 4 + 3]};
    const unsigned long long key = best_key[f];
    if (threadIdx.x == 0) {
        s_ok = 0;
        if (key) {
            const uint32_t h = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
            s_ok = solve_hypothesis(h, n, k0, k1, cam, fo, fi, s_pose) ? 1 : 0;
            out_hyp[f] = (int32_t)h;
        } else {
            out_hyp[f] = -1;
        }
        out_count[f] = s_ok ? (int32_t)(key >> 32) : 0;
        for (int i = 0; i < 12; ++i) out_pose[f * 12 + i] = s_ok ? (i < 9 ? s_pose.R[i] : s_pose.t[i - 9]) : (i % 4 == 0 && i < 9 ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int p = threadIdx.x; p < n_max; p += blockDim.x)
        out_inlier[(int64_t)f * n_max + p] = (s_ok && p < n && reproj_err2(s_pose, cam, fo + p * 3, fi + p * 2) <= thr2) ? 1 : 0;
}

}  // namespace
}  // namespace cofi

using namespace cofi;

extern "C" int cofi_pnp_ransac(const float* image_points, const float* object_points, const int32_t* count, int count_stride,
                               int n_max, int frames, const float* cam, int iterations, float reproj_threshold, uint64_t seed,
                               int32_t* out_count, int32_t* out_hypothesis, double* out_pose, uint8_t* out_inlier, void* work,
                               void* stream) {
    COFI_REQUIRE(image_points && object_points && cam && out_count && out_hypothesis && out_pose && out_inlier && work,
                 "cofi_pnp_ransac: null pointer");
    COFI_REQUIRE(n_max >= 4 && n_max <= PNP_MAXN && frames > 0 && frames < 65536 && iterations > 0,
                 "cofi_pnp_ransac: bad shape (4 <= n_max <= %d)", PNP_MAXN);
    COFI_REQUIRE(((uintptr_t)work % 8) == 0, "cofi_pnp_ransac: workspace must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* best = reinterpret_cast<unsigned long long*>(work);   // [frames]
    cudaError_t e = cudaMemsetAsync(best, 0, sizeof(unsigned long long) * frames, st);
    if (e != cudaSuccess) {
        set_error("cofi_pnp_ransac: memset: %s", cudaGetErrorString(e));
        return COFI_ECUDA;
    }
    const int smem = n_max * 5 * (int)sizeof(float);
    const double thr2 = (double)reproj_threshold * (double)reproj_threshold;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    dim3 grid((unsigned)ceil_div(iterations, 128), (unsigned)frames);
    pnp_score_kernel<<<grid, 128, smem, st>>>(image_points, object_points, count, count_stride, n_max, cam, iterations, thr2, k0,
                                             k1, best);
    int rc = check_launch("cofi_pnp_ransac(score)");
    if (rc) return rc;
    pnp_emit_kernel<<<frames, 128, 0, st>>>(image_points, object_points, count, count_stride, n_max, cam, thr2, k0, k1, best,
                                            out_count, out_hypothesis, out_pose, out_inlier);
    return check_launch("cofi_pnp_ransac(emit)");
}
