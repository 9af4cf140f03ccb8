// sample.cu -- random half-sampling of the point pyramid on the device
// (reference model/kpconv/preprocess_data.py:52-68: level i+1 = level i[np.random.choice(n, n // 2)], WITH replacement).
// The reference draws from numpy's global Mersenne Twister on the host; a device sampler cannot reproduce that stream,
// so the draw is defined here as a counter-based generator the host can restate exactly (oracle/knn.py::
// half_sample_pyramid_philox):  u = Philox4x32-10(counter = (j, frame, level, 0), key = seed)[0],
// index = (u * n) >> 32  (Lemire's multiply-shift range reduction), independently for every output row j.
// One launch builds every level: row j of level l walks its chain of draws back to level 0 and copies that point.
#include "common.cuh"

namespace cofi {

__host__ __device__ __forceinline__ uint32_t philox4x32_10_word0(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                                 uint32_t k0, uint32_t k1) {
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

constexpr int HS_MAXL = 8;
struct HalfSampleParams {
    const float* pts0;
    float* out[HS_MAXL];       // out[l] for l = 1..levels-1: [frames * (n0 >> l), 3]
    int64_t* index[HS_MAXL];   // optional: frame-local level-0 row each output row was copied from
    int64_t n0;
    int frames, levels;
    uint32_t k0, k1;
};

__global__ void __launch_bounds__(256)
half_sample_kernel(const HalfSampleParams p) {
    const int level = blockIdx.y + 1, frame = blockIdx.z;
    const int64_t nl = p.n0 >> level;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nl; j += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = j;
        for (int t = level; t >= 1; --t) {  // row i of level t was drawn from level t-1 (n >> (t-1) rows)
            const uint32_t u = philox4x32_10_word0((uint32_t)i, (uint32_t)frame, (uint32_t)t, 0u, p.k0, p.k1);
            i = (int64_t)(((uint64_t)u * (uint64_t)(p.n0 >> (t - 1))) >> 32);
        }
        const float* s = p.pts0 + ((int64_t)frame * p.n0 + i) * 3;
        float* d = p.out[level] + ((int64_t)frame * nl + j) * 3;
        d[0] = s[0], d[1] = s[1], d[2] = s[2];
        if (p.index[level]) p.index[level][(int64_t)frame * nl + j] = i;
    }
}

}  // namespace cofi

using namespace cofi;

extern "C" int cofi_half_sample_pyramid(const float* pts0, int64_t n0, int frames, int levels, uint64_t seed,
                                        float* const* out_levels, int64_t* const* out_index, void* stream) {
    COFI_REQUIRE(pts0 && out_levels && n0 > 0 && frames > 0 && frames < 65536 && levels >= 1 && levels <= HS_MAXL,
                 "cofi_half_sample_pyramid: bad argument");
    COFI_REQUIRE((n0 >> (levels - 1)) >= 1 && n0 < (1ll << 32), "cofi_half_sample_pyramid: level sizes out of range");
    if (levels == 1) return COFI_OK;
    HalfSampleParams p{};
    p.pts0 = pts0;
    p.n0 = n0;
    p.frames = frames;
    p.levels = levels;
    p.k0 = (uint32_t)seed;
    p.k1 = (uint32_t)(seed >> 32);
    for (int l = 1; l < levels; ++l) {
        COFI_REQUIRE(out_levels[l] != nullptr, "cofi_half_sample_pyramid: null output level");
        p.out[l] = out_levels[l];
        p.index[l] = out_index ? out_index[l] : nullptr;
    }
    const unsigned bx = (unsigned)ceil_div(n0 >> 1, 256);
    half_sample_kernel<<<dim3(bx, levels - 1, frames), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("cofi_half_sample_pyramid");
}
