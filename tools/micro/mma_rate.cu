// Micro-benchmark: clocks per tcgen05.mma (kind::tf32 / kind::f16, M=128) issued back to back by one thread per CTA, operands
// from shared memory (SS) or A from tensor memory (TS), N = 64/128/256; one CTA per SM.  Data is whatever shared/tensor memory
// holds (zeros): only the issue/execute rate is measured.   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../cofii2p_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace cofi::tc;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, int f16) {
    if (f16)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}

__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred;
}

// variant: the whole warp runs the loop (provably warp-uniform control flow), one elected lane issues
template <int N>
__global__ void __launch_bounds__(128, 1) rate_uniform_kernel(int mode, int f16, int iters, int nacc, unsigned long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc(f16 ? 0 : 2, 128, N);
        const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 16384;
        const long long t0 = clock64();
        uint32_t acc = 0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t bd = umma_desc_k128(b_addr + k * 32);
                const uint32_t d = tm + acc * N;
                acc = (acc + 1 == (uint32_t)nacc) ? 0 : acc + 1;
                if (elect_one()) {
                    if (mode == 0) {
                        const uint64_t ad = umma_desc_k128(a_addr + k * 32);
                        if (f16) mma_f16(d, ad, bd, idesc, 1u); else mma_tf32(d, ad, bd, idesc, 1u);
                    } else {
                        mma_ts(d, tm + 256 + k * 8, bd, idesc, f16);
                    }
                }
                __syncwarp();
            }
        }
        if (elect_one()) tc_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode /*0 SS, 1 TS*/, int f16, int iters, int nacc, unsigned long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc(f16 ? 0 : 2, 128, N);
        const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 16384;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t bd = umma_desc_k128(b_addr + k * 32);
                const uint32_t d = tm + (uint32_t)((it * 4 + k) % nacc) * N;
                if (mode == 0) {
                    const uint64_t ad = umma_desc_k128(a_addr + k * 32);
                    if (f16) mma_f16(d, ad, bd, idesc, 1u); else mma_tf32(d, ad, bd, idesc, 1u);
                } else {
                    mma_ts(d, tm + 256 + k * 8, bd, idesc, f16);
                }
            }
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

// issue-style variants with compile-time operand mode: STYLE 0 = elect per MMA (CUTLASS), 1 = elect once outside the loop and
// predicate every MMA on it, 2 = elect once per 4-MMA block, 3 = style 1 with 12 MMAs per iteration alternating TS operands as
// the 3xTF32 main loop does
template <int N, int TS, int STYLE>
__global__ void __launch_bounds__(128, 1) rate_style_kernel(int iters, unsigned long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 2 * N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 0) {
        constexpr uint32_t idesc = umma_idesc(2, 128, N);
        const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 16384;
        const bool leader = elect_one() != 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (STYLE == 3) {
                if (leader) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t bh = umma_desc_k128(b_addr + k * 32), bl = umma_desc_k128(b_addr + N * 128 + k * 32);
                        mma_ts(tm, tm + 256 + k * 8, bh, idesc, 0);
                        mma_ts(tm, tm + 288 + k * 8, bh, idesc, 0);
                        mma_ts(tm, tm + 256 + k * 8, bl, idesc, 0);
                    }
                }
            } else if (STYLE == 2) {
                if (leader) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t bd = umma_desc_k128(b_addr + k * 32);
                        if (TS) mma_ts(tm, tm + 256 + k * 8, bd, idesc, 0);
                        else mma_tf32(tm, umma_desc_k128(a_addr + k * 32), bd, idesc, 1u);
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t bd = umma_desc_k128(b_addr + k * 32);
                    const bool go = STYLE == 0 ? (elect_one() != 0) : leader;
                    if (go) {
                        if (TS) mma_ts(tm, tm + 256 + k * 8, bd, idesc, 0);
                        else mma_tf32(tm, umma_desc_k128(a_addr + k * 32), bd, idesc, 1u);
                    }
                }
            }
        }
        if (leader) tc_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int N, int TS, int STYLE>
void run_s() {
    unsigned long long* d;
    cudaMalloc(&d, 8);
    const int smem = 16384 + 2 * N * 128 + 2048, iters = 2000;
    cudaFuncSetAttribute(rate_style_kernel<N, TS, STYLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_style_kernel<N, TS, STYLE><<<148, 128, smem>>>(iters, d);
    rate_style_kernel<N, TS, STYLE><<<148, 128, smem>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const char* names[4] = {"elect per MMA", "elect once, predicate per MMA", "elect once, 4 MMAs per block", "elect once, 12 TS MMAs per block (3xTF32 pattern)"};
    printf("{\"issue\": \"%s\", \"kind\": \"tf32\", \"mode\": \"%s\", \"N\": %d, \"ctas\": 148, \"clk_per_mma\": %.1f, \"err\": \"%s\"}\n",
           names[STYLE], TS ? "TS" : "SS", N, (double)h / (iters * (STYLE == 3 ? 12.0 : 4.0)), cudaGetErrorString(e));
    cudaFree(d);
}

template <int N>
void run_u(int mode, int f16, int nacc, int grid) {
    unsigned long long* d;
    cudaMalloc(&d, 8);
    const int smem = 16384 + N * 128 + 2048, iters = 2000;
    cudaFuncSetAttribute(rate_uniform_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_uniform_kernel<N><<<grid, 128, smem>>>(mode, f16, iters, nacc, d);
    rate_uniform_kernel<N><<<grid, 128, smem>>>(mode, f16, iters, nacc, d);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("{\"issue\": \"warp-uniform + elect\", \"kind\": \"%s\", \"mode\": \"%s\", \"N\": %d, \"accumulators\": %d, \"ctas\": %d, \"clk_per_mma\": %.1f, \"err\": \"%s\"}\n",
           f16 ? "f16" : "tf32", mode ? "TS" : "SS", N, nacc, grid, (double)h / (iters * 4.0), cudaGetErrorString(e));
    cudaFree(d);
}

template <int N>
void run(int mode, int f16, int nacc, int grid) {
    unsigned long long* d;
    cudaMalloc(&d, 8);
    const int smem = 16384 + N * 128 + 2048, iters = 2000;
    cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_kernel<N><<<grid, 128, smem>>>(mode, f16, iters, nacc, d);
    rate_kernel<N><<<grid, 128, smem>>>(mode, f16, iters, nacc, d);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("{\"issue\": \"one thread in divergent code\", \"kind\": \"%s\", \"mode\": \"%s\", \"N\": %d, \"accumulators\": %d, \"ctas\": %d, \"clk_per_mma\": %.1f, \"err\": \"%s\"}\n",
           f16 ? "f16" : "tf32", mode ? "TS" : "SS", N, nacc, grid, (double)h / (iters * 4.0), cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run_s<128, 0, 0>(); run_s<128, 0, 1>(); run_s<128, 0, 2>();
    run_s<128, 1, 0>(); run_s<128, 1, 1>(); run_s<128, 1, 2>(); run_s<128, 1, 3>();
    run_s<64, 1, 2>(); run_s<64, 1, 3>(); run_s<64, 0, 2>();
    run_s<256, 0, 2>(); run_s<256, 1, 2>();
    for (int grid : {148})
        for (int f16 = 0; f16 < 2; ++f16)
            for (int mode = 0; mode < 2; ++mode) {
                run<64>(mode, f16, 1, grid);
                run<128>(mode, f16, 1, grid);
                run<256>(mode, f16, 1, grid);
                run_u<64>(mode, f16, 1, grid);
                run_u<128>(mode, f16, 1, grid);
                run_u<128>(mode, f16, 2, grid);
                run_u<256>(mode, f16, 1, grid);
            }
    return 0;
}
