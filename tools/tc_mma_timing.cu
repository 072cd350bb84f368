// tc_mma_timing.cu -- how long do 24 dependent tcgen05.mma (M128,N64,K8,tf32) take, and does spreading them over
// several TMEM accumulators pipeline them?   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I hyphy_b200/csrc -o /tmp/tcm tools/tc_mma_timing.cu
#include <cstdio>
#include <vector>
#include "hb2_kernels_tc.cuh"
using namespace hb2;

// Variant 2: the CUTLASS idiom -- warp-uniform operands (values broadcast with __shfl_sync so that ptxas can keep them in
// uniform registers) and ONE elected lane issuing inside a warp-uniform branch.  The first variant issues from
// `if (lane == 0)`: ptxas then wraps every UTCHMMA in an ELECT / R2UR.BROADCAST waterfall loop (see the SASS), ~190 cycles each.
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred;
}
template <int NMMA>
__global__ void __launch_bounds__(128) k_uniform(const float *Bt, long long *out, int nissue) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float *Bs = reinterpret_cast<float *>(smem);
    uint64_t *bar_b = reinterpret_cast<uint64_t *>(smem + 2 * 32768);
    uint64_t *bar_mma = bar_b + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_b + 2);
    __shared__ int err;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) { mbar_init(bar_b, 1); mbar_init(bar_mma, nissue); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t z[64];
    for (int i = 0; i < 64; i++) z[i] = __float_as_uint(0.25f);
    for (int o = 0; o < 64; o += 16) { HB2_TMEM_ST16(lane_addr + 384 + o, z, o); }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (tid == 0) { mbar_expect_tx(bar_b, 32768u); bulk_g2s(Bs, Bt, 32768u, bar_b); }
    tc_fence_before(); __syncthreads();
    if (warp < nissue) {
        tc_fence_after();
        mbar_wait(bar_b, 0, &err);
        const uint64_t dh = make_b_desc(__shfl_sync(0xffffffffu, smem_u32(Bs), 0));
        for (int rep = 0; rep < 3; rep++) {
            long long t0 = clock64(), t1 = 0;
            if (elect_one()) {
#pragma unroll
                for (int m = 0; m < NMMA; m++)
                    tc_mma_tf32_ts(tmem_base + (uint32_t)warp * 64u, tmem_base + 384 + (m % 8) * 8, dh + (uint64_t)((m % 8) * 2 * 1024 >> 4), TC_IDESC, 1u);
                t1 = clock64();
                tc_commit(bar_mma);
            }
            __syncwarp();
            mbar_wait(bar_mma, rep & 1, &err);
            long long t2 = clock64();
            t1 = __shfl_sync(0xffffffffu, t1, 0) ? t1 : t1;
            if (tid == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

__global__ void __launch_bounds__(128) k(const float *Bt, long long *out, int nacc, int nmma, int a_from_smem) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float *Bs = reinterpret_cast<float *>(smem);
    uint64_t *bar_b = reinterpret_cast<uint64_t *>(smem + 2 * 32768);
    uint64_t *bar_mma = bar_b + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_b + 2);
    __shared__ int err;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(bar_b, 1); mbar_init(bar_mma, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t z[64];
    for (int i = 0; i < 64; i++) z[i] = __float_as_uint(0.25f);
    for (int o = 0; o < 64; o += 16) { HB2_TMEM_ST16(lane_addr + 384 + o, z, o); }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (tid == 0) { mbar_expect_tx(bar_b, 32768u); bulk_g2s(Bs, Bt, 32768u, bar_b); }
    tc_fence_before(); __syncthreads();
    // a_from_smem is reused as the number of issuing threads (lane 0 of warps 0..n-1); bar_mma expects n commits
    const int nissue = a_from_smem < 1 ? 1 : a_from_smem;
    if (tid == 0) { mbar_init(bar_mma, nissue); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if ((tid & 31) == 0 && warp < nissue) {
        tc_fence_after();
        mbar_wait(bar_b, 0, &err);
        const uint64_t dh = make_b_desc(smem_u32(Bs));
        for (int rep = 0; rep < 3; rep++) {
            long long t0 = clock64();
#pragma unroll 1
            for (int m = warp; m < nmma; m += nissue)
                tc_mma_tf32_ts(tmem_base + (uint32_t)(m % nacc) * 64u, tmem_base + 384 + (m % 8) * 8, dh + (uint64_t)((m % 8) * 2 * 1024 >> 4), TC_IDESC, 1u);
            long long t1 = clock64();
            tc_commit(bar_mma);
            mbar_wait(bar_mma, rep & 1, &err);
            long long t2 = clock64();
            if (tid == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

int main() {
    std::vector<float> Bt(8192, 0.01f);
    float *dB; long long *dO;
    cudaMalloc(&dB, 8192 * 4); cudaMalloc(&dO, 64);
    cudaMemcpy(dB, Bt.data(), 8192 * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    for (int nmma : {24, 48})
        for (int nacc : {1, 4})
        for (int nissue : {1, 2, 4}) {
            k<<<1, 128, TC_SMEM_BYTES>>>(dB, dO, nacc, nmma, nissue);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[6]; cudaMemcpy(h, dO, 48, cudaMemcpyDeviceToHost);
            printf("issuers=%d nmma=%d accumulators=%d : issue %lld cyc, issue->complete %lld cyc (%.1f per MMA)  [%s]\n", nissue, nmma, nacc, h[4], h[5], (double)h[5] / nmma, cudaGetErrorString(e));
        }
    cudaFuncSetAttribute(k_uniform<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    cudaFuncSetAttribute(k_uniform<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    for (int nissue : {1, 2}) {
        if (nissue == 1) k_uniform<24><<<1, 128, TC_SMEM_BYTES>>>(dB, dO, nissue); else k_uniform<12><<<1, 128, TC_SMEM_BYTES>>>(dB, dO, nissue);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[6]; cudaMemcpy(h, dO, 48, cudaMemcpyDeviceToHost);
        printf("UNIFORM idiom: issuers=%d, %d MMAs each : issue %lld cyc (elected lane may not be lane 0: 0 = not lane 0), issue->complete %lld cyc  [%s]\n", nissue, nissue == 1 ? 24 : 12, h[4], h[5], cudaGetErrorString(e));
    }
    return 0;
}
