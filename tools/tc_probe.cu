// tc_probe.cu -- standalone bring-up probe for the tcgen05 building blocks used by prune64_tc_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I hyphy_b200/csrc -o /tmp/tc_probe tools/tc_probe.cu && /tmp/tc_probe
// Tests: (1) TMEM st/ld round trip; (2) bulk copy into smem; (3) one K=8 MMA, A from TMEM, B from smem, for both
// LBO/SBO assignments; (4) full 3xTF32 K=64 contraction vs fp64: max and MEAN SIGNED relative error (accumulation bias).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "hb2_kernels_tc.cuh"

using namespace hb2;

__device__ __forceinline__ uint64_t make_b_desc2(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// mode 0: roundtrip; mode 1: K=8 single MMA with (lbo,sbo); mode 2: full split contraction
__global__ void __launch_bounds__(128) probe_kernel(int mode, const float *A /*128x64*/, const float *Bt /*canonical hi|lo 8192*/,
                                                     float *out /*128x64*/, uint32_t lbo, uint32_t sbo, int nsteps, int order, int *err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float *Bs = reinterpret_cast<float *>(smem);
    uint64_t *bar_b = reinterpret_cast<uint64_t *>(smem + 2 * 32768);
    uint64_t *bar_mma = bar_b + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_b + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(bar_b, 1); mbar_init(bar_mma, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    if (tid == 0 && blockIdx.x == 0) printf("tmem_base=0x%08x smem=0x%08x\n", tmem_base, smem_u32(Bs));
    uint32_t hi[64], lo[64];
    for (int k = 0; k < 64; k++) {
        float x = A[tid * 64 + k];
        float h = (mode == 0) ? x : tf32_rn(x);
        hi[k] = __float_as_uint(h);
        lo[k] = __float_as_uint(x - h);
    }
#pragma unroll
    for (int o = 0; o < 64; o += 16) { HB2_TMEM_ST16(lane_addr + 64 + o, hi, o); HB2_TMEM_ST16(lane_addr + 128 + o, lo, o); }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (mode == 0) {
        uint32_t d[64];
#pragma unroll
        for (int o = 0; o < 64; o += 16) HB2_TMEM_LD16(lane_addr + 64 + o, d, o);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int k = 0; k < 64; k++) out[tid * 64 + k] = __uint_as_float(d[k]);
    } else {
        if (tid == 0) { mbar_expect_tx(bar_b, 32768u); bulk_g2s(Bs, Bt, 32768u, bar_b); }
        tc_fence_before(); __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            mbar_wait(bar_b, 0, err);
            const uint64_t dh = make_b_desc2(smem_u32(Bs), lbo, sbo), dl = make_b_desc2(smem_u32(Bs + 4096), lbo, sbo);
            const uint32_t kstep = (2 * lbo) >> 4;       // two K chunks per MMA
            if (mode == 1) {
                for (int kk = 0; kk < nsteps; kk++) tc_mma_tf32_ts(tmem_base, tmem_base + 64 + kk * 8, dh + (uint64_t)(kk * kstep), TC_IDESC, kk > 0);
            } else {
                bool first = true;
                auto run = [&](uint32_t a_col, uint64_t d) { for (int kk = 0; kk < 8; kk++) { tc_mma_tf32_ts(tmem_base, tmem_base + a_col + kk * 8, d + (uint64_t)(kk * kstep), TC_IDESC, first ? 0u : 1u); first = false; } };
                if (order == 0) { run(128, dh); run(64, dl); run(64, dh); }       // small terms first
                else if (order == 1) { run(64, dh); run(64, dl); run(128, dh); }  // big term first
                else { run(64, dh); }                                              // plain TF32 (hi*hi only)
            }
            tc_commit(bar_mma);
        }
        __syncwarp();
        mbar_wait(bar_mma, 0, err);
        tc_fence_after();
        uint32_t d[64];
#pragma unroll
        for (int o = 0; o < 64; o += 16) HB2_TMEM_LD16(lane_addr + o, d, o);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int k = 0; k < 64; k++) out[tid * 64 + k] = __uint_as_float(d[k]);
        if (mode == 1 && tid == 0 && blockIdx.x == 0) {      // dump smem sanity
            printf("Bs[0..3]=%g %g %g %g  Bs[4]=%g Bs[256]=%g\n", Bs[0], Bs[1], Bs[2], Bs[3], Bs[4], Bs[256]);
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
}

static float tf32_host(float x) {   // round to nearest (ties away), like cvt.rna
    uint32_t u; memcpy(&u, &x, 4);
    u += 0x1000u; u &= 0xFFFFE000u;
    float r; memcpy(&r, &u, 4); return r;
}

int main() {
    const int M = 128, N = 64, K = 64;
    std::vector<float> A(M * K), P(N * K), Bt(8192), out(M * N);
    srand(1);
    for (auto &x : A) x = (float)rand() / RAND_MAX;
    for (int m = 0; m < M; m++) { float mx = 0; for (int k = 0; k < K; k++) mx = fmaxf(mx, A[m * K + k]); for (int k = 0; k < K; k++) A[m * K + k] /= (mx * 1.3f); }
    for (auto &x : P) { float u = (float)rand() / RAND_MAX; x = u * u * u * u; }
    for (int n = 0; n < N; n++)
        for (int k = 0; k < K; k++) {
            float p = P[n * K + k], h = tf32_host(p), l = tf32_host(p - h);
            int o = ((k >> 2) * 64 + n) * 4 + (k & 3);
            Bt[o] = h; Bt[4096 + o] = l;
        }
    float *dA, *dB, *dO; int *dErr;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, Bt.size() * 4); cudaMalloc(&dO, out.size() * 4); cudaMalloc(&dErr, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, Bt.data(), Bt.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dErr, 0, 4);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    auto run = [&](int mode, uint32_t lbo, uint32_t sbo, int nsteps, int order) {
        cudaMemset(dO, 0, out.size() * 4);
        probe_kernel<<<1, 128, TC_SMEM_BYTES>>>(mode, dA, dB, dO, lbo, sbo, nsteps, order, dErr);
        cudaError_t e = cudaDeviceSynchronize();
        int herr = 0; cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
        printf("mode=%d lbo=%u sbo=%u nsteps=%d order=%d : cuda=%s barrier_timeout=%d\n", mode, lbo, sbo, nsteps, order, cudaGetErrorString(e), herr);
        return e == cudaSuccess;
    };
    // (1) round trip
    if (run(0, 0, 0, 0, 0)) {
        int bad = 0; for (int i = 0; i < M * K; i++) if (out[i] != A[i]) bad++;
        printf("  TMEM st/ld round trip mismatches: %d of %d\n", bad, M * K);
    }
    // (3) K=8 and K=64 plain tf32 with both descriptor conventions
    for (int conv = 0; conv < 2; conv++) {
        uint32_t lbo = conv == 0 ? 1024 : 128, sbo = conv == 0 ? 128 : 1024;
        for (int nsteps : {1, 8}) {
            if (!run(1, lbo, sbo, nsteps, 0)) return 1;
            double maxrel = 0; int bad = 0;
            for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) {
                double ref = 0; for (int k = 0; k < nsteps * 8; k++) ref += (double)tf32_host(A[m * K + k]) * (double)Bt[((k >> 2) * 64 + n) * 4 + (k & 3)];
                double rel = fabs(out[m * N + n] - ref) / (fabs(ref) + 1e-30);
                if (rel > 1e-4) bad++; maxrel = fmax(maxrel, rel);
            }
            printf("  conv %s K=%d: bad=%d maxrel=%.3e  out[0][0..2]=%g %g %g out[5][7]=%g\n", conv == 0 ? "LBO=1024(K),SBO=128(N)" : "LBO=128,SBO=1024", nsteps * 8, bad, maxrel, out[0], out[1], out[2], out[5 * 64 + 7]);
        }
    }
    // (4) split accuracy, both accumulation orders, and plain tf32 for scale
    for (int order = 0; order < 3; order++) {
        if (!run(2, 1024, 128, 8, order)) return 1;
        double maxrel = 0, meansigned = 0, meanabs = 0;
        for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) {
            double ref = 0; for (int k = 0; k < K; k++) ref += (double)A[m * K + k] * (double)P[n * K + k];
            double rel = (out[m * N + n] - ref) / ref;
            maxrel = fmax(maxrel, fabs(rel)); meansigned += rel; meanabs += fabs(rel);
        }
        printf("  split order=%d (0 small-first, 1 big-first, 2 plain tf32): max|rel|=%.3e mean|rel|=%.3e MEAN SIGNED rel=%.3e\n", order, maxrel, meanabs / (M * N), meansigned / (M * N));
    }
    return 0;
}
