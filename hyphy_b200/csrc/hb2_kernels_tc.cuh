// hb2_kernels_tc.cuh -- tcgen05 (5th-gen tensor core) pruning path for 33..64-state models (sm_100a only).
//
// The 61-state codon update  parent[s][n] *= sum_k P[n][k] * child[s][k]  is a dense contraction batched over
// patterns:  D[128 patterns][64 parent states] = X[128][64 child states] * P^T.  tcgen05 has no fp64 kind, so the
// contraction runs as an error-compensated split in kind::tf32 with fp32 accumulation in TMEM:
//        X = Xh + Xl,  P = Ph + Pl   (each part exactly representable in tf32, round-to-nearest splits)
//        D = Xl*Ph + Xh*Pl + Xh*Ph   (the dropped Xl*Pl term is ~2^-22 relative)
// All operands are non-negative, so there is no cancellation and the relative error of D is bounded by the
// representation error of the splits (~2^-22) plus fp32 accumulation.  Conditionals are stored in fp32, renormalised
// per (node, pattern) to max in [0.5,1) with an int32 binary exponent, so fp32 range is never an issue.
//
// Anchors.  The tensor core accumulates in fp32 with truncation (measured on B200: -8e-8 mean relative error per
// K=64 contraction), which would bias lnL by ~-1.6e-7*|lnL|.  Conditional vectors are sharply peaked (1-3 codons carry
// almost all the mass), so each thread pulls the entries >= 2^-6 of its row (at most 8, "anchors") out of the tensor
// operand and multiplies them on the CUDA cores with round-to-nearest fp32 FMAs against fp32 rows of P^T while the
// MMAs run; the tensor core only contracts the diffuse remainder, whose truncation bias is scaled down by its share.
//
// Operand staging:
//   A (= X, 128 x 64, K-major) comes from TENSOR MEMORY: each of the 128 threads owns one pattern (one TMEM lane),
//     loads its fp32 row from HBM, splits it in registers and writes Xh / Xl with tcgen05.st (columns 64..191).
//   B (= P, 64 x 64, K-major) comes from SHARED MEMORY in the canonical no-swizzle UMMA layout (8x16-byte core
//     matrices; LBO = 1024 B between K chunks, SBO = 128 B between 8-row groups).  The expm stage writes P already
//     split and tiled that way (pack_tc_kernel), so one 32 KB cp.async.bulk (TMA engine, mbarrier completion) per
//     child stages Ph and Pl.
//   D (128 x 64 fp32) lives in TMEM columns 0..63; it is read back with tcgen05.ld, one row per thread, multiplied
//     into the running product of the parent held in registers.
//
// Replaces (for D in 33..64) reference tree_evaluator.cpp:3875-4008 (_hy_mvp_blocked<61> + _hy_vvmult_sum<61>).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "hb2_kernels_fp64.cuh"

namespace hb2 {

constexpr int TC_TILE_P = 128;                 // patterns per CTA (= UMMA M = TMEM lanes)
constexpr int TC_PB_FLOATS = 2 * 4096;         // per (class, branch): Ph tile + Pl tile in canonical layout
constexpr int TC_PTF_ROW = 68;                 // floats per row of the fp32 P^T table (64 + 4 padding: 272-byte rows)
constexpr int TC_PTF_FLOATS = 64 * TC_PTF_ROW; // per (class, branch)
constexpr int TC_SMEM_BYTES = 2 * 32768 + 1024; // two B stages (one used for now; also caps residency at 2 CTAs/SM) + barriers
constexpr uint32_t TC_TMEM_COLS = 256;         // D: 0..63, Xh: 64..127, Xl: 128..191 (allocations are powers of two)
constexpr int TC_MAX_ANCHORS = 8;              // per pattern and child (list capacity of the per-level and split-row kernels)
constexpr int WALK_FAST_ANCHORS = 4;           // anchors handled by the unrolled fast path of the walk kernel; more go to a loop
constexpr float TC_ANCHOR_THR = 0.015625f;     // 2^-6 (rows are normalised to max in [0.5,1))

struct PruneTcArgs {
    const float *PB;                // [C][B][2][16][64][4] canonical K-major tiles of P (hi, lo)
    const float *PTf;               // [C][B][64][68] fp32 copy of PT, padded rows (leaf column gather, anchors)
    float *cond;                    // [C][I][Sp/128][16 chunks][128 patterns][4] fp32 conditionals (tile-wise K-major chunks)
    int *scal;                      // [C][I][Sp]
    const int *leaf;                // [L][Sp]
    const double *ambig;            // [nAmb][64]
    const double *pi;               // [64]
    double *rootL;                  // [C][Sp]
    int *rootE;                     // [C][Sp]
    TreeDev tree;
    int *err;                       // device error flag (mbarrier timeout)
    int L, I, B, D, Sp, cat0;
    const int *forced;              // forced states of ONE node (see PruneArgs), or unused when forced_node < 0
    int forced_node;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Device-side waits are bounded by TIME (4 s of globaltimer), not by an iteration count: a slow producer (another
// partition or process sharing the GPU, fewer co-resident CTAs than planned) only makes the pass slower; a protocol bug
// still surfaces as an error flag (the host returns an error), never as a hung GPU.
constexpr unsigned long long HB2_WAIT_LIMIT_NS = 4000000000ull;
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int *err_flag) {
    const uint32_t addr = smem_u32(bar);
    unsigned long long t0 = 0;
    for (int it = 0;; it++) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return true;
        if ((it & 1023) == 1023) {
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > HB2_WAIT_LIMIT_NS) break;
        }
    }
    atomicExch(err_flag, 1);
    return false;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc], kind::tf32, M=128, N=64, K=8
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// One lane of a converged warp (the CUTLASS issue idiom).  tcgen05.mma wants its operands in UNIFORM registers: issued
// from a divergent `if (lane == 0)` ptxas wraps every UTCHMMA in an ELECT / R2UR.BROADCAST waterfall loop (~190 cycles per
// instruction, profiles/r01j_mma_issue_timing.txt); issued by an elected lane inside a warp-uniform branch, with operands
// that were broadcast by __shfl_sync, the instructions go out back to back (profiles/r02_mma_issue_uniform.txt).
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred;
}
// round-to-nearest (ties away from zero) to tf32, i.e. what cvt.rna.tf32.f32 computes, but with two full-rate integer
// instructions: the conversion instruction issues at a quarter of the ALU rate and dominated the operand-split phase
__device__ __forceinline__ float tf32_rn(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
#define HB2_TMEM_ST16(taddr, r, o)                                                                                   \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
                 ::"r"(taddr), "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]),   \
                 "r"(r[o + 6]), "r"(r[o + 7]), "r"(r[o + 8]), "r"(r[o + 9]), "r"(r[o + 10]), "r"(r[o + 11]),            \
                 "r"(r[o + 12]), "r"(r[o + 13]), "r"(r[o + 14]), "r"(r[o + 15]) : "memory")
#define HB2_TMEM_LD16(taddr, r, o)                                                                                   \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]),      \
                   "=r"(r[o + 6]), "=r"(r[o + 7]), "=r"(r[o + 8]), "=r"(r[o + 9]), "=r"(r[o + 10]), "=r"(r[o + 11]),    \
                   "=r"(r[o + 12]), "=r"(r[o + 13]), "=r"(r[o + 14]), "=r"(r[o + 15])                                    \
                 : "r"(taddr) : "memory")

// UMMA shared-memory descriptor, SWIZZLE_NONE, K-major: start>>4 | (LBO>>4)<<16 | (SBO>>4)<<32 | version 1 @46
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), K-major both, N=64 (8<<17), M=128 (8<<24)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// ------------------------------------------------------------------------------------------------------------------
// PT (fp64, transposed) -> tensor-path operands: canonical hi/lo tiles of P and an fp32 copy of PT.  One CTA per slot.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_tc_kernel(const double *__restrict__ PT, const int *__restrict__ slots,
                                                       float *__restrict__ PB, float *__restrict__ PTf) {
    if (slots[blockIdx.x] < 0) return;
    const size_t slot = slots[blockIdx.x];
    const double *src = PT + slot * 4096;
    float *pb = PB + slot * TC_PB_FLOATS;
    float *pf = PTf + slot * TC_PTF_FLOATS;
    for (int o = threadIdx.x; o < 4096; o += 256) {
        const int chunk = o >> 8, n = (o >> 2) & 63, kk = chunk * 4 + (o & 3);
        const double p = src[kk * 64 + n];              // P[n][kk] = PT[kk][n]
        const float hi = tf32_rn((float)p);
        const float lo = tf32_rn((float)(p - (double)hi));
        pb[o] = hi;
        pb[4096 + o] = lo;
        pf[(o >> 6) * TC_PTF_ROW + (o & 63)] = (float)src[o];
    }
}

// Explicit-form mixtures (BS-REL, reference tree.cpp:3047-3089): PT[slot_i] = sum_k w[i][k] * comp[i*K + k], components
// added in order k = 0..K-1 (deterministic); 64-state tensor partitions also get the split tiles and the fp32 table.
// One CTA per node; comp = scratch matrices [n*K][Dp*Dp] in the PT layout (transposed).
__global__ void __launch_bounds__(256) mix_reduce_kernel(const double *__restrict__ comp, const double *__restrict__ w,
                                                          const int *__restrict__ slots, int K, int Dp, double *__restrict__ PT,
                                                          float *__restrict__ PB, float *__restrict__ PTf) {
    const size_t slot = slots[blockIdx.x];
    const int dpdp = Dp * Dp;
    double *dst = PT + slot * dpdp;
    const double *src = comp + (size_t)blockIdx.x * K * dpdp;
    const double *wk = w + (size_t)blockIdx.x * K;
    for (int o = threadIdx.x; o < dpdp; o += 256) {
        double s = 0.0;
        for (int k = 0; k < K; k++) s += wk[k] * src[(size_t)k * dpdp + o];
        dst[o] = s;
    }
    if (PB && Dp == 64) {
        __syncthreads();                         // this CTA's own global writes are visible to it after the barrier
        float *pb = PB + slot * TC_PB_FLOATS;
        float *pf = PTf + slot * TC_PTF_FLOATS;
        for (int o = threadIdx.x; o < 4096; o += 256) {
            const int chunk = o >> 8, n = (o >> 2) & 63, kk = chunk * 4 + (o & 3);
            const double p = dst[kk * 64 + n];
            const float hi = tf32_rn((float)p);
            pb[o] = hi;
            pb[4096 + o] = tf32_rn((float)(p - (double)hi));
            pf[(o >> 6) * TC_PTF_ROW + (o & 63)] = (float)dst[o];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Fused pruning update on tcgen05.  grid = (Sp/128, jobs, classes), block = 128 threads (thread t <-> pattern t).
// ------------------------------------------------------------------------------------------------------------------
// Power-of-two renormalisation of a pattern's 64 conditionals: max -> [0.5,1).  `force` = false only rescales when the
// maximum has drifted below 2^-32 (cheap guard between the children of one node); true at the end of a node.
__device__ __forceinline__ void renorm_f32(float (&v)[64], int &ex, bool force = true) {
    float m0 = v[0], m1 = v[1], m2 = v[2], m3 = v[3];          // four independent chains instead of one of length 64
#pragma unroll
    for (int k = 4; k < 64; k += 4) { m0 = fmaxf(m0, v[k]); m1 = fmaxf(m1, v[k + 1]); m2 = fmaxf(m2, v[k + 2]); m3 = fmaxf(m3, v[k + 3]); }
    const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    if (m > 0.f && m < INFINITY && (force || m < 2.3283064e-10f)) {
        const int e = max((int)((__float_as_uint(m) >> 23) & 0xffu) - 126, -125);   // m = f * 2^e, f in [0.5,1) (normal m)
        if (e != 0) {
            const float sc = __uint_as_float((uint32_t)(127 - e) << 23);           // 2^-e, exact
#pragma unroll
            for (int k = 0; k < 64; k++) v[k] *= sc;
            ex += e;
        }
    }
}

__global__ void __launch_bounds__(128, 2) prune64_tc_kernel(PruneTcArgs a, const int *__restrict__ jobs) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float *Bs = reinterpret_cast<float *>(smem);                         // [2][4096] Ph, Pl
    uint64_t *bar_b = reinterpret_cast<uint64_t *>(smem + 2 * 32768);
    uint64_t *bar_mma = bar_b + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_b + 2);
    int *s_ak = reinterpret_cast<int *>(smem + 32768);                   // [TC_MAX_ANCHORS][128] anchor state
    float *s_av = reinterpret_cast<float *>(smem + 32768 + TC_MAX_ANCHORS * 128 * 4);   // [TC_MAX_ANCHORS][128] anchor value
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int par = jobs[blockIdx.y];
    const int cat = a.cat0 + blockIdx.z;
    const size_t Sp = a.Sp;
    const size_t s = (size_t)blockIdx.x * TC_TILE_P + tid;

    if (tid == 0) {
        mbar_init(bar_b, 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);      // this warp's TMEM lane quarter
    const uint64_t bdesc_hi = make_b_desc(__shfl_sync(0xffffffffu, smem_u32(Bs), 0));
    const uint64_t bdesc_lo = make_b_desc(__shfl_sync(0xffffffffu, smem_u32(Bs + 4096), 0));

    float v[64];
#pragma unroll
    for (int k = 0; k < 64; k++) v[k] = 1.f;
    int ex = 0;
    uint32_t phase = 0;
    const int c_begin = a.tree.child_start[par], c_end = a.tree.child_start[par + 1];
    for (int ci = c_begin; ci < c_end; ci++) {
        const int child = a.tree.child_ids[ci];
        const size_t slot = (size_t)cat * a.B + child;
        if (child < a.L) {
            const int code = (child == a.forced_node) ? a.forced[s] : a.leaf[(size_t)child * Sp + s];
            const float *PTf = a.PTf + slot * TC_PTF_FLOATS;
            if (code >= 0) {
                const float4 *row = reinterpret_cast<const float4 *>(PTf + (size_t)code * TC_PTF_ROW);
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const float4 r = __ldg(row + q);
                    v[4 * q] *= r.x; v[4 * q + 1] *= r.y; v[4 * q + 2] *= r.z; v[4 * q + 3] *= r.w;
                }
            } else {
                float acc[64];
#pragma unroll
                for (int k = 0; k < 64; k++) acc[k] = 0.f;
                const double *amb = a.ambig + (size_t)(-code - 1) * 64;
                for (int j = 0; j < a.D; j++) {
                    if (__ldg(amb + j) != 0.0) {
                        const float4 *row = reinterpret_cast<const float4 *>(PTf + (size_t)j * TC_PTF_ROW);
#pragma unroll
                        for (int q = 0; q < 16; q++) {
                            const float4 r = __ldg(row + q);
                            acc[4 * q] += r.x; acc[4 * q + 1] += r.y; acc[4 * q + 2] += r.z; acc[4 * q + 3] += r.w;
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 64; k++) v[k] *= acc[k];
            }
        } else {
            const int cin = child - a.L;
            // (A) stage Ph|Pl of this branch: one 32 KB bulk copy (TMA engine), completion on bar_b
            if (tid == 0) {
                mbar_expect_tx(bar_b, 32768u);
                bulk_g2s(Bs, a.PB + slot * TC_PB_FLOATS, 32768u, bar_b);
            }
            // (B) this thread's pattern row -> anchors out -> split -> TMEM (Xh at cols 64.., Xl at cols 128..)
            int na = 0;
            {
                const float4 *xr = reinterpret_cast<const float4 *>(a.cond) + ((((size_t)cat * a.I + cin) * (Sp / 128) + blockIdx.x) * 16) * 128 + tid;
                uint32_t hi[64], lo[64];
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const float4 x = xr[(size_t)q * 128];
                    float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (xs[u] >= TC_ANCHOR_THR && na < TC_MAX_ANCHORS) {
                            s_ak[na * 128 + tid] = 4 * q + u;
                            s_av[na * 128 + tid] = xs[u];
                            na++;
                            xs[u] = 0.f;
                        }
                        const float h = tf32_rn(xs[u]);
                        hi[4 * q + u] = __float_as_uint(h);
                        lo[4 * q + u] = __float_as_uint(xs[u] - h);
                    }
                }
#pragma unroll
                for (int o = 0; o < 64; o += 16) {
                    HB2_TMEM_ST16(lane_addr + 64 + o, hi, o);
                    HB2_TMEM_ST16(lane_addr + 128 + o, lo, o);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            ex += a.scal[((size_t)cat * a.I + cin) * Sp + s];
            tc_fence_before();
            __syncthreads();                 // A operand complete in TMEM; every thread is done reading the previous D
            if (warp == 0) {
                tc_fence_after();
                mbar_wait(bar_b, phase, a.err);
                if (elect_one()) {
                    // small terms first: Xl*Ph, Xh*Pl, then Xh*Ph (8 K-steps of 8 each)
#pragma unroll
                    for (int kk = 0; kk < 8; kk++)
                        tc_mma_tf32_ts(tmem_base, tmem_base + 128 + kk * 8, bdesc_hi + (uint64_t)(kk * 2 * 1024 >> 4), TC_IDESC, kk > 0);
#pragma unroll
                    for (int kk = 0; kk < 8; kk++)
                        tc_mma_tf32_ts(tmem_base, tmem_base + 64 + kk * 8, bdesc_lo + (uint64_t)(kk * 2 * 1024 >> 4), TC_IDESC, 1u);
#pragma unroll
                    for (int kk = 0; kk < 8; kk++)
                        tc_mma_tf32_ts(tmem_base, tmem_base + 64 + kk * 8, bdesc_hi + (uint64_t)(kk * 2 * 1024 >> 4), TC_IDESC, 1u);
                    tc_commit(bar_mma);
                }
            }
            __syncwarp();
            // anchors on the CUDA cores while the tensor core works: acc[n] = sum_a x[k_a] * P[n][k_a], fp32 RN
            float acc[64];
#pragma unroll
            for (int k = 0; k < 64; k++) acc[k] = 0.f;
            {
                const float *PTf = a.PTf + slot * TC_PTF_FLOATS;
                for (int ai = 0; ai < na; ai++) {
                    const int ka = s_ak[ai * 128 + tid];
                    const float xv = s_av[ai * 128 + tid];
                    const float4 *row = reinterpret_cast<const float4 *>(PTf + (size_t)ka * TC_PTF_ROW);
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        const float4 r = __ldg(row + q);
                        acc[4 * q] = fmaf(xv, r.x, acc[4 * q]); acc[4 * q + 1] = fmaf(xv, r.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(xv, r.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(xv, r.w, acc[4 * q + 3]);
                    }
                }
            }
            mbar_wait(bar_mma, phase, a.err);
            tc_fence_after();
#pragma unroll
            for (int o = 0; o < 64; o += 16) {
                uint32_t d[16];
                HB2_TMEM_LD16(lane_addr + o, d, 0);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int k = 0; k < 16; k++) v[o + k] *= (__uint_as_float(d[k]) + acc[o + k]);
            }
            phase ^= 1u;
        }
        if (ci == c_end - 1 && a.L + par == a.forced_node) {     // pinned internal node: only the forced state survives
            const int f = a.forced[s];
#pragma unroll
            for (int k = 0; k < 64; k++) if (k != f) v[k] = 0.f;
        }
        renorm_f32(v, ex);
    }
    // write the parent's conditionals (already renormalised after the last child) and the root reduction
    {
        float4 *outp = reinterpret_cast<float4 *>(a.cond) + ((((size_t)cat * a.I + par) * (Sp / 128) + blockIdx.x) * 16) * 128 + tid;
#pragma unroll
        for (int q = 0; q < 16; q++) outp[(size_t)q * 128] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        a.scal[((size_t)cat * a.I + par) * Sp + s] = ex;
        if (par == a.I - 1) {
            double r = 0.0;
#pragma unroll
            for (int k = 0; k < 64; k++) r = fma((double)v[k], a.pi[k], r);
            a.rootL[(size_t)cat * Sp + s] = r;
            a.rootE[(size_t)cat * Sp + s] = ex;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Persistent "walk" kernel: the whole pruning pass of one evaluation in ONE launch.
//
// Patterns are independent, so all dependencies are tile-local: tile t of node p needs tile t of p's children only.
// CTA (c, t, r) owns rate class c, pattern tile t and "lane" r (one of K in-order queues of JOBS the host planner
// fills, hb2_engine.cu run_walk: a job is an internal node or the side product of a node -- all its children but the
// deepest -- kept in node slot I + n); it executes its lane's STEPS (one per child of every job) back to back.
//   * A child computed by another lane of the same (c, t) is awaited on the DATA ITSELF ("tags"): every job has a
//     generation bit that the host flips each time the job is re-pruned; the producer writes it into the sign bit of
//     every conditional (they are non-negative) and into bit 0 of the exponent word, and the consumer thread -- which
//     needs exactly the 17 words the producer thread of the same pattern wrote -- re-reads them until all carry the new
//     bit.  No flag, no fence, no barrier on either side; a tile is consumed one L2 round trip after it was written.
//   * A child computed by THIS CTA in the previous job (the spine of deep trees) is taken straight from registers.
//   * All K lanes of a (c,t) must be co-resident (the host guarantees grid <= resident capacity); with K == 1 there are
//     no cross-CTA waits at all and a CTA may loop over several (c,t) pairs.
//
// The pass is bound by (tree depth x per-step latency), so every operand of step i+1 is staged while step i runs:
//   * per step, the branch's fp32 P^T table (64 rows x 256 B, padded to 272 B rows) and -- for a contraction -- its
//     Ph|Pl UMMA tiles are bulk-copied (TMA engine, mbarrier tx completion) into a 2-stage shared-memory ring by thread 0
//     right after the step-begin barrier; leaf column gathers and anchor rows are then shared-memory reads;
//   * the 32 KB conditional block of the next contraction is prefetched into L2 (cp.async.bulk.prefetch.L2) and the next
//     leaf's state codes (or the awaited child's generation bit) into a register;
//   * conditionals are stored tile-wise as [16 chunks][128 patterns][4 floats] (the K-major UMMA core-matrix order), so
//     thread t's 16-byte accesses are perfectly coalesced for both the producer and the consumer of a tile.
// Step encoding: x = child id | WALK_WAIT | WALK_CHAIN | WALK_MUL ; y = job's node slot | STEP_FIRST | STEP_LAST.
// ------------------------------------------------------------------------------------------------------------------
constexpr int WALK_WAIT = 1 << 30;
constexpr int WALK_CHAIN = 1 << 29;
constexpr int WALK_MUL = 1 << 28;         // child is a side-product node (index >= I): multiply it in, no contraction
constexpr int WALK_ID_MASK = (1 << 27) - 1;
constexpr int STEP_FIRST = 1 << 28;
constexpr int STEP_LAST = 1 << 29;
constexpr int WALK_PT_ROW = TC_PTF_ROW;        // same padded rows in shared memory: one flat bulk copy stages the table
constexpr int WALK_STAGE_FLOATS = 64 * WALK_PT_ROW + 8192;     // P^T table + Ph|Pl tiles
constexpr int WALK_SMEM_BYTES = 2 * WALK_STAGE_FLOATS * 4 + TC_MAX_ANCHORS * 128 * 8 + 64;

struct WalkArgs {
    PruneTcArgs a;
    const int *lane_start;      // [K+1] offsets into steps
    const int2 *steps;
    const int *gen;             // [C][NI] generation bit of every job's conditionals AFTER this pass (see "tags" above)
    int K, T, ncls, nslots;
    int NI;                     // node slots per class in cond/scal/gen: I real nodes + I side products
    long long *trace;           // nullable debug buffer: 12 clock64 stamps per step of CTA `trace_cta` (HB2_WALK_TRACE)
    long long *trace_cta_times; // nullable: per CTA {smid, clock64 at entry, clock64 at exit, globaltimer at entry}
    int trace_cta;
    float anchor_thr;           // entries >= this (rows are normalised to max in [0.5,1)) bypass the tensor core; default 2^-6
};

// Tagged hand-over between CTAs: relaxed gpu-scope accesses, every 32-bit word carries its own validity bit, so no
// ordering between the words (and no fence) is needed.
__device__ __forceinline__ uint4 ld_relaxed_u4(const void *p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const void *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// float4 index of chunk q of pattern t in the conditional block of (cat, node, tile)
__device__ __forceinline__ size_t cond_f4(int cat, int node, int tile, int q, int t, int I, int T) {
    return ((((size_t)cat * I + node) * T + tile) * 16 + q) * 128 + t;
}

__global__ void __launch_bounds__(128, 2) prune64_tc_walk_kernel(WalkArgs w) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const PruneTcArgs &a = w.a;
    float *stage_base = reinterpret_cast<float *>(smem);                           // 2 x [P^T table | Ph | Pl]
    int *s_ak = reinterpret_cast<int *>(smem + 2 * WALK_STAGE_FLOATS * 4);          // [TC_MAX_ANCHORS][128]
    float *s_av = reinterpret_cast<float *>(s_ak + TC_MAX_ANCHORS * 128);
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(s_av + TC_MAX_ANCHORS * 128); // [2]
    uint64_t *bar_mma = bar_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_mma + 1);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const size_t Sp = a.Sp;
    float4 *cond4 = reinterpret_cast<float4 *>(a.cond);

    if (tid == 0) {
        mbar_init(bar_full, 1);
        mbar_init(bar_full + 1, 1);
        mbar_init(bar_mma, 1);               // one tcgen05.commit per contraction (elected lane of warp 0)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t n_step = 0, n_mma = 0;          // running counters: select ring stage / barrier parities
    bool bailed = false;                     // a dependency wait timed out: stop waiting, the host reports the error

    const int r = blockIdx.x % w.K;
    const int i_begin = w.lane_start[r], i_end = w.lane_start[r + 1];
    if (w.trace_cta_times && tid == 0) {
        uint32_t smid; unsigned long long gt;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        long long *q = w.trace_cta_times + (size_t)blockIdx.x * 4;
        q[0] = smid; q[1] = clock64(); q[3] = (long long)gt;
    }

    // one lane of warp 3: stage the operands of one step into ring slot (m & 1): two flat bulk copies + one L2 prefetch
    auto stage_step = [&](int cat, int tile, int2 st, uint32_t m) {
        const int child = st.x & WALK_ID_MASK;
        const bool internal = child >= a.L;
        float *dst = stage_base + (m & 1u) * WALK_STAGE_FLOATS;
        uint64_t *bar = bar_full + (m & 1u);
        if (st.x & WALK_MUL) {               // side product: no matrix involved; the slot's phase still has to complete
            mbar_expect_tx(bar, 0u);
            prefetch_l2_bulk(cond4 + cond_f4(cat, child - a.L, tile, 0, 0, w.NI, w.T), 32768u);
            return;
        }
        const size_t slot = (size_t)cat * a.B + child;
        mbar_expect_tx(bar, (uint32_t)(TC_PTF_FLOATS * 4) + (internal ? 32768u : 0u));
        bulk_g2s(dst, a.PTf + slot * TC_PTF_FLOATS, (uint32_t)(TC_PTF_FLOATS * 4), bar);
        if (internal) {
            bulk_g2s(dst + 64 * WALK_PT_ROW, a.PB + slot * TC_PB_FLOATS, 32768u, bar);
            if (!(st.x & WALK_CHAIN))      // pull the child's conditional block towards L2 (it may still be in DRAM)
                prefetch_l2_bulk(cond4 + cond_f4(cat, child - a.L, tile, 0, 0, w.NI, w.T), 32768u);
        }
    };

    // This thread's 17 words of another node's tile: 16 x 4 conditionals + the exponent word.  With `await` they are
    // re-read until every word carries generation bit `want` (sign bit / bit 0), i.e. until the producer lane of THIS
    // launch has written them; otherwise the tile is resident and its bits are whatever pass produced it.
    auto load_tile_row = [&](const float4 *xrow, const int *scp, bool await, uint32_t want, uint4 (&x4)[16], uint32_t &sc) {
#pragma unroll
        for (int q = 0; q < 16; q++) x4[q] = __ldcg(reinterpret_cast<const uint4 *>(xrow + (size_t)q * 128));
        sc = (uint32_t)__ldcg(scp);
        if (await) {
            unsigned long long t0 = 0;
            for (int it = 0; !bailed; it++) {
                uint32_t bad = (sc << 31) ^ want;
#pragma unroll
                for (int q = 0; q < 16; q++) bad |= (x4[q].x ^ want) | (x4[q].y ^ want) | (x4[q].z ^ want) | (x4[q].w ^ want);
                if (!(bad >> 31)) break;
                if ((it & 255) == 255) {                                       // never hang the GPU: time-bounded
                    const unsigned long long now = globaltimer_ns();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > HB2_WAIT_LIMIT_NS) { atomicExch(a.err, 2); bailed = true; }
                }
#pragma unroll
                for (int q = 0; q < 16; q++) x4[q] = ld_relaxed_u4(xrow + (size_t)q * 128);
                sc = ld_relaxed_u32(scp);
            }
        }
    };

    for (int ct = blockIdx.x / w.K; ct < w.ncls * w.T; ct += w.nslots) {
        const int cat = a.cat0 + ct / w.T;
        const int tile = ct % w.T;
        const size_t s = (size_t)tile * TC_TILE_P + tid;
        float v[64];
        int ex = 0;
        if (i_begin == i_end) continue;
        int2 st = __ldg(w.steps + i_begin);
        int2 nx = (i_begin + 1 < i_end) ? __ldg(w.steps + i_begin + 1) : make_int2(0, 0);
        __syncthreads();                      // previous (class, tile): every read of the ring is complete
        if (warp == 3 && elect_one()) stage_step(cat, tile, st, n_step);     // warp 3 stages: warp 0 issues the MMAs
        // fetched one step ahead: a leaf's state code, or (contractions) the generation bit the child's words must carry
        auto step_aux = [&](int2 q) -> int {
            const int ch = q.x & WALK_ID_MASK;
            if (ch < a.L) return (ch == a.forced_node) ? __ldg(a.forced + s) : __ldg(a.leaf + (size_t)ch * Sp + s);
            return (q.x & WALK_WAIT) ? __ldg(w.gen + (size_t)cat * w.NI + (ch - a.L)) : 0;
        };
        int next_code = step_aux(st);
        for (int i = i_begin; i < i_end; i++) {
            const int enc = st.x;
            const int child = enc & WALK_ID_MASK;
            const int par = st.y & WALK_ID_MASK;
            const int flags = st.y;
            const int code = next_code;
            const uint32_t tagbit = (flags & STEP_LAST) ? ((uint32_t)__ldg(w.gen + (size_t)cat * w.NI + par) << 31) : 0u;   // used at the end
            const bool has_next = (i + 1 < i_end);
            const int2 nx2 = (i + 2 < i_end) ? __ldg(w.steps + i + 2) : make_int2(0, 0);   // descriptors run two steps ahead
            const bool tr = w.trace && tid == 0 && (int)blockIdx.x == w.trace_cta;
            long long *trp = tr ? w.trace + (size_t)(i - i_begin) * 12 : nullptr;
            if (tr) { trp[0] = ((long long)st.x << 32) | (unsigned)st.y; trp[1] = clock64(); }
            __syncthreads();                  // (1) everyone is done with step i-1: ring slot (n_step+1)&1 is free
            if (has_next) {
                if (warp == 3 && elect_one()) stage_step(cat, tile, nx, n_step + 1);
                next_code = step_aux(nx);
            }
            if ((flags & STEP_FIRST) && !(enc & WALK_CHAIN)) {
#pragma unroll
                for (int k = 0; k < 64; k++) v[k] = 1.f;
                ex = 0;
            }
            const float *tab = stage_base + (n_step & 1u) * WALK_STAGE_FLOATS;     // P^T table of this branch
            if (tr) trp[2] = clock64();
            if (child < a.L) {
                mbar_wait(bar_full + (n_step & 1u), (n_step >> 1) & 1u, a.err);
                if (tr) trp[3] = clock64();
                if (code >= 0) {
                    const float4 *row = reinterpret_cast<const float4 *>(tab + code * WALK_PT_ROW);
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        const float4 rr = row[q];
                        v[4 * q] *= rr.x; v[4 * q + 1] *= rr.y; v[4 * q + 2] *= rr.z; v[4 * q + 3] *= rr.w;
                    }
                } else {
                    float acc[64];
#pragma unroll
                    for (int k = 0; k < 64; k++) acc[k] = 0.f;
                    const double *amb = a.ambig + (size_t)(-code - 1) * 64;
                    for (int jj = 0; jj < a.D; jj++) {
                        if (__ldg(amb + jj) != 0.0) {
                            const float4 *row = reinterpret_cast<const float4 *>(tab + jj * WALK_PT_ROW);
#pragma unroll
                            for (int q = 0; q < 16; q++) {
                                const float4 rr = row[q];
                                acc[4 * q] += rr.x; acc[4 * q + 1] += rr.y; acc[4 * q + 2] += rr.z; acc[4 * q + 3] += rr.w;
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 64; k++) v[k] *= acc[k];
                }
            } else if (enc & WALK_MUL) {
                // side product of this node (its other children, multiplied up by another lane or earlier in this one):
                // no matrix, just the element-wise product and the exponent
                const int cin = child - a.L;
                const bool await = (enc & WALK_WAIT) != 0;
                uint4 x4[16];
                uint32_t sc;
                load_tile_row(cond4 + cond_f4(cat, cin, tile, 0, tid, w.NI, w.T), a.scal + ((size_t)cat * w.NI + cin) * Sp + s, await,
                              await ? ((uint32_t)code << 31) : 0u, x4, sc);
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    v[4 * q] *= __uint_as_float(x4[q].x & 0x7fffffffu); v[4 * q + 1] *= __uint_as_float(x4[q].y & 0x7fffffffu);
                    v[4 * q + 2] *= __uint_as_float(x4[q].z & 0x7fffffffu); v[4 * q + 3] *= __uint_as_float(x4[q].w & 0x7fffffffu);
                }
                ex += (int)sc >> 1;
                if (tr) trp[3] = clock64();
            } else {
                const int cin = child - a.L;
                // Anchors without a serial dependency: one independent compare per element builds a 64-bit mask and
                // zeroes the element in the tensor operand; the (few) set bits are then enumerated with ffs and their
                // values re-read from the child's conditional block in L2 (the row this thread itself loaded or, on a
                // chain, stored a moment ago), eight loads in flight.
                uint32_t am0 = 0, am1 = 0, am2 = 0, am3 = 0;
                const float4 *xrow = cond4 + cond_f4(cat, cin, tile, 0, tid, w.NI, w.T);
                {
                    uint32_t hi[64], lo[64];
                    if (enc & WALK_CHAIN) {
#pragma unroll
                        for (int k = 0; k < 64; k++) {
                            float x = v[k];
                            v[k] = 1.f;
                            const bool big = x >= w.anchor_thr;
                            if (big) { if (k < 16) am0 |= 1u << k; else if (k < 32) am1 |= 1u << (k - 16); else if (k < 48) am2 |= 1u << (k - 32); else am3 |= 1u << (k - 48); }
                            x = big ? 0.f : x;
                            const float h = tf32_rn(x);
                            hi[k] = __float_as_uint(h);
                            lo[k] = __float_as_uint(x - h);
                        }
                    } else {
                        // A child produced by another lane during THIS launch is awaited on the data itself: every word
                        // must carry the generation bit of this pass (sign bit of a conditional, bit 0 of the exponent
                        // word).  Other children are resident; their bits are whatever pass produced them and are ignored.
                        const bool await = (enc & WALK_WAIT) != 0;
                        const uint32_t want = await ? ((uint32_t)code << 31) : 0u;
                        const int *scp = a.scal + ((size_t)cat * w.NI + cin) * Sp + s;
                        uint4 x4[16];
                        uint32_t sc;
                        load_tile_row(xrow, scp, await, want, x4, sc);
#pragma unroll
                        for (int q = 0; q < 16; q++) {
                            const uint32_t xs[4] = {x4[q].x, x4[q].y, x4[q].z, x4[q].w};
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int kq = 4 * q + u;
                                const float xv = __uint_as_float(xs[u] & 0x7fffffffu);
                                const bool big = xv >= w.anchor_thr;
                                if (big) { if (kq < 16) am0 |= 1u << kq; else if (kq < 32) am1 |= 1u << (kq - 16); else if (kq < 48) am2 |= 1u << (kq - 32); else am3 |= 1u << (kq - 48); }
                                const float x = big ? 0.f : xv;
                                const float h = tf32_rn(x);
                                hi[kq] = __float_as_uint(h);
                                lo[kq] = __float_as_uint(x - h);
                            }
                        }
                        ex += (int)sc >> 1;
                    }
                    if (tr) trp[8] = clock64();
#pragma unroll
                    for (int o = 0; o < 64; o += 16) {
                        HB2_TMEM_ST16(lane_addr + 64 + o, hi, o);
                        HB2_TMEM_ST16(lane_addr + 128 + o, lo, o);
                    }
                    if (tr) trp[9] = clock64();
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    if (tr) trp[10] = clock64();
                }
                unsigned long long amask = (unsigned long long)(am0 | (am1 << 16)) | ((unsigned long long)(am2 | (am3 << 16)) << 32);
                int ak[WALK_FAST_ANCHORS];
                float av[WALK_FAST_ANCHORS];
                const float *xrow_f = reinterpret_cast<const float *>(xrow);
#pragma unroll
                for (int ai = 0; ai < WALK_FAST_ANCHORS; ai++) {
                    ak[ai] = -1; av[ai] = 0.f;
                    if (amask) {
                        const int kk = __ffsll((long long)amask) - 1;
                        amask &= amask - 1;
                        ak[ai] = kk;
                        av[ai] = __ldcg(xrow_f + ((size_t)(kk >> 2) * 128) * 4 + (kk & 3));   // sign bit = tag: stripped at the use, so the load stays in flight
                    }
                }
                if (tr) trp[3] = clock64();
                tc_fence_before();
                __syncthreads();             // (2) A operand complete in TMEM; every thread is done reading the previous D
                if (tr) trp[4] = clock64();
                if (warp == 0) {
                    // 24 tcgen05.mma (small terms first: Xl*Ph, Xh*Pl, Xh*Ph; 8 K-steps of 8 each) into ONE accumulator
                    // (TMEM columns 0..63), issued back to back by the elected lane of warp 0 with uniform operands: the
                    // chain is bound by the tensor pipe (32 cycles per M128 N64 K8 instruction), not by the issue.
                    tc_fence_after();
                    mbar_wait(bar_full + (n_step & 1u), (n_step >> 1) & 1u, a.err);
                    if (tr) trp[5] = clock64();
                    const uint32_t baddr = __shfl_sync(0xffffffffu, smem_u32(tab + 64 * WALK_PT_ROW), 0);
                    const uint64_t bdesc_hi = make_b_desc(baddr);
                    const uint64_t bdesc_lo = make_b_desc(baddr + 4096 * 4);
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 8; kk++) tc_mma_tf32_ts(tmem_base, tmem_base + 128 + kk * 8, bdesc_hi + (uint64_t)(kk * 2 * 1024 >> 4), TC_IDESC, kk > 0);
#pragma unroll
                        for (int kk = 0; kk < 8; kk++) tc_mma_tf32_ts(tmem_base, tmem_base + 64 + kk * 8, bdesc_lo + (uint64_t)(kk * 2 * 1024 >> 4), TC_IDESC, 1u);
#pragma unroll
                        for (int kk = 0; kk < 8; kk++) tc_mma_tf32_ts(tmem_base, tmem_base + 64 + kk * 8, bdesc_hi + (uint64_t)(kk * 2 * 1024 >> 4), TC_IDESC, 1u);
                        tc_commit(bar_mma);
                    }
                }
                __syncwarp();
                // anchors on the CUDA cores while the tensor core works (rows of the P^T table in shared memory)
                float acc[64];
                mbar_wait(bar_full + (n_step & 1u), (n_step >> 1) & 1u, a.err);
                {   // first anchor (nearly every row has one: its maximum is >= 0.5): a product, no zero-fill + FMA
                    const float xv = ak[0] >= 0 ? fabsf(av[0]) : 0.f;
                    const float4 *row = reinterpret_cast<const float4 *>(tab + max(ak[0], 0) * WALK_PT_ROW);
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        const float4 rr = row[q];
                        acc[4 * q] = xv * rr.x; acc[4 * q + 1] = xv * rr.y; acc[4 * q + 2] = xv * rr.z; acc[4 * q + 3] = xv * rr.w;
                    }
                }
#pragma unroll
                for (int ai = 1; ai < WALK_FAST_ANCHORS; ai++) {
                    if (ak[ai] >= 0) {
                        const float xv = fabsf(av[ai]);
                        const float4 *row = reinterpret_cast<const float4 *>(tab + ak[ai] * WALK_PT_ROW);
#pragma unroll
                        for (int q = 0; q < 16; q++) {
                            const float4 rr = row[q];
                            acc[4 * q] = fmaf(xv, rr.x, acc[4 * q]); acc[4 * q + 1] = fmaf(xv, rr.y, acc[4 * q + 1]);
                            acc[4 * q + 2] = fmaf(xv, rr.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(xv, rr.w, acc[4 * q + 3]);
                        }
                    }
                }
                while (amask) {                  // more than TC_MAX_ANCHORS entries above the threshold (diffuse vectors): rare
                    const int kk = __ffsll((long long)amask) - 1;
                    amask &= amask - 1;
                    const float xv = fabsf(__ldcg(xrow_f + ((size_t)(kk >> 2) * 128) * 4 + (kk & 3)));
                    const float4 *row = reinterpret_cast<const float4 *>(tab + kk * WALK_PT_ROW);
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        const float4 rr = row[q];
                        acc[4 * q] = fmaf(xv, rr.x, acc[4 * q]); acc[4 * q + 1] = fmaf(xv, rr.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(xv, rr.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(xv, rr.w, acc[4 * q + 3]);
                    }
                }
                mbar_wait(bar_mma, n_mma & 1u, a.err);
                if (tr) trp[6] = clock64();
                tc_fence_after();
#pragma unroll
                for (int o = 0; o < 64; o += 16) {
                    uint32_t d[16];
                    HB2_TMEM_LD16(lane_addr + o, d, 0);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int k = 0; k < 16; k++) v[o + k] *= (__uint_as_float(d[k]) + acc[o + k]);
                }
                n_mma++;
            }
            if ((flags & STEP_LAST) && a.L + par == a.forced_node) {   // pinned internal node (never a side-product slot)
                const int f = __ldg(a.forced + s);
#pragma unroll
                for (int k = 0; k < 64; k++) if (k != f) v[k] = 0.f;
            }
            renorm_f32(v, ex, (flags & STEP_LAST) != 0);
            n_step++;
            if (flags & STEP_LAST) {
                // this tile of the parent: conditionals and exponent, every word tagged with the node's generation bit
                // (consumers in other lanes poll the data itself: no flag, no fence, no barrier); root reduction
                float4 *outp = cond4 + cond_f4(cat, par, tile, 0, tid, w.NI, w.T);
#pragma unroll
                for (int q = 0; q < 16; q++)
                    __stcg(reinterpret_cast<uint4 *>(outp + (size_t)q * 128),
                           make_uint4(__float_as_uint(v[4 * q]) | tagbit, __float_as_uint(v[4 * q + 1]) | tagbit,
                                      __float_as_uint(v[4 * q + 2]) | tagbit, __float_as_uint(v[4 * q + 3]) | tagbit));
                __stcg(a.scal + ((size_t)cat * w.NI + par) * Sp + s, (int)(((uint32_t)ex << 1) | (tagbit >> 31)));
                if (par == a.I - 1) {
                    double rr = 0.0;
#pragma unroll
                    for (int k = 0; k < 64; k++) rr = fma((double)v[k], a.pi[k], rr);
                    a.rootL[(size_t)cat * Sp + s] = rr;
                    a.rootE[(size_t)cat * Sp + s] = ex;
                }
            }
            if (tr) trp[7] = clock64();
            st = nx;
            nx = nx2;
        }
    }
    if (w.trace_cta_times && tid == 0) w.trace_cta_times[(size_t)blockIdx.x * 4 + 2] = clock64();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}


// ------------------------------------------------------------------------------------------------------------------
// prune64_tc_walk2_kernel: the same pass with TWO threads per pattern (256-thread CTA).
//
// The 128-thread kernel above is bound by each warp's own instruction stream (~830 instructions per contraction step at
// ~6 cycles each, two warps per scheduler: profiles/r2e_prune64_tc_walk_kernel.json) -- every thread carries all 64 states
// of its pattern through the split, the anchor products and the read-back.  Here warp w (0..7) serves TMEM lane quarter
// w % 4 (hardware rule for tcgen05.ld/st) and state half h = w / 4: thread (pattern t, half h) owns states 32h..32h+31 of
// pattern t everywhere (conditional chunks 8h..8h+7, TMEM operand columns 32h.., accumulator columns 32h..), so the serial
// stream per step halves while the tile, the MMAs, the staging ring and the plan stay exactly the same.
//   * anchors: each thread publishes the anchor mask of ITS half in shared memory before barrier (2); afterwards both threads
//     of a pattern enumerate the combined 64-bit mask (values re-read from the child's block in L2, as in the first kernel)
//     and accumulate their own 32 parent states;
//   * exponents: inside a job the two halves rescale independently (own guard, own exponent); when the job is complete
//     (STEP_LAST) the halves exchange (maximum, exponent) through shared memory and bring the node to ONE power-of-two scale
//     with max in [0.5,1), which is what is stored (tagged) and handed over in registers on a chain;
//   * the root dot product is combined the same way.
// Everything observable (layouts, tags, plan, results' meaning) is identical to prune64_tc_walk_kernel; the summation order
// inside a row differs, so results agree to rounding, not bitwise (both are deterministic).
// ------------------------------------------------------------------------------------------------------------------
constexpr int WALK2_XBYTES = 2 * 128 * 4 /* anchor masks */ + 2 * 128 * 4 /* half maxima */ + 2 * 128 * 4 /* half exponents */ +
                             2 * 128 * 8 /* root partials */;
constexpr int WALK2_SMEM_BYTES = 2 * WALK_STAGE_FLOATS * 4 + WALK2_XBYTES + 64;

__device__ __forceinline__ void renorm_half(float (&v)[32], int &ex) {     // guard inside a job: own half only, rarely taken
    float m0 = v[0], m1 = v[1], m2 = v[2], m3 = v[3];
#pragma unroll
    for (int k = 4; k < 32; k += 4) { m0 = fmaxf(m0, v[k]); m1 = fmaxf(m1, v[k + 1]); m2 = fmaxf(m2, v[k + 2]); m3 = fmaxf(m3, v[k + 3]); }
    const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    if (m > 0.f && m < 2.3283064e-10f) {
        const int e = max((int)((__float_as_uint(m) >> 23) & 0xffu) - 126, -125);
        const float sc = __uint_as_float((uint32_t)(127 - e) << 23);
#pragma unroll
        for (int k = 0; k < 32; k++) v[k] *= sc;
        ex += e;
    }
}

__global__ void __launch_bounds__(256, 2) prune64_tc_walk2_kernel(WalkArgs w) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const PruneTcArgs &a = w.a;
    float *stage_base = reinterpret_cast<float *>(smem);                           // 2 x [P^T table | Ph | Pl]
    uint32_t *s_am = reinterpret_cast<uint32_t *>(smem + 2 * WALK_STAGE_FLOATS * 4);   // [2][128] anchor masks of the two halves
    float *s_mx = reinterpret_cast<float *>(s_am + 2 * 128);                        // [2][128]
    int *s_ex = reinterpret_cast<int *>(s_mx + 2 * 128);                            // [2][128]
    double *s_rt = reinterpret_cast<double *>(s_ex + 2 * 128);                      // [2][128]
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(s_rt + 2 * 128);              // [2]
    uint64_t *bar_mma = bar_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_mma + 1);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lq = warp & 3, h = warp >> 2, t = 32 * lq + (tid & 31);              // TMEM lane quarter, state half, pattern in tile
    const size_t Sp = a.Sp;
    float4 *cond4 = reinterpret_cast<float4 *>(a.cond);

    if (tid == 0) {
        mbar_init(bar_full, 1);
        mbar_init(bar_full + 1, 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(lq * 32) << 16);
    uint32_t n_step = 0, n_mma = 0;
    bool bailed = false;

    const int r = blockIdx.x % w.K;
    const int i_begin = w.lane_start[r], i_end = w.lane_start[r + 1];
    if (w.trace_cta_times && tid == 0) {
        uint32_t smid; unsigned long long gt;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        long long *q = w.trace_cta_times + (size_t)blockIdx.x * 4;
        q[0] = smid; q[1] = clock64(); q[3] = (long long)gt;
    }

    auto stage_step = [&](int cat, int tile, int2 st, uint32_t m) {
        const int child = st.x & WALK_ID_MASK;
        const bool internal = child >= a.L;
        float *dst = stage_base + (m & 1u) * WALK_STAGE_FLOATS;
        uint64_t *bar = bar_full + (m & 1u);
        if (st.x & WALK_MUL) {
            mbar_expect_tx(bar, 0u);
            prefetch_l2_bulk(cond4 + cond_f4(cat, child - a.L, tile, 0, 0, w.NI, w.T), 32768u);
            return;
        }
        const size_t slot = (size_t)cat * a.B + child;
        mbar_expect_tx(bar, (uint32_t)(TC_PTF_FLOATS * 4) + (internal ? 32768u : 0u));
        bulk_g2s(dst, a.PTf + slot * TC_PTF_FLOATS, (uint32_t)(TC_PTF_FLOATS * 4), bar);
        if (internal) {
            bulk_g2s(dst + 64 * WALK_PT_ROW, a.PB + slot * TC_PB_FLOATS, 32768u, bar);
            if (!(st.x & WALK_CHAIN)) prefetch_l2_bulk(cond4 + cond_f4(cat, child - a.L, tile, 0, 0, w.NI, w.T), 32768u);
        }
    };

    // this thread's 9 words of another node's tile: its 8 chunks (32 states) + the exponent word; `await` as in the first kernel
    auto load_half_row = [&](const float4 *xrow, const int *scp, bool await, uint32_t want, uint4 (&x4)[8], uint32_t &sc) {
#pragma unroll
        for (int q = 0; q < 8; q++) x4[q] = __ldcg(reinterpret_cast<const uint4 *>(xrow + (size_t)q * 128));
        sc = (uint32_t)__ldcg(scp);
        if (await) {
            unsigned long long t0 = 0;
            for (int it = 0; !bailed; it++) {
                uint32_t bad = (sc << 31) ^ want;
#pragma unroll
                for (int q = 0; q < 8; q++) bad |= (x4[q].x ^ want) | (x4[q].y ^ want) | (x4[q].z ^ want) | (x4[q].w ^ want);
                if (!(bad >> 31)) break;
                if ((it & 255) == 255) {
                    const unsigned long long now = globaltimer_ns();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > HB2_WAIT_LIMIT_NS) { atomicExch(a.err, 2); bailed = true; }
                }
#pragma unroll
                for (int q = 0; q < 8; q++) x4[q] = ld_relaxed_u4(xrow + (size_t)q * 128);
                sc = ld_relaxed_u32(scp);
            }
        }
    };

    for (int ct = blockIdx.x / w.K; ct < w.ncls * w.T; ct += w.nslots) {
        const int cat = a.cat0 + ct / w.T;
        const int tile = ct % w.T;
        const size_t s = (size_t)tile * TC_TILE_P + t;
        float v[32];
        int ex = 0;
        if (i_begin == i_end) continue;
        int2 st = __ldg(w.steps + i_begin);
        int2 nx = (i_begin + 1 < i_end) ? __ldg(w.steps + i_begin + 1) : make_int2(0, 0);
        __syncthreads();
        if (warp == 7 && elect_one()) stage_step(cat, tile, st, n_step);
        auto step_aux = [&](int2 q) -> int {
            const int ch = q.x & WALK_ID_MASK;
            if (ch < a.L) return (ch == a.forced_node) ? __ldg(a.forced + s) : __ldg(a.leaf + (size_t)ch * Sp + s);
            return (q.x & WALK_WAIT) ? __ldg(w.gen + (size_t)cat * w.NI + (ch - a.L)) : 0;
        };
        int next_code = step_aux(st);
        for (int i = i_begin; i < i_end; i++) {
            const int enc = st.x;
            const int child = enc & WALK_ID_MASK;
            const int par = st.y & WALK_ID_MASK;
            const int flags = st.y;
            const int code = next_code;
            const uint32_t tagbit = (flags & STEP_LAST) ? ((uint32_t)__ldg(w.gen + (size_t)cat * w.NI + par) << 31) : 0u;
            const bool has_next = (i + 1 < i_end);
            const int2 nx2 = (i + 2 < i_end) ? __ldg(w.steps + i + 2) : make_int2(0, 0);
            const bool tr = w.trace && tid == 0 && (int)blockIdx.x == w.trace_cta;
            long long *trp = tr ? w.trace + (size_t)(i - i_begin) * 12 : nullptr;
            if (tr) { trp[0] = ((long long)st.x << 32) | (unsigned)st.y; trp[1] = clock64(); }
            __syncthreads();                  // (1) everyone is done with step i-1: ring slot (n_step+1)&1 and the exchange arrays are free
            if (has_next) {
                if (warp == 7 && elect_one()) stage_step(cat, tile, nx, n_step + 1);
                next_code = step_aux(nx);
            }
            if ((flags & STEP_FIRST) && !(enc & WALK_CHAIN)) {
#pragma unroll
                for (int k = 0; k < 32; k++) v[k] = 1.f;
                ex = 0;
            }
            const float *tab = stage_base + (n_step & 1u) * WALK_STAGE_FLOATS;
            if (tr) trp[2] = clock64();
            if (child < a.L) {
                mbar_wait(bar_full + (n_step & 1u), (n_step >> 1) & 1u, a.err);
                if (tr) trp[3] = clock64();
                if (code >= 0) {
                    const float4 *row = reinterpret_cast<const float4 *>(tab + code * WALK_PT_ROW + 32 * h);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float4 rr = row[q];
                        v[4 * q] *= rr.x; v[4 * q + 1] *= rr.y; v[4 * q + 2] *= rr.z; v[4 * q + 3] *= rr.w;
                    }
                } else {
                    float acc[32];
#pragma unroll
                    for (int k = 0; k < 32; k++) acc[k] = 0.f;
                    const double *amb = a.ambig + (size_t)(-code - 1) * 64;
                    for (int jj = 0; jj < a.D; jj++) {
                        if (__ldg(amb + jj) != 0.0) {
                            const float4 *row = reinterpret_cast<const float4 *>(tab + jj * WALK_PT_ROW + 32 * h);
#pragma unroll
                            for (int q = 0; q < 8; q++) {
                                const float4 rr = row[q];
                                acc[4 * q] += rr.x; acc[4 * q + 1] += rr.y; acc[4 * q + 2] += rr.z; acc[4 * q + 3] += rr.w;
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 32; k++) v[k] *= acc[k];
                }
            } else if (enc & WALK_MUL) {
                const int cin = child - a.L;
                const bool await = (enc & WALK_WAIT) != 0;
                uint4 x4[8];
                uint32_t sc;
                load_half_row(cond4 + cond_f4(cat, cin, tile, 8 * h, t, w.NI, w.T), a.scal + ((size_t)cat * w.NI + cin) * Sp + s, await,
                              await ? ((uint32_t)code << 31) : 0u, x4, sc);
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    v[4 * q] *= __uint_as_float(x4[q].x & 0x7fffffffu); v[4 * q + 1] *= __uint_as_float(x4[q].y & 0x7fffffffu);
                    v[4 * q + 2] *= __uint_as_float(x4[q].z & 0x7fffffffu); v[4 * q + 3] *= __uint_as_float(x4[q].w & 0x7fffffffu);
                }
                ex += (int)sc >> 1;
                if (tr) trp[3] = clock64();
            } else {
                const int cin = child - a.L;
                const float4 *xrow = cond4 + cond_f4(cat, cin, tile, 0, t, w.NI, w.T);      // chunk 0 of the pattern's row (all 64 states)
                uint32_t am = 0;                                                         // anchors of this half (bit k <-> state 32h + k)
                float xa[32];                                                            // this half's values (anchor values are taken from here)
                if (enc & WALK_CHAIN) {
#pragma unroll
                    for (int k = 0; k < 32; k++) { xa[k] = v[k]; v[k] = 1.f; }
                } else {
                    const bool await = (enc & WALK_WAIT) != 0;
                    const uint32_t want = await ? ((uint32_t)code << 31) : 0u;
                    uint4 x4[8];
                    uint32_t sc;
                    load_half_row(xrow + (size_t)8 * h * 128, a.scal + ((size_t)cat * w.NI + cin) * Sp + s, await, want, x4, sc);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        xa[4 * q] = __uint_as_float(x4[q].x & 0x7fffffffu); xa[4 * q + 1] = __uint_as_float(x4[q].y & 0x7fffffffu);
                        xa[4 * q + 2] = __uint_as_float(x4[q].z & 0x7fffffffu); xa[4 * q + 3] = __uint_as_float(x4[q].w & 0x7fffffffu);
                    }
                    ex += (int)sc >> 1;
                }
                // split in two 16-column pieces (registers): anchors out, tf32 hi / lo -> TMEM columns 64 + 32h.. / 128 + 32h..
#pragma unroll
                for (int o = 0; o < 32; o += 16) {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        const float xv = xa[o + k];
                        const bool big = xv >= w.anchor_thr;
                        if (big) am |= 1u << (o + k);
                        const float x = big ? 0.f : xv;
                        const float hh = tf32_rn(x);
                        hi[k] = __float_as_uint(hh);
                        lo[k] = __float_as_uint(x - hh);
                    }
                    HB2_TMEM_ST16(lane_addr + 64 + 32 * h + o, hi, 0);
                    HB2_TMEM_ST16(lane_addr + 128 + 32 * h + o, lo, 0);
                }
                if (tr) trp[8] = clock64();
                s_am[h * 128 + t] = am;                                              // for both threads of the pattern
                if (tr) trp[9] = clock64();
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                if (tr) { trp[10] = clock64(); trp[3] = trp[10]; }
                tc_fence_before();
                __syncthreads();             // (2) A operand complete in TMEM, anchor lists published; previous D consumed
                if (tr) trp[4] = clock64();
                if (warp == 0) {
                    tc_fence_after();
                    mbar_wait(bar_full + (n_step & 1u), (n_step >> 1) & 1u, a.err);
                    if (tr) trp[5] = clock64();
                    const uint32_t baddr = __shfl_sync(0xffffffffu, smem_u32(tab + 64 * WALK_PT_ROW), 0);
                    const uint64_t bdesc_hi = make_b_desc(baddr);
                    const uint64_t bdesc_lo = make_b_desc(baddr + 4096 * 4);
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 8; kk++) tc_mma_tf32_ts(tmem_base, tmem_base + 128 + kk * 8, bdesc_hi + (uint64_t)(kk * 2 * 1024 >> 4), TC_IDESC, kk > 0);
#pragma unroll
                        for (int kk = 0; kk < 8; kk++) tc_mma_tf32_ts(tmem_base, tmem_base + 64 + kk * 8, bdesc_lo + (uint64_t)(kk * 2 * 1024 >> 4), TC_IDESC, 1u);
#pragma unroll
                        for (int kk = 0; kk < 8; kk++) tc_mma_tf32_ts(tmem_base, tmem_base + 64 + kk * 8, bdesc_hi + (uint64_t)(kk * 2 * 1024 >> 4), TC_IDESC, 1u);
                        tc_commit(bar_mma);
                    }
                }
                __syncwarp();
                // anchors of BOTH halves on the CUDA cores while the tensor core works: this thread's 32 parent states.  Values
                // are re-read from the child's block in L2 (loaded, or stored by the two threads of this pattern a moment ago --
                // barrier (2) made those stores visible), a handful of loads in flight under the MMAs.
                unsigned long long amask = (unsigned long long)s_am[t] | ((unsigned long long)s_am[128 + t] << 32);
                const float *xrow_f = reinterpret_cast<const float *>(xrow);
                int ak[WALK_FAST_ANCHORS];
                float av[WALK_FAST_ANCHORS];
#pragma unroll
                for (int ai = 0; ai < WALK_FAST_ANCHORS; ai++) {
                    ak[ai] = -1; av[ai] = 0.f;
                    if (amask) {
                        const int kk = __ffsll((long long)amask) - 1;
                        amask &= amask - 1;
                        ak[ai] = kk;
                        av[ai] = __ldcg(xrow_f + ((size_t)(kk >> 2) * 128) * 4 + (kk & 3));
                    }
                }
                float acc[32];
                mbar_wait(bar_full + (n_step & 1u), (n_step >> 1) & 1u, a.err);
                const float *tabh = tab + 32 * h;
                {
                    const float xv = ak[0] >= 0 ? fabsf(av[0]) : 0.f;
                    const float4 *row = reinterpret_cast<const float4 *>(tabh + max(ak[0], 0) * WALK_PT_ROW);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float4 rr = row[q];
                        acc[4 * q] = xv * rr.x; acc[4 * q + 1] = xv * rr.y; acc[4 * q + 2] = xv * rr.z; acc[4 * q + 3] = xv * rr.w;
                    }
                }
#pragma unroll
                for (int ai = 1; ai < WALK_FAST_ANCHORS; ai++) {
                    if (ak[ai] >= 0) {
                        const float xv = fabsf(av[ai]);
                        const float4 *row = reinterpret_cast<const float4 *>(tabh + ak[ai] * WALK_PT_ROW);
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            const float4 rr = row[q];
                            acc[4 * q] = fmaf(xv, rr.x, acc[4 * q]); acc[4 * q + 1] = fmaf(xv, rr.y, acc[4 * q + 1]);
                            acc[4 * q + 2] = fmaf(xv, rr.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(xv, rr.w, acc[4 * q + 3]);
                        }
                    }
                }
                while (amask) {                  // more than four entries above the threshold (diffuse vectors): rare
                    const int kk = __ffsll((long long)amask) - 1;
                    amask &= amask - 1;
                    const float xv = fabsf(__ldcg(xrow_f + ((size_t)(kk >> 2) * 128) * 4 + (kk & 3)));
                    const float4 *row = reinterpret_cast<const float4 *>(tabh + kk * WALK_PT_ROW);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float4 rr = row[q];
                        acc[4 * q] = fmaf(xv, rr.x, acc[4 * q]); acc[4 * q + 1] = fmaf(xv, rr.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(xv, rr.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(xv, rr.w, acc[4 * q + 3]);
                    }
                }
                mbar_wait(bar_mma, n_mma & 1u, a.err);
                if (tr) trp[6] = clock64();
                tc_fence_after();
#pragma unroll
                for (int o = 0; o < 32; o += 16) {
                    uint32_t d[16];
                    HB2_TMEM_LD16(lane_addr + 32 * h + o, d, 0);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int k = 0; k < 16; k++) v[o + k] *= (__uint_as_float(d[k]) + acc[o + k]);
                }
                n_mma++;
            }
            n_step++;
            if (!(flags & STEP_LAST)) {
                renorm_half(v, ex);
            } else {
                if (a.L + par == a.forced_node) {                                      // pinned internal node
                    const int f = __ldg(a.forced + s) - 32 * h;
#pragma unroll
                    for (int k = 0; k < 32; k++) if (k != f) v[k] = 0.f;
                }
                // the job is complete: ONE power-of-two scale for the whole row.  Exchange (half maximum, half exponent).
                float m0 = v[0], m1 = v[1], m2 = v[2], m3 = v[3];
#pragma unroll
                for (int k = 4; k < 32; k += 4) { m0 = fmaxf(m0, v[k]); m1 = fmaxf(m1, v[k + 1]); m2 = fmaxf(m2, v[k + 2]); m3 = fmaxf(m3, v[k + 3]); }
                const float mh = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                s_mx[h * 128 + t] = mh;
                s_ex[h * 128 + t] = ex;
                __syncthreads();             // (3) once per job
                const float mp = s_mx[(h ^ 1) * 128 + t];
                const int exp_p = s_ex[(h ^ 1) * 128 + t];
                // true exponent of each half's maximum (m = f * 2^e, f in [0.5,1)); a half that is all zero does not count
                const bool okh = mh > 0.f && mh < INFINITY, okp = mp > 0.f && mp < INFINITY;
                const int Eh = ex + max((int)((__float_as_uint(mh) >> 23) & 0xffu) - 126, -125);
                const int Ep = exp_p + max((int)((__float_as_uint(mp) >> 23) & 0xffu) - 126, -125);
                const int Et = (okh && okp) ? max(Eh, Ep) : okh ? Eh : okp ? Ep : ex;
                if (okh || okp) {
                    const int d = max(-127, min(ex - Et, 127));                         // this half's values are multiplied by 2^d
                    if (d != 0) {
                        const float sc = d < -126 ? 0.f : __uint_as_float((uint32_t)(127 + d) << 23);
#pragma unroll
                        for (int k = 0; k < 32; k++) v[k] *= sc;
                    }
                    ex = Et;
                }
                float4 *outp = cond4 + cond_f4(cat, par, tile, 8 * h, t, w.NI, w.T);
#pragma unroll
                for (int q = 0; q < 8; q++)
                    __stcg(reinterpret_cast<uint4 *>(outp + (size_t)q * 128),
                           make_uint4(__float_as_uint(v[4 * q]) | tagbit, __float_as_uint(v[4 * q + 1]) | tagbit,
                                      __float_as_uint(v[4 * q + 2]) | tagbit, __float_as_uint(v[4 * q + 3]) | tagbit));
                if (h == 0) __stcg(a.scal + ((size_t)cat * w.NI + par) * Sp + s, (int)(((uint32_t)ex << 1) | (tagbit >> 31)));
                if (par == a.I - 1) {
                    double rr = 0.0;
#pragma unroll
                    for (int k = 0; k < 32; k++) rr = fma((double)v[k], a.pi[32 * h + k], rr);
                    s_rt[h * 128 + t] = rr;
                    __syncthreads();
                    if (h == 0) {
                        a.rootL[(size_t)cat * Sp + s] = s_rt[t] + s_rt[128 + t];
                        a.rootE[(size_t)cat * Sp + s] = ex;
                    }
                }
            }
            if (tr) trp[7] = clock64();
            st = nx;
            nx = nx2;
        }
    }
    if (w.trace_cta_times && tid == 0) w.trace_cta_times[(size_t)blockIdx.x * 4 + 2] = clock64();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
    }
}

}  // namespace hb2
