// hb2_kernels_lanes.cuh -- fp64 pruning pass for 33..64 states in ONE launch with K in-order lanes per (class, tile).
//
// Replaces, for the double-precision path, both the per-level launches of prune64_kernel (54 launches on the north-star
// tree, each bound by launch + load latency) and prune64_walk_kernel (one CTA per (class, tile) walking ALL dirty nodes
// serially: 120 CTAs x 198 nodes).  Same decomposition as the tensor walk kernel (hb2_kernels_tc.cuh): the host planner
// (hb2_engine.cu plan_walk) list-schedules the dirty jobs -- nodes and side products of nodes with two internal children
// -- on K lanes; CTA (class c, 64-pattern tile t, lane r) executes its lane's steps back to back.  Reference semantics:
// ComputeTreeBlockByBranch (tree_evaluator.cpp:3556-4171) restricted to the nodes DetermineNodesForUpdate marked.
//   * hand-over between lanes: a flag per (class, job, tile) holding the id of the pass that produced the tile
//     (st.release.gpu by the producer after its stores, ld.acquire.gpu poll by the consumer, time-bounded);
//   * a child produced by THIS CTA in its previous job goes straight from registers to the shared-memory operand;
//   * the transition matrix of the NEXT contraction is staged with cp.async while the current one is multiplied
//     (two P buffers + one X buffer = 99 KB: two CTAs per SM);
//   * tile product: tile_mm64 (4x4 register blocks, DFMA at the FP64 pipe's rate; hb2_kernels_fp64.cuh).
// Per element the products are those of prune64_kernel; the children of a node are multiplied in plan order (chain child,
// leaves, other internal children, side product), so results agree with the per-level kernel to rounding (parity bound
// 1e-10 relative on lnL, observed 1e-15), and are deterministic for a given plan.
#pragma once
#include "hb2_kernels_fp64.cuh"
#include "hb2_kernels_tc.cuh"      // step encoding (WALK_*, STEP_*), globaltimer_ns, HB2_WAIT_LIMIT_NS

namespace hb2 {

struct LaneArgs {
    PruneArgs a;
    double *cond_side;          // [C][I][Sp][64] side products (job I + n)
    int *scal_side;             // [C][I][Sp]
    const int *lane_start;      // [K+1]
    const int2 *steps;
    int *flags;                 // [C][2I][T] id of the pass that last produced (class, job, tile)
    int *err;
    int K, T, ncls, nslots, pass;
};

constexpr int LANES_SMEM_BYTES = 3 * 64 * LD64 * (int)sizeof(double);

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256, 2) prune64_lanes_kernel(LaneArgs w) {
    extern __shared__ __align__(16) double sm[];
    double *Xs = sm;                              // [64][LD64] child conditionals (pattern-major)
    double *Pb = sm + 64 * LD64;                  // 2 x [64][LD64] transition matrices (transposed), ring
    const PruneArgs &a = w.a;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const size_t Sp = a.Sp;
    const int r = blockIdx.x % w.K;
    const int i_begin = w.lane_start[r], i_end = w.lane_start[r + 1];
    if (i_begin == i_end) return;

    auto is_contraction = [&](int enc) { return (enc & WALK_ID_MASK) >= a.L && !(enc & WALK_MUL); };
    auto next_contraction = [&](int from) {
        int j = from;
        while (j < i_end && !is_contraction(__ldg(&w.steps[j].x))) j++;
        return j;
    };
    // 64 x 64 doubles, contiguous in global memory -> padded rows in shared memory, 8 x 16 bytes per thread
    auto stage_tile = [&](double *dst, const double *src) {
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int idx = tid + 256 * m, rr = idx >> 5, c2 = (idx & 31) * 2;
            cp_async16(dst + rr * LD64 + c2, src + rr * 64 + c2);
        }
    };
    // wait until (class, job, tile) has been produced by this pass.  Uniform call (contains a barrier).
    auto wait_flag = [&](int cat, int job, int tile) {
        if (tid == 0) {
            const int *f = w.flags + ((size_t)cat * 2 * a.I + job) * w.T + tile;
            unsigned long long t0 = 0;
            for (int it = 0;; it++) {
                int got;
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(got) : "l"(f) : "memory");
                if (got == w.pass) break;
                if ((it & 63) == 63) {                                          // never hang the GPU: time-bounded
                    const unsigned long long now = globaltimer_ns();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > HB2_WAIT_LIMIT_NS) { atomicExch(w.err, 2); break; }
                }
            }
        }
        __syncthreads();
    };

    for (int ct = blockIdx.x / w.K; ct < w.ncls * w.T; ct += w.nslots) {
        const int cat = a.cat0 + ct / w.T;
        const int tile = ct % w.T;
        const int s0 = tile * TILE_P;
        double v[4][4];
        int ex[4];
        int cur = 0;
        __syncthreads();                          // previous (class, tile): the buffers are free
        {
            const int j0 = next_contraction(i_begin);
            if (j0 < i_end) {
                stage_tile(Pb, a.PT + ((size_t)cat * a.B + (__ldg(&w.steps[j0].x) & WALK_ID_MASK)) * 4096);
                cp_async_commit();
            }
        }
        for (int i = i_begin; i < i_end; i++) {
            const int2 st = __ldg(w.steps + i);
            const int enc = st.x, flags = st.y;
            const int child = enc & WALK_ID_MASK;
            const int job = flags & WALK_ID_MASK;            // node (< I) or side product (I + node)
            if ((flags & STEP_FIRST) && !(enc & WALK_CHAIN)) {
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    ex[ii] = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) v[ii][j] = 1.0;
                }
            }
            if (child < a.L) {
                // leaf: column gather PT[state][k]; ambiguous: sum_j amb[j] PT[j][k]
                const double *PT = a.PT + ((size_t)cat * a.B + child) * 4096;
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    const int code = (child == a.forced_node) ? a.forced[s0 + 4 * ty + ii] : __ldg(a.leaf + (size_t)child * Sp + s0 + 4 * ty + ii);
                    double m0, m1, m2, m3;
                    if (code >= 0) {
                        const double2 *p = reinterpret_cast<const double2 *>(PT + (size_t)code * 64 + 2 * tx);
                        const double2 q0 = __ldg(p), q1 = __ldg(p + 16);
                        m0 = q0.x; m1 = q0.y; m2 = q1.x; m3 = q1.y;
                    } else {
                        const double *amb = a.ambig + (size_t)(-code - 1) * 64;
                        m0 = m1 = m2 = m3 = 0.0;
                        for (int j = 0; j < a.D; j++) {
                            const double wj = __ldg(amb + j);
                            if (wj != 0.0) {
                                const double2 *p = reinterpret_cast<const double2 *>(PT + (size_t)j * 64 + 2 * tx);
                                const double2 q0 = __ldg(p), q1 = __ldg(p + 16);
                                m0 = fma(wj, q0.x, m0); m1 = fma(wj, q0.y, m1); m2 = fma(wj, q1.x, m2); m3 = fma(wj, q1.y, m3);
                            }
                        }
                    }
                    v[ii][0] *= m0; v[ii][1] *= m1; v[ii][2] *= m2; v[ii][3] *= m3;
                }
            } else if (enc & WALK_MUL) {
                // side product of this node (its other children, contracted by another lane or earlier in this one)
                const int n = child - a.L - a.I;
                if (enc & WALK_WAIT) wait_flag(cat, child - a.L, tile);
                const double *X = w.cond_side + (((size_t)cat * a.I + n) * Sp + s0) * 64;
                const int *sc = w.scal_side + ((size_t)cat * a.I + n) * Sp + s0 + 4 * ty;
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    const double2 q0 = __ldcg(reinterpret_cast<const double2 *>(X + (size_t)(4 * ty + ii) * 64 + 2 * tx));
                    const double2 q1 = __ldcg(reinterpret_cast<const double2 *>(X + (size_t)(4 * ty + ii) * 64 + 32 + 2 * tx));
                    v[ii][0] *= q0.x; v[ii][1] *= q0.y; v[ii][2] *= q1.x; v[ii][3] *= q1.y;
                    ex[ii] += __ldcg(sc + ii);
                }
            } else {
                const int cin = child - a.L;
                __syncthreads();                  // (A) every thread is past the previous product: Xs and Pb[cur^1] are free
                int sc4[4] = {0, 0, 0, 0};
                if (enc & WALK_CHAIN) {
                    // the child is this lane's previous job: its renormalised values are still in v, its exponents in ex
#pragma unroll
                    for (int ii = 0; ii < 4; ii++) {
                        *reinterpret_cast<double2 *>(Xs + (4 * ty + ii) * LD64 + 2 * tx) = make_double2(v[ii][0], v[ii][1]);
                        *reinterpret_cast<double2 *>(Xs + (4 * ty + ii) * LD64 + 32 + 2 * tx) = make_double2(v[ii][2], v[ii][3]);
#pragma unroll
                        for (int j = 0; j < 4; j++) v[ii][j] = 1.0;
                    }
                } else {
                    if (enc & WALK_WAIT) wait_flag(cat, cin, tile);
                    stage_tile(Xs, a.cond + (((size_t)cat * a.I + cin) * Sp + s0) * 64);
                    const int *sc = a.scal + ((size_t)cat * a.I + cin) * Sp + s0 + 4 * ty;
#pragma unroll
                    for (int ii = 0; ii < 4; ii++) sc4[ii] = __ldcg(sc + ii);
                }
                cp_async_commit();                // group of X (empty on a chain)
                const int jn = next_contraction(i + 1);
                if (jn < i_end) {                 // stage the next contraction's matrix behind this step's product
                    stage_tile(Pb + (cur ^ 1) * 64 * LD64, a.PT + ((size_t)cat * a.B + (__ldg(&w.steps[jn].x) & WALK_ID_MASK)) * 4096);
                    cp_async_commit();
                    cp_async_wait<1>();           // everything but the newest group: this step's P and X have landed
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                double acc[4][4];
#pragma unroll
                for (int ii = 0; ii < 4; ii++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[ii][j] = 0.0;
                tile_mm64(Xs, Pb + cur * 64 * LD64, tx, ty, acc);
                cur ^= 1;
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    ex[ii] += sc4[ii];
#pragma unroll
                    for (int j = 0; j < 4; j++) v[ii][j] *= acc[ii][j];
                }
            }
            if (flags & STEP_LAST) {
                const bool side = job >= a.I;
                const int node = side ? job - a.I : job;
                if (!side && a.L + node == a.forced_node) {            // pinned internal node: only the forced state survives
#pragma unroll
                    for (int ii = 0; ii < 4; ii++) {
                        const int f = a.forced[s0 + 4 * ty + ii];
#pragma unroll
                        for (int j = 0; j < 4; j++) if (col_of(tx, j) != f) v[ii][j] = 0.0;
                    }
                }
                // per-pattern renormalisation: exact power of two so that max_k lands in [0.5, 1)
                const bool is_root = !side && node == a.I - 1;
                double *outp = (side ? w.cond_side : a.cond) + (((size_t)cat * a.I + node) * Sp + s0) * 64;
                int *outs = (side ? w.scal_side : a.scal) + ((size_t)cat * a.I + node) * Sp + s0;
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    double m = fmax(fmax(v[ii][0], v[ii][1]), fmax(v[ii][2], v[ii][3]));
#pragma unroll
                    for (int o = 1; o < 16; o <<= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
                    int e = 0;
                    if (m > 0.0 && m < INFINITY) {
                        e = ilogb(m) + 1;
                        const double s1 = exp2i(-(e / 2)), s2 = exp2i(-(e - e / 2));   // two steps: |e| may exceed 1022
#pragma unroll
                        for (int j = 0; j < 4; j++) v[ii][j] = v[ii][j] * s1 * s2;
                    }
                    ex[ii] += e;
                    const int row = 4 * ty + ii;
                    __stcg(reinterpret_cast<double2 *>(outp + (size_t)row * 64 + 2 * tx), make_double2(v[ii][0], v[ii][1]));
                    __stcg(reinterpret_cast<double2 *>(outp + (size_t)row * 64 + 32 + 2 * tx), make_double2(v[ii][2], v[ii][3]));
                    if (tx == 0) __stcg(outs + row, ex[ii]);
                    if (is_root) {
                        double rr = v[ii][0] * a.pi[2 * tx] + v[ii][1] * a.pi[2 * tx + 1] + v[ii][2] * a.pi[32 + 2 * tx] + v[ii][3] * a.pi[33 + 2 * tx];
#pragma unroll
                        for (int o = 1; o < 16; o <<= 1) rr += __shfl_xor_sync(0xffffffffu, rr, o);
                        if (tx == 0) {
                            a.rootL[(size_t)cat * Sp + s0 + row] = rr;
                            a.rootE[(size_t)cat * Sp + s0 + row] = ex[ii];
                        }
                    }
                }
                __syncthreads();                  // the whole tile is written (and visible to this CTA's later reads)
                if (tid == 0) {                   // publish: other lanes of this (class, tile) may consume it now
                    __threadfence();
                    int *f = w.flags + ((size_t)cat * 2 * a.I + job) * w.T + tile;
                    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(f), "r"(w.pass) : "memory");
                }
            }
        }
    }
}

}  // namespace hb2
