// hb2_kernels_fp64.cuh -- fp64 CUDA kernels of the likelihood hot path (sm_100a).
//
//   expm64_kernel / expm_small_kernel   P = exp(Q t) per (branch, rate class), one CTA per matrix
//                                       (replaces _Matrix::Exponentiate, reference matrix.cpp:5537-5951)
//   prune64_kernel / prune_small_kernel fused Felsenstein update of one internal node for a tile of patterns:
//                                       all children folded (leaf column gather, ambiguity mat-vec, internal
//                                       contraction), per-pattern renormalisation, root pi-dot
//                                       (replaces ComputeTreeBlockByBranch, reference tree_evaluator.cpp:3556-4171)
//   combine_kernel / final_sum_kernel   rate-class mixture, log, pattern-frequency weighting, deterministic sum
//                                       (replaces likefunc2.cpp:828-859,1446-1506 and tree_evaluator.cpp:4093-4149)
//
// Device data layout (Dp = padded state count, Sp = pattern count padded to 64):
//   PT    [C][B][Dp][Dp]   TRANSPOSED transition matrices: PT[j][k] = P(parent k -> child j); padding = 0
//   cond  [C][I][Sp][Dp]   conditionals of internal nodes, renormalised per (node, pattern) so max_k = [0.5,1)
//   scal  [C][I][Sp]       int32 binary exponent: true conditional = cond * 2^scal
//   leaf  [L][Sp]          int32 state code (>=0) or -(ambiguity row + 1)
//   ambig [nAmb][Dp]       0/1 rows
// Underflow handling is the engine's own (exact power-of-two renormalisation at every node) and is converted to the
// reference's 2^64-count convention only at the boundary (last bullet).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace hb2 {

constexpr int TILE_P = 64;          // patterns per CTA in the 64-state kernel
constexpr int LD64 = 66;            // smem leading dimension (doubles) for 64x64 operands: conflict-free row pairs

struct TreeDev {
    const int *child_start;         // [I+1]
    const int *child_ids;           // flat node ids of the children of each internal node
};

struct PruneArgs {
    const double *PT;               // [C][B][Dp*Dp]
    double *cond;                   // [C][I][Sp][Dp]
    int *scal;                      // [C][I][Sp]
    const int *leaf;                // [L][Sp]
    const double *ambig;            // [nAmb][Dp]
    const double *pi;               // [Dp] root frequencies (padding 0)
    double *rootL;                  // [C][Sp]
    int *rootE;                     // [C][Sp]
    TreeDev tree;
    int L, I, B, D, Sp, cat0;
    // forced states (reference setBranch / setBranchTo, tree_evaluator.cpp:3624,173-181,4059): flat id of ONE node whose
    // state is pinned per pattern to forced[s] (a leaf: its observed state is replaced; an internal node or the root: its
    // conditional vector is masked to that state), or -1
    const int *forced;
    int forced_node;
};

__device__ __forceinline__ double exp2i(int e) {   // exact 2^e for |e| < 1022
    return __longlong_as_double((long long)(e + 1023) << 52);
}

// ------------------------------------------------------------------------------------------------
// 64x64x64 fp64 tile product, 256 threads; thread (tx,ty) owns rows 4ty..4ty+3 and columns
// {2tx, 2tx+1, 32+2tx, 33+2tx} of C (col_of):  acc[i][j] += sum_k A[(4*ty+i)*LD64 + k] * B[k*LD64 + col_of(tx,j)]
// A rows are read as broadcasts (two distinct rows per warp, 16 banks apart thanks to LD64); each B read is one
// 16-byte vector per thread, 256 contiguous bytes per half-warp: conflict-free.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int col_of(int tx, int j) { return (j < 2) ? 2 * tx + j : 32 + 2 * tx + (j - 2); }

__device__ __forceinline__ void tile_mm64(const double *__restrict__ A, const double *__restrict__ Bm,
                                          int tx, int ty, double (&acc)[4][4]) {
    const double *a0 = A + (4 * ty) * LD64;
    const double *b0 = Bm + 2 * tx;
#pragma unroll 4
    for (int k = 0; k < 64; k += 2) {
        double2 a[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = *reinterpret_cast<const double2 *>(a0 + i * LD64 + k);
        double2 b00 = *reinterpret_cast<const double2 *>(b0 + k * LD64);
        double2 b01 = *reinterpret_cast<const double2 *>(b0 + k * LD64 + 32);
        double2 b10 = *reinterpret_cast<const double2 *>(b0 + (k + 1) * LD64);
        double2 b11 = *reinterpret_cast<const double2 *>(b0 + (k + 1) * LD64 + 32);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            acc[i][0] = fma(a[i].x, b00.x, acc[i][0]);
            acc[i][1] = fma(a[i].x, b00.y, acc[i][1]);
            acc[i][2] = fma(a[i].x, b01.x, acc[i][2]);
            acc[i][3] = fma(a[i].x, b01.y, acc[i][3]);
            acc[i][0] = fma(a[i].y, b10.x, acc[i][0]);
            acc[i][1] = fma(a[i].y, b10.y, acc[i][1]);
            acc[i][2] = fma(a[i].y, b11.x, acc[i][2]);
            acc[i][3] = fma(a[i].y, b11.y, acc[i][3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Matrix exponential, 64-padded states.  One CTA (256 threads) per matrix.
//   in : Q [n][D*D] row-major rate matrices (Q*t, diagonal = -rowsum), dst[n] = slot index into PT
//   out: PT[slot] = exp(Q)^T with padding rows/cols zeroed, negatives clamped, columns repaired to sum to 1
// Algorithm (engine's own, same result as the reference's Taylor+squaring to ~1e-15): scale by 2^-s so that
// ||A||_inf <= 1, degree-15 Taylor polynomial by Paterson-Stockmeyer with A^2,A^3,A^4 (6 products), s squarings.
// mix_w != nullptr: PT[slot] (+)= w * result, used for explicit-form mixtures (sum_k w_k Exp(Q_k)).
// ------------------------------------------------------------------------------------------------
struct ExpmArgs {
    const double *Q;        // [n][D*D]
    const int *dst;         // [n] destination slot; < 0: retired entry (the slot was handed over again), skipped
    double *PT;
    const double *colsum;   // nullable [n][16][64]: column-sum partials of expm_poly_kernel; non-null = expm64_dmma_kernel also
                            // does the row repair of the entries that kernel finished (no separate expm_diag_kernel launch)
    double *Qres;           // nullable [slots][D*D]: resident copy of the rate matrices, written while loading
    int D;
    int is_trans;           // 1: input already a transition matrix -> just transpose/pad
    // compiled-template input (replaces dense Q when tmpl_nnz > 0): Q[row][col] = V[blk][formula[e]] * colfreq[col],
    // diagonal = -(off-diagonal row sum); mirrors _CompiledMatrixData + MultByFreqs (reference matrix.h:69, matrix.cpp:1546)
    const double *V;        // [n][nF] formula values per matrix
    const int *tmpl_index;  // [nnz] row*D + col
    const int *tmpl_formula;// [nnz]
    const double *tmpl_colfreq;  // nullable [D]
    int tmpl_nnz, nF;
    // shared-powers path (64-state kernels, see expm_classify_kernel): per-entry group, proportionality flag and L1 weight
    const int *group;       // nullable [n]: power-table group of the entry (rate class; mixture component), < 0 none
    const int *flag;        // nullable [n]: 1 = the entry is a scalar multiple of its group's reference direction
    const double *weight;   // [n]: sum of |input values| of the entry (formula values or matrix entries)
    const struct ExpmGroup *groups;   // [G] reference info of every group
    const double *pow;      // [G][EXPM_POW_TERMS][4096] Taylor coefficient matrices of the groups (PT layout)
};

// Shared powers.  In the models the optimiser spends its time in, all branches of a rate class share ONE rate matrix up
// to the branch length: A_b = rho_b * R with ||R||_inf = 1 (VERDICT r1 "what's weak" 4).  Then
//     exp(A_b) = sum_m (R^m / m!) * rho_b^m
// is, element by element, a scalar polynomial in rho_b whose coefficient matrices S_m = R^m/m! are the same for every
// branch of the group: they are built ONCE per group and evaluation (expm_powers_kernel: 23 products of 64^3 arranged as
// a 5-level tree S_{a+b} = S_a S_b / C(a+b,a), one CTA per product, dependencies resolved through generation flags),
// after which a branch costs 24 FMAs per matrix element (expm_poly_kernel) instead of ~5 matrix products; only matrices
// with rho_b > 2.3 still need squarings (s of them, x = rho_b / 2^s <= 2.3; degree 24 truncates below 1e-16 there).  Proportionality is CHECKED on the device for every matrix
// (expm_classify_kernel), never assumed: anything that is not a multiple of its group's reference direction to 1e-13
// takes the general scaling-and-squaring path inside the same launch.
constexpr int EXPM_POW_TERMS = 25;        // S_0 = I (not stored, slot unused) .. S_24
constexpr double EXPM_POLY_THETA = 2.3;   // x^25/25! < 1e-16 for x <= 2.33: degree 24 needs no squaring up to here
struct ExpmGroup {
    double nu;        // ||A_ref||_inf of the reference matrix the powers were built from
    double weight;    // its L1 weight (same measure as ExpmArgs::weight)
    int kind;         // 0 invalid, 1 reference direction held as formula values (compiled input), 2 as dense matrix entries
    int pad;
};

// Builds A1[j][i] = Q[i][j] (transposed, zero padded, leading dimension LD) from dense or compiled input and leaves a
// dense resident copy in Qres when asked.  All threads of the CTA call it; ends with a barrier.
template <int DP, int LD, int NT>
__device__ __forceinline__ void load_rate_matrix(const ExpmArgs &a, double *A1, int tid, int entry = -1) {
    const int D = a.D;
    if (entry < 0) entry = blockIdx.x;
    const size_t slot = a.dst[entry];
    if (a.tmpl_nnz > 0) {
        for (int idx = tid; idx < DP * LD; idx += NT) A1[idx] = 0.0;
        __syncthreads();
        const double *V = a.V + (size_t)entry * a.nF;
        for (int e = tid; e < a.tmpl_nnz; e += NT) {
            const int rc = a.tmpl_index[e], r = rc / D, c = rc - r * D;
            double v = V[a.tmpl_formula[e]];
            if (a.tmpl_colfreq) v *= a.tmpl_colfreq[c];
            if (r != c) A1[c * LD + r] = v;
        }
        __syncthreads();
        if (tid < D) {
            double s = 0.0;
            for (int c = 0; c < D; c++) if (c != tid) s += A1[c * LD + tid];
            A1[tid * LD + tid] = -s;
        }
        __syncthreads();
        if (a.Qres)
            for (int idx = tid; idx < D * D; idx += NT) { const int i = idx / D, j = idx - i * D; a.Qres[slot * D * D + idx] = A1[j * LD + i]; }
    } else {
        const double *Q = a.Q + (size_t)entry * D * D;
        for (int idx = tid; idx < DP * DP; idx += NT) {
            const int i = idx / DP, j = idx - i * DP;
            const double v = (i < D && j < D) ? Q[(size_t)i * D + j] : 0.0;
            A1[j * LD + i] = v;
            if (a.Qres && i < D && j < D) a.Qres[slot * D * D + (size_t)i * D + j] = v;
        }
    }
    __syncthreads();
}

__constant__ double c_taylor[16] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
                                    1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0,
                                    1.0 / 87178291200.0, 1.0 / 1307674368000.0};
__device__ __forceinline__ double taylor_c(int k) { return c_taylor[k]; }

__global__ void __launch_bounds__(256, 1) expm64_kernel(ExpmArgs a) {
    extern __shared__ __align__(16) double sm[];
    double *A1 = sm, *A2 = A1 + 64 * LD64, *A3 = A2 + 64 * LD64, *A4 = A3 + 64 * LD64, *R = A4 + 64 * LD64;
    __shared__ double red[64];
    __shared__ int s_shift;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int D = a.D;
    if (a.dst[blockIdx.x] < 0) return;
    double *out = a.PT + (size_t)a.dst[blockIdx.x] * 4096;

    load_rate_matrix<64, LD64, 256>(a, A1, tid);
    if (a.is_trans) {
        for (int idx = tid; idx < 4096; idx += 256) out[idx] = A1[(idx >> 6) * LD64 + (idx & 63)];
        return;
    }
    // ||A||_1 of A^T == ||Q||_inf: max over columns i of sum_j |A1[j][i]|
    if (tid < 64) {
        double s = 0.0;
        for (int j = 0; j < 64; j++) s += fabs(A1[j * LD64 + tid]);
        red[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        double m = 0.0;
        bool bad = false;
        for (int i = 0; i < 64; i++) { if (!(red[i] == red[i]) || isinf(red[i])) bad = true; m = fmax(m, red[i]); }
        int e = 0;
        if (m > 0.0) frexp(m, &e);
        s_shift = bad ? -1 : max(e, 0);
    }
    __syncthreads();
    const int shift = s_shift;
    if (shift < 0 || shift > 900) {           // NaN/inf or absurd rates: propagate NaN (host sees NaN lnL)
        for (int idx = tid; idx < 4096; idx += 256) out[idx] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    if (shift > 0) {
        const double sc = exp2i(-shift);
        for (int idx = tid; idx < 64 * 64; idx += 256) A1[(idx >> 6) * LD64 + (idx & 63)] *= sc;
        __syncthreads();
    }
    double acc[4][4];
    auto zero = [&]() {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
    };
    auto store = [&](double *M) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j += 2)
                *reinterpret_cast<double2 *>(M + (4 * ty + i) * LD64 + col_of(tx, j)) = make_double2(acc[i][j], acc[i][j + 1]);
    };
    zero(); tile_mm64(A1, A1, tx, ty, acc); store(A2);
    __syncthreads();
    zero(); tile_mm64(A2, A1, tx, ty, acc); store(A3);
    zero(); tile_mm64(A2, A2, tx, ty, acc); store(A4);
    // R = B3 = c12 I + c13 A + c14 A2 + c15 A3
    auto poly_block = [&](int q, int i, int j) -> double {
        int r = 4 * ty + i, c = col_of(tx, j), o = r * LD64 + c;
        double v = taylor_c(4 * q + 1) * A1[o] + taylor_c(4 * q + 2) * A2[o];
        if (r == c) v += taylor_c(4 * q);
        return v;   // the A3 term is added by the caller after the barrier that publishes A3
    };
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int o = (4 * ty + i) * LD64 + col_of(tx, j);
            R[o] = poly_block(3, i, j) + taylor_c(15) * A3[o];
        }
    __syncthreads();
    for (int q = 2; q >= 0; q--) {               // R = R*A4 + B_q   (all matrices are polynomials in A: they commute)
        zero(); tile_mm64(R, A4, tx, ty, acc);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int o = (4 * ty + i) * LD64 + col_of(tx, j);
                R[o] = acc[i][j] + poly_block(q, i, j) + taylor_c(4 * q + 3) * A3[o];
            }
        __syncthreads();
    }
    for (int s = 0; s < shift; s++) {            // squarings
        zero(); tile_mm64(R, R, tx, ty, acc);
        __syncthreads();
        store(R);
        __syncthreads();
    }
    // clamp negatives, zero the padding, repair: column k of PT (= row k of P) must sum to 1
    for (int idx = tid; idx < 64 * 64; idx += 256) {
        int j = idx >> 6, k = idx & 63;
        double v = R[j * LD64 + k];
        if (j >= D || k >= D) v = 0.0; else if (v < 0.0) v = 0.0;
        R[j * LD64 + k] = v;
    }
    __syncthreads();
    if (tid < 64) {
        double s = 0.0;
        for (int j = 0; j < 64; j++) if (j != tid) s += R[j * LD64 + tid];
        if (tid < D) R[tid * LD64 + tid] = fmax(1.0 - s, 0.0);
    }
    __syncthreads();
    for (int idx = tid; idx < 4096; idx += 256) out[idx] = R[(idx >> 6) * LD64 + (idx & 63)];
}


// ------------------------------------------------------------------------------------------------
// expm64_dmma_kernel: same contract as expm64_kernel, products on the FP64 tensor pipe (mma.sync.m8n8k4.f64, DMMA)
// and a Paterson-Stockmeyer polynomial whose degree follows the norm:
//   p(A) = sum_{q=0..Q} A3^q * (c_{3q} I + c_{3q+1} A + c_{3q+2} A2),   A3 = A^3,   degree 3Q+2,   2+Q products,
//   Q chosen so that theta^(3Q+3)/(3Q+3)! <= 1e-16 with theta = ||A||/2^s <= 0.975; s squarings follow.
// 8 warps; warp (wr,wc) owns rows 16wr..16wr+15 and columns 32wc..32wc+31 of the product as 2x4 m8n8 tiles; lane
// (g = lane/4, q = lane%4) holds element (row 16wr+8i+g, col 32wc+8j+2q+e) in acc[i][j][e].
// Optional fused epilogue for the tensor-core pruning path: PB (hi/lo canonical tiles of P) and PTf (fp32 PT).
// ------------------------------------------------------------------------------------------------
__constant__ double c_taylor18[18] = {1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320,
                                      1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0,
                                      1.0 / 87178291200.0, 1.0 / 1307674368000.0, 1.0 / 20922789888000.0,
                                      1.0 / 355687428096000.0};

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__device__ __forceinline__ void tile_mm64_dmma(const double *__restrict__ A, const double *__restrict__ Bm, int wr, int wc,
                                               int g, int q, double (&acc)[2][4][2]) {
    const double *a0 = A + (16 * wr + g) * LD64 + q;
    const double *b0 = Bm + q * LD64 + 32 * wc + g;
#pragma unroll 4
    for (int k0 = 0; k0 < 64; k0 += 4) {
        const double x0 = a0[k0], x1 = a0[8 * LD64 + k0];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const double b = b0[k0 * LD64 + 8 * j];
            dmma884(acc[0][j], x0, b);
            dmma884(acc[1][j], x1, b);
        }
    }
}

struct ExpmTcOut {
    float *PB;      // nullable: [slots][2][4096] canonical hi/lo tiles (see hb2_kernels_tc.cuh)
    float *PTf;     // [slots][64 rows][68 floats] (rows padded: conflict-free shared-memory gathers after a flat bulk copy)
};

__device__ __forceinline__ float tf32_rn_dev(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}


// ------------------------------------------------------------------------------------------------
// Shared-powers path, stage 1: one small CTA per batch entry.  Measures the entry (weight = sum of |values|: formula
// values of a compiled hand-over, matrix entries of a dense one), keeps the resident copy of the input (Vres / Qres:
// hb2_time_resident replays from it) and decides whether the entry is a scalar multiple of its group's reference
// direction:  |v_k[f] * w_ref - v_ref[f] * w_k| <= 1e-13 * w_k * w_ref  for every value f.  The reference is either
// entry ref[k] of this batch (its powers are (re)built by expm_powers_kernel in this launch sequence) or, for small
// batches, the direction the group's power table was last built from (refvec, kept on the device).
// ------------------------------------------------------------------------------------------------
struct ExpmClassifyArgs {
    const double *V;          // compiled input [n][nV] or dense input [n][D*D] (nV = D*D)
    int nV;
    int kind;                 // 1 compiled, 2 dense
    const int *dst;           // [n]
    const int *group;         // [n]
    const int *ref;           // [n] batch index of the group's reference entry, or -1: use the cached direction
    const ExpmGroup *groups;  // [G]
    const double *refvec;     // [G][nVmax] cached reference directions
    int refvec_stride;
    double *weight;           // out [n]
    int *flag;                // out [n]
    double *res;              // nullable: resident copy [slots][nV]
};

__global__ void __launch_bounds__(128) expm_classify_kernel(ExpmClassifyArgs a) {
    __shared__ double red[2][4];
    __shared__ int bad[4];
    const int k = blockIdx.x, tid = threadIdx.x;
    const int slot = a.dst[k];
    if (slot < 0) { if (tid == 0) { a.flag[k] = 0; a.weight[k] = 0.0; } return; }
    const double *v = a.V + (size_t)k * a.nV;
    const int g = a.group[k], r = a.ref[k];
    const double *vr = nullptr;
    double wr_cached = -1.0;
    if (g >= 0) {
        if (r >= 0) vr = a.V + (size_t)r * a.nV;
        else if (a.groups[g].kind == a.kind) { vr = a.refvec + (size_t)g * a.refvec_stride; wr_cached = a.groups[g].weight; }
    }
    double wk = 0.0, wr = 0.0;
    for (int f = tid; f < a.nV; f += 128) {
        const double x = v[f];
        wk += fabs(x);
        if (vr) wr += fabs(vr[f]);
        if (a.res) a.res[(size_t)slot * a.nV + f] = x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { wk += __shfl_xor_sync(0xffffffffu, wk, o); wr += __shfl_xor_sync(0xffffffffu, wr, o); }
    if ((tid & 31) == 0) { red[0][tid >> 5] = wk; red[1][tid >> 5] = wr; }
    __syncthreads();
    wk = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    wr = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    if (wr_cached >= 0.0) wr = wr_cached;
    int ok = (vr != nullptr) && (wr > 0.0) && (wk >= 0.0) && (wk < INFINITY) && (wr < INFINITY);   // NaN fails every comparison
    if (ok) {
        const double tol = 1e-13 * wk * wr;
        for (int f = tid; f < a.nV; f += 128)
            if (!(fabs(v[f] * wr - vr[f] * wk) <= tol)) ok = 0;
    }
    ok = __all_sync(0xffffffffu, ok);
    if ((tid & 31) == 0) bad[tid >> 5] = !ok;
    __syncthreads();
    if (tid == 0) {
        a.flag[k] = !(bad[0] | bad[1] | bad[2] | bad[3]);
        a.weight[k] = wk;
    }
}

// The same classification with one WARP per entry, eight entries per CTA, for short inputs (compiled hand-overs: ~50 formula
// values): 1588 CTAs of 128 threads that each touch 47 doubles cost 17 us of launch and drain; n/8 CTAs are one wave.
// (Sums are taken lane-strided, then by xor-shuffle: a fixed order, different from the 128-thread kernel's.)
__global__ void __launch_bounds__(256) expm_classify_warp_kernel(ExpmClassifyArgs a, int n) {
    const int lane = threadIdx.x & 31, k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= n) return;
    const int slot = a.dst[k];
    if (slot < 0) { if (lane == 0) { a.flag[k] = 0; a.weight[k] = 0.0; } return; }
    const double *v = a.V + (size_t)k * a.nV;
    const int g = a.group[k], r = a.ref[k];
    const double *vr = nullptr;
    double wr_cached = -1.0;
    if (g >= 0) {
        if (r >= 0) vr = a.V + (size_t)r * a.nV;
        else if (a.groups[g].kind == a.kind) { vr = a.refvec + (size_t)g * a.refvec_stride; wr_cached = a.groups[g].weight; }
    }
    double wk = 0.0, wr = 0.0;
    for (int f = lane; f < a.nV; f += 32) {
        const double x = v[f];
        wk += fabs(x);
        if (vr) wr += fabs(vr[f]);
        if (a.res) a.res[(size_t)slot * a.nV + f] = x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { wk += __shfl_xor_sync(0xffffffffu, wk, o); wr += __shfl_xor_sync(0xffffffffu, wr, o); }
    if (wr_cached >= 0.0) wr = wr_cached;
    int ok = (vr != nullptr) && (wr > 0.0) && (wk >= 0.0) && (wk < INFINITY) && (wr < INFINITY);   // NaN fails every comparison
    if (ok) {
        const double tol = 1e-13 * wk * wr;
        for (int f = lane; f < a.nV; f += 32)
            if (!(fabs(v[f] * wr - vr[f] * wk) <= tol)) ok = 0;
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) { a.flag[k] = ok; a.weight[k] = wk; }
}

__global__ void __launch_bounds__(256, 2) expm64_dmma_kernel(ExpmArgs a, ExpmTcOut tc) {
    // Three shared 64x64 fp64 buffers (101 KB) so that TWO CTAs share an SM and one CTA's load/norm/epilogue phases
    // overlap the other's DMMA products:  X0 = A -> later R,  X1 = A^2,  X2 = A^3.  The polynomial blocks need A at the
    // thread's own output positions after X0 has become R: those values are parked in the (not yet written) output
    // slot PT[slot] in global memory (L2 resident, same [row][col] layout) and re-read per Horner step.
    extern __shared__ __align__(16) double sm[];
    double *X0 = sm, *X1 = X0 + 64 * LD64, *X2 = X1 + 64 * LD64;
    __shared__ double red[64];
    __shared__ int s_shift, s_Q;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = warp >> 1, wc = warp & 1, g = lane >> 2, q4 = lane & 3;
    const int D = a.D;
    if (a.dst[blockIdx.x] < 0) return;
    const size_t slot = a.dst[blockIdx.x];
    double *out = a.PT + slot * 4096;
    double *park = out;
    double *R = X0;

    const bool shared_powers = a.flag && a.flag[blockIdx.x];
    if (shared_powers) {
        // expm_poly_kernel has left the Taylor polynomial of A / 2^shift in the output slot; squarings and the epilogue
        // (clamp, row repair, tensor-path operands) remain.  shift is recomputed exactly as the polynomial stage chose it.
        const ExpmGroup gi = a.groups[a.group[blockIdx.x]];
        const double rho = gi.nu * (a.weight[blockIdx.x] / gi.weight);
        int shift = 0;
        if (rho > EXPM_POLY_THETA) { int e = 0; frexp(rho / EXPM_POLY_THETA, &e); shift = max(e, 0); }
        if (shift == 0 && rho == rho) {          // finished by expm_poly_kernel up to the row repair (expm_diag_kernel's arithmetic)
            if (a.colsum && tid < 64) {
                const int col = tid;
                double s = 0.0;
#pragma unroll
                for (int sl = 0; sl < 16; sl++) s += a.colsum[((size_t)blockIdx.x * 16 + sl) * 64 + col];
                const double d = col < a.D ? fmax(1.0 - s, 0.0) : 0.0;
                out[col * 64 + col] = d;
                if (tc.PB) {
                    tc.PTf[slot * 4352 + col * 68 + col] = (float)d;
                    const float hi = tf32_rn_dev((float)d);
                    float *pb = tc.PB + slot * 8192 + (col >> 2) * 256 + col * 4 + (col & 3);
                    pb[0] = hi;
                    pb[4096] = tf32_rn_dev((float)(d - (double)hi));
                }
            }
            return;
        }
        if (!(rho == rho) || shift > 900) {
            for (int idx = tid; idx < 4096; idx += 256) out[idx] = __longlong_as_double(0x7ff8000000000000LL);
            if (tc.PB) {
                for (int idx = tid; idx < 8192; idx += 256) tc.PB[slot * 8192 + idx] = __int_as_float(0x7fc00000);
                for (int idx = tid; idx < 4352; idx += 256) tc.PTf[slot * 4352 + idx] = __int_as_float(0x7fc00000);
            }
            return;
        }
        for (int idx = tid; idx < 4096; idx += 256) R[(idx >> 6) * LD64 + (idx & 63)] = __ldcg(out + idx);
        if (tid == 0) s_shift = shift;
        __syncthreads();
    } else
    load_rate_matrix<64, LD64, 256>(a, X0, tid);
    if (shared_powers || !a.is_trans) {
      double acc[2][4][2];
      auto zero = [&]() {
#pragma unroll
          for (int i = 0; i < 2; i++)
#pragma unroll
              for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
      };
      auto off = [&](int i, int j) { return (16 * wr + 8 * i + g) * LD64 + 32 * wc + 8 * j + 2 * q4; };
      auto goff = [&](int i, int j) { return (16 * wr + 8 * i + g) * 64 + 32 * wc + 8 * j + 2 * q4; };
      auto store = [&](double *M) {
#pragma unroll
          for (int i = 0; i < 2; i++)
#pragma unroll
              for (int j = 0; j < 4; j++) *reinterpret_cast<double2 *>(M + off(i, j)) = make_double2(acc[i][j][0], acc[i][j][1]);
      };
      if (!shared_powers) {
        if (tid < 64) {
            double s = 0.0;
            for (int j = 0; j < 64; j++) s += fabs(X0[j * LD64 + tid]);
            red[tid] = s;
        }
        __syncthreads();
        if (tid < 32) {                      // warp-parallel max / NaN scan of the 64 row sums
            double m = fmax(red[tid], red[tid + 32]);
            bool bad = !(red[tid] == red[tid]) || !(red[tid + 32] == red[tid + 32]) || isinf(red[tid]) || isinf(red[tid + 32]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
            bad = __any_sync(0xffffffffu, bad);
            if (tid == 0) {
                int shift = 0;
                if (!bad && m > 0.975) {
                    int e = 0;
                    frexp(m / 0.975, &e);          // m/0.975 = f*2^e, f in [0.5,1)  ->  m/2^e <= 0.975
                    shift = max(e, 0);
                }
                const double theta = bad ? 0.0 : ldexp(m, -shift);
                s_Q = theta <= 0.0064 ? 1 : theta <= 0.069 ? 2 : theta <= 0.245 ? 3 : theta <= 0.55 ? 4 : 5;
                s_shift = bad ? -1 : shift;
            }
        }
        __syncthreads();
        const int shift = s_shift, Q = s_Q;
        if (shift < 0 || shift > 900) {           // NaN/inf or absurd rates: propagate NaN (host sees NaN lnL)
            for (int idx = tid; idx < 4096; idx += 256) out[idx] = __longlong_as_double(0x7ff8000000000000LL);
            if (tc.PB) {
                for (int idx = tid; idx < 8192; idx += 256) tc.PB[slot * 8192 + idx] = __int_as_float(0x7fc00000);
                for (int idx = tid; idx < 4352; idx += 256) tc.PTf[slot * 4352 + idx] = __int_as_float(0x7fc00000);
            }
            return;
        }
        {   // scale A in place and park a copy in the output slot (coalesced)
            const double sc = exp2i(-shift);
            for (int idx = tid; idx < 64 * 64; idx += 256) {
                const int o = (idx >> 6) * LD64 + (idx & 63);
                const double v = X0[o] * sc;
                X0[o] = v;
                park[idx] = v;
            }
            __syncthreads();
        }
        zero(); tile_mm64_dmma(X0, X0, wr, wc, g, q4, acc); store(X1);
        __syncthreads();
        zero(); tile_mm64_dmma(X1, X0, wr, wc, g, q4, acc); store(X2);
        __syncthreads();                             // every warp is done reading A from X0: it becomes R
        // R = B_Q, then R = R*A3 + B_q for q = Q-1..0, with B_q = c_{3q} I + c_{3q+1} A + c_{3q+2} A^2 at own positions
        auto poly2 = [&](int qq, int i, int j) -> double2 {
            const int r = 16 * wr + 8 * i + g, c = 32 * wc + 8 * j + 2 * q4;
            const double2 a1 = __ldcg(reinterpret_cast<const double2 *>(park + goff(i, j)));
            const double2 a2 = *reinterpret_cast<const double2 *>(X1 + off(i, j));
            double2 v = make_double2(c_taylor18[3 * qq + 1] * a1.x + c_taylor18[3 * qq + 2] * a2.x,
                                     c_taylor18[3 * qq + 1] * a1.y + c_taylor18[3 * qq + 2] * a2.y);
            if (r == c) v.x += c_taylor18[3 * qq];
            if (r == c + 1) v.y += c_taylor18[3 * qq];
            return v;
        };
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) *reinterpret_cast<double2 *>(R + off(i, j)) = poly2(Q, i, j);
        __syncthreads();
        for (int qq = Q - 1; qq >= 0; qq--) {
            zero(); tile_mm64_dmma(R, X2, wr, wc, g, q4, acc);
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double2 pv = poly2(qq, i, j);
                    *reinterpret_cast<double2 *>(R + off(i, j)) = make_double2(acc[i][j][0] + pv.x, acc[i][j][1] + pv.y);
                }
            __syncthreads();
        }
      }   // general path: R = Taylor polynomial of A / 2^shift
        const int n_sq = s_shift;
        for (int s = 0; s < n_sq; s++) {
            zero(); tile_mm64_dmma(R, R, wr, wc, g, q4, acc);
            __syncthreads();
            store(R);
            __syncthreads();
        }
        // clamp negatives, zero the padding, repair: column k of PT (= row k of P) must sum to 1
        for (int idx = tid; idx < 64 * 64; idx += 256) {
            int j = idx >> 6, kk = idx & 63;
            double v = R[j * LD64 + kk];
            if (j >= D || kk >= D) v = 0.0; else if (v < 0.0) v = 0.0;
            R[j * LD64 + kk] = v;
        }
        __syncthreads();
        {   // column sums by all 256 threads: 4 threads per column, 16 rows each
            const int col = tid & 63, part = tid >> 6;
            double s = 0.0;
            for (int j = part * 16; j < part * 16 + 16; j++) if (j != col) s += R[j * LD64 + col];
            X1[part * 64 + col] = s;                 // X1 (A^2) is no longer needed
        }
        __syncthreads();
        if (tid < D) R[tid * LD64 + tid] = fmax(1.0 - (X1[tid] + X1[64 + tid] + X1[128 + tid] + X1[192 + tid]), 0.0);
        __syncthreads();
    }
    {
        for (int idx = tid; idx < 4096; idx += 256) out[idx] = R[(idx >> 6) * LD64 + (idx & 63)];
        if (tc.PB) {
            float *pb = tc.PB + slot * 8192;
            float *pf = tc.PTf + slot * 4352;           // rows padded to 68 floats (TC_PTF_ROW)
            for (int o = tid; o < 4096; o += 256) {
                const int chunk = o >> 8, n = (o >> 2) & 63, kk = chunk * 4 + (o & 3);
                const double pv = R[kk * LD64 + n];          // P[n][kk] = PT[kk][n]
                const float hi = tf32_rn_dev((float)pv);
                pb[o] = hi;
                pb[4096 + o] = tf32_rn_dev((float)(pv - (double)hi));
                pf[(o >> 6) * 68 + (o & 63)] = (float)R[(o >> 6) * LD64 + (o & 63)];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Shared-powers path, stage 2: S_m = R^m / m! (PT layout, i.e. powers of the TRANSPOSED reference matrix), m = 1..24, for
// every group whose reference is (re)established in this batch.  grid = (24, number of such groups); CTA (j, g) produces
// S_{j+1}:  j = 0 assembles the reference matrix, normalises it (R = A_ref / ||A_ref||_inf -> S_1) and records the
// group's measures and direction; CTA j >= 1 multiplies two EARLIER coefficient matrices,
//     S_m = S_a S_b / C(m, a),   (a, b) = (1,1) (2,1) (2,2) (4,1..4) (8,1..8) (16,1..8),
// which it awaits on generation flags (flag[group][m] == gen once S_m is complete; the producer fences before it
// publishes).  Blocks are dispatched in index order and only wait for lower indices of their own column, so the scheme
// cannot deadlock even if not all of them are co-resident; waits are time-bounded like every device-side wait here.
// Critical path: 5 products instead of 23.
// ------------------------------------------------------------------------------------------------
struct ExpmPowersArgs {
    ExpmArgs a;               // input description (dense or compiled); a.Qres must be null
    const int *refs;          // [gridDim.y] batch indices of the reference entries
    ExpmGroup *groups;        // out
    double *pow;              // out [G][EXPM_POW_TERMS][4096]
    double *refvec;           // out [G][refvec_stride]
    int refvec_stride, nV, kind;
    const double *Vin;        // the raw input vectors [n][nV] (formula values or dense entries)
    unsigned long long *flags;   // [G][EXPM_POW_TERMS] generation of each coefficient matrix
    unsigned long long gen;
    int *err;
};

__device__ __forceinline__ bool expm_await_flag(const unsigned long long *f, unsigned long long gen, int *err) {
    unsigned long long t0 = 0;
    for (int it = 0;; it++) {
        unsigned long long v;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        if (v == gen) return true;
        if ((it & 255) == 255) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) { atomicExch(err, 4); return false; }
        }
    }
}

__global__ void __launch_bounds__(256, 1) expm_powers_kernel(ExpmPowersArgs pa) {
    extern __shared__ __align__(16) double sm[];
    double *X0 = sm, *X1 = X0 + 64 * LD64;
    __shared__ double red[64];
    __shared__ double s_nu;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = warp >> 1, wc = warp & 1, g = lane >> 2, q4 = lane & 3;
    const int entry = pa.refs[blockIdx.y];
    const int grp = pa.a.group[entry];
    if (grp < 0 || pa.a.dst[entry] < 0) return;                       // (cannot happen: the host lists live references only)
    const int m = blockIdx.x + 1;                                      // this CTA produces S_m
    double *pw = pa.pow + (size_t)grp * EXPM_POW_TERMS * 4096;
    unsigned long long *flags = pa.flags + (size_t)grp * EXPM_POW_TERMS;
    if (m == 1) {
        load_rate_matrix<64, LD64, 256>(pa.a, X0, tid, entry);        // X0 = A_ref^T, zero padded
        if (tid < 64) {
            double s = 0.0;
            for (int j = 0; j < 64; j++) s += fabs(X0[j * LD64 + tid]);   // column sums of A^T = row sums of A
            red[tid] = s;
        }
        __syncthreads();
        if (tid == 0) {
            double mx = 0.0;
            bool bad = false;
            for (int i = 0; i < 64; i++) { if (!(red[i] == red[i]) || isinf(red[i])) bad = true; mx = fmax(mx, red[i]); }
            s_nu = (bad || !(mx > 0.0)) ? 0.0 : mx;
        }
        __syncthreads();
        const double nu = s_nu;
        if (tid == 0) {
            ExpmGroup gi;
            gi.nu = nu; gi.weight = pa.a.weight[entry]; gi.kind = nu > 0.0 ? pa.kind : 0; gi.pad = 0;
            pa.groups[grp] = gi;
        }
        for (int f = tid; f < pa.nV; f += 256) pa.refvec[(size_t)grp * pa.refvec_stride + f] = pa.Vin[(size_t)entry * pa.nV + f];
        // (no entry can have been flagged against a reference with nu == 0; the other CTAs of the column still need S_1)
        const double inv = nu > 0.0 ? 1.0 / nu : 0.0;
        for (int idx = tid; idx < 4096; idx += 256) pw[4096 + idx] = X0[(idx >> 6) * LD64 + (idx & 63)] * inv;
        __threadfence();
        __syncthreads();
        if (tid == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flags + 1), "l"(pa.gen) : "memory");
        return;
    }
    const int a = m <= 2 ? 1 : m <= 4 ? 2 : m <= 8 ? 4 : m <= 16 ? 8 : 16, b = m - a;
    if (tid == 0) expm_await_flag(flags + a, pa.gen, pa.err);
    if (tid == 32) expm_await_flag(flags + b, pa.gen, pa.err);
    __syncthreads();
    for (int idx = tid; idx < 2048; idx += 256) {                     // both operands as double2, coalesced
        const int r = idx >> 5, c2 = (idx & 31) * 2;
        *reinterpret_cast<double2 *>(X0 + r * LD64 + c2) = __ldcg(reinterpret_cast<const double2 *>(pw + (size_t)a * 4096 + r * 64 + c2));
        *reinterpret_cast<double2 *>(X1 + r * LD64 + c2) = __ldcg(reinterpret_cast<const double2 *>(pw + (size_t)b * 4096 + r * 64 + c2));
    }
    __syncthreads();
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    tile_mm64_dmma(X0, X1, wr, wc, g, q4, acc);                       // S_a * S_b (powers of one matrix commute)
    double binom = 1.0;                                                // C(m, a), exact in fp64 for m <= 24
    for (int i = 1; i <= a; i++) binom = binom * (double)(m - a + i) / (double)i;
    const double inv = 1.0 / binom;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = 16 * wr + 8 * i + g, c = 32 * wc + 8 * j + 2 * q4;
            __stcg(reinterpret_cast<double2 *>(pw + (size_t)m * 4096 + r * 64 + c), make_double2(acc[i][j][0] * inv, acc[i][j][1] * inv));
        }
    __threadfence();
    __syncthreads();
    if (tid == 0) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flags + m), "l"(pa.gen) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Shared-powers path, stage 3: P = sum_m S_m x^m with x = rho / 2^shift <= EXPM_POLY_THETA for every flagged entry, fused
// with the epilogue for the (common) entries that need no squaring.
// grid = (16 row slices, chunks of EXPM_POLY_CHUNK entries).  Thread t of slice sl owns element (row 4 sl + t%4,
// column t/4) of the PT layout (row = child state, column = parent state) and keeps that element's 24 coefficients in
// registers while it walks its chunk (reloaded only when the group changes: batches are class-major); the chunk's
// metadata is fetched up front so that the loop carries no dependent global loads.
//   shift == 0: the value is final up to the row repair.  It is clamped at 0, padding is zeroed, and PT (fp64), PTf (fp32
//               table) and PB (tf32 hi/lo tiles; chunk sl of the canonical layout is exactly this slice, offset t) are
//               written directly; the partial column sums of the slice (4 rows, diagonal excluded; 4 adjacent lanes) go to
//               colsum[entry][sl][column] for expm_diag_kernel, which sets the diagonal to 1 - (sum of the rest of P's row).
//   shift  > 0: the raw polynomial value goes to PT; expm64_dmma_kernel squares and finishes the entry.
// ------------------------------------------------------------------------------------------------
constexpr int EXPM_POLY_CHUNK = 8;
__global__ void __launch_bounds__(256, 3) expm_poly_kernel(ExpmArgs a, ExpmTcOut tc, double *__restrict__ colsum, int n) {
    __shared__ int s_slot[EXPM_POLY_CHUNK], s_grp[EXPM_POLY_CHUNK];
    __shared__ double s_wgt[EXPM_POLY_CHUNK];
    const int tid = threadIdx.x, sl = blockIdx.x;
    const int row = 4 * sl + (tid & 3), col = tid >> 2;
    const int e = row * 64 + col;                                     // element index in the 64x64 PT layout
    const double unit = (row == col) ? 1.0 : 0.0;
    const bool pad = row >= a.D || col >= a.D;
    const int k0 = blockIdx.y * EXPM_POLY_CHUNK;
    if (tid < EXPM_POLY_CHUNK) {                                      // the chunk's metadata, once per CTA
        const int k = k0 + tid;
        int sv = -1, gv = -1;
        double wv = 0.0;
        if (k < n) {
            sv = a.dst[k];
            if (sv >= 0 && !a.flag[k]) sv = -1;
            gv = a.group[k];
            wv = a.weight[k];
        }
        s_slot[tid] = sv; s_grp[tid] = gv; s_wgt[tid] = wv;
    }
    __syncthreads();
    double c[EXPM_POW_TERMS];
    int cur = -1;
    double nu = 0.0, wref = 1.0;
    const size_t out_pt = (size_t)e, out_ptf = (size_t)row * 68 + col, out_pb = (size_t)sl * 256 + tid;
    for (int i = 0; i < EXPM_POLY_CHUNK; i++) {
        const int slot = s_slot[i];
        if (slot < 0) continue;
        if (s_grp[i] != cur) {
            cur = s_grp[i];
            const double *pw = a.pow + (size_t)cur * EXPM_POW_TERMS * 4096 + e;
#pragma unroll
            for (int m = 1; m < EXPM_POW_TERMS; m++) c[m] = __ldcg(pw + (size_t)m * 4096);
            nu = a.groups[cur].nu; wref = a.groups[cur].weight;
        }
        const double rho = nu * (s_wgt[i] / wref);
        double x = rho;
        int shift = 0;
        if (rho > EXPM_POLY_THETA) {
            frexp(rho / EXPM_POLY_THETA, &shift);
            if (shift > 900) continue;                                // the finishing kernel writes NaN for absurd rates
            shift = max(shift, 0);
            x = ldexp(rho, -shift);
        }
        // p(x) = 1 + x (c1 + c3 x^2 + .. + c23 x^22) + x^2 (c2 + c4 x^2 + .. + c24 x^22): two interleaved Horner chains in
        // x^2 halve the dependent-FMA depth
        static_assert(EXPM_POW_TERMS == 25, "the even/odd split below assumes degree 24");
        const double x2 = x * x;
        double ve = c[24], vo = c[23];
#pragma unroll
        for (int m = 22; m >= 2; m -= 2) { ve = fma(ve, x2, c[m]); vo = fma(vo, x2, c[m - 1]); }
        double v = fma(vo, x, fma(ve, x2, unit));
        const size_t sidx = (size_t)slot;
        if (shift > 0) { __stcg(a.PT + sidx * 4096 + out_pt, v); continue; }
        v = (pad || v < 0.0) ? 0.0 : v;
        // column sums over this slice's 4 rows (lanes t, t^1, t^2, t^3 share a column), diagonal element excluded
        double part = (row == col) ? 0.0 : v;
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        if ((tid & 3) == 0) colsum[((size_t)(k0 + i) * 16 + sl) * 64 + col] = part;
        if (row != col) {                                             // expm_diag_kernel owns the diagonal
            __stcg(a.PT + sidx * 4096 + out_pt, v);
            if (tc.PB) {
                const float vf = (float)v;
                tc.PTf[sidx * 4352 + out_ptf] = vf;
                const float hi = tf32_rn_dev(vf);
                float *pb = tc.PB + sidx * 8192 + out_pb;             // canonical tile: chunk = row/4 = sl, offset col*4 + row%4 = tid
                pb[0] = hi;
                pb[4096] = tf32_rn_dev((float)(v - (double)hi));
            }
        }
    }
}

// Row repair of the entries expm_poly_kernel finished (flagged, no squaring): P[k][k] = max(1 - sum_{j != k} P[k][j], 0),
// i.e. PT[k][k] from the column sums of PT, added up in slice order (deterministic).  One CTA of 64 threads per entry.
__global__ void __launch_bounds__(64) expm_diag_kernel(ExpmArgs a, ExpmTcOut tc, const double *__restrict__ colsum) {
    const int k = blockIdx.x, col = threadIdx.x;
    const int slot = a.dst[k];
    if (slot < 0 || !a.flag[k]) return;
    const ExpmGroup gi = a.groups[a.group[k]];
    const double rho = gi.nu * (a.weight[k] / gi.weight);
    if (!(rho <= EXPM_POLY_THETA)) return;                            // squared and finished by expm64_dmma_kernel
    double s = 0.0;
#pragma unroll
    for (int sl = 0; sl < 16; sl++) s += colsum[((size_t)k * 16 + sl) * 64 + col];
    const double d = col < a.D ? fmax(1.0 - s, 0.0) : 0.0;
    const size_t sidx = (size_t)slot;
    a.PT[sidx * 4096 + col * 64 + col] = d;
    if (tc.PB) {
        tc.PTf[sidx * 4352 + col * 68 + col] = (float)d;
        const float hi = tf32_rn_dev((float)d);
        float *pb = tc.PB + sidx * 8192 + (col >> 2) * 256 + col * 4 + (col & 3);
        pb[0] = hi;
        pb[4096] = tf32_rn_dev((float)(d - (double)hi));
    }
}

// Small state spaces (Dp <= 32): one CTA of 128 threads per matrix, plain shared-memory products.
constexpr size_t expm_small_smem_bytes(int DP) { return (size_t)6 * DP * (DP + 1) * sizeof(double); }
template <int DP>
__global__ void __launch_bounds__(128) expm_small_kernel(ExpmArgs a) {
    constexpr int LD = DP + 1;
    extern __shared__ __align__(16) double sm[];          // 6 matrices of DP*LD doubles (expm_small_smem_bytes)
    double *A1 = sm, *A2 = A1 + DP * LD, *A3 = A2 + DP * LD, *A4 = A3 + DP * LD, *R = A4 + DP * LD, *T = R + DP * LD;
    __shared__ double red[DP];
    __shared__ int s_shift;
    const int tid = threadIdx.x, D = a.D;
    if (a.dst[blockIdx.x] < 0) return;
    double *out = a.PT + (size_t)a.dst[blockIdx.x] * DP * DP;
    load_rate_matrix<DP, LD, 128>(a, A1, tid);
    if (a.is_trans) {
        for (int idx = tid; idx < DP * DP; idx += 128) out[idx] = A1[(idx / DP) * LD + idx % DP];
        return;
    }
    if (tid < DP) {
        double s = 0.0;
        for (int j = 0; j < DP; j++) s += fabs(A1[j * LD + tid]);
        red[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        double m = 0.0; bool bad = false;
        for (int i = 0; i < DP; i++) { if (!(red[i] == red[i]) || isinf(red[i])) bad = true; m = fmax(m, red[i]); }
        int e = 0;
        if (m > 0.0) frexp(m, &e);
        s_shift = bad ? -1 : max(e, 0);
    }
    __syncthreads();
    const int shift = s_shift;
    if (shift < 0 || shift > 900) {
        for (int idx = tid; idx < DP * DP; idx += 128) out[idx] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    if (shift > 0) {
        const double sc = exp2i(-shift);
        for (int idx = tid; idx < DP * DP; idx += 128) A1[(idx / DP) * LD + idx % DP] *= sc;
        __syncthreads();
    }
    auto mm = [&](double *Cm, const double *X, const double *Y) {      // Cm = X*Y (Cm distinct from X,Y)
        for (int idx = tid; idx < DP * DP; idx += 128) {
            int r = idx / DP, c = idx % DP;
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < DP; k++) s = fma(X[r * LD + k], Y[k * LD + c], s);
            Cm[r * LD + c] = s;
        }
        __syncthreads();
    };
    mm(A2, A1, A1); mm(A3, A2, A1); mm(A4, A2, A2);
    auto poly = [&](int q, int r, int c) -> double {
        int o = r * LD + c;
        double v = taylor_c(4 * q + 1) * A1[o] + taylor_c(4 * q + 2) * A2[o] + taylor_c(4 * q + 3) * A3[o];
        if (r == c) v += taylor_c(4 * q);
        return v;
    };
    for (int idx = tid; idx < DP * DP; idx += 128) R[(idx / DP) * LD + idx % DP] = poly(3, idx / DP, idx % DP);
    __syncthreads();
    for (int q = 2; q >= 0; q--) {
        mm(T, R, A4);
        for (int idx = tid; idx < DP * DP; idx += 128) { int r = idx / DP, c = idx % DP; R[r * LD + c] = T[r * LD + c] + poly(q, r, c); }
        __syncthreads();
    }
    for (int s = 0; s < shift; s++) {
        mm(T, R, R);
        for (int idx = tid; idx < DP * DP; idx += 128) { int o = (idx / DP) * LD + idx % DP; R[o] = T[o]; }
        __syncthreads();
    }
    for (int idx = tid; idx < DP * DP; idx += 128) {
        int j = idx / DP, k = idx % DP;
        double v = R[j * LD + k];
        if (j >= D || k >= D) v = 0.0; else if (v < 0.0) v = 0.0;
        R[j * LD + k] = v;
    }
    __syncthreads();
    if (tid < DP) {
        double s = 0.0;
        for (int j = 0; j < DP; j++) if (j != tid) s += R[j * LD + tid];
        if (tid < D) R[tid * LD + tid] = fmax(1.0 - s, 0.0);
    }
    __syncthreads();
    for (int idx = tid; idx < DP * DP; idx += 128) out[idx] = R[(idx / DP) * LD + idx % DP];
}

// ------------------------------------------------------------------------------------------------
// Fused pruning update, 33..64 states (codon models), fp64.
// grid = (Sp/64, jobs, classes); block = 256 threads as a 16x16 grid, thread (tx,ty) owns patterns 4ty..4ty+3
// and parent states 4tx..4tx+3 of the tile.  jobs[blockIdx.y] = internal index of the parent to (re)compute.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) prune64_kernel(PruneArgs a, const int *__restrict__ jobs) {
    extern __shared__ __align__(16) double sm[];
    double *Xs = sm;                    // [64][LD64] child conditionals (pattern-major)
    double *Ps = sm + 64 * LD64;        // [64][LD64] PT of the child's branch
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int par = jobs[blockIdx.y];
    const int cat = a.cat0 + blockIdx.z;
    const int s0 = blockIdx.x * TILE_P;
    const size_t Sp = a.Sp;
    double v[4][4];
    int ex[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        ex[i] = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) v[i][j] = 1.0;
    }
    const int c_begin = a.tree.child_start[par], c_end = a.tree.child_start[par + 1];
    for (int ci = c_begin; ci < c_end; ci++) {
        const int child = a.tree.child_ids[ci];
        const double *PT = a.PT + ((size_t)cat * a.B + child) * 4096;
        if (child < a.L) {
            // leaf: column gather PT[state][k]; ambiguous: sum_j amb[j] PT[j][k]
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int code = (child == a.forced_node) ? a.forced[s0 + 4 * ty + i] : a.leaf[(size_t)child * Sp + s0 + 4 * ty + i];
                double m0, m1, m2, m3;
                if (code >= 0) {
                    const double2 *p = reinterpret_cast<const double2 *>(PT + (size_t)code * 64 + 2 * tx);
                    double2 q0 = __ldg(p), q1 = __ldg(p + 16);
                    m0 = q0.x; m1 = q0.y; m2 = q1.x; m3 = q1.y;
                } else {
                    const double *amb = a.ambig + (size_t)(-code - 1) * 64;
                    m0 = m1 = m2 = m3 = 0.0;
                    for (int j = 0; j < a.D; j++) {
                        const double w = __ldg(amb + j);
                        if (w != 0.0) {
                            const double2 *p = reinterpret_cast<const double2 *>(PT + (size_t)j * 64 + 2 * tx);
                            double2 q0 = __ldg(p), q1 = __ldg(p + 16);
                            m0 = fma(w, q0.x, m0); m1 = fma(w, q0.y, m1); m2 = fma(w, q1.x, m2); m3 = fma(w, q1.y, m3);
                        }
                    }
                }
                v[i][0] *= m0; v[i][1] *= m1; v[i][2] *= m2; v[i][3] *= m3;
            }
        } else {
            const int cin = child - a.L;
            const double *X = a.cond + (((size_t)cat * a.I + cin) * Sp + s0) * 64;
            __syncthreads();            // previous child's tiles fully consumed
            for (int idx = tid; idx < 2048; idx += 256) {       // 4096 doubles as double2, coalesced
                const int r = idx >> 5, c2 = (idx & 31) * 2;
                *reinterpret_cast<double2 *>(Xs + r * LD64 + c2) = *reinterpret_cast<const double2 *>(X + r * 64 + c2);
                *reinterpret_cast<double2 *>(Ps + r * LD64 + c2) = __ldg(reinterpret_cast<const double2 *>(PT + r * 64 + c2));
            }
            __syncthreads();
            double acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
            tile_mm64(Xs, Ps, tx, ty, acc);
            const int *sc = a.scal + ((size_t)cat * a.I + cin) * Sp + s0 + 4 * ty;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                ex[i] += sc[i];
#pragma unroll
                for (int j = 0; j < 4; j++) v[i][j] *= acc[i][j];
            }
        }
    }
    if (a.L + par == a.forced_node) {            // pinned internal node: only the forced state survives
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int f = a.forced[s0 + 4 * ty + i];
#pragma unroll
            for (int j = 0; j < 4; j++) if (col_of(tx, j) != f) v[i][j] = 0.0;
        }
    }
    // per-pattern renormalisation: exact power of two so that max_k lands in [0.5, 1)
    const bool is_root = (par == a.I - 1);
    double *outp = a.cond + (((size_t)cat * a.I + par) * Sp + s0) * 64;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double m = fmax(fmax(v[i][0], v[i][1]), fmax(v[i][2], v[i][3]));
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        int e = 0;
        if (m > 0.0 && m < INFINITY) {
            e = ilogb(m) + 1;
            const double s1 = exp2i(-(e / 2)), s2 = exp2i(-(e - e / 2));   // two steps: |e| may exceed 1022
#pragma unroll
            for (int j = 0; j < 4; j++) v[i][j] = v[i][j] * s1 * s2;
        }
        ex[i] += e;
        const int row = 4 * ty + i;
        *reinterpret_cast<double2 *>(outp + (size_t)row * 64 + 2 * tx) = make_double2(v[i][0], v[i][1]);
        *reinterpret_cast<double2 *>(outp + (size_t)row * 64 + 32 + 2 * tx) = make_double2(v[i][2], v[i][3]);
        if (tx == 0) a.scal[((size_t)cat * a.I + par) * Sp + s0 + row] = ex[i];
        if (is_root) {
            double r = v[i][0] * a.pi[2 * tx] + v[i][1] * a.pi[2 * tx + 1] + v[i][2] * a.pi[32 + 2 * tx] + v[i][3] * a.pi[33 + 2 * tx];
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
            if (tx == 0) {
                a.rootL[(size_t)cat * Sp + s0 + row] = r;
                a.rootE[(size_t)cat * Sp + s0 + row] = ex[i];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 33..64 states, fp64, WHOLE pruning pass in one launch: patterns are independent, so tile t of a node needs tile t of its
// children only -- CTA (tile, class) walks the dirty internal nodes in post-order (jobs ascending) and every dependency is
// CTA-local (a child's tile was written by this CTA earlier in the launch, or is resident from an earlier evaluation).  No
// grid-wide synchronisation, no launch per tree level (the north-star tree has 48 levels: 54 launches per evaluation
// before).  Same tile body as prune64_kernel.  grid = (Sp/64, classes).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) prune64_walk_kernel(PruneArgs a, const int *__restrict__ jobs, int njobs) {
    extern __shared__ __align__(16) double sm[];
    double *Xs = sm;                    // [64][LD64] child conditionals (pattern-major)
    double *Ps = sm + 64 * LD64;        // [64][LD64] PT of the child's branch
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int cat = a.cat0 + blockIdx.y;
    const int s0 = blockIdx.x * TILE_P;
    const size_t Sp = a.Sp;
  for (int jb = 0; jb < njobs; jb++) {
    const int par = __ldg(jobs + jb);
    double v[4][4];
    int ex[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        ex[i] = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) v[i][j] = 1.0;
    }
    const int c_begin = a.tree.child_start[par], c_end = a.tree.child_start[par + 1];
    for (int ci = c_begin; ci < c_end; ci++) {
        const int child = a.tree.child_ids[ci];
        const double *PT = a.PT + ((size_t)cat * a.B + child) * 4096;
        if (child < a.L) {
            // leaf: column gather PT[state][k]; ambiguous: sum_j amb[j] PT[j][k]
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int code = (child == a.forced_node) ? a.forced[s0 + 4 * ty + i] : a.leaf[(size_t)child * Sp + s0 + 4 * ty + i];
                double m0, m1, m2, m3;
                if (code >= 0) {
                    const double2 *p = reinterpret_cast<const double2 *>(PT + (size_t)code * 64 + 2 * tx);
                    double2 q0 = __ldg(p), q1 = __ldg(p + 16);
                    m0 = q0.x; m1 = q0.y; m2 = q1.x; m3 = q1.y;
                } else {
                    const double *amb = a.ambig + (size_t)(-code - 1) * 64;
                    m0 = m1 = m2 = m3 = 0.0;
                    for (int j = 0; j < a.D; j++) {
                        const double w = __ldg(amb + j);
                        if (w != 0.0) {
                            const double2 *p = reinterpret_cast<const double2 *>(PT + (size_t)j * 64 + 2 * tx);
                            double2 q0 = __ldg(p), q1 = __ldg(p + 16);
                            m0 = fma(w, q0.x, m0); m1 = fma(w, q0.y, m1); m2 = fma(w, q1.x, m2); m3 = fma(w, q1.y, m3);
                        }
                    }
                }
                v[i][0] *= m0; v[i][1] *= m1; v[i][2] *= m2; v[i][3] *= m3;
            }
        } else {
            const int cin = child - a.L;
            const double *X = a.cond + (((size_t)cat * a.I + cin) * Sp + s0) * 64;
            __syncthreads();            // previous child's tiles fully consumed
            for (int idx = tid; idx < 2048; idx += 256) {       // 4096 doubles as double2, coalesced
                const int r = idx >> 5, c2 = (idx & 31) * 2;
                *reinterpret_cast<double2 *>(Xs + r * LD64 + c2) = *reinterpret_cast<const double2 *>(X + r * 64 + c2);
                *reinterpret_cast<double2 *>(Ps + r * LD64 + c2) = __ldg(reinterpret_cast<const double2 *>(PT + r * 64 + c2));
            }
            __syncthreads();
            double acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
            tile_mm64(Xs, Ps, tx, ty, acc);
            const int *sc = a.scal + ((size_t)cat * a.I + cin) * Sp + s0 + 4 * ty;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                ex[i] += sc[i];
#pragma unroll
                for (int j = 0; j < 4; j++) v[i][j] *= acc[i][j];
            }
        }
    }
    if (a.L + par == a.forced_node) {            // pinned internal node: only the forced state survives
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int f = a.forced[s0 + 4 * ty + i];
#pragma unroll
            for (int j = 0; j < 4; j++) if (col_of(tx, j) != f) v[i][j] = 0.0;
        }
    }
    // per-pattern renormalisation: exact power of two so that max_k lands in [0.5, 1)
    const bool is_root = (par == a.I - 1);
    double *outp = a.cond + (((size_t)cat * a.I + par) * Sp + s0) * 64;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double m = fmax(fmax(v[i][0], v[i][1]), fmax(v[i][2], v[i][3]));
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        int e = 0;
        if (m > 0.0 && m < INFINITY) {
            e = ilogb(m) + 1;
            const double s1 = exp2i(-(e / 2)), s2 = exp2i(-(e - e / 2));   // two steps: |e| may exceed 1022
#pragma unroll
            for (int j = 0; j < 4; j++) v[i][j] = v[i][j] * s1 * s2;
        }
        ex[i] += e;
        const int row = 4 * ty + i;
        *reinterpret_cast<double2 *>(outp + (size_t)row * 64 + 2 * tx) = make_double2(v[i][0], v[i][1]);
        *reinterpret_cast<double2 *>(outp + (size_t)row * 64 + 32 + 2 * tx) = make_double2(v[i][2], v[i][3]);
        if (tx == 0) a.scal[((size_t)cat * a.I + par) * Sp + s0 + row] = ex[i];
        if (is_root) {
            double r = v[i][0] * a.pi[2 * tx] + v[i][1] * a.pi[2 * tx + 1] + v[i][2] * a.pi[32 + 2 * tx] + v[i][3] * a.pi[33 + 2 * tx];
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
            if (tx == 0) {
                a.rootL[(size_t)cat * Sp + s0 + row] = r;
                a.rootE[(size_t)cat * Sp + s0 + row] = ex[i];
            }
        }
    }
    __syncthreads();                 // this node's tile (global) is visible to the whole CTA before a parent reads it
  }
}

// ------------------------------------------------------------------------------------------------
// Fused pruning update for small state spaces (Dp in {4,8,16,24,32}): one thread per pattern.
// grid = (Sp/128, jobs, classes), block = 128.  PT of each child is staged in shared memory (broadcast reads).
// ------------------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(128) prune_small_kernel(PruneArgs a, const int *__restrict__ jobs) {
    __shared__ double Ps[DP * DP];
    const int tid = threadIdx.x;
    const int par = jobs[blockIdx.y];
    const int cat = a.cat0 + blockIdx.z;
    const size_t Sp = a.Sp;
    const size_t s = (size_t)blockIdx.x * 128 + tid;       // Sp is a multiple of 128 for the small kernels
    double v[DP];
#pragma unroll
    for (int k = 0; k < DP; k++) v[k] = 1.0;
    int ex = 0;
    const int c_begin = a.tree.child_start[par], c_end = a.tree.child_start[par + 1];
    for (int ci = c_begin; ci < c_end; ci++) {
        const int child = a.tree.child_ids[ci];
        const double *PT = a.PT + ((size_t)cat * a.B + child) * DP * DP;
        __syncthreads();
        for (int idx = tid; idx < DP * DP; idx += 128) Ps[idx] = __ldg(PT + idx);
        __syncthreads();
        if (child < a.L) {
            const int code = (child == a.forced_node) ? a.forced[s] : a.leaf[(size_t)child * Sp + s];
            if (code >= 0) {
#pragma unroll
                for (int k = 0; k < DP; k++) v[k] *= Ps[code * DP + k];
            } else {
                const double *amb = a.ambig + (size_t)(-code - 1) * DP;
                double acc[DP];
#pragma unroll
                for (int k = 0; k < DP; k++) acc[k] = 0.0;
                for (int j = 0; j < a.D; j++) {
                    const double w = __ldg(amb + j);
#pragma unroll
                    for (int k = 0; k < DP; k++) acc[k] = fma(w, Ps[j * DP + k], acc[k]);
                }
#pragma unroll
                for (int k = 0; k < DP; k++) v[k] *= acc[k];
            }
        } else {
            const int cin = child - a.L;
            const double *X = a.cond + (((size_t)cat * a.I + cin) * Sp + s) * DP;
            double acc[DP];
#pragma unroll
            for (int k = 0; k < DP; k++) acc[k] = 0.0;
#pragma unroll
            for (int j = 0; j < DP; j += 2) {
                const double2 x = *reinterpret_cast<const double2 *>(X + j);
#pragma unroll
                for (int k = 0; k < DP; k++) acc[k] = fma(x.x, Ps[j * DP + k], acc[k]);
#pragma unroll
                for (int k = 0; k < DP; k++) acc[k] = fma(x.y, Ps[(j + 1) * DP + k], acc[k]);
            }
#pragma unroll
            for (int k = 0; k < DP; k++) v[k] *= acc[k];
            ex += a.scal[((size_t)cat * a.I + cin) * Sp + s];
        }
    }
    if (a.L + par == a.forced_node) {
        const int f = a.forced[s];
#pragma unroll
        for (int k = 0; k < DP; k++) if (k != f) v[k] = 0.0;
    }
    double m = 0.0;
#pragma unroll
    for (int k = 0; k < DP; k++) m = fmax(m, v[k]);
    if (m > 0.0 && m < INFINITY) {
        const int e = ilogb(m) + 1;
        const double s1 = exp2i(-(e / 2)), s2 = exp2i(-(e - e / 2));
#pragma unroll
        for (int k = 0; k < DP; k++) v[k] = v[k] * s1 * s2;
        ex += e;
    }
    double *outp = a.cond + (((size_t)cat * a.I + par) * Sp + s) * DP;
#pragma unroll
    for (int k = 0; k < DP; k += 2) *reinterpret_cast<double2 *>(outp + k) = make_double2(v[k], v[k + 1]);
    a.scal[((size_t)cat * a.I + par) * Sp + s] = ex;
    if (par == a.I - 1) {
        double r = 0.0;
#pragma unroll
        for (int k = 0; k < DP; k++) r = fma(v[k], a.pi[k], r);
        a.rootL[(size_t)cat * Sp + s] = r;
        a.rootE[(size_t)cat * Sp + s] = ex;
    }
}


// ------------------------------------------------------------------------------------------------
// Small state spaces, whole pruning pass in ONE launch: thread s owns pattern s and walks the dirty internal nodes in
// post-order.  Every dependency is thread-local (a child's conditionals were written by the same thread earlier in the
// same launch, or are a valid cache from an earlier evaluation), so there is no synchronisation of any kind; P^T of a
// branch is read by all threads at the same address (L1 broadcast).  HBM traffic per evaluation is the algorithmic
// minimum: one coalesced write and (at most) one read of Dp*8+4 bytes per (node, pattern) plus 4 B per leaf state.
// grid = (Sp/128, classes), block = 128.  jobs = dirty internal indices, ascending (= post-order).
// ------------------------------------------------------------------------------------------------
// (Forcing 16 CTAs/SM for 4 states -- 32 registers, spills -- was slower: 1.39 vs 1.18 ms at 256 x 200k, r01s.)
template <int DP>
__global__ void __launch_bounds__(128) prune_small_walk_kernel(PruneArgs a, const int *__restrict__ jobs, int njobs) {
    // 16..32 states: the branch's P^T is staged in shared memory once per child (all threads of the CTA walk the same
    // job list in lockstep) and read as broadcasts; 4/8 states read the 16/64 doubles straight through L1.
    constexpr bool kStage = DP > 8;
    // The node computed by the previous job is, in post-order, very often a child of the current one (every node follows
    // its last child): its vector is handed over in registers instead of being re-read from L2 (the thread would wait a
    // full L2 round trip for data it wrote a moment ago).  Register budget allows it up to 24 padded states.
    constexpr bool kChain = DP <= 24;
    __shared__ double Ps[kStage ? DP * DP : 1];
    const int tid = threadIdx.x;
    const int cat = a.cat0 + blockIdx.y;
    const size_t Sp = a.Sp;
    const size_t s = (size_t)blockIdx.x * 128 + tid;
    double pv[kChain ? DP : 1];
    int prev = -1, pex = 0;
    for (int jb = 0; jb < njobs; jb++) {
        const int par = __ldg(jobs + jb);
        double v[DP];
#pragma unroll
        for (int k = 0; k < DP; k++) v[k] = 1.0;
        int ex = 0;
        const int c_begin = __ldg(a.tree.child_start + par), c_end = __ldg(a.tree.child_start + par + 1);
        for (int ci = c_begin; ci < c_end; ci++) {
            const int child = __ldg(a.tree.child_ids + ci);
            const double *PTg = a.PT + ((size_t)cat * a.B + child) * DP * DP;
            const double *PT = PTg;
            if (kStage) {
                __syncthreads();
                for (int idx = tid; idx < DP * DP / 2; idx += 128)
                    reinterpret_cast<double2 *>(Ps)[idx] = __ldg(reinterpret_cast<const double2 *>(PTg) + idx);
                __syncthreads();
                PT = Ps;
            }
            if (child < a.L) {
                const int code = (child == a.forced_node) ? __ldg(a.forced + s) : __ldg(a.leaf + (size_t)child * Sp + s);
                if (code >= 0) {
                    const double2 *row = reinterpret_cast<const double2 *>(PT + (size_t)code * DP);
#pragma unroll
                    for (int k = 0; k < DP; k += 2) { const double2 r = kStage ? row[k / 2] : __ldg(row + k / 2); v[k] *= r.x; v[k + 1] *= r.y; }
                } else {
                    const double *amb = a.ambig + (size_t)(-code - 1) * DP;
                    double acc[DP];
#pragma unroll
                    for (int k = 0; k < DP; k++) acc[k] = 0.0;
                    for (int j = 0; j < a.D; j++) {
                        const double wgt = __ldg(amb + j);
                        const double2 *row = reinterpret_cast<const double2 *>(PT + (size_t)j * DP);
#pragma unroll
                        for (int k = 0; k < DP; k += 2) { const double2 r = kStage ? row[k / 2] : __ldg(row + k / 2); acc[k] = fma(wgt, r.x, acc[k]); acc[k + 1] = fma(wgt, r.y, acc[k + 1]); }
                    }
#pragma unroll
                    for (int k = 0; k < DP; k++) v[k] *= acc[k];
                }
            } else {
                const int cin = child - a.L;
                double x[DP];
                if (kChain && cin == prev) {
#pragma unroll
                    for (int j = 0; j < DP; j++) x[j] = pv[j];
                    ex += pex;
                } else {
                    const double2 *X = reinterpret_cast<const double2 *>(a.cond + (((size_t)cat * a.I + cin) * Sp + s) * DP);
#pragma unroll
                    for (int j = 0; j < DP; j += 2) { const double2 t = X[j / 2]; x[j] = t.x; x[j + 1] = t.y; }
                    ex += a.scal[((size_t)cat * a.I + cin) * Sp + s];
                }
                double acc[DP];
#pragma unroll
                for (int k = 0; k < DP; k++) acc[k] = 0.0;
#pragma unroll
                for (int j = 0; j < DP; j++) {
                    const double2 *row = reinterpret_cast<const double2 *>(PT + (size_t)j * DP);
#pragma unroll
                    for (int k = 0; k < DP; k += 2) { const double2 r = kStage ? row[k / 2] : __ldg(row + k / 2); acc[k] = fma(x[j], r.x, acc[k]); acc[k + 1] = fma(x[j], r.y, acc[k + 1]); }
                }
#pragma unroll
                for (int k = 0; k < DP; k++) v[k] *= acc[k];
            }
        }
        if (a.L + par == a.forced_node) {
            const int f = __ldg(a.forced + s);
#pragma unroll
            for (int k = 0; k < DP; k++) if (k != f) v[k] = 0.0;
        }
        double m = 0.0;
#pragma unroll
        for (int k = 0; k < DP; k++) m = fmax(m, v[k]);
        if (m > 0.0 && m < INFINITY) {
            const int e = ilogb(m) + 1;
            const double s1 = exp2i(-(e / 2)), s2 = exp2i(-(e - e / 2));
#pragma unroll
            for (int k = 0; k < DP; k++) v[k] = v[k] * s1 * s2;
            ex += e;
        }
        double2 *outp = reinterpret_cast<double2 *>(a.cond + (((size_t)cat * a.I + par) * Sp + s) * DP);
#pragma unroll
        for (int k = 0; k < DP; k += 2) outp[k / 2] = make_double2(v[k], v[k + 1]);
        a.scal[((size_t)cat * a.I + par) * Sp + s] = ex;
        if (kChain) {
#pragma unroll
            for (int k = 0; k < DP; k++) pv[k] = v[k];
            prev = par; pex = ex;
        }
        if (par == a.I - 1) {
            double r = 0.0;
#pragma unroll
            for (int k = 0; k < DP; k++) r = fma(v[k], a.pi[k], r);
            a.rootL[(size_t)cat * Sp + s] = r;
            a.rootE[(size_t)cat * Sp + s] = ex;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 4 / 8 padded states, PP patterns per thread: the one-pattern kernel above is bound by the latency of each thread's serial
// node chain, and -- every CTA walking the whole tree -- by wave quantisation (256 x 200 k nucleotides: 1563 CTAs on 1332
// resident = 1.17 waves, i.e. two passes).  With two patterns per thread the loads and the arithmetic of both are in flight
// together (the pass gets 1.55x longer, not 2x) and half as many CTAs cover the alignment; the host picks the variant
// with the cheaper (passes x pass cost).  Thread t of CTA b owns patterns (b*PP + u)*128 + t: coalesced for every u.
// ------------------------------------------------------------------------------------------------
template <int DP, int PP>
__global__ void __launch_bounds__(128, 6) prune_small_walk_ilp_kernel(PruneArgs a, const int *__restrict__ jobs, int njobs) {
    const int tid = threadIdx.x;
    const int cat = a.cat0 + blockIdx.y;
    const size_t Sp = a.Sp;
    size_t s[PP];
    bool live[PP];
#pragma unroll
    for (int u = 0; u < PP; u++) {
        s[u] = ((size_t)blockIdx.x * PP + u) * 128 + tid;
        live[u] = s[u] < Sp;
        if (!live[u]) s[u] = Sp - 1;          // loads stay in range; nothing is stored for a dead slot
    }
    double pv[PP][DP];
    int prev = -1, pex[PP];
#pragma unroll
    for (int u = 0; u < PP; u++) pex[u] = 0;
    for (int jb = 0; jb < njobs; jb++) {
        const int par = __ldg(jobs + jb);
        double v[PP][DP];
        int ex[PP];
#pragma unroll
        for (int u = 0; u < PP; u++) {
            ex[u] = 0;
#pragma unroll
            for (int k = 0; k < DP; k++) v[u][k] = 1.0;
        }
        const int c_begin = __ldg(a.tree.child_start + par), c_end = __ldg(a.tree.child_start + par + 1);
        for (int ci = c_begin; ci < c_end; ci++) {
            const int child = __ldg(a.tree.child_ids + ci);
            const double *PT = a.PT + ((size_t)cat * a.B + child) * DP * DP;
            if (child < a.L) {
                int code[PP];
#pragma unroll
                for (int u = 0; u < PP; u++) code[u] = (child == a.forced_node) ? __ldg(a.forced + s[u]) : __ldg(a.leaf + (size_t)child * Sp + s[u]);
#pragma unroll
                for (int u = 0; u < PP; u++) {
                    if (code[u] >= 0) {
                        const double2 *row = reinterpret_cast<const double2 *>(PT + (size_t)code[u] * DP);
#pragma unroll
                        for (int k = 0; k < DP; k += 2) { const double2 r = __ldg(row + k / 2); v[u][k] *= r.x; v[u][k + 1] *= r.y; }
                    } else {
                        const double *amb = a.ambig + (size_t)(-code[u] - 1) * DP;
                        double acc[DP];
#pragma unroll
                        for (int k = 0; k < DP; k++) acc[k] = 0.0;
                        for (int j = 0; j < a.D; j++) {
                            const double wgt = __ldg(amb + j);
                            const double2 *row = reinterpret_cast<const double2 *>(PT + (size_t)j * DP);
#pragma unroll
                            for (int k = 0; k < DP; k += 2) { const double2 r = __ldg(row + k / 2); acc[k] = fma(wgt, r.x, acc[k]); acc[k + 1] = fma(wgt, r.y, acc[k + 1]); }
                        }
#pragma unroll
                        for (int k = 0; k < DP; k++) v[u][k] *= acc[k];
                    }
                }
            } else {
                const int cin = child - a.L;
                double x[PP][DP];
                if (cin == prev) {
#pragma unroll
                    for (int u = 0; u < PP; u++) {
#pragma unroll
                        for (int j = 0; j < DP; j++) x[u][j] = pv[u][j];
                        ex[u] += pex[u];
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < PP; u++) {
                        const double2 *X = reinterpret_cast<const double2 *>(a.cond + (((size_t)cat * a.I + cin) * Sp + s[u]) * DP);
#pragma unroll
                        for (int j = 0; j < DP; j += 2) { const double2 t = X[j / 2]; x[u][j] = t.x; x[u][j + 1] = t.y; }
                    }
#pragma unroll
                    for (int u = 0; u < PP; u++) ex[u] += a.scal[((size_t)cat * a.I + cin) * Sp + s[u]];
                }
                double acc[PP][DP];
#pragma unroll
                for (int u = 0; u < PP; u++)
#pragma unroll
                    for (int k = 0; k < DP; k++) acc[u][k] = 0.0;
#pragma unroll
                for (int j = 0; j < DP; j++) {
                    const double2 *row = reinterpret_cast<const double2 *>(PT + (size_t)j * DP);
#pragma unroll
                    for (int k = 0; k < DP; k += 2) {
                        const double2 r = __ldg(row + k / 2);           // one load serves all PP patterns
#pragma unroll
                        for (int u = 0; u < PP; u++) { acc[u][k] = fma(x[u][j], r.x, acc[u][k]); acc[u][k + 1] = fma(x[u][j], r.y, acc[u][k + 1]); }
                    }
                }
#pragma unroll
                for (int u = 0; u < PP; u++)
#pragma unroll
                    for (int k = 0; k < DP; k++) v[u][k] *= acc[u][k];
            }
        }
        const bool pinned = a.L + par == a.forced_node;
#pragma unroll
        for (int u = 0; u < PP; u++) {
            if (pinned) {
                const int f = __ldg(a.forced + s[u]);
#pragma unroll
                for (int k = 0; k < DP; k++) if (k != f) v[u][k] = 0.0;
            }
            double m = 0.0;
#pragma unroll
            for (int k = 0; k < DP; k++) m = fmax(m, v[u][k]);
            if (m > 0.0 && m < INFINITY) {
                const int e = ilogb(m) + 1;
                const double s1 = exp2i(-(e / 2)), s2 = exp2i(-(e - e / 2));
#pragma unroll
                for (int k = 0; k < DP; k++) v[u][k] = v[u][k] * s1 * s2;
                ex[u] += e;
            }
            if (live[u]) {
                double2 *outp = reinterpret_cast<double2 *>(a.cond + (((size_t)cat * a.I + par) * Sp + s[u]) * DP);
#pragma unroll
                for (int k = 0; k < DP; k += 2) outp[k / 2] = make_double2(v[u][k], v[u][k + 1]);
                a.scal[((size_t)cat * a.I + par) * Sp + s[u]] = ex[u];
                if (par == a.I - 1) {
                    double r = 0.0;
#pragma unroll
                    for (int k = 0; k < DP; k++) r = fma(v[u][k], a.pi[k], r);
                    a.rootL[(size_t)cat * Sp + s[u]] = r;
                    a.rootE[(size_t)cat * Sp + s[u]] = ex[u];
                }
            }
#pragma unroll
            for (int k = 0; k < DP; k++) pv[u][k] = v[u][k];
            pex[u] = ex[u];
        }
        prev = par;
    }
}

// ------------------------------------------------------------------------------------------------
// Root reduction: rate-class mixture + log + pattern frequency + block partial sums.
//   L_s = sum_c w_c * rootL[c][s] * 2^(rootE[c][s]);  lnL_s = log(sum_c w_c rootL 2^(e_c-emax)) + emax ln2
// Optional per-pattern outputs in the reference's convention: siteL * 2^(-64*siteScale).
// partial[block] = sum over the block's patterns of freq*lnL_s (tree reduction, deterministic);
// flag[0] |= 1 if any pattern with freq>0 has L <= 0.
// ------------------------------------------------------------------------------------------------
struct CombineArgs {
    const double *rootL; const int *rootE; const double *weights; const double *freq;
    double *partial; int *flag; double *siteL; long long *siteScale;
    int Sp, S, c0, nc;
    unsigned *counter;      // nullable: with `out`, the LAST block to finish adds up the partials (final_sum_kernel's arithmetic,
    double *out;            // same order) and re-arms counter and flag -- one launch and one memset less per evaluation
};

__global__ void __launch_bounds__(256) combine_kernel(CombineArgs a) {
    __shared__ double red[256];
    const int s = blockIdx.x * 256 + threadIdx.x;
    double term = 0.0;
    if (s < a.S) {
        int emax = INT_MIN;
        for (int c = 0; c < a.nc; c++) {
            const double l = a.rootL[(size_t)(a.c0 + c) * a.Sp + s];
            if (l > 0.0) emax = max(emax, a.rootE[(size_t)(a.c0 + c) * a.Sp + s]);
        }
        double sum = 0.0;
        bool nan_seen = false;
        for (int c = 0; c < a.nc; c++) {
            const double l = a.rootL[(size_t)(a.c0 + c) * a.Sp + s];
            const double w = a.weights ? a.weights[c] : 1.0;
            if (l != l) nan_seen = true;
            if (l > 0.0) {
                const int de = a.rootE[(size_t)(a.c0 + c) * a.Sp + s] - emax;       // <= 0
                sum += w * l * (de < -1000 ? 0.0 : exp2i(de));
            }
        }
        if (nan_seen) sum = __longlong_as_double(0x7ff8000000000000LL);
        if (emax == INT_MIN) emax = 0;
        const double f = a.freq[s];
        double lnl;
        if (sum > 0.0) lnl = log(sum) + (double)emax * 0.693147180559945309417232121458;
        else if (sum != sum) lnl = sum;
        else { lnl = -INFINITY; if (f > 0.0) atomicOr(a.flag, 1); }
        term = (f > 0.0) ? f * lnl : 0.0;
        if (a.siteL) {
            // reference convention: L_true = siteL * 2^(-64*count); keep siteL in (2^-64, 1]
            long long cnt = 0; double outL = sum;
            if (sum > 0.0) {
                int e2 = emax;                       // L_true = sum * 2^e2
                cnt = (e2 < 0) ? (long long)((-e2) / 64) : -(long long)((e2 + 63) / 64);
                const int rem = e2 + (int)(64 * cnt);    // in (-64, 0]
                outL = sum * exp2i(rem);
            }
            a.siteL[s] = outL; a.siteScale[s] = cnt;
        }
    }
    red[threadIdx.x] = term;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.partial[blockIdx.x] = red[0];
    if (a.counter) {
        __shared__ int s_last;
        if (threadIdx.x == 0) {
            __threadfence();                                   // partial[] (and flag) before the ticket
            s_last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) s += __ldcg(a.partial + i);
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            a.out[0] = __ldcg(a.flag) ? -INFINITY : red[0];
            *a.flag = 0;
            *a.counter = 0u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Single-branch shortcut (reference: ComputeBranchCache tree_evaluator.cpp:4286, ComputeLLWithBranchCache tree.cpp:3383).
// For a branch b with parent u the likelihood factorises as  L_s = sum_a rest_b[s][a] * (P_b * below_b[s])[a]  where
// below_b = conditionals of b's subtree (already resident) and rest_b[s][a] = everything else, seen from state a at u:
//   out_root[a] = pi_a ;  rest_v[a] = out_u[a] * prod_{siblings c of v at u} (P_c L_c)[a] ;  out_v[a'] = sum_a rest_v[a] P_v[a][a'].
// (No reversibility is needed: the outside vectors are propagated with the transposed matrices instead of re-rooting.)
// bc_path_kernel walks the root->u path once; bc_eval_kernel then costs O(S*D^2) per probe of P_b.
// fp64 throughout; the build runs one CTA per (class, pattern), the probe one thread per (class, pattern).  CondView abstracts the two conditional layouts (fp64 [I][Sp][Dp]; the walk kernel's
// tile-wise fp32 with generation tags).
// ------------------------------------------------------------------------------------------------
struct CondView {
    const double *c64;      // fp64 layout or null
    const float *c32;       // tile-wise fp32 layout ([I][Sp/128][16][128][4], sign bit = tag) or null
    const int *scal;        // exponents; tagged = 1 -> (value << 1) | tag
    int I, Sp, Dp, tagged;
};

__device__ __forceinline__ double cond_at(const CondView &v, int cat, int node, int s, int k) {
    if (v.c64) return v.c64[(((size_t)cat * v.I + node) * v.Sp + s) * v.Dp + k];
    const size_t tile = s >> 7, t = s & 127;
    return (double)fabsf(v.c32[(((((size_t)cat * v.I + node) * (v.Sp >> 7) + tile) * 16 + (k >> 2)) * 128 + t) * 4 + (k & 3)]);
}
__device__ __forceinline__ int exp_at(const CondView &v, int cat, int node, int s) {
    const int e = v.scal[((size_t)cat * v.I + node) * v.Sp + s];
    return v.tagged ? (e >> 1) : e;
}

struct BranchCacheArgs {
    CondView cv;
    const double *PT;       // [C][B][Dp*Dp] transposed transition matrices
    const int *leaf;        // [L][Sp]
    const double *ambig;    // [nAmb][Dp]
    const double *pi;       // [Dp]
    double *out;            // [C][Sp][Dp] outside vector of the current path node (in/out)
    int *outE;              // [C][Sp]
    int L, B, D, Dp, Sp, S, cat0, ncls;
};

// message of child `ch` (flat id) towards its parent for one pattern: m[a] = sum_k P_ch[a][k] x[k]  (PT[k][a] = P[a][k])
__device__ __forceinline__ void bc_message(const BranchCacheArgs &a, int cat, int ch, int s, double *m, int &e) {
    const double *PTc = a.PT + ((size_t)cat * a.B + ch) * a.Dp * a.Dp;
    if (ch < a.L) {
        const int code = a.leaf[(size_t)ch * a.Sp + s];
        if (code >= 0) {
            for (int q = 0; q < a.D; q++) m[q] = __ldg(PTc + (size_t)code * a.Dp + q);
        } else {
            const double *amb = a.ambig + (size_t)(-code - 1) * a.Dp;
            for (int q = 0; q < a.D; q++) m[q] = 0.0;
            for (int k = 0; k < a.D; k++)
                if (__ldg(amb + k) != 0.0)
                    for (int q = 0; q < a.D; q++) m[q] += __ldg(PTc + (size_t)k * a.Dp + q);
        }
    } else {
        const int ci = ch - a.L;
        for (int q = 0; q < a.D; q++) m[q] = 0.0;
        for (int k = 0; k < a.D; k++) {
            const double x = cond_at(a.cv, cat, ci, s, k);
            if (x != 0.0)
                for (int q = 0; q < a.D; q++) m[q] = fma(x, __ldg(PTc + (size_t)k * a.Dp + q), m[q]);
        }
        e += exp_at(a.cv, cat, ci, s);
    }
}

__device__ __forceinline__ void bc_renorm(double *v, int D, int &e) {
    double m = 0.0;
    for (int q = 0; q < D; q++) m = fmax(m, v[q]);
    if (m > 0.0 && m < INFINITY) {
        int ex;
        frexp(m, &ex);
        if (ex != 0) {
            const double sc = exp2i(-ex);
            for (int q = 0; q < D; q++) v[q] *= sc;
            e += ex;
        }
    }
}

// The whole root -> parent(b) path in ONE launch: one CTA of Dp' = 64 (or 32) threads per (class, pattern), thread q owns
// state q; vectors are exchanged through shared memory and a matrix column / row is read with consecutive threads on
// consecutive addresses (the transposed table serves messages, its rows serve the outside propagation).  Replaces one
// launch of bc_step_kernel per path node with a serial thread per pattern (17.7 ms at depth 48 on the north-star shape).
//   path[i]      internal index of the i-th path node (root first), npath nodes
//   sib_off[i]   range of its non-path children in sib[]
//   down[i]      flat id of the path child below path[i] (the next path node), -1 for the last (whose child is the branch b)
struct BranchPathArgs {
    const int *sib, *sib_off, *down;
    int npath;
};

template <int NT>
__global__ void __launch_bounds__(NT) bc_path_kernel(BranchCacheArgs a, BranchPathArgs bp) {
    __shared__ double xs[64];
    __shared__ double red[2];
    const int s = blockIdx.x, cat = a.cat0 + blockIdx.y, q = threadIdx.x;
    const bool live = q < a.D;
    double r = live ? a.pi[q] : 0.0;             // out_root = pi
    int e = 0;
    auto block_max = [&](double v) -> double {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (NT > 32) {
            __syncthreads();
            if ((q & 31) == 0) red[q >> 5] = v;
            __syncthreads();
            v = fmax(red[0], red[1]);
        }
        return v;
    };
    auto renorm = [&](double &v) {
        const double m = block_max(v);
        if (m > 0.0 && m < INFINITY) {
            int ex;
            frexp(m, &ex);
            if (ex != 0) { v *= exp2i(-(ex / 2)) * exp2i(-(ex - ex / 2)); e += ex; }
        }
    };
    for (int i = 0; i < bp.npath; i++) {
        for (int ci = bp.sib_off[i]; ci < bp.sib_off[i + 1]; ci++) {          // rest = out_u * prod over siblings (P_c L_c)
            const int ch = bp.sib[ci];
            const double *PTc = a.PT + ((size_t)cat * a.B + ch) * a.Dp * a.Dp;
            double m = 0.0;
            if (ch < a.L) {
                const int code = a.leaf[(size_t)ch * a.Sp + s];
                if (code >= 0) m = live ? PTc[(size_t)code * a.Dp + q] : 0.0;
                else {
                    const double *amb = a.ambig + (size_t)(-code - 1) * a.Dp;
                    for (int k = 0; k < a.D; k++) if (amb[k] != 0.0 && live) m += PTc[(size_t)k * a.Dp + q];
                }
            } else {
                const int cin = ch - a.L;
                __syncthreads();
                if (q < a.Dp) xs[q] = live ? cond_at(a.cv, cat, cin, s, q) : 0.0;
                __syncthreads();
                if (live) for (int k = 0; k < a.D; k++) m = fma(xs[k], PTc[(size_t)k * a.Dp + q], m);
                e += exp_at(a.cv, cat, cin, s);
            }
            r *= m;
            renorm(r);
        }
        if (bp.down[i] >= 0) {                                                  // out'[q] = sum_k rest[k] * P_down[k][q] = sum_k r[k] PT[q][k]
            const double *PTd = a.PT + ((size_t)cat * a.B + bp.down[i]) * a.Dp * a.Dp;
            __syncthreads();
            if (q < a.Dp) xs[q] = r;
            __syncthreads();
            double acc = 0.0;
            if (live) for (int k = 0; k < a.D; k++) acc = fma(xs[k], PTd[(size_t)q * a.Dp + k], acc);
            r = acc;
            renorm(r);
        }
    }
    if (q < a.Dp) a.out[((size_t)cat * a.Sp + s) * a.Dp + q] = r;
    if (q == 0) a.outE[(size_t)cat * a.Sp + s] = e;
}

// Probe: rootL/rootE <- sum_a rest[a] * (P_b * below)[a] with the CURRENT matrix of branch b.
__global__ void __launch_bounds__(128) bc_eval_kernel(BranchCacheArgs a, int b, double *rootL, int *rootE) {
    const int s = blockIdx.x * 128 + threadIdx.x;
    const int cat = a.cat0 + blockIdx.y;
    if (s >= a.S) return;
    double m[64];
    int e = a.outE[(size_t)cat * a.Sp + s];
    bc_message(a, cat, b, s, m, e);
    const double *o = a.out + ((size_t)cat * a.Sp + s) * a.Dp;
    double acc = 0.0;
    for (int q = 0; q < a.D; q++) acc = fma(o[q], m[q], acc);
    rootL[(size_t)cat * a.Sp + s] = acc;
    rootE[(size_t)cat * a.Sp + s] = e;
}

// Class-group sharding (hb2_comm_class_groups): this rank's share of sum_c w_c L_{c,s}, as (value, binary exponent) per
// pattern, written to send[0..xs) | send[xs..2xs) (exponent as a double; NaN value = numerical failure, value 0 = no
// positive class likelihood).
__global__ void __launch_bounds__(256) class_partial_kernel(const double *rootL, const int *rootE, const double *weights,
                                                            int Sp, int S, int c0, int nc, double *send, int xs) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= xs) return;
    double sum = 0.0;
    int emax = 0;
    if (s < S) {
        emax = INT_MIN;
        bool nan_seen = false;
        for (int c = 0; c < nc; c++) {
            const double l = rootL[(size_t)(c0 + c) * Sp + s];
            if (l != l) nan_seen = true;
            if (l > 0.0) emax = max(emax, rootE[(size_t)(c0 + c) * Sp + s]);
        }
        for (int c = 0; c < nc; c++) {
            const double l = rootL[(size_t)(c0 + c) * Sp + s];
            if (l > 0.0) {
                const int de = rootE[(size_t)(c0 + c) * Sp + s] - emax;
                sum += weights[c0 + c] * l * (de < -1000 ? 0.0 : exp2i(de));
            }
        }
        if (nan_seen) sum = __longlong_as_double(0x7ff8000000000000LL);
        if (emax == INT_MIN) emax = 0;
    }
    send[s] = sum;
    send[xs + s] = (double)emax;
}

// Merges the class-group partials of EVERY pattern shard (gath = gathered buffer, rank-major, 2*xs doubles per rank;
// the G ranks [j*G, j*G+G) hold the class groups of shard j; xfreq = pattern frequencies of every rank, gathered once)
// and finishes like combine_kernel.  Every rank therefore ends up with the complete lnL -- bit-identical on all ranks --
// and no second collective is needed; per-pattern outputs are written for the rank's own shard only.
__global__ void __launch_bounds__(256) class_merge_kernel(const double *gath, const double *xfreq, int xs, int nShards, int G,
                                                          int myShard, double *partial, int *flag, double *siteL,
                                                          long long *siteScale, int S) {
    __shared__ double red[256];
    const int idx = blockIdx.x * 256 + threadIdx.x;
    double term = 0.0;
    if (idx < nShards * xs) {
        const int j = idx / xs, s = idx % xs;
        const double f = xfreq[(size_t)j * G * xs + s];
        int emax = INT_MIN;
        bool nan_seen = false;
        for (int r = 0; r < G; r++) {
            const double *q = gath + (size_t)(j * G + r) * 2 * xs;
            const double m = q[s];
            if (m != m) nan_seen = true;
            if (m > 0.0) emax = max(emax, (int)q[xs + s]);
        }
        double sum = 0.0;
        for (int r = 0; r < G; r++) {
            const double *q = gath + (size_t)(j * G + r) * 2 * xs;
            const double m = q[s];
            if (m > 0.0) {
                const int de = (int)q[xs + s] - emax;
                sum += m * (de < -1000 ? 0.0 : exp2i(de));
            }
        }
        if (nan_seen) sum = __longlong_as_double(0x7ff8000000000000LL);
        if (emax == INT_MIN) emax = 0;
        double lnl;
        if (sum > 0.0) lnl = log(sum) + (double)emax * 0.693147180559945309417232121458;
        else if (sum != sum) lnl = sum;
        else { lnl = -INFINITY; if (f > 0.0) atomicOr(flag, 1); }
        term = (f > 0.0) ? f * lnl : 0.0;
        if (siteL && j == myShard && s < S) {
            long long cnt = 0; double outL = sum;
            if (sum > 0.0) {
                const int e2 = emax;
                cnt = (e2 < 0) ? (long long)((-e2) / 64) : -(long long)((e2 + 63) / 64);
                const int rem = e2 + (int)(64 * cnt);
                outL = sum * exp2i(rem);
            }
            siteL[s] = outL; siteScale[s] = cnt;
        }
    }
    red[threadIdx.x] = term;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// ------------------------------------------------------------------------------------------------
// Batched one-pattern likelihoods (the site phases of FEL / MEME fit thousands of one-pattern likelihood
// functions on the same tree, FEL.bf:1180-1233, MEME.bf:699-746).  Set s prunes pattern pat[s] with ITS OWN transition
// matrices PT[s][b] (64-padded, transposed, as the expm kernels write them).  One CTA of 64 threads per set: thread k owns
// parent state k; a child's vector is broadcast from shared memory and column k of P is read from PT[.][j][k] -- 64
// consecutive doubles per j, coalesced.  fp64 throughout; per-node power-of-two renormalisation like the main kernels.
// Internal-node vectors live in global scratch cond[s][I][64] (L1/L2 resident: 512 bytes per node).
// ------------------------------------------------------------------------------------------------
struct BatchPruneArgs {
    const double *PT;         // [nSets][B][4096]
    double *cond;             // [nSets][I][64]
    int *node_ex;             // [nSets][I] binary exponents of the node vectors
    const int *pat;           // [nSets] pattern of the partition evaluated by each set
    const int *leaf;          // [L][Sp]
    const double *ambig;      // [nAmb][64]
    const double *pi;         // [64]
    TreeDev tree;
    double *out;              // [nSets] log-likelihood of the pattern
    int L, I, B, D, Sp;
};

__global__ void __launch_bounds__(64) prune_batch_kernel(BatchPruneArgs a) {
    __shared__ double xs[64];
    __shared__ double red[2];
    const int s = blockIdx.x, k = threadIdx.x;
    const int pattern = a.pat[s];
    const double *PTs = a.PT + (size_t)s * a.B * 4096;
    double *cs = a.cond + (size_t)s * a.I * 64;
    int *node_ex = a.node_ex + (size_t)s * a.I;
    for (int par = 0; par < a.I; par++) {
        double v = k < a.D ? 1.0 : 0.0;
        int ex = 0;
        const int c_begin = a.tree.child_start[par], c_end = a.tree.child_start[par + 1];
        for (int ci = c_begin; ci < c_end; ci++) {
            const int child = a.tree.child_ids[ci];
            const double *PT = PTs + (size_t)child * 4096;
            double m = 0.0;
            if (child < a.L) {
                const int code = a.leaf[(size_t)child * a.Sp + pattern];
                if (code >= 0) m = PT[code * 64 + k];
                else {
                    const double *amb = a.ambig + (size_t)(-code - 1) * 64;
                    for (int j = 0; j < a.D; j++) { const double w = amb[j]; if (w != 0.0) m = fma(w, PT[j * 64 + k], m); }
                }
            } else {
                const int cin = child - a.L;
                __syncthreads();                 // xs free
                xs[k] = cs[(size_t)cin * 64 + k];
                __syncthreads();
                for (int j = 0; j < a.D; j++) m = fma(xs[j], PT[j * 64 + k], m);
                ex += node_ex[cin];
            }
            v *= m;
        }
        // renormalise: max over the 64 threads -> [0.5, 1)
        double mx = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        __syncthreads();
        if ((k & 31) == 0) red[k >> 5] = mx;
        __syncthreads();
        mx = fmax(red[0], red[1]);
        if (mx > 0.0 && mx < INFINITY) {
            const int e = ilogb(mx) + 1;
            v = v * exp2i(-(e / 2)) * exp2i(-(e - e / 2));
            ex += e;
        }
        cs[(size_t)par * 64 + k] = v;
        if (k == 0) node_ex[par] = ex;
        __syncthreads();                         // node_ex / cond of this node visible to the block before a parent reads them
        if (par == a.I - 1) {
            double r = v * a.pi[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
            if ((k & 31) == 0) red[k >> 5] = r;
            __syncthreads();
            if (k == 0) {
                const double L = red[0] + red[1];
                a.out[s] = L > 0.0 ? log(L) + (double)ex * 0.693147180559945309417232121458 : (L != L ? L : -INFINITY);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Peer exchange over NVLink (one process per GPU; every rank maps every other rank's exchange buffer through CUDA IPC).
// Buffer of a rank:  data[2 parities][R source ranks][payload doubles]  followed by  flags[2][R] (uint64 generation).
// A writer stores its payload into slot `rank` of EVERY rank's buffer (its own included), fences system-wide and then
// publishes the generation number in the matching flag; a reader spins (time-bounded) on its LOCAL flags.  Two
// parities suffice: a rank can run at most one evaluation ahead of the slowest one (it needs that rank's flag to finish).
// ------------------------------------------------------------------------------------------------
struct PeerBuf {
    double *const *peer;        // [R] device pointers to the ranks' buffers (entry `rank` = the local one)
    int R, rank, payload;
    unsigned long long gen;     // generation of THIS exchange (1, 2, 3, ...)
    int *err;
};
__device__ __forceinline__ double *peer_data(const PeerBuf &b, int dst_rank, int src_rank) {
    return b.peer[dst_rank] + ((size_t)(b.gen & 1ull) * b.R + src_rank) * b.payload;
}
__device__ __forceinline__ unsigned long long *peer_flag(const PeerBuf &b, int dst_rank, int src_rank) {
    return reinterpret_cast<unsigned long long *>(b.peer[dst_rank] + (size_t)2 * b.R * b.payload) + (b.gen & 1ull) * b.R + src_rank;
}
__device__ __forceinline__ void peer_publish(const PeerBuf &b, int dst_rank) {       // after __threadfence_system()
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag(b, dst_rank, b.rank)), "l"(b.gen) : "memory");
}
__device__ __forceinline__ bool peer_await(const PeerBuf &b, int src_rank) {
    const unsigned long long *f = peer_flag(b, b.rank, src_rank);
    unsigned long long t0 = 0;
    for (int it = 0;; it++) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        if (v == b.gen) return true;
        if ((it & 255) == 255) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000ull) { atomicExch(b.err, 3); return false; }      // 20 s: a rank is gone
        }
    }
}

// Pattern shards only: every rank contributes its partial lnL; all ranks add the R partials in rank order, so the
// result is bit-identical everywhere.  One block, launched after final_sum_kernel.  payload >= 1.
__global__ void __launch_bounds__(64) peer_sum_kernel(PeerBuf b, double *lnL) {
    const int t = threadIdx.x;
    if (t < b.R) *peer_data(b, t, b.rank) = *lnL;
    __threadfence_system();
    __syncthreads();
    if (t < b.R) peer_publish(b, t);
    if (t < b.R) peer_await(b, t);
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
        for (int r = 0; r < b.R; r++) s += __ldcg(peer_data(b, b.rank, r));
        *lnL = s;
    }
}

// Class groups: fused class-partial computation + exchange + merge + final sum (replaces class_partial_kernel, the
// zero-padded all-reduce used as a gather, class_merge_kernel and final_sum_kernel).  payload = 2*xs doubles per rank:
// [0,xs) partial values, [xs,2xs) binary exponents.  The block that finishes its stores LAST publishes the flags, waits
// for the other ranks' and merges every shard (fixed order: bit-identical lnL on all ranks).
struct ClassXchgArgs {
    const double *rootL; const int *rootE; const double *weights;
    int Sp, S, c0, nc, xs;
    PeerBuf pb;
    const double *xfreq;        // [nShards*G][xs] pattern frequencies of every rank (gathered once)
    int nShards, G, myShard;
    double *lnL; double *siteL; long long *siteScale;
    unsigned int *counter;
};

__global__ void __launch_bounds__(256) class_exchange_kernel(ClassXchgArgs a) {
    __shared__ double red[256];
    __shared__ int s_last, s_bad;
    const int tid = threadIdx.x;
    const PeerBuf &b = a.pb;
    const int xs = a.xs;
    for (int s = blockIdx.x * 256 + tid; s < xs; s += gridDim.x * 256) {
        double sum = 0.0;
        int emax = 0;
        if (s < a.S) {
            emax = INT_MIN;
            bool nan_seen = false;
            for (int c = 0; c < a.nc; c++) {
                const double l = a.rootL[(size_t)(a.c0 + c) * a.Sp + s];
                if (l != l) nan_seen = true;
                if (l > 0.0) emax = max(emax, a.rootE[(size_t)(a.c0 + c) * a.Sp + s]);
            }
            for (int c = 0; c < a.nc; c++) {
                const double l = a.rootL[(size_t)(a.c0 + c) * a.Sp + s];
                if (l > 0.0) {
                    const int de = a.rootE[(size_t)(a.c0 + c) * a.Sp + s] - emax;
                    sum += a.weights[a.c0 + c] * l * (de < -1000 ? 0.0 : exp2i(de));
                }
            }
            if (nan_seen) sum = __longlong_as_double(0x7ff8000000000000LL);
            if (emax == INT_MIN) emax = 0;
        }
        for (int r = 0; r < b.R; r++) {
            double *d = peer_data(b, r, b.rank);
            d[s] = sum;
            d[xs + s] = (double)emax;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    if (tid == 0) { *a.counter = 0u; s_bad = 0; }
    __threadfence_system();
    if (tid < b.R) peer_publish(b, tid);
    if (tid < b.R) peer_await(b, tid);
    __syncthreads();
    double term = 0.0;
    for (int idx = tid; idx < a.nShards * xs; idx += 256) {
        const int j = idx / xs, s = idx - j * xs;
        const double f = a.xfreq[(size_t)j * a.G * xs + s];
        int emax = INT_MIN;
        bool nan_seen = false;
        for (int r = 0; r < a.G; r++) {
            const double *q = peer_data(b, b.rank, j * a.G + r);
            const double m = __ldcg(q + s);
            if (m != m) nan_seen = true;
            if (m > 0.0) emax = max(emax, (int)__ldcg(q + xs + s));
        }
        double sum = 0.0;
        for (int r = 0; r < a.G; r++) {
            const double *q = peer_data(b, b.rank, j * a.G + r);
            const double m = __ldcg(q + s);
            if (m > 0.0) {
                const int de = (int)__ldcg(q + xs + s) - emax;
                sum += m * (de < -1000 ? 0.0 : exp2i(de));
            }
        }
        if (nan_seen) sum = __longlong_as_double(0x7ff8000000000000LL);
        if (emax == INT_MIN) emax = 0;
        double lnl;
        if (sum > 0.0) lnl = log(sum) + (double)emax * 0.693147180559945309417232121458;
        else if (sum != sum) lnl = sum;
        else { lnl = -INFINITY; if (f > 0.0) s_bad = 1; }
        if (f > 0.0) term += f * lnl;
        if (a.siteL && j == a.myShard && s < a.S) {
            long long cnt = 0; double outL = sum;
            if (sum > 0.0) {
                const int e2 = emax;
                cnt = (e2 < 0) ? (long long)((-e2) / 64) : -(long long)((e2 + 63) / 64);
                const int rem = e2 + (int)(64 * cnt);
                outL = sum * exp2i(rem);
            }
            a.siteL[s] = outL; a.siteScale[s] = cnt;
        }
    }
    red[tid] = term;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0) *a.lnL = s_bad ? -INFINITY : red[0];
}

__global__ void __launch_bounds__(256) final_sum_kernel(const double *partial, int n, const int *flag, double *out) {
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (*flag) ? -INFINITY : red[0];
}

}  // namespace hb2
