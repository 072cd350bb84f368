/*
 * hb2_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, fp64) of the reference's likelihood
 * hot path, written from the reference's algorithm, NOT compiled from or copied out of its sources.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this
 * library; the product (hyphy_b200/csrc, libhyphy_b200.so) never links or calls it.
 *
 * Parity pin: tests/test_oracle.py checks every function here against (a) lnL / per-site log-likelihood
 * vectors produced by the UNMODIFIED reference binary (oracle/_ref/hyphy, built by oracle/Makefile.ref) and
 * committed under tests/golden/ by tools/make_golden.py, and (b) the reference's own golden lnL for
 * tests/hbltests/SimpleOptimizations/SmallCodon.bf:37 at its fitted parameters.
 *
 * What follows which reference code (paths relative to /root/reference/src/core):
 *   hb2o_expm            _Matrix::Exponentiate(1., true)                matrix.cpp:5537-5951
 *                        RowAndColumnMax matrix.cpp:4901, IsMaxElement :4984, MinElement :5075
 *   hb2o_prune           _TheTree::ComputeTreeBlockByBranch             tree_evaluator.cpp:3556-4171
 *                        leaf gather / ambiguity                        tree_evaluator.cpp:162-256
 *                        __ll_loop_handle_scaling                       tree_evaluator.cpp:411-525
 *                        _computeBoostScaler/_computeReductionScaler    tree.cpp:161-205, constants :126-129
 *   hb2o_combine         PopulateConditionalProbabilities (weighted sum)  likefunc2.cpp:828-859
 *                        SumUpSiteLikelihoods                           likefunc2.cpp:1446-1506
 *   hb2o_lnl             ComputeBlock + Compute glue                    likefunc.cpp:10783-11289, 2421-2836
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SCALER_UP 18446744073709551616.0          /* 2^64  _lfScalerUpwards            tree.cpp:126 */
#define SCALER_TH (1.0 / 18446744073709551616.0)  /* 2^-64 _lfScalingFactorThreshold   tree.cpp:127 */
#define LOG_SCALER (64.0 * 0.693147180559945309417232121458) /* _logLFScaler tree.cpp:128 */

static double lf_max_scaler(void) { return sqrt(DBL_MAX * 1.e-10); }   /* tree.cpp:129 */
static double lf_min_scaler(void) { return 1.0 / lf_max_scaler(); }

/* ---------------------------------------------------------------------------------------------------- */
/* a8: matrix exponential, reference semantics                                       */
/* ---------------------------------------------------------------------------------------------------- */
static void matmul(const double *A, const double *B, double *C, int D) {
    for (int i = 0; i < D; i++) {
        double *c = C + (size_t)i * D;
        for (int j = 0; j < D; j++) c[j] = 0.0;
        for (int k = 0; k < D; k++) {
            double a = A[(size_t)i * D + k];
            if (a == 0.0) continue;
            const double *b = B + (size_t)k * D;
            for (int j = 0; j < D; j++) c[j] += a * b[j];
        }
    }
}

static int diag_ok(const double *R, int D) {        /* transition_verifier matrix.cpp:5820 */
    for (int r = 0; r < D; r++) if (R[(size_t)r * D + r] > 1.0) return 0;
    return 1;
}

static int diag_repair(double *R, int D) {          /* diag_populator matrix.cpp:5837 */
    for (int r = 0; r < D; r++) {
        double sum = 0.0;
        for (int c = 0; c < D; c++) sum += R[(size_t)r * D + c];
        if (isnan(sum)) return -1;
        R[(size_t)r * D + r] += 1.0 - sum;
    }
    return 0;
}

static int expm_scaled(const double *A, int D, int sparse_storage, double scale_to, double *R) {
    size_t n = (size_t)D * D;
    double *T = (double *)malloc(n * sizeof(double)), *T2 = (double *)malloc(n * sizeof(double));
    double *rs = (double *)calloc(2 * (size_t)D, sizeof(double));
    double *cs = rs + D;
    /* RowAndColumnMax: max absolute row sum x max absolute column sum */
    double minabs = DBL_MAX;
    for (int i = 0; i < D; i++)
        for (int j = 0; j < D; j++) {
            double v = fabs(A[(size_t)i * D + j]);
            rs[i] += v; cs[j] += v;
            /* MinElement runs over *stored* entries: all of them for dense storage, the non-zeros for sparse */
            if ((!sparse_storage || v != 0.0) && v < minabs) minabs = v;
        }
    double r = 0, c = 0;
    for (int i = 0; i < D; i++) { if (rs[i] > r) r = rs[i]; if (cs[i] > c) c = cs[i]; }
    double mx = r * c, mmax = 1.0;
    long power2 = 0;
    if (mx > 0.1) {
        mx = scale_to * (sparse_storage ? 2.0 : 8.0) * sqrt(mx);
        power2 = (long)(log(mx) / log(2.0)) + 1L;          /* C truncation toward zero */
        mmax = exp(power2 * log(2.0));
    }
    /* R = I + A/mmax (power2>0) or I + A */
    double inv = (power2 > 0) ? 1.0 / mmax : 1.0;
    for (size_t k = 0; k < n; k++) R[k] = A[k] * inv;
    for (int d = 0; d < D; d++) R[(size_t)d * D + d] += 1.0;
    int status = 0;
    if (r * c != 0.0) {
        /* Taylor: T_i = T_{i-1} * A scaled so that T_i = (A/mmax)^i / i! */
        double tMax = minabs * sqrt((double)D);
        if (!(tMax > 1e-16)) tMax = 1e-16;               /* truncPrecision matrix.cpp:78 */
        memcpy(T, A, n * sizeof(double));
        long i = 2;
        int more;
        do {
            matmul(T, A, T2, D);
            double f = (i > 2) ? 1.0 / (mmax * (double)i) : 0.5 / (mmax * mmax);
            more = 0;
            for (size_t k = 0; k < n; k++) { T[k] = T2[k] * f; R[k] += T[k]; }
            i++;
            double bench = tMax * 1e-16 * (double)i;
            for (size_t k = 0; k < n; k++) if (T[k] > bench || T[k] < -bench) { more = 1; break; }
        } while (more && i < 10000);
        if (!diag_ok(R, D)) status = 1;                    /* caller restarts with scale_to*100 */
        else if (diag_repair(R, D)) status = -1;
        if (status == 0) {
            double last_diff = 0.0;
            for (long s = 0; s < power2; s++) {            /* Sqr + early exits matrix.cpp:5873-5920 */
                matmul(R, R, T2, D);
                double maxDiff = 0.0;
                for (size_t k = 0; k < n; k++) { double d = fabs(T2[k] - R[k]); if (d > maxDiff) maxDiff = d; R[k] = T2[k]; }
                if (maxDiff < DBL_EPSILON * 1.e3 || (s >= 10 && maxDiff > last_diff * 100.)) break;
                last_diff = maxDiff;
            }
            if (power2) {
                if (!diag_ok(R, D)) status = 1;
                else if (diag_repair(R, D)) status = -1;
            }
        }
    }
    free(T); free(T2); free(rs);
    return status;
}

/* P = exp(A); A = Q*t with diagonal = -rowsum.  sparse_storage mirrors the reference's storage class of the
 * rate matrix (codon models are compressed-sparse there: scale factor 2*sqrt(m) instead of 8*sqrt(m)).
 * returns 0 ok, -1 NaN, -2 could not produce a transition matrix */
int hb2o_expm(const double *A, int D, int sparse_storage, double *P) {
    double scale_to = 1.0;
    for (;;) {
        int st = expm_scaled(A, D, sparse_storage, scale_to, P);
        if (st == 0) return 0;
        if (st < 0) return -1;
        if (scale_to >= 1e100) return -2;
        scale_to *= 100.0;                                  /* matrix.cpp:5854-5864 */
    }
}

/* ---------------------------------------------------------------------------------------------------- */
/* a10 + a11: pruning for one rate class, per-pattern outputs (storageVec mode)                          */
/* ---------------------------------------------------------------------------------------------------- */
static void handle_scaling(double sum, double *parent, int D, double *adj, int64_t *count) {
    if (sum < SCALER_TH && sum > 0.0) {
        double cur = (*adj) * SCALER_UP;
        if (cur < lf_max_scaler()) {                        /* _computeBoostScaler tree.cpp:182 */
            long did = 1;
            double s = sum * SCALER_UP, try2 = cur * SCALER_UP, scaler = SCALER_UP;
            while (s < SCALER_TH && try2 < lf_max_scaler()) { s *= SCALER_UP; try2 *= SCALER_UP; scaler *= SCALER_UP; did++; }
            for (int k = 0; k < D; k++) parent[k] *= scaler;
            *adj *= scaler; *count += did;
        }
    } else if (sum > SCALER_UP && sum < HUGE_VAL) {
        double cur = (*adj) * SCALER_TH;
        if (cur > lf_min_scaler()) {                        /* _computeReductionScaler tree.cpp:161 */
            long did = -1;
            double s = sum * SCALER_TH, try2 = cur * SCALER_TH, scaler = SCALER_TH;
            while (s > SCALER_UP && try2 > lf_min_scaler()) { s *= SCALER_TH; try2 *= SCALER_TH; scaler *= SCALER_TH; did--; }
            for (int k = 0; k < D; k++) parent[k] *= scaler;
            *adj *= scaler; *count += did;
        }
    }
}

/* flatParents: [L+I] parent's internal index, root -1; nodes 0..L-1 leaves (post-order), L.. internals (post-order).
 * leafState [L*S] original pattern order, >=0 state, <0 -> ambig row -(code+1).  P: [(L+I-1)][D*D] row=parent state.
 * Outputs: siteL[S] (root likelihood incl. pi), siteScale[S] with L_true = L * 2^(-64*count).
 * cond (nullable): caller buffer [I*S*D] receiving the internal-node conditionals. */
static int prune_impl(int64_t S, int D, int64_t L, int64_t I, const int64_t *flatParents, const int64_t *leafState,
               const double *ambig, int64_t nAmb, const double *P, const double *pi,
               double *siteL, int64_t *siteScale, double *cond_out, int64_t setBranch, const int64_t *setBranchTo) {
    /* setBranch follows the reference: internal index 0..I-1, or I + leaf index; -1 none (tree_evaluator.cpp:3624,173) */
    size_t nc = (size_t)I * S * D;
    double *cond = cond_out ? cond_out : (double *)malloc(nc * sizeof(double));
    double *adj = (double *)malloc((size_t)I * S * sizeof(double));
    char *touched = (char *)calloc((size_t)I, 1);
    double *mvs = (double *)malloc((size_t)D * sizeof(double));
    if (!cond || !adj || !touched || !mvs) return -1;
    for (size_t k = 0; k < (size_t)I * S; k++) adj[k] = 1.0;       /* likefunc.cpp:4250 */
    for (int64_t s = 0; s < S; s++) siteScale[s] = 0;
    for (int64_t node = 0; node < L + I - 1; node++) {              /* leaves then internals, post-order */
        int64_t par = flatParents[node];
        const double *Pn = P + (size_t)node * D * D;
        if (!touched[par]) {                                         /* tree_evaluator.cpp:3618-3664 */
            touched[par] = 1;
            for (int64_t s = 0; s < S; s++) {
                double *pc = cond + ((size_t)par * S + s) * D;
                if (par == setBranch) {                              /* __ll_loop_handle_leaf_case, matchSet :585-592 */
                    for (int k = 0; k < D; k++) pc[k] = 0.0;
                    pc[setBranchTo[s]] = adj[(size_t)par * S + s];
                } else
                for (int k = 0; k < D; k++) pc[k] = adj[(size_t)par * S + s];
            }
        }
        for (int64_t s = 0; s < S; s++) {
            double *pc = cond + ((size_t)par * S + s) * D;
            const double *child = NULL;
            double sum = 0.0;
            if (node < L) {
                int64_t st = (setBranch == node + I) ? setBranchTo[s] : leafState[(size_t)node * S + s];   /* :173-181 */
                if (st >= 0) {                                       /* column gather tree_evaluator.cpp:171-235 */
                    for (int k = 0; k < D; k++) { pc[k] *= Pn[(size_t)k * D + st]; sum += pc[k]; }
                    handle_scaling(sum, pc, D, &adj[(size_t)par * S + s], &siteScale[s]);
                    continue;
                }
                if (-st - 1 >= nAmb) return -2;
                child = ambig + (size_t)(-st - 1) * D;               /* tree_evaluator.cpp:237 */
            } else {
                child = cond + ((size_t)(node - L) * S + s) * D;
            }
            for (int k = 0; k < D; k++) {
                double a = 0.0;
                const double *row = Pn + (size_t)k * D;
                for (int j = 0; j < D; j++) a += row[j] * child[j];
                mvs[k] = a;
            }
            for (int k = 0; k < D; k++) { pc[k] *= mvs[k]; sum += pc[k]; }
            handle_scaling(sum, pc, D, &adj[(size_t)par * S + s], &siteScale[s]);
        }
    }
    const double *root = cond + (size_t)(I - 1) * S * D;            /* tree_evaluator.cpp:4048-4067 */
    for (int64_t s = 0; s < S; s++) {
        double acc = 0.0;
        if (setBranch + 1 == I) acc = root[(size_t)s * D + setBranchTo[s]] * pi[setBranchTo[s]];     /* :4059-4063 */
        else
        for (int p = 0; p < D; p++) acc += root[(size_t)s * D + p] * pi[p];
        siteL[s] = acc;
    }
    if (!cond_out) free(cond);
    free(adj); free(touched); free(mvs);
    return 0;
}

int hb2o_prune(int64_t S, int D, int64_t L, int64_t I, const int64_t *flatParents, const int64_t *leafState,
               const double *ambig, int64_t nAmb, const double *P, const double *pi,
               double *siteL, int64_t *siteScale, double *cond_out) {
    return prune_impl(S, D, L, I, flatParents, leafState, ambig, nAmb, P, pi, siteL, siteScale, cond_out, -1, NULL);
}

/* ComputeTreeBlockByBranch with setBranch / setBranchTo (forced states of one node, reference convention for setBranch) */
int hb2o_prune_forced(int64_t S, int D, int64_t L, int64_t I, const int64_t *flatParents, const int64_t *leafState,
               const double *ambig, int64_t nAmb, const double *P, const double *pi,
               double *siteL, int64_t *siteScale, int64_t setBranch, const int64_t *setBranchTo) {
    return prune_impl(S, D, L, I, flatParents, leafState, ambig, nAmb, P, pi, siteL, siteScale, NULL, setBranch, setBranchTo);
}

/* ---------------------------------------------------------------------------------------------------- */
/* a13: category combination (weighted sum with 2^64 scaler harmonisation) and the final sum            */
/* ---------------------------------------------------------------------------------------------------- */
static double scaler_multiplier(long s) { return exp(-LOG_SCALER * (double)s); }   /* acquireScalerMultiplier */

void hb2o_combine(int64_t S, int64_t C, const double *weights, const double *siteL /*[C][S]*/,
                  const int64_t *siteScale /*[C][S]*/, double *outL, int64_t *outScale) {
    for (int64_t s = 0; s < S; s++) {
        double buf = 0.0; int64_t sc = 0;
        for (int64_t c = 0; c < C; c++) {
            double v = siteL[(size_t)c * S + s]; int64_t scv = siteScale[(size_t)c * S + s];
            if (c == 0) { buf = weights[c] * v; sc = scv; }
            else if (scv < sc) { buf = weights[c] * v + buf * scaler_multiplier((long)(sc - scv)); sc = scv; }
            else if (scv > sc) { buf += weights[c] * v * scaler_multiplier((long)(scv - sc)); }
            else buf += weights[c] * v;
        }
        outL[s] = buf; outScale[s] = sc;
    }
}

/* lnL = sum f*log(L) - 64 ln2 * sum f*count; -inf if any L <= 0 (tree_evaluator.cpp:4094-4112).  siteLnL nullable. */
double hb2o_sum(int64_t S, const double *L, const int64_t *scale, const int64_t *freq, double *siteLnL) {
    double logL = 0.0, comp = 0.0; int64_t cum = 0; int bad = 0;
    for (int64_t s = 0; s < S; s++) {
        double l;
        if (L[s] > 0.0) l = log(L[s]); else { l = -INFINITY; bad = 1; }
        if (siteLnL) siteLnL[s] = l - LOG_SCALER * (double)scale[s];
        double term = l * (double)freq[s] - comp;               /* compensated, as the reference's Kahan loop */
        double t = logL + term; comp = (t - logL) - term; logL = t;
        cum += scale[s] * freq[s];
    }
    if (bad) return -INFINITY;
    return logL - LOG_SCALER * (double)cum;
}

/* Whole likelihood function for one partition: C rate classes, each with B = L+I-1 rate matrices Q*t.
 * Qt: [C][B][D*D].  Returns lnL; siteLnL (nullable) gets per-pattern log-likelihoods. */
double hb2o_lnl(int64_t S, int D, int64_t L, int64_t I, int64_t C, const int64_t *flatParents,
                const int64_t *leafState, const double *ambig, int64_t nAmb, const int64_t *freq,
                const double *Qt, int sparse_storage, const double *weights, const double *pi, double *siteLnL) {
    int64_t B = L + I - 1;
    size_t dd = (size_t)D * D;
    double *P = (double *)malloc((size_t)B * dd * sizeof(double));
    double *sl = (double *)malloc((size_t)C * S * sizeof(double));
    int64_t *ss = (int64_t *)malloc((size_t)C * S * sizeof(int64_t));
    double *ol = (double *)malloc((size_t)S * sizeof(double));
    int64_t *os = (int64_t *)malloc((size_t)S * sizeof(int64_t));
    double res = NAN;
    int ok = 1;
    for (int64_t c = 0; c < C && ok; c++) {
        for (int64_t b = 0; b < B && ok; b++)
            if (hb2o_expm(Qt + ((size_t)c * B + b) * dd, D, sparse_storage, P + (size_t)b * dd)) ok = 0;
        if (ok && hb2o_prune(S, D, L, I, flatParents, leafState, ambig, nAmb, P, pi, sl + (size_t)c * S, ss + (size_t)c * S, NULL)) ok = 0;
    }
    if (ok) {
        hb2o_combine(S, C, weights, sl, ss, ol, os);
        res = hb2o_sum(S, ol, os, freq, siteLnL);
    }
    free(P); free(sl); free(ss); free(ol); free(os);
    return res;
}
