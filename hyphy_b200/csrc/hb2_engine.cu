// hb2_engine.cu -- C++ host side of libhyphy_b200.so: the C ABI of include/hyphy_b200.h over the CUDA kernels.
//
// Mirrors, for one partition, what the reference does inside _LikelihoodFunction::ComputeBlock
// (likefunc.cpp:10783-11289): keep the per-node caches resident (here: in HBM), take the changed matrices,
// re-exponentiate them, re-prune the dirty part of the tree, reduce at the root, hand back one double.
// No CPU fallback exists on this path: every failure is an error code (fatal for the host).
#include "../../include/hyphy_b200.h"
#include "hb2_kernels_fp64.cuh"
#include "hb2_kernels_tc.cuh"
#include "hb2_kernels_lanes.cuh"

#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>
#include <unordered_map>

#include <nccl.h>

namespace {

thread_local std::string g_err;

int fail(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CU(x)                                                                                         \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ---- Block pool for device and pinned-host memory ----------------------------------------------------------------------
// The site phases of FEL / MEME create and destroy one small likelihood function per site: through the
// patched host that is one hb2_create / hb2_destroy pair each, i.e. ~60 cudaMalloc / cudaMallocHost and as many frees (each
// cudaFree synchronises the device).  Blocks up to 32 MB are rounded up to a power of two and recycled through a
// process-wide free list per device (caps: 1 GB device, 256 MB pinned); larger blocks go straight to the runtime.
// Recycled memory is NOT zero: every buffer whose initial contents matter is cleared explicitly at hb2_create, as it
// always had to be (cudaMalloc gives no such guarantee either).  HB2_POOL=0 switches the pool off.
struct BlockPool {
    struct Info { size_t cls; int dev; bool host; };
    std::mutex mu;
    std::unordered_map<void *, Info> live;
    std::unordered_map<unsigned long long, std::vector<void *>> free_blocks;      // key = (host, device, class)
    size_t pooled_dev = 0, pooled_host = 0;
    static constexpr size_t kMaxClass = (size_t)32 << 20, kDevCap = (size_t)1 << 30, kHostCap = (size_t)256 << 20;
    bool enabled() { static const bool on = !(getenv("HB2_POOL") && getenv("HB2_POOL")[0] == '0'); return on; }
    static size_t size_class(size_t n) { size_t c = 256; while (c < n) c <<= 1; return c; }
    static unsigned long long key(bool host, int dev, size_t cls) { return ((unsigned long long)host << 63) | ((unsigned long long)(dev & 0xff) << 52) | (unsigned long long)cls; }
    cudaError_t alloc(void **p, size_t n, bool host) {
        if (!enabled() || n > kMaxClass) return host ? (cudaMallocHost)(p, n) : (cudaMalloc)(p, n);
        int dev = 0;
        cudaGetDevice(&dev);
        const size_t cls = size_class(n);
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = free_blocks.find(key(host, host ? 0 : dev, cls));
            if (it != free_blocks.end() && !it->second.empty()) {
                *p = it->second.back();
                it->second.pop_back();
                (host ? pooled_host : pooled_dev) -= cls;
                live[*p] = Info{cls, dev, host};
                return cudaSuccess;
            }
        }
        const cudaError_t e = host ? (cudaMallocHost)(p, cls) : (cudaMalloc)(p, cls);
        if (e == cudaSuccess) { std::lock_guard<std::mutex> lk(mu); live[*p] = Info{cls, dev, host}; }
        return e;
    }
    cudaError_t release(void *p, bool host) {
        if (!p) return cudaSuccess;
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = live.find(p);
            if (it != live.end()) {
                const Info in = it->second;
                live.erase(it);
                size_t &pooled = in.host ? pooled_host : pooled_dev;
                if (pooled + in.cls <= (in.host ? kHostCap : kDevCap)) {
                    free_blocks[key(in.host, in.host ? 0 : in.dev, in.cls)].push_back(p);
                    pooled += in.cls;
                    return cudaSuccess;
                }
            }
        }
        return host ? (cudaFreeHost)(p) : (cudaFree)(p);
    }
};
BlockPool g_pool;
// every allocation of this file goes through the pool (buffers that are exported through CUDA IPC opt out with the
// parenthesised runtime name)
#define cudaMalloc(p, n) g_pool.alloc((void **)(p), (n), false)
#define cudaMallocHost(p, n) g_pool.alloc((void **)(p), (n), true)
#define cudaFree(p) g_pool.release((void *)(p), false)
#define cudaFreeHost(p) g_pool.release((void *)(p), true)

// ---- NCCL through dlopen: single-GPU users never need the library -------------------------------
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (h) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) return false;
        GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
        AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        return GetUniqueId && CommInitRank && AllReduce && AllGather && CommDestroy && GetErrorString;
    }
} g_nccl;

int pad_states(int64_t D) {
    if (D <= 4) return 4;
    if (D <= 8) return 8;
    if (D <= 16) return 16;
    if (D <= 24) return 24;
    if (D <= 32) return 32;
    if (D <= 64) return 64;
    return -1;
}

}  // namespace

struct hb2_partition {
    int device = 0, flags = 0;
    cudaStream_t stream = nullptr;
    int64_t S = 0, D = 0, L = 0, I = 0, C = 0, B = 0, nAmb = 0;
    int Dp = 0;
    int64_t Sp = 0;
    std::vector<int64_t> parents;
    std::vector<std::vector<int>> children;   // per internal node, flat ids
    std::vector<int> height;                  // per internal node: 0 = only leaf children
    int max_height = 0;
    // device
    int *d_leaf = nullptr, *d_scal = nullptr, *d_rootE = nullptr, *d_child_start = nullptr, *d_child_ids = nullptr;
    int *d_jobs = nullptr, *d_dst = nullptr, *d_flag = nullptr;
    double *d_ambig = nullptr, *d_freq = nullptr, *d_cond = nullptr, *d_PT = nullptr, *d_Q = nullptr, *d_pi = nullptr;
    double *d_rootL = nullptr, *d_weights = nullptr, *d_partial = nullptr, *d_lnL = nullptr, *d_siteL = nullptr;
    long long *d_siteScale = nullptr;
    // pinned host staging
    double *h_Q = nullptr, *h_small = nullptr;   // h_small: pi (Dp) + weights (C) + lnL (1)
    int *h_jobs = nullptr, *h_dst = nullptr;
    // pending matrices: entries [0, n_pending) of h_Q / h_dst, with kinds
    int64_t n_pending = 0;
    std::vector<int> pending_kind;
    // one pending entry per (class, node) slot across BOTH queues: >= 0 index into the dense queue, <= -2 -> index
    // -(v+2) into the compiled queue, -1 none.  A slot handed over twice before an evaluation overwrites / retires its
    // earlier entry (last one wins, like the reference's SetCompExp, calcnode.cpp:714), so one expm launch never holds
    // two CTAs with the same destination and the flush order of the two queues cannot matter.
    std::vector<int> pend_pos;
    int64_t q_capacity = 0;                   // in matrices
    std::vector<char> have_matrix;            // [C*B]
    std::vector<char> is_rate;                // [C*B] slot holds a rate matrix resident in d_Qres (for hb2_time_resident)
    double *d_Qres = nullptr;                 // [C][B][D*D] last rate matrices, resident copy
    // explicit-form mixtures (hb2_set_mixture_matrices): own staging so that nothing is shared with the plain queues
    double *d_mix_scratch = nullptr, *d_mix_Q = nullptr, *h_mix = nullptr;   // [cap][Dp*Dp], [cap][D*D], pinned [cap][D*D + 1]
    int *d_mix_dst = nullptr;
    int64_t mix_capacity = 0;
    cudaEvent_t ev_mix = nullptr;
    bool mix_busy = false;
    // shared-powers expm (64 padded states, hb2_kernels_fp64.cuh "Shared powers"): group = rate class (plain batches) or
    // C + mixture component; per-entry arrays are indexed like the batch
    struct ExBuf {
        int G = 0, stride = 0;                // groups, doubles per cached reference direction
        int *d_int = nullptr, *h_int = nullptr, *d_flag = nullptr;     // [group cap | ref cap | refs G]
        double *d_weight = nullptr, *d_pow = nullptr, *d_refvec = nullptr, *d_colsum = nullptr;   // colsum [cap][16][64]
        void *d_groups = nullptr;
        unsigned long long *d_flags = nullptr, gen = 0;   // generation flags of the coefficient matrices [G][EXPM_POW_TERMS]
        std::vector<int> kind;                // host mirror: kind of the reference direction each group holds (0 none)
        int64_t cap = 0;
    };
    ExBuf ex;                                 // the partition's own hand-overs
    ExBuf bex;                                // batched one-pattern likelihoods (hb2_batch_site_likelihoods): one group per set
    int64_t b_cap = 0;                        // sets per chunk of the batch buffers below
    double *d_bPT = nullptr, *d_bV = nullptr, *d_bcond = nullptr, *d_bout = nullptr;
    int *d_bdst = nullptr, *d_bpat = nullptr, *d_bnodeex = nullptr, *h_bdst = nullptr;
    int b_groups_per_set = 1;
    bool ex_enabled = true;                   // HB2_EXPM_SHARED=0 switches the path off (A/B testing)
    int64_t stage_launches[3] = {0, 0, 0};    // launches per evaluation of the last hb2_time_resident {expm, pruning, root}
    // compiled rate-matrix templates (hb2_set_rate_template*): static scatter map + own queue of per-evaluation formula
    // values.  A tree may carry several models (foreground / background branches, per-branch models of aBSREL): one
    // template each, up to HB2_MAX_TEMPLATES.
    struct Tmpl {
        int64_t nnz = 0, nF = 0, n_pending = 0;
        bool has_colfreq = false;
        int *d_index = nullptr, *d_formula = nullptr, *d_vdst = nullptr, *h_vdst = nullptr;
        double *d_colfreq = nullptr, *h_colfreq = nullptr, *d_V = nullptr, *h_V = nullptr, *d_Vres = nullptr;
    };
    std::vector<Tmpl> tmpls;
    std::vector<signed char> pend_tmpl;       // [C*B] template of the slot's pending compiled entry
    std::vector<signed char> res_tmpl;        // [C*B] template of the slot's resident formula values (is_rate == 2)
    // tensor-core path (33..64 states unless HB2_FLAG_FORCE_FP64): fp32 conditionals + split/tiled P operands
    bool use_tc = false;
    float *d_condf = nullptr, *d_PB = nullptr, *d_PTf = nullptr;
    int *d_err = nullptr;
    unsigned *d_root_counter = nullptr;       // ticket counter of the fused root reduction (combine_kernel's last block)
    int *d_forced = nullptr, *h_forced = nullptr;   // forced states of the current evaluation (hb2_evaluate_forced)
    int forced_node = -1;
    // persistent walk kernel (one launch per evaluation): plan buffers, generation bits of the tagged hand-over, residency
    bool small_walk = true;                   // HB2_SMALL_WALK=0: per-level launches of prune_small_kernel (A/B testing)
    int small_ilp = 0, small_resident[2] = {1, 1};   // HB2_SMALL_ILP=1/2 forces one / two patterns per thread (4 states); co-resident CTAs of both
    bool small_dmma = true;                   // HB2_SMALL_DMMA=0: one thread per pattern (prune_small_walk_kernel) also for 16..32 states
    int fp64_mode = 2;                        // HB2_FP64_WALK: 2 (default) prune64_lanes_kernel, 1 prune64_walk_kernel when it fills the machine, 0 per-level launches
    bool fp64_walk = true;                    // fp64_mode >= 1
    double *d_cond_side = nullptr;            // fp64 lanes kernel: side products [C][I][Sp][64] (exponents: second half of d_scal)
    int *d_lane_flags = nullptr;              // [C][2I][Sp/64] pass id of the last production
    int lane_pass = 0, lanes_resident[2] = {0, 0}, lanes_force_nw = 0;   // co-resident CTAs of the 4- and 8-warp shapes; HB2_LANES_WARPS
    int sm_count = 148;
    bool expm_dfma = false;                   // HB2_EXPM_DFMA=1: previous FFMA-style fp64 expm kernel (A/B testing)
    bool use_walk = false;
    bool walk_v2 = false;                     // HB2_WALK_V2=1: two threads per pattern (prune64_tc_walk2_kernel; correct, measured 20 % slower)
    int walk_max_resident = 0;
    // single-branch shortcut (hb2_branch_cache_*): outside vectors of one branch, all owned classes
    double *d_bc_out = nullptr; int *d_bc_outE = nullptr, *d_bc_sib = nullptr; int64_t bc_node = -1;
    int64_t bc_dirty_node = -1;               // first node other than bc_node whose matrix changed since the build
    std::vector<char> plan_jdirty;      // dirty jobs (nodes + side products) of the cached plan
    bool walk_split_nodes = true;       // HB2_WALK_SPLIT_NODES=0: no side products (A/B)
    std::vector<char> plan_dirty;       // cached walk plan (h_walk holds its steps): dirty set, lane count, step count
    int plan_K = 0, plan_steps = 0;
    std::vector<int> walk_gen;          // [C][I] generation bit of each node's resident conditionals (tc walk path)
    bool walk_reset = false;            // a pass was aborted: tags are inconsistent -> wipe and recompute everything
    int *d_walk = nullptr, *h_walk = nullptr;
    int walk_lane_cap = 8;              // lanes (CTAs) that share one (class, tile) pair; HB2_WALK_LANES overrides (<= 15)
    bool first_eval_done = false;
    std::vector<char> evaluated_cat;          // [C] whole tree pruned at least once
    int64_t launches = 0;
    ncclComm_t comm = nullptr;
    int n_ranks = 1, rank = 0;
    // class groups (hb2_comm_class_groups): this rank prunes classes [own0, own0 + ownN) only
    int cg_G = 1, cg_g = 0, own0 = 0, ownN = 0, xchg_len = 0;
    double *d_xsend = nullptr, *d_xrecv = nullptr, *d_xfreq = nullptr, *d_xpartial = nullptr;
    int n_partial_blocks = 0;
    // peer exchange (hb2_kernels_fp64.cuh "Peer exchange"): IPC-mapped buffers of all ranks; falls back to NCCL when the
    // mapping cannot be established (px_ok false)
    bool px_ok = false, px_enabled = true;
    int px_payload = 0;
    unsigned long long px_gen = 0;
    double *px_local = nullptr;
    std::vector<double *> px_peer;            // host copy of the mapped pointers (entry rank = px_local)
    double **d_px_peer = nullptr;
    unsigned int *d_px_counter = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_staging = nullptr;
    bool staging_busy = false;
};

namespace {

// The pinned staging buffers (both queues, h_dst/h_vdst included) may still be read by the H2D copies of the last flush.
int wait_staging(hb2_partition *p) {
    if (p->staging_busy) {
        CU(cudaEventSynchronize(p->ev_staging));
        p->staging_busy = false;
    }
    return 0;
}

// Host half of the shared-powers path for one batch: groups and references of the n entries (h_dst = their slots, < 0
// retired; kind 1 = compiled hand-over, 2 = dense).  A group whose batch holds at least EX_MIN_GROUP live entries gets its
// power table rebuilt from the first of them; smaller groups are compared with the direction the table was last built
// from.  Fills the pinned arrays and enqueues their upload: call BEFORE ev_staging is recorded.
struct SharedPlan {
    bool use = false;
    int kind = 0, n_refs = 0, tmpl = 0;
    hb2_partition::ExBuf *x = nullptr;
};
constexpr int EX_MIN_GROUP = 24;

int plan_shared(hb2_partition *p, const int *h_dst, int64_t n, int kind, const int *group_override, SharedPlan &sp, int tmpl = 0,
                hb2_partition::ExBuf *bx = nullptr, int min_group = EX_MIN_GROUP) {
    sp = SharedPlan();
    hb2_partition::ExBuf &x = bx ? *bx : p->ex;
    sp.x = &x;
    if (p->Dp != 64 || p->expm_dfma || !p->ex_enabled || !x.d_int || n <= 0 || n > x.cap) return 0;
    if (kind == 1 && !p->tmpls[tmpl].d_Vres) return 0;
    if (wait_staging(p)) return 1;            // an earlier plan's upload may still be reading the pinned arrays
    int *grp = x.h_int, *ref = x.h_int + x.cap, *refs = x.h_int + 2 * x.cap;
    std::vector<int> count(x.G, 0), first(x.G, -1);
    for (int64_t k = 0; k < n; k++) {
        int g = -1;
        if (h_dst[k] >= 0) g = group_override ? group_override[k] : (int)(h_dst[k] / p->B) + tmpl * (int)p->C;   // (template, class)
        if (g >= x.G) g = -1;
        grp[k] = g;
        if (g >= 0) { if (first[g] < 0) first[g] = (int)k; count[g]++; }
    }
    bool any = false;
    int nr = 0;
    for (int g = 0; g < x.G; g++) {
        if (count[g] >= min_group) { refs[nr++] = first[g]; any = true; }
        else { first[g] = -1; if (count[g] > 0 && x.kind[g] == kind) any = true; }
    }
    if (!any) return 0;
    for (int64_t k = 0; k < n; k++) ref[k] = grp[k] >= 0 ? first[grp[k]] : -1;
    for (int r = 0; r < nr; r++) x.kind[grp[refs[r]]] = kind;
    CU(cudaMemcpyAsync(x.d_int, x.h_int, (size_t)(2 * x.cap + x.G) * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    sp.use = true; sp.kind = kind; sp.n_refs = nr; sp.tmpl = tmpl;
    return 0;
}

// One CTA per listed matrix.  is_trans: 0 rate matrices (exponentiated), 1 transition matrices (transposed/padded only),
// 2 compiled template (dQ = formula values).  pt_override: write the results there instead of the P cache (mixture
// components go to a scratch area first); pack_tc: also emit the tensor-path operands of the slot.  sp: shared-powers plan
// of the WHOLE batch this run [off, off + n) belongs to (dQ / d_dst already point at the run).
int launch_expm(hb2_partition *p, const double *dQ, const int *d_dst, int n, int is_trans, double *qres, bool pack_tc = true,
                double *pt_override = nullptr, const SharedPlan *sp = nullptr, int64_t off = 0, int tmpl = 0) {
    if (n <= 0) return 0;
    hb2::ExpmArgs a{};
    a.Q = dQ; a.dst = d_dst; a.PT = pt_override ? pt_override : p->d_PT; a.Qres = qres; a.D = (int)p->D;
    a.is_trans = is_trans;
    if (is_trans == 2) {                       // compiled template: dQ points at the formula values [n][nF]
        const hb2_partition::Tmpl &t = p->tmpls[tmpl];
        a.is_trans = 0; a.Q = nullptr; a.V = dQ; a.tmpl_index = t.d_index; a.tmpl_formula = t.d_formula;
        a.tmpl_colfreq = t.has_colfreq ? t.d_colfreq : nullptr; a.tmpl_nnz = (int)t.nnz; a.nF = (int)t.nF;
    }
    bool packed = false;
    pack_tc = pack_tc && p->use_tc && !pt_override;
    switch (p->Dp) {
        case 64:
            if (p->expm_dfma) {
                hb2::expm64_kernel<<<n, 256, 5 * 64 * hb2::LD64 * sizeof(double), p->stream>>>(a);
            } else {
                const size_t smem = 3 * 64 * hb2::LD64 * sizeof(double);
                if (sp && sp->use && is_trans != 1) {
                    hb2_partition::ExBuf &x = *sp->x;
                    const int nV = sp->kind == 1 ? (int)p->tmpls[tmpl].nF : (int)(p->D * p->D);
                    const hb2::ExpmGroup *groups = static_cast<const hb2::ExpmGroup *>(x.d_groups);
                    hb2::ExpmClassifyArgs ca{};
                    ca.V = dQ; ca.nV = nV; ca.kind = sp->kind; ca.dst = d_dst; ca.group = x.d_int + off;
                    ca.ref = x.d_int + x.cap + off; ca.groups = groups; ca.refvec = x.d_refvec; ca.refvec_stride = x.stride;
                    ca.weight = x.d_weight + off; ca.flag = x.d_flag + off;
                    ca.res = pt_override ? nullptr : (sp->kind == 1 ? p->tmpls[tmpl].d_Vres : qres);   // scratch targets keep no resident copy
                    // references index the whole batch: classification has to see it in one piece (off == 0 for compiled batches
                    // and for dense batches that consist of a single run; mixed dense batches fall back below)
                    static const bool warp_classify = !(getenv("HB2_CLASSIFY_WARP") && getenv("HB2_CLASSIFY_WARP")[0] == '0');
                    if (nV <= 256 && warp_classify) hb2::expm_classify_warp_kernel<<<(n + 7) / 8, 256, 0, p->stream>>>(ca, n);
                    else hb2::expm_classify_kernel<<<n, 128, 0, p->stream>>>(ca);
                    a.group = ca.group; a.flag = ca.flag; a.weight = ca.weight; a.groups = groups; a.pow = x.d_pow;
                    a.Qres = nullptr;                    // the classification stage keeps the resident copy
                    if (sp->n_refs > 0) {
                        hb2::ExpmPowersArgs pa{};
                        pa.a = a; pa.refs = x.d_int + 2 * x.cap; pa.groups = static_cast<hb2::ExpmGroup *>(x.d_groups);
                        pa.pow = x.d_pow; pa.refvec = x.d_refvec; pa.refvec_stride = x.stride; pa.nV = nV; pa.kind = sp->kind; pa.Vin = dQ;
                        pa.flags = x.d_flags; pa.gen = ++x.gen; pa.err = p->d_err;
                        hb2::expm_powers_kernel<<<dim3(hb2::EXPM_POW_TERMS - 1, (unsigned)sp->n_refs), 256, smem, p->stream>>>(pa);
                        p->launches++;
                    }
                    hb2::ExpmTcOut tcs{nullptr, nullptr};
                    if (pack_tc) { tcs.PB = p->d_PB; tcs.PTf = p->d_PTf; }
                    double *colsum = x.d_colsum + (size_t)off * 16 * 64;
                    hb2::expm_poly_kernel<<<dim3(16, (unsigned)((n + hb2::EXPM_POLY_CHUNK - 1) / hb2::EXPM_POLY_CHUNK)), 256, 0, p->stream>>>(a, tcs, colsum, n);
                    // the row repair of the entries the polynomial kernel finished rides in the finishing kernel below (same
                    // arithmetic as expm_diag_kernel, one launch less); needs the tensor-path outputs to be the same set
                    static const bool fuse_diag = !(getenv("HB2_EXPM_FUSE_DIAG") && getenv("HB2_EXPM_FUSE_DIAG")[0] == '0');
                    if (fuse_diag) { a.colsum = colsum; p->launches += 2; }
                    else { hb2::expm_diag_kernel<<<n, 64, 0, p->stream>>>(a, tcs, colsum); p->launches += 3; }
                }
                hb2::ExpmTcOut tco{nullptr, nullptr};
                if (pack_tc) { tco.PB = p->d_PB; tco.PTf = p->d_PTf; packed = true; }
                hb2::expm64_dmma_kernel<<<n, 256, smem, p->stream>>>(a, tco);
            }
            break;
        case 4: hb2::expm_small_kernel<4><<<n, 128, hb2::expm_small_smem_bytes(4), p->stream>>>(a); break;
        case 8: hb2::expm_small_kernel<8><<<n, 128, hb2::expm_small_smem_bytes(8), p->stream>>>(a); break;
        case 16: hb2::expm_small_kernel<16><<<n, 128, hb2::expm_small_smem_bytes(16), p->stream>>>(a); break;
        case 24: hb2::expm_small_kernel<24><<<n, 128, hb2::expm_small_smem_bytes(24), p->stream>>>(a); break;
        case 32: hb2::expm_small_kernel<32><<<n, 128, hb2::expm_small_smem_bytes(32), p->stream>>>(a); break;
        default: return fail("unsupported padded state count %d", p->Dp);
    }
    p->launches++;
    CU(cudaGetLastError());
    if (pack_tc && !packed) {
        hb2::pack_tc_kernel<<<n, 256, 0, p->stream>>>(p->d_PT, d_dst, p->d_PB, p->d_PTf);
        p->launches++;
        CU(cudaGetLastError());
    }
    return 0;
}

int launch_prune(hb2_partition *p, const hb2::PruneArgs &a, const int *d_jobs, int njobs, int ncls) {
    if (njobs <= 0) return 0;
    if (p->use_tc) {
        hb2::PruneTcArgs t;
        t.PB = p->d_PB; t.PTf = p->d_PTf; t.cond = p->d_condf; t.scal = a.scal; t.leaf = a.leaf; t.ambig = a.ambig; t.pi = a.pi;
        t.rootL = a.rootL; t.rootE = a.rootE; t.tree = a.tree; t.err = p->d_err;
        t.L = a.L; t.I = a.I; t.B = a.B; t.D = a.D; t.Sp = a.Sp; t.cat0 = a.cat0;
        t.forced = a.forced; t.forced_node = a.forced_node;
        dim3 grid((unsigned)(p->Sp / hb2::TC_TILE_P), (unsigned)njobs, (unsigned)ncls);
        hb2::prune64_tc_kernel<<<grid, 128, hb2::TC_SMEM_BYTES, p->stream>>>(t, d_jobs);
    } else if (p->Dp == 64) {
        dim3 grid((unsigned)(p->Sp / hb2::TILE_P), (unsigned)njobs, (unsigned)ncls);
        hb2::prune64_kernel<<<grid, 256, 2 * 64 * hb2::LD64 * sizeof(double), p->stream>>>(a, d_jobs);
    } else {
        dim3 grid((unsigned)(p->Sp / 128), (unsigned)njobs, (unsigned)ncls);
        switch (p->Dp) {
            case 4: hb2::prune_small_kernel<4><<<grid, 128, 0, p->stream>>>(a, d_jobs); break;
            case 8: hb2::prune_small_kernel<8><<<grid, 128, 0, p->stream>>>(a, d_jobs); break;
            case 16: hb2::prune_small_kernel<16><<<grid, 128, 0, p->stream>>>(a, d_jobs); break;
            case 24: hb2::prune_small_kernel<24><<<grid, 128, 0, p->stream>>>(a, d_jobs); break;
            case 32: hb2::prune_small_kernel<32><<<grid, 128, 0, p->stream>>>(a, d_jobs); break;
            default: return fail("unsupported padded state count %d", p->Dp);
        }
    }
    p->launches++;
    CU(cudaGetLastError());
    return 0;
}

// Flush matrices handed over since the last evaluation: one H2D copy + one (or two) expm launches per queue.  Retired
// entries (destination -1: the slot was handed over again later) are skipped by the kernels.
int flush_compiled(hb2_partition *p) {
    for (size_t ti = 0; ti < p->tmpls.size(); ti++) {
        hb2_partition::Tmpl &t = p->tmpls[ti];
        const int64_t n = t.n_pending;
        if (n == 0) continue;
        CU(cudaMemcpyAsync(t.d_V, t.h_V, n * t.nF * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        CU(cudaMemcpyAsync(t.d_vdst, t.h_vdst, n * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        SharedPlan sp;
        if (plan_shared(p, t.h_vdst, n, 1, nullptr, sp, (int)ti)) return 1;
        CU(cudaEventRecord(p->ev_staging, p->stream));
        p->staging_busy = true;
        if (launch_expm(p, t.d_V, t.d_vdst, (int)n, 2, sp.use ? nullptr : p->d_Qres, true, nullptr, &sp, 0, (int)ti)) return 1;
        for (int64_t k = 0; k < n; k++)
            if (t.h_vdst[k] >= 0) { p->is_rate[t.h_vdst[k]] = sp.use ? 2 : 1; p->res_tmpl[t.h_vdst[k]] = (signed char)ti; p->pend_pos[t.h_vdst[k]] = -1; }
        t.n_pending = 0;
    }
    return 0;
}

int flush_matrices(hb2_partition *p) {
    if (flush_compiled(p)) return 1;
    const int64_t n = p->n_pending;
    if (n == 0) return 0;
    const size_t dd = (size_t)p->D * p->D;
    CU(cudaMemcpyAsync(p->d_Q, p->h_Q, n * dd * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CU(cudaMemcpyAsync(p->d_dst, p->h_dst, n * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    SharedPlan sp;                           // shared powers only for batches that are one run of rate matrices
    {
        bool all_rate = true;
        for (int64_t k = 0; k < n; k++) all_rate = all_rate && p->pending_kind[k] == HB2_MATRIX_RATE;
        if (all_rate && plan_shared(p, p->h_dst, n, 2, nullptr, sp)) return 1;
    }
    // the pinned staging buffers are reused by the next hb2_set_matrices: it waits on this event (copies only)
    CU(cudaEventRecord(p->ev_staging, p->stream));
    p->staging_busy = true;
    // pending entries are grouped by kind in runs; one launch per run.  Rate matrices also leave a resident copy
    // in d_Qres (written by the kernel while it loads them) so a whole evaluation can be replayed on device.
    int64_t i = 0;
    while (i < n) {
        int64_t j = i;
        while (j < n && p->pending_kind[j] == p->pending_kind[i]) j++;
        const int is_trans = p->pending_kind[i] == HB2_MATRIX_TRANS;
        if (launch_expm(p, p->d_Q + i * dd, p->d_dst + i, (int)(j - i), is_trans, is_trans ? nullptr : p->d_Qres, true, nullptr, &sp, i)) return 1;
        i = j;
    }
    for (int64_t k = 0; k < n; k++)
        if (p->h_dst[k] >= 0) { p->is_rate[p->h_dst[k]] = (p->pending_kind[k] == HB2_MATRIX_RATE); p->pend_pos[p->h_dst[k]] = -1; }
    p->n_pending = 0;
    p->pending_kind.clear();
    return 0;
}

// A matrix of `node` is about to change: the single-branch cache (if any) only survives changes of ITS branch.
inline void note_matrix_change(hb2_partition *p, int64_t node) {
    if (p->bc_node >= 0 && node != p->bc_node && p->bc_dirty_node < 0) p->bc_dirty_node = node;
}

// Retire the pending entry of `slot` if it sits in ANOTHER queue than the one it is about to join (tmpl < 0: the dense
// queue; >= 0: that template's compiled queue).  Returns the index of an entry of the SAME queue that can be overwritten
// in place, or -1.
int64_t claim_slot(hb2_partition *p, int64_t slot, int tmpl) {
    const int pos = p->pend_pos[slot];
    if (pos == -1) return -1;
    if (pos >= 0) {                            // pending in the dense queue
        if (tmpl < 0) return pos;
        p->h_dst[pos] = -1;
    } else {                                   // pending in a compiled queue
        const int pt = p->pend_tmpl[slot];
        if (pt == tmpl) return -(int64_t)pos - 2;
        p->tmpls[pt].h_vdst[-(int64_t)pos - 2] = -1;
    }
    p->pend_pos[slot] = -1;
    return -1;
}

int stage_matrix(hb2_partition *p, int64_t cat, int64_t node, const double *M, int kind) {
    if (cat < 0) cat = 0;
    if (cat >= p->C) return fail("rate class %lld out of range (C=%lld)", (long long)cat, (long long)p->C);
    if (node < 0 || node >= p->B) return fail("node id %lld has no branch (valid 0..%lld)", (long long)node, (long long)p->B - 1);
    if (kind != HB2_MATRIX_RATE && kind != HB2_MATRIX_TRANS) return fail("unknown matrix kind %d", kind);
    note_matrix_change(p, node);
    if (cat < p->own0 || cat >= p->own0 + p->ownN) { p->have_matrix[cat * p->B + node] = 1; return 0; }   // another class group's
    if (wait_staging(p)) return 1;            // previous flush's H2D copies must have left the pinned buffers
    const int64_t slot = cat * p->B + node;
    int64_t at = claim_slot(p, slot, -1);
    if (at < 0) {
        if (p->n_pending == p->q_capacity) {
            cudaSetDevice(p->device);
            if (flush_matrices(p) || wait_staging(p)) return 1;
        }
        at = p->n_pending++;
        p->pending_kind.push_back(kind);
        p->pend_pos[slot] = (int)at;
    }
    const size_t dd = (size_t)p->D * p->D;
    memcpy(p->h_Q + at * dd, M, dd * sizeof(double));
    p->h_dst[at] = (int)slot;
    p->pending_kind[at] = kind;
    p->have_matrix[slot] = 1;
    return 0;
}

// Build the per-level job lists for a set of dirty nodes (closure: parents of dirty nodes and all their ancestors).
// update == nullptr => whole tree.  Returns levels[h] = internal indices at height h to recompute.
int plan_levels(hb2_partition *p, int64_t nUpdate, const int64_t *update, std::vector<std::vector<int>> &levels) {
    levels.assign(p->max_height + 1, {});
    std::vector<char> dirty(p->I, 0);
    if (update == nullptr || nUpdate < 0) {
        std::fill(dirty.begin(), dirty.end(), 1);
    } else {
        for (int64_t k = 0; k < nUpdate; k++) {
            int64_t n = update[k];
            if (n < 0 || n >= p->L + p->I) return fail("updateNodes[%lld]=%lld out of range", (long long)k, (long long)n);
            int64_t par = p->parents[n];
            while (par >= 0 && !dirty[par]) {
                dirty[par] = 1;
                par = p->parents[p->L + par];
            }
        }
    }
    for (int64_t i = 0; i < p->I; i++)
        if (dirty[i]) levels[p->height[i]].push_back((int)i);
    return 0;
}

hb2::PruneArgs prune_args(hb2_partition *p, int cat0) {
    hb2::PruneArgs a;
    a.PT = p->d_PT; a.cond = p->d_cond; a.scal = p->d_scal; a.leaf = p->d_leaf; a.ambig = p->d_ambig; a.pi = p->d_pi;
    a.rootL = p->d_rootL; a.rootE = p->d_rootE;
    a.tree.child_start = p->d_child_start; a.tree.child_ids = p->d_child_ids;
    a.L = (int)p->L; a.I = (int)p->I; a.B = (int)p->B; a.D = (int)p->D; a.Sp = (int)p->Sp; a.cat0 = cat0;
    a.forced = p->d_forced; a.forced_node = p->forced_node;
    return a;
}

// Pure host code (no CUDA): the walk plan for a dirty set on K lanes.  lane_start[K+1]; steps = 2 ints per step
// (child encoding, job's node slot | flags); jdirty[2I] = jobs of the plan.  Returns the number of steps.  Exposed for
// CPU tests through hb2_plan_walk.
int plan_walk(const std::vector<std::vector<int>> &children, const std::vector<int> &height, int L, int I,
              const std::vector<char> &dirty, int K, bool split_nodes, int *lane_start, int *steps, std::vector<char> &jdirty_out,
              bool canonical_multi = false) {
    // canonical_multi (fp64 lanes kernel): a node with more than two children is neither split nor chained and its children
    // are multiplied in tree order, so its product is associated the same way in every plan; with two children the product
    // is exact-commutative and power-of-two rescalings are exact, hence conditionals do not depend on the plan at all (a
    // partial re-evaluation reproduces the full one bit for bit, like the per-level kernel).
    // Jobs.  Job j < I is node j; job I + n is the SIDE PRODUCT of node n: a node with two or more internal children is
    // split into "contract the child with the deepest subtree" (+ multiply the side product in: no matrix, the cheapest
    // step after a leaf) and a side job that contracts all OTHER children.  Side jobs sit off the root path, so other
    // lanes do them early and the spine of a deep tree carries one contraction per node instead of two (`split`).
    // item list of a job: children to contract (flat ids) and, for a split node, the side job to multiply in.
    const int NJ = 2 * I;
    const bool split_on = split_nodes && K > 1;
    std::vector<std::vector<int>> jch(NJ);       // children (flat ids, leaves and internals) contracted by the job
    std::vector<int> jmul(NJ, -1);               // side job multiplied in by the job (real split nodes only)
    std::vector<char> jdirty(NJ, 0);
    for (int n = 0; n < I; n++) {
        if (!dirty[n]) continue;
        jdirty[n] = 1;
        int heavy = -1, nint = 0;
        for (int ch : children[n]) if (ch >= L) { nint++; if (heavy < 0 || height[ch - L] > height[heavy - L]) heavy = ch; }
        if (split_on && nint >= 2 && !(canonical_multi && children[n].size() > 2)) {
            jch[n].push_back(heavy);
            jmul[n] = I + n;
            jdirty[I + n] = 1;
            for (int ch : children[n]) if (ch != heavy) jch[I + n].push_back(ch);
        } else {
            jch[n] = children[n];
        }
    }
    // Lane assignment = list scheduling of the dirty jobs on K in-order lanes with a rough cost model (cycles measured
    // with HB2_WALK_TRACE, profiles/): the pass is bound by the busiest lane or by the deepest root path, so the
    // (job, lane) pair that can start first is scheduled next, ties going to the longer remaining path; a job whose
    // contracted internal child was the lane's previous job continues there (register hand-over: the cheapest
    // contraction).  Every lane executes its jobs in the order they were scheduled and a job is scheduled only after
    // the jobs it consumes, so cross-lane waits cannot cycle.
    const double C_LEAF = 1.8e3, C_CHAIN = 4.7e3, C_INT = 7.0e3, C_MUL = 2.5e3, C_NODE = 1.0e3, C_XFER = 1.5e3;
    auto job_of_child = [&](int ch) { return ch - L; };          // internal child (flat id) -> job that produces it
    std::vector<int> lane_of(NJ, -1), lane_last(K, -1), chain_child(NJ, -1), parent_job(NJ, -1);
    std::vector<std::vector<int>> lanes(K);
    int njobs = 0;
    {
        std::vector<double> base(NJ, 0.0), tail(NJ, 0.0), finish(NJ, 0.0), lane_time(K, 0.0);
        std::vector<int> pending(NJ, 0);         // dirty producer jobs not yet scheduled
        std::vector<int> ready;
        for (int j = 0; j < NJ; j++) {
            if (!jdirty[j]) continue;
            njobs++;
            base[j] = C_NODE + (jmul[j] >= 0 ? C_MUL : 0.0);
            for (int ch : jch[j]) {
                base[j] += ch < L ? C_LEAF : C_INT;
                if (ch >= L && dirty[ch - L]) { pending[j]++; parent_job[job_of_child(ch)] = j; }
            }
            if (jmul[j] >= 0) { pending[j]++; parent_job[jmul[j]] = j; }
        }
        // remaining path to the root, parents before children: real nodes by decreasing height, a side job right after
        // its node
        std::vector<int> by_height;
        for (int n = 0; n < I; n++) if (dirty[n]) by_height.push_back(n);
        std::stable_sort(by_height.begin(), by_height.end(), [&](int x, int y) { return height[x] > height[y]; });
        for (int n : by_height) {
            tail[n] = base[n] + (parent_job[n] >= 0 ? tail[parent_job[n]] : 0.0);
            if (jmul[n] >= 0) tail[I + n] = base[I + n] + tail[n];
        }
        for (int j = 0; j < NJ; j++) if (jdirty[j] && pending[j] == 0) ready.push_back(j);
        auto producers = [&](int j, auto &&f) {   // dirty jobs whose output job j consumes
            for (int ch : jch[j]) if (ch >= L && dirty[ch - L]) f(job_of_child(ch));
            if (jmul[j] >= 0) f(jmul[j]);
        };
        for (int done = 0; done < njobs; done++) {
            int best = -1, best_r = 0; double best_start = 0.0;
            for (int j : ready) {
                for (int r = 0; r < K; r++) {
                    double est = lane_time[r];
                    producers(j, [&](int q) { est = std::max(est, finish[q] + (lane_of[q] == r ? 0.0 : C_XFER)); });
                    if (best < 0 || est < best_start || (est == best_start && tail[j] > tail[best])) { best = j; best_r = r; best_start = est; }
                }
            }
            const int j = best, r = best_r;
            double cost = base[j];
            for (int ch : jch[j])
                if (ch >= L && dirty[ch - L] && lane_last[r] == job_of_child(ch) && !(canonical_multi && jch[j].size() > 2)) { chain_child[j] = ch - L; cost -= C_INT - C_CHAIN; break; }
            finish[j] = best_start + cost;
            lane_time[r] = finish[j];
            lane_of[j] = r; lane_last[r] = j;
            lanes[r].push_back(j);
            ready.erase(std::find(ready.begin(), ready.end(), j));
            if (parent_job[j] >= 0 && --pending[parent_job[j]] == 0) ready.push_back(parent_job[j]);
        }
        if (getenv("HB2_DEBUG")) {
            double mx = 0; for (double t : lane_time) mx = std::max(mx, t);
            double deepest = 0; for (int j = 0; j < NJ; j++) if (jdirty[j]) deepest = std::max(deepest, tail[j]);
            fprintf(stderr, "[hb2] walk schedule: %d jobs, modelled makespan %.0f cycles, deepest root path %.0f, lanes (jobs/cycles):", njobs, mx, deepest);
            for (int r = 0; r < K; r++) fprintf(stderr, " %zu/%.0f", lanes[r].size(), lane_time[r]);
            fprintf(stderr, "\n");
        }
    }
    // flatten every lane into steps: chain child first, then leaves, then the other internal children, then the side product
    int ns = 0;
    for (int r = 0; r < K; r++) {
        lane_start[r] = ns;
        for (int j : lanes[r]) {
            const int first = ns;
            auto push = [&](int enc) { steps[2 * ns] = enc; steps[2 * ns + 1] = j; ns++; };
            if (canonical_multi && jch[j].size() > 2) {          // tree order, leaves and internal children interleaved
                for (int ch : jch[j]) push(ch < L ? ch : (ch | ((dirty[ch - L] && lane_of[ch - L] != lane_of[j]) ? hb2::WALK_WAIT : 0)));
            } else {
            if (chain_child[j] >= 0) push((chain_child[j] + L) | hb2::WALK_CHAIN);
            for (int ch : jch[j]) if (ch < L) push(ch);
            for (int ch : jch[j]) {
                if (ch < L || ch - L == chain_child[j]) continue;
                const int ci = ch - L;
                push(ch | ((dirty[ci] && lane_of[ci] != lane_of[j]) ? hb2::WALK_WAIT : 0));
            }
            }
            if (jmul[j] >= 0) push((L + jmul[j]) | hb2::WALK_MUL | (lane_of[jmul[j]] != lane_of[j] ? hb2::WALK_WAIT : 0));
            steps[2 * first + 1] |= hb2::STEP_FIRST;
            steps[2 * (ns - 1) + 1] |= hb2::STEP_LAST;
        }
    }
    lane_start[K] = ns;
    jdirty_out = jdirty;
    return ns;
}

// Walk plan for prune64_tc_walk_kernel: K lanes of jobs (nodes and side products, in schedule order), flattened into
// one step per child.  Layout of the int buffer: lane_start[K+1] (16 ints) | steps (int2 each) | generation bits [C][2I]
int run_walk(hb2_partition *p, int cat0, int ncls, const std::vector<std::vector<int>> &levels) {
    const int I = (int)p->I, L = (int)p->L;
    const int T = (int)(p->Sp / hb2::TC_TILE_P);
    const int CT = ncls * T;
    int K = std::max(1, std::min(p->walk_lane_cap, p->walk_max_resident / std::max(CT, 1)));
    int total = 0;
    for (auto &lv : levels) total += (int)lv.size();
    if (total == 0) return 0;
    K = std::min(K, total);
    const int nslots = (K == 1) ? std::min(CT, p->walk_max_resident) : CT;
    std::vector<char> dirty(I, 0);
    for (auto &lv : levels) for (int n : lv) dirty[n] = 1;
    int *buf = p->h_walk;                        // lane_start[K+1] | steps (int2, from int offset 16) | generation bits
    int ns = p->plan_steps;
    // the plan depends on the tree, the dirty set and K only: optimisers re-evaluate the same set again and again
    const bool reuse = p->plan_steps > 0 && p->plan_K == K && p->plan_dirty == dirty;
    if (!reuse) {
    ns = plan_walk(p->children, p->height, L, I, dirty, K, p->walk_split_nodes, buf, buf + 16, p->plan_jdirty);
    p->plan_steps = ns; p->plan_K = K; p->plan_dirty = dirty;
    }   // !reuse
    // generation bits: every node re-pruned by this pass flips its bit (per class); the kernel tags what it writes with the
    // new bit and awaits cross-lane children on it.  The table travels behind the steps.
    for (int c = cat0; c < cat0 + ncls; c++)
        for (int j = 0; j < 2 * I; j++)
            if (p->plan_jdirty[j]) p->walk_gen[(size_t)c * 2 * I + j] ^= 1;
    const int gen_off = 16 + 2 * ns;
    std::copy(p->walk_gen.begin(), p->walk_gen.end(), buf + gen_off);
    const int nints = gen_off + (int)p->walk_gen.size();
    {
        cudaError_t ce = cudaMemcpyAsync(p->d_walk, p->h_walk, nints * sizeof(int), cudaMemcpyHostToDevice, p->stream);
        if (ce != cudaSuccess) {
            int cur = -1; cudaGetDevice(&cur);
            cudaPointerAttributes pa{}, pb{};
            cudaError_t e1 = cudaPointerGetAttributes(&pa, p->d_walk), e2 = cudaPointerGetAttributes(&pb, p->h_walk);
            return fail("walk plan upload failed: %s; nints=%d ns=%d K=%d total=%d cur_dev=%d part_dev=%d d_walk=%p(type %d dev %d err %d) h_walk=%p(type %d dev %d err %d) stream=%p query=%d",
                        cudaGetErrorString(ce), nints, ns, K, total, cur, p->device, (void *)p->d_walk, (int)pa.type, pa.device, (int)e1,
                        (void *)p->h_walk, (int)pb.type, pb.device, (int)e2, (void *)p->stream, (int)cudaStreamQuery(p->stream));
        }
    }
    hb2::PruneArgs a = prune_args(p, cat0);
    hb2::WalkArgs w;
    hb2::PruneTcArgs &t = w.a;
    t.PB = p->d_PB; t.PTf = p->d_PTf; t.cond = p->d_condf; t.scal = a.scal; t.leaf = a.leaf; t.ambig = a.ambig; t.pi = a.pi;
    t.rootL = a.rootL; t.rootE = a.rootE; t.tree = a.tree; t.err = p->d_err;
    t.L = a.L; t.I = a.I; t.B = a.B; t.D = a.D; t.Sp = a.Sp; t.cat0 = a.cat0;
    t.forced = a.forced; t.forced_node = a.forced_node;
    w.lane_start = p->d_walk; w.steps = reinterpret_cast<const int2 *>(p->d_walk + 16);
    w.gen = p->d_walk + gen_off; w.NI = 2 * I; w.K = K; w.T = T; w.ncls = ncls; w.nslots = nslots;
    if (getenv("HB2_DEBUG")) fprintf(stderr, "[hb2] walk: jobs=%d steps=%d K=%d T=%d ncls=%d nslots=%d grid=%d resident=%d\n", total, ns, K, T, ncls, nslots, nslots * K, p->walk_max_resident);
    w.trace = nullptr; w.trace_cta = 0; w.trace_cta_times = nullptr;
    { static const float thr = getenv("HB2_ANCHOR_THR") ? (float)atof(getenv("HB2_ANCHOR_THR")) : hb2::TC_ANCHOR_THR; w.anchor_thr = thr; }
    const char *trace_path = getenv("HB2_WALK_TRACE");
    long long *d_trace = nullptr;
    if (trace_path) {          // bring-up aid: per-step clock stamps of one CTA -> text file
        const char *tc = getenv("HB2_WALK_TRACE_CTA");
        w.trace_cta = tc ? atoi(tc) : 0;
        const size_t tr_n = (size_t)(ns + 1) * 12 + (size_t)nslots * K * 4;
        CU(cudaMalloc(&d_trace, tr_n * sizeof(long long)));
        CU(cudaMemsetAsync(d_trace, 0, tr_n * sizeof(long long), p->stream));
        w.trace = d_trace;
        w.trace_cta_times = d_trace + (size_t)(ns + 1) * 12;
    }
    if (p->walk_v2) hb2::prune64_tc_walk2_kernel<<<nslots * K, 256, hb2::WALK2_SMEM_BYTES, p->stream>>>(w);
    else hb2::prune64_tc_walk_kernel<<<nslots * K, 128, hb2::WALK_SMEM_BYTES, p->stream>>>(w);
    if (d_trace) {
        std::vector<long long> ht((size_t)(ns + 1) * 12 + (size_t)nslots * K * 4);
        CU(cudaStreamSynchronize(p->stream));
        CU(cudaMemcpy(ht.data(), d_trace, ht.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(d_trace);
        if (FILE *f = fopen(trace_path, "w")) {
            const int r = w.trace_cta % K;
            fprintf(f, "# cta %d lane %d steps %d..%d ; columns: child_enc parent_flags t_begin bar1 staged/ready tableready|split bar2 bfull mma_done end (cycles rel. to first)\n", w.trace_cta, r, buf[r], buf[r + 1]);
            const long long t00 = ht[1];
            for (int i = 0; i < buf[r + 1] - buf[r]; i++) {
                const long long *q = ht.data() + (size_t)i * 12;
                fprintf(f, "%d 0x%x 0x%x", i, (unsigned)(q[0] >> 32), (unsigned)(q[0] & 0xffffffff));
                for (int c = 1; c < 12; c++) fprintf(f, " %lld", q[c] ? q[c] - t00 : -1LL);
                fprintf(f, "\n");
            }
            {   // per-CTA residence: "# C <block> <smid> <cycles> <start ns rel. to the earliest CTA>"
                const long long *ct = ht.data() + (size_t)(ns + 1) * 12;
                long long g0 = ct[3];
                for (int b = 0; b < nslots * K; b++) g0 = std::min(g0, ct[(size_t)b * 4 + 3]);
                for (int b = 0; b < nslots * K; b++)
                    fprintf(f, "# C %d %lld %lld %lld\n", b, ct[(size_t)b * 4], ct[(size_t)b * 4 + 2] - ct[(size_t)b * 4 + 1], ct[(size_t)b * 4 + 3] - g0);
            }
            fclose(f);
        }
    }
    p->launches++;
    CU(cudaGetLastError());
    return 0;
}

// 4 states: one or two patterns per thread -- (passes over the resident capacity) x (measured cost of a pass: 0.63 vs 0.98 ms
// at 256 taxa, profiles/r2x_small*.json)
static bool small_two_patterns(const hb2_partition *p, int ncls) {
    if (p->Dp != 4 || p->small_ilp == 1) return false;
    if (p->small_ilp == 2) return true;
    const long ct1 = (long)(p->Sp / 128) * ncls, ct2 = (long)((p->Sp + 255) / 256) * ncls;
    return 1.55 * std::ceil((double)ct2 / p->small_resident[1]) < std::ceil((double)ct1 / p->small_resident[0]);
}

// Small state spaces: one launch, one thread per pattern walking the dirty nodes in post-order.
int run_small_walk(hb2_partition *p, int cat0, int ncls, const std::vector<std::vector<int>> &levels) {
    std::vector<int> jobs;
    for (auto &lv : levels) jobs.insert(jobs.end(), lv.begin(), lv.end());
    if (jobs.empty()) return 0;
    std::sort(jobs.begin(), jobs.end());                 // internal indices ascending == post-order
    std::copy(jobs.begin(), jobs.end(), p->h_jobs);
    CU(cudaMemcpyAsync(p->d_jobs, p->h_jobs, jobs.size() * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    hb2::PruneArgs a = prune_args(p, cat0);
    dim3 grid((unsigned)(p->Sp / 128), (unsigned)ncls);
    const int n = (int)jobs.size();
    if (p->Dp >= 16 && p->small_dmma) {          // 16 / 24 / 32 padded states: FP64 tensor pipe, 8 patterns per warp
        // every CTA walks the whole tree: the launch is a few long waves.  Shape with the smaller last-wave loss: 8 warps
        // (4 CTAs per SM; 3 at 32 states) or, up to 24 states, 4 warps (9 per SM).
        const double warps = (double)(p->Sp / 8) * ncls;
        const int cap8 = (p->Dp == 32 ? 3 : 4) * 8 * p->sm_count, cap4 = 9 * 4 * p->sm_count;
        auto cost = [&](int cap) { const double wv = warps / cap; return std::ceil(wv - 1e-9) / wv; };      // 1 = no tail loss
        int nw = 8;
        if (p->Dp <= 24 && cost(cap4) < cost(cap8) - 0.05) nw = 4;
        if (const char *ov = getenv("HB2_SMALL_WARPS")) nw = (atoi(ov) == 4 && p->Dp <= 24) ? 4 : 8;
        dim3 g((unsigned)(p->Sp / (8 * nw)), (unsigned)ncls);
        if (nw == 8) {
            switch (p->Dp) {
                case 16: hb2::prune_small_dmma_kernel<16, 8><<<g, 256, 0, p->stream>>>(a, p->d_jobs, n); break;
                case 24: hb2::prune_small_dmma_kernel<24, 8><<<g, 256, 0, p->stream>>>(a, p->d_jobs, n); break;
                case 32: hb2::prune_small_dmma_kernel<32, 8><<<g, 256, 0, p->stream>>>(a, p->d_jobs, n); break;
                default: return fail("unsupported padded state count %d", p->Dp);
            }
        } else {
            if (p->Dp == 16) hb2::prune_small_dmma_kernel<16, 4><<<g, 128, 0, p->stream>>>(a, p->d_jobs, n);
            else hb2::prune_small_dmma_kernel<24, 4><<<g, 128, 0, p->stream>>>(a, p->d_jobs, n);
        }
        p->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    if (small_two_patterns(p, ncls)) {
        dim3 g2((unsigned)((p->Sp + 255) / 256), (unsigned)ncls);
        hb2::prune_small_walk_ilp_kernel<4, 2><<<g2, 128, 0, p->stream>>>(a, p->d_jobs, n);
        p->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    switch (p->Dp) {
        case 4: hb2::prune_small_walk_kernel<4><<<grid, 128, 0, p->stream>>>(a, p->d_jobs, n); break;
        case 8: hb2::prune_small_walk_kernel<8><<<grid, 128, 0, p->stream>>>(a, p->d_jobs, n); break;
        case 16: hb2::prune_small_walk_kernel<16><<<grid, 128, 0, p->stream>>>(a, p->d_jobs, n); break;
        case 24: hb2::prune_small_walk_kernel<24><<<grid, 128, 0, p->stream>>>(a, p->d_jobs, n); break;
        case 32: hb2::prune_small_walk_kernel<32><<<grid, 128, 0, p->stream>>>(a, p->d_jobs, n); break;
        default: return fail("unsupported padded state count %d", p->Dp);
    }
    p->launches++;
    CU(cudaGetLastError());
    return 0;
}

// 33..64 states in fp64: one launch, CTA (tile, class) walks the dirty nodes in post-order (prune64_walk_kernel).
int run_fp64_walk(hb2_partition *p, int cat0, int ncls, const std::vector<std::vector<int>> &levels) {
    std::vector<int> jobs;
    for (auto &lv : levels) jobs.insert(jobs.end(), lv.begin(), lv.end());
    if (jobs.empty()) return 0;
    std::sort(jobs.begin(), jobs.end());                 // internal indices ascending == post-order
    std::copy(jobs.begin(), jobs.end(), p->h_jobs);      // (every evaluation ends with a stream synchronisation: the buffer is free)
    CU(cudaMemcpyAsync(p->d_jobs, p->h_jobs, jobs.size() * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    hb2::PruneArgs a = prune_args(p, cat0);
    dim3 grid((unsigned)(p->Sp / hb2::TILE_P), (unsigned)ncls);
    hb2::prune64_walk_kernel<<<grid, 256, 2 * 64 * hb2::LD64 * sizeof(double), p->stream>>>(a, p->d_jobs, (int)jobs.size());
    p->launches++;
    CU(cudaGetLastError());
    return 0;
}

// 33..64 states in fp64, default: one launch, K in-order lanes per (class, 64-pattern tile) (prune64_lanes_kernel).  Same
// planner and plan cache as the tensor walk; the step list travels without a generation table (hand-over by pass id).
int run_fp64_lanes(hb2_partition *p, int cat0, int ncls, const std::vector<std::vector<int>> &levels) {
    const int I = (int)p->I, L = (int)p->L;
    int total = 0;
    for (auto &lv : levels) total += (int)lv.size();
    if (total == 0) return 0;
    // CTA shape: 4 warps x 8 patterns, 4 CTAs per SM.  The 8-warp shape (64 patterns, 3 per SM) gives the north-star shape a
    // third lane and c5 a single round, and is slower on both (1.28 vs 0.98 ms, 7.6 vs 5.6 ms: profiles/r2s_*): it stays
    // behind HB2_LANES_WARPS=8 for A/B runs.
    const int nw = p->lanes_force_nw == 8 ? 8 : 4;
    const int T = (int)(p->Sp / (8 * nw)), CT = ncls * T, resident = p->lanes_resident[nw == 8];
    const int K = std::max(1, std::min(std::min(p->walk_lane_cap, total), resident / std::max(CT, 1)));
    const int nslots = (K == 1) ? std::min(CT, resident) : CT;
    std::vector<char> dirty(I, 0);
    for (auto &lv : levels) for (int n : lv) dirty[n] = 1;
    int *buf = p->h_walk;
    int ns = p->plan_steps;
    const bool reuse = p->plan_steps > 0 && p->plan_K == K && p->plan_dirty == dirty;
    if (!reuse) {
        ns = plan_walk(p->children, p->height, L, I, dirty, K, p->walk_split_nodes, buf, buf + 16, p->plan_jdirty, true);
        p->plan_steps = ns; p->plan_K = K; p->plan_dirty = dirty;
        // the kernel's descriptors: (child, job, branch id of the lane's next contraction, -) behind the planner's pairs
        int *d4 = buf + ((16 + 2 * ns + 3) & ~3);
        std::vector<char> awaited(2 * I, 0);                 // jobs some other lane waits for: only those publish a flag
        for (int i = 0; i < ns; i++)
            if (buf[16 + 2 * i] & hb2::WALK_WAIT) awaited[(buf[16 + 2 * i] & hb2::WALK_ID_MASK) - L] = 1;
        for (int r = 0; r < K; r++) {
            int nextb = -1;
            for (int i = buf[r + 1] - 1; i >= buf[r]; i--) {
                const int enc = buf[16 + 2 * i], child = enc & hb2::WALK_ID_MASK;
                const int fl = buf[16 + 2 * i + 1];
                d4[4 * i] = enc; d4[4 * i + 1] = fl; d4[4 * i + 2] = nextb;
                d4[4 * i + 3] = ((fl & hb2::STEP_LAST) && awaited[fl & hb2::WALK_ID_MASK]) ? 1 : 0;
                if (child >= L && !(enc & hb2::WALK_MUL)) nextb = child;
            }
        }
    }
    const int o4 = (16 + 2 * ns + 3) & ~3;
    CU(cudaMemcpyAsync(p->d_walk, p->h_walk, (size_t)(o4 + 4 * ns) * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    hb2::LaneArgs w;
    w.a = prune_args(p, cat0);
    w.cond_side = p->d_cond_side;
    w.scal_side = p->d_scal + (size_t)p->C * I * p->Sp;
    w.lane_start = p->d_walk; w.steps = reinterpret_cast<const int4 *>(p->d_walk + o4);
    w.flags = p->d_lane_flags; w.err = p->d_err;
    w.K = K; w.T = T; w.ncls = ncls; w.nslots = nslots; w.pass = ++p->lane_pass;
    if (getenv("HB2_DEBUG")) fprintf(stderr, "[hb2] fp64 lanes: jobs=%d steps=%d K=%d T=%d ncls=%d nslots=%d grid=%d warps=%d resident=%d\n", total, ns, K, T, ncls, nslots, nslots * K, nw, resident);
    w.trace = nullptr; w.trace_cta = 0;
    const char *trace_path = getenv("HB2_WALK_TRACE");
    long long *d_trace = nullptr;
    if (trace_path) {          // bring-up aid: per-step clock stamps of one CTA -> text file
        const char *tc = getenv("HB2_WALK_TRACE_CTA");
        w.trace_cta = tc ? atoi(tc) : 0;
        CU(cudaMalloc(&d_trace, (size_t)(ns + 1) * 12 * sizeof(long long)));
        CU(cudaMemsetAsync(d_trace, 0, (size_t)(ns + 1) * 12 * sizeof(long long), p->stream));
        w.trace = d_trace;
    }
    if (nw == 4) hb2::prune64_lanes_kernel<4><<<nslots * K, 128, hb2::lanes_smem_bytes(4), p->stream>>>(w);
    else hb2::prune64_lanes_kernel<8><<<nslots * K, 256, hb2::lanes_smem_bytes(8), p->stream>>>(w);
    if (d_trace) {
        std::vector<long long> ht((size_t)(ns + 1) * 12);
        CU(cudaStreamSynchronize(p->stream));
        CU(cudaMemcpy(ht.data(), d_trace, ht.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(d_trace);
        if (FILE *f = fopen(trace_path, "w")) {
            const int r = w.trace_cta % K;
            fprintf(f, "# fp64 lanes kernel, cta %d lane %d steps %d..%d ; columns: child_enc job_flags begin staged landed barrier product barrier2 body_end stored barrier3 published (cycles rel. to first)\n", w.trace_cta, r, buf[r], buf[r + 1]);
            const long long t00 = ht[1];
            for (int i = 0; i < buf[r + 1] - buf[r]; i++) {
                const long long *q = ht.data() + (size_t)i * 12;
                fprintf(f, "%d 0x%x 0x%x", i, (unsigned)(q[0] >> 32), (unsigned)(q[0] & 0xffffffff));
                for (int c = 1; c < 12; c++) fprintf(f, " %lld", q[c] ? q[c] - t00 : -1LL);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    p->launches++;
    CU(cudaGetLastError());
    return 0;
}

int run_pruning(hb2_partition *p, int cat0, int ncls, const std::vector<std::vector<int>> &levels) {
    if (p->use_tc && p->use_walk) return run_walk(p, cat0, ncls, levels);
    // one launch per evaluation pays only when (tiles x classes) fills the machine about twice over (c5: 316 CTAs); below
    // that the per-level launches expose more parallelism (north-star shape: 128 CTAs walk 1.55 ms, 54 level launches take
    // 1.38 ms; a single class of it -- the patched host's per-class ComputeBlock -- would leave 116 SMs idle)
    if (!p->use_tc && p->Dp == 64 && p->fp64_mode == 2) return run_fp64_lanes(p, cat0, ncls, levels);
    if (!p->use_tc && p->Dp == 64 && p->fp64_walk && (int64_t)(p->Sp / hb2::TILE_P) * ncls >= 2 * p->sm_count)
        return run_fp64_walk(p, cat0, ncls, levels);
    if (p->Dp <= 32 && p->small_walk) return run_small_walk(p, cat0, ncls, levels);
    // upload all job lists in one copy
    int total = 0;
    for (auto &lv : levels) total += (int)lv.size();
    if (total == 0) return 0;
    int off = 0;
    for (auto &lv : levels) { std::copy(lv.begin(), lv.end(), p->h_jobs + off); off += (int)lv.size(); }
    CU(cudaMemcpyAsync(p->d_jobs, p->h_jobs, total * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    hb2::PruneArgs a = prune_args(p, cat0);
    off = 0;
    for (auto &lv : levels) {
        if (launch_prune(p, a, p->d_jobs + off, (int)lv.size(), ncls)) return 1;
        off += (int)lv.size();
    }
    return 0;
}

// (Re)creates the peer-exchange buffers with room for `payload` doubles per rank and maps every other rank's buffer through
// CUDA IPC; the 64-byte handles travel over the partition's NCCL communicator.  Collective: every rank calls it with the
// same payload.  On any failure (IPC not permitted in this container, no peer access) ALL ranks fall back to NCCL.
void peer_teardown(hb2_partition *p) {
    for (size_t r = 0; r < p->px_peer.size(); r++)
        if ((int)r != p->rank && p->px_peer[r]) cudaIpcCloseMemHandle(p->px_peer[r]);
    p->px_peer.clear();
    if (p->px_local) (cudaFree)(p->px_local);                 // IPC-exported: never pooled
    if (p->d_px_peer) cudaFree(p->d_px_peer);
    if (p->d_px_counter) cudaFree(p->d_px_counter);
    p->px_local = nullptr; p->d_px_peer = nullptr; p->d_px_counter = nullptr; p->px_ok = false;
}

int peer_setup(hb2_partition *p, int payload) {
    peer_teardown(p);
    { const char *env = getenv("HB2_PEER_XCHG"); p->px_enabled = !(env && env[0] == '0'); }
    if (!p->comm || p->n_ranks < 2 || p->n_ranks > 64) return 0;
    const int R = p->n_ranks;
    const size_t doubles = (size_t)2 * R * payload + (size_t)2 * R;          // data + flags (uint64 = double sized)
    int ok = p->px_enabled ? 1 : 0;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (ok && (cudaMalloc)((void **)&p->px_local, doubles * sizeof(double)) != cudaSuccess) { cudaGetLastError(); p->px_local = nullptr; ok = 0; }
    if (ok) {
        CU(cudaMemsetAsync(p->px_local, 0, doubles * sizeof(double), p->stream));
        if (cudaIpcGetMemHandle(&mine, p->px_local) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    // exchange {ok, handle} of every rank: one all-gather of 72 bytes per rank
    const size_t rec = 8 + sizeof(cudaIpcMemHandle_t);
    std::vector<unsigned char> sendrec(rec, 0), all(rec * R, 0);
    memcpy(sendrec.data(), &ok, sizeof(int));
    memcpy(sendrec.data() + 8, &mine, sizeof mine);
    unsigned char *d_send = nullptr, *d_all = nullptr;
    CU(cudaMalloc(&d_send, rec));
    CU(cudaMalloc(&d_all, rec * R));
    CU(cudaMemcpyAsync(d_send, sendrec.data(), rec, cudaMemcpyHostToDevice, p->stream));
    ncclResult_t nr = g_nccl.AllGather(d_send, d_all, rec, ncclChar, p->comm, p->stream);
    if (nr != ncclSuccess) { cudaFree(d_send); cudaFree(d_all); return fail("ncclAllGather (IPC handles): %s", g_nccl.GetErrorString(nr)); }
    CU(cudaMemcpyAsync(all.data(), d_all, rec * R, cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    cudaFree(d_send); cudaFree(d_all);
    int all_ok = 1;
    for (int r = 0; r < R; r++) { int o; memcpy(&o, all.data() + rec * r, sizeof(int)); all_ok = all_ok && o; }
    p->px_peer.assign(R, nullptr);
    if (all_ok) {
        for (int r = 0; r < R && all_ok; r++) {
            if (r == p->rank) { p->px_peer[r] = p->px_local; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, all.data() + rec * r + 8, sizeof h);
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); all_ok = 0; }
            p->px_peer[r] = static_cast<double *>(ptr);
        }
    }
    // the mapping must have worked on EVERY rank: agree with a max-reduce of "failed"
    int *d_flag2 = nullptr;
    CU(cudaMalloc(&d_flag2, sizeof(int)));
    int failed = all_ok ? 0 : 1;
    CU(cudaMemcpyAsync(d_flag2, &failed, sizeof(int), cudaMemcpyHostToDevice, p->stream));
    nr = g_nccl.AllReduce(d_flag2, d_flag2, 1, ncclInt32, ncclMax, p->comm, p->stream);
    if (nr != ncclSuccess) { cudaFree(d_flag2); return fail("ncclAllReduce (IPC agreement): %s", g_nccl.GetErrorString(nr)); }
    CU(cudaMemcpyAsync(&failed, d_flag2, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    cudaFree(d_flag2);
    if (failed) {
        const bool wanted = p->px_enabled;
        peer_teardown(p);
        if (wanted && getenv("HB2_DEBUG")) fprintf(stderr, "[hb2] rank %d: peer exchange unavailable (CUDA IPC), using NCCL collectives\n", p->rank);
        return 0;
    }
    CU(cudaMalloc(&p->d_px_peer, R * sizeof(double *)));
    CU(cudaMemcpy(p->d_px_peer, p->px_peer.data(), R * sizeof(double *), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&p->d_px_counter, sizeof(unsigned int)));
    CU(cudaMemset(p->d_px_counter, 0, sizeof(unsigned int)));
    p->px_payload = payload; p->px_gen = 0; p->px_ok = true;
    return 0;
}

hb2::PeerBuf peer_buf(hb2_partition *p) {
    hb2::PeerBuf b;
    b.peer = p->d_px_peer; b.R = p->n_ranks; b.rank = p->rank; b.payload = p->px_payload; b.gen = ++p->px_gen; b.err = p->d_err;
    return b;
}

int run_root(hb2_partition *p, int c0, int nc, bool use_weights, bool want_sites) {
    static const bool fuse_root = !(getenv("HB2_ROOT_FUSED") && getenv("HB2_ROOT_FUSED")[0] == '0');
    const bool plain = p->cg_G <= 1;           // single GPU or pattern shards: combine + sum in one kernel (its last block)
    if (!(plain && fuse_root)) CU(cudaMemsetAsync(p->d_flag, 0, sizeof(int), p->stream));
    hb2::CombineArgs c;
    c.counter = nullptr; c.out = nullptr;
    c.rootL = p->d_rootL; c.rootE = p->d_rootE; c.weights = use_weights ? p->d_weights : nullptr; c.freq = p->d_freq;
    c.partial = p->d_partial; c.flag = p->d_flag; c.siteL = want_sites ? p->d_siteL : nullptr;
    c.siteScale = want_sites ? p->d_siteScale : nullptr;
    c.Sp = (int)p->Sp; c.S = (int)p->S; c.c0 = c0; c.nc = nc;
    if (p->cg_G > 1 && p->px_ok) {
        // class groups over NVLink peer memory: ONE kernel computes this rank's class partials, stores them into every
        // rank's exchange buffer, and (its last block) merges all shards and classes into the complete lnL
        hb2::ClassXchgArgs x;
        x.rootL = p->d_rootL; x.rootE = p->d_rootE; x.weights = p->d_weights; x.Sp = (int)p->Sp; x.S = (int)p->S; x.c0 = p->own0; x.nc = p->ownN;
        x.xs = p->xchg_len; x.pb = peer_buf(p); x.xfreq = p->d_xfreq; x.nShards = p->n_ranks / p->cg_G; x.G = p->cg_G; x.myShard = p->rank / p->cg_G;
        x.lnL = p->d_lnL; x.siteL = c.siteL; x.siteScale = c.siteScale; x.counter = p->d_px_counter;
        const int nblk = std::max(1, std::min(16, (p->xchg_len + 255) / 256));
        hb2::class_exchange_kernel<<<nblk, 256, 0, p->stream>>>(x);
        p->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    if (p->cg_G > 1) {
        // class groups without peer mapping: partial over the owned classes into this rank's slot of a buffer that is zero
        // elsewhere; ONE sum all-reduce then acts as the gather (every rank merges every shard and holds the complete lnL)
        const int xs = p->xchg_len, N = p->n_ranks;
        hb2::class_partial_kernel<<<(xs + 255) / 256, 256, 0, p->stream>>>(p->d_rootL, p->d_rootE, p->d_weights, (int)p->Sp, (int)p->S,
                                                                          p->own0, p->ownN, p->d_xsend + (size_t)p->rank * 2 * xs, xs);
        ncclResult_t r = g_nccl.AllReduce(p->d_xsend, p->d_xrecv, (size_t)2 * xs * N, ncclDouble, ncclSum, p->comm, p->stream);
        if (r != ncclSuccess) return fail("ncclAllReduce (class partials): %s", g_nccl.GetErrorString(r));
        const int nblk = (N / p->cg_G * xs + 255) / 256;
        hb2::class_merge_kernel<<<nblk, 256, 0, p->stream>>>(p->d_xrecv, p->d_xfreq, xs, N / p->cg_G, p->cg_G, p->rank / p->cg_G, p->d_xpartial,
                                                             p->d_flag, c.siteL, c.siteScale, (int)p->S);
        hb2::final_sum_kernel<<<1, 256, 0, p->stream>>>(p->d_xpartial, nblk, p->d_flag, p->d_lnL);
        p->launches += 3;
        CU(cudaGetLastError());
        return 0;
    }
    if (fuse_root) {
        c.counter = p->d_root_counter; c.out = p->d_lnL;
        hb2::combine_kernel<<<p->n_partial_blocks, 256, 0, p->stream>>>(c);
        p->launches += 1;
    } else {
        hb2::combine_kernel<<<p->n_partial_blocks, 256, 0, p->stream>>>(c);
        hb2::final_sum_kernel<<<1, 256, 0, p->stream>>>(p->d_partial, p->n_partial_blocks, p->d_flag, p->d_lnL);
        p->launches += 2;
    }
    CU(cudaGetLastError());
    if (p->comm && p->px_ok) {                // pattern shards: R partial lnL over peer memory, summed in rank order
        hb2::peer_sum_kernel<<<1, 64, 0, p->stream>>>(peer_buf(p), p->d_lnL);
        p->launches++;
        CU(cudaGetLastError());
    } else if (p->comm) {
        ncclResult_t r = g_nccl.AllReduce(p->d_lnL, p->d_lnL, 1, ncclDouble, ncclSum, p->comm, p->stream);
        if (r != ncclSuccess) return fail("ncclAllReduce: %s", g_nccl.GetErrorString(r));
    }
    return 0;
}

int check_ready(hb2_partition *p, int c0, int nc) {
    for (int c = c0; c < c0 + nc; c++)
        for (int64_t b = 0; b < p->B; b++)
            if (!p->have_matrix[c * p->B + b])
                return fail("no matrix was ever set for rate class %d, node %lld", c, (long long)b);
    return 0;
}

int evaluate_impl(hb2_partition *p, int c0, int nc, const double *weights, int64_t nUpdate, const int64_t *updateNodes,
                  const double *rootFreqs, double *lnL, double *siteL, int64_t *siteScale, int64_t forcedNode = -1,
                  const int64_t *forcedStates = nullptr) {
    if (!p) return fail("null partition");
    if (!rootFreqs || !lnL) return fail("rootFreqs and lnL must not be null");
    CU(cudaSetDevice(p->device));
    p->forced_node = -1;
    if (forcedNode >= 0) {
        if (forcedNode >= p->L + p->I) return fail("forced node %lld out of range", (long long)forcedNode);
        if (!forcedStates) return fail("forced states are null");
        if (!p->d_forced) {
            CU(cudaMalloc(&p->d_forced, p->Sp * sizeof(int)));
            CU(cudaMallocHost(&p->h_forced, p->Sp * sizeof(int)));
        }
        CU(cudaStreamSynchronize(p->stream));         // h_forced may still be in flight from the previous forced evaluation
        for (int64_t s = 0; s < p->Sp; s++) {
            const int64_t f = s < p->S ? forcedStates[s] : 0;
            if (f < 0 || f >= p->D) return fail("forced state %lld of pattern %lld out of range", (long long)f, (long long)s);
            p->h_forced[s] = (int)f;
        }
        CU(cudaMemcpyAsync(p->d_forced, p->h_forced, p->Sp * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        p->forced_node = (int)forcedNode;
    }
    if (check_ready(p, c0, nc)) return 1;
    if (p->cg_G > 1) {
        if (!weights || c0 != 0 || nc != (int)p->C) return fail("with class groups only hb2_evaluate_classes is available");
        c0 = p->own0;                          // prune the owned classes; the weights of all C classes are still uploaded
    }
    p->bc_node = -1; p->bc_dirty_node = -1;     // conditionals are about to change: the branch cache is stale
    const int nw = nc;                         // number of class weights the caller passed
    if (p->cg_G > 1) nc = p->ownN;
    if (flush_matrices(p)) return 1;
    if (p->walk_reset) {                      // an earlier pass was aborted half-way: its tags are inconsistent
        CU(cudaMemsetAsync(p->d_condf, 0, (size_t)p->C * 2 * p->I * p->Sp * 64 * sizeof(float), p->stream));
        CU(cudaMemsetAsync(p->d_scal, 0, (size_t)p->C * 2 * p->I * p->Sp * sizeof(int), p->stream));
        CU(cudaMemsetAsync(p->d_err, 0, sizeof(int), p->stream));
        std::fill(p->walk_gen.begin(), p->walk_gen.end(), 0);
        std::fill(p->evaluated_cat.begin(), p->evaluated_cat.end(), 0);
        p->walk_reset = false;
    }
    // small inputs: pi (padded) and class weights
    double *hs = p->h_small;
    for (int k = 0; k < p->Dp; k++) hs[k] = k < p->D ? rootFreqs[k] : 0.0;
    if (weights) for (int c = 0; c < nw; c++) hs[p->Dp + c] = weights[c];
    CU(cudaMemcpyAsync(p->d_pi, hs, p->Dp * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    if (weights) CU(cudaMemcpyAsync(p->d_weights, hs + p->Dp, nw * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    // classes never pruned before need the whole tree regardless of updateNodes (likefunc.cpp:10965-10967)
    bool all = (updateNodes == nullptr || nUpdate < 0);
    for (int c = c0; c < c0 + nc; c++) if (!p->evaluated_cat[c]) all = true;
    std::vector<std::vector<int>> levels;
    if (plan_levels(p, all ? -1 : nUpdate, all ? nullptr : updateNodes, levels)) return 1;
    // the root must always be recomputed when anything changed; if nothing is dirty (only pi/weights changed)
    // the root tile products are still valid but rootL depends on pi -> recompute the root node
    bool any = false;
    for (auto &lv : levels) any = any || !lv.empty();
    if (!any) levels[p->height[p->I - 1]].push_back((int)p->I - 1);
    if (run_pruning(p, c0, nc, levels)) return 1;
    const bool want_sites = siteL != nullptr || siteScale != nullptr;
    if (run_root(p, c0, nc, weights != nullptr, want_sites)) return 1;
    CU(cudaMemcpyAsync(hs + p->Dp + p->C, p->d_lnL, sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaMemcpyAsync(hs + p->Dp + p->C + 1, p->d_err, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    if (siteL) CU(cudaMemcpyAsync(siteL, p->d_siteL, p->S * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (siteScale) CU(cudaMemcpyAsync(siteScale, p->d_siteScale, p->S * sizeof(long long), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    if (const int code = *reinterpret_cast<int *>(hs + p->Dp + p->C + 1)) {
        p->walk_reset = p->use_tc && p->use_walk;          // tags are inconsistent: the next pass starts from scratch
        cudaMemsetAsync(p->d_err, 0, sizeof(int), p->stream);   // on every path, or each later evaluation would fail too
        std::fill(p->evaluated_cat.begin(), p->evaluated_cat.end(), 0);
        return fail("a device-side wait timed out (code %d: 1 = operand staging / tensor pipe, 2 = another lane's conditionals, 3 = a peer "
                    "rank's partials, 4 = a coefficient matrix of the shared-powers exponential)", code);
    }
    *lnL = hs[p->Dp + p->C];
    for (int c = c0; c < c0 + nc; c++) p->evaluated_cat[c] = 1;
    return 0;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int hb2_abi_version(void) { return HB2_ABI_VERSION; }
const char *hb2_last_error(void) { return g_err.c_str(); }

int hb2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int hb2_create(hb2_partition **out, int64_t S, int64_t D, int64_t L, int64_t I, int64_t C, const int64_t *flatParents,
               const int64_t *leafState, const double *ambig, int64_t nAmb, const int64_t *patternFreq, int device,
               int flags) {
    if (!out) return fail("out is null");
    *out = nullptr;
    if (S <= 0 || D < 2 || L < 2 || I < 1 || C < 1) return fail("bad sizes S=%lld D=%lld L=%lld I=%lld C=%lld", (long long)S, (long long)D, (long long)L, (long long)I, (long long)C);
    if (!flatParents || !leafState || !patternFreq) return fail("null input array");
    if (nAmb > 0 && !ambig) return fail("nAmb > 0 but ambig is null");
    const int Dp = pad_states(D);
    if (Dp < 0) return fail("state count %lld not supported (max 64)", (long long)D);
    int ndev = hb2_device_count();
    if (ndev <= 0) return fail("no CUDA device visible: the B200 engine has no CPU fallback");
    if (device < 0 || device >= ndev) return fail("device %d out of range (%d visible)", device, ndev);
    CU(cudaSetDevice(device));

    hb2_partition *p = new hb2_partition();
    cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    p->device = device; p->flags = flags; p->S = S; p->D = D; p->L = L; p->I = I; p->C = C; p->B = L + I - 1; p->nAmb = nAmb;
    p->Dp = Dp;
    p->use_tc = (Dp == 64) && !(flags & HB2_FLAG_FORCE_FP64);
    const int64_t tile = (Dp == 64 && !p->use_tc) ? hb2::TILE_P : 128;
    p->Sp = (S + tile - 1) / tile * tile;
    p->parents.assign(flatParents, flatParents + L + I);
    p->children.assign(I, {});
    int roots = 0;
    for (int64_t n = 0; n < L + I; n++) {
        int64_t par = flatParents[n];
        if (par == -1) { roots++; if (n != L + I - 1) { delete p; return fail("root must be the last node (found -1 parent at %lld)", (long long)n); } continue; }
        if (par < 0 || par >= I) { delete p; return fail("flatParents[%lld]=%lld out of range", (long long)n, (long long)par); }
        if (n >= L && par <= n - L) { delete p; return fail("internal nodes must be in post-order (node %lld has parent %lld)", (long long)n, (long long)par); }
        p->children[par].push_back((int)n);
    }
    if (roots != 1) { delete p; return fail("tree must have exactly one root"); }
    p->height.assign(I, 0);
    for (int64_t i = 0; i < I; i++) {
        if (p->children[i].empty()) { delete p; return fail("internal node %lld has no children", (long long)i); }
        int h = 0;
        for (int ch : p->children[i]) if (ch >= L) h = std::max(h, p->height[ch - L] + 1);
        p->height[i] = h;
        p->max_height = std::max(p->max_height, h);
    }
    for (int64_t k = 0; k < L * S; k++) {
        int64_t c = leafState[k];
        if (c >= D || c < -nAmb) { delete p; return fail("leafState[%lld]=%lld out of range", (long long)k, (long long)c); }
    }
#define CUP(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fail("%s failed: %s", #x, cudaGetErrorString(e_)); hb2_destroy(p); return 1; } } while (0)
    CUP(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    for (auto &e : p->ev) CUP(cudaEventCreate(&e));
    CUP(cudaEventCreateWithFlags(&p->ev_staging, cudaEventDisableTiming));
    CUP(cudaEventCreateWithFlags(&p->ev_mix, cudaEventDisableTiming));
    const size_t Sp = p->Sp, dpdp = (size_t)Dp * Dp, dd = (size_t)D * D;
    CUP(cudaMalloc(&p->d_leaf, L * Sp * sizeof(int)));
    CUP(cudaMalloc(&p->d_ambig, std::max<int64_t>(nAmb, 1) * Dp * sizeof(double)));
    CUP(cudaMalloc(&p->d_freq, Sp * sizeof(double)));
    if (p->use_tc) {
        // 2I node slots per class: I nodes + I side products of the walk kernel (the per-level kernel uses the first I)
        CUP(cudaMalloc(&p->d_condf, (size_t)C * 2 * I * Sp * 64 * sizeof(float)));
        CUP(cudaMalloc(&p->d_PB, (size_t)C * p->B * hb2::TC_PB_FLOATS * sizeof(float)));
        CUP(cudaMalloc(&p->d_PTf, (size_t)C * p->B * hb2::TC_PTF_FLOATS * sizeof(float)));
        CUP(cudaMemsetAsync(p->d_condf, 0, (size_t)C * 2 * I * Sp * 64 * sizeof(float), p->stream));
        CUP(cudaMemsetAsync(p->d_PB, 0, (size_t)C * p->B * hb2::TC_PB_FLOATS * sizeof(float), p->stream));
        CUP(cudaMemsetAsync(p->d_PTf, 0, (size_t)C * p->B * hb2::TC_PTF_FLOATS * sizeof(float), p->stream));
        CUP(cudaFuncSetAttribute(hb2::prune64_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hb2::TC_SMEM_BYTES));
        CUP(cudaFuncSetAttribute(hb2::prune64_tc_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hb2::WALK_SMEM_BYTES));
        CUP(cudaFuncSetAttribute(hb2::prune64_tc_walk2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hb2::WALK2_SMEM_BYTES));
        CUP(cudaFuncSetAttribute(hb2::prune64_tc_walk2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        { const char *env = getenv("HB2_WALK_V2"); p->walk_v2 = env && env[0] == '1'; }
        {
            const char *env = getenv("HB2_TC_WALK");
            p->use_walk = !(env && env[0] == '0');
            int per_sm = 0, sms = 0;
            CUP(cudaFuncSetAttribute(hb2::prune64_tc_walk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            const void *wk = p->walk_v2 ? (const void *)hb2::prune64_tc_walk2_kernel : (const void *)hb2::prune64_tc_walk_kernel;
            const int wk_threads = p->walk_v2 ? 256 : 128, wk_smem = p->walk_v2 ? hb2::WALK2_SMEM_BYTES : hb2::WALK_SMEM_BYTES;
            CUP(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wk, wk_threads, wk_smem));
            if (getenv("HB2_DEBUG")) {
                cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, wk);
                cudaDeviceProp dp; cudaGetDeviceProperties(&dp, device);
                fprintf(stderr, "[hb2] walk kernel occupancy: %d CTAs/SM (regs/thread %d, static smem %zu, dyn smem %d, regs/SM %d, smem/SM %zu, smem/block optin %zu, reserved/block %zu)\n",
                        per_sm, fa.numRegs, fa.sharedSizeBytes, wk_smem, dp.regsPerMultiprocessor, dp.sharedMemPerMultiprocessor, dp.sharedMemPerBlockOptin, dp.reservedSharedMemPerBlock);
            }
            CUP(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
            {   // The occupancy API reports 1 CTA/SM for these kernels on B200 although ncu (launch__occupancy_limit_* = 2) and
                // the measured residence of 256-CTA grids show two are co-resident: the lane count is planned from the
                // resources.  Correctness does not depend on it: blocks are dispatched in index order and a lane only waits
                // for lanes of its own (class, tile) slot -- adjacent block indices -- with time-bounded waits, so fewer
                // resident CTAs than planned only make the pass slower.
                cudaFuncAttributes fa;
                CUP(cudaFuncGetAttributes(&fa, wk));
                cudaDeviceProp dp; CUP(cudaGetDeviceProperties(&dp, device));
                const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * wk_threads;
                const size_t smem_per_cta = (size_t)wk_smem + fa.sharedSizeBytes + dp.reservedSharedMemPerBlock;
                const int by_res = std::min(dp.regsPerMultiprocessor / std::max(regs_per_cta, 1), (int)(dp.sharedMemPerMultiprocessor / smem_per_cta));
                per_sm = std::max(per_sm, std::min(by_res, 2));
            }
            if (const char *ov = getenv("HB2_WALK_CTAS_PER_SM")) per_sm = atoi(ov);      // bring-up override
            if (const char *ov = getenv("HB2_WALK_LANES")) p->walk_lane_cap = std::max(1, std::min(atoi(ov), 15));
            p->walk_max_resident = std::min(per_sm, 2) * sms;          // TMEM: 256 of 512 columns per CTA -> at most 2 per SM
            if (p->walk_max_resident < 1) p->use_walk = false;
            const size_t T = Sp / hb2::TC_TILE_P;
            p->walk_gen.assign((size_t)C * 2 * I, 0);   // matches the zero-filled conditional and exponent buffers
            const size_t walk_ints = 16 + 2 * (size_t)(L + 2 * I) + (size_t)C * 2 * I;
            { const char *ev = getenv("HB2_WALK_SPLIT_NODES"); p->walk_split_nodes = !(ev && ev[0] == '0'); }
            CUP(cudaMalloc(&p->d_walk, walk_ints * sizeof(int)));
            CUP(cudaMallocHost(&p->h_walk, walk_ints * sizeof(int)));
        }
    } else {
        CUP(cudaMalloc(&p->d_cond, (size_t)C * I * Sp * Dp * sizeof(double)));
        CUP(cudaMemsetAsync(p->d_cond, 0, (size_t)C * I * Sp * Dp * sizeof(double), p->stream));
        { const char *env = getenv("HB2_FP64_WALK"); p->fp64_mode = (env && env[0] >= '0' && env[0] <= '2') ? env[0] - '0' : 2; }
        if (Dp == 64 && p->fp64_mode == 2) {
            // lanes kernel: side products, hand-over flags, the step list, and how many CTAs can be co-resident
            const size_t T = Sp / 32;                                        // flags are sized for the smaller tile
            CUP(cudaMalloc(&p->d_cond_side, (size_t)C * I * Sp * 64 * sizeof(double)));
            CUP(cudaMemsetAsync(p->d_cond_side, 0, (size_t)C * I * Sp * 64 * sizeof(double), p->stream));
            CUP(cudaMalloc(&p->d_lane_flags, (size_t)C * 2 * I * T * sizeof(int)));
            CUP(cudaMemsetAsync(p->d_lane_flags, 0, (size_t)C * 2 * I * T * sizeof(int), p->stream));
            const size_t walk_ints = 16 + 6 * (size_t)(L + 2 * I) + 4;       // planner's pairs + the kernel's int4 descriptors
            CUP(cudaMalloc(&p->d_walk, walk_ints * sizeof(int)));
            CUP(cudaMallocHost(&p->h_walk, walk_ints * sizeof(int)));
            { const char *ev = getenv("HB2_WALK_SPLIT_NODES"); p->walk_split_nodes = !(ev && ev[0] == '0'); }
            if (const char *ov = getenv("HB2_WALK_LANES")) p->walk_lane_cap = std::max(1, std::min(atoi(ov), 15));
            if (const char *ov = getenv("HB2_LANES_WARPS")) p->lanes_force_nw = atoi(ov) == 8 ? 8 : atoi(ov) == 4 ? 4 : 0;
            int sms = 0;
            CUP(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
            cudaDeviceProp dp; CUP(cudaGetDeviceProperties(&dp, device));
            for (int v = 0; v < 2; v++) {
                const void *fn = v ? (const void *)hb2::prune64_lanes_kernel<8> : (const void *)hb2::prune64_lanes_kernel<4>;
                const int threads = v ? 256 : 128, smem = hb2::lanes_smem_bytes(v ? 8 : 4), cap = v ? 3 : 4;
                CUP(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                CUP(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                int per_sm = 0;
                CUP(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
                // same resource arithmetic as for the tensor walk kernel (the occupancy API under-reports there)
                cudaFuncAttributes fa;
                CUP(cudaFuncGetAttributes(&fa, fn));
                const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * threads;
                const size_t smem_per_cta = (size_t)smem + fa.sharedSizeBytes + dp.reservedSharedMemPerBlock;
                const int by_res = std::min(dp.regsPerMultiprocessor / std::max(regs_per_cta, 1), (int)(dp.sharedMemPerMultiprocessor / smem_per_cta));
                if (getenv("HB2_DEBUG")) fprintf(stderr, "[hb2] fp64 lanes kernel <%d warps>: occupancy API %d, by resources %d (regs %d)\n", v ? 8 : 4, per_sm, by_res, fa.numRegs);
                per_sm = std::max(per_sm, std::min(by_res, cap));
                if (const char *ov = getenv("HB2_WALK_CTAS_PER_SM")) per_sm = atoi(ov);
                p->lanes_resident[v] = std::max(per_sm, 1) * sms;
            }
        }
    }
    CUP(cudaMalloc(&p->d_err, sizeof(int)));
    CUP(cudaMemsetAsync(p->d_err, 0, sizeof(int), p->stream));
    CUP(cudaMalloc(&p->d_scal, (size_t)C * 2 * I * Sp * sizeof(int)));      // 2I: see d_condf
    CUP(cudaMalloc(&p->d_PT, (size_t)C * p->B * dpdp * sizeof(double)));
    CUP(cudaMalloc(&p->d_Qres, (size_t)C * p->B * dd * sizeof(double)));
    p->q_capacity = C * p->B;
    CUP(cudaMalloc(&p->d_Q, (size_t)p->q_capacity * dd * sizeof(double)));
    CUP(cudaMalloc(&p->d_dst, p->q_capacity * sizeof(int)));
    CUP(cudaMalloc(&p->d_pi, Dp * sizeof(double)));
    CUP(cudaMalloc(&p->d_rootL, (size_t)C * Sp * sizeof(double)));
    CUP(cudaMalloc(&p->d_rootE, (size_t)C * Sp * sizeof(int)));
    CUP(cudaMalloc(&p->d_weights, C * sizeof(double)));
    p->n_partial_blocks = (int)((S + 255) / 256);
    CUP(cudaMalloc(&p->d_partial, p->n_partial_blocks * sizeof(double)));
    CUP(cudaMalloc(&p->d_lnL, sizeof(double)));
    CUP(cudaMalloc(&p->d_flag, sizeof(int)));
    CUP(cudaMemsetAsync(p->d_flag, 0, sizeof(int), p->stream));
    CUP(cudaMalloc(&p->d_root_counter, sizeof(unsigned)));
    CUP(cudaMemsetAsync(p->d_root_counter, 0, sizeof(unsigned), p->stream));
    CUP(cudaMalloc(&p->d_siteL, Sp * sizeof(double)));
    CUP(cudaMalloc(&p->d_siteScale, Sp * sizeof(long long)));
    CUP(cudaMalloc(&p->d_jobs, I * sizeof(int)));
    CUP(cudaMalloc(&p->d_child_start, (I + 1) * sizeof(int)));
    CUP(cudaMalloc(&p->d_child_ids, (L + I) * sizeof(int)));
    CUP(cudaMallocHost(&p->h_Q, (size_t)p->q_capacity * dd * sizeof(double)));
    CUP(cudaMallocHost(&p->h_dst, p->q_capacity * sizeof(int)));
    CUP(cudaMallocHost(&p->h_small, (Dp + C + 8) * sizeof(double)));
    CUP(cudaMallocHost(&p->h_jobs, I * sizeof(int)));
    CUP(cudaMemsetAsync(p->d_scal, 0, (size_t)C * 2 * I * Sp * sizeof(int), p->stream));
    CUP(cudaMemsetAsync(p->d_PT, 0, (size_t)C * p->B * dpdp * sizeof(double), p->stream));
    CUP(cudaMemsetAsync(p->d_rootL, 0, (size_t)C * Sp * sizeof(double), p->stream));
    CUP(cudaMemsetAsync(p->d_rootE, 0, (size_t)C * Sp * sizeof(int), p->stream));
    {   // static uploads (pageable host vectors: synchronous copies are fine at setup time)
        std::vector<int> leaf((size_t)L * Sp, 0);          // padding patterns: state 0 everywhere, frequency 0
        for (int64_t l = 0; l < L; l++)
            for (int64_t s = 0; s < S; s++) leaf[l * Sp + s] = (int)leafState[l * S + s];
        std::vector<double> amb((size_t)std::max<int64_t>(nAmb, 1) * Dp, 0.0);
        for (int64_t a = 0; a < nAmb; a++)
            for (int64_t k = 0; k < D; k++) amb[a * Dp + k] = ambig[a * D + k];
        std::vector<double> fr(Sp, 0.0);
        for (int64_t s = 0; s < S; s++) fr[s] = (double)patternFreq[s];
        std::vector<int> cs(I + 1, 0), ci;
        for (int64_t i = 0; i < I; i++) { cs[i] = (int)ci.size(); for (int ch : p->children[i]) ci.push_back(ch); }
        cs[I] = (int)ci.size();
        CUP(cudaStreamSynchronize(p->stream));
        CUP(cudaMemcpy(p->d_leaf, leaf.data(), leaf.size() * sizeof(int), cudaMemcpyHostToDevice));
        CUP(cudaMemcpy(p->d_ambig, amb.data(), amb.size() * sizeof(double), cudaMemcpyHostToDevice));
        CUP(cudaMemcpy(p->d_freq, fr.data(), fr.size() * sizeof(double), cudaMemcpyHostToDevice));
        CUP(cudaMemcpy(p->d_child_start, cs.data(), cs.size() * sizeof(int), cudaMemcpyHostToDevice));
        CUP(cudaMemcpy(p->d_child_ids, ci.data(), ci.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (Dp == 32) CUP(cudaFuncSetAttribute(hb2::expm_small_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hb2::expm_small_smem_bytes(32)));
    if (Dp == 64) {
        { const char *env = getenv("HB2_EXPM_SHARED"); p->ex_enabled = !(env && env[0] == '0'); }
        p->ex.G = std::max((int)C + 8, HB2_MAX_TEMPLATES * (int)C); p->ex.cap = (int64_t)(C + 8) * p->B; p->ex.stride = 4096;
        p->ex.kind.assign(p->ex.G, 0);
        CUP(cudaMalloc(&p->ex.d_int, (size_t)(2 * p->ex.cap + p->ex.G) * sizeof(int)));
        CUP(cudaMallocHost(&p->ex.h_int, (size_t)(2 * p->ex.cap + p->ex.G) * sizeof(int)));
        CUP(cudaMalloc(&p->ex.d_flag, (size_t)p->ex.cap * sizeof(int)));
        CUP(cudaMalloc(&p->ex.d_weight, (size_t)p->ex.cap * sizeof(double)));
        CUP(cudaMalloc(&p->ex.d_colsum, (size_t)p->ex.cap * 16 * 64 * sizeof(double)));
        CUP(cudaMalloc(&p->ex.d_groups, (size_t)p->ex.G * sizeof(hb2::ExpmGroup)));
        CUP(cudaMemsetAsync(p->ex.d_groups, 0, (size_t)p->ex.G * sizeof(hb2::ExpmGroup), p->stream));
        CUP(cudaMalloc(&p->ex.d_pow, (size_t)p->ex.G * hb2::EXPM_POW_TERMS * 4096 * sizeof(double)));
        CUP(cudaMalloc(&p->ex.d_flags, (size_t)p->ex.G * hb2::EXPM_POW_TERMS * sizeof(unsigned long long)));
        CUP(cudaMemsetAsync(p->ex.d_flags, 0, (size_t)p->ex.G * hb2::EXPM_POW_TERMS * sizeof(unsigned long long), p->stream));
        CUP(cudaMalloc(&p->ex.d_refvec, (size_t)p->ex.G * p->ex.stride * sizeof(double)));
        CUP(cudaFuncSetAttribute(hb2::expm_powers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * 64 * hb2::LD64 * sizeof(double))));
        CUP(cudaFuncSetAttribute(hb2::expm64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(5 * 64 * hb2::LD64 * sizeof(double))));
        CUP(cudaFuncSetAttribute(hb2::expm64_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * 64 * hb2::LD64 * sizeof(double))));
        CUP(cudaFuncSetAttribute(hb2::expm64_dmma_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        { const char *env = getenv("HB2_EXPM_DFMA"); p->expm_dfma = env && env[0] == '1'; }
    }
    {
        CUP(cudaFuncSetAttribute(hb2::prune64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * 64 * hb2::LD64 * sizeof(double))));
        CUP(cudaFuncSetAttribute(hb2::prune64_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * 64 * hb2::LD64 * sizeof(double))));
    }
#undef CUP
    { const char *env = getenv("HB2_SMALL_WALK"); p->small_walk = !(env && env[0] == '0'); }
    { const char *env = getenv("HB2_SMALL_DMMA"); p->small_dmma = !(env && env[0] == '0'); }
    if (Dp == 4) {
        if (const char *env = getenv("HB2_SMALL_ILP")) p->small_ilp = atoi(env);
        int b1 = 0, b2 = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, hb2::prune_small_walk_kernel<4>, 128, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b2, hb2::prune_small_walk_ilp_kernel<4, 2>, 128, 0);
        p->small_resident[0] = std::max(b1, 1) * p->sm_count;
        p->small_resident[1] = std::max(b2, 1) * p->sm_count;
    }
    p->fp64_walk = p->fp64_mode >= 1;
    p->have_matrix.assign(C * p->B, 0);
    p->is_rate.assign(C * p->B, 0);
    p->pend_pos.assign(C * p->B, -1);
    p->pend_tmpl.assign(C * p->B, 0);
    p->res_tmpl.assign(C * p->B, 0);
    p->evaluated_cat.assign(C, 0);
    p->own0 = 0; p->ownN = (int)C;
    *out = p;
    return 0;
}

int hb2_set_matrices(hb2_partition *p, int64_t cat, int64_t n, const int64_t *nodeIds, const double *const *M, int kind) {
    if (!p) return fail("null partition");
    if (n < 0 || (n > 0 && (!nodeIds || !M))) return fail("bad matrix list");
    for (int64_t k = 0; k < n; k++) {
        if (!M[k]) return fail("matrix %lld is null", (long long)k);
        if (stage_matrix(p, cat, nodeIds[k], M[k], kind)) return 1;
    }
    return 0;
}

int hb2_set_matrices_packed(hb2_partition *p, int64_t cat, int64_t n, const int64_t *nodeIds, const double *M, int kind) {
    if (!p) return fail("null partition");
    if (n < 0 || (n > 0 && (!nodeIds || !M))) return fail("bad matrix list");
    const size_t dd = (size_t)p->D * p->D;
    for (int64_t k = 0; k < n; k++)
        if (stage_matrix(p, cat, nodeIds[k], M + k * dd, kind)) return 1;
    return 0;
}

int hb2_set_mixture_matrices(hb2_partition *p, int64_t cat, int64_t n, const int64_t *nodeIds, int64_t K, const double *M,
                             const double *w) {
    if (!p) return fail("null partition");
    if (n < 0 || K < 1 || (n > 0 && (!nodeIds || !M || !w))) return fail("bad mixture arguments");
    if (cat < 0) cat = 0;
    if (cat >= p->C) return fail("rate class %lld out of range", (long long)cat);
    for (int64_t i = 0; i < n; i++) {
        if (nodeIds[i] < 0 || nodeIds[i] >= p->B) return fail("node id %lld has no branch", (long long)nodeIds[i]);
        note_matrix_change(p, nodeIds[i]);
    }
    if (cat < p->own0 || cat >= p->own0 + p->ownN) {       // another class group's
        for (int64_t i = 0; i < n; i++) p->have_matrix[cat * p->B + nodeIds[i]] = 1;
        return 0;
    }
    if (n == 0) return 0;
    CU(cudaSetDevice(p->device));
    if (flush_matrices(p)) return 1;         // keep ordering with plain matrices staged earlier (same stream)
    // Two launches, no host synchronisation: all n*K components are exponentiated side by side into a scratch area
    // (one CTA each), then one CTA per node forms sum_k w_k Exp(Q_k) in component order (deterministic) and emits the
    // tensor-path operands.  Own pinned staging: [n*K matrices | n*K weights], then the n destination slots as ints.
    const size_t dd = (size_t)p->D * p->D, dpdp = (size_t)p->Dp * p->Dp;
    const int64_t nk = n * K;
    if (p->mix_busy) { CU(cudaEventSynchronize(p->ev_mix)); p->mix_busy = false; }
    if (nk > p->mix_capacity) {
        CU(cudaStreamSynchronize(p->stream));
        for (void *d : {(void *)p->d_mix_scratch, (void *)p->d_mix_Q, (void *)p->d_mix_dst}) if (d) cudaFree(d);
        if (p->h_mix) cudaFreeHost(p->h_mix);
        p->d_mix_scratch = p->d_mix_Q = nullptr; p->d_mix_dst = nullptr; p->h_mix = nullptr; p->mix_capacity = 0;
        CU(cudaMalloc(&p->d_mix_scratch, (size_t)nk * dpdp * sizeof(double)));
        CU(cudaMalloc(&p->d_mix_Q, (size_t)nk * (dd + 1) * sizeof(double)));
        CU(cudaMalloc(&p->d_mix_dst, (size_t)2 * nk * sizeof(int)));
        CU(cudaMallocHost(&p->h_mix, (size_t)nk * (dd + 1) * sizeof(double) + (size_t)2 * nk * sizeof(int)));
        p->mix_capacity = nk;
    }
    memcpy(p->h_mix, M, (size_t)nk * dd * sizeof(double));
    memcpy(p->h_mix + (size_t)nk * dd, w, (size_t)nk * sizeof(double));
    int *h_idx = reinterpret_cast<int *>(p->h_mix + (size_t)nk * (dd + 1));      // [nk] scratch slots 0..nk-1 | [n] P-cache slots
    for (int64_t k = 0; k < nk; k++) h_idx[k] = (int)k;
    for (int64_t i = 0; i < n; i++) h_idx[nk + i] = (int)(cat * p->B + nodeIds[i]);
    CU(cudaMemcpyAsync(p->d_mix_Q, p->h_mix, (size_t)nk * (dd + 1) * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CU(cudaMemcpyAsync(p->d_mix_dst, h_idx, (size_t)(nk + n) * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    CU(cudaEventRecord(p->ev_mix, p->stream));
    p->mix_busy = true;
    // component k of every node is (in BS-REL / BUSTED models) the same rate matrix up to the branch length: the components
    // form shared-powers groups C + k
    SharedPlan sp;
    if (K <= 8) {
        if (wait_staging(p)) return 1;
        std::vector<int> grp(nk);
        for (int64_t k = 0; k < nk; k++) grp[k] = (int)(p->C + k % K);
        if (plan_shared(p, h_idx, nk, 2, grp.data(), sp)) return 1;
        CU(cudaEventRecord(p->ev_staging, p->stream));
        p->staging_busy = true;
    }
    if (launch_expm(p, p->d_mix_Q, p->d_mix_dst, (int)nk, 0, nullptr, false, p->d_mix_scratch, &sp)) return 1;
    hb2::mix_reduce_kernel<<<(unsigned)n, 256, 0, p->stream>>>(p->d_mix_scratch, p->d_mix_Q + (size_t)nk * dd, p->d_mix_dst + nk, (int)K, p->Dp,
                                                              p->d_PT, p->use_tc ? p->d_PB : nullptr, p->d_PTf);
    p->launches++;
    CU(cudaGetLastError());
    for (int64_t i = 0; i < n; i++) { p->have_matrix[cat * p->B + nodeIds[i]] = 1; p->is_rate[cat * p->B + nodeIds[i]] = 0; }
    return 0;
}

static void free_template(hb2_partition::Tmpl &t) {
    for (void *d : {(void *)t.d_index, (void *)t.d_formula, (void *)t.d_vdst, (void *)t.d_colfreq, (void *)t.d_V, (void *)t.d_Vres}) if (d) cudaFree(d);
    for (void *h : {(void *)t.h_vdst, (void *)t.h_colfreq, (void *)t.h_V}) if (h) cudaFreeHost(h);
    t = hb2_partition::Tmpl();
}

int hb2_set_rate_template_id(hb2_partition *p, int64_t templateId, int64_t nnz, const int64_t *entryIndex, const int64_t *entryFormula,
                             int64_t nFormulas, const double *colFreq) {
    if (!p) return fail("null partition");
    if (templateId < 0 || templateId >= HB2_MAX_TEMPLATES) return fail("template id %lld out of range (0..%d)", (long long)templateId, HB2_MAX_TEMPLATES - 1);
    if (nnz < 1 || nFormulas < 1 || !entryIndex || !entryFormula) return fail("bad template arguments");
    CU(cudaSetDevice(p->device));
    if (flush_matrices(p)) return 1;
    CU(cudaStreamSynchronize(p->stream));
    p->staging_busy = false;
    std::vector<int> idx(nnz), frm(nnz);
    for (int64_t e = 0; e < nnz; e++) {
        if (entryIndex[e] < 0 || entryIndex[e] >= p->D * p->D) return fail("template entry %lld: index %lld out of range", (long long)e, (long long)entryIndex[e]);
        if (entryFormula[e] < 0 || entryFormula[e] >= nFormulas) return fail("template entry %lld: formula %lld out of range", (long long)e, (long long)entryFormula[e]);
        idx[e] = (int)entryIndex[e]; frm[e] = (int)entryFormula[e];
    }
    if ((int64_t)p->tmpls.size() <= templateId) p->tmpls.resize(templateId + 1);
    hb2_partition::Tmpl &t = p->tmpls[templateId];
    free_template(t);
    CU(cudaMalloc(&t.d_index, nnz * sizeof(int)));
    CU(cudaMalloc(&t.d_formula, nnz * sizeof(int)));
    CU(cudaMalloc(&t.d_colfreq, p->D * sizeof(double)));
    CU(cudaMallocHost(&t.h_colfreq, p->D * sizeof(double)));
    CU(cudaMalloc(&t.d_V, (size_t)p->q_capacity * nFormulas * sizeof(double)));
    CU(cudaMalloc(&t.d_vdst, p->q_capacity * sizeof(int)));
    // resident formula values of every slot (replay, shared-powers classification); cached directions of this template's
    // groups are void now
    CU(cudaMalloc(&t.d_Vres, (size_t)p->C * p->B * nFormulas * sizeof(double)));
    if (p->ex.d_groups) {
        for (int c = 0; c < (int)p->C; c++) {
            const int g = (int)templateId * (int)p->C + c;
            if (g < p->ex.G) {
                CU(cudaMemset(static_cast<hb2::ExpmGroup *>(p->ex.d_groups) + g, 0, sizeof(hb2::ExpmGroup)));
                p->ex.kind[g] = 0;
            }
        }
    }
    for (size_t k = 0; k < p->is_rate.size(); k++) if (p->is_rate[k] == 2 && p->res_tmpl[k] == templateId) p->is_rate[k] = 0;
    CU(cudaMallocHost(&t.h_V, (size_t)p->q_capacity * nFormulas * sizeof(double)));
    CU(cudaMallocHost(&t.h_vdst, p->q_capacity * sizeof(int)));
    CU(cudaMemcpy(t.d_index, idx.data(), nnz * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(t.d_formula, frm.data(), nnz * sizeof(int), cudaMemcpyHostToDevice));
    t.has_colfreq = colFreq != nullptr;
    if (colFreq) CU(cudaMemcpy(t.d_colfreq, colFreq, p->D * sizeof(double), cudaMemcpyHostToDevice));
    t.nnz = nnz; t.nF = nFormulas; t.n_pending = 0;
    return 0;
}

int hb2_set_rate_template(hb2_partition *p, int64_t nnz, const int64_t *entryIndex, const int64_t *entryFormula,
                          int64_t nFormulas, const double *colFreq) {
    return hb2_set_rate_template_id(p, 0, nnz, entryIndex, entryFormula, nFormulas, colFreq);
}

int hb2_set_template_frequencies(hb2_partition *p, int64_t templateId, const double *colFreq) {
    if (!p || !colFreq) return fail("null argument");
    if (templateId < 0 || templateId >= (int64_t)p->tmpls.size() || p->tmpls[templateId].nnz == 0) return fail("template %lld has not been defined", (long long)templateId);
    hb2_partition::Tmpl &t = p->tmpls[templateId];
    if (!t.has_colfreq) return fail("template %lld was defined without column frequencies", (long long)templateId);
    CU(cudaSetDevice(p->device));
    if (flush_matrices(p)) return 1;          // matrices handed over so far were assembled with the old frequencies
    if (wait_staging(p)) return 1;
    memcpy(t.h_colfreq, colFreq, p->D * sizeof(double));
    CU(cudaMemcpyAsync(t.d_colfreq, t.h_colfreq, p->D * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CU(cudaEventRecord(p->ev_staging, p->stream));
    p->staging_busy = true;
    return 0;
}

int hb2_set_matrices_compiled_id(hb2_partition *p, int64_t templateId, int64_t cat, int64_t n, const int64_t *nodeIds, const double *formulaValues) {
    if (!p) return fail("null partition");
    if (templateId < 0 || templateId >= (int64_t)p->tmpls.size() || p->tmpls[templateId].nnz == 0) return fail("hb2_set_rate_template has not been called (template %lld)", (long long)templateId);
    if (n < 0 || (n > 0 && (!nodeIds || !formulaValues))) return fail("bad compiled matrix list");
    if (cat < 0) cat = 0;
    if (cat >= p->C) return fail("rate class %lld out of range (C=%lld)", (long long)cat, (long long)p->C);
    hb2_partition::Tmpl &t = p->tmpls[templateId];
    const bool owned = cat >= p->own0 && cat < p->own0 + p->ownN;
    for (int64_t k = 0; k < n; k++) {
        if (nodeIds[k] < 0 || nodeIds[k] >= p->B) return fail("node id %lld has no branch", (long long)nodeIds[k]);
        p->have_matrix[cat * p->B + nodeIds[k]] = 1;
        note_matrix_change(p, nodeIds[k]);
    }
    if (!owned) return 0;                                   // another class group's matrices
    if (wait_staging(p)) return 1;
    const size_t row = (size_t)t.nF * sizeof(double);
    for (int64_t k = 0; k < n; k++) {
        const int64_t slot = cat * p->B + nodeIds[k];
        int64_t at = claim_slot(p, slot, (int)templateId);
        if (at < 0) {
            if (t.n_pending == p->q_capacity) { cudaSetDevice(p->device); if (flush_compiled(p) || wait_staging(p)) return 1; }
            at = t.n_pending++;
            p->pend_pos[slot] = (int)(-at - 2);
            p->pend_tmpl[slot] = (signed char)templateId;
        }
        memcpy(t.h_V + at * t.nF, formulaValues + k * t.nF, row);
        t.h_vdst[at] = (int)slot;
    }
    return 0;
}

int hb2_set_matrices_compiled(hb2_partition *p, int64_t cat, int64_t n, const int64_t *nodeIds, const double *formulaValues) {
    return hb2_set_matrices_compiled_id(p, 0, cat, n, nodeIds, formulaValues);
}

int hb2_evaluate(hb2_partition *p, int64_t cat, int64_t nUpdate, const int64_t *updateNodes, const double *rootFreqs,
                 double *lnL, double *siteL, int64_t *siteScale) {
    if (!p) return fail("null partition");
    if (cat < 0) cat = 0;
    if (cat >= p->C) return fail("rate class %lld out of range", (long long)cat);
    return evaluate_impl(p, (int)cat, 1, nullptr, nUpdate, updateNodes, rootFreqs, lnL, siteL, siteScale);
}

int hb2_evaluate_forced(hb2_partition *p, int64_t cat, int64_t nUpdate, const int64_t *updateNodes, const double *rootFreqs,
                        int64_t forcedNode, const int64_t *forcedStates, double *lnL, double *siteL, int64_t *siteScale) {
    if (!p) return fail("null partition");
    if (cat < 0) cat = 0;
    if (cat >= p->C) return fail("rate class %lld out of range", (long long)cat);
    if (p->cg_G > 1) return fail("forced states are not available with class groups");
    return evaluate_impl(p, (int)cat, 1, nullptr, nUpdate, updateNodes, rootFreqs, lnL, siteL, siteScale, forcedNode, forcedStates);
}

int hb2_evaluate_classes(hb2_partition *p, const double *weights, int64_t nUpdate, const int64_t *updateNodes,
                         const double *rootFreqs, double *lnL, double *siteL, int64_t *siteScale) {
    if (!p) return fail("null partition");
    if (!weights) return fail("weights must not be null");
    return evaluate_impl(p, 0, (int)p->C, weights, nUpdate, updateNodes, rootFreqs, lnL, siteL, siteScale);
}

int hb2_batch_site_likelihoods(hb2_partition *p, int64_t templateId, int64_t nSets, const int64_t *patternOf,
                               const double *formulaValues, const int64_t *branchGroup, const double *rootFreqs, double *siteLnL) {
    if (!p || !patternOf || !formulaValues || !rootFreqs || !siteLnL) return fail("null argument");
    if (nSets < 0) return fail("bad set count");
    if (p->Dp != 64) return fail("hb2_batch_site_likelihoods supports 33..64-state (codon) partitions only");
    if (templateId < 0 || templateId >= (int64_t)p->tmpls.size() || p->tmpls[templateId].nnz == 0) return fail("template %lld has not been defined", (long long)templateId);
    if (p->cg_G > 1) return fail("hb2_batch_site_likelihoods is not available with class groups (shard the SETS across ranks instead)");
    int gps = 1;                                // shared-powers groups per set: branches with the same branchGroup share a direction
    if (branchGroup)
        for (int64_t b = 0; b < p->B; b++) {
            if (branchGroup[b] < 0 || branchGroup[b] >= 8) return fail("branchGroup[%lld]=%lld outside 0..7", (long long)b, (long long)branchGroup[b]);
            gps = std::max(gps, (int)branchGroup[b] + 1);
        }
    for (int64_t s = 0; s < nSets; s++)
        if (patternOf[s] < 0 || patternOf[s] >= p->S) return fail("patternOf[%lld]=%lld out of range", (long long)s, (long long)patternOf[s]);
    if (nSets == 0) return 0;
    CU(cudaSetDevice(p->device));
    const hb2_partition::Tmpl &t = p->tmpls[templateId];
    const int64_t B = p->B, nF = t.nF;
    // chunk buffers: P matrices of `cap` sets (32 KB per branch) capped at ~8 GB
    int64_t cap = std::min<int64_t>(nSets, std::max<int64_t>(1, (int64_t)(8.0e9 / ((double)B * (4096 + 1024 + (double)nF) * 8))));
    cap = std::min<int64_t>(cap, 1024);
    if (cap > p->b_cap || gps > p->b_groups_per_set) {
        CU(cudaStreamSynchronize(p->stream));
        hb2_partition::ExBuf &x = p->bex;
        for (void *d : {(void *)x.d_int, (void *)x.d_flag, (void *)x.d_weight, (void *)x.d_groups, (void *)x.d_pow, (void *)x.d_refvec, (void *)x.d_flags, (void *)x.d_colsum,
                        (void *)p->d_bPT, (void *)p->d_bV, (void *)p->d_bcond, (void *)p->d_bout, (void *)p->d_bdst, (void *)p->d_bpat, (void *)p->d_bnodeex}) if (d) cudaFree(d);
        if (x.h_int) cudaFreeHost(x.h_int);
        if (p->h_bdst) cudaFreeHost(p->h_bdst);
        x = hb2_partition::ExBuf();
        p->d_bPT = p->d_bV = p->d_bcond = p->d_bout = nullptr; p->d_bdst = p->d_bpat = p->d_bnodeex = nullptr; p->h_bdst = nullptr; p->b_cap = 0;
        const int64_t ne = cap * B;
        x.G = (int)(cap * gps); x.cap = ne; x.stride = 4096; x.kind.assign(x.G, 0);
        CU(cudaMalloc(&x.d_int, (size_t)(2 * ne + x.G) * sizeof(int)));
        CU(cudaMallocHost(&x.h_int, (size_t)(2 * ne + x.G) * sizeof(int)));
        CU(cudaMalloc(&x.d_flag, (size_t)ne * sizeof(int)));
        CU(cudaMalloc(&x.d_weight, (size_t)ne * sizeof(double)));
        CU(cudaMalloc(&x.d_colsum, (size_t)ne * 16 * 64 * sizeof(double)));
        CU(cudaMalloc(&x.d_groups, (size_t)x.G * sizeof(hb2::ExpmGroup)));
        CU(cudaMemset(x.d_groups, 0, (size_t)x.G * sizeof(hb2::ExpmGroup)));
        CU(cudaMalloc(&x.d_pow, (size_t)x.G * hb2::EXPM_POW_TERMS * 4096 * sizeof(double)));
        CU(cudaMalloc(&x.d_flags, (size_t)x.G * hb2::EXPM_POW_TERMS * sizeof(unsigned long long)));
        CU(cudaMemset(x.d_flags, 0, (size_t)x.G * hb2::EXPM_POW_TERMS * sizeof(unsigned long long)));
        CU(cudaMalloc(&x.d_refvec, (size_t)x.G * x.stride * sizeof(double)));
        CU(cudaMalloc(&p->d_bPT, (size_t)ne * 4096 * sizeof(double)));
        CU(cudaMalloc(&p->d_bV, (size_t)ne * nF * sizeof(double)));
        CU(cudaMalloc(&p->d_bcond, (size_t)cap * p->I * 64 * sizeof(double)));
        CU(cudaMalloc(&p->d_bnodeex, (size_t)cap * p->I * sizeof(int)));
        CU(cudaMalloc(&p->d_bout, (size_t)cap * sizeof(double)));
        CU(cudaMalloc(&p->d_bdst, (size_t)ne * sizeof(int)));
        CU(cudaMalloc(&p->d_bpat, (size_t)cap * sizeof(int)));
        CU(cudaMallocHost(&p->h_bdst, (size_t)(ne + cap) * sizeof(int)));
        p->b_cap = cap; p->b_groups_per_set = gps;
    }
    double *hs = p->h_small;
    CU(cudaStreamSynchronize(p->stream));
    for (int k = 0; k < p->Dp; k++) hs[k] = k < p->D ? rootFreqs[k] : 0.0;
    CU(cudaMemcpyAsync(p->d_pi, hs, p->Dp * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    std::vector<int> grp;
    for (int64_t s0 = 0; s0 < nSets; s0 += p->b_cap) {
        const int64_t nb = std::min<int64_t>(p->b_cap, nSets - s0), ne = nb * B;
        CU(cudaStreamSynchronize(p->stream));          // previous chunk done: buffers and pinned arrays are free again
        for (int64_t e = 0; e < ne; e++) p->h_bdst[e] = (int)e;                      // slot in the chunk's P area: set-major
        for (int64_t s = 0; s < nb; s++) p->h_bdst[ne + s] = (int)patternOf[s0 + s];
        CU(cudaMemcpyAsync(p->d_bdst, p->h_bdst, (size_t)ne * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        CU(cudaMemcpyAsync(p->d_bpat, p->h_bdst + ne, (size_t)nb * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        CU(cudaMemcpyAsync(p->d_bV, formulaValues + (size_t)s0 * B * nF, (size_t)ne * nF * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        grp.resize(ne);
        for (int64_t s = 0; s < nb; s++)
            for (int64_t b = 0; b < B; b++) grp[s * B + b] = (int)(s * p->b_groups_per_set + (branchGroup ? branchGroup[b] : 0));
        SharedPlan sp;
        std::fill(p->bex.kind.begin(), p->bex.kind.end(), 0);                     // every chunk rebuilds its tables
        if (plan_shared(p, p->h_bdst, ne, 1, grp.data(), sp, (int)templateId, &p->bex, 4)) return 1;
        CU(cudaEventRecord(p->ev_staging, p->stream));
        p->staging_busy = true;
        // classification writes its "resident copy" of the formula values in place (res = input) -- harmless
        if (launch_expm(p, p->d_bV, p->d_bdst, (int)ne, 2, nullptr, false, p->d_bPT, &sp, 0, (int)templateId)) return 1;
        hb2::BatchPruneArgs ba;
        ba.PT = p->d_bPT; ba.cond = p->d_bcond; ba.node_ex = p->d_bnodeex; ba.pat = p->d_bpat; ba.leaf = p->d_leaf; ba.ambig = p->d_ambig; ba.pi = p->d_pi;
        ba.tree.child_start = p->d_child_start; ba.tree.child_ids = p->d_child_ids; ba.out = p->d_bout;
        ba.L = (int)p->L; ba.I = (int)p->I; ba.B = (int)B; ba.D = (int)p->D; ba.Sp = (int)p->Sp;
        hb2::prune_batch_kernel<<<(unsigned)nb, 64, 0, p->stream>>>(ba);
        p->launches++;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(siteLnL + s0, p->d_bout, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    }
    int err = 0;
    CU(cudaMemcpyAsync(&err, p->d_err, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    if (err) { cudaMemsetAsync(p->d_err, 0, sizeof(int), p->stream); return fail("a device-side wait timed out in the batched exponential (code %d)", err); }
    return 0;
}

int hb2_read_conditionals(hb2_partition *p, int64_t cat, int64_t inode, double *cond, int32_t *exp2) {
    if (!p || !cond || !exp2) return fail("null argument");
    if (cat < 0) cat = 0;
    if (cat >= p->C || inode < 0 || inode >= p->I) return fail("index out of range");
    CU(cudaSetDevice(p->device));
    std::vector<double> tmp((size_t)p->Sp * p->Dp);
    std::vector<int> te(p->Sp);
    const int64_t nstride = (p->use_tc && p->use_walk) ? 2 * p->I : p->I;     // node slots per class (walk: + side products)
    CU(cudaStreamSynchronize(p->stream));
    if (p->use_tc) {
        std::vector<float> tf(tmp.size());
        CU(cudaMemcpy(tf.data(), p->d_condf + ((size_t)cat * nstride + inode) * p->Sp * 64, tf.size() * sizeof(float), cudaMemcpyDeviceToHost));
        // device layout per tile of 128 patterns: [16 chunks][128 patterns][4 floats]
        for (int64_t s = 0; s < p->Sp; s++)
            for (int k = 0; k < 64; k++)
                tmp[s * 64 + k] = std::fabs(tf[(((size_t)(s / 128) * 16 + k / 4) * 128 + s % 128) * 4 + k % 4]);   // sign bit = walk tag
    } else
    CU(cudaMemcpy(tmp.data(), p->d_cond + ((size_t)cat * p->I + inode) * p->Sp * p->Dp, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(te.data(), p->d_scal + ((size_t)cat * nstride + inode) * p->Sp, te.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int64_t s = 0; s < p->S; s++) {
        for (int64_t k = 0; k < p->D; k++) cond[s * p->D + k] = tmp[s * p->Dp + k];
        exp2[s] = (p->use_tc && p->use_walk) ? (te[s] >> 1) : te[s];      // walk path: bit 0 is the generation tag
    }
    return 0;
}

int hb2_read_transition(hb2_partition *p, int64_t cat, int64_t node, double *P) {
    if (!p || !P) return fail("null argument");
    if (cat < 0) cat = 0;
    if (cat >= p->C || node < 0 || node >= p->B) return fail("index out of range");
    CU(cudaSetDevice(p->device));
    if (flush_matrices(p)) return 1;
    std::vector<double> tmp((size_t)p->Dp * p->Dp);
    CU(cudaStreamSynchronize(p->stream));
    CU(cudaMemcpy(tmp.data(), p->d_PT + ((size_t)cat * p->B + node) * p->Dp * p->Dp, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < p->D; i++)
        for (int64_t j = 0; j < p->D; j++) P[i * p->D + j] = tmp[j * p->Dp + i];   // stored transposed
    return 0;
}

static hb2::BranchCacheArgs bc_args(hb2_partition *p, int c0, int nc) {
    hb2::BranchCacheArgs a{};
    a.cv.c64 = p->use_tc ? nullptr : p->d_cond;
    a.cv.c32 = p->use_tc ? p->d_condf : nullptr;
    a.cv.scal = p->d_scal; a.cv.I = (int)((p->use_tc && p->use_walk) ? 2 * p->I : p->I); a.cv.Sp = (int)p->Sp; a.cv.Dp = p->Dp; a.cv.tagged = (p->use_tc && p->use_walk) ? 1 : 0;
    a.PT = p->d_PT; a.leaf = p->d_leaf; a.ambig = p->d_ambig; a.pi = p->d_pi; a.out = p->d_bc_out; a.outE = p->d_bc_outE;
    a.L = (int)p->L; a.B = (int)p->B; a.D = (int)p->D; a.Dp = p->Dp; a.Sp = (int)p->Sp; a.S = (int)p->S; a.cat0 = c0; a.ncls = nc;
    return a;
}

int hb2_branch_cache_build(hb2_partition *p, int64_t node, const double *rootFreqs) {
    if (!p || !rootFreqs) return fail("null argument");
    if (node < 0 || node >= p->B) return fail("node id %lld has no branch (valid 0..%lld)", (long long)node, (long long)p->B - 1);
    CU(cudaSetDevice(p->device));
    for (int c = p->own0; c < p->own0 + p->ownN; c++)
        if (!p->evaluated_cat[c]) return fail("hb2_branch_cache_build: rate class %d has never been evaluated (no resident conditionals)", c);
    if (flush_matrices(p)) return 1;
    if (!p->d_bc_out) {
        CU(cudaMalloc(&p->d_bc_out, (size_t)p->C * p->Sp * p->Dp * sizeof(double)));
        CU(cudaMalloc(&p->d_bc_outE, (size_t)p->C * p->Sp * sizeof(int)));
        CU(cudaMalloc(&p->d_bc_sib, (size_t)(3 * (p->L + p->I) + 2) * sizeof(int)));
    }
    double *hs = p->h_small;
    for (int k = 0; k < p->Dp; k++) hs[k] = k < p->D ? rootFreqs[k] : 0.0;
    CU(cudaMemcpyAsync(p->d_pi, hs, p->Dp * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    // root -> parent(node) path (internal indices) and, per path node, its children that are NOT on the path
    std::vector<int> path;
    for (int64_t u = p->parents[node]; u >= 0; u = p->parents[p->L + u]) path.push_back((int)u);
    std::reverse(path.begin(), path.end());
    std::vector<int> sib, off(path.size() + 1, 0), down(path.size(), -1);
    for (size_t i = 0; i < path.size(); i++) {
        const int on_path = (i + 1 < path.size()) ? (int)p->L + path[i + 1] : (int)node;
        for (int ch : p->children[path[i]]) if (ch != on_path) sib.push_back(ch);
        off[i + 1] = (int)sib.size();
        down[i] = (i + 1 < path.size()) ? on_path : -1;
    }
    CU(cudaStreamSynchronize(p->stream));            // h_small / path upload below are synchronous with respect to earlier work
    // one launch for the whole path: [siblings | sibling offsets (npath+1) | path child below each path node (npath)]
    std::vector<int> pk(sib);
    const int o_off = (int)pk.size();
    pk.insert(pk.end(), off.begin(), off.end());
    const int o_down = (int)pk.size();
    pk.insert(pk.end(), down.begin(), down.end());
    if ((int64_t)pk.size() > 3 * (p->L + p->I) + 2) return fail("branch-cache path does not fit its buffer");
    CU(cudaMemcpy(p->d_bc_sib, pk.data(), pk.size() * sizeof(int), cudaMemcpyHostToDevice));
    hb2::BranchCacheArgs a = bc_args(p, p->own0, p->ownN);
    hb2::BranchPathArgs bp;
    bp.sib = p->d_bc_sib; bp.sib_off = p->d_bc_sib + o_off; bp.down = p->d_bc_sib + o_down; bp.npath = (int)path.size();
    dim3 grid((unsigned)p->S, (unsigned)p->ownN);
    if (p->Dp > 32) hb2::bc_path_kernel<64><<<grid, 64, 0, p->stream>>>(a, bp);
    else hb2::bc_path_kernel<32><<<grid, 32, 0, p->stream>>>(a, bp);
    p->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(p->stream));
    p->bc_node = node;
    p->bc_dirty_node = -1;
    return 0;
}

int hb2_branch_cache_evaluate(hb2_partition *p, int64_t cat, const double *weights, double *lnL, double *siteL, int64_t *siteScale) {
    if (!p || !lnL) return fail("null argument");
    if (p->bc_node < 0) return fail("hb2_branch_cache_evaluate: no valid branch cache (build it after the last full evaluation)");
    CU(cudaSetDevice(p->device));
    // only the cached branch may have a new matrix (flushed or not, plain, compiled or mixture): anything else needs a
    // regular evaluation
    if (p->bc_dirty_node >= 0)
        return fail("branch cache holds node %lld but a matrix of node %lld changed", (long long)p->bc_node, (long long)p->bc_dirty_node);
    if (flush_matrices(p)) return 1;
    int c0, nc;
    if (weights) { c0 = p->own0; nc = p->ownN; }
    else {
        if (p->cg_G > 1) return fail("with class groups only the all-class form (weights != NULL) is available");
        if (cat < 0) cat = 0;
        if (cat >= p->C) return fail("rate class %lld out of range", (long long)cat);
        c0 = (int)cat; nc = 1;
    }
    double *hs = p->h_small;
    if (weights) {
        for (int c = 0; c < p->C; c++) hs[p->Dp + c] = weights[c];
        CU(cudaMemcpyAsync(p->d_weights, hs + p->Dp, p->C * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    }
    hb2::BranchCacheArgs a = bc_args(p, c0, nc);
    dim3 grid((unsigned)((p->S + 127) / 128), (unsigned)nc);
    hb2::bc_eval_kernel<<<grid, 128, 0, p->stream>>>(a, (int)p->bc_node, p->d_rootL, p->d_rootE);
    p->launches++;
    CU(cudaGetLastError());
    const bool want_sites = siteL != nullptr || siteScale != nullptr;
    if (run_root(p, c0, nc, weights != nullptr, want_sites)) return 1;
    CU(cudaMemcpyAsync(hs + p->Dp + p->C, p->d_lnL, sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (siteL) CU(cudaMemcpyAsync(siteL, p->d_siteL, p->S * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (siteScale) CU(cudaMemcpyAsync(siteScale, p->d_siteScale, p->S * sizeof(long long), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    *lnL = hs[p->Dp + p->C];
    return 0;
}

int hb2_plan_walk(int64_t L, int64_t I, const int64_t *flatParents, int64_t nUpdate, const int64_t *updateNodes, int lanes,
                  int splitNodes, int32_t *laneStart, int32_t *steps, int64_t stepCapacity, int64_t *nSteps) {
    if (!flatParents || !laneStart || !steps || !nSteps) return fail("null argument");
    if (L < 2 || I < 1 || lanes < 1 || lanes > 15) return fail("bad sizes L=%lld I=%lld lanes=%d", (long long)L, (long long)I, lanes);
    std::vector<std::vector<int>> children(I);
    for (int64_t n = 0; n < L + I - 1; n++) {
        const int64_t par = flatParents[n];
        if (par < 0 || par >= I || (n >= L && par <= n - L)) return fail("flatParents[%lld]=%lld: not a post-ordered tree", (long long)n, (long long)par);
        children[par].push_back((int)n);
    }
    if (flatParents[L + I - 1] != -1) return fail("root must be the last node");
    std::vector<int> height(I, 0);
    for (int64_t i = 0; i < I; i++) {
        if (children[i].empty()) return fail("internal node %lld has no children", (long long)i);
        for (int ch : children[i]) if (ch >= L) height[i] = std::max(height[i], height[ch - L] + 1);
    }
    std::vector<char> dirty(I, 0);
    if (!updateNodes || nUpdate < 0) std::fill(dirty.begin(), dirty.end(), 1);
    else
        for (int64_t k = 0; k < nUpdate; k++) {                       // same closure as hb2_evaluate: parents and ancestors
            if (updateNodes[k] < 0 || updateNodes[k] >= L + I) return fail("updateNodes[%lld] out of range", (long long)k);
            for (int64_t par = flatParents[updateNodes[k]]; par >= 0 && !dirty[par]; par = flatParents[L + par]) dirty[par] = 1;
        }
    int total = 0;
    for (char d : dirty) total += d;
    if (total == 0) { *nSteps = 0; for (int r = 0; r <= lanes; r++) laneStart[r] = 0; return 0; }
    const int K = std::min(lanes, total);
    if (stepCapacity < L + 2 * I) return fail("steps needs room for %lld entries", (long long)(L + 2 * I));
    std::vector<int> ls(K + 1), st(2 * (size_t)(L + 2 * I));
    std::vector<char> jdirty;
    const int ns = plan_walk(children, height, (int)L, (int)I, dirty, K, (splitNodes & 1) != 0, ls.data(), st.data(), jdirty, (splitNodes & 2) != 0);
    for (int r = 0; r <= lanes; r++) laneStart[r] = ls[std::min(r, K)];
    for (int i = 0; i < 2 * ns; i++) steps[i] = st[i];
    *nSteps = ns;
    return 0;
}

int hb2_comm_unique_id(void *uniqueId128) {
    if (!uniqueId128) return fail("null id buffer");
    if (!g_nccl.load()) return fail("cannot load libnccl.so.2: %s", dlerror());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail("ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
    memcpy(uniqueId128, &id, 128);
    return 0;
}

int hb2_comm_init(hb2_partition *p, int nRanks, int rank, const void *uniqueId128) {
    if (!p || !uniqueId128) return fail("null argument");
    if (nRanks < 1 || rank < 0 || rank >= nRanks) return fail("bad rank %d of %d", rank, nRanks);
    if (!g_nccl.load()) return fail("cannot load libnccl.so.2: %s", dlerror());
    CU(cudaSetDevice(p->device));
    if (p->comm) { g_nccl.CommDestroy(p->comm); p->comm = nullptr; }      // re-initialisation replaces the communicator
    ncclUniqueId id;
    memcpy(&id, uniqueId128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&p->comm, nRanks, id, rank);
    if (r != ncclSuccess) { p->comm = nullptr; return fail("ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
    p->n_ranks = nRanks; p->rank = rank;
    return peer_setup(p, 2);
}

int hb2_comm_gather_sites(hb2_partition *p, const double *siteL, const int64_t *siteScale, double *allSiteL, int64_t *allSiteScale,
                          int64_t capacity, int64_t *total) {
    if (!p || !siteL || !siteScale || !allSiteL || !allSiteScale || !total) return fail("null argument");
    if (!p->comm) return fail("hb2_comm_gather_sites needs hb2_comm_init first");
    CU(cudaSetDevice(p->device));
    const int R = p->n_ranks, G = std::max(p->cg_G, 1);
    // 1. every rank's pattern count
    long long mine = (long long)p->S, *d_cnt = nullptr;
    std::vector<long long> cnt(R);
    CU(cudaMalloc(&d_cnt, (size_t)(R + 1) * sizeof(long long)));
    CU(cudaMemcpyAsync(d_cnt + R, &mine, sizeof(long long), cudaMemcpyHostToDevice, p->stream));
    ncclResult_t nr = g_nccl.AllGather(d_cnt + R, d_cnt, sizeof(long long), ncclChar, p->comm, p->stream);
    if (nr != ncclSuccess) { cudaFree(d_cnt); return fail("ncclAllGather (pattern counts): %s", g_nccl.GetErrorString(nr)); }
    CU(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)R * sizeof(long long), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    cudaFree(d_cnt);
    long long mx = 0, tot = 0;
    for (int r = 0; r < R; r++) { mx = std::max(mx, cnt[r]); if (r % G == 0) tot += cnt[r]; }    // ranks of one shard hold the same patterns
    *total = (int64_t)tot;
    if (tot > capacity) return fail("hb2_comm_gather_sites: %lld patterns in all shards, room for %lld", tot, (long long)capacity);
    // 2. (likelihood, scaler count) pairs, padded to the largest shard
    const size_t rec = (size_t)mx * 16;
    std::vector<unsigned char> send(rec, 0), all(rec * R);
    for (long long s = 0; s < mine; s++) { memcpy(&send[(size_t)s * 16], siteL + s, 8); memcpy(&send[(size_t)s * 16 + 8], siteScale + s, 8); }
    unsigned char *d_send = nullptr, *d_all = nullptr;
    CU(cudaMalloc(&d_send, std::max<size_t>(rec, 16)));
    CU(cudaMalloc(&d_all, std::max<size_t>(rec * R, 16)));
    CU(cudaMemcpyAsync(d_send, send.data(), rec, cudaMemcpyHostToDevice, p->stream));
    nr = g_nccl.AllGather(d_send, d_all, rec, ncclChar, p->comm, p->stream);
    if (nr != ncclSuccess) { cudaFree(d_send); cudaFree(d_all); return fail("ncclAllGather (per-pattern outputs): %s", g_nccl.GetErrorString(nr)); }
    CU(cudaMemcpyAsync(all.data(), d_all, rec * R, cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    cudaFree(d_send); cudaFree(d_all);
    // 3. shard order = rank order of the first rank of every shard
    int64_t at = 0;
    for (int r = 0; r < R; r += G)
        for (long long s = 0; s < cnt[r]; s++, at++) {
            memcpy(allSiteL + at, &all[(size_t)r * rec + (size_t)s * 16], 8);
            memcpy(allSiteScale + at, &all[(size_t)r * rec + (size_t)s * 16 + 8], 8);
        }
    return 0;
}

int hb2_comm_class_groups(hb2_partition *p, int nGroups) {
    if (!p) return fail("null partition");
    if (!p->comm) return fail("hb2_comm_class_groups needs hb2_comm_init first");
    if (nGroups < 1 || p->C % nGroups != 0 || p->n_ranks % nGroups != 0)
        return fail("class groups: %d must divide both the %lld rate classes and the %d ranks", nGroups, (long long)p->C, p->n_ranks);
    if (p->n_pending) return fail("hb2_comm_class_groups must be called before matrices are set");
    for (char h : p->have_matrix) if (h) return fail("hb2_comm_class_groups must be called before matrices are set");
    CU(cudaSetDevice(p->device));
    p->cg_G = nGroups; p->cg_g = p->rank % nGroups;
    p->ownN = (int)(p->C / nGroups); p->own0 = p->cg_g * p->ownN;
    if (nGroups == 1) return 0;
    // exchange length: the longest padded shard among the ranks (shards of the same communicator may differ in size)
    int *d_len = nullptr;
    CU(cudaMalloc(&d_len, sizeof(int)));
    int len = (int)p->Sp;
    CU(cudaMemcpyAsync(d_len, &len, sizeof(int), cudaMemcpyHostToDevice, p->stream));
    ncclResult_t r = g_nccl.AllReduce(d_len, d_len, 1, ncclInt32, ncclMax, p->comm, p->stream);
    if (r != ncclSuccess) { cudaFree(d_len); return fail("ncclAllReduce(max): %s", g_nccl.GetErrorString(r)); }
    CU(cudaMemcpyAsync(&len, d_len, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    cudaFree(d_len);
    p->xchg_len = len;
    for (double **d : {&p->d_xsend, &p->d_xrecv, &p->d_xfreq, &p->d_xpartial}) if (*d) { cudaFree(*d); *d = nullptr; }
    const size_t tot = (size_t)2 * len * p->n_ranks;
    CU(cudaMalloc(&p->d_xsend, tot * sizeof(double)));
    CU(cudaMalloc(&p->d_xrecv, tot * sizeof(double)));
    CU(cudaMalloc(&p->d_xfreq, (size_t)len * p->n_ranks * sizeof(double)));
    CU(cudaMalloc(&p->d_xpartial, ((size_t)len * p->n_ranks / 256 + 2) * sizeof(double)));
    // pattern frequencies of every rank, gathered once with the same zero-elsewhere sum (d_xsend doubles as scratch)
    CU(cudaMemsetAsync(p->d_xsend, 0, tot * sizeof(double), p->stream));
    CU(cudaMemcpyAsync(p->d_xsend + (size_t)p->rank * len, p->d_freq, (size_t)std::min<int64_t>(p->Sp, len) * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    r = g_nccl.AllReduce(p->d_xsend, p->d_xfreq, (size_t)len * p->n_ranks, ncclDouble, ncclSum, p->comm, p->stream);
    if (r != ncclSuccess) return fail("ncclAllReduce (frequencies): %s", g_nccl.GetErrorString(r));
    CU(cudaMemsetAsync(p->d_xsend, 0, tot * sizeof(double), p->stream));     // from now on only this rank's slot is written
    CU(cudaStreamSynchronize(p->stream));
    return peer_setup(p, 2 * len);
}

void hb2_destroy(hb2_partition *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    peer_teardown(p);
    if (p->comm) g_nccl.CommDestroy(p->comm);
    void *dev[] = {p->d_leaf, p->d_scal, p->d_rootE, p->d_child_start, p->d_child_ids, p->d_jobs, p->d_dst, p->d_flag,
                   p->d_mix_dst, p->d_mix_Q, p->d_ambig, p->d_freq, p->d_cond, p->d_PT, p->d_Q, p->d_pi, p->d_rootL, p->d_weights,
                   p->d_partial, p->d_lnL, p->d_siteL, p->d_siteScale, p->d_Qres, p->d_condf, p->d_PB, p->d_PTf, p->d_err, p->d_mix_scratch, p->d_xsend, p->d_xrecv, p->d_xfreq, p->d_xpartial, p->d_bc_out, p->d_bc_outE, p->d_bc_sib, p->d_walk, p->d_forced, p->d_cond_side, p->d_lane_flags, p->d_root_counter, p->ex.d_int, p->ex.d_flag, p->ex.d_weight, p->ex.d_groups, p->ex.d_pow, p->ex.d_refvec, p->ex.d_flags, p->ex.d_colsum, p->bex.d_colsum, p->bex.d_int, p->bex.d_flag, p->bex.d_weight, p->bex.d_groups, p->bex.d_pow, p->bex.d_refvec, p->bex.d_flags,
                   p->d_bPT, p->d_bV, p->d_bcond, p->d_bout, p->d_bdst, p->d_bpat, p->d_bnodeex};
    for (void *d : dev) if (d) cudaFree(d);
    if (p->h_Q) cudaFreeHost(p->h_Q);
    if (p->h_small) cudaFreeHost(p->h_small);
    if (p->h_jobs) cudaFreeHost(p->h_jobs);
    if (p->h_dst) cudaFreeHost(p->h_dst);
    for (auto &t : p->tmpls) free_template(t);
    if (p->h_walk) cudaFreeHost(p->h_walk);
    if (p->h_mix) cudaFreeHost(p->h_mix);
    if (p->h_forced) cudaFreeHost(p->h_forced);
    if (p->ex.h_int) cudaFreeHost(p->ex.h_int);
    if (p->bex.h_int) cudaFreeHost(p->bex.h_int);
    if (p->h_bdst) cudaFreeHost(p->h_bdst);
    if (p->ev_mix) cudaEventDestroy(p->ev_mix);
    for (auto &e : p->ev) if (e) cudaEventDestroy(e);
    if (p->ev_staging) cudaEventDestroy(p->ev_staging);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

int64_t hb2_launch_count(const hb2_partition *p) { return p ? p->launches : 0; }
int hb2_precision_mode(const hb2_partition *p) { return (p && p->use_tc) ? 1 : 0; }
int hb2_stage_launches(const hb2_partition *p, int64_t *out3) {
    if (!p || !out3) return fail("null argument");
    for (int k = 0; k < 3; k++) out3[k] = p->stage_launches[k];
    return 0;
}
const char *hb2_pruning_kernel(const hb2_partition *p) {
    if (!p) return "";
    if (p->use_tc) return p->use_walk ? (p->walk_v2 ? "prune64_tc_walk2_kernel" : "prune64_tc_walk_kernel") : "prune64_tc_kernel";
    if (p->Dp == 64 && p->fp64_mode == 2) return "prune64_lanes_kernel";
    if (p->Dp == 64) return (p->fp64_walk && (int64_t)(p->Sp / hb2::TILE_P) * p->ownN >= 2 * p->sm_count) ? "prune64_walk_kernel" : "prune64_kernel";
    if (p->small_walk && p->small_dmma && p->Dp >= 16) return "prune_small_dmma_kernel";
    if (p->small_walk && small_two_patterns(p, p->ownN)) return "prune_small_walk_ilp_kernel";
    return p->small_walk ? "prune_small_walk_kernel" : "prune_small_kernel";
}

int hb2_time_resident(hb2_partition *p, const double *weights, const double *rootFreqs, int iters, double *msPerEval,
                      double *stageMs, double *lnL) {
    if (!p || !weights || !rootFreqs || !msPerEval) return fail("null argument");
    if (iters < 1) return fail("iters must be >= 1");
    CU(cudaSetDevice(p->device));
    if (check_ready(p, 0, (int)p->C)) return 1;
    if (flush_matrices(p)) return 1;
    p->bc_node = -1; p->bc_dirty_node = -1;
    // every slot must hold a resident copy of its rate matrix (dense, or the formula values of a compiled hand-over) so
    // that the expm stage can be replayed
    std::vector<int> dst;
    int res_kind = 0;
    for (int64_t k = (int64_t)p->own0 * p->B; k < (int64_t)(p->own0 + p->ownN) * p->B; k++) {
        if (!p->is_rate[k]) return fail("hb2_time_resident needs HB2_MATRIX_RATE matrices in every slot");
        if (res_kind && res_kind != p->is_rate[k]) return fail("hb2_time_resident: hand all matrices over the same way (dense or compiled) before timing");
        if (p->is_rate[k] == 2 && p->res_tmpl[k] != 0) return fail("hb2_time_resident replays template 0 only");
        res_kind = p->is_rate[k];
        dst.push_back((int)k);
    }
    if (wait_staging(p)) return 1;           // the flush above may still be reading the pinned queues
    memcpy(p->h_dst, dst.data(), dst.size() * sizeof(int));
    CU(cudaMemcpyAsync(p->d_dst, p->h_dst, dst.size() * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    SharedPlan sp;
    if (plan_shared(p, p->h_dst, (int64_t)dst.size(), res_kind == 2 ? 1 : 2, nullptr, sp)) return 1;
    if (res_kind == 2 && !sp.use) return fail("hb2_time_resident: compiled resident copies need the shared-powers path");
    CU(cudaEventRecord(p->ev_staging, p->stream));
    p->staging_busy = true;
    double *hs = p->h_small;
    for (int k = 0; k < p->Dp; k++) hs[k] = k < p->D ? rootFreqs[k] : 0.0;
    for (int c = 0; c < p->C; c++) hs[p->Dp + c] = weights[c];
    CU(cudaMemcpyAsync(p->d_pi, hs, p->Dp * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CU(cudaMemcpyAsync(p->d_weights, hs + p->Dp, p->C * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    std::vector<std::vector<int>> levels;
    if (plan_levels(p, -1, nullptr, levels)) return 1;
    double tot = 0, st[3] = {0, 0, 0};
    for (int it = 0; it < iters; it++) {
        CU(cudaEventRecord(p->ev[0], p->stream));
        const int64_t l0 = p->launches;
        if (res_kind == 2) {
            if (launch_expm(p, p->tmpls[0].d_Vres + (size_t)p->own0 * p->B * p->tmpls[0].nF, p->d_dst, (int)dst.size(), 2, nullptr, true, nullptr, &sp)) return 1;
        } else if (launch_expm(p, p->d_Qres + (size_t)p->own0 * p->B * p->D * p->D, p->d_dst, (int)dst.size(), 0, nullptr, true, nullptr, &sp)) return 1;
        const int64_t l1 = p->launches;
        CU(cudaEventRecord(p->ev[1], p->stream));
        if (run_pruning(p, p->own0, p->ownN, levels)) return 1;
        CU(cudaEventRecord(p->ev[2], p->stream));
        const int64_t l2 = p->launches;
        if (run_root(p, p->own0, p->ownN, true, false)) return 1;
        p->stage_launches[0] = l1 - l0; p->stage_launches[1] = l2 - l1; p->stage_launches[2] = p->launches - l2;
        CU(cudaEventRecord(p->ev[3], p->stream));
        CU(cudaEventSynchronize(p->ev[3]));
        float a = 0, b = 0, c = 0;
        CU(cudaEventElapsedTime(&a, p->ev[0], p->ev[1]));
        CU(cudaEventElapsedTime(&b, p->ev[1], p->ev[2]));
        CU(cudaEventElapsedTime(&c, p->ev[2], p->ev[3]));
        st[0] += a; st[1] += b; st[2] += c; tot += a + b + c;
    }
    *msPerEval = tot / iters;
    if (stageMs) for (int k = 0; k < 3; k++) stageMs[k] = st[k] / iters;
    if (lnL) {
        CU(cudaMemcpy(lnL, p->d_lnL, sizeof(double), cudaMemcpyDeviceToHost));
    }
    for (int c = 0; c < p->C; c++) p->evaluated_cat[c] = 1;
    return 0;
}

}  // extern "C"
