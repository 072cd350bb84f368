// hb2_hyphy_hooks.cpp -- see hb2_hyphy_hooks.h.  Compiled with the HyPhy build copy's headers.
#include "hb2_hyphy_hooks.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "batchlan.h"
#include "calcnode.h"
#include "formula.h"
#include "dataset_filter.h"
#include "global_things.h"
#include "likefunc.h"
#include "matrix.h"
#include "tree.h"
#include "vector.h"

#include "hyphy_b200.h"

static_assert(sizeof(long) == sizeof(int64_t), "HyPhy's long arrays are passed to the engine as int64_t (LP64)");

// friend of _TheTree and _Matrix (one inserted line each): the flat parent table and the compiled form of a model matrix
// are not public there
struct hb2_hooks_access {
    static long const *flat_parents(_TheTree const *t) { return t->flatParents.list_data; }
    static unsigned long flat_parents_length(_TheTree const *t) { return t->flatParents.lLength; }
    static _CompiledMatrixData *compiled(_Matrix const *m) { return m->cmd; }
    static long const *stored_index(_Matrix const *m) { return m->theIndex; }
    static long stored_count(_Matrix const *m) { return m->lDim; }
};

namespace hb2_hooks {

namespace {

struct Part {
    hb2_partition *h = nullptr;
    _TheTree *tree = nullptr;
    long S = 0, D = 0, L = 0, I = 0, C = 0;
    std::unordered_map<_CalcNode const *, long> node_id;    // tree node -> flat id (leaves, then internal nodes)
    std::mutex lock;                                         // SetCompExp is called from ExponentiateMatrices' OpenMP loop
    // compiled route: one engine template per model matrix seen on this tree
    struct Template {
        int id = -1;
        _CompiledMatrixData const *cmd = nullptr;
        long stored = 0, n_formulas = 0, freq_var = -1;
        unsigned long checked_epoch = 0;
        std::vector<double> freqs;
    };
    std::unordered_map<_Matrix const *, Template> templates;
    unsigned long epoch = 0, n_compiled = 0;                 // epoch: one per ComputeBlock scope (frequencies are re-read once per epoch)
    unsigned long n_eval = 0, n_rate = 0, n_trans = 0;
    double t_handover = 0.0, t_evaluate = 0.0;               // seconds inside hb2_set_matrices* / hb2_evaluate* (HYPHY_B200_VERBOSE)
    std::chrono::steady_clock::time_point t_created;
};

inline double seconds_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

struct State {
    std::vector<Part *> parts;
};

Part *g_current = nullptr;          // ComputeBlock is entered from the single interpreter thread
std::unordered_map<_TheTree const *, Part *> g_tree_part;      // for readers that only know the tree (materialize_tree)
std::mutex g_tree_mu;
int g_bypass = 0;

bool env_true(char const *name) {
    char const *v = getenv(name);
    return v && v[0] && v[0] != '0';
}

bool engine_enabled() {
    char const *v = getenv("HYPHY_B200");
    if (v && v[0] == '0') return false;
    return true;
}

[[noreturn]] void fatal(char const *what) {
    // no CPU fallback exists once a partition is on the engine: every failure is fatal for the host
    hy_global::HandleApplicationError(_String("hyphy_b200 engine: ") & what & " -- " & hb2_last_error(), true);
    abort();
}

}  // namespace

void create(void *&state, unsigned long n_trees, unsigned long index, _TheTree *tree, _DataSetFilter const *filter,
            long const *leaf_flags, _Vector const *ambiguities, long n_ambiguities) {
    if (!engine_enabled()) return;
    State *st = static_cast<State *>(state);
    if (!st) { st = new State(); state = st; }
    if (st->parts.size() < n_trees) st->parts.resize(n_trees, nullptr);
    if (st->parts[index]) {
        { std::lock_guard<std::mutex> lk(g_tree_mu); auto it = g_tree_part.find(st->parts[index]->tree); if (it != g_tree_part.end() && it->second == st->parts[index]) g_tree_part.erase(it); }
        hb2_destroy(st->parts[index]->h); delete st->parts[index]; st->parts[index] = nullptr;
    }
    // two-sequence analyses have no internal-node cache and never reach the pruning branch of ComputeBlock
    // (likefunc.cpp:4216, :11260-11281: ComputeTwoSequenceLikelihood): nothing to take over, the host's own code runs
    if (tree->GetLeafCount() < 2 || tree->GetINodeCount() < 1) return;
    if (hb2_device_count() <= 0) fatal("no CUDA device is visible (set HYPHY_B200=0 to run the CPU path)");

    Part *p = new Part();
    p->tree = tree;
    p->S = (long)filter->GetPatternCount();
    p->D = (long)filter->GetDimension();
    p->L = tree->GetLeafCount();
    p->I = tree->GetINodeCount();
    p->C = tree->categoryCount > 0 ? tree->categoryCount : 1;
    if ((long)hb2_hooks_access::flat_parents_length(tree) != p->L + p->I) fatal("tree has not been flattened (SetUp)");
    for (long k = 0; k < p->L; k++) p->node_id[(_CalcNode const *)tree->flatCLeaves.GetItem(k)] = k;
    for (long k = 0; k < p->I; k++) p->node_id[(_CalcNode const *)tree->flatTree.GetItem(k)] = p->L + k;

    int device = 0;
    if (char const *dv = getenv("HYPHY_B200_DEVICE")) device = atoi(dv);
#ifdef __HYPHYMPI__
    else device = hy_mpi_node_rank % hb2_device_count();
#endif
    // pattern multiplicities: theFrequencies is a _SimpleList of long (dataset_filter.h), original pattern order
    if (hb2_create(&p->h, p->S, p->D, p->L, p->I, p->C, (int64_t const *)hb2_hooks_access::flat_parents(tree),
                   (int64_t const *)leaf_flags, n_ambiguities > 0 ? ambiguities->theData : nullptr, n_ambiguities,
                   (int64_t const *)filter->theFrequencies.list_data, device,
                   env_true("HYPHY_B200_TC") ? HB2_FLAG_DEFAULT : HB2_FLAG_FORCE_FP64)) {
        delete p;
        fatal("hb2_create failed");
    }
    st->parts[index] = p;
    { std::lock_guard<std::mutex> lk(g_tree_mu); g_tree_part[tree] = p; }
    p->t_created = std::chrono::steady_clock::now();
    if (env_true("HYPHY_B200_VERBOSE"))
        fprintf(stderr, "[hyphy_b200] partition %lu on device %d: %ld patterns x %ld states, %ld leaves, %ld internal nodes, %ld rate classes, %s pruning\n",
                index, device, p->S, p->D, p->L, p->I, p->C, hb2_pruning_kernel(p->h));
}

void *partition(void *state, unsigned long index) {
    State *st = static_cast<State *>(state);
    return (st && index < st->parts.size()) ? st->parts[index] : nullptr;
}

void destroy_all(void *&state) {
    State *st = static_cast<State *>(state);
    if (!st) return;
    for (Part *p : st->parts) {
        if (!p) continue;
        if (env_true("HYPHY_B200_VERBOSE"))
            fprintf(stderr, "[hyphy_b200] partition destroyed after %lu evaluations, %lu rate matrices exponentiated on the device, %lu host transition matrices, %lld kernel launches; "
                            "%lu of them handed over as formula values through %zu template(s); lifetime %.3f s of which %.3f s in matrix hand-over and %.3f s in hb2_evaluate (rest = HyPhy host code)\n",
                    p->n_eval, p->n_rate, p->n_trans, (long long)hb2_launch_count(p->h), p->n_compiled, p->templates.size(), seconds_since(p->t_created), p->t_handover, p->t_evaluate);
        if (g_current == p) g_current = nullptr;
        { std::lock_guard<std::mutex> lk(g_tree_mu); auto it = g_tree_part.find(p->tree); if (it != g_tree_part.end() && it->second == p) g_tree_part.erase(it); }
        hb2_destroy(p->h);
        delete p;
    }
    delete st;
    state = nullptr;
}

Scope::Scope(void *part) : part_(part), prev_(g_current) {
    g_current = static_cast<Part *>(part);
    if (g_current) g_current->epoch++;
}
Bypass::Bypass() { g_bypass++; }
Bypass::~Bypass() { g_bypass--; }
Scope::~Scope() { g_current = static_cast<Part *>(prev_); }

namespace {

_Matrix *placeholder(Part *p, _Matrix *existing, _Matrix const *not_this) {
    // exp(Qt) lives on the device only.  The node keeps a matrix object of the right shape so that the host's "has this
    // node ever been exponentiated" bookkeeping (NeedNewCategoryExponential, calcnode.cpp:480-523) works; its entries are
    // NaN so that any host code path that still tried to READ a transition matrix would fail loudly instead of computing
    // with stale numbers.
    if (existing && existing != not_this && (long)existing->GetHDim() == p->D && (long)existing->GetVDim() == p->D && existing->is_dense())
        return existing;
    _Matrix *ph = new _Matrix(p->D, p->D, false, true);
    for (long k = 0; k < p->D * p->D; k++) ph->theData[k] = NAN;
    return ph;
}

// equilibrium frequencies MultByFreqs would multiply the columns by (matrix.cpp:1550-1563), or nullptr
double const *model_frequencies(long model_index, long D) {
    if (model_index < 0 || model_index >= (long)modelFrequenciesIndices.lLength) return nullptr;
    long const fv = modelFrequenciesIndices.list_data[model_index];
    if (fv < 0) return nullptr;
    _Matrix *fm = (_Matrix *)LocateVar(fv)->GetValue();
    if (!fm) return nullptr;
    fm = (_Matrix *)fm->ComputeNumeric();
    if (!fm || !fm->is_dense() || (long)(fm->GetHDim() * fm->GetVDim()) != D) return nullptr;
    return fm->theData;
}

}  // namespace

_Matrix *compiled(_CalcNode *node, _Matrix *mm, long model_index, long catID, _Matrix *existing) {
    Part *p = g_current;
    static bool const dense_only = env_true("HYPHY_B200_DENSE");
    if (!p || !mm || dense_only) return nullptr;
    auto it = p->node_id.find(node);
    if (it == p->node_id.end()) return nullptr;
    _CompiledMatrixData *cmd = hb2_hooks_access::compiled(mm);
    if (!cmd || (long)mm->GetHDim() != p->D || (long)mm->GetVDim() != p->D) return nullptr;
    auto const t0 = std::chrono::steady_clock::now();
    long const stored = hb2_hooks_access::stored_count(mm), n_formulas = (long)cmd->formulasToEval.lLength;
    Part::Template &T = p->templates[mm];
    bool const wants_freqs = model_index >= 0 && model_index < (long)modelFrequenciesIndices.lLength && modelFrequenciesIndices.list_data[model_index] >= 0;
    if (T.id < 0 || T.cmd != cmd || T.stored != stored || T.n_formulas != n_formulas) {
        // (re)register the static half: which stored off-diagonal entry takes which formula
        if (T.id < 0) {
            if ((long)p->templates.size() > HB2_MAX_TEMPLATES) { p->templates.erase(mm); return nullptr; }   // dense route for the rest
            T.id = (int)p->templates.size() - 1;
        }
        long const *index = hb2_hooks_access::stored_index(mm);
        std::vector<int64_t> ei, ef;
        for (long e = 0; e < stored; e++) {
            long const flat = index ? index[e] : e, f = cmd->formulaRefs[e];
            if (flat < 0 || f < 0) continue;
            if (flat / p->D == flat % p->D) continue;                      // the diagonal is -(row sum) on the device too
            ei.push_back(flat); ef.push_back(f);
        }
        double const *fr = wants_freqs ? model_frequencies(model_index, p->D) : nullptr;
        if (ei.empty() || (wants_freqs && !fr)) { p->templates.erase(mm); return nullptr; }
        if (hb2_set_rate_template_id(p->h, T.id, (int64_t)ei.size(), ei.data(), ef.data(), n_formulas, fr)) fatal("hb2_set_rate_template_id failed");
        T.cmd = cmd; T.stored = stored; T.n_formulas = n_formulas;
        T.freq_var = wants_freqs ? modelFrequenciesIndices.list_data[model_index] : -1;
        T.freqs.assign(fr ? fr : nullptr, fr ? fr + p->D : nullptr);
        T.checked_epoch = p->epoch;
    } else if (T.freq_var >= 0 && T.checked_epoch != p->epoch) {
        // estimated equilibrium frequencies can move between evaluations: one comparison per ComputeBlock
        T.checked_epoch = p->epoch;
        double const *fr = model_frequencies(model_index, p->D);
        if (!fr) fatal("model frequencies became unavailable");
        if (memcmp(fr, T.freqs.data(), sizeof(double) * p->D) != 0) {
            if (hb2_set_template_frequencies(p->h, T.id, fr)) fatal("hb2_set_template_frequencies failed");
            T.freqs.assign(fr, fr + p->D);
        }
    }
    // this node's parameter values are already in the model's variables (RecomputeMatrix copied them): refresh the
    // compiled formulas' inputs and evaluate the unique formulas
    for (unsigned long i = 0; i < cmd->varIndex.lLength; i++) {
        _Variable *v = LocateVar(cmd->varIndex.list_data[i]);
        if (v->ObjectClass() == MATRIX) cmd->varValues[i].reference = (hyPointer)((_Matrix *)v->Compute())->theData;
        else cmd->varValues[i].value = v->IsIndependent() ? v->Value() : v->Compute()->Value();
    }
    for (long f = 0; f < n_formulas; f++)
        cmd->formulaValues[f] = ((_Formula *)cmd->formulasToEval.list_data[f])->ComputeSimple(cmd->theStack, cmd->varValues);
    int64_t const id = it->second;
    if (hb2_set_matrices_compiled_id(p->h, T.id, catID, 1, &id, cmd->formulaValues)) fatal("hb2_set_matrices_compiled_id failed");
    p->n_rate++; p->n_compiled++;
    p->t_handover += seconds_since(t0);
    return placeholder(p, existing, nullptr);
}

_Matrix *intercept(_CalcNode *node, _Matrix *m, long catID, bool do_exponentiation, _Matrix *existing) {
    Part *p = g_current;
    if (!p || !m || g_bypass) return nullptr;
    auto it = p->node_id.find(node);
    if (it == p->node_id.end()) return nullptr;                       // a node of some other tree
    if ((long)m->GetHDim() != p->D || (long)m->GetVDim() != p->D) return nullptr;
    int64_t const id = it->second;
    auto const t0 = std::chrono::steady_clock::now();
    double const *data = m->is_dense() ? m->theData : nullptr;
    if (!data) {                                                      // compressed-sparse rate matrix (codon models)
        thread_local std::vector<double> dense;
        dense.assign((size_t)p->D * p->D, 0.0);
        double *d = dense.data();
        m->ForEachCellNumeric([d](hyFloat v, long idx, long, long) -> void { d[idx] = v; });
        data = d;
    }
    {
        std::lock_guard<std::mutex> guard(p->lock);                   // one host thread per partition handle at a time
        if (hb2_set_matrices(p->h, catID, 1, &id, &data, do_exponentiation ? HB2_MATRIX_RATE : HB2_MATRIX_TRANS))
            fatal("hb2_set_matrices failed");
        if (do_exponentiation) p->n_rate++; else p->n_trans++;
        p->t_handover += seconds_since(t0);
    }
    if (!do_exponentiation) return m;                                 // host-computed P (explicit-form models): kept as is
    return placeholder(p, existing, m);
}

double compute_block(void *part, _TheTree *tree, long catID, _SimpleList const &branches, double *siteRes, long *scc,
                     long branchIndex, _SimpleList *branchValues) {
    Part *p = static_cast<Part *>(part);
    static int64_t const none = 0;
    int64_t const *upd = branches.lLength ? (int64_t const *)branches.list_data : &none;
    double lnl = 0.0;
    int rc;
    auto const t0 = std::chrono::steady_clock::now();
    if (branchIndex >= 0) {
        // reference convention (tree_evaluator.cpp:3624,173): internal index, or I + leaf index -> flat node id
        int64_t const forced = branchIndex < p->I ? branchIndex + p->L : branchIndex - p->I;
        rc = hb2_evaluate_forced(p->h, catID, (int64_t)branches.lLength, upd, tree->GetProbs(), forced,
                                 (int64_t const *)branchValues->list_data, &lnl, siteRes, (int64_t *)scc);
    } else {
        rc = hb2_evaluate(p->h, catID, (int64_t)branches.lLength, upd, tree->GetProbs(), &lnl, siteRes, (int64_t *)scc);
    }
    if (rc) fatal("evaluation failed");
    p->n_eval++;
    p->t_evaluate += seconds_since(t0);
    return lnl;
}

void materialize(void *part, _TheTree *tree, _DataSetFilter const *filter, double *inode_cache, _SimpleList const *site_ordering) {
    Part *p = static_cast<Part *>(part);
    if (!p) return;
    long const S = p->S, D = p->D, L = p->L, I = p->I, C = p->C;
    std::vector<double> buf((size_t)std::max(S * D, D * D));
    std::vector<int32_t> ex((size_t)S);
    for (long cat = 0; cat < C; cat++) {
        // transition matrices: every branch (all nodes but the root), into the matrix the node keeps for this class
        for (long n = 0; n < L + I - 1; n++) {
            _CalcNode *node = n < L ? (_CalcNode *)tree->flatCLeaves.GetItem(n) : (_CalcNode *)tree->flatTree.GetItem(n - L);
            _Matrix *m = node->GetCompExp(C > 1 ? cat : -1);
            if (!m || m->GetHDim() != D || m->GetVDim() != D || !m->theData) continue;      // nothing the host could read either
            if (hb2_read_transition(p->h, cat, n, buf.data())) fatal("hb2_read_transition failed");
            memcpy(m->theData, buf.data(), (size_t)D * D * sizeof(double));
        }
        // conditionals of the internal nodes, in the host's cache order (position s holds pattern site_ordering[s])
        if (inode_cache)
            for (long i = 0; i < I; i++) {
                if (hb2_read_conditionals(p->h, cat, i, buf.data(), ex.data())) fatal("hb2_read_conditionals failed");
                double *dst = inode_cache + ((size_t)cat * I + i) * S * D;
                for (long s = 0; s < S; s++) {
                    long const pat = site_ordering ? site_ordering->list_data[s] : s;
                    memcpy(dst + (size_t)s * D, buf.data() + (size_t)pat * D, (size_t)D * sizeof(double));
                }
            }
    }
    (void)filter;
}

void materialize_tree(_TheTree *tree) {
    Part *p = nullptr;
    { std::lock_guard<std::mutex> lk(g_tree_mu); auto it = g_tree_part.find(tree); if (it != g_tree_part.end()) p = it->second; }
    if (p) materialize(p, tree, nullptr, nullptr, nullptr);
}

}  // namespace hb2_hooks
