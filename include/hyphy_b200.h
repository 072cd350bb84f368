/*
 * hyphy_b200.h -- C ABI of the B200-native phylogenetic-likelihood engine (libhyphy_b200.so).
 *
 * This is the drop-in boundary behind the reference's `_LikelihoodFunction::ComputeBlock`
 * (veg/hyphy src/core/likefunc.cpp:10783-11289).  The reference has no plugin interface for its evaluator; the
 * only precedent is the dormant `#ifdef MDSOCL` trio (construct in SetupLFCaches likefunc.cpp:4182-4184, init
 * :4313-4316, destroy :10546-10551).  The entry points below are what a maintainer binds at those three places
 * (INTEGRATION.md shows the patch).  Plain pointers and sizes only; no C++/torch types cross this line.
 *
 * Conventions (all follow the reference so host arrays can be passed unmodified):
 *   - nodes: 0..L-1 leaves in post-order, L..L+I-1 internal nodes in post-order, root = L+I-1;
 *     flatParents[node] = parent's INTERNAL index (0..I-1), root = -1        (tree.cpp:722-766, tree.h:336)
 *   - leafState[l*S + s]: original pattern order; >=0 state index, <0 -> ambiguity row -(code+1)
 *                                                                            (likefunc.cpp:4265-4308)
 *   - matrices: D*D row-major doubles, row = parent(from) state, column = child(to) state
 *                                                                            (tree_evaluator.cpp:232,3674)
 *   - per-pattern results use the reference's scaler convention L_true = L * 2^(-64*count)
 *                                                                            (likefunc2.cpp:828-859,1484-1506)
 *   - errors: every call returns 0 on success, nonzero on failure (hb2_last_error() has the text).  There is NO
 *     CPU fallback: a missing GPU makes hb2_create fail and the host must treat that as fatal
 *     (HandleApplicationError, global_things.cpp:787).  Numerical failures are value-encoded exactly like
 *     tree_evaluator.cpp:4094-4148: lnL = -inf when a pattern has likelihood <= 0, NaN is propagated.
 *   - threading: one host thread per partition handle at a time (ComputeBlock is entered from the single
 *     interpreter thread; OpenMP lives inside it).  Calls block until the result is on the host.
 */
#ifndef HYPHY_B200_H
#define HYPHY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB2_ABI_VERSION 1

typedef struct hb2_partition hb2_partition;   /* opaque: one per (likelihood function, partition) */

/* hb2_create flags */
#define HB2_FLAG_DEFAULT        0
#define HB2_FLAG_FORCE_FP64     1   /* always use the fp64 pruning kernels (parity reference), never the tensor path */

/* matrix kinds for hb2_set_matrices */
#define HB2_MATRIX_RATE   0   /* numeric rate matrix Q*t after MultByFreqs: engine exponentiates on device
                                 (replaces _Matrix::Exponentiate matrix.cpp:5537 / ExponentiateMatrices tree.cpp:2932) */
#define HB2_MATRIX_TRANS  1   /* transition matrix P already exponentiated by the host (GetCompExp()->theData) */

int         hb2_abi_version(void);
const char *hb2_last_error(void);              /* thread-local text of the last failure */
int         hb2_device_count(void);            /* number of visible CUDA devices (0 => hb2_create will fail) */

/* Replaces the allocation half of _LikelihoodFunction::SetupLFCaches (likefunc.cpp:4163-4319) for one partition.
 * S patterns, D states, L leaves, I internal nodes, C rate classes (categoryCount, >=1).
 * ambig: nAmb*D doubles of 0/1 (dataset_filter.cpp:1594-1632); patternFreq: theFilter->theFrequencies.
 * All host arrays are copied; the caller keeps ownership.  device = CUDA ordinal. */
int hb2_create(hb2_partition **out, int64_t S, int64_t D, int64_t L, int64_t I, int64_t C,
               const int64_t *flatParents, const int64_t *leafState, const double *ambig, int64_t nAmb,
               const int64_t *patternFreq, int device, int flags);

/* Replaces ExponentiateMatrices (tree.cpp:2932-3113) + _CalcNode::SetCompExp (calcnode.cpp:714): hands the engine
 * the matrices of the nodes in `*matrices` (DetermineNodesForUpdate, tree.cpp:3117) for rate class `cat`
 * (0..C-1; -1 is accepted as 0 for partitions without category variables, likefunc.cpp:10843).
 * nodeIds[k] in 0..L+I-2; M[k] -> D*D doubles.  kind = HB2_MATRIX_RATE | HB2_MATRIX_TRANS. */
int hb2_set_matrices(hb2_partition *p, int64_t cat, int64_t n, const int64_t *nodeIds,
                     const double *const *M, int kind);
/* same, matrices packed back to back (n*D*D doubles) */
int hb2_set_matrices_packed(hb2_partition *p, int64_t cat, int64_t n, const int64_t *nodeIds,
                            const double *M, int kind);

/* Compact hand-over of rate matrices (removes the 8*D*D bytes per matrix of H2D traffic and lets the device do
 * what _Matrix::EvaluateSimple's scatter (matrix.cpp:3094-3321) and MultByFreqs (matrix.cpp:1546-1677) do on the host).
 * The reference compiles a model matrix into _CompiledMatrixData (include/matrix.h:69-80): a list of unique formulas
 * (formulasToEval), per-evaluation numeric formulaValues[], and formulaRefs[] mapping every stored off-diagonal entry
 * to its formula.  hb2_set_rate_template takes that static part once per model:
 *   entryIndex[e]   = row*D + col of stored entry e (theIndex), off-diagonal
 *   entryFormula[e] = formulaRefs[e] in [0, nFormulas)
 *   colFreq         = nullable D doubles: entry (r,c) is multiplied by colFreq[c] (MultByFreqs; NULL when the model
 *                     was declared with the "do not multiply by frequencies" flag)
 * hb2_set_matrices_compiled then hands over, per evaluation, only formulaValues (n * nFormulas doubles, one row per
 * listed node); the engine scatters them, applies colFreq, sets diagonal = -(row sum) and exponentiates. */
int hb2_set_rate_template(hb2_partition *p, int64_t nnz, const int64_t *entryIndex, const int64_t *entryFormula,
                          int64_t nFormulas, const double *colFreq);
int hb2_set_matrices_compiled(hb2_partition *p, int64_t cat, int64_t n, const int64_t *nodeIds,
                              const double *formulaValues);
/* Several models on one tree (foreground / background branches of RELAX and BUSTED, per-branch models of aBSREL; every
 * _CalcNode carries its own model index, calcnode.h): one template per model matrix, ids 0..HB2_MAX_TEMPLATES-1; the two
 * calls above are the id-0 forms.  hb2_set_template_frequencies replaces colFreq when the model's equilibrium
 * frequencies are themselves being estimated (cheap: D doubles, no reallocation). */
#define HB2_MAX_TEMPLATES 16
int hb2_set_rate_template_id(hb2_partition *p, int64_t templateId, int64_t nnz, const int64_t *entryIndex,
                             const int64_t *entryFormula, int64_t nFormulas, const double *colFreq);
int hb2_set_matrices_compiled_id(hb2_partition *p, int64_t templateId, int64_t cat, int64_t n, const int64_t *nodeIds,
                                 const double *formulaValues);
int hb2_set_template_frequencies(hb2_partition *p, int64_t templateId, const double *colFreq);

/* Explicit-form mixtures P = sum_k w_k Exp(Q_k) per branch (BS-REL; tree.cpp:3047-3089):
 * K components for each listed node, M packed [n][K][D*D], w packed [n][K].  Asynchronous like the other hand-overs
 * (two kernel launches on the partition's stream, no host synchronisation); M and w are copied before the call returns. */
int hb2_set_mixture_matrices(hb2_partition *p, int64_t cat, int64_t n, const int64_t *nodeIds, int64_t K,
                             const double *M, const double *w);

/* Replaces likefunc.cpp:10978-11123 (+ ComputeTreeBlockByBranch tree_evaluator.cpp:3556-4171) for ONE rate class:
 * prunes the nodes in updateNodes (nUpdate < 0 or updateNodes == NULL => all nodes) and reduces at the root.
 *   lnL            : sum_s f_s log L_s, already scale-corrected (what ComputeBlock returns at likefunc.cpp:11259)
 *   siteL/siteScale: nullable; per-pattern (original order) L and count with L_true = L*2^(-64*count)
 *                    (the storageVec / siteCorrectionCounts outputs, tree_evaluator.cpp:4080-4092, :84-87) */
int hb2_evaluate(hb2_partition *p, int64_t cat, int64_t nUpdate, const int64_t *updateNodes,
                 const double *rootFreqs, double *lnL, double *siteL, int64_t *siteScale);

/* hb2_evaluate with ONE node pinned to a per-pattern state -- the reference's `branchIndex` / `branchValues` arguments of
 * ComputeBlock (likefunc.cpp:10783), i.e. setBranch / setBranchTo of ComputeTreeBlockByBranch (tree_evaluator.cpp:3624:
 * a pinned leaf has its observed state replaced :173-181, a pinned internal node starts from the indicator vector of its
 * state :585-605, a pinned root contributes only that state :4059-4063).  forcedNode is a FLAT node id (0..L-1 leaves,
 * L..L+I-1 internal nodes; the reference's branchIndex b maps to b + L for b < I and to b - I otherwise), forcedStates[S]
 * in original pattern order, each in 0..D-1.  The caller lists the pinned node (a leaf) or its children (an internal
 * node) in updateNodes exactly as DetermineNodesForUpdate does with `addOne` (tree.cpp:3117,3262), and schedules the
 * recomputation of the affected conditionals afterwards (AddBranchToForcedRecomputeList, likefunc2.cpp:979-1041).
 * Everything PopulateConditionalProbabilities builds on per-class ComputeBlock calls -- ConstructCategoryMatrix run
 * modes, HMM / constant-on-partition categories, ancestral sampling -- only needs this entry point. */
int hb2_evaluate_forced(hb2_partition *p, int64_t cat, int64_t nUpdate, const int64_t *updateNodes,
                        const double *rootFreqs, int64_t forcedNode, const int64_t *forcedStates,
                        double *lnL, double *siteL, int64_t *siteScale);

/* Fused replacement for the whole category loop PopulateConditionalProbabilities(WeightedSum) +
 * SumUpSiteLikelihoods (likefunc2.cpp:484-908, 1446-1506): all C classes pruned in one pass,
 * L_s = sum_c weights[c]*L_{c,s} combined on device, one lnL comes back.  Same outputs as hb2_evaluate. */
int hb2_evaluate_classes(hb2_partition *p, const double *weights, int64_t nUpdate, const int64_t *updateNodes,
                         const double *rootFreqs, double *lnL, double *siteL, int64_t *siteScale);

/* Single-branch shortcut (ComputeBranchCache tree_evaluator.cpp:4286-4845 and ComputeLLWithBranchCache
 * tree.cpp:3383-3934, entered from ComputeBlock likefunc.cpp:10984-10993 / :11170-11177).  After a regular evaluation,
 * hb2_branch_cache_build(node) collects, for every owned class and pattern, the likelihood of everything outside the
 * subtree of `node` as seen from the parent end of its branch (the reference's branchCaches[.][1]; the subtree side is
 * the resident conditional).  While only THAT branch's matrix changes (set it with hb2_set_matrices*),
 * hb2_branch_cache_evaluate returns lnL in O(S*D^2) without touching the tree; cat/weights as in hb2_evaluate /
 * hb2_evaluate_classes (weights != NULL selects the all-class form).  Any regular evaluation invalidates the cache; a
 * pending matrix of another node makes the call fail.  Unlike the reference this needs no reversible model: the outside
 * vectors are propagated through transposed transition matrices instead of re-rooting the tree. */
int hb2_branch_cache_build(hb2_partition *p, int64_t node, const double *rootFreqs);
int hb2_branch_cache_evaluate(hb2_partition *p, int64_t cat, const double *weights, double *lnL, double *siteL,
                              int64_t *siteScale);

/* Batched one-pattern likelihoods on the partition's tree.  The site phases of FEL and MEME
 * (res/TemplateBatchFiles/SelectionAnalyses/FEL.bf:1180-1233, MEME.bf:699-746) fit thousands of independent likelihood
 * functions that all use ONE site pattern of the alignment, the same tree, and differ only in a few per-site parameters
 * (alpha, beta, ...); the reference farms them out one by one (mpi.QueueJob).  Here `nSets` such problems are evaluated in
 * one call: set s prunes pattern patternOf[s] of the partition with ITS OWN matrices for every branch, handed over in the
 * compiled form of template `templateId`: formulaValues[s][b][nFormulas], b = 0..B-1 (one rate class per set).
 * branchGroup (nullable, B entries in 0..7) tells the engine which branches of a set share one rate matrix up to a scalar
 * (FEL: tested vs background branches); it is a hint for the shared-powers exponential, proportionality is still checked.
 * siteLnL[s] = log-likelihood of the pattern under set s (NOT multiplied by the pattern's frequency).  fp64 throughout;
 * the partition's resident caches are not touched.  33..64 states only. */
int hb2_batch_site_likelihoods(hb2_partition *p, int64_t templateId, int64_t nSets, const int64_t *patternOf,
                               const double *formulaValues, const int64_t *branchGroup, const double *rootFreqs,
                               double *siteLnL);

/* FillInConditionals-style read-back (tree.cpp:3335): conditionals of internal node `inode` (0..I-1), class cat,
 * as S*D doubles (original pattern order) and S binary exponents: true value = cond * 2^exp2. */
int hb2_read_conditionals(hb2_partition *p, int64_t cat, int64_t inode, double *cond, int32_t *exp2);

/* Read back the transition matrix currently cached for (cat, node): D*D doubles, row = parent state. */
int hb2_read_transition(hb2_partition *p, int64_t cat, int64_t node, double *P);

/* Multi-GPU (one process per GPU, patterns sharded by the caller: each rank creates its partition with its own
 * pattern shard).  After hb2_comm_init every hb2_evaluate* returns the SUM over ranks of the partial lnL
 * (a single fp64 ncclAllReduce on the partition's stream; per-pattern outputs stay local to the shard).
 * uniqueId: the 128 bytes of an ncclUniqueId produced by hb2_comm_unique_id on rank 0 and broadcast by the host. */
int hb2_comm_unique_id(void *uniqueId128);
int hb2_comm_init(hb2_partition *p, int nRanks, int rank, const void *uniqueId128);

/* Optional second sharding axis for partitions with rate classes (the per-class loop of
 * ComputeSiteLikelihoodsForABlock, likefunc2.cpp:912, iterates independent ComputeBlock calls): the nRanks ranks are
 * arranged as (nRanks/nGroups pattern shards) x (nGroups class groups); rank r belongs to class group r % nGroups and
 * pattern shard r / nGroups, i.e. the nGroups consecutive ranks of one shard are created with the SAME pattern slice.
 * Group g owns classes [g*C/nGroups, (g+1)*C/nGroups): matrices handed over for other classes are ignored on this
 * rank (no expm, no pruning), hb2_evaluate_classes prunes the owned classes only, and the per-pattern class partials
 * (value, binary exponent) of all ranks are exchanged before the logarithm with ONE collective per evaluation: every
 * rank writes its partials into its own slot of a buffer that is zero elsewhere and a sum ncclAllReduce acts as the
 * gather; every rank then merges every shard and holds the complete, bit-identical lnL (no second collective).
 * Requires C % nGroups == 0 and nRanks % nGroups == 0; hb2_evaluate (single class) is refused in this mode.  Call after
 * hb2_comm_init, before the first matrix is set. */
int hb2_comm_class_groups(hb2_partition *p, int nGroups);
/* Per-pattern outputs of ALL pattern shards on every rank ("per-pattern outputs, when requested, are gathered";
 * what the host needs for ConstructCategoryMatrix / SITE_LOG_LIKELIHOODS on a sharded partition, likefunc.cpp:1791).
 * siteL / siteScale: this rank's S values as returned by the last hb2_evaluate*; allSiteL / allSiteScale: room for
 * `capacity` patterns, filled in shard order (= pattern order of the unsharded partition when shards are contiguous
 * slices); *total = patterns over all shards.  Collective over the partition's communicator (two ncclAllGather). */
int hb2_comm_gather_sites(hb2_partition *p, const double *siteL, const int64_t *siteScale, double *allSiteL,
                          int64_t *allSiteScale, int64_t capacity, int64_t *total);

/* Pair of SetupLFCaches in DeleteCaches (likefunc.cpp:10556-10601). */
void hb2_destroy(hb2_partition *p);

/* Host-side planner of the tcgen05 pruning pass, callable WITHOUT a GPU (tests/test_planner.py): for a tree (same
 * flatParents as hb2_create), a dirty set given as hb2_evaluate's updateNodes (NULL or nUpdate < 0 = whole tree) and a
 * number of lanes (1..15), returns the lanes' step lists exactly as the kernel receives them: laneStart[lanes+1] and
 * 2 ints per step -- child id | HB2_STEP_WAIT (1<<30: produced by another lane) | HB2_STEP_CHAIN (1<<29: produced by
 * the lane's previous job, taken from registers) | HB2_STEP_MUL (1<<28: side product, multiply without a matrix); job's
 * node slot (>= I: side product of node slot-I) | HB2_STEP_FIRST (1<<28) | HB2_STEP_LAST (1<<29).  `steps` must hold
 * 2*(L+2I) ints.  splitNodes: bit 0 = side products on; bit 1 = the fp64 lanes kernel's rule for nodes with more than
 * two children (never split or chained, children in tree order), which makes conditionals independent of the plan. */
int hb2_plan_walk(int64_t L, int64_t I, const int64_t *flatParents, int64_t nUpdate, const int64_t *updateNodes, int lanes,
                  int splitNodes, int32_t *laneStart, int32_t *steps, int64_t stepCapacity, int64_t *nSteps);

/* Introspection for tests/bench: kernel launches issued so far by this partition, and device pointers/timing. */
int64_t hb2_launch_count(const hb2_partition *p);
/* 0: fp64 pruning kernels (4/20-state register kernels, or HB2_FLAG_FORCE_FP64);
 * 1: tcgen05 tensor-core pruning, error-compensated 3xTF32 split with fp32 conditionals (33..64 states). */
int hb2_precision_mode(const hb2_partition *p);
/* Name of the pruning kernel this partition launches, and the kernel launches per evaluation of the three stages
 * {expm, pruning, root reduction} as counted during the last hb2_time_resident (bench.py's roofline object). */
const char *hb2_pruning_kernel(const hb2_partition *p);
int hb2_stage_launches(const hb2_partition *p, int64_t *out3);
/* Runs the same work as hb2_evaluate_classes `iters` times with inputs already resident on the device
 * (re-exponentiating every cached rate matrix and re-pruning the whole tree each time) and returns the mean
 * device time per evaluation in milliseconds (CUDA events on the partition's stream).  stageMs (nullable, 3 doubles)
 * receives the mean per-stage times {expm, pruning, root reduction}. */
int hb2_time_resident(hb2_partition *p, const double *weights, const double *rootFreqs, int iters,
                      double *msPerEval, double *stageMs, double *lnL);

#ifdef __cplusplus
}
#endif
#endif /* HYPHY_B200_H */
