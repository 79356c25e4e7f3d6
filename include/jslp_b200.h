/*
 * include/jslp_b200.h -- C ABI of libjslp_b200.so, the B200-native (sm_100a) drop-in for
 * jsLPSolver's dense-tableau simplex + branch-and-cut hot path.
 *
 * The reference (JWally/jsLPSolver, TypeScript) has no FFI of its own: its plug-in seam is the
 * `Tableau` method table plus the injected `BranchAndCutService`.  Every entry
 * point below replaces exactly one member of that seam; an N-API addon binds
 * them 1:1 from a `GpuTableau extends Tableau` subclass.  Plain pointers and sizes only.
 *
 * Conventions (mirroring the reference, SURVEY.md 8b "Conventions"):
 *   - every function returns 0 on success, a negative JSLP_E_* on misuse/driver failure;
 *     solver outcomes (infeasible / unbounded / cycle) are FLAGS in the status structs, never
 *     error codes (the reference never throws for them);
 *   - not thread-safe per handle, thread-safe across handles; one context per GPU;
 *   - all matrices are row-major fp64, stride == width, row 0 = cost row, column 0 = RHS
 *     (tableau.ts:49-54,304); index maps are int32 with -1 = "not there" (tableau.ts:306-316);
 *   - there is NO CPU fallback: without a CUDA device every call fails with JSLP_E_CUDA.
 */
#ifndef JSLP_B200_H
#define JSLP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JSLP_OK 0
#define JSLP_E_INVALID (-1)   /* bad argument / bad handle state            */
#define JSLP_E_CUDA (-2)      /* CUDA driver/runtime failure (see last_error) */
#define JSLP_E_CAPACITY (-3)  /* row capacity / shared-memory capacity exceeded */
#define JSLP_E_UNSUPPORTED (-4) /* feature outside the hot-path scope (e.g. useMIRCuts) */

typedef struct jslp_ctx jslp_ctx;
typedef struct jslp_tab jslp_tab;

/* Thread-local description of the last failure (never NULL). */
const char *jslp_last_error(void);
/* ABI version, bumped on any signature change. */
int jslp_abi_version(void);

/* One context per GPU.  `stream` is a cudaStream_t (as void*) the library launches on, or NULL
 * for a private non-blocking stream.  Replaces: nothing (the reference has no device). */
int jslp_ctx_create(int device, void *stream, jslp_ctx **out);
void jslp_ctx_destroy(jslp_ctx *ctx);
/* The stream all work of this context is enqueued on (cudaStream_t as void*). */
void *jslp_ctx_stream(jslp_ctx *ctx);
/* Number of kernel launches issued by this context so far (bench.py's gpu_launches). */
int64_t jslp_ctx_launches(jslp_ctx *ctx);
int jslp_ctx_sync(jslp_ctx *ctx);

/* Replaces Tableau.initialize (tableau.ts:292-317): allocates a device tableau of `width` x
 * `height` with room for `row_capacity` >= height rows (branching cuts append rows,
 * cutting-strategies.ts:24-34).  `precision` is Tableau.precision (default 1e-8, tableau.ts:96). */
int jslp_tab_create(jslp_ctx *ctx, int width, int height, int row_capacity, double precision,
                    jslp_tab **out);
void jslp_tab_destroy(jslp_tab *tab);

/* Replaces the result of Tableau._resetMatrix / setModel (tableau.ts:319-391): uploads the
 * initial tableau built by the (unchanged) host front end.
 *   matrix            height*width doubles, row-major, stride == width
 *   var_index_by_row  height int32 (entry 0 == -1)          (tableau.ts:63)
 *   var_index_by_col  width  int32 (entry 0 == -1)          (tableau.ts:64)
 *   unrestricted      n_index bytes, 1 = Tableau.unrestrictedVars[index] === true, or NULL
 *   n_index           width + height - 2 (= Tableau.nVars)
 *   int_var_indices   model.integerVariables[].index in model order (mip-utils.ts:47,108), or NULL
 *   n_opt, opt_obj    optionalObjectives sorted by priority; n_opt*width reducedCosts
 *                     (tableau.ts:278-290), or 0/NULL
 * All host pointers may be pageable or pinned; the copy is synchronous w.r.t. the host.  */
int jslp_tab_upload(jslp_tab *tab, const double *matrix, const int32_t *var_index_by_row,
                    const int32_t *var_index_by_col, const uint8_t *unrestricted, int n_index,
                    const int32_t *int_var_indices, int n_int, int n_opt, const double *opt_obj);

/* Solver options that the reference keeps on Model / Tableau. */
enum {
    JSLP_OPT_ENGINE = 1,       /* 0 = auto (resident when H*W <= 16384, else fused), 1 = two-kernel
                                  (select + update), 2 = fused step, 4 = single-CTA resident    */
    JSLP_OPT_BATCH = 2,        /* pivots enqueued per host poll (default 256)              */
    JSLP_OPT_PIVOT_LOG_CAP = 3, /* keep a host-side (row,col,leaving,entering) log, 0 = off */
    /* tuning / diagnostics of the fused pivot step (no effect on results) */
    JSLP_OPT_STEP_VARIANT = 4, /* kernel instantiation: threads, occupancy, rows per pass, prefetch; -1 = auto */
    JSLP_OPT_GRID_PER_SM = 5,  /* CTAs per SM of the fused step, 0 = the variant's default        */
    JSLP_OPT_LOOKAHEAD = 6,    /* 1 (default) = look-ahead ratio test, 0 = generic serial tail    */
    JSLP_OPT_TIMELINE = 7,     /* record a per-CTA timeline for the first N launches of a solve   */
    JSLP_OPT_PDL = 8,          /* 1 = chain the fused steps with programmatic dependent launch    */
    JSLP_OPT_PINGPONG = 9,     /* 1 (default) = ping-pong tableau + two selector CTAs, 0 = in place */
    /* branch-and-cut nodes too large for shared memory (no effect on results) */
    JSLP_OPT_NODE_SLOTS = 10,  /* node LPs in flight side by side in HBM: -1 = auto (default), 0 = one at a
                                  time (the reference's literal applyCuts sequence), n = at most n slots */
    JSLP_OPT_SLOT_STEPS = 11,  /* pivots per slot between host polls of the slot batch (default 32)   */
    JSLP_OPT_SLOT_VARIANT = 12, /* kernel instantiation used by the slot batch (JSLP_OPT_STEP_VARIANT values) */
    JSLP_OPT_NODE_LOG_CAP = 14, /* pivot-log entries per node LP of the shared-memory node kernel (default 512); a node
                                   that needs more is re-evaluated on the HBM path (tests lower it to hit that path) */
    JSLP_OPT_USE_MIR_CUTS = 13  /* model.useMIRCuts (model.ts:69,313): applyCuts / branchAndCut run the MIR loop of
                                   branch-and-cut.ts:38-51 after every node's simplex()                          */
};
int jslp_tab_set_option(jslp_tab *tab, int key, double value);
/* Diagnostics: per-CTA timeline (8 int64 per CTA per launch) recorded under JSLP_OPT_TIMELINE. */
int jslp_debug_timeline(jslp_tab *tab, int64_t *out, int64_t cap_values, int *launches, int *grid);
/* Diagnostics: bandwidth [GB/s, read + write] of the library's own 128-bit copy loop ping-ponging between two
 * buffers of `bytes` each, `iters` round trips -- with 2 * bytes below the L2 size this is the L2-resident roof
 * the in-solve streaming phase runs under, above it the HBM roof (bench.py reports both beside roofline.frac). */
int jslp_debug_copy_gbs(jslp_ctx *ctx, int64_t bytes, int iters, double *gbs);

/* Tableau state read by callers after simplex() (SURVEY.md 8b "State contract"). */
typedef struct {
    int32_t feasible;            /* Tableau.feasible                                     */
    int32_t bounded;             /* Tableau.bounded                                      */
    int32_t cycled;              /* 0, or the phase (1/2) in which checkForCycles fired  */
    int32_t cycle_start;         /* model.messages "Start :"                             */
    int32_t cycle_length;        /* model.messages "Length :"                            */
    int32_t phase1_pivots;       /* return value of Tableau.phase1()                     */
    int32_t phase2_pivots;       /* return value of Tableau.phase2()                     */
    int32_t unbounded_var_index; /* Tableau.unboundedVarIndex, -1 = null                 */
    int32_t simplex_iters;       /* Tableau.simplexIters                                 */
    int32_t width, height;       /* current logical size                                 */
    int32_t engine;              /* engine that actually ran (JSLP_OPT_ENGINE values)    */
    double evaluation_raw;       /* matrix[0] at exit                                    */
    double evaluation;           /* Tableau.evaluation (setEvaluation rounding, tableau.ts:420-430) */
    double best_possible_eval;   /* Tableau.bestPossibleEval                             */
    double gpu_ms;               /* CUDA-event time of the solve on the context stream   */
    int64_t kernel_launches;     /* launches issued for this call                        */
} jslp_lp_status;

/* == Tableau.simplex() (tableau.ts:103-111 -> simplex.ts:14-23): phase1, then phase2 if
 * feasible.  check_cycles == model.checkForCycles (model.ts:73,359-363).                 */
int jslp_simplex(jslp_tab *tab, int check_cycles, jslp_lp_status *out);
/* == Tableau.phase1() / phase2() (simplex.ts:25-98 / 100-325). */
int jslp_phase1(jslp_tab *tab, int check_cycles, jslp_lp_status *out);
int jslp_phase2(jslp_tab *tab, int check_cycles, jslp_lp_status *out);
/* == Tableau.pivot(row, col) (simplex.ts:330-413). */
int jslp_pivot(jslp_tab *tab, int row, int col);

/* == Tableau.save() / restore() (backup.ts:49-105): device-side snapshot, D2D restore. */
int jslp_save(jslp_tab *tab);
int jslp_restore(jslp_tab *tab);

/* BranchCut (types.ts:17-21). type: 0 = "min" (x >= value), 1 = "max" (x <= value). */
typedef struct {
    int32_t type;
    int32_t var_index;
    double value;
} jslp_cut;

/* == Tableau.addCutConstraints(cuts) (cutting-strategies.ts:16-72). */
int jslp_add_cuts(jslp_tab *tab, const jslp_cut *cuts, int n);
/* == Tableau.addLowerBoundMIRCut(row) (upper_bound = 0) / addUpperBoundMIRCut(row) (1), cutting-strategies.ts:
 * 74-196: *added = 1 when a cut row (and its slack) was appended.  == Tableau.applyMIRCuts() (198-212): lower-
 * bound cuts on the first (at most 10) eligible rows.  == Tableau.computeFractionalVolume (mip-utils.ts:67-98). */
int jslp_add_mir_cut(jslp_tab *tab, int row, int upper_bound, int *added);
int jslp_apply_mir_cuts(jslp_tab *tab, int *n_added);
int jslp_fractional_volume(jslp_tab *tab, int ignore_integer_values, double *volume);
/* == BranchAndCutService.applyCuts (branch-and-cut.ts:33-52): restore, add cuts, simplex [, MIR loop]. */
int jslp_apply_cuts(jslp_tab *tab, const jslp_cut *cuts, int n, int check_cycles, jslp_lp_status *out);
/* == Tableau.isIntegral() (mip-utils.ts:43-61) and getMostFractionalVar() (mip-utils.ts:100-126):
 * *var_index = -1 when no fractional integer variable exists.                             */
int jslp_is_integral(jslp_tab *tab, int *is_integral);
int jslp_most_fractional(jslp_tab *tab, int32_t *var_index, double *value);

/* The dynamic-modification API on the device-resident tableau (dynamic-modification.ts:16-55,78-316): edits after a
 * solve without rebuilding / re-uploading it.  Constraints and variables are named by their element index
 * (Constraint.index == its slack's index, Variable.index); availableIndexes and the Constraint / Variable objects stay
 * host bookkeeping.  opt_slot: -1 = priority 0 (cost row), k = the k-th uploaded optional objective.
 *   put_in_base / take_out_of_base   == Tableau.putInBase / takeOutOfBase (may pivot); *row / *col = where it ended
 *   update_rhs                       == updateRightHandSide(constraint, difference)
 *   update_coefficient               == updateConstraintCoefficient(constraint, variable, difference)
 *   update_cost                      == updateCost(variable, difference)
 *   add_constraint / remove_constraint == addConstraint(constraint) / removeConstraint(constraint); terms in order
 *   add_variable / remove_variable   == addVariable(variable) / removeVariable(variable); cost_entry is the signed
 *                                       cost-row entry (isMinimization ? -cost : cost)                            */
int jslp_put_in_base(jslp_tab *tab, int var_index, int *row);
int jslp_take_out_of_base(jslp_tab *tab, int var_index, int *col);
int jslp_update_rhs(jslp_tab *tab, int constraint_index, double difference);
int jslp_update_coefficient(jslp_tab *tab, int constraint_index, int var_index, double difference);
int jslp_update_cost(jslp_tab *tab, int var_index, int opt_slot, double difference);
int jslp_add_constraint(jslp_tab *tab, int is_upper_bound, double rhs, int slack_index, const int32_t *term_var,
                        const double *term_coef, int n_terms);
int jslp_remove_constraint(jslp_tab *tab, int slack_index);
int jslp_add_variable(jslp_tab *tab, int var_index, double cost_entry, int opt_slot, int is_integer, int is_unrestricted);
/* Tableau's scalar bookkeeping for the host mirror: {width, height, nVars, lastElementIndex, row stride, row capacity} */
int jslp_tab_info(jslp_tab *tab, int32_t *out6);
int jslp_remove_variable(jslp_tab *tab, int var_index);

/* Read-back for updateVariableValues / generateSolutionSet / getSolution
 * (dynamic-modification.ts:57-76, solution.ts:35-60).  Any pointer may be NULL.
 *   matrix     height*width doubles (stride == width)   rhs_col   height doubles
 *   cost_row   width doubles                             opt_obj   n_opt*width doubles   */
int jslp_download(jslp_tab *tab, double *matrix, double *rhs_col, double *cost_row,
                  int32_t *var_index_by_row, int32_t *var_index_by_col, double *opt_obj,
                  int32_t *width, int32_t *height);
/* Drains the host-side pivot log: 4 int32 per pivot (row, col, leaving var, entering var). */
int jslp_pivot_log(jslp_tab *tab, int32_t *entries, int cap, int *n);

/* Multi-GPU communicator of the branch-and-cut frontier (one process per GPU).  NCCL is loaded at run time;
 * without libnccl these calls fail with JSLP_E_UNSUPPORTED and single-GPU use is unaffected.  Rank 0 obtains a
 * 128-byte unique id and hands it to the other ranks through whatever bootstrap the host has (the N-API host:
 * its own IPC; the Python mirror: torch.distributed); every rank then creates its communicator.  Replaces
 * nothing in the reference (it is single-process); carries the two collectives of SURVEY.md 8e.  */
typedef struct jslp_comm jslp_comm;
int jslp_comm_unique_id(uint8_t *id128);
int jslp_comm_create(jslp_ctx *ctx, const uint8_t *id128, int rank, int n_ranks, jslp_comm **out);
void jslp_comm_destroy(jslp_comm *comm);
/* The collectives themselves (host buffers in and out, staged through device memory on the context stream):
 * in-place rank-major all-gather of bytes_per_rank bytes, element-wise all-reduce(min) of n doubles.       */
int jslp_comm_all_gather(jslp_comm *comm, void *buf, int64_t bytes_per_rank);
int jslp_comm_all_reduce_min(jslp_comm *comm, double *vals, int n);

/* == BranchAndCutService.branchAndCut (branch-and-cut.ts:54-199). */
typedef struct {
    double tolerance;        /* model.tolerance                                            */
    int32_t is_minimization; /* model.isMinimization                                       */
    int32_t check_cycles;    /* model.checkForCycles                                       */
    int32_t max_spec_batch;  /* nodes evaluated speculatively per round (1 = reference order,
                                one node at a time); results are committed in exact pop order */
    int32_t rank, n_ranks;   /* node sharding across GPUs; 0/1 = single GPU                */
    int64_t max_nodes;       /* safety cap on evaluated nodes, 0 = none                    */
    /* Multi-GPU exchange hook: all-gather `bytes` per rank (in-place, rank-major).  The host
     * layer implements it with torch.distributed/NCCL.  NULL when n_ranks <= 1.            */
    int (*all_gather)(void *user, void *buf, int64_t bytes_per_rank);
    void *user;
    jslp_comm *comm;        /* NCCL communicator: when set it carries the all-gather of node summaries and the
                               all-reduce(min) of the incumbent bound, and the hook above is ignored          */
    int32_t shard_policy;   /* 0 = auto: shard a round over the ranks only when its node LPs run in HBM (nodes
                               that fit shared memory cost less than a collective: every rank evaluates them
                               itself); 1 = shard every round                                                */
    int32_t keep_solutions; /* model.keep_solutions (branch-and-cut.ts:143-153): every incumbent is stored   */
    double timeout_ms;      /* model.timeout (branch-and-cut.ts:61-63,76), wall clock, 0 = none              */
    /* Which BranchAndCutService main.ts:62-83 would have injected: 0 = createBranchAndCutService (default;
     * speculative frontier, node slots, multi-GPU), 1 = createEnhancedBranchAndCutService (options.nodeSelection /
     * options.branching; pseudocosts make it sequential: one node LP at a time, one GPU), 2 =
     * createIncrementalBranchAndCutService (options.useIncremental: the enhanced loop with parent checkpoints).  */
    int32_t service;
    int32_t node_selection;    /* enhanced: 1 = "best-first", 2 = "depth-first", 3 = "hybrid" (0 = its default, hybrid) */
    int32_t branching;         /* enhanced: 1 = "most-fractional", 2 = "pseudocost", 3 = "strong" (0 = pseudocost)      */
    int32_t strong_candidates; /* enhanced: strongBranchingCandidates (0 = 5)                                          */
} jslp_bnb_opts;

typedef struct {
    int32_t feasible, bounded, is_integral; /* Tableau.feasible/bounded/__isIntegral        */
    int32_t iterations;                     /* Tableau.branchAndCutIterations               */
    int32_t n_best_cuts;                    /* cuts of the winning branch (left appended)   */
    int32_t rounds;                         /* speculative rounds                           */
    int64_t nodes_evaluated;                /* node LPs solved, including discarded speculation */
    int64_t pivots;                         /* pivots over all node LPs                     */
    double evaluation;                      /* Tableau.evaluation of the final tableau      */
    double best_possible_eval;
    double gpu_ms;
    int64_t kernel_launches;
    double host_eval_ms;   /* wall time spent evaluating node batches (launch + read-back + sync) */
    double host_commit_ms; /* wall time spent in the sequential commit loop                     */
    double host_root_ms;   /* part of host_eval_ms spent on the root relaxation                 */
    double host_final_ms;  /* wall time of the final re-solve of the winning branch             */
    double node_kernel_ms; /* sum over rounds of the slowest node CTA's lifetime (%globaltimer) */
    int32_t timed_out;     /* the loop ended because Date.now() >= terminalTime (branch-and-cut.ts:76)   */
    int32_t n_solutions;   /* incumbents stored under keep_solutions (jslp_bnb_solution)                 */
    int64_t nodes_pruned;  /* speculative node LPs dropped or aborted by the incumbent bound of the round */
    int64_t collectives;   /* all-gathers + all-reduces issued (n_ranks > 1)                             */
    int64_t slot_pivots;   /* pivots executed in HBM node slots (jslp_slots.cuh), this rank             */
    double slot_ms;        /* wall time of the slot-batch graphs, this rank                              */
    double slot_bytes;     /* algorithmic bytes of those pivots: 16 * H_node * stride each               */
} jslp_bnb_status;

int jslp_branch_and_cut(jslp_tab *root, const jslp_bnb_opts *opts, jslp_bnb_status *out,
                        jslp_cut *best_cuts, int best_cuts_cap);
/* Per-node trace of the last branch_and_cut: 8 doubles per committed node
 * (iteration, nCuts, feasible, evaluation, integral(-1/0/1), branchVar, branchValue, pivots). */
int jslp_bnb_node_log(jslp_tab *root, double *entries, int64_t cap, int64_t *n);
/* model.solutions under keep_solutions (branch-and-cut.ts:143-153): incumbent i of the last branch_and_cut as
 * the state generateSolutionSet reads (solution.ts:35-60): Tableau.evaluation, height, varIndexByRow[height],
 * right-hand-side column[height].  `cap` = entries the two arrays can hold.                                */
int jslp_bnb_solution(jslp_tab *root, int i, double *evaluation, int32_t *height, int32_t *var_index_by_row,
                      double *rhs_col, int cap);

#ifdef __cplusplus
}
#endif
#endif /* JSLP_B200_H */
