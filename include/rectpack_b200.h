/* rectpack_b200 — C ABI of the B200-native rollout engine for perfect rectangle packing.
 *
 * This is the drop-in boundary for the hot path of igorpejic/visapi (paths below are
 * relative to the reference tree).  The reference has exactly one FFI on this path,
 * the f2py module `search` (engine/search.f90:1-16, bound at
 * engine/solution_checker.py:10); everything else on the path is Python calling
 * Python.  The entry points here are what a binding for the path binds: the scan
 * itself (bp_find_first) and batched forms of the functions that call it.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are HOST pointers unless a name
 *     ends in _dev.  The caller owns every host buffer; nothing is retained after a
 *     call returns.  Device memory lives behind the opaque handles.
 *   - every function returns 0 on success or a negative bp_status; bp_last_error()
 *     gives the message of the calling thread's last failure.
 *   - one bp_context per (device, host thread); calls on a context are serialised by
 *     the caller.  Kernels run on the stream set with bp_set_stream (default: the
 *     legacy default stream).
 *   - board: uint32 words, row-major, wpr = ceil(cols / 32) words per row; bit (c % 32)
 *     of word (c / 32) set  <=>  cell (row, c) occupied.  "Occupied" is whatever the
 *     caller packed: non-zero cells (canonical, binpack/MCTS.py:42) or cells == 1
 *     (solution_checker.py:117-119).
 *   - tile table: uint8 (h, w) pairs in list order, both orientations present;
 *     (0, 0) = absent entry (the [0, 0] pads of alpha/binpack/BinPackLogic.py:59-63).
 *   - limits: rows, cols <= 128; entries <= 128 (64 tiles).
 */
#ifndef RECTPACK_B200_H
#define RECTPACK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BP_MAX_ROWS 128
#define BP_MAX_COLS 128
#define BP_MAX_ENTRIES 128

typedef enum {
    BP_OK = 0,
    BP_ERR_INVALID = -1,     /* bad argument (null pointer, size out of range) */
    BP_ERR_CUDA = -2,        /* CUDA runtime failure; see bp_last_error() */
    BP_ERR_NO_DEVICE = -3,   /* no usable CUDA device: there is no CPU fallback */
    BP_ERR_STATE = -4        /* call not valid in the handle's current state */
} bp_status;

/* per-state verdicts; replace the string sentinels of solution_checker.py:16-18 */
typedef enum {
    BP_PLACED = 1,                         /* (True, board)                              */
    BP_ALL_TILES_USED = 2,                 /* 'ALL_TILES_USED'                           */
    BP_TILE_CANNOT_BE_PLACED = 3,          /* 'TILE_CANNOT_BE_PLACED' / (False, None)    */
    BP_NO_NEXT_POSITION_TILES_UNUSED = 4,  /* 'NO_NEXT_POSITION_TILES_UNUSED'            */
    BP_PAIR_MISSING = 5                    /* ValueError of eliminate_pair_tiles :325-329 */
} bp_verdict;

typedef enum { BP_MAX_DEPTH = 0, BP_AVG_DEPTH = 1 } bp_strategy;   /* binpack/MCTS.py:144-147 */

typedef struct bp_context bp_context;
typedef struct bp_search bp_search;
typedef struct bp_tree bp_tree;

/* ---- lifetime ------------------------------------------------------------------- */
int bp_create(int device_id, bp_context **out);
int bp_destroy(bp_context *ctx);
const char *bp_last_error(void);
/* cuda_stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream), 0 = default */
int bp_set_stream(bp_context *ctx, void *cuda_stream);
int bp_synchronize(bp_context *ctx);
/* sm_count, max SM clock (kHz), kernels launched by this context so far */
int bp_device_info(bp_context *ctx, int *sm_count, int *sm_clock_khz, int64_t *kernel_launches);

/* Issue-rate micro-benchmark (measurement aid): G warp-instructions/s of a
 * dependency-free LOP3/IADD3 stream (ALU pipe only) and of a LOP3/IMAD stream (ALU + FMA
 * pipes) on every SM.  The roofline peak of bench.py is SMs x 4 x clock; these two numbers
 * show how much of it integer code can reach. */
int bp_issue_microbench(bp_context *ctx, double *alu_only_ginst, double *alu_fma_ginst);

/* ---- a1: search.find_first (engine/search.f90:1-16) ----------------------------- */
/* first 0-based index k with haystack[k] == needle, else -1 */
int bp_find_first(bp_context *ctx, int32_t needle, const int32_t *haystack, int64_t n,
                  int64_t *out_index);

/* ---- a2-a8, a13: batched transitions on n states of one shape ------------------- */
/* a2 + a3: out_lfb[i] = (row, col, free_run) of the row-major first free cell, or
 * (-1, -1, 0) on a full board (get_next_lfb_on_grid :82-92, get_next_occupied_col :290-305) */
int bp_lfb(bp_context *ctx, int n, int rows, int cols, const uint32_t *occ, int32_t *out_lfb);

/* a4 / a8: out_mask[i][e] = 1 iff entry e may be placed at the LFB
 * (get_valid_next_moves :307-317, get_valid_tile_actions_indexes_given_grid :359-381).
 * A full board gives an all-zero mask and lfb (-1, -1, 0). */
int bp_valid_masks(bp_context *ctx, int n, int rows, int cols, int n_entries,
                   const uint32_t *occ, const uint8_t *tiles, uint8_t *out_mask,
                   int32_t *out_lfb);

/* a5 + a6 + a7 (+ a13 add_tile): place entry actions[i] at the LFB of state i.
 * out_verdict[i] is a bp_verdict.  On BP_PLACED out_occ[i] has the footprint filled and
 * out_tiles[i] has the tile and its rotation removed: compact != 0 shifts the table
 * left and pads (0,0) at the end (eliminate_pair_tiles + pad, BinPackLogic.py:55-63);
 * compact == 0 zeroes the two entries in place.  Otherwise outputs copy the inputs.
 * out_occ / out_tiles may be NULL (get_only_success). */
int bp_transitions(bp_context *ctx, int n, int rows, int cols, int n_entries,
                   const uint32_t *occ, const uint8_t *tiles, const int32_t *actions,
                   int compact, uint32_t *out_occ, uint8_t *out_tiles, int32_t *out_verdict,
                   int32_t *out_lfb);

/* a13/a14 getGameEnded (BinPackLogic.py:90-101, BinPackGame.py:61-68):
 * out_value[i] = 0 while moves remain, 1 when the board is full, else filled/total. */
int bp_game_ended(bp_context *ctx, int n, int rows, int cols, int n_entries,
                  const uint32_t *occ, const uint8_t *tiles, double *out_value);

/* ---- a10: whole random rollouts (CustomMCTS.perform_simulation, MCTS.py:152-201) - */
/* For root i and sim s in [sim_begin, sim_begin + n_sims): the Philox stream is
 * key (seed, streams[i]), counter (s, tags[i], block, 0) (oracle/philox.py).
 * out_steps[i][s - sim_begin]  = placements made (the depth the reference returns)
 * out_solved[i][s - sim_begin] = 1 iff every tile was used
 * out_moves (nullable) [i][s - sim_begin][n_entries / 2][2] = (h, w) per placement */
int bp_rollouts(bp_context *ctx, int n_roots, int rows, int cols, int n_entries,
                const uint32_t *occ, const uint8_t *tiles, uint32_t seed,
                const uint32_t *streams, const uint32_t *tags, uint32_t sim_begin, int n_sims,
                uint8_t *out_steps, uint8_t *out_solved, uint8_t *out_moves);

/* ---- a11 + a12: flat Monte-Carlo search (CustomMCTS.predict, MCTS.py:44-149) ----- */
typedef struct {
    int32_t depth;               /* committed moves (2nd return value of predict)        */
    int32_t solution_found;      /* 3rd return value                                      */
    int32_t score;               /* 0 if found else n_tiles - depth (run_mcts.py:221-224) */
    int32_t n_order;             /* length of solution_tiles_order                        */
    int64_t n_tiles_placed;      /* CustomMCTS.n_tiles_placed                             */
    int64_t rollouts;            /* perform_simulation calls the reference would make     */
    int64_t rollout_placements;  /* placements inside those rollouts                      */
    int64_t rollouts_executed;   /* rollouts the device ran (no early exit inside a depth)*/
    int64_t placements_executed; /* placements inside the rollouts the device ran         */
} bp_search_result;

/* Problems of mixed shapes in one batch.
 *   rows, cols, n_entries : int32[n_problems]
 *   tiles   : uint8[n_problems][entry_stride][2], (h, w), reference order, (0,0) padded
 *   boards  : NULL (empty boards) or uint32[n_problems][max_rows][max_wpr] initial occupancy
 *             with max_rows = max(rows), max_wpr = ceil(max(cols) / 32)
 *   streams : uint32[n_problems] Philox key word 1 per problem (NULL: problem index)
 * Rollout (depth d, entry e, sim s) uses counter (s, d << 16 | e, block, 0). */
int bp_search_create(bp_context *ctx, int n_problems, const int32_t *rows, const int32_t *cols,
                     const int32_t *n_entries, const uint8_t *tiles, int entry_stride,
                     const uint32_t *boards, const uint32_t *streams, int n_sim, int strategy,
                     uint32_t seed, bp_search **out);
/* Root-parallel split: this rank runs sims [sim_begin, sim_begin + n_local) of the n_sim
 * per child; n_ranks ranks hold contiguous ascending ranges (default: 0, n_sim, 1). */
int bp_search_set_partition(bp_search *s, int rank, int n_ranks, int sim_begin, int n_local);
/* (re)load the initial states into HBM; must precede bp_search_run */
int bp_search_reset(bp_search *s);
/* How bp_search_run plays the search (results are identical either way):
 *   BP_SEARCH_STEPPED   one rollout launch + one advance launch per depth and shape class, each
 *                       class's depth loop on its own stream (the form the stepwise API uses);
 *   BP_SEARCH_RESIDENT  ONE launch runs every depth of every problem, all shape classes
 *                       together: persistent warps take work units from a device queue, and the
 *                       warp that finishes a problem's last rollout of a depth scores its
 *                       children, commits the best one (CustomMCTS.predict,
 *                       binpack/MCTS.py:95-119) and queues the next depth's rollouts, so problems
 *                       advance independently.  Needs the lane kernel (see bp_search_info) and
 *                       fewer than 2^20 problems; the only form that runs a partitioned
 *                       (root-parallel) search without a host step per move;
 *   BP_SEARCH_HYBRID    the first depths (plenty of work each) launch-per-depth, the remaining
 *                       ones -- where a launch-per-depth step is mostly launch gaps and kernel
 *                       tails -- by the resident kernel: every chain's last advance hands its live
 *                       problems to the resident kernel's queues.  Same requirements as
 *                       BP_SEARCH_RESIDENT, single rank only; falls back to one of the two pure
 *                       forms when the hand-over depth is the first or the last one;
 *   BP_SEARCH_AUTO      the fastest of these as measured on B200.
 * The environment variable BP_SEARCH=stepped|resident|hybrid overrides the mode.
 * bp_search_last_run: 0 = launch-per-depth (warp-per-board advance, rebuilt tables), 1 = resident,
 * 2 = hybrid, 3 = launch-per-depth on the skyline state (the resident kernel's playout and advance). */
typedef enum { BP_SEARCH_AUTO = 0, BP_SEARCH_STEPPED = 1, BP_SEARCH_RESIDENT = 2, BP_SEARCH_HYBRID = 3 } bp_search_mode;
int bp_search_set_mode(bp_search *s, int mode);
/* on != 0: time the launches of the next run with CUDA events on the context's stream, shape
 * classes serialised.  bp_search_profile synchronises and returns, for the last run,
 *   stepped:  milliseconds in the rollout kernels / in the advance kernels, over n_depths depths;
 *   resident: milliseconds in the resident kernel (rollout_ms; advance_ms = 0), n_depths = 1. */
int bp_search_set_profiling(bp_search *s, int on);
int bp_search_profile(bp_search *s, double *rollout_ms, double *advance_ms, int *n_depths);
/* run every depth; asynchronous; single rank only.  Each shape class runs on an internal
 * stream forked from and joined to the context's stream with events, so the call is ordered
 * with other work on the context's stream like a single kernel would be.  bp_search_results
 * fails with BP_ERR_STATE if a resident kernel gave up (its watchdog fired). */
int bp_search_run(bp_search *s);
/* stepwise form (root-parallel): for depth d = 0, 1, ... until bp_search_max_depth:
 *   bp_search_step_rollouts(s)      rollouts for every child of every live problem
 *   bp_search_step_reduce(s)        per-child statistics of this rank -> stats buffer
 *   (all-gather the stats buffers of all ranks into gathered_dev, rank-major)
 *   bp_search_step_commit(s, gathered_dev)   merge, commit the best child, expand      */
int bp_search_step_rollouts(bp_search *s);
int bp_search_step_reduce(bp_search *s);
int bp_search_step_commit(bp_search *s, const void *gathered_stats_dev);
/* resident form of the root-parallel search (no host step per move): after
 * bp_search_set_partition every rank creates its exchange buffer, the ranks swap the 64-byte
 * CUDA IPC handles (ranks in other processes; ipc_handles = uint8[n_ranks][64] in rank order) or
 * the raw pointers (ranks of one process; local_ptrs[n_ranks]) and connect; then
 * bp_search_reset + bp_search_run in BP_SEARCH_RESIDENT mode on every rank, concurrently.
 * Inside the kernel the warp that finishes a problem's depth stores this rank's per-child
 * statistics into every rank's buffer over NVLink peer memory, raises a flag there and waits
 * for the other ranks' flags: the merge of binpack/MCTS.py:126-149 split over GPUs.
 * All ranks must reset / run the same number of times; synchronise the ranks (host barrier)
 * before destroying a connected search. */
int bp_search_exchange_create(bp_search *s, void *ipc_handle_64, void **local_ptr);
int bp_search_exchange_connect(bp_search *s, const void *ipc_handles, void *const *local_ptrs);
/* resident kernel blocks per SM (0 = as many as fit): lets several searches be resident on one
 * device at once (ranks emulated on one GPU wait for each other inside their kernels) */
int bp_search_set_resident_blocks(bp_search *s, int blocks_per_sm);
int bp_search_stats_buffer(bp_search *s, void **stats_dev, int64_t *n_bytes);
int bp_search_max_depth(bp_search *s, int *max_depth);
/* lane_kernel = 1: rollouts run in the lane-per-rollout kernel (32 rollouts per warp, needs
 * empty initial boards), 0: in the warp-per-board kernel (any initial board; also forced by
 * the environment variable BP_ROLLOUT=warp).  n_classes = shape classes in the batch, each
 * with its own launch sequence. */
int bp_search_info(bp_search *s, int *lane_kernel, int *n_classes);
/* resident = 1 when the last bp_search_run used the resident kernel */
int bp_search_last_run(bp_search *s, int *resident);
/* Where the warps of the last resident run spent their time; synchronises.  out13 = SM clock
 * cycles summed over all warps: {resident, waiting for a unit, playing units, advancing
 * problems}, then {units played, problems advanced, polls of an empty slot}, cycles spent in
 * the peer-memory exchange (root-parallel), units published, warps launched, and the cycles of
 * the advance by phase {merging the unit records, commit + expand, publishing the next depth}. */
int bp_search_resident_stats(bp_search *s, double *out13);
/* synchronises the stream and copies results to the host
 *   results      : bp_search_result[n_problems]
 *   order        : NULL or uint8[n_problems][entry_stride / 2 + 2][2]  solution_tiles_order
 *   path         : NULL or int32[n_problems][entry_stride / 2]  committed entry per depth
 *                  (index into that depth's live tile list), -1 beyond depth
 *   child_scores : NULL or int32[n_problems][entry_stride / 2][entry_stride]
 *                  per depth and entry: max depth or sum of depths over the n_sim rollouts,
 *                  -1 = entry could not be placed / not evaluated, -2 = solved child      */
int bp_search_results(bp_search *s, bp_search_result *results, uint8_t *order, int32_t *path,
                      int32_t *child_scores);
int bp_search_destroy(bp_search *s);

/* one-shot convenience: create + reset + run + results + destroy (host buffers in/out) */
int bp_flat_search(bp_context *ctx, int n_problems, const int32_t *rows, const int32_t *cols,
                   const int32_t *n_entries, const uint8_t *tiles, int entry_stride,
                   const uint32_t *boards, const uint32_t *streams, int n_sim, int strategy,
                   uint32_t seed, bp_search_result *results, uint8_t *order, int32_t *path,
                   int32_t *child_scores);

/* ---- a15 + a16: PUCT tree search (MCTS.search / getActionProb, MCTS.py:204-335) -- */
/* n_trees independent trees of one shape searched in lock step, one warp per tree.
 * Nodes are keyed by the exact (board, tile list) state, like the reference's string
 * keys, so transpositions share statistics.  A tree keeps its nodes across
 * bp_tree_set_root calls (the reference keeps its dicts for a whole episode) until
 * bp_tree_reset.  One simulation of every tree:
 *   bp_tree_select  walks root -> leaf/terminal with argmax UCB (float64, lowest action
 *                   wins ties).  out_status[i] = 0: tree i stopped at an unexpanded leaf
 *                   whose board, padded tile list and valid-action mask are returned for
 *                   nnet.predict; 1: it stopped at a terminal node (nothing to evaluate).
 *   bp_tree_backup  priors[i][a] = the leaf's policy already masked with the valid moves
 *                   and normalised (MCTS.py:281-292, done by the caller next to the net),
 *                   values[i] = the net's v; both ignored for trees with status 1.
 *                   Updates Qsa/Nsa/Ns along each path (:326-334); out_values (nullable)
 *                   receives the v each search() call returns.                          */
int bp_tree_create(bp_context *ctx, int n_trees, int rows, int cols, int n_entries,
                   int max_nodes, double cpuct, bp_tree **out);
int bp_tree_reset(bp_tree *t);
int bp_tree_set_root(bp_tree *t, const uint32_t *occ, const uint8_t *tiles);
int bp_tree_select(bp_tree *t, uint32_t *out_leaf_occ, uint8_t *out_leaf_tiles,
                   uint8_t *out_valid, int32_t *out_status);
int bp_tree_backup(bp_tree *t, const double *priors, const double *values, double *out_values);
/* Device-resident form of one simulation (many trees, PyTorch net on the same stream): the
 * leaf boards (uint32[n_trees][rows][wpr]), padded tile lists (uint8[n_trees][E][2]), valid
 * masks (uint8[n_trees][E]) and status (int32[n_trees]) stay in device buffers owned by the
 * tree; bp_tree_backup_dev takes the net's raw float32 policy [n_trees][E] and value
 * [n_trees] device tensors and does the masking, normalisation and uniform fallback of
 * MCTS.py:281-292 on the device.  No host copy, no synchronisation; bp_tree_check reports
 * a node-pool overflow afterwards. */
int bp_tree_select_dev(bp_tree *t, void **leaf_occ_dev, void **leaf_tiles_dev, void **valid_dev,
                       void **status_dev);
int bp_tree_backup_dev(bp_tree *t, const void *policy_f32_dev, const void *values_f32_dev);
int bp_tree_check(bp_tree *t, int *overflow);
/* counts[i][a] = Nsa[(root_i, a)]; n_nodes (nullable)[i] = nodes in tree i */
int bp_tree_root_counts(bp_tree *t, int32_t *counts, int32_t *n_nodes);
/* same counts left on the device (int32[n_trees][n_entries]) for an NCCL all-reduce */
int bp_tree_root_counts_dev(bp_tree *t, void **counts_dev, int64_t *n_bytes);
/* root-parallel PUCT (independent trees per GPU from the same roots): Nsa (int32) and
 * Wsa = Qsa * Nsa (float64), both [n_trees][n_entries], left on the device; the ranks all-reduce
 * SUM both before getActionProb (binpack/MCTS.py:221-244 consumes the counts). */
int bp_tree_root_stats_dev(bp_tree *t, void **counts_dev, void **wsa_dev);
int bp_tree_destroy(bp_tree *t);

#ifdef __cplusplus
}
#endif
#endif /* RECTPACK_B200_H */
