/* Plain-C restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 * See rectpack_oracle.h for the contract; every function cites the reference
 * lines it follows (paths relative to /root/reference).
 */
#include "rectpack_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* ---------------------------------------------------------------- Philox ---- */
/* Philox4x32-10, Salmon et al. SC'11; same function as oracle/philox.py. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct {
    uint32_t key[2], sim, tag, words[4];
    int s, block;
} draw_stream;

static void stream_init(draw_stream *d, uint32_t seed, uint32_t stream, uint32_t tag, uint32_t sim)
{
    d->key[0] = seed; d->key[1] = stream; d->sim = sim; d->tag = tag; d->s = 0; d->block = -1;
}

/* index in [0, k): (u32 * k) >> 32, one draw per call */
static int stream_draw(draw_stream *d, int k)
{
    int b = d->s >> 2;
    if (b != d->block) {
        uint32_t ctr[4] = { d->sim, d->tag, (uint32_t)b, 0u };
        orc_philox4x32_10(ctr, d->key, d->words);
        d->block = b;
    }
    uint32_t u = d->words[d->s & 3];
    d->s++;
    return (int)(((uint64_t)u * (uint64_t)k) >> 32);
}

/* ------------------------------------------------------------------- a1 ---- */
int64_t orc_find_first(int32_t needle, const int32_t *haystack, int64_t n)
{
    for (int64_t k = 0; k < n; ++k)
        if (haystack[k] == needle) return k;
    return -1;
}

/* ------------------------------------------------------------------- a2 ---- */
int orc_next_lfb(const int32_t *grid, int rows, int cols, int *r, int *c)
{
    int64_t k = orc_find_first(0, grid, (int64_t)rows * cols);
    if (k < 0) return 0;
    *r = (int)(k / cols);
    *c = (int)(k % cols);
    return 1;
}

static int cell_occupied(int32_t v, int colorful) { return colorful ? (v != 0) : (v == 1); }

/* ------------------------------------------------------------------- a3 ---- */
int orc_next_occupied_col(const int32_t *grid, int rows, int cols, int r, int c, int colorful)
{
    (void)rows;
    for (int j = c; j < cols; ++j)
        if (cell_occupied(grid[(int64_t)r * cols + j], colorful)) return j;
    return cols;
}

/* ------------------------------------------------------------------- a4 ---- */
int orc_valid_next_moves(const int32_t *grid, int rows, int cols, const int32_t *tiles,
                         int n_entries, int colorful, uint8_t *flags)
{
    int r, c;
    if (!orc_next_lfb(grid, rows, cols, &r, &c)) return -1;
    int room = orc_next_occupied_col(grid, rows, cols, r, c, colorful) - c;
    int count = 0;
    for (int i = 0; i < n_entries; ++i) {
        int th = tiles[2 * i], tw = tiles[2 * i + 1];
        int ok = (tw <= room) && (th + r <= rows);
        flags[i] = (uint8_t)ok;
        count += ok;
    }
    return count;
}

/* ------------------------------------------------------------------- a5 ---- */
int orc_place(int32_t *grid, int rows, int cols, int th, int tw, int r, int c, int32_t val,
              int only_success, int colorful)
{
    if (c + tw > cols) return 0;
    if (r + th > rows) return 0;
    /* only the first footprint row is tested ("height is always free", :107) */
    for (int j = c; j < c + tw; ++j)
        if (cell_occupied(grid[(int64_t)r * cols + j], colorful)) return 0;
    if (only_success) return 1;
    for (int i = r; i < r + th; ++i)
        for (int j = c; j < c + tw; ++j)
            grid[(int64_t)i * cols + j] = val;
    return 1;
}

/* ------------------------------------------------------------------- a6 ---- */
int orc_next_turn(int32_t *grid, int rows, int cols, int n_entries, int th, int tw,
                  int32_t val, int only_success, int colorful)
{
    int r, c;
    if (!orc_next_lfb(grid, rows, cols, &r, &c))
        return n_entries == 0 ? ORC_ALL_TILES_USED : ORC_NO_NEXT_POSITION_TILES_UNUSED;
    if (!orc_place(grid, rows, cols, th, tw, r, c, val, only_success, colorful))
        return ORC_TILE_CANNOT_BE_PLACED;
    return ORC_PLACED;
}

/* ------------------------------------------------------------------- a7 ---- */
static int tiles_index(const int32_t *tiles, int n, int th, int tw)
{
    for (int i = 0; i < n; ++i)
        if (tiles[2 * i] == th && tiles[2 * i + 1] == tw) return i;
    return -1;
}

static void tiles_remove(int32_t *tiles, int n, int i)
{
    memmove(tiles + 2 * i, tiles + 2 * (i + 1), (size_t)(n - i - 1) * 2 * sizeof(int32_t));
}

int orc_eliminate_pair(int32_t *tiles, int n_entries, int th, int tw)
{
    int i = tiles_index(tiles, n_entries, th, tw);
    if (i < 0) return -1;
    tiles_remove(tiles, n_entries, i);
    int j = tiles_index(tiles, n_entries - 1, tw, th);
    if (j < 0) return -2;
    tiles_remove(tiles, n_entries - 1, j);
    return n_entries - 2;
}

/* ------------------------------------------------------------------- a8 ---- */
int orc_valid_action_mask(const int32_t *grid, int rows, int cols, const int32_t *tiles,
                          int n_entries, uint8_t *mask)
{
    int r, c;
    if (!orc_next_lfb(grid, rows, cols, &r, &c)) return -1;
    for (int i = 0; i < n_entries; ++i) {
        int th = tiles[2 * i], tw = tiles[2 * i + 1];
        if (th == 0 && tw == 0) { mask[i] = 0; continue; }
        /* val=1, default colorful_states=False: occupied iff == 1 (:373-376) */
        mask[i] = (uint8_t)orc_place((int32_t *)grid, rows, cols, th, tw, r, c, 1, 1, 0);
    }
    return 0;
}

/* ------------------------------------------------------------------ a10 ---- */
void orc_rollout(int32_t *grid, int rows, int cols, int32_t *tiles, int n_entries,
                 uint32_t seed, uint32_t stream, uint32_t tag, uint32_t sim, int32_t val,
                 int32_t *moves, orc_rollout_result *res)
{
    draw_stream ds;
    uint8_t flags[ORC_MAX_ENTRIES];
    stream_init(&ds, seed, stream, tag, sim);
    res->solved = 0; res->depth = 0; res->placements = 0;
    if (n_entries == 0) { res->solved = 1; return; }          /* MCTS.py:161-163 */
    int n = n_entries;
    for (;;) {
        val += 1;
        if (n == 0) { res->solved = 1; return; }              /* :167-169 */
        int k = orc_valid_next_moves(grid, rows, cols, tiles, n, 1, flags);
        if (k <= 0) return;                                    /* :171-172 */
        int pick = stream_draw(&ds, k);                        /* :174 */
        int e = -1;
        for (int i = 0; i < n; ++i)
            if (flags[i] && pick-- == 0) { e = i; break; }
        int th = tiles[2 * e], tw = tiles[2 * e + 1];
        int verdict = orc_next_turn(grid, rows, cols, n, th, tw, val, 0, 1);   /* :175-178 */
        res->placements += 1;                                  /* :179 */
        if (moves) { moves[2 * res->depth] = th; moves[2 * res->depth + 1] = tw; }
        if (verdict != ORC_PLACED) return;                     /* :187-192, unreachable */
        n = orc_eliminate_pair(tiles, n, th, tw);              /* :194 */
        res->depth += 1;                                       /* :200 */
    }
}

/* ------------------------------------------------------------- a11 + a12 ---- */
typedef struct {
    int solved;          /* 1: a rollout used all tiles */
    int64_t score;       /* max of depths, or sum of depths (avg strategy, fixed N) */
    int64_t ran;
    int64_t placements;
    int n_moves;         /* moves of the solved rollout */
} sims_result;

static void run_simulations(const int32_t *grid, int rows, int cols, const int32_t *tiles,
                            int n_entries, int n_sim, int strategy, uint32_t seed,
                            uint32_t stream, uint32_t tag, int32_t val, int32_t *scratch_grid,
                            int32_t *moves, sims_result *out)
{
    int32_t tl[2 * ORC_MAX_ENTRIES];
    orc_rollout_result rr;
    int64_t best = 0, sum = 0;
    out->solved = 0; out->ran = 0; out->placements = 0; out->n_moves = 0;
    for (int n = 0; n < n_sim; ++n) {                          /* MCTS.py:134 */
        memcpy(scratch_grid, grid, (size_t)rows * cols * sizeof(int32_t));   /* state.copy() */
        memcpy(tl, tiles, (size_t)n_entries * 2 * sizeof(int32_t));
        orc_rollout(scratch_grid, rows, cols, tl, n_entries, seed, stream, tag, (uint32_t)n,
                    val, moves, &rr);
        out->ran += 1;
        out->placements += rr.placements;
        if (rr.solved) {                                       /* :136-138 */
            out->solved = 1;
            out->n_moves = rr.depth;
            return;
        }
        if (rr.depth > best) best = rr.depth;
        sum += rr.depth;
    }
    out->score = strategy == 0 ? best : sum;                   /* :144-147 */
}

int orc_flat_search(int rows, int cols, const int32_t *tiles_in, int n_entries, int n_sim,
                    int strategy, uint32_t seed, uint32_t stream, const int32_t *board,
                    orc_search_result *res, int32_t *order, int64_t *children_log,
                    int children_cap, int32_t *path)
{
    if (n_entries > ORC_MAX_ENTRIES || n_entries < 0) return -1;
    size_t cells = (size_t)rows * cols;
    int32_t *grid = (int32_t *)calloc(cells, sizeof(int32_t));
    int32_t *child = (int32_t *)malloc(cells * sizeof(int32_t));
    int32_t *best_grid = (int32_t *)malloc(cells * sizeof(int32_t));
    int32_t *scratch = (int32_t *)malloc(cells * sizeof(int32_t));
    int32_t tiles[2 * ORC_MAX_ENTRIES], rest[2 * ORC_MAX_ENTRIES], best_rest[2 * ORC_MAX_ENTRIES];
    int32_t moves[2 * ORC_MAX_ENTRIES];
    if (board) memcpy(grid, board, cells * sizeof(int32_t));
    memcpy(tiles, tiles_in, (size_t)n_entries * 2 * sizeof(int32_t));
    memset(res, 0, sizeof(*res));

    int n = n_entries, depth = 0, n_order = 0;
    int32_t val = 1;
    int have_prev = 0, prev_h = 0, prev_w = 0;     /* state.tile_placed */
    int found = 0;

    while (n > 0) {                                            /* MCTS.py:53 */
        int placed_any = 0, best_entry = -1, best_n = 0, best_h = 0, best_w = 0;
        int64_t best_score = 0;
        for (int i = 0; i < n; ++i) {                          /* :58 */
            int th = tiles[2 * i], tw = tiles[2 * i + 1];
            memcpy(child, grid, cells * sizeof(int32_t));      /* destroy_state=False */
            int verdict = orc_next_turn(child, rows, cols, n, th, tw, val, 0, 1);
            res->n_tiles_placed += 1;                          /* :60 */
            if (verdict == ORC_ALL_TILES_USED) { found = 1; goto done; }
            if (verdict == ORC_TILE_CANNOT_BE_PLACED) continue;
            if (verdict != ORC_PLACED) continue;               /* full board, tiles left */
            placed_any = 1;
            memcpy(rest, tiles, (size_t)n * 2 * sizeof(int32_t));
            int rn = orc_eliminate_pair(rest, n, th, tw);      /* :72 */
            if (rn < 0) { free(grid); free(child); free(best_grid); free(scratch); return -2; }
            sims_result sr;
            run_simulations(child, rows, cols, rest, rn, n_sim, strategy, seed, stream,
                            ((uint32_t)depth << 16) | (uint32_t)i, val, scratch, moves, &sr);
            res->rollouts += sr.ran;
            res->rollout_placements += sr.placements;
            res->n_tiles_placed += sr.placements;
            if (children_log && res->n_children < children_cap) {
                int64_t *row = children_log + 4 * (int64_t)res->n_children;
                row[0] = depth; row[1] = i; row[2] = sr.solved; row[3] = sr.solved ? 0 : sr.score;
            }
            res->n_children += 1;
            if (sr.solved) {                                   /* :77-93 */
                found = 1;
                if (have_prev) { order[2 * n_order] = prev_h; order[2 * n_order + 1] = prev_w; n_order++; }
                order[2 * n_order] = th; order[2 * n_order + 1] = tw; n_order++;
                for (int m = 0; m < sr.n_moves; ++m) {
                    order[2 * n_order] = moves[2 * m]; order[2 * n_order + 1] = moves[2 * m + 1];
                    n_order++;
                }
                goto done;
            }
            /* get_max_index (:20-29): strictly greater, so the first maximum wins */
            if (best_entry < 0 || sr.score > best_score) {
                best_entry = i; best_score = sr.score; best_n = rn; best_h = th; best_w = tw;
                memcpy(best_grid, child, cells * sizeof(int32_t));
                memcpy(best_rest, rest, (size_t)rn * 2 * sizeof(int32_t));
            }
        }
        if (!placed_any) goto done;                            /* :101-103 */
        val += 1;
        depth += 1;
        if (have_prev) { order[2 * n_order] = prev_h; order[2 * n_order + 1] = prev_w; n_order++; }  /* :115-116 */
        if (path) path[res->n_path] = best_entry;
        res->n_path += 1;
        memcpy(grid, best_grid, cells * sizeof(int32_t));
        memcpy(tiles, best_rest, (size_t)best_n * 2 * sizeof(int32_t));
        n = best_n;
        have_prev = 1; prev_h = best_h; prev_w = best_w;
    }
    found = 1;                                                 /* :121-123 */
done:
    res->depth = depth;
    res->solution_found = found;
    res->score = found ? 0 : n_entries / 2 - depth;            /* run_mcts.py:221-224 */
    res->n_order = n_order;
    free(grid); free(child); free(best_grid); free(scratch);
    return 0;
}

int orc_num_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    atomic_int next;
    int n_problems, max_entries, n_sim, strategy, order_stride;
    uint32_t seed;
    const int32_t *rows, *cols, *tiles, *n_entries;
    const uint32_t *streams;
    orc_search_result *res;
    int32_t *orders;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    int32_t dummy_order[2 * (ORC_MAX_ENTRIES / 2 + 2)];
    for (;;) {
        int p = atomic_fetch_add(&j->next, 1);
        if (p >= j->n_problems) break;
        orc_flat_search(j->rows[p], j->cols[p], j->tiles + (int64_t)p * j->max_entries * 2,
                        j->n_entries[p], j->n_sim, j->strategy, j->seed,
                        j->streams ? j->streams[p] : (uint32_t)p, NULL, &j->res[p],
                        j->orders ? j->orders + (int64_t)p * j->order_stride : dummy_order,
                        NULL, 0, NULL);
    }
    return NULL;
}

/* one pthread per requested host thread, problems handed out one at a time */
int orc_flat_search_batch(int n_problems, const int32_t *rows, const int32_t *cols,
                          const int32_t *tiles, const int32_t *n_entries, int max_entries,
                          int n_sim, int strategy, uint32_t seed, const uint32_t *streams,
                          orc_search_result *res, int32_t *orders, int n_threads)
{
    batch_job job;
    if (n_threads <= 0) n_threads = orc_num_threads();
    if (n_threads > n_problems) n_threads = n_problems > 0 ? n_problems : 1;
    atomic_init(&job.next, 0);
    job.n_problems = n_problems; job.max_entries = max_entries; job.n_sim = n_sim;
    job.strategy = strategy; job.order_stride = 2 * (max_entries / 2 + 2); job.seed = seed;
    job.rows = rows; job.cols = cols; job.tiles = tiles; job.n_entries = n_entries;
    job.streams = streams; job.res = res; job.orders = orders;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    int started = 0;
    for (int t = 0; t < n_threads; ++t)
        if (pthread_create(&th[t], NULL, batch_worker, &job) == 0) started++; else break;
    if (started == 0) batch_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
    return 0;
}

/* ------------------------------------------------------------------ a13 ---- */
int orc_board_add_tile(int32_t *grid, int rows, int cols, int32_t *tiles, int n_entries,
                       int action)
{
    int r, c;
    if (!orc_next_lfb(grid, rows, cols, &r, &c)) return -1;    /* reference raises TypeError */
    int th = tiles[2 * action], tw = tiles[2 * action + 1];
    if (!orc_place(grid, rows, cols, th, tw, r, c, 1, 0, 0)) return 0;   /* :49-52 */
    int rn = orc_eliminate_pair(tiles, n_entries, th, tw);     /* :55-56 */
    if (rn < 0) return -2;
    for (int i = rn; i < n_entries; ++i) { tiles[2 * i] = 0; tiles[2 * i + 1] = 0; }  /* :59-63 */
    return 1;
}

int orc_board_win_state(const int32_t *grid, int rows, int cols, const int32_t *tiles,
                        int n_entries, double *value)
{
    uint8_t mask[ORC_MAX_ENTRIES];
    int all_placed = 1, any_valid = 0;
    for (int i = 0; i < n_entries; ++i)
        if (tiles[2 * i] != 0 || tiles[2 * i + 1] != 0) { all_placed = 0; break; }
    if (!all_placed) {
        if (orc_valid_action_mask(grid, rows, cols, tiles, n_entries, mask) == 0)
            for (int i = 0; i < n_entries; ++i) any_valid |= mask[i];
    }
    if (all_placed || !any_valid) {                            /* :91-98 */
        double filled = 0.0;
        for (int64_t k = 0; k < (int64_t)rows * cols; ++k) filled += grid[k];
        double total = (double)rows * cols;
        *value = (filled == total) ? 1.0 : filled / total;
        return 1;
    }
    *value = 0.0;
    return 0;
}
