// PUCT tree search (MCTS.search / getActionProb, binpack/MCTS.py:204-335) on the device.
//
// B independent trees are searched in lock step, one warp per tree.  A simulation is two
// launches: `select` walks from the root to a leaf or a terminal node (UCB argmax over the
// legal actions with lane = action, the transition of BinPackLogic.add_tile, a transposition
// lookup keyed by the exact (board, tile list) state, as the reference's string keys do), the
// host evaluates the leaves with the neural net as ONE batch, `backup` stores the priors and
// updates N/Q/Ns along each path.
//
// Tree statistics are struct-of-arrays in HBM, action-minor so that lane = action loads and
// stores are coalesced:  Nsa[tree][node][a] (int32), Qsa / Psa[tree][node][a] (float64),
// valid[tree][node] (4 x 32-bit masks), Ns / Es per node.  All UCB arithmetic is IEEE
// float64 with explicit round-to-nearest intrinsics (no FMA contraction) in the reference's
// order of operations, so visit counts match the Python implementation bit for bit.
#include <algorithm>
#include <vector>

#include "board.cuh"
#include "common.cuh"

namespace bp {

constexpr int TREE_WARPS = 4;
constexpr double PUCT_EPS = 0.1;          // EPS, binpack/MCTS.py:6

struct TreeParams {
    int n_trees, rows, cols, wpr, n_entries, cap, hcap, max_path;
    double cpuct;
    uint32_t* occ;        // [B][cap][rows*wpr]
    uint8_t* tiles;       // [B][cap][E][2]
    int* nsa;             // [B][cap][E]
    double* qsa;          // [B][cap][E]
    double* psa;          // [B][cap][E]
    uint4* valid;         // [B][cap]
    int* ns;              // [B][cap]
    double* es;           // [B][cap]
    uint8_t* flags;       // [B][cap] bit0: Es known, bit1: expanded (Ps known)
    unsigned long long* hash;  // [B][cap]
    int* table;           // [B][hcap] node index or -1
    int* n_nodes;         // [B]
    int* root;            // [B]
    int* path_node;       // [B][max_path]
    int* path_action;     // [B][max_path]
    int* path_len;        // [B]
    int* leaf;            // [B] node awaiting evaluation, -1 = terminal
    double* leaf_value;   // [B] terminal value
    int* overflow;        // [1] set when a tree runs out of nodes
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

template <int RHO, int W, int EJ>
__device__ __forceinline__ unsigned long long state_hash(const WarpState<RHO, W, EJ>& s)
{
    const int lane = lane_id();
    unsigned long long h = 0;
#pragma unroll
    for (int k = 0; k < RHO; ++k)
#pragma unroll
        for (int q = 0; q < W; ++q)
            h += mix64(((unsigned long long)(1 + (k * W + q) * 32 + lane) << 32) | s.occ[k][q]);
#pragma unroll
    for (int j = 0; j < EJ; ++j)
        h += mix64(((unsigned long long)(0x10000 + j * 32 + lane) << 32) | s.ent[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(FULLMASK, h, o);
    return mix64(h) | 1ull;
}

template <int RHO, int W, int EJ>
__device__ __forceinline__ bool state_equal(const WarpState<RHO, W, EJ>& a, const WarpState<RHO, W, EJ>& b)
{
    bool eq = true;
#pragma unroll
    for (int k = 0; k < RHO; ++k)
#pragma unroll
        for (int q = 0; q < W; ++q) eq = eq && (a.occ[k][q] == b.occ[k][q]);
#pragma unroll
    for (int j = 0; j < EJ; ++j) eq = eq && (a.ent[j] == b.ent[j]);
    return __all_sync(FULLMASK, eq);
}

template <int RHO, int W, int EJ>
__device__ __forceinline__ void load_node(const TreeParams& p, int tree, int node, WarpState<RHO, W, EJ>& s)
{
    const size_t idx = (size_t)tree * p.cap + node;
    load_board(s, p.occ + idx * p.rows * p.wpr, p.rows, p.cols, p.wpr);
    load_entries(s, p.tiles + idx * p.n_entries * 2, p.n_entries);
}

// find the node holding state s, or create it; returns -1 when the tree is full
template <int RHO, int W, int EJ>
__device__ __forceinline__ int find_or_insert(const TreeParams& p, int tree, const WarpState<RHO, W, EJ>& s)
{
    const int lane = lane_id();
    const unsigned long long h = state_hash(s);
    int* table = p.table + (size_t)tree * p.hcap;
    unsigned slot = (unsigned)(h >> 20) & (unsigned)(p.hcap - 1);
    for (int probe = 0; probe < p.hcap; ++probe) {
        const int node = table[slot];
        if (node < 0) break;
        if (p.hash[(size_t)tree * p.cap + node] == h) {
            WarpState<RHO, W, EJ> t;
            load_node(p, tree, node, t);
            if (state_equal(s, t)) return node;
        }
        slot = (slot + 1) & (unsigned)(p.hcap - 1);
    }
    const int node = p.n_nodes[tree];
    if (node >= p.cap) {
        if (lane == 0) *p.overflow = 1;
        return -1;
    }
    __syncwarp();
    const size_t idx = (size_t)tree * p.cap + node;
    store_board(s, p.occ + idx * p.rows * p.wpr, p.rows, p.cols, p.wpr);
    store_entries(s, p.tiles + idx * p.n_entries * 2, p.n_entries);
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        const int a = lane + 32 * j;
        if (a < p.n_entries) {
            p.nsa[idx * p.n_entries + a] = 0;
            p.qsa[idx * p.n_entries + a] = 0.0;
            p.psa[idx * p.n_entries + a] = 0.0;
        }
    }
    if (lane == 0) {
        p.hash[idx] = h;
        p.ns[idx] = 0;
        p.es[idx] = 0.0;
        p.flags[idx] = 0;
        p.valid[idx] = make_uint4(0, 0, 0, 0);
        table[slot] = node;
        p.n_nodes[tree] = node + 1;
    }
    __syncwarp();
    return node;
}

// shift live entries to the front, (0,0) pads at the end (BinPackLogic.py:55-63)
template <int RHO, int W, int EJ>
__device__ __forceinline__ void compact_pad(WarpState<RHO, W, EJ>& s, uint16_t* buf)
{
    const int lane = lane_id();
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < EJ; ++j) buf[lane + 32 * j] = 0;
    __syncwarp();
    int base = 0;
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        const uint32_t m = __ballot_sync(FULLMASK, s.ent[j] != 0u);
        if (s.ent[j] != 0u) buf[base + __popc(m & lt)] = (uint16_t)s.ent[j];
        base += __popc(m);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < EJ; ++j) s.ent[j] = buf[lane + 32 * j];
    __syncwarp();
}

// getGameEnded (BinPackLogic.py:90-101): 0 while moves remain, 1 when full, else filled/total.
// Also returns the legal-action masks of the state.
template <int RHO, int W, int EJ>
__device__ __forceinline__ double game_value(const WarpState<RHO, W, EJ>& s, int rows, int cols,
                                             uint32_t (&legal)[EJ])
{
#pragma unroll
    for (int j = 0; j < EJ; ++j) legal[j] = 0u;
    const int n_alive = count_alive(s);
    int r, c, run, k = 0;
    if (n_alive != 0 && find_lfb(s, r, c, run)) k = legal_masks(s, rows, r, run, legal);
    if (n_alive != 0 && k != 0) return 0.0;
    int filled = 0;
#pragma unroll
    for (int kk = 0; kk < RHO; ++kk)
#pragma unroll
        for (int q = 0; q < W; ++q) filled += __popc(s.occ[kk][q]);
    filled = __reduce_add_sync(FULLMASK, filled);
    const int total = rows * cols;
    filled -= 32 * RHO * 32 * W - total;
    return (filled == total) ? 1.0 : __ddiv_rn((double)filled, (double)total);
}

__device__ __forceinline__ double shfl_xor_f64(double v, int o)
{
    return __shfl_xor_sync(FULLMASK, v, o);
}

template <int RHO, int W, int EJ>
__global__ void __launch_bounds__(TREE_WARPS * 32)
tree_set_root_kernel(TreeParams p, const uint32_t* __restrict__ occ, const uint8_t* __restrict__ tiles)
{
    const int tree = blockIdx.x * TREE_WARPS + (threadIdx.x >> 5);
    if (tree >= p.n_trees) return;
    WarpState<RHO, W, EJ> s;
    load_board(s, occ + (size_t)tree * p.rows * p.wpr, p.rows, p.cols, p.wpr);
    load_entries(s, tiles + (size_t)tree * p.n_entries * 2, p.n_entries);
    const int node = find_or_insert(p, tree, s);
    if (lane_id() == 0) p.root[tree] = node;
}

template <int RHO, int W, int EJ>
__global__ void __launch_bounds__(TREE_WARPS * 32)
tree_select_kernel(TreeParams p, uint32_t* __restrict__ out_occ, uint8_t* __restrict__ out_tiles,
                   uint8_t* __restrict__ out_valid, int32_t* __restrict__ out_status)
{
    __shared__ uint16_t cbuf[TREE_WARPS][32 * EJ];
    const int warp = threadIdx.x >> 5;
    const int tree = blockIdx.x * TREE_WARPS + warp;
    if (tree >= p.n_trees) return;
    const int lane = lane_id();
    int node = p.root[tree];
    int len = 0;
    int status = 1;                    // 1 = terminal, 0 = leaf to evaluate, -1 = error
    double value = 0.0;
    int leaf = -1;
    WarpState<RHO, W, EJ> s;
    uint32_t legal[EJ];
    for (;;) {
        if (node < 0) { status = -1; break; }
        const size_t idx = (size_t)tree * p.cap + node;
        load_node(p, tree, node, s);
        uint8_t fl = p.flags[idx];
        if (!(fl & 1)) {                                    // Es[s] not known yet (:269-270)
            const double v = game_value(s, p.rows, p.cols, legal);
            if (lane == 0) {
                p.es[idx] = v;
                p.flags[idx] = fl | 1;
                p.valid[idx] = make_uint4(legal[0], EJ > 1 ? legal[EJ > 1 ? 1 : 0] : 0u,
                                          EJ > 2 ? legal[EJ > 2 ? 2 : 0] : 0u,
                                          EJ > 3 ? legal[EJ > 3 ? 3 : 0] : 0u);
            }
            __syncwarp();
            fl |= 1;
        }
        const double es = p.es[idx];
        if (es != 0.0) { value = es; status = 1; break; }   // terminal node (:271-273)
        const uint4 vm = p.valid[idx];
        const uint32_t vmask[4] = {vm.x, vm.y, vm.z, vm.w};
        if (!(fl & 2)) {                                    // leaf: needs nnet.predict (:275-296)
            leaf = node;
            status = 0;
            store_board(s, out_occ + (size_t)tree * p.rows * p.wpr, p.rows, p.cols, p.wpr);
            store_entries(s, out_tiles + (size_t)tree * p.n_entries * 2, p.n_entries);
#pragma unroll
            for (int j = 0; j < EJ; ++j) {
                const int a = lane + 32 * j;
                if (a < p.n_entries) out_valid[(size_t)tree * p.n_entries + a] = (vmask[j] >> lane) & 1u;
            }
            break;
        }
        if (len >= p.max_path) { status = -1; break; }
        // ---- argmax_a UCB over the valid actions, lowest a wins ties (:300-318) ----
        const double sqrt_ns = __dsqrt_rn((double)p.ns[idx]);
        const double sqrt_ns_eps = __dsqrt_rn(__dadd_rn((double)p.ns[idx], PUCT_EPS));
        double best_u = -INFINITY;
        int best_a = 0x7FFFFFFF;
#pragma unroll
        for (int j = 0; j < EJ; ++j) {
            const int a = lane + 32 * j;
            if ((vmask[j] >> lane) & 1u) {
                const int n = p.nsa[idx * p.n_entries + a];
                const double pr = p.psa[idx * p.n_entries + a];
                double u;
                if (n > 0)
                    u = __dadd_rn(p.qsa[idx * p.n_entries + a],
                                  __ddiv_rn(__dmul_rn(__dmul_rn(p.cpuct, pr), sqrt_ns),
                                            (double)(1 + n)));
                else
                    u = __dadd_rn(__dmul_rn(__dmul_rn(p.cpuct, pr), sqrt_ns_eps), 100.0);
                if (u > best_u) { best_u = u; best_a = a; }
            }
        }
        double umax = best_u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) umax = fmax(umax, shfl_xor_f64(umax, o));
        int a_sel = (best_u == umax) ? best_a : 0x7FFFFFFF;
        a_sel = __reduce_min_sync(FULLMASK, a_sel);
        if (a_sel == 0x7FFFFFFF) { status = -1; break; }    // no valid action: cannot happen
        // ---- getNextState: Board.add_tile(a) (BinPackLogic.py:46-69) ----------------
        int r, c, run;
        find_lfb(s, r, c, run);
        uint32_t mine = s.ent[0];
#pragma unroll
        for (int j = 1; j < EJ; ++j) mine = ((a_sel >> 5) == j) ? s.ent[j] : mine;
        const uint32_t v = __shfl_sync(FULLMASK, mine, a_sel & 31);
        place(s, r, c, (int)(v & 0xFFu), (int)(v >> 8));
        kill_pair(s, v);
        compact_pad(s, cbuf[warp]);
        if (lane == 0) {
            p.path_node[(size_t)tree * p.max_path + len] = node;
            p.path_action[(size_t)tree * p.max_path + len] = a_sel;
        }
        len += 1;
        node = find_or_insert(p, tree, s);
    }
    if (lane == 0) {
        p.path_len[tree] = len;
        p.leaf[tree] = leaf;
        p.leaf_value[tree] = value;
        out_status[tree] = status;
        if (status < 0) *p.overflow = 1;          // reported by bp_tree_select / bp_tree_check
    }
}

// priors [B][E] (already masked and normalised on the host, :281-292) and values [B]
__global__ void __launch_bounds__(TREE_WARPS * 32)
tree_backup_kernel(TreeParams p, const double* __restrict__ priors, const double* __restrict__ values,
                   double* __restrict__ out_value)
{
    const int tree = blockIdx.x * TREE_WARPS + (threadIdx.x >> 5);
    if (tree >= p.n_trees) return;
    const int lane = lane_id();
    const int leaf = p.leaf[tree];
    double v = p.leaf_value[tree];
    if (leaf >= 0) {
        const size_t idx = (size_t)tree * p.cap + leaf;
        for (int a = lane; a < p.n_entries; a += 32)
            p.psa[idx * p.n_entries + a] = priors[(size_t)tree * p.n_entries + a];
        if (lane == 0) {
            p.flags[idx] |= 2;
            p.ns[idx] = 0;
        }
        v = values[tree];
    }
    if (lane == 0) {
        const int len = p.path_len[tree];
        for (int i = len - 1; i >= 0; --i) {                 // (:326-334)
            const int node = p.path_node[(size_t)tree * p.max_path + i];
            const int a = p.path_action[(size_t)tree * p.max_path + i];
            const size_t e = ((size_t)tree * p.cap + node) * p.n_entries + a;
            const int n = p.nsa[e];
            if (n > 0)
                p.qsa[e] = __ddiv_rn(__dadd_rn(__dmul_rn((double)n, p.qsa[e]), v), (double)(n + 1));
            else
                p.qsa[e] = v;
            p.nsa[e] = n + 1;
            p.ns[(size_t)tree * p.cap + node] += 1;
        }
        out_value[tree] = v;
    }
}

// Device-resident leaf path: raw float32 policy [B][E] and value [B] straight from the net's
// output tensors; masking with the valid moves, normalisation (float64) and the uniform
// fallback of MCTS.py:281-292 happen here, then the same backup as tree_backup_kernel.
__global__ void __launch_bounds__(TREE_WARPS * 32)
tree_backup_dev_kernel(TreeParams p, const float* __restrict__ policy, const float* __restrict__ values)
{
    const int tree = blockIdx.x * TREE_WARPS + (threadIdx.x >> 5);
    if (tree >= p.n_trees) return;
    const int lane = lane_id();
    const int leaf = p.leaf[tree];
    double v = p.leaf_value[tree];
    if (leaf >= 0) {
        const size_t idx = (size_t)tree * p.cap + leaf;
        const uint4 vm = p.valid[idx];
        const uint32_t vmask[4] = {vm.x, vm.y, vm.z, vm.w};
        double mine[4] = {0.0, 0.0, 0.0, 0.0};
        double total = 0.0;
        int n_valid = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int a = lane + 32 * j;
            if (a < p.n_entries && ((vmask[j] >> lane) & 1u)) {
                mine[j] = (double)policy[(size_t)tree * p.n_entries + a];
                total += mine[j];
                n_valid += 1;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            total += __shfl_xor_sync(FULLMASK, total, o);
            n_valid += __shfl_xor_sync(FULLMASK, n_valid, o);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int a = lane + 32 * j;
            if (a < p.n_entries) {
                const bool ok = (vmask[j] >> lane) & 1u;
                double pr = 0.0;
                if (ok) pr = total > 0.0 ? mine[j] / total : 1.0 / (double)n_valid;
                p.psa[idx * p.n_entries + a] = pr;
            }
        }
        if (lane == 0) {
            p.flags[idx] |= 2;
            p.ns[idx] = 0;
        }
        v = (double)values[tree];
    }
    if (lane == 0) {
        const int len = p.path_len[tree];
        for (int i = len - 1; i >= 0; --i) {
            const int node = p.path_node[(size_t)tree * p.max_path + i];
            const int a = p.path_action[(size_t)tree * p.max_path + i];
            const size_t e = ((size_t)tree * p.cap + node) * p.n_entries + a;
            const int n = p.nsa[e];
            if (n > 0)
                p.qsa[e] = __ddiv_rn(__dadd_rn(__dmul_rn((double)n, p.qsa[e]), v), (double)(n + 1));
            else
                p.qsa[e] = v;
            p.nsa[e] = n + 1;
            p.ns[(size_t)tree * p.cap + node] += 1;
        }
    }
}

__global__ void tree_root_counts_kernel(TreeParams p, int32_t* __restrict__ counts)
{
    const int tree = blockIdx.x;
    const int root = p.root[tree];
    for (int a = threadIdx.x; a < p.n_entries; a += blockDim.x)
        counts[(size_t)tree * p.n_entries + a] =
            root >= 0 ? p.nsa[((size_t)tree * p.cap + root) * p.n_entries + a] : 0;
}

// root statistics for the root-parallel merge: visit counts and value sums W = Q * N per action
__global__ void tree_root_stats_kernel(TreeParams p, int32_t* __restrict__ counts, double* __restrict__ wsa)
{
    const int tree = blockIdx.x;
    const int root = p.root[tree];
    for (int a = threadIdx.x; a < p.n_entries; a += blockDim.x) {
        const size_t e = ((size_t)tree * p.cap + (root >= 0 ? root : 0)) * p.n_entries + a;
        const int n = root >= 0 ? p.nsa[e] : 0;
        counts[(size_t)tree * p.n_entries + a] = n;
        wsa[(size_t)tree * p.n_entries + a] = n > 0 ? __dmul_rn(p.qsa[e], (double)n) : 0.0;
    }
}

}  // namespace bp

// ================================================================== host side
using namespace bp;

struct bp_tree {
    bp_context* ctx = nullptr;
    TreeParams p{};
    int cls = 0;
    std::vector<void*> allocs;
    // staging buffers on the device for the per-simulation exchange
    uint32_t* d_leaf_occ = nullptr;
    uint8_t* d_leaf_tiles = nullptr;
    uint8_t* d_leaf_valid = nullptr;
    int32_t* d_status = nullptr;
    double* d_priors = nullptr;
    double* d_values = nullptr;
    double* d_out_value = nullptr;
    int32_t* d_counts = nullptr;
    double* d_wsa = nullptr;
    uint32_t* d_root_occ = nullptr;
    uint8_t* d_root_tiles = nullptr;
};

template <typename T>
static int tree_alloc(bp_tree* t, T** out, size_t count)
{
    void* ptr = nullptr;
    BP_CUDA(cudaMalloc(&ptr, std::max<size_t>(count, 1) * sizeof(T)));
    t->allocs.push_back(ptr);
    *out = static_cast<T*>(ptr);
    return BP_OK;
}

extern "C" {

int bp_tree_destroy(bp_tree* t)
{
    if (!t) return BP_OK;
    cudaSetDevice(t->ctx->device);
    cudaStreamSynchronize(t->ctx->stream);
    for (void* ptr : t->allocs) cudaFree(ptr);
    delete t;
    return BP_OK;
}

int bp_tree_reset(bp_tree* t)
{
    BP_REQUIRE(t != nullptr, "tree is null");
    BP_CUDA(cudaSetDevice(t->ctx->device));
    cudaStream_t st = t->ctx->stream;
    const TreeParams& p = t->p;
    BP_CUDA(cudaMemsetAsync(p.table, 0xFF, (size_t)p.n_trees * p.hcap * sizeof(int), st));
    BP_CUDA(cudaMemsetAsync(p.n_nodes, 0, (size_t)p.n_trees * sizeof(int), st));
    BP_CUDA(cudaMemsetAsync(p.root, 0xFF, (size_t)p.n_trees * sizeof(int), st));
    BP_CUDA(cudaMemsetAsync(p.overflow, 0, sizeof(int), st));
    return BP_OK;
}

int bp_tree_create(bp_context* ctx, int n_trees, int rows, int cols, int n_entries, int max_nodes,
                   double cpuct, bp_tree** out)
{
    BP_REQUIRE(ctx && out, "null argument");
    BP_REQUIRE(n_trees >= 1 && max_nodes >= 2, "n_trees >= 1, max_nodes >= 2");
    BP_REQUIRE(shape_ok(rows, cols, n_entries) && n_entries >= 2 && n_entries % 2 == 0,
               "rows/cols 1..128, entries even 2..128");
    BP_CUDA(cudaSetDevice(ctx->device));
    bp_tree* t = new bp_tree();
    t->ctx = ctx;
    TreeParams& p = t->p;
    p.n_trees = n_trees; p.rows = rows; p.cols = cols; p.wpr = (cols + 31) / 32;
    p.n_entries = n_entries; p.cap = max_nodes; p.max_path = n_entries / 2 + 1;
    p.cpuct = cpuct;
    int hcap = 16;
    while (hcap < 2 * max_nodes) hcap <<= 1;
    p.hcap = hcap;
    t->cls = shape_class(rows, cols, n_entries);
    const size_t B = n_trees, N = (size_t)n_trees * max_nodes, E = n_entries;
    int rc = BP_OK;
#define TRY(x) do { rc = (x); if (rc) { bp_tree_destroy(t); return rc; } } while (0)
    TRY(tree_alloc(t, &p.occ, N * rows * p.wpr));
    TRY(tree_alloc(t, &p.tiles, N * E * 2));
    TRY(tree_alloc(t, &p.nsa, N * E));
    TRY(tree_alloc(t, &p.qsa, N * E));
    TRY(tree_alloc(t, &p.psa, N * E));
    TRY(tree_alloc(t, &p.valid, N));
    TRY(tree_alloc(t, &p.ns, N));
    TRY(tree_alloc(t, &p.es, N));
    TRY(tree_alloc(t, &p.flags, N));
    TRY(tree_alloc(t, &p.hash, N));
    TRY(tree_alloc(t, &p.table, B * hcap));
    TRY(tree_alloc(t, &p.n_nodes, B));
    TRY(tree_alloc(t, &p.root, B));
    TRY(tree_alloc(t, &p.path_node, B * p.max_path));
    TRY(tree_alloc(t, &p.path_action, B * p.max_path));
    TRY(tree_alloc(t, &p.path_len, B));
    TRY(tree_alloc(t, &p.leaf, B));
    TRY(tree_alloc(t, &p.leaf_value, B));
    TRY(tree_alloc(t, &p.overflow, 1));
    TRY(tree_alloc(t, &t->d_leaf_occ, B * rows * p.wpr));
    TRY(tree_alloc(t, &t->d_leaf_tiles, B * E * 2));
    TRY(tree_alloc(t, &t->d_leaf_valid, B * E));
    TRY(tree_alloc(t, &t->d_status, B));
    TRY(tree_alloc(t, &t->d_priors, B * E));
    TRY(tree_alloc(t, &t->d_values, B));
    TRY(tree_alloc(t, &t->d_out_value, B));
    TRY(tree_alloc(t, &t->d_counts, B * E));
    TRY(tree_alloc(t, &t->d_wsa, B * E));
    TRY(tree_alloc(t, &t->d_root_occ, B * rows * p.wpr));
    TRY(tree_alloc(t, &t->d_root_tiles, B * E * 2));
#undef TRY
    rc = bp_tree_reset(t);
    if (rc) { bp_tree_destroy(t); return rc; }
    *out = t;
    return BP_OK;
}

int bp_tree_set_root(bp_tree* t, const uint32_t* occ, const uint8_t* tiles)
{
    BP_REQUIRE(t && occ && tiles, "null argument");
    BP_CUDA(cudaSetDevice(t->ctx->device));
    const TreeParams& p = t->p;
    cudaStream_t st = t->ctx->stream;
    const size_t B = p.n_trees;
    BP_CUDA(cudaMemcpyAsync(t->d_root_occ, occ, B * p.rows * p.wpr * 4, cudaMemcpyHostToDevice, st));
    BP_CUDA(cudaMemcpyAsync(t->d_root_tiles, tiles, B * p.n_entries * 2, cudaMemcpyHostToDevice, st));
    const int grid = (p.n_trees + TREE_WARPS - 1) / TREE_WARPS;
    dispatch_class(t->cls, [&](auto R, auto Wc, auto E) {
        tree_set_root_kernel<decltype(R)::value, decltype(Wc)::value, decltype(E)::value>
            <<<grid, TREE_WARPS * 32, 0, st>>>(p, t->d_root_occ, t->d_root_tiles);
    });
    t->ctx->launches++;
    BP_CUDA(cudaGetLastError());
    return BP_OK;
}

int bp_tree_select(bp_tree* t, uint32_t* out_leaf_occ, uint8_t* out_leaf_tiles, uint8_t* out_valid,
                   int32_t* out_status)
{
    BP_REQUIRE(t && out_leaf_occ && out_leaf_tiles && out_valid && out_status, "null argument");
    BP_CUDA(cudaSetDevice(t->ctx->device));
    const TreeParams& p = t->p;
    cudaStream_t st = t->ctx->stream;
    const size_t B = p.n_trees;
    const int grid = (p.n_trees + TREE_WARPS - 1) / TREE_WARPS;
    dispatch_class(t->cls, [&](auto R, auto Wc, auto E) {
        tree_select_kernel<decltype(R)::value, decltype(Wc)::value, decltype(E)::value>
            <<<grid, TREE_WARPS * 32, 0, st>>>(p, t->d_leaf_occ, t->d_leaf_tiles, t->d_leaf_valid,
                                               t->d_status);
    });
    t->ctx->launches++;
    BP_CUDA(cudaGetLastError());
    BP_CUDA(cudaMemcpyAsync(out_status, t->d_status, B * 4, cudaMemcpyDeviceToHost, st));
    BP_CUDA(cudaMemcpyAsync(out_leaf_occ, t->d_leaf_occ, B * p.rows * p.wpr * 4, cudaMemcpyDeviceToHost, st));
    BP_CUDA(cudaMemcpyAsync(out_leaf_tiles, t->d_leaf_tiles, B * p.n_entries * 2, cudaMemcpyDeviceToHost, st));
    BP_CUDA(cudaMemcpyAsync(out_valid, t->d_leaf_valid, B * p.n_entries, cudaMemcpyDeviceToHost, st));
    int overflow = 0;
    BP_CUDA(cudaMemcpyAsync(&overflow, p.overflow, sizeof(int), cudaMemcpyDeviceToHost, st));
    BP_CUDA(cudaStreamSynchronize(st));
    if (overflow) {
        set_error("bp_tree_select: a tree ran out of nodes (max_nodes = %d)", p.cap);
        return BP_ERR_STATE;
    }
    for (size_t i = 0; i < B; ++i)
        if (out_status[i] < 0) {
            set_error("bp_tree_select: tree %zu has no root or an invalid path", i);
            return BP_ERR_STATE;
        }
    return BP_OK;
}

int bp_tree_backup(bp_tree* t, const double* priors, const double* values, double* out_values)
{
    BP_REQUIRE(t && priors && values, "null argument");
    BP_CUDA(cudaSetDevice(t->ctx->device));
    const TreeParams& p = t->p;
    cudaStream_t st = t->ctx->stream;
    const size_t B = p.n_trees;
    BP_CUDA(cudaMemcpyAsync(t->d_priors, priors, B * p.n_entries * 8, cudaMemcpyHostToDevice, st));
    BP_CUDA(cudaMemcpyAsync(t->d_values, values, B * 8, cudaMemcpyHostToDevice, st));
    const int grid = (p.n_trees + TREE_WARPS - 1) / TREE_WARPS;
    tree_backup_kernel<<<grid, TREE_WARPS * 32, 0, st>>>(p, t->d_priors, t->d_values, t->d_out_value);
    t->ctx->launches++;
    BP_CUDA(cudaGetLastError());
    if (out_values) {
        BP_CUDA(cudaMemcpyAsync(out_values, t->d_out_value, B * 8, cudaMemcpyDeviceToHost, st));
        BP_CUDA(cudaStreamSynchronize(st));
    }
    return BP_OK;
}

int bp_tree_root_counts(bp_tree* t, int32_t* counts, int32_t* n_nodes)
{
    BP_REQUIRE(t && counts, "null argument");
    BP_CUDA(cudaSetDevice(t->ctx->device));
    const TreeParams& p = t->p;
    cudaStream_t st = t->ctx->stream;
    tree_root_counts_kernel<<<p.n_trees, 128, 0, st>>>(p, t->d_counts);
    t->ctx->launches++;
    BP_CUDA(cudaGetLastError());
    BP_CUDA(cudaMemcpyAsync(counts, t->d_counts, (size_t)p.n_trees * p.n_entries * 4,
                            cudaMemcpyDeviceToHost, st));
    if (n_nodes)
        BP_CUDA(cudaMemcpyAsync(n_nodes, p.n_nodes, (size_t)p.n_trees * 4, cudaMemcpyDeviceToHost, st));
    BP_CUDA(cudaStreamSynchronize(st));
    return BP_OK;
}

/* device-resident simulation step: no host copies, no synchronisation */
int bp_tree_select_dev(bp_tree* t, void** leaf_occ_dev, void** leaf_tiles_dev, void** valid_dev,
                       void** status_dev)
{
    BP_REQUIRE(t && leaf_occ_dev && leaf_tiles_dev && valid_dev && status_dev, "null argument");
    BP_CUDA(cudaSetDevice(t->ctx->device));
    const TreeParams& p = t->p;
    const int grid = (p.n_trees + TREE_WARPS - 1) / TREE_WARPS;
    dispatch_class(t->cls, [&](auto R, auto Wc, auto E) {
        tree_select_kernel<decltype(R)::value, decltype(Wc)::value, decltype(E)::value>
            <<<grid, TREE_WARPS * 32, 0, t->ctx->stream>>>(p, t->d_leaf_occ, t->d_leaf_tiles,
                                                            t->d_leaf_valid, t->d_status);
    });
    t->ctx->launches++;
    BP_CUDA(cudaGetLastError());
    *leaf_occ_dev = t->d_leaf_occ;
    *leaf_tiles_dev = t->d_leaf_tiles;
    *valid_dev = t->d_leaf_valid;
    *status_dev = t->d_status;
    return BP_OK;
}

int bp_tree_backup_dev(bp_tree* t, const void* policy_f32_dev, const void* values_f32_dev)
{
    BP_REQUIRE(t && policy_f32_dev && values_f32_dev, "null argument");
    BP_CUDA(cudaSetDevice(t->ctx->device));
    const TreeParams& p = t->p;
    const int grid = (p.n_trees + TREE_WARPS - 1) / TREE_WARPS;
    tree_backup_dev_kernel<<<grid, TREE_WARPS * 32, 0, t->ctx->stream>>>(
        p, static_cast<const float*>(policy_f32_dev), static_cast<const float*>(values_f32_dev));
    t->ctx->launches++;
    BP_CUDA(cudaGetLastError());
    return BP_OK;
}

/* 1 if any tree ran out of nodes or hit an invalid path since the last reset (synchronises) */
int bp_tree_check(bp_tree* t, int* overflow)
{
    BP_REQUIRE(t && overflow, "null argument");
    BP_CUDA(cudaSetDevice(t->ctx->device));
    int flag = 0;
    BP_CUDA(cudaMemcpyAsync(&flag, t->p.overflow, sizeof(int), cudaMemcpyDeviceToHost, t->ctx->stream));
    BP_CUDA(cudaStreamSynchronize(t->ctx->stream));
    *overflow = flag;
    return BP_OK;
}

/* device pointers of the root statistics, for the root-parallel all-reduce */
int bp_tree_root_counts_dev(bp_tree* t, void** counts_dev, int64_t* n_bytes)
{
    BP_REQUIRE(t && counts_dev && n_bytes, "null argument");
    BP_CUDA(cudaSetDevice(t->ctx->device));
    const TreeParams& p = t->p;
    tree_root_counts_kernel<<<p.n_trees, 128, 0, t->ctx->stream>>>(p, t->d_counts);
    t->ctx->launches++;
    BP_CUDA(cudaGetLastError());
    *counts_dev = t->d_counts;
    *n_bytes = (int64_t)p.n_trees * p.n_entries * 4;
    return BP_OK;
}

/* visit counts Nsa (int32[n_trees][n_entries]) and value sums Wsa = Qsa * Nsa
 * (float64[n_trees][n_entries]) of the roots, left on the device */
int bp_tree_root_stats_dev(bp_tree* t, void** counts_dev, void** wsa_dev)
{
    BP_REQUIRE(t && counts_dev && wsa_dev, "null argument");
    BP_CUDA(cudaSetDevice(t->ctx->device));
    const TreeParams& p = t->p;
    tree_root_stats_kernel<<<p.n_trees, 128, 0, t->ctx->stream>>>(p, t->d_counts, t->d_wsa);
    t->ctx->launches++;
    BP_CUDA(cudaGetLastError());
    *counts_dev = t->d_counts;
    *wsa_dev = t->d_wsa;
    return BP_OK;
}

}  // extern "C"
