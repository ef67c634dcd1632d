"""ctypes loader for oracle/rectpack_oracle.c.  TEST INFRASTRUCTURE (oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "librectpack_oracle.so")

PLACED, ALL_TILES_USED, TILE_CANNOT_BE_PLACED, NO_NEXT_POSITION_TILES_UNUSED = 1, 2, 3, 4
MAX_ENTRIES = 256


class RolloutResult(C.Structure):
    _fields_ = [("solved", C.c_int32), ("depth", C.c_int32), ("placements", C.c_int32)]


class SearchResult(C.Structure):
    _fields_ = [("depth", C.c_int32), ("solution_found", C.c_int32), ("score", C.c_int32),
                ("n_order", C.c_int32), ("n_tiles_placed", C.c_int64), ("rollouts", C.c_int64),
                ("rollout_placements", C.c_int64), ("n_children", C.c_int32),
                ("n_path", C.c_int32)]


def build(force=False):
    src = os.path.join(_HERE, "rectpack_oracle.c")
    hdr = os.path.join(_HERE, "rectpack_oracle.h")
    stale = (not os.path.exists(LIB_PATH)
             or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
        u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
        i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
        L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.orc_philox4x32_10.restype = None
        L.orc_find_first.argtypes = [C.c_int32, i32p, C.c_int64]
        L.orc_find_first.restype = C.c_int64
        L.orc_next_lfb.argtypes = [i32p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_next_occupied_col.argtypes = [i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_valid_next_moves.argtypes = [i32p, C.c_int, C.c_int, i32p, C.c_int, C.c_int, u8p]
        L.orc_place.argtypes = [i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int32, C.c_int, C.c_int]
        L.orc_next_turn.argtypes = [i32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int32,
                                    C.c_int, C.c_int]
        L.orc_eliminate_pair.argtypes = [i32p, C.c_int, C.c_int, C.c_int]
        L.orc_valid_action_mask.argtypes = [i32p, C.c_int, C.c_int, i32p, C.c_int, u8p]
        L.orc_rollout.argtypes = [i32p, C.c_int, C.c_int, i32p, C.c_int, C.c_uint32, C.c_uint32,
                                  C.c_uint32, C.c_uint32, C.c_int32, i32p,
                                  C.POINTER(RolloutResult)]
        L.orc_rollout.restype = None
        L.orc_flat_search.argtypes = [C.c_int, C.c_int, i32p, C.c_int, C.c_int, C.c_int,
                                      C.c_uint32, C.c_uint32, C.c_void_p,
                                      C.POINTER(SearchResult), i32p, i64p, C.c_int, i32p]
        L.orc_flat_search_batch.argtypes = [C.c_int, i32p, i32p, i32p, i32p, C.c_int, C.c_int,
                                            C.c_int, C.c_uint32, C.c_void_p,
                                            C.POINTER(SearchResult), C.c_void_p, C.c_int]
        L.orc_board_add_tile.argtypes = [i32p, C.c_int, C.c_int, i32p, C.c_int, C.c_int]
        L.orc_board_win_state.argtypes = [i32p, C.c_int, C.c_int, i32p, C.c_int,
                                          C.POINTER(C.c_double)]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _grid(grid):
    return np.ascontiguousarray(np.asarray(grid), dtype=np.int32)


def _tiles(tiles):
    return np.ascontiguousarray(np.asarray(tiles, dtype=np.int32).reshape(-1, 2))


def philox(ctr, key):
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), out)
    return tuple(int(v) for v in out)


def find_first(needle, haystack):
    h = np.ascontiguousarray(np.asarray(haystack).astype(np.int32).ravel())
    return int(lib().orc_find_first(int(needle), h, h.size))


def next_lfb(grid):
    g = _grid(grid)
    r, c = C.c_int(), C.c_int()
    if not lib().orc_next_lfb(g, g.shape[0], g.shape[1], C.byref(r), C.byref(c)):
        return None
    return (r.value, c.value)


def next_occupied_col(grid, pos, colorful=True):
    g = _grid(grid)
    return int(lib().orc_next_occupied_col(g, g.shape[0], g.shape[1], pos[0], pos[1],
                                           int(colorful)))


def valid_next_moves(grid, tiles, colorful=True):
    g, t = _grid(grid), _tiles(tiles)
    flags = np.zeros(max(len(t), 1), np.uint8)
    k = lib().orc_valid_next_moves(g, g.shape[0], g.shape[1], t, len(t), int(colorful), flags)
    if k < 0:
        raise TypeError("no LFB on a full board")
    return [tuple(int(v) for v in t[i]) for i in range(len(t)) if flags[i]]


def place(tile, pos, val, grid, get_only_success=False, colorful=True):
    """grid must be an int32 C-contiguous array to be modified in place."""
    g = _grid(grid)
    ok = lib().orc_place(g, g.shape[0], g.shape[1], int(tile[0]), int(tile[1]), pos[0], pos[1],
                         int(val), int(get_only_success), int(colorful))
    if not ok:
        return False, None
    return True, (None if get_only_success else g)


def next_turn(grid, n_entries, tile, val=1, get_only_success=False, colorful=True):
    g = _grid(grid).copy()
    v = lib().orc_next_turn(g, g.shape[0], g.shape[1], n_entries, int(tile[0]), int(tile[1]),
                            int(val), int(get_only_success), int(colorful))
    return int(v), g


def eliminate_pair(tiles, tile):
    t = _tiles(tiles).copy()
    n = lib().orc_eliminate_pair(t, len(t), int(tile[0]), int(tile[1]))
    if n < 0:
        raise ValueError("tile or rotated pair not in list")
    return [tuple(int(v) for v in row) for row in t[:n]]


def valid_action_mask(grid, tiles):
    g, t = _grid(grid), _tiles(tiles)
    mask = np.zeros(len(t), np.uint8)
    if lib().orc_valid_action_mask(g, g.shape[0], g.shape[1], t, len(t), mask) < 0:
        raise TypeError("no LFB on a full board")
    return [int(v) for v in mask]


def rollout(grid, tiles, seed, stream, tag, sim, val=1):
    g, t = _grid(grid).copy(), _tiles(tiles).copy()
    moves = np.zeros(2 * max(len(t), 1), np.int32)
    rr = RolloutResult()
    lib().orc_rollout(g, g.shape[0], g.shape[1], t, len(t), seed, stream, tag, sim, val, moves,
                      C.byref(rr))
    n_moves = rr.placements
    return {"solved": bool(rr.solved), "depth": -1 if rr.solved else int(rr.depth),
            "steps": int(rr.depth), "placements": int(rr.placements),
            "moves": moves[:2 * n_moves].reshape(-1, 2).tolist(),
            "final_occupancy": (g != 0).astype(np.uint8)}


def flat_search(tiles, board, n_sim, strategy="max_depth", seed=123, stream_id=0):
    t = _tiles(tiles)
    b = _grid(board)
    rows, cols = b.shape
    n = len(t) // 2
    res = SearchResult()
    order = np.zeros(2 * (n + 2), np.int32)
    cap = (n + 1) * len(t) + 8
    log = np.zeros(4 * cap, np.int64)
    path = np.zeros(max(n, 1), np.int32)
    strat = 0 if strategy == "max_depth" else 1
    rc = lib().orc_flat_search(rows, cols, t, len(t), n_sim, strat, seed, stream_id,
                               b.ctypes.data_as(C.c_void_p), C.byref(res), order, log, cap, path)
    if rc != 0:
        raise ValueError("orc_flat_search failed: %d" % rc)
    children = []
    for row in log.reshape(-1, 4)[:res.n_children]:
        if row[2]:
            children.append([int(row[0]), int(row[1]), "S"])
        else:
            sc = float(row[3]) if strat == 0 else float(row[3]) / n_sim
            children.append([int(row[0]), int(row[1]), sc])
    return {"depth": int(res.depth), "solution_found": bool(res.solution_found),
            "score": int(res.score), "n_tiles_placed": int(res.n_tiles_placed),
            "solution_tiles_order": order[:2 * res.n_order].reshape(-1, 2).tolist(),
            "children": children, "rollouts": int(res.rollouts),
            "rollout_placements": int(res.rollout_placements),
            "path": path[:res.n_path].tolist()}


def flat_search_batch(rows, cols, tiles, n_entries, n_sim, strategy="max_depth", seed=123,
                      streams=None, n_threads=0):
    """tiles: int32 [P, E, 2] padded; returns list of dicts (no children log)."""
    rows = np.ascontiguousarray(rows, np.int32)
    cols = np.ascontiguousarray(cols, np.int32)
    tiles = np.ascontiguousarray(tiles, np.int32)
    n_entries = np.ascontiguousarray(n_entries, np.int32)
    P, E = tiles.shape[0], tiles.shape[1]
    res = (SearchResult * P)()
    stride = 2 * (E // 2 + 2)
    orders = np.zeros((P, stride), np.int32)
    sp = None
    if streams is not None:
        streams = np.ascontiguousarray(streams, np.uint32)
        sp = streams.ctypes.data_as(C.c_void_p)
    strat = 0 if strategy == "max_depth" else 1
    lib().orc_flat_search_batch(P, rows, cols, tiles, n_entries, E, n_sim, strat, seed, sp, res,
                                orders.ctypes.data_as(C.c_void_p), n_threads)
    out = []
    for p in range(P):
        r = res[p]
        out.append({"depth": int(r.depth), "solution_found": bool(r.solution_found),
                    "score": int(r.score), "n_tiles_placed": int(r.n_tiles_placed),
                    "solution_tiles_order": orders[p, :2 * r.n_order].reshape(-1, 2).tolist(),
                    "rollouts": int(r.rollouts),
                    "rollout_placements": int(r.rollout_placements)})
    return out


def board_add_tile(grid, tiles, action):
    g, t = _grid(grid).copy(), _tiles(tiles).copy()
    rc = lib().orc_board_add_tile(g, g.shape[0], g.shape[1], t, len(t), int(action))
    return rc, g, t


def board_win_state(grid, tiles):
    g, t = _grid(grid), _tiles(tiles)
    v = C.c_double()
    ended = lib().orc_board_win_state(g, g.shape[0], g.shape[1], t, len(t), C.byref(v))
    return (v.value if ended else False)


def num_threads():
    return int(lib().orc_num_threads())
