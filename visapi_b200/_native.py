"""ctypes binding of librectpack_b200.so (include/rectpack_b200.h) and numpy helpers.

There is no CPU fallback: importing works anywhere, but creating an ``Engine``
without the built library or without a CUDA device raises ``EngineUnavailable``.
"""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BP_LIB") or os.path.join(_HERE, "librectpack_b200.so")

MAX_ROWS = MAX_COLS = MAX_ENTRIES = 128

# bp_verdict
PLACED, ALL_TILES_USED, TILE_CANNOT_BE_PLACED, NO_NEXT_POSITION_TILES_UNUSED, PAIR_MISSING = 1, 2, 3, 4, 5
MAX_DEPTH, AVG_DEPTH = 0, 1


class EngineUnavailable(RuntimeError):
    """The CUDA library is missing or no CUDA device is usable."""


class EngineError(RuntimeError):
    pass


class SearchResult(C.Structure):
    _fields_ = [("depth", C.c_int32), ("solution_found", C.c_int32), ("score", C.c_int32),
                ("n_order", C.c_int32), ("n_tiles_placed", C.c_int64), ("rollouts", C.c_int64),
                ("rollout_placements", C.c_int64), ("rollouts_executed", C.c_int64),
                ("placements_executed", C.c_int64)]


_lib = None
_lib_lock = threading.Lock()

# every symbol include/rectpack_b200.h declares: (name, restype, argtypes)
_vp, _i32, _i64, _u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
SYMBOLS = [
    ("bp_create", C.c_int, [C.c_int, C.POINTER(_vp)]),
    ("bp_destroy", C.c_int, [_vp]),
    ("bp_last_error", C.c_char_p, []),
    ("bp_set_stream", C.c_int, [_vp, _vp]),
    ("bp_synchronize", C.c_int, [_vp]),
    ("bp_device_info", C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(_i64)]),
    ("bp_issue_microbench", C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("bp_find_first", C.c_int, [_vp, _i32, _vp, _i64, C.POINTER(_i64)]),
    ("bp_lfb", C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    ("bp_valid_masks", C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    ("bp_transitions", C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int,
                                 _vp, _vp, _vp, _vp]),
    ("bp_game_ended", C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    ("bp_rollouts", C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _u32, _vp, _vp,
                              _u32, C.c_int, _vp, _vp, _vp]),
    ("bp_search_create", C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int,
                                   C.c_int, _u32, C.POINTER(_vp)]),
    ("bp_search_set_partition", C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("bp_search_reset", C.c_int, [_vp]),
    ("bp_search_set_mode", C.c_int, [_vp, C.c_int]),
    ("bp_search_set_profiling", C.c_int, [_vp, C.c_int]),
    ("bp_search_profile", C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_int)]),
    ("bp_search_run", C.c_int, [_vp]),
    ("bp_search_step_rollouts", C.c_int, [_vp]),
    ("bp_search_step_reduce", C.c_int, [_vp]),
    ("bp_search_step_commit", C.c_int, [_vp, _vp]),
    ("bp_search_stats_buffer", C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_i64)]),
    ("bp_search_max_depth", C.c_int, [_vp, C.POINTER(C.c_int)]),
    ("bp_search_exchange_create", C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    ("bp_search_exchange_connect", C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    ("bp_search_set_resident_blocks", C.c_int, [_vp, C.c_int]),
    ("bp_search_info", C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("bp_search_last_run", C.c_int, [_vp, C.POINTER(C.c_int)]),
    ("bp_search_resident_stats", C.c_int, [_vp, C.POINTER(C.c_double)]),
    ("bp_search_results", C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    ("bp_search_destroy", C.c_int, [_vp]),
    ("bp_flat_search", C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int,
                                 C.c_int, _u32, _vp, _vp, _vp, _vp]),
    ("bp_tree_create", C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                 C.POINTER(_vp)]),
    ("bp_tree_reset", C.c_int, [_vp]),
    ("bp_tree_set_root", C.c_int, [_vp, _vp, _vp]),
    ("bp_tree_select", C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    ("bp_tree_backup", C.c_int, [_vp, _vp, _vp, _vp]),
    ("bp_tree_select_dev", C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                                     C.POINTER(_vp)]),
    ("bp_tree_backup_dev", C.c_int, [_vp, _vp, _vp]),
    ("bp_tree_check", C.c_int, [_vp, C.POINTER(C.c_int)]),
    ("bp_tree_root_counts", C.c_int, [_vp, _vp, _vp]),
    ("bp_tree_root_counts_dev", C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_i64)]),
    ("bp_tree_root_stats_dev", C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    ("bp_tree_destroy", C.c_int, [_vp]),
]


def load_library():
    """dlopen the C ABI and bind every declared symbol (no CUDA call is made)."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise EngineUnavailable(
                    "%s is missing: run `python -m visapi_b200.build` (needs nvcc); "
                    "there is no CPU fallback" % LIB_PATH)
            lib = C.CDLL(LIB_PATH)
            for name, restype, argtypes in SYMBOLS:
                fn = getattr(lib, name)
                fn.restype = restype
                fn.argtypes = argtypes
            _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ packing helpers
def words_per_row(cols):
    return (int(cols) + 31) // 32


def pack_occupancy(grids, colorful=True):
    """[n, rows, cols] cell grids -> uint32 [n, rows, wpr]; bit (c%32) of word c/32 = col c.

    colorful=True: occupied iff non-zero; False: occupied iff == 1
    (engine/solution_checker.py:111-121 of the reference)."""
    g = np.asarray(grids)
    if g.ndim == 2:
        g = g[None]
    occ = (g != 0) if colorful else (g == 1)
    n, rows, cols = occ.shape
    wpr = words_per_row(cols)
    padded = np.zeros((n, rows, wpr * 32), np.uint8)
    padded[:, :, :cols] = occ
    packed = np.packbits(padded, axis=2, bitorder="little")         # [n, rows, wpr*4] bytes
    return np.ascontiguousarray(packed).view("<u4").reshape(n, rows, wpr)


def unpack_occupancy(words, cols):
    """uint32 [n, rows, wpr] -> uint8 [n, rows, cols] (1 = occupied)."""
    w = np.ascontiguousarray(words, dtype="<u4")
    n, rows, wpr = w.shape
    bits = np.unpackbits(w.view(np.uint8).reshape(n, rows, wpr * 4), axis=2, bitorder="little")
    return bits[:, :, :cols]


def pack_tiles(tile_lists, n_entries=None):
    """list of tile lists [(h, w), ...] -> uint8 [n, E, 2], (0, 0) padded."""
    if isinstance(tile_lists, np.ndarray) and tile_lists.ndim == 3:
        arr = tile_lists
        if n_entries is None or n_entries == arr.shape[1]:
            return np.ascontiguousarray(arr, dtype=np.uint8)
        tile_lists = [a for a in arr]
    n = len(tile_lists)
    if n_entries is None:
        n_entries = max((len(t) for t in tile_lists), default=0)
    out = np.zeros((n, n_entries, 2), np.uint8)
    for i, tl in enumerate(tile_lists):
        if len(tl):
            a = np.asarray(tl, dtype=np.int64).reshape(-1, 2)
            if a.min() < 0 or a.max() > 255:
                raise ValueError("tile sides must be in 0..255")
            out[i, :len(a)] = a
    return out


# ------------------------------------------------------------------------ engine
class _LockedLibrary(object):
    """Serialises every call made through one context: the C ABI allows one caller at a time
    per bp_context, and ctypes releases the GIL while a call runs."""

    def __init__(self, lib, lock):
        self._lib, self._lock = lib, lock

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        lock = self._lock

        def locked(*args):
            with lock:
                return fn(*args)
        locked.__name__ = name
        setattr(self, name, locked)
        return locked


class _StreamScope(object):
    def __init__(self, engine, stream):
        self.engine, self.stream, self.previous = engine, stream, None

    def __enter__(self):
        self.previous = self.engine.stream
        if self.previous != self.stream:
            self.engine.set_stream(self.stream)
        return self.engine

    def __exit__(self, *exc):
        if self.previous != self.stream:
            self.engine.set_stream(self.previous)
        return False


class Engine(object):
    """One bp_context: a device plus the stream its kernels run on.  Calls from several
    Python threads are serialised by a per-engine lock."""

    def __init__(self, device=0, stream=None):
        self._lock = threading.RLock()
        self._lib = _LockedLibrary(load_library(), self._lock)
        handle = C.c_void_p()
        rc = self._lib.bp_create(int(device), C.byref(handle))
        if rc != 0:
            msg = self._lib.bp_last_error().decode()
            if rc == -3:
                raise EngineUnavailable(msg)
            raise EngineError(msg)
        self._h = handle
        self.device = int(device)
        self.stream = 0                  # the legacy NULL stream until set_stream
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise EngineError("rectpack_b200 error %d: %s" % (rc, self._lib.bp_last_error().decode()))

    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); the
        caller owns it and keeps it alive while the engine uses it.  Work already queued on the
        previous stream is ordered before anything launched on the new one (bp_set_stream)."""
        self._check(self._lib.bp_set_stream(self._h, C.c_void_p(int(cuda_stream))))
        self.stream = int(cuda_stream)

    def on_stream(self, cuda_stream):
        """Context manager: run the engine's kernels on ``cuda_stream`` inside the block (e.g.
        torch's current stream, so that engine kernels and torch ops interleave in order)."""
        return _StreamScope(self, int(cuda_stream))

    def synchronize(self):
        self._check(self._lib.bp_synchronize(self._h))

    def device_info(self):
        sm, khz, launches = C.c_int(), C.c_int(), C.c_int64()
        self._check(self._lib.bp_device_info(self._h, C.byref(sm), C.byref(khz), C.byref(launches)))
        return {"sm_count": sm.value, "sm_clock_khz": khz.value, "kernel_launches": launches.value}

    def issue_microbench(self):
        """G warp-instructions/s of dependency-free integer streams: (ALU only, ALU + FMA)."""
        a, b = C.c_double(), C.c_double()
        self._check(self._lib.bp_issue_microbench(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def kernel_launches(self):
        return self.device_info()["kernel_launches"]

    # ---- a1 ---------------------------------------------------------------------
    def find_first(self, needle, haystack):
        hay = np.ascontiguousarray(np.asarray(haystack).astype(np.int32).ravel())
        out = C.c_int64()
        self._check(self._lib.bp_find_first(self._h, int(needle), _ptr(hay), hay.size, C.byref(out)))
        return int(out.value)

    # ---- a2-a8 ------------------------------------------------------------------
    @staticmethod
    def _shape(occ, cols):
        occ = np.ascontiguousarray(occ, dtype=np.uint32)
        if occ.ndim == 2:
            occ = occ[None]
        n, rows, wpr = occ.shape
        if wpr != words_per_row(cols):
            raise ValueError("occ has %d words per row, cols=%d needs %d" % (wpr, cols, words_per_row(cols)))
        return occ, n, rows

    def lfb(self, occ, cols):
        occ, n, rows = self._shape(occ, cols)
        out = np.zeros((n, 3), np.int32)
        self._check(self._lib.bp_lfb(self._h, n, rows, cols, _ptr(occ), _ptr(out)))
        return out

    def valid_masks(self, occ, cols, tiles):
        occ, n, rows = self._shape(occ, cols)
        tiles = np.ascontiguousarray(tiles, dtype=np.uint8)
        E = tiles.shape[1]
        mask = np.zeros((n, E), np.uint8)
        lfb = np.zeros((n, 3), np.int32)
        self._check(self._lib.bp_valid_masks(self._h, n, rows, cols, E, _ptr(occ), _ptr(tiles),
                                             _ptr(mask), _ptr(lfb)))
        return mask, lfb

    def transitions(self, occ, cols, tiles, actions, compact=False, want_state=True):
        occ, n, rows = self._shape(occ, cols)
        tiles = np.ascontiguousarray(tiles, dtype=np.uint8)
        E = tiles.shape[1]
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        out_occ = np.zeros_like(occ) if want_state else None
        out_tiles = np.zeros_like(tiles) if want_state else None
        verdict = np.zeros(n, np.int32)
        lfb = np.zeros((n, 3), np.int32)
        self._check(self._lib.bp_transitions(self._h, n, rows, cols, E, _ptr(occ), _ptr(tiles),
                                             _ptr(actions), int(bool(compact)), _ptr(out_occ),
                                             _ptr(out_tiles), _ptr(verdict), _ptr(lfb)))
        return out_occ, out_tiles, verdict, lfb

    def game_ended(self, occ, cols, tiles):
        occ, n, rows = self._shape(occ, cols)
        tiles = np.ascontiguousarray(tiles, dtype=np.uint8)
        out = np.zeros(n, np.float64)
        self._check(self._lib.bp_game_ended(self._h, n, rows, cols, tiles.shape[1], _ptr(occ),
                                            _ptr(tiles), _ptr(out)))
        return out

    # ---- a10 --------------------------------------------------------------------
    def rollouts(self, occ, cols, tiles, n_sims, seed=123, streams=None, tags=None, sim_begin=0,
                 want_moves=False):
        occ, n, rows = self._shape(occ, cols)
        tiles = np.ascontiguousarray(tiles, dtype=np.uint8)
        E = tiles.shape[1]
        streams = None if streams is None else np.ascontiguousarray(streams, dtype=np.uint32)
        tags = None if tags is None else np.ascontiguousarray(tags, dtype=np.uint32)
        steps = np.zeros((n, n_sims), np.uint8)
        solved = np.zeros((n, n_sims), np.uint8)
        moves = np.zeros((n, n_sims, E // 2, 2), np.uint8) if want_moves else None
        self._check(self._lib.bp_rollouts(self._h, n, rows, cols, E, _ptr(occ), _ptr(tiles),
                                          int(seed), _ptr(streams), _ptr(tags), int(sim_begin),
                                          int(n_sims), _ptr(steps), _ptr(solved), _ptr(moves)))
        return steps, solved, moves

    # ---- a11 + a12 --------------------------------------------------------------
    def search(self, rows, cols, tiles, n_entries, n_sim, strategy=MAX_DEPTH, seed=123,
               streams=None, boards=None):
        return Search(self, rows, cols, tiles, n_entries, n_sim, strategy, seed, streams, boards)

    def flat_search(self, rows, cols, tiles, n_entries, n_sim, strategy=MAX_DEPTH, seed=123,
                    streams=None, boards=None, want_scores=False):
        """One-shot bp_flat_search with host buffers; returns a dict of numpy arrays."""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        n_entries = np.ascontiguousarray(n_entries, dtype=np.int32)
        tiles = np.ascontiguousarray(tiles, dtype=np.uint8)
        P, ES = tiles.shape[0], tiles.shape[1]
        streams = None if streams is None else np.ascontiguousarray(streams, dtype=np.uint32)
        boards = None if boards is None else np.ascontiguousarray(boards, dtype=np.uint32)
        res = (SearchResult * P)()
        order = np.zeros((P, ES // 2 + 2, 2), np.uint8)
        path = np.zeros((P, ES // 2), np.int32)
        scores = np.zeros((P, ES // 2, ES), np.int32) if want_scores else None
        self._check(self._lib.bp_flat_search(
            self._h, P, _ptr(rows), _ptr(cols), _ptr(n_entries), _ptr(tiles), ES, _ptr(boards),
            _ptr(streams), int(n_sim), int(strategy), int(seed), C.cast(res, C.c_void_p),
            _ptr(order), _ptr(path), _ptr(scores)))
        return _results_dict(res, order, path, scores)


def _results_dict(res, order, path, scores):
    a = np.frombuffer(res, dtype=np.dtype([
        ("depth", "<i4"), ("solution_found", "<i4"), ("score", "<i4"), ("n_order", "<i4"),
        ("n_tiles_placed", "<i8"), ("rollouts", "<i8"), ("rollout_placements", "<i8"),
        ("rollouts_executed", "<i8"), ("placements_executed", "<i8")])).copy()
    out = {k: a[k] for k in a.dtype.names}
    out["order"] = order
    out["path"] = path
    out["child_scores"] = scores
    return out


class Search(object):
    """A resident flat search (bp_search): upload once, run / step many times."""

    def __init__(self, engine, rows, cols, tiles, n_entries, n_sim, strategy, seed, streams, boards):
        self.engine = engine
        self._lib = engine._lib
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        n_entries = np.ascontiguousarray(n_entries, dtype=np.int32)
        tiles = np.ascontiguousarray(tiles, dtype=np.uint8)
        self.P, self.ES = tiles.shape[0], tiles.shape[1]
        streams = None if streams is None else np.ascontiguousarray(streams, dtype=np.uint32)
        boards = None if boards is None else np.ascontiguousarray(boards, dtype=np.uint32)
        h = C.c_void_p()
        engine._check(self._lib.bp_search_create(
            engine._h, self.P, _ptr(rows), _ptr(cols), _ptr(n_entries), _ptr(tiles), self.ES,
            _ptr(boards), _ptr(streams), int(n_sim), int(strategy), int(seed), C.byref(h)))
        self._h = h
        md = C.c_int()
        engine._check(self._lib.bp_search_max_depth(self._h, C.byref(md)))
        self.max_depth = md.value

    def set_partition(self, rank, n_ranks, sim_begin, n_local):
        self.engine._check(self._lib.bp_search_set_partition(self._h, rank, n_ranks, sim_begin, n_local))

    def reset(self):
        self.engine._check(self._lib.bp_search_reset(self._h))

    def run(self):
        self.engine._check(self._lib.bp_search_run(self._h))

    def set_mode(self, mode):
        """"stepped" (one launch pair per depth and class), "resident" (one launch runs every
        depth), "hybrid" (launch-per-depth while a depth is plenty of work, then the resident
        kernel) or "auto" (the fastest one as measured)."""
        self.engine._check(self._lib.bp_search_set_mode(
            self._h, {"auto": 0, "stepped": 1, "resident": 2, "hybrid": 3}[mode]))

    def last_run_form(self):
        """"stepped" | "resident" | "hybrid": how the last bp_search_run played the search."""
        r = C.c_int()
        self.engine._check(self._lib.bp_search_last_run(self._h, C.byref(r)))
        return ("stepped", "resident", "hybrid", "stepped")[r.value]

    def last_run_skyline(self):
        """True when the last run's launch-per-depth part used the skyline kernels."""
        r = C.c_int()
        self.engine._check(self._lib.bp_search_last_run(self._h, C.byref(r)))
        return r.value in (2, 3)

    def last_run_resident(self):
        return self.last_run_form() == "resident"

    def resident_stats(self):
        """Time split of the warps of the last resident run (see bp_search_resident_stats)."""
        out = (C.c_double * 13)()
        self.engine._check(self._lib.bp_search_resident_stats(self._h, out))
        tot = max(out[0], 1.0)
        return {"warp_cycles": out[0], "wait_frac": out[1] / tot, "play_frac": out[2] / tot,
                "advance_frac": out[3] / tot, "exchange_frac": out[7] / tot,
                "other_frac": (out[0] - out[1] - out[2] - out[3] - out[7]) / tot,
                "units_played": int(out[4]), "advances": int(out[5]), "polls": int(out[6]),
                "units_published": int(out[8]), "warps": int(out[9]),
                "advance_cycles": {"merge": out[10] / max(out[5], 1.0), "commit": out[11] / max(out[5], 1.0),
                                   "publish": out[12] / max(out[5], 1.0)}}

    def set_profiling(self, on=True):
        self.engine._check(self._lib.bp_search_set_profiling(self._h, int(bool(on))))

    def profile(self):
        """(rollout_ms, advance_ms, n_depths) of the last run; synchronises."""
        r, a, n = C.c_double(), C.c_double(), C.c_int()
        self.engine._check(self._lib.bp_search_profile(self._h, C.byref(r), C.byref(a), C.byref(n)))
        return r.value, a.value, n.value

    def step_rollouts(self):
        self.engine._check(self._lib.bp_search_step_rollouts(self._h))

    def step_reduce(self):
        self.engine._check(self._lib.bp_search_step_reduce(self._h))

    def step_commit(self, gathered_ptr=None):
        self.engine._check(self._lib.bp_search_step_commit(
            self._h, C.c_void_p(int(gathered_ptr)) if gathered_ptr else None))

    def info(self):
        lanes, ncls = C.c_int(), C.c_int()
        self.engine._check(self._lib.bp_search_info(self._h, C.byref(lanes), C.byref(ncls)))
        return {"lane_kernel": bool(lanes.value), "n_classes": ncls.value}

    def stats_buffer(self):
        ptr, nbytes = C.c_void_p(), C.c_int64()
        self.engine._check(self._lib.bp_search_stats_buffer(self._h, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    def exchange_create(self):
        """Exchange buffer of a partitioned resident search: (64-byte CUDA IPC handle, device
        pointer).  See bp_search_exchange_create."""
        handle = np.zeros(64, np.uint8)
        ptr = C.c_void_p()
        self.engine._check(self._lib.bp_search_exchange_create(self._h, _ptr(handle), C.byref(ptr)))
        return handle, ptr.value

    def exchange_connect(self, ipc_handles=None, local_ptrs=None):
        """ipc_handles: uint8[n_ranks][64] in rank order (ranks in other processes), or local_ptrs:
        the exchange pointers of the ranks of this process."""
        if ipc_handles is not None:
            handles = np.ascontiguousarray(ipc_handles, dtype=np.uint8)
            self.engine._check(self._lib.bp_search_exchange_connect(self._h, _ptr(handles), None))
        else:
            arr = (C.c_void_p * len(local_ptrs))(*[C.c_void_p(int(v)) for v in local_ptrs])
            self.engine._check(self._lib.bp_search_exchange_connect(self._h, None, arr))

    def set_resident_blocks(self, blocks_per_sm):
        self.engine._check(self._lib.bp_search_set_resident_blocks(self._h, int(blocks_per_sm)))

    def results(self, want_scores=False):
        res = (SearchResult * self.P)()
        order = np.zeros((self.P, self.ES // 2 + 2, 2), np.uint8)
        path = np.zeros((self.P, self.ES // 2), np.int32)
        scores = np.zeros((self.P, self.ES // 2, self.ES), np.int32) if want_scores else None
        self.engine._check(self._lib.bp_search_results(self._h, C.cast(res, C.c_void_p), _ptr(order),
                                                       _ptr(path), _ptr(scores)))
        return _results_dict(res, order, path, scores)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bp_search_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Tree(object):
    """n_trees PUCT trees of one shape (bp_tree); see include/rectpack_b200.h."""

    def __init__(self, engine, n_trees, rows, cols, n_entries, max_nodes, cpuct):
        self.engine = engine
        self._lib = engine._lib
        self.B, self.rows, self.cols, self.E = int(n_trees), int(rows), int(cols), int(n_entries)
        self.wpr = words_per_row(cols)
        h = C.c_void_p()
        engine._check(self._lib.bp_tree_create(engine._h, self.B, self.rows, self.cols, self.E,
                                               int(max_nodes), float(cpuct), C.byref(h)))
        self._h = h
        self._occ = np.zeros((self.B, self.rows, self.wpr), np.uint32)
        self._tiles = np.zeros((self.B, self.E, 2), np.uint8)
        self._valid = np.zeros((self.B, self.E), np.uint8)
        self._status = np.zeros(self.B, np.int32)

    def reset(self):
        self.engine._check(self._lib.bp_tree_reset(self._h))

    def set_root(self, occ, tiles):
        occ = np.ascontiguousarray(occ, dtype=np.uint32).reshape(self.B, self.rows, self.wpr)
        tiles = np.ascontiguousarray(tiles, dtype=np.uint8).reshape(self.B, self.E, 2)
        self.engine._check(self._lib.bp_tree_set_root(self._h, _ptr(occ), _ptr(tiles)))
        self.engine.synchronize()      # occ/tiles are the caller's buffers

    def select(self):
        """-> (status [B], leaf occupancy [B, rows, wpr], leaf tiles [B, E, 2], valid [B, E])."""
        self.engine._check(self._lib.bp_tree_select(self._h, _ptr(self._occ), _ptr(self._tiles),
                                                    _ptr(self._valid), _ptr(self._status)))
        return self._status, self._occ, self._tiles, self._valid

    def backup(self, priors, values, want_values=True):
        priors = np.ascontiguousarray(priors, dtype=np.float64).reshape(self.B, self.E)
        values = np.ascontiguousarray(values, dtype=np.float64).reshape(self.B)
        out = np.zeros(self.B, np.float64) if want_values else None
        self.engine._check(self._lib.bp_tree_backup(self._h, _ptr(priors), _ptr(values), _ptr(out)))
        if out is None:
            self.engine.synchronize()
        return out

    def select_dev(self):
        """Device pointers (leaf occupancy, leaf tiles, valid masks, status); no sync."""
        a, b, c, d = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.engine._check(self._lib.bp_tree_select_dev(self._h, C.byref(a), C.byref(b), C.byref(c),
                                                        C.byref(d)))
        return a.value, b.value, c.value, d.value

    def backup_dev(self, policy_ptr, values_ptr):
        self.engine._check(self._lib.bp_tree_backup_dev(self._h, C.c_void_p(int(policy_ptr)),
                                                        C.c_void_p(int(values_ptr))))

    def check(self):
        flag = C.c_int()
        self.engine._check(self._lib.bp_tree_check(self._h, C.byref(flag)))
        if flag.value:
            raise EngineError("a PUCT tree ran out of nodes (raise maxNodes)")

    def root_counts(self):
        counts = np.zeros((self.B, self.E), np.int32)
        n_nodes = np.zeros(self.B, np.int32)
        self.engine._check(self._lib.bp_tree_root_counts(self._h, _ptr(counts), _ptr(n_nodes)))
        return counts, n_nodes

    def root_stats_dev(self):
        """(counts pointer int32[B][E], Wsa pointer float64[B][E]) on the device."""
        c, w = C.c_void_p(), C.c_void_p()
        self.engine._check(self._lib.bp_tree_root_stats_dev(self._h, C.byref(c), C.byref(w)))
        return c.value, w.value

    def root_counts_dev(self):
        ptr, nbytes = C.c_void_p(), C.c_int64()
        self.engine._check(self._lib.bp_tree_root_counts_dev(self._h, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bp_tree_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
