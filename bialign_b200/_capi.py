"""ctypes binding of include/bialign_b200.h.  There is no fallback: a missing or unloadable
library, or a machine without an sm_100 GPU, raises.  ctypes drops the GIL around every call."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbialign_b200.so")

BA_OK, BA_ERR_INVALID_ARG, BA_ERR_NO_DEVICE, BA_ERR_CUDA, BA_ERR_OOM, BA_ERR_SCORE_RANGE, BA_ERR_ALPHABET, BA_ERR_STATE = range(8)

SYMBOLS = ["ba_engine_create", "ba_engine_destroy", "ba_last_error", "ba_set_scoring", "ba_load_sequences",
           "ba_load_pairs", "ba_run", "ba_fetch_scores", "ba_trace_bytes", "ba_fetch_traces", "ba_align_batch",
           "ba_get_stats", "ba_set_option", "ba_debug_fetch_codes", "ba_debug_fetch_end_values", "ba_microbench_int",
           "ba_version", "ba_engine_create_multi", "ba_engine_device_count", "ba_set_pair_mu2"]


ENGINE_OPTIONS = {"kernel": -1, "pad": -1, "long": -1, "io_warp": -1, "col_chunks": 0, "p16": -1, "na_kernel": -1, "chain": -1, "rebase": -1, "rebase_window": 0, "warps_per_cta": 0, "code_arena_bytes": 0}


class BaStats(ctypes.Structure):
    _fields_ = [("pairs", ctypes.c_int64), ("cell_states", ctypes.c_int64), ("kernel_launches", ctypes.c_int64),
                ("waves", ctypes.c_int64), ("fill_ms", ctypes.c_double), ("traceback_ms", ctypes.c_double),
                ("total_ms", ctypes.c_double), ("code_bytes", ctypes.c_int64), ("kernel_kind", ctypes.c_int32),
                ("device", ctypes.c_int32), ("warps_per_cta", ctypes.c_int32), ("fallback_pairs", ctypes.c_int32)]


class BialignError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"bialign_b200 error {code}: {text}")
        self.code = code


_lib = None


def load_library():
    """Load the in-tree CUDA library; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m bialign_b200.build` "
                          "(bialign_b200 has no CPU implementation)")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    L.ba_engine_create.argtypes = [i32, ctypes.POINTER(vp)]
    L.ba_engine_create_multi.argtypes = [vp, i32, ctypes.POINTER(vp)]
    L.ba_engine_device_count.argtypes = [vp]
    L.ba_engine_destroy.argtypes = [vp]
    L.ba_engine_destroy.restype = None
    L.ba_last_error.argtypes = [vp]
    L.ba_last_error.restype = ctypes.c_char_p
    L.ba_set_scoring.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32]
    L.ba_load_sequences.argtypes = [vp, vp, vp, vp, i64]
    L.ba_load_pairs.argtypes = [vp, vp, vp, i64]
    L.ba_set_pair_mu2.argtypes = [vp, vp, vp]
    L.ba_run.argtypes = [vp, i32]
    L.ba_fetch_scores.argtypes = [vp, vp]
    L.ba_trace_bytes.argtypes = [vp, ctypes.POINTER(i64)]
    L.ba_fetch_traces.argtypes = [vp, vp, vp, vp]
    L.ba_align_batch.argtypes = [vp, vp, vp, vp, i64, vp, vp, i64, i32, vp]
    L.ba_get_stats.argtypes = [vp, ctypes.POINTER(BaStats)]
    L.ba_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    L.ba_debug_fetch_codes.argtypes = [vp, i64, vp, i64]
    L.ba_debug_fetch_end_values.argtypes = [vp, i64, vp]
    L.ba_version.restype = ctypes.c_char_p
    L.ba_microbench_int.argtypes = [i32, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i32)]
    _lib = L
    return L


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None and a.size else ctypes.c_void_p(a.ctypes.data if a is not None else 0)


class Engine:
    """Thin RAII wrapper of ba_engine: one GPU (`device`), or several GPUs of the box behind one handle
    (`devices` = list of CUDA ordinals, or "all").  Calls on one engine must not overlap (one thread at a time)."""

    def __init__(self, device=0, devices=None):
        self._L = load_library()
        h = ctypes.c_void_p()
        if devices is not None:
            ids = None if devices == "all" else np.ascontiguousarray(list(devices), dtype=np.int32)
            rc = self._L.ba_engine_create_multi(_ptr(ids) if ids is not None else None, 0 if ids is None else ids.size,
                                                ctypes.byref(h))
        else:
            rc = self._L.ba_engine_create(int(device), ctypes.byref(h))
        if rc:
            raise BialignError(rc, self._L.ba_last_error(None).decode())
        self._h = h
        self.n_devices = int(self._L.ba_engine_device_count(h))
        self.device = int(device) if devices is None else None

    def close(self):
        if getattr(self, "_h", None):
            self._L.ba_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise BialignError(rc, self._L.ba_last_error(self._h).decode())

    def set_option(self, key, value):
        self._check(self._L.ba_set_option(self._h, key.encode(), int(value)))

    def apply_options(self, options=None):
        """Every tuning option back to automatic, then `options` (dict).  Engines are shared per device, so each
        aligner calls this before it runs: settings of a previous user of the engine never leak into its alignments."""
        options = options or {}
        for key, auto in ENGINE_OPTIONS.items():
            self.set_option(key, options.get(key, auto))

    def set_scoring(self, sim, structure_weight, gap_opening_cost, gap_cost, shift_cost, max_shift):
        sim = np.ascontiguousarray(sim, dtype=np.int32)
        assert sim.ndim == 2 and sim.shape[0] == sim.shape[1]
        self._check(self._L.ba_set_scoring(self._h, _ptr(sim), sim.shape[0], int(structure_weight),
                                           int(gap_opening_cost), int(gap_cost), int(shift_cost), int(max_shift)))

    def load_sequences(self, residues, classes, offsets):
        self._res = np.ascontiguousarray(residues, dtype=np.uint8)
        self._cls = np.ascontiguousarray(classes, dtype=np.uint8)
        self._off = np.ascontiguousarray(offsets, dtype=np.int64)
        self._check(self._L.ba_load_sequences(self._h, _ptr(self._res), _ptr(self._cls), _ptr(self._off),
                                              len(self._off) - 1))

    def load_pairs(self, seq_a, seq_b):
        a = np.ascontiguousarray(seq_a, dtype=np.int32)
        b = np.ascontiguousarray(seq_b, dtype=np.int32)
        assert a.shape == b.shape
        self._npairs = int(a.size)
        self._check(self._L.ba_load_pairs(self._h, _ptr(a), _ptr(b), a.size))

    def set_pair_mu2(self, matrices):
        """Per-pair mu2 matrices (list of int arrays of shape len(A) x len(B), in pair order), or None to drop them."""
        if matrices is None:
            self._check(self._L.ba_set_pair_mu2(self._h, None, None))
            return
        flat = [np.ascontiguousarray(m, dtype=np.int32).reshape(-1) for m in matrices]
        off = np.concatenate([[0], np.cumsum([f.size for f in flat])]).astype(np.int64)
        buf = np.concatenate(flat).astype(np.int32) if flat else np.zeros(1, dtype=np.int32)
        if buf.size == 0:
            buf = np.zeros(1, dtype=np.int32)
        self._check(self._L.ba_set_pair_mu2(self._h, _ptr(buf), _ptr(off)))

    def run(self, want_trace=False):
        self._check(self._L.ba_run(self._h, 1 if want_trace else 0))

    def fetch_scores(self, out=None):
        if out is None:
            out = np.empty(self._npairs, dtype=np.int64)
        self._check(self._L.ba_fetch_scores(self._h, _ptr(out)))
        return out

    def fetch_traces(self):
        """Returns (cols uint8, offsets int64[n+1], complete uint8[n])."""
        total = ctypes.c_int64(0)
        self._check(self._L.ba_trace_bytes(self._h, ctypes.byref(total)))
        cols = np.empty(max(total.value, 1), dtype=np.uint8)
        offsets = np.zeros(self._npairs + 1, dtype=np.int64)
        complete = np.zeros(max(self._npairs, 1), dtype=np.uint8)
        self._check(self._L.ba_fetch_traces(self._h, _ptr(cols), _ptr(offsets), _ptr(complete)))
        return cols[: total.value], offsets, complete[: self._npairs]

    def align_batch(self, residues, classes, offsets, seq_a, seq_b, want_trace=False, scores_out=None):
        """End-to-end call on host buffers (H2D + fill [+ traceback] + D2H of the scores)."""
        res = np.ascontiguousarray(residues, dtype=np.uint8)
        cls = np.ascontiguousarray(classes, dtype=np.uint8)
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        a = np.ascontiguousarray(seq_a, dtype=np.int32)
        b = np.ascontiguousarray(seq_b, dtype=np.int32)
        self._npairs = int(a.size)
        out = scores_out if scores_out is not None else np.empty(a.size, dtype=np.int64)
        self._check(self._L.ba_align_batch(self._h, _ptr(res), _ptr(cls), _ptr(off), len(off) - 1, _ptr(a), _ptr(b),
                                           a.size, 1 if want_trace else 0, _ptr(out)))
        return out

    def stats(self):
        st = BaStats()
        self._check(self._L.ba_get_stats(self._h, ctypes.byref(st)))
        return {f: getattr(st, f) for f, _ in BaStats._fields_}

    def debug_codes(self, pair, words):
        out = np.empty(words, dtype=np.uint64)
        self._check(self._L.ba_debug_fetch_codes(self._h, int(pair), _ptr(out), words))
        return out

    def debug_end_values(self, pair):
        out = np.empty(9, dtype=np.int32)
        self._check(self._L.ba_debug_fetch_end_values(self._h, int(pair), _ptr(out)))
        return out


def microbench_int(device, kind):
    """Measured integer-pipe rate (thread-instructions/s) and SM count of `device`."""
    L = load_library()
    rate, sms = ctypes.c_double(0), ctypes.c_int(0)
    rc = L.ba_microbench_int(int(device), int(kind), ctypes.byref(rate), ctypes.byref(sms))
    if rc:
        raise BialignError(rc, "ba_microbench_int failed")
    return rate.value, sms.value


_engines = {}


def get_engine(device=None, devices=None):
    """Process-wide engine for one `device` (default: BIALIGN_DEVICE / LOCAL_RANK, else 0) or for a set of `devices`
    (list of ordinals, or "all") behind one multi-GPU handle.  Shared: see BatchAligner for how options are scoped."""
    if devices is not None:
        key = "all" if devices == "all" else tuple(int(d) for d in devices)
        if key not in _engines:
            _engines[key] = Engine(devices=devices)
        return _engines[key]
    if device is None:
        device = int(os.environ.get("BIALIGN_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]
