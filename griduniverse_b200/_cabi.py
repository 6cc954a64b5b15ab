"""ctypes binding of libgu_b200.so (the C ABI declared in include/gu_b200.h).

There is no CPU fallback: if the library has not been built, or a call returns a
non-zero code, this module raises.  Only raw device pointers, sizes and a stream
handle cross the boundary; PyTorch owns every buffer.
"""
import ctypes
import os

from . import build as _build

c_i32p = ctypes.c_void_p
c_ptr = ctypes.c_void_p

GU_FLAG_AUTO_RESET = 1
GU_FLAG_NO_CARE_TERMINAL = 2
GU_FLAG_ACCUMULATE = 4
GU_FLAG_PACKED_ACTIONS = 8
GU_BFS_LAVA_BLOCKS = 1
GU_TEXT_NO_START, GU_TEXT_NO_GOAL = -1, -2
GU_POLICY_PROBS, GU_POLICY_MASK, GU_POLICY_UNIFORM, GU_POLICY_GREEDY = 0, 1, 2, 3

EXPORTS = ("gu_step", "gu_rollout", "gu_pack_actions", "gu_pack_actions_host", "gu_rollout_policy", "gu_mc_episode_f64", "gu_mc_evaluate_f64", "gu_mc_finalize_f64", "gu_synth_env_levels", "gu_synth_maze", "gu_tables_bytes", "gu_pack_tables", "gu_look_step_ahead", "gu_look_server_start",
           "gu_sweep_f64", "gu_sweep_f32", "gu_greedy_f64", "gu_greedy_f32", "gu_pack_info", "gu_sweep_peer_f32", "gu_sweep_peer_f64", "gu_peer_wait", "gu_max_diff_f32", "gu_max_diff_f64", "gu_vi_small_f64",
           "gu_vi_small_max_cells", "gu_pi_small_f64", "gu_pi_small_max_cells", "gu_vi_batch_f64", "gu_pi_batch_f64", "gu_bfs_init", "gu_bfs_expand", "gu_bfs_walk", "gu_pack_level_text", "gu_render_ansi", "gu_render_rgb", "gu_version", "gu_arch", "gu_error_string")


class GuLevels(ctypes.Structure):
    """struct gu_levels (include/gu_b200.h)."""
    _fields_ = [("X", ctypes.c_int32), ("Y", ctypes.c_int32), ("per_env", ctypes.c_int32),
                ("words", ctypes.c_int32), ("wall", c_ptr), ("goal", c_ptr), ("lava", c_ptr),
                ("start", c_ptr)]


class GuGrid(ctypes.Structure):
    """struct gu_grid (include/gu_b200.h)."""
    _fields_ = [("X", ctypes.c_int32), ("Y", ctypes.c_int32), ("row_begin", ctypes.c_int32),
                ("row_end", ctypes.c_int32), ("pitch", ctypes.c_int32), ("pitch_words", ctypes.c_int32),
                ("wall", c_ptr), ("goal", c_ptr), ("lava", c_ptr), ("info", c_ptr)]


class GuGridBatch(ctypes.Structure):
    """struct gu_grid_batch (include/gu_b200.h)."""
    _fields_ = [("X", ctypes.c_int32), ("Y", ctypes.c_int32), ("n_mazes", ctypes.c_int32),
                ("pitch", ctypes.c_int32), ("pitch_words", ctypes.c_int32), ("cell_stride", ctypes.c_int64),
                ("plane_stride", ctypes.c_int64), ("wall", c_ptr), ("goal", c_ptr), ("lava", c_ptr)]


GU_MAX_PEERS = 16


class GuPeerLinks(ctypes.Structure):
    """struct gu_peer_links (include/gu_b200.h)."""
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("slot", ctypes.c_int32),
                ("n_slots", ctypes.c_int32), ("up_ghost", c_ptr), ("down_ghost", c_ptr),
                ("res_tables", c_ptr * GU_MAX_PEERS), ("done_counter", c_ptr), ("error_flag", c_ptr),
                ("threshold", ctypes.c_double), ("gate_lag", ctypes.c_int32), ("first_slot", ctypes.c_int32),
                ("halo_flags", c_ptr), ("up_flag", c_ptr), ("down_flag", c_ptr), ("edge_counters", c_ptr),
                ("stop_flag", c_ptr), ("abort_flags", c_ptr * GU_MAX_PEERS), ("timeout_cycles", ctypes.c_int64),
                ("slot_base", c_ptr)]


class GuError(RuntimeError):
    def __init__(self, fn, code, msg):
        RuntimeError.__init__(self, "%s failed with code %d: %s" % (fn, code, msg))
        self.code = code


_LIB = None


def library_path():
    # GU_B200_LIB selects a tuning variant built by build.build_variant (developer use only)
    return os.environ.get("GU_B200_LIB", _build.LIB_PATH)


def lib():
    """Load (once) and return the shared library; raise loudly if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(
            "libgu_b200.so is not built (%s). Build it with `python -m griduniverse_b200.build` "
            "or `python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback." % path)
    L = ctypes.CDLL(path)
    i32, i64, u32, f32, f64, p = (ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_float,
                                  ctypes.c_double, c_ptr)
    lvp, gp, gbp = ctypes.POINTER(GuLevels), ctypes.POINTER(GuGrid), ctypes.POINTER(GuGridBatch)
    sig = {
        "gu_step": (ctypes.c_int, [lvp, i64, p, p, p, p, p, p, p, u32, p]),
        "gu_rollout": (ctypes.c_int, [lvp, i64, i64, p, p, p, p, p, p, p, p, p, p, u32, p]),
        "gu_pack_actions": (ctypes.c_int, [p, i64, i64, p, p]),
        "gu_pack_actions_host": (ctypes.c_int, [p, i64, i64, p, i32]),
        "gu_rollout_policy": (ctypes.c_int, [lvp, i64, i64, p, p, p, p, p, p, p, p]),
        "gu_mc_episode_f64": (ctypes.c_int, [i32, i32, p, p, p, i64, p, p, i32, i32, f64, p, p, p, p, p]),
        "gu_mc_finalize_f64": (ctypes.c_int, [i32, p, p, p, p]),
        "gu_mc_evaluate_f64": (ctypes.c_int, [lvp, p, p, i64, p, i32, i32, p, p, i32, i32, f64, p, p, p, p, p, p, p, p,
                                              p, p]),
        "gu_synth_env_levels": (ctypes.c_int, [i32, i32, i64, i64, u32, p, p, p, p, p]),
        "gu_synth_maze": (ctypes.c_int, [i32, i32, i32, i32, i32, u32, p, p, p, p]),
        "gu_tables_bytes": (i64, [lvp, i64]),
        "gu_pack_tables": (ctypes.c_int, [lvp, i64, p, u32, p]),
        "gu_look_step_ahead": (ctypes.c_int, [lvp, i64, p, p, p, p, p, u32, p]),
        "gu_look_server_start": (ctypes.c_int, [lvp, p, u32, i64, i64, p]),
        "gu_sweep_f64": (ctypes.c_int, [gp, p, p, ctypes.c_int, p, f64, p, p, f64, p]),
        "gu_sweep_f32": (ctypes.c_int, [gp, p, p, ctypes.c_int, p, f32, p, p, f32, p]),
        "gu_pack_info": (ctypes.c_int, [gp, p, p]),
        "gu_sweep_peer_f32": (ctypes.c_int, [gp, p, p, ctypes.c_int, p, f32, p, ctypes.POINTER(GuPeerLinks), p]),
        "gu_sweep_peer_f64": (ctypes.c_int, [gp, p, p, ctypes.c_int, p, f64, p, ctypes.POINTER(GuPeerLinks), p]),
        "gu_peer_wait": (ctypes.c_int, [ctypes.POINTER(GuPeerLinks), ctypes.c_int, p]),
        "gu_max_diff_f32": (ctypes.c_int, [gp, p, p, p, p]),
        "gu_max_diff_f64": (ctypes.c_int, [gp, p, p, p, p]),
        "gu_greedy_f64": (ctypes.c_int, [gp, p, p, f64, p]),
        "gu_greedy_f32": (ctypes.c_int, [gp, p, p, f32, p]),
        "gu_vi_small_f64": (ctypes.c_int, [gp, p, p, p, ctypes.c_int, p, f64, f64, i32, p, p, p]),
        "gu_vi_small_max_cells": (i64, []),
        "gu_pi_small_f64": (ctypes.c_int, [gp, p, p, p, ctypes.c_int, p, f64, f64, i32, p, p, p]),
        "gu_pi_small_max_cells": (i64, []),
        "gu_vi_batch_f64": (ctypes.c_int, [gbp, p, p, p, ctypes.c_int, p, f64, f64, i32, p, p, p]),
        "gu_pi_batch_f64": (ctypes.c_int, [gbp, p, p, p, ctypes.c_int, p, f64, f64, i32, p, p, p]),
        "gu_pack_level_text": (ctypes.c_int, [p, i64, i32, i32, p, p, p, p, p, p, p]),
        "gu_render_ansi": (ctypes.c_int, [lvp, i64, p, p, p]),
        "gu_render_rgb": (ctypes.c_int, [lvp, i64, p, p, i32, i32, p, p]),
        "gu_bfs_init": (ctypes.c_int, [gp, p, p, p, p, p, u32, p]),
        "gu_bfs_expand": (ctypes.c_int, [gp, p, p, p, i32, i32, p, u32, p]),
        "gu_bfs_walk": (ctypes.c_int, [gp, p, i64, p, i32, p, p]),
        "gu_version": (ctypes.c_int, []),
        "gu_arch": (ctypes.c_char_p, []),
        "gu_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)   # AttributeError here = the library does not export the header's symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def check(fn_name, code):
    if code != 0:
        raise GuError(fn_name, code, lib().gu_error_string(code).decode())


def on_device(fn):
    """Method decorator: run with ``self.device`` as the current CUDA device.  Every C-ABI call enqueues
    on the CURRENT stream of the CURRENT device, so an object living on cuda:1 must make cuda:1
    current for the duration of the call (otherwise its kernels would be launched on another GPU's
    stream with this GPU's pointers)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        import torch
        dev = getattr(self, "device", None)
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return fn(self, *args, **kwargs)
    return wrapper


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)
