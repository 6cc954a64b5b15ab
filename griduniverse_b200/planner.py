"""Device-side driver for the Bellman-sweep kernels: sweeps, greedy extraction and the
value-iteration / policy-iteration loops of core/algorithms/dynamic_programming.py.

Value functions, tie masks and policies stay on the GPU in the padded layout of
``PlanGrid``; the host sees one residual scalar per sweep (read in chunks).  The
reference-signature wrappers live in ``griduniverse_b200/algorithms``.
"""
import warnings

import numpy as np
import torch

from . import _cabi
from .device import PlanGrid

_KINDS = {"probs": _cabi.GU_POLICY_PROBS, "mask": _cabi.GU_POLICY_MASK,
          "uniform": _cabi.GU_POLICY_UNIFORM, "greedy": _cabi.GU_POLICY_GREEDY}


def masks_to_policy(masks, dtype=np.float64):
    """Tie masks -> the [N,4] rows utils.py:69-71 writes: 1/len(ties) on ties, 0 elsewhere."""
    masks = np.asarray(masks, dtype=np.uint8)
    bits = ((masks[:, None] >> np.arange(4, dtype=np.uint8)) & 1).astype(dtype)
    cnt = bits.sum(axis=1, keepdims=True)
    inv = np.divide(1.0, cnt, out=np.zeros_like(cnt), where=cnt > 0)
    return bits * inv


def policy_to_masks(policy):
    """Tie masks if ``policy`` is exactly 'uniform on a subset' in every row, else None."""
    policy = np.asarray(policy)
    if policy.ndim != 2 or policy.shape[1] != 4:
        return None
    masks = ((policy > 0) * np.array([1, 2, 4, 8], dtype=np.uint8)).sum(axis=1).astype(np.uint8)
    if np.array_equal(masks_to_policy(masks, np.float64), np.asarray(policy, dtype=np.float64)):
        return masks
    return None


class Planner(object):
    """Sweeps / greedy / VI / PI for one level (or one row shard) on one GPU."""

    def __init__(self, level, dtype=np.float64, device="cuda", row_begin=0, row_end=None, grid=None):
        """``level``: a host Level; or pass ``grid`` = a ready PlanGrid (e.g. synth.maze_plan_grid)."""
        self.grid = PlanGrid(level, dtype, device, row_begin, row_end) if grid is None else grid
        self.level = level
        self.device = self.grid.device
        self.np_dtype = self.grid.np_dtype
        self.dtype = self.grid.dtype
        self._lib = _cabi.lib()
        self._f64 = self.np_dtype == np.dtype(np.float64)
        self._sweep_fn = self._lib.gu_sweep_f64 if self._f64 else self._lib.gu_sweep_f32
        self._greedy_fn = self._lib.gu_greedy_f64 if self._f64 else self._lib.gu_greedy_f32
        self.launches = 0

    # ------------------------------------------------------------------ policy staging
    @_cabi.on_device
    def stage_policy(self, policy):
        """Caller's policy -> (kind, device tensor or None).

        ``"uniform"`` / ``"greedy"`` need no array.  A NumPy [N,4] array that is uniform on a
        subset in every row (policy0 of the examples, any greedy policy) is shipped as 1-byte
        tie masks; anything else as T[N,4] probabilities."""
        if isinstance(policy, tuple):          # already staged: (kind, tensor)
            return policy
        if isinstance(policy, str):
            return _KINDS[policy], None
        if torch.is_tensor(policy):
            if policy.dtype == torch.uint8:
                return _cabi.GU_POLICY_MASK, policy
            return _cabi.GU_POLICY_PROBS, policy
        pol = np.asarray(policy)
        masks = policy_to_masks(pol)
        if masks is not None:
            return _cabi.GU_POLICY_MASK, self.grid.pad(masks, torch.uint8)
        return _cabi.GU_POLICY_PROBS, self.grid.pad(pol.astype(self.np_dtype), self.dtype)

    # ------------------------------------------------------------------ kernels
    @_cabi.on_device
    def sweep(self, v_in, v_out, kind, policy_t, gamma, residual=None, gate=None, threshold=0.0):
        rc = self._sweep_fn(self.grid.ref(), _cabi.ptr(v_in), _cabi.ptr(v_out), kind, _cabi.ptr(policy_t),
                            float(gamma), _cabi.ptr(residual), _cabi.ptr(gate), float(threshold),
                            _cabi.stream_ptr())
        _cabi.check("gu_sweep", rc)
        self.launches += 1

    @_cabi.on_device
    def greedy(self, v, gamma, out=None):
        out = self.grid.empty(torch.uint8) if out is None else out
        rc = self._greedy_fn(self.grid.ref(), _cabi.ptr(v), _cabi.ptr(out), float(gamma), _cabi.stream_ptr())
        _cabi.check("gu_greedy", rc)
        self.launches += 1
        return out

    @_cabi.on_device
    def max_diff(self, a, b):
        """Signed max of (a - b) over the owned cells as a device scalar (dynamic_programming.py:44)."""
        out = self.new_residuals(1)
        fn = self._lib.gu_max_diff_f64 if self._f64 else self._lib.gu_max_diff_f32
        _cabi.check("gu_max_diff", fn(self.grid.ref(), _cabi.ptr(a), _cabi.ptr(b), _cabi.ptr(out), _cabi.stream_ptr()))
        self.launches += 1
        return out

    def new_residuals(self, n):
        return torch.full((n,), float("-inf"), dtype=self.dtype, device=self.device)

    @_cabi.on_device
    def stage_value(self, value_function):
        if value_function is None:
            return self.grid.empty()
        if torch.is_tensor(value_function) and tuple(value_function.shape) == (self.grid.rows + 2, self.grid.pitch):
            return value_function.to(self.dtype).clone()
        return self.grid.pad(value_function)

    # ------------------------------------------------------------------ value iteration
    @_cabi.on_device
    def value_iteration(self, policy="uniform", value_function=None, threshold=1e-5, max_steps=1000,
                        discount_factor=1.0, chunk=16, allow_small=True, use_graph=True):
        """dynamic_programming.py:8-28.  Returns (V_padded, tie_masks_padded, sweeps, last_delta).

        The first sweep evaluates the caller's policy, every later sweep is the fused
        greedy pass (greedy of V_k and evaluation of V_k in one kernel).  Sweeps are enqueued
        `chunk` at a time; the kernels after the converged one are gated off on the device."""
        kind0, pol_t = self.stage_policy(policy)
        v0 = self.stage_value(value_function)
        g = self.grid
        n_cells = g.X * g.Y
        if (allow_small and self._f64 and g.rows == g.Y and n_cells <= self._lib.gu_vi_small_max_cells()
                and max_steps > 0):
            v_out = g.empty()
            tie = g.empty(torch.uint8)
            meta_i = torch.zeros(1, dtype=torch.int32, device=self.device)
            meta_d = torch.zeros(1, dtype=torch.float64, device=self.device)
            rc = self._lib.gu_vi_small_f64(g.ref(), _cabi.ptr(v0), _cabi.ptr(v_out), _cabi.ptr(tie), kind0,
                                           _cabi.ptr(pol_t), float(discount_factor), float(threshold),
                                           int(max_steps), _cabi.ptr(meta_i), _cabi.ptr(meta_d),
                                           _cabi.stream_ptr())
            _cabi.check("gu_vi_small_f64", rc)
            self.launches += 1
            return v_out, tie, int(meta_i.item()), float(meta_d.item())
        from .sharded import ShardedValueIteration      # one driver for one GPU and for row shards
        if getattr(self, "_solo_driver", None) is None:
            self._solo_driver = ShardedValueIteration(self, solo=True)
        v, tie, sweeps, last = self._solo_driver.value_iteration((kind0, pol_t), v0, threshold, max_steps,
                                                                 discount_factor, chunk=chunk, use_graph=use_graph)
        return v.clone(), tie, sweeps, last      # v is a view of the driver's persistent buffers

    # ------------------------------------------------------------------ policy iteration
    @_cabi.on_device
    def policy_iteration(self, policy="uniform", value_function=None, threshold=1e-5, max_steps=1000,
                         discount_factor=1.0, allow_small=True, chunk=16, use_graph=True):
        """dynamic_programming.py:31-57.  Returns (V_lastconv_padded, tie_masks or None, sweeps,
        delta_eval, exhausted): tie masks are None when no greedy update ever ran (the caller's
        policy is returned unchanged in that case, as in the reference)."""
        kind, pol_t = self.stage_policy(policy)
        g = self.grid
        if (allow_small and self._f64 and g.rows == g.Y and g.X * g.Y <= self._lib.gu_pi_small_max_cells()
                and max_steps > 0):
            v0 = self.stage_value(value_function)
            v_out = g.empty()
            tie = g.empty(torch.uint8)
            meta = torch.zeros(3, dtype=torch.int32, device=self.device)
            meta_d = torch.zeros(1, dtype=torch.float64, device=self.device)
            rc = self._lib.gu_pi_small_f64(g.ref(), _cabi.ptr(v0), _cabi.ptr(v_out), _cabi.ptr(tie), kind,
                                           _cabi.ptr(pol_t), float(discount_factor), float(threshold),
                                           int(max_steps), _cabi.ptr(meta), _cabi.ptr(meta_d), _cabi.stream_ptr())
            _cabi.check("gu_pi_small_f64", rc)
            self.launches += 1
            sweeps, improved, exhausted = (int(x) for x in meta.cpu().numpy())
            return v_out, (tie if improved else None), sweeps, float(meta_d.item()), bool(exhausted)
        if getattr(self, "_solo_driver", None) is None:
            from .sharded import ShardedValueIteration
            self._solo_driver = ShardedValueIteration(self, solo=True)
        last, tie, sweeps, delta_eval, exhausted = self._solo_driver.policy_iteration(
            (kind, pol_t), value_function, threshold, max_steps, discount_factor, chunk=chunk, use_graph=use_graph)
        # the driver's buffers are reused by the next solve: hand out copies
        return last.clone(), (None if tie is None else tie.clone()), sweeps, delta_eval, exhausted
