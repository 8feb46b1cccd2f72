"""Update functions: the server-side `initValue / updateValue` of a table.

Reference SPI: services/et/.../evaluator/api/UpdateFunction.java:25-44.  In the B200 design an
update function has up to three faces:
  * a per-key python face (`init_value`, `update_value`) used by generic (dict) blocks,
  * a vectorised numpy face (`init_rows`, `update_rows`) used by dense host blocks,
  * a device code (`device_fn`) selecting the fused epilogue of the PUSH kernel
    (harmony_b200/ops/csrc/table_ops.cu: UPD_ADD / UPD_SGD_CLAMP / UPD_ADAGRAD / UPD_ASSIGN).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional

import numpy as np

# device codes (must match table_ops.cu)
DEV_ADD, DEV_SGD_CLAMP, DEV_ADAGRAD, DEV_ASSIGN = 0, 1, 2, 3
INIT_ZERO, INIT_GAUSSIAN, INIT_UNIFORM, INIT_CONST = 0, 1, 2, 3


class UpdateFunction:
    """Base class. `params` is a plain dict so that the function serialises with the table conf."""

    name = "void"
    device_fn: Optional[int] = None

    def __init__(self, **params: Any) -> None:
        self.params: Dict[str, Any] = dict(params)

    # ---- per key ----
    def init_value(self, key: Any) -> Any:
        return None

    def update_value(self, key: Any, old: Any, delta: Any) -> Any:
        return old

    # ---- vectorised (dense rows) ----
    def init_spec(self) -> Dict[str, Any]:
        """{kind, a, b, seed} consumed by the host and device initialisers."""
        return {"kind": INIT_ZERO, "a": 0.0, "b": 0.0, "seed": 0}

    def init_rows(self, keys: np.ndarray, row_dim: int, dtype) -> np.ndarray:
        spec = self.init_spec()
        n = len(keys)
        if spec["kind"] == INIT_ZERO:
            return np.zeros((n, row_dim), dtype=dtype)
        if spec["kind"] == INIT_CONST:
            return np.full((n, row_dim), spec["a"], dtype=dtype)
        out = np.empty((n, row_dim), dtype=dtype)
        for i, k in enumerate(np.asarray(keys).tolist()):
            rng = np.random.default_rng([int(spec["seed"]) & 0x7fffffff, int(k) & 0x7fffffffffffffff])
            if spec["kind"] == INIT_GAUSSIAN:
                out[i] = (spec["a"] + spec["b"] * rng.standard_normal(row_dim)).astype(dtype)
            else:
                out[i] = rng.uniform(spec["a"], spec["b"], row_dim).astype(dtype)
        return out

    def update_rows(self, old: np.ndarray, deltas: np.ndarray) -> np.ndarray:
        raise NotImplementedError(f"{self.name} has no vectorised update")

    def device_params(self) -> Dict[str, float]:
        return {"step": 0.0, "lo": 0.0, "hi": 0.0, "eps": 0.0}

    def spec(self) -> Dict[str, Any]:
        return {"name": self.name, "params": self.params}


class VoidUpdateFunction(UpdateFunction):
    """Tables that are never `update`d (input tables). Reference: VoidUpdateFunction.java."""
    name = "void"


class AddUpdateFunction(UpdateFunction):
    """w += delta with a configurable initialiser (MLR: N(0, model_gaussian); AddVector: zeros).

    Reference: MLRETModelUpdateFunction.java:51-62, AddVectorET update function.
    """
    name = "add"
    device_fn = DEV_ADD

    def init_spec(self):
        p = self.params
        if p.get("init") == "gaussian":
            return {"kind": INIT_GAUSSIAN, "a": float(p.get("mean", 0.0)), "b": float(p.get("std", 0.001)),
                    "seed": int(p.get("seed", 0))}
        if p.get("init") == "uniform":
            return {"kind": INIT_UNIFORM, "a": float(p.get("init_lo", 0.0)), "b": float(p.get("init_hi", 1.0)),
                    "seed": int(p.get("seed", 0))}
        if p.get("init") == "const":
            return {"kind": INIT_CONST, "a": float(p.get("value", 0.0)), "b": 0.0, "seed": 0}
        return super().init_spec()

    def init_value(self, key):
        dim = self.params.get("row_dim")
        if dim is None:
            return 0
        return self.init_rows(np.asarray([key]), int(dim), np.dtype(self.params.get("dtype", "float32")))[0]

    def update_value(self, key, old, delta):
        c = self.params.get("coefficient")          # AddIntegerUpdateFunction.java:45: old + delta * coefficient
        return old + (delta if c is None else delta * c)

    def update_rows(self, old, deltas):
        c = self.params.get("coefficient")
        return old + (deltas if c is None else deltas * c)


class SGDClampUpdateFunction(AddUpdateFunction):
    """w = clamp(w - step * g, lo, hi); init U(init_min, init_max).

    Reference: NMFETModelUpdateFunction.java:43-52, NMFModelGenerator.java:59-101.
    """
    name = "sgd_clamp"
    device_fn = DEV_SGD_CLAMP

    def update_value(self, key, old, delta):
        return np.clip(old - self.params["step"] * delta, self.params.get("lo", 0.0), self.params.get("hi", 1e6))

    def update_rows(self, old, deltas):
        return np.clip(old - np.float32(self.params["step"]) * deltas, self.params.get("lo", 0.0),
                       self.params.get("hi", 1e6)).astype(old.dtype)

    def device_params(self):
        return {"step": float(self.params["step"]), "lo": float(self.params.get("lo", 0.0)),
                "hi": float(self.params.get("hi", 1e6)), "eps": 0.0}


class AdagradUpdateFunction(AddUpdateFunction):
    """Row = [w | h]; h += g^2; w -= step * g / sqrt(h + eps). (north-star optimiser option)"""
    name = "adagrad"
    device_fn = DEV_ADAGRAD

    def update_rows(self, old, deltas):
        d = old.shape[1] // 2
        w, h = old[:, :d].copy(), old[:, d:].copy()
        g = deltas[:, :d]
        h += g * g
        w -= np.float32(self.params["step"]) * g / np.sqrt(h + np.float32(self.params.get("eps", 1e-8)))
        return np.concatenate([w, h], axis=1).astype(old.dtype)

    def update_value(self, key, old, delta):
        return self.update_rows(np.asarray(old)[None, :], np.asarray(delta)[None, :])[0]

    def device_params(self):
        return {"step": float(self.params["step"]), "lo": 0.0, "hi": 0.0, "eps": float(self.params.get("eps", 1e-8))}


class AssignUpdateFunction(UpdateFunction):
    """update = overwrite (value := delta)."""

    name = "assign"
    device_fn = DEV_ASSIGN

    def update_value(self, key, old, delta):
        return delta

    def update_rows(self, old, deltas):
        return deltas.astype(old.dtype)


class LDASparseUpdateFunction(UpdateFunction):
    """int[K+1] rows; delta = flat (topic, change) pairs; clamp >= 0; last slot = #non-zero topics.

    Reference: LDAETModelUpdateFunction.java:38-64.
    """
    name = "lda_sparse"

    def init_value(self, key):
        return np.zeros(int(self.params["num_topics"]) + 1, dtype=np.int32)

    def init_rows(self, keys, row_dim, dtype):
        return np.zeros((len(keys), row_dim), dtype=dtype)

    def update_value(self, key, old, delta):
        k = int(self.params["num_topics"])
        d = np.asarray(delta, dtype=np.int64).reshape(-1, 2)
        for topic, chg in d.tolist():
            oldc = int(old[topic])
            newc = max(oldc + chg, 0)
            old[topic] = newc
            if oldc == 0 and newc != 0:
                old[k] += 1
            elif oldc != 0 and newc == 0:
                old[k] -= 1
        return old


class AppendUpdateFunction(UpdateFunction):
    """value = list; update appends (GBT trees: GBTETModelUpdateFunction.java:40-48; Pregel messages)."""
    name = "append"

    def init_value(self, key):
        return []

    def update_value(self, key, old, delta):
        if isinstance(delta, list):
            old.extend(delta)
        else:
            old.append(delta)
        return old


class CombineUpdateFunction(UpdateFunction):
    """Pregel message combiner: value = combine(old, delta) with op in {sum, min, max}."""
    name = "combine"

    def init_value(self, key):
        return None

    def update_value(self, key, old, delta):
        if old is None:
            return delta
        op = self.params.get("op", "sum")
        if op == "sum":
            return old + delta
        if op == "min":
            return min(old, delta)
        return max(old, delta)


_REGISTRY: Dict[str, Callable[..., UpdateFunction]] = {}


def register_update_function(cls) -> None:
    _REGISTRY[cls.name] = cls


for _c in (VoidUpdateFunction, AddUpdateFunction, SGDClampUpdateFunction, AdagradUpdateFunction,
           AssignUpdateFunction, LDASparseUpdateFunction, AppendUpdateFunction, CombineUpdateFunction):
    register_update_function(_c)


def make_update_function(spec: Any) -> UpdateFunction:
    """spec: UpdateFunction | name | {"name":..., "params": {...}}"""
    if isinstance(spec, UpdateFunction):
        return spec
    if spec is None:
        return VoidUpdateFunction()
    if isinstance(spec, str):
        return _REGISTRY[spec]()
    return _REGISTRY[spec["name"]](**spec.get("params", {}))
