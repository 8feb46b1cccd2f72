"""GPU tree growing / forest prediction for the GBT app (ops/csrc/k12_ops.cu).

Level-wise exact greedy growth: every level is four launches (node statistics, one warp per (feature, node)
split search over the pre-sorted columns, per-node arg-max + tree write, sample routing).  The columns are
argsorted once per mini-batch.  Reference behaviour: GBTTrainer.java:327-531 (split search), :244-320
(residual = label - forest prediction), tree/GBTree.java (array-encoded tree); the CPU twin is
`mlapps/gbt.py:GBTTrainer._build`, which the tests compare against node by node.
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch

CAT_FLAG = 1 << 63


def _i64(mask: int) -> int:
    return mask - (1 << 64) if mask >= (1 << 63) else mask


class DeviceTreeBuilder:
    """Level-wise exact greedy tree growth and forest prediction on the GPU (ops/csrc/k12_ops.cu).  Reference:
    dolphin/mlapps/gbt/GBTTrainer.java:327-531."""

    def __init__(self, device, num_features: int, types: Dict[int, int], max_depth: int, leaf_min: int,
                 lam: float, gamma: float) -> None:
        from ..ops import native
        self._n = native()
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.F, self.max_depth, self.leaf_min = int(num_features), int(max_depth), int(leaf_min)
        self.lam, self.gamma = float(lam), float(gamma)
        if any(v > 63 for v in types.values()):
            raise ValueError("categorical features with more than 63 categories are not supported on the device path")
        self.ftype = torch.tensor([1 if types.get(f, 0) else 0 for f in range(self.F)], dtype=torch.int32,
                                  device=self.device)
        self.nodes = 2 ** (self.max_depth + 1) - 1

    def _check(self, *tensors: torch.Tensor) -> None:
        for t in tensors:
            if t.device != self.device:
                raise ValueError(f"tensor on {t.device}, builder on {self.device}: kernels of one tree must run on "
                                 "one GPU's stream")

    def presort(self, X: torch.Tensor) -> torch.Tensor:
        """X [n, F] float32 on the device -> per-feature stable argsort [F, n] int32."""
        self._check(X)
        return torch.sort(X, dim=0, stable=True).indices.t().contiguous().to(torch.int32)

    def build(self, X: torch.Tensor, order: torch.Tensor, g: torch.Tensor, step: float) -> Dict[str, torch.Tensor]:
        """Grow one tree on residuals g [n] float64.  Returns the tree arrays (device tensors, complete-tree ids)."""
        self._check(X, order, g)
        n = X.shape[0]
        dev = self.device
        t = {"feature": torch.full((self.nodes,), -1, dtype=torch.int32, device=dev),
             "thr": torch.zeros(self.nodes, dtype=torch.float32, device=dev),
             "mask": torch.zeros(self.nodes, dtype=torch.int64, device=dev),
             "value": torch.zeros(self.nodes, dtype=torch.float32, device=dev),
             "leaf": torch.zeros(self.nodes, dtype=torch.int32, device=dev)}
        node_of = torch.zeros(n, dtype=torch.int32, device=dev)
        for d in range(self.max_depth + 1):
            self._n.gbt_level(X, order, self.ftype, node_of, g, 2 ** d - 1, 2 ** d, self.leaf_min, self.lam,
                              self.gamma, float(step), d < self.max_depth, t["feature"], t["thr"], t["mask"],
                              t["value"], t["leaf"])
        return t

    def predict(self, X: torch.Tensor, forest: Dict[str, torch.Tensor]) -> torch.Tensor:
        """forest: stacked [T, nodes] arrays -> float64 [n] sum of the trees' leaf values."""
        self._check(X)
        out = torch.zeros(X.shape[0], dtype=torch.float64, device=self.device)
        if forest["feature"].numel():
            self._n.gbt_predict(X, forest["feature"], forest["thr"], forest["mask"], forest["value"],
                                forest["leaf"], out)
        return out

    # ---- host tree objects <-> device arrays -------------------------------------------------
    def to_host_tree(self, t: Dict[str, torch.Tensor]):
        from ..mlapps.gbt import GBTree
        tree = GBTree(self.max_depth)
        tree.feature[:] = t["feature"].cpu().numpy()
        tree.threshold[:] = t["thr"].cpu().numpy()
        tree.value[:] = t["value"].cpu().numpy()
        tree.is_leaf[:] = t["leaf"].cpu().numpy().astype(bool)
        masks = t["mask"].cpu().numpy().astype(np.uint64)
        for i in np.nonzero(masks >> np.uint64(63))[0]:
            m = int(masks[i])
            tree.categories[int(i)] = frozenset(c for c in range(63) if (m >> c) & 1)
        return tree

    def pack_forest(self, trees: Sequence) -> Dict[str, torch.Tensor]:
        """List of host GBTree objects (possibly of different depths) -> stacked device arrays."""
        T = len(trees)
        nodes = max([self.nodes] + [len(t.feature) for t in trees])
        feat = np.full((T, nodes), -1, dtype=np.int32); thr = np.zeros((T, nodes), dtype=np.float32)
        val = np.zeros((T, nodes), dtype=np.float32); leaf = np.zeros((T, nodes), dtype=np.int32)
        mask = np.zeros((T, nodes), dtype=np.int64)
        for i, t in enumerate(trees):
            k = len(t.feature)
            feat[i, :k], thr[i, :k], val[i, :k], leaf[i, :k] = t.feature, t.threshold, t.value, t.is_leaf
            for node, cats in t.categories.items():
                m = CAT_FLAG
                for c in cats:
                    m |= 1 << int(c)
                mask[i, node] = _i64(m)
        dev = self.device
        return {"feature": torch.from_numpy(feat).to(dev), "thr": torch.from_numpy(thr).to(dev),
                "mask": torch.from_numpy(mask).to(dev), "value": torch.from_numpy(val).to(dev),
                "leaf": torch.from_numpy(leaf).to(dev)}
