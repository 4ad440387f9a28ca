"""CSR stand-in for the `tf.RaggedTensor` of per-image person boxes the reference passes to
`Patcher` / `Masker` (attacker.py:184, attack_detection.py:183): values [N,4] + row_splits [B+1]."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch


class RaggedBoxes:
    def __init__(self, values: torch.Tensor, row_splits: torch.Tensor):
        if values.dim() != 2 or values.shape[1] != 4:
            raise ValueError("values must be [N,4] (ymin,xmin,ymax,xmax)")
        if row_splits.dtype != torch.int32 or row_splits.dim() != 1:
            raise ValueError("row_splits must be int32 [B+1]")
        self.values = values
        self.row_splits = row_splits

    @classmethod
    def from_rows(cls, rows: Sequence, device) -> "RaggedBoxes":
        rows = [np.asarray(r, dtype=np.float32).reshape(-1, 4) for r in rows]
        splits = np.zeros(len(rows) + 1, dtype=np.int32)
        splits[1:] = np.cumsum([len(r) for r in rows])
        vals = np.concatenate(rows, 0) if rows else np.zeros((0, 4), np.float32)
        return cls(torch.from_numpy(vals).to(device), torch.from_numpy(splits).to(device))

    def nrows(self) -> int:
        return self.row_splits.numel() - 1

    def to_rows(self) -> List[np.ndarray]:
        s = self.row_splits.cpu().numpy()
        v = self.values.cpu().numpy()
        return [v[s[i]:s[i + 1]] for i in range(len(s) - 1)]

    def slice_rows(self, first: int, last: int) -> "RaggedBoxes":
        """Rows [first,last) -- used to shard a global batch over ranks (needs the host splits)."""
        s = self.row_splits.cpu().numpy()
        vals = self.values[int(s[first]):int(s[last])]
        return RaggedBoxes(vals, torch.from_numpy((s[first:last + 1] - s[first]).astype(np.int32)).to(self.values.device))
