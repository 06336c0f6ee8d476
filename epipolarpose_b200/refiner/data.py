"""Synthetic stand-in for the reference refiner/data.py `Human36M` dataset (the real one reads
the triangulated / ground-truth H36M pose files, which are not available offline): the same
interface -- `__getitem__ -> (inp [45] f32, tar [45] f32)`, `__len__`, `evaluate(preds) -> mean
per-joint error` -- over seeded root-relative 15-joint poses; inputs = targets + noise (the
refiner learns to undo the triangulation noise)."""
import numpy as np
import torch


class SyntheticPoses(torch.utils.data.Dataset):
    def __init__(self, is_train=True, n=None, seed=0, noise=0.05):
        n = n if n is not None else (4096 if is_train else 1024)
        rng = np.random.default_rng(seed + (0 if is_train else 1))
        tar = rng.normal(0.0, 0.3, (n, 15, 3)).astype(np.float32)
        tar[:, 0] = 0.0                                            # root-relative
        self.tar = tar.reshape(n, 45)
        self.inp = (self.tar + rng.normal(0.0, noise, (n, 45))).astype(np.float32)

    def __len__(self):
        return len(self.tar)

    def __getitem__(self, i):
        return torch.from_numpy(self.inp[i]), torch.from_numpy(self.tar[i])

    def evaluate(self, preds):
        """mean per-joint position error of [n, 45] predictions (reference data.py evaluate)."""
        p = np.asarray(preds, dtype=np.float64).reshape(-1, 15, 3)
        t = self.tar[:len(p)].astype(np.float64).reshape(-1, 15, 3)
        return float(np.sqrt(((p - t) ** 2).sum(2)).mean())


Human36M = SyntheticPoses
