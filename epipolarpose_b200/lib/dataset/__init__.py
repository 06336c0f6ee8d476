"""Dataset registry (scripts/train.py:125 picks `dataset.<DATASET.DATASET>`).
The H36M / MPII loaders of the reference need image data that is not
available offline and are CPU data-loader work outside the hot path
(SURVEY.md section 2.1 #11-12); the synthetic dataset honours the same contract."""
from .synthetic import SyntheticH36M as synthetic_h36m
