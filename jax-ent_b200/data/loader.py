"""Dataset / ExpD_Dataloader containers (reference jaxent/src/data/loader.py:24-308) and
SparseFragmentMapping (data/splitting/mapping.py:24-32), reduced to what the optimise step reads:
train/val `y_true`, the peptide<-residue map and the optional precision matrix.

Building the maps from topologies (create_sparse_map) and data splitting are host-side, run-once
steps that stay in the reference (out of scope)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional, Sequence

import numpy as np

from .._arrays import to_numpy
from ..custom_types.key import m_key


@dataclass(frozen=True, slots=True)
class SparseFragmentMapping:
    """(n_fragments, n_residues) map; dense ndarray / tensor, or anything with .todense() / .toarray()."""

    sparse_map: Any

    def dense(self) -> np.ndarray:
        m = self.sparse_map
        if hasattr(m, "todense"):
            m = m.todense()
        elif hasattr(m, "toarray"):
            m = m.toarray()
        return np.asarray(to_numpy(m), np.float32)


@dataclass(frozen=True, slots=True)
class Dataset:
    data: Sequence[Any]
    y_true: Any  # (n_fragments, n_timepoints, 1) for uptake, (n_fragments,) for protection factors
    data_mapping: Any
    covariance_matrix: Optional[Any] = None  # trace-normalised precision matrix (loader.py:202-217)

    @property
    def residue_feature_ouput_mapping(self):
        return self.data_mapping.sparse_map


class ExpD_Dataloader:
    """Holds train / val / test Datasets.  Only the metadata-free construction path of the reference
    (`ExpD_Dataloader(train=..., val=..., key=...)`, loader.py:106-127) is provided."""

    def __init__(self, data=None, covariance_matrix=None, train: Optional[Dataset] = None, val: Optional[Dataset] = None,
                 test: Optional[Dataset] = None, key=None):
        if data is not None:
            raise NotImplementedError(
                "building datasets from HDX_peptide objects / topologies stays in the reference "
                "(ExpD_Dataloader.create_datasets); pass train= and val= Datasets here"
            )
        self.data = []
        self.y_true = np.array([])
        self.top = []
        self.key = key if key is not None else m_key("unknown")
        self.covariance_matrix = covariance_matrix
        if train is not None:
            self.train = train
        if val is not None:
            self.val = val
        if test is not None:
            self.test = test
