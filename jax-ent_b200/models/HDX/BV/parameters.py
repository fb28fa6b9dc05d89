"""BV_Model_Parameters (reference jaxent/src/models/HDX/BV/parameters.py:14-58)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, ClassVar

import numpy as np

from ....custom_types.key import m_key
from ....interfaces.model import Model_Parameters


@dataclass(frozen=True, slots=True)
class BV_Model_Parameters(Model_Parameters):
    bv_bc: Any = field(default_factory=lambda: np.array([0.35], np.float32))
    bv_bh: Any = field(default_factory=lambda: np.array([2.0], np.float32))
    temperature: float = 300.0
    timepoints: Any = field(default_factory=lambda: np.array([0.167, 1.0, 10.0], np.float32))
    key: ClassVar = frozenset({m_key("HDX_resPF"), m_key("HDX_peptide")})
    static_params: ClassVar[set] = {"temperature", "key", "timepoints"}

    def __mul__(self, scalar):
        return BV_Model_Parameters(self.bv_bc * scalar, self.bv_bh * scalar, self.temperature, self.timepoints)

    __rmul__ = __mul__

    def __sub__(self, other):
        return BV_Model_Parameters(self.bv_bc - other.bv_bc, self.bv_bh - other.bv_bh, self.temperature, self.timepoints)

    def update_parameters(self, new_params):
        return BV_Model_Parameters(new_params.bv_bc, new_params.bv_bh, self.temperature, self.timepoints)
