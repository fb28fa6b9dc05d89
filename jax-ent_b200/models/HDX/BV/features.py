"""BV feature containers (reference jaxent/src/models/HDX/BV/features.py:16-94)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, ClassVar

from ....custom_types.key import m_key


@dataclass(frozen=True, slots=True)
class BV_input_features:
    heavy_contacts: Any  # (n_residues, n_frames), frame axis contiguous
    acceptor_contacts: Any
    k_ints: Any = None  # (n_residues,)
    __features__: ClassVar[set] = {"heavy_contacts", "acceptor_contacts", "k_ints"}
    key: ClassVar[set] = {m_key("HDX_resPF"), m_key("HDX_peptide")}

    @property
    def features_shape(self) -> tuple:
        h = self.heavy_contacts
        if type(h) is not type(self.acceptor_contacts):
            raise TypeError("heavy_contacts and acceptor_contacts must be of the same type")
        shp = tuple(h.shape)
        return (shp[0], 1) if len(shp) == 1 else (shp[0], shp[1])

    @property
    def feat_pred(self):
        return (self.heavy_contacts, self.acceptor_contacts)

    # .npz round trip in the reference's key layout (custom_types/features.py:91-286)
    def save(self, filepath: str) -> None:
        from ....utils.io import save_features_npz

        save_features_npz(self, filepath)

    save_features = save

    @classmethod
    def load(cls, filepath):
        from ....utils.io import load_features_npz

        return load_features_npz(str(filepath))

    load_features = load


@dataclass(frozen=True, slots=True)
class BV_output_features:
    log_Pf: Any
    k_ints: Any = None
    __features__: ClassVar[set] = {"log_Pf", "k_ints"}
    key: ClassVar = m_key("HDX_resPF")

    @property
    def output_shape(self) -> tuple:
        return (1, len(self.log_Pf))

    def y_pred(self):
        return self.log_Pf


@dataclass(frozen=True, slots=True)
class uptake_BV_output_features:
    uptake: Any  # (n_timepoints, n_residues)
    __features__: ClassVar[set] = {"uptake"}
    key: ClassVar = m_key("HDX_peptide")

    @property
    def output_shape(self) -> tuple:
        return tuple(self.uptake.shape)

    def y_pred(self):
        return self.uptake
