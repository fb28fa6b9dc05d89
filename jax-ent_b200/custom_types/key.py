"""String tags used as dictionary keys for forward-model outputs and identifiers (same names and call behaviour as the
reference's `custom_types/key.py`: `m_key("HDX_peptide")` returns the plain string, statically typed as its own kind)."""
import typing


def _string_tag(name: str):
    return typing.NewType(name, str)


m_key, m_id = _string_tag("m_key"), _string_tag("m_id")
