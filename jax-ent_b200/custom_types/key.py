"""m_key / m_id (reference jaxent/src/custom_types/key.py)."""
from typing import NewType

m_key = NewType("m_key", str)
m_id = NewType("m_id", str)
