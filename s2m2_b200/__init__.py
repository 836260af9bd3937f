"""s2m2_b200 — Blackwell-native (sm_100a) implementation of the S2M2 stereo inference path."""
from .model import S2M2  # noqa: F401
from .weights import MODEL_TABLE, param_spec, random_state_dict, synthetic_pair  # noqa: F401

__all__ = ["S2M2", "MODEL_TABLE", "param_spec", "random_state_dict", "synthetic_pair"]
