"""m3pc_b200 -- B200-native (sm_100a) implementation of the M^3PC test-time planning hot path.

Host-side mirror of the reference's model / planner API (``omtm``, ``omtmConfig``, ``TokenizerManager``,
``ContinuousTokenizer``, the inference mask creators, the forward and zero-shot ``Learner`` planners)
over a C-ABI CUDA library (``include/m3pc.h``, built from ``m3pc_b200/csrc``).  There is no CPU fallback:
every compute entry point raises if the CUDA library or a GPU is missing.
"""
__version__ = "0.2.0"

from . import synthetic  # noqa: F401  (numpy only)

_LAZY = {
    "omtm": "mtm_model", "omtmConfig": "mtm_model", "SquashedNormal": "mtm_model",
    "TokenizerManager": "tokenizers", "ContinuousTokenizer": "tokenizers", "DataStatistics": "tokenizers",
    "TwinQ": "critic", "PlanEngine": "engine", "Learner": "learner", "PlannerMixin": "learner",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        return getattr(importlib.import_module(f"{__name__}.{_LAZY[name]}"), name)
    raise AttributeError(name)
