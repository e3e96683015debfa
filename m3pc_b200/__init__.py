"""m3pc_b200 -- B200-native (sm_100a) implementation of the M^3PC test-time planning hot path.

Host-side mirror of the reference's model / planner API (``omtm``, ``omtmConfig``, ``TokenizerManager``,
``ContinuousTokenizer``, the inference mask creators, the forward and zero-shot ``Learner`` planners)
over a C-ABI CUDA library (``include/m3pc.h``, built from ``m3pc_b200/csrc``).  There is no CPU fallback:
every compute entry point raises if the CUDA library or a GPU is missing.
"""
__version__ = "0.1.0"
