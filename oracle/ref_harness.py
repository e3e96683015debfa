"""ORACLE (test infrastructure, NOT product code) -- drives the UNMODIFIED reference planner.

Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s reference / baseline legs may import this.

Finds the reference package (``oracle/_ref`` staged by oracle/stage_ref.py, else /root/reference in the dev container),
installs the four import-time stubs that do no arithmetic on this path (matplotlib, gym, d4rl, termcolor -- SURVEY.md
section 8c), and builds the reference's own objects -- ``omtm`` (research/omtm/models/mtm_model.py:324), ``TokenizerManager`` /
``ContinuousTokenizer`` (research/omtm/tokenizers), ``TwinQ`` (research/finetune_omtm/model.py:146) and ``Learner``
(research/finetune_omtm/learner.py:17; constructed with ``object.__new__`` because ``__init__`` wants a gym env and a
checkpoint file) -- with the synthetic weights of ``m3pc_b200.synthetic``.  Nothing of the reference is modified or patched.
"""
from __future__ import annotations

import os
import sys
import types
from collections import OrderedDict
from types import SimpleNamespace
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = (os.path.join(HERE, "_ref"), "/root/reference")
_ref_root: Optional[str] = None


def reference_root() -> Optional[str]:
    """Directory holding the reference's ``research`` package, or None when neither the staged copy nor /root/reference exists."""
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "research", "finetune_omtm", "learner.py")):
            return c
    return None


def available() -> bool:
    return reference_root() is not None


def _install():
    global _ref_root
    if _ref_root is not None:
        return
    root = reference_root()
    if root is None:
        raise ImportError("the reference is not staged: run `python oracle/stage_ref.py` where /root/reference exists")
    for n in ["matplotlib", "matplotlib.pyplot", "matplotlib.figure", "gym", "gym.wrappers", "gym.wrappers.pixel_observation", "d4rl", "termcolor"]:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "use"):
        mpl.use = lambda *a, **k: None
    if not hasattr(mpl, "pyplot"):
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(mpl, "figure"):
        mpl.figure = sys.modules["matplotlib.figure"]
    if not hasattr(sys.modules["matplotlib.figure"], "Figure"):
        sys.modules["matplotlib.figure"].Figure = object
    if not hasattr(sys.modules["gym"], "Env"):
        sys.modules["gym"].Env = object
    if root not in sys.path:
        sys.path.insert(0, root)
    _ref_root = root


def build_learner(shape, *, guidance: str, n_cand: int, temperature: float, device: str = "cpu", horizon: int = 4, zeroshot: bool = False,
                  sd_seed: int = 0, stat_seed: int = 1):
    """The reference ``Learner`` (finetune_omtm, or zeroshot_omtm when ``zeroshot``) on ``device`` with synthetic weights."""
    import torch
    _install()
    from research.omtm.datasets.base import DataStatistics
    from research.omtm.models.mtm_model import omtmConfig
    from research.omtm.tokenizers.base import TokenizerManager
    from research.omtm.tokenizers.continuous import ContinuousTokenizer
    from m3pc_b200 import synthetic as syn
    if zeroshot:
        from research.zeroshot_omtm.learner import Learner
    else:
        from research.finetune_omtm.learner import Learner
    cfg = omtmConfig(n_embd=shape.n_embd, n_head=shape.n_head, n_enc_layer=shape.n_enc_layer, n_dec_layer=shape.n_dec_layer, dropout=0.1, norm="none")
    model = cfg.create(shape.data_shapes, shape.traj_length, {k: False for k in shape.data_shapes}).eval()
    res = model.load_state_dict({k: torch.from_numpy(v) for k, v in syn.make_state_dict(shape, sd_seed).items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    model = model.to(device)
    stats = syn.make_tokenizer_stats(shape, stat_seed)
    toks = OrderedDict()
    for k in shape.data_shapes:
        s = stats[k]
        tk = ContinuousTokenizer(s["mean"], s["std"], DataStatistics(s["mean"], s["std"], s["min"], s["max"]), normalize=(k != "actions"))
        toks[k] = tk.to(device) if hasattr(tk, "to") else tk
    L = object.__new__(Learner)
    L.cfg = SimpleNamespace(traj_length=shape.traj_length, device=device, action_samples=n_cand, discount=0.99, temperature=temperature,
                            horizon=horizon, plan_guidance=guidance, lmbda=0.6)
    L.tokenizer_manager = TokenizerManager(toks)
    if hasattr(L.tokenizer_manager, "to"):
        L.tokenizer_manager = L.tokenizer_manager.to(device)
    L.mtm = model
    if guidance != "rtg_guiding" and not zeroshot:
        from research.finetune_omtm.model import TwinQ
        om, os_ = syn.make_obs_norm(shape)
        qf = TwinQ(shape.obs_dim, shape.act_dim, torch.from_numpy(om).to(device), torch.from_numpy(os_).to(device)).eval()
        qf.load_state_dict({k: torch.from_numpy(v) for k, v in syn.make_critic_state_dict(shape).items()}, strict=True)
        L.iql = SimpleNamespace(qf=qf.to(device))
    return L
