#!/usr/bin/env python
"""Generate tests/golden/ckpt/* -- files in the reference's on-disk formats, written BY the unmodified reference classes.

Run in the dev container only (needs /root/reference):  python tests/golden/gen_checkpoint_fixture.py

  d4rl_statistics_tiny.pkl   pickle of {"states","actions","rewards","values": research.omtm.datasets.base.DataStatistics}
                             (the pre-rename key "values", as older caches hold it; sequence_dataset.py:372-378)
  mtm_7.pt                   torch.save({"model": omtm.state_dict(), "optimizer": AdamW.state_dict(), "step": 7})  (finetune.py:319-326)
  iql_7.pt                   {"qf": TwinQ.state_dict(), "vf": ..., "actor": ..., "total_it": 7}                     (model.py:310-320)
  expect.npz                 what the reference itself makes of them: ContinuousTokenizer.create's mean / std per modality
The model is tiny (n_embd 128, 1+1 layers, T=4, obs 3 / act 2) to keep the fixture small; key names and shapes follow the
same rules as the shipped configuration.
"""
import os
import pickle
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("M3PC_REFERENCE", "/root/reference"))
for n in ["matplotlib", "matplotlib.pyplot", "matplotlib.figure", "gym", "gym.wrappers", "gym.wrappers.pixel_observation", "d4rl", "termcolor"]:
    sys.modules[n] = types.ModuleType(n)
sys.modules["matplotlib"].use = lambda *a, **k: None
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].figure = sys.modules["matplotlib.figure"]
sys.modules["matplotlib.figure"].Figure = object
sys.modules["gym"].Env = object

from research.omtm.models.mtm_model import omtm, omtmConfig  # noqa: E402
from research.omtm.tokenizers.continuous import ContinuousTokenizer  # noqa: E402
from research.omtm.datasets.base import DataStatistics  # noqa: E402
from research.finetune_omtm.model import TwinQ  # noqa: E402

out = os.path.join(HERE, "ckpt")
os.makedirs(out, exist_ok=True)
rs = np.random.RandomState(11)
obs, act, T = 3, 2, 4
dims = {"states": obs, "actions": act, "rewards": 1, "values": 1}
stats = {}
for k, d in dims.items():
    mean = rs.randn(d)
    std = rs.uniform(0.02, 1.5, size=d)  # some entries fall under continuous.py:58's 0.1 threshold
    stats[k] = DataStatistics(mean=mean, std=std, min=mean - 3 * std, max=mean + 3 * std)
with open(os.path.join(out, "d4rl_statistics_tiny.pkl"), "wb") as f:
    pickle.dump(stats, f)


class _DS:  # the slice of SequenceDataset that ContinuousTokenizer.create touches
    def trajectory_statistics(self):
        with open(os.path.join(out, "d4rl_statistics_tiny.pkl"), "rb") as f:
            d = pickle.load(f)
        d["returns"] = d.pop("values")
        return d


expect = {}
for k in ("states", "actions", "rewards", "returns"):
    tok = ContinuousTokenizer.create(k, _DS())
    expect[f"{k}/mean"] = tok._data_mean.detach().numpy()
    expect[f"{k}/std"] = tok._data_std.detach().numpy()
    expect[f"{k}/normalize"] = np.array(bool(tok.normalize))
    x = torch.from_numpy(rs.randn(2, T, dims["values" if k == "returns" else k]).astype(np.float32))
    expect[f"{k}/x"] = x.numpy()
    expect[f"{k}/encoded"] = tok.encode(x).numpy()

torch.manual_seed(0)
shapes = {"states": (1, obs), "actions": (1, act), "rewards": (1, 1), "returns": (1, 1)}
m = omtm(shapes, T, {k: False for k in shapes}, omtmConfig(n_embd=128, n_head=1, n_enc_layer=1, n_dec_layer=1, dropout=0.1, norm="none"))
opt = torch.optim.AdamW(m.parameters(), lr=1e-4)
torch.save({"model": m.state_dict(), "optimizer": opt.state_dict(), "step": 7}, os.path.join(out, "mtm_7.pt"))
expect["mtm_keys"] = np.array(sorted(m.state_dict().keys()))
expect["mtm_checksum"] = np.array(sum(float(v.double().sum()) for v in m.state_dict().values()))
q = TwinQ(obs, act, torch.zeros(obs), torch.ones(obs))
torch.save({"qf": q.state_dict(), "vf": {}, "actor": {}, "q_optimizer": {}, "v_optimizer": {}, "actor_optimizer": {}, "actor_lr_schedule": {},
            "total_it": 7}, os.path.join(out, "iql_7.pt"))
expect["qf_keys"] = np.array(sorted(q.state_dict().keys()))
expect["qf_checksum"] = np.array(sum(float(v.double().sum()) for v in q.state_dict().values()))
np.savez_compressed(os.path.join(out, "expect.npz"), **expect)
print("wrote", sorted(os.listdir(out)), {f: os.path.getsize(os.path.join(out, f)) for f in os.listdir(out)})
