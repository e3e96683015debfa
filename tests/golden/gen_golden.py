#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the dev container only (the GPU box has no /root/reference):

    python tests/golden/gen_golden.py            # writes tests/golden/*.npz

The reference (wkh923/m3pc, /root/reference, read-only) is imported as-is; the only thing
replaced is four import-time dependencies that do no arithmetic on this path (matplotlib, gym,
d4rl, termcolor -- SURVEY.md section 8c).  Weights / tokenizer statistics / histories come from
``m3pc_b200.synthetic`` (numpy RandomState, seeds stored in each fixture), so a fixture holds only
seeds, the torch-drawn noise the reference consumed, and the reference's outputs.

Noise capture: the reference draws ``SquashedNormal.sample((N,))`` and ``torch.multinomial`` from
torch's global CPU generator.  We seed, pre-draw ``eps = randn(N,1,T,1,A)`` and
``q = empty(N).exponential_(1)`` in that order, re-seed, and run the reference: it consumes the
same stream (asserted below by reproducing its sampled candidates from eps).
"""
from __future__ import annotations

import json
import os
import sys
import types
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("M3PC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

for n in ["matplotlib", "matplotlib.pyplot", "matplotlib.figure", "gym", "gym.wrappers",
          "gym.wrappers.pixel_observation", "d4rl", "termcolor"]:
    sys.modules[n] = types.ModuleType(n)
sys.modules["matplotlib"].use = lambda *a, **k: None
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].figure = sys.modules["matplotlib.figure"]
sys.modules["matplotlib.figure"].Figure = object
sys.modules["gym"].Env = object

from research.omtm.models.mtm_model import omtm, omtmConfig  # noqa: E402
from research.omtm.tokenizers.base import TokenizerManager  # noqa: E402
from research.omtm.tokenizers.continuous import ContinuousTokenizer  # noqa: E402
from research.omtm.datasets.base import DataStatistics  # noqa: E402
from research.finetune_omtm.learner import Learner as FLearner  # noqa: E402
from research.finetune_omtm.model import TwinQ  # noqa: E402
from research.finetune_omtm import masks as fmasks  # noqa: E402
from research.zeroshot_omtm.learner import Learner as ZLearner  # noqa: E402
from research.zeroshot_omtm import masks as zmasks  # noqa: E402

from m3pc_b200 import synthetic as syn  # noqa: E402


def build_reference(shape: syn.ModelShape, sd_seed=0, stat_seed=1):
    cfg = omtmConfig(n_embd=shape.n_embd, n_head=shape.n_head, n_enc_layer=shape.n_enc_layer,
                     n_dec_layer=shape.n_dec_layer, dropout=0.1, norm="none")
    model = cfg.create(shape.data_shapes, shape.traj_length, {k: False for k in shape.data_shapes}).eval()
    sd = syn.make_state_dict(shape, sd_seed)
    missing = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    stats = syn.make_tokenizer_stats(shape, stat_seed)
    toks = OrderedDict()
    for k in shape.data_shapes:
        s = stats[k]
        toks[k] = ContinuousTokenizer(s["mean"], s["std"], DataStatistics(s["mean"], s["std"], s["min"], s["max"]),
                                      normalize=(k != "actions"))
    return model, TokenizerManager(toks)


def make_learner(cls, shape, model, tm, n_cand, guidance, temperature, horizon=4, critic=False):
    L = object.__new__(cls)
    L.cfg = SimpleNamespace(traj_length=shape.traj_length, device="cpu", action_samples=n_cand, discount=0.99,
                            temperature=temperature, horizon=horizon, plan_guidance=guidance, lmbda=0.6)
    L.tokenizer_manager = tm
    L.mtm = model
    if critic:
        om, os_ = syn.make_obs_norm(shape)
        qf = TwinQ(shape.obs_dim, shape.act_dim, torch.from_numpy(om), torch.from_numpy(os_)).eval()
        qf.load_state_dict({k: torch.from_numpy(v) for k, v in syn.make_critic_state_dict(shape).items()}, strict=True)
        L.iql = SimpleNamespace(qf=qf)
    return L


def draw_noise(seed, n, T, A, kind="dist", h=4):
    torch.manual_seed(seed)
    if kind == "dist":
        eps = torch.randn(n, 1, T, 1, A)
    else:
        eps = torch.randn(n, h, A)
    q = torch.empty(n).exponential_(1)
    return eps, q


def forward_cases(out_dir):
    """omtm.forward on each inference mask, hopper shapes, B=3, idx in {4, 0, 6}."""
    shape = syn.shipped_shape("hopper")
    model, tm = build_reference(shape)
    traj = syn.make_trajectories(shape, batch=3, seed=5)
    enc = tm.encode({k: torch.from_numpy(v) for k, v in traj.items()})
    T = shape.traj_length
    arrays, meta = {}, []
    creators = {"rcbc": fmasks.create_rcbc_mask, "fd": fmasks.create_fd_mask, "pi": zmasks.create_pi_mask,
                "fid": zmasks.create_fid_mask, "gid": zmasks.create_gid_mask}
    for name, fn in creators.items():
        for idx in (4, 0, 6):
            mask = fn(T, "cpu", idx)
            with torch.no_grad():
                out = model(enc, mask)
                enc_only = model.encode(enc, mask)
            tag = f"{name}_idx{idx}"
            meta.append({"tag": tag, "mask": name, "idx": idx})
            for k in ("states", "rewards", "returns"):
                arrays[f"{tag}/{k}"] = out[k].numpy()
            arrays[f"{tag}/act_mu"] = out["actions"].loc.numpy()
            arrays[f"{tag}/act_std"] = out["actions"].std.numpy()
            arrays[f"{tag}/act_mean"] = out["actions"].mean.numpy()
            arrays[f"{tag}/enc_out"] = torch.cat([enc_only[k] for k in enc_only], dim=1).numpy()
            for k, m in mask.items():
                assert m.dtype == torch.float64
                arrays[f"{tag}/mask_{k}"] = m.numpy()
    arrays["meta"] = np.array(json.dumps({"env": "hopper", "batch": 3, "traj_seed": 5, "sd_seed": 0, "stat_seed": 1, "cases": meta}))
    np.savez_compressed(os.path.join(out_dir, "forward_hopper.npz"), **arrays)
    print("forward_hopper.npz:", len(meta), "cases")


def planner_cases(out_dir):
    arrays, meta = {}, []
    specs = [
        # tag, env, guidance, N, temperature, path_length, eval, plan
        ("hopper_rtg_h4", "hopper", "rtg_guiding", 64, 0.01, 50, True, True),
        ("hopper_rtg_h4_explore", "hopper", "rtg_guiding", 64, 0.01, 50, False, True),
        ("hopper_rtg_h8", "hopper", "rtg_guiding", 48, 0.01, 0, True, True),
        ("hopper_rtg_h6", "hopper", "rtg_guiding", 48, 0.01, 2, True, True),
        ("hopper_rtg_h5", "hopper", "rtg_guiding", 48, 0.01, 3, False, True),
        ("walker_critic_h4", "walker2d", "critic_lambda_guiding", 64, 1.0, 50, True, True),
        ("walker_critic_h7", "walker2d", "critic_lambda_guiding", 40, 1.0, 1, False, True),
        ("cheetah_rtg_h4", "halfcheetah", "rtg_guiding", 96, 0.01, 123, True, True),
        ("walker_noise_h4", "walker2d", "noise_adding_lambda", 64, 1.0, 50, True, True),
        ("hopper_sampling", "hopper", "rtg_guiding", 1, 0.01, 50, False, False),
        ("hopper_rtg_pct", "hopper", "rtg_guiding", 32, 0.01, 77, False, True),  # rtg=None -> percentage path
    ]
    built = {}
    for tag, env, guidance, n, temp, pl, ev, plan in specs:
        shape = syn.shipped_shape(env)
        if env not in built:
            built[env] = build_reference(shape)
        model, tm = built[env]
        L = make_learner(FLearner, shape, model, tm, n, guidance, temp, critic=("critic" in guidance or "noise" in guidance))
        hist = syn.make_history(shape, seed=4, path_length=pl)
        T, A = shape.traj_length, shape.act_dim
        h = 4 if pl + 4 >= T else T - pl
        seed = 7
        if plan:
            eps, q = draw_noise(seed, n, T, A, "noise" if guidance == "noise_adding_lambda" else "dist", h)
        else:
            torch.manual_seed(seed)
            eps, q = torch.randn(1, T, 1, A), torch.zeros(1)
        rtg = 3.0 if (ev or tag != "hopper_rtg_pct") else None
        kw = dict(percentage=0.8, plan=plan, eval=ev, rtg=rtg)
        torch.manual_seed(seed)
        act = L.action_sample(hist, **kw)
        # second run with eval flipped gives the other output on the same noise
        torch.manual_seed(seed)
        act_other = L.action_sample(hist, **{**kw, "eval": not ev, "rtg": 3.0 if rtg is None else rtg}) if rtg is not None else None
        meta.append({"tag": tag, "env": env, "guidance": guidance, "n_cand": n, "temperature": temp, "path_length": pl,
                     "eval": ev, "plan": plan, "rtg": rtg, "percentage": 0.8, "horizon": h, "hist_seed": 4})
        arrays[f"{tag}/eps"] = eps.numpy()
        arrays[f"{tag}/q"] = q.numpy()
        arrays[f"{tag}/action"] = act.numpy()
        if act_other is not None:
            arrays[f"{tag}/action_other"] = act_other.numpy()
    arrays["meta"] = np.array(json.dumps({"sd_seed": 0, "stat_seed": 1, "cases": meta}))
    np.savez_compressed(os.path.join(out_dir, "planner.npz"), **arrays)
    print("planner.npz:", len(meta), "cases")


def zeroshot_cases(out_dir):
    arrays, meta = {}, []
    shape = syn.shipped_shape("hopper")
    model, tm = build_reference(shape)
    for tag, fn, pl in [("id_pl50", "action_id_sample", 50), ("piid_pl50", "action_piid_sample", 50),
                        ("piid_pl2", "action_piid_sample", 2), ("piid_pl997", "action_piid_sample", 997),
                        ("id_pl998", "action_id_sample", 998)]:
        L = make_learner(ZLearner, shape, model, tm, 1, "rtg_guiding", 0.01)
        hist = syn.make_history(shape, seed=4, path_length=pl)
        T, A = shape.traj_length, shape.act_dim
        torch.manual_seed(11)
        eps = torch.randn(1, T, 1, A)
        torch.manual_seed(11)
        sample = getattr(L, fn)(hist, eval=False, rtg=2.5)
        mean = getattr(L, fn)(hist, eval=True, rtg=2.5)
        meta.append({"tag": tag, "fn": fn, "path_length": pl, "rtg": 2.5, "hist_seed": 4})
        arrays[f"{tag}/eps"] = eps.numpy()
        arrays[f"{tag}/sample_action"] = sample.numpy()
        arrays[f"{tag}/eval_action"] = mean.numpy()
    arrays["meta"] = np.array(json.dumps({"env": "hopper", "sd_seed": 0, "stat_seed": 1, "cases": meta}))
    np.savez_compressed(os.path.join(out_dir, "zeroshot_hopper.npz"), **arrays)
    print("zeroshot_hopper.npz:", len(meta), "cases")


def mask_cases(out_dir):
    """Mask layouts (bit-exact contract) for T in {8, 16}, every idx."""
    arrays = {}
    creators = {"rcbc": fmasks.create_rcbc_mask, "fd": fmasks.create_fd_mask, "pi": zmasks.create_pi_mask,
                "fid": zmasks.create_fid_mask, "gid": zmasks.create_gid_mask}
    for T in (8, 16):
        for name, fn in creators.items():
            for idx in range(T):
                m = fn(T, "cpu", idx)
                arrays[f"T{T}/{name}/{idx}"] = np.stack([m[k].numpy() for k in ("states", "actions", "rewards", "returns")])
    np.savez_compressed(os.path.join(out_dir, "masks.npz"), **arrays)
    print("masks.npz:", len(arrays), "layouts")


if __name__ == "__main__":
    torch.set_num_threads(8)
    mask_cases(HERE)
    forward_cases(HERE)
    planner_cases(HERE)
    zeroshot_cases(HERE)
