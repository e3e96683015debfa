"""Pin the oracle (oracle/*.py) against fixtures produced by the live reference (tests/golden/gen_golden.py).

CPU only.  Tolerances: the oracle restates nn.TransformerEncoderLayer's fused fast path with plain ops,
which differs from it by fp32 round-off only (SURVEY.md 3.4 measured <= 6e-6 abs).
"""
import json
import os

import numpy as np
import pytest
import torch

from m3pc_b200 import synthetic as syn
from oracle import mtm_oracle as mo
from oracle import planner_oracle as po

CREATORS = {"rcbc": po.create_rcbc_mask, "fd": po.create_fd_mask, "pi": po.create_pi_mask,
            "fid": po.create_fid_mask, "gid": po.create_gid_mask}


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return z, json.loads(str(z["meta"])) if "meta" in z.files else None


def test_mask_layouts_bit_exact(golden_dir):
    z, _ = _load(golden_dir, "masks.npz")
    n = 0
    for key in z.files:
        T, name, idx = key.split("/")
        m = CREATORS[name](int(T[1:]), int(idx))
        got = np.stack([m[k] for k in ("states", "actions", "rewards", "returns")])
        assert got.dtype == np.float64
        assert np.array_equal(got, z[key]), key
        n += 1
    assert n == 120


@pytest.fixture(scope="module")
def hopper():
    shape = syn.shipped_shape("hopper")
    return shape, syn.make_state_dict(shape, 0), syn.make_tokenizer_stats(shape, 1)


def test_forward_matches_reference(golden_dir, hopper):
    shape, sd_np, stats_np = hopper
    z, meta = _load(golden_dir, "forward_hopper.npz")
    sd, stats = mo.to_torch(sd_np), mo.stats_to_torch(stats_np)
    traj = {k: torch.from_numpy(v) for k, v in syn.make_trajectories(shape, meta["batch"], meta["traj_seed"]).items()}
    enc = mo.encode_all(traj, stats)
    for case in meta["cases"]:
        tag = case["tag"]
        masks = {k: torch.from_numpy(v) for k, v in CREATORS[case["mask"]](shape.traj_length, case["idx"]).items()}
        for k, m in masks.items():
            assert np.array_equal(m.numpy(), z[f"{tag}/mask_{k}"])
        out = mo.mtm_forward(sd, enc, masks, shape.n_head, shape.n_enc_layer, shape.n_dec_layer, return_stages=True)
        for k in ("states", "rewards", "returns"):
            np.testing.assert_allclose(out[k].numpy(), z[f"{tag}/{k}"], rtol=2e-5, atol=2e-5, err_msg=f"{tag}/{k}")
        np.testing.assert_allclose(out["actions"]["mu"].numpy(), z[f"{tag}/act_mu"], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(out["actions"]["std"].numpy(), z[f"{tag}/act_std"], rtol=5e-5, atol=1e-6)
        np.testing.assert_allclose(torch.tanh(out["actions"]["mu"]).numpy(), z[f"{tag}/act_mean"], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(out["_stages"]["enc_out"].numpy(), z[f"{tag}/enc_out"], rtol=2e-5, atol=2e-5)


def _planner(case, n_cand=None):
    shape = syn.shipped_shape(case["env"])
    need_critic = case["guidance"] in ("critic_lambda_guiding", "noise_adding_lambda")
    return shape, po.from_synthetic(
        shape, syn.make_state_dict(shape, 0), syn.make_tokenizer_stats(shape, 1),
        critic_np=syn.make_critic_state_dict(shape) if need_critic else None,
        obs_norm=syn.make_obs_norm(shape) if need_critic else None,
        action_samples=case["n_cand"], temperature=case["temperature"], plan_guidance=case["guidance"])


def test_planners_match_reference(golden_dir):
    z, meta = _load(golden_dir, "planner.npz")
    for case in meta["cases"]:
        tag = case["tag"]
        shape, P = _planner(case)
        hist = syn.make_history(shape, seed=case["hist_seed"], path_length=case["path_length"])
        eps, q = torch.from_numpy(z[f"{tag}/eps"]), torch.from_numpy(z[f"{tag}/q"])
        act, dbg = P.action_sample(hist, percentage=case["percentage"], plan=case["plan"], eval=case["eval"],
                                   rtg=case["rtg"], eps=eps, q=q if case["plan"] else None)
        assert dbg["horizon"] == case["horizon"], tag
        ref = z[f"{tag}/action"]
        assert tuple(act.shape) == tuple(ref.shape), (tag, act.shape, ref.shape)
        np.testing.assert_allclose(act.numpy(), ref, rtol=1e-4, atol=2e-5, err_msg=tag)
        if f"{tag}/action_other" in z.files:
            other = dbg["sample_action"] if case["eval"] else dbg["eval_action"]
            np.testing.assert_allclose(other.numpy(), z[f"{tag}/action_other"], rtol=1e-4, atol=2e-5, err_msg=tag + " other")


def test_zeroshot_matches_reference(golden_dir):
    z, meta = _load(golden_dir, "zeroshot_hopper.npz")
    shape = syn.shipped_shape("hopper")
    P = po.from_synthetic(shape, syn.make_state_dict(shape, 0), syn.make_tokenizer_stats(shape, 1), action_samples=1)
    for case in meta["cases"]:
        tag = case["tag"]
        hist = syn.make_history(shape, seed=case["hist_seed"], path_length=case["path_length"])
        eps = torch.from_numpy(z[f"{tag}/eps"])
        fn = getattr(P, case["fn"])
        _, dbg = fn(hist, eval=False, rtg=case["rtg"], eps1=eps)
        np.testing.assert_allclose(dbg["sample_action"].numpy(), z[f"{tag}/sample_action"], rtol=1e-4, atol=2e-5, err_msg=tag)
        np.testing.assert_allclose(dbg["eval_action"].numpy(), z[f"{tag}/eval_action"], rtol=1e-4, atol=2e-5, err_msg=tag)


def test_fp64_oracle_close_to_fp32(hopper):
    """The float64 oracle is the ground truth both the reference and the CUDA path are measured against."""
    shape, sd_np, stats_np = hopper
    traj_np = syn.make_trajectories(shape, 2, 9)
    masks = {k: torch.from_numpy(v) for k, v in po.create_fd_mask(shape.traj_length, 4).items()}
    outs = []
    for dt in (torch.float32, torch.float64):
        sd, stats = mo.to_torch(sd_np, dt), mo.stats_to_torch(stats_np, dt)
        enc = mo.encode_all({k: torch.from_numpy(v).to(dt) for k, v in traj_np.items()}, stats)
        outs.append(mo.mtm_forward(sd, enc, masks, shape.n_head, shape.n_enc_layer, shape.n_dec_layer))
    for k in ("states", "rewards", "returns"):
        np.testing.assert_allclose(outs[0][k].numpy(), outs[1][k].numpy(), rtol=1e-4, atol=1e-4)


def test_library_ops_equal_their_published_formulas():
    torch.manual_seed(0)
    x = torch.randn(7, 13, 512, dtype=torch.float64) * 2 + 0.3
    w, b = torch.rand(512, dtype=torch.float64) + 0.5, torch.randn(512, dtype=torch.float64)
    np.testing.assert_allclose(mo.layer_norm(x, w, b).numpy(), mo.layer_norm_plain(x, w, b).numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(mo.gelu_erf(x).numpy(), mo.gelu_erf_plain(x).numpy(), rtol=1e-12, atol=1e-12)
    q, k, v = (torch.randn(3, 4, 13, 128, dtype=torch.float64) for _ in range(3))
    ref = torch.softmax(q @ k.transpose(-1, -2) / 128 ** 0.5, dim=-1) @ v
    np.testing.assert_allclose(torch.nn.functional.scaled_dot_product_attention(q, k, v).numpy(), ref.numpy(), rtol=1e-10, atol=1e-12)
