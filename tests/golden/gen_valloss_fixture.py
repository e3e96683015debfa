#!/usr/bin/env python
"""Golden vectors for the validation-loss path, written by the UNMODIFIED reference (run in the dev container, where
/root/reference exists; the output ``valloss.npz`` is committed):

  * ``create_random_autoregressize_mask`` (research/finetune_omtm/masks.py:98-125) under 60 numpy seeds x 3 configurations;
  * ``Learner.compute_mtm_loss`` (research/finetune_omtm/learner.py:419-503) on synthetic walker2d / hopper batches with the
    synthetic weights of m3pc_b200.synthetic, torch CPU fp32: the masks it drew, the normal draws of its single-sample entropy
    estimate (``torch.manual_seed`` -> ``torch.randn``, the stream ``Normal.rsample`` consumes), and the 5-tuple it returned.

    python tests/golden/gen_valloss_fixture.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from m3pc_b200 import synthetic as syn  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

MASK_CONFIGS = [((0.5,), (0.1, 0.1, 0.7, 0.1)), ((0.3, 0.7, 1.0), (0, 0, 0.7, 0.3)), (0.6, (0.25, 0.25, 0.25, 0.25))]
LOSS_CASES = [("walker2d", 6, 3, (0.5,), (0.1, 0.1, 0.7, 0.1)), ("walker2d", 33, 11, (0.7,), (0, 0, 0.7, 0.3)),
              ("hopper", 5, 4, (0.3, 0.7, 1.0), (0.25, 0.25, 0.25, 0.25)), ("hopper", 40, 8, 0.6, (0.1, 0.1, 0.7, 0.1))]


def main():
    rh._install()
    from research.finetune_omtm import masks as RM
    arrays, meta = {}, {"mask_configs": [[list(r) if isinstance(r, tuple) else r, list(p)] for r, p in MASK_CONFIGS], "mask_seeds": 60, "loss_cases": []}
    shapes = {"states": (1, 17), "actions": (1, 6), "rewards": (1, 1), "returns": (1, 1)}
    for ci, (ratios, pw) in enumerate(MASK_CONFIGS):
        for seed in range(60):
            for T in (8, 16):
                np.random.seed(seed)
                m = RM.create_random_autoregressize_mask(shapes, ratios, T, "cpu", pw)
                assert all(v.dtype == torch.float64 and tuple(v.shape) == (T, 1) for v in m.values())
                arrays[f"mask/{ci}/{T}/{seed}"] = np.stack([m[k].numpy()[:, 0] for k in ("states", "actions", "rewards", "returns")])
    torch.set_num_threads(8)
    for env, B, seed, ratios, pw in LOSS_CASES:
        shape = syn.shipped_shape(env)
        L = rh.build_learner(shape, guidance="rtg_guiding", n_cand=8, temperature=1.0, device="cpu")
        L.cfg.mask_ratio, L.cfg.p_weights = ratios, pw
        batch = {k: torch.from_numpy(v) for k, v in syn.make_trajectories(shape, B, seed).items()}
        T, A = shape.traj_length, shape.act_dim
        np.random.seed(seed)
        masks = RM.create_random_autoregressize_mask(shape.data_shapes, ratios, T, "cpu", pw)
        torch.manual_seed(seed)
        eps = torch.randn(1, B, T, 1, A)
        np.random.seed(seed)
        torch.manual_seed(seed)
        with torch.no_grad():
            loss, losses, masked, masked_c, entropy = L.compute_mtm_loss(batch, shape.data_shapes, {k: False for k in shape.data_shapes},
                                                                        L.mtm.temperature().detach())
        tag = f"loss/{env}/{B}/{seed}"
        arrays[f"{tag}/masks"] = np.stack([masks[k].numpy()[:, 0] for k in ("states", "actions", "rewards", "returns")])
        arrays[f"{tag}/eps"] = eps.numpy()
        out = {"loss": float(loss), "entropy": float(entropy), "entropy_reg": float(L.mtm.temperature())}
        out.update({f"losses/{k}": float(v) for k, v in losses.items()})
        out.update({f"masked/{k}": float(v) for k, v in masked.items()})
        out.update({f"masked_c/{k}": float(v) for k, v in masked_c.items()})
        meta["loss_cases"].append({"env": env, "B": B, "seed": seed, "ratios": list(ratios) if isinstance(ratios, tuple) else ratios,
                                   "p_weights": list(pw), "tag": tag, "out": out})
        print(tag, out)
    arrays["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "valloss.npz"), **arrays)
    print("valloss.npz:", len(arrays), "arrays")


if __name__ == "__main__":
    main()
