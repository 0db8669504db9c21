"""Build-container check behind DESIGN.md Q15 (TEST INFRASTRUCTURE ONLY; needs /root/reference).

    python -m oracle.check_sdpa_vs_eager

The gradient goldens are frozen with the reference's `eager` attention branch.  For the backbones whose masks never
produce a row without an allowed key (Qwen3SessionMoe, Qwen3Moe) that IS the gradient of the default `sdpa` branch —
this script measures the difference (2e-6 relative, summation order).  For Qwen3Multi the fused SDPA backward inflates
the gradient of the fully masked ("uniform", quirk Q1) cross-attention rows by the key count, so the two branches
disagree by orders of magnitude while eager agrees with finite differences; the CUDA path implements the eager one.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_golden as mg, ref_shim  # noqa: E402
from gamer_b200 import synthetic as syn  # noqa: E402


def main():
    R = ref_shim.load_reference()
    for variant, seed, bseed in (("Qwen3SessionMoe", 44, 5), ("Qwen3Moe", 47, 8), ("Qwen3Multi", 42, 3)):
        cfg = mg.tiny_config(variant)
        batch = syn.make_train_batch(syn.make_catalogue(2000, 1), 6, max_his_len=12, seed=bseed, median_len=6)
        g = {}
        for impl in ("eager", "sdpa"):
            m, _, _ = mg.build(R, variant, cfg, seed, impl)
            m(**batch).loss.backward()
            g[impl] = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in m.named_parameters()}
        worst = max(((g["eager"][k] - g["sdpa"][k]).norm() / (g["eager"][k].norm() + 1e-12)).item() for k in g["eager"])
        print(f"{variant}: worst per-parameter relative difference, eager vs sdpa gradients = {worst:.3e}")


if __name__ == "__main__":
    main()
