#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REAL reference CUDA on a GPU box.

    gpurun -- python tests/golden/make_golden.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/            # commit them

The reference is submodules/diff-gaussian-rasterization compiled unmodified by oracle/build_ref.py
into oracle/_ref/ref_P_C.so (F=15, 15x15 tiles).  Each file holds the seeded inputs and every
output / internal buffer of one forward + backward, so the CPU oracle and the CUDA path can be
pinned against the reference without the reference being present.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _util as U  # noqa: E402

CASES = {
    # name: make_scene kwargs
    "p15_small": dict(P=2000, F=15, W=96, H=64, seed=0, view=0, scale=0.05),
    "p15_view1_bg": dict(P=3000, F=15, W=120, H=75, seed=3, view=1, scale=0.08, bg=(0.2, 0.5, 0.7)),
    "p15_dense": dict(P=6000, F=15, W=64, H=48, seed=5, view=2, scale=0.03),
}


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    mod = U.ref_module("ref_P_C")
    if mod is None:
        print("reference module not available (no GPU or oracle/_ref/ref_P_C.so missing)")
        return 1
    dev = torch.device("cuda:0")
    for name, kw in CASES.items():
        sc = U.make_scene(**kw)
        grads = U.loss_weights(sc["F"], sc["W"], sc["H"], seed=1)
        r = U.run_ref(mod, sc, dev, grads=grads)
        r2 = U.run_ref(mod, sc, dev, grads=grads)  # second run: measures the reference's own atomic-order noise
        save = {k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in sc.items()}
        save.update({"gw_color": grads[0].numpy(), "gw_language": grads[1].numpy(), "gw_depth": grads[2].numpy()})
        for k, v in r.items():
            if k == "grads":
                for gk, gv in v.items():
                    save["grad_" + gk] = gv
                    save["grad2_" + gk] = r2["grads"][gk]
            else:
                save["out_" + k] = np.asarray(v)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **save)
        print(name, "R =", r["R"], "visible =", int((r["radii"] > 0).sum()), "bytes =", os.path.getsize(path))
    return 0


if __name__ == "__main__":
    sys.exit(main())
