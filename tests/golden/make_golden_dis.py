#!/usr/bin/env python
"""Generates tests/golden/d3_*.npz by running the REAL disentangled reference CUDA on a GPU box.

    gpurun -- python tests/golden/make_golden_dis.py     # writes gpurun_out/golden/d3_*.npz
    cp gpurun_out/golden/d3_*.npz tests/golden/           # commit them

The reference is submodules/diff-gaussian-rasterization-disentangle-optim compiled unmodified by
oracle/build_ref.py into oracle/_ref/ref_D_C.so (F=3, 16x16 tiles -- the only shape D/ compiles at,
SURVEY 0.2).  Each file holds the seeded inputs (including the language footprint) and every output /
internal buffer of one forward + two backwards (the second measures the reference's own atomic noise).
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
    "d3_small": dict(P=2500, F=3, W=112, H=80, seed=0, view=0, scale=0.05),
    "d3_view1_bg": dict(P=3000, F=3, W=120, H=75, seed=3, view=1, scale=0.08, bg=(0.2, 0.5, 0.7)),
}


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    mod = U.ref_module("ref_D_C")
    if mod is None:
        print("reference module not available (no GPU or oracle/_ref/ref_D_C.so missing)")
        return 1
    dev = torch.device("cuda:0")
    for name, kw in CASES.items():
        sc = U.add_lang_footprint(U.make_scene(**kw), seed=7)
        grads = U.loss_weights(sc["F"], sc["W"], sc["H"], seed=1)
        r = U.run_ref_dis(mod, sc, dev, grads=grads)
        r2 = U.run_ref_dis(mod, sc, dev, grads=grads)
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
        print(name, "R =", r["R"], "R_lang =", r["R_lang"], "visible =", int((r["radii"] > 0).sum()),
              int((r["radii_lang"] > 0).sum()), "bytes =", os.path.getsize(path))
    return 0


if __name__ == "__main__":
    sys.exit(main())
