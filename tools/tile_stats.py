"""Per-tile traversal statistics of the forward blend at the headline shape (development aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import _util as U
dev = torch.device("cuda:0")
for view in (0, 1, 5):
    sc = U.make_scene(P=1000000, F=15, W=960, H=540, seed=0, scale=0.01, view=view)
    o = U.run_ours(sc, dev, tile=15, bitexact=False)
    ws = o["ws"]
    rg = ws["ranges"].astype(np.int64); ln = rg[:, 1] - rg[:, 0]
    gx, gy = 64, 36
    T = ws["final_T"]; nc = ws["n_contrib"].astype(np.int64)
    Tp = np.full((gy * 15, gx * 15), 0.0, np.float32); Tp[:540, :960] = T
    ncp = np.zeros((gy * 15, gx * 15), np.int64); ncp[:540, :960] = nc
    inside = np.zeros((gy * 15, gx * 15), bool); inside[:540, :960] = True
    Tt = Tp.reshape(gy, 15, gx, 15).transpose(0, 2, 1, 3).reshape(gy * gx, 225)
    nt = ncp.reshape(gy, 15, gx, 15).transpose(0, 2, 1, 3).reshape(gy * gx, 225)
    it = inside.reshape(gy, 15, gx, 15).transpose(0, 2, 1, 3).reshape(gy * gx, 225)
    # a pixel certainly ran through its whole list if its transmittance stayed well above the 1e-4 cut
    unsat = ((Tt >= 1e-2) & it).any(1)
    trav = np.where(unsat, ln, np.minimum(nt.max(1) + 1, ln))
    print("   pixels with final_T >= 1e-2:", float(((Tp >= 1e-2) & inside).sum()) / inside.sum(), " mean n_contrib", nc.mean(), "p99", np.percentile(nc, 99))
    order = np.argsort(-trav)
    print(f"view {view}: tiles {len(ln)} R {ln.sum()} mean len {ln.mean():.0f} max len {ln.max()}  traversed: sum {trav.sum()} mean {trav.mean():.0f} "
          f"p50 {np.percentile(trav,50):.0f} p90 {np.percentile(trav,90):.0f} p99 {np.percentile(trav,99):.0f} max {trav.max()}  unsat tiles {unsat.sum()}")
    print("   heaviest tiles (traversed, len, unsat px):", [(int(trav[t]), int(ln[t]), int(((Tt[t] >= 1e-2) & it[t]).sum())) for t in order[:8]])
