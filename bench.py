#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native language-splatting hot path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W
    python bench.py --config 4 | --config 5                   (BASELINE.json configs[3] / configs[4], see below)

Headline workload (BASELINE.json metric / configs[2] shape): 1,000,000 Gaussians, 15-dim language features,
960x540, 8 synthetic keyframes per GPU and step.  A step is one mapping iteration of the reference on the hot path
(utils/slam_backend.py:499-757): encode the keyframes' 192x192x768 CLIP maps with the autoencoder, activate the
Gaussian parameters, render the 8 views (ONE batched forward), back-propagate them (ONE batched backward that writes
the summed per-Gaussian gradients into the flat buffer and the per-view pose gradients), fold the densification
statistics, all-reduce the flat gradient buffer + the side statistics over NCCL (N > 1), and take the fused Adam step
(activation Jacobians applied in the kernel) on every rank.

  value    frames/s with every input resident in HBM (CUDA-graph replay, CUDA events)
  e2e      frames/s through the public API (render_batch() + AutoencoderMLP.encode + losses.mapping_loss,
           loss.backward()) with the per-frame inputs (CLIP map, ground-truth RGB-D) copied from pinned host memory
           and the loss + the 15-dim code maps read back every step; default module flags (deferred overflow check)
  roofline the forward blend kernel named by the metric: algorithmic bytes / its mean launch time
           measured with CUDA events recorded inside the library over a second, eagerly launched timed region
  cpu_baseline / --impl reference: the CPU restatement of the reference (oracle/) + the reference AE on torch-CPU,
           one keyframe, all host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "lang-splat render+AE FPS @1M Gaussians/15-dim/960×540; blend HBM GB/s vs peak"
ENC_DIMS = [384, 192, 96, 48, 24, 15]
DEC_DIMS = [24, 48, 96, 192, 384, 384, 768]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--keyframes", type=int, default=8, help="keyframes per GPU per step")
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--height", type=int, default=540)
    ap.add_argument("--tile", type=int, default=15, help="15 = the reference build's tile geometry")
    ap.add_argument("--backward-mode", default="compat", choices=["compat", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-hr", action="store_true", help="skip the separate HR up-sampler timing block")
    ap.add_argument("--no-balance", action="store_true", help="N>1: keep contiguous blocks of views per rank (no cost balancing)")
    ap.add_argument("--no-graph", action="store_true", help="launch the resident step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--config", default="headline", choices=["headline", "4", "5"],
                    help="4 = AE-only encode+decode of 32 maps (BASELINE configs[3]); 5 = Replica-room0-shaped tracking+mapping loop (configs[4])")
    ap.add_argument("--per-view", action="store_true", help="headline: one forward/backward call per keyframe (the reference's call pattern) instead of one batched call")
    ap.add_argument("--tracking-iters", type=int, default=100, help="config 5: Training.tracking_itr_num")
    ap.add_argument("--mapping-iters", type=int, default=150, help="config 5: Training.mapping_itr_num")
    ap.add_argument("--kf-interval", type=int, default=4, help="config 5: Training.kf_interval")
    return ap.parse_args()


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 5 + i and r[5 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_frame_seconds(args, keyframes: int = 1):
    """One keyframe of the same workload on the host: oracle forward+backward + reference AE encode on torch-CPU."""
    import torch
    from online_lang_splatting_b200 import synthetic as S
    from online_lang_splatting_b200 import autoencoder as AE
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _util as U
    from oracle import oracle as O
    from oracle import torch_oracle as TO
    want = os.cpu_count() or 1
    O.set_num_threads(want)          # explicit: torchrun exports OMP_NUM_THREADS=1 to its workers
    cores = O.num_threads()
    torch.set_num_threads(max(cores, 1))
    sc = U.make_scene(P=args.gaussians, F=15, W=args.width, H=args.height, seed=0, scale=0.01)
    grads = U.loss_weights(15, args.width, args.height, seed=1)
    torch.manual_seed(0)
    ae = AE.AutoencoderMLP(ENC_DIMS, DEC_DIMS).eval()
    x = S.make_clip_maps(1, seed=0)
    t0 = time.perf_counter()
    for _ in range(keyframes):
        with torch.no_grad():
            TO.reference_chain(list(ae.encoder), x)
        U.run_oracle(sc, tile=args.tile, grads=grads, compat=(args.backward_mode == "compat"))
    dt = (time.perf_counter() - t0) / keyframes
    return dt, cores


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sec = []
    if args.warmup > 0:
        cpu_frame_seconds(args)
    for _ in range(max(1, min(args.steps, 3))):
        dt, cores = cpu_frame_seconds(args)
        sec.append(dt)
    dt = sum(sec) / len(sec)
    fps = 1.0 / dt
    sample = (f"{len(sec)} keyframe(s) of the same workload (P={args.gaussians}, F=15, {args.width}x{args.height}): CPU "
              f"restatement of the reference rasterizer forward+backward (oracle/ols_oracle.cpp, OpenMP) + reference "
              f"AutoencoderMLP.encode on torch-CPU; the reference ships no CPU render path (SURVEY 0.4)")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": len(sec), "warmup": 1 if args.warmup else 0, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample,
                             "cpu_model": cpu_model()},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    gpu_ref = compiled_reference_ms(args)
    if gpu_ref is not None:
        line["gpu_reference_cuda"] = gpu_ref
    emit(line)


def compiled_reference_ms(args):
    """Context only: the reference's own CUDA (oracle/_ref/ref_P_C.so, sm_100 build) on the same scene."""
    try:
        import torch
        if not torch.cuda.is_available():
            return None
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import _util as U
        mod = U.ref_module("ref_P_C")
        if mod is None:
            return None
        dev = torch.device("cuda:0")
        sc = U.make_scene(P=args.gaussians, F=15, W=args.width, H=args.height, seed=0, scale=0.01)
        grads = [g.to(dev) for g in U.loss_weights(15, args.width, args.height, seed=1)]
        d = lambda t: t.to(dev).contiguous()
        e = torch.Tensor([])
        a = (d(sc["bg"]), d(sc["means3D"]), e, d(sc["language"]), d(sc["opacities"]), d(sc["scales"]), d(sc["rotations"]),
             1.0, e, d(sc["viewmatrix"]), d(sc["projmatrix"]), d(sc["projmatrix_raw"]), sc["tanfovx"], sc["tanfovy"],
             args.height, args.width, d(sc["shs"]), 0, d(sc["campos"]), False, False)
        tf = tb = 0.0
        n = 5
        for it in range(n + 2):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            R, color, language, radii, geom, binning, img, depth, opacity, n_touched = mod.rasterize_language_gaussians(*a)
            e1.record()
            b = (a[0], a[1], radii, e, a[3], a[5], a[6], 1.0, e, a[9], a[10], a[11], a[12], a[13], grads[0], grads[1],
                 grads[2], a[16], 0, a[18], geom, R, binning, img, False)
            mod.rasterize_language_gaussians_backward(*b)
            e2.record()
            torch.cuda.synchronize()
            if it >= 2:
                tf += e0.elapsed_time(e1)
                tb += e1.elapsed_time(e2)
        return {"forward_ms": tf / n, "backward_ms": tb / n, "R": int(R),
                "note": "reference P/ CUDA (15x15 tiles, F=15) rebuilt for sm_100, one view, rasterizer only"}
    except Exception as ex:  # context only -- never fail the arm
        return {"error": str(ex)[:200]}


def workload_config(args):
    return {"workload": f"{args.gaussians} Gaussians, 15-dim language features, {args.width}x{args.height}, "
                        f"{args.keyframes} keyframes per GPU per step (one mapping iteration): AE encode of the 192x192x768 CLIP "
                        f"maps + parameter activation + render forward + backward of the keyframes + densification statistics "
                        f"+ NCCL all-reduce of the flat gradient buffer and the statistics (N>1) + fused Adam step",
            "gaussians": args.gaussians, "feature_dim": 15, "width": args.width, "height": args.height,
            "keyframes_per_gpu": args.keyframes, "view_assignment": "cost-balanced over ranks (sharding.balanced_views)" if (args.gpus > 1 and not args.no_balance) else "views rank*K .. rank*K+K-1", "tile": args.tile, "backward_mode": args.backward_mode,
            "autoencoder": "768-384-192-96-48-24-15 (1-stage, BN folded)", "parallelism": f"frames x{args.gpus}",
            "l2_policy": "per-step inputs (8 CLIP maps = 906 MB, 56 MB of Gaussian parameters + 64 MB records per view) exceed the 126 MB L2"}


# ------------------------------------------------------------------------------------------------ GPU arm
_OUT = None


def emit(line):
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


def _claim_stdout():
    """Everything except the final JSON line goes to stderr: libraries (NCCL prints its version banner on the
    first communicator) write to file descriptor 1 directly, so fd 1 is pointed at stderr for the whole run and
    the JSON line is written to a private duplicate of the original stdout."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


LR = {"xyz": 0.00016, "f_dc": 0.0025, "f_rest": 0.0025 / 20.0, "opacity": 0.05, "scaling": 0.001, "rotation": 0.001,
      "f_language": 0.0025}   # configs/rgbd/replicav2/base_config.yaml:78-88


def count_graph_kernels(graphs):
    """Kernel nodes of captured CUDA graphs (cudaGraphGetNodes / cudaGraphNodeGetType through libcudart): the launches
    one replay performs, counted rather than derived.  Returns None when the runtime cannot be reached."""
    import ctypes
    try:
        rt = None
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                rt = ctypes.CDLL(name)
                break
            except OSError:
                continue
        if rt is None:
            return None
        total = 0
        for g in graphs:
            if g is None:
                continue
            raw = ctypes.c_void_p(g.raw_cuda_graph())
            n = ctypes.c_size_t(0)
            if rt.cudaGraphGetNodes(raw, None, ctypes.byref(n)) != 0:
                return None
            nodes = (ctypes.c_void_p * n.value)()
            if rt.cudaGraphGetNodes(raw, nodes, ctypes.byref(n)) != 0:
                return None
            for i in range(n.value):
                t = ctypes.c_int(-1)
                if rt.cudaGraphNodeGetType(ctypes.c_void_p(nodes[i]), ctypes.byref(t)) == 0 and t.value == 0:  # cudaGraphNodeTypeKernel
                    total += 1
        return total
    except Exception:
        return None


def make_graph():
    import torch
    try:
        return torch.cuda.CUDAGraph(keep_graph=True)
    except TypeError:
        return torch.cuda.CUDAGraph()


def main():
    args = parse()
    global _OUT
    _OUT = _claim_stdout()
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm would silently run single-threaded
        n = os.cpu_count() or 1
        os.environ["OMP_NUM_THREADS"] = str(n)
        os.environ["MKL_NUM_THREADS"] = str(n)
        reference_arm(args)
        return
    if args.config == "4":
        from bench_configs import config4
        config4(args, emit, peaks, ClockSampler)
        return
    if args.config == "5":
        from bench_configs import config5
        config5(args, emit, peaks, ClockSampler)
        return
    import torch
    import torch.distributed as dist
    from online_lang_splatting_b200 import _native as N
    from online_lang_splatting_b200 import autoencoder as AE
    from online_lang_splatting_b200 import diff_gaussian_rasterization as dgr
    from online_lang_splatting_b200 import synthetic as S
    from online_lang_splatting_b200.gaussian_renderer import render_batch
    from online_lang_splatting_b200.losses import mapping_loss
    from online_lang_splatting_b200.optim import FlatAdam
    from online_lang_splatting_b200.sharding import FlatGradBuffer, FlatParams, SideStats
    import online_lang_splatting_b200.gaussian_renderer as GR

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N.require_cuda()

    P, W, H, KF = args.gaussians, args.width, args.height, args.keyframes
    GR.TILE_SIZE, GR.BACKWARD_MODE = args.tile, args.backward_mode
    g = S.make_gaussians(P, 15, W, H, seed=0, scale_px_sigma=0.01)
    pc = S.SyntheticGaussianModel(g, device=dev, requires_grad=True)     # e2e: the public-API model (autograd leaves)
    pipe = S.PipelineParams()
    bg = torch.zeros(3, device=dev)
    cams = [S.make_camera(W, H, view=rank * KF + k, seed=0, device=str(dev)) for k in range(KF)]
    torch.manual_seed(0)
    ae = AE.AutoencoderMLP(ENC_DIMS, DEC_DIMS).eval().to(dev)
    for p_ in ae.parameters():
        p_.requires_grad_(False)

    # per-keyframe inputs: one device-resident block of KF CLIP maps (value) and pinned host copies (e2e)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    clip_all = torch.randn(KF * 192 * 192, 768, device=dev, generator=gen)
    clip_all = clip_all / clip_all.norm(dim=-1, keepdim=True)
    clip_dev = [clip_all[k * 36864:(k + 1) * 36864] for k in range(KF)]
    clip_host, gt_host = [], []
    if not args.no_e2e:
        for k in range(KF):
            clip_host.append(clip_dev[k].cpu().pin_memory())
            gt_host.append((torch.rand(3, H, W).pin_memory(), (torch.rand(1, H, W) * 5).pin_memory()))
    gen_w = torch.Generator(device=dev).manual_seed(4321)   # the same image-space gradients on every rank (reduce_check recomputes other ranks' views)
    wc, wl, wd = (torch.randn(s, device=dev, generator=gen_w) for s in ((3, H, W), (15, H, W), (1, H, W)))

    # resident training state: raw parameters, activated copies, flat gradient buffer (the all-reduced one), Adam, side stats
    op = g["opacities"].clamp(1e-6, 1 - 1e-6)
    raw = {"means3D": g["means3D"], "sh": g["shs"][:, :1, :], "opacity": torch.log(op / (1 - op)), "scales": torch.log(g["scales"]),
           "rotations": g["rotations"], "language": g["language"]}
    fp = FlatParams({k_: v_.to(dev) for k_, v_ in raw.items()}, 15, 1, device=dev)
    # gradients [29 P] + the step's densification sums [2 P] in ONE buffer = ONE all-reduce (SUM) per step; max_radii2D is a
    # running maximum, which commutes with the cross-rank MAX, so it is reduced only when densification reads it
    fbuf = FlatGradBuffer(P, 15, 1, device=dev, extra=2 * P)
    flat = fbuf.flat
    opt = FlatAdam(fp.flat, fbuf.grads, fbuf.adam_groups(LR), capturable=True)   # step counter on the device: graph replays advance it
    stats = SideStats(P, device=dev, delta=fbuf.extra)
    act = fp.activate()
    out_bufs = fbuf.backward_outputs({"colors": torch.empty(P, 3, device=dev), "cov3D": torch.empty(P, 6, device=dev),
                                      "means2D": torch.empty(KF, P, 3, device=dev), "tau_sum": torch.empty(KF, 6, device=dev)})

    def settings_of(cam):
        return dgr.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), bg=bg,
            scale_modifier=1.0, viewmatrix=cam.world_view_transform, projmatrix=cam.full_proj_transform,
            projmatrix_raw=cam.projection_matrix, sh_degree=0, campos=cam.camera_center, prefiltered=False, debug=False,
            tile_size=args.tile, backward_mode=args.backward_mode)

    rs_list = [settings_of(cam) for cam in cams]
    rank_render_ms = None
    empty = torch.Tensor([])
    Rs = []
    view_ids = [rank * KF + k for k in range(KF)]

    def params_tuple():
        return (act["means3D"], act["sh"], empty, act["language"], act["opacity"], act["scales"], act["rotations"], empty)

    def phase_encode():
        """the step's autoencoder work: independent of the Gaussians (and therefore of the gradient all-reduce)"""
        with torch.no_grad():
            if args.per_view:
                for k in range(KF):
                    ae.encode(clip_dev[k])
            else:
                ae.encode(clip_all)

    def phase_render(views=None, out=None):
        """activate -> forward -> backward of this rank's keyframes, gradients written (not accumulated) into the flat
        buffer; densification statistics folded per view"""
        fp.activate()
        rl = rs_list if views is None else views
        ob = out_bufs if out is None else out
        stats.begin_step()
        if args.per_view:
            flat.zero_()
            for k in range(len(rl)):
                outs, st = dgr._forward_native_batch(*params_tuple(), [rl[k]])
                sub = dict(ob)
                sub["means2D"], sub["tau_sum"] = ob["means2D"][k], ob["tau_sum"][k]
                dgr._backward_native_batch(st, [outs[0][2]], [wc], [wl], [wd], out=sub, accumulate=True)
                stats.add_view(outs[0][2], ob["means2D"][k])
                Rs.extend(r_ for r_ in st.Rs if r_ >= 0)
            return
        outs, st = dgr._forward_native_batch(*params_tuple(), rl)
        radii = [o[2] for o in outs]
        V = len(rl)
        ob = dict(ob)
        ob["stats"] = stats.fused_outputs()      # densification statistics folded into the backward's geometry kernel
        dgr._backward_native_batch(st, radii, [wc] * V, [wl] * V, [wd] * V, out=ob, accumulate=False)
        Rs.extend(r_ for r_ in st.Rs if r_ >= 0)

    def phase_update():
        """after the reduce: identical Adam step and statistics update on every rank"""
        opt.step()
        stats.apply()

    def reduce_all():
        fbuf.all_reduce()      # gradients + accum / denom deltas (one NCCL all-reduce)

    def step_resident():
        phase_encode()
        phase_render()
        reduce_all()
        phase_update()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, with_marks=False):
        barrier()
        if with_marks:
            N.timing_begin(steps * KF * 16 + 64)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        marks = N.timing_end() if with_marks else None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, marks

    # ---- value: device-resident ----
    dgr.CHECK_OVERFLOW = "sync"     # warm-up: establishes the instance capacity of this configuration
    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    if world > 1 and not args.no_balance:
        # Keyframes differ in cost (how far a tile's list is traversed), and a step ends when the slowest rank does:
        # every rank times its own views once, the costs are gathered, and the world's views are re-dealt so that
        # the per-rank sums are even (sharding.balanced_views).  Each rank still renders KF distinct keyframes.
        from online_lang_splatting_b200.sharding import balanced_views
        cost = torch.zeros(KF, device=dev)
        for k in range(KF):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(2):
                outs, st = dgr._forward_native_batch(*params_tuple(), [rs_list[k]])
                sub = dict(out_bufs)
                sub["means2D"], sub["tau_sum"] = out_bufs["means2D"][k], out_bufs["tau_sum"][k]
                dgr._backward_native_batch(st, [outs[0][2]], [wc], [wl], [wd], out=sub, accumulate=False)
            a1.record()
            torch.cuda.synchronize()
            cost[k] = a0.elapsed_time(a1)
        all_cost = [torch.zeros(KF, device=dev) for _ in range(world)]
        dist.all_gather(all_cost, cost)
        flat_cost = torch.cat(all_cost).tolist()          # index = global view id (rank-major blocks)
        # refinement: the cost of a view inside a batched launch is not exactly its stand-alone cost.  Every rank measures
        # its BATCHED render time, the per-view costs of each rank are rescaled so that they add up to it, and the views
        # are dealt again; the assignment with the smallest slowest-rank time is kept (same data on every rank).
        assign = balanced_views(flat_cost, world)
        best = None
        for rnd in range(4):
            view_ids = assign[rank]
            cams[:] = [S.make_camera(W, H, view=v, seed=0, device=str(dev)) for v in view_ids]
            rs_list[:] = [settings_of(cam) for cam in cams]
            phase_render()
            torch.cuda.synchronize()
            t_best = 1e9
            for _ in range(3):
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                phase_render()
                a1.record()
                torch.cuda.synchronize()
                t_best = min(t_best, a0.elapsed_time(a1))
            tr = [torch.zeros(1, device=dev) for _ in range(world)]
            dist.all_gather(tr, torch.tensor([t_best], device=dev))
            times = [float(t_.item()) for t_ in tr]
            if best is None or max(times) < best[0]:
                best = (max(times), [list(a_) for a_ in assign], times)
            if max(times) <= 1.01 * (sum(times) / world) or rnd == 3:
                break
            for r_ in range(world):
                scale = times[r_] / max(sum(flat_cost[v] for v in assign[r_]), 1e-9)
                for v in assign[r_]:
                    flat_cost[v] *= scale
            assign = balanced_views(flat_cost, world)
        assign, rank_render_ms = best[1], best[2]
        view_ids = assign[rank]
        cams[:] = [S.make_camera(W, H, view=v, seed=0, device=str(dev)) for v in view_ids]
        rs_list[:] = [settings_of(cam) for cam in cams]
        sys.stderr.write(f"[balance rank {rank}] batched render ms per rank: {[round(t_, 3) for t_ in rank_render_ms]}\n")
        for _ in range(2):
            step_resident()
        torch.cuda.synchronize()
    # The step has no host synchronisation and fixed launch geometry, so its launches are captured once into CUDA
    # graphs and replayed (the collectives stay outside the graphs).  N = 1: one graph.  N > 1: three graphs -- the
    # AE encodes (independent of the Gaussians), activate + render forward + backward + statistics, and the Adam
    # step.  The all-reduce of step i runs on a communication stream while the encodes of step i + 1 replay; the
    # Adam graph of step i waits for it, and the render graph of step i + 1 follows the Adam graph.
    graph = graph_enc = graph_upd = None
    split = world > 1
    if not args.no_graph:
        try:
            if split:
                graph_enc = make_graph()
                with torch.cuda.graph(graph_enc):
                    phase_encode()
                graph = make_graph()
                with torch.cuda.graph(graph):
                    phase_render()
                graph_upd = make_graph()
                with torch.cuda.graph(graph_upd):
                    phase_update()
            else:
                graph = make_graph()
                with torch.cuda.graph(graph):
                    phase_encode()
                    phase_render()
                    phase_update()
            torch.cuda.synchronize()
        except Exception as ex:  # pragma: no cover -- fall back to eager launches, say so in the JSON line
            sys.stderr.write(f"CUDA graph capture failed ({ex}); timing eager launches\n")
            graph = graph_enc = graph_upd = None
            torch.cuda.synchronize()
    comm_stream = torch.cuda.Stream(device=dev, priority=-1) if split else None  # high priority: its CTAs are placed first
    comm_done = [None]

    def step_value():
        if graph is None:
            step_resident()
            return
        main_s = torch.cuda.current_stream(dev)
        if not split:
            graph.replay()
            return
        graph_enc.replay()                        # overlaps the previous step's all-reduce
        if comm_done[0] is not None:
            main_s.wait_event(comm_done[0])       # gradients + statistics of the previous step are reduced
            graph_upd.replay()                    # ... and applied (identical on every rank)
        graph.replay()
        ready = torch.cuda.Event()
        ready.record(main_s)
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(ready)
            reduce_all()
            ev = torch.cuda.Event()
            ev.record(comm_stream)
        comm_done[0] = ev

    def step_value_traced(tr):
        """step_value with CUDA events between the phases (diagnostic, N > 1 with graphs only)"""
        main_s = torch.cuda.current_stream(dev)
        def mark():
            e = torch.cuda.Event(enable_timing=True)
            e.record(main_s)
            return e
        t0 = mark()
        graph_enc.replay()
        t1 = mark()
        if comm_done[0] is not None:
            main_s.wait_event(comm_done[0])
            t2 = mark()
            graph_upd.replay()
        else:
            t2 = mark()
        t3 = mark()
        graph.replay()
        t4 = mark()
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(t4)
            c0 = torch.cuda.Event(enable_timing=True); c0.record(comm_stream)
            reduce_all()
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(comm_stream)
        comm_done[0] = ev
        tr.append((t0, t1, t2, t3, t4, c0, ev))

    def drain():
        if comm_done[0] is not None:
            torch.cuda.current_stream(dev).wait_event(comm_done[0])
            graph_upd.replay()
            comm_done[0] = None

    dgr.CHECK_OVERFLOW = "deferred"
    for _ in range(2):
        step_value()
    drain()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_value()
    drain()                                       # the last step's all-reduce and update are inside the timed region
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t_ = torch.tensor([ms], device=dev)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        ms = float(t_.item())
    launches_per_step = count_graph_kernels([graph, graph_enc, graph_upd]) if graph is not None else None
    phase_trace = None
    if split and graph is not None and os.environ.get("OLS_BENCH_TRACE"):
        tr = []
        for _ in range(6):
            step_value_traced(tr)
        drain()
        torch.cuda.synchronize()
        rows = []
        for (t0, t1, t2, t3, t4, c0, ev) in tr[1:]:
            rows.append({"encode": t0.elapsed_time(t1), "wait_for_reduce": t1.elapsed_time(t2), "adam_stats": t2.elapsed_time(t3),
                         "render": t3.elapsed_time(t4), "all_reduce": c0.elapsed_time(ev), "reduce_start_after_render": t4.elapsed_time(c0)})
        phase_trace = {k: sum(r[k] for r in rows) / len(rows) for k in rows[0]}
        sys.stderr.write(f"[phase trace rank {rank}] {phase_trace}\n")
    # second region, same K steps launched eagerly: CUDA events between the kernels give the per-kernel times
    Rs.clear()
    ms_eager, marks = timed(step_resident, args.steps, with_marks=True)
    clocks = sampler.stop() if rank == 0 else None
    dgr.CHECK_OVERFLOW = "sync"
    Rs.clear()
    step_resident()  # one checked step: proves no instance-capacity overflow happened with this capacity
    R_mean = sum(Rs) / max(len(Rs), 1)
    frames = world * KF * args.steps
    value = frames / (ms * 1e-3)

    # ---- the true-gradient configuration (16x16 tiles, exact backward) on the same keyframes: render forward + backward only ----
    exact16 = None
    if rank == 0 and not args.per_view and (args.tile, args.backward_mode) != (16, "exact"):
        try:
            rs_alt = [r_._replace(tile_size=16, backward_mode="exact") for r_ in rs_list]
            def step_alt():
                outs, st = dgr._forward_native_batch(*params_tuple(), rs_alt)
                dgr._backward_native_batch(st, [o[2] for o in outs], [wc] * KF, [wl] * KF, [wd] * KF, out=out_bufs, accumulate=False)
            dgr.CHECK_OVERFLOW = "sync"
            for _ in range(2):
                step_alt()
            dgr.CHECK_OVERFLOW = "deferred"
            torch.cuda.synchronize()
            n_alt = 5
            N.timing_begin(n_alt * 16 + 64)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(n_alt):
                step_alt()
            a1.record()
            torch.cuda.synchronize()
            m_alt = N.timing_end()
            exact16 = {"config": "tile 16x16, backward_mode exact (mathematically exact gradients, no reference quirks)",
                       "render_fwd_bwd_ms_per_view": a0.elapsed_time(a1) / n_alt / KF,
                       "ms_per_view": {t_: tot_ / max(cnt_, 1) / KF for t_, (tot_, cnt_) in m_alt.items() if cnt_}}
        except Exception as ex:  # pragma: no cover
            exact16 = {"error": str(ex)[:200]}
            dgr.CHECK_OVERFLOW = "deferred"

    # ---- multi-GPU correctness: the N-rank reduced buffer equals the 1-rank sum over the same world x KF views ----
    reduce_check = None
    if world > 1:
        phase_render()
        torch.cuda.synchronize()
        reduce_all()
        torch.cuda.synchronize()
        ids = [torch.zeros(KF, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(ids, torch.tensor(view_ids, dtype=torch.int64, device=dev))
        if rank == 0:
            reduced = fbuf.grads.clone()
            total = torch.zeros_like(reduced)
            side = FlatGradBuffer(P, 15, 1, device=dev)
            ob2 = side.backward_outputs({"colors": out_bufs["colors"], "cov3D": out_bufs["cov3D"], "means2D": out_bufs["means2D"],
                                         "tau_sum": out_bufs["tau_sum"]})
            for r_ in range(world):
                vs = [settings_of(S.make_camera(W, H, view=int(v), seed=0, device=str(dev))) for v in ids[r_].tolist()]
                outs, st = dgr._forward_native_batch(*params_tuple(), vs)
                dgr._backward_native_batch(st, [o[2] for o in outs], [wc] * KF, [wl] * KF, [wd] * KF, out=ob2, accumulate=False)
                total += side.grads
            torch.cuda.synchronize()
            num = float((reduced.double() - total.double()).norm())
            den = float(total.double().norm())
            reduce_check = {"views": world * KF, "rel_l2_reduced_vs_single_rank_sum": num / max(den, 1e-30),
                            "checksum_reduced": float(reduced.double().sum()), "checksum_single_rank": float(total.double().sum()),
                            "note": "float atomics make both sides order-dependent in the last bits; identical views, identical parameters"}
        dist.barrier()

    # ---- e2e: public API + host buffers, default module flags ----
    dgr.CHECK_OVERFLOW = "deferred"
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, dist, dev, world, KF, H, W, ae, cams, pc, pipe, bg, clip_host, gt_host, render_batch, mapping_loss,
                      timed, frames, make_graph)
        e2e["h2d_ceiling"] = h2d_ceiling(torch, dist, dev, world)
        e2e["transfer_floor_ms_per_step"] = e2e["h2d_bytes_per_step"] / (e2e["h2d_ceiling"]["per_rank_GBps_min"] * 1e9) * 1e3
        if not args.no_hr:
            # the same step fed by what the reference's backbone really produces on the GPU side
            from online_lang_splatting_b200 import supervised_net as SN
            torch.manual_seed(11)
            hr_net = SN.HighResLanguageFeatureNet().eval().to(dev)
            for p_ in hr_net.parameters():
                p_.requires_grad_(False)
            hr_host = [(torch.randn(1, 768, 24, 24).pin_memory(), torch.randn(1, 384, 96, 96).pin_memory(),
                        torch.randn(1, 192, 192, 192).pin_memory()) for _ in range(KF)]
            for p_ in pc.parameters():
                p_.grad = None
            e2e["hr_input_variant"] = run_e2e(args, torch, dist, dev, world, KF, H, W, ae, cams, pc, pipe, bg, clip_host, gt_host,
                                              render_batch, mapping_loss, timed, frames, make_graph, hr=(hr_net, hr_host))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- accuracy of the default (ex2.approx) blend next to the bit-exact one, on keyframe 0 ----
    fast_exp = None
    try:
        from online_lang_splatting_b200.debug import workspace_arrays
        dgr.CHECK_OVERFLOW = "sync"
        with torch.no_grad():
            oa, st_a = dgr._forward_native_batch(*params_tuple(), [rs_list[0]._replace(bitexact_blend=True)])
            nc_a = workspace_arrays(st_a)["n_contrib"].clone()
            ob_, st_b = dgr._forward_native_batch(*params_tuple(), [rs_list[0]])
            nc_b = workspace_arrays(st_b)["n_contrib"]
        rel = lambda a_, b_: float((a_ - b_).abs().max() / b_.abs().max().clamp_min(1e-30))
        fast_exp = {"rel_err_color": rel(ob_[0][0], oa[0][0]), "rel_err_language": rel(ob_[0][1], oa[0][1]),
                    "rel_err_depth": rel(ob_[0][3], oa[0][3]),
                    "pixels_with_different_n_contrib": int((nc_a != nc_b).sum().item()), "pixels": H * W, "budget": 1e-4,
                    "note": "default blend: alpha = min(0.99, o * ex2.approx(power * log2e)); bit-exact mode: expf as the reference build"}
    except Exception as ex:  # pragma: no cover
        fast_exp = {"error": str(ex)[:200]}

    # ---- roofline of the blend kernel (algorithmic bytes: SURVEY 8d) ----
    peak, peak_src = peaks()
    HW = W * H
    V_launch = 1 if args.per_view else KF     # views one launch of the rasterizer kernels covers
    per_kernel = {}
    alg1 = {"blend_fwd": 104 * R_mean + 88 * HW + 4 * P, "blend_bwd": 104 * R_mean + 84 * HW + 216 * P,
            "preprocess": 131 * P, "binning": 20 * P + 12 * R_mean + 8 * P, "sort": 24 * R_mean,
            "geometry_bwd": 230 * P}
    for tag, (tot, cnt) in (marks or {}).items():
        if cnt:
            avg_ms = tot / cnt
            ab = alg1[tag] * V_launch if tag in alg1 else (KF * 192 * 192 * (768 * 4 + 15 * 4) if tag == "ae" and not args.per_view
                                                           else (192 * 192 * (768 * 4 + 15 * 4) if tag == "ae" else None))
            per_kernel[tag] = {"ms_per_launch": avg_ms, "launches": cnt, "views_per_launch": V_launch if tag in alg1 else None,
                               "ms_per_view": avg_ms / V_launch if tag in alg1 else None,
                               "algorithmic_bytes_per_launch": ab,
                               "algorithmic_GBps": ab / (avg_ms * 1e-3) / 1e9 if ab else None}
    bf = per_kernel.get("blend_fwd", {"ms_per_launch": float("nan"), "algorithmic_GBps": float("nan")})
    roofline = {"kernel": "k_blend2 (forward alpha-blend, the kernel the metric names)", "bound": "hbm",
                "achieved": bf["algorithmic_GBps"], "peak": peak, "unit": "GB/s",
                "frac": (bf["algorithmic_GBps"] or 0.0) / peak, "traffic": traffic_from_profiles(V_launch),
                "peak_source": peak_src, "ms_per_launch": bf["ms_per_launch"], "views_per_launch": V_launch,
                "algorithmic_bytes_per_launch": alg1["blend_fwd"] * V_launch, "R_mean": R_mean,
                "note": "algorithmic bytes per view = 104*R + 88*H*W + 4*P (SURVEY 8d), times the views one launch covers; the "
                        "kernel stops at per-pixel saturation, so most of the R list is never read -- see DESIGN.md for the ncu "
                        "dram traffic and the instruction-issue roofline"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        dt, cores = cpu_frame_seconds(args)
        cpu = {"value": 1.0 / dt, "unit": "frames/s", "cores": cores, "kind": "port", "cpu_model": cpu_model(),
               "sample": "1 keyframe of the same workload: oracle forward+backward (OpenMP, all cores) + reference AE "
                         "encode on torch-CPU"}
    hr = None
    if not args.no_hr and rank == 0:
        hr = hr_module_timing(dev, ae)
    if launches_per_step is None:
        # eager fall-back: AE + activate + (preprocess, offsets, scan, scatter, 3 sorts, blend) + (blend bwd, geometry) + KF stats + adam
        launches_per_step = (1 + 1 + 8 + 2 + KF + 1) if not args.per_view else KF * (1 + 8 + 2 + 1) + 2
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (rasterizer, loss, Adam); autoencoder: tf32 first layer + bf16 inner layers, fp32 accumulate",
            "data": "synthetic", "config": workload_config(args),
            "roofline": roofline, "kernels": per_kernel, "cpu_baseline": cpu, "e2e": e2e,
            "hr_module": hr, "fast_exp_blend": fast_exp, "reduce_check": reduce_check, "exact_tile16": exact16,
            "rank_render_ms_after_balancing": rank_render_ms,
            "gpu_launches": args.steps * launches_per_step,
            "gpu_launches_per_step": launches_per_step,
            "gpu_launches_source": "kernel nodes of the replayed CUDA graphs (cudaGraphGetNodes)" if graph is not None else "launch list of the eager step",
            "clocks": clocks,
            "launch_mode": {"value": "cuda_graph_replay" if graph is not None else "eager", "ms_per_step_eager": ms_eager / args.steps,
                            "rasterizer_calls": "one call per keyframe" if args.per_view else f"one batched forward + one batched backward for the {KF} keyframes (grid.y = view)",
                            "collective": ("all-reduce of gradients + side statistics of step i on a communication stream, overlapped with the "
                                           "AE encodes of step i+1; Adam + render of step i+1 wait for it") if (split and graph is not None) else
                                          ("all-reduce after the backward" if world > 1 else "none (N=1)"),
                            "note": "per-kernel times and the roofline come from a second, eagerly launched region of the same "
                                    "K steps with CUDA events recorded between the kernels"}}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def h2d_ceiling(torch, dist, dev, world):
    """Pinned-host -> device copy bandwidth with ALL ranks copying at the same time (the ceiling of the e2e number when the
    step is transfer bound): 3 x 1 GiB per rank, barrier-bracketed, CUDA events."""
    n = 1 << 28
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device=dev)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbps = 3 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbps], device=dev)
    lo = t.clone()
    if world > 1:
        dist.all_reduce(t)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    del h, d
    return {"per_rank_GBps_min": float(lo.item()), "aggregate_GBps": float(t.item()), "ranks_copying_concurrently": world}


def run_e2e(args, torch, dist, dev, world, KF, H, W, ae, cams, pc, pipe, bg, clip_host, gt_host, render_batch, mapping_loss, timed,
            frames, make_graph, hr=None):
    """`hr`: None -> the host input of a keyframe is its 192x192x768 CLIP map (what the metric names); otherwise
    (net, [(fv, f3, f2) pinned host tensors per keyframe]) -> the host input is what the reference's SED backbone hands to
    the HR module (fv 24x24x768, res3, res2: 44 MB instead of 113 MB) and the map is produced on the device by
    AutoencoderMLP.encode_hr (HR up-sampler + encoder, final_conv folded into the first Linear)."""
    """One mapping iteration through the public API with host-resident inputs: H2D of every keyframe's CLIP map and
    ground-truth RGB-D from pinned memory, encode, render_batch, mapping_loss per view, ONE backward, all-reduce of the
    parameter gradients (N > 1), D2H of the code maps and the loss."""
    from online_lang_splatting_b200.sharding import FlatGradBuffer
    params = pc.parameters()
    copy_stream = torch.cuda.Stream(device=dev)
    code_host = [torch.empty(192 * 192, 15).pin_memory() for _ in range(KF)]
    # Double-buffered device staging: a step consumes slot s (filled while the previous step computed) and, at its very
    # start, queues the NEXT step's H2D copies into slot 1-s on the copy stream, so the PCIe link never idles and every
    # step still moves exactly one step's worth of inputs inside the timed region.
    if hr is None:
        bufs = [[(torch.empty(192 * 192, 768, device=dev), torch.empty(3, H, W, device=dev), torch.empty(1, H, W, device=dev))
                 for _ in range(KF)] for _ in range(2)]
    else:
        hr_net, hr_host = hr
        bufs = [[(tuple(torch.empty_like(t, device=dev) for t in hr_host[k]), torch.empty(3, H, W, device=dev),
                  torch.empty(1, H, W, device=dev)) for k in range(KF)] for _ in range(2)]

    def issue_copies(slot):
        with torch.cuda.stream(copy_stream):
            for k in range(KF):
                x, rgb, d = bufs[slot][k]
                if hr is None:
                    x.copy_(clip_host[k], non_blocking=True)
                else:
                    for dst, src in zip(x, hr_host[k]):
                        dst.copy_(src, non_blocking=True)
                rgb.copy_(gt_host[k][0], non_blocking=True)
                d.copy_(gt_host[k][1], non_blocking=True)

    def e2e_body(slot):
        cur = torch.cuda.current_stream(dev)
        copy_stream.wait_stream(cur)                                 # the readers of slot 1-s (previous step) are done
        issue_copies(1 - slot)                                       # next step's inputs, overlapped with this step's kernels
        outs = render_batch(cams, pc, pipe, bg)
        total = torch.zeros((), device=dev)
        for k in range(KF):
            x, gt_rgb, gt_d = bufs[slot][k]
            with torch.no_grad():
                if hr is None:
                    code = ae.encode(x)                              # [36864, 15] -> gt_lang_feat (slam_backend.py:557-576)
                else:
                    code = ae.encode_hr(hr_net, *x).view(-1, 15)     # slam_backend.py:381-395: hr_model(...) then encode
            code_host[k].copy_(code, non_blocking=True)              # the reference keeps it on the CPU (:576)
            gt_lang = code.t().reshape(15, 192, 192)
            loss = mapping_loss(outs[k]["render"], outs[k]["depth"], gt_rgb, gt_d, outs[k]["language"], gt_lang, alpha=0.95,
                                rgb_boundary_threshold=0.01, lambda_lang=1.0)
            total = total + loss
        total.backward()                                             # one backward for the window (slam_backend.py:670)
        cur.wait_stream(copy_stream)                                 # slot 1-s is complete when the next step starts
        return total.detach()

    state = {"graph": [None, None], "total": [None, None], "slot": 0}
    gbuf = None
    if world > 1:   # parameter gradients of the public model, flattened once so that the reduce is ONE collective
        gbuf = torch.zeros(sum(p_.numel() for p_ in params), device=dev)

    def step_e2e():
        slot = state["slot"]
        state["slot"] = 1 - slot
        if state["graph"][slot] is not None:
            state["graph"][slot].replay()
            total = state["total"][slot]
        else:
            total = e2e_body(slot)
        if world > 1:
            o = 0
            for p_ in params:
                if p_.grad is not None and p_.numel():
                    gbuf[o:o + p_.numel()].copy_(p_.grad.reshape(-1))
                o += p_.numel()
            dist.all_reduce(gbuf)
        val = float(total.item())                                     # D2H read of the step's result
        if state["graph"][slot] is None:
            for p_ in params:
                p_.grad = None
        return val

    issue_copies(0)                                                   # prime the first slot (outside the timed region)
    torch.cuda.synchronize()
    for _ in range(2):
        step_e2e()
    if not args.no_graph:
        for p_ in params:
            p_.grad = None
        torch.cuda.synchronize()
        try:
            for slot in (0, 1):   # gradients land in the same .grad tensors in both graphs (captured back to back)
                g_ = make_graph()
                with torch.cuda.graph(g_):
                    tot = e2e_body(slot)
                state["graph"][slot], state["total"][slot] = g_, tot
                if slot == 0:
                    for p_ in params:
                        p_.grad = None
        except Exception as ex:  # pragma: no cover
            sys.stderr.write(f"e2e CUDA graph capture failed ({ex}); timing eager calls\n")
            state["graph"] = [None, None]
            torch.cuda.synchronize()
            for p_ in params:
                p_.grad = None
    for _ in range(2):
        step_e2e()
    ms_e, _ = timed(step_e2e, args.steps)
    per_kf = 192 * 192 * 768 * 4 if hr is None else sum(t.numel() * 4 for t in hr_host[0])
    h2d = KF * (per_kf + 3 * H * W * 4 + H * W * 4)
    d2h = KF * (192 * 192 * 15 * 4) + 4
    return {"value": frames / (ms_e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "ms_per_step": ms_e / args.steps,
            "api": ("gaussian_renderer.render_batch() + AutoencoderMLP.encode() + losses.mapping_loss(), one loss.backward(); default module flags"
                    if hr is None else
                    "gaussian_renderer.render_batch() + AutoencoderMLP.encode_hr(HighResLanguageFeatureNet, fv, res3, res2) + losses.mapping_loss(), one loss.backward()"),
            "host_input": "192x192x768 fp32 CLIP map per keyframe (113 MB)" if hr is None else "fv 24x24x768 + res3 96x96x384 + res2 192x192x192 per keyframe (44 MB): the HR module runs on the device",
            "h2d_GBps_achieved": h2d / (ms_e / args.steps * 1e-3) / 1e9,
            "pipelining": "the H2D copies of step i+1 run on a copy stream under the kernels of step i (two staging slots); one step's "
                          "inputs cross the link per step inside the timed region",
            "launch_mode": "cuda_graph_replay" if state["graph"][0] is not None else "eager"}


def hr_module_timing(dev, ae=None, iters=20):
    """The HR up-sampler that produces the autoencoder's input when `hr_model` is on: fv 24x24 ->
    192x192x768, random weights, timed alone with CUDA events OUTSIDE the benchmark's timed region (the headline
    metric is quoted on random CLIP maps, i.e. without this stage).  Tensor roofline: 104.9 GFLOP per frame against
    the measured dense bf16 peak."""
    import torch
    from online_lang_splatting_b200 import supervised_net as SN
    torch.manual_seed(11)
    net = SN.HighResLanguageFeatureNet().eval().to(dev)
    fv = torch.randn(1, 768, 24, 24, device=dev)
    f3 = torch.randn(1, 384, 96, 96, device=dev)
    f2 = torch.randn(1, 192, 192, 192, device=dev)
    with torch.no_grad():
        for _ in range(3):
            net(fv, f3, f2)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            net(fv, f3, f2)
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    fused = None
    if ae is not None:
        # HR + autoencoder encode of the same frame: the two calls the reference makes vs final_conv folded into the encoder
        def timed(fn):
            with torch.no_grad():
                for _ in range(3):
                    fn()
                torch.cuda.synchronize(dev)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(iters):
                    fn()
                b.record()
                torch.cuda.synchronize(dev)
            return a.elapsed_time(b) / iters
        fused = {"hr_then_encode_ms": timed(lambda: ae.encode(net(fv, f3, f2).permute(0, 2, 3, 1).view(-1, 768))),
                 "encode_hr_fused_ms": timed(lambda: ae.encode_hr(net, fv, f3, f2))}
    gflop = 104.9
    peak = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["bf16_tflops"])
    except Exception:
        peak = 2250.0
    tf = gflop / ms
    return {"ms_per_frame": ms, "frames_per_s": 1000.0 / ms, "gflop_per_frame": gflop, "achieved_tflops": tf,
            "peak_tflops": peak, "frac": tf / peak, "kernels_per_frame": 16, "with_autoencoder": fused, "dtype": "bf16 activations, fp32 accumulate",
            "note": "13 tcgen05 implicit-GEMM convolutions + 3 input conversions, eager launches with programmatic dependent launch"}


def traffic_from_profiles(views_per_launch=1):
    """dram bytes per launch of the forward blend from the committed ncu --set full capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "blend_fwd_dram_bytes.json")) as f:
            d = json.load(f)
        if int(d.get("views_per_launch", 1)) == int(views_per_launch):
            return d["dram_bytes_per_launch"]
        return d["dram_bytes_per_launch"] / d.get("views_per_launch", 1) * views_per_launch
    except Exception:
        return None


if __name__ == "__main__":
    main()
