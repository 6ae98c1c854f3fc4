#!/usr/bin/env python
"""Benchmark of the BeNeRF N-pose blur + event render (BASELINE.json metric: rays/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at N = 1: BASELINE.json configs[1] -- benerf_unreal RGB + events, 768x480 intrinsics, 19
virtual poses across the exposure (blur model) + 2 poses of an event window, coarse + fine 64 + 128
samples per ray, C = 3 -- in its throughput shape: R = 65,536 pixels, i.e. 21 * R =
1,376,256 rays per step.  One step = spline poses -> two Graph.render calls -> blur mean + event
log-difference for the fine and coarse levels -> the four loss terms.  Synthetic pixels, Xavier
weights (init_nerf), knots rand(4,6)*0.01; in-kernel Philox draws (production mode).

N > 1: pixels shard across ranks (every rank renders its own R pixels at all poses with a full
weight replica, so image formation stays local); one NCCL all-reduce of the four partial loss sums
per step is the only exchange.  Weak scaling; value = all rays / max-over-ranks device time.

--impl reference times the CPU oracle (oracle/, the restatement of the reference's PyTorch path
pinned to its outputs; the Python reference itself cannot travel to the GPU box) on the host cores,
on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time
from argparse import Namespace

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, FOCAL, CX, CY = 480, 768, 548.409, 384.0, 240.0       # configs/benerf_unreal/livingroom.txt:12-17
N_POSES, S_C, N_I, CH = 19, 64, 64, 3
EXPOSURE, WINDOW = (0.2, 0.8), (0.3, 0.4)
MACS_PER_SAMPLE = 593_408                                   # SURVEY 8-d, C = 3 (the reference's linears)
# MACs the CTA-pair kernel actually multiplies: feature_linear (65,536) and the feature block of views_linears.0 (32,768) run
# as ONE merged 256 -> 128 linear (32,768); the 27 view channels (3,456) and the two heads (256 + 384) are per-ray / FFMA work
MACS_ISSUED_PER_SAMPLE = 63 * 256 + 4 * 65536 + 319 * 256 + 2 * 65536 + 256 * 128
FLOP_PER_RAY = (S_C + S_C + N_I) * 2 * MACS_PER_SAMPLE      # 227,868,672
K_MAT = [[FOCAL, 0.0, CX], [0.0, FOCAL, CY], [0.0, 0.0, 1.0]]


def ref_args():
    return Namespace(dataset="BeNeRF_Unreal", channels=CH, N_samples=S_C, N_importance=N_I, multires=10, multires_views=4,
                     i_embed=0, use_viewdirs=True, use_barf_c2f=False, ndc=True, traj="spline", num_interpolated_pose=N_POSES,
                     rgb_crf_net_hidden=0, rgb_crf_net_width=128, event_crf_net_hidden=0, event_crf_net_width=128,
                     chunk=4096, event_threshold=0.1, event_coeff_syn=0.1, rgb_coeff=1.0, seed=0)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json: sustained bf16 TFLOP/s, copy GB/s)"
    except Exception:
        return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(flop_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r02_mlp_tc3_ncu_bench.json, written by tools/r02_profiles.py from tools/gpu_r02_final.sh), scaled by algorithmic work to this run's
    average launch (the kernel's traffic is proportional to its rows); None when no capture is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_mlp_tc3_ncu_bench.json")) as f:
            cap = json.load(f)
        return {"bytes_per_launch": cap["dram_bytes"] * flop_per_launch / cap["algorithmic_flop"],
                "over_algorithmic_bytes": cap["dram_bytes"] / cap["algorithmic_bytes"],
                "source": "ncu --set full, %s, scaled from a launch of %.3g algorithmic FLOP" % (cap["kernel"], cap["algorithmic_flop"])}
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:           # NVML unavailable: report that instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------
def run_cpu_oracle(n_pixels, repeats=1):
    """`repeats` steps of the same workload on the host CPU with the oracle; returns (rays/s, rays per step, mean seconds
    per step).  The oracle evaluates a step's samples in one batch, so the sample is bounded per step and repeated."""
    from oracle import pose, render as orender, image_formation as oif, mlp as omlp
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(0)
    coarse, fine = omlp.xavier_params(CH, g), omlp.xavier_params(CH, g)
    knots = torch.rand(4, 6, generator=g) * 0.01
    transform = torch.zeros(1, 6)
    idx_evt = torch.randint(0, H * W, (n_pixels,), generator=g)
    idx_rgb = torch.randint(0, H * W, (n_pixels,), generator=g)
    K = torch.tensor(K_MAT, dtype=torch.float32)
    best = None
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            p_evt = pose.poses_from_knots(knots, None, *WINDOW, 2)
            p_rgb = pose.poses_from_knots(knots, transform, *EXPOSURE, N_POSES)
            r_evt = orender.render(coarse, fine, p_evt, idx_evt, H, W, K, orender.draw_rng(2 * n_pixels, S_C, N_I, g))
            r_rgb = orender.render(coarse, fine, p_rgb, idx_rgb, H, W, K, orender.draw_rng(N_POSES * n_pixels, S_C, N_I, g))
            for lvl in ("rgb_map", "rgb0"):
                oif.blur_mean(r_rgb[lvl], N_POSES)
                oif.event_log_diff(r_evt[lvl], "BeNeRF_Unreal", CH)
            dt = time.perf_counter() - t0
            best = dt if best is None else best + dt
    rays = (N_POSES + 2) * n_pixels
    return rays / (best / repeats), rays, best / repeats


def staged_reference():
    """The unmodified reference staged by build() under baseline/_ref (tools/stage_reference.py), or None."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        from stage_reference import import_staged
        return import_staged()
    except Exception as e:                       # a missing third-party import on this box: fall back to the port, say why
        print(f"staged reference not importable ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)
        return None
    finally:
        sys.path.pop(0)


_REF_STATE = {}


def run_cpu_reference(ref, n_pixels, repeats=1):
    """`repeats` steps of the workload through the REFERENCE's own code on the host CPU: model/optimize.py Model/Graph, get_pose_evt /
    get_pose_rgb (spline.py), Graph.render x2 (model/nerf.py:236-343: rays, embedder, NeRF.forward coarse + sample_pdf + fine,
    raw2output), then the blur mean (train.py:307-318) and the event log-difference with the reference's RGB2Gray /
    rgb2brightlog.  Forward only, under no_grad, all host threads.  Returns (rays/s, rays per step, mean seconds per step)."""
    if "graph" not in _REF_STATE:
        torch.manual_seed(0)
        args = ref_args()
        args.max_iter, args.barf_c2f_start, args.barf_c2f_end = 80000, 0.1, 0.5
        graph = ref.optimize.Model(args).build_network(args)
        ref.helpers.init_nerf(graph.nerf)
        ref.helpers.init_nerf(graph.nerf_fine)
        _REF_STATE.update(graph=graph, args=args, gray=ref.img_utils.RGB2Gray())
    graph, args, rgb2gray = _REF_STATE["graph"], _REF_STATE["args"], _REF_STATE["gray"]
    g = torch.Generator().manual_seed(0)
    idx_evt = torch.randint(0, H * W, (n_pixels,), generator=g)
    idx_rgb = torch.randint(0, H * W, (n_pixels,), generator=g)
    K = torch.tensor(K_MAT, dtype=torch.float32)
    total = 0.0
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            p_evt = graph.get_pose_evt(args, torch.tensor(WINDOW, dtype=torch.float32))
            p_rgb = graph.get_pose_rgb(args, torch.tensor(EXPOSURE, dtype=torch.float32))
            r_evt = graph.render(0, p_evt, idx_evt, H, W, K, args, enable_crf=True, sensor_type="event", remap=None, training=True)
            r_rgb = graph.render(0, p_rgb, idx_rgb, H, W, K, args, enable_crf=True, sensor_type="rgb", remap=None, training=True)
            for lvl in ("rgb_map", "rgb0"):
                blur = 0
                for j in range(N_POSES):
                    blur = blur + r_rgb[lvl][j * n_pixels:(j + 1) * n_pixels]
                blur = blur / N_POSES
                b1 = ref.math_utils.rgb2brightlog(rgb2gray(r_evt[lvl][:n_pixels]), args.dataset)
                b2 = ref.math_utils.rgb2brightlog(rgb2gray(r_evt[lvl][n_pixels:]), args.dataset)
                _ = b2 - b1
            total += time.perf_counter() - t0
    rays = (N_POSES + 2) * n_pixels
    return rays / (total / repeats), rays, total / repeats


def run_cpu_arm(n_pixels, repeats=1):
    """The CPU arm of the bench: the staged reference when present (kind "reference"), else the oracle port (kind "port")."""
    ref = staged_reference()
    if ref is not None:
        return run_cpu_reference(ref, n_pixels, repeats) + ("reference", "unmodified reference (baseline/_ref: model/nerf.py Graph.render, spline.py, run_nerf_helpers.py), torch CPU fp32")
    return run_cpu_oracle(n_pixels, repeats) + ("port", "oracle/ (torch CPU fp32 restatement, pinned to the reference's outputs)")


def run_cpu_oracle_train(n_evt, n_rgb):
    """Forward + backward of one training iteration (configs[2] shape, scaled down) on the host CPU with the oracle under
    torch autograd -- what train.py:160-340 costs per ray on the reference's CPU path.  Returns (rays/s, rays, seconds)."""
    from oracle import pose, render as orender, image_formation as oif, mlp as omlp
    g = torch.Generator().manual_seed(0)
    Ht = Wt = 800
    K = torch.tensor([[1111.111, 0.0, 400.0], [0.0, 1111.111, 400.0], [0.0, 0.0, 1.0]])
    coarse = {k: v.requires_grad_(True) for k, v in omlp.xavier_params(CH, g).items()}
    fine = {k: v.requires_grad_(True) for k, v in omlp.xavier_params(CH, g).items()}
    knots = (torch.rand(4, 6, generator=g) * 0.01).requires_grad_(True)
    transform = torch.zeros(1, 6, requires_grad=True)
    idx_evt = torch.randint(0, Ht * Wt, (n_evt,), generator=g)
    idx_rgb = torch.randint(0, Ht * Wt, (n_rgb,), generator=g)
    accu = torch.randint(-3, 4, (Ht, Wt), generator=g).double()
    blur_t = torch.rand(n_rgb, CH, generator=g)
    t0 = time.perf_counter()
    p_evt = pose.poses_from_knots(knots, None, *WINDOW, 2)
    p_rgb = pose.poses_from_knots(knots, transform, *EXPOSURE, N_POSES)
    r_evt = orender.render(coarse, fine, p_evt, idx_evt, Ht, Wt, K, orender.draw_rng(2 * n_evt, S_C, N_I, g))
    r_rgb = orender.render(coarse, fine, p_rgb, idx_rgb, Ht, Wt, K, orender.draw_rng(N_POSES * n_rgb, S_C, N_I, g))
    loss, _ = oif.training_loss(r_evt, r_rgb, accu, idx_evt, blur_t, n_poses=N_POSES, dataset="E2NeRF_Synthetic", channels=CH, threshold=0.2)
    loss.backward()
    dt = time.perf_counter() - t0
    rays = 2 * n_evt + N_POSES * n_rgb
    return rays / dt, rays, dt


def bench_reference(opts):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_pixels = opts.cpu_pixels
    for _ in range(max(1, min(opts.warmup, 2))):   # untimed warm-up passes (thread pools, allocator)
        run_cpu_arm(max(8, n_pixels // 4))
    t = []
    rays, kind, what = 0, "port", ""
    for _ in range(opts.steps):
        rps, rays, dt, kind, what = run_cpu_arm(n_pixels)
        t.append(dt)
    ms = 1e3 * sum(t) / len(t)
    value = rays / (ms / 1e3)
    sample = f"{n_pixels} pixels x ({N_POSES}+2) poses = {rays} rays per step (the GPU arm renders 65,536 pixels per step: same poses, samples and networks, bounded pixel count), {what}"
    print(json.dumps({
        "impl": "reference", "metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": opts.gpus, "steps": opts.steps,
        "warmup": opts.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        # the SAME config as our arm (the workload the number speaks about); the bounded sample each step actually renders is stated
        # in cpu_baseline.sample / sample_rays_per_step -- a full 1.38 M-ray step would take the host about seven minutes
        "config": workload_config(opts.pixels, opts.gpus), "sample_rays_per_step": rays,
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n_pixels, n_gpus, note=None):
    c = {"workload": "benerf_unreal RGB+events 768x480, 19 blur poses + 2 event poses, coarse+fine 64+128 samples, C=3 "
                     "(BASELINE.json configs[1], throughput shape)",
         "pixels_per_gpu": n_pixels, "rays_per_step_per_gpu": (N_POSES + 2) * n_pixels, "n_poses": N_POSES,
         "samples": [S_C, S_C + N_I], "parallelism": f"pixel-sharded x{n_gpus}, weights replicated",
         "l2": "no flush needed: per-step workspace + outputs >> 126 MB L2 (1.8 KB/ray scratch: rays, view biases, sample depths)"}
    if note:
        c["note"] = note
    return c


# ----------------------------------------------------------------------------------------------
def bench_image_formation_stress(dev, steps=5, warmup=3):
    """BASELINE.json configs[4]: the image-formation stage standalone at 1920x1080 on pre-rendered frames:
    blur mean over 51 poses, log-intensity differences of 257 frames (256 event bins), scatter of 1e7 events.  HBM-bound;
    inputs (1.3 GB / 6.4 GB) are far larger than the 126 MB L2, so no flush is needed between iterations."""
    from benerf_b200 import image_formation as IF
    R, P, B = 1920 * 1080, 51, 256
    _, hbm_peak, src = peaks()
    frames = torch.rand(B + 1, R, 3, device=dev)
    out = {}

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    with torch.no_grad():
        blur_in = frames[:P].reshape(P * R, 3)
        ms = timed(lambda: IF.blur_mean(blur_in, P))
        nbytes = 4 * R * 3 * (P + 1)
        out["blur_mean"] = {"ms": ms, "bytes": nbytes, "achieved_gbs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / hbm_peak}
        ev_in = frames.reshape((B + 1) * R, 3)
        ms = timed(lambda: IF.event_logdiff(ev_in, B, "BeNeRF_Unreal"))
        nbytes = 4 * R * (3 * (B + 1) + B)
        out["event_logdiff"] = {"ms": ms, "bytes": nbytes, "achieved_gbs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / hbm_peak}
        # event scatter: E = 1e7 time-sorted events into 256 bins x [1080, 1920] float64 (4.2 GB of images) in ONE launch.  The
        # 120 MB of one event set would stay in the 126 MB L2 across repeats, so every pass reads a different one of SETS sets
        # (480 MB): the input side comes from HBM, and so do the read-modify-writes into the bin images.
        E, SETS = 10_000_000, 4
        g = torch.Generator(device=dev).manual_seed(0)
        sets = []
        for _ in range(SETS):
            sets.append((torch.randint(0, 1920, (E,), device=dev, generator=g, dtype=torch.int32),
                         torch.randint(0, 1080, (E,), device=dev, generator=g, dtype=torch.int32),
                         (torch.randint(0, 2, (E,), device=dev, generator=g) * 2 - 1).float()))
        bounds = torch.linspace(0, E, B + 1, device=dev).round().to(torch.int64)
        img = torch.zeros(B, 1080, 1920, device=dev, dtype=torch.float64)
        turn = [0]

        def scatter():
            x, y, pol = sets[turn[0] % SETS]
            turn[0] += 1
            IF.accumulate_events_binned(x, y, pol, bounds, 1080, 1920, out=img)

        ms = timed(scatter)
        total = float(img.sum())                                  # every pass adds sum(pol) of its set: a checksum of all passes
        want = sum(float(sets[i % SETS][2].sum()) for i in range(turn[0]))
        out["event_scatter"] = {"ms": ms, "events": E, "bins": B, "events_per_s": E / ms * 1e3, "input_read_gbs": 12 * E / ms / 1e6,
                                "atomic_sector_gbs": 2 * 32 * E / ms / 1e6, "input_sets": SETS, "checksum_ok": total == want,
                                "note": "inputs rotate over 4 sets (480 MB > L2); one 32 B sector read + written per event"}
        del img, sets
    out["peak_gbs"], out["peak_source"] = hbm_peak, src
    out["workload"] = "1920x1080, 51 blur poses, 257 frames -> 256 event bins, 1e7 events (BASELINE.json configs[4], per GPU)"
    del frames
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------
def bench_train_step(opts, dev, world, rank, steps=5, warmup=3, config="e2nerf_synthetic", scaling="weak"):
    """Training step per GPU (weak scaling), forward + backward + Adam + ONE gradient all-reduce
    (benerf_b200.train.Trainer.step), device-timed, max over ranks:
      e2nerf_synthetic  BASELINE.json configs[2]: 800x800, 1024 event pixels x 2 poses + 2048 // 19 = 107 blur pixels x 19
                        poses = 4081 rays, thresholded event loss
      e2nerf_real       BASELINE.json configs[3]: 346x260, 31 virtual poses, 1024 event pixels x 2 + 1024 // 31 = 33 blur pixels
                        x 31 = 3071 rays, NORMALISED event loss (its two batch norms are all-reduced, train.py:238-292)"""
    import torch.distributed as dist
    from benerf_b200 import optimize, run_nerf_helpers
    from benerf_b200.train import Trainer
    args = ref_args()
    if config == "e2nerf_real":
        Ht, Wt, f, n_poses, r_rgb = 260, 346, 653.98456, 31, 1024 // 31
        Kt = [[f, 0.0, 173.0], [0.0, f, 130.0], [0.0, 0.0, 1.0]]
        args.dataset, args.event_threshold, args.seed, args.num_interpolated_pose = "E2NeRF_Real", -1.0, 1, n_poses
        args.event_coeff_real = 2.0
    else:
        Ht = Wt = 800
        f, n_poses, r_rgb = 1111.111, N_POSES, 2048 // N_POSES
        Kt = [[f, 0.0, 400.0], [0.0, f, 400.0], [0.0, 0.0, 1.0]]
        args.dataset, args.event_threshold, args.seed = "E2NeRF_Synthetic", 0.2, 1
    args.lrate, args.pose_lrate, args.transform_lrate, args.rgb_crf_lrate, args.event_crf_lrate = 5e-4, 1e-3, 1e-6, 5e-4, 5e-4
    args.optimize_nerf, args.optimize_pose, args.optimize_trans = True, True, False
    args.fused_optimizer = os.environ.get("BNRF_FUSED_TAIL", "1") != "0"
    if os.environ.get("BNRF_GEMM_MODE"):
        args.gemm_mode = os.environ["BNRF_GEMM_MODE"]          # development aid: tc | tc_chain1 | tc_linear
    torch.manual_seed(0)
    model = optimize.Model(args)
    graph = model.build_network(args)
    run_nerf_helpers.init_nerf(graph.nerf)
    run_nerf_helpers.init_nerf(graph.nerf_fine)
    graph.to(dev)
    trainer = Trainer(model, args)
    if os.environ.get("BNRF_STEP_TRACE"):
        trainer.phase_ms = []
    g = torch.Generator().manual_seed(99 + rank)
    r_evt = 1024
    if scaling == "strong":                 # ONE reference batch split by pixel over the ranks
        r_evt, r_rgb = r_evt // world, max(r_rgb // world, 1)
    idx_evt = torch.randint(0, Ht * Wt, (r_evt,), generator=g).to(dev)
    idx_rgb = torch.randint(0, Ht * Wt, (r_rgb,), generator=g).to(dev)
    blur_t = torch.rand(r_rgb, CH, generator=g).to(dev)
    accu = torch.randint(-3, 4, (Ht, Wt), generator=g).double().to(dev)
    ts_evt, ts_rgb = torch.tensor(WINDOW), torch.tensor(EXPOSURE)

    def one():
        return trainer.step(accu, idx_evt, idx_rgb, blur_t, ts_evt, ts_rgb, Ht, Wt, Kt, Kt)

    for _ in range(warmup):
        loss, _ = one()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    eng = graph.engine(args)
    eng.profile(True)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        loss, _ = one()
        evs[i + 1].record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = evs[0].elapsed_time(evs[-1]) / steps
    per_step = [round(evs[i].elapsed_time(evs[i + 1]), 2) for i in range(steps)]
    launches = eng.profile_read()["launches"]
    eng.profile(False)
    if trainer.launches_per_step is not None:            # the iteration is a replayed CUDA graph: count the kernels it holds
        launches = trainer.launches_per_step * steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if trainer.phase_ms:
        print("phase trace:", trainer.phase_ms, file=sys.stderr)
    rays = (2 * r_evt + n_poses * r_rgb) * world
    return {"metric": "rays_per_sec (training step: forward + backward + Adam + gradient all-reduce)", "value": rays / (ms / 1e3),
            "unit": "rays/s", "ms_per_step": ms, "steps": steps, "warmup": warmup, "rays_per_step_per_gpu": rays // world,
            "workload": (f"{config} {Wt}x{Ht}, {r_evt} event px x 2 poses + {r_rgb} blur px x {n_poses} poses per GPU, 64+128 samples, C=3 "
                         f"(BASELINE.json configs[{3 if config == 'e2nerf_real' else 2}], {scaling} scaling"
                         + (", normalised event loss with all-reduced batch norms)" if config == "e2nerf_real" else ")")),
            "algorithmic_tflops": 3 * rays / world * FLOP_PER_RAY / (ms / 1e3) / 1e12,
            "ms_each_step": per_step, "gpu_launches_per_step": launches / steps, "final_loss": float(loss),
            "mem_allocated_gb": round(torch.cuda.max_memory_allocated(dev) / 2**30, 1),
            "optimizer_tail": "fused (bnrf_adam_step_sched)" if args.fused_optimizer else "torch.optim.Adam x3",
            "cuda_graph": trainer._cg is not None,
            "backward": "fused loss + analytic render gradients (loss.cu), heads_fused_kernel, tcgen05 dgrad chain on CTA pairs with the last "
                        "K-block issued in N-halves (dgrad_chain2.cu), weight gradients of the eight 256-wide layers + view layer on CTA "
                        "pairs (wgrad_pair.cu) and of the two encoded-point blocks on single CTAs (bwd_tiles.cu); bf16 hi/lo tile matrices, "
                        "3 MMAs per product, fp32 accumulate; whole step replayed as one CUDA graph"}


# ----------------------------------------------------------------------------------------------
def bench_ours(opts):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: benerf_b200 has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from benerf_b200 import optimize, run_nerf_helpers, image_formation as IF

    args = ref_args()
    args.mlp_mode = opts.mlp_mode
    torch.manual_seed(0)
    graph = optimize.Model(args).build_network(args)
    run_nerf_helpers.init_nerf(graph.nerf)
    run_nerf_helpers.init_nerf(graph.nerf_fine)
    graph.to(dev)
    eng = graph.engine(args)
    if opts.mode == "train":                 # development aid: only the training-step measurement
        t = bench_train_step(opts, dev, world, rank, steps=opts.steps, warmup=opts.warmup, config=opts.train_config, scaling=opts.train_scaling)
        if rank == 0:
            print(json.dumps(t))
        if world > 1:
            dist.destroy_process_group()
        return
    # the two secondary measurements run first, on a quiet device (the headline render phase holds ~12 GB and the power cap)
    train = None if opts.no_train_step else bench_train_step(opts, dev, world, rank)
    train_real = None if opts.no_train_step else bench_train_step(opts, dev, world, rank, config="e2nerf_real")
    train_strong = None if (opts.no_train_step or world == 1) else bench_train_step(opts, dev, world, rank, scaling="strong")
    stress = None if opts.no_train_step else bench_image_formation_stress(dev)
    torch.cuda.empty_cache()
    R = opts.pixels
    g = torch.Generator().manual_seed(1234 + rank)
    host_idx_evt = torch.randint(0, H * W, (R,), generator=g).pin_memory()
    host_idx_rgb = torch.randint(0, H * W, (R,), generator=g).pin_memory()
    host_target_blur = torch.rand(R, CH, generator=g).pin_memory()
    host_target_evt = (torch.randint(-3, 4, (R, 1), generator=g).double()).pin_memory()
    dev_inputs = [t.to(dev) for t in (host_idx_evt, host_idx_rgb, host_target_blur, host_target_evt)]
    ts_evt = torch.tensor(WINDOW, dtype=torch.float32)
    ts_rgb = torch.tensor(EXPOSURE, dtype=torch.float32)
    launches_py = 0

    @torch.no_grad()        # render + image-formation throughput; the training step is measured by bench_train_step
    def step(idx_evt, idx_rgb, target_blur, target_evt, it):
        """The hot path through the reference-facing API (model/nerf.py:208-232 + train.py:163-331)."""
        nonlocal launches_py
        poses_evt = graph.get_pose_evt(args, ts_evt)
        poses_rgb = graph.get_pose_rgb(args, ts_rgb)
        ret_evt = graph.render(it, poses_evt, idx_evt, H, W, K_MAT, args, enable_crf=True, sensor_type="event", remap=None, training=True)
        ret_rgb = graph.render(it, poses_rgb, idx_rgb, H, W, K_MAT, args, enable_crf=True, sensor_type="rgb", remap=None, training=True)
        parts, blur, diff = [], None, None
        for lvl in ("rgb_map", "rgb0"):
            blur = IF.blur_mean(ret_rgb[lvl], N_POSES)
            diff = IF.event_logdiff(ret_evt[lvl], 1, args.dataset).reshape(-1, 1)
            launches_py += 2
            parts.append(((diff - target_evt * args.event_threshold) ** 2).sum())
            parts.append(((blur - target_blur) ** 2).sum())
        sums = torch.stack(parts)                       # partial sums of squared errors: the only cross-rank exchange
        if world > 1:
            dist.all_reduce(sums)
        return sums, blur, diff

    def step_device(it):
        return step(*dev_inputs, it)

    def step_e2e(it):
        ins = [t.to(dev, non_blocking=True) for t in (host_idx_evt, host_idx_rgb, host_target_blur, host_target_evt)]
        sums, blur, diff = step(*ins, it)
        return sums.cpu(), blur.cpu(), diff.cpu()       # device -> host read of the step's results (synchronises)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        barrier()
        if profile:
            eng.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            out = fn(1000 + i)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, out

    for i in range(opts.warmup):
        step_device(i)
        step_e2e(i)
    sampler = ClockSampler(local)
    sampler.start()
    launches_py = 0
    ms, _, out = timed(step_device, opts.steps, profile=True)
    prof = eng.profile_read()
    eng.profile(False)
    launches_dev_region = prof["launches"] + launches_py
    _, wall_e2e, _ = timed(step_e2e, opts.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    rays_step = (N_POSES + 2) * R * world
    ms_step = ms / opts.steps
    value = rays_step / (ms_step / 1e3)
    e2e_value = rays_step / (wall_e2e / opts.steps)
    peak_tf, _, peak_src = peaks()
    mlp_ms_per_launch = prof["mlp_ms"] / max(prof["mlp_timed"], 1)
    achieved = prof["mlp_flops"] / max(prof["mlp_ms"], 1e-9) / 1e9          # algorithmic TFLOP/s of the MLP kernel
    # tensor-core MACs per algorithmic MAC: the pair kernel merges feature_linear into the view layer
    issued_ratio = MACS_ISSUED_PER_SAMPLE / MACS_PER_SAMPLE if opts.mlp_mode in ("tc", "tc2") else (MACS_PER_SAMPLE - 3456 - 640) / MACS_PER_SAMPLE
    cpu = None
    if rank == 0 and world == 1 and not opts.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        run_cpu_arm(16)
        reps = 10
        rps, rays, dt, kind, what = run_cpu_arm(opts.cpu_pixels, repeats=reps)
        cpu = {"value": rps, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": kind,
               "sample": f"{reps} steps of {opts.cpu_pixels} pixels x 21 poses = {rays} rays, {dt:.2f} s per step ({reps * dt:.0f} s of CPU "
                         f"work), {what}"}
    if cpu is not None and train is not None:
        run_cpu_oracle_train(8, 1)
        runs = [run_cpu_oracle_train(256, 27) for _ in range(5)]          # 1/4 of the configs[2] batch: 512 + 513 rays
        rays, dt = runs[0][1], sum(r[2] for r in runs) / len(runs)
        train["cpu_baseline"] = {"value": rays / dt, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": f"{len(runs)} x ({rays} rays, forward + backward), {dt:.2f} s each, oracle/ under torch autograd "
                                           "(CPU fp32)"}
    if rank == 0:
        h2d = sum(t.numel() * t.element_size() for t in (host_idx_evt, host_idx_rgb, host_target_blur, host_target_evt))
        d2h = 4 * 4 + R * CH * 4 + R * 4
        # DRAM bytes per launch of the dominant kernel: from the committed ncu capture, scaled to this run's launch size (a number,
        # so that it reaches the driver's record; where it comes from is in traffic_detail)
        traffic = ncu_traffic(prof["mlp_flops"] / max(prof["mlp_timed"], 1))
        line = {
            "metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": opts.steps, "warmup": opts.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (tcgen05 fp16 hi/lo split operands, 3 MMAs per product, fp32 TMEM accumulate)" if opts.mlp_mode != "simt" else "f32 (SIMT)",
            "data": "synthetic", "config": workload_config(R, world),
            "roofline": {"bound": "tensor", "kernel": {"tc": "bnrf::tc3::mlp_tc3_kernel<3> (CTA pairs, cta_group::2, A operand in tensor memory)", "tc2": "bnrf::tc2::mlp_tc2_kernel<3> (CTA pairs, cta_group::2)", "tc1": "bnrf::tc::mlp_tc_kernel<3>", "simt": "bnrf::mlp_simt_kernel<3>"}[opts.mlp_mode],
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": (traffic or {}).get("bytes_per_launch"), "traffic_detail": traffic, "peak_source": peak_src,
                         "algorithmic_flop_per_launch": prof["mlp_flops"] / max(prof["mlp_timed"], 1),
                         "ms_per_launch": mlp_ms_per_launch, "launches_timed": prof["mlp_timed"],
                         "issued_tflops": achieved * 3 * issued_ratio if opts.mlp_mode != "simt" else achieved,
                         "issued_frac": achieved * 3 * issued_ratio / peak_tf if opts.mlp_mode != "simt" else None,
                         "mlp_share_of_step": prof["mlp_ms"] / ms,
                         "note": "achieved counts the reference's 593,408 MAC/sample once; the kernel multiplies 523,776 of them "
                                 "(feature_linear merged into the view layer; view channels and heads off the tensor pipe) and the "
                                 "1e-4 parity bound needs 3 fp16 MMAs per product (issued_*)"},
            "cpu_baseline": cpu,
            "train_step": train,
            "train_step_e2nerf_real": train_real,
            "train_step_strong_scaling": train_strong,
            "image_formation_stress": stress,
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * wall_e2e / opts.steps},
            "gpu_launches": int(launches_dev_region),
            "clocks": sampler.summary(),
            "checksum": [float(x) for x in out[0].tolist()],
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pixels", type=int, default=65536, help="pixels per GPU per step (R)")
    ap.add_argument("--cpu-pixels", type=int, default=128, help="pixels of the bounded CPU sample")
    ap.add_argument("--mlp-mode", default="tc", choices=["tc", "tc2", "tc1", "simt"])
    ap.add_argument("--mode", default="render", choices=["render", "train"], help="train: print only the training-step line")
    ap.add_argument("--train-config", default="e2nerf_synthetic", choices=["e2nerf_synthetic", "e2nerf_real"])
    ap.add_argument("--train-scaling", default="weak", choices=["weak", "strong"], help="--mode train: per-GPU batch fixed, or one batch split")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-step", action="store_true", help="skip the extra training-step measurement")
    opts = ap.parse_args()
    if opts.impl == "reference":
        bench_reference(opts)
    else:
        bench_ours(opts)


if __name__ == "__main__":
    main()
