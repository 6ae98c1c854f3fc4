#!/usr/bin/env python
"""Host-side profile of Trainer.step (cProfile over N steps; bring-up aid: is the training step launch-bound?)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from benerf_b200 import optimize, run_nerf_helpers
from benerf_b200.train import Trainer


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    dev = torch.device("cuda", 0)
    args = bench.ref_args()
    args.dataset, args.event_threshold, args.seed = "E2NeRF_Synthetic", 0.2, 1
    args.lrate, args.pose_lrate, args.transform_lrate, args.rgb_crf_lrate, args.event_crf_lrate = 5e-4, 1e-3, 1e-6, 5e-4, 5e-4
    args.optimize_nerf, args.optimize_pose, args.optimize_trans = True, True, False
    model = optimize.Model(args)
    graph = model.build_network(args)
    run_nerf_helpers.init_nerf(graph.nerf); run_nerf_helpers.init_nerf(graph.nerf_fine)
    graph.to(dev)
    tr = Trainer(model, args)
    g = torch.Generator().manual_seed(0)
    idx_evt = torch.randint(0, 640000, (1024,), generator=g).to(dev)
    idx_rgb = torch.randint(0, 640000, (107,), generator=g).to(dev)
    blur_t = torch.rand(107, 3, generator=g).to(dev)
    accu = torch.randint(-3, 4, (800, 800), generator=g).double().to(dev)
    K = [[1111.111, 0, 400.0], [0, 1111.111, 400.0], [0, 0, 1]]
    one = lambda: tr.step(accu, idx_evt, idx_rgb, blur_t, torch.tensor((0.3, 0.4)), torch.tensor((0.2, 0.8)), 800, 800, K, K)
    for _ in range(5):
        one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        one()
    t_host = time.perf_counter() - t0            # time to ENQUEUE n steps (the host runs ahead of the GPU)
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"host enqueue {1e3 * t_host / n:.2f} ms/step, wall incl. GPU drain {1e3 * t_all / n:.2f} ms/step")
    # host time per phase WITHOUT synchronising (how long the host takes to enqueue each part)
    import types, collections
    acc = collections.defaultdict(float)
    state = {"t": None}
    def mark(self, name):
        now = time.perf_counter()
        if name is not None:
            acc[name] += now - state["t"]
        state["t"] = now
    tr._mark = types.MethodType(mark, tr)
    for _ in range(n):
        one()
    torch.cuda.synchronize()
    print("host ms/step by phase (no sync):", {k: round(1e3 * v / n, 2) for k, v in acc.items()})
    # the same with the GPU idle at the start of every step (pure host cost, nothing to wait for)
    acc.clear()
    for _ in range(n):
        torch.cuda.synchronize()
        one()
    torch.cuda.synchronize()
    print("host ms/step by phase (GPU idle at step start):", {k: round(1e3 * v / n, 2) for k, v in acc.items()})
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        one()
    pr.disable()
    torch.cuda.synchronize()
    out = io.StringIO()
    pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(28)
    print(out.getvalue())


if __name__ == "__main__":
    main()
