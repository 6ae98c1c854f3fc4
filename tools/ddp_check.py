#!/usr/bin/env python
"""Multi-GPU check of train.Trainer (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ddp_check.py

Every rank trains on its own pixels for 6 steps (2 eager + capture + replays).  After every step the parameters must be
IDENTICAL on all ranks (any gradient range that missed the all-reduce would make them drift apart), and the run with the
all-reduce overlapped with the backward pass (default) must match the run with one all-reduce at the end.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(overlap, rank, world, dev):
    from benerf_b200 import optimize, run_nerf_helpers
    from benerf_b200.train import Trainer
    from tests.cases import CASES
    from tests.test_gpu_backward import case_args
    case = CASES["e2nerf_syn"]
    args = case_args(case)
    args.fused_optimizer, args.cuda_graph, args.overlap_all_reduce = True, True, overlap
    args.lrate, args.pose_lrate, args.transform_lrate, args.rgb_crf_lrate, args.event_crf_lrate = 5e-4, 1e-3, 1e-6, 5e-4, 5e-4
    args.optimize_nerf, args.optimize_pose, args.optimize_trans = True, True, True
    args.event_coeff_syn, args.rgb_coeff = 0.1, 1.0
    torch.manual_seed(0)
    model = optimize.Model(args)
    graph = model.build_network(args)
    run_nerf_helpers.init_nerf(graph.nerf)
    run_nerf_helpers.init_nerf(graph.nerf_fine)
    graph.to(dev)
    tr = Trainer(model, args)
    g = torch.Generator().manual_seed(100 + rank)              # every rank its own pixels
    losses, spread = [], []
    for it in range(6):
        idx_evt = torch.randint(0, case.H * case.W, (96,), generator=g).to(dev)
        idx_rgb = torch.randint(0, case.H * case.W, (16,), generator=g).to(dev)
        blur_t = torch.rand(16, case.channels, generator=g).to(dev)
        accu = torch.randint(-3, 4, (case.H, case.W), generator=torch.Generator().manual_seed(7)).double().to(dev)
        loss, _ = tr.step(accu, idx_evt, idx_rgb, blur_t, torch.tensor(case.window), torch.tensor(case.exposure), case.H, case.W, case.K, case.K)
        losses.append(float(loss))
        print(f"[rank {rank}] overlap={overlap} step {it} loss {losses[-1]:.6f}", file=sys.stderr, flush=True)
        flat = torch.cat([p.detach().reshape(-1) for p in graph.parameters()])
        lo, hi = flat.clone(), flat.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        spread.append(float((hi - lo).abs().max()))
    return losses, spread, torch.cat([p.detach().reshape(-1) for p in graph.parameters()]).cpu(), tr


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    la, sa, pa, tra = run(True, rank, world, dev)
    lb, sb, pb, trb = run(False, rank, world, dev)
    ok = max(sa) == 0.0 and max(sb) == 0.0 and float((pa - pb).abs().max()) < 2e-3 and tra._cg is not None
    if rank == 0:
        print("overlapped :", [f"{x:.6f}" for x in la], "parameter spread over ranks", sa, "launches", tra.launches_per_step)
        print("one at end :", [f"{x:.6f}" for x in lb], "parameter spread over ranks", sb, "launches", trb.launches_per_step)
        print("max parameter difference between the two runs", float((pa - pb).abs().max()))
        print("DDP CHECK", "OK" if ok else "FAILED")
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if ok else 1)       # the captured graphs still hold NCCL work: tearing the process group down under them hangs


if __name__ == "__main__":
    main()
