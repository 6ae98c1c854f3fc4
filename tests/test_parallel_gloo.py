"""World-size-2 tests of the data-parallel host logic on CPU (gloo): pixel sharding, the single flat-gradient
all-reduce, and the normalised event loss whose norms span the whole batch (train.py:238-292)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from benerf_b200 import parallel, image_formation as IF
        g = torch.Generator().manual_seed(7)
        R = 64
        theta0 = torch.randn(5, generator=g)
        feats = torch.randn(R, 5, generator=g)
        target = torch.randn(R, 1, generator=g).double()
        blur_t = torch.randn(R, 3, generator=g)

        def losses(theta, sl):
            diff = (feats[sl] @ theta).reshape(-1, 1)                 # stands in for the rendered log-brightness difference
            ev = IF.event_loss(diff, target[sl], -1.0, 0.1, 2.0)      # normalised branch
            bl = IF.mse(torch.tanh(feats[sl] @ theta)[:, None].expand(-1, 3), blur_t[sl])
            return ev + bl

        # single-process reference on the global batch (world() == 1 inside global_sum is false here, so emulate):
        theta_ref = theta0.clone().requires_grad_(True)
        diff = (feats @ theta_ref).reshape(-1, 1)
        dn = diff / (torch.linalg.norm(diff, dim=0, keepdim=True) + 1e-9)
        tn = target / (torch.linalg.norm(target, dim=0, keepdim=True) + 1e-9)
        ref = IF.mse(dn, tn) * 2.0 + IF.mse(torch.tanh(feats @ theta_ref)[:, None].expand(-1, 3), blur_t)
        ref.backward()

        theta = theta0.clone().requires_grad_(True)
        flat = parallel.FlatGrads([theta])
        per = R // world
        sl = slice(rank * per, (rank + 1) * per)
        assert torch.equal(parallel.shard(feats), feats[sl])
        loss = losses(theta, sl)
        loss.backward()
        assert theta.grad.data_ptr() == flat.flat.data_ptr()          # gradients accumulated in place into the flat buffer
        flat.all_reduce_mean()
        loss_mean = loss.detach().clone()
        dist.all_reduce(loss_mean)
        loss_mean /= world
        out[rank] = (float((theta.grad - theta_ref.grad).abs().max()), float(abs(loss_mean - ref.detach())))
    finally:
        dist.destroy_process_group()


def test_sharded_step_equals_single_process_global_batch():
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        grad_err, loss_err = out[r]
        assert grad_err < 1e-6 and loss_err < 1e-6, (r, grad_err, loss_err)
