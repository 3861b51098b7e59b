"""World-size-2 gloo test of the data-parallel host logic (sqlx/dist.py): sharding the batch by sample and
averaging parameter gradients across ranks reproduces the single-process full-batch gradients.  The per-rank
compute here is the CPU oracle (the CUDA kernels need a GPU); the exchange step is the code the GPU path uses."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _loss_and_grads(batch, params):
    from oracle import sqldepth_oracle as O
    B = batch["x"].shape[0]
    Wp, bp, W1, b1 = params
    F = torch.nn.functional
    tail_energy, summ = O.full_query(batch["x"], batch["queries"])
    centers = O.bin_centers(F.linear(summ.reshape(B, -1), W1, b1), 0.1, 80.0)
    pred = O.bins_expectation(tail_energy, Wp, bp, centers)
    poses = [{"axisangle": batch["aa%d" % i], "translation": batch["tr%d" % i], "invert": i == 0} for i in range(2)]
    out = O.photometric_losses({0: pred}, {0: batch["target"]}, [batch["s0"], batch["s1"]], batch["K"], batch["inv_K"],
                               poses, {0: batch["noise"]}, height=32, width=48)
    grads = torch.autograd.grad(out["loss"], params)
    return out["loss"].detach(), grads


def _make(seed=0):
    from _cases import smooth_images, kitti_K
    g = torch.Generator().manual_seed(seed)
    B, H, W, E, Q, D = 4, 32, 48, 16, 6, 8
    fr = smooth_images(g, B, H, W, 3)
    K, iK = kitti_K(B, H, W)
    batch = {"x": torch.randn(B, E, H // 2, W // 2, generator=g).double(), "queries": 0.4 * torch.randn(B, Q, E, generator=g).double(),
             "target": fr[1].double(), "s0": fr[0].double(), "s1": fr[2].double(), "K": K.double(), "inv_K": iK.double(),
             "noise": torch.randn(B, 2, H, W, generator=g).double()}
    for i in range(2):
        batch["aa%d" % i] = 0.01 * torch.randn(B, 1, 1, 3, generator=g).double()
        batch["tr%d" % i] = 0.05 * torch.randn(B, 1, 1, 3, generator=g).double()
    params = [0.3 * torch.randn(D, Q, generator=g).double(), 0.1 * torch.randn(D, generator=g).double(),
              (torch.randn(D, Q * E, generator=g) / (Q * E) ** 0.5).double(), 0.1 * torch.randn(D, generator=g).double()]
    return batch, params


def _worker(rank, world, port, ret):
    for p in (ROOT, os.path.join(ROOT, "sfmnext-impl_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sqlx.dist import GradBucket, shard_batch
    torch.set_num_threads(1)
    batch, params = _make()
    params = [p.clone().requires_grad_(True) for p in params]
    shard = shard_batch(batch, rank, world)
    assert shard["x"].shape[0] == 2
    loss, grads = _loss_and_grads(shard, params)
    grads = [g.float().contiguous() for g in grads]
    # the hot path's form: every gradient is a view of ONE flat buffer that is averaged in place (no pack / unpack)
    from sqlx.dist import allreduce_flat_
    sizes = [g.numel() for g in grads]
    flat = torch.cat([g.reshape(-1) for g in grads]).clone()
    views = [v.view_as(g) for v, g in zip(flat.split(sizes), grads)]
    allreduce_flat_(flat)
    bucket = GradBucket(grads)
    bucket.allreduce_(grads)
    for v, g in zip(views, grads):
        assert torch.equal(v, g), "flat in-place exchange and the packed bucket must agree bit for bit"
    lt = loss.clone().float().reshape(1)
    dist.all_reduce(lt)
    if rank == 0:
        ret["loss"] = float(lt) / world
        ret["grads"] = [g.clone() for g in grads]
    dist.destroy_process_group()


def test_sharded_gradients_match_full_batch():
    sys.path.insert(0, os.path.join(ROOT, "sfmnext-impl_b200"))
    batch, params = _make()
    params = [p.clone().requires_grad_(True) for p in params]
    loss_full, grads_full = _loss_and_grads(batch, params)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert abs(ret["loss"] - float(loss_full)) < 1e-6
    for a, b in zip(ret["grads"], grads_full):
        assert float((a.double() - b).abs().max()) <= 1e-5 * float(b.abs().max()) + 1e-9


def test_shard_batch_rejects_ragged():
    sys.path.insert(0, os.path.join(ROOT, "sfmnext-impl_b200"))
    from sqlx.dist import shard_batch
    with pytest.raises(ValueError):
        shard_batch({"x": torch.zeros(3, 2)}, 0, 2)
