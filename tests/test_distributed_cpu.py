"""world_size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding,
the single all-reduce of the [sets*K, C+1] class sums, and the global-mean loss exchange.  The
per-rank partial sums are produced by the oracle here (the CUDA kernels cannot run on this host);
what is under test is the product's exchange + finalisation logic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_sums(feat, labels, k):
    """[K, C+1] float64: per-class feature sums | counts (what slcl_class_sums produces)."""
    rows = []
    for c in range(k):
        m = (labels == c).unsqueeze(1).to(feat.dtype)
        rows.append(torch.cat([(feat * m).sum(dim=(0, 2, 3)).double(), m.sum().double().reshape(1)]))
    return torch.stack(rows)


def _worker(rank, world, port, tmp):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "soft-labeled-contrastive-learning_b200"))
    from slcl.distributed import all_reduce_mean_loss, all_reduce_sums, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        b, c, h, w, k = 6, 8, 5, 7, 4
        feat = torch.randn(b, c, h, w, generator=g)
        labels = torch.randint(0, k, (b, h, w), generator=g)
        rows = torch.randn(b * h * w, generator=g).abs()
        sel = (torch.rand(b * h * w, generator=g) > 0.3).float()
        lo, hi = shard_range(b, rank, world)
        local = _oracle_sums(feat[lo:hi], labels[lo:hi], k)
        total = all_reduce_sums(local, True)
        whole = _oracle_sums(feat, labels, k)
        assert torch.equal(total[:, -1], whole[:, -1])                       # counts bit-exact
        torch.testing.assert_close(total, whole, rtol=1e-5, atol=1e-5)     # fp32 partial sums on the oracle side
        assert torch.equal(local, _oracle_sums(feat[lo:hi], labels[lo:hi], k))   # input not mutated
        px = slice(lo * h * w, hi * h * w)
        loss = all_reduce_mean_loss((rows[px] * sel[px]).sum(), sel[px].sum() + (1e-4 if rank == 0 else 0.0), True)
        want = (rows * sel).sum() / (sel.sum() + 1e-4)
        torch.testing.assert_close(loss, want, rtol=1e-6, atol=0)
        with open(os.path.join(tmp, f"ok{rank}"), "w") as fh:
            fh.write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_sums_and_loss(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_all_reduce_is_identity_without_process_group():
    from slcl.distributed import all_reduce_sums
    t = torch.arange(6, dtype=torch.float64).reshape(2, 3)
    assert all_reduce_sums(t, None) is t
