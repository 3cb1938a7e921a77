"""Multi-GPU check of the peer-memory exchange THROUGH THE DROP-IN API (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_api_check.py

Each rank evaluates ``mpcl_loss_calc(..., group=True)`` (NCCL all-reduce of the 8-byte pair + rescale kernel) and
``mpcl_loss_calc(..., group=PeerMailbox(dev))`` (slcl_proto_rescale_peer: exchange + rescale as one kernel over NVLink peer
mailboxes) on its own shard, forward + backward, several steps, and compares losses and gradients; the fused target step is
checked the same way.  ``bench.py`` makes the same comparison at the plan level (raw C ABI) before it times anything; this
script covers the autograd / custom-op route (``slcl.functional._exchange_loss_pair`` -> ``slcl::proto_rescale_peer``).
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))


def main():
    from slcl.loss import MPCL, mpcl_loss_calc, mpcl_target_step
    from slcl.peer import PeerMailbox
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    mbox = PeerMailbox(dev)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    b, c, h, w, k = 4, 64, 48, 48, 5
    crit = MPCL(dev, num_class=k, temperature=0.1, m=0.4, base_temperature=1.0)
    centres = torch.randn(k, c, device=dev, generator=gen)
    dist.broadcast(centres, 0)
    worst = 0.0
    for step in range(6):
        feat = torch.randn(b, c, h, w, device=dev, generator=gen)
        labels = torch.randint(0, k, (b, h, w), device=dev, generator=gen)
        sel = (torch.rand(b * h * w, device=dev, generator=gen) > 0.5).float()
        res = []
        for group in (True, mbox):
            f = feat.clone().requires_grad_(True)
            if step % 3 == 0:
                loss = mpcl_loss_calc(f, labels, centres, crit, tag="source", group=group)
            elif step % 3 == 1:
                loss = mpcl_loss_calc(f, labels.reshape(-1), centres, crit, pixel_sel_loc=sel, tag="target", group=group)
            else:
                loss, _, _ = mpcl_target_step(f, centres, crit, 0.05, group=group)
            loss.backward()
            res.append((loss.detach().clone(), f.grad.clone()))
        torch.cuda.synchronize(dev)
        rel = float((res[0][0] - res[1][0]).abs() / res[0][0].abs().clamp_min(1e-30))
        gd = float((res[0][1] - res[1][1]).abs().max() / res[0][1].abs().max().clamp_min(1e-30))
        worst = max(worst, rel, gd)
        if rank == 0:
            print(f"step {step}: loss nccl {float(res[0][0]):.8f} peer {float(res[1][0]):.8f}  rel {rel:.2e}  grad diff {gd:.2e}")
    t = torch.tensor([worst, float(mbox.timeouts())], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = float(t[0]) < 1e-6 and float(t[1]) == 0
    if rank == 0:
        print("peer exchange through the API:", "ok" if ok else f"MISMATCH (worst {float(t[0]):.3e}, timeouts {int(t[1])})")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
