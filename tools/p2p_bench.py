"""cfg3 pixel<->pixel timings only (the p2p section of bench.py), one line per entry."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
import bench  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(1234)
    for k, v in bench.p2p_kernels(dev, gen).items():
        print(f"{v['ms'] * 1e3:8.1f} us  {v['achieved_TFLOPs']:7.1f} TF  {v['frac_of_bf16_peak'] * 100:5.1f}%  {k}")
