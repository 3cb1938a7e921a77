"""Three BlockConLoss steps (forward + backward through the Python API) at the reference's documented shape
(1, 2, 32, 224, 224), 32 x 32 tiles -- the command behind profiles/r2b_blockcon_launches.csv:

    ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \
        --log-file gpurun_out/r2b_blockcon_launches.csv python tools/blockcon_step.py
"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "soft-labeled-contrastive-learning_b200"))
from slcl.loss import BlockConLoss
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
fb = torch.nn.functional.normalize(torch.randn(1, 2, 32, 224, 224, device=dev, generator=g), dim=2).requires_grad_(True)
lbk = torch.randint(0, 4, (1, 2, 224, 224), device=dev, generator=g)
crit = BlockConLoss(0.7, 32)
for _ in range(int(os.environ.get("SLCL_STEPS", "3"))):
    loss = crit(fb, lbk)
    loss.backward()
    fb.grad = None
torch.cuda.synchronize()
print("loss", float(loss))
