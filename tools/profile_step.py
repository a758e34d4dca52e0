"""One un-graphed step of the DETR hot path inside a cudaProfilerStart/Stop range.

Run under `ncu --profile-from-start off ...` so the launch list holds exactly one step.
Usage: python tools/profile_step.py [precision] [batch]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import seeded_weights  # noqa: E402
from focoos_b200 import DETRConfig, ops  # noqa: E402
from focoos_b200.fai_detr import FAIDetr  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "fp16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda", 0)
model = FAIDetr(DETRConfig(), precision=precision)
model.load_state_dict(seeded_weights(), strict=True)
model.to(dev)
rng = np.random.default_rng(1)
x = torch.from_numpy(rng.integers(0, 256, (B, 640, 640, 3), dtype=np.uint8)).to(dev)
sizes = torch.tensor([(640, 640)] * B, dtype=torch.int32, device=dev)


def step():
    out = model(x)
    return ops.detr_postprocess(out.logits, out.boxes, sizes, 300, 0.5)


for _ in range(3):
    step()
torch.cuda.synchronize()
l0 = ops.launch_count()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("launches in range:", ops.launch_count() - l0)
