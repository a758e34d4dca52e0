"""Not a test (no test_ prefix): runs the training graph on CUDA and on the CPU references side by side and prints where they diverge."""
import sys

import numpy as np
import torch

from focoos_b200 import DETRConfig, FAIDetr, ops
from oracle.gen_golden import synth_images
from oracle.ops_ref import RefBackend
from focoos_b200.utils.seeded_weights import desaturate_classifiers
from tests.parity_utils import seeded_sd

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
size, B = 192, 2
x = torch.from_numpy(np.stack(synth_images(5, [(size, size)] * B))).permute(0, 3, 1, 2).float()


def run(dev):
    m = FAIDetr(DETRConfig(), precision=prec)
    m.load_state_dict(desaturate_classifiers(seeded_sd(0)), strict=True)
    m.to(dev).train()
    g = m.train_graph()
    g.taps = {}
    with torch.no_grad():
        out = g.forward(x.to(dev))
    t = {k: v.cpu() for k, v in g.taps.items()}
    t["pred_logits"], t["pred_boxes"] = out["pred_logits"].cpu(), out["pred_boxes"].cpu()
    t["topk"] = out["_topk_ind"].cpu().float()
    return t


gpu = run("cuda")
ops._backend = RefBackend()
cpu = run("cpu")
for k in cpu:
    a, b = gpu[k].float(), cpu[k].float()
    print(f"{k:18s} max|d| {float((a - b).abs().max()):.3e}   scale {float(b.abs().max()):.3e}")
