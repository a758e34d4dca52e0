"""tests/golden/detr_l_train_b2_192.npz: one training-mode forward + backward of the UNMODIFIED reference
(FAIDetr.forward with self.training, SetCriterion, torch autograd; fp32 on the CPU) on seeded weights / images / targets.

    python -m oracle.gen_golden_train

Stored: the 21 loss scalars, for every one of the 501 trainable tensors the L2 norm and the sum of its gradient, a few small
gradients in full, and BatchNorm running statistics after the step (they are updated by the training forward)."""
import os

import numpy as np
import torch

from focoos_b200.utils.seeded_weights import desaturate_classifiers, seeded_state_dict
from oracle import ref_import
from oracle.gen_golden import synth_images

FULL = ["head.predictor.dec_score_classifier.5.bias", "head.predictor.enc_output.0.bias", "pixel_decoder.backbone.conv1.conv1_1.conv.weight",
        "pixel_decoder.backbone.res_layers.0.blocks.0.branch2a.norm.weight", "head.predictor.decoder.layers.0.cross_attn.sampling_offsets.bias",
        "head.predictor.query_pos_head.layers.0.weight", "pixel_decoder.encoder.0.layers.0.self_attn.in_proj_bias", "head.predictor.dec_bbox_classifier.2.layers.2.weight"]
BN_BUFFERS = ["pixel_decoder.backbone.conv1.conv1_1.norm.running_mean", "pixel_decoder.backbone.conv1.conv1_1.norm.running_var",
              "pixel_decoder.fpn_blocks.0.bottlenecks.1.conv2.norm.running_var", "head.predictor.input_proj.2.norm.running_mean"]


def synth_targets(seed, B, num_classes):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(B):
        n = int(torch.randint(1, 9, (1,), generator=g))
        cxcy = 0.2 + 0.6 * torch.rand((n, 2), generator=g)
        wh = 0.05 + 0.30 * torch.rand((n, 2), generator=g)
        out.append((torch.randint(0, num_classes, (n,), generator=g), torch.cat([cxcy, wh], 1)))
    return out


def main(size=192, B=2):
    fm = ref_import.get_reference_model("fai-detr-l-obj365")
    from focoos.models.fai_detr.ports import DETRTargets

    model = fm.model
    sd = desaturate_classifiers(seeded_state_dict(model.state_dict(), seed=0))
    model.load_state_dict(sd, strict=True)
    model.train()
    imgs = np.stack(synth_images(5, [(size, size)] * B))
    x = torch.from_numpy(imgs).permute(0, 3, 1, 2).float()
    targets = synth_targets(6, B, model.config.num_classes)
    recorded = []  # the matcher's assignments in call order: final layer, aux_0..aux_4 (decoder layers), aux_5 (encoder proposals)
    matcher = model.head.criterion.matcher
    orig_forward = matcher.forward

    def recording_forward(outputs, tg):
        idx = orig_forward(outputs, tg)
        row = []
        for (qi, tj), t in zip(idx, tg):
            mq = torch.empty(len(t.labels), dtype=torch.int64)
            mq[tj] = qi
            row.append(mq)
        recorded.append(torch.cat(row))
        return idx

    matcher.forward = recording_forward
    topk_calls = []  # the query selection (modelling.py:1214) is the only torch.topk of the training forward
    orig_topk = torch.topk

    def recording_topk(*a, **k):
        r = orig_topk(*a, **k)
        topk_calls.append(r[1].clone())
        return r

    torch.topk = recording_topk
    try:
        out = model(x, [DETRTargets(labels=t[0], boxes=t[1]) for t in targets])
    finally:
        torch.topk = orig_topk
        matcher.forward = orig_forward
    assert len(topk_calls) == 1 and len(recorded) == 7
    losses = out.loss
    sum(losses.values()).backward()
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    params = dict(model.named_parameters())
    has = np.array([params[n].grad is not None for n in names])
    gnorm = np.array([float(params[n].grad.norm()) if params[n].grad is not None else 0.0 for n in names], dtype=np.float64)
    gsum = np.array([float(params[n].grad.double().sum()) if params[n].grad is not None else 0.0 for n in names], dtype=np.float64)
    full = {n: params[n].grad.numpy().astype(np.float32).copy() for n in FULL}  # before clip_grad_norm_ rescales the gradients in place
    # ---- the optimiser half of TrainerLoop.run_step (trainer.py:757-773, amp off as on a CPU host): clip_grads, then the clipping AdamW's step
    from focoos.trainer.solver.build import build_optimizer

    before = {n: params[n].detach().clone() for n in names}
    opt = build_optimizer("ADAMW", learning_rate=5e-4, weight_decay=0.02, model=model, weight_decay_norm=0.0, weight_decay_embed=0.0, backbone_multiplier=0.1,
                          clip_gradients=0.1)
    ps = [p for p in model.parameters() if p.requires_grad and p.grad is not None]
    total_norm = float(torch.nn.utils.clip_grad_norm_(ps, 0.1))
    opt.step()
    dnorm = np.array([float((params[n].detach() - before[n]).norm()) for n in names], dtype=np.float64)
    dsum = np.array([float((params[n].detach() - before[n]).double().sum()) for n in names], dtype=np.float64)
    keys = sorted(losses.keys())
    bufs = dict(model.named_buffers())
    blob = {"loss_keys": np.array(keys), "loss_values": np.array([float(losses[k]) for k in keys], dtype=np.float64), "param_names": np.array(names), "grad_has": has,
            "grad_norm": gnorm, "grad_sum": gsum, "size": np.array([size, B]), "total_grad_norm": np.array(total_norm), "match_q": torch.stack(recorded).numpy().astype(np.int32), "topk_ind": topk_calls[0].numpy().astype(np.int32), "step_delta_norm": dnorm, "step_delta_sum": dsum}
    for n in FULL:
        blob["grad::" + n] = full[n]
    for n in BN_BUFFERS:
        blob["buf::" + n] = bufs[n].detach().numpy().astype(np.float32)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"detr_l_train_b{B}_{size}.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; total loss", float(sum(losses.values())), "params without grad:", [n for n, h in zip(names, has) if not h])


if __name__ == "__main__":
    main()
