"""Golden fixtures for the BisenetFormer path FROM THE UNMODIFIED REFERENCE (build container only): python -m oracle.gen_golden_bisenet"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from focoos_b200.utils.seeded_weights import seeded_state_dict  # noqa: E402
from oracle import ref_import  # noqa: E402
from oracle.gen_golden import state_dict_digest, synth_images  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
NAME = "bisenetformer-l-ade"
SIZES = [(256, 384), (256, 384)]


def main():
    fm = ref_import.get_reference_model(NAME)
    template = fm.model.state_dict()
    with open(os.path.join(GOLDEN, "bisenetformer_l_ade_state_dict_manifest.json"), "w") as f:
        json.dump({k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in template.items()}, f, indent=0, sort_keys=True)
    sd = seeded_state_dict(template, seed=0)
    fm.model.load_state_dict(sd, strict=True)
    fm.model.eval()
    imgs = synth_images(4, SIZES)
    x = torch.stack([torch.from_numpy(im).permute(2, 0, 1).float() for im in imgs])
    taps = {}
    hooks = [fm.model.head.predictor.register_forward_hook(lambda m, i, o: taps.__setitem__("pred", {k: v.detach() for k, v in o.items() if k != "aux_outputs"})),
             fm.model.pixel_decoder.register_forward_hook(lambda m, i, o: taps.__setitem__("pd", o))]
    with torch.no_grad():
        out = fm.model(x)
    for h in hooks:
        h.remove()
    mf, (cp32, cp16, cp8) = taps["pd"]
    pm = taps["pred"]["pred_masks"]
    g = {"logits": out.logits.numpy(), "pred_masks_q4": pm[:, ::4].numpy(), "pred_masks_stat": np.array([pm.mean().item(), pm.std().item(), pm.abs().max().item()], np.float32),
         "masks_q10_s4": out.masks[:, ::10, ::4, ::4].numpy(), "mask_features_tap": mf[:, ::16, ::2, ::2].numpy(), "cp32_tap": cp32[:, ::16].numpy(),
         "cp16_tap": cp16[:, ::16].numpy(), "sizes": np.array(SIZES, np.int32)}
    from focoos.models.bisenetformer.ports import BisenetFormerOutput

    thr = 0.5
    ks, kl, kb = [], [], []
    for i in range(len(imgs)):
        o1 = BisenetFormerOutput(masks=out.masks[i:i + 1], logits=out.logits[i:i + 1], loss=None)
        dets = fm.processor.postprocess(o1, [imgs[i]], class_names=[], threshold=thr)[0]
        ks.append([d.conf for d in dets.detections]); kl.append([d.cls_id for d in dets.detections]); kb.append([d.bbox for d in dets.detections])
    n = max(1, max(len(s) for s in ks))
    ds = np.zeros((len(imgs), n), np.float32); dl = np.full((len(imgs), n), -1, np.int32); db = np.zeros((len(imgs), n, 4), np.int32); dc = np.zeros(len(imgs), np.int32)
    for i in range(len(imgs)):
        k = len(ks[i]); dc[i] = k; ds[i, :k] = ks[i]; dl[i, :k] = kl[i]
        if k:
            db[i, :k] = np.array(kb[i])
    g.update(det_scores=ds, det_labels=dl, det_boxes=db, det_count=dc, threshold=np.float32(thr))
    np.savez_compressed(os.path.join(GOLDEN, "bisenetformer_l_ade_b2_256x384.npz"), **g)
    meta = {"model": NAME, "weights_seed": 0, "weights_sha256": state_dict_digest(sd), "image_seed": 4, "sizes": SIZES, "threshold": thr, "det_count": dc.tolist()}
    with open(os.path.join(GOLDEN, "golden_meta_bisenet.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(meta, "pred_masks stat", g["pred_masks_stat"], "max prob", out.logits.max().item())
    from oracle import bisenet_oracle as Bo

    with torch.no_grad():
        ot = {}
        probs, masks = Bo.bisenet_forward(sd, x, Bo.BisenetOracleConfig(), ot)
    print("oracle vs reference: logits", (probs - out.logits).abs().max().item(), "pred_masks", (ot["pred_masks"] - pm).abs().max().item(), "masks", (masks - out.masks).abs().max().item())
    for i in range(len(imgs)):
        q, s, l, bm = Bo.semantic_postprocess_tensors(probs[i:i + 1], masks[i:i + 1], Bo.BisenetOracleConfig(), thr)
        print("img", i, "oracle kept", len(q), "ref kept", int(dc[i]))


if __name__ == "__main__":
    main()
