"""Golden fixtures for the FAIMaskFormer path FROM THE UNMODIFIED REFERENCE (build container only):

    python -m oracle.gen_golden_mf

fai-mf-l-coco-ins (R101-vd, 6 pixel-decoder transformer layers, 9 masked decoder layers, 100 queries, instance
post-processing) with the seeded state_dict, run through the reference's own forward and (per image) postprocess."""
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
NAME = "fai-mf-l-coco-ins"
SIZES = [(320, 416), (320, 416)]


def main():
    fm = ref_import.get_reference_model(NAME)
    template = fm.model.state_dict()
    with open(os.path.join(GOLDEN, "fai_mf_l_coco_ins_state_dict_manifest.json"), "w") as f:
        json.dump({k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in template.items()}, f, indent=0, sort_keys=True)
    sd = seeded_state_dict(template, seed=0)
    fm.model.load_state_dict(sd, strict=True)
    fm.model.eval()
    imgs = synth_images(3, SIZES)
    x = torch.stack([torch.from_numpy(im).permute(2, 0, 1).float() for im in imgs])
    taps = {}
    hooks = [
        fm.model.pixel_decoder.mask_features.register_forward_hook(lambda m, i, o: taps.__setitem__("mask_features", o.detach())),
        fm.model.head.predictor.register_forward_hook(lambda m, i, o: taps.__setitem__("pred", {k: v.detach() for k, v in o.items() if k != "aux_outputs"})),
        fm.model.pixel_decoder.transformer.register_forward_hook(lambda m, i, o: taps.__setitem__("enc_memory", o.detach())),
    ]
    with torch.no_grad():
        out = fm.model(x)
    for h in hooks:
        h.remove()
    g = {
        "logits": out.logits.numpy(),                                   # [B,Q,K] softmax probs without no-object
        "pred_logits_raw": taps["pred"]["pred_logits"].numpy(),         # [B,Q,K+1]
        "pred_masks_q4": taps["pred"]["pred_masks"][:, ::4].numpy(),    # pre-sigmoid, every 4th query, 1/4 resolution
        "pred_masks_stat": np.array([taps["pred"]["pred_masks"].mean().item(), taps["pred"]["pred_masks"].std().item(), taps["pred"]["pred_masks"].abs().max().item()], np.float32),
        "masks_q10_s4": out.masks[:, ::10, ::4, ::4].numpy(),           # final probabilities, subsampled
        "mask_features_tap": taps["mask_features"][:, ::32, ::4, ::4].numpy(),
        "enc_memory_tap": taps["enc_memory"][:, ::32].numpy(),
        "sizes": np.array(SIZES, np.int32),
    }
    # post-process per image through the reference (B=1 only)
    from focoos.models.fai_mf.ports import MaskFormerModelOutput

    thr = 0.5
    keep_q, keep_s, keep_l, keep_px, keep_box = [], [], [], [], []
    for i in range(len(imgs)):
        o1 = MaskFormerModelOutput(masks=out.masks[i:i + 1], logits=out.logits[i:i + 1], loss=None)
        dets = fm.processor.postprocess(o1, [imgs[i]], class_names=[], threshold=thr)[0]
        keep_s.append([d.conf for d in dets.detections])
        keep_l.append([d.cls_id for d in dets.detections])
        keep_box.append([d.bbox for d in dets.detections])
    n = max(1, max(len(s) for s in keep_s))
    ds = np.zeros((len(imgs), n), np.float32); dl = np.full((len(imgs), n), -1, np.int32); db = np.zeros((len(imgs), n, 4), np.int32); dc = np.zeros(len(imgs), np.int32)
    for i in range(len(imgs)):
        k = len(keep_s[i]); dc[i] = k
        ds[i, :k] = keep_s[i]; dl[i, :k] = keep_l[i]
        if k:
            db[i, :k] = np.array(keep_box[i])
    g.update(det_scores=ds, det_labels=dl, det_boxes=db, det_count=dc, threshold=np.float32(thr))
    np.savez_compressed(os.path.join(GOLDEN, "mf_l_coco_ins_b2_320x416.npz"), **g)
    meta = {"model": NAME, "weights_seed": 0, "weights_sha256": state_dict_digest(sd), "image_seed": 3, "sizes": SIZES, "threshold": thr, "det_count": dc.tolist()}
    with open(os.path.join(GOLDEN, "golden_meta_mf.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(meta, "pred_masks stat", g["pred_masks_stat"], "max prob", out.logits.max().item())

    from oracle import mf_oracle as M

    with torch.no_grad():
        otaps = {}
        probs, masks = M.mf_forward(sd, x, M.MFOracleConfig(), otaps)
    print("oracle vs reference: logits max|d|", (probs - out.logits).abs().max().item(), "pred_masks max|d|", (otaps["pred_masks"] - taps["pred"]["pred_masks"]).abs().max().item(),
          "final masks max|d|", (masks - out.masks).abs().max().item())
    for i in range(len(imgs)):
        q, s, l, bm = M.mf_postprocess_tensors(probs[i:i + 1], masks[i:i + 1], M.MFOracleConfig(), thr)
        print("img", i, "oracle kept", len(q), "ref kept", int(dc[i]), "scores close", np.allclose(np.array(s), ds[i, :dc[i]], atol=1e-4) if len(q) == dc[i] else None)


if __name__ == "__main__":
    main()
