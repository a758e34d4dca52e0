"""Generate the committed golden fixtures under tests/golden/ FROM THE UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden            # writes tests/golden/*.npz, *.json

The reference model is built through its own `ModelManager.get` (weights_uri=None), loaded with the
seeded state_dict of `focoos_b200.utils.seeded_weights` and run through its own
`processor.preprocess -> model.forward -> processor.postprocess` on seeded synthetic images
(SURVEY.md §8d).  Intermediate tensors are captured with forward hooks / a recording torch.topk.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from focoos_b200.utils.seeded_weights import seeded_state_dict  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def state_dict_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def synth_images(seed: int, sizes):
    rng = np.random.default_rng(seed)
    # smooth-ish content (random low-res field upsampled + noise) so that resizes are non-trivial
    out = []
    for (h, w) in sizes:
        base = rng.integers(0, 256, (max(h // 16, 1), max(w // 16, 1), 3)).astype(np.float32)
        up = np.kron(base, np.ones((16, 16, 1), dtype=np.float32))[:h, :w]
        if up.shape[0] < h or up.shape[1] < w:
            up = np.pad(up, ((0, h - up.shape[0]), (0, w - up.shape[1]), (0, 0)), mode="edge")
        noise = rng.integers(-40, 41, (h, w, 3)).astype(np.float32)
        out.append(np.clip(up + noise, 0, 255).astype(np.uint8))
    return out


class TopkRecorder:
    def __init__(self):
        self.calls = []
        self._orig = torch.topk

    def __enter__(self):
        def rec(inp, k, *a, **kw):
            r = self._orig(inp, k, *a, **kw)
            self.calls.append((tuple(inp.shape), r.indices.clone(), r.values.clone()))
            return r

        torch.topk = rec
        return self

    def __exit__(self, *exc):
        torch.topk = self._orig


def run_case(fm, images, threshold, tag, full: bool):
    m, proc = fm.model, fm.processor
    taps = {}

    def hook(name):
        def f(mod, inp, out):
            taps[name] = out.detach().clone() if isinstance(out, torch.Tensor) else out
        return f

    hs = []
    names = {
        "pixel_decoder.backbone.conv1": "stem",
        "pixel_decoder.backbone.res_layers.0": "res2",
        "pixel_decoder.backbone.res_layers.1": "res3",
        "pixel_decoder.backbone.res_layers.2": "res4",
        "pixel_decoder.backbone.res_layers.3": "res5",
        "pixel_decoder.encoder.0": "aifi",
        "pixel_decoder.fpn_blocks.0": "fpn0",
        "pixel_decoder.fpn_blocks.1": "fpn1",
        "pixel_decoder.pan_blocks.0": "pan0",
        "pixel_decoder.pan_blocks.1": "pan1",
        "head.predictor.enc_output": "output_memory",
        "head.predictor.decoder.layers.0": "dec0_out",
        "head.predictor.decoder.layers.5": "dec5_out",
        "head.predictor.dec_score_classifier.5": "pred_logits",
    }
    mods = dict(m.named_modules())
    for n, t in names.items():
        hs.append(mods[n].register_forward_hook(hook(t)))
    with torch.no_grad(), TopkRecorder() as rec:
        x, _ = proc.preprocess(images, device=torch.device("cpu"), dtype=torch.float32)
        out = m(x)
        dets = proc.postprocess(out, images, class_names=[], threshold=threshold)
    for h in hs:
        h.remove()
    B = x.shape[0]
    enc_topk = [c for c in rec.calls if c[0] == (B, 8400)]
    assert len(enc_topk) == 1
    post_topk = [c for c in rec.calls if c[0] == (300 * out.logits.shape[-1],)]
    assert len(post_topk) == B
    g = {
        "scores": out.logits.numpy(),
        "boxes": out.boxes.numpy(),
        "enc_topk_ind": enc_topk[0][1].numpy().astype(np.int32),
        "enc_topk_val": enc_topk[0][2].numpy(),
        "post_topk_ind": np.stack([c[1].numpy() for c in post_topk]).astype(np.int32),
        "pre_image_mean": x.mean(dim=(2, 3)).numpy(),
        "pre_image_patch": x[:, :, 100:108, 200:208].numpy(),
        "det_count": np.array([len(d.detections) for d in dets], dtype=np.int32),
        "image_sizes": np.array([im.shape[:2] for im in images], dtype=np.int32),
        "threshold": np.float32(threshold),
    }
    nmax = max(1, int(g["det_count"].max()))
    db = np.zeros((B, nmax, 4), np.int32)
    ds = np.zeros((B, nmax), np.float32)
    dl = np.full((B, nmax), -1, np.int32)
    for i, d in enumerate(dets):
        for j, det in enumerate(d.detections):
            db[i, j] = det.bbox
            ds[i, j] = det.conf
            dl[i, j] = det.cls_id
    g.update(det_boxes=db, det_scores=ds, det_labels=dl)
    if full:  # channel-sliced intermediates, NCHW as the reference lays them out
        for t in ("stem", "res2", "res3", "res4", "res5", "fpn0", "fpn1", "pan0", "pan1"):
            v = taps[t]
            g["tap_" + t] = v[:, :: max(1, v.shape[1] // 8)][:, :8, :: max(1, v.shape[2] // 20), :: max(1, v.shape[3] // 20)].numpy()
            g["tapstat_" + t] = np.array([v.mean().item(), v.std().item(), v.abs().max().item()], np.float32)
        g["tap_aifi"] = taps["aifi"][:, ::25, ::8].numpy()
        g["tap_output_memory"] = taps["output_memory"][:, ::97, ::4].numpy()
        g["tap_dec0_out"] = taps["dec0_out"][:, :, ::8].numpy()
        g["tap_dec5_out"] = taps["dec5_out"][:, :, ::8].numpy()
        g["pred_logits_raw"] = taps["pred_logits"].numpy()
    np.savez_compressed(os.path.join(GOLDEN, f"{tag}.npz"), **g)
    return g, x, out


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    name = "fai-detr-l-obj365"
    fm = ref_import.get_reference_model(name)
    template = fm.model.state_dict()
    manifest = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in template.items()}
    with open(os.path.join(GOLDEN, "fai_detr_l_obj365_state_dict_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=0, sort_keys=True)
    sd = seeded_state_dict(template, seed=0)
    res = fm.model.load_state_dict(sd, strict=True)
    fm.model.eval()
    meta = {"model": name, "weights_seed": 0, "weights_sha256": state_dict_digest(sd), "torch": torch.__version__,
            "reference_version": "0.25.0", "cases": {}}

    # case A: B=2, 640x640 originals (no resize), intermediates kept
    imgs = synth_images(1, [(640, 640), (640, 640)])
    g, x, out = run_case(fm, imgs, 0.5, "detr_l_obj365_b2_640", full=True)
    meta["cases"]["detr_l_obj365_b2_640"] = {"image_seed": 1, "sizes": [[640, 640]] * 2, "threshold": 0.5,
                                               "det_count": g["det_count"].tolist()}
    # case B: ragged original sizes -> processor resize + post-process rescale; low threshold
    sizes = [(375, 500), (720, 1280), (640, 640)]
    imgs2 = synth_images(2, sizes)
    g2, _, _ = run_case(fm, imgs2, 0.45, "detr_l_obj365_b3_ragged", full=False)
    meta["cases"]["detr_l_obj365_b3_ragged"] = {"image_seed": 2, "sizes": [list(s) for s in sizes], "threshold": 0.45,
                                                  "det_count": g2["det_count"].tolist()}
    with open(os.path.join(GOLDEN, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta, indent=1))

    # side-by-side: oracle restatement vs the reference on case A
    from oracle import detr_oracle as O

    cfg = O.DetrOracleConfig()
    taps = {}
    with torch.no_grad():
        xs = O.detr_preprocess(imgs, (640, 640))
        s, b = O.detr_forward(sd, xs, cfg, taps)
    print("oracle vs reference: scores max|d| =", (s - out.logits).abs().max().item(), " boxes max|d| =", (b - out.boxes).abs().max().item())
    print("enc topk identical:", bool((taps["topk_ind"].numpy() == g["enc_topk_ind"]).all()))
    valid = O.generate_anchors([(20, 20), (40, 40), (80, 80)])[1][0, :, 0]
    sel_invalid = (~valid[torch.from_numpy(g["enc_topk_ind"]).long()]).sum().item()
    print("selected invalid anchors:", sel_invalid)
    ev = torch.from_numpy(g["enc_topk_val"])
    print("enc top-k min gap between consecutive kept scores:", (ev[:, :-1] - ev[:, 1:]).min().item())


if __name__ == "__main__":
    main()
