"""The FocoosModel API north_star names beyond `infer` - export / train / eval, tensor batches - exercised on a GPU-less machine with the CPU reference
operators installed as the ops backend (host logic only; the CUDA kernels are covered by the -m gpu tests)."""
import os

import numpy as np
import pytest
import torch

from focoos_b200 import DETRConfig, DETRProcessor, FAIDetr, FocoosModel, ModelInfo, ops
from focoos_b200.trainer import BoxAPEvaluator, SyntheticDetectionDataset, TrainerArgs, lr_factor, training_batch
from oracle.ops_ref import RefBackend
from tests.parity_utils import seeded_sd


@pytest.fixture()
def ref_backend():
    ops._backend = RefBackend()
    yield
    ops._backend = None


def _fm(size=128, num_classes=365, precision="fp32"):
    m = FAIDetr(DETRConfig(num_classes=num_classes), precision=precision)
    sd = seeded_sd(0)
    if num_classes != 365:
        sd = {k: v for k, v in sd.items() if tuple(v.shape) == tuple(m.state_dict()[k].shape)}
    m.load_state_dict(sd)
    return FocoosModel(m, ModelInfo(name="fai-detr-l-obj365", im_size=size, config={"num_classes": num_classes}))


def test_export_torchscript_roundtrip_is_bit_identical(ref_backend, tmp_path):
    """FocoosModel.export (focoos_model.py:418-573): trace -> model.pt (+ model_info.json) -> torch.jit.load -> same tensors as the eager model, and the
    exported graph is ONE focoos_b200::model_forward call over the module's own weights (self-contained file)."""
    fm = _fm()
    im = fm.export(out_dir=str(tmp_path), device="cpu", image_size=128)
    assert sorted(os.listdir(tmp_path)) == ["model.pt", "model_info.json"]
    x = 128 * torch.randn(1, 3, 128, 128)
    eager = fm.model(x)
    loaded = torch.jit.load(str(tmp_path / "model.pt"))
    boxes, logits = loaded(x)
    assert torch.equal(boxes, eager.boxes) and torch.equal(logits, eager.logits)
    g = str(loaded.graph)
    assert g.count("focoos_b200::model_forward") == 1 and "aten::conv" not in g and "aten::_convolution" not in g
    assert len(list(loaded.parameters())) + len(list(loaded.buffers())) >= len(fm.model.state_dict()), "the weights travel inside the file"
    # the InferModel serves it through the processor's export_postprocess (infer_model.py:223-262)
    img = np.random.default_rng(0).integers(0, 256, (100, 140, 3), dtype=np.uint8)
    d1, d2 = im.infer(img, threshold=0.3), fm.infer(img, threshold=0.3)
    assert [(d.cls_id, d.bbox) for d in d1.detections] == [(d.cls_id, d.bbox) for d in d2.detections]
    with pytest.raises(ValueError):
        fm.export(runtime_type="onnx_cuda32", out_dir=str(tmp_path))


def test_model_forward_has_a_fake_kernel():
    """register_fake: shape propagation without running anything (meta tensors)"""
    from focoos_b200 import export as E
    m = FAIDetr(DETRConfig(num_classes=80))
    meta = E.make_meta(m)
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        x = torch.empty(4, 3, 640, 640)
        out = torch.ops.focoos_b200.model_forward(x, [torch.empty(1)], meta)
    assert [tuple(t.shape) for t in out] == [(4, 300, 4), (4, 300, 80)]


def test_tensor_batches_are_batches(ref_backend):
    """a [B,3,H,W] tensor / a uint8 [B,H,W,3] tensor is a batch of B images, resized like a list of the same images"""
    proc = DETRProcessor(DETRConfig(), image_size=64)
    rng = np.random.default_rng(1)
    imgs = [rng.integers(0, 256, (48, 80, 3), dtype=np.uint8) for _ in range(3)]
    a, _ = proc.preprocess(imgs, device=torch.device("cpu"))
    b, _ = proc.preprocess(torch.from_numpy(np.stack(imgs)).permute(0, 3, 1, 2).float(), device=torch.device("cpu"))
    c, _ = proc.preprocess(torch.from_numpy(np.stack(imgs)), device=torch.device("cpu"))
    assert a.shape == b.shape == (3, 3, 64, 64) and torch.allclose(a, b, atol=1e-4)
    assert torch.allclose(c if c.dtype != torch.uint8 else c.permute(0, 3, 1, 2).float(), a, atol=1e-4)
    from focoos_b200.processor import get_image_sizes
    assert get_image_sizes(torch.zeros(3, 3, 48, 80)) == [(48, 80)] * 3


def test_input_size_error_is_clear(ref_backend):
    m = FAIDetr(DETRConfig())
    with pytest.raises(ValueError, match="multiple of 32"):
        m(torch.zeros(1, 3, 100, 128))


def test_train_and_eval_entry_points(ref_backend, tmp_path):
    """FocoosModel.train (focoos_model.py:221-275) for two iterations on synthetic COCO-shape data, then FocoosModel.eval: weights + model_info are written and
    reloaded, the loss is finite, the scheduler / scaler state is logged, eval returns AP numbers from eval_postprocess outputs."""
    fm = _fm(size=128, num_classes=5)
    data = SyntheticDetectionDataset(n=4, size=128, num_classes=5)
    before = {k: v.clone() for k, v in fm.model.state_dict().items()}
    args = TrainerArgs(run_name="t", output_dir=str(tmp_path), num_gpus=1, max_iters=2, batch_size=2, log_period=1, scheduler="MULTISTEP", scheduler_extra={"milestones": [0.5]})
    info = fm.train(args, data, data_val=None)
    assert os.path.exists(tmp_path / "t" / "model_final.pth") and os.path.exists(tmp_path / "t" / "model_info.json")
    hist = info["training_history"]
    assert len(hist) == 2 and all(np.isfinite(h["total_loss"]) for h in hist) and hist[-1]["step"] >= 1
    after = fm.model.state_dict()
    assert any(not torch.equal(before[k], after[k]) for k in before if before[k].is_floating_point()), "the trained weights were reloaded into the model"
    assert not fm.model.training
    metrics = fm.eval(args, data, save_json=True)
    assert set(metrics["bbox"]) == {"AP", "AP50", "AP75"} and metrics["num_images"] == 4 and os.path.exists(tmp_path / "t" / "eval_metrics.json")


def test_training_batch_and_schedule():
    data = SyntheticDetectionDataset(n=3, size=64, num_classes=7, seed=1)
    x, targets = training_batch([data[0], data[1]], torch.device("cpu"))
    assert x.shape == (2, 3, 64, 64) and x.dtype == torch.float32 and len(targets) == 2
    e = data[0]
    b = e["instances"].boxes.tensor / 64
    assert torch.allclose(targets[0].boxes[:, :2], (b[:, :2] + b[:, 2:]) / 2) and torch.allclose(targets[0].boxes[:, 2:], b[:, 2:] - b[:, :2])
    assert lr_factor(0, 100, "MULTISTEP", {"milestones": [0.5, 0.9]}) == 1.0 and abs(lr_factor(60, 100, "MULTISTEP", {"milestones": [0.5, 0.9]}) - 0.1) < 1e-12
    assert abs(lr_factor(95, 100, "MULTISTEP", {"milestones": [0.5, 0.9]}) - 0.01) < 1e-12 and lr_factor(5, 100, "FIXED", {"warmup_iters": 10, "warmup_factor": 0.0}) == 0.5


def test_box_ap_evaluator_known_answer():
    from focoos_b200.ports import Boxes, Instances
    ev = BoxAPEvaluator(2)
    gt = {"instances": Instances((100, 100), boxes=Boxes(torch.tensor([[10., 10, 50, 50], [60, 60, 90, 90]])), classes=torch.tensor([0, 1]))}
    perfect = [{"instances": Instances((100, 100), boxes=Boxes(torch.tensor([[10., 10, 50, 50], [60, 60, 90, 90]])), scores=torch.tensor([0.9, 0.8]), classes=torch.tensor([0, 1]))}]
    ev.process([gt], perfect)
    assert ev.evaluate()["bbox"]["AP"] == pytest.approx(100.0)
    ev.reset()
    shifted = [{"instances": Instances((100, 100), boxes=Boxes(torch.tensor([[10., 10, 50, 42], [0, 0, 5, 5]])), scores=torch.tensor([0.9, 0.8]), classes=torch.tensor([0, 1]))}]
    ev.process([gt], shifted)  # class 0: IoU 0.8 -> hit for thresholds .50-.80 (7 of 10); class 1: miss
    r = ev.evaluate()["bbox"]
    assert r["AP50"] == pytest.approx(50.0) and r["AP"] == pytest.approx(35.0)


def test_preprocess_resizes_a_whole_batch_like_interpolate(ref_backend):
    """DETRProcessor.preprocess with image_size set: a uint8 NHWC batch, a float NCHW batch and a list of differently sized images all come out as the float NCHW
    tensor F.interpolate(bilinear, align_corners=False) gives per image (processor/base_processor.py:284-294) - one image_resize call per batch / image."""
    import torch.nn.functional as F
    from focoos_b200 import DETRConfig, DETRProcessor

    proc = DETRProcessor(DETRConfig(), image_size=64)
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (3, 48, 80, 3), generator=g, dtype=torch.uint8)
    ref = F.interpolate(u8.permute(0, 3, 1, 2).float(), size=(64, 64), mode="bilinear", align_corners=False)
    x, _ = proc.preprocess(u8, device=torch.device("cpu"))
    assert x.dtype == torch.float32 and tuple(x.shape) == (3, 3, 64, 64) and torch.allclose(x, ref, atol=1e-4)
    x2, _ = proc.preprocess(u8.permute(0, 3, 1, 2).float().contiguous(), device=torch.device("cpu"))
    assert torch.allclose(x2, ref, atol=1e-4)
    imgs = [np.asarray(u8[0]), np.asarray(torch.randint(0, 256, (100, 60, 3), generator=g, dtype=torch.uint8))]
    x3, _ = proc.preprocess(imgs, device=torch.device("cpu"))
    assert tuple(x3.shape) == (2, 3, 64, 64) and torch.allclose(x3[0], ref[0], atol=1e-4)
    same = torch.randint(0, 256, (2, 64, 64, 3), generator=g, dtype=torch.uint8)
    x4, _ = proc.preprocess(same, device=torch.device("cpu"))
    assert x4.dtype == torch.uint8 and tuple(x4.shape) == (2, 64, 64, 3), "a batch that already has the model size stays uint8 NHWC (the stem kernel reads it directly)"
