"""BisenetFormer (real-time semantic segmenter) — host-side mirror of `focoos/models/bisenetformer/modelling.py` and
`focoos/nn/backbone/stdc.py`.  Parameter containers under the reference's state_dict keys
(505 entries for bisenetformer-l-ade) + `BisenetEngine`, a fused NHWC graph:

  * STDC backbone: ConvX = conv + folded BN + ReLU in one kernel; CatBottleneck is CONCAT-FREE — every branch writes its channel
    slice of the block's output buffer (widths C/2, C/4, C/8, C/8) and the next 3x3 conv reads that slice in place; stride-2 blocks
    use the depthwise 3x3/s2 + BN kernel and the 3x3/s2 average-pool skip (nn/backbone/stdc.py:153-172),
  * context path: ARM = 1x1 proj -> 3x3 ConvBNReLU -> global average pool -> 1x1 + BN + sigmoid gate (a [B,C] GEMM with the
    sigmoid in its epilogue) -> one gating kernel that also adds the global-context vector / the upsampled coarser level
    (bisenetformer/modelling.py:159-210),
  * spatial path + FFM: proj1(res3) + proj2(cp8) as one conv with the other as residual, 1x1 ConvBNReLU, gate, `feat*att + feat`
    in the gating kernel, conv_out (:224-235,266),
  * the 2-level masked transformer decoder, head and x8 sigmoid+bilinear upsample shared with the MaskFormer family.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, fields
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from .fai_detr import _bn_fold, _Conv, _CriterionStub, _Linear
from .fai_mf import MaskFormerModelOutput, MFEngine, PredictionHeads, _AttnLayer, _ConvBN, _FFNLayer
from .ports import ModelOutput


@dataclass
class STDCConfig:
    """nn/backbone/stdc.py:175-186."""

    in_chans: int = 3
    base: int = 64
    layers: List[int] = field(default_factory=lambda: [4, 5, 3])
    out_features: List[str] = field(default_factory=lambda: ["res2", "res3", "res4", "res5"])
    model_type: str = "stdc"
    block_num: int = 4
    block_type: str = "cat"
    backbone_url: Optional[str] = None
    size: Optional[str] = None
    use_conv_last: bool = False
    use_pretrained: bool = False


@dataclass
class BisenetFormerConfig:
    """models/bisenetformer/config.py (fields of the registry JSON)."""

    backbone_config: STDCConfig = field(default_factory=STDCConfig)
    num_classes: int = 150
    num_queries: int = 100
    pixel_mean: List[float] = field(default_factory=lambda: [123.675, 116.28, 103.53])
    pixel_std: List[float] = field(default_factory=lambda: [58.395, 57.12, 57.375])
    size_divisibility: int = 0
    pixel_decoder_out_dim: int = 128
    pixel_decoder_feat_dim: int = 128
    transformer_predictor_out_dim: int = 128
    transformer_predictor_hidden_dim: int = 256
    transformer_predictor_dec_layers: int = 6
    transformer_predictor_dim_feedforward: int = 1024
    head_out_dim: int = 128
    cls_sigmoid: bool = False
    postprocessing_type: str = "semantic"
    top_k: int = 100
    mask_threshold: float = 0.5
    predict_all_pixels: bool = True
    use_mask_score: bool = False
    threshold: float = 0.5
    resolution: Optional[int] = None
    criterion_deep_supervision: bool = True
    criterion_eos_coef: float = 0.1
    criterion_num_points: int = 12544
    weight_dict_loss_dice: int = 5
    weight_dict_loss_mask: int = 5
    weight_dict_loss_ce: int = 2
    matcher_cost_class: int = 2
    matcher_cost_mask: int = 5
    matcher_cost_dice: int = 5

    @classmethod
    def from_dict(cls, d: dict) -> "BisenetFormerConfig":
        d = dict(d)
        bc = d.pop("backbone_config", {}) or {}
        if isinstance(bc, dict):
            bc = STDCConfig(**{k: v for k, v in bc.items() if k in {f.name for f in fields(STDCConfig)}})
        unknown = set(d) - {f.name for f in fields(cls)}
        if unknown:
            raise ValueError(f"Invalid parameters for BisenetFormerConfig: {sorted(unknown)}")
        return cls(backbone_config=bc, **d)


BisenetFormerOutput = MaskFormerModelOutput  # models/bisenetformer/ports.py: same fields (masks, logits, loss)


# ---- parameter containers -------------------------------------------------------------------------
class ConvX(nn.Module):  # nn/backbone/stdc.py:20 — also ConvBNReLU (bisenetformer/modelling.py:122)
    def __init__(self, cin, cout, k=3, stride=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, padding=k // 2, bias=False)
        self.bn = nn.BatchNorm2d(cout)


class CatBottleneck(nn.Module):  # nn/backbone/stdc.py:109
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.stride = stride
        if stride == 2:
            self.avd_layer = nn.Sequential(nn.Conv2d(cout // 2, cout // 2, 3, 2, 1, groups=cout // 2, bias=False), nn.BatchNorm2d(cout // 2))
        self.conv_list = nn.ModuleList([ConvX(cin, cout // 2, 1), ConvX(cout // 2, cout // 4), ConvX(cout // 4, cout // 8), ConvX(cout // 8, cout // 8)])


class STDC(nn.Module):  # nn/backbone/stdc.py:189
    def __init__(self, cfg: STDCConfig):
        super().__init__()
        assert cfg.block_type == "cat" and cfg.block_num == 4, "focoos_b200 implements the CatBottleneck STDC (block_num 4)"
        base, feats = cfg.base, []
        feats += [ConvX(cfg.in_chans, base // 2, 3, 2), ConvX(base // 2, base, 3, 2)]
        for i, n in enumerate(cfg.layers):
            for j in range(n):
                if i == 0 and j == 0:
                    feats.append(CatBottleneck(base, base * 4, 2))
                elif j == 0:
                    feats.append(CatBottleneck(base * 2 ** (i + 1), base * 2 ** (i + 2), 2))
                else:
                    feats.append(CatBottleneck(base * 2 ** (i + 2), base * 2 ** (i + 2), 1))
        self.features = nn.Sequential(*feats)
        self.out_channels = [base, base * 4, base * 8, base * 16]


class AttentionRefinementModule(nn.Module):  # bisenetformer/modelling.py:149
    def __init__(self, cin, cout):
        super().__init__()
        self.proj = nn.Conv2d(cin, cout, 1, bias=False)
        self.conv = ConvX(cout, cout, 3)
        self.conv_atten = nn.Conv2d(cout, cout, 1, bias=False)
        self.bn_atten = nn.BatchNorm2d(cout)


class ContextPath(nn.Module):  # bisenetformer/modelling.py:170
    def __init__(self, ch, d):
        super().__init__()
        self.arm32 = AttentionRefinementModule(ch[3], d)
        self.conv_avg = ConvX(ch[3], d, 1)
        self.conv_head32 = ConvX(d, d, 3)
        self.arm16 = AttentionRefinementModule(ch[2], d)
        self.conv_head16 = ConvX(d, d, 3)


class FeatureFusionModule(nn.Module):  # bisenetformer/modelling.py:213
    def __init__(self, c1, c2, cout):
        super().__init__()
        self.proj1, self.proj2 = nn.Conv2d(c1, cout, 1), nn.Conv2d(c2, cout, 1)
        self.convblk = ConvX(cout, cout, 1)
        self.conv1 = nn.Conv2d(cout, cout // 4, 1, bias=False)
        self.conv2 = nn.Conv2d(cout // 4, cout, 1, bias=False)


class BiseNet(nn.Module):  # bisenetformer/modelling.py:238
    def __init__(self, backbone: STDC, feat_dim, out_dim):
        super().__init__()
        self.backbone = backbone
        ch = backbone.out_channels
        self.cp = ContextPath(ch, feat_dim)
        self.ffm = FeatureFusionModule(ch[1], feat_dim, feat_dim)
        self.conv_out = ConvX(feat_dim, out_dim, 3)


class TransformerDecoder(nn.Module):  # bisenetformer/modelling.py:285 (two feature levels)
    def __init__(self, in_ch, out_dim, num_classes, d, num_queries, nhead, dff, layers):
        super().__init__()
        self.transformer_self_attention_layers = nn.ModuleList([_AttnLayer(d, nhead, "self_attn") for _ in range(layers)])
        self.transformer_cross_attention_layers = nn.ModuleList([_AttnLayer(d, nhead, "multihead_attn") for _ in range(layers)])
        self.transformer_ffn_layers = nn.ModuleList([_FFNLayer(d, dff) for _ in range(layers)])
        self.query_feat, self.query_embed = nn.Embedding(num_queries, d), nn.Embedding(num_queries, d)
        self.input_proj = nn.ModuleList([_ConvBN(in_ch, d, 1, bias=True, norm=False) for _ in range(2)])
        self.forward_prediction_heads = PredictionHeads(d, num_classes, out_dim)


class BisenetFormerHead(nn.Module):
    def __init__(self, predictor, num_classes):
        super().__init__()
        self.criterion = _CriterionStub(num_classes)
        self.predictor = predictor


class BisenetEngine(MFEngine):
    def __init__(self, sd: Dict[str, torch.Tensor], cfg: BisenetFormerConfig, device, precision: str = "fp16", algo: int = ops.ALGO_AUTO):
        self.cfg, self.device, self.precision, self.algo = cfg, torch.device(device), precision, algo
        assert precision in ("fp32", "fp16", "fp32_tc")
        self.dt = torch.float16 if precision == "fp16" else torch.float32
        self._host_w3 = {} if precision == "fp32_tc" else None  # fp32 storage, three fp16 tensor-core products per conv / linear (fai_detr._split3_weights)
        self.nhead, self.d = 8, cfg.transformer_predictor_hidden_dim
        self._consts = {}
        # pack on the HOST (BN folding, re-parameterisation, concatenations are a few hundred tiny tensor ops: as device launches they were ~700 `at::`
        # kernels in front of the first forward); only the packed tensors travel to the device
        sd = {k: v.detach().to("cpu") for k, v in sd.items()}
        bb = "pixel_decoder.backbone.features"
        w = sd[bb + ".0.conv.weight"].float()
        s, b = _bn_fold(sd, bb + ".0.bn")
        self.stem_w, self.stem_s, self.stem_b = self._f32(w.permute(0, 2, 3, 1)), self._f32(s), self._f32(b)
        self.stem2 = self._convx(sd, bb + ".1", 2)
        self.blocks = []
        idx = 2
        for n in cfg.backbone_config.layers:
            stage = []
            for j in range(n):
                p = f"{bb}.{idx}"
                stride = 2 if j == 0 else 1
                blk = {"stride": stride, "convs": [self._convx(sd, f"{p}.conv_list.{i}", 1) for i in range(4)]}
                if stride == 2:
                    wd = sd[p + ".avd_layer.0.weight"].float()  # [C,1,3,3]
                    sa, ba = _bn_fold(sd, p + ".avd_layer.1")
                    blk["avd"] = (self._f32(wd.reshape(wd.shape[0], 9).t()), self._f32(sa), self._f32(ba))
                stage.append(blk)
                idx += 1
            self.blocks.append(stage)
        cp, ffm = "pixel_decoder.cp", "pixel_decoder.ffm"
        self.conv_avg = self._convx(sd, cp + ".conv_avg", 1)
        self.arm = {}
        for name in ("arm32", "arm16"):
            q = f"{cp}.{name}"
            sa, ba = _bn_fold(sd, q + ".bn_atten")
            self.arm[name] = {"proj": _Conv(self._to(sd[q + ".proj.weight"].float().permute(0, 2, 3, 1)), None, None, 1, 0, ops.ACT_NONE),
                              "conv": self._convx(sd, q + ".conv", 1),
                              "att": _Conv(self._to(sd[q + ".conv_atten.weight"].float().permute(0, 2, 3, 1)), self._f32(sa), self._f32(ba), 1, 0, ops.ACT_SIGMOID)}
        self.head32, self.head16 = self._convx(sd, cp + ".conv_head32", 1), self._convx(sd, cp + ".conv_head16", 1)
        self.ffm_p1 = _Conv(self._to(sd[ffm + ".proj1.weight"].float().permute(0, 2, 3, 1)), None, self._f32(sd[ffm + ".proj1.bias"]), 1, 0, ops.ACT_NONE)
        self.ffm_p2 = _Conv(self._to(sd[ffm + ".proj2.weight"].float().permute(0, 2, 3, 1)), None, self._f32(sd[ffm + ".proj2.bias"]), 1, 0, ops.ACT_NONE)
        self.ffm_blk = self._convx(sd, ffm + ".convblk", 1)
        self.ffm_c1 = _Conv(self._to(sd[ffm + ".conv1.weight"].float().permute(0, 2, 3, 1)), None, None, 1, 0, ops.ACT_RELU)
        self.ffm_c2 = _Conv(self._to(sd[ffm + ".conv2.weight"].float().permute(0, 2, 3, 1)), None, None, 1, 0, ops.ACT_SIGMOID)
        self.conv_out = self._convx(sd, "pixel_decoder.conv_out", 1)
        self._pack_decoder(sd, 2)
        self._finish_pack()

    def _convx(self, sd, p, stride):
        w = sd[p + ".conv.weight"].float()
        s, b = _bn_fold(sd, p + ".bn")
        return _Conv(self._to(w.permute(0, 2, 3, 1)), self._f32(s), self._f32(b), stride, w.shape[-1] // 2, ops.ACT_RELU)

    def _gate(self, conv, vec):
        """tiny [B,C] GEMM(s) of the channel-attention gates, SIMT path (M = batch size)."""
        B, C = vec.shape
        return conv(vec.reshape(B, 1, 1, C), algo=ops.ALGO_SIMT).reshape(B, -1)

    def _cat_bottleneck(self, x, blk, first_in_pair=False):
        A, dt = self.algo, self.dt
        c = blk["convs"]
        half = c[0].w.shape[0]
        Cout = half * 2
        B, H, W, _ = x.shape
        if first_in_pair and blk["stride"] != 2:  # a stride-1 block the pair path does not take: back to fp32 first
            x, first_in_pair = x.float(), False
        if blk["stride"] == 2:
            out1 = self._pc(c[0], x, out_pair=False) if first_in_pair else c[0](x, algo=A)
            buf = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, Cout), dtype=dt, device=x.device)
            ops.avgpool3x3s2(out1, out=buf[..., :half])
            src = ops.dwconv3x3s2(out1, *blk["avd"])
        else:
            buf = torch.empty((B, H, W, Cout), dtype=dt, device=x.device)
            c[0](x, out=buf[..., :half], algo=A)
            src = buf[..., :half]
        o = half
        for i in (1, 2, 3):
            w = c[i].w.shape[0]
            c[i](src, out=buf[..., o:o + w], algo=A)
            src = buf[..., o:o + w]
            o += w
        return buf

    def _cat_bottleneck_pair(self, x, blk):
        """stride-1 CatBottleneck in the pair format (fp32_tc): the four convs read and write fp16 [hi | lo] planes - each one's output is a channel slice of the block's
        concat buffer and the next one's input - so no split pass runs inside the block.  x: Pair or fp32 tensor (split once) -> Pair"""
        c = blk["convs"]
        half = c[0].w.shape[0]
        B, H, W, _ = x.shape
        buf = ops.Pair.empty((B, H, W, 2 * half), x.device)
        src = self._pc(c[0], x, out=buf.slice(0, half))
        o = half
        for i in (1, 2, 3):
            w = c[i].w.shape[0]
            src = self._pc(c[i], src, out=buf.slice(o, o + w))
            o += w
        return buf

    def _pair_block_ok(self, blk, H, W) -> bool:
        """conv2d_pair takes the block: every conv has its weight triple; a 32-channel 3x3 input needs the halo mode (rows of at least 64 pixels, Cout <= 64)"""
        if blk["stride"] != 1 or not self.pair_capable():
            return False
        for cv in blk["convs"]:
            cin, cout, k = cv.w.shape[3], cv.w.shape[0], cv.w.shape[1]
            if cv.w3 is None or cout % 8:
                return False
            if cin % 64 and not (cin == 32 and k == 3 and W >= 64 and cout <= 64):
                return False
        return True

    @torch.no_grad()
    def forward(self, images: torch.Tensor, taps: Optional[dict] = None):
        cfg, dt, A = self.cfg, self.dt, self.algo
        if images.dtype == torch.uint8:
            B, H, W, _ = images.shape
        else:
            assert images.dim() == 4 and images.shape[1] == 3 and images.dtype == torch.float32
            B, _, H, W = images.shape
        if H % 32 or W % 32:
            # the reference's torch graph takes any size (odd feature maps from ceil-mode pools / stride-2 convs); the B200 kernels tile the stride-2 layers on even
            # maps, so the engine takes multiples of 32 - resize or pad in the processor (image_size) for other inputs
            raise ValueError(f"focoos_b200: input size {H}x{W} is not a multiple of 32; resize/pad the image (e.g. ModelInfo.im_size) before the model")
        x = ops.stem_conv(images.contiguous(), self.stem_w, self.stem_s, self.stem_b, cfg.pixel_mean, cfg.pixel_std, ops.ACT_RELU, dt)
        x = self.stem2(x, algo=A)  # res2
        feats = []
        P = ops.Pair
        as_f32 = lambda t: t.float() if isinstance(t, P) else t  # noqa: E731  (a torch op: only ever applied to the small 1/32-resolution map below)
        for stage in self.blocks:
            for blk in stage:
                if self._pair_block_ok(blk, x.shape[1], x.shape[2]):
                    x = self._cat_bottleneck_pair(x, blk)
                elif isinstance(x, P):  # a stride-2 block after pair-native ones: its first conv reads the pair and writes fp32, the rest runs as before
                    x = self._cat_bottleneck(x, blk, first_in_pair=True)
                else:
                    x = self._cat_bottleneck(x, blk)
            feats.append(x)
        res3, res4, res5 = feats
        res5 = as_f32(res5)  # global average pool + ARM gates work on fp32 (33 M elements at bs=64 1024x512)
        pin = lambda conv, f, **kw: (self._pc(conv, f, out_pair=False, **kw) if isinstance(f, P) else conv(f, algo=A, **kw))  # noqa: E731
        # context path
        avg = self._gate(self.conv_avg, ops.global_avgpool(res5))
        a = self.arm["arm32"]
        f = a["conv"](a["proj"](res5, algo=A), algo=A)
        f32 = ops.channel_scale(f, self._gate(a["att"], ops.global_avgpool(f)), addvec=avg)
        up = self.head32(ops.resize_bilinear(f32, (res4.shape[1], res4.shape[2])), algo=A)
        a = self.arm["arm16"]
        f = a["conv"](pin(a["proj"], res4), algo=A)
        f16 = ops.channel_scale(f, self._gate(a["att"], ops.global_avgpool(f)), addt=up)
        f8 = self.head16(ops.resize_bilinear(f16, (res3.shape[1], res3.shape[2])), algo=A)
        # feature fusion
        feat = self.ffm_blk(pin(self.ffm_p1, res3, residual=self.ffm_p2(f8, algo=A)), algo=A)
        att = self._gate(self.ffm_c2, self._gate(self.ffm_c1, ops.global_avgpool(feat)))
        fuse = ops.channel_scale(feat, att, self_add=True)
        mask_features = self.conv_out(fuse, algo=A)
        if taps is not None:
            taps.update(res3=as_f32(res3), res4=as_f32(res4), res5=res5, cp32=f32, cp16=f16, cp8=f8, mask_features=mask_features)
        return self._run_decoder([f32, f16], mask_features, B, H, W, taps)


class BisenetFormer(nn.Module):
    """Drop-in for the reference `BisenetFormer(BaseModelNN)` (bisenetformer/modelling.py:534)."""

    lazy_masks = False  # True: forward() returns fai_mf.LazyMasks (low-resolution logits) instead of the upsampled [B,Q,H,W] probabilities

    def __init__(self, config: BisenetFormerConfig, precision: str = "fp16"):
        super().__init__()
        self.config = c = config
        self.pixel_decoder = BiseNet(STDC(c.backbone_config), c.pixel_decoder_feat_dim, c.pixel_decoder_out_dim)
        self.head = BisenetFormerHead(TransformerDecoder(c.pixel_decoder_out_dim, c.transformer_predictor_out_dim, c.num_classes, c.transformer_predictor_hidden_dim,
                                                         c.num_queries, 8, c.transformer_predictor_dim_feedforward, c.transformer_predictor_dec_layers), c.num_classes)
        self.register_buffer("pixel_mean", torch.tensor(c.pixel_mean, dtype=torch.float32).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor(c.pixel_std, dtype=torch.float32).view(-1, 1, 1), False)
        self.num_classes, self.precision, self.algo, self._engine = c.num_classes, precision, ops.ALGO_AUTO, None
        self.eval()

    device = property(lambda self: self.pixel_mean.device)
    dtype = property(lambda self: self.pixel_mean.dtype)

    def load_state_dict(self, state_dict, strict: bool = False, assign: bool = False):
        if "model" in state_dict and isinstance(state_dict["model"], dict):
            state_dict = state_dict["model"]
        own = self.state_dict()
        filtered = {k: v for k, v in state_dict.items() if k in own and tuple(own[k].shape) == tuple(v.shape)}
        res = super().load_state_dict(filtered, strict=False)
        self._engine = None
        if strict and (res.missing_keys or len(filtered) != len(state_dict)):
            raise RuntimeError(f"load_state_dict(strict): missing {res.missing_keys[:5]} / dropped {len(state_dict) - len(filtered)}")
        return res

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def engine(self) -> BisenetEngine:
        e = self._engine
        if e is None or e.device != self.device or e.precision != self.precision or e.algo != self.algo:
            self._engine = BisenetEngine(self.state_dict(), self.config, self.device, self.precision, self.algo)
        return self._engine

    def forward(self, images: torch.Tensor, targets: list = [], taps: Optional[dict] = None) -> MaskFormerModelOutput:
        if self.training or (targets is not None and len(targets) > 0):
            raise NotImplementedError("focoos_b200: losses / fine-tuning are not part of the inference hot path")
        if ops._backend is None and not images.is_cuda:
            raise RuntimeError("focoos_b200.BisenetFormer runs on CUDA (sm_100a) only — no CPU fallback")
        eng = self.engine()
        eng.lazy_masks = bool(getattr(self, "lazy_masks", False))
        probs, masks = eng.forward(images if images.dtype == torch.uint8 else images.to(torch.float32), taps)
        return MaskFormerModelOutput(masks=masks, logits=probs, loss=None)
