"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-PyTorch, CPU, fp32 restatement of the reference's **mFormerV0** inference path
(eval mode: BatchNorm uses its running statistics, dropout / drop-connect / DropPath are
identities), written functionally over a ``state_dict`` whose keys and shapes are exactly
the reference's.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this file.  Parity pinning: ``tests/test_oracle_v0_vs_reference.py`` runs the
unmodified reference (build container only) on the same seeded weights / inputs and asserts
equality; ``tests/golden/make_golden_v0.py`` freezes reference logits into
``tests/golden/v0_*.npz`` for the GPU box.  ``R/`` = ``/root/reference/linnaeus``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

from .mformer_oracle import _gen_for, _meta_token

__all__ = ["ArchV0", "arch_from_config", "param_shapes", "synth_state_dict", "synth_batch", "relative_position_index",
           "forward_features", "forward"]


@dataclass
class ArchV0:
    """What R/models/mFormerV0.py:85-345 derives from the config."""
    img_size: int = 224
    in_chans: int = 3
    conv_embed_dims: tuple = (64, 96)
    conv_out_channels: tuple = (96, 192)
    conv_depths: tuple = (2, 3)
    conv_strides: tuple = ((2, 1), (1, 1, 1))
    attn_dims: tuple = (384, 768)
    attn_depths: tuple = (5, 2)
    attn_strides: tuple = ((2, 1, 1, 1, 1), (2, 1))
    heads: tuple = (8, 8)
    mlp_ratio: tuple = (4.0, 4.0)
    meta: list = field(default_factory=list)   # [(name, dim, offset)]
    tasks: list = field(default_factory=list)  # [(task, n_classes)]
    only_last_cls: bool = False

    @property
    def extra_tokens(self) -> int:
        return 1 + len(self.meta)

    def hw_after_conv(self) -> int:
        """compute_hw_after_stage0_stage1_stage2, mFormerV0.py:21-47."""
        h = self.img_size // 4
        for seq in self.conv_strides:
            for s in seq:
                h //= s
        return max(h, 1)

    def stage_hw(self) -> tuple:
        h3 = self.hw_after_conv()
        for s in self.attn_strides[0]:
            h3 //= s
        h4 = max(h3, 1)
        for s in self.attn_strides[1]:
            h4 //= s
        return max(h3, 1), max(h4, 1)


def arch_from_config(cfg, num_classes: dict) -> ArchV0:
    cs, at = cfg.MODEL.CONV_STAGES, cfg.MODEL.ATTENTION_STAGES
    img = cfg.MODEL.IMG_SIZE
    img = img if isinstance(img, int) else img[0]
    meta = []
    if "COMPONENTS" in cfg.DATA.META:
        items = []
        for name, comp in cfg.DATA.META.COMPONENTS.items():
            if comp.get("ENABLED", False) and comp.get("IDX", -1) >= 0:
                items.append((comp.get("IDX"), name, comp["DIM"]))
        items.sort(key=lambda t: t[0])
        off = 0
        for _, name, dim in items:
            meta.append((name, dim, off))
            off += dim
    heads_cfg = cfg.MODEL.CLASSIFICATION.HEADS
    return ArchV0(
        img_size=img, in_chans=cfg.MODEL.IN_CHANS,
        conv_embed_dims=tuple(cs.EMBED_DIMS), conv_out_channels=tuple(cs.OUT_CHANNELS), conv_depths=tuple(cs.DEPTHS),
        conv_strides=tuple(tuple(s) for s in cs.STRIDE_SEQS),
        attn_dims=tuple(at.EMBED_DIMS), attn_depths=tuple(at.DEPTHS), attn_strides=tuple(tuple(s) for s in at.STRIDE_SEQS),
        heads=tuple(at.NUM_HEADS), mlp_ratio=tuple(at.MLP_RATIO), meta=meta,
        tasks=[(t, num_classes[t]) for t in heads_cfg.keys()], only_last_cls=bool(cfg.MODEL.ONLY_LAST_CLS),
    )


def _bn(s, p, c):
    s[p + "weight"] = (c,)
    s[p + "bias"] = (c,)
    s[p + "running_mean"] = (c,)
    s[p + "running_var"] = (c,)
    s[p + "num_batches_tracked"] = ()


def _mbconv_shapes(s, p, inp, out):
    """MBConvBlock.__init__, R/models/blocks/mb_conv.py:131-224 (expand ratio 4, SE ratio 0.25 of the INPUT filters)."""
    oup = inp * 4
    s[p + "_expand_conv.weight"] = (oup, inp, 1, 1)
    _bn(s, p + "_bn0.", oup)
    s[p + "_depthwise_conv.weight"] = (oup, 1, 3, 3)
    _bn(s, p + "_bn1.", oup)
    sq = max(1, int(inp * 0.25))
    s[p + "_se_reduce.weight"] = (sq, oup, 1, 1)
    s[p + "_se_reduce.bias"] = (sq,)
    s[p + "_se_expand.weight"] = (oup, sq, 1, 1)
    s[p + "_se_expand.bias"] = (oup,)
    s[p + "_project_conv.weight"] = (out, oup, 1, 1)
    _bn(s, p + "_bn2.", out)


def _meta_head_shapes(s, p, dim, D):
    s[p + "0.weight"] = (D, dim)
    s[p + "0.bias"] = (D,)
    s[p + "2.weight"] = (D,)
    s[p + "2.bias"] = (D,)
    for n in ("w1", "w2"):
        s[p + f"3.{n}.weight"] = (D, D)
        s[p + f"3.{n}.bias"] = (D,)
    for n in ("norm_fn1", "norm_fn2"):
        s[p + f"3.{n}.weight"] = (D,)
        s[p + f"3.{n}.bias"] = (D,)


def _tblock_shapes(s, p, cin, D, stride, hw, heads, ratio, n_extra):
    """RelativeMHSABlock.__init__, R/models/blocks/relative_mhsa.py:268-343."""
    if stride == 2:
        s[p + "patch_embed.proj.weight"] = (D, cin, 3, 3)
        s[p + "patch_embed.proj.bias"] = (D,)
        s[p + "patch_embed.norm.weight"] = (D,)
        s[p + "patch_embed.norm.bias"] = (D,)
    for n in ("norm1", "norm2"):
        s[p + n + ".weight"] = (D,)
        s[p + n + ".bias"] = (D,)
    s[p + "attn.relative_position_bias_table"] = ((2 * hw - 1) * (2 * hw - 1) + 1, heads)
    s[p + "attn.relative_position_index"] = (hw * hw + n_extra, hw * hw + n_extra)
    s[p + "attn.qkv.weight"] = (3 * D, D)
    s[p + "attn.proj.weight"] = (D, D)
    s[p + "attn.proj.bias"] = (D,)
    hid = int(D * ratio)
    s[p + "mlp.fc1.weight"] = (hid, D)
    s[p + "mlp.fc1.bias"] = (hid,)
    s[p + "mlp.fc2.weight"] = (D, hid)
    s[p + "mlp.fc2.bias"] = (D,)


def param_shapes(a: ArchV0) -> dict[str, tuple]:
    """Every state_dict key / shape of the reference mFormerV0 in registration order (checked against the real
    state_dict in tests/test_oracle_v0_vs_reference.py)."""
    s: dict[str, tuple] = {}
    e0 = a.conv_embed_dims[0]
    stem = (3 * (e0 // 4), e0)
    s["stage_0.0.weight"] = (stem[0], a.in_chans, 3, 3)
    _bn(s, "stage_0.1.", stem[0])
    s["stage_0.3.weight"] = (stem[1], stem[0], 3, 3)
    _bn(s, "stage_0.4.", stem[1])
    s["stage_0.6.weight"] = (e0, stem[1], 3, 3)
    _bn(s, "bn1.", e0)
    cin = e0
    for si in range(2):
        out = a.conv_out_channels[si]
        for i in range(a.conv_depths[si]):
            _mbconv_shapes(s, f"stage_{si + 1}.{i}.", cin if i == 0 else out, out)
        cin = out
    hw3, hw4 = a.stage_hw()
    D3, D4 = a.attn_dims
    s["cls_token_1"] = (1, 1, D3)
    for name, dim, _ in a.meta:
        _meta_head_shapes(s, f"meta_{name.lower()}_head_1.", dim, D3)
    for i in range(a.attn_depths[0]):
        _tblock_shapes(s, f"stage_3.{i}.", a.conv_out_channels[-1] if i == 0 else D3, D3, a.attn_strides[0][i], hw3, a.heads[0], a.mlp_ratio[0],
                       a.extra_tokens)
    s["norm_1.weight"] = (D3,)
    s["norm_1.bias"] = (D3,)
    s["cls_token_2"] = (1, 1, D4)
    for name, dim, _ in a.meta:
        _meta_head_shapes(s, f"meta_{name.lower()}_head_2.", dim, D4)
    for i in range(a.attn_depths[1]):
        _tblock_shapes(s, f"stage_4.{i}.", D3 if i == 0 else D4, D4, a.attn_strides[1][i], hw4, a.heads[1], a.mlp_ratio[1], a.extra_tokens)
    s["norm_2.weight"] = (D4,)
    s["norm_2.bias"] = (D4,)
    if not a.only_last_cls:
        s["cl_1_fc.0.fc1.weight"] = (D3, D3)
        s["cl_1_fc.0.fc1.bias"] = (D3,)
        s["cl_1_fc.0.fc2.weight"] = (D4, D3)
        s["cl_1_fc.0.fc2.bias"] = (D4,)
        s["cl_1_fc.1.weight"] = (D4,)
        s["cl_1_fc.1.bias"] = (D4,)
        s["aggregate.weight"] = (1, 2, 1)
        s["aggregate.bias"] = (1,)
    s["norm.weight"] = (D4,)
    s["norm.bias"] = (D4,)
    for t, c in a.tasks:
        s[f"head.{t}.fc.weight"] = (c, D4)
        s[f"head.{t}.fc.bias"] = (c,)
    return s


def relative_position_index(h: int, w: int, n_extra: int) -> torch.Tensor:
    """RelativeAttention.__init__, R/models/blocks/relative_mhsa.py:148-185: pairwise (dy, dx) offsets of the patch grid
    flattened to table rows; every pair involving an extra token shares the LAST row."""
    coords = torch.stack(torch.meshgrid([torch.arange(h), torch.arange(w)], indexing="ij")).reshape(2, -1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += h - 1
    rel[:, :, 1] += w - 1
    rel[:, :, 0] *= 2 * w - 1
    idx = rel.sum(-1)
    return F.pad(idx, (n_extra, 0, n_extra, 0), value=(2 * h - 1) * (2 * w - 1)).long()


def synth_state_dict(a: ArchV0, seed: int = 0) -> dict[str, torch.Tensor]:
    """Deterministic weights (CPU generators keyed by parameter name) chosen so that every op matters."""
    out = {}
    hw3, hw4 = a.stage_hw()
    for name, shape in param_shapes(a).items():
        g = _gen_for("v0." + name, seed)
        leaf = name.split(".")[-1]
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.int64)
        elif leaf == "relative_position_index":
            hw = hw3 if name.startswith("stage_3") else hw4
            t = relative_position_index(hw, hw, a.extra_tokens)
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            t = 0.6 + 0.8 * torch.rand(shape, generator=g)
        elif leaf == "relative_position_bias_table":
            t = 0.5 * torch.randn(shape, generator=g)
        elif name.startswith("cls_token"):
            t = 0.5 * torch.randn(shape, generator=g)
        elif name == "aggregate.weight":
            t = torch.tensor([0.6, 0.7]).view(shape) + 0.05 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s_ in shape[1:]:
                fan_in *= s_
            t = torch.randn(shape, generator=g) * (1.2 / math.sqrt(fan_in))
        out[name] = t.contiguous() if t.dtype == torch.int64 else t.float().contiguous()
    return out


def synth_batch(a: ArchV0, batch: int, seed: int = 0):
    g = torch.Generator(device="cpu")
    g.manual_seed(2000 + seed)
    x = torch.randn(batch, a.in_chans, a.img_size, a.img_size, generator=g)
    md = sum(dim for _, dim, _ in a.meta)
    meta = torch.randn(batch, md, generator=g) if md else None
    return x, meta


# ---------------------------------------------------------------------------
def _bn_eval(P, p, x, eps):
    return F.batch_norm(x, P[p + "running_mean"], P[p + "running_var"], P[p + "weight"], P[p + "bias"], False, 0.0, eps)


def _same_pad(img: int, k: int, s: int):
    """Conv2dStaticSamePadding.__init__, mb_conv.py:46-84: TF 'SAME' padding computed from the FULL image size."""
    o = math.ceil(img / s)
    pad = max((o - 1) * s + (k - 1) + 1 - img, 0)
    return pad // 2, pad - pad // 2


def _swish(x):
    return x * torch.sigmoid(x)


def _mbconv(P, p, x, img, stride, inp, out):
    """MBConvBlock._forward_impl / forward, mb_conv.py:226-288 (eval: no drop-connect)."""
    y = _swish(_bn_eval(P, p + "_bn0.", F.conv2d(x, P[p + "_expand_conv.weight"]), 0.01))
    lo, hi = _same_pad(img, 3, stride)
    y = F.pad(y, (lo, hi, lo, hi))
    y = F.conv2d(y, P[p + "_depthwise_conv.weight"], None, stride, 0, 1, y.shape[1])
    y = _swish(_bn_eval(P, p + "_bn1.", y, 0.01))
    sq = F.adaptive_avg_pool2d(y, 1)
    sq = _swish(F.conv2d(sq, P[p + "_se_reduce.weight"], P[p + "_se_reduce.bias"]))
    sq = F.conv2d(sq, P[p + "_se_expand.weight"], P[p + "_se_expand.bias"])
    y = torch.sigmoid(sq) * y
    y = _bn_eval(P, p + "_bn2.", F.conv2d(y, P[p + "_project_conv.weight"]), 0.01)
    if stride == 1 and inp == out:
        y = y + x
    return y


def _rel_attention(P, p, x, heads):
    """RelativeAttention.forward, relative_mhsa.py:201-236 (qkv without bias, scale on q, bias = table[index])."""
    B, N, C = x.shape
    qkv = F.linear(x, P[p + "qkv.weight"]).reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * ((C // heads) ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    bias = P[p + "relative_position_bias_table"][P[p + "relative_position_index"].view(-1)].view(N, N, heads).permute(2, 0, 1)
    attn = torch.softmax(attn + bias.unsqueeze(0), dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(o, P[p + "proj.weight"], P[p + "proj.bias"])


def _tblock(P, p, x, stride, heads, extras):
    """RelativeMHSABlock.forward, relative_mhsa.py:361-453 (eval)."""
    if stride == 2:
        t = F.conv2d(x, P[p + "patch_embed.proj.weight"], P[p + "patch_embed.proj.bias"], stride=2, padding=1)
        D = t.shape[1]
        t = t.flatten(2).transpose(1, 2)
        t = F.layer_norm(t, (D,), P[p + "patch_embed.norm.weight"], P[p + "patch_embed.norm.bias"], 1e-5)
        if extras:
            t = torch.cat([*extras, t], dim=1)
        x = t
    D = x.shape[-1]
    x = x + _rel_attention(P, p + "attn.", F.layer_norm(x, (D,), P[p + "norm1.weight"], P[p + "norm1.bias"], 1e-5), heads)
    h = F.layer_norm(x, (D,), P[p + "norm2.weight"], P[p + "norm2.bias"], 1e-5)
    h = F.gelu(F.linear(h, P[p + "mlp.fc1.weight"], P[p + "mlp.fc1.bias"]))
    return x + F.linear(h, P[p + "mlp.fc2.weight"], P[p + "mlp.fc2.bias"])


def _extras(P, a: ArchV0, stage: int, meta, B):
    cls = P[f"cls_token_{stage}"].expand(B, -1, -1)
    out = [cls]
    if meta is not None:
        for name, dim, off in a.meta:
            out.append(_meta_token(P, f"meta_{name.lower()}_head_{stage}.", meta[:, off:off + dim]).unsqueeze(1))
    return out


def forward_features(P: dict, a: ArchV0, x: torch.Tensor, meta: torch.Tensor | None = None) -> torch.Tensor:
    """mFormerV0.forward_features, R/models/mFormerV0.py:499-660 (eval mode)."""
    B = x.shape[0]
    x = F.relu(_bn_eval(P, "stage_0.1.", F.conv2d(x, P["stage_0.0.weight"], None, 2, 1), 1e-5))
    x = F.relu(_bn_eval(P, "stage_0.4.", F.conv2d(x, P["stage_0.3.weight"], None, 1, 1), 1e-5))
    x = F.conv2d(x, P["stage_0.6.weight"], None, 1, 1)
    x = F.relu(_bn_eval(P, "bn1.", x, 1e-5))
    x = F.max_pool2d(x, 3, 2, 1)
    cin = a.conv_embed_dims[0]
    for si in range(2):
        out = a.conv_out_channels[si]
        for i in range(a.conv_depths[si]):
            x = _mbconv(P, f"stage_{si + 1}.{i}.", x, a.img_size, a.conv_strides[si][i], cin if i == 0 else out, out)
        cin = out
    y = x
    for i in range(a.attn_depths[0]):
        y = _tblock(P, f"stage_3.{i}.", y, a.attn_strides[0][i], a.heads[0], _extras(P, a, 1, meta, B) if i == 0 else None)
    D3, D4 = a.attn_dims
    y = F.layer_norm(y, (D3,), P["norm_1.weight"], P["norm_1.bias"], 1e-5)
    if not a.only_last_cls:
        c1 = y[:, 0:1, :]
        c1 = F.linear(F.gelu(F.linear(c1, P["cl_1_fc.0.fc1.weight"], P["cl_1_fc.0.fc1.bias"])), P["cl_1_fc.0.fc2.weight"], P["cl_1_fc.0.fc2.bias"])
        c1 = F.layer_norm(c1, (D4,), P["cl_1_fc.1.weight"], P["cl_1_fc.1.bias"], 1e-5)
    hw3, _ = a.stage_hw()
    x = y[:, a.extra_tokens:, :].reshape(B, hw3, hw3, -1).permute(0, 3, 1, 2).contiguous()
    for i in range(a.attn_depths[1]):
        x = _tblock(P, f"stage_4.{i}.", x, a.attn_strides[1][i], a.heads[1], _extras(P, a, 2, meta, B) if i == 0 else None)
    x = F.layer_norm(x, (D4,), P["norm_2.weight"], P["norm_2.bias"], 1e-5)
    c2 = x[:, 0:1, :]
    if not a.only_last_cls:
        agg = F.conv1d(torch.cat([c1, c2], dim=1), P["aggregate.weight"], P["aggregate.bias"]).squeeze(1)
        return F.layer_norm(agg, (D4,), P["norm.weight"], P["norm.bias"], 1e-5)
    return c2.squeeze(1)


def forward(P: dict, a: ArchV0, x: torch.Tensor, meta: torch.Tensor | None = None) -> dict[str, torch.Tensor]:
    feats = forward_features(P, a, x, meta)
    return {t: F.linear(feats, P[f"head.{t}.fc.weight"], P[f"head.{t}.fc.bias"]) for t, _ in a.tasks}
