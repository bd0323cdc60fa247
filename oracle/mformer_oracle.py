"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch, CPU, fp32 restatement of the reference's mFormerV1 hot path
(forward, hierarchical masked loss, clip + AdamW step), written functionally
over a ``state_dict`` whose keys and shapes are exactly the reference's.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this file.  The product
(``linnaeus_b200``) never does and has no CPU fallback.

Parity pinning: the reference ships **no** golden vectors for this path
(SURVEY.md section 4), so the pin is the reference itself executed in the build
container: ``tests/test_oracle_vs_reference.py`` imports ``/root/reference``
(through the two shims in ``tests/support/ref_shims``) and asserts this file
equals it on logits, loss, every gradient and one optimizer step;
``tests/golden/make_golden.py`` then freezes reference outputs into
``tests/golden/*.npz`` so the GPU box can check without the reference tree.

Every function cites the reference file:line it restates (``R/`` =
``/root/reference/linnaeus``).
"""

from __future__ import annotations

import hashlib
import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

__all__ = [
    "Arch",
    "arch_from_config",
    "param_shapes",
    "synth_state_dict",
    "synth_batch",
    "forward_features",
    "forward",
    "per_sample_losses",
    "hierarchical_loss",
    "split_decay",
    "adamw_clip_step",
    "train_step",
    "synthetic_taxonomy_smoothing",
]


# ---------------------------------------------------------------------------
# Architecture description (what R/models/mFormerV1.py:44-349 derives from cfg)
# ---------------------------------------------------------------------------
@dataclass
class Arch:
    img_size: int = 224
    in_chans: int = 3
    dims: tuple = (96, 192, 384, 768)
    conv_depths: tuple = (3, 3)  # only DEPTHS[:2] are ever built (mFormerV1.py:134-136,177,193)
    rope_depths: tuple = (5, 2)
    heads: tuple = (6, 12)
    mlp_ratio: tuple = (4.0, 4.0)
    # [(component name, dim, offset)] ordered by IDX (mFormerV1.py:98-113)
    meta: list = field(default_factory=list)
    # [(task key, n classes)] in MODEL.CLASSIFICATION.HEADS insertion order
    tasks: list = field(default_factory=list)
    head_type: str = "Linear"  # Linear | HierarchicalSoftmax | ConditionalClassifier
    only_last_cls: bool = False

    @property
    def extra_tokens(self) -> int:
        return 1 + len(self.meta)


def arch_from_config(cfg, num_classes: dict) -> Arch:
    """mFormerV1.__init__ config reads, R/models/mFormerV1.py:49-130,314,332-343."""
    cs, rs = cfg.MODEL.CONVNEXT_STAGES, cfg.MODEL.ROPE_STAGES
    img = cfg.MODEL.IMG_SIZE
    img = img if isinstance(img, int) else img[0]
    meta = []
    if cfg.DATA.META.get("ACTIVE", False) and "COMPONENTS" in cfg.DATA.META:
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
    tasks = [(t, num_classes[t]) for t in heads_cfg.keys()]
    types = {heads_cfg[t].get("TYPE", "Linear") for t in heads_cfg.keys()}
    head_type = types.pop() if len(types) == 1 else "Linear"
    return Arch(
        img_size=img,
        in_chans=cfg.MODEL.IN_CHANS,
        dims=tuple(cs.DIMS),
        conv_depths=tuple(cs.DEPTHS[:2]),
        rope_depths=tuple(rs.DEPTHS),
        heads=tuple(rs.NUM_HEADS),
        mlp_ratio=tuple(rs.MLP_RATIO),
        meta=meta,
        tasks=tasks,
        head_type=head_type,
        only_last_cls=bool(cfg.MODEL.ONLY_LAST_CLS),
    )


def param_shapes(a: Arch) -> dict[str, tuple]:
    """Every parameter key/shape of the reference model in registration order
    (checked against the real state_dict in tests)."""
    s: dict[str, tuple] = {}
    d = a.dims
    s["cls_token_1"] = (1, 1, d[2])
    s["cls_token_2"] = (1, 1, d[3])
    s["stem.0.weight"] = (d[0], a.in_chans, 4, 4)
    s["stem.0.bias"] = (d[0],)
    s["stem.1.weight"] = (d[0],)
    s["stem.1.bias"] = (d[0],)
    for i in range(3):
        s[f"downsample_layers.{i}.norm.weight"] = (d[i],)
        s[f"downsample_layers.{i}.norm.bias"] = (d[i],)
        s[f"downsample_layers.{i}.conv.weight"] = (d[i + 1], d[i], 2, 2)
        s[f"downsample_layers.{i}.conv.bias"] = (d[i + 1],)
    for st in range(2):
        c = d[st]
        for i in range(a.conv_depths[st]):
            p = f"stages.{st}.{i}."
            s[p + "gamma"] = (c,)
            s[p + "dwconv.weight"] = (c, 1, 7, 7)
            s[p + "dwconv.bias"] = (c,)
            s[p + "norm.weight"] = (c,)
            s[p + "norm.bias"] = (c,)
            s[p + "pwconv1.weight"] = (4 * c, c)
            s[p + "pwconv1.bias"] = (4 * c,)
            s[p + "pwconv2.weight"] = (c, 4 * c)
            s[p + "pwconv2.bias"] = (c,)
    for st in range(2):
        c = d[2 + st]
        hid = int(c * a.mlp_ratio[st])
        for i in range(a.rope_depths[st]):
            p = f"stages.{2 + st}.{i}."
            s[p + "norm1.weight"] = (c,)
            s[p + "norm1.bias"] = (c,)
            s[p + "norm2.weight"] = (c,)
            s[p + "norm2.bias"] = (c,)
            s[p + "attn.freqs"] = (2, a.heads[st], (c // a.heads[st]) // 2)
            s[p + "attn.qkv.weight"] = (3 * c, c)
            s[p + "attn.qkv.bias"] = (3 * c,)
            s[p + "attn.proj.weight"] = (c, c)
            s[p + "attn.proj.bias"] = (c,)
            s[p + "mlp.fc1.weight"] = (hid, c)
            s[p + "mlp.fc1.bias"] = (hid,)
            s[p + "mlp.fc2.weight"] = (c, hid)
            s[p + "mlp.fc2.bias"] = (c,)
    s["norm_1.weight"] = (d[2],)
    s["norm_1.bias"] = (d[2],)
    s["norm_2.weight"] = (d[3],)
    s["norm_2.bias"] = (d[3],)
    for name, dim, _ in a.meta:
        for st in (1, 2):
            c = d[1 + st]
            p = f"meta_{name.lower()}_head_{st}."
            s[p + "0.weight"] = (c, dim)
            s[p + "0.bias"] = (c,)
            s[p + "2.weight"] = (c,)
            s[p + "2.bias"] = (c,)
            for sub in ("norm_fn1", "norm_fn2"):
                s[p + f"3.{sub}.weight"] = (c,)
                s[p + f"3.{sub}.bias"] = (c,)
            for sub in ("w1", "w2"):
                s[p + f"3.{sub}.weight"] = (c, c)
                s[p + f"3.{sub}.bias"] = (c,)
    if not a.only_last_cls:
        s["cl_1_fc.0.fc1.weight"] = (d[2], d[2])
        s["cl_1_fc.0.fc1.bias"] = (d[2],)
        s["cl_1_fc.0.fc2.weight"] = (d[3], d[2])
        s["cl_1_fc.0.fc2.bias"] = (d[3],)
        s["cl_1_fc.1.weight"] = (d[3],)
        s["cl_1_fc.1.bias"] = (d[3],)
        s["aggregate.weight"] = (1, 2, 1)
        s["aggregate.bias"] = (1,)
    s["final_norm.weight"] = (d[3],)
    s["final_norm.bias"] = (d[3],)
    if a.head_type == "Linear":
        for t, c in a.tasks:
            s[f"head.{t}.fc.weight"] = (c, d[3])
            s[f"head.{t}.fc.bias"] = (c,)
    else:
        # Hierarchical heads share ONE ModuleDict of per-level Linear layers
        # (R/models/heads/utils.py:218-228); state_dict lists it under every head.
        sub = "task_classifiers" if a.head_type == "HierarchicalSoftmax" else "level_classifiers"
        for t, _ in a.tasks:
            for t2, c2 in a.tasks:
                s[f"head.{t}.{sub}.{t2}.weight"] = (c2, d[3])
                s[f"head.{t}.{sub}.{t2}.bias"] = (c2,)
    return s


# ---------------------------------------------------------------------------
# Deterministic name-hashed weights / inputs (no weights are shipped)
# ---------------------------------------------------------------------------
def _gen_for(name: str, seed: int) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


def synth_state_dict(shapes: dict[str, tuple], seed: int = 0) -> dict[str, torch.Tensor]:
    """Fill any ``{name: shape}`` identically on both sides of a parity test.
    Values are *not* the reference's init (layer-scale 1e-6 would hide the conv
    blocks); they are chosen so every op contributes visibly to the output."""
    out = {}
    shared: dict[str, torch.Tensor] = {}
    for name, shape in shapes.items():
        # hierarchical heads: every head.<t>.*_classifiers.<t2>.* aliases one tensor
        key = name
        parts = name.split(".")
        if parts[0] == "head" and len(parts) == 5:
            key = "head.*." + ".".join(parts[2:])
            if key in shared:
                out[name] = shared[key]
                continue
        g = _gen_for(key, seed)
        leaf = parts[-1]
        if leaf == "gamma":
            t = 0.25 + 0.5 * torch.rand(shape, generator=g)
        elif leaf == "freqs":
            nh, half = shape[1], shape[2]
            inv = 1.0 / (10000.0 ** (torch.arange(0, 2 * half, 2).float() / (2 * half)))
            ang = torch.rand(nh, 1, generator=g) * 2 * math.pi
            t = torch.stack([inv[None, :] * torch.cos(ang), inv[None, :] * torch.sin(ang)], 0)
        elif name.startswith("cls_token"):
            t = 0.5 * torch.randn(shape, generator=g)
        elif name == "aggregate.weight":
            t = torch.tensor([0.6, 0.7]).view(shape) + 0.05 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g)
        elif len(shape) == 1:  # norm weights
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s_ in shape[1:]:
                fan_in *= s_
            t = torch.randn(shape, generator=g) * (0.9 / math.sqrt(fan_in))
        out[name] = t.float().contiguous()
        if key != name:
            shared[key] = out[name]
    return out


def synth_batch(a: Arch, batch: int, seed: int = 0, null_frac: bool = True):
    """Synthetic inputs of SURVEY.md 8(d): images randn, aux randn, int64 targets
    (label 0 = null).  CPU generator => identical on every host."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    x = torch.randn(batch, a.in_chans, a.img_size, a.img_size, generator=g)
    md = sum(dim for _, dim, _ in a.meta)
    meta = torch.randn(batch, md, generator=g) if md else None
    targets = {}
    for t, c in a.tasks:
        lo = 0 if null_frac else 1
        targets[t] = torch.randint(lo, c, (batch,), generator=g, dtype=torch.int64)
    return x, meta, targets


# ---------------------------------------------------------------------------
# Forward
# ---------------------------------------------------------------------------
def _ln_cf(x, w, b, eps=1e-6):
    """LayerNormChannelsFirst.forward, R/models/blocks/convnext.py:32-43 (biased var)."""
    mu = x.mean(1, keepdim=True)
    var = (x - mu).pow(2).mean(1, keepdim=True)
    xh = (x - mu) / torch.sqrt(var + eps)
    return w.view(1, -1, 1, 1) * xh + b.view(1, -1, 1, 1)


def _convnext_block(P, p, x):
    """ConvNeXtBlock._forward_impl, R/models/blocks/convnext.py:73-87 (DropPath = identity)."""
    c = x.shape[1]
    y = F.conv2d(x, P[p + "dwconv.weight"], P[p + "dwconv.bias"], padding=3, groups=c)
    y = y.permute(0, 2, 3, 1)
    y = F.layer_norm(y, (c,), P[p + "norm.weight"], P[p + "norm.bias"], 1e-6)
    y = F.linear(y, P[p + "pwconv1.weight"], P[p + "pwconv1.bias"])
    y = F.gelu(y)
    y = F.linear(y, P[p + "pwconv2.weight"], P[p + "pwconv2.bias"])
    y = P[p + "gamma"] * y
    return x + y.permute(0, 3, 1, 2)


def _downsample(P, i, x):
    """ConvNeXtDownsampleLayer.forward, R/models/blocks/convnext.py:112-115."""
    p = f"downsample_layers.{i}."
    x = _ln_cf(x, P[p + "norm.weight"], P[p + "norm.bias"])
    return F.conv2d(x, P[p + "conv.weight"], P[p + "conv.bias"], stride=2)


def rope_cos_table(freqs, H, W):
    """init_t_xy + compute_mixed_cis + the real-cast in _get_current_freqs_cis
    (R/models/blocks/rope_2d_mhsa.py:56-73,114-155,404-408): the complex cis is
    cast to the dtype of the real parameter, leaving cos(theta) only.
    Returns [N_img, heads, head_dim/2]."""
    t = torch.arange(H * W, dtype=torch.float32, device=freqs.device)
    tx = t % W
    ty = torch.div(t, W, rounding_mode="floor")
    theta = tx[:, None, None] * freqs[0][None] + ty[:, None, None] * freqs[1][None]
    return torch.cos(theta)


def _attention(P, p, x, H, W, heads, n_extra):
    """RoPE2DAttention.forward standard path, R/models/blocks/rope_2d_mhsa.py:422-456,492-505."""
    B, N, C = x.shape
    hd = C // heads
    qkv = F.linear(x, P[p + "qkv.weight"], P[p + "qkv.bias"]).reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]  # [B, h, N, hd]
    cos = rope_cos_table(P[p + "freqs"], H, W).permute(1, 0, 2)  # [h, N_img, hd/2]
    cos = cos.repeat_interleave(2, dim=-1)  # each (even, odd) pair scaled by the same cos
    ones = torch.ones(heads, n_extra, hd, dtype=cos.dtype, device=cos.device)
    fac = torch.cat([ones, cos], dim=1)[None]  # extras untouched
    q = q * fac * (hd ** -0.5)
    k = k * fac
    attn = torch.softmax(q.float() @ k.float().transpose(-2, -1), dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(out, P[p + "proj.weight"], P[p + "proj.bias"])


def _rope_block(P, p, x, H, W, heads, n_extra):
    """RoPE2DMHSABlock.forward, R/models/blocks/rope_2d_mhsa.py:584-645 (LN eps 1e-5)."""
    C = x.shape[-1]
    y = F.layer_norm(x, (C,), P[p + "norm1.weight"], P[p + "norm1.bias"], 1e-5)
    x = x + _attention(P, p + "attn.", y, H, W, heads, n_extra)
    y = F.layer_norm(x, (C,), P[p + "norm2.weight"], P[p + "norm2.bias"], 1e-5)
    y = F.linear(y, P[p + "mlp.fc1.weight"], P[p + "mlp.fc1.bias"])  # Mlp.forward, R/models/blocks/mlp.py:46-66
    y = F.gelu(y)
    y = F.linear(y, P[p + "mlp.fc2.weight"], P[p + "mlp.fc2.bias"])
    return x + y


def _meta_token(P, p, m):
    """Linear -> ReLU -> LN -> ResNormLayer (R/models/mFormerV1.py:288-308;
    R/models/normalization/res_norm_layer.py:22-30)."""
    C = P[p + "0.weight"].shape[0]
    x = F.relu(F.linear(m, P[p + "0.weight"], P[p + "0.bias"]))
    x = F.layer_norm(x, (C,), P[p + "2.weight"], P[p + "2.bias"], 1e-5)
    y = F.relu(F.linear(x, P[p + "3.w1.weight"], P[p + "3.w1.bias"]))
    y = F.layer_norm(y, (C,), P[p + "3.norm_fn1.weight"], P[p + "3.norm_fn1.bias"], 1e-5)
    y = F.relu(F.linear(y, P[p + "3.w2.weight"], P[p + "3.w2.bias"]))
    y = F.layer_norm(y, (C,), P[p + "3.norm_fn2.weight"], P[p + "3.norm_fn2.bias"], 1e-5)
    return x + y


def forward_features(P: dict, a: Arch, x: torch.Tensor, meta: torch.Tensor | None = None) -> torch.Tensor:
    """mFormerV1.forward_features, R/models/mFormerV1.py:407-529."""
    B = x.shape[0]
    d = a.dims
    x = F.conv2d(x, P["stem.0.weight"], P["stem.0.bias"], stride=4)
    x = _ln_cf(x, P["stem.1.weight"], P["stem.1.bias"])
    for i in range(a.conv_depths[0]):
        x = _convnext_block(P, f"stages.0.{i}.", x)
    x = _downsample(P, 0, x)
    for i in range(a.conv_depths[1]):
        x = _convnext_block(P, f"stages.1.{i}.", x)
    x = _downsample(P, 1, x)
    H, W = x.shape[2], x.shape[3]
    x = x.flatten(2).transpose(1, 2)

    def extras(stage: int):
        toks = [P[f"cls_token_{stage}"].expand(B, -1, -1)]
        if a.meta and meta is not None:
            for name, dim, off in a.meta:
                toks.append(_meta_token(P, f"meta_{name.lower()}_head_{stage}.", meta[:, off : off + dim]).unsqueeze(1))
        return toks

    # NOTE: the attention split uses the build-time extra-token count even when
    # meta is None (mFormerV1.py:130,450); like the reference that asserts.
    n_extra = a.extra_tokens
    x = torch.cat([*extras(1), x], dim=1)
    for i in range(a.rope_depths[0]):
        x = _rope_block(P, f"stages.2.{i}.", x, H, W, a.heads[0], n_extra)
    x = F.layer_norm(x, (d[2],), P["norm_1.weight"], P["norm_1.bias"], 1e-5)
    if not a.only_last_cls:
        c1 = x[:, 0:1, :]
        c1 = F.linear(c1, P["cl_1_fc.0.fc1.weight"], P["cl_1_fc.0.fc1.bias"])
        c1 = F.gelu(c1)
        c1 = F.linear(c1, P["cl_1_fc.0.fc2.weight"], P["cl_1_fc.0.fc2.bias"])
        c1 = F.layer_norm(c1, (d[3],), P["cl_1_fc.1.weight"], P["cl_1_fc.1.bias"], 1e-5)
    x = x[:, n_extra:, :].transpose(1, 2).reshape(B, -1, H, W)
    x = _downsample(P, 2, x)
    H, W = x.shape[2], x.shape[3]
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([*extras(2), x], dim=1)
    for i in range(a.rope_depths[1]):
        x = _rope_block(P, f"stages.3.{i}.", x, H, W, a.heads[1], n_extra)
    x = F.layer_norm(x, (d[3],), P["norm_2.weight"], P["norm_2.bias"], 1e-5)
    c2 = x[:, 0:1, :]
    if not a.only_last_cls:
        # aggregate = Conv1d(2 -> 1, k=1) over the two CLS vectors (mFormerV1.py:512-524)
        w = P["aggregate.weight"].view(2)
        agg = w[0] * c1[:, 0] + w[1] * c2[:, 0] + P["aggregate.bias"]
    else:
        agg = c2[:, 0]
    return F.layer_norm(agg, (d[3],), P["final_norm.weight"], P["final_norm.bias"], 1e-5)


def forward(P: dict, a: Arch, x: torch.Tensor, meta: torch.Tensor | None = None) -> dict[str, torch.Tensor]:
    """mFormerV1.forward, R/models/mFormerV1.py:531-541.  Hierarchical head types
    return their own level's linear output (the hmatrix buffer key is
    never found, R/utils/taxonomy/taxonomy_tree.py:384-404 vs
    R/models/heads/hierarchical_softmax_head.py:164-190)."""
    f = forward_features(P, a, x, meta)
    out = {}
    for t, _ in a.tasks:
        if a.head_type == "Linear":
            w, b = P[f"head.{t}.fc.weight"], P[f"head.{t}.fc.bias"]
        else:
            sub = "task_classifiers" if a.head_type == "HierarchicalSoftmax" else "level_classifiers"
            w, b = P[f"head.{t}.{sub}.{t}.weight"], P[f"head.{t}.{sub}.{t}.bias"]
        out[t] = F.linear(f, w, b)
    return out


# ---------------------------------------------------------------------------
# Loss
# ---------------------------------------------------------------------------
def _hard(target: torch.Tensor) -> torch.Tensor:
    return target.argmax(dim=1) if target.dim() == 2 else target.long()


def per_sample_losses(logits: dict, targets: dict, kind: str = "ce", smoothing: float = 0.1, soft_matrices: dict | None = None, ignore_index=None):
    """The criteria, R/loss/basic_loss.py:15-228 and
    R/loss/taxonomy_label_smoothing.py:219-408; ``reduction='none'`` -> [B]."""
    out = {}
    for t in sorted(logits.keys(), key=lambda k: int(k.split("_L")[-1])):  # R/loss/core_loss.py:46
        z, y = logits[t].float(), targets[t]
        lp = F.log_softmax(z, dim=-1)
        if kind == "ce":
            l = -lp.gather(1, _hard(y)[:, None])[:, 0]
        elif kind == "ls":
            C = z.shape[1]
            dist = torch.full_like(lp, smoothing / (C - 1))
            dist.scatter_(1, _hard(y)[:, None], 1.0 - smoothing)
            l = -(dist * lp).sum(1)
        elif kind == "soft":
            l = -(y * lp).sum(1)
        elif kind == "taxonomy":
            l = -(soft_matrices[t][_hard(y)] * lp).sum(1)
        else:
            raise ValueError(kind)
        if ignore_index is not None and kind != "soft":
            l = l.masked_fill(_hard(y) == ignore_index, 0.0)
        out[t] = l
    return out


def _class_sample_weight(cw: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Per-sample class weight: cw[y] for hard labels, sum_c y_c cw_c for [B, C] targets (R/loss/masking.py:469-518)."""
    if y.dim() == 1:
        return cw[y.long()]
    return (y.float() * cw[None]).sum(1)


def hierarchical_loss(
    logits: dict,
    targets: dict,
    kind: str = "ce",
    task_weights: dict | None = None,
    null_mask_prob: float = 1.0,
    phase1_mask_null: bool = False,
    is_validation: bool = False,
    smoothing: float = 0.1,
    soft_matrices: dict | None = None,
    class_weights: dict | None = None,
    coin_flips: dict | None = None,
    apply_cw: bool | None = None,
    criterion_weights: dict | None = None,
):
    """weighted_hierarchical_loss, R/loss/hierarchical_loss.py:138-395, with
    masking R/loss/masking.py:19-466,521-700 and static task weighting
    R/loss/gradient_weighting.py:301-358.

    total = sum_k w_k * (sum_i l_ik) / max(n_valid_k, 1e-6),  n_valid_k = #(l_ik != 0)
    (in the PHASE1 training branch ``num_valid_samples_per_task`` is absent, so the
    divisor is the batch size -- hierarchical_loss.py:241-276,337-340; in validation the
    criteria still carry ignore_index = 0 (loss/utils.py:104-145), so null samples are
    zero and drop out of n_valid).
    ``coin_flips[t]``: bool [B]; True keeps a null sample (the reference draws
    ``rand < null_mask_prob`` for null rows only; tests inject the draw).
    ``class_weights[t]``: [C] tensor = the reference's dict lookup.  It is applied once inside
    apply_loss_masking (masking.py:696-698; not on the PHASE1 training branch), once in
    hierarchical_loss.py:313-334 when ``apply_cw`` (config LOSS.GRAD_WEIGHTING.CLASS.TRAIN / .VAL;
    None = the config defaults True / False) and once in GradientWeighting.forward (:334-352).
    ``criterion_weights[t]``: [C] tensor = a criterion built with ``weight=`` and
    ``apply_class_weights=True`` (basic_loss.py:76-90,160-175,217-221).
    """
    ignore = 0 if phase1_mask_null else None  # R/loss/utils.py (prepare_loss_functions)
    raw = per_sample_losses(logits, targets, kind, smoothing, soft_matrices, ignore)
    keys = list(raw.keys())
    if criterion_weights:
        raw = {t: raw[t] * _class_sample_weight(criterion_weights[t], targets[t] if kind == "soft" else _hard(targets[t])) for t in keys}
    masked, n_valid = {}, {}
    for t in keys:
        y = targets[t]
        null = (y == 0) if y.dim() == 1 else (y[:, 0] > 0.5)
        l = raw[t]
        if phase1_mask_null and not is_validation:
            masked[t] = l * (~null).float()
            n_valid[t] = l.shape[0]
        else:
            p = 1.0 if is_validation else null_mask_prob
            if p < 1.0 and bool(null.any()):
                keep = coin_flips[t] if coin_flips is not None else (torch.rand(l.shape[0]) < p)
                l = torch.where(null & ~keep, torch.zeros_like(l), l)
            masked[t] = l
            n_valid[t] = int((l != 0).sum().item())
    cw_times = 0
    if class_weights:
        if apply_cw is None:
            apply_cw = not is_validation
        cw_times = (0 if (phase1_mask_null and not is_validation) else 1) + (1 if apply_cw else 0) + 1
    total = 0.0
    weighted = {}
    for t in keys:
        w = 1.0 if task_weights is None else float(task_weights[t])
        l = masked[t]
        if cw_times and t in class_weights:
            l = l * _class_sample_weight(class_weights[t], targets[t]).pow(cw_times)
        weighted[t] = l.sum() / max(float(n_valid[t]), 1e-6) * w
        total = total + weighted[t]
    comps = {
        "total": float(total.detach()),
        "tasks": {t: float(raw[t].mean().detach()) for t in keys},
        "weighted_tasks": {t: float(weighted[t].detach()) for t in keys},
        "num_valid_samples_per_task": n_valid,
    }
    return total, comps


def synthetic_taxonomy_smoothing(tasks: list, alpha: float = 0.1, beta: float = 1.0) -> dict[str, torch.Tensor]:
    """[C, C] soft-label matrices for the synthetic taxonomy of SURVEY.md 8(d)
    (child i -> parent 0 if i == 0 else 1 + (i-1) mod (C_parent-1)), built the way
    build_taxonomy_smoothing_matrix does (R/loss/taxonomy_label_smoothing.py:30-128):
    off-diagonal mass alpha spread as exp(-beta * tree distance); class 0 (null,
    a root) gets uniform mass.  Distance: siblings 2, otherwise 4 (two-level view)."""
    mats = {}
    for li, (t, C) in enumerate(tasks):
        if li + 1 < len(tasks):
            Cp = tasks[li + 1][1]
            idx = torch.arange(C)
            parent = torch.where(idx == 0, torch.zeros_like(idx), 1 + (idx - 1) % max(Cp - 1, 1))
            same = parent[:, None] == parent[None, :]
            dist = torch.where(same, torch.tensor(2.0), torch.tensor(4.0))
        else:
            dist = torch.full((C, C), 2.0)
        w = torch.exp(-beta * dist)
        w.fill_diagonal_(0.0)
        if C > 1:
            w[0] = 1.0 / (C - 1)
            w[0, 0] = 0.0
        m = w * (alpha / w.sum(1, keepdim=True).clamp_min(1e-9))
        m.fill_diagonal_(1.0 - alpha)
        m = m / m.sum(1, keepdim=True)
        mats[t] = m.float()
    return mats


# ---------------------------------------------------------------------------
# Optimizer step
# ---------------------------------------------------------------------------
def split_decay(names_shapes: dict[str, tuple]) -> tuple[list, list]:
    """set_weight_decay, R/optimizers/build.py:687-716: no decay for 1-D tensors
    and names ending in '.bias'."""
    decay, no_decay = [], []
    for n, s in names_shapes.items():
        (no_decay if (len(s) == 1 or n.endswith(".bias")) else decay).append(n)
    return decay, no_decay


def adamw_clip_step(P: dict, G: dict, state: dict, step: int, lr: float, wd: float = 0.05, betas=(0.9, 0.999), eps: float = 1e-8, clip: float = 5.0):
    """clip_grad_norm_(CLIP_GRAD) then torch.optim.AdamW semantics
    (R/train.py:282-313, R/optimizers/build.py:98-104).  In place on ``P``;
    ``state[name] = (m, v)``; ``step`` is 1-based.  Returns the pre-clip norm."""
    names = [n for n in P if n in G and G[n] is not None]
    total = torch.sqrt(sum((G[n].double() ** 2).sum() for n in names)).float()
    coef = torch.clamp(clip / (total + 1e-6), max=1.0) if clip > 0 else torch.tensor(1.0)
    decay, _ = split_decay({n: tuple(P[n].shape) for n in names})
    decay = set(decay)
    b1, b2 = betas
    for n in names:
        g = G[n] * coef
        if n not in state:
            state[n] = (torch.zeros_like(P[n]), torch.zeros_like(P[n]))
        m, v = state[n]
        if n in decay:
            P[n].mul_(1.0 - lr * wd)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1**step
        bc2 = 1 - b2**step
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        P[n].addcdiv_(m, denom, value=-lr / bc1)
    return float(total)


def unique_params(P: dict) -> dict:
    """Drop aliased entries (shared hierarchical level classifiers)."""
    seen, out = set(), {}
    for n, t in P.items():
        if id(t) in seen:
            continue
        seen.add(id(t))
        out[n] = t
    return out


def train_step(P: dict, a: Arch, x, meta, targets, state: dict, step: int, lr: float, **loss_kw):
    """One reference training step (R/train.py:115-377) at AMP O0, accumulation 1:
    forward -> loss -> backward -> clip 5.0 -> AdamW.  Returns (loss, grads, logits)."""
    U = unique_params(P)
    leaves = {n: t.detach().requires_grad_(True) for n, t in U.items()}
    full = {n: leaves[next(k for k, v in U.items() if v is t)] for n, t in P.items()}
    logits = forward(full, a, x, meta)
    total, comps = hierarchical_loss(logits, targets, **loss_kw)
    total.backward()
    G = {n: leaves[n].grad for n in leaves}
    with torch.no_grad():
        adamw_clip_step(U, G, state, step, lr)
    return float(total.detach()), G, {t: v.detach() for t, v in logits.items()}
