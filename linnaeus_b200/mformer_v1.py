"""mFormerV1 on the B200 kernels, behind the reference's model surface.

Drop-in for ``linnaeus/models/mFormerV1.py`` (R/models/mFormerV1.py:31-541):
same constructor ``(config, num_classes=, taxonomy_tree=)``, same
``forward(x, meta=None, force_checkpointing=None) -> {task: logits}``, same
``forward_features``, same parameter / buffer names and shapes (so checkpoints,
optimizer name filters and ``load_state_dict`` interchange), same config keys.

The ``nn.Conv2d`` / ``nn.Linear`` / ``nn.LayerNorm`` sub-modules below are parameter
holders only: their ``forward`` is never called.  Compute goes through
``linnaeus_b200.functional`` (C-ABI CUDA kernels).  Differences in *how*:
the conv trunk runs NHWC end to end (channels-first LayerNorm becomes a row LN, the
4x4/s4 and 2x2/s2 convolutions become im2col/space-to-depth + tensor-core GEMMs), the
metadata/CLS concat and split are single gather kernels, and the K head GEMMs run as
one concatenated GEMM whose output the fused loss reads directly.
"""
from __future__ import annotations

from typing import Any

import torch
import torch.nn as nn
import torch.nn.functional as TF

from . import functional as F
from .heads import configure_classification_heads
from .registry import register_model


def trunc_normal_(t: torch.Tensor, std: float = 0.02) -> torch.Tensor:
    return nn.init.trunc_normal_(t, std=std, a=-2.0, b=2.0)


class LayerNormChannelsFirst(nn.Module):
    """Parameter holder for the reference's channels-first LN (convnext.py:21-43)."""

    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))
        self.eps = eps


# bf16 shadow of the parameters for the forward in flight: id(param) -> bf16 view (set by mFormerV1._features)
_ACTIVE_SHADOW: dict | None = None


def _wc(p: torch.Tensor):
    """Compute-dtype copy of a parameter from the per-step shadow (None in fp32 mode -> the op casts itself)."""
    return None if _ACTIVE_SHADOW is None else _ACTIVE_SHADOW.get(id(p))


from .flat import WEIGHTS_EPOCH  # noqa: E402


class _Bf16Shadow:
    """bf16 copies of all parameters, refreshed with one cast launch per distinct storage: after FlatAdamW has
    re-homed the parameters that is two launches per step instead of one per weight."""

    def __init__(self):
        self.key = None
        self.stamp = None
        self.bufs = []   # (fp32 storage-wide view, bf16 buffer)
        self.views = {}

    def refresh(self, params, reuse_if_unchanged: bool = False) -> dict:
        """``reuse_if_unchanged`` (inference: grad mode off): skip the casts when no parameter was written since the last refresh -
        same storages, same ``_version`` of every parameter, same flat.WEIGHTS_EPOCH (kernel-side writes).  Without FlatAdamW every
        weight has its own storage, i.e. ~90 cast launches per forward otherwise."""
        key = tuple(p.data_ptr() for p in params)
        if reuse_if_unchanged:
            stamp = (key, tuple(p._version for p in params), WEIGHTS_EPOCH[0])
            if stamp == self.stamp:
                return self.views
        else:
            stamp = None
        self.stamp = stamp
        if key != self.key:
            by_storage = {}
            for p in params:
                st = p.untyped_storage()
                by_storage.setdefault(st.data_ptr(), (st, []))[1].append(p)
            self.bufs, self.views = [], {}
            for _, (st, ps) in by_storage.items():
                n = st.nbytes() // 4
                whole = torch.empty(0, dtype=torch.float32, device=ps[0].device).set_(st, 0, (n,))
                buf = torch.empty(n, dtype=torch.bfloat16, device=ps[0].device)
                self.bufs.append((whole, buf))
                for p in ps:
                    if p.is_contiguous():
                        self.views[id(p)] = buf[p.storage_offset():p.storage_offset() + p.numel()].view(p.shape)
            self.key = key
        for whole, buf in self.bufs:
            F.cast_bf16(whole, out=buf)
        return self.views


class _ParamViews:
    """Cached ``nn.Parameter`` views over one or several model parameters that sit back to back in one storage (after FlatAdamW has
    re-homed them they do): the stem conv weight as a [Cout, Cin*16] matrix, all K head weights as ONE [sum C_k, D] matrix.  The
    view's ``.grad`` is the matching view of the members' gradient buffer, so the Linear kernels accumulate weight gradients in
    place and report per member to a data-parallel wrapper; no torch.cat / pad / reshape copies forward, no CatBackward + K
    accumulate launches backward.  Returns None when the members are not contiguous (fresh model, no flat optimizer yet)."""

    def __init__(self):
        self._cache = {}

    def get(self, tag: str, members: list, shape: tuple):
        g0 = members[0].grad
        key = tuple(m.data_ptr() for m in members) + (None if g0 is None else g0.data_ptr(),)
        hit = self._cache.get(tag)
        if hit is not None and hit[0] == key:
            return hit[1]
        view = self._build(members, shape)
        self._cache[tag] = (key, view)
        return view

    @staticmethod
    def _contiguous(ts) -> bool:
        st = ts[0].untyped_storage().data_ptr()
        off = ts[0].storage_offset()
        for t in ts:
            if t is None or not t.is_contiguous() or t.dtype != torch.float32 or t.untyped_storage().data_ptr() != st or t.storage_offset() != off:
                return False
            off += t.numel()
        return True

    def _build(self, members, shape):
        n = 1
        for d in shape:
            n *= d
        if sum(m.numel() for m in members) != n or not self._contiguous([m.data for m in members]):
            return None
        grads = [m.grad for m in members]
        if any(g is None for g in grads) or not self._contiguous(grads):
            return None
        data = torch.empty(0, dtype=torch.float32, device=members[0].device).set_(members[0].data.untyped_storage(), members[0].storage_offset(), shape)
        view = nn.Parameter(data, requires_grad=True)
        view.grad = torch.empty(0, dtype=torch.float32, device=members[0].device).set_(grads[0].untyped_storage(), grads[0].storage_offset(), shape)
        mem = list(members)

        def ready(_p, mem=mem):
            for m in mem:  # forward "gradient final" to whoever listens on the real parameters (linnaeus_b200.DataParallel)
                h = getattr(m, "_lnx_grad_ready", None)
                if h is not None:
                    h(m)

        view._lnx_grad_ready = ready
        view._lnx_members = mem
        return view


def _shadow_view(members: list, shape: tuple, views: dict | None = None):
    """bf16 shadow of a _ParamViews view: the shadow buffer mirrors the parameter storage element for element."""
    views = _ACTIVE_SHADOW if views is None else views
    if views is None:
        return None
    first = views.get(id(members[0]))
    if first is None:
        return None
    return torch.empty(0, dtype=first.dtype, device=first.device).set_(first.untyped_storage(), first.storage_offset(), shape)


def drop_path_mask(B: int, drop_prob: float, training: bool, device) -> torch.Tensor | None:
    """Per-sample stochastic-depth multiplier floor(keep + U[0,1)) / keep (drop_path.py:11-36); None when inactive.
    The multiplier is applied inside the epilogue of the GEMM that closes the residual branch."""
    if drop_prob == 0.0 or not training:
        return None
    keep = 1.0 - drop_prob
    return torch.floor(keep + torch.rand(B, device=device, dtype=torch.float32)) / keep


class ConvNeXtBlock(nn.Module):
    """dw7x7 -> LN -> Linear 4x -> GELU -> Linear -> gamma -> + x  (convnext.py:46-100), NHWC."""

    def __init__(self, dim: int, drop_path: float = 0.0, layer_scale_init_value: float = 1e-6):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value > 0 else None
        self.drop_prob = float(drop_path)

    def run(self, x: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
        C = x.shape[-1]
        t, skip = F.dwconv7_fork(x.view(B, H, W, C), self.dwconv.weight, self.dwconv.bias)
        t = F.layernorm(t.view(-1, C), self.norm.weight, self.norm.bias, 1e-6)
        mask = drop_path_mask(B, self.drop_prob, self.training, x.device)
        return F.mlp2(t, self.pwconv1.weight, self.pwconv1.bias, self.pwconv2.weight, self.pwconv2.bias,
                      w1c=_wc(self.pwconv1.weight), w2c=_wc(self.pwconv2.weight),
                      act="gelu", residual=skip.view(-1, C), col_scale=self.gamma, row_scale=mask, rows_per_group=H * W)


class ConvNeXtDownsampleLayer(nn.Module):
    """LN (channels-first in the reference) -> Conv 2x2 s2 (convnext.py:104-115)."""

    def __init__(self, in_dim: int, out_dim: int):
        super().__init__()
        self.norm = LayerNormChannelsFirst(in_dim, eps=1e-6)
        self.conv = nn.Conv2d(in_dim, out_dim, kernel_size=2, stride=2)

    def run(self, x: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
        C = x.shape[-1]
        t = F.layernorm(x.reshape(-1, C), self.norm.weight, self.norm.bias, 1e-6)
        a = F.space_to_depth(t.view(B, H, W, C))
        # Conv2d weight [Cout, Cin, kh, kw] -> [Cout, (kh, kw, Cin)] to match the gather order
        w2d = self.conv.weight.permute(0, 2, 3, 1).reshape(self.conv.weight.shape[0], 4 * C)
        return F.linear(a, w2d, self.conv.bias)


class RoPE2DAttention(nn.Module):
    def __init__(self, dim: int, num_heads: int, rope_theta: float = 10000.0, qkv_bias: bool = True):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.freqs = nn.Parameter(init_random_2d_freqs(self.head_dim, num_heads, rope_theta))


def init_random_2d_freqs(head_dim: int, num_heads: int, theta: float = 10000.0) -> torch.Tensor:
    """Mixed-mode frequency init (rope_2d_mhsa.py:76-111): inv_freq_j * (cos phi_h, sin phi_h)."""
    half = head_dim // 2
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2)[:half].float() / head_dim))
    ang = torch.rand(num_heads, 1) * 2 * torch.pi
    return torch.stack([inv[None] * torch.cos(ang), inv[None] * torch.sin(ang)], 0).float()


class Mlp(nn.Module):
    def __init__(self, in_features: int, hidden_features: int | None = None, out_features: int | None = None):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)


class RoPE2DMHSABlock(nn.Module):
    """Pre-norm block (rope_2d_mhsa.py:511-645), LN eps 1e-5."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float, rope_theta: float, extra_token_num: int, drop_path: float = 0.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.attn = RoPE2DAttention(dim, num_heads, rope_theta, qkv_bias=True)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.extra_token_num = extra_token_num
        self.drop_prob = float(drop_path)

    def run(self, x: torch.Tensor, H: int, W: int) -> torch.Tensor:
        a = self.attn
        t, x = F.layernorm_fork(x, self.norm1.weight, self.norm1.bias, 1e-5)  # x: skip connection (gradient fused into LN bwd)
        B, N = x.shape[0], x.shape[1]
        if F.fused_qkv_rope_ok(t, a.qkv.weight.shape[1], a.num_heads, N):
            # projection with the cos factors in its epilogue, attention straight from its output (no q / k scaling pass, no copies)
            o = F.qkv_rope_attention(t, a.qkv.weight, a.qkv.bias, _wc(a.qkv.weight), a.freqs, H, W, a.num_heads, self.extra_token_num)
        else:
            qkv = F.linear(t, a.qkv.weight, a.qkv.bias, weight_c=_wc(a.qkv.weight))
            o = F.rope_attention(qkv, a.freqs, H, W, a.num_heads, self.extra_token_num)
        m1 = drop_path_mask(B, self.drop_prob, self.training, x.device)
        x = F.linear(o, a.proj.weight, a.proj.bias, weight_c=_wc(a.proj.weight), residual=x, row_scale=m1, rows_per_group=N)
        t, x = F.layernorm_fork(x, self.norm2.weight, self.norm2.bias, 1e-5)
        m2 = drop_path_mask(B, self.drop_prob, self.training, x.device)
        return F.mlp2(t, self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight, self.mlp.fc2.bias,
                      w1c=_wc(self.mlp.fc1.weight), w2c=_wc(self.mlp.fc2.weight), act="gelu", residual=x,
                      row_scale=m2, rows_per_group=N)


class ResNormLayer(nn.Module):
    """x + LN(ReLU(W2 LN(ReLU(W1 x))))  (normalization/res_norm_layer.py:7-30)."""

    def __init__(self, dim: int):
        super().__init__()
        self.norm_fn1 = nn.LayerNorm(dim)
        self.norm_fn2 = nn.LayerNorm(dim)
        self.w1 = nn.Linear(dim, dim)
        self.w2 = nn.Linear(dim, dim)


def _meta_head(in_dim: int, dim: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(in_dim, dim), nn.ReLU(inplace=True), nn.LayerNorm(dim), ResNormLayer(dim))


def _run_meta_head(seq: nn.Sequential, meta: torch.Tensor, off: int, dim: int, cdtype: torch.dtype) -> torch.Tensor:
    lin, _, ln, rn = seq[0], seq[1], seq[2], seq[3]
    m = meta[:, off:off + dim]
    wc = _wc(lin.weight)
    if wc is None:
        wc = F.compute_copy(lin.weight, cdtype)  # fixes the output dtype of the mixed f32-in GEMM
    t = F.linear(m, lin.weight, lin.bias, weight_c=wc, act="relu", x_ld=meta.shape[1])
    t = F.layernorm(t, ln.weight, ln.bias, 1e-5)
    r = F.linear(t, rn.w1.weight, rn.w1.bias, weight_c=_wc(rn.w1.weight), act="relu")
    r = F.layernorm(r, rn.norm_fn1.weight, rn.norm_fn1.bias, 1e-5)
    r = F.linear(r, rn.w2.weight, rn.w2.bias, weight_c=_wc(rn.w2.weight), act="relu")
    return F.layernorm(r, rn.norm_fn2.weight, rn.norm_fn2.bias, 1e-5, residual=t)


class LogitsDict(dict):
    """``{task: logits}`` plus the concatenated [B, sum C_k] tensor the fused loss consumes."""

    cat: torch.Tensor | None = None
    class_off: tuple | None = None


@register_model("mFormerV1")
class mFormerV1(nn.Module):
    def __init__(self, config, **kwargs):
        super().__init__()
        self.config = config
        M = config.MODEL
        # BaseModel._init_common_parameters (base_model.py:60-83)
        self.drop_rate = M.DROP_RATE
        self.drop_path_rate = M.DROP_PATH_RATE
        self.attn_drop_rate = M.get("ATTN_DROP_RATE", 0.0)
        self.label_smoothing = M.get("LABEL_SMOOTHING", 0.0)
        self.only_last_cls = M.ONLY_LAST_CLS
        img_size = M.get("IMG_SIZE", 224) if hasattr(M, "get") else M.IMG_SIZE
        self.img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        in_chans = M.IN_CHANS

        if not hasattr(M, "CONVNEXT_STAGES"):
            raise ValueError("mFormerV1 requires MODEL.CONVNEXT_STAGES config")
        cs = M.CONVNEXT_STAGES
        depths, dims = list(cs.DEPTHS), list(cs.DIMS)
        self.convnext_ls_init = cs.get("LAYER_SCALE_INIT_VALUE", 1e-6)
        if len(depths) != 4 or len(dims) != 4:
            raise ValueError("CONVNEXT_STAGES depths and dims must be lists of length 4.")
        if not hasattr(M, "ROPE_STAGES"):
            raise ValueError("mFormerV1 requires MODEL.ROPE_STAGES config")
        rs = M.ROPE_STAGES
        rdepths, rdims, rheads, rratio = list(rs.DEPTHS), list(rs.DIMS), list(rs.NUM_HEADS), list(rs.MLP_RATIO)
        self.rope_theta = rs.get("ROPE_THETA", 10000.0)
        self.rope_mixed = rs.get("ROPE_MIXED", True)
        if len(rdepths) != 2 or len(rdims) != 2 or len(rheads) != 2 or len(rratio) != 2:
            raise ValueError("ROPE_STAGES depths, dims, num_heads, mlp_ratio must be lists of length 2.")
        if not self.rope_mixed:
            raise ValueError("linnaeus_b200 implements ROPE_MIXED=True only (the reference's axial mode is broken, SURVEY F3)")
        self.use_flash_attn = M.get("USE_FLASH_ATTN", False)  # accepted; the standard-path semantics are always used

        # metadata components (mFormerV1.py:94-130)
        self.use_meta = False
        self.meta_components: dict[str, dict] = {}
        self.meta_dims: list[int] = []
        D = config.DATA
        if hasattr(D, "META") and D.META.get("ACTIVE", False):
            if hasattr(D.META, "COMPONENTS"):
                self.use_meta = True
                items = []
                for name, comp in D.META.COMPONENTS.items():
                    if comp.get("ENABLED", False) and comp.get("IDX", -1) >= 0:
                        items.append((comp.get("IDX"), name, comp))
                items.sort(key=lambda t: t[0])
                off = 0
                for _, name, comp in items:
                    self.meta_dims.append(comp.DIM)
                    self.meta_components[name] = {"dim": comp.DIM, "offset": off}
                    off += comp.DIM
            elif M.get("META_DIMS"):
                raise ValueError("legacy MODEL.META_DIMS is not supported; use DATA.META.COMPONENTS")
        self.extra_token_num = 1 + len(self.meta_dims)

        total_depth = sum(depths[:2]) + sum(rdepths)
        dpr = [x.item() for x in torch.linspace(0, self.drop_path_rate, total_depth)]

        self.stem = nn.Sequential(nn.Conv2d(in_chans, dims[0], kernel_size=4, stride=4), LayerNormChannelsFirst(dims[0], eps=1e-6))
        self.downsample_layers = nn.ModuleList([ConvNeXtDownsampleLayer(dims[i], dims[i + 1]) for i in range(3)])
        if rdims[0] != dims[2]:
            raise ValueError(f"ConvNeXt dim[2] ({dims[2]}) must match RoPE dim[0] ({rdims[0]})")
        if rdims[1] != dims[3]:
            raise ValueError(f"ConvNeXt dim[3] ({dims[3]}) must match RoPE dim[1] ({rdims[1]})")
        self.stages = nn.ModuleList()
        k = 0
        for st in range(2):
            self.stages.append(nn.ModuleList([ConvNeXtBlock(dims[st], dpr[k + i], self.convnext_ls_init) for i in range(depths[st])]))
            k += depths[st]
        for st in range(2):
            self.stages.append(nn.ModuleList([
                RoPE2DMHSABlock(rdims[st], rheads[st], rratio[st], self.rope_theta, self.extra_token_num, dpr[k + i])
                for i in range(rdepths[st])
            ]))
            k += rdepths[st]
        self.norm_1 = nn.LayerNorm(rdims[0])
        self.norm_2 = nn.LayerNorm(rdims[1])
        self.cls_token_1 = nn.Parameter(torch.zeros(1, 1, rdims[0]))
        self.cls_token_2 = nn.Parameter(torch.zeros(1, 1, rdims[1]))
        trunc_normal_(self.cls_token_1, std=0.02)
        trunc_normal_(self.cls_token_2, std=0.02)
        for name, info in self.meta_components.items():
            if info["dim"] > 0:
                setattr(self, f"meta_{name.lower()}_head_1", _meta_head(info["dim"], rdims[0]))
                setattr(self, f"meta_{name.lower()}_head_2", _meta_head(info["dim"], rdims[1]))
            else:
                raise ValueError("metadata components with DIM 0 are not supported")
        if not self.only_last_cls:
            self.cl_1_fc = nn.Sequential(Mlp(rdims[0], rdims[0], rdims[1]), nn.LayerNorm(rdims[1]))
            self.aggregate = nn.Conv1d(in_channels=2, out_channels=1, kernel_size=1)
        else:
            self.cl_1_fc = None
            self.aggregate = None
        self.final_norm = nn.LayerNorm(rdims[1])

        self.head = configure_classification_heads(
            heads_config=M.CLASSIFICATION.HEADS,
            in_features=rdims[1],
            num_classes_dict=kwargs.get("num_classes"),
            task_keys=list(D.TASK_KEYS_H5),
            taxonomy_tree=kwargs.get("taxonomy_tree"),
        )
        self.apply(self._init_weights)
        self.dims, self.in_chans = dims, in_chans
        self._compute_dtype: torch.dtype | None = None  # None: follow autocast (on -> bf16, off -> fp32)
        self._shadow: _Bf16Shadow | None = None
        self._side_stream = None  # metadata-token branch (see _extras_async)
        self._views = _ParamViews()

    # -- init / metadata properties (mFormerV1.py:351-405) ---------------------
    def _init_weights(self, m):
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            trunc_normal_(m.weight, std=0.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @property
    def parameter_groups_metadata(self) -> dict[str, Any]:
        return {
            "stages": {
                "convnext_stages": ["stem.", "stages.0.", "stages.1.", "downsample_layers.0", "downsample_layers.1"],
                "rope_stages": ["stages.2.", "stages.3.", "downsample_layers.2", "downsample_layers.3"],
                "rope_freqs": ["freqs"],
            },
            "heads": {"classification_heads": ["head."], "meta_heads": ["meta_"]},
            "embeddings": ["cls_token"],
            "norm_layers": ["norm", ".bn", "LayerNorm"],
            "aggregation": ["cl_1_fc.", "aggregate.", "final_norm."],
        }

    @property
    def pretrained_ckpt_handling_metadata(self) -> dict[str, Any]:
        return {
            "drop_buffers": [],
            "drop_params": ["head.", "meta_", "pos_embed", "norm.", "downsample_layers."],
            "interpolate_rel_pos_bias": False,
            "supports_module_prefix": True,
            "strict": False,
        }

    # -- compute dtype -----------------------------------------------------------
    def set_compute_dtype(self, dtype: torch.dtype | str | None) -> "mFormerV1":
        """``torch.bfloat16`` / ``torch.float32`` / None (= bf16 under autocast, else fp32)."""
        if isinstance(dtype, str):
            dtype = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "fp32": torch.float32, "float32": torch.float32}[dtype]
        if dtype not in (None, torch.bfloat16, torch.float32):
            raise ValueError("compute dtype must be bfloat16 or float32")
        self._compute_dtype = dtype
        return self

    def _cdtype(self) -> torch.dtype:
        if self._compute_dtype is not None:
            return self._compute_dtype
        return torch.bfloat16 if torch.is_autocast_enabled() else torch.float32

    # -- forward -----------------------------------------------------------------
    def _extras(self, stage: int, meta: torch.Tensor | None, cdtype: torch.dtype):
        if not (self.use_meta and meta is not None and self.meta_components):
            return None
        meta = meta.float().contiguous()
        toks = []
        for name, info in self.meta_components.items():
            seq = getattr(self, f"meta_{name.lower()}_head_{stage}")
            toks.append(_run_meta_head(seq, meta, info["offset"], info["dim"], cdtype))
        return torch.stack(toks, dim=1)

    def forward_features(self, x: torch.Tensor, meta: torch.Tensor | None = None, force_checkpointing: bool | None = None) -> torch.Tensor:
        """R/models/mFormerV1.py:407-529.  ``force_checkpointing`` is accepted for API
        compatibility; activations are kept (B200 has 180 GB; recompute is never needed
        at the reference's batch sizes)."""
        if not x.is_cuda:
            raise RuntimeError("linnaeus_b200.mFormerV1 runs on CUDA (sm_100a) only; there is no CPU fallback")
        global _ACTIVE_SHADOW
        cd = self._cdtype()  # read before autocast is switched off for the kernels' torch plumbing
        prev = _ACTIVE_SHADOW
        try:
            with torch.autocast("cuda", enabled=False):
                if cd == torch.bfloat16:
                    if self._shadow is None:
                        self._shadow = _Bf16Shadow()
                        self._shadow_params = [p for p in self.parameters() if p.ndim >= 2]
                    _ACTIVE_SHADOW = self._shadow.refresh(self._shadow_params, reuse_if_unchanged=not torch.is_grad_enabled())
                else:
                    _ACTIVE_SHADOW = None
                return self._features(x, meta, cd)
        finally:
            _ACTIVE_SHADOW = prev

    def _extras_async(self, meta, cd):
        """Both stages' metadata tokens on a side stream: ~40 tiny launches (3 components x 2 stages of
        Linear -> ReLU -> LN -> ResNorm on [B, D]) that depend only on ``meta``, so they overlap the convolutional
        trunk instead of sitting between its kernels.  autograd runs each backward node on its forward stream, so the
        backward of this branch overlaps too; under CUDA-graph capture the fork / join become graph edges."""
        if not (self.use_meta and meta is not None and self.meta_components):
            return None, None, None
        cur = torch.cuda.current_stream()
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=meta.device)
        side = self._side_stream
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            ex1 = self._extras(1, meta, cd)
            ex2 = self._extras(2, meta, cd)
        return ex1, ex2, side

    def _features(self, x, meta, cd):
        B, Cin, Hi, Wi = x.shape
        dims = self.dims
        ex1, ex2, side = self._extras_async(meta, cd)
        # stem: 4x4/s4 conv as im2col (K = 48 padded to 64) + GEMM, then LN (NHWC rows)
        kpad = ((Cin * 16 + 7) // 8) * 8  # K = 48: the GEMM's TMA box zero-fills the tail of its 64-wide k-block
        a = F.patchify(x, 4, kpad, cd)
        conv = self.stem[0]
        w2d = self._views.get("stem", [conv.weight], (dims[0], Cin * 16)) if kpad == Cin * 16 and torch.is_grad_enabled() else None
        if w2d is not None:
            y = F.linear(a, w2d, conv.bias, weight_c=_shadow_view([conv.weight], (dims[0], Cin * 16)))
        else:
            w2d = conv.weight.reshape(dims[0], Cin * 16)
            if kpad != Cin * 16:
                w2d = TF.pad(w2d, (0, kpad - Cin * 16))
            y = F.linear(a, w2d, conv.bias)
        y = F.layernorm(y, self.stem[1].weight, self.stem[1].bias, 1e-6)
        H, W = Hi // 4, Wi // 4
        for blk in self.stages[0]:
            y = blk.run(y, B, H, W)
        y = self.downsample_layers[0].run(y, B, H, W)
        H, W = H // 2, W // 2
        for blk in self.stages[1]:
            y = blk.run(y, B, H, W)
        y = self.downsample_layers[1].run(y, B, H, W)
        H, W = H // 2, W // 2

        n_meta = self.extra_token_num - 1
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
            for t in (ex1, ex2):
                t.record_stream(torch.cuda.current_stream())
        if (ex1 is None) != (n_meta == 0):
            raise AssertionError(f"Input sequence length {H * W + 1 + (0 if ex1 is None else ex1.shape[1])} != H*W+extra {H * W + self.extra_token_num}")
        x3 = F.tokens_assemble(self.cls_token_1, ex1, y.view(B, H * W, dims[2]))
        for blk in self.stages[2]:
            x3 = blk.run(x3, H, W)
        x3 = F.layernorm(x3, self.norm_1.weight, self.norm_1.bias, 1e-5)
        cls1, patches = F.tokens_split(x3, n_meta)
        if not self.only_last_cls:
            mlp, ln = self.cl_1_fc[0], self.cl_1_fc[1]
            c1 = F.mlp2(cls1, mlp.fc1.weight, mlp.fc1.bias, mlp.fc2.weight, mlp.fc2.bias, w1c=_wc(mlp.fc1.weight), w2c=_wc(mlp.fc2.weight),
                        act="gelu")
            c1 = F.layernorm(c1, ln.weight, ln.bias, 1e-5)
        y = self.downsample_layers[2].run(patches.view(-1, dims[2]), B, H, W)
        H, W = H // 2, W // 2
        x4 = F.tokens_assemble(self.cls_token_2, ex2, y.view(B, H * W, dims[3]))
        for blk in self.stages[3]:
            x4 = blk.run(x4, H, W)
        # norm_2 is applied to every token in the reference but only the CLS row is used
        cls2, _ = F.tokens_split(x4, n_meta)
        c2 = F.layernorm(cls2, self.norm_2.weight, self.norm_2.bias, 1e-5)
        if not self.only_last_cls:
            agg = F.aggregate2(c1, c2, self.aggregate.weight, self.aggregate.bias)
        else:
            agg = c2
        return F.layernorm(agg, self.final_norm.weight, self.final_norm.bias, 1e-5)

    def forward(self, x: torch.Tensor, meta: torch.Tensor | None = None, force_checkpointing: bool | None = None):
        """-> {task: logits [B, C_k]} in ``head`` insertion order (mFormerV1.py:531-541)."""
        feats = self.forward_features(x, meta, force_checkpointing=force_checkpointing)
        with torch.autocast("cuda", enabled=False):
            ws, bs, offs = [], [], [0]
            for t, head in self.head.items():
                w, b = head.classifier_params()
                ws.append(w)
                bs.append(b)
                offs.append(offs[-1] + w.shape[0])
            # pad the class dimension to a multiple of 8 so the tensor-core epilogue stays vectorised
            pad = (-offs[-1]) % 8
            wcat = bcat = wcat_c = None
            if not pad and all(b is not None for b in bs) and torch.is_grad_enabled():
                # the K classifiers as one matrix without copying (contiguous once the flat optimizer owns the parameters)
                wcat = self._views.get("head_w", ws, (offs[-1], ws[0].shape[1]))
                bcat = self._views.get("head_b", bs, (offs[-1],)) if wcat is not None else None
                if bcat is None:
                    wcat = None
                else:
                    wcat_c = (_shadow_view(ws, (offs[-1], ws[0].shape[1]), self._shadow.views)
                              if feats.dtype == torch.bfloat16 and self._shadow is not None else None)
            if wcat is None:
                wcat = torch.cat(ws, 0)
                bcat = torch.cat([b if b is not None else torch.zeros(w.shape[0], device=w.device) for w, b in zip(ws, bs)], 0)
                if pad:
                    wcat = TF.pad(wcat, (0, 0, 0, pad))
                    bcat = TF.pad(bcat, (0, pad))
            cat = F.linear(feats, wcat, bcat, weight_c=wcat_c, out_dtype=torch.float32)
            out = LogitsDict()
            for i, t in enumerate(self.head.keys()):
                out[t] = cat[:, offs[i]:offs[i + 1]]
            if not pad:
                out.cat = cat
                out.class_off = tuple(offs)
            return out
