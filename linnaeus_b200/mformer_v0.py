"""mFormerV0 (MBConv + RelativeAttention) on the B200 kernels -- inference path.

Drop-in for ``linnaeus/models/mFormerV0.py`` (R/models/mFormerV0.py:65-660): same constructor
``(config, num_classes=, taxonomy_tree=)``, same ``forward(x, meta=None, force_checkpointing=None) -> {task: logits}``
and ``forward_features``, same parameter / buffer names and shapes (``state_dict`` interchanges with the reference,
checked key for key in tests/test_oracle_v0_vs_reference.py + tests/test_gpu_v0.py).

The ``nn.Conv2d`` / ``nn.BatchNorm2d`` / ``nn.Linear`` / ``nn.LayerNorm`` sub-modules are parameter holders: their
``forward`` is never called.  This round builds the EVAL forward only (what config 5 measures): BatchNorm is folded
into the preceding convolution once per weight version (its running statistics are constants at inference), dropout /
drop-connect / DropPath are identities.  ``model.train()`` + forward raises: the V0 training path (batch-stat BN,
backward of MBConv / squeeze-excite / relative attention) is not built yet.

How it runs (NHWC end to end):
  dense 3x3 convs (stem, overlap patch embed)   lnx_im2col3x3 gather + tensor-core GEMM (bias, ReLU fused)
  MBConv                                        1x1 expand GEMM (+folded BN, swish epilogue) -> lnx_dwconv3_fwd (static
                                                "same" padding, folded BN, swish, squeeze-excite pool sums in the same
                                                pass) -> two tiny GEMMs for the gate -> lnx_se_scale -> 1x1 project
                                                GEMM (+folded BN, +residual epilogue)
  RelativeMHSABlock                             LN -> qkv GEMM -> lnx_attn_bias_fwd (reads q/k/v straight from the qkv
                                                output, bias = table[index] gathered once per weight version) -> proj
                                                GEMM (+residual) -> LN -> MLP GEMMs (GELU epilogue, +residual)
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as TF

from . import functional as F
from ._lib import call, dt
from .heads import configure_classification_heads
from .mformer_v1 import LogitsDict, Mlp, _meta_head, _run_meta_head, trunc_normal_
from .registry import register_model


def _same_pad(img: int, k: int, s: int) -> tuple[int, int]:
    """Conv2dStaticSamePadding (mb_conv.py:46-84): TF 'SAME' padding from the FULL image size -> (before, after)."""
    o = math.ceil(img / s)
    pad = max((o - 1) * s + (k - 1) + 1 - img, 0)
    return pad // 2, pad - pad // 2


class MBConvBlock(nn.Module):
    """Parameter holder with the reference's attribute names (mb_conv.py:131-224)."""

    def __init__(self, input_filters: int, output_filters: int, image_size: int, stride: int):
        super().__init__()
        oup = input_filters * 4
        self._input_filters, self._output_filters, self._stride, self._image_size = input_filters, output_filters, stride, image_size
        self._expand_conv = nn.Conv2d(input_filters, oup, 1, bias=False)
        self._bn0 = nn.BatchNorm2d(oup, momentum=0.1, eps=0.01)
        self._depthwise_conv = nn.Conv2d(oup, oup, 3, stride=stride, groups=oup, bias=False)
        self._bn1 = nn.BatchNorm2d(oup, momentum=0.1, eps=0.01)
        sq = max(1, int(input_filters * 0.25))
        self._se_reduce = nn.Conv2d(oup, sq, 1)
        self._se_expand = nn.Conv2d(sq, oup, 1)
        self._project_conv = nn.Conv2d(oup, output_filters, 1, bias=False)
        self._bn2 = nn.BatchNorm2d(output_filters, momentum=0.1, eps=0.01)


class OverlapPatchEmbed(nn.Module):
    def __init__(self, in_chans: int, embed_dim: int):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, 3, stride=2, padding=1)
        self.norm = nn.LayerNorm(embed_dim)


class RelativeAttention(nn.Module):
    def __init__(self, dim: int, img_size: tuple[int, int], extra_token_num: int, num_heads: int):
        super().__init__()
        h, w = img_size
        self.num_heads, self.extra_token_num, self.img_size = num_heads, extra_token_num, img_size
        self.scale = (dim // num_heads) ** -0.5
        n_rel = (2 * h - 1) * (2 * w - 1) + 1
        self.relative_position_bias_table = nn.Parameter(torch.zeros(n_rel, num_heads))
        coords = torch.stack(torch.meshgrid([torch.arange(h), torch.arange(w)], indexing="ij")).reshape(2, -1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += h - 1
        rel[:, :, 1] += w - 1
        rel[:, :, 0] *= 2 * w - 1
        idx = TF.pad(rel.sum(-1), (extra_token_num, 0, extra_token_num, 0), value=n_rel - 1)
        self.register_buffer("relative_position_index", idx.long())
        self.qkv = nn.Linear(dim, 3 * dim, bias=False)
        self.proj = nn.Linear(dim, dim)
        trunc_normal_(self.relative_position_bias_table, std=0.02)


class RelativeMHSABlock(nn.Module):
    def __init__(self, input_dim: int, output_dim: int, image_size: tuple[int, int], stride: int, num_heads: int, mlp_ratio: float,
                 extra_token_num: int):
        super().__init__()
        self.stride = stride
        if stride == 2:
            self.patch_embed = OverlapPatchEmbed(input_dim, output_dim)
            self.dim = output_dim
        else:
            self.patch_embed = None
            self.dim = input_dim
        self.norm1 = nn.LayerNorm(self.dim)
        self.norm2 = nn.LayerNorm(self.dim)
        self.attn = RelativeAttention(self.dim, image_size, extra_token_num, num_heads)
        self.mlp = Mlp(self.dim, int(self.dim * mlp_ratio), self.dim)


@register_model("mFormerV0")
class mFormerV0(nn.Module):
    def __init__(self, config, **kwargs):
        super().__init__()
        self.config = config
        M = config.MODEL
        img = M.IMG_SIZE
        self.img_size = (img, img) if isinstance(img, int) else tuple(img)
        self.in_chans = M.IN_CHANS
        self.only_last_cls = M.ONLY_LAST_CLS
        self.drop_path_rate = M.DROP_PATH_RATE
        self.drop_rate = M.get("DROP_RATE", 0.0)
        if not hasattr(M, "CONV_STAGES") or not hasattr(M, "ATTENTION_STAGES"):
            raise ValueError("mFormerV0 requires MODEL.CONV_STAGES and MODEL.ATTENTION_STAGES config")
        cs, at = M.CONV_STAGES, M.ATTENTION_STAGES
        self.conv_embed_dims, self.conv_out_channels = list(cs.EMBED_DIMS), list(cs.OUT_CHANNELS)
        self.conv_depths, self.conv_stride_seqs = list(cs.DEPTHS), [list(s) for s in cs.STRIDE_SEQS]
        self.attn_embed_dims, self.attn_depths = list(at.EMBED_DIMS), list(at.DEPTHS)
        self.attn_stride_seqs = [list(s) for s in at.STRIDE_SEQS]
        self.num_heads_list, self.mlp_ratio_list = list(at.NUM_HEADS), list(at.MLP_RATIO)
        for i in range(2):
            if len(self.conv_stride_seqs[i]) != self.conv_depths[i] or len(self.attn_stride_seqs[i]) != self.attn_depths[i]:
                raise ValueError("STRIDE_SEQS must have one stride per block")
            if self.attn_stride_seqs[i][0] != 2 or any(s != 1 for s in self.attn_stride_seqs[i][1:]):
                raise ValueError("attention stages are built as [2, 1, 1, ...] (first block embeds patches)")

        # metadata components (mFormerV0.py:113-147)
        D = config.DATA
        self.use_meta = bool(D.META.get("ACTIVE", False)) if hasattr(D, "META") else False
        self.meta_components: dict[str, dict] = {}
        self.meta_dims: list[int] = []
        if hasattr(D, "META") and hasattr(D.META, "COMPONENTS"):
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
        self.extra_token_num = 1 + len(self.meta_dims)

        e0 = self.conv_embed_dims[0]
        stem = (3 * (e0 // 4), e0)
        self.stage_0 = nn.Sequential(
            nn.Conv2d(self.in_chans, stem[0], 3, 2, 1, bias=False), nn.BatchNorm2d(stem[0]), nn.ReLU(inplace=True),
            nn.Conv2d(stem[0], stem[1], 3, 1, 1, bias=False), nn.BatchNorm2d(stem[1]), nn.ReLU(inplace=True),
            nn.Conv2d(stem[1], e0, 3, 1, 1, bias=False))
        self.bn1 = nn.BatchNorm2d(e0)
        cin = e0
        for si in range(2):
            out = self.conv_out_channels[si]
            blocks = nn.ModuleList([MBConvBlock(cin if i == 0 else out, out, self.img_size[0], self.conv_stride_seqs[si][i])
                                    for i in range(self.conv_depths[si])])
            setattr(self, f"stage_{si + 1}", blocks)
            cin = out

        h = self.img_size[0] // 4
        for seq in self.conv_stride_seqs:
            for s in seq:
                h //= s
        hw3 = max(h, 1)
        for s in self.attn_stride_seqs[0]:
            hw3 //= s
        hw3 = max(hw3, 1)
        hw4 = hw3
        for s in self.attn_stride_seqs[1]:
            hw4 //= s
        hw4 = max(hw4, 1)
        self._hw = (hw3, hw4)
        D3, D4 = self.attn_embed_dims
        self.cls_token_1 = nn.Parameter(torch.zeros(1, 1, D3))
        self.cls_token_2 = nn.Parameter(torch.zeros(1, 1, D4))
        trunc_normal_(self.cls_token_1, std=0.02)
        trunc_normal_(self.cls_token_2, std=0.02)
        for name, info in self.meta_components.items():
            if info["dim"] <= 0:
                raise ValueError("metadata components with DIM 0 are not supported")
            setattr(self, f"meta_{name.lower()}_head_1", _meta_head(info["dim"], D3))
        self.stage_3 = nn.ModuleList([
            RelativeMHSABlock(self.conv_out_channels[-1] if i == 0 else D3, D3, (hw3, hw3), self.attn_stride_seqs[0][i], self.num_heads_list[0],
                              self.mlp_ratio_list[0], self.extra_token_num) for i in range(self.attn_depths[0])])
        self.norm_1 = nn.LayerNorm(D3)
        # registration order = the reference's (R/models/mFormerV0.py:225-320): stage-3 heads, stage_3, norm_1, stage-4 heads,
        # stage_4 ... so state_dict() and named_parameters() enumerate identically (index-based optimizer state interchange)
        for name, info in self.meta_components.items():
            setattr(self, f"meta_{name.lower()}_head_2", _meta_head(info["dim"], D4))
        self.stage_4 = nn.ModuleList([
            RelativeMHSABlock(D3 if i == 0 else D4, D4, (hw4, hw4), self.attn_stride_seqs[1][i], self.num_heads_list[1], self.mlp_ratio_list[1],
                              self.extra_token_num) for i in range(self.attn_depths[1])])
        self.norm_2 = nn.LayerNorm(D4)
        if not self.only_last_cls:
            self.cl_1_fc = nn.Sequential(Mlp(D3, D3, D4), nn.LayerNorm(D4))
            self.aggregate = nn.Conv1d(in_channels=2, out_channels=1, kernel_size=1)
        else:
            self.cl_1_fc = None
            self.aggregate = None
        self.norm = nn.LayerNorm(D4)
        self.head = configure_classification_heads(
            heads_config=M.CLASSIFICATION.HEADS, in_features=D4, num_classes_dict=kwargs.get("num_classes"),
            task_keys=list(D.TASK_KEYS_H5), taxonomy_tree=kwargs.get("taxonomy_tree"))
        self.apply(self._init_weights)
        self._compute_dtype: torch.dtype | None = None
        self._fold_key = None
        self._fold: dict = {}

    # -- init / metadata properties (mFormerV0.py:382-405, 482-497) -------------
    def _init_weights(self, m: nn.Module):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)
        elif isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    @property
    def parameter_groups_metadata(self) -> dict:
        return {"stages": {"conv_stages": ["stage_0", "stage_1", "stage_2"], "transformer_stages": ["stage_3", "stage_4"]},
                "heads": {"classification_heads": ["head.taxa_L"], "meta_heads": ["meta_"]},
                "embeddings": ["cls_token"], "norm_layers": ["norm", "bn"]}

    @property
    def pretrained_ckpt_handling_metadata(self) -> dict:
        return {"drop_buffers": ["relative_position_index"], "drop_params": ["head", "meta_"], "interpolate_rel_pos_bias": True,
                "supports_module_prefix": True}

    def set_compute_dtype(self, dtype):
        if isinstance(dtype, str):
            dtype = {"fp32": torch.float32, "float32": torch.float32, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16}[dtype]
        self._compute_dtype = dtype
        return self

    def _cdtype(self) -> torch.dtype:
        if self._compute_dtype is not None:
            return self._compute_dtype
        return torch.bfloat16 if torch.is_autocast_enabled() else torch.float32

    # -- weight preparation (BatchNorm folding), cached per weight version -------
    @staticmethod
    def _bn_affine(bn: nn.BatchNorm2d):
        s = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
        return s, bn.bias.detach() - bn.running_mean * s

    def _prepare(self, cd: torch.dtype) -> dict:
        key = (cd, tuple(p._version for p in self.parameters()), tuple(b._version for b in self.buffers()),
               next(self.parameters()).device)
        if key == self._fold_key:
            return self._fold
        f: dict = {}

        def conv3(name, conv, bn, cin_pad_to=None):
            w = conv.weight.detach()
            b = conv.bias.detach() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
            if bn is not None:
                s, t = self._bn_affine(bn)
                w = w * s[:, None, None, None]
                b = b * s + t
            w2d = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)  # columns (kh, kw, cin) = lnx_im2col3x3 order
            kp = ((w2d.shape[1] + 7) // 8) * 8 if cin_pad_to is None else cin_pad_to  # GEMM K pitch: multiple of 8
            if kp != w2d.shape[1]:
                w2d = TF.pad(w2d, (0, kp - w2d.shape[1]))
            f[name] = (w2d.contiguous().to(cd), b.float().contiguous())
            if cd == torch.bfloat16 and conv.stride[0] == 1 and w.shape[1] <= 64 and w.shape[1] % 8 == 0 and w.shape[0] % 16 == 0 and w.shape[0] <= 96:
                # implicit-GEMM layout (lnx_conv3x3_s1): per output channel 9 taps x 64 input channels, zero padded
                w9 = TF.pad(w.permute(0, 2, 3, 1), (0, 64 - w.shape[1])).reshape(w.shape[0], 9 * 64)
                f[name + ".w9"] = w9.contiguous().to(cd)

        def conv1(name, conv, bn):
            w = conv.weight.detach().flatten(1)
            b = conv.bias.detach() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
            if bn is not None:
                s, t = self._bn_affine(bn)
                w = w * s[:, None]
                b = b * s + t
            f[name] = (w.contiguous().to(cd), b.float().contiguous())

        k0 = ((self.in_chans * 9 + 7) // 8) * 8
        conv3("stem0", self.stage_0[0], self.stage_0[1], k0)
        conv3("stem1", self.stage_0[3], self.stage_0[4])
        conv3("stem2", self.stage_0[6], self.bn1)
        for si in (1, 2):
            for i, blk in enumerate(getattr(self, f"stage_{si}")):
                p = f"s{si}.{i}."
                conv1(p + "expand", blk._expand_conv, blk._bn0)
                s, t = self._bn_affine(blk._bn1)
                f[p + "dw"] = (blk._depthwise_conv.weight.detach().reshape(-1, 9).t().contiguous().float(), s.float().contiguous(), t.float().contiguous())
                conv1(p + "se_reduce", blk._se_reduce, None)
                conv1(p + "se_expand", blk._se_expand, None)
                conv1(p + "project", blk._project_conv, blk._bn2)
        for si, stage in ((3, self.stage_3), (4, self.stage_4)):
            for i, blk in enumerate(stage):
                p = f"s{si}.{i}."
                if blk.patch_embed is not None:
                    conv3(p + "embed", blk.patch_embed.proj, None)
                a = blk.attn
                N = a.relative_position_index.shape[0]
                bias = a.relative_position_bias_table.detach()[a.relative_position_index.view(-1)].view(N, N, a.num_heads)
                f[p + "bias"] = bias.permute(2, 0, 1).contiguous().float()
                for nm, lin in (("qkv", a.qkv), ("proj", a.proj), ("fc1", blk.mlp.fc1), ("fc2", blk.mlp.fc2)):
                    f[p + nm] = lin.weight.detach().to(cd).contiguous()
        self._fold_key, self._fold = key, f
        return f

    # -- forward -----------------------------------------------------------------
    def _conv3x3(self, x, wb, B, H, W, C, stride, act, cd, image=False, w9=None):
        w2d, b = wb
        Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
        if w9 is not None and stride == 1 and not image and act in (None, "relu"):
            # tensor-core implicit GEMM straight from the NHWC activation (no im2col buffer)
            y = torch.empty((B * H * W, w9.shape[0]), dtype=cd, device=x.device)
            call("lnx_conv3x3_s1", x.data_ptr(), w9.data_ptr(), b.data_ptr(), y.data_ptr(), B, H, W, C, w9.shape[0], int(act == "relu"), dt(y))
            return y, H, W
        a = torch.empty((B * Ho * Wo, w2d.shape[1]), dtype=cd, device=x.device)
        call("lnx_im2col3x3", x.data_ptr(), int(image), a.data_ptr(), B, H, W, C, stride, Ho, Wo, w2d.shape[1], dt(a))
        y = F.linear(a, w2d, b, weight_c=w2d, act=act)
        return y, Ho, Wo

    def _mbconv(self, blk: MBConvBlock, f: dict, p: str, x, B, H, W, cd):
        cin, cout, stride = blk._input_filters, blk._output_filters, blk._stride
        oup = cin * 4
        w, b = f[p + "expand"]
        t = F.linear(x, w, b, weight_c=w, act="swish")
        lo, _ = _same_pad(blk._image_size, 3, stride)
        hi = _same_pad(blk._image_size, 3, stride)[1]
        Ho, Wo = (H + lo + hi - 3) // stride + 1, (W + lo + hi - 3) // stride + 1
        w9, s1, t1 = f[p + "dw"]
        y = torch.empty((B * Ho * Wo, oup), dtype=cd, device=x.device)
        pool = torch.zeros((B, oup), dtype=torch.float32, device=x.device)
        call("lnx_dwconv3_fwd", t.data_ptr(), w9.data_ptr(), s1.data_ptr(), t1.data_ptr(), y.data_ptr(), pool.data_ptr(), B, H, W, oup, stride,
             lo, lo, Ho, Wo, 1, dt(y))
        wr, br = f[p + "se_reduce"]
        we, be = f[p + "se_expand"]
        sq = F.linear((pool * (1.0 / (Ho * Wo))).to(cd), wr, br, weight_c=wr, act="swish")
        gate = F.linear(sq, we, be, weight_c=we, out_dtype=torch.float32)
        z = torch.empty_like(y)
        call("lnx_se_scale", y.data_ptr(), gate.data_ptr(), z.data_ptr(), B, Ho * Wo, oup, dt(y))
        wp, bp = f[p + "project"]
        out = F.linear(z, wp, bp, weight_c=wp, residual=x if (stride == 1 and cin == cout) else None)
        return out, Ho, Wo

    def _tblock(self, blk: RelativeMHSABlock, f: dict, p: str, x, B, cd):
        a = blk.attn
        N, D = x.shape[1], x.shape[2]
        t = F.layernorm(x, blk.norm1.weight, blk.norm1.bias, 1e-5)
        qkv = F.linear(t, a.qkv.weight, None, weight_c=f[p + "qkv"])
        o = torch.empty((B, N, D), dtype=cd, device=x.device)
        bias = f[p + "bias"]
        if bias.shape[1] != N:
            raise AssertionError(f"sequence length {N} != patch grid + extra tokens {bias.shape[1]} (pass metadata for every enabled component)")
        call("lnx_attn_bias_fwd", qkv.data_ptr(), bias.data_ptr(), o.data_ptr(), B, a.num_heads, N, D // a.num_heads, float(a.scale), dt(qkv))
        x = F.linear(o, a.proj.weight, a.proj.bias, weight_c=f[p + "proj"], residual=x)
        t = F.layernorm(x, blk.norm2.weight, blk.norm2.bias, 1e-5)
        return F.mlp2(t, blk.mlp.fc1.weight, blk.mlp.fc1.bias, blk.mlp.fc2.weight, blk.mlp.fc2.bias, w1c=f[p + "fc1"], w2c=f[p + "fc2"], act="gelu",
                      residual=x)

    def _extras(self, stage: int, meta, cd):
        if not (self.use_meta and meta is not None and self.meta_components):
            return None
        meta = meta.float().contiguous()
        toks = [_run_meta_head(getattr(self, f"meta_{n.lower()}_head_{stage}"), meta, info["offset"], info["dim"], cd)
                for n, info in self.meta_components.items()]
        return torch.stack(toks, dim=1)

    def forward_features(self, x: torch.Tensor, meta: torch.Tensor | None = None, force_checkpointing: bool | None = None) -> torch.Tensor:
        """R/models/mFormerV0.py:499-660, eval mode."""
        if not x.is_cuda:
            raise RuntimeError("linnaeus_b200.mFormerV0 runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self.training or torch.is_grad_enabled():
            raise NotImplementedError("linnaeus_b200.mFormerV0: only the inference path is built (call model.eval() under torch.no_grad()); "
                                      "the training path of the V0 variant is not implemented yet")
        cd = self._cdtype()
        with torch.autocast("cuda", enabled=False):
            f = self._prepare(cd)
            B, Cin, Hi, Wi = x.shape
            x = x.float().contiguous()
            y, H, W = self._conv3x3(x, f["stem0"], B, Hi, Wi, Cin, 2, "relu", cd, image=True)
            c = f["stem0"][0].shape[0]
            y, H, W = self._conv3x3(y, f["stem1"], B, H, W, c, 1, "relu", cd, w9=f.get("stem1.w9"))
            c = f["stem1"][0].shape[0]
            y, H, W = self._conv3x3(y, f["stem2"], B, H, W, c, 1, "relu", cd, w9=f.get("stem2.w9"))
            c = f["stem2"][0].shape[0]
            Hp, Wp = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
            z = torch.empty((B * Hp * Wp, c), dtype=cd, device=x.device)
            call("lnx_maxpool3s2", y.data_ptr(), z.data_ptr(), B, H, W, c, dt(y))
            y, H, W = z, Hp, Wp
            for si in (1, 2):
                for i, blk in enumerate(getattr(self, f"stage_{si}")):
                    y, H, W = self._mbconv(blk, f, f"s{si}.{i}.", y, B, H, W, cd)
            c = self.conv_out_channels[-1]
            D3, D4 = self.attn_embed_dims
            n_meta = self.extra_token_num - 1
            # stage 3
            blk = self.stage_3[0]
            t, H, W = self._conv3x3(y, f["s3.0.embed"], B, H, W, c, 2, None, cd)
            t = F.layernorm(t, blk.patch_embed.norm.weight, blk.patch_embed.norm.bias, 1e-5)
            x3 = F.tokens_assemble(self.cls_token_1, self._extras(1, meta, cd), t.view(B, H * W, D3))
            for i, blk in enumerate(self.stage_3):
                x3 = self._tblock(blk, f, f"s3.{i}.", x3, B, cd)
            x3 = F.layernorm(x3, self.norm_1.weight, self.norm_1.bias, 1e-5)
            cls1, patches = F.tokens_split(x3, n_meta)
            if not self.only_last_cls:
                mlp, ln = self.cl_1_fc[0], self.cl_1_fc[1]
                c1 = F.mlp2(cls1, mlp.fc1.weight, mlp.fc1.bias, mlp.fc2.weight, mlp.fc2.bias, act="gelu")
                c1 = F.layernorm(c1, ln.weight, ln.bias, 1e-5)
            # stage 4
            blk = self.stage_4[0]
            t, H, W = self._conv3x3(patches.reshape(-1, D3), f["s4.0.embed"], B, H, W, D3, 2, None, cd)
            t = F.layernorm(t, blk.patch_embed.norm.weight, blk.patch_embed.norm.bias, 1e-5)
            x4 = F.tokens_assemble(self.cls_token_2, self._extras(2, meta, cd), t.view(B, H * W, D4))
            for i, blk in enumerate(self.stage_4):
                x4 = self._tblock(blk, f, f"s4.{i}.", x4, B, cd)
            cls2, _ = F.tokens_split(x4, n_meta)
            c2 = F.layernorm(cls2, self.norm_2.weight, self.norm_2.bias, 1e-5)  # norm_2 is per token: only the CLS row is used
            agg = F.aggregate2(c1, c2, self.aggregate.weight, self.aggregate.bias) if not self.only_last_cls else c2
            return F.layernorm(agg, self.norm.weight, self.norm.bias, 1e-5) if not self.only_last_cls else agg

    def forward(self, x: torch.Tensor, meta: torch.Tensor | None = None, force_checkpointing: bool | None = None):
        """-> {task: logits [B, C_k]} in ``head`` insertion order (mFormerV0.py:482-497)."""
        feats = self.forward_features(x, meta, force_checkpointing=force_checkpointing)
        with torch.autocast("cuda", enabled=False):
            ws, bs, offs = [], [], [0]
            for t, head in self.head.items():
                w, b = head.classifier_params()
                ws.append(w)
                bs.append(b if b is not None else torch.zeros(w.shape[0], device=w.device))
                offs.append(offs[-1] + w.shape[0])
            wcat, bcat = torch.cat(ws, 0), torch.cat(bs, 0)
            pad = (-offs[-1]) % 8
            if pad:
                wcat = TF.pad(wcat, (0, 0, 0, pad))
                bcat = TF.pad(bcat, (0, pad))
            cat = F.linear(feats, wcat, bcat, out_dtype=torch.float32)
            out = LogitsDict()
            for i, t in enumerate(self.head.keys()):
                out[t] = cat[:, offs[i]:offs[i + 1]]
            if not pad:
                out.cat = cat
                out.class_off = tuple(offs)
            return out
