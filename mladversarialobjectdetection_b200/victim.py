"""Random-init EfficientDet-shaped victim on the framework's own GPU path (cuDNN through torch).

The reference's victim is the vendored Keras EfficientDet (automl/efficientdet/tf2/efficientdet_keras.py:778-994,
called with pre_mode=None, post_mode=None at attacker.py:98,125); TensorFlow is not available in this image and
the north star keeps the detector's convolutions on the framework path, so this module is a stand-in with the
same interface and cost shape: NHWC float32 images in, five class / box level tensors out in NHWC
([B,h_l,w_l,9*90], [B,h_l,w_l,9*4], levels 3..7).  It exists so the attack step can be timed end to end;
nothing in the parity suite depends on its weights.  Architecture tables: hparams_config.py:302-346.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as Fn


@dataclass
class VictimConfig:
    name: str = "efficientdet-d0"
    image_size: int = 512
    width_mult: float = 1.0           # EfficientNet width / depth coefficients
    depth_mult: float = 1.0
    fpn_channels: int = 64
    fpn_layers: int = 3
    head_layers: int = 3
    num_classes: int = 90
    min_level: int = 3
    max_level: int = 7
    num_scales: int = 3
    aspect_ratios: Tuple[float, ...] = (1.0, 2.0, 0.5)
    anchor_scale: float = 4.0
    mean_rgb: float = 127.0           # only used for on-disk de-normalisation (attacker.py:338)
    stddev_rgb: float = 128.0
    nms_configs: dict = field(default_factory=lambda: dict(method="gaussian", iou_thresh=None, score_thresh=0.0,
                                                           sigma=None, max_nms_inputs=0, max_output_size=100))

    def override(self, d: dict) -> None:
        """Config.override subset used by PatchAttacker(config_override=...) (hparams_config.py:85-115)."""
        for k, v in d.items():
            if isinstance(v, dict) and isinstance(getattr(self, k, None), dict):
                getattr(self, k).update(v)
            elif hasattr(self, k):
                setattr(self, k, v)
            else:
                raise KeyError(f"unknown config key {k}")


CONFIGS = {
    "efficientdet-d0": dict(image_size=512, width_mult=1.0, depth_mult=1.0, fpn_channels=64, fpn_layers=3, head_layers=3),
    "efficientdet-d1": dict(image_size=640, width_mult=1.0, depth_mult=1.1, fpn_channels=88, fpn_layers=4, head_layers=3),
    "efficientdet-d2": dict(image_size=768, width_mult=1.1, depth_mult=1.2, fpn_channels=112, fpn_layers=5, head_layers=3),
    "efficientdet-d3": dict(image_size=896, width_mult=1.2, depth_mult=1.4, fpn_channels=160, fpn_layers=6, head_layers=4),
    "efficientdet-d4": dict(image_size=1024, width_mult=1.4, depth_mult=1.8, fpn_channels=224, fpn_layers=7, head_layers=4),
    "efficientdet-lite4": dict(image_size=640, width_mult=1.4, depth_mult=1.8, fpn_channels=224, fpn_layers=7, head_layers=4),
}


def get_config(name: str) -> VictimConfig:
    return VictimConfig(name=name, **CONFIGS[name])


def _round_filters(c: int, mult: float, divisor: int = 8) -> int:
    c *= mult
    new = max(divisor, int(c + divisor / 2) // divisor * divisor)
    if new < 0.9 * c:
        new += divisor
    return int(new)


# ---- fused bias + activation epilogue (CUDA only) ---------------------------------------------------------------
# After BatchNorm folding every convolution carries a bias.  PyTorch's cuDNN path applies it as a separate broadcast
# add (a non-vectorised kernel in channels_last: a quarter of the victim's time at D0 / 512^2) and SiLU as a further
# pass; libeotpatch's nhwc_bias_act_fwd / nhwc_bias_silu_bwd do both in one 128-bit pass.  The convolution itself stays
# on cuDNN.  On the CPU (oracle arm, CPU tests) the plain torch ops run.
FUSED_EPILOGUE = True


class _BiasAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, act):
        from . import ops
        ctx.save_for_backward(x, bias)
        ctx.act = act
        return ops.nhwc_bias_act(x, bias, act, out=torch.empty_like(x))

    @staticmethod
    def backward(ctx, grad):
        from . import ops
        x, bias = ctx.saved_tensors
        if not ctx.act:
            return grad, None, None
        return ops.nhwc_bias_silu_backward(x, bias, grad), None, None


def _nhwc_ok(y: torch.Tensor) -> bool:
    return (y.is_cuda and y.dtype == torch.float32 and y.dim() == 4 and y.shape[1] % 4 == 0 and y.shape[2] * y.shape[3] > 1
            and y.is_contiguous(memory_format=torch.channels_last))


def _nhwc_ok2(y: torch.Tensor, act: bool) -> bool:
    """forward-only / identity epilogue also takes even channel counts (the 810-channel class logits)"""
    return (not act and y.is_cuda and y.dtype == torch.float32 and y.dim() == 4 and y.shape[1] % 2 == 0
            and y.shape[2] * y.shape[3] > 1 and y.is_contiguous(memory_format=torch.channels_last))


def conv_bias_act(x: torch.Tensor, weight, bias, stride, padding, dilation, groups, act: bool) -> torch.Tensor:
    """act(conv2d(x, weight) + bias): cuDNN convolution + one fused epilogue pass on CUDA, torch ops elsewhere."""
    if FUSED_EPILOGUE and x.is_cuda and bias is not None:
        y = Fn.conv2d(x, weight, None, stride, padding, dilation, groups)
        if _nhwc_ok(y) or _nhwc_ok2(y, act):
            if torch.is_grad_enabled() and y.requires_grad:
                return _BiasAct.apply(y, bias, act)
            from . import ops
            return ops.nhwc_bias_act(y, bias, act, out=y)                  # nothing to differentiate: in place
        y = y + bias.view(1, -1, 1, 1)
        return Fn.silu(y) if act else y
    y = Fn.conv2d(x, weight, bias, stride, padding, dilation, groups)
    return Fn.silu(y) if act else y


class _SqueezeExcite(torch.autograd.Function):
    """y * sigmoid(W2 silu(W1 mean_hw(y) + b1) + b2) with the two full-tensor passes (the gate product and, backward,
    dy = dout * gate + dmean / HW and dgate = sum_hw dout * y) in libeotpatch; the [N,C]-sized MLP stays in torch."""

    @staticmethod
    def forward(ctx, y, w1, b1, w2, b2):
        from . import ops
        s = y.mean((2, 3))
        h_pre = torch.addmm(b1, s, w1.t())
        h = Fn.silu(h_pre)
        g = torch.sigmoid(torch.addmm(b2, h, w2.t())).contiguous()
        ctx.save_for_backward(y, g, h_pre, h, w1, w2)
        return ops.nhwc_channel_scale(y, g)

    @staticmethod
    def backward(ctx, dout):
        from . import ops
        y, g, h_pre, h, w1, w2 = ctx.saved_tensors
        if not dout.is_contiguous(memory_format=torch.channels_last):
            dout = dout.contiguous(memory_format=torch.channels_last)
        dg = ops.nhwc_channel_dot(dout, y)
        dg_pre = dg * g * (1.0 - g)
        dh = dg_pre @ w2
        sg = torch.sigmoid(h_pre)
        ds = ((dh * (sg * (1.0 + h_pre * (1.0 - sg)))) @ w1).contiguous()
        dy = ops.nhwc_channel_scale(dout, g, ds, 1.0 / float(y.shape[2] * y.shape[3]))
        return dy, None, None, None, None


def squeeze_excite(y: torch.Tensor, se_reduce: nn.Conv2d, se_expand: nn.Conv2d) -> torch.Tensor:
    if FUSED_EPILOGUE and _nhwc_ok(y):
        w1, w2 = se_reduce.weight.flatten(1), se_expand.weight.flatten(1)
        if torch.is_grad_enabled() and y.requires_grad:
            return _SqueezeExcite.apply(y, w1, se_reduce.bias, w2, se_expand.bias)
        from . import ops
        h = Fn.silu(torch.addmm(se_reduce.bias, y.mean((2, 3)), w1.t()))
        g = torch.sigmoid(torch.addmm(se_expand.bias, h, w2.t())).contiguous()
        return ops.nhwc_channel_scale(y, g, out=y)                        # nothing to differentiate: in place
    s = y.mean((2, 3), keepdim=True)
    return y * torch.sigmoid(se_expand(Fn.silu(se_reduce(s))))


class _FuseSilu(torch.autograd.Function):
    """silu(sum_i w_i x_i) of the BiFPN's fast normalised fusion in one pass (weights are frozen: no dL/dw)."""

    @staticmethod
    def forward(ctx, w, *xs):
        from . import ops
        ctx.save_for_backward(w, *xs)
        return ops.nhwc_fuse_silu(xs, w)

    @staticmethod
    def backward(ctx, dout):
        from . import ops
        w, *xs = ctx.saved_tensors
        if dout.stride() != xs[0].stride():
            dout = dout.contiguous(memory_format=torch.channels_last)
        return (None, *ops.nhwc_fuse_silu_backward(xs, w, dout, ctx.needs_input_grad[1:]))


def fuse_silu(xs: Sequence[torch.Tensor], w: torch.Tensor) -> torch.Tensor:
    x0 = xs[0]
    if (FUSED_EPILOGUE and len(xs) in (2, 3) and x0.is_cuda and x0.dtype == torch.float32 and x0.numel() % 4 == 0
            and x0.is_contiguous(memory_format=torch.channels_last)
            and all(x.shape == x0.shape and x.stride() == x0.stride() and x.dtype == x0.dtype for x in xs[1:])):
        w = w.contiguous()
        if torch.is_grad_enabled() and any(x.requires_grad for x in xs):
            return _FuseSilu.apply(w, *xs)
        from . import ops
        return ops.nhwc_fuse_silu(xs, w)
    y = xs[0] * w[0]
    for i in range(1, len(xs)):
        y = y + xs[i] * w[i]
    return Fn.silu(y)


def _module_conv_bias_act(conv: nn.Conv2d, x, act: bool):
    return conv_bias_act(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups, act)


class ConvBN(nn.Sequential):
    def __init__(self, cin, cout, k=1, stride=1, groups=1, act=True):
        layers = [nn.Conv2d(cin, cout, k, stride, k // 2, groups=groups, bias=False), nn.BatchNorm2d(cout, eps=1e-3)]
        if act:
            layers.append(nn.SiLU(inplace=True))
        super().__init__(*layers)

    def forward(self, x):
        if isinstance(self[1], nn.Identity):                                # BatchNorm folded into the conv
            return _module_conv_bias_act(self[0], x, act=len(self) == 3)
        return super().forward(x)


class MBConv(nn.Module):
    def __init__(self, cin, cout, expand, k, stride, se_ratio=0.25):
        super().__init__()
        mid = cin * expand
        self.expand = ConvBN(cin, mid, 1) if expand != 1 else nn.Identity()
        self.dw = ConvBN(mid, mid, k, stride, groups=mid)
        se = max(1, int(cin * se_ratio))
        self.se_reduce = nn.Conv2d(mid, se, 1)
        self.se_expand = nn.Conv2d(se, mid, 1)
        self.project = ConvBN(mid, cout, 1, act=False)
        self.skip = stride == 1 and cin == cout

    def forward(self, x):
        y = squeeze_excite(self.dw(self.expand(x)), self.se_reduce, self.se_expand)
        y = self.project(y)
        return x + y if self.skip else y


class Backbone(nn.Module):
    """EfficientNet-B* trunk returning the stride 8/16/32 features."""
    BLOCKS = [(1, 3, 1, 16, 1), (6, 3, 2, 24, 2), (6, 5, 2, 40, 2), (6, 3, 2, 80, 3), (6, 5, 1, 112, 3),
              (6, 5, 2, 192, 4), (6, 3, 1, 320, 1)]

    def __init__(self, width, depth):
        super().__init__()
        c = _round_filters(32, width)
        self.stem = ConvBN(3, c, 3, 2)
        stages, self.out_channels, self.taps = [], [], []
        for i, (e, k, s, co, r) in enumerate(self.BLOCKS):
            co = _round_filters(co, width)
            blocks = []
            for j in range(int(math.ceil(r * depth))):
                blocks.append(MBConv(c, co, e, k, s if j == 0 else 1))
                c = co
            stages.append(nn.Sequential(*blocks))
            if i in (2, 4, 6):
                self.out_channels.append(c)
        self.stages = nn.ModuleList(stages)

    def forward(self, x):
        x = self.stem(x)
        feats = []
        for i, st in enumerate(self.stages):
            x = st(x)
            if i in (2, 4, 6):
                feats.append(x)
        return feats


class SepConvBN(nn.Sequential):
    def __init__(self, c, act=False):
        super().__init__(nn.Conv2d(c, c, 3, 1, 1, groups=c, bias=False), nn.Conv2d(c, c, 1, bias=True),
                         nn.BatchNorm2d(c, eps=1e-3))

    def forward(self, x):
        if isinstance(self[2], nn.Identity):
            return _module_conv_bias_act(self[1], self[0](x), act=False)
        return super().forward(x)


class Fuse(nn.Module):
    """Fast normalised fusion (weights relu'd and normalised) + swish + separable conv + BN."""

    def __init__(self, n, c):
        super().__init__()
        self.w = nn.Parameter(torch.ones(n))
        self.conv = SepConvBN(c)

    def forward(self, xs: Sequence[torch.Tensor]):
        w = Fn.relu(self.w)
        w = w / (w.sum() + 1e-4)
        return self.conv(fuse_silu(xs, w))


def _down(x):
    return Fn.max_pool2d(x, 3, 2, 1)


def _up(x, ref):
    return Fn.interpolate(x, size=ref.shape[-2:], mode="nearest")


class BiFPNLayer(nn.Module):
    def __init__(self, c, in_channels=None):
        super().__init__()
        self.lateral = None
        if in_channels is not None:      # first layer: project backbone features (two copies for P4/P5 like upstream)
            self.lateral = nn.ModuleList([ConvBN(ci, c, 1, act=False) for ci in in_channels])
            self.lateral2 = nn.ModuleList([ConvBN(ci, c, 1, act=False) for ci in in_channels[1:]])
            self.p6 = ConvBN(in_channels[-1], c, 1, act=False)
        self.td = nn.ModuleList([Fuse(2, c) for _ in range(4)])       # P6', P5', P4', P3 out
        self.bu = nn.ModuleList([Fuse(3, c) for _ in range(3)] + [Fuse(2, c)])   # P4, P5, P6, P7 out

    def forward(self, feats: List[torch.Tensor]):
        if self.lateral is not None:
            c3, c4, c5 = feats
            p6 = _down(self.p6(c5))
            p7 = _down(p6)
            p3, p4, p5 = self.lateral[0](c3), self.lateral[1](c4), self.lateral[2](c5)
            p4b, p5b = self.lateral2[0](c4), self.lateral2[1](c5)
        else:
            p3, p4, p5, p6, p7 = feats
            p4b, p5b = p4, p5
        t6 = self.td[0]([p6, _up(p7, p6)])
        t5 = self.td[1]([p5, _up(t6, p5)])
        t4 = self.td[2]([p4, _up(t5, p4)])
        o3 = self.td[3]([p3, _up(t4, p3)])
        o4 = self.bu[0]([p4b, t4, _down(o3)])
        o5 = self.bu[1]([p5b, t5, _down(o4)])
        o6 = self.bu[2]([p6, t6, _down(o5)])
        o7 = self.bu[3]([p7, _down(o6)])
        return [o3, o4, o5, o6, o7]


class Head(nn.Module):
    """Class / box net: shared separable convs, per-level BN, swish, then the prediction conv."""

    def __init__(self, c, layers, out, levels=5, bias_init=0.0):
        super().__init__()
        self.dw = nn.ModuleList([nn.Conv2d(c, c, 3, 1, 1, groups=c, bias=False) for _ in range(layers)])
        self.pw = nn.ModuleList([nn.Conv2d(c, c, 1) for _ in range(layers)])
        self.bn = nn.ModuleList([nn.ModuleList([nn.BatchNorm2d(c, eps=1e-3) for _ in range(levels)]) for _ in range(layers)])
        self.out_dw = nn.Conv2d(c, c, 3, 1, 1, groups=c, bias=False)
        self.out_pw = nn.Conv2d(c, out, 1)
        nn.init.constant_(self.out_pw.bias, bias_init)

        self.folded = None

    def fold_batchnorm(self):
        """Inference-mode BatchNorm is a per-channel affine map: fold it into the shared pointwise conv, one
        (weight, bias) pair per level (the reference runs the victim with is_training_bn=False)."""
        ws, bs = nn.ParameterList(), nn.ParameterList()
        for i in range(len(self.dw)):
            for bn in self.bn[i]:
                k = bn.weight / torch.sqrt(bn.running_var + bn.eps)
                ws.append(nn.Parameter((self.pw[i].weight * k.view(-1, 1, 1, 1)).detach(), requires_grad=False))
                bs.append(nn.Parameter(((self.pw[i].bias - bn.running_mean) * k + bn.bias).detach(), requires_grad=False))
        self.fold_w, self.fold_b = ws, bs
        self.folded = len(self.bn[0])

    def forward(self, feats):
        outs = []
        for lvl, x in enumerate(feats):
            for i in range(len(self.dw)):
                if self.folded:
                    k = i * self.folded + lvl
                    x = conv_bias_act(self.dw[i](x), self.fold_w[k], self.fold_b[k], 1, 0, 1, 1, act=True)
                else:
                    x = Fn.silu(self.bn[i][lvl](self.pw[i](self.dw[i](x))))
            outs.append(_module_conv_bias_act(self.out_pw, self.out_dw(x), act=False))
        return outs


class EfficientDetVictim(nn.Module):
    """`model(images, pre_mode=None, post_mode=None) -> (cls_outputs, box_outputs)` (attacker.py:98)."""

    def __init__(self, config: VictimConfig):
        super().__init__()
        self.config = config
        self.backbone = Backbone(config.width_mult, config.depth_mult)
        c = config.fpn_channels
        self.fpn = nn.ModuleList([BiFPNLayer(c, self.backbone.out_channels if i == 0 else None)
                                  for i in range(config.fpn_layers)])
        na = config.num_scales * len(config.aspect_ratios)
        self.class_net = Head(c, config.head_layers, na * config.num_classes,
                              bias_init=-math.log((1 - 0.01) / 0.01))      # efficientdet_keras.py:469
        self.box_net = Head(c, config.head_layers, na * 4)
        self.eval()                                                          # is_training_bn=False (infer_lib.py:171)
        for p in self.parameters():
            p.requires_grad_(False)                                          # only [scale, patch] train (attacker.py:63)

    def forward(self, images: torch.Tensor, pre_mode=None, post_mode=None):
        if pre_mode is not None or post_mode is not None:
            raise NotImplementedError("the attack path calls the victim with pre_mode=None, post_mode=None")
        x = images.permute(0, 3, 1, 2)                                       # NHWC memory == channels_last view
        feats = self.backbone(x)
        for layer in self.fpn:
            feats = layer(feats)
        cls = [o.permute(0, 2, 3, 1) for o in self.class_net(feats)]         # NHWC views, no copy
        box = [o.permute(0, 2, 3, 1) for o in self.box_net(feats)]
        return cls, box


def _fold_conv_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d) -> nn.Conv2d:
    k = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    out = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation,
                    conv.groups, bias=True)
    bias = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
    with torch.no_grad():
        out.weight.copy_(conv.weight * k.view(-1, 1, 1, 1))
        out.bias.copy_((bias - bn.running_mean) * k + bn.bias)
    return out


def fold_batchnorm(model: "EfficientDetVictim") -> None:
    """conv -> BatchNorm(eval) pairs become one conv with bias (exact algebra; the usual inference-graph rewrite).
    The victim always runs with is_training_bn=False (tf2/infer_lib.py:171), so nothing else ever sees the BN."""
    for mod in model.modules():
        if isinstance(mod, ConvBN):
            mod[0] = _fold_conv_bn(mod[0], mod[1])
            mod[1] = nn.Identity()
        elif isinstance(mod, SepConvBN):
            mod[1] = _fold_conv_bn(mod[1], mod[2])
            mod[2] = nn.Identity()
        elif isinstance(mod, Head):
            mod.fold_batchnorm()
    for p in model.parameters():
        p.requires_grad_(False)


def get_victim_model(name: str = "efficientdet-d0", device="cuda", seed: int = 0, image_size: int = None,
                     fold_bn: bool = True) -> EfficientDetVictim:
    """Counterpart of util.get_victim_model (util.py:177-189) with random-init weights."""
    cfg = get_config(name)
    if image_size is not None:
        cfg.image_size = image_size
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    model = EfficientDetVictim(cfg)
    _calibrate_batchnorm(model)
    if fold_bn:
        fold_batchnorm(model)
    torch.random.set_rng_state(g)
    return model.to(device=device, memory_format=torch.channels_last)


def _calibrate_batchnorm(model: EfficientDetVictim, size: int = 128) -> None:
    """Data-dependent init: one pass of random images (on the CPU, so every rank gets bit-identical weights)
    sets each BatchNorm's running statistics to the statistics it actually sees.  Without it a random-init
    network of this depth has vanishing activations (logits == bias, dL/dimage ~ 1e-15) and would be a
    degenerate victim for step-level tests, although its cost would be the same."""
    bns = [m for m in model.modules() if isinstance(m, nn.BatchNorm2d)]
    for bn in bns:
        bn.reset_running_stats()
        bn.momentum = None            # cumulative average
    model.train()
    with torch.no_grad():
        model(torch.rand(2, size, size, 3) * 2 - 1)
    model.eval()
    for bn in bns:
        bn.momentum = 0.1
    # the prediction convs have no BN behind them: scale them so logits / box regressions are O(1)
    with torch.no_grad():
        cls, box = model(torch.rand(2, size, size, 3) * 2 - 1)
        for head, outs, target in ((model.class_net, cls, 1.5), (model.box_net, box, 0.3)):
            bias = head.out_pw.bias.view(1, 1, 1, -1)
            std = torch.cat([(o - bias).reshape(-1) for o in outs]).std()
            head.out_pw.weight.mul_(target / float(std))
