"""Builds kc_model_desc structures (POD mirror of the reference's ModelDesc, cpp/neuralnet/desc.h:13-304).

Random-init weights follow SURVEY.md 8(d): He-style N(0, 2/fan_in) for layers followed by ReLU,
N(0, 1/fan_in) for final linear layers, the second conv of each block scaled by 1/sqrt(numBlocks)
so the trunk stays O(1), BN mean 0 / eps 1e-4 / scale 1 / bias N(0, 0.1^2), ReLU everywhere.
`calibrate` then fixes the BN running statistics on synthetic V1-like planes so that activations
and logits are O(1) at every depth: "rms" (default) keeps mean 0 as SURVEY.md prescribes and sets
variance = E[x^2]; "full" also sets the means (the numeric profile of a trained net, where the
mean subtraction amplifies reduced-precision error); "none" leaves mean 0 / variance 1.
Trunk shapes come from the reference's python/modelconfigs.py:129-247 (b6c96, b10c128, b15c192);
head shapes are the Coffee ones: policy 4 channels, value 2, misc 2, ownership 1.
Weight layouts are the ones desc.cpp produces: conv oc,ic,y,x; matmul ic,oc.
"""
import ctypes as C

import numpy as np

from . import capi

CONFIGS = {
    # name: (trunk, mid, regular, gpool, numBlocks, gpool block indices, v2 size[, head channels = 32])
    "b0c32": (32, 32, 16, 16, 0, (), 32),        # shallow nets: bf16 rounding has not decorrelated yet, so the
    "b1c32": (32, 32, 16, 16, 1, (), 32),        #   tensor-core path can be compared POINTWISE with the bf16-emulating oracle
    "b1c32g": (32, 32, 16, 16, 1, (0,), 32),
    "b2c32": (32, 32, 16, 16, 2, (1,), 32),      # tiny net for fast tests
    "b6c96": (96, 96, 64, 32, 6, (2, 4), 64),    # modelconfigs.py b6c96: gpool blocks 3 and 5 (1-based)
    "b10c128": (128, 128, 96, 32, 10, (4, 7), 80),
    "b15c192": (192, 192, 128, 64, 15, (6, 11), 96),
    "b1c192g": (192, 192, 128, 64, 1, (0,), 96),   # shallow 192-wide net: pointwise check of the one-tile-per-CTA kernel
    "b2c256": (256, 256, 192, 64, 2, (1,), 96),    # the 256-wide single-tile kernel (modelconfigs.py b20c256 / b40c256 shapes), shallow for pointwise checks
    "b6c256": (256, 256, 192, 64, 6, (2, 4), 96),
    "b20c256": (256, 256, 192, 64, 20, (6, 11, 16), 112, 48),   # modelconfigs.py:249-288 b20c256: gpool blocks 7, 12, 17, 48-channel heads, v2 112
    "b2c256h48": (256, 256, 192, 64, 2, (1,), 112, 48),         # shallow forms of the b20c256 / b40c256 head widths
    "b2c256h64": (256, 256, 192, 64, 2, (0,), 128, 64),
    "b2c128h48": (128, 128, 96, 32, 2, (1,), 80, 48),           # wide heads on a narrow trunk: not a tensor-path shape, must be rejected
    "b1c320": (320, 320, 256, 64, 1, (), 96),      # wider than the tensor-core kernel supports: must be rejected, not emulated
}
HEAD_C = 32


class Model:
    """Owns the numpy weight arrays and the ctypes description pointing into them."""

    def __init__(self, name, seed=0, calibrate="rms", activation="relu"):
        if name not in CONFIGS:
            raise ValueError(f"unknown net {name}")
        if activation not in ("relu", "mish"):
            raise ValueError("activation must be 'relu' or 'mish' (cpp/neuralnet/activations.h:4-6)")
        self.name = name
        self.act = act = 1 if activation == "relu" else 2
        C_, mid, reg, gp, nb, gpool_blocks, v2 = CONFIGS[name][:7]
        HEAD_C = self.head_c = CONFIGS[name][7] if len(CONFIGS[name]) > 7 else 32
        self.trunk, self.num_blocks, self.v2 = C_, nb, v2
        rng = np.random.default_rng(seed)
        self._keep = []
        self._np = {}      # ctypes struct address -> numpy arrays of that layer (for calibration)
        d = capi.ModelDesc()
        d.version = 1
        d.numInputChannels, d.numInputGlobalChannels, d.numBlocks = 15, 1, nb
        d.trunkNumChannels, d.midNumChannels, d.regularNumChannels, d.gpoolNumChannels = C_, mid, reg, gp
        d.trunkTipActivation = d.g1Activation = d.p1Activation = d.v1Activation = d.v2Activation = act
        d.initialConv = self._conv(rng, 3, 15, C_, relu=True)
        d.initialMatMul = self._matmul(rng, 1, C_, scale=0.3)
        blocks = (capi.BlockDesc * nb)()
        self.flops_per_eval_cell = 2.0 * (9 * 15 * C_ + C_)
        for i in range(nb):
            b = blocks[i]
            b.preActivation = b.gpoolActivation = b.midActivation = act
            b.preBN = self._bn(rng, C_)
            resid = 1.0 / np.sqrt(max(nb, 1))
            if i in gpool_blocks:
                b.kind = 2
                b.regularConv = self._conv(rng, 3, C_, reg, relu=True)
                b.gpoolConv = self._conv(rng, 3, C_, gp, relu=True)
                b.gpoolBN = self._bn(rng, gp)
                b.gpoolToBiasMul = self._matmul(rng, 3 * gp, reg, scale=0.5)
                b.midBN = self._bn(rng, reg)
                b.finalConv = self._conv(rng, 3, reg, C_, relu=True, mult=resid)
            else:
                b.kind = 0
                b.regularConv = self._conv(rng, 3, C_, mid, relu=True)
                b.midBN = self._bn(rng, mid)
                b.finalConv = self._conv(rng, 3, mid, C_, relu=True, mult=resid)
        self._keep.append(blocks)
        d.blocks = C.cast(blocks, C.POINTER(capi.BlockDesc))
        d.trunkTipBN = self._bn(rng, C_)
        d.p1Conv = self._conv(rng, 1, C_, HEAD_C, relu=True)
        d.g1Conv = self._conv(rng, 1, C_, HEAD_C, relu=True)
        d.g1BN = self._bn(rng, HEAD_C)
        d.gpoolToBiasMul = self._matmul(rng, 3 * HEAD_C, HEAD_C, scale=0.5)
        d.p1BN = self._bn(rng, HEAD_C)
        d.p2Conv = self._conv(rng, 1, HEAD_C, 4, relu=False)
        d.v1Conv = self._conv(rng, 1, C_, HEAD_C, relu=True)
        d.v1BN = self._bn(rng, HEAD_C)
        d.v2Mul = self._matmul(rng, 3 * HEAD_C, v2, scale=1.0)
        d.v2Bias = self._bias(rng, v2)
        d.v3Mul = self._matmul(rng, v2, 2, scale=1.0)
        d.v3Bias = self._bias(rng, 2)
        d.sv3Mul = self._matmul(rng, v2, 2, scale=1.0)
        d.sv3Bias = self._bias(rng, 2)
        d.vOwnershipConv = self._conv(rng, 1, HEAD_C, 1, relu=False)
        self.desc = d
        self._blocks = blocks
        self._gpool_blocks = gpool_blocks
        if calibrate not in ("none", "rms", "full"):
            raise ValueError("calibrate must be 'none', 'rms' or 'full'")
        if calibrate != "none":
            self._calibrate(np.random.default_rng(seed + 1000003), full=(calibrate == "full"))

    # -- helpers: every array is kept alive on self._keep --
    def _arr(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        self._keep.append(a)
        return a.ctypes.data_as(capi.c_float_p)

    def _conv(self, rng, k, ic, oc, relu, mult=1.0):
        fan_in = k * k * ic
        std = np.sqrt((2.0 if relu else 1.0) / fan_in) * mult
        w = np.ascontiguousarray(rng.standard_normal((oc, ic, k, k)) * std, np.float32)
        d = capi.ConvDesc(k, k, ic, oc, self._arr(w))
        self._np[C.addressof(d.weights.contents)] = w
        return d

    def _bn(self, rng, c):
        arrs = [np.zeros(c, np.float32), np.ones(c, np.float32), np.ones(c, np.float32),
                np.ascontiguousarray(rng.standard_normal(c) * 0.1, np.float32)]
        d = capi.BNDesc(c, 1e-4, 1, 1, *[self._arr(a) for a in arrs])
        self._np[C.addressof(d.mean.contents)] = arrs
        return d

    def _matmul(self, rng, ic, oc, scale=1.0):
        w = np.ascontiguousarray(rng.standard_normal((ic, oc)) * np.sqrt(scale / ic), np.float32)
        d = capi.MatMulDesc(ic, oc, self._arr(w))
        self._np[C.addressof(d.weights.contents)] = w
        return d

    # -- calibration of the BN running statistics on synthetic planes (host-side, torch CPU) --
    def _w(self, d):
        return self._np[C.addressof(d.weights.contents)]

    def _calibrate(self, rng, full, n=192, H=5, W=5):
        import torch
        import torch.nn.functional as F
        x = np.zeros((n, 15, H, W), np.float32)
        x[:, 0] = 1
        stones = rng.random((n, H, W))
        fill = rng.random((n, 1, 1)) * 0.8
        x[:, 1] = stones < fill / 2
        x[:, 2] = (stones >= fill / 2) & (stones < fill)
        for c in range(3, 11):                      # one-hot move planes
            idx = rng.integers(0, H * W, n)
            on = rng.random(n) < (0.25 if c < 7 else 0.8)
            x[np.arange(n)[on], c, idx[on] // W, idx[on] % W] = 1
        x[:, 11] = (rng.random((n, H, W)) < 0.2) & (x[:, 1] + x[:, 2] == 0)
        for c in (12, 13, 14):
            x[:, c] = (rng.random((n, H, W)) < 0.3) & (x[:, 1] + x[:, 2] > 0)
        x = torch.from_numpy(x)
        g = torch.full((n, 1), 4.0)

        def conv(t, d):
            w = torch.from_numpy(self._w(d))
            return F.conv2d(t, w, padding=d.convYSize // 2)

        def bn_fit(t, d):
            mean, var, scale, bias = self._np[C.addressof(d.mean.contents)]
            dims = (0, 2, 3) if t.dim() == 4 else (0,)
            if full:
                mean[:] = t.mean(dims).numpy()
                var[:] = t.var(dims, unbiased=False).numpy() + 1e-3
            else:
                var[:] = (t * t).mean(dims).numpy() + 1e-3
            s = torch.from_numpy(scale / np.sqrt(var + d.epsilon))
            b = torch.from_numpy(bias) - torch.from_numpy(mean) * s
            shape = (1, -1, 1, 1) if t.dim() == 4 else (1, -1)
            y = t * s.view(shape) + b.view(shape)
            return torch.relu(y) if self.act == 1 else F.mish(y)

        def gpool(t):
            mean = t.mean((2, 3))
            sq = float(np.sqrt(H * W))
            return torch.cat([mean, mean * ((sq - 14.0) * 0.1), t.amax((2, 3))], 1)

        d = self.desc
        trunk = conv(x, d.initialConv) + (g @ torch.from_numpy(self._w(d.initialMatMul)))[:, :, None, None]
        for i in range(self.num_blocks):
            b = self._blocks[i]
            a = bn_fit(trunk, b.preBN)
            reg = conv(a, b.regularConv)
            if b.kind == 2:
                gp = bn_fit(conv(a, b.gpoolConv), b.gpoolBN)
                reg = reg + (gpool(gp) @ torch.from_numpy(self._w(b.gpoolToBiasMul)))[:, :, None, None]
            trunk = trunk + conv(bn_fit(reg, b.midBN), b.finalConv)
        tip = bn_fit(trunk, d.trunkTipBN)
        g1 = bn_fit(conv(tip, d.g1Conv), d.g1BN)
        p1 = conv(tip, d.p1Conv) + (gpool(g1) @ torch.from_numpy(self._w(d.gpoolToBiasMul)))[:, :, None, None]
        bn_fit(p1, d.p1BN)
        bn_fit(conv(tip, d.v1Conv), d.v1BN)

    def _bias(self, rng, c):
        return capi.MatBiasDesc(c, 0, self._arr(rng.standard_normal(c) * 0.1))


def flops_per_eval(name, hw):
    """Algorithmic FLOPs of one forward = 2 x direct-conv/matmul MACs."""
    C_, mid, reg, gp, nb, gpool_blocks, v2 = CONFIGS[name][:7]
    HEAD_C = CONFIGS[name][7] if len(CONFIGS[name]) > 7 else 32
    macs = hw * 9 * 15 * C_ + C_                       # initial conv + global matmul
    for i in range(nb):
        if i in gpool_blocks:
            macs += hw * 9 * C_ * (reg + gp) + 3 * gp * reg + hw * 9 * reg * C_
        else:
            macs += hw * 9 * C_ * mid + hw * 9 * mid * C_
    macs += hw * C_ * 3 * HEAD_C + 3 * HEAD_C * HEAD_C + hw * HEAD_C * 4       # p1,g1,v1, gpool bias, p2
    macs += 3 * HEAD_C * v2 + v2 * 2 + v2 * 2 + hw * HEAD_C                     # v2, v3, sv3, ownership
    return 2 * macs
