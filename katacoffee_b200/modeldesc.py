"""Builds kc_model_desc structures (POD mirror of the reference's ModelDesc, cpp/neuralnet/desc.h:13-304).

Random-init weights follow SURVEY.md 8(d): He-style N(0, 2/fan_in) for layers followed by ReLU,
N(0, 1/fan_in) for final linear layers, the second conv of each block scaled by 1/sqrt(numBlocks)
so the trunk stays O(1), BN mean 0 / var 1 / eps 1e-4 / scale 1 / bias N(0, 0.1^2), ReLU everywhere.
Trunk shapes come from the reference's python/modelconfigs.py:129-247 (b6c96, b10c128, b15c192);
head shapes are the Coffee ones: policy 4 channels, value 2, misc 2, ownership 1.
Weight layouts are the ones desc.cpp produces: conv oc,ic,y,x; matmul ic,oc.
"""
import ctypes as C

import numpy as np

from . import capi

CONFIGS = {
    # name: (trunk, mid, regular, gpool, numBlocks, gpool block indices, v2 size)
    "b2c32": (32, 32, 16, 16, 2, (1,), 32),      # tiny net for fast tests
    "b6c96": (96, 96, 64, 32, 6, (2, 4), 64),    # modelconfigs.py b6c96: gpool blocks 3 and 5 (1-based)
    "b10c128": (128, 128, 96, 32, 10, (4, 7), 80),
    "b15c192": (192, 192, 128, 64, 15, (6, 11), 96),
}
HEAD_C = 32


class Model:
    """Owns the numpy weight arrays and the ctypes description pointing into them."""

    def __init__(self, name, seed=0):
        if name not in CONFIGS:
            raise ValueError(f"unknown net {name}")
        self.name = name
        C_, mid, reg, gp, nb, gpool_blocks, v2 = CONFIGS[name]
        self.trunk, self.num_blocks, self.v2 = C_, nb, v2
        rng = np.random.default_rng(seed)
        self._keep = []
        d = capi.ModelDesc()
        d.version = 1
        d.numInputChannels, d.numInputGlobalChannels, d.numBlocks = 15, 1, nb
        d.trunkNumChannels, d.midNumChannels, d.regularNumChannels, d.gpoolNumChannels = C_, mid, reg, gp
        d.trunkTipActivation = d.g1Activation = d.p1Activation = d.v1Activation = d.v2Activation = 1
        d.initialConv = self._conv(rng, 3, 15, C_, relu=True)
        d.initialMatMul = self._matmul(rng, 1, C_, scale=0.3)
        blocks = (capi.BlockDesc * nb)()
        self.flops_per_eval_cell = 2.0 * (9 * 15 * C_ + C_)
        for i in range(nb):
            b = blocks[i]
            b.preActivation = b.gpoolActivation = b.midActivation = 1
            b.preBN = self._bn(rng, C_)
            resid = 1.0 / np.sqrt(nb)
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

    # -- helpers: every array is kept alive on self._keep --
    def _arr(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        self._keep.append(a)
        return a.ctypes.data_as(capi.c_float_p)

    def _conv(self, rng, k, ic, oc, relu, mult=1.0):
        fan_in = k * k * ic
        std = np.sqrt((2.0 if relu else 1.0) / fan_in) * mult
        w = rng.standard_normal((oc, ic, k, k)) * std
        return capi.ConvDesc(k, k, ic, oc, self._arr(w))

    def _bn(self, rng, c):
        return capi.BNDesc(c, 1e-4, 1, 1, self._arr(np.zeros(c)), self._arr(np.ones(c)),
                           self._arr(np.ones(c)), self._arr(rng.standard_normal(c) * 0.1))

    def _matmul(self, rng, ic, oc, scale=1.0):
        w = rng.standard_normal((ic, oc)) * np.sqrt(scale / ic)
        return capi.MatMulDesc(ic, oc, self._arr(w))

    def _bias(self, rng, c):
        return capi.MatBiasDesc(c, 0, self._arr(rng.standard_normal(c) * 0.1))


def flops_per_eval(name, hw):
    """Algorithmic FLOPs of one forward = 2 x direct-conv/matmul MACs."""
    C_, mid, reg, gp, nb, gpool_blocks, v2 = CONFIGS[name]
    macs = hw * 9 * 15 * C_ + C_                       # initial conv + global matmul
    for i in range(nb):
        if i in gpool_blocks:
            macs += hw * 9 * C_ * (reg + gp) + 3 * gp * reg + hw * 9 * reg * C_
        else:
            macs += hw * 9 * C_ * mid + hw * 9 * mid * C_
    macs += hw * C_ * 3 * HEAD_C + 3 * HEAD_C * HEAD_C + hw * HEAD_C * 4       # p1,g1,v1, gpool bias, p2
    macs += 3 * HEAD_C * v2 + v2 * 2 + v2 * 2 + hw * HEAD_C                     # v2, v3, sv3, ownership
    return 2 * macs
