"""Model files: kc_modelfile_write / kc_modelfile_load against the format the reference's parser reads
(cpp/neuralnet/desc.cpp).  Host-only: no GPU, no compute calls."""
import ctypes as C
import gzip
import hashlib
import os

import numpy as np
import pytest

from katacoffee_b200 import backend, capi, modeldesc


def _arr(p, n):
    return np.ctypeslib.as_array(p, shape=(n,)).copy()


def _layers(d):
    """Flat list of (name, array) for every weight array of a ModelDesc, in a fixed order."""
    out = []
    conv = lambda n, c: out.append((n, (c.convYSize, c.convXSize, c.inChannels, c.outChannels), _arr(c.weights, c.convYSize * c.convXSize * c.inChannels * c.outChannels)))
    mm = lambda n, m: out.append((n, (m.inChannels, m.outChannels), _arr(m.weights, m.inChannels * m.outChannels)))
    mb = lambda n, m: out.append((n, (m.numChannels,), _arr(m.weights, m.numChannels)))
    def bn(n, b):
        for f in ("mean", "variance", "scale", "bias"):
            out.append((n + "." + f, (b.numChannels, round(b.epsilon, 9)), _arr(getattr(b, f), b.numChannels)))
    conv("initialConv", d.initialConv); mm("initialMatMul", d.initialMatMul)
    for i in range(d.numBlocks):
        b = d.blocks[i]
        out.append((f"b{i}.kind", (b.kind, b.preActivation, b.midActivation, b.gpoolActivation if b.kind == 2 else 0), np.zeros(0)))
        bn(f"b{i}.preBN", b.preBN); conv(f"b{i}.regularConv", b.regularConv)
        if b.kind == 2:
            conv(f"b{i}.gpoolConv", b.gpoolConv); bn(f"b{i}.gpoolBN", b.gpoolBN); mm(f"b{i}.gpoolToBiasMul", b.gpoolToBiasMul)
        bn(f"b{i}.midBN", b.midBN); conv(f"b{i}.finalConv", b.finalConv)
    bn("trunkTipBN", d.trunkTipBN)
    conv("p1Conv", d.p1Conv); conv("g1Conv", d.g1Conv); bn("g1BN", d.g1BN); mm("gpoolToBiasMul", d.gpoolToBiasMul); bn("p1BN", d.p1BN); conv("p2Conv", d.p2Conv)
    conv("v1Conv", d.v1Conv); bn("v1BN", d.v1BN); mm("v2Mul", d.v2Mul); mb("v2Bias", d.v2Bias); mm("v3Mul", d.v3Mul); mb("v3Bias", d.v3Bias)
    mm("sv3Mul", d.sv3Mul); mb("sv3Bias", d.sv3Bias); conv("vOwnershipConv", d.vOwnershipConv)
    out.append(("scalars", (d.version, d.numInputChannels, d.numInputGlobalChannels, d.numBlocks, d.trunkNumChannels, d.midNumChannels, d.regularNumChannels,
                            d.gpoolNumChannels, d.trunkTipActivation, d.g1Activation, d.p1Activation, d.v1Activation, d.v2Activation), np.zeros(0)))
    return out


@pytest.mark.parametrize("suffix", [".bin.gz", ".txt.gz", ".bin", ".txt"])
@pytest.mark.parametrize("net,activation", [("b2c32", "relu"), ("b1c32g", "mish")])
def test_modelfile_round_trip(tmp_path, suffix, net, activation):
    m = modeldesc.Model(net, seed=5, activation=activation)
    path = str(tmp_path / ("model" + suffix))
    backend.writeModelFile(m, path, name="kc-test-" + net)
    raw = open(path, "rb").read()
    mf = backend.ModelFile(path)
    assert mf.name == "kc-test-" + net
    assert mf.sha256 == hashlib.sha256(raw).hexdigest()          # SHA-256 of the file as stored (fileutils.cpp:113-140)
    a, b = _layers(m.desc), _layers(mf.desc)
    assert len(a) == len(b)
    for (na, sa, wa), (nb, sb, wb) in zip(a, b):
        assert na == nb and sa == sb, (na, sa, sb)
        assert wa.tobytes() == wb.tobytes(), na                   # text uses 9 significant digits: exact for fp32
    # the file is what the reference's parser expects: the text form starts name / version / 15 / 1 / trunk header
    text = gzip.decompress(raw) if suffix.endswith(".gz") else raw
    head = text.split(b"\n", 6)
    C_, mid, reg, gp, nb, _, _ = modeldesc.CONFIGS[net][:7]
    assert head[:5] == [b"kc-test-" + net.encode(), b"1", b"15", b"1", b"trunk"]
    assert head[5].split() == [str(x).encode() for x in (nb, C_, mid, reg, reg, gp)]
    assert (b"@BIN@" in text) == (".bin" in suffix)
    mf.close()
    # the expected hash is enforced, case-insensitively
    backend.ModelFile(path, hashlib.sha256(raw).hexdigest().upper()).close()
    with pytest.raises(capi.KCError, match="does not match the expected sha256"):
        backend.ModelFile(path, "0" * 64)


def test_modelfile_ambiguous_gz_and_errors(tmp_path):
    m = modeldesc.Model("b2c32", seed=1)
    txt = str(tmp_path / "a.txt")
    backend.writeModelFile(m, txt)
    # a text model under a plain .gz name: binary is tried first, then text (desc.cpp:1171-1195)
    amb = str(tmp_path / "b.gz")
    with open(amb, "wb") as f:
        f.write(gzip.compress(open(txt, "rb").read()))
    backend.ModelFile(amb).close()
    dat = str(tmp_path / "a.dat")
    open(dat, "wb").write(open(txt, "rb").read())
    with pytest.raises(capi.KCError, match="should end with"):
        backend.ModelFile(dat)
    with pytest.raises(capi.KCError, match="could not open"):
        backend.ModelFile(str(tmp_path / "missing.bin.gz"))
    lines = open(txt).read().split("\n")
    def variant(name, edit):
        l = list(lines); edit(l)
        p = str(tmp_path / name)
        open(p, "w").write("\n".join(l))
        return p
    # a Go-shaped or otherwise wrong model is rejected with the reference's kind of message
    with pytest.raises(capi.KCError, match="version"):
        backend.ModelFile(variant("v.txt", lambda l: l.__setitem__(1, "14")))
    with pytest.raises(capi.KCError, match="spatial"):
        backend.ModelFile(variant("c.txt", lambda l: l.__setitem__(2, "22")))
    with pytest.raises(capi.KCError, match="trunkNumChannels"):
        backend.ModelFile(variant("t.txt", lambda l: l.__setitem__(5, "2 48 32 16 16 16")))
    with pytest.raises(capi.KCError, match="unknown block kind"):
        backend.ModelFile(variant("k.txt", lambda l: l.__setitem__(l.index("ordinary_block"), "weird_block")))
    with pytest.raises(capi.KCError, match="nested bottleneck"):
        backend.ModelFile(variant("n.txt", lambda l: l.__setitem__(l.index("ordinary_block"), "nested_bottleneck_block")))
    with pytest.raises(capi.KCError, match="unexpected end|could not read float"):
        backend.ModelFile(variant("e.txt", lambda l: l.__delitem__(slice(len(l) // 2, None))))
    i = lines.index("p2/w")
    with pytest.raises(capi.KCError):     # policy head with 2 output channels (a Go net): float count no longer matches / != 4
        backend.ModelFile(variant("p.txt", lambda l: l.__setitem__(i + 1, "1 1 32 2 1 1")))
    with pytest.raises(capi.KCError, match="Nan or infinite"):
        backend.ModelFile(variant("nan.txt", lambda l: l.__setitem__(l.index("conv1") + 2, "nan " + l[l.index("conv1") + 2].split(" ", 1)[1])))
    # hostile headers: dimensions whose product wraps size_t, or that no file could fill, are errors -- not crashes, not allocations
    ci = lines.index("conv1")
    for k, dims in enumerate(("3 3 2147483647 2147483647 1 1", "3 3 65536 65536 1 1", "3 3 8192 8192 1 1", "99 99 15 32 1 1")):
        for suffix in ("txt", "bin"):
            with pytest.raises(capi.KCError, match="channels|filter size|too short"):
                backend.ModelFile(variant(f"h{k}.{suffix}", lambda l: l.__setitem__(ci + 1, dims)))
    # the writer emits ReLU activation layers as their name alone, which is the reference's version-1 format (desc.cpp:243-256)
    assert not any(x.startswith("ACTIVATION_") for x in lines)
    # ... and a file that does spell the kind out parses the same
    def add_kinds(l):
        l[:] = [y for x in l for y in ((x, "ACTIVATION_RELU") if x.endswith("/actv") or "/actv" in x else (x,))]
    mf = backend.ModelFile(variant("kinds.txt", add_kinds))
    assert mf.desc.trunkTipActivation == 1 and mf.desc.blocks[0].preActivation == 1
    mf.close()
    with pytest.raises(capi.KCError, match="at least one block"):
        backend.writeModelFile(modeldesc.Model("b0c32", seed=1), str(tmp_path / "z.bin"))
