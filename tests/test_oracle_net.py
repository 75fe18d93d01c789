"""Pins the oracle's layer arithmetic to the literal vectors of the reference's tests/testnn.cpp
(conv 1x1/3x3/5x5 :137-340, batchnorm :375-475, residual block :508-677, gpool residual block
:710-916) and its symmetry copies to tests/results/runOutputTests.txt:20220ff, with the reference's
own fp32 tolerance (testnn.cpp:8-15)."""
import numpy as np
import pytest

from conftest import golden


def approx_equal(a, b):
    # testnn.cpp:8-15: |x-y| < 1e-4 * max(|x|,|y|,1)
    tol = 1e-4 * np.maximum(np.maximum(np.abs(a), np.abs(b)), 1.0)
    return bool((np.abs(a - b) < tol).all())


def nchw_to_nhwc(v, n, c, h, w):
    return np.asarray(v, np.float32).reshape(n, c, h, w).transpose(0, 2, 3, 1).reshape(-1)


LAYER_CASES = golden("nn_layers_golden.json")


@pytest.mark.parametrize("case", LAYER_CASES, ids=[c["label"] for c in LAYER_CASES])
@pytest.mark.parametrize("nhwc", [0, 1])
@pytest.mark.parametrize("mode", [0, 1])
def test_layer_golden(oracle, case, nhwc, mode):
    from katacoffee_b200 import backend
    n, xl, yl = case["batchSize"], case["nnXLen"], case["nnYLen"]
    d = case["desc"]
    kind = case["kind"]
    if kind == "conv":
        ic, oc = d["inChannels"], d["outChannels"]
    elif kind == "batchnorm":
        ic = oc = d["numChannels"]
    else:
        ic = oc = d["preBN"]["numChannels"]
    inp = np.asarray(case["input"], np.float32)
    exp = np.asarray(case["expected"], np.float32)
    if nhwc:
        inp, exp = nchw_to_nhwc(inp, n, ic, yl, xl), nchw_to_nhwc(exp, n, oc, yl, xl)
    if kind == "conv":
        s, keep = backend._desc_conv(d)
        out = oracle.test_conv(s, n, xl, yl, nhwc, inp, oc, mode)
    elif kind == "batchnorm":
        if mode == 1:
            pytest.skip("batchnorm has one implementation")
        s, keep = backend._desc_bn(d)
        out = oracle.test_batchnorm(s, 0, n, xl, yl, nhwc, inp, case["mask"])   # identity activation (testEvaluateBatchNorm)
    else:
        s, keep = backend.block_desc_from_dict(d)
        out = oracle.test_resblock(s, n, xl, yl, nhwc, inp, case["mask"], mode)
    assert approx_equal(out[:exp.size], exp), (out, exp)


SYM_CASES = golden("nn_symmetry_golden.json")


@pytest.mark.parametrize("case", SYM_CASES, ids=[c["label"] for c in SYM_CASES])
def test_symmetry_golden(oracle, case):
    n, c = case["batchSize"], case["numChannels"]
    # the reference test passes (nnXLen, nnYLen) in the (hSize, wSize) slots (testnn.cpp:941-943)
    h, w = case["nnXLen"], case["nnYLen"]
    src = np.asarray(case["input"], np.float32)
    for e in case["inputs_sym"]:
        inp = nchw_to_nhwc(src, n, c, w, h) if e["useNHWC"] else src   # NCHWtoNHWC(input,n,c,nnYLen,nnXLen)
        out = oracle.copy_inputs_with_symmetry(inp, n, h, w, c, e["useNHWC"], e["symmetry"])
        assert out.tolist() == e["expected"], (case["label"], e["useNHWC"], e["symmetry"])
    for e in case["outputs_sym"]:
        out = oracle.copy_outputs_with_symmetry(src, n * c, h, w, e["symmetry"])
        assert out.tolist() == e["expected"], (case["label"], "OUTPUT", e["symmetry"])


def test_forward_modes_agree_and_symmetry_roundtrip(oracle):
    """Winograd (the Eigen algorithm) and direct convolution agree to fp32 tolerance on a whole net,
    and evaluating a symmetric copy of a position returns the symmetric outputs un-rotated."""
    from katacoffee_b200 import modeldesc
    m = modeldesc.Model("b2c32", seed=3)
    om = oracle.Model(m)
    recs, planes, glob = oracle.playout_run(5, 5, 4, seed=7, g0=0, n=6)
    planes, glob = planes[:40], glob[:40].reshape(-1, 1)
    p0, v0, m0, o0 = om.forward(planes, glob, 5, 5, mode=0)
    p1, v1, m1, o1 = om.forward(planes, glob, 5, 5, mode=1)
    for a, b in ((p0, p1), (v0, v1), (m0, m1), (o0, o1)):
        assert np.abs(a - b).max() < 2e-4
    # NHWC rows give the same result as NCHW rows
    nhwc = planes.reshape(-1, 15, 25).transpose(0, 2, 1).reshape(-1, 375)
    p2, v2, _, _ = om.forward(nhwc, glob, 5, 5, nhwc=True, mode=0)
    assert np.abs(p2 - p0).max() < 1e-6 and np.abs(v2 - v0).max() < 1e-6
    assert np.isfinite(p0).all() and p0.std() > 1e-3


def test_postprocess_matches_formulae(oracle):
    rng = np.random.default_rng(1)
    policy = rng.standard_normal(100).astype(np.float32)
    legal = np.zeros(4, np.uint32)
    idx = [3, 17, 40, 99]
    for i in idx:
        legal[i >> 5] |= np.uint32(1 << (i & 31))
    p, v, m = oracle.postprocess(policy, legal, np.array([0.3, -0.2], np.float32), np.array([0.5, -1.0], np.float32), next_pla=1)
    e = np.exp(policy[idx] - policy[idx].max())
    assert np.allclose(p[idx], e / e.sum(), atol=1e-6)
    assert (np.delete(p, idx) == -1).all()
    w = 1 / (1 + np.exp(-0.5))
    assert np.allclose(v, [1 - w, w], atol=1e-6)     # black to move: flipped to white's view
    assert np.allclose(m[0], np.log1p(np.exp(0.5)) * 40, rtol=1e-6)
    assert np.allclose(m[1], np.sqrt(np.log1p(np.exp(-0.5)) ** 2 * 0.25), rtol=1e-6)
