#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors into small JSON fixtures.

Run in the BUILD container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py [/root/reference]

Sources:
  cpp/core/rand.cpp:41-103      XorShift1024Mult::test   (seed array + 32 expected uint32)
  cpp/core/rand.cpp:120-149     PCG32::test              (state 123 + 16 expected uint32)
  cpp/core/rand.cpp:386-507     Rand("abc") 24 values, MD5 and SHA-256 known answers
  cpp/tests/testnn.cpp:137-916  conv / batchnorm / residual / gpool-residual layer vectors
  cpp/tests/testnn.cpp:931-1001 + cpp/tests/results/runOutputTests.txt  symmetry copies

The fixtures are data (numbers) lifted from the reference's tests; no reference code is copied.
The script is a tiny interpreter for the C++ subset those tests are written in.
"""
import json
import math
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def read(path):
    with open(os.path.join(REF, path), "r", encoding="utf-8", errors="replace") as f:
        return f.read()


# --------------------------------------------------------------------------------------------
# rand.cpp
# --------------------------------------------------------------------------------------------
def extract_hash():
    src = read("cpp/core/rand.cpp")
    out = {}

    def func_body(name):
        i = src.index(name)
        j = src.index("{", i)
        depth, k = 0, j
        while True:
            if src[k] == "{":
                depth += 1
            elif src[k] == "}":
                depth -= 1
                if depth == 0:
                    return src[j:k + 1]
            k += 1

    xs = func_body("void XorShift1024Mult::test()")
    init = re.search(r"init_a\[XORMULT_LEN\]\s*=\s*\{(.*?)\};", xs, re.S).group(1)
    exp = re.search(r"expected\[32\]\s*=\s*\{(.*?)\};", xs, re.S).group(1)
    out["xorshift1024"] = {
        "init_a": [str(int(x.rstrip("UL"))) for x in re.findall(r"\d+ULL", init)],
        "expected": [int(x.rstrip("uU"), 16) for x in re.findall(r"0x[0-9a-fA-F]+[uU]?", exp)],
    }
    pcg = func_body("void PCG32::test()")
    state = int(re.search(r"PCG32 pcg\((\d+)\)", pcg).group(1))
    exp = re.search(r"expected\[16\]\s*=\s*\{(.*?)\};", pcg, re.S).group(1)
    out["pcg32"] = {"state": state,
                    "expected": [int(x.rstrip("uU"), 16) for x in re.findall(r"0x[0-9a-fA-F]+[uU]?", exp)]}
    st = func_body("static void simpleTest()")
    seed = re.search(r'Rand rand\("([^"]*)"\)', st).group(1)
    exp = re.search(r"expected\[24\]\s*=\s*\{(.*?)\};", st, re.S).group(1)
    out["rand"] = {"seed": seed,
                   "expected": [int(x.rstrip("uU"), 16) for x in re.findall(r"0x[0-9a-fA-F]+[uU]?", exp)]}
    md5s = re.search(r'const string s = "([^"]*)";\s*MD5::get', st).group(1)
    md5v = [int(x, 16) for x in re.findall(r"hash\[\d\] == (0x[0-9A-Fa-f]+)", st)]
    out["md5"] = {"msg": md5s, "expected": md5v}
    m = re.search(r'SHA2::get256\("([^"]*)", hash\);\s*if\(string\(hash\) != string\("([0-9a-f]+)"\)', st)
    out["sha256"] = [{"msg": m.group(1), "hex": m.group(2)}]
    raw = re.search(r'R"%%\((.*?)\)%%"', st, re.S).group(1)
    lines = raw.split("\n")[1:]  # raw string starts with a newline
    # 10 (message, sha256) pairs come first, then 10 (message, sha512) pairs
    for i in range(10):
        out["sha256"].append({"msg": lines[2 * i], "hex": lines[2 * i + 1]})
    return out


# --------------------------------------------------------------------------------------------
# testnn.cpp mini interpreter
# --------------------------------------------------------------------------------------------
NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?f?"


def parse_floats(text):
    text = re.sub(r"//[^\n]*", "", text)
    return [float(x.rstrip("f")) for x in re.findall(NUM, text)]


def strip_comments(s):
    return re.sub(r"//[^\n]*", "", s)


def function_text(src, header):
    i = src.index(header)
    j = src.index("{", i)
    depth, k = 0, j
    while True:
        if src[k] == "{":
            depth += 1
        elif src[k] == "}":
            depth -= 1
            if depth == 0:
                return src[j + 1:k]
        k += 1


def split_statements(body):
    """Yield ('open',) ('close',) or ('stmt', text) walking a function body."""
    i, n = 0, len(body)
    cur = ""
    paren = 0
    while i < n:
        c = body[i]
        if c == "(":
            paren += 1
            cur += c
        elif c == ")":
            paren -= 1
            cur += c
        elif c == "{" and paren == 0:
            head = cur.strip()
            if head.startswith("auto testConfigurations") or head.startswith("for"):
                # skip/capture a whole braced region
                depth, k = 0, i
                while True:
                    if body[k] == "{":
                        depth += 1
                    elif body[k] == "}":
                        depth -= 1
                        if depth == 0:
                            break
                    k += 1
                region = body[i:k + 1]
                if head.startswith("for"):
                    yield ("for", head, region)
                i = k + 1
                # lambda definitions end with ';'
                while i < n and body[i] in " \t\n;":
                    i += 1
                cur = ""
                continue
            if head:
                # something like "ResidualBlockDesc desc" should not precede a brace; treat as stmt
                yield ("stmt", head)
            cur = ""
            yield ("open",)
        elif c == "}" and paren == 0:
            cur = ""
            yield ("close",)
        elif c == ";" and paren == 0:
            s = cur.strip()
            if s:
                yield ("stmt", s)
            cur = ""
        else:
            cur += c
        i += 1


class Env:
    def __init__(self):
        self.scopes = [{}]

    def push(self):
        self.scopes.append({})

    def pop(self):
        self.scopes.pop()

    def set(self, k, v):
        self.scopes[-1][k] = v

    def assign(self, k, v):
        for s in reversed(self.scopes):
            if k in s:
                s[k] = v
                return
        self.scopes[-1][k] = v

    def get(self, k):
        for s in reversed(self.scopes):
            if k in s:
                return s[k]
        raise KeyError(k)

    def has(self, k):
        return any(k in s for s in self.scopes)


def eval_value(env, text):
    text = text.strip()
    m = re.fullmatch(r"vector<float>\(\{(.*)\}\)", text, re.S)
    if m:
        return parse_floats(m.group(1))
    if text in ("true", "false"):
        return text == "true"
    if re.fullmatch(NUM, text):
        return float(text.rstrip("f")) if ("." in text or "f" in text or "e" in text) else int(text)
    if re.fullmatch(r'"[^"]*"', text):
        return text[1:-1]
    if re.fullmatch(r"[A-Za-z_]\w*", text):
        return env.get(text)
    raise ValueError("cannot evaluate: " + text[:60])


def set_path(d, path, v):
    parts = path.split(".")
    for p in parts[:-1]:
        d = d.setdefault(p, {})
    d[parts[-1]] = v


def run_function(body, kind):
    env = Env()
    cases = []
    for item in split_statements(strip_comments(body)):
        if item[0] == "open":
            env.push()
        elif item[0] == "close":
            env.pop()
        elif item[0] == "for":
            head, region = item[1], item[2]
            m = re.search(r"int i = (\d+); i<(\d+)", head)
            lo, hi = int(m.group(1)), int(m.group(2))
            expr = re.search(r"expected\[i\] \+= \(float\)\((.*?)\);", region, re.S).group(1)
            val = eval(expr.replace("\n", " "), {"sqrt": math.sqrt})
            expected = list(env.get("expected"))
            mask = env.get("mask")
            for i in range(lo, hi):
                # float arithmetic as in the C++ (float += float, then *= mask)
                expected[i] = (expected[i] + float(val)) * mask[i]
            env.assign("expected", expected)
        else:
            s = item[1]
            m = re.fullmatch(r"int (\w+) = (\d+)", s)
            if m:
                env.set(m.group(1), int(m.group(2)))
                continue
            m = re.fullmatch(r'string label\("([^"]*)"\)', s)
            if m:
                env.set("label", m.group(1))
                continue
            m = re.fullmatch(r"vector<float> (\w+)\(\{(.*)\}\)", s, re.S)
            if m:
                env.set(m.group(1), parse_floats(m.group(2)))
                continue
            m = re.fullmatch(r"\w+Desc desc", s)
            if m:
                env.set("desc", {})
                continue
            m = re.fullmatch(r"desc\.([\w.]+) = (.*)", s, re.S)
            if m:
                set_path(env.get("desc"), m.group(1), eval_value(env, m.group(2)))
                continue
            m = re.fullmatch(r"testConfigurations\((.*)\)", s, re.S)
            if m:
                args = [a.strip() for a in m.group(1).split(",")]
                case = {"kind": kind, "label": env.get(args[0]),
                        "batchSize": env.get(args[1]), "nnXLen": env.get(args[2]),
                        "nnYLen": env.get(args[3]),
                        "desc": json.loads(json.dumps(env.get(args[4])))}
                if len(args) == 7:
                    case["input"], case["expected"] = env.get(args[5]), env.get(args[6])
                else:
                    case["input"], case["mask"], case["expected"] = (
                        env.get(args[5]), env.get(args[6]), env.get(args[7]))
                cases.append(case)
                continue
            # anything else (ActivationLayerDesc defaults etc.) is ignored on purpose
    return cases


def extract_nn_layers():
    src = read("cpp/tests/testnn.cpp")
    cases = []
    cases += run_function(function_text(src, "static void testConvLayer("), "conv")
    cases += run_function(function_text(src, "static void testBatchNormLayer("), "batchnorm")
    cases += run_function(function_text(src, "static void testResidualBlock("), "resblock")
    cases += run_function(function_text(src, "static void testGlobalPoolingResidualBlock("), "gpoolblock")
    return cases


def extract_symmetry():
    src = strip_comments(function_text(read("cpp/tests/testnn.cpp"), "void Tests::runNNSymmetryTests()"))
    # the two input vectors and the four calls, in order
    inputs = [parse_floats(m) for m in re.findall(r"vector<float> input\(\{(.*?)\}\);", src, re.S)]
    calls = re.findall(r'testConfigurations\("([^"]+)",(\d+),(\d+),(\d+),(\d+),input\)', src)
    assert len(inputs) == 2 and len(calls) == 4
    res = read("cpp/tests/results/runOutputTests.txt").split("\n")
    out = []
    for idx, (label, n, c, xlen, ylen) in enumerate(calls):
        case = {"label": label, "batchSize": int(n), "numChannels": int(c), "nnXLen": int(xlen),
                "nnYLen": int(ylen), "input": inputs[idx // 2], "inputs_sym": [], "outputs_sym": []}
        # find the block of lines for this label
        start = next(i for i, l in enumerate(res) if l.strip() == f"{label} useNHWC 0 0")
        i = start
        for useNHWC in range(2):
            for sym in range(8):
                assert res[i].strip() == f"{label} useNHWC {useNHWC} {sym}", res[i]
                case["inputs_sym"].append({"useNHWC": useNHWC, "symmetry": sym,
                                           "expected": [float(x) for x in res[i + 1].split()]})
                i += 2
        for sym in range(8):
            assert res[i].strip() == f"{label} OUTPUT", res[i]
            case["outputs_sym"].append({"symmetry": sym,
                                        "expected": [float(x) for x in res[i + 1].split()]})
            i += 2
        out.append(case)
    return out


def main():
    with open(os.path.join(OUT, "hash_golden.json"), "w") as f:
        json.dump(extract_hash(), f, indent=1)
    with open(os.path.join(OUT, "nn_layers_golden.json"), "w") as f:
        json.dump(extract_nn_layers(), f)
    with open(os.path.join(OUT, "nn_symmetry_golden.json"), "w") as f:
        json.dump(extract_symmetry(), f)
    print("wrote hash_golden.json, nn_layers_golden.json, nn_symmetry_golden.json in", OUT)


if __name__ == "__main__":
    main()
