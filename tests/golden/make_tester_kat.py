"""Extracts the known-answer vectors of the reference's legacy self-test (src/toolbox/Tester.cpp) into
tests/golden/tester_kat.json.  Run in the build container (needs /root/reference); the JSON is committed.

The self-test is print-only and never wired to a target; its inputs are literal
`xt::xarray<T> name{...}` initialisers and its expected values live in `/**correct result(s)` comments:
    blendShape          Tester.cpp:298-470   (V = 1)
    jointRegression     Tester.cpp:541-660   (V = 5)
    worldTransformation Tester.cpp:700-860   (arbitrary 3x3 "rotations")
    linearBlendSkinning Tester.cpp:880-1041  (V = 1, sum of weights != 1, random 4x4 transforms)
"""
import json
import os
import re
import sys

SRC = "/root/reference/src/toolbox/Tester.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tester_kat.json")

NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"


def nested(text):
    """Parse a brace- or bracket-nested numeric literal into nested lists."""
    text = text.replace("{", "[").replace("}", "]")
    text = re.sub(r"\s+", "", text)
    text = re.sub(r",\]", "]", text)
    text = re.sub(r"(?<![\d.])\.(\d)", r"0.\1", text)
    text = re.sub(r"(\d)\.(?!\d)", r"\1.0", text)
    return json.loads(text)


def function_body(src, name):
    m = re.search(r"void Tester::%s\(\)[^{]*\{" % name, src)
    start = m.end()
    nxt = re.search(r"\nvoid Tester::\w+\(", src[start:])
    return src[start:start + nxt.start()] if nxt else src[start:]


def arrays(body):
    out = {}
    for m in re.finditer(r"xt::xarray<\w+>\s+(\w+)\s*\{", body):
        i = m.end() - 1
        depth, j = 0, i
        while True:
            depth += body[j] == "{"
            depth -= body[j] == "}"
            j += 1
            if depth == 0:
                break
        out[m.group(1).rstrip("_")] = nested(body[i:j])
    return out


def expected(body):
    m = re.search(r"/\*\*correct results?(.*?)\*/", body, re.S)
    text = "\n".join(line.lstrip(" *") for line in m.group(1).splitlines())
    out = {}
    for em in re.finditer(r"-\s*(\w+):\s*[\[(][^\n]*\n(.*?)(?=\n-\s*\w+:|\Z)", text, re.S):
        out[em.group(1)] = nested(em.group(2).strip())
    return out


def main():
    src = open(SRC).read()
    kat = {}
    for fn in ("blendShape", "jointRegression", "worldTransformation", "linearBlendSkinning"):
        body = function_body(src, fn)
        kat[fn] = {"inputs": arrays(body), "expected": expected(body)}
    with open(OUT, "w") as f:
        json.dump(kat, f)
    for fn, d in kat.items():
        print(fn, "inputs:", {k: _shape(v) for k, v in d["inputs"].items()}, "expected:",
              {k: _shape(v) for k, v in d["expected"].items()})


def _shape(x):
    s = []
    while isinstance(x, list):
        s.append(len(x))
        x = x[0]
    return tuple(s)


if __name__ == "__main__":
    sys.exit(main())
