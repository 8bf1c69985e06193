"""Regenerates tests/golden/out_charset.txt from the reference's constant (src/weinman/mjsynth.py:23).
Run in the container that has /root/reference; the GPU box only sees the committed copy."""
import os
import re

REF = "/root/reference/src/weinman/mjsynth.py"


def read_reference_charset(path=REF):
    src = open(path, encoding="utf-8").read()
    m = re.search(r'^out_charset\s*=\s*(".*")\s*$', src, re.M)
    return eval(m.group(1))   # a plain string literal


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "out_charset.txt")
    open(out, "w", encoding="utf-8").write(read_reference_charset())
    print("wrote", out)
