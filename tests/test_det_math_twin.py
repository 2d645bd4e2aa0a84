"""The deterministic elementary functions exist twice -- in the CPU oracle
(oracle/det_math.hpp) and in the CUDA path (magnetizer_b200/csrc/mag_det_math.cuh) --
because product code may not depend on oracle/ and vice versa.  Parity at 1e-8 needs them
to agree bit for bit (see the header comments), so the two function bodies must be the
same text; the GPU test test_device_det_math_bit_exact checks the values."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _twin(path):
    s = open(path).read()
    b = s.index("// ==== BEGIN TWIN")
    e = s.index("// ==== END TWIN ====")
    body = s[b:e].split("\n", 1)[1]
    return re.sub(r"[ \t]+\n", "\n", body)


def test_twin_blocks_are_identical():
    a = _twin(os.path.join(ROOT, "oracle", "det_math.hpp"))
    b = _twin(os.path.join(ROOT, "magnetizer_b200", "csrc", "mag_det_math.cuh"))
    assert a == b
    assert "fma(" not in a  # no fused multiply-add: CPU and GPU must round identically


def test_bessel_tables_are_identical():
    def lits(path):
        return re.findall(r"^\s+(-?[0-9][0-9.e+-]*)[,}]", open(path).read(), flags=re.M)
    a = lits(os.path.join(ROOT, "oracle", "bessel_coeffs.hpp"))
    b = lits(os.path.join(ROOT, "magnetizer_b200", "csrc", "mag_bessel_coeffs.cuh"))
    assert len(a) == 28 + 28 + 29 + 29 + 41 + 41 + 42 and a == b
