"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports
every entry point include/magnetizer_b200.h declares; the struct mirrors agree with the
header; without a GPU the solve entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "magnetizer_b200.h")


@pytest.fixture(scope="module")
def lib():
    from magnetizer_b200.solver import LIB_PATH, load_library
    if not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "magnetizer_b200", "csrc")])
    return load_library()


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mag_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared_functions()
    assert {"mag_create", "mag_destroy", "mag_solve_batch", "mag_solve_batch_device", "mag_params_default",
            "mag_last_error", "mag_last_timing", "mag_measure_fp64_peak"} <= set(names)
    for n in names:
        assert hasattr(lib, n), n


def test_params_struct_matches_header_and_oracle(lib, oracle):
    from magnetizer_b200.abi import MagParams
    p = MagParams()
    lib.mag_params_default(C.byref(p))
    q = oracle.default_params()
    assert bytes(p) == bytes(q)  # product defaults == oracle defaults, byte for byte
    # compile a probe against the header: sizeof/offsetof must agree with the ctypes mirror
    probe = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "magnetizer_b200.h"
    int main(void){ printf("%zu %zu %zu %zu\n", sizeof(mag_params_t), offsetof(mag_params_t, rng_kind),
                           offsetof(mag_params_t, reserved), sizeof(mag_outputs_t)); return 0; }
    """
    exe = os.path.join("/tmp", f"mag_probe_{os.getpid()}")
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=probe.encode(), check=True)
    out = subprocess.check_output([exe]).split()
    os.remove(exe)
    from magnetizer_b200.abi import MagOutputs
    assert [int(v) for v in out] == [C.sizeof(MagParams), MagParams.rng_kind.offset, MagParams.reserved.offset,
                                     C.sizeof(MagOutputs)]


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from magnetizer_b200.solver import DynamoSolver, MagnetizerError
    with pytest.raises(MagnetizerError, match="MAG_ERR_NO_DEVICE"):
        DynamoSolver()


def test_unsupported_switch_is_rejected(lib):
    from magnetizer_b200.abi import MagParams
    p = MagParams()
    lib.mag_params_default(C.byref(p))
    p.Damp = 1
    ctx = C.c_void_p()
    assert lib.mag_create(C.byref(p), -1, C.byref(ctx)) == -3  # MAG_ERR_UNSUPPORTED, checked before any device use


def test_synthetic_generator_is_shard_consistent(example_input):
    from magnetizer_b200.synthetic import synthetic_histories
    _, inputs = example_input
    full = synthetic_histories(inputs, 9000)
    part = synthetic_histories(inputs, 700, first=4000)
    assert np.array_equal(full[:, 4000:4700], part)
    assert np.array_equal(np.sort(full[0, :196], axis=0), np.sort(inputs[0], axis=0))  # first 196 = the example
    assert np.sign(full).tolist() == np.sign(inputs[:, np.random.Generator(np.random.PCG64(20240607)).permutation(196)[np.arange(9000) % 196]]).tolist()
