"""Parity tests proper: the CUDA path, called through the C ABI (mag_solve_batch), against
the CPU oracle on identical inputs.  Bar: tests/parity.py."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from parity import compare

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def solver():
    from magnetizer_b200.solver import DynamoSolver
    s = DynamoSolver(device=0)
    yield s
    s.close()


def _subset(inputs, ids):
    ids = np.asarray(ids, np.int32)
    return ids, np.ascontiguousarray(inputs[:, ids - 1, :])


def test_device_rng_matches_libgfortran_golden(solver):
    gold = json.load(open(os.path.join(HERE, "golden", "rng_golden.json")))
    L = solver._lib
    for entry in gold.values():
        for case in entry["cases"]:
            want = np.array([float(v) for v in case["uniforms"]])
            got = np.empty(len(want))
            rc = L.mag_test_rng(C.c_int32(case["igal"]), C.c_int32(case["p_random_seed"]), C.c_int32(1),
                                C.c_int32(len(want)), got.ctypes.data_as(C.c_void_p))
            assert rc == 0 and np.array_equal(got, want), case["igal"]


def test_device_bessel_matches_oracle(solver, oracle):
    xs = np.concatenate([np.geomspace(1e-7, 2, 200), np.linspace(2, 40, 400)])
    got = np.empty((4, xs.size))
    assert solver._lib.mag_test_bessel(xs.ctypes.data_as(C.c_void_p), C.c_int(xs.size), got.ctypes.data_as(C.c_void_p)) == 0
    L = oracle.lib()
    for which in range(4):
        want = np.array([L.oracle_bessel(which, float(x)) for x in xs])
        assert np.max(np.abs(got[which] / want - 1)) < 5e-14, which


def test_single_galaxy_mode(solver, oracle, example_input):
    """BASELINE config 1: example input, single-galaxy mode, Fortran gal_id=1."""
    t, inputs = example_input
    ids, sub = _subset(inputs, [1])
    gpu = solver.run(ids, t, sub)
    ora = oracle.solve_batch(solver.params, ids, t, sub)
    rep = compare(gpu, ora)
    assert max(rep.values()) <= 1e-8


def test_tutorial_galaxies(solver, oracle, example_input):
    """The three galaxies whose Bmax the reference's tutorial publishes (two of them hit
    p_nsteps_max), checked against the published value as well."""
    t, inputs = example_input
    ids, sub = _subset(inputs, [22, 23, 24])
    gpu = solver.run(ids, t, sub)
    ora = oracle.solve_batch(solver.params, ids, t, sub)
    compare(gpu, ora)
    bmax = gpu["scalars"][gpu["scalar_names"].index("Bmax")][:, -1]
    for got, want in zip(bmax, (0.45010823, 0.42917456, 0.03083575)):
        assert abs(got / want - 1) < 2e-6


def test_example_full_run(solver, oracle, example_input):
    """BASELINE config 2: all 196 galaxies of the example input."""
    t, inputs = example_input
    ids = np.arange(1, inputs.shape[1] + 1, dtype=np.int32)
    gpu = solver.run(ids, t, inputs)
    ora = oracle.solve_batch(solver.params, ids, t, inputs)
    rep = compare(gpu, ora)
    print("worst row-relative errors:", {k: f"{v:.2e}" for k, v in rep.items() if v > 0})


def test_generic_path_agrees_with_fast_path(solver, oracle, example_input):
    """The any-nx generic path and the register fast path are interchangeable."""
    t, inputs = example_input
    ids, sub = _subset(inputs, [1, 5, 24, 60, 101])
    solver.set_fast_path(False)
    try:
        slow = solver.run(ids, t, sub)
    finally:
        solver.set_fast_path(True)
    compare(slow, oracle.solve_batch(solver.params, ids, t, sub))


def test_reduced_quantity_list_and_null_outputs(solver, oracle, example_input):
    t, inputs = example_input
    ids, sub = _subset(inputs, [3, 4])
    prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax",)
    gpu = solver.run(ids, t, sub, profiles=prof, scalars=scal)
    ora = oracle.solve_batch(solver.params, ids, t, sub, profiles=prof, scalars=scal)
    compare(gpu, ora)


def test_invalid_and_single_snapshot_galaxies(solver, oracle, example_input):
    """status 'p' / error=.true. (input_parameters.f90:217-223); ragged validity windows."""
    t, inputs = example_input
    bad = inputs[:, :4, :].copy()
    bad[0, 0, :] = -1.0e10
    bad[0, 1, :] = -1.0e10
    bad[0, 1, 5] = 3.0
    bad[0, 2, 20:] = -1.0e10  # history ends early
    ids = np.array([1, 2, 3, 4], np.int32)
    gpu = solver.run(ids, t, bad)
    ora = oracle.solve_batch(solver.params, ids, t, bad)
    compare(gpu, ora)
    assert gpu["error"].tolist()[:2] == [1, 1]
    assert (gpu["profiles"][:, :2] == -99999.0).all()


def test_synthetic_histories_parity(solver, oracle, example_input):
    """A slice of the synthetic catalogue used by bench.py (BASELINE config 3), ids beyond
    1290 so that the int32 wrap of igal**3 (random.f90:43) is exercised."""
    from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories
    t, inputs = example_input
    first, n = 2000, 48
    syn = synthetic_histories(inputs, n, first=first)
    ids = galaxy_ids(n, first=first)
    prof, scal = ("Br", "Bp", "alp_m", "h", "n", "Omega", "Shear"), ("Bmax", "Bmax_idx", "dt")
    gpu = solver.run(ids, t, syn, profiles=prof, scalars=scal)
    ora = oracle.solve_batch(solver.params, ids, t, syn, profiles=prof, scalars=scal)
    compare(gpu, ora)


def test_linearity_property_of_kinematic_phase(solver, example_input):
    """Size-independent property: the solve is deterministic -- two runs of the same batch in
    different queue orders (different ids order) give bit-identical rows."""
    t, inputs = example_input
    ids, sub = _subset(inputs, [7, 8, 9, 10, 11, 12])
    a = solver.run(ids, t, sub, profiles=("Br", "Bp"), scalars=("Bmax",))
    perm = np.array([5, 3, 1, 0, 2, 4])
    b = solver.run(ids[perm], t, np.ascontiguousarray(sub[:, perm, :]), profiles=("Br", "Bp"), scalars=("Bmax",))
    assert np.array_equal(a["profiles"][:, perm], b["profiles"])
    assert np.array_equal(a["status"][perm], b["status"])
