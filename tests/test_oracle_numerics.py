"""Pins the oracle's restatement of the third-party arithmetic that is NOT under
/root/reference (GSL root finders, libgfortran RNG, Bessel functions) and of the small
numerical kernels of the reference (deriv.f90, interpolation.f90, constants.f90)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_rng_xorshift1024_matches_vendored_libgfortran(oracle):
    """random_seed(put=[w]*33) + random_number(REAL(8)) of the GCC-8 libgfortran shipped in
    this image (tests/golden/make_rng_golden.py) -- bit exact."""
    gold = json.load(open(os.path.join(HERE, "golden", "rng_golden.json")))
    L = oracle.lib()
    ncase = 0
    for libname, entry in gold.items():
        assert entry["seed_size"] == 33
        for case in entry["cases"]:
            want = np.array([float(v) for v in case["uniforms"]])
            got = np.empty(len(want))
            L.oracle_rng_uniforms(C.c_int32(case["word"]), C.c_int32(1), C.c_int32(len(want)),
                                  got.ctypes.data_as(C.c_void_p))
            assert np.array_equal(got, want), (libname, case["igal"])
            ncase += 1
    assert ncase >= 9


def test_rng_survey_known_answers(oracle):
    # SURVEY.md App. B known answers (igal=1 -> seed word 288, igal=2 -> 281)
    L = oracle.lib()
    got = np.empty(4)
    L.oracle_rng_uniforms(C.c_int32(288), C.c_int32(1), C.c_int32(4), got.ctypes.data_as(C.c_void_p))
    assert got.tolist() == [0.4116549659881128, 0.4758300016969982, 0.2941695546009686, 0.012520593305444416]
    L.oracle_rng_uniforms(C.c_int32(281), C.c_int32(1), C.c_int32(4), got.ctypes.data_as(C.c_void_p))
    assert got.tolist() == [0.46624808572178067, 0.8738134113888357, 0.24373077809543708, 0.5316872877499449]


def test_rng_xoshiro_is_uniform_and_deterministic(oracle):
    # no golden stream is available for GCC>=9's xoshiro256** in this image (parity unpinned
    # for that generator): check determinism, range and the first two moments only
    L = oracle.lib()
    a, b = np.empty(20000), np.empty(20000)
    L.oracle_rng_uniforms(C.c_int32(288), C.c_int32(0), C.c_int32(a.size), a.ctypes.data_as(C.c_void_p))
    L.oracle_rng_uniforms(C.c_int32(288), C.c_int32(0), C.c_int32(b.size), b.ctypes.data_as(C.c_void_p))
    assert np.array_equal(a, b)
    assert 0.0 <= a.min() and a.max() < 1.0
    assert abs(a.mean() - 0.5) < 0.01 and abs(a.var() - 1 / 12) < 0.005


@pytest.mark.parametrize("kind", [0, 1])
def test_random_normal_moments(oracle, kind):
    L = oracle.lib()
    z = np.empty(40000)
    L.oracle_rng_normals(C.c_int32(7), C.c_int32(17), C.c_int32(kind), C.c_int32(z.size), z.ctypes.data_as(C.c_void_p))
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02


def test_bessel_against_scipy(oracle):
    sp = pytest.importorskip("scipy.special")
    L = oracle.lib()
    xs = np.concatenate([np.geomspace(1e-7, 1, 40), np.linspace(1, 30, 60)])
    for which, ref in enumerate([sp.i0, sp.i1, sp.k0, sp.k1]):
        got = np.array([L.oracle_bessel(which, float(x)) for x in xs])
        want = ref(xs)
        assert np.max(np.abs(got / want - 1)) < 2e-14, which


def test_brent_and_secant(oracle):
    L = oracle.lib()
    root, iters = C.c_double(0), C.c_int32(0)
    ok = L.oracle_brent_power(C.c_double(2.0), C.c_double(2.0), C.c_double(1.0), C.c_double(2.0), C.byref(root),
                              C.byref(iters))
    assert ok == 1 and abs(root.value / np.sqrt(2) - 1) < 1e-7 and 3 <= iters.value <= 12
    # no sign change -> failure, as gsl_root_fsolver_set returns GSL_EINVAL
    ok = L.oracle_brent_power(C.c_double(2.0), C.c_double(2.0), C.c_double(2.0), C.c_double(3.0), C.byref(root),
                              C.byref(iters))
    assert ok == 0
    ok = L.oracle_secant_power(C.c_double(3.0), C.c_double(5.0), C.c_double(1.0), C.c_double(1e-6), C.byref(root))
    assert ok == 1 and abs(root.value / 5 ** (1 / 3) - 1) < 1e-5


def test_stencils_exact_on_polynomials(oracle):
    """deriv.f90: the 6th-order centred stencils and the 7-point one-sided closures
    differentiate polynomials of degree <= 6 exactly."""
    L = oracle.lib()
    nx, dx = 23, 0.125
    x = np.arange(nx) * dx
    for deg in range(7):
        f = x ** deg
        d1, d2 = np.empty(nx), np.empty(nx)
        L.oracle_stencils(C.c_int32(nx), C.c_double(dx), f.ctypes.data_as(C.c_void_p), d1.ctypes.data_as(C.c_void_p),
                          d2.ctypes.data_as(C.c_void_p))
        e1 = deg * x ** max(deg - 1, 0) if deg >= 1 else np.zeros(nx)
        e2 = deg * (deg - 1) * x ** max(deg - 2, 0) if deg >= 2 else np.zeros(nx)
        scale = max(1.0, np.abs(f).max())
        assert np.max(np.abs(d1 - e1)) < 1e-10 * scale / dx, deg
        assert np.max(np.abs(d2 - e2)) < 1e-9 * scale / dx ** 2, deg


def test_rescale_array_is_linear_interpolation(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(3)
    for n, m in [(30, 20), (30, 40), (121, 101), (101, 101), (254, 104)]:
        y = rng.normal(size=n)
        out = np.empty(m)
        L.oracle_rescale_array(y.ctypes.data_as(C.c_void_p), C.c_int32(n), out.ctypes.data_as(C.c_void_p), C.c_int32(m))
        want = np.interp(np.arange(m) / (m - 1), np.arange(n) / (n - 1), y)
        assert np.allclose(out, want, rtol=0, atol=1e-13)
        assert out[0] == y[0] and abs(out[-1] - y[-1]) < 1e-14


def test_constants(oracle):
    L = oracle.lib()
    assert L.oracle_constant(0) == 3.14156295358  # the reference's pi (constants.f90:24), NOT math.pi
    assert abs(L.oracle_constant(1) - 2.9334) < 2e-3  # t0_Gyr
    p = oracle.default_params()
    # float-rounded default literals (tutorial dump: P_COURANT_V: 0.09000000357627869)
    assert p.p_courant_v == 0.09000000357627869 and p.p_courant_eta == 0.09000000357627869
    assert p.p_ISM_turbulent_length == float(np.float32(0.1))
    assert p.p_stellarHeightToRadiusScale == 1.0 / float(np.float32(7.3))
    assert p.p_nx_ref == 101 and p.p_nsteps_max == 20000 and p.p_random_seed == 17


def test_det_math_accuracy_against_host_libm(oracle):
    """oracle/det_math.hpp: log and exp within 1 ulp of the host libm, pow within 2e-14
    relative (the reference's libm is unpinned; see the header for why these are pinned)."""
    L = oracle.lib()
    rng = np.random.default_rng(5)

    def call(which, x, y=None):
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y if y is not None else x, np.float64)
        out = np.empty_like(x)
        L.oracle_det_math(C.c_int32(which), C.c_int32(x.size), x.ctypes.data_as(C.c_void_p),
                          y.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out

    x = np.concatenate([np.exp(rng.uniform(-700, 700, 50000)), rng.uniform(0.5, 2, 50000), [2.0, 0.5, 1e-310]])
    got, want = call(0, x), np.log(x)
    assert np.max(np.abs(got - want) / np.spacing(np.abs(want))) <= 1.0
    assert call(0, np.array([1.0]))[0] == 0.0 and call(0, np.array([0.0]))[0] == -np.inf
    x = np.concatenate([rng.uniform(-745, 709, 50000), rng.uniform(-1, 1, 50000)])
    got, want = call(1, x), np.exp(x)
    assert np.max(np.abs(got - want) / np.spacing(want)) <= 1.0
    x, y = np.exp(rng.uniform(-30, 30, 50000)), rng.uniform(-3, 3, 50000)
    assert np.max(np.abs(call(2, x, y) / np.power(x, y) - 1)) < 2e-14
    assert call(2, np.array([0.0]), np.array([0.92]))[0] == 0.0
