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


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def test_device_bessel_bit_exact(solver, oracle):
    """I0, I1, K0, K1: the device twin of oracle/det_math.hpp returns identical bits."""
    xs = np.concatenate([np.geomspace(1e-7, 2, 2000), np.linspace(2, 40, 4000)])
    got, want = np.empty((4, xs.size)), np.empty((4, xs.size))
    assert solver._lib.mag_test_bessel(_vp(xs), C.c_int(xs.size), _vp(got)) == 0
    oracle.lib().oracle_bessel_all(C.c_int32(xs.size), _vp(xs), _vp(want))
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_device_det_math_bit_exact(solver, oracle):
    """log, exp, pow on 3e5 arguments each: bit-identical on the device and in the oracle
    (the Brent decisions of the profile solves depend on the last bit, oracle/det_math.hpp)."""
    rng = np.random.default_rng(11)
    n = 300000
    cases = [
        (0, np.concatenate([np.exp(rng.uniform(-700, 700, n - 4)), [1.0, 0.5, 1e-310, 2.0]]), None),
        (1, np.concatenate([rng.uniform(-745, 709, n // 2), rng.uniform(-2, 2, n // 2)]), None),
        (2, np.exp(rng.uniform(-40, 40, n)), rng.uniform(-3, 3, n)),
    ]
    for which, x, y in cases:
        x = np.ascontiguousarray(x)
        y = np.ascontiguousarray(y if y is not None else x)
        got, want = np.empty_like(x), np.empty_like(x)
        assert solver._lib.mag_test_det_math(C.c_int32(which), C.c_int32(x.size), _vp(x), _vp(y), _vp(got)) == 0
        oracle.lib().oracle_det_math(C.c_int32(which), C.c_int32(x.size), _vp(x), _vp(y), _vp(want))
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), which


def test_device_fast_sqrt(solver):
    """sqrt|alpha| of the hot loop (rsqrt seed + one Newton step): relative error < 2e-12."""
    rng = np.random.default_rng(3)
    x = np.concatenate([np.exp(rng.uniform(-60, 60, 200000)), -np.exp(rng.uniform(-20, 20, 1000)), [1.0, 4.0, 1e-300]])
    got = np.empty_like(x)
    assert solver._lib.mag_test_fast_sqrt(C.c_int32(x.size), _vp(x), _vp(got)) == 0
    err = np.abs(got / np.sqrt(np.abs(x)) - 1)
    print("fast_sqrt max rel err", err.max())
    assert err.max() < 2e-12


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
    print("fast-path blocks replayed:", solver.info()["fast_path_replays"])


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
    prof, scal = ("Br", "Bp", "alp_m", "h", "n", "Omega", "Shear", "alp"), ("Bmax", "Bmax_idx", "dt")
    # 2000..: plain slice; 1500..: contains a galaxy whose hydrostatic solve fails (status 'H',
    # construct_profiles -> finish(), dynamo.f90:134-137) and secant-fallback cases
    for first, n in ((2000, 48), (1500, 32)):
        syn = synthetic_histories(inputs, n, first=first)
        ids = galaxy_ids(n, first=first)
        gpu = solver.run(ids, t, syn, profiles=prof, scalars=scal)
        ora = oracle.solve_batch(solver.params, ids, t, syn, profiles=prof, scalars=scal)
        compare(gpu, ora)
        if first == 1500:
            assert b"H" in ora["status"].tobytes()


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


def test_partitioned_solve_matches_single_context(solver, example_input):
    """mag_solve_partitioned (static + work-stealing chunks, one host thread per device,
    host-side gather): two contexts on the same GPU must reproduce the single-context
    result bit for bit, whatever thread ends up with which chunk."""
    from magnetizer_b200.solver import solve_partitioned
    t, inputs = example_input
    ids, sub = _subset(inputs, list(range(30, 71)))
    prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax", "dt")
    one = solver.run(ids, t, sub, profiles=prof, scalars=scal)
    two = solve_partitioned(ids, t, sub, devices=[0, 0], chunk=4, profiles=prof, scalars=scal)
    assert two["chunks_done"].sum() == 11
    # which galaxies run on four-warp teams depends on the batch they are solved in; the two evolve
    # kernels agree to rounding (~1e-13), not bit for bit
    compare(two, one, rtol=1e-10)
    # with the split switched off the result is bit-identical however the catalogue is cut
    os.environ["MAG_HEAVY_MODE"] = "0"
    try:
        from magnetizer_b200.solver import DynamoSolver
        s0 = DynamoSolver(device=0)
        try:
            one = s0.run(ids, t, sub, profiles=prof, scalars=scal)
        finally:
            s0.close()
        two = solve_partitioned(ids, t, sub, devices=[0, 0], chunk=4, profiles=prof, scalars=scal)
    finally:
        del os.environ["MAG_HEAVY_MODE"]
    for k in ("profiles", "scalars", "nsteps", "error", "point_stages", "flags"):
        assert np.array_equal(one[k], two[k]), k
    assert np.array_equal(one["status"], two["status"])


def test_unstable_timestep_blow_up_and_retry_path(oracle, example_input):
    """Courant factors of 0.5 make the explicit scheme unstable: check_f_error trips, the
    retry loop runs (status 'm', nsteps_0 steps, dynamo.f90:222-242) and the run ends with
    status 'i' and a reseeded field (:245-257).  Exercises the checkpoint/replay logic of the
    fast path (the blow-up must be detected at the reference's step and stage, otherwise the
    accumulated time, the RNG stream and the reseeded field differ) and the invalidation of
    the rows written by the profile phase for the snapshots that are never reached."""
    from magnetizer_b200.solver import DynamoSolver
    t, inputs = example_input
    ids, sub = _subset(inputs, [1, 3, 24, 60, 101, 150])
    p = oracle.default_params()
    p.p_courant_v = 0.5
    p.p_courant_eta = 0.5
    s = DynamoSolver(params=p, device=0)
    try:
        gpu = s.run(ids, t, sub)
        replays = s.info()["fast_path_replays"]
    finally:
        s.close()
    ora = oracle.solve_batch(p, ids, t, sub)
    assert b"i" in ora["status"].tobytes() and (ora["flags"] & 4).any()   # the case really blows up
    compare(gpu, ora)
    assert replays > 0


def test_fine_grid_config5(oracle, example_input):
    """BASELINE config 5: 4x radial resolution (p_nx_ref = 401, nx = 407 and more after grid
    extensions) with tightened Courant factors (0.045, passed as doubles as a namelist file
    would).  Every galaxy runs on the four-warp teams (4 points per thread; 8 beyond 512
    points).  Three example galaxies and 64 of the synthetic catalogue."""
    from magnetizer_b200.solver import DynamoSolver
    from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories
    t, inputs = example_input
    ids, sub = _subset(inputs, [1, 2, 60])
    syn = synthetic_histories(inputs, 64, first=3000)
    ids = np.concatenate([ids, galaxy_ids(64, first=3000)])
    sub = np.ascontiguousarray(np.concatenate([sub, syn], axis=1))
    p = oracle.default_params()
    p.p_nx_ref = 401
    p.p_courant_v = 0.045
    p.p_courant_eta = 0.045
    s = DynamoSolver(params=p, device=0)
    try:
        gpu = s.run(ids, t, sub)
        info = s.info()
    finally:
        s.close()
    assert info["heavy_galaxies"] > 0
    ora = oracle.solve_batch(p, ids, t, sub)
    rep = compare(gpu, ora)
    print("config 5 worst errors:", {k: f"{v:.1e}" for k, v in rep.items() if v > 0}, info)


def test_pinned_outputs_are_written_in_place(solver, example_input):
    """mag_solve_batch writes rows straight into pinned (mapped) host buffers instead of staging
    them in HBM; the result is identical to the staged path with pageable buffers."""
    t, inputs = example_input
    ids, sub = _subset(inputs, [1, 5, 24, 30, 60, 101, 150])
    prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax",)
    staged = solver.run(ids, t, sub, profiles=prof, scalars=scal)
    out = solver.alloc_outputs(len(ids), t.size, prof, scal, pinned=True)
    mapped = solver.run(ids, t, sub, profiles=prof, scalars=scal, out=out)
    for k in ("profiles", "scalars", "status", "nsteps", "error", "point_stages", "flags"):
        assert np.array_equal(np.asarray(staged[k]), np.asarray(mapped[k]), equal_nan=(k in ("profiles", "scalars"))), k


def test_warp_path_every_alignment_matches_generic_path(example_input):
    """The one-warp fast path places the outer Dirichlet point in register (nx-4) % PPL of its
    lane and switches at run time on that position; p_nx_ref = 95..135 walks through every
    position of the 4-, 5- and 6-register variants (and grid growth inside the histories
    reaches the 8-register one).  Fast path and any-nx generic path must agree to rounding."""
    from magnetizer_b200.solver import DynamoSolver, default_params
    t, inputs = example_input
    ids, sub = _subset(inputs, [24, 101])
    prof, scal = ("Br", "Bp", "alp_m", "alp", "Bzmod"), ("Bmax", "Bmax_idx")
    for nref in list(range(95, 107)) + [121, 122, 123, 124, 125, 131, 135, 150, 163]:
        p = default_params()
        p.p_nx_ref = nref
        s = DynamoSolver(params=p, device=0)
        try:
            fast = s.run(ids, t, sub, profiles=prof, scalars=scal)
            assert s.info()["fast_path_replays"] == 0
            s.set_fast_path(False)
            slow = s.run(ids, t, sub, profiles=prof, scalars=scal)
        finally:
            s.close()
        assert fast["status"].tobytes() == slow["status"].tobytes(), nref
        assert np.array_equal(fast["nsteps"], slow["nsteps"]), nref
        rep = compare(fast, slow)
        assert max(rep.values()) < 1e-8, (nref, rep)


# ---------------------------------------------------------------------------------------------
# Every run-time parameter the ABI advertises (mag_params_t), off its default: 32 galaxies of the
# synthetic catalogue + the 196 example galaxies per setting, CUDA path vs oracle.
def _sweep_cases():
    def setter(**kw):
        def f(p):
            for k, v in kw.items():
                setattr(p, k, v)
        return f
    return [
        ("test_run", setter(p_no_magnetic_fields_test_run=1)),                  # dynamo.f90:102-104, status 't'
        ("xoshiro256ss", setter(rng_kind=0)),                                   # GCC >= 9 libgfortran stream
        ("R_kappa_0.5", setter(R_kappa=0.5)),                                   # shared-line loops (mag_rk_fast.cuh)
        ("no_Pdm", setter(p_use_Pdm=0)),
        ("no_Pbulge", setter(p_use_Pbulge=0)),
        ("ignore_satellite_haloes", setter(p_ignore_satellite_DM_haloes=1)),
        ("random_seed_5", setter(p_random_seed=5)),
        ("MAX_FAILS_2_courant_0.4", setter(p_MAX_FAILS=2, p_courant_v=0.4, p_courant_eta=0.4)),   # the retry loop is reached
        ("nsteps_max_5000", setter(p_nsteps_max=5000)),
        ("rmax_over_rdisk_3.5", setter(p_rmax_over_rdisk=3.5)),
        ("floor_and_seed", setter(fmag=0.3, C_floor=2.0, frac_seed=0.05, p_floor_kappa=5.0, C_alp=0.8)),
        ("ism", setter(p_ISM_sound_speed_km_s=12.0, p_ISM_kappa=0.9, p_ISM_xi=0.3, p_ISM_turbulent_length=0.08,
                       p_rreg_to_rdisk=0.15, p_Rmol_alpha=0.8, p_minimum_density=1e-29)),
    ]


@pytest.mark.parametrize("name,setp", _sweep_cases(), ids=[c[0] for c in _sweep_cases()])
def test_parameter_sweep(oracle, example_input, name, setp):
    from magnetizer_b200.solver import DynamoSolver
    from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories
    t, inputs = example_input
    n_ex = inputs.shape[1]
    syn = synthetic_histories(inputs, 32, first=5000)
    ids = np.concatenate([np.arange(1, n_ex + 1, dtype=np.int32), galaxy_ids(32, first=5000)])
    allin = np.ascontiguousarray(np.concatenate([inputs, syn], axis=1))
    prof, scal = ("Br", "Bp", "alp_m", "Bzmod", "h", "n", "Omega", "Shear", "alp", "alp_k", "etat", "Uz"), \
                 ("Bmax", "Bmax_idx", "dt", "rmax", "Bavg", "Beavg")
    p = oracle.default_params()
    setp(p)
    s = DynamoSolver(params=p, device=0)
    try:
        gpu = s.run(ids, t, allin, profiles=prof, scalars=scal)
        info = s.info()
    finally:
        s.close()
    ora = oracle.solve_batch(p, ids, t, allin, profiles=prof, scalars=scal)
    # settings that destabilise some discs (no dark-matter pressure, no satellite haloes, large Courant
    # factors) contain runs that FAIL in the reference's own terms; parity.py explains what parity means there
    rep = compare(gpu, ora, failed_runs="relaxed")
    nfail = int(((ora["flags"] & 4) != 0).sum())
    print(name, "worst:", max(rep.values()), "failed runs:", nfail, "statuses:", sorted(set(ora["status"].tobytes().decode())),
          "heavy", info["heavy_galaxies"], "replays", info["fast_path_replays"])
    if name.startswith("MAX_FAILS"):
        assert (ora["flags"] & 4).any()       # retries really happen
    if name == "test_run":
        assert set(ora["status"].tobytes().decode()) <= set("t-pHhP")


@pytest.mark.parametrize("outflow", [1, 2, 3, 4], ids=["vturb", "wind", "superbubble_simple", "superbubble"])
def test_outflow_models(oracle, example_input, outflow):
    """p_outflow_type other than 'no_outflow' (outflow.f90:68-119): Uz enters the vertical-advection
    terms of all three equations, the floor-field source and the critical dynamo number."""
    from magnetizer_b200.solver import DynamoSolver
    from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories
    t, inputs = example_input
    ex = np.arange(1, 65, dtype=np.int32)
    syn = synthetic_histories(inputs, 32, first=7000)
    ids = np.concatenate([ex, galaxy_ids(32, first=7000)])
    allin = np.ascontiguousarray(np.concatenate([inputs[:, ex - 1, :], syn], axis=1))
    prof, scal = ("Br", "Bp", "alp_m", "Bzmod", "h", "n", "alp", "Uz"), ("Bmax", "Bmax_idx", "dt")
    p = oracle.default_params()
    p.outflow_type = outflow
    s = DynamoSolver(params=p, device=0)
    try:
        gpu = s.run(ids, t, allin, profiles=prof, scalars=scal)
    finally:
        s.close()
    ora = oracle.solve_batch(p, ids, t, allin, profiles=prof, scalars=scal)
    rep = compare(gpu, ora)
    uz = ora["profiles"][prof.index("Uz")]
    assert (uz[uz > -9e4] > 0).any()          # the outflow is really on
    print("outflow", outflow, "worst:", max(rep.values()))


def test_algebraic_quenching(oracle, example_input):
    """Alg_quench = T (gutsdynamo.f90:202-203): alp = alp_k/(1 + (Br^2+Bp^2+Bzmod^2)/Beq^2) replaces the dynamical
    quenching in the RHS; Bzmod is re-estimated at every step (dynamo.f90:179).  Runs on the any-nx generic path
    of both evolve kernels."""
    from magnetizer_b200.solver import DynamoSolver
    t, inputs = example_input
    ids, sub = _subset(inputs, [1, 3, 24, 30, 60, 85, 101, 150])
    p = oracle.default_params()
    p.Alg_quench = 1
    ora = oracle.solve_batch(p, ids, t, sub)
    ref = oracle.solve_batch(oracle.default_params(), ids, t, sub)
    assert not np.array_equal(ora["profiles"][0], ref["profiles"][0])      # the switch changes Br
    s = DynamoSolver(params=p, device=0)
    try:
        for mode in (0, 2):                                                # one-warp teams, four-warp teams
            s.set_heavy_policy(mode=mode)
            rep = compare(s.run(ids, t, sub), ora)
            print("Alg_quench heavy mode", mode, "worst:", max(rep.values()))
    finally:
        s.close()


_SWITCH_REF = {}
_SWITCH_CASES = [
    # (id, settings, runs on the register loops?)
    ("neumann_outer_bc", dict(p_neumann_boundary_condition_rmax=1), False),       # gutsdynamo.f90:63-74
    ("fixed_physical_grid", dict(p_use_fixed_physical_grid=1), True),              # grid.f90:69-73, :159-162
    ("seed_fraction", dict(seed_choice=1), True),                                  # seed_field.f90:38-41
    ("seed_decaying", dict(seed_choice=2, p_r_seed_decay=8.0, p_nn_seed=3), True),  # seed_field.f90:42-46
    ("alg_quench_without_dyn_quench", dict(Alg_quench=1, Dyn_quench=0), False),    # nvar = 2, gutsdynamo.f90:28-43, :199-209
    ("kinematic_no_quenching", dict(Dyn_quench=0), False),
]


@pytest.mark.parametrize("name,settings,fast", _SWITCH_CASES, ids=[c[0] for c in _SWITCH_CASES])
def test_more_physics_switches(oracle, example_input, name, settings, fast):
    """the Neumann outer boundary, the fixed physical grid, the deterministic seed prescriptions and
    the runs without the alp_m equation (nvar = 2: no 'alp_m' rows) -- 32 example + 9 synthetic galaxies each,
    on one-warp and on four-warp teams (validated on 48 + 16, profiles/r2_parity.md; every galaxy is solved
    independently of the others in both forced modes)."""
    from magnetizer_b200.solver import DynamoSolver
    from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories
    t, inputs = example_input
    ex = np.arange(1, 33, dtype=np.int32)
    syn = synthetic_histories(inputs, 9, first=9000)     # the ninth one reaches profiles.f90:155 with the fixed grid
    ids = np.concatenate([ex, galaxy_ids(9, first=9000)])
    allin = np.ascontiguousarray(np.concatenate([inputs[:, ex - 1, :], syn], axis=1))
    prof, scal = ("Br", "Bp", "alp_m", "Bzmod", "h", "n", "alp", "rkpc"), ("Bmax", "Bmax_idx", "dt", "rmax")
    p = oracle.default_params()
    for k, v in settings.items():
        setattr(p, k, v)
    ora = oracle.solve_batch(p, ids, t, allin, profiles=prof, scalars=scal)
    if "ref" not in _SWITCH_REF:
        _SWITCH_REF["ref"] = oracle.solve_batch(oracle.default_params(), ids, t, allin, profiles=prof, scalars=scal)
    assert not np.array_equal(ora["profiles"][0], _SWITCH_REF["ref"]["profiles"][0])   # the switch changes Br
    am = ora["profiles"][prof.index("alp_m")]
    assert (am > -9e4).any() == bool(p.Dyn_quench)                                 # ts_arrays.f90:132-138
    if name == "fixed_physical_grid":                                              # one radial grid for the whole history
        rk = ora["profiles"][prof.index("rkpc")]
        for g in range(0, len(ids), 7):
            rows = rk[g][rk[g][:, -1] > -9e4]
            assert len(rows) == 0 or np.ptp(rows[:, -1]) == 0.0
        st8 = ora["status"][-1].tobytes().decode().rstrip("-")
        assert st8.endswith("MH")                          # a negative scale height extrapolated to the centre (profiles.f90:155)
    s = DynamoSolver(params=p, device=0)
    try:
        for mode in (0, 2):
            s.set_heavy_policy(mode=mode)
            gpu = s.run(ids, t, allin, profiles=prof, scalars=scal)
            rep = compare(gpu, ora, failed_runs="relaxed")
            print(name, "heavy mode", mode, "worst:", max(rep.values()), "statuses:", sorted(set(ora["status"].tobytes().decode())),
                  "replays", s.info()["fast_path_replays"])
    finally:
        s.close()


def test_heavy_teams_every_grid_class(oracle, example_input):
    """k_evolve_heavy (four warps per galaxy, mag_rk_fast.cuh): all galaxies forced onto it.  The
    example histories reach nx = 107 .. 330, i.e. the 1-, 2- and 3-points-per-thread loops; results
    must match the oracle and, bit for bit in the exactly-compared fields, the one-warp kernel."""
    from magnetizer_b200.solver import DynamoSolver
    t, inputs = example_input
    ids, sub = _subset(inputs, [1, 5, 23, 24, 30, 60, 67, 85, 101, 150])
    s = DynamoSolver(device=0)
    try:
        s.set_heavy_policy(mode=2)
        heavy = s.run(ids, t, sub)
        assert s.info()["heavy_galaxies"] == len(ids)
        s.set_heavy_policy(mode=0)
        light = s.run(ids, t, sub)
        assert s.info()["heavy_galaxies"] == 0
    finally:
        s.close()
    ora = oracle.solve_batch(s.params, ids, t, sub)
    compare(heavy, ora)
    compare(light, ora)
    compare(heavy, light)


def test_heavy_light_split_and_launch_modes(oracle, example_input):
    """The cost-based split of a batch into the two evolve kernels, launched on two streams and as
    programmatic dependents (MAG_EVOLVE_PDL=1): identical results, some galaxies in each class."""
    from magnetizer_b200.solver import DynamoSolver
    t, inputs = example_input
    ids = np.arange(1, inputs.shape[1] + 1, dtype=np.int32)
    prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax", "dt")
    res = {}
    for pdl in ("0", "1"):
        os.environ["MAG_EVOLVE_PDL"] = pdl
        try:
            s = DynamoSolver(device=0)
        finally:
            del os.environ["MAG_EVOLVE_PDL"]
        try:
            s.set_heavy_policy(mode=1, alpha=4.0)      # small batch: put the threshold where both classes are populated
            res[pdl] = s.run(ids, t, inputs, profiles=prof, scalars=scal)
            nh = s.info()["heavy_galaxies"]
            assert 0 < nh < len(ids), nh
        finally:
            s.close()
    for k in ("profiles", "scalars", "nsteps", "error", "point_stages", "flags"):   # same split, same kernels: bit-identical
        assert np.array_equal(res["0"][k], res["1"][k]), k
    ora = oracle.solve_batch(oracle.default_params(), ids, t, inputs, profiles=prof, scalars=scal)
    compare(res["1"], ora)


def test_multi_chunk_records(example_input):
    """A batch that does not fit the record tables is solved in several chunks (MAG_MAX_CHUNK caps
    the chunk size here): every chunk addresses its chunk-local tables; same result as one chunk."""
    from magnetizer_b200.solver import DynamoSolver
    t, inputs = example_input
    ids, sub = _subset(inputs, list(range(100, 131)))
    prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax", "dt")
    s = DynamoSolver(device=0)
    try:
        one = s.run(ids, t, sub, profiles=prof, scalars=scal)
    finally:
        s.close()
    os.environ["MAG_MAX_CHUNK"] = "7"
    try:
        s = DynamoSolver(device=0)
        try:
            many = s.run(ids, t, sub, profiles=prof, scalars=scal)
            assert s.last_timing()["kernel_launches"] >= 5 * 8   # 5 chunks
        finally:
            s.close()
    finally:
        del os.environ["MAG_MAX_CHUNK"]
    for k in ("profiles", "scalars", "nsteps", "error", "point_stages", "flags"):
        assert np.array_equal(one[k], many[k]), k
    assert np.array_equal(one["status"], many["status"])


def test_arena_overflow_retry_and_device_status(example_input):
    """A record arena that is too small (MAG_ARENA_NX_PCT=20: far fewer points per snapshot than the
    default grid has) overflows on the device; mag_solve_batch repeats the solve with a larger one,
    mag_solve_batch_device reports it through mag_last_device_status."""
    import torch
    from magnetizer_b200.solver import DynamoSolver, MagnetizerError
    t, inputs = example_input
    ids, sub = _subset(inputs, list(range(1, 41)))
    prof, scal = ("Br", "h"), ("Bmax",)
    s = DynamoSolver(device=0)
    try:
        want = s.run(ids, t, sub, profiles=prof, scalars=scal)
    finally:
        s.close()
    os.environ["MAG_ARENA_NX_PCT"] = "20"
    try:
        s = DynamoSolver(device=0)
    finally:
        del os.environ["MAG_ARENA_NX_PCT"]
    try:
        dev = torch.device("cuda", 0)
        d_in, d_t, d_ids = torch.from_numpy(sub).to(dev), torch.from_numpy(t).to(dev), torch.from_numpy(ids).to(dev)
        d_prof = torch.empty((2, len(ids), t.size, s.params.p_nx_ref), dtype=torch.float64, device=dev)
        d_err = torch.empty(len(ids), dtype=torch.int32, device=dev)
        s.run_device(len(ids), t.size, d_ids.data_ptr(), d_t.data_ptr(), d_in.data_ptr(), prof, (),
                     dict(profiles=d_prof.data_ptr(), error=d_err.data_ptr()), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        with pytest.raises(MagnetizerError):
            s.device_status()
        got = s.run(ids, t, sub, profiles=prof, scalars=scal)       # retries inside
    finally:
        s.close()
    for k in ("profiles", "scalars", "nsteps", "point_stages"):
        assert np.array_equal(want[k], got[k]), k


def test_partition_pinned_in_place_and_resident(example_input):
    """mag_partition_*: page-locked caller arrays are written in place by the kernels (no gather
    copy); the HBM-resident form leaves every galaxy's rows on the device that solved it."""
    import torch
    from magnetizer_b200.solver import DynamoSolver, Partition, host_free
    t, inputs = example_input
    ids, sub = _subset(inputs, list(range(1, 61)))
    prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax", "dt")
    s = DynamoSolver(device=0)
    try:
        one = s.run(ids, t, sub, profiles=prof, scalars=scal)
    finally:
        s.close()
    part = Partition([0, 0])
    try:
        out = part.alloc_outputs(len(ids), t.size, prof, scal, pinned=True)
        two = part.solve(ids, t, sub, chunk=8, profiles=prof, scalars=scal, out=out)
        assert two["chunks_done"].sum() == 8
        compare({k: (np.asarray(v) if isinstance(v, np.ndarray) else v) for k, v in two.items() if not k.startswith("chunks")}, one, rtol=1e-10)
        # resident form: one full copy of the inputs / outputs per listed device
        dev = torch.device("cuda", 0)
        nd, n, nz, nr = 2, len(ids), t.size, part.params.p_nx_ref
        d_in = [torch.from_numpy(sub).to(dev) for _ in range(nd)]
        d_t = [torch.from_numpy(t).to(dev) for _ in range(nd)]
        d_id = [torch.from_numpy(ids).to(dev) for _ in range(nd)]
        d_prof = [torch.full((len(prof), n, nz, nr), -7.0, dtype=torch.float64, device=dev) for _ in range(nd)]
        d_ps = [torch.zeros(n, dtype=torch.int64, device=dev) for _ in range(nd)]
        owner = np.full(n, -1, np.int32)
        st = part.solve_device(n, nz, [x.data_ptr() for x in d_id], [x.data_ptr() for x in d_t], [x.data_ptr() for x in d_in],
                               prof, (), [dict(profiles=d_prof[d].data_ptr(), point_stages=d_ps[d].data_ptr()) for d in range(nd)],
                               chunk=8, owner=owner)
        torch.cuda.synchronize()
        assert st["chunks_done"].sum() == 8 and (owner >= 0).all()
        got = np.stack([d_prof[owner[g]][:, g].cpu().numpy() for g in range(n)], axis=1)
        from parity import _row_err
        assert np.array_equal(got == -99999.0, one["profiles"] == -99999.0)
        assert max(float(_row_err(got[q], one["profiles"][q]).max()) for q in range(len(prof))) < 1e-10
        for k in ("profiles", "scalars", "status", "nsteps", "error", "point_stages", "flags"):
            host_free(out[k])
    finally:
        part.close()
