"""Host side of the C++ driver (magnetizer_b200/host): the namelist reader and the minimal
HDF5 reader must see the same data as the Python mirror / committed fixture.  The GPU test
runs the driver end to end on the reference's example input (BASELINE configs 1 and 2)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "magnetizer_b200", "host", "Magnetizer_b200")
PARAMS = os.path.join("tests", "golden", "example_global_parameters.in")


@pytest.fixture(scope="module")
def exe():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "magnetizer_b200", "csrc"), "host"], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    return EXE


def test_dry_run_reads_namelist_and_hdf5(exe, example_input):
    t, inputs = example_input
    out = subprocess.check_output([exe, PARAMS, "--dry-run"], cwd=ROOT, text=True)
    m = re.search(r"dry-run ngal (\d+) nz (\d+) list (\d+) p_courant_v (\S+) p_nx_ref (\d+) p_random_seed (\d+)", out)
    assert m and (int(m[1]), int(m[2]), int(m[3])) == (196, 40, 196)
    assert float(m[4]) == 0.09000000357627869 and int(m[5]) == 101 and int(m[6]) == 17  # float-rounded default kept
    from magnetizer_b200.hdf5_min import INPUT_COLUMNS
    for c, name in enumerate(INPUT_COLUMNS):
        got = float(re.search(rf"column {name} sum (\S+)", out)[1])
        want = float(np.sum(inputs[c].ravel()))  # same summation order: galaxy-major, snapshot fastest
        assert abs(got - want) <= 1e-9 * abs(want) + 1e-300, name
    assert "profiles rkpc Omega Shear etat alp alp_k n h l Br Bp Bzmod Uz" in out
    assert "scalars Bmax Bmax_idx rmax Bavg Beavg" in out


def test_single_galaxy_list_and_namelist_overrides(exe, tmp_path):
    p = tmp_path / "p.in"
    p.write_text("&run_parameters\n p_random_seed = 5 ! comment\n p_courant_v = 0.09\n/\n&io_parameters\n"
                 f" input_file_name = '{os.path.join(ROOT, 'tests', 'golden', 'example_input.hdf5')}'\n/\n"
                 "&grid_parameters\n p_nx_ref = 51\n/\n")
    out = subprocess.check_output([exe, str(p), "3", "7", "--dry-run"], cwd=ROOT, text=True)
    m = re.search(r"list (\d+) p_courant_v (\S+) p_nx_ref (\d+) p_random_seed (\d+)", out)
    # a value given in the file is parsed as a double (0.09, not the float-rounded default)
    assert (int(m[1]), float(m[2]), int(m[3]), int(m[4])) == (2, 0.09, 51, 5)


def test_resume_filtering_from_log_completed(exe, tmp_path):
    """The output file doubles as checkpoint (distributor.f90:150-156): galaxies whose Log/completed flag
    is set are skipped (IO_hdf5.f90:244-266); --restart ignores the file, --stop-after N is the graceful
    stop.  Host logic only (--dry-run); the file is written by the HDF5 writer probe the way a killed run
    leaves it: galaxies 1-50 and 101-150 completed, the rest never written (fill value)."""
    probe, out = str(tmp_path / "probe"), str(tmp_path / "out.hdf5")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-o", probe, os.path.join(ROOT, "tests", "hdf5_writer_probe.cpp"), "-lz"])
    subprocess.check_call([probe, "write1", out, "196", "101", "40", "50", "chunked"], stdout=subprocess.DEVNULL)
    run = lambda *extra: subprocess.run([exe, PARAMS, "--out", out, "--dry-run", *extra], cwd=ROOT, capture_output=True, text=True)
    assert "resume yes: 100 of 196 galaxies completed, 96 to solve now" in run().stdout
    assert "resume yes: 100 of 196 galaxies completed, 2 to solve now" in run("--stop-after", "2").stdout
    assert "resume no: 0 of 196 galaxies completed, 196 to solve now" in run("--restart").stdout
    assert "resume yes: 2 of 4 galaxies completed, 2 to solve now" in run("50", "51", "100", "101").stdout   # a sub-list
    # a file written for another grid cannot be extended
    p = tmp_path / "p.in"
    p.write_text(open(os.path.join(ROOT, PARAMS)).read().replace("&grid_parameters", "&grid_parameters\n  p_nx_ref = 51"))
    other = subprocess.run([exe, str(p), "--out", out, "--dry-run"], cwd=ROOT, capture_output=True, text=True)
    assert other.returncode == 1 and "cannot resume" in other.stderr and "--restart" in other.stderr


def test_unimplemented_switch_is_rejected_not_ignored(exe, tmp_path):
    """A namelist that sets a physics switch the CUDA path does not implement must fail loudly: running
    default physics under a non-default label would stamp wrong parameters into the output file."""
    base = open(os.path.join(ROOT, PARAMS)).read()
    for key, val in (("Krause", ".false."), ("p_scale_back_f_array", ".false."), ("Alp_squared", "T"), ("p_sech2_profile", ".true.")):
        p = tmp_path / "bad.in"
        p.write_text(base + f"\n&dynamo_parameters\n  {key} = {val}\n/\n")
        r = subprocess.run([exe, str(p), "--dry-run"], cwd=ROOT, capture_output=True, text=True)
        assert r.returncode == 1 and key.lower() in r.stderr.lower() and "not implemented" in r.stderr, key
    bad = tmp_path / "bad2.in"   # a seed prescription the path does not have: rejected by mag_create, not mapped to another one
    bad.write_text(base + "\n&dynamo_parameters\n  p_seed_choice = 'floor'\n/\n")
    r = subprocess.run([exe, str(bad), "--dry-run"], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 1 and "not implemented" in r.stderr, r.stdout + r.stderr
    ok = tmp_path / "ok.in"    # the default value spelled out, implemented switches, an observables key and an unknown key: accepted
    ok.write_text(base + "\n&dynamo_parameters\n  Krause = .true.\n  p_no_such_key = 3\n  p_neumann_boundary_condition_rmax = .true.\n"
                         "  p_seed_choice = 'decaying'\n  p_r_seed_decay = 12d0\n  p_nn_seed = 3\n  Dyn_quench = F\n  Alg_quench = T\n/\n"
                         "&grid_parameters\n  p_use_fixed_physical_grid = .true.\n/\n&observables_parameters\n  p_obs_CR_alpha = 2d0\n/\n")
    r = subprocess.run([exe, str(ok), "--dry-run"], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0 and "unknown key p_no_such_key" in r.stderr and "p_obs_cr_alpha has no effect" in r.stderr


@pytest.fixture(scope="module")
def stub_exe():
    """The real driver linked against the CPU fake of the single-GPU entry points (tests/partition_stub.cpp +
    the real csrc/mag_partition.cu): everything of the driver except the physics runs without a GPU."""
    from test_partition_host import BUILD, build_stub
    build_stub()
    out = os.path.join(BUILD, "Magnetizer_stub")
    src = os.path.join(ROOT, "magnetizer_b200", "host", "magnetizer_main.cpp")
    deps = [src] + [os.path.join(ROOT, "magnetizer_b200", "host", h) for h in ("mag_hdf5_writer.hpp", "mag_hdf5_reader.hpp", "mag_namelist.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps + [os.path.join(BUILD, "libpartstub.so")]):
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O1", "-std=c++17", "-o", out, src, "-L" + BUILD, "-lpartstub", "-lz",
                               "-Xlinker", "-rpath", "-Xlinker", BUILD], stderr=subprocess.DEVNULL)
    return out


def stub_profile(name_id, ids, inputs, nz, fac=1.0):
    """what tests/partition_stub.cpp writes for profile quantity name_id, in the file's (ngal, nx, nz) layout and units"""
    i = np.arange(101) * 1e-3
    v = ids[:, None, None] * 1000.0 + name_id * 10.0 + inputs[0][ids - 1][:, :, None] + i[None, None, :]   # (n, nz, nx)
    return np.transpose(v, (0, 2, 1)) * fac


def test_driver_cycles_layout_resume_and_stop_file(stub_exe, example_input, tmp_path):
    """The driver's side of jobs_distribute + write_output (distributor.f90:253-337, output.f90:23-236) with the
    physics faked: checkpoint cycles with a flush each, the reference's file layout (one row per galaxy of the input
    file, chunked + deflate, fill -1d50), sparse galaxy lists, a stop after a cycle (STOP file) and the in-place resume
    that merges new galaxies into blocks an earlier invocation had partly filled."""
    import h5spec
    from magnetizer_b200.abi import PROFILE_ID
    t, inputs = example_input
    FILL = -1e50
    base = open(os.path.join(ROOT, PARAMS)).read().replace('"tests/golden/', '"' + ROOT + '/tests/golden/').replace("'tests/golden/", "'" + ROOT + "/tests/golden/").replace(
        "&io_parameters", "&io_parameters\n  p_ncheckpoint = 60\n  p_IO_number_of_galaxies_in_chunks = 50")
    params = tmp_path / "cycles.in"
    params.write_text(base)
    out = str(tmp_path / "stub_output.hdf5")
    run = lambda *a: subprocess.run([stub_exe, str(params), "--devices", "0,1", "--out", out, *a], cwd=tmp_path, capture_output=True, text=True)
    # 1. a sparse list first: galaxies 3, 24, 60, 101, 150 -> one cycle (fewer than p_ncheckpoint galaxies), blocks 0, 1, 2
    r = run("3", "24", "60", "101", "150")
    assert r.returncode == 0 and "1 flush(es)" in r.stdout, r.stderr
    objs = h5spec.walk(out)
    ids = np.array([3, 24, 60, 101, 150])
    br = objs["/Output/Br"]
    assert br.shape == (196, 101, 40) and br.chunk == (50, 101, 1) and br.filters == [(1, [6])] and br.nchunks == 3 * 40
    assert np.array_equal(br.data[ids - 1], stub_profile(PROFILE_ID["Br"], ids, inputs, 40))
    rest = np.setdiff1d(np.arange(196), ids - 1)
    assert np.all(br.data[rest] == FILL)                                       # never written: the fill value (IO_hdf5.f90:44,979)
    assert np.array_equal(objs["/Output/h"].data[ids - 1], stub_profile(PROFILE_ID["h"], ids, inputs, 40, 1e3))   # kpc -> pc (output.f90)
    assert np.array_equal(objs["/Output/r"].data[ids - 1], stub_profile(PROFILE_ID["rkpc"], ids, inputs, 40))
    comp = objs["/Log/completed"].data[:, 0]
    assert np.all(comp[ids - 1] == 1.0) and np.all(comp[rest] == FILL)
    assert objs["/Log/status"].data[2].tobytes() == b"M" * 40 and objs["/Log/status"].data[0].tobytes() == b"\0" * 40
    assert np.array_equal(objs["/Log/nsteps"].data[ids - 1, 0], ids.astype(float))
    # 2. the whole catalogue, stopped by a STOP file after its first cycle (blocks 0-1: the galaxies not yet completed there)
    (tmp_path / "STOP").write_text("")
    r = run()
    assert r.returncode == 0 and "STOP file found" in r.stdout and "(5 already completed, 191 to solve)" in r.stdout
    os.remove(tmp_path / "STOP")
    comp = h5spec.walk(out)["/Log/completed"].data[:, 0]
    assert np.all(comp[:100] == 1.0) and comp[100] == 1.0 and comp[149] == 1.0 and np.all(np.delete(comp[101:], 48) == FILL)
    # 3. resume: the rest (blocks 2 and 3 in one cycle); rows of all three invocations are there, the file grew in place
    size2 = os.path.getsize(out)
    r = run()
    assert r.returncode == 0 and "(102 already completed, 94 to solve)" in r.stdout and "1 flush(es)" in r.stdout
    assert os.path.getsize(out) > size2
    objs = h5spec.walk(out)
    assert objs["trailing_bytes"] == 0
    all_ids = np.arange(1, 197)
    assert np.array_equal(objs["/Output/Br"].data, stub_profile(PROFILE_ID["Br"], all_ids, inputs, 40))
    assert np.all(objs["/Log/completed"].data == 1.0) and objs["/Output/Br"].nchunks == 4 * 40
    from magnetizer_b200.hdf5_min import H5File
    assert np.array_equal(H5File(out)["Output/Bmax"], objs["/Output/Bmax"].data)
    # 4. nothing left to do
    r = run()
    assert r.returncode == 0 and "(196 already completed, 0 to solve)" in r.stdout


def test_no_cpu_fallback(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe, PARAMS, "1"], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_driver_end_to_end_example(exe, oracle, example_input, tmp_path):
    """BASELINE configs 1/2 through the real file formats: the reference's parameter file and HDF5 input
    in, the reference's HDF5 output layout out (chunked + deflate, one row per galaxy of the input file,
    fill value -1d50 where nothing was written), several checkpoint cycles, single-galaxy mode, and an
    interrupted run that is extended in place."""
    import h5spec
    from magnetizer_b200.hdf5_min import H5File
    from parity import RTOL
    t, inputs = example_input
    FILL = -1e50
    base = open(os.path.join(ROOT, PARAMS)).read().replace(
        "&io_parameters", "&io_parameters\n  p_ncheckpoint = 60\n  p_IO_number_of_galaxies_in_chunks = 50")
    params = tmp_path / "cycles.in"
    params.write_text(base)
    out = str(tmp_path / "example_output.hdf5")
    log = subprocess.check_output([exe, str(params), "--devices", "0,0", "--out", out], cwd=ROOT, text=True)
    assert "196 galaxies x 40 snapshots" in log and "2 flush(es)" in log       # cycles of 100 + 96 galaxies (blocks of 50)
    objs = h5spec.walk(out)                                                    # the format-specification walker
    assert objs["trailing_bytes"] == 0
    br = objs["/Output/Br"]
    assert br.shape == (196, 101, 40) and br.chunk == (50, 101, 1) and br.filters == [(1, [6])] and br.nchunks == 4 * 40
    d = H5File(out)
    ids = np.arange(1, 197, dtype=np.int32)
    ora = oracle.solve_batch(oracle.default_params(), ids, t, inputs, profiles=("Br", "h", "rkpc"), scalars=("Bmax",))
    for k, name, fac in ((0, "Br", 1.0), (1, "h", 1e3), (2, "r", 1.0)):
        want = np.transpose(ora["profiles"][k], (0, 2, 1))
        want = np.where(want == -99999.0, want, want * fac)
        got = d["Output/" + name]
        assert np.array_equal(got, objs["/Output/" + name].data)
        scale = np.abs(want).max(axis=1, keepdims=True)
        assert np.all(np.abs(got - want) <= RTOL * np.maximum(scale, 1e-300)), name
    assert np.array_equal(d["Log/status"].view("S1").reshape(196, 40), ora["status"])
    assert np.array_equal(d["Log/nsteps"].ravel(), ora["nsteps"].astype(float))
    assert d["Log/nsteps"].shape == (196, 1) and d["Log/completed"].ravel().tolist() == [1.0] * 196
    assert d.attrs("Output/Br")["Units"] == "microgauss" and d.attrs("Output/h")["Description"] == "Disk scaleheight"
    # single-galaxy mode (Fortran gal_id = 24): the datasets still have one row per galaxy of the input file
    out1 = str(tmp_path / "example_output_gal24.hdf5")
    subprocess.check_call([exe, PARAMS, "24", "--out", out1], cwd=ROOT, stdout=subprocess.DEVNULL)
    f1 = H5File(out1)
    assert f1.keys("/") == ["Log", "Output"] and "Br" in f1.keys("Output") and "r" in f1.keys("Output")
    b1 = f1["Output/Br"]
    assert b1.shape == (196, 101, 40) and np.array_equal(b1[23], d["Output/Br"][23])
    assert np.all(b1[:23] == FILL) and np.all(b1[24:] == FILL)                 # IO_hdf5.f90:44,979
    comp = f1["Log/completed"].ravel()
    assert comp[23] == 1.0 and np.all(np.delete(comp, 23) == FILL)
    assert f1["Log/status"][23].tobytes() == ora["status"][23].tobytes() and f1["Log/status"][0].tobytes() == b"\0" * 40
    # interrupted run + resume == straight run (the output file is the checkpoint and is extended in place)
    part, full = str(tmp_path / "resumed.hdf5"), str(tmp_path / "straight.hdf5")
    gl = ["3", "24", "60", "101"]
    log_a = subprocess.check_output([exe, PARAMS, *gl, "--out", part, "--stop-after", "2"], cwd=ROOT, text=True)
    fa = H5File(part)
    ca = fa["Log/completed"].ravel()
    assert "(0 already completed, 2 to solve)" in log_a and ca[[2, 23]].tolist() == [1.0, 1.0] and ca[59] == FILL and ca[100] == FILL
    size_a = os.path.getsize(part)
    log_b = subprocess.check_output([exe, PARAMS, *gl, "--out", part], cwd=ROOT, text=True)
    assert "(2 already completed, 2 to solve)" in log_b and os.path.getsize(part) > size_a
    subprocess.check_call([exe, PARAMS, *gl, "--out", full], cwd=ROOT, stdout=subprocess.DEVNULL)
    h5spec.walk(part)
    fb, fc = H5File(part), H5File(full)
    assert fb["Log/completed"].ravel()[[2, 23, 59, 100]].tolist() == [1.0] * 4 and fb.keys("Output") == fc.keys("Output")
    for k in fc.keys("Output"):
        assert np.array_equal(fb["Output/" + k], fc["Output/" + k]), k
    assert fb["Log/status"].tobytes() == fc["Log/status"].tobytes() and np.array_equal(fb["Log/nsteps"], fc["Log/nsteps"])
    root = f1.attrs("/")
    assert "P_NX_REF=101," in root["grid_parameters"] and root["run_parameters"].startswith("&RUN_PARAMETERS ")
    assert float(re.search(r"P_COURANT_V=([^,]+),", root["run_parameters"])[1]) == 0.09000000357627869
    assert 'P_OUTFLOW_TYPE="no_outflow"' in root["outflow_parameters"] and "P_IO_COMPRESSION=T" in root["io_parameters"]
