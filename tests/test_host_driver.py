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
    stop.  Host logic only (--dry-run); the file is written by the HDF5 writer probe."""
    probe, out = str(tmp_path / "probe"), str(tmp_path / "out.hdf5")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", probe, os.path.join(ROOT, "tests", "hdf5_writer_probe.cpp"), "-lz"])
    subprocess.check_call([probe, out, "7", "11", "5"], stdout=subprocess.DEVNULL)   # galaxies 1..7, completed: 1, 4, 7
    ids = [str(i) for i in range(1, 8)]
    run = lambda *extra: subprocess.run([exe, PARAMS, *ids, "--out", out, "--dry-run", *extra], cwd=ROOT, capture_output=True, text=True)
    assert "resume yes: 3 of 7 galaxies completed, 4 to solve now" in run().stdout
    assert "resume yes: 3 of 7 galaxies completed, 2 to solve now" in run("--stop-after", "2").stdout
    assert "resume no: 0 of 7 galaxies completed, 7 to solve now" in run("--restart").stdout
    other = subprocess.run([exe, PARAMS, "1", "2", "3", "--out", out, "--dry-run"], cwd=ROOT, capture_output=True, text=True)
    assert other.returncode == 1 and "cannot resume" in other.stderr and "--restart" in other.stderr


def test_no_cpu_fallback(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe, PARAMS, "1"], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_driver_end_to_end_example(exe, oracle, example_input, tmp_path):
    """BASELINE configs 1/2 through the real file format: single-galaxy mode and a full run."""
    from parity import RTOL
    t, inputs = example_input
    out = str(tmp_path / "example_output.npz")
    log = subprocess.check_output([exe, PARAMS, "--devices", "0,0", "--chunk", "16", "--out", out], cwd=ROOT, text=True)
    assert "196 galaxies x 40 snapshots" in log
    d = np.load(out)
    assert d["Output/Br"].shape == (196, 101, 40) and d["Log/status"].shape == (196, 40)
    ids = np.arange(1, 197, dtype=np.int32)
    ora = oracle.solve_batch(oracle.default_params(), ids, t, inputs, profiles=("Br", "h", "rkpc"), scalars=("Bmax",))
    for k, name, fac in ((0, "Br", 1.0), (1, "h", 1e3), (2, "r", 1.0)):
        want = np.transpose(ora["profiles"][k], (0, 2, 1))
        want = np.where(want == -99999.0, want, want * fac)
        got = d["Output/" + name]
        scale = np.abs(want).max(axis=1, keepdims=True)
        assert np.all(np.abs(got - want) <= RTOL * np.maximum(scale, 1e-300)), name
    assert np.array_equal(d["Log/status"].view("S1").reshape(196, 40), ora["status"])
    assert np.array_equal(d["Log/nsteps"], ora["nsteps"].astype(float))
    # single-galaxy mode (Fortran gal_id = 1), written as HDF5 in the reference's layout
    # (mag_hdf5_writer.hpp) and read back with the pure-python reader
    from magnetizer_b200.hdf5_min import H5File
    out1 = str(tmp_path / "example_output_gal1.hdf5")
    subprocess.check_call([exe, PARAMS, "1", "--out", out1], cwd=ROOT, stdout=subprocess.DEVNULL)
    f1 = H5File(out1)
    assert f1.keys("/") == ["Log", "Output"] and "Br" in f1.keys("Output") and "r" in f1.keys("Output")
    assert np.array_equal(f1["Output/Br"][0], d["Output/Br"][0]) and f1["Log/gal_id"].tolist() == [1]
    assert f1["Output/Br"].shape == (1, 101, 40) and f1["Log/status"].shape == (1, 40) and f1["Log/nsteps"].shape == (1, 1)
    assert f1["Log/status"].tobytes() == ora["status"][0].tobytes()
    assert f1.attrs("Output/Br")["Units"] == "microgauss" and f1.attrs("Output/h")["Description"] == "Disk scaleheight"
    # interrupted run + resume == straight run (the output file is the checkpoint)
    part, full = str(tmp_path / "resumed.hdf5"), str(tmp_path / "straight.hdf5")
    gl = ["3", "24", "60", "101"]
    log_a = subprocess.check_output([exe, PARAMS, *gl, "--out", part, "--stop-after", "2"], cwd=ROOT, text=True)
    fa = H5File(part)
    assert "(0 already completed, 2 to solve)" in log_a and fa["Log/completed"].ravel().tolist() == [1.0, 1.0, 0.0, 0.0]
    assert fa["Log/status"][2].tobytes() == b"-" * 40 and np.all(fa["Output/Br"][2:] == -99999.0)
    log_b = subprocess.check_output([exe, PARAMS, *gl, "--out", part], cwd=ROOT, text=True)
    assert "(2 already completed, 2 to solve)" in log_b
    subprocess.check_call([exe, PARAMS, *gl, "--out", full], cwd=ROOT, stdout=subprocess.DEVNULL)
    fb, fc = H5File(part), H5File(full)
    assert fb["Log/completed"].ravel().tolist() == [1.0] * 4 and fb.keys("Output") == fc.keys("Output")
    for k in fc.keys("Output"):
        assert np.array_equal(fb["Output/" + k], fc["Output/" + k]), k
    assert fb["Log/status"].tobytes() == fc["Log/status"].tobytes() and np.array_equal(fb["Log/nsteps"], fc["Log/nsteps"])
    root = f1.attrs("/")
    assert "P_NX_REF=101," in root["grid_parameters"] and root["run_parameters"].startswith("&RUN_PARAMETERS ")
    assert float(re.search(r"P_COURANT_V=([^,]+),", root["run_parameters"])[1]) == 0.09000000357627869
