"""Native HDF5 writer of the host driver (magnetizer_b200/host/mag_hdf5_writer.hpp, SURVEY.md
8(f1)): a file written by it must read back identically through the two independent readers
of this repo -- the pure-python one (magnetizer_b200/hdf5_min.py, which also parses the
reference's own example_input.hdf5) and the C++ one of the driver -- and must carry what the
reference's consumer (python/magnetizer/MagnetizerRun.py:117-150, parameters.py:65-125)
looks for: /Output and /Log groups, (ngals, nx, nz) profiles, 'Units' / 'Description'
attributes, the namelists as text attributes of the root group."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def probe_file(tmp_path_factory):
    d = tmp_path_factory.mktemp("h5w")
    exe, out = str(d / "hdf5_writer_probe"), str(d / "probe.hdf5")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "hdf5_writer_probe.cpp"), "-lz"])
    txt = subprocess.check_output([exe, out, "7", "11", "5"], text=True)
    return out, txt


def test_cpp_reader_round_trip(probe_file):
    _, txt = probe_file
    n = 7 * 11 * 5
    assert f"Omega shape 7 11 5 first {6000.0:.17g} last {6000.0 + 0.5 * (n - 1):.17g}" in txt
    assert "runtime n 7 last 6.25" in txt and "gal_id n 7 last 7" in txt


def test_python_reader_round_trip(probe_file):
    from magnetizer_b200.hdf5_min import H5File
    path, _ = probe_file
    f = H5File(path)
    assert f.keys("/") == ["Log", "Output"]
    names = ["r", "Br", "Bp", "alp_m", "Bzmod", "h", "Omega", "Shear", "l", "v", "etat", "tau", "alp_k", "Uz", "n", "Beq", "alp",
             "Omega_h", "Omega_b", "Omega_d", "rkpc"]
    assert f.keys("Output") == sorted(names + ["Bmax"])
    n = 7 * 11 * 5
    for q, nm in enumerate(names):
        a = f["Output/" + nm]
        assert a.shape == (7, 11, 5) and a.dtype == np.float64
        assert np.array_equal(a.ravel(), 1000.0 * q + 0.5 * np.arange(n))
        at = f.attrs("Output/" + nm)
        assert at["Units"] == "microgauss" and at["Description"] == "profile " + nm
    assert np.array_equal(f["Output/Bmax"].ravel(), -99999.0 + np.arange(35))
    st = f["Log/status"]
    assert st.shape == (7, 5) and st.dtype == np.dtype("S1") and st.tobytes() == (b"Mm-i" * 9)[:35]
    assert f["Log/runtime"].shape == (7, 1) and f["Log/runtime"][6, 0] == 6.25
    assert f["Log/gal_id"].dtype == np.int32 and f["Log/gal_id"].tolist() == list(range(1, 8))
    root = f.attrs("/")
    assert root["Model name"] == "probe model"
    # the consumer's own parsing of a namelist attribute (parameters.py:85-118)
    m = re.match(r"(&\w+)\s+(.+)", root["run_parameters"])
    assert m.group(1) == "&RUN_PARAMETERS"
    vals = dict(re.sub(r"\s+", "", v).split("=") for v in m.group(2).split(",") if "=" in v)
    assert float(vals["P_COURANT_V"]) == 0.09000000357627869 and vals["P_VARIABLE_TIMESTEPS"] == "T"


def test_file_structure_matches_the_format_specification(probe_file):
    """Structural facts a libhdf5-based reader relies on (HDF5 File Format Specification,
    version 0 superblock / version 1 B-trees): node sizes derived from the K values of the
    superblock, end-of-file address, sorted symbol tables, 8-byte alignment."""
    path, _ = probe_file
    buf = open(path, "rb").read()
    assert buf[:8] == b"\x89HDF\r\n\x1a\n" and buf[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])
    leaf_k, int_k, flags = struct.unpack_from("<HHI", buf, 16)
    base, free, eof, drv = struct.unpack_from("<QQQQ", buf, 24)
    assert flags == 0 and base == 0 and free == drv == 0xFFFFFFFFFFFFFFFF and eof == len(buf)
    name_off, ohdr, cache, _, btree, heap = struct.unpack_from("<QQIIQQ", buf, 56)
    assert name_off == 0 and cache == 1 and ohdr % 8 == 0 and btree % 8 == 0 and heap % 8 == 0

    def check_group(btree, heap, want):
        assert buf[btree:btree + 4] == b"TREE" and buf[btree + 4] == 0 and buf[btree + 5] == 0
        nent = struct.unpack_from("<H", buf, btree + 6)[0]
        assert nent == 1 and btree + 24 + (2 * int_k + 1) * 8 + 2 * int_k * 8 <= len(buf)
        key0, snod, key1 = struct.unpack_from("<QQQ", buf, btree + 24)
        assert buf[heap:heap + 4] == b"HEAP"
        seg_size, free_head, seg = struct.unpack_from("<QQQ", buf, heap + 8)
        assert seg == heap + 32 and free_head + 16 <= seg_size and struct.unpack_from("<QQ", buf, seg + free_head) == (1, seg_size - free_head)
        assert buf[snod:snod + 4] == b"SNOD" and buf[snod + 4] == 1 and snod + 8 + 2 * leaf_k * 40 <= len(buf)
        nsym = struct.unpack_from("<H", buf, snod + 6)[0]
        names = []
        for i in range(nsym):
            no = struct.unpack_from("<Q", buf, snod + 8 + 40 * i)[0]
            names.append(buf[seg + no:buf.index(b"\0", seg + no)].decode())
        assert names == sorted(want) and key0 == 0
        assert buf[seg + key1:buf.index(b"\0", seg + key1)].decode() == names[-1]   # right key = largest name of the child
        return names

    assert check_group(btree, heap, ["Log", "Output"]) == ["Log", "Output"]
