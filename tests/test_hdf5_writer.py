"""Native HDF5 writer of the host driver (magnetizer_b200/host/mag_hdf5_writer.hpp, SURVEY.md
8(f1)) in the reference's output layout (IO_hdf5.f90:44,463,551,721-731,940-981): datasets sized
to all galaxies of the input file, chunked (gals_per_chunk x all radii x one snapshot) + deflate,
fill value -1d50 for never-written galaxies, flushes that leave a valid file behind, a later
invocation extending the same file.

A file written by it is checked three ways: (1) tests/h5spec.py walks it structure by structure
against the invariants of the HDF5 File Format Specification that libhdf5 relies on -- code that
shares nothing with the repo's readers; (2) the pure-python reader (magnetizer_b200/hdf5_min.py,
which also parses the reference's own example_input.hdf5) and (3) the C++ reader of the driver
must return the same numbers.  It must also carry what the reference's consumer
(python/magnetizer/MagnetizerRun.py:117-150, parameters.py:65-125) looks for: /Output and /Log,
(ngals, nx, nz) profiles, 'Units' / 'Description' attributes, the namelists as root attributes."""
import os
import re
import subprocess

import numpy as np
import pytest

import h5spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["r", "Br", "Bp", "alp_m", "Bzmod", "h", "Omega", "Shear", "l", "v", "etat", "tau", "alp_k", "Uz", "n", "Beq", "alp",
         "Omega_h", "Omega_b", "Omega_d", "rkpc"]
FILL = -1e50


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    d = tmp_path_factory.mktemp("h5w")
    exe = str(d / "hdf5_writer_probe")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "hdf5_writer_probe.cpp"), "-lz"])
    return exe, d


def expected(ngal, nr, nz, blocks_written, cs):
    g = np.arange(ngal)[:, None, None]
    i = np.arange(nr)[None, :, None]
    iz = np.arange(nz)[None, None, :]
    lin = (g * nr + i) * nz + iz
    written = np.zeros(ngal, bool)
    for kb in blocks_written:
        written[kb * cs:(kb + 1) * cs] = True
    prof = {nm: np.where(written[:, None, None], 1000.0 * q + 0.5 * lin, FILL) for q, nm in enumerate(NAMES)}
    gz = np.arange(ngal)[:, None] * nz + np.arange(nz)[None, :]
    bmax = np.where(written[:, None], -99999.0 + gz, FILL)
    status = np.where(written[:, None], np.frombuffer(b"Mm-i", "S1")[gz % 4], b"")
    return written, prof, bmax, status


def check_file(path, ngal, nr, nz, cs, blocks, mode, closed=True):
    written, prof, bmax, status = expected(ngal, nr, nz, blocks, cs)
    # (1) the specification walker
    objs = h5spec.walk(path)
    assert (objs["trailing_bytes"] == 0) == closed
    assert sorted(k for k in objs if k.startswith("/Output/")) == sorted("/Output/" + n for n in NAMES + ["Bmax"])
    for nm in NAMES:
        d = objs["/Output/" + nm]
        assert d.shape == (ngal, nr, nz) and d.dtype == np.float64
        assert d.attrs["Units"] == "microgauss" and d.attrs["Description"] == "profile " + nm
        assert np.array_equal(d.data, prof[nm]), nm
        if mode != "contiguous":
            assert d.layout == "chunked" and d.chunk == (min(cs, ngal), nr, 1)       # IO_hdf5.f90:951-952
            assert d.fill == (8, np.float64(FILL).tobytes())                        # IO_hdf5.f90:979
            assert d.filters == ([(1, [6])] if mode == "chunked" else [])          # deflate, level 6 (:964)
            assert d.nchunks == len(blocks) * nz
    assert np.array_equal(objs["/Output/Bmax"].data, bmax)
    assert objs["/Output/Bmax"].chunk == ((min(cs, ngal), 1) if mode != "contiguous" else None)
    st = objs["/Log/status"]
    assert st.shape == (ngal, nz) and st.dtype == np.dtype("S1") and np.array_equal(st.data, status)
    assert objs["/Log/runtime"].shape == (ngal, 1) and objs["/Log/gal_id"].dtype == np.int32
    assert np.array_equal(objs["/Log/completed"].data[:, 0], np.where(written, 1.0, FILL))
    root = objs["/"]
    assert root["Model name"] == "probe model"
    # the consumer's own parsing of a namelist attribute (parameters.py:85-118)
    m = re.match(r"(&\w+)\s+(.+)", root["run_parameters"])
    assert m.group(1) == "&RUN_PARAMETERS"
    vals = dict(re.sub(r"\s+", "", v).split("=") for v in m.group(2).split(",") if "=" in v)
    assert float(vals["P_COURANT_V"]) == 0.09000000357627869 and vals["P_VARIABLE_TIMESTEPS"] == "T"
    # (2) the pure-python reader of the repo
    from magnetizer_b200.hdf5_min import H5File
    f = H5File(path)
    assert f.keys("/") == ["Log", "Output"]
    for nm in ("Br", "rkpc", "Omega_d"):
        assert np.array_equal(f["Output/" + nm], prof[nm]), nm
    assert np.array_equal(f["Output/Bmax"], bmax)
    assert f.attrs("Output/h")["Units"] == "microgauss"
    return objs


@pytest.mark.parametrize("mode", ["chunked", "raw", "contiguous"])
def test_killed_run_leaves_last_flush_and_a_second_invocation_extends_it(probe, mode):
    exe, d = probe
    path = str(d / f"probe_{mode}.hdf5")
    ngal, nr, nz, cs = 23, 11, 5, 6          # four blocks of galaxies, the last one partial
    out = subprocess.check_output([exe, "write1", path, str(ngal), str(nr), str(nz), str(cs), mode], text=True)
    assert "blocks 4" in out
    # the process died after writing block 3 without a flush: blocks 0 and 2 are there, 1 and 3 read as fill
    # (contiguous storage is rewritten in place, so block 3 is visible there -- like a killed libhdf5 run)
    first = [0, 2] if mode != "contiguous" else [0, 2, 3]
    check_file(path, ngal, nr, nz, cs, first, mode, closed=(mode == "contiguous"))
    size1 = os.path.getsize(path)
    out = subprocess.check_output([exe, "resume", path, str(ngal), str(nr), str(nz), str(cs), mode], text=True)
    assert ("resume: completed 1 0 1 0" if mode != "contiguous" else "resume: completed 1 0 1 1") in out
    objs = check_file(path, ngal, nr, nz, cs, [0, 1, 2, 3], mode)
    assert os.path.getsize(path) > size1                                            # extended in place, not rewritten
    # (3) the C++ reader: a row range that straddles two blocks
    txt = subprocess.check_output([exe, "rows", path, "Output/Omega", "4", "5"], text=True).split("\n")
    assert txt[0] == f"shape 5 {nr} {nz}"
    got = np.array([float(x) for x in txt[1:] if x]).reshape(5, nr, nz)
    assert np.array_equal(got, objs["/Output/Omega"].data[4:9])


def test_many_chunks_build_a_multi_level_btree(probe):
    """More than 64 chunks per dataset need internal B-tree nodes: 150 blocks x 3 snapshots = 450 chunks
    -> 8 leaves under one root; sibling pointers and keys are checked by the walker."""
    exe, d = probe
    path = str(d / "probe_big.hdf5")
    ngal, nr, nz, cs = 300, 3, 3, 2
    subprocess.check_call([exe, "write1", path, str(ngal), str(nr), str(nz), str(cs), "chunked"], stdout=subprocess.DEVNULL)
    subprocess.check_call([exe, "resume", path, str(ngal), str(nr), str(nz), str(cs), "chunked"], stdout=subprocess.DEVNULL)
    objs = h5spec.walk(path)
    dset = objs["/Output/Br"]
    assert dset.nchunks == 150 * 3 and dset.btree_levels == 2
    _, prof, bmax, _ = expected(ngal, nr, nz, range(150), cs)
    assert np.array_equal(dset.data, prof["Br"]) and np.array_equal(objs["/Output/Bmax"].data, bmax)
    from magnetizer_b200.hdf5_min import H5File
    assert np.array_equal(H5File(path)["Output/Br"], prof["Br"])


def test_walker_rejects_a_damaged_file(probe):
    """The specification walker is not a rubber stamp: structural damage makes it fail."""
    exe, d = probe
    path = str(d / "probe_dmg.hdf5")
    subprocess.check_call([exe, "write1", path, "23", "11", "5", "6", "chunked"], stdout=subprocess.DEVNULL)
    subprocess.check_call([exe, "resume", path, "23", "11", "5", "6", "chunked"], stdout=subprocess.DEVNULL)
    buf = bytearray(open(path, "rb").read())
    h5spec.walk(path)
    root_ohdr = int.from_bytes(buf[64:72], "little")
    damages = {
        "truncated file (end-of-file address beyond the file)": lambda b: b[:-8],
        "symbol-table node signature": lambda b: b[:b.rindex(b"SNOD")] + b"SNOX" + b[b.rindex(b"SNOD") + 4:],
        "chunk B-tree node type": lambda b: b[:b.rindex(b"TREE\x01") + 4] + b"\x02" + b[b.rindex(b"TREE\x01") + 5:],
        "chunk keys out of order": lambda b: swap_first_two_chunk_keys(b),
        "object header message count": lambda b: b[:root_ohdr + 2] + bytes([b[root_ohdr + 2] + 1]) + b[root_ohdr + 3:],
        "deflate stream": lambda b: corrupt_first_chunk(b),
    }

    def swap_first_two_chunk_keys(b):
        i = b.rindex(b"TREE\x01\x00")            # a leaf of a chunk B-tree (rank 3: key = 8 + 4*8 bytes, then the child address)
        k = 8 + 8 * 4
        e0, e1 = i + 24, i + 24 + k + 8
        return b[:e0] + b[e1:e1 + k] + b[e0 + k:e1] + b[e0:e0 + k] + b[e1 + k:]

    def corrupt_first_chunk(b):
        i = b.rindex(b"TREE\x01\x00")
        addr = int.from_bytes(b[i + 24 + 40:i + 24 + 48], "little")
        return b[:addr] + bytes(x ^ 0xFF for x in b[addr:addr + 16]) + b[addr + 16:]

    for what, fn in damages.items():
        p2 = str(d / "bad.hdf5")
        open(p2, "wb").write(bytes(fn(bytearray(buf))))
        with pytest.raises(Exception):
            h5spec.walk(p2)
        print("rejected:", what)
