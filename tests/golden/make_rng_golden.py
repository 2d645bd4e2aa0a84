"""Generates golden RNG vectors from the libgfortran builds vendored inside the
scipy/numpy wheels of this image, driven through ctypes exactly like the reference's
set_random_seed / random_number calls (source/random.f90:39-64, 90-91, 126):

    call random_seed(size=n); seed(:) = w; call random_seed(put=seed)
    call random_number(x)   ! REAL(8)

Run in the build container:  python tests/golden/make_rng_golden.py
Writes tests/golden/rng_golden.json (seed size tells the generator: 33 = xorshift1024*
of GCC 7/8, 8 = xoshiro256** of GCC >= 9).
"""
import ctypes as C
import glob
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


class Dim(C.Structure):
    _fields_ = [("stride", C.c_ssize_t), ("lbound", C.c_ssize_t), ("ubound", C.c_ssize_t)]


class DType(C.Structure):
    _fields_ = [("elem_len", C.c_size_t), ("version", C.c_int), ("rank", C.c_byte), ("type", C.c_byte),
                ("attribute", C.c_short)]


class Desc(C.Structure):  # gfc_array descriptor, GCC >= 8 layout
    _fields_ = [("base", C.c_void_p), ("offset", C.c_ssize_t), ("dtype", DType), ("span", C.c_ssize_t),
                ("dim", Dim * 1)]


def draws(path, word, ndraw):
    for q in glob.glob(os.path.join(os.path.dirname(path), "libquadmath*")):
        try:
            C.CDLL(q, mode=C.RTLD_GLOBAL)
        except OSError:
            pass
    lib = C.CDLL(path)
    n = C.c_int32(0)
    lib._gfortran_random_seed_i4(C.byref(n), None, None)
    arr = (C.c_int32 * n.value)(*([word] * n.value))
    d = Desc()
    d.base = C.cast(arr, C.c_void_p)
    d.offset = -1
    d.dtype = DType(4, 0, 1, 1, 0)
    d.span = 4
    d.dim[0] = Dim(1, 1, n.value)
    lib._gfortran_random_seed_i4(None, C.byref(d), None)
    out = []
    x = C.c_double(0)
    for _ in range(ndraw):
        lib._gfortran_random_r8(C.byref(x))
        out.append(x.value)
    return n.value, out


def wrap32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


if __name__ == "__main__":
    libs = {}
    for path in sorted(glob.glob("/opt/prime-rl/.venv/lib/python3.12/site-packages/*.libs/libgfortran*.so*")):
        libs.setdefault(os.path.basename(path), path)  # the wheels ship duplicate copies
    libs = [libs[sorted(libs)[0]]]  # one process can only bind one libgfortran.so.5; all copies here are GCC-8 era
    res = {}
    # seed words as set_random_seed computes them: p_random_seed**2 - igal**3 (int32 wrap)
    igals = [1, 2, 22, 23, 24, 196, 1291, 5000, 100000]
    for path in libs:
        key = os.path.basename(path)
        entry = {"cases": []}
        for igal in igals:
            w = wrap32(17 * 17 - igal ** 3)
            n, d = draws(path, w, 16)
            entry["seed_size"] = n
            entry["cases"].append({"igal": igal, "p_random_seed": 17, "word": w, "uniforms": [repr(v) for v in d]})
        res[key] = entry
        print(key, "seed size", entry["seed_size"], entry["cases"][0]["uniforms"][:2])
    with open(os.path.join(HERE, "rng_golden.json"), "w") as fh:
        json.dump(res, fh, indent=1)
