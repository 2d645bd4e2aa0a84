#!/bin/bash
# run HERE before every gpurun call: the .so files that travel to the GPU box must be built from the current sources
set -e
cd "$(dirname "$0")/.."
make -C magnetizer_b200/csrc >/dev/null
make -C magnetizer_b200/csrc host >/dev/null 2>&1
make -C oracle >/dev/null
python - <<'PY'
import ctypes, re
L = ctypes.CDLL('magnetizer_b200/libmagnetizer_b200.so')
hdr = re.sub(r"/\*.*?\*/", "", open('include/magnetizer_b200.h').read(), flags=re.S)
missing = [n for n in set(re.findall(r'\b(mag_[a-z0-9_]+)\s*\(', hdr)) if not hasattr(L, n)]
assert not missing, missing
print("libraries up to date")
PY
