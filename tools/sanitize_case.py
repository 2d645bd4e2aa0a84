"""GPU box, under compute-sanitizer: one solve of the 48 example + 16 synthetic galaxies of
tests/test_gpu_parity.py::test_more_physics_switches with the settings given as key=value arguments."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from magnetizer_b200.solver import DynamoSolver, default_params  # noqa: E402
from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories  # noqa: E402

d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
t, inputs = d["t_gyr"], d["inputs"]
ex = np.arange(1, 49, dtype=np.int32)
syn = synthetic_histories(inputs, 16, first=9000)
ids = np.concatenate([ex, galaxy_ids(16, first=9000)])
allin = np.ascontiguousarray(np.concatenate([inputs[:, ex - 1, :], syn], axis=1))
lo, n = int(os.environ.get("MAG_SAN_FIRST", "0")), int(os.environ.get("MAG_SAN_COUNT", "64"))   # a slice of the 64 galaxies
ids, allin = ids[lo:lo + n], np.ascontiguousarray(allin[:, lo:lo + n])
p = default_params()
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    setattr(p, k, type(getattr(p, k))(float(v)) if isinstance(getattr(p, k), float) else int(v))
s = DynamoSolver(params=p, device=0)
s.set_heavy_policy(mode=int(os.environ.get("MAG_SAN_MODE", "0")))
r = s.run(ids, t, allin, profiles=("Br", "Bp", "alp_m", "h", "rkpc"), scalars=("Bmax",))
print("ok", lo, n, sorted(set(r["status"].tobytes().decode())), int(r["point_stages"].sum()))
s.close()
