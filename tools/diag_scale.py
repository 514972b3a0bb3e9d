import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
from spectral_b200 import api
from spectral_b200.scenarios import WEIGHTS_FILE, load_fixture, perturbed_obstacles
B = 8192
batch = perturbed_obstacles(load_fixture("c2"), B, seed=20230601)
p = api.SpectralPlanner(device=0, max_batch=B, n_max=128, r_max=8, k_max=32)
a = p.solve("trp", batch, WEIGHTS_FILE)
c = p.solve("trp", batch, tuple(2.0 * v for v in WEIGHTS_FILE))
sel = a.verified() & c.verified()
d = np.abs(a.ctrl - c.ctrl)
tol = 1e-6 + 1e-5 * np.abs(a.ctrl)
bad = sel & ((d > tol).any(axis=1))
print("verified both", sel.sum(), "bad", bad.sum(), "status0", (a.status == 0).sum())
for b in np.nonzero(bad)[0][:6]:
    K = int(a.K[b]); j = int(np.argmax(d[b] / tol[b]))
    print("b", b, "K", K, "worst idx", j, "seg", (j % (6 * K)) // 6, "axis", j // (6 * K), "vals", a.ctrl[b, j], c.ctrl[b, j], "t", a.segs[b, :K]["t"], "status", a.status[b], c.status[b], "iters", a.iters[b], c.iters[b], "obj", a.obj[b], c.obj[b] / 2)
