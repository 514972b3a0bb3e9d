"""Debug driver of the shared-KKT path (K4a): default path vs shared_kkt=1 on small config-3 batches."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spectral_b200 import api
from spectral_b200.scenarios import WEIGHTS_FILE, config3, config2, GOLDEN_W_CUB

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
groups = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pl = api.SpectralPlanner(device=0, max_batch=max(B, 64), n_max=128, r_max=8, k_max=16)
batch = config3(B, groups=groups)
ref = pl.solve("trp", batch, WEIGHTS_FILE)
t0 = time.time()
got = pl.solve("trp", batch, WEIGHTS_FILE, options=api.default_options(shared_kkt=1))
print("shared solve wall %.3f s" % (time.time() - t0))
print("K      ", ref.K[:16])
print("status ", ref.status[:16], "\n shared", got.status[:16])
print("iters  ", ref.iters[:16], "\n shared", got.iters[:16])
print("flags  ", ref.flags[:16], "\n shared", got.flags[:16])
ok = ref.ok() & got.ok()
both = ref.verified() & got.verified()
print("ok ref %d shared %d both-verified %d of %d" % (ref.ok().sum(), got.ok().sum(), both.sum(), B))
if both.any():
    d = np.abs(ref.ctrl[both] - got.ctrl[both])
    print("max |ctrl diff| over verified:", d.max(), "obj diff", np.abs(ref.obj[both] - got.obj[both]).max())
if ok.any():
    d = np.abs(ref.ctrl[ok] - got.ctrl[ok])
    print("max |ctrl diff| over ok:", d.max())
print("mean iters ref %.0f shared %.0f" % (ref.iters.mean(), got.iters.mean()))
pl.close()
