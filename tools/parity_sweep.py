#!/usr/bin/env python
"""tools/parity_sweep.py -- the parity policy of tests/helpers.py on more batches than the test-suite runs: config-2 shards
beyond the first, trapezoid-prism batches with other seeds, config 3 with other group counts.  Prints one line per batch."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import helpers as H
from spectral_b200 import api
from spectral_b200.scenarios import GOLDEN_W_CUB, GOLDEN_W_TRP, WEIGHTS_FILE, config2, config3, load_fixture, perturbed_obstacles

p = api.SpectralPlanner(device=0, max_batch=1024, n_max=256, r_max=8, k_max=32)
cases = [("cub", config2(1024, first=1024 * i), GOLDEN_W_CUB, "config2 shard %d" % i) for i in (1, 2, 3)]
cases += [("trp", perturbed_obstacles(load_fixture("c1"), 512, seed=s), GOLDEN_W_TRP, "trp c1 seed %d" % s) for s in (101, 202)]
cases += [("trp", config3(512, groups=g, first=4096), WEIGHTS_FILE, "config3 groups %d" % g) for g in (1, 64)]
cases += [("cub", perturbed_obstacles(load_fixture("c3"), 512, seed=303), WEIGHTS_FILE, "cub c3 seed 303")]
if os.environ.get("PARITY_SWEEP_MORE", "1") == "1":
    from spectral_b200.scenarios import mixed_batches
    cases += [(v, b, WEIGHTS_FILE, "mixed seed 777 #%d %s" % (i, v)) for i, (v, b) in enumerate(mixed_batches(1280, seed=777))]
    for name in ("c7", "c5", "bounds", "c4"):
        for v in ("trp", "cub"):
            cases.append((v, perturbed_obstacles(load_fixture(name), 128, seed=404, s_max=100.0 if name == "c7" else 50.0), WEIGHTS_FILE, "%s %s seed 404" % (v, name)))
    rng = np.random.default_rng(9)
    for Bodd in (1, 33, 1000):
        cases.append(("cub", config2(Bodd, first=5000), GOLDEN_W_CUB, "config2 B=%d" % Bodd))
    wr = np.tile(np.array(WEIGHTS_FILE), (256, 1))
    wr[:, :4] = rng.uniform(0.5, 50.0, (256, 4)); wr[:, 4:8] = rng.uniform(0.0, 20.0, (256, 4)); wr[:, 8:] = rng.uniform(0.0, 40.0, (256, 2))
    cases.append(("trp", perturbed_obstacles(load_fixture("c2"), 256, seed=505), wr, "trp c2 random weights"))
    cases.append(("cub", config2(256, first=9000), wr, "cub config2 random weights"))
bad = 0
for variant, batch, w, name in cases:
    t = time.time()
    got = p.solve(variant, batch, w)
    got0 = p.solve(variant, batch, w, options=api.default_options(polish=0))   # the reference's polish = 0: the raw ADMM iterate
    ref, ref0 = H.oracle_pair(variant, batch, w)
    try:
        both = H.assert_batch_parity(got, ref, name, need_verified_frac=0.4, ref0=ref0, batch=batch, variant=variant, weights=w, got0=got0)
        print("%-20s B=%4d ok %4d verified %4d compared %4d  PASS  (%.0f s)" % (name, batch.batch, got.ok().sum(), got.verified().sum(), both.sum(), time.time() - t))
    except AssertionError as e:
        bad += 1
        print("%-20s FAIL: %s" % (name, str(e)[:300]))
print("failures:", bad)
import json
log = H.PARITY_LOG
tot = lambda k: int(sum((d.get(k) or 0) for d in log))  # noqa: E731
itc = tot("iters_compared")
print(json.dumps({"parity_sweep_summary": {"batches": len(cases), "failures": bad, "scenarios": tot("B"), "decided": tot("decided"),
      "class_mismatch_decided": tot("class_mismatch_decided"), "class_mismatch_undecided": tot("class_mismatch_undecided"),
      "iters_compared": itc, "iters_equal": int(round(sum((d.get("iters_equal_frac") or 0) * (d.get("iters_compared") or 0) for d in log))),
      "compared_verified": tot("compared_verified"), "compared_iterate": tot("compared_iterate"), "iterate_outliers": tot("iterate_outliers"),
      "exceptions": tot("exceptions"), "charged_to_oracle": tot("charged_to_oracle")}}))
