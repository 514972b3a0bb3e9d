"""Wire formats of the planning hot path.

* ``read_scenario_text`` / ``write_scenario_text``: the whitespace-separated ``c_road_*.txt``
  grammar that the reference's ``find_traj`` parses (trp_wrapper.cpp:39-144, identical in
  cub_wrapper.cpp:38-140; writer: cart_frenet.py:384-453).
* ``read_trajectory_text``: the 7-column ``t s l ds dl dds ddl`` file that ``find_traj`` writes with
  ``std::fixed << std::setprecision(3)`` (trp_wrapper.cpp:288-301).
* ``ScenarioBatch``: the packed, batch-major binary layout consumed by the C-ABI
  (include/spectral.h): everything FP64, C-contiguous.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

N_SCALARS = 10  # ds_ref, dl_ref, dds_lo, dds_hi, ddds_lo, ddds_hi, ddl_lo, ddl_hi, dddl_lo, dddl_hi


@dataclass
class Scenario:
    """One planning problem, exactly the content of a ``c_road_*.txt`` file."""

    n_knots: int
    delta_t: float
    init_s: np.ndarray  # [3] s, ds, dds
    init_l: np.ndarray  # [3]
    scalars: np.ndarray  # [10], see N_SCALARS
    s_bounds: np.ndarray  # [R, N, 2]
    l_bounds: np.ndarray  # [R, N, 2]
    ds_bounds: np.ndarray  # [N, 2]
    dl_bounds: np.ndarray  # [N, 2]
    s_ref: np.ndarray  # [N]
    l_ref: np.ndarray  # [N]
    s_kappa: np.ndarray = field(default_factory=lambda: np.zeros(0))  # read, never used (:134-144)
    l_kappa: np.ndarray = field(default_factory=lambda: np.zeros(0))

    @property
    def n_regions(self) -> int:
        return int(self.s_bounds.shape[0])


def read_scenario_text(path: str) -> Scenario:
    """Parse the reference's input grammar.  Token order follows trp_wrapper.cpp:39-144."""
    with open(path) as f:
        tok = f.read().split()
    pos = 0

    def take(k: int) -> np.ndarray:
        nonlocal pos
        vals = tok[pos:pos + k]
        pos += k
        return np.array([float(v) for v in vals], dtype=np.float64)

    n = int(float(tok[0]))
    delta = float(tok[1])
    pos = 2
    init_s = take(3)
    init_l = take(3)
    r = int(float(tok[pos]))
    pos += 1
    ds_ref, dl_ref = take(2)
    dds = take(2)
    ddds = take(2)
    ddl = take(2)
    dddl = take(2)
    s_bounds = np.zeros((r, n, 2))
    l_bounds = np.zeros((r, n, 2))
    for i in range(r):
        s_bounds[i] = take(2 * n).reshape(n, 2)
        l_bounds[i] = take(2 * n).reshape(n, 2)
    ds_bounds = take(2 * n).reshape(n, 2)
    dl_bounds = take(2 * n).reshape(n, 2)
    s_ref = take(n)
    l_ref = take(n)
    if len(s_ref) != n or len(l_ref) != n:
        raise ValueError("%s: truncated before the reference trajectories" % path)
    s_kappa = take(n)
    l_kappa = take(n)
    scalars = np.array([ds_ref, dl_ref, dds[0], dds[1], ddds[0], ddds[1], ddl[0], ddl[1], dddl[0], dddl[1]])
    return Scenario(n, delta, init_s, init_l, scalars, s_bounds, l_bounds, ds_bounds, dl_bounds,
                    s_ref, l_ref, s_kappa, l_kappa)


def write_scenario_text(path: str, sc: Scenario) -> None:
    """Write a scenario in the grammar above (17 significant digits, so a round trip is exact)."""
    def row(a: Sequence[float]) -> str:
        return " ".join(repr(float(v)) for v in np.asarray(a).ravel())

    n = sc.n_knots
    kap_s = sc.s_kappa if len(sc.s_kappa) == n else np.zeros(n)
    kap_l = sc.l_kappa if len(sc.l_kappa) == n else np.zeros(n)
    lines = ["%d %r" % (n, float(sc.delta_t)), row(sc.init_s), row(sc.init_l), "", str(sc.n_regions), "",
             row(sc.scalars[0:2]), "", row(sc.scalars[2:4]), row(sc.scalars[4:6]), "",
             row(sc.scalars[6:8]), row(sc.scalars[8:10]), ""]
    for r in range(sc.n_regions):
        lines += [row(sc.s_bounds[r]), "", row(sc.l_bounds[r]), ""]
    lines += [row(sc.ds_bounds), "", row(sc.dl_bounds), "", row(sc.s_ref), "", row(sc.l_ref), "",
              row(kap_s), "", row(kap_l), ""]
    with open(path, "w") as f:
        f.write("\n".join(lines))


def read_trajectory_text(path: str) -> np.ndarray:
    """[n_points, 7] array of ``t s l ds dl dds ddl`` (trp_wrapper.cpp:298-301)."""
    rows = []
    with open(path) as f:
        for line in f:
            p = line.split()
            if len(p) == 7:
                rows.append([float(v) for v in p])
    return np.array(rows, dtype=np.float64).reshape(-1, 7)


@dataclass
class ScenarioBatch:
    """B scenarios sharing (N, R, delta_t), packed for the C-ABI (include/spectral.h)."""

    n_knots: int
    n_regions: int
    delta_t: float
    s_bounds: np.ndarray  # [B, R, N, 2]
    l_bounds: np.ndarray  # [B, R, N, 2]
    ds_bounds: np.ndarray  # [B, N, 2]
    dl_bounds: np.ndarray  # [B, N, 2]
    s_ref: np.ndarray  # [B, N]
    l_ref: np.ndarray  # [B, N]
    init: np.ndarray  # [B, 6] s, ds, dds, l, dl, ddl
    scalars: np.ndarray  # [B, 10]

    @property
    def batch(self) -> int:
        return int(self.s_ref.shape[0])

    @staticmethod
    def from_scenarios(scs: List[Scenario]) -> "ScenarioBatch":
        n, r, d = scs[0].n_knots, scs[0].n_regions, scs[0].delta_t
        for s in scs:
            if (s.n_knots, s.n_regions, s.delta_t) != (n, r, d):
                raise ValueError("a ScenarioBatch needs one (N, R, delta_t)")
        c = np.ascontiguousarray
        return ScenarioBatch(
            n, r, d,
            c(np.stack([s.s_bounds for s in scs])), c(np.stack([s.l_bounds for s in scs])),
            c(np.stack([s.ds_bounds for s in scs])), c(np.stack([s.dl_bounds for s in scs])),
            c(np.stack([s.s_ref for s in scs])), c(np.stack([s.l_ref for s in scs])),
            c(np.stack([np.concatenate([s.init_s, s.init_l]) for s in scs])),
            c(np.stack([s.scalars for s in scs])))

    def slice(self, lo: int, hi: int) -> "ScenarioBatch":
        return ScenarioBatch(self.n_knots, self.n_regions, self.delta_t, self.s_bounds[lo:hi],
                             self.l_bounds[lo:hi], self.ds_bounds[lo:hi], self.dl_bounds[lo:hi],
                             self.s_ref[lo:hi], self.l_ref[lo:hi], self.init[lo:hi], self.scalars[lo:hi])

    def arrays(self):
        return (self.s_bounds, self.l_bounds, self.ds_bounds, self.dl_bounds, self.s_ref, self.l_ref,
                self.init, self.scalars)

    def save(self, path: str) -> None:
        """Packed binary batch file (npz): lets CPU and GPU runs share identical inputs."""
        np.savez(path, n_knots=self.n_knots, n_regions=self.n_regions, delta_t=self.delta_t,
                 s_bounds=self.s_bounds, l_bounds=self.l_bounds, ds_bounds=self.ds_bounds,
                 dl_bounds=self.dl_bounds, s_ref=self.s_ref, l_ref=self.l_ref, init=self.init,
                 scalars=self.scalars)

    @staticmethod
    def load(path: str) -> "ScenarioBatch":
        z = np.load(path)
        return ScenarioBatch(int(z["n_knots"]), int(z["n_regions"]), float(z["delta_t"]), z["s_bounds"],
                             z["l_bounds"], z["ds_bounds"], z["dl_bounds"], z["s_ref"], z["l_ref"],
                             z["init"], z["scalars"])
