"""oracle/ref_python.py -- TEST INFRASTRUCTURE: runs the reference's OWN Python functions without importing its module.

/root/reference/src/cart_frenet.py cannot be imported (CommonRoad, a CommonRoad scenario file and a 141-step simulation
run at module level).  Its pure functions and the two classes around the planning hot path can: this module parses
the file with `ast`, takes the definitions it is asked for (source text untouched) and executes them in a namespace
that holds what they need (math, numpy, copy, the module's own constants from cart_frenet.py:54-64, and stand-ins
for the four CommonRoad value classes run_ego() instantiates).  Used only by the golden-vector generators
(oracle/gen_downstream_golden.py, oracle/gen_bounds_golden.py) IN THE BUILD CONTAINER; the vectors they write are
committed under tests/golden/ and nothing at test / run time reads /root/reference.
"""
import ast
import copy
import math
import os

import numpy as np

REF = "/root/reference/src/cart_frenet.py"


class _Obj:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _state(**kw):          # commonroad.scenario.trajectory.State(position=, velocity=, orientation=, time_step=)
    return _Obj(**kw)


def _trajectory(initial_time_step, state_list):   # Trajectory(curr, state_list)
    return _Obj(initial_time_step=initial_time_step, state_list=state_list)


def _rectangle(width, length):
    return _Obj(width=width, length=length)


def _prediction(trajectory, shape):               # TrajectoryPrediction(traj, shape)
    return _Obj(trajectory=trajectory, shape=shape)


def load(names, extra=None):
    """Namespace with the reference definitions `names` (functions / classes of cart_frenet.py) executed in it."""
    if not os.path.exists(REF):
        raise FileNotFoundError(REF)
    src = open(REF).read()
    tree = ast.parse(src)
    ns = {"np": np, "copy": copy, "math": math, "State": _state, "Trajectory": _trajectory, "Rectangle": _rectangle,
          "TrajectoryPrediction": _prediction,
          # cart_frenet.py:54-64
          "s_u_l": 50.0, "s_l_l": 0.0, "d_u_l": 8.0, "d_l_l": -2.0, "time_": 7.0, "num_of_knots": 71, "homotopy": "yield"}
    for k in dir(math):                                   # `from math import *` (cart_frenet.py:24)
        if not k.startswith("_"):
            ns[k] = getattr(math, k)
    if extra:
        ns.update(extra)
    want = set(names)
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in want:
            code = compile(ast.Module(body=[node], type_ignores=[]), REF, "exec")
            exec(code, ns)
            want.discard(node.name)
    if want:
        raise KeyError("not found in the reference: %s" % sorted(want))
    return ns
