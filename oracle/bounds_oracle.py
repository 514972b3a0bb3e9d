"""oracle/bounds_oracle.py -- TEST INFRASTRUCTURE: CPU restatement of the step right BEFORE the planning hot path
(SURVEY.md 8f row 1): obstacles (centre, velocity, horizon) -> space-time prisms -> lateral lane edges -> per-lane,
per-knot s-limits, i.e. the region bounds that find_traj reads.

    obstacle prism / lateral edges   /root/reference/src/cart_frenet.py:687-804  Car.getCar (vel_l >= 0: the eleven overlap
                                     cases s1..s11 against the cars created before; vel_l < 0: own edges only)
    lineFromPoints                   :819-830   s-limit per knot, round(.., 2)
    get_bounds ('yield' homotopy)    :833-1026  lanes per car and edge pair, free lanes, merge of equal lanes, gap lanes
The caller's sequence is reproduced as it runs (cart_frenet.py:1539-1557): for every obstacle `Car(...)` (whose __init__
calls getCar) followed by an explicit `.getCar()`, then get_bounds(Car._lateral).
Pinned against the reference's own classes (executed from its source by oracle/gen_bounds_golden.py):
tests/golden/bounds.npz.  Not specified by the reference (hash-ordered): the order of cars whose smallest lateral edge
ties -- such inputs are outside the pinned domain.  Only tests/ use this module.
"""
import numpy as np

L_SAFE = 5 / 3 + 5 / 3   # :693-699
W_SAFE = 2 / 3 + 2 / 3
ROAD = dict(s_l_l=0.0, s_u_l=50.0, d_l_l=-2.0, d_u_l=8.0)   # :54-58


class Prism:
    """The numbers of Car.car that get_bounds reads: rear / front face lines in (t, s) and the lateral extent."""

    def __init__(self, centre, vel_s, vel_l, horizon):
        s, l, t0 = centre
        fs, fl, ft = s + vel_s * horizon, l + vel_l * horizon, t0 + horizon          # forw_state :705-706
        if vel_l >= 0:
            self.min_l, self.max_l = l - W_SAFE, fl + W_SAFE                          # :710-711
        else:
            self.min_l, self.max_l = fl - W_SAFE, l + W_SAFE                          # :781-782
        self.t0, self.t1 = t0, ft
        self.rear0, self.rear1 = s - L_SAFE, fs - L_SAFE                              # car[0][0], car[4][0]
        self.front0, self.front1 = s + L_SAFE, fs + L_SAFE                            # car[2][0], car[6][0]
        self.vel_l = vel_l


def lateral_edges(obstacles):
    """obstacles: [(centre(s, l, t0), vel_s, vel_l, horizon)] in creation order.  Returns (prisms, edges) with edges[k] = the
    lateral values recorded for car k (Car._lateral entries with its ref), duplicates kept out."""
    prisms, edges = [], []
    for k, (centre, vel_s, vel_l, horizon) in enumerate(obstacles):
        p = Prism(centre, vel_s, vel_l, horizon)
        prisms.append(p)
        edges.append([])
        for _call in range(2):                         # Car.__init__ (:676) and the caller's explicit getCar (:1540)
            if p.vel_l >= 0:
                snapshot = [(j, v) for j in range(k + 1) for v in edges[j]]   # deepcopy(Car._lateral) :716
                for j, _v in snapshot:
                    a0, a1 = prisms[j].min_l, prisms[j].max_l                 # car[0].car[0][1], car[0].car[1][1]
                    mn, mx = p.min_l, p.max_l
                    if a0 == mn and a1 == mx:          # s9
                        continue
                    if a1 < mn:                        # s1
                        continue
                    if a0 < mn and a1 == mn:           # s7
                        continue
                    if a0 < mn and a1 > mn:
                        if a1 < mx:                    # s2
                            edges[k].append(a1); edges[j].append(mn)
                        elif a1 == mx:                 # s8
                            pass
                        else:                          # s6
                            edges[j].append(mn); edges[j].append(mx)
                        continue
                    if a0 == mn and a1 < mx:           # s10
                        edges[k].append(a1); continue
                    if a0 > mn and a1 < mx:            # s3
                        edges[k].append(a0); edges[k].append(a1); continue
                    if a0 > mn and a0 < mx and a1 > mx:  # s4
                        edges[k].append(a0); edges[j].append(mx); continue
                    if a0 > mn and a1 == mx:           # s11
                        edges[k].append(a0); continue
                    if a0 > mx:                        # s5
                        continue
            edges[k].append(p.min_l); edges[k].append(p.max_l)   # :760-761 / :784-785
    return prisms, [sorted(set(e)) for e in edges]


def line_from_points(x1, y1, x2, y2, n_knots):
    c = (y2 - y1) / (x2 - x1)                                       # :821-823
    return [round(c * i / 10 - c * x1 + y1, 2) for i in range(n_knots)]   # :826-828


def get_bounds(obstacles, n_knots=71, road=ROAD):
    """Returns [(s_bounds [N][2], l_bounds (lo, hi))] per lane, in the reference's order."""
    s_l_l, s_u_l, d_l_l, d_u_l = road["s_l_l"], road["s_u_l"], road["d_l_l"], road["d_u_l"]
    prisms, edges = lateral_edges(obstacles)
    # cars keyed in order of first appearance in the val-sorted list (:861-866): by smallest lateral edge
    order = sorted(range(len(prisms)), key=lambda k: edges[k][0])
    s_b, l_b = [], []
    flag = False
    free = [[s_l_l, s_u_l] for _ in range(n_knots)]
    for k in order:
        p, e = prisms[k], edges[k]
        for a in range(len(e) - 1):
            if e[a] > d_l_l and not flag:                                   # :878-885
                s_b.append([r[:] for r in free]); l_b.append((d_l_l, e[a])); flag = True
            if p.t0 == 0:                                                   # :901  rear face bounds s from above
                line = line_from_points(p.t0, p.rear0, p.t1, p.rear1, n_knots)
                cur = [[s_l_l, (s_u_l if (i < p.t0 * 10 or i > p.t1 * 10) else line[i])] for i in range(n_knots)]
            else:                                                           # :930  front face bounds s from below
                line = line_from_points(p.t0, p.front0, p.t1, p.front1, n_knots)
                cur = [[(s_l_l if (i < p.t0 * 10 or i > p.t1 * 10) else line[i]), s_u_l] for i in range(n_knots)]
            if not (e[a] > e[a + 1] or e[a] == e[a + 1]):                   # :921-925 pop degenerate lanes
                s_b.append(cur); l_b.append((e[a], e[a + 1]))
    max_l = max(edges[k][-1] for k in order)
    if max_l < d_u_l:                                                       # :959-965
        s_b.append([r[:] for r in free]); l_b.append((max_l, d_u_l))
    check = set()
    for l1 in range(len(l_b) - 1):                                          # :975-983 equal lanes: intersect
        for l2 in range(l1 + 1, len(l_b)):
            if l_b[l1] == l_b[l2]:
                m = [[max(s_b[l1][i][0], s_b[l2][i][0]), min(s_b[l1][i][1], s_b[l2][i][1])] for i in range(n_knots)]
                s_b[l1] = m
                s_b[l2] = [r[:] for r in m]
                check.add(l2)
    for idx in sorted(check, reverse=True):
        l_b.pop(idx); s_b.pop(idx)
    for l in range(len(l_b) - 1):                                           # :990-997 (range fixed before the inserts)
        if l + 1 < len(l_b) and l_b[l][1] != l_b[l + 1][0]:
            s_b.insert(l + 1, [r[:] for r in free]); l_b.insert(l + 1, (l_b[l][1], l_b[l + 1][0]))
    return [(np.array(s), lb) for s, lb in zip(s_b, l_b)]
