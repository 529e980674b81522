"""Motion-primitive automaton tables (host side, numpy).

Only the TABLES the search reads at run time are built here:
trims, the (time-varying) transition matrix, and per-edge maneuvers with their
three occupied-area polygons.  In the MATLAB drop-in these tables come from the
reference's own ``MotionPrimitiveAutomaton`` object and are handed to
``pdmpc_upload_mpa`` by the MEX shim; this module exists so that the synthetic
scenario harness, the tests and ``bench.py`` have MPAs of the named shapes
without MATLAB.

Follows (reference file:line):
  * trim sets / adjacency ........ hlc/model/motion_primitive_automaton/choose_trims.m:11-135,
                                    build_mpa.m:1-72
  * maneuver integration ......... generate_maneuver.m:1-64 (ode45 RelTol=AbsTol=1e-8;
                                    here scipy RK45 = the same Dormand-Prince pair,
                                    rtol=atol=1e-10, not bit-identical to MATLAB)
  * kinematic bicycle ODE ........ hlc/model/differential_equations/BicycleModel.m:26-54
  * maneuver areas ............... generate_maneuver.m:68-105, utility/translate_global.m:19-22
  * time-varying transitions ..... MotionPrimitiveAutomaton.m:133-145,238-250
  * vehicle dimensions ........... scenarios/Vehicle.m:10-13
"""
from __future__ import annotations

import dataclasses
from collections import deque

import numpy as np

AREA_STRIDE = 8  # PDMPC_AREA_STRIDE

VEH_LENGTH = 0.22
VEH_WIDTH = 0.10
VEH_LF = 0.1
VEH_LR = 0.1


def choose_trims(mpa_type: str, max_acc_per_dt: float = 0.128, max_dec_per_dt: float | None = None):
    """choose_trims.m:1-135 -> (trim_inputs [nT,2] = (steering, speed), adjacency [nT,nT])."""
    if max_dec_per_dt is None:
        max_dec_per_dt = max_acc_per_dt
    if mpa_type == "single_speed":
        n_half = 5
        steering = np.linspace(-0.6, 0.6, 2 * n_half + 1)
        v_profile = np.arange(0, 9) * 0.1  # 0:0.1:0.8
        speed_left = v_profile[-n_half:]
        speed = np.concatenate([speed_left, [0.8], speed_left[::-1]])
        ntr = steering.size + 1
        trim_inputs = np.vstack([np.zeros((1, 2)), np.column_stack([steering, speed])])
        adj = np.ones((ntr, ntr))
        m = ntr - 1
        adj[1:, 1:] -= np.triu(np.ones((m, m)), 2) + np.tril(np.ones((m, m)), -2)
        return trim_inputs, adj
    if mpa_type == "triple_speed":
        n_sixth = 5
        steering = np.linspace(-0.6, 0.6, 2 * n_sixth + 1)
        n_third = steering.size
        speed = np.concatenate([np.full(n_third, 0.5), np.full(n_third, 0.7), np.full(n_third, 0.9)])
        ntr = 3 * n_third + 1
        trim_inputs = np.vstack([np.zeros((1, 2)), np.column_stack([np.tile(steering, 3), speed])])
        adj = np.ones((ntr, ntr))
        m = ntr - 1
        adj[1:, 1:] -= np.triu(np.ones((m, m)), 2) + np.tril(np.ones((m, m)), -2)
        # MATLAB indices below are 1-based; python index = matlab - 1
        adj[0, n_third + 1:] = 0
        adj[n_third + 1:, 0] = 0
        adj[n_third, n_third + 1] = 0
        adj[n_third + 1, n_third] = 0
        adj[2 * n_third, 2 * n_third + 1] = 0
        adj[2 * n_third + 1, 2 * n_third] = 0
        first = list(range(2, n_third + 2)) + list(range(n_third + 2, 2 * n_third + 2))
        second = list(range(n_third + 2, 2 * n_third + 2)) + list(range(2 * n_third + 2, ntr + 1))
        for i, j in zip(first, second):
            adj[i - 1, j - 1] = 1
            adj[j - 1, i - 1] = 1
        return trim_inputs, adj
    if mpa_type == "realistic":
        d_speed = min(max_acc_per_dt, max_dec_per_dt)
        acc_max = 1.05 * max_acc_per_dt
        dec_max = 1.05 * max_dec_per_dt
        speed_max = d_speed * round(0.8 / d_speed)
        n_speeds = int(round(speed_max / d_speed)) + 1
        speed_vec = np.arange(n_speeds) * d_speed
        d_steer = 0.5 * np.pi / 18
        steer_lo = d_steer * round((3 * np.pi / 18) / d_steer)
        steer_hi = d_steer * round((2 * np.pi / 18) / d_steer)
        d_steer_max = 1.05 * d_steer

        def sym_range(mx):
            n = int(round(mx / d_steer))
            return np.arange(-n, n + 1) * d_steer

        steer_cla = [sym_range(steer_lo)]
        xs = [d_speed, speed_vec[2]]
        vs = [steer_lo, steer_hi]
        for i_speed in (1, 2):
            mx = np.interp(speed_vec[i_speed], xs, vs)
            mx = d_steer * round(mx / d_steer)
            steer_cla.append(sym_range(mx))
        for _ in range(3, n_speeds):
            steer_cla.append(sym_range(steer_hi))
        # build_mpa.m:24-70
        rows = []
        for sp, st in zip(speed_vec, steer_cla):
            for s in st:
                rows.append((s, sp))
        trim_inputs = np.array(rows)
        ntr = trim_inputs.shape[0]
        adj = np.zeros((ntr, ntr))
        for i in range(ntr):
            for j in range(ntr):
                if abs(trim_inputs[j, 0] - trim_inputs[i, 0]) <= d_steer_max:
                    if trim_inputs[j, 1] > trim_inputs[i, 1]:
                        ok = (trim_inputs[j, 1] - trim_inputs[i, 1]) <= acc_max
                    else:
                        ok = (trim_inputs[i, 1] - trim_inputs[j, 1]) <= dec_max
                    if ok:
                        adj[i, j] = 1
        return trim_inputs, adj
    raise ValueError(f"unknown mpa type {mpa_type!r}")


def _bicycle_ode(_t, x, steering_derivative, acceleration):
    """BicycleModel.m:26-54, centred kinematic bicycle."""
    L = VEH_LF + VEH_LR
    R = VEH_LR / L
    psi, v, delta = x[2], x[3], x[4]
    beta = np.arctan(R * np.tan(delta))
    return np.array([
        v * np.cos(psi + beta),
        v * np.sin(psi + beta),
        v / L * np.tan(delta) * np.cos(beta),
        acceleration,
        steering_derivative,
    ])


def _translate_global(yaw, x0, y0, xl, yl):
    """utility/translate_global.m:19-22"""
    c, s = np.cos(yaw), np.sin(yaw)
    return c * xl - s * yl + x0, s * xl + c * yl + y0


def _maneuver_area(x1, y1, x2, y2, signum, non_convex):
    """generate_maneuver.m:68-105 (indices there are 1-based)."""
    if signum == 0:
        xs = [x1[0], x1[1], x2[2], x2[3], x1[0]]
        ys = [y1[0], y1[1], y2[2], y2[3], y1[0]]
    elif signum > 0:
        if non_convex:
            xs = [x1[0], x1[1], x2[1], x2[2], x2[3], x1[3], x1[0]]
            ys = [y1[0], y1[1], y2[1], y2[2], y2[3], y1[3], y1[0]]
        else:
            xs = [x1[0], x1[1], x2[2], x2[3], x2[3], x1[0]]
            ys = [y1[0], y1[1], y2[2], y2[3], y1[3], y1[0]]
    else:
        if non_convex:
            xs = [x1[0], x1[1], x1[2], x2[2], x2[3], x2[0], x1[0]]
            ys = [y1[0], y1[1], y1[2], y2[2], y2[3], y2[0], y1[0]]
        else:
            xs = [x1[0], x1[1], x2[2], x2[2], x2[3], x1[0]]
            ys = [y1[0], y1[1], y1[2], y2[2], y2[3], y1[0]]
    return np.array(xs), np.array(ys)


@dataclasses.dataclass
class MotionPrimitiveAutomaton:
    """The table subset of the reference's MotionPrimitiveAutomaton object."""

    mpa_type: str
    Hp: int
    dt_seconds: float
    non_convex: bool
    recursive_feasibility: bool
    trim_steering: np.ndarray      # [nT]
    trim_speed: np.ndarray         # [nT]
    transition: np.ndarray         # uint8 [Hp, nT, nT]  (step, from, to)
    adjacency: np.ndarray          # uint8 [nT, nT] time-invariant
    edge_from: np.ndarray          # int32 [nE] 1-based
    edge_to: np.ndarray            # int32 [nE] 1-based
    edge_dx: np.ndarray
    edge_dy: np.ndarray
    edge_dyaw: np.ndarray
    area_npts: np.ndarray          # int32 [nE, 3]
    area_x: np.ndarray             # [nE, 3, 8]
    area_y: np.ndarray
    distance_to_equilibrium: np.ndarray
    edge_index: np.ndarray         # int32 [nT, nT] -> edge or -1

    @property
    def n_trims(self) -> int:
        return int(self.trim_speed.size)

    @property
    def n_edges(self) -> int:
        return int(self.edge_from.size)

    def get_straight_speeds_of_mpa(self) -> np.ndarray:
        """MotionPrimitiveAutomaton.m:187-191"""
        m = (self.trim_speed > 0) & (self.trim_steering == 0)
        return self.trim_speed[m]

    def get_max_speed_of_mpa(self) -> float:
        return float(self.trim_speed.max())

    def trim_from_values(self, speed: float, steering: float) -> int:
        """MotionPrimitiveAutomaton.m:193-236; returns a 1-based trim index."""
        if steering == 0:
            idx = np.flatnonzero(self.trim_steering == 0)
            return int(idx[np.argmin(np.abs(self.trim_speed[idx] - speed))]) + 1
        sc, ss = self.trim_speed.min(), self.trim_speed.max() - self.trim_speed.min()
        tc, ts = self.trim_steering.min(), self.trim_steering.max() - self.trim_steering.min()
        d = np.hypot((self.trim_speed - sc) / ss - (speed - sc) / ss,
                     (self.trim_steering - tc) / ts - (steering - tc) / ts)
        return int(np.argmin(d)) + 1

    def full_tree_nodes(self) -> int:
        """Nodes of the complete search tree from the worst start trim (capacity bound)."""
        nT = self.n_trims
        worst = 1
        for t0 in range(nT):
            cnt = np.zeros(nT)
            cnt[t0] = 1
            total = 1
            for k in range(self.Hp):
                cnt = cnt @ self.transition[k].astype(np.float64)
                total += cnt.sum()
            worst = max(worst, int(total))
        return worst


def build_mpa(mpa_type: str = "single_speed", Hp: int = 6, dt_seconds: float = 0.2,
              non_convex: bool = False, recursive_feasibility: bool = True,
              offset: float = 0.01) -> MotionPrimitiveAutomaton:
    """MotionPrimitiveAutomaton.m:25-153 restricted to the tables the search reads."""
    from scipy.integrate import solve_ivp

    acc = 0.64 * dt_seconds  # MotionPrimitiveAutomaton.m:38-41
    trim_inputs, adj = choose_trims(mpa_type, acc, acc)
    nT = trim_inputs.shape[0]
    steering, speed = trim_inputs[:, 0].copy(), trim_inputs[:, 1].copy()

    e_from, e_to, dxs, dys, dyaws = [], [], [], [], []
    npts = []
    ax = []
    ay = []
    for i in range(nT):
        for j in range(nT):
            if not adj[i, j]:
                continue
            sd = (steering[j] - steering[i]) / dt_seconds   # generate_maneuver.m:7-8
            ac = (speed[j] - speed[i]) / dt_seconds
            x0 = np.array([0.0, 0.0, 0.0, speed[i], steering[i]])
            sol = solve_ivp(_bicycle_ode, (0.0, dt_seconds), x0, method="RK45", args=(sd, ac),
                            rtol=1e-10, atol=1e-10)
            dx, dy, dyaw = (float(v) for v in sol.y[:3, -1])
            signum = int(np.sign(dyaw))
            kinds_x, kinds_y, kinds_n = [], [], []
            for (lx, ly) in ((VEH_LENGTH / 2 + offset, VEH_WIDTH / 2 + offset),   # :40-41
                             (VEH_LENGTH / 2, VEH_WIDTH / 2),                      # :49-50
                             (VEH_LENGTH / 2 + 0.05, VEH_WIDTH / 2 + 0.0)):        # :58-59
                x1 = np.array([-1.0, -1.0, 1.0, 1.0]) * lx
                y1 = np.array([-1.0, 1.0, 1.0, -1.0]) * ly
                x2, y2 = _translate_global(dyaw, dx, dy, x1, y1)
                px, py = _maneuver_area(x1, y1, x2, y2, signum, non_convex)
                n = px.size
                kinds_n.append(n)
                kinds_x.append(np.pad(px, (0, AREA_STRIDE - n)))
                kinds_y.append(np.pad(py, (0, AREA_STRIDE - n)))
            e_from.append(i + 1)
            e_to.append(j + 1)
            dxs.append(dx)
            dys.append(dy)
            dyaws.append(dyaw)
            npts.append(kinds_n)
            ax.append(kinds_x)
            ay.append(kinds_y)

    # distance to equilibrium: MotionPrimitiveAutomaton.m:133-136 (undirected graph distances)
    und = ((adj + adj.T) > 0)
    dist = np.full(nT, np.iinfo(np.int32).max, dtype=np.int64)
    dq = deque()
    for e in np.flatnonzero(speed == 0):
        dist[e] = 0
        dq.append(int(e))
    while dq:
        u = dq.popleft()
        for v in np.flatnonzero(und[u]):
            if dist[v] > dist[u] + 1:
                dist[v] = dist[u] + 1
                dq.append(int(v))

    trans = np.repeat(adj[None, :, :], Hp, axis=0).astype(np.uint8)
    if recursive_feasibility:  # :238-250
        for k in range(1, Hp + 1):
            trans[k - 1][:, dist > (Hp - k)] = 0

    edge_index = np.full((nT, nT), -1, dtype=np.int32)
    for e, (f, t) in enumerate(zip(e_from, e_to)):
        edge_index[f - 1, t - 1] = e

    return MotionPrimitiveAutomaton(
        mpa_type=mpa_type, Hp=Hp, dt_seconds=dt_seconds, non_convex=non_convex,
        recursive_feasibility=recursive_feasibility,
        trim_steering=steering, trim_speed=speed,
        transition=np.ascontiguousarray(trans), adjacency=adj.astype(np.uint8),
        edge_from=np.array(e_from, dtype=np.int32), edge_to=np.array(e_to, dtype=np.int32),
        edge_dx=np.array(dxs), edge_dy=np.array(dys), edge_dyaw=np.array(dyaws),
        area_npts=np.array(npts, dtype=np.int32),
        area_x=np.ascontiguousarray(np.array(ax)), area_y=np.ascontiguousarray(np.array(ay)),
        distance_to_equilibrium=dist, edge_index=edge_index,
    )


_CACHE: dict = {}


def get_mpa(mpa_type: str = "single_speed", Hp: int = 6, dt_seconds: float = 0.2,
            non_convex: bool = False, recursive_feasibility: bool = True) -> MotionPrimitiveAutomaton:
    """Cached build, the analogue of the reference's library/*.mat cache
    (MotionPrimitiveAutomaton.m:59-79)."""
    key = (mpa_type, Hp, dt_seconds, non_convex, recursive_feasibility)
    if key not in _CACHE:
        _CACHE[key] = build_mpa(mpa_type, Hp, dt_seconds, non_convex, recursive_feasibility)
    return _CACHE[key]
