"""Flat search records: the host-side data format either side of the hot path.

``IterationData`` mirrors the fields of the reference's per-vehicle ``iter_v``
that the search reads (hlc/controller/common/IterationData.m:4-33, built by
PrioritizedController.plan, hlc/controller/prioritized/PrioritizedController.m:297-324).
``SearchBatch`` is the SoA/CSR layout of ``pdmpc_batch_in`` (include/pdmpc_b200.h)
that the MEX shim produces from a cell array of such structs; ``BatchResult`` is
``pdmpc_batch_out``.
"""
from __future__ import annotations

import dataclasses
from typing import List, Sequence

import numpy as np

AREA_STRIDE = 8
CHECKER_SAT = 0
CHECKER_INTERX = 1


@dataclasses.dataclass
class IterationData:
    """nV == 1 subset of IterationData.m:4-33 (names kept)."""

    x0: np.ndarray                              # (x, y, yaw[, speed])
    trim_indices: int                           # 1-based
    reference_trajectory_points: np.ndarray     # [Hp, 2]
    v_ref: np.ndarray                           # [Hp]
    obstacles: List[np.ndarray] = dataclasses.field(default_factory=list)            # each [2, n] closed
    dynamic_obstacle_area: List[List[np.ndarray]] = dataclasses.field(default_factory=list)  # rows x Hp
    predicted_lanelet_boundary: tuple = (np.zeros((2, 0)), np.zeros((2, 0)))      # (left [2,nL], right [2,nR])
    amount: int = 1


@dataclasses.dataclass
class SearchBatch:
    Hp: int
    checker: int
    dt_seconds: float
    x0: np.ndarray
    y0: np.ndarray
    yaw0: np.ndarray
    trim0: np.ndarray
    ref_x: np.ndarray
    ref_y: np.ndarray
    v_ref: np.ndarray
    slot_ptr: np.ndarray
    poly_ptr: np.ndarray
    vert_x: np.ndarray
    vert_y: np.ndarray
    lane_ptr: np.ndarray
    lane_x: np.ndarray
    lane_y: np.ndarray

    @property
    def n(self) -> int:
        return int(self.x0.size)

    def input_bytes(self) -> int:
        return sum(int(getattr(self, f.name).nbytes) for f in dataclasses.fields(self)
                   if isinstance(getattr(self, f.name), np.ndarray))

    @staticmethod
    def from_iters(iters: Sequence[IterationData], Hp: int, checker: int, dt_seconds: float) -> "SearchBatch":
        n = len(iters)
        x0 = np.empty(n)
        y0 = np.empty(n)
        yaw0 = np.empty(n)
        trim0 = np.empty(n, dtype=np.int32)
        ref_x = np.empty((n, Hp))
        ref_y = np.empty((n, Hp))
        v_ref = np.empty((n, Hp))
        slot_ptr = [0]
        poly_ptr = [0]
        vx: list = []
        vy: list = []
        lane_ptr = [0]
        lx: list = []
        ly: list = []
        nv = 0
        nl = 0
        for i, it in enumerate(iters):
            x0[i], y0[i], yaw0[i] = it.x0[0], it.x0[1], it.x0[2]
            trim0[i] = it.trim_indices
            ref_x[i] = it.reference_trajectory_points[:, 0]
            ref_y[i] = it.reference_trajectory_points[:, 1]
            v_ref[i] = it.v_ref
            for k in range(Hp + 1):
                polys = it.obstacles if k == 0 else [row[k - 1] for row in it.dynamic_obstacle_area]
                for p in polys:
                    p = np.asarray(p, dtype=np.float64)
                    if p.shape[1] == 0:
                        continue
                    vx.append(p[0])
                    vy.append(p[1])
                    nv += p.shape[1]
                    poly_ptr.append(nv)
                slot_ptr.append(len(poly_ptr) - 1)
            for side in it.predicted_lanelet_boundary[:2]:
                side = np.asarray(side, dtype=np.float64)
                if side.size:
                    lx.append(side[0])
                    ly.append(side[1])
                    nl += side.shape[1]
                lane_ptr.append(nl)
        cat = lambda parts: np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(0)
        return SearchBatch(
            Hp=Hp, checker=checker, dt_seconds=dt_seconds, x0=x0, y0=y0, yaw0=yaw0, trim0=trim0,
            ref_x=ref_x.reshape(-1), ref_y=ref_y.reshape(-1), v_ref=v_ref.reshape(-1),
            slot_ptr=np.asarray(slot_ptr, dtype=np.int32), poly_ptr=np.asarray(poly_ptr, dtype=np.int32),
            vert_x=cat(vx), vert_y=cat(vy), lane_ptr=np.asarray(lane_ptr, dtype=np.int32),
            lane_x=cat(lx), lane_y=cat(ly))

    def to_iters(self) -> List[IterationData]:
        """Inverse of from_iters (the per-vehicle iter_v structs of a flat batch)."""
        Hp, out = self.Hp, []
        rx, ry, vr = self.ref_x.reshape(-1, Hp), self.ref_y.reshape(-1, Hp), self.v_ref.reshape(-1, Hp)

        def poly(p):
            v0, v1 = self.poly_ptr[p], self.poly_ptr[p + 1]
            return np.vstack([self.vert_x[v0:v1], self.vert_y[v0:v1]])

        for i in range(self.n):
            sp = self.slot_ptr[i * (Hp + 1): (i + 1) * (Hp + 1) + 1]
            obstacles = [poly(p) for p in range(sp[0], sp[1])]
            rows = max(int(sp[k + 1] - sp[k]) for k in range(1, Hp + 1))
            dyn = []
            for r in range(rows):   # row r = the r-th polygon of every step (empty where a step has fewer)
                dyn.append([poly(sp[k] + r) if sp[k] + r < sp[k + 1] else np.zeros((2, 0)) for k in range(1, Hp + 1)])
            lp = self.lane_ptr[2 * i: 2 * i + 3]
            sides = tuple(np.vstack([self.lane_x[lp[s]:lp[s + 1]], self.lane_y[lp[s]:lp[s + 1]]]) for s in range(2))
            out.append(IterationData(x0=np.array([self.x0[i], self.y0[i], self.yaw0[i]]), trim_indices=int(self.trim0[i]),
                                     reference_trajectory_points=np.stack([rx[i], ry[i]], axis=1), v_ref=vr[i].copy(),
                                     obstacles=obstacles, dynamic_obstacle_area=dyn, predicted_lanelet_boundary=sides))
        return out

    def select(self, idx: Sequence[int]) -> "SearchBatch":
        """Sub-batch of the given searches (re-based CSR)."""
        idx = np.asarray(idx, dtype=np.int64)
        Hp = self.Hp
        S = Hp + 1
        slot_ptr = [0]
        poly_ptr = [0]
        vsel = []
        lane_ptr = [0]
        lsel = []
        nv = 0
        nl = 0
        for i in idx:
            for s in range(S):
                a, b = self.slot_ptr[i * S + s], self.slot_ptr[i * S + s + 1]
                for p in range(a, b):
                    v0, v1 = self.poly_ptr[p], self.poly_ptr[p + 1]
                    vsel.append(np.arange(v0, v1))
                    nv += v1 - v0
                    poly_ptr.append(nv)
                slot_ptr.append(len(poly_ptr) - 1)
            for s in range(2):
                a, b = self.lane_ptr[2 * i + s], self.lane_ptr[2 * i + s + 1]
                lsel.append(np.arange(a, b))
                nl += b - a
                lane_ptr.append(nl)
        vs = np.concatenate(vsel).astype(np.int64) if vsel else np.zeros(0, dtype=np.int64)
        ls = np.concatenate(lsel).astype(np.int64) if lsel else np.zeros(0, dtype=np.int64)
        r2 = lambda a: np.ascontiguousarray(a.reshape(self.n, Hp)[idx].reshape(-1))
        return SearchBatch(
            Hp=Hp, checker=self.checker, dt_seconds=self.dt_seconds,
            x0=self.x0[idx].copy(), y0=self.y0[idx].copy(), yaw0=self.yaw0[idx].copy(),
            trim0=self.trim0[idx].copy(), ref_x=r2(self.ref_x), ref_y=r2(self.ref_y), v_ref=r2(self.v_ref),
            slot_ptr=np.asarray(slot_ptr, dtype=np.int32), poly_ptr=np.asarray(poly_ptr, dtype=np.int32),
            vert_x=self.vert_x[vs].copy(), vert_y=self.vert_y[vs].copy(),
            lane_ptr=np.asarray(lane_ptr, dtype=np.int32), lane_x=self.lane_x[ls].copy(),
            lane_y=self.lane_y[ls].copy())

    @staticmethod
    def concat(batches: Sequence["SearchBatch"]) -> "SearchBatch":
        b0 = batches[0]
        slot_ptr = [np.zeros(1, dtype=np.int64)]
        poly_ptr = [np.zeros(1, dtype=np.int64)]
        lane_ptr = [np.zeros(1, dtype=np.int64)]
        np_off = nv_off = nl_off = 0
        for b in batches:
            assert b.Hp == b0.Hp and b.checker == b0.checker and b.dt_seconds == b0.dt_seconds
            slot_ptr.append(b.slot_ptr[1:].astype(np.int64) + np_off)
            poly_ptr.append(b.poly_ptr[1:].astype(np.int64) + nv_off)
            lane_ptr.append(b.lane_ptr[1:].astype(np.int64) + nl_off)
            np_off += int(b.poly_ptr.size - 1)
            nv_off += int(b.vert_x.size)
            nl_off += int(b.lane_x.size)
        if nv_off >= 2**31 or np_off >= 2**31:
            raise ValueError("batch too large for int32 CSR offsets; split it")
        c = lambda name: np.ascontiguousarray(np.concatenate([getattr(b, name) for b in batches]))
        return SearchBatch(
            Hp=b0.Hp, checker=b0.checker, dt_seconds=b0.dt_seconds, x0=c("x0"), y0=c("y0"),
            yaw0=c("yaw0"), trim0=c("trim0"), ref_x=c("ref_x"), ref_y=c("ref_y"), v_ref=c("v_ref"),
            slot_ptr=np.concatenate(slot_ptr).astype(np.int32),
            poly_ptr=np.concatenate(poly_ptr).astype(np.int32), vert_x=c("vert_x"), vert_y=c("vert_y"),
            lane_ptr=np.concatenate(lane_ptr).astype(np.int32), lane_x=c("lane_x"), lane_y=c("lane_y"))

    def save(self, path: str) -> None:
        np.savez_compressed(path, **{f.name: getattr(self, f.name) for f in dataclasses.fields(self)})

    @staticmethod
    def load(path: str) -> "SearchBatch":
        z = np.load(path)
        kw = {k: z[k] for k in z.files}
        for k in ("Hp", "checker"):
            kw[k] = int(kw[k])
        kw["dt_seconds"] = float(kw["dt_seconds"])
        return SearchBatch(**kw)


@dataclasses.dataclass
class BatchResult:
    """pdmpc_batch_out as numpy arrays."""

    Hp: int
    status: np.ndarray
    is_exhausted: np.ndarray
    n_expanded: np.ndarray
    n_pops: np.ndarray
    pop_hash: np.ndarray
    trims: np.ndarray        # [n, Hp+1]
    tree_path: np.ndarray    # [n, Hp+1]
    y_predicted: np.ndarray  # [n, Hp, 3]
    g_path: np.ndarray       # [n, Hp+1]
    h_path: np.ndarray       # [n, Hp+1]
    shape_npts: np.ndarray   # [n, Hp]
    shape_x: np.ndarray      # [n, Hp, 8]
    shape_y: np.ndarray

    @staticmethod
    def empty(n: int, Hp: int) -> "BatchResult":
        return BatchResult(
            Hp=Hp,
            status=np.full(n, -1, dtype=np.int32),
            is_exhausted=np.zeros(n, dtype=np.uint8),
            n_expanded=np.zeros(n, dtype=np.int32),
            n_pops=np.zeros(n, dtype=np.int32),
            pop_hash=np.zeros(n, dtype=np.uint64),
            trims=np.zeros((n, Hp + 1), dtype=np.int32),
            tree_path=np.zeros((n, Hp + 1), dtype=np.int32),
            y_predicted=np.zeros((n, Hp, 3)),
            g_path=np.zeros((n, Hp + 1)),
            h_path=np.zeros((n, Hp + 1)),
            shape_npts=np.zeros((n, Hp), dtype=np.int32),
            shape_x=np.zeros((n, Hp, AREA_STRIDE)),
            shape_y=np.zeros((n, Hp, AREA_STRIDE)),
        )

    def output_bytes(self) -> int:
        return sum(int(getattr(self, f.name).nbytes) for f in dataclasses.fields(self)
                   if isinstance(getattr(self, f.name), np.ndarray))

    def shapes(self, i: int) -> List[np.ndarray]:
        """info.shapes(1,:) of search i as a list of [2, n] arrays."""
        out = []
        for k in range(self.Hp):
            n = int(self.shape_npts[i, k])
            out.append(np.vstack([self.shape_x[i, k, :n], self.shape_y[i, k, :n]]))
        return out


@dataclasses.dataclass
class TimestepDeps:
    """pdmpc_timestep_deps (include/pdmpc_b200.h): which searches of a batch wait for which,
    and the areas an exhausted search publishes (PrioritizedController.m:449-506, :568-621, :678-718)."""

    pred_ptr: np.ndarray     # [n+1]
    pred_idx: np.ndarray     # [pred_ptr[n]]
    fb_npts: np.ndarray      # [n, Hp]
    fb_x: np.ndarray         # [n, Hp, 8]
    fb_y: np.ndarray

    @staticmethod
    def build(preds: Sequence[Sequence[int]], fallback_shapes: Sequence[Sequence[np.ndarray]], Hp: int) -> "TimestepDeps":
        """preds[i] = searches search i waits for; fallback_shapes[i] = Hp closed [2, m] areas (or None)."""
        n = len(preds)
        pred_ptr = np.zeros(n + 1, dtype=np.int32)
        for i, p in enumerate(preds):
            pred_ptr[i + 1] = pred_ptr[i] + len(p)
        pred_idx = np.array([j for p in preds for j in p], dtype=np.int32).reshape(-1)
        fb_npts = np.zeros((n, Hp), dtype=np.int32)
        fb_x = np.zeros((n, Hp, AREA_STRIDE))
        fb_y = np.zeros((n, Hp, AREA_STRIDE))
        for i, shapes in enumerate(fallback_shapes):
            if shapes is None:
                continue
            for k in range(Hp):
                a = np.asarray(shapes[k], dtype=np.float64)
                m = a.shape[1]
                fb_npts[i, k] = m
                fb_x[i, k, :m] = a[0]
                fb_y[i, k, :m] = a[1]
        return TimestepDeps(pred_ptr, pred_idx, fb_npts, fb_x, fb_y)

    @staticmethod
    def concat(parts: Sequence["TimestepDeps"], sizes: Sequence[int]) -> "TimestepDeps":
        """Several independent blocks of searches (scenarios, time steps, priority permutations) as ONE call:
        predecessor indices are shifted by the searches in front of each block."""
        off, ptr, idx = 0, [np.zeros(1, dtype=np.int32)], []
        for d, n in zip(parts, sizes):
            ptr.append(d.pred_ptr[1:] + ptr[-1][-1])
            idx.append(d.pred_idx + off)
            off += int(n)
        return TimestepDeps(np.concatenate(ptr).astype(np.int32), np.concatenate(idx).astype(np.int32),
                            np.concatenate([d.fb_npts for d in parts]), np.concatenate([d.fb_x for d in parts]),
                            np.concatenate([d.fb_y for d in parts]))

    def preds(self, i: int) -> np.ndarray:
        return self.pred_idx[self.pred_ptr[i]:self.pred_ptr[i + 1]]

    def fallback_shapes(self, i: int) -> List[np.ndarray]:
        return [np.vstack([self.fb_x[i, k, :self.fb_npts[i, k]], self.fb_y[i, k, :self.fb_npts[i, k]]])
                for k in range(self.fb_npts.shape[1])]
