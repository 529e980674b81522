"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Bar: bit-exact trims / flags / node counts / pop order, doubles bit-identical
(and a fortiori within the 1e-9 relative tolerance BASELINE.json states)."""
import dataclasses

import numpy as np
import pytest

from oracle import oracle_py, parity
from pdmpc_b200 import capi
from pdmpc_b200.mpa import build_mpa, get_mpa
from pdmpc_b200.records import CHECKER_INTERX, CHECKER_SAT, SearchBatch

from helpers import GOLDEN_CASES, circle_records, load_golden, rect, road_records, straight_iter

pytestmark = pytest.mark.gpu


# one search per warp / tiles (2 and 4 searches per warp) / CTA per search / CTA with valid-only queue:
# identical results required (shape 5 reports pop_hash over the popped nodes that passed their edge check)
VARIANTS = (1, 2, 3, 4, 5)


def check(planner, mpa, batch, variants=VARIANTS, **kw):
    planner.upload_mpa(mpa)
    ref = oracle_py.plan_batch(mpa, batch)
    ref5 = None
    info = dev = None
    try:
        for variant in variants:
            planner.set_variant(variant)
            dev = planner.plan_batch(batch, raise_on_search_error=False)
            ran = int(planner.stats().shape)
            if variant:   # a requested shape runs, or falls back to shape 1 for a documented reason — never silently
                why = ((variant in (2, 3) and batch.checker != CHECKER_INTERX) or
                       (variant in (4, 5) and mpa.full_tree_nodes() + 8 > 32768))
                assert ran == (1 if why else variant), (variant, ran)
            if variant == 5:
                # pop_hash covers the valid pops only — unless the shape fell back to shape 1 (search
                # trees beyond 32768 nodes), which hashes every pop
                if ref5 is None:
                    ref5 = oracle_py.plan_batch(mpa, batch, hash_valid_pops_only=True)
                    parity.compare(ref5, ref, skip=("pop_hash",))
                info = parity.compare(dev, ref, skip=("pop_hash",), **kw)
                assert np.array_equal(dev.pop_hash, ref5.pop_hash) or np.array_equal(dev.pop_hash, ref.pop_hash)
            else:
                info = parity.compare(dev, ref, **kw)
    finally:
        planner.set_variant(0)
    return info, dev, ref


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_fixture(planner, name):
    """CUDA path through the C ABI against the committed fixtures (tests/golden, made by
    tools/make_golden.py after both CPU restatements agreed): no oracle involved at run time."""
    mpa, batch, exp = load_golden(name)
    planner.upload_mpa(mpa)
    try:
        for variant in VARIANTS:
            planner.set_variant(variant)
            dev = planner.plan_batch(batch, raise_on_search_error=False)
            parity.compare(dev, exp, skip=("pop_hash",) if variant == 5 else ())
    finally:
        planner.set_variant(0)


def test_circle_config0_sat(planner):
    mpa, batch = circle_records(30)
    info, dev, _ = check(planner, mpa, batch)
    assert info["n"] == batch.n and info["pops"] > 1000


@pytest.mark.parametrize("mpa_type", ["single_speed", "triple_speed"])
def test_road_config1_interx(planner, mpa_type):
    mpa, batch = road_records(mpa_type, 8)
    info, dev, _ = check(planner, mpa, batch)
    assert info["n"] == batch.n


def test_road_realistic_mpa(planner):
    mpa, batch = road_records("realistic", 3, amount=10, seed=5)
    check(planner, mpa, batch)


def test_sat_checker_on_road_records(planner):
    """SAT + lanelet-boundary path (used when not prioritized, OptimizerInterface.m:36-46)."""
    mpa = get_mpa("single_speed", non_convex=False)
    _, b = road_records("single_speed", 4)
    b = dataclasses.replace(b, checker=CHECKER_SAT)
    check(planner, mpa, b)


def test_pop_trace_identical(planner):
    """The full pq.pop() sequence (GraphSearch.m:55), not only its hash."""
    mpa, batch = road_records("triple_speed", 8)
    planner.upload_mpa(mpa)
    planner.stage(batch)
    ref_pops = oracle_py.plan_batch(mpa, batch).n_pops
    try:
        for tile in VARIANTS:
            planner.set_variant(tile)
            for si in (0, int(np.argmax(ref_pops)), batch.n - 1):
                want = oracle_py.plan_trace(mpa, batch, si)
                got = planner.trace(si, cap=want.size + 8)
                assert np.array_equal(got, want), (tile, si)
    finally:
        planner.set_variant(0)


@pytest.mark.parametrize("variant", [2, 3])
@pytest.mark.parametrize("points", [0, 40, 90, 150])
def test_tile_shape_unstaged_polylines(planner, variant, points):
    """Shapes 2, 3: polylines that do not fit a tile's staged points are read from HBM instead (40: neither
    lanelet bounds nor obstacles staged for most searches; 90 / 150: lanelets staged, obstacles of the crowded
    searches not).  Any split must give the oracle's results, and the requested shape must really run."""
    mpa, batch = road_records("triple_speed", 8)
    planner.set_tile_points(points)
    try:
        info, dev, ref = check(planner, mpa, batch, variants=(variant,))
        assert info["n"] == batch.n
        st = planner.stats()
        assert st.shape == variant
        assert st.total_pops == int(ref.n_pops.sum()) and st.total_nodes == int(ref.n_expanded.sum())
    finally:
        planner.set_tile_points(0)


@pytest.mark.parametrize("entries", [2, 16, 64, 0])
def test_cta_shape_heap_spills_to_arena(planner, entries):
    """Shape 4: whatever part of the queue lives in shared memory (two-level look-ahead walk)
    or in the HBM arena (single-level walk), the pop order is the oracle's."""
    mpa, batch = road_records("triple_speed", 8)
    planner.set_cta_heap_smem(entries)
    try:
        info, dev, ref = check(planner, mpa, batch, variants=(4,))
        st = planner.stats()
        assert st.total_pops == int(ref.n_pops.sum()) and st.total_nodes == int(ref.n_expanded.sum())
    finally:
        planner.set_cta_heap_smem(0)


def test_cta_shape_is_the_default_for_one_level(planner):
    """A batch with at most one search per SM (one computation level of a time step) takes the
    CTA-per-search shape automatically; many CTAs' worth of searches (more than SMs) still work
    when the shape is forced (each CTA pulls several searches one after the other)."""
    mpa, batch = road_records("single_speed", 6)
    check(planner, mpa, batch.select(np.arange(20)), variants=(0,))
    check(planner, mpa, batch, variants=(4,))
    mpa_c, batch_c = circle_records(30)
    check(planner, mpa_c, batch_c, variants=(4,))


def test_lane_shape_single_speed_and_realistic(planner):
    """5-/6-/7-point maneuver areas all run the padded 6-edge InterX of shape 3."""
    for mpa_type, kw in (("single_speed", {}), ("realistic", dict(amount=10, seed=5))):
        mpa, batch = road_records(mpa_type, 3, **kw)
        check(planner, mpa, batch, variants=(3,))


def test_empty_batch(planner):
    mpa = get_mpa("single_speed", non_convex=False)
    planner.upload_mpa(mpa)
    b = SearchBatch.from_iters([], mpa.Hp, CHECKER_SAT, mpa.dt_seconds)
    r = planner.plan_batch(b)
    assert r.status.size == 0


@pytest.mark.parametrize("checker", [CHECKER_SAT, CHECKER_INTERX])
def test_free_space_straight_plan(planner, checker):
    """No obstacles, no lanelets: the search walks straight down the tree."""
    mpa = get_mpa("single_speed", non_convex=(checker == CHECKER_INTERX))
    b = SearchBatch.from_iters([straight_iter(mpa)], mpa.Hp, checker, mpa.dt_seconds)
    info, dev, ref = check(planner, mpa, b)
    assert not dev.is_exhausted[0] and dev.n_pops[0] >= mpa.Hp + 1
    assert dev.trims[0, -1] == 1     # recursive feasibility: last trim is the equilibrium


@pytest.mark.parametrize("checker", [CHECKER_SAT, CHECKER_INTERX])
def test_blocked_world_exhausts_with_countable_tree(planner, checker):
    """A static obstacle on top of the vehicle invalidates every depth-1 node:
    pops = 1 + #children(root), tree = root + children, is_exhausted."""
    mpa = get_mpa("single_speed", non_convex=(checker == CHECKER_INTERX))
    # SAT: a big box covering everything; InterX: only crossings count, so use a thin
    # long sliver across the vehicle's footprint
    obs = rect(0.0, 0.0, 5.0, 5.0) if checker == CHECKER_SAT else rect(0.05, 0.0, 0.001, 3.0)
    it = straight_iter(mpa, obstacles=[obs])
    b = SearchBatch.from_iters([it], mpa.Hp, checker, mpa.dt_seconds)
    info, dev, ref = check(planner, mpa, b)
    n_children = int(mpa.transition[0, it.trim_indices - 1].sum())
    assert dev.is_exhausted[0] == 1
    assert dev.n_expanded[0] == 1 + n_children
    assert dev.n_pops[0] == 1 + n_children
    assert np.isnan(dev.y_predicted[0]).all() and (dev.trims[0, 1:] == 0).all()


def test_dynamic_obstacle_only_at_its_step(planner):
    """An obstacle at step 3 only (GraphSearch.m:147,176: index = child's depth)."""
    mpa = get_mpa("single_speed", non_convex=False)
    Hp = mpa.Hp
    far = rect(50, 50, 0.1, 0.1)
    row = [far] * Hp
    row[2] = rect(0.45, 0.0, 0.05, 1.0)
    it = straight_iter(mpa, dynamic_obstacle_area=[row])
    b = SearchBatch.from_iters([it], Hp, CHECKER_SAT, mpa.dt_seconds)
    check(planner, mpa, b)


def test_heap_overflow_beyond_shared_memory(planner):
    """Exhausted road searches hold far more open nodes than the shared-memory heap
    top (256 entries): exercises the HBM heap overflow."""
    mpa, batch = road_records("triple_speed", 20)
    ref = oracle_py.plan_batch(mpa, batch)
    order = np.argsort(ref.n_expanded)[::-1][:6]
    assert ref.n_expanded[order[0]] > 2000
    check(planner, mpa, batch.select(order))


def test_capacity_error_is_loud(planner):
    mpa, batch = road_records("triple_speed", 8)
    ref = oracle_py.plan_batch(mpa, batch)
    big = int(np.argmax(ref.n_expanded))
    planner.upload_mpa(mpa)
    planner.set_node_capacity(256)
    try:
        r = planner.plan_batch(batch.select([big, 0]), raise_on_search_error=False)
        assert r.status[0] == capi.PDMPC_ERR_CAPACITY and r.is_exhausted[0] == 1
        with pytest.raises(capi.PdmpcError):
            planner.plan_batch(batch.select([big]))
    finally:
        planner.set_node_capacity(0)
    check(planner, mpa, batch.select([big, 0]))


def test_bad_inputs_rejected(planner):
    mpa, batch = road_records("triple_speed", 8)
    planner.upload_mpa(mpa)
    bad = dataclasses.replace(batch, trim0=np.zeros_like(batch.trim0))
    with pytest.raises(capi.PdmpcError) as e:
        planner.plan_batch(bad)
    assert e.value.code == capi.PDMPC_ERR_BAD_INPUT
    vx = batch.vert_x.copy()
    vx[0] += 1.0        # first polygon no longer closed (vectorize_all_obstacles.m:71-76)
    with pytest.raises(capi.PdmpcError):
        planner.plan_batch(dataclasses.replace(batch, vert_x=vx))
    fresh = capi.Planner(0)
    with pytest.raises(capi.PdmpcError) as e:
        fresh.plan_batch(batch)
    assert e.value.code == capi.PDMPC_ERR_NO_MPA
    fresh.close()


def test_short_and_long_horizon(planner):
    for Hp in (3, 8):
        mpa = build_mpa("single_speed", Hp=Hp, non_convex=False)
        its = [straight_iter(mpa), straight_iter(mpa, obstacles=[rect(0.6, 0.0, 0.05, 0.3)])]
        b = SearchBatch.from_iters(its, Hp, CHECKER_SAT, mpa.dt_seconds)
        check(planner, mpa, b)


def test_staged_path_equals_plan_batch_and_is_deterministic(planner):
    mpa, batch = road_records("triple_speed", 8)
    planner.upload_mpa(mpa)
    a = planner.plan_batch(batch)
    planner.stage(batch)
    planner.run_staged()
    planner.run_staged()
    b = planner.fetch()
    parity.compare(a, b)
    st = planner.stats()
    assert st.total_pops == int(a.n_pops.sum()) and st.total_nodes == int(a.n_expanded.sum())


def test_full_size_properties(planner):
    """BASELINE-size batch (tiled road records): per-search results are independent of
    batch position and of which CTA/slot ran them (permutation + concatenation invariance)."""
    mpa, batch = road_records("triple_speed", 8)
    rng = np.random.default_rng(0)
    reps = 60
    big = SearchBatch.concat([batch] * reps)            # ~9600 searches, > resident CTA slots
    perm = rng.permutation(big.n)
    shuffled = big.select(perm)
    planner.upload_mpa(mpa)
    r0 = planner.plan_batch(batch)
    r1 = planner.plan_batch(shuffled)
    for name in ("status", "is_exhausted", "n_expanded", "n_pops", "pop_hash"):
        base = np.tile(getattr(r0, name), reps)
        assert np.array_equal(getattr(r1, name), base[perm]), name
    assert np.array_equal(r1.trims, np.tile(r0.trims, (reps, 1))[perm])
    assert np.array_equal(r1.y_predicted.view(np.uint64), np.tile(r0.y_predicted, (reps, 1, 1))[perm].view(np.uint64))


def test_packed_and_unpacked_staging_agree(planner):
    """Batches of at most 1 MiB go through one pinned staging block and one copy each way, larger ones
    through per-array copies: the same searches must give the same outputs either way, and in any
    mix of calls on one handle."""
    mpa, batch = road_records("triple_speed", 8)
    planner.upload_mpa(mpa)
    small = batch.select(np.arange(12))
    assert small.input_bytes() < (1 << 20)
    big = SearchBatch.concat([batch] * 6)
    assert big.input_bytes() > (1 << 20)
    ref = oracle_py.plan_batch(mpa, batch, 4)
    a = planner.plan_batch(small)
    b = planner.plan_batch(big)
    c = planner.plan_batch(small)
    first = lambda r, n: dataclasses.replace(r, **{f.name: getattr(r, f.name)[:n] for f in dataclasses.fields(r)
                                                   if isinstance(getattr(r, f.name), np.ndarray)})
    parity.compare(a, first(ref, 12))
    parity.compare(c, first(ref, 12))
    parity.compare(first(b, batch.n), ref)
    parity.compare(dataclasses.replace(b, **{f.name: getattr(b, f.name)[-batch.n:] for f in dataclasses.fields(b)
                                             if isinstance(getattr(b, f.name), np.ndarray)}), ref)


def test_restage_before_fetch_and_knob_validation(planner):
    """The pinned staging block is reused by the next stage call: staging again while the previous
    copy may still be in flight must not corrupt either batch."""
    mpa, batch = road_records("single_speed", 6)
    planner.upload_mpa(mpa)
    a, b = batch.select(np.arange(0, 10)), batch.select(np.arange(10, 30))
    planner.stage(a)
    planner.run_staged()
    planner.stage(b)          # no fetch in between
    planner.run_staged()
    parity.compare(planner.fetch(), oracle_py.plan_batch(mpa, b))
    planner.stage(a)
    planner.run_staged()
    parity.compare(planner.fetch(), oracle_py.plan_batch(mpa, a))
    for bad in (3, 5000, -2):
        with pytest.raises(capi.PdmpcError):
            planner.set_cta_heap_smem(bad)
    with pytest.raises(capi.PdmpcError):
        planner.set_variant(9)
    planner.set_variant(0)


@pytest.mark.parametrize("chunks", [2, 3, 7, 16])
def test_pipelined_plan_batch_equals_single_launch(planner, chunks):
    """Large host batches are cut into chunks whose copies and searches overlap
    (pdmpc_set_pipeline_chunks): same outputs for any chunk count, both warp launch shapes, both
    checkers, and the device is left holding the whole batch (run_staged / fetch afterwards)."""
    cases = [road_records("triple_speed", 8), circle_records(30)]
    try:
        for mpa, batch in cases:
            planner.upload_mpa(mpa)
            ref = oracle_py.plan_batch(mpa, batch, 4)
            for variant in (0, 1, 2):
                planner.set_variant(variant)
                planner.set_pipeline_chunks(chunks)
                dev = planner.plan_batch(batch, raise_on_search_error=False)
                parity.compare(dev, ref)
                st = planner.stats()
                assert st.total_pops == int(ref.n_pops.sum()) and st.total_nodes == int(ref.n_expanded.sum())
                assert st.kernel_launches >= chunks
            planner.set_variant(0)
            planner._staged_n, planner._staged_Hp = batch.n, batch.Hp
            planner.run_staged()
            parity.compare(planner.fetch(), ref)
            planner.set_pipeline_chunks(1)
            parity.compare(planner.plan_batch(batch, raise_on_search_error=False), ref)
        # malformed input in a LATER chunk is still rejected, and the handle survives
        mpa, batch = cases[0]
        planner.upload_mpa(mpa)
        planner.set_pipeline_chunks(chunks)
        bad = dataclasses.replace(batch, trim0=batch.trim0.copy())
        bad.trim0[-1] = 99
        with pytest.raises(capi.PdmpcError) as e:
            planner.plan_batch(bad)
        assert e.value.code == capi.PDMPC_ERR_BAD_INPUT
        parity.compare(planner.plan_batch(batch, raise_on_search_error=False), oracle_py.plan_batch(mpa, batch, 4))
    finally:
        planner.set_variant(0)
        planner.set_pipeline_chunks(0)
    with pytest.raises(capi.PdmpcError):
        planner.set_pipeline_chunks(17)


@pytest.mark.parametrize("variant", [2, 3])
@pytest.mark.parametrize("valid_only", [False, True])
def test_escalation_of_long_searches(planner, variant, valid_only):
    """pdmpc_set_escalation: the tile shapes give searches beyond a pop threshold up and the CTA shape runs them
    behind the tile kernel (both gated instances: a short list -> one master per CTA, a long list -> several).
    Same outputs as without escalation and as the oracle, staged and pipelined; pdmpc_set_cta_queue(1) makes
    pop_hash cover the valid pops in every shape, so the batch reports one kind of hash."""
    mpa, batch = road_records("triple_speed", 8)
    planner.upload_mpa(mpa)
    ref = oracle_py.plan_batch(mpa, batch, 4, hash_valid_pops_only=valid_only)
    pops = ref.n_pops.astype(np.int64)
    try:
        planner.set_cta_queue(valid_only)
        planner.set_variant(variant)
        for thr, short in ((int(np.percentile(pops, 99.5)), -1), (int(np.percentile(pops, 70)), 10), (8, -1), (8, 20)):
            # given up after its thr-th pop unless that pop was the goal
            expect = int(((pops > thr) | ((pops == thr) & (ref.is_exhausted != 0))).sum())
            planner.set_escalation(thr, short)   # short: lists up to that long -> one master per CTA, else several
            planner.set_pipeline_chunks(1)
            dev = planner.plan_batch(batch, raise_on_search_error=False)
            st = planner.stats()
            parity.compare(dev, ref)
            assert st.shape == variant and st.escalated == expect, (thr, st.escalated, expect)
            assert st.total_pops == int(pops.sum()) and st.total_nodes == int(ref.n_expanded.sum())
            planner.set_pipeline_chunks(3)
            dev = planner.plan_batch(batch, raise_on_search_error=False)
            st = planner.stats()
            parity.compare(dev, ref)
            assert st.escalated == expect
            assert st.total_pops == int(pops.sum()) and st.total_nodes == int(ref.n_expanded.sum())
        planner.set_escalation(0)
        dev = planner.plan_batch(batch, raise_on_search_error=False)
        parity.compare(dev, ref)
        assert planner.stats().escalated == 0
        # every warp shape reports the same kind of hash
        planner.set_variant(1)
        parity.compare(planner.plan_batch(batch, raise_on_search_error=False), ref)
    finally:
        planner.set_variant(0)
        planner.set_pipeline_chunks(0)
        planner.set_cta_queue(False)
        planner.set_escalation(-1)           # back to the default: by batch size
    with pytest.raises(capi.PdmpcError):
        planner.set_escalation(-2)


def _with_areas(mpa, transform):
    """A copy of `mpa` whose maneuver areas (all three kinds) went through transform(x[:n], y[:n]) -> (x', y')."""
    npts, ax, ay = mpa.area_npts.copy(), np.zeros_like(mpa.area_x), np.zeros_like(mpa.area_y)
    for e in range(mpa.n_edges):
        for k in range(3):
            n = int(mpa.area_npts[e, k])
            x, y = transform(mpa.area_x[e, k, :n].copy(), mpa.area_y[e, k, :n].copy())
            npts[e, k] = x.size
            ax[e, k, :x.size], ay[e, k, :x.size] = x, y
    return dataclasses.replace(mpa, area_npts=npts, area_x=ax, area_y=ay)


@pytest.mark.gpu
def test_open_and_eight_point_maneuver_areas(planner):
    """The InterX loop instances no reference MPA reaches: areas that are NOT closed polygons (the closing vertex
    dropped: the C2 row then computes its last vertex instead of reusing the first) and areas of 8 points (7 edges, the
    widest row; only plan_batch takes them, the hand-over of a time step needs a free column).  Every launch shape
    against the oracle on road-network records."""
    base, batch = road_records("single_speed", 3)
    batch = batch.select(np.arange(min(batch.n, 120)))
    open_mpa = _with_areas(base, lambda x, y: (x[:-1], y[:-1]))
    assert int(open_mpa.area_npts.max()) == int(base.area_npts.max()) - 1

    def eight(x, y):   # one more vertex in the middle of the longest-index edge of a 7-point area: same polygon, 7 edges
        if x.size != 7:
            return x, y
        return (np.concatenate([x[:1], [(x[0] + x[1]) / 2], x[1:]]), np.concatenate([y[:1], [(y[0] + y[1]) / 2], y[1:]]))

    wide_mpa = _with_areas(base, eight)
    assert int(wide_mpa.area_npts.max()) == 8
    for mpa in (open_mpa, wide_mpa):
        info, dev, ref = check(planner, mpa, batch)
        assert info["n"] == batch.n and info["pops"] > 0
