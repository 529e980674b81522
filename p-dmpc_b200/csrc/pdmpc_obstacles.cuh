// pdmpc_obstacles.cuh — obstacle assembly of a time step on the device (the part that does not
// depend on this time step's plans): what PrioritizedController.plan puts into iter_v.obstacles /
// iter_v.dynamic_obstacle_area before the optimizer runs, for every vehicle of every scenario in one call.
//
//   consider_successors, ConstraintFromSuccessor.area_of_standstill
//                                    hlc/controller/prioritized/PrioritizedController.m:508-540: a coupled vehicle of
//                                    LOWER priority that stands (|speed| < 0.01 m/s) is a static obstacle, its area =
//                                    iter.occupied_areas{j}.normal_offset
//   get_occupied_areas               hlc/controller/common/get_occupied_areas.m:19-25: the closed 5-point rectangle
//                                    [-1 -1 1 1 -1] * (Length/2 + offset), [-1 1 1 -1 -1] * (Width/2 + offset) placed
//                                    at the measured pose by translate_global
//   parallel_coupling_reachability   PrioritizedController.m:391-407 (via consider_predecessors :449-506): a coupled
//                                    vehicle of HIGHER priority that plans in parallel (another group) enters through its
//                                    reachable sets, one dynamic obstacle per step: iter.reachable_sets(j, :)
//   reachable_sets_at_pose           hlc/model/motion_primitive_automaton/MotionPrimitiveAutomaton.m:649-687: the local
//                                    reachable sets of the vehicle's CURRENT trim (local_reachable_sets_conv{trim, t},
//                                    reachability_analysis_offline :252-392, uploaded once) placed at the pose
//   translate_global                 utility/translate_global.m:20-23: x = c*xl - s*yl + x0, y = s*xl + c*yl + y0
//
// (The areas of SEQUENTIAL predecessors are this time step's plans: they are handed over inside the search launch,
// pdmpc_plan_timestep.)  Output = the obstacle CSR of pdmpc_batch_in: slot i*(Hp+1) holds the standstill rectangles of
// row i's standing successors in the caller's order, slot i*(Hp+1)+k the step-k reachable set of each of its parallel
// predecessors in the caller's order — the order plan() appends them in.  Three kernels: counts per slot (and cos/sin
// of every row's yaw, by the arithmetic specification of DESIGN.md §2), two prefix sums, one warp per slot placing its
// polygons.  Pure streaming of tabulated polygons: tens of KB per scenario, launch-bound.
#pragma once

#include "pdmpc_kernels.cuh"

namespace pdmpc {

struct ReachDev {                       // mpa.local_reachable_sets_conv{trim, t}, closed polygons
    int nT = 0, Hp = 0;
    const int *ptr = nullptr;           // [nT * Hp + 1] into x / y
    const double *x = nullptr, *y = nullptr;
};

struct CouplingDev {
    int n, Hp;
    const double *x, *y, *yaw, *speed;  // [n] measured state
    const int *trim;                    // [n] 1-based
    const int *succ_ptr, *succ_idx;     // CSR: coupled rows of lower priority
    const int *par_ptr, *par_idx;       // CSR: coupled rows of higher priority that plan in parallel
    double half_len, half_wid;          // Length / 2 + offset, Width / 2 + offset
    double2 *cs;                        // [n] (cos, sin) of the row's yaw
    int *cnt_poly, *cnt_vert;           // [n * (Hp + 1)]
    int *slot_ptr, *vert_base;          // [n * (Hp + 1) + 1] exclusive prefix sums
    int *poly_ptr;                      // [polygons + 1]
    double *vert_x, *vert_y;
};

constexpr double kStandstillSpeed = 0.01;   // PrioritizedController.m:527 standstill_speed_meter_per_second

// one thread per slot (row i, k = 0: static, k = 1..Hp: dynamic obstacles of step k)
__global__ void count_obstacles_kernel(ReachDev r, CouplingDev c) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= c.n * (c.Hp + 1)) return;
    const int i = s / (c.Hp + 1), k = s % (c.Hp + 1);
    int np = 0, nv = 0;
    if (k == 0) {
        double sn, cs;
        sincos_ref(c.yaw[i], sn, cs);
        c.cs[i] = make_double2(cs, sn);
        for (int q = c.succ_ptr[i]; q < c.succ_ptr[i + 1]; ++q)
            if (fabs(c.speed[c.succ_idx[q]]) < kStandstillSpeed) { ++np; nv += 5; }
    } else {
        for (int q = c.par_ptr[i]; q < c.par_ptr[i + 1]; ++q) {
            const int t = c.trim[c.par_idx[q]] - 1;
            ++np;
            nv += r.ptr[t * r.Hp + k] - r.ptr[t * r.Hp + k - 1];
        }
    }
    c.cnt_poly[s] = np;
    c.cnt_vert[s] = nv;
}

// one warp per slot: its polygons one after the other, the vertices of a polygon across the lanes
__global__ void __launch_bounds__(128) fill_obstacles_kernel(ReachDev r, CouplingDev c, int poly_capacity, int vert_capacity) {
    const int lane = threadIdx.x % kWarp;
    const int s = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) / kWarp);
    const int S = c.n * (c.Hp + 1);
    if (s >= S) return;
    if (c.slot_ptr[S] > poly_capacity || c.vert_base[S] > vert_capacity) return;   // reported by the host
    const int i = s / (c.Hp + 1), k = s % (c.Hp + 1);
    int p = c.slot_ptr[s], v = c.vert_base[s];
    if (k == 0) {
        for (int q = c.succ_ptr[i]; q < c.succ_ptr[i + 1]; ++q) {
            const int j = c.succ_idx[q];
            if (!(fabs(c.speed[j]) < kStandstillSpeed)) continue;
            if (lane == 0) c.poly_ptr[p] = v;
            if (lane < 5) {   // get_occupied_areas.m:21-23
                const double xl = (lane == 2 || lane == 3 ? 1.0 : -1.0) * c.half_len;
                const double yl = (lane == 1 || lane == 2 ? 1.0 : -1.0) * c.half_wid;
                const double2 cs = c.cs[j];
                c.vert_x[v + lane] = cs.x * xl - cs.y * yl + c.x[j];
                c.vert_y[v + lane] = cs.y * xl + cs.x * yl + c.y[j];
            }
            ++p; v += 5;
        }
    } else {
        for (int q = c.par_ptr[i]; q < c.par_ptr[i + 1]; ++q) {
            const int j = c.par_idx[q];
            const int t = c.trim[j] - 1;
            const int a0 = r.ptr[t * r.Hp + k - 1], m = r.ptr[t * r.Hp + k] - a0;
            if (lane == 0) c.poly_ptr[p] = v;
            const double2 cs = c.cs[j];
            const double x0 = c.x[j], y0 = c.y[j];
            for (int u = lane; u < m; u += kWarp) {   // MotionPrimitiveAutomaton.m:672-678
                const double xl = r.x[a0 + u], yl = r.y[a0 + u];
                c.vert_x[v + u] = cs.x * xl - cs.y * yl + x0;
                c.vert_y[v + u] = cs.y * xl + cs.x * yl + y0;
            }
            ++p; v += m;
        }
    }
    if (s == S - 1 && lane == 0) c.poly_ptr[c.slot_ptr[S]] = c.vert_base[S];
}

}  // namespace pdmpc
