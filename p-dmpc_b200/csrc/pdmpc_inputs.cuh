// pdmpc_inputs.cuh — the input side of a time step on the device: for every vehicle the
// reference trajectory over the horizon and the lanelet boundary its plan must stay in.
//
//   get_reference_trajectory        hlc/controller/common/get_reference_trajectory.m:28-46
//   sample_reference_trajectory     hlc/controller/common/sample_reference_trajectory.m:24-97
//   get_arc_distance_to_endpoint    hlc/controller/common/get_arc_distance_to_endpoint.m:41-113 (closest point, idx_next)
//   projection_2d                   hlc/controller/common/projection_2d.m:1-27
//   get_predicted_lanelets          hlc/controller/common/get_predicted_lanelets.m:25-62
//   get_lanelets_boundary           hlc/controller/common/get_lanelets_boundary.m:19-68 (left / right bound; the polyshape
//                                   in cell 3 is not read by the search)
//
// Arithmetic: the reference's expressions in the reference's order, norm([a b], 2) taken as sqrt(a*a + b*b), no FMA —
// the specification p-dmpc_b200/scenario.py restates on the host (sample_reference_trajectory, _closest_point,
// get_predicted_lanelets, get_lanelets_boundary) and the parity tests compare bit for bit.
//
// Road and reference paths are uploaded once per scenario set (pdmpc_upload_road); a call handles any number of
// (path, pose, speed) rows: one warp per row.  The closest-point search runs over the lanes; the sampling itself is
// a short serial walk along the path (Hp steps) done by lane 0; three kernels: sample + count, scan, copy bounds.
#pragma once

#include "pdmpc_kernels.cuh"

namespace pdmpc {

constexpr int kMaxPredLanelets = PDMPC_MAX_PRED_LANELETS;

struct RoadDev {
    int n_lanelets, n_paths;
    const int *bound_ptr;              // [2 * n_lanelets + 1]
    const double *bound_x, *bound_y;
    const int *path_ptr;               // [n_paths + 1]
    const double *path_x, *path_y;
    const int *lan_ptr;                // [n_paths + 1]
    const int *lanelets_index, *points_index;
    const unsigned char *is_loop;
    const double *reference_speed;
};

struct InputsDev {
    int n, Hp;
    double dt;
    const int *path_id;                // [n]
    const double *x, *y, *speed;       // [n] position and current speed (mpa.trims(trim).speed)
    double *ref_x, *ref_y, *v_ref;     // [n * Hp]
    int *ref_index, *current_index;    // [n * Hp], [n]
    int *pred_lanelets;                // [n * kMaxPredLanelets], 0 padded
    int *pre_lanelet;                  // [n] predecessor lanelet whose tail is prepended (0: none)
    int *lane_cnt;                     // [2n] points of the left / right bound
    int *lane_ptr;                     // [2n + 1]
    double *lane_x, *lane_y;
};

__device__ __forceinline__ double norm2_ref(double dx, double dy) { return sqrt(dx * dx + dy * dy); }

__global__ void __launch_bounds__(128) sample_inputs_kernel(RoadDev rd, InputsDev in) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x % kWarp;
    const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) / kWarp);
    if (row >= in.n) return;
    const int pid = in.path_id[row];
    const int p0 = rd.path_ptr[pid], n_pts = rd.path_ptr[pid + 1] - p0;
    const double *__restrict__ px = rd.path_x + p0;
    const double *__restrict__ py = rd.path_y + p0;
    const double x = in.x[row], y = in.y[row];
    // ---- get_arc_distance_to_endpoint.m:41-47: first minimum of the squared distances --------------------
    double best = 0.0;
    int ibest = -1;
    for (int i = lane; i < n_pts; i += kWarp) {
        const double dx = px[i] - x, dy = py[i] - y;
        const double d2 = dx * dx + dy * dy;
        if (ibest < 0 || d2 < best) { best = d2; ibest = i; }     // strided: a lane sees its indices ascending
    }
    for (int d = 16; d > 0; d >>= 1) {
        const double ob = __shfl_xor_sync(FULL, best, d);
        const int oi = __shfl_xor_sync(FULL, ibest, d);
        if (oi >= 0 && (ibest < 0 || ob < best || (ob == best && oi < ibest))) { best = ob; ibest = oi; }
    }
    if (lane != 0) return;
    const int Hp = in.Hp;
    const int ic = ibest + 1;                                      // 1-based idx_closest
    auto d2_at = [&](int i1) {                                     // 1-based
        const double dx = px[i1 - 1] - x, dy = py[i1 - 1] - y;
        return dx * dx + dy * dy;
    };
    int a, b;
    if (ic == 1) { a = 1; b = 2; }
    else if (ic == n_pts) { a = n_pts - 1; b = n_pts; }
    else if (d2_at(ic - 1) <= d2_at(ic + 1)) { a = ic - 1; b = ic; }   // min([left, right]): the left one on a tie
    else { a = ic; b = ic + 1; }
    // ---- projection_2d.m ------------------------------------------------------------------------------
    double cx, cy, lam;
    {
        const double x1 = px[a - 1], y1 = py[a - 1], x2 = px[b - 1], y2 = py[b - 1];
        const double bb = norm2_ref(x2 - x1, y2 - y1);
        if (bb != 0) {
            const double xn = (x2 - x1) / bb, yn = (y2 - y1) / bb;
            const double x31 = x - x1, y31 = y - y1;
            const double dot = xn * x31 + yn * y31;
            cx = x1 + dot * xn;
            cy = y1 + dot * yn;
            lam = dot / bb;
        } else { cx = x1; cy = y1; lam = 0.0; }
    }
    int point_index = ic;                                          // idx_next, :92-113
    if ((lam >= 0 && lam <= 0.5) || lam >= 1) point_index = ic < n_pts ? ic + 1 : 1;
    point_index = max(2, point_index);
    in.current_index[row] = point_index;
    // ---- sample_reference_trajectory.m:40-97 ----------------------------------------------------------
    const bool is_loop = norm2_ref(px[0] - px[n_pts - 1], py[0] - py[n_pts - 1]) < 1e-8;
    int point_index_last = point_index - 1;
    if (is_loop && point_index == n_pts) point_index = 1;
    const double v_reference = rd.reference_speed[pid];
    double v_prev = in.speed[row];                                 // get_reference_trajectory.m:35-40
    int ref_idx[kMaxHp];
    for (int i = 0; i < Hp; ++i) {
        const double step = ((v_prev + v_reference) / 2) * in.dt;
        v_prev = v_reference;
        double remaining = norm2_ref(cx - px[point_index - 1], cy - py[point_index - 1]);
        if (remaining > step || point_index == n_pts) {
            while (px[point_index - 1] == px[point_index_last - 1] && py[point_index - 1] == py[point_index_last - 1] &&
                   point_index_last > 1)
                --point_index_last;
            const double ux = px[point_index - 1] - px[point_index_last - 1], uy = py[point_index - 1] - py[point_index_last - 1];
            const double nrm = norm2_ref(ux, uy);
            cx = cx + step * (ux / nrm);
            cy = cy + step * (uy / nrm);
        } else {
            double reflength = remaining;
            while (remaining < step) {
                reflength = remaining;
                cx = px[point_index - 1];
                cy = py[point_index - 1];
                point_index_last = point_index;
                point_index = min(point_index + 1, n_pts);
                if (is_loop && point_index == n_pts) point_index = 1;
                remaining = remaining + norm2_ref(cx - px[point_index - 1], cy - py[point_index - 1]);
            }
            const double ux = px[point_index - 1] - px[point_index_last - 1], uy = py[point_index - 1] - py[point_index_last - 1];
            const double nrm = norm2_ref(ux, uy);
            cx = cx + (step - reflength) * (ux / nrm);
            cy = cy + (step - reflength) * (uy / nrm);
        }
        in.ref_x[(size_t)row * Hp + i] = cx;
        in.ref_y[(size_t)row * Hp + i] = cy;
        in.v_ref[(size_t)row * Hp + i] = v_reference;
        in.ref_index[(size_t)row * Hp + i] = point_index;
        ref_idx[i] = point_index;
    }
    // ---- get_predicted_lanelets.m:25-62 ---------------------------------------------------------------
    const int l0 = rd.lan_ptr[pid], n_lan = rd.lan_ptr[pid + 1] - l0;
    const int *__restrict__ lanelets_index = rd.lanelets_index + l0;
    const int *__restrict__ points_index = rd.points_index + l0;
    int pred[kMaxPredLanelets], n_pred = 0;
    for (int q = 0; q <= Hp; ++q) {
        int idx;
        if (q < Hp) idx = ref_idx[q];
        else {
            idx = ref_idx[Hp - 1] + 4;
            if (idx > n_pts) idx -= n_pts;
        }
        int pos = 1;                                               // sum(idx > reference_path_points_index) + 1
        for (int j = 0; j < n_lan; ++j) pos += idx > points_index[j] ? 1 : 0;
        bool seen = false;                                         // unique(..., 'stable')
        for (int j = 0; j < n_pred; ++j) seen = seen || pred[j] == pos;
        if (!seen && n_pred < kMaxPredLanelets) pred[n_pred++] = pos;
    }
    if (n_pred == 1) {
        int nxt = pred[0] + 1;
        if (nxt > n_lan) nxt = 1;
        pred[n_pred++] = nxt;
    }
    for (int j = 0; j < kMaxPredLanelets; ++j)
        in.pred_lanelets[(size_t)row * kMaxPredLanelets + j] = j < n_pred ? lanelets_index[pred[j] - 1] : 0;
    // ---- get_lanelets_boundary.m:19-68: sizes -----------------------------------------------------------
    const int first_lanelet = lanelets_index[pred[0] - 1];
    int pos_first = 0;
    while (pos_first < n_lan && lanelets_index[pos_first] != first_lanelet) ++pos_first;
    int pre = 0;
    if (pos_first != 0) pre = lanelets_index[pos_first - 1];
    else if (rd.is_loop[pid]) pre = lanelets_index[n_lan - 1];
    int nL = 1, nR = 1;
    for (int j = 0; j < n_pred; ++j) {
        const int l = lanelets_index[pred[j] - 1] - 1;
        nL += rd.bound_ptr[2 * l + 1] - rd.bound_ptr[2 * l] - 1;
        nR += rd.bound_ptr[2 * l + 2] - rd.bound_ptr[2 * l + 1] - 1;
    }
    if (pre) {
        const int l = pre - 1;
        const int k = min(4, min(rd.bound_ptr[2 * l + 2] - rd.bound_ptr[2 * l + 1] - 1, rd.bound_ptr[2 * l + 1] - rd.bound_ptr[2 * l] - 1));
        nL += k;
        nR += k;
    }
    in.pre_lanelet[row] = pre;
    in.lane_cnt[2 * row] = nL;
    in.lane_cnt[2 * row + 1] = nR;
}

// lane_ptr = exclusive prefix sum of lane_cnt (2n entries), one block
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int *cnt, int *ptr, int m) {
    __shared__ int part[1024];
    const int t = threadIdx.x, per = (m + 1023) / 1024;
    const int lo = min(t * per, m), hi = min(lo + per, m);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += cnt[i];
    part[t] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const int v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int run = t ? part[t - 1] : 0;
    for (int i = lo; i < hi; ++i) {
        ptr[i] = run;
        run += cnt[i];
    }
    if (t == 1023) ptr[m] = part[1023];
}

// get_lanelets_boundary.m:28-62: [tail of the predecessor][every predicted lanelet without its last point][the last point]
__global__ void __launch_bounds__(128) copy_bounds_kernel(RoadDev rd, InputsDev in, int lane_capacity) {
    const int lane = threadIdx.x % kWarp;
    const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) / kWarp);
    if (row >= in.n) return;
    if (in.lane_ptr[2 * in.n] > lane_capacity) return;            // reported by the host, nothing is written
    const int pre = in.pre_lanelet[row];
    for (int side = 0; side < 2; ++side) {
        int w = in.lane_ptr[2 * row + side];
        if (pre) {
            const int l = pre - 1;
            const int lenL = rd.bound_ptr[2 * l + 1] - rd.bound_ptr[2 * l], lenR = rd.bound_ptr[2 * l + 2] - rd.bound_ptr[2 * l + 1];
            const int k = min(4, min(lenR - 1, lenL - 1));
            const int base = rd.bound_ptr[2 * l + side], len = side ? lenR : lenL;
            for (int j = lane; j < k; j += kWarp) {                // (end - k : end - 1)
                in.lane_x[w + j] = rd.bound_x[base + len - 1 - k + j];
                in.lane_y[w + j] = rd.bound_y[base + len - 1 - k + j];
            }
            w += k;
        }
        int last_base = 0, last_len = 0;
        for (int q = 0; q < kMaxPredLanelets; ++q) {
            const int lid = in.pred_lanelets[(size_t)row * kMaxPredLanelets + q];
            if (!lid) break;
            const int base = rd.bound_ptr[2 * (lid - 1) + side], len = rd.bound_ptr[2 * (lid - 1) + side + 1] - base;
            for (int j = lane; j < len - 1; j += kWarp) {
                in.lane_x[w + j] = rd.bound_x[base + j];
                in.lane_y[w + j] = rd.bound_y[base + j];
            }
            w += len - 1;
            last_base = base; last_len = len;
        }
        if (lane == 0) {
            in.lane_x[w] = rd.bound_x[last_base + last_len - 1];
            in.lane_y[w] = rd.bound_y[last_base + last_len - 1];
        }
    }
}

}  // namespace pdmpc
