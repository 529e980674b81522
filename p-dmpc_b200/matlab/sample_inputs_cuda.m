function hlcs = sample_inputs_cuda(handle, hlcs, scenario, mpa, options, states)
% SAMPLE_INPUTS_CUDA  Reference trajectories and lanelet boundaries of ALL vehicles of a time step in ONE call on the GPU.
%
%   Replaces, per vehicle, the calls HighLevelController makes before the optimizer runs
%   (hlc/controller/common):
%       get_reference_trajectory.m:28-46   -> iter.reference_trajectory_points, iter.v_ref
%       sample_reference_trajectory.m:24-97, get_arc_distance_to_endpoint.m:41-113, projection_2d.m
%       get_predicted_lanelets.m:25-62     -> iter.predicted_lanelets
%       get_lanelets_boundary.m:19-68      -> iter.predicted_lanelet_boundary{:, 1:2}  (the polyshape in cell 3 is
%                                             rebuilt below: the search does not read it)
%
%   handle   : planner handle of pdmpc_b200_mex (GraphSearchCuda.handle)
%   hlcs     : 1 x N controllers whose iter is filled
%   scenario : Scenario with vehicles(i).reference_path_struct (path, lanelets_index, points_index), reference_speed,
%              and road_raw_data / lanelet_boundaries
%   states   : N x 3 current (x, y, trim) of the vehicles
%
%   The road tables are uploaded once per scenario (persistent flag), like the MPA tables.

    UPLOAD_ROAD = 8; SAMPLE_INPUTS = 9;
    persistent uploaded_for
    N = numel(hlcs);

    if isempty(uploaded_for) || uploaded_for ~= handle
        lb = vertcat(scenario.lanelet_boundaries{:});           % nL x 3 cell -> left, right as n x 2
        paths = arrayfun(@(v) v.reference_path, scenario.vehicles, UniformOutput = false);
        lidx = arrayfun(@(v) v.lanelets_index, scenario.vehicles, UniformOutput = false);
        pidx = arrayfun(@(v) v.points_index, scenario.vehicles, UniformOutput = false);
        is_loop = arrayfun(@(v) double(v.is_loop), scenario.vehicles);
        speed = arrayfun(@(v) v.reference_speed, scenario.vehicles);
        pdmpc_b200_mex(UPLOAD_ROAD, handle, lb(:, 1:2), paths(:), lidx(:), pidx(:), is_loop, speed);
        uploaded_for = handle;
    end

    speed_now = arrayfun(@(i) mpa.trims(states(i, 3)).speed, 1:N);
    [ref, v_ref, points_index, current_index, predicted, left, right] = pdmpc_b200_mex(SAMPLE_INPUTS, handle, ...
        1:N, states(:, 1), states(:, 2), speed_now, options.dt_seconds);

    for i = 1:N
        iter = hlcs(i).iter;
        iter.reference_trajectory_points(i, :, :) = ref(i, :, :);
        iter.reference_trajectory_index(i, :, :) = points_index(i, :);
        iter.v_ref(i, :) = v_ref(i, :);
        iter.predicted_lanelets{i} = predicted{i};
        x = [left{i}(1, :), right{i}(1, end:-1:1)];
        y = [left{i}(2, :), right{i}(2, end:-1:1)];
        iter.predicted_lanelet_boundary(i, :) = {left{i}, right{i}, polyshape(x, y, Simplify = false)};   % :62-67
        hlcs(i).iter = iter;
        hlcs(i).current_point_index = current_index(i); %#ok<*AGROW>
    end

end
