function [obstacles, dynamic_obstacle_area] = assemble_obstacles_cuda(handle, iter, mpa, directed_coupling, directed_coupling_sequential, vehicles, options)
% ASSEMBLE_OBSTACLES_CUDA  The obstacles of ALL vehicles of a time step that do not depend on this time step's plans,
% in ONE call on the GPU (pdmpc_assemble_obstacles).
%
%   Replaces, per vehicle i, in PrioritizedController.plan (hlc/controller/prioritized/PrioritizedController.m):
%       consider_successors :508-540 with ConstraintFromSuccessor.area_of_standstill
%                                 -> obstacles{i}: iter.occupied_areas{j}.normal_offset (get_occupied_areas.m:19-25) of
%                                    every coupled vehicle j of lower priority with |speed| < 0.01
%       consider_predecessors :449-506 -> parallel_coupling_reachability :391-407
%                                 -> dynamic_obstacle_area{i}: iter.reachable_sets(j, :) of every coupled vehicle j of
%                                    higher priority that plans in parallel (directed_coupling(j, i) set,
%                                    directed_coupling_sequential(j, i) not set), i.e.
%                                    mpa.reachable_sets_at_pose (MotionPrimitiveAutomaton.m:649-687)
%   The areas of SEQUENTIAL predecessors are handed over on the device by plan_timestep_cuda.m.
%   Not done here: HighLevelController.m:241-246 clips a reachable set by the predicted lanelets (polyshape
%   intersection); pass pre-clipped local sets or keep that step in MATLAB if it matters for the scenario.
%
%   handle            : planner handle of pdmpc_b200_mex (GraphSearchCuda.handle), MPA uploaded
%   iter              : IterationData of the time step (x0 [N x 4], trim_indices [N])
%   directed_coupling : N x N, (j, i) set = j has the higher priority
%   vehicles          : scenario.vehicles (Length, Width); options.offset
%   obstacles{i}            : {s_i x 1} cell of 2 x 5 areas   -> append to iter_v.obstacles
%   dynamic_obstacle_area{i}: {p_i x Hp} cell of 2 x m areas  -> append to iter_v.dynamic_obstacle_area

    UPLOAD_REACHABLE_SETS = 10; ASSEMBLE_OBSTACLES = 11;
    persistent uploaded_for
    N = size(iter.x0, 1);

    if isempty(uploaded_for) || uploaded_for ~= handle
        % closed 2 x m polygons, as HighLevelController.m:258-264 turns polyshapes into plain arrays
        sets = cellfun(@(c) [c.Vertices(:, 1)', c.Vertices(1, 1); c.Vertices(:, 2)', c.Vertices(1, 2)], ...
            mpa.local_reachable_sets_conv, UniformOutput = false);
        pdmpc_b200_mex(UPLOAD_REACHABLE_SETS, handle, sets);
        uploaded_for = handle;
    end

    successors = arrayfun(@(i) find(directed_coupling(i, :)), 1:N, UniformOutput = false);
    parallel = arrayfun(@(i) find(directed_coupling(:, i)' & ~directed_coupling_sequential(:, i)'), 1:N, UniformOutput = false);
    [obstacles, dynamic_obstacle_area] = pdmpc_b200_mex(ASSEMBLE_OBSTACLES, handle, iter.x0(:, 1:4), iter.trim_indices(:), ...
        successors(:), parallel(:), vehicles(1).Length / 2 + options.offset, vehicles(1).Width / 2 + options.offset);

end
