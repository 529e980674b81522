classdef GraphSearchCuda < OptimizerInterface
    % GRAPHSEARCHCUDA  B200 drop-in for GraphSearch (hlc/optimizer/graph_search/GraphSearch.m).
    %
    %   Same call surface as every optimizer of the reference
    %   (OptimizerInterface.m:14): info = run_optimizer(obj, veh_index, iter, mpa, options, time_step)
    %   returns a ControlResultsInfo with the fields PrioritizedController.plan and
    %   Simulation.apply consume (ControlResultsInfo.m:5-17).  The search itself runs on the
    %   GPU through the C ABI of include/pdmpc_b200.h, bound by pdmpc_b200_mex.cpp.
    %   There is no CPU fallback: without a CUDA device the constructor errors.
    %
    %   Selected by OptimizerType.CudaOptimal (see INTEGRATION.md for the two-line patch to
    %   config/enums/OptimizerType.m and OptimizerInterface.get_optimizer).

    properties (Constant, Hidden = true)
        CREATE = 0;
        DESTROY = 1;
        UPLOAD_MPA = 2;
        PLAN = 3;
        STATS = 4;
        PLAN_JOINT = 7;
    end

    properties (SetAccess = private, Hidden = true)
        handle = uint64(0); % pdmpc_handle* owned by the MEX
        mpa_key = ''; % which MPA tables are resident in HBM
        checker (1, 1) double = 1; % 0 = SAT, 1 = InterX (OptimizerInterface.m:36-46)
    end

    methods

        function obj = GraphSearchCuda(options, device_id)

            arguments
                options (1, 1) Config
                device_id (1, 1) double = 0
            end

            obj = obj@OptimizerInterface();
            obj.handle = pdmpc_b200_mex(GraphSearchCuda.CREATE, device_id);
            % same selection rule as OptimizerInterface.set_constraint_checker
            obj.checker = double(options.are_any_obstacles_non_convex);
        end

        function delete(obj)

            if obj.handle ~= 0
                pdmpc_b200_mex(GraphSearchCuda.DESTROY, obj.handle);
                obj.handle = uint64(0);
            end

        end

        function info = run_optimizer(obj, ~, iter, mpa, options, ~)
            obj.upload_mpa_once(mpa, options);

            if iter.amount > 1
                % CentralizedController.controller (CentralizedController.m:33-59): one joint search
                info = obj.run_joint(iter, options);
                return
            end

            % iter.predicted_lanelet_boundary{1, 1:2}: left / right bound, 2 x n (may be empty)
            left = zeros(2, 0); right = zeros(2, 0);

            if ~isempty(iter.predicted_lanelet_boundary)
                left = iter.predicted_lanelet_boundary{1, 1};
                right = iter.predicted_lanelet_boundary{1, 2};
            end

            dynamic_obstacle_area = GraphSearchCuda.with_hdv_reachable_sets(iter, obj.checker, options.Hp);

            [is_exhausted, n_expanded, trims, y_pred, g_path, h_path, shapes] = pdmpc_b200_mex( ...
                GraphSearchCuda.PLAN, obj.handle, ...
                iter.x0(1, 1:3), iter.trim_indices(1), ...
                squeeze(iter.reference_trajectory_points(1, :, :)), iter.v_ref(1, :), ...
                iter.obstacles, dynamic_obstacle_area, left, right, ...
                obj.checker, options.dt_seconds);

            info = ControlResultsInfo(iter.amount, options.Hp);
            info.is_exhausted = is_exhausted; % read by the helper below (:84)
            info.n_expanded = n_expanded;

            if is_exhausted
                % GraphSearch.m:57-61: only n_expanded, is_exhausted and the tree are set
                info.tree = OptimizerInterface.create_tree(iter);
                return
            end

            % NodeInfo order [x y yaw trim g h k exactEval] (NodeInfo.m:4-14)
            next_nodes = cell(1, options.Hp);

            for k = 1:options.Hp
                next_nodes{k} = [y_pred(1, k), y_pred(2, k), y_pred(3, k), trims(k + 1), ...
                                     g_path(k + 1), h_path(k + 1), k, 1];
            end

            y_full = {[y_pred', trims(2:end)']}; % one row per step: the helper takes the last row of each step
            info = OptimizerInterface.create_control_results_info_from_mex( ...
                info, iter, options, next_nodes, trims, y_full);
            info.shapes = shapes; % 1 x Hp cell of 2 x n (return_path_area.m:5-7)
            info.needs_fallback = false; % GraphSearch.m:88
        end

        function info = run_joint(obj, iter, options)
            % iter.amount = nV > 1: expand_node.m:15-75, are_constraints_satisfied_sat.m:15-53 on the GPU
            nV = iter.amount; Hp = options.Hp;
            assert(obj.checker == 0, 'joint searches use the SAT checker (are_constraints_satisfied_interx.m:12)');
            obstacles = cell(nV, max(1, numel(iter.obstacles)));
            obstacles(1, 1:numel(iter.obstacles)) = iter.obstacles(:)'; % obstacles of the search: row 1
            R = size(iter.dynamic_obstacle_area, 1);
            dyn = cell(nV, max(1, R) * Hp);

            for k = 1:Hp
                dyn(1, (1:R) + max(1, R) * (k - 1)) = iter.dynamic_obstacle_area(:, k)';
            end

            [is_exhausted, n_expanded, trims, y, shapes, g_path, h_path] = pdmpc_b200_mex( ...
                GraphSearchCuda.PLAN_JOINT, obj.handle, iter.x0(:, 1:3), iter.trim_indices(:), ...
                iter.reference_trajectory_points, iter.v_ref, obstacles, dyn, ...
                iter.predicted_lanelet_boundary(:, 1), iter.predicted_lanelet_boundary(:, 2), 0, options.dt_seconds);
            y = reshape(y, 3, Hp, nV);
            info = ControlResultsInfo(nV, Hp);
            info.is_exhausted = logical(is_exhausted(1));
            info.n_expanded = n_expanded(1);

            if info.is_exhausted
                info.tree = OptimizerInterface.create_tree(iter);
                return
            end

            next_nodes = cell(1, Hp);
            y_full = cell(nV, 1);

            for k = 1:Hp % nV x 8 rows in NodeInfo order; g and h are the joint values (Tree.m: one g, h per node)
                next_nodes{k} = [squeeze(y(1, k, :)), squeeze(y(2, k, :)), squeeze(y(3, k, :)), trims(:, k + 1), ...
                                     repmat([g_path(k + 1), h_path(k + 1), k, 1], nV, 1)];
            end

            for v = 1:nV
                y_full{v} = [y(:, :, v)', trims(v, 2:end)'];
            end

            info = OptimizerInterface.create_control_results_info_from_mex(info, iter, options, next_nodes, trims, y_full);
            info.shapes = shapes; % nV x Hp cell
            info.needs_fallback = false;
        end

        function s = stats(obj)
            % device-side counters of the last call: kernel_ms, pops, nodes, obstacle columns
            s = pdmpc_b200_mex(GraphSearchCuda.STATS, obj.handle);
        end

    end

    methods (Access = private)

        function upload_mpa_once(obj, mpa, options)
            % the analogue of the reference's library/*.mat cache key
            % (MotionPrimitiveAutomaton.m:59-79): tables are staged in HBM once per MPA
            key = sprintf('%s_Hp%d_T%g_nc%d', string(options.mpa_type), options.Hp, ...
                options.dt_seconds, options.are_any_obstacles_non_convex);

            if strcmp(key, obj.mpa_key)
                return
            end

            pdmpc_b200_mex(GraphSearchCuda.UPLOAD_MPA, obj.handle, ...
                double(mpa.transition_matrix_single), mpa.maneuvers);
            obj.mpa_key = key;
        end

    end

    methods (Static)

        function dyn = with_hdv_reachable_sets(iter, checker, Hp)
            % Manual (human-driven) vehicles.  InterX checker: the reference tests the shape against the
            % reachable sets of the adjacent HDVs of step k right after the vehicle obstacles of step k
            % (are_constraints_satisfied_interx.m:23-31, vectorize_all_obstacles.m:47-62); the polygons are
            % NaN-separated and the answer is "any hit", so appending them as further rows of
            % dynamic_obstacle_area gives the same boolean.  SAT checker: the reference's HDV block never
            % tests anything (are_constraints_satisfied_sat.m:55-66 loops over find(adj) only if ~any(adj)),
            % so nothing is added.
            dyn = iter.dynamic_obstacle_area;
            adjacent_hdv = find(iter.hdv_adjacency);

            if checker ~= 1 || isempty(adjacent_hdv) || isempty(iter.hdv_reachable_sets)
                return
            end

            n_steps = min(size(iter.hdv_reachable_sets, 2), Hp);
            rows = cell(numel(adjacent_hdv), Hp);
            rows(:, 1:n_steps) = iter.hdv_reachable_sets(adjacent_hdv, 1:n_steps);
            rows(cellfun(@isempty, rows)) = {zeros(2, 0)};

            if isempty(dyn)
                dyn = rows;
            else
                assert(size(dyn, 2) == Hp, 'pdmpc_b200:hdv', 'dynamic_obstacle_area must have Hp columns');
                dyn = [dyn; rows];
            end

        end

    end

end
