classdef MonteCarloTreeSearchCuda < OptimizerInterface
    % MONTECARLOTREESEARCHCUDA  B200 drop-in for MonteCarloTreeSearch
    %   (hlc/optimizer/graph_search/MonteCarloTreeSearch.m), the sampled optimizer selected by
    %   OptimizerType.MatlabSampled (OptimizerInterface.m:28-30).
    %
    %   Same call surface: info_v = run_optimizer(obj, vehicle_index, iter, mpa, options, time_step).
    %   The random stream is the reference's: RandStream('mt19937ar', Seed = time_step + vehicle_index)
    %   (MonteCarloTreeSearch.m:31); the roll-outs, the n_expansions_max budget (:8, and
    %   config/mcts.json, :17-27) and the returned ControlResultsInfo fields (:199-249) are
    %   reproduced on the GPU by pdmpc_mcts_plan_batch (include/pdmpc_b200.h).
    %   Selected by OptimizerType.CudaSampled (INTEGRATION.md).  No CPU fallback.

    properties
        n_expansions_max (1, 1) double = 250;
    end

    properties (SetAccess = private, Hidden = true)
        handle = uint64(0);
        mpa_key = '';
        checker (1, 1) double = 1;
    end

    properties (Constant, Hidden = true)
        CREATE = 0;
        DESTROY = 1;
        UPLOAD_MPA = 2;
        PLAN_SAMPLED = 5;
    end

    methods

        function obj = MonteCarloTreeSearchCuda(options, device_id)

            arguments
                options (1, 1) Config
                device_id (1, 1) double = 0
            end

            obj = obj@OptimizerInterface();
            config_file = fullfile('config/mcts.json'); % MonteCarloTreeSearch.m:15-27

            if isfile(config_file)
                mcts_config = jsondecode(fileread(config_file));

                if isfield(mcts_config, 'n_expansions_max')
                    obj.n_expansions_max = mcts_config.n_expansions_max;
                end

            end

            obj.handle = pdmpc_b200_mex(MonteCarloTreeSearchCuda.CREATE, device_id);
            obj.checker = double(options.are_any_obstacles_non_convex);
        end

        function delete(obj)

            if obj.handle ~= 0
                pdmpc_b200_mex(MonteCarloTreeSearchCuda.DESTROY, obj.handle);
                obj.handle = uint64(0);
            end

        end

        function info_v = run_optimizer(obj, vehicle_index, iter, mpa, options, time_step)
            assert(iter.amount == 1); % MonteCarloTreeSearch.m:49
            obj.upload_mpa_once(mpa, options);
            left = zeros(2, 0); right = zeros(2, 0);

            if ~isempty(iter.predicted_lanelet_boundary)
                left = iter.predicted_lanelet_boundary{1, 1};
                right = iter.predicted_lanelet_boundary{1, 2};
            end

            Hp = size(iter.v_ref, 2);
            [is_exhausted, n_expanded, trims, y_pred, g_path, ~, shapes] = pdmpc_b200_mex( ...
                MonteCarloTreeSearchCuda.PLAN_SAMPLED, obj.handle, ...
                iter.x0(1, 1:3), iter.trim_indices(1), ...
                squeeze(iter.reference_trajectory_points(1, :, :)), iter.v_ref(1, :), ...
                iter.obstacles, iter.dynamic_obstacle_area, left, right, ...
                obj.checker, options.dt_seconds, time_step + vehicle_index, obj.n_expansions_max);

            info_v = ControlResultsInfo(iter.amount, Hp);
            info_v.n_expanded = n_expanded; % :199
            info_v.is_exhausted = is_exhausted;

            if is_exhausted % :202-205
                return
            end

            % path-only tree, as the reference builds it (:207-239): k = 1..Hp+1 along the path,
            % g = -1 except the goal leaf, h = -1
            tree = Tree();
            tree.trim = trims;
            tree.parent = uint32(0:Hp);
            tree.x = [iter.x0(1, 1), y_pred(1, :)];
            tree.y = [iter.x0(1, 2), y_pred(2, :)];
            tree.yaw = [iter.x0(1, 3), y_pred(3, :)];
            tree.k = 1:(Hp + 1);
            tree.g = [-1 * ones(1, Hp), g_path(end)];
            tree.h = -1 * ones(1, Hp + 1);
            info_v.tree = tree;
            info_v.tree_path = 1:(Hp + 1);
            info_v.y_predicted = y_pred; % 3 x Hp (return_path_to.m:11-25)
            info_v.shapes = shapes;
            info_v.predicted_trims = trims(2:end);
            info_v.needs_fallback = false;
        end

    end

    methods (Access = private)

        function upload_mpa_once(obj, mpa, options)
            key = sprintf('%s_Hp%d_T%g_nc%d', string(options.mpa_type), options.Hp, ...
                options.dt_seconds, options.are_any_obstacles_non_convex);

            if strcmp(key, obj.mpa_key)
                return
            end

            pdmpc_b200_mex(MonteCarloTreeSearchCuda.UPLOAD_MPA, obj.handle, ...
                double(mpa.transition_matrix_single), mpa.maneuvers);
            obj.mpa_key = key;
        end

    end

end
