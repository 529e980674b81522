function compile_pdmpc_b200
    % COMPILE_PDMPC_B200  Build libpdmpc_b200.so (nvcc, sm_100a) and the MEX shim.
    %   Register as a project StartUp task next to compile_priority_queue
    %   (hlc/optimizer/graph_search/priority_queue/compile_priority_queue.m:1-14,
    %   resources/project/.../RzujV5F5YFjdqHayFBC08MnkBSQd.xml), so that opening the
    %   project (startup.m:2) builds it exactly like the reference's own MEX.
    folder = fileparts(mfilename('fullpath'));
    is_mex_compiled = isequal(fileparts(which('pdmpc_b200_mex')), folder);

    if is_mex_compiled
        return
    end

    root = fileparts(fileparts(folder)); % repository root
    csrc = fullfile(root, 'p-dmpc_b200', 'csrc');
    inc = fullfile(root, 'include');
    lib = fullfile(csrc, 'libpdmpc_b200.so');

    if ~isfile(lib)
        fprintf('Compiling libpdmpc_b200.so (nvcc, sm_100a)\n');
        cmd = sprintf(['nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo ' ...
                           '--fmad=false -Xcompiler -fPIC -shared -o "%s" "%s"'], ...
            lib, fullfile(csrc, 'pdmpc_capi.cu'));
        status = system(cmd);
        assert(status == 0, 'nvcc failed; the GPU optimizer has no CPU fallback');
    end

    cwd = cd(folder);
    cleanup_obj = onCleanup(@() cd(cwd));
    fprintf('Compiling pdmpc_b200_mex\n');
    mex('-R2018a', ['-I' inc], 'pdmpc_b200_mex.cpp', ['-L' csrc], '-lpdmpc_b200', ...
        ['LDFLAGS=$LDFLAGS -Wl,-rpath,' csrc]);
end
