classdef OrthogonalMatchingPursuit < handle
    % Stand-in for sparse-plex's spx.pursuit.joint.OrthogonalMatchingPursuit (external, unpinned: README.md:9)
    % so the reference's drivers run unmodified (plot_errorVSsnr.m:116-118):
    %     s = spx.pursuit.joint.OrthogonalMatchingPursuit(A, K);  r = s.solve(Y);  r.Z
    % served by the jstsp_somp MEX gateway (row-l2 simultaneous OMP on the GPU).
    properties
        Dict
        K
    end
    methods
        function self = OrthogonalMatchingPursuit(Dict, K)
            self.Dict = double(Dict);
            self.K = K;
        end
        function result = solve(self, Y)
            [Z, support, R] = jstsp_somp(self.Dict, Y, self.K);
            result.Z = Z;
            result.R = R;
            result.support = support;
            result.iterations = numel(support);
        end
    end
end
