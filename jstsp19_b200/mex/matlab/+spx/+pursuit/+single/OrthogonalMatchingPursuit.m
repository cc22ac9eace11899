classdef OrthogonalMatchingPursuit < handle
    % Stand-in for sparse-plex's spx.pursuit.single.OrthogonalMatchingPursuit (external, unpinned: README.md:9):
    %     s = spx.pursuit.single.OrthogonalMatchingPursuit(Phi, K);  r = s.solve(y);  r.z
    % (plot_time_comparisions.m:83-85), served by the OMP MEX gateway with benchmark_algorithms/OMP.m's semantics.
    properties
        Dict
        K
    end
    methods
        function self = OrthogonalMatchingPursuit(Dict, K)
            self.Dict = double(Dict);
            self.K = K;
        end
        function result = solve(self, y)
            [x_hat, indexSet] = OMP(self.Dict, y(:), self.K, 0);
            result.z = x_hat;
            result.support = cell2mat(indexSet);
            result.iterations = numel(indexSet);
        end
    end
end
