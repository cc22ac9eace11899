"""jstsp19_b200 - B200-native batched channel-estimation engine for the hot path of
vlaxose/jstsp19 (per-trial wideband hybrid-beamforming mmWave channel estimation).

The package is a thin host layer over ``libjstsp_b200.so`` (hand-written sm_100a CUDA
behind the C ABI of ``include/jstsp_b200.h``):

* ``jstsp19_b200.api``    - functions with the reference's MATLAB names/signatures
  (NumPy in, NumPy out; HOST-buffer C-ABI calls - what a MEX gateway does);
* ``jstsp19_b200.engine`` - device-resident batched engine on torch tensors
  (Monte-Carlo trials sharded over ranks, one NCCL reduction at the end).

There is no CPU fallback: without the shared library the import raises, without a
B200-class GPU every solver call raises.
"""
from . import _lib  # noqa: F401  (raises if the CUDA library is missing)
from . import api  # noqa: F401
from .api import (ls_estimate, capacity, capacity_sweep, power_model, energy_efficiency, OMP, OMP_kron, somp, createBeamformer, qam4mod, admm_parameters, hbf, log2det_rate, mc_admm, mc_svt, nmse, proposed_algorithm, proposed_algorithm_angles,  # noqa: F401
                  proposed_algorithm_pilots, proposed_algorithm_psi, proposed_hbf, sparse_admm, svt, vamp, wideband_hybBF_comm_system_training, wideband_mmwave_channel)

__all__ = ["ls_estimate", "capacity", "capacity_sweep", "power_model", "energy_efficiency", "proposed_algorithm", "proposed_algorithm_angles", "proposed_algorithm_psi", "proposed_algorithm_pilots", "svt", "mc_svt", "mc_admm", "OMP", "OMP_kron", "somp", "createBeamformer", "qam4mod", "sparse_admm", "vamp",
           "wideband_mmwave_channel", "proposed_hbf", "hbf", "wideband_hybBF_comm_system_training", "nmse", "admm_parameters", "log2det_rate"]
