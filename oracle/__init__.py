"""CPU oracle for the jstsp19 channel-estimation hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  Nothing under ``jstsp19_b200/`` imports
``oracle`` and the product path raises when its CUDA library is missing.

What it is: a line-by-line fp64 NumPy restatement of the reference's MATLAB
files (citations ``file:line`` are relative to the reference root):

* ``matlab_compat``  – MATLAB built-in semantics the path relies on
  (Hermitian ``toeplitz``, ``round`` half away from zero, ``norm`` = sigma_max,
  ``eigs`` = 6 largest, stable descending ``sort``, column-major ``vec``).
* ``system_model``   – ``wideband_mmwave_channel.m``,
  ``wideband_hybBF_comm_system_training.m``, ``proposed_hbf.m``, ``hbf.m``,
  ``createBeamformer.m``, ``qam4mod.m``.
* ``estimators``     – ``proposed_algorithm.m``, ``proposed_algorithm_angles.m``,
  ``svt.m``, ``mc_svt.m``, ``mc_admm.m``, ``sparse_admm.m``, ``OMP.m``; each in a
  *literal* mode (materialises the reference's dense Kronecker operators) and a
  *structured* mode (Kronecker-free, proven equal to the literal one in
  ``tests/test_oracle.py``).
* ``vamp``           – ``vamp.m`` + ``VampGlmEst.m`` + the three GAMPmatlab
  estimator classes it instantiates.
* ``fixtures``       – seeded per-trial inputs drawn in the reference's RNG
  consumption order.

PARITY UNPINNED: the reference ships no golden vectors, no seeds and no tests
(SURVEY.md section 4), and neither MATLAB nor Octave exists in this image, so
the oracle cannot be checked against the reference interpreter.  What pins it
instead: (1) literal == structured to <=1e-12; (2) algebraic properties
(mask cardinality, SVT/soft-threshold fixed points, OMP support monotonicity);
(3) NMSE-vs-SNR order of magnitude against ``results/errorVSsnr.fig``
(BASELINE.md section 1).
"""

from . import matlab_compat, system_model, estimators, vamp, fixtures  # noqa: F401
