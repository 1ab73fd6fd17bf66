"""CPU oracle for the EP-in-Kalman hot path of AaltoML/nonstationary-audio-gp.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
(``nonstationary-audio-gp_b200/``) never imports, links or executes anything in
here; it fails loudly when its CUDA library is missing.

What it is: a line-by-line FP64 NumPy/SciPy restatement of the reference's
MATLAB files for the hot path (each function cites the ``matlab/...m:line`` it
follows), in the reference's own *dense n-by-n* operation order.

PARITY UNPINNED.  The reference ships no golden vectors, known-answer tests or
expected lZ/RMSE values for this path, and neither
MATLAB nor GNU Octave exists in the build container or on the GPU box, so the
reference itself cannot be executed.  Third-party arithmetic the reference
calls and that is not under /root/reference (MATLAB built-ins and toolboxes,
versions unpinned: expm, balance, chol, eig, roots, unique, dare [Control
System Toolbox], normpdf/cholcov [Statistics Toolbox]) is restated through the
published algorithms' SciPy/LAPACK equivalents (scipy.linalg.expm,
matrix_balance (xGEBAL, same routine MATLAB calls), cholesky, numpy.roots,
solve_discrete_are).  The oracle is instead pinned by
  * known-answer identities (cubature moment identities, the gauher(20) table,
    Kalman/RTS versus a direct dense GP solve, info-form equivalence of the two
    update branches, exact block-diagonality of P) -- tests/test_oracle_*.py,
  * a second, independent restatement in C (oracle/c/nsagp_oracle.c) compared
    with this one on the committed golden fixtures (tests/golden/).
"""
