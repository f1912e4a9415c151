"""CPU oracle for the HYMLS preconditioner hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy / scipy) of the reference algorithm
(nlesc-smcm/hymls, `src/HYMLS_*.cpp`).  It exists to check the CUDA product
path and to serve as the timed CPU baseline in `bench.py`.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product (`hymls_b200/`) never does.

Parity status:
  * index maps (partitioner / groups): PINNED by the closed-form expectations of
    the reference's own unit tests (tests/test_oracle_partitioner.py).
  * matrix generators: PINNED to 1e-14 against the shipped MatrixMarket
    fixtures (committed, reduced, under tests/golden/).
  * exact path (Number of Levels = 0): PINNED by the reference's "1 iteration"
    integration targets on the Stokes fixtures.
  * subdomain solver of the timed CPU baseline (oracle/cpp): F-matrix ordering, scaling and static pivots of the
    reference; its fill is PINNED to the known-answer counts of unit_tests/HYMLS_SparseDirectSolver.cpp
    (78 / 397 / 2033 there, 78 / 396 / 2128 here: AMD vs exact minimum degree; tests/test_oracle_cpp.py).
  * approximate path (levels >= 1): the reference ships no golden ApplyInverse
    vectors; pinned only through iteration-count / residual targets of the
    reference's integration tests ("parity unpinned" beyond that) -- on the shipped
    fixtures (tests/test_oracle_solver.py, test_oracle_skew.py) and on the reference's
    generated multi-level problems threeD1.xml (34 iterations, bound 35) and stokes6.xml
    (28-29, bound 30) (tests/test_oracle_cpp.py).
"""
