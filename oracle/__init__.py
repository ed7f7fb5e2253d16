"""CPU oracle for the sparse-conv hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import this package,
and only as the checker or as the timed CPU baseline.  The product (doda_b200/) never imports it and has no CPU
fallback.

Pinning status
  * spconv v1.2 (rulebook + native gather-GEMM-scatter algorithm): the library is NOT in /root/reference and cannot
    be installed here -> restated from its published algorithm and
    cross-checked against an independent dense formulation (torch conv3d on the densified tensor,
    tests/test_oracle_cpu.py).  The reference repo holds no golden vector for it: **parity unpinned**.
  * PG_OP / pointops2: restated in numpy from the reference sources (file:line cited per function) and pinned
    against outputs of the reference's own code compiled from /root/reference by oracle/build_ref.py
    (fixtures under tests/golden/, generating script tests/golden/make_golden.py).
"""
