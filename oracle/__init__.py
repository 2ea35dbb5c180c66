"""CPU oracle for the FFR-Net hot path — TEST INFRASTRUCTURE ONLY.

A plain-PyTorch fp32 (and numpy float64 for the scoring) restatement of the reference algorithm, written from the
reference's behaviour with every function citing the reference file:line it follows. It is pinned against the
real reference, imported from /root/reference in the build container, by tools/make_golden.py, which commits
small fixtures under tests/golden/ (the reference ships no tests or golden vectors of its own).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package,
and only as the checker or the timed CPU baseline — never as part of the product path.
"""
