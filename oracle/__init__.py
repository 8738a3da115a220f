"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatements (plain torch fp32/fp64) of the reference's algorithms for the ViT encoder hot path,
used exclusively by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs as the checker.  Nothing under vit-of-pytorch_b200/ imports this package; the product path
raises when libvitb200.so or a B200 is missing instead of routing here.

Pinning: the reference ships no tests, golden vectors or fixtures for this path,
so the restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py imports
the unmodified /root/reference modules in the build container, runs them on seeded inputs and commits
the vectors under tests/golden/; tests/test_oracle.py checks the restatement against those vectors
everywhere and against the live reference wherever /root/reference exists.
"""
