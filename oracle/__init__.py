"""oracle/ — CPU checker for the B200 join engine.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / ``--impl reference`` legs may
import this package.  The product (flash_hash_join_b200/) never does and has no CPU fallback.

Contents
  join_oracle.c   plain-C restatement of /root/reference/hash_join.cpp (cites file:line per function)
  oracle.py       ctypes wrapper + an independent numpy restatement + loader for oracle/_ref
  build_ref.sh    compiles the unmodified reference (and a pairs-returning sed-patched temp copy)
                  into oracle/_ref/ (git-ignored; travels to the GPU box)
Parity status: PINNED against the compiled reference (tests/test_oracle.py) and tests/golden/.
"""
from .oracle import (  # noqa: F401
    ENTRY_POINTS,
    build,
    entry_point_name,
    hash64,
    join,
    load_reference,
    np_join,
    reference_available,
    sorted_pairs,
)
