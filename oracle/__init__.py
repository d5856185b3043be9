"""oracle/ — TEST INFRASTRUCTURE ONLY (parity checker + CPU baseline), never the product path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package.  `dcs-net_b200/` must never import it: the product path fails
loudly when the CUDA extension is missing.

PARITY UNPINNED: the reference (jackhwalters/DCS-Net) ships no tests, golden vectors or
known-answer fixtures for this path, and the arithmetic of its layers
lives in the un-vendored third-party package complexPyTorch==0.3.  The pins this oracle has:
  * `oracle/dcsnet_oracle.py` (standalone fp32 torch-CPU restatement) is checked in the build
    container against the reference's OWN files executed unmodified behind import stubs
    (`oracle/reference_harness.py`, `tests/test_oracle_vs_reference.py`);
  * golden vectors generated from that reference execution are committed under
    `tests/golden/` by `oracle/make_golden.py`.
"""
