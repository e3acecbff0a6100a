"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (plain PyTorch fp32 + numpy float64) of the reference's translate path
(achilles1989/NanoDecoder): signal front end, encoders, decoders, greedy and beam decode loops.
Every function cites the reference file:line it follows.

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs — as the checker or as the timed CPU baseline, never
as part of the product path.  ``nanodecoder_b200/`` must not import it.

Parity pinning: the reference has no golden vectors for this path, and its
arithmetic lives in PyTorch (pinned torch==1.0.0 upstream; torch 2.11 here).  The oracle is
therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container through
``oracle/refshim.py`` by ``oracle/make_golden.py``; the resulting vectors are committed under
``tests/golden/`` and checked by ``tests/test_oracle_golden.py`` on every CPU test run.
"""
