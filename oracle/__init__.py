"""CPU oracle for the COCO-DR hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker or
as the timed CPU baseline -- never as part of the product path.  The product
(`cocodr_b200`) must fail loudly when its CUDA library is missing.

Parity status: the reference ships no tests / golden vectors (SURVEY.md §4,
§8c: "parity unpinned" by the reference itself).  The oracle is therefore
pinned against *outputs of the reference classes run in the dev container*
(``oracle/make_golden.py`` imports ``/root/reference`` and writes
``tests/golden/*.npz``); ``tests/test_oracle_golden.py`` checks every oracle
function against those fixtures.
"""
