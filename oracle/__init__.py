"""oracle/ -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the UniRes ADMM/CG hot path used as the parity checker for
the CUDA kernels in ``unires_b200``.

PARITY UNPINNED: the arithmetic of this path lives in the third-party package
``nitorch`` (pinned in /root/reference/setup.py:11 at commit
8067d60542642a39ab6c6eb5e1157373a9d3dcc3).  nitorch is neither vendored under
/root/reference nor installed here, and the reference ships no tests, golden
vectors or fixtures for the path.  The primitives in
``oracle/nitorch_shim`` therefore restate nitorch's *published* algorithms
 and are anchored on the reference's own call sites:
the reference's unmodified control-flow files (unires/_project.py,
unires/_update.py, unires/struct.py) are imported BY PATH on top of the shim
(``oracle/load_reference.py``) and used (a) to validate the independent
in-repo port ``oracle/unires_port.py`` and (b) to generate the golden
fixtures under tests/golden/ (``oracle/gen_golden.py``).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import anything from this package.
"""
