"""CPU oracle for the CAGroup3D inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cagroup3d_b200/``, ``pcdet/`` or
``tools/`` may import this package; only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and
there only as the checker or the timed CPU baseline.

Parity status: the sparse-convolution arithmetic of the
reference lives in MinkowskiEngine v0.5.4, which is neither vendored under
/root/reference nor installable offline, and the reference ships no tests or
golden vectors -> the MinkowskiEngine part of this oracle is "parity unpinned"
(it restates SURVEY.md Appendix A).  The box / IoU / NMS / coder arithmetic IS
pinned: against the reference's own CPU IoU compiled from its sources
(``oracle/_ref``, see ``oracle/Makefile``) and against golden vectors generated
by importing the reference's pure-torch helpers (``tests/golden/make_golden.py``).
"""
