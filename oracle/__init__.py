"""Test infrastructure only: CPU restatement of the reference hot path.

Nothing under ``oracle/`` is imported by the product package
(``zeroshotsemanticsegmentation_b200``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and only as the checker
or as the timed CPU baseline.
"""
