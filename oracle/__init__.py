"""CPU oracle for the LM / Poisson-MLE hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package; nothing under ``dphtools_b200/`` does.
"""
