"""CPU oracle for the scarlet proximal-gradient path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; the product (``scarlet_b200``) never does.
"""
