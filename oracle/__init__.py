"""TEST INFRASTRUCTURE ONLY: CPU oracles for the SA/LCP construction path.

Nothing under ``psac_b200/`` may import this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.
"""
