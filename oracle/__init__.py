"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the AAURoverEnv-v0 non-physics MDP hot path.

Importable from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs only.  The product package ``isaac_rover_orbit_b200`` never imports it.
"""
