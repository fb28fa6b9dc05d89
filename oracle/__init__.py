"""CPU oracle for the BV MaxEnt optimise step.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product path
(``jaxent_b200``) never imports it and has no CPU fallback.
"""
