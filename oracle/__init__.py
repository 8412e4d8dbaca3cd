"""CPU oracle of the shortest-tokenization path.  TEST INFRASTRUCTURE ONLY: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import,
call, link or execute anything in here; nothing under ``dp-tokenization_b200/`` does."""
