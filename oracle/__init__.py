"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference (uber/fiber @ ad6faf02) ``Pool.map`` hot path and of the
workload bodies that BASELINE.json's configs map through it.  Nothing under ``fiber_b200/``
imports this package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may, and there only as the checker / the CPU arm.

Parity status: PINNED.  ``tests/golden/make_golden.py`` ran the *unmodified reference pool code*
(copied to a scratch dir outside the repo with the single transport constant
``fiber/socket.py:27`` flipped ``"nanomsg"`` -> ``"zmq"`` because ``nnpy`` is not installable
here) and the resulting vectors are committed under ``tests/golden/``; ``tests/test_oracle.py``
checks every function here against them, plus the published Philox4x32-10 / SplitMix64
known-answer vectors.
"""
