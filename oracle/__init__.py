"""TEST INFRASTRUCTURE ONLY.

CPU restatement ("oracle") of the GridUniverse hot path -- the transition
function, batched step / rollouts, the Bellman sweep, greedy extraction, the
VI / PI drivers and Monte-Carlo evaluation -- used as the *checker* by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.  Nothing in ``griduniverse_b200/`` imports this
package; the product path fails loudly when its CUDA library is missing.

Parity status: PINNED.  The oracle is checked against (a) the ten known-answer
cases of the reference's own unit tests (tests/test_griduniverse.py:7-176) and
(b) golden vectors generated in the authoring container by importing the
unmodified reference under ``oracle/ref_shim.py``
(``tests/golden/make_golden.py`` is the generating script; the vectors are
committed under ``tests/golden/``).
"""
