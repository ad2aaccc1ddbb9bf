"""CPU oracle for the Monte-Carlo certification hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``certifyingfacerecognition_b200`` may
import from here: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / reported baseline -- never as the product path.

Parity status: PINNED.  ``oracle/mc_path.py`` restates the reference's
algorithm (plain torch fp32 on CPU, every function cites the reference
file:line it follows).  ``oracle/make_golden.py`` imports the *unmodified*
reference modules from ``/root/reference`` (through the shim set of
``oracle/reference_shims.py``), runs them on the seeded fixture weights of
``oracle/fixtures.py`` and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks the restatement against those vectors.
"""
