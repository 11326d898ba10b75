"""CPU oracle for the PTv3 / spconv hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pointcept_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / the CPU
baseline -- never as the thing shipped.

Pinning status ("Oracle"):
  * serialization (z-order / Hilbert / encode / argsort), patch padding and the
    dense attention math are PINNED: ``tools/gen_golden.py`` imports the
    reference's own python files from /root/reference and writes
    ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
    restatement against those fixtures bit-exactly (integers) / to 1e-6 (fp32).
  * sparse convolution (rulebook + gather-GEMM-scatter): PARITY UNPINNED.  The
    arithmetic lives in third-party ``spconv-cu124`` (unpinned version,
    reference ``environment.yml:47``) whose source is not vendored in the
    reference and is not installable offline.  ``oracle/spconv_ref.py`` restates
    its published semantics anchored on the reference
    call sites.
"""
