"""CPU oracle for the batched explicit-RK ODE hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / CPU
baseline -- never as the thing measured or shipped.  The product package
(``torchdyn_b200``) must not import this package.

Contents
--------
``erk_oracle.py``   torch-CPU restatement of the reference algorithm
                    (tableaus, steppers, controller, adaptive / fixed loops,
                    dense output, backsolve + interpolated adjoint incl.
                    integral losses, natural cubic spline, hybrid / ALF
                    loops are pinned through fixtures of the executed
                    reference instead).  Every function cites the reference
                    file:line it follows.
``ref_shim.py``     imports the *real* reference from ``/root/reference`` with
                    the three absent third-party packages stubbed.  Only
                    usable in the build container (the GPU box has no
                    ``/root/reference``); used by ``make_golden.py`` and by the
                    CPU tests that pin the oracle against the executed
                    reference when it is present.
``make_golden.py``  regenerates ``tests/golden/*.pt`` from the executed
                    reference.

Parity status: PINNED -- the restatement is checked (bit-exact on CPU) against
outputs of the reference itself run in the build container, committed as
fixtures under ``tests/golden/`` (see ``tests/test_oracle_golden.py``).  The
one exception is the natural cubic spline used by the interpolated adjoint:
the reference delegates it to the third-party ``torchcde`` (``^0.2.3``, not
vendored, not installed, no lockfile) -- spline-level parity is UNPINNED and
anchored only on the reference's own gradient-level checks.
"""
