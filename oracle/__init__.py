"""CPU oracle for the VMLMF compressed-LSTM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``vmlmf_b200/`` may import this package.
Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` (as the checker
and as the timed CPU baseline -- never as the product path).

Two restatements live here:

* :mod:`oracle.vmlmf_oracle` -- a functional, eager-torch-on-CPU restatement of
  the reference cells, layers and networks, op order kept the same as the
  reference so fp32 results agree to rounding.  Gradients come from torch
  autograd, exactly as in the reference.
* :mod:`oracle.canonical_numpy` -- numpy (fp64 or fp32) restatement of the
  *canonical* recurrence every cell reduces to, with a hand-derived
  backward-through-time.  This is the spec the CUDA kernels implement.

Parity pin: the reference ships no golden vectors (its unit tests assert
shapes only, unittest/unit_test.py:63-93).  The oracle is therefore pinned
against outputs of the *live reference modules* executed in the build
container; those outputs are committed as fixtures under ``tests/golden/``
together with the generating script ``tests/golden/make_golden.py``.
"""
