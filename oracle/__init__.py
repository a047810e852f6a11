"""CPU oracle for the MPPI / Neural Laplace planning hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the CPU arm that is timed beside the GPU number.  The product
package (``neurallaplacecontrol_b200``) never imports this package and has no
CPU fallback.

What it restates (all ``file:line`` are relative to the reference tree):

* ``oracle.mppi``   - ``planners/mppi_delay.py:193-356`` (``MPPIDelay.command``
  and helpers), op for op, with the noise tensor injected.
* ``oracle.nl_model`` - ``w_nl.py:14-145`` (GRU history encoder, Laplace
  representation MLP, model forward) from a reference ``state_dict``.
* ``oracle.costs``  - the env reward formulas
  (``envs/oderl/envs/ctpendulum.py:139-155``, ``ctcartpole.py:289-346``,
  ``ctacrobot.py:153-166,233-255``, ``base_env.py:297-301``) negated as in
  ``mppi_with_model.py:145-171``.
* ``oracle.ilt``    - the Fourier-series inverse Laplace transform and the
  Riemann-sphere maps of the third-party ``torchlaplace`` package that
  ``w_nl.py:6,137-144`` calls.

PARITY STATUS
-------------
* Stages 1, 3, 4 and the rollout bookkeeping, the GRU encoder and the
  representation MLP are **pinned**: ``tests/golden/*.npz`` were produced by
  importing the reference's own ``MPPIDelay`` / ``ReverseGRUEncoder`` /
  ``LaplaceRepresentationFunc`` in the build container
  (``oracle/gen_golden.py``), and ``tests/test_oracle_*`` check this package
  against them.
* The inverse Laplace transform itself is **parity unpinned**:
  ``torchlaplace`` is an un-vendored, un-pinned dependency
  (``requirements.txt:17``) that is absent from the reference tree and from
  this image, and the reference ships no test or fixture at that boundary.
  ``oracle.ilt`` restates its published algorithm (Fourier series ILT with
  ``alpha=1e-3, tol=10*alpha, scale=2, eps=1e-6``; stereographic Riemann
  sphere maps) and is anchored on closed-form transform pairs only.
"""
