"""CPU oracle: a NumPy restatement of geomjax's static Riemannian transition kernels.

TEST INFRASTRUCTURE ONLY.  Nothing under ``geomjax_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / reported CPU baseline.

Why a restatement and not the reference itself: the reference is pure Python on JAX
(``/root/reference/requirements.txt:7-8``) and neither ``jax`` nor ``jaxlib`` exists in
this image or its wheelhouse, so ``import geomjax`` is impossible here and on the GPU box.
Every function cites the reference ``file:line`` it follows.

Parity pins (see ``tests/test_oracle_pins.py``):
  * threefry2x32 Random123 known-answer vectors and the JAX-documented outputs of
    ``split`` / ``uniform`` / ``normal`` in both threefry modes;
  * the reference's only author-produced artefact, the ``nutsrmhmc`` golden vector
    ``[-0.73879963, 1.2370402]`` (``tests/test_samplers.py:10-19``), reproduced through
    ``oracle.nuts_rmhmc`` which shares the PRNG, funnel target, momentum draw,
    implicit-midpoint integrator and energy with the static ``rmhmc`` kernel;
  * the reference test's cross-sampler equivalences ``rmhmc ~ hmc ~ lmc`` at rtol 1e-4
    (``tests/test_samplers.py:21-57``) and its documented Monge mismatch (``:58-59``).
For the static kernels' log-densities, energies, accept decisions and Info fields the
reference's own tests hold no vectors: those are pinned only transitively (same code
path as the golden vector) -- "parity pinned through NUTS golden vector; static-kernel
Info fields unpinned by the reference".

Batched convention: every array carries a leading chain axis ``(C, ...)``; this is the
``jax.vmap(kernel)(keys, states)`` of ``examples/funnel/main.py:19``.
"""
