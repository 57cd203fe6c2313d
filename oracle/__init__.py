"""CPU oracle for the vit4hep CFM-ViT hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``vit4hep_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` use it, as the checker and the timed CPU baseline respectively.

Parity status: the reference (luigifvr/vit4hep) ships no tests, golden vectors or
fixtures, so the restatement in ``vit_oracle.py`` is pinned against
outputs of the *unmodified reference files run in the build container*
(``oracle/make_golden.py`` imports ``/root/reference`` through the three third-party
stubs in ``ref_stubs.py`` and writes ``tests/golden/*.npz``).  Third-party arithmetic on
the path that is absent from ``/root/reference`` and therefore restated from its
published algorithm (versions unpinned in requirements.txt): timm ``Mlp``, xformers
``memory_efficient_attention``, torchdiffeq fixed-grid ``rk4`` (3/8 rule).
"""
