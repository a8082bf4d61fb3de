"""CPU oracle for the U-Net hot path — TEST INFRASTRUCTURE ONLY.

A plain-PyTorch fp32 restatement (functional, driven by a ``state_dict``) of the reference's 2D/3D
U-Net forward, its Dice/IoU/BCE losses and its numpy metrics.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it;
the product package ``mdeical_image_segmentation_b200`` never does.

Parity pinning: the reference ships no golden vectors or asserting tests, so the
oracle is pinned against outputs of the reference itself: ``oracle/make_golden.py`` imports the
reference modules from ``/root/reference`` (with the import shims of ``oracle/ref_loader.py``), runs
them on seeded inputs and stores inputs + weights + outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` replays those fixtures through this restatement (fp32, exact to 1e-5).
"""
