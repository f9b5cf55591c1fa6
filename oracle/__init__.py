"""CPU oracle for the obman_train hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker / the CPU arm that is
timed beside the CUDA path.  The product package (``obman_train_b200``) never
imports it and fails loudly when its CUDA library is missing.

Contents
--------
``icosphere``   restatement of ``trimesh.creation.icosphere`` (call site
                /root/reference/mano_train/networks/branches/atlasbranch.py:64-70).
``mano``        restatement of ``manopth.manolayer.ManoLayer.forward`` (external,
                un-vendored, un-pinned dependency; call sites
                /root/reference/mano_train/networks/branches/manobranch.py:92-105,170-182).
                PARITY UNPINNED: ``manopth`` is absent from /root/reference and from
                this image, so this restatement of the published algorithm is only
                checked for self-consistency (zero pose/shape => template, rotation
                equivariance, fp64 finite differences), not against upstream output.
``geometry``    restatement of ChamferLoss / batch_pairwise_dist /
                batch_mesh_contains_points / compute_contact_loss / masked_mean_loss /
                edge_loss / AtlasLoss / ManoLoss.  PINNED against the reference's own
                Python files executed through ``oracle.refhook`` (tests/test_oracle_vs_reference.py,
                runs only where /root/reference is mounted) and against the committed
                golden vectors under tests/golden/ generated the same way.
``nets``        functional restatement (torch CPU, fp32/fp64) of ResNet-18, ManoBranch,
                AtlasBranch/PointGenCon and HandNet.forward keyed by the reference's
                state-dict names.  PINNED the same way.
``refhook``     import hook that executes the reference's unmodified files from
                /root/reference with ``.cuda()`` stripped (container only; never on the GPU box).
"""
