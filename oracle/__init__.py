"""CPU oracle for the NOVA point-cloud diffusion-head sampling path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline.  The product (``nova_pointcloud_b200``) never imports it and
fails loudly when its CUDA library is missing.

The oracle is a from-scratch restatement (torch CPU / numpy / scipy) of the
reference's arithmetic for the hot path named in BASELINE.json:

* ``oracle.head``      -- ``DiffusionMLP.forward``         (diffnext/models/diffusion_mlp.py:26-99,
                          normalization.py:24-36, embeddings.py:139-166)
* ``oracle.scheduler`` -- ``FlowMatchEulerDiscreteScheduler`` (diffnext/schedulers/scheduling_cfm.py:39-49,92-104,125-140)
* ``oracle.loop``      -- ``Transformer3DModel.denoise`` + ``GuidanceScaler``
                          (diffnext/models/transformers/transformer_3d.py:102-113, guidance_scaler.py:46-87)
* ``oracle.chamfer``   -- Chamfer variants A/B/C (demo.py:38-55, train_newloss.py:316-349,
                          test_optimize.py:354-383)
* ``oracle.geometry``  -- kNN / local density / softmax interpolation (transformer_pointcloud_nova.py:81-89,128-152)
* ``oracle.training``  -- training tables, add_noise, get_losses forward (scheduling_cfm.py:39-49,87-117,
                          transformer_3d.py:81-95)
* ``oracle.partition`` -- set schedules (pipeline_nova.py:129-132, transformer_pointcloud_nova.py:63-78,
                          embeddings.py:262-270)

Parity pinning: the reference ships NO tests, golden vectors or fixtures for
this path, so the pins are outputs of the reference's own
modules, imported unmodified from /root/reference in the build container by
``tests/make_golden.py`` and committed under ``tests/golden/``.  The oracle is
checked against those fixtures in ``tests/test_oracle_golden.py`` (CPU), and,
when /root/reference is mounted, directly against the live reference modules in
``tests/test_oracle_vs_reference.py``.
"""
