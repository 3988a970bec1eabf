"""CPU oracle for the ViPFormer hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  Nothing under
``vipformer_b200/`` does (tests/test_boundary.py greps for it).

* ``oracle.tokenizer``  -- ctypes binding of ``tokenizer_oracle.c`` (plain C,
  pinned arithmetic; restates vipformer/model/pointcloud/utils.py:6-141).
* ``oracle.model_ref``  -- plain PyTorch fp32 restatement of the floating-point
  blocks (Group2Emb, input adapter, attention, MLP, encoder, heads) and of the
  un-vendored lightly==1.1.21 NT-Xent loss.

Parity pinning: ``tests/make_golden.py`` imports the real reference from
``/root/reference`` (in the build container only) and writes input/output
vectors to ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
oracle against them.  NT-Xent has no reference source in-tree (third-party
dependency) -> that one function is "parity unpinned" (see DESIGN.md).
"""
