"""CPU oracle for the GA3C predict/train hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import anything under this package.  The product
(`ga3c_b200/`) never imports it and has no CPU fallback.

PARITY STATUS
  * Network arithmetic (forward, loss, backward, RMSProp): **parity unpinned**.
    The reference delegates this arithmetic to TensorFlow 1.x, which is not in
    /root/reference, carries no version pin (README.md:8 says "TensorFlow 1.0")
    and is not installable here.  The reference has no tests or golden vectors.
    `oracle_np.py` restates the graph from the reference's call sites; it is
    cross-checked by finite differences and against torch CPU autograd
    (`oracle_torch.py`) as an independent second opinion on layout/padding.
  * Host-side functions (`_accumulate_rewards`, `convert_data`, `select_action`,
    predictor/trainer batching): **pinned** against the reference's own Python,
    imported from /root/reference in the build container by
    `oracle/gen_golden.py`; the outputs are committed under `tests/golden/`.
"""
