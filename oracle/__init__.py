"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's graph message-passing hot path (ecmwf/anemoi-models,
`layers/conv.py`, `layers/block.py`, `distributed/{shapes,khop_edges}.py`) and of the few
torch-geometric (>=2.3,<2.5, un-vendored dependency) entry points it calls.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package, and there only as the checker / the reported CPU baseline.
The product package `anemoi_models_b200` never imports it and has no CPU fallback.

Parity status: the reference's own tests hold NO golden vectors for this path, so the
oracle is pinned against outputs of the unmodified reference Python run in the build container over
the restated PyG shim (`oracle/make_golden.py` -> `tests/golden/*.npz`), and cross-checked against an
independent float64 dense masked-softmax formulation (`oracle/gtconv.py: gt_conv_dense_f64`).
"""
