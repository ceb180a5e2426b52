"""msrflute_b200 — a Blackwell (B200, sm_100a) native federated-learning simulation engine.

Capability parity target: microsoft/msrflute (FLUTE).  Layout:

* ``core``        config/schema, server round loop, client, trainer, strategies, evaluation
* ``models``      model zoo for every shipped task (built on ``ops``)
* ``ops``         hand-written sm_100a CUDA kernels + their PyTorch reference fallbacks
* ``parallel``    flat arenas, symmetric memory, fused broadcast/gather transports
* ``extensions``  privacy (local/global DP, RDP), quantization, RL aggregation weights
* ``utils``       optimizers, schedulers, samplers, logging, IO
* ``data``        synthetic federated datasets of the benchmark shapes
* ``dp_accountant`` PRV privacy accountant
"""
__version__ = "0.1.0"
