#!/usr/bin/env python
"""suspend()/resume() demo (reference: example/pytorch/elastic_benchmark_byteps.py:124-133):
the engine is torn down and rebuilt; tensor keys stay stable because declared
names are re-declared in their original order."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import byteps_b200.torch as bps  # noqa: E402

bps.init()
t = torch.ones(1000) * (bps.rank() + 1)
print("before:", bps.push_pull(t, name="elastic.t")[0].item())
bps.suspend()
bps.resume(int(os.environ.get("DMLC_NUM_WORKER", 1)), int(os.environ.get("DMLC_NUM_SERVER", 0)))
print("after :", bps.push_pull(t, name="elastic.t")[0].item())
bps.shutdown()
