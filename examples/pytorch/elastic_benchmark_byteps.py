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
before = bps.push_pull(t, name="elastic.t")[0].item()
if bps.rank() == 0:
    print("before: %.1f on %d workers" % (before, bps.size()))
bps.suspend()
# same topology here; a real elastic job passes the NEW number of worker boxes / servers (and restarts the
# scheduler and servers with it) - see tests/test_ps_api.py::test_elastic_suspend_resume_against_new_cluster
bps.resume(int(os.environ.get("DMLC_NUM_WORKER", 1)), int(os.environ.get("DMLC_NUM_SERVER", 0)))
after = bps.push_pull(t, name="elastic.t")[0].item()
if bps.rank() == 0:
    print("after : %.1f on %d workers" % (after, bps.size()))
bps.shutdown()
