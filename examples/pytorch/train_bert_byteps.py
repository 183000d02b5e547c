#!/usr/bin/env python
"""BERT-large pre-training step loop (the reference's bandwidth-bound headline
workload) with the fused Adam exchange: bf16 parameters/gradients, fp32 master
weights and moments sharded over the GPUs, one kernel per bucket."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import byteps_b200.torch as bps  # noqa: E402
from byteps_b200.models import get_model  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--model", default="bert_large")
p.add_argument("--batch-size", type=int, default=16)
p.add_argument("--seq-len", type=int, default=128)
p.add_argument("--steps", type=int, default=20)
args = p.parse_args()
bps.init()
torch.cuda.set_device(bps.local_rank())
model = get_model(args.model).cuda().to(torch.bfloat16)
opt = bps.DistributedOptimizer(torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01),
                               named_parameters=model.named_parameters(), fused_update=True)
bps.broadcast_parameters(model.state_dict(), root_rank=0)
ids = torch.randint(0, 30522, (args.batch_size, args.seq_len), device="cuda")
labels = torch.randint(0, 30522, (args.batch_size, args.seq_len), device="cuda")
for i in range(args.steps + 5):
    if i == 5:
        torch.cuda.synchronize()
        t0 = time.time()
    opt.zero_grad()
    loss = model(ids, mlm_labels=labels)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
if bps.rank() == 0:
    dt = time.time() - t0
    print("tokens/sec on %d GPU(s): %.0f (loss %.3f)" % (bps.size(), bps.size() * args.batch_size * args.seq_len *
                                                          args.steps / dt, loss.item()))
bps.shutdown()
