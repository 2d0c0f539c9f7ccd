#!/usr/bin/env python
"""K12 measurement: optimiser steps per second of the device trainer on the reference's default regressor
(SimpleFC(4096, [264,128,64], 1), batch 16, Adam) next to the reference's own loop (torch autograd + torch.optim.Adam)
on the same GPU (eager CUDA) and on the host CPU.  One JSON line."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clip_assisted_data_labeling_b200.scorer import SimpleFC  # noqa: E402
from clip_assisted_data_labeling_b200.trainer import DeviceTrainer  # noqa: E402


def torch_loop(model, feats, labels, order, batch, steps):
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, weight_decay=6e-4)
    crit = torch.nn.MSELoss()
    model.train()
    t0 = time.perf_counter()
    for s in range(steps):
        idx = order[s * batch:(s + 1) * batch]
        opt.zero_grad()
        loss = crit(model(feats[idx]).squeeze(), labels[idx])
        loss.backward()
        opt.step()
    if feats.is_cuda:
        torch.cuda.synchronize()
    return steps / (time.perf_counter() - t0)


def main():
    D, hidden, batch, n = 4096, [264, 128, 64], 16, 16000
    torch.manual_seed(0)
    feats, labels = torch.randn(n, D), torch.rand(n)
    order = torch.randperm(n)
    model = SimpleFC(D, hidden, 1, ["M/x"], dropout_prob=0.5)
    tr = DeviceTrainer(model, max_batch=batch, dropout_p=0.5, seed=1)
    fd, ld = feats.cuda(), labels.cuda()
    tr.epoch(fd, ld, order[:batch * 20].tolist(), batch, 2e-4, 6e-4)  # warm-up
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tr.epoch(fd, ld, order.tolist(), batch, 2e-4, 6e-4)
    b.record()
    torch.cuda.synchronize()
    steps = n // batch
    ours = steps / (a.elapsed_time(b) / 1e3)
    import copy
    m_gpu = copy.deepcopy(model).cuda()
    torch_loop(m_gpu, fd, ld, order.cuda(), batch, 50)
    eager = torch_loop(m_gpu, fd, ld, order.cuda(), batch, 400)
    m_cpu = copy.deepcopy(model)
    torch.set_num_threads(os.cpu_count())
    torch_loop(m_cpu, feats, labels, order, batch, 20)
    cpu = torch_loop(m_cpu, feats, labels, order, batch, 200)
    print(json.dumps({"config": f"SimpleFC({D},{hidden},1) batch {batch} Adam dropout 0.5", "steps_per_s_b2c": ours,
                      "us_per_step_b2c": 1e6 / ours, "launches_per_step": 3 * (len(hidden) + 1),
                      "steps_per_s_torch_eager_cuda": eager, "steps_per_s_torch_cpu": cpu, "cpu_threads": os.cpu_count(),
                      "epoch_10k_samples_60_epochs_s_b2c": 60 * (10000 / batch) / ours}))


if __name__ == "__main__":
    main()
