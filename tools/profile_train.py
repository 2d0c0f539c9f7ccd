"""A few optimiser steps of the device trainer for ncu: python tools/profile_train.py [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from clip_assisted_data_labeling_b200.scorer import SimpleFC
from clip_assisted_data_labeling_b200.trainer import DeviceTrainer
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
D, hidden, batch = 4096, [264, 128, 64], 16
torch.manual_seed(0)
n = steps * batch
feats, labels = torch.randn(n, D).cuda(), torch.rand(n).cuda()
tr = DeviceTrainer(SimpleFC(D, hidden, 1, ["M/x"], dropout_prob=0.5), max_batch=batch, dropout_p=0.5, seed=1)
print(tr.epoch(feats, labels, list(range(n)), batch, 2e-4, 6e-4))
