"""N>1 on real GPUs (skipped on a 1-GPU box): tools/dist_check.py under torchrun over NCCL — the distributed duplicate
search must return exactly the single-GPU pair list, and sharded embedding must equal unsharded embedding."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_multi_gpu_dedup_and_embed_match_single_gpu():
    """One rank per VISIBLE GPU (2, 4 or 8: an 8-GPU lease verifies the 8-rank path), plus the 2-rank case when more are there."""
    import torch
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs >= 2 GPUs")
    for world in sorted({2, n_gpus}):
        port = 29600 + (os.getpid() + world) % 300
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
                            "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py")],
                           capture_output=True, text=True, timeout=900, cwd=ROOT)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        assert line["world"] == world and line["dedup_identical"] and line["store_dedup_identical"] and line["embed_identical"] and line["dedup_pairs"] > 0
