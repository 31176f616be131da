"""ShardedTower on real GPUs over NCCL: 2 ranks, one process per GPU, launched with torchrun exactly as
the driver launches bench.py.  Needs >= 2 GPUs (skipped on a 1-GPU box; run with `gpurun --gpus 2`)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from tensorflow_web_deploy_b200 import engine, sharding, weights
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
C, B = 6, 4
path = os.path.join(%(tmp)r, "w%%d.icpw" %% rank)
weights.save_blob(path, weights.synthetic_weights(C))
eng = engine.Engine(path, C, device=lr, max_batch=B, precision="fp16", use_cuda_graph=True)
dev = torch.device("cuda", lr)
tower = sharding.ShardedTower(None, C, dist=dist, device=dev, engine=eng, sweet_spot=B, depth=2)
# device path: every rank owns B rows; five steps through two rotating buffers, gathers on the side stream
rng = np.random.default_rng(100 + rank)
xs = [torch.from_numpy(rng.uniform(-1, 1, (B, 299, 299, 3)).astype(np.float32)).to(dev) for _ in range(5)]
stream = torch.cuda.Stream(device=dev)
with torch.cuda.stream(stream):
    handles = [tower.step_device(x, stream) for x in xs]
    outs = [h.synchronize().clone().cpu().numpy() for h in handles[-2:]]      # the two buffers still alive
mine = [eng.forward(x.cpu().numpy()) for x in xs[-2:]]
ok_dev = all(np.array_equal(o[rank * B:(rank + 1) * B], m) for o, m in zip(outs, mine))
# every rank must hold every rank's rows: compare with an all_gather of the local results
gl = [torch.zeros((B, C), device=dev) for _ in range(world)]
dist.all_gather(gl, torch.from_numpy(mine[-1]).to(dev))
ok_all = np.array_equal(outs[-1], torch.cat(gl).cpu().numpy())
# host path: a batch larger than one GPU's sweet spot is split by rows and reassembled on every rank
xh = np.random.default_rng(7).uniform(-1, 1, (2 * B - 1, 299, 299, 3)).astype(np.float32)
full = tower.classify(xh)
lo, hi = sharding.shard_bounds(xh.shape[0], sharding.gpus_for_batch(xh.shape[0], world, B), rank)
ok_host = full.shape == (2 * B - 1, C) and np.array_equal(full[lo:hi], eng.forward(xh[lo:hi]))
small = tower.classify(xh[:2])                      # fits one GPU: rank 0 computes, everyone gets it
ok_small = small.shape == (2, C)
print("RESULT " + json.dumps({"rank": rank, "ok_dev": bool(ok_dev), "ok_all": bool(ok_all), "ok_host": bool(ok_host),
                              "ok_small": bool(ok_small), "sum": float(full.sum())}), flush=True)
eng.close()
dist.barrier()
dist.destroy_process_group()
'''


def test_sharded_tower_two_ranks_nccl(tmp_path, built_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "tmp": str(tmp_path)})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29631", str(script)], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    res = [json.loads(l.split("RESULT ", 1)[1]) for l in r.stdout.splitlines() if "RESULT " in l]
    assert len(res) == 2
    for x in res:
        assert x["ok_dev"] and x["ok_all"] and x["ok_host"] and x["ok_small"], x
    assert res[0]["sum"] == res[1]["sum"]
