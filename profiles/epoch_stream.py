"""BASELINE.json configs[4]: end-to-end epoch over a dataset-shaped synthetic shard stream through the native
double-buffered loader (pinned host buffers, side-stream H2D) into the fused training engine.

    python profiles/epoch_stream.py [n_pairs] [batch]            (single GPU)
    torchrun --nproc-per-node N profiles/epoch_stream.py ...      (data parallel)

The reference names its dataset splits small / medium / large without sizes (README.md:52-54); SURVEY.md assumes
medium = 131072 pairs (51.5 GB of uint8). The default here is 8192 pairs (3.2 GB) so that writing the shard takes
seconds; throughput is size-independent once the file is larger than the two staging buffers."""
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from argus_b200.data import Augmentation, AugmentationConfig  # noqa: E402
from argus_b200.engine import TrainEngine  # noqa: E402
from argus_b200.loader import ShardLoader, write_synthetic_shard  # noqa: E402
from argus_b200.models import NCameraCNN  # noqa: E402

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import datetime

    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=300))
path = os.path.join(tempfile.gettempdir(), f"argus_synth_{n_pairs}.argusraw")
t0 = time.time()
if rank == 0 and not os.path.exists(path):
    write_synthetic_shard(path, n_pairs, fast=n_pairs > 16384)
if world > 1:
    dist.barrier()
t_write = time.time() - t0

torch.manual_seed(42)
model = NCameraCNN().to(dev)
engine = TrainEngine(model, augmentation=Augmentation(AugmentationConfig(), train=True, seed=1 + rank, gpu_spaghetti=True))
loader = ShardLoader(path, batch, dev, rank=rank, world=world, seed=0, shuffle=True, drop_last=True, lookahead=True)
for epoch in range(2):  # epoch 0 warms up (plans, page cache), epoch 1 is timed
    loader.set_epoch(epoch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    n = 0
    loss = None
    # look-ahead pipeline of argus_b200/train.py: fetch batch k + 1, enqueue step k, stage batch k + 1 on the side stream
    it = iter(loader)
    current = next(it, None)
    while current is not None:
        upcoming = next(it, None)
        loss = engine.step(*current)
        if upcoming is not None:
            engine.prefetch(upcoming[0])
        n += current[0].shape[0]
        current = upcoming
    torch.cuda.synchronize()
    dt = time.time() - t0
if world > 1:
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
if rank == 0:
    print(json.dumps({"workload": "configs[4] epoch stream", "pairs_in_shard": n_pairs, "world": world,
                      "per_gpu_batch": batch, "pairs_per_rank_epoch": n, "epoch_seconds": round(dt, 3),
                      "pairs_per_s": round(n * world / dt, 1), "shard_write_seconds": round(t_write, 1),
                      "final_loss": float(loss)}))
if world > 1:
    dist.destroy_process_group()
