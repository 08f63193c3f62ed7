"""GPU: the native double-buffered shard loader delivers exactly the file's samples, partitions them across ranks
like DistributedSampler (disjoint, padded by wrapping, reshuffled per epoch) and feeds the training engine."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shard(tmp_path_factory):
    from argus_b200.loader import write_shard

    rng = np.random.default_rng(0)
    n = 37
    imgs = rng.integers(0, 256, (n, 2, 64, 64, 3), dtype=np.uint8)
    imgs[:, 0, 0, 0, 0] = np.arange(n)  # tag every sample with its index
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    poses = np.concatenate([rng.normal(size=(n, 3)), q], -1).astype(np.float32)
    path = tmp_path_factory.mktemp("shard") / "data.argusraw"
    write_shard(str(path), imgs, poses)
    return str(path), imgs, poses


def collect(loader):
    out = []
    for images, poses in loader:
        out.append((images.cpu().numpy().copy(), poses.cpu().numpy().copy()))
    return out


def test_sequential_epoch_is_exact(cuda_device, shard):
    from argus_b200.loader import ShardLoader

    path, imgs, poses = shard
    loader = ShardLoader(path, batch_size=8, device=cuda_device, shuffle=False)
    assert (loader.n_samples, loader.n_cams, loader.H, loader.W) == (37, 2, 64, 64) and len(loader) == 5
    batches = collect(loader)
    assert [b[0].shape[0] for b in batches] == [8, 8, 8, 8, 5]
    assert np.array_equal(np.concatenate([b[0] for b in batches]), imgs)
    assert np.array_equal(np.concatenate([b[1] for b in batches]), poses)
    assert len(collect(loader)) == 5  # a second epoch restarts cleanly


def test_rank_partition_and_reshuffle(cuda_device, shard):
    from argus_b200.loader import ShardLoader

    path, imgs, poses = shard
    seen = []
    for rank in range(2):
        loader = ShardLoader(path, batch_size=4, device=cuda_device, rank=rank, world=2, seed=5, shuffle=True)
        assert loader.samples_per_rank == 19
        tags = np.concatenate([b[0][:, 0, 0, 0, 0] for b in collect(loader)])
        assert len(tags) == 19
        seen.append(tags)
        loader.set_epoch(1)
        tags1 = np.concatenate([b[0][:, 0, 0, 0, 0] for b in collect(loader)])
        assert not np.array_equal(tags, tags1)  # reshuffled per epoch
    both = np.concatenate(seen)
    assert set(both.tolist()) == set(range(37))      # every sample visited
    assert len(both) == 38                            # padded by wrapping to divide evenly, as DistributedSampler does
    # samples arrive intact (pose follows its image)
    loader = ShardLoader(path, batch_size=4, device=cuda_device, rank=1, world=2, seed=5, shuffle=True)
    for images, p in loader:
        idx = images[:, 0, 0, 0, 0].cpu().numpy()
        assert np.array_equal(p.cpu().numpy(), poses[idx])


def test_loader_feeds_engine(cuda_device, shard):
    from argus_b200.data import Augmentation, AugmentationConfig
    from argus_b200.engine import TrainEngine
    from argus_b200.loader import ShardLoader
    from argus_b200.models import NCameraCNN

    path, _, _ = shard
    torch.manual_seed(0)
    model = NCameraCNN().to(cuda_device)
    engine = TrainEngine(model, distributed=False, augmentation=Augmentation(AugmentationConfig(), train=True, seed=1))
    loader = ShardLoader(path, batch_size=8, device=cuda_device, shuffle=True, drop_last=True)
    losses = [engine.step(images, poses) for images, poses in loader]
    assert len(losses) == 4 and all(torch.isfinite(l) for l in losses)


def test_lookahead_loop_reads_each_buffer_before_it_is_reused(cuda_device, shard):
    """The training loop fetches batch k + 1 BEFORE it enqueues step k (argus_b200/train.py). With lookahead=True the
    copy of batch k + 2 into the buffer of batch k waits for the work enqueued up to that fetch -- here a deliberately
    slow consumer (a long chain of matmuls, then a clone of the batch): every clone must still hold batch k."""
    from argus_b200.loader import ShardLoader

    path, imgs, poses = shard
    loader = ShardLoader(path, batch_size=4, device=cuda_device, shuffle=False, lookahead=True)
    a = torch.randn(2048, 2048, device=cuda_device)

    def slow_consume(images, targets):
        b = a
        for _ in range(40):                 # ~ a few ms of GPU work before the batch is read
            b = (b @ a) * 1e-3
        return images.clone() + (b[0, 0] * 0).to(torch.uint8), targets.clone()

    got = []
    it = iter(loader)
    current = next(it, None)
    while current is not None:
        upcoming = next(it, None)           # fetch k + 1 first ...
        got.append(slow_consume(*current))  # ... then enqueue the work on batch k
        current = upcoming
    torch.cuda.synchronize()
    assert np.array_equal(np.concatenate([g[0].cpu().numpy() for g in got]), imgs)
    assert np.array_equal(np.concatenate([g[1].cpu().numpy() for g in got]), poses)
