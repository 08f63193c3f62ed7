"""GPU integration tests mirroring the reference's tests/test_train.py:39-77 and validate.py: one epoch on a dummy
dataset runs, writes `<id>.pth` in the reference layout, a second run with the same seed lands on the same model,
and validate() consumes the checkpoint."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dummy_data_path(tmp_path_factory):
    from argus_b200.dataset import write_dataset

    rng = np.random.default_rng(0)
    root = tmp_path_factory.mktemp("data") / "dummy"

    def poses(n):
        q = rng.normal(size=(n, 4))
        q /= np.linalg.norm(q, axis=-1, keepdims=True)
        return np.concatenate([rng.normal(size=(n, 3)), q], -1)

    yy, xx = np.mgrid[0:128, 0:128]
    base = (np.sin(xx / 11.0) * 60 + np.cos(yy / 7.0) * 50 + 120)[None, None, :, :, None]
    imgs = np.clip(base + rng.normal(0, 30, (24, 2, 128, 128, 3)) + rng.uniform(-50, 50, (24, 2, 1, 1, 3)), 0, 255)
    imgs = imgs.astype(np.uint8)
    write_dataset(str(root), imgs[:16], poses(16), imgs[16:], poses(8))
    return str(root)


def run_once(dummy_data_path, save_dir):
    from argus_b200.dataset import CameraCubePoseDatasetConfig
    from argus_b200.models import NCameraCNN
    from argus_b200.train import TrainConfig, train
    from argus_b200.validate import load_checkpoint

    cfg = TrainConfig(dataset_config=CameraCubePoseDatasetConfig(dataset_path=dummy_data_path, center_crop=(128, 128)),
                      batch_size=8, n_epochs=1, device="cuda", wandb_log=False, save_dir=str(save_dir), num_workers=0,
                      num_gpus=1)
    run_id = train(cfg)
    path = save_dir / f"{run_id}.pth"
    assert path.exists()
    model = NCameraCNN()
    load_checkpoint(model, str(path))
    model.to("cuda").eval()
    with torch.no_grad():
        return model(torch.ones(1, 6, 128, 128, device="cuda")), path


def test_train_writes_reference_layout_and_is_reproducible(cuda_device, dummy_data_path, tmp_path):
    import json
    from pathlib import Path

    out1, path = run_once(dummy_data_path, tmp_path)
    sd = torch.load(path, map_location="cpu", weights_only=True)
    keys = json.loads((Path(__file__).resolve().parent / "golden" / "state_dict_keys.json").read_text())
    assert [k["name"] for k in keys] == list(sd.keys())
    assert all(list(sd[k["name"]].shape) == k["shape"] for k in keys)
    assert int(sd["resnet.bn1.num_batches_tracked"]) == 2          # 16 samples / batch 8
    assert torch.isfinite(out1).all()
    out2, _ = run_once(dummy_data_path, tmp_path)
    # same seed => same augmentation draws, same shuffling, same init, and every reduction on the device is ordered:
    # the two runs agree bit for bit (the reference asserts allclose at default tolerance, test_train.py:69-77)
    assert torch.equal(out1, out2), (out1, out2)


def test_validate_consumes_checkpoint(cuda_device, dummy_data_path, tmp_path):
    from argus_b200.dataset import CameraCubePoseDatasetConfig
    from argus_b200.validate import ValConfig, validate

    _, path = run_once(dummy_data_path, tmp_path)
    res = validate(ValConfig(model_path=str(path), dataset_config=CameraCubePoseDatasetConfig(
        dataset_path=dummy_data_path, center_crop=(128, 128))))
    assert res["losses"].shape == (8,) and torch.isfinite(res["losses"]).all()
    assert res["poses"].shape == (8, 7)
    assert torch.allclose(res["poses"][:, 3:].norm(dim=-1), torch.ones(8), atol=1e-5)  # unit quaternions


def test_loss_decreases_when_overfitting(cuda_device):
    """A few hundred fused steps on one fixed batch drive the geometric loss down (optimizer + backward sanity)."""
    from argus_b200.engine import TrainEngine
    from argus_b200.models import NCameraCNN
    from gpu_util import random_targets, structured_images

    torch.manual_seed(0)
    model = NCameraCNN().to("cuda")
    engine = TrainEngine(model, lr=1e-3, max_grad_norm=1.0, distributed=False)
    x = structured_images(8, 6, 64, 64, 1, "cuda")
    t = random_targets(8, 2, "cuda")
    losses = [engine.step(x, t).item() for _ in range(60)]
    assert losses[-1] < 0.5 * losses[0], (losses[0], losses[-1])
    assert all(np.isfinite(losses))


def test_loss_curve_matches_reference_within_its_own_spread(cuda_device):
    """North star: "a matching ... loss curve". Acceptance criterion (profiles/loss_curve.py::acceptance): for every seed
    (task + initial weights) the reference's step body runs three ways -- fp32, fp32 channels_last, TF32 (PyTorch's default
    convolution arithmetic): the same math with other cuDNN kernels / summation orders, i.e. the reference's OWN
    run-to-run spread -- and every 100-step window mean of the bf16 product run must lie inside [min, max] of those three
    runs, widened by 15 %. The committed 3-seed x 1000-step run (profiles/r2_loss_curve_acceptance.json) is at most 7.7 %
    outside the band in 30 windows, while the band itself is 5-55 % wide (the trajectories are chaotic: the reference
    does not reproduce itself any better); here 2 seeds x 500 steps to bound the test time."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "profiles"))
    from loss_curve import acceptance

    res = acceptance(steps=500, seeds=(0, 1), slack=0.15)
    for seed, r in res["seeds"].items():
        print("seed", seed, "ours", [round(v, 4) for v in r["window_means"]["ours_bf16"]], "band",
              [(round(a, 4), round(b, 4)) for a, b in zip(r["band_lo"], r["band_hi"])], "outside", r["outside"])
        ours = r["window_means"]["ours_bf16"]
        assert ours[-1] < 0.25 * ours[0]                     # and it trains: the loss falls by 4x in 500 steps
    assert res["accepted"], res["worst_outside"]


def test_prefetch_is_bitwise_equivalent(cuda_device):
    """TrainEngine.prefetch() (augmentation + staging of the next batch on a side stream, double-buffered stem input)
    must not change a single bit of the training trajectory."""
    from argus_b200.data import Augmentation, AugmentationConfig
    from argus_b200.engine import TrainEngine
    from argus_b200.models import NCameraCNN
    from gpu_util import random_targets

    g = torch.Generator().manual_seed(3)
    batches = [(torch.randint(0, 256, (4, 2, 64, 64, 3), dtype=torch.uint8, generator=g).to("cuda"),
                random_targets(4, 10 + i, "cuda")) for i in range(3)]

    def run(prefetch):
        torch.manual_seed(0)
        model = NCameraCNN().to("cuda")
        eng = TrainEngine(model, lr=1e-3, distributed=False,
                          augmentation=Augmentation(AugmentationConfig(), train=True, seed=5))
        losses = []
        for i in range(6):
            losses.append(eng.step(*batches[i % 3]).clone())
            if prefetch and i + 1 < 6:
                eng.prefetch(batches[(i + 1) % 3][0])
        torch.cuda.synchronize()
        return torch.stack(losses), model.flat_params.clone()

    l0, p0 = run(False)
    l1, p1 = run(True)
    assert torch.equal(l0, l1), (l0, l1)
    assert torch.equal(p0, p1)
    # a prefetched batch that is then NOT the one trained on is simply discarded
    torch.manual_seed(0)
    model = NCameraCNN().to("cuda")
    eng = TrainEngine(model, lr=1e-3, distributed=False, augmentation=Augmentation(AugmentationConfig(), train=True, seed=5))
    eng.step(*batches[0])
    eng.prefetch(batches[2][0])
    assert torch.isfinite(eng.step(*batches[1])).all()


def test_prefetch_without_augmentation(cuda_device):
    """uint8 batches with no augmentation configured: prefetch() stages through the same kernel with apply = 0 and a NULL
    parameter table (regression: the kernel used to read the table unconditionally), bit-identical to the direct path."""
    from argus_b200.engine import TrainEngine
    from argus_b200.models import NCameraCNN
    from gpu_util import random_targets

    g = torch.Generator().manual_seed(4)
    batches = [(torch.randint(0, 256, (4, 2, 64, 64, 3), dtype=torch.uint8, generator=g).to("cuda"),
                random_targets(4, 20 + i, "cuda")) for i in range(2)]

    def run(prefetch):
        torch.manual_seed(0)
        model = NCameraCNN().to("cuda")
        eng = TrainEngine(model, lr=1e-3, distributed=False, augmentation=None)
        losses = []
        for i in range(4):
            losses.append(eng.step(*batches[i % 2]).clone())
            if prefetch and i + 1 < 4:
                eng.prefetch(batches[(i + 1) % 2][0])
        torch.cuda.synchronize()
        return torch.stack(losses), model.flat_params.clone()

    l0, p0 = run(False)
    l1, p1 = run(True)
    assert torch.isfinite(l0).all()
    assert torch.equal(l0, l1), (l0, l1)
    assert torch.equal(p0, p1)


def test_prefetch_with_shorter_last_batch(cuda_device):
    """The last batch of an epoch is usually shorter (DataLoader without drop_last, as in the reference, train.py:168-192).
    Staging it while the full-size step is still in flight must not touch that step's stem input: the staging buffers
    belong to the model, not to the per-batch-size plan (regression for the round-1 aliasing bug). A long side-stream
    delay (ARGUS_FUZZ_DELAY_US is read at library load, so the fuzz kernel is not used here) is emulated by issuing the
    prefetch BEFORE the step's kernels can have finished: the step is large enough to still be running."""
    from argus_b200.data import Augmentation, AugmentationConfig
    from argus_b200.engine import TrainEngine
    from argus_b200.models import NCameraCNN
    from gpu_util import random_targets

    g = torch.Generator().manual_seed(11)
    sizes = [16, 16, 5, 16, 3]
    batches = [(torch.randint(0, 256, (b, 2, 64, 64, 3), dtype=torch.uint8, generator=g).to("cuda"),
                random_targets(b, 30 + i, "cuda")) for i, b in enumerate(sizes)]

    def run(prefetch):
        torch.manual_seed(0)
        model = NCameraCNN().to("cuda")
        eng = TrainEngine(model, lr=1e-3, distributed=False,
                          augmentation=Augmentation(AugmentationConfig(), train=True, seed=9))
        losses = []
        for i in range(len(batches)):
            losses.append(eng.step(*batches[i]).clone())
            if prefetch and i + 1 < len(batches):
                eng.prefetch(batches[i + 1][0])
        torch.cuda.synchronize()
        return torch.stack(losses), model.flat_params.clone()

    l0, p0 = run(False)
    l1, p1 = run(True)
    assert torch.isfinite(l0).all()
    assert torch.equal(l0, l1), (l0, l1)
    assert torch.equal(p0, p1)
