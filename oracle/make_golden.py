"""Generates tests/golden/* from the REAL reference (run only where /root/reference exists, i.e. the build
container):  python -m oracle.make_golden

1. imports /root/reference/argus/models.py unmodified, with torchvision's weight download replaced by random init
   (the literal constructor needs the network: models.py:43 `weights="DEFAULT"`);
2. asserts oracle.ref_model.RefNCameraCNN is indistinguishable from it (keys, shapes, bit-identical forward / grads);
3. writes: state_dict key/shape/dtype list; model golden vectors (seed 42 weights, seed 0 inputs) for the small
   CPU config; loss golden vectors from the pinned float64 loss oracle.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def import_reference():
    import torchvision.models as tvm

    orig = tvm.resnet50
    tvm.resnet50 = lambda *a, **k: orig(weights=None)
    sys.path.insert(0, "/root/reference")
    try:
        from argus.models import NCameraCNN, NCameraCNNConfig  # noqa
    finally:
        sys.path.remove("/root/reference")
    return NCameraCNN, NCameraCNNConfig, (tvm, orig)


def main():
    from oracle.ref_model import RefNCameraCNN, torch_loss
    from oracle import se3_loss

    NCameraCNN, NCameraCNNConfig, (tvm, orig) = import_reference()
    torch.manual_seed(42)
    ref = NCameraCNN(NCameraCNNConfig())
    tvm.resnet50 = orig
    torch.manual_seed(42)
    mine = RefNCameraCNN()
    sd_ref, sd_mine = ref.state_dict(), mine.state_dict()
    assert list(sd_ref.keys()) == list(sd_mine.keys())
    for k in sd_ref:
        assert sd_ref[k].shape == sd_mine[k].shape and sd_ref[k].dtype == sd_mine[k].dtype
        assert torch.equal(sd_ref[k], sd_mine[k]), k
    keys = [{"name": k, "shape": list(v.shape), "dtype": str(v.dtype).replace("torch.", "")} for k, v in sd_ref.items()]
    GOLDEN.mkdir(parents=True, exist_ok=True)
    (GOLDEN / "state_dict_keys.json").write_text(json.dumps(keys, indent=0))
    print("state_dict entries:", len(keys))

    # ---- model golden vectors: B=2 pairs at 128x128 (CPU-cheap): eval forward at init, train-mode forward + backward,
    # then eval forward with the updated running statistics
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 6, 128, 128, generator=g)
    q = torch.randn(2, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    target = torch.cat([torch.randn(2, 3, generator=g), q], -1)
    outs = {}
    for name, model in (("reference", ref), ("restatement", mine)):
        model.eval()
        with torch.no_grad():
            y_init = model(x)
        model.train()
        model.zero_grad()
        y = model(x)
        loss = torch_loss(y, target).mean()
        loss.backward()
        model.eval()
        with torch.no_grad():
            y_eval = model(x)
        outs[name] = {
            "eval_out_init": y_init.double().numpy().tolist(),
            "train_out": y.detach().double().numpy().tolist(),
            "loss": float(loss),
            "eval_out_after_step0": y_eval.double().numpy().tolist(),
            "grad_norms": {k: float(p.grad.double().norm()) for k, p in model.named_parameters()
                           if k in ("resnet.conv1.weight", "resnet.layer1.0.conv2.weight", "resnet.layer2.0.downsample.0.weight",
                                    "resnet.layer3.2.bn2.weight", "resnet.layer4.2.conv3.weight", "resnet.fc.weight",
                                    "resnet.fc.bias", "output_mlp.0.weight", "output_mlp.4.bias")},
            "running_mean_bn1_first4": model.resnet.bn1.running_mean[:4].double().numpy().tolist(),
        }
    assert outs["reference"] == outs["restatement"], "restatement differs from the reference module"
    # cross-check the torch loss against the pinned numpy oracle
    l_np = se3_loss.geometric_loss(np.array(outs["reference"]["train_out"]), target.double().numpy()).mean()
    assert abs(l_np - outs["reference"]["loss"]) < 1e-9, (l_np, outs["reference"]["loss"])
    gold = {"seed_weights": 42, "seed_inputs": 0, "shape": [2, 6, 128, 128], "target": target.double().numpy().tolist(),
            **outs["reference"]}
    (GOLDEN / "model_small.json").write_text(json.dumps(gold))
    print("model golden: loss", gold["loss"], "train_out[0]", gold["train_out"][0])

    # ---- loss golden vectors (float64 oracle pinned against expm/logm + the reference's known-answer test)
    rng = np.random.default_rng(7)
    n = 48
    pred = rng.normal(size=(n, 6)) * np.repeat(np.array([1e-7, 1e-3, 0.3, 1.0, 2.5, 4.0]), n // 6)[:, None]
    qq = rng.normal(size=(n, 4))
    qq /= np.linalg.norm(qq, axis=-1, keepdims=True)
    tgt = np.concatenate([rng.normal(size=(n, 3)), qq], -1)
    # identity-error rows: target = Exp(pred)  (reference tests/test_train.py:32-36)
    tgt[::6] = se3_loss.se3_exp(pred[::6])
    loss, grad = se3_loss.geometric_loss_and_grad(pred, tgt)
    lm = np.array([se3_loss.geometric_loss_matrix(pred[i], tgt[i]) for i in range(n)])
    assert np.abs(loss - lm).max() < 1e-8, np.abs(loss - lm).max()
    (GOLDEN / "loss_vectors.json").write_text(json.dumps({
        "pred": pred.tolist(), "target": tgt.tolist(), "loss": loss.tolist(), "grad": grad.tolist(),
        "pose": se3_loss.se3_exp(pred).tolist()}))
    print("loss golden:", n, "vectors; max loss", loss.max())


if __name__ == "__main__":
    main()
