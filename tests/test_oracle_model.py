"""CPU: the PyTorch restatement of the reference model (oracle/ref_model.py) against the golden vectors generated
from the REAL reference module (oracle/make_golden.py, run where /root/reference exists), and the state_dict
contract of our NCameraCNN mirror (326 keys, shapes, dtypes; strict load both ways)."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def gold():
    return json.loads((GOLDEN / "model_small.json").read_text())


def test_restatement_reproduces_reference_golden(gold):
    from oracle.ref_model import make_reference_model, torch_loss

    torch.set_num_threads(4)
    model = make_reference_model(gold["seed_weights"])
    g = torch.Generator().manual_seed(gold["seed_inputs"])
    x = torch.rand(*gold["shape"], generator=g)
    target = torch.tensor(gold["target"])
    model.eval()
    with torch.no_grad():
        y0 = model(x)
    assert np.allclose(y0.double().numpy(), gold["eval_out_init"], rtol=1e-5, atol=1e-7)
    model.train()
    y = model(x)
    loss = torch_loss(y, target).mean()
    assert np.allclose(y.detach().double().numpy(), gold["train_out"], rtol=1e-4, atol=1e-6)
    assert abs(float(loss) - gold["loss"]) < 1e-4 * abs(gold["loss"])
    loss.backward()
    for name, want in gold["grad_norms"].items():
        got = float(dict(model.named_parameters())[name].grad.double().norm())
        assert abs(got - want) <= 2e-3 * want + 1e-9, (name, got, want)


def test_reference_forward_contract():
    """Reference tests/test_model.py:7-17: 3-D input -> AssertionError; (2,6,256,256) -> (2,6) (checked at 64x64 to
    keep the CPU suite fast; the GPU suite runs the 256x256 case)."""
    from oracle.ref_model import RefNCameraCNN

    m = RefNCameraCNN().eval()
    with pytest.raises(AssertionError):
        m(torch.rand(6, 64, 64))
    with torch.no_grad():
        assert m(torch.rand(2, 6, 64, 64)).shape == (2, 6)


def test_state_dict_layout_matches_reference():
    from argus_b200.models import NCameraCNN, NCameraCNNConfig

    keys = json.loads((GOLDEN / "state_dict_keys.json").read_text())
    assert len(keys) == 326
    model = NCameraCNN(NCameraCNNConfig())
    sd = model.state_dict()
    assert [k["name"] for k in keys] == list(sd.keys())
    for k in keys:
        t = sd[k["name"]]
        assert list(t.shape) == k["shape"] and str(t.dtype).replace("torch.", "") == k["dtype"], k["name"]
    assert sum(p.numel() for p in model.parameters()) == 25_885_766


def test_state_dict_round_trip_with_reference_module():
    """validate.py:100-101 does `model.load_state_dict(torch.load(path))` strictly: both directions must work."""
    from argus_b200.models import NCameraCNN
    from oracle.ref_model import make_reference_model

    ref = make_reference_model(7)
    ours = NCameraCNN()
    res = ours.load_state_dict(ref.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for (n1, p1), (n2, p2) in zip(ref.state_dict().items(), ours.state_dict().items()):
        assert n1 == n2 and torch.equal(p1, p2), n1
    ref2 = make_reference_model(8)
    ref2.load_state_dict(ours.state_dict(), strict=True)
    assert torch.equal(ref2.resnet.layer3[2].conv2.weight, ref.resnet.layer3[2].conv2.weight)
    # parameters stay views of one flat arena (what the fused optimizer and the gradient buckets rely on)
    base = ours.flat_params.data_ptr()
    for p in ours.parameters():
        assert base <= p.data_ptr() < base + ours.flat_params.numel() * 4


def test_cpu_call_fails_loudly():
    from argus_b200 import _lib
    from argus_b200.models import NCameraCNN

    m = NCameraCNN()
    with pytest.raises(AssertionError):
        m(torch.rand(6, 64, 64))  # reference contract: non 4-D input asserts (models.py:76)
    with pytest.raises(_lib.ArgusError):
        m(torch.rand(1, 6, 64, 64))  # no CPU fallback
