"""A used model must pick up parameter / buffer updates made through torch (regression for the stale packed-weight
bug: the flat arena's version counter does not see updates made on the Parameter views)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_load_state_dict_after_first_forward(cuda_device):
    from argus_b200.models import NCameraCNN
    from gpu_util import structured_images

    x = structured_images(2, 6, 64, 64, 0, "cuda")
    torch.manual_seed(1)
    a = NCameraCNN().to("cuda").eval()
    torch.manual_seed(2)
    b = NCameraCNN().to("cuda").eval()
    with torch.no_grad():
        for m in (a, b):
            for name, buf in m.named_buffers():
                if name.endswith("running_var"):
                    buf.uniform_(0.5, 1.5)
                elif name.endswith("running_mean"):
                    buf.normal_(0.0, 0.1)
        ya, yb = a(x).clone(), b(x).clone()
        assert not torch.equal(ya, yb)
        a.load_state_dict(b.state_dict())        # second checkpoint into a model that has already run
        ya2 = a(x)
    assert torch.equal(ya2, yb), (ya2, yb)


def test_stock_optimizer_on_parameters_moves_the_output(cuda_device):
    """The advertised autograd path: loss.backward() + torch.optim.Adam(model.parameters())."""
    from argus_b200.loss import geometric_loss_fn
    from argus_b200.models import NCameraCNN
    from gpu_util import random_targets, structured_images

    torch.manual_seed(0)
    model = NCameraCNN().to("cuda").train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    x = structured_images(4, 6, 64, 64, 3, "cuda")
    t = random_targets(4, 5, "cuda")
    losses = []
    for _ in range(25):
        opt.zero_grad()
        loss = geometric_loss_fn(model(x), t).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.7 * losses[0], losses          # frozen packed weights would leave the loss where it started
