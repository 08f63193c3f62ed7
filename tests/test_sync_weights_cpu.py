"""Host logic of NCameraCNN.sync_weights(): every in-place update made through the module tree must change the
fingerprint that decides whether the packed bf16 weights / eval BN fold are refreshed (no GPU needed)."""
import torch


def test_state_version_sees_updates_through_the_module_tree():
    from argus_b200.models import NCameraCNN

    m = NCameraCNN()
    seen = [m._state_version()]
    with torch.no_grad():
        m.resnet.conv1.weight.add_(1.0)                      # manual edit of one parameter view
    seen.append(m._state_version())
    m.load_state_dict({k: v.clone() for k, v in m.state_dict().items()})   # load_state_dict copies in place
    seen.append(m._state_version())
    m.resnet.bn1.running_mean.add_(1.0)                      # buffer edit (eval-mode BN fold depends on it)
    seen.append(m._state_version())
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)          # a stock optimizer on model.parameters()
    for p in m.parameters():
        p.grad = torch.ones_like(p)
    opt.step()
    seen.append(m._state_version())
    assert all(a < b for a, b in zip(seen, seen[1:])), seen
    # the views still alias the flat arena the C library is bound to
    assert m.resnet.conv1.weight.data_ptr() == m.flat_params.data_ptr()
