"""CPU tier: the `lib` package is a drop-in for the reference's: same factories,
state_dict schema (order, names, shapes, dtypes), frozen gates, bit-identical
initialisation under the same seed, and no silent CPU fallback."""
import json
import os

import pytest
import torch

import fill
from helpers import GOLDEN

SCHEMA = json.load(open(os.path.join(GOLDEN, "schema.json")))


def _factory(name):
    import lib
    return lib.models.axialunet if name == "axialunet" else getattr(lib.models.axialnet, name)


@pytest.mark.parametrize("tag", sorted(SCHEMA))
def test_state_dict_schema_and_init_match_reference(tag):
    name, S = tag.rsplit("_", 1)
    s = SCHEMA[tag]
    torch.manual_seed(1234)
    m = _factory(name)(img_size=int(S), imgchan=3)
    sd = m.state_dict()
    assert [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()] == s["keys"]
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            assert abs(float(v.double().sum()) - s["init_sum"][k]) < 1e-9, k
    assert [k for k, p in m.named_parameters() if not p.requires_grad] == s["requires_grad_false"]
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == s["trainable"]


def test_load_state_dict_roundtrip_with_dataparallel_prefix():
    m = _factory("gated")(img_size=64, imgchan=1)
    vals = fill.fill_state_dict(m.state_dict())
    sd = {k: torch.as_tensor(v).to(m.state_dict()[k].dtype) for k, v in vals.items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("flatten_index") for k in missing)
    dp = torch.nn.DataParallel(m)           # train.py:107 wraps like this; keys gain "module."
    assert all(k.startswith("module.") for k in dp.state_dict())


def test_cpu_forward_raises_instead_of_falling_back():
    m = _factory("gated")(img_size=32, imgchan=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.rand(1, 3, 32, 32))


def test_constructor_errors_match_reference():
    import lib
    with pytest.raises(AssertionError):
        lib.models.axialnet.AxialAttention_dynamic(12, 12, groups=8, kernel_size=8)
    with pytest.raises(ValueError):
        lib.models.axialnet.gated(img_size=32, replace_stride_with_dilation=[False])


def test_grouped_bn_matches_sequential_calls():
    """Semantics of the batched LoGo branch (oracle side): grouped statistics == one BN call per patch."""
    from oracle.grouped_bn_ref import batch_norm_grouped
    torch.manual_seed(0)
    x = torch.randn(12, 5, 4, 4, dtype=torch.float64)
    a, b = torch.nn.BatchNorm2d(5).double(), torch.nn.BatchNorm2d(5).double()
    with torch.no_grad():
        for bn in (a, b):
            bn.weight.copy_(torch.linspace(0.5, 1.5, 5)); bn.bias.copy_(torch.linspace(-0.2, 0.2, 5))
    ref = torch.cat([a(xb) for xb in x.chunk(4)], 0)
    got = batch_norm_grouped(x, b, 4)
    assert torch.allclose(ref, got, atol=1e-12)
    assert torch.allclose(a.running_mean, b.running_mean, atol=1e-12)
    assert torch.allclose(a.running_var, b.running_var, atol=1e-12)
    assert int(b.num_batches_tracked) == 4
