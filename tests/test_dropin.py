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


def test_flat_group_adam_equals_per_parameter_adam():
    """TrainStep re-homes the parameters as views of one flat tensor and runs Adam on that tensor: elementwise the
    same arithmetic, and a late-activated parameter (the gates of train.py:169-171) gets its own flat tensor with its
    own step count, as torch.optim.Adam gives a parameter whose first gradient arrives late."""
    import copy
    import torch
    from medt_b200.train_step import FlatGroup
    torch.manual_seed(0)
    net_a = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    gate_a = torch.nn.Parameter(torch.tensor(0.1), requires_grad=False)
    net_b, gate_b = copy.deepcopy(net_a), copy.deepcopy(gate_a)
    kw = dict(lr=1e-2, weight_decay=1e-3)
    opt_a = torch.optim.Adam(list(net_a.parameters()) + [gate_a], **kw)
    g0 = FlatGroup(list(net_b.parameters()))
    assert all(p.data_ptr() % 16 == 0 for p in g0.params if p.numel() % 4 == 0)
    opt_b = torch.optim.Adam([g0.flat], **kw)
    groups = [g0]
    xs = [torch.randn(4, 6) for _ in range(5)]
    for it, x in enumerate(xs):
        if it == 3:                                   # the epoch-10 flip
            gate_a.requires_grad_(True)
            gate_b.requires_grad_(True)
            groups.append(FlatGroup([gate_b]))
            opt_b.param_groups[0]["params"].append(groups[-1].flat)
        opt_a.zero_grad(set_to_none=True)
        (net_a(x) * gate_a).pow(2).sum().backward()
        opt_a.step()
        for p in list(net_b.parameters()) + [gate_b]:
            p.grad = None
        (net_b(x) * gate_b).pow(2).sum().backward()
        for g in groups:
            g.gather_grads_()
            g.publish_grads_()
        opt_b.step()
        for pa, pb in zip(list(net_a.parameters()) + [gate_a], list(net_b.parameters()) + [gate_b]):
            assert torch.equal(pa.detach(), pb.detach()), it
    assert int(opt_b.state[groups[0].flat]["step"]) == 5 and int(opt_b.state[groups[1].flat]["step"]) == 2
    sd = net_b.state_dict()                            # still a normal module: state_dict / load_state_dict work in place
    net_b.load_state_dict({k: torch.zeros_like(v) for k, v in sd.items()})
    assert float(groups[0].flat.detach().abs().sum()) == 0.0


def test_inference_helpers_threshold_and_checkpoint_prefix(tmp_path):
    """medt_b200.infer: the reference's thresholding (test.py:124-146) and `module.`-prefix tolerant checkpoint loading."""
    import numpy as np
    import torch
    import lib
    from medt_b200.infer import Predictor, load_checkpoint, logits_to_mask
    torch.manual_seed(0)
    logits = torch.randn(3, 2, 5, 7)
    logits[0, 1, 0, 0] = 0.5                                    # boundary: >= 0.5 is foreground
    tmp = logits.numpy().copy()                                  # the reference's numpy sequence
    tmp[tmp >= 0.5] = 1
    tmp[tmp < 0.5] = 0
    want = tmp.astype(int)
    want[want == 1] = 255
    got = logits_to_mask(logits)
    assert got.dtype == torch.uint8 and np.array_equal(got.numpy(), want[:, 1])
    m = lib.models.axialnet.gated(img_size=32, imgchan=3)
    ref_state = {k: torch.randn_like(v) if v.is_floating_point() else v.clone() for k, v in m.state_dict().items()}
    torch.save({"module." + k: v for k, v in ref_state.items()}, tmp_path / "dp.pth")       # written under DataParallel
    load_checkpoint(m, str(tmp_path / "dp.pth"))
    assert all(torch.equal(v, ref_state[k]) for k, v in m.state_dict().items())
    dp = torch.nn.DataParallel(lib.models.axialnet.gated(img_size=32, imgchan=3))
    load_checkpoint(dp, ref_state)                                                            # plain -> wrapped model
    assert all(torch.equal(v, ref_state[k[len("module."):]]) for k, v in dp.state_dict().items())
    with pytest.raises(RuntimeError, match="no CPU path"):
        Predictor(m, (1, 3, 32, 32))
