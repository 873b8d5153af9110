"""End-to-end and stage-wise parity of the CUDA forward (through the reference-facing
nn.Module and the C ABI engine) against the CPU oracle and the committed goldens that
were produced by the real reference (tests/golden/make_golden.py).

Depth tolerance (BASELINE.json north_star): max |d - d_ref| / max(|d_ref|, 1e-6) <= 1e-3.
"""
import os

import numpy as np
import pytest
import torch

from omnifusion_b200.checkpoint import synthetic_state_dict
from oracle import model as om

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FOV = (80, 80)
NP = {3: 10, 4: 18, 5: 26, 6: 46}
REL_TOL = 1e-3


def urand(*shape, seed=0):
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed))


def max_rel(a, b):
    return ((a - b).abs() / b.abs().clamp_min(1e-6)).max().item()


_models = {}
FORMATS = {"split16_tcgen05": 1, "f32_cudacore": 0}


def model(kind, nrows, fmt=1):
    net = _model(kind, nrows)
    net.set_option("format", fmt)
    return net


def _model(kind, nrows):
    key = (kind, nrows)
    if key not in _models:
        if kind == "iterative":
            from omnifusion_b200.model.spherical_model_iterative import spherical_fusion
        else:
            from omnifusion_b200.model.spherical_model import spherical_fusion
        net = spherical_fusion(nrows, NP[nrows], (128, 128), FOV)
        net.load_state_dict(synthetic_state_dict(kind, NP[nrows], 0))
        _models[key] = net.to(DEV).eval()
    return _models[key]


def folded_to_nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("fmt", list(FORMATS), ids=list(FORMATS))
@pytest.mark.parametrize("confidence", [False, True])
def test_iterative_forward_matches_oracle_with_stages(confidence, fmt):
    net = model("iterative", 4, FORMATS[fmt])
    stage_tol = 2e-4 if FORMATS[fmt] == 0 else 5e-4
    sd = synthetic_state_dict("iterative", 18, 0)
    rgb = urand(2, 3, 64, 128, seed=123)
    trace = {}
    ref = om.forward_iterative(sd, rgb, 2, confidence, trace=trace)
    with torch.no_grad():
        got = net(rgb.to(DEV), iter=2, confidence=confidence)
    assert isinstance(got, list) and len(got) == 2
    for i, (g, r) in enumerate(zip(got, ref)):
        assert g.shape == r.shape == (2, 1, 64, 128)
        e = max_rel(g.cpu(), r)
        print(f"[parity] iterative {fmt} conf={confidence} iter{i}: max_rel={e:.3e}")
        assert e <= REL_TOL
    # stage-wise: intermediates of the last iteration against the oracle's trace
    t1 = trace["iter1"]
    for name in ["conv1", "pool", "layer1_pre", "layer1", "layer2", "layer3", "layer4",
                 "de_conv0_1", "de_conv1_1", "de_conv2_1", "de_conv3_1", "de_conv4_0"]:
        a = folded_to_nchw(net.activation(name).cpu())
        r = t1[name]
        e = ((a - r).abs().max() / r.abs().max()).item()
        print(f"[parity] stage {fmt} {name}: max_err/absmax={e:.3e}")
        assert e <= stage_tol, name
    enc = net.activation("encoded").cpu().reshape(2, 18, 512)
    e = ((enc - t1["encoded"]).abs().max() / t1["encoded"].abs().max()).item()
    print(f"[parity] stage {fmt} encoded: max_err/absmax={e:.3e}")
    assert e <= stage_tol
    tok = net.activation("tokens").cpu().permute(0, 3, 1, 2).reshape(2, 18, 512)
    assert ((tok - t1["tokens"]).abs().max() / t1["tokens"].abs().max()).item() <= stage_tol


def test_nrows3_forward_matches_oracle_including_uncovered_pixels():
    """nrows=3 (10 patches): the only layout that leaves ERP pixels uncovered; the reference returns
    0 there (pers2equi_v3.py:189-196 with all weights zero, spherical_model_iterative.py:376-378)."""
    net = model("iterative", 3)
    sd = synthetic_state_dict("iterative", 10, 0)
    rgb = urand(2, 3, 64, 128, seed=77)
    ref = om.forward_iterative(sd, rgb, 2, True, nrows=3)
    with torch.no_grad():
        got = net(rgb.to(DEV), iter=2, confidence=True)
    for i, (g, r) in enumerate(zip(got, ref)):
        e = max_rel(g.cpu(), r)
        print(f"[parity] nrows=3 iter{i}: max_rel={e:.3e}; uncovered pixels {(r == 0).sum().item()}")
        assert e <= REL_TOL
        assert torch.equal(g.cpu() == 0, r == 0)
    assert (ref[-1] == 0).any()


def test_single_stage_forward_matches_oracle():
    net = model("single", 4)
    sd = synthetic_state_dict("single", 18, 0)
    rgb = urand(2, 3, 64, 128, seed=123)
    for conf in (True, False):
        ref = om.forward_single(sd, rgb, conf)
        with torch.no_grad():
            got = net(rgb.to(DEV), confidence=conf)
        assert torch.is_tensor(got) and got.shape == (2, 1, 64, 128)
        e = max_rel(got.cpu(), ref)
        print(f"[parity] single-stage conf={conf}: max_rel={e:.3e}")
        assert e <= REL_TOL


GOLDEN_CASES = ["iter_small_conf0", "iter_small_conf1", "single_small_conf1", "single_small_conf0",
                "iter_n6_conf1", "iter_n5_conf1", "iter_full_conf1"]


@pytest.mark.parametrize("fmt", list(FORMATS), ids=list(FORMATS))
@pytest.mark.parametrize("tag", GOLDEN_CASES)
def test_forward_matches_reference_goldens(tag, fmt, golden_dir):
    """Outputs of the REAL reference (committed fixtures) vs the CUDA path."""
    z = np.load(os.path.join(golden_dir, f"model_{tag}.npz"))
    kind, nrows = str(z["kind"]), int(z["nrows"])
    net = model(kind, nrows, FORMATS[fmt])
    erp = tuple(int(v) for v in z["erp"])
    rgb = urand(int(z["bs"]), 3, *erp, seed=int(z["seed"])).to(DEV)
    s = int(z["stride"])
    with torch.no_grad():
        if kind == "iterative":
            got = net(rgb, iter=int(z["iters"]), confidence=bool(z["conf"]))
        else:
            got = [net(rgb, confidence=bool(z["conf"]))]
    for i, g in enumerate(got):
        ref = torch.from_numpy(z[f"out{i}"])
        e = max_rel(g.cpu()[:, :, ::s, ::s], ref)
        print(f"[parity] golden {fmt} {tag} iter{i}: max_rel={e:.3e}")
        assert e <= REL_TOL
        assert abs(g.double().mean().item() - float(z[f"out{i}_mean"])) <= REL_TOL * abs(float(z[f"out{i}_mean"]))


@pytest.mark.parametrize("fmt", list(FORMATS), ids=list(FORMATS))
def test_batch_invariance_chunking_dedup_and_graph(fmt):
    net = model("iterative", 4, FORMATS[fmt])
    rgb = urand(5, 3, 64, 128, seed=7).to(DEV)
    with torch.no_grad():
        base = [t.clone() for t in net(rgb, iter=2, confidence=True)]
        one = net(rgb[3:4], iter=2, confidence=True)
        assert torch.equal(one[1], base[1][3:4]), "panoramas must be independent (eval-mode BN, per-panorama attention)"
        net.set_option("chunk", 2)
        chunked = net(rgb, iter=2, confidence=True)
        assert torch.equal(chunked[1], base[1])
        net.set_option("chunk", 0)
        net.set_option("dedup", 0)
        full = net(rgb, iter=2, confidence=True)
        assert torch.equal(full[1], base[1]), "re-using the iteration-invariant stem must not change results"
        net.set_option("dedup", 1)
        # option lanes=2: the batch runs as two concurrent halves on two streams; must agree bit for bit
        net.set_option("lanes", 2)
        two_lanes = net(rgb, iter=2, confidence=True)
        net.set_option("lanes", 1)
        assert torch.equal(two_lanes[0], base[0]) and torch.equal(two_lanes[1], base[1])
        if FORMATS[fmt] == 1:
            # the heads run on the tensor pipe; the CUDA-core heads kernel must agree to rounding
            net.set_option("heads_tc", 0)
            cc = net(rgb, iter=2, confidence=True)
            net.set_option("heads_tc", 1)
            d = ((cc[1] - base[1]).abs() / base[1].abs().clamp_min(1e-6)).max().item()
            print(f"[parity] tensor-core vs CUDA-core heads: depth max rel diff {d:.3e}")
            assert d <= 2e-5
            # the last decoder upsample is folded into de_conv4_0's operand producer; unfused must agree
            net.set_option("fuse_ups", 0)
            unfused = net(rgb, iter=2, confidence=True)
            net.set_option("fuse_ups", 1)
            d = ((unfused[1] - base[1]).abs() / base[1].abs().clamp_min(1e-6)).max().item()
            print(f"[parity] fused vs materialised upsample: depth max rel diff {d:.3e}")
            assert d <= 2e-6
        g = net.forward_graphed(rgb, 2, True)
        assert torch.equal(g[1], base[1])
        g2 = net.forward_graphed(rgb.flip(0).contiguous(), 2, True)
        assert torch.equal(g2[1], base[1].flip(0))


def test_module_prefix_checkpoint_and_errors():
    from omnifusion_b200 import _lib
    from omnifusion_b200.model.spherical_model_iterative import spherical_fusion
    sd = synthetic_state_dict("iterative", 18, 0)
    net = spherical_fusion(4, 18, (128, 128), FOV)
    net.load_state_dict({"module." + k: v for k, v in sd.items()})        # DataParallel checkpoint (test.py:107-110)
    net = net.to(DEV)
    rgb = urand(1, 3, 64, 128, seed=123).to(DEV)
    with torch.no_grad():
        out = net(rgb, iter=1)
        ref = model("iterative", 4)(rgb, iter=1)
    assert torch.equal(out[0], ref[0])
    with pytest.raises(_lib.OfbError):
        net(rgb.cpu(), iter=1)
    net.train()
    with pytest.raises(RuntimeError):
        net(rgb, iter=1)
    with pytest.raises(ValueError):
        spherical_fusion(4, 18, (256, 256), FOV)
