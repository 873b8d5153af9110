"""The fused transformer stack (csrc/token_tc.cu: all Transformer_Blocks of model/blocks.py:50-88 + encoder_norm as one
launch, one group of 16 CTAs per panorama) against a float64 torch restatement of the blocks, phase by
phase (the exchange buffers after k GEMM phases) and end to end, at every token count the BASELINE configs use."""
import numpy as np
import pytest
import torch

from omnifusion_b200 import _lib
from omnifusion_b200.checkpoint import synthetic_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NROWS = {18: 4, 26: 5, 46: 6, 10: 3}
_nets = {}


def net_for(n_tok):
    if n_tok not in _nets:
        from omnifusion_b200.model.spherical_model_iterative import spherical_fusion
        net = spherical_fusion(NROWS[n_tok], n_tok, (128, 128), (80, 80))
        net.load_state_dict(synthetic_state_dict("iterative", n_tok, 0))
        net = net.to(DEV).eval()
        net._ensure_handle(torch.device(DEV))
        net._ensure_weights()
        _nets[n_tok] = net
    return _nets[n_tok]


def block_ref(sd, i, x, n_tok):
    """One Transformer_Block in float64; returns the intermediate tensors the kernel exchanges."""
    p = f"transformer.layer.{i}."
    w = lambda k: sd[p + k].double()
    ln = lambda t, g, b, eps: torch.nn.functional.layer_norm(t, (512,), g, b, eps)
    B = x.shape[0] // n_tok
    h = ln(x, w("norm1.weight"), w("norm1.bias"), 1e-5)
    q = h @ w("attn.q.weight").T
    kv = h @ w("attn.kv.weight").T
    k, v = kv[:, :512], kv[:, 512:]
    sh = lambda t: t.reshape(B, n_tok, 4, 128).permute(0, 2, 1, 3)
    s = sh(q) @ sh(k).transpose(-2, -1)                           # (B, 4, N, N) raw scores
    att = ((s * 128 ** -0.5).softmax(-1) @ sh(v)).transpose(1, 2).reshape(B * n_tok, 512)
    y = x + att @ w("attn.proj.weight").T + w("attn.proj.bias")
    h2 = ln(y, w("norm2.weight"), w("norm2.bias"), 1e-5)
    f1 = torch.nn.functional.gelu(h2 @ w("mlp.fc1.weight").T + w("mlp.fc1.bias"))
    x2 = y + f1 @ w("mlp.fc2.weight").T + w("mlp.fc2.bias")
    return dict(scores=s, att=att, y=y, x=x2)


def run(net, x, B, n_tok, nblk, stop):
    L = _lib.lib()
    nscr = L.ofb_token_stack_scratch_floats(B, n_tok)
    scratch = torch.full((nscr,), float("nan"), device=DEV)
    xd = x.float().to(DEV).contiguous()
    enc = torch.full((B * n_tok, 512), float("nan"), device=DEV)
    _lib.check(L.ofb_token_stack_f32(net._handle, _lib.ptr(xd), _lib.ptr(scratch), nscr, _lib.ptr(enc), B, n_tok, nblk,
                                     stop, _lib.stream_of(torch.device(DEV))))
    torch.cuda.synchronize()
    s_end = B * 16 * n_tok * n_tok
    a_end = s_end + B * n_tok * 512
    return dict(x=xd.cpu().double(), enc=enc.cpu().double(),
                scores=scratch[:s_end].reshape(B, 4, 4, n_tok, n_tok).cpu().double(),
                att=scratch[s_end:a_end].reshape(B * n_tok, 512).cpu().double(),
                part=scratch[a_end:].reshape(B, 16, n_tok, 512).cpu().double())


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


def test_device_runs_at_least_eight_panoramas_at_once():
    n = _lib.lib().ofb_token_stack_resident_groups(18)
    print(f"[token_stack] resident groups of 16 CTAs: {n} (NP=32), {_lib.lib().ofb_token_stack_resident_groups(46)} (NP=48)")
    assert n >= 8


@pytest.mark.parametrize("stop", [1, 2, 3, 4])
def test_first_block_phase_by_phase(stop):
    n_tok, B = 18, 3
    net = net_for(n_tok)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    x = torch.randn(B * n_tok, 512, generator=torch.Generator().manual_seed(5)).double()
    ref = block_ref(sd, 0, x, n_tok)
    got = run(net, x, B, n_tok, 1, stop)
    # phase 1 (qkv + attention): partial scores summed over the four 32-dim quarters, attention output
    e = rel(got["scores"].sum(2), ref["scores"])
    print(f"[parity] token_stack stop={stop}: scores {e:.2e}", end="")
    assert e < 2e-5
    e = rel(got["att"], ref["att"])
    print(f" att {e:.2e}", end="")
    assert e < 2e-5
    if stop >= 2:
        e = rel(got["x"], ref["y"] if stop < 4 else ref["x"])
        print(f" residual {e:.2e}", end="")
        assert e < 2e-5
    if stop >= 4:
        # the partial sums are raw accumulators of the power-of-two scaled weights: compare up to that factor
        fc2 = (ref["x"] - ref["y"] - sd["transformer.layer.0.mlp.fc2.bias"].double())
        raw = got["part"].sum(1).reshape(B * n_tok, 512)
        factor = 2.0 ** torch.log2((fc2 * raw).sum() / (raw * raw).sum()).round()
        e = rel(raw * factor, fc2)
        print(f" fc2 partial sums {e:.2e}", end="")
        assert e < 5e-5
    print()


@pytest.mark.parametrize("n_tok,B", [(18, 8), (18, 1), (26, 5), (46, 3), (10, 2), (18, 11), (18, 23)])
def test_whole_stack_matches_float64_blocks(n_tok, B):
    net = net_for(n_tok)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    x = torch.randn(B * n_tok, 512, generator=torch.Generator().manual_seed(n_tok + B)).double()
    ref = x
    for i in range(6):
        ref = block_ref(sd, i, ref, n_tok)["x"]
    enc_ref = torch.nn.functional.layer_norm(ref, (512,), sd["transformer.encoder_norm.weight"].double(),
                                             sd["transformer.encoder_norm.bias"].double(), 1e-6)
    got = run(net, x, B, n_tok, 6, 0)
    e1, e2 = rel(got["x"], ref), rel(got["enc"], enc_ref)
    print(f"[parity] token_stack N={n_tok} B={B}: residual stream {e1:.2e}, encoder_norm output {e2:.2e}")
    assert torch.isfinite(got["enc"]).all()
    assert e1 < 3e-5 and e2 < 3e-5
    # a panorama's result does not depend on the batch it is in
    one = run(net, x[:n_tok], 1, n_tok, 6, 0)
    assert torch.equal(one["enc"], got["enc"][:n_tok])
