"""Reference checkpoint layout and the synthetic checkpoint used for benchmarks.

``key_spec`` enumerates the exact ``state_dict`` of the reference models
(/root/reference/model/spherical_model_iterative.py:254-305 and
model/spherical_model.py:191-235; 375 tensors, Conv3d weights carry a trailing
unit dim) so that real OmniFusion checkpoints load unchanged, including the
``module.`` prefix nn.DataParallel adds (test.py:107-110).

``synthetic_state_dict`` builds a deterministic, non-degenerate random
checkpoint (no network access for the real one): every tensor is drawn from its
own generator seeded by a hash of its name, so the result does not depend on
construction order or on which model variant asks for it.
"""
import hashlib
from collections import OrderedDict

import torch

RESNET34_BLOCKS = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))
DECODER = (("de_conv0_0", 512, 256), ("de_conv0_1", 512, 128), ("de_conv1_0", 128, 128),
           ("de_conv1_1", 256, 64), ("de_conv2_0", 64, 64), ("de_conv2_1", 128, 64),
           ("de_conv3_0", 64, 64), ("de_conv3_1", 128, 32), ("de_conv4_0", 32, 32))
EMB, DEPTH, HEADS, MLP_RATIO = 512, 6, 4, 4


def _bn(spec, p, c):
    spec[p + ".weight"] = (c,)
    spec[p + ".bias"] = (c,)
    spec[p + ".running_mean"] = (c,)
    spec[p + ".running_var"] = (c,)
    spec[p + ".num_batches_tracked"] = ()


def key_spec(kind="iterative", npatches=18):
    """OrderedDict name -> shape, in the reference's registration order."""
    # "test": the 256x256-patch variant of network_test.py:271 (down1 512 -> 8 over 8x8 positions), otherwise = iterative
    assert kind in ("iterative", "single", "test")
    s = OrderedDict()
    s["conv1.weight"] = (64, 3, 7, 7, 1)
    _bn(s, "bn1", 64)
    cin = 64
    for li, (c, nblk, stride) in enumerate(RESNET34_BLOCKS, 1):
        for b in range(nblk):
            p = f"layer{li}.{b}"
            s[p + ".conv1.weight"] = (c, cin if b == 0 else c, 3, 3, 1)
            _bn(s, p + ".bn1", c)
            s[p + ".conv2.weight"] = (c, c, 3, 3, 1)
            _bn(s, p + ".bn2", c)
            if b == 0 and (stride != 1 or cin != c):
                s[p + ".downsample.0.weight"] = (c, cin, 1, 1, 1)
                _bn(s, p + ".downsample.1", c)
        cin = c
    down = "down" if kind == "single" else "down1"
    dch = EMB // 64 if kind == "test" else EMB // 16
    s[down + ".weight"] = (dch, 512, 1, 1, 1)
    s[down + ".bias"] = (dch,)
    s["transformer.pos_emb"] = (1, npatches, EMB)
    for i in range(DEPTH):
        p = f"transformer.layer.{i}"
        s[p + ".norm1.weight"] = (EMB,)
        s[p + ".norm1.bias"] = (EMB,)
        s[p + ".attn.q.weight"] = (EMB, EMB)
        s[p + ".attn.kv.weight"] = (2 * EMB, EMB)
        s[p + ".attn.proj.weight"] = (EMB, EMB)
        s[p + ".attn.proj.bias"] = (EMB,)
        s[p + ".norm2.weight"] = (EMB,)
        s[p + ".norm2.bias"] = (EMB,)
        s[p + ".mlp.fc1.weight"] = (MLP_RATIO * EMB, EMB)
        s[p + ".mlp.fc1.bias"] = (MLP_RATIO * EMB,)
        s[p + ".mlp.fc2.weight"] = (EMB, MLP_RATIO * EMB)
        s[p + ".mlp.fc2.bias"] = (EMB,)
    s["transformer.encoder_norm.weight"] = (EMB,)
    s["transformer.encoder_norm.bias"] = (EMB,)
    for name, ci, co in DECODER:
        s[name + ".conv.weight"] = (co, ci, 3, 3, 1)
        _bn(s, name + ".bn", co)
    for head in ("pred", "weight_pred"):
        s[head + ".weight"] = (1, 32, 3, 3, 1)
        s[head + ".bias"] = (1,)
    mlps = ("mlp_points",) if kind == "single" else ("mlp_points1", "mlp_points2")
    for m in mlps:
        s[m + ".0.weight"] = (16, 5 if kind == "single" else 3, 1, 1)
        _bn(s, m + ".1", 16)
        s[m + ".3.weight"] = (64, 16, 1, 1)
        _bn(s, m + ".4", 64)
    return s


def _gen(name, seed):
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return torch.Generator().manual_seed(int.from_bytes(h[:7], "little"))


def synthetic_state_dict(kind="iterative", npatches=18, seed=0):
    """Deterministic random checkpoint with the reference key set.

    Conv / linear weights ~ N(0, gain/fan_in) (He-style so activations keep their
    scale through ReLUs), BN statistics perturbed around identity, heads scaled so
    that depth is positive and varies and the confidence logits are not saturated.
    """
    sd = OrderedDict()
    for name, shape in key_spec(kind, npatches).items():
        g = _gen(name, seed)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t = torch.tensor(1000, dtype=torch.long)
        elif leaf == "running_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            t = 0.75 + 0.5 * torch.rand(shape, generator=g)
        elif name == "transformer.pos_emb":
            t = 0.02 * torch.randn(shape, generator=g)
        elif leaf == "weight" and len(shape) == 1:       # BN / LayerNorm scale
            t = 0.75 + 0.5 * torch.rand(shape, generator=g)
            if ".bn2." in name:
                t = t * 0.35                               # residual branch: keep the running sum bounded
            elif name.startswith("de_conv"):
                t = t * 0.85
        elif leaf == "bias":
            t = 0.1 * torch.randn(shape, generator=g)
        else:                                             # conv / linear weight
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            gain = 1.0 if ("attn.proj" in name or "mlp.fc2" in name) else 2.0
            if name in ("pred.weight", "weight_pred.weight"):
                gain = 1.5 if name == "pred.weight" else 3.0
            t = torch.randn(shape, generator=g) * (gain / fan_in) ** 0.5
        sd[name] = t.float() if t.dtype != torch.long else t
    # heads: depth ~ O(1) with spread, confidence logits O(1)
    sd["pred.bias"] = torch.tensor([2.0])
    sd["weight_pred.bias"] = torch.tensor([-3.0])
    return sd


def strip_module_prefix(sd):
    if len(sd) and all(k.startswith("module.") for k in sd):
        return OrderedDict((k[len("module."):], v) for k, v in sd.items())
    return sd


def rescale_activations(sd, s, kind="iterative"):
    """A checkpoint whose conv-tower activations are `s` times those of `sd` while the depth output is unchanged in
    exact arithmetic (up to the BatchNorm eps): every BatchNorm of the patch network gets gamma, beta * s and - when its
    input is already scaled - running_mean * s, running_var * s^2; the token path stays at its own scale (down1.weight / s,
    encoder_norm * s because its output is added to layer4); pred / weight_pred weights / s.  Used to probe the
    numeric range of the split-half activation format (tests/test_gpu_model.py)."""
    out = OrderedDict((k, v.clone()) for k, v in sd.items())
    bn = sorted({k[: -len(".running_var")] for k in sd if k.endswith(".running_var")})
    mlp = [p for p in bn if p.startswith("mlp_points")]
    for p in bn:
        # first BN of a point MLP feeds only the second one: left alone; the second one's OUTPUT joins layer1 (scaled)
        if p in mlp and not p.endswith(".4"):
            continue
        out[p + ".weight"] *= s
        out[p + ".bias"] *= s
        if p != "bn1" and p not in mlp:                      # input already scaled by s
            out[p + ".running_mean"] *= s
            out[p + ".running_var"] *= s * s
    down = "down" if kind == "single" else "down1"
    out[down + ".weight"] /= s
    out["transformer.encoder_norm.weight"] *= s
    out["transformer.encoder_norm.bias"] *= s
    out["pred.weight"] /= s
    out["weight_pred.weight"] /= s
    return out
