"""Generate the committed golden fixtures by running the REAL reference.

Authoring-container only (needs /root/reference):

    python tests/golden/make_golden.py

Imports the unmodified reference through ref_shim, feeds it seeded synthetic
inputs and the synthetic checkpoint from omnifusion_b200.checkpoint, and writes
small .npz fixtures next to this file.  The fixtures pin the oracle
(tests/test_oracle_golden.py) and, through it, the CUDA path.  The reference has
no golden vectors of its own, so these outputs of the
reference itself are the pin.
"""
import json
import os
import shutil
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import ref_shim  # noqa: E402

ref_shim.install()
from equi_pers.equi2pers_v3 import equi2pers as ref_e2p  # noqa: E402
from equi_pers.pers2equi_v3 import pers2equi as ref_p2e  # noqa: E402
from model.spherical_model import spherical_fusion as RefSingle  # noqa: E402
from model.spherical_model_iterative import spherical_fusion as RefIter  # noqa: E402

from omnifusion_b200.checkpoint import synthetic_state_dict  # noqa: E402
from oracle import equi_pers as oe  # noqa: E402
from oracle import model as om  # noqa: E402

NP = {3: 10, 4: 18, 5: 26, 6: 46}
FOV = (80, 80)
meta = {"torch": torch.__version__, "numpy": np.__version__, "checks": {}}


def rand(shape, seed):
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def eq(a, b):
    """torch.equal that treats NaN == NaN: at 1024x2048 / nrows=5 the reference's own table holds NaN weights at
    two ERP pixels (cos_c == 0 exactly -> x/0 = inf -> inf * mask 0 = NaN, pers2equi_v3.py:114,137-140), and the
    restatement must reproduce them."""
    if a.is_floating_point():
        return bool(torch.equal(torch.isnan(a), torch.isnan(b)) and
                    torch.equal(torch.nan_to_num(a, nan=0.0), torch.nan_to_num(b, nan=0.0)))
    return bool(torch.equal(a, b))


def in_fresh_cwd(fn):
    d = ref_shim.fresh_cwd()
    try:
        return fn(d)
    finally:
        os.chdir(HERE)
        shutil.rmtree(d, ignore_errors=True)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        out[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: tuple(v.shape) for k, v in out.items()})


# ---------------------------------------------------------------- resamplers
def resampler_case(tag, nrows, erp, P, stride, seed):
    def run(d):
        img = rand((2, 3, *erp), seed)
        pers, xyz, uv, center_p = ref_e2p(img, FOV, nrows, (P, P))
        back = ref_p2e(pers, FOV, nrows, (P, P), erp, "golden")
        tab = torch.load(os.path.join(d, "grid", "golden.pth"))
        # the reference's sampling grid is internal; recover its taps through the oracle helper
        geo = oe.equi2pers_geometry(FOV, nrows, (P, P))
        x0, y0, _, _ = oe.grid_sample_taps(geo["grid"], *erp)
        o_pers, o_xyz, o_uv, o_cp = oe.equi2pers(img, FOV, nrows, (P, P))
        o_back = oe.pers2equi(pers, FOV, nrows, (P, P), erp)
        o_tab = oe.pers2equi_table(FOV, nrows, (P, P), erp)
        meta["checks"][tag] = {
            "e2p_equal": bool(torch.equal(pers, o_pers) and torch.equal(xyz, o_xyz)
                              and torch.equal(uv, o_uv) and torch.equal(center_p, o_cp)),
            "p2e_equal": eq(back, o_back),
            "table_equal": bool(all(eq(tab[k], o_tab[k]) for k in tab)),
            "reference_nan_pixels": torch.nonzero(torch.isnan(back[0, 0])).tolist(),
        }
        s = stride
        ps = max(1, P // 16)
        save(f"resample_{tag}",
             nrows=nrows, erp=np.array(erp), P=P, stride=s, pstride=ps, seed=seed,
             pers=pers[:, :, ::ps, ::ps, :], xyz=xyz[:, :, ::ps, ::ps], uv=uv[:, :, ::ps, ::ps],
             center_p=center_p,
             pers_sum=pers.double().sum(), back=back[:, :, ::s, ::s], back_sum=torch.nan_to_num(back).double().sum(),
             back_nan=torch.nonzero(torch.isnan(back[0, 0])),
             e2p_x0=x0[:, ::ps, ::ps].to(torch.int16), e2p_y0=y0[:, ::ps, ::ps].to(torch.int16),
             e2p_x0_sum=x0.sum(), e2p_y0_sum=y0.sum(),
             t_x0=tab["x0"][:, ::s, ::s].to(torch.uint8), t_y0=tab["y0"][:, ::s, ::s].to(torch.uint8),
             t_x1=tab["x1"][:, ::s, ::s].to(torch.uint8), t_y1=tab["y1"][:, ::s, ::s].to(torch.uint8),
             t_mask=tab["mask"][:, ::s, ::s].to(torch.uint8), t_w=tab["w_list"][:, ::s, ::s],
             t_sums=np.array([int((tab[k] * tab["mask"]).sum()) for k in ("x0", "y0", "x1", "y1")]
                             + [int(tab["mask"].sum())], dtype=np.int64),
             t_w_sum=torch.nan_to_num(tab["w_list"]).double().sum(),
             t_w_nan=torch.nonzero(torch.isnan(tab["w_list"]).any(-1)))
    in_fresh_cwd(run)


# --------------------------------------------------------------------- models
def hooks(model, names):
    store, hs = {}, []
    for n in names:
        m = dict(model.named_modules())[n]
        hs.append(m.register_forward_hook(
            lambda mod, inp, out, n=n: store.setdefault(n, []).append(out.detach().clone())))
    return store, hs


def model_case(tag, kind, nrows, erp, bs, iters, conf, stride, seed=123):
    def run(d):
        npatch = NP[nrows]
        sd = synthetic_state_dict(kind, npatch, 0)
        Ref = RefIter if kind == "iterative" else RefSingle
        net = Ref(nrows, npatch, (128, 128), FOV).eval()
        net.load_state_dict(sd)
        rgb = rand((bs, 3, *erp), seed)
        probe = ["layer1", "layer2", "layer3", "layer4", "transformer", "de_conv4_0", "pred", "weight_pred"]
        store, hs = hooks(net, probe)
        with torch.no_grad():
            outs = net(rgb, iter=iters, confidence=conf) if kind == "iterative" else [net(rgb, confidence=conf)]
        for h in hs:
            h.remove()
        tr = {}
        if kind == "iterative":
            o_outs = om.forward_iterative(sd, rgb, iters, conf, nrows=nrows, fov=FOV, trace=tr)
        else:
            t0 = {}
            o_outs = [om.forward_single(sd, rgb, conf, nrows=nrows, fov=FOV, trace=t0)]
            tr["iter0"] = t0
        # NaN-aware: the reference itself returns NaN at the ERP pixels whose blend weights are NaN (see eq())
        same_nan = all(torch.equal(torch.isnan(a), torch.isnan(b)) for a, b in zip(outs, o_outs))
        fin = lambda t: t[~torch.isnan(t)]
        rel = max((((fin(a) - fin(b)).abs() / fin(a).abs().clamp_min(1e-6)).max().item()) for a, b in zip(outs, o_outs)) \
            if same_nan else float("inf")
        meta["checks"][tag] = {"oracle_vs_reference_max_rel": rel,
                               "reference_nan_pixels": torch.nonzero(torch.isnan(outs[-1][0, 0])).tolist(),
                               "depth_min": min(fin(o).min().item() for o in outs),
                               "depth_max": max(fin(o).max().item() for o in outs),
                               "depth_std": fin(outs[-1]).std().item()}
        arrs = {"nrows": nrows, "erp": np.array(erp), "bs": bs, "iters": iters, "conf": int(conf),
                "stride": stride, "seed": seed, "kind": kind}
        for i, o in enumerate(outs):
            arrs[f"out{i}"] = o[:, :, ::stride, ::stride]
            arrs[f"out{i}_mean"] = fin(o).double().mean()           # over the finite pixels
            arrs[f"out{i}_nan"] = torch.nonzero(torch.isnan(o))
        # probes: reference layout is (B,C,H,W,N) for conv maps, (B,N,512) for the transformer
        omap = {"layer1": "layer1_pre", "layer2": "layer2", "layer3": "layer3", "layer4": "layer4",
                "transformer": "encoded", "de_conv4_0": "de_conv4_0", "pred": "pred_raw",
                "weight_pred": "weight_raw"}
        worst = 0.0
        for n in probe:
            if n == "weight_pred" and not conf:
                continue
            for it, t in enumerate(store[n]):
                o = tr[f"iter{it}"][omap[n]]
                if t.dim() == 5:
                    t2 = om._fold(t)
                    sl = t2[::7, ::5, ::9, ::9]
                else:
                    t2 = t
                    sl = t2[:, ::3, ::17]
                worst = max(worst, ((t2 - o).abs().max() / t2.abs().max()).item())
                arrs[f"probe_{n}_{it}"] = sl
                arrs[f"probe_{n}_{it}_absmean"] = t2.double().abs().mean()
        meta["checks"][tag]["oracle_vs_reference_probe_max_rel_to_absmax"] = worst
        save(f"model_{tag}", **arrs)
    in_fresh_cwd(run)


def variant_case(tag, which, erp, bs, seed=123):
    """network_360d.py / network_test.py through the real reference classes."""
    def run(d):
        import network_360d
        import network_test
        rgb = rand((bs, 3, *erp), seed)
        if which == "360d":
            sd = synthetic_state_dict("iterative", 18, 0)
            net = network_360d.spherical_fusion().eval()
            net.load_state_dict(sd)
            with torch.no_grad():
                outs = [net(rgb, FOV, (128, 128), 4)]
            o_outs = [om.forward_360d(sd, rgb, FOV, (128, 128), 4)]
        else:
            sd = synthetic_state_dict("test", 18, 0)
            net = network_test.spherical_fusion().eval()
            net.load_state_dict(sd)
            with torch.no_grad():
                outs = net(rgb, FOV, (256, 256), 4, 2)
            o_tr = {}
            o_erp = om.forward_iterative(sd, rgb, 2, False, nrows=4, fov=FOV, patch_size=(256, 256), trace=o_tr)
            # the reference appends the PATCH prediction relu(pred) (B,1,P,P,N) for iterations >= 1 (network_test.py:441-445)
            o_outs = [o_erp[0], om._unfold(torch.relu(o_tr["iter1"]["pred_raw"]), bs)]
        rel = max((((a - b).abs() / a.abs().clamp_min(1e-6)).max().item()) for a, b in zip(outs, o_outs))
        meta["checks"][tag] = {"oracle_vs_reference_max_rel": rel, "depth_min": min(o.min().item() for o in outs),
                               "depth_max": max(o.max().item() for o in outs), "depth_std": outs[-1].std().item()}
        arrs = {"which": which, "erp": np.array(erp), "bs": bs, "seed": seed}
        for i, o in enumerate(outs):
            arrs[f"out{i}"] = o if o.dim() == 4 else o[:, :, ::8, ::8, :]      # patch-space output: strided
            arrs[f"out{i}_mean"] = o.double().mean()
        save(f"variant_{tag}", **arrs)
    in_fresh_cwd(run)


CASES = {
    "net360d_small": lambda: variant_case("net360d_small", "360d", (64, 128), 2),
    "nettest_p256_small": lambda: variant_case("nettest_p256_small", "test", (64, 128), 1),
    **{f"small_n{n}": (lambda n=n: resampler_case(f"small_n{n}", n, (16, 32), 16, 1, 10 + n)) for n in (3, 4, 5, 6)},
    "mid_n4_p32": lambda: resampler_case("mid_n4_p32", 4, (128, 256), 32, 4, 21),
    "full_n4": lambda: resampler_case("full_n4", 4, (512, 1024), 128, 16, 22),
    "full_n6": lambda: resampler_case("full_n6", 6, (512, 1024), 128, 16, 23),
    # BASELINE configs[3] geometry (1024x2048, nrows=5): the reference's dense table is 3 GB here
    "full_n5_2k": lambda: resampler_case("full_n5_2k", 5, (1024, 2048), 128, 32, 24),
    "iter_small_conf0": lambda: model_case("iter_small_conf0", "iterative", 4, (64, 128), 2, 2, False, 1),
    "iter_small_conf1": lambda: model_case("iter_small_conf1", "iterative", 4, (64, 128), 2, 2, True, 1),
    "single_small_conf1": lambda: model_case("single_small_conf1", "single", 4, (64, 128), 2, 1, True, 1),
    "single_small_conf0": lambda: model_case("single_small_conf0", "single", 4, (64, 128), 1, 1, False, 1),
    "iter_n6_conf1": lambda: model_case("iter_n6_conf1", "iterative", 6, (64, 128), 1, 2, True, 1),
    "iter_n5_conf1": lambda: model_case("iter_n5_conf1", "iterative", 5, (64, 128), 1, 2, True, 1),
    "iter_full_conf1": lambda: model_case("iter_full_conf1", "iterative", 4, (512, 1024), 1, 2, True, 8),
    # BASELINE configs[3] / [4] at their real geometry (one panorama each)
    "iter_full_n5": lambda: model_case("iter_full_n5", "iterative", 5, (1024, 2048), 1, 2, True, 16),
    "iter_full_n6": lambda: model_case("iter_full_n6", "iterative", 6, (512, 1024), 1, 2, True, 8),
}


if __name__ == "__main__":
    # python make_golden.py [case ...]   (no arguments: every case); meta.json is merged, not replaced
    torch.manual_seed(0)
    todo = sys.argv[1:] or list(CASES)
    mpath = os.path.join(HERE, "meta.json")
    if os.path.exists(mpath):
        old = json.load(open(mpath))
        meta["checks"].update(old.get("checks", {}))
    for name in todo:
        CASES[name]()
    with open(mpath, "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print(json.dumps({k: meta["checks"][k] for k in todo if k in meta["checks"]}, indent=1, sort_keys=True))
