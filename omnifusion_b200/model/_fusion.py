"""Host side of the OmniFusion patch network: an nn.Module that carries the reference's
parameters (same state_dict keys and shapes) and runs the forward through libofb's engine.

Mirrors /root/reference/model/spherical_model_iterative.py:253-456 and
model/spherical_model.py:190-314.  Inference only (BatchNorm uses running statistics, as
under ``.eval()`` in the reference's test.py:192).
"""
import ctypes as C
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import _lib, tables
from ..checkpoint import key_spec, strip_module_prefix, synthetic_state_dict

_BUFFER_LEAVES = ("running_mean", "running_var", "num_batches_tracked")


class _Holder(nn.Module):
    """Parameter container; the arithmetic happens in libofb."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("omnifusion_b200 sub-modules only hold parameters; call the model's forward")


class SphericalFusionBase(nn.Module):
    KIND = "iterative"
    PATCH_SIZES = ((128, 128),)      # what the token width allows: 32 * (P / 32)^2 == 512

    def __init__(self, nrows=4, npatches=18, patch_size=(128, 128), fov=(80, 80)):
        super().__init__()
        self.nrows = nrows
        self.npatches = npatches
        self.patch_size = tables.pair(patch_size)
        self.fov = tables.pair(fov)
        if self.patch_size not in self.PATCH_SIZES:
            # the reference hard-wires the token width to 32*(P/32)^2 == 512
            # (spherical_model_iterative.py:274,276,331): any other patch size shape-errors there too
            raise ValueError(f"patch_size must be one of {self.PATCH_SIZES}, got {self.patch_size}")
        if tables.NUM_PATCHES.get(nrows) != npatches:
            raise ValueError(f"nrows={nrows} has {tables.NUM_PATCHES.get(nrows)} patches, npatches={npatches}")
        init = synthetic_state_dict(self.KIND, npatches, seed=0)
        for name, shape in key_spec(self.KIND, npatches).items():
            *path, leaf = name.split(".")
            mod = self
            for part in path:
                if part not in mod._modules:
                    mod.add_module(part, _Holder())
                mod = mod._modules[part]
            value = init[name].clone()
            assert tuple(value.shape) == tuple(shape)
            if leaf in _BUFFER_LEAVES:
                mod.register_buffer(leaf, value)
            else:
                mod.register_parameter(leaf, nn.Parameter(value, requires_grad=False))
        self.eval()
        self._handle = None
        self._handle_device = None
        self._weights_key = None
        self._geometry_key = None
        self._keepalive = None
        self._options = {}
        self._graphs = {}
        self._tensor_list = None       # cached list of checkpoint tensors (rebuilt after load_state_dict / .to())

    # ------------------------------------------------------------------ plumbing
    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts reference checkpoints, including DataParallel's ``module.`` prefix (test.py:107-110)."""
        self._tensor_list = None
        return super().load_state_dict(strip_module_prefix(OrderedDict(state_dict)), strict=strict, **kw)

    def _apply(self, fn, *a, **k):
        self._tensor_list = None        # .to() / .cuda() / .float() may replace parameter storage
        return super()._apply(fn, *a, **k)

    def set_option(self, key, value):
        """Engine knobs: 'engine' (0 auto / 1 CUDA-core / 2 tcgen05), 'chunk', 'dedup'."""
        self._options[key] = int(value)
        if self._handle is not None:
            _lib.check(_lib.lib().ofb_set_option(self._handle, key.encode(), int(value)))
        self._graphs.clear()

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None:
                _lib.lib().ofb_destroy(self._handle)
        except Exception:
            pass

    def _ensure_handle(self, device):
        if self._handle is not None and self._handle_device == device:
            return
        if self._handle is not None:
            _lib.lib().ofb_destroy(self._handle)
        h = C.c_void_p()
        idx = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(_lib.lib().ofb_create(idx, C.byref(h)))
        self._handle, self._handle_device = h, device
        self._weights_key = self._geometry_key = None
        self._graphs.clear()            # graphs captured with the old handle replay launches into freed memory
        for k, v in self._options.items():
            _lib.check(_lib.lib().ofb_set_option(h, k.encode(), v))

    def _ensure_weights(self):
        """Re-packs the checkpoint on the device when any tensor changed.  The change check is one
        (data_ptr, version) pair per tensor over a cached tensor list - no state_dict() rebuild per forward."""
        if self._tensor_list is None:
            # storage can only move through load_state_dict / _apply (.to, .cuda, .float ...), which reset this list;
            # in between, an in-place edit shows up in the tensors' version counters - one attribute read each
            self._tensor_list = list(self.state_dict().items())
            self._weights_key = None
        key = tuple(v._version for _, v in self._tensor_list)
        if key == self._weights_key:
            return
        tensors = OrderedDict(self._tensor_list)
        host = [(k, v.detach().to("cpu", torch.float32).contiguous()) for k, v in tensors.items()
                if not k.endswith("num_batches_tracked")]
        descs = (_lib.TensorDesc * len(host))()
        for d, (k, v) in zip(descs, host):
            d.name = k.encode()
            d.data = v.data_ptr()
            d.ndim = v.dim()
            for i, s in enumerate(v.shape):
                d.shape[i] = s
        _lib.check(_lib.lib().ofb_load_weights(self._handle, descs, len(host), int(self.KIND == "single")))
        self._weights_key = key
        self._graphs.clear()

    def _point_table(self, low_geo, device):
        return low_geo["xyz"]

    def _ensure_geometry(self, erp_hw, device):
        key = (erp_hw, str(device), self.fov, self.nrows, self.patch_size)
        if key == self._geometry_key:
            return
        p = self.patch_size[0]
        hi = tables.device_patch_geometry(self.fov, self.nrows, (p, p), device)
        lo = tables.device_patch_geometry(self.fov, self.nrows, (p // 4, p // 4), device)
        blend = tables.device_blend_table(self.fov, self.nrows, (p, p), erp_hw, device)
        pts = self._point_table(lo, device).contiguous()
        g = _lib.Geometry()
        g.n_patch, g.patch, g.erp_h, g.erp_w = self.npatches, p, erp_hw[0], erp_hw[1]
        g.grid_hi, g.grid_lo = hi["grid"].data_ptr(), lo["grid"].data_ptr()
        g.pts, g.pts_c = pts.data_ptr(), pts.shape[1]
        g.blend_rowptr, g.blend_idx, g.blend_w = (blend["rowptr"].data_ptr(), blend["idx"].data_ptr(),
                                                  blend["w"].data_ptr())
        _lib.check(_lib.lib().ofb_set_geometry(self._handle, C.byref(g)))
        self._keepalive = (hi, lo, blend, pts)      # the engine keeps raw pointers into these
        self._geometry_key = key
        self._graphs.clear()

    def _run(self, rgb, iters, confidence):
        if self.training:
            raise RuntimeError("omnifusion_b200 implements inference only: call .eval() first")
        rgb = _lib.require_cuda(rgb, "input")
        if rgb.dim() != 4 or rgb.shape[1] != 3:
            raise ValueError(f"input must be (B,3,He,We), got {tuple(rgb.shape)}")
        if iters < 1:
            raise ValueError("iter must be >= 1")
        bs, _, he, we = rgb.shape
        self._ensure_handle(rgb.device)
        self._ensure_weights()
        self._ensure_geometry((he, we), rgb.device)
        outs = [torch.empty((bs, 1, he, we), dtype=torch.float32, device=rgb.device) for _ in range(iters)]
        arr = (C.c_void_p * iters)(*[o.data_ptr() for o in outs])
        _lib.check(_lib.lib().ofb_forward_f32(self._handle, _lib.ptr(rgb), bs, iters, int(bool(confidence)),
                                              arr, _lib.stream_of(rgb.device)))
        return outs

    def forward_graphed(self, rgb, iters=1, confidence=False):
        """Same as forward, replayed from a CUDA graph captured for this (shape, iters, confidence).
        The returned tensors are the graph's static outputs and are overwritten by the next call."""
        if self.training:
            raise RuntimeError("omnifusion_b200 implements inference only: call .eval() first")
        key = (tuple(rgb.shape), str(rgb.device), iters, bool(confidence))
        if self._handle is not None and self._handle_device == rgb.device:
            # anything that invalidates captured launches drops the graphs: new weights (load_state_dict,
            # in-place edits), and a workspace arena re-allocated by a larger eager forward in between
            self._ensure_weights()
            gen = _lib.lib().ofb_workspace_generation(self._handle)
            if any(e[3] != gen for e in self._graphs.values()):
                self._graphs.clear()
        ent = self._graphs.get(key)
        if ent is None:
            static_in = rgb.clone()
            side = torch.cuda.Stream(device=rgb.device)
            side.wait_stream(torch.cuda.current_stream(rgb.device))
            with torch.cuda.stream(side):
                for _ in range(2):                       # warm-up: workspace + tables allocated
                    self._run(static_in, iters, confidence)
            torch.cuda.current_stream(rgb.device).wait_stream(side)
            gen = _lib.lib().ofb_workspace_generation(self._handle)
            if any(e[3] != gen for e in self._graphs.values()):
                self._graphs.clear()                     # the warm-up grew the arena under older graphs
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self._run(static_in, iters, confidence)
            ent = (graph, static_in, static_out, gen)
            self._graphs[key] = ent
        graph, static_in, static_out, _ = ent
        static_in.copy_(rgb, non_blocking=True)
        graph.replay()
        return static_out

    def range_report(self):
        """With set_option("check_range", 1): (number of activations outside the range the storage format represents,
        {name: (max_abs, nonfinite_count)}) of the last forward.  Synchronises.  The split-half format overflows above
        65504 and loses low-order bits below ~6e-5 (include/ofb.h); the reference is fp32."""
        buf = C.create_string_buffer(4096)
        bad = _lib.check(_lib.lib().ofb_range_report(self._handle, buf, len(buf)))
        rows = {}
        for ln in buf.value.decode().strip().split("\n"):
            f = ln.split()
            if len(f) == 3:
                rows[f[0]] = (float(f[1]), int(f[2]))
        return bad, rows

    def check_numerics(self):
        """Raises OfbError when the last forward's activations left the representable range (needs check_range = 1)."""
        bad, rows = self.range_report()
        if bad:
            worst = max(rows.items(), key=lambda kv: (kv[1][1], kv[1][0]))
            raise _lib.OfbError(f"{bad} activation tensor(s) left the range of the split-half storage format "
                                f"(e.g. {worst[0]}: max |x| = {worst[1][0]:.3e}, {worst[1][1]} non-finite); "
                                "run this checkpoint with set_option('format', 0)")
        return rows

    def activation(self, name):
        """Debug/test hook: a named intermediate of the last forward as a (n,h,w,c) tensor."""
        dims = (C.c_int * 4)()
        n = _lib.check(_lib.lib().ofb_get_activation(self._handle, name.encode(), None, 0, C.byref(dims), None))
        out = torch.empty(tuple(dims), dtype=torch.float32, device=self._handle_device)
        _lib.check(_lib.lib().ofb_get_activation(self._handle, name.encode(), _lib.ptr(out), n, C.byref(dims),
                                                 _lib.stream_of(self._handle_device)))
        return out
