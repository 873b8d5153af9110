"""N>1 host logic on CPU: two gloo ranks shard a batch of panoramas, compute partial metric
sums on their shards and all-reduce them; the result must equal the unsharded metric
(mirrors AverageMeter.update(val, N) of the reference's test.py:171-177)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from omnifusion_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = parallel.init_from_env("gloo")
    g = torch.Generator().manual_seed(5)
    pred = 0.1 + 7.9 * torch.rand(total, 1, 16, 32, generator=g)
    gt = 0.1 + 7.9 * torch.rand(total, 1, 16, 32, generator=g)
    mask = (gt <= 8) & (gt > 0.1) & (torch.rand(total, 1, 16, 32, generator=g) > 0.3)
    lo, hi = parallel.shard_bounds(total, r, w)
    p, t, m = (parallel.shard_batch(x, r, w) for x in (pred, gt, mask))
    assert p.shape[0] == hi - lo
    part = torch.tensor([((p[m] - t[m]).abs() / t[m]).double().sum(), m.sum().double()], dtype=torch.float64)
    parallel.reduce_sums(part)
    slow = parallel.max_over_ranks(10.0 + r, torch.device("cpu"))
    parallel.barrier()
    if r == 0:
        whole = ((pred[mask] - gt[mask]).abs() / gt[mask]).double()
        q.put((float(part[0] / part[1]), float(whole.mean()), int(part[1]), int(mask.sum()), slow))
    dist.destroy_process_group()


def test_two_rank_sharded_metric_equals_unsharded():
    world, total = 2, 7              # ragged: 4 + 3 panoramas
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sharded, whole, n_sharded, n_whole, slow = got
    assert n_sharded == n_whole
    assert abs(sharded - whole) < 1e-12
    assert slow == 11.0               # max over ranks


def test_shard_bounds_cover_batch_exactly():
    for total in (1, 7, 8, 64, 129):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _cpu_partial(pred, gt, mask, scale):
    """The nine sums ofb_depth_metrics_partial produces (resample.cu: depth_metrics_kernel), with torch on the CPU."""
    m = mask > 0
    p, g = (pred * scale)[m].double(), gt[m].double()
    d = p - g
    ok = (p > 1e-7) & (g > 1e-7)
    r = torch.maximum(p / g, g / p)
    return torch.stack([(d.abs() / g).sum(), (d * d / g).sum(), (d * d).sum(),
                        ((p[ok].log() - g[ok].log()) ** 2).sum(), ok.sum().double(),
                        (r < 1.25).sum().double(), (r < 1.25 ** 2).sum().double(), (r < 1.25 ** 3).sum().double(),
                        m.sum().double()])


def _meter_worker(rank, world, port, total, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from omnifusion_b200 import metrics
    from oracle import model as om
    r, w, _ = parallel.init_from_env("gloo")
    g = torch.Generator().manual_seed(9)
    pred = 0.1 + 7.9 * torch.rand(total, 1, 16, 32, generator=g)
    pred[:, :, 0, :5] = 0.0                              # pred <= 1e-7: leaves rms_sq_log's own valid count below N
    gt = 0.1 + 7.9 * torch.rand(total, 1, 16, 32, generator=g)
    mask = (gt <= 8) & (gt > 0.1) & (torch.rand(total, 1, 16, 32, generator=g) > 0.3)
    p, t, m = (parallel.shard_batch(x, r, w) for x in (pred, gt, mask))
    meters = metrics.DepthMeters(torch.device("cpu"))    # the product's meter algebra + collective; sums from torch-CPU
    scale = (t[m].median() / p[m].median()).item()       # per-shard batch tensor, as documented in DepthMeters
    meters.acc += metrics.meter_update(_cpu_partial(p, t, m, scale))
    got = meters.all_reduce().result()
    if r == 0:
        # the reference: one AverageMeter.update(val, N) per (shard) batch - test.py:171-177
        want = {k: 0.0 for k in metrics.METRIC_NAMES}
        n_all = 0
        for rr in range(w):
            ps, ts, ms = (parallel.shard_batch(x, rr, w) for x in (pred, gt, mask))
            ref = om.eval_metrics(ps, ts, ms, median_scale=True)
            for k in want:
                want[k] += ref[k] * ref["n"]
            n_all += ref["n"]
        q.put((got, {k: v / n_all for k, v in want.items()}, n_all))
    dist.destroy_process_group()


def test_two_rank_depth_meters_follow_average_meter_semantics():
    """metrics.DepthMeters (meter_update, all_reduce -> parallel.reduce_sums over gloo, result) against the
    reference's N-weighted AverageMeters, incl. rms_sq_log whose per-batch value is a mean over its OWN valid pixels."""
    world, total = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_meter_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, want, n_all = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got["n"] == n_all
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-5 * max(1.0, abs(v)), (k, got[k], v)
