"""Host-side logic of the data-parallel path on CPU: two gloo ranks feed different gradients through
GradReducer in backward order; every rank must end with the cross-rank mean, bucket by bucket."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mdeical_image_segmentation_b200.parallel import DataParallelUNet, GradReducer, allreduce_mean_
    torch.manual_seed(100 + rank)                      # different init per rank: broadcast must fix it
    net = nn.Sequential(nn.Conv2d(3, 8, 3), nn.Conv2d(8, 8, 3), nn.Conv2d(8, 2, 1))
    net.final_conv = net[2]
    dp = DataParallelUNet(net, bucket_mb=0.001)        # ~260 floats per bucket -> several buckets
    ref0 = [p.detach().clone() for p in net.parameters()]
    red: GradReducer = dp.reducer
    params = list(net.parameters())
    g = torch.Generator().manual_seed(7 + rank)
    grads = {p: torch.randn(p.shape, generator=g) for p in params}
    out = {}
    for p in reversed(params):                         # backward order: head first
        out[p] = red.on_grad(p, grads[p])
    slots = red.finish()
    res = [slots[p].clone() for p in params]
    loss = allreduce_mean_(torch.tensor([float(rank)]))
    q.put((rank, [t.numpy() for t in ref0], [g_.numpy() for g_ in grads.values()],
           [t.numpy() for t in res], red.launched_buckets, loss.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, w0, g0, r0, nb0, l0), (_, w1, g1, r1, nb1, l1) = got
    for a, b in zip(w0, w1):                           # parameters broadcast from rank 0
        assert (a == b).all()
    assert nb0 == nb1 and nb0 >= 2                     # more than one bucket was exercised
    for ga, gb, ra, rb in zip(g0, g1, r0, r1):
        mean = (ga + gb) / 2
        assert abs(ra - mean).max() < 1e-6 and abs(rb - mean).max() < 1e-6
    assert l0 == l1 == 0.5


def test_reducer_is_attached_to_every_executor():
    """DataParallelUNet must hand its GradReducer to every whole-network executor it wraps — the U-Nets
    (``final_conv``) and the UNet 3+ family, also behind the HF-style ``UNetModel`` wrapper — otherwise the
    backward pass would silently skip the gradient all-reduce."""
    import mdeical_image_segmentation_b200 as B
    from mdeical_image_segmentation_b200.parallel import DataParallelUNet
    for make, inner in ((lambda: B.UNet3D(1, 1, f_maps=[8, 16]), lambda m: m),
                        (lambda: B.UNetModel(B.UNetConfig(1, 1)), lambda m: m.unet),
                        (lambda: B.UNetModel(B.UNetConfig(3, 1, "UNet_3Plus")), lambda m: m.unet),
                        (lambda: B.UNet_3Plus_DeepSup(3, 1), lambda m: m)):
        model = make()
        dp = DataParallelUNet(model)
        assert inner(model)._b200_reducer is dp.reducer, type(inner(model)).__name__
        assert sum(p.numel() for p in dp.reducer.params) == sum(p.numel() for p in model.parameters())


def test_reducer_results_do_not_alias_bucket_memory():
    """what backward returns becomes ``p.grad`` without a clone, so the reducer must hand out fresh
    memory — a second backward (gradient accumulation, ``zero_grad(set_to_none=False)``) must not overwrite
    the first result in place."""
    from mdeical_image_segmentation_b200.parallel import GradReducer
    net = nn.Sequential(nn.Linear(4, 3), nn.Linear(3, 2))
    params = list(net.parameters())
    red = GradReducer(params, bucket_mb=0.00001)
    g1 = {p: torch.full(p.shape, 1.0) for p in params}
    g2 = {p: torch.full(p.shape, 5.0) for p in params}
    for p in reversed(params):
        red.on_grad(p, g1[p])
    first = red.finish()
    keep = {p: first[p].clone() for p in params}
    for p in reversed(params):
        red.on_grad(p, g2[p])
    second = red.finish()
    for p in params:
        assert first[p].data_ptr() != second[p].data_ptr()
        assert torch.equal(first[p], keep[p]) and torch.equal(first[p], g1[p])
        assert torch.equal(second[p], g2[p])
        assert first[p].shape == p.shape
