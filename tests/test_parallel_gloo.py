"""Multi-process host logic on CPU: world_size 2, gloo backend (the data path itself has no
collective, so what is covered is the frame sharding, the single flat gradient all-reduce of the training step and the
launch contract of bench.py's reference arm under torchrun)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_covers_every_frame_once():
    from fovea.parallel import shard_range
    for n in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
    from types import SimpleNamespace as NS
    from fovea.models import CompressNet
    from fovea.parallel import FlatGradBucket, shard_batch
    from fovea.saliency_network import fov_simple
    cfg = NS(MODEL=NS(saliency_net="fovsimple", fov_deform=True))
    torch.manual_seed(0)                                   # identical weights on every rank (what DDP broadcasts)
    sal, comp = fov_simple(cfg), CompressNet(cfg)
    feed = {"x": torch.randn(6, 5, 16, 16, generator=torch.Generator().manual_seed(1)),
            "t": torch.randn(6, 1, 16, 16, generator=torch.Generator().manual_seed(2))}
    mine = shard_batch(feed, rank, world)
    assert mine["x"].shape[0] == 3
    loss = ((comp(sal(mine["x"])) - mine["t"]) ** 2).mean()
    loss.backward()
    bucket = FlatGradBucket([sal, comp])
    assert bucket.numel == sum(p.numel() for m in (sal, comp) for p in m.parameters())
    pend = bucket.allreduce(async_op=True)
    pend.wait()
    grads = torch.cat([p.grad.flatten() for p in bucket.params])
    torch.save(grads, os.path.join(out_dir, f"g{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_bucket_allreduce_world2(tmp_path):
    """The bucket's mean-all-reduce equals the gradient of the full-batch mean loss (equal shard sizes)."""
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0, g1 = torch.load(tmp_path / "g0.pt"), torch.load(tmp_path / "g1.pt")
    assert torch.equal(g0, g1)                              # every rank ends with the same gradient
    sys.path[:0] = [ROOT, os.path.join(ROOT, "foveated-instance-segmentation_b200")]
    from types import SimpleNamespace as NS
    from fovea.models import CompressNet
    from fovea.saliency_network import fov_simple
    cfg = NS(MODEL=NS(saliency_net="fovsimple", fov_deform=True))
    torch.manual_seed(0)
    sal, comp = fov_simple(cfg), CompressNet(cfg)
    sal.eval(), comp.eval()                                # BatchNorm statistics are per-rank in the reference too;
    x = torch.randn(6, 5, 16, 16, generator=torch.Generator().manual_seed(1))
    t = torch.randn(6, 1, 16, 16, generator=torch.Generator().manual_seed(2))
    # per-shard losses averaged == what the two ranks computed (train-mode BN uses per-rank batch statistics, exactly
    # like the reference's un-synchronised "SynchronizedBatchNorm" under DDP, SURVEY.md surprise 5)
    sal.train(), comp.train()
    tot = None
    for lo in (0, 3):
        for p in list(sal.parameters()) + list(comp.parameters()):
            p.grad = None
        loss = ((comp(sal(x[lo:lo + 3])) - t[lo:lo + 3]) ** 2).mean()
        loss.backward()
        g = torch.cat([p.grad.flatten() for m in (sal, comp) for p in m.parameters()])
        tot = g if tot is None else tot + g
    want = tot / 2
    assert torch.allclose(g0, want, rtol=1e-5, atol=1e-7)


def test_bench_reference_arm_under_torchrun_world2():
    """`bench.py --impl reference` launched as the driver launches it for N=2: rank 0 alone works and prints ONE JSON
    line with impl=reference; rank 1 exits 0 without output."""
    port = _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--impl", "reference", "--workload", "tiny"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
