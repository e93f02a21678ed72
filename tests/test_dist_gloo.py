"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the frame sharding and the two calibration
collectives (MAX over [-min, max], SUM over the int64 histograms) of sesr_pytorch_quantize_b200.calib.
On the GPU box the same code runs over NCCL (bench.py --gpus N, tests/test_multi_gpu.py + tests/multi_gpu_worker.py)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sesr_pytorch_quantize_b200.calib import merge_minmax, merge_histograms, shard_range
    rng = np.random.default_rng(42)
    frames_mm = rng.standard_normal((7, 6, 2)).astype(np.float32)       # 7 "frames" of per-domain (a, b)
    frames_mm.sort(axis=2)
    frames_h = rng.integers(0, 1000, size=(7, 6, 2048)).astype(np.int64)
    lo, hi = shard_range(7, rank, world)
    mm = torch.empty(6, 2)
    mm[:, 0], mm[:, 1] = float("inf"), float("-inf")
    h = torch.zeros(6, 2048, dtype=torch.int64)
    for f in range(lo, hi):
        mm[:, 0] = torch.minimum(mm[:, 0], torch.from_numpy(frames_mm[f, :, 0]))
        mm[:, 1] = torch.maximum(mm[:, 1], torch.from_numpy(frames_mm[f, :, 1]))
        h += torch.from_numpy(frames_h[f])
    mm_all = merge_minmax(mm)
    h_all = merge_histograms(h)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), mm=mm_all.numpy(), h=h_all.numpy(), lo=lo, hi=hi,
             want_mm=np.stack([frames_mm[:, :, 0].min(0), frames_mm[:, :, 1].max(0)], 1), want_h=frames_h.sum(0))
    dist.destroy_process_group()


def test_world2_merge_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    spans = []
    for r in range(world):
        z = np.load(tmp_path / ("r%d.npz" % r))
        assert np.array_equal(z["mm"], z["want_mm"])          # bit-identical to the single-process statistics
        assert np.array_equal(z["h"], z["want_h"])
        spans.append((int(z["lo"]), int(z["hi"])))
    assert spans == [(0, 4), (4, 7)]


def test_shard_range_partitions_every_count():
    from sesr_pytorch_quantize_b200.calib import shard_range
    for n in (0, 1, 7, 8, 64, 513):
        for world in (1, 2, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_merge_is_identity_without_process_group():
    from sesr_pytorch_quantize_b200.calib import merge_minmax, merge_histograms
    mm = torch.rand(6, 2)
    h = torch.randint(0, 9, (6, 2048))
    assert merge_minmax(mm) is mm and merge_histograms(h) is h
