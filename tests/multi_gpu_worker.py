"""One rank of the multi-GPU parity check (launched by tests/test_multi_gpu.py through torch.distributed.run).

    (i)  inference: ONE global seeded batch of the BASELINE config-3 shape (SESR x4, 540x960 -> 2160x3840) is sharded with
         calib.shard_range; every rank runs its contiguous shard; the per-frame sha256 of q5 and y, gathered over the
         ranks, must equal those of the same batch run on one GPU, and one frame per rank must equal the C oracle.
    (ii) calibration: MinMaxCalibrator / KLCalibrator over ONE global calibration set (seed 1234, contiguous shards) with
         the real collectives must give exactly the thresholds / KL indices a single process gets from all frames.
Backend: NCCL with one GPU per rank when the box has enough GPUs; otherwise the ranks share cuda:0 and the collectives
run over gloo (same kernels, same merge code).  Rank 0 prints one JSON line.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def sha(t):
    return hashlib.sha256(t.contiguous().cpu().numpy().tobytes()).hexdigest()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ngpu = torch.cuda.device_count()
    backend = "nccl" if ngpu >= world else "gloo"
    dev = torch.device("cuda", local if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    from sesr_pytorch_quantize_b200 import Int8Net, compile_network
    from sesr_pytorch_quantize_b200.calib import KLCalibrator, MinMaxCalibrator, shard_range
    from sesr_pytorch_quantize_b200.compile import thresholds_from_minmax
    from sesr_pytorch_quantize_b200.fixtures import load_golden
    from sesr_pytorch_quantize_b200.fq import FakeQuantNet
    res = {"world": world, "backend": backend, "gpus": ngpu}

    # ---- (i) sharded inference == single GPU == oracle -----------------------------------------------------------
    g = load_golden("x4_fixture")
    B, H, W = int(os.environ.get("SESRQ_MG_FRAMES", "8")), 540, 960
    x_all = torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(4321))
    net = Int8Net(compile_network([torch.from_numpy(w) for w in g.wf], [torch.from_numpy(b) for b in g.bf], g.scales, g.zeros, g.r),
                  device=dev)
    lo, hi = shard_range(B, rank, world)
    q5, y = net(x_all[lo:hi].to(dev))
    torch.cuda.synchronize(dev)
    mine = [(lo + i, sha(q5[i]), sha(y[i])) for i in range(hi - lo)]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # one frame of this rank's shard against the C restatement of the reference
    import c_oracle
    import sesr_oracle as orc
    cn = c_oracle.CNet(orc.derive_net(g.wf, g.bf, g.scales, g.zeros, g.r))
    ok_oracle = True
    if hi > lo:
        q_ref, y_ref = cn.forward(x_all[lo:lo + 1].numpy())
        ok_oracle = bool(np.array_equal(q5[:1].cpu().numpy(), q_ref) and np.array_equal(y[:1].cpu().numpy(), y_ref))
    oks = [None] * world
    dist.all_gather_object(oks, ok_oracle)
    if rank == 0:
        q1, y1 = net(x_all.to(dev))
        single = [(i, sha(q1[i]), sha(y1[i])) for i in range(B)]
        sharded = sorted(t for part in gathered for t in part)
        res["frames"] = B
        res["sharded_equals_single"] = sharded == single
        res["covered_every_frame_once"] = [t[0] for t in sharded] == list(range(B))
        res["oracle_ok_per_rank"] = oks
        del q1, y1
    del q5, y
    net.close()

    # ---- (ii) calibration: N-rank thresholds == 1-rank thresholds ---------------------------------------------------
    gx = load_golden("x2_fixture")
    n_cal = int(os.environ.get("SESRQ_MG_CALIB", "16"))
    frames_all = torch.rand(n_cal, 3, 80, 960, generator=torch.Generator().manual_seed(1234))
    lo, hi = shard_range(n_cal, rank, world)

    def make():
        return FakeQuantNet([torch.from_numpy(w) for w in gx.wf], [torch.from_numpy(b) for b in gx.bf], gx.r, residual=True,
                            observe_out=False, device=dev)

    fq = make()
    mmc = MinMaxCalibrator(fq)
    mmc.observe(frames_all[lo:hi].to(dev))
    mm_scales, mm_zeros, mins, maxs = mmc.finish()
    fq.reset_observers()
    klc = KLCalibrator(fq)
    klc.observe_pass1(frames_all[lo:hi].to(dev))
    klc.begin_pass2()
    klc.observe_pass2(frames_all[lo:hi].to(dev))
    kl_scales, kl_zeros, kl_idx = klc.finish()
    fq.close()
    if rank == 0:
        # the same statistics from ONE process over all frames, no collective involved
        solo = make()
        solo(frames_all.to(dev))
        mm = solo.minmax.cpu()
        s1, z1 = thresholds_from_minmax([float(v) for v in mm[:, 0]], [float(v) for v in mm[:, 1]])
        lows = [float(v) for v in mm[:, 0]]
        lows[-1] = 0.0
        solo.reset_observers()
        solo.enable_histograms([[l, float(h)] for l, h in zip(lows, mm[:, 1])])
        solo(frames_all.to(dev))
        idx1, _ = KLCalibrator(solo).search(solo.hist)
        res["minmax_thresholds_equal"] = (s1 == mm_scales and z1 == mm_zeros)
        res["kl_indices_equal"] = (idx1 == kl_idx)
        res["kl_idx"] = kl_idx
        res["threshold_hash"] = hashlib.sha256(repr((mm_scales, mm_zeros)).encode()).hexdigest()[:16]
        solo.close()
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
