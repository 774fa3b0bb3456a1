"""Worker of tests/test_gpu_multi.py (launched with torch.distributed.run, one rank per GPU): scenes sharded over the
ranks, stepped independently, results gathered through mmw_gather_nccl; rank 0 re-runs EVERY rank's scenes in one
context of its own and compares the gathered records bitwise (1-GPU and G-GPU outputs identical per
scene)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmwave_msc_b200 import _lib, pose_weights as pw, sharding, synth  # noqa: E402
from mmwave_msc_b200.batched import BatchedTracker  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S, F = 24, 14
    W = pw.make_pose_weights(pw.VARIANT_3D)
    ids = sharding.shard_scene_ids(world * S, world, rank)
    bt = BatchedTracker(S, device=local)
    bt.load_pose_weights(W)
    for b in synth.gen_batch(ids, F):
        bt.step(b.points, b.offsets, b.dt, pose=True)
    ng = sharding.NcclGather(bt)
    out = ng.gather()
    bt.sync()
    got = out.cpu().numpy().view(np.uint32).reshape(world, -1)
    ok = True
    if rank == 0:
        one = BatchedTracker(world * S, device=local)          # all scenes on one GPU
        one.load_pose_weights(W)
        for b in synth.gen_batch(list(range(world * S)), F):
            one.step(b.points, b.offsets, b.dt, pose=True)
        ref = torch.empty(world * S * one.tcap * _lib.RESULT_FLOATS, dtype=torch.float32, device="cuda")
        one.pack_results(ref.data_ptr())
        one.sync()
        want = ref.cpu().numpy().view(np.uint32).reshape(world, -1)
        ok = bool(np.array_equal(got, want)) and float(np.abs(ref.cpu().numpy()[::72] + 1).max()) > 0
        print("MULTI_GPU_IDENTITY world=%d scenes=%d tracks=%d identical=%s" % (
            world, world * S, int((ref.cpu().numpy()[::72] >= 0).sum()), ok), flush=True)
    ng.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
