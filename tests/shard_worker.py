"""Worker of tests/test_gpu_parity.py::test_sharded_pipeline_nccl (one process per GPU, torchrun):
NCCL scatter of uint8 frames from rank 0 -> the C5 chain on every rank's shard -> NCCL gather of the sobel maps,
keypoints and rects; rank 0 compares every gathered frame with the oracle chain, for the whole-shard form and
the chunk-overlapped form.  TEST INFRASTRUCTURE (uses the oracle)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _libs as L  # noqa: E402
import grayskull_b200 as g  # noqa: E402
from grayskull_b200 import api, pipeline, shard  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    g._lib.check(g.lib().gs_b200_set_device(local), "set_device")
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, h, w = 7, 192, 256          # ragged shards at world 2 (3 + 4 frames)
    params = dict(nkps=300, max_rects=512)
    cas = g.load_cascade()
    frames = np.stack([L.natural_like(w, h, 100 + i) if i % 2 == 0 else
                       np.random.default_rng(i).integers(0, 256, (h, w)).astype(np.uint8) for i in range(n)])
    root = torch.from_numpy(frames).to(dev) if rank == 0 else None
    lo, hi = shard.shard_range(n, rank, world)
    pipe = pipeline.FramePipeline(cas, hi - lo, h, w, dev, **params)
    run = shard.ShardedRun(pipe, n, h, w, dev)
    want = [L.oracle_chain(cas.ptr, frames[i], **params) for i in range(n)] if rank == 0 else None
    for mode in ("serial", "overlapped-2", "overlapped-3"):
        for t in (run.gathered or []):
            t.zero_()
        pipe.sobel.zero_(); pipe.score.zero_()
        if mode == "serial":
            run.run_serial(root)
        else:
            run.run_overlapped(root, int(mode.split("-")[1]))
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            sob, kps, kc, rects, rc = run.gathered
            got_k, got_r = api.kps_to_numpy(kps, kc), api.rects_to_numpy(rects, rc)
            for i in range(n):
                assert np.array_equal(sob[i].cpu().numpy(), want[i]["sobel"]), (mode, i, "sobel")
                assert got_k[i].tobytes() == want[i]["kps"].tobytes(), (mode, i, "kps")
                assert got_r[i].tobytes() == want[i]["rects"].tobytes(), (mode, i, "rects")
            assert sum(len(k) for k in got_k) > 50
    if rank == 0:
        print("SHARD_OK world=%d scattered=%d B gathered=%d B" % (world, run.bytes_scattered(), run.bytes_gathered()))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
