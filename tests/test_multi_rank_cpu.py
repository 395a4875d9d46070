"""world_size-2 gloo test (CPU) of the host-side multi-GPU logic: the batch shards by contiguous
slices with no exchange during the solve; a shard's inputs are generated in place from
(seed, first_config); the optional collective is an all-gather of the poses.  The per-rank
'solve' here is the sample-count helper of the product's host library plus the CPU oracle standing
in as the checker — the CUDA solve itself is covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out_dir):
    for p in (os.path.join(REPO, "ram2017-ctr-kinematics_b200"), os.path.join(REPO, "oracle"), HERE):
        sys.path.insert(0, p)
    import ctrfk
    import oracle_lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = ctrfk.desc_from_xml(ctrfk.resource("ctr_sec3_tub3.xml"), 64)
    lo, hi = ctrfk.shard_range(total, world, rank)
    per = -(-total // world)
    jv = ctrfk.sample_joints_host(d, 20172, lo, hi - lo)               # the rank's own slice
    r = oracle_lib.fk_batch(d, jv, ctrfk.MODE_STEP, step=d.max_length / 64.0, nthreads=1)
    n_host = ctrfk.count_samples_host(d, jv, ctrfk.MODE_STEP, step=d.max_length / 64.0)
    assert np.array_equal(n_host, r["n"])
    # all-gather of the poses (padded to the per-rank slice size, like the NCCL path)
    tip = torch.zeros(per, 12, dtype=torch.float64)
    tip[: hi - lo] = torch.from_numpy(r["tip"])
    gathered = [torch.zeros_like(tip) for _ in range(world)]
    dist.all_gather(gathered, tip)
    count = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(count)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = torch.cat(gathered)[:total].numpy()
        np.save(os.path.join(out_dir, "gathered.npy"), full)
        np.save(os.path.join(out_dir, "meta.npy"), np.array([count.item(), t.item()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [1001, 64])
def test_two_rank_sharding_and_gather(tmp_path, total, ctrfk, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    full = np.load(tmp_path / "gathered.npy")
    meta = np.load(tmp_path / "meta.npy")
    assert meta[0] == total and meta[1] == 2.0
    d = ctrfk.desc_from_xml(ctrfk.resource("ctr_sec3_tub3.xml"), 64)
    jv = ctrfk.sample_joints_host(d, 20172, 0, total)
    want = oracle.fk_batch(d, jv, ctrfk.MODE_STEP, step=d.max_length / 64.0)["tip"]
    assert np.array_equal(full, want)            # sharded + gathered == one global solve, bit for bit


def test_shard_ranges_tile_the_batch(ctrfk):
    for total in (0, 1, 7, 8, 1000, 16_000_000):
        for world in (1, 2, 4, 8):
            spans = [ctrfk.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            assert max(hi - lo for lo, hi in spans) == -(-total // world)


def test_reference_arm_under_torchrun_prints_once(tmp_path):
    """`bench.py --impl reference` launched like the driver launches it for N > 1 (torchrun, one rank per GPU): rank 0
    alone times the CPU arm and prints ONE JSON line, the other ranks exit 0 without work.  No GPU involved."""
    import json
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(repo, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--cpu-sample", "2000"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=repo)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
