"""N > 1 host logic on CPU (gloo, world_size 2): the ray-range partition every rank applies and
the sum-reduce of the private images.  The data path has no other collective.
The per-rank tracing here is done by the CPU oracle because this box has no GPU -- what is under
test is the partition + reduce protocol bench.py / ort_init_rank use, not the kernels."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nrays, phase, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opticalraytrace_b200 import abi, partition
    from tests import cases, oracle_lib as O
    scene = cases.scene_for(O, cases.C2, phase)
    first, cnt = partition(nrays, rank, world, first_ray=11)
    img, lost, hist = O.trace(abi.default_job(phase, cnt, first_ray=first), scene, nthreads=2)
    # the reduce the library does with ncclReduce(sum, uint64) -- int64 view, same bits
    buf = torch.from_numpy(np.concatenate([img.ravel().view(np.int64), hist.ravel()]))
    dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)
    ranges = [None] * world
    dist.all_gather_object(ranges, (first, cnt))
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), buf.numpy())
        np.save(os.path.join(out_dir, "ranges.npy"), np.array(ranges))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("phase", [1, 2])
def test_two_rank_partition_and_reduce(orc, tmp_path, phase):
    from opticalraytrace_b200 import abi
    from tests import cases
    nrays, world = 300_001, 2
    mp.spawn(_worker, args=(world, _free_port(), nrays, phase, str(tmp_path)), nprocs=world, join=True)
    red = np.load(tmp_path / "reduced.npy")
    ranges = np.load(tmp_path / "ranges.npy")
    # disjoint cover of [11, 11 + nrays)
    assert ranges[0][0] == 11 and ranges[0][0] + ranges[0][1] == ranges[1][0]
    assert ranges[1][0] + ranges[1][1] == 11 + nrays
    scene = cases.scene_for(orc, cases.C2, phase)
    img, lost, hist = orc.trace(abi.default_job(phase, nrays, first_ray=11), scene)
    assert np.array_equal(red[:abi.ORT_IMG_BINS].view(np.uint64), img.ravel())
    assert np.array_equal(red[abi.ORT_IMG_BINS:], hist.ravel())


def test_partition_properties():
    from opticalraytrace_b200 import partition
    for n in (0, 1, 7, 10 ** 11, 2 ** 40 + 3):
        for g in (1, 2, 3, 4, 8):
            parts = [partition(n, r, g, first_ray=5) for r in range(g)]
            assert parts[0][0] == 5
            for a, b in zip(parts, parts[1:]):
                assert a[0] + a[1] == b[0]
            assert parts[-1][0] + parts[-1][1] == 5 + n
            sizes = [p[1] for p in parts]
            assert max(sizes) - min(sizes) <= 1


def _rdv_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_PORT"] = str(port)
    os.environ["TORCHELASTIC_RUN_ID"] = "pytest"
    import bench
    r = bench.Rendezvous(rank, world)
    got = r.bcast(b"nccl-unique-id-stand-in" if rank == 0 else b"")
    rows = [r.allgather([rank, k, 0.5 * rank]) for k in range(40)]
    r.barrier()
    r.close()
    ok = got == b"nccl-unique-id-stand-in" and all(row == [[i, k, 0.5 * i] for i in range(world)] for k, row in enumerate(rows))
    open(os.path.join(out_dir, "rdv.%d" % rank), "w").write("ok" if ok else "bad")


def test_bench_rendezvous_without_torch_distributed(tmp_path):
    """bench.py's ranks find each other through files (no torch.distributed): broadcast of the NCCL id,
    all-gather of the timings, barrier, and a teardown in which rank 0 removes the directory only after
    every rank has finished reading (the race that hung a 2-GPU run once)."""
    world = 3
    mp.spawn(_rdv_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert [open(tmp_path / ("rdv.%d" % r)).read() for r in range(world)] == ["ok"] * world
    import glob
    import tempfile
    assert not glob.glob(os.path.join(tempfile.gettempdir(), "ort_bench_*_pytest_%d" % os.getpid()))
