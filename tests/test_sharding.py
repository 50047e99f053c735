"""Frame sharding across ranks: host logic + a world_size-2 gloo run on CPU in which
every rank processes only its frames and the reassembled result equals the unsharded one."""
import importlib
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def test_frame_range_partitions_exactly():
    sh = importlib.import_module("tce-rvos_b200.sharding")
    for n in (0, 1, 5, 10, 36, 64, 65, 1024):
        for world in (1, 2, 3, 4, 8):
            spans = [sh.frame_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.frame_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, ret):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = importlib.import_module("tce-rvos_b200.sharding")
        from oracle import msda_oracle as O
        levels = [(5, 6), (3, 3)]
        inp = O.make_inputs("stress", n_frames, levels, Lq=7, M=2, D=4, P=2, seed=3, dtype=torch.float64)
        v, loc, aw, go = sh.shard_frames(rank, world, inp["value"], inp["loc"], inp["aw"], inp["grad_out"])
        # the oracle stands in for the kernel on CPU; the thing under test is the partition
        out, gv, gl, ga = O.core_pytorch_fwd_bwd(v, inp["shapes"], loc, aw, go)
        full = [sh.gather_frames(t, n_frames) for t in (out, gv, gl, ga)]
        if rank == 0:
            ref = O.core_pytorch_fwd_bwd(inp["value"], inp["shapes"], inp["loc"], inp["aw"], inp["grad_out"])
            ret["ok"] = all(torch.allclose(a, b, rtol=0, atol=1e-14) for a, b in zip(full, ref))
            ret["sizes"] = [tuple(t.shape) for t in full]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 4])
def test_two_rank_frame_sharding_gloo(n_frames):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n_frames, ret), nprocs=2, join=True)
    assert ret.get("ok") is True, dict(ret)
    assert ret["sizes"][0][0] == n_frames
