"""CPU, world_size 2, gloo: the N>1 host logic -- frame sharding (reference convention,
lib/utils/subprocess.py:56) and the final all-gather of per-frame detections + bit-packed masks."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_clip(frames=6, D=3, h=11, w=13):
    rs = np.random.RandomState(5)
    dets = rs.uniform(size=(frames, D, 6)).astype(np.float32)
    masks = (rs.uniform(size=(frames, D, h, w)) > 0.5).astype(np.uint8)
    return dets, masks


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vosdetectron_b200.pipeline import shard_frames, all_gather_frames, pack_mask_bits, unpack_mask_bits
    dets, masks = _make_clip()
    mine = shard_frames(dets.shape[0], world, rank)
    d = torch.from_numpy(dets[mine])
    m = pack_mask_bits(torch.from_numpy(masks[mine]))
    gd, gm = all_gather_frames(d, m)
    full = unpack_mask_bits(gm, masks.shape[2], masks.shape[3])
    ok = torch.equal(gd, torch.from_numpy(dets)) and torch.equal(full, torch.from_numpy(masks))
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and float(t) == float(world)
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_frame_sharded_all_gather_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok0").read() == "1" and open(tmp_path / "ok1").read() == "1"
