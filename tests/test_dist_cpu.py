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


def _worker_rle(rank, world, port, out_dir):
    """The --gather-rle payload of bench.py: per-rank arena of COCO RLE strings + (offset, length) columns appended to
    the detection records, gathered with the same all_gather_frames (async handles included)."""
    import sys
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import region_oracle as orc
    from vosdetectron_b200.core.test import rle_encode, rle_counts_from_string
    from vosdetectron_b200.pipeline import shard_frames, all_gather_frames
    dets, masks = _make_clip()
    F, D, h, w = masks.shape
    mine = shard_frames(F, world, rank)
    cap = len(mine) * D * 256                              # fixed arena capacity per rank (bytes)
    arena = np.zeros(cap, dtype=np.uint8)
    rec = np.zeros((len(mine), D, 8), dtype=np.float32)
    rec[:, :, :6] = dets[mine]
    cur = 0
    for i, f in enumerate(mine):
        for d in range(D):
            s = rle_encode(masks[f, d])['counts'].encode('ascii')
            arena[cur:cur + len(s)] = np.frombuffer(s, dtype=np.uint8)
            rec[i, d, 6], rec[i, d, 7] = cur, len(s)
            cur += len(s)
    assert cur <= cap
    grec, garena, works = all_gather_frames(torch.from_numpy(rec), torch.from_numpy(arena).view(len(mine), -1), async_op=True)
    for wk in works:
        wk.wait()
    ok = torch.equal(grec[:, :, :6], torch.from_numpy(dets))
    per_rank = F // world
    flat = garena.reshape(world, -1).numpy()               # rank r's arena = row r
    for f in range(F):
        r = f // per_rank
        for d in range(D):
            o, n = int(grec[f, d, 6]), int(grec[f, d, 7])
            counts = rle_counts_from_string(flat[r, o:o + n].tobytes())
            ok = ok and np.array_equal(orc.rle_decode(counts, h, w), masks[f, d])
    open(os.path.join(out_dir, "rle%d" % rank), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_frame_sharded_rle_payload_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker_rle, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "rle0").read() == "1" and open(tmp_path / "rle1").read() == "1"


def test_frame_sharded_all_gather_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok0").read() == "1" and open(tmp_path / "ok1").read() == "1"


def _worker_clip(rank, world, port, out_dir):
    """run_clip (the frame-sharded clip driver) on CPU: 3 batches of 2 frames per rank through a stand-in step,
    gathered with two rotating slots; the result must be the clip in (batch, rank, frame) order on every rank."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vosdetectron_b200.pipeline import FrameGather, pack_mask_bits, run_clip, unpack_mask_bits
    nb, fr, D, h, w = 3, 2, 3, 11, 13
    dets, masks = _make_clip(frames=nb * world * fr, D=D, h=h, w=w)
    # frame f of the clip = batch f // (world * fr), rank (f // fr) % world
    mine = [[b * world * fr + rank * fr + i for i in range(fr)] for b in range(nb)]
    batches = [{"frames": idx} for idx in mine]
    calls = []

    def step_fn(batch):
        calls.append(tuple(batch["frames"]))
        return {"masks_packed": pack_mask_bits(torch.from_numpy(masks[batch["frames"]]))}

    def records_fn(batch, out):
        return torch.from_numpy(dets[batch["frames"]]), out["masks_packed"]

    gd, gm = run_clip(batches, step_fn, records_fn)
    ok = len(calls) == nb and torch.equal(gd, torch.from_numpy(dets))
    ok = ok and torch.equal(unpack_mask_bits(gm, h, w), torch.from_numpy(masks))
    # a single-rank gather is the identity and allocates nothing symmetric
    g1 = FrameGather((fr, D, 6), torch.float32, (fr, D, 4), torch.uint8, "cpu", transport="nccl")
    ok = ok and g1.transport == "nccl"
    open(os.path.join(out_dir, "clip%d" % rank), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_run_clip_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker_clip, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "clip0").read() == "1" and open(tmp_path / "clip1").read() == "1"
