"""CPU, world_size 2 over gloo: the multi-GPU host logic (frame partition, scene
broadcast, ordered gather).  The per-frame work is played by the oracle here -- the
test checks the plumbing, the GPU tests check the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_frame_block_partitions():
    from slr_sfs_b200.sharding import frame_block
    for n in (0, 1, 7, 60, 61):
        for world in (1, 2, 3, 8):
            blocks = [frame_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for a, b in zip(blocks, blocks[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert [hi - lo for lo, hi in (frame_block(60, r, 8) for r in range(8))] == [8, 8, 8, 8, 7, 7, 7, 7]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from slr_sfs_b200 import sharding, workloads
        H, W, C, N = 24, 40, 4, 7
        if rank == 0:
            scene = workloads.scene(H, W, C, "A", seed=4)
        else:
            scene = (((1, C, H, W), torch.float32), ((1, 1, H, W), torch.float32), ((1, 2, H, W), torch.float32))
        feat, Z, motion = sharding.broadcast_scene(scene, src=0)
        ref = workloads.scene(H, W, C, "A", seed=4)
        for a, b in zip((feat, Z, motion), ref):
            assert torch.equal(a, b)

        def make(lo, hi):
            frames = [oracle.joint_splat_baseline(feat.numpy(), Z.numpy(), motion.numpy(), (0, t, N - 1))
                      for t in range(lo, hi)]
            return torch.from_numpy(np.concatenate(frames, 0))

        lo, hi, mine = sharding.synthesize_sharded(make, N)
        assert (lo, hi) == sharding.frame_block(N, rank, world) and mine.shape[0] == hi - lo
        # gather per-frame checksums (small), in frame order
        sums = sharding.all_gather_frames(mine.double().sum(dim=(1, 2, 3)).reshape(-1, 1), N)
        full = make(0, N).double().sum(dim=(1, 2, 3)).reshape(-1, 1)
        assert torch.equal(sums, full)
        # SceneExchange: every rank owns one scene; the prepared buffer travels, two slots are reused
        numel = 5 * H * W
        prepared = []

        def prepare(inputs, out):           # stand-in for JointSplat.prepare_scene (the GPU suite runs the real one)
            prepared.append(1)
            out[:4 * H * W] = (inputs[0] * inputs[1].exp()).reshape(-1)
            out[4 * H * W:] = inputs[1].exp().reshape(-1)

        ex = sharding.SceneExchange(C, H, W, 0, "cpu", numel, prepare=prepare)
        n_scenes = 5
        scenes = [workloads.scene(H, W, C, "A", seed=10 + s) for s in range(n_scenes)]
        ticket = ex.post(0 % world, scenes[0] if rank == 0 % world else None)
        for sidx in range(n_scenes):
            nxt = None
            if sidx + 1 < n_scenes:
                o = (sidx + 1) % world
                nxt = ex.post(o, scenes[sidx + 1] if rank == o else None)
            buf, mot, ready, core_only = ex.take(ticket)
            assert not core_only
            f, z, m = scenes[sidx]
            assert ready is None and torch.equal(mot, m)
            assert torch.equal(buf[:4 * H * W], (f * z.exp()).reshape(-1)) and torch.equal(buf[4 * H * W:], z.exp().reshape(-1))
            ex.used(ticket, None)
            ticket = nxt
        assert len(prepared) == len([s for s in range(n_scenes) if s % world == rank])      # once per scene, on its owner only
        # rotated frame blocks: over `world` scenes every rank gets the same number of frames
        per_rank = sum(hi - lo for lo, hi in (sharding.frame_block(N, rank, world, rotate=s) for s in range(world)))
        assert per_rank == N
        with open(os.path.join(tmp, "ok%d" % rank), "w") as fh:
            fh.write("ok")
    finally:
        dist.destroy_process_group()


def test_world2_gloo_broadcast_shard_gather(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(tmp_path, "ok%d" % r)) for r in range(world))
