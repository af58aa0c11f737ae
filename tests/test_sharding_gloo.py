"""CPU, world_size 2 over gloo: the multi-GPU path end to end -- frame partition, scene broadcast, the
prepared-scene exchange, ordered gather -- with the per-frame work done by the KERNEL SOURCES (tests/emu: the CUDA
text compiled for the CPU, same C ABI) and checked against the oracle; the GPU suite and `bench.py --check` run the
same logic over NCCL."""
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

        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import emu
        from conftest import rel_err
        kernels = emu.Scene(feat.numpy(), Z.numpy(), motion.numpy())

        def make(lo, hi):            # this rank's frame block through slr_clip_table / bin / expand / gather / heavy
            return torch.from_numpy(kernels.frames(0, N - 1, lo, hi - lo, table=kernels.table(0, N - 1, lo, hi - lo)))

        lo, hi, mine = sharding.synthesize_sharded(make, N)
        assert (lo, hi) == sharding.frame_block(N, rank, world) and mine.shape[0] == hi - lo
        for i, t in enumerate(range(lo, hi)):
            want = oracle.joint_splat_baseline(feat.numpy(), Z.numpy(), motion.numpy(), (0, t, N - 1))
            assert rel_err(mine[i:i + 1].numpy(), want) <= 1e-4, t
        # gather per-frame checksums (small), in frame order, against one rank synthesising every frame alone
        sums = sharding.all_gather_frames(mine.double().sum(dim=(1, 2, 3)).reshape(-1, 1), N)
        full = make(0, N).double().sum(dim=(1, 2, 3)).reshape(-1, 1)
        assert torch.allclose(sums, full, rtol=1e-9, atol=1e-9)
        # SceneExchange: every rank owns scenes in turn; the scene buffer PREPARED by the owner (slr_reduce_max +
        # slr_scene_prep, here the emulated library) travels, two slots are reused, and every rank synthesises its
        # rotated frame block of every scene from the received buffer
        numel = (emu.lib().slr_scene_bytes(C, 0, H, W) + 3) // 4
        prepared = []

        def prepare(inputs, out):
            prepared.append(1)
            f_, z_ = emu.f32(inputs[0].numpy()), emu.f32(inputs[1].numpy())
            zmax = np.zeros(1, np.float32)
            emu.call("slr_reduce_max", emu.p(z_), z_.size, emu.p(zmax), None)
            emu.call("slr_scene_prep", emu.p(f_), emu.p(z_), emu.p(zmax), None, 0, ctypes_ptr(out), C, H, W, None)

        import ctypes

        def ctypes_ptr(t):
            assert t.data_ptr() % 32 == 0
            return ctypes.c_void_p(t.data_ptr())

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
            blo, bhi = sharding.frame_block(N, rank, world, rotate=sidx)
            rx = emu.Scene.from_buffer(buf.numpy(), mot.numpy(), C, H, W)
            got = rx.frames(0, N - 1, blo, bhi - blo)
            for i, t in enumerate(range(blo, bhi)):
                want = oracle.joint_splat_baseline(f.numpy(), z.numpy(), m.numpy(), (0, t, N - 1))
                assert rel_err(got[i:i + 1], want) <= 1e-4, (sidx, t)
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
