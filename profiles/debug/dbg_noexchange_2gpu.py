import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, "/root/repo")
import slr_sfs_b200 as pkg
from slr_sfs_b200 import workloads
from slr_sfs_b200.clip import ClipRunner
from slr_sfs_b200.sharding import frame_block
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
H, W, C, N = 768, 1024, 64, 60
own = tuple(t.to(dev) for t in workloads.scene(H, W, C, "A", seed=rank))
runner = ClipRunner(C, H, W, dev, group=30)
def step():
    for k in range(world):
        lo, hi = frame_block(N, rank, world, rotate=k)
        runner.run(pkg.JointSplat(*own, inputs_event=False), 0, N - 1, lo, hi)
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize()
    if rank == 0: print("step", i, "%.2f ms" % ((time.perf_counter() - t0) * 1e3), "alloc MB", torch.cuda.memory_allocated() // 2**20, "reserved", torch.cuda.memory_reserved() // 2**20, flush=True)
dist.destroy_process_group()
