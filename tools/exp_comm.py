"""Per-op time of the library's all-gather (pm3d_comm_all_gather) under torchrun, a few sizes; and the c5 pieces."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pixelmek-3d_b200"))
import numpy as np, torch, torch.distributed as dist, pm3d
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = pm3d.Context(lr)
uid = [pm3d.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(uid[0], rank, world)
for per in (4096, 1 << 20, (16 << 20) // world, 8 << 20):
    d_send, d_recv = ctx.device_alloc(per), ctx.device_alloc(per * world)
    call = lambda: ctx._check(ctx.lib.pm3d_comm_all_gather(ctx.h, C.c_void_p(d_send), C.c_void_p(d_recv), per), "ag")
    for _ in range(5): call()
    ctx.sync(); dist.barrier()
    ctx.timer_start(0)
    for _ in range(20): call()
    ctx.timer_stop(0)
    ms = ctx.timer_ms(0) / 20
    t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"all-gather {per} B per rank x {world}: {t.item()*1e3:.1f} us per op, {per*world/t.item()/1e6:.1f} GB/s gathered", flush=True)
    ctx.device_free(d_send); ctx.device_free(d_recv)
ctx.close(); dist.destroy_process_group()
