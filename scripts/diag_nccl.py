"""Latency of the time-sharding collectives in isolation (torchrun --nproc-per-node 2)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from poor_man_gplvm_b200.shard import TimeShard
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
sh = TimeShard()
K, N = 400, 500
a = torch.rand(4 * K, device=dev); b = torch.rand(4 * K, device=dev)
yw = torch.rand((K, N), device=dev); tw = torch.rand(K, device=dev); m = torch.rand(2, device=dev)
big = torch.rand(64 << 20, device=dev)
def t(fn, name, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); th = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    if rank == 0: print("%-44s gpu %.1f us/call   host enqueue %.1f us/call" % (name, e0.elapsed_time(e1) / n * 1e3, th), flush=True)
t(lambda: sh.boundary(a, b), "boundary exchange both ways (4K floats)")
t(lambda: sh.boundary(a, None), "boundary exchange to the left only")
t(lambda: sh.allreduce_sum_(yw, tw), "packed all-reduce yw+tw (fp64 on the wire)")
t(lambda: sh.allreduce_max_(m), "all-reduce MAX of 2 floats")
t(lambda: dist.all_reduce(yw), "plain all_reduce fp32 yw")
def with_busy(fn):
    def g():
        big.mul_(1.0001); fn()
    return g
t(with_busy(lambda: sh.boundary(a, b)), "256MB elementwise kernel + boundary exchange")
t(lambda: big.mul_(1.0001), "256MB elementwise kernel alone")
dist.destroy_process_group()
