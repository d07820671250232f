"""Multi-GPU verification (run under torchrun, one rank per GPU):
shard c3 / c4-like stacks by (image, slice) over the ranks, render each share with the fused kernel,
all_gather over NCCL and compare bit-for-bit with the full stack rendered by every rank alone.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29611 tests/gpu_multi_verify.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aadff_b200  # noqa: E402
from oracle import focal_stack_oracle as orc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    lens = aadff_b200.PSFNet(kernel_size=11, device=f"cuda:{local}")
    lens.load_net(os.path.join(ROOT, "tests/golden/rf50mm_PSFNet480x640_ks11.pkl"))
    ok = True
    for (N, S, H, W) in [(16, 5, 256, 256), (1, 10, 270, 480), (3, 7, 40, 56)]:
        img, dm = orc.synthetic_rgbd(N, H, W, seed=2024)           # same inputs on every rank
        foc = -orc.synthetic_focus(dm, S) * 1e3
        img, dep, foc = img.cuda(), -dm.cuda() * 1e3, foc.cuda()
        full, runs = aadff_b200.sharding.render_stack_sharded(lens, img, dep, foc, rank, world)
        ref = lens.render_stack(img, dep, foc)
        same = bool(torch.equal(full, ref))
        ok &= same
        print(f"rank {rank}/{world} N{N} S{S} {H}x{W}: runs={runs} gathered==single-GPU: {same}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_VERIFY", "PASS" if int(flag) == 1 else "FAIL")
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
