"""One rank of a large world split over the GPUs of one box, one process per GPU (launched by torch.distributed.run):
export -> all_gather of the 128-byte descriptions -> attach -> step; rank 0 also steps an unsplit batch and every rank
checks its state against it bit for bit (broadcast)."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import split_util as su  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="pile_10x10x20")
    ap.add_argument("--steps", type=int, default=70)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib, scenes = su.load("b200", "single")
    B = su.build(lib, scenes, a.scene, local)
    mine = torch.frombuffer(bytearray(su.export(lib, B)), dtype=torch.uint8).cuda()
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    su.attach(lib, B, rank, [bytes(t.cpu().numpy().tobytes()) for t in allh])
    dist.barrier()
    su.step(lib, B, 0.01, a.steps)
    got = su.state(lib, B)
    want = torch.zeros(got.shape, dtype=torch.float32, device="cuda")
    if rank == 0:
        R = su.build(lib, scenes, a.scene, local)
        su.step(lib, R, 0.01, a.steps)
        want.copy_(torch.from_numpy(su.state(lib, R)))
    dist.broadcast(want, 0)
    same = torch.tensor([int(want.cpu().numpy().tobytes() == got.tobytes())], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SPLIT_OK" if int(same.item()) == 1 else "SPLIT_MISMATCH", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(same.item()) == 1 else 1)


if __name__ == "__main__":
    main()
