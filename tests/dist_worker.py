"""torchrun worker of tests/test_gpu_sharded.py::test_nccl_two_gpus: a sharded run over NCCL,
checked against the CPU oracle on rank 0."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ti_sph_b200.sharded import ShardedSim, TorchDistComm
    from test_cpu_sharded import _scene
    from test_gpu_sharded import check_against_oracle
    for mode in ("reference", "summed"):
        sim = ShardedSim(_scene(), rank, world, comm=TorchDistComm(device=f"cuda:{local}"),
                         density_mode=mode, device=local)
        if rank == 0:
            print("halo path:", "p2p" if sim.p2p else "nccl", flush=True)
            check_against_oracle(lambda: sim.step(1), sim.dump, mode)
        else:
            for _ in range(4):
                sim.step(1)
                sim.dump()
        sim.engine.sync()
        sim.engine.close()
    dist.barrier()
    if rank == 0:
        print("SHARDED-NCCL-OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
