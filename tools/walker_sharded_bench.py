"""C3 star sampled by ONE ensemble sharded by walker across the GPUs of a box (strong scaling of a single
star).  Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N
--master-addr 127.0.0.1 --master-port P tools/walker_sharded_bench.py [steps].  Time = barrier +
synchronize on both sides, max over ranks (the loop is host-driven: two small syncs per half-step)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
from gptransits_b200 import Engine
from gptransits_b200.parallel import run_walker_sharded

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local)); dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)

class A:
    walkers = 0
    gpus = 1
work = bench.build_workload("c3", 0, A)          # the SAME star on every rank
t = work["t"]
np.random.seed(42)
gp, tr, spec = bench.model_of(work, t)
eng = Engine(local)
ys = bench.inject_transit(eng, work, spec)
eng.upload_star(0, t, ys[0], work["yerr"])
W = work["W"]
init = bench.draw_prior(gp, tr, W)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
d = dist if world > 1 else None
run_walker_sharded(eng, d, init, 5, seed=42)      # warm-up (allocations, halo adaptation)
if d: dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
chain, logp = run_walker_sharded(eng, d, init, steps, seed=42)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if d:
    v = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(v, op=dist.ReduceOp.MAX); dt = float(v.item())
if rank == 0:
    print(json.dumps({"workload": work["label"], "sharding": "one star, ensemble sharded by walker", "n_gpus": world,
                      "steps": steps, "ms_per_step": 1e3 * dt / steps, "evals_per_s": W * steps / dt,
                      "chunks": eng.stat("chunks"), "chunk_len": eng.stat("chunk_len"), "halo": eng.stat("halo"),
                      "note": "includes the initial whole-ensemble evaluation and the device-to-host read of the chain"}))
if d: dist.destroy_process_group()
