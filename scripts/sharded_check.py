"""torchrun worker: sharded build == single-GPU build == oracle (run with torchrun --nproc-per-node=R)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bench import latent
from pynndescent_b200 import NNDescent

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, world = dist.get_rank(), dist.get_world_size()
# (metric, n, d, k, seed, push_capacity): a small push_capacity forces several emission rounds per iteration and
# the split-by-row replay of the collected half-updates (sharded.py:_replay)
cases = [("euclidean", 20000, 32, 15, 42, None), ("cosine", 8000, 24, 10, 7, None), ("euclidean", 3000, 7, 30, 1, None),
         ("euclidean", 3000, 7, 15, 1, None), ("euclidean", 3000, 16, 30, 1, None), ("euclidean", 3000, 16, 20, 1, None),
         ("euclidean", 20000, 32, 15, 42, 1 << 16), ("euclidean", 3000, 16, 30, 1, 1 << 16), ("cosine", 8000, 24, 10, 7, 1 << 15)]
if len(sys.argv) > 1:
    cases = [("euclidean", int(sys.argv[1]), 128, 15, 42, None)]
ok = True
for metric, n, d, k, seed, cap in cases:
    X = latent(n, d, seed)
    t0 = time.perf_counter()
    sh = NNDescent(X, metric, n_neighbors=k, random_state=seed, sharded=True, push_capacity=cap)
    torch.cuda.synchronize(); t_sh = time.perf_counter() - t0
    gi, gd = sh.neighbor_graph
    if rank == 0:
        t0 = time.perf_counter()
        one = NNDescent(X, metric, n_neighbors=k, random_state=seed)
        torch.cuda.synchronize(); t_one = time.perf_counter() - t0
        oi, od = one.neighbor_graph
        same = np.array_equal(gi, oi) and np.array_equal(gd, od) and sh.c_hist == one.c_hist
        print(f"[{metric} {n}x{d} k={k} cap={cap}] world={world} sharded==single: {same}  c_hist {sh.c_hist[:4]}.. t_sharded {t_sh:.3f}s t_single {t_one:.3f}s "
              f"timings {sh.timings} comm_MB {sh._sharded.comm_bytes/1e6:.1f}", flush=True)
        if not same:
            bad = np.where((gi != oi).any(1))[0]
            print("   differing rows:", len(bad), bad[:10], "c_hist sharded", sh.c_hist, "single", one.c_hist, flush=True)
        ok = ok and same
    dist.barrier()
if rank == 0:
    print("SHARDED_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
