"""A/B of the two all-reduce routes inside ONE torchrun job (same box, same shards): NVLink peer memory
(csrc/peer_reduce.cuh) against ncclAllReduce, alternating, device-resident iterations of the headline workload.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/peer_ab.py [points] [rounds]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LCBA_PEER_REDUCE", "2")       # map the peer blocks, start on NCCL
import torch  # noqa: E402
import torch.distributed as tdist  # noqa: E402

from lasercalib_b200 import dist as D  # noqa: E402
from lasercalib_b200._cabi import Engine  # noqa: E402
from lasercalib_b200.synth import make_rig  # noqa: E402

rank, ws, local = D.init_from_env()
npts = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pb = make_rig("ring24", npts, seed=0, variant="volume", p_vis=1.0)
sh = D.shard_problem(pb["pts0"], pb["points_2d"], pb["camera_ind"], pb["point_ind"], None, rank, ws)
eng = Engine(local)
eng.set_problem(pb["cams0"], sh["pts"], sh["points_2d"], sh["camera_ind"], sh["point_ind"], pt_offset=sh["pt_offset"])
D.connect_engine(eng)
tol = dict(ftol=1e-4, xtol=1e-8, gtol=1e-8, max_nfev=200)


def run(k):
    done, ms, cost = 0, 0.0, None
    while done < k:
        eng.set_params(pb["cams0"], sh["pts"])
        r, _ = eng.solve(max_iterations=k - done, **tol)
        done += int(r.iterations)
        ms += r.solve_ms
        cost = r.cost
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    return float(t.item()) / done, cost


run(3)
out = []
for rd in range(rounds):
    for mode in (1, 0):
        active = eng.lib.lcba_debug_peer_reduce(eng.h, mode)
        tdist.barrier()
        ms, cost = run(20)
        out.append(dict(round=rd, peer=int(active), ms_per_iteration=ms, final_cost=cost))
if rank == 0:
    for o in out:
        print(json.dumps(o))
    for m in (1, 0):
        v = [o["ms_per_iteration"] for o in out if o["peer"] == m]
        if v:
            print("%s: %.4f ms per iteration (min %.4f) over %d runs of 20 iterations, %d GPUs"
                  % ("NVLink peer memory" if m else "ncclAllReduce     ", float(np.mean(v)), min(v), len(v), ws))
tdist.barrier()
tdist.destroy_process_group()
