"""world_size-N gloo worker for tests/test_host_cpu.py::test_shard_clips_gloo_world2."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist  # noqa: E402

from video_heart_rate_b200 import parallel  # noqa: E402

rank, world = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo", rank=rank, world_size=world)
n_clips = 7
mine = parallel.shard_units(n_clips, rank, world)
assert mine == list(range(rank, n_clips, world))
# each rank "measures" its clips: bpm = 60 + clip id, two windows per clip
local = {c: np.array([60.0 + c, 61.0 + c]) for c in mine}
full = parallel.gather_results(local, n_clips, width=2)
if rank == 0:
    exp = np.stack([[60.0 + c, 61.0 + c] for c in range(n_clips)])
    assert np.array_equal(full, exp), full
else:
    assert full is None
# weighted round-robin for unequal work (config c5)
costs = [5, 1, 9, 3, 7, 2, 8, 4]
a = parallel.shard_by_cost(costs, world)
assert sorted(sum(a, [])) == list(range(8))
assert abs(sum(costs[i] for i in a[0]) - sum(costs[i] for i in a[1])) <= max(costs)
dist.barrier()
dist.destroy_process_group()
print(f"OK rank {rank}")
