#!/usr/bin/env python
"""What would the north star's element-block variant pay for its exchange?  (DESIGN.md
section 6.)  Each rank would scan M/G elements and send every off-owner incidence to the
rank that owns its row: at 64M triangles and G = 8, 201M incidences, 7/8 of them off-owner,
8 bytes each if the row could be implied (176 MB per rank), 12 bytes with the row index
(264 MB per rank).  This times exactly that all-to-all (NCCL grouped send/recv over
NVLink/NVSwitch, equal splits, device events, max over ranks) -- a LOWER bound of the
variant's extra cost, to be compared with the replicated element scan it would remove
(0.37 ms per rank at any G).

    python -m torch.distributed.run --nproc-per-node 8 ... tools/exchange_probe.py
"""
import json
import os

import torch
import torch.distributed as dist

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local_rank = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local_rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
M = 67108864
out = []
for bytes_per_incidence in (8, 12):
    per_rank = 3 * M // world                               # incidences a rank produces
    per_peer = per_rank // world * bytes_per_incidence      # bytes to each rank (own share stays)
    send = torch.empty(per_peer * world, dtype=torch.uint8, device='cuda')
    recv = torch.empty_like(send)
    for _ in range(5):
        dist.all_to_all_single(recv, send)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    a.record()
    for _ in range(reps):
        dist.all_to_all_single(recv, send)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    off_owner = per_peer * (world - 1)
    out.append({'bytes_per_incidence': bytes_per_incidence, 'n_gpus': world,
                'off_owner_MB_per_rank': off_owner / 1e6, 'all_to_all_ms': float(t.item()),
                'GBs_per_rank_per_direction': off_owner / float(t.item()) / 1e6})
if rank == 0:
    print(json.dumps({'mesh': 'S(12), 201M incidences', 'rows': out}, indent=1))
dist.barrier()
dist.destroy_process_group()
