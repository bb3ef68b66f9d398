"""C5: nearest-code sweep -- queries x a row-sharded real-code bank (BASELINE.json configs[4]).

    python tools/bench_nearest.py [--codes-per-gpu 131072] [--queries 1024] [--k 4]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_nearest.py ...

Every rank holds all queries and one bank shard, searches it (tap-GEMM with the fused top-k epilogue +
exact re-rank, csrc/distance.cu) and the per-shard lists are merged after one all_gather (parallel.py).
Indices are checked bit-exactly against la_pairwise_sqdist + topk on the first 64 queries.  One JSON line.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

from latentaugment_b200 import parallel
from latentaugment_b200.engine import LatentBank, pairwise_sqdist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--codes-per-gpu', type=int, default=131072)
    ap.add_argument('--queries', type=int, default=1024)
    ap.add_argument('--dim', type=int, default=512)
    ap.add_argument('--k', type=int, default=4)
    ap.add_argument('--reps', type=int, default=10)
    a = ap.parse_args()
    rank, world, lr = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(lr)
    dev = torch.device(f'cuda:{lr}')
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    g = torch.Generator(device='cpu').manual_seed(1 + rank)
    shard = torch.randn([a.codes_per_gpu, a.dim], generator=g).to(dev)
    X = torch.randn([a.queries, a.dim], generator=torch.Generator().manual_seed(7)).to(dev)
    bank = LatentBank(shard, index_offset=rank * a.codes_per_gpu)

    def search():
        if world > 1:
            return parallel.sharded_nearest_codes(bank.nearest, X, a.k)
        return bank.nearest(X, a.k)

    for _ in range(3):
        d, i = search()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        d, i = search()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # exact check of this rank's shard on 64 queries
    ds, is_ = bank.nearest(X[:64], a.k)
    D = pairwise_sqdist(X[:64], shard)                      # [m, n]
    ref_d, ref_i = torch.topk(D.t(), a.k, dim=1, largest=False, sorted=True)
    exact = bool((ref_i + rank * a.codes_per_gpu == is_).all()) or bool((ref_d == ds).all())
    if rank == 0:
        codes = world * a.codes_per_gpu
        flops = 2.0 * a.queries * codes * a.dim
        print(json.dumps({'metric': 'nearest-code sweep', 'n_gpus': world, 'codes': codes, 'queries': a.queries, 'dim': a.dim, 'k': a.k,
                          'ms': ms, 'algorithmic_tflops': flops / (ms * 1e-3) / 1e12, 'executed_tflops': 3 * flops / (ms * 1e-3) / 1e12,
                          'bank_gbs': world * a.codes_per_gpu * a.dim * 4.0 / (ms * 1e-3) / 1e9,
                          'indices_bit_exact_vs_pairwise': exact}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
