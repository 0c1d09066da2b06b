#!/usr/bin/env python
"""bench.py -- P1 stiffness assembly to CSR on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--level L]

A "step" is one complete cold assembly of the stiffness matrix of the mesh
S(L) (unit square cut by its centre, L uniform NVB refinements; L=12 is the
64M-triangle mesh the north star quotes) from device-resident
(coordinates, elements) to device-resident CSR (indptr, indices, data):
plan (node->element lists, row pointer) + numeric fill, nothing cached between
steps.  Prints ONE JSON line (see the contract in the task description).

N > 1 (torchrun): every rank holds the mesh and assembles its share of the rows
(chunks of 4096 rows dealt round-robin; strong scaling of one 64M mesh; no
data-path collective, only the all-gather of per-rank nnz).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = 'P1 stiffness assembly Melements/s to CSR'
UNIT = 'Melements/s'


def mesh_counts(level):
    n = 2 ** level
    m = 4 * n * n
    nn = 2 * n * n + 2 * n + 1
    nnz = 14 * n * n + 6 * n + 1
    return m, nn, nnz


def algorithmic_bytes(m, nn, nnz):
    """every input read once, every output written
    once; int32 indices, fp64 values: 12 M + 16 N + 12 nnz + 4 (N+1)."""
    return 12 * m + 16 * nn + 12 * nnz + 4 * (nn + 1)


def generate_mesh(level, keep=None):
    """S(level) on the host; `keep` collects intermediate levels."""
    from p1afempy_b200 import meshgen
    c, e, b = meshgen.criss_square()
    for k in range(1, level + 1):
        c, e, b = meshgen.refine_nvb(c, e, np.arange(e.shape[0]), b)
        if keep is not None and k in keep:
            keep[k] = (c, e)
    return c, e


def cpu_info():
    model = 'unknown'
    try:
        with open('/proc/cpuinfo') as fh:
            for line in fh:
                if line.startswith('model name'):
                    model = line.split(':', 1)[1].strip()
                    break
    except OSError:
        pass
    return model, os.cpu_count(), len(os.sched_getaffinity(0))


def time_reference(c, e, steps, warmup):
    """The reference's CPU path (oracle port, bit-identical to the reference:
    tests/test_oracle_vs_reference.py): get_stiffness_matrix(c, e).tocsr().
    Single thread -- numpy/scipy do not thread this path."""
    from oracle import p1_oracle as orc
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        a = orc.stiffness_coo(c, e).tocsr()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        del a
    return times


class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix='.csv')
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.fh = open(self.path, 'w')
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.QUERY,
                 '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=self.fh, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        with open(self.path) as fh:
            for line in fh:
                parts = [p.strip() for p in line.split(',')]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


def read_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except (OSError, KeyError, ValueError):
        return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(kernel, level):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`
    from the committed `ncu --set full` capture (profiles/traffic.json; taken
    on this workload at this level), or None when there is no capture."""
    try:
        with open(os.path.join(REPO, 'profiles', 'traffic.json')) as fh:
            table = json.load(fh)
        entry = table.get('level_%d' % level, {}).get(kernel)
        return int(entry['dram_bytes']) if entry else None
    except (OSError, ValueError, KeyError):
        return None


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path
    (oracle port, pinned bit-identical to /root/reference) on host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    level = min(args.level, args.ref_level)
    c, e = generate_mesh(level)
    m = e.shape[0]
    times = time_reference(c, e, args.steps, min(args.warmup, 1))
    t = float(np.mean(times))
    model, ncpu, naff = cpu_info()
    value = m / t / 1e6
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': min(args.warmup, 1),
        'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'stiffness->CSR, S(%d) unit square, M=%d triangles '
                               '(bounded sample of the S(%d) workload)'
                               % (level, m, args.level)},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                         'sample': 'oracle port of get_stiffness_matrix(...).tocsr() on '
                                   'S(%d), M=%d, mean of %d runs; host: %s, %d cpus '
                                   '(%d usable), numpy/scipy single-threaded on this path'
                                   % (level, m, args.steps, model, ncpu, naff)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def parse_profile(text):
    out = {}
    for row in text.strip().splitlines():
        name, ms, launches = row.split()
        out[name] = (float(ms), int(launches))
    return out


def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from p1afempy_b200 import _lib, solvers
    from p1afempy_b200.device import DeviceMesh, context

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a GPU (no CPU fallback); use --impl reference '
                         'for the CPU arm')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    level = args.level
    m, nn, nnz = mesh_counts(level)

    # ---- synthetic input: the mesh the reference's refineNVB would produce,
    # generated on the device (meshgen.unit_square_device: bit-identical to the
    # host generator, tests/test_gpu_dist.py), every rank for itself
    from p1afempy_b200 import meshgen
    t0 = time.perf_counter()
    c, e, _ = meshgen.unit_square_device(level, device=local_rank)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    assert e.shape[0] == m and c.shape[0] == nn

    ctx, dev = context(local_rank)
    lib = _lib.load()
    # N > 1: chunks of 4096 rows dealt out round-robin (p1_plan_create_interleaved)
    mesh = DeviceMesh(c, e, device=local_rank,
                      interleave=(world, rank, 12) if world > 1 else None)
    stream = torch.cuda.current_stream()
    nnz_all = torch.zeros(world, dtype=torch.int32, device='cuda')
    pending = []

    def step():
        ip, ix, data = mesh.stiffness_csr()   # cold: plan + fill, nothing cached
        if world > 1:   # global row-pointer offsets: every rank learns every rank's nnz
            # (ip[-1] = local nnz, on device; asynchronous: NCCL orders it after the
            # fill on its own stream, the host does not wait)
            pending.append(dist.all_gather_into_tensor(nnz_all, ip[-1:], async_op=True))
        return ip, ix, data

    for _ in range(args.warmup):
        out = step()
    for work in pending:
        work.wait()
    pending.clear()
    torch.cuda.synchronize()
    if world == 1:
        assert mesh.nnz == nnz, (mesh.nnz, nnz)
    del out

    sampler = ClockSampler(local_rank)
    lib.p1_profile_reset(ctx)
    lib.p1_profile_enable(ctx, 0 if os.environ.get('P1_NO_PROF') else 1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0 and not os.environ.get('P1_NO_CLOCKS'):
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        out = step()
    for work in pending:
        work.wait()          # stream-level wait: the collectives are inside the timed region
    pending.clear()
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    buf = C.create_string_buffer(1 << 14)
    lib.p1_profile_report(ctx, buf, len(buf))
    lib.p1_profile_enable(ctx, 0)
    prof = parse_profile(buf.value.decode())
    if world > 1:
        tmax = torch.tensor([ms_total], dtype=torch.float64, device='cuda')
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
        total_nnz = int(nnz_all.to(torch.int64).sum().item())
        assert total_nnz == nnz, (total_nnz, nnz)
    ms_step = ms_total / args.steps
    value = m / (ms_step * 1e-3) / 1e6
    del out

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event timing)
    peak, peak_src = read_peaks()
    alg = algorithmic_bytes(m, nn, nnz)
    # kernels per timed span: a scan is reduce + scan-of-sums + downsweep
    launches = sum(n * (3 if k == 'scan' else 1) for k, (_, n) in prof.items())
    dom = max(prof, key=lambda k: prof[k][0]) if prof else None
    kernel_ms = {k: v[0] / max(1, v[1]) * (v[1] / args.steps) for k, v in prof.items()}
    roof = None
    if dom:
        dom_ms = prof[dom][0] / args.steps
        achieved = alg / world / (dom_ms * 1e-3) / 1e9
        roof = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak,
                'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': ncu_traffic(dom, level) if world == 1 else None,
                'peak_source': peak_src, 'kernel_ms_per_step': dom_ms,
                'algorithmic_bytes_per_launch': alg // world,
                'whole_step': {'achieved': alg / (ms_step * 1e-3) / 1e9,
                               'frac': alg / (ms_step * 1e-3) / 1e9 / (peak * world)},
                'all_kernels_ms_per_step': kernel_ms}

    # ---- end to end through the drop-in call: pinned host numpy in, scipy out
    e2e = None
    if world == 1 and not args.no_e2e:
        hc = torch.empty(c.shape, dtype=torch.float64, pin_memory=True)
        he = torch.empty(e.shape, dtype=torch.int64, pin_memory=True)   # what numpy users hold
        hc.copy_(c)
        he.copy_(e)
        cn, en = hc.numpy(), he.numpy()
        n_e2e = max(1, min(args.steps, 3))
        a = solvers.get_stiffness_matrix(cn, en)      # warm-up (pinned pool, pages)
        d2h = a.indptr.nbytes + a.indices.nbytes + a.data.nbytes
        del a
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            a = solvers.get_stiffness_matrix(cn, en)
            checksum = float(a.data[0])
            del a
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        e2e = {'value': m / dt / 1e6, 'unit': UNIT,
               'h2d_bytes_per_step': int(cn.nbytes + en.nbytes),
               'd2h_bytes_per_step': int(d2h), 'ms_per_step': dt * 1e3,
               'steps': n_e2e,
               'call': 'p1afempy_b200.solvers.get_stiffness_matrix(numpy, numpy) -> '
                       'scipy.sparse.csr_matrix'}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        rl = args.ref_level if args.ref_level < level else level
        rc, re_ = generate_mesh(rl)
        times = time_reference(rc, re_, 3, 1)
        model, ncpu, naff = cpu_info()
        cpu = {'value': re_.shape[0] / float(np.mean(times)) / 1e6, 'unit': UNIT,
               'cores': 1, 'kind': 'port',
               'sample': 'oracle port (bit-identical to the reference) of '
                         'get_stiffness_matrix(...).tocsr() on S(%d), M=%d, mean of 3 '
                         'runs after 1 warm-up; host: %s, %d cpus (%d usable)'
                         % (rl, re_.shape[0], model, ncpu, naff)}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'stiffness->CSR (cold: plan + fill), S(%d) unit square '
                               'uniformly refined by NVB, M=%d triangles, N=%d nodes, '
                               'nnz=%d' % (level, m, nn, nnz),
                   'parallelism': ('rows dealt to %d GPUs in chunks of 4096, inputs replicated, '
                                   'no data-path collective' % world) if world > 1 else '1 GPU',
                   'l2': 'inputs (%.0f MB) and outputs (%.0f MB) exceed the 126 MB L2; '
                         'no flush needed' % ((12 * m + 16 * nn) / 1e6, 12 * nnz / 1e6),
                   'mesh_generation_s': gen_s},
        'roofline': roof, 'cpu_baseline': cpu, 'e2e': e2e,
        'gpu_launches': launches, 'clocks': clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--level', type=int,
                    default=int(os.environ.get('P1_BENCH_LEVEL', '12')),
                    help='S(level): 9=1M, 11=16M, 12=64M triangles')
    ap.add_argument('--ref-level', type=int, default=10,
                    help='mesh of the bounded CPU sample (10 = 4M triangles)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true',
                    help='skip the host-buffer leg (256M: 25 GB of pinned host memory)')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
