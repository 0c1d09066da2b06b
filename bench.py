#!/usr/bin/env python
"""bench.py -- the P1 assembly hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--level L] [--op stiffness|mass|general|weighted|load|eta]

A "step" is one complete cold call of the operator on the mesh S(L) (unit square
cut by its centre, L uniform NVB refinements; L=12 is the 64M-triangle mesh the
north star quotes) from device-resident inputs to device-resident output, nothing
cached between steps.  The default operator is the BASELINE metric: stiffness
matrix to CSR (plan + fill, p1_assemble_cold: no host synchronisation inside a
step; the status words of every step are checked after the timed region).  The
other operators are the configurations C2 / C3 / C5 of SURVEY.md section 8(d),
with the coefficient values at the quadrature points array-fed from the device.
Prints ONE JSON line (see the contract in the task description).

N > 1 (torchrun): every rank holds the mesh and assembles its share of the rows
(chunks of 4096 rows dealt round-robin; strong scaling of one mesh); the only
collective is the all-gather of per-chunk nnz that turns the ranks' row pointers
into the global one, inside the timed region.  `value` = triangles of the mesh /
max-over-ranks device time.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

OPS = ('stiffness', 'mass', 'general', 'weighted', 'load', 'eta')
METRIC = {'stiffness': 'P1 stiffness assembly Melements/s to CSR',
          'mass': 'P1 mass assembly Melements/s to CSR',
          'general': 'P1 generalised stiffness assembly Melements/s to CSR',
          'weighted': 'P1 weighted mass assembly Melements/s to CSR',
          'load': 'P1 cubature load vector Melements/s',
          'eta': 'P1 compute_eta_r Melements/s'}
UNIT = 'Melements/s'


# ----------------------------------------------------------------- workload
def mesh_counts(level):
    n = 2 ** level
    return 4 * n * n, 2 * n * n + 2 * n + 1, 14 * n * n + 6 * n + 1


def algorithmic_bytes(op, m, nn, nnz, n_qp):
    """BASELINE.md section 2 / SURVEY.md section 8(d): every input read once, every
    output written once; int32 indices, fp64 values; coefficient values array-fed."""
    csr = 12 * m + 16 * nn + 12 * nnz + 4 * (nn + 1)
    return {'stiffness': csr, 'mass': csr,
            'general': csr + 32 * n_qp * m,
            'weighted': csr + 8 * nn + 8 * n_qp * m,
            'load': 12 * m + 16 * nn + 8 * nn + 8 * n_qp * m,
            'eta': 12 * m + 16 * nn + 8 * nn + 8 * m + 8 * m}[op]


def generate_mesh(level):
    """S(level) on the host, with its boundary edges."""
    from p1afempy_b200 import meshgen
    c, e, b = meshgen.criss_square()
    for _ in range(level):
        c, e, b = meshgen.refine_nvb(c, e, np.arange(e.shape[0]), b)
    return c, e, b


# the callbacks of C2 / C3; written so that they run on
# numpy arrays and on the package's device-resident ndarray stand-in alike
def a11(r): return -(1. + r[:, 0] ** 2)
def a12(r): return -0.25 * r[:, 0] * r[:, 1]
def a21(r): return -0.5 * r[:, 0] * r[:, 1]
def a22(r): return -(2. + np.sin(np.pi * r[:, 1]))
def phi(u): return 1. + u ** 2
def f_load(r): return np.sin(2. * np.pi * r[:, 0]) * np.cos(3. * np.pi * r[:, 1]) + 1.
def g_zero(r): return 0. * r[:, 0]
def u_nodal(c): return np.sin(np.pi * c[:, 0]) * np.sin(np.pi * c[:, 1])


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


def reference_call(op, c, e, b):
    """One call of the reference's CPU implementation of `op` (oracle port, pinned
    bit-identical to /root/reference: tests/test_oracle_vs_reference.py).  Single
    thread -- numpy/scipy do not thread this path."""
    from oracle import p1_oracle as orc
    from p1afempy_b200 import cubature
    w, p = cubature.resolve_rule(cubature.CubatureRuleEnum.CONICAL_GAUSS_4X4)
    if op == 'stiffness':
        return orc.stiffness_coo(c, e).tocsr()
    if op == 'mass':
        return orc.mass_coo(c, e).tocsr()
    if op == 'general':
        return orc.general_stiffness_coo(c, e, a11, a12, a21, a22, w, p).tocsr()
    if op == 'weighted':
        return orc.weighted_mass_coo(c, e, u_nodal(c), phi, w, p).tocsr()
    if op == 'load':
        return orc.load_vector_cubature(c, e, f_load, w, p)
    none = np.zeros((0, 2), dtype=np.int64)
    return orc.eta_r(u_nodal(c), c, e, np.vstack(b), none, f_load, g_zero)


def time_reference(op, level, steps, warmup):
    c, e, b = generate_mesh(level)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = reference_call(op, c, e, b)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        del out
    return e.shape[0], times


# seconds per triangle of the oracle port on one core (measured, rounded up): sizes
# the bounded samples
REF_S_PER_TRI = {'stiffness': 5e-7, 'mass': 4.5e-7, 'general': 1.3e-6, 'weighted': 1.3e-6,
                 'load': 1.5e-6, 'eta': 7e-7}
REF_PEAK_B_PER_TRI = {'stiffness': 520, 'mass': 520, 'general': 700, 'weighted': 700,
                      'load': 400, 'eta': 500}


def reference_level(op, level, runs, budget_s):
    """Largest S(l), l <= level, on which `runs` calls of the oracle port fit the time
    budget and the host's memory."""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    lowest = min(level, 6)
    for l in range(level, lowest - 1, -1):
        m = mesh_counts(l)[0]
        if m * REF_S_PER_TRI[op] * runs <= budget_s and m * REF_PEAK_B_PER_TRI[op] <= 0.6 * avail:
            return l
    return lowest


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU, polled through NVML from a thread
    for as long as the timed region lasts (a few thousand samples per second: also
    millisecond regions get their samples)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.stop_flag = threading.Event()
        self.recording = threading.Event()
        self.sm, self.reasons, self.sm_max = [], set(), None
        self.error = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown',
                                            getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8)),
                     'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown',
                                                    getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40)),
                     'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown',
                                                    getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20)),
                     'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap',
                                             getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4))}
            get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons',
                                  getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons', None))
            while not self.stop_flag.is_set():
                if self.recording.is_set():
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    if get_reasons is not None:
                        bits = int(get_reasons(h))
                        for name, bit in names.items():
                            if bits & bit:
                                self.reasons.add(name)
                else:
                    time.sleep(0.0005)
            nv.nvmlShutdown()
        except Exception as exc:  # no NVML: the driver's own clock record stands alone
            self.error = repr(exc)

    def sample_now(self):
        """One sample from the calling thread (regions too short for the poller)."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
        except Exception as exc:
            self.error = repr(exc)

    def result(self):
        self.stop_flag.set()
        self.join(timeout=5)
        out = {'sm_mhz': None, 'sm_max_mhz': self.sm_max, 'reasons': sorted(self.reasons),
               'samples': len(self.sm), 'source': 'NVML polled from a thread during the timed region'}
        if self.sm:
            out['sm_mhz'] = float(np.median(self.sm))
        if self.error:
            out['error'] = self.error
        return out


def read_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except (OSError, KeyError, ValueError):
        return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(level, op):
    """Per-kernel dram__bytes_read.sum + dram__bytes_write.sum of one launch from the
    committed `ncu --set full` captures (profiles/traffic.json), or {}."""
    try:
        with open(os.path.join(REPO, 'profiles', 'traffic.json')) as fh:
            table = json.load(fh)
        entry = table.get('level_%d' % level, {})
        if op != 'stiffness':
            entry = entry.get(op, {})
        return {k: int(v['dram_bytes']) for k, v in entry.items()
                if isinstance(v, dict) and 'dram_bytes' in v}
    except (OSError, ValueError, KeyError):
        return {}


# ---------------------------------------------------------------- reference arm
def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle
    port, pinned bit-identical to /root/reference) on the host cores, on the same
    mesh as the GPU arm when K + W calls of it fit a few minutes, else on the
    largest S(l) that does."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    warmup = min(args.warmup, 1)
    level = reference_level(args.op, args.level, args.steps + warmup, args.ref_budget)
    m, times = time_reference(args.op, level, args.steps, warmup)
    t = float(np.mean(times))
    model, ncpu, naff = cpu_info()
    value = m / t / 1e6
    same = level == args.level
    line = {
        'impl': 'reference', 'metric': METRIC[args.op], 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': warmup,
        'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': '%s, S(%d) unit square uniformly refined by NVB, M=%d triangles%s'
                               % (args.op, level, m, '' if same else
                                  ' (bounded sample of the S(%d) workload: %d calls of it '
                                  'would exceed the %d s budget)' % (args.level, args.steps + warmup,
                                                                     args.ref_budget)),
                   'same_mesh_as_gpu_arm': same},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                         'sample': 'oracle port of the reference\'s %s on S(%d), M=%d, mean of %d '
                                   'calls after %d warm-up; host: %s, %d cpus (%d usable), '
                                   'numpy/scipy single-threaded on this path'
                                   % (args.op, level, m, args.steps, warmup, model, ncpu, naff)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------- GPU arm
def parse_profile(text):
    out = {}
    for row in text.strip().splitlines():
        name, ms, launches = row.split()
        out[name] = (float(ms), int(launches))
    return out


class Workload:
    """Device-resident inputs of one operator and its cold step."""

    def __init__(self, op, level, world, rank, local_rank):
        import torch
        from p1afempy_b200 import _lib, cubature, meshgen
        from p1afempy_b200.device import DeviceMesh
        self.torch, self._lib, self.op = torch, _lib, op
        self.world, self.rank = world, rank
        t0 = time.perf_counter()
        c, e, bnd = meshgen.unit_square_device(level, device=local_rank)
        torch.cuda.synchronize()
        self.gen_s = time.perf_counter() - t0
        self.c, self.e, self.bnd = c, e, bnd
        self.m, self.nn = int(e.shape[0]), int(c.shape[0])
        self.w, self.p = cubature.resolve_rule(cubature.CubatureRuleEnum.CONICAL_GAUSS_4X4)
        self.n_qp = int(self.w.shape[0])
        # N > 1: matrices and load vectors are dealt out by rows (chunks of 4096), the
        # estimator by contiguous element blocks
        inter = (world, rank, 12) if world > 1 and op not in ('eta', 'load') else None
        self.block = (self.m * rank // world, self.m * (rank + 1) // world)
        if world > 1 and op == 'load':
            # the load vector is dealt out by ELEMENTS: the rank's block over all N nodes,
            # the ranks' partial vectors added by one all-reduce (dist.load_vector_reduced)
            self.mesh = DeviceMesh(c, e[self.block[0]:self.block[1]], device=local_rank,
                                   n_nodes=self.nn, check=False)
        else:
            self.mesh = DeviceMesh(c, e, device=local_rank, interleave=inter)
        self.statuses = []
        self.out = None        # the output arrays, overwritten by every step
        x, y = c[:, 0], c[:, 1]
        pi = float(np.pi)
        if op in ('general', 'load'):
            xq = self.mesh.quadrature_points(self.p)          # (n_qp, M, 2)
            qx, qy = xq[..., 0], xq[..., 1]
            if op == 'general':
                self.q = [(-(1. + qx ** 2)).contiguous(), (-0.25 * qx * qy).contiguous(),
                          (-0.5 * qx * qy).contiguous(),
                          (-(2. + torch.sin(pi * qy))).contiguous()]
                self.args = _lib.QuadArgs(self.q[0].data_ptr(), self.q[1].data_ptr(),
                                          self.q[2].data_ptr(), self.q[3].data_ptr(), None,
                                          self.w.ctypes.data, None, self.n_qp)
            else:
                self.fq = (torch.sin(2. * pi * qx) * torch.cos(3. * pi * qy) + 1.).contiguous()
            del xq, qx, qy
        if op in ('weighted', 'eta'):
            self.u = (torch.sin(pi * x) * torch.sin(pi * y)).contiguous()
        if op == 'weighted':
            uq = self.mesh.nodal_at_quadrature(self.u, self.p)
            self.phiq = (1. + uq ** 2).contiguous()
            del uq
            self.args = _lib.QuadArgs(None, None, None, None, self.phiq.data_ptr(),
                                      self.w.ctypes.data, self.p.ctypes.data, self.n_qp)
        if op == 'eta':
            cen = self.mesh.centroids()
            self.fc = (torch.sin(2. * pi * cen[:, 0]) * torch.cos(3. * pi * cen[:, 1]) + 1.
                       ).contiguous()
            from p1afempy_b200.device import to_device_indices
            self.dirichlet = to_device_indices(torch.cat([b.reshape(-1, 2) for b in bnd]),
                                               local_rank, self.mesh.ctx)
            self.none = torch.zeros((0, 2), dtype=torch.int32, device=c.device)
        torch.cuda.synchronize()

    def step(self):
        """One cold call; returns the output (kept alive by the caller until the next
        step so that the allocator behaves as in real use)."""
        _lib, mesh, op = self._lib, self.mesh, self.op
        if op in ('stiffness', 'mass', 'general', 'weighted'):
            code = {'stiffness': _lib.OP_STIFFNESS, 'mass': _lib.OP_MASS,
                    'general': _lib.OP_GENERAL, 'weighted': _lib.OP_WEIGHTED}[op]
            cold = self.out = mesh.assemble_cold(code, quad=getattr(self, 'args', None),
                                                 out=self.out)
            self.statuses.append(cold.status.clone())     # (8 words, device to device)
            if self.world > 1:   # where my chunks start in the global CSR: one all-gather
                from p1afempy_b200 import dist as pdist
                offs, total = pdist.global_chunk_offsets(cold.indptr, self.nn, self.world,
                                                         self.rank, 12)
                return cold, offs, total
            return cold
        if op == 'load':
            out = mesh.load_vector(self.fq, self.w, self.p)
            mesh.close()        # cold: nothing is kept for the next call
            if self.world > 1:
                import torch.distributed as dist
                dist.all_reduce(out)
            return out
        if self.world > 1:
            return mesh.eta_r_block(self.u, self.dirichlet, self.none, None, self.fc, *self.block)
        out = mesh.eta_r(self.u, self.dirichlet, self.none, None, self.fc)
        mesh.close()
        return out

    def check(self, nnz_expected):
        """After the timed region: the status words of every cold assembly."""
        for st in self.statuses:
            host = st.cpu().numpy()
            assert host[0] == 0, 'status bits %d: the cold path declined the mesh' % host[0]
            if self.world == 1:
                assert host[6] == nnz_expected, (host[6], nnz_expected)
        self.statuses.clear()


def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from p1afempy_b200 import _lib
    from p1afempy_b200.device import context

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a GPU (no CPU fallback); use --impl reference '
                         'for the CPU arm')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    op, level = args.op, args.level
    m, nn, nnz = mesh_counts(level)
    wl = Workload(op, level, world, rank, local_rank)
    assert wl.m == m and wl.nn == nn
    ctx, _ = context(local_rank)
    lib = _lib.load()
    stream = torch.cuda.current_stream()

    out = None
    for _ in range(args.warmup):
        out = wl.step()
    torch.cuda.synchronize()
    wl.check(nnz)
    if world > 1 and op in ('stiffness', 'mass', 'general', 'weighted'):
        assert int(out[2].item()) == nnz, (int(out[2].item()), nnz)
    if op == 'load':    # the entries of the load vector add up to the integral of f (= 1)
        assert out.shape[0] == nn and abs(float(out.sum().item()) - 1.) < 1e-6, float(out.sum())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.recording.set()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        out = wl.step()
    ev1.record(stream)
    if rank == 0:
        sampler.sample_now()         # (the GPU is still busy with the queued steps)
    torch.cuda.synchronize()
    sampler.recording.clear()
    if world > 1:
        dist.barrier()
    clocks = sampler.result() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    wl.check(nnz)
    # the kernels' shares of a step: the same steps once more with the library's event
    # profiler on (outside the timed region: it brackets every kernel with two events)
    lib.p1_profile_reset(ctx)
    lib.p1_profile_enable(ctx, 1)
    for _ in range(args.steps):
        out = wl.step()
    torch.cuda.synchronize()
    wl.check(nnz)
    buf = C.create_string_buffer(1 << 14)
    lib.p1_profile_report(ctx, buf, len(buf))
    lib.p1_profile_enable(ctx, 0)
    prof = parse_profile(buf.value.decode())
    if world > 1:
        tmax = torch.tensor([ms_total], dtype=torch.float64, device='cuda')
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
    ms_step = ms_total / args.steps
    value = m / (ms_step * 1e-3) / 1e6
    del out

    # ---- end to end through the drop-in call, host buffers in and out
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, wl, world, rank, local_rank, m)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline: the WHOLE step against the measured HBM peak; per kernel: its
    # share of the step and, where an ncu capture of this workload is committed, the
    # DRAM bytes it really moves
    peak, peak_src = read_peaks()
    alg = algorithmic_bytes(op, m, nn, nnz, wl.n_qp)
    # (counted in the profiled repeat of the same steps; "scan" and "rowptr" are three
    # kernels each: reduce / scan of the tile sums / final)
    launches = sum(n * (3 if k in ('scan', 'rowptr') else 1) for k, (_, n) in prof.items())
    if op in ('stiffness', 'mass', 'general', 'weighted'):
        launches += args.steps          # cold_init_kernel (not timed on its own)
    kernel_ms = {k: v[0] / args.steps for k, v in prof.items()}
    traffic = ncu_traffic(level, op) if world == 1 else {}
    dom = max(kernel_ms, key=kernel_ms.get) if kernel_ms else None
    achieved = alg / (ms_step * 1e-3) / 1e9
    kernels = {}
    for k, ms in kernel_ms.items():
        kernels[k] = {'ms_per_step': ms, 'share_of_step': ms / ms_step}
        if k in traffic:
            kernels[k]['dram_bytes'] = traffic[k]
            kernels[k]['dram_gbs'] = traffic[k] / (ms * 1e-3) / 1e9
            kernels[k]['dram_frac_of_peak'] = kernels[k]['dram_gbs'] / peak
    roof = {'bound': 'hbm', 'achieved': achieved, 'peak': peak * world, 'unit': 'GB/s',
            'frac': achieved / (peak * world),
            'traffic': sum(traffic.values()) if traffic else None,
            'what': 'algorithmic bytes of the whole step / step time (all kernels), against the '
                    'measured copy bandwidth%s' % (' x %d GPUs' % world if world > 1 else ''),
            'peak_source': peak_src, 'algorithmic_bytes_per_step': alg,
            'dominant_kernel': dom, 'kernels': kernels,
            'frac_of_nominal_8TBs': achieved / (8000.0 * world)}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        rl = reference_level(op, min(level, args.cpu_level), 3, 30.)
        rm, times = time_reference(op, rl, 2, 1)
        model, ncpu, naff = cpu_info()
        cpu = {'value': rm / float(np.mean(times)) / 1e6, 'unit': UNIT, 'cores': 1, 'kind': 'port',
               'sample': 'oracle port (bit-identical to the reference) of %s on S(%d), M=%d, mean '
                         'of 2 calls after 1 warm-up; host: %s, %d cpus (%d usable)'
                         % (op, rl, rm, model, ncpu, naff)}

    line = {
        'metric': METRIC[op], 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': '%s (cold call, nothing cached between steps), S(%d) unit square '
                               'uniformly refined by NVB, M=%d triangles, N=%d nodes, nnz=%d%s'
                               % (op, level, m, nn, nnz,
                                  ', %d-point rule, coefficient values array-fed' % wl.n_qp
                                  if op in ('general', 'weighted', 'load') else ''),
                   'parallelism': (('contiguous element blocks on %d GPUs, inputs replicated, no '
                                    'collective' % world) if op == 'eta' else
                                   ('contiguous element blocks on %d GPUs (f(x_q) of the block only), '
                                    'one all-reduce of the partial vectors per step: the whole '
                                    'vector on every rank' % world) if op == 'load' else
                                   ('rows dealt to %d GPUs in chunks of 4096, inputs replicated; one '
                                    'all-gather of per-chunk nnz (global row pointer) per step'
                                    % world)) if world > 1 else '1 GPU',
                   'l2': ('inputs (%.0f MB) and outputs (%.0f MB, overwritten by every step) '
                          'exceed the 126 MB L2; no flush needed' if 12 * m + 16 * nn > 252e6 else
                          'inputs (%.0f MB) and outputs (%.0f MB) do NOT exceed the L2 by a safe '
                          'margin and nothing is flushed: not a headline configuration')
                         % ((12 * m + 16 * nn) / 1e6, (4 * nn + 12 * nnz) / 1e6),
                   'mesh_generation_s': wl.gen_s},
        'roofline': roof, 'cpu_baseline': cpu, 'e2e': e2e,
        'gpu_launches': launches, 'clocks': clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, wl, world, rank, local_rank, m):
    """The same metric through the public drop-in call with HOST buffers: numpy in,
    scipy / numpy out, host<->device copies inside the timed region.  Measured with
    pageable numpy arrays (what a caller of the reference holds) and with pinned
    ones.  N > 1: every rank uploads the mesh and downloads its share of the rows
    (p1afempy_b200.dist.DistributedMesh); max over ranks."""
    import torch
    import torch.distributed as dist
    from p1afempy_b200 import cubature, indicators, solvers
    from p1afempy_b200.device import to_host
    op = args.op
    rule = cubature.CubatureRuleEnum.CONICAL_GAUSS_4X4
    n_e2e = max(1, min(args.steps, args.e2e_steps))
    c, e = wl.c, wl.e

    def host_copy(pinned):
        hc = torch.empty(c.shape, dtype=torch.float64, pin_memory=pinned)
        he = torch.empty(e.shape, dtype=torch.int64, pin_memory=pinned)   # what numpy users hold
        hc.copy_(c)
        he.copy_(e)
        return hc.numpy(), he.numpy()

    def call(cn, en):
        if world > 1:
            from p1afempy_b200.dist import DistributedMesh
            dm = DistributedMesh(cn, en, device=local_rank)
            ip, ix, data = dm.stiffness_csr() if op == 'stiffness' else dm.mass_csr()
            out = (to_host(ip), to_host(ix), to_host(data))
            dm.close()
            return out, sum(t.nbytes for t in out)
        if op == 'stiffness':
            a = solvers.get_stiffness_matrix(cn, en)
        elif op == 'mass':
            a = solvers.get_mass_matrix(cn, en)
        elif op == 'general':
            a = solvers.get_general_stiffness_matrix(cn, en, a11, a12, a21, a22, rule)
        elif op == 'weighted':
            a = solvers.get_weighted_mass_matrix(cn, en, un, phi, rule)
        elif op == 'load':
            a = solvers.get_right_hand_side(cn, en, f_load, cubature_rule=rule)
            return a, a.nbytes
        else:
            a = indicators.compute_eta_r(un, cn, en, dn, none, f_load, g_zero)
            return a, a.nbytes
        return a, a.indptr.nbytes + a.indices.nbytes + a.data.nbytes

    if world > 1 and op not in ('stiffness', 'mass'):
        return None
    un = dn = none = None
    if op in ('weighted', 'eta'):
        un = to_host(wl.u)
    if op == 'eta':
        dn = to_host(wl.dirichlet).astype(np.int64)
        none = np.zeros((0, 2), dtype=np.int64)
    result = {}
    for kind in ('pageable', 'pinned'):
        cn, en = host_copy(kind == 'pinned')
        out, d2h = call(cn, en)                      # warm-up (staging buffers, pages)
        del out
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            out, d2h = call(cn, en)
            del out
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            tmax = torch.tensor([dt], dtype=torch.float64, device='cuda')
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dt = float(tmax.item())
        h2d = int(cn.nbytes + en.nbytes + (un.nbytes if un is not None else 0))
        result[kind] = {'value': m / dt / 1e6, 'ms_per_step': dt * 1e3,
                        'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': int(d2h)}
        del cn, en
    head = result['pageable']
    names = {'stiffness': 'solvers.get_stiffness_matrix', 'mass': 'solvers.get_mass_matrix',
             'general': 'solvers.get_general_stiffness_matrix',
             'weighted': 'solvers.get_weighted_mass_matrix',
             'load': 'solvers.get_right_hand_side', 'eta': 'indicators.compute_eta_r'}
    return {'value': head['value'], 'unit': UNIT,
            'h2d_bytes_per_step': head['h2d_bytes_per_step'],
            'd2h_bytes_per_step': head['d2h_bytes_per_step'],
            'ms_per_step': head['ms_per_step'], 'steps': n_e2e,
            'host_memory': 'pageable numpy arrays (headline); pinned alongside',
            'pinned': result['pinned'],
            'call': ('p1afempy_b200.%s(numpy, ...) -> scipy.sparse.csr_matrix / ndarray'
                     % names[op]) if world == 1 else
                    'p1afempy_b200.dist.DistributedMesh(numpy, numpy).%s_csr() -> numpy pieces, '
                    'per rank: whole mesh up, own rows down' % op}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--op', default='stiffness', choices=OPS)
    ap.add_argument('--level', type=int,
                    default=int(os.environ.get('P1_BENCH_LEVEL', '12')),
                    help='S(level): 9=1M, 11=16M, 12=64M triangles')
    ap.add_argument('--cpu-level', type=int, default=11,
                    help='largest mesh of the bounded cpu_baseline sample (11 = 16M triangles)')
    ap.add_argument('--ref-budget', type=int, default=300,
                    help='--impl reference: seconds all K + W calls may take')
    ap.add_argument('--e2e-steps', type=int, default=10)
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
