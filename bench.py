#!/usr/bin/env python
"""bench.py -- CG-iterations/sec of the UniRes y-update on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload sr3_256] [--cg-iters 20] [--no-cpu-baseline]

One *step* = one pass of the y-update of `_update_admm` (unires/_update.py:122-150) over one
subject: for every channel build the right-hand side (sum tau At x - lam div(w - rho z)) and run
the CG solve with a FIXED trip count (tolerance 0, SURVEY.md 8d "throughput mode": no stop test,
36 N bytes per CG iteration: fused matvec 24 + residual update 12).  Workload at N=1: BASELINE.json configs[1], "3-channel 1 mm
super-resolution, 256^3 recon grid" (synthetic BrainWeb-like phantom, 181x217x181 scanner FOV,
each channel thick-sliced x4 along a different axis).  N>1: one independent subject per GPU
(weak scaling, no data-path collective).  value = CG iterations of all ranks / max-over-ranks
device time.  `e2e` is the same step through the public Python API with HOST (pinned) buffers:
observations uploaded, the initial estimate formed from them on the device, reconstructed
channels downloaded, every step.  `sharded` (every N): BASELINE.json configs[3], 8 channels of
384^3 sharded over the ranks -- one full ADMM iteration through `_update_admm_sharded` with its
NCCL all-reduces (strong scaling).  `energy_rule`: the same y-update in the reference's default
mode (stop='max_gain' = energy objective, tolerance 1e-3).
"""
import argparse
import json
import math
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'cg_iterations_per_sec'
UNIT = 'CG-it/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='sr3_256')
    ap.add_argument('--cg-iters', type=int, default=20)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-sample-iters', type=int, default=10)
    ap.add_argument('--channel-streams', type=int, default=None)
    ap.add_argument('--tune', action='append', default=[], help='knob=value (ur_tune)')
    ap.add_argument('--no-sharded', action='store_true',
                    help='skip the channel-sharded ADMM section (configs[3])')
    ap.add_argument('--sharded-workload', default='thickz2_384x8')
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


WORKLOADS = {
    'sr3_256': 'thick-slice super-resolution',
    'sr3_48': 'thick-slice super-resolution',
    'sr3_256_rigid': 'thick-slice super-resolution, rigidly mis-aligned scans (rotated operators)',
    'crop3_256': '1 mm observations on a larger 1 mm grid (crop / embed operator)',
    'thickz2_256': 'thick-slice super-resolution (z x2)',
    'thickz2_384': 'thick-slice super-resolution (z x2)',
    'thickz2_384x8': 'thick-slice super-resolution (z x2)',
    'iso2_512': '0.5 mm reconstruction of 1 mm isotropic data (ratio 2 on every axis)',
    'denoise_181': 'denoising (identity operator)',
}


def describe(workload):
    return WORKLOADS.get(workload, 'thick-slice super-resolution')


def matvec_kernel_desc(path):
    """What served the CG matvec (ur_last_lhs_path of the last launch)."""
    return {
        2: 'lhs_fast_kernel: CG matvec A p = sum tau AtA p + rho lam^2 DtD p with p = beta p + r, '
           'x += alpha p and p.Ap fused in (first iteration of a solve: plain matvec)',
        3: 'rotated operator: rot_forward_kernel (tile pull + slice profile + scaling + transposed '
           'profile in shared memory) + lhs_rot_kernel (gather adjoint + DtD + p.Ap); the launch '
           'pair is timed as one matvec; instruction-issue bound, not HBM bound (DESIGN.md 4.3)',
        5: 'rotated operator: rot_forward_kernel (tile pull + slice profile + scaling + transposed '
           'profile in shared memory) + rot_adjoint_cell_kernel (adjoint pull through per-cell '
           'corner coefficients in shared memory, deterministic) + lhs_fast_kernel (DtD + p.Ap with '
           'the adjoint as accumulator); the three launches are timed as one matvec; '
           'instruction-issue bound, not HBM bound (DESIGN.md 4.3)',
        4: 'several decimated axes: nd_down_spec_kernel (v -> low-resolution image) + '
           'nd_up_spec_kernel (expansion + DtD + p.Ap); the launch pair is timed as one matvec',
        1: 'lhs_stream_kernel (generic TMA streaming kernel)',
        0: 'lhs_direct_kernel (+ general-path accumulation)',
    }.get(int(path), 'lhs kernel path %d' % int(path))


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    COMMITTED `ncu --set full` capture of this workload (profiles/traffic.json), else None.  It
    is not measured in this run (ncu replays kernels; nothing under a profiler is a bench value):
    the JSON line says so in roofline.traffic_source."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return float(json.load(f)[workload]['dram_bytes_per_launch'])
    except Exception:
        return None


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {'hw_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8),
                 'hw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40),
                 'sw_thermal_slowdown': getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20),
                 'sw_power_cap': getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def result(self):
        self.stop_flag = True
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.sm_max,
                'reasons': sorted(self.reasons), 'samples': len(s)}


# ----------------------------------------------------------------------------- host placement
def bind_to_gpu_numa_node(index):
    """Pin this rank's threads to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned
    host buffers are allocated (first touch places them there).  Returns a description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(':')[0]) == 8:
            bdf = bdf[4:]
        with open('/sys/bus/pci/devices/%s/numa_node' % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return {'numa_node': None, 'bound': False}
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            cpus = set()
            for part in f.read().strip().split(','):
                a, _, b = part.partition('-')
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {'numa_node': node, 'bound': bool(allowed), 'cpus': len(allowed)}
    except Exception as e:  # placement is best effort
        return {'numa_node': None, 'bound': False, 'why': repr(e)[:80]}


def copy_ceiling(dev, h2d_bytes, d2h_bytes, steps, barrier):
    """its/s a step could reach if the host<->device copies of e2e were the ONLY cost: the same
    bytes per step, pinned buffers, H2D and D2H on two streams, all ranks at once."""
    src = torch.empty(max(h2d_bytes, 4) // 4, dtype=torch.float32).pin_memory()
    dst = torch.empty(max(d2h_bytes, 4) // 4, dtype=torch.float32).pin_memory()
    a = torch.empty_like(src, device=dev)
    b = torch.empty_like(dst, device=dev)
    s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    def once():
        with torch.cuda.stream(s_up):
            a.copy_(src, non_blocking=True)
        with torch.cuda.stream(s_dn):
            dst.copy_(b, non_blocking=True)
    once()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        once()
    torch.cuda.current_stream().wait_stream(s_up)
    torch.cuda.current_stream().wait_stream(s_dn)
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps


# ----------------------------------------------------------------------------- workload
def build_scenario(workload, device):
    from unires_b200 import synth, _project, struct
    cfg = synth.CONFIGS[workload]
    return synth.make_scenario(cfg, _project, struct, device=device, seed=0)


def y_update(x, y, z, w, rho, tmp, sett, vx, dim):
    """The y-update of one subject through the product's public functions."""
    from unires_b200 import _update
    return _update._solve_y(x, y, z, w, rho, tmp, sett, dim, vx)


def sharded_admm(args, dev, world, rank, barrier):
    """BASELINE.json configs[3]: 8 channels of 384^3 (z x2 thick slices), channels sharded
    round-robin over the ranks.  Step = ONE full ADMM iteration through the product's
    `_update_admm_sharded` (unires/_update.py:105-195): per-channel right-hand side + CG (20 fixed
    iterations), objective, JTV prox -- with its collectives: ONE SUM all-reduce of the prior-energy
    field and the JTV coupling field (2 x 226 MB float32 in one buffer) + one float64 scalar.
    Strong scaling: the same 8-channel problem on 1, 2, 4 or 8 GPUs."""
    import torch.distributed as dist
    from unires_b200 import synth, _project, struct, _update, parallel
    cfg = synth.CONFIGS[args.sharded_workload]
    C = len(cfg['thick'])
    if world > C:
        return {'skipped': '%d ranks for %d channels' % (world, C)}
    mine = parallel.channel_shard(C, world, rank)
    sc = synth.make_scenario(cfg, _project, struct, device=dev, seed=0, channels=mine,
                             phantom_device=dev)
    sett = sc.sett
    sett.cgs_max_iter, sett.cgs_tol = args.cg_iters, 0.0
    rho = sc.rho.clone()
    if world > 1:
        dist.broadcast(rho, 0)
    dim = tuple(sc.y[0].dim)
    n_vox = dim[0] * dim[1] * dim[2]
    z, w = _update._admm_aux(sc.y, sett)
    tmp = torch.zeros(dim, device=dev)
    n_steps = max(2, min(args.steps, 3))
    obj = torch.zeros(n_steps + 2, 3, dtype=torch.float64, device=dev)
    for it in range(2):  # warm-up (also moves z, w away from zero)
        _update._update_admm_sharded(sc.x, sc.y, z, w, rho, tmp, obj, it, sett)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(n_steps):
        _update._update_admm_sharded(sc.x, sc.y, z, w, rho, tmp, obj, 2 + it, sett)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / n_steps
    ar_ms = 0.0
    if world > 1:  # the iteration's (2, X, Y, Z) field all-reduce, in isolation
        field = torch.zeros((2,) + dim, device=dev)
        dist.all_reduce(field)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            dist.all_reduce(field)
        a1.record()
        barrier()
        ar_ms = a0.elapsed_time(a1) / 5
    t = torch.tensor([ms, ar_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ar_ms = t.tolist()
    finite = bool(torch.isfinite(obj[2:2 + n_steps]).all().item())
    return {'workload': '%s: %d-channel %s, recon grid %s, channels sharded round-robin over '
                        '%d rank(s); step = one full ADMM iteration (y-update with %d fixed CG '
                        'iterations per channel, objective, JTV prox) through _update_admm_sharded'
                        % (args.sharded_workload, C, describe(args.sharded_workload),
                           'x'.join(map(str, dim)), world, args.cg_iters),
            'scaling': 'strong', 'n_gpus': world, 'channels': C, 'channels_per_rank': len(mine),
            'ms_per_admm_iteration': ms, 'steps': n_steps,
            'value': C * args.cg_iters / (ms * 1e-3), 'unit': UNIT,
            'collectives_per_iteration': 0 if world == 1 else
            '1 x all-reduce(SUM) of 2 x %.0f MB float32 (prior-energy field + JTV coupling field in '
            'one buffer) + 1 float64 scalar (NCCL)' % (n_vox * 4 / 1e6),
            'allreduce_ms': ar_ms,
            'allreduce_busbw_gbs': (2 * (world - 1) / world * 2 * n_vox * 4 / (ar_ms * 1e-3) / 1e9)
            if ar_ms > 0 else None,
            'objective_finite': finite}


def run_ours(args):
    import torch.distributed as dist
    from unires_b200 import _lib, _update
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    sc = build_scenario(args.workload, dev)
    sett = sc.sett
    sett.cgs_max_iter, sett.cgs_tol = args.cg_iters, 0.0  # throughput mode: fixed trip count
    if args.channel_streams is not None:
        sett.channel_streams = args.channel_streams
    for kv in args.tune:
        k, v = kv.split('=')
        _lib.check(_lib.lib.ur_tune(k.encode(), int(v)))
    C = len(sc.x)
    dim, vx = _update._geometry(sc.y)
    n_vox = dim[0] * dim[1] * dim[2]
    rho = float(sc.rho)
    z, w = _update._admm_aux(sc.y, sett)
    tmp = torch.zeros(dim, device=dev)
    y0 = [yc.dat.clone() for yc in sc.y]
    its_per_step = C * args.cg_iters

    def reset():
        for c in range(C):
            sc.y[c].dat.copy_(y0[c])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-resident arm: inputs already in HBM ----
    for _ in range(max(args.warmup, 3)):
        reset()
        y_update(sc.x, sc.y, z, w, rho, tmp, sett, vx, dim)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.lib.ur_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        reset()  # restart from the same initial estimate (3 device copies, <1% of a step)
        y_update(sc.x, sc.y, z, w, rho, tmp, sett, vx, dim)
    host_ms = (time.perf_counter() - t_host0) * 1e3  # host enqueue time (no sync inside)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.lib.ur_launch_count() - l0

    # ---- roofline pass: the same K steps again with every matvec launch bracketed by CUDA
    # events on its stream (library instrumentation).  Channels run back to back on one stream
    # here so that a bracket holds exactly one kernel; the event records cost ~4 % of a step,
    # which is why `value` is timed without them. ----
    import ctypes as C_
    streams_cfg = getattr(sett, 'channel_streams', 1)
    sett.channel_streams = 1
    _lib.lib.ur_profile_matvec(1)
    for _ in range(args.steps):
        reset()
        y_update(sc.x, sc.y, z, w, rho, tmp, sett, vx, dim)
    tot, cnt, bpv = C_.c_double(0), C_.c_int32(0), C_.c_double(0)
    _lib.check(_lib.lib.ur_profile_matvec_read(C_.byref(tot), C_.byref(cnt), C_.byref(bpv)))
    _lib.lib.ur_profile_matvec(0)
    mv_path = _lib.lib.ur_last_lhs_path()
    sett.channel_streams = streams_cfg
    clocks = sampler.result()  # sampled over the timed region and the roofline pass (same load)

    # ---- end-to-end arm: host buffers in, host buffers out, every step ----
    # Per step: the observations of a subject are uploaded from pinned host memory, the initial
    # estimate is formed FROM THEM on the device (back-projection normalised by the operator's
    # column sums, what synth.make_scenario does on the host; round 1 uploaded it: 201 MB more
    # per step; one pass per channel through ur_backproject when the observations are lattice
    # aligned), every channel is solved and the reconstruction is downloaded.
    numa = bind_to_gpu_numa_node(local)
    hx = [[o.dat.cpu().pin_memory() for o in xc] for xc in sc.x]
    hy = [torch.empty(dim, dtype=torch.float32).pin_memory() for _ in range(C)]
    h2d = sum(t.numel() * 4 for xc in hx for t in xc)
    d2h = sum(t.numel() * 4 for t in hy)
    from unires_b200 import _project
    den = None
    if sett.do_proj:  # A' 1: a property of the operator, computed once
        den = [_project._proj_apply('At', torch.ones_like(sc.x[c][0].dat)[None, None],
                                    sc.x[c][0].po, method=sett.method)[0, 0].clamp_min(1e-3)
               for c in range(C)]

    inv_den, bp_ops = None, None
    if den is not None:  # one-pass back-projection (ur_backproject) for lattice observations
        inv_den = [1.0 / d for d in den]
        bp_ops = [_project.LhsOperator(sc.x[c], sc.y[c], method=sett.method, do=sett.do_proj,
                                       rho=rho, vx_y=vx) for c in range(C)]

    def init_y(x, y, c):
        if den is None:
            y[c].dat.copy_(x[c][0].dat)
        elif not _update._backproject(x[c], y[c].dat, bp_ops[c], inv_den[c]):
            num = _project._proj_apply('At', x[c][0].dat[None, None], x[c][0].po,
                                       method=sett.method)[0, 0]
            torch.div(num, den[c], out=y[c].dat)

    def clone_set(x, y):  # same operators (read-only), own observation / estimate volumes
        import copy
        xb = []
        for xc in x:
            row = []
            for o in xc:
                n = copy.copy(o)
                n.dat = o.dat.clone()
                row.append(n)
            xb.append(row)
        yb = []
        for yc in y:
            n = copy.copy(yc)
            n.dat = yc.dat.clone()
            yb.append(n)
        return xb, yb

    set_b = clone_set(sc.x, sc.y)
    pipe = _update.HostPipeline([(sc.x, sc.y), set_b], z, w, rho, sett, init_y=init_y)

    def e2e_step():
        pipe.submit(hx, None, hy)

    for _ in range(2):
        e2e_step()
    pipe.drain()
    barrier()
    l_e0 = _lib.lib.ur_launch_count()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    pipe.drain()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    e2e_launches = _lib.lib.ur_launch_count() - l_e0
    e2e_ok = all(torch.isfinite(t).all().item() for t in hy)
    ms_copy = copy_ceiling(dev, h2d, d2h, args.steps, barrier)

    # ---- the reference's default mode: energy stop rule, tolerance 1e-3 (device-side stop) ----
    sett.cgs_tol = 1e-3
    for _ in range(2):
        reset()
        y_update(sc.x, sc.y, z, w, rho, tmp, sett, vx, dim)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    infos = []
    for _ in range(args.steps):
        reset()
        infos.append(y_update(sc.x, sc.y, z, w, rho, tmp, sett, vx, dim))
    g1.record()
    barrier()
    ms_energy = g0.elapsed_time(g1)
    its_energy = sum(i.n_iter for step in infos for i in step)
    sett.cgs_tol = 0.0

    sharded = None
    if not args.no_sharded:
        try:
            sharded = sharded_admm(args, dev, world, rank, barrier)
        except Exception as e:  # never take the headline down
            sharded = {'error': repr(e)[:200]}

    t = torch.tensor([ms, ms_e2e, ms_copy, ms_energy], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_copy, ms_energy = t.tolist()
    t2 = torch.tensor([float(its_energy)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2)
    its_energy_all = t2.item()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_its = world * its_per_step * args.steps
    peak, peak_src = peaks()
    mv_ms = tot.value / max(cnt.value, 1)
    # algorithmic bytes per voxel of the average matvec launch: 8 (read p, write Ap) for a plain
    # launch, 24 when the direction and x updates are fused in (read p, r, x; write p, Ap, x)
    mv_bpv = bpv.value / max(cnt.value, 1)
    achieved = mv_bpv * n_vox / (mv_ms * 1e-3) / 1e9 if cnt.value else None
    line = {
        'metric': METRIC, 'value': total_its / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': '%s: %d-channel %s, recon grid %s, '
                               'CG y-update with %d fixed iterations per channel (tolerance 0); '
                               'one subject per GPU' % (args.workload, C, describe(args.workload),
                                                        'x'.join(map(str, dim)), args.cg_iters),
                   'channels': C, 'recon_grid': list(dim), 'cg_iters_per_channel': args.cg_iters,
                   'channel_streams': int(getattr(sett, 'channel_streams', 1)),
                   'l2': 'inputs larger than L2: CG working set per channel 5 volumes = %.0f MB '
                         '(L2 126 MB); no explicit flush' % (5 * n_vox * 4 / 1e6)},
        'e2e': {'value': total_its / (ms_e2e * 1e-3), 'unit': UNIT,
                'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'pipeline': 'HostPipeline: 2 device buffer sets, subject k+1 uploads / k-1 '
                            'downloads while k solves; observations uploaded, initial estimate '
                            'formed from them on the device, reconstruction downloaded',
                'result_finite': bool(e2e_ok), 'gpu_launches': int(e2e_launches),
                'host_placement': numa,
                'copy_ceiling': {'value': world * its_per_step / (ms_copy * 1e-3), 'unit': UNIT,
                                 'ms_per_step': ms_copy,
                                 'what': 'the same H2D + D2H bytes per step as plain pinned copies '
                                         'on two streams, all ranks at once, no kernels: the '
                                         'host-side ceiling of e2e on this box'}},
        'energy_rule': {'value': its_energy_all / (ms_energy * 1e-3), 'unit': UNIT,
                        'ms_per_step': ms_energy / args.steps,
                        'cg_iterations_per_step': its_energy_all / args.steps / world,
                        'what': "the same y-update in the reference's default mode: "
                                "stop='max_gain' (energy objective), tolerance 1e-3, at most 20 "
                                'iterations; iterations counted from the device-side stop'},
        'sharded': sharded,
        'gpu_launches': int(launches),
        'host_enqueue_ms_per_step': host_ms / args.steps,
        'clocks': clocks,
        'roofline': {'bound': 'hbm', 'kernel': matvec_kernel_desc(mv_path),
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': (achieved / peak) if achieved else None,
                     'traffic': ncu_traffic(args.workload),
                     'traffic_source': 'committed ncu --set full capture (profiles/traffic.json), '
                                       'not measured in this run',
                     'timing': 'CUDA events around every matvec launch in a second pass over the '
                               'same K steps (channels serialised on one stream)',
                     'algorithmic_bytes_per_launch': mv_bpv * n_vox,
                     'algorithmic_bytes_per_voxel': mv_bpv,
                     'avg_launch_ms': mv_ms, 'launches_timed': cnt.value, 'peak_source': peak_src},
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            line['cpu_baseline'] = cpu_baseline(sc, args)
        except Exception as e:  # the baseline must never take the GPU number down with it
            line['cpu_baseline'] = {'error': repr(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- CPU arms
def oracle_problem(workload, channel=0):
    """Channel `channel` of the workload built ENTIRELY with the CPU oracle (no kernels)."""
    from unires_b200 import synth
    from oracle.adapters import port_ops, port_structs
    cfg = synth.CONFIGS[workload]
    return synth.make_scenario(cfg, port_ops, port_structs, device='cpu', channels=[channel])


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1 for every rank: the CPU arms must still use every
    host core this process may run on (the reference would)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    if torch.get_num_threads() < n:
        torch.set_num_threads(n)
    return torch.get_num_threads()


def time_oracle_cg(sc, n_iters, warm=1):
    """Seconds per CG iteration of the oracle port (first channel of `sc`, tolerance 0), after
    `warm` untimed iterations."""
    use_all_host_cores()
    from oracle import unires_port as P
    from oracle.nitorch_shim.core import optim as OO
    vx = torch.ones(3) * float(sc.cfg['vx_y'])
    kw = dict(method=sc.sett.method, do=sc.sett.do_proj)
    b = sc.x[0][0].tau * P.proj('At', sc.x[0][0].dat, sc.x[0], sc.y[0], n=0, **kw)
    lhs = lambda v: P.proj('AtA', v, sc.x[0], sc.y[0], rho=sc.rho, vx_y=vx, **kw)
    stamps = []
    OO.cg(A=lhs, b=b, x=sc.y[0].dat.clone(), max_iter=warm + n_iters, tolerance=0, stop='max_gain',
          record=lambda it, xi: stamps.append(time.perf_counter()))
    return (stamps[-1] - stamps[warm - 1]) / n_iters


def cpu_baseline(sc_gpu, args):
    """Oracle port on the host cores, bounded sample: channel 0, a few CG iterations."""
    from unires_b200 import synth
    from oracle.adapters import port_ops, port_structs
    sc = synth.make_scenario(sc_gpu.cfg, port_ops, port_structs, device='cpu', channels=[0],
                             truth=[sc_gpu.truth[0].cpu()])
    sec = time_oracle_cg(sc, args.cpu_sample_iters)
    return {'value': 1.0 / sec, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': 'oracle/unires_port.py (restated reference, pure-PyTorch primitives) on CPU: '
                      'channel 0 of %s, %d CG iterations at full size (tolerance 0) after 1 '
                      'untimed one, %.1f s per iteration' % (args.workload, args.cpu_sample_iters,
                                                             sec)}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  nitorch is not
    installable (no network, not vendored), so this times the oracle port -- the reference's
    control flow (bitwise equal to its own files, tests/test_oracle_vs_reference.py) over the
    restated nitorch primitives -- on all host cores, on the SAME workload: a step is a bounded
    sample of the y-update, ONE full-size CG iteration of EVERY channel; `--warmup` untimed
    steps, then `--steps` timed ones."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from unires_b200 import synth
    C = len(synth.CONFIGS[args.workload]['thick'])
    n, warm = max(1, args.steps), max(1, args.warmup)
    sec, dim = 0.0, None
    for c in range(C):
        sc = oracle_problem(args.workload, c)
        dim = tuple(sc.y[0].dim)
        sec += time_oracle_cg(sc, n, warm=warm)  # seconds per iteration of this channel
        del sc
    val = C / sec  # CG iterations per second over a step of C iterations
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT,
            'n_gpus': int(os.environ.get('WORLD_SIZE', '1')), 'steps': n, 'warmup': warm,
            'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': '%s: %d-channel %s, recon grid %s, CG y-update (tolerance 0); '
                                   'bounded sample: each step is ONE full-size CG iteration of '
                                   'every channel on the host cores'
                                   % (args.workload, C, describe(args.workload),
                                      'x'.join(map(str, dim))),
                       'channels': C, 'recon_grid': list(dim), 'cg_iters_per_channel': 1},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': torch.get_num_threads(),
                             'kind': 'port',
                             'sample': '%d timed steps (one full-size CG iteration of each of the '
                                       '%d channels) after %d untimed ones' % (n, C, warm)},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
