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
observations and initial estimate uploaded, reconstructed channels downloaded, every step.
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
    'crop3_256': '1 mm observations on a larger 1 mm grid (crop / embed operator)',
    'thickz2_256': 'thick-slice super-resolution (z x2)',
    'thickz2_384': 'thick-slice super-resolution (z x2)',
    'thickz2_384x8': 'thick-slice super-resolution (z x2)',
    'iso2_512': '0.5 mm reconstruction of 1 mm isotropic data (ratio 2 on every axis)',
    'denoise_181': 'denoising (identity operator)',
}


def describe(workload):
    return WORKLOADS.get(workload, 'thick-slice super-resolution')


def ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed `ncu --set full` capture of this workload (profiles/traffic.json), else None."""
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
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.sm_max,
                'reasons': sorted(self.reasons), 'samples': len(s)}


# ----------------------------------------------------------------------------- workload
def build_scenario(workload, device):
    from unires_b200 import synth, _project, struct
    cfg = synth.CONFIGS[workload]
    return synth.make_scenario(cfg, _project, struct, device=device, seed=0)


def y_update(x, y, z, w, rho, tmp, sett, vx, dim):
    """The y-update of one subject through the product's public functions."""
    from unires_b200 import _update
    return _update._solve_y(x, y, z, w, rho, tmp, sett, dim, vx)


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
    clocks = sampler.result()

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
    sett.channel_streams = streams_cfg

    # ---- end-to-end arm: host buffers in, host buffers out, every step ----
    hx = [[o.dat.cpu().pin_memory() for o in xc] for xc in sc.x]
    hy0 = [t.cpu().pin_memory() for t in y0]
    hy = [torch.empty_like(t).pin_memory() for t in hy0]
    h2d = sum(t.numel() * 4 for xc in hx for t in xc) + sum(t.numel() * 4 for t in hy0)
    d2h = sum(t.numel() * 4 for t in hy)

    # public host-buffer entry point: a double-buffered pipeline over a stream of subjects
    # (uploads of subject k+1 and the download of subject k-1 overlap the solves of subject k;
    # every byte of every step still crosses PCIe inside the timed region)
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
    pipe = _update.HostPipeline([(sc.x, sc.y), set_b], z, w, rho, sett)

    def e2e_step():
        pipe.submit(hx, hy0, hy)

    for _ in range(2):
        e2e_step()
    pipe.drain()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    pipe.drain()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    e2e_ok = all(torch.isfinite(t).all().item() for t in hy)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
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
                            'downloads while k solves', 'result_finite': bool(e2e_ok)},
        'gpu_launches': int(launches),
        'host_enqueue_ms_per_step': host_ms / args.steps,
        'clocks': clocks,
        'roofline': {'bound': 'hbm', 'kernel': 'lhs_fast_kernel: CG matvec A p = sum tau AtA p + rho lam^2 '
                                               'DtD p with p = beta p + r, x += alpha p and p.Ap '
                                               'fused in (first iteration of a solve: plain matvec)',
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': (achieved / peak) if achieved else None,
                     'traffic': ncu_traffic(args.workload),
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
    cfg = dict(synth.CONFIGS[workload])
    cfg['thick'] = [cfg['thick'][channel]]
    truth = synth.phantom(tuple(cfg['dim_y']), 1, 0)
    return synth.make_scenario(cfg, port_ops, port_structs, device='cpu', truth=truth)


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


def time_oracle_cg(sc, n_iters):
    """Seconds per CG iteration of the oracle port (channel 0, tolerance 0)."""
    use_all_host_cores()
    from oracle import unires_port as P
    from oracle.nitorch_shim.core import optim as OO
    vx = torch.ones(3) * float(sc.cfg['vx_y'])
    kw = dict(method=sc.sett.method, do=sc.sett.do_proj)
    b = sc.x[0][0].tau * P.proj('At', sc.x[0][0].dat, sc.x[0], sc.y[0], n=0, **kw)
    lhs = lambda v: P.proj('AtA', v, sc.x[0], sc.y[0], rho=sc.rho, vx_y=vx, **kw)
    stamps = []
    OO.cg(A=lhs, b=b, x=sc.y[0].dat.clone(), max_iter=n_iters + 1, tolerance=0, stop='max_gain',
          record=lambda it, xi: stamps.append(time.perf_counter()))
    return (stamps[-1] - stamps[0]) / n_iters


def cpu_baseline(sc_gpu, args):
    """Oracle port on the host cores, bounded sample: channel 0, a few CG iterations."""
    from unires_b200 import synth
    from oracle.adapters import port_ops, port_structs
    cfg = dict(sc_gpu.cfg)
    cfg['thick'] = [cfg['thick'][0]]
    truth = [sc_gpu.truth[0].cpu()]
    sc = synth.make_scenario(cfg, port_ops, port_structs, device='cpu', truth=truth)
    sec = time_oracle_cg(sc, args.cpu_sample_iters)
    return {'value': 1.0 / sec, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': 'oracle/unires_port.py (restated reference, pure-PyTorch primitives) on CPU: '
                      'channel 0 of %s, %d CG iterations at full size (tolerance 0), %.1f s per '
                      'iteration' % (args.workload, args.cpu_sample_iters, sec)}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  nitorch is not
    installable (no network, not vendored) and /root/reference does not exist on the GPU box,
    so this times the oracle port (restated reference) on all host cores; each step is a
    bounded sample (one full-size CG iteration of channel 0)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sc = oracle_problem(args.workload, 0)
    n = max(1, args.steps)
    sec = time_oracle_cg(sc, n)
    val = 1.0 / sec
    dim = tuple(sc.y[0].dim)
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT,
            'n_gpus': int(os.environ.get('WORLD_SIZE', '1')), 'steps': n, 'warmup': 1,
            'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': '%s: %s, recon grid %s, CG '
                                   'y-update (tolerance 0); bounded sample: each step is ONE full-size '
                                   'CG iteration of channel 0 on the host cores'
                                   % (args.workload, describe(args.workload),
                                      'x'.join(map(str, dim))),
                       'channels': 1, 'recon_grid': list(dim), 'cg_iters_per_channel': 1},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': torch.get_num_threads(),
                             'kind': 'port',
                             'sample': '%d full-size CG iterations of channel 0 after 1 untimed '
                                       'iteration' % n},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
