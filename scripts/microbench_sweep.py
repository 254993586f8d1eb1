"""Sweep tuning knobs of the streaming lhs kernel: python scripts/microbench_sweep.py [workload]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unires_b200 import _project, struct, synth  # noqa: E402
from scripts.microbench_lhs import tune, time_op  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    workload = sys.argv[1] if len(sys.argv) > 1 else 'sr3_256'
    sc = synth.make_scenario(synth.CONFIGS[workload], _project, struct, device=dev, seed=0)
    dim = tuple(sc.y[0].dim)
    n = dim[0] * dim[1] * dim[2]
    v = torch.rand(dim, device=dev)
    vx = [float(sc.cfg['vx_y'])] * 3
    ops = [('ch%d' % c, _project.LhsOperator(sc.x[c], sc.y[c], method=sc.sett.method,
                                             do=sc.sett.do_proj, rho=sc.rho, vx_y=vx))
           for c in range(len(sc.x))]
    ops.append(('denoise', _project.LhsOperator([struct._input(tau=0.01)],
                                                struct._output(dim=dim, lam=0.1), do=False,
                                                rho=1.0, vx_y=vx)))
    for name, op in ops:
        for rpt in (1, 2):
            for pf in (0, 2, 4, 8):
                tune('stream_rpt', rpt)
                tune('stream_pf', pf)
                try:
                    us = time_op(op, v)
                except Exception as e:
                    print(name, rpt, pf, 'failed', e)
                    continue
                print('%-8s rpt %d pf %d: %7.1f us  %7.1f GB/s' % (name, rpt, pf, us, 8 * n / us / 1e3),
                      flush=True)
    tune('stream_rpt', 0)
    tune('stream_pf', 0)


if __name__ == '__main__':
    main()
