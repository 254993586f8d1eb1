#!/bin/bash
mkdir -p gpurun_out
python -c "
import cProfile, pstats, sys, io
sys.argv = ['sr_demo.py', '--max-iter', '40']
sys.path.insert(0, 'demos')
import runpy
pr = cProfile.Profile()
pr.enable()
try:
    runpy.run_path('demos/sr_demo.py', run_name='__main__')
finally:
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
    print(s.getvalue()[:9000])
" > gpurun_out/demo_profile.log 2>&1
tail -70 gpurun_out/demo_profile.log | cut -c1-160
