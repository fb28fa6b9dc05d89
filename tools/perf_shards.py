import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from perf_kernels import make, timeit
for (F,R,T) in [(500,53,3),(1000,130,8),(125_000,500,8),(250_000,500,8),(500_000,500,8),(1_000_000,500,8)]:
    eng = make(F,R,T)
    lib, plan = eng.lib, eng.plan
    tp = timeit(lambda s: lib.jaxent_bv_pass(plan,1,s))
    tr = timeit(lambda s: lib.jaxent_bv_reduce(plan,s))
    tt = timeit(lambda s: lib.jaxent_bv_tail(plan,s))
    def step(s):
        lib.jaxent_bv_pass(plan,1,s); lib.jaxent_bv_reduce(plan,s); lib.jaxent_bv_tail(plan,s)
    ts = timeit(step, reps=10)
    print(f"F={F} R={R} T={T}: warm back-to-back us: pass {tp:.1f} reduce {tr:.1f} tail {tt:.1f} | sum {tp+tr+tt:.1f} | full step {ts:.1f} | overhead {ts-tp:.1f}", flush=True)
    del eng
