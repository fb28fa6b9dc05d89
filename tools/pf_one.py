"""per-frame uptake mode (average_first=False): warm CUDA-graph time of the pass, and a few plain steps for
`ncu -k regex:bv_pass_kernel`.  usage: pf_one.py F R T [time]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from jaxent_b200 import _lib as L, synthetic
from jaxent_b200.engine import BvEngine, OptSettings, SlotSpec
from perf_kernels import timeit

F, R, T = [int(x) for x in sys.argv[1:4]]
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(0)
heavy = torch.randint(0, 40, (R, F), device=dev, generator=g, dtype=torch.int32).float()
acc = torch.randint(0, 3, (R, F), device=dev, generator=g, dtype=torch.int32).float()
small = synthetic.make_ensemble(2048, R, T, seed=0)
slots = [SlotSpec(L.LOSS_UPTAKE_EYE_MSE, small.S_train, small.y_train, small.S_val, small.y_val), SlotSpec(L.LOSS_MAXENT_CONVEX_KL)]
eng = BvEngine(heavy, acc, small.k_ints, small.timepoints, slots, True, OptSettings(learning_rate=1.0, clip_value=None),
               prior=torch.full((F,), 1.0 / F, device=dev), average_first=False)
eng.init_state(torch.ones(F, device=dev), 0.35, 2.0, [0.5, 0.5], [1000., 1000.])
eng.step(4); torch.cuda.synchronize()
if len(sys.argv) > 4:
    lib, plan = eng.lib, eng.plan
    tp = timeit(lambda s: lib.jaxent_bv_pass(plan, 1, s), reps=5, outer=4)
    def step(s):
        lib.jaxent_bv_pass(plan, 1, s); lib.jaxent_bv_reduce(plan, s); lib.jaxent_bv_tail(plan, s)
    ts = timeit(step, reps=5, outer=4)
    print(f"per-frame F={F} R={R} T={T}: pass {tp:.1f} us, full step {ts:.1f} us", flush=True)
    dbg = eng._view(8, torch.int64).cpu().numpy()
    if dbg[50] > 0:  # JX_PROFILE build: clock64 sums of CTA 0 (lane 0 of the Adam warps, consumer thread 0), cycles per tile
        n = float(dbg[50])
        print(f"  per tile: adam0 wait {dbg[48]/n:.0f} work {dbg[49]/n:.0f} | adam1 wait {dbg[51]/n:.0f} work {dbg[52]/n:.0f} | "
              f"consumer ring {dbg[56]/n:.0f} A {dbg[57]/n:.0f} waitE {dbg[58]/n:.0f} B {dbg[59]/n:.0f} | "
              f"adam0 since sync: row sums {dbg[63]/n:.0f} H {dbg[60]/n:.0f} math {dbg[61]/n:.0f} arrive {dbg[62]/n:.0f}", flush=True)
