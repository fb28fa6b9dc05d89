import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jaxent_b200 import _lib as L, synthetic
from jaxent_b200.engine import BvEngine, OptSettings, SlotSpec

def make(F,R,T):
    dev = torch.device('cuda')
    g = torch.Generator(device=dev); g.manual_seed(0)
    heavy = torch.randint(0, 40, (R,F), device=dev, generator=g, dtype=torch.int32).float()
    acc = torch.randint(0, 3, (R,F), device=dev, generator=g, dtype=torch.int32).float()
    small = synthetic.make_ensemble(2048, R, T, seed=0)
    slots = [SlotSpec(L.LOSS_UPTAKE_EYE_MSE, small.S_train, small.y_train, small.S_val, small.y_val), SlotSpec(L.LOSS_MAXENT_CONVEX_KL)]
    eng = BvEngine(heavy, acc, small.k_ints, small.timepoints, slots, True, OptSettings(learning_rate=1.0, clip_value=None), prior=torch.full((F,), 1.0/F, device=dev))
    eng.init_state(torch.ones(F, device=dev), 0.35, 2.0, [0.5,0.5], [1000.,1000.])
    eng.step(4); torch.cuda.synchronize()
    return eng

def timeit(fn, reps=20, outer=5):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps): fn(side.cuda_stream)
    torch.cuda.current_stream().wait_stream(side)
    g.replay(); torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(outer): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)*1e3/(reps*outer)

if __name__ == "__main__":
  for (F,R,T) in [(500,53,3),(1000,130,8),(100_000,500,8),(1_000_000,500,8)]:
      eng = make(F,R,T)
      lib, plan = eng.lib, eng.plan
      tp = timeit(lambda s: lib.jaxent_bv_pass(plan,1,s))
      tr = timeit(lambda s: lib.jaxent_bv_reduce(plan,s))
      tt = timeit(lambda s: lib.jaxent_bv_tail(plan,s))
      def step(s):
          lib.jaxent_bv_pass(plan,1,s); lib.jaxent_bv_reduce(plan,s); lib.jaxent_bv_tail(plan,s)
      ts = timeit(step, reps=10)
      print(f"F={F} R={R} T={T}: warm back-to-back us: pass {tp:.1f} reduce {tr:.1f} tail {tt:.1f} | sum {tp+tr+tt:.1f} | full step {ts:.1f}")
      del eng
