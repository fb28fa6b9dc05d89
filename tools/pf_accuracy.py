"""worst per-step loss / gradient error of the per-frame mode against the fp64 oracle (the cases of tests/test_gpu_perframe.py)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as H
from oracle import bv_oracle as O
from test_gpu_parity import _check_steps

def case_(F, R, T, kind=O.UPTAKE_EYE_MSE, seed=4, **kw):
    c = H.make_case(F, R, T, kind, seed=seed, **kw); c["problem"].average_first = False; return c

for kind, F, R, T, clip, om in [(O.UPTAKE_EYE_MSE, 777, 97, 5, None, False), (O.UPTAKE_MC_EYE_MSE, 1000, 53, 3, 1.0, False),
                                (O.UPTAKE_SIGMA_MSE, 515, 64, 8, None, False), (O.UPTAKE_EYE_MSE, 4099, 500, 8, None, False),
                                (O.UPTAKE_EYE_MSE, 4099, 500, 8, None, True), (O.UPTAKE_EYE_MSE, 20000, 200, 8, None, False)]:
    case = case_(F, R, T, kind, gamma=2.0, prior="random")
    cfg = O.OptConfig(learning_rate=1.0 if not om else 0.2, initial_learning_rate=1.0, initial_steps=0, clip_value=clip, optimise_model_parameters=om)
    try:
        _, wl, wg = _check_steps(case, cfg, 6, STEP_RTOL=1e-2)
        print(f"{kind[:22]:22s} F={F} R={R} T={T} opt_model={om}: worst loss rel err {wl:.2e}, worst grad err {wg:.2e}", flush=True)
    except AssertionError as e:
        print("FAILED", kind, F, R, T, str(e)[:200], flush=True)
