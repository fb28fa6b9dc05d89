import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jaxent_b200 import _lib as L, synthetic
from jaxent_b200.engine import OptSettings, SlotSpec
from jaxent_b200.batch_engine import BatchedBvEngine
F, R, T, B = 100_000, 500, 8, 256
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device('cuda')
g = torch.Generator(device=dev); g.manual_seed(0)
heavy = torch.randint(0, 40, (R, F), device=dev, generator=g, dtype=torch.int32).float()
acc = torch.randint(0, 3, (R, F), device=dev, generator=g, dtype=torch.int32).float()
small = synthetic.make_ensemble(2048, R, T, seed=0)
slots = [SlotSpec(L.LOSS_UPTAKE_EYE_MSE, small.S_train, small.y_train, small.S_val, small.y_val), SlotSpec(L.LOSS_MAXENT_CONVEX_KL)]
eng = BatchedBvEngine(heavy, acc, small.k_ints, small.timepoints, slots, True, OptSettings(learning_rate=1.0, clip_value=None), B,
                      prior=torch.full((F,), 1.0 / F, device=dev), n_pieces=NS)
del heavy, acc
gam = np.logspace(0, 3, B).astype(np.float32)
fw = np.stack([gam / (gam + 1), 1 / (gam + 1)], 1).astype(np.float32)
fs = np.full((B, 2), 1000.0, np.float32)
eng.init_state(torch.ones(F, device=dev), 0.35, 2.0, fw, fs, np.full(B, 1.0, np.float32))
eng.step(3); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
K = 20
e0.record(); eng.step(K); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
fl = 8 * R * F * B
print(f"NS={NS}: batched step {ms*1e3:.1f} us for {B} replicas = {1e3/ms:.1f} steps/s ({B*1e3/ms:.0f} replica-steps/s); algorithmic {fl/ms/1e9:.1f} TFLOP/s, executed x{NS} = {NS*fl/ms/1e9:.1f} TFLOP/s")
# per-kernel timing through the building blocks
lib, plan = eng.lib, eng.plan
r = eng.records_at(eng.steps_done)
print("loss replica 0 / 255:", r[0].total_train, r[255].total_train)
# per-phase timing (events between the four phases of a step; phases are serialised here, no side-stream overlap)
import ctypes as C
from jaxent_b200.engine import _stream_ptr
K = 10
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)]
for k in range(K):
    evs[k][0].record()
    for ph in range(4):
        L.check(lib.jaxent_bg_step_phase(plan, ph, _stream_ptr()))
        evs[k][ph + 1].record()
torch.cuda.synchronize()
names = ["gemm1 (+fused grad)", "grad+decide+adam", "gemm2", "reduce2+tail+kl+klfold"]
tot = 0.0
for ph in range(4):
    t = np.median([evs[k][ph].elapsed_time(evs[k][ph + 1]) for k in range(K)]) * 1e3
    tot += t
    print(f"  phase {ph} {names[ph]:28s} {t:7.1f} us")
print(f"  sum of phases {tot:.1f} us")
