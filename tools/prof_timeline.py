"""2-GPU timeline of one fused step (JX_PROFILE build): torchrun --nproc-per-node 2 scratch/prof2.py [nvlink|nccl]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from jaxent_b200 import _lib as L, synthetic
from jaxent_b200.engine import BvEngine, OptSettings, SlotSpec
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "nvlink"
R, T = 500, 8
Fl = int(os.environ.get("FRAMES_PER_RANK", 500_000))
F = Fl * world
use_graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
g = torch.Generator(device=dev); g.manual_seed(rank)
heavy = torch.randint(0, 40, (R, Fl), device=dev, generator=g, dtype=torch.int32).float()
acc = torch.randint(0, 3, (R, Fl), device=dev, generator=g, dtype=torch.int32).float()
small = synthetic.make_ensemble(2048, R, T, seed=0)
slots = [SlotSpec(L.LOSS_UPTAKE_EYE_MSE, small.S_train, small.y_train, small.S_val, small.y_val), SlotSpec(L.LOSS_MAXENT_CONVEX_KL)]
eng = BvEngine(heavy, acc, small.k_ints, small.timepoints, slots, True, OptSettings(learning_rate=1.0, clip_value=None),
               prior=torch.full((Fl,), 1.0 / F, device=dev), F_total=F, group=dist.group.WORLD if world > 1 else None, exchange=mode)
del heavy, acc
eng.init_state(torch.ones(Fl, device=dev), 0.35, 2.0, [0.5, 0.5], [1000., 1000.])
dbg = eng._view(8, torch.int64)
rows = []
if use_graph:
    eng.step(4); eng.capture(4); torch.cuda.synchronize()
for trial in range(4):
    dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    if use_graph: eng.step_graph(10)
    else: eng.step(40)
    e1.record(); torch.cuda.synchronize()
    period = e0.elapsed_time(e1) * 1e3 / 40
    if rank == 0: print(f"rank 0 period {period:.1f} us/step ({mode}, graph={use_graph}, world={world})", flush=True)
    rows.append(dbg.cpu().numpy().copy())
    a = rows[-1]
    if os.environ.get("COMPACT") and trial > 0:
        t0 = a[41]
        g = lambda i: (a[i] - t0) / 1e3
        per = a[64:].reshape(-1, 4).astype(np.float64); per = per[per[:, 0] > 0]
        pc = lambda v: "/".join(f"{x:.1f}" for x in np.percentile((v - t0) / 1e3, [0, 50, 100]))
        print(f"[r{rank}] per-CTA start {pc(per[:,0])} released {pc(per[:,1])} end {pc(per[:,2])}")
        print(f"[r{rank}] period {period:.1f} | lastCTA {g(44):.1f} tail.pdl {g(2):.1f} scalars {g(3):.1f} fin {g(9):.1f} ctl written {g(10):.1f} -> release gap {period - g(10):.1f} | tail body {g(9)-g(3):.1f}", flush=True)
dist.barrier()
names = {40: "pass.start", 41: "pass.pdl", 42: "pass.kl", 43: "pass.cta0end", 44: "red.start", 45: "red.pdl", 46: "red.end",
         0: "tail.start", 1: "tail.staged", 2: "tail.pdl", 3: "tail.scalars", 4: "tail.lnpf", 5: "tail.pred", 7: "tail.back", 8: "tail.sums", 9: "tail.fin", 16: "a.begin", 17: "a.loop", 18: "a.wsum", 19: "b.begin", 20: "b.loop", 13: "rep1.start", 14: "rep0.end", 10: "tail.ctl", 11: "tail.frames", 12: "tail.end",
         32: "ftail.start", 33: "ftail.pdl", 34: "ftail.scalars", 35: "ftail.frames", 36: "ftail.end"}
for r in range(world):
    if r == rank and not os.environ.get("COMPACT"):
        for a in rows[1:]:
            # last step: pass(40..43) -> reduce(44..46) -> tail(0..12); everything relative to pass.pdl
            t0 = a[41]
            order = [40, 41, 42, 43, 44, 45, 46, 0, 1, 2, 3, 4, 16, 17, 18, 5, 19, 20, 7, 8, 9, 10, 11, 12, 33, 34, 35, 36]
            print("cycles: qloop", a[21], "nnz", a[22], "epilogue", a[23])
            print(f"rank {rank}: " + " ".join(f"{names[i]}={(a[i]-t0)/1e3:.1f}" for i in order), flush=True)
            if len(a) > 64:
                per = a[64:].reshape(-1, 4).astype(np.float64)
                per = per[per[:, 0] > 0]
                q = lambda v: " ".join(f"{x:.1f}" for x in np.percentile((v - t0) / 1e3, [0, 10, 50, 90, 100]))
                print(f"   per-CTA (n={len(per)}) min/p10/p50/p90/max: start {q(per[:,0])} | released {q(per[:,1])} | end {q(per[:,2])} | busy(end-released) {' '.join(f'{x:.1f}' for x in np.percentile((per[:,2]-per[:,1])/1e3,[0,10,50,90,100]))}", flush=True)
    dist.barrier()
dist.destroy_process_group()
