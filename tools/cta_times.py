"""per-CTA busy time of the pass over several launches: is the end-of-pass straggle tied to the SM, the CTA index or the
feature region?  Needs a library built with JAXENT_NVCC_FLAGS=-DJX_PROFILE (JAXENT_B200_LIB=...); add -DJX_REVERSE_PARTITION
to reverse the CTA index -> region assignment, SHIFT_MB=n to shift the allocation.  usage: cta_times.py F   """
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from jaxent_b200 import _lib as L, synthetic
from jaxent_b200.engine import BvEngine, OptSettings, SlotSpec
F = int(sys.argv[1]); R, T = 500, 8
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
SHIFT = int(os.environ.get('SHIFT_MB', 0))
dummy = torch.empty(SHIFT << 20, dtype=torch.uint8, device=dev) if SHIFT else None
g = torch.Generator(device=dev); g.manual_seed(0)
heavy = torch.randint(0, 40, (R, F), device=dev, generator=g, dtype=torch.int32).float()
acc = torch.randint(0, 3, (R, F), device=dev, generator=g, dtype=torch.int32).float()
small = synthetic.make_ensemble(2048, R, T, seed=0)
slots = [SlotSpec(L.LOSS_UPTAKE_EYE_MSE, small.S_train, small.y_train, small.S_val, small.y_val), SlotSpec(L.LOSS_MAXENT_CONVEX_KL)]
eng = BvEngine(heavy, acc, small.k_ints, small.timepoints, slots, True, OptSettings(learning_rate=1.0, clip_value=None), prior=torch.full((F,), 1.0 / F, device=dev))
del heavy, acc
eng.init_state(torch.ones(F, device=dev), 0.35, 2.0, [0.5, 0.5], [1000., 1000.])
dbg = eng._view(8, torch.int64)
eng.step(10); torch.cuda.synchronize()
busy, ends, smids = [], [], []
for trial in range(12):
    eng.step(3); torch.cuda.synchronize()
    a = dbg.cpu().numpy().copy()
    per = a[64:].reshape(-1, 4).astype(np.float64); per = per[per[:, 0] > 0]
    t0 = per[:, 1].min()
    busy.append((per[:, 2] - per[:, 1]) / 1e3); ends.append((per[:, 2] - t0) / 1e3); smids.append(per[:, 3].astype(int))
busy, ends, smids = np.array(busy), np.array(ends), np.array(smids)
print(f"F={F}: CTAs {busy.shape[1]}; end of pass (max over CTAs) per trial: {np.round(ends.max(1),1)}")
print(f"busy per CTA: mean {busy.mean():.1f} us, min {busy.min(1).mean():.1f}, max {busy.max(1).mean():.1f}; end: first {ends.min(1).mean():.1f} median {np.median(ends,1).mean():.1f} last {ends.max(1).mean():.1f}")
same_sm = (smids == smids[0]).all()
print("CTA -> SM mapping identical across launches:", same_sm, "| SMs used:", len(np.unique(smids[0])), "| CTAs per SM:", np.bincount(np.bincount(smids[0])))
# is lateness systematic per CTA index / per SM?
m = ends.mean(0); sd = ends.std(0)
print(f"per-CTA mean end: spread (std over CTAs of the means) {m.std():.2f} us; typical run-to-run std of one CTA {sd.mean():.2f} us")
cc = np.corrcoef(ends)
print("correlation of per-CTA end times between launches (mean off-diagonal):", round((cc.sum() - len(cc)) / (len(cc) ** 2 - len(cc)), 3))
# by SM
sm_mean = {}
for sm in np.unique(smids[0]):
    sm_mean[sm] = ends[:, smids[0] == sm].mean()
v = np.array(list(sm_mean.values())); k = np.array(list(sm_mean.keys()))
o = np.argsort(v)
print("earliest SMs:", [(int(k[i]), round(v[i], 1)) for i in o[:8]])
print("latest SMs:  ", [(int(k[i]), round(v[i], 1)) for i in o[-8:]])
# by CTA index halves / position in memory
n = busy.shape[1]
print("mean end by CTA-index octile:", [round(ends[:, i * n // 8:(i + 1) * n // 8].mean(), 1) for i in range(8)])
print("mean end by smid octile:", [round(np.mean([sm_mean[s] for s in k[(k >= i * 148 // 8) & (k < (i + 1) * 148 // 8)]]), 1) for i in range(8)])
rel = []
# re-run to collect released (start of work) per CTA too
busy2, ends2, rel2, st2 = [], [], [], []
for trial in range(12):
    eng.step(3); torch.cuda.synchronize()
    a = dbg.cpu().numpy().copy()
    per = a[64:].reshape(-1, 4).astype(np.float64); per = per[per[:, 0] > 0]
    t0 = per[:, 1].min()
    st2.append((per[:, 0] - t0) / 1e3); rel2.append((per[:, 1] - t0) / 1e3); ends2.append((per[:, 2] - t0) / 1e3); busy2.append((per[:, 2] - per[:, 1]) / 1e3)
st2, rel2, ends2, busy2 = map(np.array, (st2, rel2, ends2, busy2))
n = busy2.shape[1]
oc = lambda x, k=16: [round(float(x[:, i * n // k:(i + 1) * n // k].mean()), 1) for i in range(k)]
print("by CTA-index 16-tile: kernel start", oc(st2))
print("by CTA-index 16-tile: released    ", oc(rel2))
print("by CTA-index 16-tile: busy        ", oc(busy2))
print("by CTA-index 16-tile: end         ", oc(ends2))
np.save(os.path.join(ROOT, "gpurun_out", f"cta_ends_{F}.npy"), np.stack([st2.mean(0), rel2.mean(0), busy2.mean(0), ends2.mean(0)]))
