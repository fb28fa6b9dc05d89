"""the end-to-end job bench.py times (pinned host features -> Simulation.initialise -> OptaxOptimizer -> run_optimise -> read-back),
repeated with a per-phase breakdown, then one job under cProfile.  usage: e2e_probe.py [K]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
import jaxent_b200 as J
from jaxent_b200.opt.run import run_optimise, _optimise
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
wl = B.WORKLOADS["cfg4"]; F, R, T = wl["F"], wl["R"], wl["T"]
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
small = B.small_problem(R, T)
heavy, acc, lo, hi = B.gen_features(R, F, 0, 1, dev)
host_h = torch.empty((R, F), dtype=torch.float32, pin_memory=True); host_a = torch.empty((R, F), dtype=torch.float32, pin_memory=True)
host_h.copy_(heavy); host_a.copy_(acc); del heavy, acc; torch.cuda.synchronize(); torch.cuda.empty_cache()
class _Model:
    def __init__(self, fp): self.forwardpass = fp
def job(opt_fn, tag):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    feats = J.BV_input_features(host_h, host_a, small.k_ints)
    mp = J.BV_Model_Parameters(bv_bc=np.array([0.35], np.float32), bv_bh=np.array([2.0], np.float32), timepoints=small.timepoints)
    yt = lambda y: y[:, :, None]
    loader = J.ExpD_Dataloader(train=J.Dataset([], yt(small.y_train), J.SparseFragmentMapping(small.S_train)), val=J.Dataset([], yt(small.y_val), J.SparseFragmentMapping(small.S_val)))
    params = J.Simulation_Parameters(frame_weights=np.full(F, 1.0 / F, np.float32), frame_mask=np.ones(F, np.float32), model_parameters=[mp],
        forward_model_weights=np.array([1.0, 1.0], np.float32), normalise_loss_functions=np.ones(2, np.float32), forward_model_scaling=np.full(2, 1000.0, np.float32))
    sim = J.Simulation(input_features=(feats,), forward_models=(_Model(J.BV_uptake_ForwardPass()),), params=params, device=dev)
    t1 = time.perf_counter()
    sim.initialise(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    opt = J.OptaxOptimizer(learning_rate=1.0, parameter_partition_masks={J.Optimisable_Parameters.frame_weights}, clip_value=None, optimizer="adam", initial_learning_rate=1.0, initial_steps=0)
    cfg = J.OptimiserSettings(name="bench", n_steps=K, tolerance=1e-30, convergence=1e-30, learning_rate=1.0)
    sim2, hist = run_optimise(sim, (loader, params), cfg, forward_models=(), indexes=(0, 0), loss_functions=[J.hdx_uptake_eye_MSE_loss, J.maxent_convexKL_loss], optimizer=opt, _opt_fn=opt_fn, silent=True)
    w_host = sim2.params.frame_weights.cpu(); last = float(hist.states[-1].losses.total_train_loss); torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"{tag}: objects {1e3*(t1-t0):.1f} ms, initialise {1e3*(t2-t1):.1f} ms, run_optimise+readback {1e3*(t3-t2):.1f} ms, total {1e3*(t3-t0):.1f} ms -> {K/(t3-t0):.1f} steps/s", flush=True)
    del sim, sim2, opt, w_host, hist, feats
    torch.cuda.empty_cache()
for i in range(2): job(_optimise, f"python loop #{i}")
for i in range(3): job(None, f"device loop #{i}")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); job(None, "profiled"); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
