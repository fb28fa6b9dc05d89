import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from perf_kernels import make
F, R, T = [int(x) for x in sys.argv[1:4]]
eng = make(F, R, T)
eng.step(6); torch.cuda.synchronize()
