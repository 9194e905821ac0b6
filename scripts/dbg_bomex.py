import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import breeze_b200 as bz
from oracle_lib import CPUOracle
gpu = bz.cases.bomex_model(bz.B200(), size=(32, 16, 30), extent=3200.0)
cpu = bz.cases.bomex_model(CPUOracle(), size=(32, 16, 30), extent=3200.0)
names = ["ρu", "ρv", "ρw", "ρθ", "ρq"]
for step in range(5):
    gpu.time_step(2.0); cpu.time_step(2.0)
    out = []
    for n in names:
        a, b = gpu.field(n), cpu.field(n)
        out.append(f"{n}: abs {np.abs(a-b).max():.2e} max {np.abs(b).max():.2e}")
    print(step, " | ".join(out))
