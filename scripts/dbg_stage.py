import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import breeze_b200 as bz
from conftest import make_bubble_model, bubble_theta
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
m = make_bubble_model(bz.B200(use_tma=mode), (32, 16, 24))
m.set(θ=bubble_theta(), u=1.0)
m.context.compute_tendencies()
print("mode", mode, "G_theta absmax", abs(m.context.get_tendency("ρθ")).max())
