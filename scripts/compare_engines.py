"""Debug aid (GPU box): compare every activation, activation-gradient and variable gradient of the
tensor-core engine against the FFMA engine on the same batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy, torch
from hypelcnn_b200 import engine as E
from oracle import hypelcnn_ref as R
from tests.util import ALG, synthetic_batch

case = sys.argv[1] if len(sys.argv) > 1 else "tiny"
CASES = {
    "tiny": dict(P=3, C=10, classes=4, alg={**ALG, "filter_count": 32}, B=24),
    "c5": dict(P=3, C=65, classes=11, alg=ALG, B=32),
    "c2": dict(P=7, C=145, classes=15, alg=ALG, B=16),
    "nonres": dict(P=5, C=20, classes=6, alg={**ALG, "filter_count": 64, "use_residual": False}, B=20),
}
c = CASES[case]
alg = {**c["alg"], "drop_out_ratio": 0.0}
x, y = synthetic_batch(c["B"], c["P"], c["C"], c["classes"])
xd, yd = torch.tensor(x).cuda(), torch.tensor(y).cuda()
engs = {}
for prec in ("fp32", "3xtf32"):
    e = E.PatchEngine(c["P"], c["C"], c["classes"], alg, max_batch=c["B"], precision=prec)
    e.init_variables(1234)
    e.forward(xd, True, True, 0)
    e.loss_backward(xd, yd)
    torch.cuda.synchronize()
    engs[prec] = e
a, b = engs["fp32"], engs["3xtf32"]


def rel(u, v):
    s = max(float(v.abs().max()), 1e-30)
    return float((u - v).abs().max()) / s


plan = R.build_plan(c["P"], c["C"], c["classes"], alg, True)
seen = set()
for l in reversed(plan):
    if l.kind == "level_end" or (l.kind == "conv" and l.concat_slot is None) or l.kind == "fc":
        t = l.dst
        if t in seen:
            continue
        seen.add(t)
        ga, gb = a.debug_tensor(t, 2), b.debug_tensor(t, 2)
        aa, ab = a.debug_tensor(t, 0), b.debug_tensor(t, 0)
        print(f"tensor {t:22s} act rel {rel(ab, aa):.2e}   grad rel {rel(gb, ga):.2e}")
for name in reversed(list(a.variables)):
    if a.variables[name][0] in (0, 1):
        print(f"var {name:50s} grad rel {rel(b.gradient(name), a.gradient(name)):.2e}")
