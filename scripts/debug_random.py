import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import cases as K
from tests.test_gpu_random_cases import _random_params
from tests.initstate import wavy_state
from life_b200 import capi
from oracle import oracle as O
for seed in [int(a) for a in sys.argv[1:]]:
    p, r = _random_params(seed)
    for steps in (1, 2, 5, 25):
        o = O.Oracle(p)
        f0, rho0, u0 = wavy_state(o.Nx, o.Ny, bool(p.central_moments), amp=0.03, non_equilibrium=0.01)
        o.set("f", f0); o.set("rho", rho0); o.set("u", u0)
        ctx = capi.Context(K.life_config(p, o, kernel=2))
        K.upload_from_oracle(ctx, o)
        for t in range(1, steps + 1):
            ctx.step(t)
        o.step(steps)
        st = ctx.download_state(); ctx.close()
        d = np.abs(st["f"] - o.get("f"))
        i, j, v = np.unravel_index(np.argmax(d), d.shape)
        bad = np.argwhere(d > 1e-12)
        print("seed", seed, "steps", steps, "walls", p.wall_left, p.wall_right, p.wall_bottom, p.wall_top, "N", o.Nx, o.Ny,
              "max|df|=%.3e at (i=%d,j=%d,v=%d)" % (d.max(), i, j, v), "n_bad", len(bad),
              "bad i range", (bad[:, 0].min(), bad[:, 0].max()) if len(bad) else None, "bad j", sorted(set(bad[:, 1]))[:12] if len(bad) else None,
              "bad v", sorted(set(bad[:, 2])) if len(bad) else None)
