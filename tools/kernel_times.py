"""Per-kernel device time with ONE stream active (no L/R overlap): us per image for each kernel (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iv_slam_b200 import api, synthetic as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
W, H, NF = (int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (1241, 376, 2000)
L, R = S.make_stereo_batch(W, H, n, 100, distinct=min(16, n))
a = (NF, 1.2, 8, 20, 7)
INTRO = len(sys.argv) > 6 and sys.argv[6] == 'intro'      # left eye with an introspection cost-map (BASELINE C2 style)
gL, gR = api.ORBextractor(*a, INTRO), api.ORBextractor(*a)
cost = np.stack([S.make_cost_map(W, H, 7 + i) for i in range(min(n, 4))] * ((n + 3) // 4))[:n] if INTRO else None
gL.upload(L, cost); gR.upload(R)
def step():
    gL.run(); gL.sync(); gR.run(); gR.sync()
    rc = api.lib().ivg_stereo_match_batch(gL._h, gR._h, 386.1448, 718.856, None, None, gL.cap, 1)
    assert rc == 0
step()
gL.profile_enable(True); gR.profile_enable(True)
for _ in range(steps): step()
pl, pr = gL.profile_read(), gR.profile_read()
tot = 0
for k in pl:
    ms = pl[k][0] + pr[k][0]
    per = ms * 1e3 / (steps * (n if k.startswith('k_stereo') else 2 * n))
    tot += per * (1 if k.startswith('k_stereo') else 2)
    print('%-20s %8.3f ms/step  %7.3f us per %s' % (k, ms / steps, per, 'pair' if k.startswith('k_stereo') else 'image'))
print('sum per pair %.2f us -> %.0f pairs/s if serial' % (tot, 1e6 / tot))
