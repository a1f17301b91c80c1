"""Stage-by-stage GPU vs oracle diagnostics (developer tool; run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from iv_slam_b200 import api, synthetic as S
from oracle import oracle_lib as O

def compare(cfg, with_cost):
    c = S.CONFIGS[cfg]
    L, R = S.make_stereo_pair(c['w'], c['h'], c['seed'])
    cost = S.make_cost_map(c['w'], c['h'], c.get('cost_seed', 2)) if with_cost else None
    args = (c['nfeatures'], c['scaleFactor'], c['nlevels'], c['iniThFAST'], c['minThFAST'])
    gL, gR = api.ORBextractor(*args, with_cost), api.ORBextractor(*args, False)
    oL, oR = O.OracleExtractor(*args, with_cost), O.OracleExtractor(*args, False)
    t = time.time(); kg, dg = gL(L, cost); print(cfg, 'gpu first call s', time.time() - t)
    t = time.time(); kg, dg = gL(L, cost); print(cfg, 'gpu second call s', time.time() - t)
    ko, do = oL(L, cost)
    ok = True
    for l in range(c['nlevels']):
        for which, name in ((0, 'pyr'), (1, 'blur')) + (((2, 'qual'),) if with_cost else ()):
            a, b = gL.level(l, which), oL.level(l, which)
            if b is None: continue
            if not np.array_equal(a, b):
                ok = False
                d = np.argwhere(a != b)
                print('  level', l, name, 'MISMATCH', len(d), 'first', d[:3])
        x, y, r = gL.level_keypoints(l)
        k = oL.level_keypoints(l)
        same = x.size == k.size and np.array_equal(x, k['x']) and np.array_equal(y, k['y']) and np.array_equal(r, k['response'])
        if not same:
            ok = False
            print('  level', l, 'keypoints MISMATCH n', x.size, k.size)
            sg = set(zip(x.tolist(), y.tolist())); so = set(zip(k['x'].tolist(), k['y'].tolist()))
            print('    only gpu', len(sg - so), 'only oracle', len(so - sg), list(sg - so)[:5], list(so - sg)[:5])
    print(cfg, 'cost' if with_cost else 'plain', 'n', kg.size, ko.size, 'stages ok', ok)
    if kg.size == ko.size:
        for f in kg.dtype.names:
            if not np.array_equal(kg[f], ko[f]):
                print('  field', f, 'differs at', int((kg[f] != ko[f]).sum()), 'max abs', float(np.abs(kg[f] - ko[f]).max()))
        print('  desc bits differing', int(np.unpackbits(dg ^ do).sum()), 'of', do.size * 8)
    # stereo
    kgr, dgr = gR(R); kor, dor = oR(R)
    print('  right n', kgr.size, kor.size, 'equal', kgr.size == kor.size and all(np.array_equal(kgr[f], kor[f]) for f in kgr.dtype.names), int(np.unpackbits(dgr ^ dor).sum()) if kgr.size == kor.size else -1)
    u, d = api.compute_stereo_matches(gL, gR, c['mbf'], c['maxD'])
    uo, do_ = O.stereo_match(oL, oR, ko, do, kor, dor, c['mbf'], c['maxD'])
    n = ko.size
    print('  stereo matched gpu', int((u[:n] >= 0).sum()), 'oracle', int((uo >= 0).sum()), 'equal', np.array_equal(u[:n], uo), np.array_equal(d[:n], do_),
          'max |du|', float(np.abs(u[:n] - uo).max()))

if __name__ == '__main__':
    print(api.device_info(0))
    compare('C1', False)
    compare('C2', True)
    compare('C2', False)
