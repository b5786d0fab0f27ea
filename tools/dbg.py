import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
from common import load_golden, View
from parsy_bench_b200 import executor as ex, _lib
G = load_golden("2d5_N30_c8_l1_d2")
def run(S, values=None, fill=np.nan):
    lv = np.full(G.meta["xsize"], fill)
    ok = ex.cholesky_left_par_05(S.n, S.A2_p, S.A2_i, S.A2_x if values is None else values, S.p, S.s, S.i_ptr, lv, S.super, S.nsuper, None, S.sParent, None, None, S.col2Sup, len(S.levelPtr)-1, S.levelPtr, None, 0, S.parPtr, S.partition)
    print(ok, _lib.last_error() if not ok else '', np.abs(lv-G.valL).max() if ok else '')
par = G.parPtr.tolist()
e = View(G); e["parPtr"] = np.array([0, 0] + par[1:], np.int32); lp = G.levelPtr.copy(); lp[1:] += 1; e["levelPtr"] = lp
print('plain'); run(G)
print('empty-part'); run(e)
print('empty-part zeros'); run(e, fill=0.0)
vals = G.A2_x.copy(); vals[G.A2_p[G.n // 2]] = -4.0
print('nonspd'); run(G, vals)
print('plain after nonspd'); run(G)
print('empty-part after'); run(e)
