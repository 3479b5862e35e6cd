import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpdoemd_b200 as g
from gpdoemd_b200 import synthetic
ctx = g.get_context(0)
for n in (500, 2000, 4000):
    t0 = time.time()
    ms, ds = synthetic.build_models(g.GPModel, 3, 1, 2, 3, 4, n, 'Matern52')
    t1 = time.time()
    ms[0].device_handle(); ctx.sync()
    t2 = time.time()
    X = synthetic.candidates(1, 256, 3, ds[0]['xlo'], ds[0]['xhi'])
    ms[0].predict(X)
    t3 = time.time()
    ms[0].pmean = ms[0].pmean * 1.01
    ms[0].predict(X)
    t4 = time.time()
    print('n', n, 'host build (2 GPs) %.2fs' % (t1 - t0), 'device upload+invert+pack (2 GPs) %.3fs' % (t2 - t1),
          'first predict %.3fs' % (t3 - t2), 'predict after pmean change %.4fs' % (t4 - t3))
