"""Setup time of one exact GP: host LAPACK inference + n x n factor upload (gpd_gp_upload) vs the
whole inference on the device (gpd_gp_build).  Both end with the packed L^-1 resident on the GPU."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import gpdoemd_b200 as g
from gpdoemd_b200 import kernels
from gpdoemd_b200.gp_state import ExactGP, GPState, ProductKernel

ctx = g.get_context(0)
out = []
for n in ([int(a) for a in sys.argv[1:]] or [500, 2000, 4000]):
    rng = np.random.default_rng(n)
    Z = rng.uniform(size=(n, 7))
    Y = np.sin(3 * Z.sum(1))[:, None]
    hyp = np.r_[1.3, rng.uniform(0.3, 2, 3), 0.8, rng.uniform(0.3, 2, 4), 1e-6]

    def kern():
        return ProductKernel(kernels.Matern52(3, [0, 1, 2], 'kernx'), kernels.Matern52(4, range(3, 7), 'kernp'))
    res = {'n': n}
    for rep in range(2):
        t0 = time.perf_counter()
        hg = ExactGP(Z, Y, kern())
        hg[:] = hyp
        t1 = time.perf_counter()
        h = GPState.from_gp(hg, 3, 4).upload(ctx)
        ctx.sync()
        t2 = time.perf_counter()
        ctx.lib.gpd_gp_free(h)
        dg = g.DeviceExactGP(Z, Y, kern())
        dg[:] = hyp
        t3 = time.perf_counter()
        h = dg.build(ctx, 3, 4)
        ctx.sync()
        t4 = time.perf_counter()
        ctx.lib.gpd_gp_free(h)
        res = {'n': n, 'host_inference_s': t1 - t0, 'upload_invert_pack_s': t2 - t1, 'device_build_s': t4 - t3,
               'speedup': (t2 - t0) / (t4 - t3)}
    out.append(res)
    print(json.dumps(res))
