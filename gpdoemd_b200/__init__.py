"""gpdoemd_b200 -- B200-native design-evaluation hot path of GPdoemd.

GP predictive mean/variance and p-derivatives -> Taylor marginals -> design criteria -> argmax, in
FP64 CUDA for sm_100a behind the reference's Python classes and function signatures.  Importing this
package does not need a GPU; the first call that computes anything does, and raises if the CUDA library
or device is missing (there is no CPU fallback).
"""
from . import design_criteria, discrimination_criteria, kernels, marginal, scoring, synthetic
from .context import Context, get_context, set_default_device
from .design_criteria import AW, BF, BH, HR, JR
from .gp_state import DeviceExactGP, ExactGP, GPState, SparseGP
from .design_gradients import criterion_gradient
from .marginal import taylor_first_order, taylor_first_order_dx, taylor_multi, taylor_second_order
from .models import GPModel, SparseGPModel
from .param_covar import laplace_approximation
from .scoring import criterion_with_gradient
from .scoring import LibraryComm, MultiDeviceScorer, ShardedScorer, refine_designs, score_designs, score_designs_dev, score_designs_multi

__all__ = ['GPModel', 'SparseGPModel', 'GPState', 'ExactGP', 'DeviceExactGP', 'SparseGP', 'Context', 'get_context', 'set_default_device', 'kernels', 'marginal', 'design_criteria', 'discrimination_criteria',
           'scoring', 'taylor_first_order', 'taylor_second_order', 'taylor_multi', 'taylor_first_order_dx', 'criterion_gradient', 'criterion_with_gradient', 'laplace_approximation', 'HR', 'BH', 'BF', 'AW', 'JR', 'score_designs',
           'score_designs_dev', 'score_designs_multi', 'refine_designs', 'ShardedScorer', 'MultiDeviceScorer', 'LibraryComm', 'synthetic']
