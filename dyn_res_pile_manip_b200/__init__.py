"""B200-native particle-GNN rollout path of dyn-res-pile-manip (drop-in model / planner API).

    from dyn_res_pile_manip_b200 import PropNetDiffDenModel, PlannerGD

mirrors `model.gnn_dyn.PropNetDiffDenModel` and `planners.PlannerGD` of the reference; all
compute runs in hand-written sm_100a kernels (csrc/, C ABI in include/pile_gnn.h).
"""
from . import _lib, ops, synthetic  # noqa: F401
from .propnet import PropModuleDiffDen, PropNetDiffDenModel  # noqa: F401
from .planner import Planner, PlannerGD, particle_num_to_iter_time  # noqa: F401
from .rewards import config_reward_ptcl  # noqa: F401
from .regressor import MPCResRgrNoPool  # noqa: F401

__all__ = ["PropNetDiffDenModel", "PropModuleDiffDen", "Planner", "PlannerGD", "config_reward_ptcl", "MPCResRgrNoPool",
           "particle_num_to_iter_time", "ops", "synthetic"]
