"""B200-native batched differentiable tube MPC (Dubins + discrete barrier state).

Drop-in for the solver / IFT layer of lmcggg/differentiable-tube-mpc: the reference's function-level API
(`ilqr_solve`, `ddp_sensitivity`, `ift_gradient`, `run_closed_loop_experiment`, `ILQRConfig`, ...) over
hand-written sm_100a CUDA kernels behind the C ABI of include/dtmpc.h.  There is no CPU fallback.
"""
from .config import default_dubins_config, load_config
from .spec import (AuxiliaryTheta, BoxClampControl, CircleObstacle, DBaSConfig, DubinsConfig, ILQRConfig, NominalTheta,
                   OcpSpec, SystemSpec, TubeProblem, is_paper_mode, system_from_config)

__all__ = [
    "default_dubins_config", "load_config", "AuxiliaryTheta", "BoxClampControl", "CircleObstacle", "DBaSConfig",
    "DubinsConfig", "ILQRConfig", "NominalTheta", "OcpSpec", "SystemSpec", "TubeProblem", "is_paper_mode",
    "system_from_config",
]
__version__ = "0.1.0"
