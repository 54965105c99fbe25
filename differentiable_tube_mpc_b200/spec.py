"""Declarative problem description of the Dubins + DBaS tube MPC.

The reference passes Python closures (f, stage_cost, ...) into its solver; a CUDA
kernel cannot call those, so the drop-in takes the data the closures are built
from.  Class and field names follow the reference:

  DubinsConfig      core/systems/dubins.py:10-21
  CircleObstacle    core/systems/dubins_obstacles.py:10-13
  DBaSConfig        core/barrier.py:16-33
  BoxClampControl   core/control.py:38-70
  ILQRConfig        core/ddp.py:12-20
  NominalTheta / AuxiliaryTheta   core/params.py:14-59
  SystemSpec        core/system_spec.py:10-31 (declarative variant)

`TubeProblem.from_config` extracts exactly what core/tube_mpc.py:680-768 (paper mode)
or :51-165 (full mode) extract from the YAML dict, including the quirks listed in
SURVEY App. B (Q2-Q6).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib


@dataclass(frozen=True)
class DubinsConfig:
    dt: float = 0.01
    v_max: float = 10.0
    omega_max: float = math.pi
    w_low: Tuple[float, float, float] = (-0.05, -0.05, -0.05)
    w_high: Tuple[float, float, float] = (0.05, 0.05, 0.05)
    x_target: Tuple[float, float, float] = (10.0, 10.0, math.pi / 4)


@dataclass(frozen=True)
class CircleObstacle:
    center: Tuple[float, float]
    radius: float


@dataclass(frozen=True)
class DBaSConfig:
    barrier_type: str = "inverse"
    alpha: float = 0.1
    gamma: float = 0.0
    eps: float = 1e-6

    def __post_init__(self):
        if self.barrier_type != "inverse":
            # no analytic jacobian exists for the log barrier in the reference either
            # (core/systems/dubins_aug_jac.py:119-122; core/tube_mpc.py:317 forces "inverse")
            raise ValueError(f"Unknown barrier_type: {self.barrier_type} (only the relaxed inverse barrier is supported)")
        if isinstance(self.alpha, (float, int)) and self.alpha < 0:
            raise ValueError("alpha must be >= 0")  # core/barrier.py:43-44
        if isinstance(self.gamma, (float, int)) and not (-1.0 <= self.gamma <= 1.0):
            raise ValueError("gamma must be in [-1, 1]")  # core/barrier.py:90-91


@dataclass(frozen=True)
class BoxClampControl:
    """Hard box u in [u_min, u_max] with the active-set tolerance of core/control.py:58."""
    u_min: Any
    u_max: Any
    active_tol: float = 1e-8

    def bounds(self) -> Tuple[List[float], List[float]]:
        lo = [float(v) for v in (self.u_min.tolist() if isinstance(self.u_min, Tensor) else self.u_min)]
        hi = [float(v) for v in (self.u_max.tolist() if isinstance(self.u_max, Tensor) else self.u_max)]
        return lo, hi

    def clamp(self, u: Tensor) -> Tensor:
        lo, hi = self.bounds()
        return torch.clamp(u, min=torch.tensor(lo, device=u.device, dtype=u.dtype),
                           max=torch.tensor(hi, device=u.device, dtype=u.dtype))

    def active_mask(self, u: Tensor) -> Tensor:
        lo, hi = self.bounds()
        lo_t = torch.tensor(lo, device=u.device, dtype=u.dtype)
        hi_t = torch.tensor(hi, device=u.device, dtype=u.dtype)
        return (u <= (lo_t + self.active_tol)) | (u >= (hi_t - self.active_tol))


@dataclass(frozen=True)
class ILQRConfig:
    horizon: int
    nx: int = 4
    nu: int = 2
    max_iter: int = 30
    tol: float = 1e-6
    reg: float = 1e-6
    line_search_alphas: Tuple[float, ...] = (1.0, 0.5, 0.25, 0.1)


def _softplus(x: Tensor) -> Tensor:
    return torch.nn.functional.softplus(x)


@dataclass
class NominalTheta:
    """theta_bar: raw (unconstrained) parameters, constrained by softplus / tanh (core/params.py:14-37)."""
    Q_raw: Tensor
    R_raw: Tensor
    Qf_raw: Tensor
    qb_raw: Tensor
    alpha_raw: Tensor
    gamma_raw: Tensor
    tight_raw: Tensor

    def Q(self): return _softplus(self.Q_raw)
    def R(self): return _softplus(self.R_raw)
    def Qf(self): return _softplus(self.Qf_raw)
    def qb(self): return _softplus(self.qb_raw)
    def alpha(self): return _softplus(self.alpha_raw) + 1e-6
    def gamma(self): return torch.tanh(self.gamma_raw)
    def tight(self): return _softplus(self.tight_raw)

    def tensors(self) -> List[Tensor]:
        return [self.Q_raw, self.R_raw, self.Qf_raw, self.qb_raw, self.alpha_raw, self.gamma_raw, self.tight_raw]

    def constrained(self) -> Tensor:
        """[12] vector in the C-ABI theta layout."""
        return torch.cat([self.Q(), self.R(), self.Qf(), self.qb().view(1), self.alpha().view(1),
                          self.gamma().view(1), self.tight().view(1)]).detach()

    def chain(self) -> Tensor:
        """d constrained / d raw, same layout (softplus' = sigmoid, tanh' = 1 - tanh^2)."""
        s = torch.sigmoid
        return torch.cat([s(self.Q_raw), s(self.R_raw), s(self.Qf_raw), s(self.qb_raw).view(1),
                          s(self.alpha_raw).view(1), (1.0 - torch.tanh(self.gamma_raw) ** 2).view(1),
                          s(self.tight_raw).view(1)]).detach()


@dataclass
class AuxiliaryTheta:
    """theta: core/params.py:40-59 (no tightening)."""
    Q_raw: Tensor
    R_raw: Tensor
    Qf_raw: Tensor
    qb_raw: Tensor
    alpha_raw: Tensor
    gamma_raw: Tensor

    def Q(self): return _softplus(self.Q_raw)
    def R(self): return _softplus(self.R_raw)
    def Qf(self): return _softplus(self.Qf_raw)
    def qb(self): return _softplus(self.qb_raw)
    def alpha(self): return _softplus(self.alpha_raw) + 1e-6
    def gamma(self): return torch.tanh(self.gamma_raw)

    def tensors(self) -> List[Tensor]:
        return [self.Q_raw, self.R_raw, self.Qf_raw, self.qb_raw, self.alpha_raw, self.gamma_raw]

    def constrained(self) -> Tensor:
        z = torch.zeros(1, dtype=self.Q_raw.dtype, device=self.Q_raw.device)
        return torch.cat([self.Q(), self.R(), self.Qf(), self.qb().view(1), self.alpha().view(1),
                          self.gamma().view(1), z]).detach()

    def chain(self) -> Tensor:
        s = torch.sigmoid
        z = torch.zeros(1, dtype=self.Q_raw.dtype, device=self.Q_raw.device)
        return torch.cat([s(self.Q_raw), s(self.R_raw), s(self.Qf_raw), s(self.qb_raw).view(1),
                          s(self.alpha_raw).view(1), (1.0 - torch.tanh(self.gamma_raw) ** 2).view(1), z]).detach()


@dataclass(frozen=True)
class OcpSpec:
    """One OCP family: ILQRConfig + the DBaS numerical settings its closures would capture."""
    ilqr: ILQRConfig
    eps_dyn: float = 1e-6   # DBaSConfig.eps inside f_hat (1e-6 in full mode, YAML eps in paper mode: App. B Q5)
    eps_jac: float = 1e-6   # DBaSConfig.eps inside f_jac / gradient closures
    angle_wrap: bool = False  # run_nominal.py:297-324
    tight_in_jac: bool = False  # the reference jacobian ignores the nominal tightening

    def to_c(self) -> _lib.Ocp:
        o = _lib.Ocp()
        c = self.ilqr
        if c.nx != 4 or c.nu != 2:
            raise ValueError("the Dubins + DBaS family has nx=4 (x, y, heading, b) and nu=2")
        al = [float(a) for a in c.line_search_alphas]
        if not (1 <= len(al) <= _lib.MAX_ALPHAS):
            raise ValueError(f"between 1 and {_lib.MAX_ALPHAS} line-search step sizes are supported")
        o.horizon, o.max_iter, o.n_alpha = int(c.horizon), int(c.max_iter), len(al)
        o.angle_wrap = int(self.angle_wrap)
        o.tol, o.reg = float(c.tol), float(c.reg)
        for i, a in enumerate(al):
            o.alphas[i] = a
        o.eps_dyn, o.eps_jac = float(self.eps_dyn), float(self.eps_jac)
        o.tight_in_jac = int(self.tight_in_jac)
        return o


@dataclass(frozen=True)
class SystemSpec:
    """Declarative counterpart of core/system_spec.py:10-31: the callables f / u_project / h of the reference
    are replaced by the data they are built from."""
    name: str
    nx: int
    nu: int
    dt: float
    horizon_N: int
    task_horizon_H: int
    dubins: DubinsConfig
    ctrl: BoxClampControl
    obstacles: Tuple[CircleObstacle, ...]
    obs_beta: float = 20.0
    obs_agg: str = "smoothmin"   # "smoothmin" | "min" | "single"
    y_target: Tuple[float, float, float] = (10.0, 10.0, math.pi / 4)

    def to_c(self) -> _lib.Env:
        e = _lib.Env()
        e.dt = float(self.dt)
        lo, hi = self.ctrl.bounds()
        e.u_min[0], e.u_min[1] = lo
        e.u_max[0], e.u_max[1] = hi
        e.active_tol = float(self.ctrl.active_tol)
        if len(self.obstacles) > _lib.MAX_OBS:
            raise ValueError(f"at most {_lib.MAX_OBS} obstacles")
        e.n_obs = len(self.obstacles)
        e.obs_agg = {"smoothmin": _lib.AGG_SMOOTHMIN, "min": _lib.AGG_MIN, "single": _lib.AGG_SINGLE}[self.obs_agg]
        for i, o in enumerate(self.obstacles):
            e.obs_cx[i], e.obs_cy[i], e.obs_r[i] = float(o.center[0]), float(o.center[1]), float(o.radius)
        e.obs_beta = float(self.obs_beta)
        return e


def system_from_config(cfg: Dict[str, Any]) -> SystemSpec:
    """core/tube_mpc.py:56-88 / :680-721."""
    s = cfg["system"]
    if s["name"] != "dubins":
        raise NotImplementedError("Only dubins is wired in the skeleton; other systems are added next.")
    v_max = float(s["control_bounds"]["v_max"])
    om = float(s["control_bounds"]["omega_max"])
    v_min = float(s["control_bounds"].get("v_min", -v_max))
    dub = DubinsConfig(dt=float(s["dt"]), v_max=v_max, omega_max=om,
                       w_low=tuple(s["disturbance"]["w_low"]), w_high=tuple(s["disturbance"]["w_high"]),
                       x_target=tuple(s["target"]))
    env = cfg.get("environment", {})
    agg = str(env.get("obstacle_aggregation", "min"))
    if "obstacles" in env:
        obs = tuple(CircleObstacle(center=tuple(o["center"]), radius=float(o["radius"])) for o in env["obstacles"])
        agg = "smoothmin" if agg == "smoothmin" else "min"
    else:
        o = env.get("obstacle", {"center": [5.0, 5.0], "radius": 1.5})
        obs = (CircleObstacle(center=tuple(o["center"]), radius=float(o["radius"])),)
        agg = "single"
    return SystemSpec(name="dubins", nx=3, nu=2, dt=float(s["dt"]), horizon_N=int(s["horizon_N"]),
                      task_horizon_H=int(s["task_horizon_H"]), dubins=dub,
                      ctrl=BoxClampControl(u_min=[v_min, -om], u_max=[v_max, om]), obstacles=obs,
                      obs_beta=float(env.get("obstacle_smoothmin_beta", 20.0)), obs_agg=agg,
                      y_target=tuple(float(v) for v in s["target"]))


def is_paper_mode(cfg: Dict[str, Any]) -> bool:
    """Mode dispatch of core/tube_mpc.py:48 (App. B Q1)."""
    return bool(cfg.get("paper_dubins_mode", True)) and (not bool(cfg.get("adaptation", {}).get("adapt_nominal", True)))


@dataclass
class TubeProblem:
    """Everything the two-layer solve needs besides the per-instance data."""
    system: SystemSpec
    nominal: OcpSpec
    auxiliary: OcpSpec
    mode: str                      # "paper" | "full"
    theta_nom0: List[float] = field(default_factory=list)   # paper: constrained values; full: RAW values
    theta_aux0: List[float] = field(default_factory=list)

    @staticmethod
    def from_config(cfg: Dict[str, Any], mode: Optional[str] = None) -> "TubeProblem":
        mode = mode or ("paper" if is_paper_mode(cfg) else "full")
        sysd = system_from_config(cfg)
        s = cfg["system"]
        N = int(s["horizon_N"])
        eps_yaml = float(cfg["dbas"].get("eps", DBaSConfig().eps))
        cn = cfg["cost_nominal"]
        if mode == "paper":
            # core/tube_mpc.py:706-768: alpha = gamma = 0, YAML eps everywhere, default reg, tol 1e-3, YAML alphas
            alphas = tuple(float(a) for a in s.get("line_search_alphas", [1.0]))
            mk = lambda key: OcpSpec(ILQRConfig(horizon=N, max_iter=int(s.get(key, 10)), tol=1e-3,
                                                line_search_alphas=alphas), eps_dyn=eps_yaml, eps_jac=eps_yaml)
            ca = cfg.get("cost_auxiliary", {})
            Qa = [float(v) for v in ca.get("Q", [1.0, 1.0, 1.0])]
            Ra = [float(v) for v in ca.get("R", [1.0, 1.0])]
            qba = float(ca.get("q_b", 1.0))
            th_n = [float(v) for v in cn["Q"]] + [float(v) for v in cn["R"]] + [float(v) for v in cn["Qf"]] + [
                float(cn["q_b"]), 0.0, 0.0, 0.0]
            th_a = Qa + Ra + Qa + [qba, 0.0, 0.0, 0.0]   # aux terminal weight = Qa (App. B Q6)
        elif mode == "full":
            # core/tube_mpc.py:105-165: YAML values are RAW; reg = ilqr_reg; default tol / alphas; solve-time eps 1e-6
            reg = float(s.get("ilqr_reg", 1e-6))
            mk = lambda key: OcpSpec(ILQRConfig(horizon=N, max_iter=int(s.get(key, 10)), reg=reg),
                                     eps_dyn=DBaSConfig().eps, eps_jac=eps_yaml)
            a0, g0 = float(cfg["dbas"]["alpha"]), float(cfg["dbas"]["gamma"])
            t0 = float(cfg["dbas"].get("nominal_tightening", 0.0))
            base = [float(v) for v in cn["Q"]] + [float(v) for v in cn["R"]] + [float(v) for v in cn["Qf"]]
            qba = float(cfg.get("cost_auxiliary", {}).get("q_b", float(cn["q_b"])))
            th_n = base + [float(cn["q_b"]), a0, g0, t0]
            th_a = base + [qba, a0, g0]
        else:
            raise ValueError(f"unknown mode {mode!r}")
        return TubeProblem(system=sysd, nominal=mk("nominal_max_iter"), auxiliary=mk("aux_max_iter"), mode=mode,
                           theta_nom0=th_n, theta_aux0=th_a)

    def with_horizon(self, T: int) -> "TubeProblem":
        return replace(self, nominal=replace(self.nominal, ilqr=replace(self.nominal.ilqr, horizon=T)),
                       auxiliary=replace(self.auxiliary, ilqr=replace(self.auxiliary.ilqr, horizon=T)))

    # -- full mode helpers -------------------------------------------------
    def make_thetas(self, device, dtype) -> Tuple[NominalTheta, AuxiliaryTheta]:
        """Raw parameter containers initialised as core/tube_mpc.py:105-132 does."""
        t = lambda v: torch.tensor(v, device=device, dtype=dtype)
        n, a = self.theta_nom0, self.theta_aux0
        th_n = NominalTheta(t(n[0:3]), t(n[3:5]), t(n[5:8]), t(n[8]), t(n[9]), t(n[10]), t(n[11]))
        th_a = AuxiliaryTheta(t(a[0:3]), t(a[3:5]), t(a[5:8]), t(a[8]), t(a[9]), t(a[10]))
        return th_n, th_a
