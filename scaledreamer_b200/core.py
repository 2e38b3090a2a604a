"""Host-side shell that keeps threestudio's plugin contract: name -> class registry, nested dataclass Configs
parsed strictly from yaml (unknown keys are rejected), scheduled scalars `C()`, the recursive `Updateable` walk and
the BaseObject / BaseModule constructors. Mirrors (restated, not copied):
    threestudio/__init__.py:5-32          register / find (incl. "main:mixin1,mixin2")
    threestudio/utils/config.py:11-128    resolvers, load_config, parse_structured
    threestudio/utils/base.py:11-118      Configurable / Updateable / BaseObject / BaseModule
    threestudio/utils/misc.py:18-30,66-101  get_rank / get_device / C
OmegaConf and pytorch-lightning are not required: yaml goes through PyYAML and a small resolver.
"""
from __future__ import annotations

import dataclasses
import math
import os
import re
import typing
from dataclasses import dataclass, field, fields, is_dataclass
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn
import yaml

__modules__: Dict[str, type] = {}


def register(name: str):
    def decorator(cls):
        if name in __modules__:
            raise ValueError(f"Module {name} already exists! Names of extensions conflict!")
        __modules__[name] = cls
        return cls

    return decorator


def find(name: str):
    if ":" in name:
        main_name, sub_names = name.split(":")
        subs = [s.strip() for s in sub_names.split(",")] if "," in sub_names else [sub_names]
        cls = __modules__[main_name]
        for sub in subs:
            cls = getattr(cls, sub)
            if cls is None:
                raise AttributeError(f"mixin {sub} of {main_name} is None")
        return cls
    if name not in __modules__:
        raise KeyError(f"no plugin registered under '{name}' (known: {sorted(__modules__)})")
    return __modules__[name]


# ----------------------------------------------------------------------------------------------- environment
_SYNTH_WARNED = set()


def synthetic_or_raise(what: str, path: str):
    """Pretrained weights / cached text embeddings that cannot be found are an ERROR (the reference fails in
    `from_pretrained` / runs CLIP). Seeded synthetic stand-ins are used only when SDB_SYNTHETIC_WEIGHTS=1 is set
    (bench.py, smoke() and the tests set it: there are no checkpoints on the box), with one warning per kind."""
    if os.environ.get("SDB_SYNTHETIC_WEIGHTS") != "1":
        raise FileNotFoundError(f"{what}: '{path}' not found. Point the config at real weights / a populated "
                                ".threestudio_cache, or set SDB_SYNTHETIC_WEIGHTS=1 to run on seeded synthetic "
                                "parameters (benchmarks and tests only: training against them optimises noise).")
    if what not in _SYNTH_WARNED:
        _SYNTH_WARNED.add(what)
        import warnings

        warnings.warn(f"SDB_SYNTHETIC_WEIGHTS=1: {what} replaced by seeded synthetic values ('{path}' not found)",
                      RuntimeWarning, stacklevel=3)


def get_rank() -> int:
    for key in ("RANK", "LOCAL_RANK", "SLURM_PROCID", "JSM_NAMESPACE_RANK"):
        if os.environ.get(key) is not None:
            return int(os.environ[key])
    return 0


def get_device() -> torch.device:
    return torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}")


# ----------------------------------------------------------------------------------------------- config
def _resolver_table(n_gpus: int):
    def calc_exp_lr_decay_rate(factor, n):
        return float(factor) ** (1.0 / float(n))

    return {
        "calc_exp_lr_decay_rate": calc_exp_lr_decay_rate,
        "add": lambda a, b: a + b, "sub": lambda a, b: a - b, "mul": lambda a, b: a * b,
        "div": lambda a, b: a / b, "idiv": lambda a, b: a // b,
        "basename": lambda p: os.path.basename(p),
        "rmspace": lambda s, sub: str(s).replace(" ", sub),
        "tuple2": lambda s: [float(s), float(s)],
        "gt0": lambda s: s > 0,
        "cmaxgt0": lambda s: _cmax(s) > 0,
        "cmaxgt0orcmaxgt0": lambda a, b: _cmax(a) > 0 or _cmax(b) > 0,
        "not": lambda s: not s,
        "n_gpus": lambda: n_gpus,
    }


def _cmax(value):
    """C_max (threestudio/utils/config.py:31-49): the largest value a scheduled scalar reaches."""
    if isinstance(value, (int, float)):
        return value
    value = list(value)
    if len(value) >= 6:
        max_value = value[2]
        for i in range(4, len(value), 2):
            max_value = max(max_value, value[i])
        value = [value[0], value[1], max_value, value[3]]
    if len(value) == 3:
        value = [0] + value
    if len(value) != 4:
        raise TypeError(f"Scalar specification must have 3, 4 or >= 6 entries, got {value}")
    return max(value[1], value[2])


_INTERP = re.compile(r"\$\{([^${}]+)\}")


def _lookup(root, path: str):
    node = root
    for part in path.split("."):
        node = node[int(part)] if isinstance(node, list) else node[part]
    return node


def _split_args(s: str) -> List[str]:
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            depth += ch in "([{"
            depth -= ch in ")]}"
            cur += ch
    out.append(cur)
    return [a.strip() for a in out]


def _coerce(s: str):
    try:
        return yaml.safe_load(s)
    except Exception:
        return s


def _resolve_str(s: str, root, table, depth=0):
    if depth > 32:
        raise ValueError(f"interpolation too deep: {s}")
    whole = _INTERP.fullmatch(s)
    while True:
        m = _INTERP.search(s)
        if not m:
            return s
        expr = m.group(1).strip()
        if ":" in expr and expr.split(":", 1)[0] in table:
            name, args = expr.split(":", 1)
            vals = [_resolve_value(_coerce(a), root, table, depth + 1) for a in _split_args(args)] if args != "" else []
            val = table[name](*vals)
        else:
            val = _resolve_value(_lookup(root, expr), root, table, depth + 1)
        if whole and m.span() == (0, len(s)):
            return val
        s = s[: m.start()] + str(val) + s[m.end():]
        whole = None


def _resolve_value(v, root, table, depth=0):
    if isinstance(v, str) and "${" in v:
        return _resolve_str(v, root, table, depth)
    return v


def _resolve_tree(node, root, table):
    if isinstance(node, dict):
        for k in list(node):
            node[k] = _resolve_tree(node[k], root, table)
        return node
    if isinstance(node, list):
        return [_resolve_tree(v, root, table) for v in node]
    return _resolve_value(node, root, table)


def _deep_merge(a: dict, b: dict) -> dict:
    for k, v in b.items():
        if isinstance(v, dict) and isinstance(a.get(k), dict):
            _deep_merge(a[k], v)
        else:
            a[k] = v
    return a


def _set_dotted(cfg: dict, key: str, value) -> None:
    node = cfg
    parts = key.split(".")
    for p in parts[:-1]:
        node = node.setdefault(p, {})
    node[parts[-1]] = value


MISSING = "???"


@dataclass
class ExperimentConfig:
    name: str = "default"
    description: str = ""
    tag: str = ""
    seed: int = 0
    use_timestamp: bool = True
    timestamp: Optional[str] = None
    exp_root_dir: str = "outputs"
    exp_dir: str = "outputs/default"
    trial_name: str = "exp"
    trial_dir: str = "outputs/default/exp"
    n_gpus: int = 1
    resume: Optional[str] = None
    data_type: str = ""
    data: dict = field(default_factory=dict)
    system_type: str = ""
    system: dict = field(default_factory=dict)
    trainer: dict = field(default_factory=dict)
    checkpoint: dict = field(default_factory=dict)


def load_config(*yamls: str, cli_args: Optional[List[str]] = None, from_string=False, n_gpus: int = 1, **kwargs):
    """yaml files (later override earlier) + `a.b=c` overrides -> resolved ExperimentConfig."""
    cfg: dict = {}
    for y in yamls:
        loaded = yaml.safe_load(y if from_string else open(y).read()) or {}
        _deep_merge(cfg, loaded)
    for arg in cli_args or []:
        if "=" not in arg:
            raise ValueError(f"override '{arg}' is not of the form key=value")
        k, v = arg.split("=", 1)
        _set_dotted(cfg, k, _coerce(v))
    _deep_merge(cfg, kwargs)
    cfg = _resolve_tree(cfg, cfg, _resolver_table(n_gpus))
    scfg = parse_structured(ExperimentConfig, cfg)
    scfg.n_gpus = n_gpus
    if not scfg.tag and not scfg.use_timestamp:
        raise ValueError("Either tag is specified or use_timestamp is True.")
    scfg.trial_name = scfg.tag
    # threestudio/utils/config.py:86-101: a run without an explicit `timestamp` gets "@%Y%m%d-%H%M%S" appended to its trial
    # name (single-GPU runs only: ranks of one job must agree on the directory), so successive runs never share
    # ckpts/last.ckpt or save/, and "@LAST" has timestamped directories to choose from; the directory is created here
    if scfg.timestamp is None:
        scfg.timestamp = ""
        if scfg.use_timestamp and n_gpus <= 1:
            from datetime import datetime

            scfg.timestamp = datetime.now().strftime("@%Y%m%d-%H%M%S")
    scfg.trial_name += scfg.timestamp
    scfg.exp_dir = os.path.join(scfg.exp_root_dir, scfg.name)
    scfg.trial_dir = os.path.join(scfg.exp_dir, scfg.trial_name)
    if os.environ.get("SDB_NO_TRIAL_DIRS") != "1":  # the test-suite parses dozens of configs and wants no directories
        os.makedirs(scfg.trial_dir, exist_ok=True)
    return scfg


def _check_missing(node, path=""):
    if isinstance(node, dict):
        for k, v in node.items():
            _check_missing(v, f"{path}.{k}" if path else k)
    elif node == MISSING:
        raise ValueError(f"Missing mandatory value: {path}")


def parse_structured(config_cls, cfg: Optional[Any] = None):
    """Strict dataclass construction: unknown keys raise TypeError, '???' raises, nested dataclass fields recurse."""
    if cfg is None:
        cfg = {}
    if is_dataclass(cfg) and not isinstance(cfg, type):
        cfg = dataclasses.asdict(cfg)
    cfg = dict(cfg)
    known = {f.name: f for f in fields(config_cls)}
    unknown = [k for k in cfg if k not in known]
    if unknown:
        raise TypeError(f"{config_cls.__qualname__} got unexpected config keys {unknown}")
    hints = typing.get_type_hints(config_cls)
    kwargs = {}
    for k, v in cfg.items():
        if isinstance(v, str) and v == MISSING:
            raise ValueError(f"Missing mandatory value: {k}")
        t = hints.get(k)
        if is_dataclass(t) and isinstance(v, dict):
            v = parse_structured(t, v)
        kwargs[k] = v
    obj = config_cls(**kwargs)
    for f in fields(config_cls):
        if isinstance(getattr(obj, f.name), str) and getattr(obj, f.name) == MISSING:
            raise ValueError(f"Missing mandatory value: {f.name}")
    return obj


def C(value: Any, epoch: int, global_step: int, interpolation="linear") -> float:
    """Scheduled scalar: number, or [start_step, start_value, end_value, end_step] (3 values: start_step = 0;
    >= 6 values: piecewise). An int end_step schedules on global_step, a float one on epoch."""
    if isinstance(value, (int, float)):
        return value
    value = list(value)
    if len(value) == 3:
        value = [0] + value
    if len(value) >= 6:
        select_i = 3
        for i in range(3, len(value) - 2, 2):
            if global_step >= value[i]:
                select_i = i + 2
        if select_i != 3:
            start_value, start_step = value[select_i - 3], value[select_i - 2]
        else:
            start_step, start_value = value[:2]
        end_value, end_step = value[select_i - 1], value[select_i]
        value = [start_step, start_value, end_value, end_step]
    if len(value) != 4:
        raise TypeError(f"Scalar specification must have 3, 4 or >= 6 entries, got {value}")
    start_step, start_value, end_value, end_step = value
    current = global_step if isinstance(end_step, int) else epoch
    t = max(min(1.0, (current - start_step) / (end_step - start_step)), 0.0)
    if interpolation == "linear":
        return start_value + (end_value - start_value) * t
    if interpolation == "exp":
        return math.exp(math.log(start_value) * (1 - t) + math.log(end_value) * t)
    raise ValueError(f"Unknown interpolation method: {interpolation}, only support linear and exp")


# ----------------------------------------------------------------------------------------------- base classes
class Updateable:
    def do_update_step(self, epoch: int, global_step: int, on_load_weights: bool = False):
        for attr in self.__dir__():
            if attr.startswith("_"):
                continue
            try:
                module = getattr(self, attr)
            except Exception:
                continue
            if isinstance(module, Updateable):
                module.do_update_step(epoch, global_step, on_load_weights=on_load_weights)
        self.update_step(epoch, global_step, on_load_weights=on_load_weights)

    def do_update_step_end(self, epoch: int, global_step: int):
        for attr in self.__dir__():
            if attr.startswith("_"):
                continue
            try:
                module = getattr(self, attr)
            except Exception:
                continue
            if isinstance(module, Updateable):
                module.do_update_step_end(epoch, global_step)
        self.update_step_end(epoch, global_step)

    def update_step(self, epoch: int, global_step: int, on_load_weights: bool = False):
        pass

    def update_step_end(self, epoch: int, global_step: int):
        pass


class BaseObject(Updateable):
    @dataclass
    class Config:
        pass

    cfg: Config

    def __init__(self, cfg: Optional[dict] = None, *args, **kwargs) -> None:
        super().__init__()
        self.cfg = parse_structured(self.Config, cfg)
        self.device = get_device()
        self.configure(*args, **kwargs)

    def configure(self, *args, **kwargs) -> None:
        pass


class BaseModule(nn.Module, Updateable):
    @dataclass
    class Config:
        weights: Optional[str] = None

    cfg: Config

    def __init__(self, cfg: Optional[dict] = None, *args, **kwargs) -> None:
        super().__init__()
        self.cfg = parse_structured(self.Config, cfg)
        self.device = get_device()
        self.configure(*args, **kwargs)
        if self.cfg.weights is not None:
            weights_path, module_name = self.cfg.weights.split(":")
            state_dict, epoch, global_step = load_module_weights(weights_path, module_name=module_name)
            self.load_state_dict(state_dict)
            self.do_update_step(epoch, global_step, on_load_weights=True)
        self.register_buffer("_dummy", torch.zeros(0).float(), persistent=False)

    def configure(self, *args, **kwargs) -> None:
        pass


def find_last_path(path: Optional[str]) -> Optional[str]:
    """threestudio/utils/misc.py:143-161: `outputs/exp/prompt@LAST/ckpts/last.ckpt` -> the lexicographically last
    directory starting with `outputs/exp/prompt@` (trial directories carry a timestamp), spaces read as underscores."""
    if path is None or "LAST" not in path:
        return path
    path = path.replace(" ", "_")
    prefix, suffix = path.split("LAST", 1)
    base_dir = os.path.dirname(prefix)
    prefix = os.path.join(base_dir, os.path.split(prefix)[-1])
    candidates = sorted((os.path.join(base_dir, d) for d in os.listdir(base_dir)), reverse=True)
    candidates = [d for d in candidates if d.startswith(prefix)]
    if not candidates or not os.path.exists(candidates[0] + suffix):
        raise FileNotFoundError((candidates[0] if candidates else prefix) + suffix)
    return candidates[0] + suffix


def load_module_weights(path, module_name=None, ignore_modules=None, map_location="cpu"):
    """Selects `module_name.*` entries of a Lightning-style checkpoint (threestudio/utils/misc.py:33-63)."""
    if module_name is not None and ignore_modules is not None:
        raise ValueError("module_name and ignore_modules cannot be both set")
    # Lightning checkpoints carry more than tensors (hyper-parameters, callback state): the full unpickler, as the
    # reference (torch < 2.6 semantics) uses
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    sd = ckpt["state_dict"]
    out = {}
    if ignore_modules is not None:
        for k, v in sd.items():
            if not any(k.startswith(i + ".") for i in ignore_modules):
                out[k] = v
    elif module_name is not None:
        for k, v in sd.items():
            m = re.match(rf"^{module_name}\.(.*)$", k)
            if m:
                out[m.group(1)] = v
    else:
        out = dict(sd)
    return out, ckpt.get("epoch", 0), ckpt.get("global_step", 0)
