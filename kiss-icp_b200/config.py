"""KISSConfig mirror of python/kiss_icp/config/config.py:28-48 and config/parser.py:41-81
(plain dataclasses; same field names, same defaults, same voxel_size = max_range/100 rule)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional


@dataclass
class DataConfig:
    max_range: float = 100.0
    min_range: float = 0.0
    deskew: bool = True


@dataclass
class MappingConfig:
    voxel_size: Optional[float] = None  # default: take it from data
    max_points_per_voxel: int = 20


@dataclass
class RegistrationConfig:
    max_num_iterations: int = 500
    convergence_criterion: float = 0.0001
    max_num_threads: int = 0  # accepted, ignored on the GPU


@dataclass
class AdaptiveThresholdConfig:
    fixed_threshold: Optional[float] = None
    initial_threshold: float = 2.0
    min_motion_th: float = 0.1


@dataclass
class KISSConfig:
    out_dir: str = "results"
    data: DataConfig = field(default_factory=DataConfig)
    registration: RegistrationConfig = field(default_factory=RegistrationConfig)
    mapping: MappingConfig = field(default_factory=MappingConfig)
    adaptive_threshold: AdaptiveThresholdConfig = field(default_factory=AdaptiveThresholdConfig)


def load_config(max_range: Optional[float] = None, deskew: Optional[bool] = None, voxel_size: Optional[float] = None,
                **overrides) -> KISSConfig:
    """parser.py:67-81: CLI overrides, min/max sanity, voxel_size = max_range / 100 when unset."""
    config = KISSConfig()
    if max_range is not None:
        config.data.max_range = max_range
    if deskew is not None:
        config.data.deskew = deskew
    if voxel_size is not None:
        config.mapping.voxel_size = voxel_size
    for dotted, value in overrides.items():
        section, key = dotted.split("__")
        setattr(getattr(config, section), key, value)
    if config.data.max_range < config.data.min_range:
        config.data.min_range = 0.0
    if config.mapping.voxel_size is None:
        config.mapping.voxel_size = float(config.data.max_range / 100.0)
    return config
