"""SimulationConfig and friends -- mirror of src/config.rs:10-371 (the fields the hot path reads).

JSON layout and defaults follow config/default.json and the ``Default`` impls
(src/config.rs:26-35, 56-66, 93-103, 305-316).
"""
from __future__ import annotations

import json
from dataclasses import asdict, dataclass, field


class ConfigError(Exception):
    """src/config.rs:375-388"""


@dataclass
class BubbleParameters:  # src/config.rs:10-35
    diameter: float = 0.05
    film_thickness_nm: float = 500.0
    critical_thickness_nm: float = 10.0
    refractive_index: float = 1.33


@dataclass
class FluidParameters:  # src/config.rs:38-66
    viscosity: float = 0.001
    surface_tension: float = 0.025
    density: float = 1000.0
    surfactant_diffusion: float = 1e-9
    surfactant_concentration: float = 0.5


@dataclass
class EnvironmentParameters:  # src/config.rs:69-103
    gravity: float = 9.81
    temperature: float = 293.15
    pressure: float = 101325.0
    wind: list = field(default_factory=lambda: [0.0, 0.0, 0.0])
    buoyancy: float = 1.0


@dataclass
class SimulationConfig:  # src/config.rs:283-316
    bubble: BubbleParameters = field(default_factory=BubbleParameters)
    fluid: FluidParameters = field(default_factory=FluidParameters)
    environment: EnvironmentParameters = field(default_factory=EnvironmentParameters)
    foam: dict = field(default_factory=dict)  # rigid-body foam dynamics are out of scope; kept opaque
    dt: float = 0.001
    resolution: int = 128

    @classmethod
    def default(cls) -> "SimulationConfig":
        return cls()

    @classmethod
    def from_file(cls, path) -> "SimulationConfig":  # src/config.rs:328-337
        try:
            with open(path) as f:
                raw = json.load(f)
        except OSError as e:
            raise ConfigError(f"Failed to read config file '{path}': {e}") from e
        except json.JSONDecodeError as e:
            raise ConfigError(f"Failed to parse config file '{path}': {e}") from e
        try:
            return cls(bubble=BubbleParameters(**raw["bubble"]), fluid=FluidParameters(**raw["fluid"]),
                       environment=EnvironmentParameters(**raw["environment"]), foam=raw.get("foam", {}),
                       dt=float(raw["dt"]), resolution=int(raw["resolution"]))
        except (KeyError, TypeError) as e:
            raise ConfigError(f"Failed to parse config file '{path}': {e}") from e

    def to_file(self, path) -> None:  # src/config.rs:347-354
        with open(path, "w") as f:
            json.dump(asdict(self), f, indent=2)

    def bubble_radius(self) -> float:  # src/config.rs:357-359
        return self.bubble.diameter / 2.0

    def internal_pressure_difference(self) -> float:  # src/config.rs:363-365
        return 4.0 * self.fluid.surface_tension / self.bubble_radius()

    def film_thickness_meters(self) -> float:  # src/config.rs:368-370
        return self.bubble.film_thickness_nm * 1e-9
