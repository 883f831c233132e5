"""Scenario TOML — mirrors the reference's serde structs (pedoni-simulator/src/scenario.rs:10-66).

The TOML format is kept verbatim (north star: "keeps the simulator's scenario-TOML"); unknown keys
are ignored like serde does (random.toml:3 carries a `field.unit` nobody reads).
"""
from __future__ import annotations

import tomllib
import dataclasses as dc
from dataclasses import dataclass
from pathlib import Path
from typing import List, Tuple

Vec2 = Tuple[float, float]


@dataclass
class FieldConfig:  # scenario.rs:18-20
    size: Vec2 = (0.0, 0.0)


@dataclass
class ObstacleConfig:  # scenario.rs:23-27, width default 1.0 (:4-6)
    line: Tuple[Vec2, Vec2] = ((0.0, 0.0), (0.0, 0.0))
    width: float = 1.0


@dataclass
class WaypointConfig:  # scenario.rs:39-43
    line: Tuple[Vec2, Vec2] = ((0.0, 0.0), (0.0, 0.0))
    width: float = 1.0


@dataclass
class PedestrianSpawnConfig:  # scenario.rs:63-66, tag = "kind"
    kind: str = "periodic"  # "periodic" | "once"
    frequency: float = 0.0  # periodic: pedestrians per second (f64)
    count: int = 0          # once: i32


@dataclass
class PedestrianConfig:  # scenario.rs:55-59
    origin: int = 0
    destination: int = 0
    spawn: PedestrianSpawnConfig = dc.field(default_factory=PedestrianSpawnConfig)


@dataclass
class Scenario:  # scenario.rs:10-15
    field: FieldConfig = dc.field(default_factory=FieldConfig)
    waypoints: List[WaypointConfig] = dc.field(default_factory=list)
    obstacles: List[ObstacleConfig] = dc.field(default_factory=list)
    pedestrians: List[PedestrianConfig] = dc.field(default_factory=list)

    @staticmethod
    def from_toml_str(text: str) -> "Scenario":
        doc = tomllib.loads(text)

        def vec2(v) -> Vec2:
            if len(v) != 2:
                raise ValueError(f"expected [x, y], got {v!r}")
            return (float(v[0]), float(v[1]))

        def line(v):
            if len(v) != 2:
                raise ValueError(f"expected [[x, y], [x, y]], got {v!r}")
            return (vec2(v[0]), vec2(v[1]))

        sc = Scenario(field=FieldConfig(size=vec2(doc["field"]["size"])))
        # serde: `waypoints`, `obstacles`, `pedestrians` are required Vec fields (no #[serde(default)]).
        for w in doc["waypoints"]:
            sc.waypoints.append(WaypointConfig(line=line(w["line"]), width=float(w.get("width", 1.0))))
        for o in doc["obstacles"]:
            sc.obstacles.append(ObstacleConfig(line=line(o["line"]), width=float(o.get("width", 1.0))))
        for p in doc["pedestrians"]:
            sp = p["spawn"]
            kind = sp["kind"]
            if kind == "periodic":
                spawn = PedestrianSpawnConfig(kind=kind, frequency=float(sp["frequency"]))
            elif kind == "once":
                spawn = PedestrianSpawnConfig(kind=kind, count=int(sp["count"]))
            else:
                raise ValueError(f"unknown spawn kind {kind!r}")
            sc.pedestrians.append(PedestrianConfig(origin=int(p["origin"]), destination=int(p["destination"]),
                                                   spawn=spawn))
        return sc

    @staticmethod
    def from_toml(path) -> "Scenario":
        return Scenario.from_toml_str(Path(path).read_text())
