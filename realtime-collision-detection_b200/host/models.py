"""Data contract at the boundary: the dataclasses the reference's callers pass in and get back.

``Position``, ``Vector``, ``LocationData`` follow src/common/models.py:10-64 field for field.
``Vehicle`` and the 10-field ``CollisionRisk`` are the types src/collision/* constructs but the
reference's models.py lacks: Vehicle fields per
src/collision/warning_system.py:649-670, CollisionRisk per
src/collision/collision_detection.py:156-166 and :831-842.  ``ComputeNodeCollisionRisk`` is the
9-field record of src/common/models.py:108-136 that src/compute/compute_node.py:310-317 creates.
"""
from __future__ import annotations

import time
import uuid
from dataclasses import dataclass, field


@dataclass
class Position:
    x: float
    y: float
    z: float

    def distance_to(self, other: "Position") -> float:
        return ((self.x - other.x) ** 2 + (self.y - other.y) ** 2 + (self.z - other.z) ** 2) ** 0.5


@dataclass
class Vector:
    x: float
    y: float
    z: float

    def magnitude(self) -> float:
        return (self.x ** 2 + self.y ** 2 + self.z ** 2) ** 0.5

    def normalize(self) -> "Vector":
        mag = self.magnitude()
        if mag == 0:
            return Vector(0, 0, 0)
        return Vector(self.x / mag, self.y / mag, self.z / mag)


@dataclass
class LocationData:
    vehicle_id: str
    timestamp: float
    position: Position
    velocity: Vector
    heading: float
    vehicle_type: str

    @classmethod
    def create(cls, vehicle_id, position, velocity, heading, vehicle_type) -> "LocationData":
        return cls(vehicle_id=vehicle_id, timestamp=time.time(), position=position, velocity=velocity,
                   heading=heading, vehicle_type=vehicle_type)


@dataclass
class Vehicle:
    id: str
    position: Position
    velocity: Vector
    acceleration: Vector
    heading: float
    size: float
    type: str
    timestamp: float = 0.0


@dataclass
class CollisionRisk:
    """What CollisionDetector.detect_collisions / CollisionPredictionModel.predict_collisions return."""
    id: str
    vehicle_id: str
    other_vehicle_id: str
    time_to_collision: float
    distance: float
    relative_speed: float
    risk_level: float
    collision_position: Position
    timestamp: float
    is_predicted: bool = False
    # additions of this implementation (classified on the GPU with the pair)
    alert_priority: int = -1      # warning_system.py:287-311; -1 = below RISK_LEVEL_LOW
    time_to_closest: float = 0.0  # stage-2 values of the detect path (collision_detection.py:277-284)
    closest_distance: float = 0.0


@dataclass
class ComputeNodeCollisionRisk:
    risk_id: str
    timestamp: float
    vehicle_id1: str
    vehicle_id2: str
    risk_level: float
    estimated_collision_time: float
    position: Position
    relative_velocity: float
    time_to_collision: float

    @classmethod
    def create(cls, vehicle_id1, vehicle_id2, risk_level, estimated_collision_time, position,
               relative_velocity) -> "ComputeNodeCollisionRisk":
        now = time.time()
        return cls(risk_id=str(uuid.uuid4()), timestamp=now, vehicle_id1=vehicle_id1, vehicle_id2=vehicle_id2,
                   risk_level=risk_level, estimated_collision_time=estimated_collision_time, position=position,
                   relative_velocity=relative_velocity, time_to_collision=estimated_collision_time - now)


@dataclass
class AlertInfo:
    """src/collision/warning_system.py:30-45."""
    id: str
    vehicle_id: str
    other_vehicle_id: str
    risk_level: float
    time_to_collision: float
    message: str
    priority: int
    timestamp: float
    acknowledged: bool = False

    def __lt__(self, other):
        return (self.priority, -self.timestamp) > (other.priority, -other.timestamp)
