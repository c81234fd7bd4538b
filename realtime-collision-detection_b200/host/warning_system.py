"""Classification part of src/collision/warning_system.py: the alert gate,
priority tiers and message tiers.  The priority of every emitted pair is computed on the GPU
together with the pair (rcd_pair.priority); ``AlertManager`` here only turns risks into
``AlertInfo`` records.  Queueing, resend loops and broker plumbing are out of scope (I/O glue).
"""
from __future__ import annotations

import time
import uuid
from typing import Dict, List, Optional, Tuple

from .models import AlertInfo, CollisionRisk

RISK_LEVEL_LOW = 0.3
RISK_LEVEL_MEDIUM = 0.6
RISK_LEVEL_HIGH = 0.8
PRIORITY_LOW, PRIORITY_MEDIUM, PRIORITY_HIGH, PRIORITY_CRITICAL = 0, 1, 2, 3


class AlertManager:
    def __init__(self, broker=None):
        self.broker = broker
        self.alerts: Dict[Tuple[str, str], AlertInfo] = {}
        self.stats = {"total_alerts": 0, "active_alerts": 0}

    def _get_priority(self, risk_level: float, time_to_collision: float) -> int:
        """warning_system.py:287-311 (the same thresholds the kernels apply to every pair)."""
        if risk_level >= RISK_LEVEL_HIGH and time_to_collision < 3.0:
            return PRIORITY_CRITICAL
        if risk_level >= RISK_LEVEL_HIGH or time_to_collision < 5.0:
            return PRIORITY_HIGH
        if risk_level >= RISK_LEVEL_MEDIUM:
            return PRIORITY_MEDIUM
        return PRIORITY_LOW

    def _generate_alert_message(self, risk: CollisionRisk) -> str:
        """Message tiers of warning_system.py:313-329 (English wording)."""
        if risk.risk_level >= RISK_LEVEL_HIGH:
            return (f"URGENT: possible collision with vehicle {risk.other_vehicle_id} in "
                    f"{risk.time_to_collision:.1f} s, distance {risk.distance:.1f} m - take evasive action now")
        if risk.risk_level >= RISK_LEVEL_MEDIUM:
            return (f"WARNING: possible collision with vehicle {risk.other_vehicle_id} in "
                    f"{risk.time_to_collision:.1f} s, distance {risk.distance:.1f} m - prepare to give way")
        return f"NOTICE: vehicle {risk.other_vehicle_id} is close ({risk.distance:.1f} m) - keep a safe distance"

    def process_collision_risks(self, risks: List[CollisionRisk]) -> List[AlertInfo]:
        """warning_system.py:259-285: drop risks below RISK_LEVEL_LOW, create or update one alert
        per (vehicle, other vehicle)."""
        out = []
        for risk in risks:
            if risk.risk_level < RISK_LEVEL_LOW:
                continue
            prio = getattr(risk, "alert_priority", -1)
            if prio < 0:  # a risk that did not come from the GPU path
                prio = self._get_priority(risk.risk_level, risk.time_to_collision)
            key = (risk.vehicle_id, risk.other_vehicle_id)
            alert = self.alerts.get(key)
            if alert is None:
                alert = AlertInfo(id=f"alert-{uuid.uuid4()}", vehicle_id=risk.vehicle_id,
                                  other_vehicle_id=risk.other_vehicle_id, risk_level=risk.risk_level,
                                  time_to_collision=risk.time_to_collision,
                                  message=self._generate_alert_message(risk), priority=prio, timestamp=time.time())
                self.alerts[key] = alert
                self.stats["total_alerts"] += 1
            else:
                alert.risk_level, alert.time_to_collision = risk.risk_level, risk.time_to_collision
                alert.priority, alert.message, alert.timestamp = prio, self._generate_alert_message(risk), time.time()
            out.append(alert)
        self.stats["active_alerts"] = len(self.alerts)
        return out

    # -- device table (csrc/rcd_alerts.cuh): the same lifecycle for whole frames ----------------------
    def attach_engine(self, engine, id_of, max_alerts: int = 1 << 20) -> None:
        """Keep the alert table on the GPU next to the frames of ``engine`` (a ``FrameEngine``);
        ``id_of(k)`` maps the caller ids of the pairs back to vehicle id strings."""
        self._engine, self._id_of = engine, id_of
        engine.alerts_configure(max_alerts)

    def _alert_from_event(self, e) -> AlertInfo:
        vid, other = self._id_of(int(e["i"])), self._id_of(int(e["j"]))
        risk = CollisionRisk(id="", vehicle_id=vid, other_vehicle_id=other, time_to_collision=float(e["ttc"]), distance=float("nan"),
                             relative_speed=0.0, risk_level=float(e["risk"]), collision_position=None, timestamp=float(e["timestamp"]))
        msg = self._generate_alert_message(risk) if float(e["risk"]) >= RISK_LEVEL_MEDIUM else \
            f"NOTICE: vehicle {other} is close - keep a safe distance"
        return AlertInfo(id=f"alert-{int(e['alert_id'])}", vehicle_id=vid, other_vehicle_id=other, risk_level=float(e["risk"]),
                         time_to_collision=float(e["ttc"]), message=msg, priority=int(e["priority"]),
                         timestamp=float(e["timestamp"]), acknowledged=bool(e["acknowledged"]))

    def process_frame(self, now: Optional[float] = None, report_refreshed: bool = False) -> List[AlertInfo]:
        """process_collision_risks (:259-285) for every risk of the engine's last frame, on the device.
        Returns only what the reference's queue would hear about: created alerts and priority changes
        (plus refreshed ones on request)."""
        ev, st = self._engine.alerts_update(time.time() if now is None else now, report_refreshed)
        self.stats["total_alerts"] += st["n_created"]
        self.stats["active_alerts"] = st["n_live"]
        return [self._alert_from_event(e) for e in ev]

    def cleanup_expired(self, now: Optional[float] = None, max_age: float = 30.0) -> List[Tuple[str, str]]:
        """_cleanup_expired_alerts (:488-517); returns the (vehicle, other) keys that were dropped."""
        ev, st = self._engine.alerts_expire(time.time() if now is None else now, max_age)
        self.stats["active_alerts"] = st["n_live"]
        return [(self._id_of(int(e["i"])), self._id_of(int(e["j"]))) for e in ev]

    def acknowledge(self, vehicle_id_index: int, other_vehicle_index: int) -> bool:
        """acknowledge_alert (:199-213) addressed by the pair's caller ids."""
        return self._engine.alerts_acknowledge([vehicle_id_index], [other_vehicle_index]) == 1

    def get_stats(self):
        return dict(self.stats)
