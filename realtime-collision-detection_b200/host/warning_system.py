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

    def get_stats(self):
        return dict(self.stats)
