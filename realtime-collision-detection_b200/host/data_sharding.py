"""Shard-manager adapter: the routing half of src/collision/data_sharding.py.

The reference's ``ShardManager`` (:22-588) is control plane -- broker topics, node registry, heartbeats -- and out of
scope; what the hot path needs from it is the answer to "which shard owns this vehicle"
(``get_shard_for_vehicle``, :172-201) and a place where measured loads arrive (``update_node_load``, :153-170) and
rebalancing is triggered (:513-516).  This class keeps exactly that, with the same method names, arguments and
bookkeeping (``shards[shard]['vehicle_count' | 'load' | 'node_id']``, ``vehicle_to_shard``), for both modes of
``SpatialPartitioner``:

* region mode: the reference's behaviour -- a vehicle keeps the shard it was first given (sticky), and when the
  partitioner knows no shard for the position a random one is drawn (:190-192);
* slab mode (GPUs of one box): every shard is the x-slab of one GPU and a slab must hold the objects that lie in it,
  because its neighbours' halos are cut by position (host/slabs.py).  Ownership therefore follows the position:
  a vehicle that crossed a cut *migrates* (the counts of both shards move with it; ``migrations`` counts them).
"""
from __future__ import annotations

import random
import time
from typing import Any, Dict, Optional, Sequence

from .models import Position
from .spatial_index import SpatialPartitioner


class ShardManager:
    def __init__(self, spatial_partitioner: SpatialPartitioner, node_id: str = "node-0", initial_shards: int = 10,
                 rebalance_interval: float = 60.0, rng: Optional[random.Random] = None):
        self.node_id = node_id
        self.spatial_partitioner = spatial_partitioner
        self.rebalance_interval = rebalance_interval
        self.shards: Dict[str, Dict[str, Any]] = {
            f"shard-{k}": {"node_id": None, "vehicle_count": 0, "load": 0.0, "created_at": time.time()}
            for k in range(initial_shards)}
        self.vehicle_to_shard: Dict[str, str] = {}
        self.migrations = 0
        self._rng = rng or random

    @property
    def slab_mode(self) -> bool:
        return self.spatial_partitioner.slab_cuts is not None

    def get_shard_for_vehicle(self, vehicle_id: str, position: Position) -> str:
        """data_sharding.py:172-201."""
        known = self.vehicle_to_shard.get(vehicle_id)
        if known is not None and not self.slab_mode:
            return known  # sticky
        shard = self.spatial_partitioner.get_shard_for_position(position)
        if not shard or shard not in self.shards:
            if known is not None:
                return known
            shard = self._rng.choice(list(self.shards.keys()))
        if known == shard:
            return shard
        if known is not None:  # slab mode: the vehicle crossed a cut
            self.shards[known]["vehicle_count"] -= 1
            self.migrations += 1
        self.vehicle_to_shard[vehicle_id] = shard
        self.shards[shard]["vehicle_count"] += 1
        return shard

    def remove_vehicle(self, vehicle_id: str) -> None:
        shard = self.vehicle_to_shard.pop(vehicle_id, None)
        if shard in self.shards:
            self.shards[shard]["vehicle_count"] -= 1

    def assign_shard(self, shard_id: str, node_id: str) -> bool:
        if shard_id not in self.shards:
            return False
        self.shards[shard_id]["node_id"] = node_id
        return True

    def get_node_for_shard(self, shard_id: str) -> Optional[str]:
        return self.shards[shard_id]["node_id"] if shard_id in self.shards else None

    def get_node_for_vehicle(self, vehicle_id: str, position: Position) -> Optional[str]:
        return self.get_node_for_shard(self.get_shard_for_vehicle(vehicle_id, position))

    def update_shard_loads(self, loads: Sequence[float]) -> None:
        """Slab mode: measured frame time (ms) of every GPU, in slab order -- the GPUs' answer to the CPU-usage
        heartbeats of :153-170.  Feeds the partitioner."""
        for k, load in enumerate(loads):
            sid = f"shard-{k}"
            if sid in self.shards:
                self.shards[sid]["load"] = float(load)
                self.spatial_partitioner.update_load(sid, float(load))

    def check_rebalance(self) -> bool:
        """:513-516 -- let the partitioner re-balance when its interval is over."""
        return self.spatial_partitioner.check_rebalance()

    def get_stats(self) -> Dict[str, Any]:
        return {"total_shards": len(self.shards), "total_vehicles": len(self.vehicle_to_shard), "migrations": self.migrations,
                "shards": {s: {"vehicle_count": v["vehicle_count"], "load": v["load"], "node_id": v["node_id"]}
                           for s, v in self.shards.items()}}
