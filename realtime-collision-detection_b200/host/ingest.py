"""Batched ingest of the reference's vehicle wire format.

The reference handles one broker message at a time: ``json.loads`` -> dict ->
``EarlyWarningSystem._handle_vehicle_position`` builds a ``Vehicle`` and calls
``CollisionDetector.update_vehicle`` + ``CollisionPredictionModel.update_trajectory``
(src/collision/warning_system.py:638-678; producer src/test/vehicle_simulator.py:721-752).
``VehicleIngest`` takes whole buffers of those messages instead: the native decoder
(csrc/rcd_ingest.hpp) turns them into fixed records and interns ids / types, and
``FrameEngine.apply_records`` scatters the records into the device-resident frame state and
trajectory rings (csrc/rcd_ingest.cuh).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Tuple, Union

import numpy as np

from . import _native as N


class VehicleIngest:
    """Decoder + id table.  Slot ``k`` of the frame is the k-th distinct vehicle id ever seen."""

    def __init__(self, threads: int = 0):
        self._lib = N.load()
        self._g = ctypes.c_void_p()
        self.threads = int(threads)
        rc = self._lib.rcd_ingest_create(ctypes.byref(self._g))
        if rc != N.RCD_OK:
            raise N.NativeError(rc, "rcd_ingest_create")
        self.bad_messages = 0

    def close(self) -> None:
        if getattr(self, "_g", None) is not None and self._g:
            self._lib.rcd_ingest_destroy(self._g)
            self._g = ctypes.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc != N.RCD_OK:
            msg = self._lib.rcd_ingest_last_error(self._g)
            raise N.NativeError(rc, msg.decode("utf-8", "replace") if msg else "")

    # -- decode -----------------------------------------------------------------------------
    def decode(self, data: Union[bytes, bytearray, memoryview, str], out: Optional[np.ndarray] = None
               ) -> Tuple[np.ndarray, int]:
        """Decode every message in ``data``; returns (records, max_seq).  Malformed / incomplete
        messages are dropped and counted in ``bad_messages`` (the reference logs and drops them)."""
        if isinstance(data, str):
            data = data.encode("utf-8")
        buf = (ctypes.c_char * len(data)).from_buffer_copy(data) if not isinstance(data, bytes) else data
        n_bytes = len(data)
        # a message of the reference format is > 150 bytes; 64 is a safe lower bound per record
        cap = max(16, n_bytes // 64 + 1) if out is None else int(out.shape[0])
        rec = np.empty(cap, dtype=N.RECORD_DTYPE) if out is None else out
        n, bad, mseq = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint32()
        ptr = ctypes.cast(buf, ctypes.c_void_p) if isinstance(buf, bytes) else ctypes.cast(ctypes.addressof(buf), ctypes.c_void_p)
        rc = self._lib.rcd_ingest_decode_json(self._g, ptr, n_bytes, self.threads, rec.ctypes.data_as(ctypes.c_void_p), cap,
                                              ctypes.byref(n), ctypes.byref(bad), ctypes.byref(mseq))
        if rc == N.RCD_ECAPACITY and out is None and int(n.value) > cap:
            # more (tiny) messages than the estimate: nothing was decoded and no state changed -- once more, with room
            cap = int(n.value)
            rec = np.empty(cap, dtype=N.RECORD_DTYPE)
            rc = self._lib.rcd_ingest_decode_json(self._g, ptr, n_bytes, self.threads, rec.ctypes.data_as(ctypes.c_void_p),
                                                  cap, ctypes.byref(n), ctypes.byref(bad), ctypes.byref(mseq))
        self._check(rc)
        self.bad_messages += int(bad.value)
        return rec[: int(n.value)], int(mseq.value)

    # -- id / type tables -------------------------------------------------------------------
    def set_limit(self, max_ids: int) -> None:
        """At most ``max_ids`` distinct vehicles: messages of further (unknown) vehicles are dropped and counted
        (``bad_messages``, ``rejected``) while the known ones keep being served."""
        self._check(self._lib.rcd_ingest_set_limit(self._g, int(max_ids)))

    @property
    def rejected(self) -> int:
        n = ctypes.c_uint64()
        self._check(self._lib.rcd_ingest_rejected(self._g, ctypes.byref(n)))
        return int(n.value)

    @property
    def n_objects(self) -> int:
        n = ctypes.c_uint64()
        self._check(self._lib.rcd_ingest_counts(self._g, ctypes.byref(n), None))
        return int(n.value)

    @property
    def n_types(self) -> int:
        n = ctypes.c_uint64()
        self._check(self._lib.rcd_ingest_counts(self._g, None, ctypes.byref(n)))
        return int(n.value)

    def id_of(self, slot: int) -> str:
        p, ln = ctypes.c_char_p(), ctypes.c_uint32()
        self._check(self._lib.rcd_ingest_id_name(self._g, int(slot), ctypes.byref(p), ctypes.byref(ln)))
        return ctypes.string_at(p, ln.value).decode("utf-8", "surrogatepass")

    def type_of(self, code: int) -> str:
        p, ln = ctypes.c_char_p(), ctypes.c_uint32()
        self._check(self._lib.rcd_ingest_type_name(self._g, int(code), ctypes.byref(p), ctypes.byref(ln)))
        return ctypes.string_at(p, ln.value).decode("utf-8", "surrogatepass")

    def slot_of(self, vehicle_id: str) -> Optional[int]:
        b = vehicle_id.encode("utf-8", "surrogatepass")
        s = ctypes.c_uint32()
        rc = self._lib.rcd_ingest_lookup(self._g, b, len(b), ctypes.byref(s))
        if rc == N.RCD_ESTATE:
            return None
        self._check(rc)
        return int(s.value)

    def ids(self) -> List[str]:
        return [self.id_of(k) for k in range(self.n_objects)]


class VehiclePositionStream:
    """Batched counterpart of ``EarlyWarningSystem._handle_vehicle_position`` +
    ``_detect_all_vehicles`` (src/collision/warning_system.py:638-714): message buffers in,
    one GPU frame for every vehicle out.  Object state and trajectory rings stay on the device;
    the host keeps only the id table."""

    def __init__(self, max_objects: int, max_pairs: Optional[int] = None, world_bounds=None, device: int = 0,
                 threads: int = 0, max_history: int = 100):
        from .engine import FrameEngine
        self.engine = FrameEngine(max_objects, max_pairs, device=device, world_bounds=world_bounds)
        self.engine.history_configure(max_history)
        self.ingest = VehicleIngest(threads=threads)
        self.ingest.set_limit(max_objects)  # a full frame rejects new vehicles, it does not fail every later batch
        self.messages_applied = 0

    def close(self) -> None:
        self.ingest.close()
        self.engine.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def handle_messages(self, data) -> int:
        """Decode and apply a buffer of vehicle messages; returns how many were applied."""
        rec, max_seq = self.ingest.decode(data)
        if len(rec):
            self.engine.apply_records(rec, self.ingest.n_objects, max_seq, history=True)
            self.messages_applied += len(rec)
        return len(rec)

    def detect_all_vehicles(self, predict: bool = True) -> np.ndarray:
        """``predict_collisions`` (or ``detect_collisions``) for every vehicle: pairs with i / j = slots
        (``ingest.id_of``), sorted by (i, j)."""
        if self.engine.n == 0:
            return np.zeros(0, dtype=N.PAIR_DTYPE)
        if predict:
            self.engine.history_classify(want_codes=False)
            return self.engine.predict()
        return self.engine.detect()
