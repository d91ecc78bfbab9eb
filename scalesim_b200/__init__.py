"""scalesim_b200 — B200-native Time Warp core behind ScaleSim's application surface.

The product is ``libscalesim_b200.so`` (CUDA kernels for sm_100a + the C ABI of
``include/scalesim_b200.h``). This package is the thin Python host layer over that
C ABI (ctypes): PyTorch is used only for plumbing (pinned host buffers,
``torch.distributed`` rendezvous). There is no CPU fallback anywhere: creating an
engine without the built library or without a CUDA device raises.
"""
from .capi import (  # noqa: F401
    Engine,
    EngineError,
    Config,
    Stats,
    EVENT_DTYPE,
    RECORD_DTYPE,
    RAW_RECORD_DTYPE,
    FLAG_KERNEL_TIMING,
    FLAG_HISTORY,
    SENT_RECORD_DTYPE,
    PACKED_RECORD_DTYPE,
    events_to_raw,
    packed_to_records,
    raw_to_records,
    MODEL_PHOLD,
    MODEL_TRAFFIC,
    MODEL_PHOLD_STATEFUL,
    library_path,
    load_library,
    records_to_lines,
    digest_records,
)
from . import dist, phold, scenario, traffic  # noqa: F401

__all__ = [
    "Engine", "EngineError", "Config", "Stats", "EVENT_DTYPE", "RECORD_DTYPE", "RAW_RECORD_DTYPE", "FLAG_KERNEL_TIMING", "FLAG_HISTORY", "SENT_RECORD_DTYPE", "PACKED_RECORD_DTYPE", "events_to_raw", "packed_to_records", "raw_to_records",
    "MODEL_PHOLD", "MODEL_TRAFFIC", "MODEL_PHOLD_STATEFUL", "library_path", "load_library",
    "records_to_lines", "digest_records", "dist", "phold", "scenario", "traffic",
]
