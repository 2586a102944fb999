"""B200-native particle occupancy-map update engine (the per-frame hot path of tud-amr/semantic_dsp_map).

``SemanticDSPMap`` mirrors the reference's class of the same name over ``libdspmap_b200.so`` (C ABI in
``include/dspmap_c.h``); the CUDA library must be built (``__graft_entry__.build()``) and a GPU must be present.
"""
from .configs import BASELINE, SETTINGS, MapConfig, tiny
from .labels import LabelTables, ZED2_ROWS
from .map import Engine, MaskKpts, OccupancyResult, SemanticDSPMap, connect_shards
from .synth import Frame, Scene, make_random_table
from . import wire  # noqa: E402,F401

__all__ = ["BASELINE", "SETTINGS", "MapConfig", "tiny", "LabelTables", "ZED2_ROWS", "Engine", "MaskKpts", "OccupancyResult",
           "SemanticDSPMap", "connect_shards", "Frame", "Scene", "make_random_table"]
