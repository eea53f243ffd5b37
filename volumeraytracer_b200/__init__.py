"""volumeraytracer_b200 -- B200-native (sm_100a) replacement for the ray-marching hot path of
PaulStahr/VolumeRaytracer, behind the reference's own TraceRaysCu<> boundary.  See DESIGN.md."""
from ._lib import (VRT_F32, VRT_I16, VRT_U32, VRT_OPT_KERNEL, VRT_OPT_BLOCK_THREADS, VRT_OPT_REFILL,   # noqa: F401
                   VRT_OPT_CHUNK_RAYS, VRT_OPT_STEPS_PER_POLL, VRT_OPT_MAX_CTAS_PER_SM, VRT_INFO_EMPTY_PERMILLE, VRT_INFO_NUM_SMS, VRT_INFO_STAT_BASE, VRT_TRACE_ROUND_HOST, VRT_OPT_WAVE_LOG2, VRT_OPT_WAVE_MARGIN, VRT_OPT_WAVE_CHECK, VRT_OPT_WAVE_TAIL_PERMILLE, VRT_OPT_WAVE_CTAS_PER_SM, VRT_OPT_WAVE_REFILL, VRT_OPT_ALL_CLEAR_KERNEL, VRT_INFO_ALL_CLEAR, VRT_INFO_WAVE_ROUNDS, VRT_OPT_REGION_LOG2, VRT_OPT_REGION_ROUNDS, VrtError, launch_count, lib, LIB_PATH)
from .scene import Comm, Options, RaytraceScene, TraceRaysCu                                                # noqa: F401
