"""B200-native batched NAO controller-and-simulator (drop-in for the closed-loop hot path of
Ema158/linearMpcHumanoid).  The compute path is liblmh_b200.so (hand-written sm_100a CUDA behind the
C ABI in include/lmh_capi.h); this package is the Python binding plus the scenario generators."""
from .api import (Batch, LmhError, load_library, build_library, measure_fp64_peak, export_mujoco,  # noqa: F401
                  STAGE_DTYPE, STEPLOG_DTYPE, STATS_DTYPE, SYMBOLS, LIB_PATH,
                  TAU_LIMIT, COM_ERR_LIMIT, STATE_LIMIT, TORSO_MIN)
from .scenarios import push_plan, shard_range, walk_plan, foot_coefficients  # noqa: F401
