"""rtb200: B200-native (sm_100a CUDA) transport kernels for the razoumov/radiativeTransfer hot path."""
from ._lib import LIB_PATH, RTB200Error, lib  # noqa: F401
from .solver import MATH_FAITHFUL, MATH_FAST, Transport, comm_unique_id, direction, patterns  # noqa: F401
