"""av1b200: B200-native AV1 encode backend (drop-in for the av1an + SVT-AV1 stage of IONIQ6000/av1-base)."""
from . import abi  # noqa: F401
