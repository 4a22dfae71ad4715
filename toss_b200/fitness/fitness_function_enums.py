"""FitnessFunctions -- mirrors toss/fitness/fitness_function_enums.py:4-48 (values are the `fn` id of the
C ABI's toss_fitness_eval)."""
from enum import Enum


class FitnessFunctions(Enum):
    TargetAltitudeDistance = 1
    CloseDistancePenalty = 2
    FarDistancePenalty = 3
    CoveredVolume = 4
    TotalCoveredVolume = 5
    CoveredVolumeFarDistancePenalty = 6
    CoveredVolumeCloseDistancePenaltyFarDistancePenalty = 7
    CoveredSpace = 8
    CoveredSpaceCloseDistancePenaltyFarDistancePenalty = 9
