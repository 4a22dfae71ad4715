"""IntegrationScheme -- mirrors toss/trajectory/Integrator.py:4-51 (names are desolver method names).
The B200 path implements the one scheme every cfg and test of the reference selects
(RK8713MSolver = 3, default_cfg.toml:12); the other values are kept so configs parse and are
rejected at integration time (`supported`)."""
from enum import Enum


class IntegrationScheme(Enum):
    RK1412Solver = 1
    RK108Solver = 2
    RK8713MSolver = 3
    RK45CKSolver = 4
    HeunEulerSolver = 5
    DOPRI45 = 6
    BABs9o7HSolver = 7
    ABAs5o6HSolver = 8
    RK5Solver = 9
    RK4Solver = 10
    MidpointSolver = 11
    HeunsSolver = 12
    EulerSolver = 13
    EulerTrapSolver = 14
    SymplecticEulerSolver = 15
    LobattoIIIC4 = 16
    RadauIIA5 = 17
    GaussLegendre4 = 18
    GaussLegendre6 = 19
    BackwardEuler = 20
    ImplicitMidpoint = 21
    LobattoIIIA2 = 22
    LobattoIIIA4 = 23
    LobattoIIIB2 = 24
    LobattoIIIB4 = 25
    LobattoIIIC2 = 26
    CrankNicolson = 27
    RadauIA3 = 28
    RadauIA5 = 29
    RadauIIA3 = 30
    RadauIIA19 = 31

    @property
    def supported(self) -> bool:
        return self is IntegrationScheme.RK8713MSolver
