from .setup_parameters import GravityEvaluable, setup_parameters
from .setup_state import setup_initial_state_domain
from .udp_initial_condition import udp_initial_condition

__all__ = ["GravityEvaluable", "setup_parameters", "setup_initial_state_domain", "udp_initial_condition"]
