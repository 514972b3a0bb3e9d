"""spectral_b200: B200-native planning hot path (corridor generation + Bezier QP) of Srujan-D/spectral."""
from .wire import Scenario, ScenarioBatch, read_scenario_text, write_scenario_text, read_trajectory_text  # noqa: F401
