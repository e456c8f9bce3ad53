"""Root-level module with the reference's name (quantile_forecaster.py); the implementation lives in the package."""
from shared_imports import *  # noqa: F401,F403
from neural_inventory_control_b200.quantile_forecaster import *  # noqa: F401,F403
from neural_inventory_control_b200.quantile_forecaster import FullyConnectedForecaster  # noqa: F401
