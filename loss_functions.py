"""Root-level module with the reference's name, so `from trainer import *` / user notebooks keep working.
The implementation lives in neural_inventory_control_b200.loss_functions (star-import chain as in the reference)."""
from shared_imports import *  # noqa: F401,F403
from neural_inventory_control_b200.loss_functions import *  # noqa: F401,F403
