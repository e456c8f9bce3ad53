"""Root-level module with the reference's name, so `from trainer import *` / user notebooks keep working.
The implementation lives in neural_inventory_control_b200.trainer (star-import chain as in the reference)."""
from shared_imports import *  # noqa: F401,F403
from environment import *  # noqa: F401,F403
from loss_functions import *  # noqa: F401,F403
from neural_inventory_control_b200.trainer import *  # noqa: F401,F403
