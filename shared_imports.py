"""Root-level module with the reference's name, so `from trainer import *` / user notebooks keep working.
The implementation lives in neural_inventory_control_b200.shared_imports (star-import chain as in the reference)."""
from neural_inventory_control_b200.shared_imports import *  # noqa: F401,F403
