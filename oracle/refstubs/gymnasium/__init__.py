"""Inert stand-in for `gymnasium`, which is not installed in this image.

Only used by tests/golden/make_golden.py to import the *unmodified* reference from
/root/reference (its environment.py does `import gymnasium as gym; from gymnasium import spaces`
and only uses `gym.Env`, `spaces.Box`, `spaces.Dict` as passive containers).
TEST INFRASTRUCTURE ONLY - never imported by the product package.
"""
from . import spaces  # noqa: F401


class Env:
    metadata = {}
