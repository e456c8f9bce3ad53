"""Star-import surface of the reference's shared_imports.py (torch, nn, np, pd, Dataset, DataLoader, plt, copy,
DefaultDict, datetime, os). matplotlib is optional: plotting is outside the hot path."""
import copy  # noqa: F401
import datetime  # noqa: F401
import os  # noqa: F401
from collections import defaultdict as DefaultDict  # noqa: F401

import numpy as np  # noqa: F401
import pandas as pd  # noqa: F401
import torch  # noqa: F401
from torch import nn  # noqa: F401
from torch.nn.modules.loss import _Loss  # noqa: F401
from torch.utils.data import DataLoader, Dataset  # noqa: F401

try:  # pragma: no cover - absent in the build image
    import matplotlib.pyplot as plt  # noqa: F401
except Exception:  # noqa: BLE001
    plt = None
