"""mjlab_b200: a B200-native batched rigid-body physics step behind mjlab's Simulation API.

Only the hot path is here: model compiler -> C-ABI CUDA engine -> torch views.
"""

from pathlib import Path

PKG_PATH = Path(__file__).parent
REPO_PATH = PKG_PATH.parent
__version__ = "0.1.0"
