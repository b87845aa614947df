"""B200-native hot path of the multimodal pre-routing timing predictor.

The directory doubles as the drop-in replacement for the reference's ``src/`` directory: put it on
``sys.path`` and ``import model`` / ``import Unet`` resolve to the CUDA-backed modules.  Importing
the package (``importlib.import_module("multimodal-fusion-based-pre-routing-timing-prediction-_b200")``)
does exactly that and re-exports the pieces.
"""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
if _here not in _sys.path:
    _sys.path.insert(0, _here)

import tm_lib  # noqa: E402,F401
import tm_synth  # noqa: E402,F401
import tm_graph  # noqa: E402,F401
import tm_ops  # noqa: E402,F401
import tm_unet  # noqa: E402,F401
import model  # noqa: E402,F401
import Unet  # noqa: E402,F401
import tm_engine  # noqa: E402,F401
from model import MLP, PathConv, LayoutNet, PathModel  # noqa: E402,F401
from Unet import UNet  # noqa: E402,F401
from tm_graph import TimingGraph, MaskCSR, MaskRows, Schedule  # noqa: E402,F401
from tm_ops import MaskedFeatureMap  # noqa: E402,F401
from tm_engine import DesignBatch, DesignStep, HostDesign, PreparedDesign, build_models  # noqa: E402,F401
