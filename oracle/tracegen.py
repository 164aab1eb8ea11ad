"""Synthetic trace generators live in rlgpuschedule_b200/synth.py (shared by bench.py and the tests)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlgpuschedule_b200.synth import *  # noqa: F401,F403,E402
from rlgpuschedule_b200.synth import frame_gen, frame_probe100, frame_rows, write, COLS  # noqa: F401,E402
