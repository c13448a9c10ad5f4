"""Epoch time of the persistent head kernel at the paper shapes (sessions 1, 4, 8)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "subspace-reg_b200"))
sys.path.insert(0, ROOT)
import profile_sweep  # noqa: E402

profile_sweep.head_timing()
