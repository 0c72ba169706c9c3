"""Importing this package puts psdr_cuda_b200/compat on sys.path so that `import psdr_cuda` resolves to the B200-native
module (the reference's import name, src/psdr.cpp:41)."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
