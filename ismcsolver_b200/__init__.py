"""ismcsolver_b200 — B200-native ISMCTS search engine (hot path of sfranzen/ismcsolver).

The product is the CUDA library built from csrc/ (C ABI: include/ismcts_b200.h) and the C++
host layer in include/ismcts/ that mirrors the reference's SOSolver / MOSolver API.  This Python
package only builds and loads the library for the test and benchmark harnesses.
"""
from .build import LIB_PATH, build_library  # noqa: F401

__version__ = "0.1"
