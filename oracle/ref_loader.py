"""Execute the reference's own function bodies verbatim (oracle; test infrastructure).

The reference modules cannot be imported (they import mediapipe / matplotlib at module
top and ``rppg_*.py`` run ``main()`` at import), but every
function on the signal path is plain NumPy/SciPy/cv2.  This loader parses a reference
file with ``ast``, keeps the requested ``FunctionDef`` nodes and the module-level
constant assignments, and ``exec``s them -- unmodified -- into a namespace that provides
``np``, ``sp``, ``cv`` / ``cv2``.  No reference source is copied into this repository.

``/root/reference`` exists only in the build container: this module is used by
``tests/golden/make_golden.py`` (fixture generation) and by CPU tests that skip when the
reference tree is absent.  Nothing on the GPU box reads it.
"""
from __future__ import annotations

import ast
import os
from typing import Any

REFERENCE_ROOT = os.environ.get("VHR_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "rppg_VIDEO.py"))


def load_functions(rel_path: str, names, extra_ns: dict | None = None) -> dict[str, Any]:
    """Return {name: callable} for ``names`` defined at module level of
    ``REFERENCE_ROOT/rel_path``; module-level simple constant assignments (``FREQ_LOW =
    0.7`` ...) are executed too so the functions see their globals."""
    import numpy as np
    import scipy.signal as sp
    import cv2

    path = os.path.join(REFERENCE_ROOT, rel_path)
    with open(path, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)
    keep = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            keep.append(node)
        elif isinstance(node, ast.Assign) and all(isinstance(t, ast.Name) for t in node.targets):
            # constants only: literals and arithmetic on literals (FREQ_LOW = 40 / 60)
            try:
                ast.literal_eval(node.value)
                keep.append(node)
            except Exception:
                if isinstance(node.value, ast.BinOp) and all(
                        isinstance(n, (ast.BinOp, ast.Constant, ast.operator, ast.UnaryOp, ast.unaryop))
                        for n in ast.walk(node.value)):
                    keep.append(node)
    mod = ast.Module(body=keep, type_ignores=[])
    from collections import deque
    from typing import Optional, Tuple, List, Sequence, Generator
    ns: dict[str, Any] = {"np": np, "sp": sp, "cv": cv2, "cv2": cv2, "deque": deque,
                          "Optional": Optional, "Tuple": Tuple, "List": List,
                          "Sequence": Sequence, "Generator": Generator,
                          "Landmarks": List, "ENABLE_PLOTTING": False}
    if extra_ns:
        ns.update(extra_ns)
    exec(compile(mod, path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    if missing:
        raise KeyError(f"{rel_path}: functions not found: {missing}")
    return {n: ns[n] for n in names} | {"__ns__": ns}


class Landmark:
    """Stand-in for mediapipe's NormalizedLandmark (only .x / .y are read,
    rppg_VIDEO.py:93-94; analysis/utils/roi.py:44-45)."""
    __slots__ = ("x", "y")

    def __init__(self, x: float, y: float):
        self.x = float(x)
        self.y = float(y)
