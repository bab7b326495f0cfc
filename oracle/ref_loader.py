"""Import the reference's own loss / bank classes (TEST / BENCH INFRASTRUCTURE ONLY - never on the product path).

Looks in `baseline/_ref/` (staged by oracle/make_ref.py; this is what exists on the GPU box) and then in
/root/reference (dev container).  `faiss` / `wandb` are stubbed: they are only pulled in by utils/eval_utils.py:2 and
models/*.py:3 at import time and nothing on the loss path calls them.
Returns None when no copy of the reference is available (callers then fall back to the oracle port and say so).
"""
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CACHE = {}


def find_root():
    for cand in (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("SSV_REFERENCE", "/root/reference")):
        if cand and os.path.isfile(os.path.join(cand, "utils", "losses.py")):
            return cand
    return None


def load():
    """-> namespace with .losses (utils/losses.py), .MemoryBank (models/moco.py:23), .FeatureBank / .Prototypes
    (models/swav.py:57 / :44), .root; or None."""
    if "ns" in _CACHE:
        return _CACHE["ns"]
    root = find_root()
    ns = None
    if root is not None:
        for stub in ("faiss", "wandb"):
            if stub not in sys.modules:
                try:
                    importlib.import_module(stub)
                except Exception:  # noqa: BLE001
                    sys.modules[stub] = types.ModuleType(stub)
        sys.path.insert(0, root)
        try:
            ns = types.SimpleNamespace(root=root)
            ns.losses = importlib.import_module("utils.losses")
            try:
                ns.MemoryBank = importlib.import_module("models.moco").MemoryBank
                sw = importlib.import_module("models.swav")
                ns.FeatureBank, ns.Prototypes = sw.FeatureBank, sw.Prototypes
            except Exception as e:  # noqa: BLE001  (model files import torchvision etc.; the losses alone suffice)
                ns.MemoryBank = ns.FeatureBank = ns.Prototypes = None
                ns.bank_import_error = repr(e)
        finally:
            sys.path.remove(root)
    _CACHE["ns"] = ns
    return ns
