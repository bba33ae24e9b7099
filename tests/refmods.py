"""Imports the REFERENCE's own Python modules (staged under baseline/_ref by tools/stage_reference.py; /root/reference in
the dev container) on top of abr_iod_b200.compat -- shared by the drop-in tests.  Test infrastructure."""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.path.join(ROOT, "baseline", "_ref"), os.environ.get("ABR_REFERENCE", "/root/reference")]


def reference_root():
    for c in CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "maskrcnn_benchmark", "layers", "roi_align.py")):
            return c
    return None


class ReferenceModules:
    """Context manager: `maskrcnn_benchmark` importable from the staged sources with `apex` stubbed and
    `maskrcnn_benchmark._C` provided by compat.install(); everything is removed from sys.modules on exit."""

    def __enter__(self):
        import abr_iod_b200.compat as compat

        self.root = reference_root()
        self.saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("maskrcnn_benchmark", "apex", "tools")}
        for k in self.saved:
            del sys.modules[k]
        sys.path.insert(0, self.root)
        apex, amp = types.ModuleType("apex"), types.ModuleType("apex.amp")
        amp.float_function = lambda f: f  # the ops below are called with fp32 (layers/roi_align.py:58, nms.py:8)
        apex.amp = amp
        sys.modules["apex"], sys.modules["apex.amp"] = apex, amp
        self.compat = compat
        compat.install()
        return self

    def load_by_path(self, rel, name):
        """A module whose package cannot be imported on Python 3.12 (maskrcnn_benchmark.data needs `imp`)."""
        spec = importlib.util.spec_from_file_location(name, os.path.join(self.root, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    def __exit__(self, *exc):
        from abr_iod_b200.structures import bounding_box

        bounding_box.OUTPUT_CLASS = None
        sys.path.remove(self.root)
        for k in [k for k in sys.modules if k.split(".")[0] in ("maskrcnn_benchmark", "apex", "tools") or k.startswith("ref_")]:
            del sys.modules[k]
        sys.modules.update(self.saved)
        return False
