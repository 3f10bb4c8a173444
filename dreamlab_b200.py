"""Import shim: the package directory is named after the reference repo
(`stable-diffusion-1.5-lcm-onnx-rknn2_b200/`), which is not a valid Python identifier.
`import dreamlab_b200` loads that directory as the package `dreamlab_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "stable-diffusion-1.5-lcm-onnx-rknn2_b200")
_spec = importlib.util.spec_from_file_location(
    "dreamlab_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dreamlab_b200"] = _mod
_spec.loader.exec_module(_mod)
