"""Run the child script of tests/test_gpu_zz_fold_draft.py in-process (one torch import) -- for a short gpurun call."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("fold_draft", os.path.join(ROOT, "tests", "test_gpu_zz_fold_draft.py"))
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
exec(compile(mod.CHILD % {"root": ROOT}, "fold_draft_child", "exec"))
